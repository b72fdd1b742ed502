"""Summarise an ncu source page: top stalled SASS instructions of one launch.
usage: python tools/ncu_src.py <rep> <launch-index> [top]"""
import csv, subprocess, sys, io
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
sec = rows[starts[idx]:starts[idx + 1]]
print(sec[0][:2])
hdr = sec[1]
data = [r for r in sec[2:] if len(r) == len(hdr)]
ci = {n: i for i, n in enumerate(hdr)}
S = ci["# Samples"]
tot = sum(int(r[S]) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[ci[n]] or 0) for r in data) for n in stalls}
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
order = sorted(range(len(data)), key=lambda i: -int(data[i][S]))[:top]
for i in sorted(order):
    r = data[i]
    why = sorted(((int(r[ci[n]] or 0), n[6:]) for n in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[S]):6d} {100*int(r[S])/max(tot,1):5.1f}%  {r[ci['Source']].strip()[:90]:90s} {why}")

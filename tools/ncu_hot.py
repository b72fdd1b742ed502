"""Top stall locations of every kernel in an ncu report (needs --import-source on): python tools/ncu_hot.py <rep> [top N]
Prints per launch: duration, the stall-reason mix over all warp samples, and the N SASS instructions with the most samples."""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr = rr[0]
names = [(r[0], r[hdr.index("Kernel Name")], r[hdr.index("gpu__time_duration.sum")]) for r in rr[2:]]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(src)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
if len(blocks) == 2 * len(names): blocks = blocks[::2]      # ncu prints two views per launch
for b, (kid, kname, dur) in zip(blocks, names):
    h = b["hdr"]; ix = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    S = [int(r[ix["# Samples"]]) for r in b["rows"]]
    tot = sum(S) or 1
    agg = sorted(((sum(int(r[ix[s]]) for r in b["rows"]), s) for s in stalls), reverse=True)[:6]
    print(f"==== id {kid} {kname[:70]}  {float(dur):.1f} us  samples {tot}")
    print("   " + "  ".join(f"{s[6:]} {100*v/tot:.0f}%" for v, s in agg))
    for i in sorted(sorted(range(len(S)), key=lambda i: -S[i])[:topn]):
        r = b["rows"][i]
        st = sorted(((int(r[ix[s]]), s) for s in stalls), reverse=True)[:2]
        print(f"   {i:5d} {r[ix['Source']].strip()[:58]:58s} {S[i]:6d} {100*S[i]/tot:4.1f}% x{r[ix['Instructions Executed']]:>8s}  " + " ".join(f"{s[6:]}={v}" for v, s in st))

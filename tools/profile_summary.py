"""Aggregate an `ncu --metrics ... --csv` launch list of bench.py into per-kernel shares of ONE step.
usage: python tools/profile_summary.py gpurun_out/launches_metrics.csv [profiles/conv_dram_traffic.json] > profiles/<name>.txt
With a second argument the conv_tc_* family's DRAM bytes per step are also written as JSON, together with the commit the
pass was taken at -- bench.py reports that figure as roofline.traffic instead of a constant in its source."""
import collections, csv, json, subprocess, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[h]
ki, mi, vi, ii, ui = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID'), hdr.index('Metric Unit')
per, order = {}, []
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    k = int(r[ii])
    if k not in per:
        per[k] = {'name': r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[-52:]}
        order.append(k)
    v = float(r[vi].replace(',', ''))
    if r[mi].startswith('dram__bytes'):
        v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[r[ui]]
    per[k][r[mi]] = v
seq = [per[k] for k in order]
idx = [i for i, d in enumerate(seq) if 'prep' in d['name']]
step = seq[idx[0]:idx[1]]
agg = collections.OrderedDict()
T = 'gpu__time_duration.sum'
for d in step:
    a = agg.setdefault(d['name'], [0.0, 0, 0.0, 0.0, 0.0])
    a[0] += d[T] / 1e3; a[1] += 1; a[2] += d.get('dram__bytes_read.sum', 0); a[3] += d.get('dram__bytes_write.sum', 0)
    a[4] += d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0) * d[T] / 1e3
tot = sum(a[0] for a in agg.values())
print(f"one bench step (64 tiles, YOLOv8m), {len(step)} kernel launches, {tot:.1f} us summed under ncu (cold, serialised)")
print(f"{'time us':>9} {'n':>4} {'share':>6} {'dram rd MB':>11} {'dram wr MB':>11} {'tensor act':>10}  kernel")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{a[0]:9.1f} {a[1]:4d} {100 * a[0] / tot:5.1f}% {a[2] / 1e6:11.1f} {a[3] / 1e6:11.1f} {a[4] / a[0]:9.1f}%  {n}")
conv = [a for n, a in agg.items() if n.startswith('conv_tc')]
print(f"conv_tc_* family: {sum(a[0] for a in conv):.1f} us ({100 * sum(a[0] for a in conv) / tot:.1f}% of the step), {sum(a[1] for a in conv)} launches, "
      f"dram read {sum(a[2] for a in conv) / 1e9:.3f} GB + write {sum(a[3] for a in conv) / 1e9:.3f} GB = {sum(a[2] + a[3] for a in conv) / 1e9:.3f} GB per step")
if len(sys.argv) > 2:
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
        dirty = bool(subprocess.run(["git", "status", "--porcelain", "--", "aerial_image_recognition_b200/csrc"], stdout=subprocess.PIPE, text=True).stdout.strip())
    except Exception:
        commit, dirty = "unknown", True
    json.dump({"dram_bytes_per_step": sum(a[2] + a[3] for a in conv), "dram_read_bytes": sum(a[2] for a in conv), "dram_write_bytes": sum(a[3] for a in conv),
               "launches": sum(a[1] for a in conv), "commit": commit + ("+uncommitted csrc changes" if dirty else ""),
               "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active... --clock-control none over `python bench.py --steps 2 --warmup 1 --legs ''`, one step (between two preprocess launches); " + sys.argv[1]},
              open(sys.argv[2], "w"), indent=1)

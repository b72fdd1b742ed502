"""Debug: single-rank vs emulated multi-rank dedup on a wide mosaic strip; timing of the seam path's phases."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from aerial_image_recognition_b200 import mosaic as M, synth
from aerial_image_recognition_b200.engine import Engine
from oracle import postproc as OP
GT = (2335637.62, 0.1, 0.0, 6845688.78, 0.0, -0.1)
H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 5120, 40000
eng = Engine("yolov8m", max_batch=64)
pool = torch.from_numpy(synth.mosaic_block_pool(77)).cuda()
band = synth.mosaic_band_device(pool, H, W, 0, H, 5)
det = M.MosaicDetector(eng, GT, conf=0.4, dedup_thr=1.0)
wins, ids, _ = M.shard_windows(H, W, 0, 1)
cols = det.detect_windows(band, wins, ids, 0)
x, y, conf, cls, wid, slot, py = cols
print("raw", x.numel())
single = det.dedup(*cols, 0, 1, None)
key = lambda a: a["window"].astype(np.int64) * 65536 + a["slot"]
ks = np.sort(key(single))
# CPU oracle on the same raw detections, in total order
okey = (wid.cpu().numpy().astype(np.int64) * 65536 + slot.cpu().numpy())
order = np.argsort(okey)
xs, ys, cs = x.cpu().numpy()[order], y.cpu().numpy()[order], conf.cpu().numpy()[order]
t0 = time.time(); keep = OP.dedup_greedy(xs, ys, cs, 1.0, True); print("oracle s", time.time() - t0)
ko = np.sort(okey[order][keep])
print("single vs oracle:", len(ks), len(ko), "equal" if np.array_equal(ks, ko) else f"DIFF {len(np.setxor1d(ks, ko))}")
for world in ((2,) if H > 20000 else (2, 3, 4)):
    covers = [M.shard_windows(H, W, r, world)[2] for r in range(world)]
    locals_, recs = [], []
    tt = {}
    for r in range(world):
        w_r, ids_r, _ = M.shard_windows(H, W, r, world)
        sel = torch.from_numpy(np.isin(wid.cpu().numpy(), ids_r)).cuda()
        sub = tuple(c[sel] for c in cols)
        torch.cuda.synchronize(); t0 = time.time()
        loc, rec = det.seam_split(*sub, r, covers)
        torch.cuda.synchronize(); tt[f"split{r}"] = time.time() - t0
        locals_.append(loc); recs.append(rec)
    origin = np.concatenate([np.full(len(rc), r, np.int64) for r, rc in enumerate(recs)])
    t0 = time.time()
    merged = [np.concatenate([locals_[r], det.seam_merge(recs, origin, r)]) for r in range(world)]
    tt["merge_all"] = time.time() - t0
    ku = np.sort(key(np.concatenate(merged)))
    d = np.setxor1d(ku, ks)
    print(f"world {world}: union {len(ku)} single {len(ks)} seam {sum(len(r) for r in recs)} diff {len(d)}", {k: round(v, 3) for k, v in tt.items()})
    for k in d[:6]:
        i = int(np.nonzero(okey == k)[0][0])
        xi, yi = x[i].item(), y[i].item()
        d2 = (x - xi) ** 2 + (y - yi) ** 2
        nb = torch.nonzero(d2 <= 1.0000001).flatten().cpu().numpy()
        print("   key", k, "in_union", k in ku, "in_single", k in ks, "py", py[i].item(), "conf", conf[i].item(), "neighbours:",
              [(int(okey[j]), float(conf[j]), float(py[j]), float(d2[j])) for j in nb if j != i][:5])

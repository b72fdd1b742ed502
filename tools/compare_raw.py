import sys, numpy as np
a = np.load(sys.argv[1]); parts = [np.load(f) for f in sys.argv[2:]]
def table(z):
    k = z["wid"].astype(np.int64) * 65536 + z["slot"]
    o = np.argsort(k)
    return k[o], z["x"][o], z["y"][o], z["conf"][o], z["py"][o]
ka, xa, ya, ca, pa = table(a)
cat = {f: np.concatenate([p[f] for p in parts]) for f in ("x", "y", "conf", "wid", "slot", "py")}
kb, xb, yb, cb, pb = table(cat)
print("counts", len(ka), len(kb), "keys equal", np.array_equal(ka, kb))
if np.array_equal(ka, kb):
    for n, u, v in (("x", xa, xb), ("y", ya, yb), ("conf", ca, cb), ("py", pa, pb)):
        d = np.nonzero(u != v)[0]
        print(n, "differs at", len(d), "rows", [(int(ka[i]) >> 16, int(ka[i]) & 65535, float(u[i]), float(v[i])) for i in d[:5]])
else:
    only_a, only_b = np.setdiff1d(ka, kb), np.setdiff1d(kb, ka)
    print("only in single", [(int(k) >> 16, int(k) & 65535) for k in only_a[:10]], "only in sharded", [(int(k) >> 16, int(k) & 65535) for k in only_b[:10]])

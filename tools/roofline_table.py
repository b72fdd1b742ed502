#!/usr/bin/env python
"""Per-op time against each op's own floor max(HBM bytes / peak, FLOP / peak) for a `tools/diag.py time` table.

    python tools/roofline_table.py profiles/r1_final2_per_op_times_b64.txt [--arch yolov8m --batch 64] > profiles/<name>.txt

Bytes are the op's algorithmic traffic (input slice + residual + output slice; weights are negligible at batch 64), FLOPs the
convolution's multiply-adds x 2.  Peaks from MEASURED_PEAKS.json (HBM copy bandwidth, sustained bf16)."""
import argparse
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aerial_image_recognition_b200 import graph as G   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("table")
    ap.add_argument("--arch", default="yolov8m")
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    hbm, tf = 6532.2e9, 1388.9e12
    pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        d = json.load(open(pk))
        hbm, tf = d.get("hbm_gbs", 6532.2) * 1e9, d.get("bf16_tflops_sustained", 1388.9) * 1e12
    t, desc = {}, {}
    for line in open(a.table):
        m = re.match(r"\[\s*(\d+)\]\s+([\d.]+) us\s+[\d.]+ TF/s\s+(.*)", line)
        if m:
            t[int(m.group(1))] = float(m.group(2))
            desc[int(m.group(1))] = m.group(3).split("|")[0].strip()
    g = G.build(a.arch)
    B = a.batch
    print(f"{a.arch} batch {B}: per-op time vs floor = max(bytes / {hbm / 1e9:.0f} GB/s, flop / {tf / 1e12:.0f} TF/s)   (source: {a.table})")
    print(f"{'op':>4} {'us':>8} {'floor':>8} {'x':>5} {'MB':>8} {'GFLOP':>8}  bound  what")
    tot = ftot = 0.0
    for i, op in enumerate(g.ops):
        if i not in t:
            continue
        sb, db = g.bufs[op.src.buf], g.bufs[op.dst.buf]
        if op.kind in ("conv", "dwconv"):
            cout, cing, k, _ = g.wshapes[op.weight]
            cin = cing if op.kind == "conv" else cout
            rd = B * sb.h * sb.w * cin * 2 + (B * db.h * db.w * cout * 2 if getattr(op, "res", None) is not None else 0)
            wr = B * db.h * db.w * cout * (4 if db.f32 else 2)
            fl = 2.0 * B * db.h * db.w * cout * cing * k * k
        else:
            c = op.dst.c
            rd, wr, fl = B * sb.h * sb.w * c * 2, B * db.h * db.w * c * 2, 0.0
        fb, ff = (rd + wr) / hbm * 1e6, fl / tf * 1e6
        fus = max(fb, ff)
        if t[i] < 5 and "fused into" in desc[i]:
            continue
        tot += t[i]
        ftot += fus
        print(f"{i:4d} {t[i]:8.1f} {fus:8.1f} {t[i] / fus:5.2f} {(rd + wr) / 1e6:8.1f} {fl / 1e9:8.1f}  {'hbm   ' if fb >= ff else 'tensor'} {desc[i][:90]}")
    print(f"sum  {tot:8.1f} {ftot:8.1f} {tot / ftot:5.2f}")


if __name__ == "__main__":
    main()

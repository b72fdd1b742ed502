"""GPU diagnostics: per-stage and per-op parity against the CPU oracle, and per-op timing.

Not a test and not the bench: a development tool run under gpurun, one section per
process so that a hung kernel in one section cannot hide the others:

    python tools/diag.py pre | tcops | forward | post | time [--arch yolov8m] [--batch N]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from aerial_image_recognition_b200 import graph as G, synth, weights as W  # noqa: E402
from aerial_image_recognition_b200.engine import Engine, dets_to_numpy, geodets_to_numpy  # noqa: E402


def log(*a):
    print(*a, flush=True)


def stats(name, got, ref):
    got = got.float().cpu() if torch.is_tensor(got) else torch.from_numpy(np.asarray(got)).float()
    ref = ref.float().cpu() if torch.is_tensor(ref) else torch.from_numpy(np.asarray(ref)).float()
    d = (got - ref).abs()
    rel = d.max().item() / (ref.abs().max().item() + 1e-12)
    log(f"  {name:44s} max|d| {d.max().item():.4e} mean|d| {d.mean().item():.3e} rel {rel:.3e} ref_rms {ref.pow(2).mean().sqrt().item():.3e}")
    return d.max().item(), rel


def sec_pre(args):
    import cv2
    from PIL import Image
    eng = Engine(args.arch, max_batch=2, imgsz=640, conv_impl="auto")
    rng = np.random.default_rng(0)
    cases = [rng.integers(0, 256, (864, 864, 3), dtype=np.uint8), rng.integers(0, 256, (1280, 1280, 3), dtype=np.uint8),
             rng.integers(0, 256, (1000, 1300, 3), dtype=np.uint8), synth.make_tiles(1, 640, 1)[0]]
    p = os.path.join(ROOT, "tests", "golden", "test_tile_864.png")
    if os.path.exists(p):
        cases.append(np.array(Image.open(p).convert("RGB")))
    for img in cases:
        t = torch.from_numpy(img)[None].cuda()
        h, w = img.shape[:2]
        if (h, w) == (640, 640):
            u8 = eng.preprocess(t, "identity", out="u8").cpu().numpy()[0]
            log(f"identity {h}x{w}: diff bytes {(u8 != img).sum()}")
            f = eng.preprocess(t, "identity", out="f32").cpu().numpy()[0]
            ref = (img.astype(np.float32) / 255.0).transpose(2, 0, 1)
            log(f"identity f32 bit-exact: {np.array_equal(f, ref)}")
            eng.preprocess(t, "identity")
            got = eng.buffer("input", 1).float().cpu().numpy()[0]
            ref_b = torch.from_numpy(img.astype(np.float32) / 255.0).to(torch.bfloat16).float().numpy()
            log(f"identity bf16 NHWC4 exact: {np.array_equal(got[..., :3], ref_b)} pad zero: {(got[..., 3] == 0).all()}")
            continue
        ref_pil = np.array(Image.fromarray(img).resize((640, 640)))
        ref_cv = cv2.resize(img, (640, 640))
        got_pil = eng.preprocess(t, "pil_bicubic", out="u8").cpu().numpy()[0]
        got_cv = eng.preprocess(t, "cv2_linear", out="u8").cpu().numpy()[0]
        log(f"{h}x{w}: PIL bicubic diff bytes {(got_pil != ref_pil).sum()}  cv2 linear diff bytes {(got_cv != ref_cv).sum()}")
        f = eng.preprocess(t, "pil_bicubic", out="f32").cpu().numpy()[0]
        log(f"   f32 CHW bit-exact vs PIL/255: {np.array_equal(f, (ref_pil.astype(np.float32) / 255.0).transpose(2, 0, 1))}")
        g = eng.preprocess(t, "cv2_linear", bgr=True, out="u8").cpu().numpy()[0]
        log(f"   bgr flag: {np.array_equal(g, ref_cv[..., ::-1])}")


def _cpu_op(g, op, w, bufs_cpu, emulate=True):
    import torch.nn.functional as F
    src = bufs_cpu[op.src.buf][..., op.src.c0:op.src.c0 + op.src.c].permute(0, 3, 1, 2).float()
    if op.kind in ("conv", "dwconv"):
        from aerial_image_recognition_b200.graph import op_weights
        wt, b = (torch.from_numpy(np.ascontiguousarray(a)) for a in op_weights(op, w))
        groups = g.wshapes[op.weight][3]
        if op.src.buf == "input":
            src = src[:, :3]
        y = F.conv2d(src, wt, b, stride=op.s, padding=op.k // 2, groups=groups)
        if op.act:
            y = y * torch.sigmoid(y)
        if op.res is not None:
            y = y + bufs_cpu[op.res.buf][..., op.res.c0:op.res.c0 + op.res.c].permute(0, 3, 1, 2).float()
    elif op.kind == "maxpool":
        y = F.max_pool2d(src, op.k, op.s, op.k // 2 if op.s == 1 else 0)
    else:
        y = F.interpolate(src, scale_factor=2, mode="nearest")
    return y.permute(0, 2, 3, 1)


def sec_ops(args, impl):
    """Per-op parity: run the whole plan op by op; after each op compare its dst slice with a CPU
    evaluation of that op on the engine's *own* inputs (isolates the failing op)."""
    n = args.batch
    imgsz = args.imgsz
    g = G.build(args.arch, imgsz=imgsz)
    w = W.make_synthetic_weights(g, 0)
    eng = Engine(args.arch, weights=w, max_batch=n, imgsz=imgsz, conv_impl=impl, graph=g)
    tiles = torch.from_numpy(synth.make_tiles(n, imgsz, 5)).cuda()
    eng.preprocess(tiles, "identity")
    torch.cuda.synchronize()
    worst = []
    only = set(args.only.split(",")) if args.only else None
    for i, op in enumerate(g.ops):
        desc = eng.describe_op(i)
        t0 = time.time()
        eng.run_op(i, n)
        torch.cuda.synchronize()
        if only and not any(s in desc for s in only):
            continue
        need = {op.src.buf, op.dst.buf} | ({op.res.buf} if op.res is not None else set())
        cpu = {b: eng.buffer(b, n).float().cpu() for b in need}
        ref = _cpu_op(g, op, w, cpu)
        got = cpu[op.dst.buf][..., op.dst.c0:op.dst.c0 + op.dst.c]
        d = (got - ref).abs()
        tol = 2e-2 * ref.abs().max().item() + 1e-3
        ok = d.max().item() <= tol and torch.isfinite(got).all().item()
        worst.append((d.max().item() / (ref.abs().max().item() + 1e-9), desc))
        log(f"[{i:3d}] {'ok ' if ok else 'BAD'} max|d| {d.max().item():.3e} mean|d| {d.mean().item():.2e} refmax {ref.abs().max().item():.2e} "
            f"{time.time()-t0:.2f}s  {desc}")
        if not ok:
            bad = (d > tol).nonzero()
            log(f"      {len(bad)} bad elements of {d.numel()}; first: {bad[:6].tolist()}")
            # structure of the error: per-channel and per-row/col fractions
            badmask = (d > tol)
            log(f"      bad frac by image {badmask.float().mean((1,2,3)).tolist()}")
            log(f"      bad frac by channel (first 24) {[round(x,2) for x in badmask.float().mean((0,1,2)).tolist()[:24]]}")
            log(f"      bad frac by row (first 24) {[round(x,2) for x in badmask.float().mean((0,2,3)).tolist()[:24]]}")
            log(f"      bad frac by col (first 24) {[round(x,2) for x in badmask.float().mean((0,1,3)).tolist()[:24]]}")
            log(f"      sample got {got[tuple(bad[0].tolist())].item():.4f} ref {ref[tuple(bad[0].tolist())].item():.4f}")
    worst.sort(reverse=True)
    log("worst relative errors:")
    for r, dsc in worst[:8]:
        log(f"   {r:.3e}  {dsc}")


def sec_forward(args):
    from oracle.yolo_torch import make_oracle
    n = args.batch
    g = G.build(args.arch, imgsz=args.imgsz)
    w = W.make_synthetic_weights(g, 0)
    log("weights fingerprint", W.weights_fingerprint(w))
    eng = Engine(args.arch, weights=w, max_batch=n, imgsz=args.imgsz, graph=g)
    tiles_np = synth.make_tiles(n, args.imgsz, 5)
    tiles = torch.from_numpy(tiles_np).cuda()
    eng.preprocess(tiles, "identity")
    eng.forward(n)
    torch.cuda.synchronize()
    x = torch.from_numpy(tiles_np.astype(np.float32) / 255.0).permute(0, 3, 1, 2)
    for emu in (True, False):
        orc = make_oracle(args.arch, w, emulate_bf16=emu)
        raw = orc.raw_head(x)
        log(f"oracle emulate_bf16={emu}")
        for i, lv in enumerate(g.head["levels"]):
            got = eng.buffer(lv["buf"], n).float().cpu()[..., :raw[i].shape[1]]
            stats(f"head level {i} raw", got, raw[i].permute(0, 2, 3, 1))
        rows_ref = orc.decode(raw)
        if args.arch.startswith("yolov8"):
            from oracle.postproc import v8_rows_adapter
            rows_ref = np.stack([v8_rows_adapter(r.numpy()) for r in rows_ref])
        else:
            rr = rows_ref.numpy()
            rows_ref = np.concatenate([rr[..., :5], rr[..., 5:].argmax(-1)[..., None].astype(np.float32)], -1)
        rows = eng.decode_rows(n).cpu().numpy()
        stats("rows box (px)", rows[..., :4], rows_ref[..., :4])
        stats("rows conf", rows[..., 4], rows_ref[..., 4])
        sel = rows_ref[..., 4] >= 0.3
        stats("rows box (px), conf>=0.3", rows[..., :4][sel], rows_ref[..., :4][sel])
        stats("rows conf, conf>=0.3", rows[..., 4][sel], rows_ref[..., 4][sel])
        log(f"  class id agreement {np.mean(rows[..., 5] == rows_ref[..., 5]):.5f}; keep-set (>=0.3) agreement "
            f"{np.mean((rows[..., 4] >= 0.3) == sel):.6f}; n>=0.3: {sel.sum()}")


def sec_post(args):
    from oracle import postproc as OP
    n = 3
    eng = Engine(args.arch, max_batch=n, imgsz=args.imgsz)
    A = eng.num_rows
    rng = np.random.default_rng(3)
    # synthetic rows: clustered boxes so NMS has work
    rows = np.zeros((n, A, 6), np.float32)
    centers = rng.uniform(20, 620, (n, 40, 2)).astype(np.float32)
    idx = rng.integers(0, 40, (n, A))
    rows[..., 0:2] = np.take_along_axis(centers, idx[..., None].repeat(2, -1), 1) + rng.normal(0, 3, (n, A, 2)).astype(np.float32)
    rows[..., 2:4] = rng.uniform(15, 50, (n, A, 2)).astype(np.float32)
    rows[..., 4] = (rng.random((n, A)) ** 6).astype(np.float32)
    rows[..., 5] = rng.integers(0, 2, (n, A)).astype(np.float32)
    rows[0, :50, 4] = 0.5   # exact ties
    rt = torch.from_numpy(rows).cuda()
    # reference filter
    dets, counts = eng.postprocess(n, 0.3, True, rows=rt)
    got = dets_to_numpy(dets, counts)
    okf = True
    for i in range(n):
        ref = OP.filter_rows(rows[i], 0.3)
        g = got[i]
        same = len(ref) == len(g) and np.array_equal(ref[:, 0], g["cx"]) and np.array_equal(ref[:, 4], g["conf"])
        okf &= same
        log(f"filter tile {i}: ref {len(ref)} got {len(g)} identical-order {same}")
    dets, counts = eng.postprocess(n, 0.3, True, top_k=10, rows=rt)
    got = dets_to_numpy(dets, counts)
    for i in range(n):
        ref = OP.top_k_rows(OP.filter_rows(rows[i], 0.3), 10)
        log(f"top10 tile {i}: conf equal {np.array_equal(np.sort(ref[:,4]), np.sort(got[i]['conf']))}")
    # NMS
    pred = np.zeros((n, 6, A), np.float32)
    pred[:, :4] = rows[..., :4].transpose(0, 2, 1)
    cls = rows[..., 5].astype(int)
    for c in range(2):
        pred[:, 4 + c] = np.where(cls == c, rows[..., 4], rows[..., 4] * 0.5)
    rows_nms = np.stack([OP.v8_rows_adapter(p) for p in pred])
    ref = OP.ultralytics_nms(pred, 0.25, 0.7, 300)
    dets, counts = eng.postprocess(n, 0.25, False, iou_thr=0.7, max_det=300, rows=torch.from_numpy(rows_nms).cuda())
    got = dets_to_numpy(dets, counts)
    for i in range(n):
        r = ref[i]
        gx1 = got[i]["cx"] - got[i]["w"] / 2
        same = len(r) == len(got[i]) and np.array_equal(r[:, 4], got[i]["conf"]) and np.allclose(r[:, 0], gx1, atol=1e-4)
        log(f"nms tile {i}: ref {len(r)} got {len(got[i])} identical {same}")
        if not same:
            m = min(len(r), len(got[i]))
            first = next((k for k in range(m) if r[k, 4] != got[i]["conf"][k]), m)
            log(f"    first mismatch at {first}: ref {r[first] if first < len(r) else None} got {got[i][first] if first < len(got[i]) else None}")
    # georef
    dets, counts = eng.postprocess(n, 0.3, True, rows=rt, cap=512)
    params = np.zeros((n, 16))
    for i in range(n):
        params[i, :6] = [-118.2503 + i * 1e-3, -118.2497 + i * 1e-3, 34.0497, 34.0503, 864, 640]
    geo = eng.georef(dets, counts, torch.from_numpy(params).cuda(), "bounds")
    gg = geodets_to_numpy(geo, counts)
    dd = dets_to_numpy(dets, counts)
    bad = 0
    for i in range(n):
        for k in range(len(dd[i])):
            lon, lat, xi, yi = OP.georef_bounds(dd[i]["cx"][k], dd[i]["cy"][k], *params[i, :4], 640, 864)
            if lon != gg[i]["x"][k] or lat != gg[i]["y"][k] or np.float32(xi) != gg[i]["x_img"][k]:
                bad += 1
    log(f"georef bounds: {sum(len(d) for d in dd)} dets, {bad} not bit-exact")
    params2 = np.zeros((n, 16)); params2[:, :4] = params[:, [0, 2, 1, 3]]
    geo = eng.georef(dets, counts, torch.from_numpy(params2).cuda(), "gpuhandler")
    gg = geodets_to_numpy(geo, counts)
    bad = 0
    for i in range(n):
        for k in range(len(dd[i])):
            lon, lat = OP.georef_gpuhandler(dd[i]["cx"][k], dd[i]["cy"][k], *params2[i, :4])
            bad += (lon != gg[i]["x"][k]) or (lat != gg[i]["y"][k])
    log(f"georef gpuhandler: {bad} not bit-exact")
    # dedup
    m = 20000
    x = rng.uniform(0, 300, m); y = rng.uniform(0, 300, m)
    x[:2000] = x[2000:4000] + rng.normal(0, 0.4, 2000); y[:2000] = y[2000:4000] + rng.normal(0, 0.4, 2000)
    conf = rng.random(m).astype(np.float32); conf[:500] = 0.5
    for incl in (True, False):
        keep = eng.dedup(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(conf).cuda(), 1.0, incl).cpu().numpy()
        ref = OP.dedup_greedy(x, y, conf, 1.0, incl)
        refmask = np.zeros(m, bool); refmask[ref] = True
        log(f"dedup inclusive={incl}: ref kept {refmask.sum()} got {keep.sum()} identical {np.array_equal(refmask, keep.astype(bool))}")
    lon = rng.uniform(-118.3, -118.2, 1000); lat = rng.uniform(34.0, 34.1, 1000)
    ux, uy = eng.utm_forward(torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda(), 11, True)
    rx, ry = OP.utm_forward(lon, lat, 11, True)
    log(f"utm forward max |d| {np.abs(ux.cpu().numpy()-rx).max():.3e} m, {np.abs(uy.cpu().numpy()-ry).max():.3e} m")


def sec_time(args):
    n = args.batch
    g = G.build(args.arch, imgsz=args.imgsz)
    eng = Engine(args.arch, max_batch=n, imgsz=args.imgsz, graph=g, conv_impl=args.impl)
    tiles = torch.from_numpy(synth.make_tiles(min(n, 8), args.imgsz, 5)).cuda()
    tiles = tiles.repeat((n + tiles.shape[0] - 1) // tiles.shape[0], 1, 1, 1)[:n].contiguous()
    eng.preprocess(tiles, "identity")
    for _ in range(3):
        eng.forward(n)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(g.ops) + 1)]
    reps = 5
    acc = np.zeros(len(g.ops))
    for _ in range(reps):
        ev[0].record()
        for i in range(len(g.ops)):
            eng.run_op(i, n)
            ev[i + 1].record()
        torch.cuda.synchronize()
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(g.ops))])
    acc /= reps
    tot_flop = 0
    rowsout = []
    for i, op in enumerate(g.ops):
        fl = 0
        if op.kind in ("conv", "dwconv"):
            cout, cing, k, gr = g.wshapes[op.weight]
            db = g.bufs[op.dst.buf]
            fl = 2.0 * n * db.h * db.w * cout * cing * k * k
        tot_flop += fl
        rowsout.append((acc[i], fl / (acc[i] * 1e-3) / 1e12 if acc[i] > 0 else 0, eng.describe_op(i)))
    for i, (ms, tf, d) in enumerate(rowsout):
        log(f"[{i:3d}] {ms*1e3:8.1f} us {tf:7.1f} TF/s  {d}")
    total = acc.sum()
    log(f"sum of per-op times {total:.3f} ms for batch {n}: {n/total*1e3:.0f} tiles/s, {tot_flop/(total*1e-3)/1e12:.1f} TF/s")
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        eng.forward(n)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    log(f"forward() {ms:.3f} ms for batch {n}: {n/ms*1e3:.0f} tiles/s, {tot_flop/(ms*1e-3)/1e12:.1f} TF/s")
    json.dump({"per_op_ms": acc.tolist(), "desc": [r[2] for r in rowsout], "forward_ms": ms, "batch": n},
              open(os.path.join(ROOT, "gpurun_out", f"time_{args.arch}_{args.impl}_b{n}.json"), "w"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("section")
    ap.add_argument("--arch", default="yolov8m")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--impl", default="auto")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log(f"=== diag {a.section} arch={a.arch} batch={a.batch} imgsz={a.imgsz} dev={torch.cuda.get_device_name(0)}")
    t0 = time.time()
    {"pre": sec_pre, "tcops": lambda x: sec_ops(x, "auto"), "forward": sec_forward,
     "post": sec_post, "time": sec_time}[a.section](a)
    log(f"=== done {a.section} in {time.time()-t0:.1f}s")

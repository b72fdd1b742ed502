"""world_size-2 gloo test (CPU) of the N>1 host path: window sharding, seam flags, the two-collective
seam exchange, and the claim that sharded dedup == unsharded dedup.  The device kernels are replaced
by the NumPy oracle here; the same protocol runs over NCCL on the GPUs (mosaic.MosaicDetector.dedup)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aerial_image_recognition_b200 import mosaic as M
from oracle import postproc as OP


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _closure_np(x, y, flag, thr):
    flag = flag.astype(bool).copy()
    changed = True
    while changed:
        changed = False
        fi = np.nonzero(flag)[0]
        for i in np.nonzero(~flag)[0]:
            if len(fi) and (((x[i] - x[fi]) ** 2 + (y[i] - y[fi]) ** 2) <= thr * thr).any():
                flag[i] = True; changed = True
        if changed:
            continue
    return flag


def _greedy_with_key(x, y, conf, key, thr):
    order = np.lexsort((key, ))               # ascending key = the input order the oracle expects
    k = OP.dedup_greedy(x[order], y[order], conf[order], thr, True)
    return order[k]


def _synthetic_detections(H, W, seed):
    """Detections of every window of an H x W mosaic: points in window coordinates, duplicated where
    windows overlap (each window 'sees' every true object inside it, with a little jitter)."""
    rng = np.random.default_rng(seed)
    objs = rng.uniform([0, 0], [W, H], (900, 2))
    wins = M.window_grid(H, W)
    recs = []
    for wid, (x0, y0, w0, h0) in enumerate(wins):
        inside = np.nonzero((objs[:, 0] >= x0) & (objs[:, 0] < x0 + w0) & (objs[:, 1] >= y0) & (objs[:, 1] < y0 + h0))[0]
        for slot, o in enumerate(inside):
            j = rng.normal(0, 1.5, 2)
            recs.append((objs[o, 0] + j[0], objs[o, 1] + j[1], np.float32(rng.random() * 0.5 + 0.4), wid, slot))
    return wins, np.array(recs)


def _worker(rank, world, port, H, W, res, thr, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wins, recs = _synthetic_detections(H, W, 7)
    _, ids, _ = M.shard_windows(H, W, rank, world)
    covers = [M.shard_windows(H, W, r, world)[2] for r in range(world)]
    mine = recs[np.isin(recs[:, 3].astype(np.int64), ids)]
    px, py = mine[:, 0], mine[:, 1]
    gx, gy = px * res, -py * res                               # metric CRS
    conf = mine[:, 2].astype(np.float32)
    key = mine[:, 3].astype(np.int64) * 65536 + mine[:, 4].astype(np.int64)
    flag = M.seam_flags(py, rank, covers, thr / res + 1.0)
    flag = _closure_np(gx, gy, flag, thr)
    lk = _greedy_with_key(gx[~flag], gy[~flag], conf[~flag], key[~flag], thr)
    local_keys = key[~flag][lk]
    rec = M.pack_records(torch.from_numpy(gx[flag]), torch.from_numpy(gy[flag]), torch.from_numpy(conf[flag]),
                         torch.zeros(int(flag.sum()), dtype=torch.int32), torch.from_numpy(key[flag]))      # the product's 32-byte records
    assert rec.dtype == torch.int64 and tuple(rec.shape) == (int(flag.sum()), M.RECORD_WORDS)

    def gather_counts(k):
        t = torch.tensor([k], dtype=torch.int64); out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t); return [int(o) for o in out]

    def gather_padded(r, cap):
        pad = torch.zeros((cap, M.RECORD_WORDS), dtype=torch.int64); pad[:r.shape[0]] = r
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad); return out

    parts, origin = M.exchange_seam(rec, world, gather_counts, gather_padded)
    ax, ay, ac, _, gkey = (t.numpy() for t in M.unpack_records(torch.cat(parts)))
    gk = _greedy_with_key(ax, ay, ac, gkey, thr)
    seam_keys = gkey[gk][origin[gk] == rank]
    q.put((rank, np.concatenate([local_keys, seam_keys]), int(flag.sum()), len(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_seam_exchange_equals_single_rank():
    H, W, res, thr = 2200, 1500, 0.1, 1.0
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, res, thr, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    union = np.sort(np.concatenate([g[1] for g in got]))
    wins, recs = _synthetic_detections(H, W, 7)
    key = recs[:, 3].astype(np.int64) * 65536 + recs[:, 4].astype(np.int64)
    ref = _greedy_with_key(recs[:, 0] * res, -recs[:, 1] * res, recs[:, 2].astype(np.float32), key, thr)
    assert np.array_equal(union, np.sort(key[ref]))
    assert len(union) < len(recs)                                # duplicates really were removed
    exchanged, total = sum(g[2] for g in got), sum(g[3] for g in got)
    assert 0 < exchanged < total // 2                            # only the seam travelled


def test_band_sharding_covers_every_window_once():
    for world in (1, 2, 3, 8):
        ids = np.concatenate([M.shard_windows(40000, 40000, r, world)[1] for r in range(world)])
        assert np.array_equal(ids, np.arange(6241))
    assert M.band_rows(79, 8) == [(0, 10), (10, 20), (20, 30), (30, 40), (40, 50), (50, 60), (60, 70), (70, 79)]
    w = M.window_grid(40000, 40000)
    assert tuple(w[78]) == (39936, 0, 64, 640) and tuple(w[-1]) == (39936, 39936, 64, 64)
    ref = OP.sliding_windows(40000, 40000, 640, 512)
    assert all((x1 - x0, y1 - y0) == (int(a[2]), int(a[3])) and (x0, y0) == (int(a[0]), int(a[1])) for (x0, y0, x1, y1), a in zip(ref[:200], w[:200]))


def test_seam_records_round_trip_bit_exact():
    rng = np.random.default_rng(3)
    x, y = rng.normal(2.3e6, 1e3, 257), rng.normal(6.8e6, 1e3, 257)
    conf = rng.random(257).astype(np.float32); conf[:3] = (-0.0, 1.0, np.float32(0.4) + np.float32(1e-7))
    cls = rng.integers(0, 80, 257).astype(np.int32)
    key = rng.integers(0, 2 ** 40, 257)
    rec = M.pack_records(*(torch.from_numpy(a) for a in (x, y, conf, cls, key)))
    bx, by, bc, bcls, bkey = (t.numpy() for t in M.unpack_records(rec))
    assert np.array_equal(bx, x) and np.array_equal(by, y) and np.array_equal(bc.view(np.uint32), conf.view(np.uint32))
    assert np.array_equal(bcls, cls) and np.array_equal(bkey, key)
    dev = M.seam_flags_device(torch.tensor([100.0, 5100.0, 5115.0, 5200.0, 5260.0, 9000.0]), 0, [(0, 5248), (5120, 10368)], 11.0)
    assert dev.dtype == torch.uint8 and dev.tolist() == [0, 0, 1, 1, 1, 1]


def test_seam_flags_only_near_other_ranks():
    covers = [(0, 5248), (5120, 10368)]
    py = np.array([100.0, 5100.0, 5115.0, 5200.0, 5260.0, 9000.0])
    assert M.seam_flags(py, 0, covers, 11.0).tolist() == [0, 0, 1, 1, 1, 1]
    assert M.seam_flags(py, 1, covers, 11.0).tolist() == [1, 1, 1, 1, 0, 0]

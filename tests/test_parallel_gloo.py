"""The factor-sharded layer's host logic on CPU: shard ranges, compacted shard-local F->V tables
(row order, tile_slots, out_rows, gathered edge types) and the max-all-reduce, world_size 2 over
gloo.  The per-shard arithmetic is done by the CPU oracle here (the kernels need a GPU; their
half of this identity is tests/test_gpu_parity.py::test_sharded_plan_equals_single_gpu)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fgnn_b200 import graphs
from fgnn_b200.parallel import LocalF2V, shard_range
from oracle import fgnn_oracle as orc

N_VARS, C, O, T = 300, 8, 8, 3


def _problem():
    rng = np.random.default_rng(7)
    types = graphs.synthetic_map_graph(N_VARS, 700, 150, 3, seed=3)
    data = []
    for t in types:
        x_f = rng.standard_normal((1, C, t.n_factors, 1)).astype(np.float32)
        et = rng.standard_normal((1, T, t.n_vars, t.kv)).astype(np.float32)
        et[np.broadcast_to(t.pad_f2v[None, None], et.shape)] = 0.0
        W = rng.uniform(-0.5, 0.5, (C, O * T)).astype(np.float32)
        bias = rng.uniform(-0.2, 0.2, O).astype(np.float32)
        bn = dict(weight=rng.uniform(0.8, 1.2, O).astype(np.float32), bias=rng.uniform(-0.2, 0.2, O).astype(np.float32),
                  running_mean=rng.uniform(-0.1, 0.1, O).astype(np.float32),
                  running_var=rng.uniform(0.5, 1.5, O).astype(np.float32))
        data.append(dict(x_f=x_f, et=et, W=W, bias=bias, bn=bn))
    return types, data


def _partial_raw(t, d, lo, hi):
    """Raw max over the local factors' slots, [N, O], -inf where a variable has no local slot: what the
    kernel computes from the compacted table (mask_negative, tile_slots, out_rows)."""
    loc = LocalF2V(t.idx_f2v, t.pad_f2v, lo, hi)
    raw = np.full((t.n_vars, O), -np.inf, np.float32)
    if loc.n_rows == 0:
        return raw, loc
    et_loc = loc.gather_etype(torch.from_numpy(d["et"])).numpy()
    per_slot = orc.mp_conv_forward(d["x_f"][:, :, lo:hi], np.maximum(loc.idx, 0)[None], et_loc, d["W"], None, None,
                                   extension=0, aggregator=None, activation=None)[0]          # [O, rows, kmax]
    dead = np.zeros_like(loc.idx, dtype=bool)
    for r in range(loc.n_rows):
        dead[r] = (loc.idx[r] < 0) | (np.arange(loc.kmax) >= loc.tile_slots[r // 128])
    per_slot = np.where(dead[None], -np.inf, per_slot)
    raw[loc.var] = per_slot.max(2).T
    return raw, loc


def _epilogue(raw, d):
    v = raw + d["bias"][None]
    bn = d["bn"]
    v = (v - bn["running_mean"]) / np.sqrt(bn["running_var"] + 1e-5) * bn["weight"] + bn["bias"]
    return np.maximum(v, 0).astype(np.float32)


def _full(t, d):
    y = orc.mp_conv_forward(d["x_f"], t.idx_f2v[None], d["et"], d["W"], d["bias"], d["bn"], extension=0, aggregator="max")
    return y[0, :, :, 0].T                                                                     # [N, O]


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 100, 101):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def test_compacted_table_invariants():
    types, _ = _problem()
    for t in types:
        total = 0
        for rank in range(3):
            lo, hi = shard_range(t.n_factors, rank, 3)
            loc = LocalF2V(t.idx_f2v, t.pad_f2v, lo, hi)
            live = loc.idx >= 0
            cnt = live.sum(1)
            assert (cnt >= 1).all() and (np.diff(cnt) <= 0).all()               # sorted, no empty rows
            assert (live == (np.arange(loc.kmax)[None] < cnt[:, None])).all()   # live slots first
            assert len(loc.tile_slots) == (loc.n_rows + 127) // 128
            for r in range(loc.n_rows):
                assert cnt[r] <= loc.tile_slots[r // 128] <= loc.kmax
            assert loc.idx.max() < hi - lo and len(set(loc.var.tolist())) == loc.n_rows
            # every live local slot is the original table's entry
            orig = np.take_along_axis(t.idx_f2v[loc.var], loc.slot_pos, axis=1)
            assert (np.where(live, orig - lo, -1) == loc.idx).all()
            total += int(cnt.sum())
        assert total == t.idx_f2v.size                                          # every slot live on exactly one rank


def _worker(rank, world, port, ret):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        types, data = _problem()
        raws = []
        for t, d in zip(types, data):
            lo, hi = shard_range(t.n_factors, rank, world)
            raws.append(_partial_raw(t, d, lo, hi)[0])
        raw = torch.from_numpy(np.concatenate(raws, 1))                         # [N, J*O]: one reduce per layer
        dist.all_reduce(raw, op=dist.ReduceOp.MAX)
        raw = raw.numpy()
        out = sum(_epilogue(raw[:, j * O:(j + 1) * O], d) for j, d in enumerate(data))
        ref = sum(_full(t, d) for t, d in zip(types, data))
        ret[rank] = float(np.abs(out - ref).max())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_f2v_equals_single(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert set(ret.keys()) == {0, 1}
    # max is exact and order independent; the epilogue is applied to identical values
    assert max(ret.values()) <= 1e-6, dict(ret)


def test_single_process_simulation_of_four_ranks():
    types, data = _problem()
    out = 0
    for t, d in zip(types, data):
        raw = np.full((t.n_vars, O), -np.inf, np.float32)
        for rank in range(4):
            lo, hi = shard_range(t.n_factors, rank, 4)
            raw = np.maximum(raw, _partial_raw(t, d, lo, hi)[0])
        assert np.isfinite(raw).all()
        out = out + _epilogue(raw, d)
    ref = sum(_full(t, d) for t, d in zip(types, data))
    assert np.abs(out - ref).max() <= 1e-6

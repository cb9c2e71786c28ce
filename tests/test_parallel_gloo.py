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
            # the padded slots (index 0 + zero edge type) are live only on the rank that owns factor 0, and marked there
            orig_pad = np.take_along_axis(t.pad_f2v[loc.var], loc.slot_pos, axis=1)
            assert (loc.slot_pad == (orig_pad & live)).all()
            assert loc.slot_pad.any() == (bool(t.pad_f2v.any()) and lo == 0)
            assert (loc.idx[loc.slot_pad] == 0).all()
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


# ---------------------------------------------------------------------------------------------
# owner-computes sharding with feature halos (parallel.HaloPartition): host logic over gloo
# ---------------------------------------------------------------------------------------------

def _full_layer(types, data, x_v):
    """One FGNN layer on the whole graph with the oracle: (new x_v [1,O,N,1], new x_f list)."""
    new_v, new_f = 0, []
    for t, d in zip(types, data):
        et_v2f = d["et_v2f"]
        new_f.append(orc.mp_conv_forward(x_v, t.idx_v2f[None], et_v2f, d["W"], d["bias"], d["bn"], extension=0, aggregator="max"))
        new_v = new_v + orc.mp_conv_forward(d["x_f"], t.idx_f2v[None], d["et"], d["W"], d["bias"], d["bn"], extension=0,
                                            aggregator="max")
    return new_v, new_f


def _halo_problem(band):
    rng = np.random.default_rng(11)
    types = graphs.synthetic_map_graph(N_VARS, 600, 150, 3, seed=5, local_band=band)
    if band:
        types = graphs.locality_order(types)
    types2, data = types, []
    for t in types2:
        d = dict(x_f=rng.standard_normal((1, C, t.n_factors, 1)).astype(np.float32),
                 et=rng.standard_normal((1, T, t.n_vars, t.kv)).astype(np.float32),
                 et_v2f=rng.standard_normal((1, T, t.n_factors, t.order)).astype(np.float32),
                 W=rng.uniform(-0.5, 0.5, (C, O * T)).astype(np.float32), bias=rng.uniform(-0.2, 0.2, O).astype(np.float32),
                 bn=dict(weight=np.ones(O, np.float32), bias=np.zeros(O, np.float32), running_mean=np.zeros(O, np.float32),
                         running_var=np.ones(O, np.float32)))
        d["et"][np.broadcast_to(t.pad_f2v[None, None], d["et"].shape)] = 0.0
        data.append(d)
    x_v = rng.standard_normal((1, C, N_VARS, 1)).astype(np.float32)
    return types2, data, x_v


def _halo_worker(rank, world, port, band, out_q):
    from fgnn_b200.parallel import HaloPartition
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        types, data, x_v = _halo_problem(band)
        part = HaloPartition(types, rank, world)

        def pull(own_rows, halo_src, n_rows_by_rank):
            """The halo pull, emulated: all-gather every rank's owned rows, pick (owner, row)."""
            width = own_rows.shape[1]
            pad = max(n_rows_by_rank)
            buf = torch.zeros(pad, width)
            buf[:own_rows.shape[0]] = torch.from_numpy(own_rows)
            gathered = [torch.zeros(pad, width) for _ in range(world)]
            dist.all_gather(gathered, buf)
            owner, row = halo_src
            return np.stack([gathered[int(o)][int(r)].numpy() for o, r in zip(owner, row)]) if len(owner) else np.zeros((0, width), np.float32)

        nv_by_rank = [shard_range(N_VARS, q, world)[1] - shard_range(N_VARS, q, world)[0] for q in range(world)]
        own_v = x_v[0, :, part.v0:part.v1, 0].T                                   # [N_own, C]
        loc_v = np.concatenate([own_v, pull(np.ascontiguousarray(own_v), part.src_v, nv_by_rank)], 0)
        new_v, new_f = 0, []
        for j, (t, d) in enumerate(zip(types, data)):
            f0, f1 = part.fr[j]
            nf_by_rank = [shard_range(t.n_factors, q, world)[1] - shard_range(t.n_factors, q, world)[0] for q in range(world)]
            own_f = d["x_f"][0, :, f0:f1, 0].T
            loc_f = np.concatenate([own_f, pull(np.ascontiguousarray(own_f), part.src_f[j], nf_by_rank)], 0)
            # V->F of my factors, F->V of my variables, local tables
            new_f.append(orc.mp_conv_forward(loc_v.T[None, :, :, None], part.idx_v2f[j][None], d["et_v2f"][:, :, f0:f1], d["W"],
                                             d["bias"], d["bn"], extension=0, aggregator="max"))
            new_v = new_v + orc.mp_conv_forward(loc_f.T[None, :, :, None], part.idx_f2v[j][None], d["et"][:, :, part.v0:part.v1],
                                                d["W"], d["bias"], d["bn"], extension=0, aggregator="max")
        out_q.put((rank, part.v0, part.v1, part.fr, new_v, new_f, len(part.var_halo), [len(h) for h in part.fac_halo]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("band", [0, 32])
def test_halo_partition_two_ranks_gloo(band):
    """world_size 2 over gloo: every rank evaluates its own destinations from [owned | pulled halo] rows through the
    renumbered tables; the owned rows put together equal the single-process layer bit for bit (same slots, same order)."""
    world = 2
    types, data, x_v = _halo_problem(band)
    want_v, want_f = _full_layer(types, data, x_v)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, band, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, v0, v1, fr, new_v, new_f, hv, hf in res:
        assert np.array_equal(new_v, want_v[:, :, v0:v1])
        for j, (f0, f1) in enumerate(fr):
            assert np.array_equal(new_f[j], want_f[j][:, :, f0:f1])
        if band:
            assert hv < 0.5 * (v1 - v0)                       # locality order: the halo is a fraction of the shard

"""GPU-box diagnostic (2+ GPUs): the factor-sharded layer with the fused peer-memory exchange against the
NCCL all-reduce version -- bit-identical variable / factor features over several layers, and timings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgnn_b200  # noqa: E402
from fgnn_b200 import _lib, graphs, parallel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_MAX_CTAS", "16")
dist.init_process_group("nccl", device_id=dev)

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_vars, n_pw, n_hi = 100_000 // scale, 300_000 // scale, 50_000 // scale
C = O = 64
T = 16
L = 4
types = graphs.synthetic_map_graph(n_vars, n_pw, n_hi, 3, seed=0)
J = len(types)
rng = np.random.default_rng(1)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
x_v0 = t(rng.random((1, n_vars, C), dtype=np.float32))
x_f_full = [t(np.abs(rng.standard_normal((1, ty.n_factors, C))).astype(np.float32)) for ty in types]
et_v2f = [t(rng.standard_normal((1, T, ty.n_factors, ty.order)).astype(np.float32)) for ty in types]
et_f2v = []
for ty in types:
    e = rng.standard_normal((1, T, ty.n_vars, ty.kv)).astype(np.float32)
    e[np.broadcast_to(ty.pad_f2v[None, None], e.shape)] = 0.0
    et_f2v.append(t(e))
W = [[{d: dict(filters=t(rng.uniform(-0.05, 0.05, (C, O * T)).astype(np.float32)), bias=t(rng.uniform(0, 0.05, O).astype(np.float32)),
               scale=t(rng.uniform(0.8, 1.2, O).astype(np.float32)), shift=t(rng.uniform(-0.05, 0.05, O).astype(np.float32)))
       for d in ("v2f", "f2v")} for _ in range(J)] for _ in range(L)]


def run(mode, reps):
    plan = parallel.ShardedLayerPlan(types, rank, world, dev, exchange=mode)
    x_f = plan.local_factor_features(x_f_full)
    ev, ef = plan.local_etypes(et_v2f, et_f2v)
    buf_f = [[torch.empty_like(a) for a in x_f] for _ in range(2)]
    if mode == "peer":
        xv = plan.peer_buffers(J, O)
    else:
        xv = [torch.empty_like(x_v0) for _ in range(2)]

    def forward():
        xv[0].copy_(x_v0)
        cur_f = x_f
        src = 0
        for l in range(L):
            nf = buf_f[l & 1]
            if mode == "peer":
                plan.layer_peer(src, cur_f, ev, ef, W[l], nf, last=(l == L - 1))
            else:
                plan.layer(xv[src], cur_f, ev, ef, W[l], xv[src ^ 1], nf)
            cur_f, src = nf, src ^ 1
        if mode == "peer":
            plan.peer_wait()
        return xv[src], cur_f

    out_v, out_f = forward()
    torch.cuda.synchronize()
    dist.barrier()
    res = (out_v.clone(), [a.clone() for a in out_f])
    for _ in range(2):
        forward()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        forward()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return res, float(ms.item())


def time_exchange_alone():
    """the exchange kernel by itself (all ranks launch it together): NVLink-bound"""
    px = parallel.PeerExchange(n_vars, J, O, rank, world, dev)
    for r in px.raws:
        r.normal_()
    bias = torch.zeros(J * O, device=dev)
    out = []
    for ctas in (16, 32, 64, 128):
        px.ctas = ctas
        px2 = parallel.PeerExchange(n_vars, J, O, rank, world, dev, ctas=ctas)     # own counter per grid size
        for r in px2.raws:
            r.normal_()
        torch.cuda.synchronize(); dist.barrier()
        for _ in range(3):
            px2.forward(0, bias, None, None)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            px2.forward(i & 1, bias, None, None)
        e1.record()
        torch.cuda.synchronize()
        us = torch.tensor([e0.elapsed_time(e1) / 10 * 1e3], device=dev)
        dist.all_reduce(us, op=dist.ReduceOp.MAX)
        remote = (world - 1) / world * n_vars * (J * O + O) * 4
        out.append(f"{ctas} CTAs: {float(us):.1f} us ({remote / float(us) / 1e3:.0f} GB/s per rank over the links)")
        dist.barrier()
        px2.close()
    px.close()
    return "; ".join(out)


ex_alone = time_exchange_alone()
if rank == 0:
    print("exchange kernel alone:", ex_alone, flush=True)
(ref_v, ref_f), ms_nccl = run("nccl", 10)
(got_v, got_f), ms_peer = run("peer", 10)
same = torch.equal(ref_v, got_v) and all(torch.equal(a, b) for a, b in zip(ref_f, got_f))
flag = torch.tensor([1 if same else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}: {L} layers, nccl all-reduce {ms_nccl:.3f} ms, peer exchange {ms_peer:.3f} ms per forward; "
          f"bit-identical on all ranks: {bool(flag.item())}; finite: {bool(torch.isfinite(got_v).all())}", flush=True)
dist.destroy_process_group()

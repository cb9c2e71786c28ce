"""GPU-box diagnostic (torchrun, N ranks): where a factor-sharded FGNN layer spends its time.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_probe.py
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
T, C = int(os.environ.get("T", 16)), 64
rng = np.random.default_rng(1)
types = graphs.synthetic_map_graph(100_000, 300_000, 50_000, 3, seed=0)
plan = parallel.ShardedLayerPlan(types, rank, world, dev)
J = len(types)
x_v = torch.rand(1, types[0].n_vars, C, device=dev)
x_f = plan.local_factor_features([torch.rand(1, t.n_factors, C, device=dev) for t in types])
ev, ef = plan.local_etypes([torch.randn(1, T, t.n_factors, t.order, device=dev) for t in types],
                           [torch.randn(1, T, t.n_vars, t.kv, device=dev) for t in types])
W = [{d: dict(filters=torch.randn(C, C * T, device=dev) * 0.01, bias=torch.rand(C, device=dev) * 0.05,
              scale=torch.ones(C, device=dev), shift=torch.zeros(C, device=dev)) for d in ("v2f", "f2v")} for _ in range(J)]
ws = [{"v2f": torch.zeros(C * C * T * 4 + 4096, dtype=torch.uint8, device=dev), "ver_v2f": 1 + j,
       "f2v": torch.zeros(C * C * T * 4 + 4096, dtype=torch.uint8, device=dev), "ver_f2v": 11 + j} for j in range(J)]
out_v = torch.empty_like(x_v)
out_f = [torch.empty_like(a) for a in x_f]
nm = lambda t: t.permute(0, 2, 1).unsqueeze(-1)
raw = torch.empty(1, types[0].n_vars, J * C, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def f2v(sm=0):
    for j in range(J):
        w = W[j]["f2v"]
        fgnn_b200.mp_forward(nm(x_f[j]), plan.idx_f2v[j], ef[j], w["filters"], None, None, None, extension=0, aggregator=0,
                             activation=0, mask_negative=True, out=nm(raw[:, :, j * C:(j + 1) * C]),
                             tile_slots=plan.tile_slots[j], out_rows=plan.out_rows[j], workspace=ws[j]["f2v"],
                             filters_version=ws[j]["ver_f2v"], sm_limit=sm)


def v2f(sm=0):
    for j in range(J):
        w = W[j]["v2f"]
        fgnn_b200.mp_forward(nm(x_v), plan.idx_v2f[j], ev[j], w["filters"], w["bias"], w["scale"], w["shift"], extension=0,
                             aggregator=0, out=nm(out_f[j]), workspace=ws[j]["v2f"], filters_version=ws[j]["ver_v2f"], sm_limit=sm)


res = {
    "f2v_us": timed(f2v), "v2f_us": timed(v2f), "v2f_132sm_us": timed(lambda: v2f(132)),
    "fill_us": timed(lambda: raw.fill_(float("-inf"))),
    "allreduce_us": timed(lambda: dist.all_reduce(raw, op=dist.ReduceOp.MAX)),
    "finish_us": timed(lambda: plan.finish(raw, W, out_v)),
    "layer_us": timed(lambda: plan.layer(x_v, x_f, ev, ef, W, out_v, out_f, _lib.KERNEL_AUTO, ws)),
}
info = {j: dict(rows=plan.f2v[j].n_rows, kmax=plan.f2v[j].kmax, live=plan.f2v[j].live_slots,
                items=int(plan.f2v[j].tile_slots.sum()) * 128) for j in range(J)}
for r in range(world):
    if r == rank:
        print(f"rank {rank}/{world} T={T}: " + " ".join(f"{k}={v:.0f}" for k, v in res.items()), "| f2v tables:", info, flush=True)
    dist.barrier()
dist.destroy_process_group()

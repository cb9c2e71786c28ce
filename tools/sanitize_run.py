"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck) on the GPU box:

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgnn_b200  # noqa: E402
from fgnn_b200 import _lib, graphs, parallel  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
nm = lambda a: a.permute(0, 2, 1).unsqueeze(-1)


def call(B, N, M, K, T, O=64, C=64, ext=0, plan_cap=None, dtype=torch.float32, agg=0):
    x = t(rng.standard_normal((B, N, C)).astype(np.float32)).to(dtype)
    idx = t(rng.integers(0, N, (B, M, K)))
    et = t(rng.standard_normal((B, T, M, K)).astype(np.float32)).to(dtype)
    W = t((rng.uniform(-1, 1, ((2 if ext else 1) * C, O * T)) * 0.1).astype(np.float32))
    bias = t(rng.uniform(0, 0.1, O).astype(np.float32))
    plan = fgnn_b200.SourcePlan(idx, N, row_cap=plan_cap) if plan_cap else None
    y = fgnn_b200.mp_forward(nm(x), idx, et, W, bias, None, None, extension=ext, aggregator=agg, plan=plan)
    torch.cuda.synchronize()
    assert torch.isfinite(y.float()).all()
    return y


call(1, 700, 900, 3, 16)                                   # destination-stationary tcgen05, T = 16
call(2, 300, 500, 6, 4)                                    # T = 4, batched
call(1, 700, 900, 3, 16, dtype=torch.bfloat16)             # bf16 I/O
call(1, 500, 900, 2, 16, plan_cap=3)                       # source-stationary, chunk-alternating epilogue
call(1, 200, 900, 2, 16, plan_cap=6)                       # source-stationary, edge split, hub rows
call(2, 300, 300, 4, 16, ext=2)                            # ORIG_WITH_DIFF on tensor cores (two-atom rows)
call(1, 257, 300, 5, 3, O=6, C=5)                          # SIMT kernel
# padded slots left out of the plan (slot_edge = -2), int32 table
idx = rng.integers(0, 400, (1, 700, 3))
pad = rng.random((700, 3)) < 0.3
pad[:, 0] = False
idx[0][pad] = 0
et = rng.standard_normal((1, 16, 700, 3)).astype(np.float32)
et[0][:, pad] = 0
d_idx = t(idx).int()
fgnn_b200.mp_forward(nm(t(rng.standard_normal((1, 400, 64)).astype(np.float32))), d_idx, t(et),
                     t((rng.uniform(-1, 1, (64, 1024)) * 0.1).astype(np.float32)), None, None, None, extension=0, aggregator=0,
                     plan=fgnn_b200.SourcePlan(d_idx, 400, zero_slots=t(pad)))
# fused aggregation (one batch element per tile) on LDPC-shaped tables
Bq = 70
tbl = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ldpc_factornn.npz"))["idx_v2f"]
idx = t(np.broadcast_to(tbl[None], (Bq, 48, 6)).copy())
fgnn_b200.mp_forward(nm(t(rng.standard_normal((Bq, 96, 64)).astype(np.float32))), idx, t(rng.standard_normal((Bq, 4, 48, 6)).astype(np.float32)),
                     t((rng.uniform(-1, 1, (64, 256)) * 0.1).astype(np.float32)), None, None, None, extension=0, aggregator=0,
                     plan=fgnn_b200.SourcePlan(idx, 96, batch_local=True))
# 1x1 maps through the identity table: C = 64 and the two-slot / two-type form of C = 256, accumulating
from fgnn_b200.mp_nn import conv1x1_native  # noqa: E402
for cin in (64, 256):
    conv = torch.nn.Conv2d(cin, 128, 1).to(dev)
    xm = torch.randn(40, cin, 120, 1, device=dev).contiguous(memory_format=torch.channels_last)
    acc = torch.zeros(40, 128, 120, 1, device=dev).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        assert conv1x1_native(xm, conv.weight, conv.bias, out=acc, accumulate=True) is acc
torch.cuda.synchronize()
# edge model + backward
em = torch.nn.Sequential(torch.nn.Conv2d(7, 64, 1), torch.nn.ReLU(), torch.nn.Conv2d(64, 4, 1)).to(dev)
fgnn_b200.emodel_forward(em, torch.randn(3, 7, 96, 3, device=dev))
mod = fgnn_b200.mp_conv_v2(64, 64, 4, extension=fgnn_b200.mp_conv_type.NO_EXTENSION, aggregtor="max").to(dev).train()
x = torch.randn(2, 64, 50, 1, device=dev, requires_grad=True)
et = torch.randn(2, 4, 70, 3, device=dev, requires_grad=True)
mod(x, torch.randint(0, 50, (2, 70, 3), device=dev), et).sum().backward()
# two simulated ranks: halo pull + owner-computes layers
types = graphs.locality_order(graphs.synthetic_map_graph(600, 1800, 300, 3, seed=1, local_band=64))
plans = [parallel.HaloLayerPlan(types, r, 2, dev, torch.float32, 64, ctas=4) for r in range(2)]
infos = [p.info() for p in plans]
for p in plans:
    p.connect(infos, same_process=True)
xv = t(rng.random((1, 600, 64), dtype=np.float32))
xf = [t(rng.random((1, ty.n_factors, 64), dtype=np.float32)) for ty in types]
ev = [t(rng.standard_normal((1, 16, ty.n_factors, ty.order)).astype(np.float32)) for ty in types]
ef = [t(rng.standard_normal((1, 16, ty.n_vars, ty.kv)).astype(np.float32)) for ty in types]
W = [{d: dict(filters=t(rng.uniform(-0.05, 0.05, (64, 1024)).astype(np.float32)), bias=None, scale=None, shift=None) for d in ("v2f", "f2v")} for _ in types]
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
ets = [p.local_etypes(ev, ef) for p in plans]
for p in plans:
    p.load_features(xv, xf)
torch.cuda.synchronize()
for l in range(2):
    for r, p in enumerate(plans):
        with torch.cuda.stream(streams[r]):
            p.layer(l, ets[r][0], ets[r][1], W, last=(l == 1))
torch.cuda.synchronize()
for p in plans:
    p.close()
print("sanitize_run: all kernels ran,", fgnn_b200.launch_count(), "launches")

"""GPU-box diagnostic: per-tile pipeline timeline of CTA 0 of the source-stationary kernel (pass 1).

    python factor-graph-neural-network_b200/build.py --trace        # here (builds libfgnn_b200_trace.so)
    FGNN_B200_LIB=.../libfgnn_b200_trace.so python tools/src_trace.py [fan_out] [row_cap]

Slots (ns, globaltimer): 0 loader: x stage free  1 x copies issued  2 convert: x landed  3 convert: A stage written
4 mma: A ready  5 mma: tile issued  6/10 epilogue group 0/1: tile start  7/11 edge types landed  8/12 first chunk ready
9/13 tile done  14 edge-type copy issued
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgnn_b200  # noqa: E402
from fgnn_b200 import _lib  # noqa: E402

dev = "cuda:0"
rng = np.random.default_rng(0)
if len(sys.argv) > 1 and sys.argv[1] in ("ldpc_v2f", "ldpc_f2v"):
    # the LDPC decoding graph, batch 4096, T = 4: fused aggregation (one codeword per tile)
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ldpc_factornn.npz"))
    tbl = z["idx_v2f"] if sys.argv[1] == "ldpc_v2f" else z["idx_f2v"]
    B, T, cap = 4096, 4, None
    N = 96 if sys.argv[1] == "ldpc_v2f" else 48
    M, K = tbl.shape
    fan = M * K // N
    idx = torch.from_numpy(tbl[None]).to(dev).expand(B, M, K)
    x = torch.randn(B, N, 64, device=dev).permute(0, 2, 1).unsqueeze(-1)
    et = torch.randn(B, T, M, K, device=dev)
else:
    fan = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    cap = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    T, B = 16, 1
    N = 300_000 if fan == 2 else 100_000
    K = 6 if fan == 2 else 2
    M = N * fan // K
    x = torch.randn(1, N, 64, device=dev).permute(0, 2, 1).unsqueeze(-1)
    idx_np = np.concatenate([rng.permutation(N) for _ in range(fan)]).reshape(1, M, K)
    idx = torch.from_numpy(idx_np).to(dev)
    et = torch.randn(1, T, M, K, device=dev)
W = torch.randn(64, 64 * T, device=dev) * 0.1
out = torch.empty(B, 64, M, 1, device=dev, memory_format=torch.channels_last)
ws = torch.zeros(64 * 64 * T * 4 + 4096, dtype=torch.uint8, device=dev)
plan = fgnn_b200.SourcePlan(idx, N, batch_local=True) if B > 1 else fgnn_b200.SourcePlan(idx, N, row_cap=cap)
for _ in range(3):
    fgnn_b200.mp_forward(x, idx, et, W, None, None, None, extension=0, aggregator=0, kernel=_lib.KERNEL_TCGEN05,
                         out=out, workspace=ws, filters_version=7, plan=plan)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
n = 16 * 4096
buf = np.zeros(n, dtype=np.uint64)
assert lib.fgnn_debug_src_trace_read(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n)) == 0
raw = buf.reshape(4096, 16).astype(np.int64)
tr = raw[:, :15]
items = int((tr[:, 4] > 0).sum())
tr = tr[:items]
if items > 8 and raw[items - 2, 15] > raw[3, 15]:        # SM cycles against nanoseconds between two late tiles
    print(f"SM clock during the kernel: {(raw[items - 2, 15] - raw[3, 15]) / (raw[items - 2, 5] - raw[3, 5]) * 1e3:.0f} MHz")
t0 = tr[tr > 0].min()
tr = np.where(tr > 0, tr - t0, -1)
names = ["x_free", "x_issued", "x_landed", "a_written", "m_ready", "m_issued", "e0_start", "e0_et", "e0_acc", "e0_done",
         "e1_start", "e1_et", "e1_acc", "e1_done", "et_issue"]
print(f"fan-out {fan} cap {cap}: {items} tiles in CTA 0 ({plan.n_rows} virtual rows), span {tr.max() / 1e3:.1f} us -> {tr.max() / items:.0f} ns / tile")
print("tile " + " ".join(f"{n:>9s}" for n in names))
for i in list(range(0, min(items, 12))) + list(range(max(12, items - 4), items)):
    print(f"{i:4d} " + " ".join(f"{v:9d}" for v in tr[i]))
per = np.diff(tr[3:-2], axis=0)
print("median tile-to-tile interval per slot (ns):", " ".join(f"{np.median(per[:, s]):.0f}" for s in range(15)))
m = tr[3:-2]
d = lambda a, b: np.median(m[:, b] - m[:, a])
print(f"median ns: x issue {d(0,1):.0f} | issued->landed {d(1,2):.0f} | convert {d(2,3):.0f} | A written->mma start {d(3,4):.0f} | mma issue {d(4,5):.0f} | "
      f"mma start->first chunk ready {d(4,8):.0f} | g0: start->et {d(6,7):.0f}, et->acc {d(7,8):.0f}, acc->done {d(8,9):.0f} | "
      f"g1: start->et {d(10,11):.0f}, et->acc {d(11,12):.0f}, acc->done {d(12,13):.0f} | et copy issue->g0 sees it {d(14,7):.0f}")

"""GPU-box diagnostic: per-item pipeline timeline of CTA 0 of the tcgen05 kernel.

    python factor-graph-neural-network_b200/build.py --trace        # here (builds libfgnn_b200_trace.so)
    FGNN_B200_LIB=.../libfgnn_b200_trace.so python tools/tc_trace.py [T] [K]

Slots (ns, globaltimer): 0 gather: stage free   1 gather: copies issued   2 convert: raw landed
3 convert: done   4 mma: A ready   5 mma: issued+committed   6 epilogue: accumulator ready   7 epilogue: item done
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
if len(sys.argv) > 1 and sys.argv[1] == "map":          # a 1x1 map through the identity table (mp_nn.conv1x1_native), LDPC rows
    T, K, N = 1, 1, 4096 * 96
    M = N
    idx = torch.arange(N, dtype=torch.int32, device=dev).view(1, N, 1)
else:
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    N, M = 100_000, 300_000 if K == 2 else 100_000
    idx = torch.from_numpy(rng.integers(0, N, (1, M, K))).to(dev)
x = torch.randn(1, N, 64, device=dev).permute(0, 2, 1).unsqueeze(-1)
et = torch.randn(1, T, M, K, device=dev)
W = torch.randn(64, 64 * T, device=dev) * 0.1
out = torch.empty(1, 64, M, 1, device=dev, memory_format=torch.channels_last)
ws = torch.zeros(64 * 64 * T * 4 + 4096, dtype=torch.uint8, device=dev)
for _ in range(3):
    fgnn_b200.mp_forward(x, idx, et, W, None, None, None, extension=0, aggregator=0, kernel=_lib.KERNEL_TCGEN05,
                         out=out, workspace=ws, filters_version=7)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
n = 16 * 4096
buf = np.zeros(n, dtype=np.uint64)
assert lib.fgnn_debug_trace_read(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n)) == 0
tr = buf.reshape(4096, 16).astype(np.int64)[:, :11]
items = int((tr[:, 4] > 0).sum())
tr = tr[:items]
t0 = tr[tr > 0].min()
tr = np.where(tr > 0, tr - t0, -1)
names = ["g_free", "g_issued", "c_landed", "c_done", "m_ready", "m_issued", "e_ready", "e_done", "f_start", "f_staged", "f_stored"]
print(f"T={T} K={K}: {items} items in CTA 0, span {tr.max() / 1e3:.1f} us -> {tr.max() / items:.0f} ns / item")
print("item " + " ".join(f"{n:>9s}" for n in names))
for i in list(range(0, min(items, 14))) + list(range(max(14, items - 6), items)):
    print(f"{i:4d} " + " ".join(f"{v:9d}" for v in tr[i]))
d = lambda a, b: np.median(tr[4:-2, b] - tr[4:-2, a])
print(f"median ns: stage free->copies issued {d(0,1):.0f} | issued->landed {d(1,2):.0f} | convert {d(2,3):.0f} | "
      f"converted->mma start {d(3,4):.0f} | mma issue {d(4,5):.0f} | mma start->acc ready(chunk0) {d(4,6):.0f} | "
      f"epilogue item {d(6,7):.0f}")
per = np.diff(tr[4:-2], axis=0)
print("median item-to-item interval per slot (ns):", " ".join(f"{np.median(per[:, s]):.0f}" for s in range(8)))
fin = tr[(tr[:, 8] >= 0) & (tr[:, 10] >= 0)]
print(f"finish phase (per tile): stage {np.median(fin[:, 9] - fin[:, 8]):.0f} ns, store {np.median(fin[:, 10] - fin[:, 9]):.0f} ns; e_done->f_start {np.median(fin[:, 8] - fin[:, 7]):.0f} ns")

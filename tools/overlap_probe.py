"""GPU-box diagnostic: do the two passes of INDEPENDENT source-stationary calls share the SMs?

Two calls of cfg-2 shape (F->V pairwise: tensor-bound first pass; V->F pairwise: long second pass) are timed back to back on
one stream and on two streams; a plain device copy on the second stream stands in for "any small-block kernel".

    python tools/overlap_probe.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgnn_b200  # noqa: E402
from fgnn_b200 import _lib, graphs  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(1)
types = graphs.synthetic_map_graph(100_000, 300_000, 50_000, 3, seed=0)
C = O = 64
T = 16
nm = lambda t: t.permute(0, 2, 1).unsqueeze(-1)


def make(name, n_src, idx, M, ver):
    K = idx.shape[1]
    x = torch.from_numpy(rng.random((1, n_src, C), dtype=np.float32)).to(dev)
    d_idx = torch.from_numpy(idx[None]).to(dev)
    et = torch.from_numpy(rng.standard_normal((1, T, M, K)).astype(np.float32)).to(dev)
    W = torch.from_numpy(rng.uniform(-0.01, 0.01, (C, O * T)).astype(np.float32)).to(dev)
    out = torch.empty((1, M, O), dtype=torch.float32, device=dev)
    ws = torch.zeros(C * O * T * 4 + 4096, dtype=torch.uint8, device=dev)
    plan = fgnn_b200.SourcePlan(d_idx, n_src)

    def call():
        fgnn_b200.mp_forward(nm(x), d_idx, et, W, None, None, None, extension=0, aggregator=_lib.AGG_MAX,
                             activation=_lib.ACT_RELU, kernel=_lib.KERNEL_TCGEN05, out=nm(out), workspace=ws,
                             filters_version=ver, plan=plan, validate=False)
    return call


ty = types[0]
a = make("f2v0", ty.n_factors, ty.idx_f2v, ty.n_vars, 11)
b = make("v2f0", ty.n_vars, ty.idx_v2f, ty.n_factors, 12)
big0 = torch.empty(50_000_000, dtype=torch.float32, device=dev)
big1 = torch.empty_like(big0)
copy = lambda: big1.copy_(big0)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both(f, g):
    def run():
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        for st, fn in ((s1, f), (s2, g)):
            st.wait_event(ev)
            with torch.cuda.stream(st):
                fn()
            j = torch.cuda.Event()
            j.record(st)
            main.wait_event(j)
    return run


ta, tb, tc = timed(a), timed(b), timed(copy)
print(f"alone: f2v0 {ta:.1f} us, v2f0 {tb:.1f} us, copy(200 MB) {tc:.1f} us")
print(f"one stream: f2v0 + v2f0 {timed(lambda: (a(), b())):.1f} us; f2v0 + copy {timed(lambda: (a(), copy())):.1f} us")
print(f"two streams: f2v0 | v2f0 {timed(both(a, b)):.1f} us; f2v0 | copy {timed(both(a, copy)):.1f} us; v2f0 | copy {timed(both(b, copy)):.1f} us")

"""GPU-box diagnostic: device time of every core call of one FGNN layer of a bench workload, destination-stationary
against source-stationary (row caps 3 / 6), CUDA events over back-to-back repetitions (each call's working set
exceeds L2 at the cfg-2 sizes).

    python tools/call_times.py [--config cfg2] [--edge-types 16] [--reps 20]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgnn_b200  # noqa: E402
from fgnn_b200 import _lib, graphs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vars", type=int, default=100_000)
    ap.add_argument("--pairwise", type=int, default=300_000)
    ap.add_argument("--high", type=int, default=50_000)
    ap.add_argument("--high-order", type=int, default=3)
    ap.add_argument("--edge-types", type=int, default=16)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--modes", default="dst,src3,src6,auto")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(1)
    types = graphs.synthetic_map_graph(args.vars, args.pairwise, args.high, args.high_order, seed=0)
    C = O = 64
    T = args.edge_types
    nm = lambda t: t.permute(0, 2, 1).unsqueeze(-1)
    x_v = torch.from_numpy(rng.random((1, args.vars, C), dtype=np.float32)).to(dev)
    total = {m: 0.0 for m in args.modes.split(",")}
    for j, ty in enumerate(types):
        x_f = torch.from_numpy(np.abs(rng.standard_normal((1, ty.n_factors, C))).astype(np.float32)).to(dev)
        for name, x, idx, M in (("v2f", x_v, ty.idx_v2f, ty.n_factors), ("f2v", x_f, ty.idx_f2v, ty.n_vars)):
            K = idx.shape[1]
            d_idx = torch.from_numpy(idx[None]).to(dev)
            et = torch.from_numpy(rng.standard_normal((1, T, M, K)).astype(np.float32)).to(dev)
            W = torch.from_numpy(rng.uniform(-0.01, 0.01, (C, O * T)).astype(np.float32)).to(dev)
            bias = torch.from_numpy(rng.uniform(0, 0.05, O).astype(np.float32)).to(dev)
            out = torch.empty((1, M, O), dtype=torch.float32, device=dev)
            ws = torch.zeros(C * O * T * 4 + 4096, dtype=torch.uint8, device=dev)
            n_src = x.shape[1]
            line = f"{name}{j} N={n_src} M={M} K={K}:"
            ref = None
            for mode in args.modes.split(","):
                plan = None
                if mode.startswith("src"):
                    plan = fgnn_b200.SourcePlan(d_idx, n_src, row_cap=int(mode[3:]))
                elif mode == "auto":
                    plan = fgnn_b200.SourcePlan(d_idx, n_src)
                    if plan.n_rows * 1.25 > d_idx.numel():
                        plan = None

                def call():
                    fgnn_b200.mp_forward(nm(x), d_idx, et, W, bias, None, None, extension=0, aggregator=_lib.AGG_MAX,
                                         activation=_lib.ACT_RELU, kernel=_lib.KERNEL_TCGEN05, out=nm(out), workspace=ws,
                                         filters_version=7 + j, plan=plan, validate=False)
                for _ in range(3):
                    call()
                torch.cuda.synchronize()
                if ref is None:
                    ref = out.clone()
                else:
                    assert torch.equal(out, ref), f"{name}{j} {mode}: result differs from the first mode"
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    call()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / args.reps * 1e3
                total[mode] += us
                extra = "" if plan is None else f" [cap {plan.row_cap}, {plan.n_rows} vrows]"
                line += f"  {mode} {us:7.1f} us{extra}"
            print(line, flush=True)
    print("layer total:", "  ".join(f"{m} {v:7.1f} us" for m, v in total.items()))


if __name__ == "__main__":
    main()

"""GPU-box diagnostic: the tcgen05 kernel against the SIMT kernel and the CPU oracle on seeded
shapes; prints error statistics per shape (used while bringing the kernel up).

    python tools/tc_check.py [--big]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgnn_b200  # noqa: E402
from fgnn_b200 import _lib  # noqa: E402
from oracle import fgnn_oracle as orc  # noqa: E402

dev = "cuda:0"


def run(B, N, M, K, T, agg=0, O=64, C=64, seed=0, oracle=True, mask=False):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, N, C)).astype(np.float32)
    idx = rng.integers(0, N, (B, M, K))
    if mask:
        idx[rng.random(idx.shape) < 0.3] = -1
    et = rng.standard_normal((B, T, M, K)).astype(np.float32)
    W = (rng.uniform(-1, 1, (C, O * T)) * 0.1).astype(np.float32)
    bias = rng.uniform(-0.2, 0.2, O).astype(np.float32)
    scale = rng.uniform(0.8, 1.2, O).astype(np.float32)
    shift = rng.uniform(-0.2, 0.2, O).astype(np.float32)
    xt = torch.from_numpy(x).to(dev).permute(0, 2, 1).unsqueeze(-1)
    args = (xt, torch.from_numpy(idx).to(dev), torch.from_numpy(et).to(dev), torch.from_numpy(W).to(dev),
            torch.from_numpy(bias).to(dev), torch.from_numpy(scale).to(dev), torch.from_numpy(shift).to(dev))
    kw = dict(extension=0, aggregator=agg, mask_negative=mask)
    y_simt = fgnn_b200.mp_forward(*args, kernel=_lib.KERNEL_SIMT, **kw)
    torch.cuda.synchronize()
    try:
        y_tc = fgnn_b200.mp_forward(*args, kernel=_lib.KERNEL_TCGEN05, **kw)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"B{B} N{N} M{M} K{K} T{T} agg{agg}: TC FAILED: {e}")
        return False
    a, b = y_tc.float().cpu().numpy(), y_simt.cpu().numpy()
    fin = np.isfinite(b)
    same_inf = np.array_equal(np.isfinite(a), fin)
    scale_ = max(np.abs(b[fin]).max(), 1e-30) if fin.any() else 1.0
    err = np.abs(a[fin] - b[fin]).max() / scale_ if fin.any() and same_inf else float("nan")
    msg = f"B{B} N{N} M{M} K{K} T{T} O{O} agg{agg} mask{int(mask)}: tc-vs-simt rel err {err:.3e} (scale {scale_:.3f}) inf-pattern {'ok' if same_inf else 'DIFFERS'}"
    if oracle and not mask:
        bn = dict(weight=scale, bias=shift, running_mean=np.zeros(O, np.float32), running_var=np.ones(O, np.float32) - 1e-5)
        ref = orc.mp_conv_forward_c(np.ascontiguousarray(x.transpose(0, 2, 1))[..., None], idx, et, W, bias, bn,
                                    extension=0, aggregator={0: "max", 1: "softmax", 2: "mean"}[agg])
        msg += f" | tc-vs-oracle {np.abs(a - ref).max() / scale_:.3e} simt-vs-oracle {np.abs(b - ref).max() / scale_:.3e}"
    if not (err <= 1e-4):
        bad = np.argwhere(~(np.abs(a - b) <= 1e-4 * scale_))
        msg += f" | BAD elements {len(bad)} of {a.size}; first {bad[:5].tolist()}"
        r0 = bad[0]
        msg += f" tc={a[tuple(r0)]:.5f} simt={b[tuple(r0)]:.5f}"
        rows = np.unique(bad[:, 2])
        chans = np.unique(bad[:, 1])
        msg += f" | bad rows {len(rows)} (min {rows.min()} max {rows.max()}), bad channels {len(chans)} (min {chans.min()} max {chans.max()})"
    print(msg, flush=True)
    return err <= 1e-4


def run_bf16(B, N, M, K, T, agg=0, O=64, seed=0):
    """bf16 I/O kernel against the C oracle on the bf16-rounded inputs / filters."""
    rng = np.random.default_rng(seed)
    r16 = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    x = r16(rng.standard_normal((B, N, 64)).astype(np.float32))
    idx = rng.integers(0, N, (B, M, K))
    et = r16(rng.standard_normal((B, T, M, K)).astype(np.float32))
    W = (rng.uniform(-1, 1, (64, O * T)) * 0.1).astype(np.float32)
    bias = rng.uniform(-0.2, 0.2, O).astype(np.float32)
    xt = torch.from_numpy(x).to(dev).to(torch.bfloat16).permute(0, 2, 1).unsqueeze(-1)
    try:
        y = fgnn_b200.mp_forward(xt, torch.from_numpy(idx).to(dev), torch.from_numpy(et).to(dev).to(torch.bfloat16),
                                 torch.from_numpy(W).to(dev), torch.from_numpy(bias).to(dev), None, None, extension=0,
                                 aggregator=agg, kernel=_lib.KERNEL_TCGEN05)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"bf16 B{B} N{N} M{M} K{K} T{T} agg{agg}: FAILED: {e}")
        return False
    ref = orc.mp_conv_forward_c(np.ascontiguousarray(x.transpose(0, 2, 1))[..., None], idx, et, r16(W), bias, None,
                                extension=0, aggregator={0: "max", 1: "softmax", 2: "mean"}[agg])
    a = y.float().cpu().numpy()
    scale_ = max(np.abs(ref).max(), 1e-30)
    err = np.abs(a - ref).max() / scale_
    print(f"bf16 B{B} N{N} M{M} K{K} T{T} O{O} agg{agg}: tc-vs-oracle rel err {err:.3e} (scale {scale_:.3f})", flush=True)
    if not err <= 8e-3:
        bad = np.argwhere(~(np.abs(a - ref) <= 8e-3 * scale_))
        print(f"   BAD elements {len(bad)} of {a.size}; first {bad[:5].tolist()}; rows {np.unique(bad[:, 2])[:10]} "
              f"channels {np.unique(bad[:, 1])[:16]}")
    return err <= 8e-3


def run_src(B, N, M, K, T, agg=0, O=64, seed=0):
    """source-stationary path against the destination-stationary kernel (bit-identical by construction)."""
    rng = np.random.default_rng(seed)
    x = torch.from_numpy(rng.standard_normal((B, N, 64)).astype(np.float32)).to(dev).permute(0, 2, 1).unsqueeze(-1)
    idx = torch.from_numpy(rng.integers(0, N, (B, M, K))).to(dev)
    et = torch.from_numpy(rng.standard_normal((B, T, M, K)).astype(np.float32)).to(dev)
    W = torch.from_numpy((rng.uniform(-1, 1, (64, O * T)) * 0.1).astype(np.float32)).to(dev)
    bias = torch.from_numpy(rng.uniform(-0.2, 0.2, O).astype(np.float32)).to(dev)
    try:
        plan = fgnn_b200.SourcePlan(idx, N)
        y0 = fgnn_b200.mp_forward(x, idx, et, W, bias, None, None, extension=0, aggregator=agg, kernel=_lib.KERNEL_TCGEN05)
        y1 = fgnn_b200.mp_forward(x, idx, et, W, bias, None, None, extension=0, aggregator=agg, plan=plan)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"src B{B} N{N} M{M} K{K} T{T} agg{agg}: FAILED: {e}")
        return False
    same = torch.equal(y0, y1)
    diff = float((y0 - y1).abs().max())
    msg = f"src B{B} N{N} M{M} K{K} T{T} O{O} agg{agg}: fan-out {plan.fan_out:.2f} bit-identical {same} max|diff| {diff:.3e}"
    if not same:
        bad = torch.nonzero((y0 != y1)[0, :, :, 0])
        msg += f" | {len(bad)} bad of {y0.numel()}; channels {sorted(set(bad[:, 0].tolist()))[:20]} rows {sorted(set(bad[:, 1].tolist()))[:10]}"
    print(msg, flush=True)
    return same


def timeit_src(N, M, K, T, reps=10):
    rng = np.random.default_rng(0)
    x = torch.randn(1, N, 64, device=dev).permute(0, 2, 1).unsqueeze(-1)
    idx = torch.from_numpy(rng.integers(0, N, (1, M, K))).to(dev)
    et = torch.randn(1, T, M, K, device=dev)
    W = torch.randn(64, 64 * T, device=dev) * 0.1
    bias = torch.zeros(64, device=dev)
    out = torch.empty(1, 64, M, 1, device=dev, memory_format=torch.channels_last)
    ws = torch.zeros(64 * 64 * T * 4 + 4096, dtype=torch.uint8, device=dev)
    plan = fgnn_b200.SourcePlan(idx, N)
    f = lambda: fgnn_b200.mp_forward(x, idx, et, W, bias, None, None, extension=0, aggregator=0, out=out, workspace=ws,
                                     filters_version=7, plan=plan)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"time src  N{N} M{M} K{K} T{T}: {us:9.1f} us  {M * K / us:8.1f} Mslots/s (fan-out {plan.fan_out:.1f})", flush=True)


def timeit(B, N, M, K, T, kernel, reps=10, dtype=torch.float32, C=64):
    rng = np.random.default_rng(0)
    x = torch.randn(B, N, C, device=dev).to(dtype).permute(0, 2, 1).unsqueeze(-1)
    idx = torch.from_numpy(rng.integers(0, N, (B, M, K))).to(dev)
    et = torch.randn(B, T, M, K, device=dev).to(dtype)
    W = torch.randn(C, 64 * T, device=dev) * 0.1
    bias = torch.zeros(64, device=dev)
    out = torch.empty(B, 64, M, 1, device=dev, dtype=dtype, memory_format=torch.channels_last)
    ws = torch.zeros(C * 64 * T * 4 + 4096, dtype=torch.uint8, device=dev)
    f = lambda: fgnn_b200.mp_forward(x, idx, et, W, bias, None, None, extension=0, aggregator=0, kernel=kernel, out=out,
                                     workspace=ws, filters_version=7)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    sz = 2 if dtype == torch.bfloat16 else 4
    byt = 4 * M * K * B + sz * T * M * K * B + sz * C * N * B + sz * 64 * M * B
    print(f"time {'bf16' if sz == 2 else 'fp32'} C{C} B{B} N{N} M{M} K{K} T{T} kernel{kernel}: {us:9.1f} us  {B * M * K / us:8.1f} Mslots/s  "
          f"{byt / us / 1e3:7.1f} GB/s algorithmic", flush=True)


if __name__ == "__main__":
    if "--profile-src" in sys.argv:       # source-stationary launches for ncu: the four cfg-2 calls at T=16
        for (N, M, K) in ((100_000, 300_000, 2), (300_000, 100_000, 6), (100_000, 50_000, 3), (50_000, 100_000, 2)):
            timeit_src(N, M, K, 16, reps=1)
        sys.exit(0)
    if "--profile" in sys.argv:           # two launches for ncu: V2F-pairwise shaped, T=16 and T=4
        timeit(1, 100_000, 300_000, 2, 16, _lib.KERNEL_TCGEN05, reps=1)
        timeit(1, 100_000, 300_000, 2, 4, _lib.KERNEL_TCGEN05, reps=1)
        sys.exit(0)
    ok = True
    ok &= run(1, 200, 128, 1, 16)
    ok &= run(1, 200, 128, 2, 16)
    ok &= run(1, 300, 256, 2, 4)
    ok &= run(1, 300, 1000, 3, 16)
    ok &= run(1, 300, 1000, 3, 1)
    ok &= run(1, 300, 1000, 3, 2)
    ok &= run(1, 300, 1000, 3, 8)
    ok &= run(2, 96, 48, 6, 4)
    ok &= run(1, 5000, 20000, 2, 16)
    ok &= run(1, 5000, 20000, 6, 4, agg=1)
    ok &= run(1, 5000, 20000, 6, 16, agg=1)
    ok &= run(1, 5000, 20000, 3, 16, agg=2)
    ok &= run(1, 500, 3000, 4, 16, mask=True)
    ok &= run(1, 500, 3000, 4, 4, O=128)
    ok &= run(1, 500, 3000, 4, 4, C=128)                 # two K atoms (the 128 -> 64 layer of the LDPC model)
    ok &= run(1, 500, 3000, 3, 16, C=128)
    ok &= run(2, 96, 48, 6, 1, C=128, agg=1)
    ok &= run(64, 48, 96, 3, 4, C=128)
    ok &= run(1, 700, 2000, 2, 8, C=128, agg=2, O=128)
    ok &= run_bf16(1, 200, 128, 1, 16)
    ok &= run_bf16(1, 300, 1000, 3, 16)
    ok &= run_bf16(1, 300, 1000, 3, 4)
    ok &= run_bf16(1, 300, 1000, 3, 8)
    ok &= run_bf16(2, 300, 1000, 3, 1)
    ok &= run_bf16(1, 300, 1000, 3, 2, agg=2)
    ok &= run_bf16(1, 5000, 20000, 6, 16, agg=1)
    ok &= run_bf16(1, 500, 3000, 4, 16, O=128)
    ok &= run_src(1, 200, 128, 1, 16)
    ok &= run_src(1, 200, 1000, 2, 16)
    ok &= run_src(1, 3000, 9000, 2, 16)
    ok &= run_src(1, 3000, 9000, 2, 4)
    ok &= run_src(1, 3000, 9000, 2, 8, agg=2)
    ok &= run_src(1, 50, 3000, 3, 16, agg=1)
    ok &= run_src(2, 300, 1000, 3, 16)
    ok &= run_src(1, 300, 1000, 3, 16, O=128)
    print("ALL OK" if ok else "SOME FAILED")
    if "--big" in sys.argv or ok:
        for pdl in (True, False):
            fgnn_b200.set_programmatic_launch(pdl)
            print(f"programmatic launch {'on' if pdl else 'off'}")
            for T in (16, 4):
                for (N, M, K) in ((100_000, 300_000, 2), (300_000, 100_000, 6), (100_000, 50_000, 3), (50_000, 100_000, 2)):
                    timeit(1, N, M, K, T, _lib.KERNEL_TCGEN05)
        fgnn_b200.set_programmatic_launch(True)
        for T in (16, 4):
            for (N, M, K) in ((100_000, 300_000, 2), (300_000, 100_000, 6), (100_000, 50_000, 3), (50_000, 100_000, 2)):
                timeit_src(N, M, K, T)
        for T in (16, 4):
            for (N, M, K) in ((100_000, 300_000, 2), (300_000, 100_000, 6), (1_000_000, 3_000_000, 2)):
                timeit(1, N, M, K, T, _lib.KERNEL_TCGEN05, dtype=torch.bfloat16)
        timeit(4096, 96, 48, 6, 4, _lib.KERNEL_TCGEN05, C=128)
        timeit(4096, 96, 48, 6, 4, _lib.KERNEL_SIMT, C=128, reps=2)
        timeit(4096, 96, 48, 6, 4, _lib.KERNEL_TCGEN05)
        timeit(4096, 48, 96, 3, 4, _lib.KERNEL_TCGEN05)

"""GPU-box diagnostic: a whole FactorNN forward with the dimensions of the reference's train_ldpc.py
(BASELINE configs[2]: LDPC 96.3.963, check factors T=4 + the global factor T=1/K=96, dims
[64,64,64,128,256,256,128,64,64]) on the native core, with every mp_conv_v2 on its automatically selected kernel
(tcgen05 for C in {64,128}) and, for comparison, forced onto the fp32 CUDA-core kernel.

    python tools/ldpc_model_bench.py [B]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgnn_b200  # noqa: E402
from fgnn_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = "cuda:0"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ldpc_factornn.npz"))
rng = np.random.default_rng(0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

dims = [64, 64, 64, 128, 256, 256, 128, 64, 64]                      # train_ldpc.py
model = fgnn_b200.FactorNN(2, [6, 96], dims, [4, 1], 2, skip_link={4: 3, 5: 2, 7: 0}).to(dev).eval()
for m in model.modules():
    if isinstance(m, fgnn_b200.mp_conv_v2):
        m.enable_weight_cache()

snr = rng.integers(0, 5, (B, 1, 1, 1)).astype(np.float32)
y = (-10 ** (snr / 20) + rng.standard_normal((B, 1, 96, 1))).astype(np.float32)        # all-zero codeword + noise
node = t(np.concatenate([y, np.broadcast_to(snr, y.shape)], 1))
hop = t(rng.standard_normal((B, 6, 48, 1)).astype(np.float32))
nhop = node[:, 0, :, :].reshape(B, 96, 1, 1)
rep = lambda a: t(a)[None].repeat(B, 1, 1)
idx_f2v, idx_v2f = rep(g["idx_f2v"]), rep(g["idx_v2f"])
h_idx_v2f = torch.arange(96, device=dev).reshape(1, 1, 96).repeat(B, 1, 1)
h_idx_f2v = torch.zeros(B, 96, 1, dtype=torch.long, device=dev)
et_f2v, et_v2f = t(rng.standard_normal((B, 4, 96, 3)).astype(np.float32)), t(rng.standard_normal((B, 4, 48, 6)).astype(np.float32))
ones_f2v, ones_v2f = torch.ones(B, 1, 96, 1, device=dev), torch.ones(B, 1, 1, 96, device=dev)


def forward():
    with torch.no_grad():
        return model(node, [hop, nhop], [idx_f2v, h_idx_f2v], [idx_v2f, h_idx_v2f], [et_f2v, ones_f2v], [et_v2f, ones_v2f])


def timed(reps=5):
    for _ in range(2):
        out = forward()
    torch.cuda.synchronize()
    l0 = fgnn_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = forward()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out, (fgnn_b200.launch_count() - l0) // reps


layers = len(dims) - 1
msgs = B * layers * (2 * 96 * 3 + 2 * 96)            # checks: 96x3 + 48x6 slots; global factor: 96 + 96
ms_auto, out_auto, launches = timed()
for m in model.modules():
    if isinstance(m, fgnn_b200.mp_conv_v2):
        m.kernel = _lib.KERNEL_SIMT
ms_simt, out_simt, _ = timed(2)
err = float((out_auto - out_simt).abs().max() / out_simt.abs().max())
print(f"LDPC FactorNN forward, B={B}, {layers} layers, dims {dims}: {ms_auto:.2f} ms with the auto-selected kernels "
      f"({launches} fgnn launches, {msgs / ms_auto / 1e6:.2f} G messages/s incl. the PyTorch 1x1 maps) vs {ms_simt:.2f} ms "
      f"with every core on the SIMT kernel; max relative difference {err:.2e}; "
      f"{int(((out_auto >= 0) != (out_simt >= 0)).sum())} of {out_auto.numel()} hard decisions differ (random weights: "
      f"logits within {float(out_simt.abs()[(out_auto >= 0) != (out_simt >= 0)].max()) if ((out_auto >= 0) != (out_simt >= 0)).any() else 0.0:.1e} "
      f"of zero)")

if "--profile" in sys.argv:            # where the rest of the forward goes (PyTorch's 1x1 maps / norms around the cores)
    for m in model.modules():
        if isinstance(m, fgnn_b200.mp_conv_v2):
            m.kernel = _lib.KERNEL_AUTO
    from torch.profiler import ProfilerActivity, profile
    forward()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        forward()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))

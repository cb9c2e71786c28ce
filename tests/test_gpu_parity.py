"""Parity of the CUDA path (through the C ABI) with the golden vectors of the real reference and
with the CPU oracle on seeded inputs.  Tolerance: 1e-4 relative, fp32 (north_star); decisions
(argmax / sign) bit-exact."""
import ctypes

import numpy as np
import pytest
import torch

import fgnn_b200
from fgnn_b200 import _lib, graphs
from oracle import fgnn_oracle as orc
from tests.util import RTOL, assert_close, bn_of, load_npz, load_unit_cases, rel_err, sub_sd

pytestmark = pytest.mark.gpu
CASES = load_unit_cases()
DEV = "cuda:0"
KERNELS = {"auto": _lib.KERNEL_AUTO, "simt": _lib.KERNEL_SIMT}


def t(a, dev=DEV):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def module_for(case, kernel="auto"):
    m = case["meta"]
    mod = fgnn_b200.mp_conv_v2(m["C"], m["O"], m["T"], bias=m.get("bias", True), bn=m.get("bn", True),
                               extension=fgnn_b200.mp_conv_type(m["ext"]),
                               activation_fn=m.get("act", "relu"), aggregtor=m["agg"])
    mod.load_state_dict({k: torch.from_numpy(v) for k, v in case["sd"].items()})
    mod.kernel = KERNELS[kernel]
    return mod.to(DEV).eval()


@pytest.mark.parametrize("kernel", ["auto", "simt"])
@pytest.mark.parametrize("case", CASES, ids=[c["meta"]["name"] for c in CASES])
def test_unit_cases_vs_reference_golden(case, kernel):
    mod = module_for(case, kernel)
    before = fgnn_b200.launch_count()
    with torch.no_grad():
        y = mod(t(case["x"]), t(case["idx"]), t(case["etype"]))
    assert fgnn_b200.launch_count() > before, "no fgnn_b200 kernel was launched"
    assert tuple(y.shape) == case["out"].shape
    assert_close(y.cpu().numpy(), case["out"], RTOL, case["meta"]["name"])


@pytest.mark.parametrize("layout", ["channels_first", "channels_last", "x3d", "int32_idx", "expanded_tables"])
def test_input_layouts(layout):
    case = next(c for c in CASES if c["meta"]["name"] == "noext_rect")
    mod = module_for(case)
    x, idx, et = t(case["x"]), t(case["idx"]), t(case["etype"])
    if layout == "channels_last":
        x = x.contiguous(memory_format=torch.channels_last)
    elif layout == "x3d":
        x = x[..., 0]
    elif layout == "int32_idx":
        idx = idx.int()
    elif layout == "expanded_tables":
        # batch-replicated tables without the copy (.expand instead of the scripts' .repeat)
        idx = idx[:1].expand(case["x"].shape[0], -1, -1)
        et = et[:1].expand(case["x"].shape[0], -1, -1, -1)
        ref = orc.mp_conv_forward(case["x"], np.repeat(case["idx"][:1], 2, 0), np.repeat(case["etype"][:1], 2, 0),
                                  case["sd"]["filters"], case["sd"]["bias"], bn_of(case["sd"]), 0, "max")
        with torch.no_grad():
            y = mod(x, idx, et)
        assert_close(y.cpu().numpy(), ref, RTOL, layout)
        return
    with torch.no_grad():
        y = mod(x, idx, et)
    assert_close(y.cpu().numpy(), case["out"], RTOL, layout)
    assert y.stride(1) == 1                    # node-major (channels_last) output memory


def test_error_behaviour():
    case = CASES[0]
    mod = module_for(case)
    bad = case["idx"].copy()
    bad[1, 3, 2] = case["x"].shape[2]
    with torch.no_grad(), pytest.raises(IndexError):             # reference: ATen gather raises
        mod(t(case["x"]), t(bad), t(case["etype"]))
    bad[1, 3, 2] = -1
    with torch.no_grad(), pytest.raises(IndexError):
        mod(t(case["x"]), t(bad), t(case["etype"]))
    ext = next(c for c in CASES if c["meta"]["ext"] == 2)
    m2 = module_for(ext)
    with torch.no_grad(), pytest.raises(_lib.FgnnError):         # extension modes need M == N
        m2(t(ext["x"]), t(ext["idx"][:, :5]), t(ext["etype"][:, :, :5]))
    with torch.no_grad(), pytest.raises(AssertionError):         # mp_nn.py:100
        mod(t(case["x"]), t(case["idx"][:1]), t(case["etype"]))


def test_async_index_check_raises_late():
    case = CASES[0]
    mod = module_for(case)
    mod.index_check = "async"
    bad = case["idx"].copy()
    bad[0, 0, 0] = case["x"].shape[2] + 3
    with torch.no_grad():
        mod(t(case["x"]), t(bad), t(case["etype"]))          # no error yet: the scan runs on the stream
        with pytest.raises(IndexError):
            fgnn_b200.check_async_errors(synchronize=True)
        mod(t(case["x"]), t(case["idx"]), t(case["etype"]))  # flag was cleared: good tables pass
        fgnn_b200.check_async_errors(synchronize=True)


def test_custom_aggregator_and_train_mode_bn():
    case = next(c for c in CASES if c["meta"]["name"] == "ext0_max")
    m = case["meta"]
    mod = fgnn_b200.mp_conv_v2(m["C"], m["O"], m["T"], extension=fgnn_b200.mp_conv_type(0),
                               aggregtor=lambda v: torch.max(v, dim=3, keepdim=True)[0])
    mod.load_state_dict({k: torch.from_numpy(v) for k, v in case["sd"].items()})
    mod = mod.to(DEV).eval()
    with torch.no_grad():
        y = mod(t(case["x"]), t(case["idx"]), t(case["etype"]))
    assert_close(y.cpu().numpy(), case["out"], RTOL, "callable aggregator")
    # train-mode BN: batch statistics are computed by the module's own bn on the kernel's output
    mod2 = module_for(case).train()
    with torch.no_grad():
        y2 = mod2(t(case["x"]), t(case["idx"]), t(case["etype"]))
    raw = orc.mp_conv_forward(case["x"], case["idx"], case["etype"], case["sd"]["filters"], case["sd"]["bias"],
                              None, 0, "max", None)
    mu, var = raw.mean((0, 2, 3), keepdims=True), raw.var((0, 2, 3), keepdims=True)
    ref = np.maximum((raw - mu) / np.sqrt(var + 1e-5) * case["sd"]["bn.weight"].reshape(1, -1, 1, 1)
                     + case["sd"]["bn.bias"].reshape(1, -1, 1, 1), 0)
    assert_close(y2.cpu().numpy(), ref, 5e-4, "train-mode bn")


def test_host_buffer_entry_point():
    """fgnn_mp_forward_host: the call a non-PyTorch caller binds (host pointers in, host pointers out)."""
    case = next(c for c in CASES if c["meta"]["name"] == "core64_T4")
    m, sd = case["meta"], case["sd"]
    x = np.ascontiguousarray(case["x"][..., 0])
    idx, et = np.ascontiguousarray(case["idx"]), np.ascontiguousarray(case["etype"])
    scale = (sd["bn.weight"] / np.sqrt(sd["bn.running_var"] + 1e-5)).astype(np.float32)
    shift = (sd["bn.bias"] - sd["bn.running_mean"] * scale).astype(np.float32)
    out = np.empty((m["B"], m["O"], m["M"], 1), np.float32)
    a = _lib.MpArgs()
    a.x, a.idx, a.etype, a.filters = x.ctypes.data, idx.ctypes.data, et.ctypes.data, sd["filters"].ctypes.data
    a.bias, a.bn_scale, a.bn_shift, a.out = sd["bias"].ctypes.data, scale.ctypes.data, shift.ctypes.data, out.ctypes.data
    a.B, a.N, a.M, a.K, a.C, a.O, a.T = m["B"], m["N"], m["M"], m["K"], m["C"], m["O"], m["T"]
    a.extension, a.aggregator, a.activation = 0, _lib.AGG_MAX, _lib.ACT_RELU
    a.dtype, a.idx_dtype, a.kernel, a.gamma = _lib.F32, _lib.I64, _lib.KERNEL_AUTO, 3.0
    _lib.check(_lib.lib().fgnn_mp_forward_host(ctypes.byref(a)), "mp_forward_host")
    assert_close(out, case["out"], RTOL, "host entry")
    idx[0, 0, 0] = m["N"]
    assert _lib.lib().fgnn_mp_forward_host(ctypes.byref(a)) == _lib.ERR_INDEX_RANGE


def test_cfg1_simple_gnn_map_labels_bit_exact():
    """BASELINE configs[0]: train_syn_fixed_pw_hop.py simple_gnn, 128-variable chain, B=8."""
    g = load_npz("cfg1_simple_gnn.npz")
    T_ = fgnn_b200.mp_conv_type
    model = fgnn_b200.mp_sequential(fgnn_b200.mp_conv_v2(2, 64, 16, extension=T_.ORIG_WITH_NEIGHBOR),
                                    fgnn_b200.mp_conv_residual(64, 64, 16), torch.nn.Conv2d(64, 2, 1))
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sub_sd(g, "model").items()})
    emodel = torch.nn.Sequential(torch.nn.Conv2d(1, 64, 1), torch.nn.ReLU(inplace=True), torch.nn.Conv2d(64, 16, 1))
    emodel.load_state_dict({k: torch.from_numpy(v) for k, v in sub_sd(g, "emodel").items()})
    model, emodel = model.to(DEV).eval(), emodel.to(DEV).eval()
    idx, ef = graphs.chain_knn_table(128, 4)
    B = g["x"].shape[0]
    with torch.no_grad():
        etype = emodel(t(ef))
        logits = model(t(g["x"]), t(idx).repeat(B, 1, 1), etype.repeat(B, 1, 1, 1))
    assert_close(logits.cpu().numpy(), g["logits"], RTOL, "logits")
    labels = logits.squeeze(-1).argmax(1).cpu().numpy()
    assert np.array_equal(labels, g["labels"]), "MAP labels differ from the reference"
    assert 0 < labels.sum() < labels.size
    print("cfg1 min |margin| =", float(g["min_margin"]), "max rel err =", rel_err(logits.cpu().numpy(), g["logits"]))


def test_cfg1_factornn_decisions_bit_exact():
    g = load_npz("cfg1_factornn.npz")
    model = fgnn_b200.FactorNN(2, [4], [64, 64], [16], 2)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sub_sd(g, "model").items()})
    model = model.to(DEV).eval()
    B = g["node"].shape[0]
    with torch.no_grad():
        logit = model(t(g["node"]), [t(g["hop"])], [t(g["idx_f2v"]).repeat(B, 1, 1)],
                      [t(g["idx_v2f"]).repeat(B, 1, 1)], [t(g["et_f2v"])], [t(g["et_v2f"])])
    assert_close(logit.cpu().numpy(), g["logit"], RTOL, "logit")
    assert np.array_equal((logit >= 0).cpu().numpy(), g["decision"])


def test_ldpc_factornn_hard_decisions_bit_exact():
    """BASELINE configs[2] shapes: 96.3.963 tables, check factors (T=4) + the global factor (T=1, K=96)."""
    g = load_npz("ldpc_factornn.npz")
    model = fgnn_b200.FactorNN(2, [6, 96], [64, 64, 128, 64], [4, 1], 2, skip_link={2: 0}, ret_high=True)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sub_sd(g, "model").items()})
    model = model.to(DEV).eval()
    B = g["node"].shape[0]
    node = t(g["node"])
    rep = lambda a: t(a)[None].repeat(B, 1, 1)
    h_idx_v2f = torch.arange(96, device=DEV).reshape(1, 1, 96).repeat(B, 1, 1)
    h_idx_f2v = torch.zeros(B, 96, 1, dtype=torch.long, device=DEV)
    nhop = node[:, 0, :, :].reshape(B, 96, 1, 1)
    with torch.no_grad():
        res, nhops = model(node, [t(g["hop"]), nhop], [rep(g["idx_f2v"]), h_idx_f2v],
                           [rep(g["idx_v2f"]), h_idx_v2f],
                           [t(g["et_f2v"]), torch.ones(B, 1, 96, 1, device=DEV)],
                           [t(g["et_v2f"]), torch.ones(B, 1, 1, 96, device=DEV)])
        res = res + node[:, :1]
    assert_close(res.cpu().numpy(), g["res"], RTOL, "res")
    assert_close(nhops[0].cpu().numpy(), g["nhop0"], RTOL, "nhop0")
    assert np.array_equal((res >= 0).cpu().numpy(), g["hard"])


@pytest.mark.parametrize("fe,T,M,K,B", [(7, 4, 96, 3, 5), (7, 4, 48, 6, 5), (3, 16, 60, 2, 4), (2, 16, 1000, 9, 1), (1, 16, 128, 4, 1)])
def test_fused_edge_model_matches_torch_sequential(fe, T, M, K, B):
    """`emodel_forward` (one kernel, hidden vector in registers) == the scripts' Sequential(Conv2d(Fe,64,1), ReLU,
    Conv2d(64,T,1)) (train_ldpc.py:32-38, train_syn_hop_factor.py:174-179); with a plan it also leaves the plan's
    edge-type image, bit-identical to fgnn_src_permute_etype of its own output."""
    torch.manual_seed(fe * 100 + T)
    em = torch.nn.Sequential(torch.nn.Conv2d(fe, 64, 1), torch.nn.ReLU(inplace=True), torch.nn.Conv2d(64, T, 1)).to(DEV).eval()
    ef = torch.randn(B, fe, M, K, device=DEV)
    with torch.no_grad():
        want = em(ef.clone())
    idx = torch.randint(0, 37, (B, M, K), device=DEV)
    plan = fgnn_b200.SourcePlan(idx, 37) if T % 4 == 0 else None
    got = fgnn_b200.emodel_forward(em, ef, plan=plan)
    assert_close(got.cpu().numpy(), want.cpu().numpy(), 1e-5, "fused edge model")
    if plan is not None:
        img = plan._et[3].clone()
        plan._et = None
        ref_img = plan.etype_edges(got, got.stride(0) if B > 1 else T * M * K)
        assert torch.equal(img, ref_img)


def test_fused_edge_model_cfg1_golden():
    g = load_npz("cfg1_simple_gnn.npz")
    emodel = torch.nn.Sequential(torch.nn.Conv2d(1, 64, 1), torch.nn.ReLU(inplace=True), torch.nn.Conv2d(64, 16, 1))
    emodel.load_state_dict({k: torch.from_numpy(v) for k, v in sub_sd(g, "emodel").items()})
    got = fgnn_b200.emodel_forward(emodel.to(DEV).eval(), t(g["efeature"]))
    assert_close(got.cpu().numpy(), g["etype"], 1e-5, "cfg1 etype")


@pytest.mark.parametrize("ext", [0, 1, 2])
@pytest.mark.parametrize("agg", ["max", "softmax", "mean", None])
def test_training_step_gradients_match_aten_autograd(ext, agg):
    """Train mode (train_ldpc.py:222-231 calls loss.backward()): forward with batch-statistics
    BatchNorm and the native backward (csrc/backward.cu + three GEMMs) against ATen's autograd through the reference's
    own op chain (oracle/fgnn_oracle_torch.py) on the same device: output, d x, d etype, d filters, d bias, d BN."""
    from oracle import fgnn_oracle_torch as orct
    torch.manual_seed(10 * ext + len(str(agg)))
    B, C, O, T, N, K = 3, 64, 64, 4, 50, 3
    M = N if ext else 70
    mod = fgnn_b200.mp_conv_v2(C, O, T, extension=fgnn_b200.mp_conv_type(ext), aggregtor=agg).to(DEV).train()
    with torch.no_grad():
        mod.filters.uniform_(-0.2, 0.2)
    x = torch.randn(B, C, N, 1, device=DEV, requires_grad=True)
    et = torch.randn(B, T, M, K, device=DEV, requires_grad=True)
    idx = torch.randint(0, N, (B, M, K), device=DEV)
    probe = torch.randn(B, O, M, K if agg is None else 1, device=DEV)
    before = fgnn_b200.launch_count()
    y = mod(x, idx, et)
    (y * probe).sum().backward()
    assert fgnn_b200.launch_count() >= before + 6            # forward kernel + the five backward kernels
    got = [y.detach(), x.grad.clone(), et.grad.clone(), mod.filters.grad.clone(), mod.bias.grad.clone(), mod.bn.weight.grad.clone()]
    rm, rv = mod.bn.running_mean.clone(), mod.bn.running_var.clone()
    # reference op chain under ATen autograd, same parameters, fresh BatchNorm in train mode
    x2, et2 = x.detach().clone().requires_grad_(True), et.detach().clone().requires_grad_(True)
    W2, b2 = mod.filters.detach().clone().requires_grad_(True), mod.bias.detach().clone().requires_grad_(True)
    bn2 = torch.nn.BatchNorm2d(O).to(DEV).train()
    pre = orct.mp_conv_forward_torch(x2, idx, et2, W2, b2, None, extension=ext, aggregator=agg, activation=None)
    y2 = torch.relu(bn2(pre))
    (y2 * probe).sum().backward()
    want = [y2.detach(), x2.grad, et2.grad, W2.grad, b2.grad, bn2.weight.grad]
    for name, a, b in zip(["y", "dx", "d_etype", "d_filters", "d_bias", "d_bn_weight"], got, want):
        if name == "d_bias":          # batch-statistics BN removes the mean: the bias gradient is zero up to round-off
            assert float(a.abs().max()) < 1e-4 and float(b.abs().max()) < 1e-4
            continue
        assert_close(a.cpu().numpy(), b.cpu().numpy(), 2e-4, f"ext {ext} agg {agg}: {name}")
    assert torch.allclose(rm, bn2.running_mean, rtol=1e-4, atol=1e-6) and torch.allclose(rv, bn2.running_var, rtol=1e-4, atol=1e-6)


def test_training_step_through_factornn_and_residual_wrappers():
    """loss.backward() through FactorNN (mp_conv_residual cores, InstanceNorm maps) in train mode: every parameter gets
    a finite gradient and one SGD step lowers the loss (the scripts' training loops run on the native core)."""
    torch.manual_seed(5)
    g = load_npz("cfg1_factornn.npz")
    model = fgnn_b200.FactorNN(2, [4], [64, 64, 64], [16], 2).to(DEV).train()      # two layers: the V->F modules of the first feed the output
    B = g["node"].shape[0]
    args = (t(g["node"]), [t(g["hop"])], [t(g["idx_f2v"]).repeat(B, 1, 1)], [t(g["idx_v2f"]).repeat(B, 1, 1)],
            [t(g["et_f2v"])], [t(g["et_v2f"])])
    target = (torch.rand(B, 1, 128, 1, device=DEV) > 0.5).float()
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = torch.nn.functional.binary_cross_entropy_with_logits(model(*args), target)
        loss.backward()
        for n_, p_ in model.named_parameters():
            if n_.startswith(("v2f_1_", "f2f_1_")):          # the last layer's factor features are not read by the classifier
                continue
            assert p_.grad is not None and torch.isfinite(p_.grad).all(), n_
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0]


def test_factor_mpnn_merged_tables_labels_bit_exact():
    """The merged-table model of train_syn_hop_factor.py / train_syn_pw_factor.py (factor_mpnn.py:88-133): golden from
    the real reference.  Its ORIG_WITH_DIFF residual cores (C = 64) run on the tensor-core kernel."""
    g = load_npz("factor_mpnn_merged.npz")
    model = fgnn_b200.factor_mpnn(2, [4, 4], [64, 64, 2], [16, 16])
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sub_sd(g, "model").items()})
    model = model.to(DEV).eval()
    B = g["node"].shape[0]
    before = fgnn_b200.launch_count()
    with torch.no_grad():
        out_v, out_f = model(t(g["node"]), [t(g["f_pw"]), t(g["f_hi"])],
                             [[t(g["idx_pw"]).repeat(B, 1, 1), t(g["et_pw"])], [t(g["idx_hi"]).repeat(B, 1, 1), t(g["et_hi"])]])
    assert fgnn_b200.launch_count() > before
    assert_close(out_v.cpu().numpy(), g["out_v"], RTOL, "out_v")
    assert_close(out_f[0].cpu().numpy(), g["out_f0"], RTOL, "out_f0")
    assert_close(out_f[1].cpu().numpy(), g["out_f1"], RTOL, "out_f1")
    labels = out_v.squeeze(-1).argmax(1).cpu().numpy()
    assert np.array_equal(labels, g["labels"]) and 0 < labels.sum() < labels.size


@pytest.mark.parametrize("ext", [1, 2])
@pytest.mark.parametrize("agg", ["max", "softmax", "mean"])
@pytest.mark.parametrize("shape", [dict(B=2, N=300, K=4, O=64, T=16), dict(B=8, N=128, K=4, O=64, T=16),
                                   dict(B=32, N=60, K=9, O=64, T=16), dict(B=1, N=1000, K=3, O=128, T=4),
                                   dict(B=3, N=130, K=2, O=32, T=8)],
                         ids=lambda s: "B{B}_N{N}_K{K}_O{O}_T{T}".format(**s))
def test_extension_modes_on_tensor_cores(ext, agg, shape):
    """ORIG_WITH_NEIGHBOR / ORIG_WITH_DIFF (mp_nn.py:136-159) at C = 64 select the tcgen05 kernel: the A row of a
    slot is [x_m || x_idx] (two K atoms) against the filter image [W_top ; W_bot] resp. [W_top + W_bot ; -W_bot].
    Shapes: cfg 1 (N = 128, K = 4, B = 8) and the factor_mpnn scripts (N + F = 60, K = 2 / 9, B = 32)."""
    rng = np.random.default_rng(ext * 100 + shape["N"])
    B, N, K, O, T = (shape[k] for k in "BNKOT")
    x, idx, et, W0, bias, bn = _random_call(rng, B=B, N=N, M=N, K=K, C=64, O=O, T=T)
    W = (rng.uniform(-1, 1, (128, O * T)) * 0.1).astype(np.float32)
    code = {"max": 0, "softmax": 1, "mean": 2}[agg]
    a = _lib.MpArgs()
    a.x = a.idx = a.etype = a.filters = a.out = 4096
    a.B, a.N, a.M, a.K, a.C, a.O, a.T = B, N, N, K, 64, O, T
    a.x_sc, a.x_sn, a.x_sb, a.out_so, a.out_sm, a.out_sb, a.out_sk = 1, 64, N * 64, 1, O, N * O, 1
    a.extension, a.aggregator = ext, code
    assert _lib.lib().fgnn_mp_select_kernel(ctypes.byref(a)) == _lib.KERNEL_TCGEN05
    y = _native(x, idx, et, W, bias, bn, agg=code, extension=ext)
    y_simt = _native(x, idx, et, W, bias, bn, agg=code, extension=ext, kernel=_lib.KERNEL_SIMT)
    ref = orc.mp_conv_forward_c(x, idx, et, W, bias, bn, extension=ext, aggregator=agg)
    assert_close(y.cpu().numpy(), ref, RTOL, f"ext {ext} {agg} tensor cores")
    assert_close(y_simt.cpu().numpy(), ref, RTOL, f"ext {ext} {agg} simt")


# ---------------------------------------------------------------------------------------------
# seeded inputs vs the CPU oracle at sizes it finishes in seconds; properties at full size
# ---------------------------------------------------------------------------------------------

def _random_call(rng, B, N, M, K, C, O, T, pad_frac=0.1):
    x = rng.standard_normal((B, C, N, 1)).astype(np.float32)
    idx = rng.integers(0, N, (B, M, K))
    et = rng.standard_normal((B, T, M, K)).astype(np.float32)
    et[np.broadcast_to(rng.random((B, 1, M, K)) < pad_frac, et.shape)] = 0
    W = (rng.uniform(-1, 1, (C, O * T)) * 0.3 / np.sqrt(C / 8)).astype(np.float32)
    bias = rng.uniform(-0.2, 0.2, O).astype(np.float32)
    bn = dict(weight=rng.uniform(0.8, 1.2, O).astype(np.float32), bias=rng.uniform(-0.2, 0.2, O).astype(np.float32),
              running_mean=rng.uniform(-0.1, 0.1, O).astype(np.float32),
              running_var=rng.uniform(0.5, 1.5, O).astype(np.float32))
    return x, idx, et, W, bias, bn


def _native(x, idx, et, W, bias, bn, agg=_lib.AGG_MAX, kernel=_lib.KERNEL_AUTO, extension=0, **kw):
    scale = bn["weight"] / np.sqrt(bn["running_var"] + 1e-5)
    shift = bn["bias"] - bn["running_mean"] * scale
    xt = t(x).contiguous(memory_format=torch.channels_last)
    y = fgnn_b200.mp_forward(xt, t(idx), t(et), t(W), t(bias), t(scale.astype(np.float32)),
                             t(shift.astype(np.float32)), extension=extension, aggregator=agg, kernel=kernel, **kw)
    return y


@pytest.mark.parametrize("shape", [
    dict(B=1, N=5000, M=15000, K=2, C=64, O=64, T=16),     # cfg-2 V2F pairwise at 1/20 scale
    dict(B=1, N=15000, M=5000, K=6, C=64, O=64, T=16),     # cfg-2 F2V pairwise
    dict(B=1, N=5000, M=2500, K=3, C=64, O=64, T=16),      # cfg-2 V2F order-3
    dict(B=1, N=5000, M=15000, K=2, C=64, O=64, T=4),
    dict(B=64, N=96, M=48, K=6, C=64, O=64, T=4),          # cfg-3 V2F
    dict(B=64, N=48, M=96, K=3, C=64, O=64, T=4),          # cfg-3 F2V
    dict(B=1, N=777, M=1001, K=5, C=64, O=64, T=16),       # ragged: M not a multiple of any tile
    dict(B=3, N=130, M=257, K=1, C=64, O=64, T=1),
    dict(B=1, N=300, M=129, K=16, C=128, O=64, T=4),
    dict(B=1, N=300, M=500, K=2, C=64, O=128, T=16),
], ids=lambda s: "B{B}_N{N}_M{M}_K{K}_C{C}_O{O}_T{T}".format(**s))
@pytest.mark.parametrize("agg", ["max", "softmax", "mean"])
def test_seeded_vs_oracle(shape, agg):
    rng = np.random.default_rng(hash((shape["M"], shape["K"], shape["T"])) & 0xffff)
    x, idx, et, W, bias, bn = _random_call(rng, **shape)
    ref = orc.mp_conv_forward_c(x, idx, et, W, bias, bn, extension=0, aggregator=agg)
    y = _native(x, idx, et, W, bias, bn, agg={"max": 0, "softmax": 1, "mean": 2}[agg])
    assert_close(y.cpu().numpy(), ref, RTOL, f"{shape} {agg}")


def _bf16_round(a):
    """fp32 array rounded to the nearest bf16 value (still fp32), as torch does it."""
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()


BF16_RTOL = 8e-3      # bf16 output: half an ulp is 2^-9 = 2e-3 of the value; the rest is the epilogue's scale/shift


@pytest.mark.parametrize("shape", [
    dict(B=1, N=5000, M=15000, K=2, C=64, O=64, T=16),     # cfg-4 V2F pairwise shape (scaled)
    dict(B=1, N=15000, M=5000, K=6, C=64, O=64, T=16),     # cfg-4 F2V pairwise
    dict(B=1, N=5000, M=2500, K=4, C=64, O=64, T=16),      # cfg-4 V2F order-4
    dict(B=1, N=5000, M=15000, K=2, C=64, O=64, T=4),
    dict(B=1, N=900, M=1001, K=3, C=64, O=64, T=8),
    dict(B=3, N=130, M=257, K=5, C=64, O=64, T=1),
    dict(B=2, N=130, M=300, K=2, C=64, O=64, T=2),
    dict(B=1, N=300, M=500, K=2, C=64, O=128, T=16),
], ids=lambda s: "B{B}_N{N}_M{M}_K{K}_C{C}_O{O}_T{T}".format(**s))
@pytest.mark.parametrize("agg", ["max", "softmax", "mean"])
def test_bf16_io_vs_oracle(shape, agg):
    """bf16 features / edge types / output (SURVEY 8d cfg 4): the oracle runs in fp32 on the bf16-rounded
    inputs and bf16-rounded filters (the kernel rounds the filters once, accumulates in fp32); the
    result differs by the bf16 rounding of the output and summation order only."""
    rng = np.random.default_rng(hash((shape["M"], shape["K"], shape["T"], 7)) & 0xffff)
    x, idx, et, W, bias, bn = _random_call(rng, **shape)
    x, et = _bf16_round(x), _bf16_round(et)
    ref = orc.mp_conv_forward_c(x, idx, et, _bf16_round(W), bias, bn, extension=0, aggregator=agg)
    scale = (bn["weight"] / np.sqrt(bn["running_var"] + 1e-5)).astype(np.float32)
    shift = (bn["bias"] - bn["running_mean"] * scale).astype(np.float32)
    xt = t(x).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    before = fgnn_b200.launch_count()
    y = fgnn_b200.mp_forward(xt, t(idx), t(et).to(torch.bfloat16), t(W), t(bias), t(scale), t(shift), extension=0,
                             aggregator={"max": 0, "softmax": 1, "mean": 2}[agg], kernel=_lib.KERNEL_TCGEN05)
    assert fgnn_b200.launch_count() > before
    assert y.dtype == torch.bfloat16 and y.stride(1) == 1
    assert_close(y.float().cpu().numpy(), ref, BF16_RTOL, f"bf16 {shape} {agg}")


def test_bf16_accumulate_and_masked():
    rng = np.random.default_rng(3)
    x, idx, et, W, bias, bn = _random_call(rng, B=1, N=700, M=900, K=4, C=64, O=64, T=16, pad_frac=0.0)
    x, et = _bf16_round(x), _bf16_round(et)
    idx[rng.random(idx.shape) < 0.3] = -1
    idx[:, :, 0] = np.abs(idx[:, :, 0])                  # every destination keeps one live slot
    scale = (bn["weight"] / np.sqrt(bn["running_var"] + 1e-5)).astype(np.float32)
    shift = (bn["bias"] - bn["running_mean"] * scale).astype(np.float32)
    args = (t(x).to(torch.bfloat16).contiguous(memory_format=torch.channels_last), t(idx), t(et).to(torch.bfloat16),
            t(W), t(bias), t(scale), t(shift))
    y = fgnn_b200.mp_forward(*args, extension=0, aggregator=0, mask_negative=True)
    # reference: masked slots never win the max -- give them an edge type that makes them lose
    et_m = et.copy()
    live = idx >= 0
    ref_rows = []
    refs = orc.mp_conv_forward_c(x, np.where(live, idx, 0), et_m, _bf16_round(W), None, None, extension=0,
                                 aggregator=None, activation=None)          # [B,O,M,K] raw messages
    raw = np.where(live[:, None], refs, -np.inf).max(3, keepdims=True)
    ref = np.maximum((raw + bias.reshape(1, -1, 1, 1)) * scale.reshape(1, -1, 1, 1) + shift.reshape(1, -1, 1, 1), 0)
    assert_close(y.float().cpu().numpy(), ref, BF16_RTOL, "bf16 masked")
    base = torch.full_like(y, 0.5)
    acc = base.clone()
    fgnn_b200.mp_forward(*args, extension=0, aggregator=0, mask_negative=True, out=acc, accumulate=True)
    assert_close(acc.float().cpu().numpy(), (y.float() + 0.5).cpu().numpy(), BF16_RTOL, "bf16 accumulate")


def test_programmatic_launch_chain_bit_identical():
    """Back-to-back dependent launches (layer l+1 gathers what layer l stored, ping-pong buffers, the
    F->V call accumulating into what the previous call wrote) give bit-identical results with
    programmatic dependent launch on and off."""
    rng = np.random.default_rng(21)
    N, F, K, Kv, C, T = 6000, 18000, 2, 6, 64, 16
    g = graphs.synthetic_map_graph(N, F, 0, 3, seed=4)[0]
    x_v = t(rng.random((1, N, C), dtype=np.float32))
    x_f = t(np.abs(rng.standard_normal((1, F, C))).astype(np.float32))
    et_v2f = t(rng.standard_normal((1, T, F, K)).astype(np.float32))
    et_f2v = t(rng.standard_normal((1, T, N, g.kv)).astype(np.float32))
    idx_v2f, idx_f2v = t(g.idx_v2f[None]), t(g.idx_f2v[None])
    Ws = [t((rng.uniform(-1, 1, (C, C * T)) * 0.1).astype(np.float32)) for _ in range(2)]
    bias = t(rng.uniform(0, 0.05, C).astype(np.float32))
    nm = lambda a: a.permute(0, 2, 1).unsqueeze(-1)
    wsb = [torch.zeros(C * C * T * 4 + 4096, dtype=torch.uint8, device=DEV) for _ in range(2)]

    def chain():
        bv = [x_v.clone(), torch.empty_like(x_v)]
        bf = [x_f.clone(), torch.empty_like(x_f)]
        for l in range(4):
            s, d = l & 1, (l + 1) & 1
            fgnn_b200.mp_forward(nm(bv[s]), idx_v2f, et_v2f, Ws[0], bias, None, None, extension=0, aggregator=0,
                                 out=nm(bf[d]), workspace=wsb[0], filters_version=1)
            fgnn_b200.mp_forward(nm(bf[s]), idx_f2v, et_f2v, Ws[1], bias, None, None, extension=0, aggregator=0,
                                 out=nm(bv[d]), workspace=wsb[1], filters_version=2)
            fgnn_b200.mp_forward(nm(bf[s]), idx_f2v, et_f2v, Ws[1], bias, None, None, extension=0, aggregator=0,
                                 out=nm(bv[d]), accumulate=True, workspace=wsb[1], filters_version=2)
        torch.cuda.synchronize()
        return bv[0].clone(), bf[0].clone()

    prev = fgnn_b200.set_programmatic_launch(False)
    try:
        ref_v, ref_f = chain()
        fgnn_b200.set_programmatic_launch(True)
        for _ in range(3):
            got_v, got_f = chain()
            assert torch.equal(got_v, ref_v) and torch.equal(got_f, ref_f)
    finally:
        fgnn_b200.set_programmatic_launch(prev)
    assert torch.isfinite(ref_v).all() and float(ref_v.abs().max()) > 0


@pytest.mark.parametrize("shape", [
    dict(B=1, N=5000, M=15000, K=2, C=64, O=64, T=16),     # cfg-2 V2F pairwise at 1/20 scale: fan-out 6
    dict(B=1, N=15000, M=5000, K=6, C=64, O=64, T=16),     # cfg-2 F2V pairwise: fan-out 2
    dict(B=1, N=2500, M=5000, K=2, C=64, O=64, T=16),      # cfg-2 F2V order-3 shape
    dict(B=1, N=5000, M=15000, K=2, C=64, O=64, T=4),
    dict(B=1, N=900, M=1001, K=3, C=64, O=64, T=8),
    dict(B=1, N=37, M=3000, K=5, C=64, O=64, T=16),        # fan-out ~400: hub rows split into many virtual rows
    dict(B=3, N=130, M=257, K=4, C=64, O=64, T=16),        # batched, ragged tiles
    dict(B=1, N=300, M=500, K=2, C=64, O=128, T=16),
], ids=lambda s: "B{B}_N{N}_M{M}_K{K}_C{C}_O{O}_T{T}".format(**s))
@pytest.mark.parametrize("row_cap", [3, 6])
@pytest.mark.parametrize("agg", ["max", "softmax", "mean"])
def test_source_stationary_equals_destination_stationary(shape, agg, row_cap):
    """The source-stationary evaluation (one row-product per source row, per-edge messages, second pass)
    is bit-identical to the destination-stationary kernel and within 1e-4 of the oracle."""
    rng = np.random.default_rng(hash((shape["M"], shape["K"], shape["T"], 3)) & 0xffff)
    x, idx, et, W, bias, bn = _random_call(rng, **shape)
    code = {"max": 0, "softmax": 1, "mean": 2}[agg]
    d_idx = t(idx)
    plan = fgnn_b200.SourcePlan(d_idx, shape["N"], row_cap=row_cap)
    assert plan.n_edges == idx.size and int(plan.src_ptr[-1]) == idx.size and int(plan.src_ptr.diff().max()) <= row_cap
    y_dst = _native(x, idx, et, W, bias, bn, agg=code, kernel=_lib.KERNEL_TCGEN05)
    before = fgnn_b200.launch_count()
    scale = bn["weight"] / np.sqrt(bn["running_var"] + 1e-5)
    shift = bn["bias"] - bn["running_mean"] * scale
    y_src = fgnn_b200.mp_forward(t(x).contiguous(memory_format=torch.channels_last), d_idx, t(et), t(W), t(bias),
                                 t(scale.astype(np.float32)), t(shift.astype(np.float32)), extension=0, aggregator=code,
                                 plan=plan)
    assert fgnn_b200.launch_count() >= before + 2
    assert torch.equal(y_src, y_dst), f"max |diff| = {float((y_src - y_dst).abs().max())}"
    ref = orc.mp_conv_forward_c(x, idx, et, W, bias, bn, extension=0, aggregator=agg)
    assert_close(y_src.cpu().numpy(), ref, RTOL, f"source-stationary {shape} {agg}")


def test_source_stationary_masked_accumulate_and_module_auto():
    rng = np.random.default_rng(17)
    x, idx, et, W, bias, bn = _random_call(rng, B=1, N=700, M=4000, K=4, C=64, O=64, T=16, pad_frac=0.0)
    idx[rng.random(idx.shape) < 0.3] = -1
    d_idx = t(idx)
    plan = fgnn_b200.SourcePlan(d_idx, 700, mask_negative=True)
    assert plan.n_edges == int((idx >= 0).sum())
    args = (t(x).contiguous(memory_format=torch.channels_last), d_idx, t(et), t(W), t(bias), None, None)
    y_dst = fgnn_b200.mp_forward(*args, extension=0, aggregator=0, mask_negative=True, kernel=_lib.KERNEL_TCGEN05)
    y_src = fgnn_b200.mp_forward(*args, extension=0, aggregator=0, mask_negative=True, plan=plan)
    assert torch.equal(y_src, y_dst)                       # including the -inf rows of destinations with no live slot
    acc_d, acc_s = torch.full_like(y_dst, 0.25), torch.full_like(y_dst, 0.25)
    ok = torch.isfinite(y_dst)
    fgnn_b200.mp_forward(*args, extension=0, aggregator=0, mask_negative=True, kernel=_lib.KERNEL_TCGEN05, out=acc_d,
                         accumulate=True)
    fgnn_b200.mp_forward(*args, extension=0, aggregator=0, mask_negative=True, plan=plan, out=acc_s, accumulate=True)
    assert torch.equal(acc_s[ok], acc_d[ok])
    # the module picks the plan by itself for a table whose sources feed many slots, and says so
    mod = fgnn_b200.mp_conv_v2(64, 64, 16, extension=fgnn_b200.mp_conv_type.NO_EXTENSION, aggregtor="max").to(DEV).eval()
    mod.AUTO_MIN_SLOTS, mod.AUTO_MIN_USES = 1000, 1
    big = t(rng.integers(0, 700, (1, 6000, 2)))
    xe, ete = t(x).contiguous(memory_format=torch.channels_last), t(rng.standard_normal((1, 16, 6000, 2)).astype(np.float32))
    with torch.no_grad():
        y_auto = mod(xe, big, ete)
        assert mod._plan_for(xe, big, ete, 0, _lib.AGG_MAX) is not None
        mod.source_stationary = False
        y_off = mod(xe, big, ete)
    assert torch.equal(y_auto, y_off)


@pytest.mark.parametrize("agg", ["max", "softmax", "mean"])
def test_source_plan_zero_slots_equal_evaluated_padding(agg):
    """SourcePlan(zero_slots=...): padded slots (valid index 0 + all-zero edge type, the reference's convention) are left
    out of the plan and aggregated as constant-zero messages -- bit-identical to evaluating them (destination-stationary
    kernel and a plan without the hint), while the pad target no longer collects an edge per padded slot."""
    rng = np.random.default_rng(11)
    N, M, K, T, C = 700, 1500, 3, 16, 64
    idx = rng.integers(0, N, (1, M, K))
    pad = rng.random((M, K)) < 0.35
    pad[:, 0] = False                                                  # every destination keeps a real slot
    idx[0][pad] = 0
    et = rng.standard_normal((1, T, M, K)).astype(np.float32)
    et[0][:, pad] = 0.0
    x = t(rng.standard_normal((1, N, C)).astype(np.float32)).permute(0, 2, 1).unsqueeze(-1)
    W = t(rng.uniform(-0.1, 0.1, (C, C * T)).astype(np.float32))
    bias = t(rng.uniform(0, 0.05, C).astype(np.float32))
    d_idx, d_et = t(idx), t(et)
    code = {"max": _lib.AGG_MAX, "softmax": _lib.AGG_SOFTMAX, "mean": _lib.AGG_MEAN}[agg]
    run = lambda plan: fgnn_b200.mp_forward(x, d_idx, d_et, W, bias, None, None, extension=0, aggregator=code,
                                            kernel=_lib.KERNEL_TCGEN05, plan=plan)
    ref = run(None)
    plain = fgnn_b200.SourcePlan(d_idx, N, row_cap=3)
    hinted = fgnn_b200.SourcePlan(d_idx, N, row_cap=3, zero_slots=t(pad))
    assert hinted.n_edges == int((~pad).sum()) and hinted.n_rows < plain.n_rows
    assert int((hinted.slot_edge == -2).sum()) == int(pad.sum())
    assert torch.equal(run(plain), ref)
    assert torch.equal(run(hinted), ref)
    if agg == "max":                                                   # the module-level hint reaches the plan
        mod = fgnn_b200.mp_conv_v2(C, C, T, extension=fgnn_b200.mp_conv_type.NO_EXTENSION, aggregtor="max").to(DEV).eval()
        mod.source_stationary = True
        with torch.no_grad():
            y_plain = mod(x, d_idx, d_et)
            mod2 = fgnn_b200.mp_conv_v2(C, C, T, extension=fgnn_b200.mp_conv_type.NO_EXTENSION, aggregtor="max").to(DEV).eval()
            mod2.load_state_dict(mod.state_dict())
            mod2.source_stationary = True
            mod2.zero_edge_type_slots = t(pad)
            idx2 = d_idx.clone()                                       # its own table object: its own cached plan
            y_hint = mod2(x, idx2, d_et)
        assert torch.equal(y_plain, y_hint)
        assert int((fgnn_b200.SourcePlan.for_table(idx2, N, zero_slots=mod2.zero_edge_type_slots).slot_edge == -2).sum()) == int(pad.sum())


@pytest.mark.parametrize("agg", ["max", "softmax", "mean"])
@pytest.mark.parametrize("direction", ["v2f", "f2v"])
def test_fused_aggregation_ldpc_shapes(direction, agg):
    """The LDPC decoding graph (96.3.963 tables of the golden fixture, batch of codewords, T = 4): with a plan the
    first pass keeps the messages in shared memory and aggregates per codeword inside the CTA (one launch, no message
    buffer) -- bit-identical to the two-pass evaluation and to the destination-stationary kernel, and within 1e-4 of
    the oracle.  V->F: 96 source rows per codeword, 3 edges each; F->V: 48 source rows of 6 edges = 96 virtual rows."""
    g = load_npz("ldpc_factornn.npz")
    rng = np.random.default_rng(7 + len(agg))
    B, C, O, T = 37, 64, 64, 4
    tbl = g["idx_v2f"] if direction == "v2f" else g["idx_f2v"]          # [48,6] values < 96  |  [96,3] values < 48
    M, K = tbl.shape
    N = 96 if direction == "v2f" else 48
    idx = np.broadcast_to(tbl[None], (B, M, K)).copy()
    x = rng.standard_normal((B, C, N, 1)).astype(np.float32)
    et = rng.standard_normal((B, T, M, K)).astype(np.float32)
    W = (rng.uniform(-1, 1, (C, O * T)) * 0.1).astype(np.float32)
    bias = rng.uniform(-0.2, 0.2, O).astype(np.float32)
    bn = dict(weight=rng.uniform(0.8, 1.2, O).astype(np.float32), bias=rng.uniform(-0.2, 0.2, O).astype(np.float32),
              running_mean=rng.uniform(-0.1, 0.1, O).astype(np.float32), running_var=rng.uniform(0.5, 1.5, O).astype(np.float32))
    code = {"max": 0, "softmax": 1, "mean": 2}[agg]
    d_idx = t(idx)
    plan = fgnn_b200.SourcePlan(d_idx, N, batch_local=True)
    plan2 = fgnn_b200.SourcePlan(d_idx, N)                       # ordinary plan: two passes
    assert plan.fusable(O, T) and plan.rows_per_batch == 96 and plan.n_rows == B * 96
    scale = bn["weight"] / np.sqrt(bn["running_var"] + 1e-5)
    shift = bn["bias"] - bn["running_mean"] * scale
    args = (t(x).contiguous(memory_format=torch.channels_last), d_idx, t(et), t(W), t(bias), t(scale.astype(np.float32)),
            t(shift.astype(np.float32)))
    y_dst = fgnn_b200.mp_forward(*args, extension=0, aggregator=code, kernel=_lib.KERNEL_TCGEN05)
    y_fused = fgnn_b200.mp_forward(*args, extension=0, aggregator=code, plan=plan)
    assert plan._msg is None                                       # no message buffer was ever allocated
    l0 = fgnn_b200.launch_count()
    y_fused = fgnn_b200.mp_forward(*args, extension=0, aggregator=code, plan=plan)
    fused_n = fgnn_b200.launch_count() - l0
    y_two = fgnn_b200.mp_forward(*args, extension=0, aggregator=code, plan=plan2)      # (first call also permutes the edge types)
    la = fgnn_b200.launch_count()
    y_two = fgnn_b200.mp_forward(*args, extension=0, aggregator=code, plan=plan2)
    assert plan._msg is None and plan2._msg is not None and fused_n == (fgnn_b200.launch_count() - la) - 1    # no second pass
    assert torch.equal(y_fused, y_dst), f"max |diff| = {float((y_fused - y_dst).abs().max())}"
    assert torch.equal(y_two, y_dst)
    ref = orc.mp_conv_forward_c(x, idx, et, W, bias, bn, extension=0, aggregator=agg)
    assert_close(y_fused.cpu().numpy(), ref, RTOL, f"fused {direction} {agg}")
    acc = torch.full_like(y_dst, 0.5)
    fgnn_b200.mp_forward(*args, extension=0, aggregator=code, plan=plan, out=acc, accumulate=True)
    assert torch.allclose(acc, y_dst + 0.5, rtol=0, atol=1e-6)


def test_source_stationary_column_slices():
    """O*T must be 256 or a multiple of 512 (a CTA owns 256 or 512 filter columns; round-1 advice: widths like
    O = 48 at T = 16 silently skipped the trailing 256 columns).  O = 48: the explicit plan raises and the module
    falls back to the destination-stationary kernel; O = 96 (three slices of 512) runs source-stationary."""
    rng = np.random.default_rng(5)
    for O in (48, 96):
        x, idx, et, W, bias, bn = _random_call(rng, B=1, N=300, M=900, K=2, C=64, O=O, T=16)
        d_idx = t(idx)
        plan = fgnn_b200.SourcePlan(d_idx, 300)
        args = (t(x).contiguous(memory_format=torch.channels_last), d_idx, t(et), t(W), t(bias), None, None)
        ref = orc.mp_conv_forward_c(x, idx, et, W, bias, None, extension=0, aggregator="max")
        mod = fgnn_b200.mp_conv_v2(64, O, 16, extension=fgnn_b200.mp_conv_type.NO_EXTENSION, aggregtor="max", bn=False).to(DEV).eval()
        mod.AUTO_MIN_SLOTS, mod.AUTO_MIN_USES = 100, 1
        with torch.no_grad():
            mod.filters.copy_(t(W)); mod.bias.copy_(t(bias))
            if O == 48:
                with pytest.raises(fgnn_b200.FgnnError):
                    fgnn_b200.mp_forward(*args, extension=0, aggregator=0, plan=plan)
                assert mod._plan_for(args[0], d_idx, args[2], 0, _lib.AGG_MAX) is None
            else:
                y_src = fgnn_b200.mp_forward(*args, extension=0, aggregator=0, plan=plan)
                assert_close(y_src.cpu().numpy(), ref, RTOL, f"O={O} source-stationary")
            y = mod(args[0], d_idx, args[2])
        assert_close(y.cpu().numpy(), ref, RTOL, f"O={O}")


def test_masked_slots_and_epilogue_split_equal_fused():
    """Shard-local tables: negative indices are empty slots excluded from the max; the
    raw aggregate of two half-tables, max-combined and passed through fgnn_epilogue_forward,
    equals the fused single call -- the 1-GPU == N-GPU identity of the factor-sharded layer."""
    rng = np.random.default_rng(11)
    x, idx, et, W, bias, bn = _random_call(rng, B=1, N=400, M=300, K=6, C=64, O=64, T=4, pad_frac=0.0)
    full = _native(x, idx, et, W, bias, bn)
    scale = (bn["weight"] / np.sqrt(bn["running_var"] + 1e-5)).astype(np.float32)
    shift = (bn["bias"] - bn["running_mean"] * scale).astype(np.float32)
    parts = []
    for lo, hi in ((0, 200), (200, 400)):                 # two shards of the SOURCE nodes (factors)
        local = np.where((idx >= lo) & (idx < hi), idx - lo, -1)
        y = fgnn_b200.mp_forward(t(x[:, :, lo:hi]).contiguous(memory_format=torch.channels_last), t(local), t(et),
                                 t(W), None, None, None, extension=0, aggregator=_lib.AGG_MAX,
                                 activation=_lib.ACT_NONE, mask_negative=True)
        parts.append(y)
    red = torch.maximum(parts[0], parts[1])               # what ncclAllReduce(max) computes
    rows = red.shape[0] * red.shape[2]
    flat = red.permute(0, 2, 3, 1).contiguous()
    d_bias, d_scale, d_shift = t(bias), t(scale), t(shift)     # keep the device buffers alive across the call
    _lib.check(_lib.lib().fgnn_epilogue_forward(
        flat.data_ptr(), flat.data_ptr(), rows, 64, d_bias.data_ptr(), d_scale.data_ptr(), d_shift.data_ptr(),
        _lib.ACT_RELU, 0.0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "epilogue")
    torch.cuda.synchronize()
    got = flat.permute(0, 3, 1, 2)
    assert torch.equal(got, full), "sharded max + epilogue differs from the fused call"


def test_full_size_properties_cfg2():
    """BASELINE configs[1] full size (100K vars, 300K pairwise, T=16): properties that need no
    full-size oracle -- (1) permuting the slots of every destination leaves the max unchanged,
    (2) a 2 000-row sample of destinations matches the oracle run on just those rows,
    (3) an all-zero edge type yields act(BN(bias)), (4) run-to-run bit-identical."""
    rng = np.random.default_rng(2)
    N, M, K, C, O, T = 100_000, 300_000, 2, 64, 64, 16
    x, idx, et, W, bias, bn = _random_call(rng, 1, N, M, K, C, O, T)
    y = _native(x, idx, et, W, bias, bn)
    y2 = _native(x, idx, et, W, bias, bn)
    assert torch.equal(y, y2)
    yp = _native(x, idx[:, :, ::-1].copy(), et[:, :, :, ::-1].copy(), W, bias, bn)
    assert torch.equal(y, yp)
    rows = rng.choice(M, 2000, replace=False)
    ref = orc.mp_conv_forward_c(x, idx[:, rows], et[:, :, rows], W, bias, bn, extension=0, aggregator="max")
    assert_close(y[:, :, torch.from_numpy(rows).to(DEV)].cpu().numpy(), ref, RTOL, "sampled rows")
    yz = _native(x, idx, np.zeros_like(et), W, bias, bn)
    scale = bn["weight"] / np.sqrt(bn["running_var"] + 1e-5)
    const = np.maximum(bias * scale + bn["bias"] - bn["running_mean"] * scale, 0).astype(np.float32)
    assert_close(yz.cpu().numpy(), np.broadcast_to(const.reshape(1, O, 1, 1), (1, O, M, 1)), 1e-6, "zero etype")


@pytest.mark.parametrize("T", [16, 4])
def test_full_size_cfg2_all_four_calls_vs_oracle(T):
    """BASELINE configs[1] at FULL size, all four core calls of a layer (V->F / F->V of the pairwise and the order-3
    type, the bench's own graph with its zero-etype pads): destination-stationary and source-stationary evaluation
    against the C oracle on 4 000 sampled destination rows per call (the oracle evaluates exactly those rows from the
    full source features), and bit-identical to each other."""
    rng = np.random.default_rng(20 + T)
    types = graphs.synthetic_map_graph(100_000, 300_000, 50_000, 3, seed=0)
    C = O = 64
    x_v = rng.random((1, C, 100_000, 1), dtype=np.float32)
    for ty in types:
        x_f = np.abs(rng.standard_normal((1, C, ty.n_factors, 1))).astype(np.float32)
        for name, x, idx, pad in (("v2f", x_v, ty.idx_v2f, None), ("f2v", x_f, ty.idx_f2v, ty.pad_f2v)):
            M, K = idx.shape
            et = rng.standard_normal((1, T, M, K)).astype(np.float32)
            if pad is not None:
                et[np.broadcast_to(pad[None, None], et.shape)] = 0.0
            W = (rng.uniform(-1, 1, (C, O * T)) * 0.05).astype(np.float32)
            bias = rng.uniform(0, 0.05, O).astype(np.float32)
            bn = dict(weight=rng.uniform(0.8, 1.2, O).astype(np.float32), bias=rng.uniform(-0.1, 0.1, O).astype(np.float32),
                      running_mean=rng.uniform(-0.05, 0.05, O).astype(np.float32), running_var=rng.uniform(0.5, 1.5, O).astype(np.float32))
            y = _native(x, idx[None], et, W, bias, bn, kernel=_lib.KERNEL_TCGEN05)
            d_idx = t(idx[None])
            plan = fgnn_b200.SourcePlan(d_idx, x.shape[2])
            scale = bn["weight"] / np.sqrt(bn["running_var"] + 1e-5)
            shift = bn["bias"] - bn["running_mean"] * scale
            y_src = fgnn_b200.mp_forward(t(x).contiguous(memory_format=torch.channels_last), d_idx, t(et), t(W), t(bias),
                                         t(scale.astype(np.float32)), t(shift.astype(np.float32)), extension=0, aggregator=0, plan=plan)
            assert torch.equal(y, y_src), f"{ty.name} {name}: source-stationary differs"
            rows = np.sort(rng.choice(M, 4000, replace=False))
            ref = orc.mp_conv_forward_c(x, idx[None][:, rows], et[:, :, rows], W, bias, bn, extension=0, aggregator="max")
            assert_close(y[:, :, torch.from_numpy(rows).to(DEV)].cpu().numpy(), ref, RTOL, f"{ty.name} {name} T={T}")


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_local_tables_source_stationary_equal_destination_stationary(world):
    """ShardedLayerPlan.source_plan: the shard-local tables (V->F over all variables; the compacted, masked F->V table with
    tile_slots / out_rows and the padded slots left out) evaluated source-stationary give bit for bit what the
    destination-stationary kernel gives for them -- the factor-sharded layers use these plans (cfg 2, T = 16)."""
    from fgnn_b200 import parallel
    rng = np.random.default_rng(21)
    C = O = 64
    T = 16
    types = graphs.synthetic_map_graph(10_000 * world, 30_000 * world, 5_000 * world, 3, seed=4)      # shards big enough to plan
    dev = torch.device(DEV)
    nm = lambda a: a.permute(0, 2, 1).unsqueeze(-1)
    x_v = t(rng.random((1, types[0].n_vars, C), dtype=np.float32))
    x_f = [t(np.abs(rng.standard_normal((1, ty.n_factors, C))).astype(np.float32)) for ty in types]
    et_v2f = [t(rng.standard_normal((1, T, ty.n_factors, ty.order)).astype(np.float32)) for ty in types]
    et_f2v = []
    for ty in types:
        e = rng.standard_normal((1, T, ty.n_vars, ty.kv)).astype(np.float32)
        e[np.broadcast_to(ty.pad_f2v[None, None], e.shape)] = 0.0
        et_f2v.append(t(e))
    W = t((rng.uniform(-1, 1, (C, O * T)) * 0.1).astype(np.float32))
    bias = t(rng.uniform(-0.2, 0.2, O).astype(np.float32))
    used = 0
    for rank in (0, world - 1):                                   # rank 0 owns factor 0, the pad target
        plan = parallel.ShardedLayerPlan(types, rank, world, dev)
        xf_loc = plan.local_factor_features(x_f)
        ev, ef = plan.local_etypes(et_v2f, et_f2v)
        for j in range(len(types)):
            sp = plan.source_plan("v2f", j, T)
            if sp is not None:
                used += 1
                a = fgnn_b200.mp_forward(nm(x_v), plan.idx_v2f[j], ev[j], W, bias, None, None, extension=0, aggregator=0)
                b = fgnn_b200.mp_forward(nm(x_v), plan.idx_v2f[j], ev[j], W, bias, None, None, extension=0, aggregator=0, plan=sp)
                assert torch.equal(a, b)
            sp = plan.source_plan("f2v", j, T)
            if sp is not None:
                used += 1
                raws = []
                for pl in (None, sp):
                    raw = torch.full((1, types[0].n_vars, O), float("-inf"), device=dev)
                    fgnn_b200.mp_forward(nm(xf_loc[j]), plan.idx_f2v[j], ef[j], W, None, None, None, extension=0, aggregator=0,
                                         activation=_lib.ACT_NONE, mask_negative=True, out=nm(raw), tile_slots=plan.tile_slots[j],
                                         out_rows=plan.out_rows[j], plan=pl)
                    raws.append(raw)
                assert torch.equal(raws[0], raws[1])
                if rank == 0 and plan.f2v[j].slot_pad.any():
                    assert int((sp.slot_edge == -2).sum()) == int(plan.f2v[j].slot_pad.sum())
    assert used >= 2


# ---------------------------------------------------------------------------------------------
# factor-sharded layer: N-GPU result == 1-GPU result, ranks simulated on one device
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("kernel", ["auto", "simt"])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_plan_equals_single_gpu(world, kernel):
    from fgnn_b200 import parallel
    rng = np.random.default_rng(5)
    C = O = 64
    T = 4
    types = graphs.synthetic_map_graph(3000, 9000, 1500, 3, seed=9)
    J = len(types)
    dev = torch.device(DEV)
    x_v = t(rng.random((1, types[0].n_vars, C), dtype=np.float32))
    x_f = [t(np.abs(rng.standard_normal((1, ty.n_factors, C))).astype(np.float32)) for ty in types]
    et_v2f = [t(rng.standard_normal((1, T, ty.n_factors, ty.order)).astype(np.float32)) for ty in types]
    et_f2v = []
    for ty in types:
        e = rng.standard_normal((1, T, ty.n_vars, ty.kv)).astype(np.float32)
        e[np.broadcast_to(ty.pad_f2v[None, None], e.shape)] = 0.0
        et_f2v.append(t(e))
    W = []
    for _ in range(J):
        d = {}
        for direction in ("v2f", "f2v"):
            d[direction] = dict(filters=t((rng.uniform(-1, 1, (C, O * T)) * 0.1).astype(np.float32)),
                                bias=t(rng.uniform(-0.2, 0.2, O).astype(np.float32)),
                                scale=t(rng.uniform(0.8, 1.2, O).astype(np.float32)),
                                shift=t(rng.uniform(-0.2, 0.2, O).astype(np.float32)))
        W.append(d)
    kern = KERNELS[kernel]
    nm = lambda a: a.permute(0, 2, 1).unsqueeze(-1)
    # single GPU: the unsharded reference tables
    ref_v = torch.zeros_like(x_v)
    ref_f = []
    for j, ty in enumerate(types):
        w = W[j]
        ref_f.append(fgnn_b200.mp_forward(nm(x_v), t(ty.idx_v2f[None]), et_v2f[j], w["v2f"]["filters"], w["v2f"]["bias"],
                                          w["v2f"]["scale"], w["v2f"]["shift"], extension=0, aggregator=0, kernel=kern))
        fgnn_b200.mp_forward(nm(x_f[j]), t(ty.idx_f2v[None]), et_f2v[j], w["f2v"]["filters"], w["f2v"]["bias"],
                             w["f2v"]["scale"], w["f2v"]["shift"], extension=0, aggregator=0, kernel=kern,
                             out=nm(ref_v), accumulate=j > 0)
    # `world` ranks, one after the other on this device; the all-reduce(MAX) is the elementwise maximum
    raws, facs, plans = [], [], []
    for rank in range(world):
        plan = parallel.ShardedLayerPlan(types, rank, world, dev)
        xf_loc = plan.local_factor_features(x_f)
        ev, ef = plan.local_etypes(et_v2f, et_f2v)
        out_v = torch.empty_like(x_v)
        out_f = [torch.empty_like(a) for a in xf_loc]
        plan.layer(x_v, xf_loc, ev, ef, W, out_v, out_f, kern)      # no process group: reduce is done below
        raws.append(plan.raw.clone())
        facs.append(out_f)
        plans.append(plan)
    raw = raws[0]
    for r in raws[1:]:
        raw = torch.maximum(raw, r)
    assert torch.isfinite(raw).all()
    got_v = plans[0].finish(raw, W, torch.empty_like(x_v))
    assert torch.equal(got_v, ref_v), "sharded F->V differs from the single-GPU result"
    for j in range(J):
        cat = torch.cat([f[j] for f in facs], 1)
        assert torch.equal(nm(cat), ref_f[j]), "sharded V->F differs from the single-GPU result"
    # work really is sharded: live slots per rank add up to the table size
    for j, ty in enumerate(types):
        assert sum(p.f2v[j].live_slots for p in plans) == ty.idx_f2v.size


@pytest.mark.parametrize("world,band,dtype", [(2, 0, "f32"), (3, 64, "f32"), (4, 64, "bf16"), (2, 0, "bf16")])
def test_halo_sharded_layers_equal_single_gpu(world, band, dtype):
    """Owner-computes sharding with feature halos (parallel.HaloLayerPlan): `world` ranks simulated on one
    device (their arenas addressed directly, their kernels on separate streams) run three layers; the owned rows of
    every rank, put together, are BIT-IDENTICAL to the single-GPU layers -- fp32 and bf16 I/O (cfg 4 at small scale),
    uniform-random and banded incidence with the locality order."""
    from fgnn_b200 import parallel
    rng = np.random.default_rng(world * 10 + band)
    types = graphs.synthetic_map_graph(3000, 9000, 1500, 4, seed=3, local_band=band)
    if band:
        types = graphs.locality_order(types)
    C, T, L, J = 64, 16, 3, len(types)
    td = torch.bfloat16 if dtype == "bf16" else torch.float32
    fd = lambda a: t(a).to(td)
    x_v = fd(rng.random((1, 3000, C), dtype=np.float32))
    x_f = [fd(np.abs(rng.standard_normal((1, ty.n_factors, C))).astype(np.float32)) for ty in types]
    et_v2f = [fd(rng.standard_normal((1, T, ty.n_factors, ty.order)).astype(np.float32)) for ty in types]
    et_f2v = []
    for ty in types:
        e = rng.standard_normal((1, T, ty.n_vars, ty.kv)).astype(np.float32)
        e[np.broadcast_to(ty.pad_f2v[None, None], e.shape)] = 0.0
        et_f2v.append(fd(e))
    W = [[{d: dict(filters=t(rng.uniform(-0.05, 0.05, (C, C * T)).astype(np.float32)), bias=t(rng.uniform(0, 0.05, C).astype(np.float32)),
                   scale=t(rng.uniform(0.8, 1.2, C).astype(np.float32)), shift=t(rng.uniform(-0.1, 0.1, C).astype(np.float32)))
           for d in ("v2f", "f2v")} for _ in range(J)] for _ in range(L)]
    nm = lambda a: a.permute(0, 2, 1).unsqueeze(-1)
    # single GPU
    cv, cf = x_v, x_f
    for l in range(L):
        nv = torch.empty_like(cv)
        nf = [torch.empty_like(f) for f in cf]
        for j, ty in enumerate(types):
            w = W[l][j]
            fgnn_b200.mp_forward(nm(cv), t(ty.idx_v2f[None]), et_v2f[j], w["v2f"]["filters"], w["v2f"]["bias"], w["v2f"]["scale"],
                                 w["v2f"]["shift"], extension=0, aggregator=0, out=nm(nf[j]))
            fgnn_b200.mp_forward(nm(cf[j]), t(ty.idx_f2v[None]), et_f2v[j], w["f2v"]["filters"], w["f2v"]["bias"], w["f2v"]["scale"],
                                 w["f2v"]["shift"], extension=0, aggregator=0, out=nm(nv), accumulate=j > 0)
        cv, cf = nv, nf
    torch.cuda.synchronize()
    # `world` ranks on this device
    plans = [parallel.HaloLayerPlan(types, r, world, DEV, td, C, ctas=8) for r in range(world)]
    infos = [p.info() for p in plans]
    for p in plans:
        p.connect(infos, same_process=True)
    streams = [torch.cuda.Stream(device=DEV) for _ in range(world)]
    ets = [p.local_etypes(et_v2f, et_f2v) for p in plans]
    try:
        for p in plans:
            p.load_features(x_v, x_f)
        torch.cuda.synchronize()
        outs = [None] * world
        for l in range(L):
            for r, p in enumerate(plans):
                with torch.cuda.stream(streams[r]):
                    outs[r] = p.layer(l, ets[r][0], ets[r][1], W[l], last=(l == L - 1))
        torch.cuda.synchronize()
        got = torch.cat(outs, dim=1)
        assert torch.equal(got, cv), f"max |diff| = {float((got.float() - cv.float()).abs().max())}"
        if band:                                             # locality: the halos are a small part of the rows
            assert all(len(p.var_halo) < 0.2 * p.n_own_v for p in plans)
    finally:
        for p in plans:
            p.close()


@pytest.mark.parametrize("layout", ["channels_first", "channels_last"])
def test_sharded_instance_norm_matches_full(layout):
    """InstanceNorm + ReLU with the nodes of every instance split over three shards (two-pass statistics, partial sums
    added across the shards = the 2 x C-float all-reduce of a sharded FactorNN layer) == the norm of the whole tensor."""
    from fgnn_b200 import parallel
    torch.manual_seed(1)
    B, C, N = 3, 64, 1000
    x = torch.randn(B, C, N, 1, device=DEV) * 2 + 0.5
    if layout == "channels_last":
        x = x.contiguous(memory_format=torch.channels_last)
    want = torch.relu(torch.nn.functional.instance_norm(x))
    cuts = [0, 250, 700, N]
    shards = [x[:, :, a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    # emulate the collective: the partial sums of every shard are computed up front through the C ABI, and each shard's
    # call gets a reducer that returns the total of the pass it is in (first call: sum x, second: sum (x - mean)^2)
    lib, p = _lib.lib(), (lambda tt: ctypes.c_void_p(tt.data_ptr()))
    st = ctypes.c_void_p(torch.cuda.current_stream(DEV).cuda_stream)

    def partials(mean):
        out = [torch.empty(B, C, device=DEV) for _ in shards]
        for s_, sh in zip(out, shards):
            _lib.check(lib.fgnn_instance_norm_partial(p(sh), p(mean) if mean is not None else None, p(s_), B, C, sh.shape[2],
                                                      sh.stride(0), sh.stride(1), sh.stride(2), st), "partial")
        return sum(out)
    tot1 = partials(None)
    tot2 = partials(tot1 / N)
    outs = []
    for sh in shards:
        seq = iter([tot1, tot2])
        outs.append((sh, lambda tt, seq=seq: next(seq).clone()))
    got = torch.cat([parallel.sharded_instance_norm_act(sh, N, activation="relu", reduce=red) for sh, red in outs], dim=2)
    assert_close(got.cpu().numpy(), want.cpu().numpy(), 2e-5, "sharded instance norm")


def test_peer_exchange_two_ranks_on_one_device():
    """fgnn_exchange_forward (max over ranks + per-type epilogue + sum over types + broadcast, one kernel over peer
    memory): two ranks' arenas on this device, their kernels running concurrently on two streams, against
    torch.maximum + fgnn_epilogue_sum_forward.  Three epochs exercise the flag protocol and the ping-pong buffers."""
    from fgnn_b200.parallel import PeerExchange
    dev = torch.device(DEV)
    N, J, O = 5003, 2, 64
    rng = np.random.default_rng(31)
    px = [PeerExchange(N, J, O, r, 2, dev, peers=[0, 0]) for r in range(2)]
    try:
        bases = [px[0].base, px[1].base]
        for q in px:
            q.set_peers(bases)
        bias, scale, shift = (t(rng.uniform(-0.2, 0.2, J * O).astype(np.float32)), t(rng.uniform(0.8, 1.2, J * O).astype(np.float32)),
                              t(rng.uniform(-0.2, 0.2, J * O).astype(np.float32)))
        streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        for epoch in range(3):
            dst = epoch & 1
            raws = []
            for r in range(2):
                a = rng.standard_normal((1, N, J * O)).astype(np.float32)
                a[:, rng.random(N) < 0.4] = -np.inf                      # rows this rank's shard does not touch
                a[:, :7] = -np.inf                                        # rows nobody touches
                px[r].raw.copy_(t(a))
                raws.append(t(a))
            torch.cuda.synchronize()
            raw_mask = out_mask = None
            if epoch == 2:            # static shard sparsity: untouched rows are not read, unneeded rows not written
                bits = np.zeros(N, dtype=np.uint32)
                for r in range(2):
                    fin = np.isfinite(raws[r].cpu().numpy()[0, :, 0])
                    for j in range(J):
                        bits[fin] |= np.uint32(1 << (r * J + j))
                raw_mask = t(bits.view(np.int32))
                want = rng.integers(0, 4, N).astype(np.uint32)
                out_mask = t(want.view(np.int32))
                for r in range(2):
                    px[r].xv[dst].fill_(-7.0)
                torch.cuda.synchronize()
            for r in range(2):
                px[r].forward(dst, bias, scale, shift, _lib.ACT_RELU, 0.0, stream=streams[r], raw_mask=raw_mask, out_mask=out_mask)
            torch.cuda.synchronize()
            red = torch.maximum(raws[0], raws[1]).contiguous()
            ref = torch.empty((1, N, O), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().fgnn_epilogue_sum_forward(red.data_ptr(), ref.data_ptr(), N, O, J, bias.data_ptr(), scale.data_ptr(),
                                                            shift.data_ptr(), _lib.ACT_RELU, 0.0, 0,
                                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "epilogue_sum")
            torch.cuda.synchronize()
            for r in range(2):
                if out_mask is None:
                    assert torch.equal(px[r].xv[dst], ref), f"epoch {epoch} rank {r}"
                else:                 # a rank holds the rows it asked for and the rows it owns; the rest was left alone
                    rows = np.arange(N)
                    owned = (rows >= px[r].row0) & (rows < px[r].row1)
                    got = ((want >> r) & 1).astype(bool) | owned
                    gi, ni = torch.from_numpy(np.nonzero(got)[0]).to(dev), torch.from_numpy(np.nonzero(~got)[0]).to(dev)
                    assert torch.equal(px[r].xv[dst][:, gi], ref[:, gi]), f"masked epoch rank {r}"
                    assert bool((px[r].xv[dst][:, ni] == -7.0).all())
            assert torch.isinf(ref[:, :7]).all() and torch.isfinite(ref[:, 7:]).any()
    finally:
        for q in px:
            q.close()


@pytest.mark.parametrize("shape", [(5, 64, 96), (3, 130, 48), (2, 7, 1000), (4, 256, 33)])
@pytest.mark.parametrize("layout", ["channels_first", "channels_last"])
def test_instance_norm_act_matches_torch(shape, layout):
    """fgnn_instance_norm_forward (the InstanceNorm + ReLU of FactorNN's v2v / f2f maps, base_model.py:83-90)
    against torch.nn.functional.instance_norm + relu, both memory formats."""
    from fgnn_b200.factor_nn import instance_norm_act
    B, C, N = shape
    x = torch.randn(B, C, N, 1, device=DEV) * 3 + 1
    if layout == "channels_last":
        x = x.contiguous(memory_format=torch.channels_last)
    ref = torch.relu(torch.nn.functional.instance_norm(x, eps=1e-5))
    before = fgnn_b200.launch_count()
    y = instance_norm_act(x, 1e-5, "relu")
    assert fgnn_b200.launch_count() == before + 1 and y.stride() == x.stride()
    assert_close(y.cpu().numpy(), ref.cpu().numpy(), 1e-5, f"instance norm {shape} {layout}")
    y0 = instance_norm_act(x, 1e-5, None)
    assert_close(y0.cpu().numpy(), torch.nn.functional.instance_norm(x, eps=1e-5).cpu().numpy(), 1e-5, "no activation")


@pytest.mark.parametrize("cin,cout", [(64, 64), (64, 128), (128, 64), (64, 256), (128, 256), (256, 64), (256, 128), (256, 256)])
def test_conv1x1_native_matches_pytorch(cin, cout):
    """mp_nn.conv1x1_native: a 1x1 map + bias + folded eval BatchNorm + LeakyReLU as one launch of the tensor-core
    kernel (identity table, one slot, one edge type == 1) against Conv2d -> BatchNorm2d -> LeakyReLU in fp32."""
    from fgnn_b200.mp_nn import conv1x1_native, _fold_bn
    torch.manual_seed(cin + cout)
    seq = torch.nn.Sequential(torch.nn.Conv2d(cin, cout, 1), torch.nn.BatchNorm2d(cout), torch.nn.LeakyReLU()).to(DEV).eval()
    seq[1].running_mean.uniform_(-0.3, 0.3)
    seq[1].running_var.uniform_(0.5, 1.5)
    seq[1].weight.data.uniform_(0.5, 1.5)
    seq[1].bias.data.uniform_(-0.2, 0.2)
    x = torch.randn(37, cin, 129, 1, device=DEV).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        ref = seq(x)
        scale, shift = _fold_bn(seq[1])
        before = fgnn_b200.launch_count()
        y = conv1x1_native(x, seq[0].weight, seq[0].bias, scale, shift, _lib.ACT_LEAKY_RELU, 0.01)
        assert y is not None and fgnn_b200.launch_count() > before and y.stride(1) == 1
        assert_close(y.cpu().numpy(), ref.cpu().numpy(), RTOL, f"native 1x1 map {cin}->{cout}")
        raw = conv1x1_native(x, seq[0].weight, seq[0].bias)
        assert_close(raw.cpu().numpy(), seq[0](x).cpu().numpy(), RTOL, "plain map")
        assert conv1x1_native(x[:, :, :3], seq[0].weight) is None          # too small / not node-major: caller falls back
        # fused `acc = acc + map(x)` (FactorNN's nfeature + f2v(...)): the map adds into the caller's tensor in its store
        acc = torch.randn(37, cout, 129, 1, device=DEV).contiguous(memory_format=torch.channels_last)
        want = acc + ref
        got = conv1x1_native(x, seq[0].weight, seq[0].bias, scale, shift, _lib.ACT_LEAKY_RELU, 0.01, out=acc, accumulate=True)
        assert got is acc
        assert_close(acc.cpu().numpy(), want.cpu().numpy(), RTOL, "accumulating map")


def test_factornn_wide_layers_fused_adds_match_pytorch_wrappers():
    """train_ldpc.py's wide layers (128 -> 256 -> 256 -> 128): the C = 256 maps run on the tensor-core kernel as two-slot /
    two-type calls, `nfeature + f2v(...)` is fused into the last kernel's store and the classifier head's
    Conv -> InstanceNorm -> ReLU runs natively.  Reference: the same modules with every wrapper on PyTorch ops (the cores
    stay native), i.e. factor_mpnn_sp.py:136-176 as written."""
    from fgnn_b200 import factor_nn, mp_nn
    torch.manual_seed(5)
    B = 48
    g = load_npz("ldpc_factornn.npz")
    model = fgnn_b200.FactorNN(2, [6, 96], [64, 128, 256, 256, 128], [4, 1], 2, skip_link={3: 0}).to(DEV).eval()
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.uniform_(-0.2, 0.2)
            m.running_var.uniform_(0.6, 1.4)
    rng = np.random.default_rng(3)
    node = t(rng.standard_normal((B, 2, 96, 1)).astype(np.float32))
    hop = t(rng.standard_normal((B, 6, 48, 1)).astype(np.float32))
    nhop = node[:, 0, :, :].reshape(B, 96, 1, 1)
    rep = lambda a: t(a)[None].repeat(B, 1, 1)
    args = (node, [hop, nhop], [rep(g["idx_f2v"]), torch.zeros(B, 96, 1, dtype=torch.long, device=DEV)],
            [rep(g["idx_v2f"]), torch.arange(96, device=DEV).reshape(1, 1, 96).repeat(B, 1, 1)],
            [t(rng.standard_normal((B, 4, 96, 3)).astype(np.float32)), torch.ones(B, 1, 96, 1, device=DEV)],
            [t(rng.standard_normal((B, 4, 48, 6)).astype(np.float32)), torch.ones(B, 1, 1, 96, device=DEV)])
    with torch.no_grad():
        before = fgnn_b200.launch_count()
        got = model(*args)
        native_launches = fgnn_b200.launch_count() - before
        # the same forward with the wrappers' native routes switched off
        saved = (mp_nn.conv1x1_native, factor_nn.conv1x1_native, factor_nn.conv_in_relu, factor_nn.FactorNN._add_core)
        try:
            mp_nn.conv1x1_native = factor_nn.conv1x1_native = lambda *a, **k: None
            factor_nn.conv_in_relu = lambda *a, **k: None
            factor_nn.FactorNN._add_core = staticmethod(
                lambda acc, m, x, idx, ef: acc + (m(x, idx, ef) if isinstance(m, fgnn_b200.base_mp_nn) else m(x)))
            before = fgnn_b200.launch_count()
            ref = model(*args)
            plain_launches = fgnn_b200.launch_count() - before
        finally:
            mp_nn.conv1x1_native, factor_nn.conv1x1_native, factor_nn.conv_in_relu, factor_nn.FactorNN._add_core = saved
    assert native_launches > plain_launches                      # the maps really ran on the library's kernels
    assert_close(got.cpu().numpy(), ref.cpu().numpy(), RTOL, "FactorNN with wide layers")

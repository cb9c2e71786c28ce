"""The CPU oracle (numpy restatement + C restatement) against the golden vectors produced by the
real reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import fgnn_oracle as orc
from tests.util import assert_close, bn_of, load_npz, load_unit_cases, sub_sd

CASES = load_unit_cases()
# the oracle's summation order differs from ATen's mm/bmm; fp32 round-off only
ORACLE_RTOL = 2e-5


@pytest.mark.parametrize("case", CASES, ids=[c["meta"]["name"] for c in CASES])
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_unit_cases(case, impl):
    m, sd = case["meta"], case["sd"]
    fn = orc.mp_conv_forward if impl == "numpy" else orc.mp_conv_forward_c
    got = fn(case["x"], case["idx"], case["etype"], sd["filters"], sd.get("bias"), bn_of(sd),
             extension=m["ext"], aggregator=m["agg"], activation=m.get("act", "relu"))
    assert_close(got, case["out"], ORACLE_RTOL, m["name"])


@pytest.mark.parametrize("case", CASES, ids=[c["meta"]["name"] for c in CASES])
@pytest.mark.parametrize("chunk", [0, 3])
def test_torch_port_unit_cases(case, chunk):
    """The PyTorch restatement in the reference's op order (the `port-torch` CPU baseline and the ATen-on-GPU
    baseline of bench.py) against the same reference outputs; chunking over destinations is bit-identical."""
    import torch
    from oracle import fgnn_oracle_torch as orct
    m, sd = case["meta"], case["sd"]
    t = torch.from_numpy
    bn = bn_of(sd)
    bn = {k: t(v) for k, v in bn.items()} if bn is not None else None
    got = orct.mp_conv_forward_torch(t(case["x"]), t(case["idx"]), t(case["etype"]), t(sd["filters"]),
                                     t(sd["bias"]) if "bias" in sd else None, bn, extension=m["ext"], aggregator=m["agg"],
                                     activation=m.get("act", "relu"), chunk_rows=chunk)
    assert_close(got.numpy(), case["out"], 1e-6, m["name"])
    if chunk:
        whole = orct.mp_conv_forward_torch(t(case["x"]), t(case["idx"]), t(case["etype"]), t(sd["filters"]),
                                           t(sd["bias"]) if "bias" in sd else None, bn, extension=m["ext"],
                                           aggregator=m["agg"], activation=m.get("act", "relu"))
        assert torch.equal(got, whole)


def test_out_of_range_index_raises():
    c = CASES[0]
    bad = c["idx"].copy()
    bad[0, 0, 0] = c["x"].shape[2]
    for fn in (orc.mp_conv_forward, orc.mp_conv_forward_c):
        with pytest.raises(IndexError):
            fn(c["x"], bad, c["etype"], c["sd"]["filters"], extension=c["meta"]["ext"], aggregator="max")


def test_cfg1_simple_gnn_chain():
    """BASELINE configs[0]: first layer (NEIGHBOR, softmax), the residual FGNN layer (DIFF, max),
    1x1 classifier; MAP labels bit-exact."""
    g = load_npz("cfg1_simple_gnn.npz")
    sd = sub_sd(g, "model")
    B = g["x"].shape[0]
    idx = np.repeat(g["nn_idx"], B, 0)
    et = np.repeat(g["etype"], B, 0)
    s0 = {k[2:]: v for k, v in sd.items() if k.startswith("0.")}
    h1 = orc.mp_conv_forward(g["x"], idx, et, s0["filters"], s0["bias"], bn_of(s0),
                             extension=orc.ORIG_WITH_NEIGHBOR, aggregator="softmax")
    assert_close(h1, g["h1"], ORACLE_RTOL, "h1")
    s1 = {k[2:]: v for k, v in sd.items() if k.startswith("1.")}
    h2 = orc.mp_conv_residual_forward(s1, h1, idx, et, orc.ORIG_WITH_DIFF, "max", True)
    assert_close(h2, g["h2"], ORACLE_RTOL, "h2")
    logits = orc.conv1x1(h2, sd["2.weight"], sd["2.bias"])
    assert_close(logits, g["logits"], ORACLE_RTOL, "logits")
    labels = logits[..., 0].argmax(1)
    assert np.array_equal(labels, g["labels"])
    assert 0 < labels.sum() < labels.size          # the check is not vacuous


def test_cfg1_factornn():
    g = load_npz("cfg1_factornn.npz")
    sd = sub_sd(g, "model")
    B = g["node"].shape[0]
    logit = orc.factor_nn_forward(sd, g["node"], [g["hop"]], [np.repeat(g["idx_f2v"], B, 0)],
                                  [np.repeat(g["idx_v2f"], B, 0)], [g["et_f2v"]], [g["et_v2f"]], [64, 64])
    assert_close(logit, g["logit"], 5e-5, "logit")
    assert np.array_equal(logit >= 0, g["decision"])


def test_ldpc_factornn():
    g = load_npz("ldpc_factornn.npz")
    sd = sub_sd(g, "model")
    B = g["node"].shape[0]
    rep = lambda a: np.repeat(a[None], B, 0)
    h_idx_v2f = np.tile(np.arange(96).reshape(1, 1, 96), (B, 1, 1))
    h_idx_f2v = np.zeros((B, 96, 1), dtype=np.int64)
    nhop = g["node"][:, 0].reshape(B, 96, 1, 1)
    res, nhops = orc.factor_nn_forward(
        sd, g["node"], [g["hop"], nhop], [rep(g["idx_f2v"]), h_idx_f2v], [rep(g["idx_v2f"]), h_idx_v2f],
        [g["et_f2v"], np.ones((B, 1, 96, 1), np.float32)], [g["et_v2f"], np.ones((B, 1, 1, 96), np.float32)],
        [64, 64, 128, 64], skip_link={2: 0}, ret_high=True)
    res = res + g["node"][:, :1]
    assert_close(res, g["res"], 5e-5, "res")
    assert_close(nhops[0], g["nhop0"], 5e-5, "nhop0")
    assert np.array_equal(res >= 0, g["hard"])


def test_factor_mpnn_merged_tables():
    """The merged-table model of train_syn_hop_factor.py (factor_mpnn.py:88-133): ORIG_WITH_DIFF residual cores and
    the bare 64 -> 2 last layer with the default softmax aggregator; MAP labels bit-exact."""
    g = load_npz("factor_mpnn_merged.npz")
    sd = sub_sd(g, "model")
    B = g["node"].shape[0]
    rep = lambda a: np.repeat(a, B, 0)
    out_v, out_f = orc.factor_mpnn_forward(sd, g["node"], [g["f_pw"], g["f_hi"]],
                                           [(rep(g["idx_pw"]), g["et_pw"]), (rep(g["idx_hi"]), g["et_hi"])], [64, 64, 2])
    assert_close(out_v, g["out_v"], 5e-5, "out_v")
    assert_close(out_f[0], g["out_f0"], 5e-5, "out_f0")
    assert_close(out_f[1], g["out_f1"], 5e-5, "out_f1")
    labels = out_v[..., 0].argmax(1)
    assert np.array_equal(labels, g["labels"]) and 0 < labels.sum() < labels.size


@pytest.mark.parametrize("seed", range(12))
def test_three_oracles_agree_on_seeded_random_shapes(seed):
    """numpy restatement == C restatement == torch restatement (the reference's own ATen op order) on shapes and option
    mixes the goldens do not enumerate: ragged sizes, every extension x aggregator, with / without bias and BN, batched
    tables with repeated and boundary indices.  The goldens pin each of them to the real reference; this pins them to
    each other everywhere else."""
    import torch
    from oracle import fgnn_oracle_torch as orct
    rng = np.random.default_rng(1000 + seed)
    B = int(rng.integers(1, 4))
    C = int(rng.choice([2, 3, 8, 17, 64]))
    O = int(rng.choice([2, 5, 16, 64]))
    T = int(rng.choice([1, 2, 4, 16]))
    ext = int(rng.integers(0, 3))
    N = int(rng.integers(3, 40))
    M = N if ext else int(rng.integers(1, 50))
    K = int(rng.integers(1, 7))
    agg = [None, "max", "softmax", "mean"][int(rng.integers(0, 4))] if seed % 4 else "max"
    x = rng.standard_normal((B, C, N, 1)).astype(np.float32)
    idx = rng.integers(0, N, (B, M, K)).astype(np.int64)
    idx[:, 0, 0] = N - 1                                          # boundary index
    idx[:, -1, :] = idx[:, -1, :1]                                # repeated sources in one row
    et = rng.standard_normal((B, T, M, K)).astype(np.float32)
    if K > 1:
        et[:, :, :, -1] = 0.0                                     # the reference's padding: a zero edge type on a valid index
    W = (rng.uniform(-1, 1, ((2 if ext else 1) * C, O * T)) * 0.3).astype(np.float32)
    bias = rng.uniform(-0.2, 0.2, O).astype(np.float32) if seed % 3 else None
    bn = None
    if seed % 2:
        bn = dict(weight=rng.uniform(0.5, 1.5, O).astype(np.float32), bias=rng.uniform(-0.2, 0.2, O).astype(np.float32),
                  running_mean=rng.uniform(-0.2, 0.2, O).astype(np.float32), running_var=rng.uniform(0.5, 1.5, O).astype(np.float32))
    act = "relu" if seed % 5 else None
    a = orc.mp_conv_forward(x, idx, et, W, bias, bn, extension=ext, aggregator=agg, activation=act)
    c = orc.mp_conv_forward_c(x, idx, et, W, bias, bn, extension=ext, aggregator=agg, activation=act)
    t = torch.from_numpy
    tb = {k: t(v) for k, v in bn.items()} if bn is not None else None
    p = orct.mp_conv_forward_torch(t(x), t(idx), t(et), t(W), t(bias) if bias is not None else None, tb, extension=ext,
                                   aggregator=agg, activation=act).numpy()
    assert a.shape == c.shape == p.shape == (B, O, M, K if agg is None else 1)
    assert_close(c, a, ORACLE_RTOL, f"C vs numpy (seed {seed})")
    assert_close(p, a, ORACLE_RTOL, f"torch vs numpy (seed {seed})")

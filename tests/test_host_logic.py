"""Host-side mirror of the reference interface (no GPU): constructor / state_dict / error
behaviour, graph-table builders, the drop-in install, and the no-fallback rule."""
import io
import contextlib
import os
import sys

import numpy as np
import pytest
import torch

import fgnn_b200
from fgnn_b200 import graphs
from oracle import refload
from tests.util import load_npz, load_unit_cases, sub_sd


def test_constructor_surface_and_state_dict_keys():
    m = fgnn_b200.mp_conv_v2(4, 8, 3)
    assert (m.nin, m.nou, m.nedge_types) == (4, 8, 3)
    assert m.extension == fgnn_b200.mp_conv_type.ORIG_WITH_DIFF          # reference default
    assert m.filters.shape == (8, 24) and m.bias.shape == (8,)
    assert isinstance(m.bn, torch.nn.BatchNorm2d) and isinstance(m.activation_fn, torch.nn.ReLU)
    assert list(m.state_dict().keys()) == ["filters", "bias", "bn.weight", "bn.bias", "bn.running_mean",
                                           "bn.running_var", "bn.num_batches_tracked"]
    assert isinstance(m, fgnn_b200.base_mp_nn) and m.is_mp_nn
    assert float(m.filters.detach().abs().max()) <= 0.01 and 0 <= float(m.bias.detach().min()) and float(m.bias.detach().max()) <= 0.05
    n = fgnn_b200.mp_conv_v2(4, 8, 3, bias=False, bn=False, extension=fgnn_b200.mp_conv_type.NO_EXTENSION,
                             activation_fn=None, aggregtor='max')
    assert n.filters.shape == (4, 24) and n.bias is None and n.bn is None and n.activation_fn is None
    assert list(n.state_dict().keys()) == ["filters"]


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError, match="extension must one of mp_conv_type"):
        fgnn_b200.mp_conv_v2(4, 8, 3, extension=7)
    m = fgnn_b200.mp_conv_v2(4, 8, 3, aggregtor='median')        # unknown string: attribute stays unset
    assert not hasattr(m, "aggregtor")


def test_default_aggregators():
    assert fgnn_b200.mp_conv_v2(2, 2, 1)._agg == fgnn_b200._lib.AGG_SOFTMAX         # mp_nn.py:26
    assert fgnn_b200.mp_conv_residual(4, 4, 2).mp_conv._agg == fgnn_b200._lib.AGG_MAX  # mp_nn_residual.py:15
    v = torch.randn(2, 3, 4, 5)
    m = fgnn_b200.mp_conv_v2(2, 2, 1)
    assert torch.allclose(m.aggregtor(v), torch.logsumexp(3 * v, 3, keepdim=True) / 3)


def test_cpu_tensor_has_no_fallback():
    m = fgnn_b200.mp_conv_v2(4, 8, 2).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(1, 4, 3, 1), torch.zeros(1, 3, 2, dtype=torch.long), torch.zeros(1, 2, 3, 2))


def test_training_forward_needs_cuda_too():
    """Train mode goes through the autograd.Function around the same kernel: still no CPU fallback."""
    m = fgnn_b200.mp_conv_v2(4, 8, 2).train()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 4, 3, 1), torch.zeros(1, 3, 2, dtype=torch.long), torch.zeros(1, 2, 3, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "factor-graph-neural-network_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "fgnn_oracle" not in src, f


def test_state_dicts_load_from_reference_fixtures():
    for c in load_unit_cases():
        m = c["meta"]
        mod = fgnn_b200.mp_conv_v2(m["C"], m["O"], m["T"], bias=m.get("bias", True), bn=m.get("bn", True),
                                   extension=fgnn_b200.mp_conv_type(m["ext"]),
                                   activation_fn=m.get("act", "relu"), aggregtor=m["agg"])
        mod.load_state_dict({k: torch.from_numpy(v) for k, v in c["sd"].items()})
    g = load_npz("cfg1_factornn.npz")
    fnn = fgnn_b200.FactorNN(2, [4], [64, 64], [16], 2)
    fnn.load_state_dict({k: torch.from_numpy(v) for k, v in sub_sd(g, "model").items()})
    g = load_npz("ldpc_factornn.npz")
    fnn = fgnn_b200.FactorNN(2, [6, 96], [64, 64, 128, 64], [4, 1], 2, skip_link={2: 0}, ret_high=True)
    fnn.load_state_dict({k: torch.from_numpy(v) for k, v in sub_sd(g, "model").items()})
    # layer 1 (64 -> 128) is a bare core, layers 0 and 2 are residual wrappers (factor_mpnn_sp.py:79-94)
    assert isinstance(fnn.f2v_modules[1][0], fgnn_b200.mp_conv_v2)
    assert isinstance(fnn.f2v_modules[0][0], fgnn_b200.mp_conv_residual)


def test_chain_knn_table_matches_reference_fixture():
    g = load_npz("cfg1_simple_gnn.npz")
    idx, ef = graphs.chain_knn_table(128, 4)
    assert np.array_equal(idx, g["nn_idx"]) and np.array_equal(ef, g["efeature"])
    assert (idx[0, :, 3] == 0).all() and (ef[0, 0, :, 3] == 0).all()     # the reference's pad slot


def test_synthetic_graph_tables_are_consistent():
    types = graphs.synthetic_map_graph(1000, 3000, 500, 3, seed=5)
    pw, ho = types
    assert pw.idx_v2f.shape == (3000, 2) and pw.idx_f2v.shape == (1000, 6) and not pw.pad_f2v.any()
    assert ho.idx_v2f.shape == (500, 3) and ho.idx_f2v.shape == (1000, 2)
    assert pw.real_messages + ho.real_messages == 2 * (3000 * 2 + 500 * 3)
    for t in types:
        assert t.idx_v2f.min() >= 0 and t.idx_v2f.max() < t.n_vars
        assert t.idx_f2v.min() >= 0 and t.idx_f2v.max() < t.n_factors
        # every non-pad (variable, slot) names a factor that lists the variable, and counts agree
        v, s = np.nonzero(~t.pad_f2v)
        assert (t.idx_v2f[t.idx_f2v[v, s]] == v[:, None]).any(1).all()
        assert (~t.pad_f2v).sum() == t.n_factors * t.order
        assert (t.idx_f2v[t.pad_f2v] == 0).all()
    for loc in graphs.synthetic_map_graph(1000, 3000, 500, 3, seed=5, local_band=16):
        span = (loc.idx_v2f.max(1) - loc.idx_v2f.min(1))
        assert ((span < 32) | (span > 1000 - 32)).all()                      # within two bands (or wrapped around)
        assert (~loc.pad_f2v).sum() == loc.n_factors * loc.order
    pw = graphs.synthetic_map_graph(1000, 3000, 0, 0, seed=5, local_band=16)[0]
    assert pw.kv == 6 and not pw.pad_f2v.any()                                   # degrees as even as the uniform graph's
    # locality order: factors sorted by their smallest variable, tables still consistent
    ordered = graphs.locality_order([pw])[0]
    assert np.all(np.diff(ordered.idx_v2f.min(1)) >= 0)
    v, s_ = np.nonzero(~ordered.pad_f2v)
    assert (ordered.idx_v2f[ordered.idx_f2v[v, s_]] == v[:, None]).any(1).all()


def test_point_cloud_graph_cfg5_shape():
    """BASELINE configs[4] (SURVEY 8d cfg 5): kNN pairwise + patch factors over a Z-ordered point set."""
    types = graphs.point_cloud_graph(4096, 16, 512, 16, seed=1)
    knn, patch = types
    assert knn.idx_v2f.shape == (4096 * 8, 2) and patch.idx_v2f.shape == (512, 16)
    assert knn.real_messages + patch.real_messages == 2 * (4096 * 8 * 2 + 512 * 16)
    assert (knn.idx_v2f[:, 0] != knn.idx_v2f[:, 1]).all()                        # a point is not its own neighbour
    assert (np.diff(np.sort(patch.idx_v2f, 1), axis=1) > 0).all()                # 16 distinct members
    for t in types:
        assert t.idx_v2f.min() >= 0 and t.idx_v2f.max() < 4096
        v, s = np.nonzero(~t.pad_f2v)
        assert (t.idx_v2f[t.idx_f2v[v, s]] == v[:, None]).any(1).all()
        assert (~t.pad_f2v).sum() == t.n_factors * t.order and (t.idx_f2v[t.pad_f2v] == 0).all()
    # the Z-order numbering turns spatial neighbours into index neighbours: half of the kNN pairs span < 1 % of the ids
    assert np.median(np.abs(knn.idx_v2f[:, 0] - knn.idx_v2f[:, 1])) < 41
    a = graphs.point_cloud_graph(4096, 16, 512, 16, seed=1)[0].idx_v2f
    assert np.array_equal(a, knn.idx_v2f)                                        # seeded


def test_parse_alist_small():
    text = "4 2\n2 3\n1 2 1 2\n3 3\n1 0\n1 2\n2 0\n1 2\n1 2 4\n2 4 0\n"
    v2c, c2v = graphs.parse_alist(text)
    assert v2c.tolist() == [[0, -1], [0, 1], [1, -1], [0, 1]]
    assert c2v.tolist() == [[0, 1, 3], [1, 3, -1]]


@pytest.mark.skipif(not refload.available(), reason="reference tree only exists in the build container")
def test_parse_alist_matches_reference_ldpc_tables():
    g = load_npz("ldpc_factornn.npz")
    text = open(os.path.join(refload.REF_ROOT, "ldpc_codes", "96.3.963", "96.3.963")).read()
    v2c, c2v = graphs.parse_alist(text)
    assert np.array_equal(v2c, g["idx_f2v"]) and np.array_equal(c2v, g["idx_v2f"])


@pytest.mark.skipif(not refload.available(), reason="reference tree only exists in the build container")
def test_install_drops_into_reference_package():
    mpnn = refload.load()
    orig = mpnn.mp_conv_v2
    import sys
    mods = {n: sys.modules["lib.model.mpnn." + n] for n in ("mp_nn", "mp_nn_residual", "factor_mpnn_sp", "factor_mpnn")}
    saved = {n: m.mp_conv_v2 for n, m in mods.items()}
    with contextlib.redirect_stdout(io.StringIO()):
        ref_keys = list(orig(8, 8, 2).state_dict().keys())     # before patching: the class finds itself by name
    try:
        native = fgnn_b200.install(mpnn)
        with contextlib.redirect_stdout(io.StringIO()):
            res = mpnn.mp_conv_residual(8, 8, 2)                        # reference wrapper, native core
            fnn = mpnn.FactorNN(2, [4], [16, 16], [2], 2)
        assert isinstance(res.mp_conv, native) and isinstance(res.mp_conv, fgnn_b200.mp_conv_v2)
        assert isinstance(res.mp_conv, sys.modules["lib.model.mpnn.base_model"].base_mp_nn)      # reference isinstance dispatch
        assert isinstance(fnn.f2v_modules[0][0].mp_conv, native)
        assert res.mp_conv.extension == mpnn.mp_conv_type.ORIG_WITH_DIFF
        assert list(res.mp_conv.state_dict().keys()) == ref_keys
    finally:
        mpnn.mp_conv_v2 = orig
        for n, v in saved.items():
            mods[n].mp_conv_v2 = v


def _oracle_core(x, nn_idx, etype, filters, bias=None, bn_scale=None, bn_shift=None, *, extension=0, aggregator=0,
                 activation=1, act_slope=0.01, gamma=3.0, **_ignored):
    """CPU stand-in for fgnn_b200.mp_nn.mp_forward in host-logic tests: the same call contract (folded BN as
    scale/shift, enum ints, [B,O,M,Kout] result) evaluated by the numpy oracle."""
    import torch
    from oracle import fgnn_oracle as orc
    assert isinstance(extension, int) and isinstance(aggregator, int) and isinstance(activation, int)
    y = orc.mp_conv_forward(x.detach().numpy(), nn_idx.numpy(), etype.detach().numpy(), filters.detach().numpy(),
                            None, None, extension=extension, aggregator=aggregator, activation=None, gamma=gamma)
    y = torch.from_numpy(y)
    if bias is not None:
        y = y + bias.detach().view(1, -1, 1, 1)
    if bn_scale is not None:
        y = y * bn_scale.view(1, -1, 1, 1) + bn_shift.view(1, -1, 1, 1)
    if activation == 1:
        y = torch.relu(y)
    elif activation == 2:
        y = torch.nn.functional.leaky_relu(y, act_slope)
    return y


@pytest.mark.skipif(not refload.available(), reason="reference tree only exists in the build container")
def test_installed_module_forward_through_reference_wrappers(monkeypatch):
    """After install() the REFERENCE's own mp_conv_residual / FactorNN / mp_sequential run a forward through
    fgnn_b200.mp_conv_v2.forward: enum handling (the reference's mp_conv_type lands on .extension), argument
    marshalling, BN folding and the epilogue selection are exercised on CPU with the device call replaced by the
    numpy oracle, and the result must equal the unpatched reference module with the same weights."""
    import sys
    import torch
    from fgnn_b200 import mp_nn as native_mp_nn
    mpnn = refload.load()
    orig = mpnn.mp_conv_v2
    mods = {n: sys.modules["lib.model.mpnn." + n] for n in ("mp_nn", "mp_nn_residual", "factor_mpnn_sp", "factor_mpnn")}
    saved = {n: m.mp_conv_v2 for n, m in mods.items()}
    torch.manual_seed(3)
    B, N, K, T = 2, 12, 3, 4

    def randomise(m):
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.uniform_(-0.2, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
        for name, p in m.named_parameters():
            if name.endswith("filters"):
                p.data.uniform_(-0.3, 0.3)

    with contextlib.redirect_stdout(io.StringIO()):
        ref_res = mpnn.mp_conv_residual(8, 8, T)                                    # ORIG_WITH_DIFF inside
        ref_fnn = mpnn.FactorNN(2, [4], [16, 16], [T], 2)
        ref_v2 = mpnn.mp_conv_v2(8, 6, T, extension=mpnn.mp_conv_type.ORIG_WITH_NEIGHBOR)   # softmax default
    for m in (ref_res, ref_fnn, ref_v2):
        randomise(m)
        m.eval()
    x = torch.randn(B, 8, N, 1)
    idx = torch.randint(0, N, (B, N, K))
    et = torch.randn(B, T, N, K)
    nf, hf = torch.randn(B, 2, N, 1), [torch.randn(B, 4, 7, 1)]
    idx_f2v, idx_v2f = [torch.randint(0, 7, (B, N, 2))], [torch.randint(0, N, (B, 7, 3))]
    et_f2v, et_v2f = [torch.randn(B, T, N, 2)], [torch.randn(B, T, 7, 3)]
    with torch.no_grad():
        want_res = ref_res(x, idx, et)
        want_v2 = ref_v2(x, idx, et)
        want_fnn = ref_fnn(nf, hf, idx_f2v, idx_v2f, et_f2v, et_v2f)
    try:
        native = fgnn_b200.install(mpnn)
        monkeypatch.setattr(native_mp_nn, "mp_forward", _oracle_core)
        with contextlib.redirect_stdout(io.StringIO()):
            got_res = mpnn.mp_conv_residual(8, 8, T)
            got_fnn = mpnn.FactorNN(2, [4], [16, 16], [T], 2)
            got_v2 = mpnn.mp_conv_v2(8, 6, T, extension=mpnn.mp_conv_type.ORIG_WITH_NEIGHBOR)
        assert isinstance(got_res.mp_conv, native) and isinstance(got_fnn.v2f_modules[0][0].mp_conv, native)
        for g, r in ((got_res, ref_res), (got_fnn, ref_fnn), (got_v2, ref_v2)):
            g.load_state_dict(r.state_dict())
            g.eval()
        with torch.no_grad():
            assert torch.allclose(got_res(x, idx, et), want_res, rtol=1e-5, atol=1e-6)
            assert torch.allclose(got_v2(x, idx, et), want_v2, rtol=1e-5, atol=1e-6)
            assert torch.allclose(got_fnn(nf, hf, idx_f2v, idx_v2f, et_f2v, et_v2f), want_fnn, rtol=1e-4, atol=1e-5)
            # int extension values (base_mp_nn.NO_EXTENSION style) are accepted too
            got_v2.extension = 1
            assert torch.allclose(got_v2(x, idx, et), want_v2, rtol=1e-5, atol=1e-6)
    finally:
        mpnn.mp_conv_v2 = orig
        for n, v in saved.items():
            mods[n].mp_conv_v2 = v


@pytest.mark.parametrize("row_cap", [3, 6, None])
def test_source_plan_layout(row_cap):
    """SourcePlan (host side of the source-stationary path): edges = live slots, numbered virtual row by virtual
    row; a virtual row is a source row with at most row_cap of its edges (slot order), overflow rows appended."""
    import torch
    import fgnn_b200
    rng = np.random.default_rng(0)
    B, N, M, K = 2, 7, 23, 3
    idx = rng.integers(-1, N, (B, M, K))
    idx[0, :, 0] = 2                                            # a hub: 23+ edges on source row 2 of batch 0
    plan = fgnn_b200.SourcePlan(torch.from_numpy(idx), N, mask_negative=True, row_cap=row_cap)
    cap = plan.row_cap
    assert cap in (3, 6) and (row_cap is None or cap == row_cap)
    live = idx >= 0
    R = B * N
    assert plan.n_edges == int(live.sum())
    ptr, slot_edge, edge_slot, xrow = plan.src_ptr.numpy(), plan.slot_edge.numpy(), plan.edge_slot.numpy(), plan.src_rows.numpy()
    V = plan.n_rows
    assert ptr.shape == (V + 1,) and xrow.shape == (V - R,)
    assert ptr[0] == 0 and ptr[-1] == plan.n_edges and np.all(np.diff(ptr) >= 0) and np.all(np.diff(ptr) <= cap)
    flat_src = (idx + np.arange(B)[:, None, None] * N).reshape(-1)
    seen = {g: [] for g in range(R)}
    for v in range(V):
        g = v if v < R else int(xrow[v - R])
        slots = edge_slot[ptr[v]:ptr[v + 1]]
        assert np.all(flat_src[slots] == g) and np.all(live.reshape(-1)[slots])
        if v >= R:
            assert len(slots) > 0                                # extra rows exist only for overflow edges
        seen[g].append((v, list(slots)))
    for g in range(R):
        want = np.nonzero((flat_src == g) & live.reshape(-1))[0]
        got = [s for _, sl in sorted(seen[g]) for s in sl]
        assert got == list(want)                                # slot order within a source, first row_cap on the row itself
        assert len(seen[g][0][1]) == min(cap, len(want))
    assert np.all(np.diff(xrow) >= 0)
    assert np.array_equal(slot_edge[edge_slot], np.arange(plan.n_edges))
    assert np.all(slot_edge[~live.reshape(-1)] == -1)
    # cached per table object
    tbl = torch.from_numpy(idx)
    assert fgnn_b200.SourcePlan.for_table(tbl, N, True) is fgnn_b200.SourcePlan.for_table(tbl, N, True)


@pytest.mark.parametrize("row_cap", [3, 6, None])
def test_native_plan_builder_matches_torch_builder(row_cap):
    """fgnn_plan_build_host (csrc/plan.cu, counting sort on the host) == the torch-op builder, array for array; it
    needs no GPU, so it is what SourcePlan uses for tables that still live on the host."""
    import torch
    import fgnn_b200
    rng = np.random.default_rng(3)
    B, N, M, K = 3, 40, 90, 4
    idx = rng.integers(-1, N, (B, M, K))
    idx[1, :, 1] = 5                                            # a hub
    t64 = torch.from_numpy(idx)
    native = fgnn_b200.SourcePlan(t64, N, row_cap=row_cap)                       # host table -> native builder

    class _Dev(torch.Tensor):                                   # the same table posing as a device tensor: torch-op builder
        @property
        def is_cuda(self):
            return True
    ref = fgnn_b200.SourcePlan(t64.as_subclass(_Dev), N, row_cap=row_cap)
    assert (native.row_cap, native.n_rows, native.n_edges) == (ref.row_cap, ref.n_rows, ref.n_edges)
    for name in ("src_ptr", "slot_edge", "edge_slot", "src_rows"):
        assert torch.equal(getattr(native, name), torch.as_tensor(getattr(ref, name))), name
    native32 = fgnn_b200.SourcePlan(t64.int(), N, row_cap=native.row_cap)
    assert torch.equal(native32.slot_edge, native.slot_edge)


def test_native_locality_order_matches_numpy():
    types = graphs.synthetic_map_graph(500, 1500, 250, 3, seed=9, local_band=32)
    a, b = graphs.locality_order(types), graphs.locality_order_numpy(types)
    for x, y in zip(a, b):
        assert np.array_equal(x.idx_v2f, y.idx_v2f) and np.array_equal(x.idx_f2v, y.idx_f2v)


def test_source_plan_zero_slots_need_the_device_builder():
    """SourcePlan(zero_slots=...) is a feature of the on-device builder; the host (native) builder and the batch-local plan
    say so instead of ignoring the hint."""
    idx = torch.randint(0, 50, (1, 80, 3))
    with pytest.raises(ValueError):
        fgnn_b200.SourcePlan(idx, 50, zero_slots=torch.zeros(80, 3, dtype=torch.bool))
    with pytest.raises(ValueError):
        fgnn_b200.SourcePlan(idx, 50, batch_local=True, zero_slots=torch.zeros(80, 3, dtype=torch.bool))


def test_source_plan_picks_row_cap_by_cost():
    import torch
    import fgnn_b200
    n = 4096
    six = torch.arange(n).repeat_interleave(6).view(1, -1, 2)        # every source feeds six slots
    two = torch.arange(n).repeat_interleave(2).view(1, -1, 2)
    assert fgnn_b200.SourcePlan(six, n).row_cap == 6 and fgnn_b200.SourcePlan(six, n).n_rows == n
    assert fgnn_b200.SourcePlan(six, n, row_cap=3).n_rows == 2 * n
    assert fgnn_b200.SourcePlan(two, n).row_cap == 3


def test_conv_bn_fold_and_conv1x1_match_pytorch():
    """Host-side rewrites around the core (eval mode): BatchNorm folded into the 1x1 convolution of
    mp_conv_residual's maps, and the 1x1 convolution as a row-major GEMM on node-major tensors, equal PyTorch's own
    Conv2d -> BatchNorm2d -> LeakyReLU sequence."""
    import torch
    import fgnn_b200
    from fgnn_b200.mp_nn import conv1x1
    torch.manual_seed(0)
    m = fgnn_b200.mp_conv_residual(48, 64, 4, extension=fgnn_b200.mp_conv_type.NO_EXTENSION, nout=32).eval()
    for seq in (m.conv1, m.conv2):
        seq[1].running_mean.uniform_(-0.3, 0.3)
        seq[1].running_var.uniform_(0.5, 1.5)
        seq[1].weight.data.uniform_(0.5, 1.5)
        seq[1].bias.data.uniform_(-0.2, 0.2)
    x = torch.randn(3, 48, 17, 1)
    with torch.no_grad():
        ref = m.conv1(x.clone())
        got = m._conv_bn_act(m.conv1, x.clone())
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
    # the cache follows the parameters
    with torch.no_grad():
        m.conv1[0].weight.mul_(1.5)
        assert torch.allclose(m._conv_bn_act(m.conv1, x.clone()), m.conv1(x.clone()), rtol=1e-5, atol=1e-6)
    # train mode keeps PyTorch's sequence (batch statistics)
    m.train()
    a = m._conv_bn_act(m.conv1, x.clone())
    assert a.requires_grad
    w, b = torch.randn(20, 48, 1, 1), torch.randn(20)
    xl = x.contiguous(memory_format=torch.channels_last)
    assert torch.allclose(conv1x1(xl, w, b), torch.nn.functional.conv2d(x, w, b), rtol=1e-5, atol=1e-5)


def test_static_table_use_counter():
    """A source plan is only built for an index table that is demonstrably static: the counter restarts whenever the
    tensor object is written in place (new _version) or replaced."""
    import torch
    from fgnn_b200 import mp_nn
    t = torch.zeros(1, 4, 2, dtype=torch.long)
    assert [mp_nn._table_use_count(t) for _ in range(3)] == [1, 2, 3]
    t.add_(1)                                   # in-place write: a different table as far as the cache goes
    assert mp_nn._table_use_count(t) == 1
    u = torch.zeros(1, 4, 2, dtype=torch.long)
    assert mp_nn._table_use_count(u) == 1 and mp_nn._table_use_count(t) == 2


def _bench_args(*argv):
    import bench
    old = sys.argv
    sys.argv = ["bench.py", *argv]
    try:
        return bench, bench.parse_args()
    finally:
        sys.argv = old


def test_bench_algorithmic_bytes_match_the_survey_figures():
    """The roofline numerator: bytes = B [4 M K + s T M K + s C N + s O M] per call, summed over types and both
    directions -- 387.0 MB / 1 500 000 messages per layer for cfg 2 at T = 16, 349.2 MB / 2 359 296 for the LDPC config."""
    bench, a = _bench_args()
    types = bench.build_graph(a)
    assert abs(bench.algorithmic_bytes_per_layer(a, types) / 1e6 - 387.0) < 0.05
    assert sum(t.real_messages for t in types) == 1_500_000
    bench, a = _bench_args("--config", "cfg3")
    types = bench.build_graph(a)
    assert abs(bench.algorithmic_bytes_per_layer(a, types) / 1e6 - 349.2) < 0.05
    assert sum(t.real_messages for t in types) * a.batch == 2_359_296
    bench, a = _bench_args("--edge-types", "4")
    assert abs(bench.algorithmic_bytes_per_layer(a, bench.build_graph(a)) / 1e6 - 312.6) < 0.05


def test_bench_reference_arm_contract_small():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the native one): one JSON line with the contract's
    keys, honouring --steps / --warmup, at a size that runs in seconds here."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--vars", "1000", "--pairwise", "3000",
                          "--high", "500", "--steps", "2", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    assert rec["impl"] == "reference" and rec["metric"] == "factor_messages_per_sec_per_fgnn_layer" and rec["unit"] == "messages/s"
    assert rec["steps"] == 2 and rec["warmup"] == 1 and rec["higher_is_better"] is True and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] in ("port", "port-torch") and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["d2h_bytes_per_step"] == 0 and rec["e2e"]["value"] == rec["value"]

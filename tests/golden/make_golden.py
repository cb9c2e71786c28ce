"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REAL REFERENCE
(/root/reference, lib/model/mpnn, torch CPU fp32, eval mode) on seeded inputs.

    python tests/golden/make_golden.py            # build container only; needs /root/reference

The reference ships no tests or golden vectors of its own, so these files are what
pins the oracle (tests/test_oracle_golden.py) and, through it and directly, the CUDA path
(tests/test_gpu_parity.py).  /root/reference does not exist on the GPU box; only the .npz
outputs travel.  Nothing here is imported by the product package.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

torch.set_num_threads(4)


def randomise(module, gen, filt_scale=0.3):
    """Weights that make decisions non-trivial (default init is bias-dominated, SURVEY 7 hard part 3)
    and BatchNorm running stats that are not the identity."""
    for name, p in module.named_parameters():
        if name.endswith("filters"):
            p.data.copy_((torch.rand(p.shape, generator=gen) * 2 - 1) * filt_scale / max(1.0, (p.shape[0] / 8) ** 0.5))
        elif p.dim() == 1:
            p.data.copy_((torch.rand(p.shape, generator=gen) - 0.5) * 0.4 + (1.0 if name.endswith("weight") else 0.0))
        else:
            p.data.copy_((torch.rand(p.shape, generator=gen) * 2 - 1) * (1.5 / max(1.0, p.shape[1] ** 0.5)))
    for name, b in module.named_buffers():
        if name.endswith("running_mean"):
            b.copy_((torch.rand(b.shape, generator=gen) - 0.5) * 0.2)
        elif name.endswith("running_var"):
            b.copy_(torch.rand(b.shape, generator=gen) + 0.5)
    return module.eval()


def sd_np(module, prefix=""):
    return {prefix + k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def unit_cases(mpnn):
    """mp_conv_v2 on its own: every extension x aggregator, plus the shape classes of SURVEY 8a."""
    gen = torch.Generator().manual_seed(1234)
    T_ = mpnn.mp_conv_type
    cases = []
    for ext in (0, 1, 2):
        for agg in ("max", "softmax", "mean", None):
            cases.append(dict(name=f"ext{ext}_{agg}", B=2, C=5, O=6, T=3, N=19, M=19, K=4, ext=ext, agg=agg))
    cases += [
        dict(name="noext_rect", B=2, C=7, O=4, T=2, N=11, M=23, K=3, ext=0, agg="max"),
        dict(name="noext_nobias_nobn_noact", B=1, C=4, O=3, T=2, N=9, M=5, K=2, ext=0, agg="max",
             bias=False, bn=False, act=None),
        dict(name="noext_T1_K1", B=3, C=6, O=8, T=1, N=10, M=10, K=1, ext=0, agg="max"),
        dict(name="diff_C2_first_layer", B=2, C=2, O=64, T=16, N=32, M=32, K=4, ext=1, agg="softmax"),
        dict(name="core64_T4", B=1, C=64, O=64, T=4, N=150, M=300, K=2, ext=0, agg="max"),
        dict(name="core64_T16", B=1, C=64, O=64, T=16, N=130, M=70, K=3, ext=0, agg="max"),
        dict(name="core_64_128", B=1, C=64, O=128, T=4, N=40, M=50, K=3, ext=0, agg="max"),
        dict(name="core_128_64", B=1, C=128, O=64, T=4, N=40, M=50, K=6, ext=0, agg="max"),
        dict(name="ldpc_f2v", B=3, C=64, O=64, T=4, N=48, M=96, K=3, ext=0, agg="max"),
        dict(name="ldpc_v2f", B=3, C=64, O=64, T=4, N=96, M=48, K=6, ext=0, agg="max"),
        dict(name="ldpc_global_v2f", B=2, C=64, O=64, T=1, N=96, M=1, K=96, ext=0, agg="max"),
        dict(name="ldpc_global_f2v", B=2, C=64, O=64, T=1, N=1, M=96, K=1, ext=0, agg="max"),
        dict(name="out2_softmax_last_layer", B=2, C=64, O=2, T=16, N=20, M=20, K=9, ext=2, agg="softmax",
             act=None),
    ]
    out = {}
    meta = []
    for c in cases:
        ext = T_(c["ext"])
        m = refload.quiet(mpnn.mp_conv_v2, c["C"], c["O"], c["T"], bias=c.get("bias", True),
                          bn=c.get("bn", True), extension=ext,
                          activation_fn=c.get("act", "relu"), aggregtor=c["agg"])
        randomise(m, gen)
        x = torch.randn(c["B"], c["C"], c["N"], 1, generator=gen)
        idx = torch.randint(0, c["N"], (c["B"], c["M"], c["K"]), generator=gen)
        et = torch.randn(c["B"], c["T"], c["M"], c["K"], generator=gen)
        # the reference's padding convention: some slots carry an all-zero edge type
        padmask = torch.rand(c["B"], 1, c["M"], c["K"], generator=gen) < 0.15
        et = et.masked_fill(padmask, 0.0)
        with torch.no_grad():
            y = m(x, idx, et)
        n = c["name"]
        out[f"{n}/x"], out[f"{n}/idx"], out[f"{n}/etype"], out[f"{n}/out"] = (
            x.numpy(), idx.numpy(), et.numpy(), y.numpy())
        for k, v in sd_np(m).items():
            out[f"{n}/sd/{k}"] = v
        meta.append(c)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "mp_conv_v2_cases.npz"), **out)
    print("mp_conv_v2_cases.npz:", len(cases), "cases")


def cfg1_simple_gnn(mpnn):
    """BASELINE configs[0]: train_syn_fixed_pw_hop.py `simple_gnn` on the 128-variable chain
    (generate_knn_table(128, 4) -> 512 directed slots = 256 pairwise factors), B = 8, MAP = argmax."""
    gen = torch.Generator().manual_seed(23456)
    fns = refload.script_functions("train_syn_fixed_pw_hop.py", ["generate_knn_table"])
    nn_idx, efeature = fns["generate_knn_table"](128, 4)
    T_ = mpnn.mp_conv_type
    model = refload.quiet(lambda: mpnn.mp_sequential(
        mpnn.mp_conv_v2(2, 64, 16, extension=T_.ORIG_WITH_NEIGHBOR),
        mpnn.mp_conv_residual(64, 64, 16), torch.nn.Conv2d(64, 2, 1)))
    emodel = torch.nn.Sequential(torch.nn.Conv2d(1, 64, 1), torch.nn.ReLU(inplace=True),
                                 torch.nn.Conv2d(64, 16, 1))
    randomise(model, gen)
    randomise(emodel, gen)
    B = 8
    x = torch.rand(B, 2, 128, 1, generator=gen)                 # node potentials ~U(0,1), random_pgm.py:22
    with torch.no_grad():
        etype = emodel(efeature)
        idx_b = nn_idx.repeat(B, 1, 1)
        et_b = etype.repeat(B, 1, 1, 1)
        h1 = model.module_list[0](x, idx_b, et_b)
        h2 = model.module_list[1](h1, idx_b, et_b)
        logits = model(x, idx_b, et_b)
    labels = logits.squeeze(-1).argmax(dim=1)
    margin = (logits[:, 0] - logits[:, 1]).abs().min().item()
    out = dict(x=x.numpy(), nn_idx=nn_idx.numpy(), efeature=efeature.numpy(), etype=etype.numpy(),
               h1=h1.numpy(), h2=h2.numpy(), logits=logits.numpy(), labels=labels.numpy(),
               min_margin=np.float32(margin))
    out.update(sd_np(model, "model/"))
    out.update(sd_np(emodel, "emodel/"))
    np.savez_compressed(os.path.join(HERE, "cfg1_simple_gnn.npz"), **out)
    print("cfg1_simple_gnn.npz: labels ones =", int(labels.sum()), "of", labels.numel(), "min |margin| =", margin)


def cfg1_factornn(mpnn):
    """cfg-1 secondary form: 128 variables, 256 pairwise factors f<128:(f,f+1),
    f>=128:(f-128,f-126) mod 128, FactorNN(2,[4],[64,64],[16],2), decision logit >= 0."""
    gen = torch.Generator().manual_seed(777)
    N, F = 128, 256
    f = np.arange(F)
    a = np.where(f < 128, f, f - 128)
    b = np.where(f < 128, (f + 1) % N, (f - 126) % N)
    idx_v2f = np.stack([a, b], 1).astype(np.int64)[None]            # [1,256,2]
    idx_f2v = np.zeros((1, N, 4), dtype=np.int64)
    fill = np.zeros(N, dtype=np.int64)
    for fi in range(F):
        for v in idx_v2f[0, fi]:
            idx_f2v[0, v, fill[v]] = fi
            fill[v] += 1
    assert (fill == 4).all()
    model = refload.quiet(mpnn.FactorNN, 2, [4], [64, 64], [16], 2)
    randomise(model, gen, filt_scale=0.2)
    B = 4
    node = torch.rand(B, 2, N, 1, generator=gen)
    hop = torch.rand(B, 4, F, 1, generator=gen)
    et_f2v = torch.randn(B, 16, N, 4, generator=gen)
    et_v2f = torch.randn(B, 16, F, 2, generator=gen)
    i_f2v = torch.from_numpy(idx_f2v).repeat(B, 1, 1)
    i_v2f = torch.from_numpy(idx_v2f).repeat(B, 1, 1)
    with torch.no_grad():
        logit = model(node, [hop], [i_f2v], [i_v2f], [et_f2v], [et_v2f])
    dec = (logit >= 0)
    out = dict(node=node.numpy(), hop=hop.numpy(), idx_f2v=idx_f2v, idx_v2f=idx_v2f,
               et_f2v=et_f2v.numpy(), et_v2f=et_v2f.numpy(), logit=logit.numpy(), decision=dec.numpy(),
               min_margin=np.float32(logit.abs().min().item()))
    out.update(sd_np(model, "model/"))
    np.savez_compressed(os.path.join(HERE, "cfg1_factornn.npz"), **out)
    print("cfg1_factornn.npz: decisions true =", int(dec.sum()), "of", dec.numel(),
          "min |margin| =", float(logit.abs().min()))


def ldpc(mpnn):
    """BASELINE configs[2] shapes: the MacKay 96.3.963 code tables from the reference's own
    ldpc_graph_structure_generator.get_mpnn_sp_structure (lib/data/ldpc_dataset.py:92-106) and an
    LDPC-shaped FactorNN (train_ldpc.py:23-30 with fewer layers), B = 4 noisy all-zero codewords."""
    gen = torch.Generator().manual_seed(963)
    gcls = refload.ldpc_structure()
    g = gcls()
    B = 4
    rng = np.random.default_rng(963)
    ys, hops, ef_f2v, ef_v2f = [], [], [], []
    for b in range(B):
        snr_db = float(b)
        y = (-(10 ** (snr_db / 20.0)) + rng.standard_normal(96)).astype(np.float32)   # all-zero codeword, BPSK 0 -> -g
        hop, idx_f2v, idx_v2f, e1, e2 = g.get_mpnn_sp_structure(y)
        ys.append(np.stack([y, np.full(96, snr_db, np.float32)], 0))
        hops.append(hop.astype(np.float32))
        ef_f2v.append(e1)
        ef_v2f.append(e2)
    node = torch.from_numpy(np.stack(ys)).reshape(B, 2, 96, 1)
    hop = torch.from_numpy(np.stack(hops)).permute(0, 2, 1).reshape(B, 6, 48, 1).contiguous()
    ef_f2v = torch.from_numpy(np.stack(ef_f2v)).permute(0, 3, 1, 2).contiguous()      # [B,7,96,3]
    ef_v2f = torch.from_numpy(np.stack(ef_v2f)).permute(0, 3, 1, 2).contiguous()      # [B,7,48,6]
    idx_f2v_t = torch.from_numpy(np.ascontiguousarray(idx_f2v)).long()[None].repeat(B, 1, 1)
    idx_v2f_t = torch.from_numpy(np.ascontiguousarray(idx_v2f)).long()[None].repeat(B, 1, 1)
    em_f2v = torch.nn.Sequential(torch.nn.Conv2d(7, 64, 1), torch.nn.ReLU(inplace=True), torch.nn.Conv2d(64, 4, 1))
    em_v2f = torch.nn.Sequential(torch.nn.Conv2d(7, 64, 1), torch.nn.ReLU(inplace=True), torch.nn.Conv2d(64, 4, 1))
    model = refload.quiet(mpnn.FactorNN, 2, [6, 96], [64, 64, 128, 64], [4, 1], 2, skip_link={2: 0},
                          ret_high=True)
    randomise(model, gen, filt_scale=0.2)
    randomise(em_f2v, gen)
    randomise(em_v2f, gen)
    # the script's "global" factor type (train_ldpc.py:40-55)
    h_idx_v2f = torch.arange(96).reshape(1, 1, 96).repeat(B, 1, 1)
    h_idx_f2v = torch.zeros(B, 96, 1, dtype=torch.long)
    h_et_v2f = torch.ones(B, 1, 1, 96)
    h_et_f2v = torch.ones(B, 1, 96, 1)
    nhop = node[:, 0, :, :].reshape(B, 96, 1, 1)
    with torch.no_grad():
        et_f2v = em_f2v(ef_f2v)
        et_v2f = em_v2f(ef_v2f)
        res, nhops = model(node, [hop, nhop], [idx_f2v_t, h_idx_f2v], [idx_v2f_t, h_idx_v2f],
                           [et_f2v, h_et_f2v], [et_v2f, h_et_v2f])
        res = res + node[:, :1]
    hard = (res >= 0)
    out = dict(node=node.numpy(), hop=hop.numpy(), idx_f2v=np.ascontiguousarray(idx_f2v).astype(np.int64),
               idx_v2f=np.ascontiguousarray(idx_v2f).astype(np.int64), ef_f2v=ef_f2v.numpy(), ef_v2f=ef_v2f.numpy(),
               et_f2v=et_f2v.numpy(), et_v2f=et_v2f.numpy(), res=res.numpy(), hard=hard.numpy(),
               nhop0=nhops[0].numpy(), nhop1=nhops[1].numpy(),
               min_margin=np.float32(res.abs().min().item()))
    out.update(sd_np(model, "model/"))
    np.savez_compressed(os.path.join(HERE, "ldpc_factornn.npz"), **out)
    print("ldpc_factornn.npz: hard ones =", int(hard.sum()), "of", hard.numel(), "min |margin| =",
          float(res.abs().min()), "idx_f2v", idx_f2v.shape, "idx_v2f", idx_v2f.shape)


def factor_mpnn_merged(mpnn):
    """The merged-table path of train_syn_hop_factor.py / train_syn_pw_factor.py: the reference's
    `factor_mpnn` (lib/model/mpnn/factor_mpnn.py:88-133) on a chain of n = 30 variables with the scripts' own
    `generate_pw_factor_table(n)` / `generate_high_factor_table(n, 4)` (train_syn_hop_factor.py:112-151) and edge
    models (:174-179); dims [64, 64, 2]: a residual layer (ORIG_WITH_DIFF cores, max) and the bare
    mp_conv_v2(64 -> 2) last layer with the default softmax aggregator; B = 4; MAP = argmax over the 2 channels."""
    gen = torch.Generator().manual_seed(4242)
    fns = refload.script_functions("train_syn_hop_factor.py", ["generate_pw_factor_table", "generate_high_factor_table"])
    n, hop = 30, 4
    idx_pw, ef_pw = fns["generate_pw_factor_table"](n)              # [1, 2n, 2], [1, 3, 2n, 2]
    idx_hi, ef_hi = fns["generate_high_factor_table"](n, hop)       # [1, 2n, hop], [1, 2, 2n, hop]
    model = refload.quiet(mpnn.factor_mpnn, 2, [4, hop], [64, 64, 2], [16, 16])
    em_pw = torch.nn.Sequential(torch.nn.Conv2d(3, 64, 1), torch.nn.ReLU(inplace=True), torch.nn.Conv2d(64, 16, 1))
    em_hi = torch.nn.Sequential(torch.nn.Conv2d(2, 64, 1), torch.nn.ReLU(inplace=True), torch.nn.Conv2d(64, 16, 1))
    randomise(model, gen, filt_scale=0.25)
    randomise(em_pw, gen)
    randomise(em_hi, gen)
    B = 4
    node = torch.rand(B, 2, n, 1, generator=gen)
    f_pw = torch.rand(B, 4, n, 1, generator=gen)
    f_hi = torch.rand(B, hop, n, 1, generator=gen)
    with torch.no_grad():
        et_pw = em_pw(ef_pw).repeat(B, 1, 1, 1)
        et_hi = em_hi(ef_hi).repeat(B, 1, 1, 1)
        i_pw, i_hi = idx_pw.repeat(B, 1, 1), idx_hi.repeat(B, 1, 1)
        out_v, out_f = model(node, [f_pw, f_hi], [[i_pw, et_pw], [i_hi, et_hi]])
        # centre the two logits (random weights leave the output bias-dominated: every label equal, SURVEY 7 hard
        # part 3) so that the MAP check is not vacuous
        d = torch.sort((out_v[:, 0] - out_v[:, 1]).reshape(-1)).values
        mid = slice(d.numel() // 3, 2 * d.numel() // 3)
        gap = int(torch.argmax(d[mid][1:] - d[mid][:-1])) + mid.start        # the widest gap near the median: decisions keep a margin
        model.mp_merge_modules[-1][-1].bias.data[0] -= 0.5 * (d[gap] + d[gap + 1])
        out_v, out_f = model(node, [f_pw, f_hi], [[i_pw, et_pw], [i_hi, et_hi]])
    labels = out_v.squeeze(-1).argmax(dim=1)
    margin = (out_v[:, 0] - out_v[:, 1]).abs().min().item()
    out = dict(node=node.numpy(), f_pw=f_pw.numpy(), f_hi=f_hi.numpy(), idx_pw=idx_pw.numpy(), idx_hi=idx_hi.numpy(),
               et_pw=et_pw.numpy(), et_hi=et_hi.numpy(), out_v=out_v.numpy(), out_f0=out_f[0].numpy(), out_f1=out_f[1].numpy(),
               labels=labels.numpy(), min_margin=np.float32(margin))
    out.update(sd_np(model, "model/"))
    np.savez_compressed(os.path.join(HERE, "factor_mpnn_merged.npz"), **out)
    print("factor_mpnn_merged.npz: labels ones =", int(labels.sum()), "of", labels.numel(), "min |margin| =", margin)


def main():
    if not refload.available():
        raise SystemExit("reference tree not found; golden vectors can only be regenerated in the build container")
    mpnn = refload.load()
    torch.manual_seed(0)
    np.random.seed(23456)
    if len(sys.argv) > 1 and sys.argv[1] == "factor_mpnn":
        factor_mpnn_merged(mpnn)
        return
    unit_cases(mpnn)
    cfg1_simple_gnn(mpnn)
    cfg1_factornn(mpnn)
    ldpc(mpnn)
    factor_mpnn_merged(mpnn)


if __name__ == "__main__":
    main()

import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import fgnn_b200
from fgnn_b200 import _lib, graphs
dev = torch.device("cuda:0")
rng = np.random.default_rng(1)
types = graphs.synthetic_map_graph(100000, 300000, 50000, 3, seed=0)
ty = types[1]
C = O = 64; T = 16
nm = lambda t: t.permute(0, 2, 1).unsqueeze(-1)
x = torch.from_numpy(np.abs(rng.standard_normal((1, ty.n_factors, C))).astype(np.float32)).to(dev)
idx = ty.idx_f2v; M = ty.n_vars; K = idx.shape[1]
d_idx = torch.from_numpy(idx[None]).to(dev)
et = torch.from_numpy(rng.standard_normal((1, T, M, K)).astype(np.float32)).to(dev)
W = torch.from_numpy(rng.uniform(-0.01, 0.01, (C, O * T)).astype(np.float32)).to(dev)
bias = torch.from_numpy(rng.uniform(0, 0.05, O).astype(np.float32)).to(dev)
ws = torch.zeros(C * O * T * 4 + 4096, dtype=torch.uint8, device=dev)
def run(plan):
    out = torch.empty((1, M, O), dtype=torch.float32, device=dev)
    fgnn_b200.mp_forward(nm(x), d_idx, et, W, bias, None, None, extension=0, aggregator=_lib.AGG_MAX, activation=_lib.ACT_NONE,
                         kernel=_lib.KERNEL_TCGEN05, out=nm(out), workspace=ws, filters_version=7, plan=plan, validate=False)
    torch.cuda.synchronize()
    return out
ref = run(None)
for cap in (3, 6):
    plan = fgnn_b200.SourcePlan(d_idx, ty.n_factors, row_cap=cap)
    for rep in range(3):
        got = run(plan)
        bad = (got != ref)
        rows = bad.any(-1)[0].nonzero()[:, 0]
        print("cap", cap, "rep", rep, "bad elems", int(bad.sum()), "bad rows", rows.numel(), rows[:10].tolist())
        if rows.numel():
            r = int(rows[0])
            ch = bad[0, r].nonzero()[:, 0]
            print("  row", r, "idx", idx[r].tolist(), "bad channels", ch[:16].tolist(), "n", ch.numel())
            print("  got", got[0, r, ch[:4]].tolist(), "ref", ref[0, r, ch[:4]].tolist())
            se = plan.slot_edge.view(M, K)[r].tolist()
            print("  edges", se, "vrows of edges", [int(torch.searchsorted(plan.src_ptr, torch.tensor([e], device=dev, dtype=torch.int32), right=True)[0]) - 1 for e in se])
            # distribution of bad rows by hub involvement
            hub = torch.from_numpy((idx == 0).any(1)).to(dev)
            print("  bad rows with a hub slot:", int(hub[rows].sum()), "of", rows.numel())

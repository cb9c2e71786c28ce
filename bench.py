#!/usr/bin/env python
"""Headline benchmark: factor-messages/sec per FGNN layer (BASELINE.json `metric`).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference            # the CPU restatement of the reference path

Workload (`config.workload`): BASELINE.json configs[1] -- synthetic MAP-inference factor graph,
100 000 variables, 300 000 pairwise + 50 000 order-3 factors, 5 stacked FGNN layers, fp32,
C = O = 64, T = 16 edge types (the synthetic scripts' value, train_syn_hop_factor.py:170-172).
One STEP = one pass of the hot path over that graph: for each of the 5 layers, for each factor
type, the Variable->Factor call and the Factor->Variable call of FactorNN's layer body
(factor_mpnn_sp.py:142-151) = 20 `fgnn_mp_forward` launches.  A MESSAGE is one real
(destination, slot) evaluation in one direction; per layer = sum_types 2*F*K = 1 500 000.

Prints ONE JSON line (rank 0).  `value` = messages/s with inputs resident in HBM; `e2e` = the
same through the public module API with HOST buffers (pinned), H2D of the step's inputs and D2H
of the resulting variable features inside the timed region.  `roofline` is the HBM roofline of
the message-passing kernel: algorithmic bytes (SURVEY 8d formula) / CUDA-event time, against
MEASURED_PEAKS.json (per GPU when N > 1).  `cpu_baseline` times the reference's own op chain
(oracle/fgnn_oracle_torch.py: mm -> index repeat -> gather -> bmm -> max, PyTorch CPU, all threads) on a
bounded sample of the workload, with the C/OpenMP restatement beside it; `gpu_aten_baseline` runs that same op
chain on the B200.  `checksum` = exact integer sum of the final variable
features' bit patterns: identical for every N (the sharded layer is bit-identical to the single-GPU one).
`--impl reference` runs the C restatement on the host cores at full scale, honouring --steps / --warmup.

`--config cfg3` = BASELINE configs[2]: the LDPC 96.3.963 decoding graph (check factors), batch 4096 codewords,
T = 4, 10 message-passing iterations; multi-GPU shards the batch (independent codewords: replicas, no collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "factor_messages_per_sec_per_fgnn_layer"
UNIT = "messages/s"
FALLBACK_HBM_GBS = 6650.0        # B200_PROFILING.md fallback, used only if MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--vars", type=int, default=100_000)
    ap.add_argument("--pairwise", type=int, default=300_000)
    ap.add_argument("--high", type=int, default=50_000)
    ap.add_argument("--high-order", type=int, default=3)
    ap.add_argument("--layers", type=int, default=5)
    ap.add_argument("--edge-types", type=int, default=16)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"],
                    help="f32 = BASELINE configs[1]; bf16 = features / edge types / output in bf16, fp32 accumulate "
                         "(configs[3]: --dtype bf16 --vars 1000000 --pairwise 3000000 --high 500000 --high-order 4 --layers 8)")
    ap.add_argument("--local-band", type=int, default=0, help="0 = uniform-random incidence (primary)")
    ap.add_argument("--src-calls", default="auto",
                    help="which core calls run source-stationary (one row-product per source row, csrc/mp_src.cu): "
                         "'auto' (fan-out rule of mp_conv_v2), 'none', or a comma list of v2f<j>/f2v<j> (j = factor type)")
    ap.add_argument("--cpu-sample-scale", type=int, default=2, help="cpu_baseline runs on 1/scale of the graph")
    ap.add_argument("--exchange", default="auto", choices=["auto", "halo", "peer", "nccl"],
                    help="multi-GPU: 'auto' = halo for graphs with index locality (--local-band) and for bf16, peer for "
                         "uniform-random fp32 graphs (every variable is a boundary variable there: measured, DESIGN.md 7); "
                         "'halo' = owner-computes sharding, one NVLink peer-memory kernel pulls the feature halos "
                         "per layer (parallel.HaloLayerPlan); 'peer' = factor sharding + fused max-reduce/epilogue/broadcast "
                         "kernel over peer memory; 'nccl' = factor sharding + NCCL all_reduce(MAX) + epilogue kernel")
    ap.add_argument("--order", default="locality", choices=["locality", "none"],
                    help="factor numbering before sharding: 'locality' sorts every type's factors by their smallest variable "
                         "(graphs.locality_order) so contiguous shards cut few edges when the graph has index locality")
    ap.add_argument("--concurrent-layer", action="store_true",
                    help="single GPU: issue the V->F calls of a layer on their own streams beside the F->V chain "
                         "(measured: +2 %% at T=4, nothing at T=16 -- programmatic launch already hides the hand-over)")
    ap.add_argument("--no-zero-slots", action="store_true",
                    help="source-stationary plans evaluate the padded slots (valid index + zero edge type) like any other slot "
                         "instead of treating them as constant-zero messages")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch every step from Python instead of replaying a CUDA graph of the step")
    ap.add_argument("--exchange-ctas", type=int, default=64, help="grid of the peer exchange kernel (512-thread CTAs, two per SM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--batch", type=int, default=1, help="graphs per step (cfg3: codewords)")
    ap.add_argument("--sustain-seconds", type=float, default=2.0,
                    help="also report the rate over a back-to-back run of at least this many seconds (0 = skip)")
    ap.add_argument("--no-aten-baseline", action="store_true")
    ap.add_argument("--config", default=None, choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json presets: cfg2 = the default (configs[1]); cfg4 = configs[3] sizes on this many GPUs "
                         "(1 M variables, 3 M pairwise + 500 K order-4, 8 layers, bf16 I/O: single GPU); cfg5 = configs[4] "
                         "(65 536 points in 3-D numbered along a Z-order curve, kNN-16 as 524 288 pairwise factors + 8 192 patch factors of "
                         "order 16, 12 layers: graphs.point_cloud_graph; --cfg5-random keeps the sizes with uniform-random incidence)")
    ap.add_argument("--cfg5-random", action="store_true", help="cfg5 sizes on uniform-random incidence instead of the point cloud")
    args = ap.parse_args()
    if args.config == "cfg4":
        args.vars, args.pairwise, args.high, args.high_order, args.layers, args.dtype = 1_000_000, 3_000_000, 500_000, 4, 8, "bf16"
        args.cpu_sample_scale = max(args.cpu_sample_scale, 10)
    elif args.config == "cfg5":
        args.vars, args.pairwise, args.high, args.high_order, args.layers = 65_536, 524_288, 8_192, 16, 12
        args.cpu_sample_scale = max(args.cpu_sample_scale, 4)
    elif args.config == "cfg3":
        args.layers, args.edge_types = 10, 4
        if args.batch == 1:
            args.batch = 4096
    args.point_cloud = args.config == "cfg5" and not args.cfg5_random and not args.local_band
    if args.exchange == "auto":
        args.exchange = "halo" if (args.local_band or args.point_cloud or args.dtype == "bf16") else "peer"
    return args


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------

def build_graph(args, scale=1):
    from fgnn_b200 import graphs
    if args.config == "cfg3":
        # the parity-check graph of ldpc_codes/96.3.963 as the reference's get_mpnn_sp_structure emits it
        # (lib/data/ldpc_dataset.py:92-106); the tables travel in the committed golden fixture
        z = np.load(os.path.join(ROOT, "tests", "golden", "ldpc_factornn.npz"))
        return [graphs.FactorType(z["idx_v2f"], z["idx_f2v"], np.zeros(z["idx_f2v"].shape, bool), "ldpc-checks")]
    if args.point_cloud:
        # kNN + patch factors over a seeded 3-D point set (SURVEY 8d cfg 5); the sizes follow from the point count
        types = graphs.point_cloud_graph(args.vars // scale, 16, args.high // scale, args.high_order, seed=args.seed)
    else:
        types = graphs.synthetic_map_graph(args.vars // scale, args.pairwise // scale, args.high // scale,
                                           args.high_order, seed=args.seed, local_band=args.local_band)
    if args.order == "locality" and (args.local_band or args.point_cloud):
        types = graphs.locality_order(types)
    return types


def layer_weights(args, n_types, rng):
    """Per layer / type / direction: filters ~U(-0.01,0.01), bias ~U(0,0.05) (mp_nn.py:49,53), BN
    running stats randomised, folded to scale/shift."""
    C = O = args.dim
    T = args.edge_types
    ws = []
    for _ in range(args.layers):
        per_type = []
        for _ in range(n_types):
            d = {}
            for direction in ("v2f", "f2v"):
                rv = rng.uniform(0.5, 1.5, O).astype(np.float32)
                rm = rng.uniform(-0.05, 0.05, O).astype(np.float32)
                scale = (1.0 / np.sqrt(rv + 1e-5)).astype(np.float32)
                d[direction] = dict(filters=rng.uniform(-0.01, 0.01, (C, O * T)).astype(np.float32),
                                    bias=rng.uniform(0, 0.05, O).astype(np.float32),
                                    scale=scale, shift=(-rm * scale).astype(np.float32), rm=rm, rv=rv)
            per_type.append(d)
        ws.append(per_type)
    return ws


def host_inputs(args, types, rng, batch=None):
    """Features and edge types per graph of the batch (per-sample, like LDPC's); index tables once ([1,M,K]:
    the batch sees them through a stride-0 expand, as identical tables collated by the reference's DataLoader)."""
    C, T = args.dim, args.edge_types
    B = args.batch if batch is None else batch
    x_v = rng.random((B, types[0].n_vars, C), dtype=np.float32)              # node-major [B,N,C]
    x_f = [np.abs(rng.standard_normal((B, t.n_factors, C), dtype=np.float32)) for t in types]
    et_v2f = [rng.standard_normal((B, T, t.n_factors, t.order), dtype=np.float32) for t in types]
    et_f2v = []
    for t in types:
        e = rng.standard_normal((B, T, t.n_vars, t.kv), dtype=np.float32)
        e[np.broadcast_to(t.pad_f2v[None, None], e.shape)] = 0.0               # reference padding: etype = 0
        et_f2v.append(e)
    return dict(x_v=x_v, x_f=x_f, et_v2f=et_v2f, et_f2v=et_f2v,
                idx_v2f=[t.idx_v2f[None] for t in types], idx_f2v=[t.idx_f2v[None] for t in types])


def algorithmic_bytes_per_layer(args, types):
    """per core call 4*M*K (idx as int32) + s*T*M*K (etype) + s*C*N_src + s*O*M."""
    s, C, O, T = (2 if args.dtype == "bf16" else 4), args.dim, args.dim, args.edge_types
    total = 0
    for t in types:
        total += 4 * t.n_factors * t.order + s * T * t.n_factors * t.order + s * C * t.n_vars + s * O * t.n_factors
        total += 4 * t.n_vars * t.kv + s * T * t.n_vars * t.kv + s * C * t.n_factors + s * O * t.n_vars
    return total * args.batch


def workload_name(args):
    if args.config == "cfg3":
        return (f"LDPC 96.3.963 decoding graph (96 variables, 48 check factors of order 6), batch {args.batch} codewords, "
                f"{args.layers} message-passing iterations, C=O={args.dim}, T={args.edge_types}, fp32")
    if args.point_cloud:
        return (f"point cloud: {args.vars} points in 3-D (Z-order numbering), kNN-16 as {args.pairwise} pairwise factors + {args.high} "
                f"patch factors of order {args.high_order}, {args.layers} FGNN layers, C=O={args.dim}, T={args.edge_types}, fp32")
    return (f"synthetic MAP inference: {args.vars} vars, {args.pairwise} pairwise + {args.high} order-{args.high_order} "
            f"factors, {args.layers} FGNN layers, C=O={args.dim}, T={args.edge_types}, {'bf16 I/O' if args.dtype == 'bf16' else 'fp32'}, "
            f"{'uniform-random' if not args.local_band else f'band-{args.local_band}'} incidence")


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement of the reference algorithm (bounded sample)
# ---------------------------------------------------------------------------------------------

def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which is not the machine)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_layer_pass(args, types, inp, weights, threads=0):
    """One FGNN layer (all types, both directions) on the C restatement of the reference; returns seconds."""
    from oracle import fgnn_oracle as orc
    B = inp["x_v"].shape[0]
    rep = lambda a: np.broadcast_to(a, (B,) + a.shape[1:])
    t0 = time.perf_counter()
    x_v = np.ascontiguousarray(inp["x_v"].transpose(0, 2, 1))[..., None]
    for j, ty in enumerate(types):
        w = weights[j]
        x_f = np.ascontiguousarray(inp["x_f"][j].transpose(0, 2, 1))[..., None]
        for direction, x, idx, et in (("v2f", x_v, inp["idx_v2f"][j], inp["et_v2f"][j]),
                                      ("f2v", x_f, inp["idx_f2v"][j], inp["et_f2v"][j])):
            p = w[direction]
            bn = dict(weight=np.ones_like(p["rv"]), bias=np.zeros_like(p["rv"]), running_mean=p["rm"], running_var=p["rv"])
            orc.mp_conv_forward_c(x, np.ascontiguousarray(rep(idx)), et, p["filters"], p["bias"], bn, extension=0,
                                  aggregator="max", threads=threads)
    return time.perf_counter() - t0


def torch_layer_pass(args, types, inp, weights, device="cpu", chunk_bytes=2 << 30, reps=1):
    """One FGNN layer through the reference's own op chain (oracle/fgnn_oracle_torch.py: permute -> mm -> int64 index
    repeat -> gather -> bmm -> max -> bias -> eval BatchNorm -> ReLU, mp_nn.py:115-175) in PyTorch on `device`.
    Destinations are evaluated in slices so that the O*T-wide int64 index + gathered rows of a slice stay below
    `chunk_bytes` (bit-identical in eval mode).  Returns seconds per pass (mean of `reps` after 1 warm-up)."""
    import torch
    from oracle import fgnn_oracle_torch as orct
    dev = torch.device(device)
    B = inp["x_v"].shape[0]
    OT = args.dim * args.edge_types
    calls = []
    x_v = torch.from_numpy(np.ascontiguousarray(inp["x_v"].transpose(0, 2, 1))[..., None]).to(dev)
    for j, ty in enumerate(types):
        x_f = torch.from_numpy(np.ascontiguousarray(inp["x_f"][j].transpose(0, 2, 1))[..., None]).to(dev)
        for direction, x, idx, et in (("v2f", x_v, inp["idx_v2f"][j], inp["et_v2f"][j]),
                                      ("f2v", x_f, inp["idx_f2v"][j], inp["et_f2v"][j])):
            p = weights[j][direction]
            bn = {k: torch.from_numpy(v).to(dev) for k, v in dict(weight=np.ones_like(p["rv"]), bias=np.zeros_like(p["rv"]),
                                                                   running_mean=p["rm"], running_var=p["rv"]).items()}
            K = idx.shape[2]
            rows = max(1, int(chunk_bytes // (12 * B * K * OT)))       # 8-byte index + 4-byte value per gathered element
            calls.append((x, torch.from_numpy(np.array(np.broadcast_to(idx, (B,) + idx.shape[1:]))).to(dev),
                          torch.from_numpy(et).to(dev), torch.from_numpy(p["filters"]).to(dev),
                          torch.from_numpy(p["bias"]).to(dev), bn, rows))

    def one_pass():
        with torch.no_grad():
            for x, idx, et, W, bias, bn, rows in calls:
                orct.mp_conv_forward_torch(x, idx, et, W, bias, bn, extension=0, aggregator="max", chunk_rows=rows)
        if dev.type == "cuda":
            torch.cuda.synchronize()
    one_pass()
    t0 = time.perf_counter()
    for _ in range(reps):
        one_pass()
    return (time.perf_counter() - t0) / reps


def cpu_baseline(args, budget_s=20.0):
    """The reference's CPU path beside the GPU number (rank 0, N = 1): its own ATen op chain on PyTorch CPU with all
    host threads (`kind: port-torch`; /root/reference itself does not exist on the GPU box), on a BOUNDED sample --
    one FGNN layer on a 1/scale instance of the workload, scale chosen so that the leg takes about `budget_s`.  The
    C/OpenMP restatement (the `--impl reference` arm) is timed on the same sample for comparison."""
    import torch
    from oracle import fgnn_oracle as orc
    orc.build_c()
    cores = host_threads()
    torch.set_num_threads(cores)
    rng = np.random.default_rng(args.seed + 1)
    # survey probe: ~0.2 M messages/s for the ATen chain at T=16 on 8 threads -> pick the sample for ~budget/4 per pass
    full_msgs = None
    scale, batch = 1, args.batch
    while True:
        types = build_graph(args, scale)
        msgs = sum(t.real_messages for t in types) * batch
        if full_msgs is None:
            full_msgs = msgs
        if msgs <= 0.2e6 * cores / 8 * budget_s / 4 or (scale >= 64 and batch <= 8):
            break
        if args.config == "cfg3":
            batch = max(8, batch // 2)
            if batch == 8:
                break
        else:
            scale *= 2
    inp = host_inputs(args, types, rng, batch)
    weights = layer_weights(args, len(types), rng)[0]
    t_torch = torch_layer_pass(args, types, inp, weights, "cpu", reps=2)
    cpu_layer_pass(args, types, inp, weights, cores)          # warm-up
    t_c = min(cpu_layer_pass(args, types, inp, weights, cores) for _ in range(3))
    what = (f"1 FGNN layer on {'a batch of %d of the %d codewords' % (batch, args.batch) if args.config == 'cfg3' else 'a 1/%d-scale instance of the workload' % scale} "
            f"({types[0].n_vars} vars, {msgs} messages)")
    return dict(value=msgs / t_torch, unit=UNIT, cores=cores, kind="port-torch",
                sample=(f"{what}: oracle/fgnn_oracle_torch.py = the reference's op chain (mp_nn.py:115-175: mm -> int64 index "
                        f"repeat -> gather -> bmm -> max -> BN -> ReLU) on PyTorch CPU, {cores} threads, mean of 2 passes after 1 warm-up"),
                c_port={"value": msgs / t_c, "unit": UNIT, "kind": "port",
                        "what": "oracle/fgnn_oracle.c (C restatement, OpenMP, same threads) on the same sample, best of 3"})


def gpu_aten_baseline(args, types, inp, weights):
    """The reference's unmodified op chain on the SAME B200: PyTorch/ATen CUDA kernels, fp32,
    one FGNN layer of the full workload (destinations sliced to bound the O*T-wide intermediates)."""
    import torch
    try:
        t = torch_layer_pass(args, types, inp, weights, "cuda", chunk_bytes=8 << 30, reps=3)
    except Exception as e:  # noqa: BLE001  (out of memory on a shared box: report, do not fail the bench)
        return {"unavailable": f"{type(e).__name__}: {str(e)[:120]}"}
    finally:
        torch.cuda.empty_cache()
    msgs = sum(t_.real_messages for t_ in types) * args.batch
    return {"value": msgs / t, "unit": UNIT, "ms_per_layer": t * 1e3,
            "what": "oracle/fgnn_oracle_torch.py (the reference's op chain, mp_nn.py:115-175) on this GPU through ATen's CUDA "
                    "kernels, fp32, full workload, mean of 3 layer passes after 1 warm-up, wall clock around a device sync"}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores -- the C/OpenMP restatement (oracle/fgnn_oracle.c;
    the reference itself is Python-over-ATen and /root/reference does not exist on the GPU box) with all host
    threads, on the arm's own workload at FULL scale, `--warmup` + `--steps` steps as asked.  Only if that would
    take more than ~4 minutes is every step cut to a 1/scale instance (said in `config.sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fgnn_oracle as orc
    orc.build_c()
    cores = host_threads()
    rng = np.random.default_rng(args.seed + 1)
    scale, batch = 1, args.batch
    while True:
        types = build_graph(args, scale)
        inp = host_inputs(args, types, rng, batch)
        weights = layer_weights(args, len(types), rng)
        cpu_layer_pass(args, types, inp, weights[0], cores)                       # first touch
        t1 = cpu_layer_pass(args, types, inp, weights[0], cores)
        if t1 * args.layers * (args.steps + args.warmup) <= 240.0 or scale >= 64 or batch <= 8:
            break
        if args.config == "cfg3":
            batch = max(8, batch // 2)
        else:
            scale *= 2
    msgs = sum(t.real_messages for t in types) * batch
    for _ in range(args.warmup):
        for l in range(args.layers):
            cpu_layer_pass(args, types, inp, weights[l], cores)
    t0 = time.perf_counter()
    for s in range(args.steps):
        for l in range(args.layers):
            cpu_layer_pass(args, types, inp, weights[l], cores)
    dt = time.perf_counter() - t0
    value = msgs * args.layers * args.steps / dt
    full = scale == 1 and batch == args.batch
    # beside it, for the record: the reference's OWN op chain (mm -> int64 index repeat -> gather -> bmm -> max, mp_nn.py:115-175)
    # through PyTorch on the same host threads, one layer of the same instance -- the port above is the FASTER of the two,
    # so the arm's number is the conservative denominator
    op_chain = None
    try:
        import torch
        torch.set_num_threads(cores)
        t_torch = torch_layer_pass(args, types, inp, weights[0], device="cpu", reps=1)
        op_chain = {"value": msgs / t_torch, "unit": UNIT, "kind": "port-torch", "seconds_per_layer": t_torch,
                    "what": "oracle/fgnn_oracle_torch.py (the reference's ATen op chain) on PyTorch CPU, same instance and threads, "
                            "1 layer pass after 1 warm-up"}
    except Exception as e:  # noqa: BLE001
        op_chain = {"unavailable": f"{type(e).__name__}: {e}"}
    sample = (f"each step = {args.layers} FGNN layers on {'the full workload' if full else ('a 1/%d-scale instance' % scale if args.config != 'cfg3' else 'a batch of %d' % batch)} "
              f"({types[0].n_vars} vars, {msgs} messages/layer), oracle/fgnn_oracle.c (C restatement of mp_nn.py:115-175, OpenMP, {cores} threads)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample, "full_scale": full},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "reference_op_chain_torch": op_chain},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def clocks_record(clocks, sustained, timed_ms):
    """nvidia-smi samples every 100 ms; a timed region of a few tens of milliseconds catches 0-2 of them.  When that
    happens the samples of the sustained run (the same step replayed back to back for seconds, its own timed region) are
    reported beside them, so the record always carries clocks and throttle reasons measured under this load."""
    if clocks is None:
        return None
    out = dict(clocks)
    if (out.get("samples") or 0) < 3 and sustained and sustained.get("clocks"):
        sc = sustained["clocks"]
        out["note"] = (f"timed region {timed_ms:.0f} ms against a 100 ms sampling period; under the same steps replayed for "
                       f"{sustained['seconds']:.1f} s: sm {sc.get('sm_mhz')} MHz median over {sc.get('samples')} samples, reasons {sc.get('reasons')}")
        if out.get("sm_mhz") is None:
            out["sm_mhz"], out["sm_max_mhz"] = sc.get("sm_mhz"), sc.get("sm_max_mhz")
            out["reasons"] = sc.get("reasons", [])
    return out


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def tensor_peak():
    """Sustained dense bf16 TFLOP/s (the kernels run inside a long step)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 0.0))), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1360.0, "fallback"


def executed_mma_flops_per_layer(args, types, plans):
    """bf16 MMA flops the step really issues per layer: 3 split terms x 2*C*O*T per row-product; one
    row-product per SLOT in a destination-stationary call, per SOURCE ROW in a source-stationary call."""
    C = O = args.dim
    T = args.edge_types
    rows = 0
    for j, t in enumerate(types):
        rows += t.n_vars if ("v2f%d" % j) in plans else t.n_factors * t.order
        rows += t.n_factors if ("f2v%d" % j) in plans else t.n_vars * t.kv
    return (1 if args.dtype == "bf16" else 3) * 2 * C * O * T * rows


def run_native(args):
    import torch
    import torch.distributed as dist
    import fgnn_b200
    from fgnn_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_MAX_CTAS", "16")       # the all-reduce shares the GPU with the V->F kernels
        dist.init_process_group("nccl", device_id=dev)
    kernel = {"auto": _lib.KERNEL_AUTO, "simt": _lib.KERNEL_SIMT, "tcgen05": _lib.KERNEL_TCGEN05}[args.kernel]

    rng = np.random.default_rng(args.seed + 1)
    types = build_graph(args)
    inp = host_inputs(args, types, rng)
    weights = layer_weights(args, len(types), rng)
    msgs_layer = sum(t.real_messages for t in types) * args.batch
    bytes_layer = algorithmic_bytes_per_layer(args, types)
    J, L, C = len(types), args.layers, args.dim
    # independent graphs (cfg3: codewords) shard by BATCH: replicas, no collective
    batch_sharded = world > 1 and args.batch >= world
    if batch_sharded:
        from fgnn_b200.parallel import shard_range
        b0, b1 = shard_range(args.batch, rank, world)
        for k in ("x_v",):
            inp[k] = inp[k][b0:b1]
        for k in ("x_f", "et_v2f", "et_f2v"):
            inp[k] = [a[b0:b1] for a in inp[k]]
    B_loc = inp["x_v"].shape[0]

    def to_dev(a, pin=False):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.pin_memory() if pin else t.to(dev)

    def nm(t):                                   # node-major [1,N,C] memory viewed as the API's [1,C,N,1]
        return t.permute(0, 2, 1).unsqueeze(-1)

    W = [[{d: {k: to_dev(v) for k, v in weights[l][j][d].items() if k in ("filters", "bias", "scale", "shift")}
           for d in ("v2f", "f2v")} for j in range(J)] for l in range(L)]
    ws = [[{d: torch.zeros(C * C * args.edge_types * 4 + 4096, dtype=torch.uint8, device=dev) for d in ("v2f", "f2v")}
           for j in range(J)] for l in range(L)]
    for l in range(L):
        for j in range(J):
            ws[l][j]["ver_v2f"], ws[l][j]["ver_f2v"] = 1 + l * 16 + j * 2, 2 + l * 16 + j * 2
    bf16 = args.dtype == "bf16"
    fdev = (lambda a: to_dev(a).to(torch.bfloat16)) if bf16 else to_dev
    d_in = dict(x_v=fdev(inp["x_v"]), x_f=[fdev(a) for a in inp["x_f"]],
                et_v2f=[fdev(a) for a in inp["et_v2f"]], et_f2v=[fdev(a) for a in inp["et_f2v"]],
                idx_v2f=[to_dev(a).expand(B_loc, -1, -1) for a in inp["idx_v2f"]],
                idx_f2v=[to_dev(a).expand(B_loc, -1, -1) for a in inp["idx_f2v"]])
    plan = halo = None
    if world > 1 and not batch_sharded and args.exchange == "halo":
        # owner-computes sharding: this rank keeps its variables, its factors, the renumbered tables, their edge types
        from fgnn_b200 import parallel
        halo = parallel.HaloLayerPlan(types, rank, world, dev, torch.bfloat16 if args.dtype == "bf16" else torch.float32, C,
                                      ctas=args.exchange_ctas)
        halo.connect_distributed()
        d_in["et_v2f"], d_in["et_f2v"] = halo.local_etypes(d_in["et_v2f"], d_in["et_f2v"])
        d_in["idx_v2f"], d_in["idx_f2v"] = halo.idx_v2f, halo.idx_f2v
        torch.cuda.empty_cache()
    elif world > 1 and not batch_sharded:
        # factor-sharded: this rank keeps its factor ranges, the compacted F->V tables and their edge types
        if args.dtype == "bf16":
            raise RuntimeError("bench.py: --dtype bf16 on several GPUs needs --exchange halo (the raw-aggregate exchanges are fp32)")
        from fgnn_b200 import parallel
        plan = parallel.ShardedLayerPlan(types, rank, world, dev, exchange=args.exchange, exchange_ctas=args.exchange_ctas)
        d_in["x_f"] = plan.local_factor_features(d_in["x_f"])
        d_in["et_v2f"], d_in["et_f2v"] = plan.local_etypes(d_in["et_v2f"], d_in["et_f2v"])
        d_in["idx_v2f"] = d_in["idx_f2v"] = None
    # ping-pong feature buffers, node-major
    buf_v = [torch.empty_like(d_in["x_v"]) for _ in range(2)]
    buf_f = [[torch.empty_like(x) for x in d_in["x_f"]] for _ in range(2)]

    # source-stationary plans of the index tables (static across layers and steps: built once, outside the
    # timed region, like the reference's own table construction)
    plans = {}
    if plan is None and args.src_calls != "none" and kernel != _lib.KERNEL_SIMT and C == 64 and args.edge_types in (4, 8, 16) and not bf16:
        rule = fgnn_b200.mp_conv_v2.AUTO_FAN_OUT[args.edge_types]
        for j, ty in enumerate(types):
            n_v, n_f = (ty.n_vars, ty.n_factors) if halo is None else (halo.rows_v, halo.rows_f[j])
            for name, idx, n_src in (("v2f%d" % j, d_in["idx_v2f"][j], n_v), ("f2v%d" % j, d_in["idx_f2v"][j], n_f)):
                if idx.numel() == 0:
                    continue
                fan = idx.numel() / (n_src * B_loc)
                if args.src_calls == "auto" and args.edge_types == 4 and n_src <= 128 and B_loc >= 64:
                    sp = fgnn_b200.SourcePlan(idx, n_src, batch_local=True)   # batched small graphs: aggregation fused into the first pass
                    if sp.fusable(C, args.edge_types):
                        plans[name] = sp
                    continue
                if (args.src_calls == "auto" and fan >= rule) or name in args.src_calls.split(","):
                    # the variable-side tables pad with (index 0, all-zero edge type): the generator knows which slots
                    # (FactorType.pad_f2v), the plan leaves them out and the second pass aggregates a literal 0 for them
                    zs = None
                    if not args.no_zero_slots and name.startswith("f2v") and ty.pad_f2v.any():
                        # (sharded: the rows of this rank's table are its own variables, in order)
                        zs = torch.from_numpy(ty.pad_f2v if halo is None else np.ascontiguousarray(ty.pad_f2v[halo.v0:halo.v1])).to(dev)
                    sp = fgnn_b200.SourcePlan(idx, n_src, zero_slots=zs)
                    if args.src_calls != "auto" or sp.n_rows * 1.25 <= idx.numel():
                        plans[name] = sp        # (hub sources -- the reference's pad target -- are split into virtual rows)

    def call(x, idx, et, w, out, accumulate, wsb, ver, name=None):
        fgnn_b200.mp_forward(nm(x), idx, et, w["filters"], w["bias"], w["scale"], w["shift"], extension=0,
                             aggregator=_lib.AGG_MAX, activation=_lib.ACT_RELU, kernel=kernel, out=nm(out),
                             accumulate=accumulate, workspace=wsb, filters_version=ver, plan=plans.get(name))

    peer_xv = plan.peer_buffers(J, C) if (plan is not None and args.exchange == "peer") else None
    side = [torch.cuda.Stream(device=dev) for _ in range(2)] if (plan is None and args.concurrent_layer) else None

    def step(src):
        """src: dict of device tensors (x_v, x_f, tables).  Returns the final variable features."""
        x_v, x_f = src["x_v"], src["x_f"]
        if halo is not None:
            halo.load_features(src["x_v"], src["x_f"])       # own + halo rows of the step's input features (local gather)
            for l in range(L):
                out = halo.layer(l, src["et_v2f"], src["et_f2v"], W[l], kernel, ws[l], last=(l == L - 1), plans=plans)
            return out
        if peer_xv is not None:
            # the variable features live in the peer-mapped arena: layer l reads buffer l & 1 and the fused
            # exchange kernel leaves the new features in buffer (l + 1) & 1 on every rank
            peer_xv[0].copy_(src["x_v"])
            for l in range(L):
                nf = buf_f[l & 1]
                x_v = plan.layer_peer(l & 1, x_f, src["et_v2f"], src["et_f2v"], W[l], nf, kernel, ws[l], last=(l == L - 1))
                x_f = nf
            plan.peer_wait()
            return x_v
        for l in range(L):
            nv, nf = buf_v[l & 1], buf_f[l & 1]
            if plan is not None:
                plan.layer(x_v, x_f, src["et_v2f"], src["et_f2v"], W[l], nv, nf, kernel, ws[l])
            elif side is None:
                for j in range(J):
                    call(x_v, src["idx_v2f"][j], src["et_v2f"][j], W[l][j]["v2f"], nf[j], False, ws[l][j]["v2f"], ws[l][j]["ver_v2f"], "v2f%d" % j)
                    call(x_f[j], src["idx_f2v"][j], src["et_f2v"][j], W[l][j]["f2v"], nv, j > 0, ws[l][j]["f2v"], ws[l][j]["ver_f2v"], "f2v%d" % j)
            else:
                # The V->F calls of a layer depend on nothing inside the layer and the F->V calls only on each other
                # (type j > 0 adds to what type 0 stored): three concurrent streams.  The persistent kernels each ask for
                # the whole GPU, so a later kernel's CTAs simply start where an earlier one's have finished -- the
                # drain of one call is filled by the next instead of idling behind a stream-order dependency.
                main = torch.cuda.current_stream(dev)
                fork = torch.cuda.Event()
                fork.record(main)
                for j in range(J):                           # V->F of every type: its own stream
                    st = side[j % len(side)]
                    st.wait_event(fork)
                    with torch.cuda.stream(st):
                        call(x_v, src["idx_v2f"][j], src["et_v2f"][j], W[l][j]["v2f"], nf[j], False, ws[l][j]["v2f"], ws[l][j]["ver_v2f"], "v2f%d" % j)
                for j in range(J):                           # F->V chain on the main stream
                    call(x_f[j], src["idx_f2v"][j], src["et_f2v"][j], W[l][j]["f2v"], nv, j > 0, ws[l][j]["f2v"], ws[l][j]["ver_f2v"], "f2v%d" % j)
                for st in side[:min(J, len(side))]:
                    join = torch.cuda.Event()
                    join.record(st)
                    main.wait_event(join)
            x_v, x_f = nv, nf
        return x_v

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def checksum_of(t):
        """Exact, order-independent fingerprint of a float tensor: the integer sum of its bit patterns (after adding
        +0, which turns a -0 into +0: ReLU written as max(v, 0) and as v * 0 differ in the sign of their zeros only)."""
        t = (t + 0.0).contiguous()
        return int(t.view(torch.int16 if t.dtype == torch.bfloat16 else torch.int32).to(torch.int64).sum().item())

    # ---- device-resident timing -----------------------------------------------------------
    for _ in range(2):
        final = step(d_in)
    barrier()
    checksum = checksum_of(final)
    if batch_sharded or halo is not None:        # every rank holds its own codewords / variables: the job's fingerprint is the sum
        cs = torch.tensor([checksum], dtype=torch.int64, device=dev)
        dist.all_reduce(cs)
        checksum = int(cs.item())
    # The step is launch-bound from Python once the kernels are short (20-30 launches of 20-150 us, more so when
    # sharded): capture it once -- streams, events, programmatic launch edges and all -- and replay the graph.
    run_step, graphed, launches_per_step = (lambda: step(d_in)), False, None
    if not args.no_graph and not (world > 1 and args.exchange == "nccl"):
        try:
            graph = torch.cuda.CUDAGraph()
            l0 = fgnn_b200.launch_count()
            with torch.cuda.graph(graph):
                step(d_in)
            launches_per_step = fgnn_b200.launch_count() - l0
            run_step, graphed = graph.replay, True
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"bench: CUDA graph capture failed ({type(e).__name__}: {e}); launching from Python\n")
            torch.cuda.synchronize()
    if world > 1:
        ok = torch.tensor([1 if graphed else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if graphed and not bool(ok.item()):
            run_step, graphed = (lambda: step(d_in)), False
    for _ in range(max(args.warmup, 3)):
        run_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = fgnn_b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        run_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = (launches_per_step * args.steps) if graphed else (fgnn_b200.launch_count() - launches0)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    ms_step = ms / args.steps
    value = msgs_layer * L / (ms_step * 1e-3)

    # ---- the same step back to back for a few seconds: the sustained (power-capped) condition ----------
    sustained = None
    if args.sustain_seconds > 0:
        n_sus = max(args.steps, int(args.sustain_seconds / max(ms_step * 1e-3, 1e-6)) + 1)
        if world > 1:
            tn = torch.tensor([n_sus], device=dev)
            dist.all_reduce(tn, op=dist.ReduceOp.MAX)
            n_sus = int(tn.item())
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
        barrier()
        ev0.record()
        for _ in range(n_sus):
            run_step()
        ev1.record()
        barrier()
        ms_sus = ev0.elapsed_time(ev1)
        if world > 1:
            tt = torch.tensor([ms_sus], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_sus = float(tt.item())
        sustained = {"value": msgs_layer * L * n_sus / (ms_sus * 1e-3), "unit": UNIT, "steps": n_sus, "seconds": ms_sus * 1e-3,
                     "ms_per_step": ms_sus / n_sus, "clocks": sampler2.stop() if rank == 0 else None}

    # ---- end to end through the module API with host buffers --------------------------------
    e2e = None
    if not args.no_e2e and world == 1 and not bf16:
        mods = []
        for l in range(L):
            row = []
            for j in range(J):
                pair = {}
                for d in ("v2f", "f2v"):
                    m = fgnn_b200.mp_conv_v2(C, C, args.edge_types, extension=fgnn_b200.mp_conv_type.NO_EXTENSION,
                                             aggregtor="max")
                    p = weights[l][j][d]
                    m.load_state_dict({"filters": torch.from_numpy(p["filters"]), "bias": torch.from_numpy(p["bias"]),
                                       "bn.weight": torch.ones(C), "bn.bias": torch.zeros(C),
                                       "bn.running_mean": torch.from_numpy(p["rm"]), "bn.running_var": torch.from_numpy(p["rv"]),
                                       "bn.num_batches_tracked": torch.tensor(0)})
                    m.kernel = kernel
                    m.index_check = "async"          # range scan without a host round trip (reference CUDA semantics)
                    if d == "f2v" and not args.no_zero_slots and types[j].pad_f2v.any():
                        m.zero_edge_type_slots = torch.from_numpy(types[j].pad_f2v).to(dev)     # the generator's padding
                    pair[d] = m.to(dev).eval().enable_weight_cache()
                row.append(pair)
            mods.append(row)
        pinned = dict(x_v=to_dev(inp["x_v"], True), x_f=[to_dev(a, True) for a in inp["x_f"]],
                      et_v2f=[to_dev(a, True) for a in inp["et_v2f"]], et_f2v=[to_dev(a, True) for a in inp["et_f2v"]],
                      idx_v2f=[to_dev(a, True) for a in inp["idx_v2f"]], idx_f2v=[to_dev(a, True) for a in inp["idx_f2v"]])
        host_out = torch.empty((B_loc, types[0].n_vars, C), dtype=torch.float32).pin_memory()
        h2d = sum(t.numel() * t.element_size() for v in pinned.values() for t in (v if isinstance(v, list) else [v]))
        d2h = host_out.numel() * 4

        copy_stream, out_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        order = ["x_v"] + [(k, j) for j in range(J) for k in ("idx_v2f", "et_v2f", "x_f", "idx_f2v", "et_f2v")]
        # two sets of device input buffers: step i+1 is uploaded while step i computes
        dsets = [{k: ([torch.empty_like(t, device=dev) for t in v] if isinstance(v, list) else torch.empty_like(v, device=dev))
                  for k, v in pinned.items()} for _ in range(2)]
        ready, free_ev = [{}, {}], [None, None]

        def upload(s):
            """Inputs leave pinned host memory on the copy stream in the order the layer consumes them."""
            if free_ev[s] is not None:
                copy_stream.wait_event(free_ev[s])           # the step that last read this buffer set is done
            with torch.cuda.stream(copy_stream):
                for item in order:
                    if item == "x_v":
                        dsets[s]["x_v"].copy_(pinned["x_v"], non_blocking=True)
                    else:
                        k, j = item
                        dsets[s][k][j].copy_(pinned[k][j], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    ready[s][item] = ev

        def compute(s):
            """20 module calls on the main stream; each waits only for the tensors it reads, so the first calls
            overlap the rest of the upload.  The final variable features go back to pinned host memory."""
            main = torch.cuda.current_stream(dev)
            src = dsets[s]

            def need(*items):
                for it in items:
                    main.wait_event(ready[s][it])

            need("x_v")
            x_v, x_f = nm(src["x_v"]), [None] * J
            with torch.no_grad():
                for l in range(L):
                    nv, nf = None, []
                    for j in range(J):
                        if l == 0:
                            need(("idx_v2f", j), ("et_v2f", j))
                        nf.append(mods[l][j]["v2f"](x_v, src["idx_v2f"][j].expand(B_loc, -1, -1), src["et_v2f"][j]))
                        if l == 0:
                            need(("x_f", j), ("idx_f2v", j), ("et_f2v", j))
                            x_f[j] = nm(src["x_f"][j])
                        y = mods[l][j]["f2v"](x_f[j], src["idx_f2v"][j].expand(B_loc, -1, -1), src["et_f2v"][j])
                        nv = y if nv is None else nv + y
                    x_v, x_f = nv, nf
            done = torch.cuda.Event()
            done.record(main)
            free_ev[s] = done
            out_stream.wait_event(done)
            with torch.cuda.stream(out_stream):
                host_out.copy_(x_v[..., 0].permute(0, 2, 1), non_blocking=True)
            x_v.record_stream(out_stream)

        def e2e_run(n):
            upload(0)
            for i in range(n):
                if i + 1 < n:
                    upload((i + 1) & 1)
                compute(i & 1)
            torch.cuda.synchronize()

        if world == 1:
            e2e_run(2)
            k = max(3, min(args.steps, 10))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_run(k)
            fgnn_b200.check_async_errors()
            dt = (time.perf_counter() - t0) / k
            # the link itself: the same uploads with nothing else running
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(3):
                upload(0)
            torch.cuda.synchronize()
            h2d_gbs = 3 * h2d / (time.perf_counter() - t1) / 1e9
            e2e = {"value": msgs_layer * L / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3, "steps": k, "index_check": "async",
                   "h2d_link_gbs": h2d_gbs,
                   "api": "fgnn_b200.mp_conv_v2.forward (20 module calls per step) on pinned host tensors: every step's inputs "
                          "are uploaded (copy stream, double-buffered: step i+1 uploads while step i computes; each call waits "
                          "only for its own inputs) and every step's variable features are read back to pinned host memory"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = hbm_peak()
    # roofline fractions are PER GPU: the job's algorithmic bytes (or flops) over N GPUs' worth of peak
    achieved = bytes_layer * L / (ms_step * 1e-3) / 1e9 / world
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath) and args.edge_types == 16 and world == 1 and args.config in (None, "cfg2") and args.src_calls == "auto" \
            and args.vars == 100_000 and not bf16:
        traffic = json.load(open(tpath))["traffic_bytes_per_launch_avg"]     # ncu --set full of this very workload
    tpeak, tpeak_src = tensor_peak()
    executed_rows_plans = plans if plan is None else {}
    mma_flops = (executed_mma_flops_per_layer(args, types, executed_rows_plans) * args.batch) if args.kernel != "simt" else 0
    exec_tflops = mma_flops * L / (ms_step * 1e-3) / 1e12 / world
    n_calls = 2 * J * L
    if batch_sharded:
        par = "batch-sharded x%d: every rank runs its own codewords, no collective (replicas)" % world
    elif halo is not None:
        par = ("owner-computes sharding x%d (contiguous variable / factor ranges%s) + one NVLink peer-memory halo pull per layer"
               % (world, ", factors ordered by smallest variable" if (args.order == "locality" and (args.local_band or args.point_cloud)) else ""))
    elif world > 1:
        par = "factor-sharded x%d + %s per layer" % (world, "fused max-reduce/epilogue/broadcast kernel over NVLink peer memory"
                                                      if args.exchange == "peer" else "NCCL max-all-reduce")
    else:
        par = "single GPU"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong",                # the workload is the same graph (or batch) whatever N is
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "checksum": checksum,
        "config": {"workload": workload_name(args), "messages_per_layer": msgs_layer, "layers": L,
                   "kernel": args.kernel, "source_stationary_calls": sorted(plans),
                   "launch": ("CUDA graph of the step, replayed" if graphed else "from Python, call by call") +
                             ("" if side is None else "; the V->F calls of a layer on their own streams beside the F->V chain"),
                   "l2": "per-step working set (~%d MB/layer) exceeds the 126 MB L2; no explicit flush" % (bytes_layer // 1_000_000),
                   "gather": "cp.async row gather (TMA tile::gather4 measured slower: 64 requests of 512 B per item are "
                             "request-rate bound, profiles/r01_early/r01_tc_gather4_variant_full.txt); edge types of a "
                             "source-stationary tile by one TMA bulk copy",
                   "parallelism": par},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "per_gpu": True, "traffic": traffic,
                     "traffic_note": ("DRAM bytes per core call from ncu --set full (profiles/r02_traffic.json). Above the algorithmic bytes on "
                                      "purpose: the source-stationary calls write and re-read one O-wide message per edge (2*4*O bytes) to "
                                      "cut the tensor work") if traffic else None,
                     "peak_source": peak_src,
                     "kernel": ("mp_src_kernel + mp_reduce_kernel / mp_tc_kernel (tcgen05; average over the step's %d core calls, "
                                "a source-stationary call being two launches)" % n_calls)
                     if args.kernel != "simt" else "mp_simt_kernel",
                     "algorithmic_bytes_per_launch": bytes_layer * L / n_calls / world,
                     "avg_launch_us": ms_step * 1e3 / n_calls,
                     "tensor": {"executed_tflops": exec_tflops, "peak": tpeak, "unit": "TFLOP/s",
                                "frac": exec_tflops / tpeak if tpeak else None, "peak_source": tpeak_src, "per_gpu": True,
                                "note": "bf16 MMA flops actually issued per GPU (3 split-bf16 terms for fp32 parity; one row-product per "
                                        "virtual source row in source-stationary calls, per slot otherwise) against the sustained cuBLAS "
                                        "bf16 rate: at T=16 the tensor pipe and tensor-memory reads, not HBM, bound pass 1 (DESIGN.md 3)"}},
        "gpu_launches": int(launches), "clocks": clocks_record(clocks, sustained, ms),
    }
    if sustained is not None:
        sustained["roofline_frac"] = bytes_layer * L / (sustained["ms_per_step"] * 1e-3) / 1e9 / world / peak
        line["sustained"] = sustained
    if halo is not None:
        line["exchange"] = {"kind": "halo pull", "halo_bytes_per_layer_rank0": int(halo.halo_bytes_per_layer),
                            "halo_rows_rank0": {"variables": int(len(halo.var_halo)), "factors": [int(len(h)) for h in halo.fac_halo]},
                            "owned_rows_rank0": {"variables": int(halo.n_own_v), "factors": [int(n) for n in halo.n_own_f]}}
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_aten_baseline and not bf16:
        line["gpu_aten_baseline"] = gpu_aten_baseline(args, types, inp, weights[0])
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()

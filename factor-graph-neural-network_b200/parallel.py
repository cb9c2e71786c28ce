"""Factor-sharded FGNN layer over the GPUs of one box (one process per GPU).

The factors of every type are split into `world` contiguous ranges.  Per layer and rank:

  V->F   dst = this rank's factors, src = all variables (replicated): the local rows of the
         reference table `idx_v2f [F, K]`; no communication.
  F->V   dst = variables, src = this rank's factors.  Only slots naming a local factor are live, so
         the rank builds a COMPACTED table: the variables it touches, sorted by live-slot count
         (descending), live slots first, `tile_slots` = slots per 128-row tile, `out_rows` = the
         variable each row stands for.  The kernel writes the raw max (no bias/BN/activation) of the
         local slots into a [N, J*O] buffer pre-filled with -inf; ONE all_reduce(MAX) over ranks;
         then fgnn_epilogue_forward per type (the epilogue is non-linear, so it runs after the
         reduce) and the sum over types (FactorNN: nfeature += nv, factor_mpnn_sp.py:147).

Work per rank is (sum of live slots) ~ 1/world of the single-GPU F->V work; the all-reduce volume
(N * J * O floats) does not shrink with `world` for uniform-random incidence -- every variable is a
boundary variable.  The result equals the single-GPU layer bit for bit: max is
exact and order-independent, every slot value is computed by the same kernel arithmetic.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .mp_nn import mp_forward

TILE = 128


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of `n` items owned by `rank`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class LocalF2V:
    """Compacted shard-local Factor->Variable table of one factor type (host side, numpy)."""

    def __init__(self, idx_f2v, pad_f2v, f_lo, f_hi):
        idx_f2v = np.asarray(idx_f2v)
        N, Kv = idx_f2v.shape
        live = (idx_f2v >= f_lo) & (idx_f2v < f_hi)
        # the reference's padding slots (valid index 0, all-zero edge type) contribute a 0-valued message
        # to the max; they are ordinary slots here and stay live on the rank that owns the factor they name
        count = live.sum(1)
        rows = np.nonzero(count > 0)[0]
        order = rows[np.argsort(-count[rows], kind="stable")]          # variables, most live slots first
        kmax = int(count.max()) if count.size and count.max() > 0 else 1
        self.n_rows = int(order.size)
        # per row: positions of its live slots (in the original table), live first
        pos = np.argsort(~live[order], axis=1, kind="stable")[:, :kmax]              # [rows, kmax]
        self.slot_pos = pos.astype(np.int64)
        self.slot_live = np.take_along_axis(live[order], pos, axis=1)
        local = np.take_along_axis(idx_f2v[order], pos, axis=1) - f_lo
        self.idx = np.where(self.slot_live, local, -1).astype(np.int64)              # [rows, kmax], -1 = empty
        self.var = order.astype(np.int32)                                            # out_rows
        self.kmax = kmax
        n_tiles = (self.n_rows + TILE - 1) // TILE
        cnt_sorted = count[order]
        self.tile_slots = np.array([int(cnt_sorted[t * TILE]) for t in range(n_tiles)], dtype=np.int32)
        self.live_slots = int(count.sum())

    def gather_etype(self, et):
        """etype [B,T,N,Kv] (torch, device) -> local [B,T,rows,kmax] matching self.idx (empty slots get 0)."""
        dev = et.device
        var = torch.from_numpy(self.var.astype(np.int64)).to(dev)
        pos = torch.from_numpy(self.slot_pos).to(dev)
        sel = et.index_select(2, var)                                               # [B,T,rows,Kv]
        out = torch.gather(sel, 3, pos[None, None].expand(sel.shape[0], sel.shape[1], -1, -1))
        return (out * torch.from_numpy(self.slot_live).to(dev)[None, None]).contiguous()


class ShardedLayerPlan:
    """Everything rank `rank` of `world` needs to run FGNN layers on its factor shard.

    `types` = list of fgnn_b200.graphs.FactorType (full graph, host); tables are built once.
    """

    def __init__(self, types, rank, world, device, group=None, comm_sms=16):
        self.rank, self.world, self.device, self.group = rank, world, device, group
        # SMs left to the collective while V->F runs beside it (the tensor-core kernel owns whole SMs)
        self.comm_sms = comm_sms
        self.n_sms = torch.cuda.get_device_properties(device).multi_processor_count if device.type == "cuda" else 0
        self.types = types
        self.n_vars = types[0].n_vars
        self.ranges = [shard_range(t.n_factors, rank, world) for t in types]
        self.idx_v2f, self.f2v = [], []
        for t, (lo, hi) in zip(types, self.ranges):
            self.idx_v2f.append(torch.from_numpy(np.ascontiguousarray(t.idx_v2f[lo:hi][None])).to(device))
            self.f2v.append(LocalF2V(t.idx_f2v, t.pad_f2v, lo, hi))
        self.idx_f2v = [torch.from_numpy(l.idx[None]).to(device) for l in self.f2v]
        self.tile_slots = [torch.from_numpy(l.tile_slots).to(device) for l in self.f2v]
        self.out_rows = [torch.from_numpy(l.var).to(device) for l in self.f2v]
        self._raw = None
        self._et_cache = {}

    # -- inputs ------------------------------------------------------------------------------
    def local_factor_features(self, x_f_full):
        """Slice full per-type factor features [1,F,C] (node-major) to this rank's shard."""
        return [x[:, lo:hi].contiguous() for x, (lo, hi) in zip(x_f_full, self.ranges)]

    def local_etypes(self, et_v2f_full, et_f2v_full):
        """Edge types of the local tables (layer-invariant: call once per forward)."""
        ev = [e[:, :, lo:hi].contiguous() for e, (lo, hi) in zip(et_v2f_full, self.ranges)]
        ef = [l.gather_etype(e) for l, e in zip(self.f2v, et_f2v_full)]
        return ev, ef

    # -- one layer ---------------------------------------------------------------------------
    def layer(self, x_v, x_f_local, et_v2f_local, et_f2v_local, weights, out_v, out_f_local, kernel=_lib.KERNEL_AUTO,
              workspaces=None):
        """x_v [1,N,C] node-major (replicated), x_f_local[j] [1,F_j_local,C]; weights[j][dir] = dict(filters,
        bias, scale, shift).  Writes the new variable features to out_v [1,N,O] and the new local factor
        features to out_f_local[j].  Returns out_v.

        Order: F->V partial maxima of all types, the all-reduce (asynchronous, on NCCL's stream), the
        V->F calls (no communication; they overlap the reduce), wait, fused epilogue + sum over types."""
        J = len(self.types)
        O = weights[0]["f2v"]["filters"].shape[1] // et_f2v_local[0].shape[1]
        nm = lambda t: t.permute(0, 2, 1).unsqueeze(-1)
        if self._raw is None or self._raw.shape[-1] != J * O:
            self._raw = torch.empty((1, self.n_vars, J * O), dtype=torch.float32, device=self.device)
        raw = self.raw = self._raw
        raw.fill_(float("-inf"))
        for j in range(J):
            if self.f2v[j].n_rows > 0:
                w = weights[j]["f2v"]
                wsj = workspaces[j] if workspaces is not None else {}
                view = raw[:, :, j * O:(j + 1) * O]                      # [1,N,O] slice of the [1,N,J*O] buffer
                mp_forward(nm(x_f_local[j]), self.idx_f2v[j], et_f2v_local[j], w["filters"], None, None, None,
                           extension=0, aggregator=_lib.AGG_MAX, activation=_lib.ACT_NONE, kernel=kernel,
                           mask_negative=True, out=nm(view), tile_slots=self.tile_slots[j], out_rows=self.out_rows[j],
                           workspace=wsj.get("f2v"), filters_version=wsj.get("ver_f2v", 0))
        work = None
        if self.world > 1 and torch.distributed.is_initialized():
            work = torch.distributed.all_reduce(raw, op=torch.distributed.ReduceOp.MAX, group=self.group, async_op=True)
        for j in range(J):
            if x_f_local[j].shape[1] > 0:
                w = weights[j]["v2f"]
                wsj = workspaces[j] if workspaces is not None else {}
                mp_forward(nm(x_v), self.idx_v2f[j], et_v2f_local[j], w["filters"], w["bias"], w["scale"], w["shift"],
                           extension=0, aggregator=_lib.AGG_MAX, activation=_lib.ACT_RELU, kernel=kernel,
                           out=nm(out_f_local[j]), workspace=wsj.get("v2f"), filters_version=wsj.get("ver_v2f", 0),
                           sm_limit=(self.n_sms - self.comm_sms) if work is not None else 0)
        if work is not None:
            work.wait()
        return self.finish(raw, weights, out_v)

    def finish(self, raw, weights, out_v):
        """out_v = sum_j ReLU(BN_j(raw_j + bias_j)) on the (reduced) raw aggregate [1,N,J*O]: one kernel."""
        J = len(self.types)
        O = raw.shape[-1] // J
        key = tuple(w["f2v"]["bias"].data_ptr() for w in weights)
        if getattr(self, "_epi_key", None) != key:
            self._epi = [torch.cat([w["f2v"][k] for w in weights]).contiguous() for k in ("bias", "scale", "shift")]
            self._epi_key = key
        bias, scale, shift = self._epi
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.lib().fgnn_epilogue_sum_forward(raw.data_ptr(), out_v.data_ptr(), self.n_vars, O, J, bias.data_ptr(),
                                                        scale.data_ptr(), shift.data_ptr(), _lib.ACT_RELU, 0.0, 0, stream),
                   "epilogue_sum_forward")
        return out_v

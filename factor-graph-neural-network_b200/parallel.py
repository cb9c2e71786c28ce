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
        # live slots that are the reference's padding (valid index + all-zero edge type): a source-stationary plan of this
        # table leaves them out (SourcePlan(zero_slots=...)) -- they all name the same factor, on the rank that owns it
        self.slot_pad = np.take_along_axis(np.asarray(pad_f2v, dtype=bool)[order], pos, axis=1) & self.slot_live
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


class _DevMem:
    """Raw device memory exposed through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerExchange:
    """The cross-GPU step of the factor-sharded layer as ONE kernel over NVLink peer memory
    (csrc/exchange.cu, `fgnn_exchange_forward`): the owner of a row range reads the peers' raw per-type maxima
    straight out of their memory, reduces, applies bias / BN / activation per type, sums the types and stores the
    finished rows into every rank's next-layer feature buffer.  Replaces all_reduce(MAX) + epilogue kernel.

    Each rank's ARENA (library-allocated, IPC-exported) holds: 16 epoch flags + block counter (256 B header), TWO raw
    aggregates [rows, J*O] (layer l fills raws[l & 1], so the exchange of layer l can still be reading while the F->V
    calls of layer l+1 write the other one) and two next-layer feature buffers [rows, O] (ping-pong: layer l reads
    xv[l & 1], the exchange of layer l fills xv[(l + 1) & 1] on every rank).
    """
    HEADER = 256

    def __init__(self, rows, J, O, rank, world, device, group=None, ctas=64, peers=None):
        self.rows, self.J, self.O, self.rank, self.world, self.device, self.ctas = rows, J, O, rank, world, device, ctas
        self.raw_offs = [self.HEADER, self._align(self.HEADER + rows * J * O * 4)]
        self.raw_off = self.raw_offs[0]
        self.xv_off = [self._align(self.raw_offs[1] + rows * J * O * 4)]
        self.xv_off.append(self._align(self.xv_off[0] + rows * O * 4))
        self.nbytes = self._align(self.xv_off[1] + rows * O * 4)
        lib = _lib.lib()
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            _lib.check(lib.fgnn_comm_alloc(self.nbytes, ctypes.byref(ptr), handle), "comm_alloc")
        self.base = ptr.value
        self._opened = []
        if peers is not None:                     # same-process ranks (tests): arena base pointers given directly
            self.peer_base = list(peers)
            self.peer_base[rank] = self.base
        elif world == 1:
            self.peer_base = [self.base]
        else:
            handles = [None] * world
            torch.distributed.all_gather_object(handles, bytes(handle.raw), group=group)
            self.peer_base = []
            for q, h in enumerate(handles):
                if q == rank:
                    self.peer_base.append(self.base)
                    continue
                pp = ctypes.c_void_p()
                with torch.cuda.device(device):
                    _lib.check(lib.fgnn_comm_open(h, ctypes.byref(pp)), "comm_open")
                self.peer_base.append(pp.value)
                self._opened.append(pp.value)
        mem = torch.as_tensor(_DevMem(self.base, self.nbytes), device=device)
        self._mem = mem
        self.raws = [mem[o:o + rows * J * O * 4].view(torch.float32).view(1, rows, J * O) for o in self.raw_offs]
        self.raw = self.raws[0]
        self.xv = [mem[o:o + rows * O * 4].view(torch.float32).view(1, rows, O) for o in self.xv_off]
        self.row0, self.row1 = shard_range(rows, rank, world)
        self.epoch = 0

    @staticmethod
    def _align(n):
        return (n + 255) // 256 * 256

    def set_peers(self, bases):
        self.peer_base = list(bases)

    def forward(self, dst, bias, scale, shift, activation=_lib.ACT_RELU, slope=0.0, stream=None, raw_index=0,
                raw_mask=None, out_mask=None):
        """Launch the exchange of raw buffer `raw_index` into xv[dst] of every rank (asynchronous).
        raw_mask / out_mask: uint32-as-int32 [rows] device tensors (fgnn_exchange_args), or None = dense.
        The kernel numbers its launches itself (epoch 0), so the launch can be captured in a CUDA graph."""
        a = _lib.ExchangeArgs()
        for q in range(self.world):
            a.raw[q] = self.peer_base[q] + self.raw_offs[raw_index]
            a.out[q] = self.peer_base[q] + self.xv_off[dst]
            a.flags[q] = self.peer_base[q]
        a.counter = self.base + 128
        a.bias = bias.data_ptr() if bias is not None else None
        a.bn_scale = scale.data_ptr() if scale is not None else None
        a.bn_shift = shift.data_ptr() if shift is not None else None
        a.rows, a.row0, a.row1 = self.rows, self.row0, self.row1
        a.world, a.rank, a.J, a.O = self.world, self.rank, self.J, self.O
        a.activation, a.act_slope, a.epoch, a.ctas = int(activation), float(slope), 0, self.ctas
        a.raw_mask = raw_mask.data_ptr() if raw_mask is not None else None
        a.out_mask = out_mask.data_ptr() if out_mask is not None else None
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().fgnn_exchange_forward(ctypes.byref(a), ctypes.c_void_p(st.cuda_stream)), "exchange_forward")
        return self.xv[dst]

    def close(self):
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for pp in self._opened:
                lib.fgnn_comm_close(ctypes.c_void_p(pp))
            self._opened = []
            if self.base:
                self.raw = self.raws = self.xv = self._mem = None
                lib.fgnn_comm_free(ctypes.c_void_p(self.base))
                self.base = 0


def owner_of(ids, n, world):
    """Rank owning each of `ids` under shard_range(n, ., world)."""
    bounds = np.array([shard_range(n, q, world)[1] for q in range(world)], dtype=np.int64)
    return np.searchsorted(bounds, np.asarray(ids, dtype=np.int64), side="right")


def sharded_instance_norm_act(x_local, n_total, group=None, eps=1e-5, activation=None, reduce=None):
    """InstanceNorm2d (affine = False) + activation (None | 'relu') of a [B,C,N_local,1] fp32 CUDA tensor whose N
    axis is split over the ranks of `group` (FactorNN's f2f / v2v maps on sharded factor / variable features,
    base_model.py:83-90): local partial sums -> all-reduce of 2 x C floats per instance -> normalise.  Two-pass
    variance, like the single-GPU kernel.  `reduce` replaces the all-reduce (tests: sum over simulated shards)."""
    lib = _lib.lib()
    B, C, N, W = x_local.shape
    assert W == 1 and x_local.is_cuda and x_local.dtype == torch.float32
    dev = x_local.device
    if reduce is None:
        def reduce(t):
            if torch.distributed.is_initialized():
                torch.distributed.all_reduce(t, group=group)
            return t
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    sb, sc, sn = x_local.stride(0), x_local.stride(1), x_local.stride(2)
    out = torch.empty_like(x_local)
    with torch.cuda.device(dev):
        s1 = torch.empty((B, C), dtype=torch.float32, device=dev)
        _lib.check(lib.fgnn_instance_norm_partial(p(x_local), None, p(s1), B, C, N, sb, sc, sn, st()), "instance_norm_partial")
        mean = reduce(s1) / float(n_total)
        s2 = torch.empty((B, C), dtype=torch.float32, device=dev)
        _lib.check(lib.fgnn_instance_norm_partial(p(x_local), p(mean), p(s2), B, C, N, sb, sc, sn, st()), "instance_norm_partial")
        inv = torch.rsqrt(reduce(s2) / float(n_total) + eps)
        _lib.check(lib.fgnn_instance_norm_apply(p(x_local), p(out), p(mean), p(inv), B, C, N, sb, sc, sn, out.stride(0), out.stride(1),
                                                out.stride(2), _lib.ACT_RELU if activation == "relu" else _lib.ACT_NONE, 0.0, st()),
                   "instance_norm_apply")
    return out


class HaloPartition:
    """Host side (numpy) of the owner-computes sharding: rank `rank` of `world` owns the variables
    shard_range(N, rank, world) and, per type, the factors shard_range(F_j, rank, world).  Local numbering of a
    feature buffer: owned rows first, then the halo rows (sorted by global id).

        idx_v2f[j]  [F_own_j, K]   local VARIABLE rows read by my factors           (V->F call, destinations = my factors)
        idx_f2v[j]  [N_own, Kv]    local FACTOR rows of type j read by my variables  (F->V call, destinations = my variables)
        var_halo / fac_halo[j]     global ids of the halo rows;  src_v / src_f[j] = (owner rank, row in the owner's buffer)
    """

    def __init__(self, types, rank, world):
        N = types[0].n_vars
        self.rank, self.world = rank, world
        self.v0, self.v1 = shard_range(N, rank, world)
        self.fr = [shard_range(t.n_factors, rank, world) for t in types]
        n_own = self.v1 - self.v0
        tv = [np.asarray(t.idx_v2f[f0:f1]) for t, (f0, f1) in zip(types, self.fr)]
        outside = [a[(a < self.v0) | (a >= self.v1)] for a in tv]
        self.var_halo = np.unique(np.concatenate(outside)) if outside else np.zeros(0, np.int64)

        def to_local(a, lo, hi, halo, n_owned):
            a = np.asarray(a, dtype=np.int64)
            own = (a >= lo) & (a < hi)
            return np.where(own, a - lo, n_owned + np.searchsorted(halo, a)).astype(np.int64)

        self.idx_v2f = [to_local(a, self.v0, self.v1, self.var_halo, n_own) for a in tv]
        self.fac_halo, self.idx_f2v = [], []
        for t, (f0, f1) in zip(types, self.fr):
            a = np.asarray(t.idx_f2v[self.v0:self.v1])
            h = np.unique(a[(a < f0) | (a >= f1)])
            self.fac_halo.append(h)
            self.idx_f2v.append(to_local(a, f0, f1, h, f1 - f0))
        self.n_own_v, self.n_own_f = n_own, [f1 - f0 for f0, f1 in self.fr]
        self.rows_v = n_own + len(self.var_halo)
        self.rows_f = [n + len(h) for n, h in zip(self.n_own_f, self.fac_halo)]

        def src_of(halo, n):
            own = owner_of(halo, n, world)
            starts = np.array([shard_range(n, q, world)[0] for q in range(world)], dtype=np.int64)
            return own.astype(np.uint8), (halo - starts[own]).astype(np.int32)
        self.src_v = src_of(self.var_halo, N)
        self.src_f = [src_of(h, t.n_factors) for h, t in zip(self.fac_halo, types)]


class HaloLayerPlan:
    """Owner-computes sharding with feature halos (halo-restricted exchange).

    Rank r owns the variables shard_range(N, r, G) and, per type, the factors shard_range(F_j, r, G).  It evaluates
    ALL slots of its own destinations -- the V->F call of its factors, the F->V calls of its variables -- with the
    same kernel and the same slot order as the single-GPU layer, so every row is bit-identical to it.  The sources it
    reads but does not own are its HALO: the feature buffers hold [owned rows | halo rows], the index tables are
    renumbered into that local order, and after every layer one kernel (`fgnn_halo_pull`, csrc/exchange.cu) copies the
    halo rows of the new features out of their owners' arenas over NVLink.  What crosses the links is therefore
    bounded by the graph's cut: with a locality-preserving factor order (`graphs.locality_order`) a banded graph
    exchanges a few hundred rows per layer; uniform-random incidence still needs most rows (reported as measured).

    Layer l reads buffer set l & 1 and writes set (l + 1) & 1 (owned rows by the calls, halo rows by the pull).
    """
    HEADER = 256

    def __init__(self, types, rank, world, device, dtype=torch.float32, C=64, ctas=32):
        self.types, self.rank, self.world, self.device, self.dtype, self.C, self.ctas = types, rank, world, device, dtype, C, ctas
        J = len(types)
        N = types[0].n_vars
        part = HaloPartition(types, rank, world)
        self.part = part
        self.v0, self.v1, self.fr = part.v0, part.v1, part.fr
        self.var_halo, self.fac_halo = part.var_halo, part.fac_halo
        self.n_own_v, self.n_own_f, self.rows_v, self.rows_f = part.n_own_v, part.n_own_f, part.rows_v, part.rows_f
        n_own = self.n_own_v
        self.idx_v2f = [torch.from_numpy(a[None]).to(device) for a in part.idx_v2f]
        self.idx_f2v = [torch.from_numpy(a[None]).to(device) for a in part.idx_f2v]
        dev_pair = lambda pr: (torch.from_numpy(pr[0]).to(device), torch.from_numpy(pr[1]).to(device))
        self.src_v = dev_pair(part.src_v)
        self.src_f = [dev_pair(pr) for pr in part.src_f]
        # ---- arena: header | two sets of [xv, xf_0 .. xf_J-1]
        esz = 2 if dtype == torch.bfloat16 else 4
        self.row_bytes = C * esz
        off, self.off_v, self.off_f = self.HEADER, [], []
        for _ in range(2):
            self.off_v.append(off)
            off = self._align(off + self.rows_v * self.row_bytes)
            fs = []
            for r in self.rows_f:
                fs.append(off)
                off = self._align(off + max(r, 1) * self.row_bytes)
            self.off_f.append(fs)
        self.nbytes = off
        lib = _lib.lib()
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            _lib.check(lib.fgnn_comm_alloc(self.nbytes, ctypes.byref(ptr), handle), "comm_alloc")
        self.base, self.handle = ptr.value, bytes(handle.raw)
        self._opened = []
        mem = torch.as_tensor(_DevMem(self.base, self.nbytes), device=device)
        self._mem = mem
        view = lambda o, rows: mem[o:o + rows * self.row_bytes].view(dtype).view(1, rows, C)
        self.xv = [view(self.off_v[p], self.rows_v) for p in range(2)]
        self.xf = [[view(self.off_f[p][j], self.rows_f[j]) for j in range(J)] for p in range(2)]
        self.peers = None                                     # per rank: (base, off_v, off_f), see connect()
        self._side = torch.cuda.Stream(device=device)
        self._halo_v_dev = torch.from_numpy(self.var_halo).to(device)
        self._halo_f_dev = [torch.from_numpy(h).to(device) for h in self.fac_halo]
        self.n_sms = torch.cuda.get_device_properties(device).multi_processor_count
        self.halo_bytes_per_layer = (len(self.var_halo) + sum(len(h) for h in self.fac_halo)) * self.row_bytes
        self.exchange_stats = None

    @staticmethod
    def _align(n):
        return (n + 255) // 256 * 256

    def info(self):
        """What the other ranks need to address this rank's arena."""
        return {"handle": self.handle, "base": self.base, "off_v": self.off_v, "off_f": self.off_f}

    def connect(self, infos, same_process=False):
        """infos[q] = info() of rank q (all_gather_object across processes; passed directly when the ranks live in one
        process, e.g. the single-device tests)."""
        lib = _lib.lib()
        self.peers = []
        for q, inf in enumerate(infos):
            if q == self.rank or same_process:
                base = self.base if q == self.rank else inf["base"]
            else:
                pp = ctypes.c_void_p()
                with torch.cuda.device(self.device):
                    _lib.check(lib.fgnn_comm_open(inf["handle"], ctypes.byref(pp)), "comm_open")
                base = pp.value
                self._opened.append(base)
            self.peers.append((base, inf["off_v"], inf["off_f"]))

    def connect_distributed(self, group=None):
        infos = [None] * self.world
        torch.distributed.all_gather_object(infos, self.info(), group=group)
        self.connect(infos)

    # -- inputs ------------------------------------------------------------------------------
    def load_features(self, x_v_full, x_f_full):
        """Fill buffer set 0 from full (replicated) feature arrays [1,N,C] / [1,F_j,C]: owned rows, then halo rows."""
        dev = self.device
        self.xv[0][:, :self.n_own_v].copy_(x_v_full[:, self.v0:self.v1])
        if len(self.var_halo):
            torch.index_select(x_v_full, 1, self._halo_v_dev, out=self.xv[0][:, self.n_own_v:])
        for j, (f0, f1) in enumerate(self.fr):
            self.xf[0][j][:, :self.n_own_f[j]].copy_(x_f_full[j][:, f0:f1])
            if len(self.fac_halo[j]):
                torch.index_select(x_f_full[j], 1, self._halo_f_dev[j], out=self.xf[0][j][:, self.n_own_f[j]:])

    def local_etypes(self, et_v2f_full, et_f2v_full):
        ev = [e[:, :, f0:f1].contiguous() for e, (f0, f1) in zip(et_v2f_full, self.fr)]
        ef = [e[:, :, self.v0:self.v1].contiguous() for e in et_f2v_full]
        return ev, ef

    # -- one layer ---------------------------------------------------------------------------
    def pull(self, dst_set, stream=None, what=("v", "f")):
        """Copy the halo rows of buffer set `dst_set` from their owners (asynchronous; all ranks call it alike)."""
        a = _lib.HaloArgs()
        jobs = []
        if "v" in what:
            jobs.append((self.off_v[dst_set], [pb + ov[dst_set] for pb, ov, _ in self.peers], self.src_v, self.n_own_v))
        if "f" in what:
            for j in range(len(self.types)):
                jobs.append((self.off_f[dst_set][j], [pb + of[dst_set][j] for pb, _, of in self.peers], self.src_f[j], self.n_own_f[j]))
        for i, (off, srcs, (s_rank, s_row), n_own) in enumerate(jobs):
            jb = a.jobs[i]
            for q in range(self.world):
                jb.src[q] = srcs[q]
            jb.dst = self.base + off
            jb.src_row, jb.src_rank = s_row.data_ptr(), s_rank.data_ptr()
            jb.dst_row0, jb.n, jb.row_bytes = n_own, s_row.numel(), self.row_bytes
        for q in range(self.world):
            a.flags[q] = self.peers[q][0]
        a.counter = self.base + 128
        a.n_jobs, a.world, a.rank, a.epoch, a.ctas = len(jobs), self.world, self.rank, 0, self.ctas
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().fgnn_halo_pull(ctypes.byref(a), ctypes.c_void_p(st.cuda_stream)), "halo_pull")

    def layer(self, l, et_v2f_local, et_f2v_local, weights, kernel=_lib.KERNEL_AUTO, workspaces=None, last=False, plans=None):
        """Layer l: reads buffer set l & 1, leaves the new features in set (l + 1) & 1 -- owned rows by the calls, halo
        rows by the pulls.  Order: the V->F calls, the pull of the new factor halos on a side stream BESIDE the F->V
        calls (which read the old factor features), then the pull of the new variable halos."""
        p, q = l & 1, (l + 1) & 1
        nm = lambda t: t.permute(0, 2, 1).unsqueeze(-1)
        main = torch.cuda.current_stream(self.device)
        plans = plans or {}
        for j in range(len(self.types)):
            if self.n_own_f[j] == 0:
                continue
            w = weights[j]["v2f"]
            wsj = workspaces[j] if workspaces is not None else {}
            mp_forward(nm(self.xv[p]), self.idx_v2f[j], et_v2f_local[j], w["filters"], w["bias"], w["scale"], w["shift"],
                       extension=0, aggregator=_lib.AGG_MAX, activation=_lib.ACT_RELU, kernel=kernel,
                       out=nm(self.xf[q][j][:, :self.n_own_f[j]]), workspace=wsj.get("v2f"), filters_version=wsj.get("ver_v2f", 0),
                       plan=plans.get("v2f%d" % j), validate=False)
        if not last and self.world > 1:
            ready = torch.cuda.Event()
            ready.record(main)
            self._side.wait_event(ready)
            self.pull(q, stream=self._side, what=("f",))
            done_f = torch.cuda.Event()
            done_f.record(self._side)
        # the F->V calls leave the pull its SMs (two 512-thread CTAs per SM) while it runs beside them
        sms = self.n_sms - (self.ctas + 1) // 2 if (not last and self.world > 1) else 0
        for j in range(len(self.types)):
            w = weights[j]["f2v"]
            wsj = workspaces[j] if workspaces is not None else {}
            mp_forward(nm(self.xf[p][j]), self.idx_f2v[j], et_f2v_local[j], w["filters"], w["bias"], w["scale"], w["shift"],
                       extension=0, aggregator=_lib.AGG_MAX, activation=_lib.ACT_RELU, kernel=kernel,
                       out=nm(self.xv[q][:, :self.n_own_v]), accumulate=j > 0, workspace=wsj.get("f2v"),
                       filters_version=wsj.get("ver_f2v", 0), plan=plans.get("f2v%d" % j), sm_limit=sms, validate=False)
        if not last and self.world > 1:
            main.wait_event(done_f)
            self.pull(q, what=("v",))
        return self.xv[q][:, :self.n_own_v]

    def close(self):
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for pp in self._opened:
                lib.fgnn_comm_close(ctypes.c_void_p(pp))
            self._opened = []
            if self.base:
                self.xv = self.xf = self._mem = None
                lib.fgnn_comm_free(ctypes.c_void_p(self.base))
                self.base = 0


class ShardedLayerPlan:
    """Everything rank `rank` of `world` needs to run FGNN layers on its factor shard.

    `types` = list of fgnn_b200.graphs.FactorType (full graph, host); tables are built once.
    """

    def __init__(self, types, rank, world, device, group=None, comm_sms=16, exchange="nccl", exchange_ctas=64):
        self.rank, self.world, self.device, self.group = rank, world, device, group
        # "nccl": all_reduce(MAX) + epilogue kernel; "peer": one fused kernel over NVLink peer memory (PeerExchange)
        self.exchange = exchange
        self.exchange_ctas = exchange_ctas        # 512-thread CTAs of the exchange kernel, two per SM
        # The V->F calls of layer l always run beside the exchange of layer l and leave it its SMs.  By the time the
        # F->V calls of layer l+1 start, that exchange has normally finished (it is shorter than the V->F calls), so
        # they take the whole GPU unless told otherwise; CTAs that find an SM still taken simply start a little later.
        self.limit_f2v = False
        self.px = None
        self._comm_stream = None
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
        # static sparsity of the shards for the peer exchange: which (rank, type) touches which variable.  Every rank
        # holds the whole (host) graph, so the masks need no communication.
        self.raw_mask = self.out_mask = None
        if exchange == "peer" and world > 1 and world * len(types) <= 32:
            raw_bits = np.zeros(self.n_vars, dtype=np.uint32)
            out_bits = np.zeros(self.n_vars, dtype=np.uint32)
            for q in range(world):
                for j, t in enumerate(types):
                    lo, hi = shard_range(t.n_factors, q, world)
                    touched = ((np.asarray(t.idx_f2v) >= lo) & (np.asarray(t.idx_f2v) < hi)).any(1)
                    raw_bits[touched] |= np.uint32(1 << (q * len(types) + j))
                    # variables the rank gathers in its V->F calls: those named by its factors
                    named = np.zeros(self.n_vars, dtype=bool)
                    named[np.asarray(t.idx_v2f[lo:hi]).reshape(-1)] = True
                    out_bits[named] |= np.uint32(1 << q)
            self.raw_mask = torch.from_numpy(raw_bits.view(np.int32)).to(device)
            self.out_mask = torch.from_numpy(out_bits.view(np.int32)).to(device)
            self.link_fraction = (float(np.mean([bin(int(b)).count("1") for b in raw_bits[::97]])) / (world * len(types)),
                                  float(np.mean([bin(int(b)).count("1") for b in out_bits[::97]])) / world)

    # -- inputs ------------------------------------------------------------------------------
    def local_factor_features(self, x_f_full):
        """Slice full per-type factor features [1,F,C] (node-major) to this rank's shard."""
        return [x[:, lo:hi].contiguous() for x, (lo, hi) in zip(x_f_full, self.ranges)]

    def local_etypes(self, et_v2f_full, et_f2v_full):
        """Edge types of the local tables (layer-invariant: call once per forward)."""
        ev = [e[:, :, lo:hi].contiguous() for e, (lo, hi) in zip(et_v2f_full, self.ranges)]
        ef = [l.gather_etype(e) for l, e in zip(self.f2v, et_f2v_full)]
        return ev, ef

    # -- source-stationary plans of the local tables --------------------------------------------
    def source_plan(self, direction, j, T):
        """SourcePlan of this rank's V->F / compacted F->V table of type j, or None when the call is better off
        destination-stationary (mp_conv_v2's own rule: T = 16, at least 1.4 edges per source row, enough row-products
        saved).  Built once (the tables are static)."""
        from .mp_nn import SourcePlan, mp_conv_v2
        key = (direction, j)
        cache = self.__dict__.setdefault("_src_plans", {})
        if key in cache:
            return cache[key]
        plan = None
        if self.use_source_plans and T == 16:
            lo, hi = self.ranges[j]
            if direction == "v2f":
                idx, n_src, zero = self.idx_v2f[j], self.n_vars, None
            else:
                idx, n_src = self.idx_f2v[j], hi - lo
                zero = torch.from_numpy(self.f2v[j].slot_pad).to(self.device) if self.f2v[j].slot_pad.any() else None
            edges = int((idx >= 0).sum().item()) - (int(zero.sum().item()) if zero is not None else 0)
            if idx.numel() >= 50_000 and edges >= mp_conv_v2.AUTO_FAN_OUT[16] * n_src:
                cand = SourcePlan(idx, n_src, mask_negative=True, zero_slots=zero)
                if cand.n_rows * 1.25 <= idx.numel():
                    plan = cand
        cache[key] = plan
        return plan

    use_source_plans = True

    # -- peer-memory exchange ----------------------------------------------------------------
    def peer_buffers(self, J, O):
        """The two arena-backed variable-feature buffers [1,N,O] (layer l reads [l & 1], writes [(l + 1) & 1])."""
        if self.px is None:
            self.px = PeerExchange(self.n_vars, J, O, self.rank, self.world, self.device, self.group, ctas=self.exchange_ctas)
            self._comm_stream = torch.cuda.Stream(device=self.device)
        return self.px.xv

    def layer_peer(self, src, x_f_local, et_v2f_local, et_f2v_local, weights, out_f_local, kernel=_lib.KERNEL_AUTO,
                   workspaces=None, last=False):
        """One layer with the fused peer-memory exchange: reads the variable features from peer_buffers()[src],
        leaves the new ones in peer_buffers()[src ^ 1] on every rank -- valid after peer_wait() or the next call.

        Order on the main stream: the F->V partial maxima (they need only the factor features, so they do NOT wait
        for the previous layer's exchange) into raw buffer `src`; the exchange kernel is launched on a side stream;
        then, once the PREVIOUS layer's exchange has landed, the V->F calls.  The exchange of layer l therefore runs
        beside V->F of layer l and F->V of layer l+1, all of which leave it `ctas / 2` SMs (sm_limit)."""
        J = len(self.types)
        O = weights[0]["f2v"]["filters"].shape[1] // et_f2v_local[0].shape[1]
        xv = self.peer_buffers(J, O)
        px = self.px
        x_v = xv[src]
        nm = lambda t: t.permute(0, 2, 1).unsqueeze(-1)
        main = torch.cuda.current_stream(self.device)
        sms = self.n_sms - (px.ctas + 1) // 2               # two exchange CTAs share an SM
        raw = self.raw = px.raws[src]
        if self.raw_mask is None:
            raw.fill_(float("-inf"))      # with the masks nobody reads a (row, type) this rank's calls do not write
        for j in range(J):
            if self.f2v[j].n_rows > 0:
                w = weights[j]["f2v"]
                wsj = workspaces[j] if workspaces is not None else {}
                view = raw[:, :, j * O:(j + 1) * O]
                mp_forward(nm(x_f_local[j]), self.idx_f2v[j], et_f2v_local[j], w["filters"], None, None, None,
                           extension=0, aggregator=_lib.AGG_MAX, activation=_lib.ACT_NONE, kernel=kernel,
                           mask_negative=True, out=nm(view), tile_slots=self.tile_slots[j], out_rows=self.out_rows[j],
                           workspace=wsj.get("f2v"), filters_version=wsj.get("ver_f2v", 0),
                           sm_limit=sms if self.limit_f2v else 0,
                           plan=self.source_plan("f2v", j, et_f2v_local[j].shape[1]) if kernel != _lib.KERNEL_SIMT else None)
        bias, scale, shift = self._epilogue_params(weights)
        ready = torch.cuda.Event()
        ready.record(main)
        self._comm_stream.wait_event(ready)
        # intermediate layers: a rank receives only the rows its own V->F calls will gather; the last layer's
        # features go to everyone in full (they are the result)
        px.forward(src ^ 1, bias, scale, shift, _lib.ACT_RELU, 0.0, stream=self._comm_stream, raw_index=src,
                   raw_mask=self.raw_mask, out_mask=None if last else self.out_mask)
        done = torch.cuda.Event()
        done.record(self._comm_stream)
        self.peer_wait()                                    # x_v (the previous layer's exchange) must have landed
        self._peer_done = done
        for j in range(J):
            if x_f_local[j].shape[1] > 0:
                w = weights[j]["v2f"]
                wsj = workspaces[j] if workspaces is not None else {}
                mp_forward(nm(x_v), self.idx_v2f[j], et_v2f_local[j], w["filters"], w["bias"], w["scale"], w["shift"],
                           extension=0, aggregator=_lib.AGG_MAX, activation=_lib.ACT_RELU, kernel=kernel,
                           out=nm(out_f_local[j]), workspace=wsj.get("v2f"), filters_version=wsj.get("ver_v2f", 0),
                           sm_limit=sms,
                           plan=self.source_plan("v2f", j, et_v2f_local[j].shape[1]) if kernel != _lib.KERNEL_SIMT else None)
        return xv[src ^ 1]

    def _epilogue_params(self, weights):
        """Concatenated bias / BN scale / BN shift [J*O] of the F->V modules of one layer, cached per layer on the
        (address, in-place version) of every tensor involved."""
        key = tuple((w["f2v"][k].data_ptr(), w["f2v"][k]._version) for w in weights for k in ("bias", "scale", "shift"))
        cache = self.__dict__.setdefault("_epi_cache", {})
        ent = cache.get(key)
        if ent is None:
            if len(cache) > 256:
                cache.clear()
            ent = [torch.cat([w["f2v"][k] for w in weights]).contiguous() for k in ("bias", "scale", "shift")]
            cache[key] = ent
        return ent

    def peer_wait(self):
        """Make the current stream wait for the last launched exchange (its output buffer is then complete)."""
        done = getattr(self, "_peer_done", None)
        if done is not None:
            torch.cuda.current_stream(self.device).wait_event(done)
            self._peer_done = None

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
        bias, scale, shift = self._epilogue_params(weights)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.lib().fgnn_epilogue_sum_forward(raw.data_ptr(), out_v.data_ptr(), self.n_vars, O, J, bias.data_ptr(),
                                                        scale.data_ptr(), shift.data_ptr(), _lib.ACT_RELU, 0.0, 0, stream),
                   "epilogue_sum_forward")
        return out_v

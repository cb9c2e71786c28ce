"""Seeded synthetic factor-graph index tables in the reference's input layout (host side, numpy).

The hot path takes dense rectangular neighbour tables `nn_idx [M, K]` (int64) per factor type and
direction, exactly as the reference's table builders emit them:
  * Variable->Factor: `idx_v2f [F, K]`  -- row f lists the K variables of factor f;
  * Factor->Variable: `idx_f2v [N, Kv]` -- row v lists the (up to Kv) factors variable v is in;
    unused slots hold a VALID index (0) whose edge-type vector is all-zero, so the slot contributes a
    0-valued message that takes part in the max -- the reference's padding convention
    (lib/data/ldpc_dataset.py:36-37; last slot of generate_knn_table, train_syn_fixed_pw_hop.py:86-101).
`real_*` counts exclude padding (the message count of the metric).
"""
import numpy as np


class FactorType:
    """Tables of one factor type: F factors of order K over N variables."""

    def __init__(self, idx_v2f, idx_f2v, pad_f2v, name):
        self.idx_v2f = np.ascontiguousarray(idx_v2f, dtype=np.int64)      # [F, K]  values in [0, N)
        self.idx_f2v = np.ascontiguousarray(idx_f2v, dtype=np.int64)      # [N, Kv] values in [0, F)
        self.pad_f2v = np.ascontiguousarray(pad_f2v, dtype=bool)          # [N, Kv] True = padding slot
        self.name = name

    @property
    def n_factors(self):
        return self.idx_v2f.shape[0]

    @property
    def order(self):
        return self.idx_v2f.shape[1]

    @property
    def n_vars(self):
        return self.idx_f2v.shape[0]

    @property
    def kv(self):
        return self.idx_f2v.shape[1]

    @property
    def real_messages(self):
        """Real (destination, slot) evaluations per FGNN layer, both directions: 2 * F * K."""
        return 2 * self.n_factors * self.order


def _var_side_table(idx_v2f, n_vars, kv=None):
    """Transpose a [F,K] factor->variables table into the padded [N,Kv] variable->factors table."""
    F, K = idx_v2f.shape
    flat_v = idx_v2f.reshape(-1)
    flat_f = np.repeat(np.arange(F, dtype=np.int64), K)
    order = np.argsort(flat_v, kind="stable")
    v_sorted, f_sorted = flat_v[order], flat_f[order]
    deg = np.bincount(flat_v, minlength=n_vars)
    kmax = int(deg.max()) if deg.size else 0
    if kv is None:
        kv = max(kmax, 1)
    if kmax > kv:
        raise ValueError(f"a variable has {kmax} incident factors, table width is {kv}")
    start = np.concatenate([[0], np.cumsum(deg)[:-1]])
    slot = np.arange(v_sorted.size) - np.repeat(start, deg)
    idx = np.zeros((n_vars, kv), dtype=np.int64)
    pad = np.ones((n_vars, kv), dtype=bool)
    idx[v_sorted, slot] = f_sorted
    pad[v_sorted, slot] = False
    return idx, pad


def random_factor_type(n_vars, n_factors, order, rng, local_band=0, name="factors"):
    """`n_factors` factors of `order` distinct-ish variables each, built from random permutations
    so that variable degrees are as even as possible (n_factors*order / n_vars, +-1).
    local_band > 0: every factor's variables lie within (less than) two windows of that many consecutive
    variables (graphs with index locality: chains, kNN, LDPC-like)."""
    total = n_factors * order
    reps = -(-total // n_vars)
    if local_band > 0:
        # the same construction with BLOCK-LOCAL permutations: permutation r shuffles the variables inside windows of
        # `local_band` consecutive ids (windows shifted by half a band on alternate permutations, so neighbouring
        # windows are tied together); position p of every permutation then lies within one band of p, and a factor --
        # one position of `order` different permutations -- spans less than two bands.  Degrees stay as even as in
        # the uniform case (every variable appears once per permutation).
        def block_perm(r):
            shift = (r % 2) * (local_band // 2)
            ids = (np.arange(n_vars) + shift) % n_vars
            out = np.empty(n_vars, dtype=np.int64)
            for b0 in range(0, n_vars, local_band):
                blk = ids[b0:b0 + local_band]
                out[b0:b0 + local_band] = rng.permutation(blk)
            return out
        # column j of the table walks its own sequence of R block-local permutations; factor f sits at the same
        # relative position of every column's sequence, i.e. near variable (f * R * N // F) % N in all of them
        R = -(-n_factors // n_vars)
        step = max(1, (R * n_vars) // n_factors)
        pos = (np.arange(n_factors, dtype=np.int64) * (R * n_vars)) // n_factors
        cols = []
        for j in range(order):
            seq = np.concatenate([block_perm(j * R + r) for r in range(R)])
            cols.append(seq[(pos + j % step) % (R * n_vars)])
        idx_v2f = np.stack(cols, 1)
        idx_f2v, pad = _var_side_table(idx_v2f, n_vars)
        return FactorType(idx_v2f, idx_f2v, pad, name)
    else:
        pool = np.concatenate([rng.permutation(n_vars) for _ in range(reps)])[:total]
    # column-major fill: column j of the table is (a slice of) one permutation, so the
    # variables of one factor come from different permutations
    idx_v2f = pool.reshape(order, n_factors).T
    idx_f2v, pad = _var_side_table(idx_v2f, n_vars)
    return FactorType(idx_v2f, idx_f2v, pad, name)


def synthetic_map_graph(n_vars, n_pairwise, n_high, high_order, seed=0, local_band=0):
    """The synthetic MAP-inference graph of BASELINE.json configs[1]/[3]: `n_pairwise` order-2
    factors plus `n_high` factors of order `high_order` over `n_vars` variables."""
    rng = np.random.default_rng(seed)
    types = [random_factor_type(n_vars, n_pairwise, 2, rng, local_band, "pairwise")]
    if n_high > 0:
        types.append(random_factor_type(n_vars, n_high, high_order, rng, local_band,
                                        f"order{high_order}"))
    return types


def point_cloud_graph(n_points=65_536, knn=16, n_patches=8_192, patch_order=16, seed=0):
    """The point-cloud graph of BASELINE.json configs[4] (SURVEY 8d cfg 5): a seeded uniform 3-D point set, numbered
    along a Z-order (Morton) curve so that index locality follows spatial locality; `n_points * knn / 2` pairwise
    factors (point i, its j-th nearest neighbour), j = 1 .. knn/2 -- every point is the first variable of knn/2 factors
    and the second variable of as many as name it a neighbour, knn incidences per point on average; `n_patches` factors
    of order `patch_order`: the points nearest to a patch centre (centres = a seeded subset of the points).  The
    variable-side tables are padded to the largest membership with the reference's convention (valid index + all-zero
    edge type).  Needs scipy (cKDTree); host side, seconds at the default size."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    pts = rng.random((n_points, 3))
    # Morton order: interleave the top 10 bits of every coordinate
    q = np.minimum((pts * 1024).astype(np.uint64), 1023)

    def spread(v):
        v = (v | (v << 16)) & np.uint64(0x030000FF)
        v = (v | (v << 8)) & np.uint64(0x0300F00F)
        v = (v | (v << 4)) & np.uint64(0x030C30C3)
        v = (v | (v << 2)) & np.uint64(0x09249249)
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    pts = pts[np.argsort(code, kind="stable")]
    tree = cKDTree(pts)
    half = knn // 2
    _, nb = tree.query(pts, k=half + 1)                                # column 0 is the point itself
    first = np.repeat(np.arange(n_points, dtype=np.int64), half)
    pair = np.stack([first, nb[:, 1:].reshape(-1).astype(np.int64)], 1)                 # [N*knn/2, 2]
    idx_f2v, pad = _var_side_table(pair, n_points)
    types = [FactorType(pair, idx_f2v, pad, "knn-pairwise")]
    if n_patches > 0:
        centres = pts[rng.choice(n_points, n_patches, replace=False)]
        _, members = tree.query(centres, k=patch_order)
        patch = np.sort(members.astype(np.int64), 1)
        idx_f2v, pad = _var_side_table(patch, n_points)
        types.append(FactorType(patch, idx_f2v, pad, f"patch{patch_order}"))
    return types


def locality_order(types):
    """Renumber the factors of every type by their smallest variable (stable), so that a contiguous range of factors
    touches a (nearly) contiguous range of variables when the graph has index locality (banded / chain / kNN
    graphs): the precondition for halo-sized exchanges under contiguous sharding.  The variable
    numbering is kept.  Returns new FactorType objects (tables only; features / edge types generated afterwards
    follow the new numbering)."""
    out = []
    for t in types:
        out.append(FactorType(*_locality_order_native(t), t.pad_f2v, t.name))
    return out


def _locality_order_native(t):
    """fgnn_locality_order_host (csrc/plan.cu): stable sort of the factors by smallest variable + renumbering."""
    import ctypes
    from . import _lib
    F, K = t.idx_v2f.shape
    N, Kv = t.idx_f2v.shape
    order = np.empty(F, dtype=np.int64)
    v2f = np.empty((F, K), dtype=np.int64)
    f2v = np.empty((N, Kv), dtype=np.int64)
    pad = np.ascontiguousarray(t.pad_f2v, dtype=np.uint8)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    _lib.check(_lib.lib().fgnn_locality_order_host(p(t.idx_v2f), F, K, p(t.idx_f2v), p(pad), N, Kv, p(order), p(v2f), p(f2v)),
               "locality_order_host")
    return v2f, f2v


def locality_order_numpy(types):
    """The same renumbering with numpy (cross-check of the native builder)."""
    out = []
    for t in types:
        order = np.argsort(t.idx_v2f.min(1), kind="stable")             # new factor i = old factor order[i]
        inv = np.empty_like(order)
        inv[order] = np.arange(order.size)
        idx_f2v = np.where(t.pad_f2v, 0, inv[t.idx_f2v])                # pads keep pointing at a valid row (0)
        out.append(FactorType(t.idx_v2f[order], idx_f2v, t.pad_f2v, t.name))
    return out


def chain_knn_table(n, k):
    """Chain-MRF neighbour table with the semantics of the reference's generate_knn_table
    (train_syn_fixed_pw_hop.py:86-101): node i lists its k//2 left neighbours i-k//2..i-1 and its
    k//2 - 1 right neighbours i+1..i+k//2-1, clamped to [0, n-1]; the edge feature is the signed
    distance i - j after clamping; the last slot is never filled and stays (index 0, feature 0) --
    the reference's pad.  Returns (nn_idx [1,n,k] int64, efeature [1,1,n,k] float32)."""
    hk = k // 2
    offs = np.concatenate([np.arange(-hk, 0), np.arange(1, hk)]).astype(np.int64)   # k-1 offsets
    i = np.arange(n, dtype=np.int64)[:, None]
    j = np.clip(i + offs[None, :], 0, n - 1)
    idx = np.zeros((1, n, k), dtype=np.int64)
    ef = np.zeros((1, 1, n, k), dtype=np.float32)
    idx[0, :, :offs.size] = j
    ef[0, 0, :, :offs.size] = (i - j).astype(np.float32)
    return idx, ef


def parse_alist(text):
    """Parse a MacKay `alist` parity-check description (the format of
    ldpc_codes/96.3.963/96.3.963, read by lib/data/ldpc_dataset.py:26-49).  Returns
    (var_to_checks [N, max_col_w], check_to_vars [M, max_row_w]) 0-based, -1 = unused slot."""
    tok = text.split()
    pos = 0

    def take(n):
        nonlocal pos
        vals = [int(t) for t in tok[pos:pos + n]]
        pos += n
        return vals
    n, m = take(2)
    cw, rw = take(2)
    col_w = take(n)
    row_w = take(m)
    v2c = -np.ones((n, cw), dtype=np.int64)
    c2v = -np.ones((m, rw), dtype=np.int64)
    for i in range(n):
        vals = take(cw)
        for j in range(col_w[i]):
            v2c[i, j] = vals[j] - 1
    for i in range(m):
        vals = take(rw)
        for j in range(row_w[i]):
            c2v[i, j] = vals[j] - 1
    return v2c, c2v

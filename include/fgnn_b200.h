/* fgnn_b200 -- C ABI of the B200-native FGNN message-passing hot path.
 *
 * This is the drop-in boundary for ONE path of zzhang1987/Factor-Graph-Neural-Network: the
 * Variable->Factor / Factor->Variable message-passing layer `mp_conv_v2.forward`
 * (reference lib/model/mpnn/mp_nn.py:115-175) with its gather (mp_nn.py:92-113), aggregators
 * (mp_nn.py:73-87) and bias / BatchNorm / activation epilogue (mp_nn.py:165-173).
 *
 * The reference has no FFI of its own on this path (it is Python over ATen), so the entry points
 * below are what a ctypes binding inside the reference's `mp_conv_v2.forward` calls; the binding
 * is shown in INTEGRATION.md and implemented in factor-graph-neural-network_b200/mp_nn.py.
 *
 * Conventions: plain C, no torch types; every pointer in fgnn_mp_args is a DEVICE pointer unless
 * the function name ends in `_host`; functions never allocate device memory, never synchronise the
 * device (except the `_host` and `_check_` helpers, which say so) and never throw: they return 0 or
 * a negative fgnn_status.  All launches go to the `stream` argument (a cudaStream_t).  Re-entrant.
 */
#ifndef FGNN_B200_H_
#define FGNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGNN_B200_VERSION 200 /* 0.2.0 */

typedef enum fgnn_status {
  FGNN_OK = 0,
  FGNN_ERR_INVALID_ARG = -1,    /* NULL pointer, non-positive dimension, unknown enum value       */
  FGNN_ERR_INDEX_RANGE = -2,    /* nn_idx entry outside [0,N) (reference: ATen gather error)      */
  FGNN_ERR_SHAPE = -3,          /* extension mode with M != N (reference requires M == N,
                                   mp_nn.py:136-159), or filters rows != C / 2C                    */
  FGNN_ERR_UNSUPPORTED = -4,    /* shape/dtype the selected kernel cannot run                      */
  FGNN_ERR_WORKSPACE = -5,      /* workspace too small (see fgnn_mp_workspace_bytes)               */
  FGNN_ERR_CUDA = -6,           /* a CUDA runtime call failed; fgnn_last_cuda_error() has the code */
  FGNN_ERR_NO_DEVICE = -7       /* no sm_100 device                                                */
} fgnn_status;

/* mp_conv_type, reference mp_nn.py:7-10 (same integer values; also base_model.py:6-8) */
typedef enum fgnn_extension {
  FGNN_NO_EXTENSION = 0,
  FGNN_ORIG_WITH_NEIGHBOR = 1,
  FGNN_ORIG_WITH_DIFF = 2
} fgnn_extension;

/* aggregtor, reference mp_nn.py:68-90 */
typedef enum fgnn_aggregator {
  FGNN_AGG_MAX = 0,      /* agg_max, mp_nn.py:73-75                              */
  FGNN_AGG_SOFTMAX = 1,  /* 1/gamma * logsumexp(gamma * x), mp_nn.py:80-83       */
  FGNN_AGG_MEAN = 2,     /* mp_nn.py:87                                          */
  FGNN_AGG_NONE = 3      /* aggregtor=None: output keeps the K axis, mp_nn.py:162 */
} fgnn_aggregator;

typedef enum fgnn_activation {
  FGNN_ACT_NONE = 0,
  FGNN_ACT_RELU = 1,     /* mp_nn.py:63-64 */
  FGNN_ACT_LEAKY_RELU = 2 /* slope in act_slope; used when the module was given nn.LeakyReLU */
} fgnn_activation;

typedef enum fgnn_dtype { FGNN_F32 = 0, FGNN_BF16 = 1 } fgnn_dtype;
typedef enum fgnn_index_dtype { FGNN_I64 = 0, FGNN_I32 = 1 } fgnn_index_dtype;

typedef enum fgnn_kernel {
  FGNN_KERNEL_AUTO = 0,    /* tensor-core kernel when the shape qualifies, else SIMT */
  FGNN_KERNEL_SIMT = 1,    /* fp32 CUDA-core kernel, every shape                    */
  FGNN_KERNEL_TCGEN05 = 2  /* tcgen05/TMEM kernel; FGNN_ERR_UNSUPPORTED if the shape does not qualify */
} fgnn_kernel;

enum {
  /* a negative nn_idx entry marks an EMPTY slot that is excluded from the aggregate (shard-local
     F2V tables); a destination with no live slot yields -inf (max/softmax) or 0 (mean).
     Without this flag negative entries are out of range, as in the reference. */
  FGNN_FLAG_MASK_NEGATIVE = 1u,
  /* out += result instead of out = result: fuses the caller's `nfeature = nfeature + nv`
     (FactorNN, factor_mpnn_sp.py:147,151) into the store.  Not valid with FGNN_AGG_NONE. */
  FGNN_FLAG_ACCUMULATE = 2u,
  /* source-stationary calls only: never aggregate inside the first pass (see src_edge_slot); for tests / comparisons */
  FGNN_FLAG_NO_FUSED_REDUCE = 4u
};

/* One message-passing call: out[b,o,m] = act(BN(bias[o] + AGG_k sum_t etype[b,t,m,k] *
 *                                     sum_c xin(b,m,k)[c] * filters[c, o*T+t]))
 * with xin = x[b,:,idx[b,m,k]] (NO_EXTENSION), [x[b,:,m] || x[b,:,idx]] (NEIGHBOR) or
 * [x[b,:,m] || x[b,:,m]-x[b,:,idx]] (DIFF).  Strides are in ELEMENTS. */
typedef struct fgnn_mp_args {
  const void* x;          /* logical [B,C,N]: element (b,c,n) at x[b*x_sb + c*x_sc + n*x_sn]            */
  const void* idx;        /* logical [B,M,K] int64|int32: (b,m,k) at idx[b*idx_sb + m*K + k]; idx_sb may be 0 */
  const void* etype;      /* logical [B,T,M,K]: (b,t,m,k) at etype[b*et_sb + (t*M + m)*K + k]; et_sb may be 0   */
  const float* filters;   /* [C or 2C, O*T] fp32, row-major, column o*T+t (mp_nn.py:41-46)              */
  const float* bias;      /* [O] or NULL (mp_nn.py:51-55)                                               */
  const float* bn_scale;  /* [O] or NULL: gamma / sqrt(running_var + eps)  (eval BatchNorm folded)      */
  const float* bn_shift;  /* [O] or NULL: beta - running_mean * bn_scale                                */
  void* out;              /* logical [B,O,M,Kout]: (b,o,m,k) at out[b*out_sb + o*out_so + m*out_sm + k*out_sk] */
  void* workspace;        /* device scratch, >= fgnn_mp_workspace_bytes(args), 256-byte aligned, or NULL */
  size_t workspace_bytes;
  int64_t x_sb, x_sc, x_sn;
  int64_t idx_sb, et_sb;
  int64_t out_sb, out_so, out_sm, out_sk;
  int32_t B, N, M, K, C, O, T;
  int32_t extension;      /* fgnn_extension   */
  int32_t aggregator;     /* fgnn_aggregator  */
  int32_t activation;     /* fgnn_activation  */
  int32_t dtype;          /* fgnn_dtype of x / etype / out (filters, bias, bn always fp32) */
  int32_t idx_dtype;      /* fgnn_index_dtype */
  int32_t kernel;         /* fgnn_kernel      */
  uint32_t flags;
  float gamma;            /* softmax aggregator temperature (reference: 3)  */
  float act_slope;        /* LeakyReLU negative slope                      */
  int64_t filters_version;/* any value that changes when `filters` changes (the tensor-core path caches
                             its bf16 weight image in the workspace keyed on workspace + filters pointer +
                             this value; the caller then must leave the workspace untouched between
                             calls); 0 = never cache */
  /* Compacted shard-local tables (factor-sharded F->V direction); both optional (NULL):     */
  const int32_t* tile_slots; /* [ceil(B*M/128)], each in [1,K]: in the 128-row destination tile i only slots
                                k < tile_slots[i] are evaluated (rows sorted by live-slot count, the rest
                                of the row is empty anyway)                                              */
  const int32_t* out_rows;   /* [B*M]: destination row g = b*M+m is written to out + out_rows[g]*out_sm
                                (+ o*out_so) instead of its own position; negative = not written       */
  int32_t sm_limit;          /* > 0: the persistent tensor-core kernel uses at most this many SMs, leaving the
                                rest to a concurrent collective (the kernel owns whole SMs: a CTA that cannot
                                be placed would serialise behind the collective).  0 = all SMs              */
  int32_t reserved_;
  /* Source-stationary evaluation (optional; src_ptr == NULL = destination-stationary).  The reference computes
     H = x W once per SOURCE node and gathers rows of H per slot (mp_nn.py:124-134); the default kernel instead
     recomputes x[n] W once per slot.  With a plan of the index table (built by the caller from nn_idx, layout
     below) the call computes H once per source row on the tensor cores, stores one O-wide message per edge and
     aggregates every destination's messages in a second streaming pass -- bit-identical results, fewer
     row-products when sources feed several slots.  fp32, NO_EXTENSION, C = 64, T in {4,8,16}, O*T = 256 or a
     multiple of 512, O % 8 == 0.
     An EDGE is a live slot (b,m,k).  A VIRTUAL ROW is a source row with at most src_row_cap of its edges: virtual
     row v < B*N is source row v with its first src_row_cap edges, the rows beyond (src_rows) carry the remaining
     edges of rows that have more (the reference pads with a valid index and a zero edge type, so the pad target
     collects every padded slot).  Edges are numbered virtual row by virtual row. */
  const int32_t* src_ptr;    /* [n_src_rows + 1]: the edges of virtual row v are src_ptr[v] .. src_ptr[v+1]-1 (<= src_row_cap) */
  const int32_t* slot_edge;  /* [B*M*K]: edge number of slot (b*M + m)*K + k; -1 = empty slot (not aggregated);
                                -2 = a slot the caller KNOWS to carry an all-zero edge-type vector (the reference's
                                padding: valid index + zero edge type): it is no edge of the plan, its message is the
                                constant 0 and it takes part in the aggregate like any other slot                */
  const void* etype_edges;   /* [E, T]: edge-type vector of every edge, edge-major (fgnn_src_permute_etype)      */
  void* messages;            /* [E, O] scratch for the per-edge messages, 32-byte aligned                       */
  int64_t n_edges;           /* E                                                                                */
  const int32_t* src_rows;   /* [n_src_rows - B*N]: flattened source row b*N + n of virtual row v >= B*N (may be NULL
                                when n_src_rows == B*N)                                                          */
  int64_t n_src_rows;        /* virtual rows, >= B*N                                                             */
  int32_t src_row_cap;       /* 3: tables with few edges per row; 6: the epilogue splits a row's edges over two warps
                                (tables whose rows mostly have 4-6 edges)                                        */
  int32_t src_rows_per_batch; /* > 0: BATCH-LOCAL plan for fused aggregation (below): the virtual rows are numbered batch element
                                by batch element, this many each (<= 128, row cap 3), and src_rows names the source row of
                                EVERY virtual row (n_src_rows = B * src_rows_per_batch).  0: the layout described above.    */
  const int32_t* src_edge_slot; /* [E]: slot (b*M + m)*K + k of every edge.  Fused aggregation: a table [B,M,K] is block-diagonal
                                over the batch, so with a batch-local plan, all filter columns in one CTA (O*T = 256) and
                                every slot live, a tile = one batch element keeps its messages in shared memory and aggregates
                                them inside the first pass: the per-codeword graphs of LDPC decoding (train_ldpc.py) never
                                send a message through HBM and need no second launch.                                        */
} fgnn_mp_args;

int fgnn_version(void);
const char* fgnn_strerror(int status);
int fgnn_last_cuda_error(void);          /* cudaError_t of the last failed runtime call on this thread */

/* Scratch the call needs (0 for the SIMT kernel with node-major x). */
size_t fgnn_mp_workspace_bytes(const fgnn_mp_args* args);

/* Which kernel FGNN_KERNEL_AUTO resolves to for these args (FGNN_KERNEL_SIMT / _TCGEN05), or <0. */
int fgnn_mp_select_kernel(const fgnn_mp_args* args);

/* The hot path.  Replaces mp_conv_v2.forward (mp_nn.py:115-175).  Asynchronous on `stream`. */
int fgnn_mp_forward(const fgnn_mp_args* args, void* stream);

/* Same call with HOST buffers (x, idx, etype, out in fgnn_mp_args are host pointers; filters/bias/bn
 * host too).  Allocates device buffers, copies in, runs, copies out, frees, synchronises.  This is
 * the end-to-end entry a non-PyTorch caller binds. */
int fgnn_mp_forward_host(const fgnn_mp_args* host_args);

/* 1 if `args` (with its source-stationary plan fields set) qualifies for the source-stationary path, else 0. */
int fgnn_mp_src_supported(const fgnn_mp_args* args);

/* etype [B,T,M,K] (reference layout, batch stride et_sb elements) -> the plan's edge-type image: edge-major [E,T] in
 * the plan's edge order (edge e = slot (b*M + m)*K + k = edge_slot[e]), with the T/4 16-byte pieces of edge e stored
 * at piece position j ^ key(e), key(e) = (e / (8 / (T/4))) & (T/4 - 1) -- a tile's block is staged in shared memory
 * by one bulk copy and read row-per-thread; the key spreads those reads over the banks.  Asynchronous on `stream`. */
int fgnn_src_permute_etype(const float* etype, int64_t et_sb, const int32_t* edge_slot, float* out, int32_t T,
                           int32_t M, int32_t K, int64_t n_edges, void* stream);

/* The scripts' edge model as ONE kernel: etype = Conv1x1(H -> T)(ReLU(Conv1x1(Fe -> H)(efeature))) with H = 64
 * (reference train_ldpc.py:32-38,68-69; train_syn_hop_factor.py:174-179): the H-channel hidden tensor never reaches
 * HBM.  efeature logical [B,Fe,M,K] (batch stride ef_sb elements, the rest contiguous), w1 [H,Fe], b1 [H] or NULL,
 * w2 [T,H], b2 [T] or NULL.  edge_slot == NULL: out = etype [B,T,M,K] (batch stride out_sb).  edge_slot != NULL
 * ([n_edges] slot of every edge of a source-stationary plan): out = the plan's edge-type image [n_edges, T], i.e.
 * what fgnn_src_permute_etype would produce from the etype -- that pass is fused too.  Fe <= 8, T <= 16. */
int fgnn_emodel_forward(const float* efeature, int64_t ef_sb, const float* w1, const float* b1, const float* w2,
                        const float* b2, float* out, int64_t out_sb, const int32_t* edge_slot, int64_t n_edges,
                        int32_t B, int32_t Fe, int32_t H, int32_t T, int32_t M, int32_t K, void* stream);

/* ---- backward of the call (training; reference train_ldpc.py:222-231 relies on ATen autograd) ---------------------
 * With g = dL/dy, a = dAGG/de (one-hot at the first arg-max slot | softmax_k(gamma e) | 1/K), ge = g * a:
 *     Z[(b,m,k), o*T+t] = ge[.,o] * etype[b,t,m,k]     dXin = Z W^T     dW = Xin^T Z     H = Xin W
 *     d etype[b,t,m,k] = sum_o ge[.,o] * H[., o*T+t]    dx = scatter-add of dXin through nn_idx (+ the self part)
 * The three dense products are plain GEMMs run by the caller (mp_nn.py: torch.matmul = cuBLAS); these entry points
 * are the graph-structured pieces, each over the destination rows m0 .. m0+mc-1 of every batch element (slot order
 * (b, m - m0, k)), so the O*T-wide intermediates stay bounded.  `a` describes the forward call (x, idx, etype, strides,
 * shapes, extension, aggregator, gamma, flags); fp32 only.  All asynchronous on `stream`. */
int fgnn_bwd_gather(const fgnn_mp_args* a, float* xin /* [B*mc*K, C or 2C] */, int32_t m0, int32_t mc, void* stream);
int fgnn_bwd_slot_values(const fgnn_mp_args* a, const float* H /* [B*mc*K, O*T] */, float* e /* [B*mc*K, O] */, int32_t m0,
                         int32_t mc, void* stream);
int fgnn_bwd_aggregate(const fgnn_mp_args* a, const float* e, const float* grad_out, int64_t g_sb, int64_t g_so, int64_t g_sm,
                       int64_t g_sk, float* ge /* [B*mc*K, O] */, int32_t m0, int32_t mc, void* stream);
/* d_etype (reference layout [B,T,M,K], batch stride det_sb; may be NULL) from H, then H <- Z in place */
int fgnn_bwd_outer(const fgnn_mp_args* a, float* H_inout, const float* ge, float* d_etype, int64_t det_sb, int32_t m0, int32_t mc,
                   void* stream);
/* dx [B,N,C] node-major, zero-initialised by the caller, += dXin [B*mc*K, C or 2C] (atomic adds) */
int fgnn_bwd_scatter(const fgnn_mp_args* a, const float* dxin, float* dx, int32_t m0, int32_t mc, void* stream);

/* ---- native graph preprocessing (host) ---------------------------------------------------------
 * The source-stationary plan of a HOST table nn_idx [B,M,K] (batch stride idx_sb elements) in O(E) counting-sort
 * passes, stable in slot order: call once with the output pointers NULL to get V (virtual rows, the return value) and
 * E (*n_edges_out), then with src_ptr [V+1], slot_edge [B*M*K], edge_slot [E], src_rows [V - B*n_src] (may be NULL when
 * V == B*n_src).  Entries outside [0, n_src) are empty slots.  Layout as documented at fgnn_mp_args.src_ptr. */
int64_t fgnn_plan_build_host(const void* idx, int32_t idx_dtype, int32_t B, int32_t M, int32_t K, int64_t idx_sb, int32_t n_src,
                             int32_t row_cap, int64_t* n_edges_out, int32_t* src_ptr, int32_t* slot_edge, int32_t* edge_slot,
                             int32_t* src_rows);

/* Factors renumbered by their smallest variable (stable): order [F] (new factor i = old factor order[i]), the
 * factor-side table with its rows permuted, the variable-side table [N,Kv] with its entries renumbered (pad slots --
 * pad[i] != 0 -- keep the reference's valid-index pad 0).  With index locality in the graph, contiguous factor shards
 * then touch (nearly) contiguous variable ranges: the precondition for halo-sized exchanges (fgnn_halo_pull). */
int fgnn_locality_order_host(const int64_t* idx_v2f, int64_t F, int32_t K, const int64_t* idx_f2v, const uint8_t* pad, int64_t N,
                             int32_t Kv, int64_t* order, int64_t* idx_v2f_out, int64_t* idx_f2v_out);

/* Index validation the reference gets for free from ATen's gather (mp_nn.py:111): returns
 * FGNN_ERR_INDEX_RANGE if any entry of idx[count] is outside [lo, N).  Synchronises `stream`.
 * `scratch` = 8 bytes of device memory. */
int fgnn_check_index_range(const void* idx, int idx_dtype, int64_t count, int64_t lo, int64_t n,
                           void* scratch, void* stream);

/* The same check without the round trip: only launches the scan on `stream`; on a violation the kernel
 * stores 1 to *flag (never cleared here).  `flag` may be device memory or pinned (mapped) host memory the
 * caller polls later -- the reference's own CUDA behaviour is an asynchronous device-side assert. */
int fgnn_check_index_range_async(const void* idx, int idx_dtype, int64_t count, int64_t lo, int64_t n,
                                 int32_t* flag, void* stream);

/* Elementwise epilogue out = act(bn(in + bias)) on a node-major [rows, O] buffer, used after the
 * cross-GPU max-all-reduce of the raw aggregate (the epilogue is non-linear, so it runs after the
 * reduce; reference mp_nn.py:165-173).  In-place allowed.  -inf inputs (no live slot on any shard)
 * stay -inf before bias is added. */
int fgnn_epilogue_forward(const float* in, float* out, int64_t rows, int32_t O, const float* bias,
                          const float* bn_scale, const float* bn_shift, int32_t activation,
                          float act_slope, void* stream);

/* The same for J factor types at once: in = [rows, J*O] (type j in columns j*O..j*O+O-1, the layout of
 * the single max-all-reduce per layer), bias / bn_scale / bn_shift = [J*O] (or NULL),
 * out[r,o] = sum_j act(bn_j(in[r, j*O+o] + bias_j[o]))  -- FactorNN's `nfeature += nv` over the factor
 * types (factor_mpnn_sp.py:142-147).  With accumulate != 0 the sum is added to out. */
int fgnn_epilogue_sum_forward(const float* in, float* out, int64_t rows, int32_t O, int32_t J, const float* bias,
                              const float* bn_scale, const float* bn_shift, int32_t activation, float act_slope,
                              int32_t accumulate, void* stream);

/* ---- factor-sharded layer: the cross-GPU step as one kernel over NVLink peer memory -------------
 * Every rank owns an ARENA allocated by this library (cudaMalloc, zero-filled) whose IPC handle the host code
 * passes to the other ranks (one process per GPU); `fgnn_comm_open` maps a peer's arena.  The caller lays the
 * arena out: 2*8 uint32 epoch flags + one uint64 block counter, the raw aggregate [rows, J*O] and the next-layer
 * feature buffers [rows, O] (factor-graph-neural-network_b200/parallel.py: PeerExchange). */
int fgnn_comm_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64);
int fgnn_comm_open(const unsigned char* handle64, void** dev_ptr);
int fgnn_comm_close(void* dev_ptr);
int fgnn_comm_free(void* dev_ptr);

typedef struct fgnn_exchange_args {
  const float* raw[8];   /* per rank: raw per-type maxima [rows, J*O] of that rank's factor shard (-inf = none)        */
  float* out[8];         /* per rank: next-layer variable features [rows, O]; this rank writes rows row0..row1-1 of ALL */
  uint32_t* flags[8];    /* per rank: 16 uint32 epoch flags (A[8] then B[8]) in that rank's arena                        */
  uint64_t* counter;     /* this rank's block counter (in its own arena); every launch must use the same `ctas`          */
  const float* bias;     /* [J*O] or NULL; bn_scale / bn_shift likewise (both or neither)                                 */
  const float* bn_scale;
  const float* bn_shift;
  int64_t rows, row0, row1;
  int32_t world, rank, J, O;
  int32_t activation;    /* fgnn_activation */
  float act_slope;
  uint32_t epoch;        /* 0: the kernel numbers its launches itself from `counter` (launches captured in a CUDA graph
                            keep counting on replay); else > 0, the same on every rank, growing by one per call            */
  int32_t ctas;          /* grid of 512-thread CTAs, two per SM (<= 128; 0 = 64): small, so the kernel is co-resident beside
                            sm_limit-ed launches                                                                          */
  /* Static sparsity of the shards (both optional, device, [rows] uint32): a rank's factors touch only some variables,
     so its raw row is -inf elsewhere and it never gathers the others' features.                                       */
  const uint32_t* raw_mask; /* bit q*J+j set: rank q's raw row holds data of type j (unset: never read, taken as -inf) */
  const uint32_t* out_mask; /* bit q set: rank q needs the finished row (unset: not written to it; own rank always)   */
} fgnn_exchange_args;

/* out_q[n, o] = sum_j act(bn_j(max_r raw_r[n, j*O+o] + bias_j[o])) for n in [row0,row1), written into every rank's
 * `out`; returns when launched.  When the kernel completes on a rank, all rows of its own `out` are in place and no
 * peer still reads its `raw` (epoch flags, system-scope release/acquire).  All ranks launch it with the same epoch. */
int fgnn_exchange_forward(const fgnn_exchange_args* args, void* stream);

/* ---- owner-computes sharding: the feature-halo pull (halo-restricted exchange) ---------------------------
 * Every rank owns a contiguous range of variables and of each type's factors and evaluates all slots of its own
 * destinations; the source rows other ranks own are its halo.  One launch copies the halo rows of up to four buffers
 * (variables + factor types) straight out of their owners' arenas over NVLink; flags / counter / epoch / ctas as in
 * fgnn_exchange_args (all ranks launch it the same number of times).  Rows are raw bytes: fp32 and bf16 alike. */
typedef struct fgnn_halo_job {
  const void* src[8];        /* per rank: base of that rank's buffer (its owned rows first)          */
  void* dst;                 /* this rank's buffer                                                    */
  const int32_t* src_row;    /* [n] row of halo row i in its owner's buffer                           */
  const uint8_t* src_rank;   /* [n] owner of halo row i                                               */
  int64_t dst_row0;          /* halo row i lands in row dst_row0 + i of dst                           */
  int32_t n;                 /* halo rows                                                             */
  int32_t row_bytes;         /* bytes per row, a multiple of 16                                       */
} fgnn_halo_job;

typedef struct fgnn_halo_args {
  fgnn_halo_job jobs[4];
  uint32_t* flags[8];
  uint64_t* counter;
  int32_t n_jobs, world, rank;
  uint32_t epoch;
  int32_t ctas;              /* 512-thread CTAs (<= 128; 0 = 32)                                       */
  int32_t reserved_;
} fgnn_halo_args;

int fgnn_halo_pull(const fgnn_halo_args* args, void* stream);

/* InstanceNorm2d (affine = false, biased variance, statistics of instance (b,c) over its N nodes) + activation on a
 * logical [B,C,N] tensor with element strides: the v2v / f2f maps of FactorNN's layer body (reference
 * base_model.py:83-90, iid_mapping_in) after their 1x1 convolution.  out may alias x. */
int fgnn_instance_norm_forward(const float* x, float* out, int32_t B, int32_t C, int32_t N, int64_t x_sb, int64_t x_sc,
                               int64_t x_sn, int64_t out_sb, int64_t out_sc, int64_t out_sn, float eps,
                               int32_t activation, float act_slope, void* stream);

/* The same norm when the N nodes of an instance are split over ranks (a sharded FactorNN layer's f2f / v2v maps;
 * SURVEY 8f rank 4): fgnn_instance_norm_partial reduces THIS rank's rows to sums [B,C] -- of x when mean == NULL, of
 * (x - mean)^2 when the (all-reduced) mean [B,C] is given: the two-pass variance of the single-GPU kernel -- the caller
 * all-reduces the 2 x C floats per instance (torch.distributed), and fgnn_instance_norm_apply normalises the local rows
 * with the global mean / inverse standard deviation and applies the activation. */
int fgnn_instance_norm_partial(const float* x, const float* mean, float* sums, int32_t B, int32_t C, int32_t N, int64_t x_sb,
                               int64_t x_sc, int64_t x_sn, void* stream);
int fgnn_instance_norm_apply(const float* x, float* out, const float* mean, const float* inv_std, int32_t B, int32_t C, int32_t N,
                             int64_t x_sb, int64_t x_sc, int64_t x_sn, int64_t out_sb, int64_t out_sc, int64_t out_sn,
                             int32_t activation, float act_slope, void* stream);

/* Layout helper: channel-major [B,C,N] -> node-major [B,N,C] (the reference's
 * x.permute(0,2,3,1).contiguous(), mp_nn.py:125). */
int fgnn_to_node_major(const float* x, float* out, int32_t B, int32_t C, int32_t N, int64_t x_sb,
                       int64_t x_sc, int64_t x_sn, void* stream);

/* Programmatic dependent launch between consecutive tensor-core launches on a stream (default on):
 * launch i+1 may start its set-up, filter load and index / edge-type prefetch while launch i drains;
 * it orders itself behind launch i before its first read of x and its first store to out, so results
 * do not change.  Returns the previous setting.  Process-wide; not a per-call option. */
int fgnn_set_programmatic_launch(int enabled);

/* Number of kernels this library has launched since load (all threads). */
uint64_t fgnn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FGNN_B200_H_ */

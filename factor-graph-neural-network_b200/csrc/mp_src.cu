// Source-stationary evaluation of the FGNN message-passing call on tcgen05 (NO_EXTENSION, C = 64, fp32).
//
//   out[b,o,m] = act(BN(bias[o] + AGG_k sum_t etype[b,t,m,k] * (x[b, idx[b,m,k], :] . W[:, o*T+t])))
//   reference: lib/model/mpnn/mp_nn.py:115-175
//
// The reference computes H = x W once per SOURCE node (mp_nn.py:124-127) and gathers the O*T-wide rows
// of H per slot (mp_nn.py:92-113, 128-134); the destination-stationary kernel (mp_tc.cu) never
// materialises H but recomputes x[n] W once per slot -- E row-products instead of N (6x for the
// variable->factor call of a pairwise type whose variables each sit in six factors).  At T = 16 that
// kernel runs at the tensor pipe's sustained rate, so the only way forward is fewer row-products:
//
//   pass 1 (mp_src_kernel):  a tile = 128 consecutive SOURCE rows.  H = x W for the tile is accumulated in
//            TMEM exactly as in mp_tc.cu (split-bf16, three MMA terms, stationary filter slice); each
//            epilogue thread owns one source row and walks that row's OUT-EDGES (src_ptr, source-sorted
//            edge order): it contracts the H chunk with the edge's edge-type vector and stores the
//            O-wide MESSAGE of the edge -- H itself (O*T wide) still never leaves the SM.
//   pass 2 (mp_reduce_kernel): every destination aggregates the messages of its K slots (slot_edge maps a
//            slot to its edge), then bias / eval-BN / activation -- pure streaming.
//
// Messages cost 2 * 4*O bytes per edge of extra HBM traffic; in exchange the tensor work drops from E to N
// row-products.  The caller chooses (fgnn_mp_args.src_ptr != NULL); the results are bit-identical to the
// destination-stationary kernel (same MMA terms per row, same contraction order, same aggregation order).
//
// CTA roles as in mp_tc.cu (512 threads): warps 0-7 epilogue, 8-11 converters, 12-13 gatherers (the
// "gather" is the identity here: contiguous rows), 14 MMA.  The two epilogue groups split a row's EDGES
// (group g takes edges g, g+2, ...), not the accumulator columns: every thread reads whole chunks.
#include "tc_common.cuh"

namespace fgnn {

namespace {

using namespace tc;

constexpr int kEB = 3;                 // edges per epilogue thread per pass over an accumulator chunk

struct SrcParams {
  const float* x;                      // [rows, 64] source rows, node-major
  const int32_t* src_ptr;              // [rows + 1]
  const float* et_edges;               // [E, T]
  float* msg;                          // [E, O]
  uint32_t rows;                       // B * N
  int O, T;
};

template <int T, int NCH>
__global__ void __launch_bounds__(tc::kThreads, 1)
mp_src_kernel(const SrcParams p, const uint8_t* __restrict__ wimg, const int S, const int n_workers, const int n_tiles) {
  constexpr int NC = 128;                      // accumulator columns per chunk
  constexpr int COLS = NC * NCH;               // columns of W this CTA owns
  constexpr int CPC = NC / T;                  // output channels per chunk
  constexpr int CPH = 64 / T;                  // output channels per 64-column half chunk
  constexpr int NST = 2;                       // raw-ring stages: one item per tile, the epilogue paces the CTA
  constexpr int ROWB = row_bytes(false), STAGEB = stage_bytes(false);
  static_assert(16 % T == 0 && T >= 4 && CPH % 4 == 0, "unsupported edge-type count");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                        // [2 parts][COLS rows][128 B]  UMMA K-major SW128
  uint8_t* sA = sB + w_bytes(COLS, false);                   // [NST][128 rows][256 B]      raw ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + NST * STAGEB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  const uint32_t bar0 = smem_u32(bars);
  auto raw_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto raw_empty = [&](uint32_t s) { return bar0 + 8u * (kMaxAStages + s); };
  auto ta_full = [&](uint32_t s) { return bar0 + 8u * (2 * kMaxAStages + s); };
  auto ta_empty = [&](uint32_t s) { return bar0 + 8u * (2 * kMaxAStages + kTA + s); };
  auto t_full = [&](uint32_t s) { return bar0 + 8u * (2 * kMaxAStages + 2 * kTA + s); };
  auto t_empty = [&](uint32_t s) { return bar0 + 8u * (2 * kMaxAStages + 2 * kTA + kAcc + s); };
  const uint32_t w_full = bar0 + 8u * (2 * kMaxAStages + 2 * kTA + 2 * kAcc);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x % S, worker = blockIdx.x / S;
  const int col0 = split * COLS;                             // first W column of this CTA
  const int ch0 = col0 / T;                                  // first output channel of this CTA

  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < kMaxAStages; ++s) {
      mbar_init(raw_full(s), kGatherWarps * 32);
      mbar_init(raw_empty(s), 128);
    }
    for (int s = 0; s < kTA; ++s) {
      mbar_init(ta_full(s), 128);
      mbar_init(ta_empty(s), 1);
    }
    for (int s = 0; s < kAcc; ++s) {
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), kEpiWarps * 32);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // =====================================================================================
    // EPILOGUE: thread (eg, r) owns source row tile*128 + r == TMEM lane r and that row's out-edges
    // e0 + eg, e0 + eg + 2, ...; per accumulator chunk it reads the whole chunk (two 64-column halves),
    // contracts it with each of its edges' edge-type vectors and stores CPC message channels per edge
    // =====================================================================================
    reg_inc<kRegEpi>();
    const int eg = warp >> 2, wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
    // edge range of this thread's row in a tile (0 edges past the end)
    auto edge_range = [&](int tile, int32_t& e0, int32_t& e1) {
      const uint32_t g = (uint32_t)tile * kTileM + r;
      e0 = e1 = 0;
      if (tile < n_tiles && g < p.rows) { e0 = __ldg(p.src_ptr + g); e1 = __ldg(p.src_ptr + g + 1); }
    };
    int32_t e0n, e1n;
    edge_range(worker, e0n, e1n);
    uint32_t ct = 0;
    bool waited = false;
    float* const msg_base = p.msg + ch0;
    for (int tile = worker; tile < n_tiles; tile += n_workers) {
      const int32_t e0 = e0n, e1 = e1n;
      edge_range(tile + n_workers, e0n, e1n);                // next tile's range: in flight during this tile
      const int n_mine = e1 - e0 > eg ? (e1 - e0 - eg + 1) >> 1 : 0;       // edges e0 + eg + 2i, i < n_mine
      const int n_warp = __reduce_max_sync(0xffffffffu, n_mine);
      const int n_pass = (n_warp + kEB - 1) / kEB;
      // edge-type vectors of the first pass stay in registers across the chunks of the tile
      float et[kEB][T];
      auto load_et = [&](int pass) {
#pragma unroll
        for (int i = 0; i < kEB; ++i) {
          const int ii = pass * kEB + i;
          if (ii < n_mine) {
            const float4* pe = reinterpret_cast<const float4*>(p.et_edges + (int64_t)(e0 + eg + 2 * ii) * T);
#pragma unroll
            for (int t4 = 0; t4 < T / 4; ++t4) {
              const float4 v = __ldg(pe + t4);
              et[i][4 * t4] = v.x; et[i][4 * t4 + 1] = v.y; et[i][4 * t4 + 2] = v.z; et[i][4 * t4 + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int t = 0; t < T; ++t) et[i][t] = 0.f;
          }
        }
      };
      load_et(0);
      if (!waited) { pdl_wait(); waited = true; }            // first store: the preceding launch may still read msg
#pragma unroll 1
      for (int chunk = 0; chunk < NCH; ++chunk) {
        const uint32_t st = ct % kAcc;
        mbar_wait(t_full(st), (ct / kAcc) & 1);
        tc_fence_after();
        const uint32_t taddr = lane_addr + st * kAccCols;
        for (int pass = 0; pass < (n_pass > 0 ? n_pass : 1); ++pass) {
          if (pass > 0 || (n_pass > 1 && chunk > 0)) load_et(pass);      // rows with many edges: reload per pass
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t d[4][16];
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) tmem_ld16(taddr + h * 64 + gq * 16, d[gq]);
            tmem_ld_wait();
            if (h == 1 && pass + 1 >= n_pass) {              // last read of this accumulator stage
              tc_fence_before();
              mbar_arrive(t_empty(st));
            }
#pragma unroll
            for (int i = 0; i < kEB; ++i) {
              const int ii = pass * kEB + i;
              if (ii < n_warp) {                             // warp-uniform
                float o[CPH];
#pragma unroll
                for (int c = 0; c < CPH; ++c) {              // channel c of this half: columns c*T .. c*T+T-1
                  o[c] = contract_types<T>(et[i], &d[(c * T) >> 4][(c * T) & 15]);      // same function as mp_tc.cu: bit-identical
                }
                if (ii < n_mine) {
                  float4* dst = reinterpret_cast<float4*>(msg_base + (int64_t)(e0 + eg + 2 * ii) * p.O + chunk * CPC + h * CPH);
#pragma unroll
                  for (int c4 = 0; c4 < CPH; c4 += 4) dst[c4 >> 2] = make_float4(o[c4], o[c4 + 1], o[c4 + 2], o[c4 + 3]);
                }
              }
            }
          }
        }
        ++ct;
      }
      if (e1n > e0n + eg) {                                  // next tile's first edge-type rows towards L2
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.et_edges + (int64_t)(e0n + eg) * T));
      }
    }
  } else if (warp < kGatherWarp0) {
    // =====================================================================================
    // CONVERTERS: raw fp32 row (ring) -> bf16 hi/lo pairs -> A stage in tensor memory (as mp_tc.cu)
    // =====================================================================================
    reg_dec<kRegConv>();
    const int cr = tid - kConvWarp0 * 32;
    const uint32_t row_u = smem_u32(sA) + (uint32_t)cr * ROWB;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kTACol0;
    uint32_t i = 0;
    for (int tile = worker; tile < n_tiles; tile += n_workers, ++i) {
      const uint32_t st = i % NST, use = i / NST;
      const uint32_t ta = i % kTA, tuse = i / kTA;
      mbar_wait(raw_full(st), use & 1);
      float4 v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = lds_f4(row_u + st * STAGEB + (uint32_t)((c ^ (cr & 15)) * 16));
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[c].x, v[c].y), h1 = __floats2bfloat162_rn(v[c].z, v[c].w);
        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        hi[2 * c] = pack_bf16(h0);
        hi[2 * c + 1] = pack_bf16(h1);
        lo[2 * c] = pack_bf16(__floats2bfloat162_rn(v[c].x - f0.x, v[c].y - f0.y));
        lo[2 * c + 1] = pack_bf16(__floats2bfloat162_rn(v[c].z - f1.x, v[c].w - f1.y));
      }
      mbar_arrive(raw_empty(st));
      mbar_wait(ta_empty(ta), (tuse & 1) ^ 1);
      tc_fence_after();
      tmem_st32(lane_addr + ta * kTACols, hi);
      tmem_st32(lane_addr + ta * kTACols + 32, lo);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ta_full(ta));
    }
  } else {
  reg_dec<kRegAux>();                                        // warps 12-15 together (one warpgroup)
  if (warp < kMmaWarp) {
    // =====================================================================================
    // "GATHERERS": the tile's 128 source rows are consecutive; cp.async them into the raw ring in the
    // converter's swizzled layout (chunk q of row rr at position q ^ (rr & 15))
    // =====================================================================================
    constexpr int LPR = ROWB / 16, RPW = 32 / LPR, ROWS_W = kTileM / kGatherWarps, NIT = ROWS_W / RPW, NDO = LPR / RPW;
    const int pw = warp - kGatherWarp0;
    const int sub = lane / LPR, q = lane % LPR;
    const uint8_t* xq = reinterpret_cast<const uint8_t*>(p.x) + q * 16;
    const uint32_t sA_u = smem_u32(sA);
    uint32_t dst_off[NDO];
#pragma unroll
    for (int c = 0; c < NDO; ++c)
      dst_off[c] = (uint32_t)(pw * ROWS_W + RPW * c + sub) * ROWB + (uint32_t)((q ^ ((RPW * c + sub) & (LPR - 1))) * 16);
    pdl_wait();                                              // x is the preceding launch's output: order behind it
    uint32_t i = 0;
    for (int tile = worker; tile < n_tiles; tile += n_workers, ++i) {
      const uint32_t st = i % NST, use = i / NST;
      mbar_wait(raw_empty(st), (use & 1) ^ 1);
      const uint32_t stage = sA_u + st * STAGEB;
      const uint32_t row0 = (uint32_t)tile * kTileM + pw * ROWS_W + sub;      // this lane's row in instruction 0
#pragma unroll
      for (int u = 0; u < NIT; ++u) {
        const uint32_t row = row0 + RPW * u;
        const bool ok = row < p.rows;
        const uint32_t dst = stage + dst_off[u % NDO] + (uint32_t)((u / NDO) * LPR * ROWB);
        cp_async16(dst, xq + (uint64_t)(ok ? row : 0u) * ROWB, ok ? 16u : 0u);
      }
      cp_async_arrive_noinc(raw_full(st));
    }
  } else if (warp == kMmaWarp) {
    // =====================================================================================
    // MMA ISSUER (as mp_tc.cu, one item per tile)
    // =====================================================================================
    const uint32_t sB_u = smem_u32(sB);
    if (elect_one()) {
      const int OT = p.O * p.T;
      constexpr uint32_t part = COLS * 128, piece = part < 32768u ? part : 32768u;
      mbar_expect_tx(w_full, 2 * part);
      for (int h = 0; h < 2; ++h) {
        const uint8_t* src = wimg + kHeaderBytes + (size_t)h * OT * 128 + (size_t)col0 * 128;
        for (uint32_t o = 0; o < part; o += piece) bulk_g2s(sB_u + h * part + o, src + o, piece, w_full);
      }
    }
    __syncwarp();
    mbar_wait(w_full, 0);
    constexpr uint32_t idesc = umma_idesc(NC);
    const uint64_t desc_hi = umma_desc_sw128(0);
    uint32_t ct = 0, i = 0;
    for (int tile = worker; tile < n_tiles; tile += n_workers, ++i) {
      const uint32_t ta = i % kTA, tuse = i / kTA;
      mbar_wait(ta_full(ta), tuse & 1);
      const uint32_t a_hi = tmem_base + kTACol0 + ta * kTACols, a_lo = a_hi + 32;
#pragma unroll
      for (int chunk = 0; chunk < NCH; ++chunk) {
        const uint32_t ts = ct % kAcc;
        mbar_wait(t_empty(ts), ((ct / kAcc) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ts * kAccCols;
        const uint32_t b_hi = (sB_u + (uint32_t)chunk * NC * 128u) >> 4, b_lo = b_hi + ((uint32_t)COLS * 128u >> 4);
        if (elect_one()) {
#pragma unroll
          for (int term = 0; term < 3; ++term) {             // xl*Wh + xh*Wl + xh*Wh
            const uint32_t a = term == 0 ? a_lo : a_hi;
            const uint32_t bb = term == 1 ? b_lo : b_hi;
#pragma unroll
            for (int ks = 0; ks < kC / 16; ++ks)
              umma_bf16_ts(d_tmem, a + ks * 8, desc_hi | (uint64_t)(bb + ks * 2), idesc, (term | ks) != 0);
          }
          umma_commit(t_full(ts));
          if (chunk == NCH - 1) umma_commit(ta_empty(ta));
        }
        __syncwarp();
        ++ct;
      }
    }
  }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// pass 2: every destination aggregates the messages of its K slots, then bias / BN / activation.
// One thread per (destination row, 4 channels); slots in order k = 0..K-1 like the other kernels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mp_reduce_kernel(const MpParams p, const float* __restrict__ msg, const int32_t* __restrict__ slot_edge) {
  const int O4 = p.O >> 2;
  const int64_t rows = (int64_t)p.B * p.M, total = rows * O4;
  const float neg = p.act == FGNN_ACT_NONE ? 1.f : (p.act == FGNN_ACT_RELU ? 0.f : p.slope);
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = i / O4;
    const int o = (int)(i - g * O4) * 4;
    const int32_t* se = slot_edge + g * p.K;
    const int kt = p.tile_k ? p.tile_k[g >> 7] : p.K;
    float a[4], s[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { a[c] = p.agg == FGNN_AGG_MEAN ? 0.f : -INFINITY; s[c] = 0.f; }
    float live = 0.f;
    for (int k = 0; k < kt; ++k) {
      const int32_t e = __ldg(se + k);
      if (e < 0) continue;
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(msg + (int64_t)e * p.O + o));
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      live += 1.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (p.agg == FGNN_AGG_MAX) {
          a[c] = fmaxf(a[c], v[c]);
        } else if (p.agg == FGNN_AGG_SOFTMAX) {
          const float z = p.gamma * v[c], mx = fmaxf(a[c], z);
          s[c] = s[c] * expf(a[c] - mx) + expf(z - mx);
          a[c] = mx;
        } else {
          a[c] += v[c];
        }
      }
    }
    float y[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float r;
      if (p.agg == FGNN_AGG_MAX) r = a[c];
      else if (p.agg == FGNN_AGG_SOFTMAX) r = live > 0.f ? (logf(s[c]) + a[c]) * (1.f / p.gamma) : -INFINITY;
      else r = a[c] * (live > 0.f ? 1.f / live : 0.f);
      const float bi = p.bias ? p.bias[o + c] : 0.f, sc = p.scale ? p.scale[o + c] : 1.f, sh = p.scale ? p.shift[o + c] : 0.f;
      float v = fmaf(r + bi, sc, sh);
      v = v >= 0.f ? v : v * neg;
      y[c] = r == -INFINITY ? r : v;
    }
    const int64_t orow = p.out_rows ? (int64_t)p.out_rows[g] : g;
    if (orow < 0) continue;
    float4* dst = reinterpret_cast<float4*>(p.out + orow * p.o_sm + o);
    if (p.accumulate) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]) : "memory");
    } else {
      *dst = make_float4(y[0], y[1], y[2], y[3]);
    }
  }
}

// etype [B,T,M,K] (reference layout) -> edge-major [E,T] in source-sorted edge order
__global__ void et_permute_kernel(const float* __restrict__ et, int64_t et_sb, const int32_t* __restrict__ edge_slot,
                                  float* __restrict__ out, int T, int64_t MK, int64_t n_edges) {
  const int64_t total = n_edges * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / T;
    const int t = (int)(i - e * T);
    const int64_t slot = edge_slot[e];                       // (b*M + m)*K + k
    const int64_t b = slot / MK, mk = slot - b * MK;
    out[i] = et[b * et_sb + (int64_t)t * MK + mk];
  }
}

template <int T, int NCH>
int launch_src(const SrcParams& sp, const uint8_t* wimg, int S, int workers, int tiles, cudaStream_t st) {
  auto kern = mp_src_kernel<T, NCH>;
  const size_t smem = 1024 + (size_t)tc::w_bytes(128 * NCH, false) + 2 * (size_t)tc::stage_bytes(false) + tc::kNumBars * 8 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return FGNN_ERR_CUDA;
    if (fa.numRegs != tc::kRegLaunch) return FGNN_ERR_UNSUPPORTED;      // setmaxnreg budget (tc_common.cuh)
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return FGNN_ERR_CUDA;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(S * workers));
  cfg.blockDim = dim3(tc::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc_pdl_enabled() ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, sp, wimg, S, workers, tiles);
  count_launch();
  return e == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

}  // namespace

bool src_supported(const fgnn_mp_args* a) {
  if (!a->src_ptr || !a->slot_edge || !a->etype_edges || !a->messages) return false;
  if (a->extension != FGNN_NO_EXTENSION || a->dtype != FGNN_F32 || a->C != tc::kC) return false;
  if (a->aggregator == FGNN_AGG_NONE) return false;
  if (a->T != 16 && a->T != 8 && a->T != 4) return false;
  if ((a->O * a->T) % 256 || a->O % 4 || a->O > 128) return false;
  if (a->x_sc != 1 || a->x_sn != a->C) return false;                              // node-major rows
  if (a->B > 1 && a->x_sb != (int64_t)a->N * a->C) return false;                  // batch-contiguous
  if ((int64_t)a->B * a->N * tc::row_bytes(false) >= (int64_t)UINT32_MAX) return false;
  if (a->n_edges <= 0 || a->n_edges >= INT32_MAX) return false;
  if ((reinterpret_cast<uintptr_t>(a->x) & 15) || (reinterpret_cast<uintptr_t>(a->out) & 15) ||
      (reinterpret_cast<uintptr_t>(a->messages) & 15) || (reinterpret_cast<uintptr_t>(a->etype_edges) & 15))
    return false;
  if (a->out_so != 1 || (a->out_sm & 3)) return false;
  if (a->B > 1 && a->out_sb != (int64_t)a->M * a->out_sm) return false;
  return true;
}

int launch_mp_src(const MpParams& p, const fgnn_mp_args* a, cudaStream_t stream) {
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255)) return FGNN_ERR_WORKSPACE;
  const int OT = p.O * p.T;
  const int wrc = tc_prepare_weights(p.W, ws, tc::kC, OT, a->filters_version, stream);
  if (wrc != FGNN_OK) return wrc;
  SrcParams sp;
  sp.x = p.x; sp.src_ptr = a->src_ptr; sp.et_edges = reinterpret_cast<const float*>(a->etype_edges);
  sp.msg = reinterpret_cast<float*>(a->messages);
  sp.rows = (uint32_t)((int64_t)p.B * p.N);
  sp.O = p.O; sp.T = p.T;
  // columns per CTA: 512 (the split-bf16 image of 512 columns is 128 KB), or all of them when fewer
  const int cols = OT < 512 ? OT : 512;
  const int NCH = cols / 128, S = OT / cols;
  int sms = tc_num_sms();
  if (a->sm_limit > 0 && a->sm_limit < sms) sms = a->sm_limit;
  if (S > sms) return FGNN_ERR_UNSUPPORTED;
  const int tiles = (int)((sp.rows + tc::kTileM - 1) / tc::kTileM);
  int workers = sms / S;
  if (workers > tiles) workers = tiles;
  int rc = FGNN_ERR_UNSUPPORTED;
  if (p.T == 16 && NCH == 4) rc = launch_src<16, 4>(sp, ws, S, workers, tiles, stream);
  else if (p.T == 16 && NCH == 2) rc = launch_src<16, 2>(sp, ws, S, workers, tiles, stream);
  else if (p.T == 8 && NCH == 4) rc = launch_src<8, 4>(sp, ws, S, workers, tiles, stream);
  else if (p.T == 8 && NCH == 2) rc = launch_src<8, 2>(sp, ws, S, workers, tiles, stream);
  else if (p.T == 4 && NCH == 4) rc = launch_src<4, 4>(sp, ws, S, workers, tiles, stream);
  else if (p.T == 4 && NCH == 2) rc = launch_src<4, 2>(sp, ws, S, workers, tiles, stream);
  if (rc != FGNN_OK) return rc;
  // pass 2
  const int64_t total = (int64_t)p.B * p.M * (p.O / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, mp_reduce_kernel, p, (const float*)sp.msg, a->slot_edge);
  count_launch();
  return e == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

int launch_et_permute(const float* et, int64_t et_sb, const int32_t* edge_slot, float* out, int T, int64_t MK,
                      int64_t n_edges, cudaStream_t stream) {
  const int64_t total = n_edges * T;
  if (total == 0) return FGNN_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  et_permute_kernel<<<(unsigned)blocks, 256, 0, stream>>>(et, et_sb, edge_slot, out, T, MK, n_edges);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

}  // namespace fgnn

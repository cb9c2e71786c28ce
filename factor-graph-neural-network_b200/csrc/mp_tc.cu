// tcgen05 / TMEM kernel for the FGNN message-passing call (NO_EXTENSION, C = 64; fp32 or bf16 I/O).
//
//   out[b,o,m] = act(BN(bias[o] + AGG_k sum_t etype[b,t,m,k] * (x[b, idx[b,m,k], :] . W[:, o*T+t])))
//   reference: lib/model/mpnn/mp_nn.py:115-175 (the VF and the FV module of FGNN)
//
// Formulation.  A destination tile = 128 consecutive (b,m) rows.  For every slot k the 128 source
// rows x[idx[.,k]] form the A operand [128 x C]; the filters form the B operand [C x O*T]; one
// UMMA M=128 accumulates H_k = A_k W into TMEM (fp32), NC columns at a time.  TMEM lane r == row r,
// so an epilogue thread of row r reads its own H_k row, contracts it with its slot's edge-type
// vector (sum_t et[t] * H[o*T+t]) and folds the result into a running max / logsumexp / mean held
// in registers: the K-reduction needs no shuffles and no shared memory, and nothing O*T wide ever
// reaches HBM (the reference materialises H, an int64 index expansion and the gathered rows).
//
// fp32 parity on bf16 tensor cores (fp32 I/O): both operands are split x = xh + xl, W = Wh + Wl
// (bf16 each) and three MMAs xl*Wh + xh*Wl + xh*Wh accumulate in fp32 (error ~2^-16 relative,
// inside the 1e-4 contract; plain TF32/bf16 is not -- SURVEY 7 hard part 2).
// bf16 I/O (SURVEY 8d cfg 4: bf16 features / edge types / weights, fp32 accumulate): x rows are
// the A operand as they are, W is rounded once to bf16, ONE MMA term; the filter image of all
// O*T columns then fits one CTA at T = 16 (no column split across CTAs).
//
// Shared-memory bandwidth is the scarce resource (measured: with both operands in shared memory the
// MMA reads 12 KB per 128-cycle instruction and runs at half rate while gather/convert starve), so
// the A operand lives in TENSOR MEMORY: the converter threads write the (split-)bf16 rows straight
// into TMEM with tcgen05.st (thread r == lane r == row r) and the MMA takes A from TMEM; shared
// memory only carries the raw gather ring and the stationary filter slice.
//
// CTA roles (512 threads = 4 warpgroups, 1 CTA / SM, persistent over tiles); an ITEM is one (tile, k):
//   warps 0-7   epilogue    two groups of four warps; warps w and w+4 share TMEM lanes 32(w%4).. and
//                           split every accumulator chunk's columns in halves (group 0: low half):
//                           TMEM -> registers, edge-type contraction, aggregate; bias/BN/act and the
//                           store go through a shared staging tile (both groups write their channels,
//                           then share the line-sized stores); the next item's edge-type vector is
//                           prefetched during the current one
//   warps 8-11  converters  thread r: raw row r from the ring (conflict-free swizzled chunks) ->
//                           bf16 hi/lo pairs (fp32 I/O) or as is (bf16 I/O) -> tcgen05.st into the A
//                           stage in TMEM
//   warps 12-13 gatherers   cp.async (LDGSTS) 16-byte chunks of the 128 source rows (64 per warp) into the
//                           raw ring: no register staging, every free ring stage's gather is in flight;
//                           the item's indices are loaded one item ahead
//   warp  14    MMA         one elected lane issues tcgen05.mma (A: TMEM, B: smem descriptor); owns the
//                           TMEM allocation; brings the stationary filter slice in with TMA bulk copies
//   warp  15    idle        (setmaxnreg is a warpgroup-wide instruction: warps 12-15 release registers together)
// TMEM (512 columns): 3 accumulator stages x 128 columns + 2 A stages x 64 columns (32 hi + 32 lo).
// Pipelines (all mbarrier): raw ring raw_empty -> raw_full; A stages ta_empty -> ta_full;
// accumulators t_empty -> t_full.
// The filters are stationary: each CTA keeps the bf16 image of its column slice (O*T / S columns,
// S = column split across CTAs so the slice fits in shared memory) for its whole lifetime; the
// image is produced once per weight version by w_split_kernel.
//
// Programmatic dependent launch: the kernel releases its dependents at once
// (griddepcontrol.launch_dependents) and orders itself behind its predecessor only where it must
// (griddepcontrol.wait before the first read of x and before the first store to out), so set-up,
// filter load and index / edge-type prefetch of launch i+1 overlap the tail of launch i.
#include <mutex>
#include <unordered_map>

#define FGNN_TC_TRACE_TU          // this translation unit owns the trace buffer of -DFGNN_TC_TRACE builds
#include "tc_common.cuh"

namespace fgnn {


// ---------------------------------------------------------------------------------------------
// filters [C, O*T] fp32 -> split-bf16 image in the UMMA B layout (K-major rows of 64 bf16 = 128 B,
// 16-byte chunks XOR-swizzled with row % 8): image[part][n][c], n = column o*T+t; part 0 = bf16(W)
// (all the bf16-I/O kernel uses), part 1 = bf16(W - part 0).
// ---------------------------------------------------------------------------------------------
// diff != 0 (ORIG_WITH_DIFF, C = 2 x 64 filter rows): the image holds [W_top + W_bot ; -W_bot], so that
// [x_i || x_j] . image == [x_i || x_i - x_j] . W  (mp_nn.py:136-159 evaluated as x_i (W_top + W_bot) - x_j W_bot).
__global__ void w_split_kernel(const float* __restrict__ W, uint8_t* __restrict__ ws, int C, int OT, int64_t version, int diff, int T) {
  tc::Header* h = reinterpret_cast<tc::Header*>(ws);
  if (version != 0 && h->version == version && h->filters == W && h->C == C && h->OT == OT) return;
  uint8_t* img = ws + tc::kHeaderBytes;
  const int KA = C / tc::kC;                                 // K atoms of 64 channels: image[part][ka][n][64 bf16]
  const int total = OT * (C / 8);                            // one 16-byte chunk (8 channels) per thread-iteration
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % OT, chunk = i / OT;                    // consecutive threads: consecutive columns (coalesced)
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = W[(int64_t)(chunk * 8 + 2 * j) * OT + n], b = W[(int64_t)(chunk * 8 + 2 * j + 1) * OT + n];
      if (diff) {
        const int c0 = chunk * 8 + 2 * j, half = C / 2;
        if (c0 < half) { a += W[(int64_t)(c0 + half) * OT + n]; b += W[(int64_t)(c0 + 1 + half) * OT + n]; }
        else { a = -a; b = -b; }
      }
      const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
      const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah));
      const __nv_bfloat16 bl = __float2bfloat16_rn(b - __bfloat162float(bh));
      hi[j] = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
      lo[j] = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
    }
    const int ka = chunk >> 3, cc = chunk & 7;
    const int ni = tc::image_column(n, T);                   // T = 4: channel pairs interleaved (tc_common.cuh)
    const size_t off = (size_t)ka * OT * 128 + (size_t)ni * 128 + (size_t)((cc ^ (ni & 7)) * 16);
    *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + (size_t)KA * OT * 128 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {                 // read by later launches only (stream order)
    h->version = version; h->filters = W; h->C = C; h->OT = OT;
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel
//   T   edge types (columns per output channel)        NC  accumulator columns per MMA chunk
//   NCH chunks per CTA (CTA column slice = NC*NCH)     AGG aggregator      XB  bf16 x / etype / out
//   KA  K atoms: input channels / 64 (1, or 2 for the C = 128 layers of the LDPC model, fp32 I/O only).  With two
//       atoms an A stage takes 128 TMEM columns, so there are two accumulator stages instead of three.
// ---------------------------------------------------------------------------------------------
template <int T, int NC, int NCH, int AGG, bool XB, int KA>
__global__ void __launch_bounds__(tc::kThreads, 1)
mp_tc_kernel(const MpParams p, const uint8_t* __restrict__ wimg, const int S, const int n_workers,
             const int n_tiles) {
  using namespace tc;
  constexpr int COLS = NC * NCH;               // columns of W this CTA owns
  constexpr int CH = COLS / T;                 // output channels this CTA owns
  constexpr int CH_PER_LD = 16 / T;            // channels per 16-column TMEM load
  constexpr int NLD = NC / (16 * kEpiGroups);  // 16-column loads per chunk per epilogue group
  constexpr int CPG = NC / (T * kEpiGroups);   // channels per chunk per epilogue group
  constexpr int CHG = CPG * NCH;               // channels (accumulator registers) per epilogue thread
  constexpr int NST = a_stages(COLS, CH, XB, KA);   // raw-ring stages that fit beside the filter slice and the output tile
  constexpr int ROWB = row_bytes(XB, KA), STAGEB = stage_bytes(XB, KA);
  constexpr int ACC = KA == 2 ? 2 : kAcc;      // accumulator stages
  constexpr int TACOL0 = ACC * kAccCols, TACOLS = KA * kTACols;      // TMEM: ACC x 128 accumulator + 2 x KA*64 A-stage columns
  static_assert(TACOL0 + kTA * TACOLS <= 512 && (KA == 1 || !XB), "tensor memory budget");
  constexpr int OB = XB ? 2 : 4;               // bytes per output element
  constexpr int NTERMS = XB ? 1 : 3;           // MMA terms per K step
  static_assert(NC % (16 * kEpiGroups) == 0 && NC <= kAccCols && 16 % T == 0 && CH <= 64 && NST >= 2 && CPG % 4 == 0 &&
                    (CH * OB) % 32 == 0, "unsupported shape");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1 KB alignment by OFFSET (not by integer round-trip) so the compiler keeps the shared state space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                        // [1|2 parts][COLS rows][128 B]  UMMA K-major SW128
  uint8_t* sA = sB + w_bytes(COLS, XB, KA);                  // [NST][128 rows][ROWB]          raw ring
  uint8_t* sOut = sA + NST * STAGEB;                         // [4 quarters][32 rows][CH*OB]   output staging (swizzled chunks)
  float* s_epi = reinterpret_cast<float*>(sOut + out_tile_bytes(CH, XB));   // [3][64]: bias, BN scale, BN shift (16-byte aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_epi + 3 * 64);
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
  const uint32_t rows_total = (uint32_t)p.B * (uint32_t)p.M;    // < 2^31 (tc_supported)
  const uint32_t Mu = (uint32_t)p.M;
  // (b, m) of flattened destination row g; B == 1 is the common (single graph) case
  auto split_row = [&](uint32_t g, uint32_t& b, uint32_t& m) {
    if (p.B == 1) { b = 0; m = g; } else { b = g / Mu; m = g - b * Mu; }
  };
  // slots evaluated in a tile: K, or the tile's own count for compacted shard-local tables
  auto slots_of = [&](int tile) -> int { return p.tile_k ? (tile < n_tiles ? p.tile_k[tile] : 0) : p.K; };
  // this CTA's items (tile, k) in order: tile = worker + j*n_workers, k < slots_of(tile)
  struct ItemIter {
    int tile, k, kt;
  };
  auto first_item = [&]() -> ItemIter { return ItemIter{worker, 0, slots_of(worker)}; };
  auto next_item = [&](ItemIter& it) {
    if (++it.k >= it.kt) { it.k = 0; it.tile += n_workers; it.kt = slots_of(it.tile); }
  };

  // ---- one-time setup ---------------------------------------------------------------------
  pdl_launch_dependents();                                   // the next launch may begin its own set-up now
  if (tid == 0) {
    for (int s = 0; s < kMaxAStages; ++s) {
      mbar_init(raw_full(s), kGatherWarps * 32);             // gather threads (cp.async arrive-on)
      mbar_init(raw_empty(s), 128);                          // converter threads
    }
    for (int s = 0; s < kTA; ++s) {
      mbar_init(ta_full(s), 128);                            // converter threads
      mbar_init(ta_empty(s), 1);                             // tcgen05.commit
    }
    for (int s = 0; s < kAcc; ++s) {
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), kEpiWarps * 32);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), 512);
  if (tid < CH) {                                             // absent bias / BN: exact identities (+0, *1, +0)
    s_epi[tid] = p.bias ? p.bias[ch0 + tid] : 0.f;
    s_epi[CH + tid] = p.scale ? p.scale[ch0 + tid] : 1.f;
    s_epi[2 * CH + tid] = p.scale ? p.shift[ch0 + tid] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // =====================================================================================
    // EPILOGUE: warps w and w+4 own destination rows tile*128 + 32*(w%4) .. +31 == TMEM lanes
    // 32*(w%4) .. +31; group eg = w/4 takes the columns [eg*NC/2, (eg+1)*NC/2) of every chunk
    // =====================================================================================
    reg_inc<kRegEpi>();
    const int eg = warp >> 2, wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(eg * (NC / kEpiGroups));
    const int64_t et_st = (int64_t)p.M * p.K;                // stride between edge types
    // edge-type vector + liveness of item (tile j, slot k) for this thread's row
    float et_nx[T];
    bool live_nx = false;
    auto fetch = [&](int tile, int k) {
      const uint32_t g = (uint32_t)tile * kTileM + r;
      const bool ok = tile < n_tiles && g < rows_total;
      uint32_t b = 0, m = 0;
      if (ok) split_row(g, b, m);
      const int64_t e0 = (int64_t)b * p.et_sb + (int64_t)m * p.K + k;
      if (XB) {
        const unsigned short* pe = reinterpret_cast<const unsigned short*>(p.et) + e0;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          et_nx[t] = ok ? __uint_as_float((uint32_t)__ldg(pe) << 16) : 0.f;
          pe += et_st;
        }
      } else {
        const float* pe = p.et + e0;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          et_nx[t] = ok ? __ldg(pe) : 0.f;
          pe += et_st;
        }
      }
      live_nx = ok;
      if (ok && p.mask_neg) live_nx = load_index(p.idx, p.idx64, (int64_t)b * p.idx_sb + (int64_t)m * p.K + k) >= 0;
    };
    fetch(worker, 0);
    uint32_t ct = 0;
    bool waited = false;
#ifdef FGNN_TC_TRACE
    uint32_t ei = 0;
#endif
    for (int tile = worker; tile < n_tiles; tile += n_workers) {
      float acc[CHG];                                        // max | running max of gamma*e | sum
      float acc2[AGG == FGNN_AGG_SOFTMAX ? CHG : 1];         // softmax: running sum of exp
      float live_count = 0.f;
#pragma unroll
      for (int c = 0; c < CHG; ++c) acc[c] = (AGG == FGNN_AGG_MEAN) ? 0.f : -INFINITY;
#pragma unroll
      for (int c = 0; c < (AGG == FGNN_AGG_SOFTMAX ? CHG : 1); ++c) acc2[c] = 0.f;
      const int kt = slots_of(tile);
      // output row of tile row (wq*32 + lane), fetched now so the finish phase never waits on it
      int32_t my_orow = -1;
      {
        const uint32_t go = (uint32_t)tile * kTileM + r;
        if (go < rows_total) my_orow = p.out_rows ? p.out_rows[go] : (int32_t)go;
      }
      for (int k = 0; k < kt; ++k) {
        float et[T];
#pragma unroll
        for (int t = 0; t < T; ++t) et[t] = et_nx[t];
        uint64_t et2[4];
        if constexpr (T == 4) {
#pragma unroll
          for (int t = 0; t < 4; ++t) et2[t] = pack2(et[t], et[t]);
        }
        const bool live = live_nx;
        // prefetch the next item's edge types: their latency hides behind this item's math
        if (k + 1 < kt) fetch(tile, k + 1); else fetch(tile + n_workers, 0);
#pragma unroll
        for (int chunk = 0; chunk < NCH; ++chunk) {
          const uint32_t st = ct % ACC;
          mbar_wait(t_full(st), (ct / ACC) & 1);
          tc_fence_after();
#ifdef FGNN_TC_TRACE
          if (chunk == 0 && warp == 0) TC_TRACE(ei, 6);
#endif
          const uint32_t taddr = lane_addr + st * kAccCols;
          // this group's half of the chunk comes to registers in one go, and the accumulator stage goes
          // back to the MMA warp before the arithmetic starts
          uint32_t d[NLD][16];
          if constexpr (NLD == 4) {
            tmem_ld64(taddr, reinterpret_cast<uint32_t(&)[64]>(d));
          } else {
#pragma unroll
            for (int gq = 0; gq < NLD; ++gq) tmem_ld16(taddr + gq * 16, d[gq]);
          }
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(t_empty(st));
          ++ct;
          auto fold = [&](int c, float e) {
            if (AGG == FGNN_AGG_MAX) {
              acc[c] = live ? fmaxf(acc[c], e) : acc[c];
            } else if (AGG == FGNN_AGG_SOFTMAX) {
              if (live) softmax_push(acc[c], acc2[c], e, p.gamma);
            } else {
              acc[c] += live ? e : 0.f;
            }
          };
#pragma unroll
          for (int gq = 0; gq < NLD; ++gq) {
            if constexpr (T == 4) {                          // channel pairs, one packed FMA per edge type (tc_common.cuh)
#pragma unroll
              for (int q2 = 0; q2 < 2; ++q2) {
                float e0, e1;
                unpack2(contract_pair4(et2, &d[gq][q2 * 8]), e0, e1);
                const int c = chunk * CPG + gq * CH_PER_LD + q2 * 2;
                fold(c, e0);
                fold(c + 1, e1);
              }
            } else {
#pragma unroll
              for (int q = 0; q < CH_PER_LD; ++q) {
                float e;
                if constexpr (T >= 4) {                      // packed fp32x2 chains (tc_common.cuh)
                  e = contract_types<T>(et, &d[gq][q * T]);
                } else {
                  e = 0.f;
#pragma unroll
                  for (int t = 0; t < T; ++t) e = fmaf(et[t], __uint_as_float(d[gq][q * T + t]), e);
                }
                fold(chunk * CPG + gq * CH_PER_LD + q, e);
              }
            }
          }
        }
#ifdef FGNN_TC_TRACE
        if (warp == 0) TC_TRACE(ei, 7);
        ++ei;
#endif
        live_count += live ? 1.f : 0.f;
      }
      // finish: aggregate, bias / eval-BN / activation (mp_nn.py:162-173) into this quarter's staging
      // rows (16-byte chunk c of row rr at position c ^ (rr & SW): conflict-free both ways) ...
      {
        constexpr int CPR = CH * OB / 16;                    // 16-byte chunks per output row
        constexpr int SW = (CPR < 8 ? CPR : 8) - 1;          // chunk-index bits XOR-ed with the row
        uint8_t* stage = sOut + wq * (32 * CPR * 16);        // this quarter's [32 rows][CPR chunks]
        const float4* epi4 = reinterpret_cast<const float4*>(s_epi);
        const float inv_live = live_count > 0.f ? __frcp_rn(live_count) : 0.f;
        // negative-side slope of the activation: 1 = none, 0 = ReLU, slope = LeakyReLU
        const float neg = p.act == FGNN_ACT_NONE ? 1.f : (p.act == FGNN_ACT_RELU ? 0.f : p.slope);
        named_bar_sync(1 + wq, 64);                          // both warps of the quarter are done reading the previous tile
#ifdef FGNN_TC_TRACE
        if (warp == 0) TC_TRACE(ei - 1, 8);
#endif
        // groups of four channels; the parameters of group g+1 are fetched while group g is computed
        constexpr int NG = CHG / 4, GPC = CPG / 4;           // groups per thread, groups per chunk
        auto cc_of = [&](int g) { return (g / GPC) * (NC / T) + (g % GPC) * 4 + eg * CPG; };   // first CTA-local channel of group g
        float4 prm[2][3];
        prm[0][0] = epi4[cc_of(0) >> 2]; prm[0][1] = epi4[(CH + cc_of(0)) >> 2]; prm[0][2] = epi4[(2 * CH + cc_of(0)) >> 2];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const int cc = cc_of(g);
          if (g + 1 < NG) {
            const int cn = cc_of(g + 1);
            prm[(g + 1) & 1][0] = epi4[cn >> 2]; prm[(g + 1) & 1][1] = epi4[(CH + cn) >> 2]; prm[(g + 1) & 1][2] = epi4[(2 * CH + cn) >> 2];
          }
          const float4 bi = prm[g & 1][0], sc = prm[g & 1][1], sh = prm[g & 1][2];
          const float bia[4] = {bi.x, bi.y, bi.z, bi.w}, sca[4] = {sc.x, sc.y, sc.z, sc.w}, shi[4] = {sh.x, sh.y, sh.z, sh.w};
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int la = g * 4 + j;                        // accumulator index (compile time)
            float a;
            if (AGG == FGNN_AGG_MAX) a = acc[la];
            else if (AGG == FGNN_AGG_SOFTMAX) a = live_count > 0.f ? softmax_finish(acc[la], acc2[la], p.gamma) : -INFINITY;
            else a = __fmul_rn(acc[la], inv_live);
            float y = fmaf(a + bia[j], sca[j], shi[j]);      // bias, then eval BN folded to scale/shift
            y = y >= 0.f ? y : y * neg;
            v[j] = a == -INFINITY ? a : y;                   // no live slot on this shard: stay -inf
          }
          const int byte0 = cc * OB, ck = byte0 >> 4, within = byte0 & 15;
          uint8_t* dst = stage + lane * (CPR * 16) + (((ck & ~SW) | ((ck ^ lane) & SW)) << 4) + within;
          if (XB) {
            *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(__floats2bfloat162_rn(v[0], v[1])),
                                                        pack_bf16(__floats2bfloat162_rn(v[2], v[3])));
          } else {
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
        named_bar_sync(1 + wq, 64);                          // both groups' channels are staged
#ifdef FGNN_TC_TRACE
        if (warp == 0) TC_TRACE(ei - 1, 9);
#endif
        if (!waited) { pdl_wait(); waited = true; }          // first store: the preceding launch (a reader of out) is complete
        // ... then whole lines go out: CPR lanes per row, 32/CPR rows per store instruction, the two
        // warps of the quarter alternate (out is batch-contiguous node-major: row g at out + g * o_sm)
        constexpr int RPI = 32 / CPR, NSI = CPR / 2;         // rows per store instruction, store instructions per warp
        const int cq = lane % CPR, rsub = lane / CPR;
        uint8_t* obase = reinterpret_cast<uint8_t*>(p.out) + (size_t)ch0 * OB + cq * 16;
        const uint32_t ostride = (uint32_t)p.o_sm * OB;      // bytes per output row (< 2^31, tc_supported)
        int32_t orow[NSI];
        uint4 ov[NSI];
#pragma unroll
        for (int i2 = 0; i2 < NSI; ++i2) {                   // warp eg of the quarter: store instructions 2*i2 + eg
          const int rr = (2 * i2 + eg) * RPI + rsub;
          orow[i2] = __shfl_sync(0xffffffffu, my_orow, rr);  // flattened output row (out is batch-contiguous)
          ov[i2] = *reinterpret_cast<const uint4*>(stage + rr * (CPR * 16) + (((cq & ~SW) | ((cq ^ rr) & SW)) << 4));
        }
        if (p.accumulate) {                                  // out += v without waiting on a load: one vector reduction
#pragma unroll
          for (int i2 = 0; i2 < NSI; ++i2) {
            uint8_t* dst = obase + (uint64_t)(uint32_t)orow[i2] * ostride;
            const uint4 v = ov[i2];
            if (orow[i2] >= 0) {
              if (XB) {
                asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                             : "memory");
              } else {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(v.x)),
                             "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                             : "memory");
              }
            }
          }
        } else {
#pragma unroll
          for (int i2 = 0; i2 < NSI; ++i2) {
            uint8_t* dst = obase + (uint64_t)(uint32_t)orow[i2] * ostride;
            if (orow[i2] >= 0) *reinterpret_cast<uint4*>(dst) = ov[i2];
          }
        }
#ifdef FGNN_TC_TRACE
        if (warp == 0) TC_TRACE(ei - 1, 10);
#endif
      }
    }
  } else if (warp < kGatherWarp0) {
    // =====================================================================================
    // CONVERTERS: thread cr == TMEM lane cr == row cr of every item.  raw row (ring) -> A stage in
    // tensor memory: fp32 I/O 32 columns of hi pairs + 32 columns of lo pairs, bf16 I/O 32 columns
    // =====================================================================================
    reg_dec<kRegConv>();
    const int cr = tid - kConvWarp0 * 32;
    const uint32_t row_u = smem_u32(sA) + (uint32_t)cr * ROWB;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + TACOL0;
    uint32_t i = 0;
    for (ItemIter it = first_item(); it.tile < n_tiles; next_item(it), ++i) {
      const uint32_t st = i % NST, use = i / NST;
      const uint32_t ta = i % kTA, tuse = i / kTA;
      mbar_wait(raw_full(st), use & 1);
      if (cr < 32) TC_TRACE(i, 2);
      if (XB) {
        uint32_t a[32];                                      // channels 8c .. 8c+7 in chunk c (stored at c ^ (row & 7))
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = lds_u4(row_u + st * STAGEB + (uint32_t)((c ^ (cr & 7)) * 16));
          a[4 * c] = v.x; a[4 * c + 1] = v.y; a[4 * c + 2] = v.z; a[4 * c + 3] = v.w;
        }
        mbar_arrive(raw_empty(st));                          // ring stage is free again
        mbar_wait(ta_empty(ta), (tuse & 1) ^ 1);
        tc_fence_after();
        tmem_st32(lane_addr + ta * TACOLS, a);
      } else {
        // per K atom (64 channels = 256 bytes of the row): channels 4c .. 4c+3 of the atom in chunk c, stored at
        // c ^ (row & 15) inside the atom's 16 chunks
        uint32_t hi[KA][32], lo[KA][32];
#pragma unroll
        for (int ka = 0; ka < KA; ++ka) {
          float4 v[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = lds_f4(row_u + st * STAGEB + (uint32_t)(ka * 256 + (c ^ (cr & 15)) * 16));
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[c].x, v[c].y), h1 = __floats2bfloat162_rn(v[c].z, v[c].w);
            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            hi[ka][2 * c] = pack_bf16(h0);
            hi[ka][2 * c + 1] = pack_bf16(h1);
            lo[ka][2 * c] = pack_bf16(__floats2bfloat162_rn(v[c].x - f0.x, v[c].y - f0.y));
            lo[ka][2 * c + 1] = pack_bf16(__floats2bfloat162_rn(v[c].z - f1.x, v[c].w - f1.y));
          }
          if (KA == 2 && ka == 0) {                          // two atoms: the first goes to tensor memory before the second is read
            mbar_wait(ta_empty(ta), (tuse & 1) ^ 1);
            tc_fence_after();
            tmem_st32(lane_addr + ta * TACOLS, hi[0]);
            tmem_st32(lane_addr + ta * TACOLS + KA * 32, lo[0]);
          }
        }
        mbar_arrive(raw_empty(st));                          // ring stage is free again: every chunk has been consumed above
        if (KA == 1) {
          mbar_wait(ta_empty(ta), (tuse & 1) ^ 1);
          tc_fence_after();
        }
        tmem_st32(lane_addr + ta * TACOLS + (KA - 1) * 32, hi[KA - 1]);
        tmem_st32(lane_addr + ta * TACOLS + KA * 32 + (KA - 1) * 32, lo[KA - 1]);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ta_full(ta));
      if (cr < 32) TC_TRACE(i, 3);
    }
  } else {
  reg_dec<kRegAux>();                                        // warps 12-15 together (one warpgroup)
  if (warp < kMmaWarp) {
    // =====================================================================================
    // GATHERERS: cp.async the 128 source rows of every item into the raw ring
    // =====================================================================================
    constexpr int LPR = ROWB / 16;                           // lanes per source row (16-byte chunks): 16 | 8
    constexpr int RPW = 32 / LPR;                            // rows per warp instruction: 2 | 4
    constexpr int ROWS_W = kTileM / kGatherWarps;            // rows per gather warp: 64
    constexpr int IPL = ROWS_W / 32;                         // index-table entries per lane: 2
    const int pw = warp - kGatherWarp0;                      // rows pw*64 .. pw*64+63 of the tile
    const int sub = lane / LPR, q = lane % LPR;
    // source rows (b*N + n, -1 = none) of tile rows pw*64 + j*32 + lane of an item; the index loads are
    // issued one item ahead and only CONSUMED (range check -> source row) after the stage wait
    // The loaded words are kept RAW (lo / hi halves): a warp issues in order, so widening a 32-bit entry right behind its
    // load would park the warp for the whole load latency -- 1.6 us per item on a streaming call (the trace of a 1x1 map
    // through the identity table), where "one item ahead" then hides nothing.  index_value() widens at the point of use.
    struct Idx {
      int32_t lo[IPL], hi[IPL];
      int32_t base[IPL];
    };
    auto index_of = [&](const ItemIter& it, Idx& v) {
#pragma unroll
      for (int j = 0; j < IPL; ++j) {
        v.base[j] = -1;
        v.lo[j] = -1;
        v.hi[j] = -1;
        if (it.tile >= n_tiles) continue;
        const uint32_t g = (uint32_t)it.tile * kTileM + pw * ROWS_W + j * 32 + lane;
        if (g >= rows_total) continue;
        uint32_t b, m;
        split_row(g, b, m);
        v.base[j] = (int32_t)(b * (uint32_t)p.N);
        const int64_t off = (int64_t)b * p.idx_sb + (int64_t)m * p.K + it.k;
        if (p.idx64) {
          const int2 w = __ldg(reinterpret_cast<const int2*>(p.idx) + off);
          v.lo[j] = w.x;
          v.hi[j] = w.y;
        } else {
          v.lo[j] = __ldg(reinterpret_cast<const int32_t*>(p.idx) + off);
        }
      }
    };
    auto index_value = [&](const Idx& v, int j) -> int64_t {
      return p.idx64 ? (int64_t)(((uint64_t)(uint32_t)v.hi[j] << 32) | (uint32_t)v.lo[j]) : (int64_t)v.lo[j];
    };
    const uint8_t* xq = reinterpret_cast<const uint8_t*>(p.x) + q * 16;
    const uint32_t sA_u = smem_u32(sA);
    // Extension modes (mp_nn.py:136-159; KA == 2 with C = 64): the 512-byte A row of slot (m, k) is [x_m || x_idx] --
    // K atom 0 = the destination's own row (M == N: row g of x), K atom 1 = the gathered neighbour; the filter image
    // carries the NEIGHBOR / DIFF algebra (w_split_kernel).  Source rows are then 256 bytes apart.
    const bool two_src = KA == 2 && p.ext != 0;
    const uint32_t src_rowb = two_src ? (uint32_t)(ROWB / 2) : (uint32_t)ROWB;
    // Warp instruction u (of NIT per item) copies chunk q of the RPW rows RPW*u + sub of this warp's 64:
    // chunk q of tile row rr lands at position q ^ (rr & (LPR-1)) (the converter's per-row reads are then
    // conflict-free).  The swizzle term repeats every NDO instructions, so the shared-memory offset is one
    // of NDO lane constants plus a compile-time multiple of LPR rows.
    // (two K atoms: 32 lanes per 512-byte row, the swizzle acts inside each atom's 16 chunks and the offset is
    // computed per copy)
    constexpr int NIT = ROWS_W / RPW, NDO = KA == 1 ? LPR / RPW : 1, SWZ = (LPR < 16 ? LPR : 16) - 1;
    uint32_t dst_off[NDO];
#pragma unroll
    for (int c = 0; c < NDO; ++c)
      dst_off[c] = (uint32_t)(pw * ROWS_W + RPW * c + sub) * ROWB + (uint32_t)((q ^ ((RPW * c + sub) & (LPR - 1))) * 16);
    Idx cur, nxt;
    ItemIter it = first_item();
    index_of(it, cur);
    pdl_wait();                                              // x is the preceding launch's output: order behind it
    for (uint32_t i = 0; it.tile < n_tiles; ++i) {
      const uint32_t st = i % NST, use = i / NST;
      const int cur_tile = it.tile;
      (void)cur_tile;
      next_item(it);
      index_of(it, nxt);                                     // index loads of the next item: in flight during this one
      mbar_wait(raw_empty(st), (use & 1) ^ 1);
      if (pw == 0) TC_TRACE(i, 0);
      // byte offset of the source row in x ((b*N + n) * ROWB < 2^32, checked by tc_supported); ~0 = no row
      // (tile tail, masked or out-of-range slot) -> zero-filled
      uint32_t off[IPL];
#pragma unroll
      for (int j = 0; j < IPL; ++j)
      {
        const int64_t n = index_value(cur, j);
        off[j] = (cur.base[j] >= 0 && n >= 0 && n < p.N) ? (uint32_t)(cur.base[j] + (int32_t)n) * src_rowb : 0xffffffffu;
      }

      const uint32_t stage = sA_u + st * STAGEB;
#pragma unroll
      for (int u0 = 0; u0 < NIT; u0 += 8) {                  // batches of 8: all shuffles first, then the copies
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = __shfl_sync(0xffffffffu, off[(RPW * (u0 + j)) / 32], (RPW * (u0 + j) + sub) & 31);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int u = u0 + j;
          const uint32_t dst = KA == 1 ? stage + dst_off[u % NDO] + (uint32_t)((u / NDO) * LPR * ROWB)
                                       : stage + (uint32_t)(pw * ROWS_W + u) * ROWB +
                                             (uint32_t)(((q & ~SWZ) | ((q ^ u) & SWZ)) * 16);
          const bool ok = o[j] != 0xffffffffu;
          if (KA == 2 && two_src) {
            // lanes 0-15: chunk q of the destination's own row; lanes 16-31: chunk q-16 of the neighbour
            const uint32_t g_self = (uint32_t)cur_tile * kTileM + (uint32_t)(pw * ROWS_W + u);
            const uint32_t so = (q < 16 ? g_self * src_rowb : o[j]) + (uint32_t)(q & 15) * 16u;
            cp_async16(dst, reinterpret_cast<const uint8_t*>(p.x) + (ok ? so : 0u), ok ? 16u : 0u);
          } else {
            cp_async16(dst, xq + (ok ? o[j] : 0u), ok ? 16u : 0u);
          }
        }
      }
      cp_async_arrive_noinc(raw_full(st));
      if (pw == 0) TC_TRACE(i, 1);
      cur = nxt;
    }
  } else if (warp == kMmaWarp) {
    // =====================================================================================
    // MMA ISSUER: the whole warp runs the loop (warp-uniform control flow, so addresses and
    // descriptors live in uniform registers); one elected lane issues.  First: the stationary
    // B operand (this CTA's column slice of the bf16 filter image) by TMA bulk copy.
    // =====================================================================================
    const uint32_t sB_u = smem_u32(sB);
    if (elect_one()) {
      const int OT = p.O * p.T;
      constexpr uint32_t part = COLS * 128, piece = part < 32768u ? part : 32768u;      // one (part, K atom) slab
      constexpr int parts = XB ? 1 : 2;
      mbar_expect_tx(w_full, parts * KA * part);
      for (int h = 0; h < parts * KA; ++h) {                   // image and shared memory: [part][ka][column][64 bf16]
        const uint8_t* src = wimg + kHeaderBytes + (size_t)h * OT * 128 + (size_t)col0 * 128;
        for (uint32_t o = 0; o < part; o += piece) bulk_g2s(sB_u + h * part + o, src + o, piece, w_full);
      }
    }
    __syncwarp();
    mbar_wait(w_full, 0);
    constexpr uint32_t idesc = umma_idesc(NC);
    // descriptor of a B tile = constant high word + (address >> 4) in the low word
    const uint64_t desc_hi = umma_desc_sw128(0);
    uint32_t ct = 0, i = 0;
    for (ItemIter it = first_item(); it.tile < n_tiles; next_item(it), ++i) {
      const uint32_t ta = i % kTA, tuse = i / kTA;
      mbar_wait(ta_full(ta), tuse & 1);
      TC_TRACE(i, 4);
      const uint32_t a_hi = tmem_base + TACOL0 + ta * TACOLS, a_lo = a_hi + KA * 32;
#pragma unroll
      for (int chunk = 0; chunk < NCH; ++chunk) {
        const uint32_t ts = ct % ACC;
        mbar_wait(t_empty(ts), ((ct / ACC) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ts * kAccCols;
        const uint32_t b_hi = (sB_u + (uint32_t)chunk * NC * 128u) >> 4, b_lo = b_hi + ((uint32_t)(KA * COLS) * 128u >> 4);
        if (elect_one()) {
#pragma unroll
          for (int term = 0; term < NTERMS; ++term) {        // fp32 I/O: xl*Wh + xh*Wl + xh*Wh; bf16 I/O: x*Wh
            const uint32_t a = (!XB && term == 0) ? a_lo : a_hi;
            const uint32_t bb = (!XB && term == 1) ? b_lo : b_hi;
#pragma unroll
            for (int ka = 0; ka < KA; ++ka)
#pragma unroll
              for (int ks = 0; ks < kC / 16; ++ks)           // 16 bf16 of K = 8 TMEM columns of A = 32 bytes of a B row
                umma_bf16_ts(d_tmem, a + ka * 32 + ks * 8, desc_hi | (uint64_t)(bb + ka * ((uint32_t)COLS * 128u >> 4) + ks * 2),
                             idesc, (term | ka | ks) != 0);
          }
          umma_commit(t_full(ts));                           // accumulator ready when these MMAs retire
          if (chunk == NCH - 1) umma_commit(ta_empty(ta));   // A stage reusable when its readers retire
        }
        __syncwarp();
        ++ct;
      }
      TC_TRACE(i, 5);
    }
  }
  }

  // ---- teardown -----------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: shape gating, configuration, launch
// ---------------------------------------------------------------------------------------------
namespace {

struct TcConfig {
  int T, NC, NCH;
  bool ok;
};

TcConfig pick_config(int T, int agg, int OT, bool xb, int ka = 1) {
  // accumulator chunks of NC <= 128 columns (TMEM: 3 x 128 accumulator + 2 x 64 A-stage columns);
  // channels per CTA (NC*NCH/T) <= 64; accumulator registers per epilogue thread = half of that
  // (twice for softmax).  bf16 I/O: the image has one part, so twice the columns fit a CTA.
  const bool sm = agg == FGNN_AGG_SOFTMAX;
  TcConfig c{T, 128, 1, false};
  switch (T) {
    case 16: c.NCH = xb ? (sm ? 4 : 8) : (sm ? 2 : 4); break;   // bf16: 32 / 64 channels, fp32: 16 / 32
    case 8: c.NCH = xb ? (sm ? 2 : 4) : 2; break;               // bf16: 32 / 64 channels, fp32: 32
    case 4: c.NCH = sm ? 1 : 2; break;                          // 32 / 64 channels
    case 2: c.NCH = 1; c.NC = sm ? 64 : 128; break;
    case 1: c.NC = sm ? 32 : 64; break;
    default: return c;
  }
  if (ka == 2) c.NCH = 1;       // C = 128: the image of 128 columns (two K atoms, hi + lo) is 64 KB beside a 2 x 64 KB ring
  c.ok = OT % (c.NC * c.NCH) == 0;
  return c;
}

size_t smem_bytes(const TcConfig& c, bool xb, int ka = 1) {
  const int cols = c.NC * c.NCH, ch = cols / c.T;
  return 1024 + (size_t)tc::w_bytes(cols, xb, ka) + (size_t)tc::a_stages(cols, ch, xb, ka) * tc::stage_bytes(xb, ka) +
         (size_t)tc::out_tile_bytes(ch, xb) + tc::kNumBars * 8 + 16 + 3 * 64 * 4;
}

bool g_pdl = true;      // programmatic dependent launch between consecutive tensor-core launches

template <int T, int NC, int NCH, int AGG, bool XB, int KA>
int launch_one(const MpParams& p, const uint8_t* wimg, int S, int workers, int tiles, size_t smem, cudaStream_t st) {
  auto kern = mp_tc_kernel<T, NC, NCH, AGG, XB, KA>;
  static bool attr_set = false;                              // per instantiation; the value never changes
  if (!attr_set) {
    // the in-kernel setmaxnreg budget assumes the launch register count ptxas chose (see tc::kRegLaunch):
    // refuse to launch a build that breaks it instead of hanging in setmaxnreg.inc
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return FGNN_ERR_CUDA;
    if (fa.numRegs != tc::kRegLaunch) return FGNN_ERR_UNSUPPORTED;
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
  cfg.numAttrs = g_pdl ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p, wimg, S, workers, tiles);
  count_launch();
  return e == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

// NCHm: chunks per CTA for max & mean, NCHs: for softmax (must mirror pick_config)
template <int T, int NCm, int NCHm, int NCs, int NCHs, bool XB, int KA = 1>
int launch_agg(const MpParams& p, const uint8_t* wimg, int S, int workers, int tiles, size_t smem, cudaStream_t st) {
  switch (p.agg) {
    case FGNN_AGG_MAX: return launch_one<T, NCm, NCHm, FGNN_AGG_MAX, XB, KA>(p, wimg, S, workers, tiles, smem, st);
    case FGNN_AGG_SOFTMAX: return launch_one<T, NCs, NCHs, FGNN_AGG_SOFTMAX, XB, KA>(p, wimg, S, workers, tiles, smem, st);
    case FGNN_AGG_MEAN: return launch_one<T, NCm, NCHm, FGNN_AGG_MEAN, XB, KA>(p, wimg, S, workers, tiles, smem, st);
  }
  return FGNN_ERR_UNSUPPORTED;
}

}  // namespace

void tc_set_pdl(bool on) { g_pdl = on; }

#ifdef FGNN_TC_TRACE
extern "C" int fgnn_debug_trace_read(unsigned long long* host, size_t count) {
  return cudaMemcpyFromSymbol(host, g_trace, count * sizeof(unsigned long long)) == cudaSuccess ? 0 : -6;
}
#endif

// The bf16 image of `W` in `ws` is rebuilt unless this workspace is known to hold the image of exactly these
// filters (pointer + caller-supplied version).  Host-side mirror of the device header: a cached image costs
// no launch at all.  version == 0 always rebuilds.
int tc_prepare_weights(const float* W, uint8_t* ws, int C, int OT, int64_t version, cudaStream_t stream, int diff, int T) {
  if (version != 0) version = version * 4 + (diff ? 1 : 0) + (T == 4 ? 2 : 0);      // the image depends on the transform / layout
  static std::mutex mu;
  static std::unordered_map<const void*, tc::Header> known;
  bool fresh = false;
  if (version != 0) {
    std::lock_guard<std::mutex> lock(mu);
    auto it = known.find(ws);
    fresh = it != known.end() && it->second.version == version && it->second.filters == W && it->second.OT == OT &&
            it->second.C == C;
    if (!fresh) {
      if (known.size() > 4096) known.clear();
      known[ws] = tc::Header{version, W, C, OT};
    }
  } else {
    std::lock_guard<std::mutex> lock(mu);
    known.erase(ws);
  }
  if (!fresh) {
    w_split_kernel<<<(OT * (C / 8) + 255) / 256, 256, 0, stream>>>(W, ws, C, OT, 0, diff, T);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return FGNN_ERR_CUDA;
  }
  return FGNN_OK;
}

int tc_num_sms() {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  return num_sms;
}

bool tc_pdl_enabled() { return g_pdl; }

// K atoms of the A operand: C / 64, or two for the extension modes at C = 64 ([x_m || x_idx])
static int tc_k_atoms(const fgnn_mp_args* a) { return a->extension != FGNN_NO_EXTENSION ? 2 : a->C / tc::kC; }

bool tc_supported(const fgnn_mp_args* a) {
  if (a->dtype != FGNN_F32 && a->dtype != FGNN_BF16) return false;
  const bool xb = a->dtype == FGNN_BF16;
  if (a->extension != FGNN_NO_EXTENSION) {
    // NEIGHBOR / DIFF (mp_nn.py:136-159) as a two-atom row [x_m || x_idx] against a transformed filter image
    if (a->C != tc::kC || xb || a->M != a->N || a->tile_slots || a->out_rows) return false;
  } else if (a->C != tc::kC && !(a->C == 2 * tc::kC && !xb)) {
    return false;                                                           // C = 128: fp32 I/O only
  }
  const int ka = tc_k_atoms(a);
  if (a->aggregator == FGNN_AGG_NONE) return false;
  if (a->x_sc != 1 || a->x_sn != a->C) return false;                              // node-major rows
  if (a->B > 1 && a->x_sb != (int64_t)a->N * a->C) return false;                  // batch-contiguous
  if ((int64_t)a->B * a->N * tc::row_bytes(xb, ka) >= (int64_t)UINT32_MAX) return false;   // 32-bit source byte offsets
  if (a->out_sm * (xb ? 2 : 4) >= (int64_t)INT32_MAX) return false;
  if ((reinterpret_cast<uintptr_t>(a->x) & 15) || (reinterpret_cast<uintptr_t>(a->out) & 15)) return false;
  if (a->out_so != 1 || (a->out_sm & (xb ? 7 : 3))) return false;                   // 16-byte aligned output rows
  if (a->B > 1 && a->out_sb != (int64_t)a->M * a->out_sm) return false;             // batch-contiguous rows
  if (a->O % (xb ? 8 : 4)) return false;
  const TcConfig c = pick_config(a->T, a->aggregator, a->O * a->T, xb, ka);
  if (!c.ok) return false;
  if (smem_bytes(c, xb, ka) > (size_t)tc::kSmemBudget || tc::a_stages(c.NC * c.NCH, c.NC * c.NCH / c.T, xb, ka) < 2) return false;
  if ((int64_t)a->B * a->M >= (int64_t)INT32_MAX - 256) return false;
  return true;
}

size_t tc_workspace_bytes(const fgnn_mp_args* a) {
  return tc::kHeaderBytes + (size_t)2 * tc_k_atoms(a) * a->O * a->T * 128;
}

int launch_mp_tc(const MpParams& p, const fgnn_mp_args* a, cudaStream_t stream) {
  const bool xb = a->dtype == FGNN_BF16;
  const int ka = tc_k_atoms(a);
  TcConfig c = pick_config(p.T, p.agg, p.O * p.T, xb, ka);
  if (!c.ok) return FGNN_ERR_UNSUPPORTED;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  if (reinterpret_cast<uintptr_t>(ws) & 255) return FGNN_ERR_WORKSPACE;
  const int OT = p.O * p.T;
  const int wrc = tc_prepare_weights(p.W, ws, ka * tc::kC, OT, a->filters_version, stream, p.ext == FGNN_ORIG_WITH_DIFF, p.T);
  if (wrc != FGNN_OK) return wrc;

  const int num_sms = tc_num_sms();
  const int64_t rows = (int64_t)p.B * p.M;
  const int tiles = (int)((rows + tc::kTileM - 1) / tc::kTileM);
  const int S = OT / (c.NC * c.NCH);
  int sms = num_sms;
  if (a->sm_limit > 0 && a->sm_limit < sms) sms = a->sm_limit;
  if (S > sms) return FGNN_ERR_UNSUPPORTED;
  int workers = sms / S;
  if (workers > tiles) workers = tiles;
  const size_t smem = smem_bytes(c, xb, ka);
  if (ka == 2) {                    // C = 128 (fp32 I/O): one 128-column chunk per CTA
    switch (p.T) {
      case 16: return launch_agg<16, 128, 1, 128, 1, false, 2>(p, ws, S, workers, tiles, smem, stream);
      case 8: return launch_agg<8, 128, 1, 128, 1, false, 2>(p, ws, S, workers, tiles, smem, stream);
      case 4: return launch_agg<4, 128, 1, 128, 1, false, 2>(p, ws, S, workers, tiles, smem, stream);
      case 2: return launch_agg<2, 128, 1, 64, 1, false, 2>(p, ws, S, workers, tiles, smem, stream);
      case 1: return launch_agg<1, 64, 1, 32, 1, false, 2>(p, ws, S, workers, tiles, smem, stream);
    }
    return FGNN_ERR_UNSUPPORTED;
  }
  if (xb) {
    switch (p.T) {
      case 16: return launch_agg<16, 128, 8, 128, 4, true>(p, ws, S, workers, tiles, smem, stream);
      case 8: return launch_agg<8, 128, 4, 128, 2, true>(p, ws, S, workers, tiles, smem, stream);
      case 4: return launch_agg<4, 128, 2, 128, 1, true>(p, ws, S, workers, tiles, smem, stream);
      case 2: return launch_agg<2, 128, 1, 64, 1, true>(p, ws, S, workers, tiles, smem, stream);
      case 1: return launch_agg<1, 64, 1, 32, 1, true>(p, ws, S, workers, tiles, smem, stream);
    }
  } else {
    switch (p.T) {
      case 16: return launch_agg<16, 128, 4, 128, 2, false>(p, ws, S, workers, tiles, smem, stream);
      case 8: return launch_agg<8, 128, 2, 128, 2, false>(p, ws, S, workers, tiles, smem, stream);
      case 4: return launch_agg<4, 128, 2, 128, 1, false>(p, ws, S, workers, tiles, smem, stream);
      case 2: return launch_agg<2, 128, 1, 64, 1, false>(p, ws, S, workers, tiles, smem, stream);
      case 1: return launch_agg<1, 64, 1, 32, 1, false>(p, ws, S, workers, tiles, smem, stream);
    }
  }
  return FGNN_ERR_UNSUPPORTED;
}

}  // namespace fgnn

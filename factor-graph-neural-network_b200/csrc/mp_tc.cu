// tcgen05 / TMEM kernel for the FGNN message-passing call (NO_EXTENSION, C = 64, fp32 I/O).
//
//   out[b,o,m] = act(BN(bias[o] + AGG_k sum_t etype[b,t,m,k] * (x[b, idx[b,m,k], :] . W[:, o*T+t])))
//   reference: lib/model/mpnn/mp_nn.py:115-175 (the VF and the FV module of FGNN)
//
// Formulation.  A destination tile = 128 consecutive (b,m) rows.  For every slot k the 128 source
// rows x[idx[.,k]] form the A operand [128 x C]; the filters form the B operand [C x O*T]; one
// UMMA M=128 accumulates H_k = A_k W into TMEM (fp32), NC columns at a time.  TMEM lane r == row r,
// so epilogue thread r reads its own H_k row, contracts it with its slot's edge-type vector
// (sum_t et[t] * H[o*T+t]) and folds the result into a running max / logsumexp / mean held in
// registers: the K-reduction needs no shuffles and no shared memory, and nothing O*T wide ever
// reaches HBM (the reference materialises H, an int64 index expansion and the gathered rows).
//
// fp32 parity on bf16 tensor cores: both operands are split x = xh + xl, W = Wh + Wl (bf16 each)
// and three MMAs xh*Wh + xl*Wh + xh*Wl accumulate in fp32 (error ~2^-16 relative, inside the
// 1e-4 contract; plain TF32/bf16 is not -- SURVEY 7 hard part 2).
//
// CTA roles (288 threads, 1 CTA / SM, persistent over tiles):
//   warps 0-3  epilogue   TMEM -> registers, edge-type contraction, aggregate, bias/BN/act, store
//   warps 4-7  producers  gather x rows (128-bit coalesced: 16 lanes per 256-byte row), split to
//                         bf16 hi/lo, write the UMMA K-major SWIZZLE_128B image of A into smem
//   warp  8    MMA        one elected lane issues tcgen05.mma; owns the TMEM allocation
// Pipelines: A stages (2) producers<->MMA, TMEM accumulator stages (2) MMA<->epilogue, all mbarrier.
// The filters are stationary: each CTA keeps the split-bf16 image of its column slice
// (O*T / S columns, S = column split across CTAs so the slice fits in shared memory) for its
// whole lifetime; the image is produced once per weight version by w_split_kernel.
#include <cuda_bf16.h>

#include "common.cuh"

namespace fgnn {

namespace tc {

constexpr int kC = 64;                 // input channels (K dimension of the MMA), one 128-byte swizzle atom
constexpr int kTileM = 128;            // destinations per tile == UMMA M == TMEM lanes
constexpr int kEpiWarps = 4, kProdWarps = 4;
constexpr int kThreads = (kEpiWarps + kProdWarps + 1) * 32;   // 288
constexpr int kAStages = 2;
constexpr int kAPartBytes = kTileM * kC * 2;                  // 16 KB: one bf16 part (hi or lo) of A
constexpr int kAStageBytes = 2 * kAPartBytes;                 // hi + lo
constexpr int kHeaderBytes = 256;                             // workspace header in front of the W image
constexpr uint32_t kSpinLimit = 1u << 24;                     // watchdog: trap instead of hanging the GPU

struct Header {                       // first bytes of the workspace
  int64_t version;                    // fgnn_mp_args.filters_version the image was built from
  const float* filters;               // and the pointer it was built from
  int32_t C, OT;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, M=128, N from idesc, K=16
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B, rows of 128 bytes (64 bf16), 8-row groups
// 1024 bytes apart (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64)).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32=1 [4,6), a/b_format BF16=1
// [7,10)/[10,13), K-major both (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

// Register re-balancing between warp groups (warps 0-3 epilogue, 4-7 producers): the kernel is
// compiled for 168 registers/thread; the epilogue grows to 232, the producers shrink to 96.
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// filters [C, O*T] fp32 -> split-bf16 image in the UMMA B layout (K-major rows of 64 bf16 = 128 B,
// 16-byte chunks XOR-swizzled with row % 8): image[part][n][c], n = column o*T+t.
// ---------------------------------------------------------------------------------------------
__global__ void w_split_kernel(const float* __restrict__ W, uint8_t* __restrict__ ws, int OT, int64_t version) {
  tc::Header* h = reinterpret_cast<tc::Header*>(ws);
  if (version != 0 && h->version == version && h->filters == W && h->C == tc::kC && h->OT == OT) return;
  uint8_t* img = ws + tc::kHeaderBytes;
  const int total = OT * (tc::kC / 8);                       // one 16-byte chunk (8 channels) per thread-iteration
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % OT, chunk = i / OT;                    // consecutive threads: consecutive columns (coalesced)
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = W[(int64_t)(chunk * 8 + 2 * j) * OT + n], b = W[(int64_t)(chunk * 8 + 2 * j + 1) * OT + n];
      const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
      const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah));
      const __nv_bfloat16 bl = __float2bfloat16_rn(b - __bfloat162float(bh));
      hi[j] = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
      lo[j] = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
    }
    const size_t off = (size_t)n * 128 + (size_t)((chunk ^ (n & 7)) * 16);
    *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + (size_t)OT * 128 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {                 // read by later launches only (stream order)
    h->version = version; h->filters = W; h->C = tc::kC; h->OT = OT;
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel
//   T   edge types (columns per output channel)        NC  accumulator columns per MMA chunk
//   NCH chunks per CTA (CTA column slice = NC*NCH)     AGG aggregator
// ---------------------------------------------------------------------------------------------
template <int T, int NC, int NCH, int AGG>
__global__ void __launch_bounds__(tc::kThreads, 1)
mp_tc_kernel(const MpParams p, const uint8_t* __restrict__ wimg, const int S, const int n_workers,
             const int n_tiles) {
  using namespace tc;
  constexpr int COLS = NC * NCH;               // columns of W this CTA owns
  constexpr int CH = COLS / T;                 // output channels this CTA owns
  constexpr int CH_PER_LD = 16 / T;            // channels per 16-column TMEM load
  constexpr int TMEM_COLS = (2 * NC) < 32 ? 32 : 2 * NC;
  static_assert(NC % 32 == 0 && NC <= 256 && 16 % T == 0 && CH <= 64, "unsupported shape");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                        // [2 parts][COLS rows][128 B]
  uint8_t* sA = sB + 2 * COLS * 128;                         // [kAStages][2 parts][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + kAStages * kAStageBytes);
  // bars: a_full[2], a_empty[2], t_full[2], t_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (6 + s); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x % S, worker = blockIdx.x / S;
  const int col0 = split * COLS;                             // first W column of this CTA
  const int ch0 = col0 / T;                                  // first output channel of this CTA
  const int64_t rows_total = (int64_t)p.B * p.M;

  // ---- one-time setup ---------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), kProdWarps * 32);
      mbar_init(a_empty(s), 1);
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), kEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps + kProdWarps) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  // stationary B: this CTA's column slice of the split-bf16 filter image
  {
    const int OT = p.O * p.T;
    const uint4* src_hi = reinterpret_cast<const uint4*>(wimg + kHeaderBytes + (size_t)col0 * 128);
    const uint4* src_lo = reinterpret_cast<const uint4*>(wimg + kHeaderBytes + (size_t)OT * 128 + (size_t)col0 * 128);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    constexpr int n16 = COLS * 128 / 16;
    for (int i = tid; i < n16; i += kThreads) {
      dst[i] = __ldg(src_hi + i);
      dst[n16 + i] = __ldg(src_lo + i);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // =====================================================================================
    // EPILOGUE: thread r owns destination row (tile*128 + r) and TMEM lane r
    // =====================================================================================
    reg_inc<232>();
    const int r = tid;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t ct = 0;
    for (int tile = worker; tile < n_tiles; tile += n_workers) {
      const int64_t g = (int64_t)tile * kTileM + r;
      const bool valid = g < rows_total;
      const int b = valid ? (int)(g / p.M) : 0;
      const int m = valid ? (int)(g % p.M) : 0;
      float acc[CH];                                         // max | running max of gamma*e | sum
      float acc2[AGG == FGNN_AGG_SOFTMAX ? CH : 1];          // softmax: running sum of exp
      float live_count = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = (AGG == FGNN_AGG_MEAN) ? 0.f : -INFINITY;
#pragma unroll
      for (int c = 0; c < (AGG == FGNN_AGG_SOFTMAX ? CH : 1); ++c) acc2[c] = 0.f;
      const float* et_row = p.et + (int64_t)b * p.et_sb + (int64_t)m * p.K;
      const int64_t et_st = (int64_t)p.M * p.K;              // stride between edge types
      for (int k = 0; k < p.K; ++k) {
        // this slot's edge-type vector (issued before the accumulator wait so the loads overlap the MMA)
        float et[T];
        bool live = valid;
        if (valid && p.mask_neg) live = load_index(p.idx, p.idx64, (int64_t)b * p.idx_sb + (int64_t)m * p.K + k) >= 0;
        {
          const float* pe = et_row + k;                       // one live pointer, bumped per edge type
#pragma unroll
          for (int t = 0; t < T; ++t) {
            et[t] = valid ? __ldg(pe) : 0.f;
            pe += et_st;
          }
        }
#pragma unroll
        for (int chunk = 0; chunk < NCH; ++chunk) {
          const uint32_t st = ct & 1;
          mbar_wait(t_full(st), (ct >> 1) & 1);
          tc_fence_after();
          const uint32_t taddr = lane_addr + st * NC;
          uint32_t d[2][16];
          tmem_ld16(taddr, d[0]);
#pragma unroll
          for (int gq = 0; gq < NC / 16; ++gq) {
            tmem_ld_wait();
            if (gq + 1 < NC / 16) tmem_ld16(taddr + (gq + 1) * 16, d[(gq + 1) & 1]);
#pragma unroll
            for (int q = 0; q < CH_PER_LD; ++q) {
              float e = 0.f;
#pragma unroll
              for (int t = 0; t < T; ++t) e = fmaf(et[t], __uint_as_float(d[gq & 1][q * T + t]), e);
              const int c = chunk * (NC / T) + gq * CH_PER_LD + q;
              if (AGG == FGNN_AGG_MAX) {
                acc[c] = live ? fmaxf(acc[c], e) : acc[c];
              } else if (AGG == FGNN_AGG_SOFTMAX) {
                if (live) {
                  const float z = p.gamma * e, mx = fmaxf(acc[c], z);
                  acc2[c] = acc2[c] * expf(acc[c] - mx) + expf(z - mx);
                  acc[c] = mx;
                }
              } else {
                acc[c] += live ? e : 0.f;
              }
            }
          }
          tc_fence_before();
          mbar_arrive(t_empty(st));
          ++ct;
        }
        live_count += live ? 1.f : 0.f;
      }
      // finish: aggregate, bias / eval-BN / activation (mp_nn.py:162-173), store this row's channels
      if (valid) {
        float* orow = p.out + (int64_t)b * p.o_sb + (int64_t)m * p.o_sm + ch0;
#pragma unroll
        for (int c4 = 0; c4 < CH; c4 += 4) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = c4 + j;
            float a;
            if (AGG == FGNN_AGG_MAX) a = acc[c];
            else if (AGG == FGNN_AGG_SOFTMAX) a = live_count > 0.f ? (logf(acc2[c]) + acc[c]) / p.gamma : -INFINITY;
            else a = live_count > 0.f ? acc[c] / live_count : 0.f;
            if (a != -INFINITY) a = apply_epilogue(a, ch0 + c, p);
            v[j] = a;
          }
          float4* dst = reinterpret_cast<float4*>(orow + c4);
          if (p.accumulate) {
            const float4 old = *dst;
            v[0] += old.x; v[1] += old.y; v[2] += old.z; v[3] += old.w;
          }
          *dst = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  } else if (warp < kEpiWarps + kProdWarps) {
    // =====================================================================================
    // PRODUCERS: gather + split the A operand of every (tile, k)
    // =====================================================================================
    reg_dec<96>();
    const int pw = warp - kEpiWarps;                         // rows pw*32 .. pw*32+31 of the tile
    const int sub = lane >> 4, q = lane & 15;                // 16 lanes x float4 = one 256-byte row
    uint32_t it = 0;
    for (int tile = worker; tile < n_tiles; tile += n_workers) {
      const int64_t g = (int64_t)tile * kTileM + pw * 32 + lane;   // the row this lane owns for index loads
      const bool valid = g < rows_total;
      const int b = valid ? (int)(g / p.M) : 0;
      const int m = valid ? (int)(g % p.M) : 0;
      const int64_t idx_off = (int64_t)b * p.idx_sb + (int64_t)m * p.K;
      const int64_t x_base = (int64_t)b * p.x_sb;
      for (int k = 0; k < p.K; ++k) {
        int64_t src = -1;                                    // element offset of the source row, -1 = no row
        if (valid) {
          const int64_t n = load_index(p.idx, p.idx64, idx_off + k);
          if (n >= 0 && n < p.N) src = x_base + n * p.x_sn;
        }
        const uint32_t st = it & 1;
        mbar_wait(a_empty(st), ((it >> 1) & 1) ^ 1);
        uint8_t* a_hi = sA + st * kAStageBytes;
        uint8_t* a_lo = a_hi + kAPartBytes;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int rw = 2 * i + sub;                        // row within this warp's 32
          const int64_t off = __shfl_sync(0xffffffffu, src, rw);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (off >= 0) v = ldg_f4(p.x + off + q * 4);
          const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
          const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
          const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
          const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
          const int row = pw * 32 + rw;
          const uint32_t o = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u +
                             (uint32_t)(((q >> 1) ^ (row & 7)) * 16) + (uint32_t)(q & 1) * 8u;
          *reinterpret_cast<uint2*>(a_hi + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01),
                                                           *reinterpret_cast<const uint32_t*>(&h23));
          *reinterpret_cast<uint2*>(a_lo + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01),
                                                           *reinterpret_cast<const uint32_t*>(&l23));
        }
        fence_proxy_async();                                 // generic-proxy stores -> visible to the MMA (async proxy)
        mbar_arrive(a_full(st));
        ++it;
      }
    }
  } else {
    // =====================================================================================
    // MMA ISSUER: one lane
    // =====================================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(NC);
      const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
      uint32_t it = 0, ct = 0;
      for (int tile = worker; tile < n_tiles; tile += n_workers) {
        for (int k = 0; k < p.K; ++k) {
          const uint32_t st = it & 1;
          mbar_wait(a_full(st), (it >> 1) & 1);
          tc_fence_after();
          const uint32_t a_hi = sA_u + st * kAStageBytes, a_lo = a_hi + kAPartBytes;
#pragma unroll
          for (int chunk = 0; chunk < NCH; ++chunk) {
            const uint32_t ts = ct & 1;
            mbar_wait(t_empty(ts), ((ct >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + ts * NC;
            const uint32_t b_hi = sB_u + (uint32_t)chunk * NC * 128u, b_lo = b_hi + (uint32_t)COLS * 128u;
            uint32_t accumulate = 0;
#pragma unroll
            for (int term = 0; term < 3; ++term) {           // xl*Wh + xh*Wl + xh*Wh
              const uint32_t a = term == 0 ? a_lo : a_hi;
              const uint32_t bb = term == 1 ? b_lo : b_hi;
#pragma unroll
              for (int ks = 0; ks < kC / 16; ++ks) {
                umma_bf16(d_tmem, umma_desc_sw128(a + ks * 32), umma_desc_sw128(bb + ks * 32), idesc, accumulate);
                accumulate = 1;
              }
            }
            umma_commit(t_full(ts));                         // accumulator ready when these MMAs retire
            ++ct;
          }
          umma_commit(a_empty(st));                          // A stage reusable when its readers retire
          ++it;
        }
      }
    }
    __syncwarp();
  }

  // ---- teardown -----------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + kProdWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
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

TcConfig pick_config(int T, int agg, int OT) {
  // channels per CTA (NC*NCH/T) <= 64 for max/mean, <= 32 for softmax (two registers per channel)
  const bool sm = agg == FGNN_AGG_SOFTMAX;
  TcConfig c{T, 0, 1, false};
  switch (T) {
    case 16: c.NC = 256; c.NCH = 2; break;
    case 8: c.NC = 256; c.NCH = 1; break;
    case 4: c.NC = sm ? 128 : 256; break;
    case 2: c.NC = sm ? 64 : 128; break;
    case 1: c.NC = sm ? 32 : 64; break;
    default: return c;
  }
  c.ok = OT % (c.NC * c.NCH) == 0;
  return c;
}

size_t smem_bytes(const TcConfig& c) {
  return 1024 + (size_t)2 * c.NC * c.NCH * 128 + (size_t)tc::kAStages * tc::kAStageBytes + 8 * 8 + 16;
}

template <int T, int NC, int NCH>
int launch_agg(const MpParams& p, const uint8_t* wimg, int S, int workers, int tiles, size_t smem, cudaStream_t st) {
  auto go = [&](auto kern) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return (int)FGNN_ERR_CUDA;
    kern<<<S * workers, tc::kThreads, smem, st>>>(p, wimg, S, workers, tiles);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
  };
  switch (p.agg) {
    case FGNN_AGG_MAX: return go(mp_tc_kernel<T, NC, NCH, FGNN_AGG_MAX>);
    case FGNN_AGG_SOFTMAX: return go(mp_tc_kernel<T, (NC > 32 ? NC / 2 : NC), NCH, FGNN_AGG_SOFTMAX>);
    case FGNN_AGG_MEAN: return go(mp_tc_kernel<T, NC, NCH, FGNN_AGG_MEAN>);
  }
  return FGNN_ERR_UNSUPPORTED;
}

}  // namespace

bool tc_supported(const fgnn_mp_args* a) {
  if (a->extension != FGNN_NO_EXTENSION || a->dtype != FGNN_F32) return false;
  if (a->C != tc::kC) return false;
  if (a->aggregator == FGNN_AGG_NONE) return false;
  if (a->x_sc != 1 || a->x_sn != a->C || (a->x_sb & 3)) return false;          // node-major, 16-byte rows
  if ((reinterpret_cast<uintptr_t>(a->x) & 15) || (reinterpret_cast<uintptr_t>(a->out) & 15)) return false;
  if (a->out_so != 1 || (a->out_sm & 3) || (a->out_sb & 3)) return false;
  if (a->O % 4) return false;
  const TcConfig c = pick_config(a->T, a->aggregator, a->O * a->T);
  if (!c.ok) return false;
  if (smem_bytes(c) > 227 * 1024) return false;
  if ((int64_t)a->B * a->M > (int64_t)INT32_MAX * 64) return false;
  return true;
}

size_t tc_workspace_bytes(const fgnn_mp_args* a) {
  return tc::kHeaderBytes + (size_t)2 * a->O * a->T * 128;
}

int launch_mp_tc(const MpParams& p, const fgnn_mp_args* a, cudaStream_t stream) {
  TcConfig c = pick_config(p.T, p.agg, p.O * p.T);
  if (!c.ok) return FGNN_ERR_UNSUPPORTED;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  if (reinterpret_cast<uintptr_t>(ws) & 255) return FGNN_ERR_WORKSPACE;
  const int OT = p.O * p.T;
  w_split_kernel<<<(OT * (tc::kC / 8) + 255) / 256, 256, 0, stream>>>(p.W, ws, OT, a->filters_version);
  count_launch();
  if (cudaGetLastError() != cudaSuccess) return FGNN_ERR_CUDA;

  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int64_t rows = (int64_t)p.B * p.M;
  const int tiles = (int)((rows + tc::kTileM - 1) / tc::kTileM);
  // softmax halves NC for T in {16,8} (see launch_agg), so its column split doubles
  int nc_eff = c.NC;
  if (p.agg == FGNN_AGG_SOFTMAX && (p.T == 16 || p.T == 8)) nc_eff = c.NC / 2;
  const int S = OT / (nc_eff * c.NCH);
  if (S > num_sms) return FGNN_ERR_UNSUPPORTED;
  int workers = num_sms / S;
  if (workers > tiles) workers = tiles;
  TcConfig ce = c;
  ce.NC = nc_eff;
  const size_t smem = smem_bytes(ce);
  switch (p.T) {
    case 16: return launch_agg<16, 256, 2>(p, ws, S, workers, tiles, smem, stream);
    case 8: return launch_agg<8, 256, 1>(p, ws, S, workers, tiles, smem, stream);
    case 4: return launch_agg<4, 256, 1>(p, ws, S, workers, tiles, smem, stream);
    case 2: return launch_agg<2, 128, 1>(p, ws, S, workers, tiles, smem, stream);
    case 1: return launch_agg<1, 64, 1>(p, ws, S, workers, tiles, smem, stream);
  }
  return FGNN_ERR_UNSUPPORTED;
}

}  // namespace fgnn

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
// and three MMAs xl*Wh + xh*Wl + xh*Wh accumulate in fp32 (error ~2^-16 relative, inside the
// 1e-4 contract; plain TF32/bf16 is not -- SURVEY 7 hard part 2).
//
// Shared-memory bandwidth is the scarce resource (measured: with both operands in shared memory the
// MMA reads 12 KB per 128-cycle instruction and runs at half rate while gather/convert starve), so
// the A operand lives in TENSOR MEMORY: the converter threads write the split-bf16 rows straight
// into TMEM with tcgen05.st (thread r == lane r == row r) and the MMA takes A from TMEM; shared
// memory only carries the raw gather ring and the stationary filter slice.
//
// CTA roles (416 threads, 1 CTA / SM, persistent over tiles); an ITEM is one (tile, k):
//   warps 0-3   epilogue    TMEM -> registers, edge-type contraction, aggregate, bias/BN/act, store;
//                           the next item's edge-type vector is prefetched during the current one
//   warps 4-7   converters  thread r: raw fp32 row r from the ring (conflict-free swizzled chunks) ->
//                           bf16 hi/lo pairs -> tcgen05.st into the A stage in TMEM
//   warps 8-11  gatherers   cp.async (LDGSTS) 16-byte chunks of the 128 source rows into the raw ring:
//                           no register staging, every free ring stage's gather is in flight; the
//                           item's indices are loaded one item ahead
//   warp  12    MMA         one elected lane issues tcgen05.mma (A: TMEM, B: smem descriptor); owns the
//                           TMEM allocation; brings the stationary filter slice in with TMA bulk copies
// TMEM (512 columns): 3 accumulator stages x 128 columns + 2 A stages x 64 columns (32 hi + 32 lo).
// Pipelines (all mbarrier): raw ring raw_empty -> raw_full; A stages ta_empty -> ta_full;
// accumulators t_empty -> t_full.
// The filters are stationary: each CTA keeps the split-bf16 image of its column slice
// (O*T / S columns, S = column split across CTAs so the slice fits in shared memory) for its
// whole lifetime; the image is produced once per weight version by w_split_kernel.
#include <cuda_bf16.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace fgnn {

#ifdef FGNN_TC_TRACE
// Debug builds only (-DFGNN_TC_TRACE): per-item timestamps of CTA 0, read back by tools/tc_trace.py.
__device__ unsigned long long g_trace[16 * 4096];
#define TC_TRACE(item, slot)                                                                    \
  do {                                                                                          \
    if (blockIdx.x == 0 && (item) < 4096u && (threadIdx.x & 31) == 0) {                         \
      unsigned long long _t;                                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                    \
      g_trace[(item) * 16 + (slot)] = _t;                                                       \
    }                                                                                           \
  } while (0)
#else
#define TC_TRACE(item, slot) do { } while (0)
#endif

namespace tc {

constexpr int kC = 64;                 // input channels (K dimension of the MMA), one 128-byte swizzle atom
constexpr int kTileM = 128;            // destinations per tile == UMMA M == TMEM lanes
constexpr int kEpiWarps = 4, kGatherWarps = 4, kConvWarps = 4;
constexpr int kMmaWarp = kEpiWarps + kGatherWarps + kConvWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;                 // 416
constexpr int kMaxAStages = 6;                                // raw ring stages in shared memory
constexpr int kTA = 2;                                        // A stages in tensor memory
constexpr int kAcc = 3;                                       // accumulator stages in tensor memory
constexpr int kAccCols = 128, kTACol0 = kAcc * kAccCols, kTACols = 64;   // TMEM column map: 3 x 128 + 2 x 64 = 512
constexpr int kNumBars = 2 * kMaxAStages + 2 * kTA + 2 * kAcc + 1;    // raw_full, raw_empty, ta_full, ta_empty, t_full, t_empty, w_full
constexpr int kSmemBudget = 227 * 1024;
constexpr int kAStageBytes = kTileM * kC * 4;                 // 32 KB: 128 raw fp32 rows
constexpr int kHeaderBytes = 256;                             // workspace header in front of the W image
constexpr uint32_t kSpinLimit = 1u << 20;                     // watchdog: trap instead of hanging the GPU

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
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 16-byte asynchronous copy global -> shared (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier gets one (pre-counted) arrival when all of this thread's prior cp.async have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }
// one lane of a converged warp (warp-uniform control flow keeps descriptors in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x 8 columns (16 bf16 per row, two per column)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
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
// compiled for 128 registers/thread (416 threads); setmaxnreg moves registers WITHIN the CTA's
// launch allocation (416 x 128 = 53248), so 128*200 (epilogue) + 128*120 (converters) + 128*56
// (gatherers) + 32*128 (MMA warp) = 52224 must fit in it -- an over-subscribed inc never returns.
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// raw-ring stages that fit beside a filter slice of `cols` columns and the output staging tile of `ch`
// channels (plus 1 KB alignment slack, barriers, epilogue params)
__host__ __device__ constexpr int a_stages(int cols, int ch) {
  int n = (kSmemBudget - 1024 - 2 * cols * 128 - kTileM * ch * 4 - 2048) / kAStageBytes;
  return n > kMaxAStages ? kMaxAStages : n;
}

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
  constexpr int NST = a_stages(COLS, CH);      // raw-ring stages that fit beside the filter slice and the output tile
  static_assert(NC % 16 == 0 && NC <= kAccCols && 16 % T == 0 && CH <= 64 && NST >= 2, "unsupported shape");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1 KB alignment by OFFSET (not by integer round-trip) so the compiler keeps the shared state space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                        // [2 parts][COLS rows][128 B]   UMMA K-major SW128
  uint8_t* sA = sB + 2 * COLS * 128;                         // [NST][128 rows][256 B]        raw fp32 ring
  uint8_t* sOut = sA + NST * kAStageBytes;                   // [4 warps][32 rows][CH floats]  output staging (swizzled chunks)
  float* s_epi = reinterpret_cast<float*>(sOut + kTileM * CH * 4);    // [3][64]: bias, BN scale, BN shift (16-byte aligned)
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
  const int my_tiles = worker < n_tiles ? (n_tiles - worker + n_workers - 1) / n_workers : 0;
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
  (void)my_tiles;

  // ---- one-time setup ---------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < kMaxAStages; ++s) {
      mbar_init(raw_full(s), 128);                           // gather threads (cp.async arrive-on)
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
    // EPILOGUE: thread r owns destination row (tile*128 + r) and TMEM lane r
    // =====================================================================================
    reg_inc<200>();
    const int r = tid;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int64_t et_st = (int64_t)p.M * p.K;                // stride between edge types
    // edge-type vector + liveness of item (tile j, slot k) for this thread's row
    float et_nx[T];
    bool live_nx = false;
    auto fetch = [&](int tile, int k) {
      const uint32_t g = (uint32_t)tile * kTileM + r;
      const bool ok = tile < n_tiles && g < rows_total;
      uint32_t b = 0, m = 0;
      if (ok) split_row(g, b, m);
      const float* pe = p.et + (int64_t)b * p.et_sb + (int64_t)m * p.K + k;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        et_nx[t] = ok ? __ldg(pe) : 0.f;
        pe += et_st;
      }
      live_nx = ok;
      if (ok && p.mask_neg) live_nx = load_index(p.idx, p.idx64, (int64_t)b * p.idx_sb + (int64_t)m * p.K + k) >= 0;
    };
    fetch(worker, 0);
    uint32_t ct = 0;
#ifdef FGNN_TC_TRACE
    uint32_t ei = 0;
#endif
    for (int tile = worker; tile < n_tiles; tile += n_workers) {
      float acc[CH];                                         // max | running max of gamma*e | sum
      float acc2[AGG == FGNN_AGG_SOFTMAX ? CH : 1];          // softmax: running sum of exp
      float live_count = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = (AGG == FGNN_AGG_MEAN) ? 0.f : -INFINITY;
#pragma unroll
      for (int c = 0; c < (AGG == FGNN_AGG_SOFTMAX ? CH : 1); ++c) acc2[c] = 0.f;
      const int kt = slots_of(tile);
      // output row of tile row (warp*32 + lane), fetched now so the finish phase never waits on it
      int32_t my_orow = -1;
      {
        const uint32_t go = (uint32_t)tile * kTileM + warp * 32 + lane;
        if (go < rows_total) my_orow = p.out_rows ? p.out_rows[go] : (int32_t)go;
      }
      for (int k = 0; k < kt; ++k) {
        float et[T];
#pragma unroll
        for (int t = 0; t < T; ++t) et[t] = et_nx[t];
        const bool live = live_nx;
        // prefetch the next item's edge types: their latency hides behind this item's math
        if (k + 1 < kt) fetch(tile, k + 1); else fetch(tile + n_workers, 0);
#pragma unroll
        for (int chunk = 0; chunk < NCH; ++chunk) {
          const uint32_t st = ct % kAcc;
          mbar_wait(t_full(st), (ct / kAcc) & 1);
          tc_fence_after();
#ifdef FGNN_TC_TRACE
          if (chunk == 0 && warp == 0) TC_TRACE(ei, 6);
#endif
          const uint32_t taddr = lane_addr + st * kAccCols;
          uint32_t d[2][16];
          tmem_ld16(taddr, d[0]);
#pragma unroll
          for (int gq = 0; gq < NC / 16; ++gq) {
            tmem_ld_wait();
            if (gq + 1 < NC / 16) tmem_ld16(taddr + (gq + 1) * 16, d[(gq + 1) & 1]);
#pragma unroll
            for (int q = 0; q < CH_PER_LD; ++q) {
              float e;
              if (T >= 4) {                                  // four independent chains: FMA latency, not count, binds here
                float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
#pragma unroll
                for (int t = 0; t < T; t += 4) {
                  e0 = fmaf(et[t], __uint_as_float(d[gq & 1][q * T + t]), e0);
                  e1 = fmaf(et[(t + 1) % T], __uint_as_float(d[gq & 1][q * T + (t + 1) % T]), e1);
                  e2 = fmaf(et[(t + 2) % T], __uint_as_float(d[gq & 1][q * T + (t + 2) % T]), e2);
                  e3 = fmaf(et[(t + 3) % T], __uint_as_float(d[gq & 1][q * T + (t + 3) % T]), e3);
                }
                e = (e0 + e1) + (e2 + e3);
              } else {
                e = 0.f;
#pragma unroll
                for (int t = 0; t < T; ++t) e = fmaf(et[t], __uint_as_float(d[gq & 1][q * T + t]), e);
              }
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
#ifdef FGNN_TC_TRACE
        if (warp == 0) TC_TRACE(ei, 7);
        ++ei;
#endif
        live_count += live ? 1.f : 0.f;
      }
      // finish: aggregate, bias / eval-BN / activation (mp_nn.py:162-173) into this warp's staging
      // rows (16-byte chunk c of row rr at position c ^ (rr & SW): conflict-free both ways) ...
      {
        constexpr int CPR = CH / 4;                          // 16-byte chunks per output row
        constexpr int SW = (CPR < 8 ? CPR : 8) - 1;          // chunk-index bits XOR-ed with the row
        float4* stage = reinterpret_cast<float4*>(sOut) + warp * (32 * CPR);   // this warp's [32 rows][CPR chunks]
        const float4* epi4 = reinterpret_cast<const float4*>(s_epi);
        const float inv_gamma = 1.f / p.gamma, inv_live = live_count > 0.f ? 1.f / live_count : 0.f;
        // negative-side slope of the activation: 1 = none, 0 = ReLU, slope = LeakyReLU
        const float neg = p.act == FGNN_ACT_NONE ? 1.f : (p.act == FGNN_ACT_RELU ? 0.f : p.slope);
        __syncwarp();                                        // previous tile's read-back is complete
#ifdef FGNN_TC_TRACE
        if (warp == 0) TC_TRACE(ei - 1, 8);
#endif
#pragma unroll
        for (int c4 = 0; c4 < CH; c4 += 4) {
          const float4 bi = epi4[c4 >> 2], sc = epi4[(CH + c4) >> 2], sh = epi4[(2 * CH + c4) >> 2];
          const float bia[4] = {bi.x, bi.y, bi.z, bi.w}, sca[4] = {sc.x, sc.y, sc.z, sc.w}, shi[4] = {sh.x, sh.y, sh.z, sh.w};
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = c4 + j;
            float a;
            if (AGG == FGNN_AGG_MAX) a = acc[c];
            else if (AGG == FGNN_AGG_SOFTMAX) a = live_count > 0.f ? (logf(acc2[c]) + acc[c]) * inv_gamma : -INFINITY;
            else a = acc[c] * inv_live;
            float y = fmaf(a + bia[j], sca[j], shi[j]);      // bias, then eval BN folded to scale/shift
            y = y >= 0.f ? y : y * neg;
            v[j] = a == -INFINITY ? a : y;                   // no live slot on this shard: stay -inf
          }
          stage[lane * CPR + (((c4 >> 2) & ~SW) | (((c4 >> 2) ^ lane) & SW))] = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncwarp();
#ifdef FGNN_TC_TRACE
        if (warp == 0) TC_TRACE(ei - 1, 9);
#endif
        // ... then whole 128-byte lines go out: CPR lanes per row, 32/CPR rows per store instruction
        // (out is batch-contiguous node-major, so flattened row g lives at out + g * o_sm)
        constexpr int RPI = 32 / CPR;
        const int cq = lane % CPR, rsub = lane / CPR;
        float* obase = p.out + ch0 + cq * 4;
#pragma unroll
        for (int it = 0; it < CPR; ++it) {
          const int rr = it * RPI + rsub;
          const int64_t orow = __shfl_sync(0xffffffffu, my_orow, rr);   // flattened output row (out is batch-contiguous)
          if (orow >= 0) {
            float4 v = stage[rr * CPR + ((cq & ~SW) | ((cq ^ rr) & SW))];
            float4* dst = reinterpret_cast<float4*>(obase + orow * p.o_sm);
            if (p.accumulate) {                              // out += v without waiting on a load: one vector reduction
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                           : "memory");
            } else {
              *dst = v;
            }
          }
        }
#ifdef FGNN_TC_TRACE
        if (warp == 0) TC_TRACE(ei - 1, 10);
#endif
      }
    }
  } else if (warp < kEpiWarps + kConvWarps) {
    // =====================================================================================
    // CONVERTERS: thread cr == TMEM lane cr == row cr of every item.  raw fp32 row (ring) ->
    // split bf16 -> A stage in tensor memory (32 columns of hi pairs, 32 columns of lo pairs)
    // =====================================================================================
    reg_dec<120>();
    const int cr = tid - kEpiWarps * 32;
    const uint32_t row_u = smem_u32(sA) + (uint32_t)cr * 256u;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kTACol0;
    uint32_t i = 0;
    for (ItemIter it = first_item(); it.tile < n_tiles; next_item(it), ++i) {
      const uint32_t st = i % NST, use = i / NST;
      const uint32_t ta = i % kTA, tuse = i / kTA;
      mbar_wait(raw_full(st), use & 1);
      if (cr < 32) TC_TRACE(i, 2);
      float4 v[16];                                          // channels 4c .. 4c+3 in chunk c (stored at c ^ (row & 15))
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = lds_f4(row_u + st * kAStageBytes + (uint32_t)((c ^ (cr & 15)) * 16));
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
      mbar_arrive(raw_empty(st));                            // ring stage is free again: every chunk has been consumed above
      mbar_wait(ta_empty(ta), (tuse & 1) ^ 1);
      tc_fence_after();
      tmem_st32(lane_addr + ta * kTACols, hi);
      tmem_st32(lane_addr + ta * kTACols + 32, lo);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ta_full(ta));
      if (cr < 32) TC_TRACE(i, 3);
    }
  } else if (warp < kMmaWarp) {
    // =====================================================================================
    // GATHERERS: cp.async the 128 source rows of every item into the raw ring
    // =====================================================================================
    reg_dec<56>();
    const int pw = warp - (kEpiWarps + kConvWarps);          // rows pw*32 .. pw*32+31 of the tile
    const int sub = lane >> 4, q = lane & 15;                // 16 lanes x 16 B = one 256-byte row
    // index-table entry of the row this lane owns (row pw*32+lane of item i's tile); the load is
    // issued one item ahead and only CONSUMED (range check -> source row) after the stage wait
    auto index_of = [&](const ItemIter& it, int32_t& base) -> int64_t {
      base = -1;
      if (it.tile >= n_tiles) return -1;
      const uint32_t g = (uint32_t)it.tile * kTileM + pw * 32 + lane;
      if (g >= rows_total) return -1;
      uint32_t b, m;
      split_row(g, b, m);
      base = (int32_t)(b * (uint32_t)p.N);
      return load_index(p.idx, p.idx64, (int64_t)b * p.idx_sb + (int64_t)m * p.K + it.k);
    };
    const float* xq = p.x + q * 4;
    const uint32_t sA_u = smem_u32(sA);
    int32_t base, base_next;
    ItemIter it = first_item();
    int64_t n = index_of(it, base);
    for (uint32_t i = 0; it.tile < n_tiles; ++i) {
      const uint32_t st = i % NST, use = i / NST;
      next_item(it);
      const int64_t n_next = index_of(it, base_next);        // index load of the next item: in flight during this one
      mbar_wait(raw_empty(st), (use & 1) ^ 1);
      if (pw == 0) TC_TRACE(i, 0);
      // source row (b*N + n; x is batch-contiguous, checked by tc_supported); -1 = no row (tile tail,
      // masked or out-of-range slot) -> zero-filled
      const int32_t src = (base >= 0 && n >= 0 && n < p.N) ? base + (int32_t)n : -1;
      const uint32_t stage = sA_u + st * kAStageBytes;
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int rw = 2 * u + sub;
        const int32_t row = __shfl_sync(0xffffffffu, src, rw);
        const int rr = pw * 32 + rw;
        // chunk q of row rr at position q ^ (rr & 15): the converter's per-row reads are conflict-free
        const uint32_t dst = stage + (uint32_t)rr * 256u + (uint32_t)((q ^ (rr & 15)) * 16);
        cp_async16(dst, xq + (int64_t)(row >= 0 ? row : 0) * kC, row >= 0 ? 16u : 0u);
      }
      cp_async_arrive_noinc(raw_full(st));
      if (pw == 0) TC_TRACE(i, 1);
      n = n_next;
      base = base_next;
    }
  } else {
    // =====================================================================================
    // MMA ISSUER: the whole warp runs the loop (warp-uniform control flow, so addresses and
    // descriptors live in uniform registers); one elected lane issues.  First: the stationary
    // B operand (this CTA's column slice of the split-bf16 filter image) by TMA bulk copy.
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
    // descriptor of a B tile = constant high word + (address >> 4) in the low word
    const uint64_t desc_hi = umma_desc_sw128(0);
    uint32_t ct = 0, i = 0;
    for (ItemIter it = first_item(); it.tile < n_tiles; next_item(it), ++i) {
      const uint32_t ta = i % kTA, tuse = i / kTA;
      mbar_wait(ta_full(ta), tuse & 1);
      TC_TRACE(i, 4);
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
            for (int ks = 0; ks < kC / 16; ++ks)             // 16 bf16 of K = 8 TMEM columns of A = 32 bytes of a B row
              umma_bf16_ts(d_tmem, a + ks * 8, desc_hi | (uint64_t)(bb + ks * 2), idesc, (term | ks) != 0);
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

TcConfig pick_config(int T, int agg, int OT) {
  // accumulator chunks of NC <= 128 columns (TMEM: 2 x 128 accumulator + 4 x 64 A-stage columns);
  // channels per CTA (NC*NCH/T) <= 64 for max/mean, <= 32 for softmax (two registers per channel)
  const bool sm = agg == FGNN_AGG_SOFTMAX;
  TcConfig c{T, 128, 1, false};
  switch (T) {
    case 16: c.NCH = sm ? 2 : 4; break;          // 16 / 32 channels
    case 8: c.NCH = 2; break;                    // 32 channels
    case 4: c.NCH = sm ? 1 : 2; break;           // 32 / 64 channels
    case 2: c.NCH = 1; c.NC = sm ? 64 : 128; break;
    case 1: c.NC = sm ? 32 : 64; break;
    default: return c;
  }
  c.ok = OT % (c.NC * c.NCH) == 0;
  return c;
}

size_t smem_bytes(const TcConfig& c) {
  const int cols = c.NC * c.NCH, ch = cols / c.T;
  return 1024 + (size_t)2 * cols * 128 + (size_t)tc::a_stages(cols, ch) * tc::kAStageBytes +
         (size_t)tc::kTileM * ch * 4 + tc::kNumBars * 8 + 16 + 3 * 64 * 4;
}

template <int T, int NC, int NCH, int AGG>
int launch_one(const MpParams& p, const uint8_t* wimg, int S, int workers, int tiles, size_t smem, cudaStream_t st) {
  auto kern = mp_tc_kernel<T, NC, NCH, AGG>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return FGNN_ERR_CUDA;
  kern<<<S * workers, tc::kThreads, smem, st>>>(p, wimg, S, workers, tiles);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

// NCm/NCHm: max & mean configuration, NCs/NCHs: softmax configuration (must mirror pick_config)
template <int T, int NCm, int NCHm, int NCs, int NCHs>
int launch_agg(const MpParams& p, const uint8_t* wimg, int S, int workers, int tiles, size_t smem, cudaStream_t st) {
  switch (p.agg) {
    case FGNN_AGG_MAX: return launch_one<T, NCm, NCHm, FGNN_AGG_MAX>(p, wimg, S, workers, tiles, smem, st);
    case FGNN_AGG_SOFTMAX: return launch_one<T, NCs, NCHs, FGNN_AGG_SOFTMAX>(p, wimg, S, workers, tiles, smem, st);
    case FGNN_AGG_MEAN: return launch_one<T, NCm, NCHm, FGNN_AGG_MEAN>(p, wimg, S, workers, tiles, smem, st);
  }
  return FGNN_ERR_UNSUPPORTED;
}

}  // namespace

#ifdef FGNN_TC_TRACE
extern "C" int fgnn_debug_trace_read(unsigned long long* host, size_t count) {
  return cudaMemcpyFromSymbol(host, g_trace, count * sizeof(unsigned long long)) == cudaSuccess ? 0 : -6;
}
#endif

bool tc_supported(const fgnn_mp_args* a) {
  if (a->extension != FGNN_NO_EXTENSION || a->dtype != FGNN_F32) return false;
  if (a->C != tc::kC) return false;
  if (a->aggregator == FGNN_AGG_NONE) return false;
  if (a->x_sc != 1 || a->x_sn != a->C) return false;                              // node-major rows
  if (a->B > 1 && a->x_sb != (int64_t)a->N * a->C) return false;                  // batch-contiguous
  if ((int64_t)a->B * a->N >= INT32_MAX) return false;
  if ((reinterpret_cast<uintptr_t>(a->x) & 15) || (reinterpret_cast<uintptr_t>(a->out) & 15)) return false;
  if (a->out_so != 1 || (a->out_sm & 3)) return false;
  if (a->B > 1 && a->out_sb != (int64_t)a->M * a->out_sm) return false;             // batch-contiguous rows
  if (a->O % 4) return false;
  const TcConfig c = pick_config(a->T, a->aggregator, a->O * a->T);
  if (!c.ok) return false;
  if (smem_bytes(c) > (size_t)tc::kSmemBudget || tc::a_stages(c.NC * c.NCH, c.NC * c.NCH / c.T) < 2) return false;
  if ((int64_t)a->B * a->M >= (int64_t)INT32_MAX - 256) return false;
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
  // The image is rebuilt unless this workspace is known to hold the image of exactly these filters
  // (pointer + caller-supplied version).  Host-side mirror of the device header: a cached image costs
  // no launch at all.  filters_version == 0 always rebuilds.
  {
    static std::mutex mu;
    static std::unordered_map<const void*, tc::Header> known;
    bool fresh = false;
    if (a->filters_version != 0) {
      std::lock_guard<std::mutex> lock(mu);
      auto it = known.find(ws);
      fresh = it != known.end() && it->second.version == a->filters_version && it->second.filters == p.W &&
              it->second.OT == OT;
      if (!fresh) {
        if (known.size() > 4096) known.clear();
        known[ws] = tc::Header{a->filters_version, p.W, tc::kC, OT};
      }
    } else {
      std::lock_guard<std::mutex> lock(mu);
      known.erase(ws);
    }
    if (!fresh) {
      w_split_kernel<<<(OT * (tc::kC / 8) + 255) / 256, 256, 0, stream>>>(p.W, ws, OT, 0);
      count_launch();
      if (cudaGetLastError() != cudaSuccess) return FGNN_ERR_CUDA;
    }
  }

  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int64_t rows = (int64_t)p.B * p.M;
  const int tiles = (int)((rows + tc::kTileM - 1) / tc::kTileM);
  const int S = OT / (c.NC * c.NCH);
  int sms = num_sms;
  if (a->sm_limit > 0 && a->sm_limit < sms) sms = a->sm_limit;
  if (S > sms) return FGNN_ERR_UNSUPPORTED;
  int workers = sms / S;
  if (workers > tiles) workers = tiles;
  const size_t smem = smem_bytes(c);
  switch (p.T) {
    case 16: return launch_agg<16, 128, 4, 128, 2>(p, ws, S, workers, tiles, smem, stream);
    case 8: return launch_agg<8, 128, 2, 128, 2>(p, ws, S, workers, tiles, smem, stream);
    case 4: return launch_agg<4, 128, 2, 128, 1>(p, ws, S, workers, tiles, smem, stream);
    case 2: return launch_agg<2, 128, 1, 64, 1>(p, ws, S, workers, tiles, smem, stream);
    case 1: return launch_agg<1, 64, 1, 32, 1>(p, ws, S, workers, tiles, smem, stream);
  }
  return FGNN_ERR_UNSUPPORTED;
}

}  // namespace fgnn

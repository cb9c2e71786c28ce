// Shared pieces of the tcgen05 kernels (mp_tc.cu: destination-stationary, mp_src.cu: source-stationary):
// CTA geometry, shared-memory budget helpers and the PTX wrappers (mbarrier, cp.async, TMA bulk copy,
// tcgen05 alloc / ld / st / mma / commit, setmaxnreg, programmatic dependent launch).
#pragma once
#include <cuda_bf16.h>

#include <cstdio>

#include "common.cuh"

namespace fgnn {

#if defined(FGNN_TC_TRACE) && defined(FGNN_TC_TRACE_TU)
// Debug builds only (-DFGNN_TC_TRACE), in the one translation unit that defines FGNN_TC_TRACE_TU (mp_tc.cu):
// per-item timestamps of CTA 0, read back by tools/tc_trace.py.
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
constexpr int kEpiGroups = 2;          // epilogue warp groups sharing the TMEM lanes, splitting the columns
constexpr int kEpiWarps = 4 * kEpiGroups, kConvWarps = 4, kGatherWarps = 2;
constexpr int kConvWarp0 = kEpiWarps, kGatherWarp0 = kConvWarp0 + kConvWarps;
constexpr int kMmaWarp = kGatherWarp0 + kGatherWarps;
constexpr int kThreads = 512;                                 // 16 warps: the last one only completes the fourth warpgroup
static_assert(kMmaWarp == 14, "warps 12-15 must form one warpgroup");
constexpr int kMaxAStages = 6;                                // raw ring stages in shared memory
constexpr int kTA = 2;                                        // A stages in tensor memory
constexpr int kAcc = 3;                                       // accumulator stages in tensor memory
constexpr int kAccCols = 128, kTACol0 = kAcc * kAccCols, kTACols = 64;   // TMEM column map: 3 x 128 + 2 x 64 = 512
constexpr int kNumBars = 2 * kMaxAStages + 2 * kTA + 2 * kAcc + 1;    // raw_full, raw_empty, ta_full, ta_empty, t_full, t_empty, w_full
constexpr int kSmemBudget = 227 * 1024;
constexpr int kHeaderBytes = 256;                             // workspace header in front of the W image
constexpr uint32_t kSpinLimit = 1u << 22;                     // watchdog: trap instead of hanging the GPU

struct Header {                       // first bytes of the workspace
  int64_t version;                    // fgnn_mp_args.filters_version the image was built from
  const float* filters;               // and the pointer it was built from
  int32_t C, OT;
};

// shared-memory geometry, by I/O type (xb: bf16 I/O)
// ka = K atoms: the input channels in units of kC (one 128-byte swizzle atom of bf16): 1 (C = 64) or 2 (C = 128)
__host__ __device__ constexpr int row_bytes(bool xb, int ka = 1) { return ka * (xb ? kC * 2 : kC * 4); }   // one source row in the ring
__host__ __device__ constexpr int stage_bytes(bool xb, int ka = 1) { return kTileM * row_bytes(xb, ka); }  // 16 / 32 / 64 KB
__host__ __device__ constexpr int w_bytes(int cols, bool xb, int ka = 1) { return (xb ? 1 : 2) * ka * cols * 128; }   // hi (+ lo) image rows
__host__ __device__ constexpr int out_tile_bytes(int ch, bool xb) { return kTileM * ch * (xb ? 2 : 4); }
// raw-ring stages that fit beside a filter slice of `cols` columns and the output staging tile of `ch`
// channels (plus 1 KB alignment slack, barriers, epilogue params)
__host__ __device__ constexpr int a_stages(int cols, int ch, bool xb, int ka = 1) {
  int n = (kSmemBudget - 1024 - w_bytes(cols, xb, ka) - out_tile_bytes(ch, xb) - 2048) / stage_bytes(xb, ka);
  return n > kMaxAStages ? kMaxAStages : n;
}

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
    if (++spins > kSpinLimit) {
#ifdef FGNN_TC_DEBUG
      printf("fgnn mbar timeout: block %d thread %d barrier+%u parity %u\n", (int)blockIdx.x, (int)threadIdx.x,
             bar & 0xfffu, parity);
#endif
      __trap();
    }
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
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 32 consecutive columns of this thread's lane in one instruction
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 64 consecutive columns of this thread's lane in one instruction
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
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
// Packed fp32x2 arithmetic (FFMA2 / FADD2, sm_100): two independent fp32 operations per instruction, each
// lane rounded like the scalar instruction.  The 3-register scalar FFMA issues every other cycle per SM
// sub-partition; the edge-type contraction of the epilogues is FMA-issue bound, so it runs packed.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
  return v;
}
__device__ __forceinline__ uint64_t pack2u(uint32_t lo, uint32_t hi) {
  uint64_t v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
  return v;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// sum_t et[t] * h[t] over T (a multiple of 4) accumulator columns h: lanes (t, t+1) and (t+2, t+3) of every
// group of four run as two packed chains, folded as ((c0 + c2) + (c1 + c3)).  Both tcgen05 kernels use this
// one function, which is what makes their results bit-identical.
template <int T>
__device__ __forceinline__ float contract_types(const float (&et)[T], const uint32_t* h) {
  static_assert(T % 4 == 0, "packed contraction needs a multiple of four edge types");
  uint64_t s01 = pack2(0.f, 0.f), s23 = pack2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < T; t += 4) {
    s01 = ffma2(pack2(et[t], et[t + 1]), pack2u(h[t], h[t + 1]), s01);
    s23 = ffma2(pack2(et[t + 2], et[t + 3]), pack2u(h[t + 2], h[t + 3]), s23);
  }
  const uint64_t s = fadd2(s01, s23);
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(s));
  return lo + hi;
}
// T = 4: the filter image keeps the columns of a channel PAIR interleaved -- image column (o >> 1) * 8 + 2 t + (o & 1)
// for filter column o * 4 + t (w_split_kernel) -- so an aligned 64-bit register pair of an accumulator row is
// (H[2g][t], H[2g+1][t]) and ONE packed FMA per edge type updates both channels: four FFMA2 for two channels and no
// horizontal add (the T >= 8 scheme above spends half of its instructions on T = 4 outside the multiply-adds).
// et2[t] = (et[t], et[t]).  Returns (channel 2g, channel 2g+1) = sum_t et[t] * H[.][t], accumulated in order t = 0..3.
__device__ __forceinline__ uint64_t contract_pair4(const uint64_t (&et2)[4], const uint32_t* h8) {
  uint64_t s = ffma2(et2[0], pack2u(h8[0], h8[1]), pack2(0.f, 0.f));
  s = ffma2(et2[1], pack2u(h8[2], h8[3]), s);
  s = ffma2(et2[2], pack2u(h8[4], h8[5]), s);
  s = ffma2(et2[3], pack2u(h8[6], h8[7]), s);
  return s;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// image column of filter column n = o * T + t
__host__ __device__ __forceinline__ int image_column(int n, int T) {
  return T == 4 ? ((n >> 3) << 3) + 2 * (n & 3) + ((n >> 2) & 1) : n;
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// named barrier over `count` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// programmatic dependent launch (no-ops when the launch carries no programmatic dependency)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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

// Register re-balancing between the warp roles.  setmaxnreg is executed by whole WARPGROUPS (four
// consecutive warps, all with the same value): warpgroups 0-1 = epilogue, 2 = converters, 3 = gatherers +
// MMA warp + one idle warp.  With setmaxnreg in the code ptxas gives the kernel
// floor(65536 / threads / 32) * 32 registers per thread at launch (probed: 128 for 416..512 threads, 96 for
// 544).  The instruction moves registers WITHIN the CTA's launch allocation (512 x 128 = 65536), so
// 256*160 (epilogue) + 128*120 (converters) + 128*72 (warpgroup 3) = 65536 must fit in it -- an
// over-subscribed inc never returns; the host checks the launch register count of every instantiation
// before its first launch.
constexpr int kRegEpi = 160, kRegConv = 120, kRegAux = 72, kRegLaunch = 128;
static_assert(kEpiWarps * 32 * kRegEpi + kConvWarps * 32 * kRegConv + 4 * 32 * kRegAux <= kThreads * kRegLaunch,
              "setmaxnreg budget exceeds the launch allocation");
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

}  // namespace tc

}  // namespace fgnn

// Cross-GPU step of the factor-sharded FGNN layer as ONE kernel over NVLink peer memory.
//
// Factor-sharded F->V: every rank holds the raw per-type maxima of ITS factors in a
// [rows, J*O] buffer (-inf where it has no incident factor).  FactorNN then needs
//     x_v'[n, o] = sum_j act(BN_j(max_ranks raw_r[n, j*O + o] + bias_j[o]))          (factor_mpnn_sp.py:142-147)
// on every rank.  With NCCL that is all_reduce(MAX) over [rows, J*O] + an epilogue kernel; here the
// owner of a row range reads the peers' raw rows straight out of their memory (ld over NVLink), reduces,
// applies the (non-linear, per-type) epilogue, sums the types and stores the finished O-wide row into
// EVERY rank's next-layer feature buffer: (G-1)/G * rows * (J*O + O) * 4 bytes cross the links per rank
// instead of 2 * (G-1)/G * rows * J*O * 4, and the epilogue rides along.
//
// Synchronisation is by epoch flags in the peers' arenas (system-scope release / acquire):
//   A[rank] = epoch   "my raw buffer of this epoch is complete"   (set by block 0 at kernel start: the
//                      producing kernels precede this one in stream order)
//   B[rank] = epoch   "I have read every raw buffer and written my rows everywhere"  (set by the last block)
// Every block waits for all A before reading and for all B before exiting, so when the kernel completes on
// a rank (a) all rows of its own next-layer buffer are in place and (b) nobody still reads its raw buffer.
// All ranks must launch the kernel with the same epoch.  Every block spins on flags other blocks (here and on the peers)
// set, so ALL blocks of the grid must be resident at once: the grid is small (<= 128 CTAs of 512 threads, two per SM;
// launch_exchange refuses a grid beyond what the occupancy calculator says the device can hold) and the caller gives
// the persistent tensor-core kernels that run beside it sm_limit = SMs - ceil(ctas / 2), so its SMs are free.
#include <cstring>

#include "common.cuh"

namespace fgnn {

namespace {

constexpr int kMaxRanks = 8;
constexpr uint32_t kSpinLimit = 1u << 27;        // ~10 s: trap instead of hanging the GPU when a peer never arrives

struct ExParams {
  const float* raw[kMaxRanks];
  float* out[kMaxRanks];
  uint32_t* flags[kMaxRanks];                    // [2][kMaxRanks] per arena: A then B
  unsigned long long* counter;                   // this rank's block counter (monotonic)
  const float* bias; const float* scale; const float* shift;     // [J*O] or nullptr
  const uint32_t* raw_mask;                      // [rows] bit q*J+j: rank q's raw row has data of type j (nullptr = all)
  const uint32_t* out_mask;                      // [rows] bit q: rank q needs the finished row (nullptr = all)
  int64_t rows, row0, row1;
  int world, rank, J, O, act;
  uint32_t epoch;
  float slope;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {            // peer memory: never from a stale local cache line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// the same load under a predicate, -inf when it is off: straight-line code, so a thread's loads stay batched
__device__ __forceinline__ float4 ld_peer_f4_if(const float* p, bool on) {
  float4 v;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.u32 q, %5, 0;\n\t"
      "mov.b32 %0, 0xff800000;\n\tmov.b32 %1, 0xff800000;\n\tmov.b32 %2, 0xff800000;\n\tmov.b32 %3, 0xff800000;\n\t"
      "@q ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=&f"(v.x), "=&f"(v.y), "=&f"(v.z), "=&f"(v.w)
      : "l"(p), "r"((uint32_t)on)
      : "memory");
  return v;
}

__device__ __forceinline__ void wait_all(const ExParams& p, int phase, uint32_t epoch) {
  if ((int)threadIdx.x < p.world) {
    const uint32_t* f = p.flags[p.rank] + phase * kMaxRanks + threadIdx.x;
    uint32_t spins = 0;
    // epochs only grow; a peer may already be an epoch ahead in phase A of the next layer
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
      if (++spins > kSpinLimit) __trap();
      __nanosleep(64);
    }
  }
  __syncthreads();
}

// WORLD: ranks (compile time, so the peer loads of one step sit in registers); UNROLL: row groups per thread
// step.  Everything a thread needs from the peers in one step -- UNROLL x J x (WORLD-1) 16-byte loads -- is
// issued before the first use: an NVLink round trip is ~2-3 us, so bandwidth is bytes in flight / latency.
template <int WORLD, int UNROLL, int kJ>            // kJ: factor types per load batch (J is 1-2 in FGNN; larger J loops)
__global__ void __launch_bounds__(512, 2)
exchange_kernel(const ExParams p) {
  // The epoch of this launch.  epoch == 0 in the arguments: derive it from this rank's block counter -- every launch
  // adds gridDim.x to it, and it cannot reach the next multiple before the block reading it here has finished -- so a
  // launch captured in a CUDA graph (frozen arguments) still counts up on every replay, in step on all ranks.
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0)
    s_epoch = p.epoch ? p.epoch : (uint32_t)(*reinterpret_cast<volatile unsigned long long*>(p.counter) / gridDim.x) + 1u;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  // A: my raw buffer is complete (stream order) -> tell every rank, once
  if (blockIdx.x == 0 && (int)threadIdx.x < WORLD)
    st_release_sys(p.flags[threadIdx.x] + 0 * kMaxRanks + p.rank, epoch);
  wait_all(p, 0, epoch);

  const int O4 = p.O >> 2, JO = p.J * p.O;
  const int64_t total = (p.row1 - p.row0) * O4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float neg = p.act == FGNN_ACT_NONE ? 1.f : (p.act == FGNN_ACT_RELU ? 0.f : p.slope);
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * UNROLL) {
    float acc[UNROLL][4];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f;
    for (int j0 = 0; j0 < p.J; j0 += kJ) {
      float4 mine[UNROLL][kJ], peer[UNROLL][kJ][WORLD > 1 ? WORLD - 1 : 1];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t i = i0 + u * stride;
        const bool ok = i < total;
        const int64_t r = p.row0 + (ok ? i : 0) / O4;
        const int o = (int)((ok ? i : 0) % O4) * 4;
        // rows a rank's factor shard does not touch hold -inf there: skip the trip over the link
        const uint32_t have = !ok ? 0u : (p.raw_mask ? __ldg(p.raw_mask + r) : 0xffffffffu);
#pragma unroll
        for (int jj = 0; jj < kJ; ++jj) {
          const bool okj = ok && j0 + jj < p.J;
          const int64_t off = r * JO + (okj ? j0 + jj : 0) * p.O + o;
          // my own rows: ordinary load (with a mask, rows my shard does not touch were never written: -inf)
          mine[u][jj] = (okj && ((have >> (p.rank * p.J + j0 + jj)) & 1u)) ? *reinterpret_cast<const float4*>(p.raw[p.rank] + off)
                                                                        : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
          for (int q = 0; q < WORLD - 1; ++q) {
            const int pr = q < p.rank ? q : q + 1;                                   // the q-th OTHER rank
            peer[u][jj][q] = ld_peer_f4_if(p.raw[pr] + off, okj && ((have >> (pr * p.J + j0 + jj)) & 1u));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t i = i0 + u * stride;
        const int o = (int)((i < total ? i : 0) % O4) * 4;
#pragma unroll
        for (int jj = 0; jj < kJ; ++jj) {
          if (j0 + jj >= p.J) continue;
          float4 m = mine[u][jj];
#pragma unroll
          for (int q = 0; q < WORLD - 1; ++q) {
            const float4 v = peer[u][jj][q];
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
          }
          const float mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float y = mv[c];
            if (y != -INFINITY) {                            // same arithmetic as epilogue_sum_kernel (api.cu)
              const int col = (j0 + jj) * p.O + o + c;
              if (p.bias) y += p.bias[col];
              if (p.scale) y = fmaf(y, p.scale[col], p.shift[col]);
              if (p.act == FGNN_ACT_RELU) y = fmaxf(y, 0.f);
              else if (p.act == FGNN_ACT_LEAKY_RELU) y = y >= 0.f ? y : y * neg;
            }
            acc[u][c] += y;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= total) continue;
      const int64_t r = p.row0 + i / O4;
      const int o = (int)(i % O4) * 4;
      const float4 res = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
      const uint32_t want = p.out_mask ? (__ldg(p.out_mask + r) | (1u << p.rank)) : 0xffffffffu;
#pragma unroll
      for (int q = 0; q < WORLD; ++q)
        if ((want >> q) & 1u) *reinterpret_cast<float4*>(p.out[q] + r * p.O + o) = res;
    }
  }

  // B: all of this rank's blocks are done -> the last one tells every rank
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(p.counter, 1ull) + 1ull;
    if (done % gridDim.x == 0) {                              // every launch adds gridDim.x: the last block of this launch
      __threadfence_system();
      for (int q = 0; q < WORLD; ++q) st_release_sys(p.flags[q] + 1 * kMaxRanks + p.rank, epoch);
    }
  }
  wait_all(p, 1, epoch);
}

// ---------------------------------------------------------------------------------------------
// Feature-halo pull for the owner-computes sharding (parallel.py: HaloLayerPlan).  Every rank owns a contiguous
// range of variables and of each type's factors and evaluates ALL slots of its own destinations (same kernel, same
// slot order as the single-GPU layer: bit-identical rows); what it lacks are the features of the source rows other
// ranks own -- its HALO.  After a layer's calls have written the owned rows of the next-layer buffers, this kernel
// copies every halo row straight out of its owner's buffer over NVLink (16-byte loads, all of a thread's loads in
// flight before the first store), for up to kMaxJobs buffers (variables + factor types) in one launch.
// Epoch flags as above: A = "my owned rows are written" (stream order: the producers precede this launch),
// B = "I have copied everything I need": when the kernel completes on a rank nobody still reads its buffers.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxJobs = 4;

struct HaloJob {
  const uint8_t* src[kMaxRanks];     // per rank: base of that rank's buffer (owned rows first)
  uint8_t* dst;                      // this rank's buffer
  const int32_t* src_row;            // [n] row in the owner's buffer
  const uint8_t* src_rank;           // [n] owner
  int64_t dst_row0;                  // first halo row of this rank's buffer
  int32_t n;                         // halo rows
  int32_t row_bytes;                 // multiple of 16
};

struct HaloParams {
  HaloJob job[kMaxJobs];
  uint32_t* flags[kMaxRanks];
  unsigned long long* counter;
  int n_jobs, world, rank;
  uint32_t epoch;
};

__global__ void __launch_bounds__(512, 2)
halo_kernel(const HaloParams p) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0)
    s_epoch = p.epoch ? p.epoch : (uint32_t)(*reinterpret_cast<volatile unsigned long long*>(p.counter) / gridDim.x) + 1u;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  ExParams fl;                                         // wait_all only reads flags / rank / world
  for (int q = 0; q < kMaxRanks; ++q) fl.flags[q] = p.flags[q];
  fl.rank = p.rank; fl.world = p.world;
  if (blockIdx.x == 0 && (int)threadIdx.x < p.world) st_release_sys(p.flags[threadIdx.x] + 0 * kMaxRanks + p.rank, epoch);
  wait_all(fl, 0, epoch);
  constexpr int U = 4;
  for (int j = 0; j < p.n_jobs; ++j) {
    const HaloJob& jb = p.job[j];
    const int cpr = jb.row_bytes >> 4;                 // 16-byte chunks per row
    const int64_t total = (int64_t)jb.n * cpr, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < total) {
          const int64_t r = i / cpr;
          const int c = (int)(i - r * cpr);
          const uint8_t* src = jb.src[jb.src_rank[r]] + (int64_t)jb.src_row[r] * jb.row_bytes + c * 16;
          v[u] = ld_peer_f4(reinterpret_cast<const float*>(src));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < total) {
          const int64_t r = i / cpr;
          const int c = (int)(i - r * cpr);
          *reinterpret_cast<float4*>(jb.dst + (jb.dst_row0 + r) * jb.row_bytes + c * 16) = v[u];
        }
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(p.counter, 1ull) + 1ull;
    if (done % gridDim.x == 0) {
      __threadfence_system();
      for (int q = 0; q < p.world; ++q) st_release_sys(p.flags[q] + 1 * kMaxRanks + p.rank, epoch);
    }
  }
  wait_all(fl, 1, epoch);
}

template <int WORLD, int UNROLL, int kJ>
cudaError_t launch_exchange(const ExParams& p, int ctas, cudaStream_t st) {
  // co-residency of the whole grid is a correctness condition (see the header comment): check it against the device
  static int capacity = 0;
  if (!capacity) {
    int per_sm = 0, dev = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, exchange_kernel<WORLD, UNROLL, kJ>, 512, 0) != cudaSuccess ||
        cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return cudaErrorUnknown;
    capacity = per_sm * sms;
  }
  if (ctas > capacity) return cudaErrorLaunchOutOfResources;
  exchange_kernel<WORLD, UNROLL, kJ><<<ctas, 512, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

}  // namespace fgnn

using namespace fgnn;

extern "C" {

int fgnn_comm_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
  if (!dev_ptr || !handle64 || bytes == 0) return FGNN_ERR_INVALID_ARG;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return FGNN_ERR_CUDA;
  if (cudaMemset(p, 0, bytes) != cudaSuccess) { cudaFree(p); return FGNN_ERR_CUDA; }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); return FGNN_ERR_CUDA; }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(p); return FGNN_ERR_CUDA; }
  *dev_ptr = p;
  return FGNN_OK;
}

int fgnn_comm_open(const unsigned char* handle64, void** dev_ptr) {
  if (!handle64 || !dev_ptr) return FGNN_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return FGNN_ERR_CUDA;
  *dev_ptr = p;
  return FGNN_OK;
}

int fgnn_comm_close(void* dev_ptr) {
  return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? FGNN_OK : FGNN_ERR_CUDA;
}

int fgnn_comm_free(void* dev_ptr) { return cudaFree(dev_ptr) == cudaSuccess ? FGNN_OK : FGNN_ERR_CUDA; }

int fgnn_halo_pull(const fgnn_halo_args* a, void* stream_) {
  if (!a || a->world < 1 || a->world > kMaxRanks || a->rank < 0 || a->rank >= a->world || a->n_jobs < 1 || a->n_jobs > kMaxJobs ||
      !a->counter)
    return FGNN_ERR_INVALID_ARG;
  HaloParams p;
  memset(&p, 0, sizeof(p));
  for (int q = 0; q < a->world; ++q) {
    if (!a->flags[q]) return FGNN_ERR_INVALID_ARG;
    p.flags[q] = a->flags[q];
  }
  for (int j = 0; j < a->n_jobs; ++j) {
    const fgnn_halo_job& s = a->jobs[j];
    if (s.n < 0 || s.row_bytes <= 0 || (s.row_bytes & 15) || !s.dst || (s.n > 0 && (!s.src_row || !s.src_rank))) return FGNN_ERR_INVALID_ARG;
    HaloJob& d = p.job[j];
    for (int q = 0; q < a->world; ++q) d.src[q] = reinterpret_cast<const uint8_t*>(s.src[q]);
    d.dst = reinterpret_cast<uint8_t*>(s.dst); d.src_row = s.src_row; d.src_rank = s.src_rank;
    d.dst_row0 = s.dst_row0; d.n = s.n; d.row_bytes = s.row_bytes;
  }
  p.counter = reinterpret_cast<unsigned long long*>(a->counter);
  p.n_jobs = a->n_jobs; p.world = a->world; p.rank = a->rank; p.epoch = a->epoch;
  int ctas = a->ctas > 0 ? a->ctas : 32;
  if (ctas > 128) ctas = 128;
  halo_kernel<<<ctas, 512, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(p);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

int fgnn_exchange_forward(const fgnn_exchange_args* a, void* stream_) {
  if (!a || a->world < 1 || a->world > kMaxRanks || a->rank < 0 || a->rank >= a->world) return FGNN_ERR_INVALID_ARG;
  if (a->rows <= 0 || a->J <= 0 || a->O <= 0 || (a->O & 3) || a->row0 < 0 || a->row1 < a->row0 || a->row1 > a->rows)
    return FGNN_ERR_INVALID_ARG;
  if (!a->counter) return FGNN_ERR_INVALID_ARG;
  if ((a->bn_scale == nullptr) != (a->bn_shift == nullptr)) return FGNN_ERR_INVALID_ARG;
  ExParams p;
  for (int q = 0; q < kMaxRanks; ++q) {
    const bool live = q < a->world;
    if (live && (!a->raw[q] || !a->out[q] || !a->flags[q])) return FGNN_ERR_INVALID_ARG;
    p.raw[q] = live ? a->raw[q] : nullptr;
    p.out[q] = live ? a->out[q] : nullptr;
    p.flags[q] = live ? a->flags[q] : nullptr;
  }
  p.counter = reinterpret_cast<unsigned long long*>(a->counter);
  p.bias = a->bias; p.scale = a->bn_scale; p.shift = a->bn_shift;
  if (a->raw_mask && a->world * a->J > 32) return FGNN_ERR_INVALID_ARG;
  p.raw_mask = a->raw_mask; p.out_mask = a->out_mask;
  p.rows = a->rows; p.row0 = a->row0; p.row1 = a->row1;
  p.world = a->world; p.rank = a->rank; p.J = a->J; p.O = a->O; p.act = a->activation;
  p.epoch = a->epoch; p.slope = a->act_slope;
  int ctas = a->ctas > 0 ? a->ctas : 64;           // two 512-thread CTAs fit an SM: 64 CTAs = 32 SMs
  if (ctas > 128) ctas = 128;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  cudaError_t e;
  switch (a->world) {
    case 1: e = launch_exchange<1, 2, 2>(p, ctas, st); break;      // loads in flight per thread: UNROLL * kJ * WORLD x 16 B
    case 2: e = launch_exchange<2, 2, 2>(p, ctas, st); break;
    case 3: e = launch_exchange<3, 1, 2>(p, ctas, st); break;
    case 4: e = launch_exchange<4, 1, 2>(p, ctas, st); break;
    case 5: e = launch_exchange<5, 1, 1>(p, ctas, st); break;
    case 6: e = launch_exchange<6, 1, 1>(p, ctas, st); break;
    case 7: e = launch_exchange<7, 1, 1>(p, ctas, st); break;
    default: e = launch_exchange<8, 1, 1>(p, ctas, st); break;
  }
  count_launch();
  return e == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

}  // extern "C"

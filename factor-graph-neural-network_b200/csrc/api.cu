// C ABI of libfgnn_b200.so (see include/fgnn_b200.h).  Validation, kernel selection, the small
// helper kernels (layout change, index check, epilogue) and the host-buffer entry point.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace fgnn {

static std::atomic<uint64_t> g_launches{0};
static thread_local int g_last_cuda_error = 0;

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return FGNN_ERR_CUDA;
}

#define FGNN_CUDA(call)                              \
  do {                                               \
    cudaError_t _e = (call);                         \
    if (_e != cudaSuccess) return fgnn::cuda_fail(_e); \
  } while (0)

static int validate(const fgnn_mp_args* a) {
  if (!a || !a->x || !a->idx || !a->etype || !a->filters || !a->out) return FGNN_ERR_INVALID_ARG;
  if (a->B <= 0 || a->N <= 0 || a->M <= 0 || a->K <= 0 || a->C <= 0 || a->O <= 0 || a->T <= 0)
    return FGNN_ERR_INVALID_ARG;
  if (a->extension < 0 || a->extension > 2) return FGNN_ERR_INVALID_ARG;   // mp_nn.py:47-48 ValueError
  if (a->aggregator < 0 || a->aggregator > 3) return FGNN_ERR_INVALID_ARG;
  if (a->activation < 0 || a->activation > 2) return FGNN_ERR_INVALID_ARG;
  if (a->idx_dtype != FGNN_I64 && a->idx_dtype != FGNN_I32) return FGNN_ERR_INVALID_ARG;
  if (a->dtype != FGNN_F32 && a->dtype != FGNN_BF16) return FGNN_ERR_INVALID_ARG;
  if ((a->bn_scale == nullptr) != (a->bn_shift == nullptr)) return FGNN_ERR_INVALID_ARG;
  if (a->extension != FGNN_NO_EXTENSION && a->M != a->N) return FGNN_ERR_SHAPE;  // mp_nn.py:136-159
  if ((a->flags & FGNN_FLAG_ACCUMULATE) && a->aggregator == FGNN_AGG_NONE) return FGNN_ERR_INVALID_ARG;
  if ((a->tile_slots || a->out_rows) && a->aggregator == FGNN_AGG_NONE) return FGNN_ERR_INVALID_ARG;
  return FGNN_OK;
}

static MpParams to_params(const fgnn_mp_args* a) {
  MpParams p;
  p.x = reinterpret_cast<const float*>(a->x);
  p.idx = a->idx;
  p.et = reinterpret_cast<const float*>(a->etype);
  p.W = a->filters;
  p.bias = a->bias;
  p.scale = a->bn_scale;
  p.shift = a->bn_shift;
  p.out = reinterpret_cast<float*>(a->out);
  p.tile_k = a->tile_slots;
  p.out_rows = a->out_rows;
  p.x_sb = a->x_sb; p.x_sc = a->x_sc; p.x_sn = a->x_sn;
  p.idx_sb = a->idx_sb; p.et_sb = a->et_sb;
  p.o_sb = a->out_sb; p.o_so = a->out_so; p.o_sm = a->out_sm; p.o_sk = a->out_sk;
  p.B = a->B; p.N = a->N; p.M = a->M; p.K = a->K; p.C = a->C; p.O = a->O; p.T = a->T;
  p.ext = a->extension; p.agg = a->aggregator; p.act = a->activation;
  p.idx64 = a->idx_dtype == FGNN_I64;
  p.mask_neg = (a->flags & FGNN_FLAG_MASK_NEGATIVE) ? 1 : 0;
  p.accumulate = (a->flags & FGNN_FLAG_ACCUMULATE) ? 1 : 0;
  p.gamma = a->gamma; p.slope = a->act_slope;
  return p;
}

// ---------------------------------------------------------------------------------------------
// helper kernels
// ---------------------------------------------------------------------------------------------

// [B,C,N] (any strides) -> node-major [B,N,C]; 32x32 shared-memory tile transpose.
__global__ void to_node_major_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int N,
                                     int64_t x_sb, int64_t x_sc, int64_t x_sn) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, n = n0 + threadIdx.x;
    if (c < C && n < N) tile[j][threadIdx.x] = x[(int64_t)b * x_sb + (int64_t)c * x_sc + (int64_t)n * x_sn];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int n = n0 + j, c = c0 + threadIdx.x;
    if (c < C && n < N) out[((int64_t)b * N + n) * C + c] = tile[threadIdx.x][j];
  }
}

__global__ void check_index_kernel(const void* idx, int idx64, int64_t count, int64_t lo, int64_t n,
                                   int* flag) {
  int bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = load_index(idx, idx64, i);
    bad |= (v < lo) | (v >= n);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) *reinterpret_cast<volatile int*>(flag) = 1;   // idempotent store: works on mapped host memory too
}

__global__ void epilogue_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t total, int O,
                                const float* __restrict__ bias, const float* __restrict__ scale,
                                const float* __restrict__ shift, int act, float slope) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % O);
    float v = in[i];
    if (v != -INFINITY) {
      if (bias) v += bias[o];
      if (scale) v = fmaf(v, scale[o], shift[o]);
      if (act == FGNN_ACT_RELU) v = fmaxf(v, 0.f);
      else if (act == FGNN_ACT_LEAKY_RELU) v = v >= 0.f ? v : v * slope;
    }
    out[i] = v;
  }
}

// out[r, o] (+)= sum_j act(bn_j(in[r, j*O + o] + bias_j[o])); one thread per 4 output channels
__global__ void epilogue_sum_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int O, int J,
                                    const float* __restrict__ bias, const float* __restrict__ scale,
                                    const float* __restrict__ shift, int act, float slope, int accumulate) {
  const int O4 = O >> 2;
  const int64_t total = rows * O4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / O4;
    const int o = (int)(i % O4) * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < J; ++j) {
      const float4 v4 = *reinterpret_cast<const float4*>(in + (r * J + j) * O + o);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float y = v[c];
        if (y != -INFINITY) {
          const int col = j * O + o + c;
          if (bias) y += bias[col];
          if (scale) y = fmaf(y, scale[col], shift[col]);
          if (act == FGNN_ACT_RELU) y = fmaxf(y, 0.f);
          else if (act == FGNN_ACT_LEAKY_RELU) y = y >= 0.f ? y : y * slope;
        }
        acc[c] += y;
      }
    }
    float4* dst = reinterpret_cast<float4*>(out + r * O + o);
    if (accumulate) {
      const float4 old = *dst;
      acc[0] += old.x; acc[1] += old.y; acc[2] += old.z; acc[3] += old.w;
    }
    *dst = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// InstanceNorm2d (affine = false, biased variance) + activation over [B,C,N] with arbitrary strides; the statistics
// of instance (b, c) run over its N nodes (base_model.py:83-90: the v2v / f2f maps of FactorNN).
// Channels-last memory (x_sc == 1): block = 32 channels x 8 row groups, coalesced 128-byte rows, three passes over
// a 32-channel column block that stays in L1.
__global__ void __launch_bounds__(256)
instance_norm_cl_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int N, int64_t x_sb, int64_t x_sn,
                        int64_t o_sb, int64_t o_sn, float eps, int act, float slope) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx, b = blockIdx.y;
  const bool ok = c < C;
  const float* xb = x + (int64_t)b * x_sb + c;
  float s = 0.f;
  if (ok) for (int n = ty; n < N; n += 8) s += xb[(int64_t)n * x_sn];
  red[ty][tx] = s;
  __syncthreads();
  float mean = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) mean += red[j][tx];
  mean /= (float)N;
  __syncthreads();
  float q = 0.f;
  if (ok) for (int n = ty; n < N; n += 8) { const float d = xb[(int64_t)n * x_sn] - mean; q = fmaf(d, d, q); }
  red[ty][tx] = q;
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) var += red[j][tx];
  const float inv = rsqrtf(var / (float)N + eps);
  const float neg = act == FGNN_ACT_NONE ? 1.f : (act == FGNN_ACT_RELU ? 0.f : slope);
  float* ob = out + (int64_t)b * o_sb + c;
  if (ok) for (int n = ty; n < N; n += 8) {
    const float y = (xb[(int64_t)n * x_sn] - mean) * inv;
    ob[(int64_t)n * o_sn] = y >= 0.f ? y : y * neg;
  }
}

// Any other strides (channels-first: x_sn == 1 is the coalesced case): one warp per instance (b, c).
__global__ void __launch_bounds__(256)
instance_norm_warp_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C, int N, int64_t x_sb,
                          int64_t x_sc, int64_t x_sn, int64_t o_sb, int64_t o_sc, int64_t o_sn, float eps, int act,
                          float slope) {
  const int lane = threadIdx.x & 31;
  const int64_t inst = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (inst >= (int64_t)B * C) return;
  const int b = (int)(inst / C), c = (int)(inst % C);
  const float* xi = x + (int64_t)b * x_sb + (int64_t)c * x_sc;
  float s = 0.f;
  for (int n = lane; n < N; n += 32) s += xi[(int64_t)n * x_sn];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)N;
  float q = 0.f;
  for (int n = lane; n < N; n += 32) { const float d = xi[(int64_t)n * x_sn] - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float inv = rsqrtf(q / (float)N + eps);
  const float neg = act == FGNN_ACT_NONE ? 1.f : (act == FGNN_ACT_RELU ? 0.f : slope);
  float* oi = out + (int64_t)b * o_sb + (int64_t)c * o_sc;
  for (int n = lane; n < N; n += 32) {
    const float y = (xi[(int64_t)n * x_sn] - mean) * inv;
    oi[(int64_t)n * o_sn] = y >= 0.f ? y : y * neg;
  }
}

// Sharded InstanceNorm (the f2f / v2v maps of a FactorNN layer whose nodes are split over ranks, base_model.py:83-90,
// SURVEY 8f rank 4): the statistics of instance (b, c) run over ALL ranks' rows, so every rank reduces its own rows
// to partial sums -- pass 1: sum x; pass 2 (mean given): sum (x - mean)^2, the two-pass variance of the single-GPU
// kernel -- the host all-reduces the [B, C] partials (2 * C floats per instance), and the apply kernel normalises.
__global__ void __launch_bounds__(256)
instance_norm_partial_kernel(const float* __restrict__ x, const float* __restrict__ mean, float* __restrict__ out, int C, int N,
                             int64_t x_sb, int64_t x_sc, int64_t x_sn) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx, b = blockIdx.y;
  float s = 0.f;
  if (c < C) {
    const float* xb = x + (int64_t)b * x_sb + (int64_t)c * x_sc;
    const float m = mean ? mean[(int64_t)b * C + c] : 0.f;
    for (int n = ty; n < N; n += 8) {
      const float d = xb[(int64_t)n * x_sn] - m;
      s += mean ? d * d : d;
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) a += red[j][tx];
    out[(int64_t)b * C + c] = a;
  }
}

__global__ void __launch_bounds__(256)
instance_norm_apply_kernel(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ mean,
                           const float* __restrict__ inv_std, int B, int C, int N, int64_t x_sb, int64_t x_sc, int64_t x_sn,
                           int64_t o_sb, int64_t o_sc, int64_t o_sn, int act, float slope) {
  const int64_t total = (int64_t)B * C * N;
  const float neg = act == FGNN_ACT_NONE ? 1.f : (act == FGNN_ACT_RELU ? 0.f : slope);
  const bool cl = x_sc == 1;                                  // channels fastest in memory: iterate them fastest
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int b, c, n;
    if (cl) { c = (int)(i % C); n = (int)((i / C) % N); b = (int)(i / ((int64_t)C * N)); }
    else { n = (int)(i % N); c = (int)((i / N) % C); b = (int)(i / ((int64_t)C * N)); }
    const float y = (x[b * x_sb + c * x_sc + n * x_sn] - mean[(int64_t)b * C + c]) * inv_std[(int64_t)b * C + c];
    out[b * o_sb + c * o_sc + n * o_sn] = y >= 0.f ? y : y * neg;
  }
}

static int device_ok() {
  static int cached = -100;
  if (cached != -100) return cached;
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    return FGNN_ERR_NO_DEVICE;   // not cached: a later call may have a context
  }
  cached = (prop.major == 10) ? FGNN_OK : FGNN_ERR_NO_DEVICE;
  return cached;
}

}  // namespace fgnn

using namespace fgnn;

extern "C" {

int fgnn_version(void) { return FGNN_B200_VERSION; }

const char* fgnn_strerror(int status) {
  switch (status) {
    case FGNN_OK: return "ok";
    case FGNN_ERR_INVALID_ARG: return "invalid argument (null pointer, non-positive dimension or unknown enum; extension must one of mp_conv_type)";
    case FGNN_ERR_INDEX_RANGE: return "nn_idx entry out of range";
    case FGNN_ERR_SHAPE: return "shape mismatch (extension modes need M == N; filters rows must be C or 2C)";
    case FGNN_ERR_UNSUPPORTED: return "shape or dtype not supported by the selected kernel";
    case FGNN_ERR_WORKSPACE: return "workspace too small";
    case FGNN_ERR_CUDA: return "CUDA runtime error (see fgnn_last_cuda_error)";
    case FGNN_ERR_NO_DEVICE: return "no sm_100 (B200) device";
    default: return "unknown fgnn status";
  }
}

int fgnn_last_cuda_error(void) { return g_last_cuda_error; }

uint64_t fgnn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int fgnn_set_programmatic_launch(int enabled) {
  static std::atomic<int> current{1};
  const int prev = current.exchange(enabled ? 1 : 0);
  tc_set_pdl(enabled != 0);
  return prev;
}

int fgnn_mp_select_kernel(const fgnn_mp_args* a) {
  const int v = validate(a);
  if (v != FGNN_OK) return v;
  if (a->kernel == FGNN_KERNEL_SIMT) return FGNN_KERNEL_SIMT;
  const bool tc = tc_supported(a);
  if (a->kernel == FGNN_KERNEL_TCGEN05) return tc ? (int)FGNN_KERNEL_TCGEN05 : (int)FGNN_ERR_UNSUPPORTED;
  return tc ? (int)FGNN_KERNEL_TCGEN05 : (int)FGNN_KERNEL_SIMT;
}

size_t fgnn_mp_workspace_bytes(const fgnn_mp_args* a) {
  if (a && a->src_ptr && validate(a) == FGNN_OK && src_supported(a)) return tc_workspace_bytes(a);
  const int k = fgnn_mp_select_kernel(a);
  if (k == FGNN_KERNEL_TCGEN05) return tc_workspace_bytes(a);
  return 0;
}

int fgnn_mp_forward(const fgnn_mp_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int k = fgnn_mp_select_kernel(a);
  if (k < 0) return k;
  const int d = device_ok();
  if (d != FGNN_OK) return d;
  MpParams p = to_params(a);
  int rc;
  if (a->src_ptr) {                      // the caller asked for the source-stationary evaluation
    if (!src_supported(a)) return FGNN_ERR_UNSUPPORTED;
    if (a->workspace_bytes < tc_workspace_bytes(a) || !a->workspace) return FGNN_ERR_WORKSPACE;
    rc = launch_mp_src(p, a, stream);
    if (rc == FGNN_ERR_CUDA) g_last_cuda_error = (int)cudaGetLastError();
    return rc;
  }
  if (k == FGNN_KERNEL_TCGEN05) {
    if (a->workspace_bytes < tc_workspace_bytes(a) || (tc_workspace_bytes(a) && !a->workspace))
      return FGNN_ERR_WORKSPACE;
    rc = launch_mp_tc(p, a, stream);
  } else {
    if (a->dtype != FGNN_F32) return FGNN_ERR_UNSUPPORTED;
    rc = launch_mp_simt(p, stream);
  }
  if (rc == FGNN_ERR_CUDA) g_last_cuda_error = (int)cudaGetLastError();
  return rc;
}

int fgnn_mp_src_supported(const fgnn_mp_args* a) {
  if (validate(a) != FGNN_OK) return 0;
  return src_supported(a) ? 1 : 0;
}

int fgnn_src_permute_etype(const float* etype, int64_t et_sb, const int32_t* edge_slot, float* out, int32_t T,
                           int32_t M, int32_t K, int64_t n_edges, void* stream_) {
  if (!etype || !edge_slot || !out || T <= 0 || M <= 0 || K <= 0 || n_edges < 0) return FGNN_ERR_INVALID_ARG;
  const int rc = launch_et_permute(etype, et_sb, edge_slot, out, T, (int64_t)M * K, n_edges,
                                   reinterpret_cast<cudaStream_t>(stream_));
  if (rc == FGNN_ERR_CUDA) g_last_cuda_error = (int)cudaGetLastError();
  return rc;
}

int fgnn_instance_norm_forward(const float* x, float* out, int32_t B, int32_t C, int32_t N, int64_t x_sb, int64_t x_sc,
                               int64_t x_sn, int64_t out_sb, int64_t out_sc, int64_t out_sn, float eps,
                               int32_t activation, float act_slope, void* stream_) {
  if (!x || !out || B <= 0 || C <= 0 || N <= 0 || activation < 0 || activation > 2) return FGNN_ERR_INVALID_ARG;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (x_sc == 1 && out_sc == 1 && B <= 65535) {
    dim3 grid((C + 31) / 32, B);
    instance_norm_cl_kernel<<<grid, 256, 0, stream>>>(x, out, C, N, x_sb, x_sn, out_sb, out_sn, eps, activation, act_slope);
  } else {
    const int64_t inst = (int64_t)B * C;
    if ((inst + 7) / 8 > INT32_MAX) return FGNN_ERR_UNSUPPORTED;
    instance_norm_warp_kernel<<<(unsigned)((inst + 7) / 8), 256, 0, stream>>>(x, out, B, C, N, x_sb, x_sc, x_sn, out_sb, out_sc,
                                                                          out_sn, eps, activation, act_slope);
  }
  count_launch();
  FGNN_CUDA(cudaGetLastError());
  return FGNN_OK;
}

int fgnn_instance_norm_partial(const float* x, const float* mean, float* sums, int32_t B, int32_t C, int32_t N, int64_t x_sb,
                               int64_t x_sc, int64_t x_sn, void* stream_) {
  if (!x || !sums || B <= 0 || C <= 0 || N < 0 || B > 65535) return FGNN_ERR_INVALID_ARG;
  dim3 grid((C + 31) / 32, B);
  instance_norm_partial_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(x, mean, sums, C, N, x_sb, x_sc, x_sn);
  count_launch();
  FGNN_CUDA(cudaGetLastError());
  return FGNN_OK;
}

int fgnn_instance_norm_apply(const float* x, float* out, const float* mean, const float* inv_std, int32_t B, int32_t C, int32_t N,
                             int64_t x_sb, int64_t x_sc, int64_t x_sn, int64_t out_sb, int64_t out_sc, int64_t out_sn,
                             int32_t activation, float act_slope, void* stream_) {
  if (!x || !out || !mean || !inv_std || B <= 0 || C <= 0 || N < 0 || activation < 0 || activation > 2) return FGNN_ERR_INVALID_ARG;
  const int64_t total = (int64_t)B * C * N;
  if (total == 0) return FGNN_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  instance_norm_apply_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      x, out, mean, inv_std, B, C, N, x_sb, x_sc, x_sn, out_sb, out_sc, out_sn, activation, act_slope);
  count_launch();
  FGNN_CUDA(cudaGetLastError());
  return FGNN_OK;
}

int fgnn_to_node_major(const float* x, float* out, int32_t B, int32_t C, int32_t N, int64_t x_sb,
                       int64_t x_sc, int64_t x_sn, void* stream_) {
  if (!x || !out || B <= 0 || C <= 0 || N <= 0) return FGNN_ERR_INVALID_ARG;
  if (B > 65535 || (C + 31) / 32 > 65535) return FGNN_ERR_UNSUPPORTED;
  dim3 grid((N + 31) / 32, (C + 31) / 32, B), block(32, 8);
  to_node_major_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(x, out, C, N, x_sb, x_sc, x_sn);
  count_launch();
  FGNN_CUDA(cudaGetLastError());
  return FGNN_OK;
}

int fgnn_check_index_range(const void* idx, int idx_dtype, int64_t count, int64_t lo, int64_t n,
                           void* scratch, void* stream_) {
  if (!idx || !scratch || count < 0) return FGNN_ERR_INVALID_ARG;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int* flag = reinterpret_cast<int*>(scratch);
  FGNN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), stream));
  if (count > 0) {
    int blocks = (int)((count + 255) / 256 < 1184 ? (count + 255) / 256 : 1184);
    check_index_kernel<<<blocks, 256, 0, stream>>>(idx, idx_dtype == FGNN_I64, count, lo, n, flag);
    count_launch();
    FGNN_CUDA(cudaGetLastError());
  }
  int host = 0;
  FGNN_CUDA(cudaMemcpyAsync(&host, flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
  FGNN_CUDA(cudaStreamSynchronize(stream));
  return host ? FGNN_ERR_INDEX_RANGE : FGNN_OK;
}

int fgnn_check_index_range_async(const void* idx, int idx_dtype, int64_t count, int64_t lo, int64_t n,
                                 int32_t* flag, void* stream_) {
  if (!idx || !flag || count < 0) return FGNN_ERR_INVALID_ARG;
  if (count == 0) return FGNN_OK;
  int blocks = (int)((count + 255) / 256 < 1184 ? (count + 255) / 256 : 1184);
  check_index_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(idx, idx_dtype == FGNN_I64, count, lo, n,
                                                                              reinterpret_cast<int*>(flag));
  count_launch();
  FGNN_CUDA(cudaGetLastError());
  return FGNN_OK;
}

int fgnn_epilogue_forward(const float* in, float* out, int64_t rows, int32_t O, const float* bias,
                          const float* bn_scale, const float* bn_shift, int32_t activation,
                          float act_slope, void* stream_) {
  if (!in || !out || rows < 0 || O <= 0) return FGNN_ERR_INVALID_ARG;
  if ((bn_scale == nullptr) != (bn_shift == nullptr)) return FGNN_ERR_INVALID_ARG;
  const int64_t total = rows * O;
  if (total == 0) return FGNN_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  epilogue_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      in, out, total, O, bias, bn_scale, bn_shift, activation, act_slope);
  count_launch();
  FGNN_CUDA(cudaGetLastError());
  return FGNN_OK;
}

int fgnn_epilogue_sum_forward(const float* in, float* out, int64_t rows, int32_t O, int32_t J, const float* bias,
                              const float* bn_scale, const float* bn_shift, int32_t activation, float act_slope,
                              int32_t accumulate, void* stream_) {
  if (!in || !out || rows < 0 || O <= 0 || J <= 0 || (O & 3)) return FGNN_ERR_INVALID_ARG;
  if ((bn_scale == nullptr) != (bn_shift == nullptr)) return FGNN_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return FGNN_ERR_INVALID_ARG;
  const int64_t total = rows * (O / 4);
  if (total == 0) return FGNN_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  epilogue_sum_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      in, out, rows, O, J, bias, bn_scale, bn_shift, activation, act_slope, accumulate);
  count_launch();
  FGNN_CUDA(cudaGetLastError());
  return FGNN_OK;
}

// Host-buffer entry: contiguous reference layouts only (x [B,C,N], idx [B,M,K], etype [B,T,M,K],
// out [B,O,M,Kout]); strides in the args are ignored.
int fgnn_mp_forward_host(const fgnn_mp_args* h) {
  const int v = validate(h);
  if (v != FGNN_OK) return v;
  if (h->dtype != FGNN_F32) return FGNN_ERR_UNSUPPORTED;
  const int Kc = h->extension ? 2 * h->C : h->C;
  const int Kout = h->aggregator == FGNN_AGG_NONE ? h->K : 1;
  const size_t n_x = (size_t)h->B * h->C * h->N, n_idx = (size_t)h->B * h->M * h->K;
  const size_t n_et = n_idx * h->T, n_w = (size_t)Kc * h->O * h->T, n_out = (size_t)h->B * h->O * h->M * Kout;
  const size_t isz = h->idx_dtype == FGNN_I64 ? 8 : 4;
  float *dx = nullptr, *det = nullptr, *dw = nullptr, *dbias = nullptr, *dsc = nullptr, *dsh = nullptr,
        *dout = nullptr, *dxt = nullptr;
  void *didx = nullptr, *dws = nullptr;
  int rc = FGNN_OK;
  cudaStream_t s = nullptr;
#define HCHK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { rc = cuda_fail(_e); goto done; } } while (0)
  HCHK(cudaStreamCreate(&s));
  HCHK(cudaMalloc(&dx, n_x * 4)); HCHK(cudaMalloc(&dxt, n_x * 4));
  HCHK(cudaMalloc(&didx, n_idx * isz)); HCHK(cudaMalloc(&det, n_et * 4));
  HCHK(cudaMalloc(&dw, n_w * 4)); HCHK(cudaMalloc(&dout, n_out * 4));
  HCHK(cudaMemcpyAsync(dx, h->x, n_x * 4, cudaMemcpyHostToDevice, s));
  HCHK(cudaMemcpyAsync(didx, h->idx, n_idx * isz, cudaMemcpyHostToDevice, s));
  HCHK(cudaMemcpyAsync(det, h->etype, n_et * 4, cudaMemcpyHostToDevice, s));
  HCHK(cudaMemcpyAsync(dw, h->filters, n_w * 4, cudaMemcpyHostToDevice, s));
  if (h->bias) { HCHK(cudaMalloc(&dbias, h->O * 4)); HCHK(cudaMemcpyAsync(dbias, h->bias, h->O * 4, cudaMemcpyHostToDevice, s)); }
  if (h->bn_scale) {
    HCHK(cudaMalloc(&dsc, h->O * 4)); HCHK(cudaMalloc(&dsh, h->O * 4));
    HCHK(cudaMemcpyAsync(dsc, h->bn_scale, h->O * 4, cudaMemcpyHostToDevice, s));
    HCHK(cudaMemcpyAsync(dsh, h->bn_shift, h->O * 4, cudaMemcpyHostToDevice, s));
  }
  {
    fgnn_mp_args d = *h;
    // node-major copy of x on the device (the kernels gather whole rows)
    rc = fgnn_to_node_major(dx, dxt, h->B, h->C, h->N, (int64_t)h->C * h->N, h->N, 1, s);
    if (rc != FGNN_OK) goto done;
    d.x = dxt; d.x_sb = (int64_t)h->N * h->C; d.x_sc = 1; d.x_sn = h->C;
    d.idx = didx; d.idx_sb = (int64_t)h->M * h->K;
    d.etype = det; d.et_sb = (int64_t)h->T * h->M * h->K;
    d.filters = dw; d.bias = dbias; d.bn_scale = dsc; d.bn_shift = dsh;
    d.out = dout;
    d.out_sb = (int64_t)h->O * h->M * Kout; d.out_so = (int64_t)h->M * Kout; d.out_sm = Kout; d.out_sk = 1;
    d.filters_version = 0;
    d.tile_slots = nullptr; d.out_rows = nullptr;
    d.src_ptr = nullptr; d.slot_edge = nullptr; d.etype_edges = nullptr; d.messages = nullptr; d.src_rows = nullptr;
    d.workspace = nullptr; d.workspace_bytes = 0;
    const size_t ws = fgnn_mp_workspace_bytes(&d);
    if (ws) { HCHK(cudaMalloc(&dws, ws)); d.workspace = dws; d.workspace_bytes = ws; }
    {
      void* flag = nullptr;
      HCHK(cudaMalloc(&flag, 8));
      const int64_t lo = (h->flags & FGNN_FLAG_MASK_NEGATIVE) ? INT64_MIN : 0;
      rc = fgnn_check_index_range(didx, h->idx_dtype, (int64_t)n_idx, lo, h->N, flag, s);
      cudaFree(flag);
      if (rc != FGNN_OK) goto done;
    }
    rc = fgnn_mp_forward(&d, s);
    if (rc != FGNN_OK) goto done;
  }
  HCHK(cudaMemcpyAsync(h->out, dout, n_out * 4, cudaMemcpyDeviceToHost, s));
  HCHK(cudaStreamSynchronize(s));
done:
  cudaFree(dx); cudaFree(dxt); cudaFree(didx); cudaFree(det); cudaFree(dw); cudaFree(dbias);
  cudaFree(dsc); cudaFree(dsh); cudaFree(dout); cudaFree(dws);
  if (s) cudaStreamDestroy(s);
#undef HCHK
  return rc;
}

}  // extern "C"

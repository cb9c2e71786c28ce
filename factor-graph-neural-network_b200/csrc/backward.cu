// Backward of the message-passing call (training: reference train_ldpc.py:222-231, mp_nn.py:115-175 under autograd).
//
//   y[b,o,m] = AGG_k e[b,o,m,k],   e[(b,m,k), o] = sum_t et[b,t,m,k] * H[(b,m,k), o*T+t],   H = Xin W,
//   Xin[(b,m,k), :] = x[b,:,idx]  |  [x[b,:,m] || x[b,:,idx]]  |  [x[b,:,m] || x[b,:,m] - x[b,:,idx]]
//
// With g = dL/dy and a = dAGG/de (one-hot at the arg-max slot | softmax_k(gamma e) | 1/K):
//   ge[(b,m,k), o] = g[b,o,m] * a[b,o,m,k]
//   Z[(b,m,k), o*T+t] = ge[.,o] * et[b,t,m,k]          dXin = Z W^T      dW = Xin^T Z
//   d et[b,t,m,k] = sum_o ge[.,o] * H[., o*T+t]         dx = scatter-add of dXin through idx (and the self part)
// The three dense products are plain GEMMs (the host side runs them through the framework's BLAS); everything with
// graph structure in it -- the gather, the per-slot contraction, the aggregator's derivative, the outer product /
// edge-type gradient and the scatter-add -- are the kernels below.  Slots are processed in chunks of destination
// rows, so the O*T-wide intermediates stay bounded.
#include "common.cuh"

namespace fgnn {

namespace {

// Xin rows of the slots (b, m, k), m in [m0, m0 + mc): out [B*mc*K, Cin]
__global__ void bwd_gather_kernel(const float* __restrict__ x, const void* __restrict__ idx, int idx64, float* __restrict__ out,
                                  int B, int N, int M, int K, int C, int ext, int m0, int mc, int64_t x_sb, int64_t x_sc,
                                  int64_t x_sn, int64_t idx_sb) {
  const int Cin = ext ? 2 * C : C;
  const int64_t total = (int64_t)B * mc * K * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    const int64_t s = i / Cin;                               // (b*mc + ml)*K + k
    const int k = (int)(s % K);
    const int64_t bm = s / K;
    const int ml = (int)(bm % mc), b = (int)(bm / mc), m = m0 + ml;
    const int64_t n = load_index(idx, idx64, (int64_t)b * idx_sb + (int64_t)m * K + k);
    const bool ok = n >= 0 && n < N;
    float v;
    if (!ext) {
      v = ok ? x[b * x_sb + c * x_sc + n * x_sn] : 0.f;
    } else if (c < C) {
      v = x[b * x_sb + c * x_sc + (int64_t)m * x_sn];
    } else {
      const float xj = ok ? x[b * x_sb + (c - C) * x_sc + n * x_sn] : 0.f;
      v = ext == FGNN_ORIG_WITH_DIFF ? x[b * x_sb + (c - C) * x_sc + (int64_t)m * x_sn] - xj : xj;
    }
    out[i] = v;
  }
}

// e[s, o] = sum_t et[b,t,m,k] * H[s, o*T+t]
__global__ void bwd_slot_values_kernel(const float* __restrict__ H, const float* __restrict__ et, float* __restrict__ e,
                                       int B, int M, int K, int O, int T, int m0, int mc, int64_t et_sb) {
  const int64_t total = (int64_t)B * mc * K * O;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % O);
    const int64_t s = i / O;
    const int k = (int)(s % K);
    const int64_t bm = s / K;
    const int ml = (int)(bm % mc), b = (int)(bm / mc), m = m0 + ml;
    const float* h = H + s * (int64_t)O * T + (int64_t)o * T;
    const float* pe = et + b * et_sb + (int64_t)m * K + k;
    float a = 0.f;
    for (int t = 0; t < T; ++t) a = fmaf(pe[(int64_t)t * M * K], h[t], a);
    e[i] = a;
  }
}

// ge[s, o] = g[b,o,m] * dAGG/de; one thread per (b, m, o), slots k = 0..K-1 (negative index = masked slot when mask_neg)
__global__ void bwd_aggregate_kernel(const float* __restrict__ e, const float* __restrict__ g, const void* __restrict__ idx,
                                     int idx64, float* __restrict__ ge, int B, int M, int K, int O, int agg, float gamma,
                                     int mask_neg, int m0, int mc, int64_t g_sb, int64_t g_so, int64_t g_sm, int64_t g_sk,
                                     int64_t idx_sb) {
  const int64_t total = (int64_t)B * mc * O;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % O);
    const int64_t bm = i / O;
    const int ml = (int)(bm % mc), b = (int)(bm / mc), m = m0 + ml;
    const float* pe = e + (bm * K) * O + o;                  // slot k at pe[k*O]
    float* pg = ge + (bm * K) * O + o;
    const int64_t ib = (int64_t)b * idx_sb + (int64_t)m * K;
    auto live = [&](int k) { return !mask_neg || load_index(idx, idx64, ib + k) >= 0; };
    if (agg == FGNN_AGG_NONE) {
      for (int k = 0; k < K; ++k) pg[(int64_t)k * O] = live(k) ? g[b * g_sb + o * g_so + (int64_t)m * g_sm + k * g_sk] : 0.f;
      continue;
    }
    const float go = g[b * g_sb + o * g_so + (int64_t)m * g_sm];
    if (agg == FGNN_AGG_MAX) {
      int best = -1;
      float bv = -INFINITY;
      for (int k = 0; k < K; ++k) {
        const float v = pe[(int64_t)k * O];
        if (live(k) && (best < 0 || v > bv)) { best = k; bv = v; }       // first maximum, like torch.max's backward
      }
      for (int k = 0; k < K; ++k) pg[(int64_t)k * O] = k == best ? go : 0.f;
    } else if (agg == FGNN_AGG_SOFTMAX) {
      float mx = -INFINITY;
      for (int k = 0; k < K; ++k) if (live(k)) mx = fmaxf(mx, gamma * pe[(int64_t)k * O]);
      float sum = 0.f;
      for (int k = 0; k < K; ++k) if (live(k)) sum += expf(gamma * pe[(int64_t)k * O] - mx);
      for (int k = 0; k < K; ++k) pg[(int64_t)k * O] = live(k) ? go * expf(gamma * pe[(int64_t)k * O] - mx) / sum : 0.f;
    } else {
      int cnt = 0;
      for (int k = 0; k < K; ++k) cnt += live(k) ? 1 : 0;
      for (int k = 0; k < K; ++k) pg[(int64_t)k * O] = live(k) ? go / (float)cnt : 0.f;
    }
  }
}

// per slot s: d_et[b,t,m,k] = sum_o ge[s,o] * H[s,o*T+t]; then H[s,o*T+t] <- Z = ge[s,o] * et[b,t,m,k] (in place).
// One warp per slot: lane handles columns lane, lane+32, ... of the O*T row.
__global__ void __launch_bounds__(256)
bwd_outer_kernel(float* __restrict__ H, const float* __restrict__ ge, const float* __restrict__ et, float* __restrict__ d_et,
                 int B, int M, int K, int O, int T, int m0, int mc, int64_t et_sb, int64_t det_sb, int want_det) {
  const int lane = threadIdx.x & 31;
  const int64_t slots = (int64_t)B * mc * K;
  const int OT = O * T;
  for (int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < slots; s += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int k = (int)(s % K);
    const int64_t bm = s / K;
    const int ml = (int)(bm % mc), b = (int)(bm / mc), m = m0 + ml;
    float* h = H + s * OT;
    const float* pg = ge + s * O;
    const float* pe = et + b * et_sb + (int64_t)m * K + k;
    // edge-type gradient: column n = o*T + t contributes ge[o] * h[n] to d_et[t]; T divides 32 or 32 divides T*...: reduce generically
    if (want_det) {
      for (int t = 0; t < T; ++t) {
        float a = 0.f;
        for (int o = lane; o < O; o += 32) a = fmaf(pg[o], h[o * T + t], a);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (lane == 0) d_et[b * det_sb + ((int64_t)t * M + m) * K + k] = a;
      }
    }
    __syncwarp();
    for (int n = lane; n < OT; n += 32) h[n] = pg[n / T] * pe[(int64_t)(n % T) * M * K];
  }
}

// dx[b, n, c] += dXin parts (dx node-major [B,N,C], zero-initialised by the caller)
__global__ void bwd_scatter_kernel(const float* __restrict__ dxin, const void* __restrict__ idx, int idx64, float* __restrict__ dx,
                                   int B, int N, int M, int K, int C, int ext, int m0, int mc, int64_t idx_sb) {
  const int Cin = ext ? 2 * C : C;
  const int64_t total = (int64_t)B * mc * K * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    const int64_t s = i / Cin;
    const int k = (int)(s % K);
    const int64_t bm = s / K;
    const int ml = (int)(bm % mc), b = (int)(bm / mc), m = m0 + ml;
    const int64_t n = load_index(idx, idx64, (int64_t)b * idx_sb + (int64_t)m * K + k);
    const bool ok = n >= 0 && n < N;
    const float v = dxin[i];
    if (!ext) {
      if (ok) atomicAdd(dx + ((int64_t)b * N + n) * C + c, v);
    } else if (c < C) {
      atomicAdd(dx + ((int64_t)b * N + m) * C + c, v);
    } else if (ext == FGNN_ORIG_WITH_DIFF) {
      atomicAdd(dx + ((int64_t)b * N + m) * C + (c - C), v);
      if (ok) atomicAdd(dx + ((int64_t)b * N + n) * C + (c - C), -v);
    } else if (ok) {
      atomicAdd(dx + ((int64_t)b * N + n) * C + (c - C), v);
    }
  }
}

int grid_for(int64_t total, int per_block = 256) {
  int64_t blocks = (total + per_block - 1) / per_block;
  if (blocks > 148 * 16) blocks = 148 * 16;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

}  // namespace fgnn

using namespace fgnn;

extern "C" {

int fgnn_bwd_gather(const fgnn_mp_args* a, float* xin, int32_t m0, int32_t mc, void* stream) {
  if (!a || !a->x || !a->idx || !xin || mc <= 0 || m0 < 0 || m0 + mc > a->M || a->dtype != FGNN_F32) return FGNN_ERR_INVALID_ARG;
  const int64_t total = (int64_t)a->B * mc * a->K * (a->extension ? 2 * a->C : a->C);
  bwd_gather_kernel<<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(a->x), a->idx, a->idx_dtype == FGNN_I64, xin, a->B, a->N, a->M, a->K, a->C, a->extension, m0, mc,
      a->x_sb, a->x_sc, a->x_sn, a->idx_sb);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

int fgnn_bwd_slot_values(const fgnn_mp_args* a, const float* H, float* e, int32_t m0, int32_t mc, void* stream) {
  if (!a || !a->etype || !H || !e || mc <= 0 || m0 < 0 || m0 + mc > a->M || a->dtype != FGNN_F32) return FGNN_ERR_INVALID_ARG;
  const int64_t total = (int64_t)a->B * mc * a->K * a->O;
  bwd_slot_values_kernel<<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      H, reinterpret_cast<const float*>(a->etype), e, a->B, a->M, a->K, a->O, a->T, m0, mc, a->et_sb);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

int fgnn_bwd_aggregate(const fgnn_mp_args* a, const float* e, const float* grad_out, int64_t g_sb, int64_t g_so, int64_t g_sm,
                       int64_t g_sk, float* ge, int32_t m0, int32_t mc, void* stream) {
  if (!a || !a->idx || !e || !grad_out || !ge || mc <= 0 || m0 < 0 || m0 + mc > a->M) return FGNN_ERR_INVALID_ARG;
  const int64_t total = (int64_t)a->B * mc * a->O;
  bwd_aggregate_kernel<<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      e, grad_out, a->idx, a->idx_dtype == FGNN_I64, ge, a->B, a->M, a->K, a->O, a->aggregator, a->gamma,
      (a->flags & FGNN_FLAG_MASK_NEGATIVE) ? 1 : 0, m0, mc, g_sb, g_so, g_sm, g_sk, a->idx_sb);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

int fgnn_bwd_outer(const fgnn_mp_args* a, float* H_inout, const float* ge, float* d_etype, int64_t det_sb, int32_t m0, int32_t mc,
                   void* stream) {
  if (!a || !a->etype || !H_inout || !ge || mc <= 0 || m0 < 0 || m0 + mc > a->M) return FGNN_ERR_INVALID_ARG;
  const int64_t slots = (int64_t)a->B * mc * a->K;
  bwd_outer_kernel<<<grid_for(slots, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      H_inout, ge, reinterpret_cast<const float*>(a->etype), d_etype, a->B, a->M, a->K, a->O, a->T, m0, mc, a->et_sb, det_sb,
      d_etype ? 1 : 0);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

int fgnn_bwd_scatter(const fgnn_mp_args* a, const float* dxin, float* dx, int32_t m0, int32_t mc, void* stream) {
  if (!a || !a->idx || !dxin || !dx || mc <= 0 || m0 < 0 || m0 + mc > a->M) return FGNN_ERR_INVALID_ARG;
  const int64_t total = (int64_t)a->B * mc * a->K * (a->extension ? 2 * a->C : a->C);
  bwd_scatter_kernel<<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dxin, a->idx, a->idx_dtype == FGNN_I64, dx, a->B, a->N, a->M, a->K, a->C, a->extension, m0, mc, a->idx_sb);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

}  // extern "C"

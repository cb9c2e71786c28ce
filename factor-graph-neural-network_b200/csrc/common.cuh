// Shared device-side definitions for the fgnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fgnn_b200.h"

namespace fgnn {

// Device-side view of one message-passing call (fgnn_mp_args after validation).
struct MpParams {
  const float* x;
  const void* idx;
  const float* et;
  const float* W;      // filters [Kc, O*T]
  const float* bias;   // or nullptr
  const float* scale;  // folded eval-BN or nullptr
  const float* shift;
  float* out;
  const int32_t* tile_k;    // optional per-128-row-tile slot count
  const int32_t* out_rows;  // optional destination-row -> output-row map
  int64_t x_sb, x_sc, x_sn;
  int64_t idx_sb, et_sb;
  int64_t o_sb, o_so, o_sm, o_sk;
  int B, N, M, K, C, O, T;
  int ext, agg, act, idx64, mask_neg, accumulate;
  float gamma, slope;
};

__device__ __forceinline__ int64_t load_index(const void* idx, int idx64, int64_t off) {
  return idx64 ? reinterpret_cast<const int64_t*>(idx)[off]
               : (int64_t) reinterpret_cast<const int32_t*>(idx)[off];
}

__device__ __forceinline__ float apply_epilogue(float v, int o, const MpParams& p) {
  if (p.bias) v += p.bias[o];                                 // mp_nn.py:165-168
  if (p.scale) v = fmaf(v, p.scale[o], p.shift[o]);           // mp_nn.py:169-170 (eval BN, folded)
  if (p.act == FGNN_ACT_RELU) v = fmaxf(v, 0.f);              // mp_nn.py:172-173
  else if (p.act == FGNN_ACT_LEAKY_RELU) v = v >= 0.f ? v : v * p.slope;
  return v;
}

// One step of the online logsumexp of the softmax aggregator (mp_nn.py:80-83), with every rounding spelled out
// (no FMA contraction left to the compiler): the kernels that must agree bit for bit all call this.
__device__ __forceinline__ void softmax_push(float& a, float& s, float e, float gamma) {
  const float z = __fmul_rn(gamma, e);
  const float m = fmaxf(a, z);
  s = __fmaf_rn(s, expf(__fsub_rn(a, m)), expf(__fsub_rn(z, m)));   // a = -inf on the first push: exp(-inf) = 0
  a = m;
}
__device__ __forceinline__ float softmax_finish(float a, float s, float gamma) {
  return __fmul_rn(__fadd_rn(logf(s), a), __frcp_rn(gamma));
}

// Online aggregator over the K slots of one destination (mp_nn.py:73-87).
struct AggState {
  float a;   // max: running max | softmax: running max of gamma*e | mean: running sum
  float s;   // softmax: running sum of exp(gamma*e - a) | mean: live-slot count
  __device__ __forceinline__ void init(int agg) {
    a = (agg == FGNN_AGG_MEAN) ? 0.f : -INFINITY;
    s = 0.f;
  }
  __device__ __forceinline__ void push(float e, int agg, float gamma) {
    if (agg == FGNN_AGG_MAX) {
      a = fmaxf(a, e);
    } else if (agg == FGNN_AGG_SOFTMAX) {
      softmax_push(a, s, e, gamma);
    } else {
      a += e;
      s += 1.f;
    }
  }
  __device__ __forceinline__ float finish(int agg, float gamma) const {
    if (agg == FGNN_AGG_MAX) return a;
    if (agg == FGNN_AGG_SOFTMAX) return s > 0.f ? softmax_finish(a, s, gamma) : -INFINITY;
    return s > 0.f ? a / s : 0.f;
  }
};

void count_launch();

int launch_mp_simt(const MpParams& p, cudaStream_t stream);

// tensor-core path (mp_tc.cu)
bool tc_supported(const fgnn_mp_args* a);
size_t tc_workspace_bytes(const fgnn_mp_args* a);
int launch_mp_tc(const MpParams& p, const fgnn_mp_args* a, cudaStream_t stream);
void tc_set_pdl(bool on);
bool tc_pdl_enabled();
int tc_prepare_weights(const float* W, uint8_t* ws, int C, int OT, int64_t version, cudaStream_t stream, int diff = 0, int T = 0);
int tc_num_sms();

// source-stationary path (mp_src.cu)
bool src_supported(const fgnn_mp_args* a);
int launch_mp_src(const MpParams& p, const fgnn_mp_args* a, cudaStream_t stream);
int launch_et_permute(const float* et, int64_t et_sb, const int32_t* edge_slot, float* out, int T, int64_t MK,
                      int64_t n_edges, cudaStream_t stream);

}  // namespace fgnn

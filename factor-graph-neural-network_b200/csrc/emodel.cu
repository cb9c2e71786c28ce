// The edge model of the FGNN scripts as ONE kernel: etype = Conv1x1(H -> T)(ReLU(Conv1x1(Fe -> H)(efeature)))
// (reference: train_ldpc.py:32-38,68-69; train_syn_hop_factor.py:174-179; train_syn_fixed_pw_hop.py:166-169).
//
// The reference runs conv -> ReLU -> conv as three ATen launches with an H = 64 channel hidden tensor per slot in
// between (256 B per slot written and read back: 154 MB each way for the 600 K slots of a cfg-2 pairwise table).
// Here a thread owns one slot (b, m, k): it reads the slot's Fe (1-7) edge features, keeps the hidden vector in
// registers (weights in shared memory, broadcast reads) and writes the T edge types -- either in the reference
// layout [B,T,M,K], or straight into a source-stationary plan's edge-major image (edge order + 16-byte piece
// swizzle of csrc/mp_src.cu), which also fuses fgnn_src_permute_etype.  SURVEY 8f rank 2.
//
// Why not inside the message-passing kernel itself: the edge types are layer-invariant (computed once per forward,
// read by every layer), LDPC's efeature (7 floats) is LARGER than its etype (4 floats), and at T = 16 the core is
// bound by the tensor pipe and tensor-memory reads, not by HBM -- the (Fe + T) * 64 FMAs per slot would land on the
// epilogue warps that already pace it.
#include "common.cuh"

namespace fgnn {

namespace {

constexpr int kH = 64;        // hidden width of every edge model in the reference scripts
constexpr int kMaxFe = 8, kMaxT = 16;

struct EmParams {
  const float* ef;            // [B, Fe, M, K] element (b,f,m,k) at b*ef_sb + f*M*K + m*K + k
  const float* w1; const float* b1;     // [H, Fe], [H]
  const float* w2; const float* b2;     // [T, H], [T]
  float* out;
  const int32_t* edge_slot;   // nullptr: out = [B,T,M,K]; else out = edge image [E,T] of the plan (slot of every edge)
  int64_t ef_sb, out_sb, MK, total;     // total = B*M*K slots, or E edges
  int Fe, T;
};

__device__ __forceinline__ uint32_t em_key(int T, uint32_t e) {      // et_key of mp_src.cu
  const uint32_t ppe = (uint32_t)T / 4u;
  return ppe <= 1 ? 0u : (e / (8u / ppe)) & (ppe - 1u);
}

template <int FE>
__global__ void __launch_bounds__(256)
emodel_kernel(const EmParams p) {
  __shared__ float s_w1[kH * kMaxFe], s_b1[kH], s_w2[kMaxT * kH], s_b2[kMaxT];
  for (int i = threadIdx.x; i < kH * FE; i += blockDim.x) s_w1[i] = p.w1[i];
  for (int i = threadIdx.x; i < kH; i += blockDim.x) s_b1[i] = p.b1 ? p.b1[i] : 0.f;
  for (int i = threadIdx.x; i < p.T * kH; i += blockDim.x) s_w2[i] = p.w2[i];
  for (int i = threadIdx.x; i < p.T; i += blockDim.x) s_b2[i] = p.b2 ? p.b2[i] : 0.f;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t slot = p.edge_slot ? (int64_t)p.edge_slot[i] : i;        // (b*M + m)*K + k
    const int64_t b = slot / p.MK, mk = slot - b * p.MK;
    float f[FE];
#pragma unroll
    for (int j = 0; j < FE; ++j) f[j] = p.ef[b * p.ef_sb + (int64_t)j * p.MK + mk];
    float h[kH];
#pragma unroll
    for (int c = 0; c < kH; ++c) {
      float a = s_b1[c];
#pragma unroll
      for (int j = 0; j < FE; ++j) a = fmaf(s_w1[c * FE + j], f[j], a);
      h[c] = fmaxf(a, 0.f);
    }
    for (int t = 0; t < p.T; ++t) {
      float a = s_b2[t];
#pragma unroll
      for (int c = 0; c < kH; ++c) a = fmaf(s_w2[t * kH + c], h[c], a);
      if (p.edge_slot) {
        const uint32_t key = em_key(p.T, (uint32_t)i);
        p.out[i * p.T + ((((uint32_t)t >> 2) ^ key) << 2) + (t & 3)] = a;
      } else {
        p.out[b * p.out_sb + (int64_t)t * p.MK + mk] = a;
      }
    }
  }
}

}  // namespace

}  // namespace fgnn

using namespace fgnn;

extern "C" int fgnn_emodel_forward(const float* efeature, int64_t ef_sb, const float* w1, const float* b1, const float* w2,
                                   const float* b2, float* out, int64_t out_sb, const int32_t* edge_slot, int64_t n_edges,
                                   int32_t B, int32_t Fe, int32_t H, int32_t T, int32_t M, int32_t K, void* stream_) {
  if (!efeature || !w1 || !w2 || !out || B <= 0 || Fe <= 0 || T <= 0 || M <= 0 || K <= 0) return FGNN_ERR_INVALID_ARG;
  if (H != kH || Fe > kMaxFe || T > kMaxT) return FGNN_ERR_UNSUPPORTED;
  if (edge_slot && (n_edges < 0 || (T % 4 && T != 1 && T != 2))) return FGNN_ERR_INVALID_ARG;
  EmParams p;
  p.ef = efeature; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.out = out; p.edge_slot = edge_slot;
  p.ef_sb = ef_sb; p.out_sb = out_sb; p.MK = (int64_t)M * K;
  p.total = edge_slot ? n_edges : (int64_t)B * M * K;
  p.Fe = Fe; p.T = T;
  if (p.total == 0) return FGNN_OK;
  int64_t blocks = (p.total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  switch (Fe) {
    case 1: emodel_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    case 2: emodel_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    case 3: emodel_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    case 4: emodel_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    case 5: emodel_kernel<5><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    case 6: emodel_kernel<6><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    case 7: emodel_kernel<7><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    default: emodel_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(p); break;
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

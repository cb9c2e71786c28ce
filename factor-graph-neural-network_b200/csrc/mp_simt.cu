// fp32 CUDA-core kernel for the FGNN message-passing call: every shape, every extension mode and
// aggregator of the reference's mp_conv_v2.forward (lib/model/mpnn/mp_nn.py:115-175).
//
// One thread block owns RG destinations; thread (rg, og) owns output channel `og` of destination
// `rg` and walks that destination's K slots RB at a time.  Per pass the block stages the RB*RG
// gathered input rows ([x_j], [x_i || x_j] or [x_i || x_i - x_j]) and their edge-type vectors in
// shared memory, each thread accumulates H[row, o*T + t] = sum_c A[row,c] W[c, o*T+t] for its
// channel with the weight vector held in registers across the RB rows, contracts it with
// etype[row, t], and pushes the K results through an online max / logsumexp / mean.  Nothing is
// materialised in HBM (the reference materialises H, the int64 index expansion and the gathered
// rows: SURVEY 2b k3-k5).
//
// This kernel is the reference-grade path for shapes the tcgen05 kernel does not take (C=2, O=2,
// extension modes, K>..); it is compute-bound on FFMA by design.
#include "common.cuh"

namespace fgnn {

constexpr int kSimtThreads = 256;
constexpr int kRB = 4;       // slots per pass per destination (register blocking over rows)
constexpr int kRGMax = 8;    // destinations per block

template <int TV>
__global__ void __launch_bounds__(kSimtThreads)
mp_simt_kernel(const MpParams p, const int og_count, const int RG) {
  extern __shared__ float smem[];
  const int Kc = p.ext ? 2 * p.C : p.C;
  const int OT = p.O * p.T;
  float* A_s = smem;                                   // [RG][kRB][Kc]
  float* et_s = A_s + RG * kRB * Kc;                   // [RG][kRB][T]
  int64_t* n_s = reinterpret_cast<int64_t*>(et_s + ((RG * kRB * p.T + 1) & ~1));  // [RG][kRB] source row, -1 = dead
  const int tid = threadIdx.x;
  const int og = tid % og_count, rg = tid / og_count;
  const int64_t total = (int64_t)p.B * p.M;
  const int64_t g0 = (int64_t)blockIdx.x * RG;
  const int64_t g = g0 + rg;
  const bool dest_ok = rg < RG && g < total;
  const int b = dest_ok ? (int)(g / p.M) : 0;
  const int m = dest_ok ? (int)(g % p.M) : 0;

  for (int o0 = 0; o0 < p.O; o0 += og_count) {
    const int o = o0 + og;
    const bool active = dest_ok && o < p.O;
    AggState st;
    st.init(p.agg);
    for (int k0 = 0; k0 < p.K; k0 += kRB) {                 // (slots beyond tile_k[g/128] are dead, see below)
      __syncthreads();
      // 1) slot -> source row
      if (tid < RG * kRB) {
        const int rgi = tid / kRB, r = tid % kRB;
        const int64_t gi = g0 + rgi;
        const int k = k0 + r;
        int64_t n = -1;
        if (gi < total && k < p.K) {
          const int bi = (int)(gi / p.M), mi = (int)(gi % p.M);
          n = load_index(p.idx, p.idx64, (int64_t)bi * p.idx_sb + (int64_t)mi * p.K + k);
          if (n < 0 || n >= p.N) n = -1;   // masked (FGNN_FLAG_MASK_NEGATIVE) or invalid: dead slot
          if (p.tile_k && k >= p.tile_k[gi / 128]) n = -1;   // compacted table: beyond this tile's slot count
        }
        n_s[tid] = n;
      }
      __syncthreads();
      // 2) stage gathered rows and edge-type vectors
      for (int i = tid; i < RG * kRB * Kc; i += kSimtThreads) {
        const int row = i / Kc, c = i % Kc;
        const int64_t n = n_s[row];
        float v = 0.f;
        if (n >= 0) {
          const int64_t gi = g0 + row / kRB;
          const int bi = (int)(gi / p.M), mi = (int)(gi % p.M);
          const float* xb = p.x + (int64_t)bi * p.x_sb;
          if (p.ext == FGNN_NO_EXTENSION) {
            v = xb[(int64_t)c * p.x_sc + n * p.x_sn];
          } else if (c < p.C) {
            v = xb[(int64_t)c * p.x_sc + (int64_t)mi * p.x_sn];                       // x_i
          } else {
            const float xj = xb[(int64_t)(c - p.C) * p.x_sc + n * p.x_sn];
            v = p.ext == FGNN_ORIG_WITH_DIFF
                    ? xb[(int64_t)(c - p.C) * p.x_sc + (int64_t)mi * p.x_sn] - xj      // x_i - x_j
                    : xj;
          }
        }
        A_s[i] = v;
      }
      for (int i = tid; i < RG * kRB * p.T; i += kSimtThreads) {
        const int row = i / p.T, t = i % p.T;
        float v = 0.f;
        if (n_s[row] >= 0) {
          const int64_t gi = g0 + row / kRB;
          const int bi = (int)(gi / p.M), mi = (int)(gi % p.M);
          const int k = k0 + row % kRB;
          v = p.et[(int64_t)bi * p.et_sb + ((int64_t)t * p.M + mi) * p.K + k];
        }
        et_s[i] = v;
      }
      __syncthreads();
      if (!active) continue;
      // 3) H row for channel o, contracted with the edge type (mp_nn.py:127-134)
      float e[kRB];
#pragma unroll
      for (int r = 0; r < kRB; ++r) e[r] = 0.f;
      const float* a_base = A_s + rg * kRB * Kc;
      const float* et_base = et_s + rg * kRB * p.T;
      for (int t0 = 0; t0 < p.T; t0 += TV) {
        float acc[kRB][TV];
#pragma unroll
        for (int r = 0; r < kRB; ++r)
#pragma unroll
          for (int v = 0; v < TV; ++v) acc[r][v] = 0.f;
        const float* wp = p.W + (int64_t)o * p.T + t0;
        for (int c = 0; c < Kc; ++c) {
          float w[TV];
          if (TV == 4) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)c * OT));
            w[0] = w4.x; w[1 % TV] = w4.y; w[2 % TV] = w4.z; w[3 % TV] = w4.w;
          } else {
#pragma unroll
            for (int v = 0; v < TV; ++v) w[v] = __ldg(wp + (int64_t)c * OT + v);
          }
#pragma unroll
          for (int r = 0; r < kRB; ++r) {
            const float a = a_base[r * Kc + c];
#pragma unroll
            for (int v = 0; v < TV; ++v) acc[r][v] = fmaf(a, w[v], acc[r][v]);
          }
        }
#pragma unroll
        for (int r = 0; r < kRB; ++r)
#pragma unroll
          for (int v = 0; v < TV; ++v) e[r] = fmaf(et_base[r * p.T + t0 + v], acc[r][v], e[r]);
      }
      // 4) aggregate (mp_nn.py:162-163) or write the slot straight out (aggregtor=None)
#pragma unroll
      for (int r = 0; r < kRB; ++r) {
        const int k = k0 + r;
        if (k >= p.K) break;
        const bool live = n_s[rg * kRB + r] >= 0;
        if (p.agg == FGNN_AGG_NONE) {
          const float v = live ? apply_epilogue(e[r], o, p) : 0.f;
          p.out[(int64_t)b * p.o_sb + (int64_t)o * p.o_so + (int64_t)m * p.o_sm + (int64_t)k * p.o_sk] = v;
        } else if (live) {
          st.push(e[r], p.agg, p.gamma);
        }
      }
    }
    if (active && p.agg != FGNN_AGG_NONE) {
      float v = st.finish(p.agg, p.gamma);
      if (v != -INFINITY) v = apply_epilogue(v, o, p);   // -inf = no live slot (sharded tables): keep
      float* dst = p.out + (int64_t)b * p.o_sb + (int64_t)o * p.o_so + (int64_t)m * p.o_sm;
      if (p.out_rows) {
        const int32_t row = p.out_rows[g];
        dst = row >= 0 ? p.out + (int64_t)row * p.o_sm + (int64_t)o * p.o_so : nullptr;
      }
      if (dst) *dst = p.accumulate ? *dst + v : v;
    }
  }
}

static int next_pow2(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

int launch_mp_simt(const MpParams& p, cudaStream_t stream) {
  const int Kc = p.ext ? 2 * p.C : p.C;
  int og_count = next_pow2(p.O);
  if (og_count > kSimtThreads) og_count = kSimtThreads;
  int RG = kSimtThreads / og_count;
  if (RG > kRGMax) RG = kRGMax;
  const int64_t total = (int64_t)p.B * p.M;
  if (total < RG) RG = (int)total > 0 ? (int)total : 1;
  const size_t smem = (size_t)RG * kRB * Kc * 4 + (size_t)((RG * kRB * p.T + 1) & ~1) * 4 + (size_t)RG * kRB * 8;
  if (smem > 200 * 1024) return FGNN_ERR_UNSUPPORTED;
  const int64_t blocks = (total + RG - 1) / RG;
  if (blocks <= 0 || blocks > 0x7fffffffLL) return FGNN_ERR_UNSUPPORTED;
  const bool vec4 = (p.T % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.W) & 15) == 0);
  auto kern = vec4 ? mp_simt_kernel<4> : mp_simt_kernel<1>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return FGNN_ERR_CUDA;
  }
  kern<<<(unsigned)blocks, kSimtThreads, smem, stream>>>(p, og_count, RG);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? FGNN_OK : FGNN_ERR_CUDA;
}

}  // namespace fgnn

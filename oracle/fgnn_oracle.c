/* CPU oracle (C restatement) of the FGNN message-passing hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs load this library.
 *
 * Restates mp_conv_v2.forward, /root/reference/lib/model/mpnn/mp_nn.py:115-175, in the
 * reference's own operation order (per-source linear map first, then row gather, then the
 * per-slot edge-type contraction, then aggregate / bias / eval-BN / activation), fp32 throughout:
 *   NO_EXTENSION           mp_nn.py:124-134   H = X W ; E[b,m,k,o] = sum_t H[b,idx,o*T+t] etype[b,t,m,k]
 *   ORIG_WITH_NEIGHBOR/DIFF mp_nn.py:136-159  per slot [x_i || x_j] or [x_i || x_i - x_j] times W[2C,O*T]
 *   aggregators            mp_nn.py:73-87     max | 1/g logsumexp(g .) | mean | none
 *   epilogue               mp_nn.py:165-173   + bias ; BatchNorm2d eval ; ReLU
 * Layouts are the reference's: x [B,C,N], idx [B,M,K] int64, etype [B,T,M,K],
 * filters [C or 2C, O*T] (column o*T+t), out [B,O,M,1] (or [B,O,M,K] when aggregator = none).
 *
 * Parity pinning: checked against tests/golden/ fixtures (outputs of the real reference run in the
 * build container; see tests/golden/make_golden.py) by tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { EXT_NONE = 0, EXT_NEIGHBOR = 1, EXT_DIFF = 2 };
enum { AGG_MAX = 0, AGG_SOFTMAX = 1, AGG_MEAN = 2, AGG_NONE = 3 };

int fgnn_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* out[j] (+)= a * w[j], j < n : the inner kernel of the row-times-matrix products */
static inline void axpy(float a, const float* restrict w, float* restrict out, int n) {
  for (int j = 0; j < n; ++j) out[j] += a * w[j];
}

static inline float aggregate_k(const float* e, int K, int agg, float gamma) {
  if (agg == AGG_MAX) {
    float m = e[0];
    for (int k = 1; k < K; ++k) m = e[k] > m ? e[k] : m;
    return m;
  }
  if (agg == AGG_MEAN) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += e[k];
    return s / (float)K;
  }
  /* softmax: 1/gamma * logsumexp(gamma * e)  (mp_nn.py:80-83) */
  float mx = gamma * e[0];
  for (int k = 1; k < K; ++k) { float z = gamma * e[k]; mx = z > mx ? z : mx; }
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += expf(gamma * e[k] - mx);
  return (1.0f / gamma) * (logf(s) + mx);
}

int fgnn_oracle_mp_forward(const float* x, const int64_t* idx, const float* etype,
                           const float* filters, const float* bias, const float* bn_w,
                           const float* bn_b, const float* bn_mean, const float* bn_var,
                           int B, int N, int M, int K, int C, int O, int T, int extension,
                           int agg, int act, int has_bn, float gamma, float eps, float* out,
                           int threads) {
  const int OT = O * T;
  const int Kout = agg == AGG_NONE ? K : 1;
  if (extension != EXT_NONE && M != N) return -3;
  /* index validation: the reference's gather raises on out-of-range indices */
  {
    int bad = 0;
    const int64_t total = (int64_t)B * M * K;
    for (int64_t i = 0; i < total; ++i) bad |= (idx[i] < 0) | (idx[i] >= N);
    if (bad) return -2;
  }
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  /* node-major copy of x: mp_nn.py:125 / :138  x.permute(0,2,3,1).contiguous() */
  float* xt = (float*)malloc((size_t)B * N * C * sizeof(float));
  if (!xt) return -1;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < C; ++c) xt[((size_t)b * N + n) * C + c] = x[((size_t)b * C + c) * N + n];

  float* H = NULL;
  if (extension == EXT_NONE) {
    /* mp_nn.py:127-129  H[b,n,:] = x[b,:,n] @ filters */
    H = (float*)malloc((size_t)B * N * OT * sizeof(float));
    if (!H) { free(xt); return -1; }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)B * N; ++r) {
      float* h = H + (size_t)r * OT;
      memset(h, 0, (size_t)OT * sizeof(float));
      const float* xr = xt + (size_t)r * C;
      for (int c = 0; c < C; ++c) axpy(xr[c], filters + (size_t)c * OT, h, OT);
    }
  }

  float scale_buf[1];
  (void)scale_buf;
#pragma omp parallel
  {
    float* hrow = (float*)malloc((size_t)OT * sizeof(float));      /* per-slot O*T row (ext. modes) */
    float* cat = (float*)malloc((size_t)2 * C * sizeof(float));
    float* e = (float*)malloc((size_t)K * O * sizeof(float));      /* e[k*O + o] */
    float* ek = (float*)malloc((size_t)K * sizeof(float));
#pragma omp for schedule(static)
    for (int64_t g = 0; g < (int64_t)B * M; ++g) {
      const int b = (int)(g / M), m = (int)(g % M);
      for (int k = 0; k < K; ++k) {
        const int64_t n = idx[((size_t)b * M + m) * K + k];
        const float* hsrc;
        if (extension == EXT_NONE) {
          hsrc = H + ((size_t)b * N + n) * OT;                      /* mp_nn.py:92-113 gather */
        } else {
          const float* xi = xt + ((size_t)b * N + m) * C;
          const float* xj = xt + ((size_t)b * N + n) * C;
          for (int c = 0; c < C; ++c) {
            cat[c] = xi[c];
            cat[C + c] = extension == EXT_DIFF ? xi[c] - xj[c] : xj[c];   /* mp_nn.py:142-149 */
          }
          memset(hrow, 0, (size_t)OT * sizeof(float));
          for (int c = 0; c < 2 * C; ++c) axpy(cat[c], filters + (size_t)c * OT, hrow, OT);
          hsrc = hrow;
        }
        /* bmm with the edge-type vector: mp_nn.py:133-134 / :155-159 */
        for (int o = 0; o < O; ++o) {
          float s = 0.f;
          for (int t = 0; t < T; ++t)
            s += hsrc[o * T + t] * etype[(((size_t)b * T + t) * M + m) * K + k];
          e[k * O + o] = s;
        }
      }
      for (int o = 0; o < O; ++o) {
        for (int kk = 0; kk < Kout; ++kk) {
          float v;
          if (agg == AGG_NONE) {
            v = e[kk * O + o];
          } else {
            for (int k = 0; k < K; ++k) ek[k] = e[k * O + o];
            v = aggregate_k(ek, K, agg, gamma);
          }
          if (bias) v += bias[o];                                                /* :165-168 */
          if (has_bn) v = (v - bn_mean[o]) / sqrtf(bn_var[o] + eps) * bn_w[o] + bn_b[o]; /* :169-170 */
          if (act == 1) v = v > 0.f ? v : 0.f;                                   /* :172-173 */
          out[(((size_t)b * O + o) * M + m) * Kout + kk] = v;
        }
      }
    }
    free(hrow); free(cat); free(e); free(ek);
  }
  free(H);
  free(xt);
  return 0;
}

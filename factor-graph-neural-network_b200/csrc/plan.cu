// Native graph preprocessing for the hot path: what the reference's table builders leave to the
// caller -- dense rectangular nn_idx [B,M,K] tables (lib/data/ldpc_dataset.py:92-106, train_syn_hop_factor.py:112-151)
// -- turned into the structures the kernels want, on the host, in O(E) counting-sort passes:
//
//   fgnn_plan_build_host      the source-stationary plan of a table (include/fgnn_b200.h: src_ptr / slot_edge /
//                             edge_slot / src_rows with virtual rows of at most row_cap edges), stable in slot order
//   fgnn_locality_order_host  factors renumbered by their smallest variable (contiguous shards then cut few edges
//                             when the graph has index locality) and the variable-side table rewritten to match
//
// Host code only (tables usually arrive from a DataLoader on the CPU); the torch-op builders in mp_nn.py / graphs.py
// produce the same arrays for tables that already live on the device (tests compare the two).
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "common.cuh"

extern "C" {

// Sizes first (pass out == NULL pointers), then fill: returns the number of virtual rows V (>= B*n_src) or a negative
// fgnn_status.  n_edges_out receives E.  Arrays: src_ptr [V+1], slot_edge [B*M*K], edge_slot [E], src_rows [V - B*n_src].
int64_t fgnn_plan_build_host(const void* idx, int32_t idx_dtype, int32_t B, int32_t M, int32_t K, int64_t idx_sb, int32_t n_src,
                             int32_t row_cap, int64_t* n_edges_out, int32_t* src_ptr, int32_t* slot_edge, int32_t* edge_slot,
                             int32_t* src_rows) {
  if (!idx || B <= 0 || M <= 0 || K <= 0 || n_src <= 0 || row_cap <= 0) return FGNN_ERR_INVALID_ARG;
  const int64_t R = (int64_t)B * n_src, slots = (int64_t)B * M * K, MK = (int64_t)M * K;
  if (R >= INT32_MAX || slots >= INT32_MAX) return FGNN_ERR_UNSUPPORTED;
  auto at = [&](int64_t s) -> int64_t {
    const int64_t b = s / MK, mk = s - b * MK;
    const int64_t off = b * idx_sb + mk;
    return idx_dtype == FGNN_I64 ? reinterpret_cast<const int64_t*>(idx)[off] : (int64_t) reinterpret_cast<const int32_t*>(idx)[off];
  };
  std::vector<int32_t> count(R + 1, 0);
  int64_t E = 0;
  for (int64_t s = 0; s < slots; ++s) {
    const int64_t n = at(s);
    if (n >= 0 && n < n_src) { ++count[(s / MK) * n_src + n]; ++E; }
  }
  // extra virtual rows per source row, and where they start
  std::vector<int64_t> extra_base(R + 1, 0);
  int64_t n_extra = 0;
  for (int64_t g = 0; g < R; ++g) {
    extra_base[g] = n_extra;
    const int32_t c = count[g];
    n_extra += c > row_cap ? (c + row_cap - 1) / row_cap - 1 : 0;
  }
  const int64_t V = R + n_extra;
  if (n_edges_out) *n_edges_out = E;
  if (!src_ptr || !slot_edge || !edge_slot) return V;              // sizing call
  if (V >= INT32_MAX) return FGNN_ERR_UNSUPPORTED;
  // edges per virtual row -> src_ptr
  src_ptr[0] = 0;
  for (int64_t g = 0; g < R; ++g) src_ptr[g + 1] = std::min(count[g], row_cap);
  for (int64_t g = 0; g < R; ++g) {
    int32_t left = count[g] - row_cap;
    int64_t v = R + extra_base[g];
    while (left > 0) {
      src_ptr[v + 1] = std::min(left, row_cap);
      if (src_rows) src_rows[v - R] = (int32_t)g;
      left -= row_cap;
      ++v;
    }
  }
  for (int64_t v = 0; v < V; ++v) src_ptr[v + 1] += src_ptr[v];
  // stable placement: slots in ascending order, the r-th edge of source row g goes to virtual row
  // (r < cap ? g : R + extra_base[g] + r / cap - 1), position r % cap
  std::vector<int32_t> seen(R, 0);
  for (int64_t s = 0; s < slots; ++s) {
    const int64_t n = at(s);
    if (n < 0 || n >= n_src) { slot_edge[s] = -1; continue; }
    const int64_t g = (s / MK) * n_src + n;
    const int32_t r = seen[g]++;
    const int64_t v = r < row_cap ? g : R + extra_base[g] + r / row_cap - 1;
    const int32_t e = src_ptr[v] + r % row_cap;
    slot_edge[s] = e;
    edge_slot[e] = (int32_t)s;
  }
  return V;
}

// order [F]: new factor i = old factor order[i] (ascending smallest variable, stable); idx_v2f_out [F,K] rows permuted;
// idx_f2v_out [N,Kv]: entries renumbered (pad slots -- pad[i] != 0 -- are set to 0, the reference's valid-index pad).
int fgnn_locality_order_host(const int64_t* idx_v2f, int64_t F, int32_t K, const int64_t* idx_f2v, const uint8_t* pad, int64_t N,
                             int32_t Kv, int64_t* order, int64_t* idx_v2f_out, int64_t* idx_f2v_out) {
  if (!idx_v2f || !idx_f2v || !order || !idx_v2f_out || !idx_f2v_out || F <= 0 || K <= 0 || N <= 0 || Kv <= 0) return FGNN_ERR_INVALID_ARG;
  std::vector<int64_t> key(F);
  for (int64_t f = 0; f < F; ++f) key[f] = *std::min_element(idx_v2f + f * K, idx_v2f + (f + 1) * K);
  std::iota(order, order + F, (int64_t)0);
  std::stable_sort(order, order + F, [&](int64_t a, int64_t b) { return key[a] < key[b]; });
  std::vector<int64_t> inv(F);
  for (int64_t i = 0; i < F; ++i) {
    inv[order[i]] = i;
    memcpy(idx_v2f_out + i * K, idx_v2f + order[i] * K, sizeof(int64_t) * K);
  }
  for (int64_t i = 0; i < N * Kv; ++i) {
    const int64_t f = idx_f2v[i];
    idx_f2v_out[i] = (pad && pad[i]) ? 0 : (f >= 0 && f < F ? inv[f] : f);
  }
  return FGNN_OK;
}

}  // extern "C"

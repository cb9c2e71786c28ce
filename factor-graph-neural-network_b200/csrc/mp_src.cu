// Source-stationary evaluation of the FGNN message-passing call on tcgen05 (NO_EXTENSION, C = 64, fp32).
//
//   out[b,o,m] = act(BN(bias[o] + AGG_k sum_t etype[b,t,m,k] * (x[b, idx[b,m,k], :] . W[:, o*T+t])))
//   reference: lib/model/mpnn/mp_nn.py:115-175
//
// The reference computes H = x W once per SOURCE node (mp_nn.py:124-127) and gathers the O*T-wide rows
// of H per slot (mp_nn.py:92-113, 128-134); the destination-stationary kernel (mp_tc.cu) never
// materialises H but recomputes x[n] W once per slot -- E row-products instead of N (6x for the
// variable->factor call of a pairwise type whose variables each sit in six factors).  At T = 16 that
// kernel runs at the tensor pipe's sustained rate, so the only way forward is fewer row-products:
//
//   pass 1 (mp_src_kernel):  a tile = 128 consecutive VIRTUAL source rows.  A virtual row is a source row
//            with at most RC of its out-edges (the plan splits rows with more edges -- the reference's pad
//            target collects every padded slot -- into several virtual rows, so no thread ever walks a long
//            edge list).  H = x W for the tile is accumulated in TMEM exactly as in mp_tc.cu (split-bf16,
//            three MMA terms, stationary filter slice); the epilogue thread of a row contracts its H chunk
//            with the edge-type vectors of the row's edges and stores the O-wide MESSAGE of every edge --
//            H itself (O*T wide) still never leaves the SM.
//   pass 2 (mp_reduce_kernel): every destination aggregates the messages of its K slots (slot_edge maps a
//            slot to its edge), then bias / eval-BN / activation -- pure streaming.
//
// Messages cost 2 * 4*O bytes per edge of extra HBM traffic; in exchange the tensor work drops from E to N
// row-products.  The caller chooses (fgnn_mp_args.src_ptr != NULL); the results are bit-identical to the
// destination-stationary kernel (same MMA terms per row, same contraction order, same aggregation order).
//
// What bounds pass 1 (measured, profiles/r01_full.txt -> r02): tensor memory reads (~32 B/clk per SM
// sub-partition: a 128 x 512 fp32 accumulator tile takes 2048 clk to read ONCE, against 3072 clk of MMA), the
// L1 wavefront rate of row-per-thread global accesses (one wavefront per lane: the round-1 kernel spent 9 000
// of them per tile on 16-byte message stores and edge-type loads) and shared-memory bandwidth (the B operand
// alone takes half of it).  Hence:
//   * two epilogue warps share a TMEM lane quadrant.  ES = false (tables with <= 3 edges per row): they
//     ALTERNATE over the accumulator chunks, so every accumulator element is read once.  ES = true (up to 6
//     edges per row, e.g. variables of a pairwise type): both read every chunk and split the row's edges --
//     that call has few tiles and is bound by its message traffic, not by the tensor pipe;
//   * the edge-type vectors of a tile are staged by the loader warps with COALESCED cp.async into a padded
//     shared-memory layout ([row][slot][T] with an odd row pitch: row-per-thread reads are conflict-free) and
//     live in registers for the whole tile;
//   * a message leaves as whole 32-byte sectors (one 256-bit store per edge and chunk).
//
// CTA roles (512 threads): warps 0-7 epilogue, 8-11 converters (raw fp32 row -> split-bf16 A stage in TMEM),
// 12-13 loaders (x rows and edge types, cp.async), 14 MMA.
#include "tc_common.cuh"

namespace fgnn {

#ifdef FGNN_TC_TRACE
// Debug builds only (-DFGNN_TC_TRACE): per-tile timestamps of CTA 0, read back by tools/src_trace.py.
__device__ unsigned long long g_src_trace[16 * 4096];
#define SRC_TRACE(item, slot)                                                                   \
  do {                                                                                          \
    if (blockIdx.x == 0 && (item) < 4096u && (threadIdx.x & 31) == 0) {                         \
      unsigned long long _t;                                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                    \
      g_src_trace[(item) * 16 + (slot)] = _t;                                                   \
    }                                                                                           \
  } while (0)
extern "C" int fgnn_debug_src_trace_read(unsigned long long* host, size_t count) {
  return cudaMemcpyFromSymbol(host, g_src_trace, count * sizeof(unsigned long long)) == cudaSuccess ? 0 : -6;
}
#else
#define SRC_TRACE(item, slot) do { } while (0)
#endif

namespace {

using namespace tc;

constexpr int kEB = 3;                 // edges per epilogue thread

struct SrcParams {
  const float* x;                      // [rows, 64] source rows, node-major
  const int32_t* vptr;                 // [vrows + 1]: edges of virtual row v are vptr[v] .. vptr[v+1]-1 (at most RC)
  const int32_t* xrow;                 // [vrows - rows]: source row of virtual row v >= rows (v < rows: itself)
  const float* et_edges;               // [E, T]
  float* msg;                          // [E, O]
  uint32_t rows;                       // B * N
  uint32_t vrows;                      // virtual rows (>= rows)
  int O, T;
  // fused aggregation (FUSE): a tile = ONE batch element, whose destinations only read its own source rows
  const int32_t* edge_slot;            // [E] slot (b*M + m)*K + k of every edge
  uint32_t rpt;                        // source rows per tile: N (<= 128) when fused, 128 otherwise
  int M, K;
};

// per-slot message staging of the fused path: one region per accumulator chunk, [M*K slots][channels of the chunk],
// the 16-byte pieces of a slot's row XOR-swizzled with the slot number (source threads write to scattered slots; the
// reader threads of a slot read its row whole)
__host__ __device__ constexpr int src_fuse_staging_max() { return 74 * 1024; }

// raw x stages by mode: the edge-splitting mode stages twice the edge types and its tiles take longer
#ifdef FGNN_SRC_DBG_NST1
__host__ __device__ constexpr int src_x_stages(bool es) { return 1; }
#else
__host__ __device__ constexpr int src_x_stages(bool es) { return es ? 1 : 2; }
#endif

// Edge-type image: edge-major [E][T] floats, but the T/4 16-byte pieces of edge e are stored at piece position
// j ^ et_key(e).  A tile's block is brought into shared memory as it is (one bulk copy) and read row-per-thread;
// with the key a warp's reads (edges a fixed stride apart) spread over the eight 16-byte bank groups.
template <int T>
__host__ __device__ __forceinline__ uint32_t et_key(uint32_t e) {
  constexpr uint32_t PPE = T / 4;                            // pieces per edge: 4 | 2 | 1
  return PPE == 1 ? 0u : (e / (8u / PPE)) & (PPE - 1u);      // edges per 128 bytes: 8 / PPE
}

// wait for this thread's outstanding tcgen05.ld; the registers of the load pass through the statement, so their
// consumers cannot be scheduled above it
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                 "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// a real instruction that reads `v` (a compare against one NaN payload, guarding a harmless nanosleep): the warp
// cannot run past it before the load that produces `v` has completed
__device__ __forceinline__ void touch_reg(float v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %0, 0x7fc5a5a5;\n\t@p nanosleep.u32 1;\n\t}" ::"r"(__float_as_uint(v)) : "memory");
}

__device__ __forceinline__ void stg256(float* dst, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// T edge types, NCH accumulator chunks of 128 columns per CTA, ES: the two warps of a lane quadrant split the
// row's EDGES (row cap 6) instead of alternating over the chunks (row cap 3)
template <int T, int NCH, bool ES, bool FUSE>
__global__ void __launch_bounds__(tc::kThreads, 1)
mp_src_kernel(const SrcParams p, const MpParams mp, const uint8_t* __restrict__ wimg, const int S, const int n_workers,
              const int n_tiles) {
  constexpr int NC = 128;                      // accumulator columns per chunk
  constexpr int COLS = NC * NCH;               // columns of W this CTA owns
  constexpr int CPC = NC / T;                  // output channels per chunk
  constexpr int CPH = 64 / T;                  // output channels per 64-column piece
  constexpr int RC = ES ? 2 * kEB : kEB;       // edges per virtual row (plan contract)
  constexpr int EB = T * 4;                    // bytes of one edge-type vector
  constexpr int ETB = kTileM * RC * EB;        // edge-type staging: the tile's (at most 128 RC) vectors, contiguous
  constexpr int NST = src_x_stages(ES);        // raw x stages (a tile's rows must be in flight while the previous one converts)
  constexpr int ETS = FUSE ? 2 : 1;            // edge-type staging stages (fused tiles are short: the copy needs a tile's head start)
  constexpr int ROWB = row_bytes(false), STAGEB = stage_bytes(false);
  static_assert(16 % T == 0 && T >= 4 && CPH % 4 == 0 && NCH % 2 == 0, "unsupported shape");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                        // [2 parts][COLS rows][128 B]  UMMA K-major SW128
  uint8_t* sA = sB + w_bytes(COLS, false);                   // [NST][128 rows][256 B]      raw x tiles
  uint8_t* sEt = sA + NST * STAGEB;                          // [<= 128 RC][T]              edge types of the tile's edges
  uint8_t* sMsg = sEt + ETS * ETB;                           // FUSE: [M*K slots][O floats]  messages of the tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMsg + (FUSE ? src_fuse_staging_max() : 0));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  const uint32_t bar0 = smem_u32(bars);
  // barrier slots of tc_common.cuh: the raw ring has at most two stages here, a spare slot carries the edge-type staging
  auto raw_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto raw_empty = [&](uint32_t s) { return bar0 + 8u * (kMaxAStages + s); };
  auto et_full = [&](uint32_t s) { return bar0 + 8u * (2 + s); };
  auto et_empty = [&](uint32_t s) { return bar0 + 8u * (kMaxAStages + 2 + s); };
  auto ta_full = [&](uint32_t s) { return bar0 + 8u * (2 * kMaxAStages + s); };
  auto ta_empty = [&](uint32_t s) { return bar0 + 8u * (2 * kMaxAStages + kTA + s); };
  auto t_full = [&](uint32_t s) { return bar0 + 8u * (2 * kMaxAStages + 2 * kTA + s); };
  auto t_empty = [&](uint32_t s) { return bar0 + 8u * (2 * kMaxAStages + 2 * kTA + kAcc + s); };
  const uint32_t w_full = bar0 + 8u * (2 * kMaxAStages + 2 * kTA + 2 * kAcc);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x % S, worker = blockIdx.x / S;
  const int col0 = split * COLS;                             // first W column of this CTA
  const int ch0 = col0 / T;                                  // first output channel of this CTA

  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(raw_full(s), kGatherWarps * 32);
      mbar_init(raw_empty(s), 128);
    }
    for (int s = 0; s < ETS; ++s) {
      mbar_init(et_full(s), 1);                              // one arrive.expect_tx + the bytes of the bulk copy
      mbar_init(et_empty(s), kEpiWarps * 32);
    }
    for (int s = 0; s < kTA; ++s) {
      mbar_init(ta_full(s), 128);
      mbar_init(ta_empty(s), 1);
    }
    for (int s = 0; s < kAcc; ++s) {
      mbar_init(t_full(s), 1);
#ifdef FGNN_SRC_DBG_NOALT
      mbar_init(t_empty(s), kEpiWarps * 32);
#else
      mbar_init(t_empty(s), ES ? kEpiWarps * 32 : kEpiWarps * 16);      // the readers of one chunk
#endif
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // =====================================================================================
    // EPILOGUE: thread (eg, r) works on virtual row tile*128 + r == TMEM lane r.
    //   ES = false: all (<= 3) edges of the row, accumulator chunks eg, eg+2, ...
    //   ES = true:  edges eg, eg+2, eg+4 of the row (<= 6), every chunk
    // =====================================================================================
    reg_inc<kRegEpi>();
    const int eg = warp >> 2, wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
    const uint32_t sEt_u = smem_u32(sEt);
    // first edge of the tile (= of its row 0) and this row's edge range
    auto edge_range = [&](int tile, int32_t& et0, int32_t& e0, int32_t& e1) {
      const uint32_t v = (uint32_t)tile * p.rpt + r;
      et0 = e0 = e1 = 0;
      if (tile < n_tiles) {
        et0 = __ldg(p.vptr + (uint32_t)tile * p.rpt);
        if ((uint32_t)r < p.rpt && v < p.vrows) { e0 = __ldg(p.vptr + v); e1 = __ldg(p.vptr + v + 1); }
      }
    };
    // edge ranges are fetched TWO tiles ahead, the slots of a tile's edges one tile ahead: no tile starts with a load
    // that depends on a load of the same iteration
    int32_t et0n, e0n, e1n, et0f, e0f, e1f;
    edge_range(worker, et0n, e0n, e1n);
    edge_range(worker + n_workers, et0f, e0f, e1f);
    int32_t sl_nx[kEB] = {0, 0, 0};
    // FUSE: a thread of the aggregation phase always handles the same four channels (128 threads, CPC/4 pieces per row):
    // their bias / BN scale / BN shift live in registers for the whole kernel
    float e_bi[4] = {0.f, 0.f, 0.f, 0.f}, e_sc[4] = {1.f, 1.f, 1.f, 1.f}, e_sh[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (FUSE) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int oc = eg * CPC + ((tid & 127) % (CPC / 4)) * 4 + c;
        if (mp.bias) e_bi[c] = mp.bias[oc];
        if (mp.scale) { e_sc[c] = mp.scale[oc]; e_sh[c] = mp.shift[oc]; }
      }
    }
    if constexpr (FUSE) {
      static_assert(!FUSE || !ES, "the fused path alternates over the chunks (row cap 3)");
#pragma unroll
      for (int i = 0; i < kEB; ++i) sl_nx[i] = e0n + i < e1n ? __ldg(p.edge_slot + e0n + i) : 0;
    }
    uint32_t it = 0;
    bool waited = false;
    float* const msg_base = p.msg + ch0;
    for (int tile = worker; tile < n_tiles; tile += n_workers, ++it) {
      const int32_t e0 = e0n, ne = e1n - e0n, el0 = e0n - et0n;        // el0: the row's first edge within the tile
      et0n = et0f; e0n = e0f; e1n = e1f;                     // the next tile's range (loaded a tile ago) ...
      edge_range(tile + 2 * n_workers, et0f, e0f, e1f);      // ... and the one after it goes in flight
      // my edges: slot ks(i) = ES ? eg + 2 i : i of the row, i < n_mine
      const int n_mine = ES ? (ne > eg ? (ne - eg + 1) >> 1 : 0) : ne;
      const int n_warp = __reduce_max_sync(0xffffffffu, n_mine);
      // edge-type vectors of my edges: staging -> registers, for the whole tile.  The image is edge-major with the
      // 16-byte pieces of edge e XOR-swizzled by et_key(e) (et_permute_kernel), so that the row-per-thread reads of
      // a warp spread over the banks.
      if ((warp & 3) == 0) SRC_TRACE(it, 6 + 4 * eg);         // epilogue: tile start
      // FUSE: where my edges' messages go -- their slot inside this tile's (batch element's) table; the slots of the
      // NEXT tile's edges are fetched now (their range is already known), so no tile starts with a dependent load
      int32_t sl[kEB];
      if constexpr (FUSE) {
#pragma unroll
        for (int i = 0; i < kEB; ++i) sl[i] = sl_nx[i] - tile * (p.M * p.K);
#pragma unroll
        for (int i = 0; i < kEB; ++i) sl_nx[i] = e0n + i < e1n ? __ldg(p.edge_slot + e0n + i) : 0;
      }
      const uint32_t es = it % ETS;
      mbar_wait(et_full(es), (it / ETS) & 1);
      if ((warp & 3) == 0) SRC_TRACE(it, 7 + 4 * eg);         // edge types landed
      float et[kEB][T];
#pragma unroll
      for (int i = 0; i < kEB; ++i) {
        const int ks = ES ? eg + 2 * i : i;
        if (i < n_mine) {
          const uint32_t key = et_key<T>((uint32_t)(e0 + ks));
#pragma unroll
          for (int t4 = 0; t4 < T / 4; ++t4) {
            const float4 v = lds_f4(sEt_u + es * ETB + (uint32_t)(el0 + ks) * EB + (((uint32_t)t4 ^ key) << 4));
            et[i][4 * t4] = v.x; et[i][4 * t4 + 1] = v.y; et[i][4 * t4 + 2] = v.z; et[i][4 * t4 + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int t = 0; t < T; ++t) et[i][t] = 0.f;
        }
      }
      // The staging goes back to the loader once every loaded value has ARRIVED in its register: an arrive issued
      // right behind the loads can overtake them -- they queue behind the burst of all eight warps and the MMA's
      // operand reads -- and the next tile's bulk copy (async proxy) then overwrites what the last loads have yet to
      // read (measured: the third edge's vector of a few rows per launch).  Shared-memory loads of a warp complete in
      // order, so touching the last piece of every vector is enough.
#pragma unroll
      for (int i = 0; i < kEB; ++i) touch_reg(et[i][T - 1]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(et_empty(es));                             // the staging stage may take another tile's edge types
      if (!waited) { pdl_wait(); waited = true; }            // first store: the preceding launch may still read msg
#pragma unroll 1
#ifdef FGNN_SRC_DBG_NOALT
      for (int chunk = 0; chunk < NCH; chunk += 1) {
#else
      for (int chunk = ES ? 0 : eg; chunk < NCH; chunk += ES ? 1 : 2) {
#endif
        const uint32_t ct = it * NCH + chunk;
        const uint32_t st = ct % kAcc;
        mbar_wait(t_full(st), (ct / kAcc) & 1);
        tc_fence_after();
        if ((warp & 3) == 0 && chunk < 2) SRC_TRACE(it, 8 + 4 * eg);      // first accumulator chunk of the tile ready
        const uint32_t taddr = lane_addr + st * kAccCols;
        // The chunk comes to registers in four quarters of 32 columns, double-buffered: quarter q+1 is in flight
        // while quarter q is contracted (tensor-memory reads are the scarce resource here).  A message leaves as
        // whole 32-byte sectors: eight channels = QPS quarters.
        constexpr int CPQ = 32 / T;                          // channels per quarter: 2 | 4 | 8
        constexpr int QPS = 8 / CPQ;                         // quarters per 32-byte store: 4 | 2 | 1
        float o[kEB][8];
        uint32_t d[2][32];
        tmem_ld32(taddr, d[0]);
        tmem_ld_wait_on(d[0]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q + 1 < 4) tmem_ld32(taddr + (q + 1) * 32, d[(q + 1) & 1]);
#pragma unroll
          for (int i = 0; i < kEB; ++i) {
            if (i < n_warp) {                                // warp-uniform
              if constexpr (T == 4) {                        // channel pairs (tc_common.cuh), as in mp_tc.cu: bit-identical
                uint64_t et2[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) et2[t] = pack2(et[i][t], et[i][t]);
#pragma unroll
                for (int c2 = 0; c2 < CPQ / 2; ++c2)
                  unpack2(contract_pair4(et2, &d[q & 1][c2 * 8]), o[i][(q % QPS) * CPQ + 2 * c2], o[i][(q % QPS) * CPQ + 2 * c2 + 1]);
              } else {
#pragma unroll
                for (int c = 0; c < CPQ; ++c)                // channel c of this quarter: columns c*T .. c*T+T-1
                  o[i][(q % QPS) * CPQ + c] = contract_types<T>(et[i], &d[q & 1][c * T]);      // same function as mp_tc.cu: bit-identical
              }
            }
          }
          if ((q + 1) % QPS == 0) {
#pragma unroll
            for (int i = 0; i < kEB; ++i) {
              if (i < n_mine) {
                const int ks = ES ? eg + 2 * i : i;
                if constexpr (FUSE) {
                  // eight channels = two 16-byte pieces of the slot's row in this chunk's region (CPC channels per row),
                  // piece index XOR (slot & 7)
                  constexpr uint32_t PPR = CPC / 4;                     // pieces per row: 8 at T = 4
                  const uint32_t pc = (uint32_t)(q / QPS) * 2, sk = (uint32_t)sl[i] & (PPR - 1);
                  const uint32_t row = smem_u32(sMsg) + (uint32_t)chunk * (uint32_t)(p.M * p.K) * (CPC * 4) + (uint32_t)sl[i] * (CPC * 4);
                  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((pc ^ sk) << 4)), "f"(o[i][0]), "f"(o[i][1]), "f"(o[i][2]), "f"(o[i][3]) : "memory");
                  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((pc + 1) ^ sk) << 4)), "f"(o[i][4]), "f"(o[i][5]), "f"(o[i][6]), "f"(o[i][7]) : "memory");
                } else {
                  stg256(msg_base + (int64_t)(e0 + ks) * p.O + chunk * CPC + (q / QPS) * 8, o[i]);
                }
              }
            }
          }
          if (q + 1 < 4) tmem_ld_wait_on(d[(q + 1) & 1]);
          if (q == 2) {                                      // the last read of this accumulator stage has landed
            tc_fence_before();
            mbar_arrive(t_empty(st));
          }
        }
      }
      if constexpr (FUSE) {
        // ---- fused pass 2, per warp group and chunk: group eg has staged the messages of chunk eg (its CPC channels of
        // every slot); its 128 threads now aggregate every destination's K slots for those channels (slot order, like
        // the other kernels), apply bias / eval-BN / activation and store.  The two groups never wait for each other,
        // so one group's tensor-memory reads overlap the other's aggregation.  (NCH == 2: chunk == eg.)
        static_assert(!FUSE || (NCH == 2 && 128 % (CPC / 4) == 0), "one chunk per warp group; fixed channels per thread");
        constexpr int PPR = CPC / 4;                         // 16-byte pieces per staged row
        named_bar_sync(1 + eg, 4 * 32);                      // the group's messages of this tile are staged
        const int tasks = p.M * PPR, gt = tid & 127;
        const float neg = mp.act == FGNN_ACT_NONE ? 1.f : (mp.act == FGNN_ACT_RELU ? 0.f : mp.slope);
        const uint32_t region = smem_u32(sMsg) + (uint32_t)eg * (uint32_t)(p.M * p.K) * (CPC * 4);
        for (int task = gt; task < tasks; task += 128) {
          const int m = task / PPR, c4 = task - m * PPR;
          float y[4];
          const uint32_t slot0 = (uint32_t)(m * p.K);
          auto piece = [&](int k) {
            const uint32_t slot = slot0 + (uint32_t)k;
            return lds_f4(region + slot * (CPC * 4) + ((((uint32_t)c4) ^ (slot & (PPR - 1))) << 4));
          };
          if (mp.agg == FGNN_AGG_MAX) {                        // the FGNN aggregator: a lean loop of its own
            float4 a = piece(0);
            for (int k = 1; k < p.K; ++k) {
              const float4 v = piece(k);
              a.x = fmaxf(a.x, v.x); a.y = fmaxf(a.y, v.y); a.z = fmaxf(a.z, v.z); a.w = fmaxf(a.w, v.w);
            }
            y[0] = a.x; y[1] = a.y; y[2] = a.z; y[3] = a.w;
          } else {
            float a[4], sx[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { a[c] = mp.agg == FGNN_AGG_MEAN ? 0.f : -INFINITY; sx[c] = 0.f; }
            for (int k = 0; k < p.K; ++k) {
              const float4 v4 = piece(k);
              const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (mp.agg == FGNN_AGG_SOFTMAX) softmax_push(a[c], sx[c], v[c], mp.gamma);
                else a[c] += v[c];
              }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
              y[c] = mp.agg == FGNN_AGG_SOFTMAX ? softmax_finish(a[c], sx[c], mp.gamma) : __fmul_rn(a[c], __frcp_rn((float)p.K));
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float vv = fmaf(y[c] + e_bi[c], e_sc[c], e_sh[c]);
            y[c] = vv >= 0.f ? vv : vv * neg;
          }
          float* dst = mp.out + ((int64_t)tile * p.M + m) * mp.o_sm + eg * CPC + c4 * 4;
          if (mp.accumulate) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]) : "memory");
          } else {
            *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
          }
        }
        named_bar_sync(1 + eg, 4 * 32);                      // the region may take the next tile's messages
      }
      if ((warp & 3) == 0) SRC_TRACE(it, 9 + 4 * eg);         // tile done
    }
  } else if (warp < kGatherWarp0) {
    // =====================================================================================
    // CONVERTERS: raw fp32 row -> bf16 hi/lo pairs -> A stage in tensor memory (as mp_tc.cu)
    // =====================================================================================
    reg_dec<kRegConv>();
    const int cr = tid - kConvWarp0 * 32;
    const uint32_t row_u = smem_u32(sA) + (uint32_t)cr * ROWB;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kTACol0;
    uint32_t i = 0;
    for (int tile = worker; tile < n_tiles; tile += n_workers, ++i) {
      const uint32_t st = i % NST, use = i / NST;
      const uint32_t ta = i % kTA, tuse = i / kTA;
      mbar_wait(raw_full(st), use & 1);
      if (cr < 32) SRC_TRACE(i, 2);                           // x rows landed
      float4 v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = lds_f4(row_u + st * STAGEB + (uint32_t)((c ^ (cr & 15)) * 16));
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
      mbar_arrive(raw_empty(st));
      mbar_wait(ta_empty(ta), (tuse & 1) ^ 1);
      tc_fence_after();
      tmem_st32(lane_addr + ta * kTACols, hi);
      tmem_st32(lane_addr + ta * kTACols + 32, lo);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ta_full(ta));
      if (cr < 32) SRC_TRACE(i, 3);                           // A stage written
    }
  } else {
  reg_dec<kRegAux>();                                        // warps 12-15 together (one warpgroup)
  if (warp < kMmaWarp) {
    // =====================================================================================
    // X LOADERS (two warps, 64 rows each): the x rows of the tile's 128 virtual rows (consecutive source rows for
    // v < rows, xrow[] beyond) into a raw stage in the converter's swizzled layout (chunk q of row rr at position
    // q ^ (rr & 15)), 16 lanes per row
    // =====================================================================================
    constexpr int LPR = ROWB / 16, RPW = 32 / LPR, ROWS_W = kTileM / kGatherWarps, NIT = ROWS_W / RPW, NDO = LPR / RPW;
    const int pw = warp - kGatherWarp0;
    const int sub = lane / LPR, q = lane % LPR;
    const uint8_t* xq = reinterpret_cast<const uint8_t*>(p.x) + q * 16;
    const uint32_t sA_u = smem_u32(sA);
    uint32_t dst_off[NDO];
#pragma unroll
    for (int c = 0; c < NDO; ++c)
      dst_off[c] = (uint32_t)(pw * ROWS_W + RPW * c + sub) * ROWB + (uint32_t)((q ^ ((RPW * c + sub) & (LPR - 1))) * 16);
    // per tile and lane: the x row of virtual rows pw*64 + lane and pw*64 + 32 + lane; loaded one tile ahead
    struct Meta { uint32_t xr0, xr1; };
    auto meta_of = [&](int tile, Meta& m) {
      m.xr0 = m.xr1 = 0xffffffffu;
      if (tile >= n_tiles) return;
      const uint32_t l0 = (uint32_t)(pw * ROWS_W + lane), l1 = l0 + 32;          // rows of the tile
      const uint32_t v0 = (uint32_t)tile * p.rpt + l0, v1 = (uint32_t)tile * p.rpt + l1;
      if (l0 < p.rpt && v0 < p.vrows) m.xr0 = v0 < p.rows ? v0 : (uint32_t)__ldg(p.xrow + (v0 - p.rows));
      if (l1 < p.rpt && v1 < p.vrows) m.xr1 = v1 < p.rows ? v1 : (uint32_t)__ldg(p.xrow + (v1 - p.rows));
    };
    Meta cur, nxt;
    meta_of(worker, cur);
    pdl_wait();                                              // x is the preceding launch's output: order behind it
    uint32_t i = 0;
    for (int tile = worker; tile < n_tiles; tile += n_workers, ++i) {
      const uint32_t st = i % NST, use = i / NST;
      meta_of(tile + n_workers, nxt);
      mbar_wait(raw_empty(st), (use & 1) ^ 1);
      if (pw == 0) SRC_TRACE(i, 0);                           // x stage free
      const uint32_t stage = sA_u + st * STAGEB;
      const uint32_t off0 = cur.xr0 == 0xffffffffu ? 0xffffffffu : cur.xr0 * ROWB;     // rows * 256 < 2^32 (src_supported)
      const uint32_t off1 = cur.xr1 == 0xffffffffu ? 0xffffffffu : cur.xr1 * ROWB;
#pragma unroll
      for (int u0 = 0; u0 < NIT; u0 += 8) {
        uint32_t o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rl = RPW * (u0 + j);                     // + sub: this lane's row among the warp's 64
          o8[j] = __shfl_sync(0xffffffffu, rl < 32 ? off0 : off1, (rl + sub) & 31);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int u = u0 + j;
          const uint32_t dst = stage + dst_off[u % NDO] + (uint32_t)((u / NDO) * LPR * ROWB);
          const bool ok = o8[j] != 0xffffffffu;
          cp_async16(dst, xq + (ok ? o8[j] : 0u), ok ? 16u : 0u);
        }
      }
      cp_async_arrive_noinc(raw_full(st));
      if (pw == 0) SRC_TRACE(i, 1);                           // x copies issued
      cur = nxt;
    }
  } else if (warp == kMmaWarp + 1) {
    // =====================================================================================
    // EDGE-TYPE LOADER: the tile's edges are consecutive, so their vectors are ONE contiguous block of the
    // edge-major image: a single TMA bulk copy per tile, issued as soon as the epilogue has taken the previous
    // tile's vectors into registers
    // =====================================================================================
    if (lane == 0) {
      const uint32_t sEt_u = smem_u32(sEt);
      auto range = [&](int tile, int32_t& a, int32_t& b) {
        a = b = 0;
        if (tile >= n_tiles) return;
        const uint32_t v0 = (uint32_t)tile * p.rpt, v1 = v0 + p.rpt < p.vrows ? v0 + p.rpt : p.vrows;
        a = __ldg(p.vptr + v0); b = __ldg(p.vptr + v1);
      };
      int32_t a, b, an, bn;
      range(worker, a, b);
      uint32_t i = 0;
      for (int tile = worker; tile < n_tiles; tile += n_workers, ++i) {
        range(tile + n_workers, an, bn);
        const uint32_t es = i % ETS;
        mbar_wait(et_empty(es), ((i / ETS) & 1) ^ 1);
        SRC_TRACE(i, 14);                                    // edge-type staging free: bulk copy issued
        const uint32_t bytes = (uint32_t)(b - a) * EB;
        if (bytes) {
          mbar_expect_tx(et_full(es), bytes);
          bulk_g2s(sEt_u + es * ETB, reinterpret_cast<const uint8_t*>(p.et_edges) + (int64_t)a * EB, bytes, et_full(es));
        } else {
          mbar_arrive(et_full(es));
        }
        a = an; b = bn;
      }
    }
  } else if (warp == kMmaWarp) {
    // =====================================================================================
    // MMA ISSUER (as mp_tc.cu, one item per tile)
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
    const uint64_t desc_hi = umma_desc_sw128(0);
    uint32_t ct = 0, i = 0;
    for (int tile = worker; tile < n_tiles; tile += n_workers, ++i) {
      const uint32_t ta = i % kTA, tuse = i / kTA;
      mbar_wait(ta_full(ta), tuse & 1);
      SRC_TRACE(i, 4);                                       // A stage ready
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
            for (int ks = 0; ks < kC / 16; ++ks)
              umma_bf16_ts(d_tmem, a + ks * 8, desc_hi | (uint64_t)(bb + ks * 2), idesc, (term | ks) != 0);
          }
          umma_commit(t_full(ts));
          if (chunk == NCH - 1) umma_commit(ta_empty(ta));
        }
        __syncwarp();
        ++ct;
      }
      SRC_TRACE(i, 5);                                       // all chunks of the tile issued
#ifdef FGNN_TC_TRACE
      if (blockIdx.x == 0 && i < 4096u && lane == 0) g_src_trace[i * 16 + 15] = (unsigned long long)clock64();      // SM clock beside the wall clock
#endif
    }
  }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// pass 2: every destination aggregates the messages of its K slots, then bias / BN / activation.
// One thread per (destination row, 8 channels): the slot -> edge map of the row is read first, then all of a
// batch's 32-byte message pieces are in flight before the first is used; slots in order k = 0..K-1 like the
// other kernels.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const float* src, float (&v)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(src));
}

// slot_edge value of a slot that is present but known to carry an all-zero edge-type vector (the reference's padding:
// a valid index + zero edge type): its message is the constant 0 -- it takes part in the aggregate without having an
// edge, a virtual row or a stored message
constexpr int32_t kZeroSlot = -2;

// aggregate -> bias / eval BN / activation -> store (or accumulate) of eight channels of destination row g
template <int AGG>
__device__ __forceinline__ void reduce_finish(const MpParams& p, const float (&a)[8], const float (&s)[8], float live, int64_t g, int o,
                                              float neg) {
  float y[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float r;
    if (AGG == FGNN_AGG_MAX) r = a[c];
    else if (AGG == FGNN_AGG_SOFTMAX) r = live > 0.f ? softmax_finish(a[c], s[c], p.gamma) : -INFINITY;
    else r = __fmul_rn(a[c], live > 0.f ? __frcp_rn(live) : 0.f);
    const float bi = p.bias ? p.bias[o + c] : 0.f, sc = p.scale ? p.scale[o + c] : 1.f, sh = p.scale ? p.shift[o + c] : 0.f;
    float v = fmaf(r + bi, sc, sh);
    v = v >= 0.f ? v : v * neg;
    y[c] = r == -INFINITY ? r : v;
  }
  const int64_t orow = p.out_rows ? (int64_t)p.out_rows[g] : g;
  if (orow < 0) return;
  float* dst = p.out + orow * p.o_sm + o;
  if (p.accumulate) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(y[4]), "f"(y[5]), "f"(y[6]), "f"(y[7]) : "memory");
  } else {
    *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
  }
}

template <int AGG>
__global__ void __launch_bounds__(256)
mp_reduce_kernel(const MpParams p, const float* __restrict__ msg, const int32_t* __restrict__ slot_edge) {
  constexpr int KB = 6;                                      // slots per batch of loads
  const int O8 = p.O >> 3;
  const int64_t rows = (int64_t)p.B * p.M, total = rows * O8;
  const float neg = p.act == FGNN_ACT_NONE ? 1.f : (p.act == FGNN_ACT_RELU ? 0.f : p.slope);
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = i / O8;
    const int o = (int)(i - g * O8) * 8;
    const int32_t* se = slot_edge + g * p.K;
    const int kt = p.tile_k ? p.tile_k[g >> 7] : p.K;
    float a[8], s[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { a[c] = AGG == FGNN_AGG_MEAN ? 0.f : -INFINITY; s[c] = 0.f; }
    float live = 0.f;
    for (int k0 = 0; k0 < kt; k0 += KB) {
      int32_t e[KB];
#pragma unroll
      for (int j = 0; j < KB; ++j) e[j] = k0 + j < kt ? __ldg(se + k0 + j) : -1;
      float v[KB][8];
#pragma unroll
      for (int j = 0; j < KB; ++j) {
        if (e[j] >= 0) ldg256(msg + (int64_t)e[j] * p.O + o, v[j]);
        else if (e[j] == kZeroSlot) {
#pragma unroll
          for (int c = 0; c < 8; ++c) v[j][c] = 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < KB; ++j) {
        if (e[j] < 0 && e[j] != kZeroSlot) continue;
        live += 1.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (AGG == FGNN_AGG_MAX) {
            a[c] = fmaxf(a[c], v[j][c]);
          } else if (AGG == FGNN_AGG_SOFTMAX) {
            softmax_push(a[c], s[c], v[j][c], p.gamma);
          } else {
            a[c] += v[j][c];
          }
        }
      }
    }
    reduce_finish<AGG>(p, a, s, live, g, o, neg);
  }
}

// Tables with few slots (K = 2, 3: the factor side of pairwise / order-3 types): a thread of the kernel above would have
// two or three loads in flight behind a dependent slot -> edge lookup.  Here a thread takes U = 6 / K items per
// iteration, so six message pieces are always in flight (everything indexed at compile time).
template <int AGG, int K>
__global__ void __launch_bounds__(256)
mp_reduce_small_kernel(const MpParams p, const float* __restrict__ msg, const int32_t* __restrict__ slot_edge) {
  constexpr int U = 6 / K;
  const int O8 = p.O >> 3;
  const int64_t rows = (int64_t)p.B * p.M, total = rows * O8;
  const float neg = p.act == FGNN_ACT_NONE ? 1.f : (p.act == FGNN_ACT_RELU ? 0.f : p.slope);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  pdl_wait();
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
    int64_t g[U];
    int o[U];
    int32_t e[U][K];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      g[u] = i < total ? i / O8 : -1;
      o[u] = g[u] >= 0 ? (int)(i - g[u] * O8) * 8 : 0;
#pragma unroll
      for (int k = 0; k < K; ++k) e[u][k] = g[u] >= 0 ? __ldg(slot_edge + g[u] * K + k) : -1;
    }
    float v[U][K][8];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < K; ++k)
      {
        if (e[u][k] >= 0) ldg256(msg + (int64_t)e[u][k] * p.O + o[u], v[u][k]);
        else if (e[u][k] == kZeroSlot) {
#pragma unroll
          for (int c = 0; c < 8; ++c) v[u][k][c] = 0.f;
        }
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (g[u] < 0) continue;
      float a[8], s[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { a[c] = AGG == FGNN_AGG_MEAN ? 0.f : -INFINITY; s[c] = 0.f; }
      float live = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (e[u][k] < 0 && e[u][k] != kZeroSlot) continue;
        live += 1.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (AGG == FGNN_AGG_MAX) a[c] = fmaxf(a[c], v[u][k][c]);
          else if (AGG == FGNN_AGG_SOFTMAX) softmax_push(a[c], s[c], v[u][k][c], p.gamma);
          else a[c] += v[u][k][c];
        }
      }
      reduce_finish<AGG>(p, a, s, live, g[u], o[u], neg);
    }
  }
}

// etype [B,T,M,K] (reference layout) -> the plan's edge-type image: edge-major [E,T] in the plan's edge order,
// 16-byte pieces swizzled by et_key (above)
__global__ void et_permute_kernel(const float* __restrict__ et, int64_t et_sb, const int32_t* __restrict__ edge_slot,
                                  float* __restrict__ out, int T, int64_t MK, int64_t n_edges) {
  const int64_t total = n_edges * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / T;
    const int t = (int)(i - e * T);
    const int64_t slot = edge_slot[e];                       // (b*M + m)*K + k
    const int64_t b = slot / MK, mk = slot - b * MK;
    const uint32_t key = T == 16 ? et_key<16>((uint32_t)e) : (T == 8 ? et_key<8>((uint32_t)e) : 0u);
    out[e * T + ((((uint32_t)t >> 2) ^ key) << 2) + (t & 3)] = et[b * et_sb + (int64_t)t * MK + mk];
  }
}

template <int T, int NCH, bool ES, bool FUSE>
constexpr size_t src_smem_bytes() {
  constexpr int RC = ES ? 2 * kEB : kEB;
  return 1024 + (size_t)tc::w_bytes(128 * NCH, false) + (size_t)src_x_stages(ES) * tc::stage_bytes(false) +
         (size_t)(FUSE ? 2 : 1) * tc::kTileM * RC * T * 4 + (FUSE ? src_fuse_staging_max() : 0) + tc::kNumBars * 8 + 16;
}

template <int T, int NCH, bool ES, bool FUSE = false>
int launch_src(const SrcParams& sp, const MpParams& mp, const uint8_t* wimg, int S, int workers, int tiles, cudaStream_t st) {
  auto kern = mp_src_kernel<T, NCH, ES, FUSE>;
  constexpr size_t smem = src_smem_bytes<T, NCH, ES, FUSE>();
  static_assert(smem <= (size_t)tc::kSmemBudget, "shared memory budget");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return FGNN_ERR_CUDA;
    if (fa.numRegs != tc::kRegLaunch) return FGNN_ERR_UNSUPPORTED;      // setmaxnreg budget (tc_common.cuh)
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return FGNN_ERR_CUDA;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(S * workers));
  cfg.blockDim = dim3(tc::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc_pdl_enabled() ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, sp, mp, wimg, S, workers, tiles);
  count_launch();
  return e == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

template <int T, int NCH>
int launch_src_es(bool es, const SrcParams& sp, const MpParams& mp, const uint8_t* wimg, int S, int workers, int tiles, cudaStream_t st) {
  return es ? launch_src<T, NCH, true>(sp, mp, wimg, S, workers, tiles, st) : launch_src<T, NCH, false>(sp, mp, wimg, S, workers, tiles, st);
}

}  // namespace

bool src_supported(const fgnn_mp_args* a) {
  if (!a->src_ptr || !a->slot_edge || !a->etype_edges || (!a->messages && !a->src_edge_slot)) return false;
  if (a->extension != FGNN_NO_EXTENSION || a->dtype != FGNN_F32 || a->C != tc::kC) return false;
  if (a->aggregator == FGNN_AGG_NONE) return false;
  if (a->T != 16 && a->T != 8 && a->T != 4) return false;
  const int OT = a->O * a->T;
  if ((OT != 256 && OT % 512) || a->O % 8 || a->O > 128) return false;            // a CTA owns 256 or 512 columns: no ragged slice
  if (a->src_row_cap != kEB && a->src_row_cap != 2 * kEB) return false;
  const int64_t rows = (int64_t)a->B * a->N;
  if (a->n_src_rows < rows || a->n_src_rows >= INT32_MAX - 256 || (a->n_src_rows > rows && !a->src_rows)) return false;
  if (a->x_sc != 1 || a->x_sn != a->C) return false;                              // node-major rows
  if (a->B > 1 && a->x_sb != (int64_t)a->N * a->C) return false;                  // batch-contiguous
  if ((int64_t)a->B * a->N * tc::row_bytes(false) >= (int64_t)UINT32_MAX) return false;
  if (a->n_edges <= 0 || a->n_edges >= INT32_MAX) return false;
  if ((reinterpret_cast<uintptr_t>(a->x) & 15) || (reinterpret_cast<uintptr_t>(a->out) & 15) ||
      (a->messages && (reinterpret_cast<uintptr_t>(a->messages) & 31)) || (reinterpret_cast<uintptr_t>(a->etype_edges) & 15))
    return false;
  if (a->out_so != 1 || (a->out_sm & 3)) return false;
  if (a->B > 1 && a->out_sb != (int64_t)a->M * a->out_sm) return false;
  return true;
}

int launch_mp_src(const MpParams& p, const fgnn_mp_args* a, cudaStream_t stream) {
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255)) return FGNN_ERR_WORKSPACE;
  const int OT = p.O * p.T;
  const int wrc = tc_prepare_weights(p.W, ws, tc::kC, OT, a->filters_version, stream, 0, p.T);
  if (wrc != FGNN_OK) return wrc;
  SrcParams sp;
  sp.x = p.x; sp.vptr = a->src_ptr; sp.xrow = a->src_rows; sp.et_edges = reinterpret_cast<const float*>(a->etype_edges);
  sp.msg = reinterpret_cast<float*>(a->messages);
  sp.rows = (uint32_t)((int64_t)p.B * p.N);
  sp.vrows = (uint32_t)a->n_src_rows;
  sp.O = p.O; sp.T = p.T;
  sp.edge_slot = a->src_edge_slot; sp.rpt = tc::kTileM; sp.M = p.M; sp.K = p.K;
  const bool es = a->src_row_cap == 2 * kEB;
  // columns per CTA: 512 (the split-bf16 image of 512 columns is 128 KB), or all 256 of them
  const int cols = OT < 512 ? OT : 512;
  const int NCH = cols / 128, S = OT / cols;
  int sms = tc_num_sms();
  if (a->sm_limit > 0 && a->sm_limit < sms) sms = a->sm_limit;
  if (S > sms) return FGNN_ERR_UNSUPPORTED;
  // Fused aggregation: a table [B,M,K] is block-diagonal over the batch -- the destinations of batch element b read
  // source rows of b only -- so when one batch element's sources fit a tile (N <= 128), the CTA owns all filter
  // columns (O*T = 256: T = 4 at O = 64, the LDPC shape) and every slot is live, the messages go to shared memory and
  // the CTA aggregates them itself: no message round trip through HBM, no second launch.
  // The plan for it (SourcePlan(batch_local=True)) numbers the virtual rows batch element by batch element --
  // src_rows_per_batch of them each, row cap 3, src_rows naming the source row of EVERY virtual row.
  const int rpb = a->src_rows_per_batch;
  const bool fuse = !(a->flags & FGNN_FLAG_NO_FUSED_REDUCE) && a->src_edge_slot && rpb > 0 && rpb <= tc::kTileM && !es && S == 1 &&
                    p.T == 4 && NCH == 2 && a->src_rows && a->n_src_rows == (int64_t)p.B * rpb &&
                    a->n_edges == (int64_t)p.B * p.M * p.K && !p.tile_k && !p.out_rows && !p.mask_neg &&
                    (int64_t)p.M * p.K * p.O * 4 <= src_fuse_staging_max();
  if (rpb > 0) {
    if (!fuse) return FGNN_ERR_UNSUPPORTED;                   // a batch-local plan only serves the fused path
    sp.rpt = (uint32_t)rpb;
    sp.rows = 0;                                              // every virtual row takes its source row from src_rows
    int workers = sms < p.B ? sms : p.B;
    return launch_src<4, 2, false, true>(sp, p, ws, 1, workers, p.B, stream);
  }
  if (!sp.msg) return FGNN_ERR_INVALID_ARG;
  const int tiles = (int)((sp.vrows + tc::kTileM - 1) / tc::kTileM);
  int workers = sms / S;
  if (workers > tiles) workers = tiles;
  int rc = FGNN_ERR_UNSUPPORTED;
  if (p.T == 16 && NCH == 4) rc = launch_src_es<16, 4>(es, sp, p, ws, S, workers, tiles, stream);
  else if (p.T == 16 && NCH == 2) rc = launch_src_es<16, 2>(es, sp, p, ws, S, workers, tiles, stream);
  else if (p.T == 8 && NCH == 4) rc = launch_src_es<8, 4>(es, sp, p, ws, S, workers, tiles, stream);
  else if (p.T == 8 && NCH == 2) rc = launch_src_es<8, 2>(es, sp, p, ws, S, workers, tiles, stream);
  else if (p.T == 4 && NCH == 4) rc = launch_src_es<4, 4>(es, sp, p, ws, S, workers, tiles, stream);
  else if (p.T == 4 && NCH == 2) rc = launch_src_es<4, 2>(es, sp, p, ws, S, workers, tiles, stream);
  if (rc != FGNN_OK) return rc;
  // pass 2
  const int64_t total = (int64_t)p.B * p.M * (p.O / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaError_t e;
  const float* m = sp.msg;
  const bool small = !p.tile_k && (p.K == 2 || p.K == 3);
  if (small && p.K == 2) {
    if (p.agg == FGNN_AGG_MAX) e = cudaLaunchKernelEx(&cfg, mp_reduce_small_kernel<FGNN_AGG_MAX, 2>, p, m, a->slot_edge);
    else if (p.agg == FGNN_AGG_SOFTMAX) e = cudaLaunchKernelEx(&cfg, mp_reduce_small_kernel<FGNN_AGG_SOFTMAX, 2>, p, m, a->slot_edge);
    else e = cudaLaunchKernelEx(&cfg, mp_reduce_small_kernel<FGNN_AGG_MEAN, 2>, p, m, a->slot_edge);
  } else if (small) {
    if (p.agg == FGNN_AGG_MAX) e = cudaLaunchKernelEx(&cfg, mp_reduce_small_kernel<FGNN_AGG_MAX, 3>, p, m, a->slot_edge);
    else if (p.agg == FGNN_AGG_SOFTMAX) e = cudaLaunchKernelEx(&cfg, mp_reduce_small_kernel<FGNN_AGG_SOFTMAX, 3>, p, m, a->slot_edge);
    else e = cudaLaunchKernelEx(&cfg, mp_reduce_small_kernel<FGNN_AGG_MEAN, 3>, p, m, a->slot_edge);
  } else if (p.agg == FGNN_AGG_MAX) e = cudaLaunchKernelEx(&cfg, mp_reduce_kernel<FGNN_AGG_MAX>, p, m, a->slot_edge);
  else if (p.agg == FGNN_AGG_SOFTMAX) e = cudaLaunchKernelEx(&cfg, mp_reduce_kernel<FGNN_AGG_SOFTMAX>, p, m, a->slot_edge);
  else e = cudaLaunchKernelEx(&cfg, mp_reduce_kernel<FGNN_AGG_MEAN>, p, m, a->slot_edge);
  count_launch();
  return e == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

int launch_et_permute(const float* et, int64_t et_sb, const int32_t* edge_slot, float* out, int T, int64_t MK,
                      int64_t n_edges, cudaStream_t stream) {
  const int64_t total = n_edges * T;
  if (total == 0) return FGNN_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  et_permute_kernel<<<(unsigned)blocks, 256, 0, stream>>>(et, et_sb, edge_slot, out, T, MK, n_edges);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? (int)FGNN_OK : (int)FGNN_ERR_CUDA;
}

}  // namespace fgnn

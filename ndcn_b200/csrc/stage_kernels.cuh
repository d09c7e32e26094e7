// Stage kernels: one right-hand-side evaluation fused with the Runge-Kutta stage algebra
// that consumes it.  k never makes an extra HBM round trip: it is produced in registers
// (gather -> [SMEM tile -> W GEMM] -> bias/ReLU) and handed straight to epi_apply().
//
//   k_stage_ndcn_gemm   relu((Phi x) W^T + b), H in {32,64,128,256}     neural_dynamics.py:20-39
//   k_stage_ndcn_row    same with no_control (no W) or as plain SpMM, H in {32,64,128,256}
//   k_stage_ndcn_any    any H <= 1024 (H=20 default of the dynamics scripts, H=16 of dgnn, H=1)
//   k_stage_dyn1        Heat / Gene / Mutualistic on a [N,1] state       *_dynamics.py:186-232
//   k_stage_dynv        same on a [N,d] state, d > 1
//   k_epi_only          epilogue on a k that already sits in HBM (callback RHS, dopri5 pre-stage)
#pragma once
#include "ndcn_common.cuh"

namespace ndcn {

constexpr int kTileRows = 64;   // rows per CTA in the GEMM kernel
constexpr int kKChunk = 16;     // W^T rows per TMA bulk chunk

// ---------------------------------------------------------------------------------------
// Sparse row gather: acc[ch][i] = sum_j val_j * x[col_j, ch*32*VW + lane*VW + i]
// One warp per row; the warp reads 32 (col,val) pairs coalesced, broadcasts them by
// shuffle and keeps U independent 16-byte row loads per lane in flight.
// Accumulation order along the row is the CSR order, i.e. the order torch.sparse.mm's COO
// worker visits a row's entries on the CPU.
// ---------------------------------------------------------------------------------------
// Entries [start, end) of one row, visited in batches of 32 starting at `start` with a stride of
// `batch_stride` entries (32 for a whole row; 32 * n_warps when the warps of a CTA share a long row).
template <int VW, int NCH, int U = 4>
__device__ __forceinline__ void gather_range(const GraphView& g, int start, int end, int batch_stride,
                                             const float* __restrict__ x, int lane, float (&acc)[NCH][VW]) {
  constexpr int H = 32 * VW * NCH;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int i = 0; i < VW; ++i) acc[ch][i] = 0.f;
  const float* xl = x + lane * VW;
  auto load_row = [&](int cc, float (&dst)[NCH][VW]) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) ldv<VW>(xl + (int64_t)cc * H + ch * 32 * VW, dst[ch]);
  };
  for (int base = start; base < end; base += batch_stride) {
    const int idx = base + lane;
    int my_c = 0;
    float my_v = 0.f;
    if (idx < end) {
      my_c = __ldcs(g.col + idx);
      my_v = __ldcs(g.val + idx);
    }
    const int cnt = min(32, end - base);
    int j = 0;
    for (; j + U <= cnt; j += U) {
      int c[U];
      float v[U];
      float xv[U][NCH][VW];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        c[u] = __shfl_sync(0xffffffffu, my_c, j + u);
        v[u] = __shfl_sync(0xffffffffu, my_v, j + u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) load_row(c[u], xv[u]);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
          for (int i = 0; i < VW; ++i) acc[ch][i] = fmaf(v[u], xv[u][ch][i], acc[ch][i]);
    }
    for (; j < cnt; ++j) {
      const int c = __shfl_sync(0xffffffffu, my_c, j);
      const float v = __shfl_sync(0xffffffffu, my_v, j);
      float xv[NCH][VW];
      load_row(c, xv);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
        for (int i = 0; i < VW; ++i) acc[ch][i] = fmaf(v, xv[ch][i], acc[ch][i]);
    }
  }
}

template <int VW, int NCH>
__device__ __forceinline__ void gather_row(const GraphView& g, int64_t row, const float* __restrict__ x,
                                           int lane, float (&acc)[NCH][VW]) {
  const int start = __ldg(g.rowptr + row), end = __ldg(g.rowptr + row + 1);
  gather_range<VW, NCH>(g, start, end, 32, x, lane, acc);
}

constexpr int kLongRow = 256;  // entries; rows above this are shared by all warps of a CTA

struct NdcnArgs {
  GraphView g;
  PtrPair x;          // gather source [n_cols, H] (parity-selected)
  const float* Wt;    // [H(k), H(n)] = W^T, row-major (prepared once per solve)
  const float* bias;  // [H]
  uint32_t flags;     // NDCN_F_*
  const int32_t* long_rows;  // rows with more than kLongRow entries (may be null)
  int n_long;
  // row-chunked store-only gather: rows [row_begin, row_end) only (row_end == 0: all rows); long_rows / n_long then
  // name the chunk's long rows; keep_l2: plain stores (the consumer reads the chunk out of L2), not streaming ones
  int64_t row_begin, row_end;
  int keep_l2;
};

// ---------------------------------------------------------------------------------------
// no_control / plain SpMM: one warp per row, k = [relu](Phi x) (or relu(x) with no_graph),
// epilogue straight from the gather registers.
// ---------------------------------------------------------------------------------------
// Power-law hubs (degree ~ m sqrt(N), thousands of entries) would keep a single warp busy long
// after every other row is done: rows above kLongRow entries are skipped by the row-per-warp
// CTAs and handled by the first n_long CTAs of the grid, whose 8 warps interleave 32-entry
// batches of the row and add their partial sums in warp order (fixed, reproducible).
// STORE_ONLY: the epilogue is a plain streaming store of z = Phi x into e.k_out (what feeds the tcgen05 GEMM
// kernel); none of the stage algebra's pointers and coefficients then occupies registers.
template <int VW, int NCH, int U = 4, int MINB = 3, bool STORE_ONLY = false>
__global__ void __launch_bounds__(kStageThreads, MINB) k_stage_ndcn_row(NdcnArgs a, EpiArgs e) {
  constexpr int H = 32 * VW * NCH;
  __shared__ float s_part[kWarpsPerCta][H];
  EpiCtx c;
  if constexpr (STORE_ONLY) {
    if (e.ctrl != nullptr && ((volatile Ctrl*)e.ctrl)->done) return;
  } else {
    if (!epi_resolve(e, c)) return;
  }
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ x = sel(a.x, par);
  float* __restrict__ zout = STORE_ONLY ? sel(e.k_out, par) : nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool relu = !(a.flags & NDCN_F_NO_RELU);
  const bool graph = !(a.flags & NDCN_F_NO_GRAPH);
  const int n_long = graph ? a.n_long : 0;
  double err_acc = 0.0;
  auto finish = [&](int64_t off, float (&v)[VW]) {
    if constexpr (STORE_ONLY) {
      if (a.keep_l2) stv<VW>(zout + off, v);
      else stv_stream<VW>(zout + off, v);
    } else {
      epi_apply<VW>(c, off, v, err_acc);
    }
  };
  const int64_t row_end = a.row_end > 0 ? a.row_end : a.g.n_rows;
  if ((int)blockIdx.x < n_long) {
    const int64_t row = __ldg(a.long_rows + blockIdx.x);
    const int start = __ldg(a.g.rowptr + row), end = __ldg(a.g.rowptr + row + 1);
    float acc[NCH][VW];
    gather_range<VW, NCH, U>(a.g, start + warp * 32, end, 32 * kWarpsPerCta, x, lane, acc);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) stv<VW>(&s_part[warp][ch * 32 * VW + lane * VW], acc[ch]);
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float tot[VW];
#pragma unroll
        for (int i = 0; i < VW; ++i) tot[i] = 0.f;
        for (int w = 0; w < kWarpsPerCta; ++w) {
          float p[VW];
          ldv<VW>(&s_part[w][ch * 32 * VW + lane * VW], p);
#pragma unroll
          for (int i = 0; i < VW; ++i) tot[i] += p[i];
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < VW; ++i) tot[i] = fmaxf(tot[i], 0.f);
        }
        finish(row * H + ch * 32 * VW + lane * VW, tot);
      }
    }
  } else {
    const int64_t row = a.row_begin + (int64_t)(blockIdx.x - n_long) * kWarpsPerCta + warp;
    if (row < row_end) {
      float acc[NCH][VW];
      bool mine = true;
      if (!graph) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) ldv<VW>(x + row * H + ch * 32 * VW + lane * VW, acc[ch]);
      } else {
        const int start = __ldg(a.g.rowptr + row), end = __ldg(a.g.rowptr + row + 1);
        if (n_long > 0 && end - start > kLongRow) mine = false;  // produced by a long-row CTA
        else gather_range<VW, NCH, U>(a.g, start, end, 32, x, lane, acc);
      }
      if (mine) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          if (relu) {
#pragma unroll
            for (int i = 0; i < VW; ++i) acc[ch][i] = fmaxf(acc[ch][i], 0.f);
          }
          finish(row * H + ch * 32 * VW + lane * VW, acc[ch]);
        }
      }
    }
  }
  if constexpr (!STORE_ONLY) epi_finish_block(e, err_acc);
}

// ---------------------------------------------------------------------------------------
// Full ODEFunc: per CTA a tile of 64 rows.
//   phase 1  warps pull rows from a CTA-local counter, gather Phi x into the SMEM tile z
//   phase 2  z[64,H] @ W^T[H,H] on FP32 FMA pipes; W^T streams through SMEM in 16-row
//            chunks with cp.async.bulk (TMA 1-D) double-buffered on mbarriers
//   phase 3  + bias, ReLU, stage epilogue from registers
// Thread tile: 8 rows x (NCH*VW) columns, columns interleaved so that every global and
// shared access of a warp is one contiguous 32*VW*4-byte segment.
// ---------------------------------------------------------------------------------------
template <int VW, int NCH>
struct GemmSmem {
  static constexpr int H = 32 * VW * NCH;
  static constexpr int ZLD = H + 4;
  static constexpr size_t z_bytes = sizeof(float) * kTileRows * ZLD;
  static constexpr size_t w_bytes = sizeof(float) * 2 * kKChunk * H;
  static constexpr size_t total = z_bytes + w_bytes + 64;
};

// One tile of 64 rows starting at row0, by the whole CTA.  `chunk_it` counts the W^T chunks this CTA has consumed
// so far (buffer = it & 1, mbarrier parity = (it >> 1) & 1): a persistent CTA may call this for tile after tile.
// The caller has initialised bars[0..1] (count 1) and made the initialisation visible (mbar_fence_init + barrier).
template <int VW, int NCH>
__device__ __forceinline__ void ndcn_gemm_tile(const NdcnArgs& a, const EpiCtx& c, const float* __restrict__ x,
                                               int64_t row0, unsigned char* smem_raw, uint32_t& chunk_it,
                                               double& err_acc) {
  constexpr int H = 32 * VW * NCH;
  using S = GemmSmem<VW, NCH>;
  constexpr int ZLD = S::ZLD;
  constexpr int NCHUNKS = H / kKChunk;
  constexpr uint32_t kChunkBytes = kKChunk * H * sizeof(float);
  constexpr int RPW = kTileRows / kWarpsPerCta;  // rows per warp in the GEMM phase (8)

  float* z = reinterpret_cast<float*>(smem_raw);
  float* wbuf = reinterpret_cast<float*>(smem_raw + S::z_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + S::z_bytes + S::w_bytes);
  int* row_counter = reinterpret_cast<int*>(bars + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_here = (int)min((int64_t)kTileRows, a.g.n_rows - row0);
  const uint32_t it0 = chunk_it;

  if (threadIdx.x == 0) {
    *row_counter = 0;
    // W^T chunk 0 lands while the tile is being gathered
    mbar_arrive_expect_tx(&bars[it0 & 1], kChunkBytes);
    bulk_g2s(wbuf + (it0 & 1) * kKChunk * H, a.Wt, kChunkBytes, &bars[it0 & 1]);
  }
  __syncthreads();

  // ---- phase 1: gather ----
  if (a.flags & NDCN_F_NO_GRAPH) {
    for (int r = warp; r < kTileRows; r += kWarpsPerCta) {
      float v[NCH][VW];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        if (r < rows_here) ldv<VW>(x + (row0 + r) * H + ch * 32 * VW + lane * VW, v[ch]);
        else
#pragma unroll
          for (int i = 0; i < VW; ++i) v[ch][i] = 0.f;
        stv<VW>(z + r * ZLD + ch * 32 * VW + lane * VW, v[ch]);
      }
    }
  } else {
    for (;;) {
      int r = 0;
      if (lane == 0) r = atomicAdd(row_counter, 1);
      r = __shfl_sync(0xffffffffu, r, 0);
      if (r >= kTileRows) break;
      float acc[NCH][VW];
      if (r < rows_here) {
        gather_row<VW, NCH>(a.g, row0 + r, x, lane, acc);
      } else {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
          for (int i = 0; i < VW; ++i) acc[ch][i] = 0.f;
      }
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) stv<VW>(z + r * ZLD + ch * 32 * VW + lane * VW, acc[ch]);
    }
  }
  __syncthreads();

  // ---- phase 2: out[r][n] = sum_k z[r][k] * Wt[k][n] ----
  float acc[RPW][NCH][VW];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int i = 0; i < VW; ++i) acc[r][ch][i] = 0.f;

  const float* zw = z + (warp * RPW) * ZLD;
  for (int kc = 0; kc < NCHUNKS; ++kc) {
    const uint32_t it = it0 + (uint32_t)kc;
    if (threadIdx.x == 0 && kc + 1 < NCHUNKS) {
      const uint32_t nb = (it + 1) & 1;
      mbar_arrive_expect_tx(&bars[nb], kChunkBytes);
      bulk_g2s(wbuf + nb * kKChunk * H, a.Wt + (size_t)(kc + 1) * kKChunk * H, kChunkBytes, &bars[nb]);
    }
    mbar_wait(&bars[it & 1], (it >> 1) & 1);
    const float* wb = wbuf + (it & 1) * kKChunk * H + lane * VW;
#pragma unroll
    for (int kk = 0; kk < kKChunk; kk += 4) {
      float4 av[RPW];
#pragma unroll
      for (int r = 0; r < RPW; ++r)
        av[r] = *reinterpret_cast<const float4*>(zw + r * ZLD + kc * kKChunk + kk);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        float bv[NCH][VW];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) ldv<VW>(wb + (kk + k4) * H + ch * 32 * VW, bv[ch]);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          const float ar = k4 == 0 ? av[r].x : (k4 == 1 ? av[r].y : (k4 == 2 ? av[r].z : av[r].w));
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
            for (int i = 0; i < VW; ++i) acc[r][ch][i] = fmaf(ar, bv[ch][i], acc[r][ch][i]);
        }
      }
    }
    __syncthreads();  // everyone is done with this W buffer before it is refilled
  }
  chunk_it = it0 + NCHUNKS;

  // ---- phase 3: bias, ReLU, stage epilogue ----
  float bv[NCH][VW];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) ldv<VW>(a.bias + ch * 32 * VW + lane * VW, bv[ch]);
  const bool relu = !(a.flags & NDCN_F_NO_RELU);
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int rr = warp * RPW + r;
    if (rr < rows_here) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float kv[VW];
#pragma unroll
        for (int i = 0; i < VW; ++i) {
          kv[i] = acc[r][ch][i] + bv[ch][i];
          if (relu) kv[i] = fmaxf(kv[i], 0.f);
        }
        epi_apply<VW>(c, (row0 + rr) * H + ch * 32 * VW + lane * VW, kv, err_acc);
      }
    }
  }
}

template <int VW, int NCH>
__device__ __forceinline__ void ndcn_gemm_init_bars(unsigned char* smem_raw) {
  using S = GemmSmem<VW, NCH>;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + S::z_bytes + S::w_bytes);
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
}

template <int VW, int NCH>
__global__ void __launch_bounds__(kStageThreads, 2) k_stage_ndcn_gemm(NdcnArgs a, EpiArgs e) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  EpiCtx c;
  if (!epi_resolve(e, c)) return;
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ x = sel(a.x, par);
  ndcn_gemm_init_bars<VW, NCH>(smem_raw);
  uint32_t chunk_it = 0;
  double err_acc = 0.0;
  ndcn_gemm_tile<VW, NCH>(a, c, x, (int64_t)blockIdx.x * kTileRows, smem_raw, chunk_it, err_acc);
  epi_finish_block(e, err_acc);
}

// ---------------------------------------------------------------------------------------
// Any width H <= 1024: one warp per row, lanes stride the columns.  Used for the small
// widths of the reference scripts (H=20 dynamics default heat_dynamics.py:33, H=16 dgnn.py:42,
// H=1 with --baseline no_embed heat_dynamics.py:251-256); throughput is irrelevant there
// (N is a few hundred to a few thousand), launch count is what matters.
// ---------------------------------------------------------------------------------------
// one row of relu((Phi x) W^T + b) for any width, by one warp; zr = the warp's [H] scratch row in shared memory
__device__ __forceinline__ void ndcn_any_row(const NdcnArgs& a, int H, const float* __restrict__ W,
                                             const float* __restrict__ x, float* zr, int64_t row, int lane,
                                             const EpiCtx& c, double& err_acc) {
  if (a.flags & NDCN_F_NO_GRAPH) {
    for (int col = lane; col < H; col += 32) zr[col] = x[row * H + col];
  } else {
    const int start = a.g.rowptr[row], end = a.g.rowptr[row + 1];
    for (int col = lane; col < H; col += 32) {
      float s = 0.f;
      for (int j = start; j < end; ++j)
        s = fmaf(__ldg(a.g.val + j), x[(int64_t)__ldg(a.g.col + j) * H + col], s);
      zr[col] = s;
    }
  }
  __syncwarp();
  const bool relu = !(a.flags & NDCN_F_NO_RELU);
  for (int n = lane; n < H; n += 32) {
    float kv[1];
    if (a.flags & NDCN_F_NO_CONTROL) {
      kv[0] = zr[n];
    } else {
      float s = 0.f;
      const float* wr = W + (size_t)n * H;
      for (int k = 0; k < H; ++k) s = fmaf(zr[k], __ldg(wr + k), s);
      kv[0] = s + __ldg(a.bias + n);
    }
    if (relu) kv[0] = fmaxf(kv[0], 0.f);
    epi_apply<1>(c, row * H + n, kv, err_acc);
  }
  __syncwarp();  // zr is rewritten by the warp's next row
}

__global__ void __launch_bounds__(kStageThreads) k_stage_ndcn_any(NdcnArgs a, int H, const float* W, EpiArgs e) {
  extern __shared__ float zs[];  // [warps][H]
  EpiCtx c;
  if (!epi_resolve(e, c)) return;
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ x = sel(a.x, par);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  double err_acc = 0.0;
  if (row < a.g.n_rows) ndcn_any_row(a, H, W, x, zs + warp * H, row, lane, c, err_acc);
  epi_finish_block(e, err_acc);
}

// ---------------------------------------------------------------------------------------
// Ground-truth dynamics.
// ---------------------------------------------------------------------------------------
struct DynArgs {
  GraphView g;
  PtrPair x;
  int kind;   // NDCN_RHS_HEAT / GENE / MUTUAL
  int d;      // state width
  float p[8];
  // [N,1] kernel: rows above kLongRow entries get a CTA each (the first n_long CTAs of the grid): a power-law hub
  // (thousands of entries) would otherwise be walked by LPR lanes, one dependent load chain after the other --
  // measured 0.3 ms per evaluation at 1M nodes for 0.05 ms of real work
  const int32_t* long_rows;
  int n_long;
};

// torch.pow semantics for the exponents the scripts use (1, 2) are exact products
__device__ __forceinline__ float pow_like_torch(float x, float e) {
  if (e == 1.f) return x;
  if (e == 2.f) return fmul(x, x);
  if (e == 3.f) return fmul(fmul(x, x), x);
  if (e == 0.5f) return sqrtf(x);
  return powf(x, e);
}

template <int KIND>
__device__ __forceinline__ float dyn_neighbour(const float (&p)[8], float a, float xi, float xj, bool d1) {
  if constexpr (KIND == NDCN_RHS_HEAT) {
    return fmul(a, xj);
  } else if constexpr (KIND == NDCN_RHS_GENE) {
    const float xh = pow_like_torch(xj, p[2]);
    return fmul(a, fdiv(xh, fadd(xh, 1.0f)));  // gene_dynamics.py:202
  } else {
    // mutualistic_dynamics.py: d==1 branch (:206-216) puts e on the neighbour, the d>1 loop
    // (:217-231) on the row itself -- reproduce each
    const float dd = p[3], ee = p[4], hh = p[5];
    if (d1) return fmul(a, fdiv(fmul(xj, xi), fadd(fadd(dd, fmul(ee, xj)), fmul(hh, xi))));
    return fdiv(fmul(a, fmul(xi, xj)), fadd(fadd(dd, fmul(ee, xi)), fmul(hh, xj)));
  }
}

template <int KIND>
__device__ __forceinline__ float dyn_local(const float (&p)[8], float xi, float nb) {
  if constexpr (KIND == NDCN_RHS_HEAT) {
    return fmul(p[0], nb);  // self.k * f           heat_dynamics.py:204
  } else if constexpr (KIND == NDCN_RHS_GENE) {
    return fadd(fmul(-p[0], pow_like_torch(xi, p[1])), nb);
  } else {
    // b + x*(1 - x/k)*(x/c - 1)   mutualistic_dynamics.py:205
    const float t = fmul(fmul(xi, fsub(1.0f, fdiv(xi, p[1]))), fsub(fdiv(xi, p[2]), 1.0f));
    return fadd(fadd(p[0], t), nb);
  }
}

// [N,1] state: LPR lanes cooperate on one row (degree ~10), shuffle-reduce, then the
// results are compacted so that the epilogue's loads/stores are contiguous.
// the 32 / LPR rows starting at warp_global * (32 / LPR), by one warp
template <int KIND, int LPR>
__device__ __forceinline__ void dyn1_warp_rows(const DynArgs& a, const float* __restrict__ x, int64_t warp_global,
                                               int lane, const EpiCtx& c, double& err_acc) {
  constexpr int RPWARP = 32 / LPR;
  const int64_t row_base = warp_global * RPWARP;
  const int64_t row = row_base + lane / LPR;
  const int sub = lane % LPR;
  float s = 0.f, xi = 0.f;
  bool mine = true;  // false: a long row, produced by its own CTA
  if (row < a.g.n_rows) {
    xi = x[row];
    const int start = __ldg(a.g.rowptr + row);
    int end = __ldg(a.g.rowptr + row + 1);
    if (a.n_long > 0 && end - start > kLongRow) {
      mine = false;
      end = start;
    }
    for (int j = start + sub; j < end; j += LPR) {
      const float xj = x[__ldg(a.g.col + j)];
      s = fadd(s, dyn_neighbour<KIND>(a.p, __ldg(a.g.val + j), xi, xj, true));
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s = fadd(s, __shfl_xor_sync(0xffffffffu, s, o));
  float kval = dyn_local<KIND>(a.p, xi, s);
  // lane i < RPWARP takes the result of row row_base + i (held by lane i*LPR)
  kval = __shfl_sync(0xffffffffu, kval, (lane * LPR) & 31);
  mine = __shfl_sync(0xffffffffu, mine ? 1 : 0, (lane * LPR) & 31) != 0;
  const int64_t my_row = row_base + lane;
  if (lane < RPWARP && my_row < a.g.n_rows && mine) {
    float kv[1] = {kval};
    epi_apply<1>(c, my_row, kv, err_acc);
  }
}

template <int KIND, int LPR>
__global__ void __launch_bounds__(kStageThreads) k_stage_dyn1(DynArgs a, EpiArgs e) {
  __shared__ float s_sum[kStageThreads / 32];
  EpiCtx c;
  if (!epi_resolve(e, c)) return;
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ x = sel(a.x, par);
  double err_acc = 0.0;
  if ((int)blockIdx.x < a.n_long) {
    // one long row: the CTA's threads stride over its entries; lanes, then warps, are added in a fixed order
    const int64_t row = __ldg(a.long_rows + blockIdx.x);
    const float xi = x[row];
    const int start = __ldg(a.g.rowptr + row), end = __ldg(a.g.rowptr + row + 1);
    float s = 0.f;
    for (int j = start + (int)threadIdx.x; j < end; j += kStageThreads)
      s = fadd(s, dyn_neighbour<KIND>(a.p, __ldg(a.g.val + j), xi, x[__ldg(a.g.col + j)], true));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fadd(s, __shfl_xor_sync(0xffffffffu, s, o));
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < kStageThreads / 32; ++w) tot = fadd(tot, s_sum[w]);
      float kv[1] = {dyn_local<KIND>(a.p, xi, tot)};
      epi_apply<1>(c, row, kv, err_acc);
    }
  } else {
    dyn1_warp_rows<KIND, LPR>(a, x, ((int64_t)(blockIdx.x - a.n_long) * blockDim.x + threadIdx.x) >> 5, threadIdx.x & 31,
                              c, err_acc);
  }
  epi_finish_block(e, err_acc);
}

// [N,1] state at scale ("CSR-stream"): a CTA owns 256 consecutive rows.  Their (col, val) slice is contiguous: it is
// staged in shared memory with coalesced loads, then the CTA's threads evaluate the per-ENTRY terms
// a_ij g(x_i, x_j) entry-parallel -- thread t takes entries t, t + 256, ... whatever row they belong to, with 4
// independent x[col] loads in flight -- and park them in shared memory; finally thread r adds the terms of row r in
// CSR order (the order torch.sparse.mm's CPU kernel uses) and runs the stage epilogue on consecutive rows (coalesced).
// The lane-group kernel above walks every row behind its own chain of dependent loads (rowptr -> col -> x): at 1M rows
// of ~11 entries that chain, not bandwidth, set its 0.25 ms per evaluation.  Slices larger than the staging buffer are
// processed in rounds; rows above kLongRow entries keep their own CTAs.
constexpr int kDynRows = 256;
constexpr int kDynCap = 4096;

template <int KIND>
__global__ void __launch_bounds__(kStageThreads) k_stage_dyn1_stream(DynArgs a, EpiArgs e) {
  __shared__ int s_rp[kDynRows + 1];
  __shared__ float s_xi[kDynRows];
  __shared__ int s_col[kDynCap];
  __shared__ float s_val[kDynCap];  // values, then the per-entry terms
  __shared__ float s_sum[kStageThreads / 32];
  EpiCtx c;
  if (!epi_resolve(e, c)) return;
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ x = sel(a.x, par);
  double err_acc = 0.0;
  if ((int)blockIdx.x < a.n_long) {
    const int64_t row = __ldg(a.long_rows + blockIdx.x);
    const float xi = x[row];
    const int start = __ldg(a.g.rowptr + row), end = __ldg(a.g.rowptr + row + 1);
    float s = 0.f;
    for (int j = start + (int)threadIdx.x; j < end; j += kStageThreads)
      s = fadd(s, dyn_neighbour<KIND>(a.p, __ldg(a.g.val + j), xi, x[__ldg(a.g.col + j)], true));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fadd(s, __shfl_xor_sync(0xffffffffu, s, o));
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < kStageThreads / 32; ++w) tot = fadd(tot, s_sum[w]);
      float kv[1] = {dyn_local<KIND>(a.p, xi, tot)};
      epi_apply<1>(c, row, kv, err_acc);
    }
  } else {
    const int64_t r0 = (int64_t)(blockIdx.x - a.n_long) * kDynRows;
    const int nr = (int)min((int64_t)kDynRows, a.g.n_rows - r0);
    const int t = threadIdx.x;
    for (int i = t; i <= nr; i += kStageThreads) s_rp[i] = __ldg(a.g.rowptr + r0 + i);
    if (t < nr) s_xi[t] = x[r0 + t];
    __syncthreads();
    const int e0 = s_rp[0], e1 = s_rp[nr];
    int my_start = 0, my_end = 0;
    bool mine = false;
    if (t < nr) {
      my_start = s_rp[t];
      my_end = s_rp[t + 1];
      mine = !(a.n_long > 0 && my_end - my_start > kLongRow);
    }
    float s = 0.f;
    for (int base = e0; base < e1; base += kDynCap) {
      const int cnt = min(kDynCap, e1 - base);
      for (int i = t; i < cnt; i += kStageThreads) {
        s_col[i] = __ldcs(a.g.col + base + i);
        s_val[i] = __ldcs(a.g.val + base + i);
      }
      __syncthreads();
      // entry-parallel terms, 4 independent x loads in flight per thread
      for (int i0 = t; i0 < cnt; i0 += 4 * kStageThreads) {
        float xj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kStageThreads;
          xj[u] = i < cnt ? x[s_col[i]] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kStageThreads;
          if (i < cnt) {
            // row of entry (base + i): the last r with s_rp[r] <= base + i
            const int ge = base + i;
            int lo = 0, hi = nr;
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if (s_rp[mid] <= ge) lo = mid;
              else hi = mid;
            }
            s_val[i] = dyn_neighbour<KIND>(a.p, s_val[i], s_xi[lo], xj[u], true);
          }
        }
      }
      __syncthreads();
      if (mine) {
        const int lo = max(my_start, base) - base, hi = min(my_end, base + cnt) - base;
        for (int i = lo; i < hi; ++i) s = fadd(s, s_val[i]);
      }
      __syncthreads();
    }
    if (mine) {
      float kv[1] = {dyn_local<KIND>(a.p, s_xi[t], s)};
      epi_apply<1>(c, r0 + t, kv, err_acc);
    }
  }
  epi_finish_block(e, err_acc);
}

// [N,d] state, d > 1: one warp per row, lanes over the d columns.
template <int KIND>
__device__ __forceinline__ void dynv_row(const DynArgs& a, const float* __restrict__ x, int64_t row, int lane,
                                         const EpiCtx& c, double& err_acc) {
  const int d = a.d;
  const int start = __ldg(a.g.rowptr + row), end = __ldg(a.g.rowptr + row + 1);
  for (int col = lane; col < d; col += 32) {
    const float xi = x[row * d + col];
    float s = 0.f;
    for (int j = start; j < end; ++j) {
      const float xj = x[(int64_t)__ldg(a.g.col + j) * d + col];
      s = fadd(s, dyn_neighbour<KIND>(a.p, __ldg(a.g.val + j), xi, xj, false));
    }
    float kv[1] = {dyn_local<KIND>(a.p, xi, s)};
    epi_apply<1>(c, row * d + col, kv, err_acc);
  }
}

template <int KIND>
__global__ void __launch_bounds__(kStageThreads) k_stage_dynv(DynArgs a, EpiArgs e) {
  EpiCtx c;
  if (!epi_resolve(e, c)) return;
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ x = sel(a.x, par);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  double err_acc = 0.0;
  if (row < a.g.n_rows) dynv_row<KIND>(a, x, row, lane, c, err_acc);
  epi_finish_block(e, err_acc);
}

// ---------------------------------------------------------------------------------------
// Epilogue on a k that is already in HBM: dopri5's first stage input y0 + dt*b10*k0 (k0 is
// the FSAL derivative), the initial-step probe y0 + h0*f0 (misc.py:132), and every stage of
// a callback RHS.  Pure streaming, 16-byte accesses.
// ---------------------------------------------------------------------------------------
// the elements [tid, numel) in steps of `stride` threads; returns whether a non-finite y0 was seen (check_finite)
__device__ __forceinline__ bool epi_only_range(const float* __restrict__ kin, int64_t numel, const EpiCtx& c,
                                               int check_finite, int vec, int64_t tid, int64_t stride, double& err_acc) {
  bool bad = false;
  const int64_t n4 = vec ? (numel >> 2) : 0;  // vec: every buffer 16-byte aligned
  for (int64_t i = tid; i < n4; i += stride) {
    float kv[4];
    ldv<4>(kin + i * 4, kv);
    if (check_finite) {
      float yv[4];
      ldv<4>(c.y0 + i * 4, yv);
#pragma unroll
      for (int q = 0; q < 4; ++q) bad |= !isfinite(yv[q]);
    }
    epi_apply<4>(c, i * 4, kv, err_acc);
  }
  for (int64_t i = (n4 << 2) + tid; i < numel; i += stride) {
    float kv[1] = {kin[i]};
    if (check_finite) bad |= !isfinite(c.y0[i]);
    epi_apply<1>(c, i, kv, err_acc);
  }
  return bad;
}

__global__ void __launch_bounds__(kStageThreads) k_epi_only(PtrPair k_in, int64_t numel, EpiArgs e, int vec) {
  EpiCtx c;
  if (!epi_resolve(e, c)) return;
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ kin = sel(k_in, par);
  double err_acc = 0.0;
  const bool bad = epi_only_range(kin, numel, c, e.check_finite, vec, (int64_t)blockIdx.x * blockDim.x + threadIdx.x,
                                  (int64_t)gridDim.x * blockDim.x, err_acc);
  if (e.check_finite && bad && e.ctrl) {
    atomicExch(&e.ctrl->status, NDCN_E_NONFINITE);
  }
  epi_finish_block(e, err_acc);
}

// fp32 stage time handed to a callback RHS (the reference converts t0 and dt to the state's dtype before it forms
// the stage times, rk_common.py:44-49):
//   mode 0  t0                      func(t, y) of every first stage; dopri5.py:78
//   mode 1  t0 + alpha*dt           rk_common.py:49; rk_common.py:78 (alpha = 1)
//   mode 2  t0 + (dt*num)/den       fixed_grid.py:20 (dt/2), rk_common.py:75-76 (dt/3, dt*2/3)
//   mode 3  base + h0 in float64    the probe of _select_initial_step, misc.py:126 (t0 is float64 there)
// from_ctrl: t0 / dt are the adaptive solver's current time and step (Ctrl::t1, Ctrl::dt) instead of host values.
__global__ void k_stage_time(const Ctrl* ctrl, int from_ctrl, float base, float dt_host, float alpha, float num,
                             float den, int mode, float* t_stage) {
  float t0 = base, dt = dt_host;
  if (from_ctrl) {
    t0 = (float)ctrl->t1;
    dt = (float)ctrl->dt;
  }
  if (mode == 0) *t_stage = t0;
  else if (mode == 1) *t_stage = fadd(t0, fmul(alpha, dt));
  else if (mode == 2) *t_stage = fadd(t0, fdiv(fmul(dt, num), den));
  else *t_stage = (float)((double)base + (double)ctrl->h0);
}

}  // namespace ndcn

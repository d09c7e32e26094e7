// Chunk-major sparse gather: z = Phi x (optionally + ReLU + RK stage epilogue) for states that do
// not fit in L2.
//
// A full-row gather of a [N, 256] fp32 state touches 1 KB per nonzero; on a graph without
// locality (power-law, ER) nearly every one of those reads misses the 126 MB L2 and the kernel
// moves ~E*H*4 bytes from HBM (measured 13 GB for 2.1 GB algorithmic at N=1M, profiles/README.md).
// Here the grid walks the state in COLUMN CHUNKS of CW floats: all CTAs that are resident at the
// same time gather from the same [N, CW] slab (N*CW*4 bytes, e.g. 64 MB for N=1M, CW=16), which
// stays L2-resident for the whole pass; the slab is read from HBM once and the E*H*4 gather
// bytes are served by L2.  Cost: (col, val) are re-read once per chunk.
//
//   blockIdx.x = chunk * (n_rb + n_long) + b      chunk-major, so passes follow each other
//   b <  n_rb : LPR = CW/4 lanes per row (16-byte loads), 32/LPR rows per warp, entries are
//               accumulated in CSR order (= the order torch.sparse.mm's CPU kernel visits a row,
//               neural_dynamics.py:29), rows longer than kLongRow are skipped here ...
//   b >= n_rb : ... and handled by one CTA per (long row, chunk): power-law hubs (degree ~ m*sqrt(N))
//               would otherwise serialise thousands of dependent loads in one lane group.
#pragma once
#include "ndcn_common.cuh"
#include "stage_kernels.cuh"

namespace ndcn {


template <int CW>
__global__ void __launch_bounds__(kStageThreads) k_stage_gather_chunk(NdcnArgs a, int H, int n_rb, int n_long,
                                                                       const int32_t* __restrict__ long_rows, EpiArgs e) {
  constexpr int LPR = CW / 4;               // lanes per row
  constexpr int RPW = 32 / LPR;             // rows per warp
  constexpr int RPC = RPW * kWarpsPerCta;   // rows per CTA
  constexpr int G = kStageThreads / LPR;    // lane groups per CTA (long-row path)
  __shared__ float s_part[G * CW];          // long-row partial sums, 4 KB

  EpiCtx c;
  if (!epi_resolve(e, c)) return;
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ x = sel(a.x, par);
  const int bpc = n_rb + n_long;
  const int chunk = blockIdx.x / bpc;
  const int b = blockIdx.x - chunk * bpc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = threadIdx.x % LPR;
  const float* __restrict__ xc = x + chunk * CW + sub * 4;  // this lane's 4 columns of row 0
  const bool relu = !(a.flags & NDCN_F_NO_RELU);
  double err_acc = 0.0;

  if (b < n_rb) {
    const int64_t row = (int64_t)b * RPC + warp * RPW + lane / LPR;
    int start = 0, end = 0;
    bool mine = false;
    if (row < a.g.n_rows) {
      start = __ldg(a.g.rowptr + row);
      end = __ldg(a.g.rowptr + row + 1);
      mine = true;
      if (end - start > kLongRow) {  // a long-row CTA produces this row
        mine = false;
        end = start;
      }
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.flags & NDCN_F_NO_GRAPH) {
      if (mine) ldv<4>(xc + row * H, acc);
    } else {
      const int n_it = (end - start + LPR - 1) / LPR;
      const int n_it_w = __reduce_max_sync(0xffffffffu, n_it);
      int my_c = 0;
      float my_v = 0.f;
      if (start + sub < end) {
        my_c = __ldg(a.g.col + start + sub);
        my_v = __ldg(a.g.val + start + sub);
      }
      for (int it = 0; it < n_it_w; ++it) {
        const int base = start + it * LPR;
        const int cnt = min(LPR, max(end - base, 0));
        // prefetch the next batch of (col, val) before the dependent row loads
        int nx_c = 0;
        float nx_v = 0.f;
        if (base + LPR + sub < end) {
          nx_c = __ldg(a.g.col + base + LPR + sub);
          nx_v = __ldg(a.g.val + base + LPR + sub);
        }
        float xv[LPR][4];
        float vv[LPR];
#pragma unroll
        for (int j = 0; j < LPR; ++j) {
          const int cj = __shfl_sync(0xffffffffu, my_c, j, LPR);
          vv[j] = __shfl_sync(0xffffffffu, my_v, j, LPR);
          if (j < cnt) ldv<4>(xc + (int64_t)cj * H, xv[j]);
        }
#pragma unroll
        for (int j = 0; j < LPR; ++j) {
          if (j < cnt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(vv[j], xv[j][i], acc[i]);
          }
        }
        my_c = nx_c;
        my_v = nx_v;
      }
    }
    if (mine) {
      if (relu) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaxf(acc[i], 0.f);
      }
      epi_apply<4>(c, row * H + chunk * CW + sub * 4, acc, err_acc);
    }
  } else {
    // ---- one long row: G lane groups stride over its entries, fixed-order reduction ----
    const int64_t row = __ldg(long_rows + (b - n_rb));
    const int start = __ldg(a.g.rowptr + row), end = __ldg(a.g.rowptr + row + 1);
    const int g = threadIdx.x / LPR;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int idx = start + g; idx < end; idx += 4 * G) {
      float xv[4][4];
      float vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int id = idx + u * G;
        vv[u] = 0.f;
        if (id < end) {
          vv[u] = __ldg(a.g.val + id);
          ldv<4>(xc + (int64_t)__ldg(a.g.col + id) * H, xv[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (idx + u * G < end) {
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(vv[u], xv[u][i], acc[i]);
        }
      }
    }
    stv<4>(s_part + g * CW + sub * 4, acc);
    __syncthreads();
    if (threadIdx.x < LPR) {
      float tot[4] = {0.f, 0.f, 0.f, 0.f};
      for (int gg = 0; gg < G; ++gg) {
        float p[4];
        ldv<4>(s_part + gg * CW + sub * 4, p);
#pragma unroll
        for (int i = 0; i < 4; ++i) tot[i] += p[i];
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 4; ++i) tot[i] = fmaxf(tot[i], 0.f);
      }
      epi_apply<4>(c, row * H + chunk * CW + sub * 4, tot, err_acc);
    }
  }
  epi_finish_block(e, err_acc);
}


// =========================================================================================
// v2: persistent, TMA-staged chunk-major gather.
//
// The v1 kernels above give every lane group exactly one row, so a CTA lives for three
// dependent memory round trips (rowptr -> (col, val) -> x rows) and has row loads in flight for
// only a third of its life (ncu: DRAM 35-59 %, L2 26-38 %, i.e. latency bound).  Here CTAs are
// persistent and walk work items (chunk, block of 128 rows) in chunk-major order; a loader warp
// stages the block's CSR slice -- rowptr by coalesced loads, the contiguous (col, val) ranges
// by cp.async.bulk (TMA 1-D) onto an mbarrier -- two items ahead of the 8 gather warps, which
// read indices from shared memory and keep up to 8 independent 16-byte row loads per lane in
// flight.  Rows longer than kLongRow are items of their own (whole CTA, fixed-order reduction).
// Accumulation order of a regular row = CSR order, as in v1.
// =========================================================================================
constexpr int kG2Rows = 128;      // rows per work item
constexpr int kG2Cap = 2048;      // (col, val) entries staged per item; the rest is read from global
constexpr int kG2Threads = kStageThreads + 32;  // 8 gather warps + 1 loader warp

template <int CW>
struct G2Smem {
  int32_t col[2][kG2Cap + 8];
  float val[2][kG2Cap + 8];
  int32_t rp[2][kG2Rows + 4];
  float part[(kStageThreads / (CW / 4)) * CW];  // long-row partial sums (4 KB)
  uint64_t full[2], empty[2];
};

__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// kStoreOnly: the epilogue is a plain streaming store of z (the tcgen05 GEMM kernel consumes it);
// the stage algebra's pointers and coefficients then never occupy registers.
template <int CW, bool kStoreOnly>
__global__ void __launch_bounds__(kG2Threads, kStoreOnly ? 3 : 2)
k_stage_gather_v2(NdcnArgs a, int H, int n_blocks, int n_long, const int32_t* __restrict__ long_rows, EpiArgs e) {
  constexpr int LPR = CW / 4;               // lanes per row
  constexpr int G = kStageThreads / LPR;    // lane groups per CTA
  constexpr int RPG = kG2Rows / G;          // rows per group and item
  constexpr int U = 8;                      // row loads in flight per lane
  __shared__ __align__(16) G2Smem<CW> sm;

  EpiCtx c;
  if (!epi_resolve(e, c)) return;
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ x = sel(a.x, par);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ipc = n_blocks + n_long;                       // items per chunk
  const int64_t total = (int64_t)(H / CW) * ipc;
  const bool relu = !(a.flags & NDCN_F_NO_RELU);
  const int64_t nnz = a.g.nnz;
  // the staged window never reads past the arrays: whole 16-byte groups only
  const int64_t nnz_vec = nnz & ~(int64_t)3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], kWarpsPerCta);
    }
    mbar_fence_init();
  }
  __syncthreads();

  double err_acc = 0.0;
  if (warp == kWarpsPerCta) {
    // ------------------------------ loader warp ------------------------------
    uint32_t cnt = 0;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
      const int b = (int)(w % ipc);
      if (b >= n_blocks) continue;  // long-row items read (col, val) from global
      const int buf = cnt & 1;
      mbar_wait(&sm.empty[buf], ((cnt >> 1) & 1) ^ 1);
      const int64_t r0 = (int64_t)b * kG2Rows;
      const int nr = (int)min((int64_t)kG2Rows, a.g.n_rows - r0);
      // rowptr slice: coalesced loads by the whole warp
      for (int i = lane; i <= nr; i += 32) sm.rp[buf][i] = __ldg(a.g.rowptr + r0 + i);
      __syncwarp();
      if (lane == 0) {
        const int64_t e0 = sm.rp[buf][0], e1 = sm.rp[buf][nr];
        const int64_t s0 = e0 & ~(int64_t)3;                                // 16-byte aligned window start
        int64_t s1 = min(min((e1 + 3) & ~(int64_t)3, nnz_vec), s0 + kG2Cap);
        if (s1 < s0) s1 = s0;
        const uint32_t bytes = (uint32_t)(s1 - s0) * 4u;
        // rp[] stores were made by this warp before the arrive below releases them to the gather warps
        if (bytes > 0) {
          mbar_arrive_expect_tx(&sm.full[buf], 2 * bytes);
          bulk_g2s(&sm.col[buf][0], a.g.col + s0, bytes, &sm.full[buf]);
          bulk_g2s(&sm.val[buf][0], a.g.val + s0, bytes, &sm.full[buf]);
        } else {
          mbar_arrive(&sm.full[buf]);
        }
      }
      __syncwarp();
      ++cnt;
    }
  } else {
    // ------------------------------ gather warps ------------------------------
    const int sub = threadIdx.x % LPR;
    const int g = threadIdx.x / LPR;
    uint32_t cnt = 0;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
      const int chunk = (int)(w / ipc);
      const int b = (int)(w % ipc);
      const float* __restrict__ xc = x + chunk * CW + sub * 4;
      if (b < n_blocks) {
        const int buf = cnt & 1;
        mbar_wait(&sm.full[buf], (cnt >> 1) & 1);
        const int64_t r0 = (int64_t)b * kG2Rows;
        const int nr = (int)min((int64_t)kG2Rows, a.g.n_rows - r0);
        const int e0 = sm.rp[buf][0];
        const int s0 = e0 & ~3;
        const int s1 = (int)min(min(((int64_t)sm.rp[buf][nr] + 3) & ~(int64_t)3, nnz_vec), (int64_t)s0 + kG2Cap);
        const int32_t* __restrict__ cs = &sm.col[buf][0] - s0;  // index with the global entry number
        const float* __restrict__ vs = &sm.val[buf][0] - s0;
#pragma unroll 1
        for (int rr = 0; rr < RPG; ++rr) {
          const int rl = g + rr * G;
          if (rl >= nr) break;
          const int start = sm.rp[buf][rl];
          int end = sm.rp[buf][rl + 1];
          if (end - start > kLongRow) continue;  // a long-row item produces this row
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          for (int k = start; k < end; k += U) {
            float xv[U][4];
            float vv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int idx = k + u;
              if (idx < end) {
                int cj;
                if (idx >= s0 && idx < s1) {
                  cj = cs[idx];
                  vv[u] = vs[idx];
                } else {
                  cj = __ldg(a.g.col + idx);
                  vv[u] = __ldg(a.g.val + idx);
                }
                ldv<4>(xc + (int64_t)cj * H, xv[u]);
              }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (k + u < end) {
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fmaf(vv[u], xv[u][i], acc[i]);
              }
            }
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaxf(acc[i], 0.f);
          }
          const int64_t off = (r0 + rl) * H + chunk * CW + sub * 4;
          if constexpr (kStoreOnly) __stcs(reinterpret_cast<float4*>(c.k_out + off), make_float4(acc[0], acc[1], acc[2], acc[3]));
          else epi_apply<4>(c, off, acc, err_acc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[buf]);
        ++cnt;
      } else {
        // ---- one long row: G lane groups stride over its entries, fixed-order reduction ----
        const int64_t row = __ldg(long_rows + (b - n_blocks));
        const int start = __ldg(a.g.rowptr + row), end = __ldg(a.g.rowptr + row + 1);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int idx = start + g; idx < end; idx += 4 * G) {
          float xv[4][4];
          float vv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int id = idx + u * G;
            vv[u] = 0.f;
            if (id < end) {
              vv[u] = __ldg(a.g.val + id);
              ldv<4>(xc + (int64_t)__ldg(a.g.col + id) * H, xv[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (idx + u * G < end) {
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = fmaf(vv[u], xv[u][i], acc[i]);
            }
          }
        }
        stv<4>(sm.part + g * CW + sub * 4, acc);
        named_bar_sync(1, kStageThreads);
        if (threadIdx.x < LPR) {
          float tot[4] = {0.f, 0.f, 0.f, 0.f};
          for (int gg = 0; gg < G; ++gg) {
            float p[4];
            ldv<4>(sm.part + gg * CW + sub * 4, p);
#pragma unroll
            for (int i = 0; i < 4; ++i) tot[i] += p[i];
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < 4; ++i) tot[i] = fmaxf(tot[i], 0.f);
          }
          const int64_t off = row * H + chunk * CW + sub * 4;
          if constexpr (kStoreOnly) __stcs(reinterpret_cast<float4*>(c.k_out + off), make_float4(tot[0], tot[1], tot[2], tot[3]));
          else epi_apply<4>(c, off, tot, err_acc);
        }
        named_bar_sync(1, kStageThreads);  // part[] is reused by the next long item
      }
    }
  }
  epi_finish_block(e, err_acc);
}

// =========================================================================================
// Slab gather (feature-sharded multi-GPU slice gather): the source is COLUMN-BLOCKED, [n_slabs][n_src][16] -- one
// 16-column slab is n_src * 64 bytes and contiguous (64 MB at 1M nodes), so, unlike a 16/32-column chunk of a row-major
// state whose 128-byte lines span twice that, it stays L2-resident while the grid walks it (measured with
// tests/cuda/slab_gather_probe.cu: DRAM reads 3.7 GB instead of 9.7 GB per 256-column gather, 0.12 ms per slab).
//   blockIdx.x = slab * (n_rb + n_long) + b   slab-major: the CTAs resident together read the same slab
//   b <  n_rb : 128 rows; the block's CSR slice is staged in shared memory once (coalesced, streaming), 4 lanes per
//               row (64 bytes per entry), lane groups pull rows from a CTA-local counter, 4 entries in flight per lane,
//               CSR order per row (= torch.sparse.mm's CPU order, neural_dynamics.py:29)
//   b >= n_rb : one row above kLongRow entries per CTA, fixed-order reduction over the lane groups
// z leaves through store_z_owner (FEAT_Z_OWNERS): block `rank` of the blocked Z of the rank that owns the row.
// =========================================================================================
constexpr int kSlabRows = 128;
constexpr int kSlabCap = 2048;

// `rb_shift`: the grid walks the row blocks starting at this one (wrapping around).  Every rank of the feature-sharded
// push gathers ALL rows and stores z to the rows' owners in the same order; a per-rank start (rank q at its own block)
// keeps the ranks on different owners at any moment.  Measured: no gain (see the caller), default 0.
// `spread` > 1: consecutive CTAs take row blocks of DIFFERENT owners (block index b -> owner b % spread, the owner's
// (b / spread)-th block): the CTAs resident at any moment store to all owners at once, so that no GPU's NVLink ingress
// is the target of every sender at the same time; n_rb is then spread * ceil(real blocks / spread) and the surplus
// indices do nothing.
__global__ void __launch_bounds__(kStageThreads, 6) k_gather_slab(GraphView g, const float* __restrict__ x,
                                                                  int64_t n_src, int n_rb, int n_long,
                                                                  const int32_t* __restrict__ long_rows, int rb_shift,
                                                                  int spread, int n_rb_real, EpiArgs e) {
  __shared__ __align__(16) float s_part[(kStageThreads / 4) * 16];
  __shared__ int s_col[kSlabCap];
  __shared__ float s_val[kSlabCap];
  __shared__ int s_rp[kSlabRows + 1];
  __shared__ int s_next;
  if (e.ctrl != nullptr && ((volatile Ctrl*)e.ctrl)->done) return;
  constexpr int G = kStageThreads / 4;
  const int bpc = n_rb + n_long;
  const int slab = blockIdx.x / bpc;
  const int b0 = blockIdx.x - slab * bpc;
  const bool row_block = b0 < n_rb;  // else: one of the long rows
  int b = b0;
  if (row_block) {
    if (spread > 1) {
      const int per = n_rb / spread;  // n_rb is a multiple of spread here
      b = (b0 % spread) * per + (b0 / spread);
      if (b >= n_rb_real) return;
    } else {
      b += rb_shift;
      if (b >= n_rb) b -= n_rb;
    }
  }
  const int sub = threadIdx.x & 3;
  const int gidx = threadIdx.x >> 2;
  const float* __restrict__ xs = x + (size_t)slab * n_src * 16 + sub * 4;
  const int hc_log2 = e.feat_hc_log2;
  auto emit = [&](int64_t row, const float4& v) {
    store_z_owner<4>(e.feat, e.feat_rank, hc_log2, (row << hc_log2) + slab * 16 + sub * 4, v.x, v.y, v.z, v.w);
  };
  if (row_block) {
    const int64_t r0 = (int64_t)b * kSlabRows;
    const int nr = (int)min((int64_t)kSlabRows, g.n_rows - r0);
    for (int i = threadIdx.x; i <= nr; i += kStageThreads) s_rp[i] = __ldg(g.rowptr + r0 + i);
    if (threadIdx.x == 0) s_next = G;
    __syncthreads();
    const int e0 = s_rp[0];
    const int cnt = min(s_rp[nr] - e0, kSlabCap);
    for (int i = threadIdx.x; i < cnt; i += kStageThreads) {
      s_col[i] = __ldcs(g.col + e0 + i);
      s_val[i] = __ldcs(g.val + e0 + i);
    }
    __syncthreads();
    const unsigned gmask = 0xFu << ((threadIdx.x & 31) & ~3);
    int row = gidx;
    while (row < nr) {
      const int start = s_rp[row] - e0, end = s_rp[row + 1] - e0;
      if (!(n_long > 0 && end - start > kLongRow)) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = start; k < end; k += 4) {
          float4 xv[4];
          float vv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int idx = k + u;
            if (idx < end) {
              int cj;
              if (idx < kSlabCap) { cj = s_col[idx]; vv[u] = s_val[idx]; }
              else { cj = __ldg(g.col + e0 + idx); vv[u] = __ldg(g.val + e0 + idx); }
              xv[u] = *reinterpret_cast<const float4*>(xs + (size_t)cj * 16);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (k + u < end) {
              acc.x = fmaf(vv[u], xv[u].x, acc.x); acc.y = fmaf(vv[u], xv[u].y, acc.y);
              acc.z = fmaf(vv[u], xv[u].z, acc.z); acc.w = fmaf(vv[u], xv[u].w, acc.w);
            }
          }
        }
        emit(r0 + row, acc);
      }
      int nxt = 0;
      if (sub == 0) nxt = atomicAdd(&s_next, 1);
      row = __shfl_sync(gmask, nxt, (threadIdx.x & 31) & ~3);
    }
  } else {
    const int64_t row = __ldg(long_rows + (b0 - n_rb));
    const int start = __ldg(g.rowptr + row), end = __ldg(g.rowptr + row + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int idx = start + gidx; idx < end; idx += 4 * G) {
      float4 xv[4];
      float vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int id = idx + u * G;
        vv[u] = 0.f;
        xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id < end) {
          vv[u] = __ldg(g.val + id);
          xv[u] = *reinterpret_cast<const float4*>(xs + (size_t)__ldg(g.col + id) * 16);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x = fmaf(vv[u], xv[u].x, acc.x); acc.y = fmaf(vv[u], xv[u].y, acc.y);
        acc.z = fmaf(vv[u], xv[u].z, acc.z); acc.w = fmaf(vv[u], xv[u].w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(s_part + gidx * 16 + sub * 4) = acc;
    __syncthreads();
    if (threadIdx.x < 4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int gg = 0; gg < G; ++gg) {
        const float4 p = *reinterpret_cast<const float4*>(s_part + gg * 16 + sub * 4);
        t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
      }
      emit(row, t);
    }
  }
}

}  // namespace ndcn

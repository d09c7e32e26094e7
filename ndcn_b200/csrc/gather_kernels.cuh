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

constexpr int kLongRow = 256;  // entries; rows above this get a CTA of their own per chunk

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

}  // namespace ndcn

// Persistent whole-solve kernel for small graphs (SURVEY.md section 7.1-7, section 8(b) ndcn_odeint_small_f32).
//
// BASELINE configs 1-2 integrate a few hundred to a few thousand nodes (400-node grid, H=20, 99 Euler steps;
// Cora, 2708 nodes, H=256, dopri5): every tensor of the solve fits L2 many times over and a launch-per-stage
// schedule is nothing but launch latency (round 1: 63 launches for a 3-step dopri5 solve on 400 nodes).  Here ONE
// cooperative launch runs the whole odeint(): the same right-hand-side row code, stage epilogues, controller
// arithmetic and dense output as the launch-per-stage path (the device functions are shared, so the two paths
// round identically per element), with cooperative-groups grid barriers where the launch boundaries were.  The k_i,
// the stage inputs and the state never leave L2 -- on these sizes the north star's "no intermediate k_i hits HBM"
// holds literally.
//
// Reference: torchdiffeq/_impl/solvers.py:25-33,79-99 (drivers), dopri5.py:58-122, rk_common.py:22-78,
// fixed_grid.py:5-29, misc.py:84-170, interp.py:5-65; right-hand sides neural_dynamics.py:20-39 and the three
// ground-truth dynamics.
#pragma once
#include <cooperative_groups.h>

#include "ndcn_common.cuh"
#include "solver_kernels.cuh"
#include "stage_kernels.cuh"

namespace ndcn {

namespace cg = cooperative_groups;

struct SmallArgs {
  // right-hand side
  GraphView g;
  int kind;          // NDCN_RHS_*
  int H;             // state width
  uint32_t flags;    // NDCN_F_*
  const float* W;    // [H,H] nn.Linear weight (row-major, y = x W^T + b)
  const float* Wt;   // [H,H] its transpose (tiled-GEMM instantiations: H in {32,64,128,256} with the Linear)
  const float* bias;
  float p[8];
  // solve
  int method;        // NDCN_EULER .. NDCN_DOPRI5
  int n_t;
  int terminal, forced, given_first;
  int in_slab;       // fixed grid: the output slab doubles as state storage
  const float* y0;
  float* out;
  float* Y[2];
  float* YS[2];
  float* KF[2];
  float* K[5];
  const float* dts;      // fixed grid: n_t - 1 fp32 step sizes (fp32 differences of the fp32-rounded grid)
  const double* t_out;   // dopri5: requested times, device
  Ctrl* ctrl;
  double* partials;      // >= 2 * gridDim.x doubles
  float beta32[6][8];
  float err32[8];
  float c_mid[7];
  float rtol, atol;
  double t_first;
  const float* dec_W;    // fused decoder (NDCN.output_layer) or null
  const float* dec_b;
  int dec_C;
  int64_t numel;
  int n_rows;
  int vec;               // elementwise phases may use 16-byte accesses
  int err_prefix;        // stage 5 leaves the error-estimate prefix in YS[1]
};

__device__ __forceinline__ PtrPair spp(float* a) { return PtrPair{{a, a}}; }
__device__ __forceinline__ PtrPair spp(float* a, float* b) { return PtrPair{{a, b}}; }

__device__ __forceinline__ void small_blank(EpiArgs& e) {
  // EpiArgs has no constructor (it travels as a kernel parameter): clear it field-wise
  e.mode = EPI_STORE; e.n_prev = 0; e.dt_src = DT_HOST; e.check_finite = 0;
  e.k_out = spp(nullptr); e.y_out = spp(nullptr); e.y0 = spp(nullptr); e.y1 = spp(nullptr);
#pragma unroll
  for (int j = 0; j < 6; ++j) e.kprev[j] = spp(nullptr);
#pragma unroll
  for (int j = 0; j < 8; ++j) { e.beta[j] = 0.f; e.ebeta[j] = 0.f; }
  e.ctrl = nullptr; e.dt_host = 0.f; e.rtol = 0.f; e.atol = 0.f; e.partials = nullptr;
  e.n_peers = 0;
#pragma unroll
  for (int j = 0; j < kMaxPeers; ++j) e.peer_delta[j] = 0;
  e.feat = nullptr; e.feat_mode = FEAT_OFF; e.feat_rank = 0; e.feat_row0 = 0; e.feat_hc_log2 = 0; e.feat_h_log2 = 0;
  e.e_out = nullptr; e.err_prefix = 0;
}

// descriptor copy between two shared-memory EpiArgs by the whole CTA (word-wise)
__device__ __forceinline__ void small_copy_desc(EpiArgs* dst, const EpiArgs* src) {
  static_assert(sizeof(EpiArgs) % 4 == 0, "EpiArgs is copied word-wise");
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(EpiArgs) / 4); i += blockDim.x) d[i] = s[i];
}

// the zero-coefficient streams of a stage are not read (see drop_zero_terms in ndcn_api.cu: same rule, same bits)
__device__ __forceinline__ void small_drop_zero_terms(EpiArgs& e) {
  if (e.mode != EPI_LINCOMB && e.mode != EPI_ERR && e.mode != EPI_LINCOMB_E) return;
  if (e.n_prev < 2 || e.err_prefix) return;
  const bool two = e.mode == EPI_LINCOMB_E;
  int w = 0;
  for (int j = 0; j < e.n_prev; ++j) {
    const bool zero = e.beta[j] == 0.0f && (!two || e.ebeta[j] == 0.0f);
    if (zero && !(w == 0 && j == e.n_prev - 1)) continue;
    e.kprev[w] = e.kprev[j];
    e.beta[w] = e.beta[j];
    e.ebeta[w] = e.ebeta[j];
    ++w;
  }
  if (w == e.n_prev) return;
  e.beta[w] = e.beta[e.n_prev];
  e.ebeta[w] = e.ebeta[e.n_prev];
  for (int j = w + 1; j < 8; ++j) e.beta[j] = e.ebeta[j] = 0.0f;
  e.n_prev = w;
}

// relu((Phi x) W^T + b) for one row at H <= 32: one column per lane.  The loads of a row's neighbours are all in
// flight together (the row's (col, val) pairs sit in the lanes and are broadcast by shuffle), the stage-algebra
// streams are requested BEFORE the gather so that their latency overlaps it, and W^T comes from shared memory.
// Same accumulation order as ndcn_any_row (CSR order; k = 0..H-1): bit-identical results.
__device__ __forceinline__ void ndcn_row_narrow(const NdcnArgs& a, int H, const float* __restrict__ wt_s /* [H][33] */,
                                                const float* __restrict__ x, float* zr, int64_t row, int lane,
                                                const EpiCtx& c, double& err_acc) {
  const bool act = lane < H;
  EpiIn<1> in;
  if (act) epi_load<1>(c, row * H + lane, in);
  float s = 0.f;
  if (a.flags & NDCN_F_NO_GRAPH) {
    if (act) s = x[row * H + lane];
  } else {
    const int start = a.g.rowptr[row], end = a.g.rowptr[row + 1];
    for (int base = start; base < end; base += 32) {
      int my_c = 0;
      float my_v = 0.f;
      if (base + lane < end) {
        my_c = __ldg(a.g.col + base + lane);
        my_v = __ldg(a.g.val + base + lane);
      }
      const int cnt = min(32, end - base);
      for (int j0 = 0; j0 < cnt; j0 += 8) {
        float xv[8], vv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int cj = __shfl_sync(0xffffffffu, my_c, (j0 + u) & 31);
          vv[u] = __shfl_sync(0xffffffffu, my_v, (j0 + u) & 31);
          xv[u] = (act && j0 + u < cnt) ? x[(int64_t)cj * H + lane] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (j0 + u < cnt) s = fmaf(vv[u], xv[u], s);
      }
    }
  }
  float kv[1] = {s};
  if (!(a.flags & NDCN_F_NO_CONTROL)) {
    if (act) zr[lane] = s;
    __syncwarp();
    float t = 0.f;
    if (act) {
#pragma unroll 4
      for (int k = 0; k < H; ++k) t = fmaf(zr[k], wt_s[k * 33 + lane], t);
      kv[0] = t + __ldg(a.bias + lane);
    }
    __syncwarp();
  }
  if (!(a.flags & NDCN_F_NO_RELU)) kv[0] = fmaxf(kv[0], 0.f);
  if (act) epi_math<1>(c, row * H + lane, kv, in, err_acc);
}

// one right-hand-side evaluation over all rows, fused with epilogue e (the launch-per-stage kernels' row code).
//   VW > 0, CONTROL : relu((Phi x) W^T + b) at H = 32 VW NCH through k_stage_ndcn_gemm's 64-row tiles (gather -> SMEM
//                     tile -> FP32-FMA GEMM with W^T streamed by cp.async.bulk -> bias/ReLU -> epilogue); `chunk_it` =
//                     the CTA's running W^T chunk count (mbarrier phases)
//   VW > 0, !CONTROL: no_control at the same widths: k_stage_ndcn_row's vectorised warp-per-row gather
//   VW == 0         : one warp per row, any width, any right-hand side (H <= 32: ndcn_row_narrow)
// `e` lives in shared memory (built by thread 0, published by a block barrier): no per-thread copies.
template <int VW, int NCH, bool CONTROL>
__device__ __noinline__ void small_stage(const SmallArgs& a, PtrPair src, const EpiArgs& e, float* zs,
                                         const float* wt_s, uint32_t& chunk_it) {
  EpiCtx c;
  if (!epi_resolve(e, c)) return;  // finished solve: uniform over the grid
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ x = sel(src, par);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double err_acc = 0.0;
  const int64_t w0 = (int64_t)blockIdx.x * kWarpsPerCta + warp, wn = (int64_t)gridDim.x * kWarpsPerCta;
  if (a.kind == NDCN_RHS_NDCN) {
    NdcnArgs na;
    na.g = a.g; na.x = src; na.Wt = a.Wt; na.bias = a.bias; na.flags = a.flags; na.long_rows = nullptr; na.n_long = 0;
    na.row_begin = 0; na.row_end = 0; na.keep_l2 = 0;
    if constexpr (VW > 0 && CONTROL) {
      const int64_t n_tiles = ((int64_t)a.n_rows + kTileRows - 1) / kTileRows;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        ndcn_gemm_tile<VW, NCH>(na, c, x, tile * kTileRows, reinterpret_cast<unsigned char*>(zs), chunk_it, err_acc);
    } else if constexpr (VW > 0) {
      constexpr int H = 32 * VW * NCH;
      const bool relu = !(a.flags & NDCN_F_NO_RELU);
      for (int64_t row = w0; row < a.n_rows; row += wn) {
        float acc[NCH][VW];
        if (a.flags & NDCN_F_NO_GRAPH) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) ldv<VW>(x + row * H + ch * 32 * VW + lane * VW, acc[ch]);
        } else {
          gather_row<VW, NCH>(a.g, row, x, lane, acc);
        }
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          if (relu) {
#pragma unroll
            for (int i = 0; i < VW; ++i) acc[ch][i] = fmaxf(acc[ch][i], 0.f);
          }
          epi_apply<VW>(c, row * H + ch * 32 * VW + lane * VW, acc[ch], err_acc);
        }
      }
    } else if (a.H <= 32) {
      for (int64_t row = w0; row < a.n_rows; row += wn) ndcn_row_narrow(na, a.H, wt_s, x, zs + warp * 32, row, lane, c, err_acc);
    } else {
      for (int64_t row = w0; row < a.n_rows; row += wn) ndcn_any_row(na, a.H, a.W, x, zs + warp * a.H, row, lane, c, err_acc);
    }
  } else {
    DynArgs da;
    da.g = a.g; da.x = src; da.kind = a.kind; da.d = a.H; da.long_rows = nullptr; da.n_long = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) da.p[i] = a.p[i];
    if (a.H == 1) {
      // k_stage_dyn1's row code: 4 lanes per row, 8 rows per warp
      const int64_t n_w = ((int64_t)a.n_rows + 7) / 8;
      for (int64_t w = w0; w < n_w; w += wn) {
        if (a.kind == NDCN_RHS_HEAT) dyn1_warp_rows<NDCN_RHS_HEAT, 4>(da, x, w, lane, c, err_acc);
        else if (a.kind == NDCN_RHS_GENE) dyn1_warp_rows<NDCN_RHS_GENE, 4>(da, x, w, lane, c, err_acc);
        else dyn1_warp_rows<NDCN_RHS_MUTUAL, 4>(da, x, w, lane, c, err_acc);
      }
    } else {
      for (int64_t row = w0; row < a.n_rows; row += wn) {
        if (a.kind == NDCN_RHS_HEAT) dynv_row<NDCN_RHS_HEAT>(da, x, row, lane, c, err_acc);
        else if (a.kind == NDCN_RHS_GENE) dynv_row<NDCN_RHS_GENE>(da, x, row, lane, c, err_acc);
        else dynv_row<NDCN_RHS_MUTUAL>(da, x, row, lane, c, err_acc);
      }
    }
  }
  epi_finish_block(e, err_acc);
}

// TINY: a whole [N,d] ground-truth solve on ONE CTA (<= 8192 state elements: the [400,1] states of the dynamics
// scripts, heat_dynamics.py:207-209).  Element-parallel -- thread e owns element (r, c) = (e / d, e % d) and walks its
// row's entries in CSR order -- with block barriers in place of the grid barriers: these solves take ~60-140 dopri5
// steps of 7-9 phases whose work is a few hundred rows, so the barrier is the cost (measured: 1.71 ms against 2.84 ms
// for the cooperative variant and 2.61 ms for 565 launches, 374 RHS evaluations of the heat ground truth).
// (The NDCN right-hand side at H = 20 was tried the same way and lost -- 3.9 ms against 0.85 ms for 99 Euler steps:
// one SM's 512 threads serialise 8000 x 28 multiply-adds per stage behind L2-latency loads of a state the CTA has
// just written, where 50 cooperative CTAs spread them -- so NDCN stays on the cooperative variant.)
__device__ __noinline__ void tiny_stage(const SmallArgs& a, PtrPair src, const EpiArgs& e) {
  EpiCtx c;
  if (!epi_resolve(e, c)) return;
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ x = sel(src, par);
  double err_acc = 0.0;
  const int d = a.H;
  const int numel = (int)a.numel;
  for (int el = threadIdx.x; el < numel; el += blockDim.x) {
    const int r = el / d, cc = el - r * d;
    const float xi = x[el];
    const int start = a.g.rowptr[r], end = a.g.rowptr[r + 1];
    float s = 0.f;
    for (int j = start; j < end; ++j) {
      const float xj = x[__ldg(a.g.col + j) * d + cc];
      const float av = __ldg(a.g.val + j);
      if (a.kind == NDCN_RHS_HEAT) s = fadd(s, dyn_neighbour<NDCN_RHS_HEAT>(a.p, av, xi, xj, d == 1));
      else if (a.kind == NDCN_RHS_GENE) s = fadd(s, dyn_neighbour<NDCN_RHS_GENE>(a.p, av, xi, xj, d == 1));
      else s = fadd(s, dyn_neighbour<NDCN_RHS_MUTUAL>(a.p, av, xi, xj, d == 1));
    }
    float kv[1];
    if (a.kind == NDCN_RHS_HEAT) kv[0] = dyn_local<NDCN_RHS_HEAT>(a.p, xi, s);
    else if (a.kind == NDCN_RHS_GENE) kv[0] = dyn_local<NDCN_RHS_GENE>(a.p, xi, s);
    else kv[0] = dyn_local<NDCN_RHS_MUTUAL>(a.p, xi, s);
    epi_apply<1>(c, el, kv, err_acc);
  }
  epi_finish_block(e, err_acc);
}

__device__ __noinline__ void small_epi_only(const SmallArgs& a, PtrPair k_in, const EpiArgs& e) {
  EpiCtx c;
  if (!epi_resolve(e, c)) return;
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  double err_acc = 0.0;
  const bool bad = epi_only_range(sel(k_in, par), a.numel, c, e.check_finite, a.vec,
                                  (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, err_acc);
  if (e.check_finite && bad && e.ctrl) atomicExch(&e.ctrl->status, NDCN_E_NONFINITE);
}

__device__ __forceinline__ void small_copy(const float* __restrict__ src, float* __restrict__ dst, int64_t numel) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// out slice `slot` <- state y: a copy, or y W_d^T + b_d with the fused decoder
__device__ __forceinline__ void small_put_state(const SmallArgs& a, int64_t slot, const float* y) {
  if (a.dec_C > 0) {
    decode_rows_range(y, a.n_rows, a.H, a.dec_W, a.dec_b, a.dec_C, a.out + slot * (int64_t)a.n_rows * a.dec_C,
                      ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, ((int64_t)gridDim.x * blockDim.x) >> 5,
                      threadIdx.x & 31);
  } else {
    float* dst = a.out + slot * a.numel;
    if (dst != y) small_copy(y, dst, a.numel);
  }
}

// fixed-order sum of the per-CTA partials by block 0 (what k_controller / k_init_scalar do in their own launch)
__device__ __forceinline__ double small_sum_partials(const double* partials, int n, int stride2, int q, double* s_tmp) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[stride2 * i + q];
  s_tmp[threadIdx.x] = v;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_tmp[threadIdx.x] += s_tmp[threadIdx.x + o];
    __syncthreads();
  }
  const double r = s_tmp[0];
  __syncthreads();
  return r;
}

// thread 0 edits the stage descriptor in shared memory; PUBLISH makes it visible to the CTA (the grid barrier
// that follows every stage keeps the next edit from overtaking a reader)
#define NDCN_T0 if (threadIdx.x == 0)
#define NDCN_PUBLISH()                                                                        \
  do {                                                                                        \
    __syncthreads();                                                                          \
    small_copy_desc(&s_run, &s_e); /* all threads, one word each: not 110 serial copies */   \
    __syncthreads();                                                                          \
    NDCN_T0 {                                                                                 \
      s_run.partials = a.partials;                                                            \
      small_drop_zero_terms(s_run);                                                           \
    }                                                                                         \
    __syncthreads();                                                                          \
  } while (0)

constexpr int kTinyThreads = 512;
constexpr int kTinyMaxNumel = 8192;

template <int VW, int NCH, bool CONTROL, bool TINY = false>
__global__ void __launch_bounds__(TINY ? kTinyThreads : kStageThreads, (TINY || (VW > 0 && CONTROL)) ? 1 : 2)
k_solve_small(const __grid_constant__ SmallArgs a) {
  // VW == 0: [warps][max(H, 32)] scratch rows of the row-per-warp right-hand side; tiled GEMM: GemmSmem<VW, NCH>
  extern __shared__ __align__(128) float zs[];
  __shared__ double s_tmp[TINY ? kTinyThreads : kStageThreads];
  __shared__ float s_xs[kEmitMaxPerLaunch * 4];
  __shared__ float s_wt[32 * 33];  // W^T of a narrow Linear (H <= 32), padded rows
  __shared__ EpiArgs s_e, s_run;
  cg::grid_group grid = cg::this_grid();
  Ctrl* ctrl = a.ctrl;
  uint32_t chunk_it = 0;
  if constexpr (VW > 0 && CONTROL) ndcn_gemm_init_bars<VW, NCH>(reinterpret_cast<unsigned char*>(zs));
  if (VW == 0 && a.kind == NDCN_RHS_NDCN && a.H <= 32 && !(a.flags & NDCN_F_NO_CONTROL)) {
    for (int i = threadIdx.x; i < a.H * a.H; i += blockDim.x) {
      const int n = i / a.H, k = i % a.H;
      s_wt[k * 33 + n] = a.W[i];  // W[n][k] -> W^T[k][n]
    }
  }
  __syncthreads();
#define NDCN_STAGE(src)                                                             \
  do {                                                                              \
    if constexpr (TINY) tiny_stage(a, src, s_run);                        \
    else small_stage<VW, NCH, CONTROL>(a, src, s_run, zs, s_wt, chunk_it);          \
  } while (0)
#define NDCN_BAR()                                  \
  do {                                              \
    if constexpr (TINY) __syncthreads();            \
    else grid.sync();                               \
  } while (0)

  if (a.method != NDCN_DOPRI5) {
    // ---------------- fixed grid: euler / midpoint / rk4 (3/8 rule), solvers.py:79-99 ----------------
    float* cur = a.in_slab ? a.out : a.Y[0];
    small_copy(a.y0, cur, a.numel);
    if (!a.in_slab && !a.terminal) small_put_state(a, 0, a.y0);
    NDCN_BAR();
    for (int i = 0; i + 1 < a.n_t; ++i) {
      float* nxt = a.in_slab ? a.out + (int64_t)(i + 1) * a.numel : a.Y[(i + 1) & 1];
      NDCN_T0 {
        small_blank(s_e);
        s_e.dt_src = DT_HOST;
        s_e.dt_host = a.dts[i];
        s_e.y0 = spp(cur);
      }
      if (a.method == NDCN_EULER) {  // fixed_grid.py:7-8
        NDCN_T0 { s_e.mode = EPI_LINCOMB; s_e.beta[0] = 1.0f; s_e.y_out = spp(nxt); }
        NDCN_PUBLISH();
        NDCN_STAGE(spp(cur));
        NDCN_BAR();
      } else if (a.method == NDCN_MIDPOINT) {  // fixed_grid.py:17-20
        NDCN_T0 { s_e.mode = EPI_LINCOMB; s_e.beta[0] = 0.5f; s_e.y_out = spp(a.YS[0]); }
        NDCN_PUBLISH();
        NDCN_STAGE(spp(cur));
        NDCN_BAR();
        NDCN_T0 { s_e.beta[0] = 1.0f; s_e.y_out = spp(nxt); }
        NDCN_PUBLISH();
        NDCN_STAGE(spp(a.YS[0]));
        NDCN_BAR();
      } else {  // rk_common.py:72-78
        NDCN_T0 { s_e.mode = EPI_RK4_1; s_e.k_out = spp(a.K[0]); s_e.y_out = spp(a.YS[0]); }
        NDCN_PUBLISH();
        NDCN_STAGE(spp(cur));
        NDCN_BAR();
        NDCN_T0 { s_e.mode = EPI_RK4_2; s_e.k_out = spp(a.K[1]); s_e.kprev[0] = spp(a.K[0]); s_e.y_out = spp(a.YS[1]); }
        NDCN_PUBLISH();
        NDCN_STAGE(spp(a.YS[0]));
        NDCN_BAR();
        NDCN_T0 { s_e.mode = EPI_RK4_3; s_e.k_out = spp(a.K[2]); s_e.kprev[1] = spp(a.K[1]); s_e.y_out = spp(a.YS[0]); }
        NDCN_PUBLISH();
        NDCN_STAGE(spp(a.YS[1]));
        NDCN_BAR();
        NDCN_T0 { s_e.mode = EPI_RK4_4; s_e.k_out = spp(nullptr); s_e.kprev[2] = spp(a.K[2]); s_e.y_out = spp(nxt); }
        NDCN_PUBLISH();
        NDCN_STAGE(spp(a.YS[0]));
        NDCN_BAR();
      }
      // the new state goes to its output slot while the next step already reads it (nobody writes `nxt` again
      // before the grid barrier that ends the next step)
      if (!a.in_slab && !a.terminal) small_put_state(a, i + 1, nxt);
      cur = nxt;
    }
    if (a.terminal) small_put_state(a, 0, cur);
    return;
  }

  // ---------------- dopri5: solvers.py:25-33, dopri5.py:58-122 ----------------
  const PtrPair Ycur = spp(a.Y[0], a.Y[1]), Yoth = spp(a.Y[1], a.Y[0]);
  const PtrPair KFcur = spp(a.KF[0], a.KF[1]), KFoth = spp(a.KF[1], a.KF[0]);
  small_copy(a.y0, a.Y[0], a.numel);
  if (!a.terminal) small_put_state(a, 0, a.y0);
  NDCN_T0 { small_blank(s_e); s_e.k_out = spp(a.KF[0]); }  // f0 = func(t0, y0)     dopri5.py:78
  NDCN_PUBLISH();
  NDCN_BAR();
  NDCN_STAGE(spp(a.Y[0]));
  NDCN_BAR();
  if (!a.forced && !a.given_first) {
    // _select_initial_step(order=4)     dopri5.py:80, misc.py:84-143
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    init_norms_block(a.Y[0], a.KF[0], nullptr, a.numel, a.rtol, a.atol, 0, a.partials, tid, stride);
    NDCN_BAR();
    if (blockIdx.x == 0) {
      const double s0 = small_sum_partials(a.partials, gridDim.x, 2, 0, s_tmp);
      const double s1 = small_sum_partials(a.partials, gridDim.x, 2, 1, s_tmp);
      if (threadIdx.x == 0) { init_scalar_decide(*ctrl, s0, s1, 0, a.t_first); __threadfence(); }
    }
    NDCN_BAR();
    NDCN_T0 {
      small_blank(s_e);
      s_e.ctrl = ctrl;
      s_e.dt_src = DT_CTRL_H0;
      s_e.mode = EPI_LINCOMB;
      s_e.beta[0] = 1.0f;
      s_e.y0 = spp(a.Y[0]);
      s_e.y_out = spp(a.YS[0]);
    }
    NDCN_PUBLISH();
    small_epi_only(a, spp(a.KF[0]), s_run);  // y0 + h0*f0
    NDCN_BAR();
    NDCN_T0 { small_blank(s_e); s_e.k_out = spp(a.K[0]); }
    NDCN_PUBLISH();
    NDCN_STAGE(spp(a.YS[0]));
    NDCN_BAR();
    init_norms_block(a.Y[0], a.KF[0], a.K[0], a.numel, a.rtol, a.atol, 1, a.partials, tid, stride);
    NDCN_BAR();
    if (blockIdx.x == 0) {
      const double s0 = small_sum_partials(a.partials, gridDim.x, 2, 0, s_tmp);
      if (threadIdx.x == 0) { init_scalar_decide(*ctrl, s0, 0.0, 1, a.t_first); __threadfence(); }
    }
    NDCN_BAR();
  }

  EmitArgs em;
  em.ctrl = ctrl;
  em.t_out = a.t_out;
  em.y0 = Ycur;
  em.y1 = Yoth;
  em.k0 = KFcur;
  em.k6 = KFoth;
#pragma unroll
  for (int j = 0; j < 5; ++j) em.k[j] = a.K[j];
#pragma unroll
  for (int j = 0; j < 7; ++j) em.c_mid[j] = a.c_mid[j];
  em.out = a.out;
  em.numel = a.numel;
  em.dec_W = a.dec_C > 0 ? a.dec_W : nullptr;
  em.dec_b = a.dec_C > 0 ? a.dec_b : nullptr;
  em.dec_C = a.dec_C;
  em.H = a.H;
  em.n_rows = a.n_rows;

  for (;;) {
    if (((volatile Ctrl*)ctrl)->done) break;  // written before the last grid barrier: uniform
    // stage input 1: y0 + (dt*b10) k0, k0 = FSAL derivative; also the finite-state guard (dopri5.py:101-102)
    NDCN_T0 {
      small_blank(s_e);
      s_e.ctrl = ctrl;
      s_e.dt_src = DT_CTRL;
      s_e.y0 = Ycur;
      s_e.mode = EPI_LINCOMB;
      s_e.n_prev = 0;
      s_e.beta[0] = a.beta32[0][0];
      s_e.check_finite = 1;
      s_e.y_out = spp(a.YS[0]);
    }
    NDCN_PUBLISH();
    small_epi_only(a, KFcur, s_run);
    NDCN_BAR();
    NDCN_T0 { s_e.check_finite = 0; s_e.kprev[0] = KFcur; }
    for (int s = 1; s <= 5; ++s) {
      NDCN_T0 {
        s_e.n_prev = s;
        for (int j = 0; j < 8; ++j) s_e.beta[j] = a.beta32[s][j];
        s_e.k_out = spp(a.K[s - 1]);
        if (s >= 2) s_e.kprev[s - 1] = spp(a.K[s - 2]);
        s_e.y_out = (s < 5) ? spp(a.YS[s & 1]) : Yoth;  // stage 5 forms y1 (FSAL: c_sol == beta[-1])
        if (s == 5 && a.err_prefix) {
          s_e.mode = EPI_LINCOMB_E;
          s_e.e_out = a.YS[1];
          for (int j = 0; j < 8; ++j) s_e.ebeta[j] = j < 6 ? a.err32[j] : 0.0f;
        }
      }
      NDCN_PUBLISH();
      NDCN_STAGE(spp(a.YS[(s - 1) & 1]));
      NDCN_BAR();
    }
    // stage 6: k7 = f(y1) + error estimate
    NDCN_T0 {
      s_e.mode = EPI_ERR;
      s_e.e_out = nullptr;
      if (a.err_prefix) {
        s_e.n_prev = 1;
        s_e.err_prefix = 1;
        for (int j = 0; j < 8; ++j) s_e.beta[j] = 0.0f;
        s_e.beta[1] = a.err32[6];
        s_e.kprev[0] = spp(a.YS[1]);
      } else {
        s_e.n_prev = 6;
        for (int j = 0; j < 8; ++j) s_e.beta[j] = a.err32[j];
        s_e.kprev[5] = spp(a.K[4]);
      }
      s_e.k_out = KFoth;
      s_e.y_out = spp(nullptr);
      s_e.y1 = Yoth;
      s_e.rtol = a.rtol;
      s_e.atol = a.atol;
    }
    NDCN_PUBLISH();
    NDCN_STAGE(Yoth);
    NDCN_BAR();
    // accept / reject + next step size: block 0 (k_controller's arithmetic)
    if (blockIdx.x == 0) {
      if (((volatile Ctrl*)ctrl)->status != 0) {  // non-finite state flagged by the pre-stage
        if (threadIdx.x == 0) { ctrl->done = 1; ctrl->emit_lo = ctrl->emit_hi; }
      } else {
        const double sum = small_sum_partials(a.partials, gridDim.x, 1, 0, s_tmp);
        if (threadIdx.x == 0) controller_decide(*ctrl, sum, a.t_out);
      }
      __threadfence();
    }
    NDCN_BAR();
    // dense output of the step just accepted, for the requested times inside it; reads y0/y1/k_j only, the next
    // attempt's pre-stage writes YS[0] only: no barrier needed in between
    if (a.dec_C > 0) emit_decode_range(em, ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, ((int64_t)gridDim.x * blockDim.x) >> 5);
    else emit_range(em, a.vec, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, s_xs);
  }
#undef NDCN_STAGE
#undef NDCN_BAR
}
#undef NDCN_T0
#undef NDCN_PUBLISH

// =========================================================================================================
// Persistent discrete adjoint of the fixed-grid solvers for the narrow widths of the dynamics scripts (H <= 32):
// the backward pass of `loss.backward()` through odeint (heat_dynamics.py:317-334, `--method euler` default) as ONE
// cooperative launch.  For every grid step, last to first, the step's stages are recomputed from the saved state
// slab and each RHS evaluation k = relu((Phi x) W^T + b) is differentiated in two row-local phases
//   P1  z = Phi x, mask = (pre-activation > 0), gp = scale * gk (.) mask, u = gp W, dW += gp^T z, db += sum gp
//   P2  out = sum of the given addends + Phi^T u
// (what autograd records through solvers.py:79-99 / rk_common.py:72-78, up to fp32 summation order).  The dW / db
// contributions stay in the warps' registers over ALL steps and are reduced once, in warp order, at the end.
// =========================================================================================================
struct AdjArgs {
  GraphView g, gt;        // Phi and Phi^T (pass g twice for a symmetric operator)
  int H, n_rows, n_t, method;
  uint32_t flags;
  const float* W;         // [H,H]
  const float* bias;      // [H]
  const float* dts;       // [n_t - 1]
  const float* slab;      // [n_t, N, H] forward states
  const float* g_slab;    // [n_t, N, H] cotangents of the outputs
  float* lam;             // [N, H] running adjoint; result = dL/dy0
  float* U;               // [N, H] scratch
  float* S[8];            // scratch states: K1, K2, Y2, Y3, Y4, G4, G3, G2 (rk4) / YM, GM (midpoint)
  float* part;            // [n_warps][H*H + H] per-warp dW / db contributions
  float* dW;              // [H,H] out
  float* db;              // [H] out
  int64_t numel;
};

struct AdjP1 { const float* x; const float* t[4]; float c[4]; int nt; float scale; };
struct AdjP2 { float* out; const float* add[5]; int na; };

__device__ __forceinline__ float adj_gather_narrow(const GraphView& g, const float* __restrict__ x, int H, int64_t row,
                                                   int lane, bool act) {
  float s = 0.f;
  const int start = g.rowptr[row], end = g.rowptr[row + 1];
  for (int base = start; base < end; base += 32) {
    int my_c = 0;
    float my_v = 0.f;
    if (base + lane < end) {
      my_c = __ldg(g.col + base + lane);
      my_v = __ldg(g.val + base + lane);
    }
    const int cnt = min(32, end - base);
    for (int j0 = 0; j0 < cnt; j0 += 8) {
      float xv[8], vv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int cj = __shfl_sync(0xffffffffu, my_c, (j0 + u) & 31);
        vv[u] = __shfl_sync(0xffffffffu, my_v, (j0 + u) & 31);
        xv[u] = (act && j0 + u < cnt) ? x[(int64_t)cj * H + lane] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (j0 + u < cnt) s = fmaf(vv[u], xv[u], s);
    }
  }
  return s;
}

__global__ void __launch_bounds__(kStageThreads, 2) k_adjoint_small(const __grid_constant__ AdjArgs a,
                                                                    const __grid_constant__ SmallArgs fw) {
  extern __shared__ __align__(128) float zs[];  // [warps][64]: z row | gp row
  __shared__ float s_wt[32 * 33];               // W^T[k][n]
  __shared__ float s_w[32 * 33];                // W[o][i]
  __shared__ EpiArgs s_e, s_run;
  __shared__ AdjP1 s_p1;
  __shared__ AdjP2 s_p2;
  cg::grid_group grid = cg::this_grid();
  const int H = a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool act = lane < H;
  const bool no_graph = a.flags & NDCN_F_NO_GRAPH, no_control = a.flags & NDCN_F_NO_CONTROL;
  if (!no_control) {
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
      const int o = i / H, k = i % H;
      s_wt[k * 33 + o] = a.W[i];
      s_w[o * 33 + k] = a.W[i];
    }
  }
  __syncthreads();
  const int64_t w0 = (int64_t)blockIdx.x * kWarpsPerCta + warp, wn = (int64_t)gridDim.x * kWarpsPerCta;
  float* zr = zs + warp * 64;
  float* gpr = zr + 32;
  float dWacc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) dWacc[i] = 0.f;
  float dbacc = 0.f;
  uint32_t chunk_it = 0;

  // P1 over all rows, parameters in s_p1
  auto phase1 = [&]() {
    const AdjP1 p = s_p1;
    for (int64_t row = w0; row < a.n_rows; row += wn) {
      float z = 0.f;
      if (no_graph) { if (act) z = p.x[row * H + lane]; }
      else z = adj_gather_narrow(a.g, p.x, H, row, lane, act);
      float gk = 0.f;
      if (act) {
        for (int m = 0; m < p.nt; ++m) gk = fmaf(p.c[m], p.t[m][row * H + lane], gk);
      }
      if (act) zr[lane] = z;
      __syncwarp();
      float pre = z;
      if (!no_control && act) {
        float t = 0.f;
#pragma unroll 4
        for (int k = 0; k < H; ++k) t = fmaf(zr[k], s_wt[k * 33 + lane], t);
        pre = t + __ldg(a.bias + lane);
      }
      const float gp = (act && pre > 0.f) ? p.scale * gk : 0.f;
      float u = gp;
      if (!no_control) {
        if (act) gpr[lane] = gp;
        __syncwarp();
        if (act) {
          float t = 0.f;
#pragma unroll 4
          for (int o = 0; o < H; ++o) t = fmaf(gpr[o], s_w[o * 33 + lane], t);
          u = t;
          // dW[o = lane][i] += gp[o] z[i]
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < H) dWacc[i] = fmaf(gp, zr[i], dWacc[i]);
          dbacc += gp;
        }
      }
      if (act) a.U[row * H + lane] = u;
      __syncwarp();
    }
  };
  // P2 over all rows: out = sum of addends + Phi^T U
  auto phase2 = [&]() {
    const AdjP2 p = s_p2;
    for (int64_t row = w0; row < a.n_rows; row += wn) {
      float s = no_graph ? (act ? a.U[row * H + lane] : 0.f) : adj_gather_narrow(a.gt, a.U, H, row, lane, act);
      if (act) {
        for (int m = 0; m < p.na; ++m) s += p.add[m][row * H + lane];
        p.out[row * H + lane] = s;
      }
    }
  };
#define ADJ_T0 if (threadIdx.x == 0)
#define ADJ_SYNC() __syncthreads()
#define ADJ_STAGE(src)                                                                    \
  do {                                                                                    \
    __syncthreads();                                                                      \
    small_copy_desc(&s_run, &s_e);                                                        \
    __syncthreads();                                                                      \
    small_stage<0, 0, false>(fw, src, s_run, zs, s_wt, chunk_it);                         \
    grid.sync();                                                                          \
  } while (0)

  // lam = g_slab[n_t - 1]
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.numel; i += (int64_t)gridDim.x * blockDim.x)
    a.lam[i] = a.g_slab[(int64_t)(a.n_t - 1) * a.numel + i];
  grid.sync();

  for (int i = a.n_t - 2; i >= 0; --i) {
    const float* y = a.slab + (int64_t)i * a.numel;
    const float* gi = a.g_slab + (int64_t)i * a.numel;
    const float dt = a.dts[i];
    if (a.method == NDCN_EULER) {  // y' = y + dt f(y)
      ADJ_T0 { s_p1.x = y; s_p1.t[0] = a.lam; s_p1.c[0] = 1.f; s_p1.nt = 1; s_p1.scale = dt; }
      ADJ_SYNC();
      phase1();
      grid.sync();
      ADJ_T0 { s_p2.out = a.lam; s_p2.add[0] = a.lam; s_p2.add[1] = gi; s_p2.na = 2; }
      ADJ_SYNC();
      phase2();
      grid.sync();
    } else if (a.method == NDCN_MIDPOINT) {  // y' = y + dt f(ym), ym = y + dt/2 f(y)
      float* YM = a.S[0];
      float* GM = a.S[1];
      ADJ_T0 {
        small_blank(s_e);
        s_e.dt_src = DT_HOST; s_e.dt_host = dt; s_e.y0 = spp(const_cast<float*>(y));
        s_e.mode = EPI_LINCOMB; s_e.beta[0] = 0.5f; s_e.y_out = spp(YM);
      }
      ADJ_STAGE(spp(const_cast<float*>(y)));
      ADJ_T0 { s_p1.x = YM; s_p1.t[0] = a.lam; s_p1.c[0] = 1.f; s_p1.nt = 1; s_p1.scale = dt; }
      ADJ_SYNC();
      phase1();
      grid.sync();
      ADJ_T0 { s_p2.out = GM; s_p2.na = 0; }
      ADJ_SYNC();
      phase2();
      grid.sync();
      ADJ_T0 { s_p1.x = y; s_p1.t[0] = GM; s_p1.c[0] = 1.f; s_p1.nt = 1; s_p1.scale = dt * 0.5f; }
      ADJ_SYNC();
      phase1();
      grid.sync();
      ADJ_T0 { s_p2.out = a.lam; s_p2.add[0] = a.lam; s_p2.add[1] = GM; s_p2.add[2] = gi; s_p2.na = 3; }
      ADJ_SYNC();
      phase2();
      grid.sync();
    } else {  // rk4 = 3/8 rule, rk_common.py:72-78
      float *K1 = a.S[0], *K2 = a.S[1], *Y2 = a.S[2], *Y3 = a.S[3], *Y4 = a.S[4], *G4 = a.S[5], *G3 = a.S[6], *G2 = a.S[7];
      ADJ_T0 {
        small_blank(s_e);
        s_e.dt_src = DT_HOST; s_e.dt_host = dt; s_e.y0 = spp(const_cast<float*>(y));
        s_e.mode = EPI_RK4_1; s_e.k_out = spp(K1); s_e.y_out = spp(Y2);
      }
      ADJ_STAGE(spp(const_cast<float*>(y)));
      ADJ_T0 { s_e.mode = EPI_RK4_2; s_e.k_out = spp(K2); s_e.kprev[0] = spp(K1); s_e.y_out = spp(Y3); }
      ADJ_STAGE(spp(Y2));
      ADJ_T0 { s_e.mode = EPI_RK4_3; s_e.k_out = spp(nullptr); s_e.kprev[1] = spp(K2); s_e.y_out = spp(Y4); }
      ADJ_STAGE(spp(Y3));
      // g4 = dL/dy4 through k4:  y' = y + (k1 + 3 k2 + 3 k3 + k4) dt/8
      ADJ_T0 { s_p1.x = Y4; s_p1.t[0] = a.lam; s_p1.c[0] = 1.f; s_p1.nt = 1; s_p1.scale = dt / 8.f; }
      ADJ_SYNC(); phase1(); grid.sync();
      ADJ_T0 { s_p2.out = G4; s_p2.na = 0; }
      ADJ_SYNC(); phase2(); grid.sync();
      // k3 feeds y' (3 dt/8) and y4 (dt)
      ADJ_T0 { s_p1.x = Y3; s_p1.t[0] = a.lam; s_p1.c[0] = 3.f * dt / 8.f; s_p1.t[1] = G4; s_p1.c[1] = dt; s_p1.nt = 2; s_p1.scale = 1.f; }
      ADJ_SYNC(); phase1(); grid.sync();
      ADJ_T0 { s_p2.out = G3; s_p2.na = 0; }
      ADJ_SYNC(); phase2(); grid.sync();
      // k2 feeds y' (3 dt/8), y4 (-dt) and y3 (dt)
      ADJ_T0 {
        s_p1.x = Y2; s_p1.t[0] = a.lam; s_p1.c[0] = 3.f * dt / 8.f; s_p1.t[1] = G4; s_p1.c[1] = -dt; s_p1.t[2] = G3; s_p1.c[2] = dt;
        s_p1.nt = 3; s_p1.scale = 1.f;
      }
      ADJ_SYNC(); phase1(); grid.sync();
      ADJ_T0 { s_p2.out = G2; s_p2.na = 0; }
      ADJ_SYNC(); phase2(); grid.sync();
      // k1 feeds y' (dt/8), y4 (dt), y3 (-dt/3) and y2 (dt/3)
      ADJ_T0 {
        s_p1.x = y; s_p1.t[0] = a.lam; s_p1.c[0] = dt / 8.f; s_p1.t[1] = G4; s_p1.c[1] = dt; s_p1.t[2] = G3; s_p1.c[2] = -dt / 3.f;
        s_p1.t[3] = G2; s_p1.c[3] = dt / 3.f; s_p1.nt = 4; s_p1.scale = 1.f;
      }
      ADJ_SYNC(); phase1(); grid.sync();
      ADJ_T0 { s_p2.out = a.lam; s_p2.add[0] = a.lam; s_p2.add[1] = G4; s_p2.add[2] = G3; s_p2.add[3] = G2; s_p2.add[4] = gi; s_p2.na = 5; }
      ADJ_SYNC(); phase2(); grid.sync();
    }
  }
  // parameter gradients: every warp parks its contribution, then a fixed-order sum over the warps
  if (!no_control) {
    float* mine = a.part + (size_t)w0 * (H * H + H);
    if (act) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < H) mine[lane * H + i] = dWacc[i];
      mine[H * H + lane] = dbacc;
    }
    grid.sync();
    const int total = H * H + H;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
      float sum = 0.f;
      for (int64_t w = 0; w < wn; ++w) sum += a.part[(size_t)w * total + j];
      if (j < H * H) a.dW[j] = sum;
      else a.db[j - H * H] = sum;
    }
  }
#undef ADJ_T0
#undef ADJ_SYNC
#undef ADJ_STAGE
}

}  // namespace ndcn

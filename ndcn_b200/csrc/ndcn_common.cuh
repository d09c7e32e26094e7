// Shared device-side definitions for libndcn_b200 (sm_100a).
//
// Rounding discipline: everything that mirrors the reference's *solver algebra*
// (torchdiffeq/_impl/{rk_common,misc,interp,dopri5}.py) is written with explicit
// __fmul_rn/__fadd_rn so that nvcc cannot contract it into FMAs: the reference issues one
// full-tensor mul and one full-tensor add per term, each rounded to fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ndcn_b200.h"

namespace ndcn {

constexpr int kStageThreadsCtl = 256;  // block size of the single-block scalar kernels
constexpr int kStageThreads = 256;     // block size of every stage / elementwise kernel
constexpr int kWarpsPerCta = kStageThreads / 32;
constexpr int kMaxPeers = 7;           // other ranks of one NVSwitch domain (8 GPUs)

// ------------------------------------------------------------------------------------
// Device-resident controller block: the adaptive solver's scalars never visit the host
// between polls (the reference syncs >=3 times per step, dopri5.py:88,100-102,109).
// ------------------------------------------------------------------------------------
struct Ctrl {
  double t0, t1, dt;                 // _RungeKuttaState.t0/.t1/.dt  (rk_common.py:8-19), float64
  double emit_t0, emit_t1, emit_dt;  // interval of the step whose dense output is pending
  double rtol, atol;
  double safety, ifactor, dfactor;   // dopri5.py:71-73 (already fp32-rounded by the host)
  double forced_dt;
  double first_step;
  double sum_sq;                     // last sum of squared error ratios (diagnostic)
  double numel_global;               // element count of the (global) state, for the mean
  long long n_accept, n_reject, n_attempt;
  long long steps_this_interval, max_num_steps;
  float h0, d0, d1, msr_last;        // _select_initial_step scratch (misc.py:84-143), fp32
  int parity;                        // which of Y[2]/KF[2] holds y0/f0
  int done, status, forced;
  int next_out, n_out;               // next requested time index to emit
  int emit_lo, emit_hi, emit_parity; // outputs [emit_lo, emit_hi) fall in the pending step
  int terminal_only;
};

struct PtrPair {  // buffer chosen by Ctrl::parity (both equal for parity-free buffers)
  float* p[2];
};

enum EpiMode : int {
  EPI_STORE = 0,    // k_out = k
  EPI_LINCOMB = 1,  // y_out = y0 + sum_j (dt*beta_j) k_j, fresh k last   (rk_common.py:50)
  EPI_ERR = 2,      // dopri5 last stage: error estimate + squared-ratio partial sums
  EPI_RK4_1 = 3,    // y + dt*k1/3                                (rk_common.py:75)
  EPI_RK4_2 = 4,    // y + dt*(k1/-3 + k2)                        (rk_common.py:76)
  EPI_RK4_3 = 5,    // y + dt*(k1 - k2 + k3)                      (rk_common.py:77)
  EPI_RK4_4 = 6,    // y + (k1 + 3k2 + 3k3 + k4)*(dt/8)           (rk_common.py:78, solvers.py:91)
  EPI_MASK = 7,     // backward of the ReLU: y_out = k > 0 ? (dt*beta_0) * aux : 0, aux read through y0
  // EPI_LINCOMB that also writes e_out = sum_j (dt*ebeta_j) k_j over the same k_j (fresh k last): dopri5's last
  // Runge-Kutta stage holds k1..k6 anyway, so it leaves the left-to-right PREFIX of the error estimate
  // (rk_common.py:60) behind and the error stage reads one stream instead of six (EpiArgs::err_prefix)
  EPI_LINCOMB_E = 8,
};

enum DtSrc : int { DT_HOST = 0, DT_CTRL = 1, DT_CTRL_H0 = 2 };

// Feature-sharded peer push (ndcn_solver_set_feature_peers): device-resident pointer table, one per solver.
// Rank q gathers ALL rows on its column slice [q Hc, (q+1) Hc): producers of a gather source scatter every
// new row into the slice buffers of all ranks (FEAT_Y_SLICES), the slice gather stores z = Phi x straight
// into the blocked Z of the rank that owns the row (FEAT_Z_OWNERS).
struct FeatTable {
  float* xcs[8];  // rank q's slice buffer [N, Hc]           (IPC-mapped)
  float* z[8];    // rank q's Z [world][n_local_q][Hc]        (IPC-mapped)
  int bounds[9];  // row-block boundaries of the ranks (entries past `world` repeat the last one)
  int world;
  int nl_uniform; // > 0: every block but the last has exactly this many rows (owner = row / nl_uniform)
  int n_total;    // all rows
  int slab;       // 1: slice buffers are column-blocked [Hc/16][N][16] (one 16-column slab = N*64 B contiguous, L2-sized
                  //    at N = 1M: k_gather_slab walks it slab by slab); 0: row-major [N][Hc]
};
enum FeatMode : int { FEAT_OFF = 0, FEAT_Y_SLICES = 1, FEAT_Z_OWNERS = 2 };

struct EpiArgs {
  int mode;
  int n_prev;         // previous stages read from HBM (<= 6)
  int dt_src;         // DtSrc
  int check_finite;   // also flag non-finite y0 (dopri5.py:101-102)
  PtrPair k_out;      // may be {0,0}
  PtrPair y_out;
  PtrPair y0;
  PtrPair y1;         // EPI_ERR only
  PtrPair kprev[6];
  float beta[8];      // fp32(beta_j); coefficient = fp32(dt) * beta_j   (misc.py:25)
  Ctrl* ctrl;         // may be null (fixed-grid solvers, stand-alone ops)
  float dt_host;
  float rtol, atol;
  double* partials;   // [gridDim.x] per-CTA partial sums (EPI_ERR)
  // multi-GPU peer push (ndcn_solver_set_peers): every y_out element is also stored into the gather-source
  // buffers of the other ranks, peer_delta[j] BYTES away from the local address (NVLink peer stores)
  int n_peers;
  long long peer_delta[kMaxPeers];
  // feature-sharded peer push
  const FeatTable* feat;
  int feat_mode;        // FeatMode
  int feat_rank;
  int feat_row0;        // first global row of this rank's block
  int feat_hc_log2, feat_h_log2;
  // EPI_LINCOMB_E: second coefficient set and its output; EPI_ERR with err_prefix: kprev[0] already holds
  // sum_{j<6} (dt*c_err_j) k_j and is added as it is
  float* e_out;
  float ebeta[8];
  int err_prefix;
};

struct EpiCtx {  // EpiArgs resolved against the controller, per thread
  int mode, n_prev;
  float* k_out;
  float* y_out;
  const float* y0;
  const float* y1;
  const float* kprev[6];
  float coef[8];
  float coef_fresh;  // coefficient of the k still in registers (= coef[n_prev])
  float dt;
  float rtol, atol;
  int n_peers;
  long long peer_delta[kMaxPeers];
  const FeatTable* feat;
  int feat_mode, feat_rank, feat_row0, feat_hc_log2, feat_h_log2;
  float* e_out;
  float ecoef[8];
  float ecoef_fresh;
  int err_prefix;
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// kernel parameters live in constant memory: select, do not index dynamically (that would
// force a local-memory copy of the whole argument struct)
__device__ __forceinline__ float* sel(const PtrPair& q, int par) { return par ? q.p[1] : q.p[0]; }

// returns false when the solve is already finished (kernel should exit)
__device__ __forceinline__ bool epi_resolve(const EpiArgs& a, EpiCtx& c) {
  int par = 0;
  float dt = a.dt_host;
  if (a.ctrl != nullptr) {
    const volatile Ctrl* ct = a.ctrl;
    if (ct->done) return false;
    par = ct->parity;
    if (a.dt_src == DT_CTRL) dt = (float)ct->dt;  // dt.type(fp32)   rk_common.py:46
    else if (a.dt_src == DT_CTRL_H0) dt = ct->h0;
  }
  c.mode = a.mode;
  c.n_prev = a.n_prev;
  // the RK4 (3/8) epilogues read a fixed number of earlier stages (rk_common.py:75-78)
  if (a.mode == EPI_RK4_1) c.n_prev = 0;
  else if (a.mode == EPI_RK4_2) c.n_prev = 1;
  else if (a.mode == EPI_RK4_3) c.n_prev = 2;
  else if (a.mode == EPI_RK4_4) c.n_prev = 3;
  else if (a.mode == EPI_MASK) c.n_prev = 0;
  c.k_out = sel(a.k_out, par);
  c.y_out = sel(a.y_out, par);
  c.y0 = sel(a.y0, par);
  c.y1 = sel(a.y1, par);
#pragma unroll
  for (int j = 0; j < 6; ++j) c.kprev[j] = sel(a.kprev[j], par);
#pragma unroll
  for (int j = 0; j < 8; ++j) c.coef[j] = fmul(dt, a.beta[j]);
  c.coef_fresh = fmul(dt, a.beta[a.n_prev & 7]);
  c.dt = dt;
  c.rtol = a.rtol;
  c.atol = a.atol;
  c.n_peers = a.n_peers;
#pragma unroll
  for (int j = 0; j < kMaxPeers; ++j) c.peer_delta[j] = a.peer_delta[j];
  c.feat = a.feat;
  c.feat_mode = a.feat_mode;
  c.feat_rank = a.feat_rank;
  c.feat_row0 = a.feat_row0;
  c.feat_hc_log2 = a.feat_hc_log2;
  c.feat_h_log2 = a.feat_h_log2;
  c.e_out = a.e_out;
#pragma unroll
  for (int j = 0; j < 8; ++j) c.ecoef[j] = fmul(dt, a.ebeta[j]);
  c.ecoef_fresh = fmul(dt, a.ebeta[a.n_prev & 7]);
  c.err_prefix = a.err_prefix;
  return true;
}

// ---- vector load/store helpers (VW in {1,2,4}) ---------------------------------------
template <int VW>
__device__ __forceinline__ void ldv(const float* __restrict__ p, float (&v)[VW]) {
  if constexpr (VW == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VW == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = *p;
  }
}
template <int VW>
__device__ __forceinline__ void ldv_stream(const float* __restrict__ p, float (&v)[VW]) {
  if constexpr (VW == 4) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VW == 2) {
    float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = __ldcs(p);
  }
}
template <int VW>
__device__ __forceinline__ void stv_stream(float* __restrict__ p, const float (&v)[VW]) {
  if constexpr (VW == 4) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  } else if constexpr (VW == 2) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  } else {
    __stcs(p, v[0]);
  }
}
template <int VW>
__device__ __forceinline__ void stv(float* __restrict__ p, const float (&v)[VW]) {
  if constexpr (VW == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (VW == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
    *p = v[0];
  }
}

// feature-sharded peer push: off = row * H + col of this rank's [n_local, H] block; the VW <= 4 elements lie
// in one column slice (Hc >= 32).  Out of line on purpose: the single-GPU epilogues keep their registers.
template <int VW>
__device__ __noinline__ void store_y_slices(const FeatTable* __restrict__ f, int row0, int hc_log2, int h_log2,
                                            int64_t off, float v0, float v1, float v2, float v3) {
  const int64_t row = off >> h_log2;
  const int col = (int)(off & (((int64_t)1 << h_log2) - 1));
  const int cs = col & ((1 << hc_log2) - 1);  // column inside the owner's slice
  float* dst = f->xcs[col >> hc_log2];
  if (f->slab) dst += (((int64_t)(cs >> 4) * f->n_total + row0 + row) << 4) + (cs & 15);
  else dst += (((int64_t)row0 + row) << hc_log2) + cs;
  if constexpr (VW == 4) *reinterpret_cast<float4*>(dst) = make_float4(v0, v1, v2, v3);
  else if constexpr (VW == 2) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
  else *dst = v0;
}

// y_out store: the local copy and, on a peer-push multi-GPU solve, the same elements in every other
// rank's gather-source buffer (plain stores over NVLink; ordered by the k_peer_barrier that follows)
template <int VW>
__device__ __forceinline__ void store_y(const EpiCtx& c, int64_t off, const float (&v)[VW], bool stream_out) {
  if (stream_out) stv_stream<VW>(c.y_out + off, v);
  else stv<VW>(c.y_out + off, v);
  if (c.n_peers > 0) {
    char* base = reinterpret_cast<char*>(c.y_out + off);
#pragma unroll
    for (int j = 0; j < kMaxPeers; ++j)
      if (j < c.n_peers) stv<VW>(reinterpret_cast<float*>(base + c.peer_delta[j]), v);
  }
  if (c.feat_mode == FEAT_Y_SLICES)
    store_y_slices<VW>(c.feat, c.feat_row0, c.feat_hc_log2, c.feat_h_log2, off, v[0], v[VW > 1 ? 1 : 0], v[VW > 2 ? 2 : 0],
                       v[VW > 3 ? 3 : 0]);
}

// slice gather (FEAT_Z_OWNERS): off = row * Hc + col with row a GLOBAL node id; the value goes to block
// `rank` of the Z of the rank that owns the row: z_o[rank][row - row0_o][col]
template <int VW>
__device__ __noinline__ void store_z_owner(const FeatTable* __restrict__ f, int rank, int hc_log2, int64_t off,
                                           float v0, float v1, float v2, float v3) {
  const int64_t row = off >> hc_log2;
  const int col = (int)(off & (((int64_t)1 << hc_log2) - 1));
  int o, lo, hi;
  const int nl = f->nl_uniform;
  if (nl > 0) {  // uniform blocks: no table walk, one dependent load (the owner's Z pointer) per store
    o = (int)((uint32_t)row / (uint32_t)nl);
    lo = o * nl;
    hi = min(lo + nl, f->n_total);
  } else {
    o = 0;
#pragma unroll
    for (int j = 1; j < 8; ++j) o += (row >= f->bounds[j] && j < f->world) ? 1 : 0;  // bounds ascend: owner = # of cuts <= row
    lo = f->bounds[o];
    hi = f->bounds[o + 1];
  }
  float* dst = f->z[o] + (((int64_t)rank * (hi - lo) + (row - lo)) << hc_log2) + col;
  if constexpr (VW == 4) *reinterpret_cast<float4*>(dst) = make_float4(v0, v1, v2, v3);
  else if constexpr (VW == 2) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
  else *dst = v0;
}

// ------------------------------------------------------------------------------------
// The stage epilogue: consumes freshly computed k values (still in registers) for VW
// consecutive elements starting at element offset `off`.  Split in two phases so that a
// caller can issue the loads of several element groups before doing any arithmetic
// (memory-level parallelism); epi_apply() is load + math for one group.
// ------------------------------------------------------------------------------------
template <int VW>
struct EpiIn {
  float y0[VW];
  float y1[VW];
  float kp[6][VW];
};

template <int VW>
__device__ __forceinline__ void epi_load(const EpiCtx& c, int64_t off, EpiIn<VW>& in) {
  if (c.mode == EPI_STORE) return;
  ldv_stream<VW>(c.y0 + off, in.y0);
  if (c.mode == EPI_ERR) ldv<VW>(c.y1 + off, in.y1);
#pragma unroll
  for (int j = 0; j < 6; ++j)
    if (j < c.n_prev) ldv_stream<VW>(c.kprev[j] + off, in.kp[j]);
}

// stream_out: k_out / y_out are written with streaming (evict-first) stores -- for states far larger
// than L2, where a written line is gone long before the next kernel reads it
template <int VW>
__device__ __forceinline__ void epi_math(const EpiCtx& c, int64_t off, const float (&k)[VW], const EpiIn<VW>& in,
                                         double& err_acc, bool stream_out = false) {
  if (c.k_out != nullptr) {
    if (stream_out) stv_stream<VW>(c.k_out + off, k);
    else stv<VW>(c.k_out + off, k);
  }
  if (c.feat_mode == FEAT_Z_OWNERS)
    store_z_owner<VW>(c.feat, c.feat_rank, c.feat_hc_log2, off, k[0], k[VW > 1 ? 1 : 0], k[VW > 2 ? 2 : 0], k[VW > 3 ? 3 : 0]);
  if (c.mode == EPI_STORE) return;

  if (c.mode == EPI_LINCOMB || c.mode == EPI_LINCOMB_E) {
    if (c.mode == EPI_LINCOMB_E) {
      // prefix of the error estimate over the same stages, same summation order as EPI_ERR below
      float ea[VW];
      if (c.n_prev == 0) {
#pragma unroll
        for (int i = 0; i < VW; ++i) ea[i] = fmul(c.ecoef[0], k[i]);
      } else {
#pragma unroll
        for (int i = 0; i < VW; ++i) ea[i] = fmul(c.ecoef[0], in.kp[0][i]);
#pragma unroll
        for (int j = 1; j < 6; ++j) {
          if (j < c.n_prev) {
#pragma unroll
            for (int i = 0; i < VW; ++i) ea[i] = fadd(ea[i], fmul(c.ecoef[j], in.kp[j][i]));
          }
        }
        const float ef = c.ecoef_fresh;
#pragma unroll
        for (int i = 0; i < VW; ++i) ea[i] = fadd(ea[i], fmul(ef, k[i]));
      }
      if (stream_out) stv_stream<VW>(c.e_out + off, ea);
      else stv<VW>(c.e_out + off, ea);
    }
    // y_out = y0 + sum_j (dt*beta_j) k_j, summed left to right from the first term
    // (misc.py:22-25: sum() of the per-term products; rk_common.py:50)
    float acc[VW];
    if (c.n_prev == 0) {
#pragma unroll
      for (int i = 0; i < VW; ++i) acc[i] = fmul(c.coef[0], k[i]);
    } else {
#pragma unroll
      for (int i = 0; i < VW; ++i) acc[i] = fmul(c.coef[0], in.kp[0][i]);
#pragma unroll
      for (int j = 1; j < 6; ++j) {
        if (j < c.n_prev) {
#pragma unroll
          for (int i = 0; i < VW; ++i) acc[i] = fadd(acc[i], fmul(c.coef[j], in.kp[j][i]));
        }
      }
      const float cf = c.coef_fresh;
#pragma unroll
      for (int i = 0; i < VW; ++i) acc[i] = fadd(acc[i], fmul(cf, k[i]));
    }
    float out[VW];
#pragma unroll
    for (int i = 0; i < VW; ++i) out[i] = fadd(in.y0[i], acc[i]);
    store_y<VW>(c, off, out, stream_out);
    return;
  }

  if (c.mode == EPI_ERR) {
    // err = sum_j (dt*c_err_j) k_j (n_prev earlier stages, then the fresh one)      rk_common.py:60
    // ratio = err / (atol + rtol*max(|y0|,|y1|)); sum ratio^2      misc.py:146-157
    float acc[VW];
#pragma unroll
    for (int i = 0; i < VW; ++i) acc[i] = c.err_prefix ? in.kp[0][i] : fmul(c.coef[0], in.kp[0][i]);
#pragma unroll
    for (int j = 1; j < 6; ++j) {
      if (j < c.n_prev) {
#pragma unroll
        for (int i = 0; i < VW; ++i) acc[i] = fadd(acc[i], fmul(c.coef[j], in.kp[j][i]));
      }
    }
    {
      const float cf = c.coef_fresh;
#pragma unroll
      for (int i = 0; i < VW; ++i) acc[i] = fadd(acc[i], fmul(cf, k[i]));
    }
    float group = 0.f;  // the VW squared ratios are summed in fp32 first: one fp64 add per group
#pragma unroll
    for (int i = 0; i < VW; ++i) {
      float tol = fadd(c.atol, fmul(c.rtol, fmaxf(fabsf(in.y0[i]), fabsf(in.y1[i]))));
      // 0 / tol is 0 either way; behind the ReLU the error estimate is exactly 0 for every element whose
      // stages are all clipped, and a zero numerator sends IEEE division down its slow path (measured:
      // ~15 % of the error-stage kernel's samples)
      float r = 0.f;
      // (an approximate division here was measured: 1.32 -> 1.25 ms for the error stage, not worth giving up the
      // reference's IEEE quotient)
      if (acc[i] != 0.f) r = fdiv(acc[i], tol);
      float r2 = fmul(r, r);
      // torch.max propagates NaN, fmaxf does not: keep the poison visible to the controller
      if (!(r2 == r2) || in.y0[i] != in.y0[i] || in.y1[i] != in.y1[i]) r2 = __int_as_float(0x7fc00000);
      group += r2;
    }
    err_acc += (double)group;
    return;
  }

  if (c.mode == EPI_MASK) {
    // vjp of relu(pre): the cotangent passes where the forward value is positive (k = relu(pre) > 0 <=> pre > 0)
    float out[VW];
#pragma unroll
    for (int i = 0; i < VW; ++i) out[i] = k[i] > 0.f ? fmul(c.coef[0], in.y0[i]) : 0.f;
    if (stream_out) stv_stream<VW>(c.y_out + off, out);
    else stv<VW>(c.y_out + off, out);
    return;
  }

  // ---- RK4 (3/8 rule), literal operation order of rk_common.py:72-78 ----
  float out[VW];
  const float dt = c.dt;
  if (c.mode == EPI_RK4_1) {
#pragma unroll
    for (int i = 0; i < VW; ++i) out[i] = fadd(in.y0[i], fdiv(fmul(dt, k[i]), 3.0f));
  } else if (c.mode == EPI_RK4_2) {
#pragma unroll
    for (int i = 0; i < VW; ++i) out[i] = fadd(in.y0[i], fmul(dt, fadd(fdiv(in.kp[0][i], -3.0f), k[i])));
  } else if (c.mode == EPI_RK4_3) {
#pragma unroll
    for (int i = 0; i < VW; ++i) out[i] = fadd(in.y0[i], fmul(dt, fadd(fsub(in.kp[0][i], in.kp[1][i]), k[i])));
  } else {  // EPI_RK4_4
    const float dt8 = fdiv(dt, 8.0f);
#pragma unroll
    for (int i = 0; i < VW; ++i) {
      float s = fadd(fadd(fadd(in.kp[0][i], fmul(3.0f, in.kp[1][i])), fmul(3.0f, in.kp[2][i])), k[i]);
      out[i] = fadd(in.y0[i], fmul(s, dt8));
    }
  }
  store_y<VW>(c, off, out, stream_out);
}

template <int VW>
__device__ __forceinline__ void epi_apply(const EpiCtx& c, int64_t off, const float (&k)[VW], double& err_acc) {
  EpiIn<VW> in;
  epi_load<VW>(c, off, in);
  epi_math<VW>(c, off, k, in, err_acc);
}

// per-CTA reduction of the error partials; every CTA writes its slot (zeros included)
__device__ __forceinline__ void epi_finish_block(const EpiArgs& a, double err_acc) {
  if (a.mode != EPI_ERR) return;
  __shared__ double s_red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) err_acc += __shfl_xor_sync(0xffffffffu, err_acc, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarp = (blockDim.x + 31) >> 5;
  if (lane == 0) s_red[warp] = err_acc;
  __syncthreads();
  if (warp == 0) {
    double v = lane < nwarp ? s_red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) a.partials[blockIdx.x] = v;
  }
}

struct GraphView {
  const int32_t* rowptr;
  const int32_t* col;
  const float* val;
  int64_t n_rows, n_cols, nnz;
};

// ---- mbarrier / bulk-copy (TMA, 1-D) PTX wrappers --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// cp.async.bulk (SASS: UBLKCP): contiguous global -> shared, completion on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace ndcn

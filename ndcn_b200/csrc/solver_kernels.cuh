// Device-side step control for dopri5 and the dense-output / initial-step kernels.
// Reference: torchdiffeq/_impl/dopri5.py:58-122, misc.py:84-170, interp.py:5-65.
#pragma once
#include "ndcn_common.cuh"

namespace ndcn {

// deterministic fixed-order sum of the per-CTA partials (one block)
__device__ __forceinline__ double block_sum_partials(const double* __restrict__ partials, int n) {
  __shared__ double s_tmp[kStageThreadsCtl];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
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

// ---------------------------------------------------------------------------------------
// Multi-GPU peer push (one process per GPU, buffers mapped into each other with CUDA IPC, NVLink):
// the stage kernels store every new gather-source row straight into the other ranks' buffers
// (store_y, ndcn_common.cuh); this one-block kernel is the barrier that separates those stores from
// the gathers that read them, and -- with a payload -- the all-reduce (SUM, fixed rank order, so every
// rank gets the same bits) of the two doubles the step controller needs.  Each rank owns one PeerPad
// in IPC-shared memory: rank r announces epoch e by a system-scope release store into flag[r] of every
// OTHER rank's pad and spins (acquire) on its own pad, i.e. on local memory.  A rank can be at most one
// epoch ahead of the slowest one, so two payload slots (epoch parity) are enough.
// ---------------------------------------------------------------------------------------
struct PeerPad {
  unsigned long long flag[8];   // flag[r]: last epoch rank r announced (written remotely by rank r)
  double pay[2][8][2];          // pay[epoch & 1][r]: rank r's addends
  unsigned long long epoch;     // barriers this rank has executed (owner only)
  int timed_out;                // owner only: a peer never arrived
  int pad_;
};
static_assert(sizeof(PeerPad) <= 4096, "PeerPad must fit the 4 KB head of the shared allocation");

struct PeerArgs {
  PeerPad* self;
  PeerPad* peer[8];  // [world], entry `rank` unused
  int rank, world;
};

constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(32) k_peer_barrier(PeerArgs a, double* payload, Ctrl* ctrl) {
  __shared__ unsigned long long s_epoch;
  const int j = threadIdx.x;
  PeerPad* self = a.self;
  if (j == 0) {
    s_epoch = self->epoch + 1;
    self->epoch = s_epoch;
  }
  __syncwarp();
  const unsigned long long ep = s_epoch;
  const bool active = j < a.world && j != a.rank;
  const int slot = (int)(ep & 1ull);
  if (active) {
    PeerPad* p = a.peer[j];
    if (payload != nullptr) {
      volatile double* dst = &p->pay[slot][a.rank][0];
      dst[0] = payload[0];
      dst[1] = payload[1];
    }
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&p->flag[a.rank]), "l"(ep) : "memory");
    // wait for rank j on local memory
    if (!*(volatile int*)&self->timed_out) {
      const unsigned long long t0 = global_timer_ns();
      unsigned long long seen = 0;
      unsigned int spins = 0;
      for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(&self->flag[j]) : "memory");
        if (seen >= ep) break;
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > kPeerTimeoutNs) {
          *(volatile int*)&self->timed_out = 1;
          if (ctrl != nullptr) {
            ((volatile Ctrl*)ctrl)->status = NDCN_E_PEER_TIMEOUT;
            ((volatile Ctrl*)ctrl)->done = 1;
          }
          break;
        }
      }
    }
  }
  __syncwarp();
  if (payload != nullptr && j == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int r = 0; r < a.world; ++r) {
      if (r == a.rank) {
        s0 += payload[0];
        s1 += payload[1];
      } else {
        const volatile double* src = &self->pay[slot][r][0];
        s0 += src[0];
        s1 += src[1];
      }
    }
    payload[0] = s0;
    payload[1] = s1;
  }
}

// initial state of a peer-push solve: dst (own rows of the gather source) and the same rows in every
// peer's buffer, through the same store path as every later push
struct PushDeltas {
  int n;
  long long delta[kMaxPeers];  // bytes
};
__global__ void __launch_bounds__(kStageThreads) k_copy_push(const float* __restrict__ src, float* __restrict__ dst,
                                                             int64_t numel, PushDeltas pd, int vec) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = vec ? (numel >> 2) : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<float4*>(dst)[i] = v;
#pragma unroll
    for (int j = 0; j < kMaxPeers; ++j)
      if (j < pd.n) *reinterpret_cast<float4*>(reinterpret_cast<char*>(dst + i * 4) + pd.delta[j]) = v;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    const float v = src[i];
    dst[i] = v;
#pragma unroll
    for (int j = 0; j < kMaxPeers; ++j)
      if (j < pd.n) *reinterpret_cast<float*>(reinterpret_cast<char*>(dst + i) + pd.delta[j]) = v;
  }
}

// initial state of a feature-sharded peer-push solve: dst (own [n_local, H] block, row-major) and, column
// slice by column slice, the slice buffers of all ranks.  16-byte pieces (H % 4 == 0, all bases aligned).
__global__ void __launch_bounds__(kStageThreads) k_copy_slices(const float* __restrict__ src, float* __restrict__ dst,
                                                               int64_t numel, const FeatTable* __restrict__ feat, int row0,
                                                               int h_log2, int hc_log2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = numel >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<float4*>(dst)[i] = v;
    const int64_t off = i * 4;
    const int64_t row = off >> h_log2;
    const int col = (int)(off & (((int64_t)1 << h_log2) - 1));
    const int cs = col & ((1 << hc_log2) - 1);
    float* p = feat->xcs[col >> hc_log2];
    if (feat->slab) p += (((int64_t)(cs >> 4) * feat->n_total + row0 + row) << 4) + (cs & 15);
    else p += (((int64_t)row0 + row) << hc_log2) + cs;
    *reinterpret_cast<float4*>(p) = v;
  }
}

// ---------------------------------------------------------------------------------------
// Accept / reject + next step size.  One block; thread 0 does the float64 scalar work.
// `reduce_stage`: 0 = sum partials and (single GPU) decide immediately;
//                 1 = only sum partials into xchg[0..1] (multi-GPU: the host hook all-reduces);
//                 2 = decide from xchg[0] (after the all-reduce).
// ---------------------------------------------------------------------------------------
// the float64 scalar work of one attempt (one thread), given the sum of squared error ratios
__device__ __forceinline__ void controller_decide(Ctrl& c, double sum, const double* __restrict__ t_out) {
  c.sum_sq = sum;
  c.n_attempt += 1;
  // torch.mean of the squared ratios, an fp32 scalar   (misc.py:155-156)
  const float msr = (float)(sum / c.numel_global);
  c.msr_last = msr;
  const bool accept = c.forced ? true : (msr <= 1.0f);  // dopri5.py:109
  const double dt = c.dt;
  const double t_start = c.t1;
  c.emit_lo = c.emit_hi = c.next_out;
  if (accept) {
    const double t_new = t_start + dt;  // dopri5.py:116
    c.t0 = t_start;
    c.t1 = t_new;
    c.emit_t0 = t_start;
    c.emit_t1 = t_new;
    c.emit_dt = dt;
    c.emit_parity = c.parity;
    c.parity ^= 1;  // y1 -> y0, k7 -> k1 (FSAL)
    c.n_accept += 1;
    // every requested time already covered by this step is emitted now (dopri5.py:85-92:
    // the while loop is not entered again for them)
    int hi = c.next_out;
    while (hi < c.n_out && !(t_out[hi] > t_new)) ++hi;
    if (hi > c.next_out) c.steps_this_interval = 0;
    else c.steps_this_interval += 1;
    c.emit_hi = hi;
    c.next_out = hi;
  } else {
    c.t0 = t_start;  // rejected: (t0, t1) collapse onto the step start (dopri5.py:120)
    c.n_reject += 1;
    c.steps_this_interval += 1;
  }
  if (!c.forced) {
    // _optimal_step_size, misc.py:160-170
    double dt_next;
    if (msr == 0.0f) {
      dt_next = dt * c.ifactor;
    } else {
      const double dfac = (msr < 1.0f) ? 1.0 : c.dfactor;
      const double root = (double)sqrtf(msr);
      const double expo = (double)0.2f;  // torch.tensor(1/order) is fp32 before .to(float64)
      double factor = fmin(pow(root, expo) / c.safety, 1.0 / dfac);
      factor = fmax(1.0 / c.ifactor, factor);
      if (msr != msr) factor = (double)msr;  // torch.min/max propagate NaN
      dt_next = dt / factor;
    }
    c.dt = dt_next;
  }
  if (c.next_out >= c.n_out) {
    c.done = 1;
  } else {
    if (!(c.t1 + c.dt > c.t1)) {  // dopri5.py:100
      c.status = NDCN_E_DT_UNDERFLOW;
      c.done = 1;
    } else if (c.steps_this_interval >= c.max_num_steps) {  // dopri5.py:89
      c.status = NDCN_E_MAX_STEPS;
      c.done = 1;
    }
  }
}

__global__ void __launch_bounds__(kStageThreadsCtl) k_controller(Ctrl* ctrl, const double* partials, int n_partials,
                                                                 const double* t_out, double* xchg, int reduce_stage) {
  if (((volatile Ctrl*)ctrl)->done) {
    // attempts enqueued past the end are no-ops: also retire the already-emitted outputs
    if (threadIdx.x == 0) ctrl->emit_lo = ctrl->emit_hi;
    return;
  }
  if (((volatile Ctrl*)ctrl)->status != 0) {  // e.g. non-finite state flagged by the pre-stage
    if (threadIdx.x == 0) { ctrl->done = 1; ctrl->emit_lo = ctrl->emit_hi; }
    return;
  }
  double sum = 0.0;
  if (reduce_stage != 2) sum = block_sum_partials(partials, n_partials);
  if (threadIdx.x != 0) return;
  if (reduce_stage == 1) {
    xchg[0] = sum;
    xchg[1] = 0.0;
    return;
  }
  if (reduce_stage == 2) sum = xchg[0];
  controller_decide(*ctrl, sum, t_out);
}

// ---------------------------------------------------------------------------------------
// Dense output of the accepted step for every requested time inside it.
// _interp_fit_dopri5 (dopri5.py:39-45) + _interp_fit (interp.py:5-35) +
// _interp_evaluate (interp.py:38-65), fused per element; nothing but the outputs is stored.
// ---------------------------------------------------------------------------------------
struct EmitArgs {
  Ctrl* ctrl;
  const double* t_out;
  PtrPair y0, y1, k0, k6;  // parity-selected by Ctrl::emit_parity
  const float* k[5];       // k1..k5
  float c_mid[7];          // fp32(DPS_C_MID)
  float* out;              // [n_out, numel] or [numel] (terminal only); with a decoder [n_out, n_rows, C] / [n_rows, C]
  int64_t numel;
  // optional fused decoder (NDCN.output_layer, neural_dynamics.py:148,159): out = y W_d^T + b_d, C <= kDecMaxC
  const float* dec_W;      // [C, H] row-major (nn.Linear weight) or null
  const float* dec_b;      // [C] or null
  int dec_C, H;
  int64_t n_rows;
};

constexpr int kDecMaxC = 8;

template <int VW>
__device__ __forceinline__ void emit_elem(const EmitArgs& a, const float* y0p, const float* y1p, const float* k0p,
                                          const float* k6p, int64_t off, float dt, int lo, int hi, const float* xs_all,
                                          int terminal_only, int n_out) {
  float y0[VW], y1[VW], kk[7][VW];
  ldv<VW>(y0p + off, y0);
  ldv<VW>(y1p + off, y1);
  ldv<VW>(k0p + off, kk[0]);
#pragma unroll
  for (int j = 0; j < 5; ++j) ldv<VW>(a.k[j] + off, kk[j + 1]);
  ldv<VW>(k6p + off, kk[6]);
  float ca[VW], cb[VW], cc[VW], cd[VW];
  const float m2dt = fmul(-2.f, dt), p2dt = fmul(2.f, dt), p5dt = fmul(5.f, dt), m3dt = fmul(-3.f, dt),
              m4dt = fmul(-4.f, dt);
#pragma unroll
  for (int i = 0; i < VW; ++i) {
    float acc = fmul(fmul(dt, a.c_mid[0]), kk[0][i]);
#pragma unroll
    for (int j = 1; j < 7; ++j) acc = fadd(acc, fmul(fmul(dt, a.c_mid[j]), kk[j][i]));
    const float ymid = fadd(y0[i], acc);
    const float f0 = kk[0][i], f1 = kk[6][i];
    // a = -2dt f0 + 2dt f1 - 8 y0 - 8 y1 + 16 ymid
    ca[i] = fadd(fadd(fadd(fadd(fmul(m2dt, f0), fmul(p2dt, f1)), fmul(-8.f, y0[i])), fmul(-8.f, y1[i])), fmul(16.f, ymid));
    // b = 5dt f0 - 3dt f1 + 18 y0 + 14 y1 - 32 ymid
    cb[i] = fadd(fadd(fadd(fadd(fmul(p5dt, f0), fmul(m3dt, f1)), fmul(18.f, y0[i])), fmul(14.f, y1[i])), fmul(-32.f, ymid));
    // c = -4dt f0 + dt f1 - 11 y0 - 5 y1 + 16 ymid
    cc[i] = fadd(fadd(fadd(fadd(fmul(m4dt, f0), fmul(dt, f1)), fmul(-11.f, y0[i])), fmul(-5.f, y1[i])), fmul(16.f, ymid));
    cd[i] = fmul(dt, f0);
  }
  for (int j = lo; j < hi; ++j) {
    if (terminal_only && j != n_out - 1) continue;
    const float* xs = xs_all + (j - lo) * 4;  // x, x^2, x^3, x^4
    float o[VW];
#pragma unroll
    for (int i = 0; i < VW; ++i)
      o[i] = fadd(fadd(fadd(fadd(fmul(ca[i], xs[3]), fmul(cb[i], xs[2])), fmul(cc[i], xs[1])), fmul(cd[i], xs[0])),
                  fmul(y0[i], 1.0f));
    float* dst = terminal_only ? a.out : a.out + (int64_t)j * a.numel;
    stv<VW>(dst + off, o);
  }
}

constexpr int kEmitMaxPerLaunch = 32;

// dense output of the pending step for the elements [tid, numel) in steps of `stride` threads; every thread of the
// CTA must call it (block barriers around the shared abscissa table `xs`, kEmitMaxPerLaunch * 4 floats)
__device__ __forceinline__ void emit_range(const EmitArgs& a, int vec, int64_t tid, int64_t stride, float* xs) {
  const volatile Ctrl* ct = a.ctrl;
  const int lo0 = ct->emit_lo, hi0 = ct->emit_hi;
  if (lo0 >= hi0) return;
  const int par = ct->emit_parity;
  const int terminal_only = ct->terminal_only, n_out = ct->n_out;
  if (terminal_only && hi0 != n_out) return;
  const float t0 = (float)ct->emit_t0, t1 = (float)ct->emit_t1;  // interp.py:54-56
  const float dt = (float)ct->emit_dt;                            // dopri5.py:41
  const float* y0p = sel(a.y0, par);
  const float* y1p = sel(a.y1, par);
  const float* k0p = sel(a.k0, par);
  const float* k6p = sel(a.k6, par);
  for (int lo = lo0; lo < hi0; lo += kEmitMaxPerLaunch) {
    const int hi = min(hi0, lo + kEmitMaxPerLaunch);
    __syncthreads();
    if ((int)threadIdx.x < hi - lo) {
      const float t = (float)a.t_out[lo + threadIdx.x];
      const float x = fdiv(fsub(t, t0), fsub(t1, t0));
      float p = x;
      xs[threadIdx.x * 4 + 0] = p;
      p = fmul(p, x); xs[threadIdx.x * 4 + 1] = p;
      p = fmul(p, x); xs[threadIdx.x * 4 + 2] = p;
      p = fmul(p, x); xs[threadIdx.x * 4 + 3] = p;
    }
    __syncthreads();
    if (vec) {
      const int64_t n4 = a.numel >> 2;
      for (int64_t i = tid; i < n4; i += stride)
        emit_elem<4>(a, y0p, y1p, k0p, k6p, i * 4, dt, lo, hi, xs, terminal_only, n_out);
      for (int64_t i = (n4 << 2) + tid; i < a.numel; i += stride)
        emit_elem<1>(a, y0p, y1p, k0p, k6p, i, dt, lo, hi, xs, terminal_only, n_out);
    } else {
      for (int64_t i = tid; i < a.numel; i += stride)
        emit_elem<1>(a, y0p, y1p, k0p, k6p, i, dt, lo, hi, xs, terminal_only, n_out);
    }
  }
}

__global__ void __launch_bounds__(kStageThreads) k_emit(EmitArgs a, int vec) {
  __shared__ float xs[kEmitMaxPerLaunch * 4];
  emit_range(a, vec, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, xs);
}

// ---------------------------------------------------------------------------------------
// Decoder-fused emission: NDCN applies output_layer = Linear(H -> C)
// to every one of the T returned states (neural_dynamics.py:159); at N=1M, H=256, T=100 the
// [T, N, H] slab would be 102 GB.  Here the decoder runs where the state is produced and only
// [T, N, C] is ever written.  One warp per row, lanes stride the columns, shuffle reduction.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void decode_store(const float (&acc)[kDecMaxC], int C, const float* dec_b, float* dst, int lane) {
#pragma unroll
  for (int c = 0; c < kDecMaxC; ++c) {
    if (c < C) {
      float v = acc[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) dst[c] = v + (dec_b ? dec_b[c] : 0.f);
    }
  }
}

// out[r, :] = y[r, :] W_d^T + b_d    (slot 0 = y0, fixed-grid states, terminal states); warps warp0, warp0 + n_warps, ...
__device__ __forceinline__ void decode_rows_range(const float* __restrict__ y, int64_t n_rows, int H,
                                                  const float* __restrict__ dec_W, const float* __restrict__ dec_b, int C,
                                                  float* __restrict__ out, int64_t warp0, int64_t n_warps, int lane) {
  for (int64_t r = warp0; r < n_rows; r += n_warps) {
    float acc[kDecMaxC];
#pragma unroll
    for (int c = 0; c < kDecMaxC; ++c) acc[c] = 0.f;
    for (int col = lane; col < H; col += 32) {
      const float v = y[r * H + col];
#pragma unroll
      for (int c = 0; c < kDecMaxC; ++c)
        if (c < C) acc[c] = fmaf(v, __ldg(dec_W + c * H + col), acc[c]);
    }
    decode_store(acc, C, dec_b, out + r * C, lane);
  }
}

__global__ void __launch_bounds__(kStageThreads) k_decode_rows(const float* __restrict__ y, int64_t n_rows, int H,
                                                               const float* __restrict__ dec_W,
                                                               const float* __restrict__ dec_b, int C,
                                                               float* __restrict__ out) {
  decode_rows_range(y, n_rows, H, dec_W, dec_b, C, out, ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5,
                    ((int64_t)gridDim.x * blockDim.x) >> 5, threadIdx.x & 31);
}

// dense output of the pending dopri5 step, decoded: same fit/evaluate arithmetic as emit_elem<1>
__device__ __forceinline__ void emit_decode_range(const EmitArgs& a, int64_t warp0, int64_t n_warps) {
  const volatile Ctrl* ct = a.ctrl;
  const int lo0 = ct->emit_lo, hi0 = ct->emit_hi;
  if (lo0 >= hi0) return;
  const int par = ct->emit_parity;
  const int terminal_only = ct->terminal_only, n_out = ct->n_out;
  if (terminal_only && hi0 != n_out) return;
  const float t0 = (float)ct->emit_t0, t1 = (float)ct->emit_t1;
  const float dt = (float)ct->emit_dt;
  const float* y0p = sel(a.y0, par);
  const float* y1p = sel(a.y1, par);
  const float* k0p = sel(a.k0, par);
  const float* k6p = sel(a.k6, par);
  const int lane = threadIdx.x & 31, H = a.H, C = a.dec_C;
  const float m2dt = fmul(-2.f, dt), p2dt = fmul(2.f, dt), p5dt = fmul(5.f, dt), m3dt = fmul(-3.f, dt),
              m4dt = fmul(-4.f, dt);
  // outputs in groups of kJB: the nine streams of the step are read once per group, not once per output
  constexpr int kJB = 4;
  for (int j0 = terminal_only ? n_out - 1 : lo0; j0 < hi0; j0 += kJB) {
    const int nj = min(kJB, hi0 - j0);
    float xp[kJB][4];
#pragma unroll
    for (int jj = 0; jj < kJB; ++jj) {
      const float t = (float)a.t_out[min(j0 + jj, hi0 - 1)];
      const float x1 = fdiv(fsub(t, t0), fsub(t1, t0));  // interp.py:59
      xp[jj][0] = x1;
      xp[jj][1] = fmul(x1, x1);
      xp[jj][2] = fmul(xp[jj][1], x1);
      xp[jj][3] = fmul(xp[jj][2], x1);
    }
    for (int64_t r = warp0; r < a.n_rows; r += n_warps) {
      float acc[kJB][kDecMaxC];
#pragma unroll
      for (int jj = 0; jj < kJB; ++jj)
#pragma unroll
        for (int c = 0; c < kDecMaxC; ++c) acc[jj][c] = 0.f;
      for (int col = lane; col < H; col += 32) {
        const int64_t off = r * H + col;
        const float y0 = y0p[off], y1 = y1p[off];
        float kk[7];
        kk[0] = k0p[off];
#pragma unroll
        for (int q = 0; q < 5; ++q) kk[q + 1] = a.k[q][off];
        kk[6] = k6p[off];
        float s = fmul(fmul(dt, a.c_mid[0]), kk[0]);
#pragma unroll
        for (int q = 1; q < 7; ++q) s = fadd(s, fmul(fmul(dt, a.c_mid[q]), kk[q]));
        const float ymid = fadd(y0, s);
        const float f0 = kk[0], f1 = kk[6];
        const float ca = fadd(fadd(fadd(fadd(fmul(m2dt, f0), fmul(p2dt, f1)), fmul(-8.f, y0)), fmul(-8.f, y1)), fmul(16.f, ymid));
        const float cb = fadd(fadd(fadd(fadd(fmul(p5dt, f0), fmul(m3dt, f1)), fmul(18.f, y0)), fmul(14.f, y1)), fmul(-32.f, ymid));
        const float cc = fadd(fadd(fadd(fadd(fmul(m4dt, f0), fmul(dt, f1)), fmul(-11.f, y0)), fmul(-5.f, y1)), fmul(16.f, ymid));
        const float cd = fmul(dt, f0);
        float wd[kDecMaxC];
#pragma unroll
        for (int c = 0; c < kDecMaxC; ++c) wd[c] = c < C ? __ldg(a.dec_W + c * H + col) : 0.f;
#pragma unroll
        for (int jj = 0; jj < kJB; ++jj) {
          if (jj < nj) {
            const float o = fadd(fadd(fadd(fadd(fmul(ca, xp[jj][3]), fmul(cb, xp[jj][2])), fmul(cc, xp[jj][1])),
                                      fmul(cd, xp[jj][0])), fmul(y0, 1.0f));
#pragma unroll
            for (int c = 0; c < kDecMaxC; ++c)
              if (c < C) acc[jj][c] = fmaf(o, wd[c], acc[jj][c]);
          }
        }
      }
#pragma unroll
      for (int jj = 0; jj < kJB; ++jj) {
        if (jj < nj) {
          float* dst_slice = terminal_only ? a.out : a.out + (int64_t)(j0 + jj) * a.n_rows * C;
          decode_store(acc[jj], C, a.dec_b, dst_slice + r * C, lane);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kStageThreads) k_emit_decode(EmitArgs a) {
  emit_decode_range(a, ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, ((int64_t)gridDim.x * blockDim.x) >> 5);
}

// ---------------------------------------------------------------------------------------
// _select_initial_step (misc.py:84-143), fp32, RMS norms, order = 4 for dopri5.
// ---------------------------------------------------------------------------------------
// partials[2*b+0] += sum (u/scale)^2, partials[2*b+1] += sum (v/scale)^2, scale = atol+|y0|*rtol
// mode 0: u = y0, v = f0          (d0, d1)
// mode 1: u = f1 - f0, v unused   (d2)
// every thread of the CTA must call it (block reduction); writes partials[2 * blockIdx.x + {0,1}]
__device__ __forceinline__ void init_norms_block(const float* y0, const float* f0, const float* f1, int64_t numel,
                                                 float rtol, float atol, int mode, double* partials, int64_t tid,
                                                 int64_t stride) {
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = tid; i < numel; i += stride) {
    const float y = y0[i];
    const float scale = fadd(atol, fmul(fabsf(y), rtol));
    if (mode == 0) {
      const float u = fdiv(y, scale), v = fdiv(f0[i], scale);
      s0 += (double)fmul(u, u);
      s1 += (double)fmul(v, v);
    } else {
      const float u = fdiv(fsub(f1[i], f0[i]), scale);
      s0 += (double)fmul(u, u);
    }
  }
  __shared__ double r0[32], r1[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // r0/r1 may still be read from an earlier call
  if (lane == 0) { r0[warp] = s0; r1[warp] = s1; }
  __syncthreads();
  if (warp == 0) {
    double a = lane < (int)(blockDim.x >> 5) ? r0[lane] : 0.0;
    double b = lane < (int)(blockDim.x >> 5) ? r1[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { partials[2 * blockIdx.x] = a; partials[2 * blockIdx.x + 1] = b; }
  }
}

__global__ void __launch_bounds__(kStageThreads) k_init_norms(const float* y0, const float* f0, const float* f1,
                                                              int64_t numel, float rtol, float atol, int mode,
                                                              double* partials) {
  init_norms_block(y0, f0, f1, numel, rtol, atol, mode, partials, (int64_t)blockIdx.x * blockDim.x + threadIdx.x,
                   (int64_t)gridDim.x * blockDim.x);
}

// one thread: h0 from the norms d0, d1 (phase 0); dt from d1, d2 and the controller's start time (phase 1)
__device__ __forceinline__ void init_scalar_decide(Ctrl& c, double sum0, double sum1, int phase, double t_first) {
  const float rootn = (float)sqrt(c.numel_global);  // numel ** 0.5, a Python float -> fp32 divisor
  if (phase == 0) {
    const float d0 = fdiv((float)sqrt(sum0), rootn);
    const float d1 = fdiv((float)sqrt(sum1), rootn);
    c.d0 = d0;
    c.d1 = d1;
    c.h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : fmul(0.01f, fdiv(d0, d1));
  } else {
    const float h0 = c.h0, d1 = c.d1;
    const float d2 = fdiv(fdiv((float)sqrt(sum0), rootn), h0);
    float h1;
    if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, fmul(h0, 1e-3f));
    else h1 = powf(fdiv(0.01f, fmaxf(d1, d2)), 1.0f / 5.0f);
    const float h = fminf(fmul(100.f, h0), h1);
    c.dt = (double)h;
    c.first_step = (double)h;
    c.t0 = c.t1 = t_first;
    if (!(c.t1 + c.dt > c.t1)) { c.status = NDCN_E_DT_UNDERFLOW; c.done = 1; }
  }
}

// phase 0: h0 from d0,d1.   phase 1: dt from d1,d2 and controller initialisation.
// xchg (multi-GPU): the two sums are staged there for the host hook's all-reduce;
// xchg_stage 0 = reduce+decide, 1 = reduce only, 2 = decide from xchg.
__global__ void __launch_bounds__(kStageThreadsCtl) k_init_scalar(Ctrl* ctrl, const double* partials, int n_partials,
                                                                  int phase, double t_first, double* xchg,
                                                                  int xchg_stage) {
  __shared__ double s_tmp[kStageThreadsCtl];
  double sums[2] = {0.0, 0.0};
  if (xchg_stage != 2) {
    for (int q = 0; q < 2; ++q) {
      double v = 0.0;
      for (int i = threadIdx.x; i < n_partials; i += blockDim.x) v += partials[2 * i + q];
      s_tmp[threadIdx.x] = v;
      __syncthreads();
      for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_tmp[threadIdx.x] += s_tmp[threadIdx.x + o];
        __syncthreads();
      }
      sums[q] = s_tmp[0];
      __syncthreads();
    }
  }
  if (threadIdx.x != 0) return;
  if (xchg_stage == 1) { xchg[0] = sums[0]; xchg[1] = sums[1]; return; }
  if (xchg_stage == 2) { sums[0] = xchg[0]; sums[1] = xchg[1]; }
  init_scalar_decide(*ctrl, sums[0], sums[1], phase, t_first);
}

// error-ratio sum as a stand-alone op (C ABI ndcn_error_ratio_f32)
__global__ void __launch_bounds__(kStageThreads) k_error_ratio(const float* err, const float* y0, const float* y1,
                                                               int64_t numel, float rtol, float atol, double* partials) {
  double s = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    const float tol = fadd(atol, fmul(rtol, fmaxf(fabsf(y0[i]), fabsf(y1[i]))));
    const float r = fdiv(err[i], tol);
    s += (double)fmul(r, r);
  }
  __shared__ double red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    double v = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) partials[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(kStageThreadsCtl) k_sum_partials(const double* partials, int n, double* out) {
  const double s = block_sum_partials(partials, n);
  if (threadIdx.x == 0) *out = s;
}

// Halo pack for the multi-GPU exchange: out[i, :] = x[idx[i], :].  One warp per row, 16-byte
// accesses when H % 4 == 0 (vec), so every warp moves one contiguous row segment per iteration.
__global__ void __launch_bounds__(kStageThreads) k_pack_rows(const float* __restrict__ x, const int32_t* __restrict__ idx,
                                                             int64_t n_idx, int H, float* __restrict__ out, int vec) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n_idx; i += nwarps) {
    const float* src = x + (int64_t)__ldg(idx + i) * H;
    float* dst = out + i * H;
    if (vec) {
      for (int c = lane * 4; c < H; c += 128) {
        float v[4];
        ldv<4>(src + c, v);
        stv<4>(dst + c, v);
      }
    } else {
      for (int c = lane; c < H; c += 32) dst[c] = src[c];
    }
  }
}

// W [n][k] -> Wt [k][n]   (once per solve; H*H elements)
__global__ void k_transpose(const float* __restrict__ W, float* __restrict__ Wt, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H * H) {
    const int n = i / H, k = i % H;
    Wt[(size_t)k * H + n] = W[i];
  }
}

// ---------------------------------------------------------------------------------------
// Parameter gradients of the Linear: dW[o][i] = sum_r gp[r][o] z[r][i], db[o] = sum_r gp[r][o]
// (what autograd computes through nn.Linear, neural_dynamics.py:33).  Split over row chunks (blockIdx.z): every CTA
// forms a 64x64 tile of the partial product of its chunk on FP32 FMA pipes (both operands are read as [rows, 64]
// tiles, i.e. coalesced along H); k_weight_grads_reduce then adds the chunks in chunk order -> reproducible.
// ---------------------------------------------------------------------------------------
constexpr int kWgTile = 64;
constexpr int kWgK = 16;

__global__ void __launch_bounds__(256) k_weight_grads_partial(const float* __restrict__ gp, const float* __restrict__ z,
                                                              int64_t n, int H, int64_t rows_per_chunk,
                                                              float* __restrict__ part /* [nz][H][H] */,
                                                              float* __restrict__ part_b /* [nz][H] or null */) {
  __shared__ float As[kWgK][kWgTile + 4];
  __shared__ float Bs[kWgK][kWgTile + 4];
  const int o0 = blockIdx.y * kWgTile, i0 = blockIdx.x * kWgTile;
  const int64_t r0 = (int64_t)blockIdx.z * rows_per_chunk;
  const int64_t r1 = min(n, r0 + rows_per_chunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float bsum = 0.f;  // column sum of gp for db: threads ty == 0 .. of the i-tile 0 CTAs
  const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;  // loader: 16 rows x 64 columns, 4 floats per thread
  for (int64_t r = r0; r < r1; r += kWgK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t rr = r + lr;
      float va = 0.f, vb = 0.f;
      if (rr < r1) {
        if (o0 + lc + q < H) va = gp[rr * H + o0 + lc + q];
        if (i0 + lc + q < H) vb = z[rr * H + i0 + lc + q];
      }
      As[lr][lc + q] = va;
      Bs[lr][lc + q] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kWgK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[k][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[k][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    if (part_b != nullptr && blockIdx.x == 0 && threadIdx.x < kWgTile) {
#pragma unroll
      for (int k = 0; k < kWgK; ++k) bsum += As[k][threadIdx.x];
    }
    __syncthreads();
  }
  float* dst = part + (size_t)blockIdx.z * H * H;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int o = o0 + ty * 4 + a;
    if (o >= H) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + tx * 4 + b;
      if (i < H) dst[(size_t)o * H + i] = acc[a][b];
    }
  }
  if (part_b != nullptr && blockIdx.x == 0 && threadIdx.x < kWgTile && o0 + (int)threadIdx.x < H)
    part_b[(size_t)blockIdx.z * H + o0 + threadIdx.x] = bsum;
}

// ---------------------------------------------------------------------------------------
// The same partial products on the tensor cores for H % 128 == 0 (the widths of BASELINE configs 2-5):
// 128x128 output tile per CTA, 16 warps of 32x32, K = 32 node rows per pipeline stage brought in by 16-byte
// cp.async (two stages), mma.sync m16n8k8 tf32 with both operands split into tf32 hi + lo in registers and
// lo*hi + hi*lo + hi*hi accumulated (3xTF32: the products keep ~21 mantissa bits).
// The tensor core adds into its fp32 accumulator by TRUNCATION, a bias that grows with the length of the chain
// (measured: relative error 1e-5 at 1 350 rows per CTA, 1e-4 at 13 500), so the MMA chain is restarted from zero
// every stage (12 accumulations) and the stage's partial is added to the running sum with a rounded FADD.
// Row pitch 136 floats: a fragment load touches rows t = 0..3 and columns g = 0..7 -> bank 8t + g, conflict-free.
// The K dimension (node rows) is the slow index of both operands, so no tcgen05 K-major image exists for them
// without a transposing producer; the legacy tensor path already takes this reduction from 0.79 ms (64x64
// FP32-FMA tiles; cuBLAS SGEMM 0.48 ms) to under 0.3 ms at 100k x 256.
// ---------------------------------------------------------------------------------------
constexpr int kWgmTile = 128;
constexpr int kWgmK = 32;
constexpr int kWgmThreads = 512;
constexpr int kWgmPitch = kWgmTile + 8;
constexpr int kWgmStageFloats = 2 * kWgmK * kWgmPitch;
constexpr int kWgmSmemBytes = 2 * kWgmStageFloats * (int)sizeof(float);

__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// rows [r, r + 32) of the two 128-column operand tiles -> one pipeline stage; rows past r1 are zero-filled
__device__ __forceinline__ void wgm_load_stage(float* stage, const float* __restrict__ gp, const float* __restrict__ z,
                                               int H, int o0, int i0, int64_t r, int64_t r1) {
  for (int v = threadIdx.x; v < 2 * kWgmK * (kWgmTile / 4); v += kWgmThreads) {
    const int which = v / (kWgmK * (kWgmTile / 4));
    const int w = v - which * (kWgmK * (kWgmTile / 4));
    const int row = w / (kWgmTile / 4), c4 = w - row * (kWgmTile / 4);
    float* dst = stage + which * (kWgmK * kWgmPitch) + row * kWgmPitch + c4 * 4;
    const int64_t rr = r + row;
    if (rr < r1) {
      const float* src = (which ? z + rr * H + i0 : gp + rr * H + o0) + c4 * 4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kWgmThreads, 1) k_weight_grads_mma(const float* __restrict__ gp,
                                                                     const float* __restrict__ z, int64_t n, int H,
                                                                     int64_t rows_per_chunk,
                                                                     float* __restrict__ part /* [nz][H][H] */,
                                                                     float* __restrict__ part_b /* [nz][H] or null */) {
  extern __shared__ __align__(16) float wgm_smem[];
  const int o0 = blockIdx.y * kWgmTile, i0 = blockIdx.x * kWgmTile;
  const int64_t r0 = (int64_t)blockIdx.z * rows_per_chunk;
  const int64_t r1 = min(n, r0 + rows_per_chunk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;  // the warp's 32 x 32 block of the tile
  float sum[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) sum[a][b][c] = 0.f;
  float bsum = 0.f;
  const bool want_b = part_b != nullptr && blockIdx.x == 0 && threadIdx.x < kWgmTile;
  const int64_t n_it = r1 > r0 ? (r1 - r0 + kWgmK - 1) / kWgmK : 0;
  if (n_it > 0) wgm_load_stage(wgm_smem, gp, z, H, o0, i0, r0, r1);
  for (int64_t it = 0; it < n_it; ++it) {
    float* cur = wgm_smem + (it & 1) * kWgmStageFloats;
    if (it + 1 < n_it) {
      wgm_load_stage(wgm_smem + ((it + 1) & 1) * kWgmStageFloats, gp, z, H, o0, i0, r0 + (it + 1) * kWgmK, r1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* As = cur;                       // [k][o]
    const float* Bs = cur + kWgmK * kWgmPitch;  // [k][i]
    float acc[2][4][4];  // this stage's 32 rows, chain restarted from zero
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
#pragma unroll
    for (int k0 = 0; k0 < kWgmK; k0 += 8) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const float* p = As + (k0 + t) * kWgmPitch + wm + a * 16 + g;
        tf32_split(p[0], ah[a][0], al[a][0]);
        tf32_split(p[8], ah[a][1], al[a][1]);
        tf32_split(p[4 * kWgmPitch], ah[a][2], al[a][2]);
        tf32_split(p[4 * kWgmPitch + 8], ah[a][3], al[a][3]);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float* p = Bs + (k0 + t) * kWgmPitch + wn + b * 8 + g;
        uint32_t bh[2], bl[2];
        tf32_split(p[0], bh[0], bl[0]);
        tf32_split(p[4 * kWgmPitch], bh[1], bl[1]);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          mma_tf32_16x8x8(acc[a][b], al[a], bh);
          mma_tf32_16x8x8(acc[a][b], ah[a], bl);
          mma_tf32_16x8x8(acc[a][b], ah[a], bh);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) sum[a][b][c] += acc[a][b][c];
    if (want_b) {
#pragma unroll
      for (int k = 0; k < kWgmK; ++k) bsum += As[k * kWgmPitch + threadIdx.x];
    }
    __syncthreads();
  }
  float* dst = part + (size_t)blockIdx.z * H * H;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int o = o0 + wm + a * 16 + g, i = i0 + wn + b * 8 + 2 * t;
      *reinterpret_cast<float2*>(dst + (size_t)o * H + i) = make_float2(sum[a][b][0], sum[a][b][1]);
      *reinterpret_cast<float2*>(dst + (size_t)(o + 8) * H + i) = make_float2(sum[a][b][2], sum[a][b][3]);
    }
  if (want_b) part_b[(size_t)blockIdx.z * H + o0 + threadIdx.x] = bsum;
}

// out[j] (+)= sum_c part[c][j], chunks added in order
__global__ void __launch_bounds__(256) k_weight_grads_reduce(const float* __restrict__ part, int nz, int64_t count,
                                                             float* __restrict__ out, int accumulate) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  float s = accumulate ? out[j] : 0.f;
  for (int c = 0; c < nz; ++c) s += part[(size_t)c * count + j];
  out[j] = s;
}

// out[q][r][c] = x[r][q * bc + c]: the row-sharded state as H / bc column blocks, block q being what
// peer q gathers from (feature-sharded multi-GPU exchange).  16-byte accesses, grid-stride.
__global__ void __launch_bounds__(kStageThreads) k_pack_cols(const float* __restrict__ x, int64_t n_rows, int H, int bc,
                                                             float* __restrict__ out) {
  const int h4 = H / 4, bc4 = bc / 4;
  const int64_t n4 = n_rows * h4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int64_t r = i / h4;
    const int c4 = (int)(i - r * h4);
    const int q = c4 / bc4, cc = c4 - q * bc4;
    const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
    reinterpret_cast<float4*>(out)[((int64_t)q * n_rows + r) * bc4 + cc] = v;
  }
}

}  // namespace ndcn

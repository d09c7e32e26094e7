// C ABI of libndcn_b200.so + the host-side solve drivers (see include/ndcn_b200.h).
//
// The drivers only ENQUEUE: the adaptive dopri5 loop keeps t, dt, accept/reject, the FSAL
// buffer parity and the dense-output bookkeeping in a device-resident controller block and
// polls it once per batch of step attempts (reference: >=3 host syncs per step,
// torchdiffeq/_impl/dopri5.py:88-109).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include <cstdlib>

#include "ndcn_common.cuh"
#include "solver_kernels.cuh"
#include "stage_kernels.cuh"
#include "gather_kernels.cuh"
#include "umma_kernels.cuh"
#include "small_solver.cuh"

using namespace ndcn;

#define CU_TRY(expr)                          \
  do {                                        \
    cudaError_t _e = (expr);                  \
    if (_e != cudaSuccess) return (int)_e;    \
  } while (0)
#define RC_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != 0) return _rc;   \
  } while (0)

struct ndcn_graph {
  GraphView v;
  int32_t* long_rows = nullptr;  // rows with more than kLongRow entries (device)
  int n_long = 0;
  int max_deg = 0;
  std::vector<int32_t> long_host;  // the same list on the host, ascending (row-chunked gathers take sub-ranges)
};

// ---------------------------------------------------------------------------------------
// library configuration (ndcn_config_set / environment, read once)
// ---------------------------------------------------------------------------------------
struct Config {
  int stage_impl = NDCN_IMPL_AUTO;  // which kernel family evaluates relu((Phi x) W^T + b)
  int gather_cw = 0;                // 0 auto, -1 full-row gather, else chunk width in floats (16/32/64)
  int64_t umma_min_rows = 8192;     // auto: tcgen05 path from this many rows
  int gather_v = 1;                 // chunk-major gather flavour: 1 one row per lane group, 2 persistent + TMA-staged CSR
  int64_t z_chunk_rows = 0;         // > 0: gather and tcgen05 stage kernel alternate over row chunks of this size (Z in L2)
  int small_solver = 1;             // 1: solves that fit the persistent whole-solve kernel use it (ndcn_odeint_f32 auto)
  int64_t small_max_rows = 16384;   // ... up to this many rows
  int64_t small_max_numel = 1 << 21;  // ... and this many state elements (8 MB per buffer: everything stays in L2)
};
static Config& cfg() {
  static Config c = [] {
    Config k;
    if (const char* v = std::getenv("NDCN_STAGE_IMPL")) k.stage_impl = std::atoi(v);
    if (const char* v = std::getenv("NDCN_GATHER_CW")) k.gather_cw = std::atoi(v);
    if (const char* v = std::getenv("NDCN_UMMA_MIN_ROWS")) k.umma_min_rows = std::atoll(v);
    if (const char* v = std::getenv("NDCN_GATHER_V")) k.gather_v = std::atoi(v);
    if (const char* v = std::getenv("NDCN_SMALL_SOLVER")) k.small_solver = std::atoi(v) != 0;
    if (const char* v = std::getenv("NDCN_Z_CHUNK_ROWS")) k.z_chunk_rows = std::atoll(v) / kUmmaM * kUmmaM;
    return k;
  }();
  return c;
}

extern "C" int ndcn_config_set(int32_t key, int64_t value) {
  switch (key) {
    case NDCN_CFG_STAGE_IMPL:
      if (value < NDCN_IMPL_AUTO || value > NDCN_IMPL_UMMA) return NDCN_E_ARG;
      cfg().stage_impl = (int)value;
      return NDCN_OK;
    case NDCN_CFG_GATHER_CW:
      if (!(value == 0 || value == -1 || value == 16 || value == 32 || value == 64)) return NDCN_E_ARG;
      cfg().gather_cw = (int)value;
      return NDCN_OK;
    case NDCN_CFG_UMMA_MIN_ROWS:
      if (value < 0) return NDCN_E_ARG;
      cfg().umma_min_rows = value;
      return NDCN_OK;
    case NDCN_CFG_GATHER_VERSION:
      if (value != 1 && value != 2) return NDCN_E_ARG;
      cfg().gather_v = (int)value;
      return NDCN_OK;
    case NDCN_CFG_SMALL_SOLVER:
      if (value != 0 && value != 1) return NDCN_E_ARG;
      cfg().small_solver = (int)value;
      return NDCN_OK;
    default:
      return NDCN_E_ARG;
  }
}

extern "C" int64_t ndcn_config_get(int32_t key) {
  switch (key) {
    case NDCN_CFG_STAGE_IMPL: return cfg().stage_impl;
    case NDCN_CFG_GATHER_CW: return cfg().gather_cw;
    case NDCN_CFG_UMMA_MIN_ROWS: return cfg().umma_min_rows;
    case NDCN_CFG_GATHER_VERSION: return cfg().gather_v;
    case NDCN_CFG_SMALL_SOLVER: return cfg().small_solver;
    default: return NDCN_E_ARG;
  }
}

// Dormand-Prince / Shampine tableau, torchdiffeq/_impl/dopri5.py:11-36 (doubles, cast to fp32
// exactly where the reference's tensor*python-float multiplication does, misc.py:25)
static const double kDpBeta[6][6] = {
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84},
};
static const double kDpAlpha[6] = {1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};  // stage times, dopri5.py:12
static const double kDpErr[7] = {
    35.0 / 384 - 1951.0 / 21600, 0, 500.0 / 1113 - 22642.0 / 50085, 125.0 / 192 - 451.0 / 720,
    -2187.0 / 6784 - -12231.0 / 42400, 11.0 / 84 - 649.0 / 6300, -1.0 / 60.0,
};
static const double kDpMid[7] = {
    6025192743.0 / 30085553152.0 / 2, 0, 51252292925.0 / 65400821598.0 / 2, -2691868925.0 / 45128329728.0 / 2,
    187940372067.0 / 1594534317056.0 / 2, -1776094331.0 / 19743644256.0 / 2, 11237099.0 / 235043384.0 / 2,
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
// SM count of the current device (148 on B200), queried once per device
static int sm_count_now() {
  static int cached[64] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
// cudaFuncSetAttribute is per device: remember it per (kernel instantiation, device)
// (marked only AFTER the call has returned: ranks that run as threads of one process may race here, and a
// duplicate cudaFuncSetAttribute is harmless while a launch that overtakes the first one is not)
struct PerDeviceOnce {
  volatile bool done[64] = {};
  bool need() const {
    const int dev = current_device();
    return dev < 0 || dev >= 64 || !done[dev];
  }
  void mark() {
    const int dev = current_device();
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
};

// Stream-ordered scratch (cudaMallocAsync) comes from the device's default memory pool; with the
// default release threshold of 0 the pool hands its memory back to the OS at every synchronisation
// and the next allocation pays for a fresh mapping (~100 us).  Keep freed scratch in the pool.
static void keep_scratch_pooled() {
  static bool done = false;
  if (done) return;
  done = true;
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = 1ull << 30;  // up to 1 GiB of idle scratch stays mapped
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
}
static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }
static inline int ilog2(int64_t v) {
  int l = 0;
  while (v > 1) {
    v >>= 1;
    ++l;
  }
  return l;
}
static inline PtrPair pp(float* a) { return PtrPair{{a, a}}; }
static inline PtrPair pp(float* a, float* b) { return PtrPair{{a, b}}; }

struct ndcn_solver {
  const ndcn_graph* g = nullptr;
  ndcn_rhs_desc_t rhs{};
  int method = 0;
  int64_t n_rows = 0, n_cols = 0;
  int H = 0;
  int64_t numel = 0;      // n_rows * H
  int64_t numel_src = 0;  // n_cols * H
  // caller-provided workspace
  float* Y[2] = {nullptr, nullptr};   // state ping-pong (gather sources, n_cols rows)
  float* YS[2] = {nullptr, nullptr};  // stage inputs (gather sources, n_cols rows)
  float* KF[2] = {nullptr, nullptr};  // f0 / k7 (FSAL pair)
  float* K[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float* Wt = nullptr;
  float* Z = nullptr;     // Phi x between the gather and the tcgen05 GEMM kernel
  float* Wimg = nullptr;  // W split into tf32 hi/lo, swizzled K-atoms (k_prep_w_image)
  // library-owned control scratch
  double* partials = nullptr;
  int max_partials = 0;
  double* xchg = nullptr;     // 2 doubles for the multi-GPU all-reduce hook
  float* t_stage = nullptr;   // fp32 stage time for callback RHS
  double* t_out = nullptr;
  int t_cap = 0;
  Ctrl* ctrl = nullptr;
  Ctrl* ctrl_host = nullptr;  // pinned
  int64_t launches = 0;
  int sm_count = 148;
  // multi-GPU peer push (ndcn_solver_set_peers): world > 1 switches it on
  PeerArgs peers{};
  int n_push = 0;                      // world - 1
  long long push_delta[kMaxPeers] = {};  // byte distance local gather-source element -> the same element at peer j
  // feature-sharded peer push (ndcn_solver_set_feature_peers)
  bool feat_on = false;
  FeatTable* feat_dev = nullptr;       // library-owned copy of the pointer table
  const ndcn_graph* full_graph = nullptr;
  float* xcs_self = nullptr;           // this rank's slice buffer [N, Hc]
  float* Z_own = nullptr;              // workspace Z, restored when the scheme is switched off
  int feat_rank = 0, feat_world = 0, feat_hc = 0;
  int feat_bounds[9] = {};
  bool feat_slab = true;               // slice buffers column-blocked [Hc/16][N][16] + k_gather_slab (NDCN_FEAT_SLAB=0: off)
};

// ---------------------------------------------------------------------------------------
// RHS stage dispatch
// ---------------------------------------------------------------------------------------
static int grid_for_elems(int64_t n, int sm_count) {
  int64_t blocks = (n + kStageThreads * 4 - 1) / (kStageThreads * 4);
  int64_t cap = (int64_t)sm_count * 16;
  return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

template <int VW, int NCH>
static int launch_ndcn_fast(const NdcnArgs& a, EpiArgs& e, int* grid_out, cudaStream_t st) {
  const int64_t n = a.g.n_rows;
  if (a.flags & NDCN_F_NO_CONTROL) {
    const int n_long = (a.flags & NDCN_F_NO_GRAPH) ? 0 : a.n_long;
    const int grid = (int)((n + kWarpsPerCta - 1) / kWarpsPerCta) + n_long;
    *grid_out = grid;
    // H=256: 4 CTAs/SM (64 registers) x 4 row loads in flight per lane measured best on B200
    // (1.69 ms vs 1.86 ms at 3 CTAs/SM for the 1M-node power-law gather, profiles/README.md)
    // plain z = Phi x feeding the tcgen05 GEMM kernel: the instantiation without the stage algebra
    const bool store_only = e.mode == EPI_STORE && e.feat_mode == FEAT_OFF && e.k_out.p[0] != nullptr && !(a.flags & NDCN_F_NO_GRAPH);
    if constexpr (VW == 4 && NCH == 2) {
      if (store_only) k_stage_ndcn_row<VW, NCH, 4, 4, true><<<grid, kStageThreads, 0, st>>>(a, e);
      else k_stage_ndcn_row<VW, NCH, 4, 4><<<grid, kStageThreads, 0, st>>>(a, e);
    } else if constexpr (VW == 4 && NCH == 1) {
      // H=128: 8 row loads in flight per lane at 5 CTAs/SM measured best (0.98 ms vs 1.34 ms for the
      // default instantiation on the 1M-node power-law gather)
      if (store_only) k_stage_ndcn_row<VW, NCH, 8, 5, true><<<grid, kStageThreads, 0, st>>>(a, e);
      else k_stage_ndcn_row<VW, NCH, 8, 5><<<grid, kStageThreads, 0, st>>>(a, e);
    } else {
      k_stage_ndcn_row<VW, NCH><<<grid, kStageThreads, 0, st>>>(a, e);
    }
  } else {
    using S = GemmSmem<VW, NCH>;
    static PerDeviceOnce attr;
    if (attr.need()) {
      CU_TRY(cudaFuncSetAttribute(k_stage_ndcn_gemm<VW, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)S::total));
      attr.mark();
    }
    const int grid = (int)((n + kTileRows - 1) / kTileRows);
    *grid_out = grid;
    k_stage_ndcn_gemm<VW, NCH><<<grid, kStageThreads, S::total, st>>>(a, e);
  }
  return (int)cudaGetLastError();
}

template <int KIND>
static int launch_dyn(const DynArgs& a, EpiArgs& e, double avg_deg, int* grid_out, cudaStream_t st) {
  const int64_t n = a.g.n_rows;
  if (a.d == 1 && n >= 32768 && KIND != NDCN_RHS_MUTUAL) {
    // at scale: CSR-stream kernel (slice staged in shared memory, entry-parallel terms, CSR-order row sums).  Measured
    // at 1M nodes, ms per dopri5 step: heat 1.89 -> 1.24, gene 1.97 -> 1.50, mutualistic 1.91 -> 2.60 (its per-entry
    // term is division-heavy and gains nothing from the staging): the mutualistic RHS keeps the lane-group kernel
    const int grid = (int)((n + kDynRows - 1) / kDynRows) + a.n_long;
    *grid_out = grid;
    k_stage_dyn1_stream<KIND><<<grid, kStageThreads, 0, st>>>(a, e);
  } else if (a.d == 1) {
    int lpr = 4;
    if (avg_deg > 24) lpr = 32;
    else if (avg_deg > 12) lpr = 16;
    else if (avg_deg > 6) lpr = 8;
    const int64_t rows_per_block = (int64_t)kStageThreads / lpr;
    const int grid = (int)((n + rows_per_block - 1) / rows_per_block) + a.n_long;
    *grid_out = grid;
    switch (lpr) {
      case 4: k_stage_dyn1<KIND, 4><<<grid, kStageThreads, 0, st>>>(a, e); break;
      case 8: k_stage_dyn1<KIND, 8><<<grid, kStageThreads, 0, st>>>(a, e); break;
      case 16: k_stage_dyn1<KIND, 16><<<grid, kStageThreads, 0, st>>>(a, e); break;
      default: k_stage_dyn1<KIND, 32><<<grid, kStageThreads, 0, st>>>(a, e); break;
    }
  } else {
    const int grid = (int)((n + kWarpsPerCta - 1) / kWarpsPerCta);
    *grid_out = grid;
    k_stage_dynv<KIND><<<grid, kStageThreads, 0, st>>>(a, e);
  }
  return (int)cudaGetLastError();
}

struct StageTimer {  // optional per-launch timing hook (Driver implements it)
  virtual void begin(int cls, bool count = true) = 0;  // count: this launch opens a new RHS evaluation (vs another chunk of it)
  virtual void end() = 0;
};

struct RhsBinding {  // everything needed to launch one RHS stage
  const ndcn_graph* g;
  const ndcn_rhs_desc_t* rhs;
  const float* Wt;
  double* partials;
  int max_partials;
  int sm_count;
  float* Z = nullptr;           // [n_rows, H] scratch (tcgen05 path)
  const float* Wimg = nullptr;  // tf32 hi/lo W image (tcgen05 path)
  StageTimer* timer = nullptr;
  int64_t* launches = nullptr;
  int z_block_cols = 0;         // > 0: Z already holds Phi x in column blocks of this width (external gather)
};

// ---- chunk-major gather ------------------------------------------------------------------
static bool fast_width(int H) { return H == 256 || H == 128 || H == 64 || H == 32; }

// chunk width (floats) of the gather for this problem; 0 = one pass over full rows
static int pick_gather_cw(int64_t n_cols, int H) {
  if (H % 16 != 0) return 0;
  const int want = cfg().gather_cw;
  if (want < 0) return fast_width(H) ? 0 : 16;
  if (want > 0) return (H % want == 0) ? want : 16;
  // measured on B200 (profiles/README.md, round 1): without a way to pin the [N, cw] slab in L2 the
  // chunk-major order does not beat one pass over full rows, so auto = full rows whenever the width
  // has a full-row kernel
  if (fast_width(H)) {
    // mid-size states: a 32-column slab (n_cols x 128 B) that fits L2 comfortably makes the chunk-major
    // order pay (100k nodes: 0.17 ms vs 0.45 ms, 250k: 0.49 vs 0.88 ms for H=256)
    const double state_mb = (double)n_cols * H * 4.0 / 1048576.0;
    const double slab_mb = (double)n_cols * 128.0 / 1048576.0;
    if (H > 32 && state_mb > 64.0 && slab_mb <= 48.0) return 32;
    // narrow states (the column slices of the feature-sharded multi-GPU gather): 8 lanes x 16 bytes per
    // row beat the one-warp-per-row kernels whose lanes load 4 / 8 bytes (1M nodes: H=32 0.35 vs 1.17 ms,
    // H=64 0.62 vs 0.89 ms; at H=128 the tuned full-row kernel wins, 0.98 vs 1.17 ms)
    if (H <= 64 && n_cols >= 4096) return 32;
    return 0;
  }
  const int cands[3] = {64, 32, 16};
  for (int cw : cands)
    if (H % cw == 0 && (double)n_cols * cw * 4.0 / 1048576.0 <= 72.0) return cw;
  return 16;
}

static bool gather_uses_v2(int cw, uint32_t flags) {
  return cfg().gather_v == 2 && (cw == 32 || cw == 64) && !(flags & NDCN_F_NO_GRAPH);
}

static int gather_grid(const ndcn_graph* g, int H, int cw) {
  const int lpr = cw / 4;
  const int rpc = (32 / lpr) * kWarpsPerCta;
  const int64_t n_rb = (g->v.n_rows + rpc - 1) / rpc;
  return (int)((H / cw) * (n_rb + g->n_long));
}

// z/k = [relu](Phi x) fused with epilogue e (NdcnArgs::flags: NO_RELU / NO_GRAPH honoured)
static int launch_gather(const RhsBinding& b, const NdcnArgs& a, int H, EpiArgs& e, int* grid_out, cudaStream_t st) {
  const int cw = pick_gather_cw(a.g.n_cols, H);
  if (cw == 0) {
    NdcnArgs a2 = a;
    a2.flags |= NDCN_F_NO_CONTROL;
    switch (H) {
      case 256: return launch_ndcn_fast<4, 2>(a2, e, grid_out, st);
      case 128: return launch_ndcn_fast<4, 1>(a2, e, grid_out, st);
      case 64: return launch_ndcn_fast<2, 1>(a2, e, grid_out, st);
      default: return launch_ndcn_fast<1, 1>(a2, e, grid_out, st);
    }
  }
  if (gather_uses_v2(cw, a.flags) && e.feat_mode != FEAT_Z_OWNERS) {
    const int n_blocks = (int)((a.g.n_rows + kG2Rows - 1) / kG2Rows);
    const int64_t total = (int64_t)(H / cw) * (n_blocks + b.g->n_long);
    const bool store_only = e.mode == EPI_STORE;
    const int grid = (int)std::min<int64_t>(total, (int64_t)b.sm_count * (store_only ? 3 : 2));
    *grid_out = grid;
    if (grid == 0) return 0;
    if (cw == 64 && store_only)
      k_stage_gather_v2<64, true><<<grid, kG2Threads, 0, st>>>(a, H, n_blocks, b.g->n_long, b.g->long_rows, e);
    else if (cw == 64)
      k_stage_gather_v2<64, false><<<grid, kG2Threads, 0, st>>>(a, H, n_blocks, b.g->n_long, b.g->long_rows, e);
    else if (store_only)
      k_stage_gather_v2<32, true><<<grid, kG2Threads, 0, st>>>(a, H, n_blocks, b.g->n_long, b.g->long_rows, e);
    else
      k_stage_gather_v2<32, false><<<grid, kG2Threads, 0, st>>>(a, H, n_blocks, b.g->n_long, b.g->long_rows, e);
    return (int)cudaGetLastError();
  }
  const int lpr = cw / 4;
  const int rpc = (32 / lpr) * kWarpsPerCta;
  const int n_rb = (int)((a.g.n_rows + rpc - 1) / rpc);
  const int grid = gather_grid(b.g, H, cw);
  *grid_out = grid;
  if (grid == 0) return 0;
  switch (cw) {
    case 64: k_stage_gather_chunk<64><<<grid, kStageThreads, 0, st>>>(a, H, n_rb, b.g->n_long, b.g->long_rows, e); break;
    case 32: k_stage_gather_chunk<32><<<grid, kStageThreads, 0, st>>>(a, H, n_rb, b.g->n_long, b.g->long_rows, e); break;
    default: k_stage_gather_chunk<16><<<grid, kStageThreads, 0, st>>>(a, H, n_rb, b.g->n_long, b.g->long_rows, e); break;
  }
  return (int)cudaGetLastError();
}

// ---- tcgen05 GEMM + epilogue ---------------------------------------------------------------
template <int H, int MODE, int NPREV, int B>
static int launch_umma_inst(const UmmaArgs& u, EpiArgs& e, int grid, cudaStream_t st) {
  using Cf = UmmaCfg<H>;
  static PerDeviceOnce attr;
  if (attr.need()) {
    CU_TRY(cudaFuncSetAttribute(k_stage_gemm_umma<H, MODE, NPREV, B>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)Cf::kSmemBytes));
    attr.mark();
  }
  k_stage_gemm_umma<H, MODE, NPREV, B><<<grid, kUmmaThreads, Cf::kSmemBytes, st>>>(u, e);
  return (int)cudaGetLastError();
}

// one instantiation per Runge-Kutta stage shape (epilogue mode x number of earlier stages read)
template <int H>
static int launch_umma(const UmmaArgs& u, EpiArgs& e, int sm_count, int* grid_out, cudaStream_t st) {
  const int64_t n_tiles = u.tile_end > 0 ? u.tile_end - u.tile_begin : (u.n_rows + kUmmaM - 1) / kUmmaM;
  const int grid = (int)std::min<int64_t>(n_tiles, sm_count);
  *grid_out = grid;
  if (grid == 0) return 0;
  const bool deep = u.deep_batch != 0;  // experiment switch: the other epilogue batch depth
#define NDCN_UMMA_CASE(MODE, NPREV, B0, B1) \
  return deep ? launch_umma_inst<H, MODE, NPREV, B1>(u, e, grid, st) : launch_umma_inst<H, MODE, NPREV, B0>(u, e, grid, st)
  switch (e.mode) {
    case EPI_STORE: NDCN_UMMA_CASE(EPI_STORE, 0, 2, 1);
    case EPI_LINCOMB:
      switch (e.n_prev) {
        case 0: NDCN_UMMA_CASE(EPI_LINCOMB, 0, 2, 1);
        case 1: NDCN_UMMA_CASE(EPI_LINCOMB, 1, 2, 1);
        case 2: NDCN_UMMA_CASE(EPI_LINCOMB, 2, 2, 1);
        // batch depth 2 measured 2-5 % faster than 1 at 3-4 earlier stages in round 2 (ncu, profiles/README.md)
        case 3: NDCN_UMMA_CASE(EPI_LINCOMB, 3, 2, 1);
        case 4: NDCN_UMMA_CASE(EPI_LINCOMB, 4, 2, 1);
        case 5: NDCN_UMMA_CASE(EPI_LINCOMB, 5, 1, 2);
        default: return NDCN_E_ARG;
      }
    case EPI_LINCOMB_E:
      switch (e.n_prev) {
        case 4: NDCN_UMMA_CASE(EPI_LINCOMB_E, 4, 2, 1);
        case 5: NDCN_UMMA_CASE(EPI_LINCOMB_E, 5, 1, 2);
        default: return NDCN_E_ARG;
      }
    case EPI_ERR:
      switch (e.n_prev) {
        case 1: NDCN_UMMA_CASE(EPI_ERR, 1, 2, 1);
        case 5: NDCN_UMMA_CASE(EPI_ERR, 5, 1, 2);
        case 6: NDCN_UMMA_CASE(EPI_ERR, 6, 1, 2);
        default: return NDCN_E_ARG;
      }
    case EPI_RK4_1: NDCN_UMMA_CASE(EPI_RK4_1, 0, 2, 1);
    case EPI_RK4_2: NDCN_UMMA_CASE(EPI_RK4_2, 1, 2, 1);
    case EPI_RK4_3: NDCN_UMMA_CASE(EPI_RK4_3, 2, 2, 1);
    case EPI_RK4_4: NDCN_UMMA_CASE(EPI_RK4_4, 3, 1, 2);
    case EPI_MASK: NDCN_UMMA_CASE(EPI_MASK, 0, 2, 1);
    default: return NDCN_E_ARG;
  }
#undef NDCN_UMMA_CASE
}

static bool umma_eligible(const ndcn_rhs_desc_t& r, int64_t n_rows) {
  if (r.kind != NDCN_RHS_NDCN || (r.flags & NDCN_F_NO_CONTROL)) return false;
  if (r.H != 256 && r.H != 128) return false;
  const int impl = cfg().stage_impl;
  if (impl == NDCN_IMPL_SIMT) return false;
  if (impl == NDCN_IMPL_UMMA) return true;
  return n_rows >= cfg().umma_min_rows;
}

static EpiArgs store_only(float* out);

// The reference multiplies every tableau coefficient in, zeros included (misc.py:18-25).  A term
// (dt*0)*k_j is +-0 and leaves the fp32 running sum unchanged, so the stage that carries it can
// skip reading k_j: dopri5's b_72 and c_err_2 are zero -> 2 of 27 state reads per step.  (Only a
// non-finite k_j would make the product NaN, and such a k_j has already poisoned the stage inputs
// built from it with non-zero coefficients, so the non-finite guard fires either way.)
static void drop_zero_terms(EpiArgs& e) {
  if (e.mode != EPI_LINCOMB && e.mode != EPI_ERR && e.mode != EPI_LINCOMB_E) return;
  if (e.n_prev < 2 || e.err_prefix) return;
  const bool two = e.mode == EPI_LINCOMB_E;  // a stream is skipped only if BOTH sums carry a zero for it
  int w = 0;
  for (int j = 0; j < e.n_prev; ++j) {
    const bool zero = e.beta[j] == 0.0f && (!two || e.ebeta[j] == 0.0f);
    if (zero && !(w == 0 && j == e.n_prev - 1)) continue;  // keep at least one earlier term
    e.kprev[w] = e.kprev[j];
    e.beta[w] = e.beta[j];
    e.ebeta[w] = e.ebeta[j];
    ++w;
  }
  if (w == e.n_prev) return;
  e.beta[w] = e.beta[e.n_prev];  // coefficient of the fresh k
  e.ebeta[w] = e.ebeta[e.n_prev];
  for (int j = w + 1; j < 8; ++j) e.beta[j] = e.ebeta[j] = 0.0f;
  e.n_prev = w;
}

// Launches f(src) fused with epilogue `e`; returns the number of per-CTA partial slots the
// kernel writes in EPI_ERR mode through *n_partials.
static int launch_stage(const RhsBinding& b, PtrPair src, EpiArgs e, int* n_partials, cudaStream_t st) {
  const ndcn_rhs_desc_t& r = *b.rhs;
  e.partials = b.partials;
  drop_zero_terms(e);
  int grid = 0;
  int rc = 0;
  auto tick = [&](int cls) {
    if (b.timer) b.timer->begin(cls);
    if (b.launches) *b.launches += 1;
  };
  auto tock = [&]() {
    if (b.timer) b.timer->end();
  };
  if (r.kind == NDCN_RHS_NDCN) {
    NdcnArgs a;
    a.row_begin = a.row_end = 0;
    a.keep_l2 = 0;
    a.g = b.g->v;
    a.x = src;
    a.Wt = b.Wt;
    a.bias = r.b;
    a.flags = r.flags;
    a.long_rows = b.g->long_rows;
    a.n_long = b.g->n_long;
    const bool need_w = !(r.flags & NDCN_F_NO_CONTROL);
    if (need_w && (r.W == nullptr || r.b == nullptr)) return NDCN_E_ARG;
    if (r.H < 1 || r.H > 1024) return NDCN_E_ARG;
    const bool src_aligned = aligned16(src.p[0]) && aligned16(src.p[1]);
    const bool external_z = b.z_block_cols > 0;
    if (external_z && !(umma_eligible(r, ((int64_t)1) << 40) && b.Z != nullptr && b.Wimg != nullptr)) return NDCN_E_ARG;
    if (external_z || (umma_eligible(r, a.g.n_rows) && b.Z != nullptr && b.Wimg != nullptr && src_aligned)) {
      // (1) z = Phi x  -> Z (skipped with no_graph)   (2) k = relu(z W^T + b) + stage epilogue
      UmmaArgs u;
      u.z = src;
      u.z_block_log2 = 0;
      u.tile_begin = u.tile_end = 0;
      u.partials_off = 0;
      for (int w = external_z ? b.z_block_cols : r.H; w > 1; w >>= 1) u.z_block_log2 += 1;
      u.wimg = b.Wimg;
      u.bias = r.b;
      u.n_rows = a.g.n_rows;
      u.flags = r.flags;
      {
        static const uint32_t dbg = [] {
          const char* v = std::getenv("NDCN_UMMA_DBG");
          return v ? (uint32_t)std::atoi(v) : 0u;
        }();
        u.deep_batch = (dbg >> 8) & 1u;
      }
      const int64_t chunk = cfg().z_chunk_rows;
      if (!external_z && !(r.flags & NDCN_F_NO_GRAPH) && chunk > 0 && a.g.n_rows > 2 * chunk &&
          pick_gather_cw(a.g.n_cols, r.H) == 0 && (r.H == 256 || r.H == 128)) {
        // Z L2-resident: gather and stage kernel alternate over row chunks; the chunk's z lives in one of two
        // chunk-sized halves of Z (plain stores, read back out of L2 by the A producers' streaming loads)
        const std::vector<int32_t>& lh = b.g->long_host;
        size_t lo = 0;
        int ci = 0, p_off = 0;
        for (int64_t r0 = 0; r0 < a.g.n_rows; r0 += chunk, ++ci) {
          const int64_t r1 = std::min<int64_t>(a.g.n_rows, r0 + chunk);
          float* buf = b.Z + (size_t)(ci & 1) * chunk * r.H;
          float* zbase = buf - (size_t)r0 * r.H;  // row index stays global
          size_t hi = lo;
          while (hi < lh.size() && lh[hi] < r1) ++hi;
          NdcnArgs ga = a;
          ga.flags = NDCN_F_NO_CONTROL | NDCN_F_NO_RELU;
          ga.row_begin = r0;
          ga.row_end = r1;
          ga.keep_l2 = 1;
          ga.long_rows = b.g->long_rows + lo;
          ga.n_long = (int)(hi - lo);
          lo = hi;
          EpiArgs se = store_only(zbase);
          se.ctrl = e.ctrl;
          const int ggrid = (int)((r1 - r0 + kWarpsPerCta - 1) / kWarpsPerCta) + ga.n_long;
          if (b.timer) b.timer->begin(NDCN_K_GATHER, ci == 0);
          if (b.launches) *b.launches += 1;
          if (r.H == 256) k_stage_ndcn_row<4, 2, 4, 4, true><<<ggrid, kStageThreads, 0, st>>>(ga, se);
          else k_stage_ndcn_row<4, 1, 8, 5, true><<<ggrid, kStageThreads, 0, st>>>(ga, se);
          tock();
          u.z = pp(zbase);
          u.tile_begin = r0 / kUmmaM;
          u.tile_end = (r1 + kUmmaM - 1) / kUmmaM;
          u.partials_off = p_off;
          if (b.timer) b.timer->begin(NDCN_K_STAGE, ci == 0);
          if (b.launches) *b.launches += 1;
          int cgrid = 0;
          rc = r.H == 256 ? launch_umma<256>(u, e, b.sm_count, &cgrid, st) : launch_umma<128>(u, e, b.sm_count, &cgrid, st);
          tock();
          if (rc != 0) return rc;
          p_off += cgrid;
        }
        grid = p_off;
      } else {
        if (external_z) {
          u.z = pp(b.Z);  // the exchange hook has filled it (Driver::stage)
        } else if (!(r.flags & NDCN_F_NO_GRAPH)) {
          NdcnArgs ga = a;
          ga.flags = NDCN_F_NO_CONTROL | NDCN_F_NO_RELU;
          EpiArgs se = store_only(b.Z);
          se.ctrl = e.ctrl;  // same buffer parity / done flag as the stage that consumes Z
          int g2 = 0;
          tick(NDCN_K_GATHER);
          rc = launch_gather(b, ga, r.H, se, &g2, st);
          tock();
          if (rc != 0) return rc;
          u.z = pp(b.Z);
        }
        tick(NDCN_K_STAGE);
        rc = r.H == 256 ? launch_umma<256>(u, e, b.sm_count, &grid, st) : launch_umma<128>(u, e, b.sm_count, &grid, st);
        tock();
      }
    } else if (!need_w && src_aligned && pick_gather_cw(a.g.n_cols, r.H) > 0) {
      tick(NDCN_K_STAGE);
      rc = launch_gather(b, a, r.H, e, &grid, st);
      tock();
    } else {
      tick(NDCN_K_STAGE);
      switch (r.H) {
        case 256: rc = launch_ndcn_fast<4, 2>(a, e, &grid, st); break;
        case 128: rc = launch_ndcn_fast<4, 1>(a, e, &grid, st); break;
        case 64: rc = launch_ndcn_fast<2, 1>(a, e, &grid, st); break;
        case 32: rc = launch_ndcn_fast<1, 1>(a, e, &grid, st); break;
        default: {
          grid = (int)((a.g.n_rows + kWarpsPerCta - 1) / kWarpsPerCta);
          const size_t smem = sizeof(float) * kWarpsPerCta * r.H;
          k_stage_ndcn_any<<<grid, kStageThreads, smem, st>>>(a, r.H, r.W, e);
          rc = (int)cudaGetLastError();
        }
      }
      tock();
    }
  } else if (r.kind == NDCN_RHS_HEAT || r.kind == NDCN_RHS_GENE || r.kind == NDCN_RHS_MUTUAL) {
    DynArgs a;
    a.g = b.g->v;
    a.x = src;
    a.kind = r.kind;
    a.d = r.H;
    for (int i = 0; i < 8; ++i) a.p[i] = r.p[i];
    a.long_rows = b.g->long_rows;
    a.n_long = r.H == 1 ? b.g->n_long : 0;
    const double avg = a.g.n_rows > 0 ? (double)a.g.nnz / (double)a.g.n_rows : 0.0;
    tick(NDCN_K_STAGE);
    if (r.kind == NDCN_RHS_HEAT) rc = launch_dyn<NDCN_RHS_HEAT>(a, e, avg, &grid, st);
    else if (r.kind == NDCN_RHS_GENE) rc = launch_dyn<NDCN_RHS_GENE>(a, e, avg, &grid, st);
    else rc = launch_dyn<NDCN_RHS_MUTUAL>(a, e, avg, &grid, st);
    tock();
  } else {
    return NDCN_E_ARG;
  }
  if (rc != 0) return rc;
  if (e.mode == EPI_ERR && grid > b.max_partials) return NDCN_E_WORKSPACE;
  if (n_partials) *n_partials = grid;
  return 0;
}

static int max_partials_for(const ndcn_graph* g, int H) {
  // upper bound over every stage kernel's grid: the [N,1] kernels with 4 lanes per row, the
  // warp-per-row kernels, and the chunk-major gather (H/cw chunks x (row blocks + long rows))
  const int64_t n_rows = g->v.n_rows;
  int64_t by_rows = (n_rows + (kStageThreads / 32) - 1) / (kStageThreads / 32) + g->n_long;
  int64_t by_lpr4 = (n_rows + 63) / 64 + g->n_long;
  // grid_for_elems' cap; row-chunked tcgen05 launches write one partial per CTA and chunk
  int64_t best = std::max<int64_t>(std::max(by_rows, by_lpr4), (int64_t)sm_count_now() * 64);
  if (H % 16 == 0) {
    const int cws[3] = {16, 32, 64};
    for (int cw : cws)
      if (H % cw == 0) best = std::max<int64_t>(best, gather_grid(g, H, cw));
  }
  return (int)best + 8;
}

// ---------------------------------------------------------------------------------------
// graph handle
// ---------------------------------------------------------------------------------------
extern "C" int ndcn_graph_create(int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* rowptr,
                                 const int32_t* col, const float* val, ndcn_graph_t** out) {
  if (!out || n_rows < 0 || n_cols < n_rows || nnz < 0 || !rowptr) return NDCN_E_ARG;
  if (nnz > 0 && (!col || !val)) return NDCN_E_ARG;
  if (nnz > 0x7fffffffLL || n_cols > 0x7fffffffLL) return NDCN_E_ARG;  // int32 CSR
  ndcn_graph* g = new (std::nothrow) ndcn_graph();
  if (!g) return NDCN_E_ARG;
  g->v.rowptr = rowptr;
  g->v.col = col;
  g->v.val = val;
  g->v.n_rows = n_rows;
  g->v.n_cols = n_cols;
  g->v.nnz = nnz;
  // rows above kLongRow entries get their own CTAs in the chunk-major gather: list them once
  if (n_rows > 0) {
    std::vector<int32_t> rp((size_t)n_rows + 1);
    cudaError_t ce = cudaMemcpy(rp.data(), rowptr, sizeof(int32_t) * (n_rows + 1), cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) {
      delete g;
      return (int)ce;
    }
    std::vector<int32_t> longs;
    for (int64_t r = 0; r < n_rows; ++r) {
      const int deg = rp[r + 1] - rp[r];
      if (deg < 0 || rp[r + 1] > nnz) {
        delete g;
        return NDCN_E_ARG;
      }
      g->max_deg = std::max(g->max_deg, deg);
      if (deg > kLongRow) longs.push_back((int32_t)r);
    }
    g->n_long = (int)longs.size();
    g->long_host = longs;
    if (g->n_long > 0) {
      ce = cudaMalloc((void**)&g->long_rows, sizeof(int32_t) * longs.size());
      if (ce == cudaSuccess)
        ce = cudaMemcpy(g->long_rows, longs.data(), sizeof(int32_t) * longs.size(), cudaMemcpyHostToDevice);
      if (ce != cudaSuccess) {
        if (g->long_rows) cudaFree(g->long_rows);
        delete g;
        return (int)ce;
      }
    }
  }
  *out = g;
  return NDCN_OK;
}

extern "C" int ndcn_graph_destroy(ndcn_graph_t* g) {
  if (g && g->long_rows) cudaFree(g->long_rows);

  delete g;
  return NDCN_OK;
}

// ---------------------------------------------------------------------------------------
// stand-alone operators
// ---------------------------------------------------------------------------------------
static EpiArgs store_only(float* out) {
  EpiArgs e;
  std::memset(&e, 0, sizeof(e));
  e.mode = EPI_STORE;
  e.k_out = pp(out);
  return e;
}

template <int H>
static void prep_w_image(const float* W, float* img, cudaStream_t st) {
  k_prep_w_image<H><<<(H * H + 255) / 256, 256, 0, st>>>(W, img);
}

// layout of ndcn_rhs_desc_t::prepared
struct PreparedView {
  const float* Wt = nullptr;       // [H,H] = W^T
  const float* img_fwd = nullptr;  // tf32 hi/lo image of W    (H in {128, 256})
  const float* img_bwd = nullptr;  // tf32 hi/lo image of W^T  (H in {128, 256})
  const float* zeros = nullptr;    // [H]
};
static size_t prepared_off_img(int H) { return align_up(sizeof(float) * (size_t)H * H, 1024); }
static size_t prepared_img_bytes(int H) { return (H == 256 || H == 128) ? align_up(sizeof(float) * 2 * (size_t)H * H, 1024) : 0; }
static PreparedView prepared_view(const void* p, int H) {
  PreparedView v;
  if (!p) return v;
  const unsigned char* b = (const unsigned char*)p;
  v.Wt = (const float*)b;
  const size_t img = prepared_img_bytes(H);
  if (img) {
    v.img_fwd = (const float*)(b + prepared_off_img(H));
    v.img_bwd = (const float*)(b + prepared_off_img(H) + img);
  }
  v.zeros = (const float*)(b + prepared_off_img(H) + 2 * img);
  return v;
}

extern "C" size_t ndcn_prepared_weights_bytes(int32_t H) {
  if (H < 1) return 0;
  return prepared_off_img(H) + 2 * prepared_img_bytes(H) + align_up(sizeof(float) * (size_t)H, 1024);
}

extern "C" int ndcn_prepare_weights_f32(const float* W, int32_t H, void* prepared, ndcn_stream_t s) {
  if (!W || !prepared || H < 1 || H > 1024 || ((uintptr_t)prepared & 1023u)) return NDCN_E_ARG;
  cudaStream_t st = (cudaStream_t)s;
  PreparedView v = prepared_view(prepared, H);
  float* Wt = const_cast<float*>(v.Wt);
  k_transpose<<<(H * H + 255) / 256, 256, 0, st>>>(W, Wt, H);
  CU_TRY(cudaMemsetAsync(const_cast<float*>(v.zeros), 0, sizeof(float) * H, st));
  if (H == 256) {
    prep_w_image<256>(W, const_cast<float*>(v.img_fwd), st);
    prep_w_image<256>(Wt, const_cast<float*>(v.img_bwd), st);
  } else if (H == 128) {
    prep_w_image<128>(W, const_cast<float*>(v.img_fwd), st);
    prep_w_image<128>(Wt, const_cast<float*>(v.img_bwd), st);
  }
  return (int)cudaGetLastError();
}

extern "C" int ndcn_weight_grads_f32(const float* gp, const float* z, int64_t n, int32_t H, float* dW, float* db,
                                     int32_t accumulate, ndcn_stream_t s) {
  if (!gp || !z || !dW || n < 0 || H < 1 || H > 1024) return NDCN_E_ARG;
  cudaStream_t st = (cudaStream_t)s;
  keep_scratch_pooled();
  const int64_t hh = (int64_t)H * H;
  if (n == 0) {
    if (!accumulate) {
      CU_TRY(cudaMemsetAsync(dW, 0, sizeof(float) * hh, st));
      if (db) CU_TRY(cudaMemsetAsync(db, 0, sizeof(float) * H, st));
    }
    return NDCN_OK;
  }
  // tensor-core tiles (128 x 128, mma.sync 3xTF32) for the wide layers, 64 x 64 FP32-FMA tiles otherwise
  static const bool mma_on = [] { const char* e = std::getenv("NDCN_WG_MMA"); return !(e && e[0] == '0'); }();
  const bool use_mma = mma_on && H % kWgmTile == 0 && n >= 1024 && aligned16(gp) && aligned16(z);
  const int tile = use_mma ? kWgmTile : kWgTile, kstep = use_mma ? kWgmK : kWgK;
  const int tiles = (H + tile - 1) / tile;
  // enough row chunks to fill the chip twice, at least 64 rows each
  int64_t nz = std::max<int64_t>(1, (2 * (int64_t)sm_count_now() + tiles * tiles - 1) / (tiles * tiles));
  nz = std::min<int64_t>(nz, (n + 63) / 64);
  int64_t rows_per_chunk = (n + nz - 1) / nz;
  rows_per_chunk = (rows_per_chunk + kstep - 1) / kstep * kstep;
  nz = (n + rows_per_chunk - 1) / rows_per_chunk;
  if (use_mma) {
    static PerDeviceOnce attr;
    if (attr.need()) {
      CU_TRY(cudaFuncSetAttribute(k_weight_grads_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgmSmemBytes));
      attr.mark();
    }
  }
  float* part = nullptr;
  CU_TRY(cudaMallocAsync((void**)&part, sizeof(float) * (size_t)nz * (hh + H), st));
  float* part_b = db ? part + (size_t)nz * hh : nullptr;
  if (use_mma) {
    k_weight_grads_mma<<<dim3(tiles, tiles, (unsigned)nz), kWgmThreads, kWgmSmemBytes, st>>>(gp, z, n, H, rows_per_chunk, part,
                                                                                   part_b);
  } else {
    k_weight_grads_partial<<<dim3(tiles, tiles, (unsigned)nz), 256, 0, st>>>(gp, z, n, H, rows_per_chunk, part, part_b);
  }
  k_weight_grads_reduce<<<(unsigned)((hh + 255) / 256), 256, 0, st>>>(part, (int)nz, hh, dW, accumulate ? 1 : 0);
  if (db) k_weight_grads_reduce<<<(H + 255) / 256, 256, 0, st>>>(part_b, (int)nz, H, db, accumulate ? 1 : 0);
  cudaFreeAsync(part, st);
  return (int)cudaGetLastError();
}

extern "C" int ndcn_rhs_eval_f32(const ndcn_graph_t* g, const ndcn_rhs_desc_t* rhs, const float* x, float* out,
                                 ndcn_stream_t s) {
  if (!g || !rhs || !x || !out) return NDCN_E_ARG;
  keep_scratch_pooled();
  cudaStream_t st = (cudaStream_t)s;
  float* Wt = nullptr;
  float* Z = nullptr;
  float* Wimg = nullptr;
  const bool use_umma = umma_eligible(*rhs, g->v.n_rows) && aligned16(x) && aligned16(out);
  const bool need_wt = rhs->kind == NDCN_RHS_NDCN && !(rhs->flags & NDCN_F_NO_CONTROL) && fast_width(rhs->H) && !use_umma;
  if (need_wt || use_umma) {
    if (!rhs->W) return NDCN_E_ARG;
  }
  const PreparedView pv = prepared_view(rhs->kind == NDCN_RHS_NDCN ? rhs->prepared : nullptr, rhs->H);
  const float* Wt_use = pv.Wt;
  const float* Wimg_use = pv.img_fwd;
  if (need_wt && !Wt_use) {
    CU_TRY(cudaMallocAsync((void**)&Wt, sizeof(float) * rhs->H * rhs->H, st));
    const int n = rhs->H * rhs->H;
    k_transpose<<<(n + 255) / 256, 256, 0, st>>>(rhs->W, Wt, rhs->H);
    Wt_use = Wt;
  }
  if (use_umma) {
    if (!(rhs->flags & NDCN_F_NO_GRAPH))
      CU_TRY(cudaMallocAsync((void**)&Z, sizeof(float) * (size_t)g->v.n_rows * rhs->H, st));
    if (!Wimg_use) {
      CU_TRY(cudaMallocAsync((void**)&Wimg, sizeof(float) * 2 * rhs->H * rhs->H, st));
      if (rhs->H == 256) prep_w_image<256>(rhs->W, Wimg, st);
      else prep_w_image<128>(rhs->W, Wimg, st);
      Wimg_use = Wimg;
    }
  }
  RhsBinding b{g, rhs, Wt_use, nullptr, 0, sm_count_now()};
  // without Z the dispatcher falls back to the SIMT kernels; no_graph needs no Z: give it a non-null tag
  b.Z = use_umma ? (Z ? Z : out) : nullptr;
  b.Wimg = Wimg_use;
  int rc = launch_stage(b, pp(const_cast<float*>(x)), store_only(out), nullptr, st);
  if (Wt) cudaFreeAsync(Wt, st);
  if (Z) cudaFreeAsync(Z, st);
  if (Wimg) cudaFreeAsync(Wimg, st);
  return rc;
}

// ---------------------------------------------------------------------------------------
// vjp of one RHS evaluation: what autograd computes for
// ODEFunc.forward (neural_dynamics.py:20-39) given the cotangent of its output, on the same kernels
// as the forward pass -- gather, (tcgen05 or FP32-FMA) GEMM with a ReLU-mask epilogue, GEMM with W
// untransposed, gather with Phi^T fused with the accumulation into the running adjoint.
// ---------------------------------------------------------------------------------------
extern "C" int ndcn_rhs_vjp_f32(const ndcn_graph_t* g, const ndcn_graph_t* g_t, const ndcn_rhs_desc_t* rhs,
                                const float* x, const float* gk, float scale, float* gx, int32_t accumulate,
                                float* gp, float* z, ndcn_stream_t s) {
  if (!g || !g_t || !rhs || !x || !gk || !gx || !gp) return NDCN_E_ARG;
  if (rhs->kind != NDCN_RHS_NDCN || rhs->H < 1 || rhs->H > 1024) return NDCN_E_ARG;
  keep_scratch_pooled();
  const int H = rhs->H;
  const int64_t n = g->v.n_rows;
  if (g->v.n_cols != n || g_t->v.n_rows != n || g_t->v.n_cols != n) return NDCN_E_ARG;  // single-GPU graphs
  const bool no_graph = (rhs->flags & NDCN_F_NO_GRAPH) != 0, no_control = (rhs->flags & NDCN_F_NO_CONTROL) != 0;
  if (!no_graph && !z) return NDCN_E_ARG;
  if (!no_control && (!rhs->W || !rhs->b)) return NDCN_E_ARG;
  if (n == 0) return NDCN_OK;
  cudaStream_t st = (cudaStream_t)s;
  const int sm_count = sm_count_now();
  const int64_t numel = n * H;
  const bool vec = aligned16(x) && aligned16(gk) && aligned16(gx) && aligned16(gp) && (!z || aligned16(z)) && H % 4 == 0;
  auto elementwise = [&](const float* k_in, EpiArgs e) {
    k_epi_only<<<grid_for_elems(numel, sm_count), kStageThreads, 0, st>>>(pp(const_cast<float*>(k_in)), numel, e, vec ? 1 : 0);
    return (int)cudaGetLastError();
  };
  auto blank = [] {
    EpiArgs e;
    std::memset(&e, 0, sizeof(e));
    e.dt_src = DT_HOST;
    return e;
  };
  // (a) z = Phi x
  const float* zin = x;
  if (!no_graph) {
    ndcn_rhs_desc_t rg = *rhs;
    rg.flags = NDCN_F_NO_CONTROL | NDCN_F_NO_RELU;
    RhsBinding bg{g, &rg, nullptr, nullptr, 0, sm_count};
    RC_TRY(launch_stage(bg, pp(const_cast<float*>(x)), store_only(z), nullptr, st));
    zin = z;
  }
  // (b) gp = scale * gk where relu'(.) = 1
  EpiArgs em = blank();
  em.mode = EPI_MASK;
  em.dt_host = scale;
  em.beta[0] = 1.0f;
  em.y0 = pp(const_cast<float*>(gk));
  em.y_out = pp(gp);
  float* u = gp;          // no_control: the masked cotangent is what goes through Phi^T
  float* scratch = nullptr;  // [n,H] u | [H,H] W^T | [H] zeros | two W images
  if (no_control) {
    RC_TRY(elementwise(zin, em));
  } else {
    const bool use_umma = umma_eligible(*rhs, n) && vec;
    const PreparedView pv = prepared_view(rhs->prepared, H);
    const bool have_prep = pv.Wt != nullptr && (!use_umma || pv.img_fwd != nullptr);
    const size_t img = (use_umma && !have_prep) ? align_up(sizeof(float) * 2 * (size_t)H * H, 1024) : 0;
    const size_t bytes = align_up(sizeof(float) * (size_t)numel, 1024) +
                         (have_prep ? 0 : align_up(sizeof(float) * (size_t)H * H, 1024) + align_up(sizeof(float) * (size_t)H, 1024)) +
                         2 * img + 1024;
    CU_TRY(cudaMallocAsync((void**)&scratch, bytes, st));
    unsigned char* p = (unsigned char*)align_up((size_t)scratch, 1024);
    u = (float*)p;
    p += align_up(sizeof(float) * (size_t)numel, 1024);
    const float* Wt = pv.Wt;
    const float* zeros = pv.zeros;
    const float* img_fwd = use_umma ? pv.img_fwd : nullptr;
    const float* img_bwd = use_umma ? pv.img_bwd : nullptr;
    if (!have_prep) {
      float* Wt_w = (float*)p;
      p += align_up(sizeof(float) * (size_t)H * H, 1024);
      float* zeros_w = (float*)p;
      p += align_up(sizeof(float) * (size_t)H, 1024);
      float* img_fwd_w = use_umma ? (float*)p : nullptr;
      float* img_bwd_w = use_umma ? (float*)(p + img) : nullptr;
      k_transpose<<<(H * H + 255) / 256, 256, 0, st>>>(rhs->W, Wt_w, H);
      CU_TRY(cudaMemsetAsync(zeros_w, 0, sizeof(float) * H, st));
      if (use_umma) {
        if (H == 256) {
          prep_w_image<256>(rhs->W, img_fwd_w, st);
          prep_w_image<256>(Wt_w, img_bwd_w, st);
        } else {
          prep_w_image<128>(rhs->W, img_fwd_w, st);
          prep_w_image<128>(Wt_w, img_bwd_w, st);
        }
      }
      Wt = Wt_w; zeros = zeros_w; img_fwd = img_fwd_w; img_bwd = img_bwd_w;
    }
    // pre-activation mask: the forward GEMM on z with the mask epilogue
    ndcn_rhs_desc_t rf = *rhs;
    rf.flags = NDCN_F_NO_GRAPH;
    RhsBinding bf{g, &rf, Wt, nullptr, 0, sm_count};
    bf.Z = use_umma ? u : nullptr;  // non-null tag: no_graph needs no Z
    bf.Wimg = img_fwd;
    int rc = launch_stage(bf, pp(const_cast<float*>(zin)), em, nullptr, st);
    // u = gp W: the same GEMM kernels with W^T in the role of W, no bias, no ReLU
    if (rc == 0) {
      ndcn_rhs_desc_t rb = *rhs;
      rb.flags = NDCN_F_NO_GRAPH | NDCN_F_NO_RELU;
      rb.W = Wt;
      rb.b = zeros;
      RhsBinding bb{g, &rb, rhs->W /* (W^T)^T */, nullptr, 0, sm_count};
      bb.Z = use_umma ? u : nullptr;
      bb.Wimg = img_bwd;
      rc = launch_stage(bb, pp(gp), store_only(u), nullptr, st);
    }
    if (rc != 0) {
      cudaFreeAsync(scratch, st);
      return rc;
    }
  }
  // (c) gx (+)= Phi^T u
  EpiArgs eo = accumulate ? blank() : store_only(gx);
  if (accumulate) {
    eo.mode = EPI_LINCOMB;
    eo.n_prev = 0;
    eo.dt_host = 1.0f;
    eo.beta[0] = 1.0f;
    eo.y0 = pp(gx);
    eo.y_out = pp(gx);
  }
  int rc = 0;
  if (no_graph) {
    rc = elementwise(u, eo);
  } else {
    ndcn_rhs_desc_t rt = *rhs;
    rt.flags = NDCN_F_NO_CONTROL | NDCN_F_NO_RELU;
    RhsBinding bt{g_t, &rt, nullptr, nullptr, 0, sm_count};
    rc = launch_stage(bt, pp(u), eo, nullptr, st);
  }
  if (scratch) cudaFreeAsync(scratch, st);
  return rc;
}

// ---------------------------------------------------------------------------------------
// discrete adjoint of a fixed-grid solve as ONE cooperative launch (narrow widths, small graphs)
// ---------------------------------------------------------------------------------------
extern "C" int ndcn_fixed_grid_adjoint_small_f32(const ndcn_graph_t* g, const ndcn_graph_t* g_t,
                                                 const ndcn_rhs_desc_t* rhs, int32_t method, const double* t_host,
                                                 int32_t n_t, const float* slab, const float* g_slab, float* lam,
                                                 float* dW, float* db, ndcn_stream_t s) {
  if (!g || !g_t || !rhs || !t_host || !slab || !g_slab || !lam || n_t < 2) return NDCN_E_ARG;
  if (rhs->kind != NDCN_RHS_NDCN || rhs->H < 1 || rhs->H > 32) return NDCN_E_ARG;
  if (method != NDCN_EULER && method != NDCN_MIDPOINT && method != NDCN_RK4) return NDCN_E_METHOD;
  const bool no_control = (rhs->flags & NDCN_F_NO_CONTROL) != 0;
  if (!no_control && (!rhs->W || !rhs->b || !dW || !db)) return NDCN_E_ARG;
  const int64_t n = g->v.n_rows;
  if (n < 1 || g->v.n_cols != n || g_t->v.n_rows != n || g_t->v.n_cols != n) return NDCN_E_ARG;
  if (n > cfg().small_max_rows) return NDCN_E_ARG;
  int coop = 0;
  if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, current_device()) != cudaSuccess || !coop) return NDCN_E_ARG;
  keep_scratch_pooled();
  cudaStream_t st = (cudaStream_t)s;
  const int H = rhs->H;
  const int64_t numel = n * H;
  const int sm_count = sm_count_now();

  const size_t smem = sizeof(float) * kWarpsPerCta * 64;
  int per_sm = 0;
  CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_adjoint_small, kStageThreads, smem));
  if (per_sm < 1) return NDCN_E_ARG;
  const int by_rows = (int)((n + kWarpsPerCta - 1) / kWarpsPerCta);
  const int grid = std::max(1, std::min(by_rows, std::min(per_sm * sm_count, 2 * sm_count)));
  const int64_t n_warps = (int64_t)grid * kWarpsPerCta;

  // scratch: step sizes | U | up to 8 state-sized buffers | per-warp parameter-gradient partials
  const size_t state_b = align_up(sizeof(float) * (size_t)numel, 256);
  const int n_states = method == NDCN_RK4 ? 8 : (method == NDCN_MIDPOINT ? 2 : 0);
  const size_t dts_b = align_up(sizeof(float) * (size_t)n_t, 256);
  const size_t part_b = no_control ? 0 : align_up(sizeof(float) * (size_t)n_warps * (H * H + H), 256);
  unsigned char* scratch = nullptr;
  CU_TRY(cudaMallocAsync((void**)&scratch, dts_b + (1 + n_states) * state_b + part_b + 256, st));
  unsigned char* p = (unsigned char*)align_up((size_t)scratch, 256);
  float* dts_dev = (float*)p;
  p += dts_b;
  std::vector<float> dts((size_t)n_t - 1);
  for (int i = 0; i + 1 < n_t; ++i) dts[i] = (float)t_host[i + 1] - (float)t_host[i];
  cudaError_t ce = cudaMemcpyAsync(dts_dev, dts.data(), sizeof(float) * dts.size(), cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);  // `dts` is a host temporary
  if (ce != cudaSuccess) {
    cudaFreeAsync(scratch, st);
    return (int)ce;
  }
  AdjArgs a;
  std::memset(&a, 0, sizeof(a));
  a.g = g->v;
  a.gt = g_t->v;
  a.H = H;
  a.n_rows = (int)n;
  a.n_t = n_t;
  a.method = method;
  a.flags = rhs->flags;
  a.W = rhs->W;
  a.bias = rhs->b;
  a.dts = dts_dev;
  a.slab = slab;
  a.g_slab = g_slab;
  a.lam = lam;
  a.U = (float*)p;
  p += state_b;
  for (int i = 0; i < n_states; ++i) {
    a.S[i] = (float*)p;
    p += state_b;
  }
  a.part = no_control ? nullptr : (float*)p;
  a.dW = dW;
  a.db = db;
  a.numel = numel;
  SmallArgs fw;  // the forward right-hand side, for the stage recomputation
  std::memset(&fw, 0, sizeof(fw));
  fw.g = g->v;
  fw.kind = NDCN_RHS_NDCN;
  fw.H = H;
  fw.flags = rhs->flags;
  fw.W = rhs->W;
  fw.bias = rhs->b;
  fw.numel = numel;
  fw.n_rows = (int)n;
  void* params[] = {&a, &fw};
  ce = cudaLaunchCooperativeKernel((void*)k_adjoint_small, dim3(grid), dim3(kStageThreads), params, smem, st);
  cudaFreeAsync(scratch, st);
  return (int)ce;
}

extern "C" int ndcn_spmm_f32(const ndcn_graph_t* g, const float* x, float* y, int32_t H, ndcn_stream_t s) {
  ndcn_rhs_desc_t r;
  std::memset(&r, 0, sizeof(r));
  r.kind = NDCN_RHS_NDCN;
  r.flags = NDCN_F_NO_CONTROL | NDCN_F_NO_RELU;
  r.H = H;
  return ndcn_rhs_eval_f32(g, &r, x, y, s);
}

// ---------------------------------------------------------------------------------------
// solver object
// ---------------------------------------------------------------------------------------
static size_t state_bytes(int64_t rows, int H) { return align_up((size_t)rows * H * sizeof(float), 256); }

extern "C" size_t ndcn_solver_workspace_bytes(int64_t n_rows, int64_t n_cols, int32_t H, int32_t method) {
  (void)method;
  if (n_cols < n_rows) n_cols = n_rows;
  size_t b = 0;
  b += 4 * state_bytes(n_cols, H);                      // Y[2], YS[2]
  b += 7 * state_bytes(n_rows, H);                      // KF[2], K[5]
  b += align_up(sizeof(float) * (size_t)H * H, 256);    // W^T
  if (H == 256 || H == 128) {                           // tcgen05 path: Z = Phi x, W image
    b += state_bytes(n_rows, H);
    b += align_up(sizeof(float) * 2 * (size_t)H * H, 1024);
  }
  return b + 1280;
}

extern "C" int ndcn_solver_create(const ndcn_graph_t* g, const ndcn_rhs_desc_t* rhs, int32_t method,
                                  void* workspace, size_t workspace_bytes, ndcn_solver_t** out) {
  if (!g || !rhs || !out || !workspace) return NDCN_E_ARG;
  if (method < NDCN_EULER || method > NDCN_DOPRI5) return NDCN_E_METHOD;
  if (rhs->H < 1) return NDCN_E_ARG;
  const int64_t n_rows = g->v.n_rows, n_cols = g->v.n_cols;
  if (workspace_bytes < ndcn_solver_workspace_bytes(n_rows, n_cols, rhs->H, method)) return NDCN_E_WORKSPACE;
  ndcn_solver* sv = new (std::nothrow) ndcn_solver();
  if (!sv) return NDCN_E_ARG;
  sv->g = g;
  sv->rhs = *rhs;
  sv->method = method;
  sv->n_rows = n_rows;
  sv->n_cols = n_cols;
  sv->H = rhs->H;
  sv->numel = n_rows * rhs->H;
  sv->numel_src = n_cols * rhs->H;
  unsigned char* p = (unsigned char*)align_up((size_t)workspace, 256);
  auto take = [&](size_t bytes) {
    float* r = (float*)p;
    p += bytes;
    return r;
  };
  for (int i = 0; i < 2; ++i) sv->Y[i] = take(state_bytes(n_cols, sv->H));
  for (int i = 0; i < 2; ++i) sv->YS[i] = take(state_bytes(n_cols, sv->H));
  for (int i = 0; i < 2; ++i) sv->KF[i] = take(state_bytes(n_rows, sv->H));
  for (int i = 0; i < 5; ++i) sv->K[i] = take(state_bytes(n_rows, sv->H));
  sv->Wt = take(align_up(sizeof(float) * (size_t)sv->H * sv->H, 256));
  if (sv->H == 256 || sv->H == 128) {
    sv->Z = take(state_bytes(n_rows, sv->H));
    p = (unsigned char*)align_up((size_t)p, 1024);
    sv->Wimg = take(align_up(sizeof(float) * 2 * (size_t)sv->H * sv->H, 1024));
  }
  sv->sm_count = sm_count_now();
  sv->max_partials = max_partials_for(g, sv->H);
  int rc = (int)cudaMalloc((void**)&sv->partials, sizeof(double) * 2 * sv->max_partials);
  if (!rc) rc = (int)cudaMalloc((void**)&sv->xchg, sizeof(double) * 4);
  if (!rc) rc = (int)cudaMalloc((void**)&sv->t_stage, sizeof(float) * 4);
  sv->t_cap = 128;  // requested output times; grown on demand (never inside a peer-push solve's barriers)
  if (!rc) rc = (int)cudaMalloc((void**)&sv->t_out, sizeof(double) * sv->t_cap);
  if (!rc) rc = (int)cudaMalloc((void**)&sv->ctrl, sizeof(Ctrl));
  if (!rc) rc = (int)cudaMallocHost((void**)&sv->ctrl_host, sizeof(Ctrl));
  if (rc) {
    ndcn_solver_destroy(sv);
    return rc;
  }
  *out = sv;
  return NDCN_OK;
}

extern "C" int ndcn_solver_destroy(ndcn_solver_t* sv) {
  if (!sv) return NDCN_OK;
  if (sv->partials) cudaFree(sv->partials);
  if (sv->xchg) cudaFree(sv->xchg);
  if (sv->t_stage) cudaFree(sv->t_stage);
  if (sv->t_out) cudaFree(sv->t_out);
  if (sv->ctrl) cudaFree(sv->ctrl);
  if (sv->feat_dev) cudaFree(sv->feat_dev);
  if (sv->ctrl_host) cudaFreeHost(sv->ctrl_host);
  delete sv;
  return NDCN_OK;
}

// ---------------------------------------------------------------------------------------
// drivers
// ---------------------------------------------------------------------------------------
namespace {

struct Driver : StageTimer {
  ndcn_solver* sv;
  const ndcn_solve_opts_t* o;
  cudaStream_t st;
  RhsBinding bind;
  bool vec_ok = false;  // all elementwise buffers are 16-byte aligned and numel % 4 == 0 handled
  int64_t nfe = 0;

  Driver(ndcn_solver* sv_, const ndcn_solve_opts_t* o_, cudaStream_t st_) : sv(sv_), o(o_), st(st_) {
    bind = RhsBinding{sv->g, &sv->rhs, sv->Wt, sv->partials, sv->max_partials, sv->sm_count};
    bind.Z = sv->Z;
    bind.Wimg = sv->Wimg;
    bind.timer = this;
    bind.launches = &sv->launches;
  }

  // ---- optional per-class kernel timing (NDCN_O_TIME_KERNELS): CUDA events on the launch stream
  struct Ev { cudaEvent_t a, b; int cls; int64_t attempt; bool count; };
  bool timing = false;
  int64_t cur_attempt = -1;  // dopri5 attempt index the next launches belong to (-1: prologue)
  std::vector<Ev> evs;
  void begin(int cls, bool count) override { t_begin(cls, count); }
  void end() override { t_end(); }
  void t_begin(int cls, bool count = true) {
    if (!timing) return;
    Ev e{nullptr, nullptr, cls, cur_attempt, count};
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    cudaEventRecord(e.a, st);
    evs.push_back(e);
  }
  void t_end() {
    if (timing) cudaEventRecord(evs.back().b, st);
  }
  // call after the stream is synchronised; launches of attempts >= n_real were device-side no-ops
  void t_collect(int64_t n_real, ndcn_solve_stats_t* stats) {
    for (Ev& e : evs) {
      float ms = 0.f;
      if (stats && (e.attempt < 0 || e.attempt < n_real) && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
        stats->class_ms[e.cls] += (double)ms;
        if (e.count) stats->class_launches[e.cls] += 1;
      }
      cudaEventDestroy(e.a);
      cudaEventDestroy(e.b);
    }
    evs.clear();
  }

  bool feat() const { return sv->feat_on; }
  bool push() const { return sv->n_push > 0 || sv->feat_on; }  // peers reached through mapped memory + device barriers
  bool multi() const { return o->exchange != nullptr || push(); }

  void fill_feat(EpiArgs& e, int mode) const {
    e.feat = sv->feat_dev;
    e.feat_mode = mode;
    e.feat_rank = sv->feat_rank;
    e.feat_row0 = sv->feat_bounds[sv->feat_rank];
    e.feat_hc_log2 = ilog2(sv->feat_hc);
    e.feat_h_log2 = ilog2(sv->H);
  }

  // every y_out of a solve is a gather source: on a peer-push solve it is also stored at the peers
  // (whole rows at a byte delta, or -- feature-sharded -- column slice by column slice)
  EpiArgs blank() const {
    EpiArgs e;
    std::memset(&e, 0, sizeof(e));
    e.n_peers = sv->n_push;
    for (int j = 0; j < sv->n_push; ++j) e.peer_delta[j] = sv->push_delta[j];
    if (sv->feat_on) fill_feat(e, FEAT_Y_SLICES);
    return e;
  }

  // peer push: all ranks' stores into each other's gather sources are complete and visible after this;
  // with a payload also the all-reduce (SUM) of two doubles in place
  int peer_barrier(double* payload) {
    sv->launches += 1;
    t_begin(NDCN_K_EXCHANGE);
    k_peer_barrier<<<1, 32, 0, st>>>(sv->peers, payload, sv->ctrl);
    t_end();
    return (int)cudaGetLastError();
  }

  // all-reduce (SUM) of the 2 doubles at sv->xchg across ranks: device barrier kernel or the host hook
  int allreduce_xchg() {
    if (push()) return peer_barrier(sv->xchg);
    return o->exchange(o->exchange_user, 1, sv->xchg);
  }

  // initial state -> own rows of gather source `dst` (and, peer push, the same rows at every peer)
  int put_y0(float* dst, const float* y0) {
    if (!push()) {
      CU_TRY(cudaMemcpyAsync(dst, y0, sizeof(float) * (size_t)sv->numel, cudaMemcpyDeviceToDevice, st));
      return 0;
    }
    if (feat()) {
      sv->launches += 1;
      k_copy_slices<<<grid_for_elems(sv->numel, sv->sm_count), kStageThreads, 0, st>>>(
          y0, dst, sv->numel, sv->feat_dev, sv->feat_bounds[sv->feat_rank], ilog2(sv->H), ilog2(sv->feat_hc));
      return (int)cudaGetLastError();
    }
    PushDeltas pd;
    pd.n = sv->n_push;
    for (int j = 0; j < kMaxPeers; ++j) pd.delta[j] = j < sv->n_push ? sv->push_delta[j] : 0;
    sv->launches += 1;
    k_copy_push<<<grid_for_elems(sv->numel, sv->sm_count), kStageThreads, 0, st>>>(y0, dst, sv->numel, pd, vec_ok ? 1 : 0);
    return (int)cudaGetLastError();
  }

  int exchange(float* buf) {  // make the halo rows of a gather source valid (multi-GPU)
    if (feat()) return 0;  // stage() brackets the slice gather with its own barriers
    if (push()) return peer_barrier(nullptr);
    if (o->exchange && sv->n_cols > sv->n_rows) return o->exchange(o->exchange_user, 0, buf);
    return 0;
  }

  // time of the next RHS evaluation, for a callback RHS (k_stage_time); the built-in right-hand sides are autonomous
  int tm_mode = 0, tm_from_ctrl = 0;
  float tm_base = 0.f, tm_dt = 0.f, tm_alpha = 0.f, tm_num = 0.f, tm_den = 1.f;
  void time_host(float t0, float dt, int mode, float a = 0.f, float num = 0.f, float den = 1.f) {
    tm_from_ctrl = 0; tm_base = t0; tm_dt = dt; tm_mode = mode; tm_alpha = a; tm_num = num; tm_den = den;
  }
  void time_ctrl(int mode, float a = 0.f, float base = 0.f) {
    tm_from_ctrl = mode == 3 ? 0 : 1; tm_mode = mode; tm_alpha = a; tm_base = base; tm_dt = 0.f; tm_num = 0.f; tm_den = 1.f;
  }

  // one RHS evaluation fused with epilogue e. `src_host` is the buffer the host knows to be the
  // source (needed for the exchange/callback hooks; the kernels themselves select by parity).
  int stage(PtrPair src, float* src_host, EpiArgs e, float* k_host, int* n_partials = nullptr) {
    if (o->gather_mode == NDCN_GATHER_EXTERNAL) {
      // feature-sharded multi-GPU gather: the hook turns this rank's rows of the gather source into
      // this rank's rows of z = Phi x (two all-to-alls around a column-slice gather of the whole graph)
      if (!o->exchange || sv->rhs.kind != NDCN_RHS_NDCN || (sv->rhs.flags & (NDCN_F_NO_GRAPH | NDCN_F_NO_CONTROL)) ||
          !sv->Z || o->z_block_cols < 32 || sv->H % o->z_block_cols != 0)
        return NDCN_E_ARG;
      ndcn_gather_request_t req{src_host, sv->Z};
      nfe += 1;
      t_begin(NDCN_K_GATHER);
      const int rc = o->exchange(o->exchange_user, 2, &req);
      t_end();
      if (rc != 0) return rc;
      RhsBinding b2 = bind;
      b2.z_block_cols = o->z_block_cols;
      return launch_stage(b2, src, e, n_partials, st);
    }
    if (feat()) {
      // feature-sharded peer push: the producers of `src` have scattered it into every rank's slice buffer;
      // gather ALL rows on the own slice, storing z into the owners' blocked Z, then the tcgen05 stage kernel
      nfe += 1;
      RC_TRY(peer_barrier(nullptr));  // slices complete
      ndcn_rhs_desc_t rg;
      std::memset(&rg, 0, sizeof(rg));
      rg.kind = NDCN_RHS_NDCN;
      rg.flags = NDCN_F_NO_CONTROL | NDCN_F_NO_RELU;
      rg.H = sv->feat_hc;
      RhsBinding bg{sv->full_graph, &rg, nullptr, nullptr, 0, sv->sm_count};
      bg.launches = &sv->launches;
      EpiArgs se = store_only(nullptr);
      se.ctrl = e.ctrl;
      fill_feat(se, FEAT_Z_OWNERS);
      t_begin(NDCN_K_GATHER);
      int rcg = 0;
      if (sv->feat_slab) {
        const ndcn_graph* fg = sv->full_graph;
        const int n_rb_real = (int)((fg->v.n_rows + kSlabRows - 1) / kSlabRows);
        // NDCN_FEAT_SPREAD=1 (default): consecutive CTAs take row blocks of different owners, so that the CTAs resident
        // at any moment store z to ALL ranks at once instead of all ranks storing to the same owner (k_gather_slab)
        static const int spread_on = [] { const char* v = std::getenv("NDCN_FEAT_SPREAD"); return v ? std::atoi(v) : 1; }();
        const int spread = spread_on ? sv->feat_world : 1;
        const int n_rb = spread > 1 ? spread * ((n_rb_real + spread - 1) / spread) : n_rb_real;
        const int64_t grid = (int64_t)(sv->feat_hc / 16) * (n_rb + fg->n_long);
        sv->launches += 1;
        // NDCN_FEAT_ROTATE=1 (with NDCN_FEAT_SPREAD=0): every rank starts the walk at its own row block.  Measured on
        // 8 / 4 B200s (profiles/README.md): 5.47 / 9.06 ms per step against 5.33 / 8.76 ms with the common start
        int rb_shift = 0;
        if (const char* v = std::getenv("NDCN_FEAT_ROTATE"))
          if (std::atoi(v)) rb_shift = (int)((int64_t)sv->feat_bounds[sv->feat_rank] / kSlabRows);
        k_gather_slab<<<(unsigned)grid, kStageThreads, 0, st>>>(fg->v, sv->xcs_self, fg->v.n_cols, n_rb, fg->n_long, fg->long_rows,
                                                                rb_shift % std::max(n_rb, 1), spread, n_rb_real, se);
        rcg = (int)cudaGetLastError();
      } else {
        rcg = launch_stage(bg, pp(sv->xcs_self), se, nullptr, st);
      }
      t_end();
      if (rcg != 0) return rcg;
      RC_TRY(peer_barrier(nullptr));  // every block of Z has arrived
      RhsBinding b2 = bind;
      b2.z_block_cols = sv->feat_hc;
      return launch_stage(b2, src, e, n_partials, st);
    }
    RC_TRY(exchange(src_host));
    nfe += 1;
    if (sv->rhs.kind == NDCN_RHS_CALLBACK) {
      if (!sv->rhs.callback || !k_host) return NDCN_E_ARG;
      sv->launches += 1;
      k_stage_time<<<1, 1, 0, st>>>(sv->ctrl, tm_from_ctrl, tm_base, tm_dt, tm_alpha, tm_num, tm_den, tm_mode, sv->t_stage);
      CU_TRY(cudaGetLastError());
      RC_TRY(sv->rhs.callback(sv->rhs.callback_user, src_host, k_host, sv->t_stage));
      // epilogue on the k the callback produced; k is already where it belongs
      EpiArgs e2 = e;
      e2.k_out = pp(nullptr);
      return epi_only(pp(k_host), e2, n_partials);
    }
    return launch_stage(bind, src, e, n_partials, st);
  }

  int epi_only(PtrPair k_in, EpiArgs e, int* n_partials = nullptr) {
    e.partials = sv->partials;
    drop_zero_terms(e);
    const int grid = grid_for_elems(sv->numel, sv->sm_count);
    if (n_partials) *n_partials = grid;
    sv->launches += 1;
    t_begin(NDCN_K_ALGEBRA);
    k_epi_only<<<grid, kStageThreads, 0, st>>>(k_in, sv->numel, e, vec_ok ? 1 : 0);
    t_end();
    return (int)cudaGetLastError();
  }

  bool decoding() const { return o->dec_classes > 0; }
  size_t out_slice_elems() const { return decoding() ? (size_t)sv->n_rows * o->dec_classes : (size_t)sv->numel; }
  // out slice `slot` <- state y: a copy, or y W_d^T + b_d with the fused decoder (NDCN.output_layer)
  int put_state(float* out, int64_t slot, const float* y) {
    float* dst = out + (size_t)slot * out_slice_elems();
    if (!decoding()) {
      if (dst != y) CU_TRY(cudaMemcpyAsync(dst, y, sizeof(float) * (size_t)sv->numel, cudaMemcpyDeviceToDevice, st));
      return 0;
    }
    const int64_t blocks = (sv->n_rows + kWarpsPerCta - 1) / kWarpsPerCta;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)sv->sm_count * 16));
    sv->launches += 1;
    k_decode_rows<<<grid, kStageThreads, 0, st>>>(y, sv->n_rows, sv->H, o->dec_W, o->dec_b, o->dec_classes, dst);
    return (int)cudaGetLastError();
  }

  int poll() {
    CU_TRY(cudaMemcpyAsync(sv->ctrl_host, sv->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return 0;
  }
};


// ---- fixed grid: euler / midpoint / rk4 (3/8) -------------------------------------------
// solvers.py:79-99 with the default grid (= t cast to the state's dtype)
int run_fixed_grid(Driver& d, const float* y0, const double* t, int n_t, float* out) {
  ndcn_solver* sv = d.sv;
  cudaStream_t st = d.st;
  const bool terminal = (d.o->flags & NDCN_O_TERMINAL_ONLY) != 0;
  const bool multi = sv->n_cols > sv->n_rows || d.push();
  const bool in_slab = !terminal && !multi && !d.decoding();  // the output slab doubles as state storage
  const size_t bytes = sizeof(float) * (size_t)sv->numel;
  float* cur;
  if (in_slab) {
    CU_TRY(cudaMemcpyAsync(out, y0, bytes, cudaMemcpyDeviceToDevice, st));
    cur = out;
  } else {
    if (d.push()) RC_TRY(d.peer_barrier(nullptr));
    RC_TRY(d.put_y0(sv->Y[0], y0));
    if (!terminal) RC_TRY(d.put_state(out, 0, y0));
    cur = sv->Y[0];
  }
  for (int i = 0; i + 1 < n_t; ++i) {
    const float t0 = (float)t[i], t1 = (float)t[i + 1];
    const float dt = t1 - t0;  // fp32 subtraction of fp32-rounded times (solvers.py:81,89)
    float* nxt = in_slab ? out + (size_t)(i + 1) * sv->numel : sv->Y[(i + 1) & 1];
    EpiArgs e = d.blank();
    e.dt_src = DT_HOST;
    e.dt_host = dt;
    e.y0 = pp(cur);
    if (sv->method == NDCN_EULER) {  // fixed_grid.py:7-8
      e.mode = EPI_LINCOMB;
      e.beta[0] = 1.0f;
      e.y_out = pp(nxt);
      d.time_host(t0, dt, 0);
      RC_TRY(d.stage(pp(cur), cur, e, sv->K[0]));
    } else if (sv->method == NDCN_MIDPOINT) {  // fixed_grid.py:17-20
      e.mode = EPI_LINCOMB;
      e.beta[0] = 0.5f;  // y + f*dt/2 == y + (dt*0.5)*f bit for bit (power-of-two scaling)
      e.y_out = pp(sv->YS[0]);
      d.time_host(t0, dt, 0);
      RC_TRY(d.stage(pp(cur), cur, e, sv->K[0]));
      e.beta[0] = 1.0f;
      e.y_out = pp(nxt);
      d.time_host(t0, dt, 2, 0.f, 1.f, 2.f);  // t + dt / 2
      RC_TRY(d.stage(pp(sv->YS[0]), sv->YS[0], e, sv->K[0]));
    } else {  // rk_common.py:72-78
      e.mode = EPI_RK4_1;
      e.k_out = pp(sv->K[0]);
      e.y_out = pp(sv->YS[0]);
      d.time_host(t0, dt, 0);
      RC_TRY(d.stage(pp(cur), cur, e, sv->K[0]));
      e.mode = EPI_RK4_2;
      e.k_out = pp(sv->K[1]);
      e.kprev[0] = pp(sv->K[0]);
      e.y_out = pp(sv->YS[1]);
      d.time_host(t0, dt, 2, 0.f, 1.f, 3.f);  // t + dt / 3
      RC_TRY(d.stage(pp(sv->YS[0]), sv->YS[0], e, sv->K[1]));
      e.mode = EPI_RK4_3;
      e.k_out = pp(sv->K[2]);
      e.kprev[1] = pp(sv->K[1]);
      e.y_out = pp(sv->YS[0]);
      d.time_host(t0, dt, 2, 0.f, 2.f, 3.f);  // t + dt * 2 / 3
      RC_TRY(d.stage(pp(sv->YS[1]), sv->YS[1], e, sv->K[2]));
      e.mode = EPI_RK4_4;
      e.k_out = pp(nullptr);
      e.kprev[2] = pp(sv->K[2]);
      e.y_out = pp(nxt);
      d.time_host(t0, dt, 1, 1.f);  // t + dt
      RC_TRY(d.stage(pp(sv->YS[0]), sv->YS[0], e, sv->K[3]));
    }
    if (!in_slab && !terminal) RC_TRY(d.put_state(out, i + 1, nxt));
    cur = nxt;
    sv->ctrl_host->n_accept += 1;
  }
  if (terminal) RC_TRY(d.put_state(out, 0, cur));
  CU_TRY(cudaStreamSynchronize(st));
  sv->ctrl_host->t1 = t[n_t - 1];
  return 0;
}

// controller block at the start of a dopri5 solve (single-GPU element count; the multi-GPU drivers overwrite it)
static void fill_ctrl(Ctrl& h, const Driver& d, const double* t, int n_t) {
  const ndcn_solve_opts_t* o = d.o;
  const bool forced = (o->flags & NDCN_O_FORCED_DT) != 0;
  std::memset(&h, 0, sizeof(h));
  h.t0 = h.t1 = t[0];
  h.rtol = o->rtol;
  h.atol = o->atol;
  // _convert_to_tensor(0.9, float64) goes through an fp32 tensor first (misc.py:39-47)
  h.safety = (double)(float)(o->safety > 0 ? o->safety : 0.9);
  h.ifactor = (double)(float)(o->ifactor > 0 ? o->ifactor : 10.0);
  h.dfactor = (double)(float)(o->dfactor > 0 ? o->dfactor : 0.2);
  h.forced = forced ? 1 : 0;
  h.forced_dt = o->forced_dt;
  const bool given_first = !forced && o->first_step > 0.0;
  h.dt = forced ? o->forced_dt : (given_first ? o->first_step : 0.0);
  h.first_step = h.dt;
  h.numel_global = (double)d.sv->numel;
  h.max_num_steps = o->max_num_steps > 0 ? o->max_num_steps : 2147483647LL;
  h.next_out = 1;
  h.n_out = n_t;
  h.emit_lo = h.emit_hi = 1;
  h.terminal_only = (o->flags & NDCN_O_TERMINAL_ONLY) ? 1 : 0;
}

// ---- dopri5 ------------------------------------------------------------------------------
struct Dopri {
  Driver& d;
  ndcn_solver* sv;
  float beta32[6][8];
  float err32[8];
  float alpha32[6];
  EmitArgs emit;
  bool host_parity;  // hooks need the host to know the buffer parity: poll every attempt
  int par = 0;
  bool err_prefix = true;  // stage 5 leaves the error-estimate prefix behind (NDCN_ERR_PREFIX=0: off, A/B runs)

  explicit Dopri(Driver& dr) : d(dr), sv(dr.sv) {
    std::memset(beta32, 0, sizeof(beta32));
    std::memset(err32, 0, sizeof(err32));
    for (int s = 0; s < 6; ++s)
      for (int j = 0; j <= s; ++j) beta32[s][j] = (float)kDpBeta[s][j];
    for (int j = 0; j < 7; ++j) err32[j] = (float)kDpErr[j];
    for (int j = 0; j < 6; ++j) alpha32[j] = (float)kDpAlpha[j];
    if (const char* v = std::getenv("NDCN_ERR_PREFIX")) err_prefix = std::atoi(v) != 0;
  }

  int reduce_and_control(int n_partials) {
    const bool multi = d.multi();
    if (!multi) {
      sv->launches += 1;
      d.t_begin(NDCN_K_CONTROL);
      k_controller<<<1, kStageThreadsCtl, 0, d.st>>>(sv->ctrl, sv->partials, n_partials, sv->t_out, sv->xchg, 0);
      d.t_end();
    } else {
      sv->launches += 2;
      k_controller<<<1, kStageThreadsCtl, 0, d.st>>>(sv->ctrl, sv->partials, n_partials, sv->t_out, sv->xchg, 1);
      RC_TRY(d.allreduce_xchg());
      k_controller<<<1, kStageThreadsCtl, 0, d.st>>>(sv->ctrl, sv->partials, n_partials, sv->t_out, sv->xchg, 2);
    }
    return (int)cudaGetLastError();
  }

  int attempt() {
    float* Yq = sv->Y[par ^ 1];
    float* KFq = sv->KF[par ^ 1];
    const PtrPair Ycur = pp(sv->Y[0], sv->Y[1]), Yoth = pp(sv->Y[1], sv->Y[0]);
    const PtrPair KFcur = pp(sv->KF[0], sv->KF[1]), KFoth = pp(sv->KF[1], sv->KF[0]);
    EpiArgs e = d.blank();
    e.ctrl = sv->ctrl;
    e.dt_src = DT_CTRL;
    e.y0 = Ycur;
    // stage input 1: y0 + (dt*b10) k0, k0 = FSAL derivative; also the finite-state guard
    e.mode = EPI_LINCOMB;
    e.n_prev = 0;
    e.beta[0] = beta32[0][0];
    e.check_finite = 1;
    e.y_out = pp(sv->YS[0]);
    RC_TRY(d.epi_only(KFcur, e));
    e.check_finite = 0;
    e.kprev[0] = KFcur;
    for (int s = 1; s <= 5; ++s) {
      float* src = sv->YS[(s - 1) & 1];
      e.n_prev = s;
      for (int j = 0; j < 8; ++j) e.beta[j] = beta32[s][j];
      e.k_out = pp(sv->K[s - 1]);
      if (s >= 2) e.kprev[s - 1] = pp(sv->K[s - 2]);
      e.y_out = (s < 5) ? pp(sv->YS[s & 1]) : Yoth;  // stage 5 forms y1 (FSAL: c_sol == beta[-1])
      if (s == 5 && err_prefix) {
        // this stage reads k1..k5 and holds k6: it also leaves sum_{j<=6} (dt*c_err_j) k_j (the left-to-right prefix
        // of rk_common.py:60) in YS[1] (own rows; free since stage 4 consumed it), so the error stage reads 1 stream, not 6
        e.mode = EPI_LINCOMB_E;
        e.e_out = sv->YS[1];
        for (int j = 0; j < 8; ++j) e.ebeta[j] = j < 6 ? err32[j] : 0.0f;
      }
      d.time_ctrl(1, alpha32[s - 1]);
      RC_TRY(d.stage(pp(src), src, e, sv->K[s - 1]));
    }
    // stage 6: k7 = f(y1) + error estimate
    e.mode = EPI_ERR;
    e.e_out = nullptr;
    if (err_prefix) {
      e.n_prev = 1;
      e.err_prefix = 1;
      for (int j = 0; j < 8; ++j) e.beta[j] = 0.0f;
      e.beta[1] = err32[6];
      e.kprev[0] = pp(sv->YS[1]);
    } else {
      e.n_prev = 6;
      for (int j = 0; j < 8; ++j) e.beta[j] = err32[j];
      e.kprev[5] = pp(sv->K[4]);
    }
    e.k_out = KFoth;
    e.y_out = pp(nullptr);
    e.y1 = Yoth;
    e.rtol = (float)d.o->rtol;
    e.atol = (float)d.o->atol;
    int n_partials = 0;
    d.time_ctrl(1, alpha32[5]);
    RC_TRY(d.stage(Yoth, Yq, e, KFq, &n_partials));
    RC_TRY(reduce_and_control(n_partials));
    sv->launches += 1;
    d.t_begin(NDCN_K_EMIT);
    if (d.decoding()) {
      const int64_t blocks = (sv->n_rows + kWarpsPerCta - 1) / kWarpsPerCta;
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)sv->sm_count * 16));
      k_emit_decode<<<grid, kStageThreads, 0, d.st>>>(emit);
    } else {
      k_emit<<<grid_for_elems(sv->numel, sv->sm_count), kStageThreads, 0, d.st>>>(emit, d.vec_ok ? 1 : 0);
    }
    d.t_end();
    return (int)cudaGetLastError();
  }

  int init_scalar(int n_partials, int phase, double t_first) {
    const bool multi = d.multi();
    if (!multi) {
      sv->launches += 1;
      k_init_scalar<<<1, kStageThreadsCtl, 0, d.st>>>(sv->ctrl, sv->partials, n_partials, phase, t_first, sv->xchg, 0);
    } else {
      sv->launches += 2;
      k_init_scalar<<<1, kStageThreadsCtl, 0, d.st>>>(sv->ctrl, sv->partials, n_partials, phase, t_first, sv->xchg, 1);
      RC_TRY(d.allreduce_xchg());
      k_init_scalar<<<1, kStageThreadsCtl, 0, d.st>>>(sv->ctrl, sv->partials, n_partials, phase, t_first, sv->xchg, 2);
    }
    return (int)cudaGetLastError();
  }

  int run(const float* y0, const double* t, int n_t, float* out) {
    cudaStream_t st = d.st;
    const bool terminal = (d.o->flags & NDCN_O_TERMINAL_ONLY) != 0;
    const bool forced = (d.o->flags & NDCN_O_FORCED_DT) != 0;
    host_parity = d.o->exchange != nullptr || sv->rhs.kind == NDCN_RHS_CALLBACK;

    // requested times -> device (float64, already fp32-rounded by ODEBlock when it applies)
    if (sv->t_cap < n_t) {
      if (sv->t_out) cudaFree(sv->t_out);
      sv->t_cap = std::max(n_t, 128);
      CU_TRY(cudaMalloc((void**)&sv->t_out, sizeof(double) * sv->t_cap));
    }
    CU_TRY(cudaMemcpyAsync(sv->t_out, t, sizeof(double) * n_t, cudaMemcpyHostToDevice, st));

    Ctrl& h = *sv->ctrl_host;
    fill_ctrl(h, d, t, n_t);
    const bool given_first = !forced && d.o->first_step > 0.0;
    if (d.feat()) {
      h.numel_global = (double)sv->feat_bounds[sv->feat_world] * (double)sv->H;
    } else if (d.push()) {
      h.numel_global = (double)sv->n_cols * (double)sv->H;  // full halo: n_cols = all nodes
    } else if (d.o->exchange) {
      // global element count = all-reduce of the local one (same hook, what = 1)
      double tmp[2] = {(double)sv->numel, 0.0};
      CU_TRY(cudaMemcpyAsync(sv->xchg, tmp, sizeof(tmp), cudaMemcpyHostToDevice, st));
      RC_TRY(d.o->exchange(d.o->exchange_user, 1, sv->xchg));
      CU_TRY(cudaMemcpyAsync(tmp, sv->xchg, sizeof(tmp), cudaMemcpyDeviceToHost, st));
      CU_TRY(cudaStreamSynchronize(st));
      h.numel_global = tmp[0];
    }
    CU_TRY(cudaMemcpyAsync(sv->ctrl, &h, sizeof(Ctrl), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));  // ctrl_host is reused as the poll target below

    // peer push: nobody may still be reading the halo rows of an earlier solve when y0 arrives
    if (d.push()) RC_TRY(d.peer_barrier(nullptr));
    RC_TRY(d.put_y0(sv->Y[0], y0));
    if (!terminal) RC_TRY(d.put_state(out, 0, y0));

    emit.ctrl = sv->ctrl;
    emit.t_out = sv->t_out;
    emit.y0 = pp(sv->Y[0], sv->Y[1]);
    emit.y1 = pp(sv->Y[1], sv->Y[0]);
    emit.k0 = pp(sv->KF[0], sv->KF[1]);
    emit.k6 = pp(sv->KF[1], sv->KF[0]);
    for (int j = 0; j < 5; ++j) emit.k[j] = sv->K[j];
    for (int j = 0; j < 7; ++j) emit.c_mid[j] = (float)kDpMid[j];
    emit.out = out;
    emit.numel = sv->numel;
    emit.dec_W = d.decoding() ? d.o->dec_W : nullptr;
    emit.dec_b = d.decoding() ? d.o->dec_b : nullptr;
    emit.dec_C = d.decoding() ? d.o->dec_classes : 0;
    emit.H = sv->H;
    emit.n_rows = sv->n_rows;

    // f0 = func(t0, y0)     dopri5.py:78
    d.time_host((float)t[0], 0.f, 0);
    RC_TRY(d.stage(pp(sv->Y[0]), sv->Y[0], store_only(sv->KF[0]), sv->KF[0]));
    if (!forced && !given_first) {
      // _select_initial_step(order=4)     dopri5.py:80, misc.py:84-143
      const int grid = grid_for_elems(sv->numel, sv->sm_count);
      if (grid > sv->max_partials) return NDCN_E_WORKSPACE;  // k_init_norms writes 2 doubles per CTA
      sv->launches += 1;
      k_init_norms<<<grid, kStageThreads, 0, st>>>(sv->Y[0], sv->KF[0], nullptr, sv->numel, (float)d.o->rtol,
                                                    (float)d.o->atol, 0, sv->partials);
      RC_TRY(init_scalar(grid, 0, t[0]));
      EpiArgs e = d.blank();
      e.ctrl = sv->ctrl;
      e.dt_src = DT_CTRL_H0;
      e.mode = EPI_LINCOMB;
      e.beta[0] = 1.0f;
      e.y0 = pp(sv->Y[0]);
      e.y_out = pp(sv->YS[0]);
      RC_TRY(d.epi_only(pp(sv->KF[0]), e));  // y0 + h0*f0
      d.time_ctrl(3, 0.f, (float)t[0]);      // func(t0 + h0, .), misc.py:126
      RC_TRY(d.stage(pp(sv->YS[0]), sv->YS[0], store_only(sv->K[0]), sv->K[0]));
      sv->launches += 1;
      k_init_norms<<<grid, kStageThreads, 0, st>>>(sv->Y[0], sv->KF[0], sv->K[0], sv->numel, (float)d.o->rtol,
                                                    (float)d.o->atol, 1, sv->partials);
      RC_TRY(init_scalar(grid, 1, t[0]));
    }
    CU_TRY(cudaGetLastError());

    // attempts are enqueued in growing batches; kernels of attempts past the end are no-ops.
    // With a forced dt the host can replay the controller's float64 time accumulation, so
    // exactly the needed attempts are enqueued and the controller is polled once at the end.
    int64_t forced_left = -1;
    if (forced) {
      forced_left = 0;
      double t1 = t[0];
      while (t[n_t - 1] > t1 && forced_left < (int64_t)1 << 40) {
        t1 += d.o->forced_dt;
        ++forced_left;
      }
    }
    int batch = host_parity ? 1 : 2;
    int64_t issued = 0;
    for (;;) {
      int n_now = batch;
      if (forced_left >= 0 && !host_parity) n_now = (int)std::min<int64_t>(forced_left - issued, 64);
      if (n_now < 1) n_now = 1;
      for (int i = 0; i < n_now; ++i) {
        d.cur_attempt = issued++;
        RC_TRY(attempt());
      }
      if (forced_left >= 0 && !host_parity && issued < forced_left) continue;  // no poll needed yet
      RC_TRY(d.poll());
      const Ctrl& c = *sv->ctrl_host;
      par = c.parity;
      if (c.done) break;
      if (!host_parity) batch = std::min(batch * 2, 16);
    }
    return 0;
  }
};


// ---- persistent whole-solve kernel (small graphs) ---------------------------------------------------------
static bool small_eligible(const ndcn_solver* sv, const ndcn_solve_opts_t* o) {
  if (sv->n_push > 0 || sv->feat_on || o->exchange != nullptr || o->gather_mode != NDCN_GATHER_LOCAL) return false;
  if (sv->n_cols != sv->n_rows || sv->rhs.kind == NDCN_RHS_CALLBACK) return false;
  if (sv->H < 1 || sv->H > 1024 || sv->n_rows < 1) return false;
  if (sv->n_rows > cfg().small_max_rows || sv->numel > cfg().small_max_numel) return false;
  int coop = 0;
  if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, current_device()) != cudaSuccess || !coop) return false;
  return true;
}

template <int VW, int NCH, bool CONTROL>
static int launch_small(const SmallArgs& a, size_t smem, int sm_count, int want_blocks, int* grid_out, cudaStream_t st) {
  static PerDeviceOnce attr;
  if (attr.need() && smem > 48 * 1024) {
    CU_TRY(cudaFuncSetAttribute(k_solve_small<VW, NCH, CONTROL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr.mark();
  }
  int per_sm = 0;
  CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_small<VW, NCH, CONTROL>, kStageThreads, smem));
  if (per_sm < 1) return NDCN_E_ARG;
  // all CTAs must be co-resident (grid barriers); more CTAs than the work needs only make the barriers dearer
  const int grid = std::max(1, std::min(want_blocks, per_sm * sm_count));
  *grid_out = grid;
  SmallArgs copy = a;
  void* params[] = {&copy};
  return (int)cudaLaunchCooperativeKernel((void*)k_solve_small<VW, NCH, CONTROL>, dim3(grid), dim3(kStageThreads), params,
                                          smem, st);
}

static int run_small(Driver& d, const float* y0, const double* t, int n_t, float* out) {
  ndcn_solver* sv = d.sv;
  cudaStream_t st = d.st;
  const ndcn_solve_opts_t* o = d.o;
  const bool terminal = (o->flags & NDCN_O_TERMINAL_ONLY) != 0;
  const bool forced = (o->flags & NDCN_O_FORCED_DT) != 0;
  SmallArgs a;
  std::memset(&a, 0, sizeof(a));
  a.g = sv->g->v;
  a.kind = sv->rhs.kind;
  a.H = sv->H;
  a.flags = sv->rhs.flags;
  a.W = sv->rhs.W;
  a.Wt = sv->Wt;
  a.bias = sv->rhs.b;
  for (int i = 0; i < 8; ++i) a.p[i] = sv->rhs.p[i];
  a.method = sv->method;
  a.n_t = n_t;
  a.terminal = terminal ? 1 : 0;
  a.forced = forced ? 1 : 0;
  a.given_first = (!forced && o->first_step > 0.0) ? 1 : 0;
  a.in_slab = (sv->method != NDCN_DOPRI5 && !terminal && !d.decoding()) ? 1 : 0;
  a.y0 = y0;
  a.out = out;
  for (int i = 0; i < 2; ++i) { a.Y[i] = sv->Y[i]; a.YS[i] = sv->YS[i]; a.KF[i] = sv->KF[i]; }
  for (int i = 0; i < 5; ++i) a.K[i] = sv->K[i];
  a.ctrl = sv->ctrl;
  a.partials = sv->partials;
  for (int s = 0; s < 6; ++s)
    for (int j = 0; j <= s; ++j) a.beta32[s][j] = (float)kDpBeta[s][j];
  for (int j = 0; j < 7; ++j) { a.err32[j] = (float)kDpErr[j]; a.c_mid[j] = (float)kDpMid[j]; }
  a.rtol = (float)o->rtol;
  a.atol = (float)o->atol;
  a.t_first = t[0];
  a.dec_W = d.decoding() ? o->dec_W : nullptr;
  a.dec_b = d.decoding() ? o->dec_b : nullptr;
  a.dec_C = d.decoding() ? o->dec_classes : 0;
  a.numel = sv->numel;
  a.n_rows = (int)sv->n_rows;
  a.vec = d.vec_ok ? 1 : 0;
  a.err_prefix = 1;
  if (const char* v = std::getenv("NDCN_ERR_PREFIX")) a.err_prefix = std::atoi(v) != 0;
  if (sv->rhs.kind == NDCN_RHS_NDCN && !(sv->rhs.flags & NDCN_F_NO_CONTROL) && (!sv->rhs.W || !sv->rhs.b)) return NDCN_E_ARG;

  // requested times (dopri5, float64) or step sizes (fixed grid, fp32) -> the solver's device scratch
  if (sv->t_cap < n_t) {
    if (sv->t_out) cudaFree(sv->t_out);
    sv->t_cap = std::max(n_t, 128);
    CU_TRY(cudaMalloc((void**)&sv->t_out, sizeof(double) * sv->t_cap));
  }
  std::vector<float> dts;
  if (sv->method == NDCN_DOPRI5) {
    CU_TRY(cudaMemcpyAsync(sv->t_out, t, sizeof(double) * n_t, cudaMemcpyHostToDevice, st));
    a.t_out = sv->t_out;
    Ctrl& h = *sv->ctrl_host;
    fill_ctrl(h, d, t, n_t);
    CU_TRY(cudaMemcpyAsync(sv->ctrl, &h, sizeof(Ctrl), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));  // ctrl_host is the read-back target below; `t` may be a temporary
  } else {
    dts.resize(std::max(1, n_t - 1));
    for (int i = 0; i + 1 < n_t; ++i) dts[i] = (float)t[i + 1] - (float)t[i];  // solvers.py:81,89: fp32 grid, fp32 difference
    CU_TRY(cudaMemcpyAsync(sv->t_out, dts.data(), sizeof(float) * dts.size(), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));
    a.dts = reinterpret_cast<const float*>(sv->t_out);
  }

  const bool wide = sv->rhs.kind == NDCN_RHS_NDCN && fast_width(sv->H) && sv->H > 32 && aligned16(y0) && aligned16(out);
  const bool tiled = wide && !(sv->rhs.flags & NDCN_F_NO_CONTROL);
  const bool rowvec = wide && (sv->rhs.flags & NDCN_F_NO_CONTROL);
  const bool dyn1 = sv->rhs.kind != NDCN_RHS_NDCN && sv->H == 1;  // 8 rows per warp
  const int by_rows = (int)((sv->n_rows + (dyn1 ? 64 : kWarpsPerCta) - 1) / (dyn1 ? 64 : kWarpsPerCta));
  const int by_tiles = (int)((sv->n_rows + kTileRows - 1) / kTileRows);
  const int by_elems = (int)((sv->numel + kStageThreads * 4 - 1) / (kStageThreads * 4));
  const int want = std::max(tiled ? by_tiles : std::min(by_rows, 4 * sv->sm_count), std::min(by_elems, sv->sm_count));
  int grid = 0, rc = 0;
  sv->launches += 1;
  d.t_begin(NDCN_K_STAGE);
  // one CTA, element-parallel, block barriers: the small [N,d] ground-truth states (csrc/small_solver.cuh::tiny_stage)
  static const bool tiny_on = [] { const char* v = std::getenv("NDCN_TINY"); return !v || std::atoi(v) != 0; }();
  const bool tiny = tiny_on && sv->numel <= kTinyMaxNumel && sv->rhs.kind != NDCN_RHS_NDCN;
  if (tiny) {
    const size_t smem = 16;
    SmallArgs copy = a;
    grid = 1;
    k_solve_small<0, 0, false, true><<<1, kTinyThreads, smem, st>>>(copy);
    rc = (int)cudaGetLastError();
  } else if (tiled) {
    switch (sv->H) {
      case 256: rc = launch_small<4, 2, true>(a, GemmSmem<4, 2>::total, sv->sm_count, want, &grid, st); break;
      case 128: rc = launch_small<4, 1, true>(a, GemmSmem<4, 1>::total, sv->sm_count, want, &grid, st); break;
      default: rc = launch_small<2, 1, true>(a, GemmSmem<2, 1>::total, sv->sm_count, want, &grid, st); break;
    }
  } else if (rowvec) {
    switch (sv->H) {
      case 256: rc = launch_small<4, 2, false>(a, 128, sv->sm_count, want, &grid, st); break;
      case 128: rc = launch_small<4, 1, false>(a, 128, sv->sm_count, want, &grid, st); break;
      default: rc = launch_small<2, 1, false>(a, 128, sv->sm_count, want, &grid, st); break;
    }
  } else {
    rc = launch_small<0, 0, false>(a, sizeof(float) * kWarpsPerCta * std::max(sv->H, 32), sv->sm_count, want, &grid, st);
  }
  d.t_end();
  if (rc != 0) return rc;
  if (2 * grid > 2 * sv->max_partials) return NDCN_E_WORKSPACE;
  if (sv->method == NDCN_DOPRI5) {
    RC_TRY(d.poll());
  } else {
    CU_TRY(cudaStreamSynchronize(st));
    sv->ctrl_host->n_accept = n_t - 1;
    sv->ctrl_host->t1 = t[n_t - 1];
    const int per_step = sv->method == NDCN_EULER ? 1 : (sv->method == NDCN_MIDPOINT ? 2 : 4);
    d.nfe = (int64_t)per_step * (n_t - 1);
  }
  return 0;
}

}  // namespace

// small_mode: 0 = the persistent whole-solve kernel when the problem is eligible and the configuration allows it,
//             1 = require it (ndcn_odeint_small_f32), -1 = never
static int odeint_impl(ndcn_solver_t* sv, const float* y0, const double* t_host, int32_t n_t, float* out,
                       const ndcn_solve_opts_t* opts, ndcn_solve_stats_t* stats, ndcn_stream_t s, int small_mode) {
  if (!sv || !y0 || !t_host || !out || !opts || n_t < 1) return NDCN_E_ARG;
  if (opts->method != sv->method) return NDCN_E_METHOD;
  for (int i = 0; i + 1 < n_t; ++i)
    if (!(t_host[i + 1] > t_host[i])) return NDCN_E_ARG;  // misc.py:59-60
  cudaStream_t st = (cudaStream_t)s;
  sv->launches = 0;
  Driver d(sv, opts, st);
  d.vec_ok = aligned16(out) && aligned16(y0) && (sv->numel % 4 == 0);
  for (int j = 0; j < sv->n_push; ++j) d.vec_ok = d.vec_ok && (sv->push_delta[j] % 16 == 0);  // peer copies of 16-byte stores
  d.timing = (opts->flags & NDCN_O_TIME_KERNELS) != 0;
  if (stats) std::memset(stats, 0, sizeof(*stats));
  std::memset(sv->ctrl_host, 0, sizeof(Ctrl));

  const bool small_ok = small_eligible(sv, opts);
  if (small_mode == 1 && !small_ok) return NDCN_E_ARG;
  // auto: where the persistent kernel measured faster than a launch per stage on B200 (profiles/README.md): the
  // narrow NDCN widths of the dynamics scripts (H <= 32) and the [N,1] ground-truth dynamics; the wide Cora block
  // (H = 256) runs as fast launch by launch (its phases are long enough to hide the launch latency)
  const bool small_pays = (sv->rhs.kind == NDCN_RHS_NDCN && sv->H <= 32) || (sv->rhs.kind != NDCN_RHS_NDCN && sv->H == 1);
  const bool use_small = n_t > 1 && small_ok &&
                         (small_mode == 1 || (small_mode == 0 && cfg().small_solver && small_pays));
  const bool fast_h = sv->H == 256 || sv->H == 128 || sv->H == 64 || sv->H == 32;
  if (sv->rhs.kind == NDCN_RHS_NDCN && !(sv->rhs.flags & NDCN_F_NO_CONTROL) && fast_h) {
    const int n = sv->H * sv->H;
    k_transpose<<<(n + 255) / 256, 256, 0, st>>>(sv->rhs.W, sv->Wt, sv->H);
    sv->launches += 1;
  }
  if (sv->rhs.kind == NDCN_RHS_NDCN && fast_h && !(aligned16(out) && aligned16(y0))) return NDCN_E_ARG;
  if (opts->gather_mode != NDCN_GATHER_LOCAL && opts->gather_mode != NDCN_GATHER_EXTERNAL) return NDCN_E_ARG;
  if (opts->dec_classes < 0 || opts->dec_classes > kDecMaxC || (opts->dec_classes > 0 && !opts->dec_W)) return NDCN_E_ARG;
  const bool external = opts->gather_mode == NDCN_GATHER_EXTERNAL || sv->feat_on;
  if (sv->feat_on && (opts->gather_mode != NDCN_GATHER_LOCAL || opts->exchange)) return NDCN_E_ARG;
  if (external && !(umma_eligible(sv->rhs, ((int64_t)1) << 40) && sv->Wimg && sv->Z)) return NDCN_E_ARG;
  if (!use_small && (external || umma_eligible(sv->rhs, sv->n_rows)) && sv->Wimg) {
    if (sv->H == 256) prep_w_image<256>(sv->rhs.W, sv->Wimg, st);
    else prep_w_image<128>(sv->rhs.W, sv->Wimg, st);
    sv->launches += 1;
  }

  int rc = 0;
  if (n_t == 1) {
    rc = d.put_state(out, 0, y0);
    if (!rc) rc = (int)cudaStreamSynchronize(st);
  } else if (use_small) {
    rc = run_small(d, y0, t_host, n_t, out);
  } else if (sv->method == NDCN_DOPRI5) {
    Dopri dp(d);
    rc = dp.run(y0, t_host, n_t, out);
  } else {
    rc = run_fixed_grid(d, y0, t_host, n_t, out);
  }
  if (d.push() && rc == 0) {
    // a barrier that timed out has flagged it in this rank's signal pad (the fixed-grid drivers have no
    // controller block that could carry the status)
    int timed_out = 0;
    cudaStreamSynchronize(st);
    if (cudaMemcpy(&timed_out, &sv->peers.self->timed_out, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess && timed_out)
      rc = NDCN_E_PEER_TIMEOUT;
  }
  if (d.timing) {
    cudaStreamSynchronize(st);
    d.t_collect(sv->method == NDCN_DOPRI5 ? (int64_t)sv->ctrl_host->n_attempt : (int64_t)1 << 60, stats);
  }
  if (stats) {
    const Ctrl& c = *sv->ctrl_host;
    // speculative attempts enqueued past the end are device-side no-ops: count the evaluations
    // that really ran (f0 [+ the initial-step probe] + 6 per attempt the controller saw)
    if (sv->method == NDCN_DOPRI5 && n_t > 1)
      stats->nfe = (((opts->flags & NDCN_O_FORCED_DT) || opts->first_step > 0.0) ? 1 : 2) + 6 * (int64_t)c.n_attempt;
    else
      stats->nfe = d.nfe;
    stats->n_accepted = c.n_accept;
    stats->n_rejected = c.n_reject;
    stats->n_launches = sv->launches;
    stats->first_step = c.first_step;
    stats->last_dt = c.dt;
    stats->t_final = c.t1;
    stats->status = rc < 0 ? rc : c.status;
    stats->reserved = 0;
  }
  if (rc != 0) return rc;
  return sv->ctrl_host->status;
}

extern "C" int ndcn_odeint_f32(ndcn_solver_t* sv, const float* y0, const double* t_host, int32_t n_t, float* out,
                               const ndcn_solve_opts_t* opts, ndcn_solve_stats_t* stats, ndcn_stream_t s) {
  return odeint_impl(sv, y0, t_host, n_t, out, opts, stats, s, 0);
}

extern "C" int ndcn_odeint_small_f32(ndcn_solver_t* sv, const float* y0, const double* t_host, int32_t n_t, float* out,
                                     const ndcn_solve_opts_t* opts, ndcn_solve_stats_t* stats, ndcn_stream_t s) {
  return odeint_impl(sv, y0, t_host, n_t, out, opts, stats, s, 1);
}

extern "C" int ndcn_odeint_staged_f32(ndcn_solver_t* sv, const float* y0, const double* t_host, int32_t n_t, float* out,
                                      const ndcn_solve_opts_t* opts, ndcn_solve_stats_t* stats, ndcn_stream_t s) {
  return odeint_impl(sv, y0, t_host, n_t, out, opts, stats, s, -1);
}

// ---------------------------------------------------------------------------------------
// multi-GPU peer push: configuration and IPC-shareable allocations
// ---------------------------------------------------------------------------------------
extern "C" int ndcn_solver_set_peers(ndcn_solver_t* sv, const ndcn_peer_config_t* cfg) {
  if (!sv) return NDCN_E_ARG;
  if (sv->feat_on) return NDCN_E_ARG;  // one scheme at a time (the feature-sharded push owns sv->peers)
  sv->n_push = 0;
  std::memset(&sv->peers, 0, sizeof(sv->peers));
  if (!cfg || cfg->world <= 1) return NDCN_OK;
  if (cfg->world > kMaxPeers + 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return NDCN_E_ARG;
  if (sv->n_cols <= sv->n_rows) return NDCN_E_ARG;  // needs halo rows to push into
  if (sv->rhs.kind == NDCN_RHS_CALLBACK) return NDCN_E_ARG;
  for (int r = 0; r < cfg->world; ++r)
    if (!cfg->pad[r] || ((uintptr_t)cfg->pad[r] & 15u)) return NDCN_E_ARG;
  sv->peers.self = (PeerPad*)cfg->pad[cfg->rank];
  sv->peers.rank = cfg->rank;
  sv->peers.world = cfg->world;
  int j = 0;
  for (int r = 0; r < cfg->world; ++r) {
    sv->peers.peer[r] = (PeerPad*)cfg->pad[r];
    if (r == cfg->rank) continue;
    if (cfg->delta_bytes[r] % 4 != 0) return NDCN_E_ARG;
    sv->push_delta[j++] = (long long)cfg->delta_bytes[r];
  }
  sv->n_push = cfg->world - 1;
  return NDCN_OK;
}

extern "C" int ndcn_solver_set_feature_peers(ndcn_solver_t* sv, const ndcn_graph_t* full_graph,
                                             const ndcn_feature_peer_config_t* cfg) {
  if (!sv) return NDCN_E_ARG;
  if (sv->feat_on) {  // off / reconfigure: give the workspace Z back
    sv->Z = sv->Z_own;
    sv->feat_on = false;
  }
  if (!cfg || cfg->world <= 1) return NDCN_OK;
  const int P = cfg->world;
  if (!full_graph || P > 8 || cfg->rank < 0 || cfg->rank >= P) return NDCN_E_ARG;
  if (sv->rhs.kind != NDCN_RHS_NDCN || (sv->rhs.flags & (NDCN_F_NO_GRAPH | NDCN_F_NO_CONTROL))) return NDCN_E_ARG;
  if ((sv->H != 128 && sv->H != 256) || sv->H % P != 0 || !sv->Wimg || !sv->Z) return NDCN_E_ARG;
  const int hc = sv->H / P;
  if (hc < 32 || (hc & (hc - 1)) != 0) return NDCN_E_ARG;
  if (sv->n_push > 0) return NDCN_E_ARG;  // one scheme at a time
  if (cfg->row_bounds[0] != 0 || cfg->row_bounds[P] != full_graph->v.n_rows) return NDCN_E_ARG;
  for (int r = 0; r < P; ++r) {
    if (cfg->row_bounds[r + 1] <= cfg->row_bounds[r]) return NDCN_E_ARG;
    if (!cfg->pad[r] || !cfg->xcs[r] || !cfg->z[r] || ((uintptr_t)cfg->xcs[r] & 15u) || ((uintptr_t)cfg->z[r] & 15u))
      return NDCN_E_ARG;
  }
  if (cfg->row_bounds[cfg->rank + 1] - cfg->row_bounds[cfg->rank] != sv->n_rows) return NDCN_E_ARG;
  if (full_graph->v.n_cols != full_graph->v.n_rows) return NDCN_E_ARG;
  FeatTable t;
  std::memset(&t, 0, sizeof(t));
  for (int r = 0; r < P; ++r) {
    t.xcs[r] = (float*)cfg->xcs[r];
    t.z[r] = (float*)cfg->z[r];
  }
  for (int r = 0; r <= 8; ++r) t.bounds[r] = (int)cfg->row_bounds[r <= P ? r : P];
  t.world = P;
  t.n_total = (int)cfg->row_bounds[P];
  // slice rows of 128 / 256 bytes (Hc = 32 / 64) are what the slab layout fixes; from Hc = 128 on the warp-per-row
  // gather on row-major slices is faster (measured at 2 GPUs: 1.29 ms vs 1.53 ms per gather at Hc = 128)
  sv->feat_slab = hc <= 64;
  if (const char* v = std::getenv("NDCN_FEAT_SLAB")) sv->feat_slab = std::atoi(v) != 0;
  t.slab = sv->feat_slab ? 1 : 0;
  t.nl_uniform = (int)(cfg->row_bounds[1] - cfg->row_bounds[0]);
  for (int r = 0; r < P; ++r) {
    const int64_t want = std::min<int64_t>((int64_t)r * t.nl_uniform, t.n_total);
    if (cfg->row_bounds[r] != want) t.nl_uniform = 0;  // ragged blocks: owner by table walk
  }
  if (!sv->feat_dev) CU_TRY(cudaMalloc((void**)&sv->feat_dev, sizeof(FeatTable)));
  CU_TRY(cudaMemcpy(sv->feat_dev, &t, sizeof(t), cudaMemcpyHostToDevice));
  std::memset(&sv->peers, 0, sizeof(sv->peers));
  sv->peers.self = (PeerPad*)cfg->pad[cfg->rank];
  sv->peers.rank = cfg->rank;
  sv->peers.world = P;
  for (int r = 0; r < P; ++r) sv->peers.peer[r] = (PeerPad*)cfg->pad[r];
  for (int r = 0; r <= 8; ++r) sv->feat_bounds[r] = r <= P ? (int)cfg->row_bounds[r] : (int)cfg->row_bounds[P];
  sv->feat_rank = cfg->rank;
  sv->feat_world = P;
  sv->feat_hc = hc;
  sv->full_graph = full_graph;
  sv->xcs_self = (float*)cfg->xcs[cfg->rank];
  sv->Z_own = sv->Z;
  sv->Z = (float*)cfg->z[cfg->rank];
  sv->feat_on = true;
  return NDCN_OK;
}

extern "C" int ndcn_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out /* 64 bytes */) {
  if (!ptr_out || !handle_out || bytes == 0) return NDCN_E_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  void* p = nullptr;
  CU_TRY(cudaMalloc(&p, bytes));
  cudaError_t ce = cudaMemset(p, 0, std::min<size_t>(bytes, 4096));  // the PeerPad head starts at epoch 0
  cudaIpcMemHandle_t h;
  if (ce == cudaSuccess) ce = cudaIpcGetMemHandle(&h, p);
  if (ce != cudaSuccess) {
    cudaFree(p);
    return (int)ce;
  }
  std::memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return NDCN_OK;
}

extern "C" int ndcn_peer_open(const unsigned char* handle, void** ptr_out) {
  if (!handle || !ptr_out) return NDCN_E_ARG;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  CU_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = p;
  return NDCN_OK;
}

extern "C" int ndcn_peer_close(void* ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : NDCN_OK; }
extern "C" int ndcn_peer_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : NDCN_OK; }

// same-process, same-device stand-in for the IPC mapping (tests: two "ranks" as two threads on one GPU)
extern "C" int ndcn_peer_enable_access(int peer_device) {
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  if (dev == peer_device) return NDCN_OK;
  cudaError_t ce = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (ce == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return NDCN_OK;
  }
  return (int)ce;
}

// ---------------------------------------------------------------------------------------
// solver algebra as stand-alone kernels
// ---------------------------------------------------------------------------------------
extern "C" int ndcn_rk_combine_f32(float* out, const float* y0, const float* const* k_host_ptrs,
                                   const double* beta_host, int32_t n_k, float dt, int64_t numel, ndcn_stream_t s) {
  if (!out || !y0 || !k_host_ptrs || !beta_host || n_k < 1 || n_k > 7 || numel < 0) return NDCN_E_ARG;
  cudaStream_t st = (cudaStream_t)s;
  EpiArgs e;
  std::memset(&e, 0, sizeof(e));
  e.mode = EPI_LINCOMB;
  e.n_prev = n_k - 1;
  e.dt_src = DT_HOST;
  e.dt_host = dt;
  e.y0 = pp(const_cast<float*>(y0));
  e.y_out = pp(out);
  for (int j = 0; j < n_k - 1; ++j) e.kprev[j] = pp(const_cast<float*>(k_host_ptrs[j]));
  for (int j = 0; j < n_k; ++j) e.beta[j] = (float)beta_host[j];
  bool vec = aligned16(out) && aligned16(y0) && numel % 4 == 0;
  for (int j = 0; j < n_k; ++j) vec = vec && aligned16(k_host_ptrs[j]);
  k_epi_only<<<grid_for_elems(numel, sm_count_now()), kStageThreads, 0, st>>>(pp(const_cast<float*>(k_host_ptrs[n_k - 1])), numel, e,
                                                                  vec ? 1 : 0);
  return (int)cudaGetLastError();
}

extern "C" int ndcn_error_ratio_f32(const float* err, const float* y0, const float* y1, double rtol, double atol,
                                    int64_t numel, double* sum_out_dev, ndcn_stream_t s) {
  if (!err || !y0 || !y1 || !sum_out_dev || numel < 0) return NDCN_E_ARG;
  cudaStream_t st = (cudaStream_t)s;
  const int grid = grid_for_elems(numel, sm_count_now());
  double* partials = nullptr;
  CU_TRY(cudaMallocAsync((void**)&partials, sizeof(double) * grid, st));
  k_error_ratio<<<grid, kStageThreads, 0, st>>>(err, y0, y1, numel, (float)rtol, (float)atol, partials);
  k_sum_partials<<<1, kStageThreadsCtl, 0, st>>>(partials, grid, sum_out_dev);
  cudaFreeAsync(partials, st);
  return (int)cudaGetLastError();
}

extern "C" int ndcn_pack_rows_f32(const float* x, const int32_t* idx, int64_t n_idx, int32_t H, float* out,
                                  ndcn_stream_t s) {
  if (n_idx < 0 || H < 1 || (n_idx > 0 && (!x || !idx || !out))) return NDCN_E_ARG;
  if (n_idx == 0) return NDCN_OK;
  const int vec = (H % 4 == 0) && aligned16(x) && aligned16(out);
  const int64_t blocks = (n_idx + kWarpsPerCta - 1) / kWarpsPerCta;
  const int grid = (int)std::min<int64_t>(blocks, (int64_t)sm_count_now() * 16);
  k_pack_rows<<<grid, kStageThreads, 0, (cudaStream_t)s>>>(x, idx, n_idx, H, out, vec);
  return (int)cudaGetLastError();
}

extern "C" int ndcn_pack_cols_f32(const float* x, int64_t n_rows, int32_t H, int32_t block_cols, float* out,
                                  ndcn_stream_t s) {
  if (n_rows < 0 || H < 4 || block_cols < 4 || H % block_cols != 0 || block_cols % 4 != 0) return NDCN_E_ARG;
  if (n_rows == 0) return NDCN_OK;
  if (!x || !out || !aligned16(x) || !aligned16(out)) return NDCN_E_ARG;
  const int64_t n4 = n_rows * (H / 4);
  const int grid = (int)std::min<int64_t>((n4 + kStageThreads - 1) / kStageThreads, (int64_t)sm_count_now() * 16);
  k_pack_cols<<<grid, kStageThreads, 0, (cudaStream_t)s>>>(x, n_rows, H, block_cols, out);
  return (int)cudaGetLastError();
}

extern "C" int ndcn_debug_umma_trace(void* buf_dev) {
  unsigned long long* p = (unsigned long long*)buf_dev;
  return (int)cudaMemcpyToSymbol(g_umma_trace, &p, sizeof(p));
}

extern "C" int ndcn_sizeof(int32_t which) {
  switch (which) {
    case 0: return (int)sizeof(ndcn_rhs_desc_t);
    case 1: return (int)sizeof(ndcn_solve_opts_t);
    case 2: return (int)sizeof(ndcn_solve_stats_t);
    case 3: return (int)sizeof(ndcn_peer_config_t);
    case 4: return (int)sizeof(ndcn_feature_peer_config_t);
    case 5: return (int)sizeof(ndcn_gather_request_t);
    default: return -1;
  }
}

extern "C" const char* ndcn_version(void) { return "ndcn_b200 0.1.0 (sm_100a)"; }
extern "C" int ndcn_sm_arch(void) { return 100; }

/*
 * ndcn_b200.h -- C ABI of libndcn_b200.so (hand-written sm_100a CUDA kernels for the
 * NDCN ODE-integrated graph-convolution hot path).
 *
 * The reference (calvin-zcx/ndcn) is pure Python and has NO native/FFI boundary; its hot
 * path is reached through the Python operator surface (neural_dynamics.ODEFunc / ODEBlock /
 * NDCN, torchdiffeq.odeint).  This header is the boundary a maintainer would bind instead
 * (ctypes stub in INTEGRATION.md): each entry point names the reference call site it
 * replaces (file:line into the upstream tree).
 *
 * Conventions
 *   - every function returns int: 0 ok, <0 invalid argument (NDCN_E_*), >0 a cudaError_t;
 *   - no exceptions, no torch types, plain pointers and sizes;
 *   - all data pointers are BORROWED DEVICE pointers unless the name ends in _host;
 *     the caller (PyTorch) owns every buffer, including the solver workspace;
 *   - every launch goes to the explicit stream argument (a cudaStream_t);
 *   - fp32 state, int32 CSR indices, float64 solver times;
 *   - handles are thread-compatible: one handle, one stream at a time.
 */
#ifndef NDCN_B200_H
#define NDCN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ndcn_stream_t;            /* cudaStream_t */
typedef struct ndcn_graph ndcn_graph_t; /* CSR operator Phi on the device */
typedef struct ndcn_solver ndcn_solver_t;

/* ---- status codes ------------------------------------------------------------------ */
#define NDCN_OK 0
#define NDCN_E_ARG (-1)        /* null pointer / bad size / unsupported combination */
#define NDCN_E_WORKSPACE (-2)  /* workspace too small */
#define NDCN_E_METHOD (-3)     /* method not in {euler, midpoint, rk4, dopri5} */
#define NDCN_E_NONFINITE (-10) /* "non-finite values in state `y`"  (dopri5.py:102)   */
#define NDCN_E_DT_UNDERFLOW (-11) /* "underflow in dt"              (dopri5.py:100)   */
#define NDCN_E_MAX_STEPS (-12) /* "max_num_steps exceeded"          (dopri5.py:89)    */
#define NDCN_E_PEER_TIMEOUT (-13) /* peer push: another rank never reached a barrier (20 s)  */

/* ---- right-hand sides -------------------------------------------------------------- */
enum ndcn_rhs_kind {
  NDCN_RHS_NDCN = 0,   /* relu((Phi x) W^T + b)      neural_dynamics.py:20-39            */
  NDCN_RHS_HEAT = 1,   /* k * (Phi x), Phi = -L      heat_dynamics.py:186-204            */
  NDCN_RHS_GENE = 2,   /* -b x^f + Phi(x^h/(x^h+1))  gene_dynamics.py:186-205            */
  NDCN_RHS_MUTUAL = 3, /* b + x(1-x/k)(x/c-1) + sum_j a_ij x_i x_j/(d + ...)
                          mutualistic_dynamics.py:186-232 (both branches, chosen by H==1) */
  NDCN_RHS_CALLBACK = 4 /* any callable func(t, y): the solver algebra stays fused, the
                          RHS is produced by the host callback   (odeint.py:20 `func`)    */
};

/* flags for NDCN_RHS_NDCN (ODEFunc ctor switches, neural_dynamics.py:9-18) */
#define NDCN_F_NO_GRAPH 1u   /* skip Phi          */
#define NDCN_F_NO_CONTROL 2u /* skip Linear(W,b)  */
#define NDCN_F_NO_RELU 4u    /* (stand-alone SpMM / Linear only; never set by ODEFunc) */

typedef int (*ndcn_rhs_callback_t)(void* user, const float* y_dev, float* k_dev,
                                   const float* t_dev /* fp32 stage time on the device */);
/* multi-GPU hook, called on the launching thread between kernels:
 *   what = 0: `buf` ([n_cols, H], first n_rows rows valid) needs its halo rows
 *             (rows n_rows..n_cols-1) filled from the owning ranks;
 *   what = 1: `buf` points at 2 doubles {sum of squared error ratios, element count}
 *             to be all-reduced (SUM) in place;
 *   what = 2: (only with ndcn_solve_opts_t::gather_mode = NDCN_GATHER_EXTERNAL) `buf` points at a
 *             HOST ndcn_gather_request_t: the callback must leave z = Phi * src for this rank's
 *             rows in `z_dev`, as H / z_block_cols column blocks [n_rows, z_block_cols] stored one
 *             after the other, enqueued on the solver's stream.  This is the feature-sharded
 *             multi-GPU gather: all-to-all the state into column slices, gather every row of the
 *             (replicated) graph on the local slice, all-to-all the result back.              */
typedef int (*ndcn_exchange_callback_t)(void* user, int what, void* buf_dev);

typedef struct ndcn_gather_request {
  const float* src_dev; /* [n_rows, H] gather source, this rank's rows */
  float* z_dev;         /* [H / z_block_cols][n_rows][z_block_cols]    */
} ndcn_gather_request_t;

typedef struct ndcn_rhs_desc {
  int32_t kind;  /* enum ndcn_rhs_kind */
  uint32_t flags;
  int32_t H;     /* state width (columns of x) */
  int32_t reserved;
  const float* W; /* [H,H] row-major, y = x W^T + b (nn.Linear)  -- NDCN only */
  const float* b; /* [H]                                          -- NDCN only */
  /* HEAT: p[0]=k.  GENE: p[0]=b, p[1]=f, p[2]=h.  MUTUAL: p[0..5] = b,k,c,d,e,h */
  float p[8];
  ndcn_rhs_callback_t callback; /* CALLBACK only */
  void* callback_user;
  /* optional, NDCN only: the derived forms of W the kernels consume (W^T for the FP32-FMA GEMMs, the tf32 hi/lo
   * swizzled images of W and W^T for the tcgen05 kernels), produced once by ndcn_prepare_weights_f32 and valid while
   * W is unchanged.  NULL: ndcn_rhs_eval_f32 / ndcn_rhs_vjp_f32 derive them on every call (the solver always
   * derives its own, once per solve).  Training loops evaluate the RHS dozens of times per optimiser step.      */
  const void* prepared;
} ndcn_rhs_desc_t;

/* ---- graph handle ------------------------------------------------------------------
 * Replaces the operator tensor captured by ODEFunc.A / HeatDiffusion.L / *.A
 * (neural_dynamics.py:14, heat_dynamics.py:190): dense or uncoalesced-COO there, int32 CSR
 * here (built once by the host layer).  n_cols >= n_rows: rows [n_rows, n_cols) of every
 * gather source are halo rows owned by other ranks (n_cols == n_rows on one GPU).        */
int ndcn_graph_create(int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* rowptr,
                      const int32_t* col, const float* val, ndcn_graph_t** out);
int ndcn_graph_destroy(ndcn_graph_t* g);

/* ---- stand-alone operators (one RHS evaluation) --------------------------------------- */
/* y = Phi x           torch.sparse.mm / torch.mm at neural_dynamics.py:27-31            */
int ndcn_spmm_f32(const ndcn_graph_t* g, const float* x, float* y, int32_t H, ndcn_stream_t s);
/* out = f(x) for any rhs kind except CALLBACK: ODEFunc.forward neural_dynamics.py:20-39,
 * HeatDiffusion/GeneDynamics/MutualDynamics.forward (files above)                        */
int ndcn_rhs_eval_f32(const ndcn_graph_t* g, const ndcn_rhs_desc_t* rhs, const float* x,
                      float* out, ndcn_stream_t s);

/* derived weight forms for ndcn_rhs_desc_t::prepared (device buffer of ndcn_prepared_weights_bytes(H) bytes,
 * 1024-byte aligned) */
size_t ndcn_prepared_weights_bytes(int32_t H);
int ndcn_prepare_weights_f32(const float* W, int32_t H, void* prepared, ndcn_stream_t s);

/* vjp of one ODEFunc evaluation k = relu((Phi x) W^T + b): what autograd computes through
 * neural_dynamics.py:20-39 for the cotangent gk of k (training loops: heat_dynamics.py:333).
 *   gp = scale * gk where k > 0, else 0                 [n, H]  (dW = gp^T z, db = column sums of gp)
 *   z  = Phi x                                          [n, H]  (not written with NDCN_F_NO_GRAPH: z = x)
 *   gx = Phi^T (gp W)    or gx += ... with accumulate   [n, H]
 * g_t is the handle of Phi^T (pass g again for a symmetric operator).  Single-GPU graphs.        */
int ndcn_rhs_vjp_f32(const ndcn_graph_t* g, const ndcn_graph_t* g_t, const ndcn_rhs_desc_t* rhs,
                     const float* x, const float* gk, float scale, float* gx, int32_t accumulate,
                     float* gp, float* z, ndcn_stream_t s);

/* Parameter gradients of the Linear inside ODEFunc from what ndcn_rhs_vjp_f32 leaves behind
 * (autograd through nn.Linear at neural_dynamics.py:33, training loops heat_dynamics.py:333, dgnn.py:204):
 *   dW[o][i] (+)= sum_r gp[r][o] * z[r][i]        [H, H] row-major, like nn.Linear.weight.grad
 *   db[o]    (+)= sum_r gp[r][o]                  [H]
 * over the n rows, summed in a fixed order (row chunks, then chunk by chunk): run-to-run reproducible.
 * accumulate != 0 adds to dW / db, else overwrites.  db may be NULL.                                    */
int ndcn_weight_grads_f32(const float* gp, const float* z, int64_t n, int32_t H, float* dW, float* db,
                          int32_t accumulate, ndcn_stream_t s);

/* Backward pass of a fixed-grid solve (euler | midpoint | rk4) for the narrow widths of the dynamics scripts (H <= 32,
 * heat_dynamics.py:33) as ONE cooperative launch: what `loss.backward()` computes through odeint's step loop
 * (solvers.py:79-99, rk_common.py:72-78; training loops heat_dynamics.py:317-334), i.e. the discrete adjoint of the scheme.
 *   slab    [n_t, N, H]  the states the forward solve returned (out of ndcn_odeint_f32 on the same grid)
 *   g_slab  [n_t, N, H]  cotangents of those outputs
 *   lam     [N, H]       out: dL/dy0
 *   dW [H,H], db [H]     out: parameter gradients of ODEFunc's Linear (ignored with NDCN_F_NO_CONTROL)
 * Every stage of every step is recomputed from the slab; dW / db are accumulated in registers over all steps and
 * reduced once in a fixed order.  NDCN right-hand sides, single-GPU graphs with at most 16384 rows; anything else
 * returns NDCN_E_ARG (callers fall back to ndcn_rhs_vjp_f32 + ndcn_weight_grads_f32 step by step).            */
int ndcn_fixed_grid_adjoint_small_f32(const ndcn_graph_t* g, const ndcn_graph_t* g_t, const ndcn_rhs_desc_t* rhs,
                                      int32_t method, const double* t_host, int32_t n_t, const float* slab,
                                      const float* g_slab, float* lam, float* dW, float* db, ndcn_stream_t s);

/* ---- solver -------------------------------------------------------------------------- */
enum ndcn_method { NDCN_EULER = 0, NDCN_MIDPOINT = 1, NDCN_RK4 = 2, NDCN_DOPRI5 = 3 };

#define NDCN_O_TERMINAL_ONLY 1u /* out holds only y(t[-1])  (ODEBlock terminal=True)      */
#define NDCN_O_FORCED_DT 2u     /* dopri5: every step accepted, dt = forced_dt            */
#define NDCN_GATHER_LOCAL 0    /* the library gathers Phi x itself (halo rows via exchange what=0)   */
#define NDCN_GATHER_EXTERNAL 1 /* exchange what=2 produces z = Phi x (feature-sharded multi-GPU);
                                  NDCN_RHS_NDCN with W, H in {128, 256} only                         */
#define NDCN_O_TIME_KERNELS 4u  /* bracket every launch with CUDA events on `s`; per-class sums
                                   come back in ndcn_solve_stats_t (profiling aid for bench.py)   */

/* kernel classes for ndcn_solve_stats_t::class_ms / class_launches */
enum ndcn_kernel_class {
  NDCN_K_STAGE = 0,   /* fused RHS + stage-algebra kernels (the dominant kernel) */
  NDCN_K_ALGEBRA = 1, /* stage algebra on a k already in HBM (dopri5 pre-stage, callback RHS) */
  NDCN_K_CONTROL = 2, /* step-size controller / scalar reductions */
  NDCN_K_EMIT = 3,    /* dense output */
  NDCN_K_INIT = 4,    /* initial-step norms */
  NDCN_K_GATHER = 5,  /* gather z = Phi x feeding the tcgen05 GEMM stage kernel */
  NDCN_K_EXCHANGE = 6, /* peer-push barrier / all-reduce kernel (includes the wait for the slowest rank) */
  NDCN_K_CLASSES = 8
};

typedef struct ndcn_solve_opts {
  int32_t method;   /* enum ndcn_method */
  uint32_t flags;
  double rtol, atol;           /* dopri5 only (fixed-grid solvers ignore them, solvers.py:40-41) */
  double forced_dt;            /* with NDCN_O_FORCED_DT */
  int64_t max_num_steps;       /* per output interval, dopri5.py:89; <=0: 2^31-1 */
  ndcn_exchange_callback_t exchange; /* NULL on one GPU */
  void* exchange_user;
  double safety, ifactor, dfactor; /* 0 => dopri5.py:60 defaults .9 / 10 / .2 */
  double first_step;           /* > 0: initial dt given, _select_initial_step skipped (dopri5.py:79-82;
                                  the reference replaces ANY user first_step by 0.01 -- the host layer
                                  reproduces that quirk, the library takes the value as given)        */
  int32_t gather_mode;         /* NDCN_GATHER_LOCAL (0) or NDCN_GATHER_EXTERNAL */
  int32_t z_block_cols;        /* NDCN_GATHER_EXTERNAL: column-block width of z (multiple of 32, divides H) */
  /* fused decoder: NDCN.output_layer = Linear(H -> C) applied to every returned state
   * (neural_dynamics.py:148,159).  With dec_classes = C in 1..8 `out` is [n_t, n_rows, C]
   * ([n_rows, C] with NDCN_O_TERMINAL_ONLY) and the [n_t, n_rows, H] slab is never written.   */
  const float* dec_W;          /* [C, H] row-major (nn.Linear weight), device */
  const float* dec_b;          /* [C] device, may be NULL */
  int32_t dec_classes;         /* 0: no decoder */
  int32_t reserved;
} ndcn_solve_opts_t;

typedef struct ndcn_solve_stats {
  int64_t nfe;        /* RHS evaluations                   */
  int64_t n_accepted; /* accepted steps (all, fixed grid)  */
  int64_t n_rejected;
  int64_t n_launches; /* kernels launched by this solve    */
  double first_step;  /* dopri5 initial dt (misc.py:84)    */
  double last_dt;
  double t_final;     /* end of last accepted step         */
  int32_t status;     /* NDCN_OK or an NDCN_E_* solver error */
  int32_t reserved;
  double class_ms[8];        /* with NDCN_O_TIME_KERNELS: device time per kernel class, ms  */
  int64_t class_launches[8]; /* ... and the launches that time covers (no-op launches of
                                speculative attempts past the end are excluded)            */
} ndcn_solve_stats_t;

/* bytes of device workspace ndcn_solver_create needs for this problem */
size_t ndcn_solver_workspace_bytes(int64_t n_rows, int64_t n_cols, int32_t H, int32_t method);

int ndcn_solver_create(const ndcn_graph_t* g, const ndcn_rhs_desc_t* rhs, int32_t method,
                       void* workspace_dev, size_t workspace_bytes, ndcn_solver_t** out);
int ndcn_solver_destroy(ndcn_solver_t* sv);

/* torchdiffeq.odeint(func, y0, t, rtol, atol, method)    torchdiffeq/_impl/odeint.py:20-76
 *   y0     [n_rows, H] fp32 device
 *   t_host [n_t] float64 HOST, strictly increasing; the caller has already applied the
 *          reference's dtype round trips (ODEBlock rounds vt to fp32 first,
 *          neural_dynamics.py:71; the adaptive driver then promotes, solvers.py:28)
 *   out    [n_t, n_rows, H] (or [n_rows, H] with NDCN_O_TERMINAL_ONLY); out[0] = y0
 *          (last dimension dec_classes instead of H when the fused decoder is on)
 * Enqueues on `s`; synchronises `s` before returning (stats are final on return).      */
int ndcn_odeint_f32(ndcn_solver_t* sv, const float* y0, const double* t_host, int32_t n_t,
                    float* out, const ndcn_solve_opts_t* opts, ndcn_solve_stats_t* stats,
                    ndcn_stream_t s);

/* The same solve as ONE cooperative launch (persistent whole-solve kernel, csrc/small_solver.cuh): for problems whose
 * every tensor stays in L2 -- BASELINE configs 1-2: the 400-node grid of the dynamics scripts (heat_dynamics.py:313-344,
 * 99 Euler steps per model call, 2000 calls) and Cora (dgnn.py:159-237) -- where a launch per stage is nothing but
 * launch latency.  Same arguments, same results element for element (shared device code), grid barriers where the
 * launch boundaries were; the k_i never leave the chip's caches.  Eligible: one GPU, no callback RHS, at most 16384
 * rows and 2^21 state elements, H <= 1024; otherwise NDCN_E_ARG.  ndcn_odeint_f32 picks it by itself when eligible
 * (NDCN_CFG_SMALL_SOLVER = 0 switches that off); ndcn_odeint_staged_f32 always takes the launch-per-stage path.   */
int ndcn_odeint_small_f32(ndcn_solver_t* sv, const float* y0, const double* t_host, int32_t n_t,
                          float* out, const ndcn_solve_opts_t* opts, ndcn_solve_stats_t* stats,
                          ndcn_stream_t s);
int ndcn_odeint_staged_f32(ndcn_solver_t* sv, const float* y0, const double* t_host, int32_t n_t,
                           float* out, const ndcn_solve_opts_t* opts, ndcn_solve_stats_t* stats,
                           ndcn_stream_t s);

/* ---- multi-GPU peer push (new; the reference is single-device) ---------------------------
 * The third exchange scheme of a 1-D row partition, and the one without a collective call on the
 * path: the graph handle is built with a FULL halo (n_cols = all nodes; gather sources are laid out
 * [own rows | rows of rank 0, 1, ... without the own block]), every rank maps the other ranks'
 * workspaces (CUDA IPC, NVLink), and the kernels that produce a gather source (tcgen05 stage
 * kernels, pre-stage algebra) store each new row into the local buffer AND into the halo region of
 * every peer.  A one-block barrier kernel over IPC-shared signal pads orders those stores before
 * the next gather and doubles as the 2-double all-reduce of the dopri5 controller, so the solve needs
 * no exchange hook, no NCCL call and no host synchronisation per step.
 *   pad[r]          device address (in THIS process) of rank r's 4 KB signal pad; pad[rank] = own,
 *                   zero-initialised before the first solve (ndcn_peer_alloc does it)
 *   delta_bytes[r]  distance in bytes from an element of one of this rank's gather-source buffers
 *                   (row i of the own block) to the same element inside rank r's mapped workspace:
 *                   (workspace_r - workspace_own) + halo_row_offset_of_my_block_at_r * H * 4.
 *                   Gather sources sit at the same offset in every rank's workspace (they are sized
 *                   by n_cols = all nodes), so one delta per peer covers Y[2] and YS[2].
 * All ranks must run the same solves in the same order (identical control flow: the controller
 * sees all-reduced sums).  A rank that waits 20 s at a barrier fails with NDCN_E_PEER_TIMEOUT.     */
typedef struct ndcn_peer_config {
  int32_t rank, world;      /* world <= 8 */
  void* pad[8];
  int64_t delta_bytes[8];   /* entry `rank` ignored */
} ndcn_peer_config_t;
int ndcn_solver_set_peers(ndcn_solver_t* sv, const ndcn_peer_config_t* cfg /* NULL: off */);
/* Feature-sharded peer push: the exchange volume of the feature-sharded gather (2 (P-1)/P^2 of the state
 * per RHS and rank instead of (P-1)/P for any all-gather) with the push mechanism instead of two NCCL
 * all-to-alls.  The state, the tcgen05 GEMM and the solver algebra stay row-sharded; rank q additionally
 * owns the column slice [q Hc, (q+1) Hc), Hc = H / world, of EVERY node:
 *   - every kernel that produces a gather source scatters each new row, slice by slice, into the slice
 *     buffers xcs[q] ([N, Hc], rows in global order) of all ranks;
 *   - barrier; every rank gathers ALL rows of `full_graph` on its own slice and stores z = Phi x straight
 *     into block `rank` of the blocked Z ([world][n_local_o][Hc]) of the rank o that owns the row;
 *   - barrier; the tcgen05 stage kernel reads that blocked Z (as with NDCN_GATHER_EXTERNAL).
 * The solver is created on a graph handle with n_rows = n_cols = n_local and no entries (it never
 * gathers itself).  NDCN right-hand sides with W, H in {128, 256}, H / world a power of two >= 32.
 *   pad[r], xcs[r], z[r]  addresses (in THIS process) of rank r's signal pad / slice buffer / blocked Z
 *   row_bounds            block boundaries [world + 1]; rank r owns rows [row_bounds[r], row_bounds[r+1])  */
typedef struct ndcn_feature_peer_config {
  int32_t rank, world;
  void* pad[8];
  void* xcs[8];
  void* z[8];
  int64_t row_bounds[9];
} ndcn_feature_peer_config_t;
int ndcn_solver_set_feature_peers(ndcn_solver_t* sv, const ndcn_graph_t* full_graph,
                                  const ndcn_feature_peer_config_t* cfg /* NULL: off */);
/* cudaMalloc + cudaIpcGetMemHandle (64-byte handle); the first 4 KB are zeroed (signal pad) */
int ndcn_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out);
int ndcn_peer_open(const unsigned char* handle, void** ptr_out); /* cudaIpcOpenMemHandle */
int ndcn_peer_close(void* ptr);
int ndcn_peer_free(void* ptr);
int ndcn_peer_enable_access(int peer_device); /* same-process peers (tests): cudaDeviceEnablePeerAccess */

/* ---- solver algebra as stand-alone kernels (used by tests and by generic callers) ----- */
/* out = y0 + sum_j (dt*beta_j) k_j, reference rounding (misc.py:22-25, rk_common.py:50) */
int ndcn_rk_combine_f32(float* out, const float* y0, const float* const* k_host_ptrs,
                        const double* beta_host, int32_t n_k, float dt, int64_t numel,
                        ndcn_stream_t s);
/* sum over elements of (err/(atol+rtol*max(|y0|,|y1|)))^2 -> *sum_out (double, device)
 * misc.py:146-157                                                                       */
int ndcn_error_ratio_f32(const float* err, const float* y0, const float* y1, double rtol,
                         double atol, int64_t numel, double* sum_out_dev, ndcn_stream_t s);

/* ---- multi-GPU plumbing ------------------------------------------------------------------
 * out[i, :] = x[idx[i], :], i < n_idx: packs the boundary rows another rank needs into a
 * contiguous send buffer (the reference is single-device; new with the 1-D row partition). */
int ndcn_pack_rows_f32(const float* x, const int32_t* idx, int64_t n_idx, int32_t H, float* out,
                       ndcn_stream_t s);

/* out[q][r][c] = x[r][q*block_cols + c]: the [n_rows, H] state as H/block_cols contiguous column
 * blocks -- the send buffer of the feature-sharded exchange (block q goes to peer q).  New.        */
int ndcn_pack_cols_f32(const float* x, int64_t n_rows, int32_t H, int32_t block_cols, float* out,
                       ndcn_stream_t s);

/* ---- library configuration -------------------------------------------------------------
 * Process-wide knobs (also read once from the environment: NDCN_STAGE_IMPL, NDCN_GATHER_CW,
 * NDCN_UMMA_MIN_ROWS).  They select between kernel families that compute the same function;
 * results agree within the fp32 parity tolerance (rtol 1e-4 / atol 1e-6).  New here: the
 * reference has one code path (ATen).                                                      */
#define NDCN_IMPL_AUTO 0 /* tcgen05 path for H in {128,256} from NDCN_CFG_UMMA_MIN_ROWS rows */
#define NDCN_IMPL_SIMT 1 /* fused FP32-FMA stage kernel (gather + GEMM + epilogue in one launch) */
#define NDCN_IMPL_UMMA 2 /* chunk-major gather + tcgen05 3xTF32 GEMM/epilogue kernel */
enum ndcn_config_key {
  NDCN_CFG_STAGE_IMPL = 0,   /* NDCN_IMPL_* */
  NDCN_CFG_GATHER_CW = 1,    /* 0 auto, -1 one pass over full rows, 16/32/64 floats per chunk */
  NDCN_CFG_UMMA_MIN_ROWS = 2,
  NDCN_CFG_GATHER_VERSION = 3, /* chunk-major gather: 1 one row per lane group, 2 persistent CTAs + TMA-staged CSR */
  NDCN_CFG_SMALL_SOLVER = 4   /* 1 (default): ndcn_odeint_f32 runs eligible solves as one persistent kernel; 0: never */
};
int ndcn_config_set(int32_t key, int64_t value);
int64_t ndcn_config_get(int32_t key);

/* Profiling aid: when buf_dev is non-NULL, CTA 0 of every later tcgen05 stage-kernel launch
 * writes a timeline into it: 4 roles (W loader, A producer, MMA issuer, epilogue) x 4096
 * uint64 entries, entry 0 = count, then (event_code << 56 | SM clock).  NULL switches it off. */
int ndcn_debug_umma_trace(void* buf_dev);

/* library / build information */
/* sizeof() of the structs the library reads or writes, so that a binding can check its own declarations:
 * 0 ndcn_rhs_desc_t, 1 ndcn_solve_opts_t, 2 ndcn_solve_stats_t, 3 ndcn_peer_config_t,
 * 4 ndcn_feature_peer_config_t, 5 ndcn_gather_request_t; anything else: -1 */
int ndcn_sizeof(int32_t which);
const char* ndcn_version(void);
int ndcn_sm_arch(void); /* 100: built for sm_100a */

#ifdef __cplusplus
}
#endif
#endif /* NDCN_B200_H */

// tcgen05 stage kernel: k = relu(z W^T + b) fused with the Runge-Kutta stage epilogue, for the
// Linear(H,H) of ODEFunc (neural_dynamics.py:16,32-36) at H in {128, 256}.
//
// z is a dense [N, H] fp32 tensor (the chunk-major gather's output, or x itself with no_graph).
// The contraction runs on the 5th-generation tensor cores as 3xTF32:
//     z W^T  ~=  z_hi W_hi^T + z_lo W_hi^T + z_hi W_lo^T,   x_hi = tf32(x), x_lo = tf32(x - x_hi)
// with fp32 accumulation in TMEM, which keeps the result at fp32-level accuracy (the reference
// computes this Linear in fp32: TF32 is off by default in PyTorch).
//
// One persistent CTA per SM, 14 warps, warp-specialised:
//   warps 0-7   epilogue: tcgen05.ld (thread = row) -> +bias, ReLU -> per-warp SMEM transpose ->
//               lane = column, coalesced 128-byte loads/stores of y0 / k_j / k_out / y_out, the
//               loads of 8 rows are issued before any arithmetic (bytes in flight)
//   warp  8     TMEM allocation; one elected lane issues tcgen05.mma kind::tf32, M=128, N=H, K=8
//   warp  9     one elected lane streams the pre-split, pre-swizzled W image through SMEM with
//               cp.async.bulk (TMA 1-D), one K-atom (32 k-values) of W_hi|W_lo per stage
//   warps 10-13 A producers: coalesced 16-byte loads of the z tile's K-atom -> hi/lo split in
//               registers -> SWIZZLE_128B K-major SMEM (next atom's loads are already in flight)
// Pipelines: full/empty mbarriers per SMEM stage (producers <-> MMA), tmem_full/tmem_empty per
// accumulator (MMA <-> epilogue); two accumulators of H columns each, so the MMAs of tile i+1
// overlap the epilogue of tile i.
#pragma once
#include "ndcn_common.cuh"
#include "stage_kernels.cuh"

namespace ndcn {

constexpr int kUmmaM = 128;          // rows per tile
constexpr int kUmmaEpiWarps = 8;
constexpr int kUmmaMmaWarp = 8;
constexpr int kUmmaLoadWarp = 9;
constexpr int kUmmaProdWarp0 = 10;
constexpr int kUmmaProdWarps = 4;
constexpr int kUmmaThreads = 32 * (kUmmaProdWarp0 + kUmmaProdWarps);  // 448: 128 registers per thread
// (A setmaxnreg split -- 168 registers for the epilogue warpgroups, 88 for the rest -- was measured: the
// epilogue gained nothing from deeper register batches and the A producers, squeezed to 88 registers,
// slowed the operand pipeline from 2.4k to 3.7k cycles per K-atom.)
constexpr int kUmmaStages = 2;
constexpr int kUmmaStagePitch = 32;  // floats; per-warp 32x32 transpose tile, 16-byte chunks XOR-swizzled by row

// Optional device-side timeline of CTA 0 (ndcn_debug_umma_trace): per role a list of
// (event << 56 | clock64) entries; profiling aid, null in production.
constexpr int kTraceRoleCap = 4096;
__device__ unsigned long long* g_umma_trace = nullptr;
struct Tracer {
  unsigned long long* p;
  int n;
  __device__ __forceinline__ Tracer(int role) : p(nullptr), n(0) {
    unsigned long long* t = g_umma_trace;
    if (t != nullptr && role >= 0 && blockIdx.x == 0) p = t + (size_t)role * kTraceRoleCap;
  }
  __device__ __forceinline__ void ev(int code) {
    if (p != nullptr && n < kTraceRoleCap - 1) {
      p[1 + n++] = ((unsigned long long)code << 56) | ((unsigned long long)clock64() & 0x00ffffffffffffffull);
      p[0] = (unsigned long long)n;
    }
  }
};

template <int H>
struct UmmaCfg {
  static constexpr int kAtoms = H / 32;                        // K-atoms of 32 tf32 (128 bytes)
  static constexpr uint32_t kABytes = kUmmaM * 128;            // one of hi / lo
  static constexpr uint32_t kBBytes = H * 128;                 // one of hi / lo
  static constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr uint32_t kOffStaging = kUmmaStages * kStageBytes;
  static constexpr uint32_t kStagingBytes = kUmmaEpiWarps * 32 * kUmmaStagePitch * 4;
  static constexpr uint32_t kOffBias = kOffStaging + kStagingBytes;
  static constexpr uint32_t kOffBars = kOffBias + H * 4;
  static constexpr uint32_t kSmemBytes = kOffBars + 128;
  static constexpr uint32_t kTmemCols = 2 * H;                 // 256 or 512: a power of two
  static constexpr size_t kImageFloats = (size_t)2 * H * H;    // W image: [atom][hi|lo][H rows x 32]
};

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// K-major SWIZZLE_128B operand tile: 8-row groups of 1024 B, row pitch 128 B, the 16-byte chunk
// index is XORed with row % 8.  (Validated on B200 by tests/cuda/umma_tf32x3_probe.cu.)
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  const int chunk = k >> 2;
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4) + (k & 3) * 4);
}

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);  // start address
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l2_last(const void* p) {
  asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
}
// L2 prefetch (evict_last) of the 128-byte line that holds elements [off, off + 32) of every stream the
// epilogue will read, one line per lane and stream: the demand loads a chunk later then see L2 latency,
// not HBM latency.  Measured alternatives (profiles/README.md): evict_normal lines 1.6 % slower, two
// chunks of lead 13 % slower (L2 thrash: 8 streams x 2 chunks x 148 SMs ~ 78 MB), the TMA flavour
// (cp.async.bulk.prefetch.L2) 30-50 % slower, whole-tile prefetch 50 % slower.
__device__ __forceinline__ void epi_prefetch_l2_last(const EpiCtx& c, int64_t off) {
  if (c.mode == EPI_STORE) return;
  prefetch_l2_last(c.y0 + off);
  if (c.mode == EPI_ERR) prefetch_l2_last(c.y1 + off);
#pragma unroll
  for (int j = 0; j < 6; ++j)
    if (j < c.n_prev) prefetch_l2_last(c.kprev[j] + off);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------
// W [H(n), H(k)] row-major (nn.Linear weight) -> image [atom][hi|lo][H rows x 128 B swizzled]:
// exactly the bytes the B operand stages hold, so a stage is one contiguous bulk copy.
// ---------------------------------------------------------------------------------------
template <int H>
__global__ void k_prep_w_image(const float* __restrict__ W, float* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * H) return;
  const int n = i / H, k = i % H;
  const int atom = k >> 5, kk = k & 31;
  const float x = W[i];
  const float hi = tf32_rna(x);
  const float lo = tf32_rna(x - hi);
  const size_t base = (size_t)atom * (2 * H * 32);
  const uint32_t off = sw128_offset(n, kk) >> 2;
  img[base + off] = hi;
  img[base + H * 32 + off] = lo;
}

struct UmmaArgs {
  PtrPair z;          // dense GEMM input [n_rows(+), H] (parity-selected when it is the state itself)
  const float* wimg;  // k_prep_w_image output
  const float* bias;  // [H]
  int64_t n_rows;
  uint32_t flags;     // NDCN_F_NO_RELU
  // z layout: plain row-major [n_rows, H] (z_block_log2 = log2 H), or column blocks
  // [H / bc][n_rows][bc] with bc = 1 << z_block_log2 >= 32 (the feature-sharded multi-GPU gather
  // delivers one [n_rows, H/P] block per peer and nobody has to interleave them)
  int z_block_log2;
  uint32_t deep_batch; // experiment switch (NDCN_UMMA_DBG=256): the other epilogue batch depth
  // row-chunked launches (Z kept L2-resident between its gather and this kernel): tiles [tile_begin, tile_end) only
  // (0, 0 = all tiles); the error partial of CTA b goes to partials[partials_off + b]
  int64_t tile_begin, tile_end;
  int partials_off;
};

// MODE / NPREV: the epilogue mode and the number of earlier stages it reads are compile-time
// constants (one instantiation per Runge-Kutta stage shape): the epilogue is then straight-line code
// with exactly NPREV + 1 loads per element group, which keeps the kernel small enough for the
// instruction cache (a single kernel with every mode inlined four times measured 20-30 % slower).
// kUmmaEpiBatch (1 or 2): groups of 4 rows per load batch; two batches are in flight (software pipeline).
template <int H, int MODE, int NPREV, int kUmmaEpiBatch>
__global__ void __launch_bounds__(kUmmaThreads, 1) k_stage_gemm_umma(UmmaArgs a, EpiArgs e) {
  using Cf = UmmaCfg<H>;
  extern __shared__ __align__(1024) unsigned char smem[];

  EpiCtx c;
  if (!epi_resolve(e, c)) return;
  c.mode = MODE;      // == e.mode, == the resolved n_prev: the launcher picks the instantiation
  c.n_prev = NPREV;
  const int par = e.ctrl ? ((volatile Ctrl*)e.ctrl)->parity : 0;
  const float* __restrict__ z = sel(a.z, par);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* staging = reinterpret_cast<float*>(smem + Cf::kOffStaging);
  float* bias_s = reinterpret_cast<float*>(smem + Cf::kOffBias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cf::kOffBars);
  uint64_t* full = bars;                 // [kUmmaStages]
  uint64_t* empty = bars + 2;            // [kUmmaStages]
  uint64_t* tmem_full = bars + 4;        // [2]
  uint64_t* tmem_empty = bars + 6;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int64_t all_tiles = (a.n_rows + kUmmaM - 1) / kUmmaM;
  const int64_t n_tiles = a.tile_end > 0 ? a.tile_end - a.tile_begin : all_tiles;
  const int64_t tile0 = a.tile_begin + (int64_t)blockIdx.x;  // this CTA's first tile
  const int64_t my_tiles = (n_tiles - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x;  // blockIdx.x < n_tiles

  if (threadIdx.x == 0) {
    for (int s = 0; s < kUmmaStages; ++s) {
      mbar_init(&full[s], kUmmaProdWarps + 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kUmmaEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == kUmmaMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(Cf::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < H; i += kUmmaThreads) bias_s[i] = a.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  double err_acc = 0.0;

  if (warp < kUmmaEpiWarps) {
    // =========================== epilogue ===========================
    const int q = warp & 3;        // TMEM lane quadrant this warp may read
    const int half = warp >> 2;    // which half of the H columns
    float4* st4 = reinterpret_cast<float4*>(staging + warp * 32 * kUmmaStagePitch);
    const bool relu = !(a.flags & NDCN_F_NO_RELU);
    const bool stream_out = (int64_t)a.n_rows * H * 4 > ((int64_t)256 << 20);
    const int rsub = lane >> 3;    // after the transpose: 8 lanes x 16 bytes per row, 4 rows per instruction
    const int cg = lane & 7;
    constexpr int kChunks = H / 2 / 32;
    constexpr int kAhead = 1;      // chunks of L2 prefetch lead
    Tracer tr(warp == 0 && lane == 0 ? 3 : -1000000);
    // lane r prefetches row r of this warp's quadrant for chunk number g of the warp's own sequence
    auto prefetch_chunk = [&](int64_t g) {
      const int64_t ti = g / kChunks;
      if (ti >= my_tiles) return;
      const int64_t row = (tile0 + ti * gridDim.x) * kUmmaM + q * 32 + lane;
      if (row < a.n_rows) epi_prefetch_l2_last(c, row * H + half * (H / 2) + (int)(g % kChunks) * 32);
    };
    for (int g = 0; g < kAhead; ++g) prefetch_chunk(g);
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int64_t tile = tile0 + i * gridDim.x;
      const int acc = (int)(i & 1);
      const int64_t row_base = tile * kUmmaM + q * 32;
      mbar_wait(&tmem_full[acc], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      tr.ev(6);
#pragma unroll 1
      for (int cc = 0; cc < kChunks; ++cc) {
        const int c0 = half * (H / 2) + cc * 32;
        prefetch_chunk(i * kChunks + cc + kAhead);
        // the stage-algebra loads run one batch ahead of the arithmetic (two register sets): the first
        // batch of a chunk is already in flight while the accumulators are read and transposed
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * cg);
        auto load_batch = [&](int it0, EpiIn<4>(&in)[kUmmaEpiBatch]) {
#pragma unroll
          for (int u = 0; u < kUmmaEpiBatch; ++u) {
            const int64_t row = row_base + (it0 + u) * 4 + rsub;
            if (row < a.n_rows) epi_load<4>(c, row * H + c0 + 4 * cg, in[u]);
          }
        };
        auto math_batch = [&](int it0, const EpiIn<4>(&in)[kUmmaEpiBatch]) {
#pragma unroll
          for (int u = 0; u < kUmmaEpiBatch; ++u) {
            const int r = (it0 + u) * 4 + rsub;
            const int64_t row = row_base + r;
            if (row < a.n_rows) {
              const float4 kk = st4[r * 8 + (cg ^ (r & 7))];
              float kv[4] = {kk.x + b4.x, kk.y + b4.y, kk.z + b4.z, kk.w + b4.w};
              if (relu) {
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) kv[e4] = fmaxf(kv[e4], 0.f);
              }
              epi_math<4>(c, row * H + c0 + 4 * cg, kv, in[u], err_acc, stream_out);
            }
          }
        };
        EpiIn<4> in_a[kUmmaEpiBatch], in_b[kUmmaEpiBatch];
        load_batch(0, in_a);

        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * H + c0), v);
        if (cc == kChunks - 1) {
          // every tcgen05.ld of this accumulator has completed: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        // thread = row: park the 32 accumulators of this row in the transpose tile
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        // lane = (row rsub of a group of 4, columns c0 + 4 cg .. +3): 128-byte segments, 16 bytes per lane
#pragma unroll 1
        for (int it0 = 0; it0 < 8; it0 += 2 * kUmmaEpiBatch) {
          load_batch(it0 + kUmmaEpiBatch, in_b);
          math_batch(it0, in_a);
          if (it0 + 2 * kUmmaEpiBatch < 8) load_batch(it0 + 2 * kUmmaEpiBatch, in_a);
          math_batch(it0 + kUmmaEpiBatch, in_b);
        }
        __syncwarp();  // the transpose tile is rewritten by the next chunk
        tr.ev(7);
      }
    }
  } else if (warp == kUmmaMmaWarp) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // instruction descriptor: D = f32, A = B = tf32, both K-major, N = H, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | ((128u >> 4) << 24);
      Tracer tr(2);
      uint32_t it = 0;
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int acc = (int)(i & 1);
        mbar_wait(&tmem_empty[acc], (uint32_t)(((i >> 1) & 1) ^ 1));
        tc_fence_after();
        tr.ev(5);
        const uint32_t d_tmem = tmem + (uint32_t)(acc * H);
        for (int atom = 0; atom < Cf::kAtoms; ++atom, ++it) {
          const int s = it % kUmmaStages;
          mbar_wait(&full[s], (it / kUmmaStages) & 1);
          tc_fence_after();
          tr.ev(3);
          const uint32_t sa = smem_u32(smem + s * Cf::kStageBytes);
          const uint32_t a_hi = sa, a_lo = sa + Cf::kABytes;
          const uint32_t b_hi = sa + 2 * Cf::kABytes, b_lo = b_hi + Cf::kBBytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 4 x (K = 8 tf32 = 32 bytes) per 128-byte atom
            const uint64_t dah = umma_desc_sw128(a_hi + kk * 32), dal = umma_desc_sw128(a_lo + kk * 32);
            const uint64_t dbh = umma_desc_sw128(b_hi + kk * 32), dbl = umma_desc_sw128(b_lo + kk * 32);
            umma_tf32(d_tmem, dah, dbh, idesc, (atom | kk) ? 1u : 0u);
            umma_tf32(d_tmem, dal, dbh, idesc, 1u);
            umma_tf32(d_tmem, dah, dbl, idesc, 1u);
          }
          umma_commit(&empty[s]);  // implies tcgen05.fence::before_thread_sync
          tr.ev(4);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else if (warp == kUmmaLoadWarp) {
    // =========================== W image loader (TMA 1-D) ===========================
    if (lane == 0) {
      Tracer tr(0);
      uint32_t it = 0;
      for (int64_t i = 0; i < my_tiles; ++i) {
        for (int atom = 0; atom < Cf::kAtoms; ++atom, ++it) {
          const int s = it % kUmmaStages;
          mbar_wait(&empty[s], ((it / kUmmaStages) & 1) ^ 1);
          tr.ev(0);
          mbar_arrive_expect_tx(&full[s], 2 * Cf::kBBytes);
          bulk_g2s(smem + s * Cf::kStageBytes + 2 * Cf::kABytes, a.wimg + (size_t)atom * (2 * H * 32), 2 * Cf::kBBytes,
                   &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp >= kUmmaProdWarp0) {
    // =========================== A producers ===========================
    const int pw = warp - kUmmaProdWarp0;  // rows [32 pw, 32 pw + 32) of the tile
    const int rsub = lane >> 3;            // 4 rows per load instruction
    const int ch = lane & 7;               // 16-byte chunk of the 128-byte atom row
    const int64_t total = my_tiles * Cf::kAtoms;
    float4 cur[8], nxt[8];
    const int bl = a.z_block_log2;
    const int64_t blk_stride = a.n_rows << bl;  // elements per column block
    // first element of the 32-column K-atom `atom` of row `row`
    auto z_atom = [&](int64_t row, int atom) -> const float* {
      const int col0 = atom * 32;
      return z + (int64_t)(col0 >> bl) * blk_stride + (row << bl) + (col0 & ((1 << bl) - 1));
    };
    auto load_atom = [&](int64_t it, float4(&dst)[8]) {
      const int64_t tile = tile0 + (it / Cf::kAtoms) * gridDim.x;
      const int atom = (int)(it % Cf::kAtoms);
      const int64_t row0 = tile * kUmmaM + pw * 32 + rsub;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t row = row0 + u * 4;
        dst[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < a.n_rows) dst[u] = __ldcs(reinterpret_cast<const float4*>(z_atom(row, atom) + ch * 4));
      }
    };
    // lane r of producer warp pw prefetches (L2) the 128-byte atom row r of its 32 rows
    auto prefetch_atom = [&](int64_t it) {
      if (it >= total) return;
      const int64_t tile = tile0 + (it / Cf::kAtoms) * gridDim.x;
      const int64_t row = tile * kUmmaM + pw * 32 + lane;
      if (row < a.n_rows) prefetch_l2(z_atom(row, (int)(it % Cf::kAtoms)));
    };
    constexpr int kAheadA = 4;
    Tracer tr(pw == 0 && lane == 0 ? 1 : -1000000);
    for (int it = 1; it < kAheadA; ++it) prefetch_atom(it);
    if (total > 0) load_atom(0, cur);
    for (int64_t it = 0; it < total; ++it) {
      prefetch_atom(it + kAheadA);
      if (it + 1 < total) load_atom(it + 1, nxt);
      const int s = (int)(it % kUmmaStages);
      mbar_wait(&empty[s], (uint32_t)(((it / kUmmaStages) & 1) ^ 1));
      tr.ev(1);
      unsigned char* a_hi = smem + s * Cf::kStageBytes;
      unsigned char* a_lo = a_hi + Cf::kABytes;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = pw * 32 + u * 4 + rsub;
        const uint32_t off = sw128_offset(r, ch * 4);
        const float4 x = cur[u];
        float4 hi, lo;
        hi.x = tf32_rna(x.x); hi.y = tf32_rna(x.y); hi.z = tf32_rna(x.z); hi.w = tf32_rna(x.w);
        lo.x = tf32_rna(x.x - hi.x); lo.y = tf32_rna(x.y - hi.y); lo.z = tf32_rna(x.z - hi.z); lo.w = tf32_rna(x.w - hi.w);
        *reinterpret_cast<float4*>(a_hi + off) = hi;
        *reinterpret_cast<float4*>(a_lo + off) = lo;
      }
      fence_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      tr.ev(2);
#pragma unroll
      for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
    }
  }

  // ---- teardown: error partial of this CTA, TMEM release ----
  tc_fence_before();
  __syncthreads();
  if (MODE == EPI_ERR) {
    double* red = reinterpret_cast<double*>(staging);  // the transpose tiles are idle now
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err_acc += __shfl_xor_sync(0xffffffffu, err_acc, o);
    if (lane == 0) red[warp] = err_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double v = 0.0;
      for (int w = 0; w < kUmmaThreads / 32; ++w) v += red[w];
      e.partials[a.partials_off + blockIdx.x] = v;
    }
  }
  if (warp == kUmmaMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(Cf::kTmemCols) : "memory");
  }
}

}  // namespace ndcn

// Probe (round 2): does a COLUMN-BLOCKED gather source [H/16][N][16] make the sparse gather z = Phi x
// L2-resident on B200?  One slab is N*64 B (64 MB at N = 1M) and contiguous, so -- unlike a 16/32-column
// chunk of a row-major [N,256] state, whose 128-byte lines span 128 MB -- it fits the 126 MB L2.
//
//   mode 0  row-major source, one warp per row, 2 x 16 B per lane and entry (the round-1 kernel's shape)
//   mode 1  slab source, 4 lanes per row (64 B per entry), slab-major grid, U entries in flight per lane
//   mode 2  as 1, x loads with L2::evict_last, (col,val) streaming (evict_first)
// Graph: binary file written by scripts/exp_slab_probe.py (int64 n, int64 nnz, rowptr, col, val).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o slab_probe slab_gather_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int H = 256;
constexpr int BW = 16;           // slab width (floats)
constexpr int NSLAB = H / BW;
constexpr int kThreads = 256;
constexpr int kLongRow = 256;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));       \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ld_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}

// ---- mode 0: row-major, warp per row ---------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 4) k_rowmajor(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                           const float* __restrict__ val, const float* __restrict__ x,
                                                           float* __restrict__ z, int n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + warp;
  if (row >= n) return;
  const int start = rowptr[row], end = rowptr[row + 1];
  float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
  const float* xl = x + lane * 4;
  for (int base = start; base < end; base += 32) {
    int my_c = 0;
    float my_v = 0.f;
    if (base + lane < end) {
      my_c = __ldcs(col + base + lane);
      my_v = __ldcs(val + base + lane);
    }
    const int cnt = min(32, end - base);
    int j = 0;
    for (; j + 4 <= cnt; j += 4) {
      float4 p[4], q[4];
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = __shfl_sync(0xffffffffu, my_c, j + u);
        v[u] = __shfl_sync(0xffffffffu, my_v, j + u);
        p[u] = *reinterpret_cast<const float4*>(xl + (long)c * H);
        q[u] = *reinterpret_cast<const float4*>(xl + (long)c * H + 128);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a0.x = fmaf(v[u], p[u].x, a0.x); a0.y = fmaf(v[u], p[u].y, a0.y); a0.z = fmaf(v[u], p[u].z, a0.z); a0.w = fmaf(v[u], p[u].w, a0.w);
        a1.x = fmaf(v[u], q[u].x, a1.x); a1.y = fmaf(v[u], q[u].y, a1.y); a1.z = fmaf(v[u], q[u].z, a1.z); a1.w = fmaf(v[u], q[u].w, a1.w);
      }
    }
    for (; j < cnt; ++j) {
      const int c = __shfl_sync(0xffffffffu, my_c, j);
      const float v = __shfl_sync(0xffffffffu, my_v, j);
      const float4 p = *reinterpret_cast<const float4*>(xl + (long)c * H);
      const float4 q = *reinterpret_cast<const float4*>(xl + (long)c * H + 128);
      a0.x = fmaf(v, p.x, a0.x); a0.y = fmaf(v, p.y, a0.y); a0.z = fmaf(v, p.z, a0.z); a0.w = fmaf(v, p.w, a0.w);
      a1.x = fmaf(v, q.x, a1.x); a1.y = fmaf(v, q.y, a1.y); a1.z = fmaf(v, q.z, a1.z); a1.w = fmaf(v, q.w, a1.w);
    }
  }
  __stcs(reinterpret_cast<float4*>(z + row * H + lane * 4), a0);
  __stcs(reinterpret_cast<float4*>(z + row * H + 128 + lane * 4), a1);
}

// ---- modes 1/2: slab source ------------------------------------------------------------------
// blockIdx.x = slab * (n_rb + n_long) + b ; b < n_rb: 64 rows (4 lanes each); else one long row
template <int U, bool HINT, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_slab(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                          const float* __restrict__ val, const float* __restrict__ x,
                                                          float* __restrict__ z, int n, int n_rb, int n_long,
                                                          const int* __restrict__ long_rows) {
  __shared__ __align__(16) float s_part[(kThreads / 4) * BW];
  const int bpc = n_rb + n_long;
  const int slab = blockIdx.x / bpc;
  const int b = blockIdx.x - slab * bpc;
  const int sub = threadIdx.x & 3;
  const float* __restrict__ xs = x + (size_t)slab * n * BW + sub * 4;
  float* __restrict__ zs = z + (size_t)slab * n * BW + sub * 4;
  uint64_t pol = 0;
  if constexpr (HINT) pol = policy_evict_last();
  auto ldx = [&](int c) -> float4 {
    const float4* p = reinterpret_cast<const float4*>(xs + (size_t)c * BW);
    if constexpr (HINT) return ld_hint(p, pol);
    else return *p;
  };
  if (b < n_rb) {
    const long row = (long)b * (kThreads / 4) + (threadIdx.x >> 2);
    if (row >= n) return;
    const int start = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    if (n_long > 0 && end - start > kLongRow) return;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int k = start; k < end; k += U) {
      float4 xv[U];
      float vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k + u < end) {
          const int c = HINT ? __ldcs(col + k + u) : __ldg(col + k + u);
          vv[u] = HINT ? __ldcs(val + k + u) : __ldg(val + k + u);
          xv[u] = ldx(c);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k + u < end) {
          acc.x = fmaf(vv[u], xv[u].x, acc.x); acc.y = fmaf(vv[u], xv[u].y, acc.y);
          acc.z = fmaf(vv[u], xv[u].z, acc.z); acc.w = fmaf(vv[u], xv[u].w, acc.w);
        }
      }
    }
    __stcs(reinterpret_cast<float4*>(zs + (size_t)row * BW), acc);
  } else {
    const long row = long_rows[b - n_rb];
    const int start = rowptr[row], end = rowptr[row + 1];
    const int g = threadIdx.x >> 2;
    constexpr int G = kThreads / 4;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int idx = start + g; idx < end; idx += 4 * G) {
      float4 xv[4];
      float vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int id = idx + u * G;
        vv[u] = 0.f;
        xv[u] = make_float4(0, 0, 0, 0);
        if (id < end) {
          vv[u] = __ldg(val + id);
          xv[u] = ldx(__ldg(col + id));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x = fmaf(vv[u], xv[u].x, acc.x); acc.y = fmaf(vv[u], xv[u].y, acc.y);
        acc.z = fmaf(vv[u], xv[u].z, acc.z); acc.w = fmaf(vv[u], xv[u].w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(s_part + g * BW + sub * 4) = acc;
    __syncthreads();
    if (threadIdx.x < 4) {
      float4 t = make_float4(0, 0, 0, 0);
      for (int gg = 0; gg < G; ++gg) {
        const float4 p = *reinterpret_cast<const float4*>(s_part + gg * BW + sub * 4);
        t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
      }
      __stcs(reinterpret_cast<float4*>(zs + (size_t)row * BW), t);
    }
  }
}


// ---- mode 10+: slab source, CSR slice of the CTA's row block staged in shared memory (coalesced, read once per
// CTA instead of once per lane group), lane groups pull rows from a CTA-local counter (long rows of a power-law
// graph no longer idle the other groups), U independent 16-byte loads per lane in flight.
template <int U, int ROWS, int CAP, int MINB, int POL>
__global__ void __launch_bounds__(kThreads, MINB) k_slab3(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                           const float* __restrict__ val, const float* __restrict__ x,
                                                           float* __restrict__ z, int n, int n_rb, int n_long,
                                                           const int* __restrict__ long_rows) {
  __shared__ __align__(16) float s_part[(kThreads / 4) * BW];
  __shared__ int s_col[CAP];
  __shared__ float s_val[CAP];
  __shared__ int s_rp[ROWS + 1];
  __shared__ int s_next;
  const int bpc = n_rb + n_long;
  const int slab = blockIdx.x / bpc;
  const int b = blockIdx.x - slab * bpc;
  const int sub = threadIdx.x & 3;
  const int g = threadIdx.x >> 2;
  constexpr int G = kThreads / 4;
  const float* __restrict__ xs = x + (size_t)slab * n * BW + sub * 4;
  float* __restrict__ zs = z + (size_t)slab * n * BW + sub * 4;
  uint64_t pol = 0;
  if constexpr (POL == 1) pol = policy_evict_last();
  auto ldx = [&](int c) -> float4 {
    const float4* p = reinterpret_cast<const float4*>(xs + (size_t)c * BW);
    if constexpr (POL == 1) return ld_hint(p, pol);
    else return *p;
  };
  if (b < n_rb) {
    const int r0 = b * ROWS;
    const int nr = min(ROWS, n - r0);
    for (int i = threadIdx.x; i <= nr; i += kThreads) s_rp[i] = __ldg(rowptr + r0 + i);
    if (threadIdx.x == 0) s_next = G;
    __syncthreads();
    const int e0 = s_rp[0];
    const int cnt = min(s_rp[nr] - e0, CAP);
    for (int i = threadIdx.x; i < cnt; i += kThreads) {
      s_col[i] = __ldcs(col + e0 + i);
      s_val[i] = __ldcs(val + e0 + i);
    }
    __syncthreads();
    const unsigned gmask = 0xFu << ((threadIdx.x & 31) & ~3);
    int row = g;
    while (row < nr) {
      const int start = s_rp[row] - e0, end = s_rp[row + 1] - e0;
      if (!(n_long > 0 && end - start > kLongRow)) {
        float4 acc = make_float4(0, 0, 0, 0);
        for (int k = start; k < end; k += U) {
          float4 xv[U];
          float vv[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int idx = k + u;
            if (idx < end) {
              int c;
              if (idx < CAP) { c = s_col[idx]; vv[u] = s_val[idx]; }
              else { c = __ldg(col + e0 + idx); vv[u] = __ldg(val + e0 + idx); }
              xv[u] = ldx(c);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (k + u < end) {
              acc.x = fmaf(vv[u], xv[u].x, acc.x); acc.y = fmaf(vv[u], xv[u].y, acc.y);
              acc.z = fmaf(vv[u], xv[u].z, acc.z); acc.w = fmaf(vv[u], xv[u].w, acc.w);
            }
          }
        }
        __stcs(reinterpret_cast<float4*>(zs + (size_t)(r0 + row) * BW), acc);
      }
      int nxt = 0;
      if (sub == 0) nxt = atomicAdd(&s_next, 1);
      row = __shfl_sync(gmask, nxt, (threadIdx.x & 31) & ~3);
    }
  } else {
    const long row = long_rows[b - n_rb];
    const int start = rowptr[row], end = rowptr[row + 1];
    float4 acc = make_float4(0, 0, 0, 0);
    for (int idx = start + g; idx < end; idx += 4 * G) {
      float4 xv[4];
      float vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int id = idx + u * G;
        vv[u] = 0.f;
        xv[u] = make_float4(0, 0, 0, 0);
        if (id < end) {
          vv[u] = __ldg(val + id);
          xv[u] = ldx(__ldg(col + id));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x = fmaf(vv[u], xv[u].x, acc.x); acc.y = fmaf(vv[u], xv[u].y, acc.y);
        acc.z = fmaf(vv[u], xv[u].z, acc.z); acc.w = fmaf(vv[u], xv[u].w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(s_part + g * BW + sub * 4) = acc;
    __syncthreads();
    if (threadIdx.x < 4) {
      float4 t = make_float4(0, 0, 0, 0);
      for (int gg = 0; gg < G; ++gg) {
        const float4 p = *reinterpret_cast<const float4*>(s_part + gg * BW + sub * 4);
        t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
      }
      __stcs(reinterpret_cast<float4*>(zs + (size_t)row * BW), t);
    }
  }
}

__global__ void k_to_slab(const float* __restrict__ x, float* __restrict__ xs, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index in row-major
  if (i >= n * (H / 4)) return;
  const long row = i / (H / 4);
  const int c4 = (int)(i % (H / 4));
  const int slab = c4 / (BW / 4), cc = c4 % (BW / 4);
  reinterpret_cast<float4*>(xs)[((size_t)slab * n + row) * (BW / 4) + cc] = reinterpret_cast<const float4*>(x)[i];
}

__global__ void k_fill(float* x, long n, unsigned seed) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = (unsigned)i * 2654435761u ^ seed;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  x[i] = (float)(h & 0xffff) / 65536.f - 0.5f;
}

__global__ void k_maxdiff(const float* __restrict__ zr, const float* __restrict__ zs, long n, float* out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // element index in row-major
  if (i >= n * H) return;
  const long row = i / H;
  const int c = (int)(i % H);
  const float d = fabsf(zr[i] - zs[((size_t)(c / BW) * n + row) * BW + (c % BW)]);
  if (d > 1e-5f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(d));
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "gpurun_out/graph.bin";
  const int only = argc > 2 ? atoi(argv[2]) : -1;  // run only this mode (for ncu)
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); return 1; }
  int64_t n = 0, nnz = 0;
  if (fread(&n, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1) return 1;
  std::vector<int> rp(n + 1), col(nnz);
  std::vector<float> val(nnz);
  if (fread(rp.data(), 4, n + 1, f) != (size_t)n + 1 || fread(col.data(), 4, nnz, f) != (size_t)nnz ||
      fread(val.data(), 4, nnz, f) != (size_t)nnz) return 1;
  fclose(f);
  std::vector<int> longs;
  for (int64_t r = 0; r < n; ++r)
    if (rp[r + 1] - rp[r] > kLongRow) longs.push_back((int)r);
  printf("n=%ld nnz=%ld long_rows=%zu\n", (long)n, (long)nnz, longs.size());

  int *d_rp, *d_col, *d_long;
  float *d_val, *x, *xs, *zr, *zs, *d_diff;
  CK(cudaMalloc(&d_rp, 4 * (n + 1)));
  CK(cudaMalloc(&d_col, 4 * nnz));
  CK(cudaMalloc(&d_val, 4 * nnz));
  CK(cudaMalloc(&d_long, 4 * (longs.size() + 1)));
  CK(cudaMalloc(&x, 4 * n * H));
  CK(cudaMalloc(&xs, 4 * n * H));
  CK(cudaMalloc(&zr, 4 * n * H));
  CK(cudaMalloc(&zs, 4 * n * H));
  CK(cudaMalloc(&d_diff, 4));
  CK(cudaMemcpy(d_rp, rp.data(), 4 * (n + 1), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_col, col.data(), 4 * nnz, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_val, val.data(), 4 * nnz, cudaMemcpyHostToDevice));
  if (!longs.empty()) CK(cudaMemcpy(d_long, longs.data(), 4 * longs.size(), cudaMemcpyHostToDevice));
  k_fill<<<(unsigned)((n * H + 255) / 256), 256>>>(x, n * H, 12345u);
  k_to_slab<<<(unsigned)((n * (H / 4) + 255) / 256), 256>>>(x, xs, n);
  CK(cudaDeviceSynchronize());

  const int n_long = (int)longs.size();
  const int n_rb = (int)((n + 63) / 64);
  const unsigned grid_slab = (unsigned)(NSLAB * (n_rb + n_long));
  const int reps = 5;
  const double gb_alg = (2.0 * 4 * n * H + 8.0 * nnz + 4.0 * (n + 1)) / 1e9;

  auto report = [&](const char* name, float ms) {
    printf("%-34s %.3f ms   algorithmic %.2f GB -> %.0f GB/s\n", name, ms, gb_alg, gb_alg / (ms * 1e-3));
  };
  if (only < 0 || only == 0) {
    float ms = time_ms([&] { k_rowmajor<<<(unsigned)((n + 7) / 8), kThreads>>>(d_rp, d_col, d_val, x, zr, (int)n); }, reps);
    report("mode0 row-major warp/row", ms);
  }
#define RUN_SLAB(U, HINT, MINB, tag)                                                                                   \
  {                                                                                                                    \
    float ms = time_ms([&] { k_slab<U, HINT, MINB><<<grid_slab, kThreads>>>(d_rp, d_col, d_val, xs, zs, (int)n, n_rb,  \
                                                                            n_long, d_long); }, reps);               \
    report(tag, ms);                                                                                                   \
  }
  if (only < 0 || only == 1) RUN_SLAB(4, false, 4, "mode1 slab U=4 plain 4cta");
  if (only < 0 || only == 2) RUN_SLAB(4, true, 4, "mode2 slab U=4 hints 4cta");
  if (only < 0 || only == 3) RUN_SLAB(8, true, 4, "mode3 slab U=8 hints 4cta");
  if (only < 0 || only == 4) RUN_SLAB(8, true, 5, "mode4 slab U=8 hints 5cta");
  if (only < 0 || only == 5) RUN_SLAB(12, true, 5, "mode5 slab U=12 hints 5cta");
  if (only < 0 || only == 6) RUN_SLAB(4, true, 6, "mode6 slab U=4 hints 6cta");

#define RUN_SLAB3(U, ROWS, CAP, MINB, POL, tag)                                                                       \
  {                                                                                                                    \
    const int nrb3 = (int)((n + ROWS - 1) / ROWS);                                                                     \
    const unsigned grid3 = (unsigned)(NSLAB * (nrb3 + n_long));                                                        \
    float ms = time_ms([&] { k_slab3<U, ROWS, CAP, MINB, POL><<<grid3, kThreads>>>(d_rp, d_col, d_val, xs, zs, (int)n, \
                                                                                 nrb3, n_long, d_long); }, reps);    \
    report(tag, ms);                                                                                                   \
  }
  if (only < 0 || only == 10) RUN_SLAB3(4, 128, 2048, 4, 0, "mode10 slab3 U=4 R=128 4cta");
  if (only < 0 || only == 11) RUN_SLAB3(8, 128, 2048, 4, 0, "mode11 slab3 U=8 R=128 4cta");
  if (only < 0 || only == 12) RUN_SLAB3(8, 256, 4096, 4, 0, "mode12 slab3 U=8 R=256 4cta");
  if (only < 0 || only == 13) RUN_SLAB3(8, 128, 2048, 4, 1, "mode13 slab3 U=8 R=128 evict_last");
  if (only < 0 || only == 14) RUN_SLAB3(4, 128, 2048, 6, 0, "mode14 slab3 U=4 R=128 6cta");
  if (only < 0 || only == 15) RUN_SLAB3(6, 192, 3072, 5, 0, "mode15 slab3 U=6 R=192 5cta");
  if (only < 0 || only == 16) RUN_SLAB3(8, 384, 4096, 3, 0, "mode16 slab3 U=8 R=384 3cta");
  if (only < 0 || only == 17) RUN_SLAB3(2, 128, 2048, 8, 0, "mode17 slab3 U=2 R=128 8cta");
  if (only < 0 || only == 18) RUN_SLAB3(4, 64, 1024, 6, 0, "mode18 slab3 U=4 R=64 6cta");
  if (only < 0 || only == 19) RUN_SLAB3(4, 128, 2048, 6, 1, "mode19 slab3 U=4 R=128 6cta evict_last");
  if (only < 0) {
    CK(cudaMemset(d_diff, 0, 4));
    k_rowmajor<<<(unsigned)((n + 7) / 8), kThreads>>>(d_rp, d_col, d_val, x, zr, (int)n);
    k_slab3<8, 128, 2048, 4, 0><<<(unsigned)(NSLAB * ((int)((n + 127) / 128) + n_long)), kThreads>>>(d_rp, d_col, d_val, xs, zs, (int)n, (int)((n + 127) / 128), n_long, d_long);
    k_maxdiff<<<(unsigned)((n * H + 255) / 256), 256>>>(zr, zs, n, d_diff);
    float diff;
    CK(cudaMemcpy(&diff, d_diff, 4, cudaMemcpyDeviceToHost));
    printf("max |z_rowmajor - z_slab| above 1e-5: %g\n", diff);
  }
  return 0;
}

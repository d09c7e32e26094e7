// Probe: does HBM3e deliver less bandwidth when a [N, 256] fp32 tensor is streamed in 128-byte
// pieces at a 1 KB stride (the access shape of a 32-column chunk of 32 rows: what the tcgen05
// epilogue and the per-atom A producers issue) than in contiguous 4 KB pieces?
//   mode 0: warp reads 4 KB contiguous per stream and step (rows r..r+3, all 256 columns)
//   mode 1: warp reads 32 rows x 128 B per stream and step, chunk-major inside 128-row tiles
// S read streams + 1 write stream, persistent grid, 8 independent 16-byte loads per lane in flight.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_pattern_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int H = 256;

template <int S, int MODE>
__global__ void __launch_bounds__(256) k_stream(const float* __restrict__ base, float* __restrict__ out, long n_rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long n_tiles = n_rows / 128;
  const long stream_elems = n_rows * H;
  for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // 8 warps: quadrant q = warp & 3 (32 rows), half = warp >> 2 (128 columns) -- as in the epilogue
    const int q = warp & 3, half = warp >> 2;
    const long row0 = tile * 128 + q * 32;
    for (int step = 0; step < 4; ++step) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        long off;
        if (MODE == 0) {
          // contiguous: the warp's 32 rows x 128 columns region = 16 KB, walked 512 B at a time;
          // row = 8 per step, 128 columns = 512 B contiguous per row, 1 row per instruction
          const int r = step * 8 + it;
          off = (row0 + r) * H + half * 128 + lane * 4;
        } else {
          // chunk-major: step = 32-column chunk, it = group of 4 rows, lane = (row in group, 16-byte piece)
          const int r = it * 4 + (lane >> 3);
          off = (row0 + r) * H + half * 128 + step * 32 + (lane & 7) * 4;
        }
        float4 v[S];
#pragma unroll
        for (int s = 0; s < S; ++s) v[s] = __ldcs(reinterpret_cast<const float4*>(base + s * stream_elems + off));
#pragma unroll
        for (int s = 0; s < S; ++s) {
          acc.x += v[s].x; acc.y += v[s].y; acc.z += v[s].z; acc.w += v[s].w;
        }
        __stcs(reinterpret_cast<float4*>(out + off), acc);
      }
    }
  }
}

template <int S, int MODE>
float run(const float* base, float* out, long n_rows) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_stream<S, MODE><<<148 * 4, 256>>>(base, out, n_rows);
  cudaEventRecord(a);
  for (int i = 0; i < 3; ++i) k_stream<S, MODE><<<148 * 4, 256>>>(base, out, n_rows);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / 3;
}

int main() {
  const long n_rows = 1000064;  // multiple of 128
  const size_t bytes = (size_t)n_rows * H * 4;
  float *base, *out;
  if (cudaMalloc(&base, bytes * 6) != cudaSuccess || cudaMalloc(&out, bytes) != cudaSuccess) { printf("alloc failed\n"); return 2; }
  cudaMemset(base, 0, bytes * 6);
  printf("streams  contiguous-512B   chunked-128B@1KB   (GB/s incl. the write stream)\n");
#define ROW(S) { float t0 = run<S, 0>(base, out, n_rows), t1 = run<S, 1>(base, out, n_rows); \
  printf("  %d      %7.3f ms %6.0f   %7.3f ms %6.0f\n", S, t0, (S + 1) * bytes / t0 / 1e6, t1, (S + 1) * bytes / t1 / 1e6); }
  ROW(1) ROW(2) ROW(4) ROW(6)
  cudaError_t e = cudaDeviceSynchronize();
  printf(e == cudaSuccess ? "PROBE OK\n" : "CUDA error\n");
  return e == cudaSuccess ? 0 : 1;
}

// Stand-alone probe for the tcgen05 (UMMA) building blocks used by the fused RHS kernel:
// smem descriptors for K-major SWIZZLE_128B tf32 operands, the kind::tf32 instruction descriptor,
// TMEM allocation, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b epilogue, and the 3xTF32 split
// (hi*hi + lo*hi + hi*lo) that restores fp32-level accuracy.
//   C[128,256] = A[128,256] * W[256,256]^T      (one CTA, 128 threads)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_tf32x3_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// K-major, SWIZZLE_128B: 8-row groups of 1024 B, row pitch 128 B, 16-byte chunks XOR-swizzled by row%8
__device__ __forceinline__ uint32_t sw128_offset(int row, int k /*0..31 floats*/) {
  const int chunk = k >> 2;
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4) + (k & 3) * 4);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);       // start address
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ A, const float* __restrict__ W,
                                                float* __restrict__ C, int terms) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  float* a_hi = (float*)base;                   // 128 x 32 : 16 KB
  float* a_lo = (float*)(base + 16384);
  float* b_hi = (float*)(base + 32768);         // 256 x 32 : 32 KB
  float* b_lo = (float*)(base + 65536);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  // instruction descriptor: D=f32, A=B=tf32, K-major both, N=256, M=128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

  uint32_t phase = 0;
  for (int atom = 0; atom < 8; ++atom) {
    // stage the K-atom (32 columns) of A and W as tf32 hi / lo in the swizzled layout
    for (int i = tid; i < 128 * 32; i += 128) {
      const int r = i >> 5, k = i & 31;
      const float x = A[r * 256 + atom * 32 + k];
      const float h = tf32_rna(x);
      const uint32_t off = sw128_offset(r, k);
      *(float*)((unsigned char*)a_hi + off) = h;
      *(float*)((unsigned char*)a_lo + off) = tf32_rna(x - h);
    }
    for (int i = tid; i < 256 * 32; i += 128) {
      const int n = i >> 5, k = i & 31;
      const float x = W[n * 256 + atom * 32 + k];
      const float h = tf32_rna(x);
      const uint32_t off = sw128_offset(n, k);
      *(float*)((unsigned char*)b_hi + off) = h;
      *(float*)((unsigned char*)b_lo + off) = tf32_rna(x - h);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kk = 0; kk < 4; ++kk) {  // 4 x (K = 8 tf32 = 32 bytes) per 128-byte atom
        const uint64_t dah = make_desc(smem_u32(a_hi) + kk * 32), dal = make_desc(smem_u32(a_lo) + kk * 32);
        const uint64_t dbh = make_desc(smem_u32(b_hi) + kk * 32), dbl = make_desc(smem_u32(b_lo) + kk * 32);
        mma_tf32(tmem, dah, dbh, idesc, (atom | kk) ? 1u : 0u);
        if (terms == 3) {
          mma_tf32(tmem, dal, dbh, idesc, 1u);
          mma_tf32(tmem, dah, dbl, idesc, 1u);
        } else if (terms == 4) {  // cross terms in their own (small-magnitude) accumulator
          mma_tf32(tmem + 256, dal, dbh, idesc, (atom | kk) ? 1u : 0u);
          mma_tf32(tmem + 256, dah, dbl, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs of this atom before the stage is overwritten
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&bar)), "r"(phase)
          : "memory");
    }
    phase ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w reads TMEM lanes [32w, 32w+32): thread = row, 32 consecutive columns per load
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
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
    if (terms == 4) {
      uint32_t w[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
            "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]),
            "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]),
            "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
          : "r"(taddr + 256)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
    }
    for (int j = 0; j < 32; ++j) C[row * 256 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// MMA issue-rate probe: n back-to-back M128 N256 K8 tf32 MMAs on resident (garbage) operands
__global__ void __launch_bounds__(128, 1) mma_rate(int n, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 98304 / 4; i += 128) ((float*)base)[i] = 0.f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
  if (tid == 0) {
    const uint64_t da = make_desc(smem_u32(base)), db = make_desc(smem_u32(base + 32768));
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) mma_tf32(tmem + ((i & 1) ? 256 : 0), da, db, idesc, i > 1);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&bar)), "r"(0)
          : "memory");
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  const int M = 128, N = 256, K = 256;
  std::vector<float> A(M * K), W(N * K), C(M * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : W) x = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.0625f;
  float *dA, *dW, *dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dC, C.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 98304 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int rc = 0;
  for (int terms : {1, 3, 4}) {
    cudaMemset(dC, 0, C.size() * 4);
    probe<<<1, 128, smem>>>(dA, dW, dC, terms);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    double max_abs = 0, max_ref = 0, max_f32 = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0; float sf = 0.f;
        for (int k = 0; k < K; ++k) { s += (double)A[m * K + k] * (double)W[n * K + k]; sf = fmaf(A[m * K + k], W[n * K + k], sf); }
        max_abs = fmax(max_abs, fabs(s - (double)C[m * N + n]));
        max_f32 = fmax(max_f32, fabs(s - (double)sf));
        max_ref = fmax(max_ref, fabs(s));
      }
    printf("terms=%d  max|C-ref|=%.3e  (fp32 fma chain vs ref: %.3e)  max|ref|=%.3f\n", terms, max_abs, max_f32, max_ref);
    if (terms == 3 && !(max_abs < 6e-6)) rc = 1;
    if (terms == 4 && !(max_abs < 2e-6)) rc = 1;
    if (terms == 1 && !(max_abs < 5e-3)) rc = 1;
  }
  {
    long long* dcy; cudaMalloc(&dcy, 148 * 8);
    cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {1, 148}) {
      for (int n : {96, 960}) {
        mma_rate<<<grid, 128, smem>>>(n, dcy);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        mma_rate<<<grid, 128, smem>>>(n, dcy);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long cy[148]; cudaMemcpy(cy, dcy, grid * 8, cudaMemcpyDeviceToHost);
        printf("mma_rate grid=%d n=%d: %.1f cycles/MMA (CTA0), kernel %.3f ms, %.1f TFLOP/s aggregate\n", grid, n,
               (double)cy[0] / n, ms, 2.0 * 128 * 256 * 8 * n * grid / (ms * 1e-3) / 1e12);
      }
    }
  }
  printf(rc ? "PROBE FAILED\n" : "PROBE OK\n");
  return rc;
}

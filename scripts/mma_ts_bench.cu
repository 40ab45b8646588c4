// Micro-benchmark + layout probe: tcgen05.mma with the A operand in TENSOR MEMORY (kind::f16, cta_group::1, M = 128, N = 16,
// K = 16 per instruction) against the same product with A in shared memory -- the shape of the recurrent kernels' per-step
// product (W_hh slice x h tile).  A is written to TMEM with tcgen05.st (lane = row, one 32-bit column = two consecutive k).
//   1. probe: one MMA with B = identity shows which (row, k) element the tensor core reads from which TMEM column half;
//   2. check: D [128 x 16] = A [128 x 256] x B^T [16 x 256] over 16 MMAs against the host, A from TMEM and A from smem;
//   3. timing: 32 back-to-back MMAs (two M = 128 halves x 16 k-steps), cycles per MMA, TS against SS.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ss_asr_b200/csrc -o build_tmp/mma_ts_bench scripts/mma_ts_bench.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "tc_common.cuh"
using namespace ssasr::tc;

constexpr int M = 128, N = 16, K = 256;
// shared: A (SW128 K-major, 4 k-blocks of 128 rows x 128 B) | B (4 k-blocks of 16 rows x 128 B)
constexpr int A_BLK = 128 * 128, B_BLK = N * 128;

__device__ __forceinline__ int sw128_off(int row, int k) {   // byte offset of element (row, k % 64) inside a k-block
  return row * 128 + ((((k & 63) >> 3) ^ (row & 7)) << 4) + (k & 7) * 2;
}

// mode 0: probe (one MMA, B = identity); mode 1: full check TS; mode 2: full check SS; mode 3: timing
__global__ void __launch_bounds__(192, 1) kern(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, long long* cyc, int mode) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Asm = sm;
  uint8_t* Bsm = sm + 4 * A_BLK;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 4) tmem_alloc<512>(&slot);
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(Asm + (k >> 6) * A_BLK + sw128_off(r, k)) = A[i];
  }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(Bsm + (k >> 6) * B_BLK + sw128_off(r, k)) = B[i];
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t tA = tmem + 64;          // A image: 128 columns (256 k) from column 64; a second copy (timing) at 64 + 128
  if (warp < 4) {                         // lane = row: 128 packed words of the thread's row -> TMEM
    const int row = warp * 32 + lane;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(A + (size_t)row * K);
    for (int c0 = 0; c0 < K / 2; c0 += 32) {
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = src[c0 + j];
      tmem_st32(tA + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_st32(tA + 128 + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 5 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(M, N);
    const uint32_t b0 = smem_u32(Bsm), a0 = smem_u32(Asm);
    if (mode == 0) {
      mma_bf16_ts(tmem, tA, umma_desc_k128(b0), idesc, 0);
      mma_commit(&bar);
      mbar_wait(&bar, 0);
    } else if (mode == 1 || mode == 2) {
      for (int kk = 0; kk < K / 16; ++kk) {
        const uint64_t db = umma_desc_k128(b0 + (kk >> 2) * B_BLK) + (uint64_t)((kk & 3) * 2);
        if (mode == 1) mma_bf16_ts(tmem, tA + kk * 8, db, idesc, kk != 0);
        else mma_bf16_ss(tmem, umma_desc_k128(a0 + (kk >> 2) * A_BLK) + (uint64_t)((kk & 3) * 2), db, idesc, kk != 0);
      }
      mma_commit(&bar);
      mbar_wait(&bar, 0);
    } else {
      // descriptors precomputed (uniform registers only inside the timed loops), every loop fully unrolled
      uint64_t db[16], da[16];
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        db[kk] = umma_desc_k128(b0 + (kk >> 2) * B_BLK) + (uint64_t)((kk & 3) * 2);
        da[kk] = umma_desc_k128(a0 + (kk >> 2) * A_BLK) + (uint64_t)((kk & 3) * 2);
      }
      int ph = 0;
      for (int rep = 0; rep < 3; ++rep) {            // TS, two accumulators (the two M = 128 halves of a CTA's gate rows)
        const long long t0 = clock64();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          mma_bf16_ts(tmem, tA + kk * 8, db[kk], idesc, kk != 0);
          mma_bf16_ts(tmem + 16, tA + 128 + kk * 8, db[kk], idesc, kk != 0);
        }
        mma_commit(&bar);
        const long long t1 = clock64();
        mbar_wait(&bar, ph); ph ^= 1;
        const long long t2 = clock64();
        if (rep == 2) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
      }
      for (int rep = 0; rep < 3; ++rep) {            // SS
        const long long t0 = clock64();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          mma_bf16_ss(tmem, da[kk], db[kk], idesc, kk != 0);
          mma_bf16_ss(tmem + 16, da[kk], db[kk], idesc, kk != 0);
        }
        mma_commit(&bar);
        const long long t1 = clock64();
        mbar_wait(&bar, ph); ph ^= 1;
        const long long t2 = clock64();
        if (rep == 2) { cyc[2] = t1 - t0; cyc[3] = t2 - t0; }
      }
      for (int rep = 0; rep < 3; ++rep) {            // TS, one accumulator, N = 32 (16 MMAs: R = 32 rows per tile)
        const uint32_t idesc32 = umma_idesc_bf16(M, 32);
        const long long t0 = clock64();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) mma_bf16_ts(tmem, tA + kk * 8, db[kk], idesc32, kk != 0);
        mma_commit(&bar);
        const long long t1 = clock64();
        mbar_wait(&bar, ph); ph ^= 1;
        const long long t2 = clock64();
        if (rep == 2) { cyc[4] = t1 - t0; cyc[5] = t2 - t0; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < N; ++j) D[(warp * 32 + lane) * N + j] = __uint_as_float(v[j]);
    // probe of the 16x256b load shape (columns 0..7 of the accumulator): what lands in which thread
    uint32_t w[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(tmem + ((uint32_t)(warp * 32 + 16) << 16)) : "memory");
    tmem_ld_wait();
    for (int j = 0; j < 8; ++j) D[M * N + (warp * 32 + lane) * 8 + j] = __uint_as_float(w[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem);
}

int main() {
  std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), hD(M * N + M * 8);
  __nv_bfloat16 *dA, *dB;
  float* dD;
  long long* dC;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, (M * N + M * 8) * 4); cudaMalloc(&dC, 64);
  const int smem = 4 * A_BLK + 4 * B_BLK + 2048;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // ---- probe: A[m][k] = (m % 8) * 16 + k for k < 16, B = identity on the first 16 k
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { fA[m * K + k] = k < 16 ? (float)((m % 8) * 16 + k) : 0.f; hA[m * K + k] = __float2bfloat16(fA[m * K + k]); }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { fB[n * K + k] = n == k ? 1.f : 0.f; hB[n * K + k] = __float2bfloat16(fB[n * K + k]); }
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  kern<<<1, 192, smem>>>(dA, dB, dD, dC, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  for (int m : {0, 1, 9, 33, 127}) {
    printf("  row %3d (want %d + k):", m, (m % 8) * 16);
    for (int n = 0; n < N; ++n) printf(" %g", hD[m * N + n]);
    printf("\n");
  }
  // ---- full check with random small integers
  srand(1);
  for (int i = 0; i < M * K; ++i) { fA[i] = (float)(rand() % 7 - 3); hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < N * K; ++i) { fB[i] = (float)(rand() % 5 - 2); hB[i] = __float2bfloat16(fB[i]); }
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  for (int mode = 1; mode <= 2; ++mode) {
    kern<<<1, 192, smem>>>(dA, dB, dD, dC, mode);
    e = cudaDeviceSynchronize();
    cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)fA[m * K + k] * fB[n * K + k];
      const double d = fabs(ref - hD[m * N + n]);
      if (d > worst) worst = d;
    }
    printf("check %s: %s, max |D - ref| = %g\n", mode == 1 ? "A in TMEM" : "A in smem", cudaGetErrorString(e), worst);
    if (mode == 1) {
      // 16x256b: thread t of a warp is expected to hold rows t/4 and t/4 + 8 of the addressed 16-lane block, columns 2 (t % 4), + 1
      cudaMemcpy(hD.data(), dD, (M * N + M * 8) * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int w = 0; w < 4; ++w) for (int t = 0; t < 32; ++t) for (int blk = 0; blk < 2; ++blk) for (int j = 0; j < 4; ++j) {
        const int row = w * 32 + blk * 16 + t / 4 + (j >> 1) * 8, col = 2 * (t % 4) + (j & 1);
        if (hD[M * N + (w * 32 + t) * 8 + blk * 4 + j] != hD[row * N + col]) ++bad;
      }
      printf("tcgen05.ld.16x256b.x1 layout (rows t/4, t/4+8; columns 2(t%%4), +1): %d mismatches of 1024\n", bad);
    }
  }
  kern<<<1, 192, smem>>>(dA, dB, dD, dC, 3);
  e = cudaDeviceSynchronize();
  long long c[6];
  cudaMemcpy(c, dC, 48, cudaMemcpyDeviceToHost);
  printf("timing %s: 32 MMAs M=128 N=16 K=16:  A in TMEM issue %lld complete %lld cyc = %.1f cyc/MMA;  A in smem issue %lld complete %lld = %.1f cyc/MMA\n",
         cudaGetErrorString(e), c[0], c[1], c[1] / 32.0, c[2], c[3], c[3] / 32.0);
  printf("        16 MMAs M=128 N=32 (B rows 16..31 = garbage) A in TMEM issue %lld complete %lld = %.1f cyc/MMA\n", c[4], c[5], c[5] / 16.0);
  return 0;
}

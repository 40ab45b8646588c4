// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, K=16) issued back to back into one accumulator, for the
// operand shapes / shared-memory swizzle modes of the recurrent kernels.  Operands are garbage; only timing matters.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ss_asr_b200/csrc -o build_tmp/mma_bench scripts/mma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "tc_common.cuh"
using namespace ssasr::tc;

__device__ __forceinline__ uint64_t desc(uint32_t addr, int sw_bytes) {
  // K-major: rows of sw_bytes, 8-row groups 8*sw_bytes apart
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * sw_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(sw_bytes == 128 ? 2 : sw_bytes == 64 ? 4 : 6) << 61;
  return d;
}

// n_mma MMAs of shape M x N x 16; A k-steps advance through `a_sw`-byte swizzled k-blocks, B likewise
// spin_warps: extra warps that wait on an mbarrier that completes only at the end (the epilogue warps of the recurrent
// kernels do this during the MMA phase); grid > 1: the same CTA on many SMs at once
__global__ void __launch_bounds__(384, 1) bench(int M, int N, int n_mma, int a_sw, int b_sw, int spin_warps, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<256>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(M, N);
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm + 96 * 1024);
    const int a_steps = a_sw / 32, b_steps = b_sw / 32;           // K=16 steps per k-block
    const int a_blk = 64 * a_sw, b_blk = N * b_sw;                // bytes per k-block (64 A rows reserved even for M=128 reads)
    // descriptors precomputed: the timed loop only loads two 64-bit words per MMA
    uint64_t* dtab = reinterpret_cast<uint64_t*>(sm + 190 * 1024);
    for (int i = 0; i < n_mma; ++i) {
      dtab[2 * i] = desc(a0 + (i / a_steps) * a_blk, a_sw) + (uint64_t)((i % a_steps) * 2);
      dtab[2 * i + 1] = desc(b0 + (i / b_steps) * b_blk, b_sw) + (uint64_t)((i % b_steps) * 2);
    }
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
#pragma unroll 8
      for (int i = 0; i < n_mma; ++i) mma_bf16_ss(tmem, dtab[2 * i], dtab[2 * i + 1], idesc, i != 0);
      mma_commit(&bar);
      long long t1 = clock64();
      mbar_wait(&bar, rep & 1);
      long long t2 = clock64();
      if (rep == 2 && blockIdx.x == gridDim.x / 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    mbar_arrive(&bar2);
  } else if (warp >= 2 && warp < 2 + spin_warps) {
    mbar_wait(&bar2, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int M, N, n, a_sw, b_sw; const char* what; };
  Cfg cfgs[] = {
      {64, 32, 64, 128, 128, "bwd now   M=64  N=32  A SW128 B SW128"}, {64, 32, 64, 32, 128, "bwd       M=64  N=32  A SW32  B SW128"},
      {64, 32, 64, 32, 32, "bwd       M=64  N=32  A SW32  B SW32 "},   {64, 32, 64, 64, 64, "bwd       M=64  N=32  A SW64  B SW64 "},
      {128, 32, 64, 128, 128, "bwd       M=128 N=32  A SW128 B SW128"},
      {64, 128, 16, 64, 128, "fwd now   M=64  N=128 A SW64  B SW128"}, {64, 128, 16, 32, 128, "fwd       M=64  N=128 A SW32  B SW128"},
      {64, 128, 16, 32, 32, "fwd       M=64  N=128 A SW32  B SW32 "},  {64, 128, 16, 128, 128, "fwd       M=64  N=128 A SW128 B SW128"},
      {128, 128, 16, 128, 128, "gemm      M=128 N=128 A SW128 B SW128"}, {128, 256, 8, 128, 128, "gemm      M=128 N=256 A SW128 B SW128"},
      {64, 64, 16, 32, 128, "          M=64  N=64  A SW32  B SW128"},   {64, 16, 64, 128, 128, "          M=64  N=16  A SW128 B SW128"},
  };
  const int grids[] = {1, 2, 64, 148};
  for (int spin = 0; spin <= 8; spin += 8)
    for (int gi = 0; gi < 4; ++gi)
      for (auto& c : cfgs) {
        bench<<<grids[gi], 384, 192 * 1024>>>(c.M, c.N, c.n, c.a_sw, c.b_sw, spin, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", c.what, cudaGetErrorString(e)); return 1; }
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("grid %3d spin %d %s: %3d MMAs issue %6lld cyc, complete %6lld cyc = %5.1f cyc/MMA\n", grids[gi], spin, c.what, c.n, h[0], h[1],
               (double)h[1] / c.n);
      }
  return 0;
}

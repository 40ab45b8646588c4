// Micro-benchmark: per-step cost of the h / dG exchange between the CTAs that share one (direction, batch tile) of the
// recurrent BLSTM kernels.  Every step each CTA produces `bytes` and must receive the `bytes` of every CTA of its group
// before it can start the next step (a dependent chain, like the recurrence).
//   V1  cluster, DSMEM bulk copy  (cp.async.bulk.shared::cluster.shared::cta + complete_tx on the peer's mbarrier)
//   V2  cluster, st.global + multicast bulk copy back (cp.async.bulk ... .multicast::cluster), no inter-CTA barrier
//   V3  cluster, st.shared::cluster from all threads + remote mbarrier arrive (release.cluster)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exch_bench scripts/exch_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) if (++spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) if (++spins > (1u << 26)) __trap();
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------- V1: DSMEM bulk copy ----------------
template <int CS>
__global__ void __launch_bounds__(128, 1) v1_kernel(int steps, int bytes, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* recv = sm;                           // [2][CS][bytes]
  uint8_t* send = sm + 2 * CS * bytes;          // [2][bytes]
  uint64_t* full = reinterpret_cast<uint64_t*>(send + 2 * bytes);
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(full + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    const int b = s & 1;
    uint8_t* sb = send + b * bytes;
    for (int i = threadIdx.x * 16; i < bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sb + i) = make_uint4(s, rank, i, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) mbar_expect_tx(full + b, CS * bytes);
    if (threadIdx.x < CS) {
      const uint32_t peer = threadIdx.x;
      const uint32_t dst = mapa(smem_u32(recv + (b * CS + rank) * bytes), peer);
      const uint32_t rbar = mapa(smem_u32(full + b), peer);
      asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "r"(smem_u32(sb)), "r"(bytes), "r"(rbar) : "memory");
    }
    mbar_wait(full + b, (s >> 1) & 1);
  }
  long long t1 = clock64();
  cluster_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = *reinterpret_cast<int*>(recv + (CS - 1) * bytes);
}

// ---------------- V2: st.global + multicast bulk readback ----------------
template <int CS>
__global__ void __launch_bounds__(128, 1) v2_kernel(int steps, int bytes, uint8_t* gbuf, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* recv = sm;                           // [2][CS][bytes]
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + 2 * CS * bytes);
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(full + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    const int b = s & 1;
    uint8_t* gb = gbuf + ((size_t)(s & 3) * gridDim.x + blockIdx.x) * bytes;
    for (int i = threadIdx.x * 16; i < bytes; i += 128 * 16) *reinterpret_cast<uint4*>(gb + i) = make_uint4(s, rank, i, 0);
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(full + b, CS * bytes);
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
          ::"r"(smem_u32(recv + (b * CS + rank) * bytes)), "l"(gb), "r"(bytes), "r"(smem_u32(full + b)), "h"((uint16_t)((1u << CS) - 1))
          : "memory");
    }
    mbar_wait(full + b, (s >> 1) & 1);
  }
  long long t1 = clock64();
  cluster_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = *reinterpret_cast<int*>(recv + (CS - 1) * bytes);
}

// ---------------- V3: st.shared::cluster from threads + remote arrive ----------------
template <int CS>
__global__ void __launch_bounds__(128, 1) v3_kernel(int steps, int bytes, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* recv = sm;                           // [2][CS][bytes]
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + 2 * CS * bytes);
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) {
    mbar_init(full, CS * 4);                    // one arrive per warp per peer
    mbar_init(full + 1, CS * 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    const int b = s & 1;
    const uint32_t loc = smem_u32(recv + (b * CS + rank) * bytes);
#pragma unroll
    for (int peer = 0; peer < CS; ++peer) {
      const uint32_t dst = mapa(loc, peer);
      for (int i = threadIdx.x * 16; i < bytes; i += 128 * 16)
        asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + i), "r"(s), "r"(rank), "r"(i), "r"(0) : "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int peer = 0; peer < CS; ++peer) {
        const uint32_t rbar = mapa(smem_u32(full + b), peer);
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
      }
    }
    mbar_wait_cluster(full + b, (s >> 1) & 1);
  }
  long long t1 = clock64();
  cluster_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = *reinterpret_cast<int*>(recv + (CS - 1) * bytes);
}


// ---------------- V4: smem staging -> bulk store to global -> wait -> multicast bulk readback (no generic global stores) ----------------
template <int CS>
__global__ void __launch_bounds__(128, 1) v4_kernel(int steps, int bytes, uint8_t* gbuf, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* recv = sm;                           // [2][CS][bytes]
  uint8_t* send = sm + 2 * CS * bytes;          // [bytes]
  uint64_t* full = reinterpret_cast<uint64_t*>(send + bytes);
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(full + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    const int b = s & 1;
    uint8_t* gb = gbuf + ((size_t)(s & 3) * gridDim.x + blockIdx.x) * bytes;
    for (int i = threadIdx.x * 16; i < bytes; i += 128 * 16) *reinterpret_cast<uint4*>(send + i) = make_uint4(s, rank, i, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(full + b, CS * bytes);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gb), "r"(smem_u32(send)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
          ::"r"(smem_u32(recv + (b * CS + rank) * bytes)), "l"(gb), "r"(bytes), "r"(smem_u32(full + b)), "h"((uint16_t)((1u << CS) - 1))
          : "memory");
    }
    mbar_wait(full + b, (s >> 1) & 1);
  }
  long long t1 = clock64();
  cluster_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = *reinterpret_cast<int*>(recv + (CS - 1) * bytes + 4);
}

template <typename K, typename... Args>
static double run(K kern, int cs, int nclusters, size_t smem, int steps, long long* dout, Args... args) {
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cs > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * nclusters);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
    if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  run failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long h[2];
  CK(cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost));
  return (double)h[0] / steps;
}

int main() {
  long long* dout;
  uint8_t* gbuf;
  CK(cudaMalloc(&dout, 64));
  CK(cudaMalloc(&gbuf, 64 << 20));
  const int steps = 2000;
  const int sizes[] = {2048, 4096, 8192, 16384};
  for (int nclusters : {1, 8}) {
    for (int bytes : sizes) {
      {
        const int cs = 8;
        size_t sm1 = (size_t)2 * cs * bytes + 2 * bytes + 64, sm2 = (size_t)2 * cs * bytes + 64;
        if (sm1 > 227 * 1024) continue;
        double a = run(v1_kernel<8>, cs, nclusters, sm1, steps, dout, steps, bytes, dout);
        double b = run(v2_kernel<8>, cs, nclusters, sm2, steps, dout, steps, bytes, gbuf, dout);
        double c = run(v3_kernel<8>, cs, nclusters, sm2, steps, dout, steps, bytes, dout);
        double d = run(v4_kernel<8>, cs, nclusters, sm1, steps, dout, steps, bytes, gbuf, dout);
        printf("cs=8 clusters=%2d bytes/cta=%5d (ingest %6d B): V1 dsmem-bulk %7.0f cyc  V2 global+mcast %7.0f cyc  V3 st.cluster %7.0f cyc  V4 s2g+mcast %7.0f cyc\n",
               nclusters, bytes, cs * bytes, a, b, c, d);
      }
      {
        const int cs = 4;
        size_t sm1 = (size_t)2 * cs * bytes + 2 * bytes + 64, sm2 = (size_t)2 * cs * bytes + 64;
        double a = run(v1_kernel<4>, cs, nclusters * 2, sm1, steps, dout, steps, bytes, dout);
        double b = run(v2_kernel<4>, cs, nclusters * 2, sm2, steps, dout, steps, bytes, gbuf, dout);
        double c = run(v3_kernel<4>, cs, nclusters * 2, sm2, steps, dout, steps, bytes, dout);
        printf("cs=4 clusters=%2d bytes/cta=%5d (ingest %6d B): V1 dsmem-bulk %7.0f cyc  V2 global+mcast %7.0f cyc  V3 st.cluster %7.0f cyc\n",
               nclusters * 2, bytes, cs * bytes, a, b, c);
      }
    }
  }
  // 16-CTA clusters (non-portable)
  for (int bytes : {2048, 4096}) {
    const int cs = 16;
    size_t sm1 = (size_t)2 * cs * bytes + 2 * bytes + 64, sm2 = (size_t)2 * cs * bytes + 64;
    double a = run(v1_kernel<16>, cs, 8, sm1, steps, dout, steps, bytes, dout);
    double b = run(v2_kernel<16>, cs, 8, sm2, steps, dout, steps, bytes, gbuf, dout);
    printf("cs=16 clusters= 8 bytes/cta=%5d (ingest %6d B): V1 dsmem-bulk %7.0f cyc  V2 global+mcast %7.0f cyc\n", bytes, cs * bytes, a, b);
  }
  return 0;
}

// Cluster-exchange primitives shared by the cluster-resident kernels (rec_cl.cu, spell_cl.cu): time-bounded mbarrier waits,
// remote arrives, bulk store / multicast bulk load, TMA tensor load / store, small-swizzle UMMA descriptors.
#pragma once
#include "tc_common.cuh"

namespace ssasr {
namespace clx {

using namespace tc;

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float tanh_apx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_apx(float x) { return fmaf(tanh_apx(0.5f * x), 0.5f, 0.5f); }
// exact-path activations of the cluster kernels: ex2.approx / rcp.approx based, ABSOLUTE error <= 1e-7 (sigmoid: s (1 - s) x the
// 2^-22 + |x| 2^-24 relative error of the exponential; tanh: from exp(-2|x|) <= 1, no cancellation beyond one rounding) -- an order
// of magnitude below the 2^-17 operand split of the products they follow, at a quarter of the instructions of expf / tanhf (the cell
// math of 4 cells per thread was the longest phase of the exact recurrent step)
__device__ __forceinline__ float sigmoid_x(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_x(float x) {
  const float e = __expf(-2.f * fabsf(x));
  return copysignf(__fdividef(1.f - e, 1.f + e), x);
}

// time-bounded mbarrier wait (a protocol bug must trap, not hang the GPU): ~2 s at 2 GHz
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_cluster_t(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
// arrive on the mbarrier at the same CTA-relative offset in cluster CTA `rank`
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// Same, relaxed: for "I have finished READING" notifications.  The reads in question have completed before the arriving thread
// observed their completion barrier, and the arrive is control-dependent on that observation, so no release fence is needed
// -- and a cluster-scope release costs 1-3 k cycles here (it drains every outstanding global access of the thread).
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// image (shared) -> ring slot (global), then wait until the bulk store has completed
__device__ __forceinline__ void bulk_store_wait(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// ring slot (global) -> the same CTA-relative smem offset in every CTA of `mask`; completes its bytes on the mbarrier at
// the same offset in each of them
__device__ __forceinline__ void bulk_load_mc(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// own shared memory -> the same CTA-relative offset `sdst` in cluster CTA `rank`, completing its bytes on the mbarrier at offset
// `bar` of that CTA (DSMEM bulk copy: no round trip through L2)
__device__ __forceinline__ void bulk_copy_dsmem(void* sdst, const void* ssrc, uint32_t bytes, uint64_t* bar, uint32_t rank) {
  uint32_t rdst, rbar;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(smem_u32(sdst)), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(rdst), "r"(smem_u32(ssrc)), "r"(bytes), "r"(rbar) : "memory");
}
// swizzled shared-memory tile -> global tensor (rows outside the tensor are clipped); joins the thread's bulk group
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* ssrc, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(ssrc)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// K-major operand k-block of 16 bf16 (32-byte rows, 32-byte swizzle), 8-row groups 256 B apart
__device__ __forceinline__ uint64_t umma_desc_k32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;            // layout type SWIZZLE_32B
  return d;
}
// K-major operand k-block of 32 bf16 (64-byte rows, 64-byte swizzle), 8-row groups 512 B apart
__device__ __forceinline__ uint64_t umma_desc_k64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;            // layout type SWIZZLE_64B
  return d;
}

__device__ __forceinline__ void tmem_ld1(uint32_t ta, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(ta) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t ta, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(ta) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
// 4 x 4 transpose inside a lane quad: before, lane g holds columns 0..3 of ITS row; afterwards lane gp holds rows 0..3 of column gp
__device__ __forceinline__ void quad_transpose(float& a0, float& a1, float& a2, float& a3, int gp) {
  {
    const float x = (gp & 1) ? a0 : a1, y = (gp & 1) ? a2 : a3;
    const float xr = __shfl_xor_sync(0xffffffffu, x, 1), yr = __shfl_xor_sync(0xffffffffu, y, 1);
    if (gp & 1) { a0 = xr; a2 = yr; } else { a1 = xr; a3 = yr; }
  }
  {
    const float x = (gp & 2) ? a0 : a2, y = (gp & 2) ? a1 : a3;
    const float xr = __shfl_xor_sync(0xffffffffu, x, 2), yr = __shfl_xor_sync(0xffffffffu, y, 2);
    if (gp & 2) { a0 = xr; a1 = yr; } else { a2 = xr; a3 = yr; }
  }
}


}  // namespace clx
}  // namespace ssasr

// bf16 tensor-core GEMM for the batched-over-time LSTM gate products of the training path:
//   C[M,N] (fp32) (+)= A[M,K] (bf16, K contiguous) x B[N,K]^T (bf16, K contiguous) (+ bias[N])
// tcgen05.mma (cta_group::1, 128 x BN x 16 atoms, fp32 accumulators in TMEM) fed by TMA through a multi-stage
// mbarrier ring; warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 = epilogue (tcgen05.ld -> bias -> global).  Out-of-range rows/columns/reduction indices are
// zero-filled by TMA, so M, N, K need no padding (only 16-byte row pitches).
//
// Reference call sites replaced: the cuDNN/ATen input-projection GEMMs inside nn.LSTM (asr.py:234-238,414,262)
// and their dgrad/wgrad in loss.backward() (trainer.py:437) -- SURVEY.md §2.2 K2a/K4/K17.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace ssasr {

using namespace tc;

constexpr int GT_BM = 128;
constexpr int GT_BK = 64;             // 64 bf16 = 128 B = one swizzle row

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GT_BM * GT_BK * 2;
  static constexpr int B_BYTES = BN * GT_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16;
};


// Coalesced epilogue of one 32-row x 32-column fp32 chunk.  tcgen05.ld hands every lane one ROW of the accumulator; written
// straight to global that is 32 different 128-byte lines per store instruction (LSU-bound: a 128 x 128 tile took as long as
// its K = 1024 main loop).  The chunk is therefore transposed through a per-warp 4 KB shared-memory tile (16-byte chunk c of
// row r at chunk c ^ (r & 7): conflict-free both ways) so that every global instruction covers 4 rows x 128 contiguous bytes.
template <bool ATOMIC>
__device__ __forceinline__ void epi_chunk(float* stage, const uint32_t* v, float* __restrict__ C, int ldc, int row0, int col0,
                                          int M, int N, const float* __restrict__ bias, int accumulate, int act_tanh) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(stage + lane * 32 + ((c ^ (lane & 7)) << 2)) =
        make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
  __syncwarp();
  const int rr = lane >> 3, cc = lane & 7;
  const int n = col0 + cc * 4;
  const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && n + 3 < N;
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) {
    if (n < N) bb.x = __ldg(bias + n);
    if (n + 1 < N) bb.y = __ldg(bias + n + 1);
    if (n + 2 < N) bb.z = __ldg(bias + n + 2);
    if (n + 3 < N) bb.w = __ldg(bias + n + 3);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + rr, row = row0 + r;
    float4 o = *reinterpret_cast<const float4*>(stage + r * 32 + ((cc ^ (r & 7)) << 2));
    if (row >= M) continue;
    float* dst = C + (size_t)row * ldc + n;
    o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
    if (vec_ok) {
      if (ATOMIC) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
      } else {
        if (accumulate) { const float4 q = *reinterpret_cast<const float4*>(dst); o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
        if (act_tanh) { o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w); }
        *reinterpret_cast<float4*>(dst) = o;
      }
    } else {
      const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (n + e < N) {
          float x = ov[e];
          if (ATOMIC) {
            atomicAdd(dst + e, x);
          } else {
            if (accumulate) x += dst[e];
            if (act_tanh) x = tanhf(x);
            dst[e] = x;
          }
        }
      }
    }
  }
  __syncwarp();
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(256, (BN * 64 + 8192) * STAGES * 2 <= 100 * 1024 ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
               int ldc, const float* __restrict__ bias, int M, int N, int K, int a_koff, int b_koff, int accumulate, int act_tanh,
               int kb_per_split) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  // blockIdx.x (fastest in launch order) walks the N tiles of one 128-row block of A: the block is read from HBM once and
  // serves all its column tiles out of L2, B (the weights) stays L2-resident throughout
  const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * BN;
  // split-K (grid.z > 1, the latency-bound per-step products of the decoder loop: halves the bytes every SM has to pull
  // through its L2 port): each CTA reduces kb_per_split k-blocks and adds its partial to the pre-zeroed output with red.add
  const int kb_all = (K + GT_BK - 1) / GT_BK;
  const int kb_first = blockIdx.z * kb_per_split;
  const int num_k = min(kb_per_split, kb_all - kb_first);
  const bool atomic_out = gridDim.z > 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty_bar + s, ph ^ 1);
        mbar_expect_tx(full_bar + s, L::STAGE_BYTES);
        uint8_t* st = smem + s * L::STAGE_BYTES;
        tma_load_2d(&tmA, full_bar + s, st, a_koff + (kb_first + kb) * GT_BK, m0);
        tma_load_2d(&tmB, full_bar + s, st + L::A_BYTES, b_koff + (kb_first + kb) * GT_BK, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GT_BM, BN);
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(full_bar + s, ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
        const uint64_t da = umma_desc_k128(a_addr);
        const uint64_t db = umma_desc_k128(a_addr + L::A_BYTES);
#pragma unroll
        for (int k = 0; k < GT_BK / 16; ++k)   // +32 B per K=16 step inside the 128-B swizzle row
          mma_bf16_ss(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        mma_commit(empty_bar + s);
      }
      mma_commit(tmem_full);
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;                 // TMEM lanes [32*ew, 32*ew+32)
    mbar_wait(tmem_full, 0);                 // every MMA has completed: the pipeline stages are free, stage 0 becomes staging
    tc_fence_after();
    float* stage = reinterpret_cast<float*>(smem) + ew * 1024;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= N) break;
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (atomic_out) epi_chunk<true>(stage, v, C, ldc, m0 + ew * 32, n0 + c0, M, N, blockIdx.z == 0 ? bias : nullptr, 0, 0);
      else epi_chunk<false>(stage, v, C, ldc, m0 + ew * 32, n0 + c0, M, N, bias, accumulate, act_tanh);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<BN>(tmem_base);
}

// ---- persistent variant for the large products (input projections, dgrad): one CTA per SM walks 128 x 256 output tiles
// (N fastest: the CTAs that share a 128-row block of A run at the same time, B = the weights stays in L2), 4 x 48 KB TMA stages,
// and TWO TMEM accumulators (2 x 256 columns) so that the epilogue of tile i (tcgen05.ld -> transpose -> coalesced stores, as
// above) runs under the MMAs of tile i + 1 instead of relying on a second co-resident CTA.  Twice the operand reuse per tile
// of the 128 x 128 kernel (48 KB of operands per 2 x 128 x 256 x 64 MACs).
constexpr int GP_BN = 256;
constexpr int GP_STAGES = 4;
constexpr int GP_A_BYTES = GT_BM * GT_BK * 2;
constexpr int GP_B_BYTES = GP_BN * GT_BK * 2;
constexpr int GP_STAGE_BYTES = GP_A_BYTES + GP_B_BYTES;
constexpr int GP_EPW = 8;                                             // epilogue warps: two per TMEM sub-partition, 128 columns each
constexpr int GP_THREADS = 128 + 32 * GP_EPW;
constexpr int GP_STAGING_OFF = GP_STAGES * GP_STAGE_BYTES;          // one 4 KB transpose tile per epilogue warp
constexpr int GP_BAR_OFF = GP_STAGING_OFF + GP_EPW * 4096;
constexpr int GP_TOTAL = GP_BAR_OFF + (2 * GP_STAGES + 4) * 8 + 16;

__global__ void __launch_bounds__(GP_THREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
                       int ldc, const float* __restrict__ bias, int M, int N, int K, int a_koff, int b_koff, int accumulate,
                       int act_tanh, int n_mt, int n_nt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GP_BAR_OFF);
  uint64_t* empty_bar = full_bar + GP_STAGES;
  uint64_t* tfull = empty_bar + GP_STAGES;       // [2]
  uint64_t* tempty = tfull + 2;                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5;
  const int num_k = (K + GT_BK - 1) / GT_BK;
  const int total = n_mt * n_nt;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < GP_STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tfull + 1, 1);
    mbar_init(tempty, GP_EPW);
    mbar_init(tempty + 1, GP_EPW);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m0 = (tile / n_nt) * GT_BM, n0 = (tile % n_nt) * GP_BN;
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % GP_STAGES;
          mbar_wait(empty_bar + s, ((it / GP_STAGES) & 1) ^ 1);
          mbar_expect_tx(full_bar + s, GP_STAGE_BYTES);
          uint8_t* st = smem + s * GP_STAGE_BYTES;
          tma_load_2d(&tmA, full_bar + s, st, a_koff + kb * GT_BK, m0);
          tma_load_2d(&tmB, full_bar + s, st + GP_A_BYTES, b_koff + kb * GT_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GT_BM, GP_BN);
      int it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tl) {
        const int acc = tl & 1;
        mbar_wait(tempty + acc, ((tl >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator (two tiles ago)
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * GP_BN);
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % GP_STAGES;
          mbar_wait(full_bar + s, (it / GP_STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * GP_STAGE_BYTES);
          const uint64_t da = umma_desc_k128(a_addr), db = umma_desc_k128(a_addr + GP_A_BYTES);
#pragma unroll
          for (int k = 0; k < GT_BK / 16; ++k) mma_bf16_ss(d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          mma_commit(empty_bar + s);
        }
        mma_commit(tfull + acc);
      }
    }
  } else if (warp >= 4) {
    const int ew = (warp - 4) & 3, ch = (warp - 4) >> 2;      // TMEM sub-partition, column half of the tile
    float* stage = reinterpret_cast<float*>(smem + GP_STAGING_OFF) + (warp - 4) * 1024;
    int tl = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tl) {
      const int m0 = (tile / n_nt) * GT_BM, n0 = (tile % n_nt) * GP_BN;
      const int acc = tl & 1;
      mbar_wait(tfull + acc, (tl >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = ch * (GP_BN / 2); c0 < (ch + 1) * (GP_BN / 2); c0 += 32) {
        if (n0 + c0 >= N) break;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * GP_BN + c0), v);
        tmem_ld_wait();
        epi_chunk<false>(stage, v, C, ldc, m0 + ew * 32, n0 + c0, M, N, bias, accumulate, act_tanh);
      }
      tc_fence_before();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(tempty + acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ---- fp32-accurate variant ("tf32x3"): C = A·B^T with fp32 operands split as x = hi + lo, hi = the upper 19 bits the
// tf32 tensor core reads, lo = x - hi (computed once per operand by split_lo_kernel); three MMAs per K step
// (hi·hi + hi·lo + lo·hi) accumulate in the same TMEM tile.  Dropped terms are O(2^-21) relative: fp32-level accuracy
// for the exact path (validation / greedy decode) at tensor-core speed.
template <int STAGES>
__global__ void __launch_bounds__(256, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBl, float* __restrict__ C, int ldc,
                   const float* __restrict__ bias, int M, int N, int K, int act_tanh) {
  constexpr int BN = 128, BKF = 32;             // 32 fp32 = 128 B = one swizzle row
  constexpr int T_BYTES = 128 * 128;            // one 128-row operand tile
  constexpr int STAGE_BYTES = 4 * T_BYTES;      // A_hi, A_lo, B_hi, B_lo
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * BN;     // N tiles fastest: a 128-row block of A (hi + lo) is read from HBM once
  const int num_k = (K + BKF - 1) / BKF;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmBl);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(empty_bar + s, ((kb / STAGES) & 1) ^ 1);
        mbar_expect_tx(full_bar + s, STAGE_BYTES);
        uint8_t* st = smem + s * STAGE_BYTES;
        tma_load_2d(&tmA, full_bar + s, st, kb * BKF, m0);
        tma_load_2d(&tmAl, full_bar + s, st + T_BYTES, kb * BKF, m0);
        tma_load_2d(&tmB, full_bar + s, st + 2 * T_BYTES, kb * BKF, n0);
        tma_load_2d(&tmBl, full_bar + s, st + 3 * T_BYTES, kb * BKF, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_tf32(GT_BM, BN);
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(full_bar + s, (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t dah = umma_desc_k128(base), dal = umma_desc_k128(base + T_BYTES);
        const uint64_t dbh = umma_desc_k128(base + 2 * T_BYTES), dbl = umma_desc_k128(base + 3 * T_BYTES);
#pragma unroll
        for (int k = 0; k < BKF / 8; ++k) {      // K = 8 fp32 = 32 B per step
          const uint64_t o = (uint64_t)(k * 2);
          mma_tf32_ss(tmem_base, dah + o, dbh + o, idesc, (kb | k) != 0);
          mma_tf32_ss(tmem_base, dah + o, dbl + o, idesc, 1);
          mma_tf32_ss(tmem_base, dal + o, dbh + o, idesc, 1);
        }
        mma_commit(empty_bar + s);
      }
      mma_commit(tmem_full);
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;                 // TMEM lanes [32*ew, 32*ew+32)
    mbar_wait(tmem_full, 0);                 // every MMA has completed: the pipeline stages are free, stage 0 becomes staging
    tc_fence_after();
    float* stage = reinterpret_cast<float*>(smem) + ew * 1024;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= N) break;
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      epi_chunk<false>(stage, v, C, ldc, m0 + ew * 32, n0 + c0, M, N, bias, 0, act_tanh);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<BN>(tmem_base);
}

// hi = rn_tf32(x), lo = rn_tf32(x - hi): both exactly representable in tf32, so the tensor core's own truncation of
// its fp32 inputs is a no-op and the only errors left are the unbiased roundings (2^-22 relative) and lo*lo.
__device__ __forceinline__ float rn_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__global__ void split_hi_lo_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float h = rn_tf32(v);
    hi[i] = h;
    lo[i] = rn_tf32(v - h);
  }
}

// ---- TN variant: C[M,N] = A^T B with A stored [K,M] and B stored [K,N] (both M/N-contiguous, "MN-major") --------
// Used for the weight gradients dW = dG^T X directly from the row-major bf16 activations / gate gradients, so no
// transposed copies are needed.  a_koff / b_koff shift the reduction ROW window of each operand (any integer:
// rows are the outer TMA dimension).
template <int BN, int STAGES>
__global__ void __launch_bounds__(256, (BN * 64 + 8192) * STAGES * 2 <= 100 * 1024 ? 2 : 1)
gemm_tc_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
                  int ldc, int M, int N, int K, int a_koff, int b_koff, int accumulate, int kb_per_split) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  constexpr int BLK = 64 * GT_BK * 2;      // one {64 mn, 64 k} box = 8 KB

  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * GT_BM, n0 = blockIdx.y * BN;
  // split-K: grid.z CTAs share one output tile, each reduces kb_per_split k-blocks and adds its partial with atomics
  const int kb_all = (K + GT_BK - 1) / GT_BK;
  const int kb_first = blockIdx.z * kb_per_split;
  const int num_k = min(kb_per_split, kb_all - kb_first);
  const bool atomic_out = gridDim.z > 1;
  a_koff += kb_first * GT_BK;
  b_koff += kb_first * GT_BK;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty_bar + s, ph ^ 1);
        mbar_expect_tx(full_bar + s, L::STAGE_BYTES);
        uint8_t* st = smem + s * L::STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < GT_BM / 64; ++i) tma_load_2d(&tmA, full_bar + s, st + i * BLK, m0 + 64 * i, a_koff + kb * GT_BK);
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) tma_load_2d(&tmB, full_bar + s, st + L::A_BYTES + i * BLK, n0 + 64 * i, b_koff + kb * GT_BK);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GT_BM, BN, 1, 1);
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(full_bar + s, ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
        const uint64_t da = umma_desc_mn128(a_addr, BLK);
        const uint64_t db = umma_desc_mn128(a_addr + L::A_BYTES, BLK);
#pragma unroll
        for (int k = 0; k < GT_BK / 16; ++k)   // 16 k rows = 2 swizzle atoms = 2048 B per K=16 step
          mma_bf16_ss(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, (kb | k) != 0);
        mma_commit(empty_bar + s);
      }
      mma_commit(tmem_full);
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* stage = reinterpret_cast<float*>(smem) + ew * 1024;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= N) break;
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (atomic_out) epi_chunk<true>(stage, v, C, ldc, m0 + ew * 32, n0 + c0, M, N, nullptr, 0, 0);
      else epi_chunk<false>(stage, v, C, ldc, m0 + ew * 32, n0 + c0, M, N, nullptr, accumulate, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<BN>(tmem_base);
}

// ---- host: tensor maps ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 tensor [rows, cols] with row pitch `ld` elements; box = {64 cols, box_rows}; 128-B swizzle
int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  SSASR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  SSASR_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld * 2) % 16 == 0, "TMA operand needs a 16-byte aligned base and row pitch (ld=%lld)", ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)GT_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSASR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
  return 0;
}

// 2-D bf16 tensor [rows, cols] (row pitch ld) read as MN-major operand blocks: box = {64 cols, 64 rows}
int make_tmap_bf16_mn(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld) {
  return make_tmap_bf16(m, ptr, rows, cols, ld, GT_BK);
}

// C[M,N] fp32 (+)= A^T B;  A stored [>= a_koff+K rows, M cols] (lda), B stored [>= b_koff+K rows, N cols] (ldb), bf16
int gemm_bf16_tc_tn(cudaStream_t st, int M, int N, int K, const void* A, long long lda, int a_koff, const void* B, long long ldb,
                    int b_koff, float* C, int ldc, int accumulate) {
  if (M <= 0 || N <= 0) return 0;
  SSASR_REQUIRE(K > 0, "gemm_bf16_tc_tn: K must be positive");
  constexpr int BN = 128, STAGES = 3;      // 3 x 32 KB stages: two CTAs per SM, one's epilogue overlaps the other's main loop
  using L = GemmSmem<BN, STAGES>;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_mn(&tmA, A, (long long)a_koff + K, M, lda);
  if (rc) return rc;
  rc = make_tmap_bf16_mn(&tmB, B, (long long)b_koff + K, N, ldb);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_tn_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL + 1024));
    attr_set = true;
  }
  dim3 grid((M + GT_BM - 1) / GT_BM, (N + BN - 1) / BN);
  SSASR_REQUIRE(grid.y <= 65535, "gemm_bf16_tc_tn: N=%d too large", N);
  // split the (usually very long) reduction over grid.z when the output has too few tiles to fill the GPU
  const int tiles = grid.x * grid.y, kb_all = (K + GT_BK - 1) / GT_BK;
  int splits = (2 * sm_count()) / tiles;
  if (splits > kb_all / 16) splits = kb_all / 16;
  if (splits < 1) splits = 1;
  const int kb_per = (kb_all + splits - 1) / splits;
  splits = (kb_all + kb_per - 1) / kb_per;
  grid.z = splits;
  if (splits > 1 && !accumulate) {
    if (ldc == N) {
      SSASR_CHECK_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
    } else {
      SSASR_CHECK_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st));
    }
  }
  ProfScope ps(F_GEMM_TC, st);
  gemm_tc_tn_kernel<BN, STAGES><<<grid, 256, L::TOTAL + 1024, st>>>(tmA, tmB, C, ldc, M, N, K, a_koff, b_koff, accumulate, kb_per);
  SSASR_LAUNCH_CHECK();
  return 0;
}

static int make_tmap_f32(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld) {
  EncodeTiledFn enc = get_encode();
  SSASR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  SSASR_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld * 4) % 16 == 0, "TMA operand needs a 16-byte aligned base and row pitch (ld=%lld)", ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSASR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(f32) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
  return 0;
}

__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16(v);
    hi[i] = h;
    lo[i] = __float2bfloat16(v - __bfloat162float(h));
  }
}
// x = hi + lo with hi, lo in bf16 (relative residual 2^-17)
int split_bf16(cudaStream_t st, const float* x, void* hi, void* lo, size_t n) {
  if (n == 0) return 0;
  ProfScope ps(F_PACK, st);
  split_bf16_kernel<<<592, 256, 0, st>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
  SSASR_LAUNCH_CHECK();
  return 0;
}

// Split-operand product as ONE plain bf16 GEMM over a three times longer reduction axis:
//   A B^T ~= A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T = [A_hi | A_hi | A_lo] [B_hi | B_lo | B_hi]^T
// (x = hi + lo in bf16, residual 2^-17 relative; the dropped lo*lo term is 2^-18): the persistent tcgen05 kernel runs it at the
// bf16 rate (2 x the tf32 rate of gemm_tf32x3_kernel) with its double-buffered accumulators.  This kernel writes the
// concatenated operand: rows of 3K bf16, `second` selects the [hi | lo | hi] order of the B side.  K % 8 == 0.
__global__ void split3_bf16_kernel(const float* __restrict__ x, long long ld, long long rows, int K, __nv_bfloat16* __restrict__ out,
                                   int second) {
  const int kc = K / 8;
  const long long n = rows * kc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / kc;
    const int c = (int)(i % kc) * 8;
    const float4 v0 = *reinterpret_cast<const float4*>(x + r * ld + c), v1 = *reinterpret_cast<const float4*>(x + r * ld + c + 4);
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 h0 = __float2bfloat16(v[2 * j]), h1 = __float2bfloat16(v[2 * j + 1]);
      const __nv_bfloat16 l0 = __float2bfloat16(v[2 * j] - __bfloat162float(h0)), l1 = __float2bfloat16(v[2 * j + 1] - __bfloat162float(h1));
      h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      l[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const uint4 H = make_uint4(h[0], h[1], h[2], h[3]), L = make_uint4(l[0], l[1], l[2], l[3]);
    __nv_bfloat16* o = out + r * 3 * K + c;
    *reinterpret_cast<uint4*>(o) = H;
    *reinterpret_cast<uint4*>(o + K) = second ? L : H;
    *reinterpret_cast<uint4*>(o + 2 * K) = second ? H : L;
  }
}
int split3_bf16(cudaStream_t st, const float* x, long long ld, long long rows, int K, void* out, int second) {
  if (rows <= 0 || K <= 0) return 0;
  SSASR_REQUIRE(K % 8 == 0 && ld % 4 == 0, "split3_bf16: K %% 8 == 0 and ld %% 4 == 0 required (K=%d ld=%lld)", K, ld);
  ProfScope ps(F_PACK, st);
  const long long n = rows * (K / 8);
  split3_bf16_kernel<<<(unsigned)((n + 255) / 256 > 2368 ? 2368 : (n + 255) / 256), 256, 0, st>>>(x, ld, rows, K, (__nv_bfloat16*)out, second);
  SSASR_LAUNCH_CHECK();
  return 0;
}
// 1: split-operand products run as bf16 GEMMs over [hi | hi | lo] x [hi | lo | hi] (default); 0 (SSASR_X3_GEMM=tf32): tf32 x 3 kernel
int x3_gemm_bf16() {
  const char* e = getenv("SSASR_X3_GEMM");
  return (e && e[0] == 't') ? 0 : 1;
}

__global__ void split_hi_lo_2d_kernel(const float* __restrict__ x, long long ld, int rows, int cols, float* __restrict__ hi,
                                      float* __restrict__ lo) {
  const long long n = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[(i / cols) * ld + (i % cols)];
    const float h = rn_tf32(v);
    hi[i] = h;
    lo[i] = rn_tf32(v - h);
  }
}
// strided rows [rows, cols] (pitch ld) -> dense tf32-exact hi / lo
int split_hi_lo_2d(cudaStream_t st, const float* x, long long ld, int rows, int cols, float* hi, float* lo) {
  if (rows <= 0 || cols <= 0) return 0;
  ProfScope ps(F_PACK, st);
  const long long n = (long long)rows * cols;
  split_hi_lo_2d_kernel<<<(unsigned)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256), 256, 0, st>>>(x, ld, rows, cols, hi, lo);
  SSASR_LAUNCH_CHECK();
  return 0;
}

int split_hi_lo(cudaStream_t st, const float* x, float* hi, float* lo, size_t n) {
  if (n == 0) return 0;
  ProfScope ps(F_PACK, st);
  split_hi_lo_kernel<<<1184, 256, 0, st>>>(x, hi, lo, n);
  SSASR_LAUNCH_CHECK();
  return 0;
}

// C[M,N] fp32 = A[M,K] x B[N,K]^T (+bias)(+tanh); operands given as tf32-exact (hi, lo) pairs (K contiguous, ld % 4 == 0)
int gemm_tf32x3(cudaStream_t st, int M, int N, int K, const float* A, const float* A_lo, long long lda, const float* B,
                const float* B_lo, long long ldb, float* C, int ldc, const float* bias, int act_tanh) {
  if (M <= 0 || N <= 0) return 0;
  constexpr int STAGES = 3;
  constexpr size_t SMEM = (size_t)STAGES * 4 * 128 * 128 + (2 * STAGES + 1) * 8 + 16 + 1024;
  CUtensorMap tA, tAl, tB, tBl;
  int rc = make_tmap_f32(&tA, A, M, K, lda);
  if (rc) return rc;
  rc = make_tmap_f32(&tAl, A_lo, M, K, lda);
  if (rc) return rc;
  rc = make_tmap_f32(&tB, B, N, K, ldb);
  if (rc) return rc;
  rc = make_tmap_f32(&tBl, B_lo, N, K, ldb);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    attr_set = true;
  }
  dim3 grid((N + 127) / 128, (M + GT_BM - 1) / GT_BM);
  SSASR_REQUIRE(grid.y <= 65535, "gemm_tf32x3: M=%d too large", M);
  ProfScope ps(F_GEMM_TC, st);
  gemm_tf32x3_kernel<STAGES><<<grid, 256, SMEM, st>>>(tA, tAl, tB, tBl, C, ldc, bias, M, N, K, act_tanh);
  SSASR_LAUNCH_CHECK();
  return 0;
}

// 3-D bf16 tensor: dim0 = cols (contiguous), dim1 = nA entries strideA elements apart, dim2 = nB entries strideB
// elements apart (strideA <= strideB expected); box = {64, boxA, boxB}; 128-B swizzle.
int make_tmap_bf16_3d(CUtensorMap* m, const void* ptr, long long cols, long long nA, long long strideA, long long nB,
                      long long strideB, int boxA, int boxB) {
  EncodeTiledFn enc = get_encode();
  SSASR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  SSASR_REQUIRE(((uintptr_t)ptr & 15) == 0 && (strideA * 2) % 16 == 0 && (strideB * 2) % 16 == 0,
                "TMA operand needs 16-byte aligned base and strides");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)nA, (cuuint64_t)nB};
  cuuint64_t strides[2] = {(cuuint64_t)strideA * 2, (cuuint64_t)strideB * 2};
  cuuint32_t box[3] = {(cuuint32_t)GT_BK, (cuuint32_t)boxA, (cuuint32_t)boxB};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSASR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed (%d) cols=%lld nA=%lld sA=%lld nB=%lld sB=%lld", (int)r, cols,
                nA, strideA, nB, strideB);
  return 0;
}

// Same 3-D view with a caller-chosen box width and swizzle (box_cols * 2 bytes must equal swizzle_bytes: 32 / 64 / 128):
// the cluster recurrent kernels multicast a 16-column (32 B, SWIZZLE_32B) slice of h per producer CTA.
int make_tmap_bf16_3d_ex(CUtensorMap* m, const void* ptr, long long cols, long long nA, long long strideA, long long nB,
                         long long strideB, int box_cols, int boxA, int boxB, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  SSASR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  SSASR_REQUIRE(((uintptr_t)ptr & 15) == 0 && (strideA * 2) % 16 == 0 && (strideB * 2) % 16 == 0,
                "TMA operand needs 16-byte aligned base and strides");
  SSASR_REQUIRE(box_cols * 2 == swizzle_bytes && (swizzle_bytes == 32 || swizzle_bytes == 64 || swizzle_bytes == 128),
                "make_tmap_bf16_3d_ex: box of %d columns does not match a %d-byte swizzle", box_cols, swizzle_bytes);
  const CUtensorMapSwizzle sw = swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)nA, (cuuint64_t)nB};
  cuuint64_t strides[2] = {(cuuint64_t)strideA * 2, (cuuint64_t)strideB * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)boxA, (cuuint32_t)boxB};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSASR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d ex) failed (%d) cols=%lld nA=%lld sA=%lld nB=%lld sB=%lld", (int)r,
                cols, nA, strideA, nB, strideB);
  return 0;
}

template <int BN, int STAGES>
static int launch_gemm_tc(cudaStream_t st, int M, int N, int K, const void* A, long long lda, int a_koff, const void* B, long long ldb,
                          int b_koff, float* C, int ldc, const float* bias, int accumulate, int act_tanh, int splits) {
  using L = GemmSmem<BN, STAGES>;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16(&tmA, A, M, (long long)a_koff + K, lda, GT_BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, B, N, (long long)b_koff + K, ldb, BN);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL + 1024));
    attr_set = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + GT_BM - 1) / GT_BM);
  SSASR_REQUIRE(grid.y <= 65535, "gemm_bf16_tc: M=%d too large", M);
  const int kb_all = (K + GT_BK - 1) / GT_BK;
  if (splits > kb_all) splits = kb_all;
  if (splits < 1) splits = 1;
  const int kb_per = (kb_all + splits - 1) / splits;
  grid.z = (kb_all + kb_per - 1) / kb_per;
  SSASR_REQUIRE(grid.z == 1 || (!accumulate && !act_tanh), "gemm_bf16_tc: split-K adds into a pre-zeroed output (no accumulate / tanh)");
  ProfScope ps(F_GEMM_TC, st);
  gemm_tc_kernel<BN, STAGES><<<grid, 256, L::TOTAL + 1024, st>>>(tmA, tmB, C, ldc, bias, M, N, K, a_koff, b_koff, accumulate, act_tanh,
                                                                  kb_per);
  SSASR_LAUNCH_CHECK();
  return 0;
}

static int gemm_persist_on() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSASR_GEMM_PERSIST");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}
static int launch_gemm_tc_persist(cudaStream_t st, int M, int N, int K, const void* A, long long lda, int a_koff, const void* B,
                                  long long ldb, int b_koff, float* C, int ldc, const float* bias, int accumulate, int act_tanh) {
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16(&tmA, A, M, (long long)a_koff + K, lda, GT_BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, B, N, (long long)b_koff + K, ldb, GP_BN);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_TOTAL + 1024));
    attr_set = true;
  }
  const int n_mt = (M + GT_BM - 1) / GT_BM, n_nt = (N + GP_BN - 1) / GP_BN;
  const int total = n_mt * n_nt;
  const int grid = total < sm_count() ? total : sm_count();
  ProfScope ps(F_GEMM_TC, st);
  gemm_tc_persist_kernel<<<grid, GP_THREADS, GP_TOTAL + 1024, st>>>(tmA, tmB, C, ldc, bias, M, N, K, a_koff, b_koff, accumulate, act_tanh,
                                                             n_mt, n_nt);
  SSASR_LAUNCH_CHECK();
  return 0;
}

int gemm_bf16_tc(cudaStream_t st, int M, int N, int K, const void* A, long long lda, int a_koff, const void* B, long long ldb,
                 int b_koff, float* C, int ldc, const float* bias, int accumulate, int act_tanh, int splits) {
  if (M <= 0 || N <= 0) return 0;
  SSASR_REQUIRE(K > 0, "gemm_bf16_tc: K must be positive");
  SSASR_REQUIRE(a_koff % 8 == 0 && b_koff % 8 == 0, "gemm_bf16_tc: reduction offsets must be multiples of 8 elements (TMA 16-byte "
                "box alignment), got %d / %d", a_koff, b_koff);
  // 128 x 128 tiles, 3 x 32 KB stages: two CTAs per SM, so one CTA's epilogue overlaps the other's main loop.
  // Small products (the per-step speller GEMMs, 256 rows) use 128 x 32 tiles to spread over 4x as many SMs.
  const long long tiles128 = (long long)((M + GT_BM - 1) / GT_BM) * ((N + 127) / 128);
  if (tiles128 < sm_count() / 2 && N >= 64)
    return launch_gemm_tc<32, 4>(st, M, N, K, A, lda, a_koff, B, ldb, b_koff, C, ldc, bias, accumulate, act_tanh, splits);
  // large products: persistent 128 x 256 tiles with double-buffered accumulators (at least two tiles per SM, full-width tiles)
  const long long tiles256 = (long long)((M + GT_BM - 1) / GT_BM) * ((N + GP_BN - 1) / GP_BN);
  if (gemm_persist_on() && tiles256 >= 2LL * sm_count() && N % 64 == 0 && N >= GP_BN && K >= 2 * GT_BK)
    return launch_gemm_tc_persist(st, M, N, K, A, lda, a_koff, B, ldb, b_koff, C, ldc, bias, accumulate, act_tanh);
  return launch_gemm_tc<128, 3>(st, M, N, K, A, lda, a_koff, B, ldb, b_koff, C, ldc, bias, accumulate, act_tanh, 1);
}

// ---- fp32 -> bf16 conversion (optionally transposed, optionally masking a periodic column) ----------
__global__ void cvt_bf16_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst, long long ld_dst,
                                long long rows, int cols) {
  const long long total = rows * (long long)((cols + 1) / 2);
  const int half = (cols + 1) / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / half;
    const int c = (int)(i % half) * 2;
    const float a = src[r * ld_src + c];
    const float b = (c + 1 < cols) ? src[r * ld_src + c + 1] : 0.f;
    if (c + 1 < cols || (ld_dst & 1) == 0) {
      *reinterpret_cast<__nv_bfloat162*>(dst + r * ld_dst + c) = __floats2bfloat162_rn(a, b);
    } else {
      dst[r * ld_dst + c] = __float2bfloat16(a);
    }
  }
}

// dst[c, r + shift(c)] = bf16(src[r, c]) for r < rows, c < cols; dst row pitch ld_dst; shift(c) = shift_lo for
// c < mask_split, shift_hi otherwise (destinations outside [0, rows) are dropped; the caller pre-zeroes dst).
// mask_period > 0: source rows with (r % mask_period) == mask_pos_lo are written as 0 for c < mask_split, and rows
// with (r % mask_period) == mask_pos_hi are written as 0 for c >= mask_split (drops the frame that has no
// forward-order predecessor from the recurrent-weight gradient; see ssasr_blstm_bwd_bf16).
__global__ void cvt_bf16_t_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst,
                                  long long ld_dst, long long rows, int cols, int mask_period, int mask_pos_lo, int mask_pos_hi,
                                  int mask_split, int shift_lo, int shift_hi) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long r = r0 + i;
    const int c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = src[r * ld_src + c];
      if (mask_period > 0) {
        const int ph = (int)(r % mask_period);
        if ((c < mask_split && ph == mask_pos_lo) || (c >= mask_split && ph == mask_pos_hi)) v = 0.f;
      }
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long r = r0 + threadIdx.x;
    const long long rd = r + (c < mask_split ? shift_lo : shift_hi);
    if (c < cols && r < rows && rd >= 0 && rd < rows) dst[(long long)c * ld_dst + rd] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

// zeroes, in a bf16 [rows, cols] matrix, the rows with (r % period) == pos_lo for columns < split and the rows with
// (r % period) == pos_hi for columns >= split
__global__ void mask_rows_bf16_kernel(__nv_bfloat16* __restrict__ x, long long rows, int cols, int period, int pos_lo, int pos_hi,
                                      int split) {
  const long long nper = (rows + period - 1) / period;
  const long long total = nper * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long u = i / cols;
    const int c = (int)(i % cols);
    const long long r = u * period + (c < split ? pos_lo : pos_hi);
    if (r < rows) x[r * cols + c] = __float2bfloat16(0.f);
  }
}
int mask_rows_bf16(cudaStream_t st, void* x, long long rows, int cols, int period, int pos_lo, int pos_hi, int split) {
  if (period <= 0 || rows <= 0) return 0;
  ProfScope ps(F_PACK, st);
  mask_rows_bf16_kernel<<<256, 256, 0, st>>>((__nv_bfloat16*)x, rows, cols, period, pos_lo, pos_hi, split);
  SSASR_LAUNCH_CHECK();
  return 0;
}

int cvt_bf16(cudaStream_t st, const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  ProfScope ps(F_PACK, st);
  const long long pairs = rows * (long long)((cols + 1) / 2);
  const int blocks = (int)(pairs / 256 + 1 < 1184 ? pairs / 256 + 1 : 1184);      // small operands: no idle CTAs to schedule
  cvt_bf16_kernel<<<blocks, 256, 0, st>>>(src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols);
  SSASR_LAUNCH_CHECK();
  return 0;
}
int cvt_bf16_t(cudaStream_t st, const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols,
               int mask_period, int mask_pos_lo, int mask_pos_hi, int mask_split, int shift_lo, int shift_hi) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid((unsigned)((rows + 31) / 32), (cols + 31) / 32);
  SSASR_REQUIRE(grid.y <= 65535, "cvt_bf16_t: too many columns (%d)", cols);
  ProfScope ps(F_PACK, st);
  cvt_bf16_t_kernel<<<grid, dim3(32, 8), 0, st>>>(src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols, mask_period, mask_pos_lo,
                                                  mask_pos_hi, mask_split, shift_lo, shift_hi);
  SSASR_LAUNCH_CHECK();
  return 0;
}

}  // namespace ssasr

using namespace ssasr;

extern "C" {

int ssasr_gemm_bf16_tc(int M, int N, int K, const void* A, long long lda, int a_koff, const void* B, long long ldb, int b_koff,
                       float* C, int ldc, const float* bias, int accumulate, void* stream) {
  return gemm_bf16_tc((cudaStream_t)stream, M, N, K, A, lda, a_koff, B, ldb, b_koff, C, ldc, bias, accumulate, 0);
}
// fp32-accurate tensor-core GEMM (tf32 x 3); A_ws / B_ws are scratch buffers of TWICE the operands' sizes (hi | lo)
int ssasr_gemm_tf32x3(int M, int N, int K, const float* A, float* A_ws, long long lda, const float* B, float* B_ws, long long ldb,
                      float* C, int ldc, const float* bias, int act_tanh, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_REQUIRE(lda == K && ldb == K && K % 4 == 0, "ssasr_gemm_tf32x3: operands must be dense (lda == ldb == K) with K %% 4 == 0");
  float* A_lo = A_ws + (size_t)M * K;
  float* B_lo = B_ws + (size_t)N * K;
  int rc = split_hi_lo(st, A, A_ws, A_lo, (size_t)M * K);
  if (rc) return rc;
  rc = split_hi_lo(st, B, B_ws, B_lo, (size_t)N * K);
  if (rc) return rc;
  return gemm_tf32x3(st, M, N, K, A_ws, A_lo, lda, B_ws, B_lo, ldb, C, ldc, bias, act_tanh);
}
int ssasr_gemm_bf16_tc_tn(int M, int N, int K, const void* A, long long lda, int a_koff, const void* B, long long ldb, int b_koff,
                          float* C, int ldc, int accumulate, void* stream) {
  return gemm_bf16_tc_tn((cudaStream_t)stream, M, N, K, A, lda, a_koff, B, ldb, b_koff, C, ldc, accumulate);
}
int ssasr_cvt_bf16(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols, void* stream) {
  return cvt_bf16((cudaStream_t)stream, src, ld_src, dst, ld_dst, rows, cols);
}
int ssasr_cvt_bf16_t(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols, int mask_period,
                     int mask_pos_lo, int mask_pos_hi, int mask_split, void* stream) {
  return cvt_bf16_t((cudaStream_t)stream, src, ld_src, dst, ld_dst, rows, cols, mask_period, mask_pos_lo, mask_pos_hi, mask_split,
                    0, 0);
}

}  // extern "C"

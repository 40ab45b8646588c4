// Tensor-core recurrent BLSTM kernels (bf16 training path): the per-step recurrent gate product runs on
// tcgen05 with the W_hh slice resident in shared memory for the whole sequence.
//
//   forward  step:  G[n, 64 gate cols] = h_{t-1}[n, :S] (bf16, TMA from L2)  x  W_hh[slice]^T   (+ xp, cell update)
//   backward step:  dh[n, 16 units]    = dG_{t+1}[n, :4S] (bf16, TMA from L2) x  W_hh[:, slice]  (+ dhout, cell grad)
//
// One persistent cooperative launch per layer and pass.  CTA = (slice of 16 hidden units, direction, batch-tile
// group); each CTA owns 128-row batch tiles, accumulates in TMEM (M=128 lanes = batch rows, so a thread owns one
// batch row and all four gates of its units: the cell math needs no cross-thread exchange), and the CTAs of one
// (direction, tile group) meet at a monotonic-counter barrier once per time step; h / dG are exchanged through
// L2 as bf16 row buffers that the next step's TMA loads read back.
//
// Reference semantics: nn.LSTM(bidirectional) packed (asr.py:410-418) and blstm_4 (asr.py:262); same masking and
// row-stride conventions as lstm_rec.cu.  Replaces the cuDNN per-step recurrent GEMM + pointwise kernels.
#include "common.cuh"
#include "tc_common.cuh"

namespace ssasr {

using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows);
int make_tmap_bf16_3d(CUtensorMap* m, const void* ptr, long long cols, long long nA, long long strideA, long long nB,
                      long long strideB, int boxA, int boxB);

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 2.f * sigmoid_fast(2.f * x) - 1.f; }

__device__ __forceinline__ void group_barrier(unsigned* counter, unsigned target, long long* stamp = nullptr) {
  fence_proxy_async_all();          // generic-proxy global stores -> visible to other CTAs' TMA (async proxy) reads
  __syncthreads();
  if (threadIdx.x == 0) {
    if (stamp) stamp[6] = clock64();
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (++spins > (1u << 26)) __trap();
    } while (v < target);
    fence_proxy_async_all();
    if (stamp) stamp[7] = clock64();
  }
  __syncthreads();
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps
__device__ __forceinline__ float tanh_apx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_apx(float x) { return fmaf(tanh_apx(0.5f * x), 0.5f, 0.5f); }

// Step barrier, split: the epilogue ARRIVES as soon as this CTA's slice of the exchange buffer is written; only
// the TMA producer WAITS (it is the only consumer of other CTAs' data), so the rest of the epilogue (saved-tensor
// write-out, next-step prefetch) overlaps the barrier latency and the next step's TMA + MMA.
__device__ __forceinline__ void step_arrive(unsigned* counter) {
  // release: orders this CTA's exchange-buffer stores (made visible to this thread by the preceding bar.sync) before
  // the counter increment; no L1 invalidation needed on the producer side
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
}
__device__ __forceinline__ void step_wait(unsigned* counter, unsigned target) {
  // poll with a relaxed load (an acquire load in the loop makes ptxas emit CCTL.IVALL + MEMBAR.GPU per iteration,
  // which stalls the epilogue warps' global traffic); one acquire fence once the target is reached
  unsigned v, spins = 0;
  do {
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (++spins > (1u << 26)) __trap();
  } while (v < target);
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  fence_proxy_async_all();          // acquired generic-proxy writes -> ordered before this thread's TMA (async proxy) reads
}

constexpr int RT_UNITS = 16;           // hidden units per CTA slice
constexpr int RT_THREADS = 192;        // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue

struct RecTcParams {
  float* xp;                 // fwd: [rows,8S] pre-activations in / activations out.  bwd: activations in / dG out
  float* hout;               // fwd out [rows,2S] fp32
  float* cbuf;               // fwd out / bwd in
  __nv_bfloat16* xb;         // fwd: h exchange buffer [rows,2S] bf16.  bwd: dG exchange buffer [rows,8S] bf16
  __nv_bfloat16* xb2;        // fwd, X3 mode: low part of the h exchange buffer (h = hi + lo, both bf16)
  const float* dhout;        // bwd
  float* dcstate;            // bwd [n_batch,2S]
  float* dbias;              // bwd: [8S] bias gradient (column sums of dG), pre-zeroed, accumulated with atomics; may be null
  const int* lens;
  int S, n_seq, n_batch, n_tiles;
  long long rs_seq, rs_batch;
  int seq_inner;             // 1: tensor-map dim1 = seq, dim2 = batch (time-major layers); 0: dim1 = batch, dim2 = seq
  unsigned* bar;             // [2 * gridDim.z]
  long long* dbg;            // optional [n_seq][12] clock64 stamps of CTA (0,0,0); null = off
};

constexpr int XS_P = 68;    // fp32 row pitch of the 128 x 64 gate tile in smem (conflict-free float4 row access)
constexpr int HS_P = 20;    // fp32 row pitch of the 128 x 16 state tiles
constexpr int HB_P = 24;    // bf16 row pitch of the 128 x 16 h exchange tile
constexpr int GB_P = 72;    // bf16 row pitch of the 128 x 64 dG exchange tile

static long long* g_dbg = nullptr;
#define DBG_STAMP(idx)                                                                   \
  do {                                                                                   \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg[(size_t)s * 12 + (idx)] = clock64(); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// Epilogue data movement is staged through shared memory so that every global access is a contiguous 64-256 B
// row segment: the xp tile of the NEXT step is prefetched with cp.async while the step barrier / TMA / MMA of
// that step are in flight, the thread-per-batch-row cell math works on smem, results leave as coalesced stores.
// X3: fp32-accurate mode for the exact (validation / decode) path: h and W_hh are split into bf16 hi + lo parts and the
// product is hi*hi + hi*lo + lo*hi (3 MMAs per K step); precise expf/tanhf activations.
template <int TM, int KB, bool X3>   // TM = batch rows per tile (64 or 128), KB = S / 64 resident k-blocks
__global__ void __launch_bounds__(RT_THREADS, 1)
rec_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmH2, const __grid_constant__ CUtensorMap tmW2, RecTcParams p) {
  constexpr int NP = X3 ? 2 : 1;         // operand parts (hi [, lo])
  constexpr int W_BLK = 64 * 128;        // 64 gate rows x 128 B
  constexpr int A_BLK = TM * 128;        // TM batch rows x 128 B
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Wsm = smem;
  uint8_t* Asm = smem + NP * KB * W_BLK;                                 // parts are laid out [part][kb]
  float* xs = reinterpret_cast<float*>(Asm + NP * KB * A_BLK);            // [TM][XS_P]
  float* hst = xs + TM * XS_P;                                           // [128][HS_P]
  float* cst = hst + TM * HS_P;                                          // [128][HS_P]
  __nv_bfloat16* hbst = reinterpret_cast<__nv_bfloat16*>(cst + TM * HS_P);   // [NP][TM][HB_P]
  uint64_t* bars = reinterpret_cast<uint64_t*>(hbst + NP * TM * HB_P);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* mma_done = bars + 2;
  uint64_t* tmem_free = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int S = p.S;
  const int slice = blockIdx.x, dir = blockIdx.y, z = blockIdx.z, Z = gridDim.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1);
    mbar_init(a_full, 1);
    mbar_init(mma_done, 1);
    mbar_init(tmem_free, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<64>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(w_full, NP * KB * W_BLK);
    for (int kb = 0; kb < KB; ++kb) tma_load_2d(&tmW, w_full, Wsm + kb * W_BLK, kb * 64, dir * 4 * S + slice * 64);
    if (X3)
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(&tmW2, w_full, Wsm + (KB + kb) * W_BLK, kb * 64, dir * 4 * S + slice * 64);
  }
  unsigned* gbar = p.bar + dir * Z + z;
  const unsigned G = gridDim.x;
  uint32_t it = 0;                         // completed (tile, s>0) iterations: phase of a_full / mma_done
  bool w_ready = false;
  const int eg = warp & 3;                 // TMEM lane group of an epilogue warp
  const int te = threadIdx.x - 64;         // epilogue thread id 0..127 (warps 2-5)
  const bool single = p.n_tiles <= Z;      // one batch tile per CTA: cell state lives in registers
  constexpr uint32_t idesc = umma_idesc_bf16(TM, 64);
  float creg[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) creg[j] = 0.f;

  // coalesced prefetch of the xp tile of (time t_, tile bt_) into xs
  auto prefetch_x = [&](int t_, int bt_) {
#pragma unroll 4
    for (int i = 0; i < TM / 8; ++i) {
      const int idx = i * 128 + te, r = idx >> 4, c4 = idx & 15;
      const int n = bt_ * TM + r;
      float* dst = xs + r * XS_P + c4 * 4;
      if (n < p.n_batch)
        cp_async16(dst, p.xp + ((size_t)t_ * p.rs_seq + (size_t)n * p.rs_batch) * 8 * S + (size_t)dir * 4 * S + slice * 64 + c4 * 4);
      else
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };
  if (warp >= 2) prefetch_x(dir == 0 ? 0 : p.n_seq - 1, z);

  for (int s = 0; s < p.n_seq; ++s) {
    const int t = dir == 0 ? s : p.n_seq - 1 - s;
    const int tp = dir == 0 ? t - 1 : t + 1;
    for (int bt = z; bt < p.n_tiles; bt += Z) {
      if (s > 0) {
        if (warp == 0) {
          if (elect_one()) {
            if (bt == z) step_wait(gbar, (unsigned)s * G);       // every CTA of the group has published step s-1
            if (it > 0) mbar_wait(mma_done, (it - 1) & 1);       // previous MMAs have consumed the A tile
            DBG_STAMP(0);
            mbar_expect_tx(a_full, NP * KB * A_BLK);
            for (int kb = 0; kb < KB; ++kb)
              tma_load_3d(&tmH, a_full, Asm + kb * A_BLK, dir * S + kb * 64, p.seq_inner ? tp : bt * TM, p.seq_inner ? bt * TM : tp);
            if (X3)
              for (int kb = 0; kb < KB; ++kb)
                tma_load_3d(&tmH2, a_full, Asm + (KB + kb) * A_BLK, dir * S + kb * 64, p.seq_inner ? tp : bt * TM,
                            p.seq_inner ? bt * TM : tp);
          }
        } else if (warp == 1) {
          if (elect_one()) {
            if (!w_ready) { mbar_wait(w_full, 0); w_ready = true; }
            mbar_wait(a_full, it & 1);
            if (it > 0) mbar_wait(tmem_free, (it - 1) & 1);      // epilogue has drained the accumulator
            DBG_STAMP(1);
            tc_fence_after();
            const uint32_t a0 = smem_u32(Asm), w0 = smem_u32(Wsm);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
              const uint64_t da = umma_desc_k128(a0 + kb * A_BLK), db = umma_desc_k128(w0 + kb * W_BLK);
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_bf16_ss(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              if (X3) {
                const uint64_t dal = umma_desc_k128(a0 + (KB + kb) * A_BLK), dbl = umma_desc_k128(w0 + (KB + kb) * W_BLK);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  mma_bf16_ss(tmem, da + (uint64_t)(k * 2), dbl + (uint64_t)(k * 2), idesc, 1);
                  mma_bf16_ss(tmem, dal + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1);
                }
              }
            }
            mma_commit(mma_done);
            DBG_STAMP(2);
          }
        }
      }
      if (warp >= 2) {
        // tile row <-> TMEM lane: M=128 uses all 32 lanes of each sub-partition, M=64 the lower 16 (rows 16q..16q+15)
        const int r = (TM == 128) ? eg * 32 + lane : eg * 16 + (lane & 15);
        const int n = bt * TM + r;
        const bool in_range = (TM == 128 || lane < 16) && n < p.n_batch;
        const bool valid = in_range && (p.lens ? (t < p.lens[in_range ? n : 0]) : true);
        const size_t hoff = (size_t)dir * S + slice * RT_UNITS;
        cp_async_wait_all();
        epi_bar();                                // xs holds the xp tile of (t, bt)
        float cpv[16];
        if (single) {
#pragma unroll
          for (int j = 0; j < 16; ++j) cpv[j] = creg[j];
        } else {
          const size_t rowp = (size_t)(s > 0 ? tp : t) * p.rs_seq + (size_t)(in_range ? n : 0) * p.rs_batch;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 c4 = (s > 0 && valid) ? *(reinterpret_cast<const float4*>(p.cbuf + rowp * 2 * S + hoff) + q)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
            cpv[q * 4 + 0] = c4.x; cpv[q * 4 + 1] = c4.y; cpv[q * 4 + 2] = c4.z; cpv[q * 4 + 3] = c4.w;
          }
        }
        uint32_t v[64];
        if (s > 0) {
          mbar_wait(mma_done, it & 1);
          if (warp == 2 && lane == 0) DBG_STAMP(3);
          tc_fence_after();
          const uint32_t ta = tmem + ((uint32_t)(eg * 32) << 16);
          tmem_ld16(ta, v);
          tmem_ld16(ta + 16, v + 16);
          tmem_ld16(ta + 32, v + 32);
          tmem_ld16(ta + 48, v + 48);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(tmem_free);
          if (warp == 2 && lane == 0) DBG_STAMP(4);
        } else {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = 0u;
        }
        float* xrow = xs + r * XS_P;
        if (TM == 128 || lane < 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float hv[4], cv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = q * 4 + u;
            float4 g = *reinterpret_cast<const float4*>(xrow + j * 4);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            hv[u] = 0.f; cv[u] = 0.f;
            if (valid) {
              g.x += __uint_as_float(v[j * 4 + 0]); g.y += __uint_as_float(v[j * 4 + 1]);
              g.z += __uint_as_float(v[j * 4 + 2]); g.w += __uint_as_float(v[j * 4 + 3]);
              if (X3) {
                a.x = sigmoidf_acc(g.x); a.y = sigmoidf_acc(g.y); a.z = tanhf(g.z); a.w = sigmoidf_acc(g.w);
                cv[u] = a.y * cpv[j] + a.x * a.z;
                hv[u] = a.w * tanhf(cv[u]);
              } else {
                a.x = sigmoid_apx(g.x); a.y = sigmoid_apx(g.y); a.z = tanh_apx(g.z); a.w = sigmoid_apx(g.w);
                cv[u] = fmaf(a.y, cpv[j], a.x * a.z);
                hv[u] = a.w * tanh_apx(cv[u]);
              }
            }
            creg[j] = cv[u];
            *reinterpret_cast<float4*>(xrow + j * 4) = a;
          }
          *reinterpret_cast<float4*>(hst + r * HS_P + q * 4) = make_float4(hv[0], hv[1], hv[2], hv[3]);
          *reinterpret_cast<float4*>(cst + r * HS_P + q * 4) = make_float4(cv[0], cv[1], cv[2], cv[3]);
          __nv_bfloat162 b01 = __floats2bfloat162_rn(hv[0], hv[1]), b23 = __floats2bfloat162_rn(hv[2], hv[3]);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&b01);
          pk.y = *reinterpret_cast<uint32_t*>(&b23);
          *reinterpret_cast<uint2*>(hbst + r * HB_P + q * 4) = pk;
          if (X3) {      // low parts: h - float(bf16(h))
            const float l0 = hv[0] - __bfloat162float(b01.x), l1 = hv[1] - __bfloat162float(b01.y);
            const float l2 = hv[2] - __bfloat162float(b23.x), l3 = hv[3] - __bfloat162float(b23.y);
            __nv_bfloat162 c01 = __floats2bfloat162_rn(l0, l1), c23 = __floats2bfloat162_rn(l2, l3);
            uint2 pl;
            pl.x = *reinterpret_cast<uint32_t*>(&c01);
            pl.y = *reinterpret_cast<uint32_t*>(&c23);
            *reinterpret_cast<uint2*>(hbst + (TM + r) * HB_P + q * 4) = pl;
          }
        }
        }
        if (warp == 2 && lane == 0) DBG_STAMP(5);
        epi_bar();
        // coalesced write-out: exchange buffer first, then the saved tensors
        {
          const size_t rbase = (size_t)t * p.rs_seq;
#pragma unroll
          for (int i = 0; i < TM / 64; ++i) {
            const int idx = i * 128 + te, rr = idx >> 1, hh = idx & 1, nn = bt * TM + rr;
            if (nn < p.n_batch)
              *reinterpret_cast<uint4*>(p.xb + (rbase + (size_t)nn * p.rs_batch) * 2 * S + hoff + hh * 8) =
                  *reinterpret_cast<const uint4*>(hbst + rr * HB_P + hh * 8);
            if (X3 && nn < p.n_batch)
              *reinterpret_cast<uint4*>(p.xb2 + (rbase + (size_t)nn * p.rs_batch) * 2 * S + hoff + hh * 8) =
                  *reinterpret_cast<const uint4*>(hbst + (TM + rr) * HB_P + hh * 8);
          }
          if (bt + Z >= p.n_tiles && s + 1 < p.n_seq) {   // last tile of this step: publish
            epi_bar();
            if (te == 0) {
              step_arrive(gbar);
              DBG_STAMP(6);
            }
          }
#pragma unroll
          for (int i = 0; i < TM / 32; ++i) {
            const int idx = i * 128 + te, rr = idx >> 2, c4 = idx & 3, nn = bt * TM + rr;
            if (nn < p.n_batch) {
              const size_t o = (rbase + (size_t)nn * p.rs_batch) * 2 * S + hoff + c4 * 4;
              *reinterpret_cast<float4*>(p.hout + o) = *reinterpret_cast<const float4*>(hst + rr * HS_P + c4 * 4);
              *reinterpret_cast<float4*>(p.cbuf + o) = *reinterpret_cast<const float4*>(cst + rr * HS_P + c4 * 4);
            }
          }
          if (warp == 2 && lane == 0) DBG_STAMP(8);
#pragma unroll 4
          for (int i = 0; i < TM / 8; ++i) {
            const int idx = i * 128 + te, rr = idx >> 4, c4 = idx & 15, nn = bt * TM + rr;
            if (nn < p.n_batch)
              __stcs(reinterpret_cast<float4*>(p.xp + (rbase + (size_t)nn * p.rs_batch) * 8 * S + (size_t)dir * 4 * S + slice * 64 + c4 * 4),
                     *reinterpret_cast<const float4*>(xs + rr * XS_P + c4 * 4));
          }
        }
        if (warp == 2 && lane == 0) DBG_STAMP(9);
        epi_bar();                                // staging buffers drained
        if (warp == 2 && lane == 0) DBG_STAMP(10);
        // prefetch the xp tile of the next (step, tile) iteration
        {
          int bn = bt + Z, sn = s;
          if (bn >= p.n_tiles) { bn = z; sn = s + 1; }
          if (sn < p.n_seq) prefetch_x(dir == 0 ? sn : p.n_seq - 1 - sn, bn);
        }
        if (warp == 2 && lane == 0) DBG_STAMP(7);
      }
      if (s > 0) ++it;
    }
  }
  cp_async_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(tmem);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int TM, int NST>   // TM = batch rows per tile (64 or 128), NST = TMA ring stages
__global__ void __launch_bounds__(RT_THREADS, 1)
rec_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmW, RecTcParams p) {
  constexpr int A_BLK = TM * 128;        // TM batch rows x 128 B
  constexpr int W_BLK = 16 * 128;        // 16 unit rows x 128 B
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.S;
  const int KB = 4 * S / 64;              // k-blocks of the 4S-long reduction
  uint8_t* Asm = smem;                    // NST stages
  uint8_t* Wsm = smem + NST * A_BLK;      // KB blocks of 2 KB
  float* as = reinterpret_cast<float*>(Wsm + KB * W_BLK);                 // [128][XS_P] activations -> dG
  float* dhs = as + TM * XS_P;                                           // [128][HS_P]
  float* cs = dhs + TM * HS_P;                                           // [128][HS_P] c(t)
  float* cps = cs + TM * HS_P;                                           // [128][HS_P] c(t_prev)
  __nv_bfloat16* gbs = reinterpret_cast<__nv_bfloat16*>(cps + TM * HS_P);    // [128][GB_P]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbs + TM * GB_P);
  uint64_t* w_full = bars;
  uint64_t* mma_done = bars + 1;
  uint64_t* tmem_free = bars + 2;
  uint64_t* full = bars + 3;
  uint64_t* empty = full + NST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + NST);

  const int slice = blockIdx.x, dir = blockIdx.y, z = blockIdx.z, Z = gridDim.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1);
    mbar_init(mma_done, 1);
    mbar_init(tmem_free, 128);
    for (int i = 0; i < NST; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<32>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(w_full, KB * W_BLK);
    for (int kb = 0; kb < KB; ++kb) tma_load_2d(&tmW, w_full, Wsm + kb * W_BLK, kb * 64, dir * S + slice * RT_UNITS);
  }
  unsigned* gbar = p.bar + dir * Z + z;
  const unsigned G = gridDim.x;
  uint32_t it = 0;                         // tile iterations with a matmul (phase of mma_done)
  uint32_t kcount = 0;                     // k-blocks streamed so far (ring position / phase), same in producer and MMA
  bool w_ready = false;
  const int eg = warp & 3;
  const int te = threadIdx.x - 64;
  const bool single = p.n_tiles <= Z;
  constexpr uint32_t idesc = umma_idesc_bf16(TM, 16);
  const size_t hoff = (size_t)dir * S + slice * RT_UNITS;
  float dcreg[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) dcreg[j] = 0.f;
  float bsum[64];            // per-thread running column sums of dG (bias gradient), reduced across rows at the end
#pragma unroll
  for (int j = 0; j < 64; ++j) bsum[j] = 0.f;

  // coalesced prefetch of everything step (s_) of tile bt_ needs that does not depend on the recurrence
  auto prefetch_in = [&](int s_, int bt_) {
    const int t_ = dir == 0 ? p.n_seq - 1 - s_ : s_;
    const int tp_ = dir == 0 ? t_ - 1 : t_ + 1;
    const bool hp = dir == 0 ? (t_ > 0) : (t_ < p.n_seq - 1);
#pragma unroll 4
    for (int i = 0; i < TM / 8; ++i) {
      const int idx = i * 128 + te, r = idx >> 4, c4 = idx & 15, n = bt_ * TM + r;
      float* dst = as + r * XS_P + c4 * 4;
      if (n < p.n_batch)
        cp_async16(dst, p.xp + ((size_t)t_ * p.rs_seq + (size_t)n * p.rs_batch) * 8 * S + (size_t)dir * 4 * S + slice * 64 + c4 * 4);
      else
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < TM / 32; ++i) {
      const int idx = i * 128 + te, r = idx >> 2, c4 = idx & 3, n = bt_ * TM + r;
      if (n < p.n_batch) {
        const size_t o = ((size_t)t_ * p.rs_seq + (size_t)n * p.rs_batch) * 2 * S + hoff + c4 * 4;
        cp_async16(dhs + r * HS_P + c4 * 4, p.dhout + o);
        cp_async16(cs + r * HS_P + c4 * 4, p.cbuf + o);
        if (hp) cp_async16(cps + r * HS_P + c4 * 4, p.cbuf + ((size_t)tp_ * p.rs_seq + (size_t)n * p.rs_batch) * 2 * S + hoff + c4 * 4);
        else *reinterpret_cast<float4*>(cps + r * HS_P + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        *reinterpret_cast<float4*>(dhs + r * HS_P + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(cs + r * HS_P + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(cps + r * HS_P + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    cp_async_commit();
  };
  if (warp >= 2) prefetch_in(0, z);

  for (int s = 0; s < p.n_seq; ++s) {
    const int t = dir == 0 ? p.n_seq - 1 - s : s;
    const int tn = dir == 0 ? t + 1 : t - 1;
    const int tp = dir == 0 ? t - 1 : t + 1;
    const bool has_prev = dir == 0 ? (t > 0) : (t < p.n_seq - 1);
    for (int bt = z; bt < p.n_tiles; bt += Z) {
      if (s > 0) {
        if (warp == 0) {
          if (elect_one()) {
            if (bt == z) step_wait(gbar, (unsigned)s * G);
            DBG_STAMP(0);
            uint32_t kc = kcount;
            for (int kb = 0; kb < KB; ++kb, ++kc) {
              const int st = kc % NST;
              mbar_wait(empty + st, ((kc / NST) & 1) ^ 1);
              mbar_expect_tx(full + st, A_BLK);
              tma_load_3d(&tmG, full + st, Asm + st * A_BLK, dir * 4 * S + kb * 64, p.seq_inner ? tn : bt * TM,
                          p.seq_inner ? bt * TM : tn);
            }
          }
        } else if (warp == 1) {
          if (elect_one()) {
            if (!w_ready) { mbar_wait(w_full, 0); w_ready = true; }
            uint32_t kc = kcount;
            const uint32_t a0 = smem_u32(Asm), w0 = smem_u32(Wsm);
            for (int kb = 0; kb < KB; ++kb, ++kc) {
              const int st = kc % NST;
              mbar_wait(full + st, (kc / NST) & 1);
              if (kb == 0) {
                if (it > 0) mbar_wait(tmem_free, (it - 1) & 1);
                DBG_STAMP(1);
              }
              tc_fence_after();
              const uint64_t da = umma_desc_k128(a0 + st * A_BLK), db = umma_desc_k128(w0 + kb * W_BLK);
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_bf16_ss(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              mma_commit(empty + st);
            }
            mma_commit(mma_done);
            DBG_STAMP(2);
          }
        }
        kcount += KB;
      }
      if (warp >= 2) {
        const int r = (TM == 128) ? eg * 32 + lane : eg * 16 + (lane & 15);
        const int n = bt * TM + r;
        const bool in_range = (TM == 128 || lane < 16) && n < p.n_batch;
        const int nn = in_range ? n : 0;
        const bool valid = in_range && (p.lens ? (t < p.lens[nn]) : true);
        const bool pv = valid && has_prev && (p.lens ? (tp < p.lens[nn]) : true);
        float* dcs = p.dcstate + (size_t)nn * 2 * S + hoff;
        cp_async_wait_all();
        epi_bar();                                // as / dhs / cs / cps hold the tiles of (t, bt)
        float dcr[16];
        if (single) {
#pragma unroll
          for (int j = 0; j < 16; ++j) dcr[j] = dcreg[j];
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 d4 = (valid && s > 0) ? *(reinterpret_cast<const float4*>(dcs) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            dcr[q * 4 + 0] = d4.x; dcr[q * 4 + 1] = d4.y; dcr[q * 4 + 2] = d4.z; dcr[q * 4 + 3] = d4.w;
          }
        }
        uint32_t v[16];
        if (s > 0) {
          mbar_wait(mma_done, it & 1);
          if (warp == 2 && lane == 0) DBG_STAMP(3);
          tc_fence_after();
          tmem_ld16(tmem + ((uint32_t)(eg * 32) << 16), v);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(tmem_free);
          if (warp == 2 && lane == 0) DBG_STAMP(4);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        float* arow = as + r * XS_P;
        if (TM == 128 || lane < 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 dh4 = *reinterpret_cast<const float4*>(dhs + r * HS_P + q * 4);
          const float4 c4 = *reinterpret_cast<const float4*>(cs + r * HS_P + q * 4);
          const float4 cp4 = *reinterpret_cast<const float4*>(cps + r * HS_P + q * 4);
          const float dhv[4] = {dh4.x, dh4.y, dh4.z, dh4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
          const float cpv[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = q * 4 + u;
            float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
            float dco = 0.f;
            if (valid) {
              const float4 a = *reinterpret_cast<const float4*>(arow + j * 4);
              const float dh = dhv[u] + __uint_as_float(v[j]);
              const float tc_ = tanh_apx(cv[u]);
              const float dc = fmaf(dh * a.w, 1.f - tc_ * tc_, dcr[j]);
              dg.w = dh * tc_ * a.w * (1.f - a.w);
              dg.x = dc * a.z * a.x * (1.f - a.x);
              dg.z = dc * a.x * (1.f - a.z * a.z);
              dg.y = pv ? dc * cpv[u] * a.y * (1.f - a.y) : 0.f;
              dco = dc * a.y;
            }
            dcreg[j] = dco;
            bsum[j * 4 + 0] += dg.x; bsum[j * 4 + 1] += dg.y; bsum[j * 4 + 2] += dg.z; bsum[j * 4 + 3] += dg.w;
            *reinterpret_cast<float4*>(arow + j * 4) = dg;
            __nv_bfloat162 b01 = __floats2bfloat162_rn(dg.x, dg.y), b23 = __floats2bfloat162_rn(dg.z, dg.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&b01);
            pk.y = *reinterpret_cast<uint32_t*>(&b23);
            *reinterpret_cast<uint2*>(gbs + r * GB_P + j * 4) = pk;
          }
        }
        }
        if (!single && in_range) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *(reinterpret_cast<float4*>(dcs) + q) = make_float4(dcreg[q * 4], dcreg[q * 4 + 1], dcreg[q * 4 + 2], dcreg[q * 4 + 3]);
        }
        if (warp == 2 && lane == 0) DBG_STAMP(5);
        epi_bar();
        {
          const size_t rbase = (size_t)t * p.rs_seq;
#pragma unroll
          for (int i = 0; i < TM / 16; ++i) {    // bf16 exchange tile first: TM rows x 128 B
            const int idx = i * 128 + te, rr = idx >> 3, c8 = idx & 7, nb = bt * TM + rr;
            if (nb < p.n_batch)
              *reinterpret_cast<uint4*>(p.xb + (rbase + (size_t)nb * p.rs_batch) * 8 * S + (size_t)dir * 4 * S + slice * 64 + c8 * 8) =
                  *reinterpret_cast<const uint4*>(gbs + rr * GB_P + c8 * 8);
          }
          if (bt + Z >= p.n_tiles && s + 1 < p.n_seq) {   // last tile of this step: publish
            epi_bar();
            if (te == 0) {
              step_arrive(gbar);
              DBG_STAMP(6);
            }
          }
#pragma unroll 4
          for (int i = 0; i < TM / 8; ++i) {
            const int idx = i * 128 + te, rr = idx >> 4, c4 = idx & 15, nb = bt * TM + rr;
            if (nb < p.n_batch)
              __stcs(reinterpret_cast<float4*>(p.xp + (rbase + (size_t)nb * p.rs_batch) * 8 * S + (size_t)dir * 4 * S + slice * 64 + c4 * 4),
                     *reinterpret_cast<const float4*>(as + rr * XS_P + c4 * 4));
          }
        }
        epi_bar();
        {
          int bn = bt + Z, sn = s;
          if (bn >= p.n_tiles) { bn = z; sn = s + 1; }
          if (sn < p.n_seq) prefetch_in(sn, bn);
        }
        if (warp == 2 && lane == 0) DBG_STAMP(7);
      }
      if (s > 0) ++it;
    }
  }
  cp_async_wait_all();
  if (p.dbias) {                      // bias gradient: reduce the per-thread partial sums across the tile rows
    __syncthreads();
    if (warp >= 2 && (TM == 128 || lane < 16)) {
      float* brow = as + ((TM == 128) ? eg * 32 + lane : eg * 16 + lane) * XS_P;
#pragma unroll
      for (int j = 0; j < 16; ++j) *reinterpret_cast<float4*>(brow + j * 4) = make_float4(bsum[j * 4], bsum[j * 4 + 1], bsum[j * 4 + 2], bsum[j * 4 + 3]);
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      float acc = 0.f;
      for (int rr = 0; rr < TM; ++rr) acc += as[rr * XS_P + threadIdx.x];
      atomicAdd(p.dbias + (size_t)dir * 4 * S + slice * 64 + threadIdx.x, acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<32>(tmem);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int max_z(int S) {
  int z = sm_count() / (2 * (S / RT_UNITS));
  return z < 1 ? 1 : z;
}
// 64-row tiles whenever that still gives every tile its own CTA: twice the CTAs, half the per-SM exchange traffic
static int pick_tm(int S, int n_batch) { return (S > 256 || (n_batch + 63) / 64 <= max_z(S)) ? 64 : 128; }   // S > 256: only 64-row tiles fit in smem
static int pick_z(int S, int n_tiles) {
  int z = max_z(S);
  return z > n_tiles ? n_tiles : z;
}

int rec_tc_supported(int S) { return (S % 64 == 0 && S >= 64 && S <= 512) ? 1 : 0; }

template <int TM, int KB, bool X3>
static int launch_fwd_tc_x(const CUtensorMap& tmH, const CUtensorMap& tmW, const CUtensorMap& tmH2, const CUtensorMap& tmW2,
                           RecTcParams& p, dim3 grid, cudaStream_t st) {
  constexpr int NP = X3 ? 2 : 1;
  const size_t smem = (size_t)NP * KB * (64 * 128 + TM * 128) + (size_t)TM * (XS_P + 2 * HS_P) * 4 + (size_t)NP * TM * HB_P * 2 + 64 + 1024;
  SSASR_REQUIRE(smem <= 227 * 1024, "rec_tc_fwd: %zu B shared memory needed (S=%d)", smem, p.S);
  SSASR_CHECK_CUDA(cudaFuncSetAttribute(rec_tc_fwd_kernel<TM, KB, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&tmH, (void*)&tmW, (void*)&tmH2, (void*)&tmW2, (void*)&p};
  ProfScope ps(F_REC_TC_FWD, st);
  SSASR_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)rec_tc_fwd_kernel<TM, KB, X3>, grid, dim3(RT_THREADS), args, smem, st));
  return 0;
}
template <int TM, int KB>
static int launch_fwd_tc(const CUtensorMap& tmH, const CUtensorMap& tmW, RecTcParams& p, dim3 grid, cudaStream_t st) {
  return launch_fwd_tc_x<TM, KB, false>(tmH, tmW, tmH, tmW, p, grid, st);
}
template <int TM>
static int launch_fwd_tc_kb(int kb, const CUtensorMap& tmH, const CUtensorMap& tmW, RecTcParams& p, dim3 grid, cudaStream_t st) {
  switch (kb) {
    case 1: return launch_fwd_tc<TM, 1>(tmH, tmW, p, grid, st);
    case 2: return launch_fwd_tc<TM, 2>(tmH, tmW, p, grid, st);
    case 3: return launch_fwd_tc<TM, 3>(tmH, tmW, p, grid, st);
    case 4: return launch_fwd_tc<TM, 4>(tmH, tmW, p, grid, st);
  }
  if constexpr (TM == 64) {
    switch (kb) {
      case 5: return launch_fwd_tc<TM, 5>(tmH, tmW, p, grid, st);
      case 6: return launch_fwd_tc<TM, 6>(tmH, tmW, p, grid, st);
      case 7: return launch_fwd_tc<TM, 7>(tmH, tmW, p, grid, st);
      case 8: return launch_fwd_tc<TM, 8>(tmH, tmW, p, grid, st);
    }
  }
  return -1;
}

// hb: bf16 [rows,2S] h exchange buffer; whh_bf: bf16 [2*4S, S] (interleaved rows, both directions)
int rec_tc_fwd(cudaStream_t st, float* xp, const void* whh_bf, float* hout, float* cbuf, void* hb, const int* lens, int S,
               int n_seq, int n_batch, long long rs_seq, long long rs_batch, unsigned* bar) {
  SSASR_REQUIRE(rec_tc_supported(S), "rec_tc_fwd: unsupported state size %d", S);
  if (rec_cl_supported(S, n_batch, 0)) return rec_cl_fwd(st, xp, whh_bf, hout, cbuf, hb, lens, S, n_seq, n_batch, rs_seq, rs_batch);
  if (rec_wide_supported(S, n_batch, 0)) return rec_wide_fwd(st, xp, whh_bf, hout, cbuf, hb, lens, S, n_seq, n_batch, rs_seq, rs_batch);
  const int TM = pick_tm(S, n_batch);
  RecTcParams p;
  p.xp = xp; p.hout = hout; p.cbuf = cbuf; p.xb = (__nv_bfloat16*)hb; p.xb2 = nullptr; p.dhout = nullptr; p.dcstate = nullptr; p.dbias = nullptr;
  p.lens = lens;
  p.S = S; p.n_seq = n_seq; p.n_batch = n_batch; p.n_tiles = (n_batch + TM - 1) / TM;
  p.rs_seq = rs_seq; p.rs_batch = rs_batch; p.bar = bar; p.dbg = g_dbg;
  const int Z = pick_z(S, p.n_tiles);
  SSASR_CHECK_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned) * 2 * Z, st));
  CUtensorMap tmH, tmW;
  p.seq_inner = rs_seq < rs_batch ? 1 : 0;
  int rc = p.seq_inner ? make_tmap_bf16_3d(&tmH, hb, 2 * S, n_seq, rs_seq * 2 * S, n_batch, rs_batch * 2 * S, 1, TM)
                       : make_tmap_bf16_3d(&tmH, hb, 2 * S, n_batch, rs_batch * 2 * S, n_seq, rs_seq * 2 * S, TM, 1);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW, whh_bf, 8 * S, S, S, 64);
  if (rc) return rc;
  dim3 grid(S / RT_UNITS, 2, Z);
  return TM == 64 ? launch_fwd_tc_kb<64>(S / 64, tmH, tmW, p, grid, st) : launch_fwd_tc_kb<128>(S / 64, tmH, tmW, p, grid, st);
}

// fp32-accurate forward (exact path): hb_hi/hb_lo exchange buffers, whh_hi/whh_lo bf16 [8S,S]; 64-row tiles, S <= 256
int rec_tc_fwd_x3(cudaStream_t st, float* xp, const void* whh_hi, const void* whh_lo, float* hout, float* cbuf, void* hb_hi,
                  void* hb_lo, const int* lens, int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch, unsigned* bar) {
  SSASR_REQUIRE(S % 64 == 0 && S >= 64 && S <= 256, "rec_tc_fwd_x3: unsupported state size %d", S);
  constexpr int TM = 64;
  RecTcParams p;
  p.xp = xp; p.hout = hout; p.cbuf = cbuf; p.xb = (__nv_bfloat16*)hb_hi; p.xb2 = (__nv_bfloat16*)hb_lo; p.dhout = nullptr;
  p.dcstate = nullptr; p.dbias = nullptr; p.lens = lens;
  p.S = S; p.n_seq = n_seq; p.n_batch = n_batch; p.n_tiles = (n_batch + TM - 1) / TM;
  p.rs_seq = rs_seq; p.rs_batch = rs_batch; p.bar = bar; p.dbg = nullptr;
  const int Z = pick_z(S, p.n_tiles);
  SSASR_CHECK_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned) * 2 * Z, st));
  CUtensorMap tmH, tmH2, tmW, tmW2;
  p.seq_inner = rs_seq < rs_batch ? 1 : 0;
  int rc = 0;
  for (int i = 0; i < 2 && !rc; ++i) {
    void* hb = i ? hb_lo : hb_hi;
    CUtensorMap* tm = i ? &tmH2 : &tmH;
    rc = p.seq_inner ? make_tmap_bf16_3d(tm, hb, 2 * S, n_seq, rs_seq * 2 * S, n_batch, rs_batch * 2 * S, 1, TM)
                     : make_tmap_bf16_3d(tm, hb, 2 * S, n_batch, rs_batch * 2 * S, n_seq, rs_seq * 2 * S, TM, 1);
  }
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW, whh_hi, 8 * S, S, S, 64);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW2, whh_lo, 8 * S, S, S, 64);
  if (rc) return rc;
  dim3 grid(S / RT_UNITS, 2, Z);
  switch (S / 64) {
    case 1: return launch_fwd_tc_x<TM, 1, true>(tmH, tmW, tmH2, tmW2, p, grid, st);
    case 2: return launch_fwd_tc_x<TM, 2, true>(tmH, tmW, tmH2, tmW2, p, grid, st);
    case 3: return launch_fwd_tc_x<TM, 3, true>(tmH, tmW, tmH2, tmW2, p, grid, st);
    case 4: return launch_fwd_tc_x<TM, 4, true>(tmH, tmW, tmH2, tmW2, p, grid, st);
  }
  return -1;
}

template <int TM, int NST>
static int launch_bwd_tc(const CUtensorMap& tmG, const CUtensorMap& tmW, RecTcParams& p, dim3 grid, cudaStream_t st) {
  const int S = p.S;
  const size_t smem = (size_t)NST * TM * 128 + (size_t)(4 * S / 64) * 16 * 128 + (size_t)TM * (XS_P + 3 * HS_P) * 4 +
                      (size_t)TM * GB_P * 2 + (3 + 2 * NST) * 8 + 16 + 1024;
  SSASR_REQUIRE(smem <= 227 * 1024, "rec_tc_bwd: %zu B shared memory needed (S=%d)", smem, S);
  SSASR_CHECK_CUDA(cudaFuncSetAttribute(rec_tc_bwd_kernel<TM, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&tmG, (void*)&tmW, (void*)&p};
  ProfScope ps(F_REC_TC_BWD, st);
  SSASR_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)rec_tc_bwd_kernel<TM, NST>, grid, dim3(RT_THREADS), args, smem, st));
  return 0;
}

// dgb: bf16 [rows,8S] dG exchange buffer (on return: the complete bf16 copy of dG); whhT_bf: bf16 [2*S, 4S]
int rec_tc_bwd(cudaStream_t st, float* act, const void* whhT_bf, const float* cbuf, const float* dhout, void* dgb, float* dcstate,
               const int* lens, int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch, unsigned* bar, float* dbias,
               int need_dg32) {
  SSASR_REQUIRE(rec_tc_supported(S), "rec_tc_bwd: unsupported state size %d", S);
  if (!need_dg32 && rec_ks_supported(S, n_batch))        // K-split cluster kernel (rec_wide.cu): bf16 dG only
    return rec_wide_bwd(st, act, whhT_bf, cbuf, dhout, dgb, lens, S, n_seq, n_batch, rs_seq, rs_batch, dbias);
  if (!need_dg32 && rec_cl_supported(S, n_batch, 1))     // the cluster kernel only produces the bf16 dG
    return rec_cl_bwd(st, act, whhT_bf, cbuf, dhout, dgb, lens, S, n_seq, n_batch, rs_seq, rs_batch, dbias);
  if (!need_dg32 && rec_wide_supported(S, n_batch, 1))
    return rec_wide_bwd(st, act, whhT_bf, cbuf, dhout, dgb, lens, S, n_seq, n_batch, rs_seq, rs_batch, dbias);
  const int TM = pick_tm(S, n_batch);
  RecTcParams p;
  p.xp = act; p.hout = nullptr; p.cbuf = const_cast<float*>(cbuf); p.xb = (__nv_bfloat16*)dgb; p.xb2 = nullptr; p.dhout = dhout; p.dcstate = dcstate;
  p.dbias = dbias;
  p.lens = lens; p.S = S; p.n_seq = n_seq; p.n_batch = n_batch; p.n_tiles = (n_batch + TM - 1) / TM;
  p.rs_seq = rs_seq; p.rs_batch = rs_batch; p.bar = bar; p.dbg = g_dbg;
  const int Z = pick_z(S, p.n_tiles);
  SSASR_CHECK_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned) * 2 * Z, st));
  CUtensorMap tmG, tmW;
  p.seq_inner = rs_seq < rs_batch ? 1 : 0;
  int rc = p.seq_inner ? make_tmap_bf16_3d(&tmG, dgb, 8 * S, n_seq, rs_seq * 8 * S, n_batch, rs_batch * 8 * S, 1, TM)
                       : make_tmap_bf16_3d(&tmG, dgb, 8 * S, n_batch, rs_batch * 8 * S, n_seq, rs_seq * 8 * S, TM, 1);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW, whhT_bf, 2 * S, 4 * S, 4 * S, 16);
  if (rc) return rc;
  dim3 grid(S / RT_UNITS, 2, Z);
  return TM == 64 ? launch_bwd_tc<64, 12>(tmG, tmW, p, grid, st) : launch_bwd_tc<128, 6>(tmG, tmW, p, grid, st);
}

}  // namespace ssasr

extern "C" {
// debug: device buffer [n_seq][12] of clock64 stamps written by CTA (0,0,0) of the next tensor-core recurrent launches
void ssasr_rec_tc_set_debug(long long* dev_buf) { ssasr::g_dbg = dev_buf; }
}

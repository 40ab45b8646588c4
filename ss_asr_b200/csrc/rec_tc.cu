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

__device__ __forceinline__ void group_barrier(unsigned* counter, unsigned target) {
  fence_proxy_async_all();          // generic-proxy global stores -> visible to other CTAs' TMA (async proxy) reads
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (++spins > (1u << 26)) __trap();
    } while (v < target);
    fence_proxy_async_all();
  }
  __syncthreads();
}

constexpr int RT_UNITS = 16;           // hidden units per CTA slice
constexpr int RT_THREADS = 192;        // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue

struct RecTcParams {
  float* xp;                 // fwd: [rows,8S] pre-activations in / activations out.  bwd: activations in / dG out
  float* hout;               // fwd out [rows,2S] fp32
  float* cbuf;               // fwd out / bwd in
  __nv_bfloat16* xb;         // fwd: h exchange buffer [rows,2S] bf16.  bwd: dG exchange buffer [rows,8S] bf16
  const float* dhout;        // bwd
  float* dcstate;            // bwd [n_batch,2S]
  const int* lens;
  int S, n_seq, n_batch, n_tiles;
  long long rs_seq, rs_batch;
  int seq_inner;             // 1: tensor-map dim1 = seq, dim2 = batch (time-major layers); 0: dim1 = batch, dim2 = seq
  unsigned* bar;             // [2 * gridDim.z]
};

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int KB>   // KB = S / 64 resident k-blocks
__global__ void __launch_bounds__(RT_THREADS, 1)
rec_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW, RecTcParams p) {
  constexpr int W_BLK = 64 * 128;        // 64 gate rows x 128 B
  constexpr int A_BLK = 128 * 128;       // 128 batch rows x 128 B
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Wsm = smem;
  uint8_t* Asm = smem + KB * W_BLK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Asm + KB * A_BLK);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* mma_done = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int S = p.S;
  const int slice = blockIdx.x, dir = blockIdx.y, z = blockIdx.z, Z = gridDim.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1);
    mbar_init(a_full, 1);
    mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<64>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(w_full, KB * W_BLK);
    for (int kb = 0; kb < KB; ++kb) tma_load_2d(&tmW, w_full, Wsm + kb * W_BLK, kb * 64, dir * 4 * S + slice * 64);
  }
  unsigned* gbar = p.bar + dir * Z + z;
  const unsigned G = gridDim.x;
  uint32_t it = 0;                         // completed (tile, s>0) iterations: phase of a_full / mma_done
  bool w_ready = false;
  const int eg = warp & 3;                 // TMEM lane group of an epilogue warp
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64);

  for (int s = 0; s < p.n_seq; ++s) {
    const int t = dir == 0 ? s : p.n_seq - 1 - s;
    const int tp = dir == 0 ? t - 1 : t + 1;
    for (int bt = z; bt < p.n_tiles; bt += Z) {
      if (s > 0) {
        if (warp == 0) {
          if (elect_one()) {
            mbar_expect_tx(a_full, KB * A_BLK);
            for (int kb = 0; kb < KB; ++kb)
              tma_load_3d(&tmH, a_full, Asm + kb * A_BLK, dir * S + kb * 64, p.seq_inner ? tp : bt * 128, p.seq_inner ? bt * 128 : tp);
          }
        } else if (warp == 1) {
          if (elect_one()) {
            if (!w_ready) { mbar_wait(w_full, 0); w_ready = true; }
            mbar_wait(a_full, it & 1);
            tc_fence_after();
            const uint32_t a0 = smem_u32(Asm), w0 = smem_u32(Wsm);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
              const uint64_t da = umma_desc_k128(a0 + kb * A_BLK), db = umma_desc_k128(w0 + kb * W_BLK);
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_bf16_ss(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            }
            mma_commit(mma_done);
          }
        }
      }
      if (warp >= 2) {
        const int n = bt * 128 + eg * 32 + lane;
        if (s > 0) {
          mbar_wait(mma_done, it & 1);
          tc_fence_after();
        }
        const bool in_range = n < p.n_batch;
        const bool valid = in_range && (p.lens ? (t < p.lens[in_range ? n : 0]) : true);
        const size_t row = (size_t)t * p.rs_seq + (size_t)(in_range ? n : 0) * p.rs_batch;
        const size_t rowp = (size_t)(s > 0 ? tp : t) * p.rs_seq + (size_t)(in_range ? n : 0) * p.rs_batch;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {     // 4 units (16 gate columns) per chunk
          uint32_t v[16];
          if (s > 0) {
            tmem_ld16(tmem + ((uint32_t)(eg * 32) << 16) + (uint32_t)(ch * 16), v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
          }
          if (in_range) {
            float* xptr = p.xp + row * 8 * S + (size_t)dir * 4 * S + slice * 64 + ch * 16;
            const size_t hoff = (size_t)dir * S + slice * RT_UNITS + ch * 4;
            float4 cp4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s > 0 && valid) cp4 = *reinterpret_cast<const float4*>(p.cbuf + rowp * 2 * S + hoff);
            const float cpv[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
            float hv[4], cv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float4 g = *reinterpret_cast<const float4*>(xptr + u * 4);
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
              hv[u] = 0.f; cv[u] = 0.f;
              if (valid) {
                g.x += __uint_as_float(v[u * 4 + 0]); g.y += __uint_as_float(v[u * 4 + 1]);
                g.z += __uint_as_float(v[u * 4 + 2]); g.w += __uint_as_float(v[u * 4 + 3]);
                a.x = sigmoid_fast(g.x); a.y = sigmoid_fast(g.y); a.z = tanh_fast(g.z); a.w = sigmoid_fast(g.w);
                cv[u] = a.y * cpv[u] + a.x * a.z;
                hv[u] = a.w * tanh_fast(cv[u]);
              }
              *reinterpret_cast<float4*>(xptr + u * 4) = a;
            }
            *reinterpret_cast<float4*>(p.hout + row * 2 * S + hoff) = make_float4(hv[0], hv[1], hv[2], hv[3]);
            *reinterpret_cast<float4*>(p.cbuf + row * 2 * S + hoff) = make_float4(cv[0], cv[1], cv[2], cv[3]);
            __nv_bfloat162 b01 = __floats2bfloat162_rn(hv[0], hv[1]), b23 = __floats2bfloat162_rn(hv[2], hv[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&b01);
            pk.y = *reinterpret_cast<uint32_t*>(&b23);
            *reinterpret_cast<uint2*>(p.xb + row * 2 * S + hoff) = pk;
          }
        }
        tc_fence_before();
      }
      if (s > 0) ++it;
      __syncthreads();                      // A tile and TMEM accumulator are free again
      tc_fence_after();
    }
    if (s + 1 < p.n_seq) group_barrier(gbar, (unsigned)(s + 1) * G);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(tmem);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int NST>   // ring stages of 16 KB
__global__ void __launch_bounds__(RT_THREADS, 1)
rec_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmW, RecTcParams p) {
  constexpr int A_BLK = 128 * 128;       // 128 batch rows x 128 B
  constexpr int W_BLK = 16 * 128;        // 16 unit rows x 128 B
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.S;
  const int KB = 4 * S / 64;              // k-blocks of the 4S-long reduction
  uint8_t* Asm = smem;                    // NST stages
  uint8_t* Wsm = smem + NST * A_BLK;      // KB blocks of 2 KB, padded to 1024-B alignment each -> use 2 KB pitch
  uint64_t* bars = reinterpret_cast<uint64_t*>(Wsm + KB * W_BLK);
  uint64_t* w_full = bars;
  uint64_t* mma_done = bars + 1;
  uint64_t* full = bars + 2;
  uint64_t* empty = full + NST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + NST);

  const int slice = blockIdx.x, dir = blockIdx.y, z = blockIdx.z, Z = gridDim.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1);
    mbar_init(mma_done, 1);
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
  constexpr uint32_t idesc = umma_idesc_bf16(128, 16);

  for (int s = 0; s < p.n_seq; ++s) {
    const int t = dir == 0 ? p.n_seq - 1 - s : s;
    const int tn = dir == 0 ? t + 1 : t - 1;
    const int tp = dir == 0 ? t - 1 : t + 1;
    const bool has_prev = dir == 0 ? (t > 0) : (t < p.n_seq - 1);
    for (int bt = z; bt < p.n_tiles; bt += Z) {
      if (s > 0) {
        if (warp == 0) {
          if (elect_one()) {
            uint32_t kc = kcount;
            for (int kb = 0; kb < KB; ++kb, ++kc) {
              const int st = kc % NST;
              mbar_wait(empty + st, ((kc / NST) & 1) ^ 1);
              mbar_expect_tx(full + st, A_BLK);
              tma_load_3d(&tmG, full + st, Asm + st * A_BLK, dir * 4 * S + kb * 64, p.seq_inner ? tn : bt * 128,
                          p.seq_inner ? bt * 128 : tn);
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
              tc_fence_after();
              const uint64_t da = umma_desc_k128(a0 + st * A_BLK), db = umma_desc_k128(w0 + kb * W_BLK);
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_bf16_ss(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              mma_commit(empty + st);
            }
            mma_commit(mma_done);
          }
        }
        kcount += KB;
      }
      if (warp >= 2) {
        const int n = bt * 128 + eg * 32 + lane;
        uint32_t v[16];
        if (s > 0) {
          mbar_wait(mma_done, it & 1);
          tc_fence_after();
          tmem_ld16(tmem + ((uint32_t)(eg * 32) << 16), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        tc_fence_before();
        if (n < p.n_batch) {
          const bool valid = p.lens ? (t < p.lens[n]) : true;
          const size_t row = (size_t)t * p.rs_seq + (size_t)n * p.rs_batch;
          const size_t rowp = (size_t)(has_prev ? tp : t) * p.rs_seq + (size_t)n * p.rs_batch;
          const bool pv = has_prev && (p.lens ? (tp < p.lens[n]) : true);
          const size_t hoff = (size_t)dir * S + slice * RT_UNITS;
          float* aptr = p.xp + row * 8 * S + (size_t)dir * 4 * S + slice * 64;
          __nv_bfloat16* gb = p.xb + row * 8 * S + (size_t)dir * 4 * S + slice * 64;
          float* dcs = p.dcstate + (size_t)n * 2 * S + hoff;
#pragma unroll
          for (int q = 0; q < 4; ++q) {      // 4 units per group
            float4 dh4 = make_float4(0.f, 0.f, 0.f, 0.f), c4 = dh4, cp4 = dh4, dcr4 = dh4;
            if (valid) {
              dh4 = *reinterpret_cast<const float4*>(p.dhout + row * 2 * S + hoff + q * 4);
              c4 = *reinterpret_cast<const float4*>(p.cbuf + row * 2 * S + hoff + q * 4);
              if (pv) cp4 = *reinterpret_cast<const float4*>(p.cbuf + rowp * 2 * S + hoff + q * 4);
              if (s > 0) dcr4 = *reinterpret_cast<const float4*>(dcs + q * 4);
            }
            const float dhv[4] = {dh4.x, dh4.y, dh4.z, dh4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
            const float cpv[4] = {cp4.x, cp4.y, cp4.z, cp4.w}, dcrv[4] = {dcr4.x, dcr4.y, dcr4.z, dcr4.w};
            float dco[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
              dco[u] = 0.f;
              if (valid) {
                const float4 a = *reinterpret_cast<const float4*>(aptr + (q * 4 + u) * 4);
                const float dh = dhv[u] + __uint_as_float(v[q * 4 + u]);
                const float tc_ = tanh_fast(cv[u]);
                const float dc = dh * a.w * (1.f - tc_ * tc_) + dcrv[u];
                dg.w = dh * tc_ * a.w * (1.f - a.w);
                dg.x = dc * a.z * a.x * (1.f - a.x);
                dg.z = dc * a.x * (1.f - a.z * a.z);
                dg.y = dc * cpv[u] * a.y * (1.f - a.y);
                dco[u] = dc * a.y;
              }
              *reinterpret_cast<float4*>(aptr + (q * 4 + u) * 4) = dg;
              __nv_bfloat162 b01 = __floats2bfloat162_rn(dg.x, dg.y), b23 = __floats2bfloat162_rn(dg.z, dg.w);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&b01);
              pk.y = *reinterpret_cast<uint32_t*>(&b23);
              *reinterpret_cast<uint2*>(gb + (q * 4 + u) * 4) = pk;
            }
            *reinterpret_cast<float4*>(dcs + q * 4) = make_float4(dco[0], dco[1], dco[2], dco[3]);
          }
        }
      }
      if (s > 0) ++it;
      __syncthreads();
      tc_fence_after();
    }
    if (s + 1 < p.n_seq) group_barrier(gbar, (unsigned)(s + 1) * G);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<32>(tmem);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int pick_z(int S, int n_tiles) {
  int slices2 = 2 * (S / RT_UNITS);
  int z = sm_count() / slices2;
  if (z < 1) z = 1;
  if (z > n_tiles) z = n_tiles;
  return z;
}

// scratch words needed for the step barriers of one launch
int rec_tc_bar_words(int S, int n_batch) { return 2 * pick_z(S, (n_batch + 127) / 128); }

int rec_tc_supported(int S) { return (S % 64 == 0 && S >= 64 && S <= 512) ? 1 : 0; }

template <int KB>
static int launch_fwd_tc(const CUtensorMap& tmH, const CUtensorMap& tmW, RecTcParams& p, dim3 grid, cudaStream_t st) {
  const size_t smem = (size_t)KB * (64 * 128 + 128 * 128) + 64 + 1024;
  SSASR_CHECK_CUDA(cudaFuncSetAttribute(rec_tc_fwd_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&tmH, (void*)&tmW, (void*)&p};
  ProfScope ps(F_REC_TC_FWD, st);
  SSASR_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)rec_tc_fwd_kernel<KB>, grid, dim3(RT_THREADS), args, smem, st));
  return 0;
}

// hb: bf16 [rows,2S] h exchange buffer; whh_bf: bf16 [2*4S, S] (interleaved rows, both directions)
int rec_tc_fwd(cudaStream_t st, float* xp, const void* whh_bf, float* hout, float* cbuf, void* hb, const int* lens, int S,
               int n_seq, int n_batch, long long rs_seq, long long rs_batch, unsigned* bar) {
  SSASR_REQUIRE(rec_tc_supported(S), "rec_tc_fwd: unsupported state size %d", S);
  RecTcParams p;
  p.xp = xp; p.hout = hout; p.cbuf = cbuf; p.xb = (__nv_bfloat16*)hb; p.dhout = nullptr; p.dcstate = nullptr; p.lens = lens;
  p.S = S; p.n_seq = n_seq; p.n_batch = n_batch; p.n_tiles = (n_batch + 127) / 128;
  p.rs_seq = rs_seq; p.rs_batch = rs_batch; p.bar = bar;
  const int Z = pick_z(S, p.n_tiles);
  SSASR_CHECK_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned) * 2 * Z, st));
  CUtensorMap tmH, tmW;
  p.seq_inner = rs_seq < rs_batch ? 1 : 0;
  int rc = p.seq_inner ? make_tmap_bf16_3d(&tmH, hb, 2 * S, n_seq, rs_seq * 2 * S, n_batch, rs_batch * 2 * S, 1, 128)
                       : make_tmap_bf16_3d(&tmH, hb, 2 * S, n_batch, rs_batch * 2 * S, n_seq, rs_seq * 2 * S, 128, 1);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW, whh_bf, 8 * S, S, S, 64);
  if (rc) return rc;
  dim3 grid(S / RT_UNITS, 2, Z);
  switch (S / 64) {
    case 1: return launch_fwd_tc<1>(tmH, tmW, p, grid, st);
    case 2: return launch_fwd_tc<2>(tmH, tmW, p, grid, st);
    case 3: return launch_fwd_tc<3>(tmH, tmW, p, grid, st);
    case 4: return launch_fwd_tc<4>(tmH, tmW, p, grid, st);
    case 5: return launch_fwd_tc<5>(tmH, tmW, p, grid, st);
    case 6: return launch_fwd_tc<6>(tmH, tmW, p, grid, st);
    case 7: return launch_fwd_tc<7>(tmH, tmW, p, grid, st);
    case 8: return launch_fwd_tc<8>(tmH, tmW, p, grid, st);
  }
  return -1;
}

// dgb: bf16 [rows,8S] dG exchange buffer (on return: the complete bf16 copy of dG); whhT_bf: bf16 [2*S, 4S]
int rec_tc_bwd(cudaStream_t st, float* act, const void* whhT_bf, const float* cbuf, const float* dhout, void* dgb, float* dcstate,
               const int* lens, int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch, unsigned* bar) {
  SSASR_REQUIRE(rec_tc_supported(S), "rec_tc_bwd: unsupported state size %d", S);
  constexpr int NST = 6;
  RecTcParams p;
  p.xp = act; p.hout = nullptr; p.cbuf = const_cast<float*>(cbuf); p.xb = (__nv_bfloat16*)dgb; p.dhout = dhout; p.dcstate = dcstate;
  p.lens = lens; p.S = S; p.n_seq = n_seq; p.n_batch = n_batch; p.n_tiles = (n_batch + 127) / 128;
  p.rs_seq = rs_seq; p.rs_batch = rs_batch; p.bar = bar;
  const int Z = pick_z(S, p.n_tiles);
  SSASR_CHECK_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned) * 2 * Z, st));
  CUtensorMap tmG, tmW;
  p.seq_inner = rs_seq < rs_batch ? 1 : 0;
  int rc = p.seq_inner ? make_tmap_bf16_3d(&tmG, dgb, 8 * S, n_seq, rs_seq * 8 * S, n_batch, rs_batch * 8 * S, 1, 128)
                       : make_tmap_bf16_3d(&tmG, dgb, 8 * S, n_batch, rs_batch * 8 * S, n_seq, rs_seq * 8 * S, 128, 1);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmW, whhT_bf, 2 * S, 4 * S, 4 * S, 16);
  if (rc) return rc;
  const size_t smem = (size_t)NST * 128 * 128 + (size_t)(4 * S / 64) * 16 * 128 + (2 + 2 * NST) * 8 + 16 + 1024;
  SSASR_CHECK_CUDA(cudaFuncSetAttribute(rec_tc_bwd_kernel<NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(S / RT_UNITS, 2, Z);
  void* args[] = {(void*)&tmG, (void*)&tmW, (void*)&p};
  ProfScope ps(F_REC_TC_BWD, st);
  SSASR_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)rec_tc_bwd_kernel<NST>, grid, dim3(RT_THREADS), args, smem, st));
  return 0;
}

}  // namespace ssasr

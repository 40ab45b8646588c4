// Cluster recurrent BLSTM kernels (bf16 training path).  The CTAs that share one (direction, 64-row batch tile) form ONE
// thread-block cluster (S/32 CTAs, 32 hidden units each) and exchange the per-step h / dG tile without any grid-wide
// barrier and without a generic-proxy global store on the dependent chain:
//
//   epilogue warps   TMEM accumulator -> cell math in registers -> this CTA's slice of the exchange tile written to a
//                    shared-memory IMAGE (already in the swizzled UMMA operand layout) -> arrive on `stage_ready`
//   exchange thread  bulk-stores the image to a small L2-resident ring slot (cp.async.bulk.global.shared::cta), waits for
//                    that bulk group, then reads the slot back with ONE multicast bulk copy
//                    (cp.async.bulk.shared::cluster.global ... .multicast::cluster) that lands it in the operand buffer of
//                    EVERY CTA of the cluster and completes transaction bytes on every CTA's mbarrier
//   MMA thread       waits on its local mbarrier(s) until all S/32 slices have landed, issues the step's tcgen05.mma
//
// so the producer->consumer dependency is carried entirely by mbarrier transaction counts.  Measured on B200
// (scripts/exch_bench.cu, 8-CTA cluster, 4 KB per CTA): bulk store + multicast read-back 1.0 k cycles per dependent step,
// st.global + fence + multicast 1.6 k, DSMEM bulk copies 2.1 k, st.shared::cluster 5.3 k; the counter barrier + per-CTA TMA
// reload of rec_tc.cu costs ~8 k.  16-CTA clusters (16 units per CTA) would halve the per-CTA work but only 7 of them
// are co-resident on a B200 (one GPC is short), 8 are needed for 256 utterances x 2 directions.
//
// Everything that is NOT on the dependent chain (pre-activations / saved activations / c / dhout in, h / c / activations /
// row-major bf16 copies out) moves with plain 128-bit global loads issued at the top of a step (their latency hides
// behind the exchange + MMA) and plain stores issued after the `stage_ready` arrive.  The step loop has no block barrier.
//
// Row <-> thread mapping: M=64 accumulators occupy the lower 16 lanes of each TMEM sub-partition; lane l < 16 of an
// epilogue warp reads row r's values and hands half of them to lane l + 16 by shuffle, so both half-warps work on the
// same row (8 hidden units each).  Two epilogue groups of 4 warps own the two 16-unit halves of the CTA's 32 units.
//
// Operand layouts.  Forward: A = h(t-1) tile, [64 rows x S] bf16 as S/32 k-blocks of 64 x 64 B with the 64-byte swizzle
// (one k-block = one producer's image), double-buffered; B = W_hh slice, 128 gate rows x S, 128-byte swizzle, resident.
// Backward: A = dG(t+1) tile, [64 x 4S] bf16 as S/16 k-blocks of 64 x 128 B (a producer's image = two k-blocks = its
// 32 units x 4 gates), single-buffered behind the `a_free` cluster barrier; B = W_hh^T slice, 32 unit rows x 4S.
//
// Reference semantics: nn.LSTM(bidirectional) packed (asr.py:410-418) and blstm_4 (asr.py:262); masking and row-stride
// conventions, operand rounding (bf16), accumulation (fp32) and cell math (tanh.approx) as in rec_tc.cu.
#include <limits.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"
#include "cl_common.cuh"

namespace ssasr {

using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows);
int make_tmap_bf16_3d_ex(CUtensorMap* m, const void* ptr, long long cols, long long nA, long long strideA, long long nB,
                         long long strideB, int box_cols, int boxA, int boxB, int swizzle_bytes);

namespace {

constexpr int CL_UNITS = 32;       // hidden units per CTA (two 16-unit halves, one per epilogue warp group)
constexpr int CL_THREADS = 320;    // warp 0: exchange, warp 1: MMA + TMEM alloc, warps 2-5 / 6-9: epilogue groups 0 / 1
constexpr int CL_TM = 64;          // batch rows per tile
constexpr int CL_RING = 4;         // global exchange ring depth (2 would do, see the hazard notes)
constexpr int FW_IMG = CL_TM * 64;        // forward image: 64 rows x 32 bf16
constexpr int BW_IMG = 2 * CL_TM * 128;   // backward image: two k-blocks of 64 rows x 64 bf16
constexpr int BW_MAXNC = 8;
constexpr int FW_MAXKX = 5;         // fused input projection: up to 80 input features
constexpr int FW_WSTG = 4096 + 1024 + 1024;   // forward per-warp staging: gate tile + h tile + c tile

struct RecClParams {
  float* xp;                 // fwd: [rows,8S] pre-activations in / activations out.  bwd: activations in
  float* hout;               // fwd out [rows,2S]
  float* cbuf;               // fwd out / bwd in [rows,2S]
  __nv_bfloat16* xb;         // fwd out: bf16 h [rows,2S].  bwd out: bf16 dG [rows,8S]
  const float* dhout;        // bwd
  float* dbias;              // bwd: [8S] pre-zeroed, atomically accumulated; may be null
  uint8_t* ring;             // [CL_RING][n_cta][image bytes] exchange slots
  const float* bias;         // fwd, fused input projection: packed gate bias [8S]
  int nkx;                   // fwd, fused input projection: 16-column k-blocks of the layer input (ceil(Kp / 16) <= FW_MAXKX)
  int seq_inner;             // fwd, fused input projection: tensor-map dim1 = seq (1) or batch (0)
  const int* lens;
  int S, n_seq, n_batch;
  long long rs_seq, rs_batch;
  long long* dbg;            // optional [n_seq][12] clock64 stamps of CTA (0,0,0)
  long long gld, hld;        // quad kernels: row pitch of the gate buffers (n_dir * 4S) and of the h / c / dh buffers (n_dir * S)
  const __nv_bfloat16* whh;  // rec_q_fwd: packed W_hh [n_dir * 4S, S] bf16 (row-major): the resident slice goes to TENSOR memory
  const __nv_bfloat16* wih;  // rec_q_fwd, fused input projection: packed W_ih [n_dir * 4S, kp] bf16
  int kp;
  int dsmem;                 // rec_q_fwd: the h image goes to the peers by DSMEM bulk copies instead of through the L2 ring
  const __nv_bfloat16* whh_lo;   // rec_q_fwd<.., X3>: low parts of W_hh (whh holds the high parts)
};

#define CL_STAMP(idx)                                                                                            \
  do {                                                                                                           \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg[(size_t)s * 12 + (idx)] = clock64(); \
  } while (0)

using namespace clx;

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// XF: the layer's INPUT projection is fused in (layer 1: <= 80 features).  The CTA keeps its 128 x Kp slice of W_ih in smem,
// the layer input arrives by TMA two steps ahead, and the x(t) W_ih^T MMAs are issued BEFORE the wait for the exchanged h tile
// (they do not depend on it), so they run during the exchange: the [rows, 8S] fp32 pre-activation buffer is neither written by
// a GEMM nor read here (2 x 1.07 GB per step at layer 1 of C4).
template <bool XF>
__global__ void __launch_bounds__(CL_THREADS, 1)
rec_cl_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ CUtensorMap tmWx, RecClParams p) {
  constexpr int TM = CL_TM;
  constexpr int W_BLK = 128 * 128;       // 128 gate rows x 64 bf16
  constexpr int WX_BLK = 128 * 32;       // fused projection: 128 gate rows x 16 bf16
  constexpr int X_BLK = TM * 32;         // fused projection: TM rows x 16 bf16
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.S, KB = S / 64, NC = S / 32;                           // NC = cluster size = producers per tile
  uint8_t* Wsm = smem;                                                   // [KB] blocks of W_BLK
  uint8_t* Asm = Wsm + KB * W_BLK;                                       // [2][NC] k-blocks of FW_IMG
  uint8_t* img = Asm + 2 * NC * FW_IMG;                                  // this CTA's outgoing image
  uint8_t* stg = img + FW_IMG;                                           // [8 epilogue warps] staging tiles of FW_WSTG bytes
  uint8_t* Wx = stg + 8 * FW_WSTG;                                       // XF: [nkx] k-blocks of WX_BLK
  uint8_t* Xs = Wx + (XF ? FW_MAXKX * WX_BLK : 0);                       // XF: [2][nkx] k-blocks of X_BLK
  float* bsm = reinterpret_cast<float*>(Xs + (XF ? 2 * FW_MAXKX * X_BLK : 0));   // XF: [128] gate bias slice
  uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + (XF ? 128 : 0));
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;           // [2]
  uint64_t* mma_done = bars + 3;
  uint64_t* tmem_free = bars + 4;
  uint64_t* stage_ready = bars + 5;
  uint64_t* x_full = bars + 6;           // [2] XF
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int slice = blockIdx.x, dir = blockIdx.y, bt = blockIdx.z;      // cluster = the NC CTAs along x
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tile_bytes = (uint32_t)TM * S * 2;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1);
    mbar_init(a_full, 1);
    mbar_init(a_full + 1, 1);
    mbar_init(mma_done, 1);
    mbar_init(tmem_free, 256);
    mbar_init(stage_ready, 256);
    mbar_init(x_full, 1);
    mbar_init(x_full + 1, 1);
    if (XF) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWx); }
    fence_barrier_init();
  }
  if (XF && threadIdx.x >= 64 && threadIdx.x < 192) bsm[threadIdx.x - 64] = p.bias[(size_t)dir * 4 * S + slice * 128 + threadIdx.x - 64];
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(w_full, KB * W_BLK + (XF ? p.nkx * WX_BLK : 0));
    for (int kb = 0; kb < KB; ++kb) tma_load_2d(&tmW, w_full, Wsm + kb * W_BLK, kb * 64, dir * 4 * S + slice * 128);
    if (XF)
      for (int kx = 0; kx < p.nkx; ++kx) tma_load_3d(&tmWx, w_full, Wx + kx * WX_BLK, kx * 16, dir * 4 * S + slice * 128, 0);
    mbar_expect_tx(a_full, tile_bytes);          // h(0) and h(1); re-armed by the MMA thread after each wait
    mbar_expect_tx(a_full + 1, tile_bytes);
  }
  cluster_sync_all();                            // every CTA's barriers exist before any multicast can signal them

  if (warp == 0) {
    // ---------------- exchange thread ----------------
    if (elect_one()) {
      const uint16_t cmask = (uint16_t)((1u << NC) - 1u);
      const size_t n_cta = (size_t)gridDim.x * gridDim.y * gridDim.z;
      const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      // XF: the layer-input tile of step s_ -> Xs[s_ & 1] (one 16-column box per k-block)
      auto load_x = [&](int s_) {
        const int t_ = dir == 0 ? s_ : p.n_seq - 1 - s_;
        uint64_t* xb = x_full + (s_ & 1);
        mbar_expect_tx(xb, p.nkx * X_BLK);
        for (int kx = 0; kx < p.nkx; ++kx)
          tma_load_3d(&tmX, xb, Xs + ((s_ & 1) * FW_MAXKX + kx) * X_BLK, kx * 16, p.seq_inner ? t_ : bt * TM, p.seq_inner ? bt * TM : t_);
      };
      if (XF) {
        load_x(0);
        if (p.n_seq > 1) load_x(1);
      }
      for (int s = 0; s + 1 < p.n_seq; ++s) {
        uint8_t* slot = p.ring + ((size_t)(s % CL_RING) * n_cta + cta) * FW_IMG;
        mbar_wait_t(stage_ready, s & 1);                           // all 256 epilogue threads have written the image
        CL_STAMP(8);
        bulk_store_wait(slot, img, FW_IMG);
        CL_STAMP(9);
        bulk_load_mc(Asm + ((s & 1) * NC + slice) * FW_IMG, slot, FW_IMG, a_full + (s & 1), cmask);
        CL_STAMP(6);
        if (XF && s + 2 < p.n_seq) load_x(s + 2);                  // the MMAs of step s (reading Xs[s & 1]) completed before its epilogue
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA thread ----------------
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(TM, 128);
      const int nk = S / 16;
      for (int s = XF ? 0 : 1; s < p.n_seq; ++s) {
        if (s == (XF ? 0 : 1)) mbar_wait_t(w_full, 0);
        if (XF) {
          // input projection of step s: independent of the recurrence, issued (and executing) while h(s-1) is exchanged
          if (s > 0) mbar_wait_t(tmem_free, (s - 1) & 1);          // epilogue has drained the accumulator
          mbar_wait_t(x_full + (s & 1), (s >> 1) & 1);
          tc_fence_after();
          const uint32_t x0 = smem_u32(Xs + (s & 1) * FW_MAXKX * X_BLK), wx0 = smem_u32(Wx);
          for (int kx = 0; kx < p.nkx; ++kx)
            mma_bf16_ss(tmem, umma_desc_k32(x0 + kx * X_BLK), umma_desc_k32(wx0 + kx * WX_BLK), idesc, kx != 0);
        }
        if (s > 0) {
          const int b = (s - 1) & 1;
          mbar_wait_t(a_full + b, ((s - 1) >> 1) & 1);               // all NC slices of h(s-1) have landed
          if (s + 2 < p.n_seq) mbar_expect_tx(a_full + b, tile_bytes);   // this buffer next receives h(s+1)
          if (!XF && s > 1) mbar_wait_t(tmem_free, (s - 2) & 1);     // epilogue has drained the accumulator
          CL_STAMP(1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(Asm + b * NC * FW_IMG), w0 = smem_u32(Wsm);
#pragma unroll 4
          for (int kk = 0; kk < nk; ++kk) {
            const uint64_t da = umma_desc_k64(a0 + (kk >> 1) * FW_IMG) + (uint64_t)((kk & 1) * 2);
            const uint64_t db = umma_desc_k128(w0 + (kk >> 2) * W_BLK) + (uint64_t)((kk & 3) * 2);
            mma_bf16_ss(tmem, da, db, idesc, (XF || kk != 0) ? 1u : 0u);
          }
        }
        mma_commit(mma_done);
        CL_STAMP(2);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: lane = (row r, 8 of the group's 16 units) ----------------
    const int q = warp & 3;                        // TMEM sub-partition
    const int hh = warp >= 6 ? 1 : 0;              // epilogue group = 16-unit half of the CTA's 32 units
    const int uh = lane >> 4;                      // 8-unit half inside the group
    const int l16 = lane & 15;
    const int r = q * 16 + l16;
    const int n = bt * TM + r;
    const bool in_range = n < p.n_batch;
    const int len = (in_range && p.lens) ? p.lens[n] : INT_MAX;
    const int wcol = slice * CL_UNITS + hh * 16;                      // first of this WARP's 16 units inside the direction
    const size_t hoff_w = (size_t)dir * S + wcol;
    const size_t goff_w = (size_t)dir * 4 * S + (size_t)wcol * 4;
    // image: row r, 16-byte chunk c = hh*2 + uh of the 64-byte row, 64-byte swizzle
    uint4* img_dst = reinterpret_cast<uint4*>(img + r * 64 + (((hh * 2 + uh) ^ ((r >> 1) & 3)) << 4));
    // Per-warp staging: the warp's 16 rows x 64 gate columns (fp32, 256 B per row) and 16 rows x 16 units of h / c.
    // Global accesses are coalesced (half-warp = one contiguous 256-byte / 64-byte row segment), the lane <-> (row, 8 units)
    // re-distribution happens through these tiles; 16-byte chunks are XOR-swizzled so both access patterns are
    // bank-conflict free.  Only __syncwarp is needed: a warp only ever touches its own tiles.
    uint8_t* wst = stg + (warp - 2) * FW_WSTG;
    float* gs = reinterpret_cast<float*>(wst);                         // [16][64]  chunk c of row l at (c ^ l)
    float* hs = reinterpret_cast<float*>(wst + 4096);                  // [16][16]  chunk c of row l at (c ^ ((l >> 1) & 3))
    float* cs = reinterpret_cast<float*>(wst + 5120);                  // [16][16]
    float creg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) creg[j] = 0.f;

    // coalesced async copy of the warp's pre-activation tile of time t_ into gs (instruction i: rows 2i, 2i+1)
    auto prefetch_g = [&](int t_) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = 2 * i + uh, nn = bt * TM + q * 16 + rl;
        if (nn < p.n_batch)
          cp_async16(gs + rl * 64 + ((l16 ^ rl) << 2),
                     p.xp + ((size_t)t_ * p.rs_seq + (size_t)nn * p.rs_batch) * 8 * S + goff_w + l16 * 4);
      }
      cp_async_commit();
    };
    if (!XF) prefetch_g(dir == 0 ? 0 : p.n_seq - 1);

    for (int s = 0; s < p.n_seq; ++s) {
      const int t = dir == 0 ? s : p.n_seq - 1 - s;
      const bool valid = in_range && t < len;
      float4 g[8];                                 // pre-activations of this lane's 8 units (i,f,g,o each)
      if (XF) {                                    // the projection comes out of the accumulator: start from the bias
#pragma unroll
        for (int j = 0; j < 8; ++j)
          g[j] = valid ? *reinterpret_cast<const float4*>(bsm + (hh * 16 + uh * 8 + j) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        cp_async_wait_all();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          g[j] = valid ? *reinterpret_cast<const float4*>(gs + l16 * 64 + (((uh * 8 + j) ^ l16) << 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (XF || s > 0) {
        uint32_t v[32];
        mbar_wait_t(mma_done, XF ? (s & 1) : ((s - 1) & 1));
        if (threadIdx.x == 64) CL_STAMP(3);
        tc_fence_after();
        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(hh * 64);
        // columns 32..63 (units 8..15 of the row) belong to the upper half-warp: fetched by the row's lane, handed over
        tmem_ld16(ta + 32, v);
        tmem_ld16(ta + 48, v + 16);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float u0 = __shfl_sync(0xffffffffu, __uint_as_float(v[j * 4 + 0]), l16);
          const float u1 = __shfl_sync(0xffffffffu, __uint_as_float(v[j * 4 + 1]), l16);
          const float u2 = __shfl_sync(0xffffffffu, __uint_as_float(v[j * 4 + 2]), l16);
          const float u3 = __shfl_sync(0xffffffffu, __uint_as_float(v[j * 4 + 3]), l16);
          if (uh) { g[j].x += u0; g[j].y += u1; g[j].z += u2; g[j].w += u3; }
        }
        tmem_ld16(ta, v);
        tmem_ld16(ta + 16, v + 16);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(tmem_free);
        if (!uh) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            g[j].x += __uint_as_float(v[j * 4 + 0]); g[j].y += __uint_as_float(v[j * 4 + 1]);
            g[j].z += __uint_as_float(v[j * 4 + 2]); g[j].w += __uint_as_float(v[j * 4 + 3]);
          }
        }
      }
      float hv[8], cv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        hv[j] = 0.f; cv[j] = 0.f;
        if (valid) {
          a.x = sigmoid_apx(g[j].x); a.y = sigmoid_apx(g[j].y); a.z = tanh_apx(g[j].z); a.w = sigmoid_apx(g[j].w);
          cv[j] = fmaf(a.y, creg[j], a.x * a.z);
          hv[j] = a.w * tanh_apx(cv[j]);
        }
        creg[j] = cv[j];
        g[j] = a;
      }
      uint4 hb;
      {
        __nv_bfloat162 b0 = __floats2bfloat162_rn(hv[0], hv[1]), b1 = __floats2bfloat162_rn(hv[2], hv[3]);
        __nv_bfloat162 b2 = __floats2bfloat162_rn(hv[4], hv[5]), b3 = __floats2bfloat162_rn(hv[6], hv[7]);
        hb = make_uint4(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1), *reinterpret_cast<uint32_t*>(&b2),
                        *reinterpret_cast<uint32_t*>(&b3));
      }
      if (s + 1 < p.n_seq) {
        *img_dst = hb;
        fence_proxy_async();                       // generic-proxy smem write -> visible to the bulk (async proxy) store
        mbar_arrive(stage_ready);
      }
      if (threadIdx.x == 64) CL_STAMP(5);
      // ---- off the dependent chain: saved tensors, re-distributed through the warp's staging tiles ----
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(gs + l16 * 64 + (((uh * 8 + j) ^ l16) << 2)) = g[j];
      {
        const int sw = (l16 >> 1) & 3;
        *reinterpret_cast<float4*>(hs + l16 * 16 + (((uh * 2 + 0) ^ sw) << 2)) = make_float4(hv[0], hv[1], hv[2], hv[3]);
        *reinterpret_cast<float4*>(hs + l16 * 16 + (((uh * 2 + 1) ^ sw) << 2)) = make_float4(hv[4], hv[5], hv[6], hv[7]);
        *reinterpret_cast<float4*>(cs + l16 * 16 + (((uh * 2 + 0) ^ sw) << 2)) = make_float4(cv[0], cv[1], cv[2], cv[3]);
        *reinterpret_cast<float4*>(cs + l16 * 16 + (((uh * 2 + 1) ^ sw) << 2)) = make_float4(cv[4], cv[5], cv[6], cv[7]);
      }
      __syncwarp();
      {
        if (in_range) *reinterpret_cast<uint4*>(p.xb + ((size_t)t * p.rs_seq + (size_t)n * p.rs_batch) * 2 * S + hoff_w + uh * 8) = hb;
#pragma unroll
        for (int i = 0; i < 2; ++i) {            // h / c: 8 rows x 64 contiguous bytes per instruction
          const int rl = 8 * i + (lane >> 2), c = lane & 3, nn = bt * TM + q * 16 + rl;
          if (nn < p.n_batch) {
            const size_t o = ((size_t)t * p.rs_seq + (size_t)nn * p.rs_batch) * 2 * S + hoff_w + c * 4;
            const int so = rl * 16 + ((c ^ ((rl >> 1) & 3)) << 2);
            *reinterpret_cast<float4*>(p.hout + o) = *reinterpret_cast<const float4*>(hs + so);
            *reinterpret_cast<float4*>(p.cbuf + o) = *reinterpret_cast<const float4*>(cs + so);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {            // activations: 2 rows x 256 contiguous bytes per instruction
          const int rl = 2 * i + uh, nn = bt * TM + q * 16 + rl;
          if (nn < p.n_batch)
            __stcs(reinterpret_cast<float4*>(p.xp + ((size_t)t * p.rs_seq + (size_t)nn * p.rs_batch) * 8 * S + goff_w + l16 * 4),
                   *reinterpret_cast<const float4*>(gs + rl * 64 + ((l16 ^ rl) << 2)));
        }
      }
      __syncwarp();                                // staging tiles drained
      if (!XF && s + 1 < p.n_seq) prefetch_g(dir == 0 ? s + 1 : p.n_seq - 2 - s);
      if (threadIdx.x == 64) CL_STAMP(7);
    }
  }
  cp_async_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem);
  cluster_sync_all();                            // no CTA leaves while a peer could still address its shared memory
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CL_THREADS, 1)
rec_cl_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmG, RecClParams p) {
  constexpr int TM = CL_TM;
  constexpr int A_BLK = TM * 128;        // one k-block: TM rows x 64 bf16 (a producer contributes two)
  constexpr int W_BLK = 32 * 128;        // 32 unit rows x 64 bf16
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.S, NC = S / 32, NKB = S / 16;   // NKB = 4S/64 k-blocks of the 4S-long reduction
  const int slice_ = blockIdx.x;
  uint8_t* Asm = smem;                    // [NKB] k-blocks of A_BLK
  uint8_t* Wsm = Asm + NKB * A_BLK;       // [NKB] blocks of W_BLK
  uint8_t* img = Asm + slice_ * BW_IMG;   // this CTA's outgoing image = its OWN two k-blocks of the A tile: free once the
                                          // local MMAs are done; the multicast later re-delivers the same bytes to it
  uint8_t* stg = Wsm + NKB * W_BLK;       // [8 epilogue warps] activation staging tiles of 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + 8 * 4096);
  uint64_t* w_full = bars;
  uint64_t* mma_done = bars + 1;
  uint64_t* tmem_free = bars + 2;
  uint64_t* a_free = bars + 3;            // NC remote arrivals per step: every CTA's MMAs have consumed the A tile
  uint64_t* stage_ready = bars + 4;
  uint64_t* img_free = bars + 5;          // the row-major copy of the image has been read out: the image may be rewritten
  uint64_t* full = bars + 6;              // [NC] one per producer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + BW_MAXNC);

  const int slice = blockIdx.x, dir = blockIdx.y, bt = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmG);
    mbar_init(w_full, 1);
    mbar_init(mma_done, 1);
    mbar_init(tmem_free, 256);
    mbar_init(a_free, NC);
    mbar_init(stage_ready, 256);
    mbar_init(img_free, 1);
    for (int i = 0; i < NC; ++i) mbar_init(full + i, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<32>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(w_full, NKB * W_BLK);
    for (int kb = 0; kb < NKB; ++kb) tma_load_2d(&tmW, w_full, Wsm + kb * W_BLK, kb * 64, dir * S + slice * CL_UNITS);
    if (p.n_seq > 1)
      for (int i = 0; i < NC; ++i) mbar_expect_tx(full + i, BW_IMG);
  }
  cluster_sync_all();

  float bsum[32];            // epilogue lanes: running column sums of dG (bias gradient), reduced across rows at the end
#pragma unroll
  for (int j = 0; j < 32; ++j) bsum[j] = 0.f;

  if (warp == 0) {
    // ---------------- exchange thread ----------------
    if (elect_one()) {
      const uint16_t cmask = (uint16_t)((1u << NC) - 1u);
      const size_t n_cta = (size_t)gridDim.x * gridDim.y * gridDim.z;
      const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      const int seq_inner = p.rs_seq < p.rs_batch ? 1 : 0;
      for (int s = 0; s < p.n_seq; ++s) {
        const int t = dir == 0 ? p.n_seq - 1 - s : s;
        uint8_t* slot = p.ring + ((size_t)(s % CL_RING) * n_cta + cta) * BW_IMG;
        mbar_wait_t(stage_ready, s & 1);
        CL_STAMP(8);
        if (s + 1 < p.n_seq) {
          bulk_store_wait(slot, img, BW_IMG);
          CL_STAMP(9);
          if (s > 0) mbar_wait_cluster_t(a_free, (s - 1) & 1);   // every CTA of the cluster has finished reading dG(t_next)
          bulk_load_mc(Asm + slice * BW_IMG, slot, BW_IMG, full + slice, cmask);
          CL_STAMP(6);
        }
        // row-major bf16 dG for the batched weight / input gradient GEMMs: the image IS the swizzled TMA box
        {
          const int c1 = seq_inner ? t : bt * TM, c2 = seq_inner ? bt * TM : t;
          tma_store_3d(&tmG, img, dir * 4 * S + slice * 128, c1, c2);
          tma_store_3d(&tmG, img + A_BLK, dir * 4 * S + slice * 128 + 64, c1, c2);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        mbar_arrive(img_free);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA thread ----------------
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(TM, 32);
      const uint32_t a0 = smem_u32(Asm), w0 = smem_u32(Wsm);
      for (int s = 1; s < p.n_seq; ++s) {
        if (s == 1) mbar_wait_t(w_full, 0);
        if (s > 1) mbar_wait_t(tmem_free, (s - 2) & 1);
        for (int pc = 0; pc < NC; ++pc) {
          mbar_wait_t(full + pc, (s - 1) & 1);                          // producer pc's two k-blocks of dG(t_next) have landed
          if (s + 1 < p.n_seq) mbar_expect_tx(full + pc, BW_IMG);       // re-arm for the next step
          if (pc == 0) CL_STAMP(1);
          tc_fence_after();
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int kb = pc * 2 + h2;
            const uint64_t da = umma_desc_k128(a0 + kb * A_BLK), db = umma_desc_k128(w0 + kb * W_BLK);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_bf16_ss(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
        }
        mma_commit(mma_done);
        CL_STAMP(2);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: lane = (row r, 8 of the group's 16 units) ----------------
    const int q = warp & 3;
    const int hh = warp >= 6 ? 1 : 0;
    const int uh = lane >> 4;
    const int r = q * 16 + (lane & 15);
    const int n = bt * TM + r;
    const bool in_range = n < p.n_batch;
    const int len = (in_range && p.lens) ? p.lens[n] : INT_MAX;
    const int l16 = lane & 15;
    const int ucol = slice * CL_UNITS + hh * 16 + uh * 8;
    const size_t hoff = (size_t)dir * S + ucol;
    const size_t goff_w = (size_t)dir * 4 * S + (size_t)(slice * CL_UNITS + hh * 16) * 4;
    // image: k-block hh, row r, 16-byte chunks uh*4 .. uh*4+3 of the 128-byte row, 128-byte swizzle
    uint8_t* img_row = img + hh * A_BLK + r * 128;
    // per-warp staging of the warp's 16 rows x 64 saved activations (coalesced cp.async in, row-per-lane reads out;
    // chunk c of row l sits at chunk c ^ l: both patterns are bank-conflict free)
    float* gs = reinterpret_cast<float*>(stg + (warp - 2) * 4096);
    auto prefetch_a = [&](int t_) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = 2 * i + uh, nn = bt * TM + q * 16 + rl;
        if (nn < p.n_batch)
          cp_async16(gs + rl * 64 + ((l16 ^ rl) << 2),
                     p.xp + ((size_t)t_ * p.rs_seq + (size_t)nn * p.rs_batch) * 8 * S + goff_w + l16 * 4);
      }
      cp_async_commit();
    };
    prefetch_a(dir == 0 ? p.n_seq - 1 : 0);
    float dcreg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dcreg[j] = 0.f;

    for (int s = 0; s < p.n_seq; ++s) {
      const int t = dir == 0 ? p.n_seq - 1 - s : s;
      const int tp = dir == 0 ? t - 1 : t + 1;
      const bool has_prev = dir == 0 ? (t > 0) : (t < p.n_seq - 1);
      const bool valid = in_range && t < len;
      const bool pv = valid && has_prev && tp < len;
      const size_t row = (size_t)t * p.rs_seq + (size_t)(in_range ? n : 0) * p.rs_batch;
      // everything of this step that does not depend on the recurrence, to registers
      float4 a[8], dh4[2], c4r[2], cp4[2];
      {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        cp_async_wait_all();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          a[j] = valid ? *reinterpret_cast<const float4*>(gs + l16 * 64 + (((uh * 8 + j) ^ l16) << 2)) : z4;
        __syncwarp();
        if (s + 1 < p.n_seq) prefetch_a(dir == 0 ? t - 1 : t + 1);     // a whole step of lead time
        const float4* dp = reinterpret_cast<const float4*>(p.dhout + row * 2 * S + hoff);
        const float4* cp = reinterpret_cast<const float4*>(p.cbuf + row * 2 * S + hoff);
        const float4* pp = reinterpret_cast<const float4*>(
            p.cbuf + ((size_t)(pv ? tp : t) * p.rs_seq + (size_t)(in_range ? n : 0) * p.rs_batch) * 2 * S + hoff);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dh4[j] = valid ? __ldcs(dp + j) : z4;
          c4r[j] = valid ? __ldg(cp + j) : z4;
          cp4[j] = pv ? __ldg(pp + j) : z4;
        }
      }
      float dhm[8];                             // recurrent part of dh for this lane's 8 units
      if (s > 0) {
        uint32_t v[16];
        mbar_wait_t(mma_done, (s - 1) & 1);
        if (warp == 2 && lane < NC) mbar_arrive_remote(a_free, lane);   // this CTA no longer reads its A tile
        if (threadIdx.x == 64) CL_STAMP(3);
        tc_fence_after();
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(hh * 16), v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(tmem_free);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float up = __shfl_sync(0xffffffffu, __uint_as_float(v[8 + k]), lane & 15);
          dhm[k] = uh ? up : __uint_as_float(v[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) dhm[k] = 0.f;
      }
      uint4 gq[4];                              // this lane's 8 units x 4 gates of dG, bf16
#pragma unroll
      for (int qd = 0; qd < 2; ++qd) {
        const float dhv[4] = {dh4[qd].x, dh4[qd].y, dh4[qd].z, dh4[qd].w};
        const float cv[4] = {c4r[qd].x, c4r[qd].y, c4r[qd].z, c4r[qd].w};
        const float cpv[4] = {cp4[qd].x, cp4[qd].y, cp4[qd].z, cp4[qd].w};
        uint32_t gb[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = qd * 4 + u;
          float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
          float dco = 0.f;
          if (valid) {
            const float4 aa = a[j];
            const float dh = dhv[u] + dhm[j];
            const float tc_ = tanh_apx(cv[u]);
            const float dc = fmaf(dh * aa.w, 1.f - tc_ * tc_, dcreg[j]);
            dg.w = dh * tc_ * aa.w * (1.f - aa.w);
            dg.x = dc * aa.z * aa.x * (1.f - aa.x);
            dg.z = dc * aa.x * (1.f - aa.z * aa.z);
            dg.y = pv ? dc * cpv[u] * aa.y * (1.f - aa.y) : 0.f;
            dco = dc * aa.y;
          }
          dcreg[j] = dco;
          bsum[j * 4 + 0] += dg.x; bsum[j * 4 + 1] += dg.y; bsum[j * 4 + 2] += dg.z; bsum[j * 4 + 3] += dg.w;
          __nv_bfloat162 b01 = __floats2bfloat162_rn(dg.x, dg.y), b23 = __floats2bfloat162_rn(dg.z, dg.w);
          gb[u * 2 + 0] = *reinterpret_cast<uint32_t*>(&b01);
          gb[u * 2 + 1] = *reinterpret_cast<uint32_t*>(&b23);
        }
        gq[qd * 2 + 0] = make_uint4(gb[0], gb[1], gb[2], gb[3]);
        gq[qd * 2 + 1] = make_uint4(gb[4], gb[5], gb[6], gb[7]);
      }
      if (s > 0) mbar_wait_t(img_free, (s - 1) & 1);     // the previous image has been read out (long done in practice)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(img_row + (((uh * 4 + c) ^ (r & 7)) << 4)) = gq[c];
      fence_proxy_async();
      mbar_arrive(stage_ready);
      if (threadIdx.x == 64) CL_STAMP(5);
    }
  }
  cp_async_wait_all();
  if (p.dbias) {                      // bias gradient: reduce the per-lane partial sums across the tile rows
    __syncthreads();                  // all MMAs have been consumed: the A tile is free to serve as scratch
    float* scr = reinterpret_cast<float*>(Asm);     // [TM][128]
    if (warp >= 2) {
      const int q = warp & 3, hh = warp >= 6 ? 1 : 0, uh = lane >> 4, r = q * 16 + (lane & 15);
      float* dst = scr + r * 128 + hh * 64 + uh * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(dst + ((j ^ (r & 7)) << 2)) = make_float4(bsum[j * 4], bsum[j * 4 + 1], bsum[j * 4 + 2], bsum[j * 4 + 3]);
    }
    __syncthreads();
    if (threadIdx.x < 128) {
      const int col = threadIdx.x, blk = col >> 5, j = (col >> 2) & 7, e = col & 3;
      float acc = 0.f;
      for (int rr = 0; rr < TM; ++rr) acc += scr[rr * 128 + blk * 32 + ((j ^ (rr & 7)) << 2) + e];
      atomicAdd(p.dbias + (size_t)dir * 4 * S + (size_t)slice * 128 + col, acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<32>(tmem);
  cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------
// backward, "quad" formulation: 64 hidden units per CTA, S/64 CTAs per cluster, R (32 or 16) batch rows per tile,
// TRANSPOSED product.  Why: the kernel above is bound by what every SM has to ingest per dependent step -- the all-gathered
// dG tile is rows x 4S bf16 = 128 KB at 64 rows, and 64 MMAs of N=32 wait on it.  Here
//     dh^T [64 units (M) x R rows (N)] = W_hh^T slice [64 x 4S] (A, resident, K-major)  x  dG tile [R rows x 4S] (B, K-major)
// so the tile an SM ingests per step shrinks with R (64 KB at R=32, 32 KB at R=16) while the MMA count stays 4S/16 and the
// per-instruction cost stays at the ~27-cycle issue floor (N <= 32).  Half as many CTAs per cluster (4 at S=256) also halves
// the number of producers every step has to wait for.  The accumulator's TMEM lane is now a UNIT and its column a batch
// row: an epilogue thread owns one unit x R/4 rows, so (a) all streaming operands (saved activations, dhout, c) are read
// straight from global memory with unit-contiguous (coalesced) accesses one whole exchange ahead of their use -- no
// staging tiles -- and (b) the bias gradient is 4 registers per thread instead of 32.
// Exchange protocol, image = own slot of the operand tile, `a_free` / `img_free` hand-shakes: exactly as above.
// ------------------------------------------------------------------------------------------------
constexpr int Q_UNITS = 64;
constexpr int Q_MAXNC = 4;
// 16 epilogue warps (4 per scheduler): with 8 the per-step cell math of a tile is issue- / latency-bound on two warps per
// scheduler (measured: ~300 cycles per cell and thread); TMEM sub-partition = warp % 4, the four warps of a sub-partition
// split the accumulator columns (and, forward, the two M=128 halves)
constexpr int Q_EPW = 16;
constexpr int Q_THREADS = 64 + 32 * Q_EPW;

template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t ta, uint32_t* v) {
  if constexpr (N == 32) {
    tmem_ld32(ta, v);
  } else if constexpr (N == 16) {
    tmem_ld16(ta, v);
  } else if constexpr (N == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(ta) : "memory");
  } else {
    static_assert(N == 4, "tmem_ld_n: 4 / 8 / 16 / 32 columns");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(ta) : "memory");
  }
}

template <int R>
__global__ void __launch_bounds__(Q_THREADS, 1)
rec_q_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmG, RecClParams p) {
  constexpr int B_BLK = R * 128;          // one k-block of the dG tile: R rows x 64 bf16
  constexpr int W_BLK = Q_UNITS * 128;    // one k-block of the weight slice: 64 unit rows x 64 bf16
  constexpr int IMG = 4 * B_BLK;          // a producer's image: its 64 units x 4 gates = 4 k-blocks
  constexpr int CPT = R / 8;              // cells (batch rows of ONE unit) per epilogue thread
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.S, NC = S / Q_UNITS, NKB = S / 16;     // NKB = 4S/64 k-blocks of the 4S-long reduction
  const int slice = blockIdx.x, dir = blockIdx.y, bt = blockIdx.z;
  uint8_t* Wsm = smem;                    // [NKB] blocks of W_BLK
  uint8_t* Bsm = Wsm + NKB * W_BLK;       // [NKB] blocks of B_BLK
  uint8_t* img = Bsm + slice * IMG;       // own slot of the tile = outgoing image
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bsm + NKB * B_BLK);
  uint64_t* w_full = bars;
  uint64_t* mma_done = bars + 1;
  uint64_t* tmem_free = bars + 2;
  uint64_t* a_free = bars + 3;            // NC remote arrivals per step: every CTA's MMAs have consumed its tile
  uint64_t* stage_ready = bars + 4;
  uint64_t* img_free = bars + 5;
  uint64_t* full = bars + 6;              // [NC] one per producer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + Q_MAXNC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmG);
    mbar_init(w_full, 1);
    mbar_init(mma_done, 1);
    mbar_init(tmem_free, Q_EPW);          // one arrival per epilogue WARP (512 per-thread arrivals serialise on the barrier word)
    mbar_init(a_free, NC);
    mbar_init(stage_ready, Q_EPW);
    mbar_init(img_free, 1);
    for (int i = 0; i < NC; ++i) mbar_init(full + i, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<32>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(w_full, NKB * W_BLK);
    for (int kb = 0; kb < NKB; ++kb) tma_load_2d(&tmW, w_full, Wsm + kb * W_BLK, kb * 64, dir * S + slice * Q_UNITS);
    if (p.n_seq > 1)
      for (int i = 0; i < NC; ++i) mbar_expect_tx(full + i, IMG);
  }
  cluster_sync_all();

  if (warp == 0) {
    // ---------------- exchange thread ----------------
    if (elect_one()) {
      const uint16_t cmask = (uint16_t)((1u << NC) - 1u);
      const size_t n_cta = (size_t)gridDim.x * gridDim.y * gridDim.z;
      const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      const int seq_inner = p.rs_seq < p.rs_batch ? 1 : 0;
      for (int s = 0; s < p.n_seq; ++s) {
        const int t = dir == 0 ? p.n_seq - 1 - s : s;
        uint8_t* slot = p.ring + ((size_t)(s % CL_RING) * n_cta + cta) * IMG;
        mbar_wait_t(stage_ready, s & 1);
        CL_STAMP(8);
        if (s + 1 < p.n_seq) {
          bulk_store_wait(slot, img, IMG);
          CL_STAMP(9);
          if (s > 0) mbar_wait_t(a_free, (s - 1) & 1);           // every CTA of the cluster has finished reading dG(t_next)
          CL_STAMP(11);
          bulk_load_mc(img, slot, IMG, full + slice, cmask);
          CL_STAMP(6);
        }
        // row-major bf16 dG for the batched weight / input gradient GEMMs: each image k-block IS a swizzled TMA box
        {
          const int c1 = seq_inner ? t : bt * R, c2 = seq_inner ? bt * R : t;
#pragma unroll
          for (int i = 0; i < 4; ++i) tma_store_3d(&tmG, img + i * B_BLK, dir * 4 * S + slice * 256 + i * 64, c1, c2);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        mbar_arrive(img_free);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA warp: one elected lane issues, lanes 0..NC-1 release the tile to the cluster ----------------
    constexpr uint32_t idesc = umma_idesc_bf16(Q_UNITS, R);
    const uint32_t b0 = smem_u32(Bsm), w0 = smem_u32(Wsm);
    const bool leader = elect_one();
    for (int s = 1; s < p.n_seq; ++s) {
      if (leader) {
        if (s == 1) mbar_wait_t(w_full, 0);
        if (s > 1) mbar_wait_t(tmem_free, (s - 2) & 1);
        // The issuing thread runs only ~6 MMAs ahead of the tensor pipe, so whatever it does between two producers' groups
        // must stay below ~150 cycles or the pipe drains: one (usually already complete) barrier wait and a fence only; the
        // barriers are re-armed after the last MMA has been issued (no producer can deliver the next tile before this CTA's
        // a_free arrive, which follows mma_done).
        for (int pc = 0; pc < NC; ++pc) {
          mbar_wait_t(full + pc, (s - 1) & 1);                          // producer pc's four k-blocks of dG(t_next) have landed
          if (pc == 0) CL_STAMP(1);
          tc_fence_after();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int kb = pc * 4 + i;
            const uint64_t da = umma_desc_k128(w0 + kb * W_BLK), db = umma_desc_k128(b0 + kb * B_BLK);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_bf16_ss(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
        }
        mma_commit(mma_done);
        CL_STAMP(2);
        if (s + 1 < p.n_seq)
          for (int pc = 0; pc < NC; ++pc) mbar_expect_tx(full + pc, IMG);   // re-arm for the next step
      }
      __syncwarp();
      // this CTA's tile is free again once its MMAs have completed: tell every producer of the cluster (one lane each).  Done
      // HERE and not by an epilogue warp: a cluster-scope release there has to drain the warp's outstanding global loads
      // (~1 k cycles on the dependent chain, measured); this warp has none and nothing waits for it before the next exchange.
      if (lane < NC) {
        mbar_wait_t(mma_done, (s - 1) & 1);
        mbar_arrive_remote_relaxed(a_free, lane);
      }
      __syncwarp();
    }
  } else {
    // ---------------- epilogue: thread = (unit u, CPT batch rows) ----------------
    const int q = warp & 3;                        // TMEM sub-partition: units q*16 .. q*16+15 of the CTA
    const int cg = (warp - 2) >> 2;                // column group: batch rows cg*R/4 .. +R/4 of the tile
    const int uh = lane >> 4, l16 = lane & 15;
    const int ul = q * 16 + l16;                   // unit inside the CTA
    const int r0 = cg * (R / 4) + uh * CPT;        // first of this thread's CPT tile rows
    const size_t hcol = (size_t)dir * S + slice * Q_UNITS + ul;
    const size_t gcol = (size_t)dir * 4 * S + (size_t)(slice * Q_UNITS + ul) * 4;
    int len[CPT];
    size_t rowb[CPT];                              // batch part of the row index
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int n = bt * R + r0 + j;
      const bool in_range = n < p.n_batch;
      len[j] = in_range ? (p.lens ? p.lens[n] : INT_MAX) : 0;          // out-of-range rows: never valid
      rowb[j] = (size_t)(in_range ? n : 0) * p.rs_batch;
    }
    // image: k-block q, tile row r, bytes l16*8 .. +8 of the 128-byte row (16-byte chunk l16/2), 128-byte swizzle
    uint8_t* img_u = img + q * B_BLK + (l16 & 1) * 8;
    float dcreg[CPT], bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < CPT; ++j) dcreg[j] = 0.f;

    for (int s = 0; s < p.n_seq; ++s) {
      const int t = dir == 0 ? p.n_seq - 1 - s : s;
      const int tp = dir == 0 ? t - 1 : t + 1;
      const bool has_prev = dir == 0 ? (t > 0) : (t < p.n_seq - 1);
      // everything of this step that does not depend on the recurrence, to registers (in flight during the exchange)
      float4 a[CPT];
      float dhv[CPT], cv[CPT], cpv[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const bool valid = t < len[j];
        const bool pv = valid && has_prev && tp < len[j];
        const size_t row = (size_t)t * p.rs_seq + rowb[j];
        a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        dhv[j] = 0.f; cv[j] = 0.f; cpv[j] = 0.f;
        if (valid) {
          a[j] = __ldcs(reinterpret_cast<const float4*>(p.xp + row * p.gld + gcol));
          dhv[j] = __ldcs(p.dhout + row * p.hld + hcol);
          cv[j] = __ldg(p.cbuf + row * p.hld + hcol);
        }
        if (pv) cpv[j] = __ldg(p.cbuf + ((size_t)tp * p.rs_seq + rowb[j]) * p.hld + hcol);
      }
      if (s > 0) mbar_wait_t(img_free, (s - 1) & 1);     // the previous image has been read out: checked off the dependent chain
      float dhm[CPT];                           // recurrent part of dh for this thread's cells
      if (s > 0) {
        uint32_t v[2 * CPT];
        mbar_wait_t(mma_done, (s - 1) & 1);
        if (threadIdx.x == 64) CL_STAMP(3);
        tc_fence_after();
        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * (R / 4));
        tmem_ld_n<2 * CPT>(ta, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_free);
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const float up = __shfl_sync(0xffffffffu, __uint_as_float(v[CPT + k]), l16);
          dhm[k] = uh ? up : __uint_as_float(v[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < CPT; ++k) dhm[k] = 0.f;
      }
      uint2 gq[CPT];                            // 4 gates of dG per cell, bf16
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const bool valid = t < len[j];
        const bool pv = valid && has_prev && tp < len[j];
        float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
        float dco = 0.f;
        if (valid) {
          const float4 aa = a[j];
          const float dh = dhv[j] + dhm[j];
          const float tc_ = tanh_apx(cv[j]);
          const float dc = fmaf(dh * aa.w, 1.f - tc_ * tc_, dcreg[j]);
          dg.w = dh * tc_ * aa.w * (1.f - aa.w);
          dg.x = dc * aa.z * aa.x * (1.f - aa.x);
          dg.z = dc * aa.x * (1.f - aa.z * aa.z);
          dg.y = pv ? dc * cpv[j] * aa.y * (1.f - aa.y) : 0.f;
          dco = dc * aa.y;
        }
        dcreg[j] = dco;
        bsum[0] += dg.x; bsum[1] += dg.y; bsum[2] += dg.z; bsum[3] += dg.w;
        __nv_bfloat162 b01 = __floats2bfloat162_rn(dg.x, dg.y), b23 = __floats2bfloat162_rn(dg.z, dg.w);
        gq[j] = make_uint2(*reinterpret_cast<uint32_t*>(&b01), *reinterpret_cast<uint32_t*>(&b23));
      }
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int r = r0 + j;
        *reinterpret_cast<uint2*>(img_u + r * 128 + (((l16 >> 1) ^ (r & 7)) << 4)) = gq[j];
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(stage_ready);
      if (threadIdx.x == 64) CL_STAMP(5);
    }
    if (p.dbias) {                      // bias gradient: the two half-warps hold the same unit, then one atomic per gate
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float v = bsum[g] + __shfl_xor_sync(0xffffffffu, bsum[g], 16);
        if (!uh) atomicAdd(p.dbias + gcol + g, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<32>(tmem);
  cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------
// forward, "quad" formulation (same decomposition as rec_q_bwd_kernel): 64 hidden units = 256 gate rows per CTA as two
// M=128 halves, transposed product
//     gates^T [128 gate rows (M) x R batch rows (N)] (+)= W_hh slice [128 x S] (A, resident)  x  h(t-1) tile [R rows x S] (B)
// Per step an SM ingests R x S bf16 (16 KB at R=32) instead of 32 KB, exchanges a 4 KB image instead of 4 KB x 8 producers,
// and waits for S/64 producers instead of S/32.  TMEM lane = gate row (unit*4 + gate), column = batch row: the four gates of
// a unit sit in four adjacent lanes, a 4x4 shuffle transpose per four batch rows hands lane g' of the quad the cells
// (unit, rows 4m + g'); every thread then owns ONE unit x R/4 rows, and all global traffic (pre-activations in; activations,
// h, c out) is issued straight from registers with 8 consecutive units per row segment (128-byte lines for the gate
// buffers) -- no staging tiles.  The bf16 h copy (next layer's input / weight-gradient operand) is TMA-stored from the image.
// XF: the layer's input projection is fused in exactly as in rec_cl_fwd_kernel<true>.
// ------------------------------------------------------------------------------------------------
// X3 (forward-only exact path: greedy decoding / validation): W_hh and h are split into bf16 hi + lo parts and the product is
// W_hi h_hi + W_hi h_lo + W_lo h_hi (fp32-accurate: what rec_tc_fwd_kernel<.., X3> computes, at the cluster kernel's step
// latency).  W_hi is the TMEM-resident A operand of the first two terms, W_lo lives in shared memory (SS form) in the SAME
// permuted row order; a producer's image carries the hi and the lo k-block of its 64 units; ex2 / rcp based activations with 1e-7 absolute error (cl_common.cuh); only
// `hout` is written (no activations / cell states: nothing runs backward through this path).
template <int R, bool XF, bool X3 = false>
__global__ void __launch_bounds__(Q_THREADS, 1)
rec_q_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmWx, const __grid_constant__ CUtensorMap tmH, RecClParams p) {
  constexpr int W_BLK = 256 * 128;       // one k-block of the weight slice: 256 gate rows x 64 bf16
  constexpr int IMG = R * 128;           // one k-block of the h tile: R rows x 64 units bf16
  constexpr int PARTS = X3 ? 2 : 1;      // operand parts of h (hi [, lo])
  constexpr int IMGX = PARTS * IMG;      // one producer's image: its k-block of every part
  static_assert(!(X3 && XF), "the exact path takes its pre-activations from the split-operand GEMM");
  constexpr int WX_BLK = 256 * 32;       // fused projection: 256 gate rows x 16 bf16
  constexpr int X_BLK = R * 32;          // fused projection: R rows x 16 bf16
  constexpr int CPT = R / 8;             // cells (batch rows of ONE unit) per epilogue thread
  constexpr int WCOL = 64;               // TMEM: accumulators in columns [0, 2R), the W_hh slice from column 64:
  constexpr int TCOLS = 512;             // half 0 (gate rows 0..127) in [64, 64 + S/2), half 1 behind it (S <= 256)
  constexpr int WXCOL = WCOL + 256;      // XF: the W_ih slice, FW_MAXKX * 8 columns per half
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.S, KB = S / 64, NC = S / Q_UNITS;
  uint8_t* Wsm = smem;                                                   // [KB] blocks of W_BLK
  uint8_t* Hsm = Wsm + KB * W_BLK;                                       // [2][NC][PARTS] k-blocks of IMG
  uint8_t* img = Hsm + 2 * NC * IMGX;                                    // (spare: the image is written in place)
  uint8_t* Wx = img + (IMG < 1024 ? 1024 : IMG);                         // XF: [nkx] k-blocks of WX_BLK
  uint8_t* Xs = Wx + (XF ? FW_MAXKX * WX_BLK : 0);                       // XF: [2][nkx] k-blocks of X_BLK
  float* bsm = reinterpret_cast<float*>(Xs + (XF ? 2 * FW_MAXKX * X_BLK : 0));   // XF: [256] gate bias slice
  uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + (XF ? 256 : 0));
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;           // [2]
  uint64_t* mma_done = bars + 3;
  uint64_t* tmem_free = bars + 4;
  uint64_t* stage_ready = bars + 5;
  uint64_t* x_full = bars + 6;           // [2] XF
  uint64_t* img_free = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int slice = blockIdx.x, dir = blockIdx.y, bt = blockIdx.z;      // cluster = the NC CTAs along x
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tile_bytes = (uint32_t)(NC - 1) * IMGX;     // the CTA's own slice of h is written in place by its epilogue
  const int seq_inner = p.rs_seq < p.rs_batch ? 1 : 0;
  if (warp == 0 && elect_one()) {
    if (!X3) {
      tma_prefetch_desc(&tmW);
      tma_prefetch_desc(&tmH);
    }
    mbar_init(w_full, 1);
    mbar_init(a_full, 1);
    mbar_init(a_full + 1, 1);
    mbar_init(mma_done, 1);
    mbar_init(tmem_free, Q_EPW);          // one arrival per epilogue WARP (512 per-thread arrivals serialise on the barrier word)
    mbar_init(stage_ready, Q_EPW);
    mbar_init(x_full, 1);
    mbar_init(x_full + 1, 1);
    mbar_init(img_free, 1);
    if (XF) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWx); }
    fence_barrier_init();
  }
  if (XF && threadIdx.x >= 64 && threadIdx.x < 320) bsm[threadIdx.x - 64] = p.bias[(size_t)dir * 4 * S + slice * 256 + threadIdx.x - 64];
  if (warp == 1) tmem_alloc<TCOLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    if (!X3) {
      mbar_expect_tx(w_full, KB * W_BLK + (XF ? p.nkx * WX_BLK : 0));
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(&tmW, w_full, Wsm + kb * W_BLK, kb * 64, dir * 4 * S + slice * 256);
    }
    if (XF)
      for (int kx = 0; kx < p.nkx; ++kx) tma_load_3d(&tmWx, w_full, Wx + kx * WX_BLK, kx * 16, dir * 4 * S + slice * 256, 0);
    mbar_expect_tx(a_full, tile_bytes);          // h(0) and h(1); re-armed by the MMA thread after each wait
    mbar_expect_tx(a_full + 1, tile_bytes);
  }
  if (warp >= 2) {
    // resident W_hh slice -> tensor memory (A operand of every step's product): thread = one gate row, the 16 warps split
    // the two M = 128 halves and the two halves of K.  Row order inside a 32-lane sub-partition: lane = gate * 8 + unit
    // (not unit * 4 + gate): a 16-lane x 256-bit accumulator load then hands thread t gates {0,1} / {2,3} of unit t/4 for
    // batch columns 2 (t % 4), + 1 -- all four gates of its cells without any shuffle.
    const int q_ = warp & 3, half_ = ((warp - 2) >> 2) & 1, kh_ = (warp - 2) >> 3;
    const size_t grow = (size_t)dir * 4 * S + slice * 256 + half_ * 128 + (q_ * 8 + (lane & 7)) * 4 + (lane >> 3);
    const uint32_t tl = tmem + ((uint32_t)(q_ * 32) << 16);
    tmem_store_row(tl + (uint32_t)(WCOL + half_ * (S / 2) + kh_ * (S / 4)), p.whh + grow * S + kh_ * (S / 2), S / 4);
    if (X3) {
      // the low part of the slice -> shared memory, SW128 K-major k-blocks of 256 rows, row i = the TMEM lane order above
      const uint4* src = reinterpret_cast<const uint4*>(p.whh_lo + grow * S + kh_ * (S / 2));
      const int i = half_ * 128 + q_ * 32 + lane;
      for (int j = 0; j < S / 16; ++j) {           // 16-byte chunks of this thread's half row
        const int kg = kh_ * (S / 2) + j * 8, kb = kg >> 6, c = (kg & 63) >> 3;
        *reinterpret_cast<uint4*>(Wsm + kb * W_BLK + i * 128 + ((c ^ (i & 7)) << 4)) = __ldg(src + j);
      }
      fence_proxy_async();
    }
    if (XF && kh_ == 0) {                        // the W_ih slice of the fused input projection, same row order
      const uint4* xr = reinterpret_cast<const uint4*>(p.wih + grow * p.kp);
      const int nq = p.kp / 8;                   // 16-byte groups of the row
      for (int kx = 0; kx < p.nkx; ++kx) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 w4 = make_uint4(0u, 0u, 0u, 0u);
          if (2 * kx + j < nq) w4 = __ldg(xr + 2 * kx + j);
          v[4 * j] = w4.x; v[4 * j + 1] = w4.y; v[4 * j + 2] = w4.z; v[4 * j + 3] = w4.w;
        }
        tmem_st8(tl + (uint32_t)(WXCOL + half_ * (FW_MAXKX * 8) + kx * 8), v);
      }
      tmem_st_wait();
    }
    tc_fence_before();
  }
  cluster_sync_all();                            // every CTA's barriers exist before any multicast can signal them

  if (warp == 0) {
    // ---------------- exchange thread ----------------
    if (elect_one()) {
      const uint16_t cmask = (uint16_t)(((1u << NC) - 1u) & ~(1u << slice));   // the peers
      const size_t n_cta = (size_t)gridDim.x * gridDim.y * gridDim.z;
      const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      auto load_x = [&](int s_) {
        const int t_ = dir == 0 ? s_ : p.n_seq - 1 - s_;
        uint64_t* xb = x_full + (s_ & 1);
        mbar_expect_tx(xb, p.nkx * X_BLK);
        for (int kx = 0; kx < p.nkx; ++kx)
          tma_load_3d(&tmX, xb, Xs + ((s_ & 1) * FW_MAXKX + kx) * X_BLK, kx * 16, seq_inner ? t_ : bt * R, seq_inner ? bt * R : t_);
      };
      if (XF) {
        load_x(0);
        if (p.n_seq > 1) load_x(1);
      }
      for (int s = 0; s < p.n_seq; ++s) {
        const int t = dir == 0 ? s : p.n_seq - 1 - s;
        uint8_t* slot = p.ring + ((size_t)(s % CL_RING) * n_cta + cta) * IMGX;
        mbar_wait_t(stage_ready, s & 1);                           // all 256 epilogue threads have written the image
        CL_STAMP(8);
        uint8_t* img_s = Hsm + ((s & 1) * NC + slice) * IMGX;      // the image = this CTA's k-block(s) of the tile h(s)
        if (s + 1 < p.n_seq) {
          if (p.dsmem) {
            // shared -> shared of every peer.  The image (buffer s & 1) is rewritten at step s + 2, after this CTA has received
            // the peers' h(s + 1), which they computed after consuming THESE copies
            for (int d = 0; d < NC; ++d)
              if (d != slice) bulk_copy_dsmem(img_s, img_s, IMGX, a_full + (s & 1), (uint32_t)d);
          } else {
            bulk_store_wait(slot, img_s, IMGX);
            CL_STAMP(9);
            bulk_load_mc(img_s, slot, IMGX, a_full + (s & 1), cmask);
          }
          CL_STAMP(6);
        }
        // bf16 copy of h (next layer's input, weight-gradient operand): the image IS the swizzled TMA box
        if (!X3) {
          tma_store_3d(&tmH, img_s, dir * S + slice * Q_UNITS, seq_inner ? t : bt * R, seq_inner ? bt * R : t);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        mbar_arrive(img_free);
        if (XF && s + 2 < p.n_seq) load_x(s + 2);                  // the MMAs of step s (reading Xs[s & 1]) completed before its epilogue
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA thread ----------------
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, R);
      const int nk = S / 16;
      // every operand of the step's MMAs is a base + a compile-time offset: nothing but the instructions themselves sits
      // between the arrival of the exchanged tile and the commit (the issue loop with computed descriptors took 28 cycles
      // per instruction, three times the tensor core's own 8 - 13)
      const uint64_t dH[2] = {umma_desc_k128(smem_u32(Hsm)), umma_desc_k128(smem_u32(Hsm + NC * IMGX))};
      const uint64_t dWl = umma_desc_k128(smem_u32(Wsm));           // X3: low part of the slice (half 1: + 128 rows = 16 KB)
      const uint32_t aw0 = tmem + WCOL, aw1 = tmem + WCOL + S / 2;
      const uint64_t own_h = (uint64_t)((slice * IMGX) >> 4);
      const uint32_t own_a = (uint32_t)(slice * 32);
      for (int s = XF ? 0 : 1; s < p.n_seq; ++s) {
        if (!X3 && s == (XF ? 0 : 1)) mbar_wait_t(w_full, 0);
        if (XF) {
          // input projection of step s: independent of the recurrence, issued (and executing) while h(s-1) is exchanged
          if (s > 0) mbar_wait_t(tmem_free, (s - 1) & 1);          // epilogue has drained the accumulators
          mbar_wait_t(x_full + (s & 1), (s >> 1) & 1);
          tc_fence_after();
          const uint64_t dx0 = umma_desc_k32(smem_u32(Xs + (s & 1) * FW_MAXKX * X_BLK));
#pragma unroll
          for (int kx = 0; kx < FW_MAXKX; ++kx) {
            if (kx < p.nkx) {
              mma_bf16_ts(tmem, tmem + WXCOL + kx * 8, dx0 + (uint64_t)((kx * X_BLK) >> 4), idesc, kx != 0);
              mma_bf16_ts(tmem + R, tmem + WXCOL + FW_MAXKX * 8 + kx * 8, dx0 + (uint64_t)((kx * X_BLK) >> 4), idesc, kx != 0);
            }
          }
        }
        if (s > 0) {
          const int b = (s - 1) & 1;
          const uint64_t dhb = dH[b];
          // the CTA's OWN k-block of h(s-1) was written in place by its epilogue: its MMAs run while the peers' slices are
          // still on their way through the exchange
          mbar_wait_t(stage_ready, (s - 1) & 1);
          if (!XF && s > 1) mbar_wait_t(tmem_free, (s - 2) & 1);     // epilogue has drained the accumulators
          tc_fence_after();
#pragma unroll
          for (int kq = 0; kq < 4; ++kq) {
            mma_bf16_ts(tmem, aw0 + own_a + kq * 8, dhb + own_h + (uint64_t)(kq * 2), idesc, (XF || kq != 0) ? 1u : 0u);
            mma_bf16_ts(tmem + R, aw1 + own_a + kq * 8, dhb + own_h + (uint64_t)(kq * 2), idesc, (XF || kq != 0) ? 1u : 0u);
            if (X3) {
              const uint64_t dlo = dhb + own_h + (uint64_t)(IMG >> 4) + (uint64_t)(kq * 2);
              const uint64_t dw = dWl + (uint64_t)((slice * W_BLK) >> 4) + (uint64_t)(kq * 2);
              mma_bf16_ts(tmem, aw0 + own_a + kq * 8, dlo, idesc, 1u);
              mma_bf16_ts(tmem + R, aw1 + own_a + kq * 8, dlo, idesc, 1u);
              mma_bf16_ss(tmem, dw, dhb + own_h + (uint64_t)(kq * 2), idesc, 1u);
              mma_bf16_ss(tmem + R, dw + (uint64_t)(16384 >> 4), dhb + own_h + (uint64_t)(kq * 2), idesc, 1u);
            }
          }
          mbar_wait_t(a_full + b, ((s - 1) >> 1) & 1);               // the NC - 1 remote slices of h(s-1) have landed
          if (s + 2 < p.n_seq) mbar_expect_tx(a_full + b, tile_bytes);   // this buffer next receives h(s+1)
          CL_STAMP(1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            if (kk < nk && (kk >> 2) != slice) {
              const uint64_t dh = dhb + (uint64_t)(((kk >> 2) * IMGX) >> 4) + (uint64_t)((kk & 3) * 2);
              mma_bf16_ts(tmem, aw0 + kk * 8, dh, idesc, 1u);
              mma_bf16_ts(tmem + R, aw1 + kk * 8, dh, idesc, 1u);
              if (X3) {
                const uint64_t dw = dWl + (uint64_t)(((kk >> 2) * W_BLK) >> 4) + (uint64_t)((kk & 3) * 2);
                mma_bf16_ts(tmem, aw0 + kk * 8, dh + (uint64_t)(IMG >> 4), idesc, 1u);
                mma_bf16_ts(tmem + R, aw1 + kk * 8, dh + (uint64_t)(IMG >> 4), idesc, 1u);
                mma_bf16_ss(tmem, dw, dh, idesc, 1u);
                mma_bf16_ss(tmem + R, dw + (uint64_t)(16384 >> 4), dh, idesc, 1u);
              }
            }
          }
        }
        mma_commit(mma_done);
        CL_STAMP(2);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: thread = (unit u, CPT batch rows 4m + gp) ----------------
    const int q = warp & 3;                        // TMEM sub-partition: gate rows q*32 .. q*32+31 of the half
    const int half = ((warp - 2) >> 2) & 1;        // M=128 half of the CTA's 256 gate rows
    const int cgp = (warp - 2) >> 3;               // column group: batch rows cgp*R/2 .. +R/2 of the tile
    const int uq = lane >> 2, gp = lane & 3;
    const int ul = half * 32 + q * 8 + uq;         // unit inside the CTA
    const size_t hcol = (size_t)dir * S + slice * Q_UNITS + ul;
    const size_t gcol = (size_t)dir * 4 * S + (size_t)(slice * Q_UNITS + ul) * 4;
    int len[CPT];
    size_t rowb[CPT];
    bool inr[CPT];
#pragma unroll
    for (int m = 0; m < CPT; ++m) {
      const int n = bt * R + cgp * (R / 2) + 8 * (m >> 1) + 2 * gp + (m & 1);
      inr[m] = n < p.n_batch;
      len[m] = inr[m] ? (p.lens ? p.lens[n] : INT_MAX) : 0;
      rowb[m] = (size_t)(inr[m] ? n : 0) * p.rs_batch;
    }
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (XF) bias4 = *reinterpret_cast<const float4*>(bsm + ul * 4);
    float creg[CPT];
#pragma unroll
    for (int m = 0; m < CPT; ++m) creg[m] = 0.f;
    uint8_t* img_u0 = Hsm + slice * IMGX + (ul & 7) * 2;   // own k-block of tile buffer 0: + row*128 + swizzled 16-byte chunk (ul >> 3)

    // pre-activations are fetched ONE WHOLE STEP ahead (registers): a load issued at the top of its own step is not back
    // when the accumulator is (measured: the epilogue then stalls on HBM latency under the kernel's own store bursts)
    float4 gn[CPT];
    auto fetch_g = [&](int s_) {
      const int t_ = dir == 0 ? s_ : p.n_seq - 1 - s_;
#pragma unroll
      for (int m = 0; m < CPT; ++m) {
        gn[m] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t_ < len[m]) gn[m] = __ldcs(reinterpret_cast<const float4*>(p.xp + ((size_t)t_ * p.rs_seq + rowb[m]) * p.gld + gcol));
      }
    };
    if (!XF) fetch_g(0);

    for (int s = 0; s < p.n_seq; ++s) {
      const int t = dir == 0 ? s : p.n_seq - 1 - s;
      float4 g[CPT];                               // pre-activations of this thread's cells (i,f,g,o)
#pragma unroll
      for (int m = 0; m < CPT; ++m) g[m] = XF ? ((t < len[m]) ? bias4 : make_float4(0.f, 0.f, 0.f, 0.f)) : gn[m];
      if (!XF && s + 1 < p.n_seq) fetch_g(s + 1);
      if (s > 0) mbar_wait_t(img_free, (s - 1) & 1);     // the previous image has been read out: checked off the dependent chain
      if (XF || s > 0) {
        uint32_t v[4 * CPT];
        mbar_wait_t(mma_done, XF ? (s & 1) : ((s - 1) & 1));
        if (threadIdx.x == 64) CL_STAMP(3);
        tc_fence_after();
        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * R + cgp * (R / 2));
#pragma unroll
        for (int rep = 0; rep < CPT / 2; ++rep) {
          tmem_ld_16x256(ta + 8 * rep, v + 8 * rep);                    // gates i, f of unit uq, batch columns 2 gp, 2 gp + 1
          tmem_ld_16x256(ta + (16u << 16) + 8 * rep, v + 8 * rep + 4);  // gates g, o
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_free);
#pragma unroll
        for (int m = 0; m < CPT; ++m) {
          const int o = 8 * (m >> 1) + (m & 1);
          if (t < len[m]) {
            g[m].x += __uint_as_float(v[o]); g[m].y += __uint_as_float(v[o + 2]);
            g[m].z += __uint_as_float(v[o + 4]); g[m].w += __uint_as_float(v[o + 6]);
          }
        }
      }
      float hv[CPT], cv[CPT];
#pragma unroll
      for (int m = 0; m < CPT; ++m) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        hv[m] = 0.f; cv[m] = 0.f;
        if (t < len[m]) {
          if (X3) {
            a.x = sigmoid_x(g[m].x); a.y = sigmoid_x(g[m].y); a.z = tanh_x(g[m].z); a.w = sigmoid_x(g[m].w);
            cv[m] = a.y * creg[m] + a.x * a.z;
            hv[m] = a.w * tanh_x(cv[m]);
          } else {
            a.x = sigmoid_apx(g[m].x); a.y = sigmoid_apx(g[m].y); a.z = tanh_apx(g[m].z); a.w = sigmoid_apx(g[m].w);
            cv[m] = fmaf(a.y, creg[m], a.x * a.z);
            hv[m] = a.w * tanh_apx(cv[m]);
          }
        }
        creg[m] = cv[m];
        g[m] = a;
      }
#pragma unroll
      for (int m = 0; m < CPT; ++m) {
        const int r = cgp * (R / 2) + 8 * (m >> 1) + 2 * gp + (m & 1);
        const __nv_bfloat16 hi = __float2bfloat16_rn(hv[m]);
        uint8_t* dst = img_u0 + (s & 1) * NC * IMGX + r * 128 + (((ul >> 3) ^ (r & 7)) << 4);
        *reinterpret_cast<__nv_bfloat16*>(dst) = hi;
        if (X3) *reinterpret_cast<__nv_bfloat16*>(dst + IMG) = __float2bfloat16_rn(hv[m] - __bfloat162float(hi));
      }
      fence_proxy_async();                         // generic-proxy smem writes -> visible to the bulk (async proxy) copies
      __syncwarp();
      if (lane == 0) mbar_arrive(stage_ready);
      if (threadIdx.x == 64) CL_STAMP(5);
      // ---- off the dependent chain: saved tensors ----
#pragma unroll
      for (int m = 0; m < CPT; ++m) {
        if (inr[m]) {
          const size_t row = (size_t)t * p.rs_seq + rowb[m];
          p.hout[row * p.hld + hcol] = hv[m];
          if (!X3) {
            __stcs(reinterpret_cast<float4*>(p.xp + row * p.gld + gcol), g[m]);
            p.cbuf[row * p.hld + hcol] = cv[m];
          }
        }
      }
      if (threadIdx.x == 64) CL_STAMP(7);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tmem);
  cluster_sync_all();                            // no CTA leaves while a peer could still address its shared memory
}

static long long* g_cl_dbg = nullptr;

template <typename Kern, typename... Args>
static int cluster_launch_t(int threads, Kern kern, dim3 grid, int cluster_x, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  return 0;
}

template <typename Kern, typename... Args>
static int cluster_launch(Kern kern, dim3 grid, int cluster_x, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  return 0;
}

// quad-cluster kernels (rec_q_*): S in {128, 256}; batch rows per tile from SSASR_REC_Q_ROWS / SSASR_REC_Q_ROWS_FWD
// (16 default, 32; 0 = the 8-CTA kernels above)
static int g_q_rows = -1, g_q_rows_fwd = -1;
static int parse_rows(const char* name, int dflt) {
  const char* e = getenv(name);
  int r = e ? atoi(e) : dflt;
  if (r != 0 && r != 16 && r != 32) r = dflt;
  return r;
}
static int q_rows() {
  if (g_q_rows < 0) g_q_rows = parse_rows("SSASR_REC_Q_ROWS", 16);
  return g_q_rows;
}
static int q_rows_fwd() {
  if (g_q_rows_fwd < 0) g_q_rows_fwd = parse_rows("SSASR_REC_Q_ROWS_FWD", q_rows());
  return g_q_rows_fwd;
}

static size_t fwd_smem(int S, bool xf = false) {
  return (size_t)(S / 64) * 16384 + (size_t)2 * CL_TM * S * 2 + FW_IMG + 8 * FW_WSTG + 8 * 8 + 16 + 1024 +
         (xf ? (size_t)FW_MAXKX * 4096 + 2 * FW_MAXKX * CL_TM * 32 + 512 : 0);
}
static size_t bwd_smem(int S) {
  return (size_t)CL_TM * 4 * S * 2 + (size_t)(S / 16) * 4096 + 8 * 4096 + (6 + BW_MAXNC) * 8 + 16 + 1024;
}

template <typename Kern>
static int prepare(Kern kern, int cluster_x, size_t smem, size_t smem_max, int* max_clusters) {
  // the attribute is per kernel, not per launch: always allow the largest supported state size
  SSASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster_x, 2, 1);
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int nc = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); nc = 0; }
  *max_clusters = nc;
  return 0;
}

// co-resident cluster capacity per state size, queried once (index = S / 64): [0] forward, [1] backward
static int g_cap[2][5] = {{-1, -1, -1, -1, -1}, {-1, -1, -1, -1, -1}};

static int g_cl_on = -1;      // -1: not decided yet (environment SSASR_REC_CLUSTER=0 disables), 0 / 1: set
static bool cl_enabled() {
  if (g_cl_on < 0) {
    const char* e = getenv("SSASR_REC_CLUSTER");
    g_cl_on = (e && e[0] == '0') ? 0 : 1;
  }
  return g_cl_on == 1;
}

// exchange ring: internal, L2-resident scratch (a few MB), one per stream that ever launched a cluster recurrence
constexpr size_t RING_BYTES = (size_t)CL_RING * 148 * BW_IMG;
struct RingSlot { cudaStream_t st; int dev; uint8_t* buf; };
static RingSlot g_rings[16];
static int g_nrings = 0;

static std::mutex g_ring_mu;       // the pool is shared by every host thread that launches a cluster recurrence
static uint8_t* ring_for(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_ring_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < g_nrings; ++i)
    if (g_rings[i].st == st && g_rings[i].dev == dev) return g_rings[i].buf;
  if (g_nrings == 16) {
    // pool full (17 distinct streams): hand the oldest entry of this device to the new stream once nothing can still be using it
    for (int i = 0; i < g_nrings; ++i)
      if (g_rings[i].dev == dev) {
        if (cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); return nullptr; }
        RingSlot r = g_rings[i];
        for (int j = i; j + 1 < g_nrings; ++j) g_rings[j] = g_rings[j + 1];
        r.st = st;
        g_rings[g_nrings - 1] = r;
        return r.buf;
      }
    return nullptr;
  }
  uint8_t* b = nullptr;
  if (cudaMalloc(&b, RING_BYTES) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  g_rings[g_nrings++] = {st, dev, b};
  return b;
}

}  // namespace

// 1 when the cluster kernels can run this layer with every (direction, 64-row tile) cluster co-resident
int rec_cl_supported(int S, int n_batch, int backward) {
  if (!cl_enabled()) return 0;
  if (!(S == 64 || S == 128 || S == 256)) return 0;
  const int idx = S / 64, w = backward ? 1 : 0;
  if (g_cap[w][idx] < 0) {
    int nc = 0;
    const int rc = backward ? prepare(rec_cl_bwd_kernel, S / CL_UNITS, bwd_smem(S), bwd_smem(256), &nc)
                            : prepare(rec_cl_fwd_kernel<false>, S / CL_UNITS, fwd_smem(S), fwd_smem(256), &nc);
    if (!backward && !rc)
      SSASR_CHECK_CUDA(cudaFuncSetAttribute(rec_cl_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem(256, true)));
    g_cap[w][idx] = rc ? 0 : nc;
  }
  const int tiles = (n_batch + CL_TM - 1) / CL_TM;
  return (2 * tiles <= g_cap[w][idx] && 2 * tiles * (S / CL_UNITS) <= 148) ? 1 : 0;
}

void rec_cl_enable(int on) { g_cl_on = on ? 1 : 0; }
int rec_cl_is_enabled() { return cl_enabled() ? 1 : 0; }

// co-resident cluster capacity (0 = cluster kernels unavailable for this state size)
int rec_cl_capacity(int S, int backward) {
  if (!(S == 64 || S == 128 || S == 256)) return 0;
  rec_cl_supported(S, 1, backward);
  return g_cap[backward ? 1 : 0][S / 64];
}

// x_bf / wih_bf / bias non-null: the input projection is fused in (x_bf [rows, Kp] bf16, wih_bf [8S, Kp] bf16, bias [8S]);
// xp then is an output only (saved activations).  Otherwise xp holds the pre-activations computed by the batched GEMM.
static size_t fwdq_smem(int S, int R, bool xf, int parts = 1) {
  return (size_t)(S / 64) * 32768 + (size_t)2 * (S / Q_UNITS) * R * 128 * parts + (R * 128 < 1024 ? 1024 : R * 128) + 10 * 8 + 16 + 1024 +
         (xf ? (size_t)FW_MAXKX * 256 * 32 + 2 * FW_MAXKX * R * 32 + 1024 : 0);
}

template <int R, bool XF>
static int rec_q_fwd_launch(cudaStream_t st, RecClParams& p, const void* whh_bf, void* hb, const void* x_bf, int Kp, const void* wih_bf,
                            int ndir = 2) {
  const int S = p.S;
  p.gld = (long long)ndir * 4 * S;
  p.hld = (long long)ndir * S;
  static bool attr_set = false;
  if (!attr_set) {
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(rec_q_fwd_kernel<R, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwdq_smem(256, R, XF)));
    attr_set = true;
  }
  CUtensorMap tmW, tmX, tmWx, tmH;
  p.whh = (const __nv_bfloat16*)whh_bf;
  p.wih = (const __nv_bfloat16*)wih_bf;
  p.kp = Kp;
  p.dsmem = rec_dsmem_enabled();
  int rc = make_tmap_bf16(&tmW, whh_bf, ndir * 4 * S, S, S, 256);
  if (rc) return rc;
  const bool si = p.rs_seq < p.rs_batch;
  rc = si ? make_tmap_bf16_3d_ex(&tmH, hb, ndir * S, p.n_seq, p.rs_seq * ndir * S, p.n_batch, p.rs_batch * ndir * S, 64, 1, R, 128)
          : make_tmap_bf16_3d_ex(&tmH, hb, ndir * S, p.n_batch, p.rs_batch * ndir * S, p.n_seq, p.rs_seq * ndir * S, 64, R, 1, 128);
  if (rc) return rc;
  tmX = tmW; tmWx = tmW;
  if (XF) {
    rc = si ? make_tmap_bf16_3d_ex(&tmX, x_bf, Kp, p.n_seq, p.rs_seq * Kp, p.n_batch, p.rs_batch * Kp, 16, 1, R, 32)
            : make_tmap_bf16_3d_ex(&tmX, x_bf, Kp, p.n_batch, p.rs_batch * Kp, p.n_seq, p.rs_seq * Kp, 16, R, 1, 32);
    if (rc) return rc;
    rc = make_tmap_bf16_3d_ex(&tmWx, wih_bf, Kp, 8 * S, Kp, 1, (long long)8 * S * Kp, 16, 256, 1, 32);
    if (rc) return rc;
  }
  dim3 grid(S / Q_UNITS, ndir, (p.n_batch + R - 1) / R);
  ProfScope ps(F_REC_TC_FWD, st);
  return cluster_launch_t(Q_THREADS, rec_q_fwd_kernel<R, XF>, grid, S / Q_UNITS, fwdq_smem(S, R, XF), st, tmW, tmX, tmWx, tmH, p);
}

// x_bf / wih_bf / bias non-null: the input projection is fused in (x_bf [rows, Kp] bf16, wih_bf [8S, Kp] bf16, bias [8S]);
// xp then is an output only (saved activations).  Otherwise xp holds the pre-activations computed by the batched GEMM.
int rec_cl_fwd(cudaStream_t st, float* xp, const void* whh_bf, float* hout, float* cbuf, void* hb, const int* lens, int S,
               int n_seq, int n_batch, long long rs_seq, long long rs_batch, const void* x_bf, int Kp, const void* wih_bf,
               const float* bias) {
  RecClParams p = {};
  p.xp = xp; p.hout = hout; p.cbuf = cbuf; p.xb = (__nv_bfloat16*)hb; p.lens = lens;
  p.S = S; p.n_seq = n_seq; p.n_batch = n_batch; p.rs_seq = rs_seq; p.rs_batch = rs_batch;
  p.dbg = g_cl_dbg;
  p.ring = ring_for(st);
  SSASR_REQUIRE(p.ring != nullptr, "rec_cl_fwd: cannot allocate the exchange ring");
  const bool xf = x_bf && wih_bf && bias;
  if (xf) {
    SSASR_REQUIRE(Kp % 8 == 0 && Kp > 0 && Kp <= 16 * FW_MAXKX, "rec_cl_fwd: fused input projection needs Kp %% 8 == 0 and Kp <= %d (got %d)",
                  16 * FW_MAXKX, Kp);
    p.bias = bias;
    p.nkx = (Kp + 15) / 16;
    p.seq_inner = rs_seq < rs_batch ? 1 : 0;
  }
  const int R = q_rows_fwd();
  if (R && (S == 128 || S == 256) && 2 * ((n_batch + R - 1) / R) * (S / Q_UNITS) <= 148) {
    if (R == 32) return xf ? rec_q_fwd_launch<32, true>(st, p, whh_bf, hb, x_bf, Kp, wih_bf) : rec_q_fwd_launch<32, false>(st, p, whh_bf, hb, x_bf, Kp, wih_bf);
    return xf ? rec_q_fwd_launch<16, true>(st, p, whh_bf, hb, x_bf, Kp, wih_bf) : rec_q_fwd_launch<16, false>(st, p, whh_bf, hb, x_bf, Kp, wih_bf);
  }
  CUtensorMap tmW, tmX, tmWx;
  int rc = make_tmap_bf16(&tmW, whh_bf, 8 * S, S, S, 128);
  if (rc) return rc;
  dim3 grid(S / CL_UNITS, 2, (n_batch + CL_TM - 1) / CL_TM);
  if (!xf) {
    ProfScope ps(F_REC_TC_FWD, st);
    return cluster_launch(rec_cl_fwd_kernel<false>, grid, S / CL_UNITS, fwd_smem(S), st, tmW, tmW, tmW, p);
  }
  rc = p.seq_inner ? make_tmap_bf16_3d_ex(&tmX, x_bf, Kp, n_seq, rs_seq * Kp, n_batch, rs_batch * Kp, 16, 1, CL_TM, 32)
                   : make_tmap_bf16_3d_ex(&tmX, x_bf, Kp, n_batch, rs_batch * Kp, n_seq, rs_seq * Kp, 16, CL_TM, 1, 32);
  if (rc) return rc;
  rc = make_tmap_bf16_3d_ex(&tmWx, wih_bf, Kp, 8 * S, Kp, 1, (long long)8 * S * Kp, 16, 128, 1, 32);
  if (rc) return rc;
  ProfScope ps(F_REC_TC_FWD, st);
  return cluster_launch(rec_cl_fwd_kernel<true>, grid, S / CL_UNITS, fwd_smem(S, true), st, tmW, tmX, tmWx, p);
}

// Exact (split-operand) forward recurrence on the quad clusters: xp [rows, 8S] fp32 pre-activations (split-operand GEMM),
// whh_hi / whh_lo [8S, S] bf16 parts of the packed W_hh; writes hout only.  Clusters are independent of one another, so the
// grid may exceed what is co-resident (it then runs in waves).  Returns -1 where the geometry is not covered.
template <int R>
static int rec_q_fwd_x3_launch(cudaStream_t st, RecClParams& p) {
  const int S = p.S;
  p.gld = (long long)8 * S;
  p.hld = (long long)2 * S;
  const size_t smem = fwdq_smem(S, R, false, 2);
  static bool attr_set = false;
  if (!attr_set) {
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(rec_q_fwd_kernel<R, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)fwdq_smem(256, R, false, 2)));
    attr_set = true;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));                    // the exact path uses no tensor map (operands by DSMEM / plain loads)
  dim3 grid(S / Q_UNITS, 2, (p.n_batch + R - 1) / R);
  const size_t n_cta = (size_t)grid.x * grid.y * grid.z;
  p.dsmem = (rec_dsmem_enabled() || (size_t)CL_RING * n_cta * 2 * R * 128 > RING_BYTES) ? 1 : 0;
  ProfScope ps(F_REC_TC_FWD, st);
  return cluster_launch_t(Q_THREADS, rec_q_fwd_kernel<R, false, true>, grid, S / Q_UNITS, smem, st, tm, tm, tm, tm, p);
}

int rec_q_fwd_x3(cudaStream_t st, float* xp, const void* whh_hi, const void* whh_lo, float* hout, const int* lens, int S, int n_seq,
                 int n_batch, long long rs_seq, long long rs_batch) {
  if (!cl_enabled() || !(S == 128 || S == 256) || n_seq < 4) return -1;
  const char* e = getenv("SSASR_REC_Q_X3");
  if (e && e[0] == '0') return -1;
  RecClParams p = {};
  p.xp = xp; p.hout = hout; p.lens = lens;
  p.S = S; p.n_seq = n_seq; p.n_batch = n_batch; p.rs_seq = rs_seq; p.rs_batch = rs_batch;
  p.whh = (const __nv_bfloat16*)whh_hi; p.whh_lo = (const __nv_bfloat16*)whh_lo;
  p.dbg = g_cl_dbg;
  p.ring = ring_for(st);
  SSASR_REQUIRE(p.ring != nullptr, "rec_q_fwd_x3: cannot allocate the exchange ring");
  // 16-row tiles while all their clusters are co-resident (small batches are bound by the step latency: 125 utterances 14.3 -> 13.3 ms
  // per decode), 32-row tiles beyond (the MMAs sit at the N <= 32 issue floor either way: twice the rows per step)
  const int R = e ? (atoi(e) == 16 ? 16 : 32) : (2 * ((n_batch + 15) / 16) * (S / Q_UNITS) <= 148 ? 16 : 32);
  return R == 16 ? rec_q_fwd_x3_launch<16>(st, p) : rec_q_fwd_x3_launch<32>(st, p);
}

// largest padded input width the fused forward projection accepts
int rec_cl_fused_kp_max() { return 16 * FW_MAXKX; }

static size_t bwdq_smem(int S, int R) { return (size_t)8 * R * S + (size_t)512 * S + (6 + Q_MAXNC) * 8 + 16 + 1024; }

template <int R>
static int rec_q_bwd_launch(cudaStream_t st, RecClParams& p, const void* whhT_bf, void* dgb, int ndir = 2) {
  const int S = p.S;
  p.gld = (long long)ndir * 4 * S;
  p.hld = (long long)ndir * S;
  static bool attr_set = false;
  if (!attr_set) {
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(rec_q_bwd_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwdq_smem(256, R)));
    attr_set = true;
  }
  CUtensorMap tmW, tmG;
  int rc = make_tmap_bf16(&tmW, whhT_bf, ndir * S, 4 * S, 4 * S, Q_UNITS);
  if (rc) return rc;
  const long long gl = (long long)ndir * 4 * S;
  rc = p.rs_seq < p.rs_batch ? make_tmap_bf16_3d_ex(&tmG, dgb, gl, p.n_seq, p.rs_seq * gl, p.n_batch, p.rs_batch * gl, 64, 1, R, 128)
                             : make_tmap_bf16_3d_ex(&tmG, dgb, gl, p.n_batch, p.rs_batch * gl, p.n_seq, p.rs_seq * gl, 64, R, 1, 128);
  if (rc) return rc;
  dim3 grid(S / Q_UNITS, ndir, (p.n_batch + R - 1) / R);
  ProfScope ps(F_REC_TC_BWD, st);
  return cluster_launch_t(Q_THREADS, rec_q_bwd_kernel<R>, grid, S / Q_UNITS, bwdq_smem(S, R), st, tmW, tmG, p);
}

int rec_cl_bwd(cudaStream_t st, float* act, const void* whhT_bf, const float* cbuf, const float* dhout, void* dgb, const int* lens,
               int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch, float* dbias) {
  RecClParams p = {};
  p.xp = act; p.cbuf = const_cast<float*>(cbuf); p.xb = (__nv_bfloat16*)dgb; p.dhout = dhout; p.dbias = dbias; p.lens = lens;
  p.S = S; p.n_seq = n_seq; p.n_batch = n_batch; p.rs_seq = rs_seq; p.rs_batch = rs_batch;
  p.dbg = g_cl_dbg;
  p.ring = ring_for(st);
  SSASR_REQUIRE(p.ring != nullptr, "rec_cl_bwd: cannot allocate the exchange ring");
  const int R = q_rows();
  if (R && (S == 128 || S == 256) && 2 * ((n_batch + R - 1) / R) * (S / Q_UNITS) <= 148)
    return R == 32 ? rec_q_bwd_launch<32>(st, p, whhT_bf, dgb) : rec_q_bwd_launch<16>(st, p, whhT_bf, dgb);
  CUtensorMap tmW, tmG;
  int rc = make_tmap_bf16(&tmW, whhT_bf, 2 * S, 4 * S, 4 * S, 32);
  if (rc) return rc;
  rc = rs_seq < rs_batch ? make_tmap_bf16_3d_ex(&tmG, dgb, 8 * S, n_seq, rs_seq * 8 * S, n_batch, rs_batch * 8 * S, 64, 1, CL_TM, 128)
                         : make_tmap_bf16_3d_ex(&tmG, dgb, 8 * S, n_batch, rs_batch * 8 * S, n_seq, rs_seq * 8 * S, 64, CL_TM, 1, 128);
  if (rc) return rc;
  dim3 grid(S / CL_UNITS, 2, (n_batch + CL_TM - 1) / CL_TM);
  ProfScope ps(F_REC_TC_BWD, st);
  return cluster_launch(rec_cl_bwd_kernel, grid, S / CL_UNITS, bwd_smem(S), st, tmW, tmG, p);
}

void rec_q_set_rows(int rows) { g_q_rows = g_q_rows_fwd = (rows == 16 || rows == 32) ? rows : 0; }

}  // namespace ssasr

extern "C" {
// debug: device buffer [n_seq][12] of clock64 stamps written by CTA (0,0,0) of the next cluster recurrent launches
void ssasr_rec_cl_set_debug(long long* dev_buf) { ssasr::g_cl_dbg = dev_buf; }
// how many (direction, 64-row tile) clusters of the cluster recurrent kernels can be co-resident (0: not available)
int ssasr_rec_cl_capacity(int S, int backward) { return ssasr::rec_cl_capacity(S, backward); }
// 0: use the counter-barrier kernels of rec_tc.cu even where the cluster kernels apply (A/B comparison in tests and scripts)
void ssasr_rec_cl_enable(int on) { ssasr::rec_cl_enable(on); }
// batch rows per tile of the quad-cluster recurrent kernels (64 units per CTA): 32 (default) or 16; 0 selects the older
// 8-CTA / 64-row kernels (A/B comparison in tests and scripts)
void ssasr_rec_q_set_rows(int rows) { ssasr::rec_q_set_rows(rows); }
}

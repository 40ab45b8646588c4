// Cluster recurrent BLSTM kernels for the WIDE state size (S = 512: the long-utterance configuration, BASELINE configs[4]).
// The quad-cluster kernels of rec_cl.cu keep a 256-gate-row slice of W_hh resident per CTA, which at S = 512 is 256 KB; here
// a (direction, 16-utterance tile) belongs to ONE 16-CTA thread-block cluster (non-portable size), CTA r owning hidden units
// [32 r, 32 r + 32) = 128 gate rows:
//   forward   gates^T [128 rows x 16 utt] = W_hh slice [128 x 512] (resident, 128 KB) x h(t-1)^T tile (all-gathered: every CTA
//             multicasts its 1 KB slice); cell math from TMEM, one cell per epilogue thread
//   backward  K-split, as in spell_cl.cu: every CTA contracts ITS 128 gate rows,
//             dh^T partial [512 units x 16 utt] = W_hh^T slice [512 x 128] (resident, 128 KB) x dG slice^T,
//             and pushes the bf16 partial of each 32-unit block to the block's owner (bulk store + read-back with a one-CTA
//             multicast mask); the owner adds the 16 partials to dh_out(t) and runs the cell backward
// so neither direction streams weights and the per-step MMA count is 32 (N = 16) instead of 128 on the all-gathered dG tile.
//
// Reference semantics: nn.LSTM(bidirectional) over a packed batch (asr.py:410-418): masking and row-stride conventions, bf16
// operand rounding, fp32 accumulation and tanh.approx cell math exactly as in rec_cl.cu / rec_tc.cu.
#include <limits.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"
#include "cl_common.cuh"

namespace ssasr {

using namespace tc;
using namespace clx;

int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows);
int rec_cl_is_enabled();      // rec_cl.cu: SSASR_REC_CLUSTER / ssasr_rec_cl_enable (A/B switch of all cluster recurrent kernels)

namespace {

constexpr int RW_S = 512;
constexpr int RW_NC = RW_S / 32;          // 16 CTAs per cluster, 32 hidden units each
constexpr int RW_NT = 16;                 // utterances per tile
constexpr int RW_EPW = 16;
constexpr int RW_THREADS = 128 + 32 * RW_EPW;     // warp 0 exchange, 1 MMA / TMEM, 2-3 idle, 4.. compute
constexpr int RW_HBLK = RW_NT * 64;       // one producer's k-block of the h tile (32 units): [16 rows x 64 B], 64-byte swizzle
constexpr int RW_RING = 4;
constexpr int RW_SLOT = 16384;            // exchange slot per CTA and ring position (backward: up to 16 KB; forward uses 1 KB of it)

constexpr int FOFF_W = 0;                                 // W_hh slice [128 gate rows x 512]: 8 k-blocks of [128 x 128 B]
constexpr int FOFF_H = FOFF_W + 131072;                   // h tile [2 buffers][16 producers][HBLK]
constexpr int FOFF_HIMG = FOFF_H + 2 * RW_NC * RW_HBLK;
constexpr int FOFF_BARS = FOFF_HIMG + RW_HBLK;
constexpr int RWF_SMEM = FOFF_BARS + 16 * 8 + 1024;
constexpr int RWF_SMEM_X3 = FOFF_H + 2 * (2 * RW_NC * RW_HBLK) + 2 * RW_HBLK + 16 * 8 + 1024;   // hi + lo h images
static_assert(RWF_SMEM_X3 <= 232448, "shared-memory map (exact path)");

static_assert(RWF_SMEM <= 232448, "shared-memory map");

struct RecWideP {
  float* xp;                 // fwd: [rows, 8S] pre-activations in / activations out.  bwd: activations in
  float* hout;               // fwd out [rows, 2S]
  float* cbuf;               // fwd out / bwd in [rows, 2S]
  __nv_bfloat16* xb;         // fwd out: bf16 h [rows, 2S].  bwd out: bf16 dG [rows, 8S]
  const float* dhout;        // bwd in [rows, 2S]
  float* dbias;              // bwd: [8S] pre-zeroed, atomically accumulated; may be null
  uint8_t* ring;
  const int* lens;
  int n_seq, n_batch;
  long long rs_seq, rs_batch;
  long long* dbg;            // optional [n_seq][12] clock64 stamps of CTA (0,0,0)
  const __nv_bfloat16* w;    // fwd: packed W_hh [8S, S]; bwd: W_hh^T [2S, 4S] (bf16, row-major): the resident slice -> TENSOR memory
  int dsmem;                 // bwd: partial sums go to their owners by DSMEM bulk copies instead of through the L2 ring
};
constexpr int RW_WCOL = 64;  // TMEM: accumulators in columns [0, 64), the resident weight slice (A operand) from column 64
#define RW_STAMP(idx)                                                                                            \
  do {                                                                                                           \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg[(size_t)s * 12 + (idx)] = clock64(); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// X3 (forward-only exact path, greedy decoding at S = 512): W_hh and h split into bf16 hi + lo parts, product
// W_hi h_hi + W_hi h_lo + W_lo h_hi; W_hi is the TMEM-resident A operand, W_lo the shared-memory copy (tmW then maps the LOW
// parts), a producer's image carries its hi and lo k-block; ex2 / rcp based activations with 1e-7 absolute error (cl_common.cuh); only `hout` is written.
template <bool X3>
__global__ void __launch_bounds__(RW_THREADS, 1) rec_wide_fwd_kernel(const __grid_constant__ CUtensorMap tmW, RecWideP p) {
  constexpr int S = RW_S;
  constexpr int HB = (X3 ? 2 : 1) * RW_HBLK;       // one producer's image: its k-block of every operand part
  constexpr int SLOT = X3 ? HB : RW_SLOT;          // exchange slot stride (the exact path packs them: grids beyond 160 CTAs)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Wsm = smem + FOFF_W;
  uint8_t* Hsm = smem + FOFF_H;                    // [2 buffers][16 producers][parts][HBLK]
  uint8_t* himg = Hsm + 2 * RW_NC * HB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(himg + HB);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;            // [2]
  uint64_t* g_done = bars + 3;
  uint64_t* stage_ready = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int r = blockIdx.x, dir = blockIdx.y, b0 = blockIdx.z * RW_NT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_steps = p.n_seq;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1);
    mbar_init(a_full, 1);
    mbar_init(a_full + 1, 1);
    mbar_init(g_done, 1);
    mbar_init(stage_ready, RW_EPW);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(w_full, 131072);
    for (int kb = 0; kb < 8; ++kb) tma_load_2d(&tmW, w_full, Wsm + kb * 16384, kb * 64, dir * 4 * S + r * 128);
    mbar_expect_tx(a_full, RW_NC * HB);
    mbar_expect_tx(a_full + 1, RW_NC * HB);
  }
  if (warp >= 4) {
    // resident W_hh slice [128 gate rows x S] -> tensor memory: thread = one gate row, the four warps of a sub-partition
    // split K into quarters (S / 8 packed words each)
    const int sp_ = warp & 3, part_ = (warp - 4) >> 2;
    const __nv_bfloat16* wrow = p.w + ((size_t)dir * 4 * S + r * 128 + sp_ * 32 + lane) * S + part_ * (S / 4);
    tmem_store_row(tmem + ((uint32_t)(sp_ * 32) << 16) + (uint32_t)(RW_WCOL + part_ * (S / 8)), wrow, S / 8);
    tc_fence_before();
  }
  cluster_sync_all();

  if (warp == 0) {
    if (elect_one()) {
      const uint16_t cmask = 0xFFFFu;
      const size_t n_cta = (size_t)gridDim.x * gridDim.y * gridDim.z;
      const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      for (int s = 0; s + 1 < n_steps; ++s) {
        uint8_t* slot = p.ring + ((size_t)(s % RW_RING) * n_cta + cta) * SLOT;
        mbar_wait_t(stage_ready, s & 1);
        bulk_store_wait(slot, himg, HB);
        bulk_load_mc(Hsm + ((s & 1) * RW_NC + r) * HB, slot, HB, a_full + (s & 1), cmask);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, RW_NT);
      mbar_wait_t(w_full, 0);
      for (int s = 1; s < n_steps; ++s) {                  // step 0 starts from the zero state: no product
        const int hb = (s - 1) & 1;
        mbar_wait_t(a_full + hb, ((s - 1) >> 1) & 1);
        if (s + 2 < n_steps) mbar_expect_tx(a_full + hb, RW_NC * HB);
        tc_fence_after();
        const uint64_t dhb = umma_desc_k64(smem_u32(Hsm + hb * RW_NC * HB));
        const uint64_t dWl = umma_desc_k128(smem_u32(Wsm));
#pragma unroll
        for (int kk = 0; kk < S / 16; ++kk) {
          const uint64_t dh = dhb + (uint64_t)(((kk >> 1) * HB) >> 4) + (uint64_t)((kk & 1) * 2);
          mma_bf16_ts(tmem, tmem + RW_WCOL + kk * 8, dh, idesc, kk != 0);
          if (X3) {
            mma_bf16_ts(tmem, tmem + RW_WCOL + kk * 8, dh + (uint64_t)(RW_HBLK >> 4), idesc, 1u);
            mma_bf16_ss(tmem, dWl + (uint64_t)(((kk >> 2) * 16384) >> 4) + (uint64_t)((kk & 3) * 2), dh, idesc, 1u);
          }
        }
        mma_commit(g_done);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int sp = warp & 3, cg = (warp - 4) >> 2;
    const int uq = lane >> 2, gp = lane & 3;
    const int unit = 32 * r + 8 * sp + uq;         // hidden unit (inside the direction) of this thread's cell
    const int slot = 4 * cg + gp;                  // utterance slot of the tile
    const int n = b0 + slot;
    const bool inr = n < p.n_batch;
    const int len = inr ? (p.lens ? p.lens[n] : INT_MAX) : 0;
    const size_t rowb = (size_t)(inr ? n : 0) * p.rs_batch;
    const size_t gcol = (size_t)dir * 4 * S + (size_t)unit * 4, hcol = (size_t)dir * S + unit;
    uint8_t* himg_dst = himg + slot * 64 + ((sp ^ ((slot >> 1) & 3)) << 4) + uq * 2;
    float creg = 0.f;
    float4 gn = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](int s_) {
      const int t_ = dir == 0 ? s_ : n_steps - 1 - s_;
      gn = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t_ < len) gn = __ldcs(reinterpret_cast<const float4*>(p.xp + ((size_t)t_ * p.rs_seq + rowb) * 8 * S + gcol));
    };
    fetch(0);
    for (int s = 0; s < n_steps; ++s) {
      const int t = dir == 0 ? s : n_steps - 1 - s;
      const bool valid = t < len;
      float4 g = gn;
      if (s + 1 < n_steps) fetch(s + 1);
      if (s > 0) {
        uint32_t v[4];
        mbar_wait_t(g_done, (s - 1) & 1);
        tc_fence_after();
        tmem_ld4(tmem + ((uint32_t)(sp * 32) << 16) + (uint32_t)(4 * cg), v);
        tmem_ld_wait();
        tc_fence_before();
        float a0 = __uint_as_float(v[0]), a1 = __uint_as_float(v[1]), a2 = __uint_as_float(v[2]), a3 = __uint_as_float(v[3]);
        quad_transpose(a0, a1, a2, a3, gp);
        g.x += a0; g.y += a1; g.z += a2; g.w += a3;
      }
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      float cv = 0.f, hv = 0.f;
      if (valid) {
        if (X3) {
          a.x = sigmoid_x(g.x); a.y = sigmoid_x(g.y); a.z = tanh_x(g.z); a.w = sigmoid_x(g.w);
          cv = a.y * creg + a.x * a.z;
          hv = a.w * tanh_x(cv);
        } else {
          a.x = sigmoid_apx(g.x); a.y = sigmoid_apx(g.y); a.z = tanh_apx(g.z); a.w = sigmoid_apx(g.w);
          cv = fmaf(a.y, creg, a.x * a.z);
          hv = a.w * tanh_apx(cv);
        }
      }
      creg = cv;
      const __nv_bfloat16 hhi = __float2bfloat16_rn(hv);
      *reinterpret_cast<__nv_bfloat16*>(himg_dst) = hhi;
      if (X3) *reinterpret_cast<__nv_bfloat16*>(himg_dst + RW_HBLK) = __float2bfloat16_rn(hv - __bfloat162float(hhi));
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(stage_ready);
      if (inr) {
        const size_t row = (size_t)t * p.rs_seq + rowb;
        p.hout[row * 2 * S + hcol] = hv;
        if (!X3) {
          __stcs(reinterpret_cast<float4*>(p.xp + row * 8 * S + gcol), a);
          p.cbuf[row * 2 * S + hcol] = cv;
          p.xb[row * 2 * S + hcol] = hhi;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
  cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------
// backward, K-split.  Template: state size S and hidden units per CTA (32: the 16-CTA clusters of S = 512; 64: the quad
// clusters of S = 256 / 128, same geometry as rec_q_bwd_kernel of rec_cl.cu, which all-gathers the dG tile and issues 4S/16
// MMAs per step -- here S/128 * UNITS/4 on the CTA's own dG slice).
// ------------------------------------------------------------------------------------------------
template <int S, int UNITS>
struct KsGeom {
  static constexpr int NC = S / UNITS;              // CTAs per cluster
  static constexpr int G = 4 * UNITS;               // gate rows per CTA
  static constexpr int KB = G / 64;                 // k-blocks of the dG slice operand
  static constexpr int NH = S / 128;                // 128-unit blocks of the accumulator
  static constexpr int CPT = UNITS / 32;            // cells per thread
  static constexpr int DHB = RW_NT * UNITS * 2;     // one (source, destination) block: [16 utterances][UNITS] bf16
  static constexpr int OFF_W = 0;                   // W_hh^T slice: [NH][KB] blocks of [128 units x 128 B]
  static constexpr int OFF_DG = OFF_W + NH * KB * 16384;
  static constexpr int OFF_DHOUT = OFF_DG + KB * 2048;
  static constexpr int OFF_DHIN = OFF_DHOUT + 2 * NC * DHB;   // outgoing partial sums double-buffered (DSMEM mode)
  static constexpr int OFF_BARS = OFF_DHIN + 2 * NC * DHB;
  static constexpr int SMEM = OFF_BARS + 16 * 8 + 1024;
  static constexpr int TCOLS = 512;                 // accumulators [0, NH * 16) + the W_hh^T slice from column RW_WCOL: NH x G/2 columns
  static constexpr int SLOT = NC * DHB;             // exchange slot per CTA and ring position
};

template <int S, int UNITS>
__global__ void __launch_bounds__(RW_THREADS, 1) rec_ks_bwd_kernel(const __grid_constant__ CUtensorMap tmWT, RecWideP p) {
  using GE = KsGeom<S, UNITS>;
  constexpr int NC = GE::NC, KB = GE::KB, NH = GE::NH, CPT = GE::CPT, DHB = GE::DHB;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Wsm = smem + GE::OFF_W;
  uint8_t* dGsm = smem + GE::OFF_DG;
  uint8_t* dhout = smem + GE::OFF_DHOUT;
  uint8_t* dhin = smem + GE::OFF_DHIN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GE::OFF_BARS);
  uint64_t* w_full = bars;
  uint64_t* dh_full = bars + 1;           // [2]
  uint64_t* d_done = bars + 3;
  uint64_t* dg_ready = bars + 4;
  uint64_t* dh_ready = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int r = blockIdx.x, dir = blockIdx.y, b0 = blockIdx.z * RW_NT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_steps = p.n_seq;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmWT);
    mbar_init(w_full, 1);
    mbar_init(dh_full, 1);
    mbar_init(dh_full + 1, 1);
    mbar_init(d_done, 1);
    mbar_init(dg_ready, RW_EPW);
    mbar_init(dh_ready, RW_EPW);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<GE::TCOLS>(tmem_slot);
  for (int i = threadIdx.x; i < KB * 2048 / 16; i += RW_THREADS) reinterpret_cast<uint4*>(dGsm)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(w_full, NH * KB * 16384);
    for (int h = 0; h < NH; ++h)
      for (int kb = 0; kb < KB; ++kb)
        tma_load_2d(&tmWT, w_full, Wsm + (h * KB + kb) * 16384, r * GE::G + kb * 64, dir * S + h * 128);
    mbar_expect_tx(dh_full, NC * DHB);
    mbar_expect_tx(dh_full + 1, NC * DHB);
  }
  if (warp >= 4) {
    // resident W_hh^T slice [S units x G own gate rows] -> tensor memory: thread = one unit row of a 128-unit block, the 16
    // warps = 4 sub-partitions x NH blocks x 4 / NH parts of K
    constexpr int PARTS = 4 / NH, PW = GE::G / 2 / PARTS;      // packed words per part
    const int sp_ = warp & 3, j_ = (warp - 4) >> 2, h_ = j_ / PARTS, part_ = j_ % PARTS;
    const __nv_bfloat16* wrow = p.w + ((size_t)dir * S + h_ * 128 + sp_ * 32 + lane) * (4 * S) + r * GE::G + part_ * (2 * PW);
    tmem_store_row(tmem + ((uint32_t)(sp_ * 32) << 16) + (uint32_t)(RW_WCOL + h_ * (GE::G / 2) + part_ * PW), wrow, PW);
    tc_fence_before();
  }
  cluster_sync_all();

  if (warp == 0) {
    if (elect_one()) {
      const size_t n_cta = (size_t)gridDim.x * gridDim.y * gridDim.z;
      const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      for (int s = 0; s + 1 < n_steps; ++s) {
        uint8_t* slot = p.ring + ((size_t)(s % RW_RING) * n_cta + cta) * GE::SLOT;
        const uint8_t* out = dhout + (s & 1) * NC * DHB;
        mbar_wait_t(dh_ready, s & 1);
        RW_STAMP(8);
        if (p.dsmem) {
          // shared -> shared of the owner CTA, completing its bytes on the owner's barrier.  The outgoing buffer of step s is
          // rewritten at step s + 2, after this CTA has received the peers' partial sums of step s + 1, which they computed
          // after consuming THESE copies
          // (one thread, one copy after the other: a lane per destination issuing them at once took 740 instead of 250 cycles);
          // the peers first, the CTA's own block last
          for (int k = 1; k <= NC; ++k) {
            const int d = (r + k) % NC;
            bulk_copy_dsmem(dhin + ((s & 1) * NC + r) * DHB, out + d * DHB, DHB, dh_full + (s & 1), (uint32_t)d);
          }
          RW_STAMP(10);
        } else {
          bulk_store_wait(slot, out, NC * DHB);
          RW_STAMP(9);
          for (int d = 0; d < NC; ++d)
            bulk_load_mc(dhin + ((s & 1) * NC + r) * DHB, slot + d * DHB, DHB, dh_full + (s & 1), (uint16_t)(1u << d));
          RW_STAMP(10);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, RW_NT);
      mbar_wait_t(w_full, 0);
      for (int s = 0; s + 1 < n_steps; ++s) {
        mbar_wait_t(dg_ready, s & 1);
        RW_STAMP(2);
        tc_fence_after();
        const uint32_t g0 = smem_u32(dGsm);
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
          for (int kk = 0; kk < 4 * KB; ++kk)
            mma_bf16_ts(tmem + h * RW_NT, tmem + RW_WCOL + h * (GE::G / 2) + kk * 8,
                        umma_desc_k128(g0 + (kk >> 2) * 2048) + (uint64_t)((kk & 3) * 2), idesc, kk != 0);
        mma_commit(d_done);
        RW_STAMP(3);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int cw = warp - 4;                       // utterance slot of this warp's cells; lane (+ 32) = unit of the CTA
    const int n = b0 + cw;
    const bool inr = n < p.n_batch;
    const int len = inr ? (p.lens ? p.lens[n] : INT_MAX) : 0;
    const size_t rowb = (size_t)(inr ? n : 0) * p.rs_batch;
    const size_t gcol0 = (size_t)dir * 4 * S + (size_t)(UNITS * r) * 4, hcol0 = (size_t)dir * S + UNITS * r;
    float dcreg[CPT], bsum[CPT][4];
    float4 a_n[CPT];
    float c_n[CPT], cp_n[CPT], dh_n[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) { dcreg[j] = 0.f; bsum[j][0] = bsum[j][1] = bsum[j][2] = bsum[j][3] = 0.f; }
    // the backward chain walks the steps in the opposite order of the forward pass of its direction
    auto fetch = [&](int s_) {
      const int t_ = dir == 0 ? n_steps - 1 - s_ : s_;
      const int tp_ = dir == 0 ? t_ - 1 : t_ + 1;
      const bool v = t_ < len, pv_ = v && tp_ >= 0 && tp_ < n_steps && tp_ < len;
      const size_t row = (size_t)t_ * p.rs_seq + rowb, rowp = (size_t)(pv_ ? tp_ : 0) * p.rs_seq + rowb;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int ul = lane + 32 * j;
        a_n[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        c_n[j] = 0.f; cp_n[j] = 0.f; dh_n[j] = 0.f;
        if (v) {
          a_n[j] = __ldcs(reinterpret_cast<const float4*>(p.xp + row * 8 * S + gcol0 + (size_t)ul * 4));
          dh_n[j] = __ldcs(p.dhout + row * 2 * S + hcol0 + ul);
          c_n[j] = __ldg(p.cbuf + row * 2 * S + hcol0 + ul);
          if (pv_) cp_n[j] = __ldg(p.cbuf + rowp * 2 * S + hcol0 + ul);
        }
      }
    };
    fetch(0);
    for (int s = 0; s < n_steps; ++s) {
      const int t = dir == 0 ? n_steps - 1 - s : s;
      const int tp = dir == 0 ? t - 1 : t + 1;
      const bool valid = t < len;
      const bool pv = valid && tp >= 0 && tp < n_steps && tp < len;
      float4 a[CPT];
      float cv[CPT], cpv[CPT], dh[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) { a[j] = a_n[j]; cv[j] = c_n[j]; cpv[j] = cp_n[j]; dh[j] = dh_n[j]; }
      if (s + 1 < n_steps) fetch(s + 1);
      // everything of the cell backward that does not need dh(t) is computed BEFORE the wait for the partial sums:
      //   dc = dh * k_dc + dc_carry,  dG_o = dh * k_o,  dG_i = dc * k_i,  dG_g = dc * k_g,  dG_f = dc * k_f,  carry' = dc * f
      float k_dc[CPT], k_o[CPT], k_i[CPT], k_g[CPT], k_f[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const float tc_ = tanh_apx(cv[j]);
        k_dc[j] = a[j].w * (1.f - tc_ * tc_);
        k_o[j] = tc_ * a[j].w * (1.f - a[j].w);
        k_i[j] = a[j].z * a[j].x * (1.f - a[j].x);
        k_g[j] = a[j].x * (1.f - a[j].z * a[j].z);
        k_f[j] = pv ? cpv[j] * a[j].y * (1.f - a[j].y) : 0.f;
      }
      if (s > 0) {
        mbar_wait_t(dh_full + ((s - 1) & 1), ((s - 1) >> 1) & 1);
        if (threadIdx.x == 128) RW_STAMP(1);
        if (threadIdx.x == 128 && s + 2 < n_steps) mbar_expect_tx(dh_full + ((s - 1) & 1), NC * DHB);
        const uint8_t* base = dhin + ((s - 1) & 1) * NC * DHB + cw * (UNITS * 2) + lane * 2;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int src = 0; src < NC; ++src) acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(base + src * DHB + j * 64));
          dh[j] += acc;
        }
      }
      uint2 pk[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
        float dco = 0.f;
        if (valid) {
          const float dc = fmaf(dh[j], k_dc[j], dcreg[j]);
          dg.w = dh[j] * k_o[j];
          dg.x = dc * k_i[j];
          dg.z = dc * k_g[j];
          dg.y = dc * k_f[j];
          dco = dc * a[j].y;
        }
        dcreg[j] = dco;
        bsum[j][0] += dg.x; bsum[j][1] += dg.y; bsum[j][2] += dg.z; bsum[j][3] += dg.w;
        const __nv_bfloat162 b01 = __floats2bfloat162_rn(dg.x, dg.y), b23 = __floats2bfloat162_rn(dg.z, dg.w);
        pk[j].x = *reinterpret_cast<const uint32_t*>(&b01);
        pk[j].y = *reinterpret_cast<const uint32_t*>(&b23);
        // operand tile: row = utterance slot, gate column 4 ul + g of the CTA's G: k-block ul / 16, 16-byte chunk (ul % 16) / 2
        const int ul = lane + 32 * j;
        *reinterpret_cast<uint2*>(dGsm + (ul >> 4) * 2048 + cw * 128 + ((((ul & 15) >> 1) ^ (cw & 7)) << 4) + (ul & 1) * 8) = pk[j];
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(dg_ready);
      if (threadIdx.x == 128) RW_STAMP(4);
      if (inr) {
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          *reinterpret_cast<uint2*>(p.xb + ((size_t)t * p.rs_seq + rowb) * 8 * S + gcol0 + (size_t)(lane + 32 * j) * 4) = pk[j];
      }
      if (s + 1 < n_steps) {
        mbar_wait_t(d_done, s & 1);
        if (threadIdx.x == 128) RW_STAMP(5);
        tc_fence_after();
        const int sp = warp & 3, cgrp = cw >> 2;   // TMEM sub-partition = 32 units; this warp converts 4 utterance columns
        uint32_t v[NH][4];
#pragma unroll
        for (int h = 0; h < NH; ++h) tmem_ld4(tmem + ((uint32_t)(sp * 32) << 16) + (uint32_t)(h * RW_NT + cgrp * 4), v[h]);
        tmem_ld_wait();
        if (threadIdx.x == 128) RW_STAMP(7);
        tc_fence_before();
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int unit = 128 * h + 32 * sp + lane;             // unit inside the direction
          uint8_t* dst = dhout + (s & 1) * NC * DHB + (unit / UNITS) * DHB + (unit % UNITS) * 2;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<__nv_bfloat16*>(dst + (cgrp * 4 + c) * (UNITS * 2)) = __float2bfloat16_rn(__uint_as_float(v[h][c]));
        }
        if (threadIdx.x == 128) RW_STAMP(11);
        fence_proxy_async();
        if (threadIdx.x == 128) RW_STAMP(0);
        __syncwarp();
        if (lane == 0) mbar_arrive(dh_ready);
        if (threadIdx.x == 128) RW_STAMP(6);
      }
    }
    if (p.dbias) {
#pragma unroll
      for (int j = 0; j < CPT; ++j)
#pragma unroll
        for (int g = 0; g < 4; ++g) atomicAdd(p.dbias + gcol0 + (size_t)(lane + 32 * j) * 4 + g, bsum[j][g]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<GE::TCOLS>(tmem);
  cluster_sync_all();
}

long long* g_rw_dbg = nullptr;
struct RwRing { cudaStream_t st; int dev; uint8_t* buf; };
RwRing g_rw_rings[16];
int g_rw_nrings = 0;
std::mutex g_rw_ring_mu;
uint8_t* rw_ring_for(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_rw_ring_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < g_rw_nrings; ++i)
    if (g_rw_rings[i].st == st && g_rw_rings[i].dev == dev) return g_rw_rings[i].buf;
  if (g_rw_nrings == 16) {
    // pool full: hand the oldest entry of this device to the new stream once nothing can still be using it
    for (int i = 0; i < g_rw_nrings; ++i)
      if (g_rw_rings[i].dev == dev) {
        if (cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); return nullptr; }
        RwRing r = g_rw_rings[i];
        for (int j = i; j + 1 < g_rw_nrings; ++j) g_rw_rings[j] = g_rw_rings[j + 1];
        r.st = st;
        g_rw_rings[g_rw_nrings - 1] = r;
        return r.buf;
      }
    return nullptr;
  }
  uint8_t* b = nullptr;
  if (cudaMalloc(&b, (size_t)RW_RING * 160 * RW_SLOT) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  g_rw_rings[g_rw_nrings++] = {st, dev, b};
  return b;
}

template <typename Kern>
int rw_query(Kern kern, int smem_bytes, int cluster) {
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (cluster > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster, 2, 1);
  cfg.blockDim = dim3(RW_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess) { cudaGetLastError(); nc = 0; }
  return nc;
}
// co-resident cluster capacity: [0] wide forward, [1] wide backward (S = 512), [2] / [3] K-split backward at S = 256 / 128
int g_rw_cap[4] = {-1, -1, -1, -1};
int rw_capacity(int which) {
  if (g_rw_cap[which] < 0) {
    switch (which) {
      case 0: g_rw_cap[0] = rw_query(rec_wide_fwd_kernel<false>, RWF_SMEM, RW_NC); break;
      case 1: g_rw_cap[1] = rw_query(rec_ks_bwd_kernel<512, 32>, KsGeom<512, 32>::SMEM, 16); break;
      case 2: g_rw_cap[2] = rw_query(rec_ks_bwd_kernel<256, 64>, KsGeom<256, 64>::SMEM, 4); break;
      default: g_rw_cap[3] = rw_query(rec_ks_bwd_kernel<128, 64>, KsGeom<128, 64>::SMEM, 2); break;
    }
  }
  return g_rw_cap[which];
}

template <typename Kern>
int rw_launch(Kern kern, int smem_bytes, dim3 grid, int cluster, cudaStream_t st, const CUtensorMap& tm, const RecWideP& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(RW_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tm, p));
  return 0;
}

int env_on(const char* name) {
  const char* e = getenv(name);
  return (e && e[0] == '0') ? 0 : 1;
}

}  // namespace

// cluster exchange by DSMEM bulk copies (default) or through the L2 ring (SSASR_REC_DSMEM=0): measured 2 818 against 3 251
// cycles per backward step (S = 256, quad clusters)
static int g_rec_dsmem = -1;
int rec_dsmem_enabled() {
  if (g_rec_dsmem < 0) {
    const char* e = getenv("SSASR_REC_DSMEM");
    g_rec_dsmem = e ? atoi(e) : 1;                // 2: the 16-CTA clusters of S = 512 too (A/B switch)
  }
  return g_rec_dsmem;
}

// 1 when the 16-CTA cluster kernels can run this layer with every (direction, 16-utterance tile) cluster co-resident
int rec_wide_supported(int S, int n_batch, int backward) {
  static int on = -1;
  if (on < 0) on = env_on("SSASR_REC_WIDE");
  if (!on || !rec_cl_is_enabled() || S != RW_S || n_batch < 1) return 0;
  const int tiles = (n_batch + RW_NT - 1) / RW_NT;
  return 2 * tiles <= rw_capacity(backward ? 1 : 0) ? 1 : 0;
}

// 1 when the K-split backward kernel (64 units per CTA, 16-row tiles) can run this layer with every cluster co-resident
int rec_ks_supported(int S, int n_batch) {
  static int on = -1;
  if (on < 0) on = env_on("SSASR_REC_KSPLIT");
  if (!on || !rec_cl_is_enabled() || (S != 256 && S != 128) || n_batch < 1) return 0;
  const int tiles = (n_batch + RW_NT - 1) / RW_NT;
  return (2 * tiles <= rw_capacity(S == 256 ? 2 : 3) && 2 * tiles * (S / 64) <= 148) ? 1 : 0;
}

int rec_wide_fwd(cudaStream_t st, float* xp, const void* whh_bf, float* hout, float* cbuf, void* hb, const int* lens, int S,
                 int n_seq, int n_batch, long long rs_seq, long long rs_batch) {
  SSASR_REQUIRE(rec_wide_supported(S, n_batch, 0), "rec_wide_fwd: unsupported shape S=%d n_batch=%d", S, n_batch);
  RecWideP p = {};
  p.xp = xp; p.hout = hout; p.cbuf = cbuf; p.xb = (__nv_bfloat16*)hb; p.lens = lens;
  p.n_seq = n_seq; p.n_batch = n_batch; p.rs_seq = rs_seq; p.rs_batch = rs_batch;
  p.ring = rw_ring_for(st);
  p.w = (const __nv_bfloat16*)whh_bf;
  SSASR_REQUIRE(p.ring != nullptr, "rec_wide_fwd: cannot allocate the exchange ring");
  CUtensorMap tmW;
  int rc = make_tmap_bf16(&tmW, whh_bf, 8 * S, S, S, 128);
  if (rc) return rc;
  ProfScope ps(F_REC_TC_FWD, st);
  return rw_launch(rec_wide_fwd_kernel<false>, RWF_SMEM, dim3(RW_NC, 2, (n_batch + RW_NT - 1) / RW_NT), RW_NC, st, tmW, p);
}

// Exact (split-operand) forward recurrence at S = 512 on the 16-CTA clusters: xp [rows, 8S] fp32 pre-activations, whh_hi / whh_lo
// [8S, S] bf16 parts of the packed W_hh; writes hout only.  Clusters are independent, the grid may run in waves.  -1: not covered.
int rec_wide_fwd_x3(cudaStream_t st, float* xp, const void* whh_hi, const void* whh_lo, float* hout, const int* lens, int S, int n_seq,
                    int n_batch, long long rs_seq, long long rs_batch) {
  static int cap = -1;
  const int on = env_on("SSASR_REC_WIDE") && env_on("SSASR_REC_Q_X3");      // read per call: A/B switch of tests and scripts
  if (!on || !rec_cl_is_enabled() || S != RW_S || n_batch < 1 || n_seq < 4) return -1;
  if (cap < 0) cap = rw_query(rec_wide_fwd_kernel<true>, RWF_SMEM_X3, RW_NC);
  const int tiles = (n_batch + RW_NT - 1) / RW_NT;
  if (cap < 1 || (size_t)RW_NC * 2 * tiles * (2 * RW_HBLK) > (size_t)160 * RW_SLOT) return -1;   // packed slots of the exchange ring
  RecWideP p = {};
  p.xp = xp; p.hout = hout; p.lens = lens;
  p.n_seq = n_seq; p.n_batch = n_batch; p.rs_seq = rs_seq; p.rs_batch = rs_batch;
  p.ring = rw_ring_for(st);
  p.w = (const __nv_bfloat16*)whh_hi;
  SSASR_REQUIRE(p.ring != nullptr, "rec_wide_fwd_x3: cannot allocate the exchange ring");
  CUtensorMap tmW;
  int rc = make_tmap_bf16(&tmW, whh_lo, 8 * S, S, S, 128);
  if (rc) return rc;
  ProfScope ps(F_REC_TC_FWD, st);
  return rw_launch(rec_wide_fwd_kernel<true>, RWF_SMEM_X3, dim3(RW_NC, 2, tiles), RW_NC, st, tmW, p);
}

// K-split backward: S = 512 (16-CTA clusters of 32 units) or S = 256 / 128 (clusters of S / 64 CTAs, 64 units each)
int rec_wide_bwd(cudaStream_t st, float* act, const void* whhT_bf, const float* cbuf, const float* dhout, void* dgb, const int* lens,
                 int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch, float* dbias) {
  SSASR_REQUIRE(S == 512 ? rec_wide_supported(S, n_batch, 1) : rec_ks_supported(S, n_batch),
                "rec_wide_bwd: unsupported shape S=%d n_batch=%d", S, n_batch);
  RecWideP p = {};
  p.xp = act; p.cbuf = const_cast<float*>(cbuf); p.xb = (__nv_bfloat16*)dgb; p.dhout = dhout; p.dbias = dbias; p.lens = lens;
  p.n_seq = n_seq; p.n_batch = n_batch; p.rs_seq = rs_seq; p.rs_batch = rs_batch;
  p.ring = rw_ring_for(st);
  p.dbg = g_rw_dbg;
  p.w = (const __nv_bfloat16*)whhT_bf;
  p.dsmem = rec_dsmem_enabled() >= (S == 512 ? 2 : 1);
  SSASR_REQUIRE(p.ring != nullptr, "rec_wide_bwd: cannot allocate the exchange ring");
  CUtensorMap tmWT;
  int rc = make_tmap_bf16(&tmWT, whhT_bf, 2 * S, 4 * S, 4 * S, 128);
  if (rc) return rc;
  const int tiles = (n_batch + RW_NT - 1) / RW_NT;
  ProfScope ps(F_REC_TC_BWD, st);
  if (S == 512) return rw_launch(rec_ks_bwd_kernel<512, 32>, KsGeom<512, 32>::SMEM, dim3(16, 2, tiles), 16, st, tmWT, p);
  if (S == 256) return rw_launch(rec_ks_bwd_kernel<256, 64>, KsGeom<256, 64>::SMEM, dim3(4, 2, tiles), 4, st, tmWT, p);
  return rw_launch(rec_ks_bwd_kernel<128, 64>, KsGeom<128, 64>::SMEM, dim3(2, 2, tiles), 2, st, tmWT, p);
}

}  // namespace ssasr

extern "C" {
// debug: device buffer [n_seq][12] of clock64 stamps written by CTA (0,0,0) of the next K-split backward launches
void ssasr_rec_wide_set_debug(long long* dev_buf) { ssasr::g_rw_dbg = dev_buf; }
void ssasr_rec_set_dsmem(int mode) { ssasr::g_rec_dsmem = mode < 0 ? 0 : mode; }
}

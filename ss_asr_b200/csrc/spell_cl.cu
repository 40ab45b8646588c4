// Cluster-persistent attend-and-spell step kernels (bf16 training path): ONE launch runs a whole run of teacher-forced
// decoder steps -- attention query, energies, masked softmax, context, layer-1 LSTM cell -- with the recurrent weights, the
// query projection, the decoder state (h1 tile, c1 registers) and the query resident on chip across the steps.
//
// Reference semantics: Attention.forward asr.py:343-392 and the layer-1 LSTMCell of Speller.forward asr.py:314-326 inside the
// loop of ASR.forward asr.py:65-110 (the attention query of step t is the layer-1 state of step t-1, asr.py:84).
//
// Decomposition.  A cluster of 8 CTAs owns 16 utterances; CTA r owns
//   * gate rows [128 r, 128 r + 128) of the layer-1 cell = hidden units [32 r, 32 r + 32) x (i, f, g, o), for all 16 utterances
//   * the attention (query -> energies -> softmax) of utterances 2 r and 2 r + 1.
// The context never materialises: with P[b, j, :] = W_ctx enc[b, j, :] (one batched tensor-core GEMM before the loop)
//     W_ctx ctx_t = W_ctx sum_j alpha_tj enc_j = sum_j alpha_tj P_j
// so the K = E part of the gate product becomes an alpha-weighted sum of P rows, private to an utterance.  The embedding
// part is a table lookup (G_emb[token] = W_emb emb[token] + b, 50 rows), and only the K = S_d recurrent part W_hh h1(t-1)
// is a true cluster-wide product.  Per step and CTA, everything runs on the tensor core out of shared memory:
//   acc_q [128 m x 16 utt]   = phi  [128 x 256] (resident)        x  h1(t-1)^T tile (exchanged)
//   acc_g [128 rows x 16 utt] = W_hh slice [128 x 256] (resident)  x  h1(t-1)^T tile
//                             + sum_u P_u^T [128 x 64 frames] (TMA ring, MN-major)  x  A_u [64 x 16] (alpha_u in column u, 0 elsewhere)
// The second term is ONE accumulation over K = 16 x 64: the B operand is a block-diagonal [16 x 1024] bf16 matrix whose
// diagonal rows are the attention weights of the step.  Two all-gathers per step inside the cluster (h1 slices, alpha rows)
// use the bulk-store + multicast read-back exchange of rec_cl.cu; the P / psi~ tiles stream from L2 through a 5 x 8 KB ring.
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "cl_common.cuh"

namespace ssasr {

using namespace tc;
using namespace clx;

int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows);

namespace {

constexpr int SP_NC = 8;                  // CTAs per cluster
constexpr int SP_MAXU = 24;               // utterances per cluster at most (TMEM: one 16-column accumulator per utterance)
constexpr int SP_SD = 256;                // decoder state size
constexpr int SP_M = 128;                 // attention MLP size
constexpr int SP_TP = 64;                 // encoder frames per frame block; TPB = 1 or 4 blocks per utterance (T' <= 64 / 256)
constexpr int SP_EPW = 16;                // epilogue warps
constexpr int SP_THREADS = 128 + 32 * SP_EPW;     // warp 0 exchange, 1 MMA / TMEM, 2 ring producer, 3 idle, 4.. epilogue
constexpr int SP_STAGE = 16384;           // ring stage: a phi k-block [128 m x 64 k], the psi~ tile [64 frames x 128 m] of an utterance
                                          // or its P tile [64 frames x 128 gate rows]
constexpr int SP_NSTAGE = 7;
constexpr int SP_RING = 4;                // global exchange ring depth
constexpr int SP_SLOT = 4096;             // exchange slot per CTA and ring position: h image (<= 2 KB) + alpha image (<= 512 B)
constexpr int SP_MAXOWN = 4;              // attention utterances per CTA at most

// shared-memory map (bytes from the 1024-aligned base)
constexpr int OFF_W = 0;                          // W_hh slice: 4 k-blocks [128 rows x 128 B], 128-byte swizzle
constexpr int OFF_AL = OFF_W + 65536;             // attention weights of the step [32 utterances x 64 frames] bf16, 128-byte swizzle
constexpr int OFF_H = OFF_AL + 8192;              // (4 frame blocks x 16 rows at TPB = 4)  h1 tile [2 buffers][8 producers][NT rows x 64 B], 64-byte swizzle
constexpr int OFF_RING = OFF_H + 32768;           // P / psi~ ring
constexpr int OFF_HIMG = OFF_RING + SP_NSTAGE * SP_STAGE;
constexpr int OFF_AIMG = OFF_HIMG + 2048;
constexpr int OFF_QS = OFF_AIMG + SP_MAXOWN * 128;   // [4][128] fp32 queries of the CTA's own utterances
constexpr int OFF_RED = OFF_QS + SP_MAXOWN * 512;    // [16] fp32 softmax scratch
constexpr int OFF_BARS = OFF_RED + 64;
constexpr int SP_SMEM = OFF_BARS + 32 * 8 + 1024;    // + alignment slack
static_assert(OFF_RING % 1024 == 0 && SP_SMEM <= 232448, "shared-memory map");

struct SpellClP {
  int B, U, Tp, t0, t1;
  int per;                                // utterances per cluster (the last cluster may hold fewer)
  const float* gemb;                      // [C, 4Sd]
  const int* tok; long long tok_ld;       // [B, U]
  const int* enc_lens;
  float* act1; long long act1_ldb, act1_ldt;
  float* c1; long long c1_ldb, c1_ldt;
  float* h1; long long h1_ldb, h1_ldt;
  __nv_bfloat16* h1b; long long h1b_ldb, h1b_ldt;
  float* q; long long q_ldb, q_ldt;
  float* alpha; long long al_ldb, al_ldt;
  // plain recurrence mode (ATT = false, the layer-2 cell chain): gate pre-activations of the input projection (+ bias)
  const float* xpre; long long xpre_ldb, xpre_ldt;
  float* h2nd; long long h2nd_ldb, h2nd_ldt; int h2nd_toff;   // optional second copy of h(t), written at step t + h2nd_toff (< U)
  uint8_t* ring;
  long long* dbg;                         // optional [steps][8] clock64 stamps of CTA 0
};

#define SP_STAMP(idx)                                                                  \
  do {                                                                                 \
    if (p.dbg && blockIdx.x == 0 && blockIdx.z == 0) p.dbg[(size_t)s * 8 + (idx)] = clock64(); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// forward.  NT = 16 or 32: rows of the h1 / alpha tiles = N of the cluster-wide products (>= utterances of the cluster).
// Utterance slot i of the cluster (b = b0 + i): its attention belongs to CTA i % 8 (own index i / 8), its cells to every CTA
// (32 units each).  TMEM: acc_h [128 x NT] at column 0, acc_q [128 x NT] at NT, one [128 x 16] accumulator per utterance from
// 2 NT on -- the context part of utterance i is column i % 16 of accumulator i (B operand = the 16-row alpha sub-tile holding
// row i; the other 15 columns are cross terms nobody reads).
// ------------------------------------------------------------------------------------------------
// ATT = false: the same cluster recurrence WITHOUT the attention part -- gates = W_hh h(t-1) + xpre[b, t] -- used for the
// layer-2 cell chain of the Speller (its input projection W_ih h1(t) is one batched GEMM per run of steps).
// TPB: 64-frame blocks per utterance (1: T' <= 64; 4: T' <= 256, then with 16-row tiles and at most ONE attention utterance per CTA)
template <int NT, bool ATT, int TPB>
__global__ void __launch_bounds__(SP_THREADS, 1)
spell_cl_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmPhi,
                    const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmPsi, int w_col0, SpellClP p) {
  constexpr int HBLK = NT * 64;           // one producer's k-block of the h1 tile = its outgoing image
  constexpr int NJ = NT / 16;             // cells per epilogue thread
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Wsm = smem + OFF_W;
  uint8_t* Al = smem + OFF_AL;
  uint8_t* Hsm = smem + OFF_H;
  uint8_t* Ring = smem + OFF_RING;
  uint8_t* himg = smem + OFF_HIMG;
  uint8_t* aimg = smem + OFF_AIMG;
  float* qs = reinterpret_cast<float*>(smem + OFF_QS);
  float* red = reinterpret_cast<float*>(smem + OFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;            // [2]
  uint64_t* al_full = bars + 3;
  uint64_t* q_done = bars + 4;
  uint64_t* g_done = bars + 5;
  uint64_t* alpha_ready = bars + 6;
  uint64_t* stage_ready = bars + 7;
  uint64_t* r_full = bars + 8;            // [SP_NSTAGE]
  uint64_t* r_empty = bars + 8 + SP_NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * SP_NSTAGE);

  const int r = blockIdx.x;               // rank in the cluster
  const int b0 = blockIdx.z * p.per;      // first utterance of the cluster
  const int NU = min(p.per, p.B - b0);    // utterances of the cluster (>= 1 by construction of the grid)
  const int n_own = (ATT && NU > r) ? (NU - r + SP_NC - 1) / SP_NC : 0;    // attention utterances of this CTA: slots r, r + 8, ...
  const int step_stages = 4 + (n_own + NU) * TPB;   // ring stages per step: 4 phi k-blocks, own psi~ tiles, all P tiles
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_steps = p.t1 - p.t0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmW);
    if (ATT) {
      tma_prefetch_desc(&tmPhi);
      tma_prefetch_desc(&tmP);
      tma_prefetch_desc(&tmPsi);
    }
    mbar_init(w_full, 1);
    mbar_init(a_full, 1);
    mbar_init(a_full + 1, 1);
    mbar_init(al_full, 1);
    mbar_init(q_done, 1);
    mbar_init(g_done, 1);
    mbar_init(alpha_ready, 1);
    mbar_init(stage_ready, SP_EPW);
    for (int i = 0; i < SP_NSTAGE; ++i) {
      mbar_init(r_full + i, 1);
      mbar_init(r_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < 8192 / 16; i += SP_THREADS) reinterpret_cast<uint4*>(Al)[i] = make_uint4(0u, 0u, 0u, 0u);
  // h1(t0 - 1) of the cluster's utterances -> buffer 1 of the tile (64-byte swizzle), zeros at t0 == 0 and in unused rows
  for (int i = threadIdx.x; i < NT * SP_SD / 2; i += SP_THREADS) {
    const int row = i >> 7, k = (i & 127) * 2;
    float v0 = 0.f, v1 = 0.f;
    if (p.t0 > 0 && row < NU) {
      const float2 hv = *reinterpret_cast<const float2*>(p.h1 + (size_t)(b0 + row) * p.h1_ldb + (size_t)(p.t0 - 1) * p.h1_ldt + k);
      v0 = hv.x; v1 = hv.y;
    }
    const int kb = k >> 5, kk = k & 31;     // producer block (32 units), unit inside it
    *reinterpret_cast<__nv_bfloat162*>(Hsm + (SP_NC + kb) * HBLK + row * 64 + (((kk >> 3) ^ ((row >> 1) & 3)) << 4) + (kk & 7) * 2) =
        __floats2bfloat162_rn(v0, v1);
  }
  for (int i = threadIdx.x; i < HBLK / 4; i += SP_THREADS) reinterpret_cast<uint32_t*>(himg)[i] = 0u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(w_full, 65536);
    for (int kb = 0; kb < 4; ++kb) tma_load_2d(&tmW, w_full, Wsm + kb * 16384, w_col0 + kb * 64, r * 128);
    mbar_expect_tx(a_full, SP_NC * HBLK);
    mbar_expect_tx(a_full + 1, SP_NC * HBLK);
    mbar_expect_tx(al_full, NU * TPB * 128);
  }
  cluster_sync_all();                      // every CTA's barriers exist (and are armed) before any multicast can signal them

  if (warp == 0) {
    // ---------------- exchange thread ----------------
    if (elect_one()) {
      const uint16_t cmask = (uint16_t)((1u << SP_NC) - 1u);
      const size_t n_cta = (size_t)gridDim.x * gridDim.z;
      const size_t cta = (size_t)blockIdx.z * gridDim.x + blockIdx.x;
      for (int s = 0; s < n_steps; ++s) {
        uint8_t* slot = p.ring + ((size_t)(s % SP_RING) * n_cta + cta) * SP_SLOT;
        if (ATT && n_own > 0) {
          mbar_wait_t(alpha_ready, s & 1);
          SP_STAMP(2);
          bulk_store_wait(slot + 2048, aimg, n_own * TPB * 128);
          for (int o = 0; o < n_own; ++o)      // row r + 8 o of the alpha tile, one 128-byte piece per frame block
            for (int f = 0; f < TPB; ++f)
              bulk_load_mc(Al + f * (NT * 128) + (r + 8 * o) * 128, slot + 2048 + (o * TPB + f) * 128, 128, al_full, cmask);
        }
        mbar_wait_t(stage_ready, s & 1);
        SP_STAMP(5);
        if (s + 1 < n_steps) {
          bulk_store_wait(slot, himg, HBLK);
          bulk_load_mc(Hsm + ((s & 1) * SP_NC + r) * HBLK, slot, HBLK, a_full + (s & 1), cmask);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA thread ----------------
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, NT);               // A, B K-major
      constexpr uint32_t idesc_p = umma_idesc_bf16(128, 16, 1, 0);       // A = P tile, MN-major; B = 16 rows of the alpha tile
      const uint32_t acc_h = tmem, acc_q = tmem + NT, acc_u = tmem + 2 * NT;
      mbar_wait_t(w_full, 0);
      int rp = 0;                           // ring position of the first stage of the step
      for (int s = 0; s < n_steps; ++s, rp += step_stages) {
        const int hb = (s + 1) & 1;         // buffer holding h1(t - 1): the initial tile sits in buffer 1
        if (s > 0) {
          mbar_wait_t(a_full + hb, ((s - 1) >> 1) & 1);
          if (s + 2 < n_steps) mbar_expect_tx(a_full + hb, SP_NC * HBLK);     // this buffer next receives h1 of step s + 1
        }
        SP_STAMP(0);
        tc_fence_after();
        const uint32_t h0 = smem_u32(Hsm + hb * SP_NC * HBLK);
        const uint32_t w0 = smem_u32(Wsm);
        // query pre-activations: phi streams through the ring (4 k-blocks, prefetched while the previous step finished)
        for (int kb = 0; ATT && kb < 4; ++kb) {
          const int pos = rp + kb, stg = pos % SP_NSTAGE;
          mbar_wait_t(r_full + stg, (pos / SP_NSTAGE) & 1);
          tc_fence_after();
          const uint32_t f0 = smem_u32(Ring + stg * SP_STAGE);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int kk = kb * 4 + k4;
            const uint64_t dh = umma_desc_k64(h0 + (kk >> 1) * HBLK) + (uint64_t)((kk & 1) * 2);
            mma_bf16_ss(acc_q, umma_desc_k128(f0) + (uint64_t)(k4 * 2), dh, idesc, kk != 0);
          }
          mma_commit(r_empty + stg);
        }
        if (ATT) mma_commit(q_done);
#pragma unroll 4
        for (int kk = 0; kk < 16; ++kk) {
          const uint64_t dh = umma_desc_k64(h0 + (kk >> 1) * HBLK) + (uint64_t)((kk & 1) * 2);
          mma_bf16_ss(acc_h, umma_desc_k128(w0 + (kk >> 2) * 16384) + (uint64_t)((kk & 3) * 2), dh, idesc, kk != 0);
        }
        if (ATT) {
          mbar_wait_t(al_full, s & 1);      // the attention rows of every utterance of the cluster have landed
          if (s + 1 < n_steps) mbar_expect_tx(al_full, NU * TPB * 128);
          SP_STAMP(3);
          tc_fence_after();
        }
        const uint32_t al0 = smem_u32(Al);
        for (int i = 0; ATT && i < NU * TPB; ++i) {
          const int u = i / TPB, f = i % TPB;
          const int pos = rp + 4 + n_own * TPB + i, stg = pos % SP_NSTAGE;
          mbar_wait_t(r_full + stg, (pos / SP_NSTAGE) & 1);
          tc_fence_after();
          const uint64_t da = umma_desc_mn128(smem_u32(Ring + stg * SP_STAGE), 8192);
          const uint64_t db = umma_desc_k128(al0 + f * (NT * 128) + (u >> 4) * 2048);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            mma_bf16_ss(acc_u + 16 * u, da + (uint64_t)(k4 * 128), db + (uint64_t)(k4 * 2), idesc_p, (f | k4) != 0);
          mma_commit(r_empty + stg);
        }
        mma_commit(g_done);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ---------------- ring producer ----------------
    if (ATT && elect_one()) {
      int pos = 0;
      for (int s = 0; s < n_steps; ++s) {
        for (int i = 0; i < step_stages; ++i, ++pos) {
          const int stg = pos % SP_NSTAGE;
          mbar_wait_t(r_empty + stg, ((pos / SP_NSTAGE) & 1) ^ 1);
          mbar_expect_tx(r_full + stg, SP_STAGE);
          uint8_t* dst = Ring + stg * SP_STAGE;
          if (i < 4) {                      // phi k-block i: [128 m x 64 k]
            tma_load_2d(&tmPhi, r_full + stg, dst, i * 64, 0);
          } else if (i < 4 + n_own * TPB) { // psi~ of own utterance o (slot r + 8 o), frame block f: two k-blocks of [64 frames x 64 m]
            const int o = (i - 4) / TPB, f = (i - 4) % TPB;
            const int row = (b0 + r + 8 * o) * p.Tp + 64 * f;
            tma_load_2d(&tmPsi, r_full + stg, dst, 0, row);
            tma_load_2d(&tmPsi, r_full + stg, dst + 8192, 64, row);
          } else {                          // P of utterance u, frame block f, this CTA's 128 gate rows: two blocks of [64 frames x 64 rows]
            const int j = i - 4 - n_own * TPB;
            const int row = (b0 + j / TPB) * p.Tp + 64 * (j % TPB);
            tma_load_2d(&tmP, r_full + stg, dst, r * 128, row);
            tma_load_2d(&tmP, r_full + stg, dst + 8192, r * 128 + 64, row);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    const int sp = warp & 3;                       // TMEM sub-partition: accumulator rows 32 sp .. 32 sp + 31
    const int cg = (warp - 4) >> 2;                // column group: utterance slots 4 cg .. 4 cg + 3 (+ 16 j)
    const int uq = lane >> 2, gp = lane & 3;
    const int unit = 32 * r + 8 * sp + uq;         // hidden unit of this thread's cells
    float creg[NJ];
    bool cvalid[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int slot = 4 * cg + gp + 16 * j;
      cvalid[j] = slot < NU;
      creg[j] = 0.f;
      if (cvalid[j] && p.t0 > 0) creg[j] = p.c1[(size_t)(b0 + slot) * p.c1_ldb + (size_t)(p.t0 - 1) * p.c1_ldt + unit];
    }
    // attention role (warps 4-11): thread = (own utterance ua, frame ja)
    const int ta = threadIdx.x - 128;
    const int ua = (ta >> 6) & 3, ja = ta & 63;
    const bool att_warp = warp < 12;
    const bool att_on = att_warp && ua < n_own;
    const int sa_slot = r + 8 * ua;                // utterance slot of the attention role = row of the alpha tile
    const int ba = b0 + sa_slot;
    int len_a = 0;
    if (att_on) len_a = min(p.enc_lens[ba], p.Tp);

    for (int s = 0; s < n_steps; ++s) {
      const int t = p.t0 + s;
      // the embedding part of the gate pre-activations (+ bias): a table row, fetched before anything it could wait behind
      float4 ge[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        ge[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cvalid[j]) {
          const size_t bq = (size_t)(b0 + 4 * cg + gp + 16 * j);
          if (ATT) {
            const int tk = p.tok[bq * p.tok_ld + t];
            ge[j] = __ldg(reinterpret_cast<const float4*>(p.gemb + (size_t)tk * 4 * SP_SD + (size_t)unit * 4));
          } else {
            ge[j] = __ldcs(reinterpret_cast<const float4*>(p.xpre + bq * p.xpre_ldb + (size_t)t * p.xpre_ldt + (size_t)unit * 4));
          }
        }
      }
      if (ATT && att_warp && n_own > 0) {
        if (warp < 8) {
          // ---- queries of the CTA's own utterances: rows (m) of acc_q, column = utterance slot ----
          mbar_wait_t(q_done, s & 1);
          if (ta == 0) SP_STAMP(1);
          tc_fence_after();
          uint32_t v[SP_MAXOWN];
#pragma unroll
          for (int o = 0; o < SP_MAXOWN; ++o)
            if (o < n_own) tmem_ld1(tmem + ((uint32_t)(sp * 32) << 16) + (uint32_t)(NT + r + 8 * o), v + o);
          tmem_ld_wait();
          tc_fence_before();
          const int m = sp * 32 + lane;
#pragma unroll
          for (int o = 0; o < SP_MAXOWN; ++o)
            if (o < n_own) {
              const float qv = tanhf(__uint_as_float(v[o]));
              qs[o * SP_M + m] = qv;
              p.q[(size_t)(b0 + r + 8 * o) * p.q_ldb + (size_t)t * p.q_ldt + m] = qv;
            }
        }
        named_bar(1, 256);
        // ---- energies of frames ja + 64 f of own utterance ua: psi~ rows (ring stages, one per frame block) . q ----
        float e[TPB];
#pragma unroll
        for (int f = 0; f < TPB; ++f) e[f] = -INFINITY;
        if (att_on) {
#pragma unroll
          for (int f = 0; f < TPB; ++f) {
            const int pos = s * step_stages + 4 + ua * TPB + f, stg = pos % SP_NSTAGE;
            mbar_wait_t(r_full + stg, (pos / SP_NSTAGE) & 1);
            float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;      // four independent chains
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const uint8_t* row = Ring + stg * SP_STAGE + kb * 8192 + ja * 128;
              const float* qq = qs + ua * SP_M + kb * 64;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 w = *reinterpret_cast<const uint4*>(row + ((c ^ (ja & 7)) << 4));
                const float4 qa = *reinterpret_cast<const float4*>(qq + c * 8), qb = *reinterpret_cast<const float4*>(qq + c * 8 + 4);
                e0 = fmaf(bf_lo(w.x), qa.x, e0); e1 = fmaf(bf_hi(w.x), qa.y, e1);
                e2 = fmaf(bf_lo(w.y), qa.z, e2); e3 = fmaf(bf_hi(w.y), qa.w, e3);
                e0 = fmaf(bf_lo(w.z), qb.x, e0); e1 = fmaf(bf_hi(w.z), qb.y, e1);
                e2 = fmaf(bf_lo(w.w), qb.z, e2); e3 = fmaf(bf_hi(w.w), qb.w, e3);
              }
            }
            e[f] = (ja + 64 * f < len_a) ? (e0 + e1) + (e2 + e3) : -INFINITY;
          }
        }
        // ---- masked softmax over the utterance's frames (two warps) ----
        float em = e[0];
#pragma unroll
        for (int f = 1; f < TPB; ++f) em = fmaxf(em, e[f]);
        const float wm = warp_max(em);
        if (lane == 0) red[ta >> 5] = wm;
        named_bar(1, 256);
        const float mx = fmaxf(red[ua * 2], red[ua * 2 + 1]);
        float ex[TPB], es = 0.f;
#pragma unroll
        for (int f = 0; f < TPB; ++f) {
          ex[f] = (e[f] == -INFINITY) ? 0.f : expf(e[f] - mx);
          es += ex[f];
        }
        const float wsum = warp_sum(es);
        if (lane == 0) red[8 + (ta >> 5)] = wsum;
        named_bar(1, 256);
        if (att_on) {
          const float sum = red[8 + ua * 2] + red[8 + ua * 2 + 1];
          const float inv = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
          for (int f = 0; f < TPB; ++f) {
            const float al = ex[f] * inv;
            if (ja + 64 * f < p.Tp) p.alpha[(size_t)ba * p.al_ldb + (size_t)t * p.al_ldt + ja + 64 * f] = al;
            *reinterpret_cast<__nv_bfloat16*>(aimg + (ua * TPB + f) * 128 + (((ja >> 3) ^ (sa_slot & 7)) << 4) + (ja & 7) * 2) = __float2bfloat16_rn(al);
          }
        }
        fence_proxy_async();
        named_bar(1, 256);
        if (ta == 0) {
          mbar_arrive(alpha_ready);
          for (int i = 0; i < n_own * TPB; ++i) mbar_arrive(r_empty + (s * step_stages + 4 + i) % SP_NSTAGE);   // psi~ stages are free
        }
      }
      // ---- layer-1 cells of (unit, slots 4 cg + gp + 16 j): gates = acc_h + context column of the slot's accumulator + G_emb ----
      mbar_wait_t(g_done, s & 1);
      if (ta == 0) SP_STAMP(4);
      tc_fence_after();
      uint32_t vh[NJ][4], vu[NJ][4];
      const uint32_t lane_addr = tmem + ((uint32_t)(sp * 32) << 16);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        tmem_ld4(lane_addr + (uint32_t)(4 * cg + 16 * j), vh[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int sl = 4 * cg + i + 16 * j;
          vu[j][i] = 0u;
          if (ATT && sl < NU) tmem_ld1(lane_addr + (uint32_t)(2 * NT + 16 * sl + (sl & 15)), &vu[j][i]);
        }
      }
      tmem_ld_wait();
      tc_fence_before();
      float4 act[NJ];
      float cvs[NJ], hvs[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float a0 = __uint_as_float(vh[j][0]) + __uint_as_float(vu[j][0]), a1 = __uint_as_float(vh[j][1]) + __uint_as_float(vu[j][1]);
        float a2 = __uint_as_float(vh[j][2]) + __uint_as_float(vu[j][2]), a3 = __uint_as_float(vh[j][3]) + __uint_as_float(vu[j][3]);
        quad_transpose(a0, a1, a2, a3, gp);
        float4 a;
        a.x = sigmoid_apx(ge[j].x + a0); a.y = sigmoid_apx(ge[j].y + a1); a.z = tanh_apx(ge[j].z + a2); a.w = sigmoid_apx(ge[j].w + a3);
        const float cv = cvalid[j] ? fmaf(a.y, creg[j], a.x * a.z) : 0.f;
        const float hv = cvalid[j] ? a.w * tanh_apx(cv) : 0.f;
        creg[j] = cv;
        act[j] = a; cvs[j] = cv; hvs[j] = hv;
        const int slot = 4 * cg + gp + 16 * j;
        *reinterpret_cast<__nv_bfloat16*>(himg + slot * 64 + ((sp ^ ((slot >> 1) & 3)) << 4) + uq * 2) = __float2bfloat16_rn(hv);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(stage_ready);
      // ---- off the dependent chain: tensors saved for the backward pass / layer 2 ----
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (cvalid[j]) {
          const int bc = b0 + 4 * cg + gp + 16 * j;
          *reinterpret_cast<float4*>(p.act1 + (size_t)bc * p.act1_ldb + (size_t)t * p.act1_ldt + (size_t)unit * 4) = act[j];
          p.c1[(size_t)bc * p.c1_ldb + (size_t)t * p.c1_ldt + unit] = cvs[j];
          p.h1[(size_t)bc * p.h1_ldb + (size_t)t * p.h1_ldt + unit] = hvs[j];
          if (p.h1b) p.h1b[(size_t)bc * p.h1b_ldb + (size_t)t * p.h1b_ldt + unit] = __float2bfloat16_rn(hvs[j]);
          if (p.h2nd && t + p.h2nd_toff < p.U) p.h2nd[(size_t)bc * p.h2nd_ldb + (size_t)(t + p.h2nd_toff) * p.h2nd_ldt + unit] = hvs[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
  cluster_sync_all();                      // no CTA leaves while a peer could still address its shared memory
}

// ------------------------------------------------------------------------------------------------
// backward.  Same decomposition: CTA r owns gate rows [128 r, 128 r + 128) (its 32 hidden units) for every utterance of the
// cluster and the attention of utterance slots r, r + 8, ...  Per step t (descending):
//   1. cells      dh1(t) = dh_in[b, t] (from the layer above) + the 8 partial sums pushed by the previous step; cell backward
//                 -> the CTA's slice of dG(t) [utterances x 128 gate rows]: fp32 in place of the saved activations, bf16 to the
//                 weight-gradient operand and to a shared-memory operand tile
//   2. dalpha     (ATT) partial dalpha[u, j] = sum over the CTA's 128 gate rows of dG[u, .] P[u, j, .] on the CUDA cores, the P
//                 tiles streaming through the TMA ring exactly as in the forward kernel; partials pushed to the utterance's owner
//   3. attention  (ATT, owner) softmax backward, dq = (sum_j de_j psi~_j)(1 - q^2); dq rows all-gathered (multicast)
//   4. dh         K-split: acc_d [256 units x utterances] = W_hh slice^T (the resident forward image read MN-major) x dG slice^T
//                 (+ phi slice^T x dq slice^T, 16 of the 128 query rows per CTA); bf16 partials pushed to the units' owners
// Exchanges are bulk store -> (multi)cast read-back with a one-CTA mask for the targeted pushes.
// ------------------------------------------------------------------------------------------------
constexpr int SB_NSTAGE = 4;
constexpr int BOFF_W = 0;                             // W_hh slice, forward layout (4 unit blocks of [128 gate rows x 128 B])
constexpr int BOFF_PHI = BOFF_W + 65536;              // phi rows [16 r, 16 r + 16): 4 unit blocks of [16 rows x 128 B]
constexpr int BOFF_DG = BOFF_PHI + 8192;              // dG slice operand [NT utterances x 128 gate rows] bf16, 2 k-blocks
constexpr int BOFF_DQ = BOFF_DG + 8192;               // all-gathered dq tile [NT x 128 m] bf16, 2 k-blocks
constexpr int BOFF_DHOUT = BOFF_DQ + 8192;            // outgoing dh partials [8 destinations][NT][32 units] bf16
constexpr int BOFF_DHIN = BOFF_DHOUT + 16384;         // incoming dh partials [2 buffers][8 sources][NT][32 units] bf16
constexpr int BOFF_DAOUT = BOFF_DHIN + 32768;         // outgoing dalpha partials [8 owners][4][64] fp32
constexpr int BOFF_DAIN = BOFF_DAOUT + 8192;          // incoming dalpha partials [8 sources][4][64] fp32
constexpr int BOFF_RING = BOFF_DAIN + 8192;           // P / psi~ ring
constexpr int BOFF_DQIMG = BOFF_RING + SB_NSTAGE * SP_STAGE;   // outgoing dq rows [4][2 k-blocks][128 B]
constexpr int BOFF_DES = BOFF_DQIMG + 1024;           // [4][64] fp32 de of the own utterances
constexpr int BOFF_RED = BOFF_DES + 1024;
constexpr int BOFF_BARS = BOFF_RED + 64;
constexpr int SB_SMEM = BOFF_BARS + 32 * 8 + 1024;
static_assert(BOFF_RING % 1024 == 0 && SB_SMEM <= 232448, "shared-memory map (backward)");
constexpr int SB_SLOT = 32768;                        // exchange slot: dh partials 16 KB | dalpha partials 8 KB | dq rows 1 KB

struct SpellClBwdP {
  int B, U, Tp, t0, t1, per;
  const int* enc_lens;
  float* act; long long act_ldb, act_ldt;             // in: gate activations, out: gate gradients (in place)
  const float* c; long long c_ldb, c_ldt;
  const float* dh_in; long long dh_ldb, dh_ldt;       // gradient arriving on h(t) from outside the chain
  __nv_bfloat16* dgb; long long dgb_ldb, dgb_ldt;     // out: bf16 copy of the gate gradients [.., 4Sd]
  const float* alpha; long long al_ldb, al_ldt;
  const float* q; long long q_ldb, q_ldt;
  float* de; long long de_ldb, de_ldt;                // out [.., Tp]
  float* dqpre; long long dq_ldb, dq_ldt;             // out [.., M]
  uint8_t* ring;
  long long* dbg;                                     // optional [steps][8] clock64 stamps of CTA 0 (thread 128)
};
#define SB_STAMP(idx)                                                                                       \
  do {                                                                                                      \
    if (p.dbg && blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 128) p.dbg[(size_t)s * 8 + (idx)] = clock64(); \
  } while (0)

template <int NT, bool ATT, int TPB>
__global__ void __launch_bounds__(SP_THREADS, 1)
spell_cl_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmPhiS,
                    const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmPsi, int w_col0, SpellClBwdP p) {
  constexpr int DHB = NT * 64;              // one (source, destination) block of dh partials: [NT utterances][32 units] bf16
  constexpr int NJ = NT / 16;               // cells per thread
  constexpr int KBLK = NT * 128;            // one k-block of the dG / dq operand tiles
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Wsm = smem + BOFF_W;
  uint8_t* Phis = smem + BOFF_PHI;
  uint8_t* dGsm = smem + BOFF_DG;
  uint8_t* dQsm = smem + BOFF_DQ;
  uint8_t* dhout = smem + BOFF_DHOUT;
  uint8_t* dhin = smem + BOFF_DHIN;
  float* daout = reinterpret_cast<float*>(smem + BOFF_DAOUT);
  float* dain = reinterpret_cast<float*>(smem + BOFF_DAIN);
  uint8_t* Ring = smem + BOFF_RING;
  uint8_t* dqimg = smem + BOFF_DQIMG;
  float* des = reinterpret_cast<float*>(smem + BOFF_DES);
  float* red = reinterpret_cast<float*>(smem + BOFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BOFF_BARS);
  uint64_t* w_full = bars;
  uint64_t* dh_full = bars + 1;             // [2]
  uint64_t* da_full = bars + 3;
  uint64_t* dq_full = bars + 4;
  uint64_t* d_done = bars + 5;
  uint64_t* dg_ready = bars + 6;
  uint64_t* da_ready = bars + 7;
  uint64_t* dq_ready = bars + 8;
  uint64_t* dh_ready = bars + 9;
  uint64_t* r_full = bars + 10;             // [SB_NSTAGE]
  uint64_t* r_empty = bars + 10 + SB_NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10 + 2 * SB_NSTAGE);

  const int r = blockIdx.x;
  const int b0 = blockIdx.z * p.per;
  const int NU = min(p.per, p.B - b0);
  const int n_own = (ATT && NU > r) ? (NU - r + SP_NC - 1) / SP_NC : 0;
  const int step_stages = ATT ? (NU + n_own) * TPB : 0;   // ring stages per step: all P tiles, then the own psi~ tiles (n_own * TPB <= 4)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_steps = p.t1 - p.t0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmW);
    if (ATT) {
      tma_prefetch_desc(&tmPhiS);
      tma_prefetch_desc(&tmP);
      tma_prefetch_desc(&tmPsi);
    }
    mbar_init(w_full, 1);
    mbar_init(dh_full, 1);
    mbar_init(dh_full + 1, 1);
    mbar_init(da_full, 1);
    mbar_init(dq_full, 1);
    mbar_init(d_done, 1);
    mbar_init(dg_ready, SP_EPW);
    mbar_init(da_ready, SP_EPW);
    mbar_init(dq_ready, 1);
    mbar_init(dh_ready, SP_EPW);
    for (int i = 0; i < SB_NSTAGE; ++i) {
      mbar_init(r_full + i, 1);
      mbar_init(r_empty + i, SP_EPW);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<64>(tmem_slot);
  // operand tiles start out as zeros: rows of unused utterance slots are never written
  for (int i = threadIdx.x; i < (8192 + 8192 + 16384) / 16; i += SP_THREADS) reinterpret_cast<uint4*>(dGsm)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < 8192 / 16; i += SP_THREADS) reinterpret_cast<uint4*>(daout)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    mbar_expect_tx(w_full, 65536 + (ATT ? 8192 : 0));
    for (int kb = 0; kb < 4; ++kb) {
      tma_load_2d(&tmW, w_full, Wsm + kb * 16384, w_col0 + kb * 64, r * 128);
      if (ATT) tma_load_2d(&tmPhiS, w_full, Phis + kb * 2048, kb * 64, r * 16);
    }
    mbar_expect_tx(dh_full, SP_NC * DHB);
    mbar_expect_tx(dh_full + 1, SP_NC * DHB);
    if (ATT) {
      mbar_expect_tx(da_full, SP_NC * n_own * TPB * 256);
      mbar_expect_tx(dq_full, NU * 256);
    }
  }
  cluster_sync_all();

  if (warp == 0) {
    // ---------------- exchange thread ----------------
    if (elect_one()) {
      const uint16_t cmask = (uint16_t)((1u << SP_NC) - 1u);
      const size_t n_cta = (size_t)gridDim.x * gridDim.z;
      const size_t cta = (size_t)blockIdx.z * gridDim.x + blockIdx.x;
      for (int s = 0; s < n_steps; ++s) {
        uint8_t* slot = p.ring + ((size_t)(s % SP_RING) * n_cta + cta) * SB_SLOT;
        if (ATT) {
          mbar_wait_t(da_ready, s & 1);
          bulk_store_wait(slot + 16384, daout, 8192);
          for (int o = 0; o < SP_NC; ++o) {            // partial dalpha rows of owner o's utterances -> its row block r
            const int no = (NU > o) ? (NU - o + SP_NC - 1) / SP_NC : 0;
            if (no > 0) bulk_load_mc(dain + r * 256, slot + 16384 + o * 1024, no * TPB * 256, da_full, (uint16_t)(1u << o));
          }
          if (n_own > 0 && s + 1 < n_steps) {      // (the dq of the last step feeds nothing inside the chain)
            mbar_wait_t(dq_ready, s & 1);
            bulk_store_wait(slot + 24576, dqimg, n_own * 256);
            for (int o = 0; o < n_own; ++o)
              for (int kb = 0; kb < 2; ++kb)
                bulk_load_mc(dQsm + kb * KBLK + (r + 8 * o) * 128, slot + 24576 + (o * 2 + kb) * 128, 128, dq_full, cmask);
          }
        }
        if (s + 1 < n_steps) {
          mbar_wait_t(dh_ready, s & 1);
          bulk_store_wait(slot, dhout, SP_NC * DHB);
          for (int d = 0; d < SP_NC; ++d)
            bulk_load_mc(dhin + ((s & 1) * SP_NC + r) * DHB, slot + d * DHB, DHB, dh_full + (s & 1), (uint16_t)(1u << d));
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA thread ----------------
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, NT, 1, 0);        // A MN-major (units contiguous), B K-major
      mbar_wait_t(w_full, 0);
      for (int s = 0; s + 1 < n_steps; ++s) {                            // the last step has no predecessor to feed
        mbar_wait_t(dg_ready, s & 1);
        tc_fence_after();
        const uint32_t w0 = smem_u32(Wsm), g0 = smem_u32(dGsm);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            mma_bf16_ss(tmem + h * NT, umma_desc_mn128(w0 + h * 32768, 16384) + (uint64_t)(kk * 128),
                        umma_desc_k128(g0 + (kk >> 2) * KBLK) + (uint64_t)((kk & 3) * 2), idesc, kk != 0);
        if (ATT) {
          mbar_wait_t(dq_full, s & 1);
          if (s + 2 < n_steps) mbar_expect_tx(dq_full, NU * 256);
          tc_fence_after();
          const uint32_t f0 = smem_u32(Phis), q0 = smem_u32(dQsm);
#pragma unroll
          for (int h = 0; h < 2; ++h)
            mma_bf16_ss(tmem + h * NT, umma_desc_mn128(f0 + h * 4096, 2048), umma_desc_k128(q0 + (r >> 2) * KBLK) + (uint64_t)((r & 3) * 2),
                        idesc, 1u);
        }
        mma_commit(d_done);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ---------------- ring producer ----------------
    if (ATT && elect_one()) {
      int pos = 0;
      for (int s = 0; s < n_steps; ++s) {
        for (int i = 0; i < step_stages; ++i, ++pos) {
          const int stg = pos % SB_NSTAGE;
          mbar_wait_t(r_empty + stg, ((pos / SB_NSTAGE) & 1) ^ 1);
          mbar_expect_tx(r_full + stg, SP_STAGE);
          uint8_t* dst = Ring + stg * SP_STAGE;
          if (i < NU * TPB) {               // P of utterance i / TPB, frame block i % TPB, this CTA's 128 gate rows
            const int row = (b0 + i / TPB) * p.Tp + 64 * (i % TPB);
            tma_load_2d(&tmP, r_full + stg, dst, r * 128, row);
            tma_load_2d(&tmP, r_full + stg, dst + 8192, r * 128 + 64, row);
          } else {                          // psi~ of own utterance o, frame block f
            const int j = i - NU * TPB;
            const int row = (b0 + r + 8 * (j / TPB)) * p.Tp + 64 * (j % TPB);
            tma_load_2d(&tmPsi, r_full + stg, dst, 0, row);
            tma_load_2d(&tmPsi, r_full + stg, dst + 8192, 64, row);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------- compute warps ----------------
    const int cw = warp - 4;                       // 0 .. 15
    const int ct = threadIdx.x - 128;              // 0 .. 511
    const int unit = 32 * r + lane;                // cells: warp cw owns utterance slots cw (+ 16), lane = unit of the CTA
    bool cvalid[NJ];
    float dcreg[NJ];
    float4 act_n[NJ];
    float c_n[NJ], cp_n[NJ], dh_n[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      cvalid[j] = cw + 16 * j < NU;
      dcreg[j] = 0.f;
    }
    // everything of a step that does not depend on the chain is fetched one whole step ahead
    auto fetch = [&](int t_) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (cvalid[j]) {
          const size_t bq = (size_t)(b0 + cw + 16 * j);
          act_n[j] = __ldcs(reinterpret_cast<const float4*>(p.act + bq * p.act_ldb + (size_t)t_ * p.act_ldt + (size_t)unit * 4));
          c_n[j] = __ldg(p.c + bq * p.c_ldb + (size_t)t_ * p.c_ldt + unit);
          cp_n[j] = t_ > 0 ? __ldg(p.c + bq * p.c_ldb + (size_t)(t_ - 1) * p.c_ldt + unit) : 0.f;
          dh_n[j] = __ldcs(p.dh_in + bq * p.dh_ldb + (size_t)t_ * p.dh_ldt + unit);
        }
      }
    };
    fetch(p.t1 - 1);
    // attention role (warps 4-11): thread = (own utterance ua, frame ja)
    const int ua = (ct >> 6) & 3, ja = ct & 63;
    const bool att_warp = ATT && cw < 8;
    const bool att_on = att_warp && ua < n_own;
    const int sa_slot = r + 8 * ua;
    const int ba = b0 + sa_slot;
    int len_a = 0;
    if (att_on) len_a = min(p.enc_lens[ba], p.Tp);
    // dalpha role: thread = (frame 4 cw + (lane >> 3), 16-column part lane & 7 of the CTA's 128 gate rows)
    const int jf = 4 * cw + (lane >> 3), part = lane & 7;
    int rpos = 0;                                  // ring position of the compute warps

    for (int s = 0; s < n_steps; ++s) {
      const int t = p.t1 - 1 - s;
      float4 a[NJ];
      float cv[NJ], cpv[NJ], dh[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) { a[j] = act_n[j]; cv[j] = c_n[j]; cpv[j] = cp_n[j]; dh[j] = dh_n[j]; }
      if (s + 1 < n_steps) fetch(t - 1);
      float al_a[TPB], q_a0 = 0.f, q_a1 = 0.f;
#pragma unroll
      for (int f = 0; f < TPB; ++f) al_a[f] = 0.f;
      if (att_on) {
#pragma unroll
        for (int f = 0; f < TPB; ++f)
          if (ja + 64 * f < p.Tp) al_a[f] = __ldg(p.alpha + (size_t)ba * p.al_ldb + (size_t)t * p.al_ldt + ja + 64 * f);
        const float2 qq = *reinterpret_cast<const float2*>(p.q + (size_t)ba * p.q_ldb + (size_t)t * p.q_ldt + 2 * ja);
        q_a0 = qq.x; q_a1 = qq.y;
      }
      // ---- 1. dh1(t): the partial sums of the 8 CTAs of the previous step, then the cell backward ----
      if (s > 0) {
        mbar_wait_t(dh_full + ((s - 1) & 1), ((s - 1) >> 1) & 1);
        SB_STAMP(0);
        if (ct == 0 && s + 2 < n_steps) mbar_expect_tx(dh_full + ((s - 1) & 1), SP_NC * DHB);   // re-armed for the pushes of step s + 1
        const uint8_t* base = dhin + ((s - 1) & 1) * SP_NC * DHB;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int src = 0; src < SP_NC; ++src)
            acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(base + src * DHB + (cw + 16 * j) * 64 + lane * 2));
          dh[j] += acc;
        }
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int slot = cw + 16 * j;
        float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cvalid[j]) {
          const float tc = tanh_apx(cv[j]);
          const float dc = dh[j] * a[j].w * (1.f - tc * tc) + dcreg[j];
          dg.w = dh[j] * tc * a[j].w * (1.f - a[j].w);
          dg.x = dc * a[j].z * a[j].x * (1.f - a[j].x);
          dg.z = dc * a[j].x * (1.f - a[j].z * a[j].z);
          dg.y = dc * cpv[j] * a[j].y * (1.f - a[j].y);
          dcreg[j] = dc * a[j].y;
        }
        const __nv_bfloat162 b01 = __floats2bfloat162_rn(dg.x, dg.y), b23 = __floats2bfloat162_rn(dg.z, dg.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&b01);
        pk.y = *reinterpret_cast<const uint32_t*>(&b23);
        // operand tile: row = utterance slot, gate column 4 lane + g of the CTA's 128: k-block lane / 16, 16-byte chunk (lane % 16) / 2
        *reinterpret_cast<uint2*>(dGsm + (lane >> 4) * KBLK + slot * 128 + ((((lane & 15) >> 1) ^ (slot & 7)) << 4) + (lane & 1) * 8) = pk;
        if (cvalid[j]) {
          const size_t bq = (size_t)(b0 + slot);
          __stcs(reinterpret_cast<float4*>(p.act + bq * p.act_ldb + (size_t)t * p.act_ldt + (size_t)unit * 4), dg);
          *reinterpret_cast<uint2*>(p.dgb + bq * p.dgb_ldb + (size_t)t * p.dgb_ldt + (size_t)unit * 4) = pk;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(dg_ready);
      SB_STAMP(1);
      if (ATT) {
        // ---- 2. partial dalpha over the CTA's 128 gate rows, all utterances of the cluster ----
        mbar_wait_t(dg_ready, s & 1);               // every warp's rows of the operand tile are in place
        for (int i = 0; i < NU * TPB; ++i, ++rpos) {
          const int u = i / TPB, fb = i % TPB;
          const int stg = rpos % SB_NSTAGE;
          mbar_wait_t(r_full + stg, (rpos / SB_NSTAGE) & 1);
          // P row jf, columns 16 part .. 16 part + 15: k-block part / 4, chunks 2 (part % 4) and + 1 (order swapped in the upper
          // k-block so that the 8 lanes of a quarter warp hit 8 different bank groups)
          const uint8_t* prow = Ring + stg * SP_STAGE + (part >> 2) * 8192 + jf * 128;
          const uint8_t* grow = dGsm + (part >> 2) * KBLK + u * 128;
          const int c0 = 2 * (part & 3) + (part >> 2), c1 = c0 ^ 1;
          const uint4 w0 = *reinterpret_cast<const uint4*>(prow + ((c0 ^ (jf & 7)) << 4));
          const uint4 w1 = *reinterpret_cast<const uint4*>(prow + ((c1 ^ (jf & 7)) << 4));
          const uint4 g0 = *reinterpret_cast<const uint4*>(grow + ((c0 ^ (u & 7)) << 4));
          const uint4 g1 = *reinterpret_cast<const uint4*>(grow + ((c1 ^ (u & 7)) << 4));
          float d0 = 0.f, d1 = 0.f;
          d0 = fmaf(bf_lo(w0.x), bf_lo(g0.x), d0); d1 = fmaf(bf_hi(w0.x), bf_hi(g0.x), d1);
          d0 = fmaf(bf_lo(w0.y), bf_lo(g0.y), d0); d1 = fmaf(bf_hi(w0.y), bf_hi(g0.y), d1);
          d0 = fmaf(bf_lo(w0.z), bf_lo(g0.z), d0); d1 = fmaf(bf_hi(w0.z), bf_hi(g0.z), d1);
          d0 = fmaf(bf_lo(w0.w), bf_lo(g0.w), d0); d1 = fmaf(bf_hi(w0.w), bf_hi(g0.w), d1);
          d0 = fmaf(bf_lo(w1.x), bf_lo(g1.x), d0); d1 = fmaf(bf_hi(w1.x), bf_hi(g1.x), d1);
          d0 = fmaf(bf_lo(w1.y), bf_lo(g1.y), d0); d1 = fmaf(bf_hi(w1.y), bf_hi(g1.y), d1);
          d0 = fmaf(bf_lo(w1.z), bf_lo(g1.z), d0); d1 = fmaf(bf_hi(w1.z), bf_hi(g1.z), d1);
          d0 = fmaf(bf_lo(w1.w), bf_lo(g1.w), d0); d1 = fmaf(bf_hi(w1.w), bf_hi(g1.w), d1);
          float d = d0 + d1;
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          if (part == 0) daout[(u & 7) * 256 + ((u >> 3) * TPB + fb) * 64 + jf] = d;
          __syncwarp();
          if (lane == 0) mbar_arrive(r_empty + stg);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(da_ready);
        SB_STAMP(2);
        // ---- 3. owner: softmax backward and dq of the own utterances ----
        if (att_warp && n_own > 0) {
          mbar_wait_t(da_full, s & 1);
          SB_STAMP(3);
          if (ct == 0 && s + 1 < n_steps) mbar_expect_tx(da_full, SP_NC * n_own * TPB * 256);
          float da[TPB], ad = 0.f;
#pragma unroll
          for (int f = 0; f < TPB; ++f) {
            da[f] = 0.f;
            if (att_on) {
#pragma unroll
              for (int src = 0; src < SP_NC; ++src) da[f] += dain[src * 256 + (ua * TPB + f) * 64 + ja];
            }
            ad = fmaf(al_a[f], da[f], ad);
          }
          const float wd = warp_sum(ad);
          if (lane == 0) red[ct >> 5] = wd;
          named_bar(1, 256);
          if (att_on) {
            const float dot = red[ua * 2] + red[ua * 2 + 1];
#pragma unroll
            for (int f = 0; f < TPB; ++f) {
              const float de = ja + 64 * f < len_a ? al_a[f] * (da[f] - dot) : 0.f;
              if (ja + 64 * f < p.Tp) p.de[(size_t)ba * p.de_ldb + (size_t)t * p.de_ldt + ja + 64 * f] = de;
              des[(ua * TPB + f) * 64 + ja] = de;
            }
          }
          named_bar(1, 256);
        }
        // the psi~ stages are walked by every warp in ring order (only the owner warps read them)
        float s0 = 0.f, s1 = 0.f;
        const int mc = (2 * ja) & 63;                                          // column of m = 2 ja inside its k-block
        for (int i = 0; i < n_own * TPB; ++i, ++rpos) {
          const int o = i / TPB, fb = i % TPB;
          const int stg = rpos % SB_NSTAGE;
          mbar_wait_t(r_full + stg, (rpos / SB_NSTAGE) & 1);
          if (att_on && o == ua) {
            // dq_pre[m] for m = 2 ja, 2 ja + 1: column sums of de_j psi~[j, m] over the frames of this block
            const uint8_t* pt = Ring + stg * SP_STAGE + (ja >> 5) * 8192;      // k-block of m
#pragma unroll 8
            for (int j = 0; j < 64; ++j) {
              const uint32_t w = *reinterpret_cast<const uint32_t*>(pt + j * 128 + (((mc >> 3) ^ (j & 7)) << 4) + (mc & 7) * 2);
              const float dj = des[(ua * TPB + fb) * 64 + j];
              s0 = fmaf(dj, bf_lo(w), s0);
              s1 = fmaf(dj, bf_hi(w), s1);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(r_empty + stg);
          if (att_on && o == ua && fb == TPB - 1) {
            const float dq0 = s0 * (1.f - q_a0 * q_a0), dq1 = s1 * (1.f - q_a1 * q_a1);
            *reinterpret_cast<float2*>(p.dqpre + (size_t)ba * p.dq_ldb + (size_t)t * p.dq_ldt + 2 * ja) = make_float2(dq0, dq1);
            const __nv_bfloat162 pk = __floats2bfloat162_rn(dq0, dq1);
            *reinterpret_cast<__nv_bfloat162*>(dqimg + (ua * 2 + (ja >> 5)) * 128 + (((mc >> 3) ^ (sa_slot & 7)) << 4) + (mc & 7) * 2) = pk;
          }
        }
        if (att_warp && n_own > 0) {
          fence_proxy_async();
          named_bar(1, 256);
          if (ct == 0) mbar_arrive(dq_ready);
          SB_STAMP(4);
        }
      }
      // ---- 4. partial dh of all 256 units (this CTA's gate rows / query rows contracted) -> bf16 -> the units' owners ----
      if (s + 1 < n_steps) {
        mbar_wait_t(d_done, s & 1);
        SB_STAMP(5);
        tc_fence_after();
        const int sp = warp & 3, cgrp = cw >> 2;   // TMEM sub-partition = 32 units = one destination per half; 8 (4) columns
        constexpr int NCOL = NT / 4;
        uint32_t v[2][NCOL];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t ta = tmem + ((uint32_t)(sp * 32) << 16) + (uint32_t)(h * NT + cgrp * NCOL);
          if constexpr (NCOL == 8) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[h][0]), "=r"(v[h][1]), "=r"(v[h][2]), "=r"(v[h][3]), "=r"(v[h][4]), "=r"(v[h][5]), "=r"(v[h][6]), "=r"(v[h][7])
                         : "r"(ta) : "memory");
          } else {
            tmem_ld4(ta, v[h]);
          }
        }
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int c = 0; c < NCOL; ++c)
            *reinterpret_cast<__nv_bfloat16*>(dhout + (h * 4 + sp) * DHB + (cgrp * NCOL + c) * 64 + lane * 2) =
                __float2bfloat16_rn(__uint_as_float(v[h][c]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(dh_ready);
        SB_STAMP(6);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(tmem);
  cluster_sync_all();
}

// xin1 rows for the weight-gradient products: [emb(tok) ; context ; h1(t-1)].  One CTA per utterance: the attention maps of all
// steps sit in shared memory (transposed, 16 steps per 64-byte line), a thread owns context columns and accumulates 16 steps at a
// time over the utterance's valid frames.
constexpr int FX_TB = 16;
__global__ void __launch_bounds__(512) spell_fill_xin1_kernel(int U, int Tp, int E, int Sd, const float* __restrict__ alpha,
                                                              const float* __restrict__ enc, const int* __restrict__ lens,
                                                              const float* __restrict__ emb_w, const int* __restrict__ tok,
                                                              const float* __restrict__ h1, long long h1_ldb, long long h1_ldt,
                                                              float* __restrict__ xin1) {
  extern __shared__ float als[];            // [Tp][Up], Up = U rounded up to 16
  const int b = blockIdx.x, Up = (U + FX_TB - 1) / FX_TB * FX_TB, X1 = 2 * Sd + E;
  for (int i = threadIdx.x; i < Tp * Up; i += blockDim.x) {
    const int j = i / Up, t = i - j * Up;
    als[i] = t < U ? alpha[((size_t)b * U + t) * Tp + j] : 0.f;
  }
  __syncthreads();
  const int len = min(lens[b], Tp);
  const float* encb = enc + (size_t)b * Tp * E;
  float* xb = xin1 + (size_t)b * U * X1;
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    for (int t0 = 0; t0 < U; t0 += FX_TB) {
      float acc[FX_TB];
#pragma unroll
      for (int k = 0; k < FX_TB; ++k) acc[k] = 0.f;
      for (int j = 0; j < len; ++j) {
        const float e = __ldg(encb + (size_t)j * E + c);
        const float4* a4 = reinterpret_cast<const float4*>(als + j * Up + t0);
#pragma unroll
        for (int k4 = 0; k4 < FX_TB / 4; ++k4) {
          const float4 a = a4[k4];
          acc[4 * k4] = fmaf(a.x, e, acc[4 * k4]); acc[4 * k4 + 1] = fmaf(a.y, e, acc[4 * k4 + 1]);
          acc[4 * k4 + 2] = fmaf(a.z, e, acc[4 * k4 + 2]); acc[4 * k4 + 3] = fmaf(a.w, e, acc[4 * k4 + 3]);
        }
      }
#pragma unroll
      for (int k = 0; k < FX_TB; ++k)
        if (t0 + k < U) xb[(size_t)(t0 + k) * X1 + Sd + c] = acc[k];
    }
  }
  for (int i = threadIdx.x; i < U * Sd; i += blockDim.x) {
    const int t = i / Sd, k = i - t * Sd;
    xb[(size_t)t * X1 + k] = __ldg(emb_w + (size_t)tok[(size_t)b * U + t] * Sd + k);
    xb[(size_t)t * X1 + Sd + E + k] = t ? h1[(size_t)b * h1_ldb + (size_t)(t - 1) * h1_ldt + k] : 0.f;
  }
}

// exchange ring: internal, L2-resident scratch, one per (device, stream) that ever launched a decoder-loop kernel
struct SpRing { cudaStream_t st; int dev; uint8_t* buf; };
SpRing g_sp_rings[16];
int g_sp_nrings = 0;
uint8_t* sp_ring_for(cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < g_sp_nrings; ++i)
    if (g_sp_rings[i].st == st && g_sp_rings[i].dev == dev) return g_sp_rings[i].buf;
  if (g_sp_nrings == 16) return nullptr;
  uint8_t* b = nullptr;
  if (cudaMalloc(&b, (size_t)SP_RING * 160 * SB_SLOT) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  g_sp_rings[g_sp_nrings++] = {st, dev, b};
  return b;
}

long long* g_sp_dbg = nullptr;
long long* g_sp_dbg_bwd = nullptr;
int g_sp_dbg_mode = 0;
int g_sp_cap = -1;                         // co-resident clusters of the forward kernel (queried once)

template <typename Kern>
int sp_query(Kern kern, int smem_bytes = SP_SMEM) {
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(SP_NC, 1, 1);
  cfg.blockDim = dim3(SP_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = SP_NC;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess) { cudaGetLastError(); nc = 0; }
  return nc;
}
int sp_capacity() {
  if (g_sp_cap < 0) {
    const int q[10] = {sp_query(spell_cl_fwd_kernel<16, true, 1>), sp_query(spell_cl_fwd_kernel<32, true, 1>),
                       sp_query(spell_cl_fwd_kernel<16, false, 1>), sp_query(spell_cl_fwd_kernel<32, false, 1>),
                       sp_query(spell_cl_fwd_kernel<16, true, 4>),
                       sp_query(spell_cl_bwd_kernel<16, true, 1>, SB_SMEM), sp_query(spell_cl_bwd_kernel<32, true, 1>, SB_SMEM),
                       sp_query(spell_cl_bwd_kernel<16, false, 1>, SB_SMEM), sp_query(spell_cl_bwd_kernel<32, false, 1>, SB_SMEM),
                       sp_query(spell_cl_bwd_kernel<16, true, 4>, SB_SMEM)};
    g_sp_cap = q[0];
    for (int i = 1; i < 10; ++i)
      if (q[i] < g_sp_cap) g_sp_cap = q[i];
  }
  return g_sp_cap;
}
// clusters and utterances per cluster for a batch of B: as many co-resident clusters as there are (the step time grows with
// the utterances per cluster: their P tiles stream through every CTA)
void sp_split(int B, int* clusters, int* per) {
  const int cap = sp_capacity();
  int nc = B < cap ? B : cap;
  if (nc < 1) nc = 1;
  *per = (B + nc - 1) / nc;
  *clusters = (B + *per - 1) / *per;
}
int sp_tpb(int Tp) { return Tp <= SP_TP ? 1 : 4; }

}  // namespace

// 1 when the cluster-persistent decoder-step kernels cover these dimensions with every cluster co-resident
int spell_cl_supported(int B, int Tp, int E, int Sd, int M) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SSASR_SPELL_CL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  if (!on || Sd != SP_SD || M != SP_M || Tp < 1 || Tp > 4 * SP_TP || E % 8 != 0 || B < 1) return 0;
  if (sp_capacity() < 1) return 0;
  int clusters, per;
  sp_split(B, &clusters, &per);
  return per <= (sp_tpb(Tp) == 1 ? SP_MAXU : 8) ? 1 : 0;       // four frame blocks: at most one attention utterance per CTA
}

int spell_cl_fwd(cudaStream_t st, const SpellClFwdArgs& a) {
  SSASR_REQUIRE(a.t1 > a.t0 && a.t0 >= 0 && a.t1 <= a.U, "spell_cl_fwd: bad step range [%d, %d) of %d", a.t0, a.t1, a.U);
  SSASR_REQUIRE(sp_capacity() > 0, "spell_cl_fwd: the cluster kernel cannot be launched on this device");
  int clusters, per;
  sp_split(a.B, &clusters, &per);
  SSASR_REQUIRE(per <= SP_MAXU, "spell_cl_fwd: %d utterances do not fit %d co-resident clusters", a.B, sp_capacity());
  const bool att = a.xpre == nullptr;
  SpellClP p = {};
  p.B = a.B; p.U = a.U; p.Tp = a.Tp; p.t0 = a.t0; p.t1 = a.t1; p.per = per;
  p.gemb = a.gemb; p.tok = a.tok; p.tok_ld = a.tok_ld; p.enc_lens = a.enc_lens;
  p.act1 = a.act1; p.act1_ldb = a.act1_ldb; p.act1_ldt = a.act1_ldt;
  p.c1 = a.c1; p.c1_ldb = a.c1_ldb; p.c1_ldt = a.c1_ldt;
  p.h1 = a.h1; p.h1_ldb = a.h1_ldb; p.h1_ldt = a.h1_ldt;
  p.h1b = (__nv_bfloat16*)a.h1b; p.h1b_ldb = a.h1b_ldb; p.h1b_ldt = a.h1b_ldt;
  p.q = a.q; p.q_ldb = a.q_ldb; p.q_ldt = a.q_ldt;
  p.alpha = a.alpha; p.al_ldb = a.al_ldb; p.al_ldt = a.al_ldt;
  p.xpre = a.xpre; p.xpre_ldb = a.xpre_ldb; p.xpre_ldt = a.xpre_ldt;
  p.h2nd = a.h2nd; p.h2nd_ldb = a.h2nd_ldb; p.h2nd_ldt = a.h2nd_ldt; p.h2nd_toff = a.h2nd_toff;
  p.dbg = (att != (g_sp_dbg_mode != 0)) ? g_sp_dbg : nullptr;      // mode 1: stamps of the plain-recurrence launches instead
  p.ring = sp_ring_for(st);
  SSASR_REQUIRE(p.ring != nullptr, "spell_cl_fwd: cannot allocate the exchange ring");
  CUtensorMap tmW, tmPhi, tmP, tmPsi;
  int rc = make_tmap_bf16(&tmW, a.w1cat_bf, 4 * SP_SD, a.X1, a.X1, 128);
  if (rc) return rc;
  tmPhi = tmW; tmP = tmW; tmPsi = tmW;
  if (att) {
    rc = make_tmap_bf16(&tmPhi, a.phi_bf, SP_M, SP_SD, SP_SD, 128);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmP, a.P_bf, (long long)a.B * a.Tp, 4 * SP_SD, 4 * SP_SD, 64);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmPsi, a.psi_bf, (long long)a.B * a.Tp, SP_M, SP_M, 64);
    if (rc) return rc;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(SP_NC, 1, clusters);
  cfg.blockDim = dim3(SP_THREADS);
  cfg.dynamicSmemBytes = SP_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = SP_NC;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  ProfScope ps(F_SPELL_FWD, st);
  if (att && sp_tpb(a.Tp) == 4) {
    SSASR_REQUIRE(per <= 8, "spell_cl_fwd: T' = %d needs at most 8 utterances per cluster (got %d)", a.Tp, per);
    SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, spell_cl_fwd_kernel<16, true, 4>, tmW, tmPhi, tmP, tmPsi, a.K1, p));
  } else if (att) {
    if (per <= 16) SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, spell_cl_fwd_kernel<16, true, 1>, tmW, tmPhi, tmP, tmPsi, a.K1, p));
    else SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, spell_cl_fwd_kernel<32, true, 1>, tmW, tmPhi, tmP, tmPsi, a.K1, p));
  } else {
    if (per <= 16) SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, spell_cl_fwd_kernel<16, false, 1>, tmW, tmPhi, tmP, tmPsi, a.K1, p));
    else SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, spell_cl_fwd_kernel<32, false, 1>, tmW, tmPhi, tmP, tmPsi, a.K1, p));
  }
  return 0;
}

int spell_cl_bwd(cudaStream_t st, const SpellClBwdArgs& a) {
  SSASR_REQUIRE(a.t0 == 0 && a.t1 == a.U && a.U >= 1, "spell_cl_bwd: the backward chain runs over all %d steps", a.U);
  SSASR_REQUIRE(sp_capacity() > 0, "spell_cl_bwd: the cluster kernel cannot be launched on this device");
  int clusters, per;
  sp_split(a.B, &clusters, &per);
  SSASR_REQUIRE(per <= SP_MAXU, "spell_cl_bwd: %d utterances do not fit %d co-resident clusters", a.B, sp_capacity());
  const bool att = a.P_bf != nullptr;
  SpellClBwdP p = {};
  p.B = a.B; p.U = a.U; p.Tp = a.Tp; p.t0 = a.t0; p.t1 = a.t1; p.per = per; p.enc_lens = a.enc_lens;
  p.act = a.act; p.act_ldb = a.act_ldb; p.act_ldt = a.act_ldt;
  p.c = a.c; p.c_ldb = a.c_ldb; p.c_ldt = a.c_ldt;
  p.dh_in = a.dh_in; p.dh_ldb = a.dh_ldb; p.dh_ldt = a.dh_ldt;
  p.dgb = (__nv_bfloat16*)a.dgb; p.dgb_ldb = a.dgb_ldb; p.dgb_ldt = a.dgb_ldt;
  p.alpha = a.alpha; p.al_ldb = a.al_ldb; p.al_ldt = a.al_ldt;
  p.q = a.q; p.q_ldb = a.q_ldb; p.q_ldt = a.q_ldt;
  p.de = a.de; p.de_ldb = a.de_ldb; p.de_ldt = a.de_ldt;
  p.dqpre = a.dqpre; p.dq_ldb = a.dq_ldb; p.dq_ldt = a.dq_ldt;
  p.ring = sp_ring_for(st);
  p.dbg = (att != (g_sp_dbg_mode != 0)) ? g_sp_dbg_bwd : nullptr;
  SSASR_REQUIRE(p.ring != nullptr, "spell_cl_bwd: cannot allocate the exchange ring");
  CUtensorMap tmW, tmPhiS, tmP, tmPsi;
  int rc = make_tmap_bf16(&tmW, a.wcat_bf, 4 * SP_SD, a.X, a.X, 128);
  if (rc) return rc;
  tmPhiS = tmW; tmP = tmW; tmPsi = tmW;
  if (att) {
    rc = make_tmap_bf16(&tmPhiS, a.phi_bf, SP_M, SP_SD, SP_SD, 16);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmP, a.P_bf, (long long)a.B * a.Tp, 4 * SP_SD, 4 * SP_SD, 64);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmPsi, a.psi_bf, (long long)a.B * a.Tp, SP_M, SP_M, 64);
    if (rc) return rc;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(SP_NC, 1, clusters);
  cfg.blockDim = dim3(SP_THREADS);
  cfg.dynamicSmemBytes = SB_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = SP_NC;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  ProfScope ps(F_SPELL_BWD, st);
  if (att && sp_tpb(a.Tp) == 4) {
    SSASR_REQUIRE(per <= 8, "spell_cl_bwd: T' = %d needs at most 8 utterances per cluster (got %d)", a.Tp, per);
    SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, spell_cl_bwd_kernel<16, true, 4>, tmW, tmPhiS, tmP, tmPsi, a.Kcol, p));
  } else if (att) {
    if (per <= 16) SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, spell_cl_bwd_kernel<16, true, 1>, tmW, tmPhiS, tmP, tmPsi, a.Kcol, p));
    else SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, spell_cl_bwd_kernel<32, true, 1>, tmW, tmPhiS, tmP, tmPsi, a.Kcol, p));
  } else {
    if (per <= 16) SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, spell_cl_bwd_kernel<16, false, 1>, tmW, tmPhiS, tmP, tmPsi, a.Kcol, p));
    else SSASR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, spell_cl_bwd_kernel<32, false, 1>, tmW, tmPhiS, tmP, tmPsi, a.Kcol, p));
  }
  return 0;
}

int spell_fill_xin1(cudaStream_t st, int B, int U, int Tp, int E, int Sd, const float* alpha, const float* enc, const int* lens,
                    const float* emb_w, const int* tok, const float* h1, long long h1_ldb, long long h1_ldt, float* xin1) {
  const int Up = (U + FX_TB - 1) / FX_TB * FX_TB;
  const size_t smem = (size_t)Tp * Up * sizeof(float);
  SSASR_REQUIRE(smem <= 96 * 1024, "spell_fill_xin1: attention maps of one utterance (%d x %d) do not fit in shared memory", U, Tp);
  if (smem > 48 * 1024)
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(spell_fill_xin1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(F_POINTWISE, st);
  spell_fill_xin1_kernel<<<B, 512, smem, st>>>(U, Tp, E, Sd, alpha, enc, lens, emb_w, tok, h1, h1_ldb, h1_ldt, xin1);
  SSASR_LAUNCH_CHECK();
  return 0;
}

}  // namespace ssasr

extern "C" {
// debug: device buffer [steps][8] of clock64 stamps written by CTA 0 of the next cluster decoder-loop launches
void ssasr_spell_cl_set_debug(long long* dev_buf) { ssasr::g_sp_dbg = dev_buf; }
void ssasr_spell_cl_set_debug_bwd(long long* dev_buf) { ssasr::g_sp_dbg_bwd = dev_buf; }
void ssasr_spell_cl_set_debug_mode(int plain_recurrence) { ssasr::g_sp_dbg_mode = plain_recurrence; }
}

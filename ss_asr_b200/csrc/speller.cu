// Attend-and-spell decoder loop (fp32 path): attention energy / masked softmax / context, two stacked
// LSTM cells, character projection, next-token selection, and the full backward pass.
//
// Reference semantics: Attention.forward asr.py:343-392, Speller.forward asr.py:314-326, the decode loop
// of ASR.forward asr.py:65-110 (query = layer-1 state BEFORE this step's update), loss trainer.py:426-434.
// Everything stays on the device for all U steps (the reference syncs to the host every step,
// asr.py:103); attention maps are written straight into the stacked [B,U,T'] layout.
#include "common.cuh"
#include <stdlib.h>
#include <mutex>
#include <cuda_bf16.h>
#include <curand_kernel.h>

namespace ssasr {

int colsum(cudaStream_t st, const float* src, float* out, int R, int C, int ld, int accumulate);

// ------------------------------------------------------------------------------------------------
// attention step, forward.  One CTA (256 threads) per utterance.
// ------------------------------------------------------------------------------------------------
struct AttnFwd {
  int Tp, E, Sd, M;
  const float* h1prev; long long h1_ld;      // [B,Sd] rows (null at t == 0)
  const float* phi_w;                        // [M,Sd]
  const float* psi;                          // [B,Tp,M]  tanh(psi(enc))
  const float* enc;                          // [B,Tp,E]
  const int* enc_lens;                       // [B]
  const float* emb_w; const int* tok; long long tok_ld;   // embedding gather for the step input
  float* xin1; long long xin1_ld;            // row b: [emb(Sd) ; ctx(E) ; h1prev(Sd)]
  __nv_bfloat16* xin1b; long long xin1b_ld;  // optional bf16 copy of the same row (tensor-core GEMM operand)
  float* x3h; float* x3l; long long x3_ld;   // optional tf32 hi / lo split of the same row (tf32 x 3 GEMM operands)
  float* q; long long q_ld;                  // [B,M]
  float* alpha; long long alpha_ld;          // [B,Tp]
  // optional: the layer-1 cell of the PREVIOUS step fused into this kernel's prologue (row b is private to the CTA): h1prev
  // is then computed from the gate pre-activations instead of being read, and the cell's outputs are written from here
  float* cell_gates; long long cg_ld;        // [B,4Sd] pre-activations (col = unit*4 + gate) -> activations, in place
  const float* cell_cprev; long long ccp_ld; // [B,Sd] c1 of the step before, or null (zero state)
  float* cell_cout; long long cco_ld;        // [B,Sd] c1 of the previous step
  float* cell_hout; long long cho_ld;        // [B,Sd] h1 of the previous step, fp32 (layer-2 input row)
  __nv_bfloat16* cell_hb; long long chb_ld;  // [B,Sd] the same in bf16 (layer-2 GEMM operand), or null
  int skip_emb;                              // 1: the embedding part of the row is written by emb_rows_kernel (two-stream greedy loop)
};

// LSTM cell, one unit: gates (i,f,g,o pre-activations) -> activations in place; returns h, writes c
__device__ __forceinline__ float lstm_cell_unit(float4* gp, float cp, float* c_out) {
  const float4 g = *gp;
  float4 a;
  a.x = sigmoidf_acc(g.x); a.y = sigmoidf_acc(g.y); a.z = tanhf(g.z); a.w = sigmoidf_acc(g.w);
  const float c = a.y * cp + a.x * a.z;
  *gp = a;
  *c_out = c;
  return a.w * tanhf(c);
}

// hi = rn_tf32(v), lo = rn_tf32(v - hi): the operand split of the tf32 x 3 GEMM (gemm_tc.cu), fused into the producers
// lo == nullptr: `hi` is the bf16 operand of the tripled-K bf16 GEMM instead (rows of 3 ld: [hi | hi | lo], split3_bf16)
__device__ __forceinline__ void put_hi_lo(float* hi, float* lo, size_t b, long long ld, size_t col, float v) {
  if (lo == nullptr) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(hi) + b * 3 * ld + col;
    const __nv_bfloat16 h = __float2bfloat16(v);
    o[0] = h;
    o[ld] = h;
    o[2 * ld] = __float2bfloat16(v - __bfloat162float(h));
    return;
  }
  const size_t i = b * ld + col;
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  const float h = __uint_as_float(r);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v - h));
  hi[i] = h;
  lo[i] = __uint_as_float(r);
}
// 16-byte alignment of a row pointer / leading dimension pair
__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float s) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, s))));
}

// shared memory (floats): hs[Sd] qs[M] es[Tp] scratch[32] part[2*E]
static size_t attn_fwd_smem_bytes(int Sd, int M, int Tp, int E) { return (size_t)(Sd + M + ((Tp + 3) & ~3) + 32 + 2 * E) * sizeof(float); }

// rows[m0 .. m0+R) . x for a warp: lane holds I float4 of each row (row length K <= 128 * I), all R * I loads are issued
// before the first use; returns the R warp-reduced dot products in out[] (every lane holds all of them).
// R x I is chosen per shape so that a lane keeps 16 independent 128-bit loads in flight.
template <int R, int I>
__device__ __forceinline__ void warp_rows_dot(const float* __restrict__ rows, int ld, int K, int r0, int r_end, const float* xs /*smem*/,
                                              int lane, float* out) {
  float4 w[R][I];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int i = 0; i < I; ++i) {
      const int k = (i * 32 + lane) * 4;
      w[r][i] = (r0 + r < r_end && k < K) ? __ldg(reinterpret_cast<const float4*>(rows + (size_t)(r0 + r) * ld + k))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
  for (int r = 0; r < R; ++r) out[r] = 0.f;
#pragma unroll
  for (int i = 0; i < I; ++i) {
    const int k = (i * 32 + lane) * 4;
    if (k < K) {
      const float4 x4 = *reinterpret_cast<const float4*>(xs + k);
#pragma unroll
      for (int r = 0; r < R; ++r) out[r] = dot4(w[r][i], x4, out[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) out[r] = warp_sum(out[r]);
}
template <int R>
__device__ __forceinline__ float pick_lane(const float* v, int lane) {
  float s = v[0];
#pragma unroll
  for (int r = 1; r < R; ++r) s = (lane == r) ? v[r] : s;
  return s;
}

// Every phase is a batch of INDEPENDENT 128-bit loads issued before the first use (the per-utterance working set - phi 128 KB,
// psi~ 32 KB, encoder states 128 KB at the default sizes - streams from L2, so the phases are latency-bound unless many loads
// are in flight), followed by the arithmetic and a shuffle / shared-memory reduction.
__global__ void __launch_bounds__(256, 3) attn_fwd_kernel(AttnFwd a) {
  extern __shared__ float sm[];
  float* hs = sm;                 // [Sd]
  float* qs = hs + a.Sd;          // [M]
  float* es = qs + a.M;           // [Tp]
  float* scratch = es + ((a.Tp + 3) & ~3);     // [32] (kept 16-byte aligned together with everything behind it)
  float* part = scratch + 32;     // [2][E] context partial sums of the two frame halves
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  float* xrow = a.xin1 + (size_t)b * a.xin1_ld;
  const int tk = a.tok ? a.tok[(size_t)b * a.tok_ld] : 0;
  for (int k = tid; k < a.Sd; k += blockDim.x) {
    float h;
    if (a.cell_gates) {
      const float cp = a.cell_cprev ? a.cell_cprev[(size_t)b * a.ccp_ld + k] : 0.f;
      h = lstm_cell_unit(reinterpret_cast<float4*>(a.cell_gates + (size_t)b * a.cg_ld) + k, cp, a.cell_cout + (size_t)b * a.cco_ld + k);
      a.cell_hout[(size_t)b * a.cho_ld + k] = h;
      if (a.cell_hb) a.cell_hb[(size_t)b * a.chb_ld + k] = __float2bfloat16(h);
    } else {
      h = a.h1prev ? a.h1prev[(size_t)b * a.h1_ld + k] : 0.f;
    }
    hs[k] = h;
    xrow[a.Sd + a.E + k] = h;
    const float e = (a.emb_w && !a.skip_emb) ? a.emb_w[(size_t)tk * a.Sd + k] : 0.f;
    if (!a.skip_emb) xrow[k] = e;
    if (a.xin1b) {
      a.xin1b[(size_t)b * a.xin1b_ld + a.Sd + a.E + k] = __float2bfloat16(h);
      if (!a.skip_emb) a.xin1b[(size_t)b * a.xin1b_ld + k] = __float2bfloat16(e);
    }
    if (a.x3h) {
      put_hi_lo(a.x3h, a.x3l, (size_t)b, a.x3_ld, a.Sd + a.E + k, h);
      if (!a.skip_emb) put_hi_lo(a.x3h, a.x3l, (size_t)b, a.x3_ld, k, e);
    }
  }
  __syncthreads();
  const int len = a.enc_lens[b];
  const float* psib = a.psi + (size_t)b * a.Tp * a.M;
  const float* encb = a.enc + (size_t)b * a.Tp * a.E;
  const bool vec = ((a.Sd | a.M | a.E) & 3) == 0 && al16(a.phi_w) && al16(psib) && al16(encb);
  // ---- q = tanh(phi h): a warp takes 8 (Sd <= 256) or 4 rows of phi at a time ----
  if (vec && a.Sd <= 512) {
    auto put_q = [&](int m, float acc) {
      const float qv = tanhf(acc);
      qs[m] = qv;
      a.q[(size_t)b * a.q_ld + m] = qv;
    };
    if (a.Sd <= 256) {
      for (int m0 = warp * 8; m0 < a.M; m0 += nwarp * 8) {
        float acc[8];
        warp_rows_dot<8, 2>(a.phi_w, a.Sd, a.Sd, m0, a.M, hs, lane, acc);
        if (lane < 8 && m0 + lane < a.M) put_q(m0 + lane, pick_lane<8>(acc, lane));
      }
    } else {
      for (int m0 = warp * 4; m0 < a.M; m0 += nwarp * 4) {
        float acc[4];
        warp_rows_dot<4, 4>(a.phi_w, a.Sd, a.Sd, m0, a.M, hs, lane, acc);
        if (lane < 4 && m0 + lane < a.M) put_q(m0 + lane, pick_lane<4>(acc, lane));
      }
    }
  } else {
    for (int m = warp; m < a.M; m += nwarp) {
      float s = 0.f;
      for (int k = lane; k < a.Sd; k += 32) s = fmaf(a.phi_w[(size_t)m * a.Sd + k], hs[k], s);
      s = warp_sum(s);
      if (lane == 0) {
        const float qv = tanhf(s);
        qs[m] = qv;
        a.q[(size_t)b * a.q_ld + m] = qv;
      }
    }
  }
  __syncthreads();
  // ---- energies e_j = psi~_j . q: a warp takes 8 (M <= 128) or 4 frames at a time; frames >= len are never loaded ----
  if (vec && a.M <= 256) {
    if (a.M <= 128) {
      for (int j0 = warp * 8; j0 < a.Tp; j0 += nwarp * 8) {
        float acc[8];
        warp_rows_dot<8, 1>(psib, a.M, a.M, j0, len, qs, lane, acc);
        if (lane < 8 && j0 + lane < a.Tp) es[j0 + lane] = (j0 + lane < len) ? pick_lane<8>(acc, lane) : -INFINITY;
      }
    } else {
      for (int j0 = warp * 4; j0 < a.Tp; j0 += nwarp * 4) {
        float acc[4];
        warp_rows_dot<4, 2>(psib, a.M, a.M, j0, len, qs, lane, acc);
        if (lane < 4 && j0 + lane < a.Tp) es[j0 + lane] = (j0 + lane < len) ? pick_lane<4>(acc, lane) : -INFINITY;
      }
    }
  } else {
    for (int j = warp; j < a.Tp; j += nwarp) {
      float s = 0.f;
      if (j < len) {
        for (int m = lane; m < a.M; m += 32) s = fmaf(psib[(size_t)j * a.M + m], qs[m], s);
        s = warp_sum(s);
      } else {
        s = -INFINITY;
      }
      if (lane == 0) es[j] = s;
    }
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < a.Tp; j += blockDim.x) mx = fmaxf(mx, es[j]);
  mx = block_max(mx, scratch);
  float sum = 0.f;
  for (int j = tid; j < a.Tp; j += blockDim.x) {
    const float e = (es[j] == -INFINITY) ? 0.f : expf(es[j] - mx);
    es[j] = e;
    sum += e;
  }
  sum = block_sum(sum, scratch);
  const float inv = 1.0f / sum;
  __syncthreads();
  for (int j = tid; j < a.Tp; j += blockDim.x) {
    const float al = es[j] * inv;
    es[j] = al;
    a.alpha[(size_t)b * a.alpha_ld + j] = al;
  }
  __syncthreads();
  // ---- context c = sum_j alpha_j h_j: thread = (4 columns, frame parity), 16 frames of loads in flight ----
  if (vec) {
    const int half = tid >> 7, c4 = tid & 127;
    for (int c = c4 * 4; c < a.E; c += 512) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j0 = half; j0 < len; j0 += 32) {
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int j = j0 + 2 * u;
          v[u] = (j < len) ? __ldg(reinterpret_cast<const float4*>(encb + (size_t)j * a.E + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int j = j0 + 2 * u;
          const float al = (j < len) ? es[j] : 0.f;
          acc.x = fmaf(al, v[u].x, acc.x); acc.y = fmaf(al, v[u].y, acc.y);
          acc.z = fmaf(al, v[u].z, acc.z); acc.w = fmaf(al, v[u].w, acc.w);
        }
      }
      *reinterpret_cast<float4*>(part + half * a.E + c) = acc;
    }
    __syncthreads();
    for (int c = tid; c < a.E; c += blockDim.x) {
      const float sv = part[c] + part[a.E + c];
      xrow[a.Sd + c] = sv;
      if (a.xin1b) a.xin1b[(size_t)b * a.xin1b_ld + a.Sd + c] = __float2bfloat16(sv);
      if (a.x3h) put_hi_lo(a.x3h, a.x3l, (size_t)b, a.x3_ld, a.Sd + c, sv);
    }
  } else {
    for (int c = tid; c < a.E; c += blockDim.x) {
      float sv = 0.f;
      for (int j = 0; j < len; ++j) sv = fmaf(es[j], encb[(size_t)j * a.E + c], sv);
      xrow[a.Sd + c] = sv;
      if (a.xin1b) a.xin1b[(size_t)b * a.xin1b_ld + a.Sd + c] = __float2bfloat16(sv);
      if (a.x3h) put_hi_lo(a.x3h, a.x3l, (size_t)b, a.x3_ld, a.Sd + c, sv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attention step, backward.  One CTA per utterance.
// ------------------------------------------------------------------------------------------------
struct AttnBwd {
  int Tp, E, Sd, M;
  const float* dctx; long long dctx_ld;      // [B,E]
  const float* alpha; long long alpha_ld;    // [B,Tp]
  const float* q; long long q_ld;            // [B,M]
  const float* phi_w;                        // [M,Sd]
  const float* psi;                          // [B,Tp,M]
  const float* enc;                          // [B,Tp,E]
  const int* enc_lens;
  const float* dalpha; long long dalpha_ld;  // optional [B,Tp]: extra gradient arriving on the attention scores
  float* de; long long de_ld;                // [B,Tp] out: dL/d(energy) of this step (accumulated into denc/dpsi later)
  float* dqpre; long long dqpre_ld;          // [B,M] out
  float* dh1att;                             // [B,Sd] out
  // optional: the layer-1 cell backward of the PREVIOUS step (the one whose h1 was this step's query) fused into the epilogue:
  // dh1 = dh1att (local) + cb_dh_a + cb_dh_b; dh1att is then not written
  float* cb_act; long long cba_ld;           // [B,4Sd] activations -> gate gradients, in place
  const float* cb_c; long long cbc_ld;       // [B,Sd] c1 of that step
  const float* cb_cprev; long long cbp_ld;   // [B,Sd] c1 of the step before it, or null
  const float* cb_dh_a; long long cbda_ld;   // [B,Sd] dh1 from layer 2
  const float* cb_dh_b; long long cbdb_ld;   // [B,Sd] dh1 from the recurrent input of this step's layer-1 cell
  float* cb_dcstate;                         // [B,Sd] running dc1
  __nv_bfloat16* cb_dgb;                     // [B,4Sd] bf16 copy of the gate gradients, or null
};

// LSTM cell backward, one unit (act = activations i,f,g,o -> gate gradients in place)
__device__ __forceinline__ void lstm_cell_unit_bwd(float4* ap, float dh, float cv, float cp, float dcr, float* dc_out,
                                                   __nv_bfloat16* dgb4 /*4 bf16 of this unit, or null*/) {
  const float4 a = *ap;
  const float tc = tanhf(cv);
  const float dc = dh * a.w * (1.f - tc * tc) + dcr;
  float4 dg;
  dg.w = dh * tc * a.w * (1.f - a.w);
  dg.x = dc * a.z * a.x * (1.f - a.x);
  dg.z = dc * a.x * (1.f - a.z * a.z);
  dg.y = dc * cp * a.y * (1.f - a.y);
  *ap = dg;
  *dc_out = dc * a.y;
  if (dgb4) {
    __nv_bfloat162 b01 = __floats2bfloat162_rn(dg.x, dg.y), b23 = __floats2bfloat162_rn(dg.z, dg.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&b01);
    pk.y = *reinterpret_cast<uint32_t*>(&b23);
    *reinterpret_cast<uint2*>(dgb4) = pk;
  }
}

// shared memory (floats): dcs[E] als[Tp] das[Tp] dqs[M] scratch[32] part[max(8*M, 4*Sd)]
static size_t attn_bwd_smem_bytes(int Sd, int M, int Tp, int E) {
  const int pm = 8 * M > 4 * Sd ? 8 * M : 4 * Sd;
  return (size_t)(E + 2 * ((Tp + 3) & ~3) + M + 32 + pm) * sizeof(float);
}

__global__ void __launch_bounds__(256) attn_bwd_kernel(AttnBwd a) {
  extern __shared__ float sm[];
  float* dcs = sm;                 // [E]
  float* als = dcs + a.E;          // [Tp]  alpha, then de
  float* das = als + ((a.Tp + 3) & ~3);    // [Tp]  dalpha
  float* dqs = das + ((a.Tp + 3) & ~3);    // [M]
  float* scratch = dqs + a.M;      // [32]
  float* part = scratch + 32;      // [8][M] / [4][Sd] partial sums
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int len = a.enc_lens[b];
  for (int c = tid; c < a.E; c += blockDim.x) dcs[c] = a.dctx[(size_t)b * a.dctx_ld + c];
  for (int j = tid; j < a.Tp; j += blockDim.x) als[j] = a.alpha[(size_t)b * a.alpha_ld + j];
  __syncthreads();
  const float* encb = a.enc + (size_t)b * a.Tp * a.E;
  const float* psib = a.psi + (size_t)b * a.Tp * a.M;
  const bool vec = ((a.Sd | a.M | a.E) & 3) == 0 && al16(a.phi_w) && al16(psib) && al16(encb);
  // ---- dalpha_j = dctx . h_j: a warp takes 4 (E <= 512) or 2 frames at a time, 16 float4 per lane in flight ----
  if (vec && a.E <= 1024) {
    auto put_da = [&](int j, float dot) {
      das[j] = ((j < len) ? dot : 0.f) + ((a.dalpha && j < len) ? a.dalpha[(size_t)b * a.dalpha_ld + j] : 0.f);
    };
    if (a.E <= 512) {
      for (int j0 = warp * 4; j0 < a.Tp; j0 += nwarp * 4) {
        float acc[4];
        warp_rows_dot<4, 4>(encb, a.E, a.E, j0, len, dcs, lane, acc);
        if (lane < 4 && j0 + lane < a.Tp) put_da(j0 + lane, pick_lane<4>(acc, lane));
      }
    } else {
      for (int j0 = warp * 2; j0 < a.Tp; j0 += nwarp * 2) {
        float acc[2];
        warp_rows_dot<2, 8>(encb, a.E, a.E, j0, len, dcs, lane, acc);
        if (lane < 2 && j0 + lane < a.Tp) put_da(j0 + lane, pick_lane<2>(acc, lane));
      }
    }
  } else {
    for (int j = warp; j < a.Tp; j += nwarp) {
      float sv = 0.f;
      if (j < len) {
        for (int c = lane; c < a.E; c += 32) sv = fmaf(dcs[c], encb[(size_t)j * a.E + c], sv);
        sv = warp_sum(sv);
      }
      if (lane == 0) das[j] = sv + ((a.dalpha && j < len) ? a.dalpha[(size_t)b * a.dalpha_ld + j] : 0.f);
    }
  }
  __syncthreads();
  float dot = 0.f;
  for (int j = tid; j < len; j += blockDim.x) dot = fmaf(als[j], das[j], dot);
  dot = block_sum(dot, scratch);
  __syncthreads();
  for (int j = tid; j < a.Tp; j += blockDim.x) {
    const float de = (j < len) ? als[j] * (das[j] - dot) : 0.f;
    als[j] = de;
    a.de[(size_t)b * a.de_ld + j] = de;
  }
  __syncthreads();
  // ---- dq_m = (sum_j de_j psi~_jm)(1 - q_m^2): thread = (4 columns of psi~, one of 8 frame groups) ----
  if (vec && a.M <= 128) {
    const int grp = tid >> 5, m = lane * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < a.M) {
      for (int j0 = grp; j0 < len; j0 += 64) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + 8 * u;
          v[u] = (j < len) ? __ldg(reinterpret_cast<const float4*>(psib + (size_t)j * a.M + m)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + 8 * u;
          const float de = (j < len) ? als[j] : 0.f;
          acc.x = fmaf(de, v[u].x, acc.x); acc.y = fmaf(de, v[u].y, acc.y);
          acc.z = fmaf(de, v[u].z, acc.z); acc.w = fmaf(de, v[u].w, acc.w);
        }
      }
      *reinterpret_cast<float4*>(part + grp * a.M + m) = acc;
    }
    __syncthreads();
    for (int mm = tid; mm < a.M; mm += blockDim.x) {
      float sv = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) sv += part[g * a.M + mm];
      const float qv = a.q[(size_t)b * a.q_ld + mm];
      const float dq = sv * (1.f - qv * qv);
      dqs[mm] = dq;
      a.dqpre[(size_t)b * a.dqpre_ld + mm] = dq;
    }
  } else {
    for (int m = tid; m < a.M; m += blockDim.x) {
      const float qv = a.q[(size_t)b * a.q_ld + m];
      float sv = 0.f;
      for (int j = 0; j < len; ++j) sv = fmaf(als[j], psib[(size_t)j * a.M + m], sv);
      const float dq = sv * (1.f - qv * qv);
      dqs[m] = dq;
      a.dqpre[(size_t)b * a.dqpre_ld + m] = dq;
    }
  }
  __syncthreads();
  // unit k of this utterance: publish dh1att, or run the previous step's layer-1 cell backward with it right here
  auto finish = [&](int k, float dh_att) {
    if (!a.cb_act) {
      a.dh1att[(size_t)b * a.Sd + k] = dh_att;
      return;
    }
    // same summation order as cell_bwd_kernel: dh_a + dh_b + dh_c
    float dh = a.cb_dh_a[(size_t)b * a.cbda_ld + k];
    dh += a.cb_dh_b[(size_t)b * a.cbdb_ld + k];
    dh += dh_att;
    const size_t i = (size_t)b * a.Sd + k;
    lstm_cell_unit_bwd(reinterpret_cast<float4*>(a.cb_act + (size_t)b * a.cba_ld) + k, dh, a.cb_c[(size_t)b * a.cbc_ld + k],
                       a.cb_cprev ? a.cb_cprev[(size_t)b * a.cbp_ld + k] : 0.f, a.cb_dcstate[i], a.cb_dcstate + i,
                       a.cb_dgb ? a.cb_dgb + i * 4 : nullptr);
  };
  // ---- dh = phi^T dq: thread = (4 columns of phi, one of 4 row groups) ----
  if (vec && a.Sd <= 256) {
    const int grp = tid >> 6, k = (tid & 63) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < a.Sd) {
      for (int m0 = grp; m0 < a.M; m0 += 64) {
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int m = m0 + 4 * u;
          v[u] = (m < a.M) ? __ldg(reinterpret_cast<const float4*>(a.phi_w + (size_t)m * a.Sd + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int m = m0 + 4 * u;
          const float dq = (m < a.M) ? dqs[m] : 0.f;
          acc.x = fmaf(dq, v[u].x, acc.x); acc.y = fmaf(dq, v[u].y, acc.y);
          acc.z = fmaf(dq, v[u].z, acc.z); acc.w = fmaf(dq, v[u].w, acc.w);
        }
      }
      *reinterpret_cast<float4*>(part + grp * a.Sd + k) = acc;
    }
    __syncthreads();
    for (int kk = tid; kk < a.Sd; kk += blockDim.x)
      finish(kk, part[kk] + part[a.Sd + kk] + part[2 * a.Sd + kk] + part[3 * a.Sd + kk]);
  } else {
    for (int k = tid; k < a.Sd; k += blockDim.x) {
      float sv = 0.f;
      for (int m = 0; m < a.M; ++m) sv = fmaf(dqs[m], a.phi_w[(size_t)m * a.Sd + k], sv);
      finish(k, sv);
    }
  }
}

// After the loop: out[b,j,:] = sum_t w[b,t,j] * v[b,t,:]   (denc from alpha x dctx, dpsi from de x q).
// One CTA per (utterance, 16-frame block): w tile in smem; thread = (4 columns of v, one group of the block's frames), v
// streamed as 128-bit loads, four decoding steps in flight per thread.  D % 4 == 0 and D / 4 <= 256.
constexpr int OA_FB = 16;     // frames per CTA
constexpr int OA_FPG = 8;     // at most this many frames per thread
__global__ void __launch_bounds__(256) attn_outer_accum_kernel(int U, int Tp, int D, const float* __restrict__ w /*[B,U,Tp]*/,
                                                               const float* __restrict__ v, long long v_ld_t, long long v_ld_b,
                                                               float* __restrict__ out /*[B,Tp,D]*/, const int* __restrict__ lens) {
  extern __shared__ float ws[];     // [U][OA_FB]
  const int b = blockIdx.y, j0 = blockIdx.x * OA_FB;
  const int len = lens[b];
  float* ob = out + ((size_t)b * Tp + j0) * D;
  const int nf = min(OA_FB, Tp - j0);      // frames of this block
  if (j0 >= len) {                  // fully padded frame block: gradient is exactly zero
    for (int i = threadIdx.x; i < nf * D; i += blockDim.x) ob[i] = 0.f;
    return;
  }
  for (int i = threadIdx.x; i < U * OA_FB; i += blockDim.x) {
    const int t = i / OA_FB, jj = i % OA_FB;
    ws[i] = (jj < nf) ? w[((size_t)b * U + t) * Tp + j0 + jj] : 0.f;
  }
  __syncthreads();
  const int d4 = D >> 2;                           // float4 columns
  const int groups = max(1, min((int)blockDim.x / d4, OA_FB));   // frame groups that fit beside the columns
  const int fpg = (OA_FB + groups - 1) / groups;   // frames per group (<= OA_FPG by the launcher's contract)
  const int c4 = threadIdx.x % d4, g = threadIdx.x / d4;
  if (g >= groups) return;
  const int f0 = g * fpg;
  const float* vb = v + (size_t)b * v_ld_b + (size_t)c4 * 4;
  float4 acc[OA_FPG];
#pragma unroll
  for (int jj = 0; jj < OA_FPG; ++jj) acc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t0 = 0; t0 < U; t0 += 4) {
    float4 d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      d[k] = (t0 + k < U) ? __ldg(reinterpret_cast<const float4*>(vb + (size_t)(t0 + k) * v_ld_t)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (t0 + k >= U) break;
      const float* wr = ws + (t0 + k) * OA_FB + f0;
#pragma unroll
      for (int jj = 0; jj < OA_FPG; ++jj) {
        if (jj < fpg) {
          const float wv = wr[jj];
          acc[jj].x = fmaf(wv, d[k].x, acc[jj].x); acc[jj].y = fmaf(wv, d[k].y, acc[jj].y);
          acc[jj].z = fmaf(wv, d[k].z, acc[jj].z); acc[jj].w = fmaf(wv, d[k].w, acc[jj].w);
        }
      }
    }
  }
#pragma unroll
  for (int jj = 0; jj < OA_FPG; ++jj)
    if (jj < fpg && f0 + jj < nf) *reinterpret_cast<float4*>(ob + (size_t)(f0 + jj) * D + (size_t)c4 * 4) = acc[jj];
}
// generic fallback (any D): one CTA per (utterance, 8-frame block), scalar columns
__global__ void __launch_bounds__(256) attn_outer_accum8_kernel(int U, int Tp, int D, const float* __restrict__ w /*[B,U,Tp]*/,
                                                                const float* __restrict__ v, long long v_ld_t, long long v_ld_b,
                                                                float* __restrict__ out /*[B,Tp,D]*/, const int* __restrict__ lens) {
  extern __shared__ float ws[];     // [U][8]
  const int b = blockIdx.y, j0 = blockIdx.x * 8;
  const int len = lens[b];
  float* ob = out + ((size_t)b * Tp + j0) * D;
  if (j0 >= len) {                  // fully padded frame block: gradient is exactly zero
    for (int i = threadIdx.x; i < 8 * D && j0 + i / D < Tp; i += blockDim.x) ob[i] = 0.f;
    return;
  }
  for (int i = threadIdx.x; i < U * 8; i += blockDim.x) {
    const int t = i >> 3, jj = i & 7;
    ws[i] = (j0 + jj < Tp) ? w[((size_t)b * U + t) * Tp + j0 + jj] : 0.f;
  }
  __syncthreads();
  const float* vb = v + (size_t)b * v_ld_b;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) acc[jj] = 0.f;
    for (int t = 0; t < U; ++t) {
      const float d = vb[(size_t)t * v_ld_t + c];
      const float4 w0 = *reinterpret_cast<const float4*>(ws + t * 8), w1 = *reinterpret_cast<const float4*>(ws + t * 8 + 4);
      acc[0] = fmaf(w0.x, d, acc[0]); acc[1] = fmaf(w0.y, d, acc[1]); acc[2] = fmaf(w0.z, d, acc[2]); acc[3] = fmaf(w0.w, d, acc[3]);
      acc[4] = fmaf(w1.x, d, acc[4]); acc[5] = fmaf(w1.y, d, acc[5]); acc[6] = fmaf(w1.z, d, acc[6]); acc[7] = fmaf(w1.w, d, acc[7]);
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
      if (j0 + jj < Tp) ob[(size_t)jj * D + c] = acc[jj];
  }
}
// out[b,j,:] = sum_t w[b,t,j] * v[b,t,:] on stream st
static int attn_outer_accum(cudaStream_t st, int B, int U, int Tp, int D, const float* w, const float* v, long long v_ld_t, long long v_ld_b,
                            float* out, const int* lens) {
  const int d4 = D / 4;
  const bool fast = D % 4 == 0 && d4 >= 32 && d4 <= 256 && 256 % d4 == 0 && (v_ld_t % 4) == 0 && (v_ld_b % 4) == 0 &&
                    (reinterpret_cast<uintptr_t>(v) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                    (OA_FB + (256 / d4 > OA_FB ? OA_FB : 256 / d4) - 1) / (256 / d4 > OA_FB ? OA_FB : 256 / d4) <= OA_FPG;
  ProfScope ps(F_ATTN_BWD, st);
  if (fast)
    attn_outer_accum_kernel<<<dim3((Tp + OA_FB - 1) / OA_FB, B), 256, (size_t)U * OA_FB * sizeof(float), st>>>(U, Tp, D, w, v, v_ld_t, v_ld_b,
                                                                                                              out, lens);
  else
    attn_outer_accum8_kernel<<<dim3((Tp + 7) / 8, B), 256, (size_t)U * 8 * sizeof(float), st>>>(U, Tp, D, w, v, v_ld_t, v_ld_b, out, lens);
  SSASR_LAUNCH_CHECK();
  return 0;
}

// embedding part of the step-input rows (two-stream greedy loop: the token is selected on the layer-2 stream while the
// attention step that writes the rest of the row already runs): xin1[b, 0:Sd] = emb[tok[b]] (+ the split-operand copy)
__global__ void emb_rows_kernel(int B, int Sd, const float* __restrict__ emb_w, const int* __restrict__ tok, long long tok_ld,
                                float* __restrict__ xin1, long long xin1_ld, float* __restrict__ x3h, float* __restrict__ x3l,
                                long long x3_ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Sd) return;
  const int b = i / Sd, k = i % Sd;
  const float e = emb_w[(size_t)tok[(size_t)b * tok_ld] * Sd + k];
  xin1[(size_t)b * xin1_ld + k] = e;
  if (x3h) put_hi_lo(x3h, x3l, (size_t)b, x3_ld, k, e);
}

// ------------------------------------------------------------------------------------------------
// LSTM cell pointwise (gates interleaved: col = unit*4 + gate)
// ------------------------------------------------------------------------------------------------
__global__ void cell_fwd_kernel(int B, int S, float* __restrict__ gates, long long g_ld, const float* __restrict__ cprev,
                                long long cp_ld, float* __restrict__ cout, long long c_ld, float* __restrict__ hout,
                                long long h_ld, const float* __restrict__ cp_src, long long cps_ld, float* __restrict__ cp_dst,
                                long long cpd_ld, __nv_bfloat16* __restrict__ hb_out, long long hb_ld, int hb_cp_off,
                                float* __restrict__ x3h = nullptr, float* __restrict__ x3l = nullptr, long long x3_ld = 0,
                                float* __restrict__ h_dst2 = nullptr, long long hd2_ld = 0,
                                __nv_bfloat16* __restrict__ hb_dst2 = nullptr, long long hbd2_ld = 0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * S) return;
  const int b = i / S, u = i % S;
  float4* gp = reinterpret_cast<float4*>(gates + (size_t)b * g_ld) + u;
  const float cp = cprev ? cprev[(size_t)b * cp_ld + u] : 0.f;
  const float h = lstm_cell_unit(gp, cp, cout + (size_t)b * c_ld + u);
  hout[(size_t)b * h_ld + u] = h;
  if (hb_out) hb_out[(size_t)b * hb_ld + u] = __float2bfloat16(h);
  if (x3h) put_hi_lo(x3h, x3l, (size_t)b, x3_ld, u, h);
  if (h_dst2) h_dst2[(size_t)b * hd2_ld + u] = h;
  if (hb_dst2) hb_dst2[(size_t)b * hbd2_ld + u] = __float2bfloat16(h);
  if (cp_dst) {
    const float v = cp_src ? cp_src[(size_t)b * cps_ld + u] : 0.f;
    cp_dst[(size_t)b * cpd_ld + u] = v;
    if (hb_out) hb_out[(size_t)b * hb_ld + hb_cp_off + u] = __float2bfloat16(v);
    if (x3h) put_hi_lo(x3h, x3l, (size_t)b, x3_ld, hb_cp_off + u, v);
  }
}

__global__ void cell_bwd_kernel(int B, int S, float* __restrict__ act, long long a_ld, const float* __restrict__ c, long long c_ld,
                                const float* __restrict__ cprev, long long cp_ld, const float* __restrict__ dh_a, long long da_ld,
                                const float* __restrict__ dh_b, long long db_ld, const float* __restrict__ dh_c, long long dc_ld,
                                float* __restrict__ dcstate, int first, __nv_bfloat16* __restrict__ dgb /*[B,4S] or null*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * S) return;
  const int b = i / S, u = i % S;
  float dh = dh_a[(size_t)b * da_ld + u];
  if (dh_b) dh += dh_b[(size_t)b * db_ld + u];
  if (dh_c) dh += dh_c[(size_t)b * dc_ld + u];
  lstm_cell_unit_bwd(reinterpret_cast<float4*>(act + (size_t)b * a_ld) + u, dh, c[(size_t)b * c_ld + u],
                     cprev ? cprev[(size_t)b * cp_ld + u] : 0.f, first ? 0.f : dcstate[i], dcstate + i,
                     dgb ? dgb + (size_t)b * 4 * S + (size_t)u * 4 : nullptr);
}

__global__ void emb_grad_add_kernel(int B, int Sd, const float* __restrict__ demb, long long ld, const int* __restrict__ tok,
                                    long long tok_ld, float* __restrict__ gemb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Sd) return;
  const int b = i / Sd, k = i % Sd;
  atomicAdd(gemb + (size_t)tok[(size_t)b * tok_ld] * Sd + k, demb[(size_t)b * ld + k]);
}

__global__ void dtanh_inplace_kernel(float* __restrict__ d, const float* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = y[i];
    d[i] *= (1.f - v * v);
  }
}

// Character projection (asr.py:79 char_trans) fused with the next-token selection: one warp per utterance, the [C,Sd]
// weight matrix staged once per CTA in shared memory, fp32 throughout.  Replaces a 16-CTA SIMT GEMM + pick launch per step
// of greedy decoding / sampling.  mode 0: logits only, 1: argmax (first max index), 2: sample from softmax (Philox).
constexpr int LP_UPW = 1;      // utterances per warp (8 per CTA): the projection is shared-memory-bandwidth bound per SM, so spread wide
__host__ __device__ inline int lp_cp(int C) { return C | 1; }     // odd row pitch of the transposed weights: conflict-free staging
static size_t lp_smem_bytes(int C, int Sd) { return ((size_t)Sd * lp_cp(C) + 8 * (size_t)Sd + C) * sizeof(float); }

// lane = class (C <= 64: classes lane and lane + 32): no cross-lane reduction in the projection, h broadcast from smem
__global__ void __launch_bounds__(256) logits_pick_kernel(int B, int C, int Sd, const float* __restrict__ h, long long h_ld,
                                                          const float* __restrict__ wc, const float* __restrict__ bc,
                                                          float* __restrict__ logits, long long l_ld, int mode,
                                                          unsigned long long seed, unsigned long long step,
                                                          int* __restrict__ tok_out, long long tok_ld) {
  extern __shared__ float sm[];
  const int Cp = lp_cp(C);
  float* wt = sm;                           // [Sd][Cp]  transposed weights
  float* hs = wt + (size_t)Sd * Cp;         // [8][Sd]   one h row per warp
  float* bs = hs + 8 * (size_t)Sd;          // [C]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = C * Sd;
  if ((Sd & 3) == 0 && (reinterpret_cast<uintptr_t>(wc) & 15) == 0) {
#pragma unroll 4
    for (int i = tid * 4; i < n; i += 1024) {
      const int c = i / Sd, k = i - c * Sd;
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(wc + i));
      wt[(size_t)k * Cp + c] = w4.x; wt[(size_t)(k + 1) * Cp + c] = w4.y;
      wt[(size_t)(k + 2) * Cp + c] = w4.z; wt[(size_t)(k + 3) * Cp + c] = w4.w;
    }
  } else {
    for (int i = tid; i < n; i += 256) {
      const int c = i / Sd, k = i - c * Sd;
      wt[(size_t)k * Cp + c] = __ldg(wc + i);
    }
  }
  for (int c = tid; c < C; c += 256) bs[c] = bc ? __ldg(bc + c) : 0.f;
  __syncthreads();
  float* hrow = hs + (size_t)warp * Sd;
  const int c0 = lane, c1 = lane + 32;
  for (int u = 0; u < LP_UPW; ++u) {
    const int b = (blockIdx.x * 8 + warp) * LP_UPW + u;
    if (b >= B) return;
    const float* hb = h + (size_t)b * h_ld;
    for (int k = lane; k < Sd; k += 32) hrow[k] = __ldg(hb + k);
    __syncwarp();
    float s0 = 0.f, s1 = 0.f;
    if (c1 < C) {
#pragma unroll 8
      for (int k = 0; k < Sd; ++k) {
        const float hv = hrow[k];
        s0 = fmaf(hv, wt[(size_t)k * Cp + c0], s0);
        s1 = fmaf(hv, wt[(size_t)k * Cp + c1], s1);
      }
    } else if (c0 < C) {
#pragma unroll 8
      for (int k = 0; k < Sd; ++k) s0 = fmaf(hrow[k], wt[(size_t)k * Cp + c0], s0);
    }
    float best_v = -INFINITY;
    int best = 0x7fffffff;
    if (c0 < C) {
      s0 += bs[c0];
      logits[(size_t)b * l_ld + c0] = s0;
      best_v = s0; best = c0;
    }
    if (c1 < C) {
      s1 += bs[c1];
      logits[(size_t)b * l_ld + c1] = s1;
      if (s1 > best_v) { best_v = s1; best = c1; }
    }
    if (mode != 0) {
      // argmax with torch.argmax tie-breaking (first maximal index)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best, o);
        if (ov > best_v || (ov == best_v && oi < best)) { best_v = ov; best = oi; }
      }
      if (mode == 2) {      // sample from softmax (Philox; the one intentionally non-bit-reproducible branch, asr.py:97)
        const float mx = best_v;
        const float e0 = c0 < C ? expf(s0 - mx) : 0.f, e1 = c1 < C ? expf(s1 - mx) : 0.f;
        // inclusive prefix sums in class order: classes 0..31 first, then 32..63
        float p0 = e0, p1 = e1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float a0 = __shfl_up_sync(0xffffffffu, p0, o), a1 = __shfl_up_sync(0xffffffffu, p1, o);
          if (lane >= o) { p0 += a0; p1 += a1; }
        }
        const float tot0 = __shfl_sync(0xffffffffu, p0, 31), tot = tot0 + __shfl_sync(0xffffffffu, p1, 31);
        p1 += tot0;
        curandStatePhilox4_32_10_t rs;
        curand_init(seed, (unsigned long long)b, step, &rs);
        const float r = curand_uniform(&rs) * tot;
        // first class whose inclusive prefix reaches r
        int cand = 0x7fffffff;
        if (c0 < C && r <= p0) cand = c0;
        else if (c1 < C && r <= p1) cand = c1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        best = cand == 0x7fffffff ? C - 1 : cand;
      }
      if (lane == 0) tok_out[(size_t)b * tok_ld] = best;
    }
    __syncwarp();
  }
}

// greedy decoding: number of utterances whose selected tokens tok[b, 1 .. upto] do not contain `eos` yet (a warp per utterance)
__global__ void count_unfinished_kernel(int B, const int* __restrict__ tok, long long tok_ld, int upto, int eos, int* __restrict__ out) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  bool hit = false;
  for (int t = 1 + lane; t <= upto; t += 32) hit = hit || (tok[(size_t)b * tok_ld + t] == eos);
  if (__ballot_sync(0xffffffffu, hit) == 0u && lane == 0) atomicAdd(out, 1);
}

// next-token selection from logits row (C <= 1024): mode 1 = argmax (first max index, torch.argmax),
// mode 2 = sample from softmax (Philox; the one intentionally non-bit-reproducible branch, asr.py:97)
__global__ void pick_token_kernel(int B, int C, const float* __restrict__ logits, long long ld, int mode, unsigned long long seed,
                                  unsigned long long step, int* __restrict__ tok_out, long long tok_ld) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* l = logits + (size_t)b * ld;
  int best = 0;
  float mx = l[0];
  for (int c = 1; c < C; ++c)
    if (l[c] > mx) { mx = l[c]; best = c; }
  if (mode == 2) {
    curandStatePhilox4_32_10_t rs;
    curand_init(seed, (unsigned long long)b, step, &rs);
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(l[c] - mx);
    const float r = curand_uniform(&rs) * sum;
    float acc = 0.f;
    best = C - 1;
    for (int c = 0; c < C; ++c) {
      acc += expf(l[c] - mx);
      if (r <= acc) { best = c; break; }
    }
  }
  tok_out[(size_t)b * tok_ld] = best;
}

// ------------------------------------------------------------------------------------------------
// greedy step with the character LM of ASR.decode (asr.py:153-162, charlm.py:46-57): one CTA per utterance runs
// Embedding -> GRUCell -> GRUCell -> Linear, combines log_softmax(asr) + w * log_softmax(lm) and takes the argmax.
// Weights are passed transposed ([in, out]) so that the per-unit dot products read coalesced columns.
// ------------------------------------------------------------------------------------------------
struct LmStep {
  int C, H;
  const float *emb, *w1i, *w1h, *b1i, *b1h, *w2i, *w2h, *b2i, *b2h, *wo, *bo;   // w*: [H,3H] transposed, wo: [H,C]
  float *h1, *h2;                  // [B,H] state, updated in place
  float weight;
  const float* logits; long long logits_ld;
  const int* tok_in; int* tok_out; long long tok_ld;
  float* lm_out;                   // lm_step_kernel: [B,C] language-model logits of the step (lm_mix_pick_kernel reads them)
};

__device__ __forceinline__ float gru_unit(const float* __restrict__ x, const float* __restrict__ h, const float* __restrict__ wi,
                                          const float* __restrict__ wh, const float* __restrict__ bi, const float* __restrict__ bh,
                                          int H, int j) {
  float ir = bi[j], iz = bi[H + j], in = bi[2 * H + j], hr = bh[j], hz = bh[H + j], hn = bh[2 * H + j];
  for (int k = 0; k < H; ++k) {
    const float xv = x[k], hv = h[k];
    const float* wik = wi + (size_t)k * 3 * H;
    const float* whk = wh + (size_t)k * 3 * H;
    ir = fmaf(xv, wik[j], ir); iz = fmaf(xv, wik[H + j], iz); in = fmaf(xv, wik[2 * H + j], in);
    hr = fmaf(hv, whk[j], hr); hz = fmaf(hv, whk[H + j], hz); hn = fmaf(hv, whk[2 * H + j], hn);
  }
  const float r = sigmoidf_acc(ir + hr), z = sigmoidf_acc(iz + hz);
  const float n = tanhf(in + r * hn);
  return (1.f - z) * n + z * h[j];
}

__global__ void lm_pick_kernel(LmStep a) {
  extern __shared__ float sm[];
  float* x = sm;                 // [H]
  float* h1 = x + a.H;           // [H]
  float* h2 = h1 + a.H;          // [H]
  float* h1n = h2 + a.H;         // [H]
  float* h2n = h1n + a.H;        // [H]
  float* lm = h2n + a.H;         // [C]
  float* scratch = lm + a.C;     // [32]
  const int b = blockIdx.x, tid = threadIdx.x, H = a.H, C = a.C;
  const int tok = a.tok_in[(size_t)b * a.tok_ld];
  for (int j = tid; j < H; j += blockDim.x) {
    x[j] = a.emb[(size_t)tok * H + j];
    h1[j] = a.h1[(size_t)b * H + j];
    h2[j] = a.h2[(size_t)b * H + j];
  }
  __syncthreads();
  for (int j = tid; j < H; j += blockDim.x) h1n[j] = gru_unit(x, h1, a.w1i, a.w1h, a.b1i, a.b1h, H, j);
  __syncthreads();
  for (int j = tid; j < H; j += blockDim.x) h2n[j] = gru_unit(h1n, h2, a.w2i, a.w2h, a.b2i, a.b2h, H, j);
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    float s = a.bo[c];
    for (int k = 0; k < H; ++k) s = fmaf(h2n[k], a.wo[(size_t)k * C + c], s);
    lm[c] = s;
  }
  for (int j = tid; j < H; j += blockDim.x) {
    a.h1[(size_t)b * H + j] = h1n[j];
    a.h2[(size_t)b * H + j] = h2n[j];
  }
  __syncthreads();
  const float* lg = a.logits + (size_t)b * a.logits_ld;
  float ml = -INFINITY, ma = -INFINITY;
  for (int c = tid; c < C; c += blockDim.x) { ml = fmaxf(ml, lm[c]); ma = fmaxf(ma, lg[c]); }
  ml = block_max(ml, scratch);
  ma = block_max(ma, scratch);
  float sl = 0.f, sa = 0.f;
  for (int c = tid; c < C; c += blockDim.x) { sl += expf(lm[c] - ml); sa += expf(lg[c] - ma); }
  sl = block_sum(sl, scratch);
  sa = block_sum(sa, scratch);
  const float lse_l = ml + logf(sl), lse_a = ma + logf(sa);
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) lm[c] = (lg[c] - lse_a) + a.weight * (lm[c] - lse_l);
  __syncthreads();
  if (tid == 0) {
    int best = 0;
    float mx = lm[0];
    for (int c = 1; c < C; ++c)
      if (lm[c] > mx) { mx = lm[c]; best = c; }
    a.tok_out[(size_t)b * a.tok_ld] = best;
  }
}

// The same step in two kernels for the two-stream greedy loop: the CharLM recurrences of step t depend on token t only (not on
// the attention / speller state), so they run on a THIRD stream under the attention step and the gate products of step t;
// what is left on the layer-2 stream after the character projection is the mix + argmax.
__global__ void lm_step_kernel(LmStep a) {
  extern __shared__ float sm[];
  float* x = sm;                 // [H]
  float* h1 = x + a.H;           // [H]
  float* h2 = h1 + a.H;          // [H]
  float* h1n = h2 + a.H;         // [H]
  float* h2n = h1n + a.H;        // [H]
  const int b = blockIdx.x, tid = threadIdx.x, H = a.H, C = a.C;
  const int tok = a.tok_in[(size_t)b * a.tok_ld];
  for (int j = tid; j < H; j += blockDim.x) {
    x[j] = a.emb[(size_t)tok * H + j];
    h1[j] = a.h1[(size_t)b * H + j];
    h2[j] = a.h2[(size_t)b * H + j];
  }
  __syncthreads();
  for (int j = tid; j < H; j += blockDim.x) h1n[j] = gru_unit(x, h1, a.w1i, a.w1h, a.b1i, a.b1h, H, j);
  __syncthreads();
  for (int j = tid; j < H; j += blockDim.x) h2n[j] = gru_unit(h1n, h2, a.w2i, a.w2h, a.b2i, a.b2h, H, j);
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    float s = a.bo[c];
    for (int k = 0; k < H; ++k) s = fmaf(h2n[k], a.wo[(size_t)k * C + c], s);
    a.lm_out[(size_t)b * C + c] = s;
  }
  for (int j = tid; j < H; j += blockDim.x) {
    a.h1[(size_t)b * H + j] = h1n[j];
    a.h2[(size_t)b * H + j] = h2n[j];
  }
}
// final = log_softmax(asr) + weight * log_softmax(lm) (asr.py:153-156), argmax with torch's first-maximum rule; a warp per utterance
__global__ void __launch_bounds__(256) lm_mix_pick_kernel(int B, int C, const float* __restrict__ logits, long long logits_ld,
                                                          const float* __restrict__ lm, float weight, int* __restrict__ tok_out,
                                                          long long tok_ld) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* lg = logits + (size_t)b * logits_ld;
  const float* lr = lm + (size_t)b * C;
  float ml = -INFINITY, ma = -INFINITY;
  for (int c = lane; c < C; c += 32) { ml = fmaxf(ml, lr[c]); ma = fmaxf(ma, lg[c]); }
  ml = warp_max(ml);
  ma = warp_max(ma);
  float sl = 0.f, sa = 0.f;
  for (int c = lane; c < C; c += 32) { sl += expf(lr[c] - ml); sa += expf(lg[c] - ma); }
  sl = warp_sum(sl);
  sa = warp_sum(sa);
  const float lse_l = ml + logf(sl), lse_a = ma + logf(sa);
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float f = (lg[c] - lse_a) + weight * (lr[c] - lse_l);
    if (f > bv) { bv = f; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) tok_out[(size_t)b * tok_ld] = bi == 0x7fffffff ? 0 : bi;
}
// per-device scratch of the language-model logits of one step ([B, C] floats, grown on demand)
static float* lm_out_buffer(size_t n_floats) {
  static float* buf[32] = {nullptr};
  static size_t cap[32] = {0};
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return nullptr;
  if (cap[dev] < n_floats) {
    if (buf[dev]) { cudaDeviceSynchronize(); cudaFree(buf[dev]); buf[dev] = nullptr; cap[dev] = 0; }
    if (cudaMalloc(&buf[dev], n_floats * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    cap[dev] = n_floats;
  }
  return buf[dev];
}

// ------------------------------------------------------------------------------------------------
// fused cross-entropy (trainer.py:426-434): loss and dL/dlogits in one pass.  One CTA per utterance.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ce_kernel(int B, int U, int C, int L, const float* __restrict__ logits,
                                                 const long long* __restrict__ y, float* __restrict__ loss_b,
                                                 float* __restrict__ dlogits, float grad_scale) {
  __shared__ float scratch[32];
  __shared__ float part[4];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float cnt = 0.f;
  for (int i = tid; i < L; i += blockDim.x) cnt += (y[(size_t)b * L + i] != 0) ? 1.f : 0.f;
  cnt = block_sum(cnt, scratch);
  const float w = grad_scale / (cnt * (float)B);
  float acc = 0.f;
  for (int t = warp; t < U; t += 4) {
    const float* l = logits + ((size_t)b * U + t) * C;
    const int label = (int)y[(size_t)b * L + t + 1];
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, l[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(l[c] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    if (label != 0 && lane == 0) acc += lse - l[label];
    if (dlogits) {
      float* d = dlogits + ((size_t)b * U + t) * C;
      for (int c = lane; c < C; c += 32) {
        float g = 0.f;
        if (label != 0) g = (expf(l[c] - lse) - (c == label ? 1.f : 0.f)) * w;
        d[c] = g;
      }
    }
  }
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (tid == 0) loss_b[b] = (part[0] + part[1] + part[2] + part[3]) / cnt;
}
__global__ void mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += v[i];
    *out = s / (float)n;
  }
}

// Second stream + event pool of the decoder loops (bf16 training mode).  The layer-2 cell chain of the Speller never feeds
// the attention query (asr.py:84 takes state_list[0], the layer-1 state), so it only joins the layer-1 / attention chain
// where a sampled token is needed; run on its own stream it disappears from the dependent chain of a step.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[512];
  int n_ev = 0;
};
static SideStream* side_stream(int n_events, int which = 0) {
  static SideStream pool[2][32];
  static std::mutex mu;              // creation of the per-device streams / events from several host threads
  std::lock_guard<std::mutex> lk(mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32 || n_events > 512 || which < 0 || which > 1) return nullptr;
  SideStream* ss = &pool[which][dev];
  if (!ss->s) {
    // highest priority: its short kernels take the SMs that the long kernel of the main stream (attention over all utterances,
    // several waves of CTAs) frees, instead of queueing behind its undispatched CTAs
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&ss->s, cudaStreamNonBlocking, hi) != cudaSuccess) { ss->s = nullptr; cudaGetLastError(); return nullptr; }
  }
  while (ss->n_ev < n_events) {
    if (cudaEventCreateWithFlags(&ss->ev[ss->n_ev], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ++ss->n_ev;
  }
  return ss;
}
// 0 (SSASR_DECODE_DUAL=0): greedy decoding keeps every launch of a step on one stream
static int decode_dual_on() {
  const char* e = getenv("SSASR_DECODE_DUAL");
  return (e && e[0] == '0') ? 0 : 1;
}
// 0 (SSASR_DECODE_LM_SPLIT=0): CharLM recurrences + mix + argmax as one kernel on the layer-2 stream
static int lm_split_on() {
  const char* e = getenv("SSASR_DECODE_LM_SPLIT");
  return (e && e[0] == '0') ? 0 : 1;
}
static int x3_gemm_splits(int B) {
  const char* e = getenv("SSASR_X3_GEMM_SPLITS");
  const int v = e ? atoi(e) : (B <= 256 ? 2 : 1);
  return v < 1 ? 1 : v;
}
static int step_gemm_splits() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSASR_STEP_GEMM_SPLITS");
    v = e ? atoi(e) : 2;
    if (v < 1) v = 1;
  }
  return v;
}
// `waiter` continues only after everything enqueued on `src` so far
#define SSASR_HANDOVER(event, src, waiter)                   \
  do {                                                       \
    SSASR_CHECK_CUDA(cudaEventRecord((event), (src)));       \
    SSASR_CHECK_CUDA(cudaStreamWaitEvent((waiter), (event), 0)); \
  } while (0)

}  // namespace ssasr

using namespace ssasr;

extern "C" {

// Plain-C argument block for the decoder loop (all device pointers unless noted).
typedef struct {
  int B, Tp, E, Sd, M, C, U;
  // parameters (kernel layout, see ssasr_pack_lstmcell)
  const float *phi_w, *psi_w, *psi_b, *w1cat, *b1, *w2cat, *b2, *emb_w, *wc, *bc;
  // inputs
  const float* enc;          // [B,Tp,E]
  const int* enc_lens;       // [B]
  int* tok_in;               // [B,U] input token of every step (col 0 = SOS; teacher columns pre-filled)
  const int* step_mode;      // HOST [U]: how the token AFTER step t is chosen: 0 teacher, 1 argmax, 2 sample
  unsigned long long seed;
  // outputs / saved state
  float *psi, *xin1, *xin2, *act1, *act2, *c1, *c2, *h2all, *q, *alpha, *logits;
  // bf16 tensor-core mode (all three non-null): bf16 copies of w1cat / w2cat and a [B, max(X1,X2)] bf16 scratch
  const void *w1cat_bf, *w2cat_bf;
  void* ws_bf;
  void* enc_bf;   // bf16 scratch [(B*Tp + M) * E], or NULL
  // character LM for step_mode 3 (ASR.decode with lm_weight != 0): transposed GRU / output weights, state [B,H]
  int lm_H;
  float lm_weight;
  const float *lm_emb, *lm_w1i, *lm_w1h, *lm_b1i, *lm_b1h, *lm_w2i, *lm_w2h, *lm_b2i, *lm_b2h, *lm_wo, *lm_bo;
  float *lm_h1, *lm_h2;
  // forward-only fast exact mode: gate GEMMs on tensor cores with the tf32 x 3 split;
  // scratch of 2*B*max(X1,X2) + 2*4Sd*(X1+X2) floats, or NULL
  float* x3_ws;
  int skip_final_logits;
  // bf16 mode only: ws_bf holds B*X1 + U*B*X2 elements (one layer-2 input block per step) and the layer-2 chain may run on
  // an internal second stream, joined into `stream` before the call returns
  int dual_stream;
  // greedy decoding (every step selects its successor's token): stop as soon as EVERY utterance has emitted `stop_token`
  // (asr.py:161-162 stops each bs=1 call at EOS; a batch is done when all of its utterances are).  Checked every
  // `stop_check_every` steps with a 4-byte read-back and a stream synchronisation; 0 = run all U steps.
  int stop_token, stop_check_every;
  int* stop_scratch;      // device int
  int* steps_run;         // HOST int out (may be NULL): decoding steps actually executed
  // bf16 mode, optional: workspace of ssasr_speller_cl_ws_bytes() bytes.  When given (and the dimensions are covered) the
  // teacher-forced runs of the loop execute in the cluster-persistent step kernel of spell_cl.cu: one launch per run of steps
  // instead of two launches per step.  The caller keeps it alive until the backward pass has run (it holds P and psi~ in bf16).
  void* cl_ws;
  long long cl_ws_bytes;
} ssasr_speller_fwd_args;

// workspace layout of the cluster path
struct SpellClWs {
  size_t p_bf, psi_bf, phi_bf, gemb, p_f32, xp2, total;
};
static SpellClWs spell_cl_ws_layout(int B, int Tp, int Sd, int M, int C, int U) {
  SpellClWs w;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t o = 0;
  w.p_bf = o; o += up((size_t)B * Tp * 4 * Sd * 2);
  w.psi_bf = o; o += up((size_t)B * Tp * M * 2);
  w.phi_bf = o; o += up((size_t)M * Sd * 2);
  w.gemb = o; o += up((size_t)C * 4 * Sd * 4);
  w.p_f32 = o; o += up((size_t)B * Tp * 4 * Sd * 4);
  w.xp2 = o; o += up((size_t)B * U * 4 * Sd * 4);      // layer-2 input projection, [U][B][4Sd]
  w.total = o;
  return w;
}
// bytes of `cl_ws` for these dimensions; 0 when the cluster path does not cover them (the per-step kernels run instead)
long long ssasr_speller_cl_ws_bytes(int B, int Tp, int E, int Sd, int M, int C, int U) {
  if (!spell_cl_supported(B, Tp, E, Sd, M)) return 0;
  return (long long)spell_cl_ws_layout(B, Tp, Sd, M, C, U).total;
}

int ssasr_speller_fwd_f32(const ssasr_speller_fwd_args* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int B = a->B, Tp = a->Tp, E = a->E, Sd = a->Sd, M = a->M, C = a->C, U = a->U;
  const int K1 = Sd + E, X1 = K1 + Sd, X2 = 2 * Sd;
  SSASR_REQUIRE(Sd % 4 == 0, "speller: decoder state size %d must be a multiple of 4", Sd);
  int rc;
  if (a->w1cat_bf && a->w2cat_bf && a->ws_bf && a->enc_bf && E % 8 == 0) {
    // psi~ = tanh(enc @ Wpsi^T + b) on tensor cores: enc_bf [B*Tp,E] and (after it) Wpsi bf16 [M,E] live in enc_bf
    __nv_bfloat16* eb = (__nv_bfloat16*)a->enc_bf;
    __nv_bfloat16* wb = eb + (size_t)B * Tp * E;
    rc = cvt_bf16(st, a->enc, E, eb, E, (long long)B * Tp, E);
    if (rc) return rc;
    rc = cvt_bf16(st, a->psi_w, E, wb, E, M, E);
    if (rc) return rc;
    rc = gemm_bf16_tc(st, B * Tp, M, E, eb, E, 0, wb, E, 0, a->psi, M, a->psi_b, 0, 1);
  } else {
    rc = gemm_f32(st, B * Tp, M, E, a->enc, E, 1, a->psi_w, E, 1, a->psi, M, a->psi_b, 0, 1);
  }
  if (rc) return rc;
  const size_t attn_smem = attn_fwd_smem_bytes(Sd, M, Tp, E);
  SSASR_REQUIRE(attn_smem <= 200 * 1024, "speller: attention working set too large (Tp=%d)", Tp);
  if (attn_smem > 48 * 1024)
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem));
  const int cell_blocks = (B * Sd + 255) / 256;
  const bool tc = a->w1cat_bf && a->w2cat_bf && a->ws_bf && X1 % 8 == 0 && X2 % 8 == 0;
  // gates = x_t @ Wcat^T + b, fp32 SIMT or (bf16 mode) tcgen05 after a bf16 copy of the step's input rows
  // bf16 mode: the attention / cell kernels also emit bf16 copies of the step's input rows into ws_bf
  // ([B, X1] followed by [B, X2]), which feed the tcgen05 gate GEMMs directly
  __nv_bfloat16* x1b = tc ? (__nv_bfloat16*)a->ws_bf : nullptr;
  __nv_bfloat16* x2b = tc ? x1b + (size_t)B * X1 : nullptr;
  const bool x3 = !tc && a->x3_ws && X1 % 4 == 0 && X2 % 4 == 0;
  float *xh = nullptr, *xl = nullptr, *w1h = nullptr, *w1l = nullptr, *w2h = nullptr, *w2l = nullptr;
  float* xh2 = nullptr;                              // x3b: operand rows of the layer-2 product (else they share xh / xl)
  // x3b: the split-operand gate products as bf16 GEMMs over the tripled reduction axis (the producers then write bf16
  // [hi | hi | lo] rows into the same workspace and `xl` stays null); else the tf32 x 3 kernel on fp32 hi / lo pairs
  const bool x3b = x3 && X1 % 8 == 0 && X2 % 8 == 0 && x3_gemm_bf16();
  if (x3) {
    const int Xm = X1 > X2 ? X1 : X2;
    xh = a->x3_ws; xl = xh + (size_t)B * Xm;
    w1h = xh + (size_t)2 * B * (X1 + X2); w1l = w1h + (size_t)4 * Sd * X1;
    w2h = w1l + (size_t)4 * Sd * X1; w2l = w2h + (size_t)4 * Sd * X2;
    if (x3b) {
      xl = nullptr;
      xh2 = xh + ((size_t)3 * B * X1 + 1) / 2;        // the layer-2 input rows have their own block: [B, 3 X1] bf16, then [B, 3 X2]
      rc = split3_bf16(st, a->w1cat, X1, 4 * Sd, X1, w1h, 1);
      if (rc) return rc;
      rc = split3_bf16(st, a->w2cat, X2, 4 * Sd, X2, w2h, 1);
      if (rc) return rc;
    } else {
      rc = split_hi_lo(st, a->w1cat, w1h, w1l, (size_t)4 * Sd * X1);
      if (rc) return rc;
      rc = split_hi_lo(st, a->w2cat, w2h, w2l, (size_t)4 * Sd * X2);
      if (rc) return rc;
    }
  }
  // bf16 mode: layer-2 chain on a second stream (one layer-2 input block per step in ws_bf)
  // cluster-persistent step kernel for the teacher-forced runs (spell_cl.cu); needs one layer-2 input block per step in ws_bf
  const bool cl = tc && a->dual_stream && a->cl_ws && a->enc_bf && E % 8 == 0 && spell_cl_supported(B, Tp, E, Sd, M) &&
                  (size_t)Tp * ((U + 15) / 16 * 16) * sizeof(float) <= 96 * 1024 &&      // spell_fill_xin1's attention-map tile
                  a->cl_ws_bytes >= (long long)spell_cl_ws_layout(B, Tp, Sd, M, C, U).total;
  SideStream* side = (!cl && tc && a->dual_stream && U > 1) ? side_stream(2 * U + 2) : nullptr;
  const bool dual = side != nullptr;
  cudaStream_t sb = dual ? side->s : st;
  auto x2b_at = [&](int t) { return ((dual || cl) && x2b) ? x2b + (size_t)t * B * X2 : x2b; };
  // the layer-1 product sits on the dependent chain and is bound by what one SM can pull through its L2 port (128 x 32 tiles:
  // 320 KB per CTA, 64 CTAs): split-K over twice the SMs, partials added into the pre-zeroed gate buffer
  const int chain_splits = (tc && !cl) ? step_gemm_splits() : 1;
  if (chain_splits > 1) SSASR_CHECK_CUDA(cudaMemsetAsync(a->act1, 0, sizeof(float) * (size_t)B * U * 4 * Sd, st));
  // exact per-step gate products (tripled-K bf16) of SMALL batches (a shard of an utterance-sharded decode: <= 256 utterances are
  // <= 64 CTAs of 128 x 32 tiles): split-K over twice the CTAs, the two partials added into the pre-zeroed gate buffers (two
  // addends: order-independent).  Measured: 125 utterances 15.2 -> 14.2 ms, 1000 utterances 35.0 -> 36.7 ms (hence the threshold)
  const int x3_splits = x3b ? x3_gemm_splits(B) : 1;
  if (x3_splits > 1) {
    SSASR_CHECK_CUDA(cudaMemsetAsync(a->act1, 0, sizeof(float) * (size_t)B * U * 4 * Sd, st));
    SSASR_CHECK_CUDA(cudaMemsetAsync(a->act2, 0, sizeof(float) * (size_t)B * U * 4 * Sd, st));
  }
  auto gate_gemm = [&](cudaStream_t st, const float* x, int ldx, int K, const float* w, const void* w_bf, const float* bias, float* out,
                       const __nv_bfloat16* xb, int splits = 1) -> int {
    if (tc) return gemm_bf16_tc(st, B, 4 * Sd, K, xb, K, 0, w_bf, K, 0, out, U * 4 * Sd, bias, 0, 0, splits);
    if (x3) {                  // xh / xl were written by the kernel that produced x (attention step / layer-1 cell)
      const bool first = (K == X1);
      if (x3b) return gemm_bf16_tc(st, B, 4 * Sd, 3 * K, first ? xh : xh2, 3 * K, 0, first ? w1h : w2h, 3 * K, 0, out, U * 4 * Sd, bias, 0, 0, x3_splits);
      return gemm_tf32x3(st, B, 4 * Sd, K, xh, xl, K, first ? w1h : w2h, first ? w1l : w2l, K, out, U * 4 * Sd, bias, 0);
    }
    return gemm_f32(st, B, 4 * Sd, K, x, ldx, 1, w, K, 1, out, U * 4 * Sd, bias, 0, 0);
  };
  const size_t lp_smem = lp_smem_bytes(C, Sd);
  const bool lp_fused = C <= 64 && lp_smem <= 160 * 1024;
  if (lp_fused && lp_smem > 48 * 1024) {
    static size_t lp_attr = 0;
    if (lp_smem > lp_attr) {
      SSASR_CHECK_CUDA(cudaFuncSetAttribute(logits_pick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp_smem));
      lp_attr = lp_smem;
    }
  }
  auto mode_of = [&](int t) { return a->step_mode ? a->step_mode[t] : 0; };
  // attention step t (+ the step's input row); fuse_prev: the layer-1 cell of step t-1 runs in its prologue
  auto attn_step = [&](int t, bool fuse_prev, bool skip_emb = false) {
    AttnFwd f{};
    f.skip_emb = skip_emb ? 1 : 0;
    f.Tp = Tp; f.E = E; f.Sd = Sd; f.M = M;
    f.h1prev = t ? a->xin2 + (size_t)(t - 1) * X2 : nullptr; f.h1_ld = (long long)U * X2;
    f.phi_w = a->phi_w; f.psi = a->psi; f.enc = a->enc; f.enc_lens = a->enc_lens;
    f.emb_w = a->emb_w; f.tok = a->tok_in + t; f.tok_ld = U;
    f.xin1 = a->xin1 + (size_t)t * X1; f.xin1_ld = (long long)U * X1;
    f.xin1b = x1b; f.xin1b_ld = X1;
    f.x3h = x3 ? xh : nullptr; f.x3l = x3 ? xl : nullptr; f.x3_ld = X1;
    f.q = a->q + (size_t)t * M; f.q_ld = (long long)U * M;
    f.alpha = a->alpha + (size_t)t * Tp; f.alpha_ld = (long long)U * Tp;
    if (fuse_prev) {
      f.cell_gates = a->act1 + (size_t)(t - 1) * 4 * Sd; f.cg_ld = (long long)U * 4 * Sd;
      f.cell_cprev = t > 1 ? a->c1 + (size_t)(t - 2) * Sd : nullptr; f.ccp_ld = (long long)U * Sd;
      f.cell_cout = a->c1 + (size_t)(t - 1) * Sd; f.cco_ld = (long long)U * Sd;
      f.cell_hout = a->xin2 + (size_t)(t - 1) * X2; f.cho_ld = (long long)U * X2;
      f.cell_hb = x2b_at(t - 1); f.chb_ld = X2;
    }
    ProfScope ps(F_ATTN_FWD, st);
    attn_fwd_kernel<<<B, 256, attn_smem, st>>>(f);
  };
  // layer-1 cell of step t as its own launch: h1(t) into the layer-2 input row; copy_h2: h2(t-1) (or zeros) next to it
  auto cell1 = [&](int t, bool copy_h2) {
    ProfScope ps(F_POINTWISE, st);
    cell_fwd_kernel<<<cell_blocks, 256, 0, st>>>(B, Sd, a->act1 + (size_t)t * 4 * Sd, (long long)U * 4 * Sd,
                                                 t ? a->c1 + (size_t)(t - 1) * Sd : nullptr, (long long)U * Sd,
                                                 a->c1 + (size_t)t * Sd, (long long)U * Sd, a->xin2 + (size_t)t * X2,
                                                 (long long)U * X2, t ? a->h2all + (size_t)(t - 1) * Sd : nullptr,
                                                 (long long)U * Sd, copy_h2 ? a->xin2 + (size_t)t * X2 + Sd : nullptr,
                                                 (long long)U * X2, x2b_at(t), X2, Sd, x3 ? (x3b ? xh2 : xh) : nullptr, x3 ? xl : nullptr, X2);
  };
  // layer 2 of step t on stream s2 (+ token selection for step t+1, which consumes h2(t))
  auto layer2_cell = [&](int t, cudaStream_t s2) -> int {
    int r = gate_gemm(s2, a->xin2 + (size_t)t * X2, U * X2, X2, a->w2cat, a->w2cat_bf, a->b2, a->act2 + (size_t)t * 4 * Sd, x2b_at(t));
    if (r) return r;
    const bool fwd_h2 = (dual || cl) && t + 1 < U;   // per-step input blocks: h2(t) goes straight into the next step's layer-2 input row
    {
      ProfScope ps(F_POINTWISE, s2);
      cell_fwd_kernel<<<cell_blocks, 256, 0, s2>>>(B, Sd, a->act2 + (size_t)t * 4 * Sd, (long long)U * 4 * Sd,
                                                   t ? a->c2 + (size_t)(t - 1) * Sd : nullptr, (long long)U * Sd,
                                                   a->c2 + (size_t)t * Sd, (long long)U * Sd, a->h2all + (size_t)t * Sd,
                                                   (long long)U * Sd, nullptr, 0, nullptr, 0, nullptr, 0, 0, nullptr, nullptr, 0,
                                                   fwd_h2 ? a->xin2 + (size_t)(t + 1) * X2 + Sd : nullptr, (long long)U * X2,
                                                   fwd_h2 ? x2b_at(t + 1) + Sd : nullptr, X2);
    }
    return 0;
  };
  const float* lm_ext = nullptr;     // two-stream greedy loop with the CharLM: the step's LM logits computed on the third stream
  // selection of the token that follows step t (consumes h2(t)) on stream s2; no-op for teacher-forced steps
  auto select_token = [&](int t, cudaStream_t s2) -> int {
    int r = 0;
    const int mode = mode_of(t);
    if (mode == 0 || t + 1 >= U) return 0;
    if (lp_fused) {          // projection + selection in one launch (mode 3: projection only, the LM kernel selects)
      ProfScope ps(F_POINTWISE, s2);
      logits_pick_kernel<<<(B + 8 * LP_UPW - 1) / (8 * LP_UPW), 256, lp_smem, s2>>>(
          B, C, Sd, a->h2all + (size_t)t * Sd, (long long)U * Sd, a->wc, a->bc, a->logits + (size_t)t * C, (long long)U * C,
          mode == 3 ? 0 : mode, a->seed, (unsigned long long)t, a->tok_in + t + 1, U);
    } else {
      r = gemm_f32(s2, B, C, Sd, a->h2all + (size_t)t * Sd, U * Sd, 1, a->wc, Sd, 1, a->logits + (size_t)t * C, U * C, a->bc, 0, 0);
      if (r) return r;
    }
    if (mode == 3 && lm_ext) {                // the recurrences of this step ran on the third stream: mix + argmax only
      ProfScope ps(F_POINTWISE, s2);
      lm_mix_pick_kernel<<<(B + 7) / 8, 256, 0, s2>>>(B, C, a->logits + (size_t)t * C, (long long)U * C, lm_ext, a->lm_weight,
                                                      a->tok_in + t + 1, U);
    } else if (mode == 3) {
      SSASR_REQUIRE(a->lm_emb && a->lm_h1 && a->lm_h2 && a->lm_H > 0, "speller: step mode 3 needs the language-model arguments");
      LmStep l;
      l.C = C; l.H = a->lm_H;
      l.emb = a->lm_emb; l.w1i = a->lm_w1i; l.w1h = a->lm_w1h; l.b1i = a->lm_b1i; l.b1h = a->lm_b1h;
      l.w2i = a->lm_w2i; l.w2h = a->lm_w2h; l.b2i = a->lm_b2i; l.b2h = a->lm_b2h; l.wo = a->lm_wo; l.bo = a->lm_bo;
      l.h1 = a->lm_h1; l.h2 = a->lm_h2; l.weight = a->lm_weight;
      l.logits = a->logits + (size_t)t * C; l.logits_ld = (long long)U * C;
      l.tok_in = a->tok_in + t; l.tok_out = a->tok_in + t + 1; l.tok_ld = U;
      const int nt = ((a->lm_H + 31) / 32) * 32 > 256 ? 256 : ((a->lm_H + 31) / 32) * 32;
      ProfScope ps(F_POINTWISE, s2);
      lm_pick_kernel<<<B, nt, (size_t)(5 * a->lm_H + C + 32) * sizeof(float), s2>>>(l);
    } else if (!lp_fused) {
      ProfScope ps(F_POINTWISE, s2);
      pick_token_kernel<<<(B + 127) / 128, 128, 0, s2>>>(B, C, a->logits + (size_t)t * C, (long long)U * C, mode, a->seed,
                                                         (unsigned long long)t, a->tok_in + t + 1, U);
    }
    if (dual) SSASR_CHECK_CUDA(cudaEventRecord(side->ev[U + t], s2));   // the next step's token is ready
    return 0;
  };
  auto layer2 = [&](int t, cudaStream_t s2) -> int {
    int r = layer2_cell(t, s2);
    if (r) return r;
    return select_token(t, s2);
  };
  int steps_done = U;
  if (cl) {
    // ---- cluster-persistent path: P = enc W_ctx^T, psi~ and phi in bf16, the embedding table, then one launch per run ----
    const SpellClWs wl = spell_cl_ws_layout(B, Tp, Sd, M, C, U);
    uint8_t* ws = (uint8_t*)a->cl_ws;
    __nv_bfloat16* p_bf = (__nv_bfloat16*)(ws + wl.p_bf);
    __nv_bfloat16* psi_bf = (__nv_bfloat16*)(ws + wl.psi_bf);
    __nv_bfloat16* phi_bf = (__nv_bfloat16*)(ws + wl.phi_bf);
    float* gemb = (float*)(ws + wl.gemb);
    float* p_f32 = (float*)(ws + wl.p_f32);
    rc = cvt_bf16(st, a->psi, M, psi_bf, M, (long long)B * Tp, M);
    if (rc) return rc;
    rc = cvt_bf16(st, a->phi_w, Sd, phi_bf, Sd, M, Sd);
    if (rc) return rc;
    // G_emb[c, :] = W_emb emb[c] + b1 (fp32): the embedding part of the layer-1 gates is a table row per token
    rc = gemm_f32(st, C, 4 * Sd, Sd, a->emb_w, Sd, 1, a->w1cat, X1, 1, gemb, 4 * Sd, a->b1, 0, 0);
    if (rc) return rc;
    // P[b, j, :] = W_ctx enc[b, j, :] on tensor cores (enc_bf was filled for the psi~ product above)
    rc = gemm_bf16_tc(st, B * Tp, 4 * Sd, E, a->enc_bf, E, 0, a->w1cat_bf, X1, Sd, p_f32, 4 * Sd, nullptr, 0);
    if (rc) return rc;
    rc = cvt_bf16(st, p_f32, 4 * Sd, p_bf, 4 * Sd, (long long)B * Tp, 4 * Sd);
    if (rc) return rc;
    // h2(-1) = 0: the recurrent half of the first layer-2 input block
    SSASR_CHECK_CUDA(cudaMemset2DAsync(a->xin2 + Sd, sizeof(float) * (size_t)U * X2, 0, sizeof(float) * Sd, B, st));
    SpellClFwdArgs f = {};
    f.B = B; f.U = U; f.Tp = Tp;
    f.w1cat_bf = a->w1cat_bf; f.X1 = X1; f.K1 = K1; f.phi_bf = phi_bf; f.P_bf = p_bf; f.psi_bf = psi_bf; f.gemb = gemb;
    f.tok = a->tok_in; f.tok_ld = U; f.enc_lens = a->enc_lens;
    f.act1 = a->act1; f.act1_ldb = (long long)U * 4 * Sd; f.act1_ldt = 4 * Sd;
    f.c1 = a->c1; f.c1_ldb = (long long)U * Sd; f.c1_ldt = Sd;
    f.h1 = a->xin2; f.h1_ldb = (long long)U * X2; f.h1_ldt = X2;
    f.h1b = x2b; f.h1b_ldb = X2; f.h1b_ldt = (long long)B * X2;
    f.q = a->q; f.q_ldb = (long long)U * M; f.q_ldt = M;
    f.alpha = a->alpha; f.al_ldb = (long long)U * Tp; f.al_ldt = Tp;
    float* xp2 = (float*)(ws + wl.xp2);
    SpellClFwdArgs g = {};                   // the layer-2 chain: plain recurrence mode of the same kernel
    g.B = B; g.U = U; g.Tp = Tp;
    g.w1cat_bf = a->w2cat_bf; g.X1 = X2; g.K1 = Sd;
    g.xpre = xp2; g.xpre_ldb = 4 * Sd; g.xpre_ldt = (long long)B * 4 * Sd;
    g.act1 = a->act2; g.act1_ldb = (long long)U * 4 * Sd; g.act1_ldt = 4 * Sd;
    g.c1 = a->c2; g.c1_ldb = (long long)U * Sd; g.c1_ldt = Sd;
    g.h1 = a->h2all; g.h1_ldb = (long long)U * Sd; g.h1_ldt = Sd;
    g.h2nd = a->xin2 + Sd; g.h2nd_ldb = (long long)U * X2; g.h2nd_ldt = X2; g.h2nd_toff = 1;
    for (int t0 = 0; t0 < U;) {
      int t1 = t0 + 1;                       // a run ends behind the first step whose successor token is selected on the device
      while (t1 < U && mode_of(t1 - 1) == 0) ++t1;
      f.t0 = t0; f.t1 = t1;
      rc = spell_cl_fwd(st, f);
      if (rc) return rc;
      // layer 2 of the run: input projection of all its steps in one product (rows [t0 B, t1 B) of the time-major blocks), then
      // the cell chain in the cluster recurrence, then the token that follows the run if it is selected on the device
      rc = gemm_bf16_tc(st, (t1 - t0) * B, 4 * Sd, Sd, x2b + (size_t)t0 * B * X2, X2, 0, a->w2cat_bf, X2, 0,
                        xp2 + (size_t)t0 * B * 4 * Sd, 4 * Sd, a->b2, 0);
      if (rc) return rc;
      g.t0 = t0; g.t1 = t1;
      rc = spell_cl_fwd(st, g);
      if (rc) return rc;
      rc = select_token(t1 - 1, st);
      if (rc) return rc;
      t0 = t1;
    }
    // the operand rows of the weight-gradient products (the loop never materialises the context)
    rc = spell_fill_xin1(st, B, U, Tp, E, Sd, a->alpha, a->enc, a->enc_lens, a->emb_w, a->tok_in, a->xin2, (long long)U * X2, X2, a->xin1);
    if (rc) return rc;
  } else
  {
  const int stop_every = (!dual && a->stop_check_every > 0 && a->stop_scratch && a->skip_final_logits && a->step_mode &&
                          a->step_mode[0] != 0 && a->step_mode[0] != 2) ? a->stop_check_every : 0;
  // greedy / LM decoding (every token selected on the device): the attention query of step t+1 is the LAYER-1 state of step t
  // (asr.py:84), so the layer-2 chain of step t -- gate product, cell, character projection + selection, embedding row of the
  // next step's input -- runs on a second stream UNDER the attention step t+1 (the longest kernel of a step) and joins before
  // its gate product.  Needs the operand rows of the two gate products in separate blocks (x3b, or the fp32 SIMT path).
  SideStream* gside = nullptr;
  if (!dual && U > 1 && 2 * U + 2 <= 512 && (!x3 || x3b) && a->step_mode && decode_dual_on()) {
    bool all_sel = true;
    for (int t = 0; t + 1 < U; ++t) all_sel = all_sel && (a->step_mode[t] == 1 || a->step_mode[t] == 3);
    if (all_sel) gside = side_stream(2 * U + 2);
  }
  if (gside) {
    cudaStream_t s2 = gside->s;
    int last_b = -1;
    // CharLM rescoring: the LM recurrences of step t need token t only -> third stream, under the attention step / gate products
    SideStream* lside = nullptr;
    float* lm_buf = nullptr;
    if (a->step_mode[0] == 3 && a->lm_emb && a->lm_h1 && a->lm_h2 && a->lm_H > 0 && U + 2 <= 512 && lm_split_on()) {
      lm_buf = lm_out_buffer((size_t)B * C);
      if (lm_buf) lside = side_stream(U + 2, 1);
    }
    if (lside) SSASR_HANDOVER(lside->ev[U], st, lside->s);       // everything enqueued so far (weights, LM state) is visible
    for (int t = 0; t < U; ++t) {
      if (lside && t + 1 < U && mode_of(t) == 3) {
        cudaStream_t s3 = lside->s;
        if (t > 0) SSASR_CHECK_CUDA(cudaStreamWaitEvent(s3, gside->ev[U + t - 1], 0));    // token t (and the mix that read lm_buf)
        LmStep l;
        l.C = C; l.H = a->lm_H;
        l.emb = a->lm_emb; l.w1i = a->lm_w1i; l.w1h = a->lm_w1h; l.b1i = a->lm_b1i; l.b1h = a->lm_b1h;
        l.w2i = a->lm_w2i; l.w2h = a->lm_w2h; l.b2i = a->lm_b2i; l.b2h = a->lm_b2h; l.wo = a->lm_wo; l.bo = a->lm_bo;
        l.h1 = a->lm_h1; l.h2 = a->lm_h2; l.weight = a->lm_weight;
        l.logits = nullptr; l.logits_ld = 0;
        l.tok_in = a->tok_in + t; l.tok_out = nullptr; l.tok_ld = U;
        l.lm_out = lm_buf;
        const int nt = ((a->lm_H + 31) / 32) * 32 > 256 ? 256 : ((a->lm_H + 31) / 32) * 32;
        {
          ProfScope ps(F_POINTWISE, s3);
          lm_step_kernel<<<B, nt, (size_t)(5 * a->lm_H) * sizeof(float), s3>>>(l);
        }
        SSASR_CHECK_CUDA(cudaEventRecord(lside->ev[t], s3));
      }
      attn_step(t, false, t > 0);
      if (t > 0) SSASR_CHECK_CUDA(cudaStreamWaitEvent(st, gside->ev[U + t - 1], 0));   // token t, its embedding row, h2(t-1)
      rc = gate_gemm(st, a->xin1 + (size_t)t * X1, U * X1, X1, a->w1cat, a->w1cat_bf, a->b1, a->act1 + (size_t)t * 4 * Sd, x1b, 1);
      if (rc) return rc;
      cell1(t, true);
      SSASR_HANDOVER(gside->ev[t], st, s2);
      lm_ext = nullptr;
      if (lside && t + 1 < U && mode_of(t) == 3) {
        SSASR_CHECK_CUDA(cudaStreamWaitEvent(s2, lside->ev[t], 0));
        lm_ext = lm_buf;
      }
      rc = layer2(t, s2);
      lm_ext = nullptr;
      if (rc) return rc;
      if (t + 1 < U) {
        ProfScope ps(F_POINTWISE, s2);
        emb_rows_kernel<<<cell_blocks, 256, 0, s2>>>(B, Sd, a->emb_w, a->tok_in + t + 1, U, a->xin1 + (size_t)(t + 1) * X1,
                                                     (long long)U * X1, x3 ? xh : nullptr, x3 ? xl : nullptr, X1);
      }
      SSASR_CHECK_CUDA(cudaEventRecord(gside->ev[U + t], s2));
      last_b = t;
      steps_done = t + 1;
      if (stop_every > 0 && t + 1 < U && (t + 1) % stop_every == 0) {
        SSASR_CHECK_CUDA(cudaStreamWaitEvent(st, gside->ev[U + t], 0));
        SSASR_CHECK_CUDA(cudaMemsetAsync(a->stop_scratch, 0, sizeof(int), st));
        count_unfinished_kernel<<<(B + 7) / 8, 256, 0, st>>>(B, a->tok_in, U, t + 1, a->stop_token, a->stop_scratch);
        int open_utts = 1;
        SSASR_CHECK_CUDA(cudaMemcpyAsync(&open_utts, a->stop_scratch, sizeof(int), cudaMemcpyDeviceToHost, st));
        SSASR_CHECK_CUDA(cudaStreamSynchronize(st));
        if (open_utts == 0) break;
      }
    }
    if (last_b >= 0) SSASR_CHECK_CUDA(cudaStreamWaitEvent(st, gside->ev[U + last_b], 0));
  } else if (!dual) {
    for (int t = 0; t < U; ++t) {
      attn_step(t, false);
      rc = gate_gemm(st, a->xin1 + (size_t)t * X1, U * X1, X1, a->w1cat, a->w1cat_bf, a->b1, a->act1 + (size_t)t * 4 * Sd, x1b,
                     chain_splits);
      if (rc) return rc;
      cell1(t, true);
      rc = layer2(t, st);
      if (rc) return rc;
      steps_done = t + 1;
      if (stop_every > 0 && t + 1 < U && (t + 1) % stop_every == 0) {
        // tokens 1 .. t+1 have been selected: has every utterance emitted the stop token?
        SSASR_CHECK_CUDA(cudaMemsetAsync(a->stop_scratch, 0, sizeof(int), st));
        count_unfinished_kernel<<<(B + 7) / 8, 256, 0, st>>>(B, a->tok_in, U, t + 1, a->stop_token, a->stop_scratch);
        int open_utts = 1;
        SSASR_CHECK_CUDA(cudaMemcpyAsync(&open_utts, a->stop_scratch, sizeof(int), cudaMemcpyDeviceToHost, st));
        SSASR_CHECK_CUDA(cudaStreamSynchronize(st));
        if (open_utts == 0) break;
      }
    }
  } else {
    // Two streams.  `st`: attention -> layer-1 gate GEMM (-> layer-1 cell); `sb`: layer-2 GEMM -> cell (-> token selection).
    // The layer-1 cell of a teacher-forced step is deferred into the prologue of the next attention step (one launch less on
    // the dependent chain); where the next token is sampled from this step's output it must run now, because the selection
    // (on `sb`) needs h1(t) -> h2(t) before the next attention step can start.
    auto deferred = [&](int t) { return t >= 1 && t + 1 < U && mode_of(t) == 0; };
    for (int t = 0; t < U; ++t) {
      const bool prev_def = t > 0 && deferred(t - 1);
      // the token of this step was selected from the previous step's layer-2 output
      if (t > 0 && mode_of(t - 1) != 0) SSASR_CHECK_CUDA(cudaStreamWaitEvent(st, side->ev[U + t - 1], 0));
      attn_step(t, prev_def);
      if (prev_def) {
        SSASR_HANDOVER(side->ev[t - 1], st, sb);
        rc = layer2(t - 1, sb);
        if (rc) return rc;
      }
      rc = gate_gemm(st, a->xin1 + (size_t)t * X1, U * X1, X1, a->w1cat, a->w1cat_bf, a->b1, a->act1 + (size_t)t * 4 * Sd, x1b,
                     chain_splits);
      if (rc) return rc;
      if (!deferred(t)) {
        cell1(t, t == 0);
        SSASR_HANDOVER(side->ev[t], st, sb);
        rc = layer2(t, sb);
        if (rc) return rc;
      }
    }
  }
  if (dual) SSASR_HANDOVER(side->ev[2 * U], sb, st);
  }
  if (a->steps_run) *a->steps_run = steps_done;
  if (a->skip_final_logits) {        // greedy decoding only consumes the tokens
    SSASR_LAUNCH_CHECK();
    return 0;
  }
  rc = gemm_f32(st, B * U, C, Sd, a->h2all, Sd, 1, a->wc, Sd, 1, a->logits, C, a->bc, 0, 0);
  if (rc) return rc;
  SSASR_LAUNCH_CHECK();
  return 0;
}

typedef struct {
  int B, Tp, E, Sd, M, C, U;
  const float *phi_w, *psi_w, *w1cat, *w2cat, *wc;
  const float* enc;
  const int* enc_lens;
  const int* tok_in;
  // saved by forward (act1/act2 are overwritten with gate gradients)
  const float *psi, *xin1, *xin2, *c1, *c2, *h2all, *q, *alpha;
  float *act1, *act2;
  const float* dlogits;      // [B,U,C]
  // gradient outputs (overwritten), kernel layout
  float *d_phi_w, *d_psi_w, *d_psi_b, *d_w1cat, *d_b1, *d_w2cat, *d_b2, *d_emb_w, *d_wc, *d_bc, *denc;
  // scratch
  float *dh2all /*[B,U,Sd]*/, *dxin1 /*[B,U,X1]*/, *dxin2 /*[B,X2]*/, *dc1s, *dc2s, *dh1att /*[B,Sd] each*/,
      *dpsi /*[B,Tp,M]*/, *dqpre /*[B,U,M]*/, *de_all /*[B,U,Tp]*/;
  // bf16 tensor-core mode (all non-null): transposed bf16 weights [X1,4Sd] / [X2,4Sd], scratch wsA [4Sd,BUp],
  // wsB [X1,BUp] (BUp = B*U rounded up to 8)
  const void *w1catT_bf, *w2catT_bf;
  void *wsA, *wsB;   // element counts: wsA >= max(4Sd*BUp, M*BTp), wsB >= max(X1*BUp, E*BTp + E*M)
  long long BUp, BTp;
  // bf16 mode only: dxin2 holds U*B*X2 floats (one block per step) and the layer-2 chain may run on an internal second
  // stream, joined into `stream` before the call returns
  int dual_stream;
  // optional (two-stream bf16 mode only): every product that only the optimiser consumes (all weight / bias / embedding
  // gradients) is enqueued on this stream, ordered after the loop, and NOT joined into `stream`: the caller joins it before it
  // reads the gradients.  `stream` then carries just what the encoder's backward needs (denc).
  void* wgrad_stream;
  // cluster-persistent path (csrc/spell_cl.cu), all optional: the forward call's `cl_ws` (P, psi~, phi in bf16), the forward
  // bf16 weights [4Sd, X1] / [4Sd, X2] and a scratch of ssasr_speller_cl_bwd_ws_bytes() bytes.  When given (and the forward ran
  // on that path) both cell chains and the attention backward run as ONE launch each over all steps.
  const void* cl_ws;
  const void *w1cat_bf, *w2cat_bf;
  void* cl_ws_bwd;
  long long cl_ws_bwd_bytes;
} ssasr_speller_bwd_args;

struct SpellClBwdWs {
  size_t dgb1, dgb2, dh1, total;
};
static SpellClBwdWs spell_cl_bwd_ws_layout(int B, int Sd, int U) {
  SpellClBwdWs w;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t o = 0;
  w.dgb1 = o; o += up((size_t)B * U * 4 * Sd * 2);
  w.dgb2 = o; o += up((size_t)B * U * 4 * Sd * 2);
  w.dh1 = o; o += up((size_t)B * U * Sd * 4);
  w.total = o;
  return w;
}
long long ssasr_speller_cl_bwd_ws_bytes(int B, int Tp, int E, int Sd, int M, int U) {
  if (!spell_cl_supported(B, Tp, E, Sd, M)) return 0;
  return (long long)spell_cl_bwd_ws_layout(B, Sd, U).total;
}

int ssasr_speller_bwd_f32(const ssasr_speller_bwd_args* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int B = a->B, Tp = a->Tp, E = a->E, Sd = a->Sd, M = a->M, C = a->C, U = a->U;
  const int K1 = Sd + E, X1 = K1 + Sd, X2 = 2 * Sd;
  int rc;
  // through the character projection, all steps at once
  rc = gemm_f32(st, B * U, Sd, C, a->dlogits, C, 1, a->wc, Sd, 0, a->dh2all, Sd, nullptr, 0, 0);
  if (rc) return rc;
  const bool tc0 = a->w1catT_bf && a->w2catT_bf && a->wsA && a->wsB && a->BUp >= (long long)B * U && a->BUp % 8 == 0 &&
                   a->BTp >= (long long)B * Tp && a->BTp % 8 == 0 && Sd % 8 == 0 && E % 8 == 0 && M % 8 == 0;
  // out[Mo,No] = X^T Y over R rows (X [R,Mo] ldx, Y [R,No] ldy): tensor cores through transposed bf16 copies, or fp32
  auto xty = [&](cudaStream_t st, const float* X, int ldx, int Mo, const float* Y, int ldy, int No, long long R, long long Rp, float* out) -> int {
    if (tc0) {     // row-major bf16 copies, MN-major tcgen05 operands
      const int Mp = (Mo + 7) / 8 * 8, Np = (No + 7) / 8 * 8;
      int r = cvt_bf16(st, X, ldx, a->wsA, Mp, R, Mo);
      if (r) return r;
      r = cvt_bf16(st, Y, ldy, a->wsB, Np, R, No);
      if (r) return r;
      return gemm_bf16_tc_tn(st, Mo, No, (int)R, a->wsA, Mp, 0, a->wsB, Np, 0, out, No, 0);
    }
    return gemm_f32(st, Mo, No, (int)R, X, ldx, 0, Y, ldy, 0, out, No, nullptr, 0, 0);
  };
  const size_t attn_smem = attn_bwd_smem_bytes(Sd, M, Tp, E);
  SSASR_REQUIRE(attn_smem <= 200 * 1024 && (size_t)U * OA_FB * sizeof(float) <= 48 * 1024, "speller bwd: attention working set too large (Tp=%d, U=%d)", Tp, U);
  if (attn_smem > 48 * 1024)
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem));
  const int cell_blocks = (B * Sd + 255) / 256;
  const bool tc = a->w1catT_bf && a->w2catT_bf && a->wsA && a->wsB && X1 % 8 == 0 && X2 % 8 == 0 && a->BUp >= (long long)B * U &&
                  a->BUp % 8 == 0;
  // dx = dG_t @ Wcat (through the cells' input weights), fp32 SIMT or tcgen05
  __nv_bfloat16* dgb = tc ? (__nv_bfloat16*)a->wsA : nullptr;     // [B,4Sd] bf16 copy of the step's gate gradients
  // bf16 mode: the layer-2 backward chain depends on nothing but dh2all and itself -- it runs ahead on a second stream
  // (own gate-gradient scratch, one dxin2 block per step) and hands dh1(t) to the layer-1 / attention chain by event
  SideStream* side = (tc && a->dual_stream && U > 1) ? side_stream(U + 2) : nullptr;
  const bool dual = side != nullptr;
  cudaStream_t sb = dual ? side->s : st;
  __nv_bfloat16* dgb2 = dual ? dgb + (size_t)B * 4 * Sd : dgb;
  auto dxin2_at = [&](int t) { return dual ? a->dxin2 + (size_t)t * B * X2 : a->dxin2; };
  const bool cl = tc && a->dual_stream && a->cl_ws && a->cl_ws_bwd && a->w1cat_bf && a->w2cat_bf && spell_cl_supported(B, Tp, E, Sd, M) &&
                  a->cl_ws_bwd_bytes >= (long long)spell_cl_bwd_ws_layout(B, Sd, U).total;
  const int chain_splits = (tc && !cl) ? step_gemm_splits() : 1;        // see ssasr_speller_fwd_f32
  if (chain_splits > 1) SSASR_CHECK_CUDA(cudaMemsetAsync(a->dxin1, 0, sizeof(float) * (size_t)B * U * X1, st));
  auto dgrad_gemm = [&](cudaStream_t st, const __nv_bfloat16* dgb, const float* dg, int N, const float* w, const void* wT_bf,
                        float* out, int ldo, int splits = 1) -> int {
    if (tc) return gemm_bf16_tc(st, B, N, 4 * Sd, dgb, 4 * Sd, 0, wT_bf, 4 * Sd, 0, out, ldo, nullptr, 0, 0, splits);
    return gemm_f32(st, B, N, 4 * Sd, dg, U * 4 * Sd, 1, w, N, 0, out, ldo, nullptr, 0, 0);
  };
  if (dual) SSASR_HANDOVER(side->ev[U], st, sb);
  // dW = dG_all^T @ X_all over all B*U rows
  auto wgrad_gemm = [&](cudaStream_t st, const float* dg, const float* x, int N, float* out) -> int {
    if (tc) {
      int r = cvt_bf16(st, dg, 4 * Sd, a->wsA, 4 * Sd, (long long)B * U, 4 * Sd);
      if (r) return r;
      r = cvt_bf16(st, x, N, a->wsB, N, (long long)B * U, N);
      if (r) return r;
      return gemm_bf16_tc_tn(st, 4 * Sd, N, B * U, a->wsA, 4 * Sd, 0, a->wsB, N, 0, out, N, 0);
    }
    return gemm_f32(st, 4 * Sd, N, B * U, dg, 4 * Sd, 0, x, N, 0, out, N, nullptr, 0, 0);
  };
  // layer-2 cell backward + its dgrad GEMM for step t on stream s2 (two streams: own gate-gradient scratch, one dxin2 block per step)
  auto layer2_bwd = [&](int t, cudaStream_t s2) -> int {
    const int last = (t == U - 1);
    {
      ProfScope ps(F_POINTWISE, s2);
      cell_bwd_kernel<<<cell_blocks, 256, 0, s2>>>(B, Sd, a->act2 + (size_t)t * 4 * Sd, (long long)U * 4 * Sd,
                                                   a->c2 + (size_t)t * Sd, (long long)U * Sd,
                                                   t ? a->c2 + (size_t)(t - 1) * Sd : nullptr, (long long)U * Sd,
                                                   a->dh2all + (size_t)t * Sd, (long long)U * Sd,
                                                   last ? nullptr : dxin2_at(t + 1) + Sd, (long long)X2, nullptr, 0, a->dc2s, last, dgb2);
    }
    return dgrad_gemm(s2, dgb2, a->act2 + (size_t)t * 4 * Sd, X2, a->w2cat, a->w2catT_bf, dxin2_at(t), X2);
  };
  // layer-1 cell backward of step t as its own launch
  auto cell1_bwd = [&](int t) {
    const int last = (t == U - 1);
    ProfScope ps(F_POINTWISE, st);
    cell_bwd_kernel<<<cell_blocks, 256, 0, st>>>(B, Sd, a->act1 + (size_t)t * 4 * Sd, (long long)U * 4 * Sd,
                                                 a->c1 + (size_t)t * Sd, (long long)U * Sd,
                                                 t ? a->c1 + (size_t)(t - 1) * Sd : nullptr, (long long)U * Sd, dxin2_at(t),
                                                 (long long)X2, last ? nullptr : a->dxin1 + (size_t)(t + 1) * X1 + K1,
                                                 (long long)U * X1, last ? nullptr : a->dh1att, (long long)Sd, a->dc1s, last, dgb);
  };
  // attention backward of step t; fuse_prev: the layer-1 cell backward of step t-1 (whose h1 was this step's query) in its epilogue
  auto attn_bwd = [&](int t, bool fuse_prev) {
    AttnBwd g{};
    g.Tp = Tp; g.E = E; g.Sd = Sd; g.M = M;
    g.dctx = a->dxin1 + (size_t)t * X1 + Sd; g.dctx_ld = (long long)U * X1;
    g.alpha = a->alpha + (size_t)t * Tp; g.alpha_ld = (long long)U * Tp;
    g.q = a->q + (size_t)t * M; g.q_ld = (long long)U * M;
    g.phi_w = a->phi_w; g.psi = a->psi; g.enc = a->enc; g.enc_lens = a->enc_lens;
    g.dalpha = nullptr; g.dalpha_ld = 0;
    g.de = a->de_all + (size_t)t * Tp; g.de_ld = (long long)U * Tp;
    g.dqpre = a->dqpre + (size_t)t * M; g.dqpre_ld = (long long)U * M;
    g.dh1att = a->dh1att;
    if (fuse_prev) {
      g.cb_act = a->act1 + (size_t)(t - 1) * 4 * Sd; g.cba_ld = (long long)U * 4 * Sd;
      g.cb_c = a->c1 + (size_t)(t - 1) * Sd; g.cbc_ld = (long long)U * Sd;
      g.cb_cprev = t > 1 ? a->c1 + (size_t)(t - 2) * Sd : nullptr; g.cbp_ld = (long long)U * Sd;
      g.cb_dh_a = dxin2_at(t - 1); g.cbda_ld = X2;
      g.cb_dh_b = a->dxin1 + (size_t)t * X1 + K1; g.cbdb_ld = (long long)U * X1;
      g.cb_dcstate = a->dc1s;
      g.cb_dgb = dgb;
    }
    ProfScope ps(F_ATTN_BWD, st);
    attn_bwd_kernel<<<B, 256, attn_smem, st>>>(g);
  };
  if (cl) {
    // ---- cluster-persistent path: layer-2 chain, its input gradient in one product, layer-1 chain + attention backward ----
    const SpellClWs wl = spell_cl_ws_layout(B, Tp, Sd, M, C, U);
    const SpellClBwdWs bl = spell_cl_bwd_ws_layout(B, Sd, U);
    const uint8_t* fws = (const uint8_t*)a->cl_ws;
    uint8_t* bws = (uint8_t*)a->cl_ws_bwd;
    __nv_bfloat16* dgb1 = (__nv_bfloat16*)(bws + bl.dgb1);
    __nv_bfloat16* dgb2 = (__nv_bfloat16*)(bws + bl.dgb2);
    float* dh1 = (float*)(bws + bl.dh1);
    SpellClBwdArgs g = {};
    g.B = B; g.U = U; g.Tp = Tp; g.t0 = 0; g.t1 = U; g.enc_lens = a->enc_lens;
    g.wcat_bf = a->w2cat_bf; g.X = X2; g.Kcol = Sd;
    g.act = a->act2; g.act_ldb = (long long)U * 4 * Sd; g.act_ldt = 4 * Sd;
    g.c = a->c2; g.c_ldb = (long long)U * Sd; g.c_ldt = Sd;
    g.dh_in = a->dh2all; g.dh_ldb = (long long)U * Sd; g.dh_ldt = Sd;
    g.dgb = dgb2; g.dgb_ldb = (long long)U * 4 * Sd; g.dgb_ldt = 4 * Sd;
    rc = spell_cl_bwd(st, g);
    if (rc) return rc;
    // dh1(t) from layer 2, all steps: dG2 W_ih2
    rc = gemm_bf16_tc(st, B * U, Sd, 4 * Sd, dgb2, 4 * Sd, 0, a->w2catT_bf, 4 * Sd, 0, dh1, Sd, nullptr, 0);
    if (rc) return rc;
    g.wcat_bf = a->w1cat_bf; g.X = X1; g.Kcol = K1;
    g.phi_bf = fws + wl.phi_bf; g.P_bf = fws + wl.p_bf; g.psi_bf = fws + wl.psi_bf;
    g.act = a->act1; g.c = a->c1;
    g.dh_in = dh1;
    g.dgb = dgb1;
    g.alpha = a->alpha; g.al_ldb = (long long)U * Tp; g.al_ldt = Tp;
    g.q = a->q; g.q_ldb = (long long)U * M; g.q_ldt = M;
    g.de = a->de_all; g.de_ldb = (long long)U * Tp; g.de_ldt = Tp;
    g.dqpre = a->dqpre; g.dq_ldb = (long long)U * M; g.dq_ldt = M;
    rc = spell_cl_bwd(st, g);
    if (rc) return rc;
    // gradient of the step inputs [emb ; ctx] of all steps: dG1 [W_emb | W_ctx] (the recurrent part stayed inside the kernel)
    rc = gemm_bf16_tc(st, B * U, K1, 4 * Sd, dgb1, 4 * Sd, 0, a->w1catT_bf, 4 * Sd, 0, a->dxin1, X1, nullptr, 0);
    if (rc) return rc;
  } else if (!dual) {
    for (int t = U - 1; t >= 0; --t) {
      rc = layer2_bwd(t, st);
      if (rc) return rc;
      cell1_bwd(t);
      rc = dgrad_gemm(st, dgb, a->act1 + (size_t)t * 4 * Sd, X1, a->w1cat, a->w1catT_bf, a->dxin1 + (size_t)t * X1, U * X1,
                      chain_splits);
      if (rc) return rc;
      attn_bwd(t, false);
    }
  } else {
    // Two streams.  `sb`: the layer-2 chain, issued one step ahead of its consumer; `st`: layer-1 dgrad GEMM -> attention
    // backward, whose epilogue runs the layer-1 cell backward of the step below it (one launch less on the dependent chain).
    rc = layer2_bwd(U - 1, sb);
    if (rc) return rc;
    SSASR_CHECK_CUDA(cudaEventRecord(side->ev[U - 1], sb));
    for (int t = U - 1; t >= 0; --t) {
      if (t > 0) {
        rc = layer2_bwd(t - 1, sb);
        if (rc) return rc;
        SSASR_CHECK_CUDA(cudaEventRecord(side->ev[t - 1], sb));
      }
      if (t == U - 1) {
        SSASR_CHECK_CUDA(cudaStreamWaitEvent(st, side->ev[t], 0));
        cell1_bwd(t);
      }
      rc = dgrad_gemm(st, dgb, a->act1 + (size_t)t * 4 * Sd, X1, a->w1cat, a->w1catT_bf, a->dxin1 + (size_t)t * X1, U * X1,
                      chain_splits);
      if (rc) return rc;
      if (t > 0) SSASR_CHECK_CUDA(cudaStreamWaitEvent(st, side->ev[t - 1], 0));   // dh1(t-1) from layer 2
      attn_bwd(t, t > 0);
    }
  }
  // attention memory gradients, accumulated over all steps at once; then denc += dpsi_pre @ Wpsi completes what the encoder needs
  // (two streams: the dpsi branch -- accumulate, tanh', bf16 copy -- runs beside the denc accumulation)
  cudaStream_t sp = dual ? sb : st;
  if (dual) SSASR_HANDOVER(side->ev[U], st, sp);
  rc = attn_outer_accum(sp, B, U, Tp, M, a->de_all, a->q, M, (long long)U * M, a->dpsi, a->enc_lens);
  if (rc) return rc;
  {
    ProfScope ps(F_POINTWISE, sp);
    dtanh_inplace_kernel<<<256, 256, 0, sp>>>(a->dpsi, a->psi, (size_t)B * Tp * M);
  }
  if (tc0) {
    rc = cvt_bf16(sp, a->dpsi, M, a->wsA, M, (long long)B * Tp, M);
    if (rc) return rc;
  }
  rc = attn_outer_accum(st, B, U, Tp, E, a->alpha, a->dxin1 + Sd, X1, (long long)U * X1, a->denc, a->enc_lens);
  if (rc) return rc;
  if (tc0) {
    __nv_bfloat16* wT = (__nv_bfloat16*)a->wsB;              // [E, M]
    rc = cvt_bf16_t(st, a->psi_w, E, wT, M, M, E, 0, 0, 0, 0);
    if (rc) return rc;
    if (dual) SSASR_HANDOVER(side->ev[U], sp, st);
    rc = gemm_bf16_tc(st, B * Tp, E, M, a->wsA, M, 0, wT, M, 0, a->denc, E, nullptr, 1);
  } else {
    if (dual) SSASR_HANDOVER(side->ev[U], sp, st);
    rc = gemm_f32(st, B * Tp, E, M, a->dpsi, M, 1, a->psi_w, E, 0, a->denc, E, nullptr, 1, 0);
  }
  if (rc) return rc;
  // everything below is consumed by the optimiser only
  cudaStream_t sw = st;
  if (dual && a->wgrad_stream && (cudaStream_t)a->wgrad_stream != st) {
    sw = (cudaStream_t)a->wgrad_stream;
    SSASR_HANDOVER(side->ev[U + 1], st, sw);      // also orders the scratch (wsA / wsB) reuse below after the GEMM above
  }
  rc = xty(sw, a->dlogits, C, C, a->h2all, Sd, Sd, (long long)B * U, a->BUp, a->d_wc);
  if (rc) return rc;
  rc = colsum(sw, a->dlogits, a->d_bc, B * U, C, C, 0);
  if (rc) return rc;
  SSASR_CHECK_CUDA(cudaMemsetAsync(a->d_emb_w, 0, sizeof(float) * (size_t)C * Sd, sw));
  {
    ProfScope ps(F_POINTWISE, sw);
    emb_grad_add_kernel<<<(B * U * Sd + 255) / 256, 256, 0, sw>>>(B * U, Sd, a->dxin1, X1, a->tok_in, 1, a->d_emb_w);
  }
  // weight gradients, batched over all steps
  rc = wgrad_gemm(sw, a->act1, a->xin1, X1, a->d_w1cat);
  if (rc) return rc;
  rc = colsum(sw, a->act1, a->d_b1, B * U, 4 * Sd, 4 * Sd, 0);
  if (rc) return rc;
  rc = wgrad_gemm(sw, a->act2, a->xin2, X2, a->d_w2cat);
  if (rc) return rc;
  rc = colsum(sw, a->act2, a->d_b2, B * U, 4 * Sd, 4 * Sd, 0);
  if (rc) return rc;
  rc = xty(sw, a->dqpre, M, M, a->xin1 + K1, X1, Sd, (long long)B * U, a->BUp, a->d_phi_w);
  if (rc) return rc;
  rc = xty(sw, a->dpsi, M, M, a->enc, E, E, (long long)B * Tp, a->BTp, a->d_psi_w);
  if (rc) return rc;
  rc = colsum(sw, a->dpsi, a->d_psi_b, B * Tp, M, M, 0);
  if (rc) return rc;
  SSASR_LAUNCH_CHECK();
  return 0;
}

// ---- single-step module API (Attention.forward asr.py:343-392, Speller.forward asr.py:314-326 as called one step at a
// time by the reference's other trainers, text_autoencoder.py:52-94) -------------------------------------------------
// xrow scratch [B, Sd+E+Sd]: on return ctx = xrow[:, Sd:Sd+E]
int ssasr_attn_step_fwd(int B, int Tp, int E, int Sd, int M, const float* h, const float* phi_w, const float* psi,
                        const float* enc, const int* enc_lens, float* xrow, float* q, float* alpha, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AttnFwd f{};
  f.Tp = Tp; f.E = E; f.Sd = Sd; f.M = M;
  f.h1prev = h; f.h1_ld = Sd; f.phi_w = phi_w; f.psi = psi; f.enc = enc; f.enc_lens = enc_lens;
  f.emb_w = nullptr; f.tok = nullptr; f.tok_ld = 0;
  f.xin1 = xrow; f.xin1_ld = 2 * Sd + E; f.xin1b = nullptr; f.xin1b_ld = 0;
  f.x3h = nullptr; f.x3l = nullptr; f.x3_ld = 0;
  f.q = q; f.q_ld = M; f.alpha = alpha; f.alpha_ld = Tp;
  const size_t smem = attn_fwd_smem_bytes(Sd, M, Tp, E);
  SSASR_REQUIRE(smem <= 200 * 1024, "attn_step_fwd: working set too large (Tp=%d)", Tp);
  if (smem > 48 * 1024) SSASR_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(F_ATTN_FWD, st);
  attn_fwd_kernel<<<B, 256, smem, st>>>(f);
  SSASR_LAUNCH_CHECK();
  return 0;
}
// dctx [B,E], dalpha [B,Tp] or NULL -> dh [B,Sd], dqpre [B,M], de [B,Tp], denc [B,Tp,E] (overwritten), dpsi [B,Tp,M] (overwritten)
int ssasr_attn_step_bwd(int B, int Tp, int E, int Sd, int M, const float* dctx, const float* dalpha, const float* alpha,
                        const float* q, const float* phi_w, const float* psi, const float* enc, const int* enc_lens, float* de,
                        float* dqpre, float* dh, float* denc, float* dpsi, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AttnBwd g{};
  g.Tp = Tp; g.E = E; g.Sd = Sd; g.M = M;
  g.dctx = dctx; g.dctx_ld = E; g.alpha = alpha; g.alpha_ld = Tp; g.q = q; g.q_ld = M;
  g.phi_w = phi_w; g.psi = psi; g.enc = enc; g.enc_lens = enc_lens;
  g.dalpha = dalpha; g.dalpha_ld = Tp; g.de = de; g.de_ld = Tp; g.dqpre = dqpre; g.dqpre_ld = M; g.dh1att = dh;
  const size_t smem = attn_bwd_smem_bytes(Sd, M, Tp, E);
  SSASR_REQUIRE(smem <= 200 * 1024, "attn_step_bwd: working set too large (Tp=%d)", Tp);
  if (smem > 48 * 1024) SSASR_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(F_ATTN_BWD, st);
  attn_bwd_kernel<<<B, 256, smem, st>>>(g);
  int rc = attn_outer_accum(st, B, 1, Tp, E, alpha, dctx, E, E, denc, enc_lens);
  if (rc) return rc;
  rc = attn_outer_accum(st, B, 1, Tp, M, de, q, M, M, dpsi, enc_lens);
  if (rc) return rc;
  SSASR_LAUNCH_CHECK();
  return 0;
}
// gates [B,4S] interleaved pre-activations (in) -> activations (out); c_prev may be NULL (zero state)
int ssasr_lstmcell_fwd(int B, int S, float* gates, const float* c_prev, float* c_out, float* h_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_REQUIRE(S % 4 == 0, "lstmcell: S=%d must be a multiple of 4", S);
  ProfScope ps(F_POINTWISE, st);
  cell_fwd_kernel<<<(B * S + 255) / 256, 256, 0, st>>>(B, S, gates, 4 * S, c_prev, S, c_out, S, h_out, S, nullptr, 0, nullptr, 0, nullptr,
                                                       0, 0);
  SSASR_LAUNCH_CHECK();
  return 0;
}
// act [B,4S] activations (in) -> dL/d(pre-activations) (out); dc [B,S]: in = dL/dc_out, out = dL/dc_prev
int ssasr_lstmcell_bwd(int B, int S, float* act, const float* c, const float* c_prev, const float* dh, float* dc, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(F_POINTWISE, st);
  cell_bwd_kernel<<<(B * S + 255) / 256, 256, 0, st>>>(B, S, act, 4 * S, c, S, c_prev, S, dh, S, nullptr, 0, nullptr, 0, dc, 0, nullptr);
  SSASR_LAUNCH_CHECK();
  return 0;
}
// d *= (1 - y*y)  (backward of tanh given its output y)
int ssasr_dtanh_mul(float* d, const float* y, long long n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(F_POINTWISE, st);
  dtanh_inplace_kernel<<<256, 256, 0, st>>>(d, y, (size_t)n);
  SSASR_LAUNCH_CHECK();
  return 0;
}
int ssasr_colsum(const float* src, float* out, int R, int C, int ld, int accumulate, void* stream) {
  return colsum((cudaStream_t)stream, src, out, R, C, ld, accumulate);
}

// Fused loss of trainer.py:426-434: per-utterance sum of CE(ignore_index=0) / count(y != 0), batch mean.
//   logits [B,U,C], y int64 [B,L] (label of step t is y[:, t+1]); loss_b [B] scratch; loss_out scalar;
//   dlogits [B,U,C] or null: d(loss * grad_scale)/dlogits.
int ssasr_ce_loss_f32(const float* logits, const long long* y, int B, int U, int C, int L, float* loss_b, float* loss_out,
                      float* dlogits, float grad_scale, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_REQUIRE(L >= U + 1, "ce_loss: targets have %d columns, need at least U+1=%d", L, U + 1);
  ProfScope ps(F_CE, st);
  ce_kernel<<<B, 128, 0, st>>>(B, U, C, L, logits, y, loss_b, dlogits, grad_scale);
  mean_kernel<<<1, 32, 0, st>>>(loss_b, B, loss_out);
  SSASR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

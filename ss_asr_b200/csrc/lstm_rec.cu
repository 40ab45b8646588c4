// Persistent recurrent BLSTM kernels, fp32 path (exact: greedy decode / validation / tight parity).
//
// Reference semantics: nn.LSTM(bidirectional) on a packed sequence (asr.py:410-418, layers 1-3) and
// encoder.blstm_4, the seq-first quirk (asr.py:237-238,262); SURVEY.md Appendix A.
//
// Layout ("interleaved gates"): every gate buffer has 8S columns, col = dir*4S + unit*4 + gate
// (gate order i,f,g,o), so the four gates of one hidden unit are one float4 and a CTA that owns a
// slice of units owns a contiguous slice of columns.  Rows are addressed by (seq step, batch
// element) through two row strides, which lets one kernel run both the time-major layers 1-3
// (seq = time, batch = utterance, per-utterance lengths) and layer 4 (seq = utterance, batch =
// frame, no lengths).
//
// Structure: ONE cooperative launch per layer direction-pair.  grid = (S/UPC, 2); a CTA keeps its
// 4*UPC rows of W_hh resident in shared memory for the whole sequence, the hidden state is exchanged
// through L2 (the output tensor itself) and the CTAs of one direction meet at a per-direction
// monotonic-counter barrier once per step.  No per-step launches, no pack/unpack copies: lengths
// are masks.
#include <stdlib.h>

#include "common.cuh"
#include <cuda_bf16.h>

namespace ssasr {

constexpr int TBN = 64;        // batch rows per tile
constexpr int KG = 4;          // k-split groups (256 threads = 64 rows x 4 groups)

__device__ __forceinline__ void dir_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

struct RecFwdParams {
  float* xp;            // [rows, 8S]  pre-activations in / gate activations out
  const float* whh;     // [2, 4S, S]  interleaved rows
  float* hout;          // rows x 2S
  float* cbuf;          // rows x 2S
  const int* lens;      // [n_batch] or null
  int S, UPC, n_seq, n_batch;
  long long xs_seq, xs_batch, hs_seq, hs_batch;   // row strides
  unsigned* bar;        // [2] zero-initialised
};

template <int UPC>
__global__ void __launch_bounds__(256) rec_fwd_kernel(RecFwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int S = p.S, SP = S + 4;
  float* Wsl = smem;                       // [4*UPC][S]
  float* hs = Wsl + 4 * UPC * S;           // [TBN][SP]
  float* red = hs + TBN * SP;              // [KG][TBN][4*UPC]
  const int dir = blockIdx.y, u0 = blockIdx.x * UPC;
  const int tid = threadIdx.x, nl = tid % TBN, kq = tid / TBN;
  const float* whh = p.whh + (size_t)dir * 4 * S * S + (size_t)u0 * 4 * S;
  for (int i = tid; i < 4 * UPC * S / 4; i += 256)
    reinterpret_cast<float4*>(Wsl)[i] = reinterpret_cast<const float4*>(whh)[i];
  __syncthreads();
  const int kper = S / KG;                 // S % 16 == 0 required
  const unsigned G = gridDim.x;

  for (int s = 0; s < p.n_seq; ++s) {
    const int t = dir == 0 ? s : p.n_seq - 1 - s;
    const int tp = dir == 0 ? t - 1 : t + 1;
    for (int n0 = 0; n0 < p.n_batch; n0 += TBN) {
      const int n = n0 + nl;
      float acc[4 * UPC];
#pragma unroll
      for (int i = 0; i < 4 * UPC; ++i) acc[i] = 0.0f;
      if (s > 0) {
        __syncthreads();   // previous tile's readers of hs / red are done
        for (int i = tid; i < TBN * (S / 4); i += 256) {
          const int r = i / (S / 4), c4 = i % (S / 4);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n0 + r < p.n_batch)
            v = __ldcg(reinterpret_cast<const float4*>(p.hout + ((size_t)tp * p.hs_seq + (size_t)(n0 + r) * p.hs_batch) * 2 * S +
                                                       dir * S) + c4);
          *reinterpret_cast<float4*>(hs + r * SP + c4 * 4) = v;
        }
        __syncthreads();
        const float* hrow = hs + nl * SP + kq * kper;
        const float* wbase = Wsl + kq * kper;
#pragma unroll 2
        for (int k = 0; k < kper; k += 4) {
          const float4 h4 = *reinterpret_cast<const float4*>(hrow + k);
#pragma unroll
          for (int r = 0; r < 4 * UPC; ++r) {
            const float4 w4 = *reinterpret_cast<const float4*>(wbase + r * S + k);
            acc[r] = fmaf(h4.x, w4.x, acc[r]);
            acc[r] = fmaf(h4.y, w4.y, acc[r]);
            acc[r] = fmaf(h4.z, w4.z, acc[r]);
            acc[r] = fmaf(h4.w, w4.w, acc[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < 4 * UPC; ++r) red[(kq * TBN + nl) * (4 * UPC) + r] = acc[r];
        __syncthreads();
      }
      // finalize: thread (nl, kq) owns units kq, kq+KG, ... of this CTA's slice
      if (n < p.n_batch) {
        const size_t xrow = (size_t)t * p.xs_seq + (size_t)n * p.xs_batch;
        const size_t hrow_o = (size_t)t * p.hs_seq + (size_t)n * p.hs_batch;
        const bool valid = p.lens ? (t < p.lens[n]) : true;
        for (int ul = kq; ul < UPC; ul += KG) {
          const int u = u0 + ul;
          float4* xptr = reinterpret_cast<float4*>(p.xp + xrow * 8 * S + (size_t)dir * 4 * S + (size_t)u * 4);
          float4 g = *xptr;
          if (s > 0) {
#pragma unroll
            for (int q = 0; q < KG; ++q) {
              const float* rp = red + (q * TBN + nl) * (4 * UPC) + ul * 4;
              g.x += rp[0]; g.y += rp[1]; g.z += rp[2]; g.w += rp[3];
            }
          }
          float hval = 0.f, cval = 0.f;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) {
            float cprev = 0.f;
            if (s > 0) cprev = p.cbuf[((size_t)tp * p.hs_seq + (size_t)n * p.hs_batch) * 2 * S + dir * S + u];
            a.x = sigmoidf_acc(g.x); a.y = sigmoidf_acc(g.y); a.z = tanhf(g.z); a.w = sigmoidf_acc(g.w);
            cval = a.y * cprev + a.x * a.z;
            hval = a.w * tanhf(cval);
          }
          *xptr = a;
          p.hout[hrow_o * 2 * S + dir * S + u] = hval;
          p.cbuf[hrow_o * 2 * S + dir * S + u] = cval;
        }
      }
    }
    if (s + 1 < p.n_seq) dir_barrier(p.bar + dir, (unsigned)(s + 1) * G);
  }
}

struct RecBwdParams {
  float* act;           // [rows, 8S] gate activations in / gate gradients (pre-activation) out
  const float* whhT;    // [2, S, 4S]  whhT[d][u][r] = whh[d][r][u]
  const float* cbuf;    // rows x 2S
  const float* dhout;   // rows x 2S  gradient w.r.t. hout
  float* dcstate;       // [n_batch, 2S] scratch (running dc), any initial content
  const int* lens;
  int S, UPC, n_seq, n_batch;
  long long xs_seq, xs_batch, hs_seq, hs_batch;
  unsigned* bar;
};

template <int UPC>
__global__ void __launch_bounds__(256) rec_bwd_kernel(RecBwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int S = p.S, SP = S + 4;
  float* Wsl = smem;                       // [UPC][4S]
  float* gs = Wsl + UPC * 4 * S;           // [TBN][SP]  one S-wide chunk of the previous step's dG
  float* red = gs + TBN * SP;              // [KG][TBN][UPC]
  const int dir = blockIdx.y, u0 = blockIdx.x * UPC;
  const int tid = threadIdx.x, nl = tid % TBN, kq = tid / TBN;
  const float* wt = p.whhT + (size_t)dir * S * 4 * S + (size_t)u0 * 4 * S;
  for (int i = tid; i < UPC * S; i += 256) reinterpret_cast<float4*>(Wsl)[i] = reinterpret_cast<const float4*>(wt)[i];
  __syncthreads();
  const int kper = S / KG;
  const unsigned G = gridDim.x;

  for (int s = 0; s < p.n_seq; ++s) {
    const int t = dir == 0 ? p.n_seq - 1 - s : s;       // reverse of the forward order
    const int tn = dir == 0 ? t + 1 : t - 1;            // step processed just before (its dG feeds dh)
    const int tp = dir == 0 ? t - 1 : t + 1;            // forward-order predecessor (c_prev)
    const bool has_prev = dir == 0 ? (t > 0) : (t < p.n_seq - 1);
    for (int n0 = 0; n0 < p.n_batch; n0 += TBN) {
      const int n = n0 + nl;
      float acc[UPC];
#pragma unroll
      for (int i = 0; i < UPC; ++i) acc[i] = 0.0f;
      if (s > 0) {
        for (int kc = 0; kc < 4; ++kc) {               // 4S reduction in chunks of S
          __syncthreads();
          for (int i = tid; i < TBN * (S / 4); i += 256) {
            const int r = i / (S / 4), c4 = i % (S / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + r < p.n_batch)
              v = __ldcg(reinterpret_cast<const float4*>(p.act + ((size_t)tn * p.xs_seq + (size_t)(n0 + r) * p.xs_batch) * 8 * S +
                                                         (size_t)dir * 4 * S + (size_t)kc * S) + c4);
            *reinterpret_cast<float4*>(gs + r * SP + c4 * 4) = v;
          }
          __syncthreads();
          const float* grow = gs + nl * SP + kq * kper;
          const float* wbase = Wsl + kc * S + kq * kper;
#pragma unroll 2
          for (int k = 0; k < kper; k += 4) {
            const float4 g4 = *reinterpret_cast<const float4*>(grow + k);
#pragma unroll
            for (int ul = 0; ul < UPC; ++ul) {
              const float4 w4 = *reinterpret_cast<const float4*>(wbase + ul * 4 * S + k);
              acc[ul] = fmaf(g4.x, w4.x, acc[ul]);
              acc[ul] = fmaf(g4.y, w4.y, acc[ul]);
              acc[ul] = fmaf(g4.z, w4.z, acc[ul]);
              acc[ul] = fmaf(g4.w, w4.w, acc[ul]);
            }
          }
        }
        __syncthreads();
#pragma unroll
        for (int ul = 0; ul < UPC; ++ul) red[(kq * TBN + nl) * UPC + ul] = acc[ul];
        __syncthreads();
      }
      if (n < p.n_batch) {
        const size_t xrow = (size_t)t * p.xs_seq + (size_t)n * p.xs_batch;
        const size_t hrow = (size_t)t * p.hs_seq + (size_t)n * p.hs_batch;
        const bool valid = p.lens ? (t < p.lens[n]) : true;
        for (int ul = kq; ul < UPC; ul += KG) {
          const int u = u0 + ul;
          float4* aptr = reinterpret_cast<float4*>(p.act + xrow * 8 * S + (size_t)dir * 4 * S + (size_t)u * 4);
          float* dcs = p.dcstate + (size_t)n * 2 * S + dir * S + u;
          float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
          float dc_out = 0.f;
          if (valid) {
            float dh = p.dhout[hrow * 2 * S + dir * S + u];
            float dc_rec = 0.f;
            if (s > 0) {
#pragma unroll
              for (int q = 0; q < KG; ++q) dh += red[(q * TBN + nl) * UPC + ul];
              dc_rec = *dcs;
            }
            const float4 a = *aptr;
            const float c = p.cbuf[hrow * 2 * S + dir * S + u];
            float cprev = 0.f;
            if (has_prev) {
              // forward-order predecessor is valid whenever this step is (fwd: t-1 < len; rev: only if t+1 < len)
              const bool pv = p.lens ? (tp < p.lens[n]) : true;
              if (pv) cprev = p.cbuf[((size_t)tp * p.hs_seq + (size_t)n * p.hs_batch) * 2 * S + dir * S + u];
            }
            const float tc = tanhf(c);
            const float dc = dh * a.w * (1.f - tc * tc) + dc_rec;
            dg.w = dh * tc * a.w * (1.f - a.w);
            dg.x = dc * a.z * a.x * (1.f - a.x);
            dg.z = dc * a.x * (1.f - a.z * a.z);
            dg.y = dc * cprev * a.y * (1.f - a.y);
            dc_out = dc * a.y;
          }
          *aptr = dg;
          *dcs = dc_out;
        }
      }
    }
    if (s + 1 < p.n_seq) dir_barrier(p.bar + dir, (unsigned)(s + 1) * G);
  }
}

// ---------------------------------------------------------------------------------------------
// weight packing (PyTorch [4S,K] gate-major rows -> interleaved rows) and gradient unpacking
// ---------------------------------------------------------------------------------------------
__global__ void pack_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int S, int K, int dst_ld, int dst_col0) {
  // dst[(u*4+g), dst_col0 + k] = src[(g*S+u), k]
  const size_t total = (size_t)4 * S * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int rr = (int)(i / K);
    const int u = rr >> 2, g = rr & 3;
    dst[(size_t)rr * dst_ld + dst_col0 + k] = src[((size_t)g * S + u) * K + k];
  }
}
__global__ void pack_bias_kernel(const float* __restrict__ b1, const float* __restrict__ b2, float* __restrict__ dst, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * S) {
    const int u = i >> 2, g = i & 3;
    dst[i] = b1[g * S + u] + (b2 ? b2[g * S + u] : 0.f);
  }
}
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C) {
  // dst[c][r] = src[r][c]
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? src[(size_t)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}
// grad[(g*S+u), k] += src[(u*4+g), src_col0 + k]
__global__ void unpack_rows_add_kernel(const float* __restrict__ src, float* __restrict__ grad, int S, int K, int src_ld,
                                       int src_col0) {
  const size_t total = (size_t)4 * S * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int row = (int)(i / K);          // gate-major row g*S+u
    const int g = row / S, u = row % S;
    grad[i] += src[(size_t)(u * 4 + g) * src_ld + src_col0 + k];
  }
}
__global__ void unpack_bias_add_kernel(const float* __restrict__ src, float* __restrict__ g1, float* __restrict__ g2, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * S) {
    const int g = i / S, u = i % S;
    const float v = src[u * 4 + g];
    g1[i] += v;
    if (g2) g2[i] += v;
  }
}
// out[c] (+)= sum_r src[r*ld + c].  grid (C/32, row chunks): each block reduces a 32-column x CS_ROWS-row panel
// (coalesced 128-B row segments) and adds its 32 partial sums with atomics.
constexpr int CS_ROWS = 512;
__global__ void colsum_kernel(const float* __restrict__ src, float* __restrict__ out, int R, int C, int ld) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * CS_ROWS;
  const int r1 = min(R, r0 + CS_ROWS);
  float s = 0.f;
  if (c < C)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s += src[(size_t)r * ld + c];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += part[i][threadIdx.x];
    atomicAdd(out + c, v);
  }
}

int colsum(cudaStream_t st, const float* src, float* out, int R, int C, int ld, int accumulate) {
  if (!accumulate) SSASR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
  if (R <= 0 || C <= 0) return 0;
  ProfScope ps(F_POINTWISE, st);
  colsum_kernel<<<dim3((C + 31) / 32, (R + CS_ROWS - 1) / CS_ROWS), dim3(32, 8), 0, st>>>(src, out, R, C, ld);
  SSASR_LAUNCH_CHECK();
  return 0;
}

// ---- bf16 training path: every packed operand of one BLSTM layer in ONE launch (was: 8 pack launches + 4 conversions) -----
// blocks [0, nt_ih): 32 x 32 tiles of W_ih -> wih_bf [8S,Kp] (packed rows, zero-padded columns) and wihT_bf [K,8S];
// blocks [nt_ih, nt_ih + nt_hh): tiles of W_hh -> whh_bf [8S,S] and whhT_bf [2S,4S]; the remaining blocks: bias_p = b_ih + b_hh.
struct PackBf16 {
  const float* wih[2]; const float* whh[2]; const float* bih[2]; const float* bhh[2];
  int S, K, Kp, nt_ih, nt_hh;
  float* bias_p; __nv_bfloat16 *wih_bf, *whh_bf, *wihT_bf, *whhT_bf;
};
__global__ void __launch_bounds__(256) pack_blstm_bf16_kernel(PackBf16 a) {
  __shared__ float tile[32][33];
  const int S = a.S, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8 threads
  int b = blockIdx.x;
  if (b < a.nt_ih + a.nt_hh) {
    const bool ih = b < a.nt_ih;
    if (!ih) b -= a.nt_ih;
    const int C = ih ? a.K : S;                        // source columns
    const int ct = ih ? (a.Kp + 31) / 32 : (S + 31) / 32;
    const int c0 = (b % ct) * 32, pr0 = (b / ct) * 32; // packed-row tile never straddles a direction (4S % 32 == 0)
    const int d = pr0 / (4 * S);
    const float* src = ih ? a.wih[d] : a.whh[d];
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
      const int r4 = (pr0 + i) % (4 * S), u = r4 >> 2, g = r4 & 3, c = c0 + tx;
      const float v = c < C ? src[((size_t)g * S + u) * C + c] : 0.f;
      tile[i][tx] = v;
      if (ih) { if (c < a.Kp) a.wih_bf[(size_t)(pr0 + i) * a.Kp + c] = __float2bfloat16(v); }
      else if (c < S) a.whh_bf[(size_t)(pr0 + i) * S + c] = __float2bfloat16(v);
    }
    __syncthreads();
#pragma unroll
    for (int i = ty; i < 32; i += 8) {                 // transposed copies: column c0 + i, packed rows pr0 + tx
      const int c = c0 + i;
      if (c >= C) continue;
      const __nv_bfloat16 v = __float2bfloat16(tile[tx][i]);
      if (ih) a.wihT_bf[(size_t)c * 8 * S + pr0 + tx] = v;
      else a.whhT_bf[((size_t)d * S + c) * 4 * S + (pr0 % (4 * S)) + tx] = v;
    }
  } else {
    const int i = (b - a.nt_ih - a.nt_hh) * 256 + threadIdx.x;
    if (i < 8 * S) {
      const int d = i / (4 * S), r4 = i % (4 * S), u = r4 >> 2, g = r4 & 3;
      a.bias_p[i] = a.bih[d][g * S + u] + a.bhh[d][g * S + u];
    }
  }
}

// all eight PyTorch-layout gradients of one BLSTM layer WRITTEN (not added) in one launch
struct UnpackSet {
  const float *dwih_p, *dbias_p, *dwhh_p;
  float* gwih[2]; float* gwhh[2]; float* gbih[2]; float* gbhh[2];
  int S, K;
};
__global__ void __launch_bounds__(256) unpack_blstm_set_kernel(UnpackSet a) {
  const int S = a.S, K = a.K;
  const size_t n_ih = (size_t)8 * S * K, n_hh = (size_t)8 * S * S, total = n_ih + n_hh + 8 * S;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n_ih + n_hh) {
      const bool ih = i < n_ih;
      const size_t j = ih ? i : i - n_ih;
      const int C = ih ? K : S;
      const int c = (int)(j % C);
      const int row = (int)(j / C);                    // destination row: d*4S + g*S + u
      const int d = row / (4 * S), rg = row % (4 * S), g = rg / S, u = rg % S;
      const float v = (ih ? a.dwih_p : a.dwhh_p)[((size_t)d * 4 * S + u * 4 + g) * C + c];
      (ih ? a.gwih[d] : a.gwhh[d])[(size_t)rg * C + c] = v;
    } else {
      const int r = (int)(i - n_ih - n_hh);
      const int d = r / (4 * S), rg = r % (4 * S), g = rg / S, u = rg % S;
      const float v = a.dbias_p[d * 4 * S + u * 4 + g];
      a.gbih[d][rg] = v;
      a.gbhh[d][rg] = v;
    }
  }
}

static int pick_upc(int S) {
  // smallest UPC in {4, 8} such that both directions fit co-resident on the device
  if (2 * (S / 4) <= sm_count()) return 4;
  return 8;
}

template <int UPC>
static int launch_fwd(RecFwdParams& p, cudaStream_t st) {
  const size_t smem = ((size_t)4 * UPC * p.S + (size_t)TBN * (p.S + 4) + (size_t)KG * TBN * 4 * UPC) * sizeof(float);
  SSASR_CHECK_CUDA(cudaFuncSetAttribute(rec_fwd_kernel<UPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  SSASR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rec_fwd_kernel<UPC>, 256, smem));
  dim3 grid(p.S / UPC, 2);
  SSASR_REQUIRE((int)(grid.x * 2) <= per_sm * sm_count(), "rec_fwd: %d CTAs cannot be co-resident (S=%d)", grid.x * 2, p.S);
  void* args[] = {&p};
  ProfScope ps(F_REC_FWD, st);
  SSASR_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)rec_fwd_kernel<UPC>, grid, dim3(256), args, smem, st));
  return 0;
}
template <int UPC>
static int launch_bwd(RecBwdParams& p, cudaStream_t st) {
  const size_t smem = ((size_t)UPC * 4 * p.S + (size_t)TBN * (p.S + 4) + (size_t)KG * TBN * UPC) * sizeof(float);
  SSASR_CHECK_CUDA(cudaFuncSetAttribute(rec_bwd_kernel<UPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  SSASR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rec_bwd_kernel<UPC>, 256, smem));
  dim3 grid(p.S / UPC, 2);
  SSASR_REQUIRE((int)(grid.x * 2) <= per_sm * sm_count(), "rec_bwd: %d CTAs cannot be co-resident (S=%d)", grid.x * 2, p.S);
  void* args[] = {&p};
  ProfScope ps(F_REC_BWD, st);
  SSASR_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)rec_bwd_kernel<UPC>, grid, dim3(256), args, smem, st));
  return 0;
}

}  // namespace ssasr

using namespace ssasr;

extern "C" {

// Packs one direction-pair of nn.LSTM parameters (asr.py:234-238, 403-404) into the kernel layout.
//   wih_p [8S,K], bias_p [8S], whh_p [2,4S,S], whhT_p [2,S,4S]
int ssasr_pack_blstm(const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
                     const float* w_ih_r, const float* w_hh_r, const float* b_ih_r, const float* b_hh_r, int S, int K,
                     float* wih_p, float* bias_p, float* whh_p, float* whhT_p, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_REQUIRE(S > 0 && S % 16 == 0, "pack_blstm: state size %d must be a positive multiple of 16", S);
  const float* wih[2] = {w_ih_f, w_ih_r};
  const float* whh[2] = {w_hh_f, w_hh_r};
  const float* bih[2] = {b_ih_f, b_ih_r};
  const float* bhh[2] = {b_hh_f, b_hh_r};
  ProfScope ps(F_PACK, st);
  for (int d = 0; d < 2; ++d) {
    pack_rows_kernel<<<256, 256, 0, st>>>(wih[d], wih_p + (size_t)d * 4 * S * K, S, K, K, 0);
    pack_rows_kernel<<<256, 256, 0, st>>>(whh[d], whh_p + (size_t)d * 4 * S * S, S, S, S, 0);
    pack_bias_kernel<<<(4 * S + 255) / 256, 256, 0, st>>>(bih[d], bhh[d], bias_p + (size_t)d * 4 * S, S);
    if (whhT_p)
      transpose_kernel<<<dim3((S + 31) / 32, (4 * S + 31) / 32), dim3(32, 8), 0, st>>>(whh_p + (size_t)d * 4 * S * S,
                                                                                       whhT_p + (size_t)d * 4 * S * S, 4 * S, S);
  }
  SSASR_LAUNCH_CHECK();
  return 0;
}

// Adds packed-layout gradients into the eight PyTorch-layout .grad tensors of one BLSTM.
int ssasr_unpack_blstm_grads(const float* dwih_p, const float* dbias_p, const float* dwhh_p, int S, int K, float* g_w_ih_f,
                             float* g_w_hh_f, float* g_b_ih_f, float* g_b_hh_f, float* g_w_ih_r, float* g_w_hh_r,
                             float* g_b_ih_r, float* g_b_hh_r, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  float* gwih[2] = {g_w_ih_f, g_w_ih_r};
  float* gwhh[2] = {g_w_hh_f, g_w_hh_r};
  float* gbih[2] = {g_b_ih_f, g_b_ih_r};
  float* gbhh[2] = {g_b_hh_f, g_b_hh_r};
  ProfScope ps(F_PACK, st);
  for (int d = 0; d < 2; ++d) {
    unpack_rows_add_kernel<<<256, 256, 0, st>>>(dwih_p + (size_t)d * 4 * S * K, gwih[d], S, K, K, 0);
    unpack_rows_add_kernel<<<256, 256, 0, st>>>(dwhh_p + (size_t)d * 4 * S * S, gwhh[d], S, S, S, 0);
    unpack_bias_add_kernel<<<(4 * S + 255) / 256, 256, 0, st>>>(dbias_p + (size_t)d * 4 * S, gbih[d], gbhh[d], S);
  }
  SSASR_LAUNCH_CHECK();
  return 0;
}

// bf16 training path: bias_p [8S] fp32, wih_bf [8S,Kp], whh_bf [8S,S], wihT_bf [K,8S], whhT_bf [2S,4S] (all bf16) in one launch
int ssasr_pack_blstm_bf16(const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
                          const float* w_ih_r, const float* w_hh_r, const float* b_ih_r, const float* b_hh_r, int S, int K, int Kp,
                          float* bias_p, void* wih_bf, void* whh_bf, void* wihT_bf, void* whhT_bf, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_REQUIRE(S > 0 && S % 8 == 0 && Kp >= K && Kp % 8 == 0, "pack_blstm_bf16: bad S=%d / K=%d / Kp=%d", S, K, Kp);
  PackBf16 a;
  a.wih[0] = w_ih_f; a.wih[1] = w_ih_r; a.whh[0] = w_hh_f; a.whh[1] = w_hh_r;
  a.bih[0] = b_ih_f; a.bih[1] = b_ih_r; a.bhh[0] = b_hh_f; a.bhh[1] = b_hh_r;
  a.S = S; a.K = K; a.Kp = Kp;
  a.nt_ih = ((Kp + 31) / 32) * (8 * S / 32);
  a.nt_hh = ((S + 31) / 32) * (8 * S / 32);
  a.bias_p = bias_p;
  a.wih_bf = (__nv_bfloat16*)wih_bf; a.whh_bf = (__nv_bfloat16*)whh_bf;
  a.wihT_bf = (__nv_bfloat16*)wihT_bf; a.whhT_bf = (__nv_bfloat16*)whhT_bf;
  ProfScope ps(F_PACK, st);
  pack_blstm_bf16_kernel<<<a.nt_ih + a.nt_hh + (8 * S + 255) / 256, 256, 0, st>>>(a);
  SSASR_LAUNCH_CHECK();
  return 0;
}

// Writes (does not add) the eight PyTorch-layout gradients of one BLSTM from the packed-layout ones, in one launch.
int ssasr_unpack_blstm_grads_set(const float* dwih_p, const float* dbias_p, const float* dwhh_p, int S, int K, float* g_w_ih_f,
                                 float* g_w_hh_f, float* g_b_ih_f, float* g_b_hh_f, float* g_w_ih_r, float* g_w_hh_r,
                                 float* g_b_ih_r, float* g_b_hh_r, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UnpackSet a;
  a.dwih_p = dwih_p; a.dbias_p = dbias_p; a.dwhh_p = dwhh_p;
  a.gwih[0] = g_w_ih_f; a.gwih[1] = g_w_ih_r; a.gwhh[0] = g_w_hh_f; a.gwhh[1] = g_w_hh_r;
  a.gbih[0] = g_b_ih_f; a.gbih[1] = g_b_ih_r; a.gbhh[0] = g_b_hh_f; a.gbhh[1] = g_b_hh_r;
  a.S = S; a.K = K;
  const size_t total = (size_t)8 * S * (K + S + 1);
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  ProfScope ps(F_PACK, st);
  unpack_blstm_set_kernel<<<blocks, 256, 0, st>>>(a);
  SSASR_LAUNCH_CHECK();
  return 0;
}

// nn.LSTMCell parameters (asr.py:277-283) -> wcat [4S, Kin+S] = [W_ih | W_hh] with interleaved rows,
// bcat [4S] = b_ih + b_hh (interleaved).
int ssasr_pack_lstmcell(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int S, int Kin,
                        float* wcat, float* bcat, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(F_PACK, st);
  pack_rows_kernel<<<256, 256, 0, st>>>(w_ih, wcat, S, Kin, Kin + S, 0);
  pack_rows_kernel<<<256, 256, 0, st>>>(w_hh, wcat, S, S, Kin + S, Kin);
  pack_bias_kernel<<<(4 * S + 255) / 256, 256, 0, st>>>(b_ih, b_hh, bcat, S);
  SSASR_LAUNCH_CHECK();
  return 0;
}
int ssasr_unpack_lstmcell_grads(const float* dwcat, const float* dbcat, int S, int Kin, float* g_w_ih, float* g_w_hh,
                                float* g_b_ih, float* g_b_hh, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(F_PACK, st);
  unpack_rows_add_kernel<<<256, 256, 0, st>>>(dwcat, g_w_ih, S, Kin, Kin + S, 0);
  unpack_rows_add_kernel<<<256, 256, 0, st>>>(dwcat, g_w_hh, S, S, Kin + S, Kin);
  unpack_bias_add_kernel<<<(4 * S + 255) / 256, 256, 0, st>>>(dbcat, g_b_ih, g_b_hh, S);
  SSASR_LAUNCH_CHECK();
  return 0;
}

// A "sequence" of ONE step from the zero state (what bs=1 decoding makes of encoder.blstm_4, asr.py:262: every frame is its own
// batch element): no recurrent product at all, h = o * tanh(i * g) per (row, direction, unit) straight from the pre-activations.
// Forward-only (exact decode path): writes hout only.
__global__ void blstm_single_step_kernel(const float* __restrict__ xp, float* __restrict__ hout, const int* __restrict__ lens, int S,
                                         long long n_batch, long long rs_batch) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_batch * 2 * S) return;
  const long long b = i / (2 * S);
  const int c = (int)(i % (2 * S));                 // dir * S + unit
  const long long row = b * rs_batch;
  float h = 0.f;
  if (!lens || lens[b] > 0) {
    const float4 g = __ldcs(reinterpret_cast<const float4*>(xp + row * 8 * S) + c);     // col = dir*4S + unit*4
    h = sigmoidf_acc(g.w) * tanhf(sigmoidf_acc(g.x) * tanhf(g.z));
  }
  hout[row * 2 * S + c] = h;
}

// One bidirectional LSTM layer, forward (fp32 path).
//   x    [n_rows, K]   input rows; row index = seq*xs_seq + batch*xs_batch  (same indexing for xp/hout/cbuf)
//   xp   [n_rows, 8S]  workspace; on return holds the gate activations (saved for backward)
//   hout [n_rows, 2S]  must be zero-filled for rows the sequence never visits; cbuf same shape
//   lens int32 [n_batch] (device) or null: per-batch-element sequence lengths (packed-sequence semantics)
//   bar  2 x uint32 scratch
int ssasr_blstm_fwd_f32(const float* x, int n_rows, int K, const float* wih_p, const float* bias_p, const float* whh_p,
                        int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch, const int* lens, float* xp,
                        float* hout, float* cbuf, unsigned* bar, float* tf32_ws, void* x3_ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_REQUIRE(S % 16 == 0, "blstm_fwd: S=%d must be a multiple of 16", S);
  int rc;
  if (tf32_ws && K % 8 == 0 && x3_gemm_bf16()) {
    // input projection as ONE bf16 tcgen05 GEMM over the tripled reduction axis [hi | hi | lo] x [hi | lo | hi] (6 bytes per
    // operand element of the 8 the workspace holds)
    __nv_bfloat16* xa = (__nv_bfloat16*)tf32_ws;
    __nv_bfloat16* wb = xa + (size_t)n_rows * 3 * K;
    rc = split3_bf16(st, x, K, n_rows, K, xa, 0);
    if (rc) return rc;
    rc = split3_bf16(st, wih_p, K, 8 * S, K, wb, 1);
    if (rc) return rc;
    rc = gemm_bf16_tc(st, n_rows, 8 * S, 3 * K, xa, 3 * K, 0, wb, 3 * K, 0, xp, 8 * S, bias_p, 0, 0, 1);
  } else if (tf32_ws && K % 4 == 0) {
    // input projection on tensor cores with the tf32 x 3 split (hi/lo pairs of x and W_ih in tf32_ws)
    float* xh = tf32_ws;
    float* xl = xh + (size_t)n_rows * K;
    float* wh = xl + (size_t)n_rows * K;
    float* wl = wh + (size_t)8 * S * K;
    rc = split_hi_lo(st, x, xh, xl, (size_t)n_rows * K);
    if (rc) return rc;
    rc = split_hi_lo(st, wih_p, wh, wl, (size_t)8 * S * K);
    if (rc) return rc;
    rc = gemm_tf32x3(st, n_rows, 8 * S, K, xh, xl, K, wh, wl, K, xp, 8 * S, bias_p, 0);
  } else {
    rc = gemm_f32(st, n_rows, 8 * S, K, x, K, 1, wih_p, K, 1, xp, 8 * S, bias_p, 0, 0);
  }
  if (rc) return rc;
  if (x3_ws && n_seq == 1) {
    ProfScope ps(F_POINTWISE, st);
    const long long n = (long long)n_batch * 2 * S;
    blstm_single_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(xp, hout, lens, S, n_batch, rs_batch);
    SSASR_LAUNCH_CHECK();
    return 0;
  }
  if (x3_ws && S % 64 == 0 && S <= 256) {
    // recurrence on tensor cores with the bf16 hi/lo split of h and W_hh (fp32-accurate: 3 MMAs per K step)
    __nv_bfloat16* wh = (__nv_bfloat16*)x3_ws;
    __nv_bfloat16* wl = wh + (size_t)8 * S * S;
    __nv_bfloat16* hh = wl + (size_t)8 * S * S;
    __nv_bfloat16* hl = hh + (size_t)n_rows * 2 * S;
    rc = split_bf16(st, whh_p, wh, wl, (size_t)8 * S * S);
    if (rc) return rc;
    // the quad-cluster kernel where it applies (S = 128 / 256, more than a few dependent steps): hout only
    rc = rec_q_fwd_x3(st, xp, wh, wl, hout, lens, S, n_seq, n_batch, rs_seq, rs_batch);
    if (rc >= 0) return rc;
    return rec_tc_fwd_x3(st, xp, wh, wl, hout, cbuf, hh, hl, lens, S, n_seq, n_batch, rs_seq, rs_batch, bar);
  }
  if (x3_ws && S == 512) {
    // the 16-CTA cluster kernel with split operands (x3_ws: 16 S^2 bf16); the SIMT recurrence below where it does not apply
    __nv_bfloat16* wh = (__nv_bfloat16*)x3_ws;
    __nv_bfloat16* wl = wh + (size_t)8 * S * S;
    rc = split_bf16(st, whh_p, wh, wl, (size_t)8 * S * S);
    if (rc) return rc;
    rc = rec_wide_fwd_x3(st, xp, wh, wl, hout, lens, S, n_seq, n_batch, rs_seq, rs_batch);
    if (rc >= 0) return rc;
  }
  SSASR_CHECK_CUDA(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), st));
  RecFwdParams p;
  p.xp = xp; p.whh = whh_p; p.hout = hout; p.cbuf = cbuf; p.lens = lens;
  p.S = S; p.n_seq = n_seq; p.n_batch = n_batch;
  p.xs_seq = rs_seq; p.xs_batch = rs_batch; p.hs_seq = rs_seq; p.hs_batch = rs_batch;
  p.bar = bar;
  p.UPC = pick_upc(S);
  return p.UPC == 4 ? launch_fwd<4>(p, st) : launch_fwd<8>(p, st);
}

// Backward of ssasr_blstm_fwd_f32.
//   act   [n_rows, 8S]  gate activations from forward; overwritten with dL/d(pre-activation gates)
//   dhout [n_rows, 2S]  dL/dhout
//   dx    [n_rows, K]   out (may be null)
//   dwih_p [8S,K], dbias_p [8S], dwhh_p [2,4S,S] out (overwritten)
//   dcstate [n_batch, 2S] scratch
//   shift_rows: row distance between a step and its forward-order predecessor (= rs_seq)
//   zero_period/zero_pos: rows r with (r % zero_period) == zero_pos have no forward predecessor inside the
//                         flat shifted pairing (time-major layers: period = rows per utterance); 0 = none.
int ssasr_blstm_bwd_f32(const float* x, int n_rows, int K, const float* wih_p, const float* whhT_p, int S, int n_seq,
                        int n_batch, long long rs_seq, long long rs_batch, const int* lens, float* act, const float* hout,
                        const float* cbuf, const float* dhout, float* dx, float* dwih_p, float* dbias_p, float* dwhh_p,
                        float* dcstate, unsigned* bar, int zero_period, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_CHECK_CUDA(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), st));
  RecBwdParams p;
  p.act = act; p.whhT = whhT_p; p.cbuf = cbuf; p.dhout = dhout; p.dcstate = dcstate; p.lens = lens;
  p.S = S; p.n_seq = n_seq; p.n_batch = n_batch;
  p.xs_seq = rs_seq; p.xs_batch = rs_batch; p.hs_seq = rs_seq; p.hs_batch = rs_batch;
  p.bar = bar;
  p.UPC = pick_upc(S);
  int rc = p.UPC == 4 ? launch_bwd<4>(p, st) : launch_bwd<8>(p, st);
  if (rc) return rc;
  const float* dg = act;
  if (dx) {
    rc = gemm_f32(st, n_rows, K, 8 * S, dg, 8 * S, 1, wih_p, K, 0, dx, K, nullptr, 0, 0);
    if (rc) return rc;
  }
  rc = gemm_f32(st, 8 * S, K, n_rows, dg, 8 * S, 0, x, K, 0, dwih_p, K, nullptr, 0, 0);
  if (rc) return rc;
  rc = colsum(st, dg, dbias_p, n_rows, 8 * S, 8 * S, 0);
  if (rc) return rc;
  // recurrent weights: dWhh[d] = sum over rows of dG_d[row]^T h_d[predecessor(row)]
  const long long sh = rs_seq;
  const int nred = n_rows - (int)sh;
  if (nred > 0) {
    // forward direction: predecessor = row - sh  -> A rows [sh, n_rows), B rows [0, n_rows - sh)
    rc = gemm_f32(st, 4 * S, S, nred, dg + (size_t)sh * 8 * S, 8 * S, 0, hout, 2 * S, 0, dwhh_p, S, nullptr, 0, 0, zero_period,
                  zero_period > 0 ? zero_period - (int)sh : 0);
    if (rc) return rc;
    // reverse direction: predecessor = row + sh -> A rows [0, n_rows - sh), B rows [sh, n_rows)
    rc = gemm_f32(st, 4 * S, S, nred, dg + 4 * S, 8 * S, 0, hout + (size_t)sh * 2 * S + S, 2 * S, 0,
                  dwhh_p + (size_t)4 * S * S, S, nullptr, 0, 0, zero_period, zero_period > 0 ? zero_period - (int)sh : 0);
    if (rc) return rc;
  } else {
    SSASR_CHECK_CUDA(cudaMemsetAsync(dwhh_p, 0, sizeof(float) * 8 * S * S, st));
  }
  return 0;
}


// ---- bf16 tensor-core variants: the batched-over-time gate GEMMs run on tcgen05 (gemm_tc.cu); the recurrence
// itself still uses the fp32 persistent kernel above.  Extra arguments are caller-provided bf16 workspaces.
//   wih_bf [8S, Kp] bf16 (Kp = K rounded up to 8), xb_ws [n_rows, Kp] bf16
int ssasr_blstm_fwd_bf16(const float* x, int n_rows, int K, int Kp, const void* wih_bf, const float* bias_p,
                         const float* whh_p, int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch,
                         const int* lens, void* xb_ws, float* xp, float* hout, float* cbuf, unsigned* bar,
                         const void* whh_bf, void* hb_ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_REQUIRE(S % 16 == 0 && Kp % 8 == 0 && Kp >= K, "blstm_fwd_bf16: bad S=%d / Kp=%d (K=%d)", S, Kp, K);
  int rc = 0;
  if (x) rc = cvt_bf16(st, x, K, xb_ws, Kp, n_rows, K);    // x == NULL: xb_ws already holds the bf16 input (the previous layer's bf16 h)
  if (rc) return rc;
  if (whh_bf && hb_ws && rec_tc_supported(S) && Kp <= rec_cl_fused_kp_max() && rec_cl_supported(S, n_batch, 0))
    // narrow layer input (layer 1: the fbank features): the projection runs inside the cluster recurrent kernel, the
    // [rows, 8S] pre-activation buffer is never materialised (xp only receives the saved activations)
    return rec_cl_fwd(st, xp, whh_bf, hout, cbuf, hb_ws, lens, S, n_seq, n_batch, rs_seq, rs_batch, xb_ws, Kp, wih_bf, bias_p);
  rc = gemm_bf16_tc(st, n_rows, 8 * S, K, xb_ws, Kp, 0, wih_bf, Kp, 0, xp, 8 * S, bias_p, 0);
  if (rc) return rc;
  if (whh_bf && hb_ws && rec_tc_supported(S))   // recurrence on tensor cores
    return rec_tc_fwd(st, xp, whh_bf, hout, cbuf, hb_ws, lens, S, n_seq, n_batch, rs_seq, rs_batch, bar);
  SSASR_CHECK_CUDA(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), st));
  RecFwdParams p;
  p.xp = xp; p.whh = whh_p; p.hout = hout; p.cbuf = cbuf; p.lens = lens;
  p.S = S; p.n_seq = n_seq; p.n_batch = n_batch;
  p.xs_seq = rs_seq; p.xs_batch = rs_batch; p.hs_seq = rs_seq; p.hs_batch = rs_batch;
  p.bar = bar;
  p.UPC = pick_upc(S);
  return p.UPC == 4 ? launch_fwd<4>(p, st) : launch_fwd<8>(p, st);
}

//   wihT_bf [K, 8S] bf16 (transposed packed input weights); Rp = n_rows rounded up to 8
//   workspaces: dgb_ws [n_rows, 8S] bf16, dgT_ws [8S, Rp] bf16, xT_ws [K, Rp] bf16, hT_ws [2S, Rp] bf16
int ssasr_blstm_bwd_bf16(const float* x, int n_rows, int K, const void* wihT_bf, const float* whhT_p, int S, int n_seq,
                         int n_batch, long long rs_seq, long long rs_batch, const int* lens, float* act, const float* hout,
                         const float* cbuf, const float* dhout, float* dx, float* dwih_p, float* dbias_p, float* dwhh_p,
                         float* dcstate, unsigned* bar, int zero_period, long long Rp, void* dgb_ws, void* dgT_ws, void* xT_ws,
                         void* hT_ws, const void* whhT_bf, const void* xb_saved, int Kp, void* hb_saved, void* stream,
                         void* wgrad_stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_REQUIRE(Rp % 8 == 0 && Rp >= n_rows, "blstm_bwd_bf16: bad Rp=%lld (n_rows=%d)", Rp, n_rows);
  int rc;
  const bool tc_rec = whhT_bf && dgb_ws && rec_tc_supported(S);
  if (tc_rec) {
    SSASR_CHECK_CUDA(cudaMemsetAsync(dbias_p, 0, sizeof(float) * 8 * S, st));     // bias gradient is reduced inside the kernel
    const int direct = (xb_saved && hb_saved && Kp % 8 == 0) ? 1 : 0;     // weight gradients read only the bf16 dG
    rc = rec_tc_bwd(st, act, whhT_bf, cbuf, dhout, dgb_ws, dcstate, lens, S, n_seq, n_batch, rs_seq, rs_batch, bar, dbias_p,
                    direct ? 0 : 1);
    if (rc) return rc;
  } else {
    SSASR_CHECK_CUDA(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), st));
    RecBwdParams p;
    p.act = act; p.whhT = whhT_p; p.cbuf = cbuf; p.dhout = dhout; p.dcstate = dcstate; p.lens = lens;
    p.S = S; p.n_seq = n_seq; p.n_batch = n_batch;
    p.xs_seq = rs_seq; p.xs_batch = rs_batch; p.hs_seq = rs_seq; p.hs_batch = rs_batch;
    p.bar = bar;
    p.UPC = pick_upc(S);
    rc = p.UPC == 4 ? launch_bwd<4>(p, st) : launch_bwd<8>(p, st);
    if (rc) return rc;
  }
  const float* dg = act;
  if (!tc_rec) {
    rc = colsum(st, dg, dbias_p, n_rows, 8 * S, 8 * S, 0);
    if (rc) return rc;
  }
  if (dx) {
    if (!tc_rec) {
      SSASR_REQUIRE(dgb_ws != nullptr, "blstm_bwd_bf16: dgb_ws required when dx is requested");
      rc = cvt_bf16(st, dg, 8 * S, dgb_ws, 8 * S, n_rows, 8 * S);
      if (rc) return rc;
    }
    rc = gemm_bf16_tc(st, n_rows, K, 8 * S, dgb_ws, 8 * S, 0, wihT_bf, 8 * S, 0, dx, K, nullptr, 0);
    if (rc) return rc;
  }
  const int sh = (int)rs_seq;
  // The weight gradients are not needed before the optimiser: on request they run on a second stream, ordered after
  // everything enqueued so far, so that they overlap the NEXT layer's recurrent kernel (which occupies 64 of the 148 SMs).
  // The caller joins that stream before it consumes dwih_p / dwhh_p (see functional.join_deferred).
  if (wgrad_stream && tc_rec && xb_saved && hb_saved && Kp % 8 == 0) {
    static cudaEvent_t evs[32];
    static int ev_i = -1;
    if (ev_i < 0) {
      for (int i = 0; i < 32; ++i) SSASR_CHECK_CUDA(cudaEventCreateWithFlags(&evs[i], cudaEventDisableTiming));
      ev_i = 0;
    }
    cudaEvent_t ev = evs[ev_i];
    ev_i = (ev_i + 1) & 31;
    SSASR_CHECK_CUDA(cudaEventRecord(ev, st));
    st = (cudaStream_t)wgrad_stream;
    SSASR_CHECK_CUDA(cudaStreamWaitEvent(st, ev, 0));
  }
  if (tc_rec && xb_saved && hb_saved && Kp % 8 == 0) {
    // weight gradients straight from the row-major bf16 buffers (MN-major tcgen05 operands): no transposed copies.
    // dgb_ws holds the complete bf16 dG (exchange buffer of the recurrent kernel), xb_saved / hb_saved the bf16 x and h
    // of the forward pass; h rows without a forward-order predecessor are zeroed per direction first.
    rc = gemm_bf16_tc_tn(st, 8 * S, K, n_rows, dgb_ws, 8 * S, 0, xb_saved, Kp, 0, dwih_p, K, 0);
    if (rc) return rc;
    rc = mask_rows_bf16(st, hb_saved, n_rows, 2 * S, zero_period, zero_period > 0 ? zero_period - 1 : 0, 0, S);
    if (rc) return rc;
    if (n_rows - sh > 0) {
      const __nv_bfloat16* dgb = (const __nv_bfloat16*)dgb_ws;
      const __nv_bfloat16* hb = (const __nv_bfloat16*)hb_saved;
      rc = gemm_bf16_tc_tn(st, 4 * S, S, n_rows - sh, dgb, 8 * S, sh, hb, 2 * S, 0, dwhh_p, S, 0);
      if (rc) return rc;
      rc = gemm_bf16_tc_tn(st, 4 * S, S, n_rows - sh, dgb + 4 * S, 8 * S, 0, hb + S, 2 * S, sh, dwhh_p + (size_t)4 * S * S, S, 0);
      if (rc) return rc;
    } else {
      SSASR_CHECK_CUDA(cudaMemsetAsync(dwhh_p, 0, sizeof(float) * 8 * S * S, st));
    }
    return 0;
  }
  rc = cvt_bf16_t(st, dg, 8 * S, dgT_ws, Rp, n_rows, 8 * S, 0, 0, 0, 0);
  if (rc) return rc;
  rc = cvt_bf16_t(st, x, K, xT_ws, Rp, n_rows, K, 0, 0, 0, 0);
  if (rc) return rc;
  rc = gemm_bf16_tc(st, 8 * S, K, n_rows, dgT_ws, Rp, 0, xT_ws, Rp, 0, dwih_p, K, nullptr, 0);
  if (rc) return rc;
  // h transposed AND shifted to its consumer row (fwd: h[r-sh] -> column r, rev: h[r+sh] -> column r), with the frame
  // that has no forward-order predecessor zeroed per direction; then dW_hh[d] = dG_d^T . hshift_d over all rows.
  SSASR_CHECK_CUDA(cudaMemsetAsync(hT_ws, 0, (size_t)2 * S * Rp * 2, st));
  rc = cvt_bf16_t(st, hout, 2 * S, hT_ws, Rp, n_rows, 2 * S, zero_period, zero_period > 0 ? zero_period - 1 : 0, 0, S, sh, -sh);
  if (rc) return rc;
  {
    const char* dgT = (const char*)dgT_ws;
    const char* hT = (const char*)hT_ws;
    rc = gemm_bf16_tc(st, 4 * S, S, n_rows, dgT, Rp, 0, hT, Rp, 0, dwhh_p, S, nullptr, 0);
    if (rc) return rc;
    rc = gemm_bf16_tc(st, 4 * S, S, n_rows, dgT + (size_t)4 * S * Rp * 2, Rp, 0, hT + (size_t)S * Rp * 2, Rp, 0,
                      dwhh_p + (size_t)4 * S * S, S, nullptr, 0);
    if (rc) return rc;
  }
  return 0;
}

}  // extern "C"

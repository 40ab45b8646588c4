// Beam search bookkeeping for ASR.beam_decode.  The reference configures a beam (conf/default.yaml:16-19
// `decode_beam_size`, trainer.py:552-554) but decodes greedily (trainer.py:590 "TODO: we are using simple decoding, not beam
// decoding"), so there is no reference behaviour to match; the semantics implemented here are stated in oracle/las_oracle.py
// `decode_beam` (the checker) and reduce to ASR.decode (asr.py:143-172) at beam size 1:
//   * a hypothesis' score is the sum of the per-step `final = log_softmax(asr) + lm_weight * log_softmax(lm)` (asr.py:153-156)
//     values of its tokens, EOS included;
//   * every step each live hypothesis is extended by all C tokens, a finished one (EOS emitted) stays a single candidate with its
//     score; the W best candidates survive, ties broken by the lower (parent, token) pair;
//   * W hypotheses of an utterance live in rows [n W, n W + W) of every state tensor.
// Kernels: candidate scoring + top-W selection (a warp per utterance), row gather by parent / by token (state re-ordering and the
// embedding lookup of the chosen tokens).  The step itself (attention, cells, character projection) runs on the per-step kernels
// of speller.cu.
#include <math.h>

#include "common.cuh"

namespace ssasr {
namespace {

constexpr int BM_MAXW = 16;      // beam width limit (candidates per lane: 2 W)

// logits / lm_logits [N W, C] (C <= 64), score_in / fin_in [N, W] -> score_out / fin_out / parent / token [N, W]
__global__ void __launch_bounds__(256) beam_select_kernel(const float* __restrict__ logits, const float* __restrict__ lm_logits,
                                                          float lm_weight, int N, int W, int C, int eos,
                                                          const float* __restrict__ score_in, const int* __restrict__ fin_in,
                                                          float* __restrict__ score_out, int* __restrict__ fin_out,
                                                          int* __restrict__ parent, int* __restrict__ token) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  const int c0 = lane, c1 = lane + 32;
  float v[2 * BM_MAXW];
#pragma unroll
  for (int w = 0; w < BM_MAXW; ++w) {
    v[2 * w] = -INFINITY;
    v[2 * w + 1] = -INFINITY;
    if (w < W) {
      const float s = score_in[n * W + w];
      if (fin_in[n * W + w]) {
        if (c0 == eos) v[2 * w] = s;
        if (c1 == eos) v[2 * w + 1] = s;
      } else {
        const float* lr = logits + (size_t)(n * W + w) * C;
        const float a0 = c0 < C ? lr[c0] : -INFINITY, a1 = c1 < C ? lr[c1] : -INFINITY;
        float mx = warp_max(fmaxf(a0, a1));
        float se = warp_sum((c0 < C ? expf(a0 - mx) : 0.f) + (c1 < C ? expf(a1 - mx) : 0.f));
        const float lse = mx + logf(se);
        float f0 = a0 - lse, f1 = a1 - lse;
        if (lm_logits) {
          const float* mr = lm_logits + (size_t)(n * W + w) * C;
          const float b0 = c0 < C ? mr[c0] : -INFINITY, b1 = c1 < C ? mr[c1] : -INFINITY;
          mx = warp_max(fmaxf(b0, b1));
          se = warp_sum((c0 < C ? expf(b0 - mx) : 0.f) + (c1 < C ? expf(b1 - mx) : 0.f));
          const float lsm = mx + logf(se);
          f0 += lm_weight * (b0 - lsm);
          f1 += lm_weight * (b1 - lsm);
        }
        if (c0 < C) v[2 * w] = s + f0;
        if (c1 < C) v[2 * w + 1] = s + f1;
      }
    }
  }
  // W rounds of "best remaining candidate": value first, then the lower flat index w * C + c
  for (int k = 0; k < W; ++k) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int w = 0; w < BM_MAXW; ++w) {
      if (w < W) {
        const int i0 = w * C + c0, i1 = w * C + c1;
        if (c0 < C && (v[2 * w] > bv || (v[2 * w] == bv && i0 < bi))) { bv = v[2 * w]; bi = i0; }
        if (c1 < C && (v[2 * w + 1] > bv || (v[2 * w + 1] == bv && i1 < bi))) { bv = v[2 * w + 1]; bi = i1; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    const int pw = bi / C, pc = bi - pw * C;
    // the winner leaves the candidate set (its owner is lane pc % 32, slot 2 pw + pc / 32)
#pragma unroll
    for (int w = 0; w < BM_MAXW; ++w) {
      if (w == pw) {
        if (pc == c0) v[2 * w] = -INFINITY;
        if (pc == c1) v[2 * w + 1] = -INFINITY;
      }
    }
    if (lane == 0) {
      score_out[n * W + k] = bv;
      parent[n * W + k] = pw;
      token[n * W + k] = pc;
      fin_out[n * W + k] = (fin_in[n * W + pw] || pc == eos) ? 1 : 0;
    }
  }
}

// dst row i = src row (group ? (i / group) * group : 0) + idx[i]   (rows of `row_words` 32-bit words)
__global__ void gather_rows_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const int* __restrict__ idx,
                                   long long n_rows, int row_words, int group) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_rows * row_words) return;
  const long long r = i / row_words;
  const int c = (int)(i - r * row_words);
  const long long s = (group ? (r / group) * group : 0) + idx[r];
  dst[i] = src[s * row_words + c];
}

}  // namespace
}  // namespace ssasr

using namespace ssasr;

extern "C" {

int ssasr_beam_select(const float* logits, const float* lm_logits, float lm_weight, int N, int W, int C, int eos, const float* score_in,
                      const int* fin_in, float* score_out, int* fin_out, int* parent, int* token, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_REQUIRE(W >= 1 && W <= BM_MAXW && C >= 1 && C <= 64 && eos >= 0 && eos < C,
                "beam_select: beam width 1..%d, at most 64 classes, eos inside them (W=%d C=%d eos=%d)", BM_MAXW, W, C, eos);
  if (N <= 0) return 0;
  ProfScope ps(F_POINTWISE, st);
  beam_select_kernel<<<(N + 7) / 8, 256, 0, st>>>(logits, lm_logits, lm_weight, N, W, C, eos, score_in, fin_in, score_out, fin_out,
                                                  parent, token);
  SSASR_LAUNCH_CHECK();
  return 0;
}

int ssasr_gather_rows(const void* src, void* dst, const int* idx, long long n_rows, long long row_bytes, int group, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_REQUIRE(row_bytes > 0 && row_bytes % 4 == 0, "gather_rows: rows of whole 32-bit words (row_bytes=%lld)", row_bytes);
  if (n_rows <= 0) return 0;
  const long long n = n_rows * (row_bytes / 4);
  ProfScope ps(F_POINTWISE, st);
  gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, idx, n_rows, (int)(row_bytes / 4),
                                                                  group);
  SSASR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

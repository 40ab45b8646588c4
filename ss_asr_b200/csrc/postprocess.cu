// Validation metrics of the reference (src/postprocess.py:7-50, called from ASRTrainer.valid, src/trainer.py:493-494;
// SURVEY.md §8f row f4) on the device: per utterance, the argmax token of every decoded step, the character-accuracy
// counts of calc_acc and the word-level Levenshtein distance of calc_err.  The reference copies the [B, U, C] prediction to
// the host and runs Python double loops over it; here only 4 int32 per utterance travel back.  Integer work, bit-exact.
//
// One CTA per utterance:
//   1. argmax over the C classes of each step, one warp per step (np.argmax semantics: first maximum, a NaN wins).
//   2. calc_acc: walk (prediction, label) pairs until the first label 0 (postprocess.py:21-26): total = that length,
//      correct = matches inside it.
//   3. calc_err: Mapper.translate (ASRDataset.py:240-252) on both sequences = cut after the first token 1 (trim_eos,
//      postprocess.py:68-75), drop the SOS / EOS characters, then str.split(' '): words are the runs between space
//      tokens (empty words kept).  Equal words get equal ids (first occurrence), and the Levenshtein table over the two
//      id sequences is filled one anti-diagonal per barrier.
#include "common.cuh"

namespace ssasr {
namespace {

struct ArgBest {
  float v;
  int i;
};
// np.argmax order: a NaN beats every number, otherwise the larger value; ties (and NaN vs NaN) go to the lower index
__device__ __forceinline__ bool arg_better(float v, int i, float bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn != bn) return vn;
  if (!vn && v != bv) return v > bv;
  return i < bi;
}

// splits toks[0..n) (already trimmed / filtered) into words at `space_id`; returns the number of words (>= 1)
__device__ int split_words(const int* toks, int n, int space_id, int* wstart, int* wlen) {
  int w = 0, s = 0;
  for (int k = 0; k < n; ++k)
    if (toks[k] == space_id) {
      wstart[w] = s;
      wlen[w] = k - s;
      ++w;
      s = k + 1;
    }
  wstart[w] = s;
  wlen[w] = n - s;
  return w + 1;
}

__global__ void __launch_bounds__(128) calc_acc_err_kernel(const float* __restrict__ predict, long long p_bstride, long long p_ustride,
                                                           int U, int C, const long long* __restrict__ label, long long l_bstride,
                                                           int L, int sos_id, int eos_id, int space_id, int* __restrict__ stats,
                                                           int* __restrict__ tokens_out) {
  extern __shared__ int sm[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* tokp = sm;                  // [U]   argmax tokens
  int* fp = tokp + U;              // [U]   prediction after translate()
  int* fl = fp + U;                // [L]   label after translate()
  int* ws_p = fl + L;              // [U+1] word starts / lengths
  int* wl_p = ws_p + (U + 1);
  int* ws_l = wl_p + (U + 1);      // [L+1]
  int* wl_l = ws_l + (L + 1);
  int* id_p = wl_l + (L + 1);      // [U+1] canonical word ids
  int* id_l = id_p + (U + 1);      // [L+1]
  int* diag = id_l + (L + 1);      // 3 x [U+2]
  __shared__ int s_n[6];           // n_valid, correct, len fp, len fl, words p, words l

  const float* pb = predict + (long long)b * p_bstride;
  const long long* lb = label + (long long)b * l_bstride;
  for (int u = warp; u < U; u += 4) {
    const float* row = pb + (long long)u * p_ustride;
    float bv = 0.f;
    int bi = 0x7fffffff;
    for (int k = lane; k < C; k += 32) {
      const float v = row[k];
      if (bi == 0x7fffffff || arg_better(v, k, bv, bi)) { bv = v; bi = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || arg_better(ov, oi, bv, bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      tokp[u] = bi;
      if (tokens_out) tokens_out[(long long)b * U + u] = bi;
    }
  }
  const int n_pair = U < L ? U : L;     // zip(p, l) stops at the shorter sequence
  if (tid == 0) { s_n[0] = n_pair; s_n[1] = 0; }
  __syncthreads();
  for (int j = tid; j < n_pair; j += blockDim.x)
    if (lb[j] == 0) atomicMin(&s_n[0], j);
  __syncthreads();
  const int n_valid = s_n[0];
  int correct = 0;
  for (int j = tid; j < n_valid; j += blockDim.x) correct += ((long long)tokp[j] == lb[j]);
  if (correct) atomicAdd(&s_n[1], correct);

  // translate(): one thread per sequence (a few hundred tokens at most)
  if (tid == 0 || tid == 32) {
    const bool is_p = tid == 0;
    const int n = is_p ? U : L;
    int* dst = is_p ? fp : fl;
    int m = 0;
    for (int k = 0; k < n; ++k) {
      const long long t = is_p ? (long long)tokp[k] : lb[k];
      if (t != sos_id && t != eos_id) dst[m++] = (int)t;
      if (t == 1) break;                 // trim_eos keeps the EOS and stops (postprocess.py:73-74)
    }
    s_n[is_p ? 2 : 3] = m;
    s_n[is_p ? 4 : 5] = is_p ? split_words(fp, m, space_id, ws_p, wl_p) : split_words(fl, m, space_id, ws_l, wl_l);
  }
  __syncthreads();
  const int Wp = s_n[4], Wl = s_n[5], W = Wp + Wl;
  // canonical ids: index (in the concatenated word list) of the first identical word
  for (int w = tid; w < W; w += blockDim.x) {
    const int* tw = w < Wp ? fp + ws_p[w] : fl + ws_l[w - Wp];
    const int lw = w < Wp ? wl_p[w] : wl_l[w - Wp];
    int id = w;
    for (int v = 0; v < w; ++v) {
      const int lv = v < Wp ? wl_p[v] : wl_l[v - Wp];
      if (lv != lw) continue;
      const int* tv = v < Wp ? fp + ws_p[v] : fl + ws_l[v - Wp];
      bool eq = true;
      for (int k = 0; k < lw && eq; ++k) eq = tv[k] == tw[k];
      if (eq) { id = v; break; }
    }
    if (w < Wp) id_p[w] = id; else id_l[w - Wp] = id;
  }
  __syncthreads();
  // Levenshtein over (id_p[0..Wp), id_l[0..Wl)): D[i][j], anti-diagonal k = i + j held as d[i]
  int* d2 = diag;
  int* d1 = diag + (U + 2);
  int* d0 = diag + 2 * (U + 2);
  for (int k = 0; k <= Wp + Wl; ++k) {
    const int i_lo = k > Wl ? k - Wl : 0, i_hi = k < Wp ? k : Wp;
    for (int i = i_lo + tid; i <= i_hi; i += blockDim.x) {
      const int j = k - i;
      int v;
      if (i == 0) v = j;
      else if (j == 0) v = i;
      else {
        const int del = d1[i - 1] + 1, ins = d1[i] + 1, sub = d2[i - 1] + (id_p[i - 1] != id_l[j - 1]);
        v = del < ins ? del : ins;
        v = sub < v ? sub : v;
      }
      d0[i] = v;
    }
    __syncthreads();
    int* t = d2; d2 = d1; d1 = d0; d0 = t;
  }
  if (tid == 0) {
    stats[4 * b + 0] = s_n[1];
    stats[4 * b + 1] = n_valid;
    stats[4 * b + 2] = d1[Wp];
    stats[4 * b + 3] = Wl;
  }
}

}  // namespace
}  // namespace ssasr

extern "C" {
// predict: [B, U, C] fp32 device (element strides p_bstride / p_ustride, classes contiguous); label: [B, L] int64 device (row
// stride l_bstride); stats: int32 [B, 4] = {correct, total (calc_acc), word edit distance, label words (calc_err)};
// tokens_out: int32 [B, U] argmax tokens or NULL.
int ssasr_calc_acc_err(const float* predict, long long p_bstride, long long p_ustride, int B, int U, int C, const long long* label,
                       long long l_bstride, int L, int sos_id, int eos_id, int space_id, int* stats, int* tokens_out, void* stream) {
  using namespace ssasr;
  cudaStream_t st = (cudaStream_t)stream;
  if (B <= 0) return 0;
  SSASR_REQUIRE(predict && label && stats && U > 0 && L > 0 && C > 0, "calc_acc_err: bad arguments");
  SSASR_REQUIRE(U <= 4096 && L <= 4096, "calc_acc_err: at most 4096 decoded steps / label tokens (got U=%d L=%d)", U, L);
  const size_t smem = sizeof(int) * ((size_t)2 * U + L + 3 * (size_t)(U + 1) + 3 * (size_t)(L + 1) + 3 * (size_t)(U + 2));
  if (smem > 48 * 1024)
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(calc_acc_err_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(F_POINTWISE, st);
  calc_acc_err_kernel<<<B, 128, smem, st>>>(predict, p_bstride, p_ustride, U, C, label, l_bstride, L, sos_id, eos_id, space_id, stats,
                                            tokens_out);
  SSASR_LAUNCH_CHECK();
  return 0;
}
}

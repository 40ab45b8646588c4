// Shared helpers for the ss_asr_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace ssasr {

void set_error(const char* fmt, ...);

#define SSASR_CHECK_CUDA(expr)                                                        \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ssasr::set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)_e, \
                       cudaGetErrorString(_e), #expr);                                \
      return -2;                                                                      \
    }                                                                                 \
  } while (0)

#define SSASR_REQUIRE(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      ssasr::set_error(__VA_ARGS__);    \
      return -1;                        \
    }                                   \
  } while (0)

#define SSASR_LAUNCH_CHECK() SSASR_CHECK_CUDA(cudaGetLastError())

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reductions through a 32-float shared scratch (blockDim.x multiple of 32, <= 1024).
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : 0.0f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

int sm_count();

// Launch accounting / optional per-family CUDA-event timing (bench.py reads it through ssasr_profile_*).
enum Family { F_GEMM_F32 = 0, F_REC_FWD, F_REC_BWD, F_ATTN_FWD, F_ATTN_BWD, F_POINTWISE, F_CE, F_FBANK, F_PACK,
              F_GEMM_TC, F_REC_TC_FWD, F_REC_TC_BWD, F_OPTIM, F_SPELL_FWD, F_SPELL_BWD, F_COUNT };
struct ProfScope {
  int fam; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr;
  ProfScope(int family, cudaStream_t stream);
  ~ProfScope();
};

// internal fp32 GEMM launcher (gemm_f32.cu): C[M,N] = (accumulate ? C : 0) + op(A) * op(B) (+ bias[N]) (+tanh)
//   a_kmajor: element A(m,k) at A[m*lda + k] (1) or A[k*lda + m] (0)
//   b_kmajor: element B(k,n) at B[n*ldb + k] (1, the "weights [N,K]" form) or B[k*ldb + n] (0)
int gemm_f32(cudaStream_t st, int M, int N, int K, const float* A, int lda, int a_kmajor, const float* B, int ldb,
             int b_kmajor, float* C, int ldc, const float* bias, int accumulate, int act_tanh, int zero_period = 0,
             int zero_pos = 0);
//   zero_period > 0: reduction index k with (k % zero_period) == zero_pos contributes nothing (used to drop
//   the first/last frame of every utterance from the recurrent-weight gradient).

// bf16 tensor-core GEMM (gemm_tc.cu): C[M,N] fp32 (+)= A[M,K] x B[N,K]^T (+bias); A,B bf16 with K contiguous;
// a_koff/b_koff shift the reduction window inside each operand's rows.
int gemm_bf16_tc(cudaStream_t st, int M, int N, int K, const void* A, long long lda, int a_koff, const void* B, long long ldb,
                 int b_koff, float* C, int ldc, const float* bias, int accumulate, int act_tanh = 0, int splits = 1);
//   splits > 1 (small products on 128 x 32 tiles only): split-K, every split ADDS its partial to C with red.add -- the caller
//   pre-zeroes C; with two splits the sum is order-independent, i.e. still deterministic
// C[M,N] fp32 (+)= A^T B with A stored [K,M], B stored [K,N] (row-major bf16, 16-byte row pitches); koffs shift rows
int gemm_bf16_tc_tn(cudaStream_t st, int M, int N, int K, const void* A, long long lda, int a_koff, const void* B, long long ldb,
                    int b_koff, float* C, int ldc, int accumulate);
// fp32-accurate tensor-core GEMM: tf32 x 3 split (gemm_tc.cu)
int split_bf16(cudaStream_t st, const float* x, void* hi, void* lo, size_t n);
int split_hi_lo_2d(cudaStream_t st, const float* x, long long ld, int rows, int cols, float* hi, float* lo);
int split3_bf16(cudaStream_t st, const float* x, long long ld, long long rows, int K, void* out, int second);
int x3_gemm_bf16();
int split_hi_lo(cudaStream_t st, const float* x, float* hi, float* lo, size_t n);
int gemm_tf32x3(cudaStream_t st, int M, int N, int K, const float* A, const float* A_lo, long long lda, const float* B,
                const float* B_lo, long long ldb, float* C, int ldc, const float* bias, int act_tanh);
int mask_rows_bf16(cudaStream_t st, void* x, long long rows, int cols, int period, int pos_lo, int pos_hi, int split);
int cvt_bf16(cudaStream_t st, const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols);
int cvt_bf16_t(cudaStream_t st, const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols,
               int mask_period, int mask_pos_lo, int mask_pos_hi, int mask_split, int shift_lo = 0, int shift_hi = 0);

// tensor-core recurrent kernels (rec_tc.cu)
int rec_tc_supported(int S);
int rec_tc_fwd(cudaStream_t st, float* xp, const void* whh_bf, float* hout, float* cbuf, void* hb, const int* lens, int S,
               int n_seq, int n_batch, long long rs_seq, long long rs_batch, unsigned* bar);
// exact split-operand forward recurrence on the quad clusters (rec_cl.cu); -1: geometry not covered, use rec_tc_fwd_x3
int rec_q_fwd_x3(cudaStream_t st, float* xp, const void* whh_hi, const void* whh_lo, float* hout, const int* lens, int S, int n_seq,
                 int n_batch, long long rs_seq, long long rs_batch);
int rec_wide_fwd_x3(cudaStream_t st, float* xp, const void* whh_hi, const void* whh_lo, float* hout, const int* lens, int S, int n_seq,
                    int n_batch, long long rs_seq, long long rs_batch);     // the same at S = 512 (rec_wide.cu)
int rec_tc_fwd_x3(cudaStream_t st, float* xp, const void* whh_hi, const void* whh_lo, float* hout, float* cbuf, void* hb_hi,
                  void* hb_lo, const int* lens, int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch, unsigned* bar);
int rec_tc_bwd(cudaStream_t st, float* act, const void* whhT_bf, const float* cbuf, const float* dhout, void* dgb, float* dcstate,
               const int* lens, int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch, unsigned* bar,
               float* dbias /*[8S] pre-zeroed, or null*/, int need_dg32 = 1 /*0: only the bf16 dG (dgb) is consumed*/);

// cluster recurrent kernels (rec_cl.cu): multicast-TMA exchange inside one thread-block cluster per (direction, batch tile)
int rec_cl_supported(int S, int n_batch, int backward);
int rec_cl_fwd(cudaStream_t st, float* xp, const void* whh_bf, float* hout, float* cbuf, void* hb, const int* lens, int S,
               int n_seq, int n_batch, long long rs_seq, long long rs_batch, const void* x_bf = nullptr, int Kp = 0,
               const void* wih_bf = nullptr, const float* bias = nullptr);
int rec_cl_fused_kp_max();
int rec_cl_bwd(cudaStream_t st, float* act, const void* whhT_bf, const float* cbuf, const float* dhout, void* dgb, const int* lens,
               int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch, float* dbias);


// 16-CTA cluster recurrent kernels for S = 512 (rec_wide.cu): the long-utterance configuration
int rec_wide_supported(int S, int n_batch, int backward);
int rec_ks_supported(int S, int n_batch);     // K-split backward at S = 256 / 128 (rec_wide_bwd runs it)
int rec_dsmem_enabled();                      // cluster exchange by DSMEM bulk copies (SSASR_REC_DSMEM=0: through the L2 ring)
int rec_wide_fwd(cudaStream_t st, float* xp, const void* whh_bf, float* hout, float* cbuf, void* hb, const int* lens, int S,
                 int n_seq, int n_batch, long long rs_seq, long long rs_batch);
int rec_wide_bwd(cudaStream_t st, float* act, const void* whhT_bf, const float* cbuf, const float* dhout, void* dgb, const int* lens,
                 int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch, float* dbias);

// cluster-persistent decoder-step kernels (spell_cl.cu): steps [t0, t1) of the attend-and-spell loop (attention + layer-1 cell)
// in ONE launch.  Strides (`*_ldb` per utterance, `*_ldt` per step) are in elements.
struct SpellClFwdArgs {
  int B, U, Tp, t0, t1;
  const void* w1cat_bf; int X1, K1;       // [4Sd, X1] bf16 (rows = unit*4 + gate); the recurrent block starts at column K1
  const void* phi_bf;                     // [M, Sd] bf16
  const void* P_bf;                       // [B*Tp, 4Sd] bf16: W_ctx enc
  const void* psi_bf;                     // [B*Tp, M] bf16: tanh(psi(enc))
  const float* gemb;                      // [C, 4Sd]: W_emb emb + b
  const int* tok; long long tok_ld;       // [B, U] input token of every step
  const int* enc_lens;
  float* act1; long long act1_ldb, act1_ldt;     // out: gate activations (i, f, g, o per unit)
  float* c1; long long c1_ldb, c1_ldt;           // out (and in at t0 - 1)
  float* h1; long long h1_ldb, h1_ldt;           // out (and in at t0 - 1)
  void* h1b; long long h1b_ldb, h1b_ldt;         // out, optional: bf16 copy of h1
  float* q; long long q_ldb, q_ldt;              // out [.., M]
  float* alpha; long long al_ldb, al_ldt;        // out [.., Tp]
  // plain recurrence mode (xpre != NULL; the layer-2 cell chain): gates = W_hh h(t-1) + xpre[b, t]; the recurrent block of
  // `w1cat_bf` (row pitch X1, starting at column K1) is then W_hh of that cell; the attention arguments are unused
  const float* xpre; long long xpre_ldb, xpre_ldt;
  float* h2nd; long long h2nd_ldb, h2nd_ldt; int h2nd_toff;   // optional second copy of h(t), written at step t + h2nd_toff (< U)
};
// backward chain of the same loop over ALL steps (t1 - 1 down to 0).  P_bf == NULL: plain recurrence (the layer-2 chain).
struct SpellClBwdArgs {
  int B, U, Tp, t0, t1;
  const void* wcat_bf; int X, Kcol;       // [4Sd, X] bf16 forward weights; the recurrent block starts at column Kcol
  const void* phi_bf;                     // [M, Sd] bf16
  const void* P_bf;                       // [B*Tp, 4Sd] bf16
  const void* psi_bf;                     // [B*Tp, M] bf16
  const int* enc_lens;
  float* act; long long act_ldb, act_ldt;           // in: gate activations; out: gate gradients (in place)
  const float* c; long long c_ldb, c_ldt;
  const float* dh_in; long long dh_ldb, dh_ldt;     // gradient arriving on h(t) from outside the chain
  void* dgb; long long dgb_ldb, dgb_ldt;            // out: bf16 copy of the gate gradients
  const float* alpha; long long al_ldb, al_ldt;
  const float* q; long long q_ldb, q_ldt;
  float* de; long long de_ldb, de_ldt;              // out: dL/d(energy) [.., Tp]
  float* dqpre; long long dq_ldb, dq_ldt;           // out: dL/d(query pre-activation) [.., M]
};
int spell_cl_bwd(cudaStream_t st, const SpellClBwdArgs& a);
int spell_cl_supported(int B, int Tp, int E, int Sd, int M);
int spell_cl_fwd(cudaStream_t st, const SpellClFwdArgs& a);
// xin1[b, t, :] = [emb(tok[b, t]) ; sum_j alpha[b, t, j] enc[b, j, :] ; h1[b, t - 1, :]] for all steps (the operand of the
// weight-gradient products; the loop kernels never materialise the context)
int spell_fill_xin1(cudaStream_t st, int B, int U, int Tp, int E, int Sd, const float* alpha, const float* enc, const int* lens,
                    const float* emb_w, const int* tok, const float* h1, long long h1_ldb, long long h1_ldt, float* xin1);

}  // namespace ssasr

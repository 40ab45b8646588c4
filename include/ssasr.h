/* ssasr.h -- C ABI of libssasr.so: the B200 (sm_100a) kernels behind the Listen-Attend-Spell hot path of
 * cadia-lvl/ss_asr.  The reference has no native/FFI layer (it is pure Python on torch); each
 * entry point below names the reference Python interface it replaces (paths relative to /root/reference).
 *
 * Conventions: plain pointers and sizes only; every pointer is a DEVICE pointer unless marked HOST; buffers are
 * caller-owned (no hidden allocation except the cached fbank tables); `stream` is a cudaStream_t passed as
 * void*; every function returns 0 on success, <0 on error (ssasr_last_error() gives a thread-local message);
 * nothing throws across the boundary.  Distinct streams may run concurrently; one stream is serial.
 *
 * Gate layout used by all LSTM buffers ("interleaved"): column = dir*4S + unit*4 + gate, gate order i,f,g,o.
 */
#ifndef SSASR_H_
#define SSASR_H_
#ifdef __cplusplus
extern "C" {
#endif

const char* ssasr_last_error(void);
/* Bumped whenever a signature or an argument struct of this header changes; the ctypes binding (ss_asr_b200/_lib.py) refuses
 * to bind a library whose version differs from the one it was written for (a stale .so would silently mis-read structs). */
#define SSASR_ABI_VERSION 4
int ssasr_abi_version(void);

/* ---- log-mel filterbank: preprocess.py:187-208 log_fbank(y, sample_rate) (librosa 0.6.3 melspectrogram) ---- */
long long ssasr_fbank_num_frames(long long n_samples, int sample_rate);
int ssasr_fbank(const float* audio, const long long* offsets /*[n_utt+1]*/, int n_utt, int sample_rate, int n_mels,
                float* out /*[sum frames, n_mels]*/, const long long* out_offsets /*[n_utt+1] rows*/, int max_frames,
                void* stream);

/* ---- fp32 GEMM (the nn.Linear / LSTM gate products of asr.py on the exact path) ---- */
int ssasr_gemm_f32(int M, int N, int K, const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor,
                   float* C, int ldc, const float* bias, int accumulate, int act_tanh, void* stream);

/* ---- parameter packing: nn.LSTM / nn.LSTMCell tensors of asr.py:234-238,277-283,403-404 -> kernel layout ---- */
int ssasr_pack_blstm(const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
                     const float* w_ih_r, const float* w_hh_r, const float* b_ih_r, const float* b_hh_r, int S, int K,
                     float* wih_p /*[8S,K]*/, float* bias_p /*[8S]*/, float* whh_p /*[2,4S,S]*/,
                     float* whhT_p /*[2,S,4S] or NULL*/, void* stream);
/* bf16 training path: the same packing straight to the bf16 operands of the tensor-core kernels, one launch:
 * bias_p [8S] fp32, wih_bf [8S,Kp] (Kp = K rounded up to 8, zero-padded), whh_bf [8S,S], wihT_bf [K,8S], whhT_bf [2S,4S] */
int ssasr_pack_blstm_bf16(const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
                          const float* w_ih_r, const float* w_hh_r, const float* b_ih_r, const float* b_hh_r, int S, int K, int Kp,
                          float* bias_p, void* wih_bf, void* whh_bf, void* wihT_bf, void* whhT_bf, void* stream);
/* like ssasr_unpack_blstm_grads but WRITES the eight gradients (no zero-fill needed), one launch */
int ssasr_unpack_blstm_grads_set(const float* dwih_p, const float* dbias_p, const float* dwhh_p, int S, int K, float* g_w_ih_f,
                                 float* g_w_hh_f, float* g_b_ih_f, float* g_b_hh_f, float* g_w_ih_r, float* g_w_hh_r,
                                 float* g_b_ih_r, float* g_b_hh_r, void* stream);
int ssasr_unpack_blstm_grads(const float* dwih_p, const float* dbias_p, const float* dwhh_p, int S, int K,
                             float* g_w_ih_f, float* g_w_hh_f, float* g_b_ih_f, float* g_b_hh_f, float* g_w_ih_r,
                             float* g_w_hh_r, float* g_b_ih_r, float* g_b_hh_r, void* stream);
int ssasr_pack_lstmcell(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int S, int Kin,
                        float* wcat /*[4S,Kin+S]*/, float* bcat /*[4S]*/, void* stream);
int ssasr_unpack_lstmcell_grads(const float* dwcat, const float* dbcat, int S, int Kin, float* g_w_ih, float* g_w_hh,
                                float* g_b_ih, float* g_b_hh, void* stream);

/* ---- one bidirectional LSTM layer: pBLSTM.forward asr.py:406-427 (packed-sequence semantics through `lens`)
 *      and encoder.blstm_4 asr.py:237-238,262 (seq-first quirk: n_seq = batch size, n_batch = frames, lens NULL).
 *      Row index of x/xp/hout/cbuf = seq*rs_seq + batch*rs_batch.  The pair-concat down-sampling
 *      (asr.py:429-450) is a view of hout and costs nothing. ---- */
int ssasr_blstm_fwd_f32(const float* x, int n_rows, int K, const float* wih_p, const float* bias_p, const float* whh_p,
                        int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch, const int* lens,
                        float* xp /*[n_rows,8S] out: gate activations*/, float* hout /*[n_rows,2S], pre-zeroed*/,
                        float* cbuf /*[n_rows,2S]*/, unsigned* bar /*2 words scratch*/,
                        float* tf32_ws /*NULL: fp32 SIMT input projection; else 2*(n_rows+8S)*K floats scratch: the
                                         projection runs on tensor cores with the tf32 x 3 split (K % 4 == 0)*/,
                        void* x3_ws /*NULL: fp32 SIMT recurrence; else (16*S*S + 4*n_rows*S) bf16 scratch (S == 512: 16*S*S) and `bar` of
                                      4096 words: recurrence on tensor cores with the bf16 hi/lo x 3 split (S % 64 == 0, S <= 256,
                                      or S == 512 on the 16-CTA clusters)*/,
                        void* stream);
int ssasr_blstm_bwd_f32(const float* x, int n_rows, int K, const float* wih_p, const float* whhT_p, int S, int n_seq,
                        int n_batch, long long rs_seq, long long rs_batch, const int* lens,
                        float* act /*in: activations, out: gate grads*/, const float* hout, const float* cbuf,
                        const float* dhout, float* dx /*[n_rows,K] or NULL*/, float* dwih_p, float* dbias_p,
                        float* dwhh_p, float* dcstate /*[n_batch,2S] scratch*/, unsigned* bar, int zero_period,
                        void* stream);

/* ---- bf16 tensor-core path (tcgen05 + TMA) for the batched-over-time gate GEMMs of training:
 *      C[M,N] fp32 (+)= A[M,K] x B[N,K]^T (+bias); A, B bf16 with K contiguous, 16-byte aligned base and row pitch;
 *      a_koff / b_koff shift the reduction window inside the operands' rows.  Replaces the GEMMs cuDNN/ATen run
 *      inside nn.LSTM forward/backward (asr.py:414,262; trainer.py:437). ---- */
int ssasr_gemm_bf16_tc(int M, int N, int K, const void* A, long long lda, int a_koff, const void* B, long long ldb,
                       int b_koff, float* C, int ldc, const float* bias, int accumulate, void* stream);
/* fp32-ACCURATE tensor-core GEMM for the exact path: tf32 x 3 split (hi*hi + hi*lo + lo*hi), dense fp32 operands with
   K % 4 == 0; A_ws / B_ws are caller scratch of TWICE the operands' sizes (hi | lo).  C = A B^T (+bias)(+tanh). */
int ssasr_gemm_tf32x3(int M, int N, int K, const float* A, float* A_ws, long long lda, const float* B, float* B_ws,
                      long long ldb, float* C, int ldc, const float* bias, int act_tanh, void* stream);
/* C[M,N] fp32 (+)= A^T B; A stored [K,M], B stored [K,N] row-major bf16 (the weight-gradient form dW = dG^T X) */
int ssasr_gemm_bf16_tc_tn(int M, int N, int K, const void* A, long long lda, int a_koff, const void* B, long long ldb,
                          int b_koff, float* C, int ldc, int accumulate, void* stream);
int ssasr_cvt_bf16(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols, void* stream);
int ssasr_cvt_bf16_t(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols,
                     int mask_period, int mask_pos_lo, int mask_pos_hi, int mask_split, void* stream);
/* same contracts as ssasr_blstm_fwd_f32 / _bwd_f32 plus caller-provided bf16 operands and workspaces */
int ssasr_blstm_fwd_bf16(const float* x /*NULL: xb_ws already holds the bf16 input*/, int n_rows, int K, int Kp, const void* wih_bf /*[8S,Kp]*/, const float* bias_p,
                         const float* whh_p, int S, int n_seq, int n_batch, long long rs_seq, long long rs_batch,
                         const int* lens, void* xb_ws /*[n_rows,Kp]*/, float* xp, float* hout, float* cbuf,
                         unsigned* bar /*512 words*/, const void* whh_bf /*[8S,S] bf16 or NULL*/,
                         void* hb_ws /*[n_rows,2S] bf16 or NULL; both set => tensor-core recurrence*/, void* stream);
int ssasr_blstm_bwd_bf16(const float* x, int n_rows, int K, const void* wihT_bf /*[K,8S]*/, const float* whhT_p, int S,
                         int n_seq, int n_batch, long long rs_seq, long long rs_batch, const int* lens, float* act,
                         const float* hout, const float* cbuf, const float* dhout, float* dx, float* dwih_p,
                         float* dbias_p, float* dwhh_p, float* dcstate, unsigned* bar, int zero_period, long long Rp,
                         void* dgb_ws /*[n_rows,8S]*/, void* dgT_ws /*[8S,Rp]*/, void* xT_ws /*[K,Rp]*/,
                         void* hT_ws /*[2S,Rp]*/, const void* whhT_bf /*[2S,4S] bf16 or NULL => fp32 recurrence*/,
                         const void* xb_saved /*[n_rows,Kp] bf16 x of the forward pass, or NULL*/, int Kp,
                         void* hb_saved /*[n_rows,2S] bf16 h of the forward pass (masked in place), or NULL;
                                          with both set the weight gradients use the MN-major GEMM and dgT/xT/hT_ws
                                          may be NULL*/,
                         void* stream,
                         void* wgrad_stream /*NULL, or a second stream for the weight-gradient GEMMs: the caller joins it
                                              before reading dwih_p / dwhh_p*/);

/* ---- attend-and-spell loop: Attention.forward asr.py:343-392 + Speller.forward asr.py:314-326 + the decode loop
 *      of ASR.forward asr.py:65-110 (teacher forcing / greedy / sampled), all U steps on the device ---- */
typedef struct {
  int B, Tp, E, Sd, M, C, U;
  const float *phi_w, *psi_w, *psi_b, *w1cat, *b1, *w2cat, *b2, *emb_w, *wc, *bc;
  const float* enc;         /* [B,Tp,E] listener output */
  const int* enc_lens;      /* [B] */
  int* tok_in;              /* [B,U] input token per step (col 0 = <sos>=0; teacher columns pre-filled) */
  const int* step_mode;     /* HOST [U] or NULL: token after step t: 0 teacher, 1 argmax, 2 sample, 3 argmax with LM */
  unsigned long long seed;
  float *psi, *xin1, *xin2, *act1, *act2, *c1, *c2, *h2all, *q, *alpha, *logits; /* outputs + saved state */
  const void *w1cat_bf, *w2cat_bf; /* bf16 copies of w1cat/w2cat, or NULL (fp32 path) */
  void* ws_bf;                     /* bf16 scratch [B, X1 + X2], or NULL */
  void* enc_bf;                    /* bf16 scratch [(B*Tp + M) * E], or NULL */
  /* character LM for step mode 3 (ASR.decode with lm_weight != 0, asr.py:153-162 + charlm.py:46-57):
     GRU / output weights TRANSPOSED to [in,out], biases, and the two [B,H] hidden states (updated in place) */
  int lm_H;
  float lm_weight;
  const float *lm_emb, *lm_w1i, *lm_w1h, *lm_b1i, *lm_b1h, *lm_w2i, *lm_w2h, *lm_b2i, *lm_b2h, *lm_wo, *lm_bo;
  float *lm_h1, *lm_h2;
  float* x3_ws; /* NULL, or 2*B*(X1+X2) + 8*Sd*(X1+X2) floats: forward-only gate GEMMs on tensor cores with split operands
                   (bf16 [hi|hi|lo] x [hi|lo|hi] rows by default; tf32 x 3 pairs with SSASR_X3_GEMM=tf32) */
  int skip_final_logits; /* 1: do not recompute the [B,U,C] logits after the loop (greedy decoding only needs tok_in) */
  int dual_stream;       /* bf16 mode: ws_bf holds B*X1 + U*B*X2 elements (a layer-2 input block per step) and the layer-2
                            chain (asr.py:322-324) runs on an internal second stream, joined into `stream` before returning */
  int stop_token, stop_check_every; /* greedy decoding: every `stop_check_every` steps (0 = never) the call reads back whether EVERY
                                       utterance has emitted `stop_token` (asr.py:161-162) and stops early if so */
  int* stop_scratch;                /* device int */
  int* steps_run;                   /* HOST int out or NULL: decoding steps executed */
  /* bf16 mode, optional: workspace of ssasr_speller_cl_ws_bytes() bytes.  When given (and the dimensions are covered) the
     teacher-forced runs of the loop execute in the cluster-persistent step kernel (csrc/spell_cl.cu): attention query, energies,
     masked softmax, context and the layer-1 LSTMCell of asr.py:79-103 for a whole run of steps in ONE launch, decoder state and
     query weights resident on chip.  Keep it alive until the backward pass has run. */
  void* cl_ws;
  long long cl_ws_bytes;
} ssasr_speller_fwd_args;
int ssasr_speller_fwd_f32(const ssasr_speller_fwd_args* a, void* stream);
/* bytes of `cl_ws` for these dimensions; 0 = not covered by the cluster-persistent kernels (S_d = 256, mlp = 128, T' <= 64) */
long long ssasr_speller_cl_ws_bytes(int B, int Tp, int E, int Sd, int M, int C, int U);
long long ssasr_speller_cl_bwd_ws_bytes(int B, int Tp, int E, int Sd, int M, int U);

typedef struct {
  int B, Tp, E, Sd, M, C, U;
  const float *phi_w, *psi_w, *w1cat, *w2cat, *wc;
  const float* enc;
  const int* enc_lens;
  const int* tok_in;
  const float *psi, *xin1, *xin2, *c1, *c2, *h2all, *q, *alpha;
  float *act1, *act2;
  const float* dlogits;     /* [B,U,C] */
  float *d_phi_w, *d_psi_w, *d_psi_b, *d_w1cat, *d_b1, *d_w2cat, *d_b2, *d_emb_w, *d_wc, *d_bc, *denc;
  float *dh2all, *dxin1 /*[B,U,X1]*/, *dxin2, *dc1s, *dc2s, *dh1att, *dpsi, *dqpre, *de_all /*[B,U,Tp]*/;   /* scratch */
  const void *w1catT_bf, *w2catT_bf; /* bf16 [X1,4Sd] / [X2,4Sd] transposed weights, or NULL (fp32 path) */
  void *wsA, *wsB;                   /* bf16 scratch: >= max(4Sd*BUp, M*BTp) and >= max(X1*BUp, E*BTp + E*M) elements */
  long long BUp, BTp;                /* B*U and B*Tp rounded up to multiples of 8 */
  int dual_stream;                   /* bf16 mode: dxin2 holds U*B*X2 floats; layer-2 chain on an internal second stream */
  void* wgrad_stream;                /* optional, with dual_stream: the products only the optimiser consumes (all weight / bias /
                                        embedding gradients) go to this stream, ordered after the loop and NOT joined into `stream` */
  /* cluster-persistent path (csrc/spell_cl.cu), all optional: the forward call's `cl_ws`, the forward bf16 weights
     [4Sd, X1] / [4Sd, X2] and a scratch of ssasr_speller_cl_bwd_ws_bytes() bytes: both cell chains and the attention backward
     (asr.py:79-103 under loss.backward(), trainer.py:437) then run as ONE launch each over all steps */
  const void* cl_ws;
  const void *w1cat_bf, *w2cat_bf;
  void* cl_ws_bwd;
  long long cl_ws_bwd_bytes;
} ssasr_speller_bwd_args;
int ssasr_speller_bwd_f32(const ssasr_speller_bwd_args* a, void* stream);

/* ---- single-step module API: Attention.forward (asr.py:343-392) and Speller.forward / nn.LSTMCell (asr.py:314-326) as
 *      called one decoding step at a time by the reference's other trainers (text_autoencoder.py:52-94) ---- */
int ssasr_attn_step_fwd(int B, int Tp, int E, int Sd, int M, const float* h, const float* phi_w, const float* psi,
                        const float* enc, const int* enc_lens, float* xrow /*[B,Sd+E+Sd] scratch; ctx = cols Sd..Sd+E*/,
                        float* q /*[B,M]*/, float* alpha /*[B,Tp]*/, void* stream);
int ssasr_attn_step_bwd(int B, int Tp, int E, int Sd, int M, const float* dctx, const float* dalpha /*or NULL*/,
                        const float* alpha, const float* q, const float* phi_w, const float* psi, const float* enc,
                        const int* enc_lens, float* de /*[B,Tp]*/, float* dqpre /*[B,M]*/, float* dh /*[B,Sd]*/,
                        float* denc /*[B,Tp,E]*/, float* dpsi /*[B,Tp,M]*/, void* stream);
int ssasr_lstmcell_fwd(int B, int S, float* gates /*[B,4S] interleaved: pre-activations in, activations out*/,
                       const float* c_prev /*or NULL*/, float* c_out, float* h_out, void* stream);
int ssasr_lstmcell_bwd(int B, int S, float* act /*activations in, gate gradients out*/, const float* c, const float* c_prev,
                       const float* dh, float* dc /*in: dL/dc_out, out: dL/dc_prev*/, void* stream);
int ssasr_dtanh_mul(float* d, const float* y, long long n, void* stream);
int ssasr_colsum(const float* src, float* out, int R, int C, int ld, int accumulate, void* stream);

/* ---- ASR loss: trainer.py:394-395,426-434 (CrossEntropyLoss(ignore_index=0,'none'), per-utterance length
 *      normalisation, batch mean), fused with its gradient ---- */
int ssasr_ce_loss_f32(const float* logits, const long long* y, int B, int U, int C, int L, float* loss_b /*[B]*/,
                      float* loss_out /*[1]*/, float* dlogits /*[B,U,C] or NULL*/, float grad_scale, void* stream);

/* ---- Solver.step (trainer.py:131-148: clip_grad_norm_ + NaN skip + optimiser step) with the Adadelta of trainer.py:401-403,
 *      for all parameter tensors in two launches and without the host sync of the NaN test ---- */
typedef struct {
  float* p;          /* parameter, updated in place */
  float* g;          /* gradient (rescaled in place only if write_clipped_grads) */
  float* sq;         /* Adadelta square_avg */
  float* acc;        /* Adadelta acc_delta */
  long long n;       /* elements */
} ssasr_optim_tensor;
long long ssasr_adadelta_scratch_floats(const ssasr_optim_tensor* tensors /*HOST array*/, int n_tensors);
int ssasr_adadelta_clip_step(const ssasr_optim_tensor* tensors /*HOST array of device pointers*/, int n_tensors, float lr, float rho,
                             float eps, float max_norm /*<= 0: no clipping*/, float* scratch, float* norm_out /*device [2]: norm, applied*/,
                             int write_clipped_grads, void* stream);

/* ---- prepare_x (ASRDataset.py:297-316): fp64/fp32 batch -> fp32 + per-utterance count of the frames
 *      whose feature sum is non-zero, on the device (the reference copies the whole batch back to the host to count) ---- */
int ssasr_prepare_x(const void* src /*[B,T,F] device*/, int src_is_f64, long long B, long long T, int F, float* dst /*[B,T,F] or NULL*/,
                    int* lens /*int32 [B], overwritten*/, void* stream);

/* ---- validation metrics (postprocess.py:7-50 calc_acc / calc_err, called at trainer.py:493-494):
 *      per-utterance argmax, character-accuracy counts and word-level Levenshtein distance on the device; the reference
 *      copies the whole [B,U,C] prediction to the host and loops in Python.  Mapper.translate (ASRDataset.py:240-252) and
 *      trim_eos (postprocess.py:68-75) define the token -> word-list rule. ---- */
int ssasr_calc_acc_err(const float* predict /*[B,U,C] device*/, long long p_bstride, long long p_ustride, int B, int U, int C,
                       const long long* label /*int64 [B,L] device*/, long long l_bstride, int L, int sos_id, int eos_id,
                       int space_id /*-1: no word separator*/, int* stats /*int32 [B,4]: correct, total, edit distance, label words*/,
                       int* tokens_out /*int32 [B,U] or NULL*/, void* stream);

/* ---- launch accounting and per-family CUDA-event timing (used by bench.py; no reference counterpart) ---- */
int ssasr_num_families(void);
const char* ssasr_family_name(int i);
/* ---- beam search bookkeeping (semantics: oracle/las_oracle.py decode_beam, = asr.py:143-172 at W = 1) ----
   logits / lm_logits(or NULL) [N*W, C] (C <= 64), score_in / fin_in [N, W] -> the W best candidates per utterance:
   score_out, fin_out, parent (hypothesis index 0..W-1 inside the utterance), token [N, W].  W <= 16. */
int ssasr_beam_select(const float* logits, const float* lm_logits, float lm_weight, int N, int W, int C, int eos, const float* score_in,
                      const int* fin_in, float* score_out, int* fin_out, int* parent, int* token, void* stream);
/* dst row i = src row ((group ? (i / group) * group : 0) + idx[i]); rows of row_bytes (multiple of 4) */
int ssasr_gather_rows(const void* src, void* dst, const int* idx, long long n_rows, long long row_bytes, int group, void* stream);
/* strided host -> device copy (cudaMemcpy2DAsync): `height` rows of `width` bytes; asynchronous for pinned host memory */
int ssasr_memcpy2d_h2d(void* dst, long long dpitch, const void* src, long long spitch, long long width, long long height, void* stream);
long long ssasr_launch_count(void);
void ssasr_launch_count_reset(void);
void ssasr_profile_enable(int enable);
int ssasr_profile_read(double* ms /*[num_families]*/, long long* launches /*[num_families]*/);
void ssasr_rec_tc_set_debug(long long* dev_buf /*[n_seq][12] clock64 stamps of CTA 0, or NULL*/);
void ssasr_rec_cl_set_debug(long long* dev_buf /*[n_seq][12], cluster recurrent kernels (rec_cl.cu)*/);
int ssasr_rec_cl_capacity(int S, int backward); /* co-resident (direction, tile) clusters of the cluster recurrence; 0 = unavailable */
void ssasr_rec_cl_enable(int on);              /* 0: counter-barrier recurrent kernels everywhere (A/B comparison) */
void ssasr_rec_wide_set_debug(long long* dev_buf /*[n_seq][12], K-split backward kernels (rec_wide.cu)*/);
void ssasr_rec_set_dsmem(int mode);            /* cluster exchange of the recurrent kernels: 1 DSMEM bulk copies (default), 0 through the L2 ring, 2 DSMEM for the 16-CTA clusters too (A/B comparison) */
void ssasr_spell_cl_set_debug(long long* dev_buf /*[steps][8], cluster decoder-step kernels (spell_cl.cu)*/);
void ssasr_spell_cl_set_debug_bwd(long long* dev_buf /*[steps][8], backward kernel*/);
void ssasr_spell_cl_set_debug_mode(int plain_recurrence /*1: stamp the layer-2 (plain recurrence) launches instead*/);
void ssasr_rec_q_set_rows(int rows);           /* batch rows per tile of the quad-cluster recurrence: 32 / 16; 0 = 8-CTA kernels */

#ifdef __cplusplus
}
#endif
#endif /* SSASR_H_ */

// Batched log-mel filterbank frontend: framing (reflect-padded, centred), periodic Hann window, real FFT,
// power spectrum, Slaney mel projection and log, in ONE kernel, output already [frames, n_mels].
//
// Reference semantics: log_fbank, /root/reference/src/preprocess.py:187-208, i.e.
// librosa(0.6.3).feature.melspectrogram(y, sr, n_mels=N_DIMS, n_fft=ws, hop_length=st) -> log(. + eps) -> T.
//
// One CTA handles groups of 16 consecutive frames of one utterance: the audio span they share is read once
// (hop < window, every sample is used by 2.5 frames); at 16 kHz (25 ms = 400 samples) two real frames are packed
// into one 400-point complex FFT (20 x 20, two register-resident 20-point DFTs per thread with one exchange through
// shared memory) followed by the real-pair split; other window sizes (the reference's default 22.05 kHz gives
// 551 = 19*29) take a direct-DFT kernel.  The mel matrix is applied in its sparse (two triangles per bin) form.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include <mutex>
#include <vector>

namespace ssasr {

constexpr int FPB = 16;            // frames per CTA
constexpr int NT = 256;
constexpr int TWP = 10;            // twiddles per column of the 20 x 20 register FFT (fbank400q)

struct FbankTables {
  int sr = 0, n_mels = 0, ws = 0, st = 0, nbins = 0;
  float* window = nullptr;     // [ws]
  float2* tw = nullptr;        // fast path: [200] W_200^m then [201] W_400^k ; generic: [ws] W_ws^m
  float2* tw400 = nullptr;     // 16 kHz register-FFT path: [400] W_400^e
  float* win_half = nullptr;   // fbank400q: [400] 0.5 * window (the 1/2 of the frame separation folded into the input)
  float2* twc = nullptr;       // fbank400q: [20][TWP] W_400^(c k1), k1 = 1..10; row c = the thread's column
  int2* mel_meta = nullptr;    // fbank400q: [n_mels] {first bin | padded count << 16, offset into mel_w4}
  float* mel_w4 = nullptr;     // fbank400q: zero-padded filter weights (see get_tables)
  int* mel_start = nullptr;    // [n_mels]
  int* mel_cnt = nullptr;      // [n_mels]
  int* mel_off = nullptr;      // [n_mels]
  float* mel_w = nullptr;      // packed non-zero weights
  int max_cnt = 0;
};

static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static std::mutex g_tab_mu;
static std::vector<FbankTables> g_tabs;

static int get_tables(int sr, int n_mels, FbankTables* out) {
  std::lock_guard<std::mutex> lk(g_tab_mu);
  for (auto& t : g_tabs)
    if (t.sr == sr && t.n_mels == n_mels) { *out = t; return 0; }
  FbankTables t;
  t.sr = sr; t.n_mels = n_mels;
  t.ws = (int)(sr * 0.001 * 25);
  t.st = (int)(sr * 0.001 * 10);
  t.nbins = 1 + t.ws / 2;
  const int ws = t.ws;
  std::vector<float> win(ws);
  for (int n = 0; n < ws; ++n) win[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / ws));
  std::vector<float2> tw;
  if (ws == 400) {
    for (int m = 0; m < 200; ++m) tw.push_back(make_float2((float)cos(2.0 * M_PI * m / 200), (float)-sin(2.0 * M_PI * m / 200)));
    for (int k = 0; k <= 200; ++k) tw.push_back(make_float2((float)cos(2.0 * M_PI * k / 400), (float)-sin(2.0 * M_PI * k / 400)));
  } else {
    for (int m = 0; m < ws; ++m) tw.push_back(make_float2((float)cos(2.0 * M_PI * m / ws), (float)-sin(2.0 * M_PI * m / ws)));
  }
  // Slaney mel basis (librosa.filters.mel, htk=False, norm=1), fp64 then rounded to fp32
  std::vector<double> mel_f(n_mels + 2);
  const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(sr / 2.0);
  for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
  std::vector<int> start(n_mels), cnt(n_mels), off(n_mels);
  std::vector<float> wts;
  for (int i = 0; i < n_mels; ++i) {
    int first = -1, last = -1;
    std::vector<double> row(t.nbins);
    for (int b = 0; b < t.nbins; ++b) {
      const double f = (sr / 2.0) * b / (t.nbins - 1);
      const double lower = (f - mel_f[i]) / (mel_f[i + 1] - mel_f[i]);
      const double upper = (mel_f[i + 2] - f) / (mel_f[i + 2] - mel_f[i + 1]);
      double w = lower < upper ? lower : upper;
      if (w < 0) w = 0;
      w *= 2.0 / (mel_f[i + 2] - mel_f[i]);
      row[b] = w;
      if (w > 0) { if (first < 0) first = b; last = b; }
    }
    start[i] = first < 0 ? 0 : first;
    cnt[i] = first < 0 ? 0 : last - first + 1;
    off[i] = (int)wts.size();
    for (int b = 0; b < cnt[i]; ++b) wts.push_back((float)row[start[i] + b]);
    if (cnt[i] > t.max_cnt) t.max_cnt = cnt[i];
  }
  if (wts.empty()) wts.push_back(0.f);
  if (ws == 400) {
    std::vector<float2> t4(400);
    for (int e = 0; e < 400; ++e) t4[e] = make_float2((float)cos(2.0 * M_PI * e / 400), (float)-sin(2.0 * M_PI * e / 400));
    SSASR_CHECK_CUDA(cudaMalloc(&t.tw400, sizeof(float2) * 400));
    SSASR_CHECK_CUDA(cudaMemcpy(t.tw400, t4.data(), sizeof(float2) * 400, cudaMemcpyHostToDevice));
    std::vector<float> wh(400);
    for (int n = 0; n < 400; ++n) wh[n] = 0.5f * win[n];
    std::vector<float2> tc(20 * TWP);                      // row c: W_400^(c k1) for k1 = 1..9, then k1 = 10
    for (int c = 0; c < 20; ++c)
      for (int j = 0; j < TWP; ++j) tc[c * TWP + j] = t4[(c * (j + 1)) % 400];
    // mel filters for the (mel, frame pair) work items of fbank400q: the two filters of a half-warp (mels 2m, 2m+1) start
    // on bins of different parity (one leading zero weight where needed: conflict-free 64-bit loads of the interleaved
    // power spectrum) and every filter is padded with zero weights to a multiple of 4 bins (no remainder loop)
    std::vector<int> qs(start), qc(cnt);
    for (int i = 0; i + 1 < n_mels; i += 2)
      if (((qs[i] ^ qs[i + 1]) & 1) == 0) {
        const int v = qs[i + 1] > 0 ? i + 1 : (qs[i] > 0 ? i : -1);
        if (v >= 0) { qs[v] -= 1; qc[v] += 1; }
      }
    std::vector<float> qw;
    std::vector<int2> meta(n_mels);
    for (int i = 0; i < n_mels; ++i) {
      const int lead = start[i] - qs[i];
      const int c4 = (qc[i] + 3) & ~3;
      meta[i] = make_int2(qs[i] | (c4 << 16), (int)qw.size());
      for (int b = 0; b < c4; ++b) {
        const int src = b - lead;
        qw.push_back(src >= 0 && src < cnt[i] ? wts[off[i] + src] : 0.f);
      }
    }
    if (qw.empty()) qw.push_back(0.f);
    SSASR_CHECK_CUDA(cudaMalloc(&t.win_half, sizeof(float) * 400));
    SSASR_CHECK_CUDA(cudaMalloc(&t.twc, sizeof(float2) * tc.size()));
    SSASR_CHECK_CUDA(cudaMalloc(&t.mel_meta, sizeof(int2) * n_mels));
    SSASR_CHECK_CUDA(cudaMalloc(&t.mel_w4, sizeof(float) * qw.size()));
    SSASR_CHECK_CUDA(cudaMemcpy(t.win_half, wh.data(), sizeof(float) * 400, cudaMemcpyHostToDevice));
    SSASR_CHECK_CUDA(cudaMemcpy(t.twc, tc.data(), sizeof(float2) * tc.size(), cudaMemcpyHostToDevice));
    SSASR_CHECK_CUDA(cudaMemcpy(t.mel_meta, meta.data(), sizeof(int2) * n_mels, cudaMemcpyHostToDevice));
    SSASR_CHECK_CUDA(cudaMemcpy(t.mel_w4, qw.data(), sizeof(float) * qw.size(), cudaMemcpyHostToDevice));
  }
  SSASR_CHECK_CUDA(cudaMalloc(&t.window, sizeof(float) * ws));
  SSASR_CHECK_CUDA(cudaMalloc(&t.tw, sizeof(float2) * tw.size()));
  SSASR_CHECK_CUDA(cudaMalloc(&t.mel_start, sizeof(int) * n_mels));
  SSASR_CHECK_CUDA(cudaMalloc(&t.mel_cnt, sizeof(int) * n_mels));
  SSASR_CHECK_CUDA(cudaMalloc(&t.mel_off, sizeof(int) * n_mels));
  SSASR_CHECK_CUDA(cudaMalloc(&t.mel_w, sizeof(float) * wts.size()));
  SSASR_CHECK_CUDA(cudaMemcpy(t.window, win.data(), sizeof(float) * ws, cudaMemcpyHostToDevice));
  SSASR_CHECK_CUDA(cudaMemcpy(t.tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
  SSASR_CHECK_CUDA(cudaMemcpy(t.mel_start, start.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
  SSASR_CHECK_CUDA(cudaMemcpy(t.mel_cnt, cnt.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
  SSASR_CHECK_CUDA(cudaMemcpy(t.mel_off, off.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
  SSASR_CHECK_CUDA(cudaMemcpy(t.mel_w, wts.data(), sizeof(float) * wts.size(), cudaMemcpyHostToDevice));
  g_tabs.push_back(t);
  *out = t;
  return 0;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// ---- register-resident 20-point DFT (4 x 5 Cooley-Tukey, compile-time twiddles) --------------------------------
__device__ __forceinline__ float2 w20(int e) {     // W_20^e = exp(-2 pi i e / 20); e is a compile-time constant after unrolling
  switch (e) {
    case 0: return make_float2(1.000000000e+00f, -0.000000000e+00f);
    case 1: return make_float2(9.510565163e-01f, -3.090169944e-01f);
    case 2: return make_float2(8.090169944e-01f, -5.877852523e-01f);
    case 3: return make_float2(5.877852523e-01f, -8.090169944e-01f);
    case 4: return make_float2(3.090169944e-01f, -9.510565163e-01f);
    case 5: return make_float2(6.123233996e-17f, -1.000000000e+00f);
    case 6: return make_float2(-3.090169944e-01f, -9.510565163e-01f);
    case 7: return make_float2(-5.877852523e-01f, -8.090169944e-01f);
    case 8: return make_float2(-8.090169944e-01f, -5.877852523e-01f);
    case 9: return make_float2(-9.510565163e-01f, -3.090169944e-01f);
    case 10: return make_float2(-1.000000000e+00f, -1.224646799e-16f);
    case 11: return make_float2(-9.510565163e-01f, 3.090169944e-01f);
    case 12: return make_float2(-8.090169944e-01f, 5.877852523e-01f);
  }
  return make_float2(1.f, 0.f);
}
__device__ __forceinline__ void dft4r(float2& a, float2& b, float2& c, float2& d) {     // in place, forward
  const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
  a = cadd(t0, t2);
  c = csub(t0, t2);
  b = make_float2(t1.x + t3.y, t1.y - t3.x);
  d = make_float2(t1.x - t3.y, t1.y + t3.x);
}
__device__ __forceinline__ void dft5r(float2& v0, float2& v1, float2& v2, float2& v3, float2& v4) {
  const float c1 = 0.30901699437494745f, c2 = -0.8090169943749475f, s1 = 0.9510565162951535f, s2 = 0.5877852522924731f;
  const float2 t1 = cadd(v1, v4), t2 = cadd(v2, v3), t3 = csub(v1, v4), t4 = csub(v2, v3);
  const float2 a1 = make_float2(v0.x + c1 * t1.x + c2 * t2.x, v0.y + c1 * t1.y + c2 * t2.y);
  const float2 a2 = make_float2(v0.x + c2 * t1.x + c1 * t2.x, v0.y + c2 * t1.y + c1 * t2.y);
  const float2 b1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
  const float2 b2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
  v0 = make_float2(v0.x + t1.x + t2.x, v0.y + t1.y + t2.y);
  v1 = make_float2(a1.x + b1.y, a1.y - b1.x);
  v4 = make_float2(a1.x - b1.y, a1.y + b1.x);
  v2 = make_float2(a2.x + b2.y, a2.y - b2.x);
  v3 = make_float2(a2.x - b2.y, a2.y + b2.x);
}
// x[n], n = 5a + b  ->  X[k], k = c + 4d, returned in place as x[c + 4d]
__device__ __forceinline__ void dft20(float2 (&x)[20]) {
  float2 t[5][4];
#pragma unroll
  for (int b = 0; b < 5; ++b) {
    float2 r0 = x[b], r1 = x[5 + b], r2 = x[10 + b], r3 = x[15 + b];
    dft4r(r0, r1, r2, r3);
    t[b][0] = r0;
    t[b][1] = b ? cmul(r1, w20(b)) : r1;
    t[b][2] = b ? cmul(r2, w20(2 * b)) : r2;
    t[b][3] = b ? cmul(r3, w20(3 * b)) : r3;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    dft5r(t[0][c], t[1][c], t[2][c], t[3][c], t[4][c]);
#pragma unroll
    for (int d = 0; d < 5; ++d) x[c + 4 * d] = t[d][c];
  }
}

struct FbankParams {
  const float* audio; const long long* offsets; int n_utt;
  float* out; const long long* out_offsets;
  int ws, st, nbins, n_mels;
  const float* window; const float2* tw;
  const int *mel_start, *mel_cnt, *mel_off; const float* mel_w;
  const float* win_half; const float2* twc; const int2* mel_meta; const float* mel_w4;      // fbank400q
};

// ---- 16 kHz / 25 ms fast path, register FFT: two real frames are packed into one 400-point complex FFT
// (z = x_a + i x_b, |X_a[k]|^2 = |Z[k] + conj Z[400-k]|^2 / 4, |X_b[k]|^2 = |Z[k] - conj Z[400-k]|^2 / 4), and
// 400 = 20 x 20: 20 threads per frame pair each run one register-resident 20-point DFT per pass, with one exchange
// through shared memory in between.  8 pairs (16 frames) per CTA of 160 threads.
constexpr int FPAIRS = 8;
constexpr int NTP = FPAIRS * 20;
constexpr int EXP = 21;            // row pitch (complex) of the 20 x 20 exchange tile: conflict-free both ways
__global__ void __launch_bounds__(NTP, 4) fbank400p_kernel(FbankParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int ws = 400, st = 160, half = 200, NF = 2 * FPAIRS, PB = 202;
  float* span = smem;                                             // [(NF-1)*160 + 400]
  float* win = span + (NF - 1) * st + ws;                         // [400]
  float2* tw = reinterpret_cast<float2*>(win + ws);               // [400] W_400^e
  float2* ex = tw + 400;                                          // [FPAIRS][20*EXP]
  float* P = reinterpret_cast<float*>(ex + FPAIRS * 20 * EXP);    // [NF][PB]
  const int u = blockIdx.y;
  const long long a0 = p.offsets[u];
  const int n = (int)(p.offsets[u + 1] - a0);
  const int nframes = 1 + n / st;
  const int f0 = blockIdx.x * NF;
  if (f0 >= nframes) return;
  const int nfr = min(NF, nframes - f0);
  const int tid = threadIdx.x;
  for (int i = tid; i < 400; i += NTP) { tw[i] = p.tw[i]; win[i] = p.window[i]; }
  const float* au = p.audio + a0;
  const int s0 = f0 * st - half;
  const int need = (nfr - 1) * st + ws;
  if (s0 >= 0 && s0 + need <= n) {
    for (int i = tid; i < need; i += NTP) span[i] = __ldcs(au + s0 + i);
  } else {
    for (int i = tid; i < need; i += NTP) {
      int idx = s0 + i;
      if (idx < 0) idx = -idx;
      if (idx >= n) idx = 2 * (n - 1) - idx;
      span[i] = au[idx];
    }
  }
  __syncthreads();
  const int pr = tid / 20, c = tid - pr * 20;       // pair, column (n2 in pass 1, k1 in pass 2)
  const bool has_a = 2 * pr < nfr, has_b = 2 * pr + 1 < nfr;
  float2 x[20];
  {
    const float* sa = span + (2 * pr) * st + c;
    const float* sb = sa + st;
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
      const float w = win[20 * n1 + c];
      x[n1] = make_float2(has_a ? sa[20 * n1] * w : 0.f, has_b ? sb[20 * n1] * w : 0.f);
    }
  }
  dft20(x);                                         // over n1 (stride 20): X[k1] for this n2
  float2* exp_ = ex + pr * 20 * EXP;
#pragma unroll
  for (int k1 = 0; k1 < 20; ++k1) exp_[k1 * EXP + c] = k1 ? cmul(x[k1], tw[c * k1]) : x[k1];
  __syncthreads();
#pragma unroll
  for (int n2 = 0; n2 < 20; ++n2) x[n2] = exp_[c * EXP + n2];
  dft20(x);                                         // over n2: Z[k1 + 20 k2] = x[k2]
  __syncthreads();
#pragma unroll
  for (int k2 = 0; k2 < 20; ++k2) exp_[c + 20 * k2] = x[k2];      // linear [400] inside the pair's region
  __syncthreads();
  for (int k = c; k <= 200; k += 20) {
    const float2 zk = exp_[k == 400 ? 0 : k];
    const float2 zn = exp_[k == 0 ? 0 : 400 - k];
    const float ar = zk.x + zn.x, ai = zk.y - zn.y;              // Z[k] + conj Z[400-k] = 2 X_a[k]
    const float br = zk.x - zn.x, bi = zk.y + zn.y;              // Z[k] - conj Z[400-k] = 2i X_b[k]
    P[(2 * pr) * PB + k] = 0.25f * (ar * ar + ai * ai);
    P[(2 * pr + 1) * PB + k] = 0.25f * (br * br + bi * bi);
  }
  __syncthreads();
  float* outp = p.out + (size_t)(p.out_offsets[u] + f0) * p.n_mels;
  for (int w = tid; w < nfr * p.n_mels; w += NTP) {
    const int f = w / p.n_mels, i = w - f * p.n_mels;
    const int b0 = __ldg(p.mel_start + i), cnt = __ldg(p.mel_cnt + i);
    const float* wt = p.mel_w + __ldg(p.mel_off + i);
    const float* prw = P + f * PB + b0;
    float s = 0.f;
    for (int b = 0; b < cnt; ++b) s = fmaf(__ldg(wt + b), prw[b], s);
    __stcs(outp + w, logf(s + 2.220446049250313e-16f));
  }
}

// ---- 16 kHz / 25 ms fast path, second generation (round 2).  The first-generation kernel above issued 844 warp
// instructions per frame for ~ 310 of FFT work (issue-bound at 0.11 of the HBM roofline).  Same 20 x 20 register FFT of
// frame PAIRS (z = a + i b), everything around it rebuilt:
//  * the two frames are separated BETWEEN the passes, in registers: after the first 20-point DFT (over n1, thread = column
//    n2) a real frame's rows obey Ya[20-k1] = conj Ya[k1], so Ya[k1] = Y[k1] + conj Y[20-k1], Yb[k1] ~ Y[k1] - conj Y[20-k1]
//    (k1 = 1..9; the 1/2 is folded into the window) and rows 0 and 10 are real for both frames, i.e. Y[0] and Y[10] ARE the
//    packed pairs.  The second pass then runs exactly 20 row DFTs per pair again -- a1..a9, b1..b9, packed row 0, packed
//    row 10 -- whose 20 outputs are 20 DISTINCT wanted bins (k1 + 20 k2 for k2 < 10, the mirror bin (20-k1) + 20 (19-k2)
//    otherwise): no mirror exchange through shared memory, no barrier for it, |X|^2 straight from the DFT registers.  Only the
//    two packed rows need the conj-symmetric split, inside one thread; all 16 of them sit in warp 4, so four warps of five
//    never execute that path;
//  * a CTA walks QG groups of 16 frames; the audio span of group g+1 is prefetched with cp.async into the other of two
//    buffers while group g computes (the exposed load was 25 % of all stall samples), stored with 20 floats of padding per
//    320 so that the strided window reads of adjacent pairs fall on disjoint banks;
//  * twiddles: 10 per column (rows a_k and b_k share theirs), five 128-bit loads;
//  * power spectrum interleaved (a, b) per bin, pair stride = 4 banks (mod 32), living in the dead span buffer (no barrier
//    between the second-pass reads and the stores); mel projection: work item = (mel, pair), one weight feeds both frames,
//    the lanes of a half-warp are 8 pairs x 2 adjacent mels whose first bins differ in parity (padded on the host) =
//    conflict-free 64-bit loads; 4 adjacent mels per warp have near-equal widths (the first kernel's thread per
//    (frame, mel) made every warp wait for its widest filter); weights padded to 4 bins: one 128-bit load, no remainder loop;
//    log via lg2.approx (abs. error < 1e-5 over the range of log-mel energies).
constexpr int QG = 8;              // groups of 16 frames per CTA
constexpr int QPB = 210;           // power-spectrum row pitch (float2): 420 words = 4 banks mod 32
constexpr int QSPAN = 2960;        // span with 20 floats of padding after every 320 samples
constexpr int QBUF = FPAIRS * QPB * 2;           // floats per span / power-spectrum buffer (3360 >= QSPAN)
constexpr int QSMEM = (2 * QBUF + 400) * 4 + (20 * TWP + FPAIRS * 20 * EXP) * 8;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// ln(x) for normal positive x (here x >= 2^-52): one MUFU and one multiply, abs. error < 1e-5 over the log-mel range
__device__ __forceinline__ float ln_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * 0.6931471805599453f;
}

// span of the 16 frames from f0 on -> buf (padded layout), asynchronously
__device__ __forceinline__ void q_prefetch(float* buf, const float* au, int n, int nframes, int f0, int tid) {
  constexpr int ws = 400, st = 160, half = 200, NF = 2 * FPAIRS, SPAN = (NF - 1) * st + ws;
  const int nfr = min(NF, nframes - f0);
  const int s0 = f0 * st - half;
  if (nfr == NF && s0 >= 0 && s0 + SPAN <= n && ((reinterpret_cast<uintptr_t>(au + s0) & 15) == 0)) {
    // 700 chunks of 16 bytes, chunk j = tid + 160 t at float 4 j + 20 (j / 80) = 4 tid + 20 (tid / 80) + 680 t
    float* d = buf + 4 * tid + (tid >= 80 ? 20 : 0);
    const float* sp = au + s0 + 4 * tid;
#pragma unroll
    for (int t = 0; t < 4; ++t) cp_async16(d + 680 * t, sp + 640 * t);
    if (tid < SPAN / 4 - 4 * NTP) cp_async16(d + 680 * 4, sp + 640 * 4);
  } else {                                                        // utterance edges (numpy 'reflect' padding), short last
    const int need = (nfr - 1) * st + ws;                         // group (zeros: a pair partner reads them), unaligned audio
    for (int i = tid; i < SPAN; i += NTP) {
      int idx = s0 + i;
      if (idx < 0) idx = -idx;
      if (idx >= n) idx = 2 * (n - 1) - idx;
      float* d = buf + i + 20 * (i / 320);
      if (i < need) cp_async4(d, au + idx); else *d = 0.f;
    }
  }
}

__global__ void __launch_bounds__(NTP, 4) fbank400q_kernel(FbankParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int st = 160, NF = 2 * FPAIRS;
  float* win = smem + 2 * QBUF;                                   // [400]
  float2* twc = reinterpret_cast<float2*>(win + 400);             // [20][TWP]
  float2* ex = twc + 20 * TWP;                                    // [FPAIRS][20*EXP]: rows a1..a8, b1..b8, a9, b9, r0, r10
  const int u = blockIdx.y;
  const long long a0 = p.offsets[u];
  const int n = (int)(p.offsets[u + 1] - a0);
  const int nframes = 1 + n / st;
  const int tid = threadIdx.x;
  const int fbase = blockIdx.x * (QG * NF);
  if (fbase >= nframes) return;
  const float* au = p.audio + a0;
  q_prefetch(smem, au, n, nframes, fbase, tid);
  for (int i = tid; i < 400; i += NTP) win[i] = __ldg(p.win_half + i);
  for (int i = tid; i < 20 * TWP; i += NTP) twc[i] = __ldg(p.twc + i);
  for (int i = QSPAN + tid; i < QBUF; i += NTP) { smem[i] = 0.f; smem[QBUF + i] = 0.f; }   // never written by a span: the
  const int pr = tid / 20, c = tid - pr * 20;       // first pass: pair, column n2         // padded filters read them as P
  smem[340 * pr + 320 + c] = 0.f;                   // (the padding gaps too)
  smem[QBUF + 340 * pr + 320 + c] = 0.f;
  // second pass: pair q, row.  warps 0-3: pairs (w, w+4) x rows 0..15; warp 4: 4 pairs x rows 16..19 per half-warp
  int q, row;
  if (tid < 128) { q = (tid >> 5) + 4 * ((tid >> 4) & 1); row = tid & 15; }
  else { const int l = tid - 128; q = (l & 3) + 4 * (l >> 4); row = 16 + ((l >> 2) & 3); }
  const bool special = row >= 18, is_r0 = row == 18;
  const int fr = row < 16 ? row >> 3 : row & 1;     // frame of the pair (normal rows)
  const int k1 = row < 16 ? (row & 7) + 1 : 9;
  float* const out_u = p.out + (size_t)p.out_offsets[u] * p.n_mels;
#pragma unroll 1
  for (int g = 0; g < QG; ++g) {
    const int f0 = fbase + g * NF;
    if (f0 >= nframes) break;
    const int nfr = min(NF, nframes - f0);
    float* span = smem + (g & 1) * QBUF;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                                              // B1: span g landed; everybody is done with group g-1
    if (g + 1 < QG && f0 + NF < nframes) q_prefetch(smem + ((g + 1) & 1) * QBUF, au, n, nframes, f0 + NF, tid);
    float2 x[20];
    {
      const float* sa = span + 340 * pr + c;
      const float* wv = win + c;
#pragma unroll
      for (int n1 = 0; n1 < 20; ++n1) {
        const float w = wv[20 * n1];
        x[n1] = make_float2(sa[20 * n1 + (n1 >= 16 ? 20 : 0)] * w, sa[st + 20 * n1 + (n1 >= 8 ? 20 : 0)] * w);
      }
    }
    dft20(x);                                                     // over n1 (stride 20): Y[k1] for this n2
    {
      float2* e = ex + pr * 20 * EXP + c;
      const float4* tq = reinterpret_cast<const float4*>(twc + c * TWP);
      e[18 * EXP] = x[0];                                         // packed row 0: Ya[0] + i Yb[0], both real
      float4 t2;
#pragma unroll
      for (int k = 1; k <= 9; ++k) {
        if (k & 1) t2 = tq[k >> 1];                               // W^(c k), W^(c (k+1))
        const float2 tw = (k & 1) ? make_float2(t2.x, t2.y) : make_float2(t2.z, t2.w);
        const float2 y = x[k], m = x[20 - k];
        const float2 ya = make_float2(y.x + m.x, y.y - m.y);      // Y[k] + conj Y[20-k]
        const float2 yb = make_float2(y.x - m.x, y.y + m.y);      // Y[k] - conj Y[20-k] = i * (row of frame b)
        e[(k <= 8 ? k - 1 : 16) * EXP] = cmul(ya, tw);
        e[(k <= 8 ? k + 7 : 17) * EXP] = cmul(yb, tw);
      }
      e[19 * EXP] = cmul(x[10], make_float2(t2.z, t2.w));         // packed row 10, W^(10 c)
    }
    __syncthreads();                                              // B2: the span buffer is dead from here on
    {
      const float2* e = ex + q * 20 * EXP + row * EXP;
#pragma unroll
      for (int n2 = 0; n2 < 20; ++n2) x[n2] = e[n2];
    }
    dft20(x);                                                     // over n2: X[k1 + 20 k2] = x[k2]
    float* Pw = span;                                             // power spectrum [FPAIRS][QPB] x (a, b) in the dead span
    if (!special) {
      float* lo = Pw + q * (2 * QPB) + fr + 2 * k1;               // bins k1 + 20 k2, k2 < 10
      float* hi = Pw + q * (2 * QPB) + fr + 2 * (20 - k1);        // bins (20 - k1) + 20 (19 - k2), k2 >= 10
#pragma unroll
      for (int k2 = 0; k2 < 10; ++k2) lo[40 * k2] = x[k2].x * x[k2].x + x[k2].y * x[k2].y;
#pragma unroll
      for (int k2 = 10; k2 < 20; ++k2) hi[40 * (19 - k2)] = x[k2].x * x[k2].x + x[k2].y * x[k2].y;
    } else {
      // packed rows: C[k2] = Xa[k] + i Xb[k] (k = 20 k2 or 10 + 20 k2), mirror C[(20-k2) % 20] or C[19-k2]
      float2* po = reinterpret_cast<float2*>(Pw) + q * QPB + (is_r0 ? 0 : 10);
#pragma unroll
      for (int k2 = 0; k2 < 10; ++k2) {
        const float2 zk = x[k2];
        const float2 z0 = x[(20 - k2) % 20], z1 = x[19 - k2];
        const float2 zn = is_r0 ? z0 : z1;
        const float ar = zk.x + zn.x, ai = zk.y - zn.y, br = zk.x - zn.x, bi = zk.y + zn.y;
        po[20 * k2] = make_float2(ar * ar + ai * ai, br * br + bi * bi);
      }
      if (is_r0) po[200] = make_float2(4.f * x[10].x * x[10].x, 4.f * x[10].y * x[10].y);   // bin 200: its own mirror
    }
    __syncthreads();                                              // B4
    float* outp = out_u + (size_t)f0 * p.n_mels;
    if (2 * (tid & (FPAIRS - 1)) < nfr) {
      const int pq = tid & (FPAIRS - 1);
      const float2* prow = reinterpret_cast<const float2*>(Pw) + pq * QPB;
      float* orow = outp + (2 * pq) * p.n_mels;
      const bool has_b = 2 * pq + 1 < nfr;
      constexpr int IS = NTP / FPAIRS;           // mel stride between a thread's items
      for (int i0 = tid >> 3; i0 < p.n_mels; i0 += 4 * IS) {
        // four items at a time: their filter metadata, then their first four weights, are all in flight together (one item
        // after the other exposed two dependent L1 latencies per item to a warp that has only 19 others to hide behind)
        int2 mm[4];
        float4 w0[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) mm[j] = i0 + j * IS < p.n_mels ? __ldg(p.mel_meta + i0 + j * IS) : make_int2(0, 0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          w0[j] = (mm[j].x >> 16) > 0 ? __ldg(reinterpret_cast<const float4*>(p.mel_w4 + mm[j].y)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j * IS;
          if (i >= p.n_mels) break;
          const float4* wt = reinterpret_cast<const float4*>(p.mel_w4 + mm[j].y) + 1;
          const float2* pw = prow + (mm[j].x & 0xffff);
          const float2* pe = pw + (mm[j].x >> 16);
          float sa0 = 0.f, sb0 = 0.f, sa1 = 0.f, sb1 = 0.f;
          float4 w4 = w0[j];
#pragma unroll 1
          for (; pw < pe; pw += 4, ++wt) {
            const float4 wn = pw + 4 < pe ? __ldg(wt) : w4;          // next four weights: under this iteration's FMAs
            const float2 p0 = pw[0], p1 = pw[1], p2 = pw[2], p3 = pw[3];
            sa0 = fmaf(w4.x, p0.x, sa0); sb0 = fmaf(w4.x, p0.y, sb0);
            sa1 = fmaf(w4.y, p1.x, sa1); sb1 = fmaf(w4.y, p1.y, sb1);
            sa0 = fmaf(w4.z, p2.x, sa0); sb0 = fmaf(w4.z, p2.y, sb0);
            sa1 = fmaf(w4.w, p3.x, sa1); sb1 = fmaf(w4.w, p3.y, sb1);
            w4 = wn;
          }
          __stcs(orow + i, ln_fast(sa0 + sa1 + 2.220446049250313e-16f));
          if (has_b) __stcs(orow + p.n_mels + i, ln_fast(sb0 + sb1 + 2.220446049250313e-16f));
        }
      }
    }
    // group g+1: its prefetch into THIS buffer is issued after its B1, which no thread passes before everybody has finished
    // this mel phase; its ex writes come after that B1 too
  }
}

// ---- any other window length (the reference's default 22.05 kHz gives 551 = 19*29): direct DFT ---------------
__global__ void __launch_bounds__(NT) fbank_generic_kernel(FbankParams p) {
  extern __shared__ __align__(16) float smem[];
  const int u = blockIdx.y;
  const long long a0 = p.offsets[u];
  const int n = (int)(p.offsets[u + 1] - a0);
  const int ws = p.ws, st = p.st, half = ws / 2;
  const int nframes = 1 + (n + 2 * half - ws) / st;
  const int f0 = blockIdx.x * FPB;
  if (f0 >= nframes) return;
  const int nfr = min(FPB, nframes - f0);
  const int span_len = (FPB - 1) * st + ws;
  float* span = smem;                                            // [span_len]
  float* bufA = span + ((span_len + 3) & ~3);                    // [FPB][ws] windowed frames
  float* P = bufA + ((FPB * ws + 3) & ~3);                       // [FPB][nbins+1]
  float2* tws = reinterpret_cast<float2*>(P + ((FPB * (p.nbins + 1) + 3) & ~3));
  for (int i = threadIdx.x; i < ws; i += NT) tws[i] = p.tw[i];
  const float* au = p.audio + a0;
  const int s0 = f0 * st - half;
  const int need = (nfr - 1) * st + ws;
  for (int i = threadIdx.x; i < need; i += NT) {
    int idx = s0 + i;
    if (idx < 0) idx = -idx;
    if (idx >= n) idx = 2 * (n - 1) - idx;
    span[i] = au[idx];
  }
  __syncthreads();
  const int PB = p.nbins + 1;
  for (int w = threadIdx.x; w < nfr * ws; w += NT) {
    const int f = w / ws, i = w % ws;
    bufA[f * ws + i] = span[f * st + i] * p.window[i];
  }
  __syncthreads();
  for (int w = threadIdx.x; w < nfr * p.nbins; w += NT) {
    const int f = w / p.nbins, k = w % p.nbins;
    const float* x = bufA + f * ws;
    float re = 0.f, im = 0.f;
    int idx = 0;
    for (int i = 0; i < ws; ++i) {
      const float2 t = tws[idx];
      re = fmaf(x[i], t.x, re);
      im = fmaf(x[i], t.y, im);
      idx += k;
      if (idx >= ws) idx -= ws;
    }
    P[f * PB + k] = re * re + im * im;
  }
  __syncthreads();
  float* outp = p.out + (size_t)(p.out_offsets[u] + f0) * p.n_mels;
  for (int w = threadIdx.x; w < nfr * p.n_mels; w += NT) {
    const int f = w / p.n_mels, i = w % p.n_mels;
    const int b0 = p.mel_start[i], c = p.mel_cnt[i];
    const float* wt = p.mel_w + p.mel_off[i];
    const float* pr = P + f * PB + b0;
    float s = 0.f;
    for (int b = 0; b < c; ++b) s = fmaf(wt[b], pr[b], s);
    outp[w] = logf(s + 2.220446049250313e-16f);
  }
}

}  // namespace ssasr

using namespace ssasr;

extern "C" {

// frames produced for an utterance of n_samples (preprocess.py:194-198 + librosa centred STFT)
long long ssasr_fbank_num_frames(long long n_samples, int sample_rate) {
  const int ws = (int)(sample_rate * 0.001 * 25), st = (int)(sample_rate * 0.001 * 10);
  return 1 + (n_samples + 2 * (ws / 2) - ws) / st;
}

// audio: concatenated fp32 samples; offsets int64 [n_utt+1] (device): utterance u = audio[offsets[u]:offsets[u+1]]
// out: fp32 [total_frames, n_mels]; out_offsets int64 [n_utt+1] (device): first output row of every utterance
// max_frames: largest per-utterance frame count (host value, sizes the grid)
int ssasr_fbank(const float* audio, const long long* offsets, int n_utt, int sample_rate, int n_mels, float* out,
                const long long* out_offsets, int max_frames, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSASR_REQUIRE(n_utt > 0 && n_utt <= 65535, "fbank: n_utt=%d out of range (1..65535 per call)", n_utt);
  SSASR_REQUIRE(n_mels > 0 && sample_rate >= 1000, "fbank: bad n_mels=%d / sample_rate=%d", n_mels, sample_rate);
  FbankTables t;
  int rc = get_tables(sample_rate, n_mels, &t);
  if (rc) return rc;
  FbankParams p;
  p.audio = audio; p.offsets = offsets; p.n_utt = n_utt; p.out = out; p.out_offsets = out_offsets;
  p.ws = t.ws; p.st = t.st; p.nbins = t.nbins; p.n_mels = n_mels;
  p.window = t.window; p.tw = t.tw; p.mel_start = t.mel_start; p.mel_cnt = t.mel_cnt; p.mel_off = t.mel_off; p.mel_w = t.mel_w;
  dim3 grid((max_frames + FPB - 1) / FPB, n_utt);
  ProfScope ps(F_FBANK, st);
  static const bool use_gen1 = getenv("SSASR_FBANK_GEN1") != nullptr;   // first-generation register-FFT kernel, kept for A/B timing
  if (t.ws == 400 && !use_gen1) {
    p.win_half = t.win_half; p.twc = t.twc; p.mel_meta = t.mel_meta; p.mel_w4 = t.mel_w4;
    constexpr int NF = 2 * FPAIRS;
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(fbank400q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, QSMEM));
    dim3 g2((max_frames + QG * NF - 1) / (QG * NF), n_utt);
    fbank400q_kernel<<<g2, NTP, QSMEM, st>>>(p);
  } else if (t.ws == 400) {
    p.tw = t.tw400;
    constexpr int NF = 2 * FPAIRS;
    const size_t smem = (size_t)((NF - 1) * 160 + 400 + 400) * sizeof(float) + (size_t)(400 + FPAIRS * 20 * EXP) * sizeof(float2) +
                        (size_t)NF * 202 * sizeof(float);
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(fbank400p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 g2((max_frames + NF - 1) / NF, n_utt);
    fbank400p_kernel<<<g2, NTP, smem, st>>>(p);
  } else {
    const int span_len = (FPB - 1) * t.st + t.ws;
    const size_t floats = ((span_len + 3) & ~3) + ((FPB * t.ws + 3) & ~3) + ((FPB * (t.nbins + 1) + 3) & ~3) + 2 * (size_t)t.ws;
    const size_t smem = floats * sizeof(float);
    SSASR_REQUIRE(smem <= 227 * 1024, "fbank: window of %d samples needs %zu B shared memory", t.ws, smem);
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(fbank_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fbank_generic_kernel<<<grid, NT, smem, st>>>(p);
  }
  SSASR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

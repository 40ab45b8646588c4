// Batched log-mel filterbank frontend: framing (reflect-padded, centred), periodic Hann window, real FFT,
// power spectrum, Slaney mel projection and log, in ONE kernel, output already [frames, n_mels].
//
// Reference semantics: log_fbank, /root/reference/src/preprocess.py:187-208, i.e.
// librosa(0.6.3).feature.melspectrogram(y, sr, n_mels=N_DIMS, n_fft=ws, hop_length=st) -> log(. + eps) -> T.
//
// One CTA handles FPB consecutive frames of one utterance: the audio span they share is read once
// (hop < window, every sample is used by 2.5 frames), frames are built in shared memory, the
// 25 ms window at 16 kHz (400 samples) is transformed as a 200-point complex FFT (radix 5*5*4*2
// Stockham, one butterfly per thread) plus the real-input split; other window sizes (the reference's
// default 22.05 kHz gives 551 = 19*29) take a direct-DFT path in the same kernel.  The mel matrix is
// applied in its sparse (two triangles per bin) form.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include <mutex>
#include <vector>

namespace ssasr {

constexpr int FPB = 16;            // frames per CTA
constexpr int NT = 256;

struct FbankTables {
  int sr = 0, n_mels = 0, ws = 0, st = 0, nbins = 0;
  float* window = nullptr;     // [ws]
  float2* tw = nullptr;        // fast path: [200] W_200^m then [201] W_400^k ; generic: [ws] W_ws^m
  float2* tw400 = nullptr;     // 16 kHz register-FFT path: [400] W_400^e
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

// one Stockham stage of radix R (sub-transform length Ns so far) over `nfr` frames of N=200 complex points
// (src -> dst); everything but nfr is a compile-time constant, so the index arithmetic is mul/shift only.
template <int R, int Ns>
__device__ __forceinline__ void stockham_stage(const float2* __restrict__ src, float2* __restrict__ dst, const float2* __restrict__ tw200,
                                               int nfr) {
  constexpr int N = 200;
  constexpr int NB = N / R;
  constexpr int TS = N / (Ns * R);      // twiddle index step per unit k: W_{Ns*R}^{k r} = W_200^{k r TS}  (k r TS < N)
  for (int w = threadIdx.x; w < nfr * NB; w += NT) {
    const int f = w / NB, j = w - f * NB;
    const float2* x = src + f * N;
    float2* y = dst + f * N;
    const int k = j % Ns;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      v[r] = x[j + r * NB];
      if (r > 0 && Ns > 1) v[r] = cmul(v[r], tw200[k * TS * r]);
    }
    float2 o[R];
    if constexpr (R == 2) {
      o[0] = cadd(v[0], v[1]);
      o[1] = csub(v[0], v[1]);
    } else if constexpr (R == 4) {
      const float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]), t2 = cadd(v[1], v[3]), t3 = csub(v[1], v[3]);
      o[0] = cadd(t0, t2);
      o[2] = csub(t0, t2);
      o[1] = make_float2(t1.x + t3.y, t1.y - t3.x);   // t1 - i*t3
      o[3] = make_float2(t1.x - t3.y, t1.y + t3.x);   // t1 + i*t3
    } else {  // R == 5
      const float c1 = 0.30901699437494745f, c2 = -0.8090169943749475f, s1 = 0.9510565162951535f, s2 = 0.5877852522924731f;
      const float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
      o[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
      const float2 a1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
      const float2 a2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
      const float2 b1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
      const float2 b2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
      o[1] = make_float2(a1.x + b1.y, a1.y - b1.x);   // a1 - i*b1
      o[4] = make_float2(a1.x - b1.y, a1.y + b1.x);
      o[2] = make_float2(a2.x + b2.y, a2.y - b2.x);
      o[3] = make_float2(a2.x - b2.y, a2.y + b2.x);
    }
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) y[j0 + r * Ns] = o[r];
  }
}

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

// ---- 16 kHz / 25 ms fast path: 400-sample frames, hop 160 ----------------------------------------------------
// shared: A [FPB][200] float2 | B [FPB][200] float2 (also: audio span before the FFT, power spectrum after) |
//         twiddles [401] float2 | window [400] float
__global__ void __launch_bounds__(NT, 4) fbank400_kernel(FbankParams p) {
  extern __shared__ __align__(16) float smem[];
  float2* A = reinterpret_cast<float2*>(smem);
  float2* Bf = A + FPB * 200;
  float* span = reinterpret_cast<float*>(Bf);                 // aliases B until the first FFT stage writes it
  float* P = reinterpret_cast<float*>(Bf);                    // aliases B after the last FFT stage (result is in A)
  float2* tws = Bf + FPB * 200;
  float* win = reinterpret_cast<float*>(tws + 401);
  const int u = blockIdx.y;
  const long long a0 = p.offsets[u];
  const int n = (int)(p.offsets[u + 1] - a0);
  constexpr int ws = 400, st = 160, half = 200;
  const int nframes = 1 + n / st;
  const int f0 = blockIdx.x * FPB;
  if (f0 >= nframes) return;
  const int nfr = min(FPB, nframes - f0);
  for (int i = threadIdx.x; i < 401; i += NT) tws[i] = p.tw[i];
  for (int i = threadIdx.x; i < ws; i += NT) win[i] = p.window[i];
  const float* au = p.audio + a0;
  const int s0 = f0 * st - half;
  const int need = (nfr - 1) * st + ws;
  if (s0 >= 0 && s0 + need <= n) {          // interior block: straight coalesced copy
    for (int i = threadIdx.x; i < need; i += NT) span[i] = __ldcs(au + s0 + i);
  } else {                                  // utterance edges: numpy 'reflect' padding
    for (int i = threadIdx.x; i < need; i += NT) {
      int idx = s0 + i;
      if (idx < 0) idx = -idx;
      if (idx >= n) idx = 2 * (n - 1) - idx;
      span[i] = au[idx];
    }
  }
  __syncthreads();
  for (int w = threadIdx.x; w < nfr * 200; w += NT) {          // window and pack z[m] = xw[2m] + i*xw[2m+1]
    const int f = w / 200, m = w - f * 200;
    const float2 sv = *reinterpret_cast<const float2*>(span + f * st + 2 * m);
    const float2 wv = *reinterpret_cast<const float2*>(win + 2 * m);
    A[w] = make_float2(sv.x * wv.x, sv.y * wv.y);
  }
  __syncthreads();
  stockham_stage<5, 1>(A, Bf, tws, nfr);
  __syncthreads();
  stockham_stage<5, 5>(Bf, A, tws, nfr);
  __syncthreads();
  stockham_stage<4, 25>(A, Bf, tws, nfr);
  __syncthreads();
  stockham_stage<2, 100>(Bf, A, tws, nfr);
  __syncthreads();
  // real-input split: X[k] = E[k] + W_400^k * O[k], k = 0..200; power spectrum
  constexpr int PB = 202;
  const float2* w400 = tws + 200;
  for (int w = threadIdx.x; w < nfr * 201; w += NT) {
    const int f = w / 201, k = w - f * 201;
    const float2 zk = A[f * 200 + (k == 200 ? 0 : k)];
    const float2 zn = A[f * 200 + (k == 0 ? 0 : 200 - k)];
    const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
    const float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));   // (zk - conj(zn)) / (2i)
    const float2 x = cadd(e, cmul(w400[k], o));
    P[f * PB + k] = x.x * x.x + x.y * x.y;
  }
  __syncthreads();
  float* outp = p.out + (size_t)(p.out_offsets[u] + f0) * p.n_mels;
  for (int w = threadIdx.x; w < nfr * p.n_mels; w += NT) {
    const int f = w / p.n_mels, i = w - f * p.n_mels;
    const int b0 = __ldg(p.mel_start + i), c = __ldg(p.mel_cnt + i);
    const float* wt = p.mel_w + __ldg(p.mel_off + i);
    const float* pr = P + f * PB + b0;
    float s = 0.f;
    for (int b = 0; b < c; ++b) s = fmaf(__ldg(wt + b), pr[b], s);
    __stcs(outp + w, logf(s + 2.220446049250313e-16f));
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
  static const bool use_stockham = getenv("SSASR_FBANK_STOCKHAM") != nullptr;   // previous smem-FFT kernel, kept for A/B timing
  if (t.ws == 400 && !use_stockham) {
    p.tw = t.tw400;
    constexpr int NF = 2 * FPAIRS;
    const size_t smem = (size_t)((NF - 1) * 160 + 400 + 400) * sizeof(float) + (size_t)(400 + FPAIRS * 20 * EXP) * sizeof(float2) +
                        (size_t)NF * 202 * sizeof(float);
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(fbank400p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 g2((max_frames + NF - 1) / NF, n_utt);
    fbank400p_kernel<<<g2, NTP, smem, st>>>(p);
  } else if (t.ws == 400) {
    const size_t smem = (size_t)(2 * FPB * 200 + 401) * sizeof(float2) + 400 * sizeof(float);
    SSASR_CHECK_CUDA(cudaFuncSetAttribute(fbank400_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fbank400_kernel<<<grid, NT, smem, st>>>(p);
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

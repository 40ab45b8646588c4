// fp32 SIMT GEMM used by the exact (fp32) path: small / odd-shaped products and every product whose
// result feeds a greedy argmax (greedy transcripts need fp32-accurate math).
// The batched-over-time gate GEMMs of the bf16 training path run on tcgen05 (gemm_tc.cu).
#include "common.cuh"
#include <stdarg.h>

namespace ssasr {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

constexpr int BM = 64, BN = 64, BK = 16;

// C[M,N] = (acc ? C : 0) + A·B (+bias) (+tanh).  256 threads, 4x4 outputs per thread.
__global__ void __launch_bounds__(256) gemm_f32_kernel(int M, int N, int K, const float* __restrict__ A, int lda, int akm,
                                                       const float* __restrict__ B, int ldb, int bkm, float* __restrict__ C,
                                                       int ldc, const float* __restrict__ bias, int accumulate, int act_tanh,
                                                       int zero_period, int zero_pos) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int m, k;
      if (akm) { m = idx / BK; k = idx % BK; } else { k = idx / BM; m = idx % BM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.0f;
      if (gm < M && gk < K && !(zero_period > 0 && (gk % zero_period) == zero_pos))
        v = akm ? A[(size_t)gm * lda + gk] : A[(size_t)gk * lda + gm];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int n, k;
      if (bkm) { n = idx / BK; k = idx % BK; } else { k = idx / BN; n = idx % BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.0f;
      if (gn < N && gk < K) v = bkm ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      float* p = C + (size_t)gm * ldc + gn;
      if (accumulate) v += *p;
      if (act_tanh) v = tanhf(v);
      *p = v;
    }
  }
}

int gemm_f32(cudaStream_t st, int M, int N, int K, const float* A, int lda, int a_kmajor, const float* B, int ldb,
             int b_kmajor, float* C, int ldc, const float* bias, int accumulate, int act_tanh, int zero_period,
             int zero_pos) {
  if (M <= 0 || N <= 0) return 0;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
  SSASR_REQUIRE(grid.y <= 65535, "gemm_f32: N=%d too large for grid.y", N);
  ProfScope ps(F_GEMM_F32, st);
  gemm_f32_kernel<<<grid, 256, 0, st>>>(M, N, K, A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, bias, accumulate, act_tanh,
                                        zero_period, zero_pos);
  SSASR_LAUNCH_CHECK();
  return 0;
}

}  // namespace ssasr

extern "C" {
const char* ssasr_last_error(void) { return ssasr::last_error(); }

int ssasr_gemm_f32(int M, int N, int K, const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor,
                   float* C, int ldc, const float* bias, int accumulate, int act_tanh, void* stream) {
  return ssasr::gemm_f32((cudaStream_t)stream, M, N, K, A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, bias, accumulate,
                         act_tanh, 0, 0);
}
}

// Solver.step of the reference (trainer.py:131-148 with the Adadelta of trainer.py:401-403) as ONE multi-tensor pass:
//   grad_norm = clip_grad_norm_(params, max_norm);  if isnan(grad_norm): skip  else: Adadelta step
// Two launches for all parameter tensors together (the 46 tensors of the LAS model took ~40 launches of torch's foreach
// kernels plus a host sync for the NaN test): (1) per-chunk sums of squares into a scratch array, (2) every block re-reduces
// the scratch in a fixed order (deterministic total, no atomics), derives the clip coefficient and the NaN decision on the
// device and applies torch.optim.Adadelta's update (torch/optim/adadelta.py, weight_decay = 0, maximize = False):
//   sq = rho*sq + (1-rho)*g*g;  delta = sqrt(acc+eps)/sqrt(sq+eps)*g;  acc = rho*acc + (1-rho)*delta*delta;  p -= lr*delta
#include "common.cuh"

namespace ssasr {

constexpr int OPT_MAXT = 64;              // tensors per launch (kernel-parameter table)
constexpr int OPT_CHUNK = 16384;          // elements per block (16 float4 per thread)

struct OptTable {
  float* p[OPT_MAXT];
  float* g[OPT_MAXT];
  float* sq[OPT_MAXT];
  float* acc[OPT_MAXT];
  long long n[OPT_MAXT];
  int first_chunk[OPT_MAXT + 1];          // prefix sums of ceil(n / OPT_CHUNK)
  int nt;
};

__device__ __forceinline__ void locate(const OptTable& t, int blk, int& ti, long long& off, long long& cnt) {
  int i = 0;
  while (i + 1 < t.nt && blk >= t.first_chunk[i + 1]) ++i;
  ti = i;
  off = (long long)(blk - t.first_chunk[i]) * OPT_CHUNK;
  cnt = t.n[i] - off;
  if (cnt > OPT_CHUNK) cnt = OPT_CHUNK;
}

__global__ void __launch_bounds__(256) optim_sumsq_kernel(const __grid_constant__ OptTable t, float* __restrict__ partial, int part_off) {
  __shared__ float scratch[32];
  int ti;
  long long off, cnt;
  locate(t, blockIdx.x, ti, off, cnt);
  const float* g = t.g[ti] + off;
  float s = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const long long n4 = cnt >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
      s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < cnt; i += 256) s = fmaf(g[i], g[i], s);
  } else {
    for (long long i = threadIdx.x; i < cnt; i += 256) s = fmaf(g[i], g[i], s);
  }
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) partial[part_off + blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) optim_adadelta_kernel(const __grid_constant__ OptTable t, const float* __restrict__ partial,
                                                             int n_partial, float lr, float rho, float eps, float max_norm,
                                                             float* __restrict__ norm_out, int write_grads, int first_launch) {
  __shared__ float scratch[32];
  __shared__ float s_coef;
  // total norm: fixed-order reduction of the per-chunk sums (identical in every block)
  float s = 0.f;
  for (int i = threadIdx.x; i < n_partial; i += 256) s += partial[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    const float total = sqrtf(s);
    float coef = 1.f;
    if (max_norm > 0.f) {
      coef = max_norm / (total + 1e-6f);
      if (coef > 1.f) coef = 1.f;
    }
    const bool skip = isnan(total);                        // trainer.py:145-148: the step is cancelled
    s_coef = skip ? __int_as_float(0x7fc00000) : coef;
    if (blockIdx.x == 0 && first_launch) {
      norm_out[0] = total;
      norm_out[1] = skip ? 0.f : 1.f;
    }
  }
  __syncthreads();
  const float coef = s_coef;
  if (isnan(coef)) return;
  int ti;
  long long off, cnt;
  locate(t, blockIdx.x, ti, off, cnt);
  float* p = t.p[ti] + off;
  float* g = t.g[ti] + off;
  float* sq = t.sq[ti] + off;
  float* acc = t.acc[ti] + off;
  const float omr = 1.f - rho;
  auto upd = [&](float& pv, float& gv_io, float& sqv, float& accv) {
    const float gv = gv_io * coef;
    const float sv = sqv * rho + omr * (gv * gv);
    const float stdv = sqrtf(sv + eps);
    const float delta = sqrtf(accv + eps) / stdv * gv;
    sqv = sv;
    accv = accv * rho + omr * (delta * delta);
    pv = pv - lr * delta;
    gv_io = gv;
  };
  const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(sq) |
                    reinterpret_cast<uintptr_t>(acc)) & 15) == 0;
  long long done = 0;
  if (al) {
    const long long n4 = cnt >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* g4 = reinterpret_cast<float4*>(g);
    float4* s4 = reinterpret_cast<float4*>(sq);
    float4* a4 = reinterpret_cast<float4*>(acc);
#pragma unroll 2
    for (long long i = threadIdx.x; i < n4; i += 256) {
      float4 pv = p4[i], gv = g4[i], sv = s4[i], av = a4[i];
      upd(pv.x, gv.x, sv.x, av.x); upd(pv.y, gv.y, sv.y, av.y); upd(pv.z, gv.z, sv.z, av.z); upd(pv.w, gv.w, sv.w, av.w);
      p4[i] = pv; s4[i] = sv; a4[i] = av;
      if (write_grads) g4[i] = gv;
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < cnt; i += 256) {
    float pv = p[i], gv = g[i], sv = sq[i], av = acc[i];
    upd(pv, gv, sv, av);
    p[i] = pv; sq[i] = sv; acc[i] = av;
    if (write_grads) g[i] = gv;
  }
}

}  // namespace ssasr

using namespace ssasr;

extern "C" {

typedef struct {
  float* p;          // parameter (updated in place)
  float* g;          // gradient (rescaled in place only if write_clipped_grads)
  float* sq;         // Adadelta square_avg state
  float* acc;        // Adadelta acc_delta state
  long long n;       // elements
} ssasr_optim_tensor;

// scratch floats needed for `tensors`
long long ssasr_adadelta_scratch_floats(const ssasr_optim_tensor* tensors, int n_tensors) {
  long long c = 0;
  for (int i = 0; i < n_tensors; ++i) c += (tensors[i].n + OPT_CHUNK - 1) / OPT_CHUNK;
  return c + 2;
}

// tensors: HOST array of device pointers.  scratch: device floats (ssasr_adadelta_scratch_floats).  norm_out: device [2]:
// [0] = total gradient norm before clipping, [1] = 1 if the update was applied, 0 if it was skipped (NaN norm).
// max_norm <= 0: no clipping.
int ssasr_adadelta_clip_step(const ssasr_optim_tensor* tensors, int n_tensors, float lr, float rho, float eps, float max_norm,
                             float* scratch, float* norm_out, int write_clipped_grads, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_tensors <= 0) return 0;
  SSASR_REQUIRE(scratch && norm_out, "adadelta_clip_step: scratch / norm_out missing");
  const int n_launch = (n_tensors + OPT_MAXT - 1) / OPT_MAXT;
  OptTable tabs[16];
  SSASR_REQUIRE(n_launch <= 16, "adadelta_clip_step: too many tensors (%d)", n_tensors);
  int part_off[17];
  part_off[0] = 0;
  for (int l = 0; l < n_launch; ++l) {
    OptTable& t = tabs[l];
    memset(&t, 0, sizeof(t));
    const int lo = l * OPT_MAXT, hi = n_tensors < lo + OPT_MAXT ? n_tensors : lo + OPT_MAXT;
    t.nt = hi - lo;
    t.first_chunk[0] = 0;
    for (int i = lo; i < hi; ++i) {
      const int k = i - lo;
      SSASR_REQUIRE(tensors[i].n >= 0 && tensors[i].p && tensors[i].g && tensors[i].sq && tensors[i].acc,
                    "adadelta_clip_step: tensor %d has a null pointer", i);
      t.p[k] = tensors[i].p; t.g[k] = tensors[i].g; t.sq[k] = tensors[i].sq; t.acc[k] = tensors[i].acc; t.n[k] = tensors[i].n;
      t.first_chunk[k + 1] = t.first_chunk[k] + (int)((tensors[i].n + OPT_CHUNK - 1) / OPT_CHUNK);
    }
    part_off[l + 1] = part_off[l] + t.first_chunk[t.nt];
  }
  const int n_partial = part_off[n_launch];
  for (int l = 0; l < n_launch; ++l) {
    const int blocks = tabs[l].first_chunk[tabs[l].nt];
    if (blocks == 0) continue;
    ProfScope ps(F_OPTIM, st);
    optim_sumsq_kernel<<<blocks, 256, 0, st>>>(tabs[l], scratch, part_off[l]);
  }
  for (int l = 0; l < n_launch; ++l) {
    const int blocks = tabs[l].first_chunk[tabs[l].nt];
    if (blocks == 0) continue;
    ProfScope ps(F_OPTIM, st);
    optim_adadelta_kernel<<<blocks, 256, 0, st>>>(tabs[l], scratch, n_partial, lr, rho, eps, max_norm, norm_out, write_clipped_grads,
                                                  l == 0 ? 1 : 0);
  }
  SSASR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// prepare_x of the reference (ASRDataset.py:297-316): cast the DataLoader's (float64) batch to fp32 and count, per utterance,
// the frames whose feature sum is not zero (the zero-padding rule of preprocess.py).  The reference copies the whole batch back
// to the host for the count; here only the B lengths travel.  One warp per frame.
// ------------------------------------------------------------------------------------------------
namespace ssasr {
template <typename T>
__global__ void __launch_bounds__(256) prepare_x_kernel(const T* __restrict__ src, float* __restrict__ dst, int* __restrict__ lens,
                                                        long long n_frames, long long frames_per_utt, int F) {
  const long long f = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (f >= n_frames) return;
  const int lane = threadIdx.x & 31;
  const T* s = src + f * F;
  float* d = dst ? dst + f * F : nullptr;
  float sum = 0.f;
  for (int k = lane; k < F; k += 32) {
    const float v = (float)s[k];
    if (d) d[k] = v;
    sum += v;
  }
  sum = warp_sum(sum);
  if (lane == 0 && sum != 0.f) atomicAdd(lens + f / frames_per_utt, 1);
}
}  // namespace ssasr

extern "C" {
// src: [B, T, F] float64 (src_is_f64) or float32, device; dst: [B, T, F] fp32 or NULL (count only); lens: int32 [B], overwritten
int ssasr_prepare_x(const void* src, int src_is_f64, long long B, long long T, int F, float* dst, int* lens, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B <= 0) return 0;
  SSASR_REQUIRE(T > 0 && F > 0 && src && lens, "prepare_x: bad arguments");
  SSASR_CHECK_CUDA(cudaMemsetAsync(lens, 0, sizeof(int) * B, st));
  const long long n = B * T;
  const unsigned blocks = (unsigned)((n + 7) / 8);
  ProfScope ps(F_PACK, st);
  if (src_is_f64) prepare_x_kernel<double><<<blocks, 256, 0, st>>>((const double*)src, dst, lens, n, T, F);
  else prepare_x_kernel<float><<<blocks, 256, 0, st>>>((const float*)src, dst, lens, n, T, F);
  SSASR_LAUNCH_CHECK();
  return 0;
}
}

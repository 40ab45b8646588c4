// Launch counters and optional CUDA-event timing per kernel family.
#include "common.cuh"
#include "ssasr.h"
#include <mutex>
#include <vector>

namespace ssasr {

static std::mutex g_mu;
static long long g_launches[F_COUNT] = {0};
static bool g_prof = false;
struct Pair { int fam; cudaEvent_t e0, e1; };
static std::vector<Pair> g_pairs;

ProfScope::ProfScope(int family, cudaStream_t stream) : fam(family), st(stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_launches[fam]++;
  if (g_prof) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
  }
}
ProfScope::~ProfScope() {
  if (e0) {
    cudaEventRecord(e1, st);
    std::lock_guard<std::mutex> lk(g_mu);
    g_pairs.push_back({fam, e0, e1});
  }
}

}  // namespace ssasr

using namespace ssasr;

extern "C" {

static const char* kFamilyNames[F_COUNT] = {"gemm_f32", "rec_fwd_f32", "rec_bwd_f32", "attn_fwd", "attn_bwd", "pointwise",
                                            "ce_loss", "fbank", "pack", "gemm_tc", "rec_fwd_tc", "rec_bwd_tc", "optim", "spell_fwd", "spell_bwd"};

int ssasr_abi_version(void) { return SSASR_ABI_VERSION; }
int ssasr_num_families(void) { return F_COUNT; }
const char* ssasr_family_name(int i) { return (i >= 0 && i < F_COUNT) ? kFamilyNames[i] : ""; }

// strided host -> device copy on `stream` (rows of `width` bytes out of pitched buffers; asynchronous when the host buffer is
// pinned): the upload of a Listener group that is shallower than the padded host batch (ASR.decode_batch with a host tensor)
int ssasr_memcpy2d_h2d(void* dst, long long dpitch, const void* src, long long spitch, long long width, long long height, void* stream) {
  if (width <= 0 || height <= 0) return 0;
  SSASR_CHECK_CUDA(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height, cudaMemcpyHostToDevice,
                                     (cudaStream_t)stream));
  return 0;
}

// total kernel launches issued by this library since the last reset
long long ssasr_launch_count(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  long long n = 0;
  for (int i = 0; i < F_COUNT; ++i) n += g_launches[i];
  return n;
}
void ssasr_launch_count_reset(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < F_COUNT; ++i) g_launches[i] = 0;
}
// enable != 0: every subsequent launch is bracketed by CUDA events on its own stream
void ssasr_profile_enable(int enable) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_prof = enable != 0;
}
// Synchronises the device, sums elapsed ms and launch counts per family since the last read, clears the records.
int ssasr_profile_read(double* ms /*[F_COUNT]*/, long long* launches /*[F_COUNT]*/) {
  SSASR_CHECK_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < F_COUNT; ++i) { ms[i] = 0; launches[i] = 0; }
  for (auto& p : g_pairs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, p.e0, p.e1);
    ms[p.fam] += t;
    launches[p.fam]++;
    cudaEventDestroy(p.e0);
    cudaEventDestroy(p.e1);
  }
  g_pairs.clear();
  return 0;
}
}

// extern "C" surface of libdynamo_b200.so (see include/dynamo_b200.h).
#include <stdarg.h>

#include <atomic>
#include <string.h>

#include "dd_common.cuh"

namespace dd {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launches(int n) { g_launches += n; }
long long launches() { return g_launches.load(); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int validate_desc(const dd_warp_desc* d) {
  DD_REQUIRE(d != nullptr, "dd_warp_desc is NULL");
  DD_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0, "bad shape B=%d H=%d W=%d", d->B, d->H, d->W);
  DD_REQUIRE(d->H % 32 == 0 && d->W % 32 == 0, "H=%d, W=%d must be multiples of 32 (Trainer.py:25-26)", d->H, d->W);
  DD_REQUIRE(d->num_scales >= 1 && d->num_scales <= DD_MAX_SCALES, "num_scales=%d out of range", d->num_scales);
  DD_REQUIRE(d->num_frames >= 1 && d->num_frames <= DD_MAX_FRAMES, "num_frames=%d out of range", d->num_frames);
  DD_REQUIRE(d->min_depth > 0.f && d->max_depth > d->min_depth, "bad depth range [%g, %g]", d->min_depth, d->max_depth);
  DD_REQUIRE(d->target && d->K && d->inv_K, "target / K / inv_K must not be NULL");
  DD_REQUIRE(!((d->flags & DD_FLAG_MOTMASK) && !(d->flags & DD_FLAG_CMPFLOW)), "MOTMASK requires CMPFLOW (Trainer.py:466-490)");
  for (int f = 0; f < d->num_frames; ++f) DD_REQUIRE(d->source[f] && d->T[f], "source[%d] / T[%d] is NULL", f, f);
  for (int s = 0; s < d->num_scales; ++s) {
    DD_REQUIRE(d->scale[s] >= 0 && d->scale[s] <= 3, "scale[%d]=%d unsupported (0..3)", s, d->scale[s]);
    DD_REQUIRE(d->disp[s] != nullptr, "disp[%d] is NULL", s);
    for (int f = 0; f < d->num_frames; ++f) {
      if (d->flags & DD_FLAG_CMPFLOW) DD_REQUIRE(d->flow[s][f] != nullptr, "flow[%d][%d] is NULL", s, f);
      if (d->flags & DD_FLAG_MOTMASK) DD_REQUIRE(d->mask[s][f] != nullptr, "mask[%d][%d] is NULL", s, f);
    }
  }
  return DD_OK;
}

int warp_photo_fwd_impl(const dd_warp_desc*, const dd_warp_aux*, float*, void*, size_t, cudaStream_t);
int warp_photo_bwd_impl(const dd_warp_desc*, const float*, const dd_warp_grads*, const dd_warp_aux*, void*, size_t, cudaStream_t);

}  // namespace dd

extern "C" {

const char* dd_last_error(void) { return dd::g_err; }

int dd_version(void) { return 100; }

long long dd_launch_count(void) { return dd::launches(); }

int dd_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DD_ERR_CUDA;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return DD_ERR_CUDA;
  return n;
}

size_t dd_warp_photo_workspace_bytes(const dd_warp_desc* d) {
  if (!d || d->B <= 0 || d->H <= 0 || d->W <= 0) return 0;
  const size_t ctas = (size_t)(d->W / 32) * (d->H / 32) * d->B;
  // forward: [num_scales*DD_NSUM][ctas] partial sums; backward: [ctas][frames][12] pose partials
  const size_t fwd = (size_t)DD_MAX_SCALES * DD_NSUM * ctas * sizeof(float);
  const size_t bwd = 2 * ctas * DD_MAX_FRAMES * 12 * sizeof(float);   // backward tiles are 32x16
  return fwd > bwd ? fwd : bwd;
}

int dd_warp_photo_fwd(const dd_warp_desc* desc, const dd_warp_aux* aux, float* sums, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return dd::warp_photo_fwd_impl(desc, aux, sums, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dd_warp_photo_bwd(const dd_warp_desc* desc, const float* grad_sums, const dd_warp_aux* saved,
                      const dd_warp_grads* grads, void* workspace, size_t workspace_bytes, void* stream) {
  return dd::warp_photo_bwd_impl(desc, grad_sums, grads, saved, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"

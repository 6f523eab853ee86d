// extern "C" surface of libdynamo_b200.so (see include/dynamo_b200.h).
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <tuple>

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

// Texture objects over source tensors, cached by (device, pointer, shape): a texture object is only a descriptor of an
// address range (no device memory is allocated, nothing is copied), so an entry stays valid for as long as a tensor of that
// shape lives at that address.  Returns 0 (kernels then gather with plain loads) when the tensor does not meet the pitch-linear
// requirements or DD_NO_TEX is set.
cudaTextureObject_t source_texture(const float* ptr, int B, int H, int W) {
  static const bool off = getenv("DD_NO_TEX") != nullptr;
  const long long rows = (long long)B * 3 * H;
  if (off || ptr == nullptr || ((uintptr_t)ptr & 511u) != 0 || (W * 4) % 32 != 0 || rows > 65000 || W > 65000) return 0;
  struct Key {
    int dev;
    const float* ptr;
    int W;
    long long rows;
    bool operator<(const Key& o) const { return std::tie(dev, ptr, W, rows) < std::tie(o.dev, o.ptr, o.W, o.rows); }
  };
  static std::map<Key, cudaTextureObject_t> cache;
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  std::lock_guard<std::mutex> lock(mu);
  const Key key{dev, ptr, W, rows};
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  if (cache.size() >= 256) {   // bounded: drop everything (descriptors only) and start over
    for (auto& kv : cache) cudaDestroyTextureObject(kv.second);
    cache.clear();
  }
  cudaResourceDesc res;
  memset(&res, 0, sizeof(res));
  res.resType = cudaResourceTypePitch2D;
  res.res.pitch2D.devPtr = const_cast<float*>(ptr);
  res.res.pitch2D.desc = cudaCreateChannelDesc<float>();
  res.res.pitch2D.width = (size_t)W;
  res.res.pitch2D.height = (size_t)rows;
  res.res.pitch2D.pitchInBytes = (size_t)W * sizeof(float);
  cudaTextureDesc td;
  memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  cudaTextureObject_t tex = 0;
  if (cudaCreateTextureObject(&tex, &res, &td, nullptr) != cudaSuccess) {
    cudaGetLastError();   // not an error of the call: fall back to plain loads
    tex = 0;
  }
  cache[key] = tex;
  return tex;
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

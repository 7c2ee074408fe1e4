// pixie_cuda.so runtime: device/stream state, image handles, copies, fill, checksum, timers.
// Mirrors newImage / copy / fill of the reference (common.nim:39-54, pixie.nim:120-131) for
// device-resident canvases; errors follow the bindings' lastError convention
// (bindings/bindings.nim:3-10).
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace pixie {

static thread_local std::string g_err;

Runtime& rt() {
  static Runtime r;
  return r;
}
std::recursive_mutex& api_mutex() {
  static std::recursive_mutex m;
  return m;
}
void set_error(const std::string& msg) { g_err = msg; }
int fail_pixie(const std::string& msg) {
  g_err = msg;
  return 1;
}
int fail_cuda(cudaError_t e, const char* what) {
  cudaGetLastError();  // the error is reported here: do not let it surface again at the next launch check
  g_err = std::string("CUDA error: ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ") at " + what;
  return 2;
}

ProfScope::ProfScope(int s) : slot(s) {
  Runtime& r = rt();
  if (r.profiling) cudaEventRecord(r.prof[slot][0], r.stream);
}
ProfScope::~ProfScope() {
  Runtime& r = rt();
  if (r.profiling) cudaEventRecord(r.prof[slot][1], r.stream);
}

int ensure_init() {
  Runtime& r = rt();
  if (r.inited) return 0;
  return pixie_cuda_init(0);
}

Image* find_image(uint64_t h) {
  Runtime& r = rt();
  auto it = r.images.find(h);
  if (it == r.images.end()) {
    set_error("invalid image handle");
    return nullptr;
  }
  return &it->second;
}

int get_scratch(int slot, size_t bytes, void** out) {
  Runtime& r = rt();
  if (r.scratch_bytes[slot] < bytes) {
    if (r.scratch[slot]) {
      PX_CUDA(cudaStreamSynchronize(r.stream));
      PX_CUDA(cudaFree(r.scratch[slot]));
      r.scratch[slot] = nullptr;
      r.scratch_bytes[slot] = 0;
    }
    size_t want = bytes + bytes / 8 + 4096;
    PX_CUDA(cudaMalloc(&r.scratch[slot], want));
    r.scratch_bytes[slot] = want;
  }
  *out = r.scratch[slot];
  return 0;
}

int get_pinned(size_t bytes, void** out) {
  Runtime& r = rt();
  if (r.pinned_bytes < bytes) {
    if (r.pinned) {
      PX_CUDA(cudaStreamSynchronize(r.stream));
      PX_CUDA(cudaFreeHost(r.pinned));
      r.pinned = nullptr;
      r.pinned_bytes = 0;
    }
    size_t want = bytes + bytes / 8 + 4096;
    PX_CUDA(cudaMallocHost(&r.pinned, want));
    r.pinned_bytes = want;
  }
  *out = r.pinned;
  return 0;
}

// Pinned staging buffer shared by every host->device parameter copy.  acquire() waits until the
// previous copy that read it has completed; release() marks the copy just enqueued.
int staging_acquire(size_t bytes, void** out) {
  Runtime& r = rt();
  if (r.staging_busy) {
    PX_CUDA(cudaEventSynchronize(r.staging_done));
    r.staging_busy = false;
  }
  return get_pinned(bytes, out);
}
int staging_release() {
  Runtime& r = rt();
  if (!r.staging_done) PX_CUDA(cudaEventCreateWithFlags(&r.staging_done, cudaEventDisableTiming));
  PX_CUDA(cudaEventRecord(r.staging_done, r.stream));
  r.staging_busy = true;
  return 0;
}

__global__ void fill_kernel(uint4* __restrict__ p, size_t n16, uint32_t v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint4 q = make_uint4(v, v, v, v);
  for (; i < n16; i += stride) p[i] = q;
}
__global__ void fill_tail_kernel(uint32_t* __restrict__ p, size_t n, uint32_t v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void checksum_kernel(const uint32_t* __restrict__ p, size_t n, unsigned long long* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (; i < n; i += stride) acc += (unsigned long long)p[i] * (unsigned long long)(i | 1);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

static int new_image(int w, int h, int layers, int bpp, void* wrap, pixie_image_t* out, bool zero = true) {
  if (int rc = ensure_init()) return rc;
  if (w <= 0 || h <= 0) return fail_pixie("Image width and height must be > 0");  // common.nim:41-42
  if (layers <= 0) return fail_pixie("Image layers must be > 0");
  if (bpp != 4 && bpp != 1) return fail_pixie("bytes_per_pixel must be 4 (RGBX) or 1 (A8)");
  Runtime& r = rt();
  Image im;
  im.w = w;
  im.h = h;
  im.layers = layers;
  im.bpp = bpp;
  if (wrap) {
    im.data = (uint8_t*)wrap;
    im.owned = false;
  } else {
    // stream-ordered allocation from the device pool (its memory stays cached: no driver call per newImage)
    PX_CUDA(cudaMallocAsync(&im.data, im.bytes(), r.stream));
    if (zero) PX_CUDA(cudaMemsetAsync(im.data, 0, im.bytes(), r.stream));
  }
  std::lock_guard<std::mutex> lk(r.mu);
  uint64_t hd = r.next_handle++;
  r.images[hd] = im;
  *out = hd;
  return 0;
}

int new_image_uninit(int w, int h, int layers, int bpp, pixie_image_t* out) {
  return new_image(w, h, layers, bpp, nullptr, out, false);
}

}  // namespace pixie

using namespace pixie;

extern "C" {

int pixie_cuda_init(int device) {
  PX_API_GUARD;
  Runtime& r = rt();
  if (r.inited && r.device == device) return 0;
  // the stream, events, scratch, pinned staging and every image handle belong to the first device: a process drives
  // ONE GPU (one process per GPU is how the path shards)
  if (r.inited)
    return fail_pixie("pixie_cuda_init: already initialised on device " + std::to_string(r.device) +
                      "; one process drives one GPU (start one process per GPU)");
  int n = 0;
  PX_CUDA(cudaGetDeviceCount(&n));
  if (n <= 0) return fail_pixie("pixie_cuda: no CUDA device visible (there is no CPU fallback)");
  if (device < 0 || device >= n) return fail_pixie("pixie_cuda: invalid device ordinal");
  PX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PX_CUDA(cudaGetDeviceProperties(&prop, device));
  r.device = device;
  r.num_sms = prop.multiProcessorCount;
  if (!r.own_stream) PX_CUDA(cudaStreamCreateWithFlags(&r.own_stream, cudaStreamNonBlocking));
  r.stream = r.own_stream;
  {  // temporaries of draw() (minify / magnify chain) come from the stream-ordered pool: keep its memory cached
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (!r.ev0) PX_CUDA(cudaEventCreate(&r.ev0));
  if (!r.ev1) PX_CUDA(cudaEventCreate(&r.ev1));
  r.inited = true;
  return 0;
}

const char* pixie_cuda_last_error(void) { return g_err.c_str(); }

int pixie_cuda_set_stream(void* s) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  Runtime& r = rt();
  cudaStream_t next = s ? (cudaStream_t)s : r.own_stream;
  if (next == r.stream) return 0;
  // everything already queued on the old stream (image zero-fills, staged H2D copies, LUT uploads to __constant__
  // memory) happens before anything issued on the new one
  if (!r.switch_ev) PX_CUDA(cudaEventCreateWithFlags(&r.switch_ev, cudaEventDisableTiming));
  PX_CUDA(cudaEventRecord(r.switch_ev, r.stream));
  PX_CUDA(cudaStreamWaitEvent(next, r.switch_ev, 0));
  r.stream = next;
  return 0;
}

int pixie_cuda_set_sm_reserve(int sms) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  Runtime& r = rt();
  r.sm_reserve = std::max(0, std::min(sms, r.num_sms - 1));
  return 0;
}

int pixie_cuda_sync(void) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  PX_CUDA(cudaStreamSynchronize(rt().stream));
  return 0;
}

int pixie_cuda_device_count(int* out) {
  PX_API_GUARD;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *out = 0;
    return fail_cuda(e, "cudaGetDeviceCount");
  }
  *out = n;
  return 0;
}

int pixie_cuda_image_create(int w, int h, pixie_image_t* out) { PX_API_GUARD; return new_image(w, h, 1, 4, nullptr, out); }
int pixie_cuda_image_create_layers(int w, int h, int layers, pixie_image_t* out) {
  PX_API_GUARD;
  return new_image(w, h, layers, 4, nullptr, out);
}
int pixie_cuda_image_create_a8(int w, int h, pixie_image_t* out) { PX_API_GUARD; return new_image(w, h, 1, 1, nullptr, out); }
int pixie_cuda_image_wrap(void* p, int w, int h, int layers, int bpp, pixie_image_t* out) {
  PX_API_GUARD;
  if (!p) return fail_pixie("pixie_cuda_image_wrap: null device pointer");
  return new_image(w, h, layers, bpp, p, out);
}

int pixie_cuda_image_destroy(pixie_image_t h) {
  PX_API_GUARD;
  Runtime& r = rt();
  std::lock_guard<std::mutex> lk(r.mu);
  auto it = r.images.find(h);
  if (it == r.images.end()) return fail_pixie("invalid image handle");
  if (it->second.owned) {
    // the memory goes back to the pool once everything issued so far has run; with a caller-provided
    // stream other streams may still be using the image, so wait for the current one first
    if (r.stream != r.own_stream) PX_CUDA(cudaStreamSynchronize(r.stream));
    PX_CUDA(cudaFreeAsync(it->second.data, r.stream));
  }
  r.images.erase(it);
  return 0;
}

int pixie_cuda_image_info(pixie_image_t h, int* w, int* ht, int* layers, int* bpp, void** ptr) {
  PX_API_GUARD;
  Image* im = find_image(h);
  if (!im) return 1;
  if (w) *w = im->w;
  if (ht) *ht = im->h;
  if (layers) *layers = im->layers;
  if (bpp) *bpp = im->bpp;
  if (ptr) *ptr = im->data;
  return 0;
}

int pixie_cuda_image_upload(pixie_image_t h, const uint8_t* host) {
  PX_API_GUARD;
  Image* im = find_image(h);
  if (!im) return 1;
  PX_CUDA(cudaMemcpyAsync(im->data, host, im->bytes(), cudaMemcpyHostToDevice, rt().stream));
  PX_CUDA(cudaStreamSynchronize(rt().stream));  // pageable host memory may not be kept after return
  return 0;
}
int pixie_cuda_image_download(pixie_image_t h, uint8_t* host) {
  PX_API_GUARD;
  Image* im = find_image(h);
  if (!im) return 1;
  PX_CUDA(cudaMemcpyAsync(host, im->data, im->bytes(), cudaMemcpyDeviceToHost, rt().stream));
  PX_CUDA(cudaStreamSynchronize(rt().stream));
  return 0;
}
int pixie_cuda_image_upload_async(pixie_image_t h, const uint8_t* host) {
  PX_API_GUARD;
  Image* im = find_image(h);
  if (!im) return 1;
  PX_CUDA(cudaMemcpyAsync(im->data, host, im->bytes(), cudaMemcpyHostToDevice, rt().stream));
  return 0;
}
int pixie_cuda_image_download_async(pixie_image_t h, uint8_t* host) {
  PX_API_GUARD;
  Image* im = find_image(h);
  if (!im) return 1;
  PX_CUDA(cudaMemcpyAsync(host, im->data, im->bytes(), cudaMemcpyDeviceToHost, rt().stream));
  return 0;
}
int pixie_cuda_image_download_rows(pixie_image_t h, int layer, int y0, int y1, uint8_t* host) {
  PX_API_GUARD;
  Image* im = find_image(h);
  if (!im) return 1;
  if (layer < 0 || layer >= im->layers || y0 < 0 || y1 > im->h || y0 > y1) return fail_pixie("row range out of bounds");
  const uint8_t* src = im->data + (size_t)layer * im->layer_bytes() + (size_t)y0 * im->w * im->bpp;
  PX_CUDA(cudaMemcpyAsync(host, src, (size_t)(y1 - y0) * im->w * im->bpp, cudaMemcpyDeviceToHost, rt().stream));
  PX_CUDA(cudaStreamSynchronize(rt().stream));
  return 0;
}

int pixie_cuda_image_fill(pixie_image_t h, uint32_t rgbx) {
  PX_API_GUARD;
  Image* im = find_image(h);
  if (!im) return 1;
  Runtime& r = rt();
  if (im->bpp == 1) {
    PX_CUDA(cudaMemsetAsync(im->data, (int)(rgbx >> 24), im->bytes(), r.stream));
    return 0;
  }
  uint8_t b0 = rgbx & 255;
  if (((rgbx >> 8) & 255) == b0 && ((rgbx >> 16) & 255) == b0 && (rgbx >> 24) == b0) {
    PX_CUDA(cudaMemsetAsync(im->data, b0, im->bytes(), r.stream));  // internal.nim:61-63
    return 0;
  }
  size_t npx = im->bytes() / 4;
  size_t n16 = npx / 4;
  if (n16) {
    int blocks = (int)std::min<size_t>((n16 + 255) / 256, (size_t)r.num_sms * 16);
    fill_kernel<<<blocks, 256, 0, r.stream>>>((uint4*)im->data, n16, rgbx);
    PX_LAUNCHED();
  }
  size_t tail = npx - n16 * 4;
  if (tail) {
    fill_tail_kernel<<<1, 32, 0, r.stream>>>((uint32_t*)im->data + n16 * 4, tail, rgbx);
    PX_LAUNCHED();
  }
  return 0;
}

int pixie_cuda_image_copy(pixie_image_t dst, pixie_image_t src) {
  PX_API_GUARD;
  Image* d = find_image(dst);
  Image* s = find_image(src);
  if (!d || !s) return 1;
  if (d->w != s->w || d->h != s->h || d->layers != s->layers || d->bpp != s->bpp)
    return fail_pixie("pixie_cuda_image_copy: shape mismatch");
  PX_CUDA(cudaMemcpyAsync(d->data, s->data, s->bytes(), cudaMemcpyDeviceToDevice, rt().stream));
  return 0;
}

int pixie_cuda_image_checksum(pixie_image_t h, uint64_t* out) {
  PX_API_GUARD;
  Image* im = find_image(h);
  if (!im) return 1;
  if (im->bpp != 4) return fail_pixie("checksum needs an RGBX image");
  Runtime& r = rt();
  void* d;
  if (int rc = get_scratch(3, 8, &d)) return rc;
  PX_CUDA(cudaMemsetAsync(d, 0, 8, r.stream));
  size_t n = im->bytes() / 4;
  int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)r.num_sms * 16);
  checksum_kernel<<<blocks, 256, 0, r.stream>>>((const uint32_t*)im->data, n, (unsigned long long*)d);
  PX_LAUNCHED();
  PX_CUDA(cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, r.stream));
  PX_CUDA(cudaStreamSynchronize(r.stream));
  return 0;
}

int pixie_cuda_host_alloc(size_t bytes, void** out) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  PX_CUDA(cudaMallocHost(out, bytes));
  return 0;
}
int pixie_cuda_host_free(void* p) {
  PX_API_GUARD;
  PX_CUDA(cudaFreeHost(p));
  return 0;
}

int pixie_cuda_set_profiling(int enabled) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  Runtime& r = rt();
  if (enabled && !r.prof[0][0])
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 2; j++) PX_CUDA(cudaEventCreate(&r.prof[i][j]));
  if (enabled && !r.band_prof[0][0])
    for (int i = 0; i < Runtime::kBands; i++)
      for (int j = 0; j < 4; j++) PX_CUDA(cudaEventCreate(&r.band_prof[i][j]));
  r.profiling = enabled != 0;
  return 0;
}
int pixie_cuda_profile_read(int slot, float* ms) {
  PX_API_GUARD;
  Runtime& r = rt();
  if (slot < 0 || slot >= 8 || !r.prof[0][0]) return fail_pixie("profiling slot out of range or profiling never enabled");
  if (r.prof_bands > 0 && (slot == kProfPlan || slot == kProfRaster)) {
    // banded run: the sum over the bands' launches (they overlap other bands' kernels, so this is an upper bound of
    // the time the kernel would take alone)
    const int k0 = slot == kProfPlan ? 0 : 2;
    float sum = 0.0f;
    for (int b = 0; b < r.prof_bands; b++) {
      if (cudaEventQuery(r.band_prof[b][k0 + 1]) == cudaErrorInvalidResourceHandle) continue;
      float t = 0.0f;
      PX_CUDA(cudaEventSynchronize(r.band_prof[b][k0 + 1]));
      PX_CUDA(cudaEventElapsedTime(&t, r.band_prof[b][k0], r.band_prof[b][k0 + 1]));
      sum += t;
    }
    *ms = sum;
    return 0;
  }
  PX_CUDA(cudaEventSynchronize(r.prof[slot][1]));
  PX_CUDA(cudaEventElapsedTime(ms, r.prof[slot][0], r.prof[slot][1]));
  return 0;
}

int pixie_cuda_launch_count(uint64_t* out) {
  PX_API_GUARD;
  *out = rt().launches;
  return 0;
}
int pixie_cuda_timer_begin(void) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  PX_CUDA(cudaEventRecord(rt().ev0, rt().stream));
  return 0;
}
int pixie_cuda_timer_end(float* ms) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  PX_CUDA(cudaEventRecord(rt().ev1, rt().stream));
  PX_CUDA(cudaEventSynchronize(rt().ev1));
  PX_CUDA(cudaEventElapsedTime(ms, rt().ev0, rt().ev1));
  return 0;
}

}  // extern "C"

// Strict drop-in variants of the C ABI: host pixels in, host pixels out (upload -> run -> download),
// for callers whose Image.data is a host seq (treeform/pixie src/pixie/common.nim:34-37).
// Device images that are about to be overwritten by an upload (or entirely by the operation) are created
// without the zero fill of newImage.
#include "common.cuh"

using namespace pixie;

namespace {
struct TmpImage {
  pixie_image_t h = 0;
  ~TmpImage() {
    if (h) pixie_cuda_image_destroy(h);
  }
};
}  // namespace

extern "C" {

int pixie_cuda_fill_segments_host(uint8_t* pixels, int w, int h, const float* seg, const int16_t* wind, int n,
                                  uint32_t rgbx, int rule, int mode) {
  PX_API_GUARD;
  TmpImage im;
  if (int rc = new_image_uninit(w, h, 1, 4, &im.h)) return rc;
  if (int rc = pixie_cuda_image_upload(im.h, pixels)) return rc;
  if (int rc = pixie_cuda_fill_segments(im.h, seg, wind, n, rgbx, rule, mode)) return rc;
  return pixie_cuda_image_download(im.h, pixels);
}

int pixie_cuda_blend_rect_host(uint8_t* dst, int dw, int dh, const uint8_t* src, int sw, int sh, int px, int py,
                               int mode) {
  PX_API_GUARD;
  TmpImage d, s;
  if (int rc = new_image_uninit(dw, dh, 1, 4, &d.h)) return rc;
  if (int rc = new_image_uninit(sw, sh, 1, 4, &s.h)) return rc;
  if (int rc = pixie_cuda_image_upload(d.h, dst)) return rc;
  if (int rc = pixie_cuda_image_upload(s.h, src)) return rc;
  if (int rc = pixie_cuda_blend_rect(d.h, s.h, px, py, mode)) return rc;
  return pixie_cuda_image_download(d.h, dst);
}

int pixie_cuda_blur_host(uint8_t* pixels, int w, int h, const uint16_t* lut, int radius, uint32_t oob) {
  PX_API_GUARD;
  if (radius == 0) return 0;
  if (radius < 0) return fail_pixie("Cannot apply negative blur");
  TmpImage im;
  if (int rc = new_image_uninit(w, h, 1, 4, &im.h)) return rc;
  if (int rc = pixie_cuda_image_upload(im.h, pixels)) return rc;
  if (int rc = pixie_cuda_blur(im.h, lut, radius, oob)) return rc;
  return pixie_cuda_image_download(im.h, pixels);
}

int pixie_cuda_shadow_host(const uint8_t* src, uint8_t* dst, int w, int h, float ox, float oy, int spread,
                           const uint16_t* lut, int radius, uint32_t rgbx) {
  PX_API_GUARD;
  TmpImage s, d;
  if (int rc = new_image_uninit(w, h, 1, 4, &s.h)) return rc;
  if (int rc = new_image_uninit(w, h, 1, 4, &d.h)) return rc;
  if (int rc = pixie_cuda_image_upload(s.h, src)) return rc;
  if (int rc = pixie_cuda_shadow(s.h, d.h, ox, oy, spread, lut, radius, rgbx)) return rc;
  return pixie_cuda_image_download(d.h, dst);
}

int pixie_cuda_spread_host(uint8_t* pixels, int w, int h, int spread) {  // images.nim:700-758
  PX_API_GUARD;
  if (spread == 0) return 0;
  TmpImage im;
  if (int rc = new_image_uninit(w, h, 1, 4, &im.h)) return rc;
  if (int rc = pixie_cuda_image_upload(im.h, pixels)) return rc;
  if (int rc = pixie_cuda_spread(im.h, spread)) return rc;
  return pixie_cuda_image_download(im.h, pixels);
}

int pixie_cuda_apply_opacity_host(uint8_t* pixels, int w, int h, float opacity) {  // images.nim:261-277
  PX_API_GUARD;
  TmpImage im;
  if (int rc = new_image_uninit(w, h, 1, 4, &im.h)) return rc;
  if (int rc = pixie_cuda_image_upload(im.h, pixels)) return rc;
  if (int rc = pixie_cuda_apply_opacity(im.h, opacity)) return rc;
  return pixie_cuda_image_download(im.h, pixels);
}

int pixie_cuda_blend_rect_masked_host(uint8_t* dst, int dw, int dh, const uint8_t* src, const uint8_t* mask,
                                      int mask_bytes_per_pixel, int sw, int sh, int px, int py, int mode) {
  PX_API_GUARD;
  if (mask_bytes_per_pixel != 1 && mask_bytes_per_pixel != 4) return fail_pixie("mask_bytes_per_pixel must be 1 (A8) or 4 (RGBX)");
  TmpImage d, s, m;
  if (int rc = new_image_uninit(dw, dh, 1, 4, &d.h)) return rc;
  if (int rc = new_image_uninit(sw, sh, 1, 4, &s.h)) return rc;
  if (int rc = new_image_uninit(sw, sh, 1, mask_bytes_per_pixel, &m.h)) return rc;
  if (int rc = pixie_cuda_image_upload(d.h, dst)) return rc;
  if (int rc = pixie_cuda_image_upload(s.h, src)) return rc;
  if (int rc = pixie_cuda_image_upload(m.h, mask)) return rc;
  if (int rc = pixie_cuda_blend_rect_masked(d.h, s.h, m.h, px, py, mode)) return rc;
  return pixie_cuda_image_download(d.h, dst);
}

int pixie_cuda_draw_host(uint8_t* dst, int dw, int dh, const uint8_t* src, int sw, int sh, const float* mat, int mode,
                         int tiled) {
  PX_API_GUARD;
  TmpImage d, s;
  if (int rc = new_image_uninit(dw, dh, 1, 4, &d.h)) return rc;
  if (int rc = new_image_uninit(sw, sh, 1, 4, &s.h)) return rc;
  if (int rc = pixie_cuda_image_upload(d.h, dst)) return rc;
  if (int rc = pixie_cuda_image_upload(s.h, src)) return rc;
  if (int rc = tiled ? pixie_cuda_draw_tiled(d.h, s.h, mat, mode) : pixie_cuda_draw(d.h, s.h, mat, mode)) return rc;
  return pixie_cuda_image_download(d.h, dst);
}

int pixie_cuda_fill_gradient_host(uint8_t* pixels, int w, int h, int kind, const float* handles, int n_handles,
                                  const float* stop_pos, const float* stop_rgba, int n_stops, float opacity) {
  PX_API_GUARD;
  TmpImage im;
  if (int rc = new_image_uninit(w, h, 1, 4, &im.h)) return rc;
  if (int rc = pixie_cuda_image_upload(im.h, pixels)) return rc;  // opacity 0 leaves the image as it is
  if (int rc = pixie_cuda_fill_gradient(im.h, kind, handles, n_handles, stop_pos, stop_rgba, n_stops, opacity)) return rc;
  return pixie_cuda_image_download(im.h, pixels);
}

int pixie_cuda_minify_by2_host(const uint8_t* src, int w, int h, int power, uint8_t* dst) {
  PX_API_GUARD;
  TmpImage s, d;
  if (int rc = new_image_uninit(w, h, 1, 4, &s.h)) return rc;
  if (int rc = pixie_cuda_image_upload(s.h, src)) return rc;
  if (int rc = pixie_cuda_minify_by2(s.h, power, &d.h)) return rc;
  return pixie_cuda_image_download(d.h, dst);
}

int pixie_cuda_magnify_by2_host(const uint8_t* src, int w, int h, int power, uint8_t* dst) {
  PX_API_GUARD;
  TmpImage s, d;
  if (int rc = new_image_uninit(w, h, 1, 4, &s.h)) return rc;
  if (int rc = pixie_cuda_image_upload(s.h, src)) return rc;
  if (int rc = pixie_cuda_magnify_by2(s.h, power, &d.h)) return rc;
  return pixie_cuda_image_download(d.h, dst);
}

}  // extern "C"

/*
 * pixie_oracle.h — CPU oracle for the raster hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (pixie_b200/, pixie_cuda.so) never links, imports or calls it.
 *
 * It is a plain C++ restatement (kind = "port": the Nim reference cannot be compiled in this
 * image, no nim/nimble and its dependencies are not vendored) of treeform/pixie:
 *   src/pixie/paths.nim   :1098-1117 computeBounds, :1127-1144 PartitionEntry, :1168-1262
 *                         partitionSegments, :1268-1348 fixed32/sortHits/walk, :1350-1431
 *                         computeCoverage, :1442-1591 fillCoverage/fillHits, :1593-1912 fillShapes
 *   src/pixie/blends.nim  :15-299 (all 20 BlendMode enumerators)
 *   src/pixie/common.nim  :67-101 ColorRGBX*float32, ColorRGBX*uint8, snapToPixels
 *   src/pixie/images.nim  :261-277 applyOpacity, :304-365 blur, :468-529 blendRect,
 *                         :700-758 spread, :760-776 shadow
 *                         :168-259 minifyBy2/magnifyBy2, :367-449 getRgbaSmooth/drawCorrect,
 *                         :531-683 drawSmooth/draw/drawTiled
 *   src/pixie/paints.nim  :68-248 gradientColor, fillGradient{Linear,Radial,Angular}
 *   src/pixie/simd/sse2.nim :6-46, :510-524 (the x86 numerics of the row kernels)
 *
 * Parity status: pinned against the reference's own golden PNGs (tests/golden/, see
 * tests/test_oracle_goldens.py) for Normal/Overwrite/Mask/ExcludeMask/Exclusion fills, strokes,
 * blur and the shadow pipeline.  PARITY UNPINNED for SoftLight/Hue/Saturation/Color/Luminosity
 * (delegated to the un-vendored chroma package, restated from the W3C compositing spec) and for
 * the un-premultiply step of ColorBurn/ColorDodge (chroma rgba(); Pixie's in-repo
 * straightAlphaTable, internal.nim:68-74, is used as the stand-in).
 *
 * sem: 0 = canonical semantics (x86 SSE2/AVX2 row-kernel bodies applied to every pixel,
 *          SURVEY.md section 2.3) — this is what the goldens pin and what the GPU must match;
 *      1 = the reference's scalar (-d:pixieNoSimd) roundings, for the "<= 1 LSB" report.
 */
#ifndef PIXIE_ORACLE_H
#define PIXIE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* orc_last_error(void);

/* fillShapes from the segment list down.  img: w*h*4 bytes premultiplied RGBX, modified in place.
 * covered_px (optional) += number of pixels this fill touched with non-zero coverage.
 * returns 0 ok, 1 = PixieError (message in orc_last_error). */
int orc_fill_segments(uint8_t* img, int w, int h, const float* seg_xyxy, const int16_t* winding, int n,
                      uint32_t rgbx, int winding_rule, int blend_mode, int sem, uint64_t* covered_px);

/* blender(mode)(backdrop, source) — scalar blend functions of blends.nim. */
uint32_t orc_blend_px(int blend_mode, uint32_t backdrop, uint32_t source);

/* blendRect (images.nim:468-529): dst(dw x dh) <- blend(dst, src(sw x sh) at (px,py)). */
int orc_blend_rect(uint8_t* dst, int dw, int dh, const uint8_t* src, int sw, int sh, int px, int py,
                   int blend_mode);
/* The non-solid-paint composite (paths.nim:2141-2142) in one step:
 * tmp = src; tmp.draw(mask, MaskBlend); dst.draw(tmp, mode).  mask is either an RGBX image
 * (mask_is_rgbx=1, alpha used) or an 8-bit coverage plane, same size as src. */
int orc_blend_rect_masked(uint8_t* dst, int dw, int dh, const uint8_t* src, const uint8_t* mask,
                          int mask_is_rgbx, int sw, int sh, int px, int py, int blend_mode);

int orc_apply_opacity(uint8_t* img, int w, int h, float opacity);
int orc_blur(uint8_t* img, int w, int h, const uint16_t* lut, int radius, uint32_t oob_rgbx);
int orc_spread(uint8_t* img, int w, int h, int spread);
/* shadow: out (w*h*4) = shadow(img, offset, spread, blur LUT, colour). */
int orc_shadow(const uint8_t* img, int w, int h, float ox, float oy, int spread, const uint16_t* lut,
               int radius, uint32_t rgbx, uint8_t* out);

/* minifyBy2 / magnifyBy2 (images.nim:168-259).  minify: out may be NULL to query the size. */
int orc_minify_by2(const uint8_t* src, int w, int h, int power, uint8_t* out, int* out_w, int* out_h);
int orc_magnify_by2(const uint8_t* src, int w, int h, int power, uint8_t* out);
/* draw (images.nim:636-678) with any transform: minify/magnify chain, drawSmooth (:531-634) or
 * blendRect.  mat = vmath Mat3 storage (column-major, 9 floats). */
int orc_draw(uint8_t* dst, int dw, int dh, const uint8_t* src, int sw, int sh, const float* mat, int blend_mode);
/* drawCorrect (images.nim:405-449); tiled != 0 is drawTiled (:680-683). */
int orc_draw_correct(uint8_t* dst, int dw, int dh, const uint8_t* src, int sw, int sh, const float* mat,
                     int blend_mode, int tiled);
/* fillGradient (paints.nim:68-248).  kind = ord(PaintKind): 3 linear, 4 radial, 5 angular. */
int orc_fill_gradient(uint8_t* img, int w, int h, int kind, const float* handles_xy, int n_handles,
                      const float* stop_pos, const float* stop_rgba, int n_stops, float opacity);

#ifdef __cplusplus
}
#endif
#endif

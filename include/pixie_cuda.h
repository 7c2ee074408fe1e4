/*
 * pixie_cuda.h — C ABI of pixie_cuda.so, the B200 (sm_100a) raster hot path behind Pixie's
 * public procs.  This is the drop-in boundary: a Nim shim ({.importc, dynlib: "pixie_cuda.so".},
 * see INTEGRATION.md and pixie_b200/nim/pixie_cuda.nim) replaces the BODIES of the procs cited
 * below and calls these entry points; signatures of the public procs do not change.
 *
 * Conventions
 *   - every function returns 0 on success; 1 = the condition under which the reference raises
 *     PixieError (src/pixie/common.nim:4) — the shim does
 *     `raise newException(PixieError, $pixie_cuda_last_error())`; 2 = CUDA runtime error (sticky,
 *     message verbatim).  There is NO CPU fallback.
 *   - images are premultiplied-alpha RGBX, 4 bytes/pixel, row-major, index = width*y + x
 *     (common.nim:34-37,56-57).  A8 images (1 byte/pixel) exist only as coverage masks.
 *   - enums are passed as int: ord(BlendMode) per common.nim:6-29 (Normal=0 .. Luminosity=15,
 *     Mask=16, Overwrite=17, SubtractMask=18, ExcludeMask=19); WindingRule NonZero=0, EvenOdd=1
 *     (paths.nim:5-8).
 *   - colours are packed ColorRGBX: r | g<<8 | b<<16 | a<<24, already premultiplied
 *     (what color.asRgbx() yields at paths.nim:1603).
 *   - the library never keeps a host pointer after a call returns.  Device work is issued on one
 *     stream (pixie_cuda_set_stream) in call order = the reference's sequential semantics; only
 *     *_download, *_host and pixie_cuda_sync block the host.
 *   - threading: the reference is single-threaded.  Here ONE coarse recursive lock serialises every
 *     entry point, so calls from several host threads — on distinct handles or on the same one — are safe; their
 *     device work is issued on the library's one stream in lock-acquisition order.  One process drives one GPU
 *     (pixie_cuda_init on a second device ordinal is an error): several GPUs = one process per GPU.
 *   - numerics: bit-exact with the reference's x86 row kernels applied to every pixel
 *     (src/pixie/simd/sse2.nim:6-46,510-524); float32 geometry without FMA.
 */
#ifndef PIXIE_CUDA_H
#define PIXIE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t pixie_image_t;   /* opaque handle; 0 is never valid */
typedef uint64_t pixie_cmdlist_t; /* opaque handle of a device-resident fill command list */

/* ---- runtime ------------------------------------------------------------------------- */
int pixie_cuda_init(int device);               /* select device, create the stream; idempotent */
const char* pixie_cuda_last_error(void);       /* bindings/bindings.nim:3-10 takeError analogue */
int pixie_cuda_set_stream(void* cuda_stream);  /* cudaStream_t to issue on (NULL = library stream; the legacy default
                                                * stream is named by cudaStreamLegacy = (cudaStream_t)0x1) */
int pixie_cuda_sync(void);                     /* wait for everything issued so far */
/* Persistent kernels that fill the GPU with one CTA per SM (the fused blur) leave `sms` SMs free, so that a
 * collective's kernels on another stream (NCCL send / recv of the halo rows) can run beside them.  Default 0. */
int pixie_cuda_set_sm_reserve(int sms);
int pixie_cuda_device_count(int* out);

/* ---- images: newImage / copy / fill (common.nim:39-54, pixie.nim:120-131) -------------- */
int pixie_cuda_image_create(int width, int height, pixie_image_t* out);       /* RGBX, zeroed */
int pixie_cuda_image_create_layers(int width, int height, int layers, pixie_image_t* out);
                                      /* `layers` independent RGBX canvases in one allocation */
int pixie_cuda_image_create_a8(int width, int height, pixie_image_t* out);    /* coverage plane */
int pixie_cuda_image_wrap(void* device_ptr, int width, int height, int layers, int bytes_per_pixel,
                          pixie_image_t* out); /* borrow caller-owned device memory (not freed) */
int pixie_cuda_image_destroy(pixie_image_t image);
int pixie_cuda_image_info(pixie_image_t image, int* width, int* height, int* layers, int* bytes_per_pixel,
                          void** device_ptr);
int pixie_cuda_image_upload(pixie_image_t image, const uint8_t* host_pixels);  /* all layers */
int pixie_cuda_image_download(pixie_image_t image, uint8_t* host_pixels);      /* all layers; blocks */
int pixie_cuda_image_upload_async(pixie_image_t image, const uint8_t* pinned_host_pixels);
int pixie_cuda_image_download_async(pixie_image_t image, uint8_t* pinned_host_pixels);
int pixie_cuda_image_download_rows(pixie_image_t image, int layer, int y0, int y1, uint8_t* host_rows);
int pixie_cuda_image_fill(pixie_image_t image, uint32_t rgbx);                 /* image.fill(color) */
int pixie_cuda_image_copy(pixie_image_t dst, pixie_image_t src);               /* same shape */
/* sum over pixels of (pixel_u32 * (index|1)) mod 2^64 — device-side checksum for batch tests */
int pixie_cuda_image_checksum(pixie_image_t image, uint64_t* out);

/* ---- fillShapes (paths.nim:1593-1912) ------------------------------------------------------
 * Replaces the body of the private proc fillShapes() called by fillPath (:2112) and strokePath
 * (:2175) from the point where shapesToSegments (:1604) has produced the segment list:
 *   seg_xyxy  n x 4 float32 {at.x, at.y, to.x, to.y}, y quantised to 1/256, at.y < to.y, no
 *             horizontals (exactly the output of shapesToSegments :1059-1090)
 *   winding   n int16 (+1 / -1)
 * Device kernels: partition/bin (partitionSegments :1168-1262), per-scanline mode decision +
 * trapezoid / 5-sample coverage (:1631-1908, computeCoverage :1350-1431) fused with the blend
 * (fillCoverage / fillHits :1479-1591, blends.nim).
 * Returns 1 with "Path int overflow detected" where the reference raises it (:1618-1619). */
int pixie_cuda_fill_segments(pixie_image_t image, const float* seg_xyxy, const int16_t* winding, int n,
                             uint32_t rgbx, int winding_rule, int blend_mode);

/* Ordered command list: fill k uses segments [seg_offsets[k], seg_offsets[k+1]) and goes to layer
 * layer_of_fill[k] (NULL = layer 0; must be non-decreasing).  Fills of one layer are applied in
 * list order (svg.nim:563-604 newImage(svg), fonts.nim:566-596); layers are independent.
 * covered_px (optional, may be NULL) receives the number of pixels touched with non-zero coverage
 * summed over fills — the work unit of the Mpixel/s metric (blocks the host when non-NULL).
 * Lists of up to 8192 segments (a fillPath, a glyph, an icon) are sized on the host and only enqueue work;
 * larger lists wait once inside the call for a 24-byte device readback (the sizes of the band arrays). */
int pixie_cuda_fill_batch(pixie_image_t image, int num_fills, const int32_t* layer_of_fill,
                          const float* seg_xyxy, const int16_t* winding, const int32_t* seg_offsets,
                          const uint32_t* rgbx, const uint8_t* winding_rule, const uint8_t* blend_mode,
                          uint64_t* covered_px);

/* One-shot newImage(svg)-style render with host pixels out: a width x height canvas (transparent when
 * clear != 0, else initialised from `pixels`), the ordered fills, and the result written to `pixels`.
 * The canvas is rasterised in row bands on concurrent streams and each band is copied out as soon as it
 * is done, so the D2H copy overlaps the rendering when `pixels` is page-locked (pixie_cuda_host_alloc).
 * Blocks until `pixels` is complete. */
int pixie_cuda_render_batch_host(uint8_t* pixels, int width, int height, int clear, int num_fills,
                                 const float* seg_xyxy, const int16_t* winding, const int32_t* seg_offsets,
                                 const uint32_t* rgbx, const uint8_t* winding_rule, const uint8_t* blend_mode,
                                 uint64_t* covered_px);

/* Same, split so the inputs can stay resident in HBM: create uploads segments + per-fill headers
 * for canvases of (width, height, layers); run rasterises them into `image`. */
int pixie_cuda_cmdlist_create(int width, int height, int layers, int num_fills, const int32_t* layer_of_fill,
                              const float* seg_xyxy, const int16_t* winding, const int32_t* seg_offsets,
                              const uint32_t* rgbx, const uint8_t* winding_rule, const uint8_t* blend_mode,
                              pixie_cmdlist_t* out);
int pixie_cuda_cmdlist_run(pixie_cmdlist_t list, pixie_image_t image, uint64_t* covered_px);
/* newImage(width, height) + run in one call (what `newImage(svg)` svg.nim:557-608 does before its first fill): the
 * canvas is cleared to transparent by the raster kernel itself — every (row, tile) of the canvas is zeroed by the warp
 * that rasterises it next — instead of by a separate pass over the canvas.  Same pixels as image_fill(0) + run. */
int pixie_cuda_cmdlist_run_cleared(pixie_cmdlist_t list, pixie_image_t image, uint64_t* covered_px);
/* One GPU's row band of a canvas split across GPUs: plans and rasterises only rows [y0, y1) of a
 * single-layer list.  `image` is the whole canvas (rows outside the range are left alone) or an image of exactly
 * y1 - y0 rows holding that band.  The list itself — bounds, partition boundaries (paths.nim:1172-1192), band
 * entries — is the whole canvas's, so the pixels equal the undivided render's. */
int pixie_cuda_cmdlist_run_rows(pixie_cmdlist_t list, pixie_image_t image, int y0, int y1, uint64_t* covered_px);
int pixie_cuda_cmdlist_info(pixie_cmdlist_t list, int64_t* num_segments, int64_t* num_partitions,
                            int64_t* num_entries, int64_t* launches_per_run);
int pixie_cuda_cmdlist_destroy(pixie_cmdlist_t list);
/* With PIXIE_CUDA_BANDS=k in the environment single-canvas lists of 2048 rows or more are planned and rasterised in k
 * row bands on concurrent streams (the plan kernels of band b + 1 beside the raster kernel of band b; measured slower
 * than one band on the tiger, so the default is 1).  enabled = 0 runs such a list as one band again. */
int pixie_cuda_cmdlist_set_overlap(pixie_cmdlist_t list, int enabled);

/* ---- path commands -> segments on the device -----------------------------
 * What fillPath (paths.nim:2084-2113) / strokePath (:2144-2176) do BEFORE fillShapes, for a whole document at once:
 * commandsToShapes (:654-1057: lines, quadratic and cubic Beziers with the reference's adaptive halving),
 * strokeShapes (:1922-2082: butt / square caps, miter / bevel joins), transform + shapesToSegments (:1059-1096).
 * The result stays in HBM and becomes the segment block of a command list; only the per-path bounds
 * (computeBounds :1098-1117, 20 bytes per path) come back to the host to lay out the fills.
 *
 * `commands` is the reference's Path.commands stream (seq[float32]: PathCommandKind ordinal followed by its
 * parameters, paths.nim:18-30,73-81) of all paths concatenated.  Arcs, round caps / joins and dashes need the host
 * libm's sin / cos / arccos, which differ from CUDA's in the last bit: such paths are flattened by the caller
 * (kind 2: their finished segments are passed through, `begin`/`end` index raw_xyxy / raw_winding).
 * Returns 1 with the reference's messages ("Unable to discretize ...", "Invalid path command",
 * "Path int overflow detected"). */
typedef struct pixie_path_desc {
  int32_t kind;          /* 0 fill (closeSubpaths = true), 1 stroke, 2 pre-flattened segments */
  int32_t begin, end;    /* float range of this path in `commands` (kind 0 / 1), segment range in raw_* (kind 2) */
  int32_t num_commands;  /* path commands in [begin, end) */
  float transform[9];    /* vmath Mat3, column-major (m[c * 3 + r]); pixelScale is derived from it (:61-66) */
  float stroke_width;    /* kind 1: strokeWidth, LineCap / LineJoin ordinals (paths.nim:10-16), miterLimit */
  int32_t line_cap, line_join;
  float miter_limit;
  uint32_t rgbx;         /* the fill of the resulting shape: colour, WindingRule, BlendMode, layer */
  uint8_t winding_rule, blend_mode;
  uint16_t reserved;
  int32_t layer;
} pixie_path_desc;
int pixie_cuda_cmdlist_create_from_paths(int width, int height, int layers, int num_paths, const pixie_path_desc* paths,
                                         const float* commands, int64_t num_command_floats, const float* raw_xyxy,
                                         const int16_t* raw_winding, int64_t num_raw_segments, pixie_cmdlist_t* out);
/* newImage(svg) (svg.nim:557-608) end to end from path commands: a width x height canvas (transparent when clear != 0,
 * else initialised from `pixels`), flattening / stroking on the device, the ordered fills, the result in `pixels`
 * (row bands on concurrent streams, each band copied out behind its raster kernel; page-locked `pixels` overlap).
 * Every path must have layer 0.  Blocks until `pixels` is complete. */
int pixie_cuda_render_paths_host(uint8_t* pixels, int width, int height, int clear, int num_paths,
                                 const pixie_path_desc* paths, const float* commands, int64_t num_command_floats,
                                 const float* raw_xyxy, const int16_t* raw_winding, int64_t num_raw_segments,
                                 uint64_t* covered_px);
/* The segments a list holds, copied to the host (tests, and callers that want shapesToSegments' output):
 * seg_offsets gets num_fills + 1 entries; any pointer may be NULL.  Blocks. */
int pixie_cuda_cmdlist_segments(pixie_cmdlist_t list, float* seg_xyxy, int16_t* winding, int32_t* seg_offsets);

/* ---- draw -> blendRect (images.nim:636-678 -> :468-529) -------------------------------------
 * Integer-translate fast path of draw(): dst <- blend(dst, src at (px, py)), all 20 modes
 * (blends.nim blender() :275-299); MaskBlend clears dst outside the drawn rect (:499-520) and the
 * whole image when src lies outside (:473-476). */
int pixie_cuda_blend_rect(pixie_image_t dst, pixie_image_t src, int px, int py, int blend_mode);
/* Fused form of the non-solid paint composite (paths.nim:2141-2142):
 *   fill.draw(mask, MaskBlend); image.draw(fill, blendMode)
 * = dst <- blend(dst, floor(src * mask.a / 255)) in one pass.  mask is an RGBX image (alpha used)
 * or an A8 coverage plane of src's size; src itself is left unchanged. */
int pixie_cuda_blend_rect_masked(pixie_image_t dst, pixie_image_t src, pixie_image_t mask, int px, int py,
                                 int blend_mode);
int pixie_cuda_apply_opacity(pixie_image_t image, float opacity);              /* images.nim:261-277 */

/* ---- draw with any transform (images.nim:636-678) -------------------------------------------
 * Replaces the body of draw(a, b, transform, blendMode): mat is vmath Mat3 storage (9 float32, column-major:
 * mat[0..2] = column 0, mat[6], mat[7] = translation).  Exactly the reference's flow: movement vectors from
 * the inverse transform, minifyBy2 / magnifyBy2 of the source while the scale is >= 2 or <= 0.5
 * (:649-663), then drawSmooth (:531-634: per-row x range from the transformed perimeter, bilinear
 * getRgbaSmooth :367-403, blendLineNormal / Overwrite / Mask or blender()) or, for integer translations,
 * blendRect (:678).  dst and src must be different single-layer RGBX images. */
int pixie_cuda_draw(pixie_image_t dst, pixie_image_t src, const float* mat, int blend_mode);
/* drawTiled (images.nim:680-683) = drawCorrect(..., tiled = true) (:405-449): every dst pixel samples the
 * wrapped source through blender(); used by TiledImagePaint fills (paths.nim:2131-2132). */
int pixie_cuda_draw_tiled(pixie_image_t dst, pixie_image_t src, const float* mat, int blend_mode);
int pixie_cuda_draw_correct(pixie_image_t dst, pixie_image_t src, const float* mat, int blend_mode); /* untiled */
/* minifyBy2 / magnifyBy2 (images.nim:168-259): *out receives a new image handle.  power < 0 -> 1
 * "Cannot minifyBy2 with negative power" (:172-173, :242-243). */
int pixie_cuda_minify_by2(pixie_image_t src, int power, pixie_image_t* out);
int pixie_cuda_magnify_by2(pixie_image_t src, int power, pixie_image_t* out);

/* ---- gradient paints: fillGradient (paints.nim:68-248) ---------------------------------------
 * kind = ord(PaintKind) (paints.nim:4-10): 3 LinearGradientPaint (2 handles), 4 RadialGradientPaint (3),
 * 5 AngularGradientPaint (3).  handles_xy: n_handles x {x, y} (gradientHandlePositions); stops:
 * stop_pos[n_stops] and stop_rgba[n_stops][4] = chroma Color (straight float32 r, g, b, a) in list order
 * (gradientStops); opacity = paint.opacity (clamped to 0..1, 0 is a no-op).  Errors as the reference:
 * "Linear gradient requires 2 handles", "Gradient must have at least 1 color stop", ...
 * Every pixel of the image is overwritten (image.unsafe[x, y] = gradientColor(t)). */
int pixie_cuda_fill_gradient(pixie_image_t image, int kind, const float* handles_xy, int n_handles,
                             const float* stop_pos, const float* stop_rgba, int n_stops, float opacity);
/* The composite of a non-solid fillPath / strokePath whose paint is a gradient (paths.nim:2115-2142) in one pass:
 * `fill.fillGradient(paint at opacity 1); mask.applyOpacity(paint.opacity); fill.draw(mask, MaskBlend);
 * image.draw(fill, blend_mode)` with the gradient evaluated inside the blend — the fill image never exists and, for
 * NormalBlend, pixels where the mask is 0 are not touched.  mask: canvas-sized RGBX (alpha used) or A8 image holding
 * the shape's coverage (what fillPath(mask, white) leaves).  Bit-identical to the four separate calls. */
int pixie_cuda_fill_gradient_masked(pixie_image_t image, pixie_image_t mask, int kind, const float* handles_xy,
                                    int n_handles, const float* stop_pos, const float* stop_rgba, int n_stops,
                                    float opacity, int blend_mode);

/* ---- blur / spread / shadow (images.nim:304-365, :700-758, :760-776) ------------------------
 * lut = gaussianKernel(radius) (internal.nim:17-34), 2*radius+1 uint16 taps, computed by the
 * caller.  radius < 0 -> 1 "Cannot apply negative blur" (:311-312); radius == 0 is a no-op. */
/* Radii whose taps stay below 2048 (29..32) on images whose width is a multiple of 4 run as ONE fused pass on the tcgen05
 * tensor cores (blur_tc.cu); that pass is out of place, so a whole-image blur of a library-owned image moves its pixels
 * to a fresh device buffer (pixie_cuda_image_info returns the new pointer); wrapped images keep their memory. */
int pixie_cuda_blur(pixie_image_t image, const uint16_t* lut, int radius, uint32_t out_of_bounds_rgbx);
/* Row-band form used when one canvas is split across GPUs: computes the blur of the whole image
 * but writes only rows [y0, y1); the other rows (the halo received from the neighbours) are left
 * unchanged.  With `radius` halo rows above and below, rows [y0, y1) equal the global blur. */
int pixie_cuda_blur_rows(pixie_image_t image, const uint16_t* lut, int radius, uint32_t out_of_bounds_rgbx,
                         int y0, int y1);
/* Out-of-place row-band blur: rows [y0, y1) of dst <- rows [y0, y1) of blur(src); src and the other rows of dst are
 * left unchanged.  Several calls on one src (disjoint row ranges) compose: a band split across GPUs blurs its interior
 * rows while the halo rows are still on their way and its edge rows afterwards (pixie_b200/multi.py). */
int pixie_cuda_blur_rows_to(pixie_image_t src, pixie_image_t dst, const uint16_t* lut, int radius,
                            uint32_t out_of_bounds_rgbx, int y0, int y1);
/* The same for a band whose halo rows (rows [0, y0) and [y1, height) of src) are being stored by neighbour GPUs
 * (pixie_cuda_halo_exchange): only the tiles that read halo rows wait — inside the kernel — until the flag the
 * neighbour publishes after its rows has reached `epoch`; the band's interior is blurred meanwhile.  NULL = that side
 * has no neighbour. */
int pixie_cuda_blur_rows_to_flags(pixie_image_t src, pixie_image_t dst, const uint16_t* lut, int radius,
                                  uint32_t out_of_bounds_rgbx, int y0, int y1, const void* top_flag,
                                  const void* bottom_flag, uint32_t epoch);
/* The two halves of pixie_cuda_blur_rows, for callers that overlap the halo exchange with the X pass (which needs no
 * halo): _x blurs rows [r0, r1) horizontally into the library's scratch plane; _y then produces image rows
 * [y0, y1) from scratch rows [y0 - radius, y1 + radius) — each of which an _x call must have produced since the last
 * other library call on an image of another size.  Radii 1..64 with gaussianKernel LUTs only (else status 1). */
int pixie_cuda_blur_rows_x(pixie_image_t image, const uint16_t* lut, int radius, uint32_t out_of_bounds_rgbx,
                           int r0, int r1);
int pixie_cuda_blur_rows_y(pixie_image_t image, const uint16_t* lut, int radius, uint32_t out_of_bounds_rgbx,
                           int y0, int y1);
int pixie_cuda_spread(pixie_image_t image, int spread);
/* Row-band forms of spread / shadow: `image` (src / dst) is [halo ; band ; halo], the caller keeps rows [y0, y1);
 * rows outside are left unspecified.  An edge of [y0, y1) that is not an edge of the image needs a halo of
 * |spread| rows (spread) resp. ceil|offset_y| + |spread| + radius rows (shadow), else status 1. */
int pixie_cuda_spread_rows(pixie_image_t image, int spread, int y0, int y1);
/* dst <- shadow(src, offset, spread, blur, color); the offset copy is mask.draw(image, translate(offset),
 * OverwriteBlend) (:768-769): blendRect for integral offsets, drawSmooth otherwise.  Only the mask's alpha reaches
 * the result (:760-776), so with an integral offset and radius <= 64 the whole pipeline runs on an 8-bit alpha plane
 * (same bytes out; other cases go through the RGBX mask as the reference does).  src and dst must differ. */
int pixie_cuda_shadow(pixie_image_t src, pixie_image_t dst, float offset_x, float offset_y, int spread,
                      const uint16_t* lut, int radius, uint32_t rgbx);

int pixie_cuda_shadow_rows(pixie_image_t src, pixie_image_t dst, float offset_x, float offset_y, int spread,
                           const uint16_t* lut, int radius, uint32_t rgbx, int y0, int y1);

/* ---- halo rows between the GPUs of one box, without a collective library (one process per GPU) ---------------
 * peer_alloc: a device buffer other processes can map (cudaMalloc + CUDA IPC handle, 64 bytes, to be sent to the
 * neighbours by any means); peer_open maps a neighbour's buffer into this process (peer access over NVLink).
 * halo_push (stream-ordered): copies `bytes` from local `src_rows` into the neighbour's mapped memory and then stores
 * `value` at `peer_flag` (a uint32 in the neighbour's buffer) with release semantics at system scope; halo_wait
 * (stream-ordered): everything issued after it runs once the local flag has reached `value` (epochs only grow).
 * Used by pixie_b200/multi.py RowBand(transport="peer") for blur / spread / shadow on a canvas split in row bands. */
int pixie_cuda_peer_alloc(size_t bytes, void** device_ptr, uint8_t* ipc_handle_64);
int pixie_cuda_peer_open(const uint8_t* ipc_handle_64, void** device_ptr);
int pixie_cuda_peer_close(void* device_ptr);
int pixie_cuda_peer_free(void* device_ptr);
int pixie_cuda_halo_push(const void* src_rows, void* peer_dst_rows, size_t bytes, void* peer_flag, uint32_t value);
int pixie_cuda_halo_wait(const void* local_flag, uint32_t value);
/* One epoch's exchange with the upper and / or lower neighbour (NULL = none) in ONE launch: store `epoch` at the
 * neighbours' ready flags ("my margin is free"), wait until the local ready flags written by the neighbours reach it,
 * store the rows into their margins, store `epoch` at their data flags.  halo_wait2 then holds the stream until both
 * local data flags have reached `epoch`. */
typedef struct {
  const void* src_rows;         /* local rows to send */
  void* peer_dst_rows;          /* the neighbour's margin (mapped with pixie_cuda_peer_open) */
  size_t bytes;
  void* peer_ready_flag;        /* in the neighbour's buffer: this rank's margin is free for `epoch` */
  void* peer_data_flag;         /* in the neighbour's buffer: the rows of `epoch` have arrived */
  const void* local_ready_flag; /* in this rank's buffer, written by that neighbour */
} pixie_halo_dir_t;
int pixie_cuda_halo_exchange(const pixie_halo_dir_t* up, const pixie_halo_dir_t* down, uint32_t epoch);
int pixie_cuda_halo_wait2(const void* local_flag_a, const void* local_flag_b, uint32_t value);

/* ---- strict drop-in variants: host pixels in, host pixels out (upload -> run -> download) ---- */
int pixie_cuda_fill_segments_host(uint8_t* pixels, int width, int height, const float* seg_xyxy,
                                  const int16_t* winding, int n, uint32_t rgbx, int winding_rule, int blend_mode);
int pixie_cuda_blend_rect_host(uint8_t* dst_pixels, int dst_width, int dst_height, const uint8_t* src_pixels,
                               int src_width, int src_height, int px, int py, int blend_mode);
int pixie_cuda_blur_host(uint8_t* pixels, int width, int height, const uint16_t* lut, int radius,
                         uint32_t out_of_bounds_rgbx);
int pixie_cuda_shadow_host(const uint8_t* src_pixels, uint8_t* dst_pixels, int width, int height, float offset_x,
                           float offset_y, int spread, const uint16_t* lut, int radius, uint32_t rgbx);

int pixie_cuda_spread_host(uint8_t* pixels, int width, int height, int spread);                 /* images.nim:700-758 */
int pixie_cuda_apply_opacity_host(uint8_t* pixels, int width, int height, float opacity);      /* images.nim:261-277 */
/* the non-solid paint composite (paths.nim:2141-2142) on host pixels; mask: src-sized A8 (1) or RGBX (4) plane */
int pixie_cuda_blend_rect_masked_host(uint8_t* dst_pixels, int dst_width, int dst_height, const uint8_t* src_pixels,
                                      const uint8_t* mask_pixels, int mask_bytes_per_pixel, int src_width,
                                      int src_height, int px, int py, int blend_mode);

/* draw(a, b, transform, blendMode) / drawTiled on host pixels (tiled != 0), images.nim:636-683 */
int pixie_cuda_draw_host(uint8_t* dst_pixels, int dst_width, int dst_height, const uint8_t* src_pixels, int src_width,
                         int src_height, const float* mat, int blend_mode, int tiled);
/* image.fillGradient(paint) on host pixels, paints.nim:236-248 */
int pixie_cuda_fill_gradient_host(uint8_t* pixels, int width, int height, int kind, const float* handles_xy,
                                  int n_handles, const float* stop_pos, const float* stop_rgba, int n_stops,
                                  float opacity);
/* minifyBy2 / magnifyBy2 on host pixels; dst holds ceil(w / 2^power) x ceil(h / 2^power) resp. (w << power) x
 * (h << power) pixels (images.nim:168-259) */
int pixie_cuda_minify_by2_host(const uint8_t* src_pixels, int width, int height, int power, uint8_t* dst_pixels);
int pixie_cuda_magnify_by2_host(const uint8_t* src_pixels, int width, int height, int power, uint8_t* dst_pixels);

/* page-locked host staging memory for the *_async copies */
int pixie_cuda_host_alloc(size_t bytes, void** out);
int pixie_cuda_host_free(void* ptr);

/* ---- instrumentation ----------------------------------------------------------------------- */
/* number of kernels this library has launched since init (bench.py's gpu_launches) */
int pixie_cuda_launch_count(uint64_t* out);
/* per-kernel CUDA-event timing on the library's stream.  Slots: 0 partition kernel, 1 raster kernel,
 * 2 blur X pass, 3 blur Y pass, 4 blend_rect kernel, 5 spread kernels,
 * 6 plan kernel.  profile_read waits for the slot's last launch and returns its duration in milliseconds. */
int pixie_cuda_set_profiling(int enabled);
int pixie_cuda_profile_read(int slot, float* elapsed_ms);
/* CUDA-event timing on the library's stream: begin/end bracket, elapsed in milliseconds */
int pixie_cuda_timer_begin(void);
int pixie_cuda_timer_end(float* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif

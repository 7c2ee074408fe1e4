## pixie_cuda.nim — thin shim that routes Pixie's raster hot path to pixie_cuda.so.
##
## NOT compiled or tested in this repository (no Nim toolchain in the build image); it is the
## binding a Pixie maintainer would add.  It replaces the BODIES of the private procs the public
## API funnels through — `fillShapes` (src/pixie/paths.nim:1593), `blendRect`
## (src/pixie/images.nim:468), `blur` (:304), `spread` (:700), `shadow` (:760) — and leaves every
## public signature (`fillPath`, `strokePath`, `draw`, `blur`, `shadow`, `Paint`, `WindingRule`,
## `BlendMode`) untouched.  See INTEGRATION.md.

## Imports: only modules that do NOT import this one.  `WindingRule` lives in pixie/paths (which imports the shim), so
## fillShapesCuda takes `ord(windingRule)`; `Paint` likewise (pixie/paints), so fillGradientCuda takes its fields.
import std/math                # round
import bumpy, chroma, vmath    # Segment; Color, ColorRGBX; Vec2, Ivec2, Mat3
import pixie/common            # Image, BlendMode, PixieError, newImage
import pixie/internal          # gaussianKernel (internal.nim:17-34)

const lib = "pixie_cuda.so"

type PixieImageT = uint64

proc pixie_cuda_last_error(): cstring {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_fill_segments_host(pixels: ptr uint8, width, height: cint,
    segXyxy: ptr float32, winding: ptr int16, n: cint, rgbx: uint32,
    windingRule, blendMode: cint): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_blend_rect_host(dst: ptr uint8, dw, dh: cint, src: ptr uint8, sw, sh: cint,
    px, py, blendMode: cint): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_blur_host(pixels: ptr uint8, width, height: cint, lut: ptr uint16,
    radius: cint, outOfBounds: uint32): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_shadow_host(src, dst: ptr uint8, width, height: cint, ox, oy: cfloat,
    spread: cint, lut: ptr uint16, radius: cint, rgbx: uint32): cint {.importc, dynlib: lib, cdecl.}

proc pixie_cuda_spread_host(pixels: ptr uint8, width, height, spread: cint): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_apply_opacity_host(pixels: ptr uint8, width, height: cint, opacity: cfloat): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_blend_rect_masked_host(dst: ptr uint8, dw, dh: cint, src, mask: ptr uint8,
    maskBytesPerPixel, sw, sh, px, py, blendMode: cint): cint {.importc, dynlib: lib, cdecl.}

proc pixie_cuda_draw_host(dst: ptr uint8, dw, dh: cint, src: ptr uint8, sw, sh: cint,
    mat: ptr float32, blendMode, tiled: cint): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_fill_gradient_host(pixels: ptr uint8, width, height, kind: cint,
    handlesXy: ptr float32, nHandles: cint, stopPos, stopRgba: ptr float32, nStops: cint,
    opacity: cfloat): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_minify_by2_host(src: ptr uint8, width, height, power: cint,
    dst: ptr uint8): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_magnify_by2_host(src: ptr uint8, width, height, power: cint,
    dst: ptr uint8): cint {.importc, dynlib: lib, cdecl.}

# device-resident variants (handles), for callers that keep canvases in HBM between calls
proc pixie_cuda_image_create(width, height: cint, outH: ptr PixieImageT): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_image_upload(image: PixieImageT, pixels: ptr uint8): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_image_download(image: PixieImageT, pixels: ptr uint8): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_image_destroy(image: PixieImageT): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_fill_batch(image: PixieImageT, numFills: cint, layerOfFill: ptr int32,
    segXyxy: ptr float32, winding: ptr int16, segOffsets: ptr int32, rgbx: ptr uint32,
    windingRule, blendMode: ptr uint8, coveredPx: ptr uint64): cint {.importc, dynlib: lib, cdecl.}

template check(rc: cint) =
  if rc != 0:
    raise newException(PixieError, $pixie_cuda_last_error())

proc asU32(c: ColorRGBX): uint32 {.inline.} = cast[uint32](c)

# ---- paths.nim: body of fillShapes ------------------------------------------------------------
proc fillShapesCuda*(
  image: Image,
  segments: seq[(Segment, int16)],   # output of shapesToSegments (paths.nim:1059-1090)
  rgbx: ColorRGBX,                   # color.asRgbx() (paths.nim:1603)
  windingRule: int,                  # ord(WindingRule): NonZero = 0, EvenOdd = 1 (paths.nim:5-8)
  blendMode: BlendMode
) {.raises: [PixieError].} =
  var
    xyxy = newSeq[float32](segments.len * 4)
    winding = newSeq[int16](segments.len)
  for i, (segment, w) in segments:
    xyxy[i * 4 + 0] = segment.at.x
    xyxy[i * 4 + 1] = segment.at.y
    xyxy[i * 4 + 2] = segment.to.x
    xyxy[i * 4 + 3] = segment.to.y
    winding[i] = w
  if segments.len == 0:
    return
  check pixie_cuda_fill_segments_host(
    cast[ptr uint8](image.data[0].addr), image.width.cint, image.height.cint,
    xyxy[0].addr, winding[0].addr, segments.len.cint, rgbx.asU32,
    windingRule.cint, blendMode.ord.cint)

# ---- images.nim: body of blendRect -------------------------------------------------------------
proc blendRectCuda*(a, b: Image, pos: Ivec2, blendMode: BlendMode) {.raises: [PixieError].} =
  check pixie_cuda_blend_rect_host(
    cast[ptr uint8](a.data[0].addr), a.width.cint, a.height.cint,
    cast[ptr uint8](b.data[0].addr), b.width.cint, b.height.cint,
    pos.x.cint, pos.y.cint, blendMode.ord.cint)

# ---- images.nim: body of blur ------------------------------------------------------------------
proc blurCuda*(image: Image, radius: float32, outOfBounds: ColorRGBX) {.raises: [PixieError].} =
  let radius = round(radius).int
  if radius == 0:
    return
  if radius < 0:
    raise newException(PixieError, "Cannot apply negative blur")
  var kernel = gaussianKernel(radius)   # internal.nim:17-34, unchanged
  check pixie_cuda_blur_host(
    cast[ptr uint8](image.data[0].addr), image.width.cint, image.height.cint,
    kernel[0].addr, radius.cint, outOfBounds.asU32)

# ---- images.nim: body of spread (:700-758) -----------------------------------------------------------
proc spreadCuda*(image: Image, spread: float32) {.raises: [PixieError].} =
  let spread = round(spread).int
  if spread == 0:
    return
  check pixie_cuda_spread_host(
    cast[ptr uint8](image.data[0].addr), image.width.cint, image.height.cint, spread.cint)

# ---- images.nim: body of applyOpacity (:261-277) ------------------------------------------------------
proc applyOpacityCuda*(image: Image, opacity: float32) {.raises: [PixieError].} =
  check pixie_cuda_apply_opacity_host(
    cast[ptr uint8](image.data[0].addr), image.width.cint, image.height.cint, opacity.cfloat)

# ---- paths.nim:2141-2142: fill.draw(mask, MaskBlend); image.draw(fill, blendMode) in one pass ---------
proc drawMaskedCuda*(image, fill, mask: Image, blendMode: BlendMode) {.raises: [PixieError].} =
  check pixie_cuda_blend_rect_masked_host(
    cast[ptr uint8](image.data[0].addr), image.width.cint, image.height.cint,
    cast[ptr uint8](fill.data[0].addr), cast[ptr uint8](mask.data[0].addr), 4.cint,
    fill.width.cint, fill.height.cint, 0.cint, 0.cint, blendMode.ord.cint)

# ---- images.nim: body of shadow ----------------------------------------------------------------
proc shadowCuda*(image: Image, offset: Vec2, spread, blur: float32, color: ColorRGBX): Image
    {.raises: [PixieError].} =
  result = newImage(image.width, image.height)
  let radius = round(blur).int
  var kernel = gaussianKernel(max(radius, 0))
  check pixie_cuda_shadow_host(
    cast[ptr uint8](image.data[0].addr), cast[ptr uint8](result.data[0].addr),
    image.width.cint, image.height.cint, offset.x.cfloat, offset.y.cfloat,
    round(spread).cint, kernel[0].addr, radius.cint, color.asU32)

# ---- images.nim: body of draw (:636-678) and drawTiled (:680-683) -------------------------------
# vmath Mat3 is array[3, Vec3] in column-major order: its memory is the 9 float32 the C ABI takes.
proc drawCuda*(a, b: Image, transform: Mat3, blendMode: BlendMode, tiled = false)
    {.raises: [PixieError].} =
  var m = transform
  check pixie_cuda_draw_host(
    cast[ptr uint8](a.data[0].addr), a.width.cint, a.height.cint,
    cast[ptr uint8](b.data[0].addr), b.width.cint, b.height.cint,
    cast[ptr float32](m.addr), blendMode.ord.cint, tiled.ord.cint)

# ---- images.nim: bodies of minifyBy2 (:168-236) / magnifyBy2 (:238-259) --------------------------
proc minifyBy2Cuda*(image: Image, power = 1): Image {.raises: [PixieError].} =
  if power < 0:
    raise newException(PixieError, "Cannot minifyBy2 with negative power")
  var (w, h) = (image.width, image.height)
  for _ in 1 .. power:
    w = (w + 1) div 2
    h = (h + 1) div 2
  result = newImage(w, h)
  check pixie_cuda_minify_by2_host(
    cast[ptr uint8](image.data[0].addr), image.width.cint, image.height.cint, power.cint,
    cast[ptr uint8](result.data[0].addr))

proc magnifyBy2Cuda*(image: Image, power = 1): Image {.raises: [PixieError].} =
  if power < 0:
    raise newException(PixieError, "Cannot magnifyBy2 with negative power")
  result = newImage(image.width shl power, image.height shl power)
  check pixie_cuda_magnify_by2_host(
    cast[ptr uint8](image.data[0].addr), image.width.cint, image.height.cint, power.cint,
    cast[ptr uint8](result.data[0].addr))

# ---- paints.nim: body of fillGradient (:236-248) --------------------------------------------------
# `Paint` is declared in paints.nim, which imports this module's callers; pass its fields.
proc fillGradientCuda*(image: Image, kind: int, handles: seq[Vec2], stopPositions: seq[float32],
                       stopColors: seq[Color], opacity: float32) {.raises: [PixieError].} =
  var
    hx = newSeq[float32](handles.len * 2)
    col = newSeq[float32](stopColors.len * 4)
    pos = stopPositions
  for i, p in handles:
    hx[i * 2] = p.x
    hx[i * 2 + 1] = p.y
  for i, c in stopColors:
    col[i * 4] = c.r; col[i * 4 + 1] = c.g; col[i * 4 + 2] = c.b; col[i * 4 + 3] = c.a
  check pixie_cuda_fill_gradient_host(
    cast[ptr uint8](image.data[0].addr), image.width.cint, image.height.cint, kind.cint,
    (if hx.len > 0: hx[0].addr else: nil), handles.len.cint,
    (if pos.len > 0: pos[0].addr else: nil), (if col.len > 0: col[0].addr else: nil),
    stopColors.len.cint, opacity.cfloat)

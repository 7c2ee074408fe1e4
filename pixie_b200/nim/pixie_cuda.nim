## pixie_cuda.nim — thin shim that routes Pixie's raster hot path to pixie_cuda.so.
##
## NOT compiled or tested in this repository (no Nim toolchain in the build image); it is the
## binding a Pixie maintainer would add.  It replaces the BODIES of the private procs the public
## API funnels through — `fillShapes` (src/pixie/paths.nim:1593), `blendRect`
## (src/pixie/images.nim:468), `blur` (:304), `spread` (:700), `shadow` (:760) — and leaves every
## public signature (`fillPath`, `strokePath`, `draw`, `blur`, `shadow`, `Paint`, `WindingRule`,
## `BlendMode`) untouched.  See INTEGRATION.md.

import chroma, vmath
import pixie/common   # Image, BlendMode, PixieError

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
  windingRule: WindingRule,
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
    windingRule.ord.cint, blendMode.ord.cint)

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

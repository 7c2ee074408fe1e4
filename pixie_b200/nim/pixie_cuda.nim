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

# ---- paths.nim: the gradient branch of the non-solid composite (:2115-2142) in one pass ----------------
proc pixie_cuda_fill_gradient_masked(image, mask: PixieImageT, kind: cint, handlesXy: ptr float32, nHandles: cint,
    stopPos, stopRgba: ptr float32, nStops: cint, opacity: cfloat, blendMode: cint): cint {.importc, dynlib: lib, cdecl.}

proc fillGradientMaskedCuda*(image, mask: Image, kind: int, handles: seq[Vec2], stopPositions: seq[float32],
                             stopColors: seq[Color], opacity: float32, blendMode: BlendMode) {.raises: [PixieError].} =
  ## `fill.fillGradient(paint at opacity 1); mask.applyOpacity(paint.opacity); fill.draw(mask, MaskBlend);
  ## image.draw(fill, blendMode)` — `mask` is what `mask.fillPath(path, white)` left (:2112-2113).
  var
    hx = newSeq[float32](handles.len * 2)
    col = newSeq[float32](stopColors.len * 4)
    pos = stopPositions
    hImage, hMask: PixieImageT
  for i, p in handles:
    hx[i * 2] = p.x
    hx[i * 2 + 1] = p.y
  for i, c in stopColors:
    col[i * 4] = c.r; col[i * 4 + 1] = c.g; col[i * 4 + 2] = c.b; col[i * 4 + 3] = c.a
  check pixie_cuda_image_create(image.width.cint, image.height.cint, hImage.addr)
  check pixie_cuda_image_create(mask.width.cint, mask.height.cint, hMask.addr)
  try:
    check pixie_cuda_image_upload(hImage, cast[ptr uint8](image.data[0].addr))
    check pixie_cuda_image_upload(hMask, cast[ptr uint8](mask.data[0].addr))
    check pixie_cuda_fill_gradient_masked(hImage, hMask, kind.cint,
      (if hx.len > 0: hx[0].addr else: nil), handles.len.cint,
      (if pos.len > 0: pos[0].addr else: nil), (if col.len > 0: col[0].addr else: nil),
      stopColors.len.cint, opacity.cfloat, blendMode.ord.cint)
    check pixie_cuda_image_download(hImage, cast[ptr uint8](image.data[0].addr))
  finally:
    discard pixie_cuda_image_destroy(hImage)
    discard pixie_cuda_image_destroy(hMask)

# ---- paths.nim: fillPath / strokePath of a whole document from path COMMANDS ---------------------------
# commandsToShapes (:654-1057), strokeShapes (:1922-2082), transform, shapesToSegments (:1059-1096) run on the device.
type
  PixiePathDesc* {.bycopy.} = object   # pixie_path_desc (include/pixie_cuda.h), 80 bytes
    kind*, begin*, `end`*, numCommands*: int32
    transform*: array[9, float32]
    strokeWidth*: float32
    lineCap*, lineJoin*: int32
    miterLimit*: float32
    rgbx*: uint32
    windingRule*, blendMode*: uint8
    reserved*: uint16
    layer*: int32
  PathBatchCuda* = object
    descs*: seq[PixiePathDesc]
    commands*: seq[float32]            # Path.commands of the device-flattened paths, concatenated
    rawXyxy*: seq[float32]             # segments of the paths flattened by the Nim code (kind = 2)
    rawWinding*: seq[int16]

proc pixie_cuda_cmdlist_create_from_paths(width, height, layers, numPaths: cint, paths: ptr PixiePathDesc,
    commands: ptr float32, numCommandFloats: int64, rawXyxy: ptr float32, rawWinding: ptr int16, numRaw: int64,
    outList: ptr uint64): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_cmdlist_run(list: uint64, image: PixieImageT, coveredPx: ptr uint64): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_cmdlist_run_cleared(list: uint64, image: PixieImageT, coveredPx: ptr uint64): cint {.importc, dynlib: lib, cdecl.}
proc pixie_cuda_cmdlist_destroy(list: uint64): cint {.importc, dynlib: lib, cdecl.}

const
  parameterCounts = [0, 2, 2, 1, 1, 6, 4, 4, 2, 7, 2, 2, 1, 1, 6, 4, 4, 2, 7]   # paths.nim:73-81, by ord(PathCommandKind)
  arcKinds = {9, 18}

proc scanCommands(commands: seq[float32]): (int, bool) =
  var i = 0
  while i < commands.len:
    let k = commands[i].int
    if k in arcKinds: result[1] = true
    i += 1 + parameterCounts[k]
    inc result[0]

proc addDesc(batch: var PathBatchCuda, kind, b, e, n: int, transform: Mat3, rgbx: ColorRGBX, rule: int, mode: BlendMode) =
  var d = PixiePathDesc(kind: kind.int32, begin: b.int32, `end`: e.int32, numCommands: n.int32, rgbx: rgbx.asU32,
                        windingRule: rule.uint8, blendMode: mode.ord.uint8, miterLimit: 4)
  copyMem(d.transform[0].addr, transform.unsafeAddr, 9 * sizeof(float32))
  batch.descs.add d

proc addFill*(batch: var PathBatchCuda, commands: seq[float32], transform: Mat3, rgbx: ColorRGBX, windingRule: int,
              blendMode: BlendMode, hostSegments: proc(): seq[(Segment, int16)]) =
  ## `hostSegments` = the unchanged Nim code (parseSomePath + transform + shapesToSegments), called for paths with arcs
  let (n, arc) = scanCommands(commands)
  if arc:
    let b = batch.rawWinding.len
    for (s, w) in hostSegments():
      batch.rawXyxy.add [s.at.x, s.at.y, s.to.x, s.to.y]
      batch.rawWinding.add w
    batch.addDesc(2, b, batch.rawWinding.len, 0, mat3(), rgbx, windingRule, blendMode)
  else:
    let b = batch.commands.len
    batch.commands.add commands
    batch.addDesc(0, b, batch.commands.len, n, transform, rgbx, windingRule, blendMode)

proc addStroke*(batch: var PathBatchCuda, commands: seq[float32], transform: Mat3, strokeWidth: float32,
                lineCap, lineJoin: int, miterLimit: float32, dashes: seq[float32], rgbx: ColorRGBX,
                hostSegments: proc(): seq[(Segment, int16)]) =
  let (n, arc) = scanCommands(commands)
  if arc or lineCap == 1 or lineJoin == 1 or dashes.len > 0:   # RoundCap / RoundJoin (paths.nim:10-16), dashes
    let b = batch.rawWinding.len
    for (s, w) in hostSegments():
      batch.rawXyxy.add [s.at.x, s.at.y, s.to.x, s.to.y]
      batch.rawWinding.add w
    batch.addDesc(2, b, batch.rawWinding.len, 0, mat3(), rgbx, 0, NormalBlend)
  else:
    let b = batch.commands.len
    batch.commands.add commands
    batch.addDesc(1, b, batch.commands.len, n, transform, rgbx, 0, NormalBlend)
    batch.descs[^1].strokeWidth = strokeWidth
    batch.descs[^1].lineCap = lineCap.int32
    batch.descs[^1].lineJoin = lineJoin.int32
    batch.descs[^1].miterLimit = miterLimit

proc pixie_cuda_render_paths_host(pixels: ptr uint8, width, height, clear, numPaths: cint, paths: ptr PixiePathDesc,
    commands: ptr float32, numCommandFloats: int64, rawXyxy: ptr float32, rawWinding: ptr int16, numRaw: int64,
    coveredPx: ptr uint64): cint {.importc, dynlib: lib, cdecl.}

proc renderPathsCuda*(image: Image, batch: var PathBatchCuda) {.raises: [PixieError].} =
  ## the same in ONE call on host pixels: commands in, flattening / stroking / rasterising on the device, the band-wise
  ## copy back overlapped with the rendering (pixie_cuda_render_paths_host)
  check pixie_cuda_render_paths_host(cast[ptr uint8](image.data[0].addr), image.width.cint, image.height.cint, 0,
    batch.descs.len.cint, (if batch.descs.len > 0: batch.descs[0].addr else: nil),
    (if batch.commands.len > 0: batch.commands[0].addr else: nil), batch.commands.len.int64,
    (if batch.rawXyxy.len > 0: batch.rawXyxy[0].addr else: nil),
    (if batch.rawWinding.len > 0: batch.rawWinding[0].addr else: nil), batch.rawWinding.len.int64, nil)

proc fillPathsCuda*(image: Image, batch: var PathBatchCuda) {.raises: [PixieError].} =
  ## the ordered fills / strokes of `batch` over `image` (newImage(svg), svg.nim:557-608)
  var
    h: PixieImageT
    list: uint64
  check pixie_cuda_image_create(image.width.cint, image.height.cint, h.addr)
  try:
    check pixie_cuda_image_upload(h, cast[ptr uint8](image.data[0].addr))
    check pixie_cuda_cmdlist_create_from_paths(image.width.cint, image.height.cint, 1, batch.descs.len.cint,
      (if batch.descs.len > 0: batch.descs[0].addr else: nil),
      (if batch.commands.len > 0: batch.commands[0].addr else: nil), batch.commands.len.int64,
      (if batch.rawXyxy.len > 0: batch.rawXyxy[0].addr else: nil),
      (if batch.rawWinding.len > 0: batch.rawWinding[0].addr else: nil), batch.rawWinding.len.int64, list.addr)
    try:
      check pixie_cuda_cmdlist_run(list, h, nil)
      check pixie_cuda_image_download(h, cast[ptr uint8](image.data[0].addr))
    finally:
      discard pixie_cuda_cmdlist_destroy(list)
  finally:
    discard pixie_cuda_image_destroy(h)

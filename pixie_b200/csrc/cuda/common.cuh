// Shared runtime state + device helpers of pixie_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "pixie_cuda.h"

namespace pixie {

// ------------------------------------------------------------------ host runtime
struct Image {
  uint8_t* data = nullptr;
  int w = 0, h = 0, layers = 1, bpp = 4;
  bool owned = true;
  size_t bytes() const { return (size_t)w * h * layers * bpp; }
  size_t layer_bytes() const { return (size_t)w * h * bpp; }
};

struct Runtime {
  bool inited = false;
  int device = 0;
  int num_sms = 148;
  int sm_reserve = 0;  // SMs the persistent one-CTA-per-SM kernels leave free (pixie_cuda_set_sm_reserve)
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  std::unordered_map<uint64_t, Image> images;
  uint64_t next_handle = 1;
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t switch_ev = nullptr;  // orders a new current stream after the old one (pixie_cuda_set_stream)
  // scratch buffers that grow on demand (stream-ordered reuse)
  void* scratch[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[6] = {0, 0, 0, 0, 0, 0};
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  cudaEvent_t staging_done = nullptr;  // recorded after the H2D copy that reads `pinned`
  bool staging_busy = false;
  // row-band streams of pixie_cuda_render_batch_host (rasterise band b while band b-1 goes to the host)
  static constexpr int kBands = 8;
  cudaStream_t band_stream[kBands] = {};
  cudaEvent_t band_done[kBands] = {};
  cudaEvent_t band_start = nullptr;
  // auxiliary streams: the two plan kernels of a launch run side by side ([0] library stream, [1 + b] band b)
  cudaStream_t aux_stream[kBands + 1] = {};
  cudaEvent_t aux_fork[kBands + 1] = {}, aux_join[kBands + 1] = {};
  // per-kernel CUDA-event timing (pixie_cuda_set_profiling): slot -> (begin, end) of the last launch
  bool profiling = false;
  cudaEvent_t prof[8][2] = {};
  // a command list run in row bands (plan of band b + 1 beside the raster of band b) has one plan span and one raster
  // launch per band, each on the band's stream: [band][0..1] plan, [band][2..3] raster; prof_bands = bands of the
  // last run (0: the run was not banded and slots kProfPlan / kProfRaster hold its times)
  cudaEvent_t band_prof[kBands][4] = {};
  int prof_bands = 0;
};

enum ProfSlot { kProfPartition = 0, kProfRaster = 1, kProfBlurX = 2, kProfBlurY = 3, kProfBlend = 4, kProfSpread = 5, kProfPlan = 6 };
struct ProfScope {  // brackets one kernel launch with events on the library stream when profiling is on
  int slot;
  explicit ProfScope(int s);
  ~ProfScope();
};

Runtime& rt();
// Threading contract of the C ABI (include/pixie_cuda.h "Threading"): ONE coarse recursive lock serialises every
// entry point — the library has one stream, one scratch set and one pinned staging buffer, so calls from several
// host threads (on the same or on distinct handles) are safe and are issued in lock-acquisition order.
std::recursive_mutex& api_mutex();
#define PX_API_GUARD std::lock_guard<std::recursive_mutex> _px_api_lock(pixie::api_mutex())
int new_image_uninit(int w, int h, int layers, int bpp, pixie_image_t* out);  // contents undefined (callers overwrite)
void set_error(const std::string& msg);
int fail_pixie(const std::string& msg);            // returns 1
int fail_cuda(cudaError_t e, const char* what);    // returns 2
int ensure_init();
Image* find_image(uint64_t h);
int get_scratch(int slot, size_t bytes, void** out);
int get_pinned(size_t bytes, void** out);
int staging_acquire(size_t bytes, void** out);  // pinned staging, safe to overwrite on return
int staging_release();                          // call right after enqueuing the copy that reads it

// path commands -> device-resident segments (flatten.cu); segs / wind come from the stream-ordered pool
struct FlattenedPaths {
  float4* segs = nullptr;
  int16_t* wind = nullptr;
  std::vector<int> segBegin;   // [numPaths + 1]
  std::vector<float> bounds;   // per path: xMin, xMax, yMin, yMax of its segments, NaN flag
  int numSegs = 0, numPoints = 0, numPrims = 0;
  size_t h2dBytes = 0;
};
int flatten_paths(int numPaths, const pixie_path_desc* descs, const float* commands, int64_t numCommandFloats, const float* rawXyxy,
                  const int16_t* rawWinding, int64_t numRaw, FlattenedPaths& out);
void free_flattened(FlattenedPaths& f);

#define PX_CUDA(call)                                  \
  do {                                                 \
    cudaError_t _e = (call);                           \
    if (_e != cudaSuccess) return pixie::fail_cuda(_e, #call); \
  } while (0)

#define PX_LAUNCHED()                                  \
  do {                                                 \
    pixie::rt().launches++;                            \
    cudaError_t _e = cudaGetLastError();               \
    if (_e != cudaSuccess) return pixie::fail_cuda(_e, "kernel launch"); \
  } while (0)

enum BlendMode {
  NormalBlend = 0, DarkenBlend, MultiplyBlend, ColorBurnBlend, LightenBlend, ScreenBlend,
  ColorDodgeBlend, OverlayBlend, SoftLightBlend, HardLightBlend, DifferenceBlend, ExclusionBlend,
  HueBlend, SaturationBlend, ColorBlend, LuminosityBlend, MaskBlend, OverwriteBlend,
  SubtractMaskBlend, ExcludeMaskBlend, NumBlendModes
};  // common.nim:6-29

// ------------------------------------------------------------------ device pixel helpers
typedef uint32_t px_t;  // r | g<<8 | b<<16 | a<<24
#define PXD __device__ __forceinline__

PXD uint32_t pR(px_t p) { return p & 255u; }
PXD uint32_t pG(px_t p) { return (p >> 8) & 255u; }
PXD uint32_t pB(px_t p) { return (p >> 16) & 255u; }
PXD uint32_t pA(px_t p) { return p >> 24; }
PXD px_t mk(uint32_t r, uint32_t g, uint32_t b, uint32_t a) {
  return (r & 255u) | ((g & 255u) << 8) | ((b & 255u) << 16) | (a << 24);
}

// two 16-bit lanes holding values <= 65025: floor(x / 255) per lane = (x + 1 + (x >> 8)) >> 8
PXD uint32_t div255x2(uint32_t x) {
  uint32_t t = x + 0x00010001u + ((x >> 8) & 0x00FF00FFu);
  return (t >> 8) & 0x00FF00FFu;
}
// floor(p * k / 255) on all four channels (k in 0..255)
PXD px_t mul_div255(px_t p, uint32_t k) {
  uint32_t rb = (p & 0x00FF00FFu) * k;
  uint32_t ga = ((p >> 8) & 0x00FF00FFu) * k;
  return div255x2(rb) | (div255x2(ga) << 8);
}
// byte-wise wrapping add (mm_add_epi8)
PXD px_t add_bytes(px_t a, px_t b) {
  uint32_t lo = (a & 0x00FF00FFu) + (b & 0x00FF00FFu);
  uint32_t hi = ((a >> 8) & 0x00FF00FFu) + ((b >> 8) & 0x00FF00FFu);
  return (lo & 0x00FF00FFu) | ((hi & 0x00FF00FFu) << 8);
}

// The x86 row-kernel bodies (sse2.nim:13-46): used wherever the reference calls blendLine*.
PXD px_t line_normal(px_t b, px_t s) { return add_bytes(s, mul_div255(b, 255u - pA(s))); }
PXD px_t line_mask(px_t b, px_t s) { return mul_div255(b, pA(s)); }
// applyCoverage (sse2.nim:510-524): floor(c * cov / 255)
PXD px_t mul_cov_floor(px_t c, uint32_t cov) { return mul_div255(c, cov); }
// ColorRGBX * uint8 (common.nim:79-90): rounding form, used by the generic modes (paths.nim:1519-1526)
PXD px_t mul_cov_round(px_t c, uint32_t cov) {
  if (cov == 0) return 0;
  if (cov == 255) return c;
  return mk((pR(c) * cov + 127u) / 255u, (pG(c) * cov + 127u) / 255u, (pB(c) * cov + 127u) / 255u,
            (pA(c) * cov + 127u) / 255u);
}
// applyOpacity(M128, area) (sse2.nim:6-11): cvtps_epi32 (round half even) + unsigned saturation
PXD uint32_t cvt_sat(float v) {
  float r = rintf(v);
  if (!(r > 0.0f)) return 0u;
  if (r > 255.0f) return 255u;
  return (uint32_t)r;
}
PXD px_t mul_area(px_t c, float area) {
  return mk(cvt_sat((float)pR(c) * area), cvt_sat((float)pG(c) * area), cvt_sat((float)pB(c) * area),
            cvt_sat((float)pA(c) * area));
}

// ------------------------------------------------------------------ blends.nim (scalar forms, blender())
PXD uint32_t blend_alpha(uint32_t ba, uint32_t sa) { return (sa + (ba * (255u - sa)) / 255u) & 255u; }  // :41-43
PXD uint32_t screen_(uint32_t b, uint32_t s) { return ((b + s) - (b * s) / 255u) & 255u; }             // :45-46
PXD uint32_t hard_light(uint32_t bc, uint32_t ba, uint32_t sc, uint32_t sa) {                          // :48-58
  if (sc * 2u <= sa) return ((2u * sc * bc + (sc * (255u - ba)) + (bc * (255u - sa))) / 255u) & 255u;
  return screen_(bc, sc);
}
PXD px_t blend_normal(px_t b, px_t s) {  // :60-70
  if (pA(b) == 0u || pA(s) == 255u) return s;
  if (pA(s) == 0u) return b;
  return line_normal(b, s);
}
// 256-entry tables of the two float divisions every un-premultiply / to-Color step makes: x / 255.0f and
// 255.0f / a are IEEE-exact quotients of small integers, so a table lookup returns the very same float as the
// division the reference performs (the divisions were most of the float blend modes' instructions).
struct BlendTables {
  float div255[256];   // (float)x / 255.0f
  float mul255[256];   // 255.0f / (float)a   (a = 0 unused)
};
constexpr BlendTables make_blend_tables() {
  BlendTables t{};
  for (int i = 0; i < 256; i++) {
    t.div255[i] = (float)i / 255.0f;
    t.mul255[i] = i ? 255.0f / (float)i : 0.0f;
  }
  return t;
}
__device__ const BlendTables g_blend_tables = make_blend_tables();

PXD uint32_t round_half_away_u(float v) {  // roundf() of 0 <= v < 2^22 as an integer: floor(v + 0.5), see round_bits
  return __float_as_uint(__fadd_rz(__fadd_rz(v, 0.5f), 8388608.0f)) - 0x4B000000u;
}
PXD uint32_t straight_(uint32_t c, uint32_t a) {  // internal.nim:68-74 stand-in for chroma rgba()
  if (a == 0u) return 0u;
  const float multiplier = __ldg(&g_blend_tables.mul255[a]);
  const uint32_t v = round_half_away_u((float)c * multiplier);
  return v > 255u ? 255u : v;
}
PXD px_t to_straight(px_t p) {
  uint32_t a = pA(p);
  return mk(straight_(pR(p), a), straight_(pG(p), a), straight_(pB(p), a), a);
}
PXD px_t to_premul(px_t p) {  // chroma rgbx(ColorRGBA)
  uint32_t a = pA(p);
  if (a == 255u) return p;
  return mk((pR(p) * a + 127u) / 255u, (pG(p) * a + 127u) / 255u, (pB(p) * a + 127u) / 255u, a);
}
PXD px_t alpha_fix(px_t backdrop, px_t source, px_t mixed) {  // blends.nim:18-39
  uint32_t sa = pA(source), ba = pA(backdrop);
  uint32_t t0 = sa * (255u - ba), t1 = sa * ba, t2 = (255u - sa) * ba;
  uint32_t r = t0 * pR(source) + t1 * pR(mixed) + t2 * pR(backdrop);
  uint32_t g = t0 * pG(source) + t1 * pG(mixed) + t2 * pG(backdrop);
  uint32_t b = t0 * pB(source) + t1 * pB(mixed) + t2 * pB(backdrop);
  uint32_t a = sa + ba * (255u - sa) / 255u;
  if (a == 0u) return 0u;
  return mk(r / a / 255u, g / a / 255u, b / a / 255u, a);
}

struct Col {
  float r, g, b, a;
};
// ---- the float blend modes' arithmetic, instruction by instruction what the reference computes, in fewer instructions:
// * roundf(v) of 0 <= v < 2^22: floor(v + 0.5) with BOTH additions rounded toward zero — RZ(t) >= n exactly when
//   t >= n for an integer n, so the truncated sum has the exact sum's floor; the second addition (2^23) leaves that
//   floor in the low mantissa bits (no float -> int conversion, no compare);
// * u / 255.0f for an integer 0 <= u <= 255: q = u * RN(1/255) corrected once with the exact FMA residual
//   (Markstein); equal to the IEEE quotient for all 256 inputs (tests/test_float_blend_identities.py);
// * x / d for three numerators that share d: one correctly rounded reciprocal, then per quotient the classical
//   two-step FMA refinement, which returns the correctly rounded quotient whenever nothing leaves the normal range —
//   denominators outside [1e-18, 1e18] (incl. 0, negatives, NaN) take the IEEE division itself.
PXD float round_bits(float v) { return __fadd_rz(__fadd_rz(v, 0.5f), 8388608.0f); }  // 2^23 + roundf(v) for 0 <= v < 2^22
PXD float div255f(float u) {
  const float r = 1.0f / 255.0f;
  const float q = u * r;
  return __fmaf_rn(__fmaf_rn(-q, 255.0f, u), r, q);
}
PXD float fdiv_r(float x, float d, float r) {  // RN(x / d) given r = RN(1 / d)
  float q = x * r;
  q = __fmaf_rn(__fmaf_rn(-q, d, x), r, q);
  return __fmaf_rn(__fmaf_rn(-q, d, x), r, q);
}
PXD void div3(float& x, float& y, float& z, float d) {  // x / d, y / d, z / d
  if (d > 1e-18f && d < 1e18f) {
    const float r = __frcp_rn(d);
    x = fdiv_r(x, d, r);
    y = fdiv_r(y, d, r);
    z = fdiv_r(z, d, r);
  } else {
    x = x / d;
    y = y / d;
    z = z / d;
  }
}
// to_color: chroma rgba() (straightAlphaTable stand-in, straight_ above) then .color (x / 255)
PXD Col to_color(px_t p) {
  const uint32_t a = pA(p);
  const float multiplier = __ldg(&g_blend_tables.mul255[a]);  // 0 for a == 0: every channel becomes 0 as in straight_
  Col c;
  c.r = div255f(fminf(round_bits((float)pR(p) * multiplier) - 8388608.0f, 255.0f));
  c.g = div255f(fminf(round_bits((float)pG(p) * multiplier) - 8388608.0f, 255.0f));
  c.b = div255f(fminf(round_bits((float)pB(p) * multiplier) - 8388608.0f, 255.0f));
  c.a = div255f((float)a);
  return c;
}
PXD uint32_t f2u8(float v) {  // roundf(v * 255) clamped to 0..255 (NaN and negatives -> 0)
  const float x = fminf(fmaxf(v * 255.0f, 0.0f), 255.0f);  // fmaxf(NaN, 0) = 0
  return __float_as_uint(round_bits(x)) & 0x1FFu;
}
PXD px_t from_color(Col c) { return to_premul(mk(f2u8(c.r), f2u8(c.g), f2u8(c.b), f2u8(c.a))); }
PXD float min3f(float a, float b, float c) { return fminf(a, fminf(b, c)); }
PXD float max3f(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }
PXD float lum(Col c) { return 0.3f * c.r + 0.59f * c.g + 0.11f * c.b; }
PXD Col clip_color(Col c) {
  float L = lum(c), n = min3f(c.r, c.g, c.b), x = max3f(c.r, c.g, c.b);
  if (n < 0) {
    float r_ = (c.r - L) * L, g_ = (c.g - L) * L, b_ = (c.b - L) * L;
    div3(r_, g_, b_, L - n);
    c.r = L + r_;
    c.g = L + g_;
    c.b = L + b_;
  }
  if (x > 1) {
    float r_ = (c.r - L) * (1 - L), g_ = (c.g - L) * (1 - L), b_ = (c.b - L) * (1 - L);
    div3(r_, g_, b_, x - L);
    c.r = L + r_;
    c.g = L + g_;
    c.b = L + b_;
  }
  return c;
}
PXD Col set_lum(Col c, float l) {
  float d = l - lum(c);
  c.r += d;
  c.g += d;
  c.b += d;
  return clip_color(c);
}
PXD float sat(Col c) { return max3f(c.r, c.g, c.b) - min3f(c.r, c.g, c.b); }
PXD Col set_sat(Col c, float s) {
  float satC = sat(c);
  Col r = {0.f, 0.f, 0.f, c.a};
  if (satC > 0) {
    float mn = min3f(c.r, c.g, c.b);
    r.r = (c.r - mn) * s;
    r.g = (c.g - mn) * s;
    r.b = (c.b - mn) * s;
    div3(r.r, r.g, r.b, satC);
  }
  return r;
}
PXD Col alpha_fix_f(Col cb, Col cs, Col mixed) {
  Col r;
  r.a = cs.a + cb.a * (1.0f - cs.a);
  if (r.a == 0) {
    r.r = r.g = r.b = 0.f;
    return r;
  }
  float t0 = cs.a * (1 - cb.a), t1 = cs.a * cb.a, t2 = (1 - cs.a) * cb.a;
  r.r = t0 * cs.r + t1 * mixed.r + t2 * cb.r;
  r.g = t0 * cs.g + t1 * mixed.g + t2 * cb.g;
  r.b = t0 * cs.b + t1 * mixed.b + t2 * cb.b;
  div3(r.r, r.g, r.b, r.a);
  return r;
}

// ---- the same five modes for TWO pixels at a time on Blackwell's packed fp32 instructions (add / mul / fma .f32x2 =
// FADD2 / FMUL2 / FFMA2: one issue slot for both pixels).  Every packed operation is the IEEE operation of its two
// halves, so the pair function returns exactly what two calls of the scalar path return; the kernels of blend.cu use
// it for the 16-byte groups of their fast path.
struct F2 {
  unsigned long long v;
};
PXD F2 f2(float x, float y) {  // the two halves of a 64-bit register are ordinary registers: no instruction
  F2 r;
  r.v = ((unsigned long long)__float_as_uint(y) << 32) | (unsigned long long)__float_as_uint(x);
  return r;
}
PXD float f2lo(F2 a) { return __uint_as_float((uint32_t)a.v); }
PXD float f2hi(F2 a) { return __uint_as_float((uint32_t)(a.v >> 32)); }
PXD F2 splat2(float x) { return f2(x, x); }
PXD F2 add2(F2 a, F2 b) {
  F2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
PXD F2 add2rz(F2 a, F2 b) {
  F2 r;
  asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
PXD F2 mul2(F2 a, F2 b) {
  F2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
PXD F2 fma2(F2 a, F2 b, F2 c) {
  F2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
PXD F2 sub2(F2 a, F2 b) { return fma2(b, splat2(-1.0f), a); }  // RN(a - b): b * -1 is exact
// A sum whose operand is a packed PRODUCT goes through two scalar additions: ptxas (CUDA 12.9) contracts
// mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under --fmad=false and despite the explicit .rn (it also folds
// fma(a, b, -0) and fma(p, 1, c) first), which changes the last bit about once in 10^5 pixels; scalar FADDs are
// left alone and cost no moves (the halves of a 64-bit register are ordinary registers).
PXD F2 add2s(F2 a, F2 b) { return f2(__fadd_rn(f2lo(a), f2lo(b)), __fadd_rn(f2hi(a), f2hi(b))); }
PXD F2 add2rzs(F2 a, float b) { return f2(__fadd_rz(f2lo(a), b), __fadd_rz(f2hi(a), b)); }
PXD F2 sel2(bool c0, bool c1, F2 a, F2 b) { return f2(c0 ? f2lo(a) : f2lo(b), c1 ? f2hi(a) : f2hi(b)); }
PXD F2 round_bits2(F2 v) { return add2rz(add2rzs(v, 0.5f), splat2(8388608.0f)); }  // v may be a packed product: scalar first addition
PXD F2 div255f2(F2 u) {
  const F2 r = splat2(1.0f / 255.0f);
  const F2 q = mul2(u, r);
  return fma2(fma2(q, splat2(-255.0f), u), r, q);
}
struct Col2 {
  F2 r, g, b, a;
};
// three quotients per pixel over a shared denominator: d must be inside div3's fast range in both halves
struct Rcp2 {
  F2 r, nd;
};
PXD Rcp2 rcp2(float d0, float d1) {
  Rcp2 q;
  q.r = f2(__frcp_rn(d0), __frcp_rn(d1));
  q.nd = f2(-d0, -d1);
  return q;
}
PXD F2 fdiv2(F2 x, const Rcp2& k) {
  F2 q = mul2(x, k.r);
  q = fma2(fma2(q, k.nd, x), k.r, q);
  return fma2(fma2(q, k.nd, x), k.r, q);
}
PXD bool div_fast_ok(float d) { return d > 1e-18f && d < 1e18f; }
PXD Col2 to_color2(px_t p0, px_t p1) {
  const F2 M = f2(__ldg(&g_blend_tables.mul255[p0 >> 24]), __ldg(&g_blend_tables.mul255[p1 >> 24]));
  const F2 nbig = splat2(-8388608.0f);
  Col2 c;
  F2* out[3] = {&c.r, &c.g, &c.b};
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    // byte ch of the pixel in the mantissa of 2^23: the integer as a float after one subtraction
    const F2 z = f2(__uint_as_float(__byte_perm(p0, 0x4B000000u, 0x7650u + ch)), __uint_as_float(__byte_perm(p1, 0x4B000000u, 0x7650u + ch)));
    const F2 rb = add2(round_bits2(mul2(add2(z, nbig), M)), nbig);  // roundf(c * multiplier)
    *out[ch] = div255f2(f2(fminf(f2lo(rb), 255.0f), fminf(f2hi(rb), 255.0f)));
  }
  const F2 za = f2(__uint_as_float(__byte_perm(p0, 0x4B000000u, 0x7653u)), __uint_as_float(__byte_perm(p1, 0x4B000000u, 0x7653u)));
  c.a = div255f2(add2(za, nbig));
  return c;
}
PXD void f2u8_2(F2 v, uint32_t& n0, uint32_t& n1) {
  const F2 x = mul2(v, splat2(255.0f));
  const F2 rb = round_bits2(f2(fminf(fmaxf(f2lo(x), 0.0f), 255.0f), fminf(fmaxf(f2hi(x), 0.0f), 255.0f)));
  n0 = __float_as_uint(f2lo(rb)) & 0x1FFu;
  n1 = __float_as_uint(f2hi(rb)) & 0x1FFu;
}
PXD F2 lum2(const Col2& c) { return add2s(add2s(mul2(splat2(0.3f), c.r), mul2(splat2(0.59f), c.g)), mul2(splat2(0.11f), c.b)); }
PXD F2 min3f2(const Col2& c) { return f2(min3f(f2lo(c.r), f2lo(c.g), f2lo(c.b)), min3f(f2hi(c.r), f2hi(c.g), f2hi(c.b))); }
PXD F2 max3f2(const Col2& c) { return f2(max3f(f2lo(c.r), f2lo(c.g), f2lo(c.b)), max3f(f2hi(c.r), f2hi(c.g), f2hi(c.b))); }
// c + (c - L) * k / d per channel where `on`, unchanged elsewhere; false when an active half's d is outside the fast range
PXD bool clip_step2(Col2& c, F2 L, F2 k, F2 d, bool on0, bool on1) {
  const float d0 = on0 ? f2lo(d) : 1.0f, d1 = on1 ? f2hi(d) : 1.0f;
  if (!(div_fast_ok(d0) && div_fast_ok(d1))) return false;
  const Rcp2 q = rcp2(d0, d1);
  c.r = sel2(on0, on1, add2(L, fdiv2(mul2(sub2(c.r, L), k), q)), c.r);
  c.g = sel2(on0, on1, add2(L, fdiv2(mul2(sub2(c.g, L), k), q)), c.g);
  c.b = sel2(on0, on1, add2(L, fdiv2(mul2(sub2(c.b, L), k), q)), c.b);
  return true;
}
PXD bool clip_color2(Col2& c) {
  const F2 L = lum2(c), n = min3f2(c), x = max3f2(c);
  bool ok = true;
  const bool n0 = f2lo(n) < 0, n1 = f2hi(n) < 0;
  if (n0 || n1) ok = clip_step2(c, L, L, sub2(L, n), n0, n1);
  const bool x0 = f2lo(x) > 1, x1 = f2hi(x) > 1;
  if (x0 || x1) ok = clip_step2(c, L, sub2(splat2(1.0f), L), sub2(x, L), x0, x1) && ok;
  return ok;
}
PXD bool set_lum2(Col2& c, F2 l) {
  const F2 d = sub2(l, lum2(c));
  c.r = add2(c.r, d);
  c.g = add2(c.g, d);
  c.b = add2(c.b, d);
  return clip_color2(c);
}
PXD F2 sat2(const Col2& c) { return sub2(max3f2(c), min3f2(c)); }
PXD bool set_sat2(Col2& c, F2 s) {
  const F2 satC = sat2(c), mn = min3f2(c);
  const bool on0 = f2lo(satC) > 0, on1 = f2hi(satC) > 0;
  const float d0 = on0 ? f2lo(satC) : 1.0f, d1 = on1 ? f2hi(satC) : 1.0f;
  if (!(div_fast_ok(d0) && div_fast_ok(d1))) return false;
  const Rcp2 q = rcp2(d0, d1);
  const F2 zero = splat2(0.0f);
  c.r = sel2(on0, on1, fdiv2(mul2(sub2(c.r, mn), s), q), zero);
  c.g = sel2(on0, on1, fdiv2(mul2(sub2(c.g, mn), s), q), zero);
  c.b = sel2(on0, on1, fdiv2(mul2(sub2(c.b, mn), s), q), zero);
  return true;
}
template <int MODE>
PXD px_t blend_px(px_t b, px_t s);
template <int MODE>
PXD void blend_px2_float(px_t b0, px_t b1, px_t s0, px_t s1, px_t& o0, px_t& o1) {
  const Col2 cb = to_color2(b0, b1), cs = to_color2(s0, s1);
  Col2 m;
  bool ok = true;
  if (MODE == SoftLightBlend) {
    const F2 one = splat2(1.0f), m2 = splat2(-2.0f);
    m.r = add2s(mul2(fma2(cs.r, m2, one), mul2(cb.r, cb.r)), mul2(add2(cs.r, cs.r), cb.r));
    m.g = add2s(mul2(fma2(cs.g, m2, one), mul2(cb.g, cb.g)), mul2(add2(cs.g, cs.g), cb.g));
    m.b = add2s(mul2(fma2(cs.b, m2, one), mul2(cb.b, cb.b)), mul2(add2(cs.b, cs.b), cb.b));
  } else if (MODE == HueBlend) {
    m = cs;
    ok = set_sat2(m, sat2(cb));
    ok = set_lum2(m, lum2(cb)) && ok;
  } else if (MODE == SaturationBlend) {
    m = cb;
    ok = set_sat2(m, sat2(cs));
    ok = set_lum2(m, lum2(cb)) && ok;
  } else if (MODE == ColorBlend) {
    m = cs;
    ok = set_lum2(m, lum2(cb));
  } else {
    m = cb;
    ok = set_lum2(m, lum2(cs));
  }
  if (!ok) {  // a denominator outside the refinement's range: the scalar path, which then divides the IEEE way
    o0 = blend_px<MODE>(b0, s0);
    o1 = blend_px<MODE>(b1, s1);
    return;
  }
  // alpha_fix_f: r.a == 0 only when both alphas are 0, then t0 = t1 = t2 = 0 and every numerator is +0: dividing by
  // the substitute 1e-10 leaves the 0 the reference returns
  const F2 one = splat2(1.0f);
  const F2 omsa = sub2(one, cs.a), omba = sub2(one, cb.a);
  const F2 ra = add2s(cs.a, mul2(cb.a, omsa));
  const F2 t0 = mul2(cs.a, omba), t1 = mul2(cs.a, cb.a), t2 = mul2(omsa, cb.a);
  const Rcp2 q = rcp2(fmaxf(f2lo(ra), 1e-10f), fmaxf(f2hi(ra), 1e-10f));  // 1 / 65025 <= r.a <= 2 otherwise
  const F2 rr = fdiv2(add2s(add2s(mul2(t0, cs.r), mul2(t1, m.r)), mul2(t2, cb.r)), q);
  const F2 rg = fdiv2(add2s(add2s(mul2(t0, cs.g), mul2(t1, m.g)), mul2(t2, cb.g)), q);
  const F2 rbl = fdiv2(add2s(add2s(mul2(t0, cs.b), mul2(t1, m.b)), mul2(t2, cb.b)), q);
  uint32_t r0, r1, g0, g1, bb0, bb1, a0, a1;
  f2u8_2(rr, r0, r1);
  f2u8_2(rg, g0, g1);
  f2u8_2(rbl, bb0, bb1);
  f2u8_2(ra, a0, a1);
  o0 = to_premul(mk(r0, g0, bb0, a0));
  o1 = to_premul(mk(r1, g1, bb1, a1));
}
__host__ __device__ constexpr bool mode_is_float(int mode) {
  return mode == SoftLightBlend || mode == HueBlend || mode == SaturationBlend || mode == ColorBlend || mode == LuminosityBlend;
}
// which of them take the two-pixel path in blend.cu: measured at 8192^2, SoftLight gains 8 % (0.505 -> 0.463 ms); the four
// non-separable modes do not (their per-pixel clip / saturation branches become selects over both halves)
#ifndef PIXIE_PACKED_NONSEP
#define PIXIE_PACKED_NONSEP 0
#endif
__host__ __device__ constexpr bool mode_is_packed(int mode) { return mode == SoftLightBlend || (PIXIE_PACKED_NONSEP && mode_is_float(mode)); }

// blender(MODE)(backdrop, source), blends.nim:275-299 — MODE is a compile-time constant
template <int MODE>
PXD px_t blend_px(px_t b, px_t s) {
  if (MODE == NormalBlend) return blend_normal(b, s);
  if (MODE == OverwriteBlend) return s;
  if (MODE == MaskBlend) return mul_div255(b, pA(s));
  const uint32_t ba = pA(b), sa = pA(s);
  if (MODE == SubtractMaskBlend) {
    uint32_t a = (ba * (255u - sa)) / 255u;
    return (mul_div255(b, a) & 0x00FFFFFFu) | (a << 24);
  }
  if (MODE == ExcludeMaskBlend) {
    uint32_t a = max(ba, sa) - min(ba, sa);
    return (mul_div255(s, a) & 0x00FFFFFFu) | (a << 24);
  }
  if (MODE == ColorBurnBlend || MODE == ColorDodgeBlend) {
    px_t bd = to_straight(b), sr = to_straight(s);
    uint32_t o[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      uint32_t bc = (bd >> (8 * c)) & 255u, sc = (sr >> (8 * c)) & 255u;
      if (MODE == ColorBurnBlend) {
        if (bc == 255u) o[c] = 255u;
        else if (sc == 0u) o[c] = 0u;
        else o[c] = 255u - (min(255u, (255u * (255u - bc)) / sc) & 255u);
      } else {
        if (bc == 0u) o[c] = 0u;
        else if (sc == 255u) o[c] = 255u;
        else o[c] = min(255u, (255u * bc) / (255u - sc));
      }
    }
    return to_premul(alpha_fix(bd, sr, mk(o[0], o[1], o[2], 0u)));
  }
  if (MODE == SoftLightBlend || MODE == HueBlend || MODE == SaturationBlend || MODE == ColorBlend ||
      MODE == LuminosityBlend) {
    Col cb = to_color(b), cs = to_color(s), m = {0.f, 0.f, 0.f, 0.f};
    if (MODE == SoftLightBlend) {
      m.r = (1 - 2 * cs.r) * (cb.r * cb.r) + 2 * cs.r * cb.r;
      m.g = (1 - 2 * cs.g) * (cb.g * cb.g) + 2 * cs.g * cb.g;
      m.b = (1 - 2 * cs.b) * (cb.b * cb.b) + 2 * cs.b * cb.b;
    } else if (MODE == HueBlend) {
      m = set_lum(set_sat(cs, sat(cb)), lum(cb));
    } else if (MODE == SaturationBlend) {
      m = set_lum(set_sat(cb, sat(cs)), lum(cb));
    } else if (MODE == ColorBlend) {
      m = set_lum(cs, lum(cb));
    } else {
      m = set_lum(cb, lum(cs));
    }
    return from_color(alpha_fix_f(cb, cs, m));
  }
  // separable integer modes: per channel f(bc, ba, sc, sa), alpha = blendAlpha
  uint32_t o[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    uint32_t bc = (b >> (8 * c)) & 255u, sc = (s >> (8 * c)) & 255u;
    uint32_t v = 0;
    if (MODE == DarkenBlend) v = min(bc + ((255u - ba) * sc) / 255u, sc + ((255u - sa) * bc) / 255u);
    else if (MODE == LightenBlend) v = max(bc + ((255u - ba) * sc) / 255u, sc + ((255u - sa) * bc) / 255u);
    else if (MODE == MultiplyBlend) v = ((255u - ba) * sc + (255u - sa) * bc + bc * sc) / 255u;
    else if (MODE == ScreenBlend) v = screen_(bc, sc);
    else if (MODE == OverlayBlend) v = hard_light(sc, sa, bc, ba);
    else if (MODE == HardLightBlend) v = hard_light(bc, ba, sc, sa);
    else if (MODE == DifferenceBlend) v = (bc + sc) - 2u * (min(bc * sa, sc * ba) / 255u);
    else if (MODE == ExclusionBlend) {
      int32_t t = (int32_t)(bc + sc) - (int32_t)((2u * bc * sc) / 255u);
      v = (uint32_t)(t < 0 ? 0 : t);
    }
    o[c] = v;
  }
  return mk(o[0], o[1], o[2], blend_alpha(ba, sa));
}

// ---- table forms of the seven ALU-bound modes (blend.cu stages the tables in shared memory) -----------------
// ColorBurn / ColorDodge spend their time in to_straight's float multiply + round per channel and in integer
// divisions by a variable (burn / dodge quotient, alpha_fix's `div a`); the five float modes in to_straight + the
// x / 255 conversions.  straight[a << 8 | c] holds straight_(c, a) for every (a, c) — built on the device by that very
// function, so the lookup IS the reference value; inv[d] = ceil(2^32 / d) turns floor(x / d) into one multiply-high,
// exact whenever x * d < 2^32 (error term x * (inv[d] - 2^32 / d) / 2^32 < 1 / d, and frac(x / d) <= 1 - 1 / d).
struct BlendTab {
  const uint8_t* straight;  // [65536]
  const uint32_t* inv;      // [256], inv[0] = inv[1] = 0 (d == 1 is selected around)
  const float* div255;      // [256]
};
struct InvTable {
  uint32_t v[256];
};
constexpr InvTable make_inv_table() {
  InvTable t{};
  for (uint32_t d = 2; d < 256; d++) t.v[d] = (uint32_t)((0x100000000ull + d - 1) / d);
  return t;
}
__device__ const InvTable g_inv_table = make_inv_table();

PXD uint32_t udiv_tab(uint32_t x, uint32_t d, const uint32_t* inv) {  // floor(x / d): 1 <= d <= 255, x * d < 2^32
  const uint32_t q = __umulhi(x, inv[d]);
  return d == 1u ? x : q;
}
PXD px_t to_straight_tab(px_t p, const uint8_t* st) {
  const uint32_t a = pA(p);
  const uint8_t* row = st + (a << 8);
  return (uint32_t)row[pR(p)] | ((uint32_t)row[pG(p)] << 8) | ((uint32_t)row[pB(p)] << 16) | (a << 24);
}
PXD px_t alpha_fix_tab(px_t backdrop, px_t source, px_t mixed, const uint32_t* inv) {  // alpha_fix with `div a` from the table
  const uint32_t sa = pA(source), ba = pA(backdrop);
  const uint32_t t0 = sa * (255u - ba), t1 = sa * ba, t2 = (255u - sa) * ba;
  const uint32_t a = sa + ba * (255u - sa) / 255u;
  if (a == 0u) return 0u;
  // t0 + t1 + t2 <= 65025, so every sum below is < 2^24 and sum * a < 2^32
  const uint32_t r = t0 * pR(source) + t1 * pR(mixed) + t2 * pR(backdrop);
  const uint32_t g = t0 * pG(source) + t1 * pG(mixed) + t2 * pG(backdrop);
  const uint32_t b = t0 * pB(source) + t1 * pB(mixed) + t2 * pB(backdrop);
  return mk(udiv_tab(r, a, inv) / 255u, udiv_tab(g, a, inv) / 255u, udiv_tab(b, a, inv) / 255u, a);
}
PXD Col to_color_tab(px_t p, const BlendTab& T) {
  const px_t s = to_straight_tab(p, T.straight);
  Col c = {T.div255[pR(s)], T.div255[pG(s)], T.div255[pB(s)], T.div255[pA(s)]};
  return c;
}
// blend_px<MODE> for ColorBurn, ColorDodge, SoftLight, Hue, Saturation, Color, Luminosity — same values, tables in `T`
template <int MODE>
PXD px_t blend_px_tab(px_t b, px_t s, const BlendTab& T) {
  if (MODE == ColorBurnBlend || MODE == ColorDodgeBlend) {
    const px_t bd = to_straight_tab(b, T.straight), sr = to_straight_tab(s, T.straight);
    uint32_t o[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const uint32_t bc = (bd >> (8 * c)) & 255u, sc = (sr >> (8 * c)) & 255u;
      if (MODE == ColorBurnBlend) {
        const uint32_t q = udiv_tab(255u * (255u - bc), sc, T.inv);  // sc == 0: inv[0] = 0, the value is not used
        o[c] = bc == 255u ? 255u : (sc == 0u ? 0u : 255u - (min(255u, q) & 255u));
      } else {
        const uint32_t q = udiv_tab(255u * bc, 255u - sc, T.inv);
        o[c] = bc == 0u ? 0u : (sc == 255u ? 255u : min(255u, q));
      }
    }
    return to_premul(alpha_fix_tab(bd, sr, mk(o[0], o[1], o[2], 0u), T.inv));
  }
  Col cb = to_color_tab(b, T), cs = to_color_tab(s, T), m = {0.f, 0.f, 0.f, 0.f};
  if (MODE == SoftLightBlend) {
    m.r = (1 - 2 * cs.r) * (cb.r * cb.r) + 2 * cs.r * cb.r;
    m.g = (1 - 2 * cs.g) * (cb.g * cb.g) + 2 * cs.g * cb.g;
    m.b = (1 - 2 * cs.b) * (cb.b * cb.b) + 2 * cs.b * cb.b;
  } else if (MODE == HueBlend) {
    m = set_lum(set_sat(cs, sat(cb)), lum(cb));
  } else if (MODE == SaturationBlend) {
    m = set_lum(set_sat(cb, sat(cs)), lum(cb));
  } else if (MODE == ColorBlend) {
    m = set_lum(cs, lum(cb));
  } else {
    m = set_lum(cb, lum(cs));
  }
  return from_color(alpha_fix_f(cb, cs, m));
}
// Measured at 8192^2 with an A8 mask: ColorBurn 0.787 -> 0.473 ms, ColorDodge 0.754 -> 0.399 ms.  The five float modes
// were tried too and are NOT switched over: their time is in the IEEE divisions of clip_color / set_sat / alpha_fix_f,
// which a table cannot replace, and the 66 KB of tables cost them a resident CTA (SoftLight 0.815 -> 0.838 ms,
// Hue 1.54 -> 1.65 ms).
__host__ __device__ constexpr bool mode_uses_tables(int mode) { return mode == ColorBurnBlend || mode == ColorDodgeBlend; }

// run-time mode -> compile-time dispatch (mode is warp-uniform at every call site)
#define PX_DISPATCH_MODE(mode, EXPR)                                                          \
  switch (mode) {                                                                             \
    case 0: { constexpr int MODE = 0; EXPR; } break;                                          \
    case 1: { constexpr int MODE = 1; EXPR; } break;                                          \
    case 2: { constexpr int MODE = 2; EXPR; } break;                                          \
    case 3: { constexpr int MODE = 3; EXPR; } break;                                          \
    case 4: { constexpr int MODE = 4; EXPR; } break;                                          \
    case 5: { constexpr int MODE = 5; EXPR; } break;                                          \
    case 6: { constexpr int MODE = 6; EXPR; } break;                                          \
    case 7: { constexpr int MODE = 7; EXPR; } break;                                          \
    case 8: { constexpr int MODE = 8; EXPR; } break;                                          \
    case 9: { constexpr int MODE = 9; EXPR; } break;                                          \
    case 10: { constexpr int MODE = 10; EXPR; } break;                                        \
    case 11: { constexpr int MODE = 11; EXPR; } break;                                        \
    case 12: { constexpr int MODE = 12; EXPR; } break;                                        \
    case 13: { constexpr int MODE = 13; EXPR; } break;                                        \
    case 14: { constexpr int MODE = 14; EXPR; } break;                                        \
    case 15: { constexpr int MODE = 15; EXPR; } break;                                        \
    case 16: { constexpr int MODE = 16; EXPR; } break;                                        \
    case 17: { constexpr int MODE = 17; EXPR; } break;                                        \
    case 18: { constexpr int MODE = 18; EXPR; } break;                                        \
    default: { constexpr int MODE = 19; EXPR; } break;                                        \
  }

}  // namespace pixie

// CPU oracle for the raster hot path — TEST INFRASTRUCTURE ONLY (see pixie_oracle.h).
// Restates treeform/pixie's rasteriser / blends / blur on the CPU; every function cites the
// reference lines it follows.  IEEE float32, one rounding per op: build with -ffp-contract=off.
#include "pixie_oracle.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

typedef uint32_t px_t;  // ColorRGBX packed little-endian: r | g<<8 | b<<16 | a<<24
inline uint32_t R(px_t p) { return p & 255; }
inline uint32_t G(px_t p) { return (p >> 8) & 255; }
inline uint32_t B(px_t p) { return (p >> 16) & 255; }
inline uint32_t A(px_t p) { return p >> 24; }
inline px_t mk(uint32_t r, uint32_t g, uint32_t b, uint32_t a) {
  return (r & 255) | ((g & 255) << 8) | ((b & 255) << 16) | ((a & 255) << 24);
}

enum {
  NormalBlend = 0, DarkenBlend, MultiplyBlend, ColorBurnBlend, LightenBlend, ScreenBlend,
  ColorDodgeBlend, OverlayBlend, SoftLightBlend, HardLightBlend, DifferenceBlend, ExclusionBlend,
  HueBlend, SaturationBlend, ColorBlend, LuminosityBlend, MaskBlend, OverwriteBlend,
  SubtractMaskBlend, ExcludeMaskBlend, NumBlendModes
};  // common.nim:6-29

// ------------------------------------------------------------------ blends.nim
inline uint32_t blendAlpha(uint32_t ba, uint32_t sa) {  // :41-43
  return (sa + ((ba * (255 - sa)) / 255)) & 255;
}
inline uint32_t screen_(uint32_t b, uint32_t s) {  // :45-46
  return (uint32_t)((int32_t)(b + s) - (int32_t)((b * s) / 255)) & 255;
}
inline uint32_t hardLight(uint32_t bc, uint32_t ba, uint32_t sc, uint32_t sa) {  // :48-58
  if (sc * 2 <= sa) return ((2 * sc * bc + (sc * (255 - ba)) + (bc * (255 - sa))) / 255) & 255;
  return screen_(bc, sc);
}
px_t blendNormal(px_t b, px_t s) {  // :60-70
  if (A(b) == 0 || A(s) == 255) return s;
  if (A(s) == 0) return b;
  uint32_t k = 255 - A(s);
  return mk(R(s) + (R(b) * k) / 255, G(s) + (G(b) * k) / 255, B(s) + (B(b) * k) / 255, blendAlpha(A(b), A(s)));
}
template <typename F>
inline px_t sep4(px_t b, px_t s, F f) {  // per-channel blend(bc, ba, sc, sa) + blendAlpha
  return mk(f(R(b), A(b), R(s), A(s)), f(G(b), A(b), G(s), A(s)), f(B(b), A(b), B(s), A(s)),
            blendAlpha(A(b), A(s)));
}
px_t blendDarken(px_t b, px_t s) {  // :72-84
  return sep4(b, s, [](uint32_t bc, uint32_t ba, uint32_t sc, uint32_t sa) {
    uint32_t x = bc + ((255 - ba) * sc) / 255, y = sc + ((255 - sa) * bc) / 255;
    return x < y ? x : y;
  });
}
px_t blendMultiply(px_t b, px_t s) {  // :86-99
  return sep4(b, s, [](uint32_t bc, uint32_t ba, uint32_t sc, uint32_t sa) {
    return ((255 - ba) * sc + (255 - sa) * bc + bc * sc) / 255;
  });
}
px_t blendLighten(px_t b, px_t s) {  // :128-140
  return sep4(b, s, [](uint32_t bc, uint32_t ba, uint32_t sc, uint32_t sa) {
    uint32_t x = bc + ((255 - ba) * sc) / 255, y = sc + ((255 - sa) * bc) / 255;
    return x > y ? x : y;
  });
}
px_t blendScreen(px_t b, px_t s) {  // :142-146
  return mk(screen_(R(b), R(s)), screen_(G(b), G(s)), screen_(B(b), B(s)), blendAlpha(A(b), A(s)));
}
px_t blendOverlay(px_t b, px_t s) {  // :175-179
  return mk(hardLight(R(s), A(s), R(b), A(b)), hardLight(G(s), A(s), G(b), A(b)),
            hardLight(B(s), A(s), B(b), A(b)), blendAlpha(A(b), A(s)));
}
px_t blendHardLight(px_t b, px_t s) {  // :184-188
  return mk(hardLight(R(b), A(b), R(s), A(s)), hardLight(G(b), A(b), G(s), A(s)),
            hardLight(B(b), A(b), B(s), A(s)), blendAlpha(A(b), A(s)));
}
px_t blendDifference(px_t b, px_t s) {  // :190-204
  return sep4(b, s, [](uint32_t bc, uint32_t ba, uint32_t sc, uint32_t sa) {
    uint32_t x = bc * sa, y = sc * ba;
    uint32_t m = x < y ? x : y;
    return (uint32_t)((int32_t)(bc + sc) - 2 * (int32_t)(m / 255));
  });
}
px_t blendExclusion(px_t b, px_t s) {  // :206-213
  auto f = [](uint32_t bc, uint32_t sc) {
    int32_t v = (int32_t)(bc + sc) - (int32_t)((2 * bc * sc) / 255);
    return (uint32_t)(v < 0 ? 0 : v);
  };
  return mk(f(R(b), R(s)), f(G(b), G(s)), f(B(b), B(s)), blendAlpha(A(b), A(s)));
}
px_t blendMask(px_t b, px_t s) {  // :227-232
  uint32_t k = A(s);
  return mk((R(b) * k) / 255, (G(b) * k) / 255, (B(b) * k) / 255, (A(b) * k) / 255);
}
px_t blendSubtractMask(px_t b, px_t s) {  // :234-239
  uint32_t a = (A(b) * (255 - A(s))) / 255;
  return mk((R(b) * a) / 255, (G(b) * a) / 255, (B(b) * a) / 255, a);
}
px_t blendExcludeMask(px_t b, px_t s) {  // :241-246
  uint32_t mx = A(b) > A(s) ? A(b) : A(s), mn = A(b) < A(s) ? A(b) : A(s);
  uint32_t a = mx - mn;
  return mk((R(s) * a) / 255, (G(s) * a) / 255, (B(s) * a) / 255, a);
}

// chroma rgba(ColorRGBX) stand-in = Pixie's straightAlphaTable (internal.nim:68-74). UNPINNED.
inline uint32_t straight(uint32_t c, uint32_t a) {
  if (a == 0) return 0;
  float multiplier = 255.0f / (float)a;
  float v = roundf((float)c * multiplier);
  return v > 255.0f ? 255u : (uint32_t)v;
}
inline px_t toStraight(px_t p) { return mk(straight(R(p), A(p)), straight(G(p), A(p)), straight(B(p), A(p)), A(p)); }
// chroma rgbx(ColorRGBA): (c*a + 127) div 255 — pinned by tests/test_images.nim:204-228.
inline px_t toPremul(px_t p) {
  uint32_t a = A(p);
  if (a == 255) return p;
  return mk((R(p) * a + 127) / 255, (G(p) * a + 127) / 255, (B(p) * a + 127) / 255, a);
}
px_t alphaFix(px_t backdrop, px_t source, px_t mixed) {  // blends.nim:18-39 (straight alpha in/out)
  uint32_t sa = A(source), ba = A(backdrop);
  uint32_t t0 = sa * (255 - ba), t1 = sa * ba, t2 = (255 - sa) * ba;
  uint32_t r = t0 * R(source) + t1 * R(mixed) + t2 * R(backdrop);
  uint32_t g = t0 * G(source) + t1 * G(mixed) + t2 * G(backdrop);
  uint32_t b = t0 * B(source) + t1 * B(mixed) + t2 * B(backdrop);
  uint32_t a = sa + ba * (255 - sa) / 255;
  if (a == 0) return 0;
  return mk(r / a / 255, g / a / 255, b / a / 255, a);
}
px_t blendColorBurn(px_t b, px_t s) {  // :111-126
  px_t bd = toStraight(b), sr = toStraight(s);
  auto f = [](uint32_t bc, uint32_t sc) -> uint32_t {
    if (bc == 255) return 255;
    if (sc == 0) return 0;
    uint32_t q = (255 * (255 - bc)) / sc;
    return 255 - ((q < 255 ? q : 255) & 255);
  };
  px_t blended = mk(f(R(bd), R(sr)), f(G(bd), G(sr)), f(B(bd), B(sr)), 0);
  return toPremul(alphaFix(bd, sr, blended));
}
px_t blendColorDodge(px_t b, px_t s) {  // :158-173
  px_t bd = toStraight(b), sr = toStraight(s);
  auto f = [](uint32_t bc, uint32_t sc) -> uint32_t {
    if (bc == 0) return 0;
    if (sc == 255) return 255;
    uint32_t q = (255 * bc) / (255 - sc);
    return q < 255 ? q : 255;
  };
  px_t blended = mk(f(R(bd), R(sr)), f(G(bd), G(sr)), f(B(bd), B(sr)), 0);
  return toPremul(alphaFix(bd, sr, blended));
}

// ---- chroma float blends (W3C compositing-1 non-separable modes + Pegtop soft light). UNPINNED.
struct Col {
  float r, g, b, a;
};
inline Col toColor(px_t p) {
  px_t s = toStraight(p);
  Col c = {(float)R(s) / 255.0f, (float)G(s) / 255.0f, (float)B(s) / 255.0f, (float)A(s) / 255.0f};
  return c;
}
inline uint32_t f2u8(float v) {
  float x = roundf(v * 255.0f);
  if (!(x > 0.0f)) return 0;
  if (x > 255.0f) return 255;
  return (uint32_t)x;
}
inline px_t fromColor(Col c) { return toPremul(mk(f2u8(c.r), f2u8(c.g), f2u8(c.b), f2u8(c.a))); }
inline float min3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
inline float max3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }
inline float Lum(Col c) { return 0.3f * c.r + 0.59f * c.g + 0.11f * c.b; }
inline Col ClipColor(Col c) {
  float L = Lum(c), n = min3(c.r, c.g, c.b), x = max3(c.r, c.g, c.b);
  if (n < 0) {
    c.r = L + (((c.r - L) * L) / (L - n));
    c.g = L + (((c.g - L) * L) / (L - n));
    c.b = L + (((c.b - L) * L) / (L - n));
  }
  if (x > 1) {
    c.r = L + (((c.r - L) * (1 - L)) / (x - L));
    c.g = L + (((c.g - L) * (1 - L)) / (x - L));
    c.b = L + (((c.b - L) * (1 - L)) / (x - L));
  }
  return c;
}
inline Col SetLum(Col c, float l) {
  float d = l - Lum(c);
  c.r += d;
  c.g += d;
  c.b += d;
  return ClipColor(c);
}
inline float Sat(Col c) { return max3(c.r, c.g, c.b) - min3(c.r, c.g, c.b); }
inline Col SetSat(Col c, float s) {
  float satC = Sat(c);
  Col r = {0, 0, 0, c.a};
  if (satC > 0) {
    float mn = min3(c.r, c.g, c.b);
    r.r = (c.r - mn) * s / satC;
    r.g = (c.g - mn) * s / satC;
    r.b = (c.b - mn) * s / satC;
  }
  return r;
}
inline Col alphaFixF(Col cb, Col cs, Col mixed) {
  Col r;
  r.a = cs.a + cb.a * (1.0f - cs.a);
  if (r.a == 0) {
    r.r = r.g = r.b = 0;
    return r;
  }
  float t0 = cs.a * (1 - cb.a), t1 = cs.a * cb.a, t2 = (1 - cs.a) * cb.a;
  r.r = (t0 * cs.r + t1 * mixed.r + t2 * cb.r) / r.a;
  r.g = (t0 * cs.g + t1 * mixed.g + t2 * cb.g) / r.a;
  r.b = (t0 * cs.b + t1 * mixed.b + t2 * cb.b) / r.a;
  return r;
}
px_t blendFloatMode(int mode, px_t b, px_t s) {
  Col cb = toColor(b), cs = toColor(s), m = {0, 0, 0, 0};
  switch (mode) {
    case SoftLightBlend: {
      auto f = [](float bd, float sr) { return (1 - 2 * sr) * (bd * bd) + 2 * sr * bd; };
      m.r = f(cb.r, cs.r);
      m.g = f(cb.g, cs.g);
      m.b = f(cb.b, cs.b);
    } break;
    case HueBlend: m = SetLum(SetSat(cs, Sat(cb)), Lum(cb)); break;
    case SaturationBlend: m = SetLum(SetSat(cb, Sat(cs)), Lum(cb)); break;
    case ColorBlend: m = SetLum(cs, Lum(cb)); break;
    case LuminosityBlend: m = SetLum(cb, Lum(cs)); break;
  }
  return fromColor(alphaFixF(cb, cs, m));
}

px_t blendPx(int mode, px_t b, px_t s) {  // blender(), blends.nim:275-299
  switch (mode) {
    case NormalBlend: return blendNormal(b, s);
    case DarkenBlend: return blendDarken(b, s);
    case MultiplyBlend: return blendMultiply(b, s);
    case ColorBurnBlend: return blendColorBurn(b, s);
    case LightenBlend: return blendLighten(b, s);
    case ScreenBlend: return blendScreen(b, s);
    case ColorDodgeBlend: return blendColorDodge(b, s);
    case OverlayBlend: return blendOverlay(b, s);
    case HardLightBlend: return blendHardLight(b, s);
    case DifferenceBlend: return blendDifference(b, s);
    case ExclusionBlend: return blendExclusion(b, s);
    case SoftLightBlend: case HueBlend: case SaturationBlend: case ColorBlend: case LuminosityBlend:
      return blendFloatMode(mode, b, s);
    case MaskBlend: return blendMask(b, s);
    case OverwriteBlend: return s;
    case SubtractMaskBlend: return blendSubtractMask(b, s);
    case ExcludeMaskBlend: return blendExcludeMask(b, s);
  }
  return b;
}

// ---- the x86 row-kernel bodies (sse2.nim:13-46): floor roundings, byte-wrapping add
inline px_t lineNormal(px_t b, px_t s) {
  uint32_t k = 255 - A(s);
  return mk(R(s) + (R(b) * k) / 255, G(s) + (G(b) * k) / 255, B(s) + (B(b) * k) / 255, A(s) + (A(b) * k) / 255);
}
inline px_t lineMask(px_t b, px_t s) { return blendMask(b, s); }

// rgbx * coverage: scalar common.nim:79-90 (round) vs x86 applyCoverage sse2.nim:510-524 (floor)
inline px_t mulCovScalar(px_t c, uint32_t cov) {
  if (cov == 0) return 0;
  if (cov == 255) return c;
  return mk((R(c) * cov + 127) / 255, (G(c) * cov + 127) / 255, (B(c) * cov + 127) / 255, (A(c) * cov + 127) / 255);
}
inline px_t mulCovFloor(px_t c, uint32_t cov) {
  return mk((R(c) * cov) / 255, (G(c) * cov) / 255, (B(c) * cov) / 255, (A(c) * cov) / 255);
}
// rgbx * area: scalar common.nim:67-77 vs x86 applyOpacity(M128) sse2.nim:6-11
inline px_t mulAreaScalar(px_t c, float opacity) {
  if (opacity == 0) return 0;
  uint32_t x = (uint32_t)(int64_t)roundf(opacity * 255);
  return mk((R(c) * x + 127) / 255, (G(c) * x + 127) / 255, (B(c) * x + 127) / 255, (A(c) * x + 127) / 255);
}
inline uint32_t cvtSat(float v) {  // cvtps_epi32 (round half even) + packus
  float r = nearbyintf(v);
  if (!(r > 0.0f)) return 0;
  if (r > 255.0f) return 255;
  return (uint32_t)r;
}
inline px_t mulAreaSse(px_t c, float area) {
  return mk(cvtSat((float)R(c) * area), cvtSat((float)G(c) * area), cvtSat((float)B(c) * area),
            cvtSat((float)A(c) * area));
}

// ------------------------------------------------------------------ rasteriser
struct Entry {
  float ax, ay, bx, by;  // segment.at, segment.to
  float m, b;
  int16_t winding;
};
struct Partition {
  std::vector<Entry> entries;
  bool requiresAA = false, twoSpanning = false;
  int64_t top = 0, bottom = 0;
};

inline int64_t f2i(float f) {  // Nim float32 -> int (x86 cvttss2si semantics)
  if (!(f > -9.2e18f && f < 9.2e18f)) return INT64_MIN;
  return (int64_t)f;
}
inline int32_t fixed32(float f) {  // paths.nim:1268-1269
  float v = f * 256;
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT32_MIN;
  return (int32_t)v;
}

Entry initEntry(float ax, float ay, float bx, float by, int16_t w) {  // :1127-1135
  Entry e;
  e.ax = ax; e.ay = ay; e.bx = bx; e.by = by;
  e.winding = w;
  e.m = 0;
  e.b = 0;
  float d = ax - bx;
  if (d == 0) {
    e.b = ax;
  } else {
    e.m = (ay - by) / d;
    e.b = ay - e.m * ax;
  }
  return e;
}
inline float solveX(const Entry& e, float y) { return e.m == 0 ? e.b : (y - e.b) / e.m; }  // :1137-1141
inline float solveY(const Entry& e, float x) { return e.m * x + e.b; }                      // :1143-1144

inline bool segRequiresAA(const Entry& e) {  // :1149-1160
  auto frac = [](float v) { return v - truncf(v) != 0; };
  return e.ax != e.bx || frac(e.ax) || frac(e.ay) || frac(e.by);
}

// bumpy intersects(Segment, Line) with the horizontal line (0,y)-(1000,y)  
inline bool segLine(const Entry& s, float y, float& ox, float& oy) {
  float s1x = 1000.0f - 0.0f, s1y = y - y;
  float s2x = s.bx - s.ax, s2y = s.by - s.ay;
  float den = (-s2x * s1y + s1x * s2y);
  float num = s1x * (y - s.ay) - s1y * (0.0f - s.ax);
  float u = num / den;
  if (u >= 0 && u <= 1) {
    ox = s.ax + u * s2x;
    oy = s.ay + u * s2y;
    return true;
  }
  return false;
}
// internal.nim:36-48
inline bool intersectsInside(const Entry& a, const Entry& b) {
  float s1x = a.bx - a.ax, s1y = a.by - a.ay, s2x = b.bx - b.ax, s2y = b.by - b.ay;
  float den = (-s2x * s1y + s1x * s2y);
  float s = (-s1y * (a.ax - b.ax) + s1x * (a.ay - b.ay)) / den;
  float t = (s2x * (a.ay - b.ay) - s2y * (a.ax - b.ax)) / den;
  return s > 0 && s < 1 && t > 0 && t < 1;
}

void partitionSegments(const float* seg, const int16_t* wind, int n, int64_t top, int64_t height,
                       std::vector<Partition>& parts) {  // :1168-1262
  int64_t h4 = height / 4;
  uint32_t maxPartitions = (uint32_t)(h4 > 1 ? h4 : 1);
  int64_t n2 = n / 2;
  uint32_t numPartitions = (uint32_t)(n2 > 1 ? n2 : 1);
  if (maxPartitions < numPartitions) numPartitions = maxPartitions;
  parts.assign(numPartitions, Partition());
  uint32_t startY = (uint32_t)top;
  uint32_t partitionHeight = (uint32_t)height / numPartitions;
  parts[0].top = top;
  parts[0].bottom = top + (int64_t)partitionHeight;
  for (size_t i = 1; i < parts.size(); i++) {
    parts[i].top = parts[i - 1].bottom;
    parts[i].bottom = parts[i - 1].bottom + (int64_t)partitionHeight;
  }
  parts.back().bottom = top + height;

  std::vector<Entry> entries(n);
  for (int i = 0; i < n; i++) entries[i] = initEntry(seg[4 * i], seg[4 * i + 1], seg[4 * i + 2], seg[4 * i + 3], wind[i]);

  if (numPartitions == 1) {
    parts[0].entries = entries;
  } else {
    auto prange = [&](const Entry& e, uint32_t& a, uint32_t& b) {
      float fa = e.ay - (float)startY, fb = e.by - (float)startY;
      fa = fa > 0 ? fa : 0;  // max(0, x)
      fb = fb > 0 ? fb : 0;
      a = (uint32_t)fa / partitionHeight;
      b = (uint32_t)fb / partitionHeight;
      if (a > numPartitions - 1) a = numPartitions - 1;
      if (b > numPartitions - 1) b = numPartitions - 1;
    };
    for (int i = 0; i < n; i++) {
      uint32_t a, b;
      prange(entries[i], a, b);
      for (uint32_t p = a; p <= b; p++) parts[p].entries.push_back(entries[i]);
    }
  }

  for (auto& part : parts) {
    part.requiresAA = false;
    for (const auto& e : part.entries)
      if (segRequiresAA(e)) {
        part.requiresAA = true;
        break;
      }
    float top_ = (float)part.top, bottom_ = (float)part.bottom;
    for (auto& e : part.entries) {
      if (e.ay <= top_ && e.by >= bottom_) {
        float atx = 0, aty = 0;
        segLine(e, top_, atx, aty);
        e.ax = atx;
        e.ay = aty;
        segLine(e, bottom_, atx, aty);
        e.bx = atx;
        e.by = aty;
      }
    }
    if (part.entries.size() == 2) {
      const Entry& e0 = part.entries[0];
      const Entry& e1 = part.entries[1];
      if (!intersectsInside(e0, e1)) {
        if (e0.ay <= top_ && e0.by >= bottom_ && e1.ay <= top_ && e1.by >= bottom_) {
          part.twoSpanning = true;
          float m0 = (e0.ax + e0.bx) * 0.5f, m1 = (e1.ax + e1.bx) * 0.5f;
          if (m0 > m1) std::swap(part.entries[0], part.entries[1]);
        }
      }
    }
  }
}

struct Hit {
  int32_t at;
  int16_t winding;
};

inline bool shouldFill(int rule, int64_t count) {  // :1288-1296
  return rule == 0 ? count != 0 : (count % 2) != 0;
}
inline int32_t fxInteger(int32_t p) { return p / 256; }        // :1271-1272 (trunc toward zero)
inline int32_t fxTrunc(int32_t p) { return (p / 256) * 256; }  // :1274-1275

void sortHits(std::vector<Hit>& hits, int len) {  // :1277-1286 insertion sort (stable)
  for (int i = 1; i < len; i++) {
    int j = i - 1, k = i;
    while (j >= 0 && hits[j].at > hits[k].at) {
      std::swap(hits[j + 1], hits[j]);
      j--;
      k--;
    }
  }
}

// walk (:1298-1330): calls f(prevAt, at) for every yielded span.
template <typename F>
void walk(const std::vector<Hit>& hits, int numHits, int rule, F f) {
  int i = 0;
  int64_t count = 0;
  int32_t prevAt = 0;
  while (i < numHits) {
    int32_t at = hits[i].at;
    int16_t winding = hits[i].winding;
    if (at > 0) {
      if (shouldFill(rule, count)) {
        if (i < numHits - 1) {
          int32_t nextAt = hits[i + 1].at;
          int16_t nextWinding = hits[i + 1].winding;
          if (nextAt == at && winding + nextWinding == 0) {
            i += 2;
            continue;
          }
          if (rule == 0 && count + winding != 0) {
            count += winding;
            i++;
            continue;
          }
        }
        f(prevAt, at);
      }
      prevAt = at;
    }
    count += winding;
    i++;
  }
}
template <typename F>
void walkInteger(const std::vector<Hit>& hits, int numHits, int rule, F f) {  // :1336-1348
  walk(hits, numHits, rule, [&](int32_t prevAt, int32_t at) {
    int64_t fillStart = fxInteger(prevAt);
    int64_t fillLen = fxInteger(at) - fillStart;
    if (fillLen <= 0) return;
    f(fillStart, fillLen);
  });
}

struct Canvas {
  px_t* data;
  int64_t w, h;
  int sem;
  uint64_t covered;
  inline px_t& at(int64_t x, int64_t y) { return data[w * y + x]; }
  void fillRange(int64_t start, int64_t len, px_t c) {
    if (start < 0) { len += start; start = 0; }
    if (start + len > w * h) len = w * h - start;
    for (int64_t i = 0; i < len; i++) data[start + i] = c;
  }
  void clearUnsafe(int64_t sx, int64_t sy, int64_t tx, int64_t ty) {  // :1433-1440
    if (sx == w || sy == h) return;
    int64_t start = w * sy + sx;
    int64_t len = (w * ty + tx) - start;
    fillRange(start, len, 0);
  }
};

// fillHits (:1540-1591)
void fillHits(Canvas& im, px_t rgbx, int64_t startX, int64_t y, const std::vector<Hit>& hits, int numHits,
              int rule, int mode, bool maskClears = true) {
  if (mode == OverwriteBlend) {
    walkInteger(hits, numHits, rule, [&](int64_t start, int64_t len) {
      im.fillRange(im.w * y + start, len, rgbx);
      im.covered += len;
    });
  } else if (mode == NormalBlend) {
    walkInteger(hits, numHits, rule, [&](int64_t start, int64_t len) {
      im.covered += len;
      if (A(rgbx) == 255) {
        im.fillRange(im.w * y + start, len, rgbx);
      } else {
        for (int64_t i = 0; i < len; i++) {
          px_t& p = im.at(start + i, y);
          p = im.sem == 0 ? lineNormal(p, rgbx) : blendNormal(p, rgbx);
        }
      }
    });
  } else if (mode == MaskBlend) {
    int64_t filledTo = startX;
    walkInteger(hits, numHits, rule, [&](int64_t start, int64_t len) {
      im.covered += len;
      if (maskClears) {
        int64_t gap = start - filledTo;
        if (gap > 0) im.fillRange(im.w * y + filledTo, gap, 0);
      }
      if (A(rgbx) != 255)
        for (int64_t i = 0; i < len; i++) {
          px_t& p = im.at(start + i, y);
          p = lineMask(p, rgbx);
        }
      filledTo = start + len;
    });
    if (maskClears) {
      im.clearUnsafe(0, y, startX, y);
      im.clearUnsafe(filledTo, y, im.w, y);
    }
  } else {
    walkInteger(hits, numHits, rule, [&](int64_t start, int64_t len) {
      im.covered += len;
      for (int64_t i = 0; i < len; i++) {
        px_t& p = im.at(start + i, y);
        p = blendPx(mode, p, rgbx);
      }
    });
  }
}

// fillCoverage (:1479-1526) + blendLineCoverage* (scalar :1442-1477, x86 sse2.nim:526-566,618-667,717-769)
void fillCoverage(Canvas& im, px_t rgbx, int64_t startX, int64_t y, const std::vector<uint8_t>& cov, int mode) {
  const int64_t len = (int64_t)cov.size();
  for (int64_t i = 0; i < len; i++)
    if (cov[i] != 0) im.covered++;
  if (mode == OverwriteBlend) {
    for (int64_t i = 0; i < len; i++) {
      uint32_t c = cov[i];
      if (c != 0) im.at(startX + i, y) = im.sem == 0 ? mulCovFloor(rgbx, c) : mulCovScalar(rgbx, c);
    }
  } else if (mode == NormalBlend) {
    for (int64_t i = 0; i < len; i++) {
      uint32_t c = cov[i];
      if (c == 0) continue;
      px_t& p = im.at(startX + i, y);
      p = im.sem == 0 ? lineNormal(p, mulCovFloor(rgbx, c)) : blendNormal(p, mulCovScalar(rgbx, c));
    }
  } else if (mode == MaskBlend) {
    for (int64_t i = 0; i < len; i++) {
      uint32_t c = cov[i];
      px_t& p = im.at(startX + i, y);
      if (im.sem == 0) {
        p = lineMask(p, mulCovFloor(rgbx, c));
      } else {
        if (c == 255) continue;
        p = blendMask(p, mulCovScalar(rgbx, c));
      }
    }
    im.clearUnsafe(0, y, startX, y);
    im.clearUnsafe(startX + len, y, im.w, y);
  } else {
    for (int64_t i = 0; i < len; i++) {
      uint32_t c = cov[i];
      if (c != 0) {
        px_t& p = im.at(startX + i, y);
        p = blendPx(mode, p, mulCovScalar(rgbx, c));
      }
    }
  }
}

const float kEpsilon = (float)(0.0001 * 3.141592653589793238462643383279502884);  // paths.nim:44

// computeCoverage (:1350-1431)
void computeCoverage(std::vector<uint8_t>& cov, std::vector<Hit>& hits, int& numHits, int64_t width, int64_t y,
                     int64_t startX, const Partition& part, const std::vector<int>& entryIndices,
                     int numEntryIndices, int rule) {
  const bool aa = part.requiresAA;
  const int quality = aa ? 5 : 1;
  const uint32_t sampleCoverage = 255 / quality;
  const float offset = 1 / (float)quality;
  const float initialOffset = offset / 2 + kEpsilon;
  const int64_t covLen = (int64_t)cov.size();
  float yLine = (float)y + initialOffset - offset;
  for (int m = 0; m < quality; m++) {
    yLine += offset;
    numHits = 0;
    for (int i = 0; i < numEntryIndices; i++) {
      const Entry& e = part.entries[entryIndices[i]];
      if (e.ay <= yLine && e.by >= yLine) {
        float x = e.m == 0 ? e.b : (yLine - e.b) / e.m;
        float wf = (float)width;
        float mn = x < wf ? x : wf;  // min(x, width.float32): NaN x -> width
        if (x != x) mn = wf;
        hits[numHits].at = fixed32(mn);
        hits[numHits].winding = e.winding;
        numHits++;
      }
    }
    if (numHits > 0) sortHits(hits, numHits);
    if (aa) {
      walk(hits, numHits, rule, [&](int32_t prevAt, int32_t at) {
        int64_t fillStart = fxInteger(prevAt);
        bool pixelCrossed = fxInteger(at) != fxInteger(prevAt);
        int32_t leftCover = pixelCrossed ? fxTrunc(prevAt) + 256 - prevAt : at - prevAt;
        if (leftCover != 0) {
          fillStart++;
          int64_t idx = fxInteger(prevAt) - startX;
          if (idx >= 0 && idx < covLen) cov[idx] = (uint8_t)(cov[idx] + (uint8_t)fxInteger(leftCover * (int32_t)sampleCoverage));
        }
        if (pixelCrossed) {
          int32_t rightCover = at - fxTrunc(at);
          if (rightCover > 0) {
            int64_t idx = fxInteger(at) - startX;
            if (idx >= 0 && idx < covLen) cov[idx] = (uint8_t)(cov[idx] + (uint8_t)fxInteger(rightCover * (int32_t)sampleCoverage));
          }
        }
        int64_t fillLen = fxInteger(at) - fillStart;
        for (int64_t j = fillStart; j < fillStart + fillLen; j++) {
          int64_t idx = j - startX;
          if (idx >= 0 && idx < covLen) cov[idx] = (uint8_t)(cov[idx] + sampleCoverage);
        }
      });
    }
  }
}

int fillSegments(Canvas& im, const float* seg, const int16_t* wind, int n, px_t rgbx, int rule, int mode) {
  // An empty segment list is a no-op.  (With no segments computeBounds yields +/-Inf and the
  // float->int conversions that follow are undefined behaviour in the reference; its own tiger
  // SVG contains the empty path "M-65.4,9z" and renders in its CI, so "nothing drawn" is the
  // behaviour the reference's tests pin.)
  if (n == 0) return 0;
  // computeBounds (:1098-1117) + snapToPixels (common.nim:92-101) + clip (:1605-1619)
  float xMin = INFINITY, xMax = -INFINITY, yMin = INFINITY, yMax = -INFINITY;
  for (int i = 0; i < n; i++) {
    float ax = seg[4 * i], ay = seg[4 * i + 1], bx = seg[4 * i + 2], by = seg[4 * i + 3];
    xMin = fminf(xMin, fminf(ax, bx));
    xMax = fmaxf(xMax, fmaxf(ax, bx));
    yMin = fminf(yMin, ay);
    yMax = fmaxf(yMax, by);
  }
  float bx_ = 0, by_ = 0, bw_ = 0, bh_ = 0;
  if (!(xMin != xMin || xMax != xMax || yMin != yMin || yMax != yMax)) {
    bx_ = xMin;
    by_ = yMin;
    bw_ = xMax - xMin;
    bh_ = yMax - yMin;
  }
  float sx = floorf(bx_), sw = ceilf(bx_ + bw_) - sx;
  float sy = floorf(by_), sh = ceilf(by_ + bh_) - sy;
  int64_t startX = f2i(sx) > 0 ? f2i(sx) : 0;
  int64_t startY = f2i(sy) > 0 ? f2i(sy) : 0;
  int64_t pathWidth = 0;
  if (startX < im.w) {
    int64_t wI = f2i(sw);
    pathWidth = wI < im.w - startX ? wI : im.w - startX;
  }
  int64_t pathHeight = f2i(sy + sh) < im.h ? f2i(sy + sh) : im.h;
  if (pathWidth == 0) return 0;
  if (pathWidth < 0) {
    g_err = "Path int overflow detected";
    return 1;
  }

  std::vector<Partition> parts;
  if (pathHeight > startY) {
    partitionSegments(seg, wind, n, startY, pathHeight - startY, parts);
    size_t maxEntries = 0;
    for (auto& p : parts) maxEntries = p.entries.size() > maxEntries ? p.entries.size() : maxEntries;
    std::vector<int> entryIndices(maxEntries);
    int numEntryIndices = 0;
    struct Trap { float ax, ay, bx, by; };
    std::vector<Trap> trap(maxEntries);
    std::vector<uint8_t> coverages((size_t)pathWidth, 0);
    std::vector<Hit> hits(maxEntries > 2 ? maxEntries : 2);
    int numHits = 0;
    size_t partitionIndex = 0;

    int64_t y = startY;
    while (y < pathHeight) {
      if (y >= parts[partitionIndex].bottom) partitionIndex++;
      Partition& part = parts[partitionIndex];
      const int64_t partitionHeight = part.bottom - part.top;
      if (partitionHeight == 0) break;  // (:1641-1642 would spin forever; unreachable for height > 0)

      if (part.twoSpanning && !part.requiresAA) {  // mode A (:1644-1668)
        int64_t left = f2i(part.entries[0].ax), right = f2i(part.entries[1].ax);
        int64_t minX = left < 0 ? 0 : (left > im.w ? im.w : left);
        int64_t maxX = right < 0 ? 0 : (right > im.w ? im.w : right);
        for (int64_t r = 0; r < partitionHeight; r++) {
          hits[0].at = (int32_t)(minX * 256);
          hits[0].winding = 1;
          hits[1].at = (int32_t)(maxX * 256);
          hits[1].winding = -1;
          fillHits(im, rgbx, 0, y + r, hits, 2, 0, mode);
        }
        y += partitionHeight;
        continue;
      }

      const float scanTop = (float)y, scanBottom = (float)(y + 1);
      bool allSpan = true;
      numEntryIndices = 0;
      if (part.twoSpanning) {
        numEntryIndices = 2;
        entryIndices[0] = 0;
        entryIndices[1] = 1;
      } else {
        for (size_t i = 0; i < part.entries.size(); i++) {
          const Entry& e = part.entries[i];
          if (e.by <= scanTop || e.ay >= scanBottom) continue;
          if (e.ay > scanTop || e.by < scanBottom) allSpan = false;
          entryIndices[numEntryIndices++] = (int)i;
        }
      }

      bool done = false;
      if (allSpan && numEntryIndices % 2 == 0) {  // mode B (:1691-1872)
        for (int i = 0; i < numEntryIndices; i++) {
          int idx = entryIndices[i];
          trap[idx].ay = scanTop;
          trap[idx].by = scanBottom;
          trap[idx].ax = solveX(part.entries[idx], scanTop);
          trap[idx].bx = solveX(part.entries[idx], scanBottom);
        }
        auto midX = [&](int idx) { return (trap[idx].ax + trap[idx].bx) * 0.5f; };
        for (int i = 1; i < numEntryIndices; i++) {
          int j = i - 1, k = i;
          while (j >= 0 && midX(entryIndices[j]) > midX(entryIndices[k])) {
            std::swap(entryIndices[j + 1], entryIndices[j]);
            j--;
            k--;
          }
        }
        bool noOverlap = true;
        for (int i = 0; i < numEntryIndices - 1; i++) {
          const Trap& l = trap[entryIndices[i]];
          const Trap& r = trap[entryIndices[i + 1]];
          float leftMaxX = fmaxf(l.ax, l.bx), rightMinX = fminf(r.ax, r.bx);
          if (f2i(ceilf(leftMaxX)) > f2i(rightMinX)) {
            noOverlap = false;
            break;
          }
        }
        if (noOverlap) {
          bool simple = true;
          int64_t windingCount = 0;
          for (int i = 0; i < numEntryIndices; i += 2) {
            windingCount += part.entries[entryIndices[i]].winding;
            if (!shouldFill(rule, windingCount)) { simple = false; break; }
            windingCount += part.entries[entryIndices[i + 1]].winding;
            if (shouldFill(rule, windingCount)) { simple = false; break; }
          }
          if (simple) {
            int64_t filledTo = 0;
            for (int i = 0; i < numEntryIndices; i += 2) {
              const Entry& left = part.entries[entryIndices[i]];
              const Entry& right = part.entries[entryIndices[i + 1]];
              const Trap tl = trap[entryIndices[i]], tr = trap[entryIndices[i + 1]];
              const float leftMaxX = fmaxf(tl.ax, tl.bx), rightMinX = fminf(tr.ax, tr.bx);
              const int64_t leftCoverEnd = f2i(ceilf(leftMaxX)), rightCoverBegin = f2i(truncf(rightMinX));
              auto edgePixel = [&](int64_t x, float area) {
                px_t src = im.sem == 0 ? mulAreaSse(rgbx, area) : mulAreaScalar(rgbx, area);
                px_t& p = im.at(x, y);
                p = blendPx(mode, p, src);
                im.covered++;
              };
              {  // left-side partial coverage (:1772-1809)
                const bool inverted = tl.ax < tl.bx;
                const float sliverStart = fminf(tl.ax, tl.bx), rectStart = leftMaxX;
                float pen = sliverStart, prevPen = pen;
                float penY = inverted ? (float)y : (float)(y + 1), prevPenY = penY;
                for (int64_t x = f2i(sliverStart); x < f2i(ceilf(rectStart)); x++) {
                  prevPen = pen;
                  pen = (float)(x + 1);
                  float rightRectArea = 0;
                  if (pen > rectStart) {
                    rightRectArea = pen - rectStart;
                    pen = rectStart;
                  }
                  prevPenY = penY;
                  penY = solveY(left, pen);
                  if (x < 0 || x >= im.w) continue;
                  float run = pen - prevPen;
                  float triangleArea = 0.5f * run * fabsf(penY - prevPenY);
                  float rectArea = inverted ? (prevPenY - (float)y) * run : ((float)(y + 1) - prevPenY) * run;
                  float area = triangleArea + rectArea + rightRectArea;
                  edgePixel(x, area);
                }
              }
              {  // right-side partial coverage (:1811-1847)
                const bool inverted = tr.ax > tr.bx;
                const float rectEnd = rightMinX, sliverEnd = fmaxf(tr.ax, tr.bx);
                float pen = rectEnd, prevPen = pen;
                float penY = inverted ? (float)(y + 1) : (float)y, prevPenY = penY;
                for (int64_t x = f2i(rectEnd); x < f2i(ceilf(sliverEnd)); x++) {
                  prevPen = pen;
                  pen = (float)(x + 1);
                  float leftRectArea = fabsf(prevPen) - floorf(fabsf(prevPen));  // vmath fractional(): abs(v) - floor(abs(v))
                  if (pen > sliverEnd) pen = sliverEnd;
                  prevPenY = penY;
                  penY = solveY(right, pen);
                  if (x < 0 || x >= im.w) continue;
                  float run = pen - prevPen;
                  float triangleArea = 0.5f * run * fabsf(penY - prevPenY);
                  float rectArea = inverted ? (penY - (float)y) * run : ((float)(y + 1) - penY) * run;
                  float area = leftRectArea + triangleArea + rectArea;
                  edgePixel(x, area);
                }
              }
              int64_t fillBegin = leftCoverEnd < 0 ? 0 : (leftCoverEnd > im.w ? im.w : leftCoverEnd);
              int64_t fillEnd = rightCoverBegin < 0 ? 0 : (rightCoverBegin > im.w ? im.w : rightCoverBegin);
              hits[0].at = fixed32((float)fillBegin);
              hits[0].winding = 1;
              hits[1].at = fixed32((float)fillEnd);
              hits[1].winding = -1;
              fillHits(im, rgbx, 0, y, hits, 2, 0, mode, false);
              if (mode == MaskBlend) {
                int64_t clearTo = f2i(fminf(tl.ax, tl.bx));
                int64_t a = filledTo < im.w ? filledTo : im.w, b = clearTo < im.w ? clearTo : im.w;
                im.clearUnsafe(a, y, b, y);
              }
              filledTo = f2i(ceilf(fmaxf(tr.ax, tr.bx)));
            }
            if (mode == MaskBlend) im.clearUnsafe(filledTo < im.w ? filledTo : im.w, y, im.w, y);
            y++;
            done = true;
          }
        }
      }
      if (done) continue;

      // mode C (:1874-1906)
      computeCoverage(coverages, hits, numHits, im.w, y, startX, part, entryIndices, numEntryIndices, rule);
      if (part.requiresAA) {
        fillCoverage(im, rgbx, startX, y, coverages, mode);
        std::fill(coverages.begin(), coverages.end(), 0);
      } else {
        fillHits(im, rgbx, startX, y, hits, numHits, rule, mode);
      }
      y++;
    }
  }

  if (mode == MaskBlend) {  // :1910-1912 (clamped to the image; the reference would write out of bounds)
    int64_t sY = startY < im.h ? startY : im.h;
    int64_t pH = pathHeight < sY ? sY : pathHeight;
    im.fillRange(0, im.w * sY, 0);
    im.fillRange(im.w * pH, im.w * (im.h - pH), 0);
  }
  return 0;
}

}  // namespace

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

int orc_fill_segments(uint8_t* img, int w, int h, const float* seg, const int16_t* wind, int n, uint32_t rgbx,
                      int rule, int mode, int sem, uint64_t* covered_px) {
  if (mode < 0 || mode >= NumBlendModes || rule < 0 || rule > 1) {
    g_err = "invalid enum";
    return 2;
  }
  Canvas im;
  im.data = (px_t*)img;
  im.w = w;
  im.h = h;
  im.sem = sem;
  im.covered = 0;
  int rc = fillSegments(im, seg, wind, n, rgbx, rule, mode);
  if (covered_px) *covered_px += im.covered;
  return rc;
}

uint32_t orc_blend_px(int mode, uint32_t backdrop, uint32_t source) { return blendPx(mode, backdrop, source); }

// blendRect (images.nim:468-529); Normal/Mask rows use the x86 row-kernel bodies (canonical).
int orc_blend_rect(uint8_t* dst_, int dw, int dh, const uint8_t* src_, int sw, int sh, int px, int py, int mode) {
  px_t* a = (px_t*)dst_;
  const px_t* b = (const px_t*)src_;
  if (px >= dw || px + sw <= 0 || py >= dh || py + sh <= 0) {
    if (mode == MaskBlend) memset(a, 0, (size_t)dw * dh * 4);
    return 0;
  }
  int xStart = -px > 0 ? -px : 0, yStart = -py > 0 ? -py : 0;
  int xEnd = sw < dw - px ? sw : dw - px, yEnd = sh < dh - py ? sh : dh - py;
  if (mode == MaskBlend) {
    if (yStart + py > 0) memset(a, 0, (size_t)(yStart + py) * dw * 4);
    for (int y = yStart; y < yEnd; y++) {
      px_t* row = a + (size_t)dw * (y + py);
      for (int x = 0; x < xStart + px; x++) row[x] = 0;
      for (int x = xStart; x < xEnd; x++) row[x + px] = lineMask(row[x + px], b[(size_t)sw * y + x]);
      for (int x = xEnd + px; x < dw; x++) row[x] = 0;
    }
    if (yEnd + py < dh) memset(a + (size_t)dw * (yEnd + py), 0, (size_t)(dh - (yEnd + py)) * dw * 4);
    return 0;
  }
  for (int y = yStart; y < yEnd; y++) {
    px_t* row = a + (size_t)dw * (y + py) + px;
    const px_t* srow = b + (size_t)sw * y;
    for (int x = xStart; x < xEnd; x++) {
      if (mode == NormalBlend) row[x] = lineNormal(row[x], srow[x]);
      else if (mode == OverwriteBlend) row[x] = srow[x];
      else row[x] = blendPx(mode, row[x], srow[x]);
    }
  }
  return 0;
}

int orc_blend_rect_masked(uint8_t* dst, int dw, int dh, const uint8_t* src, const uint8_t* mask, int mask_is_rgbx,
                          int sw, int sh, int px, int py, int mode) {
  std::vector<px_t> tmp((size_t)sw * sh);
  const px_t* s = (const px_t*)src;
  for (size_t i = 0; i < tmp.size(); i++) {
    uint32_t k = mask_is_rgbx ? mask[4 * i + 3] : mask[i];
    tmp[i] = lineMask(s[i], (px_t)k << 24);
  }
  return orc_blend_rect(dst, dw, dh, (const uint8_t*)tmp.data(), sw, sh, px, py, mode);
}

int orc_apply_opacity(uint8_t* img, int w, int h, float opacity) {  // images.nim:261-277
  uint32_t o = (uint32_t)(uint16_t)(int64_t)roundf(255 * opacity);
  if (o == 255) return 0;
  size_t n = (size_t)w * h * 4;
  if (o == 0) {
    memset(img, 0, n);
    return 0;
  }
  for (size_t i = 0; i < n; i++) img[i] = (uint8_t)((img[i] * o) / 255);
  return 0;
}

int orc_blur(uint8_t* img, int w, int h, const uint16_t* lut, int radius, uint32_t oob) {  // images.nim:304-365
  if (radius == 0) return 0;
  if (radius < 0) {
    g_err = "Cannot apply negative blur";
    return 1;
  }
  std::vector<uint8_t> blurX((size_t)w * h * 4);  // kept un-transposed; same values
  const uint8_t oobc[4] = {(uint8_t)R(oob), (uint8_t)G(oob), (uint8_t)B(oob), (uint8_t)A(oob)};
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      uint32_t v[4] = {0, 0, 0, 0};
      for (int xx = x - radius; xx <= x + radius; xx++) {
        uint32_t k = lut[xx - x + radius];
        const uint8_t* s = (xx < 0 || xx >= w) ? oobc : img + ((size_t)w * y + xx) * 4;
        for (int c = 0; c < 4; c++) v[c] += s[c] * k;
      }
      for (int c = 0; c < 4; c++) blurX[((size_t)w * y + x) * 4 + c] = (uint8_t)(v[c] / 256 / 255);
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      uint32_t v[4] = {0, 0, 0, 0};
      for (int yy = y - radius; yy <= y + radius; yy++) {
        uint32_t k = lut[yy - y + radius];
        const uint8_t* s = (yy < 0 || yy >= h) ? oobc : blurX.data() + ((size_t)w * yy + x) * 4;
        for (int c = 0; c < 4; c++) v[c] += s[c] * k;
      }
      for (int c = 0; c < 4; c++) img[((size_t)w * y + x) * 4 + c] = (uint8_t)(v[c] / 256 / 255);
    }
  return 0;
}

int orc_spread(uint8_t* img, int w, int h, int spread) {  // images.nim:700-758
  if (spread == 0) return 0;
  const bool grow = spread > 0;
  const int s = grow ? spread : -spread;
  std::vector<uint8_t> tmp((size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      uint8_t v = grow ? 0 : 255;
      int lo = x - s > 0 ? x - s : 0, hi = x + s < w - 1 ? x + s : w - 1;
      for (int xx = lo; xx <= hi; xx++) {
        uint8_t a = img[((size_t)w * y + xx) * 4 + 3];
        if (grow ? a > v : a < v) v = a;
      }
      tmp[(size_t)w * y + x] = v;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      uint8_t v = grow ? 0 : 255;
      int lo = y - s > 0 ? y - s : 0, hi = y + s < h - 1 ? y + s : h - 1;
      for (int yy = lo; yy <= hi; yy++) {
        uint8_t a = tmp[(size_t)w * yy + x];
        if (grow ? a > v : a < v) v = a;
      }
      uint8_t* p = img + ((size_t)w * y + x) * 4;
      p[0] = p[1] = p[2] = 0;
      p[3] = v;
    }
  return 0;
}

int orc_draw(uint8_t* dst, int dw, int dh, const uint8_t* src, int sw, int sh, const float* mat, int mode);

int orc_shadow(const uint8_t* img, int w, int h, float ox, float oy, int spread, const uint16_t* lut, int radius,
               uint32_t rgbx, uint8_t* out) {  // images.nim:760-776
  std::vector<uint8_t> mask((size_t)w * h * 4, 0);
  if (ox == 0 && oy == 0) {
    memcpy(mask.data(), img, mask.size());
  } else {  // mask.draw(image, translate(offset), OverwriteBlend)
    const float t[9] = {1, 0, 0, 0, 1, 0, ox, oy, 1};
    orc_draw(mask.data(), w, h, img, w, h, t, OverwriteBlend);
  }
  orc_spread(mask.data(), w, h, spread);
  int rc = orc_blur(mask.data(), w, h, lut, radius, 0);
  if (rc) return rc;
  px_t* o = (px_t*)out;
  for (size_t i = 0; i < (size_t)w * h; i++) o[i] = rgbx;
  return orc_blend_rect(out, w, h, mask.data(), w, h, 0, 0, MaskBlend);
}


// =====================================================================================================
// draw with any transform (images.nim:636-678): minifyBy2 / magnifyBy2 (:168-259), getRgbaSmooth
// (:367-403), drawSmooth (:531-634), drawCorrect / drawTiled (:405-449, :680-683); gradient paints
// (paints.nim:68-248).  vmath / bumpy pieces restated from their published definitions.
// =====================================================================================================
}  // extern "C" (internal helpers follow)

namespace {

struct V2 { float x, y; };
inline V2 v2(float x, float y) { V2 r = {x, y}; return r; }
inline V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
inline V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
inline V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
inline V2 operator/(V2 a, float s) { return v2(a.x / s, a.y / s); }
inline float vlen(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
struct M3 { float m[9]; };  // vmath Mat3, column-major: m[c*3+r]
inline V2 mulV(const M3& a, V2 b) { return v2(a.m[0] * b.x + a.m[3] * b.y + a.m[6], a.m[1] * b.x + a.m[4] * b.y + a.m[7]); }
inline M3 mulM(const M3& a, const M3& b) {  // vmath `*`(a, b: Mat3)
  M3 r;
  for (int c = 0; c < 3; c++)
    for (int row = 0; row < 3; row++)
      r.m[c * 3 + row] = b.m[c * 3 + 0] * a.m[0 * 3 + row] + b.m[c * 3 + 1] * a.m[1 * 3 + row] + b.m[c * 3 + 2] * a.m[2 * 3 + row];
  return r;
}
inline M3 scaleM(float x, float y) { M3 r = {{x, 0, 0, 0, y, 0, 0, 0, 1}}; return r; }
inline M3 translateM(float x, float y) { M3 r = {{1, 0, 0, 0, 1, 0, x, y, 1}}; return r; }
inline M3 rotateM(float angle) {
  const float s = sinf(angle), c = cosf(angle);
  M3 r = {{c, s, 0, -s, c, 0, 0, 0, 1}};
  return r;
}
M3 inverseM(const M3& a) {  // vmath inverse(Mat3): adjugate * (1 / determinant); A(c, r) = a.m[c*3+r]
#define A_(c, r) a.m[(c) * 3 + (r)]
  const float det = A_(0, 0) * (A_(1, 1) * A_(2, 2) - A_(2, 1) * A_(1, 2)) - A_(0, 1) * (A_(1, 0) * A_(2, 2) - A_(1, 2) * A_(2, 0)) +
                    A_(0, 2) * (A_(1, 0) * A_(2, 1) - A_(1, 1) * A_(2, 0));
  const float inv = 1.0f / det;
  M3 r;
#define R_(c, r_) r.m[(c) * 3 + (r_)]
  R_(0, 0) = +(A_(1, 1) * A_(2, 2) - A_(2, 1) * A_(1, 2)) * inv;
  R_(0, 1) = -(A_(0, 1) * A_(2, 2) - A_(0, 2) * A_(2, 1)) * inv;
  R_(0, 2) = +(A_(0, 1) * A_(1, 2) - A_(0, 2) * A_(1, 1)) * inv;
  R_(1, 0) = -(A_(1, 0) * A_(2, 2) - A_(1, 2) * A_(2, 0)) * inv;
  R_(1, 1) = +(A_(0, 0) * A_(2, 2) - A_(0, 2) * A_(2, 0)) * inv;
  R_(1, 2) = -(A_(0, 0) * A_(1, 2) - A_(1, 0) * A_(0, 2)) * inv;
  R_(2, 0) = +(A_(1, 0) * A_(2, 1) - A_(2, 0) * A_(1, 1)) * inv;
  R_(2, 1) = -(A_(0, 0) * A_(2, 1) - A_(2, 0) * A_(0, 1)) * inv;
  R_(2, 2) = +(A_(0, 0) * A_(1, 1) - A_(1, 0) * A_(0, 1)) * inv;
#undef A_
#undef R_
  return r;
}
inline float fractionalV(float v) { v = fabsf(v); return v - truncf(v); }
inline float fixAngleV(float a) {
  const float pi = (float)3.141592653589793238462643383279502884, tau = (float)(2 * 3.141592653589793238462643383279502884);
  while (a > pi) a -= tau;
  while (a < -pi) a += tau;
  return a;
}

struct Img {
  int w = 0, h = 0;
  std::vector<px_t> d;
  Img() {}
  Img(int w_, int h_) : w(w_), h(h_), d((size_t)w_ * h_, 0) {}
};

inline px_t mixPx(px_t a, px_t b, float t) {  // common.nim:59-65
  const uint32_t x = (uint32_t)(int64_t)roundf(t * 255);
  return mk((R(a) * (255 - x) + R(b) * x + 127) / 255, (G(a) * (255 - x) + G(b) * x + 127) / 255,
            (B(a) * (255 - x) + B(b) * x + 127) / 255, (A(a) * (255 - x) + A(b) * x + 127) / 255);
}

Img minifyOnce(const Img& src) {  // images.nim:181-236 (the SSE2 body :363-468 computes the same bytes)
  const bool wOdd = src.w % 2 != 0, hOdd = src.h % 2 != 0;
  const int ew = src.w / 2, eh = src.h / 2;
  Img r(wOdd ? ew + 1 : ew, hOdd ? eh + 1 : eh);
  auto at = [&](int x, int y) { return src.d[(size_t)src.w * y + x]; };
  for (int y = 0; y < eh; y++) {
    for (int x = 0; x < ew; x++) {
      const px_t a = at(2 * x, 2 * y), b = at(2 * x + 1, 2 * y), c = at(2 * x + 1, 2 * y + 1), d = at(2 * x, 2 * y + 1);
      r.d[(size_t)r.w * y + x] = mk((R(a) + R(b) + R(c) + R(d) + 2) / 4, (G(a) + G(b) + G(c) + G(d) + 2) / 4,
                                    (B(a) + B(b) + B(c) + B(d) + 2) / 4, (A(a) + A(b) + A(c) + A(d) + 2) / 4);
    }
    if (wOdd) r.d[(size_t)r.w * y + r.w - 1] = mulAreaScalar(mixPx(at(src.w - 1, 2 * y), at(src.w - 1, 2 * y + 1), 0.5f), 0.5f);
  }
  if (hOdd) {
    for (int x = 0; x < ew; x++)
      r.d[(size_t)r.w * (r.h - 1) + x] = mulAreaScalar(mixPx(at(2 * x, src.h - 1), at(2 * x + 1, src.h - 1), 0.5f), 0.5f);
    if (wOdd) r.d[(size_t)r.w * (r.h - 1) + r.w - 1] = mulAreaScalar(at(src.w - 1, src.h - 1), 0.25f);
  }
  return r;
}

Img magnifyOnce(const Img& src) {  // images.nim:238-259 with power = 1
  Img r(src.w * 2, src.h * 2);
  for (int y = 0; y < r.h; y++)
    for (int x = 0; x < r.w; x++) r.d[(size_t)r.w * y + x] = src.d[(size_t)src.w * (y / 2) + x / 2];
  return r;
}

inline px_t getPx(const Img& im, int64_t x, int64_t y) {  // image[x, y]: transparent outside
  if (x < 0 || y < 0 || x >= im.w || y >= im.h) return 0;
  return im.d[(size_t)im.w * y + x];
}
inline px_t getPxWrapped(const Img& im, int64_t x, int64_t y) {
  // image.unsafe[x mod w, y mod h]: Nim's mod keeps the sign of the dividend, so negative coordinates
  // address the pixel `linear index` w * (y mod h) + (x mod w); outside the buffer reads as transparent
  const int64_t idx = (int64_t)im.w * (y % im.h) + (x % im.w);
  if (idx < 0 || idx >= (int64_t)im.d.size()) return 0;
  return im.d[(size_t)idx];
}

px_t getRgbaSmooth(const Img& im, float x, float y, bool wrapped) {  // images.nim:367-403
  const float fx = floorf(x), fy = floorf(y);
  const int64_t x0 = f2i(fx), y0 = f2i(fy), x1 = x0 + 1, y1 = y0 + 1;
  const float xFrac = x - fx, yFrac = y - fy;
  px_t x0y0, x1y0, x0y1, x1y1;
  if (wrapped) {
    x0y0 = getPxWrapped(im, x0, y0); x1y0 = getPxWrapped(im, x1, y0);
    x0y1 = getPxWrapped(im, x0, y1); x1y1 = getPxWrapped(im, x1, y1);
  } else {
    x0y0 = getPx(im, x0, y0); x1y0 = getPx(im, x1, y0);
    x0y1 = getPx(im, x0, y1); x1y1 = getPx(im, x1, y1);
  }
  px_t top = x0y0;
  if (xFrac > 0 && x0y0 != x1y0) top = mixPx(x0y0, x1y0, xFrac);
  px_t bottom = x0y1;
  if (xFrac > 0 && x0y1 != x1y1) bottom = mixPx(x0y1, x1y1, xFrac);
  if (yFrac != 0 && top != bottom) return mixPx(top, bottom, yFrac);
  return top;
}

// bumpy intersects(Line, Segment, at): line a-b against the segment at-to
inline bool lineSegment(V2 la, V2 lb, V2 sat, V2 sto, V2& at) {
  const V2 s1 = lb - la, s2 = sto - sat;
  const float den = (-s2.x * s1.y + s1.x * s2.y);
  const float num = s1.x * (la.y - sat.y) - s1.y * (la.x - sat.x);
  const float u = num / den;
  if (u >= 0 && u <= 1) {
    at = sat + s2 * u;
    return true;
  }
  return false;
}
inline int64_t clampI(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

void drawSmooth(Img& a, const Img& b, const M3& transform, int mode) {  // images.nim:531-634
  const float hh = 0.5f;
  const V2 corners[4] = {mulV(transform, v2(0, 0)), mulV(transform, v2((float)b.w, 0)),
                         mulV(transform, v2((float)b.w, (float)b.h)), mulV(transform, v2(0, (float)b.h))};
  const M3 inv = inverseM(transform);
  const V2 p = mulV(inv, v2(0 + hh, 0 + hh));
  const V2 dx = mulV(inv, v2(1 + hh, 0 + hh)) - p;
  const V2 dy = mulV(inv, v2(0 + hh, 1 + hh)) - p;
  int64_t yStart = a.h, yEnd = 0;
  for (int k = 0; k < 4; k++) {
    yStart = std::min<int64_t>(yStart, f2i(floorf(corners[k].y)));
    yEnd = std::max<int64_t>(yEnd, f2i(ceilf(corners[k].y)));
  }
  yStart = clampI(yStart, 0, a.h);
  yEnd = clampI(yEnd, 0, a.h);
  if (mode == MaskBlend && yStart > 0) std::fill(a.d.begin(), a.d.begin() + (size_t)yStart * a.w, 0u);
  std::vector<px_t> sampleLine(a.w);
  for (int64_t y = yStart; y < yEnd; y++) {
    float xMin = (float)a.w, xMax = 0.0f;
    for (int yo = 0; yo < 2; yo++) {
      const V2 la = v2(-1000, (float)y + (float)yo), lb = v2(1000, (float)y + (float)yo);
      for (int k = 0; k < 4; k++) {
        const V2 sat = corners[k], sto = corners[(k + 1) & 3];
        V2 at = v2(0, 0);
        if (lineSegment(la, lb, sat, sto, at) && (sto.x != at.x || sto.y != at.y)) {
          xMin = xMin <= at.x ? xMin : at.x;  // Nim min/max
          xMax = at.x <= xMax ? xMax : at.x;
        }
      }
    }
    const int64_t xStart = clampI(f2i(floorf(xMin)), 0, a.w), xEnd = clampI(f2i(ceilf(xMax)), 0, a.w);
    if (xEnd - xStart == 0) continue;
    V2 srcPos = (p + dx * (float)xStart) + dy * (float)y;
    srcPos = v2(srcPos.x - hh, srcPos.y - hh);
    for (int64_t x = xStart; x < xEnd; x++) {
      sampleLine[x] = getRgbaSmooth(b, srcPos.x, srcPos.y, false);
      srcPos = srcPos + dx;
    }
    px_t* row = a.d.data() + (size_t)a.w * y;
    if (mode == MaskBlend) {
      for (int64_t x = 0; x < xStart; x++) row[x] = 0;
      for (int64_t x = xStart; x < xEnd; x++) row[x] = lineMask(row[x], sampleLine[x]);
      for (int64_t x = xEnd; x < a.w; x++) row[x] = 0;
    } else {
      for (int64_t x = xStart; x < xEnd; x++) {
        if (mode == NormalBlend) row[x] = lineNormal(row[x], sampleLine[x]);
        else if (mode == OverwriteBlend) row[x] = sampleLine[x];
        else row[x] = blendPx(mode, row[x], sampleLine[x]);
      }
    }
  }
  if (mode == MaskBlend && a.h - yEnd > 0) std::fill(a.d.begin() + (size_t)yEnd * a.w, a.d.end(), 0u);
}

bool drawAny(Img& a, const Img& b0, M3 transform, int mode) {  // draw, images.nim:636-678
  const float hh = 0.5f;
  const M3 inv = inverseM(transform);
  V2 p = mulV(inv, v2(0 + hh, 0 + hh));
  V2 dx = mulV(inv, v2(1 + hh, 0 + hh)) - p;
  V2 dy = mulV(inv, v2(0 + hh, 1 + hh)) - p;
  float filterBy2 = std::max(vlen(dx), vlen(dy));
  Img tmp;
  const Img* b = &b0;
  while (filterBy2 >= 2.0f) {
    tmp = minifyOnce(*b);
    b = &tmp;
    p = p / 2; dx = dx / 2; dy = dy / 2;
    filterBy2 /= 2;
    transform = mulM(transform, scaleM(2, 2));
  }
  while (filterBy2 <= 0.5f) {
    // the reference keeps doubling until it runs out of memory; both sides stop at 2^28 pixels (1 GiB) with an error
    if ((long long)b->w * 2 * b->h * 2 > (1ll << 28)) return false;
    tmp = magnifyOnce(*b);
    b = &tmp;
    p = p * 2; dx = dx * 2; dy = dy * 2;
    filterBy2 *= 2;
    transform = mulM(transform, scaleM(1.0f / 2, 1.0f / 2));
  }
  const bool hasRotationOrScaling = !(dx.x == 1 && dx.y == 0 && dy.x == 0 && dy.y == 1);
  const bool smooth = !(vlen(dx) == 1.0f && vlen(dy) == 1.0f && fractionalV(transform.m[6]) == 0.0f &&
                        fractionalV(transform.m[7]) == 0.0f);
  if (hasRotationOrScaling || smooth) {
    drawSmooth(a, *b, transform, mode);
  } else {
    orc_blend_rect((uint8_t*)a.d.data(), a.w, a.h, (const uint8_t*)b->d.data(), b->w, b->h, (int)transform.m[6],
                   (int)transform.m[7], mode);
  }
  return true;
}

bool drawCorrect(Img& a, const Img& b0, const M3& transform, int mode, bool tiled) {  // images.nim:405-449
  const float hh = 0.5f;
  M3 inv = inverseM(transform);
  V2 p = mulV(inv, v2(0 + hh, 0 + hh));
  V2 dx = mulV(inv, v2(1 + hh, 0 + hh)) - p;
  V2 dy = mulV(inv, v2(0 + hh, 1 + hh)) - p;
  float filterBy2 = std::max(vlen(dx), vlen(dy));
  Img tmp;
  const Img* b = &b0;
  while (filterBy2 >= 2.0f) {
    tmp = minifyOnce(*b);
    b = &tmp;
    p = p / 2; dx = dx / 2; dy = dy / 2;
    filterBy2 /= 2;
    inv = mulM(scaleM(0.5f, 0.5f), inv);
  }
  while (filterBy2 <= 0.5f) {
    if ((long long)b->w * 2 * b->h * 2 > (1ll << 28)) return false;
    tmp = magnifyOnce(*b);
    b = &tmp;
    p = p * 2; dx = dx * 2; dy = dy * 2;
    filterBy2 *= 2;
    inv = mulM(scaleM(2, 2), inv);
  }
  for (int y = 0; y < a.h; y++)
    for (int x = 0; x < a.w; x++) {
      const V2 sp = mulV(inv, v2((float)x + hh, (float)y + hh));
      const px_t sample = getRgbaSmooth(*b, sp.x - hh, sp.y - hh, tiled);
      px_t& d = a.d[(size_t)a.w * y + x];
      d = blendPx(mode, d, sample);
    }
  return true;
}

// ---- gradient paints (paints.nim:68-248); chroma Color = 4 x float32, straight alpha
struct ColF { float r, g, b, a; };
inline uint32_t quantF(float v) {  // chroma Color -> ColorRGBA channel: round(v * 255), clamped
  float r = floorf(v * 255.0f + 0.5f);
  if (!(r > 0.0f)) return 0;
  return r > 255.0f ? 255u : (uint32_t)r;
}
inline px_t colorToRgbx(ColF c) {  // chroma Color.rgbx(): quantise, then premultiply (c*a+127) div 255
  const uint32_t a = quantF(c.a);
  const uint32_t r = quantF(c.r), g = quantF(c.g), b = quantF(c.b);
  if (a == 255) return mk(r, g, b, a);
  return mk((r * a + 127) / 255, (g * a + 127) / 255, (b * a + 127) / 255, a);
}
struct Gradient {
  int n;
  const float* pos;
  const ColF* col;
  float opacity;
};
px_t gradientColor(const Gradient& g, float t) {  // paints.nim:68-94
  int index = -1;
  for (int i = 0; i < g.n; i++) {
    if (g.pos[i] < t) index = i;
    if (g.pos[i] > t) break;
  }
  ColF c;
  if (index == -1) c = g.col[0];
  else if (index + 1 >= g.n) c = g.col[index];
  else {
    const ColF a = g.col[index], b = g.col[index + 1];
    const float v = (t - g.pos[index]) / (g.pos[index + 1] - g.pos[index]);
    c.r = a.r * (1.0f - v) + b.r * v;  // chroma mix(a, b: Color, v): lerp per channel
    c.g = a.g * (1.0f - v) + b.g * v;
    c.b = a.b * (1.0f - v) + b.b * v;
    c.a = a.a * (1.0f - v) + b.a * v;
  }
  c.a *= g.opacity;
  return colorToRgbx(c);
}

}  // namespace

extern "C" {

int orc_minify_by2(const uint8_t* src, int w, int h, int power, uint8_t* out, int* ow, int* oh) {
  if (power < 0) { g_err = "Cannot minifyBy2 with negative power"; return 1; }
  Img cur(w, h);
  memcpy(cur.d.data(), src, (size_t)w * h * 4);
  for (int i = 0; i < power; i++) cur = minifyOnce(cur);
  *ow = cur.w; *oh = cur.h;
  if (out) memcpy(out, cur.d.data(), cur.d.size() * 4);
  return 0;
}
int orc_magnify_by2(const uint8_t* src, int w, int h, int power, uint8_t* out) {
  if (power < 0) { g_err = "Cannot magnifyBy2 with negative power"; return 1; }
  const int scale = 1 << power;
  const px_t* s = (const px_t*)src;
  px_t* o = (px_t*)out;
  for (int y = 0; y < h * scale; y++)
    for (int x = 0; x < w * scale; x++) o[(size_t)w * scale * y + x] = s[(size_t)w * (y / scale) + x / scale];
  return 0;
}
/* draw(a, b, transform, blendMode): mat = vmath Mat3 storage (column-major, 9 floats). */
int orc_draw(uint8_t* dst, int dw, int dh, const uint8_t* src, int sw, int sh, const float* mat, int mode) {
  Img a(dw, dh), b(sw, sh);
  memcpy(a.d.data(), dst, a.d.size() * 4);
  memcpy(b.d.data(), src, b.d.size() * 4);
  M3 t;
  memcpy(t.m, mat, sizeof t.m);
  if (!drawAny(a, b, t, mode)) {
    g_err = "draw: magnified source image too large";
    return 1;
  }
  memcpy(dst, a.d.data(), a.d.size() * 4);
  return 0;
}
/* drawTiled(dst, src, mat, blendMode) = drawCorrect(..., tiled = true) (images.nim:680-683); tiled = 0
 * gives drawCorrect's untiled form. */
int orc_draw_correct(uint8_t* dst, int dw, int dh, const uint8_t* src, int sw, int sh, const float* mat, int mode,
                     int tiled) {
  Img a(dw, dh), b(sw, sh);
  memcpy(a.d.data(), dst, a.d.size() * 4);
  memcpy(b.d.data(), src, b.d.size() * 4);
  M3 t;
  memcpy(t.m, mat, sizeof t.m);
  if (!drawCorrect(a, b, t, mode, tiled != 0)) {
    g_err = "draw: magnified source image too large";
    return 1;
  }
  memcpy(dst, a.d.data(), a.d.size() * 4);
  return 0;
}
/* fillGradient (paints.nim:236-248).  kind: 3 linear, 4 radial, 5 angular (ord(PaintKind)); handles:
 * n_handles x {x, y}; stops: n_stops positions + n_stops x {r, g, b, a} float32 straight colours. */
int orc_fill_gradient(uint8_t* img, int w, int h, int kind, const float* handles, int n_handles, const float* stop_pos,
                      const float* stop_rgba, int n_stops, float opacity) {
  if (kind < 3 || kind > 5) { g_err = "Paint must be a gradient"; return 1; }
  const int need = kind == 3 ? 2 : 3;
  if (n_handles != need) {
    g_err = kind == 3 ? "Linear gradient requires 2 handles" : kind == 4 ? "Radial gradient requires 3 handles"
                                                                         : "Angular gradient requires 2 handles";
    return 1;
  }
  if (n_stops == 0) { g_err = "Gradient must have at least 1 color stop"; return 1; }
  opacity = opacity < 0 ? 0 : (opacity > 1 ? 1 : opacity);
  if (opacity == 0) return 0;
  Gradient g = {n_stops, stop_pos, (const ColF*)stop_rgba, opacity};
  px_t* d = (px_t*)img;
  const V2 h0 = v2(handles[0], handles[1]), h1 = v2(handles[2], handles[3]);
  if (kind == 3) {
    auto toLineSpace = [](V2 at, V2 to, V2 pt) {
      const V2 dd = to - at;
      const float det = dd.x * dd.x + dd.y * dd.y;
      return (dd.y * (pt.y - at.y) + dd.x * (pt.x - at.x)) / det;
    };
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        // the horizontal / vertical special cases (:115-165) evaluate the same expression at (x, 0) / (0, y)
        V2 xy = v2((float)x, (float)y);
        if (h0.y == h1.y) xy.y = 0;
        else if (h0.x == h1.x) xy.x = 0;
        d[(size_t)w * y + x] = gradientColor(g, toLineSpace(h0, h1, xy));
      }
    return 0;
  }
  const V2 h2 = v2(handles[4], handles[5]);
  if (kind == 4) {
    const V2 center = h0, edge = h1, skew = h2;
    const float distanceX = vlen(center - edge), distanceY = vlen(center - skew);
    const V2 n = (center - edge) / vlen(center - edge);
    const float gradientAngle = fixAngleV(atan2f(n.y, n.x));
    const M3 mat = inverseM(mulM(mulM(translateM(center.x, center.y), rotateM(gradientAngle)), scaleM(distanceX, distanceY)));
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) d[(size_t)w * y + x] = gradientColor(g, vlen(mulV(mat, v2((float)x, (float)y))));
    return 0;
  }
  const V2 center = h0, edge = h1;
  const float pi = (float)3.141592653589793238462643383279502884;
  const V2 n = (edge - center) / vlen(edge - center);
  const float gradientAngle = fixAngleV(atan2f(n.y, n.x));
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const V2 dlt = v2((float)x, (float)y) - center;
      const V2 nn = dlt / vlen(dlt);
      // arctan2 in float32: evaluated in double and rounded, which is what a correctly rounded atan2f
      // returns (the GPU kernel does the same, so the two agree bit for bit)
      const float angle = (float)atan2((double)nn.y, (double)nn.x);
      const float t = fixAngleV(angle + gradientAngle + pi / 2) / 2 / pi + 0.5f;
      d[(size_t)w * y + x] = gradientColor(g, t);
    }
  return 0;
}

}  // extern "C"

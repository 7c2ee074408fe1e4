// K7-K10 — draw with any transform and the non-solid paints
// (treeform/pixie src/pixie/images.nim: minifyBy2 :168-236, magnifyBy2 :238-259, getRgbaSmooth :367-403,
//  drawCorrect :405-449, drawSmooth :531-634, draw :636-678, drawTiled :680-683;
//  src/pixie/paints.nim: gradientColor :68-94, fillGradient{Linear,Radial,Angular} :96-248).
//
//   minify_kernel        2x2 box filter, (a + b + c + d + 2) div 4 per channel, odd edges via mix() * 0.5
//   magnify_kernel       pixel replication by 2^power
//   draw_smooth_kernel   one warp per destination row: the row's x range from the transformed perimeter, the
//                        source position accumulated pixel by pixel in float32 exactly as the reference does
//                        (srcPos += dx — the sum is not associative, so every lane replays the 32 additions of
//                        its chunk and keeps the value of its own step), bilinear sample, blendLine* / blender()
//   draw_correct_kernel  drawCorrect / drawTiled: one thread per destination pixel (positions are independent)
//   gradient_kernel      one thread per pixel: t from the handle geometry, colour from the stop list
//
// All HBM-bound gathers / streams; float32 geometry with one rounding per operation (-fmad=false).
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace pixie {

PXD long long f2ll_(float f) { return (long long)f; }
PXD int clampll(long long v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : (int)v); }

// ColorRGBX mix (common.nim:59-65): (a * (255 - x) + b * x + 127) div 255 per channel, two channels per
// register (16-bit lanes: the sums are <= 65152, below both the lane width and div255x2's 65534 limit)
PXD uint32_t round_half_away(float v) {  // Nim round() on a non-negative value below 2^23: exact
  uint32_t r = __float2uint_rz(v);
  if (v - (float)r >= 0.5f) r++;
  return r;
}
PXD px_t mix_px(px_t a, px_t b, float t) {
  const uint32_t x = round_half_away(t * 255.0f), ix = 255u - x;
  const uint32_t rb = (a & 0x00FF00FFu) * ix + (b & 0x00FF00FFu) * x + 0x007F007Fu;
  const uint32_t ga = ((a >> 8) & 0x00FF00FFu) * ix + ((b >> 8) & 0x00FF00FFu) * x + 0x007F007Fu;
  return div255x2(rb) | (div255x2(ga) << 8);
}
// ColorRGBX * float32 (common.nim:67-77)
PXD px_t mul_opacity(px_t c, float opacity) {
  if (opacity == 0.0f) return 0u;
  const uint32_t x = round_half_away(opacity * 255.0f);
  return mk((pR(c) * x + 127u) / 255u, (pG(c) * x + 127u) / 255u, (pB(c) * x + 127u) / 255u, (pA(c) * x + 127u) / 255u);
}

// ---------------------------------------------------------------------------------------------
// minifyBy2 / magnifyBy2
// ---------------------------------------------------------------------------------------------
PXD uint2 load2(const px_t* p) {  // two neighbouring pixels; rows of odd-width images start on odd pixels
  if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) return *reinterpret_cast<const uint2*>(p);
  return make_uint2(p[0], p[1]);
}
__global__ void __launch_bounds__(256) minify_kernel(const px_t* __restrict__ src, int sw, int sh, px_t* __restrict__ dst,
                                                     int dw, int dh) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= dw) return;
  const int ew = sw >> 1, eh = sh >> 1;
  for (int y = blockIdx.y; y < dh; y += gridDim.y) {
    px_t out;
    if (x < ew && y < eh) {
      const uint2 t = load2(src + (size_t)sw * (2 * y) + 2 * x);
      const uint2 b = load2(src + (size_t)sw * (2 * y + 1) + 2 * x);
      // four channels as two pairs of 16-bit lanes: sums <= 1022 never carry across lanes
      const uint32_t rb = (t.x & 0x00FF00FFu) + (t.y & 0x00FF00FFu) + (b.x & 0x00FF00FFu) + (b.y & 0x00FF00FFu) + 0x00020002u;
      const uint32_t ga = ((t.x >> 8) & 0x00FF00FFu) + ((t.y >> 8) & 0x00FF00FFu) + ((b.x >> 8) & 0x00FF00FFu) +
                          ((b.y >> 8) & 0x00FF00FFu) + 0x00020002u;
      out = ((rb >> 2) & 0x00FF00FFu) | (((ga >> 2) & 0x00FF00FFu) << 8);
    } else if (y < eh) {  // last column of an odd-width source (:216-222)
      out = mul_opacity(mix_px(src[(size_t)sw * (2 * y) + sw - 1], src[(size_t)sw * (2 * y + 1) + sw - 1], 0.5f), 0.5f);
    } else if (x < ew) {  // last row of an odd-height source (:224-231)
      out = mul_opacity(mix_px(src[(size_t)sw * (sh - 1) + 2 * x], src[(size_t)sw * (sh - 1) + 2 * x + 1], 0.5f), 0.5f);
    } else {  // the corner (:233-235)
      out = mul_opacity(src[(size_t)sw * (sh - 1) + sw - 1], 0.25f);
    }
    dst[(size_t)dw * y + x] = out;
  }
}

__global__ void __launch_bounds__(256) magnify_kernel(const px_t* __restrict__ src, int sw, px_t* __restrict__ dst, int dw,
                                                      int dh, int shift) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= dw) return;
  for (int y = blockIdx.y; y < dh; y += gridDim.y) dst[(size_t)dw * y + x] = src[(size_t)sw * (y >> shift) + (x >> shift)];
}

// ---------------------------------------------------------------------------------------------
// getRgbaSmooth (images.nim:367-403)
// ---------------------------------------------------------------------------------------------
struct SrcView {
  const px_t* d;
  int w, h;
};
PXD px_t get_px(const SrcView& s, long long x, long long y) {  // image[x, y]: transparent outside
  if (x < 0 || y < 0 || x >= s.w || y >= s.h) return 0u;
  return s.d[(size_t)s.w * (size_t)y + (size_t)x];
}
PXD px_t get_px_wrapped(const SrcView& s, long long x, long long y) {
  // image.unsafe[x mod w, y mod h] with Nim's sign-of-dividend mod: the linear index w*(y mod h) + (x mod w);
  // outside the buffer reads as transparent (the reference reads out of bounds there)
  const long long idx = (long long)s.w * (y % s.h) + (x % s.w);
  if (idx < 0 || idx >= (long long)s.w * s.h) return 0u;
  return s.d[idx];
}
template <bool WRAPPED>
PXD px_t rgba_smooth(const SrcView& s, float x, float y) {
  const float fx = floorf(x), fy = floorf(y);
  const float xFrac = x - fx, yFrac = y - fy;
  px_t x0y0, x1y0, x0y1, x1y1;
  if (WRAPPED) {
    const long long x0 = f2ll_(fx), y0 = f2ll_(fy), x1 = x0 + 1, y1 = y0 + 1;
    x0y0 = get_px_wrapped(s, x0, y0); x1y0 = get_px_wrapped(s, x1, y0);
    x0y1 = get_px_wrapped(s, x0, y1); x1y1 = get_px_wrapped(s, x1, y1);
  } else {
    // image[x, y] is transparent outside; positions beyond +-2^30 (or NaN) are outside whatever their low bits
    const bool sane = fabsf(fx) < 1073741824.0f && fabsf(fy) < 1073741824.0f;
    const int x0 = sane ? (int)fx : -2, y0 = sane ? (int)fy : -2, x1 = x0 + 1, y1 = y0 + 1;
    const bool cx0 = (unsigned)x0 < (unsigned)s.w, cx1 = (unsigned)x1 < (unsigned)s.w;
    const bool cy0 = (unsigned)y0 < (unsigned)s.h, cy1 = (unsigned)y1 < (unsigned)s.h;
    const px_t* r0 = s.d + (size_t)s.w * (size_t)(cy0 ? y0 : 0);
    const px_t* r1 = s.d + (size_t)s.w * (size_t)(cy1 ? y1 : 0);
    x0y0 = (cx0 && cy0) ? r0[x0] : 0u;
    x1y0 = (cx1 && cy0) ? r0[x1] : 0u;
    x0y1 = (cx0 && cy1) ? r1[x0] : 0u;
    x1y1 = (cx1 && cy1) ? r1[x1] : 0u;
  }
  px_t top = x0y0;
  if (xFrac > 0.0f && x0y0 != x1y0) top = mix_px(x0y0, x1y0, xFrac);
  px_t bottom = x0y1;
  if (xFrac > 0.0f && x0y1 != x1y1) bottom = mix_px(x0y1, x1y1, xFrac);
  if (yFrac != 0.0f && top != bottom) return mix_px(top, bottom, yFrac);
  return top;
}

// getRgbaSmooth in two halves, so that a thread can have the gathers of several pixels in flight before it mixes
struct Quad {
  px_t x0y0, x1y0, x0y1, x1y1;
  float xFrac, yFrac;
};
PXD Quad fetch_quad(const SrcView& s, float x, float y) {
  Quad q;
  const float fx = floorf(x), fy = floorf(y);
  q.xFrac = x - fx;
  q.yFrac = y - fy;
  const bool sane = fabsf(fx) < 1073741824.0f && fabsf(fy) < 1073741824.0f;
  const int x0 = sane ? (int)fx : -2, y0 = sane ? (int)fy : -2, x1 = x0 + 1, y1 = y0 + 1;
  const bool cx0 = (unsigned)x0 < (unsigned)s.w, cx1 = (unsigned)x1 < (unsigned)s.w;
  const bool cy0 = (unsigned)y0 < (unsigned)s.h, cy1 = (unsigned)y1 < (unsigned)s.h;
  const px_t* r0 = s.d + (size_t)s.w * (size_t)(cy0 ? y0 : 0);
  const px_t* r1 = s.d + (size_t)s.w * (size_t)(cy1 ? y1 : 0);
  q.x0y0 = (cx0 && cy0) ? __ldg(r0 + x0) : 0u;
  q.x1y0 = (cx1 && cy0) ? __ldg(r0 + x1) : 0u;
  q.x0y1 = (cx0 && cy1) ? __ldg(r1 + x0) : 0u;
  q.x1y1 = (cx1 && cy1) ? __ldg(r1 + x1) : 0u;
  return q;
}
PXD px_t resolve_quad(const Quad& q) {
  px_t top = q.x0y0;
  if (q.xFrac > 0.0f && q.x0y0 != q.x1y0) top = mix_px(q.x0y0, q.x1y0, q.xFrac);
  px_t bottom = q.x0y1;
  if (q.xFrac > 0.0f && q.x0y1 != q.x1y1) bottom = mix_px(q.x0y1, q.x1y1, q.xFrac);
  if (q.yFrac != 0.0f && top != bottom) return mix_px(top, bottom, q.yFrac);
  return top;
}

__device__ __noinline__ px_t blend_px_any(int mode, px_t b, px_t s) {
  px_t r = b;
  PX_DISPATCH_MODE(mode, r = blend_px<MODE>(b, s));
  return r;
}

// ---------------------------------------------------------------------------------------------
// drawSmooth (images.nim:531-634)
// ---------------------------------------------------------------------------------------------
struct SmoothArgs {
  px_t* a;
  int aw, ah;
  SrcView b;
  float cx[4], cy[4];  // transform * corners of b
  float px, py, dxx, dxy, dyx, dyy;
  int yStart, yEnd;
  int mode;
};

// bumpy intersects(Line, Segment, at) for the scanline (-1000, ly) - (1000, ly)
PXD bool line_segment(float ly, float sax, float say, float sbx, float sby, float& ox, float& oy) {
  const float s1x = 1000.0f - (-1000.0f), s1y = ly - ly;
  const float s2x = sbx - sax, s2y = sby - say;
  const float den = (-s2x * s1y + s1x * s2y);
  const float num = s1x * (ly - say) - s1y * (-1000.0f - sax);
  const float u = num / den;
  if (u >= 0.0f && u <= 1.0f) {
    ox = sax + s2x * u;
    oy = say + s2y * u;
    return true;
  }
  return false;
}

// The reference accumulates the source position along a row with `srcPos += dx` (:588): a float32 sum whose
// roundings depend on every previous step.  It still has a closed form, piecewise: while pos and pos + d stay in
// one binade [2^e, 2^(e+1)) every result is a multiple of the binade's ulp u, so fl(pos + d) = pos + D u with a
// constant integer D (ties-to-even settles after one step: the result of a tie is even, and from an even
// mantissa the same neighbour wins every time).  build_chain() walks one component of one row binade by binade
// — a handful of real additions at each crossing, one division for the length of the run — and leaves segments
// {first index, value, exact increment}; every pixel then gets its position as base + (k - k0) * inc, two exact
// float operations, in any order.  Rows whose chain does not fit the table (non-finite or wildly changing
// positions) take the sequential form below.
struct ChainSeg {
  int k0;
  float base, inc;
};
constexpr int kMaxSegs = 128;

__device__ __noinline__ int build_chain(float v0, float d, int n, ChainSeg* seg) {
  int ns = 0, k = 0;
  float cur = v0;
#pragma unroll 1
  while (k < n) {
    if (ns > kMaxSegs - 4) return -1;
    const float p1 = cur + d, p2 = p1 + d, p3 = p2 + d;
    const uint32_t b2 = __float_as_uint(p2);
    const uint32_t e1 = __float_as_uint(p1) >> 23, e2 = b2 >> 23, e3 = __float_as_uint(p3) >> 23;  // sign + exponent
    const bool run = e1 == e2 && e2 == e3 && (e2 & 0xFFu) != 0u && (e2 & 0xFFu) != 0xFFu;
    seg[ns].k0 = k; seg[ns].base = cur; seg[ns].inc = 0.0f; ns++;
    if (!run) {
      cur = p1;
      k++;
      continue;
    }
    seg[ns].k0 = k + 1; seg[ns].base = p1; seg[ns].inc = 0.0f; ns++;
    const float inc = p3 - p2;  // exact: same binade
    long long J = n;            // steps of the run after p2; inc == 0 never leaves the binade
    if (inc != 0.0f) {
      const float lo = __uint_as_float(b2 & 0x7F800000u);                 // 2^e
      const double u = (double)lo * (1.0 / 8388608.0);                     // ulp of the binade
      const double a2 = fabs((double)p2), ai = (p2 < 0.0f) ? -(double)inc : (double)inc;  // magnitude and its step
      if (ai > 0.0) {  // growing: results must stay below 2^(e+1)
        const double hi = 2.0 * (double)lo;
        J = (long long)floor((hi - a2) / ai);
        while (J > 0 && a2 + (double)J * ai >= hi) J--;
        while (a2 + (double)(J + 1) * ai < hi) J++;
      } else {  // shrinking: a result equal to 2^e may come from an exact sum below it, where the grid is finer
        const double lo1 = (double)lo + u;
        J = (long long)floor((a2 - lo1) / -ai);
        while (J > 0 && a2 + (double)J * ai < lo1) J--;
        while (a2 + (double)(J + 1) * ai >= lo1) J++;
      }
      if (J < 1) J = 1;  // p3 is in the binade
      if (J > n) J = n;
    }
    seg[ns].k0 = k + 2; seg[ns].base = p2; seg[ns].inc = inc; ns++;
    cur = (p2 + (float)J * inc) + d;  // the value after the run: a real addition again
    k = k + 2 + (int)J + 1;
  }
  seg[ns].k0 = INT_MAX; seg[ns].base = 0.0f; seg[ns].inc = 0.0f;
  return ns;
}

// MODE: NormalBlend / OverwriteBlend / MaskBlend get the blendLine* bodies, -1 = blender() chosen at run time
template <int MODE>
__global__ void __launch_bounds__(256) draw_smooth_kernel(const SmoothArgs A) {
  __shared__ ChainSeg s_segs[8][2][kMaxSegs];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rowsBegin = MODE == MaskBlend ? 0 : A.yStart, rowsEnd = MODE == MaskBlend ? A.ah : A.yEnd;
  const int y = rowsBegin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (y >= rowsEnd) return;
  px_t* row = A.a + (size_t)A.aw * y;
  if (y < A.yStart || y >= A.yEnd) {  // MaskBlend clears the rows the image does not reach (:556-557, :629-633)
    for (int x = lane; x < A.aw; x += 32) row[x] = 0u;
    return;
  }
  float xMin = (float)A.aw, xMax = 0.0f;
#pragma unroll
  for (int yo = 0; yo < 2; yo++) {
    const float ly = (float)y + (float)yo;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int k1 = (k + 1) & 3;
      float atx = 0.0f, aty = 0.0f;
      if (line_segment(ly, A.cx[k], A.cy[k], A.cx[k1], A.cy[k1], atx, aty) && (A.cx[k1] != atx || A.cy[k1] != aty)) {
        xMin = xMin <= atx ? xMin : atx;
        xMax = atx <= xMax ? xMax : atx;
      }
    }
  }
  const int xStart = clampll(f2ll_(floorf(xMin)), 0, A.aw), xEnd = clampll(f2ll_(ceilf(xMax)), 0, A.aw);
  if (xEnd - xStart == 0) return;
  if (MODE == MaskBlend) {
    for (int x = lane; x < xStart; x += 32) row[x] = 0u;
    for (int x = max(xEnd, 0) + lane; x < A.aw; x += 32) row[x] = 0u;
  }
  if (xEnd < xStart) return;
  // srcPos = p + dx * xStart + dy * y - h, then += dx per pixel (:584-588)
  float sx = (A.px + A.dxx * (float)xStart) + A.dyx * (float)y;
  float sy = (A.py + A.dxy * (float)xStart) + A.dyy * (float)y;
  sx = sx - 0.5f;
  sy = sy - 0.5f;
  const int n = xEnd - xStart;
  int ns = 0;
  if (lane < 2) ns = build_chain(lane == 0 ? sx : sy, lane == 0 ? A.dxx : A.dxy, n, s_segs[warp][lane]);
  __syncwarp();
  const bool closedForm = __shfl_sync(0xffffffffu, ns, 0) >= 0 && __shfl_sync(0xffffffffu, ns, 1) >= 0;
  const ChainSeg* segx = s_segs[warp][0];
  const ChainSeg* segy = s_segs[warp][1];
  int jx = 0, jy = 0;
  constexpr int U = 4;  // chunks of 32 pixels per iteration: 4 U gathers + U destination loads in flight per thread
#pragma unroll 1
  for (int base = xStart; base < xEnd; base += 32 * U) {
    float mx[U], my[U];
    if (closedForm) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int k = base - xStart + 32 * u + lane;
        while (segx[jx + 1].k0 <= k) jx++;
        while (segy[jy + 1].k0 <= k) jy++;
        mx[u] = segx[jx].base + (float)(k - segx[jx].k0) * segx[jx].inc;
        my[u] = segy[jy].base + (float)(k - segy[jy].k0) * segy[jy].inc;
      }
    } else {  // sequential form: every lane replays the additions of the chunk and keeps its own steps
#pragma unroll
      for (int u = 0; u < U; u++) {
        mx[u] = sx;
        my[u] = sy;
#pragma unroll
        for (int i = 0; i < 32; i++) {
          if (i == lane) {
            mx[u] = sx;
            my[u] = sy;
          }
          sx += A.dxx;
          sy += A.dxy;
        }
      }
    }
    Quad q[U];
    px_t d[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int x = base + 32 * u + lane;
      if (x < xEnd) {
        q[u] = fetch_quad(A.b, mx[u], my[u]);
        if (MODE != OverwriteBlend) d[u] = row[x];
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int x = base + 32 * u + lane;
      if (x < xEnd) {
        const px_t s = resolve_quad(q[u]);
        if (MODE == OverwriteBlend) row[x] = s;
        else if (MODE == NormalBlend) row[x] = line_normal(d[u], s);
        else if (MODE == MaskBlend) row[x] = line_mask(d[u], s);
        else row[x] = blend_px_any(A.mode, d[u], s);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// drawCorrect (images.nim:405-449), tiled = drawTiled (:680-683)
// ---------------------------------------------------------------------------------------------
struct CorrectArgs {
  px_t* a;
  int aw, ah;
  SrcView b;
  float m[9];  // inverse transform (after the minify / magnify adjustments)
  int mode;
};
template <bool TILED>
__global__ void __launch_bounds__(256) draw_correct_kernel(const CorrectArgs A) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= A.aw) return;
  for (int y = blockIdx.y; y < A.ah; y += gridDim.y) {
    const float vx = (float)x + 0.5f, vy = (float)y + 0.5f;
    const float spx = A.m[0] * vx + A.m[3] * vy + A.m[6], spy = A.m[1] * vx + A.m[4] * vy + A.m[7];
    const px_t s = rgba_smooth<TILED>(A.b, spx - 0.5f, spy - 0.5f);
    px_t* p = A.a + (size_t)A.aw * y + x;
    *p = blend_px_any(A.mode, *p, s);
  }
}

// ---------------------------------------------------------------------------------------------
// gradients (paints.nim:68-248)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxStops = 64;
struct GradientArgs {
  px_t* img;
  int w, h;
  int kind;  // 3 linear, 4 radial, 5 angular
  int n;
  float opacity;
  float h0x, h0y, h1x, h1y;
  float m[9];           // radial: inverse(translate(center) * rotate(angle) * scale(dx, dy))
  float gradientAngle;  // angular
  float pos[kMaxStops];
  float col[kMaxStops][4];
  // RN(1 / (pos[i + 1] - pos[i])) where that difference is inside fdiv_r's range (common.cuh), else 0: the division of
  // the interpolation parameter then costs one multiply and two FMA refinement steps instead of an IEEE division
  float rstep[kMaxStops];
};
PXD uint32_t quant_f(float v) {  // chroma Color channel -> uint8: floor(v * 255 + 0.5) clamped to 0..255, NaN -> 0
  // branch-free: clamp first (fmaxf(NaN, 0) = 0; floor is monotone), then floor through a round-down addition of 2^23
  const float r = fminf(fmaxf(v * 255.0f + 0.5f, 0.0f), 255.0f);
  return __float_as_uint(__fadd_rd(r, 8388608.0f)) & 0x1FFu;
}
PXD float fix_angle(float a) {
  const float pi = (float)3.141592653589793238462643383279502884, tau = (float)(2 * 3.141592653589793238462643383279502884);
  while (a > pi) a -= tau;
  while (a < -pi) a += tau;
  return a;
}
PXD px_t gradient_color(const GradientArgs& G, float t) {  // :68-94
  int index = -1;
  for (int i = 0; i < G.n; i++) {
    if (G.pos[i] < t) index = i;
    if (G.pos[i] > t) break;
  }
  float r, g, b, a;
  if (index == -1) {
    r = G.col[0][0]; g = G.col[0][1]; b = G.col[0][2]; a = G.col[0][3];
  } else if (index + 1 >= G.n) {
    r = G.col[index][0]; g = G.col[index][1]; b = G.col[index][2]; a = G.col[index][3];
  } else {
    const float vn = t - G.pos[index], vd = G.pos[index + 1] - G.pos[index], vr = G.rstep[index];
    const float v = vr != 0.0f ? fdiv_r(vn, vd, vr) : vn / vd;
    const float iv = 1.0f - v;
    r = G.col[index][0] * iv + G.col[index + 1][0] * v;
    g = G.col[index][1] * iv + G.col[index + 1][1] * v;
    b = G.col[index][2] * iv + G.col[index + 1][2] * v;
    a = G.col[index][3] * iv + G.col[index + 1][3] * v;
  }
  a *= G.opacity;
  const uint32_t a8 = quant_f(a), r8 = quant_f(r), g8 = quant_f(g), b8 = quant_f(b);
  if (a8 == 255u) return mk(r8, g8, b8, a8);
  // (c * a + 127) div 255 on two channels at once (16-bit lanes, sums <= 65152)
  const uint32_t rb = div255x2((r8 | (b8 << 16)) * a8 + 0x007F007Fu);
  const uint32_t g_ = div255x2(g8 * a8 + 0x7Fu);
  return (rb & 0xFFu) | ((g_ & 0xFFu) << 8) | (rb & 0x00FF0000u) | (a8 << 24);
}

PXD float gradient_t(const GradientArgs& G, int x, int y) {
  if (G.kind == 3) {  // toLineSpace (:107-113); the horizontal / vertical fast paths evaluate it at (x, 0) / (0, y)
    float qx = (float)x, qy = (float)y;
    if (G.h0y == G.h1y) qy = 0.0f;
    else if (G.h0x == G.h1x) qx = 0.0f;
    const float ddx = G.h1x - G.h0x, ddy = G.h1y - G.h0y;
    const float det = ddx * ddx + ddy * ddy;
    const float num = ddy * (qy - G.h0y) + ddx * (qx - G.h0x);
    // det is the same for every pixel: its reciprocal is hoisted out of the pixel loops by the compiler
    return div_fast_ok(det) ? fdiv_r(num, det, __frcp_rn(det)) : num / det;
  }
  if (G.kind == 4) {
    const float vx = (float)x, vy = (float)y;
    const float mx = G.m[0] * vx + G.m[3] * vy + G.m[6], my = G.m[1] * vx + G.m[4] * vy + G.m[7];
    return sqrtf(mx * mx + my * my);
  }
  const float pi = (float)3.141592653589793238462643383279502884;
  const float ex = (float)x - G.h0x, ey = (float)y - G.h0y;
  const float len = sqrtf(ex * ex + ey * ey);
  const float nx = ex / len, ny = ey / len;
  // arctan2 in float32 = the double result rounded (what a correctly rounded atan2f returns)
  const float angle = (float)atan2((double)ny, (double)nx);
  return fix_angle(angle + G.gradientAngle + pi / 2.0f) / 2.0f / pi + 0.5f;
}

__global__ void __launch_bounds__(256) gradient_kernel(const __grid_constant__ GradientArgs G) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= G.w) return;
  if (G.kind == 3 && G.h0y == G.h1y) {  // horizontal gradient: one colour per column (:115-147)
    const px_t c = gradient_color(G, gradient_t(G, x, 0));
    for (int y = blockIdx.y; y < G.h; y += gridDim.y) G.img[(size_t)G.w * y + x] = c;
    return;
  }
  for (int y = blockIdx.y; y < G.h; y += gridDim.y) G.img[(size_t)G.w * y + x] = gradient_color(G, gradient_t(G, x, y));
}

// Non-solid fillPath / strokePath with a gradient paint (paths.nim:2115-2142) in ONE pass over the canvas: the
// reference fills a canvas-sized `fill` image with the gradient, masks it (`fill.draw(mask, MaskBlend)`) and draws it
// (`image.draw(fill, blendMode)`).  Here the gradient colour of a pixel is evaluated where it is needed — inside the
// blend, by the same gradient_t / gradient_color as gradient_kernel — so the fill image never exists (8 B/px less
// traffic, one launch less) and, for NormalBlend, pixels the mask does not cover cost a mask read and nothing else:
// the work follows the area of the shape, not of the canvas.  paint.opacity scales the mask (`mask.applyOpacity`,
// :2138-2139) and is folded in as floor(m * o / 255).
constexpr int GradGeneric = -1;
__device__ __noinline__ px_t grad_blend_rt(int mode, px_t b, px_t s) {
  px_t r = b;
  PX_DISPATCH_MODE(mode, r = blend_px<MODE>(b, s));
  return r;
}
struct GradBlendArgs {
  GradientArgs G;       // G.img = the canvas (dst)
  const uint8_t* mask;  // RGBX (alpha used) or A8, canvas-sized
  int maskBpp;
  uint32_t maskOpacity;  // 0..255; 255 = leave the mask as it is
  int mode;
};
template <int MODE>
__global__ void __launch_bounds__(256) gradient_blend_kernel(const __grid_constant__ GradBlendArgs A) {
  const GradientArgs& G = A.G;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x0 >= G.w) return;
  const int nx = min(4, G.w - x0);
  const bool vec = nx == 4 && (G.w & 3) == 0;
  for (int y = blockIdx.y; y < G.h; y += gridDim.y) {
    const size_t idx = (size_t)G.w * y + x0;
    uint32_t m[4] = {0u, 0u, 0u, 0u};
    if (A.maskBpp == 4) {
      const px_t* mp = reinterpret_cast<const px_t*>(A.mask) + idx;
      if (vec) {
        const uint4 v = *reinterpret_cast<const uint4*>(mp);
        m[0] = v.x >> 24; m[1] = v.y >> 24; m[2] = v.z >> 24; m[3] = v.w >> 24;
      } else {
        for (int k = 0; k < nx; k++) m[k] = mp[k] >> 24;
      }
    } else {
      const uint8_t* mp = A.mask + idx;
      if (vec) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(mp);
        m[0] = v & 255u; m[1] = (v >> 8) & 255u; m[2] = (v >> 16) & 255u; m[3] = v >> 24;
      } else {
        for (int k = 0; k < nx; k++) m[k] = mp[k];
      }
    }
    if (A.maskOpacity != 255u) {
#pragma unroll
      for (int k = 0; k < 4; k++) m[k] = (m[k] * A.maskOpacity) / 255u;  // mul_div255 on the alpha channel
    }
    // blendNormal with a transparent source leaves the backdrop alone (sse2.nim:590-616: d * 255 div 255 + 0)
    if (MODE == NormalBlend && (m[0] | m[1] | m[2] | m[3]) == 0u) continue;
    px_t* dp = G.img + idx;
    uint4 dv = make_uint4(0u, 0u, 0u, 0u);
    uint32_t* d = reinterpret_cast<uint32_t*>(&dv);
    if (vec) dv = *reinterpret_cast<const uint4*>(dp);
    else for (int k = 0; k < nx; k++) d[k] = dp[k];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (k >= nx) continue;
      if (MODE == NormalBlend && m[k] == 0u) continue;
      px_t sx = 0u;
      if (m[k] != 0u) {
        sx = gradient_color(G, gradient_t(G, x0 + k, y));
        if (m[k] != 255u) sx = mul_div255(sx, m[k]);
      }
      if (MODE == NormalBlend) d[k] = line_normal(d[k], sx);   // images.nim:485-500 row kernels, as blend_rect
      else if (MODE == MaskBlend) d[k] = line_mask(d[k], sx);
      else if (A.mode == OverwriteBlend) d[k] = sx;
      else d[k] = grad_blend_rt(A.mode, d[k], sx);
    }
    if (vec) *reinterpret_cast<uint4*>(dp) = dv;
    else for (int k = 0; k < nx; k++) dp[k] = d[k];
  }
}

// ---------------------------------------------------------------------------------------------
// host: vmath Mat3 pieces (column-major m[c*3+r]) and the draw() decision logic
// ---------------------------------------------------------------------------------------------
struct M3 {
  float m[9];
};
static inline void mul_v(const M3& a, float x, float y, float& ox, float& oy) {
  ox = a.m[0] * x + a.m[3] * y + a.m[6];
  oy = a.m[1] * x + a.m[4] * y + a.m[7];
}
static M3 mul_m(const M3& a, const M3& b) {
  M3 r;
  for (int c = 0; c < 3; c++)
    for (int row = 0; row < 3; row++)
      r.m[c * 3 + row] = b.m[c * 3 + 0] * a.m[0 * 3 + row] + b.m[c * 3 + 1] * a.m[1 * 3 + row] + b.m[c * 3 + 2] * a.m[2 * 3 + row];
  return r;
}
static M3 scale_m(float x, float y) {
  M3 r = {{x, 0, 0, 0, y, 0, 0, 0, 1}};
  return r;
}
static M3 inverse_m(const M3& a) {
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
static inline float vlen(float x, float y) { return sqrtf(x * x + y * y); }
static inline float fractional_v(float v) {
  v = fabsf(v);
  return v - truncf(v);
}
static inline long long f2i_h(float f) {
  if (!(f > -9.2e18f && f < 9.2e18f)) return INT64_MIN;
  return (long long)f;
}

// temporary images of the minify / magnify chain: stream-ordered allocations
struct TempImage {
  px_t* d = nullptr;
  int w = 0, h = 0;
};
static int temp_alloc(TempImage& t, int w, int h) {
  t.w = w;
  t.h = h;
  PX_CUDA(cudaMallocAsync(&t.d, (size_t)w * h * 4, rt().stream));
  return 0;
}
static void temp_free(TempImage& t) {
  if (t.d) cudaFreeAsync(t.d, rt().stream);
  t.d = nullptr;
}

static dim3 grid_2d(int w, int h) {
  dim3 g((w + 255) / 256, 1);
  int gy = rt().num_sms * 8 / (int)g.x;
  g.y = std::max(1, std::min(gy, h));
  return g;
}

static int minify_once(const px_t* src, int sw, int sh, TempImage& out) {
  const int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
  if (int rc = temp_alloc(out, dw, dh)) return rc;
  minify_kernel<<<grid_2d(dw, dh), 256, 0, rt().stream>>>(src, sw, sh, out.d, dw, dh);
  PX_LAUNCHED();
  return 0;
}
static int magnify_pow(const px_t* src, int sw, int sh, int power, TempImage& out) {
  const int dw = sw << power, dh = sh << power;
  if (int rc = temp_alloc(out, dw, dh)) return rc;
  magnify_kernel<<<grid_2d(dw, dh), 256, 0, rt().stream>>>(src, sw, out.d, dw, dh, power);
  PX_LAUNCHED();
  return 0;
}

static int launch_smooth(Image* a, const px_t* b, int bw, int bh, const M3& transform, int mode) {
  SmoothArgs S;
  S.a = (px_t*)a->data; S.aw = a->w; S.ah = a->h;
  S.b.d = b; S.b.w = bw; S.b.h = bh;
  S.mode = mode;
  const float cxs[4] = {0.0f, (float)bw, (float)bw, 0.0f}, cys[4] = {0.0f, 0.0f, (float)bh, (float)bh};
  for (int k = 0; k < 4; k++) mul_v(transform, cxs[k], cys[k], S.cx[k], S.cy[k]);
  const M3 inv = inverse_m(transform);
  float px, py, ax, ay, bx, by;
  mul_v(inv, 0 + 0.5f, 0 + 0.5f, px, py);
  mul_v(inv, 1 + 0.5f, 0 + 0.5f, ax, ay);
  mul_v(inv, 0 + 0.5f, 1 + 0.5f, bx, by);
  S.px = px; S.py = py;
  S.dxx = ax - px; S.dxy = ay - py;
  S.dyx = bx - px; S.dyy = by - py;
  long long yStart = a->h, yEnd = 0;
  for (int k = 0; k < 4; k++) {
    yStart = std::min<long long>(yStart, f2i_h(floorf(S.cy[k])));
    yEnd = std::max<long long>(yEnd, f2i_h(ceilf(S.cy[k])));
  }
  S.yStart = (int)std::min<long long>(std::max<long long>(yStart, 0), a->h);
  S.yEnd = (int)std::min<long long>(std::max<long long>(yEnd, 0), a->h);
  const int rows = mode == MaskBlend ? a->h : S.yEnd - S.yStart;
  if (rows <= 0) return 0;
  const int blocks = (rows + 7) / 8;
  cudaStream_t st = rt().stream;
  if (mode == NormalBlend) draw_smooth_kernel<NormalBlend><<<blocks, 256, 0, st>>>(S);
  else if (mode == OverwriteBlend) draw_smooth_kernel<OverwriteBlend><<<blocks, 256, 0, st>>>(S);
  else if (mode == MaskBlend) draw_smooth_kernel<MaskBlend><<<blocks, 256, 0, st>>>(S);
  else draw_smooth_kernel<-1><<<blocks, 256, 0, st>>>(S);
  PX_LAUNCHED();
  return 0;
}

int blend_rect_raw(Image* d, const px_t* src, int sw, int sh, int px, int py, int mode);  // blend.cu

static int draw_impl(Image* a, Image* b, const float* mat, int mode) {
  M3 transform;
  memcpy(transform.m, mat, sizeof transform.m);
  const M3 inv = inverse_m(transform);
  float px, py, ax, ay, bx, by;
  mul_v(inv, 0 + 0.5f, 0 + 0.5f, px, py);
  mul_v(inv, 1 + 0.5f, 0 + 0.5f, ax, ay);
  mul_v(inv, 0 + 0.5f, 1 + 0.5f, bx, by);
  float dxx = ax - px, dxy = ay - py, dyx = bx - px, dyy = by - py;
  float filterBy2 = std::max(vlen(dxx, dxy), vlen(dyx, dyy));
  const px_t* src = (const px_t*)b->data;
  int sw = b->w, sh = b->h;
  TempImage cur;
  int rc = 0;
  while (filterBy2 >= 2.0f) {  // :649-655
    TempImage next;
    if ((rc = minify_once(src, sw, sh, next))) break;
    temp_free(cur);
    cur = next;
    src = cur.d; sw = cur.w; sh = cur.h;
    dxx /= 2; dxy /= 2; dyx /= 2; dyy /= 2;
    filterBy2 /= 2;
    transform = mul_m(transform, scale_m(2, 2));
  }
  while (!rc && filterBy2 <= 0.5f) {  // :657-663
    if ((long long)sw * 2 * sh * 2 > (1ll << 28)) {  // the reference would double until it runs out of memory
      rc = fail_pixie("draw: magnified source image too large");
      break;
    }
    TempImage next;
    if ((rc = magnify_pow(src, sw, sh, 1, next))) break;
    temp_free(cur);
    cur = next;
    src = cur.d; sw = cur.w; sh = cur.h;
    dxx *= 2; dxy *= 2; dyx *= 2; dyy *= 2;
    filterBy2 *= 2;
    transform = mul_m(transform, scale_m(1.0f / 2, 1.0f / 2));
  }
  if (!rc) {
    const bool hasRotationOrScaling = !(dxx == 1.0f && dxy == 0.0f && dyx == 0.0f && dyy == 1.0f);
    const bool smooth = !(vlen(dxx, dxy) == 1.0f && vlen(dyx, dyy) == 1.0f && fractional_v(transform.m[6]) == 0.0f &&
                          fractional_v(transform.m[7]) == 0.0f);
    if (hasRotationOrScaling || smooth) rc = launch_smooth(a, src, sw, sh, transform, mode);
    else rc = blend_rect_raw(a, src, sw, sh, (int)transform.m[6], (int)transform.m[7], mode);
  }
  temp_free(cur);
  return rc;
}

static int draw_correct_impl(Image* a, Image* b, const float* mat, int mode, bool tiled) {
  M3 transform;
  memcpy(transform.m, mat, sizeof transform.m);
  M3 inv = inverse_m(transform);
  float px, py, ax, ay, bx, by;
  mul_v(inv, 0 + 0.5f, 0 + 0.5f, px, py);
  mul_v(inv, 1 + 0.5f, 0 + 0.5f, ax, ay);
  mul_v(inv, 0 + 0.5f, 1 + 0.5f, bx, by);
  float dxx = ax - px, dxy = ay - py, dyx = bx - px, dyy = by - py;
  float filterBy2 = std::max(vlen(dxx, dxy), vlen(dyx, dyy));
  const px_t* src = (const px_t*)b->data;
  int sw = b->w, sh = b->h;
  TempImage cur;
  int rc = 0;
  while (filterBy2 >= 2.0f) {
    TempImage next;
    if ((rc = minify_once(src, sw, sh, next))) break;
    temp_free(cur);
    cur = next;
    src = cur.d; sw = cur.w; sh = cur.h;
    dxx /= 2; dxy /= 2; dyx /= 2; dyy /= 2;
    filterBy2 /= 2;
    inv = mul_m(scale_m(0.5f, 0.5f), inv);
  }
  while (!rc && filterBy2 <= 0.5f) {
    if ((long long)sw * 2 * sh * 2 > (1ll << 28)) {  // the reference would double until it runs out of memory
      rc = fail_pixie("draw: magnified source image too large");
      break;
    }
    TempImage next;
    if ((rc = magnify_pow(src, sw, sh, 1, next))) break;
    temp_free(cur);
    cur = next;
    src = cur.d; sw = cur.w; sh = cur.h;
    dxx *= 2; dxy *= 2; dyx *= 2; dyy *= 2;
    filterBy2 *= 2;
    inv = mul_m(scale_m(2, 2), inv);
  }
  if (!rc) {
    CorrectArgs C;
    C.a = (px_t*)a->data; C.aw = a->w; C.ah = a->h;
    C.b.d = src; C.b.w = sw; C.b.h = sh;
    memcpy(C.m, inv.m, sizeof C.m);
    C.mode = mode;
    if (tiled) draw_correct_kernel<true><<<grid_2d(a->w, a->h), 256, 0, rt().stream>>>(C);
    else draw_correct_kernel<false><<<grid_2d(a->w, a->h), 256, 0, rt().stream>>>(C);
    rt().launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail_cuda(e, "kernel launch");
  }
  temp_free(cur);
  return rc;
}

static int two_images(pixie_image_t dsth, pixie_image_t srch, Image** d, Image** s, int mode) {
  if (int rc = ensure_init()) return rc;
  if (mode < 0 || mode >= NumBlendModes) return fail_pixie("invalid blend mode");
  *d = find_image(dsth);
  *s = find_image(srch);
  if (!*d || !*s) return 1;
  if ((*d)->bpp != 4 || (*s)->bpp != 4 || (*d)->layers != 1 || (*s)->layers != 1)
    return fail_pixie("draw needs single-layer RGBX images");
  if ((*d)->data == (*s)->data) return fail_pixie("draw: dst and src must be different images");
  return 0;
}

}  // namespace pixie

using namespace pixie;

extern "C" {

int pixie_cuda_draw(pixie_image_t dst, pixie_image_t src, const float* mat, int mode) {
  PX_API_GUARD;
  Image *d, *s;
  if (int rc = two_images(dst, src, &d, &s, mode)) return rc;
  return draw_impl(d, s, mat, mode);
}

int pixie_cuda_draw_tiled(pixie_image_t dst, pixie_image_t src, const float* mat, int mode) {
  PX_API_GUARD;
  Image *d, *s;
  if (int rc = two_images(dst, src, &d, &s, mode)) return rc;
  return draw_correct_impl(d, s, mat, mode, true);
}

int pixie_cuda_draw_correct(pixie_image_t dst, pixie_image_t src, const float* mat, int mode) {
  PX_API_GUARD;
  Image *d, *s;
  if (int rc = two_images(dst, src, &d, &s, mode)) return rc;
  return draw_correct_impl(d, s, mat, mode, false);
}

int pixie_cuda_minify_by2(pixie_image_t src, int power, pixie_image_t* out) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  if (power < 0) return fail_pixie("Cannot minifyBy2 with negative power");
  Image* s = find_image(src);
  if (!s) return 1;
  if (s->bpp != 4 || s->layers != 1) return fail_pixie("minifyBy2 needs a single-layer RGBX image");
  int w = s->w, h = s->h;
  for (int i = 0; i < power; i++) {
    w = (w + 1) / 2;
    h = (h + 1) / 2;
  }
  if (int rc = pixie_cuda_image_create(w, h, out)) return rc;
  s = find_image(src);  // the map may have rehashed
  Image* o = find_image(*out);
  if (power == 0) return pixie_cuda_image_copy(*out, src);
  const px_t* cur = (const px_t*)s->data;
  int cw = s->w, ch = s->h;
  TempImage tmp;
  int rc = 0;
  for (int i = 0; i < power && !rc; i++) {
    const int dw = (cw + 1) / 2, dh = (ch + 1) / 2;
    if (i == power - 1) {
      minify_kernel<<<grid_2d(dw, dh), 256, 0, rt().stream>>>(cur, cw, ch, (px_t*)o->data, dw, dh);
      rt().launches++;
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) rc = fail_cuda(e, "kernel launch");
    } else {
      TempImage next;
      rc = minify_once(cur, cw, ch, next);
      temp_free(tmp);
      tmp = next;
      cur = tmp.d;
    }
    cw = dw;
    ch = dh;
  }
  temp_free(tmp);
  return rc;
}

int pixie_cuda_magnify_by2(pixie_image_t src, int power, pixie_image_t* out) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  if (power < 0) return fail_pixie("Cannot magnifyBy2 with negative power");
  Image* s = find_image(src);
  if (!s) return 1;
  if (s->bpp != 4 || s->layers != 1) return fail_pixie("magnifyBy2 needs a single-layer RGBX image");
  if (power > 15 || ((long long)s->w << power) * ((long long)s->h << power) > (1ll << 31))
    return fail_pixie("magnifyBy2: result too large");
  const int dw = s->w << power, dh = s->h << power;
  if (int rc = pixie_cuda_image_create(dw, dh, out)) return rc;
  s = find_image(src);
  Image* o = find_image(*out);
  magnify_kernel<<<grid_2d(dw, dh), 256, 0, rt().stream>>>((const px_t*)s->data, s->w, (px_t*)o->data, dw, dh, power);
  PX_LAUNCHED();
  return 0;
}

static int gradient_setup(GradientArgs& G, Image* im, int kind, const float* handles, int n_handles, const float* stop_pos,
                          const float* stop_rgba, int n_stops, float opacity);

int pixie_cuda_fill_gradient(pixie_image_t image, int kind, const float* handles, int n_handles, const float* stop_pos,
                             const float* stop_rgba, int n_stops, float opacity) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  Image* im = find_image(image);
  if (!im) return 1;
  opacity = opacity < 0.0f ? 0.0f : (opacity > 1.0f ? 1.0f : opacity);
  if (opacity == 0.0f) return 0;
  GradientArgs G;
  if (int rc = gradient_setup(G, im, kind, handles, n_handles, stop_pos, stop_rgba, n_stops, opacity)) return rc;
  gradient_kernel<<<grid_2d(im->w, im->h), 256, 0, rt().stream>>>(G);
  PX_LAUNCHED();
  return 0;
}

int pixie_cuda_fill_gradient_masked(pixie_image_t image, pixie_image_t maskh, int kind, const float* handles, int n_handles,
                                    const float* stop_pos, const float* stop_rgba, int n_stops, float opacity, int blend_mode) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  Image* im = find_image(image);
  Image* mk_ = find_image(maskh);
  if (!im || !mk_) return 1;
  if (blend_mode < 0 || blend_mode >= NumBlendModes) return fail_pixie("invalid blend mode");
  if (mk_->w != im->w || mk_->h != im->h || mk_->layers != 1 || (mk_->bpp != 4 && mk_->bpp != 1))
    return fail_pixie("mask must be a single-layer RGBX or A8 image of the canvas size");
  if (mk_->data == im->data) return fail_pixie("fill_gradient_masked: mask and image must differ");
  opacity = opacity < 0.0f ? 0.0f : (opacity > 1.0f ? 1.0f : opacity);
  if (opacity == 0.0f) return 0;  // paths.nim:2096-2097
  GradBlendArgs A;
  // the gradient itself is filled at opacity 1 (:2123-2136); paint.opacity scales the mask (:2138-2139, images.nim:261-277)
  if (int rc = gradient_setup(A.G, im, kind, handles, n_handles, stop_pos, stop_rgba, n_stops, 1.0f)) return rc;
  A.mask = mk_->data; A.maskBpp = mk_->bpp; A.mode = blend_mode;
  A.maskOpacity = (uint32_t)(uint16_t)(int64_t)roundf(255 * opacity);
  if (A.maskOpacity > 255u) return fail_pixie("opacity out of range");
  const dim3 grid = grid_2d((im->w + 3) / 4, im->h);
  if (blend_mode == NormalBlend) gradient_blend_kernel<NormalBlend><<<grid, 256, 0, rt().stream>>>(A);
  else if (blend_mode == MaskBlend) gradient_blend_kernel<MaskBlend><<<grid, 256, 0, rt().stream>>>(A);
  else gradient_blend_kernel<GradGeneric><<<grid, 256, 0, rt().stream>>>(A);
  PX_LAUNCHED();
  return 0;
}

static int gradient_setup(GradientArgs& G, Image* im, int kind, const float* handles, int n_handles, const float* stop_pos,
                          const float* stop_rgba, int n_stops, float opacity) {
  if (im->bpp != 4 || im->layers != 1) return fail_pixie("fillGradient needs a single-layer RGBX image");
  if (kind < 3 || kind > 5) return fail_pixie("Paint must be a gradient");  // paints.nim:247-248
  if (kind == 3 && n_handles != 2) return fail_pixie("Linear gradient requires 2 handles");
  if (kind == 4 && n_handles != 3) return fail_pixie("Radial gradient requires 3 handles");
  if (kind == 5 && n_handles != 3) return fail_pixie("Angular gradient requires 2 handles");
  if (n_stops == 0) return fail_pixie("Gradient must have at least 1 color stop");
  if (n_stops < 0 || n_stops > kMaxStops) return fail_pixie("too many gradient stops (64 at most)");
  memset(&G, 0, sizeof G);
  G.img = (px_t*)im->data; G.w = im->w; G.h = im->h; G.kind = kind; G.n = n_stops; G.opacity = opacity;
  memcpy(G.pos, stop_pos, (size_t)n_stops * 4);
  memcpy(G.col, stop_rgba, (size_t)n_stops * 16);
  for (int i = 0; i + 1 < n_stops; i++) {
    volatile float d = G.pos[i + 1] - G.pos[i];  // one IEEE subtraction, as the kernel's
    const float dd = d;
    G.rstep[i] = (dd > 1e-18f && dd < 1e18f) ? 1.0f / dd : 0.0f;
  }
  G.h0x = handles[0]; G.h0y = handles[1]; G.h1x = handles[2]; G.h1y = handles[3];
  const float pi = (float)3.141592653589793238462643383279502884, tau = (float)(2 * 3.141592653589793238462643383279502884);
  auto fix = [&](float a) {
    while (a > pi) a -= tau;
    while (a < -pi) a += tau;
    return a;
  };
  if (kind == 4) {  // :180-192
    const float cx = handles[0], cy = handles[1], ex = handles[2], ey = handles[3], kx = handles[4], ky = handles[5];
    const float distanceX = vlen(cx - ex, cy - ey), distanceY = vlen(cx - kx, cy - ky);
    const float nl = vlen(cx - ex, cy - ey);
    const float nx = (cx - ex) / nl, ny = (cy - ey) / nl;
    const float ang = fix(atan2f(ny, nx));
    const float s = sinf(ang), c = cosf(ang);
    const M3 tr = {{1, 0, 0, 0, 1, 0, cx, cy, 1}}, rot = {{c, s, 0, -s, c, 0, 0, 0, 1}};
    const M3 m = inverse_m(mul_m(mul_m(tr, rot), scale_m(distanceX, distanceY)));
    memcpy(G.m, m.m, sizeof G.m);
  } else if (kind == 5) {  // :210-216
    const float ex = handles[2] - handles[0], ey = handles[3] - handles[1];
    const float nl = vlen(ex, ey);
    G.gradientAngle = fix(atan2f(ey / nl, ex / nl));
  }
  return 0;
}

}  // extern "C"

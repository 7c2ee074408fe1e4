// K0 — path commands -> segments on the device: what fillPath / strokePath do before fillShapes.
//
//   commandsToShapes (treeform/pixie src/pixie/paths.nim:654-1057)  lines, quadratics, cubics (the adaptive halving
//                    loops of addCubic :676-722 and addQuadratic :724-766, addSegment's zero-length rule :669-674)
//   strokeShapes     (:1922-2082)  butt / square caps, miter / bevel joins (makeRect :1943-1958, addJoin :1960-2010)
//   transform + shapesToSegments (:1092-1096, :1059-1090)  y quantised to 1/256, horizontals dropped, winding
//
// Everything is IEEE float32 + - * / sqrt with one rounding per operation (-fmad=false -prec-div=true -prec-sqrt=true),
// in the reference's order of operations, so the segment list equals the host's.  atan2 (only its sign and two
// threshold comparisons decide a join's shape) is evaluated in double and rounded to float32, which is what a
// correctly rounded arctan2 returns.  Arcs, round caps / joins (sin, cos, arccos of the host libm) and dashes are NOT
// done here: the caller flattens those paths on the host and passes their segments through (kind 2).
//
// Parallel decomposition.  A *primitive* is one drawing command resolved to absolute control points (resolve_kernel:
// one thread per path walks its command stream once — relative coordinates and smooth control points chain through
// float additions, so that walk is sequential — and also emits the implicit closing lines).  After that every
// primitive is independent: count_kernel runs the halving loop of each curve once to count what it produces,
// exclusive scans turn counts into offsets, emit_kernel runs the loops again and writes.  Fill paths write their
// segments directly (every polygon edge is one addSegment call and closed polygons need no wrap-around edge); stroke
// paths write the polygon points, stroke_count / stroke_emit then give every point a thread (cap, rectangle of the
// edge that ends there, join).  bounds_kernel reduces the path bounds (computeBounds :1098-1117) for the fill headers.
#include <cfloat>
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace pixie {

namespace {

enum CmdKind { Close = 0, Move, Line, HLine, VLine, Cubic, SCubic, Quad, TQuad, Arc, RMove, RLine, RHLine, RVLine, RCubic, RSCubic, RQuad, RTQuad, RArc };
enum PrimType { PrimNone = 0, PrimLine = 1, PrimQuad = 2, PrimCubic = 3, PrimRaw = 4 };

struct DPath {  // device copy of pixie_path_desc + slots
  int cmdBegin, cmdEnd;
  int primBase, primCap;
  int kind;  // 0 fill, 1 stroke, 2 raw segments
  int lineCap, lineJoin;
  float halfStroke, miterAngleLimit, errorMarginSq, minArea;
  float m[9];
  int identity;
};

struct __align__(16) Prim {
  float ax, ay, c1x, c1y, c2x, c2y, tx, ty;
  int type;        // PrimType
  int path;
  int shapeBegin;  // first primitive slot of the shape this one belongs to
  int shapeEnd;    // valid at slot shapeBegin: one past the shape's last primitive slot
};

struct V2 {
  float x, y;
};
PXD V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
PXD V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
PXD V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
PXD V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
PXD V2 operator/(V2 a, float s) { return v2(a.x / s, a.y / s); }
PXD bool veq(V2 a, V2 b) { return a.x == b.x && a.y == b.y; }
PXD float length_sq(V2 a) { return a.x * a.x + a.y * a.y; }
PXD float length(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
PXD V2 normalize(V2 a) { return a / length(a); }

PXD int param_count(int kind) {  // paths.nim:73-81
  switch (kind) {
    case Close: return 0;
    case Move: case Line: case RMove: case RLine: case TQuad: case RTQuad: return 2;
    case HLine: case VLine: case RHLine: case RVLine: return 1;
    case Cubic: case RCubic: return 6;
    case SCubic: case RSCubic: case Quad: case RQuad: return 4;
    default: return 7;
  }
}
PXD bool is_cubic_kind(int k) { return k == Cubic || k == SCubic || k == RCubic || k == RSCubic; }
PXD bool is_quad_kind(int k) { return k == Quad || k == TQuad || k == RQuad || k == RTQuad; }

// ---------------------------------------------------------------------------------------------
// resolve: commands -> primitives (one thread per path)
// ---------------------------------------------------------------------------------------------
// One WARP per path: the walk itself is sequential (lane 0), but a lone thread reading its commands from HBM pays a
// full memory round trip per command (115 us for the tiger's 305 paths); here the warp stages the command stream in
// shared memory 512 floats at a time and lane 0 reads from there.
constexpr int kResolveChunk = 512;
constexpr int kResolveWarps = 4;
__global__ void __launch_bounds__(kResolveWarps * 32) resolve_kernel(const DPath* __restrict__ paths, int numPaths, const float* __restrict__ cmdsG,
                                                                     Prim* __restrict__ prims, int* __restrict__ err) {
  __shared__ float sCmd[kResolveWarps][kResolveChunk + 8];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int pi = blockIdx.x * kResolveWarps + wi;
  if (pi >= numPaths) return;
  const DPath P = paths[pi];
  float* cmds = sCmd[wi];
  int chunkBase = P.cmdBegin;  // cmds[k] holds command float chunkBase + k
  auto stage = [&](int from) {  // the warp loads floats [from, from + chunk + 8) of the stream (a command has <= 8 floats)
    __syncwarp();
    for (int k = lane; k < kResolveChunk + 8; k += 32) cmds[k] = from + k < P.cmdEnd ? cmdsG[from + k] : 0.0f;
    chunkBase = from;
    __syncwarp();
  };
  int slot = P.primBase;
  const int slotEnd = P.primBase + P.primCap;
  if (P.kind == 2) {  // pre-flattened segments: one pass-through primitive
    Prim q;
    memset(&q, 0, sizeof(q));
    q.type = PrimRaw; q.path = pi; q.shapeBegin = slot; q.shapeEnd = slot + 1;
    if (lane == 0) prims[slot] = q;
    slot++;
    for (; slot < slotEnd; slot++) { q.type = PrimNone; q.shapeBegin = slot; q.shapeEnd = slot + 1; if (lane == 0) prims[slot] = q; }
    return;
  }
  // every lane walks the stream (same values, uniform control flow); lane 0 writes
  stage(P.cmdBegin);
  const bool closeSubpaths = P.kind == 0;
  V2 start = v2(0.f, 0.f), at = v2(0.f, 0.f), prevCtrl = v2(0.f, 0.f), prevCtrl2 = v2(0.f, 0.f);
  int prevKind = Move;
  int shapeBegin = slot;
  auto put = [&](int type, V2 a, V2 c1, V2 c2, V2 to) {
    Prim q;
    q.ax = a.x; q.ay = a.y; q.c1x = c1.x; q.c1y = c1.y; q.c2x = c2.x; q.c2y = c2.y; q.tx = to.x; q.ty = to.y;
    q.type = type; q.path = pi; q.shapeBegin = shapeBegin; q.shapeEnd = 0;
    if (lane == 0 && slot < slotEnd) prims[slot] = q;
    slot++;
  };
  auto end_shape = [&]() {
    if (lane == 0 && slot > shapeBegin && shapeBegin < slotEnd) prims[shapeBegin].shapeEnd = min(slot, slotEnd);
    shapeBegin = slot;
  };
  int i = P.cmdBegin;
  while (i < P.cmdEnd) {
    if (i - chunkBase >= kResolveChunk) stage(i);  // warp-uniform: every lane tracks i
    const int kind = (int)cmds[i - chunkBase];
    i++;
    const float* c = cmds + (i - chunkBase);
    switch (kind) {
      case Move:
        // `if shape.len > 0: if closeSubpaths: addSegment(at, start)` — when the shape is still empty at == start and
        // the closing line is degenerate, so it can be emitted unconditionally
        if (closeSubpaths) put(PrimLine, at, at, at, start);
        end_shape();
        at = v2(c[0], c[1]);
        start = at;
        break;
      case RMove:  // (:955-961: no closing segment here)
        end_shape();
        at = v2(at.x + c[0], at.y + c[1]);
        start = at;
        break;
      case Line: { V2 to = v2(c[0], c[1]); put(PrimLine, at, at, at, to); at = to; } break;
      case HLine: { V2 to = v2(c[0], at.y); put(PrimLine, at, at, at, to); at = to; } break;
      case VLine: { V2 to = v2(at.x, c[0]); put(PrimLine, at, at, at, to); at = to; } break;
      case RLine: { V2 to = v2(at.x + c[0], at.y + c[1]); put(PrimLine, at, at, at, to); at = to; } break;
      case RHLine: { V2 to = v2(at.x + c[0], at.y); put(PrimLine, at, at, at, to); at = to; } break;
      case RVLine: { V2 to = v2(at.x, at.y + c[0]); put(PrimLine, at, at, at, to); at = to; } break;
      case Cubic: {
        V2 c1 = v2(c[0], c[1]), c2 = v2(c[2], c[3]), to = v2(c[4], c[5]);
        put(PrimCubic, at, c1, c2, to); at = to; prevCtrl2 = c2;
      } break;
      case SCubic: {
        V2 c2 = v2(c[0], c[1]), to = v2(c[2], c[3]);
        V2 c1 = is_cubic_kind(prevKind) ? at * 2.0f - prevCtrl2 : at;
        put(PrimCubic, at, c1, c2, to); at = to; prevCtrl2 = c2;
      } break;
      case RCubic: {
        V2 c1 = v2(at.x + c[0], at.y + c[1]), c2 = v2(at.x + c[2], at.y + c[3]), to = v2(at.x + c[4], at.y + c[5]);
        put(PrimCubic, at, c1, c2, to); at = to; prevCtrl2 = c2;
      } break;
      case RSCubic: {
        V2 c2 = v2(at.x + c[0], at.y + c[1]), to = v2(at.x + c[2], at.y + c[3]);
        V2 c1 = is_cubic_kind(prevKind) ? at * 2.0f - prevCtrl2 : at;
        put(PrimCubic, at, c1, c2, to); at = to; prevCtrl2 = c2;
      } break;
      case Quad: {
        V2 ctrl = v2(c[0], c[1]), to = v2(c[2], c[3]);
        put(PrimQuad, at, ctrl, ctrl, to); at = to; prevCtrl = ctrl;
      } break;
      case TQuad: {
        V2 to = v2(c[0], c[1]);
        V2 ctrl = is_quad_kind(prevKind) ? at * 2.0f - prevCtrl : at;
        put(PrimQuad, at, ctrl, ctrl, to); at = to; prevCtrl = ctrl;
      } break;
      case RQuad: {
        V2 ctrl = v2(at.x + c[0], at.y + c[1]), to = v2(at.x + c[2], at.y + c[3]);
        put(PrimQuad, at, ctrl, ctrl, to); at = to; prevCtrl = ctrl;
      } break;
      case RTQuad: {
        V2 to = v2(at.x + c[0], at.y + c[1]);
        V2 ctrl = is_quad_kind(prevKind) ? at * 2.0f - prevCtrl : at;
        put(PrimQuad, at, ctrl, ctrl, to); at = to; prevCtrl = ctrl;
      } break;
      case Close:
        if (!veq(at, start)) {
          put(PrimLine, at, at, at, start);
          at = start;
        }
        end_shape();
        break;
      default:  // arcs are flattened on the host (kind 2); an unknown command is the reference's "Invalid path command"
        if (lane == 0) atomicMax(err, kind == Arc || kind == RArc ? 2 : 3);
        i = P.cmdEnd;
        break;
    }
    if (i < P.cmdEnd) i += param_count(kind);
    prevKind = kind;
  }
  if (closeSubpaths) put(PrimLine, at, at, at, start);
  end_shape();
  if (lane == 0 && slot > slotEnd) atomicMax(err, 4);  // more primitives than the caller reserved slots for
  Prim q;
  memset(&q, 0, sizeof(q));
  for (slot += lane; slot < slotEnd; slot += 32) { q.type = PrimNone; q.path = pi; q.shapeBegin = slot; q.shapeEnd = slot + 1; prims[slot] = q; }
}

// ---------------------------------------------------------------------------------------------
// the flattening loops; `seg(prev, next)` is called for every addSegment whose ends differ
// ---------------------------------------------------------------------------------------------
PXD V2 cubic_point(V2 at, V2 c1, V2 c2, V2 to, float t) {  // compute (:678-686)
  const float t2 = t * t, t3 = t2 * t;
  return at * (-t3 + 3.0f * t2 - 3.0f * t + 1.0f) + c1 * (3.0f * t3 - 6.0f * t2 + 3.0f * t) + c2 * (-3.0f * t3 + 3.0f * t2) + to * (t3);
}
PXD V2 cubic_deriv(V2 at, V2 c1, V2 c2, V2 to, float t) {  // computeDeriv (:688-694)
  const float t2 = t * t;
  return at * (-3.0f * t2 + 6.0f * t - 3.0f) + c1 * (9.0f * t2 - 12.0f * t + 3.0f) + c2 * (-9.0f * t2 + 6.0f * t) + to * (3.0f * t2);
}
PXD V2 quad_point(V2 at, V2 ctrl, V2 to, float t) {  // (:727-732)
  const float t2 = t * t;
  return at * (t2 - 2.0f * t + 1.0f) + ctrl * (-2.0f * t2 + 2.0f * t) + to * t2;
}

template <typename F>
PXD bool add_segment(V2 at, V2 to, F& seg) {  // addSegment (:669-674)
  const V2 d = at - to;
  if (d.x != 0.0f || d.y != 0.0f) {
    seg(at, to);
    return true;
  }
  return false;
}

// returns false where the reference raises "Unable to discretize ..."
template <typename F>
PXD bool flatten_prim(const Prim& q, float errorMarginSq, F& seg) {
  const V2 at = v2(q.ax, q.ay), to = v2(q.tx, q.ty);
  if (q.type == PrimLine) {
    add_segment(at, to, seg);
    return true;
  }
  if (q.type == PrimCubic) {  // addCubic (:696-722)
    const V2 c1 = v2(q.c1x, q.c1y), c2 = v2(q.c2x, q.c2y);
    float t = 0.0f, step = 1.0f;
    V2 prev = at;
    V2 next = cubic_point(at, c1, c2, to, t + step);
    V2 halfway = cubic_point(at, c1, c2, to, t + step / 2.0f);
    while (true) {
      if (step <= FLT_EPSILON) return false;
      const V2 midpoint = (prev + next) / 2.0f;
      const V2 lineTangent = midpoint - prev;
      const V2 curveTangent = cubic_deriv(at, c1, c2, to, t + step / 2.0f);
      const V2 curveTangentScaled = normalize(curveTangent) * length(lineTangent);
      const float error = length_sq(midpoint - halfway);
      const float errorTangent = length_sq(lineTangent - curveTangentScaled);
      if (error + errorTangent > errorMarginSq) {
        next = halfway;
        halfway = cubic_point(at, c1, c2, to, t + step / 4.0f);
        step /= 2.0f;
      } else {
        add_segment(prev, next, seg);
        t += step;
        if (t == 1.0f) break;
        prev = next;
        step = fminf(step * 2.0f, 1.0f - t);
        next = cubic_point(at, c1, c2, to, t + step);
        halfway = cubic_point(at, c1, c2, to, t + step / 2.0f);
      }
    }
    return true;
  }
  if (q.type == PrimQuad) {  // addQuadratic (:734-766)
    const V2 ctrl = v2(q.c1x, q.c1y);
    float t = 0.0f, step = 1.0f;
    V2 prev = at;
    V2 next = quad_point(at, ctrl, to, t + step);
    V2 halfway = quad_point(at, ctrl, to, t + step / 2.0f);
    bool halfStepping = false;
    while (true) {
      if (step <= FLT_EPSILON) return false;
      const V2 midpoint = (prev + next) / 2.0f;
      const float error = length_sq(midpoint - halfway);
      if (error > errorMarginSq) {
        next = halfway;
        halfway = quad_point(at, ctrl, to, t + step / 4.0f);
        halfStepping = true;
        step /= 2.0f;
      } else {
        add_segment(prev, next, seg);
        t += step;
        if (t == 1.0f) break;
        prev = next;
        if (halfStepping) step = fminf(step, 1.0f - t);
        else step = fminf(step * 2.0f, 1.0f - t);
        next = quad_point(at, ctrl, to, t + step);
        halfway = quad_point(at, ctrl, to, t + step / 2.0f);
      }
    }
    return true;
  }
  return true;
}

// transform (:1092-1096) + the y quantisation of shapesToSegments (:1064-1071)
PXD float quantize_y(float v) {  // vmath quantize(v, 1 / 256) = sign(v) * floor(|v| / n) * n
  const float n = 1.0f / 256.0f;
  const float sg = v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f);
  return sg * floorf(fabsf(v) / n) * n;
}
PXD V2 xform(const DPath& P, V2 v) {
  if (P.identity) return v;
  return v2(P.m[0] * v.x + P.m[3] * v.y + P.m[6], P.m[1] * v.x + P.m[4] * v.y + P.m[7]);
}
// one polygon edge (already transformed) -> at most one segment (:1072-1090); returns 1 when it is kept
PXD int edge_segment(V2 a, V2 b, float4* seg, int16_t* wind) {
  const float ya = quantize_y(a.y), yb = quantize_y(b.y);
  if (ya == yb) return 0;  // horizontal after quantisation
  if (seg) {
    if (ya > yb) {
      *seg = make_float4(b.x, yb, a.x, ya);
      *wind = (int16_t)-1;
    } else {
      *seg = make_float4(a.x, ya, b.x, yb);
      *wind = (int16_t)1;
    }
  }
  return 1;
}

// ---------------------------------------------------------------------------------------------
// count / emit over primitives
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) count_kernel_f(const DPath* __restrict__ paths, const Prim* __restrict__ prims, int numPrims,
                                                      const int* __restrict__ rawCount, int* __restrict__ cntPts, int* __restrict__ cntSeg,
                                                      int* __restrict__ err) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= numPrims) return;
  const Prim q = prims[k];
  int pts = 0, segs = 0;
  if (q.type == PrimRaw) {
    segs = rawCount[q.path];
  } else if (q.type != PrimNone) {
    const DPath& P = paths[q.path];
    const bool fill = P.kind == 0;
    auto seg = [&](V2 a, V2 b) {
      pts++;
      if (fill) segs += edge_segment(xform(P, a), xform(P, b), nullptr, nullptr);
    };
    if (!flatten_prim(q, P.errorMarginSq, seg)) atomicMax(err, 1);
    if (fill) pts = 0;
  }
  cntPts[k] = pts;
  cntSeg[k] = segs;
}

// stroke paths: the first primitive of every shape that produces points also pushes its `at` (shape.len == 0, :672)
__global__ void __launch_bounds__(128) shape_first_kernel(const DPath* __restrict__ paths, int numPaths, const Prim* __restrict__ prims,
                                                          int* __restrict__ cntPts, uint8_t* __restrict__ first) {
  const int pi = blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= numPaths) return;
  const DPath P = paths[pi];
  if (P.kind != 1) return;
  int k = P.primBase;
  const int end = P.primBase + P.primCap;
  while (k < end) {
    int se = prims[k].shapeEnd;
    if (se <= k) se = k + 1;
    for (int j = k; j < se; j++)
      if (cntPts[j] > 0) {
        cntPts[j] += 1;
        first[j] = 1;
        break;
      }
    k = se;
  }
}

struct EmitArgs {
  const DPath* paths;
  const Prim* prims;
  int numPrims;
  const int* ptOff;     // exclusive scan of cntPts  [numPrims + 1]
  const int* segOff;    // exclusive scan of cntSeg  [numPrims + 1]
  const int* seg2Off;   // exclusive scan of the per-point stroke segment counts [numPoints + 1] (null before it exists)
  const uint8_t* first;
  float2* points;
  int* pointPrim;
  float4* segs;
  int16_t* wind;
  const float4* rawSegs;
  const int16_t* rawWind;
  const int* rawBegin;  // per path
};

// first output segment of path `pi`: fill / raw segments of earlier paths + stroke segments of earlier paths
PXD int path_seg_begin(const EmitArgs& A, int pi) {
  const int pb = A.paths[pi].primBase;
  return A.segOff[pb] + (A.seg2Off ? A.seg2Off[A.ptOff[pb]] : 0);
}

// STROKE: the polygon points of stroke paths; otherwise the segments of fill paths (whose place in the output depends
// on the stroke segments of the paths before them, so that launch comes after the stroke counts are scanned)
template <bool STROKE>
__global__ void __launch_bounds__(128) emit_kernel_f(const EmitArgs A) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= A.numPrims) return;
  const Prim q = A.prims[k];
  if (q.type == PrimNone || q.type == PrimRaw) return;
  const DPath& P = A.paths[q.path];
  if ((P.kind == 1) != STROKE) return;
  if (!STROKE) {
    int out = path_seg_begin(A, q.path) + (A.segOff[k] - A.segOff[P.primBase]);
    auto seg = [&](V2 a, V2 b) { out += edge_segment(xform(P, a), xform(P, b), A.segs + out, A.wind + out); };
    flatten_prim(q, P.errorMarginSq, seg);
  } else {
    int out = A.ptOff[k];
    if (A.first[k]) {
      A.points[out] = make_float2(q.ax, q.ay);
      A.pointPrim[out] = k;
      out++;
    }
    auto seg = [&](V2 a, V2 b) {
      A.points[out] = make_float2(b.x, b.y);
      A.pointPrim[out] = k;
      out++;
    };
    flatten_prim(q, P.errorMarginSq, seg);
  }
}

// pass-through segments of host-flattened paths: one block per path
__global__ void __launch_bounds__(256) raw_copy_kernel(const EmitArgs A, int numPaths) {
  const int pi = blockIdx.x;
  if (pi >= numPaths || A.paths[pi].kind != 2) return;
  const int pb = A.paths[pi].primBase;
  const int n = A.segOff[pb + 1] - A.segOff[pb];
  const int dst = path_seg_begin(A, pi), src = A.rawBegin[pi];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    A.segs[dst + i] = A.rawSegs[src + i];
    A.wind[dst + i] = A.rawWind[src + i];
  }
}

// ---------------------------------------------------------------------------------------------
// strokeShapes: one thread per polygon point
// ---------------------------------------------------------------------------------------------
// shapesToSegments over one small closed polygon (already transformed); emits through `out`
template <int N>
PXD int poly_segments(const V2 (&p)[N], int n, float4* segs, int16_t* wind) {
  int cnt = 0;
  V2 vec1 = v2(p[n - 1].x, quantize_y(p[n - 1].y));
  for (int i = 0; i < n; i++) {
    const V2 vec2_ = v2(p[i].x, quantize_y(p[i].y));
    if (i == 0 && veq(vec1, vec2_)) continue;
    V2 sa = vec1, sb = vec2_;
    vec1 = vec2_;
    if (sa.y == sb.y) continue;
    int16_t w = 1;
    if (sa.y > sb.y) {
      const V2 t = sa;
      sa = sb;
      sb = t;
      w = -1;
    }
    if (segs) {
      segs[cnt] = make_float4(sa.x, sa.y, sb.x, sb.y);
      wind[cnt] = w;
    }
    cnt++;
  }
  return cnt;
}

PXD bool line_line_intersects(V2 aa, V2 ab, V2 ba, V2 bb, V2& at) {  // bumpy intersects(Line, Line, at)
  const V2 s1 = ab - aa, s2 = bb - ba;
  const float den = (-s2.x * s1.y + s1.x * s2.y);
  const float t = (s2.x * (aa.y - ba.y) - s2.y * (aa.x - ba.x)) / den;
  if (den == 0.0f) return false;
  at = aa + s1 * t;
  return true;
}
PXD float atan2_rn(float y, float x) { return (float)atan2((double)y, (double)x); }
PXD float fix_angle(float angle) {  // (:54-59)
  const double kPI = 3.141592653589793238462643383279502884;
  float r = angle;
  while ((double)r > kPI) r -= (float)(2.0 * kPI);
  while ((double)r < -kPI) r += (float)(2.0 * kPI);
  return r;
}

struct StrokeOut {
  const DPath* P;
  float4* segs;  // null: count only
  int16_t* wind;
  int n;
  __device__ __forceinline__ void poly(V2 (&p)[5], int np) {
#pragma unroll
    for (int i = 0; i < 5; i++)
      if (i < np) p[i] = xform(*P, p[i]);
    n += poly_segments(p, np, segs ? segs + n : nullptr, segs ? wind + n : nullptr);
  }
};

PXD void make_rect(StrokeOut& o, V2 at, V2 to, float hs) {  // (:1943-1958)
  const V2 tangent = normalize(to - at);
  const V2 normal = v2(tangent.y, tangent.x);
  const V2 a = v2(at.x + normal.x * hs, at.y - normal.y * hs);
  const V2 b = v2(to.x + normal.x * hs, to.y - normal.y * hs);
  const V2 c = v2(to.x - normal.x * hs, to.y + normal.y * hs);
  const V2 d = v2(at.x - normal.x * hs, at.y + normal.y * hs);
  V2 p[5] = {a, b, c, d, a};
  o.poly(p, 5);
}
PXD void add_join(StrokeOut& o, V2 prevPos, V2 pos, V2 nextPos, float hs) {  // (:1960-2010), miter / bevel
  const DPath& P = *o.P;
  const double kPI = 3.141592653589793238462643383279502884;
  const float kEpsilon = (float)(0.0001 * kPI);
  const V2 dn = nextPos - pos, dp = prevPos - pos;
  const float angle = fix_angle(atan2_rn(dn.y, dn.x) - atan2_rn(dp.y, dp.x));
  if (fabs(fabs((double)angle) - kPI) > (double)kEpsilon) {
    V2 a = normalize(pos - prevPos) * hs;
    V2 b = normalize(pos - nextPos) * hs;
    if (angle >= 0.0f) {
      a = v2(-a.y, a.x);
      b = v2(b.y, -b.x);
    } else {
      a = v2(a.y, -a.x);
      b = v2(-b.y, b.x);
    }
    int lineJoin = P.lineJoin;  // 0 miter, 2 bevel (LineJoin, paths.nim:14-17; round = 1 never reaches the device)
    if (lineJoin == 0 && fabsf(angle) < P.miterAngleLimit) lineJoin = 2;
    if (lineJoin == 0) {
      V2 at;
      if (line_line_intersects(prevPos + a, pos + a, nextPos + b, pos + b, at)) {
        const float bisectorLengthSq = length_sq(at - pos);
        const float areaSq = 0.25f * (length_sq(a) * bisectorLengthSq + length_sq(b) * bisectorLengthSq);
        if (areaSq > (P.minArea * P.minArea)) {
          V2 p[5] = {pos + a, at, pos + b, pos, pos + a};
          o.poly(p, 5);
        }
      }
    } else if (lineJoin == 2) {
      const float areaSq = 0.25f * length_sq(a) * length_sq(b);
      if (areaSq > (P.minArea * P.minArea)) {
        V2 p[5] = {a + pos, b + pos, pos, a + pos, a + pos};
        o.poly(p, 4);
      }
    }
  }
}

// everything point j of a stroke polygon contributes, in the reference's order (:2013-2080)
PXD void stroke_point(const EmitArgs& A, int j, StrokeOut& o) {
  const int k = A.pointPrim[j];
  const Prim& q = A.prims[k];
  const DPath& P = A.paths[q.path];
  o.P = &P;
  const int s = A.ptOff[q.shapeBegin], e = A.ptOff[A.prims[q.shapeBegin].shapeEnd];
  const int n = e - s, i = j - s;
  const float hs = P.halfStroke;
  auto pt = [&](int idx) { const float2 v = A.points[s + idx]; return v2(v.x, v.y); };
  const V2 p0 = pt(0), pl = pt(n - 1);
  const bool open = !veq(p0, pl);
  if (i == 0) {
    if (open && P.lineCap == 2) {  // SquareCap (LineCap: Butt 0, Round 1, Square 2; paths.nim:10-12)
      const V2 tangent = normalize(pt(1) - p0);
      make_rect(o, p0 - tangent * hs, p0, hs);
    }
    return;
  }
  const V2 pos = pt(i), prevPos = pt(i - 1);
  make_rect(o, prevPos, pos, hs);
  if (i < n - 1) add_join(o, prevPos, pos, pt(i + 1), hs);
  if (i == n - 1) {
    if (!open) {
      add_join(o, pt(n - 2), pl, pt(1), hs);
    } else if (P.lineCap == 2) {
      const V2 tangent = normalize(pl - pt(n - 2));
      make_rect(o, pl + tangent * hs, pl, hs);
    }
  }
}

__global__ void __launch_bounds__(128) stroke_count_kernel(const EmitArgs A, int numPoints, int* __restrict__ cnt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= numPoints) return;
  StrokeOut o;
  o.segs = nullptr; o.wind = nullptr; o.n = 0;
  stroke_point(A, j, o);
  cnt[j] = o.n;
}
__global__ void __launch_bounds__(128) stroke_emit_kernel(const EmitArgs A, int numPoints) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= numPoints) return;
  const int pi = A.prims[A.pointPrim[j]].path;
  const int pb = A.paths[pi].primBase;
  const int out = path_seg_begin(A, pi) + (A.seg2Off[j] - A.seg2Off[A.ptOff[pb]]);
  StrokeOut o;
  o.segs = A.segs + out; o.wind = A.wind + out; o.n = 0;
  stroke_point(A, j, o);
}

// per path: first segment, and computeBounds (:1098-1117) over its segments — one warp per path
__global__ void __launch_bounds__(128) bounds_kernel(const EmitArgs A, int numPaths, int fillSegs, const int* __restrict__ strokeSegs,
                                                     int* __restrict__ segBegin, float* __restrict__ bounds) {
  const int totalSegs = fillSegs + (strokeSegs ? *strokeSegs : 0);  // the stroke total is still on the device
  // one block per path (a path of the tiger has up to ten thousand segments: a single warp took 88 us over them)
  __shared__ float red[4][4];
  __shared__ int rnan[4];
  const int pi = blockIdx.x, lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  if (pi == numPaths) {
    if (threadIdx.x == 0) segBegin[pi] = totalSegs;
    return;
  }
  const int b = path_seg_begin(A, pi), e = pi + 1 < numPaths ? path_seg_begin(A, pi + 1) : totalSegs;
  float xMin = INFINITY, xMax = -INFINITY, yMin = INFINITY, yMax = -INFINITY;
  bool nan = false;
  for (int i = b + (int)threadIdx.x; i < e; i += 128) {
    const float4 s = A.segs[i];
    nan = nan || s.x != s.x || s.y != s.y || s.z != s.z || s.w != s.w;
    xMin = fminf(xMin, fminf(s.x, s.z));
    xMax = fmaxf(xMax, fmaxf(s.x, s.z));
    yMin = fminf(yMin, s.y);  // at.y < to.y for every segment
    yMax = fmaxf(yMax, s.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xMin = fminf(xMin, __shfl_xor_sync(0xffffffffu, xMin, o));
    xMax = fmaxf(xMax, __shfl_xor_sync(0xffffffffu, xMax, o));
    yMin = fminf(yMin, __shfl_xor_sync(0xffffffffu, yMin, o));
    yMax = fmaxf(yMax, __shfl_xor_sync(0xffffffffu, yMax, o));
  }
  nan = __any_sync(0xffffffffu, nan);
  if (lane == 0) {
    red[wi][0] = xMin; red[wi][1] = xMax; red[wi][2] = yMin; red[wi][3] = yMax;
    rnan[wi] = nan ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 4; k++) {
      xMin = fminf(xMin, red[k][0]); xMax = fmaxf(xMax, red[k][1]); yMin = fminf(yMin, red[k][2]); yMax = fmaxf(yMax, red[k][3]);
      nan = nan || rnan[k] != 0;
    }
    segBegin[pi] = b;
    bounds[5 * pi + 0] = xMin; bounds[5 * pi + 1] = xMax; bounds[5 * pi + 2] = yMin; bounds[5 * pi + 3] = yMax;
    bounds[5 * pi + 4] = nan ? 1.0f : 0.0f;
  }
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of an int array in place (data[n] receives the total): chunk sums, one block scans them, apply
// ---------------------------------------------------------------------------------------------
constexpr int kChunk = 2048;  // 256 threads x 8
__global__ void __launch_bounds__(256) scan_sum_kernel(const int* __restrict__ data, int n, int* __restrict__ chunkSum) {
  __shared__ int ws[8];
  const int base = blockIdx.x * kChunk;
  int s = 0;
  for (int i = base + threadIdx.x; i < min(base + kChunk, n); i += 256) s += data[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) s += ws[k];
    chunkSum[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(1024) scan_chunks_kernel(int* __restrict__ chunkSum, int numChunks) {  // one block; exclusive, total at [numChunks]
  __shared__ int carry;
  __shared__ int buf[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < numChunks; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int v = i < numChunks ? chunkSum[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < numChunks) chunkSum[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) chunkSum[numChunks] = carry;
}
__global__ void __launch_bounds__(256) scan_apply_kernel(int* __restrict__ data, int n, const int* __restrict__ chunkSum, int numChunks) {
  __shared__ int ws[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * kChunk + tid * 8;
  int v[8], s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    v[k] = i0 + k < n ? data[i0 + k] : 0;
    s += v[k];
  }
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  int run = chunkSum[blockIdx.x] + inc - s;
  for (int k = 0; k < warp; k++) run += ws[k];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (i0 + k < n) data[i0 + k] = run;
    run += v[k];
  }
  if (blockIdx.x == 0 && tid == 0) data[n] = chunkSum[numChunks];
}

// arrays up to 64 K entries (the tiger's 3280 primitives, 60 000 polygon points): one block, one launch
__global__ void __launch_bounds__(1024) scan_small_kernel(int* __restrict__ data, int n) {
  __shared__ int sums[1024];
  const int tid = threadIdx.x, per = (n + 1023) / 1024;
  const int b = min(tid * per, n), e = min(b + per, n);
  int s = 0;
  for (int i = b; i < e; i++) s += data[i];
  sums[tid] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int t = tid >= o ? sums[tid - o] : 0;
    __syncthreads();
    sums[tid] += t;
    __syncthreads();
  }
  int run = sums[tid] - s;
  for (int i = b; i < e; i++) {
    const int c = data[i];
    data[i] = run;
    run += c;
  }
  if (tid == 1023) data[n] = sums[1023];
}

static int exclusive_scan(int* data, int n, cudaStream_t st) {
  if (n <= 65536) {
    scan_small_kernel<<<1, 1024, 0, st>>>(data, n);
    PX_LAUNCHED();
    return 0;
  }
  const int numChunks = std::max(1, (n + kChunk - 1) / kChunk);
  int* chunkScratch = nullptr;
  PX_CUDA(cudaMallocAsync((void**)&chunkScratch, ((size_t)numChunks + 1) * 4, st));
  scan_sum_kernel<<<numChunks, 256, 0, st>>>(data, n, chunkScratch);
  PX_LAUNCHED();
  scan_chunks_kernel<<<1, 1024, 0, st>>>(chunkScratch, numChunks);
  PX_LAUNCHED();
  scan_apply_kernel<<<numChunks, 256, 0, st>>>(data, n, chunkScratch, numChunks);
  PX_LAUNCHED();
  PX_CUDA(cudaFreeAsync(chunkScratch, st));
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
void free_flattened(FlattenedPaths& F) {
  Runtime& r = rt();
  if (F.segs) cudaFreeAsync(F.segs, r.stream);
  if (F.wind) cudaFreeAsync(F.wind, r.stream);
  F.segs = nullptr;
  F.wind = nullptr;
}

int flatten_paths(int numPaths, const pixie_path_desc* descs, const float* commands, int64_t numCommandFloats, const float* rawXyxy,
                  const int16_t* rawWinding, int64_t numRaw, FlattenedPaths& F) {
  Runtime& r = rt();
  F = FlattenedPaths();
  if (numPaths < 0 || numCommandFloats < 0 || numRaw < 0 || numCommandFloats > 0x3fffffffll || numRaw > 0x3fffffffll)
    return fail_pixie("flatten: invalid sizes");
  F.segBegin.assign((size_t)numPaths + 1, 0);
  F.bounds.assign((size_t)numPaths * 5, 0.0f);
  if (numPaths == 0) return 0;
  std::vector<DPath> paths((size_t)numPaths);
  std::vector<int> rawBegin((size_t)numPaths, 0), rawCount((size_t)numPaths, 0);
  int64_t prims = 0;
  for (int k = 0; k < numPaths; k++) {
    const pixie_path_desc& d = descs[k];
    DPath& P = paths[(size_t)k];
    memset(&P, 0, sizeof(P));
    if (d.kind < 0 || d.kind > 2) return fail_pixie("flatten: invalid path kind");
    if (d.begin < 0 || d.end < d.begin || d.end > (d.kind == 2 ? numRaw : numCommandFloats)) return fail_pixie("flatten: path range out of bounds");
    if (d.kind == 1 && (d.line_cap == 1 || d.line_join == 1)) return fail_pixie("flatten: round caps / joins are flattened on the host (kind 2)");
    if (d.kind == 1 && (d.line_cap < 0 || d.line_cap > 2 || d.line_join < 0 || d.line_join > 2)) return fail_pixie("flatten: invalid cap / join");
    P.cmdBegin = d.begin; P.cmdEnd = d.end; P.kind = d.kind;
    P.primBase = (int)prims;
    // every command yields at most one primitive, plus the closing line of the last shape; Move also closes the shape
    // before it, which the count below covers because a Move itself draws nothing
    P.primCap = d.kind == 2 ? 1 : d.num_commands + 1;
    if (d.num_commands < 0) return fail_pixie("flatten: negative command count");
    prims += P.primCap;
    P.lineCap = d.line_cap; P.lineJoin = d.line_join;
    P.halfStroke = d.stroke_width / 2;                     // (:1935)
    P.miterAngleLimit = asinf(1 / d.miter_limit) * 2;      // (:1937)
    // transform.pixelScale (paths.nim:61-66)
    const float psa = sqrtf(d.transform[0] * d.transform[0] + d.transform[1] * d.transform[1]);
    const float psb = sqrtf(d.transform[3] * d.transform[3] + d.transform[4] * d.transform[4]);
    const float ps = psa > psb ? psa : psb;
    P.errorMarginSq = powf(0.2f / ps, 2.0f);               // pixelErrorMargin = 0.2 (:45, :667)
    P.minArea = 0.2f / ps;                                 // (:1961)
    memcpy(P.m, d.transform, sizeof(P.m));
    static const float ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    P.identity = memcmp(P.m, ident, sizeof(ident)) == 0 ? 1 : 0;
    if (!P.identity) {  // -0.0 entries compare equal to 0.0 in the reference's `!=` (:1093)
      bool same = true;
      for (int i = 0; i < 9; i++) same = same && P.m[i] == ident[i];
      P.identity = same ? 1 : 0;
    }
    if (d.kind == 2) {
      rawBegin[(size_t)k] = d.begin;
      rawCount[(size_t)k] = d.end - d.begin;
    }
    if (prims > 0x3fffffffll) return fail_pixie("flatten: too many commands");
  }
  const int numPrims = (int)prims;

  // one staging copy: path table, raw tables, commands, raw segments
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t off = 0;
  const size_t oPaths = off;   off = al(off + paths.size() * sizeof(DPath));
  const size_t oRawB = off;    off = al(off + rawBegin.size() * 4);
  const size_t oRawC = off;    off = al(off + rawCount.size() * 4);
  const size_t oCmds = off;    off = al(off + (size_t)numCommandFloats * 4);
  const size_t oRawS = off;    off = al(off + (size_t)numRaw * 16);
  const size_t oRawW = off;    off = al(off + (size_t)numRaw * 2);
  const size_t inBytes = off;
  const size_t oPrims = off;   off = al(off + (size_t)numPrims * sizeof(Prim));
  const size_t oCntPts = off;  off = al(off + ((size_t)numPrims + 1) * 4);
  const size_t oCntSeg = off;  off = al(off + ((size_t)numPrims + 1) * 4);
  const size_t oFirst = off;   off = al(off + (size_t)numPrims);
  const size_t oErr = off;     off = al(off + 16);
  const size_t oSegBegin = off; off = al(off + ((size_t)numPaths + 1) * 4);
  const size_t oBounds = off;  off = al(off + (size_t)numPaths * 5 * 4);
  const size_t totalA = off;
  uint8_t* blk = nullptr;
  PX_CUDA(cudaMallocAsync((void**)&blk, totalA, r.stream));
  struct Guard {  // temporaries go back to the pool on every exit path
    std::vector<void*> p;
    ~Guard() { for (void* q : p) if (q) cudaFreeAsync(q, rt().stream); }
  } guard;
  guard.p.push_back(blk);
  {
    void* pin;
    if (int rc = staging_acquire(inBytes, &pin)) return rc;
    uint8_t* st = (uint8_t*)pin;
    memcpy(st + oPaths, paths.data(), paths.size() * sizeof(DPath));
    memcpy(st + oRawB, rawBegin.data(), rawBegin.size() * 4);
    memcpy(st + oRawC, rawCount.data(), rawCount.size() * 4);
    if (numCommandFloats) memcpy(st + oCmds, commands, (size_t)numCommandFloats * 4);
    if (numRaw) {
      memcpy(st + oRawS, rawXyxy, (size_t)numRaw * 16);
      memcpy(st + oRawW, rawWinding, (size_t)numRaw * 2);
    }
    PX_CUDA(cudaMemcpyAsync(blk, st, inBytes, cudaMemcpyHostToDevice, r.stream));
    if (int rc = staging_release()) return rc;
  }
  F.h2dBytes = inBytes;
  const DPath* dPaths = (const DPath*)(blk + oPaths);
  Prim* dPrims = (Prim*)(blk + oPrims);
  int* cntPts = (int*)(blk + oCntPts);
  int* cntSeg = (int*)(blk + oCntSeg);
  uint8_t* first = blk + oFirst;
  int* err = (int*)(blk + oErr);
  PX_CUDA(cudaMemsetAsync(blk + oFirst, 0, totalA - oFirst, r.stream));  // first flags, error word, outputs

  resolve_kernel<<<(numPaths + kResolveWarps - 1) / kResolveWarps, kResolveWarps * 32, 0, r.stream>>>(dPaths, numPaths, (const float*)(blk + oCmds), dPrims, err);
  PX_LAUNCHED();
  count_kernel_f<<<(numPrims + 127) / 128, 128, 0, r.stream>>>(dPaths, dPrims, numPrims, (const int*)(blk + oRawC), cntPts, cntSeg, err);
  PX_LAUNCHED();
  shape_first_kernel<<<(numPaths + 127) / 128, 128, 0, r.stream>>>(dPaths, numPaths, dPrims, cntPts, first);
  PX_LAUNCHED();
  if (int rc = exclusive_scan(cntPts, numPrims, r.stream)) return rc;
  if (int rc = exclusive_scan(cntSeg, numPrims, r.stream)) return rc;
  int head[3] = {0, 0, 0};  // points, fill + raw segments, error
  PX_CUDA(cudaMemcpyAsync(&head[0], cntPts + numPrims, 4, cudaMemcpyDeviceToHost, r.stream));
  PX_CUDA(cudaMemcpyAsync(&head[1], cntSeg + numPrims, 4, cudaMemcpyDeviceToHost, r.stream));
  PX_CUDA(cudaMemcpyAsync(&head[2], err, 4, cudaMemcpyDeviceToHost, r.stream));
  PX_CUDA(cudaStreamSynchronize(r.stream));
  if (head[2] == 1) return fail_pixie("Unable to discretize curve");  // paths.nim:707, :744
  if (head[2] == 2) return fail_pixie("flatten: arcs are flattened on the host (kind 2)");
  if (head[2] == 3) return fail_pixie("Invalid path command");
  if (head[2] == 4) return fail_pixie("flatten: num_commands does not match the command stream");
  const int numPoints = head[0], fillSegs = head[1];
  // a polygon point contributes at most a cap / closing join (4) + a rectangle (4) + a join (4) edges
  const int64_t segCap = (int64_t)fillSegs + 12ll * numPoints;
  if (segCap > 0x3fffffffll) return fail_pixie("flatten: too many segments");
  float2* points = nullptr;
  int *pointPrim = nullptr, *cnt2 = nullptr;
  PX_CUDA(cudaMallocAsync((void**)&F.segs, std::max<size_t>(16, (size_t)segCap * 16), r.stream));
  PX_CUDA(cudaMallocAsync((void**)&F.wind, std::max<size_t>(16, (size_t)segCap * 2), r.stream));
  if (numPoints > 0) {
    PX_CUDA(cudaMallocAsync((void**)&points, (size_t)numPoints * 8, r.stream));
    guard.p.push_back(points);
    PX_CUDA(cudaMallocAsync((void**)&pointPrim, (size_t)numPoints * 4, r.stream));
    guard.p.push_back(pointPrim);
    PX_CUDA(cudaMallocAsync((void**)&cnt2, ((size_t)numPoints + 1) * 4, r.stream));
    guard.p.push_back(cnt2);
  }
  EmitArgs A;
  A.paths = dPaths; A.prims = dPrims; A.numPrims = numPrims; A.ptOff = cntPts; A.segOff = cntSeg; A.seg2Off = nullptr;
  A.first = first; A.points = points; A.pointPrim = pointPrim; A.segs = F.segs; A.wind = F.wind;
  A.rawSegs = (const float4*)(blk + oRawS); A.rawWind = (const int16_t*)(blk + oRawW); A.rawBegin = (const int*)(blk + oRawB);
  int totalSegs = fillSegs;
  if (numPoints > 0) {
    emit_kernel_f<true><<<(numPrims + 127) / 128, 128, 0, r.stream>>>(A);
    PX_LAUNCHED();
    stroke_count_kernel<<<(numPoints + 127) / 128, 128, 0, r.stream>>>(A, numPoints, cnt2);
    PX_LAUNCHED();
    if (int rc = exclusive_scan(cnt2, numPoints, r.stream)) return rc;
    A.seg2Off = cnt2;
  }
  emit_kernel_f<false><<<(numPrims + 127) / 128, 128, 0, r.stream>>>(A);
  PX_LAUNCHED();
  if (numRaw > 0) {
    raw_copy_kernel<<<numPaths, 256, 0, r.stream>>>(A, numPaths);
    PX_LAUNCHED();
  }
  if (numPoints > 0) {
    stroke_emit_kernel<<<(numPoints + 127) / 128, 128, 0, r.stream>>>(A, numPoints);
    PX_LAUNCHED();
  }
  int* dSegBegin = (int*)(blk + oSegBegin);
  float* dBounds = (float*)(blk + oBounds);
  bounds_kernel<<<numPaths + 1, 128, 0, r.stream>>>(A, numPaths, fillSegs, numPoints > 0 ? cnt2 + numPoints : nullptr, dSegBegin, dBounds);
  PX_LAUNCHED();
  // one readback for the path table: first segment of every path (the last entry = the total) and the bounds
  PX_CUDA(cudaMemcpyAsync(F.segBegin.data(), dSegBegin, ((size_t)numPaths + 1) * 4, cudaMemcpyDeviceToHost, r.stream));
  PX_CUDA(cudaMemcpyAsync(F.bounds.data(), dBounds, (size_t)numPaths * 5 * 4, cudaMemcpyDeviceToHost, r.stream));
  PX_CUDA(cudaStreamSynchronize(r.stream));
  totalSegs = F.segBegin[(size_t)numPaths];
  F.numSegs = totalSegs;
  F.numPoints = numPoints;
  F.numPrims = numPrims;
  return 0;
}

}  // namespace pixie

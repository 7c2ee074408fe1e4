// libpixie_host: C++ mirror of the Nim host-side producers above the raster hot path.
//
// Follows (behaviour, not code) treeform/pixie src/pixie/paths.nim:
//   parsePath :119-262, builders :339-652, commandsToShapes :654-1057,
//   shapesToSegments :1059-1090, strokeShapes :1922-2082; internal.nim gaussianKernel :17-34.
// Third-party math (vmath / bumpy, not vendored in the reference) follows SURVEY.md Appendix A.
//
// Numerics contract: IEEE float32, one rounding per operation, no FMA contraction
// (build with -ffp-contract=off).  Mixed float32 x float64-constant expressions are evaluated
// in float64 and narrowed where the Nim source does so (PI, TAU, splineCircleK).
#include "pixie_host.h"

#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

struct PixieError {
  std::string msg;
};

struct V2 {
  float x = 0.f, y = 0.f;
};
inline V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
inline V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
inline V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
inline V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
inline V2 operator/(V2 a, float s) { return v2(a.x / s, a.y / s); }
inline bool operator==(V2 a, V2 b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(V2 a, V2 b) { return !(a == b); }
inline float lengthSq(V2 a) { return a.x * a.x + a.y * a.y; }
inline float length(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
inline V2 normalize(V2 a) { return a / length(a); }

typedef std::vector<V2> Polygon;

const double kPI = 3.141592653589793238462643383279502884;
const double kTAU = 2.0 * kPI;
const float kEpsilon = (float)(0.0001 * kPI);  // paths.nim:44
const float kPixelErrorMargin = 0.2f;           // paths.nim:45
const double kSplineCircleK = 4.0 * (-1.0 + sqrt(2.0)) / 3;  // paths.nim:529

// vmath Mat3, column-major: m[c*3+r]
struct Mat3 {
  float m[9];
};
inline Mat3 identity() {
  Mat3 r = {{1, 0, 0, 0, 1, 0, 0, 0, 1}};
  return r;
}
inline bool isIdentity(const Mat3& a) {
  Mat3 i = identity();
  for (int k = 0; k < 9; k++)
    if (a.m[k] != i.m[k]) return false;
  return true;
}
inline V2 mul(const Mat3& a, V2 b) {
  return v2(a.m[0] * b.x + a.m[3] * b.y + a.m[6], a.m[1] * b.x + a.m[4] * b.y + a.m[7]);
}
inline float pixelScale(const Mat3& t) {  // paths.nim:61-66
  float a = length(v2(t.m[0], t.m[1]));
  float b = length(v2(t.m[3], t.m[4]));
  return a > b ? a : b;
}

enum Kind {
  Close = 0, Move, Line, HLine, VLine, Cubic, SCubic, Quad, TQuad, Arc,
  RMove, RLine, RHLine, RVLine, RCubic, RSCubic, RQuad, RTQuad, RArc
};

int parameterCount(int kind) {  // paths.nim:73-81
  switch (kind) {
    case Close: return 0;
    case Move: case Line: case RMove: case RLine: case TQuad: case RTQuad: return 2;
    case HLine: case VLine: case RHLine: case RVLine: return 1;
    case Cubic: case RCubic: return 6;
    case SCubic: case RSCubic: case Quad: case RQuad: return 4;
    case Arc: case RArc: return 7;
  }
  return 0;
}

}  // namespace

struct pixie_path {
  std::vector<float> commands;
  V2 start, at;
};

struct pixie_segments {
  std::vector<float> xyxy;
  std::vector<int16_t> winding;
};

namespace {

// ---------------------------------------------------------------- parsePath (paths.nim:119-262)
void parsePathInto(const std::string& path, pixie_path* result) {
  if (path.empty()) return;
  size_t p = 0, numberStart = 0;
  bool armed = false, hitDecimal = false;
  int kind = Close;
  std::vector<float> numbers;

  auto finishNumber = [&]() {
    if (numberStart > 0) {
      std::string tok = path.substr(numberStart, p - numberStart);
      char* end = nullptr;
      double v = strtod(tok.c_str(), &end);
      if (end == tok.c_str() || *end != '\0')
        throw PixieError{"Invalid path, parsing parameter failed"};
      numbers.push_back((float)v);
    }
    numberStart = 0;
    hitDecimal = false;
  };
  auto finishCommand = [&]() {
    finishNumber();
    if (armed) {
      int paramCount = parameterCount(kind);
      if (paramCount == 0) {
        if (!numbers.empty()) throw PixieError{"Invalid path, unexpected parameters"};
        result->commands.push_back((float)kind);
      } else {
        if (numbers.size() % paramCount != 0)
          throw PixieError{"Invalid path, wrong number of parameters"};
        size_t batches = numbers.size() / paramCount;
        for (size_t batch = 0; batch < batches; batch++) {
          if (batch > 0) {
            if (kind == Move) kind = Line;
            else if (kind == RMove) kind = RLine;
          }
          result->commands.push_back((float)kind);
          for (int i = 0; i < paramCount; i++)
            result->commands.push_back(numbers[batch * paramCount + i]);
        }
        numbers.clear();
      }
    }
    armed = true;
  };
  auto expectsArcFlag = [&]() {
    if (!(kind == Arc || kind == RArc)) return false;
    size_t m = numbers.size() % 7;
    return m == 3 || m == 4;
  };

  // NB: the reference uses numberStart == 0 as "no number in progress", so a number that
  // starts at index 0 of the string is not tracked; kept (paths.nim:133,246,257).
  while (p < path.size()) {
    char c = path[p];
    int newKind = -1;
    switch (c) {
      case 'm': newKind = RMove; break;
      case 'l': newKind = RLine; break;
      case 'h': newKind = RHLine; break;
      case 'v': newKind = RVLine; break;
      case 'c': newKind = RCubic; break;
      case 's': newKind = RSCubic; break;
      case 'q': newKind = RQuad; break;
      case 't': newKind = RTQuad; break;
      case 'a': newKind = RArc; break;
      case 'z': newKind = Close; break;
      case 'M': newKind = Move; break;
      case 'L': newKind = Line; break;
      case 'H': newKind = HLine; break;
      case 'V': newKind = VLine; break;
      case 'C': newKind = Cubic; break;
      case 'S': newKind = SCubic; break;
      case 'Q': newKind = Quad; break;
      case 'T': newKind = TQuad; break;
      case 'A': newKind = Arc; break;
      case 'Z': newKind = Close; break;
      default: break;
    }
    if (newKind >= 0) {
      finishCommand();
      kind = newKind;
    } else if (c == '-' || c == '+') {
      if (numberStart > 0 && (path[p - 1] == 'e' || path[p - 1] == 'E')) {
        // exponent sign
      } else {
        finishNumber();
        numberStart = p;
      }
    } else if (c == '.') {
      if (hitDecimal || expectsArcFlag()) finishNumber();
      hitDecimal = true;
      if (numberStart == 0) numberStart = p;
    } else if (c == ' ' || c == ',' || c == '\r' || c == '\n' || c == '\t') {
      finishNumber();
    } else {
      if (numberStart > 0 && expectsArcFlag()) finishNumber();
      if (p >= 1 && p - 1 == numberStart && path[p - 1] == '0') finishNumber();
      if (numberStart == 0) numberStart = p;
    }
    p++;
  }
  finishCommand();
}

// ---------------------------------------------------------------- builders (paths.nim:339-652)
void moveTo(pixie_path* p, float x, float y) {
  p->commands.push_back((float)Move);
  p->commands.push_back(x);
  p->commands.push_back(y);
  p->start = v2(x, y);
  p->at = p->start;
}
void lineTo(pixie_path* p, float x, float y) {
  p->commands.push_back((float)Line);
  p->commands.push_back(x);
  p->commands.push_back(y);
  p->at = v2(x, y);
}
void bezierCurveTo(pixie_path* p, float x1, float y1, float x2, float y2, float x3, float y3) {
  float c[7] = {(float)Cubic, x1, y1, x2, y2, x3, y3};
  p->commands.insert(p->commands.end(), c, c + 7);
  p->at = v2(x3, y3);
}
void quadraticCurveTo(pixie_path* p, float x1, float y1, float x2, float y2) {
  float c[5] = {(float)Quad, x1, y1, x2, y2};
  p->commands.insert(p->commands.end(), c, c + 5);
  p->at = v2(x2, y2);
}
void ellipticalArcTo(pixie_path* p, float rx, float ry, float rot, bool large, bool sweep, float x, float y) {
  float c[8] = {(float)Arc, rx, ry, rot, large ? 1.f : 0.f, sweep ? 1.f : 0.f, x, y};
  p->commands.insert(p->commands.end(), c, c + 8);
  p->at = v2(x, y);
}
void closePath(pixie_path* p) {
  p->commands.push_back((float)Close);
  p->at = p->start;
}
void rect(pixie_path* p, float x, float y, float w, float h, bool clockwise) {
  if (clockwise) {
    moveTo(p, x, y);
    lineTo(p, x + w, y);
    lineTo(p, x + w, y + h);
    lineTo(p, x, y + h);
    closePath(p);
  } else {
    moveTo(p, x, y);
    lineTo(p, x, y + h);
    lineTo(p, x + w, y + h);
    lineTo(p, x + w, y);
    closePath(p);
  }
}
void ellipse(pixie_path* p, float cx, float cy, float rx, float ry) {
  // magicX/magicY are float64 (float64 const * float32), sums evaluated in float64 then
  // narrowed at the float32 call boundary (paths.nim:608-619).
  double magicX = kSplineCircleK * (double)rx;
  double magicY = kSplineCircleK * (double)ry;
  moveTo(p, cx + rx, cy);
  bezierCurveTo(p, cx + rx, (float)((double)cy + magicY), (float)((double)cx + magicX), cy + ry, cx, cy + ry);
  bezierCurveTo(p, (float)((double)cx - magicX), cy + ry, cx - rx, (float)((double)cy + magicY), cx - rx, cy);
  bezierCurveTo(p, cx - rx, (float)((double)cy - magicY), (float)((double)cx - magicX), cy - ry, cx, cy - ry);
  bezierCurveTo(p, (float)((double)cx + magicX), cy - ry, cx + rx, (float)((double)cy - magicY), cx + rx, cy);
  closePath(p);
}
void roundedRect(pixie_path* p, float x, float y, float w, float h, float nw, float ne, float se,
                 float sw, bool clockwise) {
  float maxRadius = fminf(w / 2, h / 2);
  nw = fmaxf(0, fminf(nw, maxRadius));
  ne = fmaxf(0, fminf(ne, maxRadius));
  se = fmaxf(0, fminf(se, maxRadius));
  sw = fmaxf(0, fminf(sw, maxRadius));
  if (nw == 0 && ne == 0 && se == 0 && sw == 0) {
    rect(p, x, y, w, h, clockwise);
    return;
  }
  const double s = kSplineCircleK;
  V2 t1 = v2(x + nw, y), t2 = v2(x + w - ne, y), r1 = v2(x + w, y + ne), r2 = v2(x + w, y + h - se);
  V2 b1 = v2(x + w - se, y + h), b2 = v2(x + sw, y + h), l1 = v2(x, y + h - sw), l2 = v2(x, y + nw);
  V2 t1h = t1 + v2((float)(-(double)nw * s), 0);
  V2 t2h = t2 + v2((float)(+(double)ne * s), 0);
  V2 r1h = r1 + v2(0, (float)(-(double)ne * s));
  V2 r2h = r2 + v2(0, (float)(+(double)se * s));
  V2 b1h = b1 + v2((float)(+(double)se * s), 0);
  V2 b2h = b2 + v2((float)(-(double)sw * s), 0);
  V2 l1h = l1 + v2(0, (float)(+(double)sw * s));
  V2 l2h = l2 + v2(0, (float)(-(double)nw * s));
  auto bez = [&](V2 a, V2 b, V2 c) { bezierCurveTo(p, a.x, a.y, b.x, b.y, c.x, c.y); };
  if (clockwise) {
    moveTo(p, t1.x, t1.y);
    lineTo(p, t2.x, t2.y);
    bez(t2h, r1h, r1);
    lineTo(p, r2.x, r2.y);
    bez(r2h, b1h, b1);
    lineTo(p, b2.x, b2.y);
    bez(b2h, l1h, l1);
    lineTo(p, l2.x, l2.y);
    bez(l2h, t1h, t1);
  } else {
    moveTo(p, t1.x, t1.y);
    bez(t1h, l2h, l2);
    lineTo(p, l1.x, l1.y);
    bez(l1h, b2h, b2);
    lineTo(p, b1.x, b1.y);
    bez(b1h, r2h, r2);
    lineTo(p, r1.x, r1.y);
    bez(r1h, t2h, t2);
    lineTo(p, t1.x, t1.y);
  }
  closePath(p);
}
void polygon(pixie_path* p, float x, float y, float size, int sides) {  // paths.nim:633-646
  if (sides <= 2) throw PixieError{"Invalid polygon sides value"};
  moveTo(p, (float)((double)x + (double)size * sin(0.0)), (float)((double)y - (double)size * cos(0.0)));
  for (int side = 1; side <= sides - 1; side++) {
    double ang = (double)(float)side * 2.0 * kPI / (double)(float)sides;
    lineTo(p, (float)((double)x + (double)size * sin(ang)), (float)((double)y - (double)size * cos(ang)));
  }
  closePath(p);
}
void arcBuilder(pixie_path* p, float x, float y, float r, float a0, float a1, bool ccw) {  // :416-453
  if (r == 0) return;
  if (r < 0) throw PixieError{"Invalid arc, negative radius: " + std::to_string(r)};
  float dx = r * cosf(a0), dy = r * sinf(a0);
  float x0 = x + dx, y0 = y + dy;
  bool cw = !ccw;
  if (p->commands.empty()) moveTo(p, x0, y0);
  else if (fabsf(p->at.x - x0) > kEpsilon || fabsf(p->at.y - y0) > kEpsilon) lineTo(p, x0, y0);
  float angle = ccw ? a0 - a1 : a1 - a0;
  if (angle < 0) angle = (float)(fmod((double)angle, kTAU) + kTAU);
  if ((double)angle > kTAU - (double)kEpsilon) {
    ellipticalArcTo(p, r, r, 0, true, cw, x - dx, y - dy);
    p->at.x = x0;
    p->at.y = y0;
    ellipticalArcTo(p, r, r, 0, true, cw, p->at.x, p->at.y);
  } else if (angle > kEpsilon) {
    p->at.x = x + r * cosf(a1);
    p->at.y = y + r * sinf(a1);
    ellipticalArcTo(p, r, r, 0, (double)angle >= kPI, cw, p->at.x, p->at.y);
  }
}
void arcToBuilder(pixie_path* p, float x1, float y1, float x2, float y2, float r) {  // :461-500
  if (r < 0) throw PixieError{"Invalid arc, negative radius: " + std::to_string(r)};
  float x0 = p->at.x, y0 = p->at.y;
  float x21 = x2 - x1, y21 = y2 - y1, x01 = x0 - x1, y01 = y0 - y1;
  float l01_2 = x01 * x01 + y01 * y01;
  if (p->commands.empty()) {
    moveTo(p, x0, y0);
  } else if (!(l01_2 > kEpsilon)) {
  } else if (!(fabsf(y01 * x21 - y21 * x01) > kEpsilon) || r == 0) {
    lineTo(p, x1, y1);
  } else {
    float x20 = x2 - x0, y20 = y2 - y0;
    float l21_2 = x21 * x21 + y21 * y21, l20_2 = x20 * x20 + y20 * y20;
    float l21 = sqrtf(l21_2), l01 = sqrtf(l01_2);
    double inner = (kPI - (double)acosf((l21_2 + l01_2 - l20_2) / (2 * l21 * l01))) / 2;
    float l = (float)((double)r * tan(inner));
    float t01 = l / l01, t21 = l / l21;
    if (fabsf(t01 - 1) > kEpsilon) lineTo(p, x1 + t01 * x01, y1 + t01 * y01);
    p->at.x = x1 + t21 * x21;
    p->at.y = y1 + t21 * y21;
    ellipticalArcTo(p, r, r, 0, false, y01 * x20 > x01 * y20, p->at.x, p->at.y);
  }
}

// ---------------------------------------------------------------- commandsToShapes (:654-1057)
struct Flattener {
  float errorMarginSq;

  static void addSegment(Polygon& shape, V2 at, V2 to) {
    V2 d = at - to;
    if (d != v2(0, 0)) {
      if (shape.empty() || shape.back() != at) shape.push_back(at);
      shape.push_back(to);
    }
  }

  static V2 cubicPoint(V2 at, V2 c1, V2 c2, V2 to, float t) {
    float t2 = t * t, t3 = t2 * t;
    return at * (-t3 + 3 * t2 - 3 * t + 1) + c1 * (3 * t3 - 6 * t2 + 3 * t) + c2 * (-3 * t3 + 3 * t2) + to * (t3);
  }
  static V2 cubicDeriv(V2 at, V2 c1, V2 c2, V2 to, float t) {
    float t2 = t * t;
    return at * (-3 * t2 + 6 * t - 3) + c1 * (9 * t2 - 12 * t + 3) + c2 * (-9 * t2 + 6 * t) + to * (3 * t2);
  }
  void addCubic(Polygon& shape, V2 at, V2 c1, V2 c2, V2 to) const {
    float t = 0, step = 1;
    V2 prev = at;
    V2 next = cubicPoint(at, c1, c2, to, t + step);
    V2 halfway = cubicPoint(at, c1, c2, to, t + step / 2);
    while (true) {
      if (step <= FLT_EPSILON) throw PixieError{"Unable to discretize cubic"};
      V2 midpoint = (prev + next) / 2;
      V2 lineTangent = midpoint - prev;
      V2 curveTangent = cubicDeriv(at, c1, c2, to, t + step / 2);
      V2 curveTangentScaled = normalize(curveTangent) * length(lineTangent);
      float error = lengthSq(midpoint - halfway);
      float errorTangent = lengthSq(lineTangent - curveTangentScaled);
      if (error + errorTangent > errorMarginSq) {
        next = halfway;
        halfway = cubicPoint(at, c1, c2, to, t + step / 4);
        step /= 2;
      } else {
        addSegment(shape, prev, next);
        t += step;
        if (t == 1) break;
        prev = next;
        step = fminf(step * 2, 1 - t);
        next = cubicPoint(at, c1, c2, to, t + step);
        halfway = cubicPoint(at, c1, c2, to, t + step / 2);
      }
    }
  }

  static V2 quadPoint(V2 at, V2 ctrl, V2 to, float t) {
    float t2 = t * t;
    return at * (t2 - 2 * t + 1) + ctrl * (-2 * t2 + 2 * t) + to * t2;
  }
  void addQuadratic(Polygon& shape, V2 at, V2 ctrl, V2 to) const {
    float t = 0, step = 1;
    V2 prev = at;
    V2 next = quadPoint(at, ctrl, to, t + step);
    V2 halfway = quadPoint(at, ctrl, to, t + step / 2);
    bool halfStepping = false;
    while (true) {
      if (step <= FLT_EPSILON) throw PixieError{"Unable to discretize quadratic"};
      V2 midpoint = (prev + next) / 2;
      float error = lengthSq(midpoint - halfway);
      if (error > errorMarginSq) {
        next = halfway;
        halfway = quadPoint(at, ctrl, to, t + step / 4);
        halfStepping = true;
        step /= 2;
      } else {
        addSegment(shape, prev, next);
        t += step;
        if (t == 1) break;
        prev = next;
        if (halfStepping) step = fminf(step, 1 - t);
        else step = fminf(step * 2, 1 - t);
        next = quadPoint(at, ctrl, to, t + step);
        halfway = quadPoint(at, ctrl, to, t + step / 2);
      }
    }
  }

  struct ArcParams {
    V2 radii;
    float rc, rs;  // rotMat = rotate(radians): applied as (c*x - s*y, s*x + c*y)
    V2 center;
    float theta, delta;
  };
  static float svgAngle(V2 u, V2 v) {
    float dot = u.x * v.x + u.y * v.y;
    float len = length(u) * length(v);
    float q = dot / len;
    float cl = q < -1.f ? -1.f : (q > 1.f ? 1.f : q);  // clamp(x,-1,1) = max(-1, min(1, x))
    float result = acosf(cl);
    if ((u.x * v.y - u.y * v.x) < 0) result = -result;
    return result;
  }
  static ArcParams endpointToCenter(V2 at, V2 radiiIn, float rotation, bool large, bool sweep, V2 to) {
    V2 radii = v2(fabsf(radiiIn.x), fabsf(radiiIn.y));
    V2 radiiSq = v2(radii.x * radii.x, radii.y * radii.y);
    float radians = (float)((double)(rotation / 180) * kPI);
    V2 d = v2((at.x - to.x) / 2.0f, (at.y - to.y) / 2.0f);
    float cr_ = cosf(radians), sr_ = sinf(radians);
    V2 p = v2(cr_ * d.x + sr_ * d.y, -sr_ * d.x + cr_ * d.y);
    V2 pSq = v2(p.x * p.x, p.y * p.y);
    float cr = pSq.x / radiiSq.x + pSq.y / radiiSq.y;
    if (cr > 1) {
      radii = radii * sqrtf(cr);
      radiiSq = v2(radii.x * radii.x, radii.y * radii.y);
    }
    float dq = radiiSq.x * pSq.y + radiiSq.y * pSq.x;
    float pq = (radiiSq.x * radiiSq.y - dq) / dq;
    float q = sqrtf(fmaxf(0.f, pq));  // max(0, pq): NaN pq -> 0 in Nim's max (a<b ? b : a)
    if (!(pq > 0.f) && !(pq <= 0.f)) q = sqrtf(0.f);
    if (large == sweep) q = -q;
    V2 cp = v2(q * radii.x * p.y / radii.y, -q * radii.y * p.x / radii.x);
    V2 center = v2(cr_ * cp.x - sr_ * cp.y + (at.x + to.x) / 2, sr_ * cp.x + cr_ * cp.y + (at.y + to.y) / 2);
    float theta = svgAngle(v2(1, 0), v2((p.x - cp.x) / radii.x, (p.y - cp.y) / radii.y));
    float delta = svgAngle(v2((p.x - cp.x) / radii.x, (p.y - cp.y) / radii.y),
                           v2((-p.x - cp.x) / radii.x, (-p.y - cp.y) / radii.y));
    delta = (float)fmod((double)delta, kPI * 2);
    if (sweep && delta < 0) delta += (float)(2 * kPI);
    else if (!sweep && delta > 0) delta -= (float)(2 * kPI);
    while ((double)delta > kPI * 2) delta -= (float)(kPI * 2);
    while ((double)delta < -kPI * 2) delta += (float)(kPI * 2);
    ArcParams a;
    a.radii = radii;
    a.rc = cosf(radians);
    a.rs = sinf(radians);
    a.center = center;
    a.theta = theta;
    a.delta = delta;
    return a;
  }
  static V2 arcPoint(const ArcParams& arc, float a) {
    V2 r = v2(cosf(a) * arc.radii.x, sinf(a) * arc.radii.y);
    // rotMat * r: m00*x + m10*y + m20 with m00=c, m10=-s, m20=0; m01=s, m11=c
    V2 rot = v2(arc.rc * r.x + (-arc.rs) * r.y + 0.f, arc.rs * r.x + arc.rc * r.y + 0.f);
    return rot + arc.center;
  }
  void addArc(Polygon& shape, V2 at, V2 radii, float rotation, bool large, bool sweep, V2 to) const {
    ArcParams arc = endpointToCenter(at, radii, rotation, large, sweep, to);
    float t = 0, step = 1;
    V2 prev = at;
    while (t != 1) {
      if (step <= FLT_EPSILON) throw PixieError{"Unable to discretize arc"};
      float aPrev = arc.theta + arc.delta * t;
      float a = arc.theta + arc.delta * (t + step);
      V2 next = arcPoint(arc, a);
      V2 halfway = arcPoint(arc, aPrev + (a - aPrev) / 2);
      V2 midpoint = (prev + next) / 2;
      float error = lengthSq(midpoint - halfway);
      if (error > errorMarginSq) {
        V2 quarterway = arcPoint(arc, aPrev + (a - aPrev) / 4);
        V2 midpoint2 = (prev + halfway) / 2;
        float halfwayError = lengthSq(midpoint2 - quarterway);
        if (halfwayError < errorMarginSq) {
          addSegment(shape, prev, halfway);
          prev = halfway;
          t += step / 2;
          step = fminf(step / 2, 1 - t);
        } else {
          step = step / 4;
        }
      } else {
        addSegment(shape, prev, next);
        prev = next;
        t += step;
        step = fminf(step * 2, 1 - t);
      }
    }
  }
};

bool isCubicKind(int k) { return k == Cubic || k == SCubic || k == RCubic || k == RSCubic; }
bool isQuadKind(int k) { return k == Quad || k == TQuad || k == RQuad || k == RTQuad; }

std::vector<Polygon> commandsToShapes(const std::vector<float>& cmds, bool closeSubpaths, float pixelScale) {
  std::vector<Polygon> result;
  V2 start, at;
  Polygon shape;
  int prevCommandKind = Move;
  V2 prevCtrl, prevCtrl2;
  Flattener fl;
  fl.errorMarginSq = powf(kPixelErrorMargin / pixelScale, 2.0f);

  size_t i = 0;
  while (i < cmds.size()) {
    int kind = (int)cmds[i];
    i++;
    const float* c = cmds.data() + i;
    switch (kind) {
      case Move:
        if (!shape.empty()) {
          if (closeSubpaths) Flattener::addSegment(shape, at, start);
          result.push_back(shape);
          shape.clear();
        }
        at.x = c[0];
        at.y = c[1];
        start = at;
        break;
      case Line: {
        V2 to = v2(c[0], c[1]);
        Flattener::addSegment(shape, at, to);
        at = to;
      } break;
      case HLine: {
        V2 to = v2(c[0], at.y);
        Flattener::addSegment(shape, at, to);
        at = to;
      } break;
      case VLine: {
        V2 to = v2(at.x, c[0]);
        Flattener::addSegment(shape, at, to);
        at = to;
      } break;
      case Cubic: {
        V2 c1 = v2(c[0], c[1]), c2 = v2(c[2], c[3]), to = v2(c[4], c[5]);
        fl.addCubic(shape, at, c1, c2, to);
        at = to;
        prevCtrl2 = c2;
      } break;
      case SCubic: {
        V2 c2 = v2(c[0], c[1]), to = v2(c[2], c[3]);
        if (isCubicKind(prevCommandKind)) {
          V2 c1 = at * 2 - prevCtrl2;
          fl.addCubic(shape, at, c1, c2, to);
        } else {
          fl.addCubic(shape, at, at, c2, to);
        }
        at = to;
        prevCtrl2 = c2;
      } break;
      case Quad: {
        V2 ctrl = v2(c[0], c[1]), to = v2(c[2], c[3]);
        fl.addQuadratic(shape, at, ctrl, to);
        at = to;
        prevCtrl = ctrl;
      } break;
      case TQuad: {
        V2 to = v2(c[0], c[1]);
        V2 ctrl = isQuadKind(prevCommandKind) ? at * 2 - prevCtrl : at;
        fl.addQuadratic(shape, at, ctrl, to);
        at = to;
        prevCtrl = ctrl;
      } break;
      case Arc: {
        V2 radii = v2(c[0], c[1]);
        float rotation = c[2];
        bool large = c[3] == 1, sweep = c[4] == 1;
        V2 to = v2(c[5], c[6]);
        fl.addArc(shape, at, radii, rotation, large, sweep, to);
        at = to;
      } break;
      case RMove:
        if (!shape.empty()) {
          result.push_back(shape);
          shape.clear();
        }
        at.x += c[0];
        at.y += c[1];
        start = at;
        break;
      case RLine: {
        V2 to = v2(at.x + c[0], at.y + c[1]);
        Flattener::addSegment(shape, at, to);
        at = to;
      } break;
      case RHLine: {
        V2 to = v2(at.x + c[0], at.y);
        Flattener::addSegment(shape, at, to);
        at = to;
      } break;
      case RVLine: {
        V2 to = v2(at.x, at.y + c[0]);
        Flattener::addSegment(shape, at, to);
        at = to;
      } break;
      case RCubic: {
        V2 c1 = v2(at.x + c[0], at.y + c[1]), c2 = v2(at.x + c[2], at.y + c[3]), to = v2(at.x + c[4], at.y + c[5]);
        fl.addCubic(shape, at, c1, c2, to);
        at = to;
        prevCtrl2 = c2;
      } break;
      case RSCubic: {
        V2 c2 = v2(at.x + c[0], at.y + c[1]), to = v2(at.x + c[2], at.y + c[3]);
        V2 c1 = isCubicKind(prevCommandKind) ? at * 2 - prevCtrl2 : at;
        fl.addCubic(shape, at, c1, c2, to);
        at = to;
        prevCtrl2 = c2;
      } break;
      case RQuad: {
        V2 ctrl = v2(at.x + c[0], at.y + c[1]), to = v2(at.x + c[2], at.y + c[3]);
        fl.addQuadratic(shape, at, ctrl, to);
        at = to;
        prevCtrl = ctrl;
      } break;
      case RTQuad: {
        V2 to = v2(at.x + c[0], at.y + c[1]);
        V2 ctrl = isQuadKind(prevCommandKind) ? at * 2 - prevCtrl : at;
        fl.addQuadratic(shape, at, ctrl, to);
        at = to;
        prevCtrl = ctrl;
      } break;
      case RArc: {
        V2 radii = v2(c[0], c[1]);
        float rotation = c[2];
        bool large = c[3] == 1, sweep = c[4] == 1;
        V2 to = v2(at.x + c[5], at.y + c[6]);
        fl.addArc(shape, at, radii, rotation, large, sweep, to);
        at = to;
      } break;
      case Close:
        if (at != start) {
          Flattener::addSegment(shape, at, start);
          at = start;
        }
        if (!shape.empty()) {
          result.push_back(shape);
          shape.clear();
        }
        break;
      default:
        throw PixieError{"Invalid path command"};
    }
    i += parameterCount(kind);
    prevCommandKind = kind;
  }
  if (!shape.empty()) {
    if (closeSubpaths) Flattener::addSegment(shape, at, start);
    result.push_back(shape);
  }
  return result;
}

void transformShapes(std::vector<Polygon>& shapes, const Mat3& t) {  // paths.nim:1092-1096
  if (isIdentity(t)) return;
  for (auto& s : shapes)
    for (auto& v : s) v = mul(t, v);
}

// ---------------------------------------------------------------- shapesToSegments (:1059-1090)
inline float quantizeY(float v) {
  // vmath quantize(v, 1/256) = sign(v) * floor(|v| / n) * n  
  const float n = 1.0f / 256.0f;
  float sg = v > 0 ? 1.f : (v < 0 ? -1.f : 0.f);
  return sg * floorf(fabsf(v) / n) * n;
}
void shapesToSegments(const std::vector<Polygon>& shapes, pixie_segments* out) {
  for (const auto& poly : shapes) {
    if (poly.empty()) continue;
    V2 vec1 = v2(poly.back().x, quantizeY(poly.back().y));
    for (size_t i = 0; i < poly.size(); i++) {
      V2 vec2_ = v2(poly[i].x, quantizeY(poly[i].y));
      if (i == 0 && vec1 == vec2_) continue;
      V2 sa = vec1, sb = vec2_;
      vec1 = vec2_;
      if (sa.y == sb.y) continue;  // skip horizontal
      int16_t winding = 1;
      if (sa.y > sb.y) {
        V2 tmp = sa;
        sa = sb;
        sb = tmp;
        winding = -1;
      }
      out->xyxy.push_back(sa.x);
      out->xyxy.push_back(sa.y);
      out->xyxy.push_back(sb.x);
      out->xyxy.push_back(sb.y);
      out->winding.push_back(winding);
    }
  }
}

// ---------------------------------------------------------------- strokeShapes (:1922-2082)
// bumpy intersects(Line, Line, at) — SURVEY Appendix A
bool lineLineIntersects(V2 aa, V2 ab, V2 ba, V2 bb, V2& at) {
  V2 s1 = ab - aa, s2 = bb - ba;
  float den = (-s2.x * s1.y + s1.x * s2.y);
  float t = (s2.x * (aa.y - ba.y) - s2.y * (aa.x - ba.x)) / den;
  if (den == 0) return false;
  at = aa + s1 * t;
  return true;
}
float fixAngle(float angle) {
  float r = angle;
  while ((double)r > kPI) r -= (float)kTAU;
  while ((double)r < -kPI) r += (float)kTAU;
  return r;
}

std::vector<Polygon> strokeShapes(const std::vector<Polygon>& shapes, float strokeWidth, int lineCap,
                                  int lineJoinIn, float miterLimit, const std::vector<float>& dashesIn,
                                  float pixelScale) {
  std::vector<Polygon> result;
  if (strokeWidth <= 0) return result;
  const float halfStroke = strokeWidth / 2;
  const float miterAngleLimit = asinf(1 / miterLimit) * 2;

  auto makeCircle = [&](V2 at) -> Polygon {
    pixie_path p;
    ellipse(&p, at.x, at.y, halfStroke, halfStroke);
    return commandsToShapes(p.commands, true, pixelScale)[0];
  };
  auto makeRect = [&](V2 at, V2 to) -> Polygon {
    V2 tangent = normalize(to - at);
    V2 normal = v2(tangent.y, tangent.x);
    V2 a = v2(at.x + normal.x * halfStroke, at.y - normal.y * halfStroke);
    V2 b = v2(to.x + normal.x * halfStroke, to.y - normal.y * halfStroke);
    V2 c = v2(to.x - normal.x * halfStroke, to.y + normal.y * halfStroke);
    V2 d = v2(at.x - normal.x * halfStroke, at.y + normal.y * halfStroke);
    return Polygon{a, b, c, d, a};
  };
  auto addJoin = [&](std::vector<Polygon>& shape, V2 prevPos, V2 pos, V2 nextPos) {
    float minArea = kPixelErrorMargin / pixelScale;
    if (lineJoinIn == PIXIE_ROUND_JOIN) {
      float area = (float)kPI * halfStroke * halfStroke;
      if (area > minArea) shape.push_back(makeCircle(pos));
      return;
    }
    V2 dn = nextPos - pos, dp = prevPos - pos;
    float angle = fixAngle(atan2f(dn.y, dn.x) - atan2f(dp.y, dp.x));
    if (fabs(fabs((double)angle) - kPI) > (double)kEpsilon) {
      V2 a = normalize(pos - prevPos) * halfStroke;
      V2 b = normalize(pos - nextPos) * halfStroke;
      if (angle >= 0) {
        a = v2(-a.y, a.x);
        b = v2(b.y, -b.x);
      } else {
        a = v2(a.y, -a.x);
        b = v2(-b.y, b.x);
      }
      int lineJoin = lineJoinIn;
      if (lineJoin == PIXIE_MITER_JOIN && fabsf(angle) < miterAngleLimit) lineJoin = PIXIE_BEVEL_JOIN;
      if (lineJoin == PIXIE_MITER_JOIN) {
        V2 at;
        if (lineLineIntersects(prevPos + a, pos + a, nextPos + b, pos + b, at)) {
          float bisectorLengthSq = lengthSq(at - pos);
          float areaSq = 0.25f * (lengthSq(a) * bisectorLengthSq + lengthSq(b) * bisectorLengthSq);
          if (areaSq > (minArea * minArea)) shape.push_back(Polygon{pos + a, at, pos + b, pos, pos + a});
        }
      } else if (lineJoin == PIXIE_BEVEL_JOIN) {
        float areaSq = 0.25f * lengthSq(a) * lengthSq(b);
        if (areaSq > (minArea * minArea)) shape.push_back(Polygon{a + pos, b + pos, pos, a + pos});
      }
    }
  };

  for (const auto& shape : shapes) {
    std::vector<Polygon> shapeStroke;
    const size_t n = shape.size();
    if (shape[0] != shape[n - 1]) {
      if (lineCap == PIXIE_ROUND_CAP) {
        shapeStroke.push_back(makeCircle(shape[0]));
      } else if (lineCap == PIXIE_SQUARE_CAP) {
        V2 tangent = normalize(shape[1] - shape[0]);
        shapeStroke.push_back(makeRect(shape[0] - tangent * halfStroke, shape[0]));
      }
    }
    std::vector<float> dashes = dashesIn;
    if (dashes.size() % 2 != 0) dashes.insert(dashes.end(), dashesIn.begin(), dashesIn.end());
    for (float d : dashes)
      if (d <= 0.0f) throw PixieError{"Invalid line dash value"};

    for (size_t i = 1; i < n; i++) {
      V2 pos = shape[i], prevPos = shape[i - 1];
      if (!dashes.empty()) {
        float distance = length(prevPos - pos);
        V2 dir = normalize(pos - prevPos);
        V2 currPos = prevPos;
        bool done = false;
        while (!done) {
          for (size_t k = 0; k < dashes.size(); k++) {
            float d = dashes[k];
            if (k % 2 == 0) {
              float dd = fminf(distance, d);
              shapeStroke.push_back(makeRect(currPos, currPos + dir * dd));
            }
            currPos = currPos + dir * d;
            distance -= d;
            if (distance <= 0) {
              done = true;
              break;
            }
          }
        }
      } else {
        shapeStroke.push_back(makeRect(prevPos, pos));
      }
      if (i < n - 1) addJoin(shapeStroke, prevPos, pos, shape[i + 1]);
    }
    if (shape[0] == shape[n - 1]) {
      addJoin(shapeStroke, shape[n - 2], shape[n - 1], shape[1]);
    } else {
      if (lineCap == PIXIE_ROUND_CAP) {
        shapeStroke.push_back(makeCircle(shape[n - 1]));
      } else if (lineCap == PIXIE_SQUARE_CAP) {
        V2 tangent = normalize(shape[n - 1] - shape[n - 2]);
        shapeStroke.push_back(makeRect(shape[n - 1] + tangent * halfStroke, shape[n - 1]));
      }
    }
    result.insert(result.end(), shapeStroke.begin(), shapeStroke.end());
  }
  return result;
}

Mat3 matFrom(const float* mat) {
  if (!mat) return identity();
  Mat3 m;
  memcpy(m.m, mat, sizeof(float) * 9);
  return m;
}

template <typename F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const PixieError& e) {
    g_err = e.msg;
    return 1;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 2;
  }
}

}  // namespace

extern "C" {

const char* pixie_host_last_error(void) { return g_err.c_str(); }

pixie_path* pixie_host_path_new(void) { return new pixie_path(); }
void pixie_host_path_free(pixie_path* p) { delete p; }
int pixie_host_path_parse(const char* s, pixie_path** out) {
  pixie_path* p = new pixie_path();
  int rc = guarded([&]() { parsePathInto(std::string(s ? s : ""), p); });
  if (rc) {
    delete p;
    *out = nullptr;
    return rc;
  }
  *out = p;
  return 0;
}
int pixie_host_path_num_commands(const pixie_path* p) { return (int)p->commands.size(); }
int pixie_host_path_commands(const pixie_path* p, float* out, int cap) {
  int n = (int)p->commands.size();
  if (cap < n) n = cap;
  memcpy(out, p->commands.data(), sizeof(float) * n);
  return n;
}
void pixie_host_path_move_to(pixie_path* p, float x, float y) { moveTo(p, x, y); }
void pixie_host_path_line_to(pixie_path* p, float x, float y) { lineTo(p, x, y); }
void pixie_host_path_bezier_curve_to(pixie_path* p, float x1, float y1, float x2, float y2, float x3, float y3) {
  bezierCurveTo(p, x1, y1, x2, y2, x3, y3);
}
void pixie_host_path_quadratic_curve_to(pixie_path* p, float x1, float y1, float x2, float y2) {
  quadraticCurveTo(p, x1, y1, x2, y2);
}
void pixie_host_path_elliptical_arc_to(pixie_path* p, float rx, float ry, float rot, int large, int sweep, float x, float y) {
  ellipticalArcTo(p, rx, ry, rot, large != 0, sweep != 0, x, y);
}
int pixie_host_path_arc(pixie_path* p, float x, float y, float r, float a0, float a1, int ccw) {
  return guarded([&]() { arcBuilder(p, x, y, r, a0, a1, ccw != 0); });
}
int pixie_host_path_arc_to(pixie_path* p, float x1, float y1, float x2, float y2, float r) {
  return guarded([&]() { arcToBuilder(p, x1, y1, x2, y2, r); });
}
void pixie_host_path_rect(pixie_path* p, float x, float y, float w, float h, int clockwise) {
  rect(p, x, y, w, h, clockwise != 0);
}
void pixie_host_path_rounded_rect(pixie_path* p, float x, float y, float w, float h, float nw, float ne,
                                  float se, float sw, int clockwise) {
  roundedRect(p, x, y, w, h, nw, ne, se, sw, clockwise != 0);
}
void pixie_host_path_ellipse(pixie_path* p, float cx, float cy, float rx, float ry) { ellipse(p, cx, cy, rx, ry); }
int pixie_host_path_polygon(pixie_path* p, float x, float y, float size, int sides) {
  return guarded([&]() { polygon(p, x, y, size, sides); });
}
void pixie_host_path_close(pixie_path* p) { closePath(p); }

int pixie_host_fill_segments(const pixie_path* p, const float* mat, pixie_segments** out) {
  pixie_segments* s = new pixie_segments();
  int rc = guarded([&]() {
    Mat3 m = matFrom(mat);
    std::vector<Polygon> shapes = commandsToShapes(p->commands, true, pixelScale(m));
    transformShapes(shapes, m);
    shapesToSegments(shapes, s);
  });
  if (rc) {
    delete s;
    *out = nullptr;
    return rc;
  }
  *out = s;
  return 0;
}

int pixie_host_stroke_segments(const pixie_path* p, const float* mat, float stroke_width, int line_cap,
                               int line_join, float miter_limit, const float* dashes, int num_dashes,
                               pixie_segments** out) {
  pixie_segments* s = new pixie_segments();
  int rc = guarded([&]() {
    Mat3 m = matFrom(mat);
    float ps = pixelScale(m);
    std::vector<float> d(dashes, dashes + (num_dashes > 0 ? num_dashes : 0));
    std::vector<Polygon> shapes = commandsToShapes(p->commands, false, ps);
    std::vector<Polygon> stroke = strokeShapes(shapes, stroke_width, line_cap, line_join, miter_limit, d, ps);
    transformShapes(stroke, m);
    shapesToSegments(stroke, s);
  });
  if (rc) {
    delete s;
    *out = nullptr;
    return rc;
  }
  *out = s;
  return 0;
}

int pixie_host_segments_count(const pixie_segments* s) { return (int)s->winding.size(); }
const float* pixie_host_segments_xyxy(const pixie_segments* s) { return s->xyxy.data(); }
const int16_t* pixie_host_segments_winding(const pixie_segments* s) { return s->winding.data(); }
void pixie_host_segments_free(pixie_segments* s) { delete s; }

int pixie_host_gaussian_kernel(int radius, uint16_t* out) {  // internal.nim:17-34
  if (radius < 0) {
    g_err = "negative radius";
    return 1;
  }
  int n = radius * 2 + 1;
  std::vector<float> floats(n);
  double total = 0.0;
  for (int step = -radius; step <= radius; step++) {
    float s = (float)radius / 2.2f;
    float s2 = s * s;
    float st = (float)step;
    float arg = -1 * (st * st) / (2 * s2);
    double a = 1 / sqrt(2 * kPI * (double)s2) * (double)expf(arg);
    floats[step + radius] = (float)a;
    total += a;
  }
  for (int i = 0; i < n; i++) floats[i] = (float)((double)floats[i] / total);
  for (int i = 0; i < n; i++) out[i] = (uint16_t)roundf(floats[i] * 255 * 256);
  return 0;
}

}  // extern "C"

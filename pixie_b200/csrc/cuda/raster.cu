// K1-K3 — the path rasteriser behind fillPath / strokePath:
//   fillShapes (treeform/pixie src/pixie/paths.nim:1593-1912) from the segment list down.
//
//   K1  partition_kernel   partitionSegments (:1168-1262): order-preserving binning of segments
//                          into equal-height y bands (one warp per band, ballot compaction keeps
//                          segment order), per-band requiresAntiAliasing (:1149-1166), clipping of
//                          spanning entries (:1236-1248), twoNonintersectingSpanningSegments (:1250-1262)
//   K2+K3 raster_kernel    the scanline loop (:1631-1908) fused with the blend:
//                          one warp owns one (layer, scanline) and walks the ordered fill list, so
//                          fills of one canvas keep the reference's sequential semantics with a
//                          single launch and no inter-fill synchronisation; per fill it picks
//                          mode A (pixel-aligned pair :1644-1668), mode B (exact-area trapezoids
//                          :1691-1872) or mode C (computeCoverage :1350-1431, 5 sample lines, with
//                          `walk` :1298-1330 emulated literally) and blends straight into the
//                          canvas (fillCoverage / fillHits :1479-1591, blends.nim).
//
// All geometry is IEEE float32 with one rounding per operation (compiled with -fmad=false,
// -prec-div=true) so that every mode decision equals the CPU reference's.
#include <xmmintrin.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace pixie {

struct FillHeader {
  int segBegin, segCount;
  int startX, startY, pathWidth, pathHeight;
  int numPartitions, partitionHeight;
  int partBase;
  uint32_t rgbx;
  int rule, mode;
  int active;   // 0: pathWidth == 0, the reference returns before touching the image (:1615-1616)
  int wrapRows; // MaskBlend only: rows above a scanline that its negative-x clears can reach (see mask_wrap_clears)
};

struct __align__(16) Entry {
  float ax, ay, bx, by;  // segment.at, segment.to (clipped to the band when spanning)
  float m, b;
  int winding;
  int pad;
};

struct CmdList {
  int w = 0, h = 0, layers = 1;
  int numFills = 0;
  int64_t numSegs = 0, numParts = 0, numEntries = 0;
  int maxEntries = 0;
  float4* segs = nullptr;
  int16_t* wind = nullptr;
  FillHeader* fills = nullptr;
  int2* rowRange = nullptr;
  int* entryOff = nullptr;
  int* layerFillBegin = nullptr;
  Entry* entries = nullptr;
  uint8_t* flags = nullptr;
  uint32_t* scratch = nullptr;  // per-warp spill area when a band has more entries than fit in smem
  unsigned long long* counters = nullptr;  // [0] row ticket, [1] covered px
  int rasterBlocks = 0, warpsPerBlock = 0, scratchWords = 0, covBytes = 0, smemCap = 0;
  size_t smemBytes = 0, h2dBytes = 0;
  uint8_t* block = nullptr;   // block A: host-written inputs + device-made offsets/flags/counters
  uint8_t* blockB = nullptr;  // block B: band entries + spill scratch (sized after the device-side count)
  bool owned = false;
};

static std::unordered_map<uint64_t, CmdList> g_lists;
static uint64_t g_next_list = 1;

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
PXD long long f2ll(float f) { return (long long)f; }          // Nim float32 -> int
PXD int fixed32(float f) { return __float2int_rz(f * 256.0f); }  // paths.nim:1268-1269
PXD int fx_integer(int p) { return p / 256; }                  // :1271-1272 (truncating div)
PXD int fx_trunc(int p) { return (p / 256) * 256; }            // :1274-1275
PXD bool should_fill(int rule, int count) { return rule == 0 ? count != 0 : (count % 2) != 0; }  // :1288-1296
PXD float solve_x(float m, float b, float y) { return m == 0.0f ? b : (y - b) / m; }             // :1137-1141
PXD float frac_vmath(float v) {  // vmath fractional(): abs(v) - floor(abs(v))
  float a = fabsf(v);
  return a - floorf(a);
}
PXD int clampi(long long v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : (int)v); }

// bumpy intersects(Segment, Line) against the horizontal line (0,y)-(1000,y)
PXD bool seg_line(float ax, float ay, float bx, float by, float y, float& ox, float& oy) {
  const float s1x = 1000.0f - 0.0f, s1y = y - y;
  const float s2x = bx - ax, s2y = by - ay;
  const float den = (-s2x * s1y + s1x * s2y);
  const float num = s1x * (y - ay) - s1y * (0.0f - ax);
  const float u = num / den;
  if (u >= 0.0f && u <= 1.0f) {
    ox = ax + u * s2x;
    oy = ay + u * s2y;
    return true;
  }
  return false;
}
PXD bool intersects_inside(const Entry& a, const Entry& b) {  // internal.nim:36-48
  const float s1x = a.bx - a.ax, s1y = a.by - a.ay, s2x = b.bx - b.ax, s2y = b.by - b.ay;
  const float den = (-s2x * s1y + s1x * s2y);
  const float s = (-s1y * (a.ax - b.ax) + s1x * (a.ay - b.ay)) / den;
  const float t = (s2x * (a.ay - b.ay) - s2y * (a.ax - b.ax)) / den;
  return s > 0.0f && s < 1.0f && t > 0.0f && t < 1.0f;
}

// ---------------------------------------------------------------------------------------------
// K1: partitionSegments
// ---------------------------------------------------------------------------------------------
// last fill whose field (segBegin / partBase, non-decreasing over fills) is <= v
template <bool BY_PART>
PXD int find_fill(const FillHeader* __restrict__ fills, int numFills, int v) {
  int lo = 0, hi = numFills;  // first index with key > v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int key = BY_PART ? fills[mid].partBase : fills[mid].segBegin;
    if (key <= v) lo = mid + 1;
    else hi = mid;
  }
  return lo - 1;
}

// K1a: entries per band (one thread per segment; order is irrelevant for counting)
__global__ void __launch_bounds__(256) count_kernel(const FillHeader* __restrict__ fills, int numFills,
                                                    const float4* __restrict__ segs, int numSegs, int* __restrict__ cnt) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < numSegs; i += gridDim.x * blockDim.x) {
    int f = find_fill<false>(fills, numFills, i);
    while (f > 0 && fills[f].segCount == 0) f--;  // empty fills share their segBegin with the next one
    const FillHeader H = fills[f];
    if (!H.active || H.numPartitions <= 0 || i >= H.segBegin + H.segCount) continue;
    if (H.numPartitions == 1) {
      atomicAdd(&cnt[H.partBase], 1);
      continue;
    }
    const float4 s = segs[i];
    const float startYf = (float)(unsigned)H.startY;
    const unsigned ph = (unsigned)H.partitionHeight, lastP = (unsigned)(H.numPartitions - 1);
    unsigned atP = __float2uint_rz(fmaxf(0.0f, s.y - startYf)) / ph;
    unsigned toP = __float2uint_rz(fmaxf(0.0f, s.w - startYf)) / ph;
    atP = min(atP, lastP);
    toP = min(toP, lastP);
    for (unsigned p = atP; p <= toP; p++) atomicAdd(&cnt[H.partBase + (int)p], 1);
  }
}

// K1b: exclusive scan of the band counts in place (cnt[n] receives the total); meta = {total, max}
__global__ void __launch_bounds__(1024) scan_kernel(int* __restrict__ cnt, int n, unsigned long long* __restrict__ meta) {
  __shared__ long long sums[1024];
  __shared__ int maxs[1024];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int b = min(tid * per, n), e = min(b + per, n);
  long long sum = 0;
  int mx = 0;
  for (int i = b; i < e; i++) {
    const int c = cnt[i];
    sum += c;
    mx = max(mx, c);
  }
  sums[tid] = sum;
  maxs[tid] = mx;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan of the per-thread sums
    const long long v = tid >= o ? sums[tid - o] : 0;
    const int m = tid >= o ? maxs[tid - o] : 0;
    __syncthreads();
    sums[tid] += v;
    maxs[tid] = max(maxs[tid], m);
    __syncthreads();
  }
  long long run = sums[tid] - sum;
  for (int i = b; i < e; i++) {
    const int c = cnt[i];
    cnt[i] = (int)run;
    run += c;
  }
  if (tid == 1023) {
    cnt[n] = (int)sums[1023];
    meta[0] = (unsigned long long)sums[1023];
    meta[1] = (unsigned long long)maxs[1023];
  }
}

__global__ void __launch_bounds__(256) partition_kernel(const FillHeader* __restrict__ fills, int numFills,
                                                        const int* __restrict__ entryOff,
                                                        const float4* __restrict__ segs,
                                                        const int16_t* __restrict__ wind, Entry* __restrict__ entries,
                                                        uint8_t* __restrict__ flags, int numParts) {
  const int lane = threadIdx.x & 31;
  const int warpsTotal = (gridDim.x * blockDim.x) >> 5;
  for (int gp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gp < numParts; gp += warpsTotal) {
    const FillHeader H = fills[find_fill<true>(fills, numFills, gp)];
    const int p = gp - H.partBase;
    const int top = H.startY + p * H.partitionHeight;
    const int bottom = (p == H.numPartitions - 1) ? H.pathHeight : top + H.partitionHeight;
    const float topf = (float)top, botf = (float)bottom;
    const float startYf = (float)(unsigned)H.startY;
    const unsigned ph = (unsigned)H.partitionHeight, lastP = (unsigned)(H.numPartitions - 1);
    const int outBase = entryOff[gp];
    int out = outBase;
    bool aa = false;
    constexpr int kChunks = 4;  // 128 segments in flight per iteration: the scan is latency bound
    for (int base0 = 0; base0 < H.segCount; base0 += 32 * kChunks) {
      float4 sv[kChunks];
      int wv[kChunks];
#pragma unroll
      for (int q = 0; q < kChunks; q++) {
        const int i = base0 + q * 32 + lane;
        sv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        wv[q] = 0;
        if (i < H.segCount) {
          sv[q] = segs[H.segBegin + i];
          wv[q] = (int)wind[H.segBegin + i];
        }
      }
#pragma unroll
      for (int q = 0; q < kChunks; q++) {
        const int i = base0 + q * 32 + lane;
        const float4 s = sv[q];
        bool touches = false;
        if (i < H.segCount) {
          if (H.numPartitions == 1) {
            touches = true;
          } else {  // partitionRange (:1201-1213)
            unsigned atP = __float2uint_rz(fmaxf(0.0f, s.y - startYf)) / ph;
            unsigned toP = __float2uint_rz(fmaxf(0.0f, s.w - startYf)) / ph;
            atP = min(atP, lastP);
            toP = min(toP, lastP);
            touches = (unsigned)p >= atP && (unsigned)p <= toP;
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, touches);
        if (touches) {
          Entry e;  // initPartitionEntry (:1127-1135)
          e.ax = s.x; e.ay = s.y; e.bx = s.z; e.by = s.w;
          e.winding = wv[q];
          e.pad = 0;
          e.m = 0.0f;
          e.b = 0.0f;
          const float d = s.x - s.z;
          if (d == 0.0f) {
            e.b = s.x;
          } else {
            e.m = (s.y - s.w) / d;
            e.b = s.y - e.m * s.x;
          }
          // requiresAntiAliasing (:1149-1160) — on the unclipped segment
          if (s.x != s.z || (s.x - truncf(s.x) != 0.0f) || (s.y - truncf(s.y) != 0.0f) || (s.w - truncf(s.w) != 0.0f))
            aa = true;
          // clip entries that span the whole band (:1242-1248)
          if (e.ay <= topf && e.by >= botf) {
            float atx = 0.0f, aty = 0.0f;
            seg_line(e.ax, e.ay, e.bx, e.by, topf, atx, aty);
            e.ax = atx; e.ay = aty;
            seg_line(e.ax, e.ay, e.bx, e.by, botf, atx, aty);
            e.bx = atx; e.by = aty;
          }
          entries[out + __popc(bal & ((1u << lane) - 1u))] = e;
        }
        out += __popc(bal);
      }
    }
    const bool aaAny = __any_sync(0xffffffffu, aa);
    __syncwarp();
    if (lane == 0) {
      bool two = false;
      if (out - outBase == 2) {  // :1250-1262
        const Entry e0 = entries[outBase], e1 = entries[outBase + 1];
        if (!intersects_inside(e0, e1)) {
          if (e0.ay <= topf && e0.by >= botf && e1.ay <= topf && e1.by >= botf) {
            two = true;
            if ((e0.ax + e0.bx) * 0.5f > (e1.ax + e1.bx) * 0.5f) {
              entries[outBase] = e1;
              entries[outBase + 1] = e0;
            }
          }
        }
      }
      flags[gp] = (uint8_t)((aaAny ? 1 : 0) | (two ? 2 : 0));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K2+K3: scanlines
// ---------------------------------------------------------------------------------------------
struct RasterArgs {
  px_t* canvas;
  int w, h, layers;
  const FillHeader* fills;
  const int2* rowRange;        // rows [x, y) of the canvas each fill has to visit (empty when inactive)
  const int* layerFillBegin;
  const int* entryOff;
  const Entry* entries;
  const uint8_t* flags;
  uint32_t* gscratch;          // per-warp global spill (scratchWords words each), may be null
  unsigned long long* counters;
  long long rowBegin, rowEnd;  // flattened (layer * h + y) rows this launch owns
  int ticketSlot;              // counters[ticketSlot] hands out rows
  int scratchBlockBase;        // first block of this launch in the global spill area
  int smemCap;                 // entries whose scratch fits in shared memory
  int scratchCap;              // capacity of the global spill
  int covBytes;                // bytes of the per-warp coverage row in shared memory
  int countCovered;
};

constexpr int kScratchArrays = 8;

constexpr int GenericMode = -1;  // every mode that goes through blender() per pixel, chosen at run time

__device__ __noinline__ px_t blend_px_rt(int mode, px_t b, px_t s) {
  px_t r = b;
  PX_DISPATCH_MODE(mode, r = blend_px<MODE>(b, s));
  return r;
}
template <int MODE>
PXD px_t blend_any(int mode, px_t b, px_t s) {
  if (MODE == GenericMode) return blend_px_rt(mode, b, s);
  return blend_px < MODE == GenericMode ? 0 : MODE > (b, s);
}

struct WarpCtx {
  int mode;         // run-time blend mode (used by the GenericMode instantiation)
  px_t* row;        // canvas row of this warp
  int w;            // canvas width
  int y;
  int lane;
  uint8_t* cov;     // coverage row (index 0 = pixel covBase)
  bool vec_ok;      // rows are 16-byte aligned
  unsigned covered; // per-lane count of pixels touched with non-zero coverage
};

template <int MODE>
PXD px_t span_op(int mode, px_t d, px_t rgbx) {  // fillHits per-pixel op (:1551-1591)
  if (MODE == NormalBlend) return line_normal(d, rgbx);   // blendLineNormal (sse2.nim:568-588)
  if (MODE == MaskBlend) return line_mask(d, rgbx);       // blendLineMask (sse2.nim:669-688)
  if (MODE == OverwriteBlend) return rgbx;
  return blend_any<MODE>(mode, d, rgbx);
}

PXD void clear_span(WarpCtx& c, int x0, int x1) {  // [x0, x1) -> transparent
  x0 = max(x0, 0);
  x1 = min(x1, c.w);
  for (int x = x0 + c.lane; x < x1; x += 32) c.row[x] = 0u;
}

// interior span of fillHits: [x0, x1) gets the solid colour blended in
template <int MODE>
PXD void fill_span(WarpCtx& c, int x0, int x1, px_t rgbx) {
  x0 = max(x0, 0);
  x1 = min(x1, c.w);
  if (x1 <= x0) return;
  const bool store_only = MODE == OverwriteBlend || (MODE == NormalBlend && pA(rgbx) == 255u);
  if (MODE == MaskBlend && pA(rgbx) == 255u) {  // :1576-1577: opaque mask leaves the span untouched
    if (c.lane == 0) c.covered += (unsigned)(x1 - x0);
    return;
  }
  if (c.vec_ok && x1 - x0 >= 64) {
    const int xa = (x0 + 3) & ~3, xb = x1 & ~3;
    for (int x = x0 + c.lane; x < xa; x += 32) {
      c.row[x] = store_only ? rgbx : span_op<MODE>(c.mode, c.row[x], rgbx);
      c.covered++;
    }
    for (int x = xa + 4 * c.lane; x < xb; x += 128) {
      uint4* p = reinterpret_cast<uint4*>(c.row + x);
      uint4 v;
      if (store_only) {
        v = make_uint4(rgbx, rgbx, rgbx, rgbx);
      } else {
        v = *p;
        v.x = span_op<MODE>(c.mode, v.x, rgbx); v.y = span_op<MODE>(c.mode, v.y, rgbx);
        v.z = span_op<MODE>(c.mode, v.z, rgbx); v.w = span_op<MODE>(c.mode, v.w, rgbx);
      }
      *p = v;
      c.covered += 4;
    }
    for (int x = xb + c.lane; x < x1; x += 32) {
      c.row[x] = store_only ? rgbx : span_op<MODE>(c.mode, c.row[x], rgbx);
      c.covered++;
    }
  } else {
    for (int x = x0 + c.lane; x < x1; x += 32) {
      c.row[x] = store_only ? rgbx : span_op<MODE>(c.mode, c.row[x], rgbx);
      c.covered++;
    }
  }
}

// trapezoid edge pixel: blender()(backdrop, rgbx * area) (:1803-1809, :1841-1847)
template <int MODE>
PXD void edge_pixel(WarpCtx& c, int x, float area, px_t rgbx) {
  const px_t src = mul_area(rgbx, area);
  c.row[x] = blend_any<MODE>(c.mode, c.row[x], src);
  c.covered++;
}

// walk (:1298-1330) executed by one lane over sorted hits; spans are appended to spanA/spanB.
PXD int walk_spans(const int* hitAt, const int* hitW, int numHits, int rule, int* spanA, int* spanB) {
  int i = 0, count = 0, prevAt = 0, ns = 0;
  while (i < numHits) {
    const int at = hitAt[i], winding = hitW[i];
    if (at > 0) {
      if (should_fill(rule, count)) {
        if (i < numHits - 1) {
          const int nextAt = hitAt[i + 1], nextWinding = hitW[i + 1];
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
        spanA[ns] = prevAt;
        spanB[ns] = at;
        ns++;
      }
      prevAt = at;
    }
    count += winding;
    i++;
  }
  return ns;
}

// stable sort of n (key, val) pairs by key, ascending; equal keys keep their order
// (= the reference's insertion sorts, :1277-1286 and :1707-1716).  src -> dst arrays.
template <typename K>
PXD void warp_stable_sort(int n, int lane, const K* key, const int* val, K* keyOut, int* valOut) {
  for (int i = lane; i < n; i += 32) {
    const K ki = key[i];
    int r = 0;
    for (int j = 0; j < n; j++) {
      const K kj = key[j];
      r += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
    }
    keyOut[r] = ki;
    valOut[r] = val[i];
  }
}

// fillCoverage (:1479-1526) on the accumulated coverage row, then zero it (:1896).
template <int MODE>
PXD void blend_coverage_row(WarpCtx& c, int startX, int pathWidth, int covBase, px_t rgbx) {
  const int x0 = startX, x1 = startX + pathWidth;
  if (c.vec_ok) {
    // covBase is a multiple of 4: coverage word j <-> pixels covBase + 4j .. +3 (one uint4 of the row)
    const int words = (x1 - covBase + 3) >> 2;
    uint32_t* cw = reinterpret_cast<uint32_t*>(c.cov);
    for (int j = c.lane; j < words; j += 32) {
      const uint32_t cv = cw[j];
      const int x = covBase + 4 * j;
      if (MODE != MaskBlend && cv == 0u) continue;
      cw[j] = 0u;
      uint4* p = reinterpret_cast<uint4*>(c.row + x);
      const bool full = (x >= x0) && (x + 4 <= x1);
      uint4 v;
      if ((MODE == OverwriteBlend || (MODE == NormalBlend && pA(rgbx) == 255u)) && cv == 0xFFFFFFFFu && full) {
        v = make_uint4(rgbx, rgbx, rgbx, rgbx);  // sse2.nim:552-555, 648-651
        *p = v;
        c.covered += 4;
        continue;
      }
      if (MODE == MaskBlend && pA(rgbx) == 255u && cv == 0xFFFFFFFFu && full) {  // sse2.nim:750-751
        c.covered += 4;
        continue;
      }
      v = *p;
      uint32_t* vp = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t cov = (cv >> (8 * k)) & 255u;
        const int xx = x + k;
        if (xx < x0 || xx >= x1) continue;
        if (cov != 0u) c.covered++;
        if (MODE == OverwriteBlend) {
          if (cov != 0u) vp[k] = mul_cov_floor(rgbx, cov);
        } else if (MODE == NormalBlend) {
          if (cov != 0u) vp[k] = line_normal(vp[k], mul_cov_floor(rgbx, cov));
        } else if (MODE == MaskBlend) {
          vp[k] = line_mask(vp[k], mul_cov_floor(rgbx, cov));
        } else {
          if (cov != 0u) vp[k] = blend_any<MODE>(c.mode, vp[k], mul_cov_round(rgbx, cov));
        }
      }
      *p = v;
    }
  } else {
    for (int x = x0 + c.lane; x < x1; x += 32) {
      const uint32_t cov = c.cov[x - covBase];
      c.cov[x - covBase] = 0;
      if (cov != 0u) c.covered++;
      if (MODE == OverwriteBlend) {
        if (cov != 0u) c.row[x] = mul_cov_floor(rgbx, cov);
      } else if (MODE == NormalBlend) {
        if (cov != 0u) c.row[x] = line_normal(c.row[x], mul_cov_floor(rgbx, cov));
      } else if (MODE == MaskBlend) {
        c.row[x] = line_mask(c.row[x], mul_cov_floor(rgbx, cov));
      } else {
        if (cov != 0u) c.row[x] = blend_any<MODE>(c.mode, c.row[x], mul_cov_round(rgbx, cov));
      }
    }
  }
  if (MODE == MaskBlend) {  // :1516-1517
    clear_span(c, 0, x0);
    clear_span(c, x1, c.w);
  }
}

// MaskBlend + trapezoid shortcut with geometry left of the canvas.  The reference clears the gaps between
// fill pairs with clearUnsafe(min(filledTo, width), y, min(clearTo, width), y) (:1856-1866, :1433-1440),
// which addresses the canvas LINEARLY (dataIndex = width * y + x): when filledTo / clearTo are negative
// the cleared range lies in the rows above y, which the same fill has already finished.  Those writes are
// inside the image, so they are part of the reference's result; a warp owns one row, so after its own
// work on row c.y it replays the mode-B decision of row yy > c.y and applies the part of row yy's clears
// that lands on its row.  Clears commute, so the order among the rows below does not matter.
__device__ __noinline__ void mask_wrap_clears(WarpCtx& c, const FillHeader& H, const RasterArgs& A, int yy, uint32_t* sscr,
                                              uint32_t* gscr) {
  const int lane = c.lane, W = c.w;
  int p = (yy - H.startY) / H.partitionHeight;
  if (p > H.numPartitions - 1) p = H.numPartitions - 1;
  const int gp = H.partBase + p;
  const int eBeg = A.entryOff[gp];
  const int eCnt = A.entryOff[gp + 1] - eBeg;
  const unsigned fl = A.flags[gp];
  const bool aa = (fl & 1u) != 0, two = (fl & 2u) != 0;
  if (two && !aa) return;  // mode A clamps its spans to the row
  const Entry* ent = A.entries + eBeg;
  uint32_t* scr = eCnt > A.smemCap ? gscr : sscr;
  const int cap = eCnt > A.smemCap ? A.scratchCap : A.smemCap;
  int* sel = reinterpret_cast<int*>(scr);
  float* tax = reinterpret_cast<float*>(scr + 2 * cap);
  float* tbx = reinterpret_cast<float*>(scr + 3 * cap);
  float* mid = reinterpret_cast<float*>(scr + 4 * cap);
  int* idx = reinterpret_cast<int*>(scr + 5 * cap);
  float* midS = reinterpret_cast<float*>(scr + 6 * cap);
  int* order = reinterpret_cast<int*>(scr + 7 * cap);
  const float scanTop = (float)yy, scanBottom = (float)(yy + 1);
  bool allSpan = true;
  int nsel = 0;
  if (two) {
    nsel = 2;
    if (lane < 2) sel[lane] = lane;
  } else {
    for (int base = 0; base < eCnt; base += 32) {
      const int i = base + lane;
      bool take = false, partial = false;
      if (i < eCnt) {
        const float ay = ent[i].ay, by = ent[i].by;
        take = !(by <= scanTop || ay >= scanBottom);
        partial = take && (ay > scanTop || by < scanBottom);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, take);
      if (__any_sync(0xffffffffu, partial)) allSpan = false;
      if (take) sel[nsel + __popc(bal & ((1u << lane) - 1u))] = i;
      nsel += __popc(bal);
    }
  }
  __syncwarp();
  if (!allSpan || (nsel % 2) != 0) return;  // computeCoverage path: its clears stay inside the row
  for (int s = lane; s < nsel; s += 32) {
    const Entry e = ent[sel[s]];
    const float xa = solve_x(e.m, e.b, scanTop), xb = solve_x(e.m, e.b, scanBottom);
    tax[s] = xa;
    tbx[s] = xb;
    mid[s] = (xa + xb) * 0.5f;
    idx[s] = s;
  }
  __syncwarp();
  warp_stable_sort<float>(nsel, lane, mid, idx, midS, order);
  __syncwarp();
  bool ok = true;
  for (int i = lane; i < nsel - 1; i += 32) {
    const int l = order[i], r = order[i + 1];
    if (f2ll(ceilf(fmaxf(tax[l], tbx[l]))) > f2ll(fminf(tax[r], tbx[r]))) ok = false;
  }
  ok = __all_sync(0xffffffffu, ok);
  if (ok) {
    int carry = 0;
    for (int base = 0; base < nsel; base += 32) {
      const int i = base + lane;
      int pre = (i < nsel) ? ent[sel[order[i]]].winding : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += t;
      }
      pre += carry;
      if (i < nsel) {
        const bool f = should_fill(H.rule, pre);
        if (((i & 1) == 0) ? !f : f) ok = false;
      }
      carry = __shfl_sync(0xffffffffu, pre, 31);
    }
    ok = __all_sync(0xffffffffu, ok);
  }
  if (!ok) return;
  // linear pixel index of x on row yy, relative to the start of row c.y
  const long long rowOff = (long long)(yy - c.y) * W;
  long long filledTo = 0;
  for (int i = 0; i < nsel; i += 2) {
    const int ls = order[i], rs = order[i + 1];
    const long long clearTo = f2ll(fminf(tax[ls], tbx[ls]));
    const long long a = filledTo < W ? filledTo : W, b = clearTo < W ? clearTo : W;
    if (a != W) {
      const long long x0 = rowOff + a, x1 = rowOff + b;
      if (x1 > 0 && x0 < W) clear_span(c, (int)(x0 > 0 ? x0 : 0), (int)(x1 < W ? x1 : W));
    }
    filledTo = f2ll(ceilf(fmaxf(tax[rs], tbx[rs])));
  }
  const long long a = filledTo < W ? filledTo : W;
  if (a != W) {
    const long long x0 = rowOff + a;
    if (x0 < W) clear_span(c, (int)(x0 > 0 ? x0 : 0), W);
  }
  __syncwarp();
}

// One fill on one scanline.  Returns nothing; all lanes of the warp participate.
template <int MODE>
__device__ __noinline__ void fill_row(WarpCtx& c, const FillHeader& H, const RasterArgs& A, uint32_t* scr, int cap) {
  const int lane = c.lane, y = c.y, W = c.w;
  const px_t rgbx = H.rgbx;
  const int rule = H.rule;

  // rows the path does not reach: only MaskBlend touches them (:1910-1912)
  if (y < H.startY || y >= H.pathHeight) {
    if (MODE == MaskBlend) clear_span(c, 0, W);
    return;
  }
  int p = (y - H.startY) / H.partitionHeight;
  if (p > H.numPartitions - 1) p = H.numPartitions - 1;
  const int gp = H.partBase + p;
  const int eBeg = A.entryOff[gp];
  const int eCnt = A.entryOff[gp + 1] - eBeg;
  const unsigned fl = A.flags[gp];
  const bool aa = (fl & 1u) != 0, two = (fl & 2u) != 0;
  const Entry* ent = A.entries + eBeg;

  if (two && !aa) {  // mode A (:1644-1668): two vertical pixel-aligned lines
    const int minX = clampi(f2ll(ent[0].ax), 0, W), maxX = clampi(f2ll(ent[1].ax), 0, W);
    if (maxX > minX) {
      if (MODE == MaskBlend) clear_span(c, 0, minX);
      fill_span<MODE>(c, minX, maxX, rgbx);
      if (MODE == MaskBlend) clear_span(c, maxX, W);
    } else if (MODE == MaskBlend) {
      clear_span(c, 0, W);
    }
    return;
  }

  int* sel = reinterpret_cast<int*>(scr);
  int* sel2 = sel + cap;
  float* tax = reinterpret_cast<float*>(scr + 2 * cap);
  float* tbx = reinterpret_cast<float*>(scr + 3 * cap);
  int* hitAt = reinterpret_cast<int*>(scr + 4 * cap);
  int* hitW = reinterpret_cast<int*>(scr + 5 * cap);
  int* hitAt2 = reinterpret_cast<int*>(scr + 6 * cap);
  int* hitW2 = reinterpret_cast<int*>(scr + 7 * cap);

  const float scanTop = (float)y, scanBottom = (float)(y + 1);
  bool allSpan = true;
  int nsel = 0;
  const int* selC = sel;  // entry order seen by computeCoverage
  if (two) {
    nsel = 2;
    if (lane < 2) sel[lane] = lane;
  } else {  // :1681-1689
    for (int base = 0; base < eCnt; base += 32) {
      const int i = base + lane;
      bool take = false, partial = false;
      if (i < eCnt) {
        const float ay = ent[i].ay, by = ent[i].by;
        take = !(by <= scanTop || ay >= scanBottom);
        partial = take && (ay > scanTop || by < scanBottom);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, take);
      if (__any_sync(0xffffffffu, partial)) allSpan = false;
      if (take) sel[nsel + __popc(bal & ((1u << lane) - 1u))] = i;
      nsel += __popc(bal);
    }
  }
  __syncwarp();

#ifdef PIXIE_DEBUG_ROW
  if (y == PIXIE_DEBUG_ROW && lane == 0) printf("DBG y=%d eCnt=%d nsel=%d allSpan=%d aa=%d two=%d cap=%d\n", y, eCnt, nsel, (int)allSpan, (int)aa, (int)two, cap);
#endif
  if (allSpan && (nsel % 2) == 0) {  // mode B (:1691-1872)
    float* mid = reinterpret_cast<float*>(hitAt);  // sort keys
    for (int s = lane; s < nsel; s += 32) {
      const Entry e = ent[sel[s]];
      const float xa = solve_x(e.m, e.b, scanTop), xb = solve_x(e.m, e.b, scanBottom);
      tax[s] = xa;
      tbx[s] = xb;
      mid[s] = (xa + xb) * 0.5f;
      hitW[s] = s;  // payload: position in sel
    }
    __syncwarp();
    float* midS = reinterpret_cast<float*>(hitAt2);
    warp_stable_sort<float>(nsel, lane, mid, hitW, midS, hitW2);  // hitW2[r] = selected position, sorted by mid x
    __syncwarp();
    bool ok = true;
    for (int i = lane; i < nsel - 1; i += 32) {  // partial-coverage areas must not overlap (:1720-1728)
      const int l = hitW2[i], r = hitW2[i + 1];
      const float leftMaxX = fmaxf(tax[l], tbx[l]), rightMinX = fminf(tax[r], tbx[r]);
      if (f2ll(ceilf(leftMaxX)) > f2ll(rightMinX)) ok = false;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (ok) {  // only simple fill pairs (:1732-1744): prefix winding count, filled after even, empty after odd
      int carry = 0;
      for (int base = 0; base < nsel; base += 32) {
        const int i = base + lane;
        int wv = (i < nsel) ? ent[sel[hitW2[i]]].winding : 0;
        int pre = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, pre, o);
          if (lane >= o) pre += t;
        }
        pre += carry;
        if (i < nsel) {
          const bool f = should_fill(rule, pre);
          if (((i & 1) == 0) ? !f : f) ok = false;
        }
        carry = __shfl_sync(0xffffffffu, pre, 31);
      }
      ok = __all_sync(0xffffffffu, ok);
    }
#ifdef PIXIE_DEBUG_ROW
    if (y == PIXIE_DEBUG_ROW && lane == 0) {
      printf("DBG modeB ok=%d order:", (int)ok);
      for (int i = 0; i < nsel; i++) printf(" %d(%.6f)", sel[hitW2[i]], midS[i]);
      printf("\n");
    }
#endif
    if (ok) {
      int filledTo = 0;
      for (int i = 0; i < nsel; i += 2) {
        const int ls = hitW2[i], rs = hitW2[i + 1];
        const Entry left = ent[sel[ls]], right = ent[sel[rs]];
        const float lax = tax[ls], lbx = tbx[ls], rax = tax[rs], rbx = tbx[rs];
        const float leftMaxX = fmaxf(lax, lbx), rightMinX = fminf(rax, rbx);
        const long long leftCoverEnd = f2ll(ceilf(leftMaxX)), rightCoverBegin = f2ll(truncf(rightMinX));
        {  // left-side partial coverage (:1772-1809)
          const bool inverted = lax < lbx;
          const float sliverStart = fminf(lax, lbx), rectStart = leftMaxX;
          const long long xFirst = f2ll(sliverStart), xEnd = f2ll(ceilf(rectStart));
          const long long xb = xFirst > 0 ? xFirst : 0, xe = xEnd < W ? xEnd : W;
          for (long long xl = xb + lane; xl < xe; xl += 32) {
            const int x = (int)xl;
            const float prevPen = (xl == xFirst) ? sliverStart : (float)x;
            const float prevPenY = (xl == xFirst) ? (inverted ? (float)y : (float)(y + 1)) : (left.m * (float)x + left.b);
            float pen = (float)(x + 1), rightRectArea = 0.0f;
            if (pen > rectStart) {
              rightRectArea = pen - rectStart;
              pen = rectStart;
            }
            const float penY = left.m * pen + left.b;
            const float run = pen - prevPen;
            const float triangleArea = 0.5f * run * fabsf(penY - prevPenY);
            const float rectArea = inverted ? (prevPenY - (float)y) * run : ((float)(y + 1) - prevPenY) * run;
            edge_pixel<MODE>(c, x, triangleArea + rectArea + rightRectArea, rgbx);
          }
        }
        {  // right-side partial coverage (:1811-1847)
          const bool inverted = rax > rbx;
          const float rectEnd = rightMinX, sliverEnd = fmaxf(rax, rbx);
          const long long xFirst = f2ll(rectEnd), xEnd = f2ll(ceilf(sliverEnd));
          const long long xb = xFirst > 0 ? xFirst : 0, xe = xEnd < W ? xEnd : W;
          for (long long xl = xb + lane; xl < xe; xl += 32) {
            const int x = (int)xl;
            const float prevPen = (xl == xFirst) ? rectEnd : (float)x;
            const float prevPenY = (xl == xFirst) ? (inverted ? (float)(y + 1) : (float)y) : (right.m * (float)x + right.b);
            float pen = (float)(x + 1);
            const float leftRectArea = frac_vmath(prevPen);
            if (pen > sliverEnd) pen = sliverEnd;
            const float penY = right.m * pen + right.b;
            const float run = pen - prevPen;
            const float triangleArea = 0.5f * run * fabsf(penY - prevPenY);
            const float rectArea = inverted ? (penY - (float)y) * run : ((float)(y + 1) - penY) * run;
            edge_pixel<MODE>(c, x, leftRectArea + triangleArea + rectArea, rgbx);
          }
        }
        const int fillBegin = clampi(leftCoverEnd, 0, W), fillEnd = clampi(rightCoverBegin, 0, W);
        fill_span<MODE>(c, fillBegin, fillEnd, rgbx);  // fillHits(..., maskClears = false) (:1849-1854)
        if (MODE == MaskBlend) {
          const long long clearTo = f2ll(fminf(lax, lbx));
          clear_span(c, min(filledTo, W), (int)(clearTo < W ? clearTo : W));
        }
        filledTo = clampi(f2ll(ceilf(fmaxf(rax, rbx))), INT_MIN / 2, W);
      }
      if (MODE == MaskBlend) clear_span(c, min(filledTo, W), W);
      return;
    }
    // The reference sorts entryIndices in place (:1707-1716) before it decides whether the shortcut
    // applies, so when it falls through, computeCoverage walks the entries in mid-x order.
    for (int i = lane; i < nsel; i += 32) sel2[i] = sel[hitW2[i]];
    selC = sel2;
    __syncwarp();
  }

  // mode C: computeCoverage (:1350-1431)
  const int quality = aa ? 5 : 1;
  const int sampleCoverage = 255 / quality;
  const float offset = 1.0f / (float)quality;
  const float initialOffset = offset / 2.0f + (float)(0.0001 * 3.141592653589793238462643383279502884);
  const int covBase = c.vec_ok ? (H.startX & ~3) : H.startX;
  const int covX0 = H.startX, covX1 = H.startX + H.pathWidth;  // coverages[] of the reference spans [covX0, covX1)
  float yLine = (float)y + initialOffset - offset;
  const float wf = (float)W;
  int numHits = 0;
  for (int m = 0; m < quality; m++) {
    yLine += offset;
    numHits = 0;
    for (int base = 0; base < nsel; base += 32) {  // :1373-1385 (entry order preserved)
      const int s = base + lane;
      bool hit = false;
      int at = 0, wv = 0;
      if (s < nsel) {
        const Entry e = ent[selC[s]];
        if (e.ay <= yLine && e.by >= yLine) {
          float x = e.m == 0.0f ? e.b : (yLine - e.b) / e.m;
          x = (x != x) ? wf : (x < wf ? x : wf);  // min(x, width.float32)
          at = fixed32(x);
          wv = e.winding;
          hit = true;
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int pos = numHits + __popc(bal & ((1u << lane) - 1u));
        hitAt[pos] = at;
        hitW[pos] = wv;
      }
      numHits += __popc(bal);
    }
    __syncwarp();
    if (numHits > 0) warp_stable_sort<int>(numHits, lane, hitAt, hitW, hitAt2, hitW2);
    __syncwarp();
    // walk -> spans (re-using tax/tbx as int span arrays)
    int* spanA = reinterpret_cast<int*>(tax);
    int* spanB = reinterpret_cast<int*>(tbx);
    int ns = 0;
    if (lane == 0) ns = walk_spans(hitAt2, hitW2, numHits, rule, spanA, spanB);
    ns = __shfl_sync(0xffffffffu, ns, 0);
    __syncwarp();
#ifdef PIXIE_DEBUG_ROW
    if (y == PIXIE_DEBUG_ROW && lane == 0) {
      printf("DBG m=%d yLine=%.9g hits:", m, yLine);
      for (int i = 0; i < numHits; i++) printf(" %d/%d", hitAt2[i], hitW2[i]);
      printf(" spans:");
      for (int i = 0; i < ns; i++) printf(" [%d,%d)", spanA[i], spanB[i]);
      printf("\n");
    }
#endif
    if (aa) {
      for (int k = 0; k < ns; k++) {  // :1391-1431
        const int prevAt = spanA[k], at = spanB[k];
        int fillStart = fx_integer(prevAt);
        const bool pixelCrossed = fx_integer(at) != fx_integer(prevAt);
        const int leftCover = pixelCrossed ? fx_trunc(prevAt) + 256 - prevAt : at - prevAt;
        if (leftCover != 0) {
          fillStart++;
          if (lane == 0) {
            const int px = fx_integer(prevAt), idx = px - covBase;
            if (px >= covX0 && px < covX1) c.cov[idx] = (uint8_t)(c.cov[idx] + (uint8_t)fx_integer(leftCover * sampleCoverage));
          }
        }
        if (pixelCrossed && lane == 0) {
          const int rightCover = at - fx_trunc(at);
          if (rightCover > 0) {
            const int px = fx_integer(at), idx = px - covBase;
            if (px >= covX0 && px < covX1) c.cov[idx] = (uint8_t)(c.cov[idx] + (uint8_t)fx_integer(rightCover * sampleCoverage));
          }
        }
        // interior of the span: +sampleCoverage per pixel.  Spans of one sample line are disjoint, so
        // a coverage byte never exceeds 255 and whole words can be added without carries.
        const int i0 = max(fillStart, covX0) - covBase, i1 = min(fx_integer(at), covX1) - covBase;
        if (i1 - i0 >= 96) {
          const int a0 = (i0 + 3) & ~3, a1 = i1 & ~3;
          for (int j = i0 + lane; j < a0; j += 32) c.cov[j] = (uint8_t)(c.cov[j] + sampleCoverage);
          uint32_t* cw = reinterpret_cast<uint32_t*>(c.cov);
          const uint32_t add4 = 0x01010101u * (uint32_t)sampleCoverage;
          for (int w = (a0 >> 2) + lane; w < (a1 >> 2); w += 32) cw[w] += add4;
          for (int j = a1 + lane; j < i1; j += 32) c.cov[j] = (uint8_t)(c.cov[j] + sampleCoverage);
        } else {
          for (int j = i0 + lane; j < i1; j += 32) c.cov[j] = (uint8_t)(c.cov[j] + sampleCoverage);
        }
        __syncwarp();
      }
    } else {  // fillHits over walkInteger (:1897-1906, :1540-1591)
      int filledTo = H.startX;
      for (int k = 0; k < ns; k++) {
        const int start = fx_integer(spanA[k]);
        const int len = fx_integer(spanB[k]) - start;
        if (len <= 0) continue;
        if (MODE == MaskBlend) clear_span(c, filledTo, start);
        fill_span<MODE>(c, start, start + len, rgbx);
        filledTo = start + len;
      }
      if (MODE == MaskBlend) {
        clear_span(c, 0, H.startX);
        clear_span(c, filledTo, W);
      }
    }
  }
  if (aa) {
    __syncwarp();
    blend_coverage_row<MODE>(c, H.startX, H.pathWidth, covBase, rgbx);
  }
}

__global__ void __launch_bounds__(256) raster_kernel(const RasterArgs A) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warpsPerBlock = blockDim.x >> 5;
  const int scratchWordsSmem = A.smemCap * kScratchArrays;
  uint8_t* mine = smem_raw + (size_t)warp * ((size_t)A.covBytes + (size_t)scratchWordsSmem * 4);
  uint8_t* cov = mine;
  uint32_t* sscr = reinterpret_cast<uint32_t*>(mine + A.covBytes);
  uint32_t* gscr = A.gscratch
                       ? A.gscratch + (size_t)((A.scratchBlockBase + blockIdx.x) * warpsPerBlock + warp) * ((size_t)A.scratchCap * kScratchArrays)
                       : nullptr;
  for (int i = lane * 4; i < A.covBytes; i += 128) *reinterpret_cast<uint32_t*>(cov + i) = 0u;
  __syncwarp();

  unsigned covered = 0;
  while (true) {
    unsigned long long ticket = 0;
    if (lane == 0) ticket = atomicAdd(&A.counters[A.ticketSlot], 1ull);
    ticket = __shfl_sync(0xffffffffu, ticket, 0) + (unsigned long long)A.rowBegin;
    if ((long long)ticket >= A.rowEnd) break;
    const int layer = (int)(ticket / (unsigned)A.h), y = (int)(ticket % (unsigned)A.h);
    WarpCtx c;
    c.row = A.canvas + ((size_t)layer * A.h + y) * A.w;
    c.w = A.w;
    c.y = y;
    c.lane = lane;
    c.cov = cov;
    c.vec_ok = (A.w & 3) == 0 && (reinterpret_cast<uintptr_t>(A.canvas) & 15) == 0;
    c.covered = 0;
    const int f0 = A.layerFillBegin[layer], f1 = A.layerFillBegin[layer + 1];
    for (int fb = f0; fb < f1; fb += 32) {  // 32 fills per step: which of them touch this row?
      const int fl = fb + lane;
      bool act = false;
      if (fl < f1) {
        const int2 rr = A.rowRange[fl];
        act = y >= rr.x && y < rr.y;
      }
      unsigned todo = __ballot_sync(0xffffffffu, act);
      while (todo) {  // ascending order = the reference's sequential order of fills
        const int f = fb + __ffs(todo) - 1;
        todo &= todo - 1;
        const FillHeader H = A.fills[f];
        // scratch: shared memory when the band's entries fit, else the global spill area
        uint32_t* scr = sscr;
        int cap = A.smemCap;
        if (y >= H.startY && y < H.pathHeight) {
          int p = (y - H.startY) / H.partitionHeight;
          if (p > H.numPartitions - 1) p = H.numPartitions - 1;
          const int gp = H.partBase + p;
          const int eCnt = A.entryOff[gp + 1] - A.entryOff[gp];
          if (eCnt > A.smemCap) {
            scr = gscr;
            cap = A.scratchCap;
          }
        }
        c.mode = H.mode;
        if (H.mode == NormalBlend) fill_row<NormalBlend>(c, H, A, scr, cap);
        else if (H.mode == OverwriteBlend) fill_row<OverwriteBlend>(c, H, A, scr, cap);
        else if (H.mode == MaskBlend) fill_row<MaskBlend>(c, H, A, scr, cap);
        else fill_row<GenericMode>(c, H, A, scr, cap);
        __syncwarp();
        if (H.wrapRows > 0 && y >= H.startY) {  // only MaskBlend fills reaching left of the canvas
          const int yEnd = min(H.pathHeight, y + 1 + H.wrapRows);
          for (int yy = y + 1; yy < yEnd; yy++) mask_wrap_clears(c, H, A, yy, sscr, gscr);
          __syncwarp();
        }
      }
    }
    covered += c.covered;
  }
  if (A.countCovered) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) covered += __shfl_xor_sync(0xffffffffu, covered, o);
    if (lane == 0 && covered) atomicAdd(&A.counters[1], (unsigned long long)covered);
  }
}

// ---------------------------------------------------------------------------------------------
// host: command lists
// ---------------------------------------------------------------------------------------------
static inline int64_t f2i_host(float f) {  // Nim float32 -> int on x86-64 (cvttss2si)
  if (!(f > -9.2e18f && f < 9.2e18f)) return INT64_MIN;
  return (int64_t)f;
}
static inline uint32_t f2u_host(float f) {  // matches __float2uint_rz (saturating)
  if (!(f > 0.0f)) return 0u;
  if (f >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)f;
}

static void free_list(CmdList& L) {
  if (L.owned && L.block) cudaFree(L.block);
  if (L.owned && L.blockB) cudaFree(L.blockB);
  L.block = L.blockB = nullptr;
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static const bool g_trace = getenv("PIXIE_CUDA_TRACE") != nullptr;

static int build_list(CmdList& L, bool arena, int w, int h, int layers, int numFills, const int32_t* layerOf,
                      const float* seg, const int16_t* wind, const int32_t* segOff, const uint32_t* rgbx,
                      const uint8_t* rule, const uint8_t* mode) {
  Runtime& r = rt();
  const double t0 = now_ms();
  if (w <= 0 || h <= 0 || layers <= 0) return fail_pixie("Image width and height must be > 0");
  if (numFills < 0) return fail_pixie("negative fill count");
  L.w = w; L.h = h; L.layers = layers; L.numFills = numFills;
  std::vector<FillHeader> fills(numFills);
  std::vector<int> layerBegin(layers + 1, 0);
  int64_t numPartsTotal = 0;
  int prevLayer = 0;
  double tBounds = 0;
  for (int k = 0; k < numFills; k++) {
    const int layer = layerOf ? layerOf[k] : 0;
    if (layer < prevLayer || layer >= layers || layer < 0) return fail_pixie("layer_of_fill must be non-decreasing and < layers");
    prevLayer = layer;
    layerBegin[layer + 1] = k + 1;
    if (mode[k] >= NumBlendModes || rule[k] > 1) return fail_pixie("invalid blend mode / winding rule");
    const int s0 = segOff[k], s1 = segOff[k + 1];
    if (s1 < s0) return fail_pixie("seg_offsets must be non-decreasing");
    const int n = s1 - s0;
    FillHeader& H = fills[k];
    memset(&H, 0, sizeof(H));
    H.segBegin = s0; H.segCount = n; H.rgbx = rgbx[k]; H.rule = rule[k]; H.mode = mode[k];
    if (n == 0) {  // empty path: nothing is drawn (the reference's tiger has one, "M-65.4,9z")
      H.active = 0;
      H.partBase = (int)numPartsTotal;
      continue;
    }
    // computeBounds (:1098-1117) + snapToPixels (common.nim:92-101) + clip to the image (:1605-1613)
    float xMin = INFINITY, xMax = -INFINITY, yMin = INFINITY, yMax = -INFINITY;
    const double tb0 = g_trace ? now_ms() : 0;
    // Nim's min/max (`if x <= y: x else: y`) == MINPS/MAXPS(acc, v) including their NaN behaviour
    // (second operand when unordered); one 4-lane min + max per segment {at.x, at.y, to.x, to.y}.
    {
      const float* sp = seg + 4 * (size_t)s0;
      __m128 vmin = _mm_set1_ps(INFINITY), vmax = _mm_set1_ps(-INFINITY);
      for (int i = 0; i < n; i++, sp += 4) {
        const __m128 v = _mm_loadu_ps(sp);
        vmin = _mm_min_ps(vmin, v);
        vmax = _mm_max_ps(vmax, v);
      }
      float mn[4], mx[4];
      _mm_storeu_ps(mn, vmin);
      _mm_storeu_ps(mx, vmax);
      xMin = mn[0] <= mn[2] ? mn[0] : mn[2];
      xMax = mx[2] <= mx[0] ? mx[0] : mx[2];
      yMin = mn[1];  // at.y < to.y for every segment
      yMax = mx[3];
    }
    if (g_trace) tBounds += now_ms() - tb0;
    float bx_ = 0, by_ = 0, bw_ = 0, bh_ = 0;
    if (!(xMin != xMin || xMax != xMax || yMin != yMin || yMax != yMax)) {
      bx_ = xMin; by_ = yMin; bw_ = xMax - xMin; bh_ = yMax - yMin;
    }
    const float sx = floorf(bx_), sw = ceilf(bx_ + bw_) - sx;
    const float sy = floorf(by_), sh = ceilf(by_ + bh_) - sy;
    const int64_t startX = std::max<int64_t>(0, f2i_host(sx)), startY = std::max<int64_t>(0, f2i_host(sy));
    int64_t pathWidth = 0;
    if (startX < w) pathWidth = std::min<int64_t>(f2i_host(sw), w - startX);
    const int64_t pathHeight = std::min<int64_t>(h, f2i_host(sy + sh));
    if (pathWidth == 0) {  // :1615-1616
      H.active = 0;
      H.partBase = (int)numPartsTotal;
      continue;
    }
    if (pathWidth < 0) return fail_pixie("Path int overflow detected");  // :1618-1619
    H.active = 1;
    H.startX = (int)startX;
    H.pathWidth = (int)pathWidth;
    H.partBase = (int)numPartsTotal;
    if (pathHeight <= startY) {  // no scanline is touched; MaskBlend still clears the canvas
      H.startY = (int)std::min<int64_t>(startY, h);
      H.pathHeight = H.startY;
      H.numPartitions = 0;
      H.partitionHeight = 1;
      continue;
    }
    H.startY = (int)startY;
    H.pathHeight = (int)pathHeight;
    // partitionSegments sizing (:1172-1180)
    const int64_t height = pathHeight - startY;
    const uint32_t maxPartitions = (uint32_t)std::max<int64_t>(1, height / 4);
    const uint32_t numPartitions = std::min<uint32_t>(maxPartitions, (uint32_t)std::max<int>(1, n / 2));
    const uint32_t partitionHeight = (uint32_t)height / numPartitions;
    H.numPartitions = (int)numPartitions;
    H.partitionHeight = (int)partitionHeight;
    if (H.mode == MaskBlend && xMin < 0.0f) {  // rows a negative-x clearUnsafe range can reach back (mask_wrap_clears)
      const double reach = (-(double)floorf(xMin) + 1.0 + (double)w - 1.0) / (double)w;
      H.wrapRows = reach >= (double)h ? h : (int)reach;
    }
    numPartsTotal += numPartitions;
    if (numPartsTotal > 0x3fffffff) return fail_pixie("command list too large");
  }
  for (int l = 1; l <= layers; l++) layerBegin[l] = std::max(layerBegin[l], layerBegin[l - 1]);

  const double t1 = now_ms();
  const int64_t numSegs = numFills ? segOff[numFills] : 0;
  L.numSegs = numSegs;
  L.numParts = numPartsTotal;

  // raster launch geometry: persistent warps, one (layer, row) ticket at a time
  L.covBytes = ((w + 7) & ~3) + 4;             // coverage row, word aligned, with the covBase slack
  L.smemCap = 64;                              // entries per band handled from shared memory
  while (L.smemCap > 8 && (size_t)L.covBytes + (size_t)L.smemCap * kScratchArrays * 4 > 24 * 1024) L.smemCap /= 2;
  size_t perWarp = (size_t)L.covBytes + (size_t)L.smemCap * kScratchArrays * 4;
  L.warpsPerBlock = 8;
  while (L.warpsPerBlock > 1 && perWarp * L.warpsPerBlock > 96 * 1024) L.warpsPerBlock /= 2;
  if (perWarp * L.warpsPerBlock > 200 * 1024) return fail_pixie("canvas too wide for the shared-memory coverage row");
  L.smemBytes = perWarp * L.warpsPerBlock;
  const long long totalRows = (long long)layers * h;
  if (L.smemBytes > 48 * 1024)
    PX_CUDA(cudaFuncSetAttribute(raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemBytes));
  int blocksPerSm = 1;  // resident CTAs per SM for this shared-memory footprint: one wave, rows by ticket
  PX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, raster_kernel, L.warpsPerBlock * 32, L.smemBytes));
  blocksPerSm = std::max(1, blocksPerSm);
  long long wantBlocks = (totalRows + L.warpsPerBlock - 1) / L.warpsPerBlock;
  L.rasterBlocks = (int)std::min<long long>(wantBlocks, (long long)r.num_sms * blocksPerSm);
  L.rasterBlocks = std::max(L.rasterBlocks, 1);

  // Device block A.  The host-written part (segments, windings, headers, row ranges, layer table)
  // is contiguous so that it moves with a single H2D copy from a staging buffer; band counts,
  // entry offsets, flags and counters are produced on the device.
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t P = (size_t)numPartsTotal;
  size_t off = 0;
  const size_t oSegs = off;      off = al(off + (size_t)numSegs * 16);
  const size_t oWind = off;      off = al(off + (size_t)numSegs * 2);
  const size_t oFills = off;     off = al(off + fills.size() * sizeof(FillHeader));
  const size_t oRowRange = off;  off = al(off + fills.size() * sizeof(int2));
  const size_t oLayer = off;     off = al(off + layerBegin.size() * 4);
  const size_t h2dBytes = off;
  const size_t oEntryOff = off;  off = al(off + (P + 1) * 4);   // band counts, scanned in place to offsets
  const size_t oFlags = off;     off = al(off + std::max<size_t>(1, P));
  const size_t oCounters = off;  off = al(off + 128);           // [0] row ticket, [1] covered px, [2] entries, [3] max, [4..] band tickets
  const size_t totalA = off;

  uint8_t* stage = nullptr;
  std::vector<uint8_t> pageable;
  if (arena) {  // per-call lists: library-owned growing arena + pinned staging, no malloc/free per call
    void *blk, *pin;
    if (int rc = get_scratch(2, totalA, &blk)) return rc;
    if (int rc = staging_acquire(h2dBytes, &pin)) return rc;
    L.block = (uint8_t*)blk;
    L.owned = false;
    stage = (uint8_t*)pin;
  } else {
    PX_CUDA(cudaMalloc(&L.block, totalA));
    L.owned = true;
    pageable.resize(h2dBytes);
    stage = pageable.data();
  }
  if (numSegs) {
    memcpy(stage + oSegs, seg, (size_t)numSegs * 16);
    memcpy(stage + oWind, wind, (size_t)numSegs * 2);
  }
  if (!fills.empty()) {
    memcpy(stage + oFills, fills.data(), fills.size() * sizeof(FillHeader));
    int2* rr = reinterpret_cast<int2*>(stage + oRowRange);
    for (size_t k = 0; k < fills.size(); k++) {
      const FillHeader& F = fills[k];
      if (!F.active) rr[k] = make_int2(0, 0);
      else if (F.mode == MaskBlend) rr[k] = make_int2(0, h);  // clears every row it does not cover
      else rr[k] = make_int2(F.startY, F.pathHeight);
    }
  }
  memcpy(stage + oLayer, layerBegin.data(), layerBegin.size() * 4);
  PX_CUDA(cudaMemcpyAsync(L.block, stage, h2dBytes, cudaMemcpyHostToDevice, r.stream));
  if (arena) {
    if (int rc = staging_release()) return rc;
  }
  L.segs = (float4*)(L.block + oSegs);
  L.wind = (int16_t*)(L.block + oWind);
  L.fills = (FillHeader*)(L.block + oFills);
  L.rowRange = (int2*)(L.block + oRowRange);
  L.layerFillBegin = (int*)(L.block + oLayer);
  L.entryOff = (int*)(L.block + oEntryOff);
  L.flags = L.block + oFlags;
  L.counters = (unsigned long long*)(L.block + oCounters);
  L.h2dBytes = h2dBytes;

  // K1a/K1b on the device: how many entries each band gets (partitionRange :1201-1213), exclusive
  // scan to entry offsets, total and maximum -> 16 bytes back to the host to size block B.
  long long meta[2] = {0, 2};
  if (P > 0) {
    PX_CUDA(cudaMemsetAsync(L.entryOff, 0, (P + 1) * 4, r.stream));
    const int cblocks = (int)std::min<int64_t>((numSegs + 255) / 256, (int64_t)r.num_sms * 8);
    count_kernel<<<std::max(cblocks, 1), 256, 0, r.stream>>>(L.fills, numFills, L.segs, (int)numSegs, L.entryOff);
    PX_LAUNCHED();
    scan_kernel<<<1, 1024, 0, r.stream>>>(L.entryOff, (int)P, L.counters + 2);
    PX_LAUNCHED();
    PX_CUDA(cudaMemcpyAsync(meta, L.counters + 2, 16, cudaMemcpyDeviceToHost, r.stream));
  }
  PX_CUDA(cudaStreamSynchronize(r.stream));  // also retires the pageable staging vector of owned lists
  if (meta[0] > 0x7fffffffll) return fail_pixie("command list too large");
  L.numEntries = meta[0];
  L.maxEntries = (int)std::max<long long>(2, meta[1]);
  L.scratchWords = L.maxEntries > L.smemCap ? L.maxEntries * kScratchArrays : 0;

  // Device block B: band entries + the per-warp spill scratch for bands with more entries than fit in smem.
  const size_t entriesBytes = al(std::max<size_t>(1, (size_t)L.numEntries) * sizeof(Entry));
  const size_t totalB = entriesBytes + al((size_t)L.rasterBlocks * L.warpsPerBlock * L.scratchWords * 4);
  if (arena) {
    void* blk;
    if (int rc = get_scratch(4, totalB, &blk)) return rc;
    L.blockB = (uint8_t*)blk;
  } else {
    PX_CUDA(cudaMalloc(&L.blockB, totalB));
  }
  L.entries = (Entry*)L.blockB;
  L.scratch = L.scratchWords ? (uint32_t*)(L.blockB + entriesBytes) : nullptr;
  if (g_trace)
    fprintf(stderr, "[pixie_cuda] build_list: host plan %.3f ms (bounds %.3f), stage + device count/scan + readback %.3f ms "
            "(%zu B staged, blocks %zu + %zu B, %lld entries, max %d per band)\n",
            t1 - t0, tBounds, now_ms() - t1, h2dBytes, totalA, totalB, (long long)L.numEntries, L.maxEntries);
  return 0;
}

static int run_list(CmdList& L, Image* im, uint64_t* covered_px, uint8_t* host_pixels = nullptr) {
  Runtime& r = rt();
  if (im->bpp != 4 || im->w != L.w || im->h != L.h || im->layers != L.layers)
    return fail_pixie("command list was built for a different canvas shape");
  if (L.numFills == 0) {
    if (covered_px) *covered_px = 0;
    return 0;
  }
  PX_CUDA(cudaMemsetAsync(L.counters, 0, 128, r.stream));  // row tickets + covered px
  if (L.numParts > 0) {
    const int warps = (int)std::min<int64_t>(L.numParts, (int64_t)r.num_sms * 32);
    const int blocks = (warps + 7) / 8;
    {
      ProfScope ps(kProfPartition);
      partition_kernel<<<blocks, 256, 0, r.stream>>>(L.fills, L.numFills, L.entryOff, L.segs, L.wind, L.entries,
                                                     L.flags, (int)L.numParts);
    }
    PX_LAUNCHED();
  }
  RasterArgs A;
  A.canvas = (px_t*)im->data;
  A.w = L.w; A.h = L.h; A.layers = L.layers;
  A.fills = L.fills; A.rowRange = L.rowRange; A.layerFillBegin = L.layerFillBegin; A.entryOff = L.entryOff; A.entries = L.entries;
  A.flags = L.flags; A.gscratch = L.scratch; A.counters = L.counters;
  A.smemCap = L.smemCap; A.scratchCap = L.maxEntries; A.covBytes = L.covBytes;
  A.countCovered = covered_px ? 1 : 0;
  static size_t configured = 0;
  if (L.smemBytes > 48 * 1024 && configured < L.smemBytes) {
    PX_CUDA(cudaFuncSetAttribute(raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemBytes));
    configured = L.smemBytes;
  }
  const long long totalRows = (long long)L.layers * L.h;
  if (!host_pixels) {
    A.rowBegin = 0; A.rowEnd = totalRows; A.ticketSlot = 0; A.scratchBlockBase = 0;
    {
      ProfScope ps(kProfRaster);
      raster_kernel<<<L.rasterBlocks, L.warpsPerBlock * 32, L.smemBytes, r.stream>>>(A);
    }
    PX_LAUNCHED();
  } else {
    // Row bands on concurrent streams: each band's pixels start their way to the host as soon as the
    // band is rasterised, while the heavier bands are still being worked on.
    constexpr int K = Runtime::kBands;
    if (!r.band_stream[0]) {
      for (int b = 0; b < K; b++) {
        PX_CUDA(cudaStreamCreateWithFlags(&r.band_stream[b], cudaStreamNonBlocking));
        PX_CUDA(cudaEventCreateWithFlags(&r.band_done[b], cudaEventDisableTiming));
      }
      PX_CUDA(cudaEventCreateWithFlags(&r.band_start, cudaEventDisableTiming));
    }
    PX_CUDA(cudaEventRecord(r.band_start, r.stream));
    const int gridB = std::max(1, L.rasterBlocks / K);
    const size_t rowBytes = (size_t)L.w * 4;
    for (int b = 0; b < K; b++) {
      A.rowBegin = totalRows * b / K;
      A.rowEnd = totalRows * (b + 1) / K;
      A.ticketSlot = 4 + b;
      A.scratchBlockBase = b * gridB;
      if (A.rowEnd <= A.rowBegin) continue;
      PX_CUDA(cudaStreamWaitEvent(r.band_stream[b], r.band_start, 0));
      raster_kernel<<<gridB, L.warpsPerBlock * 32, L.smemBytes, r.band_stream[b]>>>(A);
      PX_LAUNCHED();
      PX_CUDA(cudaMemcpyAsync(host_pixels + (size_t)A.rowBegin * rowBytes, im->data + (size_t)A.rowBegin * rowBytes,
                              (size_t)(A.rowEnd - A.rowBegin) * rowBytes, cudaMemcpyDeviceToHost, r.band_stream[b]));
      PX_CUDA(cudaEventRecord(r.band_done[b], r.band_stream[b]));
      PX_CUDA(cudaStreamWaitEvent(r.stream, r.band_done[b], 0));
    }
  }
  if (covered_px) {
    unsigned long long host[2];
    PX_CUDA(cudaMemcpyAsync(host, L.counters, 16, cudaMemcpyDeviceToHost, r.stream));
    PX_CUDA(cudaStreamSynchronize(r.stream));
    *covered_px = host[1];
  }
  return 0;
}

}  // namespace pixie

using namespace pixie;

extern "C" {

int pixie_cuda_cmdlist_create(int w, int h, int layers, int numFills, const int32_t* layerOf, const float* seg,
                              const int16_t* wind, const int32_t* segOff, const uint32_t* rgbx, const uint8_t* rule,
                              const uint8_t* mode, pixie_cmdlist_t* out) {
  if (int rc = ensure_init()) return rc;
  CmdList L;
  int rc = build_list(L, false, w, h, layers, numFills, layerOf, seg, wind, segOff, rgbx, rule, mode);
  if (rc) {
    cudaStreamSynchronize(rt().stream);
    free_list(L);
    return rc;
  }
  std::lock_guard<std::mutex> lk(rt().mu);
  const uint64_t hd = g_next_list++;
  g_lists[hd] = L;
  *out = hd;
  return 0;
}

int pixie_cuda_cmdlist_run(pixie_cmdlist_t list, pixie_image_t image, uint64_t* covered_px) {
  if (int rc = ensure_init()) return rc;
  auto it = g_lists.find(list);
  if (it == g_lists.end()) return fail_pixie("invalid command list handle");
  Image* im = find_image(image);
  if (!im) return 1;
  return run_list(it->second, im, covered_px);
}

int pixie_cuda_cmdlist_info(pixie_cmdlist_t list, int64_t* numSegs, int64_t* numParts, int64_t* numEntries,
                            int64_t* launches) {
  auto it = g_lists.find(list);
  if (it == g_lists.end()) return fail_pixie("invalid command list handle");
  if (numSegs) *numSegs = it->second.numSegs;
  if (numParts) *numParts = it->second.numParts;
  if (numEntries) *numEntries = it->second.numEntries;
  if (launches) *launches = it->second.numParts > 0 ? 2 : 1;
  return 0;
}

int pixie_cuda_cmdlist_destroy(pixie_cmdlist_t list) {
  auto it = g_lists.find(list);
  if (it == g_lists.end()) return fail_pixie("invalid command list handle");
  cudaStreamSynchronize(rt().stream);
  free_list(it->second);
  g_lists.erase(it);
  return 0;
}

int pixie_cuda_fill_batch(pixie_image_t image, int numFills, const int32_t* layerOf, const float* seg,
                          const int16_t* wind, const int32_t* segOff, const uint32_t* rgbx, const uint8_t* rule,
                          const uint8_t* mode, uint64_t* covered_px) {
  if (int rc = ensure_init()) return rc;
  Image* im = find_image(image);
  if (!im) return 1;
  if (im->bpp != 4) return fail_pixie("fill needs an RGBX image");
  CmdList L;  // lives in the library arena: no allocation, no synchronisation, nothing to free
  int rc = build_list(L, true, im->w, im->h, im->layers, numFills, layerOf, seg, wind, segOff, rgbx, rule, mode);
  if (!rc) rc = run_list(L, im, covered_px);
  return rc;
}

int pixie_cuda_render_batch_host(uint8_t* pixels, int width, int height, int clear, int numFills, const float* seg,
                                 const int16_t* wind, const int32_t* segOff, const uint32_t* rgbx, const uint8_t* rule,
                                 const uint8_t* mode, uint64_t* covered_px) {
  if (int rc = ensure_init()) return rc;
  if (width <= 0 || height <= 0) return fail_pixie("Image width and height must be > 0");
  Runtime& r = rt();
  void* canvas;
  const size_t bytes = (size_t)width * height * 4;
  if (int rc = get_scratch(5, bytes, &canvas)) return rc;
  if (clear) PX_CUDA(cudaMemsetAsync(canvas, 0, bytes, r.stream));                        // newImage(width, height)
  else PX_CUDA(cudaMemcpyAsync(canvas, pixels, bytes, cudaMemcpyHostToDevice, r.stream));  // draw over existing pixels
  Image im;
  im.data = (uint8_t*)canvas; im.w = width; im.h = height; im.layers = 1; im.bpp = 4; im.owned = false;
  CmdList L;
  int rc = build_list(L, true, width, height, 1, numFills, nullptr, seg, wind, segOff, rgbx, rule, mode);
  if (rc) return rc;
  if (numFills == 0) {
    PX_CUDA(cudaMemcpyAsync(pixels, canvas, bytes, cudaMemcpyDeviceToHost, r.stream));
  } else {
    rc = run_list(L, &im, covered_px, pixels);
    if (rc) return rc;
  }
  PX_CUDA(cudaStreamSynchronize(r.stream));
  return 0;
}

int pixie_cuda_fill_segments(pixie_image_t image, const float* seg, const int16_t* wind, int n, uint32_t rgbx,
                             int rule, int mode) {
  if (rule < 0 || rule > 1 || mode < 0 || mode >= NumBlendModes) return fail_pixie("invalid blend mode / winding rule");
  const int32_t segOff[2] = {0, n};
  const uint8_t r8 = (uint8_t)rule, m8 = (uint8_t)mode;
  return pixie_cuda_fill_batch(image, 1, nullptr, seg, wind, segOff, &rgbx, &r8, &m8, nullptr);
}

}  // extern "C"

// K1-K3 — the path rasteriser behind fillPath / strokePath:
//   fillShapes (treeform/pixie src/pixie/paths.nim:1593-1912) from the segment list down.
//
//   K1  partition_kernel   partitionSegments (:1168-1262): order-preserving binning of segments
//                          into equal-height y bands (one warp per band, ballot compaction keeps
//                          segment order), per-band requiresAntiAliasing (:1149-1166), clipping of
//                          spanning entries (:1236-1248), twoNonintersectingSpanningSegments (:1250-1262)
//   K2  plan_light_kernel  the scanline loop (:1631-1908) up to the canvas: every (fill, scanline) job picks
//       plan_kernel        mode A (pixel-aligned pair :1644-1668), mode B (exact-area trapezoids :1691-1872) or
//                          mode C (computeCoverage :1350-1431, `walk` :1298-1330 emulated literally) and stores
//                          a small plan; one thread per job for bands with <= 16 entries, one warp per job (5
//                          sample lines side by side) for crowded bands, the two kernels side by side
//   K3  raster_kernel      one warp per canvas row applies the plans of the row's fills in order, so fills
//                          of one canvas keep the reference's sequential semantics with a single launch:
//                          coverage accumulation in shared memory, fillCoverage / fillHits (:1479-1591),
//                          trapezoid edge pixels, blends.nim
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

struct JobHdr;
struct CmdList {
  int w = 0, h = 0, layers = 1;
  int numFills = 0;
  int64_t numSegs = 0, numParts = 0, numEntries = 0;
  int maxEntries = 0;
  float4* segs = nullptr;
  int16_t* wind = nullptr;
  FillHeader* fills = nullptr;
  int2* rowRange = nullptr;
  int* fillJobBase = nullptr;
  unsigned* payOff = nullptr;
  JobHdr* jobs = nullptr;
  uint2* payload = nullptr;
  int totalJobs = 0, planBlocks = 0;
  // row bands of pixie_cuda_render_batch_host: band b plans + rasterises rows [h*b/bands, h*(b+1)/bands)
  int bands = 1;
  int* bandJobBase = nullptr;              // [bands][numFills + 1] job prefix of each band
  int bandJobs[Runtime::kBands] = {};
  int* heavyList = nullptr;                // [totalJobs] jobs left for plan_kernel, one region per plan launch
  int* splitArrive = nullptr;              // [totalJobs] see RasterArgs
  int* monsterList = nullptr;              // [totalJobs] see RasterArgs
  unsigned* scratchSlots = nullptr;        // bitmap: spill-scratch block slots in use
  int scratchSlotCount = 0;
  size_t planSmem = 0;
  int* entryOff = nullptr;
  int* layerFillBegin = nullptr;
  Entry* entries = nullptr;
  uint8_t* flags = nullptr;
  uint32_t* ranges = nullptr;  // per segment: first | last << 16 band it touches (count_kernel)
  uint32_t* groupRange = nullptr;  // the same per aligned group of 32 segments
  uint32_t* scratch = nullptr;  // per-warp spill area when a band has more entries than fit in smem
  unsigned long long* counters = nullptr;  // [0] row ticket, [1] covered px
  bool clearFirst = false;  // set around one run: the raster kernel clears the canvas as it goes
  int subShift = 0;         // see RasterArgs
  int rasterBlocks = 0, warpsPerBlock = 0, scratchWords = 0, covBytes = 0, smemCap = 0, tileW = 0, tiles = 1;
  size_t smemBytes = 0, h2dBytes = 0;
  uint8_t* block = nullptr;   // block A: host-written inputs + device-made offsets/flags/counters
  uint8_t* blockB = nullptr;  // block B: band entries + spill scratch (sized after the device-side count)
  bool owned = false;
  bool serial = false;         // run a banded list as one band (pixie_cuda_cmdlist_set_overlap(list, 0))
  bool deviceCounted = false;  // band counts / offsets were made by count_kernel + scans (lists above 8192 segments)
  bool countFresh = false;     // ... and the arrays still hold the offsets of build_list's own pass
  // host copy of what a row-band run needs to lay out its jobs (pixie_cuda_cmdlist_run_rows)
  std::vector<int> hostStartY, hostPathHeight;  // per fill; pathHeight <= startY when the fill has no jobs
  int maxWrapRows = 0;                          // MaskBlend fills reaching left of the canvas read jobs of later rows
  int* rowsJobBase = nullptr;                   // device [numFills + 1], made per run_rows call
  int* bandRows = nullptr;                      // device [numParts]: rows of every band (static)
  int* rowOrder = nullptr;                      // device [h]: rows by descending work estimate, null: row order
  unsigned long long* chunkSums = nullptr;      // device [2 * chunks]: scratch of the band scans
};

static std::unordered_map<uint64_t, CmdList> g_lists;
static uint64_t g_next_list = 1;

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
PXD long long f2ll(float f) { return (long long)f; }          // Nim float32 -> int
PXD int f2i_sat(float f) { return __float2int_rz(f); }         // the same, saturated to int32 (NaN -> 0 in both)
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

// K1a: entries per band (one thread per segment; order is irrelevant for counting).  The band range of
// every segment (partitionRange :1201-1213) is kept as `atP | toP << 16` for partition_kernel, whose band
// warps then scan 4 bytes per segment instead of the segment itself; kNoBand marks segments that belong to
// no band (inactive fills) and kWideBands fills with more than 65535 bands (recomputed from the segment).
constexpr uint32_t kNoBand = 0x0000FFFFu;  // atP = 65535 > toP = 0
constexpr int kMaxPackedBands = 65535;
PXD void band_range(const FillHeader& H, float ay, float by, unsigned& atP, unsigned& toP) {
  const float startYf = (float)(unsigned)H.startY;
  const unsigned ph = (unsigned)H.partitionHeight, lastP = (unsigned)(H.numPartitions - 1);
  atP = min(__float2uint_rz(fmaxf(0.0f, ay - startYf)) / ph, lastP);
  toP = min(__float2uint_rz(fmaxf(0.0f, by - startYf)) / ph, lastP);
}
__global__ void __launch_bounds__(256) count_kernel(const FillHeader* __restrict__ fills, int numFills,
                                                    const float4* __restrict__ segs, int numSegs, int* __restrict__ cnt,
                                                    uint32_t* __restrict__ ranges, uint32_t* __restrict__ groupRange) {
  // a warp = 32 consecutive segments, aligned to 32: besides the per-segment band range it leaves one summary word
  // for the group (lowest first band | highest last band << 16), so that partition_kernel's band warps skip
  // groups that cannot touch their band — consecutive segments of a path are neighbours in space
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < numSegs; base += gridDim.x * blockDim.x) {
    const int i = base + (threadIdx.x & 31);
    unsigned atP = 0xFFFFu, toP = 0u;  // kNoBand
    int f = -1;
    bool plain = false;  // a segment of an active, packed fill: its range can go into the summary
    if (i < numSegs) {
      f = find_fill<false>(fills, numFills, i);
      while (f > 0 && fills[f].segCount == 0) f--;  // empty fills share their segBegin with the next one
      const FillHeader H = fills[f];
      if (!H.active || H.numPartitions <= 0 || i >= H.segBegin + H.segCount) {
        ranges[i] = kNoBand;
      } else if (H.numPartitions == 1) {
        atomicAdd(&cnt[H.partBase], 1);
        ranges[i] = 0u;
        atP = toP = 0u;
        plain = true;
      } else {
        const float4 s = segs[i];
        band_range(H, s.y, s.w, atP, toP);
        plain = H.numPartitions <= kMaxPackedBands;
        ranges[i] = plain ? (atP | (toP << 16)) : kNoBand;
        for (unsigned p = atP; p <= toP; p++) atomicAdd(&cnt[H.partBase + (int)p], 1);
      }
    }
    const int f0 = __shfl_sync(0xffffffffu, f, 0);
    const bool uniform = __all_sync(0xffffffffu, plain && f == f0);
    unsigned lo = atP, hi = toP;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    // groups that mix fills (or hold segments without a packed range) are always scanned
    if ((threadIdx.x & 31) == 0) groupRange[base >> 5] = uniform ? (lo | (hi << 16)) : 0xFFFF0000u;
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
                                                        const int16_t* __restrict__ wind,
                                                        const uint32_t* __restrict__ ranges,
                                                        const uint32_t* __restrict__ groupRange, Entry* __restrict__ entries,
                                                        uint8_t* __restrict__ flags, int numParts) {
  const int lane = threadIdx.x & 31;
  const int warpsTotal = (gridDim.x * blockDim.x) >> 5;
  for (int gp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gp < numParts; gp += warpsTotal) {
    const FillHeader H = fills[find_fill<true>(fills, numFills, gp)];
    const int p = gp - H.partBase;
    const int top = H.startY + p * H.partitionHeight;
    const int bottom = (p == H.numPartitions - 1) ? H.pathHeight : top + H.partitionHeight;
    const float topf = (float)top, botf = (float)bottom;
    const bool packed = H.numPartitions <= kMaxPackedBands;
    const int outBase = entryOff[gp];
    int out = outBase;
    bool aa = false;
    // The fill's segments in groups of 32 aligned to the global segment index: 32 group summaries per step decide
    // which groups can touch this band; only those are scanned (in order: groups ascending, lanes ascending).
    const int gBegin = H.segBegin >> 5, gEnd = (H.segBegin + H.segCount + 31) >> 5;
    const unsigned pk = (unsigned)p & 0xFFFFu;
    for (int g0 = gBegin; g0 < gEnd; g0 += 32) {
      bool need = false;
      if (g0 + lane < gEnd) {
        const uint32_t gr = packed ? groupRange[g0 + lane] : 0xFFFF0000u;
        need = pk >= (gr & 0xFFFFu) && pk <= (gr >> 16);
      }
      unsigned todo = __ballot_sync(0xffffffffu, need);
      while (todo) {
        const int g = g0 + __ffs(todo) - 1;
        todo &= todo - 1;
        const int gi = g * 32 + lane;          // global segment index
        const int i = gi - H.segBegin;         // index within the fill
        bool touches = false;
        if (i >= 0 && i < H.segCount) {
          if (packed) {
            const uint32_t rv = ranges[gi];
            touches = pk >= (rv & 0xFFFFu) && pk <= (rv >> 16);
          } else {  // more bands than the packed form holds: partitionRange from the segment itself
            const float4 s = segs[gi];
            unsigned atP, toP;
            band_range(H, s.y, s.w, atP, toP);
            touches = (unsigned)p >= atP && (unsigned)p <= toP;
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, touches);
        if (touches) {
          const float4 s = segs[gi];
          Entry e;  // initPartitionEntry (:1127-1135)
          e.ax = s.x; e.ay = s.y; e.bx = s.z; e.by = s.w;
          e.winding = (int)wind[gi];
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
// K2 + K3: scanlines.
//
// A *job* is one fill on one scanline inside the path's rows.  What a job does splits cleanly:
//   plan_kernel   (K2) everything that depends on the geometry only — the scanline loop of fillShapes
//                 (:1631-1908) up to the point where it touches the canvas: entry selection, the
//                 trapezoid shortcut's sort and checks, computeCoverage's hits / sort / walk.  Jobs are
//                 independent, so all (fill, scanline) pairs of a command list run in parallel; the
//                 result is a 16-byte JobHdr plus a short payload (sorted edges or spans).
//   raster_kernel (K3) one warp per canvas row walks the row's fills IN ORDER and applies their plans:
//                 trapezoid edge pixels and interiors, the coverage row (accumulated in shared memory
//                 from the spans) and its blend, integer spans, MaskBlend's clears.
// Two kernels instead of one fused one because of the instruction cache: the fused kernel with every
// primitive inlined was 0.6 MB of SASS, 24 warps per SM all in different places of it, and ncu showed
// `no_instruction` as the top stall reason (icc hit rate 70-79 %).  Each half now fits the 32 KB L1.5
// I-cache, and the expensive half no longer sits on the critical path of the heaviest row.
// ---------------------------------------------------------------------------------------------
struct JobHdr {
  int kind;    // PlanKind
  int n;       // Trapezoids: edges (even); Coverage / Spans: spans
  int pa, pb;  // Aligned: the span [pa, pb); the other kinds: the pixels [pa, pb) the plan can touch (its edges / spans),
               // so that a raster warp whose columns they miss skips the job (MaskBlend fills clear and are never skipped)
};
PXD int extent_lo(int lo, float xa, float xb) { return min(lo, __float2int_rz(fminf(xa, xb))); }          // trapezoid edge pixels
PXD int extent_hi(int hi, float xa, float xb) { return max(hi, __float2int_rz(ceilf(fmaxf(xa, xb)))); }  // [trunc(min x), ceil(max x))

struct RasterArgs {
  px_t* canvas;
  int w, h, layers;
  int numFills;
  const FillHeader* fills;
  const int2* rowRange;        // rows [x, y) of the canvas each fill has to visit (empty when inactive)
  const int* layerFillBegin;
  const int* entryOff;
  const Entry* entries;
  const uint8_t* flags;
  const int* fillJobBase;      // [numFills + 1] first job of each fill; job = fillJobBase[f] + (y - startY)
  const int* planJobBase;      // [numFills + 1] the same restricted to the rows [planY0, planY1) of this plan launch
  int planY0, planJobs;        // jobs of this plan launch (whole list: planJobBase = fillJobBase, planY0 = 0)
  int* monsterList;            // jobs whose scanline is planned by one warp per sample line (plan_classify_kernel), count in heavyCount[24]
  int* splitArrive;            // [totalJobs], zero between runs: sample lines of a split job that have been planned (atomicInc wraps it back to 0)
  int* heavyList;              // jobs of crowded bands (> kLightMax entries), compacted by plan_light_kernel for
  unsigned long long* heavyCount;  // plan_kernel; one list region + counter per plan launch
  unsigned* scratchSlots;      // bitmap of the spill-scratch block slots in use (plan launches of several row
  int scratchSlotCount;        // bands run concurrently and share one pool sized for the resident blocks)
  const unsigned* payOff;      // [numParts + 1] payload offset of each band, in entry-rows
  JobHdr* jobs;
  uint2* payload;              // kPaySlots 8-byte slots per entry-row
  int totalJobs;
  uint32_t* gscratch;          // per-warp global spill (scratchCap * kScratchArrays words each), may be null
  unsigned long long* counters;
  long long rowBegin, rowEnd;  // flattened (layer * h + y) rows this launch owns
  int ticketSlot;              // counters[ticketSlot] hands out rows
  int clearFirst;              // every ticket zeroes its row tile before the first fill (the canvas clear rides along)
  int subShift;                // rowOrder lists: a tile is handed out in 1 << subShift pieces; rowOrder[i] >> 28 = how many of them the row uses (as a shift)
  int smemCap;                 // entries whose scratch fits in shared memory
  int scratchCap;              // capacity of the global spill
  const int* rowOrder;         // [h] rows by descending work estimate (build_list): the raster kernel's ticket order
  int tileW, tiles;            // a canvas row is rasterised in `tiles` pieces of tileW columns, one warp each
  int covBytes;                // bytes of the per-warp coverage row in shared memory
  int countCovered;
};

constexpr int kScratchArrays = 18;  // scratch words per band entry (see the layout in plan_row)
constexpr int kPaySlots = 5;        // payload slots per entry-row: <= 5 spans (one per sample line) or 2 per edge

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

enum PlanKind {
  PlanOutside = -1,    // the path does not reach this row (only MaskBlend touches it, :1910-1912)
  PlanNothing = 0,
  PlanAligned = 1,     // mode A (:1644-1668): one pixel-aligned span [pa, pb)
  PlanTrapezoids = 2,  // mode B (:1691-1872): n sorted edges, exact-area edge pixels + interiors
  PlanCoverage = 3,    // computeCoverage with anti-aliasing: n spans (24.8 fixed) of the 5 sample lines
  PlanSpans = 4,       // computeCoverage without: n spans of the single sample line (fillHits over walkInteger)
};

// walk (:1298-1330) executed by one lane over sorted hits; spans are appended to spanA/spanB, which may be
// the hit arrays themselves: span k is emitted while reading hit i > k.
PXD int walk_spans(const int* hitAt, const int* hitW, int numHits, int rule, int* spanA, int* spanB) {
  // The hits of crowded scanlines live in HBM scratch and the state machine below consumes them one by one: reading
  // them 8 (+1 of lookahead) at a time keeps 18 loads in flight instead of one dependent load per step.
  int count = 0, prevAt = 0, ns = 0;
  bool skipNext = false;  // the previous hit cancelled this one (`i += 2`, :1306-1308)
#pragma unroll 1
  for (int b = 0; b < numHits; b += 8) {
    int ca[9], cw[9];
#pragma unroll
    for (int u = 0; u < 9; u++) {
      const bool in = b + u < numHits;
      ca[u] = in ? hitAt[b + u] : 0;
      cw[u] = in ? hitW[b + u] : 0;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int i = b + u;
      if (i < numHits) {
        if (skipNext) {
          skipNext = false;
        } else {
          const int at = ca[u], winding = cw[u];
          bool consumed = false;  // `continue` of the reference's loop: count already updated or hit skipped
          if (at > 0) {
            if (should_fill(rule, count)) {
              bool emit = true;
              if (i < numHits - 1) {
                const int nextAt = ca[u + 1], nextWinding = cw[u + 1];
                if (nextAt == at && winding + nextWinding == 0) {
                  skipNext = true;
                  consumed = true;
                  emit = false;
                } else if (rule == 0 && count + winding != 0) {
                  count += winding;
                  consumed = true;
                  emit = false;
                }
              }
              if (emit) {
                spanA[ns] = prevAt;
                spanB[ns] = at;
                ns++;
              }
            }
            if (!consumed) prevAt = at;
          }
          if (!consumed) count += winding;
        }
      }
    }
  }
  return ns;
}

// Warp-wide bitonic sort of P (a power of two) 64-bit keys held as two word arrays, ascending.  The crowded
// bands' stable sorts use it with key = {sortable value, original position}: unique keys, so the result is the
// stable order, in O(P log^2 P / 32) steps per lane instead of the quadratic rank count.
template <int U>  // U = P / 64 compare-exchanges per lane and step
PXD void warp_bitonic_sort_u(uint2* key, int lane) {  // key = {low word, high word}
  constexpr int P = 64 * U;
  constexpr int B = U < 4 ? U : 4;  // pairs in flight per lane
#pragma unroll 1
  for (int k = 2; k <= P; k <<= 1) {
#pragma unroll 1
    for (int j = k >> 1; j > 0; j >>= 1) {
      // pair t: i = t with a zero inserted at bit j, partner i | j; the loads of a batch first (they are
      // independent, but the compiler cannot tell that from the stores), then the exchanges
#pragma unroll 1
      for (int u0 = 0; u0 < U; u0 += B) {
        uint2 a[B], b[B];
        int ii[B];
#pragma unroll
        for (int u = 0; u < B; u++) {
          const int t = lane + 32 * (u0 + u);
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          ii[u] = i;
          a[u] = key[i];
          b[u] = key[i | j];
        }
#pragma unroll
        for (int u = 0; u < B; u++) {
          const int i = ii[u];
          const bool aGreater = a[u].y > b[u].y || (a[u].y == b[u].y && a[u].x > b[u].x);
          if (aGreater == ((i & k) == 0)) {
            key[i] = b[u];
            key[i | j] = a[u];
          }
        }
      }
      __syncwarp();
    }
  }
}
PXD void warp_bitonic_sort(uint2* key, int P, int lane) {  // P = 128, 256 or 512
  if (P == 128) warp_bitonic_sort_u<2>(key, lane);
  else if (P == 256) warp_bitonic_sort_u<4>(key, lane);
  else warp_bitonic_sort_u<8>(key, lane);
}
PXD uint32_t sortable_float(float f) {  // order-preserving map to uint32; -0 and +0 compare equal in the reference
  if (f == 0.0f) return 0x80000000u;
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
constexpr int kBitonicMin = 64;   // above this many selected entries the sorts go through warp_bitonic_sort
constexpr int kBitonicMax = 512;  // 2 x 512 words of keys fit a warp's shared-memory scratch (smemCap * kScratchArrays)

// last fill whose first job is <= j (fills without jobs share their base with the next one)
PXD int find_fill_by_job(const int* __restrict__ jobBase, int numFills, int j) {
  int lo = 0, hi = numFills;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (jobBase[mid] <= j) lo = mid + 1;
    else hi = mid;
  }
  return lo - 1;
}

// ---------------------------------------------------------------------------------------------
// K2a: plan_light_kernel — one THREAD per job for bands with at most kLightMax entries (nine jobs out of ten in
// the tiger: a handful of edges cross a scanline).  A warp per job leaves most lanes idle on those and pays
// several hundred warp instructions of fixed overhead each; here a lane runs the reference's sequential code for
// its own scanline (selection, the trapezoid checks with their insertion sort, hits / sortHits / walk per
// sample line), 32 neighbouring scanlines of a path per warp, which mostly take the same branches.  Jobs of
// crowded bands are appended to a list for plan_kernel (K2b), where the quadratic sorts get a whole warp.
// Per-thread scratch: kLightArrays arrays of kLightMax words in shared memory, interleaved by thread.
// ---------------------------------------------------------------------------------------------
// 12: swept on the tiger and on the icon batch (4 / 8 / 12 / 16 -> plan 0.150 / 0.133 / 0.124 / 0.147 ms on the tiger);
// below that too many small jobs pay a whole warp, above it the longest thread-per-job chains set the kernel's duration
#ifndef PIXIE_LIGHT_MAX
#define PIXIE_LIGHT_MAX 12
#endif
#ifndef PIXIE_SPLIT_MIN
#define PIXIE_SPLIT_MIN 128
#endif
constexpr int kLightMax = PIXIE_LIGHT_MAX;
constexpr int kVeryHeavy = 64;  // bands with more entries than this are planned first
constexpr int kSplitLines = PIXIE_SPLIT_MIN;  // ... and above this many, one warp per SAMPLE LINE plans the scanline (plan_kernel)
constexpr int kLightArrays = 5;
constexpr int kLightThreads = 128;

// which jobs are heavy: one thread per job, before the two plan kernels (which then run side by side)
__global__ void __launch_bounds__(256) plan_classify_kernel(const RasterArgs A) {
  const int bj = blockIdx.x * blockDim.x + threadIdx.x;
  if (bj >= A.planJobs) return;
  const int f = find_fill_by_job(A.planJobBase, A.numFills, bj);
  const FillHeader* Hp = A.fills + f;
  const int startY = Hp->startY;
  const int y = max(startY, A.planY0) + (bj - A.planJobBase[f]);
  int p = (int)((unsigned)(y - startY) / (unsigned)Hp->partitionHeight);
  if (p > Hp->numPartitions - 1) p = Hp->numPartitions - 1;
  const int gp = Hp->partBase + p;
  const int eCnt = A.entryOff[gp + 1] - A.entryOff[gp];
  // The list is filled from both ends: jobs of very crowded bands (their sorts and walks are the longest single
  // pieces of work of the whole launch) from the front, so that plan_kernel starts them first, the rest from the
  // back.  One atomic per warp and class: the jobs of a warp take consecutive slots.
  const bool mon = eCnt > kSplitLines && (A.flags[gp] & 3u) == 1u;  // anti-aliased, not the two-spanning-segments case
  const bool heavy = eCnt > kLightMax && !mon, very = eCnt > kVeryHeavy;
  const int lane = threadIdx.x & 31;
  const unsigned act = __activemask();
  const unsigned balV = __ballot_sync(act, heavy && very), balH = __ballot_sync(act, heavy && !very);
  const unsigned balM = __ballot_sync(act, mon);
  if (mon) {
    const int leader = __ffs(balM) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(A.heavyCount + 24, (unsigned long long)__popc(balM));
    base = __shfl_sync(balM, base, leader);
    A.monsterList[base + __popc(balM & ((1u << lane) - 1u))] = bj;
  } else if (heavy && very) {
    const int leader = __ffs(balV) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(A.heavyCount, (unsigned long long)__popc(balV));
    base = __shfl_sync(balV, base, leader);
    A.heavyList[base + __popc(balV & ((1u << lane) - 1u))] = bj;
  } else if (heavy) {
    const int leader = __ffs(balH) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(A.heavyCount + 8, (unsigned long long)__popc(balH));
    base = __shfl_sync(balH, base, leader);
    A.heavyList[A.planJobs - 1 - (int)(base + __popc(balH & ((1u << lane) - 1u)))] = bj;
  }
}

PXD unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void plan_light_impl(const RasterArgs& A);
__global__ void __launch_bounds__(kLightThreads) plan_light_kernel(const RasterArgs A) {
#ifdef PIXIE_RASTER_TIMING
  if ((threadIdx.x & 31) == 0) atomicMax(&A.counters[52], ~gtime());
#endif
  plan_light_impl(A);
#ifdef PIXIE_RASTER_TIMING
  if ((threadIdx.x & 31) == 0) atomicMax(&A.counters[53], gtime());
#endif
}
__device__ __forceinline__ void plan_light_impl(const RasterArgs& A) {
  __shared__ uint32_t lsm[kLightArrays * kLightMax * kLightThreads];
  const int tid = threadIdx.x;
  auto at = [&](int arr, int i) -> uint32_t& { return lsm[(arr * kLightMax + i) * kLightThreads + tid]; };
  auto atf = [&](int arr, int i) -> float& { return reinterpret_cast<float*>(lsm)[(arr * kLightMax + i) * kLightThreads + tid]; };
  const int W = A.w;
  const float wf = (float)W;
  const int bj = blockIdx.x * kLightThreads + tid;
  if (bj >= A.planJobs) return;
  const int f = find_fill_by_job(A.planJobBase, A.numFills, bj);
  const FillHeader* Hp = A.fills + f;
  const int startY = Hp->startY, rule = Hp->rule;
  const int y = max(startY, A.planY0) + (bj - A.planJobBase[f]);
  const int job = A.fillJobBase[f] + (y - startY);
  const int ph = Hp->partitionHeight;
  int p = (int)((unsigned)(y - startY) / (unsigned)ph);
  if (p > Hp->numPartitions - 1) p = Hp->numPartitions - 1;
  const int gp = Hp->partBase + p;
  const int eBeg = A.entryOff[gp];
  const int eCnt = A.entryOff[gp + 1] - eBeg;
  if (eCnt > kLightMax) return;  // plan_kernel's (plan_classify_kernel has listed it)
  const unsigned fl = A.flags[gp];
  const Entry* ent = A.entries + eBeg;
  uint2* pay = A.payload + ((size_t)A.payOff[gp] + (size_t)(y - (startY + p * ph)) * (size_t)eCnt) * kPaySlots;
  const bool aa = (fl & 1u) != 0, two = (fl & 2u) != 0;
  JobHdr hdr;
  hdr.kind = PlanNothing;
  hdr.n = 0;
  hdr.pa = 0;
  hdr.pb = 0;
  if (two && !aa) {  // mode A (:1644-1668)
    hdr.kind = PlanAligned;
    hdr.pa = clampi(f2ll(ent[0].ax), 0, W);
    hdr.pb = clampi(f2ll(ent[1].ax), 0, W);
    A.jobs[job] = hdr;
    return;
  }
  // arrays: 0 sel (entry indices, in the order computeCoverage will see them)
  //         mode B: 1 tax, 2 tbx, 3 mid, 4 order        coverage: 1 hit x, 2 hit winding
  const float scanTop = (float)y, scanBottom = (float)(y + 1);
  bool allSpan = true;
  int nsel = 0;
  if (two) {
    nsel = 2;
    at(0, 0) = 0u;
    at(0, 1) = 1u;
  } else {  // :1681-1689
#pragma unroll 1
    for (int i = 0; i < eCnt; i++) {
      const float ay = ent[i].ay, by = ent[i].by;
      if (!(by <= scanTop || ay >= scanBottom)) {
        if (ay > scanTop || by < scanBottom) allSpan = false;
        at(0, nsel++) = (uint32_t)i;
      }
    }
  }
  if (allSpan && (nsel % 2) == 0) {  // mode B (:1691-1872)
#pragma unroll 1
    for (int s = 0; s < nsel; s++) {
      const Entry* e = ent + at(0, s);
      const float em = e->m, eb = e->b;
      const float xa = solve_x(em, eb, scanTop), xb = solve_x(em, eb, scanBottom);
      atf(1, s) = xa;
      atf(2, s) = xb;
      atf(3, s) = (xa + xb) * 0.5f;
    }
    // insertion sort of the positions by mid x (:1707-1716), stable
#pragma unroll 1
    for (int i = 0; i < nsel; i++) {
      const float ki = atf(3, i);
      int j = i - 1;
      while (j >= 0 && atf(3, (int)at(4, j)) > ki) {
        at(4, j + 1) = at(4, j);
        j--;
      }
      at(4, j + 1) = (uint32_t)i;
    }
    bool ok = true;
#pragma unroll 1
    for (int i = 0; i + 1 < nsel; i++) {  // partial-coverage areas must not overlap (:1720-1728)
      const int l = (int)at(4, i), r = (int)at(4, i + 1);
      const float leftMaxX = fmaxf(atf(1, l), atf(2, l)), rightMinX = fminf(atf(1, r), atf(2, r));
      if (f2ll(ceilf(leftMaxX)) > f2ll(rightMinX)) ok = false;
    }
    if (ok) {  // only simple fill pairs (:1732-1744)
      int pre = 0;
#pragma unroll 1
      for (int i = 0; i < nsel; i++) {
        pre += ent[at(0, (int)at(4, i))].winding;
        const bool fillIt = should_fill(rule, pre);
        if (((i & 1) == 0) ? !fillIt : fillIt) ok = false;
      }
    }
    if (ok) {
      hdr.kind = PlanTrapezoids;
      hdr.n = nsel;
      int lo = INT_MAX, hi = INT_MIN;
#pragma unroll 1
      for (int i = 0; i < nsel; i++) {
        const int es = (int)at(4, i);
        const Entry* e = ent + at(0, es);
        pay[2 * i] = make_uint2(__float_as_uint(e->m), __float_as_uint(e->b));
        pay[2 * i + 1] = make_uint2(__float_as_uint(atf(1, es)), __float_as_uint(atf(2, es)));
        lo = extent_lo(lo, atf(1, es), atf(2, es));
        hi = extent_hi(hi, atf(1, es), atf(2, es));
      }
      hdr.pa = lo;
      hdr.pb = hi;
      A.jobs[job] = hdr;
      return;
    }
    // the reference has sorted entryIndices in place (:1707-1716): computeCoverage sees them in mid-x order
#pragma unroll 1
    for (int i = 0; i < nsel; i++) at(1, i) = at(0, (int)at(4, i));
#pragma unroll 1
    for (int i = 0; i < nsel; i++) at(0, i) = at(1, i);
  }

  // mode C: computeCoverage (:1350-1431), sample line after sample line
  const int quality = aa ? 5 : 1;
  const float offset = 1.0f / (float)quality;
  const float initialOffset = offset / 2.0f + (float)(0.0001 * 3.141592653589793238462643383279502884);
  hdr.kind = aa ? PlanCoverage : PlanSpans;
  int S = 0, extLo = INT_MAX, extHi = INT_MIN;
  if (nsel > 0) {
    float yLine = (float)y + initialOffset - offset;
#pragma unroll 1
    for (int m = 0; m < quality; m++) {
      yLine += offset;
      int nh = 0;
#pragma unroll 1
      for (int s = 0; s < nsel; s++) {  // hits in entry order, inserted by x: the stable insertion sort of :1277-1286
        const Entry* e = ent + at(0, s);
        if (e->ay <= yLine && e->by >= yLine) {
          const float em = e->m, eb = e->b;
          float x = em == 0.0f ? eb : (yLine - eb) / em;
          x = (x != x) ? wf : (x < wf ? x : wf);  // min(x, width.float32)
          const int hx = fixed32(x);
          int j = nh - 1;
          while (j >= 0 && (int)at(1, j) > hx) {
            at(1, j + 1) = at(1, j);
            at(2, j + 1) = at(2, j);
            j--;
          }
          at(1, j + 1) = (uint32_t)hx;
          at(2, j + 1) = (uint32_t)e->winding;
          nh++;
        }
      }
      // walk (:1298-1330)
      int i = 0, count = 0, prevAt = 0;
#pragma unroll 1
      while (i < nh) {
        const int hat = (int)at(1, i), winding = (int)at(2, i);
        if (hat > 0) {
          if (should_fill(rule, count)) {
            if (i < nh - 1) {
              const int nextAt = (int)at(1, i + 1), nextWinding = (int)at(2, i + 1);
              if (nextAt == hat && winding + nextWinding == 0) {
                i += 2;
                continue;
              }
              if (rule == 0 && count + winding != 0) {
                count += winding;
                i++;
                continue;
              }
            }
            pay[S++] = make_uint2((uint32_t)prevAt, (uint32_t)hat);
            extLo = min(extLo, fx_integer(prevAt));
            extHi = max(extHi, fx_integer(hat) + 1);
          }
          prevAt = hat;
        }
        count += winding;
        i++;
      }
    }
  }
  hdr.n = S;
  hdr.pa = extLo;
  hdr.pb = extHi;
  A.jobs[job] = hdr;
}

// ---------------------------------------------------------------------------------------------
// K2: plan_kernel — one warp per job.
//
// Scratch per warp (cap = entries the scratch was sized for, kScratchArrays * cap words):
//   [0, cap) sel      indices of the band entries that cross the scanline, in entry order (:1681-1689)
//   [cap, 2cap) sel2  the same in trapezoid mid-x order (what computeCoverage sees after a failed mode B)
//   mode B:   tax, tbx, mid, wnd, -, order  at 2cap .. 8cap
//   coverage: uAt[T], sAt[T], sW[T], lineHits[8] at 2cap ..   with T = quality * nsel <= 5 cap
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) plan_kernel(const RasterArgs A) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
#ifdef PIXIE_RASTER_TIMING
  if ((threadIdx.x & 31) == 0) atomicMax(&A.counters[50], ~gtime());
#endif
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warpsPerBlock = blockDim.x >> 5;
  uint32_t* sscr = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warp * A.smemCap * kScratchArrays;
  // Spill scratch for bands with more entries than fit in shared memory: a block claims one slot of a pool
  // that has as many slots as plan_kernel blocks can be resident on the GPU, whatever launch they belong to.
  __shared__ int s_slot;
  uint32_t* gscr = nullptr;
  if (A.gscratch) {
    if (threadIdx.x == 0) {
      const int words = (A.scratchSlotCount + 31) >> 5;
      int wi = blockIdx.x % words, bit = (blockIdx.x / words) & 31, slot = -1;
      while (slot < 0) {
        const int cand = wi * 32 + bit;
        if (cand < A.scratchSlotCount && !(atomicOr(A.scratchSlots + wi, 1u << bit) & (1u << bit))) slot = cand;
        else if (++bit == 32) { bit = 0; wi = (wi + 1) % words; }
      }
      s_slot = slot;
    }
    __syncthreads();
    gscr = A.gscratch + (size_t)(s_slot * warpsPerBlock + warp) * ((size_t)A.scratchCap * kScratchArrays);
  }
  const int W = A.w;
  const float wf = (float)W;
  const int warpsTotal = gridDim.x * warpsPerBlock;
  // jobs plan_classify_kernel left for a whole warp: the very crowded ones from the front of the list, then the rest
  const int numFront = (int)A.heavyCount[0], numHeavy = numFront + (int)A.heavyCount[8], numMon = (int)A.heavyCount[24];
  // Splitting pays when a few such scanlines would otherwise outlast the rest of the launch (the tiger: 362 of 93 k jobs);
  // when there are thousands of them (an icon batch: 6 400) the launch is bound by its total work and they are planned
  // whole like the others, first.
  const bool doSplit = 5 * numMon <= 2 * warpsTotal;
  const int monTickets = doSplit ? 5 * numMon : numMon;
  // Jobs are handed out one at a time from a counter (heavyCount[16] = counters[32 + launch]): they differ tenfold in
  // cost and the list starts with the most crowded ones, so whichever warp is free takes the next — a fixed stride left
  // most warps idle while a few finished their share.
#pragma unroll 1
  while (true) {
    int hj = 0;
    if (lane == 0) hj = (int)atomicAdd(A.heavyCount + 16, 1ull);
    hj = __shfl_sync(0xffffffffu, hj, 0);
    // The scanlines plan_classify_kernel put on the monster list are handed out five tickets per job, one per sample
    // line: a scanline with hundreds of band entries is the longest single piece of work of the launch and its five
    // lines are independent until their spans are concatenated.
    if (hj >= numHeavy + monTickets) break;
    int splitLine = 0;
    int bj;
    const bool isMon = doSplit && hj < monTickets;
    if (hj < monTickets) {
      const int idx = doSplit ? hj / 5 : hj;
      splitLine = doSplit ? hj - idx * 5 : 0;
      bj = A.monsterList[idx];
    } else {
      const int h2 = hj - monTickets;
      bj = A.heavyList[h2 < numFront ? h2 : A.planJobs - 1 - (h2 - numFront)];
    }
    const int f = find_fill_by_job(A.planJobBase, A.numFills, bj);
    const FillHeader* Hp = A.fills + f;
    const int startY = Hp->startY, rule = Hp->rule;
    const int y = max(startY, A.planY0) + (bj - A.planJobBase[f]);
    const int job = A.fillJobBase[f] + (y - startY);
    const int ph = Hp->partitionHeight;
    int p = (y - startY) / ph;
    if (p > Hp->numPartitions - 1) p = Hp->numPartitions - 1;
    const int gp = Hp->partBase + p;
    const int eBeg = A.entryOff[gp];
    const int eCnt = A.entryOff[gp + 1] - eBeg;
    const unsigned fl = A.flags[gp];
    const Entry* ent = A.entries + eBeg;
    uint2* pay = A.payload + ((size_t)A.payOff[gp] + (size_t)(y - (startY + p * ph)) * (size_t)eCnt) * kPaySlots;
    uint32_t* scr = eCnt > A.smemCap ? gscr : sscr;
    const int cap = eCnt > A.smemCap ? A.scratchCap : A.smemCap;
    const bool aa = (fl & 1u) != 0, two = (fl & 2u) != 0;
    JobHdr hdr;
    hdr.kind = PlanNothing;
    hdr.n = 0;
    hdr.pa = 0;
    hdr.pb = 0;
    __syncwarp();

    if (two && !aa) {  // mode A (:1644-1668): two vertical pixel-aligned lines
      hdr.kind = PlanAligned;
      hdr.pa = clampi(f2ll(ent[0].ax), 0, W);
      hdr.pb = clampi(f2ll(ent[1].ax), 0, W);
      if (lane == 0) A.jobs[job] = hdr;
      continue;
    }

    int* sel = reinterpret_cast<int*>(scr);
    int* sel2 = sel + cap;
    const float scanTop = (float)y, scanBottom = (float)(y + 1);
    bool allSpan = true;
    int nsel = 0;
    const int* selC = sel;  // entry order seen by computeCoverage
    if (two) {
      nsel = 2;
      if (lane < 2) sel[lane] = lane;
    } else {  // :1681-1689
#pragma unroll 1
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
    // one warp per sample line: anti-aliased scanlines that go through computeCoverage with the bitonic sort
    const bool modeB = allSpan && (nsel % 2) == 0;
    const bool split = isMon && !modeB && nsel > kBitonicMin && nsel <= kBitonicMax && scr != sscr;
    if (!split && splitLine != 0) continue;  // planned whole by the line-0 ticket

    if (allSpan && (nsel % 2) == 0) {  // mode B (:1691-1872)
      float* tax = reinterpret_cast<float*>(scr + 2 * cap);
      float* tbx = reinterpret_cast<float*>(scr + 3 * cap);
      float* mid = reinterpret_cast<float*>(scr + 4 * cap);
      int* wnd = reinterpret_cast<int*>(scr + 5 * cap);
      int* order = reinterpret_cast<int*>(scr + 7 * cap);
#pragma unroll 1
      for (int s = lane; s < nsel; s += 32) {
        const Entry e = ent[sel[s]];
        const float xa = solve_x(e.m, e.b, scanTop), xb = solve_x(e.m, e.b, scanBottom);
        tax[s] = xa;
        tbx[s] = xb;
        mid[s] = (xa + xb) * 0.5f;
        wnd[s] = e.winding;
      }
      __syncwarp();
      // stable sort by mid x (= the reference's insertion sort, :1707-1716): order[rank] = selected position
      if (nsel > kBitonicMin && nsel <= kBitonicMax && scr != sscr) {
        int P = 2 * kBitonicMin;
        while (P < nsel) P <<= 1;
        uint2* key = reinterpret_cast<uint2*>(sscr);  // the warp's shared-memory scratch is free while the band's arrays live in HBM
#pragma unroll 1
        for (int i = lane; i < P; i += 32) key[i] = make_uint2((uint32_t)i, i < nsel ? sortable_float(mid[i]) : 0xFFFFFFFFu);
        __syncwarp();
        warp_bitonic_sort(key, P, lane);
#pragma unroll 1
        for (int i = lane; i < nsel; i += 32) order[i] = (int)key[i].x;
      } else {
#pragma unroll 1
        for (int i = lane; i < nsel; i += 32) {
          const float ki = mid[i];
          int r = 0;
#pragma unroll 4
          for (int j = 0; j < nsel; j++) {
            const float kj = mid[j];
            r += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
          }
          order[r] = i;
        }
      }
      __syncwarp();
      bool ok = true;
#pragma unroll 1
      for (int i = lane; i < nsel - 1; i += 32) {  // partial-coverage areas must not overlap (:1720-1728)
        const int l = order[i], r = order[i + 1];
        const float leftMaxX = fmaxf(tax[l], tbx[l]), rightMinX = fminf(tax[r], tbx[r]);
        if (f2ll(ceilf(leftMaxX)) > f2ll(rightMinX)) ok = false;
      }
      ok = __all_sync(0xffffffffu, ok);
      if (ok) {  // only simple fill pairs (:1732-1744): prefix winding count, filled after even, empty after odd
        int carry = 0;
#pragma unroll 1
        for (int base = 0; base < nsel; base += 32) {
          const int i = base + lane;
          int pre = (i < nsel) ? wnd[order[i]] : 0;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += t;
          }
          pre += carry;
          if (i < nsel) {
            const bool fillIt = should_fill(rule, pre);
            if (((i & 1) == 0) ? !fillIt : fillIt) ok = false;
          }
          carry = __shfl_sync(0xffffffffu, pre, 31);
        }
        ok = __all_sync(0xffffffffu, ok);
      }
      if (ok) {  // payload: per sorted edge {m, b} {x at the top, x at the bottom of the scanline}
        hdr.kind = PlanTrapezoids;
        hdr.n = nsel;
        int lo = INT_MAX, hi = INT_MIN;
#pragma unroll 1
        for (int i = lane; i < nsel; i += 32) {
          const int es = order[i];
          const Entry e = ent[sel[es]];
          pay[2 * i] = make_uint2(__float_as_uint(e.m), __float_as_uint(e.b));
          pay[2 * i + 1] = make_uint2(__float_as_uint(tax[es]), __float_as_uint(tbx[es]));
          lo = extent_lo(lo, tax[es], tbx[es]);
          hi = extent_hi(hi, tax[es], tbx[es]);
        }
        hdr.pa = __reduce_min_sync(0xffffffffu, lo);
        hdr.pb = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) A.jobs[job] = hdr;
        continue;
      }
      // The reference sorts entryIndices in place (:1707-1716) before it decides whether the shortcut
      // applies, so when it falls through, computeCoverage walks the entries in mid-x order.
#pragma unroll 1
      for (int i = lane; i < nsel; i += 32) sel2[i] = sel[order[i]];
      selC = sel2;
      __syncwarp();
    }

    // mode C: computeCoverage (:1350-1431).  The reference walks the `quality` sample lines one after the
    // other; they are independent until their spans meet in the coverage row, so item t = m * nsel + s
    // (line m, selected entry s) gets its own lane: hits, the stable sort by x (rank = number of hits of
    // the same line that sort before it, ties by entry order = the insertion sort of :1277-1286) and the
    // `walk` of each line (lane m) run concurrently.
    const int quality = aa ? 5 : 1;
    const float offset = 1.0f / (float)quality;
    const float initialOffset = offset / 2.0f + (float)(0.0001 * 3.141592653589793238462643383279502884);
    const int n = nsel, T = quality * n;
    int* uAt = reinterpret_cast<int*>(scr + 2 * cap);  // hit x per item, entry order (kNoHit: none)
    int* sAt = uAt + T;                                // per line: hits sorted by x, later the spans' begin
    int* sW = sAt + T;                                 //           their windings,    later the spans' end
    int* lineHits = sW + T;                            // [quality]
    constexpr int kNoHit = INT_MAX;
    hdr.kind = aa ? PlanCoverage : PlanSpans;
    if (split) {
      const int m = splitLine;
      int P = 2 * kBitonicMin;
      while (P < n) P <<= 1;
      uint2* key = reinterpret_cast<uint2*>(sscr);
      float yLine = (float)y + initialOffset - offset;
#pragma unroll 1
      for (int k = 0; k <= m; k++) yLine += offset;  // the reference accumulates yLine line by line (:1362-1371)
#pragma unroll 1
      for (int s_ = lane; s_ < P; s_ += 32) {
        int at = kNoHit;
        if (s_ < n) {
          const Entry e = ent[selC[s_]];
          if (e.ay <= yLine && e.by >= yLine) {
            float x = e.m == 0.0f ? e.b : (yLine - e.b) / e.m;
            x = (x != x) ? wf : (x < wf ? x : wf);  // min(x, width.float32)
            at = fixed32(x);
          }
        }
        key[s_] = make_uint2((uint32_t)s_, (uint32_t)at ^ 0x80000000u);  // kNoHit -> 0xFFFFFFFF sorts last
      }
      __syncwarp();
      warp_bitonic_sort(key, P, lane);
      int cnt = 0;
#pragma unroll 1
      for (int base = 0; base < n; base += 32) {
        const int r_ = base + lane;
        const uint2 kv = r_ < n ? key[r_] : make_uint2(0u, 0xFFFFFFFFu);
        const bool hit = kv.y != 0xFFFFFFFFu;
        if (hit) {
          sAt[r_] = (int)(kv.y ^ 0x80000000u);
          sW[r_] = ent[selC[kv.x]].winding;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, hit));
      }
      __syncwarp();
      int ns = 0;
      if (lane == 0) ns = walk_spans(sAt, sW, cnt, rule, sAt, sW);
      ns = __shfl_sync(0xffffffffu, ns, 0);
      uint2* lp = pay + (size_t)m * (size_t)eCnt;
#pragma unroll 1
      for (int g = lane; g < ns; g += 32) lp[g] = make_uint2((uint32_t)sAt[g], (uint32_t)sW[g]);
      if (lane == 0) lp[eCnt - 1] = make_uint2((uint32_t)ns, 0u);
      __threadfence();
      __syncwarp();
      unsigned old = 0;
      if (lane == 0) old = atomicInc(reinterpret_cast<unsigned*>(A.splitArrive + job), (unsigned)(quality - 1));
      old = __shfl_sync(0xffffffffu, old, 0);
      if (old != (unsigned)(quality - 1)) continue;  // other lines of the scanline are still being planned
      __threadfence();
      int cntOf[5], S = 0;
#pragma unroll
      for (int q = 0; q < 5; q++) cntOf[q] = (int)__ldcg(reinterpret_cast<const unsigned*>(&pay[(size_t)q * eCnt + eCnt - 1].x));
      int lo = INT_MAX, hi = INT_MIN;
#pragma unroll 1
      for (int q = 0; q < 5; q++) {
        const uint2* srcp = pay + (size_t)q * (size_t)eCnt;
#pragma unroll 1
        for (int base = 0; base < cntOf[q]; base += 32) {
          const int g = base + lane;
          uint2 v = make_uint2(0u, 0u);
          if (g < cntOf[q]) {
            v.x = __ldcg(&srcp[g].x);
            v.y = __ldcg(&srcp[g].y);
          }
          __syncwarp();
          if (g < cntOf[q]) {
            pay[S + g] = v;
            lo = min(lo, fx_integer((int)v.x));
            hi = max(hi, fx_integer((int)v.y) + 1);
          }
          __syncwarp();
        }
        S += cntOf[q];
      }
      hdr.n = S;
      hdr.pa = __reduce_min_sync(0xffffffffu, lo);
      hdr.pb = __reduce_max_sync(0xffffffffu, hi);
      if (lane == 0) A.jobs[job] = hdr;
      continue;
    }
    if (n > 0) {
      if (n > kBitonicMin && n <= kBitonicMax && scr != sscr) {
        // crowded scanline: line after line, hits keyed {x, entry position} through the bitonic sort; the keys
        // live in the warp's shared-memory scratch, which is free while the band's arrays are in HBM
        int P = 2 * kBitonicMin;
        while (P < n) P <<= 1;
        uint2* key = reinterpret_cast<uint2*>(sscr);
        float yLine = (float)y + initialOffset - offset;
#pragma unroll 1
        for (int m = 0; m < quality; m++) {
          yLine += offset;  // the reference accumulates yLine line by line (:1362-1371)
#pragma unroll 1
          for (int s = lane; s < P; s += 32) {
            int at = kNoHit;
            if (s < n) {
              const Entry e = ent[selC[s]];
              if (e.ay <= yLine && e.by >= yLine) {
                float x = e.m == 0.0f ? e.b : (yLine - e.b) / e.m;
                x = (x != x) ? wf : (x < wf ? x : wf);  // min(x, width.float32)
                at = fixed32(x);
              }
            }
            key[s] = make_uint2((uint32_t)s, (uint32_t)at ^ 0x80000000u);  // kNoHit -> 0xFFFFFFFF sorts last
          }
          __syncwarp();
          warp_bitonic_sort(key, P, lane);
          int cnt = 0;
#pragma unroll 1
          for (int base = 0; base < n; base += 32) {
            const int r = base + lane;
            const uint2 kv = r < n ? key[r] : make_uint2(0u, 0xFFFFFFFFu);
            const bool hit = kv.y != 0xFFFFFFFFu;
            if (hit) {
              sAt[m * n + r] = (int)(kv.y ^ 0x80000000u);
              sW[m * n + r] = ent[selC[kv.x]].winding;
            }
            cnt += __popc(__ballot_sync(0xffffffffu, hit));
          }
          if (lane == 0) lineHits[m] = cnt;
          __syncwarp();
        }
      } else {
#pragma unroll 1
      for (int t = lane; t < T; t += 32) {  // :1373-1385
        const int m = t / n, s = t - m * n;
        float yLine = (float)y + initialOffset - offset;
#pragma unroll 1
        for (int k = 0; k <= m; k++) yLine += offset;  // the reference accumulates yLine line by line (:1362-1371)
        const Entry e = ent[selC[s]];
        int at = kNoHit;
        if (e.ay <= yLine && e.by >= yLine) {
          float x = e.m == 0.0f ? e.b : (yLine - e.b) / e.m;
          x = (x != x) ? wf : (x < wf ? x : wf);  // min(x, width.float32)
          at = fixed32(x);
        }
        uAt[t] = at;
      }
      __syncwarp();
      int myHits = 0;  // lane m < quality: hits of sample line m (items whose position sorts before the kNoHit ones)
#pragma unroll 1
      for (int t = lane; t < T; t += 32) {
        const int m = t / n, s = t - m * n;
        const int at = uAt[t];
        const int* line = uAt + m * n;
        // stable rank: entries before s count when <= at, entries after it when < at (at + 1 cannot overflow: a hit
        // is at most width * 256, and an item without a hit is not placed)
        int r = 0;
        const int atB = at == kNoHit ? at : at + 1;
#pragma unroll 4
        for (int j = 0; j < n; j++) r += line[j] < (j < s ? atB : at) ? 1 : 0;
        const bool hit = at != kNoHit;
        if (hit) {
          sAt[m * n + r] = at;
          sW[m * n + r] = ent[selC[s]].winding;
        }
      }
      if (lane < quality) {  // hits per line: one lane per line scans its n items (this sat in the rank loop above, n times over)
        const int* line = uAt + lane * n;
#pragma unroll 4
        for (int j = 0; j < n; j++) myHits += line[j] != kNoHit ? 1 : 0;
      }
      if (lane < quality) lineHits[lane] = myHits;
      }
      __syncwarp();
      int ns = 0;  // spans of line `lane`
      if (lane < quality) ns = walk_spans(sAt + lane * n, sW + lane * n, lineHits[lane], rule, sAt + lane * n, sW + lane * n);
      __syncwarp();
      const int n0 = __shfl_sync(0xffffffffu, ns, 0), n1 = __shfl_sync(0xffffffffu, ns, 1);
      const int n2 = __shfl_sync(0xffffffffu, ns, 2), n3 = __shfl_sync(0xffffffffu, ns, 3);
      const int n4 = __shfl_sync(0xffffffffu, ns, 4);
      const int pre1 = n0, pre2 = pre1 + n1, pre3 = pre2 + n2, pre4 = pre3 + n3, S = pre4 + n4;
      hdr.n = S;
      int lo = INT_MAX, hi = INT_MIN;
#pragma unroll 1
      for (int g = lane; g < S; g += 32) {  // payload: the spans of all lines, {begin, end} in 24.8 fixed point
        const int m = (g >= pre1) + (g >= pre2) + (g >= pre3) + (g >= pre4);
        const int k = g - (m == 0 ? 0 : m == 1 ? pre1 : m == 2 ? pre2 : m == 3 ? pre3 : pre4);
        const int b_ = sAt[m * n + k], e_ = sW[m * n + k];
        pay[g] = make_uint2((uint32_t)b_, (uint32_t)e_);
        lo = min(lo, fx_integer(b_));
        hi = max(hi, fx_integer(e_) + 1);
      }
      hdr.pa = __reduce_min_sync(0xffffffffu, lo);
      hdr.pb = __reduce_max_sync(0xffffffffu, hi);
    }
    if (lane == 0) A.jobs[job] = hdr;
  }
#ifdef PIXIE_RASTER_TIMING
  if ((threadIdx.x & 31) == 0) atomicMax(&A.counters[51], gtime());
#endif
  if (A.gscratch) {
    __syncthreads();
    if (threadIdx.x == 0) atomicAnd(A.scratchSlots + (s_slot >> 5), ~(1u << (s_slot & 31)));
  }
}

// ---------------------------------------------------------------------------------------------
// K3: raster_kernel — apply the plans to the canvas, row by row, fills in order
// ---------------------------------------------------------------------------------------------
constexpr int kPartCap = 128;
constexpr int kMaxSubShift = 3;  // a tile of a busy row in up to 8 pieces
struct WarpCtx {
  int mode;         // run-time blend mode (used by the GenericMode instantiation)
  px_t* row;        // canvas row of this warp
  int w;            // canvas width
  int tx0, tx1;     // the columns [tx0, tx1) this warp owns (a tile of the row); everything it writes is clamped to them
  int y;
  int lane;
  uint8_t* cov;     // coverage row in shared memory (index 0 = pixel covBase)
  uint16_t* plist;  // kPartCap word indices: the partially covered words of the job being blended
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

__device__ __noinline__ void clear_span(WarpCtx& c, int x0, int x1) {  // [x0, x1) -> transparent
  x0 = max(x0, c.tx0);
  x1 = min(x1, c.tx1);
  px_t* row = c.row;
#pragma unroll 1
  for (int x = x0 + c.lane; x < x1; x += 32) row[x] = 0u;
}

// interior span of fillHits: [x0, x1) gets the solid colour blended in
template <int MODE>
__device__ __noinline__ void fill_span(WarpCtx& c, int x0, int x1, px_t rgbx) {
  x0 = max(x0, c.tx0);
  x1 = min(x1, c.tx1);
  if (x1 <= x0) return;
  const bool store_only = MODE == OverwriteBlend || (MODE == NormalBlend && pA(rgbx) == 255u);
  const int lane = c.lane, mode = c.mode;
  if (MODE == MaskBlend && pA(rgbx) == 255u) {  // :1576-1577: opaque mask leaves the span untouched
    if (lane == 0) c.covered += (unsigned)(x1 - x0);
    return;
  }
  px_t* row = c.row;
  unsigned cnt = 0;
  int xa = x1, xb = x1;  // [xa, xb): the 16-byte aligned middle
  if (c.vec_ok && x1 - x0 >= 64) {
    xa = (x0 + 3) & ~3;
    xb = x1 & ~3;
  }
#pragma unroll 1
  for (int x = x0 + lane; x < xa; x += 32) {
    row[x] = store_only ? rgbx : span_op<MODE>(mode, row[x], rgbx);
    cnt++;
  }
  if (store_only) {
    const uint4 v = make_uint4(rgbx, rgbx, rgbx, rgbx);
#pragma unroll 1
    for (int x = xa + 4 * lane; x < xb; x += 128) {
      *reinterpret_cast<uint4*>(row + x) = v;
      cnt += 4;
    }
  } else {  // read-modify-write: two 16-byte groups per lane in flight
#pragma unroll 1
    for (int xs = xa + 4 * lane; xs < xb; xs += 256) {
      const bool two = xs + 128 < xb;
      uint4 v0 = *reinterpret_cast<const uint4*>(row + xs), v1 = v0;
      if (two) v1 = *reinterpret_cast<const uint4*>(row + xs + 128);
      v0.x = span_op<MODE>(mode, v0.x, rgbx); v0.y = span_op<MODE>(mode, v0.y, rgbx);
      v0.z = span_op<MODE>(mode, v0.z, rgbx); v0.w = span_op<MODE>(mode, v0.w, rgbx);
      *reinterpret_cast<uint4*>(row + xs) = v0;
      cnt += 4;
      if (two) {
        v1.x = span_op<MODE>(mode, v1.x, rgbx); v1.y = span_op<MODE>(mode, v1.y, rgbx);
        v1.z = span_op<MODE>(mode, v1.z, rgbx); v1.w = span_op<MODE>(mode, v1.w, rgbx);
        *reinterpret_cast<uint4*>(row + xs + 128) = v1;
        cnt += 4;
      }
    }
  }
#pragma unroll 1
  for (int x = xb + lane; x < x1; x += 32) {
    row[x] = store_only ? rgbx : span_op<MODE>(mode, row[x], rgbx);
    cnt++;
  }
  c.covered += cnt;
}

// One pixel of a trapezoid edge of scanline y: blender()(backdrop, rgbx * area).  Left edges (:1772-1809)
// cover pixels [trunc(min x), ceil(max x)), right edges (:1811-1847) likewise; xa / xb are the edge's x at
// the top and the bottom of the scanline, `first` the first pixel of the edge (where the pen starts).
template <int MODE>
__device__ __noinline__ void edge_px(px_t* row, int mode, int y, int xl, bool left, float em, float eb, float xa,
                                     float xb, int first, px_t rgbx) {
  const int x = xl;
  float area;
  if (left) {
    const bool inverted = xa < xb;
    const float sliverStart = fminf(xa, xb), rectStart = fmaxf(xa, xb);
    const float prevPen = (xl == first) ? sliverStart : (float)x;
    const float prevPenY = (xl == first) ? (inverted ? (float)y : (float)(y + 1)) : (em * (float)x + eb);
    float pen = (float)(x + 1), rightRectArea = 0.0f;
    if (pen > rectStart) {
      rightRectArea = pen - rectStart;
      pen = rectStart;
    }
    const float penY = em * pen + eb;
    const float run = pen - prevPen;
    const float triangleArea = 0.5f * run * fabsf(penY - prevPenY);
    const float rectArea = inverted ? (prevPenY - (float)y) * run : ((float)(y + 1) - prevPenY) * run;
    area = triangleArea + rectArea + rightRectArea;
  } else {
    const bool inverted = xa > xb;
    const float rectEnd = fminf(xa, xb), sliverEnd = fmaxf(xa, xb);
    const float prevPen = (xl == first) ? rectEnd : (float)x;
    const float prevPenY = (xl == first) ? (inverted ? (float)(y + 1) : (float)y) : (em * (float)x + eb);
    float pen = (float)(x + 1);
    const float leftRectArea = frac_vmath(prevPen);
    if (pen > sliverEnd) pen = sliverEnd;
    const float penY = em * pen + eb;
    const float run = pen - prevPen;
    const float triangleArea = 0.5f * run * fabsf(penY - prevPenY);
    const float rectArea = inverted ? (penY - (float)y) * run : ((float)(y + 1) - penY) * run;
    area = leftRectArea + triangleArea + rectArea;
  }
  row[x] = blend_any<MODE>(mode, row[x], mul_area(rgbx, area));  // (:1803-1809, :1841-1847)
}

PXD void smem_add(uint32_t* p, uint32_t v) {  // shared-memory reduction, no return value
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// computeCoverage's accumulation (:1391-1431) of S spans into the coverage row, then fillCoverage
// (:1479-1526), then the row is zero again (:1896).  Spans of one sample line are disjoint and a line adds
// at most 255 div 5 to a pixel, so a coverage byte never exceeds 255 and word-wide adds carry nothing:
// every span gets a lane, long interiors are handed to the whole warp.
template <int MODE>
__device__ __noinline__ void coverage_row(WarpCtx& c, const uint2* __restrict__ spans, int S, int startX, int pathWidth,
                                          px_t rgbx) {
  const int lane = c.lane, mode = c.mode;
  // coverages[] of the reference spans [startX, startX + pathWidth); this warp keeps the part inside its tile
  const int covX0 = max(startX, c.tx0), covX1 = min(startX + pathWidth, c.tx1);
  const int covBase = c.vec_ok ? (covX0 & ~3) : covX0;
  constexpr int sampleCoverage = 255 / 5;
  uint8_t* cov = c.cov;
  uint32_t* cw = reinterpret_cast<uint32_t*>(cov);
  int pxLo = INT_MAX, pxHi = INT_MIN;
  const uint32_t add4 = 0x01010101u * (uint32_t)sampleCoverage;
#pragma unroll 1
  for (int base = 0; base < S; base += 32) {
    const int g = base + lane;
    int i0 = 0, i1 = 0;
    if (g < S) {
      const uint2 sp = spans[g];
      const int prevAt = (int)sp.x, at = (int)sp.y;
      int fillStart = fx_integer(prevAt);
      const bool pixelCrossed = fx_integer(at) != fx_integer(prevAt);
      const int leftCover = pixelCrossed ? fx_trunc(prevAt) + 256 - prevAt : at - prevAt;
      pxLo = min(pxLo, fillStart);
      pxHi = max(pxHi, fx_integer(at) + 1);
      if (leftCover != 0) {
        fillStart++;
        const int px = fx_integer(prevAt), idx = px - covBase;
        const uint32_t v = (uint32_t)(uint8_t)fx_integer(leftCover * sampleCoverage);
        if (px >= covX0 && px < covX1 && v) smem_add(&cw[idx >> 2], v << (8 * (idx & 3)));
      }
      if (pixelCrossed) {
        const int rightCover = at - fx_trunc(at);
        if (rightCover > 0) {
          const int px = fx_integer(at), idx = px - covBase;
          const uint32_t v = (uint32_t)(uint8_t)fx_integer(rightCover * sampleCoverage);
          if (px >= covX0 && px < covX1 && v) smem_add(&cw[idx >> 2], v << (8 * (idx & 3)));
        }
      }
      i0 = max(fillStart, covX0) - covBase;
      i1 = min(fx_integer(at), covX1) - covBase;
    }
    // interiors: +sampleCoverage on bytes [i0, i1) as masked word adds
    const bool some = i1 > i0;
    const bool isLong = some && (((i1 - 1) >> 2) - (i0 >> 2) >= 6);
    if (some && !isLong) {
      const int wl = (i1 - 1) >> 2;
#pragma unroll 1
      for (int w = i0 >> 2; w <= wl; w++) {
        uint32_t mask = 0xFFFFFFFFu;
        if (w == (i0 >> 2)) mask &= 0xFFFFFFFFu << (8 * (i0 & 3));
        if (w == wl) mask &= 0xFFFFFFFFu >> (8 * (3 - ((i1 - 1) & 3)));
        smem_add(&cw[w], add4 & mask);
      }
    }
    // long interiors: the owner adds its first and last (masked) word, the warp the whole words in between
    int wA = 0, wB = 0;
    if (isLong) {
      wA = (i0 >> 2) + 1;
      wB = (i1 - 1) >> 2;
      smem_add(&cw[wA - 1], add4 & (0xFFFFFFFFu << (8 * (i0 & 3))));
      smem_add(&cw[wB], add4 & (0xFFFFFFFFu >> (8 * (3 - ((i1 - 1) & 3)))));
    }
    unsigned longs = __ballot_sync(0xffffffffu, isLong);
#pragma unroll 1
    while (longs) {
      const int src = __ffs(longs) - 1;
      longs &= longs - 1;
      const int a_ = __shfl_sync(0xffffffffu, wA, src), b_ = __shfl_sync(0xffffffffu, wB, src);
#pragma unroll 1
      for (int w = a_ + lane; w < b_; w += 32) smem_add(&cw[w], add4);
    }
  }
  pxLo = __reduce_min_sync(0xffffffffu, pxLo);  // redux.sync
  pxHi = __reduce_max_sync(0xffffffffu, pxHi);
  __syncwarp();
  if (MODE != MaskBlend && S == 0) return;

  // fillCoverage; [pxLo, pxHi) bounds the pixels whose coverage can be non-zero (MaskBlend visits all)
  const int x0 = covX0, x1 = covX1;
  px_t* row = c.row;
  unsigned cnt = 0;
  if (c.vec_ok) {
    // covBase is a multiple of 4: coverage word j <-> pixels covBase + 4j .. +3 (one uint4 of the row)
    int words = (x1 - covBase + 3) >> 2, word0 = 0;
    if (MODE != MaskBlend) {
      word0 = (max(pxLo, x0) - covBase) >> 2;
      words = min(words, (min(pxHi, x1) - covBase + 3) >> 2);
    }
    // one word (4 pixels) of coverage against its 16 bytes of canvas
    auto blend_word = [&](uint32_t cv, int x, uint4 v) -> uint4 {
      uint32_t* vp = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t cvk = (cv >> (8 * k)) & 255u;
        const int xx = x + k;
        if (xx < x0 || xx >= x1) continue;
        if (cvk != 0u) cnt++;
        if (MODE == OverwriteBlend) {
          if (cvk != 0u) vp[k] = mul_cov_floor(rgbx, cvk);
        } else if (MODE == NormalBlend) {
          if (cvk != 0u) vp[k] = line_normal(vp[k], mul_cov_floor(rgbx, cvk));
        } else if (MODE == MaskBlend) {
          vp[k] = line_mask(vp[k], mul_cov_floor(rgbx, cvk));
        } else {
          if (cvk != 0u) vp[k] = blend_any<MODE>(mode, vp[k], mul_cov_round(rgbx, cvk));
        }
      }
      return v;
    };
    if (MODE == NormalBlend || MODE == OverwriteBlend) {
      // Words whose four pixels are all fully covered are handled word-wise by every lane (a plain 16-byte store
      // for opaque colours).  The others — the anti-aliased span ends, a handful per job — used to run the
      // four-pixel blend on two or three lanes of the warp; they are listed in shared memory instead and then
      // blended one PIXEL per lane, so the edge arithmetic fills the warp.
      const bool solid = MODE == OverwriteBlend || pA(rgbx) == 255u;
      const uint4 colour4 = make_uint4(rgbx, rgbx, rgbx, rgbx);
      uint16_t* plist = c.plist;
      int nPart = 0;  // warp-uniform
      auto flush = [&]() {
        __syncwarp();
#pragma unroll 1
        for (int t = lane; t < 4 * nPart; t += 32) {
          const int j = plist[t >> 2], x = covBase + 4 * j + (t & 3);
          const uint32_t cvk = cov[4 * j + (t & 3)];
          if (cvk != 0u && x >= x0 && x < x1) {
            cnt++;
            const px_t s_ = mul_cov_floor(rgbx, cvk);
            row[x] = MODE == OverwriteBlend ? s_ : line_normal(row[x], s_);
          }
        }
        __syncwarp();
#pragma unroll 1
        for (int t = lane; t < nPart; t += 32) cw[plist[t]] = 0u;
        nPart = 0;
        __syncwarp();
      };
#pragma unroll 1
      for (int j0 = word0; j0 < words; j0 += 32) {
        const int j = j0 + lane;
        const uint32_t cv = j < words ? cw[j] : 0u;
        const int xw = covBase + 4 * j;
        const bool full = cv == 0xFFFFFFFFu && (xw >= x0) && (xw + 4 <= x1);
        if (full) {
          cw[j] = 0u;
          cnt += 4;
          uint4 v = colour4;
          if (!solid) {  // translucent colour over four fully covered pixels
            v = *reinterpret_cast<const uint4*>(row + xw);
            v.x = line_normal(v.x, rgbx); v.y = line_normal(v.y, rgbx);
            v.z = line_normal(v.z, rgbx); v.w = line_normal(v.w, rgbx);
          }
          *reinterpret_cast<uint4*>(row + xw) = v;  // sse2.nim:552-555, 648-651
        }
        const bool part = cv != 0u && !full;
        const unsigned pm = __ballot_sync(0xffffffffu, part);
        if (pm) {
          if (part) plist[nPart + __popc(pm & ((1u << lane) - 1u))] = (uint16_t)j;
          nPart += __popc(pm);
          if (nPart > kPartCap - 32) flush();
        }
      }
      if (nPart) flush();
    } else {
#pragma unroll 1
      for (int j = word0 + lane; j < words; j += 32) {
        const uint32_t cv = cw[j];
        const int x = covBase + 4 * j;
        if (MODE != MaskBlend && cv == 0u) continue;
        cw[j] = 0u;
        uint4* p = reinterpret_cast<uint4*>(row + x);
        const bool full = (x >= x0) && (x + 4 <= x1);
        if (MODE == MaskBlend && pA(rgbx) == 255u && cv == 0xFFFFFFFFu && full) {  // sse2.nim:750-751
          cnt += 4;
          continue;
        }
        *p = blend_word(cv, x, *p);
      }
    }
  } else {
    const int xlo = MODE != MaskBlend ? max(pxLo, x0) : x0, xhi = MODE != MaskBlend ? min(pxHi, x1) : x1;
#pragma unroll 1
    for (int x = xlo + lane; x < xhi; x += 32) {
      const uint32_t cvk = cov[x - covBase];
      cov[x - covBase] = 0;
      if (cvk != 0u) cnt++;
      if (MODE == OverwriteBlend) {
        if (cvk != 0u) row[x] = mul_cov_floor(rgbx, cvk);
      } else if (MODE == NormalBlend) {
        if (cvk != 0u) row[x] = line_normal(row[x], mul_cov_floor(rgbx, cvk));
      } else if (MODE == MaskBlend) {
        row[x] = line_mask(row[x], mul_cov_floor(rgbx, cvk));
      } else {
        if (cvk != 0u) row[x] = blend_any<MODE>(mode, row[x], mul_cov_round(rgbx, cvk));
      }
    }
  }
  c.covered += cnt;
  if (MODE == MaskBlend) {  // :1516-1517
    clear_span(c, 0, startX);
    clear_span(c, startX + pathWidth, c.w);
  }
}

// apply_row: what the planned fill does to the canvas row (fillHits :1540-1591, the trapezoid pixels
// :1772-1866, fillCoverage :1479-1526), for one class of blend modes.
template <int MODE>
__device__ __noinline__ void apply_row(WarpCtx& c, px_t rgbx, int startX, int pathWidth, int kind, int n, int pa, int pb,
                                       const uint2* __restrict__ pay) {
  const int lane = c.lane, W = c.w, y = c.y;
  if (kind == PlanOutside) {
    if (MODE == MaskBlend) clear_span(c, 0, W);
  } else if (kind == PlanAligned) {
    const int minX = pa, maxX = pb;
    if (maxX > minX) {
      if (MODE == MaskBlend) clear_span(c, 0, minX);
      fill_span<MODE>(c, minX, maxX, rgbx);
      if (MODE == MaskBlend) clear_span(c, maxX, W);
    } else if (MODE == MaskBlend) {
      clear_span(c, 0, W);
    }
  } else if (kind == PlanTrapezoids) {
    px_t* row = c.row;
    const int mode = c.mode;
    if (MODE != MaskBlend) {
      // The partial-coverage pixels of the sorted edges and the interiors between them are pairwise
      // disjoint (that is what the two checks of the plan establish), so they can be written in any
      // order: every edge gets a lane, long edges and long interiors are handed to the whole warp.
      // Pixel positions are kept as saturated int32 (f2i_sat): every use is clamped to the tile or compared with a
      // pixel inside it, which gives what the reference's 64-bit values give.  Rows with up to four edges (most of
      // them: one or two shapes' worth) spread each edge over eight lanes, one pixel per lane.
      unsigned cnt = 0;
      const int sh = n <= 4 ? 3 : 0, per = 32 >> sh, sub = lane & ((1 << sh) - 1), slot = lane >> sh;
#pragma unroll 1
      for (int base = 0; base < n; base += per) {
        const int i = base + slot;
        float em = 0.0f, eb = 0.0f, xa = 0.0f, xb = 0.0f;
        int xFirst = 0, pxB = 0, pxE = 0, xEnd = 0;
        if (i < n) {
          const uint2 mb = pay[2 * i], ab = pay[2 * i + 1];
          em = __uint_as_float(mb.x); eb = __uint_as_float(mb.y);
          xa = __uint_as_float(ab.x); xb = __uint_as_float(ab.y);
          xFirst = f2i_sat(fminf(xa, xb));
          xEnd = f2i_sat(ceilf(fmaxf(xa, xb)));
          pxB = min(max(xFirst, c.tx0), c.tx1);  // inside the image and inside this warp's tile
          pxE = min(max(xEnd, c.tx0), c.tx1);
        }
        const bool isLeft = (slot & 1) == 0;  // base is a multiple of `per`: even sorted positions are left edges
        // interior of the pair: [ceil(left max x), trunc(right min x)) (:1849-1854); the right edge is the next slot
        const int rightMin = __shfl_down_sync(0xffffffffu, xFirst, 1 << sh);
        int fillBegin = 0, fillEnd = 0;
        if (i < n && isLeft && sub == 0) {
          fillBegin = pxE;
          fillEnd = min(max(rightMin, c.tx0), c.tx1);
        }
        const bool longEdge = pxE - pxB > 6;
        if (!longEdge) {
#pragma unroll 1
          for (int xl = pxB + sub; xl < pxE; xl += 1 << sh) {
            edge_px<MODE>(row, mode, y, xl, isLeft, em, eb, xa, xb, xFirst, rgbx);
            cnt++;
          }
        }
        unsigned todoE = __ballot_sync(0xffffffffu, longEdge && sub == 0);
#pragma unroll 1
        while (todoE) {
          const int src = __ffs(todoE) - 1;
          todoE &= todoE - 1;
          const float sm_ = __shfl_sync(0xffffffffu, em, src), sb_ = __shfl_sync(0xffffffffu, eb, src);
          const float sa_ = __shfl_sync(0xffffffffu, xa, src), sx_ = __shfl_sync(0xffffffffu, xb, src);
          const int sf_ = __shfl_sync(0xffffffffu, xFirst, src);
          const int b_ = __shfl_sync(0xffffffffu, pxB, src), e_ = __shfl_sync(0xffffffffu, pxE, src);
#pragma unroll 1
          for (int xl = b_ + lane; xl < e_; xl += 32) {
            edge_px<MODE>(row, mode, y, xl, ((src >> sh) & 1) == 0, sm_, sb_, sa_, sx_, sf_, rgbx);
            cnt++;
          }
        }
        const bool store_only = MODE == OverwriteBlend || (MODE == NormalBlend && pA(rgbx) == 255u);
        const bool longFill = fillEnd - fillBegin > 12;
        if (!longFill) {
#pragma unroll 1
          for (int x = fillBegin; x < fillEnd; x++) {
            row[x] = store_only ? rgbx : span_op<MODE>(mode, row[x], rgbx);
            cnt++;
          }
        }
        unsigned todoF = __ballot_sync(0xffffffffu, longFill);
#pragma unroll 1
        while (todoF) {
          const int src = __ffs(todoF) - 1;
          todoF &= todoF - 1;
          fill_span<MODE>(c, __shfl_sync(0xffffffffu, fillBegin, src), __shfl_sync(0xffffffffu, fillEnd, src), rgbx);
        }
      }
      c.covered += cnt;
    } else {  // MaskBlend: pair by pair, with the clears of :1856-1872
      int filledTo = 0;
      unsigned cnt = 0;
#pragma unroll 1
      for (int i = 0; i < n; i += 2) {
        float lax = 0.f, lbx = 0.f, rax = 0.f, rbx = 0.f;
#pragma unroll 1
        for (int side = 0; side < 2; side++) {
          const uint2 mb = pay[2 * (i + side)], ab = pay[2 * (i + side) + 1];
          const float em = __uint_as_float(mb.x), eb = __uint_as_float(mb.y);
          const float xa = __uint_as_float(ab.x), xb = __uint_as_float(ab.y);
          if (side == 0) { lax = xa; lbx = xb; } else { rax = xa; rbx = xb; }
          const int xFirst = f2i_sat(fminf(xa, xb)), xEnd = f2i_sat(ceilf(fmaxf(xa, xb)));
          const int b_ = min(max(xFirst, c.tx0), c.tx1), e_ = min(max(xEnd, c.tx0), c.tx1);
#pragma unroll 1
          for (int xl = b_ + lane; xl < e_; xl += 32) {
            edge_px<MODE>(row, mode, y, xl, side == 0, em, eb, xa, xb, xFirst, rgbx);
            cnt++;
          }
        }
        const int fillBegin = clampi(f2ll(ceilf(fmaxf(lax, lbx))), 0, W);
        const int fillEnd = clampi(f2ll(truncf(fminf(rax, rbx))), 0, W);
        fill_span<MODE>(c, fillBegin, fillEnd, rgbx);  // fillHits(..., maskClears = false) (:1849-1854)
        const long long clearTo = f2ll(fminf(lax, lbx));
        clear_span(c, min(filledTo, W), (int)(clearTo < W ? clearTo : W));
        filledTo = clampi(f2ll(ceilf(fmaxf(rax, rbx))), INT_MIN / 2, W);
      }
      clear_span(c, min(filledTo, W), W);
      c.covered += cnt;
    }
  } else if (kind == PlanCoverage) {
    coverage_row<MODE>(c, pay, n, startX, pathWidth, rgbx);
  } else if (kind == PlanSpans) {  // fillHits over walkInteger (:1897-1906, :1540-1591)
    int filledTo = startX;
#pragma unroll 1
    for (int k = 0; k < n; k++) {
      const uint2 sp = pay[k];
      const int start = fx_integer((int)sp.x);
      const int len = fx_integer((int)sp.y) - start;
      if (len <= 0) continue;
      if (MODE == MaskBlend) clear_span(c, filledTo, start);
      fill_span<MODE>(c, start, start + len, rgbx);
      filledTo = start + len;
    }
    if (MODE == MaskBlend) {
      clear_span(c, 0, startX);
      clear_span(c, filledTo, W);
    }
  }
}

// MaskBlend + trapezoid shortcut with geometry left of the canvas.  The reference clears the gaps between
// fill pairs with clearUnsafe(min(filledTo, width), y, min(clearTo, width), y) (:1856-1866, :1433-1440),
// which addresses the canvas LINEARLY (dataIndex = width * y + x): when filledTo / clearTo are negative
// the cleared range lies in the rows above y, which the same fill has already finished.  Those writes are
// inside the image, so they are part of the reference's result; a warp owns one row, so after its own
// work on row c.y it looks at the trapezoid plan of row yy > c.y and applies the part of row yy's clears
// that lands on its row.  Clears commute, so the order among the rows below does not matter.
__device__ __noinline__ void mask_wrap_clears(WarpCtx& c, int yy, int n, const uint2* __restrict__ pay) {
  const int W = c.w;
  const long long rowOff = (long long)(yy - c.y) * W;  // linear pixel index of row yy relative to row c.y
  long long filledTo = 0;
#pragma unroll 1
  for (int i = 0; i < n; i += 2) {
    const uint2 l = pay[2 * i + 1], r = pay[2 * i + 3];
    const long long clearTo = f2ll(fminf(__uint_as_float(l.x), __uint_as_float(l.y)));
    const long long a = filledTo < W ? filledTo : W, b = clearTo < W ? clearTo : W;
    if (a != W) {
      const long long x0 = rowOff + a, x1 = rowOff + b;
      if (x1 > 0 && x0 < W) clear_span(c, (int)(x0 > 0 ? x0 : 0), (int)(x1 < W ? x1 : W));
    }
    filledTo = f2ll(ceilf(fmaxf(__uint_as_float(r.x), __uint_as_float(r.y))));
  }
  const long long a = filledTo < W ? filledTo : W;
  if (a != W) {
    const long long x0 = rowOff + a;
    if (x0 < W) clear_span(c, (int)(x0 > 0 ? x0 : 0), W);
  }
  __syncwarp();
}

PXD const uint2* job_payload(const RasterArgs& A, const FillHeader* Hp, int y, int* jobOut) {
  const int startY = Hp->startY, ph = Hp->partitionHeight;
  int p = (y - startY) / ph;
  const int np = Hp->numPartitions;
  if (p > np - 1) p = np - 1;
  const int gp = Hp->partBase + p;
  const int eBeg = A.entryOff[gp];
  const int eCnt = A.entryOff[gp + 1] - eBeg;
  return A.payload + ((size_t)A.payOff[gp] + (size_t)(y - (startY + p * ph)) * (size_t)eCnt) * kPaySlots;
}

// CUT: rows are handed out in pieces (RasterArgs::subShift > 0); a separate instantiation, so that lists without cut
// rows (icon batches, row bands) run the kernel without the extra state
template <bool CUT>
__global__ void __launch_bounds__(256) raster_kernel(const RasterArgs A) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* cov = smem_raw + (size_t)warp * (size_t)A.covBytes;
  for (int i = lane * 4; i < A.covBytes; i += 128) *reinterpret_cast<uint32_t*>(cov + i) = 0u;
  __syncwarp();
  const bool vecGlobal = (A.w & 3) == 0 && (reinterpret_cast<uintptr_t>(A.canvas) & 15) == 0;
  const unsigned H_ = (unsigned)A.h;

#ifdef PIXIE_RASTER_TIMING
  const long long kern0_ = clock64();
#endif
  unsigned covered = 0;
  const int subShift = CUT ? A.subShift : 0, pieceW = A.tileW >> subShift;
  const unsigned per = (unsigned)A.tiles << subShift;
  const unsigned long long nTickets = (unsigned long long)(A.rowEnd > A.rowBegin ? A.rowEnd - A.rowBegin : 0) * per;
  while (true) {
    unsigned long long ticket = 0;
    if (lane == 0) ticket = atomicAdd(&A.counters[A.ticketSlot], 1ull);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket >= nTickets) break;
#ifdef PIXIE_RASTER_TIMING
    const long long tk0_ = clock64();
#endif
    unsigned long long rowOfTicket;
    int tx0_, txW_;
    if (!CUT) {
      // ticket -> (row, tile) without a 64-bit division (75 instructions per ticket, 7 % of an icon batch's kernel)
      unsigned long long rowRel = ticket;
      int tile = 0;
      if (A.tiles != 1) {
        if ((ticket >> 32) == 0ull) {
          const unsigned q = (unsigned)ticket / (unsigned)A.tiles;
          rowRel = q;
          tile = (int)((unsigned)ticket - q * (unsigned)A.tiles);
        } else {
          rowRel = ticket / (unsigned)A.tiles;
          tile = (int)(ticket - rowRel * (unsigned)A.tiles);
        }
      }
      rowOfTicket = rowRel + (unsigned long long)A.rowBegin;
      if (A.rowOrder) rowOfTicket = (unsigned long long)((unsigned)A.rowOrder[rowOfTicket] & 0x0FFFFFFFu);  // whole-canvas launches only (rowBegin = 0)
      tx0_ = tile * A.tileW;
      txW_ = A.tileW;
    } else {
      // A row is handed out in `per` pieces of pieceW columns; a row that is cut less finely (most rows: the cut
      // exists for the few busiest ones, whose single warp would otherwise outlast the rest of the kernel) is
      // rasterised by the ticket of the first piece of each group and the other tickets of the group return at once.
      const unsigned q = (unsigned)ticket / per;  // cut lists are single canvases: tickets fit 32 bits
      const unsigned sub = (unsigned)ticket - q * per;
      const unsigned ro = (unsigned)A.rowOrder[q];
      rowOfTicket = ro & 0x0FFFFFFFu;
      const int group = subShift - (int)(ro >> 28);  // log2 of the pieces this ticket covers
      if (sub & ((1u << group) - 1u)) continue;
      tx0_ = (int)sub * pieceW;
      if (tx0_ >= A.w) continue;
      txW_ = pieceW << group;
    }
    const unsigned t32 = (unsigned)rowOfTicket;  // layers * h < 2^31 (checked by the host)
    const int layer = t32 < H_ ? 0 : (int)(t32 / H_), y = (int)(t32 - (unsigned)layer * H_);
    WarpCtx c;
    c.row = A.canvas + (size_t)t32 * (size_t)A.w;
    c.w = A.w;
    c.tx0 = tx0_;
    c.tx1 = min(tx0_ + txW_, A.w);
    c.y = y;
    c.lane = lane;
    c.cov = cov;
    c.plist = reinterpret_cast<uint16_t*>(cov + A.covBytes - 2 * kPartCap);
    c.vec_ok = vecGlobal;
    c.covered = 0;
    if (A.clearFirst) {  // newImage(): transparent pixels, written by the warp that rasterises them next
      if (vecGlobal) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);  // tx0 and the width are multiples of 4 here
#pragma unroll 4
        for (int x = c.tx0 + 4 * lane; x < c.tx1; x += 128) *reinterpret_cast<uint4*>(c.row + x) = z;
      } else {
        for (int x = c.tx0 + lane; x < c.tx1; x += 32) c.row[x] = 0u;
      }
      __syncwarp();
    }
    const int f0 = A.layerFillBegin[layer], f1 = A.layerFillBegin[layer + 1];
#pragma unroll 1
    for (int fb = f0; fb < f1; fb += 32) {  // 32 fills per step: which of them touch this row?
      const int fl = fb + lane;
      bool act = false;
      // every lane fetches the plan of its own fill: no dependent loads in the ordered loop below
      int kind = PlanOutside, n = 0, pa = 0, pb = 0, startX = 0, pathWidth = 0, bmode = 0, wrapRows = 0, rowsBelow = 0;
      px_t rgbx = 0;
      const uint2* pay = nullptr;
      if (fl < f1) {
        const int2 rr = A.rowRange[fl];
        act = y >= rr.x && y < rr.y;
        if (act) {
          const FillHeader* Hp = A.fills + fl;
          const int sY = Hp->startY, pH = Hp->pathHeight;
          rgbx = Hp->rgbx; bmode = Hp->mode; startX = Hp->startX; pathWidth = Hp->pathWidth;
          // a fill whose columns [startX, startX + pathWidth) (+-1: an edge pixel can round one past the bounds)
          // miss this warp's tile has nothing to do here — except MaskBlend, which clears what it does not cover
          act = bmode == MaskBlend || (startX - 1 < c.tx1 && startX + pathWidth + 1 > c.tx0);
          if (act && y >= sY && y < pH) {
            const JobHdr hd = A.jobs[A.fillJobBase[fl] + (y - sY)];
            kind = hd.kind; n = hd.n; pa = hd.pa; pb = hd.pb;
            // the plan's pixels [pa, pb) miss this warp's columns, or there is no plan: nothing to do here
            if (bmode != MaskBlend && (kind < PlanAligned || pb <= c.tx0 || pa >= c.tx1)) act = false;
            pay = job_payload(A, Hp, y, nullptr);
            wrapRows = Hp->wrapRows;
            rowsBelow = pH - 1 - y;
          } else if (bmode != MaskBlend) {
            act = false;
          }
        }
      }
      unsigned todo = __ballot_sync(0xffffffffu, act);
#pragma unroll 1
      while (todo) {  // ascending order = the reference's sequential order of fills
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int k_ = __shfl_sync(0xffffffffu, kind, src), n_ = __shfl_sync(0xffffffffu, n, src);
        const int pa_ = __shfl_sync(0xffffffffu, pa, src), pb_ = __shfl_sync(0xffffffffu, pb, src);
        const int sx_ = __shfl_sync(0xffffffffu, startX, src), pw_ = __shfl_sync(0xffffffffu, pathWidth, src);
        const int md_ = __shfl_sync(0xffffffffu, bmode, src);
        const px_t col_ = __shfl_sync(0xffffffffu, rgbx, src);
        const uint2* pay_ = reinterpret_cast<const uint2*>(__shfl_sync(0xffffffffu, (unsigned long long)pay, src));
        c.mode = md_;
        if (md_ == NormalBlend) apply_row<NormalBlend>(c, col_, sx_, pw_, k_, n_, pa_, pb_, pay_);
        else if (md_ == OverwriteBlend) apply_row<OverwriteBlend>(c, col_, sx_, pw_, k_, n_, pa_, pb_, pay_);
        else if (md_ == MaskBlend) apply_row<MaskBlend>(c, col_, sx_, pw_, k_, n_, pa_, pb_, pay_);
        else apply_row<GenericMode>(c, col_, sx_, pw_, k_, n_, pa_, pb_, pay_);
        __syncwarp();
        const int wr_ = __shfl_sync(0xffffffffu, wrapRows, src);
        if (wr_ > 0) {  // only MaskBlend fills reaching left of the canvas
          const int below = min(wr_, __shfl_sync(0xffffffffu, rowsBelow, src));
          const FillHeader* Hp = A.fills + (fb + src);
          const int jb = A.fillJobBase[fb + src] - Hp->startY;
#pragma unroll 1
          for (int yy = y + 1; yy <= y + below; yy++) {
            const JobHdr hd = A.jobs[jb + yy];
            if (hd.kind == PlanTrapezoids) mask_wrap_clears(c, yy, hd.n, job_payload(A, Hp, yy, nullptr));
          }
          __syncwarp();
        }
      }
    }
    covered += c.covered;
#ifdef PIXIE_RASTER_TIMING  // tools/build_variant.sh tk pixie_b200/csrc/cuda/raster.cu -DPIXIE_RASTER_TIMING
    if (lane == 0) {
      const unsigned long long dt = (unsigned long long)(clock64() - tk0_);
      atomicMax(&A.counters[40], dt);
      atomicAdd(&A.counters[41], dt);
      atomicAdd(&A.counters[42], 1ull);
      if (dt > 100000ull) atomicAdd(&A.counters[43], 1ull);
      if (dt > 50000ull) atomicAdd(&A.counters[44], 1ull);
      if (dt > 20000ull) atomicAdd(&A.counters[45], 1ull);
    }
#endif
  }
#ifdef PIXIE_RASTER_TIMING
  if (lane == 0) atomicMax(&A.counters[46], (unsigned long long)(clock64() - kern0_));
#endif
  if (A.countCovered) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) covered += __shfl_xor_sync(0xffffffffu, covered, o);
    if (lane == 0 && covered) atomicAdd(&A.counters[1], (unsigned long long)covered);
  }
}

// K1c: payload offsets.  Band gp owns rows(gp) * entries(gp) entry-rows of plan payload; exclusive scan
// over the bands (one block), total -> meta[0].
__global__ void __launch_bounds__(1024) payload_scan_kernel(const FillHeader* __restrict__ fills, int numFills,
                                                            const int* __restrict__ entryOff, int numParts,
                                                            unsigned* __restrict__ payOff, unsigned long long* __restrict__ meta) {
  __shared__ unsigned long long sums[1024];
  const int tid = threadIdx.x;
  const int per = (numParts + 1023) / 1024;
  const int b = min(tid * per, numParts), e = min(b + per, numParts);
  auto weight = [&](int gp) -> unsigned long long {
    const FillHeader H = fills[find_fill<true>(fills, numFills, gp)];
    const int p = gp - H.partBase;
    const int top = H.startY + p * H.partitionHeight;
    const int bottom = (p == H.numPartitions - 1) ? H.pathHeight : top + H.partitionHeight;
    return (unsigned long long)(bottom - top) * (unsigned long long)(entryOff[gp + 1] - entryOff[gp]);
  };
  unsigned long long sum = 0;
  for (int i = b; i < e; i++) sum += weight(i);
  sums[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const unsigned long long v = tid >= o ? sums[tid - o] : 0;
    __syncthreads();
    sums[tid] += v;
    __syncthreads();
  }
  unsigned long long run = sums[tid] - sum;
  for (int i = b; i < e; i++) {
    payOff[i] = (unsigned)run;
    run += weight(i);
  }
  if (tid == 1023) {
    payOff[numParts] = (unsigned)sums[1023];
    meta[0] = sums[1023];
  }
}

// Rows of every band (static per list: the last band of a fill absorbs the remainder, :1192) — made once at build.
__global__ void __launch_bounds__(256) band_rows_kernel(const FillHeader* __restrict__ fills, int numFills, int numParts,
                                                        int* __restrict__ bandRows) {
  const int gp = blockIdx.x * blockDim.x + threadIdx.x;
  if (gp >= numParts) return;
  const FillHeader H = fills[find_fill<true>(fills, numFills, gp)];
  const int p = gp - H.partBase;
  const int top = H.startY + p * H.partitionHeight;
  const int bottom = (p == H.numPartitions - 1) ? H.pathHeight : top + H.partitionHeight;
  bandRows[gp] = bottom - top;
}

// K1b / K1c as two short multi-block kernels (the single-block scans took 1 ms on a 118 000-band icon batch):
// band_sum_kernel: per chunk of kScanChunk bands the sums of entries and of payload entry-rows and the largest
// count; the last block to finish turns the chunk sums into exclusive prefixes and leaves {entries, max, payload}
// in meta.  band_scan_kernel: each block rescans its chunk from its prefix: counts -> entry offsets (in place),
// payload offsets.
constexpr int kScanChunk = 2048;  // 256 threads x 8 bands
__global__ void __launch_bounds__(256) band_sum_kernel(const int* __restrict__ cnt, const int* __restrict__ bandRows, int n,
                                                       unsigned long long* __restrict__ chunkCnt, unsigned long long* __restrict__ chunkPay,
                                                       unsigned* __restrict__ ticket, unsigned long long* __restrict__ meta, int numChunks) {
  __shared__ unsigned long long sc[8], sp[8];
  __shared__ int sm[8];
  __shared__ bool last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base = blockIdx.x * kScanChunk;
  unsigned long long c = 0, pay = 0;
  int mx = 0;
  for (int i = base + tid; i < min(base + kScanChunk, n); i += 256) {
    const int v = cnt[i];
    c += (unsigned long long)v;
    pay += (unsigned long long)v * (unsigned long long)bandRows[i];
    mx = max(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, o);
    pay += __shfl_xor_sync(0xffffffffu, pay, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { sc[warp] = c; sp[warp] = pay; sm[warp] = mx; }
  __syncthreads();
  if (tid == 0) {
    for (int k = 1; k < 8; k++) { c += sc[k]; pay += sp[k]; mx = max(mx, sm[k]); }
    chunkCnt[blockIdx.x] = c;
    chunkPay[blockIdx.x] = pay;
    atomicMax(reinterpret_cast<unsigned long long*>(&meta[1]), (unsigned long long)mx);
    __threadfence();
    last = atomicAdd(ticket, 1u) == (unsigned)numChunks - 1u;
  }
  __syncthreads();
  if (last && warp == 0) {  // exclusive prefixes of the chunk sums, 32 chunks per step
    __threadfence();
    unsigned long long runC = 0, runP = 0;
    for (int b0 = 0; b0 < numChunks; b0 += 32) {
      const int k = b0 + lane;
      const unsigned long long vc = k < numChunks ? chunkCnt[k] : 0, vp = k < numChunks ? chunkPay[k] : 0;
      unsigned long long ic = vc, ip = vp;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long tc = __shfl_up_sync(0xffffffffu, ic, o), tp = __shfl_up_sync(0xffffffffu, ip, o);
        if (lane >= o) { ic += tc; ip += tp; }
      }
      if (k < numChunks) { chunkCnt[k] = runC + ic - vc; chunkPay[k] = runP + ip - vp; }
      runC += __shfl_sync(0xffffffffu, ic, 31);
      runP += __shfl_sync(0xffffffffu, ip, 31);
    }
    if (lane == 0) {
      meta[0] = runC;
      meta[2] = runP;
      *ticket = 0u;
    }
  }
}

__global__ void __launch_bounds__(256) band_scan_kernel(int* __restrict__ cnt, const int* __restrict__ bandRows, unsigned* __restrict__ payOff,
                                                        int n, const unsigned long long* __restrict__ chunkCnt,
                                                        const unsigned long long* __restrict__ chunkPay, const unsigned long long* __restrict__ meta) {
  __shared__ unsigned long long wc[8], wp[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * kScanChunk + tid * 8;  // 8 consecutive bands per thread
  int v[8];
  unsigned long long w[8], c = 0, pay = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    v[k] = i0 + k < n ? cnt[i0 + k] : 0;
    w[k] = i0 + k < n ? (unsigned long long)v[k] * (unsigned long long)bandRows[i0 + k] : 0;
    c += (unsigned long long)v[k];
    pay += w[k];
  }
  unsigned long long ic = c, ip = pay;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long tc = __shfl_up_sync(0xffffffffu, ic, o), tp = __shfl_up_sync(0xffffffffu, ip, o);
    if (lane >= o) { ic += tc; ip += tp; }
  }
  if (lane == 31) { wc[warp] = ic; wp[warp] = ip; }
  __syncthreads();
  unsigned long long runC = chunkCnt[blockIdx.x] + ic - c, runP = chunkPay[blockIdx.x] + ip - pay;
  for (int k = 0; k < warp; k++) { runC += wc[k]; runP += wp[k]; }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (i0 + k < n) {
      cnt[i0 + k] = (int)runC;
      payOff[i0 + k] = (unsigned)runP;
    }
    runC += (unsigned long long)v[k];
    runP += w[k];
  }
  if (blockIdx.x == 0 && tid == 0) {
    cnt[n] = (int)meta[0];
    payOff[n] = (unsigned)meta[2];
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

// Row bands of pixie_cuda_render_batch_host.  The copy engine can start once the first band is rasterised and
// is the bottleneck from then on, so the first bands are short (their latency is what the whole call waits
// for) and the rest share the canvas evenly: edges in 32nds of the height.
static int band_edge(long long rows, int b, int K) {
  static const int edges8[9] = {0, 1, 3, 7, 12, 17, 22, 27, 32};
  if (K == Runtime::kBands && K == 8) return (int)(rows * edges8[b] / 32);
  return (int)(rows * b / K);
}

static void free_list(CmdList& L) {
  if (L.rowsJobBase) cudaFree(L.rowsJobBase);
  L.rowsJobBase = nullptr;
  // owned blocks come from the stream-ordered pool (its memory stays cached: no driver allocation per list)
  if (L.owned && L.block) cudaFreeAsync(L.block, rt().stream);
  if (L.owned && L.blockB) cudaFreeAsync(L.blockB, rt().stream);
  L.block = L.blockB = nullptr;
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static const bool g_trace = getenv("PIXIE_CUDA_TRACE") != nullptr;

// The count pass of partitionSegments (:1201-1223) on the device: band counts per segment range, exclusive scan to
// entry offsets, payload offsets.  Totals land in counters[2..4].  Run by build_list (to size block B) and again
// at the start of every later run of a resident list, so that one run = the whole of partitionSegments.
static int device_count(CmdList& L) {
  Runtime& r = rt();
  const size_t P = (size_t)L.numParts;
  PX_CUDA(cudaMemsetAsync(L.entryOff, 0, (P + 1) * 4, r.stream));
  const int cblocks = (int)std::min<int64_t>((L.numSegs + 255) / 256, (int64_t)r.num_sms * 8);
  count_kernel<<<std::max(cblocks, 1), 256, 0, r.stream>>>(L.fills, L.numFills, L.segs, (int)L.numSegs, L.entryOff, L.ranges, L.groupRange);
  PX_LAUNCHED();
  const int numChunks = (int)((P + kScanChunk - 1) / kScanChunk);
  // {entries, max, payload} and the scan's block ticket — counters[5] doubles as band 1's row ticket of a banded run,
  // which leaves it non-zero
  PX_CUDA(cudaMemsetAsync(L.counters + 2, 0, 32, r.stream));
  band_sum_kernel<<<numChunks, 256, 0, r.stream>>>(L.entryOff, L.bandRows, (int)P, L.chunkSums, L.chunkSums + numChunks,
                                                   reinterpret_cast<unsigned*>(L.counters + 5), L.counters + 2, numChunks);
  PX_LAUNCHED();
  band_scan_kernel<<<numChunks, 256, 0, r.stream>>>(L.entryOff, L.bandRows, L.payOff, (int)P, L.chunkSums, L.chunkSums + numChunks, L.counters + 2);
  PX_LAUNCHED();
  return 0;
}

constexpr int64_t kHostCountMaxSegs = 8192;  // lists up to this size are counted on the host (no sync in the call)

// `dev` != null: the segments are already in HBM (flatten.cu) together with the bounds of every path; seg / wind are
// null then and segOff = dev->segBegin.
static int build_list(CmdList& L, bool arena, int bands, int w, int h, int layers, int numFills, const int32_t* layerOf,
                      const float* seg, const int16_t* wind, const int32_t* segOff, const uint32_t* rgbx,
                      const uint8_t* rule, const uint8_t* mode, const FlattenedPaths* dev = nullptr) {
  Runtime& r = rt();
  const double t0 = now_ms();
  if (w <= 0 || h <= 0 || layers <= 0) return fail_pixie("Image width and height must be > 0");
  if (numFills < 0) return fail_pixie("negative fill count");
  L.w = w; L.h = h; L.layers = layers; L.numFills = numFills;
  std::vector<FillHeader> fills(numFills);
  std::vector<int> layerBegin(layers + 1, 0);
  int64_t numPartsTotal = 0, jobsTotal = 0;  // jobs: (fill, scanline) pairs inside the paths' rows
  std::vector<int> jobBase(numFills + 1, 0);
  int prevLayer = 0;
  double tBounds = 0;
  const int64_t numSegs = numFills ? segOff[numFills] : 0;
  if (numSegs < 0 || numSegs > 0x7fffffffll) return fail_pixie("invalid seg_offsets");
  L.numSegs = numSegs;
  L.bands = (bands > 1 && layers == 1) ? std::min(bands, (int)Runtime::kBands) : 1;
  if (!arena) {
    L.hostStartY.assign(numFills, 0);
    L.hostPathHeight.assign(numFills, 0);
  }

  // Device block A.  The host-written part (segments, windings, headers, row ranges, layer table) is
  // contiguous and goes through a staging buffer; the segments are written into it by the same pass that
  // computes the bounds of each path, and start their way to the device before the headers are finished.
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t off = 0;
  const size_t oSegs = off;      off = al(off + (size_t)numSegs * 16);
  const size_t oWind = off;      off = al(off + (size_t)numSegs * 2);
  const size_t oFills = off;     off = al(off + fills.size() * sizeof(FillHeader));
  const size_t oRowRange = off;  off = al(off + fills.size() * sizeof(int2));
  const size_t oLayer = off;     off = al(off + layerBegin.size() * 4);
  const size_t oJobBase = off;   off = al(off + jobBase.size() * 4);
  const size_t oBandJobs = off;  off = al(off + (L.bands > 1 ? (size_t)L.bands * jobBase.size() * 4 : 0));
  // Longest row first (single canvases of 256..65536 rows with at least 16 fills): the raster kernel's persistent warps
  // take (row, tile) tickets from a counter, and the rows of a drawing differ tenfold in work — handed out top to
  // bottom, the busy middle rows of the tiger start late and the last of them run alone (12 % of the kernel's warp
  // slots idle under ncu).  The order comes from a per-row estimate made here from the fill headers, once per list.
  const bool lpt = layers == 1 && h >= 256 && h <= 65536 && numFills >= 16 && L.bands <= 1;
  const size_t oRowOrder = off;  off = al(off + (lpt ? (size_t)h * 4 : 0));
  const size_t h2dBytes = off;
  // Small lists (a single fillPath, a glyph, an icon): the band counts, their scans and the packed band ranges are
  // made on the host while it stages the segments anyway, so the call needs no device round trip before it can
  // size block B — fill_segments / fill_batch then only enqueue work (the reference's callers issue thousands of
  // small fills; a synchronisation per call would cost more than the fill).
  const bool hostCount = arena && numSegs <= kHostCountMaxSegs && !dev;
  size_t stageBytes = h2dBytes;
  if (hostCount) {
    const size_t pMax = (size_t)numSegs / 2 + (size_t)numFills + 1;
    stageBytes += 2 * al((pMax + 1) * 4) + al(pMax) + al(std::max<size_t>(1, (size_t)numSegs) * 4) + al(((size_t)numSegs / 32 + 2) * 4);
  }
  // the host-written part is assembled in the library's pinned staging buffer (grown on demand, reused): the H2D
  // copies run at PCIe speed and nothing is allocated or page-faulted per list
  uint8_t* stage = nullptr;
  {
    void* pin;
    if (int rc = staging_acquire(stageBytes, &pin)) return rc;
    stage = (uint8_t*)pin;
  }
  for (int k = 0; k < numFills; k++) {
    const int layer = layerOf ? layerOf[k] : 0;
    if (layer < prevLayer || layer >= layers || layer < 0) return fail_pixie("layer_of_fill must be non-decreasing and < layers");
    prevLayer = layer;
    layerBegin[layer + 1] = k + 1;
    if (mode[k] >= NumBlendModes || rule[k] > 1) return fail_pixie("invalid blend mode / winding rule");
    const int s0 = segOff[k], s1 = segOff[k + 1];
    if (s1 < s0 || s0 < 0 || s1 > numSegs) return fail_pixie("seg_offsets must be non-decreasing");
    const int n = s1 - s0;
    FillHeader& H = fills[k];
    memset(&H, 0, sizeof(H));
    jobBase[k] = (int)jobsTotal;
    H.segBegin = s0; H.segCount = n; H.rgbx = rgbx[k]; H.rule = rule[k]; H.mode = mode[k];
    if (n == 0) {  // empty path: nothing is drawn (the reference's tiger has one, "M-65.4,9z")
      H.active = 0;
      H.partBase = (int)numPartsTotal;
      continue;
    }
    // computeBounds (:1098-1117) + snapToPixels (common.nim:92-101) + clip to the image (:1605-1613)
    float xMin = INFINITY, xMax = -INFINITY, yMin = INFINITY, yMax = -INFINITY;
    const double tb0 = g_trace ? now_ms() : 0;
    if (dev) {  // computeBounds was reduced on the device (bounds_kernel)
      const float* b = dev->bounds.data() + 5 * (size_t)k;
      if (b[4] != 0.0f) return fail_pixie("flatten: non-finite path coordinates");
      xMin = b[0]; xMax = b[1]; yMin = b[2]; yMax = b[3];
    } else
    // Nim's min/max (`if x <= y: x else: y`) == MINPS/MAXPS(acc, v) including their NaN behaviour
    // (second operand when unordered); one 4-lane min + max per segment {at.x, at.y, to.x, to.y}, and the
    // segment goes to the staging buffer on the way (streaming stores: the DMA engine is the only reader).
    {
      const float* sp = seg + 4 * (size_t)s0;
      float* dp = reinterpret_cast<float*>(stage + oSegs) + 4 * (size_t)s0;
      __m128 vmin = _mm_set1_ps(INFINITY), vmax = _mm_set1_ps(-INFINITY);
      // four independent min/max chains (the loop is bound by their latency); without NaNs the result does
      // not depend on the order, with a NaN anywhere the path is folded again in the reference's order
      __m128 mn1 = vmin, mn2 = vmin, mn3 = vmin, mx1 = vmax, mx2 = vmax, mx3 = vmax, nan = _mm_setzero_ps();
      int i = 0;
      for (; i + 4 <= n; i += 4, sp += 16, dp += 16) {
        const __m128 v0 = _mm_loadu_ps(sp), v1 = _mm_loadu_ps(sp + 4), v2 = _mm_loadu_ps(sp + 8), v3 = _mm_loadu_ps(sp + 12);
        _mm_stream_ps(dp, v0); _mm_stream_ps(dp + 4, v1); _mm_stream_ps(dp + 8, v2); _mm_stream_ps(dp + 12, v3);
        vmin = _mm_min_ps(vmin, v0); mn1 = _mm_min_ps(mn1, v1); mn2 = _mm_min_ps(mn2, v2); mn3 = _mm_min_ps(mn3, v3);
        vmax = _mm_max_ps(vmax, v0); mx1 = _mm_max_ps(mx1, v1); mx2 = _mm_max_ps(mx2, v2); mx3 = _mm_max_ps(mx3, v3);
        nan = _mm_or_ps(nan, _mm_or_ps(_mm_or_ps(_mm_cmpunord_ps(v0, v0), _mm_cmpunord_ps(v1, v1)),
                                       _mm_or_ps(_mm_cmpunord_ps(v2, v2), _mm_cmpunord_ps(v3, v3))));
      }
      vmin = _mm_min_ps(_mm_min_ps(vmin, mn1), _mm_min_ps(mn2, mn3));
      vmax = _mm_max_ps(_mm_max_ps(vmax, mx1), _mm_max_ps(mx2, mx3));
      for (; i < n; i++, sp += 4, dp += 4) {
        const __m128 v = _mm_loadu_ps(sp);
        _mm_stream_ps(dp, v);
        vmin = _mm_min_ps(vmin, v);
        vmax = _mm_max_ps(vmax, v);
        nan = _mm_or_ps(nan, _mm_cmpunord_ps(v, v));
      }
      if (_mm_movemask_ps(nan)) {
        sp = seg + 4 * (size_t)s0;
        vmin = _mm_set1_ps(INFINITY);
        vmax = _mm_set1_ps(-INFINITY);
        for (i = 0; i < n; i++, sp += 4) {
          const __m128 v = _mm_loadu_ps(sp);
          vmin = _mm_min_ps(vmin, v);
          vmax = _mm_max_ps(vmax, v);
        }
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
    jobsTotal += height;
    L.maxWrapRows = std::max(L.maxWrapRows, H.wrapRows);
    if (!arena) {
      L.hostStartY[k] = H.startY;
      L.hostPathHeight[k] = H.pathHeight;
    }
    if (numPartsTotal > 0x3fffffff || jobsTotal > 0x3fffffff) return fail_pixie("command list too large");
  }
  // A MaskBlend fill reaching left of the canvas makes row y read the plans of rows y + 1 .. y + wrapRows
  // (mask_wrap_clears); with row bands those belong to the NEXT band's plan launch on another stream, which
  // band b's raster kernel does not wait for: such lists are planned and rasterised as one band.
  if (L.maxWrapRows > 0) L.bands = 1;
  for (int l = 1; l <= layers; l++) layerBegin[l] = std::max(layerBegin[l], layerBegin[l - 1]);
  jobBase[numFills] = (int)jobsTotal;
  L.totalJobs = (int)jobsTotal;
  if ((long long)layers * h > 0x7fffffffll) return fail_pixie("canvas stack too large");

  const double t1 = now_ms();
  L.numParts = numPartsTotal;

  // launch geometry: plan_kernel strides over the jobs, raster_kernel's persistent warps take one
  // (layer, row) ticket at a time; both sized to what is resident on the GPU
  // canvas rows are rasterised in tiles of 2048 columns, one warp each: fills that miss a tile are skipped there,
  // which shortens the ordered chain a warp walks, and a row's work runs on several warps at once
  L.tileW = w > 3072 ? 2048 : ((w + 3) & ~3);
  L.tiles = (w + L.tileW - 1) / L.tileW;
  L.covBytes = ((L.tileW + 7) & ~3) + 4 + 2 * kPartCap;  // coverage row of a tile, word aligned, with the covBase slack, + the partial-word list
  L.smemCap = 64;                              // entries per band planned from shared memory
  L.warpsPerBlock = 8;
  while (L.warpsPerBlock > 1 && (size_t)L.covBytes * L.warpsPerBlock > 96 * 1024) L.warpsPerBlock /= 2;
  if ((size_t)L.covBytes * L.warpsPerBlock > 200 * 1024) return fail_pixie("canvas too wide for the shared-memory coverage row");
  L.smemBytes = (size_t)L.covBytes * L.warpsPerBlock;
  L.planSmem = (size_t)L.smemCap * kScratchArrays * 4 * 8;
  const long long totalRows = (long long)layers * h;
  static size_t occSmem = ~(size_t)0;  // occupancy of the two kernels, looked up once per coverage-row size
  static int occRaster = 1, occPlan = 1;
  if (occSmem != L.smemBytes) {
    if (L.smemBytes > 48 * 1024) {
      PX_CUDA(cudaFuncSetAttribute(raster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemBytes));
      PX_CUDA(cudaFuncSetAttribute(raster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemBytes));
    }
    PX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occRaster, raster_kernel<false>, L.warpsPerBlock * 32, L.smemBytes));
    PX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occPlan, plan_kernel, 256, L.planSmem));
    occSmem = L.smemBytes;
  }
  const int blocksPerSm = std::max(1, occRaster), planPerSm = std::max(1, occPlan);
  long long wantBlocks = (totalRows * L.tiles + L.warpsPerBlock - 1) / L.warpsPerBlock;
  L.rasterBlocks = (int)std::max<long long>(1, std::min<long long>(wantBlocks, (long long)r.num_sms * blocksPerSm));
  // plan_kernel takes three blocks per SM, not the four its registers allow: the rest of the register file is
  // plan_light_kernel's, which runs beside it (launch_plan_kernels)
  static const int planCap = getenv("PIXIE_CUDA_PLAN_PER_SM") ? atoi(getenv("PIXIE_CUDA_PLAN_PER_SM")) : 3;
  const int planRes = planCap > 0 ? std::min(planCap, planPerSm) : planPerSm;
  L.planBlocks = (int)std::max<long long>(1, std::min<long long>((jobsTotal + 7) / 8, (long long)r.num_sms * planRes));
  L.scratchSlotCount = r.num_sms * planPerSm;

  // device-made part of block A: band counts, entry offsets, flags and counters
  const size_t P = (size_t)numPartsTotal;
  const size_t oEntryOff = off;  off = al(off + (P + 1) * 4);   // band counts, scanned in place to offsets
  const size_t oFlags = off;     off = al(off + std::max<size_t>(1, P));
  const size_t oPayOff = off;    off = al(off + (P + 1) * 4);   // plan payload offset of each band
  const size_t oRanges = off;    off = al(off + std::max<size_t>(1, (size_t)numSegs) * 4);  // packed band range of each segment
  const size_t oGroups = off;    off = al(off + ((size_t)numSegs / 32 + 2) * 4);           // ... and of each group of 32
  const size_t oSlots = off;     off = al(off + (size_t)((L.scratchSlotCount + 31) / 32) * 4);
  const size_t oCounters = off;  off = al(off + 512);           // [32..39] plan_kernel's job tickets (one per plan launch); [0] row ticket, [1] covered px, [2] entries, [3] max, [4..11] band tickets, [16..31] heavy-job counts (front / back of each launch's list)
  // device-only scratch of the band scans (not part of what a host-counted list copies in)
  const size_t oBandRows = off;  off = al(off + std::max<size_t>(1, P) * 4);
  const size_t oChunks = off;    off = al(off + (2 * ((P + kScanChunk - 1) / kScanChunk) + 2) * 8);
  const size_t totalA = off;
  if (arena) {
    void* blk;
    if (int rc = get_scratch(2, totalA, &blk)) return rc;
    L.block = (uint8_t*)blk;
    L.owned = false;
  } else {
    PX_CUDA(cudaMallocAsync((void**)&L.block, totalA, r.stream));
    L.owned = true;
  }
  if (numSegs && dev) {
    PX_CUDA(cudaMemcpyAsync(L.block + oSegs, dev->segs, (size_t)numSegs * 16, cudaMemcpyDeviceToDevice, r.stream));
    PX_CUDA(cudaMemcpyAsync(L.block + oWind, dev->wind, (size_t)numSegs * 2, cudaMemcpyDeviceToDevice, r.stream));
  } else if (numSegs) {  // segments (staged by the bounds pass) + windings go first
    memcpy(stage + oWind, wind, (size_t)numSegs * 2);
    _mm_sfence();
    PX_CUDA(cudaMemcpyAsync(L.block, stage, oFills, cudaMemcpyHostToDevice, r.stream));
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
  if (lpt) {
    // what a raster warp spends on one (fill, row): a fixed part, the spans (about two band entries per segment and
    // band, capped) and the covered width in 128-pixel vector steps; summed per row through a difference array
    std::vector<long long> diff((size_t)h + 1, 0);
    for (const FillHeader& F : fills) {
      if (!F.active || F.numPartitions <= 0 || F.pathHeight <= F.startY) continue;
      const long long c = 6 + std::min<long long>(64, 2ll * F.segCount / F.numPartitions) + (F.pathWidth >> 7);
      diff[(size_t)F.startY] += c;
      diff[(size_t)F.pathHeight] -= c;
    }
    std::vector<long long> cost((size_t)h);
    long long run = 0, mx = 1;
    for (int y = 0; y < h; y++) {
      run += diff[(size_t)y];
      cost[(size_t)y] = run;
      mx = std::max(mx, run);
    }
    int hist[257] = {0};
    auto bucket = [&](long long c) { return 255 - (int)(c * 255 / mx); };  // 0 = the most expensive rows
    for (int y = 0; y < h; y++) hist[bucket(cost[(size_t)y]) + 1]++;
    for (int b = 0; b < 256; b++) hist[b + 1] += hist[b];
    // How finely a row is cut: the kernel cannot end before its longest ticket does, and on the tiger the busiest
    // half-row took as long as the whole kernel (358 k of 375 k cycles; the mean ticket 119 k).  A row whose ticket
    // is estimated above cutPct % of an even share of the work is cut in two, four, eight; pieces repeat the per-ticket
    // scan of the fills, so cutting everything is slower (jobs that miss a piece cost it nothing: JobHdr::pa / pb).
    long long total = 0;
    for (int y = 0; y < h; y++) total += cost[(size_t)y];
    const long long even = std::max<long long>(1, total / std::max(1, r.num_sms * 24));
    int maxShift = 0;
    while (maxShift < kMaxSubShift && (L.tileW >> (maxShift + 1)) >= 256 && (L.tileW % (8 << maxShift)) == 0) maxShift++;
    std::vector<uint8_t> shiftOf((size_t)h, 0);
    // swept on the tiger (PIXIE_CUDA_CUT = percent): 4096^2 raster 0.195 ms uncut, 0.183 at 80, 0.189 at 90, 0.22 at 40;
    // 2048^2 0.166 uncut (before the jobs carried their extents), 0.131 at 90, 0.142 at 80; 8192^2 unchanged
    static const long long cutPct = getenv("PIXIE_CUDA_CUT") ? atoll(getenv("PIXIE_CUDA_CUT")) : 85;
    int used = 0;
    for (int y = 0; y < h; y++) {
      int s_ = 0;
      while (s_ < maxShift && cost[(size_t)y] / L.tiles > (even << s_) * cutPct / 100) s_++;
      shiftOf[(size_t)y] = (uint8_t)s_;
      used = std::max(used, s_);
    }
    L.subShift = used;
    int* order = reinterpret_cast<int*>(stage + oRowOrder);
    for (int y = 0; y < h; y++) order[hist[bucket(cost[(size_t)y])]++] = y | ((int)shiftOf[(size_t)y] << 28);
  }
  memcpy(stage + oLayer, layerBegin.data(), layerBegin.size() * 4);
  memcpy(stage + oJobBase, jobBase.data(), jobBase.size() * 4);
  if (L.bands > 1) {  // per band: prefix over the fills of the scanlines of each path inside the band
    int* bj = reinterpret_cast<int*>(stage + oBandJobs);
    for (int b = 0; b < L.bands; b++, bj += jobBase.size()) {
      const int y0 = band_edge(h, b, L.bands), y1 = band_edge(h, b + 1, L.bands);
      int run = 0;
      for (int k = 0; k < numFills; k++) {
        const FillHeader& F = fills[k];
        bj[k] = run;
        if (F.active && F.numPartitions > 0) run += std::max(0, std::min(F.pathHeight, y1) - std::max(F.startY, y0));
      }
      bj[numFills] = run;
      L.bandJobs[b] = run;
    }
  }
  long long meta[4] = {0, 2, 0, 0};  // entries, max entries per band, payload entry-rows
  size_t copyEnd = h2dBytes;
  if (hostCount && P > 0) {
    int* cnt = reinterpret_cast<int*>(stage + oEntryOff);
    unsigned* pOff = reinterpret_cast<unsigned*>(stage + oPayOff);
    uint32_t* rng = reinterpret_cast<uint32_t*>(stage + oRanges);
    uint32_t* grp = reinterpret_cast<uint32_t*>(stage + oGroups);
    memset(cnt, 0, (P + 1) * 4);
    for (int64_t i = 0; i < numSegs; i++) rng[i] = kNoBand;
    for (int k = 0; k < numFills; k++) {  // count_kernel on the host: partitionRange (:1201-1213) per segment
      const FillHeader& F = fills[k];
      if (!F.active || F.numPartitions <= 0) continue;
      const float startYf = (float)(unsigned)F.startY;
      const unsigned ph = (unsigned)F.partitionHeight, lastP = (unsigned)(F.numPartitions - 1);
      const float* sp = seg + 4 * (size_t)F.segBegin;
      for (int i = 0; i < F.segCount; i++, sp += 4) {
        unsigned atP = 0, toP = 0;
        if (F.numPartitions > 1) {
          atP = std::min(f2u_host(fmaxf(0.0f, sp[1] - startYf)) / ph, lastP);
          toP = std::min(f2u_host(fmaxf(0.0f, sp[3] - startYf)) / ph, lastP);
        }
        rng[F.segBegin + i] = F.numPartitions <= kMaxPackedBands ? (atP | (toP << 16)) : kNoBand;
        for (unsigned p_ = atP; p_ <= toP; p_++) cnt[F.partBase + (int)p_]++;
      }
    }
    {  // group summaries (count_kernel's warp reduction): uniform groups of one plain fill get lo | hi << 16
      int k = 0;
      for (int64_t g0 = 0; g0 * 32 < numSegs; g0++) {
        const int64_t b = g0 * 32, e = b + 32;
        uint32_t out = 0xFFFF0000u;
        while (k < numFills && (int64_t)fills[k].segBegin + fills[k].segCount <= b) k++;
        if (e <= numSegs && k < numFills) {
          const FillHeader& F = fills[k];
          if (F.active && F.numPartitions > 0 && F.numPartitions <= kMaxPackedBands && F.segBegin <= b && (int64_t)F.segBegin + F.segCount >= e) {
            uint32_t lo = 0xFFFFu, hi = 0u;
            for (int64_t i = b; i < e; i++) {
              lo = std::min(lo, rng[i] & 0xFFFFu);
              hi = std::max(hi, rng[i] >> 16);
            }
            out = lo | (hi << 16);
          }
        }
        grp[g0] = out;
      }
    }
    long long run = 0, mx = 0, pay = 0;
    {  // scan_kernel + payload_scan_kernel on the host
      for (int k = 0; k < numFills; k++) {
        const FillHeader& F = fills[k];
        if (!F.active || F.numPartitions <= 0) continue;
        for (int p_ = 0; p_ < F.numPartitions; p_++) {
          const size_t gp = (size_t)F.partBase + (size_t)p_;
          const int c = cnt[gp];
          const int top = F.startY + p_ * F.partitionHeight;
          const int bottom = (p_ == F.numPartitions - 1) ? F.pathHeight : top + F.partitionHeight;
          cnt[gp] = (int)run;
          pOff[gp] = (unsigned)pay;
          run += c;
          mx = std::max<long long>(mx, c);
          pay += (long long)(bottom - top) * c;
        }
      }
      cnt[P] = (int)run;
      pOff[P] = (unsigned)pay;
    }
    meta[0] = run; meta[1] = mx; meta[2] = pay;
    copyEnd = oSlots;  // entry offsets, (flags), payload offsets, ranges and group summaries travel with the headers
  }
  PX_CUDA(cudaMemcpyAsync(L.block + oFills, stage + oFills, copyEnd - oFills, cudaMemcpyHostToDevice, r.stream));
  if (int rc = staging_release()) return rc;
  L.segs = (float4*)(L.block + oSegs);
  L.wind = (int16_t*)(L.block + oWind);
  L.fills = (FillHeader*)(L.block + oFills);
  L.rowRange = (int2*)(L.block + oRowRange);
  L.layerFillBegin = (int*)(L.block + oLayer);
  L.fillJobBase = (int*)(L.block + oJobBase);
  L.bandJobBase = L.bands > 1 ? (int*)(L.block + oBandJobs) : nullptr;
  L.scratchSlots = (unsigned*)(L.block + oSlots);
  L.payOff = (unsigned*)(L.block + oPayOff);
  L.entryOff = (int*)(L.block + oEntryOff);
  L.flags = L.block + oFlags;
  L.ranges = (uint32_t*)(L.block + oRanges);
  L.groupRange = (uint32_t*)(L.block + oGroups);
  L.counters = (unsigned long long*)(L.block + oCounters);
  L.bandRows = (int*)(L.block + oBandRows);
  L.rowOrder = lpt ? (int*)(L.block + oRowOrder) : nullptr;
  L.chunkSums = (unsigned long long*)(L.block + oChunks);
  L.h2dBytes = h2dBytes;

  // K1a/K1b on the device (lists too large to count on the host): how many entries each band gets
  // (partitionRange :1201-1213), exclusive scan to entry offsets, total and maximum -> 24 bytes back to the host
  // to size block B.
  PX_CUDA(cudaMemsetAsync(L.scratchSlots, 0, (size_t)((L.scratchSlotCount + 31) / 32) * 4, r.stream));
  if (!hostCount) {
    if (P > 0) {
      L.numParts = numPartsTotal;
      PX_CUDA(cudaMemsetAsync(L.counters, 0, 512, r.stream));
      band_rows_kernel<<<(int)((P + 255) / 256), 256, 0, r.stream>>>(L.fills, numFills, (int)P, L.bandRows);
      PX_LAUNCHED();
      if (int rc = device_count(L)) return rc;
      L.deviceCounted = true;
      L.countFresh = true;
      PX_CUDA(cudaMemcpyAsync(meta, L.counters + 2, 24, cudaMemcpyDeviceToHost, r.stream));
    }
    PX_CUDA(cudaStreamSynchronize(r.stream));  // also retires the pageable staging vector of owned lists
  }
  if (meta[0] > 0x7fffffffll) return fail_pixie("command list too large");
  L.numEntries = meta[0];
  L.maxEntries = (int)std::max<long long>(2, meta[1]);
  L.scratchWords = L.maxEntries > L.smemCap ? L.maxEntries * kScratchArrays : 0;
  if (meta[2] > 0xffffffffll) return fail_pixie("command list too large");

  // Device block B: band entries, the plans (header per job + payload per entry-row) and plan_kernel's
  // per-warp spill scratch for bands with more entries than fit in shared memory.
  const size_t entriesBytes = al(std::max<size_t>(1, (size_t)L.numEntries) * sizeof(Entry));
  const size_t jobsBytes = al(std::max<size_t>(1, (size_t)L.totalJobs) * sizeof(JobHdr));
  const size_t payBytes = al(std::max<size_t>(1, (size_t)meta[2]) * kPaySlots * sizeof(uint2));
  const size_t heavyBytes = al(std::max<size_t>(1, (size_t)L.totalJobs) * sizeof(int));
  const size_t totalB = entriesBytes + jobsBytes + payBytes + 3 * heavyBytes + al((size_t)L.scratchSlotCount * 8 * L.scratchWords * 4);
  if (arena) {
    void* blk;
    if (int rc = get_scratch(4, totalB, &blk)) return rc;
    L.blockB = (uint8_t*)blk;
  } else {
    PX_CUDA(cudaMallocAsync((void**)&L.blockB, totalB, r.stream));
  }
  L.entries = (Entry*)L.blockB;
  L.jobs = (JobHdr*)(L.blockB + entriesBytes);
  L.payload = (uint2*)(L.blockB + entriesBytes + jobsBytes);
  L.heavyList = (int*)(L.blockB + entriesBytes + jobsBytes + payBytes);
  L.splitArrive = (int*)(L.blockB + entriesBytes + jobsBytes + payBytes + heavyBytes);
  PX_CUDA(cudaMemsetAsync(L.splitArrive, 0, heavyBytes, r.stream));
  L.monsterList = (int*)(L.blockB + entriesBytes + jobsBytes + payBytes + 2 * heavyBytes);
  L.scratch = L.scratchWords ? (uint32_t*)(L.blockB + entriesBytes + jobsBytes + payBytes + 3 * heavyBytes) : nullptr;
  if (g_trace)
    fprintf(stderr, "[pixie_cuda] build_list: host plan %.3f ms (bounds %.3f), stage + device count/scan + readback %.3f ms "
            "(%zu B staged, blocks %zu + %zu B, %lld entries, max %d per band)\n",
            t1 - t0, tBounds, now_ms() - t1, h2dBytes, totalA, totalB, (long long)L.numEntries, L.maxEntries);
  return 0;
}

// K2 on stream `st`: classify, then the thread-per-job kernel on an auxiliary stream next to the warp-per-job one
// (both are latency bound and leave most of the machine idle on their own)
static int launch_plan_kernels(const CmdList& L, const RasterArgs& A, int jobs, int heavyBlocks, cudaStream_t st, int auxIndex) {
  Runtime& r = rt();
  if (jobs <= 0) return 0;
  if (!r.aux_stream[auxIndex]) {
    PX_CUDA(cudaStreamCreateWithFlags(&r.aux_stream[auxIndex], cudaStreamNonBlocking));
    PX_CUDA(cudaEventCreateWithFlags(&r.aux_fork[auxIndex], cudaEventDisableTiming));
    PX_CUDA(cudaEventCreateWithFlags(&r.aux_join[auxIndex], cudaEventDisableTiming));
  }
  cudaStream_t aux = r.aux_stream[auxIndex];
  static bool carve = false;
  if (!carve && !getenv("PIXIE_CUDA_NO_CARVE")) {
    // the two plan kernels are meant to share the SMs: with the shared-memory / L1 split the driver picks for the first
    // of them alone, the blocks of the second do not fit until the first ones leave (measured: plan_light_kernel started
    // 40 us after plan_kernel, when its blocks began to exit)
    PX_CUDA(cudaFuncSetAttribute(plan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    PX_CUDA(cudaFuncSetAttribute(plan_light_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carve = true;
  }
  plan_classify_kernel<<<(jobs + 255) / 256, 256, 0, st>>>(A);
  PX_LAUNCHED();
  PX_CUDA(cudaEventRecord(r.aux_fork[auxIndex], st));
  PX_CUDA(cudaStreamWaitEvent(aux, r.aux_fork[auxIndex], 0));
  plan_light_kernel<<<(jobs + kLightThreads - 1) / kLightThreads, kLightThreads, 0, aux>>>(A);
  PX_LAUNCHED();
  PX_CUDA(cudaEventRecord(r.aux_join[auxIndex], aux));
  plan_kernel<<<heavyBlocks, 256, L.planSmem, st>>>(A);
  PX_LAUNCHED();
  PX_CUDA(cudaStreamWaitEvent(st, r.aux_join[auxIndex], 0));
  return 0;
}

// rows [rowY0, rowY1) of a single-layer list (rowY1 < 0: everything).  `im` is the whole canvas, or — bandImage —
// an image of exactly those rows (one GPU's band of a canvas split across GPUs).
static int run_list(CmdList& L, Image* im, uint64_t* covered_px, uint8_t* host_pixels = nullptr, int rowY0 = 0, int rowY1 = -1,
                    bool bandImage = false) {
  Runtime& r = rt();
  const bool rows = rowY1 >= 0;
  if (!rows) {
    if (im->bpp != 4 || im->w != L.w || im->h != L.h || im->layers != L.layers)
      return fail_pixie("command list was built for a different canvas shape");
  }
  if (L.numFills == 0 || (rows && rowY1 <= rowY0)) {
    if (covered_px) *covered_px = 0;
    return 0;
  }
  if (L.deviceCounted && !L.countFresh && L.numParts > 0) {  // the count pass belongs to every run (see device_count)
    if (int rc = device_count(L)) return rc;
  }
  L.countFresh = false;
  r.prof_bands = 0;
  PX_CUDA(cudaMemsetAsync(L.counters, 0, 512, r.stream));  // row tickets, covered px, heavy-job counts, plan tickets
  if (L.numParts > 0) {
    const int warps = (int)std::min<int64_t>(L.numParts, (int64_t)r.num_sms * 32);
    const int blocks = (warps + 7) / 8;
    {
      ProfScope ps(kProfPartition);
      partition_kernel<<<blocks, 256, 0, r.stream>>>(L.fills, L.numFills, L.entryOff, L.segs, L.wind, L.ranges, L.groupRange, L.entries,
                                                     L.flags, (int)L.numParts);
    }
    PX_LAUNCHED();
  }
  RasterArgs A;
  A.canvas = (px_t*)im->data;
  A.w = L.w; A.h = L.h; A.layers = L.layers;
  A.fills = L.fills; A.rowRange = L.rowRange; A.layerFillBegin = L.layerFillBegin; A.entryOff = L.entryOff; A.entries = L.entries;
  A.flags = L.flags; A.gscratch = L.scratch; A.counters = L.counters;
  A.smemCap = L.smemCap; A.scratchCap = L.maxEntries; A.covBytes = L.covBytes; A.tileW = L.tileW; A.tiles = L.tiles;
  A.countCovered = covered_px ? 1 : 0;
  A.numFills = L.numFills;
  A.fillJobBase = L.fillJobBase; A.payOff = L.payOff; A.jobs = L.jobs; A.payload = L.payload; A.totalJobs = L.totalJobs;
  A.splitArrive = L.splitArrive;
  A.rowBegin = 0; A.rowEnd = 0; A.ticketSlot = 0; A.clearFirst = L.clearFirst ? 1 : 0; A.subShift = 0;
  A.scratchSlots = L.scratchSlots; A.scratchSlotCount = L.scratchSlotCount;
  A.planJobBase = L.fillJobBase; A.planY0 = 0; A.planJobs = L.totalJobs;
  A.rowOrder = nullptr;
  static size_t configured = 0;
  if (L.smemBytes > 48 * 1024 && configured < L.smemBytes) {
    PX_CUDA(cudaFuncSetAttribute(raster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemBytes));
    PX_CUDA(cudaFuncSetAttribute(raster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smemBytes));
    configured = L.smemBytes;
  }
  const long long totalRows = (long long)L.layers * L.h;
  if (rows) {
    // One GPU's row band of a canvas split across GPUs: the list — bounds, partition boundaries,
    // band entries — is that of the WHOLE canvas (numPartitions / partitionHeight are defined on the whole path
    // height, paths.nim:1172-1192, so every anti-aliasing decision is the one the undivided render makes); only
    // the jobs (fill, scanline) of rows [rowY0, rowY1) are planned and only those rows are rasterised.  Rows
    // below the band whose plans a MaskBlend fill's negative-x clears read (mask_wrap_clears) are planned too.
    const int planY1 = std::min(L.h, rowY1 + L.maxWrapRows);
    const size_t nb = ((size_t)L.numFills + 1) * 4;
    if (!L.rowsJobBase) PX_CUDA(cudaMalloc(&L.rowsJobBase, nb));
    void* pin;
    if (int rc = staging_acquire(nb, &pin)) return rc;
    int* bj = (int*)pin;
    int run = 0;
    for (int k = 0; k < L.numFills; k++) {
      bj[k] = run;
      run += std::max(0, std::min(L.hostPathHeight[k], planY1) - std::max(L.hostStartY[k], rowY0));
    }
    bj[L.numFills] = run;
    PX_CUDA(cudaMemcpyAsync(L.rowsJobBase, pin, nb, cudaMemcpyHostToDevice, r.stream));
    if (int rc = staging_release()) return rc;
    if (run > 0) {
      ProfScope ps(kProfPlan);
      A.planJobBase = L.rowsJobBase; A.planY0 = rowY0; A.planJobs = run;
      A.heavyList = L.heavyList;
      A.monsterList = L.monsterList;
      A.heavyCount = L.counters + 16;
      const int blocks = std::max(1, std::min((run + 7) / 8, L.planBlocks));
      if (int rc = launch_plan_kernels(L, A, run, blocks, r.stream, 0)) return rc;
    }
    if (bandImage) A.canvas = (px_t*)im->data - (size_t)rowY0 * (size_t)L.w;  // row y of the canvas = row y - rowY0 of the band
    A.rowBegin = rowY0; A.rowEnd = rowY1; A.ticketSlot = 0;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(((long long)(rowY1 - rowY0) * L.tiles + L.warpsPerBlock - 1) / L.warpsPerBlock,
                                                                     L.rasterBlocks));
    {
      ProfScope ps(kProfRaster);
      raster_kernel<false><<<blocks, L.warpsPerBlock * 32, L.smemBytes, r.stream>>>(A);
    }
    PX_LAUNCHED();
  } else if (L.bands <= 1 || L.serial) {
    if (L.totalJobs > 0) {
      ProfScope ps(kProfPlan);
      A.heavyList = L.heavyList;
      A.monsterList = L.monsterList;
      A.heavyCount = L.counters + 16;
      if (int rc = launch_plan_kernels(L, A, L.totalJobs, L.planBlocks, r.stream, 0)) return rc;
    }
    A.rowBegin = 0; A.rowEnd = totalRows; A.ticketSlot = 0;
    A.rowOrder = L.bands <= 1 ? L.rowOrder : nullptr;
    A.subShift = A.rowOrder ? L.subShift : 0;
    {
      ProfScope ps(kProfRaster);
      if (A.subShift > 0) raster_kernel<true><<<L.rasterBlocks, L.warpsPerBlock * 32, L.smemBytes, r.stream>>>(A);
      else raster_kernel<false><<<L.rasterBlocks, L.warpsPerBlock * 32, L.smemBytes, r.stream>>>(A);
    }
    PX_LAUNCHED();
#ifdef PIXIE_RASTER_TIMING  // per-ticket durations: the kernel cannot end before its longest (row, tile) ticket does
    {
      unsigned long long hc[64];
      cudaMemcpyAsync(hc, L.counters, 512, cudaMemcpyDeviceToHost, r.stream);
      cudaStreamSynchronize(r.stream);
      fprintf(stderr, "[plan kernels, ns] heavy: start 0 end %llu | light: start %lld end %lld\n", hc[51] - ~hc[50], (long long)(~hc[52] - ~hc[50]),
              (long long)(hc[53] - ~hc[50]));
      fprintf(stderr, "[raster tickets] n %llu  mean %.0f cycles  max %llu  >100k %llu  >50k %llu  >20k %llu  longest warp lifetime %llu cycles\n", hc[42],
              hc[42] ? (double)hc[41] / (double)hc[42] : 0.0, hc[40], hc[43], hc[44], hc[45], hc[46]);
    }
#endif
    if (host_pixels)
      PX_CUDA(cudaMemcpyAsync(host_pixels, im->data, (size_t)totalRows * L.w * 4, cudaMemcpyDeviceToHost, r.stream));
  } else {
    // Row bands on concurrent streams: band b is planned and rasterised while band b-1 is on its way to the
    // host, so the copy engine starts after 1/bands of the work and stays busy from then on.  Kernels are
    // enqueued in pipeline order (plan b+1 before raster b) with full-size grids; the hardware runs them
    // roughly in that order and fills the GPU with whatever is next.
    const int K = L.bands;
    if (!r.band_stream[0]) {
      int least = 0, greatest = 0;  // earlier bands first: their pixels are the next ones the copy engine needs
      PX_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      for (int b = 0; b < Runtime::kBands; b++) {
        PX_CUDA(cudaStreamCreateWithPriority(&r.band_stream[b], cudaStreamNonBlocking, std::min(least, greatest + b)));
        PX_CUDA(cudaEventCreateWithFlags(&r.band_done[b], cudaEventDisableTiming));
      }
      PX_CUDA(cudaEventCreateWithFlags(&r.band_start, cudaEventDisableTiming));
    }
    static cudaEvent_t tev[1 + 3 * Runtime::kBands] = {};  // PIXIE_CUDA_TRACE: band timeline
    if (g_trace && !tev[0])
      for (auto& e : tev) PX_CUDA(cudaEventCreate(&e));
    if (g_trace) PX_CUDA(cudaEventRecord(tev[0], r.stream));
    PX_CUDA(cudaEventRecord(r.band_start, r.stream));
    const size_t rowBytes = (size_t)L.w * 4;
    const size_t fillsP1 = (size_t)L.numFills + 1;
    const bool prof = r.profiling && r.band_prof[0][0];
    if (prof) r.prof_bands = K;
    auto launch_plan = [&](int b) -> int {
      PX_CUDA(cudaStreamWaitEvent(r.band_stream[b], r.band_start, 0));
      if (prof) PX_CUDA(cudaEventRecord(r.band_prof[b][0], r.band_stream[b]));
      if (prof && L.bandJobs[b] <= 0) PX_CUDA(cudaEventRecord(r.band_prof[b][1], r.band_stream[b]));
      if (L.bandJobs[b] <= 0) return 0;
      RasterArgs B = A;
      B.planJobBase = L.bandJobBase + (size_t)b * fillsP1;
      B.planY0 = band_edge(L.h, b, K);
      B.planJobs = L.bandJobs[b];
      int before = 0;
      for (int bb = 0; bb < b; bb++) before += L.bandJobs[bb];
      B.heavyList = L.heavyList + before;
      B.monsterList = L.monsterList + before;
      B.heavyCount = L.counters + 16 + b;
      const int blocks = std::max(1, std::min((L.bandJobs[b] + 7) / 8, L.planBlocks));
      if (int rc = launch_plan_kernels(L, B, L.bandJobs[b], blocks, r.band_stream[b], 1 + b)) return rc;
      if (prof) PX_CUDA(cudaEventRecord(r.band_prof[b][1], r.band_stream[b]));
      if (g_trace) PX_CUDA(cudaEventRecord(tev[1 + 3 * b], r.band_stream[b]));
      return 0;
    };
    auto launch_raster = [&](int b) -> int {
      RasterArgs B = A;
      B.rowBegin = band_edge(totalRows, b, K);
      B.rowEnd = band_edge(totalRows, b + 1, K);
      B.ticketSlot = 4 + b;
      if (prof) PX_CUDA(cudaEventRecord(r.band_prof[b][2], r.band_stream[b]));
      if (prof && B.rowEnd <= B.rowBegin) PX_CUDA(cudaEventRecord(r.band_prof[b][3], r.band_stream[b]));
      if (B.rowEnd <= B.rowBegin) return 0;
      const int blocks = (int)std::max<long long>(1, std::min<long long>(((B.rowEnd - B.rowBegin) * L.tiles + L.warpsPerBlock - 1) / L.warpsPerBlock,
                                                                       L.rasterBlocks));
      raster_kernel<false><<<blocks, L.warpsPerBlock * 32, L.smemBytes, r.band_stream[b]>>>(B);
      PX_LAUNCHED();
      if (prof) PX_CUDA(cudaEventRecord(r.band_prof[b][3], r.band_stream[b]));
      if (g_trace) PX_CUDA(cudaEventRecord(tev[2 + 3 * b], r.band_stream[b]));
      if (host_pixels)
        PX_CUDA(cudaMemcpyAsync(host_pixels + (size_t)B.rowBegin * rowBytes, im->data + (size_t)B.rowBegin * rowBytes,
                                (size_t)(B.rowEnd - B.rowBegin) * rowBytes, cudaMemcpyDeviceToHost, r.band_stream[b]));
      if (g_trace) PX_CUDA(cudaEventRecord(tev[3 + 3 * b], r.band_stream[b]));
      PX_CUDA(cudaEventRecord(r.band_done[b], r.band_stream[b]));
      PX_CUDA(cudaStreamWaitEvent(r.stream, r.band_done[b], 0));
      return 0;
    };
    if (int rc = launch_plan(0)) return rc;
    for (int b = 0; b < K; b++) {
      if (b + 1 < K)
        if (int rc = launch_plan(b + 1)) return rc;
      if (int rc = launch_raster(b)) return rc;
    }
    if (g_trace) {
      PX_CUDA(cudaStreamSynchronize(r.stream));
      fprintf(stderr, "[pixie_cuda] bands (ms after the prologue; plan / raster / on host):");
      for (int b = 0; b < K; b++) {
        float t[3] = {0, 0, 0};
        for (int k = 0; k < 3; k++)
          if (cudaEventQuery(tev[1 + 3 * b + k]) == cudaSuccess) cudaEventElapsedTime(&t[k], tev[0], tev[1 + 3 * b + k]);
        fprintf(stderr, " [%d] %.3f %.3f %.3f", b, t[0], t[1], t[2]);
      }
      fprintf(stderr, "\n");
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
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  CmdList L;
  // measured on the tiger at 4096^2: 0.465 ms as one band, 0.515 ms in 4 bands (four launch sets, and the raster kernels
  // of neighbouring bands compete for the same SMs) — so one band is the default; PIXIE_CUDA_BANDS is kept for experiments
  static const int residentBands = getenv("PIXIE_CUDA_BANDS") ? atoi(getenv("PIXIE_CUDA_BANDS")) : 1;
  // single-canvas lists are planned and rasterised in row bands on concurrent streams (plan of band b + 1 beside the
  // raster of band b: both kernels are latency bound and leave most of the machine idle on their own)
  // (small canvases keep one band: four more launch sets would cost more than the overlap gains)
  const int bands = (h >= 2048 && numFills >= 16) ? residentBands : 1;
  int rc = build_list(L, false, bands, w, h, layers, numFills, layerOf, seg, wind, segOff, rgbx, rule, mode);
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

int pixie_cuda_cmdlist_create_from_paths(int w, int h, int layers, int numPaths, const pixie_path_desc* paths, const float* commands,
                                         int64_t numCommandFloats, const float* rawXyxy, const int16_t* rawWinding, int64_t numRaw,
                                         pixie_cmdlist_t* out) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  if (numPaths < 0) return fail_pixie("negative path count");
  FlattenedPaths F;
  int rc = flatten_paths(numPaths, paths, commands, numCommandFloats, rawXyxy, rawWinding, numRaw, F);
  CmdList L;
  if (!rc) {
    std::vector<int32_t> layerOf((size_t)numPaths);
    std::vector<uint32_t> rgbx((size_t)numPaths);
    std::vector<uint8_t> rule((size_t)numPaths), mode((size_t)numPaths);
    for (int k = 0; k < numPaths; k++) {
      layerOf[(size_t)k] = paths[k].layer; rgbx[(size_t)k] = paths[k].rgbx;
      rule[(size_t)k] = paths[k].winding_rule; mode[(size_t)k] = paths[k].blend_mode;
    }
    rc = build_list(L, false, 1, w, h, layers, numPaths, layerOf.data(), nullptr, nullptr, F.segBegin.data(), rgbx.data(),
                    rule.data(), mode.data(), &F);
  }
  free_flattened(F);
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

int pixie_cuda_render_paths_host(uint8_t* pixels, int width, int height, int clear, int numPaths, const pixie_path_desc* paths,
                                 const float* commands, int64_t numCommandFloats, const float* rawXyxy, const int16_t* rawWinding,
                                 int64_t numRaw, uint64_t* covered_px) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  if (width <= 0 || height <= 0) return fail_pixie("Image width and height must be > 0");
  if (numPaths < 0) return fail_pixie("negative path count");
  for (int k = 0; k < numPaths; k++)
    if (paths[k].layer != 0) return fail_pixie("render_paths_host renders one canvas (layer 0)");
  Runtime& r = rt();
  void* canvas;
  const size_t bytes = (size_t)width * height * 4;
  if (int rc = get_scratch(5, bytes, &canvas)) return rc;
  // newImage(width, height): with paths to render, the raster kernel zeroes each row tile itself (clearFirst)
  if (clear && numPaths == 0) PX_CUDA(cudaMemsetAsync(canvas, 0, bytes, r.stream));
  if (!clear) PX_CUDA(cudaMemcpyAsync(canvas, pixels, bytes, cudaMemcpyHostToDevice, r.stream));  // draw over existing pixels
  Image im;
  im.data = (uint8_t*)canvas; im.w = width; im.h = height; im.layers = 1; im.bpp = 4; im.owned = false;
  FlattenedPaths F;
  int rc = flatten_paths(numPaths, paths, commands, numCommandFloats, rawXyxy, rawWinding, numRaw, F);
  CmdList L;
  if (!rc) {
    std::vector<uint32_t> rgbx((size_t)numPaths);
    std::vector<uint8_t> rule((size_t)numPaths), mode((size_t)numPaths);
    for (int k = 0; k < numPaths; k++) {
      rgbx[(size_t)k] = paths[k].rgbx; rule[(size_t)k] = paths[k].winding_rule; mode[(size_t)k] = paths[k].blend_mode;
    }
    // row bands on concurrent streams with each band's D2H behind its raster kernel, as pixie_cuda_render_batch_host
    rc = build_list(L, false, Runtime::kBands, width, height, 1, numPaths, nullptr, nullptr, nullptr, F.segBegin.data(), rgbx.data(),
                    rule.data(), mode.data(), &F);
  }
  free_flattened(F);
  if (!rc) {
    L.clearFirst = clear != 0;
    if (numPaths == 0) PX_CUDA(cudaMemcpyAsync(pixels, canvas, bytes, cudaMemcpyDeviceToHost, r.stream));
    else rc = run_list(L, &im, covered_px, pixels);
  }
  cudaStreamSynchronize(r.stream);
  free_list(L);
  return rc;
}

int pixie_cuda_cmdlist_segments(pixie_cmdlist_t list, float* segXyxy, int16_t* winding, int32_t* segOffsets) {
  PX_API_GUARD;
  auto it = g_lists.find(list);
  if (it == g_lists.end()) return fail_pixie("invalid command list handle");
  const CmdList& L = it->second;
  Runtime& r = rt();
  if (segXyxy && L.numSegs) PX_CUDA(cudaMemcpyAsync(segXyxy, L.segs, (size_t)L.numSegs * 16, cudaMemcpyDeviceToHost, r.stream));
  if (winding && L.numSegs) PX_CUDA(cudaMemcpyAsync(winding, L.wind, (size_t)L.numSegs * 2, cudaMemcpyDeviceToHost, r.stream));
  std::vector<FillHeader> fills((size_t)L.numFills);
  if (segOffsets && L.numFills)
    PX_CUDA(cudaMemcpyAsync(fills.data(), L.fills, fills.size() * sizeof(FillHeader), cudaMemcpyDeviceToHost, r.stream));
  PX_CUDA(cudaStreamSynchronize(r.stream));
  if (segOffsets) {
    for (int k = 0; k < L.numFills; k++) segOffsets[k] = fills[(size_t)k].segBegin;
    segOffsets[L.numFills] = (int32_t)L.numSegs;
  }
  return 0;
}

int pixie_cuda_cmdlist_run(pixie_cmdlist_t list, pixie_image_t image, uint64_t* covered_px) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  auto it = g_lists.find(list);
  if (it == g_lists.end()) return fail_pixie("invalid command list handle");
  Image* im = find_image(image);
  if (!im) return 1;
  return run_list(it->second, im, covered_px);
}

// newImage() + run in one: the canvas is cleared to transparent by the raster kernel itself — every (row, tile) ticket
// zeroes its pixels before it applies the first fill — instead of by a separate pass over the canvas.
int pixie_cuda_cmdlist_run_cleared(pixie_cmdlist_t list, pixie_image_t image, uint64_t* covered_px) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  auto it = g_lists.find(list);
  if (it == g_lists.end()) return fail_pixie("invalid command list handle");
  Image* im = find_image(image);
  if (!im) return 1;
  CmdList& L = it->second;
  if (L.numFills == 0) {
    if (covered_px) *covered_px = 0;
    return pixie_cuda_image_fill(image, 0u);
  }
  L.clearFirst = true;
  const int rc = run_list(L, im, covered_px);
  L.clearFirst = false;
  return rc;
}

int pixie_cuda_cmdlist_run_rows(pixie_cmdlist_t list, pixie_image_t image, int y0, int y1, uint64_t* covered_px) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  auto it = g_lists.find(list);
  if (it == g_lists.end()) return fail_pixie("invalid command list handle");
  CmdList& L = it->second;
  Image* im = find_image(image);
  if (!im) return 1;
  if (L.layers != 1) return fail_pixie("cmdlist_run_rows needs a single-layer command list");
  if (y0 < 0 || y1 > L.h || y0 > y1) return fail_pixie("row range out of bounds");
  if (im->bpp != 4 || im->layers != 1 || im->w != L.w) return fail_pixie("command list was built for a different canvas shape");
  const bool band = im->h != L.h;
  if (band && im->h != y1 - y0) return fail_pixie("cmdlist_run_rows: the image must be the whole canvas or exactly rows [y0, y1)");
  return run_list(L, im, covered_px, nullptr, y0, y1, band);
}

int pixie_cuda_cmdlist_set_overlap(pixie_cmdlist_t list, int enabled) {
  PX_API_GUARD;
  auto it = g_lists.find(list);
  if (it == g_lists.end()) return fail_pixie("invalid command list handle");
  it->second.serial = enabled == 0;
  return 0;
}

int pixie_cuda_cmdlist_info(pixie_cmdlist_t list, int64_t* numSegs, int64_t* numParts, int64_t* numEntries,
                            int64_t* launches) {
  PX_API_GUARD;
  auto it = g_lists.find(list);
  if (it == g_lists.end()) return fail_pixie("invalid command list handle");
  if (numSegs) *numSegs = it->second.numSegs;
  if (numParts) *numParts = it->second.numParts;
  if (numEntries) *numEntries = it->second.numEntries;
  // count + two scans + partition, then per band: classify, two plan kernels, raster
  if (launches) *launches = it->second.numParts > 0 ? (it->second.deviceCounted ? 4 : 1) + 4 * std::max(1, it->second.bands) : 1;
  return 0;
}

int pixie_cuda_cmdlist_destroy(pixie_cmdlist_t list) {
  PX_API_GUARD;
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
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  Image* im = find_image(image);
  if (!im) return 1;
  if (im->bpp != 4) return fail_pixie("fill needs an RGBX image");
  CmdList L;  // lives in the library arena: no allocation, no synchronisation, nothing to free
  int rc = build_list(L, true, 1, im->w, im->h, im->layers, numFills, layerOf, seg, wind, segOff, rgbx, rule, mode);
  if (!rc) rc = run_list(L, im, covered_px);
  return rc;
}

int pixie_cuda_render_batch_host(uint8_t* pixels, int width, int height, int clear, int numFills, const float* seg,
                                 const int16_t* wind, const int32_t* segOff, const uint32_t* rgbx, const uint8_t* rule,
                                 const uint8_t* mode, uint64_t* covered_px) {
  PX_API_GUARD;
  if (int rc = ensure_init()) return rc;
  if (width <= 0 || height <= 0) return fail_pixie("Image width and height must be > 0");
  Runtime& r = rt();
  void* canvas;
  const size_t bytes = (size_t)width * height * 4;
  if (int rc = get_scratch(5, bytes, &canvas)) return rc;
  // newImage(width, height): with fills to render, the raster kernel zeroes each row tile itself (clearFirst)
  if (clear && numFills == 0) PX_CUDA(cudaMemsetAsync(canvas, 0, bytes, r.stream));
  if (!clear) PX_CUDA(cudaMemcpyAsync(canvas, pixels, bytes, cudaMemcpyHostToDevice, r.stream));  // draw over existing pixels
  Image im;
  im.data = (uint8_t*)canvas; im.w = width; im.h = height; im.layers = 1; im.bpp = 4; im.owned = false;
  CmdList L;
  int rc = build_list(L, true, Runtime::kBands, width, height, 1, numFills, nullptr, seg, wind, segOff, rgbx, rule, mode);
  if (rc) return rc;
  if (numFills == 0) {
    PX_CUDA(cudaMemcpyAsync(pixels, canvas, bytes, cudaMemcpyDeviceToHost, r.stream));
  } else {
    L.clearFirst = clear != 0;
    rc = run_list(L, &im, covered_px, pixels);
    if (rc) return rc;
  }
  PX_CUDA(cudaStreamSynchronize(r.stream));
  return 0;
}

int pixie_cuda_fill_segments(pixie_image_t image, const float* seg, const int16_t* wind, int n, uint32_t rgbx,
                             int rule, int mode) {
  PX_API_GUARD;
  if (rule < 0 || rule > 1 || mode < 0 || mode >= NumBlendModes) return fail_pixie("invalid blend mode / winding rule");
  const int32_t segOff[2] = {0, n};
  const uint8_t r8 = (uint8_t)rule, m8 = (uint8_t)mode;
  return pixie_cuda_fill_batch(image, 1, nullptr, seg, wind, segOff, &rgbx, &r8, &m8, nullptr);
}

}  // extern "C"

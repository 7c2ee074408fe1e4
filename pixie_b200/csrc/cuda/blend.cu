// K4 — image-over-image blend: draw() integer-translate fast path = blendRect
// (treeform/pixie src/pixie/images.nim:468-529) for all 20 BlendMode enumerators
// (blends.nim:275-299), plus the fused mask*fill composite of non-solid paints
// (paths.nim:2141-2142) and applyOpacity (images.nim:261-277).
//
// Pure streaming, HBM-bound for the integer modes: 16-byte vector loads/stores on dst (and on
// src/mask when the translate keeps 16-byte alignment), 4 pixels per thread per row, 4 rows in
// flight per thread, grid sized to a multiple of the SM count.
#include "common.cuh"

namespace pixie {

struct RectArgs {
  px_t* dst;
  const px_t* src;
  const uint8_t* mask;  // RGBX (alpha used) or A8, same size as src
  int dw, dh, sw, sh, px, py;
  int xs, xe, ys, ye;      // dst-space region this launch processes
  int rx0, rx1, ry0, ry1;  // dst-space rect actually covered by src (clipped)
  int src_aligned;         // px % 4 == 0 && sw % 4 == 0 -> 16-byte src (and mask) loads
};

template <int MODE>
PXD px_t rect_op(px_t d, px_t s, const BlendTab& T) {
  // Normal / Mask rows go through the x86 row kernels (images.nim:485-520 -> sse2.nim:590-616,690-715),
  // Overwrite is a copy, everything else is blender() per pixel (images.nim:521-529).
  if (MODE == NormalBlend) return line_normal(d, s);
  if (MODE == MaskBlend) return line_mask(d, s);
  if (MODE == OverwriteBlend) return s;
  if (mode_uses_tables(MODE)) return blend_px_tab<MODE>(d, s, T);
  return blend_px<MODE>(d, s);
}

// straight[a << 8 | c] = straight_(c, a): built once per process by the function the table replaces
__device__ uint8_t g_straight_table[65536];
__global__ void __launch_bounds__(256) build_straight_table_kernel() {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  g_straight_table[i] = (uint8_t)straight_(i & 255u, i >> 8);
}
constexpr size_t kTabSmem = 65536 + 1024 + 1024;  // straight, inv, div255
static bool g_straight_built = false;             // per process (one GPU per process); the API lock serialises callers

PXD uint4 ld16(const void* p) { return *reinterpret_cast<const uint4*>(p); }
PXD uint4 ld16_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// One thread = one 16-byte group of dst (4 px) x several rows.
#ifndef PIXIE_BLEND_ROWS
#define PIXIE_BLEND_ROWS 2   // rows in flight per thread (swept on B200: 2 rows x 4 CTAs/SM is best)
#endif
#ifndef PIXIE_BLEND_MINB
#define PIXIE_BLEND_MINB 4
#endif
#ifndef PIXIE_BLEND_MINB_FLOAT
#define PIXIE_BLEND_MINB_FLOAT 4
#endif
#ifndef PIXIE_BLEND_MINB_PACKED
#define PIXIE_BLEND_MINB_PACKED 3   // the two-pixel packed path wants registers (swept: 2 -> 0.472 ms, 3 -> 0.437 ms for SoftLight at 8192^2)
#endif
template <int MODE, int MASK>
__global__ void __launch_bounds__(256, mode_uses_tables(MODE) ? 3 : (mode_is_packed(MODE) ? PIXIE_BLEND_MINB_PACKED : (mode_is_float(MODE) ? PIXIE_BLEND_MINB_FLOAT : PIXIE_BLEND_MINB))) blend_rect_vec4(const RectArgs a) {
  BlendTab T = {nullptr, nullptr, nullptr};
  if (mode_uses_tables(MODE)) {  // the ALU-bound modes: tables into shared memory (66 KB, three CTAs per SM, one wave)
    extern __shared__ __align__(16) uint8_t tab_smem[];
    const uint4* src = reinterpret_cast<const uint4*>(g_straight_table);
    uint4* dst = reinterpret_cast<uint4*>(tab_smem);
#pragma unroll 4
    for (int i = threadIdx.x; i < 65536 / 16; i += 256) dst[i] = src[i];
    uint32_t* inv = reinterpret_cast<uint32_t*>(tab_smem + 65536);
    float* d255 = reinterpret_cast<float*>(tab_smem + 65536 + 1024);
    inv[threadIdx.x] = g_inv_table.v[threadIdx.x];
    d255[threadIdx.x] = g_blend_tables.div255[threadIdx.x];
    __syncthreads();
    T.straight = tab_smem;
    T.inv = inv;
    T.div255 = d255;
  }
  const int g0 = a.xs >> 2;
  const int g = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int x = g << 2;
  if (x >= a.xe) return;
  const bool full_x = (x >= a.xs) && (x + 4 <= a.xe) && (x >= a.rx0) && (x + 4 <= a.rx1);
  constexpr int ROWS = PIXIE_BLEND_ROWS;
  for (int yb = a.ys + blockIdx.y * ROWS; yb < a.ye; yb += gridDim.y * ROWS) {
    // ---- fast path: this 4x4 block of pixels lies entirely inside the drawn rect: no predicates
    if (full_x && a.src_aligned && yb >= a.ry0 && yb + ROWS <= a.ry1 && yb + ROWS <= a.ye) {
      uint4 dv[ROWS], sv[ROWS];
      uint32_t mw[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; r++) {
        const int y = yb + r;
        const size_t sidx = (size_t)a.sw * (y - a.py) + (x - a.px);
        if (MODE != OverwriteBlend) dv[r] = ld16(a.dst + (size_t)a.dw * y + x);
        sv[r] = ld16_stream(a.src + sidx);
        if (MASK == 1) {
          const uint4 m = ld16_stream(reinterpret_cast<const px_t*>(a.mask) + sidx);
          mw[r] = (m.x >> 24) | ((m.y >> 24) << 8) | ((m.z >> 24) << 16) | (m.w & 0xFF000000u);
        } else if (MASK == 2) {
          mw[r] = __ldg(reinterpret_cast<const uint32_t*>(a.mask + sidx));
        } else {
          mw[r] = 0xFFFFFFFFu;
        }
      }
#pragma unroll
      for (int r = 0; r < ROWS; r++) {
        uint32_t* dp = reinterpret_cast<uint32_t*>(&dv[r]);
        const uint32_t* sp = reinterpret_cast<const uint32_t*>(&sv[r]);
        if (mode_is_packed(MODE)) {  // two pixels per step on the packed fp32 instructions
#pragma unroll
          for (int k = 0; k < 4; k += 2) {
            px_t s0 = sp[k], s1 = sp[k + 1];
            if (MASK != 0) {
              const uint32_t m0 = (mw[r] >> (8 * k)) & 255u, m1 = (mw[r] >> (8 * k + 8)) & 255u;
              if (m0 != 255u) s0 = mul_div255(s0, m0);
              if (m1 != 255u) s1 = mul_div255(s1, m1);
            }
            blend_px2_float<MODE>(dp[k], dp[k + 1], s0, s1, dp[k], dp[k + 1]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            px_t sx = sp[k];
            if (MASK != 0) {
              const uint32_t m = (mw[r] >> (8 * k)) & 255u;
              if (m != 255u) sx = mul_div255(sx, m);
            }
            dp[k] = rect_op<MODE>(MODE == OverwriteBlend ? 0u : dp[k], sx, T);
          }
        }
        *reinterpret_cast<uint4*>(a.dst + (size_t)a.dw * (yb + r) + x) = dv[r];
      }
      continue;
    }
    uint4 dv[ROWS], sv[ROWS];
    uint32_t mv[ROWS][4];
    bool rowin[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      const int y = yb + r;
      rowin[r] = y < a.ye;
      if (!rowin[r]) continue;
      const bool in_y = (y >= a.ry0) && (y < a.ry1);
      px_t* drow = a.dst + (size_t)a.dw * y + x;
      const bool need_dst = !(MODE == OverwriteBlend && full_x && in_y) && !(MODE == MaskBlend && !in_y);
      dv[r] = need_dst ? ld16(drow) : make_uint4(0, 0, 0, 0);
      sv[r] = make_uint4(0, 0, 0, 0);
      mv[r][0] = mv[r][1] = mv[r][2] = mv[r][3] = 255u;
      if (in_y) {
        const size_t sidx = (size_t)a.sw * (y - a.py) + (x - a.px);
        uint32_t* sp = reinterpret_cast<uint32_t*>(&sv[r]);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int xx = x + k;
          if (xx >= a.rx0 && xx < a.rx1 && xx >= a.xs && xx < a.xe) {
            sp[k] = __ldg(a.src + sidx + k);
            if (MASK == 1) mv[r][k] = __ldg(reinterpret_cast<const px_t*>(a.mask) + sidx + k) >> 24;
            else if (MASK == 2) mv[r][k] = a.mask[sidx + k];
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      if (!rowin[r]) continue;
      const int y = yb + r;
      const bool in_y = (y >= a.ry0) && (y < a.ry1);
      uint32_t* dp = reinterpret_cast<uint32_t*>(&dv[r]);
      const uint32_t* sp = reinterpret_cast<const uint32_t*>(&sv[r]);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int xx = x + k;
        const bool in_region = xx >= a.xs && xx < a.xe;
        const bool in_rect = in_y && xx >= a.rx0 && xx < a.rx1;
        if (!in_region) continue;
        if (in_rect) {
          px_t sx = sp[k];
          if (MASK != 0) sx = mul_div255(sx, mv[r][k]);
          dp[k] = rect_op<MODE>(dp[k], sx, T);
        } else if (MODE == MaskBlend) {
          dp[k] = 0u;  // images.nim:501-520: MaskBlend clears everything the source does not cover
        }
      }
      *reinterpret_cast<uint4*>(a.dst + (size_t)a.dw * y + x) = dv[r];
    }
  }
}

// Fallback for canvases whose rows are not 16-byte aligned (width % 4 != 0): one pixel per thread.
template <int MODE, int MASK>
__global__ void __launch_bounds__(256) blend_rect_scalar(const RectArgs a) {
  const int x = a.xs + blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= a.xe) return;
  for (int y = a.ys + blockIdx.y; y < a.ye; y += gridDim.y) {
    px_t* dp = a.dst + (size_t)a.dw * y + x;
    const bool in_rect = y >= a.ry0 && y < a.ry1 && x >= a.rx0 && x < a.rx1;
    if (in_rect) {
      const size_t sidx = (size_t)a.sw * (y - a.py) + (x - a.px);
      px_t s = a.src[sidx];
      if (MASK == 1) s = mul_div255(s, reinterpret_cast<const px_t*>(a.mask)[sidx] >> 24);
      else if (MASK == 2) s = mul_div255(s, a.mask[sidx]);
      const BlendTab T0 = {nullptr, nullptr, nullptr};
      *dp = mode_uses_tables(MODE) ? blend_px<MODE>(*dp, s) : rect_op<MODE>(*dp, s, T0);
    } else if (MODE == MaskBlend) {
      *dp = 0u;
    }
  }
}

template <int MODE, int MASK>
static int launch_rect(const RectArgs& a) {
  Runtime& r = rt();
  const int rows = a.ye - a.ys;
  if (rows <= 0 || a.xe <= a.xs) return 0;
  constexpr bool TAB = mode_uses_tables(MODE);
  // table modes: one wave of three CTAs per SM, each staging the tables once and striding over the rows
  const int target_blocks = TAB ? r.num_sms * 3 : r.num_sms * 8;
  if ((a.dw & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dst) & 15) == 0) {
    const int groups = ((a.xe + 3) >> 2) - (a.xs >> 2);
    dim3 grid((groups + 255) / 256, 1);
    int gy = target_blocks / (int)grid.x;
    if (gy < 1) gy = 1;
    int max_gy = (rows + PIXIE_BLEND_ROWS - 1) / PIXIE_BLEND_ROWS;
    if (gy > max_gy) gy = max_gy;
    grid.y = gy;
    if (TAB) {
      if (!g_straight_built) {
        build_straight_table_kernel<<<256, 256, 0, r.stream>>>();
        PX_LAUNCHED();
        g_straight_built = true;
      }
      static bool configured = false;  // per instantiation
      if (!configured) {
        PX_CUDA(cudaFuncSetAttribute(blend_rect_vec4<MODE, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTabSmem));
        configured = true;
      }
    }
    ProfScope ps(kProfBlend);
    blend_rect_vec4<MODE, MASK><<<grid, 256, TAB ? kTabSmem : 0, r.stream>>>(a);
  } else {
    dim3 grid((a.xe - a.xs + 255) / 256, 1);
    int gy = target_blocks / (int)grid.x;
    if (gy < 1) gy = 1;
    if (gy > rows) gy = rows;
    grid.y = gy;
    blend_rect_scalar<MODE, MASK><<<grid, 256, 0, r.stream>>>(a);
  }
  PX_LAUNCHED();
  return 0;
}

template <int MASK>
static int dispatch_rect(int mode, const RectArgs& a) {
  int rc = 0;
  PX_DISPATCH_MODE(mode, rc = (launch_rect<MODE, MASK>(a)));
  return rc;
}

// dst <- blend(dst, src at (px, py)); src / mask given as raw device pointers (draw.cu passes minified copies)
static int blend_rect_core(Image* d, const px_t* src, int sw, int sh, const uint8_t* mask, int mask_bpp, int px, int py,
                           int mode) {
  RectArgs a;
  a.dst = (px_t*)d->data;
  a.src = src;
  a.mask = mask;
  a.dw = d->w; a.dh = d->h; a.sw = sw; a.sh = sh; a.px = px; a.py = py;
  a.src_aligned = ((px & 3) == 0 && (sw & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                   (!mask || (reinterpret_cast<uintptr_t>(mask) & 15) == 0)) ? 1 : 0;
  // images.nim:473-476
  const bool outside = (int64_t)px >= d->w || (int64_t)px + sw <= 0 || (int64_t)py >= d->h || (int64_t)py + sh <= 0;
  if (outside) {
    if (mode == MaskBlend) PX_CUDA(cudaMemsetAsync(d->data, 0, d->layer_bytes(), rt().stream));
    return 0;
  }
  // images.nim:478-482 (in dst space)
  a.rx0 = px > 0 ? px : 0;
  a.ry0 = py > 0 ? py : 0;
  a.rx1 = (px + sw < d->w) ? px + sw : d->w;
  a.ry1 = (py + sh < d->h) ? py + sh : d->h;
  if (mode == MaskBlend) {
    a.xs = 0; a.xe = d->w; a.ys = 0; a.ye = d->h;
  } else {
    a.xs = a.rx0; a.xe = a.rx1; a.ys = a.ry0; a.ye = a.ry1;
  }
  const int mk_ = !mask ? 0 : (mask_bpp == 4 ? 1 : 2);
  if (mk_ == 0) return dispatch_rect<0>(mode, a);
  if (mk_ == 1) return dispatch_rect<1>(mode, a);
  return dispatch_rect<2>(mode, a);
}

int blend_rect_raw(Image* d, const px_t* src, int sw, int sh, int px, int py, int mode) {
  return blend_rect_core(d, src, sw, sh, nullptr, 0, px, py, mode);
}

static int blend_rect_impl(pixie_image_t dsth, pixie_image_t srch, pixie_image_t maskh, bool masked, int px, int py,
                           int mode) {
  if (int rc = ensure_init()) return rc;
  if (mode < 0 || mode >= NumBlendModes) return fail_pixie("invalid blend mode");
  Image* d = find_image(dsth);
  Image* s = find_image(srch);
  if (!d || !s) return 1;
  if (d->bpp != 4 || s->bpp != 4) return fail_pixie("blend_rect needs RGBX images");
  Image* m = nullptr;
  if (masked) {
    m = find_image(maskh);
    if (!m) return 1;
    if (m->w != s->w || m->h != s->h) return fail_pixie("mask must have the size of src");
  }
  if (d->data == s->data) return fail_pixie("blend_rect: dst and src must be different images");
  return blend_rect_core(d, (const px_t*)s->data, s->w, s->h, m ? m->data : nullptr, m ? m->bpp : 0, px, py, mode);
}

__global__ void __launch_bounds__(256) apply_opacity_kernel(uint4* __restrict__ p, size_t n16, uint32_t o) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) {
    uint4 v = p[i];
    v.x = mul_div255(v.x, o); v.y = mul_div255(v.y, o); v.z = mul_div255(v.z, o); v.w = mul_div255(v.w, o);
    p[i] = v;
  }
}
__global__ void apply_opacity_tail(px_t* p, size_t n, uint32_t o) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = mul_div255(p[i], o);
}

}  // namespace pixie

using namespace pixie;

extern "C" {

int pixie_cuda_blend_rect(pixie_image_t dst, pixie_image_t src, int px, int py, int mode) {
  PX_API_GUARD;
  return blend_rect_impl(dst, src, 0, false, px, py, mode);
}
int pixie_cuda_blend_rect_masked(pixie_image_t dst, pixie_image_t src, pixie_image_t mask, int px, int py, int mode) {
  PX_API_GUARD;
  return blend_rect_impl(dst, src, mask, true, px, py, mode);
}

int pixie_cuda_apply_opacity(pixie_image_t h, float opacity) { PX_API_GUARD;  // images.nim:261-277
  if (int rc = ensure_init()) return rc;
  Image* im = find_image(h);
  if (!im) return 1;
  if (im->bpp != 4) return fail_pixie("apply_opacity needs an RGBX image");
  const uint32_t o = (uint32_t)(uint16_t)(int64_t)roundf(255 * opacity);
  if (o == 255) return 0;
  if (o == 0) return pixie_cuda_image_fill(h, 0u);
  if (o > 255) return fail_pixie("opacity out of range");
  Runtime& r = rt();
  const size_t npx = im->bytes() / 4, n16 = npx / 4;
  if (n16) {
    int blocks = (int)std::min<size_t>((n16 + 255) / 256, (size_t)r.num_sms * 16);
    apply_opacity_kernel<<<blocks, 256, 0, r.stream>>>((uint4*)im->data, n16, o);
    PX_LAUNCHED();
  }
  if (npx - n16 * 4) {
    apply_opacity_tail<<<1, 32, 0, r.stream>>>((px_t*)im->data + n16 * 4, npx - n16 * 4, o);
    PX_LAUNCHED();
  }
  return 0;
}

}  // extern "C"

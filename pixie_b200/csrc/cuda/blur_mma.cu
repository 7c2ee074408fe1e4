// K5b — the separable Gaussian blur (images.nim:304-365) as an exact banded contraction on the tensor cores.
//
// A blur pass is  out[a] = (sum_t lut[t] * in[a - r + t]) div 256 div 255  per channel and line: a Toeplitz
// (banded) matrix times the pixel columns.  With 65 taps x 4 channels x 2 passes per pixel the CUDA-core
// kernels of blur.cu are ALU-bound at ~4 % of the HBM roofline; the contraction itself is
// exact in the tensor cores' number formats:
//   * pixel bytes 0..255 are exact in fp16 — and need no conversion: a byte b next to a zero byte IS the fp16
//     subnormal b * 2^-24, so the planes are built with byte permutes only and the whole contraction runs scaled by
//     2^-24 (exact: every partial sum is a multiple of 2^-24 below 1);
//   * a uint16 tap k splits into  k = lo + hi * 2048  with lo < 2048 (11 significant bits) and hi * 2048 <= 63488,
//     both exact in fp16 (Gaussian LUTs of radius >= 29 have hi == 0 everywhere: one MMA per tile);
//   * every product is an integer (times 2^-24) below 2^24 and so is every partial sum (sum lut * 255 < 2^24 is
//     checked by the caller), so the fp32 accumulation never rounds.
// The results are therefore bit-identical to the reference's uint32 arithmetic; `div 256 div 255` is one FFMA.RZ on
// the accumulator: floor(a * float(1 / 65280)) == a div 65280 for every integer a < 2^24 (exhaustive check in
// tests/test_chain_closed_form.py), the FMA forms the product exactly, and adding 2^23 puts the truncation at ulp 1.
//
// Shape: mma.sync.m16n8k16 (fp16 x fp16 -> fp32).  M = 16 outputs along the blur axis, K = 16 inputs, N = 8
// lines.  The A operand is the Toeplitz block  A_q[m][k] = lut[16 q + k - m]  (q = 0 .. KT-1 with
// KT = ceil((2r + 16) / 16)); it is the same for every tile, lives in registers, and only KT of the
// (outputs/16 + KT - 1) k-tiles of a row of tiles are non-zero — 65/80 of the multiply-adds are useful at r = 32.
// B is the pixel data, staged global -> shared as four planar fp16 channel planes (cp.async of the raw tile, then
// 1.5 PRMT per two bytes) and read back with ldmatrix (.trans for the vertical pass), each k-tile once per warp
// for all the m-tiles it feeds.  One CTA = 128 outputs x 32 lines x 4 channels; a warp owns 4 m-tiles x 8 lines.
#include <cuda_fp16.h>

#include "common.cuh"

namespace pixie {

struct MmaBlurArgs {
  const px_t* src;
  px_t* dst;
  int w, h;
  int radius;
  uint32_t oob;
  int y0, y1;    // vertical pass: output rows
  int sy0, sy1;  // horizontal pass: rows to produce
  int pitch;     // shared-memory row pitch in halfs
  int hasHi;     // some tap >= 2048
};

// float(1 / 65280): floor(a * kInv65280) == a div 256 div 255 for all integers 0 <= a < 2^24.  kInv65280s is the
// same times 2^24, for accumulators that hold a * 2^-24 (pixel bytes entering the contraction as fp16 subnormals).
#define kInv65280 __uint_as_float(0x37808081u)
#define kInv65280s __uint_as_float(0x43808081u)
constexpr int kMmaOut = 128;   // outputs per CTA along the blur axis
constexpr int kMmaLines = 32;  // lines per CTA
constexpr int kMaxTaps = 2 * 64 + 1;
__constant__ uint16_t c_mma_lut[kMaxTaps + 3];

PXD void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool TRANS>
PXD void ldmatrix_x2(uint32_t& r0, uint32_t& r1, const __half* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  if (TRANS) asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
  else asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}
PXD uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
PXD int tap_at(int t, int ntaps) { return (t >= 0 && t < ntaps) ? (int)c_mma_lut[t] : 0; }

PXD void cp_async_4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
PXD void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
PXD void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// VERTICAL = false: blur along x: a tile is 32 rows (lines) x IN_A pixels, planes are [line][a];
// VERTICAL = true:  blur along y: a tile is IN_A rows x 32 pixels (lines), planes are [a][line], read with
//                   ldmatrix.trans.  Either way the raw tile and the planes are [tile row][tile column] with the
// image's x running along the columns, so staging and conversion are the same code.
//
// Persistent CTAs: the Toeplitz fragments are built once; per tile the raw RGBX bytes of the NEXT tile are
// fetched with cp.async while the tensor cores work on the current one.
template <bool VERTICAL, int KT, bool HI>
__global__ void __launch_bounds__(256, 2) blur_mma_kernel(const MmaBlurArgs a, int tilesA, int numTiles) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int IN_A = kMmaOut - 16 + 16 * KT;  // inputs along the blur axis
  constexpr int ROWS = VERTICAL ? IN_A : kMmaLines, COLS = VERTICAL ? kMmaLines : IN_A;  // tile shape in pixels
  const int pitch = a.pitch;
  const int planeHalfs = ROWS * pitch;
  __half* planes = reinterpret_cast<__half*>(smem_raw);
  px_t* raw = reinterpret_cast<px_t*>(smem_raw + (size_t)4 * planeHalfs * sizeof(__half));  // ROWS x COLS pixels
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntaps = 2 * a.radius + 1;
  const int a_len = VERTICAL ? a.h : a.w;
  const int l_end = VERTICAL ? a.w : a.sy1;
  const bool vec_ok = (a.w & 3) == 0 && (reinterpret_cast<uintptr_t>(a.src) & 15) == 0 && (VERTICAL || (a.radius & 3) == 0);

  // ---- Toeplitz fragments (row-major m16 x k16): A_q[m][k] = lut[16 q + k - m], split lo + hi * 2048
  const int g = lane >> 2, t = lane & 3;
  uint32_t Alo[KT][4], Ahi[HI ? KT : 1][4];
#pragma unroll
  for (int q = 0; q < KT; q++) {
#pragma unroll
    for (int rIdx = 0; rIdx < 4; rIdx++) {
      const int m = g + ((rIdx & 1) ? 8 : 0), k = 2 * t + ((rIdx & 2) ? 8 : 0);
      const int k0 = tap_at(16 * q + k - m, ntaps), k1 = tap_at(16 * q + k + 1 - m, ntaps);
      Alo[q][rIdx] = pack_h2((float)(k0 & 2047), (float)(k1 & 2047));
      if (HI) Ahi[q][rIdx] = pack_h2((float)((k0 >> 11) << 11), (float)((k1 >> 11) << 11));
    }
  }

  auto tile_origin = [&](int tile, int& a0, int& l0) {
    const int ta = tile % tilesA, tl = tile / tilesA;
    a0 = ta * kMmaOut + (VERTICAL ? a.y0 : 0);     // first output along the blur axis
    l0 = tl * kMmaLines + (VERTICAL ? 0 : a.sy0);  // first line
  };
  // raw tile <- global (cp.async), out-of-image pixels <- the out-of-bounds colour
  auto prefetch = [&](int a0, int l0) {
    const int x0 = VERTICAL ? l0 : a0 - a.radius, y0 = VERTICAL ? a0 - a.radius : l0;
    const int xEnd = VERTICAL ? l_end : a_len, yEnd = VERTICAL ? a_len : l_end;
    constexpr int G = COLS / 4;  // groups of 4 pixels per tile row
    if (vec_ok && x0 >= 0 && x0 + COLS <= xEnd && y0 >= 0 && y0 + ROWS <= yEnd) {
      // tile inside the image: no per-piece checks, and the piece -> (row, column) map is the same for every tile
      const px_t* base = a.src + (size_t)a.w * y0 + x0;
#pragma unroll
      for (int u = 0; u < (ROWS * G + 255) / 256; u++) {
        const int idx = tid + 256 * u;
        const int row = idx / G, c4 = (idx - row * G) * 4;
        if (idx < ROWS * G) cp_async_16(raw + 4 * idx, base + (a.w * row + c4));
      }
      return;
    }
    for (int idx = tid; idx < ROWS * G; idx += 256) {
      const int row = idx / G, c4 = (idx - row * G) * 4;
      const int y = y0 + row, x = x0 + c4;
      px_t* dst = raw + row * COLS + c4;
      const bool rowIn = y >= 0 && y < yEnd;
      if (rowIn && vec_ok && x >= 0 && x + 4 <= xEnd) {
        cp_async_16(dst, a.src + (size_t)a.w * y + x);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (rowIn && x + j >= 0 && x + j < xEnd) cp_async_4(dst + j, a.src + (size_t)a.w * y + x + j);
          else dst[j] = a.oob;
        }
      }
    }
  };

  const int nt = warp & 3, mg = warp >> 2;

  int tile = blockIdx.x;
  int a0n = 0, l0n = 0;  // origin of the tile being fetched: one division per tile
  if (tile < numTiles) {
    tile_origin(tile, a0n, l0n);
    prefetch(a0n, l0n);
  }
#pragma unroll 1
  for (; tile < numTiles; tile += gridDim.x) {
    cp_async_wait_all();
    __syncthreads();  // the raw tile is complete, and nobody reads the planes of the previous tile any more
    {                 // raw RGBX bytes -> four planar fp16 planes, 4 pixels per thread and step
      constexpr int G = COLS / 4;
      for (int idx = tid; idx < ROWS * G; idx += 256) {
        const int row = idx / G, c4 = (idx - row * G) * 4;
        const uint4 p = *reinterpret_cast<const uint4*>(raw + row * COLS + c4);
        const uint32_t rg01 = __byte_perm(p.x, p.y, 0x5140), ba01 = __byte_perm(p.x, p.y, 0x7362);
        const uint32_t rg23 = __byte_perm(p.z, p.w, 0x5140), ba23 = __byte_perm(p.z, p.w, 0x7362);
        // a byte b next to a zero byte is the fp16 SUBNORMAL b * 2^-24: exact, no int -> float conversion at all; the
        // whole contraction is scaled by 2^-24 (exact: every partial sum is a multiple of 2^-24 below 1) and the
        // epilogue's multiplier takes the scale back
        uint32_t w[8];
        w[0] = __byte_perm(rg01, 0u, 0x4140); w[1] = __byte_perm(rg23, 0u, 0x4140);  // r0 r1 | r2 r3
        w[2] = __byte_perm(rg01, 0u, 0x4342); w[3] = __byte_perm(rg23, 0u, 0x4342);  // g
        w[4] = __byte_perm(ba01, 0u, 0x4140); w[5] = __byte_perm(ba23, 0u, 0x4140);  // b
        w[6] = __byte_perm(ba01, 0u, 0x4342); w[7] = __byte_perm(ba23, 0u, 0x4342);  // a
        // vertical pass: rows are 32 halfs with no padding; the 16-byte chunk of a row is XOR-swizzled with
        // (row >> 1) & 3, which keeps both these stores (two rows per half-warp) and ldmatrix.trans (8 rows of one
        // chunk) on distinct banks
        __half* d = VERTICAL ? planes + row * pitch + ((((c4 >> 3) ^ (row >> 1)) & 3) << 3) + (c4 & 7) : planes + row * pitch + c4;
#pragma unroll
        for (int c = 0; c < 4; c++) *reinterpret_cast<uint2*>(d + c * planeHalfs) = make_uint2(w[2 * c], w[2 * c + 1]);
      }
    }
    __syncthreads();  // planes ready; the raw buffer is free again
    const int a0 = a0n, l0 = l0n;
    if (tile + (int)gridDim.x < numTiles) {
      tile_origin(tile + gridDim.x, a0n, l0n);
      prefetch(a0n, l0n);
    }

    // ---- contraction: warp = 8 lines (nt) x 4 m-tiles (mg), all four channels
    // Quantised outputs are kept two per register while the channels come in: rg[i][h] / ba[i][h] hold the bytes
    // {c0 of slot 2h, c1 of slot 2h, c0 of slot 2h+1, c1 of slot 2h+1} of m-tile i.
    uint32_t rg[4][2], ba[4][2];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
      const __half* plane = planes + c * planeHalfs;
#pragma unroll
      for (int kt = 0; kt < 4 + KT - 1; kt++) {
        const int kbase = (mg * 4 + kt) * 16;  // first input of this k-tile, relative to the tile origin
        uint32_t b0, b1;
        if (VERTICAL) {  // rows of the stored matrix = inputs (k), columns = lines
          ldmatrix_x2<true>(b0, b1, plane + (kbase + (lane & 15)) * pitch + (((nt ^ (lane >> 1)) & 3) << 3));  // kbase % 16 == 0
        } else {  // rows of the stored matrix = lines, columns = inputs (k)
          ldmatrix_x2<false>(b0, b1, plane + (nt * 8 + (lane & 7)) * pitch + kbase + ((lane & 8) ? 8 : 0));
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int q = kt - i;
          if (q >= 0 && q < KT) {
            mma_16816(acc[i], Alo[q], b0, b1);
            if (HI) mma_16816(acc[i], Ahi[q], b0, b1);
          }
        }
      }
      // `div 256 div 255` (images.nim:332-338) = floor(acc / 65280) in ONE instruction: with kInv65280 = 0x37808081
      // (1 / 65280 rounded to nearest, which lies above it), floor(a * kInv65280) == a div 65280 for every integer
      // a < 2^24 (checked exhaustively; tests/test_chain_closed_form.py repeats the check), and FFMA.RZ computes
      // a * kInv65280 + 2^23 exactly before truncating at ulp 1, so the quotient sits in the low bits of the result.
      // The two result bytes are merged with the other channel of the pair by one PRMT.
#pragma unroll
      for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          const uint32_t t0 = __float_as_uint(__fmaf_rz(acc[i][2 * hh], kInv65280s, 8388608.0f));
          const uint32_t t1 = __float_as_uint(__fmaf_rz(acc[i][2 * hh + 1], kInv65280s, 8388608.0f));
          const uint32_t qq = __byte_perm(t0, t1, 0x5410);  // bytes {q0, q0 >> 8, q1, q1 >> 8}
          if (c == 0) rg[i][hh] = qq;
          else if (c == 1) rg[i][hh] = __byte_perm(rg[i][hh], qq, 0x6240);  // {r0, g0, r1, g1}
          else if (c == 2) ba[i][hh] = qq;
          else ba[i][hh] = __byte_perm(ba[i][hh], qq, 0x6240);
        }
      }
    }
    uint32_t pix[4][4];  // [m-tile][fragment slot]: packed RGBX of the lane's 4 outputs per m-tile
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        pix[i][2 * hh] = __byte_perm(rg[i][hh], ba[i][hh], 0x5410);
        pix[i][2 * hh + 1] = __byte_perm(rg[i][hh], ba[i][hh], 0x7632);
      }
    }

    // ---- store: fragment slot s of m-tile i is output (m = g + 8 (s >> 1), n = 2 t + (s & 1))
    const bool inside = VERTICAL ? (a0 + kMmaOut <= min(a.y1, a.h) && l0 + kMmaLines <= a.w && (a.w & 1) == 0 &&
                                    (reinterpret_cast<uintptr_t>(a.dst) & 7) == 0)
                                 : (a0 + kMmaOut <= a.w && l0 + kMmaLines <= l_end);
    if (inside) {  // whole tile inside the image: one base pointer, constant strides, no checks
      if (VERTICAL) {
        px_t* p = a.dst + (size_t)a.w * (a0 + mg * 64 + g) + (l0 + nt * 8 + 2 * t);
        const size_t w8 = (size_t)a.w * 8;
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
          for (int hrow = 0; hrow < 2; hrow++, p += w8) *reinterpret_cast<uint2*>(p) = make_uint2(pix[i][2 * hrow], pix[i][2 * hrow + 1]);
        }
      } else {
        px_t* p0 = a.dst + (size_t)a.w * (l0 + nt * 8 + 2 * t) + (a0 + mg * 64 + g);
        px_t* p1 = p0 + a.w;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          p0[16 * i] = pix[i][0];
          p1[16 * i] = pix[i][1];
          p0[16 * i + 8] = pix[i][2];
          p1[16 * i + 8] = pix[i][3];
        }
      }
      continue;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int abase = a0 + (mg * 4 + i) * 16;
      if (VERTICAL) {
        const int x = l0 + nt * 8 + 2 * t;
#pragma unroll
        for (int hrow = 0; hrow < 2; hrow++) {
          const int y = abase + g + 8 * hrow;
          if (y < a.y1 && y < a.h) {
            px_t* p = a.dst + (size_t)a.w * y + x;
            if (x + 1 < a.w && ((reinterpret_cast<uintptr_t>(p) & 7) == 0)) {
              *reinterpret_cast<uint2*>(p) = make_uint2(pix[i][2 * hrow], pix[i][2 * hrow + 1]);
            } else {
              if (x < a.w) p[0] = pix[i][2 * hrow];
              if (x + 1 < a.w) p[1] = pix[i][2 * hrow + 1];
            }
          }
        }
      } else {
#pragma unroll
        for (int s_ = 0; s_ < 4; s_++) {
          const int x = abase + g + 8 * (s_ >> 1), y = l0 + nt * 8 + 2 * t + (s_ & 1);
          if (x < a.w && y < l_end) a.dst[(size_t)a.w * y + x] = pix[i][s_];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The same contraction on an 8-bit plane (one channel): shadow() only ever uses the alpha of its mask
// (images.nim:760-776: spread leaves rgbx(0, 0, 0, a), blur keeps the channels apart, the final MaskBlend draw reads
// mask.a), so its blur runs on the alpha plane — a quarter of the staging, MMA and epilogue work.
// src / dst are uint8 planes of w x h.  With one byte per pixel the per-tile bookkeeping is what counts, so this
// kernel differs from the RGBX one around the contraction:
//   * the tile is fetched global -> registers in 16-byte pieces (two per thread, issued before the MMAs of the
//     previous tile and converted after them), no raw copy in shared memory;
//   * the horizontal tile starts at a0 - roundup16(radius) so that every piece is 16-byte aligned for any radius;
//     the `shift` = roundup16(radius) - radius extra inputs on the left are folded into the Toeplitz fragments
//     (A_q[m][k] = lut[16 q + k - m - shift]);
//   * results leave through a shared-memory tile in image orientation and go out as 16-byte stores, one tile late
//     (while the next tile is being converted), which keeps two barriers per tile.
// ---------------------------------------------------------------------------------------------
struct MmaBlurA8Args {
  const uint8_t* src;
  uint8_t* dst;
  int w, h, radius;
  int shift;     // extra inputs staged before a0 - radius (horizontal pass)
  uint32_t oob;  // alpha of the out-of-bounds colour, replicated to 4 bytes
  int pitch;     // plane row pitch in halfs
  px_t* comp;    // COMP: shadow's composite fused into the last pass: comp[i] = color MaskBlend alpha (images.nim:774-776)
  px_t color;
};

template <bool VERTICAL, int KT, bool HI, bool COMP>
__global__ void __launch_bounds__(256) blur_mma_a8_kernel(const MmaBlurA8Args a, int tilesA, int numTiles) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int IN_A = kMmaOut - 16 + 16 * KT;
  constexpr int ROWS = VERTICAL ? IN_A : kMmaLines, COLS = VERTICAL ? kMmaLines : IN_A;
  constexpr int CH = COLS / 16;                     // 16-byte pieces per tile row
  constexpr int NLD = (ROWS * CH + 255) / 256;      // pieces per thread
  constexpr int OROWS = VERTICAL ? kMmaOut : kMmaLines, OCOLS = VERTICAL ? kMmaLines : kMmaOut;
  constexpr int OPITCH = OCOLS + 16;                // bytes; keeps the fragment-order byte stores off each other's banks
  constexpr int OCH = OCOLS / 16;
  static_assert(OROWS * OCH == 256, "one 16-byte piece of the output tile per thread");
  const int pitch = a.pitch;
  __half* plane = reinterpret_cast<__half*>(smem_raw);
  uint8_t* outb = smem_raw + (size_t)ROWS * pitch * sizeof(__half);
  px_t* comp_lut = reinterpret_cast<px_t*>(outb + OROWS * OPITCH);  // COMP: color * alpha / 255 for every alpha
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (COMP) comp_lut[tid] = mul_div255(a.color, (uint32_t)tid);  // read after the loop's barriers
  const int ntaps = 2 * a.radius + 1;
  const bool vec4 = (a.w & 3) == 0 && (reinterpret_cast<uintptr_t>(a.src) & 3) == 0;
  const bool vec16 = (a.w & 15) == 0 && (reinterpret_cast<uintptr_t>(a.src) & 15) == 0;
  const bool st16 = (a.w & 15) == 0 && (reinterpret_cast<uintptr_t>(a.dst) & 15) == 0;
  const int g = lane >> 2, t = lane & 3;
  uint32_t Alo[KT][4], Ahi[HI ? KT : 1][4];
#pragma unroll
  for (int q = 0; q < KT; q++) {
#pragma unroll
    for (int rIdx = 0; rIdx < 4; rIdx++) {
      const int m = g + ((rIdx & 1) ? 8 : 0), k = 2 * t + ((rIdx & 2) ? 8 : 0);
      const int k0 = tap_at(16 * q + k - m - a.shift, ntaps), k1 = tap_at(16 * q + k + 1 - m - a.shift, ntaps);
      Alo[q][rIdx] = pack_h2((float)(k0 & 2047), (float)(k1 & 2047));
      if (HI) Ahi[q][rIdx] = pack_h2((float)((k0 >> 11) << 11), (float)((k1 >> 11) << 11));
    }
  }
  auto tile_origin = [&](int tile, int& a0, int& l0) {
    const int tl = tile / tilesA, ta = tile - tl * tilesA;
    a0 = ta * kMmaOut;
    l0 = tl * kMmaLines;
  };
  uint4 v[NLD];
  auto fetch = [&](int a0, int l0) {  // the tile's bytes -> registers
    const int x0 = VERTICAL ? l0 : a0 - a.radius - a.shift, y0 = VERTICAL ? a0 - a.radius : l0;
    if (vec16 && x0 >= 0 && x0 + COLS <= a.w && y0 >= 0 && y0 + ROWS <= a.h) {  // tile inside the image: no checks
      const uint8_t* base = a.src + (size_t)a.w * y0 + x0;
#pragma unroll
      for (int u = 0; u < NLD; u++) {
        const int c = tid + 256 * u;
        const int row = c / CH, c16 = (c - row * CH) * 16;
        if (c < ROWS * CH) v[u] = *reinterpret_cast<const uint4*>(base + (a.w * row + c16));
      }
      return;
    }
#pragma unroll
    for (int u = 0; u < NLD; u++) {
      const int c = tid + 256 * u;
      const int row = c / CH, c16 = (c - row * CH) * 16;
      const int y = y0 + row, x = x0 + c16;
      v[u] = make_uint4(a.oob, a.oob, a.oob, a.oob);
      if (c < ROWS * CH && y >= 0 && y < a.h && x + 16 > 0 && x < a.w) {
        const uint8_t* p = a.src + (size_t)a.w * y + x;
        if (vec16 && x >= 0 && x + 16 <= a.w) {
          v[u] = *reinterpret_cast<const uint4*>(p);
        } else {
          uint32_t wd[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int xj = x + 4 * j;
            wd[j] = a.oob;
            if (vec4 && xj >= 0 && xj + 4 <= a.w) {
              wd[j] = *reinterpret_cast<const uint32_t*>(p + 4 * j);
            } else if (xj + 4 > 0 && xj < a.w) {
#pragma unroll
              for (int b = 0; b < 4; b++)
                if (xj + b >= 0 && xj + b < a.w) wd[j] = (wd[j] & ~(0xFFu << (8 * b))) | ((uint32_t)p[4 * j + b] << (8 * b));
            }
          }
          v[u] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        }
      }
    }
  };
  // 0x6400 | b is the half 1024 + b.  (The subnormal encoding of the RGBX kernel measured slower here: with one
  // channel the MMAs are a larger share of the tile and they take longer on subnormal operands.)
  const uint32_t magic = 0x64006400u;
  const __half2 magic_h = *reinterpret_cast<const __half2*>(&magic);
  auto to_h2 = [&](uint32_t p, uint32_t sel) {
    const uint32_t wv = __byte_perm(p, 0x64u, sel);
    const __half2 hv = __hsub2(*reinterpret_cast<const __half2*>(&wv), magic_h);
    return *reinterpret_cast<const uint32_t*>(&hv);
  };
  auto convert = [&]() {  // registers -> the fp16 plane
#pragma unroll
    for (int u = 0; u < NLD; u++) {
      const int c = tid + 256 * u;
      if (c < ROWS * CH) {
        const int row = c / CH, c16 = (c - row * CH) * 16;
        uint4* d = reinterpret_cast<uint4*>(plane + row * pitch + c16);
        d[0] = make_uint4(to_h2(v[u].x, 0x4140), to_h2(v[u].x, 0x4342), to_h2(v[u].y, 0x4140), to_h2(v[u].y, 0x4342));
        d[1] = make_uint4(to_h2(v[u].z, 0x4140), to_h2(v[u].z, 0x4342), to_h2(v[u].w, 0x4140), to_h2(v[u].w, 0x4342));
      }
    }
  };
  auto flush = [&](int a0, int l0) {  // output tile (shared) -> global, one 16-byte piece per thread
    const int row = tid / OCH, c16 = (tid - row * OCH) * 16;
    const int y = (VERTICAL ? a0 : l0) + row, x = (VERTICAL ? l0 : a0) + c16;
    if (y < a.h && x < a.w) {
      const uint4 q = *reinterpret_cast<const uint4*>(outb + row * OPITCH + c16);
      if (COMP) {
        px_t* cp = a.comp + (size_t)a.w * y + x;
        const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
        if (st16 && x + 16 <= a.w && (reinterpret_cast<uintptr_t>(a.comp) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 4; j++)
            reinterpret_cast<uint4*>(cp)[j] = make_uint4(comp_lut[wd[j] & 255u], comp_lut[(wd[j] >> 8) & 255u],
                                                         comp_lut[(wd[j] >> 16) & 255u], comp_lut[wd[j] >> 24]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; j++)
            if (x + j < a.w) cp[j] = comp_lut[(wd[j >> 2] >> (8 * (j & 3))) & 255u];
        }
        return;
      }
      uint8_t* p = a.dst + (size_t)a.w * y + x;
      if (st16 && x + 16 <= a.w) {
        *reinterpret_cast<uint4*>(p) = q;
      } else {
        const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 16; j++)
          if (x + j < a.w) p[j] = (uint8_t)(wd[j >> 2] >> (8 * (j & 3)));
      }
    }
  };
  const int nt = warp & 3, mg = warp >> 2;
  int tile = blockIdx.x;
  int a0 = 0, l0 = 0, a0p = 0, l0p = 0;
  bool pending = false;
  if (tile < numTiles) {
    tile_origin(tile, a0, l0);
    fetch(a0, l0);
  }
#pragma unroll 1
  for (; tile < numTiles; tile += gridDim.x) {
    __syncthreads();  // nobody reads the plane of the previous tile any more; its output tile is complete
    convert();
    if (pending) flush(a0p, l0p);
    __syncthreads();  // plane ready; output tile free
    a0p = a0;
    l0p = l0;
    pending = true;
    const int a0c = a0, l0c = l0;
    (void)l0c;
    if (tile + (int)gridDim.x < numTiles) {
      tile_origin(tile + gridDim.x, a0, l0);
      fetch(a0, l0);
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
#pragma unroll
    for (int kt = 0; kt < 4 + KT - 1; kt++) {
      const int kbase = (mg * 4 + kt) * 16;
      uint32_t b0, b1;
      if (VERTICAL) ldmatrix_x2<true>(b0, b1, plane + (kbase + (lane & 15)) * pitch + nt * 8);
      else ldmatrix_x2<false>(b0, b1, plane + (nt * 8 + (lane & 7)) * pitch + kbase + ((lane & 8) ? 8 : 0));
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int q = kt - i;
        if (q >= 0 && q < KT) {
          mma_16816(acc[i], Alo[q], b0, b1);
          if (HI) mma_16816(acc[i], Ahi[q], b0, b1);
        }
      }
    }
    (void)a0c;
    // fragment slots 2hh, 2hh + 1 of m-tile i: outputs (m = g + 8 hh, n = 2t, 2t + 1) -> the output tile
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        const uint32_t t0 = __float_as_uint(__fmaf_rz(acc[i][2 * hh], kInv65280, 8388608.0f));
        const uint32_t t1 = __float_as_uint(__fmaf_rz(acc[i][2 * hh + 1], kInv65280, 8388608.0f));
        const uint32_t qq = __byte_perm(t0, t1, 0x5410);  // bytes {q0, q0 >> 8, q1, q1 >> 8}
        const int ao = (mg * 4 + i) * 16 + g + 8 * hh, lo = nt * 8 + 2 * t;  // position along the axis / line
        if (VERTICAL) {
          *reinterpret_cast<uint16_t*>(outb + ao * OPITCH + lo) = (uint16_t)__byte_perm(qq, 0u, 0x4420);
        } else {
          outb[lo * OPITCH + ao] = (uint8_t)qq;
          outb[(lo + 1) * OPITCH + ao] = (uint8_t)(qq >> 16);
        }
      }
    }
  }
  if (pending) {
    __syncthreads();
    flush(a0p, l0p);
  }
}

template <bool VERTICAL, int KT, bool HI, bool COMP>
static int launch_pass_a8(const MmaBlurA8Args& a, int tilesA, int tilesL, cudaStream_t st) {
  constexpr int IN_A = kMmaOut - 16 + 16 * KT;
  constexpr int ROWS = VERTICAL ? IN_A : kMmaLines;
  constexpr int OROWS = VERTICAL ? kMmaOut : kMmaLines, OCOLS = VERTICAL ? kMmaLines : kMmaOut;
  const size_t smem = (size_t)ROWS * a.pitch * sizeof(__half) + (size_t)OROWS * (OCOLS + 16) + (COMP ? 256 * sizeof(px_t) : 0);
  static int perSm = 0;
  if (perSm == 0) {
    if (smem > 48 * 1024)
      PX_CUDA(cudaFuncSetAttribute(blur_mma_a8_kernel<VERTICAL, KT, HI, COMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, blur_mma_a8_kernel<VERTICAL, KT, HI, COMP>, 256, smem));
    perSm = std::max(1, perSm);
  }
  const int numTiles = tilesA * tilesL;
  if (numTiles <= 0) return 0;
  blur_mma_a8_kernel<VERTICAL, KT, HI, COMP><<<std::min(numTiles, rt().num_sms * perSm), 256, smem, st>>>(a, tilesA, numTiles);
  PX_LAUNCHED();
  return 0;
}

template <bool VERTICAL, bool HI, bool COMP>
static int dispatch_a8(int KT, const MmaBlurA8Args& a, int tilesA, int tilesL, cudaStream_t st) {
  switch (KT) {
    case 2: return launch_pass_a8<VERTICAL, 2, HI, COMP>(a, tilesA, tilesL, st);
    case 3: return launch_pass_a8<VERTICAL, 3, HI, COMP>(a, tilesA, tilesL, st);
    case 4: return launch_pass_a8<VERTICAL, 4, HI, COMP>(a, tilesA, tilesL, st);
    case 5: return launch_pass_a8<VERTICAL, 5, HI, COMP>(a, tilesA, tilesL, st);
    case 6: return launch_pass_a8<VERTICAL, 6, HI, COMP>(a, tilesA, tilesL, st);
    case 7: return launch_pass_a8<VERTICAL, 7, HI, COMP>(a, tilesA, tilesL, st);
    case 8: return launch_pass_a8<VERTICAL, 8, HI, COMP>(a, tilesA, tilesL, st);
    case 9: return launch_pass_a8<VERTICAL, 9, HI, COMP>(a, tilesA, tilesL, st);
    default: return -1;
  }
}

static int upload_mma_lut(const uint16_t* lut_host, int ntaps);

// Tensor-core blur of an 8-bit plane in place (through `tmp`, a second w x h plane).  -1: radius / LUT outside the
// exact domain (the caller falls back to the RGBX path).
// comp != nullptr: the last pass writes comp[i] = color MaskBlend blurred alpha instead of the plane.
int blur_mma_a8(uint8_t* plane, uint8_t* tmp, int w, int h, const uint16_t* lut_host, int radius, uint32_t oobAlpha,
                px_t* comp, px_t color) {
  const int ntaps = 2 * radius + 1;
  if (radius < 1 || ntaps > kMaxTaps) return -1;
  unsigned long long sum = 0;
  bool hasHi = false;
  for (int i = 0; i < ntaps; i++) {
    sum += lut_host[i];
    if (lut_host[i] >= 2048) hasHi = true;
  }
  if (sum * 255ull >= (1ull << 24)) return -1;
  if (int rc = upload_mma_lut(lut_host, ntaps)) return rc;
  Runtime& r = rt();
  MmaBlurA8Args a;
  a.w = w; a.h = h; a.radius = radius;
  a.oob = (oobAlpha & 255u) * 0x01010101u;
  a.comp = comp;
  a.color = color;
  {  // X pass: plane -> tmp
    a.src = plane; a.dst = tmp;
    a.shift = ((radius + 15) & ~15) - radius;
    const int KT = (2 * radius + a.shift + 16 + 15) / 16;
    const int IN_A = kMmaOut - 16 + 16 * KT;
    a.pitch = IN_A + ((8 - IN_A) % 64 + 64) % 64;  // = 8 mod 64 halfs: ldmatrix rows 16 bytes apart in bank space
    ProfScope ps(kProfBlurX);
    const int tA = (w + kMmaOut - 1) / kMmaOut, tL = (h + kMmaLines - 1) / kMmaLines;
    const int rc = hasHi ? dispatch_a8<false, true, false>(KT, a, tA, tL, r.stream)
                         : dispatch_a8<false, false, false>(KT, a, tA, tL, r.stream);
    if (rc) return rc;
  }
  {  // Y pass: tmp -> plane
    a.src = tmp; a.dst = plane;
    a.shift = 0;
    const int KT = (2 * radius + 16 + 15) / 16;
    a.pitch = kMmaLines + 8;
    ProfScope ps(kProfBlurY);
    const int tA = (h + kMmaOut - 1) / kMmaOut, tL = (w + kMmaLines - 1) / kMmaLines;
    const int rc = comp ? (hasHi ? dispatch_a8<true, true, true>(KT, a, tA, tL, r.stream)
                                 : dispatch_a8<true, false, true>(KT, a, tA, tL, r.stream))
                        : (hasHi ? dispatch_a8<true, true, false>(KT, a, tA, tL, r.stream)
                                 : dispatch_a8<true, false, false>(KT, a, tA, tL, r.stream));
    if (rc) return rc;
  }
  return 0;
}

template <bool VERTICAL, int KT, bool HI>
static int launch_pass(const MmaBlurArgs& a, int tilesA, int tilesL, size_t smem, cudaStream_t st) {
  static size_t configured = 0;
  static int perSm = 1;
  if (configured != smem) {
    if (smem > 48 * 1024)
      PX_CUDA(cudaFuncSetAttribute(blur_mma_kernel<VERTICAL, KT, HI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, blur_mma_kernel<VERTICAL, KT, HI>, 256, smem));
    perSm = std::max(1, perSm);
    configured = smem;
  }
  const int numTiles = tilesA * tilesL;
  if (numTiles <= 0) return 0;
  const int grid = std::min(numTiles, rt().num_sms * perSm);
  blur_mma_kernel<VERTICAL, KT, HI><<<grid, 256, smem, st>>>(a, tilesA, numTiles);
  PX_LAUNCHED();
  return 0;
}

// phase 0: both passes; 1: X pass only (rows [sy0, sy1) -> tmp); 2: Y pass only (tmp -> image rows [y0, y1))
template <int KT, bool HI>
static int launch_both(MmaBlurArgs a, Image* im, void* tmp, int y0, int y1, int phase) {
  Runtime& r = rt();
  constexpr int IN_A = kMmaOut - 16 + 16 * KT;
  const size_t rawBytes = (size_t)IN_A * kMmaLines * 4;
  if (phase != 2) {  // X pass: image -> tmp, rows [sy0, sy1).  Plane [line][a]: pitch = 8 mod 64 halfs keeps ldmatrix conflict-free
    a.src = (const px_t*)im->data; a.dst = (px_t*)tmp;
    a.pitch = IN_A + ((8 - IN_A) % 64 + 64) % 64;
    const size_t smem = (size_t)4 * kMmaLines * a.pitch * sizeof(__half) + rawBytes;
    ProfScope ps(kProfBlurX);
    if (int rc = launch_pass<false, KT, HI>(a, (im->w + kMmaOut - 1) / kMmaOut, (a.sy1 - a.sy0 + kMmaLines - 1) / kMmaLines, smem,
                                        r.stream))
      return rc;
  }
  if (phase != 1) {  // Y pass: tmp -> image rows [y0, y1).  Plane [a][line]: 32-half rows, 16-byte chunks XOR-swizzled by row
    a.src = (const px_t*)tmp; a.dst = (px_t*)im->data;
    a.pitch = kMmaLines;  // no padding: chunks are swizzled
    const size_t smem = (size_t)4 * IN_A * a.pitch * sizeof(__half) + rawBytes;
    ProfScope ps(kProfBlurY);
    if (int rc = launch_pass<true, KT, HI>(a, (y1 - y0 + kMmaOut - 1) / kMmaOut, (im->w + kMmaLines - 1) / kMmaLines, smem, r.stream))
      return rc;
  }
  return 0;
}

static int upload_mma_lut(const uint16_t* lut_host, int ntaps) {
  Runtime& r = rt();
  void* pin;
  if (int rc = staging_acquire(sizeof(uint16_t) * (kMaxTaps + 3), &pin)) return rc;
  memset(pin, 0, sizeof(uint16_t) * (kMaxTaps + 3));
  memcpy(pin, lut_host, (size_t)ntaps * 2);
  PX_CUDA(cudaMemcpyToSymbolAsync(c_mma_lut, pin, sizeof(uint16_t) * (kMaxTaps + 3), 0, cudaMemcpyHostToDevice, r.stream));
  return staging_release();
}

// Tensor-core blur of rows [y0, y1) of `im` through the scratch plane `tmp`.  Returns -1 when the radius / LUT is
// outside what this path holds exactly (the caller then takes the CUDA-core kernels).
int blur_mma(Image* im, void* tmp, const uint16_t* lut_host, int radius, uint32_t oob, int y0, int y1, int phase) {
  const int ntaps = 2 * radius + 1;
  if (radius < 1 || ntaps > kMaxTaps) return -1;
  unsigned long long sum = 0;
  int hasHi = 0;
  for (int i = 0; i < ntaps; i++) {
    sum += lut_host[i];
    if (lut_host[i] >= 2048) hasHi = 1;
  }
  if (sum * 255ull >= (1ull << 24)) return -1;  // partial sums must stay exact in fp32
  if (int rc = upload_mma_lut(lut_host, ntaps)) return rc;
  MmaBlurArgs a;
  a.w = im->w; a.h = im->h; a.radius = radius; a.oob = oob; a.hasHi = hasHi;
  a.y0 = y0; a.y1 = y1;
  a.sy0 = std::max(0, y0 - radius);
  a.sy1 = std::min(im->h, y1 + radius);
  if (phase == 1) {  // X pass of exactly the rows asked for
    a.sy0 = y0;
    a.sy1 = y1;
  }
  a.pitch = 0; a.src = nullptr; a.dst = nullptr;
  const int KT = (2 * radius + 16 + 15) / 16;
  switch (KT) {
    case 2: return hasHi ? launch_both<2, true>(a, im, tmp, y0, y1, phase) : launch_both<2, false>(a, im, tmp, y0, y1, phase);
    case 3: return hasHi ? launch_both<3, true>(a, im, tmp, y0, y1, phase) : launch_both<3, false>(a, im, tmp, y0, y1, phase);
    case 4: return hasHi ? launch_both<4, true>(a, im, tmp, y0, y1, phase) : launch_both<4, false>(a, im, tmp, y0, y1, phase);
    case 5: return hasHi ? launch_both<5, true>(a, im, tmp, y0, y1, phase) : launch_both<5, false>(a, im, tmp, y0, y1, phase);
    case 6: return hasHi ? launch_both<6, true>(a, im, tmp, y0, y1, phase) : launch_both<6, false>(a, im, tmp, y0, y1, phase);
    case 7: return hasHi ? launch_both<7, true>(a, im, tmp, y0, y1, phase) : launch_both<7, false>(a, im, tmp, y0, y1, phase);
    case 8: return hasHi ? launch_both<8, true>(a, im, tmp, y0, y1, phase) : launch_both<8, false>(a, im, tmp, y0, y1, phase);
    case 9: return hasHi ? launch_both<9, true>(a, im, tmp, y0, y1, phase) : launch_both<9, false>(a, im, tmp, y0, y1, phase);
    default: return -1;
  }
}

}  // namespace pixie

// K5c — the Gaussian blur (images.nim:304-365) as ONE fused pass on Blackwell's 5th-generation tensor cores.
//
// The two-pass kernels of blur_mma.cu move 16 B/px through HBM and spend their issue slots on mma.sync, ldmatrix and
// staging.  Here a CTA owns a strip of 128 output columns and marches down the image in blocks of 32 rows:
//
//   TMA (cp.async.bulk.tensor.2d) brings the raw RGBX rows of the block, 192 columns wide, into shared memory;
//   the threads split them into four planar fp16 planes (a byte next to a zero byte IS the fp16 subnormal b * 2^-24:
//     byte permutes only) laid out as a K-major, 128-byte-swizzled tcgen05 operand:  A_x[line = (channel, row)][k = x];
//   X pass:  D_x[line][n] = sum_k A_x[line][k] * T[k][n]   (tcgen05.mma kind::f16, M = 128 lines, N = 64 outputs,
//     K = 128 inputs, accumulator in TMEM), T = the banded Toeplitz matrix of the LUT, T[k][n] = lut[k - n - (32 - r)],
//     an fp16 operand in shared memory built once per CTA (taps >= 2048 split into two exact fp16 parts);
//   X epilogue: tcgen05.ld, `div 256 div 255` as one FFMA.RZ (see blur_mma.cu), and the 8-bit row goes — again as fp16
//     subnormals — straight into a RING of 96 X-blurred rows in shared memory, which is at the same time the MN-major
//     operand of the Y pass (rows = K): the intermediate image of the reference (images.nim:341) never exists in HBM;
//   Y pass:  D_y[line = x][n] = sum_k ring[k][x] * T[k][n]  per channel (M = 128 columns, N = 32 output rows, K = 96);
//   Y epilogue: tcgen05.ld of the four channels, quantise, pack RGBX, 128-byte coalesced stores.
//
// One thread issues the MMAs and goes on; they run while all 8 warps convert the next block / drain the previous
// accumulators.  Exactness is that of blur_mma.cu: every operand is an integer (times 2^-24) exact in fp16, every
// partial sum an integer below 2^24 (the caller checks sum(lut) * 255 < 2^24), the fp32 accumulator never rounds, so
// the result is bit-identical to the reference's uint32 arithmetic.  Rows outside the image enter the ring as the
// out-of-bounds colour itself (the reference's Y pass reads the colour, not an X-blurred row of it).
// HBM traffic: the image is read 1.5 times (64 halo columns per strip; neighbouring strips run at the same time and
// share them through L2) and written once.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace pixie {

using namespace umma;

constexpr int kTcStripW = 128;            // output columns per strip
constexpr int kTcRows = 32;               // rows per marching step
constexpr int kTcInW = kTcStripW + 64;    // input columns per strip (radius <= 32 on either side)
constexpr int kTcRingRows = 128;          // four blocks of 32 rows: the Y window (three) + the block being written
constexpr uint32_t kToeBytes = 2 * 64 * 128;            // Toeplitz operand: 2 K-blocks of [64 n][128 B]
constexpr uint32_t kAxBlock = 128 * 128;                // one K-block of A_x: [128 lines][128 B = 64 columns]
constexpr uint32_t kAxBytes = 3 * kAxBlock;
constexpr uint32_t kRingBlock = kTcRingRows * 128;      // [96 rows][128 B = 64 columns] of one channel
constexpr uint32_t kRingBytes = 8 * kRingBlock;         // 4 channels x 2 column blocks
constexpr uint32_t kRawBytes = kTcRows * kTcInW * 4;
constexpr int kTcMaxRadius = 32;

#define kTcInv65280s __uint_as_float(0x43808081u)  // float(2^24 / 65280), see blur_mma.cu

struct TcBlurArgs {
  px_t* dst;
  int w, h, radius;
  uint32_t oob;
  int y0, y1;  // output rows
  int strips, chunks, chunkRows;
  unsigned* ticket;
  unsigned long long* dbg;  // PIXIE_CUDA_TC_DEBUG: per-phase cycle sums of worker warp 0 and the two issuers
};

__constant__ uint16_t c_tc_lut[2 * kTcMaxRadius + 1 + 3];

PXD uint32_t tc_quant(uint32_t accBits) {  // (acc * 2^24) div 65280 in the low byte of the result
  return __float_as_uint(__fmaf_rz(__uint_as_float(accBits), kTcInv65280s, 8388608.0f));
}

// Roles: warps 0..15 are workers (convert, X epilogue, Y epilogue; TMEM lane quarter = warp % 4); lane 0 of warp 16
// issues the TMA loads and the X MMAs, lane 0 of warp 17 the Y MMAs.  A tcgen05.mma with both operands in shared
// memory takes max(N / 2, (4096 + 32 N) / 128) cycles — its operand fetch runs at 128 B/clk (measured,
// tools/umma_probe.cu: 40 cycles at N = 32, 48 at N = 64) — and one thread cannot issue faster than one per ~52
// cycles, so two issuers keep the pipe fed and do nothing else.  The roles meet only through mbarriers:
//   barRaw   TMA -> workers        raw block landed (transaction bytes)
//   barAx    workers -> X issuer   A_x planes of block i written (so: raw buffer free, D_x drained)
//   barX     X issuer -> workers   X MMAs of block i complete (tcgen05.commit): D_x readable, A_x free
//   barRing  workers -> Y issuer   X epilogue of block i finished: its 32 rows are in ring slot i % 4
//   barY     Y issuer -> workers   Y MMAs complete: D_y readable
//   barYFree workers -> Y issuer   Y epilogue has loaded D_y
// The ring has four slots of 32 rows: the Y MMAs of block i - 1 read slots i - 3 .. i - 1 while the workers convert
// block i and drain X(i) into slot i % 4; D_y is double-buffered so that the workers drain Y(i - 2) while X(i) and
// Y(i - 1) compute, and the X MMAs commit per output column block so that its epilogue starts while the other runs.
constexpr int kTcWorkers = 512;

__global__ void __launch_bounds__(kTcWorkers + 64, 1) blur_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcBlurArgs a) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t* sToe = base;
  uint8_t* sAx = sToe + kToeBytes;
  uint8_t* sRing = sAx + kAxBytes;
  uint8_t* sRaw = sRing + kRingBytes;
  __shared__ uint64_t barRaw, barAx, barX[2], barRing, barY[2], barYFree[2];
  __shared__ uint32_t tmemSlot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  {  // ---- Toeplitz operand, K-major: row n holds T[k][n] for k = 0..127, two halfs per store
    const int shift = kTcMaxRadius - a.radius, ntaps = 2 * a.radius + 1;
    for (int i = tid; i < 64 * 64; i += kTcWorkers + 64) {
      const int n = i >> 6, k = (i & 63) * 2;
      const int t0 = k - n - shift, t1 = t0 + 1;
      const int v0 = (t0 >= 0 && t0 < ntaps) ? (int)c_tc_lut[t0] : 0, v1 = (t1 >= 0 && t1 < ntaps) ? (int)c_tc_lut[t1] : 0;
      const uint32_t off = (uint32_t)(k >> 6) * (64u * 128u) + sw128_off((uint32_t)n, (uint32_t)(k & 63) >> 3) + (uint32_t)(k & 7) * 2u;
      *reinterpret_cast<__half2*>(sToe + off) = __halves2half2(__int2half_rn(v0), __int2half_rn(v1));  // taps < 2048: exact
    }
  }
  if (tid == 0) {
    mbar_init(&barRaw, 1);
    mbar_init(&barAx, kTcWorkers / 32);
    mbar_init(&barX[0], 1);
    mbar_init(&barX[1], 1);
    mbar_init(&barRing, kTcWorkers / 32);
    mbar_init(&barY[0], 1);
    mbar_init(&barY[1], 1);
    mbar_init(&barYFree[0], kTcWorkers / 32);
    mbar_init(&barYFree[1], kTcWorkers / 32);
    fence_barrier_init();
    tma_prefetch_desc(&tmap);
  }
  if (warp == 0) tmem_alloc(&tmemSlot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmemSlot;  // D_x: columns [0, 128); D_y (two buffers): [128, 256), [256, 384)
  const int total = a.strips * a.chunks;

  if (warp == kTcWorkers / 32) {
    // ================================================================ X issuer (+ TMA)
    if (lane == 0) {
      const uint32_t sToeA = smem_u32(sToe), sAxA = smem_u32(sAx);
      constexpr uint32_t idescX = idesc_f16(128, 64, false, false);
      uint64_t bd[8], ad[2][8];
#pragma unroll
      for (int s = 0; s < 8; s++) {
        bd[s] = smem_desc_sw128(sToeA + (uint32_t)(s >> 2) * (64u * 128u) + (uint32_t)(s & 3) * 32u, 16, 1024);
#pragma unroll
        for (int j = 0; j < 2; j++)  // output column block j reads input column blocks j, j + 1
          ad[j][s] = smem_desc_sw128(sAxA + (uint32_t)(j + (s >> 2)) * kAxBlock + (uint32_t)(s & 3) * 32u, 16, 1024);
      }
      uint32_t pAx = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int chunk = t / a.strips, strip = t - chunk * a.strips;
        const int x0 = strip * kTcStripW;
        const int cy0 = a.y0 + chunk * a.chunkRows, cy1 = min(a.y1, cy0 + a.chunkRows);
        const int nb = (cy1 - cy0 + kTcRows - 1) / kTcRows + 2;
        const int rowBase = cy0 - kTcRows;
        mbar_arrive_expect_tx(&barRaw, kRawBytes);  // the raw buffer is free: every earlier block has been converted
        tma_load_2d(sRaw, &tmap, &barRaw, x0 - 32, rowBase);
#pragma unroll 1
        for (int i = 0; i < nb; i++) {
          const long long w0_ = a.dbg ? clock64() : 0;
          mbar_wait(&barAx, pAx);
          pAx ^= 1;
          const long long w1_ = a.dbg ? clock64() : 0;
          tc_fence_after_sync();
          if (i + 1 < nb) {
            mbar_arrive_expect_tx(&barRaw, kRawBytes);
            tma_load_2d(sRaw, &tmap, &barRaw, x0 - 32, rowBase + kTcRows * (i + 1));
          }
#pragma unroll
          for (int j = 0; j < 2; j++) {  // one commit per output column block: its X epilogue starts while the other computes
#pragma unroll
            for (int s = 0; s < 8; s++) mma_f16_ss(tmem + (uint32_t)(j * 64), ad[j][s], bd[s], idescX, s > 0 ? 1u : 0u);
            mma_commit(&barX[j]);
          }
          if (a.dbg) { atomicAdd(&a.dbg[16], (unsigned long long)(w1_ - w0_)); atomicAdd(&a.dbg[17], (unsigned long long)(clock64() - w1_)); }
        }
      }
    }
  } else if (warp == kTcWorkers / 32 + 1) {
    // ================================================================ Y issuer
    if (lane == 0) {
      const uint32_t sToeA = smem_u32(sToe), sRingA = smem_u32(sRing);
      constexpr uint32_t idescY = idesc_f16(128, 32, true, false);
      uint64_t bd[6];  // the Y pass's Toeplitz matrix is the top-left corner (32 x 96) of the X pass's
#pragma unroll
      for (int s = 0; s < 6; s++) bd[s] = smem_desc_sw128(sToeA + (uint32_t)(s >> 2) * (64u * 128u) + (uint32_t)(s & 3) * 32u, 16, 1024);
      const uint64_t ad0 = smem_desc_sw128(sRingA, kRingBlock, 1024);  // + ring row * 128 / 16 per K step, + 2 blocks per channel
      uint32_t pRing = 0, pYFree[2] = {0, 0};
      uint32_t g = 0;  // running number of the Y block: accumulator buffer g & 1
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int chunk = t / a.strips;
        const int cy0 = a.y0 + chunk * a.chunkRows, cy1 = min(a.y1, cy0 + a.chunkRows);
        const int nb = (cy1 - cy0 + kTcRows - 1) / kTcRows + 2;
#pragma unroll 1
        for (int k = 0; k < nb; k++) {
          const long long w0_ = a.dbg ? clock64() : 0;
          mbar_wait(&barRing, pRing);
          pRing ^= 1;
          if (a.dbg) atomicAdd(&a.dbg[18], (unsigned long long)(clock64() - w0_));
          if (k < 2) continue;
          const long long w1_ = a.dbg ? clock64() : 0;
          const uint32_t b = g & 1u;
          if (g >= 2) {  // the epilogue of Y block g - 2 has loaded this buffer
            mbar_wait(&barYFree[b], pYFree[b]);
            pYFree[b] ^= 1;
          }
          g++;
          tc_fence_after_sync();
          const uint32_t start = (uint32_t)(kTcRows * ((k - 2) & 3));  // ring blocks k - 2, k - 1, k
#pragma unroll
          for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int s = 0; s < 6; s++) {
              const uint32_t row = (start + 16u * (uint32_t)s) & (uint32_t)(kTcRingRows - 1);
              const uint64_t ad = ad0 + (uint64_t)(((uint32_t)(c * 2) * kRingBlock + row * 128u) >> 4);
              mma_f16_ss(tmem + 128u + 128u * b + (uint32_t)(c * 32), ad, bd[s], idescY, s > 0 ? 1u : 0u);
            }
          }
          mma_commit(&barY[b]);
          if (a.dbg) atomicAdd(&a.dbg[19], (unsigned long long)(clock64() - w1_));
        }
      }
    }
  } else {
    // ================================================================ workers
    uint32_t pRaw = 0, pX = 0, pY[2] = {0, 0}, g = 0;
    long long tm[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc0 = 0;
#define TC_MARK(slot) if (a.dbg) { const long long n_ = clock64(); tm[slot] += n_ - tc0; tc0 = n_; }
    const int q = warp & 3, wg = warp >> 2;  // TMEM lane quarter; which quarter of the columns / rows this warp drains
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int chunk = t / a.strips, strip = t - chunk * a.strips;
      const int x0 = strip * kTcStripW;
      const int cy0 = a.y0 + chunk * a.chunkRows, cy1 = min(a.y1, cy0 + a.chunkRows);
      const int outBlocks = (cy1 - cy0 + kTcRows - 1) / kTcRows, nb = outBlocks + 2;
      const int rowBase = cy0 - kTcRows;  // first input row of block 0
      const bool patch = a.oob != 0u && (x0 - 32 < 0 || x0 + kTcInW - 32 > a.w);

      // Y epilogue of output block `ob`: lane = column, 8 rows per warp, four channels from four accumulators
      auto y_epilogue = [&](int ob) {
        const uint32_t b = g & 1u;
        g++;
        if (a.dbg) { const long long n_ = clock64(); tm[2] += n_ - tc0; tc0 = n_; }
        mbar_wait(&barY[b], pY[b]);
        pY[b] ^= 1;
        if (a.dbg) { const long long n_ = clock64(); tm[5] += n_ - tc0; tc0 = n_; }
        tc_fence_after_sync();
        const int x = x0 + 32 * q + lane;
        const int orow0 = cy0 + kTcRows * ob + 8 * wg;
        uint32_t v[4][8];
#pragma unroll
        for (int c = 0; c < 4; c++) tmem_ld8(tmem + ((uint32_t)(32 * q) << 16) + 128u + 128u * b + (uint32_t)(c * 32 + 8 * wg), v[c]);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&barYFree[b]);  // this D_y buffer is in registers: Y block g + 1 may overwrite it
        if (x < a.w) {
          px_t* p = a.dst + (size_t)a.w * (size_t)orow0 + x;
          const int rows = min(8, cy1 - orow0);
#pragma unroll
          for (int rr = 0; rr < 8; rr++) {
            const uint32_t rg = __byte_perm(tc_quant(v[0][rr]), tc_quant(v[1][rr]), 0x0040);
            const uint32_t ba = __byte_perm(tc_quant(v[2][rr]), tc_quant(v[3][rr]), 0x0040);
            if (rr < rows) *p = __byte_perm(rg, ba, 0x5410);
            p += a.w;
          }
        }
      };

#pragma unroll 1
      for (int i = 0; i < nb; i++) {
        // ---- raw RGBX block -> planar fp16 planes A_x
        if (a.dbg) tc0 = clock64();
        mbar_wait(&barRaw, pRaw);
        pRaw ^= 1;
        TC_MARK(0)
#pragma unroll
        for (int u = 0; u < 3; u++) {
          const int qd = tid + kTcWorkers * u;
          const int row = qd / 48, g = qd - row * 48;
          uint4 p = *reinterpret_cast<const uint4*>(sRaw + ((size_t)row * kTcInW + 4 * g) * 4);
          if (patch) {  // columns outside the image carry the out-of-bounds colour (TMA filled them with zeros)
            const int x = x0 - 32 + 4 * g;
            if (x < 0 || x >= a.w) p.x = a.oob;
            if (x + 1 < 0 || x + 1 >= a.w) p.y = a.oob;
            if (x + 2 < 0 || x + 2 >= a.w) p.z = a.oob;
            if (x + 3 < 0 || x + 3 >= a.w) p.w = a.oob;
          }
          const uint32_t rg01 = __byte_perm(p.x, p.y, 0x5140), ba01 = __byte_perm(p.x, p.y, 0x7362);
          const uint32_t rg23 = __byte_perm(p.z, p.w, 0x5140), ba23 = __byte_perm(p.z, p.w, 0x7362);
          uint32_t wv[8];
          wv[0] = __byte_perm(rg01, 0u, 0x4140); wv[1] = __byte_perm(rg23, 0u, 0x4140);  // r0 r1 | r2 r3 as fp16 subnormals
          wv[2] = __byte_perm(rg01, 0u, 0x4342); wv[3] = __byte_perm(rg23, 0u, 0x4342);  // g
          wv[4] = __byte_perm(ba01, 0u, 0x4140); wv[5] = __byte_perm(ba23, 0u, 0x4140);  // b
          wv[6] = __byte_perm(ba01, 0u, 0x4342); wv[7] = __byte_perm(ba23, 0u, 0x4342);  // a
          // line = channel * 32 + row; column block g / 16, 16-byte chunk (g % 16) / 2, upper or lower half of it
          uint8_t* d = sAx + (uint32_t)(g >> 4) * kAxBlock + sw128_off((uint32_t)row, (uint32_t)(g & 15) >> 1) + (uint32_t)(g & 1) * 8u;
#pragma unroll
          for (int c = 0; c < 4; c++) *reinterpret_cast<uint2*>(d + c * (32 * 128)) = make_uint2(wv[2 * c], wv[2 * c + 1]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&barAx);
        TC_MARK(1)
        // ---- while the X MMAs of this block run: drain the Y accumulators of block i - 2 (output block i - 4)
        if (i >= 4) y_epilogue(i - 4);
        TC_MARK(2)
        // ---- X epilogue: quantised rows -> ring slot i % 4.  Warp = (channel q, 32 of the 128 output columns).
        mbar_wait(&barX[wg >> 1], pX);
        pX ^= 1;
        TC_MARK(3)
        tc_fence_after_sync();
        {
          const int grow = rowBase + kTcRows * i + lane;  // lane = row of the block
          const uint32_t ringRow = (uint32_t)(kTcRows * (i & 3) + lane);
          uint8_t* dplane = sRing + (uint32_t)(q * 2 + (wg >> 1)) * kRingBlock;
          const bool rowOut = grow < 0 || grow >= a.h;
          const uint32_t oobc = (a.oob >> (8 * q)) & 255u, oobw = oobc | (oobc << 16);
          uint32_t v0[16], v1[16];
          tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * wg), v0);
          tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * wg + 16), v1);
          tmem_ld_wait();
          uint32_t w0[8], w1[8];
#pragma unroll
          for (int k = 0; k < 8; k++) {
            w0[k] = rowOut ? oobw : __byte_perm(tc_quant(v0[2 * k]), tc_quant(v0[2 * k + 1]), 0x5410);
            w1[k] = rowOut ? oobw : __byte_perm(tc_quant(v1[2 * k]), tc_quant(v1[2 * k + 1]), 0x5410);
          }
          const uint32_t c0 = (uint32_t)(4 * (wg & 1));  // 16-byte chunk of the 128-byte row
          *reinterpret_cast<uint4*>(dplane + sw128_off(ringRow, c0)) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
          *reinterpret_cast<uint4*>(dplane + sw128_off(ringRow, c0 + 1)) = make_uint4(w0[4], w0[5], w0[6], w0[7]);
          *reinterpret_cast<uint4*>(dplane + sw128_off(ringRow, c0 + 2)) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
          *reinterpret_cast<uint4*>(dplane + sw128_off(ringRow, c0 + 3)) = make_uint4(w1[4], w1[5], w1[6], w1[7]);
        }
        tc_fence_before_sync();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&barRing);
        TC_MARK(4)
      }
      if (nb >= 4) y_epilogue(nb - 4);  // Y of block nb - 2
      y_epilogue(nb - 3);               // Y of the last block
    }
    if (a.dbg && lane == 0 && (warp == 0 || warp == 9))
      for (int k = 0; k < 6; k++) atomicAdd(&a.dbg[(warp ? 8 : 0) + k], (unsigned long long)tm[k]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int launch_tc(const CUtensorMap& tmap, const TcBlurArgs& a, int blocks, cudaStream_t st) {
  const size_t smem = 1024 + kToeBytes + kAxBytes + kRingBytes + kRawBytes;
  static bool configured = false;
  if (!configured) {
    PX_CUDA(cudaFuncSetAttribute(blur_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  blur_tc_kernel<<<blocks, kTcWorkers + 64, smem, st>>>(tmap, a);
  PX_LAUNCHED();
  return 0;
}

// Fused tensor-core blur of rows [y0, y1) of `src` (w x h RGBX) into `dst` (same geometry, a different buffer).
// -1: outside this kernel's domain (radius > 32, LUT not exact in fp32, width not a multiple of 4, no TMA entry point).
int blur_tc(const px_t* src, px_t* dst, int w, int h, const uint16_t* lut_host, int radius, uint32_t oob, int y0, int y1) {
  if (radius < 1 || radius > kTcMaxRadius || (w & 3) != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0) return -1;
  const int ntaps = 2 * radius + 1;
  unsigned long long sum = 0;
  bool hasHi = false;
  for (int i = 0; i < ntaps; i++) {
    sum += lut_host[i];
    if (lut_host[i] >= 2048) hasHi = true;
  }
  if (sum * 255ull >= (1ull << 24) || hasHi) return -1;  // taps >= 2048 (radii below 29) would need two fp16 parts: blur_mma.cu
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return -1;
  Runtime& r = rt();
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
  const cuuint64_t strides[1] = {(cuuint64_t)w * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kTcInW, (cuuint32_t)kTcRows};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<px_t*>(src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return -1;
  {  // LUT -> constant memory through the pinned staging buffer
    void* pin;
    if (int rc = staging_acquire(sizeof(uint16_t) * (2 * kTcMaxRadius + 4), &pin)) return rc;
    memset(pin, 0, sizeof(uint16_t) * (2 * kTcMaxRadius + 4));
    memcpy(pin, lut_host, (size_t)ntaps * 2);
    PX_CUDA(cudaMemcpyToSymbolAsync(c_tc_lut, pin, sizeof(uint16_t) * (2 * kTcMaxRadius + 4), 0, cudaMemcpyHostToDevice, r.stream));
    if (int rc = staging_release()) return rc;
  }
  TcBlurArgs a;
  a.dst = dst; a.w = w; a.h = h; a.radius = radius; a.oob = oob; a.y0 = y0; a.y1 = y1;
  a.strips = (w + kTcStripW - 1) / kTcStripW;
  // row chunks: a chunk re-blurs 64 warm-up rows horizontally, so long chunks are cheaper; but the tickets
  // (strip, chunk) are handed out dynamically to one CTA per SM and should outnumber the SMs a few times over
  const int rows = y1 - y0;
  int chunkRows = 1024;
  while (chunkRows > 128 && (long long)a.strips * ((rows + chunkRows - 1) / chunkRows) < 2ll * r.num_sms) chunkRows /= 2;
  a.chunkRows = chunkRows;
  a.chunks = (rows + chunkRows - 1) / chunkRows;
  a.ticket = nullptr;  // tickets are dealt round-robin: every (strip, chunk) costs the same
  a.dbg = nullptr;
  static const bool dbgOn = getenv("PIXIE_CUDA_TC_DEBUG") != nullptr;
  if (dbgOn) {
    void* d;
    if (int rc = get_scratch(3, 32 * 8, &d)) return rc;
    PX_CUDA(cudaMemsetAsync(d, 0, 32 * 8, r.stream));
    a.dbg = (unsigned long long*)d;
  }
  const int blocks = std::min(a.strips * a.chunks, r.num_sms);
  ProfScope ps(kProfBlurX);
  const int rcl = launch_tc(tmap, a, blocks, r.stream);
  if (dbgOn && rcl == 0) {
    unsigned long long hd[32];
    PX_CUDA(cudaMemcpyAsync(hd, a.dbg, sizeof(hd), cudaMemcpyDeviceToHost, r.stream));
    PX_CUDA(cudaStreamSynchronize(r.stream));
    const double nbk = (double)a.strips * a.chunks * ((double)a.chunkRows / kTcRows + 2) ;
    const char* nm[6] = {"wait raw", "convert", "y epilogue", "wait X", "x epilogue", "wait Y"};
    fprintf(stderr, "[blur_tc] cycles per block (sum over CTAs / blocks): ");
    for (int k = 0; k < 6; k++) fprintf(stderr, "w0 %s %.0f | ", nm[k], hd[k] / nbk);
    for (int k = 0; k < 6; k++) fprintf(stderr, "w9 %s %.0f | ", nm[k], hd[8 + k] / nbk);
    fprintf(stderr, "Xiss wait %.0f issue %.0f | Yiss wait %.0f issue %.0f\n", hd[16] / nbk, hd[17] / nbk, hd[18] / nbk, hd[19] / nbk);
  }
  return rcl;
}

}  // namespace pixie

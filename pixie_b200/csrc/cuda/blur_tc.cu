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
constexpr uint32_t kToeBytes = 3 * 128 * 128;           // Toeplitz operand: 3 K-blocks of [128 m][128 B]
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
  // row bands across GPUs: rows [0, y0) / [y1, h) are halo rows a neighbour GPU stores into this image; a ticket that
  // reads them first waits until the flag (written by that neighbour after its rows) has reached `epoch`
  const unsigned* flagTop;
  const unsigned* flagBottom;
  unsigned epoch;
};

// tickets in the order interior chunks first, the two edge chunks (the only ones that read halo rows) last
PXD int tc_chunk_of(int order, int chunks) {
  if (chunks < 3) return order;
  return order < chunks - 2 ? order + 1 : (order == chunks - 2 ? 0 : chunks - 1);
}

__constant__ uint16_t c_tc_lut[2 * kTcMaxRadius + 1 + 3];

PXD unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
PXD uint32_t tc_quant(uint32_t accBits) {  // (acc * 2^24) div 65280 in the low byte of the result
  return __float_as_uint(__fmaf_rz(__uint_as_float(accBits), kTcInv65280s, 8388608.0f));
}

// Roles: warps 0..15 are workers (convert, X epilogue, Y epilogue; TMEM lane quarter = warp % 4); lane 0 of warp 16
// issues the TMA loads and the X MMAs, lane 0 of warps 17 / 18 the Y MMAs of channels 0-1 / 2-3.  What limits this
// kernel is the shared-memory port (128 B/clk, shared by the threads' loads / stores, TMA writes and the tensor
// pipe's operand fetch — a tcgen05.mma with both operands in shared memory takes max(N / 2, (4096 + 32 N) / 128)
// cycles, tools/umma_probe.cu) and the rate at which one thread can issue MMAs (~52 cycles each), hence: the ring of
// X-blurred rows lives in TENSOR MEMORY (the Y pass's A operand is read from TMEM, nothing of it crosses the
// shared-memory port), three issuers, and workers that never wait for the tensor pipe in steady state.
// The roles meet only through mbarriers:
//   barRaw[b]  TMA -> workers         raw block landed in buffer b (transaction bytes)
//   barAx[b]   workers -> X issuer    A_x planes of a block written to buffer b (so: raw buffer b is free)
//   barX       X issuer -> workers    X MMAs of a block complete (tcgen05.commit): D_x readable, its A_x buffer free
//   barRing    workers -> issuers     X epilogue finished: the block's rows are in the TMEM ring, D_x is free
//   barY       Y issuers -> workers   both halves of the Y MMAs complete: D_y readable
//   barYFree   workers -> Y issuers   Y epilogue has loaded D_y
// Worker step i: convert block i, X epilogue of block i - 1, Y epilogue of block i - 2 — by then the tensor pipe has
// finished what each of them needs.  The ring has four slots of 32 rows: the Y MMAs of block k read slots k - 2 .. k
// while the X epilogue of block k + 1 writes slot (k + 1) % 4.
constexpr int kTcWorkers = 512;
constexpr int kTcThreads = kTcWorkers + 96;
constexpr uint32_t kTmDx = 0, kTmRing = 128, kTmDy = 384;  // TMEM columns: D_x 128, ring 4 channels x 64 (128 rows / 2), D_y 4 x 32

template <bool DBG>
__global__ void __launch_bounds__(kTcThreads, 1) blur_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcBlurArgs a) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t* sToe = base;                  // [128 m][192 k] fp16, K-major: 3 K-blocks of [128][128 B]
  uint8_t* sAx = sToe + kToeBytes;       // 2 buffers of 3 K-blocks [128 lines][128 B]
  uint8_t* sRaw = sAx + 2 * kAxBytes;    // 2 buffers of [32 rows][192 px] RGBX
  __shared__ uint64_t barRaw[2], barAx[2], barX, barRing, barY, barYFree, barTicket[2];
  __shared__ uint32_t tmemSlot;
  __shared__ int sTicket[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  {  // ---- Toeplitz operand T[m][k] = lut[k - m - (32 - r)], K-major rows m = 0..127, k = 0..191, two halfs per store
    const int shift = kTcMaxRadius - a.radius, ntaps = 2 * a.radius + 1;
    for (int i = tid; i < 128 * 96; i += kTcThreads) {
      const int m = i / 96, k = (i - m * 96) * 2;
      const int t0 = k - m - shift, t1 = t0 + 1;
      const int v0 = (t0 >= 0 && t0 < ntaps) ? (int)c_tc_lut[t0] : 0, v1 = (t1 >= 0 && t1 < ntaps) ? (int)c_tc_lut[t1] : 0;
      const uint32_t off = (uint32_t)(k >> 6) * kAxBlock + sw128_off((uint32_t)m, (uint32_t)(k & 63) >> 3) + (uint32_t)(k & 7) * 2u;
      *reinterpret_cast<__half2*>(sToe + off) = __halves2half2(__int2half_rn(v0), __int2half_rn(v1));  // taps < 2048: exact
    }
  }
  if (tid == 0) {
    mbar_init(&barRaw[0], 1);
    mbar_init(&barRaw[1], 1);
    mbar_init(&barAx[0], kTcWorkers / 32);
    mbar_init(&barAx[1], kTcWorkers / 32);
    mbar_init(&barX, 1);
    mbar_init(&barRing, kTcWorkers / 32);
    mbar_init(&barY, 2);
    mbar_init(&barYFree, kTcWorkers / 32);
    mbar_init(&barTicket[0], 1);
    mbar_init(&barTicket[1], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap);
  }
  if (warp == 0) tmem_alloc(&tmemSlot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmemSlot;
  const int total = a.strips * a.chunks;

  if (warp == kTcWorkers / 32) {
    // ================================================================ X issuer (+ TMA)
    if (lane == 0) {
      const uint32_t sToeA = smem_u32(sToe), sAxA = smem_u32(sAx);
      constexpr uint32_t idescX = idesc_f16(128, 128, false, false);
      uint64_t ad[12], bd0[12];  // A = Toeplitz (M = 128 output columns), B = the block's planes (N = 128 lines)
#pragma unroll
      for (int s = 0; s < 12; s++) {
        ad[s] = smem_desc_sw128(sToeA + (uint32_t)(s >> 2) * kAxBlock + (uint32_t)(s & 3) * 32u, 16, 1024);
        bd0[s] = smem_desc_sw128(sAxA + (uint32_t)(s >> 2) * kAxBlock + (uint32_t)(s & 3) * 32u, 16, 1024);
      }
      uint32_t pAx[2] = {0, 0}, pRing = 0;
      bool anyX = false;
      // Tickets (strip, row chunk) are taken from a global counter — a CTA that starts late (a collective's kernels
      // may hold its SM for a while) simply takes fewer — and published to the other roles through sTicket[k & 1] /
      // barTicket[k & 1]; the X issuer is the first role that needs the next ticket.
      for (int kt = 0;; kt++) {
        const int t = (int)atomicAdd(a.ticket, 1u);
        sTicket[kt & 1] = t < total ? t : -1;
        mbar_arrive(&barTicket[kt & 1]);
        if (t >= total) break;
        const int ord = t / a.strips, strip = t - ord * a.strips, chunk = tc_chunk_of(ord, a.chunks);
        const int x0 = strip * kTcStripW;
        const int cy0 = a.y0 + chunk * a.chunkRows, cy1 = min(a.y1, cy0 + a.chunkRows);
        const int nb = (cy1 - cy0 + kTcRows - 1) / kTcRows + 2;
        const int rowBase = cy0 - kTcRows;
        if ((a.flagTop && rowBase < a.y0) || (a.flagBottom && rowBase + kTcRows * nb > a.y1)) {
          // this ticket reads halo rows: they are complete once the neighbour's epoch flag is there
          const unsigned* f = (a.flagTop && rowBase < a.y0) ? a.flagTop : nullptr;
          const unsigned* g2 = (a.flagBottom && rowBase + kTcRows * nb > a.y1) ? a.flagBottom : nullptr;
          const long long t0_ = clock64();
          while ((f && (int)(ld_acquire_sys_u32(f) - a.epoch) < 0) || (g2 && (int)(ld_acquire_sys_u32(g2) - a.epoch) < 0)) {
            __nanosleep(100);
            if (clock64() - t0_ > 8000000000ll) __trap();
          }
          asm volatile("fence.proxy.async;" ::: "memory");  // the TMA reads below come after the flag
        }
#pragma unroll 1
        for (int i = 0; i < 2 && i < nb; i++) {  // both raw buffers are free: every earlier block has been converted
          mbar_arrive_expect_tx(&barRaw[i], kRawBytes);
          tma_load_2d(sRaw + i * kRawBytes, &tmap, &barRaw[i], x0 - 32, rowBase + kTcRows * i);
        }
#pragma unroll 1
        for (int i = 0; i < nb; i++) {
          const int b = i & 1;
          const long long w0_ = DBG ? clock64() : 0;
          mbar_wait(&barAx[b], pAx[b]);
          pAx[b] ^= 1;
          const long long w1_ = DBG ? clock64() : 0;
          if (i + 2 < nb) {
            mbar_arrive_expect_tx(&barRaw[b], kRawBytes);
            tma_load_2d(sRaw + b * kRawBytes, &tmap, &barRaw[b], x0 - 32, rowBase + kTcRows * (i + 2));
          }
          if (anyX) {  // D_x drained: the X epilogue of the previous block has finished
            mbar_wait(&barRing, pRing);
            pRing ^= 1;
          }
          anyX = true;
          const long long w2_ = DBG ? clock64() : 0;
          tc_fence_after_sync();
#pragma unroll
          for (int s = 0; s < 12; s++) mma_f16_ss(tmem + kTmDx, ad[s], bd0[s] + (uint64_t)((uint32_t)b * (kAxBytes >> 4)), idescX, s > 0 ? 1u : 0u);
          mma_commit(&barX);
          if (DBG) { atomicAdd(&a.dbg[16], (unsigned long long)(w1_ - w0_)); atomicAdd(&a.dbg[17], (unsigned long long)(w2_ - w1_)); atomicAdd(&a.dbg[18], (unsigned long long)(clock64() - w2_)); }
        }
      }
    }
  } else if (warp > kTcWorkers / 32) {
    // ================================================================ Y issuers: channels 2 h, 2 h + 1
    if (lane == 0) {
      const int h = warp - (kTcWorkers / 32 + 1);
      const uint32_t sToeA = smem_u32(sToe);
      constexpr uint32_t idescY = idesc_f16(128, 32, false, false);
      uint64_t bd[6];  // B = T[n][k], n < 32, k < 96: the first 32 rows of the X pass's Toeplitz operand
#pragma unroll
      for (int s = 0; s < 6; s++) bd[s] = smem_desc_sw128(sToeA + (uint32_t)(s >> 2) * kAxBlock + (uint32_t)(s & 3) * 32u, 16, 1024);
      uint32_t pRing = 0, pYFree = 0, pT[2] = {0, 0};
      bool anyY = false;
      for (int kt = 0;; kt++) {
        mbar_wait(&barTicket[kt & 1], pT[kt & 1]);
        pT[kt & 1] ^= 1;
        const int t = sTicket[kt & 1];
        if (t < 0) break;
        const int chunk = tc_chunk_of(t / a.strips, a.chunks);
        const int cy0 = a.y0 + chunk * a.chunkRows, cy1 = min(a.y1, cy0 + a.chunkRows);
        const int nb = (cy1 - cy0 + kTcRows - 1) / kTcRows + 2;
#pragma unroll 1
        for (int k = 0; k < nb; k++) {
          const long long w0_ = (DBG && h == 0) ? clock64() : 0;
          mbar_wait(&barRing, pRing);
          pRing ^= 1;
          const long long w1_ = (DBG && h == 0) ? clock64() : 0;
          if (DBG && h == 0) atomicAdd(&a.dbg[19], (unsigned long long)(w1_ - w0_));
          if (k < 2) continue;
          if (anyY) {  // D_y drained
            mbar_wait(&barYFree, pYFree);
            pYFree ^= 1;
          }
          anyY = true;
          const long long w2_ = (DBG && h == 0) ? clock64() : 0;
          tc_fence_after_sync();
          const uint32_t start = (uint32_t)(kTcRows * ((k - 2) & 3));  // ring blocks k - 2, k - 1, k
#pragma unroll
          for (int cc = 0; cc < 2; cc++) {
            const uint32_t c = (uint32_t)(2 * h + cc);
#pragma unroll
            for (int s = 0; s < 6; s++) {
              const uint32_t row = (start + 16u * (uint32_t)s) & (uint32_t)(kTcRingRows - 1);
              mma_f16_ts(tmem + kTmDy + c * 32u, tmem + kTmRing + c * 64u + (row >> 1), bd[s], idescY, s > 0 ? 1u : 0u);
            }
          }
          mma_commit(&barY);
          if (DBG && h == 0) { atomicAdd(&a.dbg[20], (unsigned long long)(w2_ - w1_)); atomicAdd(&a.dbg[21], (unsigned long long)(clock64() - w2_)); }
        }
      }
    }
  } else {
    // ================================================================ workers
    uint32_t pRaw[2] = {0, 0}, pX = 0, pY = 0;
    long long tm[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc0 = DBG ? clock64() : 0;
#define TC_MARK(slot) if (DBG) { const long long n_ = clock64(); tm[slot] += n_ - tc0; tc0 = n_; }
    const int q = warp & 3, wg = warp >> 2;  // TMEM lane quarter (32 of the strip's columns); channel (X) / 8 rows (Y)
    uint32_t pT[2] = {0, 0};
    for (int kt = 0;; kt++) {
      mbar_wait(&barTicket[kt & 1], pT[kt & 1]);
      pT[kt & 1] ^= 1;
      const int t = sTicket[kt & 1];
      if (t < 0) break;
      const int ord = t / a.strips, strip = t - ord * a.strips, chunk = tc_chunk_of(ord, a.chunks);
      const int x0 = strip * kTcStripW;
      const int cy0 = a.y0 + chunk * a.chunkRows, cy1 = min(a.y1, cy0 + a.chunkRows);
      const int outBlocks = (cy1 - cy0 + kTcRows - 1) / kTcRows, nb = outBlocks + 2;
      const int rowBase = cy0 - kTcRows;  // first input row of block 0
      const bool patch = a.oob != 0u && (x0 - 32 < 0 || x0 + kTcInW - 32 > a.w);

      // X epilogue of block ib: lane = output column, this warp's 32 accumulator columns = the 32 rows of channel wg;
      // quantised, two rows per word (fp16 subnormals), into the block's 16 ring columns of that channel
      auto x_epilogue = [&](int ib) {
        TC_MARK(6)
        mbar_wait(&barX, pX);
        pX ^= 1;
        TC_MARK(2)
        tc_fence_after_sync();
        uint32_t v0[16], v1[16];
        tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + kTmDx + (uint32_t)(32 * wg), v0);
        tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + kTmDx + (uint32_t)(32 * wg + 16), v1);
        tmem_ld_wait();
        uint32_t wd[16];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          wd[k] = __byte_perm(tc_quant(v0[2 * k]), tc_quant(v0[2 * k + 1]), 0x5410);
          wd[8 + k] = __byte_perm(tc_quant(v1[2 * k]), tc_quant(v1[2 * k + 1]), 0x5410);
        }
        const int grow0 = rowBase + kTcRows * ib;
        if (grow0 < 0 || grow0 + kTcRows > a.h) {  // rows outside the image are the out-of-bounds colour itself
          const uint32_t oobc = (a.oob >> (8 * wg)) & 255u;
#pragma unroll
          for (int k = 0; k < 16; k++) {
            const int g0 = grow0 + 2 * k, g1 = g0 + 1;
            if (g0 < 0 || g0 >= a.h) wd[k] = (wd[k] & 0xFFFF0000u) | oobc;
            if (g1 < 0 || g1 >= a.h) wd[k] = (wd[k] & 0x0000FFFFu) | (oobc << 16);
          }
        }
        tmem_st16(tmem + ((uint32_t)(32 * q) << 16) + kTmRing + (uint32_t)(64 * wg + 16 * (ib & 3)), wd);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&barRing);
        TC_MARK(3)
      };
      // Y epilogue of output block `ob`: lane = column, 8 rows per warp, four channels from four accumulators
      auto y_epilogue = [&](int ob) {
        TC_MARK(6)
        mbar_wait(&barY, pY);
        pY ^= 1;
        TC_MARK(4)
        tc_fence_after_sync();
        const int x = x0 + 32 * q + lane;
        const int orow0 = cy0 + kTcRows * ob + 8 * wg;
        uint32_t v[4][8];
#pragma unroll
        for (int c = 0; c < 4; c++) tmem_ld8(tmem + ((uint32_t)(32 * q) << 16) + kTmDy + (uint32_t)(c * 32 + 8 * wg), v[c]);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&barYFree);  // D_y is in registers: the next Y MMAs may overwrite it
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
        TC_MARK(5)
      };

#pragma unroll 1
      for (int i = 0; i < nb; i++) {
        // ---- raw RGBX block -> planar fp16 planes A_x (buffer i & 1)
        const int b = i & 1;
        TC_MARK(6)
        mbar_wait(&barRaw[b], pRaw[b]);
        pRaw[b] ^= 1;
        TC_MARK(0)
        const uint8_t* raw = sRaw + b * kRawBytes;
        uint8_t* ax = sAx + b * kAxBytes;
        // 768 tasks of 8 pixels (row, group g8 of the 24 per row): two 16-byte loads, byte permutes, one 16-byte store
        // per channel plane (8 halfs = one swizzle chunk)
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int qd = tid + kTcWorkers * u;
          if (qd < kTcRows * (kTcInW / 8)) {
            const int row = qd / (kTcInW / 8), g8 = qd - row * (kTcInW / 8);
            const uint4* src = reinterpret_cast<const uint4*>(raw + ((size_t)row * kTcInW + 8 * g8) * 4);
            uint4 p = src[0], p2 = src[1];
            if (patch) {  // columns outside the image carry the out-of-bounds colour (TMA filled them with zeros)
              const int x = x0 - 32 + 8 * g8;
              uint32_t* pp = &p.x;
              uint32_t* pq = &p2.x;
#pragma unroll
              for (int e = 0; e < 4; e++) {
                if (x + e < 0 || x + e >= a.w) pp[e] = a.oob;
                if (x + 4 + e < 0 || x + 4 + e >= a.w) pq[e] = a.oob;
              }
            }
            const uint32_t rg01 = __byte_perm(p.x, p.y, 0x5140), ba01 = __byte_perm(p.x, p.y, 0x7362);
            const uint32_t rg23 = __byte_perm(p.z, p.w, 0x5140), ba23 = __byte_perm(p.z, p.w, 0x7362);
            const uint32_t rg45 = __byte_perm(p2.x, p2.y, 0x5140), ba45 = __byte_perm(p2.x, p2.y, 0x7362);
            const uint32_t rg67 = __byte_perm(p2.z, p2.w, 0x5140), ba67 = __byte_perm(p2.z, p2.w, 0x7362);
            // line = channel * 32 + row; column block g8 / 8, 16-byte chunk g8 % 8
            uint8_t* d = ax + (uint32_t)(g8 >> 3) * kAxBlock + sw128_off((uint32_t)row, (uint32_t)(g8 & 7));
            // a byte next to a zero byte is the fp16 subnormal b * 2^-24
            *reinterpret_cast<uint4*>(d) = make_uint4(__byte_perm(rg01, 0u, 0x4140), __byte_perm(rg23, 0u, 0x4140),
                                                      __byte_perm(rg45, 0u, 0x4140), __byte_perm(rg67, 0u, 0x4140));
            *reinterpret_cast<uint4*>(d + 32 * 128) = make_uint4(__byte_perm(rg01, 0u, 0x4342), __byte_perm(rg23, 0u, 0x4342),
                                                                 __byte_perm(rg45, 0u, 0x4342), __byte_perm(rg67, 0u, 0x4342));
            *reinterpret_cast<uint4*>(d + 64 * 128) = make_uint4(__byte_perm(ba01, 0u, 0x4140), __byte_perm(ba23, 0u, 0x4140),
                                                                 __byte_perm(ba45, 0u, 0x4140), __byte_perm(ba67, 0u, 0x4140));
            *reinterpret_cast<uint4*>(d + 96 * 128) = make_uint4(__byte_perm(ba01, 0u, 0x4342), __byte_perm(ba23, 0u, 0x4342),
                                                                 __byte_perm(ba45, 0u, 0x4342), __byte_perm(ba67, 0u, 0x4342));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&barAx[b]);
        TC_MARK(1)
        if (i >= 1) x_epilogue(i - 1);
        if (i >= 4) y_epilogue(i - 4);  // Y of block i - 2 = output block i - 4
      }
      x_epilogue(nb - 1);
      if (nb >= 4) y_epilogue(nb - 4);  // Y of block nb - 2
      y_epilogue(nb - 3);               // Y of the last block
    }
    if (DBG && lane == 0 && (warp == 0 || warp == 9))
      for (int k = 0; k < 7; k++) atomicAdd(&a.dbg[(warp ? 8 : 0) + k], (unsigned long long)tm[k]);
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
  const size_t smem = 1024 + kToeBytes + 2 * kAxBytes + 2 * kRawBytes;
  static bool configured = false;
  if (!configured) {
    PX_CUDA(cudaFuncSetAttribute(blur_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PX_CUDA(cudaFuncSetAttribute(blur_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  if (a.dbg) blur_tc_kernel<true><<<blocks, kTcThreads, smem, st>>>(tmap, a);
  else blur_tc_kernel<false><<<blocks, kTcThreads, smem, st>>>(tmap, a);
  PX_LAUNCHED();
  return 0;
}

// Fused tensor-core blur of rows [y0, y1) of `src` (w x h RGBX) into `dst` (same geometry, a different buffer).
// -1: outside this kernel's domain (radius > 32, LUT not exact in fp32, width not a multiple of 4, no TMA entry point).
int blur_tc(const px_t* src, px_t* dst, int w, int h, const uint16_t* lut_host, int radius, uint32_t oob, int y0, int y1,
            const unsigned* flagTop, const unsigned* flagBottom, unsigned epoch) {
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
  a.flagTop = flagTop; a.flagBottom = flagBottom; a.epoch = epoch;
  a.strips = (w + kTcStripW - 1) / kTcStripW;
  // row chunks: a chunk re-blurs 64 warm-up rows horizontally, so long chunks are cheaper; but the tickets
  // (strip, chunk) are handed out dynamically to one CTA per SM and should outnumber the SMs a few times over
  const int rows = y1 - y0;
  int chunkRows = 1024;
  while (chunkRows > 128 && (long long)a.strips * ((rows + chunkRows - 1) / chunkRows) < 2ll * r.num_sms) chunkRows /= 2;
  a.chunkRows = chunkRows;
  a.chunks = (rows + chunkRows - 1) / chunkRows;
  void* tk;
  if (int rc = get_scratch(3, 512, &tk)) return rc;
  PX_CUDA(cudaMemsetAsync(tk, 0, 8, r.stream));
  a.ticket = (unsigned*)tk;
  a.dbg = nullptr;
  static const bool dbgOn = getenv("PIXIE_CUDA_TC_DEBUG") != nullptr;
  if (dbgOn) {
    void* d;
    d = (uint8_t*)tk + 64;
    PX_CUDA(cudaMemsetAsync(d, 0, 32 * 8, r.stream));
    a.dbg = (unsigned long long*)d;
  }
  const int blocks = std::max(1, std::min(a.strips * a.chunks, r.num_sms - r.sm_reserve));
  ProfScope ps(kProfBlurX);
  const int rcl = launch_tc(tmap, a, blocks, r.stream);
  if (dbgOn && rcl == 0) {  // per-phase cycles of worker warps 0 and 9 and of the issuers, averaged per 32-row block
    unsigned long long hd[32];
    PX_CUDA(cudaMemcpyAsync(hd, a.dbg, sizeof(hd), cudaMemcpyDeviceToHost, r.stream));
    PX_CUDA(cudaStreamSynchronize(r.stream));
    const double nbk = (double)a.strips * a.chunks * ((double)a.chunkRows / kTcRows + 2);
    const char* nm[7] = {"wait raw", "convert", "wait X", "x epi", "wait Y", "y epi", "other"};
    fprintf(stderr, "[blur_tc] cycles per block:");
    for (int k = 0; k < 7; k++) fprintf(stderr, " w0 %s %.0f |", nm[k], hd[k] / nbk);
    for (int k = 0; k < 7; k++) fprintf(stderr, " w9 %s %.0f |", nm[k], hd[8 + k] / nbk);
    fprintf(stderr, " X issuer: wait planes %.0f, wait D_x %.0f, issue %.0f | Y issuer: wait ring %.0f, wait D_y %.0f, issue %.0f\n", hd[16] / nbk,
            hd[17] / nbk, hd[18] / nbk, hd[19] / nbk, hd[20] / nbk, hd[21] / nbk);
  }
  return rcl;
}

}  // namespace pixie

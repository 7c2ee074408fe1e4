// Standalone probe of the tcgen05 / TMEM building blocks of blur_tc.cu (run on the B200 box through gpurun):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pixie_b200/csrc/cuda -o /tmp/umma_probe tools/umma_probe.cu
// Checks, against exact integer arithmetic on the host:
//   T1  D = A * B^T with A K-major (128 lines x 128 K) and B K-major (64 x 128), 128-byte swizzle, fp16 operands where
//       A holds pixel bytes as fp16 SUBNORMALS (the bit pattern 0x00bb) and B integer taps < 2048;
//   T2  the same product with A MN-major (rows = K, 128 bytes = 64 lines per row, two 64-line blocks), i.e. the
//       layout of the vertical pass's ring of rows, addressed at a row offset;
//   and prints the cycles a batch of MMAs of the X-pass / Y-pass shapes takes.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "umma.cuh"

using namespace pixie::umma;

constexpr int M = 128, N = 64, K = 128;

struct Out {
  float d[M * N];
  long long cycles[8];
};

// mode 0: A K-major; mode 1: A MN-major with the K window starting at ring row `krow0`
__global__ void __launch_bounds__(128) probe(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, Out* out, int mode,
                                             int krow0, int ringRows, int nOut, int kSteps, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                        // mode 0: 2 blocks of [128 lines][128 B]; mode 1: 2 blocks of [ringRows][128 B]
  uint8_t* sB = base + 2 * 128 * 128 * 2;    // generous: 64 KB for A
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nB = nOut > 64 ? 256 : 64;  // rows of a B block (timing-only cases read garbage, inside the buffer)
  const uint32_t blkA = mode == 0 ? 128u * 128u : (uint32_t)ringRows * 128u;
  // ---- stage A and B into the swizzled layouts
  if (mode == 0) {
    for (int i = tid; i < M * K; i += 128) {
      const int m = i / K, k = i % K;
      const uint32_t off = (k / 64) * blkA + sw128_off(m, (k % 64) / 8) + (k % 8) * 2;
      *reinterpret_cast<uint16_t*>(sA + off) = A[m * K + k];
    }
  } else {
    for (int i = tid; i < M * ringRows; i += 128) {
      const int m = i % M, row = i / M;  // ring row holds logical K index (row - krow0 + ringRows) % ringRows
      const int k = (row - krow0 + ringRows) % ringRows;
      const uint16_t v = k < K ? A[m * K + k] : (uint16_t)0x3C00;  // rows outside the window hold 1.0 (must not be read)
      const uint32_t off = (m / 64) * blkA + sw128_off(row, (m % 64) / 8) + (m % 8) * 2;
      *reinterpret_cast<uint16_t*>(sA + off) = v;
    }
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const uint32_t off = (k / 64) * (64u * 128u) + sw128_off(n, (k % 64) / 8) + (k % 8) * 2;
    *reinterpret_cast<uint16_t*>(sB + off) = B[n * K + k];
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (mode == 2) {  // A in tensor memory: lane = line, column 128 + k / 2 holds elements k, k + 1 (low, high half)
    const int m = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t r[8];
      for (int j = 0; j < 8; j++) r[j] = (uint32_t)A[m * K + 2 * (c0 + j)] | ((uint32_t)A[m * K + 2 * (c0 + j) + 1] << 16);
      tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 128u + (uint32_t)c0, r);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }
  uint32_t phase = 0;
  long long t0 = 0, t1 = 0;
  for (int rep = 0; rep < reps; rep++) {
    if (tid == 0) {
      const uint32_t idesc = idesc_f16(128, nOut, mode == 1, false);
      // descriptors first: the timed loop below only issues (a lone thread that also divides by run-time values
      // issues one MMA every ~200 cycles and hides what the tensor pipe can do)
      uint64_t ad[8], bd[8];
#pragma unroll
      for (int s = 0; s < 8; s++) {
        if (mode == 0) {
          ad[s] = smem_desc_sw128(smem_u32(sA) + (s / 4) * blkA + (s % 4) * 32, 16, 1024);
        } else {
          const int row = (krow0 + 16 * s) % ringRows;
          ad[s] = smem_desc_sw128(smem_u32(sA) + row * 128, blkA, 1024);
        }
        bd[s] = smem_desc_sw128(smem_u32(sB) + (s / 4) * (nB * 128) + (s % 4) * 32, 16, 1024);
      }
      const int rounds = rep >= 4 ? (1 << (rep - 3)) : 1;  // timing reps: 2x, 4x, ... the MMAs per commit
      t0 = clock64();
      for (int q = 0; q < rounds; q++) {
        if (mode == 2) {
          if (kSteps == 8) {
#pragma unroll
            for (int s = 0; s < 8; s++) mma_f16_ts(tmem, tmem + 128u + 8u * s, bd[s], idesc, (s > 0 || rep >= 4) ? 1u : 0u);
          } else {
#pragma unroll
            for (int s = 0; s < 6; s++) mma_f16_ts(tmem, tmem + 128u + 8u * s, bd[s], idesc, (s > 0 || rep >= 4) ? 1u : 0u);
          }
        } else if (kSteps == 8) {
#pragma unroll
          for (int s = 0; s < 8; s++) mma_f16_ss(tmem, ad[s], bd[s], idesc, (s > 0 || rep >= 4) ? 1u : 0u);
        } else {
#pragma unroll
          for (int s = 0; s < 6; s++) mma_f16_ss(tmem, ad[s], bd[s], idesc, (s > 0 || rep >= 4) ? 1u : 0u);
        }
      }
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    if (tid == 0) {
      t1 = clock64();
      if (rep < 8) out->cycles[rep] = t1 - t0;
    }
    if (rep == 3) {
  tc_fence_after_sync();
  // ---- read D: warp w reads lanes 32w .. 32w + 31
  for (int c0 = 0; c0 < (nOut > 64 ? 0 : nOut); c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; j++) out->d[(warp * 32 + (tid & 31)) * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Issue-rate probe: `nIssuers` warps (lane 0 of each) issue `count` MMAs each (N = nOut, A K-major) into their own
// accumulator columns, then commit to one barrier (count = nIssuers).  cycles = first issue .. all complete.
__global__ void __launch_bounds__(128) issue_probe(Out* out, int nIssuers, int nOut, int count) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32 * 1024; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0x00010001u;
  if (tid == 0) {
    mbar_init(&bar, nIssuers);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const long long t0 = clock64();
  if (warp < nIssuers && (tid & 31) == 0) {
    const uint32_t idesc = idesc_f16(128, nOut, false, false);
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
      ad[s] = smem_desc_sw128(smem_u32(base) + s * 32, 16, 1024);
      bd[s] = smem_desc_sw128(smem_u32(base) + 64 * 1024 + s * 32, 16, 1024);
    }
    for (int q = 0; q < count / 4; q++) {
#pragma unroll
      for (int s = 0; s < 4; s++) mma_f16_ss(tmem + warp * 64, ad[s], bd[s], idesc, 1u);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) out->cycles[0] = clock64() - t0;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main(int argc, char** argv) {
  std::vector<uint16_t> A(M * K), B(N * K);
  srand(7);
  const bool normal = argc > 1 && !strcmp(argv[1], "normal");  // timing A/B: bytes as NORMAL halfs 1024 + b (0x6400 | b)
  for (auto& v : A) v = (uint16_t)((rand() & 255) | (normal ? 0x6400 : 0));  // pixel byte as an fp16 subnormal: value = byte * 2^-24
  for (int n = 0; n < N; n++)
    for (int k = 0; k < K; k++) {
      const int t = k - n;  // Toeplitz band of 65 taps, integers < 2048 exact in fp16
      int tap = (t >= 0 && t <= 64) ? (rand() % 1900) : 0;
      // exact integer -> fp16 bits by hand
      int e = 0, m = tap;
      uint16_t bits = 0;
      if (tap) {
        e = 31 - __builtin_clz(tap);
        m = tap << (10 - e);  // 1.xxx with 10 fraction bits (tap < 2048)
        bits = (uint16_t)(((e + 15) << 10) | (m & 0x3FF));
      }
      B[n * K + k] = bits;
    }
  auto half_val = [](uint16_t b) -> double {
    const int e = (b >> 10) & 31, m = b & 0x3FF;
    if (e == 0) return m * (1.0 / 16777216.0);
    return (double)(1024 + m) * (1.0 / 1024.0) * (double)(1ull << e) / 32768.0;
  };
  uint16_t *dA, *dB;
  Out* dOut;
  cudaMalloc(&dA, A.size() * 2);
  cudaMalloc(&dB, B.size() * 2);
  cudaMalloc(&dOut, sizeof(Out));
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  const size_t smemBytes = 1024 + 64 * 1024 + 64 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  Out* h = new Out;
  int fails = 0;
  struct Case { int mode, krow0, ringRows, nOut, kSteps; const char* name; };
  const Case cases[] = {
      {0, 0, 0, 64, 8, "T1 K-major A, N=64, K=128 (X pass shape)"},
      {0, 0, 0, 32, 6, "T1b K-major A, N=32, K=96"},
      {1, 0, 128, 64, 8, "T2 MN-major A, ring offset 0, N=64, K=128"},
      {1, 48, 96, 32, 6, "T2b MN-major A, ring of 96 rows, window from row 48 (wraps), N=32, K=96 (Y pass shape)"},
      {1, 80, 96, 32, 6, "T2c MN-major A, ring of 96 rows, window from row 80"},
      {2, 0, 0, 64, 8, "T3 A in TMEM (TS), N=64, K=128"},
      {2, 0, 0, 32, 6, "T3b A in TMEM (TS), N=32, K=96 (Y pass shape)"},
      {2, 0, 0, 16, 6, "T3c A in TMEM (TS), N=16, K=96"},
      {0, 0, 0, 16, 8, "timing only: N=16"},
      {0, 0, 0, 128, 8, "timing only: N=128"},
      {0, 0, 0, 256, 8, "timing only: N=256"},
      {1, 0, 128, 256, 8, "timing only: MN-major A, N=256"},
  };
  for (const Case& c : cases) {
    cudaMemset(dOut, 0, sizeof(Out));
    probe<<<1, 128, smemBytes>>>(dA, dB, dOut, c.mode, c.krow0, c.ringRows, c.nOut, c.kSteps, 8);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e));
      return 2;
    }
    cudaMemcpy(h, dOut, sizeof(Out), cudaMemcpyDeviceToHost);
    int bad = 0;
    double worst = 0;
    const int Kc = c.kSteps * 16;
    for (int m = 0; m < (c.nOut <= 64 ? M : 0); m++)
      for (int n = 0; n < c.nOut; n++) {
        double want = 0;
        for (int k = 0; k < Kc; k++) want += half_val(A[m * K + k]) * half_val(B[n * K + k]);
        const double got = h->d[m * N + n];
        if (got != want) {
          if (bad < 5) printf("  mismatch m=%d n=%d got %.10g want %.10g (x2^24: %.1f vs %.1f)\n", m, n, got, want, got * 16777216.0, want * 16777216.0);
          bad++;
          worst = fmax(worst, fabs(got - want));
        }
      }
    printf("%s: %s (%d of %d differ) cycles per batch of %d MMAs: %lld %lld; x2 %lld x4 %lld x8 %lld x16 %lld -> %.1f cycles per MMA\n", c.name,
           bad ? "FAIL" : "PASS", bad, M * c.nOut, c.kSteps, h->cycles[2], h->cycles[3], h->cycles[4], h->cycles[5], h->cycles[6], h->cycles[7],
           (double)(h->cycles[7] - h->cycles[6]) / (8.0 * c.kSteps));
    fails += bad != 0;
  }
  cudaFuncSetAttribute(issue_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  for (int nOut : {32, 64}) {
    for (int nIss : {1, 2, 4}) {
      for (int count : {64, 256}) {
        issue_probe<<<1, 128, smemBytes>>>(dOut, nIss, nOut, count);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("issue_probe: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
        cudaMemcpy(h, dOut, sizeof(Out), cudaMemcpyDeviceToHost);
        printf("issue probe N=%d: %d issuer warp(s) x %d MMAs: %lld cycles -> %.1f cycles per MMA overall\n", nOut, nIss, count, h->cycles[0],
               (double)h->cycles[0] / (nIss * count));
      }
    }
  }
  printf(fails ? "PROBE FAILED\n" : "PROBE OK\n");
  return fails ? 1 : 0;
}

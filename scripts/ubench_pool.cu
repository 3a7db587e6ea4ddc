// ubench_pool.cu — build the phase-1 consumer up step by step to see what limits one SM.
//   LEVEL 0: producer (TMA bulk, 32 KB tiles, D stages) + consumers that only wait and release
//   LEVEL 1: + LDS of the thread's 8 float4
//   LEVEL 2: + partial dots, butterfly reduce, smem exchange, named barrier, second-level reduce
//   LEVEL 3: + the weighted accumulation (the full phase-1 inner loop)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench_pool.bin scripts/ubench_pool.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../attentionalpoolingaction_b200/csrc/apa_common.cuh"
using namespace apa;

constexpr int C = 2048, ROWS = 4, TILE = ROWS * C * 4;

template <int LEVEL, int D, int MINB>
__global__ void __launch_bounds__(288, MINB)
k_pool(const float* X, int rows_per_cta, const float* w, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[8], empty_bar[8];
  __shared__ float partS[2][ROWS * 8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = rows_per_cta / ROWS;
  const float* base = X + size_t(blockIdx.x) * rows_per_cta * C;
  if (tid == 0) {
    for (int i = 0; i < D; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 8); }
    fence_mbar_init();
  }
  __syncthreads();
  long long t0 = clock64();
  float keep = 0.f;
  if (warp == 8) {
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full_bar[stage], TILE);
        bulk_g2s(smem + stage * TILE, base + size_t(t) * ROWS * C, TILE, &full_bar[stage], pol);
        if (++stage == D) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    const int col0 = warp * 256 + lane * 4;
    float4 wv[2];
    wv[0] = *reinterpret_cast<const float4*>(w + col0);
    wv[1] = *reinterpret_cast<const float4*>(w + col0 + 128);
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float s = 0.f;
    uint32_t stage = 0, phase = 0;
    for (int t = 0; t < ntiles; ++t) {
      const float* st = reinterpret_cast<const float*>(smem + stage * TILE) + col0;
      mbar_wait(&full_bar[stage], phase);
      float4 x[ROWS][2];
      if (LEVEL >= 1) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          x[r][0] = *reinterpret_cast<const float4*>(st + r * C);
          x[r][1] = *reinterpret_cast<const float4*>(st + r * C + 128);
        }
      }
      float a[ROWS] = {1.f, 1.f, 1.f, 1.f};
      if (LEVEL >= 2) {
        float pd[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          pd[r] = fmaf(x[r][0].x, wv[0].x, fmaf(x[r][0].y, wv[0].y, fmaf(x[r][0].z, wv[0].z, x[r][0].w * wv[0].w)));
          pd[r] += fmaf(x[r][1].x, wv[1].x, fmaf(x[r][1].y, wv[1].y, fmaf(x[r][1].z, wv[1].z, x[r][1].w * wv[1].w)));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
        float k0 = hi16 ? pd[2] : pd[0], s0 = hi16 ? pd[0] : pd[2];
        float k1 = hi16 ? pd[3] : pd[1], s1 = hi16 ? pd[1] : pd[3];
        k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
        k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        float kk = hi8 ? k1 : k0;
        const float ss = hi8 ? k0 : k1;
        kk += __shfl_xor_sync(0xffffffffu, ss, 8);
        kk += __shfl_xor_sync(0xffffffffu, kk, 4);
        kk += __shfl_xor_sync(0xffffffffu, kk, 2);
        kk += __shfl_xor_sync(0xffffffffu, kk, 1);
        float* part = partS[t & 1];
        if ((lane & 7) == 0) part[(lane >> 3) * 8 + warp] = kk;
        named_bar_sync(1, 256);
        float v = part[lane];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) a[r] = __shfl_sync(0xffffffffu, v, 8 * r);
      } else {
        if (LEVEL >= 1) keep += x[0][0].x + x[1][0].y + x[2][1].z + x[3][1].w;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
      }
      if (LEVEL >= 3) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          s += a[r];
          f[0] = fmaf(a[r], x[r][0].x, f[0]); f[1] = fmaf(a[r], x[r][0].y, f[1]);
          f[2] = fmaf(a[r], x[r][0].z, f[2]); f[3] = fmaf(a[r], x[r][0].w, f[3]);
          f[4] = fmaf(a[r], x[r][1].x, f[4]); f[5] = fmaf(a[r], x[r][1].y, f[5]);
          f[6] = fmaf(a[r], x[r][1].z, f[6]); f[7] = fmaf(a[r], x[r][1].w, f[7]);
        }
      } else if (LEVEL == 2) {
        keep += a[0] + a[1] + a[2] + a[3];
      }
      if (++stage == D) { stage = 0; phase ^= 1u; }
    }
    keep += s + f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7];
  }
  long long t1 = clock64();
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (keep == 123.456f) sink[0] = keep;
}

template <int LEVEL, int D, int MINB>
void run(const float* X, int rows_per_cta, const float* w, long long* d_cycles, float* sink, int grid, const char* what) {
  auto fn = k_pool<LEVEL, D, MINB>;
  const size_t smem = size_t(D) * TILE;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  fn<<<grid, 288, smem>>>(X, rows_per_cta, w, d_cycles, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(sizeof(long long) * grid);
  cudaMemcpy(h, d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
  const double bytes = double(rows_per_cta) * C * 4;
  printf("level %d D=%d minb=%d grid %3d: %6.1f B/clk/CTA (%5.1f GB/s) | %5.0f clk per 4-row tile | total %5.2f TB/s  %s  [%s]\n",
         LEVEL, D, MINB, grid, bytes / mean, bytes / mean * 1.965, mean / (rows_per_cta / ROWS),
         bytes * grid / (ms * 1e-3) / 1e12, what, cudaGetErrorString(cudaGetLastError()));
  free(h);
}

int main() {
  const int rows_per_cta = 384;  // 3 MB per CTA
  const int max_grid = 296;
  const size_t total = size_t(rows_per_cta) * C * 4 * max_grid;
  float* bufs[2];
  for (int i = 0; i < 2; ++i) { cudaMalloc(&bufs[i], total); cudaMemset(bufs[i], 0, total); }
  float* w; cudaMalloc(&w, C * 4); cudaMemset(w, 0, C * 4);
  long long* d_cycles; cudaMalloc(&d_cycles, sizeof(long long) * max_grid);
  float* sink; cudaMalloc(&sink, 4);
  int it = 0;
#define RUN(L, D, MB, G, WHAT) run<L, D, MB>(bufs[it++ & 1], rows_per_cta, w, d_cycles, sink, G, WHAT)
  for (int g : {32, 128, 148}) {
    RUN(0, 3, 1, g, "wait+release only");
    RUN(1, 3, 1, g, "+LDS");
    RUN(2, 3, 1, g, "+dot/reduce/barrier");
    RUN(3, 3, 1, g, "+accumulate (full)");
    RUN(3, 6, 1, g, "full, 6 stages");
    RUN(3, 2, 1, g, "full, 2 stages");
  }
  RUN(3, 3, 2, 256, "full, 2 CTAs/SM");
  RUN(3, 3, 2, 296, "full, 2 CTAs/SM");
  RUN(0, 3, 2, 296, "wait+release only, 2 CTAs/SM");
  cudaError_t err = cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}

// ubench_copy.cu — how fast can ONE SM pull cold HBM data into shared memory?
//   mode 0: cp.async.bulk (TMA engine, UBLKCP): S bytes per op, `lanes` ops per warp instruction, D tiles in flight
//   mode 1: cp.async 16 B (LDGSTS) by 256 threads, tile = T bytes, D tiles in flight
//   mode 2: plain LDG.128 into registers, U loads in flight per thread
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_copy scripts/ubench_copy.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../attentionalpoolingaction_b200/csrc/apa_common.cuh"
using namespace apa;

__global__ void __launch_bounds__(288, 1)
k_bulk(const uint8_t* src, size_t bytes_per_cta, int S, int lanes, int D, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t tile = size_t(S) * lanes;
  const int ntiles = int(bytes_per_cta / tile);
  const uint8_t* base = src + size_t(blockIdx.x) * bytes_per_cta;
  if (tid == 0) {
    for (int i = 0; i < D; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  float acc = 0.f;
  long long t0 = clock64();
  if (warp == 0) {
    const uint64_t pol = l2_policy_evict_first();
    auto issue = [&](int t) {
      const int st = t % D;
      if (lane == 0) mbar_arrive_expect_tx(&full[st], uint32_t(tile));
      __syncwarp();
      if (lane < lanes)
        bulk_g2s(smem + size_t(st) * tile + size_t(lane) * S, base + size_t(t) * tile + size_t(lane) * S, S, &full[st],
                 pol);
    };
    for (int t = 0; t < D && t < ntiles; ++t) issue(t);
    for (int t = 0; t < ntiles; ++t) {
      const int st = t % D;
      mbar_wait(&full[st], (t / D) & 1);
      acc += reinterpret_cast<const float*>(smem + size_t(st) * tile)[lane];
      __syncwarp();
      if (t + D < ntiles) issue(t + D);
    }
  }
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

__global__ void __launch_bounds__(256, 1)
k_cpasync(const uint8_t* src, size_t bytes_per_cta, int T, int D, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const int ntiles = int(bytes_per_cta / T);
  const uint8_t* base = src + size_t(blockIdx.x) * bytes_per_cta;
  const int per_thread = T / (256 * 16);
  float acc = 0.f;
  long long t0 = clock64();
  auto issue = [&](int t) {
    if (t < ntiles) {
      const uint8_t* s = base + size_t(t) * T;
      uint8_t* d = smem + size_t(t % D) * T;
      for (int i = 0; i < per_thread; ++i) {
        const size_t off = (size_t(i) * 256 + tid) * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d + off)), "l"(s + off) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int t = 0; t < D - 1; ++t) issue(t);
  for (int t = 0; t < ntiles; ++t) {
    issue(t + D - 1);
    // all but the newest D-1 groups complete -> tile t landed (own pieces)
    switch (D - 1) {
      case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
      case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
      case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
      case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
      case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
      case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
      default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
    acc += reinterpret_cast<const float*>(smem + size_t(t % D) * T)[tid * 4];
  }
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

template <int U>
__global__ void __launch_bounds__(256, 1)
k_ldg(const uint8_t* src, size_t bytes_per_cta, long long* cycles, float* sink) {
  const int tid = threadIdx.x;
  const float4* base = reinterpret_cast<const float4*>(src + size_t(blockIdx.x) * bytes_per_cta);
  const int n = int(bytes_per_cta / 16 / 256 / U);
  float acc = 0.f;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcs(base + (size_t(i) * U + u) * 256 + tid);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].w;
  }
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

static double run_stats(long long* d_cycles, int grid, size_t bytes_per_cta, float ms, const char* tag) {
  long long* h = (long long*)malloc(sizeof(long long) * grid);
  cudaMemcpy(h, d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mean = 0, mx = 0;
  for (int i = 0; i < grid; ++i) { mean += h[i]; if (h[i] > mx) mx = h[i]; }
  mean /= grid;
  const double bpc = bytes_per_cta / mean;  // bytes per clock per CTA
  printf("%-44s grid %3d  %6.1f B/clk/SM  (%5.1f GB/s/SM @1.965GHz)  total %6.2f TB/s  kernel %.1f us\n", tag, grid, bpc,
         bpc * 1.965, bytes_per_cta * (double)grid / (ms * 1e-3) / 1e12, ms * 1e3);
  free(h);
  return bpc;
}

int main() {
  const size_t bytes_per_cta = 3u << 20;  // 3 MB per CTA
  const int max_grid = 148;
  const size_t total = bytes_per_cta * max_grid;
  const int nbuf = 3;  // rotate buffers so every run is cold (3 x 444 MB > L2)
  uint8_t* bufs[nbuf];
  for (int i = 0; i < nbuf; ++i) { cudaMalloc(&bufs[i], total); cudaMemset(bufs[i], 1, total); }
  long long* d_cycles; cudaMalloc(&d_cycles, sizeof(long long) * max_grid);
  float* sink; cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int it = 0;
  char tag[128];
  const int grids[2] = {32, 148};
  for (int gi = 0; gi < 2; ++gi) {
    const int grid = grids[gi];
    const int S_list[] = {2048, 8192, 16384, 32768};
    for (int si = 0; si < 4; ++si)
      for (int lanes = 1; lanes <= 16; lanes *= 4)
        for (int D = 1; D <= 6; D += (D == 1 ? 1 : (D == 2 ? 2 : 2))) {
          const int S = S_list[si];
          const size_t smem = size_t(S) * lanes * D;
          if (smem > 192 * 1024) continue;
          cudaEventRecord(e0);
          k_bulk<<<grid, 288, smem>>>(bufs[it++ % nbuf], bytes_per_cta, S, lanes, D, d_cycles, sink);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          snprintf(tag, sizeof tag, "bulk  S=%5d lanes=%2d D=%d (in flight %3zu KB)", S, lanes, D, smem / 1024);
          run_stats(d_cycles, grid, bytes_per_cta, ms, tag);
        }
    for (int T = 16384; T <= 65536; T *= 2)
      for (int D = 2; D <= 6; D += 2) {
        const size_t smem = size_t(T) * D;
        if (smem > 192 * 1024) continue;
        cudaEventRecord(e0);
        k_cpasync<<<grid, 256, smem>>>(bufs[it++ % nbuf], bytes_per_cta, T, D, d_cycles, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        snprintf(tag, sizeof tag, "cp.async tile=%5d D=%d (in flight %3zu KB)", T, D, size_t(T) * (D - 1) / 1024);
        run_stats(d_cycles, grid, bytes_per_cta, ms, tag);
      }
    {
      float ms;
      cudaEventRecord(e0); k_ldg<4><<<grid, 256>>>(bufs[it++ % nbuf], bytes_per_cta, d_cycles, sink); cudaEventRecord(e1);
      cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); run_stats(d_cycles, grid, bytes_per_cta, ms, "ldg U=4  (16 KB in flight)");
      cudaEventRecord(e0); k_ldg<8><<<grid, 256>>>(bufs[it++ % nbuf], bytes_per_cta, d_cycles, sink); cudaEventRecord(e1);
      cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); run_stats(d_cycles, grid, bytes_per_cta, ms, "ldg U=8  (32 KB in flight)");
      cudaEventRecord(e0); k_ldg<16><<<grid, 256>>>(bufs[it++ % nbuf], bytes_per_cta, d_cycles, sink); cudaEventRecord(e1);
      cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); run_stats(d_cycles, grid, bytes_per_cta, ms, "ldg U=16 (64 KB in flight)");
    }
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}

// ubench_stream.cu — round-2 experiment bed for the STREAM side of the head (no tail, no exchange).
// Question: how fast can SHORT streams (one launch = 32 images x 196 rows x 8 KB, i.e. only
// 12..66 tiles per CTA) be pulled through an SM when launches overlap, as a function of
//   TR    rows per tile                      D     ring stages            MINB  CTAs per SM (launch bound)
//   PAIR  1: one tile per consumer chain      2: TWO tiles per chain (one butterfly/barrier round for 2 TR rows)
//   map   0: CTA = (image, row slice)         1: balanced — the launch's rows are cut into gridDim equal runs that
//                                                cross image boundaries (every SM streams the same byte count)
// and of the launch pattern (L back-to-back launches on S streams, with / without programmatic dependent launch).
// Every CTA stores its partial g at the end (8 KB) like the real STREAM kernel.  Prints aggregate TB/s over the
// whole train of launches (CUDA events), the per-CTA tile cadence, and how many SMs held 0/1/2/3+ CTAs of a launch.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench_stream.bin scripts/ubench_stream.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../attentionalpoolingaction_b200/csrc/apa_common.cuh"
using namespace apa;

constexpr int C = 2048, HW = 196, B = 32;
constexpr int kThreads = 288;

struct Args {
  const float* X;      // [B * HW, C]
  const float* w;      // [C]
  float* gpart;        // [gridDim][C]
  unsigned int* smid;  // [gridDim]
  long long* cyc;      // [gridDim][2]: stream cycles, tiles
  int lp;              // map 0: row slices per image
  int balanced;        // map 1
  int pdl_wait;        // 1: griddepcontrol.wait before the final store
};

// warp-transposing reduction of 8 per-lane values: afterwards lane (4 r) holds the warp sum of value r
__device__ __forceinline__ float butterfly8(const float (&pd)[8], int lane) {
  const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
  float k[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = h16 ? pd[i + 4] : pd[i], send = h16 ? pd[i] : pd[i + 4];
    k[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float m[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = h8 ? k[i + 2] : k[i], send = h8 ? k[i] : k[i + 2];
    m[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float keep = h4 ? m[1] : m[0], send = h4 ? m[0] : m[1];
  float v = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

template <int TR, int D, int MINB, int PAIR>
__global__ void __launch_bounds__(kThreads, MINB) k_stream(const Args a) {
  static_assert(TR * PAIR <= 8, "one chain handles at most 8 rows");
  constexpr int NR = TR * PAIR;                    // rows per consumer chain
  constexpr size_t kTile = size_t(TR) * C * 4;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[8], empty_bar[8];
  __shared__ float partS[2][64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // this CTA's run of rows [row_lo, row_hi) of the flattened [B * HW] row space
  int row_lo, row_hi;
  if (a.balanced) {
    const long long total = (long long)B * HW;
    row_lo = int(total * blockIdx.x / gridDim.x);
    row_hi = int(total * (blockIdx.x + 1) / gridDim.x);
  } else {
    const int b = blockIdx.x / a.lp, q = blockIdx.x % a.lp;
    row_lo = b * HW + (HW * q) / a.lp;
    row_hi = b * HW + (HW * (q + 1)) / a.lp;
  }
  const int nrows = row_hi - row_lo, ntiles = (nrows + TR - 1) / TR;
  if (tid == 0) {
    for (int i = 0; i < D; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 8); }
    fence_mbar_init();
    unsigned int s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    a.smid[blockIdx.x] = s;
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 8) {
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int rows = min(TR, nrows - t * TR);
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full_bar[stage], uint32_t(rows) * C * 4u);
        bulk_g2s(smem + stage * kTile, a.X + (size_t(row_lo) + size_t(t) * TR) * C, uint32_t(rows) * C * 4u,
                 &full_bar[stage], pol);
        if (++stage == D) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
    griddep_launch_dependents();
    return;
  }
  griddep_launch_dependents();  // trigger at once: the next launch may co-reside as soon as slots allow
  const int col0 = warp * 256 + lane * 4;
  float4 wv[2];
  wv[0] = __ldg(reinterpret_cast<const float4*>(a.w + col0));
  wv[1] = __ldg(reinterpret_cast<const float4*>(a.w + col0 + 128));
  float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t stage = 0, phase = 0, iter = 0;
  for (int t = 0; t < ntiles; t += PAIR, ++iter) {
    float4 x[NR][2];
    // wait for PAIR consecutive stages and pull the rows into registers
#pragma unroll
    for (int pi = 0; pi < PAIR; ++pi) {
      uint32_t st_i = stage + pi, ph_i = phase;
      if (st_i >= D) { st_i -= D; ph_i ^= 1u; }
      const bool have = (t + pi < ntiles);
      const int rows = have ? min(TR, nrows - (t + pi) * TR) : 0;
      if (have) mbar_wait(&full_bar[st_i], ph_i);
      const float* st = reinterpret_cast<const float*>(smem + st_i * kTile) + col0;
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        const bool ok = r < rows;
        x[pi * TR + r][0] = ok ? *reinterpret_cast<const float4*>(st + r * C) : make_float4(0.f, 0.f, 0.f, 0.f);
        x[pi * TR + r][1] = ok ? *reinterpret_cast<const float4*>(st + r * C + 128) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float pd[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      pd[r] = 0.f;
      if (r < NR) {
        pd[r] = fmaf(x[r][0].x, wv[0].x, fmaf(x[r][0].y, wv[0].y, fmaf(x[r][0].z, wv[0].z, x[r][0].w * wv[0].w)));
        pd[r] += fmaf(x[r][1].x, wv[1].x, fmaf(x[r][1].y, wv[1].y, fmaf(x[r][1].z, wv[1].z, x[r][1].w * wv[1].w)));
      }
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int pi = 0; pi < PAIR; ++pi)
        if (t + pi < ntiles) {
          uint32_t st_i = stage + pi;
          if (st_i >= D) st_i -= D;
          mbar_arrive(&empty_bar[st_i]);
        }
    }
    const float kk = butterfly8(pd, lane);                 // lane 4 r: warp sum of row r
    float* part = partS[iter & 1];
    if ((lane & 3) == 0) part[(lane >> 2) * 8 + warp] = kk;   // [row][warp]
    named_bar_sync(1, 256);
    float v0 = part[lane], v1 = part[32 + lane];            // rows 0..3 / 4..7, 8 warps each
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const float ar = __shfl_sync(0xffffffffu, (r < 4) ? v0 : v1, 8 * (r & 3));
      f[0] = fmaf(ar, x[r][0].x, f[0]); f[1] = fmaf(ar, x[r][0].y, f[1]);
      f[2] = fmaf(ar, x[r][0].z, f[2]); f[3] = fmaf(ar, x[r][0].w, f[3]);
      f[4] = fmaf(ar, x[r][1].x, f[4]); f[5] = fmaf(ar, x[r][1].y, f[5]);
      f[6] = fmaf(ar, x[r][1].z, f[6]); f[7] = fmaf(ar, x[r][1].w, f[7]);
    }
    stage += PAIR;
    if (stage >= D) { stage -= D; phase ^= 1u; }
  }
  const long long t1 = clock64();
  if (a.pdl_wait) griddep_wait();
  float* dst = a.gpart + size_t(blockIdx.x) * C + col0;
  *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(dst + 128) = make_float4(f[4], f[5], f[6], f[7]);
  if (tid == 0) {
    a.cyc[2 * blockIdx.x + 0] = t1 - t0;
    a.cyc[2 * blockIdx.x + 1] = ntiles;
  }
}

struct Bufs {
  std::vector<float*> X;  // ring of inputs larger than 2x L2
  float* w;
  float* gpart;
  unsigned int* smid;
  long long* cyc;
};

template <int TR, int D, int MINB, int PAIR>
void run(const Bufs& bufs, int grid, int lp, int balanced, int nstreams, int pdl, int launches, const char* what) {
  static_assert(D % PAIR == 0, "a pair of tiles must not wrap the ring in the middle");
  auto fn = k_stream<TR, D, MINB, PAIR>;
  const size_t smem = size_t(D) * TR * C * 4;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  std::vector<cudaStream_t> st(nstreams);
  for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<cudaEvent_t> done(nstreams);
  for (auto& e : done) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  auto enqueue = [&](int n) {
    for (int i = 0; i < n; ++i) {
      Args a{bufs.X[i % bufs.X.size()], bufs.w, bufs.gpart + size_t(i % nstreams) * 512 * C, bufs.smid, bufs.cyc,
             lp, balanced, pdl};
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st[i % nstreams];
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
      cudaLaunchKernelEx(&cfg, fn, a);
    }
  };
  enqueue(2 * nstreams);  // warm-up
  cudaDeviceSynchronize();
  cudaEventRecord(e0, st[0]);
  for (int s = 1; s < nstreams; ++s) cudaStreamWaitEvent(st[s], e0, 0);
  enqueue(launches);
  for (int s = 1; s < nstreams; ++s) { cudaEventRecord(done[s], st[s]); cudaStreamWaitEvent(st[0], done[s], 0); }
  cudaEventRecord(e1, st[0]);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(2 * grid);
  std::vector<unsigned int> hs(grid);
  cudaMemcpy(h.data(), bufs.cyc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
  cudaMemcpy(hs.data(), bufs.smid, sizeof(unsigned int) * grid, cudaMemcpyDeviceToHost);
  double cyc = 0, tiles = 0;
  for (int i = 0; i < grid; ++i) { cyc += double(h[2 * i]); tiles += double(h[2 * i + 1]); }
  int cnt[256] = {0}, hist[4] = {0, 0, 0, 0};
  for (int i = 0; i < grid; ++i) cnt[hs[i] & 255]++;
  for (int i = 0; i < 148; ++i) hist[cnt[i] > 3 ? 3 : cnt[i]]++;
  const double bytes = double(B) * HW * C * 4;
  printf("TR=%d D=%d minb=%d pair=%d grid=%3d %s lanes=%d pdl=%d: %6.2f us/launch  %5.2f TB/s | last launch: %4.0f clk/tile, "
         "%5.1f GB/s/CTA, SMs with 0/1/2/3+ CTAs %d/%d/%d/%d  %s [%s]\n",
         TR, D, MINB, PAIR, grid, balanced ? "balanced" : "sliced  ", nstreams, pdl, ms * 1e3 / launches,
         bytes * launches / (ms * 1e-3) / 1e12, cyc / tiles, double(TR) * C * 4 / (cyc / tiles) * 1.965,
         hist[0], hist[1], hist[2], hist[3], what, cudaGetErrorString(err));
  for (auto& s : st) cudaStreamDestroy(s);
}

__global__ void fill_random(float* p, size_t n, unsigned int seed) {  // relu(pseudo-normal): like real conv5 features
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    unsigned int h = (unsigned int)(i * 2654435761u) ^ seed;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    const float u = (h >> 8) * (1.0f / 16777216.0f) - 0.5f;
    p[i] = u > 0.f ? 3.f * u : 0.f;
  }
}

int main(int argc, char** argv) {
  const bool zeros = (argc > 1 && argv[1][0] == 'z');
  printf("input data: %s\n", zeros ? "all zeros (cudaMemset)" : "relu(pseudo-random)");
  Bufs bufs;
  const size_t xbytes = size_t(B) * HW * C * 4;
  for (int i = 0; i < 7; ++i) {  // 7 x 51 MB = 360 MB > 2x L2
    float* p;
    cudaMalloc(&p, xbytes);
    if (zeros) cudaMemset(p, 0, xbytes);
    else fill_random<<<1184, 256>>>(p, xbytes / 4, 17u * i + 1u);
    bufs.X.push_back(p);
  }
  cudaMalloc(&bufs.w, C * 4);
  if (zeros) cudaMemset(bufs.w, 0, C * 4);
  else fill_random<<<8, 256>>>(bufs.w, C, 999u);
  cudaMalloc(&bufs.gpart, size_t(8) * 512 * C * 4);
  cudaMalloc(&bufs.smid, 4 * 1024);
  cudaMalloc(&bufs.cyc, 16 * 1024);
  const int L = 210;
  run<3, 3, 3, 1>(bufs, 128, 4, 0, 1, 0, L, "as shipped, no overlap");
  run<3, 3, 3, 1>(bufs, 128, 4, 0, 1, 1, L, "as shipped, PDL");
  run<3, 3, 3, 1>(bufs, 148, 0, 1, 1, 0, L, "balanced 148, no overlap");
  run<3, 3, 3, 1>(bufs, 148, 0, 1, 1, 1, L, "balanced 148");
  run<3, 3, 3, 1>(bufs, 148, 0, 1, 4, 1, L, "balanced 148, 4 lanes");
  run<3, 3, 3, 1>(bufs, 296, 0, 1, 1, 1, L, "balanced 296");
  run<4, 3, 2, 1>(bufs, 148, 0, 1, 1, 1, L, "32 KB x 3, 2/SM, balanced");
  run<4, 3, 2, 1>(bufs, 148, 0, 1, 2, 1, L, "32 KB x 3, 2/SM, balanced, 2 lanes");
  run<4, 6, 1, 1>(bufs, 148, 0, 1, 1, 1, L, "32 KB x 6, 1/SM, balanced");
  run<3, 4, 2, 2>(bufs, 148, 0, 1, 1, 1, L, "pair, 24 KB x 4, 2/SM, balanced");
  run<2, 6, 2, 1>(bufs, 148, 0, 1, 2, 1, L, "16 KB x 6, 2/SM, balanced, 2 lanes");
  run<2, 4, 3, 1>(bufs, 148, 0, 1, 4, 1, L, "16 KB x 4, 3/SM, balanced, 4 lanes");
  cudaError_t err = cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}

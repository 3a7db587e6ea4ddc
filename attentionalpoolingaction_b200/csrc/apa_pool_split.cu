// apa_pool_split.cu — the single-map head as a two-kernel pipeline that keeps HBM busy across launches.
// Same math as apa_pool_fused.cu / apa_pool_local.cu (reference sites nets_factory.py:259-328, :350-352):
//     g_r[b] = sum_hw A_r[b,hw] X[b,hw,:]      s_r[b] = sum_hw A_r[b,hw]
//     logits[v] = sum_frames sum_r (g_r . B_r + d_r s_r) / (HW F)
//
//   STREAM kernel    one CTA per (image, row slice); no cluster, no exchange.  A producer lane feeds 3-row x C
//                    tiles (24 KB, cp.async.bulk) through a 3-stage mbarrier ring; 8 consumer warps form a[row]
//                    and accumulate the slice's g in registers; at the end the CTA stores its PARTIAL g (8 KB per
//                    rank) and (attention mass, running max) to the workspace and exits.  74 KB of shared memory
//                    and 72 registers: three CTAs per SM, so the next launch's CTAs stream while this launch's
//                    finish, and nothing latency-bound holds a streaming slot.
//   CONTRACT kernel  16-CTA clusters split the channels (C/16 each); grid = class slices (<= 56 classes) x 16.
//                    A CTA sums the row slices (softmax max/sum merge) and frames of its channels for 32 videos,
//                    contracts them with its [C/16 x 56] block of B_r from shared memory, and the 16 partial
//                    results meet in a DSMEM push + one cluster barrier; fixed-order sum, bias, logits (or the
//                    peer stores of the fused all-gather).  53 KB of shared memory, 256 threads: it fits beside
//                    two streaming CTAs.
// Both are launched as programmatic dependents: STREAM(n) -> CONTRACT(n) -> STREAM(n+1) -> ...  CONTRACT
// triggers at once, prefetches its block of B_r, then waits for STREAM(n); STREAM(n+1) therefore starts while
// STREAM(n) is still running and only waits (griddepcontrol.wait) for CONTRACT(n) before it overwrites the
// partial-g workspace at its very end.
#include "apa_common.cuh"
#include "apa_kernels.h"

#include <atomic>
#include <mutex>

namespace apa {

namespace {

constexpr int kSTR = 3;       // rows per X tile
constexpr int kSStages = 3;   // ring depth
constexpr int kKT = 56;       // classes per contract CTA (lane, lane + 32)
constexpr int kCC = 16;       // contract cluster size (channel split)
constexpr int kCPer = kVideoTile * kKT / kCC;   // outputs reduced by one CTA per pass (112 = 2 videos)
constexpr int kContractThreads = 256;

__device__ __forceinline__ void sts_mark(unsigned long long* ts, int row, int slot) {
  if (ts != nullptr && row < 320) ts[size_t(row) * 64 + slot] = clock64();
}
__device__ __forceinline__ unsigned long long sglobaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float sdot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void saxpy4(float* f, float a, const float4& x) {
  f[0] = fmaf(a, x.x, f[0]);
  f[1] = fmaf(a, x.y, f[1]);
  f[2] = fmaf(a, x.z, f[2]);
  f[3] = fmaf(a, x.w, f[3]);
}

__host__ __device__ inline size_t stream_smem_bytes(int C) {
  return size_t(kSStages) * kSTR * C * 4 + 2 * 32 * 4 + 8 * 8;
}

// ============================================================================================ STREAM
template <int Q, int RMAX, int MODE>
__global__ void __launch_bounds__(kFusedThreads, (RMAX == 1) ? 3 : 1)
apa_stream_kernel(const FusedParams p) {
  constexpr int C = Q * 1024, TR = kSTR;
  constexpr size_t kTileBytes = size_t(TR) * C * 4;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* partS = reinterpret_cast<float*>(smem + kSStages * kTileBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kSStages * kTileBytes + 256);
  uint64_t* empty_bar = full_bar + 4;
  const int R = p.R, LP = p.lp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int units = p.B * LP;

  if (tid == kConsumerThreads) {
    for (int i = 0; i < kSStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kConsumerWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  // =========================================================================== producer warp
  if (warp == kConsumerWarps) {
    if (lane == 0) {
      const uint64_t pol_x = l2_policy_evict_first();
      uint32_t stage = 0, phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int b = u / LP, q = u - b * LP;
        const int row_lo = (p.HW * q) / LP, row_hi = (p.HW * (q + 1)) / LP;
        const int nrows = row_hi - row_lo, ntiles = (nrows + TR - 1) / TR;
        const float* xb = p.X + (size_t(b) * p.HW + row_lo) * C;
        for (int t = 0; t < ntiles; ++t) {
          const int rows = min(TR, nrows - t * TR);
          const uint32_t bytes = uint32_t(rows) * C * 4u;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          bulk_g2s(smem + stage * kTileBytes, xb + size_t(t) * TR * C, bytes, &full_bar[stage], pol_x);
          if (++stage == kSStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
    griddep_launch_dependents();
    return;
  }
  if (tid == 0) {
    sts_mark(p.ts, blockIdx.x, 0);
    if (p.ts && blockIdx.x < 320) {
      p.ts[size_t(blockIdx.x) * 64 + 62] = sglobaltimer_ns();
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.ts[size_t(blockIdx.x) * 64 + 61] = smid;
    }
  }

  // =========================================================================== consumers
  const int col0 = warp * (C / kConsumerWarps) + lane * 4;  // + qq*128
  float4 wv[Q];
  float c0 = 0.f;
  const float c_in = (p.attn_in != nullptr && p.attn_in_bias) ? __ldg(p.attn_b[0]) : 0.f;
  if (p.attn_in == nullptr) {
#pragma unroll
    for (int qq = 0; qq < Q; ++qq) wv[qq] = __ldg(reinterpret_cast<const float4*>(p.attn_w[0] + col0 + qq * 128));
    c0 = __ldg(p.attn_b[0]);
  }
  float cw[RMAX], cb[RMAX];
#pragma unroll
  for (int j = 0; j < RMAX; ++j) {
    cw[j] = (j >= 1 && j < R) ? __ldg(p.attn_w[j]) : 0.f;
    cb[j] = (j >= 1 && j < R) ? __ldg(p.attn_b[j]) : 0.f;
  }
  uint32_t stage = 0, phase = 0, giter = 0;
  bool waited = false;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int b = u / LP, q = u - b * LP;
    const int row_lo = (p.HW * q) / LP, row_hi = (p.HW * (q + 1)) / LP;
    const int nrows = row_hi - row_lo, ntiles = (nrows + TR - 1) / TR;
    const bool last_unit = (u + int(gridDim.x) >= units);
    float f[RMAX][Q * 4];
    float s[RMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      s[j] = 0.f;
#pragma unroll
      for (int i = 0; i < Q * 4; ++i) f[j][i] = 0.f;
    }
    for (int it = 0; it < ntiles; ++it, ++giter) {
      const int rows = min(TR, nrows - it * TR);
      const bool full = (rows == TR);
      if (last_unit && it == max(0, ntiles - 1 - p.trigger_lead)) griddep_launch_dependents();
      const float* st = reinterpret_cast<const float*>(smem + stage * kTileBytes) + col0;
      mbar_wait(&full_bar[stage], phase);
      if (tid == 0 && giter < 24) sts_mark(p.ts, blockIdx.x, 8 + int(giter));
      float4 x[TR][Q];
      if (full) {
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
          for (int qq = 0; qq < Q; ++qq) x[r][qq] = *reinterpret_cast<const float4*>(st + r * C + qq * 128);
      } else {
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
          for (int qq = 0; qq < Q; ++qq)
            x[r][qq] = (r < rows) ? *reinterpret_cast<const float4*>(st + r * C + qq * 128)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float a[TR];
      if (p.attn_in == nullptr) {
        float pd[TR];
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          pd[r] = sdot4(x[r][0], wv[0]);
#pragma unroll
          for (int qq = 1; qq < Q; ++qq) pd[r] += sdot4(x[r][qq], wv[qq]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        // 3 row sums per warp with one transposing butterfly: after it, lane 8r holds row r's warp sum
        const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
        float k0 = hi16 ? pd[2] : pd[0], s0 = hi16 ? pd[0] : pd[2];
        float k1 = hi16 ? 0.f : pd[1], s1 = hi16 ? pd[1] : 0.f;
        k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
        k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        float kk = hi8 ? k1 : k0;
        const float ss = hi8 ? k0 : k1;
        kk += __shfl_xor_sync(0xffffffffu, ss, 8);
        kk += __shfl_xor_sync(0xffffffffu, kk, 4);
        kk += __shfl_xor_sync(0xffffffffu, kk, 2);
        kk += __shfl_xor_sync(0xffffffffu, kk, 1);
        float* part = partS + (giter & 1u) * 32;
        if ((lane & 7) == 0) part[(lane >> 3) * 8 + warp] = kk;
        named_bar_sync(1, kConsumerThreads);
        float v = part[lane];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
#pragma unroll
        for (int r = 0; r < TR; ++r) a[r] = c0 + __shfl_sync(0xffffffffu, v, 8 * r);
      } else {
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          float v = 0.f;
          if (r < rows) {
            const size_t gi = size_t(b) * p.HW + row_lo + it * TR + r;
            v = c_in;
            for (int pp = 0; pp < p.attn_in_parts; ++pp) v += __ldg(p.attn_in + size_t(pp) * p.B * p.HW + gi);
          }
          a[r] = v;
        }
      }
      if (p.attn_raw != nullptr && tid == 0) {
#pragma unroll
        for (int r = 0; r < TR; ++r)
          if (r < rows) p.attn_raw[size_t(b) * p.HW + row_lo + it * TR + r] = a[r];
      }
      float valid[TR];
#pragma unroll
      for (int r = 0; r < TR; ++r) valid[r] = (full || r < rows) ? 1.f : 0.f;
      if (MODE == MODE_SOFTMAX) {
        float tmax = -INFINITY;
#pragma unroll
        for (int r = 0; r < TR; ++r) tmax = fmaxf(tmax, valid[r] != 0.f ? a[r] : -INFINITY);
        const float mnew = fmaxf(mx, tmax);
        const float sc = (mnew == -INFINITY) ? 1.f : expf(mx - mnew);
        mx = mnew;
        s[0] *= sc;
#pragma unroll
        for (int i = 0; i < Q * 4; ++i) f[0][i] *= sc;
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          const float e = valid[r] != 0.f ? expf(a[r] - mx) : 0.f;
          s[0] += e;
#pragma unroll
          for (int qq = 0; qq < Q; ++qq) saxpy4(&f[0][qq * 4], e, x[r][qq]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          float ar = a[r];
#pragma unroll
          for (int j = 0; j < RMAX; ++j) {
            if (j < R) {
              if (j >= 1) ar = fmaf(cw[j], ar, cb[j]);
              const float A = ((MODE == MODE_RELU) ? fmaxf(ar, 0.f) : ar) * valid[r];
              s[j] += A;
#pragma unroll
              for (int qq = 0; qq < Q; ++qq) saxpy4(&f[j][qq * 4], A, x[r][qq]);
            }
          }
        }
      }
      if (p.attn_in != nullptr) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
      }
      if (++stage == kSStages) { stage = 0; phase ^= 1u; }
    }
    if (last_unit && ntiles == 0) griddep_launch_dependents();
    if (tid == 0 && u == int(blockIdx.x)) sts_mark(p.ts, blockIdx.x, 2);
    // the previous CONTRACT launch (which reads this workspace) must have completely finished
    if (!waited) {
      griddep_wait();
      waited = true;
    }
#pragma unroll
    for (int j = 0; j < RMAX; ++j)
      if (j < R) {
        float* dst = p.gpart + (size_t(u) * R + j) * C + col0;
#pragma unroll
        for (int qq = 0; qq < Q; ++qq)
          *reinterpret_cast<float4*>(dst + qq * 128) =
              make_float4(f[j][qq * 4 + 0], f[j][qq * 4 + 1], f[j][qq * 4 + 2], f[j][qq * 4 + 3]);
      }
    if (tid == 0) {
#pragma unroll
      for (int j = 0; j < RMAX; ++j) p.gstat[size_t(u) * 8 + j] = s[j];
      p.gstat[size_t(u) * 8 + 4] = mx;
    }
  }
  if (!waited) griddep_wait();
  if (tid == 0) {
    sts_mark(p.ts, blockIdx.x, 5);
    if (p.ts && blockIdx.x < 320) p.ts[size_t(blockIdx.x) * 64 + 63] = sglobaltimer_ns();
  }
}

// ============================================================================================ CONTRACT
__host__ __device__ inline size_t contract_smem_bytes(int C) {
  const int CH = C / kCC;
  return size_t(CH) * kKT * 4 + size_t(kVideoTile) * CH * 4 + size_t(kCC) * kCPer * 4 + kVideoTile * 4 * 4;
}

template <bool SOFTMAX>
__global__ void __launch_bounds__(kContractThreads, 3)
apa_contract_kernel(const FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int C = p.C, K = p.K, R = p.R, F = p.F, LP = p.lp;
  const int CH = C / kCC;                                     // 128 (C = 2048) or 64
  float* Ws = reinterpret_cast<float*>(smem);                 // [CH][kKT]   block of B_r
  float* gs = Ws + CH * kKT;                                  // [32 videos][CH]
  float* slots = gs + kVideoTile * CH;                        // [16 sources][kCPer] partial results pushed to me
  float* sS = slots + kCC * kCPer;                            // [32 videos][4] attention mass per rank
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  const int slice = blockIdx.x / kCC, nslices = gridDim.x / kCC;
  const int k_lo = int((long long)K * slice / nslices), k_hi = int((long long)K * (slice + 1) / nslices);
  const int width = k_hi - k_lo;                              // <= kKT
  const int c_base = int(crank) * CH;
  const int Bv = p.B / F;
  const int ts_row = 160 + blockIdx.x;

  griddep_launch_dependents();  // the next STREAM launch may start streaming right away
  if (tid == 0) {
    sts_mark(p.ts, ts_row, 0);
    if (p.ts && ts_row < 320) p.ts[size_t(ts_row) * 64 + 62] = sglobaltimer_ns();
  }
  auto load_w = [&](int j) {
    const float* W = p.td_w[j];
    for (int idx = tid; idx < CH * kKT; idx += kContractThreads) {
      const int c = idx / kKT, t = idx - c * kKT;
      Ws[idx] = (t < width) ? __ldg(W + size_t(c_base + c) * K + k_lo + t) : 0.f;
    }
  };
  load_w(0);        // does not depend on the STREAM launch
  griddep_wait();   // STREAM(n) has finished and its partial g / stats are visible
  if (tid == 0) sts_mark(p.ts, ts_row, 1);

  if (SOFTMAX && p.smx != nullptr && blockIdx.x == 0) {  // (max, sum) per image for the attention end point
    for (int b = tid; b < p.B; b += kContractThreads) {
      float M = -INFINITY;
      for (int q = 0; q < LP; ++q) M = fmaxf(M, __ldcg(p.gstat + (size_t(b) * LP + q) * 8 + 4));
      float S = 0.f;
      for (int q = 0; q < LP; ++q)
        S += expf(__ldcg(p.gstat + (size_t(b) * LP + q) * 8 + 4) - M) * __ldcg(p.gstat + (size_t(b) * LP + q) * 8);
      p.smx[2 * b + 0] = M;
      p.smx[2 * b + 1] = S;
    }
  }

  const int nsl = F * LP;  // row slices (over all frames) summed into one video
  for (int v0 = 0; v0 < Bv; v0 += kVideoTile) {
    // attention mass per (video, rank): sum over frames and row slices (softmax: 1 per frame)
    if (tid < kVideoTile * 4) {
      const int vi = tid >> 2, j = tid & 3, v = v0 + vi;
      float mass = 0.f;
      if (v < Bv && j < R) {
        if (SOFTMAX) mass = float(F);
        else
          for (int e = 0; e < nsl; ++e) mass += __ldcg(p.gstat + (size_t(v) * nsl + e) * 8 + j);
      }
      sS[tid] = mass;
    }
    float acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
    for (int j = 0; j < R; ++j) {
      if (v0 != 0 || j != 0) {
        __syncthreads();  // previous operands fully consumed
        if (R > 1) load_w(j);
      }
      // ---- operand g: sum the row slices (softmax: rescaled to the image's max / sum) and the frames.
      //      warp = 4 videos, lane = 4 channels; slices of a video are contiguous: [(v F + fr) LP + q]
      if (lane * 4 < CH) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int vi = warp * 4 + i, v = v0 + vi;
          float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (v < Bv) {
            const float* gp = p.gpart + (size_t(v) * nsl * R + j) * C + c_base + lane * 4;
            const float* stp = p.gstat + size_t(v) * nsl * 8;
            for (int e0 = 0; e0 < nsl; e0 += 4) {
              float4 gv[4];
              float sc[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const bool ok = (e0 + u < nsl);
                gv[u] = ok ? __ldcg(reinterpret_cast<const float4*>(gp + size_t(e0 + u) * R * C))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
                sc[u] = 1.f;
                if (SOFTMAX && ok) {  // scale of slice e within its frame: exp(m_e - M) / S
                  const float* fs = stp + size_t((e0 + u) / LP) * LP * 8;
                  float M = -INFINITY;
                  for (int q = 0; q < LP; ++q) M = fmaxf(M, __ldcg(fs + q * 8 + 4));
                  float S = 0.f;
                  for (int q = 0; q < LP; ++q) S += expf(__ldcg(fs + q * 8 + 4) - M) * __ldcg(fs + q * 8);
                  sc[u] = expf(__ldcg(stp + size_t(e0 + u) * 8 + 4) - M) / S;
                }
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                a4.x = fmaf(sc[u], gv[u].x, a4.x);
                a4.y = fmaf(sc[u], gv[u].y, a4.y);
                a4.z = fmaf(sc[u], gv[u].z, a4.z);
                a4.w = fmaf(sc[u], gv[u].w, a4.w);
              }
            }
          }
          *reinterpret_cast<float4*>(gs + vi * CH + lane * 4) = a4;
        }
      }
      __syncthreads();
      if (tid == 0 && v0 == 0 && j == 0) sts_mark(p.ts, ts_row, 2);
      // ---- contraction: warp = 4 videos, lane = classes {2 lane, 2 lane + 1} (28 lanes x 2 = 56)
      if (lane * 2 < kKT) {
        const float* g0 = gs + (warp * 4) * CH;
        const float* w0p = Ws + lane * 2;
#pragma unroll 2
        for (int c = 0; c < CH; c += 4) {
          float4 gv[4];
          float2 wv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) gv[i] = *reinterpret_cast<const float4*>(g0 + i * CH + c);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) wv[cc] = *reinterpret_cast<const float2*>(w0p + (c + cc) * kKT);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(gv[i].x, wv[0].x, acc[i][0]);
            acc[i][1] = fmaf(gv[i].x, wv[0].y, acc[i][1]);
            acc[i][0] = fmaf(gv[i].y, wv[1].x, acc[i][0]);
            acc[i][1] = fmaf(gv[i].y, wv[1].y, acc[i][1]);
            acc[i][0] = fmaf(gv[i].z, wv[2].x, acc[i][0]);
            acc[i][1] = fmaf(gv[i].z, wv[2].y, acc[i][1]);
            acc[i][0] = fmaf(gv[i].w, wv[3].x, acc[i][0]);
            acc[i][1] = fmaf(gv[i].w, wv[3].y, acc[i][1]);
          }
        }
      }
    }
    if (tid == 0 && v0 == 0) sts_mark(p.ts, ts_row, 3);
    // ---- the 16 channel slices meet: push my partial results to the CTA that reduces those outputs
    if (lane * 2 < kKT) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int o = (warp * 4 + i) * kKT + lane * 2 + n;
          const uint32_t red = uint32_t(o / kCPer);
          st_dsmem_f(dsmem_addr(slots + size_t(crank) * kCPer + (o - int(red) * kCPer), red), acc[i][n]);
        }
    }
    cluster_sync_all();  // release my pushes / acquire everybody's
    if (tid < kCPer) {
      const int o = int(crank) * kCPer + tid;
      const int vi = o / kKT, t = o - vi * kKT;
      const int v = v0 + vi, k = k_lo + t;
      if (v < Bv && t < width) {
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < kCC; ++c) sum += slots[c * kCPer + tid];
        for (int j = 0; j < R; ++j) sum = fmaf(__ldg(p.td_b[j] + k), sS[vi * 4 + j], sum);
        const float val = sum * p.inv_scale;
        if (p.gw_world > 0) {
          const size_t off = (peer_gather_step(p.gw_done) % p.gw_depth) * size_t(p.gw_world) * Bv * p.K + (size_t(p.gw_rank) * Bv + size_t(v)) * K + k;
          for (int qq = 0; qq < p.gw_world; ++qq) p.gw_logits[qq][off] = val;
        } else {
          p.logits[size_t(v) * K + k] = val;
        }
      }
    }
    if (v0 + kVideoTile < Bv) cluster_sync_all();  // slots / sS are reused by the next pass
  }
  __syncthreads();  // this CTA's logits stores precede the gather bookkeeping below
  if (p.gw_world > 0 && tid == 0)
    peer_gather_finish(p.gw_world, p.gw_rank, p.gw_flags, p.gw_done, gridDim.x, peer_gather_step(p.gw_done));
  if (tid == 0) {
    sts_mark(p.ts, ts_row, 5);
    if (p.ts && ts_row < 320) p.ts[size_t(ts_row) * 64 + 63] = sglobaltimer_ns();
  }
}

using SKernelFn = void (*)(const FusedParams);

SKernelFn stream_select(int Q, int rmax, int mode) {
  if (Q == 2) {
    if (rmax == 1) {
      if (mode == MODE_PLAIN) return &apa_stream_kernel<2, 1, MODE_PLAIN>;
      if (mode == MODE_RELU) return &apa_stream_kernel<2, 1, MODE_RELU>;
      return &apa_stream_kernel<2, 1, MODE_SOFTMAX>;
    }
    if (mode == MODE_PLAIN) return &apa_stream_kernel<2, APA_MAX_RANK, MODE_PLAIN>;
    if (mode == MODE_RELU) return &apa_stream_kernel<2, APA_MAX_RANK, MODE_RELU>;
    return nullptr;
  }
  if (Q == 1) {
    if (rmax == 1) {
      if (mode == MODE_PLAIN) return &apa_stream_kernel<1, 1, MODE_PLAIN>;
      if (mode == MODE_RELU) return &apa_stream_kernel<1, 1, MODE_RELU>;
      return &apa_stream_kernel<1, 1, MODE_SOFTMAX>;
    }
    if (mode == MODE_PLAIN) return &apa_stream_kernel<1, APA_MAX_RANK, MODE_PLAIN>;
    if (mode == MODE_RELU) return &apa_stream_kernel<1, APA_MAX_RANK, MODE_RELU>;
  }
  return nullptr;
}

std::mutex g_smu;
struct SEntry { int dev; const void* fn; int sms; };
SEntry g_scache[32];
int g_sn = 0;
std::atomic<int> g_split_pdl_refused{0};

// one-time attributes per (device, kernel); returns the SM count
cudaError_t split_prepare(const void* fn, size_t smem, bool cluster16, int* sms_out) {
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  {
    std::lock_guard<std::mutex> lk(g_smu);
    for (int i = 0; i < g_sn; ++i)
      if (g_scache[i].dev == dev && g_scache[i].fn == fn) {
        *sms_out = g_scache[i].sms;
        return cudaSuccess;
      }
  }
  err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (err != cudaSuccess) return err;
  if (cluster16) {
    err = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (err != cudaSuccess) return err;
  }
  (void)cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int sms = 0;
  err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (err != cudaSuccess) return err;
  std::lock_guard<std::mutex> lk(g_smu);
  if (g_sn < 32) g_scache[g_sn++] = {dev, fn, sms};
  *sms_out = sms;
  return cudaSuccess;
}

}  // namespace

bool split_eligible(int C, int R) { return (C == 1024 || C == 2048) && R >= 1 && R <= APA_MAX_RANK; }

int split_lp_cap(int B, int HW) {
  int cap = 16;
  while (cap > 1 && (size_t(B) * cap > 1024 || HW / cap < 3)) cap >>= 1;
  return cap;
}

int split_lp(int B, int HW, const ApaTuning* tuning) {
  const int cap = split_lp_cap(B, HW);
  int lp = 1;
  while (lp < cap && B * lp < 120) lp <<= 1;  // enough CTAs to cover the machine once
  if (tuning && tuning->cluster_size > 0) {
    lp = 1;
    while (lp * 2 <= tuning->cluster_size && lp < 16) lp <<= 1;
    if (lp > cap) lp = cap;
  }
  return lp;
}

int split_contract_grid(int K) { return kCC * ((K + kKT - 1) / kKT); }
size_t split_stream_smem(int C) { return stream_smem_bytes(C); }

cudaError_t split_launch(const FusedParams& p_in, int mode, const ApaTuning* tuning, cudaStream_t stream,
                         int* stream_grid, int* contract_grid) {
  FusedParams p = p_in;
  const int Q = p.C / 1024, rmax = (p.R == 1) ? 1 : APA_MAX_RANK;
  SKernelFn fn = stream_select(Q, rmax, mode);
  if (fn == nullptr) return cudaErrorInvalidValue;
  p.lp = split_lp(p.B, p.HW, tuning);
  const size_t smem1 = stream_smem_bytes(p.C), smem2 = contract_smem_bytes(p.C);
  int sms = 0;
  cudaError_t err = split_prepare(reinterpret_cast<const void*>(fn), smem1, false, &sms);
  if (err != cudaSuccess) return err;
  SKernelFn cfn = (mode == MODE_SOFTMAX) ? &apa_contract_kernel<true> : &apa_contract_kernel<false>;
  err = split_prepare(reinterpret_cast<const void*>(cfn), smem2, true, &sms);
  if (err != cudaSuccess) return err;
  const int units = p.B * p.lp;
  int grid1 = units;
  const int cap1 = sms * ((rmax == 1) ? 3 : 1);
  if (tuning && tuning->max_ctas > 0 && grid1 > tuning->max_ctas) grid1 = tuning->max_ctas;
  if (grid1 > cap1) grid1 = cap1;
  const int grid2 = split_contract_grid(p.K);
  if (stream_grid) *stream_grid = grid1;
  if (contract_grid) *contract_grid = grid2;
  const bool pdl = !(tuning && tuning->no_pdl) && !g_split_pdl_refused.load(std::memory_order_relaxed);

  cudaLaunchAttribute at[2];
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid1, 1, 1);
  cfg.blockDim = dim3(kFusedThreads, 1, 1);
  cfg.dynamicSmemBytes = smem1;
  cfg.stream = stream;
  int na = 0;
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  err = cudaLaunchKernelEx(&cfg, fn, p);
  if (err != cudaSuccess && pdl) {
    (void)cudaGetLastError();
    g_split_pdl_refused.store(1, std::memory_order_relaxed);
    cfg.numAttrs = 0;
    err = cudaLaunchKernelEx(&cfg, fn, p);
  }
  if (err != cudaSuccess) return err;

  cudaLaunchConfig_t cfg2 = {};
  cfg2.gridDim = dim3(grid2, 1, 1);
  cfg2.blockDim = dim3(kContractThreads, 1, 1);
  cfg2.dynamicSmemBytes = smem2;
  cfg2.stream = stream;
  cudaLaunchAttribute at2[2];
  at2[0].id = cudaLaunchAttributeClusterDimension;
  at2[0].val.clusterDim.x = kCC; at2[0].val.clusterDim.y = 1; at2[0].val.clusterDim.z = 1;
  int na2 = 1;
  const bool pdl2 = pdl && !g_split_pdl_refused.load(std::memory_order_relaxed);
  if (pdl2) {
    at2[na2].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at2[na2].val.programmaticStreamSerializationAllowed = 1;
    ++na2;
  }
  cfg2.attrs = at2; cfg2.numAttrs = na2;
  err = cudaLaunchKernelEx(&cfg2, cfn, p);
  if (err != cudaSuccess && pdl2) {
    (void)cudaGetLastError();
    g_split_pdl_refused.store(1, std::memory_order_relaxed);
    cfg2.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfg2, cfn, p);
  }
  return err;
}

}  // namespace apa

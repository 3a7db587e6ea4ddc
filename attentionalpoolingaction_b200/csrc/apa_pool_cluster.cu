// apa_pool_cluster.cu — the cluster form of the single-map head: everything after the HBM stream stays inside
// one 16-CTA thread-block cluster, so there is NO grid barrier and the attention-weighted features g never
// leave shared memory.  Same math as apa_pool_flow.cu (single attention map, rank 1..4, plain / relu /
// spatial softmax); reference sites nets_factory.py:259-328, :350-352.
//
//   cluster = 16 CTAs = LI images x LP row slices (LI x LP = 16; LI = 4 or 8).  CTA rank r: image r / LP, slice r % LP.
//   phase 1   producer lane streams TR-row x C tiles (cp.async.bulk) through an
//             STG-stage mbarrier ring; 8 consumer warps form a[row] and accumulate g in registers.
//   exchange 1  phase 2 splits the contraction g.B over the 16 CTAs by CHANNEL: CTA j owns channels
//             [j C/16, (j+1) C/16) for all LI images and all K classes.  So every CTA PUSHES (st.shared::cluster)
//             each 128-channel block of its partial g straight to the CTA that will contract it, plus its
//             (attention mass, running max) to everybody; one 16-arrival mbarrier round publishes them and
//             the owner sums the row slices of each image (softmax max/sum merge) into the operand layout.
//   phase 2   CTA j's rows of B_r are one contiguous, 16-byte aligned span, so the same producer lane streams
//             them with cp.async.bulk through the same ring right behind the last X tile (in flight while
//             exchange 1 runs).  Each thread owns classes {t, t+256, t+512}; LI images x 3 classes of fp32
//             accumulators; g is broadcast LDS.128s per channel.
//   exchange 2  every CTA pushes its partial logits to the CTA that reduces those outputs (1/16 each); a second
//             16-arrival round, then a local fixed-order sum (bit-reproducible), bias, logits.
// Cluster-scope mbarriers (remote arrive.release / local try_wait.acquire) order both exchanges; after its last
// wait nothing remote touches a CTA, so CTAs exit without a trailing cluster barrier.  A third barrier
// (slots_free) exists only for clusters that process more than one group of LI images.
// (TR, STG, LI) come in three shapes, see kShapes below and DESIGN.md section 4.2.
#include "apa_common.cuh"
#include "apa_kernels.h"

#include <atomic>
#include <mutex>

namespace apa {

namespace {

constexpr int kLC = 16;   // cluster size
constexpr int kLI = 4;    // images per cluster in the default shapes (8 in the throughput shape)
// Two shapes of the ring: (TR=4 rows, 3 stages) -> 2 CTAs per SM, and (TR=3 rows, 2 stages) -> 3 CTAs per SM
// (<= 76,800 B of shared memory and <= 72 registers per thread), which keeps three launches resident per SM.

template <int Q, int RMAX, int TR_, int STG>
struct LSmem {
  static constexpr int C = Q * 1024;
  static constexpr int CH = C / kLC;      // channels contracted by one CTA in phase 2
  static constexpr int kStages = STG;
};

struct LLayout {  // byte offsets after the ring (computed on host and device identically)
  size_t off_gslots, off_outslots, off_stat, off_part, off_scl, off_bars, total;
  size_t stage_bytes;  // one ring stage: TR rows of X, or WR rows of B_r, whichever is larger
  int per;             // outputs reduced by one CTA
};
__host__ __device__ inline int local_wr(int K) { return (K <= 512) ? 16 : 8; }  // weight rows per ring stage
__host__ __device__ inline LLayout local_layout(int C, int R, int K, int TR, int STG, int LI) {
  LLayout l;
  l.per = (LI * K + kLC - 1) / kLC;
  const size_t xb = size_t(TR) * C * 4, wb = size_t(local_wr(K)) * K * 4;
  l.stage_bytes = ((xb > wb ? xb : wb) + 1023) / 1024 * 1024;
  size_t off = size_t(STG) * l.stage_bytes;
  // partial g pushed by every CTA of the cluster: [image][quarter][r][C/16]; after the local sum the head of
  // this region is reused for gS[r][C/16][image] (the operand layout of phase 2)
  l.off_gslots = off;   off += size_t(kLC) * R * (C / kLC) * 4;
  l.off_outslots = off; off += (size_t(kLC) * l.per * 4 + 127) / 128 * 128;   // partial logits [source][per]
  l.off_stat = off;     off += kLC * 8 * 4;                                   // (s_r, max) of every CTA
  l.off_part = off;     off += 2 * 4 * 8 * 4;
  l.off_scl = off;      off += kLC * 4;
  l.off_bars = off;     off += 8 * 16;
  l.total = off;
  return l;
}

__device__ __forceinline__ void lts_mark(unsigned long long* ts, int slot) {
  if (ts != nullptr) ts[size_t(blockIdx.x) * 64 + slot] = clock64();
}
__device__ __forceinline__ unsigned long long lglobaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float ldot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void laxpy4(float* f, float a, const float4& x) {
  f[0] = fmaf(a, x.x, f[0]);
  f[1] = fmaf(a, x.y, f[1]);
  f[2] = fmaf(a, x.z, f[2]);
  f[3] = fmaf(a, x.w, f[3]);
}

template <int Q, int RMAX, int MODE, int NC, int TR, int STG, int LI>
__global__ void __launch_bounds__(kFusedThreads, (RMAX == 1) ? ((STG == 2) ? 3 : 2) : 1)
apa_pool_local_kernel(const FusedParams p) {
  using L = LSmem<Q, RMAX, TR, STG>;
  constexpr int C = L::C, CH = L::CH;
  constexpr int LP = kLC / LI;  // CTAs (row slices) per image
  static_assert(LI == 4 || LI == 8, "images per cluster");
  static_assert(TR == 3 || TR == 4, "the row butterfly below is written for up to 4 rows per tile");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int R = p.R, K = p.K;
  const LLayout lay = local_layout(C, R, K, TR, STG, LI);
  const size_t stage_bytes = lay.stage_bytes;
  float* gslots = reinterpret_cast<float*>(smem + lay.off_gslots);      // [4 images][4 quarters][R][CH]
  float* gS = gslots;                                                   // [R][CH][4 images] after the sum
  float* outslots = reinterpret_cast<float*>(smem + lay.off_outslots);  // [16 sources][per]
  float* statS = reinterpret_cast<float*>(smem + lay.off_stat);         // [16 CTAs][8]
  const int per = lay.per;
  float* partS = reinterpret_cast<float*>(smem + lay.off_part);
  float* scl = reinterpret_cast<float*>(smem + lay.off_scl);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.off_bars);
  uint64_t* full_bar = bars;          // [3]
  uint64_t* empty_bar = bars + 4;     // [3]
  uint64_t* parts_ready = bars + 8;   // 16 arrivals: every CTA pushed its partial g blocks and stats here
  uint64_t* out_ready = bars + 9;     // 16 arrivals: every CTA pushed its partial logits here
  uint64_t* slots_free = bars + 10;   // 16 arrivals: every CTA consumed what was pushed to it (multi-group)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  const int img_in_cluster = int(crank) / LP, q = int(crank) % LP;
  const int cluster_id = blockIdx.x / kLC, nclusters = gridDim.x / kLC;
  const int ngroups = p.B / LI;
  const int row_lo = (p.HW * q) / LP, row_hi = (p.HW * (q + 1)) / LP;
  const int nrows = row_hi - row_lo, ntiles = (nrows + TR - 1) / TR;
  const int WR = local_wr(K);                    // weight rows per ring stage
  const int wchunks = CH / WR;                   // per rank

  // Reading X (or a_0 from the pose GEMM, the kernel right before this one) before the previous kernel in the
  // stream has completed is only legal when the caller vouches for it (ApaTuning.inputs_stable)
  if (!p.inputs_stable || p.attn_in != nullptr || p.attn_raw != nullptr) griddep_wait();
  // =========================================================================== producer warp
  if (warp == kConsumerWarps) {
    if (lane == 0) {
      for (int i = 0; i < L::kStages; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], kConsumerWarps);
      }
      mbar_init(parts_ready, kLC);
      mbar_init(out_ready, kLC);
      mbar_init(slots_free, kLC);
      fence_mbar_init();
    }
    __syncwarp();
    cluster_arrive();
    if (lane == 0) {
      const uint64_t pol_x = l2_policy_evict_first(), pol_w = l2_policy_evict_last();
      uint32_t stage = 0, phase = 0;
      for (int g = cluster_id; g < ngroups; g += nclusters) {
        const int b = g * LI + img_in_cluster;
        const float* xb = p.X + (size_t(b) * p.HW + row_lo) * C;
        for (int t = 0; t < ntiles; ++t) {
          const int rows = min(TR, nrows - t * TR);
          const uint32_t bytes = uint32_t(rows) * C * 4u;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          bulk_g2s(smem + stage * stage_bytes, xb + size_t(t) * TR * C, bytes, &full_bar[stage], pol_x);
          if (++stage == L::kStages) { stage = 0; phase ^= 1u; }
        }
        // this CTA's weight rows for phase 2 ride the same ring (they do not depend on phase 1)
        for (int j = 0; j < R; ++j) {
          const float* wsrc = p.td_w[j] + size_t(int(crank) * CH) * K;
          for (int ch = 0; ch < wchunks; ++ch) {
            const uint32_t bytes = uint32_t(WR) * K * 4u;
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full_bar[stage], bytes);
            bulk_g2s(smem + stage * stage_bytes, wsrc + size_t(ch) * WR * K, bytes, &full_bar[stage], pol_w);
            if (++stage == L::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
    griddep_launch_dependents();
    cluster_wait();
    return;  // the producer warp takes no part in the exchanges (they are mbarrier-based)
  }
  if (tid == 0) {
    lts_mark(p.ts, 0);
    if (p.ts) p.ts[size_t(blockIdx.x) * 64 + 62] = lglobaltimer_ns();
  }
  cluster_arrive();
  cluster_wait();  // every CTA's barriers are initialised
  if (tid == 0) lts_mark(p.ts, 1);

  // =========================================================================== consumers
  const int col0 = warp * (C / kConsumerWarps) + lane * 4;  // + q*128
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
  uint32_t stage = 0, phase = 0, giter = 0, grp_iter = 0;
  for (int g = cluster_id; g < ngroups; g += nclusters, ++grp_iter) {
    const int b = g * LI + img_in_cluster;
    const bool last_group = (g + nclusters >= ngroups);
    float f[RMAX][Q * 4];
    float s[RMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      s[j] = 0.f;
#pragma unroll
      for (int i = 0; i < Q * 4; ++i) f[j][i] = 0.f;
    }
    // ------------------------------------------------------------------ phase 1: stream this CTA's rows
    for (int it = 0; it < ntiles; ++it, ++giter) {
      const int rows = min(TR, nrows - it * TR);
      const bool full = (rows == TR);
      if (last_group && it == max(0, ntiles - 1 - p.trigger_lead)) griddep_launch_dependents();
      const float* st = reinterpret_cast<const float*>(smem + stage * stage_bytes) + col0;
      mbar_wait(&full_bar[stage], phase);
      if (tid == 0 && giter < 24) lts_mark(p.ts, 8 + int(giter));
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
          pd[r] = ldot4(x[r][0], wv[0]);
#pragma unroll
          for (int qq = 1; qq < Q; ++qq) pd[r] += ldot4(x[r][qq], wv[qq]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
        const float pd3 = (TR > 3) ? pd[TR - 1] : 0.f;
        float k0 = hi16 ? pd[2] : pd[0], s0 = hi16 ? pd[0] : pd[2];
        float k1 = hi16 ? pd3 : pd[1], s1 = hi16 ? pd[1] : pd3;
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
          for (int qq = 0; qq < Q; ++qq) laxpy4(&f[0][qq * 4], e, x[r][qq]);
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
              for (int qq = 0; qq < Q; ++qq) laxpy4(&f[j][qq * 4], A, x[r][qq]);
            }
          }
        }
      }
      if (p.attn_in != nullptr) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
      }
      if (++stage == L::kStages) { stage = 0; phase ^= 1u; }
    }

    if (tid == 0 && grp_iter == 0) lts_mark(p.ts, 2);
    // ------------------------------------------------------------------ exchange 1: partial g -> phase-2 owners
    if (grp_iter > 0) mbar_wait_cluster(slots_free, (grp_iter - 1) & 1u);  // everyone consumed the last group
    {
#pragma unroll
      for (int qq = 0; qq < Q; ++qq) {
        const int c = col0 + qq * 128;
        const uint32_t owner = uint32_t(c / CH);
        const int cl = c - int(owner) * CH;
#pragma unroll
        for (int j = 0; j < RMAX; ++j)
          if (j < R)
            st_dsmem_f4(dsmem_addr(gslots + ((size_t(img_in_cluster) * LP + q) * R + j) * CH + cl, owner),
                        make_float4(f[j][qq * 4 + 0], f[j][qq * 4 + 1], f[j][qq * 4 + 2], f[j][qq * 4 + 3]));
      }
      if (tid < kLC) {  // my (s_r, max) to every CTA's stat table
        const uint32_t ra = dsmem_addr(statS + size_t(crank) * 8, uint32_t(tid));
#pragma unroll
        for (int j = 0; j < RMAX; ++j) st_dsmem_f(ra + 4u * j, s[j]);
        st_dsmem_f(ra + 16u, mx);
      }
      fence_acq_rel_cluster();
      named_bar_sync(1, kConsumerThreads);
      if (tid < kLC) mbar_arrive_cluster(parts_ready, uint32_t(tid));
      mbar_wait_cluster(parts_ready, grp_iter & 1u);
    }
    if (tid == 0 && grp_iter == 0) lts_mark(p.ts, 3);
    // per (image, quarter) softmax scale (1 otherwise); every CTA holds the stats of all 16 CTAs
    if (tid < kLC) {
      float v = 1.f;
      if (MODE == MODE_SOFTMAX) {
        const int i4 = tid & ~(LP - 1);
        float M = -INFINITY;
        for (int pp = 0; pp < LP; ++pp) M = fmaxf(M, statS[(i4 + pp) * 8 + 4]);
        float S = 0.f;
        for (int pp = 0; pp < LP; ++pp) S += expf(statS[(i4 + pp) * 8 + 4] - M) * statS[(i4 + pp) * 8 + 0];
        v = expf(statS[tid * 8 + 4] - M) / S;
        if ((tid % LP) == 0 && tid == int(crank) && p.smx != nullptr) {  // slice 0 of each image records (max, sum)
          p.smx[2 * (g * LI + (tid / LP)) + 0] = M;
          p.smx[2 * (g * LI + (tid / LP)) + 1] = S;
        }
      }
      scl[tid] = v;
    }
    named_bar_sync(1, kConsumerThreads);
    {
      // sum the row slices of every (image, r, channel) into registers, then rewrite the region as gS[r][ch][image]
      constexpr int PT = (LI * RMAX * CH + kConsumerThreads - 1) / kConsumerThreads;
      float sums[PT];
      const int total = LI * R * CH;
#pragma unroll
      for (int u = 0; u < PT; ++u) {
        const int e = tid + u * kConsumerThreads;  // e = (i * R + j) * CH + c
        float acc = 0.f;
        if (e < total) {
          const int i = e / (R * CH), rem = e - i * (R * CH);
#pragma unroll
          for (int pp = 0; pp < LP; ++pp)
            acc = fmaf(scl[i * LP + pp], gslots[(size_t(i) * LP + pp) * R * CH + rem], acc);
        }
        sums[u] = acc;
      }
      named_bar_sync(1, kConsumerThreads);
#pragma unroll
      for (int u = 0; u < PT; ++u) {
        const int e = tid + u * kConsumerThreads;
        if (e < total) {
          const int i = e / (R * CH), rem = e - i * (R * CH);  // rem = j * CH + c
          gS[size_t(rem) * LI + i] = sums[u];
        }
      }
      named_bar_sync(1, kConsumerThreads);
    }
    if (tid == 0 && grp_iter == 0) lts_mark(p.ts, 6);

    // ------------------------------------------------------------------ phase 2: my channels x all classes
    float acc[LI][NC];
#pragma unroll
    for (int i = 0; i < LI; ++i)
#pragma unroll
      for (int n = 0; n < NC; ++n) acc[i][n] = 0.f;
    for (int j = 0; j < R; ++j) {
      for (int ch = 0; ch < wchunks; ++ch) {
        const float* ws = reinterpret_cast<const float*>(smem + stage * stage_bytes);
        mbar_wait(&full_bar[stage], phase);
        const float* gj = gS + (size_t(j) * CH + ch * WR) * LI;
        for (int c = 0; c < WR; ++c) {
          float4 gv[LI / 4];  // all images of the cluster, broadcast
#pragma unroll
          for (int h = 0; h < LI / 4; ++h) gv[h] = *reinterpret_cast<const float4*>(gj + c * LI + h * 4);
#pragma unroll
          for (int n = 0; n < NC; ++n) {
            const int k = tid + n * kConsumerThreads;
            const float w = (k < K) ? ws[c * K + k] : 0.f;
#pragma unroll
            for (int h = 0; h < LI / 4; ++h) {
              acc[h * 4 + 0][n] = fmaf(gv[h].x, w, acc[h * 4 + 0][n]);
              acc[h * 4 + 1][n] = fmaf(gv[h].y, w, acc[h * 4 + 1][n]);
              acc[h * 4 + 2][n] = fmaf(gv[h].z, w, acc[h * 4 + 2][n]);
              acc[h * 4 + 3][n] = fmaf(gv[h].w, w, acc[h * 4 + 3][n]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == L::kStages) { stage = 0; phase ^= 1u; }
      }
    }
    if (last_group && ntiles == 0) griddep_launch_dependents();
    if (tid == 0 && grp_iter == 0) lts_mark(p.ts, 7);

    // ------------------------------------------------------------------ exchange 2: partial logits -> reducers
    {
#pragma unroll
      for (int i = 0; i < LI; ++i)
#pragma unroll
        for (int n = 0; n < NC; ++n) {
          const int k = tid + n * kConsumerThreads;
          if (k < K) {
            const int o = i * K + k;
            const uint32_t red = uint32_t(o / per);
            st_dsmem_f(dsmem_addr(outslots + size_t(crank) * per + (o - int(red) * per), red), acc[i][n]);
          }
        }
      fence_acq_rel_cluster();
      named_bar_sync(1, kConsumerThreads);
      if (tid < kLC) mbar_arrive_cluster(out_ready, uint32_t(tid));
      mbar_wait_cluster(out_ready, grp_iter & 1u);
    }
    if (tid == 0 && grp_iter == 0) lts_mark(p.ts, 4);
    if (grp_iter == 0) griddep_wait();  // first global write of this launch: previous launch fully done
    {
      const int nout = LI * K;
      const int o = int(crank) * per + tid;
      if (tid < per && o < nout) {
        const int i = o / K, k = o - i * K;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < kLC; ++c) sum += outslots[c * per + tid];
        for (int j = 0; j < R; ++j) {
          float S = 1.f;
          if (MODE != MODE_SOFTMAX) {
            S = 0.f;
            for (int pp = 0; pp < LP; ++pp) S += statS[(i * LP + pp) * 8 + j];
          }
          sum = fmaf(__ldg(p.td_b[j] + k), S, sum);
        }
        const float val = sum * p.inv_scale;
        if (p.gw_world > 0) {
          const size_t off = (peer_gather_step(p.gw_done) % p.gw_depth) * size_t(p.gw_world) * p.B * p.K + (size_t(p.gw_rank) * p.B + size_t(g * LI + i)) * K + k;
          for (int qq = 0; qq < p.gw_world; ++qq) p.gw_logits[qq][off] = val;
        } else {
          p.logits[size_t(g * LI + i) * K + k] = val;
        }
      }
    }
    if (!last_group) {  // tell everyone my slots / stat table may be overwritten by the next group's pushes
      named_bar_sync(1, kConsumerThreads);
      if (tid < kLC) mbar_arrive_cluster(slots_free, uint32_t(tid));
    }
  }
  named_bar_sync(1, kConsumerThreads);  // this CTA's logits stores precede the gather bookkeeping below
  if (p.gw_world > 0 && tid == 0) peer_gather_finish(p.gw_world, p.gw_rank, p.gw_flags, p.gw_done, gridDim.x, peer_gather_step(p.gw_done));
  if (tid == 0) {
    lts_mark(p.ts, 5);
    if (p.ts) p.ts[size_t(blockIdx.x) * 64 + 63] = lglobaltimer_ns();
  }
}

using LKernelFn = void (*)(const FusedParams);

template <int Q, int RMAX, int MODE, int TR, int STG, int LI>
LKernelFn lsel_nc(int nc) {
  if (nc == 1) return &apa_pool_local_kernel<Q, RMAX, MODE, 1, TR, STG, LI>;
  if (nc == 2) return &apa_pool_local_kernel<Q, RMAX, MODE, 2, TR, STG, LI>;
  return &apa_pool_local_kernel<Q, RMAX, MODE, 3, TR, STG, LI>;
}
// The three shapes of the kernel.  slots = resident CTAs per SM the ring is sized for.
//   0: 4-row tiles x 3 stages, 4 images x 4 row-quarters per cluster   (2 CTAs / SM; every rank)
//   1: 3-row tiles x 2 stages, 4 images x 4 row-quarters per cluster   (3 CTAs / SM; rank 1)
//   2: 3-row tiles x 2 stages, 8 images x 2 row-halves per cluster     (3 CTAs / SM; rank 1, B % 8 == 0):
//      half as many CTAs per launch, each streaming twice as long before the same latency-bound tail, so
//      more launches co-reside and the tail costs half the CTA-slot time (throughput shape)
struct LShape { int tr, stg, li; };
constexpr LShape kShapes[3] = {{4, 3, 4}, {3, 2, 4}, {3, 2, 8}};

LKernelFn local_select(int Q, int rmax, int mode, int nc, int shape) {
  if (Q != 2) return nullptr;
  if (rmax == 1) {
    if (shape == 2) {
      if (mode == MODE_PLAIN) return lsel_nc<2, 1, MODE_PLAIN, 3, 2, 8>(nc);
      if (mode == MODE_RELU) return lsel_nc<2, 1, MODE_RELU, 3, 2, 8>(nc);
      return lsel_nc<2, 1, MODE_SOFTMAX, 3, 2, 8>(nc);
    }
    if (shape == 1) {
      if (mode == MODE_PLAIN) return lsel_nc<2, 1, MODE_PLAIN, 3, 2, 4>(nc);
      if (mode == MODE_RELU) return lsel_nc<2, 1, MODE_RELU, 3, 2, 4>(nc);
      return lsel_nc<2, 1, MODE_SOFTMAX, 3, 2, 4>(nc);
    }
    if (mode == MODE_PLAIN) return lsel_nc<2, 1, MODE_PLAIN, 4, 3, 4>(nc);
    if (mode == MODE_RELU) return lsel_nc<2, 1, MODE_RELU, 4, 3, 4>(nc);
    return lsel_nc<2, 1, MODE_SOFTMAX, 4, 3, 4>(nc);
  }
  if (mode == MODE_PLAIN) return lsel_nc<2, APA_MAX_RANK, MODE_PLAIN, 4, 3, 4>(nc);
  if (mode == MODE_RELU) return lsel_nc<2, APA_MAX_RANK, MODE_RELU, 4, 3, 4>(nc);
  return nullptr;
}
constexpr size_t kThreeSlotSmem = 76800;  // (228 KB - 3 x 1 KB reserved) / 3
size_t shape_smem(int C, int R, int K, int shape) {
  return local_layout(C, R, K, kShapes[shape].tr, kShapes[shape].stg, kShapes[shape].li).total;
}

std::mutex g_lmu;
struct LEntry { int dev; LKernelFn fn; size_t smem; int nmax; };
LEntry g_lcache[32];
int g_ln = 0;
std::atomic<int> g_local_pdl_refused{0};

}  // namespace

bool cluster_eligible(int B, int HW, int C, int K, int R, int F) {
  (void)HW;
  return C == 2048 && F == 1 && B % kLI == 0 && B >= kLI && K >= 1 && K <= 3 * kConsumerThreads && R >= 1 &&
         R <= APA_MAX_RANK && (size_t(K) * 8 * 4) <= size_t(32) * 1024;
}

int cluster_shape_of(int B, int C, int R, int K, const ApaTuning* tuning) {
  // cluster_shape 1 / 2 / 3 forces shape 0 / 1 / 2 (when the problem allows it); otherwise the widest that fits
  const int force = tuning ? tuning->cluster_shape : 0;
  if (R != 1 || force == 1) return 0;
  if (force == 2) return 1;
  if (B % 8 == 0 && (force == 3 || shape_smem(C, R, K, 2) <= kThreeSlotSmem)) return 2;
  return (shape_smem(C, R, K, 1) <= kThreeSlotSmem) ? 1 : 0;
}

cudaError_t cluster_launch(const FusedParams& p_in, int mode, const ApaTuning* tuning, cudaStream_t stream, int* grid_out) {
  FusedParams p = p_in;
  // measured best for this form: let the next launch be scheduled as soon as a CTA starts consuming
  p.trigger_lead = (tuning && tuning->pdl_trigger_lead > 0) ? tuning->pdl_trigger_lead : 1 << 20;
  const int Q = p.C / 1024, rmax = (p.R == 1) ? 1 : APA_MAX_RANK;
  const int nc = (p.K + kConsumerThreads - 1) / kConsumerThreads;
  const int shape = cluster_shape_of(p.B, p.C, p.R, p.K, tuning);
  const int li = kShapes[shape].li;
  LKernelFn fn = local_select(Q, rmax, mode, nc, shape);
  if (fn == nullptr) return cudaErrorInvalidValue;
  const size_t smem = shape_smem(p.C, p.R, p.K, shape);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  int nmax = -1;
  {
    std::lock_guard<std::mutex> lk(g_lmu);
    for (int i = 0; i < g_ln; ++i)
      if (g_lcache[i].dev == dev && g_lcache[i].fn == fn && g_lcache[i].smem == smem) nmax = g_lcache[i].nmax;
  }
  if (nmax < 0) {
    err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (err != cudaSuccess) return err;
    (void)cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kLC, 1, 1);
    cfg.blockDim = dim3(kFusedThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kLC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms > 0 && n > sms / kLC) n = sms / kLC;  // one CTA slot per SM: the next launch co-resides
    nmax = n;
    std::lock_guard<std::mutex> lk(g_lmu);
    if (g_ln < 32) g_lcache[g_ln++] = {dev, fn, smem, nmax};
  }
  if (nmax <= 0) return cudaErrorInvalidConfiguration;
  int ncl = p.B / li;
  if (ncl > nmax) ncl = nmax;
  if (grid_out) *grid_out = ncl * kLC;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ncl * kLC, 1, 1);
  cfg.blockDim = dim3(kFusedThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kLC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  int na = 1;
  const bool pdl = !(tuning && tuning->no_pdl) && !g_local_pdl_refused.load(std::memory_order_relaxed);
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  err = cudaLaunchKernelEx(&cfg, fn, p);
  if (err != cudaSuccess && pdl) {
    (void)cudaGetLastError();
    g_local_pdl_refused.store(1, std::memory_order_relaxed);
    cfg.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfg, fn, p);
  }
  return err;
}

void cluster_plan(int B, int C, int R, int K, const ApaTuning* tuning, int* images_per_cluster, size_t* smem) {
  const int shape = cluster_shape_of(B, C, R, K, tuning);
  if (images_per_cluster) *images_per_cluster = kShapes[shape].li;
  if (smem) *smem = shape_smem(C, R, K, shape);
}

}  // namespace apa

// apa_pool_fused.cu — the headline kernel: single-map (M == 1) attentional pooling, rank 1..4,
// optional relu / spatial softmax, frame mean, for X[B,HW,C] fp32 -> logits[B/F,K].
//
// Replaces (reference, TF-1.1 graph): nets_factory.py:259-274 attention conv(s), :276-286 softmax /
// relu, :299-309 top-down conv(s), :322-328 pool + rank sum, :350-352 squeeze, :354-374 frame mean.
//
// Math (exact re-association of the reference graph for M == 1):
//   a_0[hw] = X[hw,:].w_0 + c_0 ;  a_r = w_r a_{r-1} + c_r  (the reference CHAINS the maps, Q1)
//   A_r = relu / softmax_hw / identity (a_r)
//   g_r[c] = sum_hw A_r[hw] X[hw,c] ,  s_r = sum_hw A_r[hw]
//   logits[k] = 1/(HW F) sum_f sum_r ( g_r . B_r[:,k] + d_r[k] s_r )
// i.e. the [HW,K] top-down map is never formed and X is read from HBM exactly once.
//
// Structure (one launch):
//   phase 1  one thread-block CLUSTER of p CTAs per image; CTA q owns rows [q HW/p, (q+1) HW/p).
//            Warp 8 streams 4-row x C tiles HBM -> smem with cp.async.bulk (TMA engine, UBLKCP)
//            through a 4/6-stage mbarrier ring; 8 consumer warps each own C/8 channels: per tile they
//            form partial dots, one named barrier combines them into a[row], then they accumulate
//            g (8 channels per thread per rank, in registers).  At the end of an image every CTA
//            PUSHES its partial g (st.shared::cluster) into the slot buffers of the CTA that owns
//            those C/p channels, one cluster-scope mbarrier round publishes them, and the owner
//            sums the p slots (softmax max/sum merge included) and writes g once to a [B,R,C] workspace.
//   barrier  grid barrier (monotonic 64-bit arrival counter; all CTAs co-resident, cooperative launch
//            when available), split into arrive / wait so the first weight tile loads in between
//   phase 2  logits = g . B  as a register-tiled fp32 SIMT GEMM: cluster -> class slice, CTA in
//            cluster -> channel slice, partial sums combined through DSMEM (deterministic, no atomics).
#include "apa_common.cuh"
#include "apa_kernels.h"

#include <atomic>
#include <mutex>

namespace apa {

namespace {

constexpr int kTileRows = 4;   // rows per TMA stage == rows per consumer iteration (32 KB at C = 2048)

template <int Q, int RMAX>
struct Smem {
  static constexpr int C = Q * 1024;
  static constexpr int kStages = (Q == 1) ? 6 : 3;                            // 96 KB ring
  static constexpr size_t kTileBytes = size_t(kTileRows) * C * 4;
  static constexpr size_t kRing = kStages * kTileBytes;
  static constexpr size_t kOffSlots = kRing;                                  // float slots[p][R][C/p]
  static constexpr size_t kPhase1End = kOffSlots + size_t(RMAX) * C * 4;
  // phase-2 carve over the ring + slot buffer (both idle by then)
  static constexpr int kChunk = 512;                             // channels per smem chunk
  static constexpr int kChunkPad = kChunk + 4;
  static constexpr size_t kOffG = 0;                             // float gS[32][kChunkPad]
  static constexpr size_t kOffAcc = 0;                           // float accS[8][32][16] (aliases gS after use)
  static constexpr size_t kOffW = kOffG + size_t(kVideoTile) * kChunkPad * 4;  // float Ws[kChunk][16]
  static constexpr size_t kOffOut = kOffW + size_t(kChunk) * kColTile * 4;     // float outP[32][16]
  static constexpr size_t kPhase2End = kOffOut + size_t(kVideoTile) * kColTile * 4;
  static_assert(size_t(kConsumerWarps) * kVideoTile * kColTile * 4 <= kOffW, "accS must fit inside gS");
  static constexpr size_t kOffStat = (kPhase1End > kPhase2End) ? kPhase1End : kPhase2End;  // float stat[16][8]
  static constexpr size_t kOffPart = kOffStat + 16 * 8 * 4;                   // float part[2][4][8]
  static constexpr size_t kOffScl = kOffPart + 2 * kTileRows * 8 * 4;         // float scl[16]
  static constexpr size_t kOffBars = kOffScl + 64;                            // u64 full[8] empty[8] ready free
  static constexpr size_t kTotal = kOffBars + 8 * 18;
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
  // two CTAs per SM (two launches in flight overlap each other's latency-bound tail) when R == 1
  static constexpr int kMinCtasPerSm = (RMAX == 1 && kTotal <= 113 * 1024) ? 2 : 1;
};

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void axpy4(float* f, float a, const float4& x) {
  f[0] = fmaf(a, x.x, f[0]);
  f[1] = fmaf(a, x.y, f[1]);
  f[2] = fmaf(a, x.z, f[2]);
  f[3] = fmaf(a, x.w, f[3]);
}
// Debug timeline (ApaTuning.debug_timestamps): 64 x u64 per CTA in the workspace.
//  0 start | 1 after cluster sync | 2 phase-1 end | 3 after DSMEM combine | 4 after grid barrier | 5 end
//  6 grid-barrier arrive | 7 phase-2 operands in smem
//  8..31 consumer: tile i ready (full barrier passed) | 32..55 producer: copy i issued | 62,63 globaltimer
__device__ __forceinline__ void ts_mark(unsigned long long* ts, int slot) {
  if (ts != nullptr) ts[size_t(blockIdx.x) * 64 + slot] = clock64();
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Phase-2 operand loaders (global/L2 -> smem), loads batched in registers so they overlap.
template <int CHP>
__device__ __forceinline__ void load_g_tile(float* gS, const float* g, int v0, int Bv, int F, int R, int j,
                                            int C, int c_off, int CHc, int tid) {
  const int c4n = CHc >> 2;              // a power of two (C / cluster size, capped at kChunk)
  const int sh = 31 - __clz(c4n);
  const int total = kVideoTile * c4n;
  if (F == 1) {
    constexpr int U = 8;
    for (int base = tid; base < total; base += U * kFusedThreads) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = base + u * kFusedThreads;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < total) {
          const int row = idx >> sh, c4 = idx & (c4n - 1);
          if (v0 + row < Bv)
            v[u] = __ldcg(reinterpret_cast<const float4*>(g + (size_t(v0 + row) * R + j) * C + c_off + c4 * 4));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = base + u * kFusedThreads;
        if (idx < total) {
          const int row = idx >> sh, c4 = idx & (c4n - 1);
          *reinterpret_cast<float4*>(gS + row * CHP + c4 * 4) = v[u];
        }
      }
    }
  } else {
    for (int idx = tid; idx < total; idx += kFusedThreads) {
      const int row = idx >> sh, c4 = idx & (c4n - 1);
      const int v = v0 + row;
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < Bv) {
        const float* src = g + (size_t(v) * F * R + j) * C + c_off + c4 * 4;
        int fr = 0;
        for (; fr + 4 <= F; fr += 4) {
          float4 t[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) t[u] = __ldcg(reinterpret_cast<const float4*>(src + size_t(fr + u) * R * C));
#pragma unroll
          for (int u = 0; u < 4; ++u) { sum.x += t[u].x; sum.y += t[u].y; sum.z += t[u].z; sum.w += t[u].w; }
        }
        for (; fr < F; ++fr) {
          const float4 t = __ldcg(reinterpret_cast<const float4*>(src + size_t(fr) * R * C));
          sum.x += t.x; sum.y += t.y; sum.z += t.z; sum.w += t.w;
        }
      }
      *reinterpret_cast<float4*>(gS + row * CHP + c4 * 4) = sum;
    }
  }
}

__device__ __forceinline__ void load_w_tile(float* Ws, const float* Wj, int K, int c_off, int CHc, int kp,
                                            int k_hi, int tid) {
  const int total = CHc * kColTile;
  constexpr int U = 8;
  for (int base = tid; base < total; base += U * kFusedThreads) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * kFusedThreads;
      v[u] = 0.f;
      if (idx < total) {
        const int c = idx >> 4, k = kp + (idx & 15);
        if (k < k_hi) v[u] = __ldg(Wj + size_t(c_off + c) * K + k);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * kFusedThreads;
      if (idx < total) Ws[idx] = v[u];
    }
  }
}

template <int Q, int RMAX, int MODE>
__global__ void __launch_bounds__(kFusedThreads, Smem<Q, RMAX>::kMinCtasPerSm)
apa_pool_fused_kernel(const FusedParams p) {
  using L = Smem<Q, RMAX>;
  constexpr int C = L::C;
  constexpr int TR = kTileRows;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* slots = reinterpret_cast<float*>(smem + L::kOffSlots);
  float* statS = reinterpret_cast<float*>(smem + L::kOffStat);
  float* partS = reinterpret_cast<float*>(smem + L::kOffPart);
  float* scl = reinterpret_cast<float*>(smem + L::kOffScl);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBars);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 8;
  uint64_t* parts_ready = bars + 16;
  uint64_t* parts_free = bars + 17;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t csize = cluster_nctarank();
  const uint32_t crank = cluster_ctarank();
  const int cluster_id = blockIdx.x / csize;
  const int nclusters = gridDim.x / csize;
  const int R = p.R;
  const int Cp = C / int(csize);  // channels owned by one CTA of the cluster (combine + phase 2)
  const int row_lo = int((long long)p.HW * crank / csize);
  const int row_hi = int((long long)p.HW * (crank + 1) / csize);
  const int nrows = row_hi - row_lo;
  const int ntiles = (nrows + TR - 1) / TR;

  if (tid == 0) {
    ts_mark(p.ts, 0);
    if (p.ts) p.ts[size_t(blockIdx.x) * 64 + 62] = globaltimer_ns();
  }

  // =========================================================================== phase 1
  if (warp == kConsumerWarps) {
    // -------------------------------------------------- producer warp: owns the stage barriers
    if (lane == 0) {
      for (int i = 0; i < L::kStages; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], kConsumerWarps);
      }
      mbar_init(parts_ready, csize);
      mbar_init(parts_free, csize);
      fence_mbar_init();
    }
    __syncwarp();
    cluster_arrive();  // publishes the barrier inits; this warp's matching wait comes after its copy loop
    if (p.do_phase1 && lane == 0) {
      // Start streaming before the rest of the CTA has even synchronised: the first ring fill only
      // needs the barriers this lane just initialised.  One 32 KB bulk copy per 4-row tile (issuing a
      // bulk copy costs ~0.3 us of this thread's time, so big copies: scripts/ubench_copy.cu).
      const uint64_t pol = l2_policy_evict_first();
      uint32_t stage = 0, phase = 0;
      int issued = 0;
      for (int b = cluster_id; b < p.B; b += nclusters) {
        const float* xb = p.X + (size_t(b) * p.HW + row_lo) * C;
        for (int t = 0; t < ntiles; ++t) {
          const int rows = min(TR, nrows - t * TR);
          const uint32_t bytes = uint32_t(rows) * C * 4u;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          bulk_g2s(smem + stage * L::kTileBytes, xb + size_t(t) * TR * C, bytes, &full_bar[stage], pol);
          if (issued < 24) ts_mark(p.ts, 32 + issued);
          ++issued;
          if (++stage == L::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
    griddep_launch_dependents();
    cluster_wait();
  } else {
    cluster_arrive();
    cluster_wait();  // barrier inits of every CTA of the cluster are visible (remote arrives come later)
  }
  if (tid == 0) ts_mark(p.ts, 1);

  if (p.do_phase1 && warp < kConsumerWarps) {
    // -------------------------------------------------- consumers: warp w owns channels [w C/8, (w+1) C/8)
    const int col0 = warp * (C / kConsumerWarps) + lane * 4;  // + q*128

    // Warm L2 with this CTA's phase-2 weight slice while phase 1 streams X.
    if (p.do_phase2) {
      const int k_lo = int((long long)p.K * cluster_id / nclusters);
      const int k_hi = int((long long)p.K * (cluster_id + 1) / nclusters);
      if (k_hi > k_lo) {
        for (int j = 0; j < R; ++j)
          for (int c = tid; c < Cp; c += kConsumerThreads) {
            const float* row = p.td_w[j] + size_t(crank * Cp + c) * p.K;
            prefetch_l2(row + k_lo);
            prefetch_l2(row + k_hi - 1);
          }
      }
    }

    float4 wv[Q];
    float c0 = 0.f;
    const float c_in = (p.attn_in != nullptr && p.attn_in_bias) ? __ldg(p.attn_b[0]) : 0.f;
    if (p.attn_in == nullptr) {
#pragma unroll
      for (int q = 0; q < Q; ++q) wv[q] = __ldg(reinterpret_cast<const float4*>(p.attn_w[0] + col0 + q * 128));
      c0 = __ldg(p.attn_b[0]);
    }
    float cw[RMAX], cb[RMAX];  // chain coefficients a_r = cw[r] a_{r-1} + cb[r], r >= 1
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      cw[j] = (j >= 1 && j < R) ? __ldg(p.attn_w[j]) : 0.f;
      cb[j] = (j >= 1 && j < R) ? __ldg(p.attn_b[j]) : 0.f;
    }
    uint32_t stage = 0, phase = 0, giter = 0, img_iter = 0;
    for (int b = cluster_id; b < p.B; b += nclusters, ++img_iter) {
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
        const bool full = (rows == TR);  // CTA-uniform: fast path without per-row predicates
        // programmatic dependent launch: let the next launch become resident a few tiles before this CTA
        // stops streaming, so its prologue and first copies overlap the end of this stream
        if (b + nclusters >= p.B && it == max(0, ntiles - 1 - p.trigger_lead)) griddep_launch_dependents();
        const float* st = reinterpret_cast<const float*>(smem + stage * L::kTileBytes) + col0;
        mbar_wait(&full_bar[stage], phase);
        if (tid == 0 && giter < 24) ts_mark(p.ts, 8 + int(giter));
        float4 x[TR][Q];
        if (full) {
#pragma unroll
          for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int q = 0; q < Q; ++q) x[r][q] = *reinterpret_cast<const float4*>(st + r * C + q * 128);
        } else {
#pragma unroll
          for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int q = 0; q < Q; ++q)
              x[r][q] = (r < rows) ? *reinterpret_cast<const float4*>(st + r * C + q * 128)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float a[TR];
        if (p.attn_in == nullptr) {
          float pd[TR];
#pragma unroll
          for (int r = 0; r < TR; ++r) {
            pd[r] = dot4(x[r][0], wv[0]);
#pragma unroll
            for (int q = 1; q < Q; ++q) pd[r] += dot4(x[r][q], wv[q]);
          }
          // the tile is in registers now: hand the stage back to the copy engine
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
          // 4 partial dots -> one value per 8-lane group (transposed butterfly: 5 shuffles, not 20)
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
          kk += __shfl_xor_sync(0xffffffffu, kk, 1);  // lanes 8r..8r+7 hold row r's partial of this warp
          float* part = partS + (giter & 1u) * (TR * 8);
          if ((lane & 7) == 0) part[(lane >> 3) * 8 + warp] = kk;
          named_bar_sync(1, kConsumerThreads);
          // second level: lane l holds part[l] = (row l/8, warp l%8); sum the 8-lane groups, broadcast
          float v = part[lane];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
#pragma unroll
          for (int r = 0; r < TR; ++r) a[r] = c0 + __shfl_sync(0xffffffffu, v, 8 * r);
        } else {
#pragma unroll
          for (int r = 0; r < TR; ++r)
          {
            float v = 0.f;
            if (r < rows) {
              const size_t gi = size_t(b) * p.HW + row_lo + it * TR + r;
              v = c_in;
              for (int q = 0; q < p.attn_in_parts; ++q) v += __ldg(p.attn_in + size_t(q) * p.B * p.HW + gi);
            }
            a[r] = v;
          }
        }
        if (p.attn_raw != nullptr && tid == 0) {
#pragma unroll
          for (int r = 0; r < TR; ++r)
            if (r < rows) p.attn_raw[size_t(b) * p.HW + row_lo + it * TR + r] = a[r];
        }
        // validity as a multiplier keeps the accumulation branch-free (x of an absent row is zero)
        float valid[TR];
#pragma unroll
        for (int r = 0; r < TR; ++r) valid[r] = (full || r < rows) ? 1.f : 0.f;
        if (MODE == MODE_SOFTMAX) {
          float tmax = -INFINITY;
#pragma unroll
          for (int r = 0; r < TR; ++r) tmax = fmaxf(tmax, valid[r] != 0.f ? a[r] : -INFINITY);
          const float mnew = fmaxf(mx, tmax);
          const float sc = (mnew == -INFINITY) ? 1.f : expf(mx - mnew);  // exp(-inf) == 0 on the first tile
          mx = mnew;
          s[0] *= sc;
#pragma unroll
          for (int i = 0; i < Q * 4; ++i) f[0][i] *= sc;
#pragma unroll
          for (int r = 0; r < TR; ++r) {
            const float e = valid[r] != 0.f ? expf(a[r] - mx) : 0.f;
            s[0] += e;
#pragma unroll
            for (int q = 0; q < Q; ++q) axpy4(&f[0][q * 4], e, x[r][q]);
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
                for (int q = 0; q < Q; ++q) axpy4(&f[j][q * 4], A, x[r][q]);
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
      if (tid == 0 && img_iter == 0) ts_mark(p.ts, 2);
      // Programmatic dependent launch: once this CTA has streamed its last tile the NEXT launch may start
      // streaming (it co-resides, 2 CTAs/SM) while this one finishes its latency-bound tail; before the
      // first write to the (shared) workspace, make sure the PREVIOUS launch has completely finished.
      if (b + nclusters >= p.B && ntiles == 0) griddep_launch_dependents();
      if (img_iter == 0) griddep_wait();

      // ------------------------------------------------ combine the p row-slices of image b (DSMEM push)
      // Every CTA pushes its partial g into the slot buffer of the CTA that OWNS those C/p channels;
      // one cluster-scope mbarrier round (parts_ready) publishes them.  The slot buffer is single:
      // before pushing the next image a CTA waits on parts_free, which every owner arrives on once it
      // has consumed its slots (only when there IS a next image, so nothing remote touches a CTA that
      // has moved on).
      if (img_iter > 0) mbar_wait_cluster(parts_free, (img_iter - 1) & 1u);
      {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int c = col0 + q * 128;
          const uint32_t owner = uint32_t(c / Cp);
          const int cl = c - int(owner) * Cp;
#pragma unroll
          for (int j = 0; j < RMAX; ++j) {
            if (j < R) {
              const uint32_t ra = dsmem_addr(slots + (size_t(crank) * R + j) * Cp + cl, owner);
              st_dsmem_f4(ra, make_float4(f[j][q * 4 + 0], f[j][q * 4 + 1], f[j][q * 4 + 2], f[j][q * 4 + 3]));
            }
          }
        }
        if (tid < int(csize)) {  // my (s_r, max) to every peer's stat table
          const uint32_t ra = dsmem_addr(statS + size_t(crank) * 8, uint32_t(tid));
#pragma unroll
          for (int j = 0; j < RMAX; ++j) st_dsmem_f(ra + 4u * j, s[j]);
          st_dsmem_f(ra + 16u, mx);
        }
      }
      fence_acq_rel_cluster();
      named_bar_sync(1, kConsumerThreads);
      if (tid < int(csize)) mbar_arrive_cluster(parts_ready, tid);
      mbar_wait_cluster(parts_ready, img_iter & 1u);  // every CTA's partial has landed here

      if (MODE == MODE_SOFTMAX) {
        if (tid < 16) {
          float v = 0.f;
          if (tid < int(csize)) {
            float M = -INFINITY;
            for (uint32_t q = 0; q < csize; ++q) M = fmaxf(M, statS[q * 8 + 4]);
            float S = 0.f;
            for (uint32_t q = 0; q < csize; ++q) S += expf(statS[q * 8 + 4] - M) * statS[q * 8 + 0];
            v = expf(statS[tid * 8 + 4] - M) / S;
            if (tid == 0 && crank == 0) {
              p.smx[2 * b + 0] = M;
              p.smx[2 * b + 1] = S;
            }
          }
          scl[tid] = v;
        }
        named_bar_sync(1, kConsumerThreads);
      }
      {
        const int n4 = Cp / 4;
        for (int idx = tid; idx < R * n4; idx += kConsumerThreads) {
          const int j = idx / n4;
          const int cl = (idx - j * n4) * 4;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (uint32_t q = 0; q < csize; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(slots + (size_t(q) * R + j) * Cp + cl);
            const float sc = (MODE == MODE_SOFTMAX) ? scl[q] : 1.f;
            acc.x = fmaf(sc, v.x, acc.x);
            acc.y = fmaf(sc, v.y, acc.y);
            acc.z = fmaf(sc, v.z, acc.z);
            acc.w = fmaf(sc, v.w, acc.w);
          }
          *reinterpret_cast<float4*>(p.g + (size_t(b) * R + j) * C + int(crank) * Cp + cl) = acc;
        }
        if (crank == 0 && tid < R) {
          float S = 1.f;
          if (MODE != MODE_SOFTMAX) {
            S = 0.f;
            for (uint32_t q = 0; q < csize; ++q) S += statS[q * 8 + tid];
          }
          p.sn[size_t(b) * R + tid] = S;
        }
      }
      if (b + nclusters < p.B) {  // slots / stat table consumed: peers may push the next image
        named_bar_sync(1, kConsumerThreads);
        if (tid < int(csize)) mbar_arrive_cluster(parts_free, tid);
      }
    }
    if (tid == 0) ts_mark(p.ts, 3);
  }

  if (!p.do_phase1) {
    griddep_launch_dependents();
    griddep_wait();
  }
  if (!p.do_phase2) return;

  // =========================================================================== phase 2
  {
    float* gS = reinterpret_cast<float*>(smem + L::kOffG);
    float* Ws = reinterpret_cast<float*>(smem + L::kOffW);
    float* accS = reinterpret_cast<float*>(smem + L::kOffAcc);
    float* outP = reinterpret_cast<float*>(smem + L::kOffOut);
    constexpr int CHP = L::kChunkPad;
    const int c_base = int(crank) * Cp;
    const int CHc = min(L::kChunk, Cp);
    const int Bv = p.B / p.F;
    const int k_lo = int((long long)p.K * cluster_id / nclusters);
    const int k_hi = int((long long)p.K * (cluster_id + 1) / nclusters);
    const int ig = lane >> 2, cg = lane & 3;

    // grid barrier, split: arrive -> (load the first weight tile, which does not depend on phase 1) -> wait
    unsigned long long target = 0;
    bool w_preloaded = false;
    if (p.do_phase1) {
      __syncthreads();  // this CTA's g stores are done (and the stage ring is free)
      if (tid == 0) {
        target = grid_arrive(p.barrier, gridDim.x);
        ts_mark(p.ts, 6);
      }
      if (k_hi > k_lo) {
        load_w_tile(Ws, p.td_w[0], p.K, c_base, CHc, k_lo, k_hi, tid);
        w_preloaded = true;
      }
      if (tid == 0) grid_wait(p.barrier, target);
      __syncthreads();
    }
    if (tid == 0) ts_mark(p.ts, 4);

    for (int v0 = 0; v0 < Bv; v0 += kVideoTile) {
      for (int kp = k_lo; kp < k_hi; kp += kColTile) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int t = 0; t < 4; ++t) acc[i][t] = 0.f;

        for (int j = 0; j < R; ++j) {
          for (int cc = 0; cc < Cp; cc += CHc) {
            const bool first = (v0 == 0 && kp == k_lo && j == 0 && cc == 0);
            if (!first) __syncthreads();  // previous chunk fully consumed
            load_g_tile<CHP>(gS, p.g, v0, Bv, p.F, R, j, C, c_base + cc, CHc, tid);
            if (!(first && w_preloaded)) load_w_tile(Ws, p.td_w[j], p.K, c_base + cc, CHc, kp, k_hi, tid);
            __syncthreads();
            if (first && tid == 0) ts_mark(p.ts, 7);
            if (tid < kConsumerThreads) {
              const int cw = CHc / kConsumerWarps;
              const int c_begin = warp * cw;
              for (int c = c_begin; c < c_begin + cw; c += 4) {
                float4 gv[4], wv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) gv[i] = *reinterpret_cast<const float4*>(gS + (ig + 8 * i) * CHP + c);
#pragma unroll
                for (int t = 0; t < 4; ++t) wv[t] = *reinterpret_cast<const float4*>(Ws + (c + t) * kColTile + cg * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  acc[i][0] = fmaf(gv[i].x, wv[0].x, acc[i][0]);
                  acc[i][1] = fmaf(gv[i].x, wv[0].y, acc[i][1]);
                  acc[i][2] = fmaf(gv[i].x, wv[0].z, acc[i][2]);
                  acc[i][3] = fmaf(gv[i].x, wv[0].w, acc[i][3]);
                  acc[i][0] = fmaf(gv[i].y, wv[1].x, acc[i][0]);
                  acc[i][1] = fmaf(gv[i].y, wv[1].y, acc[i][1]);
                  acc[i][2] = fmaf(gv[i].y, wv[1].z, acc[i][2]);
                  acc[i][3] = fmaf(gv[i].y, wv[1].w, acc[i][3]);
                  acc[i][0] = fmaf(gv[i].z, wv[2].x, acc[i][0]);
                  acc[i][1] = fmaf(gv[i].z, wv[2].y, acc[i][1]);
                  acc[i][2] = fmaf(gv[i].z, wv[2].z, acc[i][2]);
                  acc[i][3] = fmaf(gv[i].z, wv[2].w, acc[i][3]);
                  acc[i][0] = fmaf(gv[i].w, wv[3].x, acc[i][0]);
                  acc[i][1] = fmaf(gv[i].w, wv[3].y, acc[i][1]);
                  acc[i][2] = fmaf(gv[i].w, wv[3].z, acc[i][2]);
                  acc[i][3] = fmaf(gv[i].w, wv[3].w, acc[i][3]);
                }
              }
            }
          }
        }
        // combine the 8 channel-split warps, then the p channel-split CTAs
        __syncthreads();  // accS aliases gS: every warp is done reading its operands
        if (tid < kConsumerThreads) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(accS + (warp * kVideoTile + ig + 8 * i) * kColTile + cg * 4) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        for (int o = tid; o < kVideoTile * kColTile; o += kFusedThreads) {
          float sum = 0.f;
#pragma unroll
          for (int w = 0; w < kConsumerWarps; ++w) sum += accS[w * kVideoTile * kColTile + o];
          outP[o] = sum;
        }
        cluster_sync_all();
        for (int o = int(crank) + int(csize) * tid; o < kVideoTile * kColTile; o += int(csize) * kFusedThreads) {
          const int v = v0 + (o >> 4);
          const int k = kp + (o & 15);
          if (v < Bv && k < k_hi) {
            float part[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) part[q] = (q < int(csize)) ? ld_dsmem_f_nv(dsmem_addr(outP + o, q)) : 0.f;
            float sum = 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) sum += part[q];
            for (int j = 0; j < R; ++j) {
              float S = 0.f;
              for (int fr = 0; fr < p.F; ++fr) S += __ldcg(p.sn + size_t(v * p.F + fr) * R + j);
              sum = fmaf(__ldg(p.td_b[j] + k), S, sum);
            }
            const float val = sum * p.inv_scale;
            if (p.gw_world > 0) {
              const size_t off = (peer_gather_step(p.gw_done) % p.gw_depth) * size_t(p.gw_world) * Bv * p.K + (size_t(p.gw_rank) * Bv + size_t(v)) * p.K + k;
              for (int qq = 0; qq < p.gw_world; ++qq) p.gw_logits[qq][off] = val;
            } else {
              p.logits[size_t(v) * p.K + k] = val;
            }
          }
        }
        cluster_sync_all();  // outP / accS are reused by the next pass; also guards exit
      }
    }
  }
  if (p.gw_world > 0 && tid == 0) peer_gather_finish(p.gw_world, p.gw_rank, p.gw_flags, p.gw_done, gridDim.x, peer_gather_step(p.gw_done));
  if (tid == 0) {
    ts_mark(p.ts, 5);
    if (p.ts) p.ts[size_t(blockIdx.x) * 64 + 63] = globaltimer_ns();
  }
}

// ------------------------------------------------------------------------------------- host side
using KernelFn = void (*)(const FusedParams);

template <int Q, int RMAX, int MODE>
KernelFn kfn() {
  return &apa_pool_fused_kernel<Q, RMAX, MODE>;
}

KernelFn select_kernel(int Q, int rmax, int mode) {
  if (Q == 2) {
    if (rmax == 1) {
      if (mode == MODE_PLAIN) return kfn<2, 1, MODE_PLAIN>();
      if (mode == MODE_RELU) return kfn<2, 1, MODE_RELU>();
      return kfn<2, 1, MODE_SOFTMAX>();
    }
    if (mode == MODE_PLAIN) return kfn<2, APA_MAX_RANK, MODE_PLAIN>();
    if (mode == MODE_RELU) return kfn<2, APA_MAX_RANK, MODE_RELU>();
    return nullptr;
  }
  if (Q == 1) {
    if (rmax == 1) {
      if (mode == MODE_PLAIN) return kfn<1, 1, MODE_PLAIN>();
      if (mode == MODE_RELU) return kfn<1, 1, MODE_RELU>();
      return kfn<1, 1, MODE_SOFTMAX>();
    }
    if (mode == MODE_PLAIN) return kfn<1, APA_MAX_RANK, MODE_PLAIN>();
    if (mode == MODE_RELU) return kfn<1, APA_MAX_RANK, MODE_RELU>();
    return nullptr;
  }
  return nullptr;
}

size_t smem_bytes(int Q, int rmax) {
  if (Q == 2) return rmax == 1 ? Smem<2, 1>::kTotal : Smem<2, APA_MAX_RANK>::kTotal;
  return rmax == 1 ? Smem<1, 1>::kTotal : Smem<1, APA_MAX_RANK>::kTotal;
}

// max co-resident clusters per (device, kernel, cluster size); filled lazily, immutable afterwards.
struct OccCache {
  std::mutex mu;
  struct Entry { int dev; KernelFn fn; int cluster; int nmax; };
  Entry e[64];
  int n = 0;
};
OccCache g_occ;

cudaError_t max_clusters(KernelFn fn, size_t smem, int cluster, int* out) {
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  {
    std::lock_guard<std::mutex> lk(g_occ.mu);
    for (int i = 0; i < g_occ.n; ++i)
      if (g_occ.e[i].dev == dev && g_occ.e[i].fn == fn && g_occ.e[i].cluster == cluster) {
        *out = g_occ.e[i].nmax;
        return cudaSuccess;
      }
  }
  err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (err != cudaSuccess) return err;
  // max shared-memory carve-out: the SM must be able to hold a CTA of the next launch beside this one
  (void)cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (cluster > 8) {
    err = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (err != cudaSuccess) return err;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster, 1, 1);
  cfg.blockDim = dim3(kFusedThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  err = cudaOccupancyMaxActiveClusters(&n, fn, &cfg);
  if (err != cudaSuccess) {
    (void)cudaGetLastError();
    n = 0;  // this cluster size is not launchable here
  }
  // The kernel fits two CTAs per SM so that the NEXT launch (programmatic dependent launch) can stream
  // beside this launch's tail; one launch therefore only ever takes one CTA slot per SM.
  std::lock_guard<std::mutex> lk(g_occ.mu);
  if (g_occ.n < 64) g_occ.e[g_occ.n++] = {dev, fn, cluster, n};
  *out = n;
  return cudaSuccess;
}

// nominal B200 capacity (148 SMs in 8 GPCs of 16/18/20) for device-less planning
int nominal_max_clusters(int cluster) {
  switch (cluster) {
    case 1: return 148;
    case 2: return 74;
    case 4: return 33;
    case 8: return 16;
    default: return 8;
  }
}

void choose(int B, int HW, int C, const ApaTuning* tuning, const int nmax[5], FusedPlan* plan) {
  static const int cand[5] = {1, 2, 4, 8, 16};
  int best = -1;
  long long best_cost = 0;
  const int forced = tuning ? tuning->cluster_size : 0;
  for (int i = 0; i < 5; ++i) {
    const int cs = cand[i];
    if (forced > 0 && cs != forced) continue;
    if (nmax[i] <= 0) continue;
    if ((C / cs) % 32 != 0 || C % cs != 0) continue;
    int ncl = nmax[i];
    if (tuning && tuning->max_ctas > 0) ncl = ncl < tuning->max_ctas / cs ? ncl : tuning->max_ctas / cs;
    if (ncl <= 0) continue;
    const long long waves = (B + ncl - 1) / ncl;
    const long long rows = (HW + cs - 1) / cs;
    // time ~ waves * (rows + fixed per-image cost in row units); ties -> smaller cluster
    const long long cost = waves * (rows + 3);
    if (best < 0 || cost < best_cost) {
      best = i;
      best_cost = cost;
    }
  }
  if (best < 0) {
    plan->cluster = 0;
    plan->nclusters = 0;
    return;
  }
  int ncl = nmax[best];
  if (tuning && tuning->max_ctas > 0 && ncl > tuning->max_ctas / cand[best]) ncl = tuning->max_ctas / cand[best];
  if (ncl > B) ncl = B;
  plan->cluster = cand[best];
  plan->nclusters = ncl;
}

void fill_common(int C, int R, int mode, const ApaTuning* tuning, FusedPlan* plan) {
  plan->Q = C / 1024;
  plan->rmax = (R == 1) ? 1 : APA_MAX_RANK;
  plan->mode = mode;
  plan->smem = smem_bytes(plan->Q, plan->rmax);
  plan->fused = (tuning && tuning->fused == 0) ? 0 : 1;
  plan->cooperative = 0;
  plan->pdl = (tuning && tuning->no_pdl) ? 0 : 1;
}

}  // namespace

bool fused_supported(int C, int R) {
  return (C == 1024 || C == 2048) && R >= 1 && R <= APA_MAX_RANK;
}

void fused_make_plan_nodevice(int B, int HW, int C, int K, int R, int mode, const ApaTuning* tuning,
                              FusedPlan* plan) {
  (void)K;
  fill_common(C, R, mode, tuning, plan);
  int nmax[5];
  static const int cand[5] = {1, 2, 4, 8, 16};
  for (int i = 0; i < 5; ++i) nmax[i] = nominal_max_clusters(cand[i]);
  choose(B, HW, C, tuning, nmax, plan);
}

cudaError_t fused_make_plan(int B, int HW, int C, int K, int R, int mode, const ApaTuning* tuning,
                            FusedPlan* plan) {
  (void)K;
  fill_common(C, R, mode, tuning, plan);
  KernelFn fn = select_kernel(plan->Q, plan->rmax, mode);
  if (fn == nullptr) return cudaErrorInvalidValue;
  int nmax[5];
  static const int cand[5] = {1, 2, 4, 8, 16};
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int i = 0; i < 5; ++i) {
    cudaError_t err = max_clusters(fn, plan->smem, cand[i], &nmax[i]);
    if (err != cudaSuccess) return err;
    // The kernel fits two CTAs per SM so that the NEXT launch (programmatic dependent launch) can stream
    // beside this launch's tail; one launch therefore only takes one CTA slot per SM (max_ctas < 0 lifts it).
    if (!(tuning && tuning->max_ctas < 0) && sms > 0 && nmax[i] > sms / cand[i]) nmax[i] = sms / cand[i];
  }
  choose(B, HW, C, tuning, nmax, plan);
  if (plan->cluster == 0) return cudaErrorInvalidConfiguration;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  plan->cooperative = (coop && !(tuning && tuning->no_cooperative)) ? 1 : 0;
  return cudaSuccess;
}

namespace {
std::atomic<int> g_coop_refused{0};  // set once if the driver refuses cooperative + cluster launches
std::atomic<int> g_pdl_refused{0};   // set once if the driver refuses programmatic dependent launch here

cudaError_t launch_one(KernelFn fn, const FusedParams& p, const FusedPlan& plan, bool coop,
                       cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.cluster * plan.nclusters, 1, 1);
  cfg.blockDim = dim3(kFusedThreads, 1, 1);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[3];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = plan.cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  int na = 1;
  if (coop) {
    at[na].id = cudaLaunchAttributeCooperative;
    at[na].val.cooperative = 1;
    ++na;
  }
  if (plan.pdl && !g_pdl_refused.load(std::memory_order_relaxed)) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  cudaError_t err = cudaLaunchKernelEx(&cfg, fn, p);
  if (err != cudaSuccess && !coop && plan.pdl && !g_pdl_refused.load(std::memory_order_relaxed)) {
    (void)cudaGetLastError();
    g_pdl_refused.store(1, std::memory_order_relaxed);
    cfg.numAttrs = 1;  // cluster dimension only
    err = cudaLaunchKernelEx(&cfg, fn, p);
  }
  return err;
}
}  // namespace

int fused_grid(const FusedPlan& plan) { return plan.cluster * plan.nclusters; }

cudaError_t fused_launch(const FusedParams& p, const FusedPlan& plan, cudaStream_t stream) {
  KernelFn fn = select_kernel(plan.Q, plan.rmax, plan.mode);
  if (fn == nullptr) return cudaErrorInvalidValue;
  if (plan.fused) {
    FusedParams q = p;
    q.do_phase1 = 1;
    q.do_phase2 = 1;
    if (plan.cooperative && !g_coop_refused.load(std::memory_order_relaxed)) {
      cudaError_t err = launch_one(fn, q, plan, true, stream);
      if (err == cudaSuccess) return err;
      (void)cudaGetLastError();  // cooperative+cluster refused: the grid is still sized to be co-resident
      g_coop_refused.store(1, std::memory_order_relaxed);
    }
    return launch_one(fn, q, plan, false, stream);
  }
  FusedParams q = p;
  q.do_phase1 = 1;
  q.do_phase2 = 0;
  cudaError_t err = launch_one(fn, q, plan, false, stream);
  if (err != cudaSuccess) return err;
  q.do_phase1 = 0;
  q.do_phase2 = 1;
  return launch_one(fn, q, plan, false, stream);
}

}  // namespace apa

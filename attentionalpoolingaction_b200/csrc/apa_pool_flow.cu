// apa_pool_flow.cu — the single-map head as a STREAM kernel that does nothing but stream, followed by two small
// kernels that wait in hardware.  Same math as the other forms (reference sites nets_factory.py:259-328,
// :350-352, :354-374):
//     g_r[b] = sum_hw A_r[b,hw] X[b,hw,:]      s_r[b] = sum_hw A_r[b,hw]
//     logits[v] = sum_frames sum_r (g_r . B_r + d_r s_r) / (HW F)
//
//   STREAM    G CTAs (G = 148: one per SM, and its 129 KB of shared memory keep it the only STREAM CTA on its
//             SM).  The launch's B*HW rows are cut into G equal runs that cross image boundaries, so every SM
//             streams the same byte count whatever the batch.  A producer lane feeds 3-row x C tiles (24 KB,
//             cp.async.bulk) through a 4-stage mbarrier ring; two groups of 4 consumer warps take alternate tiles
//             (two dot -> reduce -> accumulate chains in flight), form a[row] and accumulate g in registers.  A
//             group's PARTIAL g of an image (8 KB per rank) and (attention mass, running max) go to slot
//             2 (CTA + image) + group of the workspace: partials of non-final images wait in shared memory, the
//             CTA executes griddepcontrol.wait once, after its last tile, then stores.  No flags, atomics, fences.
//   REDUCE    one 128-thread CTA per (image, 512 channels): adds the image's partials in run order (softmax:
//             max/sum merge) into g[b], the attention mass into gsum[b].
//   CONTRACT  16 channel slices x ceil(K/32) class slices, 128 threads, 25 KB.  A unit loads its [C/16 x 32] block
//             of B_r into shared memory, executes griddepcontrol.wait, loads g of its channels for 32 videos (frames
//             summed, one round trip), contracts (2 x 4 register tiles, two channel halves, 16 videos at a time) and
//             writes its partial [32 x 32] block.
//   FINISH    one 128-thread CTA per (32 videos, class slice): adds the 16 partials in index order
//             (bit-reproducible), bias, scale, and stores the logits (or the peer stores of the fused all-gather).
// All four are programmatic dependents that trigger as soon as they know which call they belong to:
// STREAM(n) -> REDUCE(n) -> CONTRACT(n) -> FINISH(n) -> STREAM(n+1) -> ...  The three tail kernels are small enough to be
// resident (parked in griddepcontrol.wait, a hardware wait) beside STREAM(n) on every SM, so STREAM(n+1)'s CTA takes
// an SM the moment STREAM(n)'s CTA leaves it, on one stream as on four.  No kernel ever waits for a kernel it does
// not read from: every per-call buffer of the workspace (partials, g, partial logits, arrival counters) exists
// kFlowDepth = 2 times and call L uses copy L % 2.  L comes from ticket counters in the workspace (one atomic per
// CTA; every launch pads its counter to the next multiple of 2^20, so CUDA-graph replays and changing shapes need
// no host state), and a STREAM CTA of call L only checks (one acquire load, true in steady state) that call L - 2
// has completed before it stores its first partial.  Reading X before the previous kernel in the stream has
// completed is only legal when X was complete before that kernel started: ApaTuning.inputs_stable (default 0 =
// griddepcontrol.wait first, which serialises back-to-back calls).
#include "apa_common.cuh"
#include "apa_kernels.h"

#include <algorithm>
#include <mutex>

namespace apa {

namespace {

constexpr int kFTR = 3;        // rows per X tile
constexpr int kNJ = 16;        // channel slices of the contraction
constexpr int kKW = 32;        // classes per contract unit
constexpr int kGPad = 4;       // gs[video][C/16 + 4]: consecutive videos start 4 banks apart
constexpr int kCThreads = 128; // CONTRACT block
constexpr int kRThreads = 128; // REDUCE block: 128 float4 columns = 512 channels
constexpr unsigned long long kStride = 1ull << 20;
// control words (u64) of the flow form in the workspace control block
enum { kCtrTS = 0, kCtrTR = 1, kCtrTC = 2, kCtrDone = 3, kCtrGwStep = 4 /* [kFlowDepth] */, kCtrTF = 6 };

__device__ __forceinline__ float fdot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void faxpy4(float* f, float a, const float4& x) {
  f[0] = fmaf(a, x.x, f[0]);
  f[1] = fmaf(a, x.y, f[1]);
  f[2] = fmaf(a, x.z, f[2]);
  f[3] = fmaf(a, x.w, f[3]);
}
__device__ __forceinline__ unsigned long long atom_add_relaxed(unsigned long long* p, unsigned long long v) {
  unsigned long long prev;
  asm volatile("atom.add.relaxed.gpu.global.u64 %0, [%1], %2;" : "=l"(prev) : "l"(p), "l"(v) : "memory");
  return prev;
}
__device__ __forceinline__ unsigned long long atom_add_acq_rel(unsigned long long* p, unsigned long long v) {
  unsigned long long prev;
  asm volatile("atom.add.acq_rel.gpu.global.u64 %0, [%1], %2;" : "=l"(prev) : "l"(p), "l"(v) : "memory");
  return prev;
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// One thread per CTA: which call (launch index on this workspace) this launch belongs to.  `ctr` only ever grows;
// the CTA that takes a launch's last ticket pads it to the next multiple of 2^20.  Must complete before the CTA
// triggers its dependents (the next launch of the same kernel must not draw tickets before this one has all of its).
__device__ __forceinline__ unsigned long long flow_ticket(unsigned long long* ctr, unsigned int nctas) {
  const unsigned long long prev = atom_add_relaxed(ctr, 1ull);
  if ((prev & (kStride - 1)) == (unsigned long long)(nctas - 1)) (void)atom_add_relaxed(ctr, kStride - nctas);
  return prev >> 20;
}
__device__ __forceinline__ unsigned long long fglobaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// rows [flow_row_lo(t), flow_row_lo(t + 1)) of the flattened [B * HW] row space belong to STREAM CTA t
__device__ __forceinline__ int flow_row_lo(long long N, int G, int t) { return int(N * t / G); }
// the STREAM CTAs that hold rows of image b are [lo, hi]; CTA e keeps its partial of image b in slot e + b
__device__ __forceinline__ void flow_image_ctas(long long N, int G, int HW, int b, int* lo, int* hi) {
  const long long x0 = (long long)b * HW, x1 = x0 + HW - 1;
  long long a = ((x0 + 1) * G - 1) / N, c = ((x1 + 1) * G - 1) / N;
  *lo = int(a < G - 1 ? a : G - 1);
  *hi = int(c < G - 1 ? c : G - 1);
}

// Shapes of the STREAM kernel (ApaTuning.stream_shape picks one for A/B runs; 0 = default):
//   shape 0: two consumer groups, 4-stage ring (96 KB), <= 128 registers: one CTA per SM
//   shape 1: two consumer groups, 3-stage ring (72 KB), <= 96 registers: two CTAs per SM
//   shape 2: one consumer group,  3-stage ring (72 KB), <= 72 registers: three CTAs per SM
constexpr int kNG = 2;                        // partial-g slots per (CTA, image): sized for two consumer groups
__host__ __device__ inline size_t stream_smem_bytes(int C, int stages) {
  return size_t(stages) * kFTR * C * 4 + 512 + 16 * 8 + 64;
}

// ================================================================================================ STREAM
template <int Q, int RMAX, int MODE, int NG, int STG, int MAXREG>
__global__ void __launch_bounds__(kFusedThreads, 1) __maxnreg__((RMAX == 1) ? MAXREG : 216)
apa_flow_stream_kernel(const FusedParams p) {
  constexpr int kFStages = STG, kWPG = kConsumerWarps / NG;   // warps per group: a warp covers C / kWPG channels
  constexpr int C = Q * 1024, TR = kFTR, QT = Q * NG;    // QT float4s of a row per thread
  constexpr size_t kTileBytes = size_t(TR) * C * 4;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* partS = reinterpret_cast<float*>(smem + kFStages * kTileBytes);            // [group][2][32]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(partS + 128);
  uint64_t* empty_bar = full_bar + 8;
  unsigned long long* call_s = reinterpret_cast<unsigned long long*>(empty_bar + 8);
  static_assert(kFStages <= 8, "barrier arrays");
  const int R = p.R, G = p.fl_G, HW = p.HW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long N = (long long)p.B * HW;
  const int t = blockIdx.x;

  const bool wait_first = !p.inputs_stable || p.attn_in != nullptr || p.attn_raw != nullptr;
  if (wait_first) griddep_wait();

  const int row_lo = flow_row_lo(N, G, t), row_hi = flow_row_lo(N, G, t + 1);
  const int b_first = row_lo / HW, b_last = (row_hi - 1) / HW;
  if (tid == kConsumerThreads) {   // the producer lane sets up the ring and starts streaming at once
    for (int i = 0; i < kFStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kWPG);
    }
    fence_mbar_init();
  }
  __syncthreads();
  // =========================================================================== producer warp
  if (warp == kConsumerWarps) {
    if (lane == 0) {
      const uint64_t pol_x = l2_policy_evict_first();
      uint32_t stage = 0, phase = 0;
      for (int b = b_first; b <= b_last; ++b) {
        const int lo = max(row_lo, b * HW), hi = min(row_hi, (b + 1) * HW);
        const int nrows = hi - lo, ntiles = (nrows + TR - 1) / TR;
        const float* xb = p.X + size_t(lo) * C;
        for (int tt = 0; tt < ntiles; ++tt) {
          const int rows = min(TR, nrows - tt * TR);
          const uint32_t bytes = uint32_t(rows) * C * 4u;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          bulk_g2s(smem + stage * kTileBytes, xb + size_t(tt) * TR * C, bytes, &full_bar[stage], pol_x);
          if (++stage == kFStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;   // (the consumers trigger the dependents, after the CTA's ticket is drawn)
  }
  if (tid == 0) {
    // which call this is (copy L % 2 of the per-call buffers); the copy is free once call L - 2 has completed
    const unsigned long long L = flow_ticket(p.fl_ctr + kCtrTS, gridDim.x);
    if (t == 0 && p.gw_world > 0) {
      // the step of this gather call; everything before this call in the stream that is not one of our own
      // kernels has completed (it does not trigger early), so this rank no longer reads older gathered buffers
      const unsigned long long step = atom_add_relaxed(p.gw_done + 2, 1ull) + 1;
      peer_gather_publish_started(p.gw_world, p.gw_rank, p.gw_flags, step);
      p.fl_ctr[kCtrGwStep + (L & 1)] = step;   // read by CONTRACT(L) after its griddepcontrol.wait
    }
    if (L >= kFlowDepth && !p.fl_debug_nowait) {
      const long long t0 = clock64();
      while ((ld_acquire(p.fl_ctr + kCtrDone) >> 20) + kFlowDepth < L + 1) {
        __nanosleep(64);
        if (clock64() - t0 > APA_SPIN_TIMEOUT_CYCLES) __trap();
      }
    }
    *call_s = L;
    if (p.ts != nullptr && t < 160) {
      p.ts[size_t(t) * 64 + 0] = clock64();
      p.ts[size_t(t) * 64 + 62] = fglobaltimer_ns();
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.ts[size_t(t) * 64 + 61] = smid;
    }
  }
  named_bar_sync(3, kConsumerThreads);
  // When to let REDUCE / CONTRACT of this call (and, behind them, the next call's STREAM) be scheduled: they park in
  // griddepcontrol.wait until this grid has finished, and parked CTAs pile up on whichever SMs have room -- which
  // then have none for a STREAM CTA.  So by default a CTA triggers only `lead` tiles before its last one.
  const int lead = p.trigger_lead;
  if (lead < 0) griddep_launch_dependents();
  const size_t cp = size_t(*call_s & (kFlowDepth - 1));   // which copy of the per-call buffers
  float* const gpart = p.gpart + cp * p.fl_gpart_stride;
  float* const gstat = p.gstat + cp * p.fl_gstat_stride;

  // =========================================================================== consumers: two groups of 4 warps;
  // group g takes the CTA's tiles g, g + 2, ... and keeps its own partial sums (slot 2 (CTA + image) + g)
  const int grp = warp / kWPG, wig = warp % kWPG, gtid = tid - grp * (kWPG * 32);
  const int col0 = wig * (C / kWPG) + lane * 4;  // + qq*128
  float4 wv[QT];
  float c0 = 0.f;
  const float c_in = (p.attn_in != nullptr && p.attn_in_bias) ? __ldg(p.attn_b[0]) : 0.f;
  if (p.attn_in == nullptr) {
#pragma unroll
    for (int qq = 0; qq < QT; ++qq) wv[qq] = __ldg(reinterpret_cast<const float4*>(p.attn_w[0] + col0 + qq * 128));
    c0 = __ldg(p.attn_b[0]);
  }
  float cw[RMAX], cb[RMAX];
#pragma unroll
  for (int j = 0; j < RMAX; ++j) {
    cw[j] = (j >= 1 && j < R) ? __ldg(p.attn_w[j]) : 0.f;
    cb[j] = (j >= 1 && j < R) ? __ldg(p.attn_b[j]) : 0.f;
  }
  float* part0 = partS + grp * 64;
  uint32_t tile0 = 0, flip = 0;       // index (within the CTA) of the current segment's first tile
  const uint32_t cta_tiles = [&] {     // tiles of this CTA (all segments)
    uint32_t n = 0;
    for (int b = b_first; b <= b_last; ++b)
      n += uint32_t(min(row_hi, (b + 1) * HW) - max(row_lo, b * HW) + TR - 1) / TR;
    return n;
  }();
  const uint32_t trig_tile = (lead >= 0 && cta_tiles > uint32_t(lead) + 2u) ? cta_tiles - 2u - uint32_t(lead) : 0u;
  bool triggered = (lead < 0);
  for (int b = b_first; b <= b_last; ++b) {
    const int seg_lo = max(row_lo, b * HW);
    const int nrows = min(row_hi, (b + 1) * HW) - seg_lo;
    const uint32_t ntiles = uint32_t(nrows + TR - 1) / TR;
    float f[RMAX][QT * 4];
    float s[RMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      s[j] = 0.f;
#pragma unroll
      for (int i = 0; i < QT * 4; ++i) f[j][i] = 0.f;
    }
    // my first tile of this segment: the smallest i >= tile0 with i % 2 == grp
    for (uint32_t i = (NG == 1) ? tile0 : tile0 + ((tile0 ^ uint32_t(grp)) & 1u); i < tile0 + ntiles; i += NG, flip ^= 32u) {
      const uint32_t stage = i % kFStages, phase = (i / kFStages) & 1u;
      if (!triggered && i >= trig_tile) { griddep_launch_dependents(); triggered = true; }
      const int off = int(i - tile0) * TR;
      const int rows = min(TR, nrows - off);
      const int gi = seg_lo + off;
      const bool full = (rows == TR);
      const float* st = reinterpret_cast<const float*>(smem + stage * kTileBytes) + col0;
      mbar_wait(&full_bar[stage], phase);
      float4 x[TR][QT];
      if (full) {
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
          for (int qq = 0; qq < QT; ++qq) x[r][qq] = *reinterpret_cast<const float4*>(st + r * C + qq * 128);
      } else {
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
          for (int qq = 0; qq < QT; ++qq)
            x[r][qq] = (r < rows) ? *reinterpret_cast<const float4*>(st + r * C + qq * 128)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float a[TR];
      if (p.attn_in == nullptr) {
        float pd[TR];
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          pd[r] = fdot4(x[r][0], wv[0]);
#pragma unroll
          for (int qq = 1; qq < QT; ++qq) pd[r] += fdot4(x[r][qq], wv[qq]);
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
        float* part = part0 + flip;
        if ((lane & 7) == 0) part[(lane >> 3) * kWPG + wig] = kk;      // [row][warp of the group]
        if (grp == 0) named_bar_sync(1, kWPG * 32);
        else named_bar_sync(2, kWPG * 32);
        float v = (lane < TR * kWPG) ? part[lane] : 0.f;
        if (kWPG == 8) v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
#pragma unroll
        for (int r = 0; r < TR; ++r) a[r] = c0 + __shfl_sync(0xffffffffu, v, kWPG * r);
      } else {
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          float v = 0.f;
          if (r < rows) {
            v = c_in;
            for (int pp = 0; pp < p.attn_in_parts; ++pp)
              v += ldg_cg_f(p.attn_in + size_t(pp) * p.B * HW + size_t(gi + r));
          }
          a[r] = v;
        }
      }
      if (p.attn_raw != nullptr && gtid == 0) {
#pragma unroll
        for (int r = 0; r < TR; ++r)
          if (r < rows) p.attn_raw[size_t(gi + r)] = a[r];
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
        for (int i = 0; i < QT * 4; ++i) f[0][i] *= sc;
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          const float e = valid[r] != 0.f ? expf(a[r] - mx) : 0.f;
          s[0] += e;
#pragma unroll
          for (int qq = 0; qq < QT; ++qq) faxpy4(&f[0][qq * 4], e, x[r][qq]);
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
              for (int qq = 0; qq < QT; ++qq) faxpy4(&f[j][qq * 4], A, x[r][qq]);
            }
          }
        }
      }
      if (p.attn_in != nullptr) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
      }
    }
    tile0 += ntiles;
    // ---- this group's partial of image b
    const int slot = NG * (t + b) + grp;
#pragma unroll
    for (int j = 0; j < RMAX; ++j)
      if (j < R) {
        float* dst = gpart + (size_t(slot) * R + j) * C + col0;
#pragma unroll
        for (int qq = 0; qq < QT; ++qq)
          *reinterpret_cast<float4*>(dst + qq * 128) =
              make_float4(f[j][qq * 4 + 0], f[j][qq * 4 + 1], f[j][qq * 4 + 2], f[j][qq * 4 + 3]);
      }
    if (gtid == 0) {
#pragma unroll
      for (int j = 0; j < RMAX; ++j) gstat[size_t(slot) * 8 + j] = s[j];
      gstat[size_t(slot) * 8 + 4] = mx;
    }
  }
  if (!triggered) griddep_launch_dependents();
  if (tid == 0 && p.ts != nullptr && t < 160) {
    p.ts[size_t(t) * 64 + 5] = clock64();
    p.ts[size_t(t) * 64 + 63] = fglobaltimer_ns();
  }
}

// ================================================================================================ REDUCE
template <bool SOFTMAX>
__global__ void __launch_bounds__(kRThreads, 8)
apa_flow_reduce_kernel(const FusedParams p) {
  __shared__ float scl[320];   // per partial of this image: softmax scale exp(m_e - M) / S (1 otherwise)
  __shared__ unsigned long long call_s;
  const int R = p.R, G = p.fl_G, HW = p.HW, C = p.C;
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int c = (blockIdx.y * kRThreads + tid) * 4;
  const long long N = (long long)p.B * HW;
  if (tid == 0) call_s = flow_ticket(p.fl_ctr + kCtrTR, gridDim.x * gridDim.y);
  __syncthreads();
  griddep_launch_dependents();
  griddep_wait();   // STREAM of this call has finished and its partial g / stats are visible
  const size_t cp = size_t(call_s & (kFlowDepth - 1));
  const float* const gpart = p.gpart + cp * p.fl_gpart_stride;
  const float* const gstat = p.gstat + cp * p.fl_gstat_stride;
  float* const gout = p.g + cp * p.fl_g_stride;
  float* const gsum = p.gsum + cp * p.fl_gsum_stride;
  int lo, hi;
  flow_image_ctas(N, G, HW, b, &lo, &hi);
  const int n = p.fl_ng * (hi - lo + 1), s0 = p.fl_ng * (lo + b);   // partials of image b (runs x consumer groups), first slot
  // the image's statistics: lanes stride the partials (all loads of a lane in flight together), butterfly in a fixed
  // order; a serial loop here would cost one L2 round trip per partial
  const int lane = tid & 31, warp = tid >> 5;
  if (SOFTMAX) {
    float M = -INFINITY;
    for (int e = lane; e < n; e += 32) M = fmaxf(M, __ldcg(gstat + size_t(s0 + e) * 8 + 4));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    float S = 0.f;
    for (int e = lane; e < n; e += 32)
      S += expf(__ldcg(gstat + size_t(s0 + e) * 8 + 4) - M) * __ldcg(gstat + size_t(s0 + e) * 8);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
    for (int e = tid; e < n; e += kRThreads) scl[e] = expf(__ldcg(gstat + size_t(s0 + e) * 8 + 4) - M) / S;
    if (tid == 0 && blockIdx.y == 0) {
      gsum[size_t(b) * 4] = 1.f;
      if (p.smx != nullptr) { p.smx[2 * b] = M; p.smx[2 * b + 1] = S; }
    }
    __syncthreads();
  } else if (warp < 4 && blockIdx.y == 0) {
    float mass = 0.f;
    if (warp < R)
      for (int e = lane; e < n; e += 32) mass += __ldcg(gstat + size_t(s0 + e) * 8 + warp);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mass += __shfl_xor_sync(0xffffffffu, mass, o);
    if (lane == 0) gsum[size_t(b) * 4 + warp] = mass;
  }
  if (c < C) {
    for (int j = 0; j < R; ++j) {
      const float* gp = gpart + (size_t(s0) * R + j) * C + c;
      float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int e0 = 0; e0 < n; e0 += 8) {   // eight loads in flight, added in run order
        float4 gv[8];
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8)
          gv[q8] = (e0 + q8 < n) ? __ldcg(reinterpret_cast<const float4*>(gp + size_t(e0 + q8) * R * C))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) {
          const float sc = (SOFTMAX && e0 + q8 < n) ? scl[e0 + q8] : 1.f;
          a4.x = fmaf(sc, gv[q8].x, a4.x);
          a4.y = fmaf(sc, gv[q8].y, a4.y);
          a4.z = fmaf(sc, gv[q8].z, a4.z);
          a4.w = fmaf(sc, gv[q8].w, a4.w);
        }
      }
      *reinterpret_cast<float4*>(gout + (size_t(b) * R + j) * C + c) = a4;
    }
  }
}

// ================================================================================================ CONTRACT
constexpr int kVH = kVideoTile / 2;   // videos staged in shared memory at a time
__host__ __device__ inline size_t contract_smem_bytes(int C) {
  const int CH = C / kNJ;
  return size_t(CH) * kKW * 4 + size_t(kVH) * (CH + kGPad) * 4 + 64;
}

template <int Q>
__global__ void __launch_bounds__(kCThreads, 6)
apa_flow_contract_kernel(const FusedParams p) {
  constexpr int C = Q * 1024, CH = C / kNJ, GS = CH + kGPad;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* Ws = reinterpret_cast<float*>(smem);       // [CH][kKW]   block of B_r
  float* gs = Ws + CH * kKW;                        // [16 videos][GS]; later red[2][16][kKW]
  unsigned long long* call_s = reinterpret_cast<unsigned long long*>(gs + kVH * GS);
  const int R = p.R, K = p.K, F = p.F, NS = p.fl_NS;
  const int tid = threadIdx.x;
  const int u = blockIdx.x;            // unit: channel slice cj (fastest), class slice sl
  const int cj = u % kNJ, sl = u / kNJ;
  const int k_lo = int((long long)K * sl / NS), k_hi = int((long long)K * (sl + 1) / NS);
  const int width = k_hi - k_lo;       // <= kKW
  const int c_base = cj * CH;
  const int Bv = p.B / F;
  const int ntile = (Bv + kVideoTile - 1) / kVideoTile;
  const int ts_row = 160 + u;

  if (tid == 0) *call_s = flow_ticket(p.fl_ctr + kCtrTC, gridDim.x);
  __syncthreads();
  griddep_launch_dependents();  // FINISH of this call and the next call's STREAM may be scheduled
  if (!p.inputs_stable) griddep_wait();
  const size_t cp = size_t(*call_s & (kFlowDepth - 1));   // which copy of the per-call buffers
  const float* const gin = p.g + cp * p.fl_g_stride;
  float* const cpart = p.cpart + cp * p.fl_cpart_stride;
  if (tid == 0 && p.ts != nullptr && ts_row < 320) {
    p.ts[size_t(ts_row) * 64 + 0] = clock64();
    p.ts[size_t(ts_row) * 64 + 62] = fglobaltimer_ns();
  }
  auto load_w = [&](int j) {
    const float* W = p.td_w[j];
#pragma unroll 8
    for (int idx = tid; idx < CH * kKW; idx += kCThreads) {
      const int c = idx / kKW, kk = idx - c * kKW;
      Ws[idx] = (kk < width) ? __ldg(W + size_t(c_base + c) * K + k_lo + kk) : 0.f;
    }
  };
  load_w(0);        // does not depend on this call's STREAM / REDUCE
  griddep_wait();   // REDUCE (and so STREAM) of this call has finished: g[b] is visible
  if (tid == 0 && p.ts != nullptr && ts_row < 320) p.ts[size_t(ts_row) * 64 + 1] = clock64();

  // thread tile of the contraction: 2 videos x 4 classes of a 16-video half, one half of the unit's channels
  const int ks = tid >> 6;                       // channel half
  const int vr = (tid & 63) >> 3, cq = tid & 7;  // video pair, class group (4 classes)
  constexpr int NP = kVideoTile * (CH / 4) / kCThreads;   // float4s of g per thread, rank and tile (8, or 4 for C = 1024)
  constexpr int NH = NP / 2;                              // ... per 16-video half
  for (int tile = 0; tile < ntile; ++tile) {
    const int v0 = tile * kVideoTile;
    float acc[2][2][4];   // [half][video of the pair][class]
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) acc[h][i][m] = 0.f;
    for (int j = 0; j < R; ++j) {
      if (j != 0 || tile != 0) {
        __syncthreads();                      // previous operands fully consumed
        if (R > 1) load_w(j);                 // rank 1: the block loaded before the wait stays valid
      }
      // ---- operand g: this unit's channels of 32 videos, frames summed.  All NP loads of a frame are in flight
      //      together (one round trip); the two 16-video halves then pass through shared memory one after the other.
      float4 a4[NP];
#pragma unroll
      for (int h = 0; h < NP; ++h) a4[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int fr = 0; fr < F; ++fr) {
#pragma unroll
        for (int h = 0; h < NP; ++h) {
          const int idx = tid + h * kCThreads;            // half = h / NH, then (video in half, 4 channels)
          const int vi = idx / (CH / 4), c4 = idx - vi * (CH / 4);
          if (v0 + vi < Bv) {
            const float4 gv = __ldcg(reinterpret_cast<const float4*>(
                gin + ((size_t(v0 + vi) * F + fr) * R + j) * C + c_base + c4 * 4));
            a4[h].x += gv.x; a4[h].y += gv.y; a4[h].z += gv.z; a4[h].w += gv.w;
          }
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half != 0) __syncthreads();       // first half fully consumed
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int idx = tid + h * kCThreads;            // index within the half
          const int vi = idx / (CH / 4), c4 = idx - vi * (CH / 4);
          *reinterpret_cast<float4*>(gs + vi * GS + c4 * 4) = a4[half * NH + h];
        }
        __syncthreads();
        // ---- contraction: acc[half][i][m] += g[2 vr + i][c] * W[c][4 cq + m] over this thread's channel half
        const float* g0 = gs + (vr * 2) * GS + ks * (CH / 2);
        const float* w0 = Ws + (ks * (CH / 2)) * kKW + cq * 4;
#pragma unroll 2
        for (int c = 0; c < CH / 2; c += 4) {
          float4 gv[2], wv[4];
#pragma unroll
          for (int i = 0; i < 2; ++i) gv[i] = *reinterpret_cast<const float4*>(g0 + i * GS + c);
#pragma unroll
          for (int m = 0; m < 4; ++m) wv[m] = *reinterpret_cast<const float4*>(w0 + (c + m) * kKW);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            float* a = acc[half][i];
            a[0] = fmaf(gv[i].x, wv[0].x, a[0]); a[1] = fmaf(gv[i].x, wv[0].y, a[1]);
            a[2] = fmaf(gv[i].x, wv[0].z, a[2]); a[3] = fmaf(gv[i].x, wv[0].w, a[3]);
            a[0] = fmaf(gv[i].y, wv[1].x, a[0]); a[1] = fmaf(gv[i].y, wv[1].y, a[1]);
            a[2] = fmaf(gv[i].y, wv[1].z, a[2]); a[3] = fmaf(gv[i].y, wv[1].w, a[3]);
            a[0] = fmaf(gv[i].z, wv[2].x, a[0]); a[1] = fmaf(gv[i].z, wv[2].y, a[1]);
            a[2] = fmaf(gv[i].z, wv[2].z, a[2]); a[3] = fmaf(gv[i].z, wv[2].w, a[3]);
            a[0] = fmaf(gv[i].w, wv[3].x, a[0]); a[1] = fmaf(gv[i].w, wv[3].y, a[1]);
            a[2] = fmaf(gv[i].w, wv[3].z, a[2]); a[3] = fmaf(gv[i].w, wv[3].w, a[3]);
          }
        }
      }
      if (tid == 0 && tile == 0 && j == 0 && p.ts != nullptr && ts_row < 320) p.ts[size_t(ts_row) * 64 + 2] = clock64();
    }
    // ---- the two channel halves meet in shared memory (red aliases gs), fixed order; then this unit's partial
    //      [32 videos][32 classes] block goes to the workspace (FINISH adds the 16 channel slices)
    const int cslot = tile * NS + sl;
    float* cpo = cpart + (size_t(cslot) * kNJ + cj) * (kVideoTile * kKW);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      __syncthreads();
      float* red = gs;  // [2][16 videos][kKW]
#pragma unroll
      for (int i = 0; i < 2; ++i)
        *reinterpret_cast<float4*>(red + (ks * kVH + vr * 2 + i) * kKW + cq * 4) =
            make_float4(acc[half][i][0], acc[half][i][1], acc[half][i][2], acc[half][i][3]);
      __syncthreads();
      const int o = tid * 4;   // 16 x 32 outputs of the half = one float4 per thread
      float4 r0 = *reinterpret_cast<const float4*>(red + o);
      const float4 r1 = *reinterpret_cast<const float4*>(red + kVH * kKW + o);
      r0.x += r1.x; r0.y += r1.y; r0.z += r1.z; r0.w += r1.w;
      *reinterpret_cast<float4*>(cpo + half * (kVH * kKW) + o) = r0;
    }
  }
  if (tid == 0 && p.ts != nullptr && ts_row < 320) {
    p.ts[size_t(ts_row) * 64 + 3] = clock64();
    p.ts[size_t(ts_row) * 64 + 5] = clock64();
    p.ts[size_t(ts_row) * 64 + 63] = fglobaltimer_ns();
  }
}

// ================================================================================================ FINISH
// One 128-thread CTA per (video tile, class slice): adds the 16 channel-slice partials in index order
// (bit-reproducible), bias x attention mass, scale, and stores the logits (or the peer stores of the fused all-gather).
__global__ void __launch_bounds__(kCThreads, 8)
apa_flow_finish_kernel(const FusedParams p) {
  __shared__ float sS[kVideoTile * 4];
  __shared__ float outS[kVideoTile * kKW];   // the finished [32 videos][32 classes] block
  __shared__ unsigned long long call_s[2];   // [0] call, [1] gather step
  const int R = p.R, K = p.K, F = p.F, NS = p.fl_NS;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x / NS, sl = blockIdx.x - tile * NS;
  const int k_lo = int((long long)K * sl / NS), k_hi = int((long long)K * (sl + 1) / NS);
  const int width = k_hi - k_lo;
  const int Bv = p.B / F, v0 = tile * kVideoTile;
  if (tid == 0) call_s[0] = flow_ticket(p.fl_ctr + kCtrTF, gridDim.x);
  __syncthreads();
  griddep_launch_dependents();
  griddep_wait();   // CONTRACT of this call has finished: the partial blocks are visible
  const size_t cp = size_t(call_s[0] & (kFlowDepth - 1));
  const float* const gsum = p.gsum + cp * p.fl_gsum_stride;
  const float* const cb0 = p.cpart + cp * p.fl_cpart_stride + size_t(blockIdx.x) * kNJ * (kVideoTile * kKW);
  if (tid < kVideoTile * 4) {      // attention mass per (video, rank): sum over frames
    const int vi = tid >> 2, j = tid & 3, v = v0 + vi;
    float mass = 0.f;
    if (v < Bv && j < R)
      for (int fr = 0; fr < F; ++fr) mass += __ldcg(gsum + size_t(v * F + fr) * 4 + j);
    sS[tid] = mass;
  }
  size_t gw_off = 0;
  if (p.gw_world > 0) {  // this step's ring buffer must be free on every peer (they started step - depth + 1)
    if (tid == 0) {
      const unsigned long long step = __ldcg(p.fl_ctr + kCtrGwStep + cp);   // drawn by STREAM CTA 0 of this call
      if (p.gw_world > 1 && step >= (unsigned long long)p.gw_depth)
        peer_gather_wait_all(p.gw_world, p.gw_flags[p.gw_rank], p.gw_world, step - p.gw_depth + 1);
      call_s[1] = step;
    }
  }
  __syncthreads();
  if (p.gw_world > 0) gw_off = size_t(call_s[1] % (unsigned long long)p.gw_depth) * size_t(p.gw_world) * Bv * K;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int o = (tid + h * kCThreads) * 4;
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c0 = 0; c0 < kNJ; c0 += 8) {   // eight loads in flight at a time, summed in index order
      float4 pv[8];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        pv[c] = __ldcg(reinterpret_cast<const float4*>(cb0 + size_t(c0 + c) * (kVideoTile * kKW) + o));
#pragma unroll
      for (int c = 0; c < 8; ++c) { sum[0] += pv[c].x; sum[1] += pv[c].y; sum[2] += pv[c].z; sum[3] += pv[c].w; }
    }
    const int vi = o / kKW, kk0 = o % kKW;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int kk = kk0 + m;
      float val = sum[m];
      if (kk < width)
        for (int j = 0; j < R; ++j) val = fmaf(__ldg(p.td_b[j] + k_lo + kk), sS[vi * 4 + j], val);
      outS[o + m] = val * p.inv_scale;
    }
  }
  __syncthreads();
  // rows of `width` consecutive floats: consecutive lanes store consecutive addresses (whole sectors, which matters
  // most for the peer stores: NVLink delivers partial-sector writes as read-modify-writes on the receiving GPU)
  const int nvid = min(kVideoTile, Bv - v0);
  for (int idx = tid; idx < nvid * width; idx += kCThreads) {
    const int vi = idx / width, kk = idx - vi * width;
    const float val = outS[vi * kKW + kk];
    const size_t row = size_t(v0 + vi) * K + k_lo + kk;
    if (p.gw_world > 0) {
      const size_t off = gw_off + size_t(p.gw_rank) * Bv * K + row;
      for (int qq = 0; qq < p.gw_world; ++qq) p.gw_logits[qq][off] = val;
    } else {
      p.logits[row] = val;
    }
  }
  if (p.gw_world > 1) __threadfence_system();  // my peer stores are visible system-wide before I am counted
  __syncthreads();
  if (tid == 0) {
    const unsigned long long prev2 = atom_add_acq_rel(p.fl_ctr + kCtrDone, 1ull);
    if ((prev2 & (kStride - 1)) == (unsigned long long)(gridDim.x - 1)) {
      // the call's last logits are stored: the gather step is complete, and so is the call (pad the counter to the
      // next multiple of 2^20: its upper bits count the completed calls, which frees copy L % 2 for call L + 2)
      if (p.gw_world > 0) peer_gather_complete(p.gw_world, p.gw_rank, p.gw_flags, p.gw_done, call_s[1]);
      (void)atom_add_acq_rel(p.fl_ctr + kCtrDone, kStride - (unsigned long long)gridDim.x);
    }
  }
}

using FKernelFn = void (*)(const FusedParams);

template <int NG, int STG, int MAXREG>
FKernelFn flow_stream_select_shape(int Q, int rmax, int mode) {
  if (Q == 2) {
    if (rmax == 1) {
      if (mode == MODE_PLAIN) return &apa_flow_stream_kernel<2, 1, MODE_PLAIN, NG, STG, MAXREG>;
      if (mode == MODE_RELU) return &apa_flow_stream_kernel<2, 1, MODE_RELU, NG, STG, MAXREG>;
      return &apa_flow_stream_kernel<2, 1, MODE_SOFTMAX, NG, STG, MAXREG>;
    }
    if (mode == MODE_PLAIN) return &apa_flow_stream_kernel<2, APA_MAX_RANK, MODE_PLAIN, NG, STG, MAXREG>;
    if (mode == MODE_RELU) return &apa_flow_stream_kernel<2, APA_MAX_RANK, MODE_RELU, NG, STG, MAXREG>;
    return nullptr;
  }
  if (Q == 1) {
    if (rmax == 1) {
      if (mode == MODE_PLAIN) return &apa_flow_stream_kernel<1, 1, MODE_PLAIN, NG, STG, MAXREG>;
      if (mode == MODE_RELU) return &apa_flow_stream_kernel<1, 1, MODE_RELU, NG, STG, MAXREG>;
      return &apa_flow_stream_kernel<1, 1, MODE_SOFTMAX, NG, STG, MAXREG>;
    }
    if (mode == MODE_PLAIN) return &apa_flow_stream_kernel<1, APA_MAX_RANK, MODE_PLAIN, NG, STG, MAXREG>;
    if (mode == MODE_RELU) return &apa_flow_stream_kernel<1, APA_MAX_RANK, MODE_RELU, NG, STG, MAXREG>;
  }
  return nullptr;
}
FKernelFn flow_stream_select(int Q, int rmax, int mode, int shape, int* stages, int* ng) {
  if (shape == 1) { *stages = 3; *ng = 2; return flow_stream_select_shape<2, 3, 96>(Q, rmax, mode); }
  if (shape == 2) { *stages = 3; *ng = 1; return flow_stream_select_shape<1, 3, 72>(Q, rmax, mode); }
  *stages = 4; *ng = 2;
  return flow_stream_select_shape<2, 4, 128>(Q, rmax, mode);
}
FKernelFn flow_contract_select(int Q) {
  if (Q == 2) return &apa_flow_contract_kernel<2>;
  if (Q == 1) return &apa_flow_contract_kernel<1>;
  return nullptr;
}

std::mutex g_fmu;
struct FEntry { int dev; const void* fn; };
FEntry g_fcache[64];
int g_fn = 0;

cudaError_t flow_prepare(const void* fn, size_t smem) {
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  {
    std::lock_guard<std::mutex> lk(g_fmu);
    for (int i = 0; i < g_fn; ++i)
      if (g_fcache[i].dev == dev && g_fcache[i].fn == fn) return cudaSuccess;
  }
  err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (err != cudaSuccess) return err;
  (void)cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  std::lock_guard<std::mutex> lk(g_fmu);
  if (g_fn < 64) g_fcache[g_fn++] = {dev, fn};
  return cudaSuccess;
}

}  // namespace

int flow_class_slices(int K) { return (K + kKW - 1) / kKW; }
int flow_contract_grid(int K) { return kNJ * flow_class_slices(K); }

int flow_stream_ctas(int B, int HW, const ApaTuning* tuning) {
  const long long n = (long long)B * HW;
  int g = kFlowStreamCtas;
  if (tuning && tuning->stream_ctas > 0 && tuning->stream_ctas < g) g = tuning->stream_ctas;
  // small batches: at least 16 rows (128 KB) per CTA, so that an image is cut into few partials (REDUCE adds them
  // one by one) and a CTA's stream is more than the first-tile latency
  if (n / 16 < g) g = int(n / 16 > 0 ? n / 16 : 1);
  return g;
}

bool flow_eligible(int B, int HW, int C, int K, int R, int F) {
  if (!((C == 1024 || C == 2048) && R >= 1 && R <= APA_MAX_RANK && F >= 1 && B % F == 0 && HW >= 1)) return false;
  return (long long)flow_slices(B, K, F) < (long long)kStride / 2 && flow_contract_grid(K) <= 65535;
}

size_t flow_gpart_slots(int B) { return kNG * (size_t(kFlowStreamCtas) + size_t(B)); }
size_t flow_slices(int B, int K, int F) {  // (video tile, class slice) pairs = arrival counters
  const int ntile = (B / F + kVideoTile - 1) / kVideoTile;
  return size_t(ntile) * flow_class_slices(K);
}
size_t flow_cpart_floats(int B, int K, int F) { return flow_slices(B, K, F) * kNJ * kVideoTile * kKW; }
size_t flow_smem(int C, const ApaTuning* tuning) {
  return stream_smem_bytes(C, (tuning && tuning->stream_shape != 0) ? 3 : 4);
}

cudaError_t flow_launch(const FusedParams& p_in, int mode, const ApaTuning* tuning, cudaStream_t stream,
                        int* stream_grid, int* contract_grid) {
  static_assert(kVideoTile * kKW == 8 * kCThreads, "two float4s of the partial block per thread");
  static_assert(kVH * kKW == 4 * kCThreads, "one float4 of a 16-video half per thread");
  FusedParams p = p_in;
  const int Q = p.C / 1024, rmax = (p.R == 1) ? 1 : APA_MAX_RANK;
  int stages = 4, ng = 2;
  const int shape = tuning ? tuning->stream_shape : 0;
  FKernelFn sfn = flow_stream_select(Q, rmax, mode, shape, &stages, &ng), cfn = flow_contract_select(Q);
  p.fl_ng = ng;
  FKernelFn rfn = (mode == MODE_SOFTMAX) ? &apa_flow_reduce_kernel<true> : &apa_flow_reduce_kernel<false>;
  if (sfn == nullptr || cfn == nullptr) return cudaErrorInvalidValue;
  const size_t pad = (tuning && tuning->tail_smem_kb > 0) ? size_t(tuning->tail_smem_kb) * 1024 : 0;
  const size_t smem1 = stream_smem_bytes(p.C, stages), smem2 = std::max(contract_smem_bytes(p.C), pad);
  cudaError_t err = flow_prepare(reinterpret_cast<const void*>(sfn), smem1);
  if (err != cudaSuccess) return err;
  err = flow_prepare(reinterpret_cast<const void*>(cfn), size_t(200) * 1024);   // room for any tail_smem_kb
  if (err != cudaSuccess) return err;
  err = flow_prepare(reinterpret_cast<const void*>(rfn), size_t(200) * 1024);
  if (err != cudaSuccess) return err;
  err = flow_prepare(reinterpret_cast<const void*>(&apa_flow_finish_kernel), size_t(200) * 1024);
  if (err != cudaSuccess) return err;
  p.fl_G = flow_stream_ctas(p.B, p.HW, tuning);
  {
    const size_t slots = flow_gpart_slots(p.B);
    p.fl_gpart_stride = slots * p.R * p.C;
    p.fl_gstat_stride = slots * 8;
    p.fl_g_stride = size_t(p.B) * p.R * p.C;
    p.fl_gsum_stride = size_t(p.B) * 4;
    p.fl_cpart_stride = flow_cpart_floats(p.B, p.K, p.F);
  }
  p.fl_NS = flow_class_slices(p.K);
  p.inputs_stable = (tuning && tuning->inputs_stable) ? 1 : 0;
  p.fl_debug_nowait = (tuning && tuning->debug_cut_tail > 0) ? 1 : 0;
  const int grid2 = kNJ * p.fl_NS;
  if (stream_grid) *stream_grid = p.fl_G;
  if (contract_grid) *contract_grid = grid2;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  const int na = (tuning && tuning->no_pdl) ? 0 : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.fl_G, 1, 1);
  cfg.blockDim = dim3(kFusedThreads, 1, 1);
  cfg.dynamicSmemBytes = smem1;
  cfg.stream = stream;
  cfg.attrs = at;
  cfg.numAttrs = na;
  err = cudaLaunchKernelEx(&cfg, sfn, p);
  if (err != cudaSuccess) return err;
  const int dbg = tuning ? tuning->debug_cut_tail : 0;   // measurement only: launch just the first 1 / 2 / 3 kernels
  if (dbg == 1) return cudaSuccess;
  cudaLaunchConfig_t cfgr = {};
  cfgr.gridDim = dim3(p.B, (p.C / 4 + kRThreads - 1) / kRThreads, 1);
  cfgr.blockDim = dim3(kRThreads, 1, 1);
  cfgr.dynamicSmemBytes = pad;
  cfgr.stream = stream;
  cfgr.attrs = at;
  cfgr.numAttrs = na;
  err = cudaLaunchKernelEx(&cfgr, rfn, p);
  if (err != cudaSuccess) return err;
  if (dbg == 2) return cudaSuccess;
  cudaLaunchConfig_t cfg2 = {};
  cfg2.gridDim = dim3(grid2, 1, 1);
  cfg2.blockDim = dim3(kCThreads, 1, 1);
  cfg2.dynamicSmemBytes = smem2;
  cfg2.stream = stream;
  cfg2.attrs = at;
  cfg2.numAttrs = na;
  err = cudaLaunchKernelEx(&cfg2, cfn, p);
  if (err != cudaSuccess) return err;
  if (dbg == 3) return cudaSuccess;
  cudaLaunchConfig_t cfg3 = {};
  cfg3.gridDim = dim3(unsigned(flow_slices(p.B, p.K, p.F)), 1, 1);
  cfg3.blockDim = dim3(kCThreads, 1, 1);
  cfg3.dynamicSmemBytes = pad;
  cfg3.stream = stream;
  cfg3.attrs = at;
  cfg3.numAttrs = na;
  return cudaLaunchKernelEx(&cfg3, &apa_flow_finish_kernel, p);
}

}  // namespace apa

// apa_pool_flow.cu — the single-map head as ONE launch of two kinds of CTA that never hold each other up.
// Same math as the other forms (reference sites nets_factory.py:259-328, :350-352, :354-374):
//     g_r[b] = sum_hw A_r[b,hw] X[b,hw,:]      s_r[b] = sum_hw A_r[b,hw]
//     logits[v] = sum_frames sum_r (g_r . B_r + d_r s_r) / (HW F)
//
//   blocks [0, G)       STREAM CTAs.  The launch's B*HW rows are cut into G equal runs that cross image
//                       boundaries (G = 148: every SM streams the same byte count).  A producer lane feeds
//                       3-row x C tiles (24 KB, cp.async.bulk) through a 3-stage mbarrier ring, 8 consumer warps
//                       form a[row] and accumulate g in registers.  At every image boundary the CTA stores its
//                       PARTIAL g (8 KB per rank) + (attention mass, running max) and counts itself in that
//                       image's arrival counter; the CTA that completes an image adds the image's partials in
//                       fixed order (softmax: max/sum merge) into g[b].  An image that lies inside one run is
//                       written to g[b] directly.  At the end the CTA counts itself in `stream-done` and exits.
//                       74 KB / 72 registers: three CTAs per SM.
//   blocks [G, G + U)   CONTRACT CTAs, U = 8 channel slices x ceil(K/32) class slices.  A unit loads its
//                       [C/8 x 32] block of B_r into shared memory, spins (one thread, nanosleep) until
//                       `stream-done` reaches G, loads g of its channels for 32 videos (summing frames), contracts
//                       (4 x 4 register tiles, 4 channel quarters), and writes its partial [32 x 32] block.  The 8th
//                       unit to arrive for a class slice adds the 8 partials in fixed order (bit-reproducible),
//                       bias, scale, and stores the logits (or the peer stores of the fused all-gather).
// A CONTRACT CTA waits only for STREAM CTAs of its own launch, which have lower block indices (CTAs are
// dispatched in index order, so they are resident or done) and which themselves wait for nothing but the
// previous launch; every spin has a trap timeout.  All counters live in the workspace and only ever grow: each
// launch leaves each of them at the next multiple of 2^20, so CUDA-graph replays and changing shapes need no host
// state and no reset (CONTRACT CTAs learn the launch index from a ticket counter).
//
// Programmatic dependent launch: every CTA triggers at once, so launch n+1's STREAM CTAs take the free CTA
// slots while launch n is still streaming / contracting; a STREAM CTA executes griddepcontrol.wait before its
// first global write (launch n has then completed, so neither the workspace nor the counters are ever shared by
// two launches).  Reading X before that wait is only legal when X was complete before the previous kernel in
// the stream started: ApaTuning.inputs_stable (default 0 = wait first).
#include "apa_common.cuh"
#include "apa_kernels.h"

#include <mutex>

namespace apa {

namespace {

constexpr int kFTR = 3;        // rows per X tile
constexpr int kFStages = 3;    // ring depth
constexpr int kNJ = 8;         // channel slices of the contraction
constexpr int kKW = 32;        // classes per contract unit
constexpr int kGPad = 4;       // gs[video][C/8 + 4]: consecutive videos start 4 banks apart
constexpr unsigned long long kStride = 1ull << 20;  // every launch leaves every counter at a multiple of this

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
__device__ __forceinline__ unsigned long long fglobaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// rows [flow_row_lo(t), flow_row_lo(t + 1)) of the flattened [B * HW] row space belong to STREAM CTA t
__device__ __forceinline__ int flow_row_lo(long long N, int G, int t) { return int(N * t / G); }
// the STREAM CTAs that hold rows of image b are [lo, hi]
__device__ __forceinline__ void flow_image_ctas(long long N, int G, int HW, int b, int* lo, int* hi) {
  const long long x0 = (long long)b * HW, x1 = x0 + HW - 1;
  long long a = ((x0 + 1) * G - 1) / N, c = ((x1 + 1) * G - 1) / N;
  *lo = int(a < G - 1 ? a : G - 1);
  *hi = int(c < G - 1 ? c : G - 1);
}

__host__ __device__ constexpr size_t flow_main_bytes(int C) {
  const size_t ring = size_t(kFStages) * kFTR * C * 4;
  const size_t contract = size_t(C / kNJ) * kKW * 4 + size_t(kVideoTile) * (C / kNJ + kGPad) * 4 + kVideoTile * 4 * 4;
  return ((ring > contract ? ring : contract) + 127) / 128 * 128;
}
__host__ __device__ inline size_t flow_smem_bytes(int C) { return flow_main_bytes(C) + 2 * 32 * 4 + 16 * 8 + 64; }

template <int Q, int RMAX, int MODE>
__global__ void __launch_bounds__(kFusedThreads, (RMAX == 1) ? 3 : 1)
apa_flow_kernel(const FusedParams p) {
  constexpr int C = Q * 1024, TR = kFTR, CH = C / kNJ;
  constexpr size_t kTileBytes = size_t(TR) * C * 4;
  constexpr size_t kMainBytes = flow_main_bytes(C);
  extern __shared__ __align__(1024) uint8_t smem[];
  float* partS = reinterpret_cast<float*>(smem + kMainBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kMainBytes + 256);
  uint64_t* empty_bar = full_bar + 4;
  unsigned long long* tick_s = reinterpret_cast<unsigned long long*>(full_bar + 8);
  int* flag_s = reinterpret_cast<int*>(full_bar + 10);
  const int R = p.R, G = p.fl_G, K = p.K, F = p.F, HW = p.HW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long N = (long long)p.B * HW;
  unsigned long long* ctr = p.fl_ctr;  // [0] ticket (CONTRACT CTAs)  [1] stream-done  [2] slices-done
  const int t = blockIdx.x;

  griddep_launch_dependents();  // the next launch may take free CTA slots from now on
  if (!p.inputs_stable || p.attn_in != nullptr || p.attn_raw != nullptr) griddep_wait();

  if (t < G) {
    // ========================================================================================= STREAM
    const int row_lo = flow_row_lo(N, G, t), row_hi = flow_row_lo(N, G, t + 1);
    const int b_first = row_lo / HW, b_last = (row_hi - 1) / HW;
    if (tid == 0) {
      for (int i = 0; i < kFStages; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], kConsumerWarps);
      }
      fence_mbar_init();
    }
    __syncthreads();
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
      return;
    }
    if (tid == 0 && p.ts != nullptr && t < 320) {
      p.ts[size_t(t) * 64 + 0] = clock64();
      p.ts[size_t(t) * 64 + 62] = fglobaltimer_ns();
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.ts[size_t(t) * 64 + 61] = smid;
    }
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
    uint32_t stage = 0, phase = 0, flip = 0;
    bool waited = false;
    for (int b = b_first; b <= b_last; ++b) {
      int gi = max(row_lo, b * HW);                       // running global row index of this image segment
      int remaining = min(row_hi, (b + 1) * HW) - gi;     // rows left in it
      float f[RMAX][Q * 4];
      float s[RMAX];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        s[j] = 0.f;
#pragma unroll
        for (int i = 0; i < Q * 4; ++i) f[j][i] = 0.f;
      }
      for (; remaining > 0; remaining -= TR, gi += TR, flip ^= 32u) {
        const int rows = min(TR, remaining);
        const bool full = (rows == TR);
        const float* st = reinterpret_cast<const float*>(smem + stage * kTileBytes) + col0;
        mbar_wait(&full_bar[stage], phase);
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
            pd[r] = fdot4(x[r][0], wv[0]);
#pragma unroll
            for (int qq = 1; qq < Q; ++qq) pd[r] += fdot4(x[r][qq], wv[qq]);
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
          float* part = partS + flip;
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
              v = c_in;
              for (int pp = 0; pp < p.attn_in_parts; ++pp)
                v += ldg_cg_f(p.attn_in + size_t(pp) * p.B * HW + size_t(gi + r));
            }
            a[r] = v;
          }
        }
        if (p.attn_raw != nullptr && tid == 0) {
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
          for (int i = 0; i < Q * 4; ++i) f[0][i] *= sc;
#pragma unroll
          for (int r = 0; r < TR; ++r) {
            const float e = valid[r] != 0.f ? expf(a[r] - mx) : 0.f;
            s[0] += e;
#pragma unroll
            for (int qq = 0; qq < Q; ++qq) faxpy4(&f[0][qq * 4], e, x[r][qq]);
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
                for (int qq = 0; qq < Q; ++qq) faxpy4(&f[j][qq * 4], A, x[r][qq]);
              }
            }
          }
        }
        if (p.attn_in != nullptr) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
        }
        if (++stage == kFStages) { stage = 0; phase ^= 1u; }
      }
      // the previous launch (whose CONTRACT CTAs read this workspace) must have completely finished
      if (!waited) {
        if (tid == 0 && p.ts != nullptr && t < 320) p.ts[size_t(t) * 64 + 2] = clock64();
        griddep_wait();
        waited = true;
        // everything before this call in the stream is complete: this rank no longer reads older gathered buffers
        if (t == 0 && tid == 0 && p.gw_world > 0)
          peer_gather_publish_started(p.gw_world, p.gw_rank, p.gw_flags, peer_gather_step(p.gw_done));
      }
      int e_lo, e_hi;                                  // STREAM CTAs holding rows of image b
      flow_image_ctas(N, G, HW, b, &e_lo, &e_hi);
      const bool whole = (e_lo == e_hi);               // the image lies inside this CTA's run: g[b] directly
      if (whole && MODE == MODE_SOFTMAX) {
        const float inv = 1.f / s[0];
#pragma unroll
        for (int i = 0; i < Q * 4; ++i) f[0][i] *= inv;
      }
      {
        float* base_g = whole ? p.g + size_t(b) * R * C : p.gpart + size_t(t + b) * R * C;
#pragma unroll
        for (int j = 0; j < RMAX; ++j)
          if (j < R) {
            float* dst = base_g + size_t(j) * C + col0;
#pragma unroll
            for (int qq = 0; qq < Q; ++qq)
              *reinterpret_cast<float4*>(dst + qq * 128) =
                  make_float4(f[j][qq * 4 + 0], f[j][qq * 4 + 1], f[j][qq * 4 + 2], f[j][qq * 4 + 3]);
          }
      }
      if (tid == 0) {
        if (whole) {
#pragma unroll
          for (int j = 0; j < RMAX; ++j) p.gsum[size_t(b) * 4 + j] = (MODE == MODE_SOFTMAX) ? 1.f : s[j];
          if (MODE == MODE_SOFTMAX && p.smx != nullptr) { p.smx[2 * b] = mx; p.smx[2 * b + 1] = s[0]; }
        } else {
#pragma unroll
          for (int j = 0; j < RMAX; ++j) p.gstat[size_t(t + b) * 8 + j] = s[j];
          p.gstat[size_t(t + b) * 8 + 4] = mx;
        }
      }
      if (!whole) {
        // count this CTA in image b; the CTA that completes the image reduces its partials into g[b]
        named_bar_sync(1, kConsumerThreads);
        if (tid == 0) {
          const unsigned long long cnt = (unsigned long long)(e_hi - e_lo + 1);
          const unsigned long long prev = atom_add_acq_rel(p.fl_img + b, 1ull);
          const bool last = ((prev & (kStride - 1)) == cnt - 1);
          if (last) (void)atom_add_relaxed(p.fl_img + b, kStride - cnt);
          *flag_s = last ? 1 : 0;
        }
        named_bar_sync(1, kConsumerThreads);
        if (*flag_s) {
          float M = 0.f, S = 1.f;
          if (MODE == MODE_SOFTMAX) {
            M = -INFINITY;
            for (int e = e_lo; e <= e_hi; ++e) M = fmaxf(M, __ldcg(p.gstat + size_t(e + b) * 8 + 4));
            S = 0.f;
            for (int e = e_lo; e <= e_hi; ++e)
              S += expf(__ldcg(p.gstat + size_t(e + b) * 8 + 4) - M) * __ldcg(p.gstat + size_t(e + b) * 8);
          }
#pragma unroll
          for (int j = 0; j < RMAX; ++j)
            if (j < R) {
              float4 acc4[Q];
#pragma unroll
              for (int qq = 0; qq < Q; ++qq) acc4[qq] = make_float4(0.f, 0.f, 0.f, 0.f);
              const float* gp = p.gpart + (size_t(e_lo + b) * R + j) * C + col0;
              const int nsl = e_hi - e_lo + 1;
              for (int e0 = 0; e0 < nsl; e0 += 4) {
                float4 gv[4][Q];
                float sc[4];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  const bool ok = (e0 + q4 < nsl);
#pragma unroll
                  for (int qq = 0; qq < Q; ++qq)
                    gv[q4][qq] = ok ? __ldcg(reinterpret_cast<const float4*>(gp + size_t(e0 + q4) * R * C + qq * 128))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                  sc[q4] = 1.f;
                  if (MODE == MODE_SOFTMAX && ok)
                    sc[q4] = expf(__ldcg(p.gstat + size_t(e_lo + b + e0 + q4) * 8 + 4) - M) / S;
                }
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
                  for (int qq = 0; qq < Q; ++qq) {
                    acc4[qq].x = fmaf(sc[q4], gv[q4][qq].x, acc4[qq].x);
                    acc4[qq].y = fmaf(sc[q4], gv[q4][qq].y, acc4[qq].y);
                    acc4[qq].z = fmaf(sc[q4], gv[q4][qq].z, acc4[qq].z);
                    acc4[qq].w = fmaf(sc[q4], gv[q4][qq].w, acc4[qq].w);
                  }
              }
              float* dst = p.g + (size_t(b) * R + j) * C + col0;
#pragma unroll
              for (int qq = 0; qq < Q; ++qq) *reinterpret_cast<float4*>(dst + qq * 128) = acc4[qq];
            }
          if (tid < RMAX) {
            float mass = 1.f;
            if (MODE != MODE_SOFTMAX) {
              mass = 0.f;
              if (tid < R)
                for (int e = e_lo; e <= e_hi; ++e) mass += __ldcg(p.gstat + size_t(e + b) * 8 + tid);
            }
            p.gsum[size_t(b) * 4 + tid] = mass;
          }
          if (MODE == MODE_SOFTMAX && p.smx != nullptr && tid == 0) { p.smx[2 * b] = M; p.smx[2 * b + 1] = S; }
        }
      }
    }
    named_bar_sync(1, kConsumerThreads);  // every consumer's stores (partials, completed images) are issued
    if (tid == 0) {
      (void)atom_add_acq_rel(ctr + 1, 1ull);  // release: they are visible to whoever sees the count
      if (p.ts != nullptr && t < 320) {
        p.ts[size_t(t) * 64 + 5] = clock64();
        p.ts[size_t(t) * 64 + 63] = fglobaltimer_ns();
      }
    }
    return;
  }

  // =========================================================================================== CONTRACT
  if (warp == kConsumerWarps) return;  // 256 threads do the contraction
  const int NS = p.fl_NS, U = kNJ * NS;
  if (tid == 0) {  // ticket: which launch this is (base) and which unit I am
    const unsigned long long prev = atom_add_relaxed(ctr + 0, 1ull);
    if ((prev & (kStride - 1)) == (unsigned long long)(U - 1))  // last ticket: pad to the next multiple
      (void)atom_add_relaxed(ctr + 0, kStride - (unsigned long long)U);
    *tick_s = prev;
  }
  named_bar_sync(1, kConsumerThreads);
  const unsigned long long base = *tick_s & ~(kStride - 1);   // = launch index * 2^20
  const int u = int(*tick_s & (kStride - 1));                 // unit: channel slice cj (fastest), class slice sl
  const int cj = u % kNJ, sl = u / kNJ;
  const int k_lo = int((long long)K * sl / NS), k_hi = int((long long)K * (sl + 1) / NS);
  const int width = k_hi - k_lo;       // <= kKW
  const int c_base = cj * CH;
  const int Bv = p.B / F;
  const int ntile = (Bv + kVideoTile - 1) / kVideoTile;
  constexpr int GS = CH + kGPad;
  float* Ws = reinterpret_cast<float*>(smem);       // [CH][kKW]
  float* gs = Ws + CH * kKW;                        // [32 videos][GS]; later red[4][32][kKW]
  float* sS = gs + kVideoTile * GS;                 // [32 videos][4] attention mass per rank
  const int ts_row = 160 + u;
  if (tid == 0 && p.ts != nullptr && ts_row < 320) {
    p.ts[size_t(ts_row) * 64 + 0] = clock64();
    p.ts[size_t(ts_row) * 64 + 62] = fglobaltimer_ns();
  }
  auto load_w = [&](int j) {
    const float* W = p.td_w[j];
    for (int idx = tid; idx < CH * kKW; idx += kConsumerThreads) {
      const int c = idx / kKW, kk = idx - c * kKW;
      Ws[idx] = (kk < width) ? __ldg(W + size_t(c_base + c) * K + k_lo + kk) : 0.f;
    }
  };
  load_w(0);  // does not depend on the STREAM CTAs
  if (tid == 0) {  // wait until all G STREAM CTAs of THIS launch have published their images
    const unsigned long long target = base + (unsigned long long)G;
    const long long t0 = clock64();
    while (ld_acquire(ctr + 1) < target) {
      __nanosleep(32);
      if (clock64() - t0 > APA_SPIN_TIMEOUT_CYCLES) __trap();
    }
  }
  named_bar_sync(1, kConsumerThreads);
  if (tid == 0 && p.ts != nullptr && ts_row < 320) p.ts[size_t(ts_row) * 64 + 1] = clock64();
  griddep_wait();  // previous launch complete (it is: this launch's STREAM CTAs waited for it) before any global write

  // thread tile of the contraction: 4 videos x 4 classes, one quarter of the unit's channels
  const int ks = warp >> 1;                    // channel quarter
  const int vr = (tid & 63) >> 3, cq = tid & 7;  // video group (4 videos), class group (4 classes)
  for (int tile = 0; tile < ntile; ++tile) {
    const int v0 = tile * kVideoTile;
    if (tile != 0) named_bar_sync(1, kConsumerThreads);  // sS / red of the previous tile fully consumed
    if (tid < kVideoTile * 4) {  // attention mass per (video, rank): sum over frames
      const int vi = tid >> 2, j = tid & 3, v = v0 + vi;
      float mass = 0.f;
      if (v < Bv && j < R)
        for (int fr = 0; fr < F; ++fr) mass += __ldcg(p.gsum + size_t(v * F + fr) * 4 + j);
      sS[tid] = mass;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int m = 0; m < 4; ++m) acc[i][m] = 0.f;
    for (int j = 0; j < R; ++j) {
      if (j != 0 || tile != 0) {
        named_bar_sync(1, kConsumerThreads);  // previous operands fully consumed
        if (R > 1) load_w(j);                 // rank 1: the block loaded before the wait stays valid (red aliases gs only)
      }
      // ---- operand g: this unit's channels of 32 videos, frames summed.  thread = (video, 4 channels), CH / 32 of them,
      //      four loads in flight at a time
      constexpr int NP = CH / 32;
#pragma unroll
      for (int h = 0; h < NP; h += 4) {
        float4 a4[4];
        const float* gp[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int idx = tid + (h + q4) * kConsumerThreads;
          const int vi = idx / (CH / 4), c4 = idx - vi * (CH / 4);
          a4[q4] = make_float4(0.f, 0.f, 0.f, 0.f);
          gp[q4] = (v0 + vi < Bv) ? p.g + (size_t(v0 + vi) * F * R + j) * C + c_base + c4 * 4 : nullptr;
        }
        for (int fr = 0; fr < F; ++fr) {
          float4 gv[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            gv[q4] = (gp[q4] != nullptr) ? __ldcg(reinterpret_cast<const float4*>(gp[q4] + size_t(fr) * R * C))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            a4[q4].x += gv[q4].x; a4[q4].y += gv[q4].y; a4[q4].z += gv[q4].z; a4[q4].w += gv[q4].w;
          }
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int idx = tid + (h + q4) * kConsumerThreads;
          const int vi = idx / (CH / 4), c4 = idx - vi * (CH / 4);
          *reinterpret_cast<float4*>(gs + vi * GS + c4 * 4) = a4[q4];
        }
      }
      named_bar_sync(1, kConsumerThreads);
      if (tid == 0 && tile == 0 && j == 0 && p.ts != nullptr && ts_row < 320) p.ts[size_t(ts_row) * 64 + 2] = clock64();
      // ---- contraction: acc[i][m] += g[4 vr + i][c] * W[c][4 cq + m] over this thread's channel quarter
      {
        const float* g0 = gs + (vr * 4) * GS + ks * (CH / 4);
        const float* w0 = Ws + (ks * (CH / 4)) * kKW + cq * 4;
#pragma unroll 2
        for (int c = 0; c < CH / 4; c += 4) {
          float4 gv[4], wv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) gv[i] = *reinterpret_cast<const float4*>(g0 + i * GS + c);
#pragma unroll
          for (int m = 0; m < 4; ++m) wv[m] = *reinterpret_cast<const float4*>(w0 + (c + m) * kKW);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(gv[i].x, wv[0].x, acc[i][0]); acc[i][1] = fmaf(gv[i].x, wv[0].y, acc[i][1]);
            acc[i][2] = fmaf(gv[i].x, wv[0].z, acc[i][2]); acc[i][3] = fmaf(gv[i].x, wv[0].w, acc[i][3]);
            acc[i][0] = fmaf(gv[i].y, wv[1].x, acc[i][0]); acc[i][1] = fmaf(gv[i].y, wv[1].y, acc[i][1]);
            acc[i][2] = fmaf(gv[i].y, wv[1].z, acc[i][2]); acc[i][3] = fmaf(gv[i].y, wv[1].w, acc[i][3]);
            acc[i][0] = fmaf(gv[i].z, wv[2].x, acc[i][0]); acc[i][1] = fmaf(gv[i].z, wv[2].y, acc[i][1]);
            acc[i][2] = fmaf(gv[i].z, wv[2].z, acc[i][2]); acc[i][3] = fmaf(gv[i].z, wv[2].w, acc[i][3]);
            acc[i][0] = fmaf(gv[i].w, wv[3].x, acc[i][0]); acc[i][1] = fmaf(gv[i].w, wv[3].y, acc[i][1]);
            acc[i][2] = fmaf(gv[i].w, wv[3].z, acc[i][2]); acc[i][3] = fmaf(gv[i].w, wv[3].w, acc[i][3]);
          }
        }
      }
    }
    // ---- the four channel quarters meet in shared memory (red aliases gs), fixed order
    named_bar_sync(1, kConsumerThreads);
    float* red = gs;  // [4][32 videos][kKW]
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(red + (ks * kVideoTile + vr * 4 + i) * kKW + cq * 4) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    named_bar_sync(1, kConsumerThreads);
    const int cslot = tile * NS + sl;
    {
      float4 r0 = *reinterpret_cast<const float4*>(red + tid * 4);
#pragma unroll
      for (int q4 = 1; q4 < 4; ++q4) {
        const float4 rq = *reinterpret_cast<const float4*>(red + q4 * kVideoTile * kKW + tid * 4);
        r0.x += rq.x; r0.y += rq.y; r0.z += rq.z; r0.w += rq.w;
      }
      *reinterpret_cast<float4*>(p.cpart + (size_t(cslot) * kNJ + cj) * (kVideoTile * kKW) + tid * 4) = r0;
    }
    named_bar_sync(1, kConsumerThreads);
    if (tid == 0) {
      const unsigned long long prev = atom_add_acq_rel(p.fl_arr + cslot, 1ull);
      *flag_s = ((prev + 1) % kNJ == 0) ? 1 : 0;
      if (tile == 0 && p.ts != nullptr && ts_row < 320) p.ts[size_t(ts_row) * 64 + 3] = clock64();
    }
    named_bar_sync(1, kConsumerThreads);
    if (*flag_s) {
      // ---- 8th arrival for this (tile, class slice): fixed-order sum of the 8 partial blocks, bias, logits
      const float* cb0 = p.cpart + size_t(cslot) * kNJ * (kVideoTile * kKW) + tid * 4;
      float4 pv[kNJ];
#pragma unroll
      for (int c = 0; c < kNJ; ++c) pv[c] = __ldcg(reinterpret_cast<const float4*>(cb0 + size_t(c) * (kVideoTile * kKW)));
      float sum[4] = {pv[0].x, pv[0].y, pv[0].z, pv[0].w};
#pragma unroll
      for (int c = 1; c < kNJ; ++c) { sum[0] += pv[c].x; sum[1] += pv[c].y; sum[2] += pv[c].z; sum[3] += pv[c].w; }
      const int vi = (tid * 4) / kKW, kk0 = (tid * 4) % kKW;
      const int v = v0 + vi;
      size_t gw_off = 0;
      if (p.gw_world > 0) {  // this step's ring buffer must be free on every peer (they started step - depth + 1)
        if (tid == 0) {
          const unsigned long long step = peer_gather_step(p.gw_done);
          if (p.gw_world > 1 && step >= (unsigned long long)p.gw_depth)
            peer_gather_wait_all(p.gw_world, p.gw_flags[p.gw_rank], p.gw_world, step - p.gw_depth + 1);
          *tick_s = step;
        }
        named_bar_sync(1, kConsumerThreads);
        gw_off = size_t(*tick_s % (unsigned long long)p.gw_depth) * size_t(p.gw_world) * Bv * K;
      }
      if (v < Bv) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int kk = kk0 + m, k = k_lo + kk;
          if (kk < width) {
            float val = sum[m];
            for (int j = 0; j < R; ++j) val = fmaf(__ldg(p.td_b[j] + k), sS[vi * 4 + j], val);
            val *= p.inv_scale;
            if (p.gw_world > 0) {
              const size_t off = gw_off + (size_t(p.gw_rank) * Bv + size_t(v)) * K + k;
              for (int qq = 0; qq < p.gw_world; ++qq) p.gw_logits[qq][off] = val;
            } else {
              p.logits[size_t(v) * K + k] = val;
            }
          }
        }
      }
      if (p.gw_world > 1) __threadfence_system();  // my peer stores are visible system-wide before I am counted
      named_bar_sync(1, kConsumerThreads);
      if (tid == 0) {
        const unsigned long long prev2 = atom_add_acq_rel(ctr + 2, 1ull);
        const unsigned long long nslices = (unsigned long long)ntile * NS;
        if ((prev2 & (kStride - 1)) == nslices - 1) {
          // the launch's last logits are stored: pad the two per-launch counters to the next multiple
          (void)atom_add_relaxed(ctr + 2, kStride - nslices);
          (void)atom_add_acq_rel(ctr + 1, kStride - (unsigned long long)G);
          if (p.gw_world > 0) peer_gather_complete(p.gw_world, p.gw_rank, p.gw_flags, p.gw_done, *tick_s);
        }
      }
    }
  }
  if (tid == 0 && p.ts != nullptr && ts_row < 320) {
    p.ts[size_t(ts_row) * 64 + 5] = clock64();
    p.ts[size_t(ts_row) * 64 + 63] = fglobaltimer_ns();
  }
}

using FKernelFn = void (*)(const FusedParams);

FKernelFn flow_select(int Q, int rmax, int mode) {
  if (Q == 2) {
    if (rmax == 1) {
      if (mode == MODE_PLAIN) return &apa_flow_kernel<2, 1, MODE_PLAIN>;
      if (mode == MODE_RELU) return &apa_flow_kernel<2, 1, MODE_RELU>;
      return &apa_flow_kernel<2, 1, MODE_SOFTMAX>;
    }
    if (mode == MODE_PLAIN) return &apa_flow_kernel<2, APA_MAX_RANK, MODE_PLAIN>;
    if (mode == MODE_RELU) return &apa_flow_kernel<2, APA_MAX_RANK, MODE_RELU>;
    return nullptr;
  }
  if (Q == 1) {
    if (rmax == 1) {
      if (mode == MODE_PLAIN) return &apa_flow_kernel<1, 1, MODE_PLAIN>;
      if (mode == MODE_RELU) return &apa_flow_kernel<1, 1, MODE_RELU>;
      return &apa_flow_kernel<1, 1, MODE_SOFTMAX>;
    }
    if (mode == MODE_PLAIN) return &apa_flow_kernel<1, APA_MAX_RANK, MODE_PLAIN>;
    if (mode == MODE_RELU) return &apa_flow_kernel<1, APA_MAX_RANK, MODE_RELU>;
  }
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

int flow_stream_ctas(int B, int HW, const ApaTuning* tuning) {
  long long n = (long long)B * HW;
  int g = kFlowStreamCtas;
  if (tuning && tuning->max_ctas > 0 && tuning->max_ctas < g) g = tuning->max_ctas;
  if (n < g) g = int(n);
  return g;
}

bool flow_eligible(int B, int HW, int C, int K, int R, int F) {
  if (!((C == 1024 || C == 2048) && R >= 1 && R <= APA_MAX_RANK && F >= 1 && B % F == 0 && HW >= 1)) return false;
  const long long slices = (long long)flow_slices(B, K, F);
  return slices < (long long)kStride / 2 && (long long)kNJ * flow_class_slices(K) < 60000;
}

size_t flow_gpart_slots(int B) { return size_t(kFlowStreamCtas) + size_t(B); }
size_t flow_slices(int B, int K, int F) {  // (video tile, class slice) pairs = arrival counters
  const int ntile = (B / F + kVideoTile - 1) / kVideoTile;
  return size_t(ntile) * flow_class_slices(K);
}
size_t flow_cpart_floats(int B, int K, int F) { return flow_slices(B, K, F) * kNJ * kVideoTile * kKW; }
size_t flow_smem(int C) { return flow_smem_bytes(C); }

cudaError_t flow_launch(const FusedParams& p_in, int mode, const ApaTuning* tuning, cudaStream_t stream, int* grid_out) {
  FusedParams p = p_in;
  const int Q = p.C / 1024, rmax = (p.R == 1) ? 1 : APA_MAX_RANK;
  FKernelFn fn = flow_select(Q, rmax, mode);
  if (fn == nullptr) return cudaErrorInvalidValue;
  const size_t smem = flow_smem_bytes(p.C);
  cudaError_t err = flow_prepare(reinterpret_cast<const void*>(fn), smem);
  if (err != cudaSuccess) return err;
  p.fl_G = flow_stream_ctas(p.B, p.HW, tuning);
  p.fl_NS = flow_class_slices(p.K);
  p.inputs_stable = (tuning && tuning->inputs_stable) ? 1 : 0;
  const int grid = p.fl_G + kNJ * p.fl_NS;
  static_assert(kVideoTile * kKW == 4 * kConsumerThreads, "one float4 of the partial block per thread");
  if (grid_out) *grid_out = grid;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(kFusedThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (tuning && tuning->no_pdl) ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, fn, p);
}

}  // namespace apa

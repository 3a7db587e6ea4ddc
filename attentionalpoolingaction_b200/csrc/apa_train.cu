// apa_train.cu — the head in TRAINING mode: forward with slim.dropout on the top-down branch and the backward.
//
// Reference sites: nets_factory.py:143-146 (dropout arg-scope: keep_prob 0.2 when cfg.NET.DROPOUT < 0, config.py:231),
// :296 (`slim.dropout(last_conv)` feeds the top-down convs only; the attention convs read the un-dropped map),
// :259-286 (attention convs, softmax / relu), :299-328 (top-down convs, multiply, spatial mean, rank sum), :354-374
// (frame mean); the callers are train.py:399-420 (clone_fn builds the graph with is_training=True) and :509-513
// (optimizer.compute_gradients differentiates it).  TF differentiates the graph automatically, so there is no
// reference source for the backward; the algebra is the re-associated closed form stated and tested on the CPU in
// oracle/attn_pool_backward.py (backward_reassociated), for the single-map (PER_CLASS off), SINGLE_LAYER_ATT head:
//
//   Xd = X * mask                     (mask = 0 or 1/keep; a counter-based hash of (seed, element index))
//   a_0 = X.w_0 + c_0;  a_r = w_r a_{r-1} + c_r;  A_r = act(a_r)                       [batch, HW]
//   g_r[v] = sum_{frames, hw} A_r Xd      s_r[v] = sum A_r                             [videos, C], [videos]
//   logits[v] = (sum_r g_r.B_r + d_r s_r) / (HW F)
//   G[v] = dlogits[v] / (HW F);   u_r[v] = B_r G[v];   t_r[v] = d_r.G[v]
//   dA_r[b,hw] = Xd[b,hw].u_r[v] + t_r[v]          (pass 1 over X)
//   delta_r    = chain / relu / softmax backward of dA_r                               (per image, tiny)
//   dX[b,hw]   = mask * sum_r A_r u_r[v] + delta_0 w_0;   dw_0 = sum delta_0 X         (pass 2 over X, dX written)
//   dB_r = g_r^T G;  dd_r = s_r^T G;  dc_r = sum delta_r;  dw_r = sum delta_r a_{r-1}  (r >= 1)
//
// Memory traffic: forward 2 x |X| (attention values, then the masked pooling), backward 2 x |X| read + |X| written;
// the [HW, K] top-down maps never exist.  All kernels are plain grids (no counters, no PDL): every per-call buffer is
// fully written before it is read in stream order, all reductions have a fixed order (bit-reproducible).
#include <stdint.h>

#include <algorithm>

#include "../../include/apa.h"
#include "apa_common.cuh"
#include "apa_kernels.h"

namespace apa {
namespace {

constexpr int kPoolThreads = 512;   // one float4 of channels per thread: 2048 channels per sweep
constexpr int kPF = 8;              // rows a thread keeps in flight in the passes over X
static_assert(kPF == 8, "bwd_fused_kernel packs 4 mask bits per row into one 32-bit word");
constexpr int kSlices = 16;         // channel slices of the logits contraction
constexpr int kSMs = 148;
constexpr int kSlots = kSMs;        // CTAs of the X passes resident at once: the fused kernels run one (persistent) per SM

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// murmur3's 32-bit finaliser: a bijection of 32-bit words with full avalanche; two of them (two keys) give the 64 bits a
// group of four elements needs.  ~12 integer instructions per float4 of X.
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}

struct Drop {
  uint32_t key0, key1;  // low / high word of splitmix64(seed)
  uint32_t thr;         // keep iff 16-bit field < thr;  thr = round(keep * 65536)
  float scale;          // 1 / keep
  int on;               // 0: keep == 1, no hashing at all
};

// mask of the four consecutive elements 4*g4 .. 4*g4+3 of X (row-major [batch, hw, channels], channels % 4 == 0;
// g4 < 2^32 is checked on the host)
__device__ __forceinline__ uint32_t drop_bits4(const Drop& d, uint64_t g4) {
  const uint32_t a = fmix32(uint32_t(g4) ^ d.key0), b = fmix32(uint32_t(g4) ^ d.key1);
  return ((a & 0xFFFFu) < d.thr ? 1u : 0u) | ((a >> 16) < d.thr ? 2u : 0u) | ((b & 0xFFFFu) < d.thr ? 4u : 0u) |
         ((b >> 16) < d.thr ? 8u : 0u);
}
__device__ __forceinline__ float4 bits_mask4(uint32_t bits, float scale) {
  return make_float4((bits & 1u) ? scale : 0.f, (bits & 2u) ? scale : 0.f, (bits & 4u) ? scale : 0.f, (bits & 8u) ? scale : 0.f);
}
__device__ __forceinline__ float4 drop_mask4(const Drop& d, uint64_t g4) { return bits_mask4(drop_bits4(d, g4), d.scale); }

struct TrainParams {
  const float* X;
  int N, HW, C, K, F, R, mode, J, rows;   // N images = videos * F;  J row chunks of `rows` rows per image
  int parts;                              // backward: rows of dw0part / pch that were written (CTAs or chunks)
  const float* aw[APA_MAX_RANK];    // aw[0]: [C];  aw[r >= 1]: scalar
  const float* ab[APA_MAX_RANK];    // scalars
  const float* tw[APA_MAX_RANK];    // [C, K]
  const float* tb[APA_MAX_RANK];    // [K]
  Drop drop;
  // saved by the forward for the backward
  float* a0;     // [N, HW]
  float* gv;     // [videos, R, C]
  float* sv;     // [videos, R]
  // forward scratch
  float* gpart;  // [N, J, R, C]
  float* spart;  // [N, J, R]
  float* cpart;  // [kSlices, videos, K]
  float* logits; // [videos, K]
  // backward
  const float* dlogits;  // [videos, K]
  float* u;      // [videos, R, C]
  float* tt;     // [videos, R]
  float* dA;     // [N, R, HW]
  float* dw0part;// [N, J, C], then [G1, C] (first level of the sum)
  float* pch;    // [N * J, 2, R]: per chunk, sum delta_r and sum delta_r a_{r-1}
  float* dX;     // [N, HW, C] or nullptr
  float* d_aw[APA_MAX_RANK];
  float* d_ab[APA_MAX_RANK];
  float* d_tw[APA_MAX_RANK];
  float* d_tb[APA_MAX_RANK];
};

// 4-byte asynchronous global -> shared copies: a thread issues all of its tile loads back to back (no registers, one
// round trip) and waits once
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// fixed-order block reductions (every thread gets the result); `red` holds 32 floats
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = red[0];
  for (int i = 1; i < nw; ++i) t = fmaxf(t, red[i]);
  return t;
}

// ------------------------------------------------------------------------------------------------ forward
// a_0[row] = X[row].w_0 + c_0 (un-dropped features, nets_factory.py:259-270): one warp per row.
__global__ void __launch_bounds__(256) train_attn_kernel(TrainParams p) {
  griddep_launch_dependents();   // the next kernel of the call may be scheduled; it waits for this one to finish
  griddep_wait();
  extern __shared__ float4 w0s[];
  const int C4 = p.C >> 2;
  for (int i = threadIdx.x; i < C4; i += blockDim.x) w0s[i] = __ldg(reinterpret_cast<const float4*>(p.aw[0]) + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float c0 = __ldg(p.ab[0]);
  const long rows = long(p.N) * p.HW;
  for (long row = long(blockIdx.x) * 8 + warp; row < rows; row += long(gridDim.x) * 8) {
    const float4* x = reinterpret_cast<const float4*>(p.X) + row * C4;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
    for (int c4 = lane; c4 < C4; c4 += 64) {
      const float4 v = __ldg(x + c4), w = w0s[c4];
      acc0 += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
      if (c4 + 32 < C4) {
        const float4 v2 = __ldg(x + c4 + 32), w2 = w0s[c4 + 32];
        acc1 += v2.x * w2.x + v2.y * w2.y + v2.z * w2.z + v2.w * w2.w;
      }
    }
    const float acc = warp_sum(acc0 + acc1);
    if (lane == 0) p.a0[row] = acc + c0;
  }
}

// The attention values of image b for every rank, into shared memory (A[r * HW + hw]); softmax over HW when asked.
// Optionally also the pre-activation chain values (araw, may be nullptr).
__device__ __forceinline__ void image_attention(const TrainParams& p, int b, float* A, float* araw, float* red) {
  const int HW = p.HW, R = p.R;
  float wr[APA_MAX_RANK], cr[APA_MAX_RANK];
#pragma unroll
  for (int r = 1; r < APA_MAX_RANK; ++r)
    if (r < R) { wr[r] = __ldg(p.aw[r]); cr[r] = __ldg(p.ab[r]); }
  for (int hw = threadIdx.x; hw < HW; hw += blockDim.x) {
    float a = p.a0[size_t(b) * HW + hw];
#pragma unroll
    for (int r = 0; r < APA_MAX_RANK; ++r)
      if (r < R) {
        if (r > 0) a = wr[r] * a + cr[r];
        if (araw) araw[r * HW + hw] = a;
        A[r * HW + hw] = (p.mode == MODE_RELU) ? fmaxf(a, 0.f) : a;
      }
  }
  __syncthreads();
  if (p.mode == MODE_SOFTMAX) {   // R == 1 (checked on the host)
    float mx = -INFINITY;
    for (int hw = threadIdx.x; hw < HW; hw += blockDim.x) mx = fmaxf(mx, A[hw]);
    mx = block_max(mx, red);
    float s = 0.f;
    for (int hw = threadIdx.x; hw < HW; hw += blockDim.x) { const float e = __expf(A[hw] - mx); A[hw] = e; s += e; }
    s = block_sum(s, red);
    const float inv = 1.f / s;
    for (int hw = threadIdx.x; hw < HW; hw += blockDim.x) A[hw] *= inv;
    __syncthreads();
  }
}

// kPF rows of one float4 column, issued together (rows past h1 read as zero)
__device__ __forceinline__ void load_rows(float4 (&xr)[kPF], const float4* col, int C4, int hb, int h1) {
#pragma unroll
  for (int i = 0; i < kPF; ++i)
    xr[i] = (hb + i < h1) ? __ldg(col + size_t(hb + i) * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// partial g_r of (image b, row chunk j): sum over the chunk's rows of A_r[hw] * mask * X[b, hw, :].  The first rows
// are in flight while the CTA works out the image's attention values.
template <int RM>
__global__ void __launch_bounds__(kPoolThreads, RM == 1 ? 2 : 1) train_pool_kernel(TrainParams p) {
  griddep_launch_dependents();   // the next kernel of the call may be scheduled; it waits for this one to finish
  griddep_wait();
  extern __shared__ float smem[];
  float* A = smem;                       // [R * HW]
  float* red = smem + p.R * p.HW;        // [32]
  const int j = blockIdx.x, b = blockIdx.y, HW = p.HW, R = p.R, C4 = p.C >> 2;
  const int h0 = j * p.rows, h1 = min(HW, h0 + p.rows);
  const float4* Xb = reinterpret_cast<const float4*>(p.X) + size_t(b) * HW * C4;
  const uint64_t g4b = uint64_t(b) * HW * C4;
  float4 xr[kPF];
  if (threadIdx.x < C4) load_rows(xr, Xb + threadIdx.x, C4, h0, h1);
  image_attention(p, b, A, nullptr, red);
  for (int c4 = threadIdx.x; c4 < C4; c4 += blockDim.x) {
    float4 acc[RM];
#pragma unroll
    for (int r = 0; r < RM; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int hb = h0; hb < h1; hb += kPF) {
      if (hb != h0 || c4 != threadIdx.x) load_rows(xr, Xb + c4, C4, hb, h1);
#pragma unroll
      for (int i = 0; i < kPF; ++i) {
        const int hw = min(hb + i, HW - 1);   // rows past h1 carry x == 0
        float4 x = xr[i];
        if (p.drop.on) {
          const float4 m = drop_mask4(p.drop, g4b + uint64_t(hw) * C4 + c4);
          x.x *= m.x; x.y *= m.y; x.z *= m.z; x.w *= m.w;
        }
#pragma unroll
        for (int r = 0; r < RM; ++r)
          if (r < R) {
            const float a = A[r * HW + hw];
            acc[r].x += a * x.x; acc[r].y += a * x.y; acc[r].z += a * x.z; acc[r].w += a * x.w;
          }
      }
    }
#pragma unroll
    for (int r = 0; r < RM; ++r)
      if (r < R)
        reinterpret_cast<float4*>(p.gpart)[((size_t(b) * p.J + j) * R + r) * C4 + c4] = acc[r];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < R) {
    float s = 0.f;
    for (int hw = h0 + lane; hw < h1; hw += 32) s += A[warp * HW + hw];
    s = warp_sum(s);
    if (lane == 0) p.spart[(size_t(b) * p.J + j) * R + warp] = s;
  }
}

// g_r[v] = sum over the video's frames and row chunks (fixed order); s_r[v] likewise
__global__ void __launch_bounds__(256) train_reduce_kernel(TrainParams p) {
  griddep_launch_dependents();   // the next kernel of the call may be scheduled; it waits for this one to finish
  griddep_wait();
  const int v = blockIdx.y, RC4 = p.R * (p.C >> 2);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int parts = p.F * p.J;
  if (i < RC4) {
    const float4* src = reinterpret_cast<const float4*>(p.gpart) + size_t(v) * parts * RC4 + i;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int q = 0; q < parts; ++q) {
      const float4 t = src[size_t(q) * RC4];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    reinterpret_cast<float4*>(p.gv)[size_t(v) * RC4 + i] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x < p.R) {
    float s = 0.f;
    for (int q = 0; q < parts; ++q) s += p.spart[(size_t(v) * parts + q) * p.R + threadIdx.x];
    p.sv[size_t(v) * p.R + threadIdx.x] = s;
  }
}

// logits[v, k] = (sum_r g_r[v].B_r[:, k] + d_r[k] s_r[v]) / (HW F) in two steps: CONTRACT forms the partial products of
// one channel slice (kSlices of them) for 32 videos x 32 classes per CTA out of shared memory, FINISH adds the slices in
// order, the bias term and the scale.
__global__ void __launch_bounds__(256) train_contract_kernel(TrainParams p) {
  griddep_launch_dependents();   // the next kernel of the call may be scheduled; it waits for this one to finish
  griddep_wait();
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int V = p.N / p.F, C = p.C, K = p.K, R = p.R;
  const int k0 = blockIdx.x * 32, sl = blockIdx.y, v0 = blockIdx.z * 32;
  const int cs = (((C + kSlices - 1) / kSlices) + 3) & ~3;
  const int c0 = sl * cs, cn = max(0, min(cs, C - c0));
  float* tws = sm;              // [cs][32]
  float* gs = sm + cs * 32;     // [32][cs]
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bool kok = k0 + lane < K;
  for (int r = 0; r < R; ++r) {
    __syncthreads();
    const float* tw = p.tw[r] + size_t(c0) * K + k0 + lane;
    for (int c = warp; c < cn; c += 8) {
      if (kok) cp_async4(tws + c * 32 + lane, tw + size_t(c) * K);
      else tws[c * 32 + lane] = 0.f;
    }
#pragma unroll
    for (int vv = 0; vv < 4; ++vv) {
      const int v = warp * 4 + vv;
      const float* g = p.gv + (size_t(min(v0 + v, V - 1)) * R + r) * C + c0;
      for (int c = lane; c < cn; c += 32) cp_async4(gs + v * cs + c, g + c);
    }
    cp_async_wait_all();
    __syncthreads();
    const float* g0 = gs + (warp * 4) * cs;
#pragma unroll 2
    for (int c = 0; c < cn; c += 4) {   // cn % 4 == 0
      const float w0 = tws[c * 32 + lane], w1 = tws[(c + 1) * 32 + lane], w2 = tws[(c + 2) * 32 + lane],
                  w3 = tws[(c + 3) * 32 + lane];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 g = *reinterpret_cast<const float4*>(g0 + i * cs + c);
        acc[i] += g.x * w0 + g.y * w1 + g.z * w2 + g.w * w3;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = v0 + warp * 4 + i;
    if (v < V && kok) p.cpart[(size_t(sl) * V + v) * K + k0 + lane] = acc[i];
  }
}
__global__ void __launch_bounds__(256) train_finish_kernel(TrainParams p) {
  griddep_launch_dependents();   // the next kernel of the call may be scheduled; it waits for this one to finish
  griddep_wait();
  const int V = p.N / p.F, K = p.K, R = p.R;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V * K) return;
  const int v = i / K, k = i % K;
  float t = 0.f;
#pragma unroll
  for (int s = 0; s < kSlices; ++s) t += p.cpart[size_t(s) * V * K + i];
  for (int r = 0; r < R; ++r) t += __ldg(p.tb[r] + k) * p.sv[size_t(v) * R + r];
  p.logits[i] = t / (float(p.HW) * float(p.F));
}

// kPF per-thread partial sums -> kPF block-wide sums in every thread (fixed order: xor butterfly inside a warp, then
// the warps in order).  red: [17][kPF] floats.
__device__ __forceinline__ void block_sum_rows(float (&v)[kPF], float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // 8 values over 32 lanes in 4 + 2 + 1 + 1 + 1 exchanges: each step halves the values a lane still carries
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  if ((lane & 3) == 0) red[warp * kPF + (lane >> 2)] = v[0];   // lane 4m carries value m
  __syncthreads();
  if (warp == 0 && lane < kPF) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += red[w * kPF + lane];
    red[16 * kPF + lane] = t;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kPF; ++i) v[i] = red[16 * kPF + i];
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// ---- One pass over X when the attention is local to a row (rank 1, relu / plain attention, channels <= 4 * blockDim).
// Persistent CTAs (one per SM) walk items = (image, row chunk); thread = one float4 column, kPF rows per batch, and the
// next batch's rows are in flight while the current batch is reduced and consumed (two register sets, ping-pong).
struct Walk {
  int item, hb;   // current item (b * J + j) and first row of the current batch
};
__device__ __forceinline__ bool walk_valid(const TrainParams& p, const Walk& w) { return w.item < p.N * p.J; }
__device__ __forceinline__ int walk_h1(const TrainParams& p, const Walk& w) { return min(p.HW, (w.item % p.J + 1) * p.rows); }
__device__ __forceinline__ Walk walk_next(const TrainParams& p, Walk w) {
  w.hb += kPF;
  if (w.hb >= walk_h1(p, w)) {
    w.item += gridDim.x;
    w.hb = (w.item % p.J) * p.rows;
  }
  return w;
}
// kPF rows of the thread's column for batch w.  No predicates: rows past the chunk re-read its last row and columns
// past C re-read the last column; the consumers give those zero weight.
__device__ __forceinline__ void walk_load(const TrainParams& p, const Walk& w, float4 (&xr)[kPF]) {
  const int C4 = p.C >> 2;
  if (walk_valid(p, w)) {
    const int hl = walk_h1(p, w) - 1;
    const float4* col = reinterpret_cast<const float4*>(p.X) + size_t(w.item / p.J) * p.HW * C4 + min(int(threadIdx.x), C4 - 1);
#pragma unroll
    for (int i = 0; i < kPF; ++i) xr[i] = __ldg(col + size_t(min(w.hb + i, hl)) * C4);
  }
}

struct FwdState {
  float4 w04, acc;
  float c0, ssum;
};
// forward: the rows' dot products with w_0 are summed across the CTA, then the same registers feed the masked pooling.
__device__ __forceinline__ void fwd_batch(const TrainParams& p, const Walk& w, bool act, const float4 (&xr)[kPF], FwdState& st,
                                          float* red, bool item_ends) {
  const int C4 = p.C >> 2, b = w.item / p.J, nr = walk_h1(p, w) - w.hb;   // rows of this batch that count
  float part[kPF];
#pragma unroll
  for (int i = 0; i < kPF; ++i) part[i] = dot4(xr[i], st.w04);   // w04 == 0 in the columns past C
  block_sum_rows(part, red);
  const uint32_t g0 = uint32_t(b) * p.HW * C4 + uint32_t(w.hb) * C4 + threadIdx.x;   // group index of row hb (< 2^32, host check)
#pragma unroll
  for (int i = 0; i < kPF; ++i) {
    const float a = part[i] + st.c0;
    if (int(threadIdx.x) == i && i < nr) p.a0[size_t(b) * p.HW + w.hb + i] = a;
    float A = (p.mode == MODE_RELU) ? fmaxf(a, 0.f) : a;
    A = (i < nr) ? A : 0.f;
    st.ssum += A;
    const float4 x = xr[i];
    if (p.drop.on) {
      const uint32_t g = g0 + uint32_t(i) * C4;
      const uint32_t ha = fmix32(g ^ p.drop.key0), hb2 = fmix32(g ^ p.drop.key1);
      const float As = A * p.drop.scale;
      st.acc.x = fmaf((ha & 0xFFFFu) < p.drop.thr ? As : 0.f, x.x, st.acc.x);
      st.acc.y = fmaf((ha >> 16) < p.drop.thr ? As : 0.f, x.y, st.acc.y);
      st.acc.z = fmaf((hb2 & 0xFFFFu) < p.drop.thr ? As : 0.f, x.z, st.acc.z);
      st.acc.w = fmaf((hb2 >> 16) < p.drop.thr ? As : 0.f, x.w, st.acc.w);
    } else {
      st.acc.x = fmaf(A, x.x, st.acc.x); st.acc.y = fmaf(A, x.y, st.acc.y);
      st.acc.z = fmaf(A, x.z, st.acc.z); st.acc.w = fmaf(A, x.w, st.acc.w);
    }
  }
  if (item_ends) {
    if (act) reinterpret_cast<float4*>(p.gpart)[size_t(w.item) * C4 + threadIdx.x] = st.acc;
    if (threadIdx.x == 0) p.spart[w.item] = st.ssum;
    st.acc = make_float4(0.f, 0.f, 0.f, 0.f);
    st.ssum = 0.f;
  }
}
__global__ void __launch_bounds__(kPoolThreads, 1) train_fused_fwd_kernel(TrainParams p) {
  __shared__ float red[17 * kPF];
  griddep_launch_dependents();
  griddep_wait();
  const bool act = int(threadIdx.x) < (p.C >> 2);
  FwdState st;
  st.w04 = act ? __ldg(reinterpret_cast<const float4*>(p.aw[0]) + threadIdx.x) : make_float4(0.f, 0.f, 0.f, 0.f);
  st.c0 = __ldg(p.ab[0]);
  st.acc = make_float4(0.f, 0.f, 0.f, 0.f);
  st.ssum = 0.f;
  Walk cur;
  cur.item = blockIdx.x;
  cur.hb = (cur.item % p.J) * p.rows;
  float4 xa[kPF], xb[kPF];
  walk_load(p, cur, xa);
  while (walk_valid(p, cur)) {
    Walk nxt = walk_next(p, cur);
    walk_load(p, nxt, xb);
    fwd_batch(p, cur, act, xa, st, red, nxt.item != cur.item);
    cur = nxt;
    if (!walk_valid(p, cur)) break;
    nxt = walk_next(p, cur);
    walk_load(p, nxt, xa);
    fwd_batch(p, cur, act, xb, st, red, nxt.item != cur.item);
    cur = nxt;
  }
}

struct BwdState {
  float4 w04, u4, dw;
  float tt0, sd;
  int par, v;
};
// backward: the rows' dot products with u[v] give dA, the row's delta follows locally (relu / plain), and the registers
// that hold the rows produce dX and the CTA's part of dw_0.
__device__ __forceinline__ void bwd_batch(const TrainParams& p, const Walk& w, bool act, const float4 (&xr)[kPF], BwdState& st,
                                          float* red) {
  const int C4 = p.C >> 2, HW = p.HW, b = w.item / p.J, nr = walk_h1(p, w) - w.hb;
  if (b / p.F != st.v) {   // a new video: its u and t (u == 0 in the columns past C)
    st.v = b / p.F;
    st.u4 = act ? reinterpret_cast<const float4*>(p.u)[size_t(st.v) * C4 + threadIdx.x] : make_float4(0.f, 0.f, 0.f, 0.f);
    st.tt0 = p.tt[st.v];
  }
  float* avs = red + (17 + (st.par & 1)) * kPF;   // the rows' a_0, read after the barriers inside block_sum_rows
  st.par ^= 1;
  if (threadIdx.x >= 32 && threadIdx.x < 32 + kPF)
    avs[threadIdx.x - 32] = (int(threadIdx.x) - 32 < nr) ? p.a0[size_t(b) * HW + w.hb + threadIdx.x - 32] : 0.f;
  const uint32_t g0 = uint32_t(b) * HW * C4 + uint32_t(w.hb) * C4 + threadIdx.x;
  uint32_t bits = 0xFFFFFFFFu;
  float part[kPF];
#pragma unroll
  for (int i = 0; i < kPF; ++i) {
    const float4 x = xr[i];
    if (p.drop.on) {
      const uint32_t g = g0 + uint32_t(i) * C4;
      const uint32_t ha = fmix32(g ^ p.drop.key0), hb2 = fmix32(g ^ p.drop.key1);
      const bool k0 = (ha & 0xFFFFu) < p.drop.thr, k1 = (ha >> 16) < p.drop.thr, k2 = (hb2 & 0xFFFFu) < p.drop.thr,
                 k3 = (hb2 >> 16) < p.drop.thr;
      bits = (bits << 4) | (k0 ? 1u : 0u) | (k1 ? 2u : 0u) | (k2 ? 4u : 0u) | (k3 ? 8u : 0u);   // row i ends up at nibble kPF-1-i
      part[i] = ((k0 ? x.x * st.u4.x : 0.f) + (k1 ? x.y * st.u4.y : 0.f) + (k2 ? x.z * st.u4.z : 0.f) +
                 (k3 ? x.w * st.u4.w : 0.f)) * p.drop.scale;
    } else {
      part[i] = dot4(x, st.u4);
    }
  }
  block_sum_rows(part, red);
  float4* dcol = (p.dX && act) ? reinterpret_cast<float4*>(p.dX) + size_t(b) * HW * C4 + size_t(w.hb) * C4 + threadIdx.x : nullptr;
#pragma unroll
  for (int i = 0; i < kPF; ++i) {
    const float a = avs[i];
    float d = part[i] + st.tt0;
    if (p.mode == MODE_RELU) d = a > 0.f ? d : 0.f;
    d = (i < nr) ? d : 0.f;
    const float A = (p.mode == MODE_RELU) ? fmaxf(a, 0.f) : a;
    st.sd += d;
    const float4 x = xr[i];
    st.dw.x = fmaf(d, x.x, st.dw.x); st.dw.y = fmaf(d, x.y, st.dw.y);
    st.dw.z = fmaf(d, x.z, st.dw.z); st.dw.w = fmaf(d, x.w, st.dw.w);
    if (dcol && i < nr) {
      const float As = p.drop.on ? A * p.drop.scale : A;
      const uint32_t nib = bits >> (4 * (kPF - 1 - i));
      float4 t;
      t.x = fmaf((nib & 1u) ? As : 0.f, st.u4.x, d * st.w04.x);
      t.y = fmaf((nib & 2u) ? As : 0.f, st.u4.y, d * st.w04.y);
      t.z = fmaf((nib & 4u) ? As : 0.f, st.u4.z, d * st.w04.z);
      t.w = fmaf((nib & 8u) ? As : 0.f, st.u4.w, d * st.w04.w);
      __stcs(dcol + size_t(i) * C4, t);
    }
  }
}
__global__ void __launch_bounds__(kPoolThreads, 1) bwd_fused_kernel(TrainParams p) {
  __shared__ float red[19 * kPF];   // [17] block_sum_rows, [2] the rows' a_0 (double-buffered)
  griddep_launch_dependents();
  griddep_wait();
  const int C4 = p.C >> 2;
  const bool act = int(threadIdx.x) < C4;
  BwdState st;
  st.w04 = act ? __ldg(reinterpret_cast<const float4*>(p.aw[0]) + threadIdx.x) : make_float4(0.f, 0.f, 0.f, 0.f);
  st.u4 = st.dw = make_float4(0.f, 0.f, 0.f, 0.f);
  st.tt0 = st.sd = 0.f;
  st.par = 0;
  st.v = -1;
  Walk cur;
  cur.item = blockIdx.x;
  cur.hb = (cur.item % p.J) * p.rows;
  float4 xa[kPF], xb[kPF];
  walk_load(p, cur, xa);
  while (walk_valid(p, cur)) {
    Walk nxt = walk_next(p, cur);
    walk_load(p, nxt, xb);
    bwd_batch(p, cur, act, xa, st, red);
    cur = nxt;
    if (!walk_valid(p, cur)) break;
    nxt = walk_next(p, cur);
    walk_load(p, nxt, xa);
    bwd_batch(p, cur, act, xb, st, red);
    cur = nxt;
  }
  // the CTA's parts of dw_0 and of dc_0 (one part per CTA)
  if (act) reinterpret_cast<float4*>(p.dw0part)[size_t(blockIdx.x) * C4 + threadIdx.x] = st.dw;
  if (threadIdx.x == 0) {
    p.pch[size_t(blockIdx.x) * 2 + 0] = st.sd;
    p.pch[size_t(blockIdx.x) * 2 + 1] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ backward
// u_r[v, c] = B_r[c, :].G[v],  t_r[v] = d_r.G[v]  with G = dlogits / (HW F): 32 channels x 8 videos per CTA, the 32 rows
// of B_r (contiguous in memory) and the 8 rows of G in shared memory; warp = video, lane = channel
__device__ __forceinline__ void bwd_u_block(const TrainParams& p, float* sm, int bx, int by, int r) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v0 = by * 8, c0 = bx * 32, V = p.N / p.F, K = p.K, C = p.C, R = p.R;
  const int Kp = K | 1;             // odd row stride: lanes (channels) fall into distinct banks
  const int K4 = (K + 3) & ~3;      // 16-byte aligned rows of G
  float* tws = sm;                  // [32][Kp]
  float* Gs = sm + 32 * Kp;         // [8][K4]
  const float scale = 1.f / (float(p.HW) * float(p.F));
  const int nc = min(32, C - c0);
  for (int c = warp; c < nc; c += 8) {
    const float* src = p.tw[r] + size_t(c0 + c) * K;
    for (int k = lane; k < K; k += 32) cp_async4(tws + c * Kp + k, src + k);
  }
  {
    const int v = min(v0 + warp, V - 1);
    for (int k = lane; k < K; k += 32) cp_async4(Gs + warp * K4 + k, p.dlogits + size_t(v) * K + k);   // scaled at the end
  }
  cp_async_wait_all();
  __syncthreads();
  if (v0 + warp < V) {
    const float* g = Gs + warp * K4;
    if (lane < nc) {
      const float* t = tws + lane * Kp;
      float a0 = 0.f, a1 = 0.f;
      int k = 0;
#pragma unroll 2
      for (; k + 3 < K; k += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(g + k);
        a0 += g4.x * t[k] + g4.z * t[k + 2];
        a1 += g4.y * t[k + 1] + g4.w * t[k + 3];
      }
      for (; k < K; ++k) a0 += g[k] * t[k];
      p.u[(size_t(v0 + warp) * R + r) * C + c0 + lane] = (a0 + a1) * scale;
    }
    if (bx == 0) {
      float t = 0.f;
      for (int k = lane; k < K; k += 32) t += g[k] * __ldg(p.tb[r] + k);
      t = warp_sum(t);
      if (lane == 0) p.tt[size_t(v0 + warp) * R + r] = t * scale;
    }
  }
}

// pass 1 over X: dA_r[b, hw] = (mask * X[b, hw]).u_r[v] + t_r[v]; warp per row, u_r[v] through L1
__global__ void __launch_bounds__(256) bwd_dattn_kernel(TrainParams p) {
  griddep_launch_dependents();   // the next kernel of the call may be scheduled; it waits for this one to finish
  griddep_wait();
  const int j = blockIdx.x, b = blockIdx.y, v = b / p.F, HW = p.HW, R = p.R, C4 = p.C >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h0 = j * p.rows, h1 = min(HW, h0 + p.rows);
  const float4* Xb = reinterpret_cast<const float4*>(p.X) + size_t(b) * HW * C4;
  const float4* uv = reinterpret_cast<const float4*>(p.u) + size_t(v) * R * C4;
  const uint64_t g4b = uint64_t(b) * HW * C4;
  for (int hw = h0 + warp; hw < h1; hw += 8) {
    float acc[APA_MAX_RANK];
#pragma unroll
    for (int r = 0; r < APA_MAX_RANK; ++r) acc[r] = 0.f;
#pragma unroll 8
    for (int c4 = lane; c4 < C4; c4 += 32) {
      float4 x = __ldg(Xb + size_t(hw) * C4 + c4);
      if (p.drop.on) {
        const float4 m = drop_mask4(p.drop, g4b + uint64_t(hw) * C4 + c4);
        x.x *= m.x; x.y *= m.y; x.z *= m.z; x.w *= m.w;
      }
#pragma unroll
      for (int r = 0; r < APA_MAX_RANK; ++r)
        if (r < R) {
          const float4 w = uv[r * C4 + c4];
          acc[r] += x.x * w.x + x.y * w.y + x.z * w.z + x.w * w.w;
        }
    }
#pragma unroll
    for (int r = 0; r < APA_MAX_RANK; ++r)
      if (r < R) {
        const float t = warp_sum(acc[r]);
        if (lane == 0) p.dA[(size_t(b) * R + r) * HW + hw] = t + p.tt[size_t(v) * R + r];
      }
  }
}

// pass 2 over X: per image the chain / relu / softmax backward (tiny, recomputed by each of the image's CTAs), then
//   dX[b, hw, :] = mask * sum_r A_r[hw] u_r[v] + delta_0[hw] w_0      and the CTA's part of dw_0 = sum delta_0 X
template <int RM>
__global__ void __launch_bounds__(kPoolThreads, RM == 1 ? 2 : 1) bwd_dx_kernel(TrainParams p) {
  griddep_launch_dependents();   // the next kernel of the call may be scheduled; it waits for this one to finish
  griddep_wait();
  extern __shared__ float smem[];
  const int HW = p.HW, R = p.R, C4 = p.C >> 2;
  float* A = smem;                 // [R * HW]
  float* araw = A + R * HW;        // [R * HW]
  float* d0 = araw + R * HW;       // [HW]
  float* red = d0 + HW;            // [32]
  const int j = blockIdx.x, b = blockIdx.y, v = b / p.F;
  const int h0 = j * p.rows, h1 = min(HW, h0 + p.rows);
  const float4* Xb = reinterpret_cast<const float4*>(p.X) + size_t(b) * HW * C4;
  float4 xr[kPF];
  if (threadIdx.x < C4) load_rows(xr, Xb + threadIdx.x, C4, h0, h1);
  image_attention(p, b, A, araw, red);
  float wr[4];
#pragma unroll
  for (int r = 1; r < 4; ++r) wr[r] = (r < R) ? __ldg(p.aw[r]) : 0.f;
  float sAd = 0.f;
  if (p.mode == MODE_SOFTMAX) {
    float t = 0.f;
    for (int hw = threadIdx.x; hw < HW; hw += blockDim.x) t += A[hw] * p.dA[size_t(b) * HW + hw];
    sAd = block_sum(t, red);
  }
  float sd[RM], sda[RM];
#pragma unroll
  for (int r = 0; r < RM; ++r) sd[r] = sda[r] = 0.f;
  const int h0c = j * p.rows, h1c = min(HW, h0c + p.rows);
  for (int hw = h0c + threadIdx.x; hw < h1c; hw += blockDim.x) {   // delta of the CTA's own rows
    float da[RM];
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      da[r] = 0.f;
      if (r < R) {
        float d = p.dA[(size_t(b) * R + r) * HW + hw];
        if (p.mode == MODE_RELU) d = araw[r * HW + hw] > 0.f ? d : 0.f;
        if (p.mode == MODE_SOFTMAX) d = A[hw] * (d - sAd);
        da[r] = d;
      }
    }
    float delta = 0.f;
#pragma unroll
    for (int r = RM - 1; r >= 0; --r)
      if (r < R) {
        delta = da[r] + ((r + 1 < R) ? wr[(r + 1) & 3] * delta : 0.f);
        sd[r] += delta;
        if (r > 0) sda[r] += delta * araw[(r - 1) * HW + hw];
      }
    d0[hw] = delta;
  }
  // the chunk's sums for dc_r and dw_r (r >= 1)
#pragma unroll
  for (int r = 0; r < RM; ++r)
    if (r < R) {
      const float s1 = block_sum(sd[r], red);
      const float s2 = (r > 0) ? block_sum(sda[r], red) : 0.f;
      if (threadIdx.x == 0) {
        p.pch[((size_t(b) * p.J + j) * 2 + 0) * R + r] = s1;
        p.pch[((size_t(b) * p.J + j) * 2 + 1) * R + r] = s2;
      }
    }
  __syncthreads();
  float4* dXb = p.dX ? reinterpret_cast<float4*>(p.dX) + size_t(b) * HW * C4 : nullptr;
  const uint64_t g4b = uint64_t(b) * HW * C4;
  for (int c4 = threadIdx.x; c4 < C4; c4 += blockDim.x) {
    float4 u4[RM];
#pragma unroll
    for (int r = 0; r < RM; ++r)
      if (r < R) u4[r] = reinterpret_cast<const float4*>(p.u)[(size_t(v) * R + r) * C4 + c4];
    const float4 w04 = __ldg(reinterpret_cast<const float4*>(p.aw[0]) + c4);
    float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int hb = h0; hb < h1; hb += kPF) {
      if (hb != h0 || c4 != threadIdx.x) load_rows(xr, Xb + c4, C4, hb, h1);
#pragma unroll
      for (int i = 0; i < kPF; ++i) {
        const int hw = hb + i;
        if (hw < h1) {
          const float4 x = xr[i];
          const float d = d0[hw];
          dw.x += d * x.x; dw.y += d * x.y; dw.z += d * x.z; dw.w += d * x.w;
          if (dXb) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < RM; ++r)
              if (r < R) {
                const float a = A[r * HW + hw];
                t.x += a * u4[r].x; t.y += a * u4[r].y; t.z += a * u4[r].z; t.w += a * u4[r].w;
              }
            if (p.drop.on) {
              const float4 m = drop_mask4(p.drop, g4b + uint64_t(hw) * C4 + c4);
              t.x *= m.x; t.y *= m.y; t.z *= m.z; t.w *= m.w;
            }
            t.x += d * w04.x; t.y += d * w04.y; t.z += d * w04.z; t.w += d * w04.w;
            __stcs(dXb + size_t(hw) * C4 + c4, t);
          }
        }
      }
    }
    reinterpret_cast<float4*>(p.dw0part)[(size_t(b) * p.J + j) * C4 + c4] = dw;
  }
}

// dB_r[c, k] = sum_v g_r[v, c] G[v, k];  dd_r[k] = sum_v s_r[v] G[v, k]: 32 channels x 256 classes per CTA pass, 32
// videos at a time in shared memory; thread = class, 32 channel accumulators
__device__ __forceinline__ void bwd_wgrad_block(const TrainParams& p, float* sm, int bx, int r) {
  float4(*gvs)[8] = reinterpret_cast<float4(*)[8]>(sm);            // [32][8] float4
  float(*Gs)[256] = reinterpret_cast<float(*)[256]>(sm + 32 * 32);  // [32][256]
  const int c0 = bx * 32, V = p.N / p.F, K = p.K, C = p.C, R = p.R;
  const float scale = 1.f / (float(p.HW) * float(p.F));
  const int nc = min(32, C - c0);   // a multiple of 4
  for (int kb = 0; kb < K; kb += 256) {
    const int k = kb + threadIdx.x;
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    for (int vt = 0; vt < V; vt += 32) {
      __syncthreads();
#pragma unroll 8
      for (int vv = 0; vv < 32; ++vv) {
        if (vt + vv < V && k < K) cp_async4(&Gs[vv][threadIdx.x], p.dlogits + size_t(vt + vv) * K + k);
        else Gs[vv][threadIdx.x] = 0.f;
      }
      {  // 32 videos x 8 float4 of channels: one float4 per thread
        const int vv = threadIdx.x >> 3, c4 = threadIdx.x & 7;
        gvs[vv][c4] = (vt + vv < V && c4 * 4 < nc)
                          ? *reinterpret_cast<const float4*>(p.gv + (size_t(vt + vv) * R + r) * C + c0 + c4 * 4)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      cp_async_wait_all();
      __syncthreads();
#pragma unroll 4
      for (int vv = 0; vv < 32; ++vv) {
        const float g = Gs[vv][threadIdx.x];
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 q = gvs[vv][c4];
          acc[c4 * 4 + 0] += q.x * g; acc[c4 * 4 + 1] += q.y * g; acc[c4 * 4 + 2] += q.z * g; acc[c4 * 4 + 3] += q.w * g;
        }
      }
    }
    if (k < K) {
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < nc) p.d_tw[r][size_t(c0 + c) * K + k] = acc[c] * scale;
      if (bx == 0) {
        float t = 0.f;
#pragma unroll 8
        for (int v = 0; v < V; ++v) t += p.sv[size_t(v) * R + r] * __ldg(p.dlogits + size_t(v) * K + k);
        p.d_tb[r][k] = t * scale;
      }
    }
  }
}

// Both need only dlogits and what the forward saved, and neither is bound by HBM: one launch, the CTAs of the two roles
// side by side (first nW CTAs: the top-down convs' gradients, the longer role; the rest: u / t).
__global__ void __launch_bounds__(256) bwd_small_kernel(TrainParams p, int nW, int cx, int vy) {
  griddep_launch_dependents();   // the next kernel of the call may be scheduled; it waits for this one to finish
  griddep_wait();
  extern __shared__ __align__(16) float sm_small[];
  const int i = blockIdx.x;
  if (i < nW) bwd_wgrad_block(p, sm_small, i % cx, i / cx);
  else bwd_u_block(p, sm_small, (i - nW) % cx, ((i - nW) / cx) % vy, (i - nW) / (cx * vy));
}

// Sum of `parts` rows of C floats in a fixed order: CTA (x, y) adds the parts of group y (`per` of them) for 32 float4
// columns, the group's parts split over the 8 warps, the warps' sums added in order.  Two launches: parts -> groups -> 1.
// The second one also adds the images' sums for dc_r and dw_r (r >= 1).
__global__ void __launch_bounds__(256) bwd_sum_kernel(TrainParams p, const float* __restrict__ src, float* __restrict__ dst,
                                                      int parts, int per, int final) {
  griddep_launch_dependents();   // the next kernel of the call may be scheduled; it waits for this one to finish
  griddep_wait();
  __shared__ float4 red[8][32];
  const int C4 = p.C >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const int g0 = blockIdx.y * per, g1 = min(parts, g0 + per);
  const int pw = (per + 7) / 8, q0 = g0 + warp * pw, q1 = min(g1, q0 + pw);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < C4) {
    const float4* s4 = reinterpret_cast<const float4*>(src) + i;
#pragma unroll 8
    for (int q = q0; q < q1; ++q) {
      const float4 t = s4[size_t(q) * C4];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && i < C4) {
    float4 t = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) { t.x += red[w][lane].x; t.y += red[w][lane].y; t.z += red[w][lane].z; t.w += red[w][lane].w; }
    reinterpret_cast<float4*>(dst)[size_t(blockIdx.y) * C4 + i] = t;
  }
  if (final && blockIdx.x == 0 && warp >= 1 && warp < 1 + 2 * p.R) {   // one warp per sum: lanes stride the chunks
    const int jj = warp - 1, which = jj / p.R, r = jj % p.R;
    float t = 0.f;
#pragma unroll 4
    for (int q = lane; q < p.parts; q += 32) t += p.pch[(size_t(q) * 2 + which) * p.R + r];
    t = warp_sum(t);
    if (lane == 0) {
      if (which == 0) p.d_ab[r][0] = t;
      else if (r > 0) p.d_aw[r][0] = t;
    }
  }
}

__global__ void dropout_mask_kernel(Drop d, float* out, size_t n4) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
    if (d.on) m = drop_mask4(d, i);
    reinterpret_cast<float4*>(out)[i] = m;
  }
}

inline size_t al(size_t n) { return (n + 63) & ~size_t(63); }   // floats, 256-byte granules

Drop make_drop(const ApaTrain* t) {
  Drop d;
  const float keep = t ? t->dropout_keep : 1.f;
  d.on = keep < 1.f;
  const uint64_t key = splitmix64(t ? t->dropout_seed : 0ull);
  d.key0 = uint32_t(key);
  d.key1 = uint32_t(key >> 32);
  double th = double(keep) * 65536.0 + 0.5;
  d.thr = th < 0 ? 0u : (th > 65536.0 ? 65536u : uint32_t(th));
  d.scale = 1.f / keep;
  return d;
}

int train_supported(const ApaDesc* d) {
  int st = apa_attn_pool_check(d);
  if (st != APA_OK) return st;
  if ((d->flags & APA_FLAG_PER_CLASS) || !(d->flags & APA_FLAG_SINGLE_LAYER_ATT)) return APA_ERR_UNSUPPORTED;
  if (double(d->batch) * d->hw * d->channels / 4.0 >= 4294967296.0) return APA_ERR_SHAPE;   // dropout counter is 32-bit
  if (size_t(d->rank) * d->hw * 2 + d->hw + 64 > 40000) return APA_ERR_SHAPE;   // per-image attention lives in smem
  if (size_t(d->num_classes) * 164 > 220 * 1024) return APA_ERR_SHAPE;   // bwd_u keeps 40 rows of K floats in smem
  return APA_OK;
}

// Row chunks per image for the passes over X: the (image, chunk) CTAs should fill whole waves of kSlots resident CTAs
// with as little left over as possible; each chunk costs a prologue worth a few rows.
int chunks_per_image(int N, int HW) {
  int best = 1;
  double best_cost = 1e300;
  for (int J = 1; J <= std::min(HW, 64); ++J) {
    const int rows = (HW + J - 1) / J;
    if ((HW + rows - 1) / rows != J) continue;   // J must be the chunk count its row count produces
    const double waves = double((long(N) * J + kSlots - 1) / kSlots);
    const double cost = waves * (rows + 4.0);
    if (cost < best_cost) { best_cost = cost; best = J; }
  }
  return best;
}
int sum_groups(size_t parts) { return int(std::min<size_t>(32, (parts + 31) / 32)); }

struct Carve {
  size_t a0, gv, sv, saved_total;                       // saved
  size_t gpart, spart, cpart, u, tt, dA, dw0part, pch, ws_total;   // workspace (floats)
};
Carve carve(const ApaDesc* d) {
  Carve c;
  const size_t N = d->batch, V = d->batch / d->frames, R = d->rank, C = d->channels, HW = d->hw;
  const size_t J = chunks_per_image(d->batch, d->hw);
  size_t o = 0;
  c.a0 = o; o += al(N * HW);
  c.gv = o; o += al(V * R * C);
  c.sv = o; o += al(V * R);
  c.saved_total = o;
  o = 0;
  c.gpart = o; o += al(N * J * R * C);
  c.spart = o; o += al(N * J * R);
  c.cpart = o; o += al(size_t(kSlices) * V * d->num_classes);
  c.u = o; o += al(V * R * C);
  c.tt = o; o += al(V * R);
  c.dA = o; o += al(N * R * HW);
  c.dw0part = o; o += al(N * J * C + size_t(sum_groups(N * J)) * C);
  c.pch = o; o += al(N * J * 2 * R);
  c.ws_total = o;
  return c;
}

int fill(TrainParams& p, const ApaDesc* d, const ApaWeights* w, const ApaTrain* t, const float* X, float* saved,
         float* ws) {
  const Carve c = carve(d);
  p.X = X;
  p.N = d->batch; p.HW = d->hw; p.C = d->channels; p.K = d->num_classes; p.F = d->frames; p.R = d->rank;
  p.mode = (d->flags & APA_FLAG_SOFTMAX_ATT) ? MODE_SOFTMAX : ((d->flags & APA_FLAG_RELU_ATT) ? MODE_RELU : MODE_PLAIN);
  p.J = chunks_per_image(d->batch, d->hw);
  p.rows = (d->hw + p.J - 1) / p.J;
  for (int r = 0; r < d->rank; ++r) {
    if (!w->attn_w[r] || !w->attn_b[r] || !w->topdown_w[r] || !w->topdown_b[r]) return APA_ERR_NULL;
    p.aw[r] = w->attn_w[r]; p.ab[r] = w->attn_b[r]; p.tw[r] = w->topdown_w[r]; p.tb[r] = w->topdown_b[r];
  }
  if ((reinterpret_cast<uintptr_t>(X) & 15u) || (reinterpret_cast<uintptr_t>(w->attn_w[0]) & 15u)) return APA_ERR_ALIGN;
  p.drop = make_drop(t);
  p.a0 = saved + c.a0; p.gv = saved + c.gv; p.sv = saved + c.sv;
  p.gpart = ws + c.gpart; p.spart = ws + c.spart; p.cpart = ws + c.cpart; p.u = ws + c.u; p.tt = ws + c.tt; p.dA = ws + c.dA;
  p.dw0part = ws + c.dw0part; p.pch = ws + c.pch;
  return APA_OK;
}

#define APA_CUDA(expr)                       \
  do {                                       \
    cudaError_t e__ = (expr);                \
    if (e__ != cudaSuccess) return int(e__); \
  } while (0)

// every kernel here starts with griddepcontrol.launch_dependents + griddepcontrol.wait: launched as programmatic
// dependents, the launch latency of kernel n+1 hides behind kernel n
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <typename Kern>
cudaError_t big_smem(Kern k, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
}

int check_train(const ApaTrain* t) {
  if (t == nullptr) return APA_OK;
  if (!(t->dropout_keep > 0.f) || t->dropout_keep > 1.f || t->reserved0 != 0) return APA_ERR_UNSUPPORTED;
  return APA_OK;
}

}  // namespace
}  // namespace apa

using namespace apa;

extern "C" {

size_t apa_attn_pool_train_saved_bytes(const ApaDesc* desc) {
  if (train_supported(desc) != APA_OK) return 0;
  return carve(desc).saved_total * sizeof(float);
}

size_t apa_attn_pool_train_workspace(const ApaDesc* desc) {
  if (train_supported(desc) != APA_OK) return 0;
  return carve(desc).ws_total * sizeof(float);
}

int apa_attn_pool_fwd_train(const ApaDesc* desc, const float* X, const ApaWeights* weights, const ApaTrain* train,
                            float* logits, void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes,
                            void* stream_) {
  int st = train_supported(desc);
  if (st != APA_OK) return st;
  if ((st = check_train(train)) != APA_OK) return st;
  if (!X || !weights || !logits || !saved || !workspace) return APA_ERR_NULL;
  const Carve c = carve(desc);
  if (saved_bytes < c.saved_total * 4 || workspace_bytes < c.ws_total * 4) return APA_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(saved) & 255u) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return APA_ERR_WORKSPACE;
  TrainParams p{};
  if ((st = fill(p, desc, weights, train, X, static_cast<float*>(saved), static_cast<float*>(workspace))) != APA_OK) return st;
  p.logits = logits;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long rows = long(p.N) * p.HW;
  const int V = p.N / p.F;
  const bool fused = p.R == 1 && p.mode != MODE_SOFTMAX && (p.C / 4) <= kPoolThreads;
  if (fused) {
    APA_CUDA(launch_pdl(train_fused_fwd_kernel, dim3(std::min(p.N * p.J, kSMs)), dim3(kPoolThreads), 0, stream, p));
  } else {
    const int grid = int(std::min<long>((rows + 7) / 8, 148 * 8));
    APA_CUDA(launch_pdl(train_attn_kernel, dim3(grid), dim3(256), size_t(p.C) * 4, stream, p));
    const size_t smem = (size_t(p.R) * p.HW + 32) * 4;
    if (p.R == 1) {
      APA_CUDA(big_smem(train_pool_kernel<1>, smem));
      APA_CUDA(launch_pdl(train_pool_kernel<1>, dim3(p.J, p.N), dim3(kPoolThreads), smem, stream, p));
    } else {
      APA_CUDA(big_smem(train_pool_kernel<APA_MAX_RANK>, smem));
      APA_CUDA(launch_pdl(train_pool_kernel<APA_MAX_RANK>, dim3(p.J, p.N), dim3(kPoolThreads), smem, stream, p));
    }
  }
  APA_CUDA(launch_pdl(train_reduce_kernel, dim3((p.R * (p.C / 4) + 255) / 256, V), dim3(256), 0, stream, p));
  {
    const int cs = (((p.C + kSlices - 1) / kSlices) + 3) & ~3;
    const size_t smem = size_t(cs) * 64 * 4;
    APA_CUDA(big_smem(train_contract_kernel, smem));
    APA_CUDA(launch_pdl(train_contract_kernel, dim3((p.K + 31) / 32, kSlices, (V + 31) / 32), dim3(256), smem, stream, p));
    APA_CUDA(launch_pdl(train_finish_kernel, dim3((V * p.K + 255) / 256), dim3(256), 0, stream, p));
  }
  APA_CUDA(cudaGetLastError());
  return APA_OK;
}

int apa_attn_pool_bwd(const ApaDesc* desc, const float* X, const ApaWeights* weights, const ApaTrain* train,
                      const float* dlogits, const void* saved, size_t saved_bytes, const ApaGrads* grads,
                      void* workspace, size_t workspace_bytes, void* stream_) {
  int st = train_supported(desc);
  if (st != APA_OK) return st;
  if ((st = check_train(train)) != APA_OK) return st;
  if (!X || !weights || !dlogits || !saved || !grads || !workspace) return APA_ERR_NULL;
  const Carve c = carve(desc);
  if (saved_bytes < c.saved_total * 4 || workspace_bytes < c.ws_total * 4) return APA_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(saved) & 255u) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return APA_ERR_WORKSPACE;
  TrainParams p{};
  if ((st = fill(p, desc, weights, train, X, const_cast<float*>(static_cast<const float*>(saved)),
                 static_cast<float*>(workspace))) != APA_OK) return st;
  p.dlogits = dlogits;
  p.dX = grads->dX;
  if (p.dX && (reinterpret_cast<uintptr_t>(p.dX) & 15u)) return APA_ERR_ALIGN;
  for (int r = 0; r < p.R; ++r) {
    if (!grads->attn_w[r] || !grads->attn_b[r] || !grads->topdown_w[r] || !grads->topdown_b[r]) return APA_ERR_NULL;
    p.d_aw[r] = grads->attn_w[r]; p.d_ab[r] = grads->attn_b[r]; p.d_tw[r] = grads->topdown_w[r]; p.d_tb[r] = grads->topdown_b[r];
  }
  if (reinterpret_cast<uintptr_t>(p.d_aw[0]) & 15u) return APA_ERR_ALIGN;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int V = p.N / p.F;
  const bool fused = p.R == 1 && p.mode != MODE_SOFTMAX && (p.C / 4) <= kPoolThreads;
  p.parts = fused ? std::min(p.N * p.J, kSMs) : p.N * p.J;
  {
    const int cx = (p.C + 31) / 32, vy = (V + 7) / 8, nU = cx * vy * p.R, nW = cx * p.R;
    const size_t smem = std::max((size_t(32) * (p.K | 1) + size_t(8) * ((p.K + 3) & ~3)) * 4, size_t(36) * 1024);
    APA_CUDA(big_smem(bwd_small_kernel, smem));
    APA_CUDA(launch_pdl(bwd_small_kernel, dim3(nU + nW), dim3(256), smem, stream, p, nW, cx, vy));
  }
  if (fused) {
    APA_CUDA(launch_pdl(bwd_fused_kernel, dim3(p.parts), dim3(kPoolThreads), 0, stream, p));
  } else {
    APA_CUDA(launch_pdl(bwd_dattn_kernel, dim3(p.J, p.N), dim3(256), 0, stream, p));
    const size_t smem = (size_t(2) * p.R * p.HW + p.HW + 32) * 4;
    if (p.R == 1) {
      APA_CUDA(big_smem(bwd_dx_kernel<1>, smem));
      APA_CUDA(launch_pdl(bwd_dx_kernel<1>, dim3(p.J, p.N), dim3(kPoolThreads), smem, stream, p));
    } else {
      APA_CUDA(big_smem(bwd_dx_kernel<APA_MAX_RANK>, smem));
      APA_CUDA(launch_pdl(bwd_dx_kernel<APA_MAX_RANK>, dim3(p.J, p.N), dim3(kPoolThreads), smem, stream, p));
    }
  }
  {
    const int parts = p.parts, groups = sum_groups(parts), per = (parts + groups - 1) / groups;
    float* tmp = p.dw0part + size_t(p.N) * p.J * p.C;
    const float* src = p.dw0part;
    APA_CUDA(launch_pdl(bwd_sum_kernel, dim3((p.C / 4 + 31) / 32, groups), dim3(256), 0, stream, p, src, tmp, parts, per, 0));
    const float* tmpc = tmp;
    APA_CUDA(launch_pdl(bwd_sum_kernel, dim3((p.C / 4 + 31) / 32, 1), dim3(256), 0, stream, p, tmpc, p.d_aw[0], groups, groups, 1));
  }
  APA_CUDA(cudaGetLastError());
  return APA_OK;
}

int apa_dropout_mask(const ApaTrain* train, size_t count, float* mask, void* stream) {
  int st = check_train(train);
  if (st != APA_OK) return st;
  if (!mask) return APA_ERR_NULL;
  if (count % 4 != 0) return APA_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(mask) & 15u) return APA_ERR_ALIGN;
  const size_t n4 = count / 4;
  if (n4 == 0) return APA_OK;
  dropout_mask_kernel<<<unsigned((n4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(make_drop(train), mask, n4);
  APA_CUDA(cudaGetLastError());
  return APA_OK;
}

}  // extern "C"

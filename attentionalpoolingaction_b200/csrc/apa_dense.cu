// apa_dense.cu — the dense form of the head, for the flag sets the re-associated kernel cannot
// serve: PER_CLASS attention (M == K), materialised end-points ('TopDownAttention',
// 'PoseLogits'), the 2048->768 pose pre-logits of config 003, and the paper's full Gram.
//
// Reference sites: nets_factory.py:151-160 (pose convs), :259-274 (attention convs incl. the
// K x K chain conv when PER_CLASS), :276-286 (softmax / relu), :299-309 (top-down convs),
// :322-328 (multiply, spatial mean, rank sum), :354-374 (frame mean).
//
// dense_gemm dispatches to the tcgen05 3xTF32 kernel (apa_gemm_tc.cu) when the problem is big enough;
// the fp32 SIMT kernel below stays for tiny shapes (N == 1 attention conv, chain convs) and as the A/B
// reference (APA_DENSE_SIMT=1).
#include <stdlib.h>

#include "apa_common.cuh"
#include "apa_kernels.h"

namespace apa {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

// Y[m, col_off + n*col_stride] = act(sum_k X[m,k] W[k,n] + bias[n]);  256 threads, 4x4 per thread.
__global__ void __launch_bounds__(256) dense_gemm_kernel(const float* __restrict__ X, int ldx, int xs,
                                                         const float* __restrict__ W, int ldw,
                                                         const float* __restrict__ bias,
                                                         float* __restrict__ Y, int ldy, int col_stride,
                                                         int col_off, int M, int N, int Kd, int relu) {
  __shared__ float Xs[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Kd; k0 += TK) {
    // X tile: 64 rows x 16 k  (thread -> row = tid/4, 4 consecutive k)
    {
      const int r = tid >> 2, kk = (tid & 3) * 4;
      const int m = m0 + r;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int k = k0 + kk + t;
        Xs[kk + t][r] = (m < M && k < Kd) ? __ldg(X + size_t(m) * ldx + size_t(k) * xs) : 0.f;
      }
    }
    // W tile: 16 k x 64 n  (thread -> k = tid/16, 4 consecutive n)
    {
      const int kk = tid >> 4, nn = (tid & 15) * 4;
      const int k = k0 + kk;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int n = n0 + nn + t;
        Ws[kk][nn + t] = (k < Kd && n < N) ? __ldg(W + size_t(k) * ldw + n) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Xs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      Y[size_t(m) * ldy + col_off + size_t(n) * col_stride] = v;
    }
  }
}

// attn_out[b,hw,r] from raw a_0: chain (Q1), then softmax / relu.
__global__ void attn_finalize_m1_kernel(const float* __restrict__ a0, const float* w1, const float* b1,
                                        const float* w2, const float* b2, const float* w3,
                                        const float* b3, const float* __restrict__ smx,
                                        float* __restrict__ out, int B, int HW, int R, int mode) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= size_t(B) * HW) return;
  const int b = int(i / HW);
  float a = a0[i];
  if (mode == MODE_SOFTMAX) {
    out[i] = expf(a - smx[2 * b]) / smx[2 * b + 1];
    return;
  }
  const float* cw[3] = {w1, w2, w3};
  const float* cb[3] = {b1, b2, b3};
  for (int r = 0; r < R; ++r) {
    if (r >= 1) a = fmaf(__ldg(cw[r - 1]), a, __ldg(cb[r - 1]));
    out[i * R + r] = (mode == MODE_RELU) ? fmaxf(a, 0.f) : a;
  }
}

// in-place softmax over hw for A[b, hw, m]; one thread per (b, m), coalesced over m.
__global__ void softmax_hw_kernel(float* __restrict__ A, int HW, int M) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float* base = A + size_t(blockIdx.y) * HW * M + m;
  float mx = -INFINITY;
  for (int h = 0; h < HW; ++h) mx = fmaxf(mx, base[size_t(h) * M]);
  float s = 0.f;
  for (int h = 0; h < HW; ++h) s += expf(base[size_t(h) * M] - mx);
  for (int h = 0; h < HW; ++h) base[size_t(h) * M] = expf(base[size_t(h) * M] - mx) / s;
}

__global__ void relu_kernel(float* __restrict__ A, size_t n) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) A[i] = fmaxf(A[i], 0.f);
}

// logits[v,k] = 1/(HW F) sum_f sum_hw sum_r A[b,hw,m(k),r] T[b,hw,k,r]   (b = v F + f)
// grid (ceil(K/128), Bv, hw-splits) with a smem-free two-step: each block reduces a slice of hw
// and atomically... no: deterministic — one block per (k-tile, v) loops all hw.
__global__ void __launch_bounds__(128) dense_pool_kernel(const float* __restrict__ A,
                                                         const float* __restrict__ T,
                                                         float* __restrict__ logits, int HW, int K, int M,
                                                         int R, int F, float inv) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int v = blockIdx.y;
  const int m = (M == 1) ? 0 : k;
  float acc = 0.f;
  for (int f = 0; f < F; ++f) {
    const size_t b = size_t(v) * F + f;
    for (int h = 0; h < HW; ++h) {
      const float* a = A + ((b * HW + h) * M + m) * R;
      const float* t = T + ((b * HW + h) * K + k) * R;
      for (int r = 0; r < R; ++r) acc = fmaf(a[r], t[r], acc);
    }
  }
  logits[size_t(v) * K + k] = acc * inv;
}

// G[b][i][j] = sum_hw X[b][hw][i] X[b][hw][j]; 64x64 output tile per block, 4x4 per thread.
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ X, float* __restrict__ G,
                                                   int HW, int C) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  const float* Xb = X + size_t(blockIdx.z) * HW * C;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int h0 = 0; h0 < HW; h0 += TK) {
    const int hh = tid >> 4, cc = (tid & 15) * 4;
    const int h = h0 + hh;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      As[hh][cc + t] = (h < HW && i0 + cc + t < C) ? __ldg(Xb + size_t(h) * C + i0 + cc + t) : 0.f;
      Bs[hh][cc + t] = (h < HW && j0 + cc + t < C) ? __ldg(Xb + size_t(h) * C + j0 + cc + t) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* Gb = G + size_t(blockIdx.z) * C * C;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = i0 + ty * 4 + i, gj = j0 + tx * 4 + j;
      if (gi < C && gj < C) Gb[size_t(gi) * C + gj] = acc[i][j];
    }
}

// ---- 1-warp kernels of the fused all-gather protocol (apa_common.cuh, include/apa.h)
struct PeerParams {
  unsigned long long* flags[APA_MAX_PEERS];
  unsigned long long* counters;
  int world, rank, depth, lag;
};
// Waits (one thread) until the data flags of all `world` ranks in MY flag array have reached (latest step - lag).
// Launched as a programmatic dependent of the head kernel: triggers its own dependents at once (the next head
// launch may start streaming), then waits for the head kernel to finish (stream order) before reading the count.
__global__ void peer_gather_wait_kernel(const PeerParams w) {
  griddep_launch_dependents();
  griddep_wait();
  if (threadIdx.x != 0) return;
  const unsigned long long latest = peer_gather_step(w.counters) - 1;   // calls completed on this rank
  if (latest > (unsigned long long)w.lag) peer_gather_wait_all(w.world, w.flags[w.rank], 0, latest - w.lag);
}
// Everything before this kernel in the stream has completed: tell every peer that this rank has started `step`
// (it no longer reads the buffers it had to consume before this call), then wait until this step's ring buffer
// is free on every peer.
__global__ void peer_gather_begin_kernel(const PeerParams w) {
  if (threadIdx.x != 0) return;
  const unsigned long long step = peer_gather_step(w.counters);
  peer_gather_publish_started(w.world, w.rank, w.flags, step);
  if (step >= (unsigned long long)w.depth) peer_gather_wait_all(w.world, w.flags[w.rank], w.world, step - w.depth + 1);
}

PeerParams peer_params(const ApaPeerGather* gw, int lag) {
  PeerParams w;
  for (int q = 0; q < APA_MAX_PEERS; ++q) w.flags[q] = q < gw->world ? gw->peer_flags[q] : nullptr;
  w.counters = gw->counters;
  w.world = gw->world; w.rank = gw->rank; w.depth = gw->depth; w.lag = lag;
  return w;
}

}  // namespace

cudaError_t peer_gather_begin_launch(const ApaPeerGather* gw, cudaStream_t stream) {
  peer_gather_begin_kernel<<<1, 32, 0, stream>>>(peer_params(gw, 0));
  return cudaGetLastError();
}

cudaError_t peer_gather_wait_launch(const ApaPeerGather* gw, int lag, bool pdl, cudaStream_t stream) {
  const PeerParams w = peer_params(gw, lag);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1, 1, 1);
  cfg.blockDim = dim3(32, 1, 1);
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t err = cudaLaunchKernelEx(&cfg, peer_gather_wait_kernel, w);
  if (err != cudaSuccess && pdl) {
    (void)cudaGetLastError();
    cfg.numAttrs = 0;
    err = cudaLaunchKernelEx(&cfg, peer_gather_wait_kernel, w);
  }
  return err;
}

bool dense_force_simt() {
  static const bool force_simt = (getenv("APA_DENSE_SIMT") != nullptr);
  return force_simt;
}

cudaError_t dense_gemm(const float* X, int ldx, int xs, const float* W, int ldw, const float* bias,
                       float* Y, int ldy, int col_stride, int col_off, int M, int N, int Kd, int relu,
                       cudaStream_t stream) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (!dense_force_simt() && dense_gemm_tc_eligible(M, N, Kd))
    return dense_gemm_tc(X, ldx, xs, W, ldw, bias, Y, ldy, col_stride, col_off, M, N, Kd, relu, stream);
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, 1);
  dense_gemm_kernel<<<grid, 256, 0, stream>>>(X, ldx, xs, W, ldw, bias, Y, ldy, col_stride, col_off, M, N, Kd,
                                              relu);
  return cudaGetLastError();
}

cudaError_t attn_finalize_m1(const float* a0, const float* const* attn_w, const float* const* attn_b,
                             const float* smx, float* attn_out, int B, int HW, int R, int mode,
                             cudaStream_t stream) {
  const size_t n = size_t(B) * HW;
  attn_finalize_m1_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(
      a0, attn_w[1], attn_b[1], attn_w[2], attn_b[2], attn_w[3], attn_b[3], smx, attn_out, B, HW, R, mode);
  return cudaGetLastError();
}

cudaError_t dense_softmax_hw(float* A, int B, int HW, int M, cudaStream_t stream) {
  dim3 grid((M + 127) / 128, B, 1);
  softmax_hw_kernel<<<grid, 128, 0, stream>>>(A, HW, M);
  return cudaGetLastError();
}

cudaError_t dense_relu(float* A, size_t n, cudaStream_t stream) {
  relu_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(A, n);
  return cudaGetLastError();
}

cudaError_t dense_pool(const float* A, const float* T, float* logits, int B, int HW, int K, int M, int R,
                       int F, cudaStream_t stream) {
  dim3 grid((K + 127) / 128, B / F, 1);
  dense_pool_kernel<<<grid, 128, 0, stream>>>(A, T, logits, HW, K, M, R, F, 1.0f / (float(HW) * float(F)));
  return cudaGetLastError();
}

cudaError_t gram_launch(int B, int HW, int C, const float* X, float* G, cudaStream_t stream) {
  if (!dense_force_simt() && C >= 64 && HW >= 16) return gram_tc(B, HW, C, X, G, stream);
  dim3 grid((C + TN - 1) / TN, (C + TM - 1) / TM, B);
  gram_kernel<<<grid, 256, 0, stream>>>(X, G, HW, C);
  return cudaGetLastError();
}

}  // namespace apa

// apa_eval.cu — the evaluation tail that follows the head.
//
// Reference sites: eval.py:193-197 (predictions = argmax(logits, 1); logits = softmax(logits) or sigmoid for the
// multi-label losses), :200-206 (streaming accuracy), :303-306 (accuracy and mAP over all collected scores);
// eval/utils.py:4-16 (compute_map: one-vs-rest AP per class, classes without positives skipped, mean);
// eval/cap_eval_utils.py:55-88 (calc_pr_ovr_noref: sort by score descending, cumulative precision / recall) and
// :91-109 (voc_ap: precision envelope from the right, area over the recall steps).
//
// Two kernels: apa_eval_tail_kernel (warp per row: max / first argmax / softmax or sigmoid / hit count) and
// apa_class_ap_kernel (CTA per class: bitonic sort of (score, index) keys in shared memory, block scans in float64).
#include <stdint.h>

#include "../../include/apa.h"
#include "apa_common.cuh"

namespace apa {
namespace {

__global__ void __launch_bounds__(256) apa_eval_tail_kernel(int rows, int K, const float* __restrict__ logits, int activation,
                                                             float* __restrict__ probs, int32_t* __restrict__ pred,
                                                             const int32_t* __restrict__ labels,
                                                             unsigned long long* __restrict__ counts) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const float* x = logits + size_t(row) * K;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    const float v = x[k];
    if (v > mx) { mx = v; arg = k; }   // strict: the first maximum of this lane's (increasing) indices
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }   // ties: smallest index, like tf.argmax / np.argmax
  }
  if (probs) {
    float* y = probs + size_t(row) * K;
    if (activation == APA_EVAL_SIGMOID) {
      for (int k = lane; k < K; k += 32) y[k] = 1.f / (1.f + expf(-x[k]));
    } else {
      float s = 0.f;
      for (int k = lane; k < K; k += 32) s += expf(x[k] - mx);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float inv = 1.f / s;
      for (int k = lane; k < K; k += 32) y[k] = expf(x[k] - mx) * inv;
    }
  }
  if (lane == 0) {
    if (pred) pred[row] = arg;
    if (labels && counts) {
      atomicAdd(counts + 0, (unsigned long long)(labels[row] == arg ? 1 : 0));   // integer adds: order-free
      atomicAdd(counts + 1, 1ull);
    }
  }
}

// float -> uint32 whose unsigned order is the float order (-inf < ... < -0 < +0 < ... < +inf; NaNs at the ends)
__device__ __forceinline__ uint32_t orderable(float f) {
  const uint32_t u = __float_as_uint(f + 0.f);   // -0 -> +0: the two compare equal on the CPU (a tie)
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int kApThreads = 1024;

// One class per CTA: keys (orderable score << 32 | index) sorted ascending in shared memory, then walked from the top
// (score descending; equal scores by index descending, which is what a stable argsort read backwards gives,
// cap_eval_utils.py:70-71).  P[i] = tp_cum / (i + 1), envelope from the right, AP = sum over the positives of the
// envelope / number of positives (voc_ap's recall steps are exactly the positives).
__global__ void __launch_bounds__(kApThreads) apa_class_ap_kernel(int n, int npad, int K, const float* __restrict__ scores,
                                                                  const int32_t* __restrict__ labels, double* __restrict__ ap) {
  extern __shared__ unsigned long long keys[];   // [npad]
  __shared__ double redd[32];
  __shared__ int redi[32];
  const int cid = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < npad; i += kApThreads)
    keys[i] = (i < n) ? ((unsigned long long)orderable(scores[size_t(i) * K + cid]) << 32) | (unsigned)i : 0ull;   // pads sort first
  __syncthreads();
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npad; i += kApThreads) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], b = keys[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  // rank q = 0 .. n-1 (descending) is keys[npad - 1 - q]; each thread owns a run of consecutive ranks
  const int per = (n + kApThreads - 1) / kApThreads, q0 = min(n, tid * per), q1 = min(n, q0 + per);
  int cnt = 0;
  for (int q = q0; q < q1; ++q) cnt += labels[int(keys[npad - 1 - q] & 0xffffffffu)] == cid ? 1 : 0;
  // exclusive scan of the runs' positive counts
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) redi[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = redi[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    redi[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  const int before = incl - cnt + (warp > 0 ? redi[warp - 1] : 0);
  const int npos = redi[31];
  if (npos == 0) {   // eval/utils.py:10-12: the class is ignored
    if (tid == 0) ap[cid] = -1.0;
    return;
  }
  // the run's largest precision, then the suffix maximum over the runs to the right
  double runmax = 0.0;
  {
    int tp = before;
    for (int q = q0; q < q1; ++q) {
      tp += labels[int(keys[npad - 1 - q] & 0xffffffffu)] == cid ? 1 : 0;
      runmax = fmax(runmax, double(tp) / double(q + 1));
    }
  }
  double suf = runmax;   // inclusive suffix max within the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf = fmax(suf, t);
  }
  if (lane == 0) redd[warp] = suf;
  __syncthreads();
  if (warp == 0) {
    double w = redd[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_down_sync(0xffffffffu, w, o);
      if (lane + o < 32) w = fmax(w, t);
    }
    redd[lane] = w;   // max over warps lane .. 31
  }
  __syncthreads();
  double right = (warp < 31) ? redd[warp + 1] : 0.0;   // everything in the warps to the right
  {
    const double t = __shfl_down_sync(0xffffffffu, suf, 1);   // the threads to the right inside the warp
    if (lane < 31) right = fmax(right, t);
  }
  // walk the run backwards with the running envelope, adding it at every positive
  double sum = 0.0;
  {
    int tp = before + cnt;
    double env = right;
    for (int q = q1 - 1; q >= q0; --q) {
      const bool pos = labels[int(keys[npad - 1 - q] & 0xffffffffu)] == cid;
      env = fmax(env, double(tp) / double(q + 1));
      if (pos) { sum += env; --tp; }
    }
  }
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) redd[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += redd[w];
    ap[cid] = t / double(npos);
  }
}

// mean over the classes that have positives (eval/utils.py:15)
__global__ void apa_mean_ap_kernel(int K, const double* __restrict__ ap, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    int c = 0;
    for (int k = 0; k < K; ++k)
      if (ap[k] >= 0.0) { t += ap[k]; ++c; }
    out[0] = c ? t / c : 0.0;
    out[1] = double(c);
  }
}

#define APA_CUDA(expr)                       \
  do {                                       \
    cudaError_t e__ = (expr);                \
    if (e__ != cudaSuccess) return int(e__); \
  } while (0)

}  // namespace
}  // namespace apa

using namespace apa;

extern "C" {

int apa_eval_tail(int32_t rows, int32_t num_classes, const float* logits, int32_t activation, float* probs,
                  int32_t* predictions, const int32_t* labels, unsigned long long* counts, void* stream) {
  if (rows < 0 || num_classes <= 0) return APA_ERR_SHAPE;
  if (activation != APA_EVAL_SOFTMAX && activation != APA_EVAL_SIGMOID) return APA_ERR_UNSUPPORTED;
  if (!logits) return APA_ERR_NULL;
  if ((labels == nullptr) != (counts == nullptr)) return APA_ERR_NULL;
  if (rows == 0) return APA_OK;
  apa_eval_tail_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, num_classes, logits, activation,
                                                                                    probs, predictions, labels, counts);
  APA_CUDA(cudaGetLastError());
  return APA_OK;
}

int apa_mean_ap(int32_t samples, int32_t num_classes, const float* scores, const int32_t* labels, double* class_ap,
                double* mean_ap, void* stream) {
  if (samples <= 0 || num_classes <= 0) return APA_ERR_SHAPE;
  if (samples > APA_AP_MAX_SAMPLES) return APA_ERR_UNSUPPORTED;
  if (!scores || !labels || !class_ap || !mean_ap) return APA_ERR_NULL;
  int npad = 2;
  while (npad < samples) npad <<= 1;
  const size_t smem = size_t(npad) * sizeof(unsigned long long);
  if (smem > 48 * 1024)
    APA_CUDA(cudaFuncSetAttribute(apa_class_ap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  apa_class_ap_kernel<<<num_classes, kApThreads, smem, st>>>(samples, npad, num_classes, scores, labels, class_ap);
  apa_mean_ap_kernel<<<1, 32, 0, st>>>(num_classes, class_ap, mean_ap);
  APA_CUDA(cudaGetLastError());
  return APA_OK;
}

}  // extern "C"

// apa_loss.cu — the two losses that sit behind the head in training, each with the
// gradient that the backward of the head (apa_attn_pool_bwd) or of the pose branch starts from.
//
// Reference sites: loss.py:72-79 (`tf.losses.softmax_cross_entropy(onehot(labels), logits, weights=action_loss_wt)`:
// TF 1.1's compute_weighted_loss = sum(w * loss_i) / number of rows for a scalar w != 0) and loss.py:29-70, the
// un-sampled branch (cfg.TRAIN.LOSS_FN_POSE_SAMPLED = False): per key-point channel c,
// 0.5 * sum_hw (label - logit)^2 / reduce_sum(ones([B, H, W])) per image (the divisor counts the batch too, :55-57), kept
// where labels_pose_valid[b, c], mean over the batch, summed over the channels, times pose_loss_wt.  (The sampled branch draws a random mask and a differing label size is resized
// there: both are outside this entry point -- pass labels of the logits' size.)
//
// Per-row losses go to a caller-provided buffer and a single CTA adds them in index order: bit-reproducible.
#include <stdint.h>

#include "../../include/apa.h"
#include "apa_common.cuh"

namespace apa {
namespace {

__global__ void __launch_bounds__(256) xent_rows_kernel(int rows, int K, const float* __restrict__ logits,
                                                        const int32_t* __restrict__ labels, float gscale,
                                                        float* __restrict__ row_loss, float* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const float* x = logits + size_t(row) * K;
  float mx = -INFINITY;
  for (int k = lane; k < K; k += 32) mx = fmaxf(mx, x[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += expf(x[k] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int lab = labels[row];
  if (lane == 0) row_loss[row] = (lab >= 0 && lab < K) ? (logf(s) + mx - x[lab]) : 0.f;
  if (dlogits) {
    const float inv = 1.f / s;
    float* d = dlogits + size_t(row) * K;
    for (int k = lane; k < K; k += 32) d[k] = gscale * (expf(x[k] - mx) * inv - (k == lab ? 1.f : 0.f));
  }
}

// one CTA per image: sum over the valid key-point channels of sum_hw (logit - label)^2; gradient on request
__global__ void __launch_bounds__(256) pose_rows_kernel(int HW, int Kp, const float* __restrict__ logits,
                                                        const float* __restrict__ labels, const uint8_t* __restrict__ valid,
                                                        float gscale, float* __restrict__ row_loss, float* __restrict__ dpose) {
  __shared__ float red[8];
  const int b = blockIdx.x, n = HW * Kp;
  const float* x = logits + size_t(b) * n;
  const float* y = labels + size_t(b) * n;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % Kp;
    const float ok = valid[size_t(b) * Kp + c] ? 1.f : 0.f;
    const float d = (x[i] - y[i]) * ok;
    acc += d * d;
    if (dpose) dpose[size_t(b) * n + i] = gscale * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    row_loss[b] = t;
  }
}

// loss = scale * sum of the rows, in index order (lanes stride, then a fixed butterfly, then the warps in order)
__global__ void __launch_bounds__(256) sum_rows_kernel(int rows, const float* __restrict__ row_loss, float scale,
                                                       float* __restrict__ loss) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) acc += row_loss[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    loss[0] = t * scale;
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

int apa_softmax_xent_loss(int32_t rows, int32_t num_classes, const float* logits, const int32_t* labels, float weight,
                          float* row_loss, float* loss, float* dlogits, void* stream) {
  if (rows <= 0 || num_classes <= 0) return APA_ERR_SHAPE;
  if (!logits || !labels || !row_loss || !loss) return APA_ERR_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float scale = weight / float(rows);   // compute_weighted_loss with a scalar weight: sum(w l_i) / rows
  xent_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(rows, num_classes, logits, labels, scale, row_loss, dlogits);
  sum_rows_kernel<<<1, 256, 0, st>>>(rows, row_loss, scale, loss);
  APA_CUDA(cudaGetLastError());
  return APA_OK;
}

int apa_pose_l2_loss(int32_t batch, int32_t hw, int32_t keypoints, const float* pose_logits, const float* pose_labels,
                     const uint8_t* valid, float weight, float* row_loss, float* loss, float* dpose, void* stream) {
  if (batch <= 0 || hw <= 0 || keypoints <= 0) return APA_ERR_SHAPE;
  if (!pose_logits || !pose_labels || !valid || !row_loss || !loss) return APA_ERR_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // loss.py:56-57 divides the per-image sum by tf.reduce_sum(mask) with mask = ones([B, H, W]) -- B*HW, not HW -- and
  // :60-64 then takes the mean over the batch: the divisor is B * B * HW (reproduced, not corrected)
  const float scale = 0.5f * weight / (float(hw) * float(batch) * float(batch));
  pose_rows_kernel<<<batch, 256, 0, st>>>(hw, keypoints, pose_logits, pose_labels, valid, 2.f * scale, row_loss, dpose);
  sum_rows_kernel<<<1, 256, 0, st>>>(batch, row_loss, scale, loss);
  APA_CUDA(cudaGetLastError());
  return APA_OK;
}

}  // extern "C"

// apa_kernels.h — internal (non-ABI) interface between the C-ABI dispatcher and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/apa.h"

namespace apa {

enum { MODE_PLAIN = 0, MODE_RELU = 1, MODE_SOFTMAX = 2 };

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = 32 * kConsumerWarps;
constexpr int kFusedThreads = kConsumerThreads + 32;  // + 1 bulk-copy producer warp
constexpr int kVideoTile = 32;  // videos per tile of the contraction

// Parameters of the single-map (M == 1) head kernels (apa_pool_flow.cu, apa_pool_cluster.cu).
struct FusedParams {
  const float* X;        // [B, HW, C]
  const float* attn_in;  // optional precomputed a_0 as attn_in_parts partial sums [parts][B*HW] (attention
                         // source is not X): a_0 = attn_b[0] * attn_in_bias + sum_p attn_in[p]; else nullptr
  int attn_in_parts, attn_in_bias;
  const float* attn_w[APA_MAX_RANK];  // r==0: [C] (unused if attn_in); r>=1: scalar
  const float* attn_b[APA_MAX_RANK];  // scalars
  const float* td_w[APA_MAX_RANK];    // [C, K]
  const float* td_b[APA_MAX_RANK];    // [K]
  float* logits;                      // [B/F, K]
  float* attn_raw;                    // optional [B*HW]: a_0 before chain/relu/softmax
  float* smx;                         // workspace [B, 2]: softmax (max, sum) per image (attention end point)
  unsigned long long* ts;             // optional debug timeline, 64 words per CTA
  int B, HW, C, K, R, F;
  int trigger_lead;                   // PDL: trigger the next launch this many tiles before a CTA's last tile
  int inputs_stable;                  // X / weights may be read before the previous kernel in the stream completed
  float inv_scale;                    // 1 / (HW * F)
  // fused all-gather over peer memory (gw_world == 0: plain store to `logits`)
  int gw_world, gw_rank, gw_depth;    // gw_depth: buffers in every rank's gathered ring (call `step` lands in step % depth)
  float* gw_logits[APA_MAX_PEERS];
  unsigned long long* gw_flags[APA_MAX_PEERS];
  unsigned long long* gw_done;        // the gather object's local counters (include/apa.h ApaPeerGather.counters)
  // flow form (apa_pool_flow.cu)
  float* gpart;                       // workspace [copy][slot][R][C]: partial g per (STREAM CTA, image, consumer group)
  float* gstat;                       // workspace [copy][slot][8]: (s_r[4], running max) per partial
  float* g;                           // workspace [copy][B][R][C]: attention-weighted feature sums per image
  float* gsum;                        // workspace [copy][B][4]: attention mass per rank of the reduced image
  float* cpart;                       // workspace [copy][tiles * class slices][16][32 videos][32 classes] partial logits
  unsigned long long* fl_ctr;         // control words: STREAM / REDUCE / CONTRACT / FINISH tickets, completed calls, gather steps
  size_t fl_gpart_stride, fl_gstat_stride, fl_g_stride, fl_gsum_stride, fl_cpart_stride;
                                      // distance between the kFlowDepth copies of the per-call buffers
  int fl_G, fl_NS, fl_ng;             // STREAM CTAs, class slices, consumer groups (partials per CTA and image)
  int fl_debug_nowait;                // measurement runs of the chain without its tail: skip the completed-calls check
};

// ---- cluster form (apa_pool_cluster.cu): 16-CTA clusters, everything after the stream stays in shared memory
bool cluster_eligible(int B, int HW, int C, int K, int R, int F);
void cluster_plan(int B, int C, int R, int K, const ApaTuning* tuning, int* images_per_cluster, size_t* smem);
cudaError_t cluster_launch(const FusedParams& p, int mode, const ApaTuning* tuning, cudaStream_t stream, int* grid_out);

// ---- flow form (apa_pool_flow.cu): row-balanced STREAM kernel + three small tail kernels parked in griddepcontrol.wait
constexpr int kFlowStreamCtas = 148;   // one STREAM CTA per SM of a B200
constexpr int kFlowDepth = 2;          // copies of the per-call buffers: call L uses copy L % 2
bool flow_eligible(int B, int HW, int C, int K, int R, int F);
int flow_stream_ctas(int B, int HW, const ApaTuning* tuning);
int flow_class_slices(int K);
int flow_contract_grid(int K);
size_t flow_gpart_slots(int B);                 // partial-g slots one copy of the workspace holds
size_t flow_slices(int B, int K, int F);        // (video tile, class slice) pairs
size_t flow_cpart_floats(int B, int K, int F);  // partial-logit floats one copy of the workspace holds
size_t flow_smem(int C, const ApaTuning* tuning);
cudaError_t flow_launch(const FusedParams& p, int mode, const ApaTuning* tuning, cudaStream_t stream,
                        int* stream_grid, int* contract_grid);

// ---- dense path (apa_dense.cu)
bool dense_force_simt();  // APA_DENSE_SIMT=1 in the environment (read once): A/B switch for tests and profiling
// Y[m, col_off + n*col_stride] (row stride ldy) = act(sum_k X[m*ldx + k*xs] W[k,n] + bias[n])
cudaError_t dense_gemm(const float* X, int ldx, int xs, const float* W, int ldw, const float* bias,
                       float* Y, int ldy, int col_stride, int col_off, int M, int N, int Kd, int relu,
                       cudaStream_t stream);
// same contract on tcgen05 tensor cores (3xTF32 split, fp32 accumulate in TMEM); apa_gemm_tc.cu
bool dense_gemm_tc_eligible(int M, int N, int Kd);
// Y may be null; with dot_w/dot_out the epilogue also writes dense_gemm_tc_dot_parts(N) x [M] partial row dots
// sum_n act(y[m,n]) dot_w[n] (one per half N-tile) so a following 1-wide conv never has to re-read Y.
int dense_gemm_tc_dot_parts(int N);
cudaError_t dense_gemm_tc(const float* X, int ldx, int xs, const float* W, int ldw, const float* bias, float* Y,
                          int ldy, int col_stride, int col_off, int M, int N, int Kd, int relu, cudaStream_t stream,
                          const float* dot_w = nullptr, float* dot_out = nullptr);
cudaError_t gram_tc(int B, int HW, int C, const float* X, float* G, cudaStream_t stream);
// attention finalisation for M == 1 given raw a_0 (chain, relu, softmax) -> attn_out [B,HW,(R)]
cudaError_t attn_finalize_m1(const float* a0, const float* const* attn_w, const float* const* attn_b,
                             const float* smx, float* attn_out, int B, int HW, int R, int mode,
                             cudaStream_t stream);
// in-place spatial softmax / relu on A [B, HW, M]
cudaError_t dense_softmax_hw(float* A, int B, int HW, int M, cudaStream_t stream);
cudaError_t dense_relu(float* A, size_t n, cudaStream_t stream);
// logits[v,k] = 1/(HW F) sum_f sum_hw sum_r A[b,hw,m(k),r] T[b,hw,k,r]
cudaError_t dense_pool(const float* A, const float* T, float* logits, int B, int HW, int K, int M,
                       int R, int F, cudaStream_t stream);
// 1-warp kernels of the fused all-gather (include/apa.h ApaPeerGather):
//   begin  (forms without in-kernel flow control) publishes "rank started `step`" to every peer and waits until every
//          peer has started step - depth + 1, i.e. until this step's ring buffer may be overwritten everywhere
//   wait   waits until every rank's data flag in MY flag array has reached `target`
cudaError_t peer_gather_begin_launch(const ApaPeerGather* gw, cudaStream_t stream);
cudaError_t peer_gather_wait_launch(const ApaPeerGather* gw, int lag, bool pdl, cudaStream_t stream);
cudaError_t gram_launch(int B, int HW, int C, const float* X, float* G, cudaStream_t stream);

}  // namespace apa

// apa_api.cu — the C ABI declared in include/apa.h: validation, workspace carving, path selection.
// Host-side only; the kernels live in apa_pool_flow.cu / apa_pool_cluster.cu (single-map head), apa_gemm_tc.cu and
// apa_dense.cu (dense path, Gram).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "apa_kernels.h"

namespace {

using namespace apa;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline int mode_of(int flags) {
  if (flags & APA_FLAG_SOFTMAX_ATT) return MODE_SOFTMAX;  // relu after a softmax is the identity
  if (flags & APA_FLAG_RELU_ATT) return MODE_RELU;
  return MODE_PLAIN;
}

constexpr size_t kCtrlBytes = 4096;         // control block; the flow form's counters start at kFlowCtrOff
constexpr size_t kFlowCtrOff = 2048;
constexpr size_t kTsBytes = 320 * 64 * 8;   // debug timeline: up to 320 CTAs x 64 stamps

struct Carve {
  size_t ts, smx, attn_raw, pose_pre, a0, A, T, gpart, gstat, g, gsum, cpart, total;
  int path;  // 0 single-map fast path, 1 dense
};

// the single-map fast path: one attention map, no materialised top-down end point, a shape the kernels take
bool use_fast_path(const ApaDesc* d, int want_topdown) {
  return !(d->flags & APA_FLAG_PER_CLASS) && !want_topdown &&
         flow_eligible(d->batch, d->hw, d->channels, d->num_classes, d->rank, d->frames);
}

Carve carve(const ApaDesc* d, int want_attn, int want_topdown, int want_pose) {
  Carve c;
  memset(&c, 0, sizeof(c));
  const size_t B = d->batch, HW = d->hw, C = d->channels, K = d->num_classes, R = d->rank;
  const size_t M = (d->flags & APA_FLAG_PER_CLASS) ? K : 1;
  const bool single = (d->flags & APA_FLAG_SINGLE_LAYER_ATT) != 0;
  c.path = use_fast_path(d, want_topdown) ? 0 : 1;
  size_t off = kCtrlBytes;
  auto take = [&](size_t floats) {
    size_t o = off;
    off = align_up(off + floats * sizeof(float), 256);
    return o;
  };
  if (c.path == 0) {
    c.ts = take(kTsBytes / sizeof(float));
    c.smx = take(B * 2);
    if (want_attn) c.attn_raw = take(B * HW);
    if (!single) c.a0 = take(B * HW * 8);  // up to 8 partial-dot planes from the pose GEMM epilogue
    // the flow form keeps kFlowDepth copies of every per-call buffer (call L uses copy L % kFlowDepth)
    const size_t slots = flow_gpart_slots(d->batch);
    c.gpart = take(kFlowDepth * slots * R * C);
    c.gstat = take(kFlowDepth * slots * 8);
    c.g = take(kFlowDepth * B * R * C);
    c.gsum = take(kFlowDepth * B * 4);
    c.cpart = take(kFlowDepth * flow_cpart_floats(d->batch, d->num_classes, d->frames));
  } else {
    if (!want_attn) c.A = take(B * HW * M * R);
    if (!want_topdown) c.T = take(B * HW * K * R);
  }
  if (!single || want_pose) c.pose_pre = take(B * HW * size_t(d->pose_channels));
  c.total = off;
  return c;
}

// APA_FORM_CLUSTER only on request and only where that kernel applies; everything else takes the flow form
bool use_cluster(const ApaDesc* d, const ApaWeights* w, const ApaTuning* t) {
  if (t == nullptr || t->form != APA_FORM_CLUSTER) return false;
  if (!cluster_eligible(d->batch, d->hw, d->channels, d->num_classes, d->rank, d->frames)) return false;
  if (w)
    for (int r = 0; r < d->rank; ++r)
      if (reinterpret_cast<uintptr_t>(w->topdown_w[r]) & 15u) return false;  // its bulk copies of B_r need 16 B alignment
  return true;
}

int check_desc(const ApaDesc* d) {
  if (d == nullptr) return APA_ERR_NULL;
  if (d->abi_version != APA_ABI_VERSION) return APA_ERR_UNSUPPORTED;
  if (d->batch <= 0 || d->hw <= 0 || d->channels <= 0 || d->num_classes <= 0) return APA_ERR_SHAPE;
  if (d->frames <= 0 || d->batch % d->frames != 0) return APA_ERR_SHAPE;
  if (d->channels % 4 != 0) return APA_ERR_SHAPE;
  if (d->rank < 1 || d->rank > APA_MAX_RANK) return APA_ERR_RANK;
  if ((d->flags & APA_FLAG_SOFTMAX_ATT) && d->rank > 1) return APA_ERR_SOFTMAX_RANK;
  if (d->flags & ~(APA_FLAG_SINGLE_LAYER_ATT | APA_FLAG_SOFTMAX_ATT | APA_FLAG_RELU_ATT | APA_FLAG_PER_CLASS))
    return APA_ERR_UNSUPPORTED;
  if (!(d->flags & APA_FLAG_SINGLE_LAYER_ATT) && d->pose_channels <= 0) return APA_ERR_SHAPE;
  for (int i = 0; i < 6; ++i)
    if (d->reserved[i] != 0) return APA_ERR_UNSUPPORTED;
  // 32-bit index headroom inside the kernels
  if (double(d->batch) * d->hw >= 2147483647.0) return APA_ERR_SHAPE;
  return APA_OK;
}

int check_tuning(const ApaTuning* t) {
  if (t == nullptr) return APA_OK;
  if (t->form != APA_FORM_AUTO && t->form != APA_FORM_FLOW && t->form != APA_FORM_CLUSTER) return APA_ERR_UNSUPPORTED;
  if (t->stream_ctas < 0 || t->stream_shape < 0 || t->stream_shape > 2 || t->cluster_shape < 0 || t->cluster_shape > 3)
    return APA_ERR_UNSUPPORTED;
  if (t->tail_smem_kb < 0 || t->tail_smem_kb > 200) return APA_ERR_UNSUPPORTED;
  for (int i = 0; i < 5; ++i)
    if (t->reserved[i] != 0) return APA_ERR_UNSUPPORTED;
  return APA_OK;
}

// sm_100 check, cached per device (immutable after first query).
std::mutex g_dev_mu;
int g_dev_ok[64];  // 0 unknown, 1 ok, -1 not sm_100

int check_device() {
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return int(err);
  if (dev < 0 || dev >= 64) return APA_ERR_DEVICE;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_dev_ok[dev] == 0) {
    int major = 0;
    err = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (err != cudaSuccess) return int(err);
    g_dev_ok[dev] = (major == 10) ? 1 : -1;
  }
  return g_dev_ok[dev] == 1 ? APA_OK : APA_ERR_DEVICE;
}

#define APA_CUDA(expr)                         \
  do {                                         \
    cudaError_t e__ = (expr);                  \
    if (e__ != cudaSuccess) return int(e__);   \
  } while (0)

}  // namespace

extern "C" {

int apa_abi_version(void) { return APA_ABI_VERSION; }

const char* apa_strerror(int status) {
  switch (status) {
    case APA_OK: return "ok";
    case APA_ERR_NULL: return "apa: required pointer is NULL";
    case APA_ERR_SHAPE: return "apa: invalid batch/hw/channels/classes/frames";
    case APA_ERR_RANK: return "apa: rank must be in [1, APA_MAX_RANK]";
    case APA_ERR_SOFTMAX_RANK:
      return "apa: SOFTMAX_ATT with RANK>1 is a graph-build error in the reference (nets_factory.py:278)";
    case APA_ERR_WORKSPACE: return "apa: workspace too small or not 256-byte aligned";
    case APA_ERR_UNSUPPORTED:
      return "apa: unsupported descriptor or tuning (abi version / flags / form / reserved fields)";
    case APA_ERR_ALIGN: return "apa: X and the weight matrices must be 16-byte aligned";
    case APA_ERR_DEVICE: return "apa: current device is not sm_100 (B200)";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(cudaError_t(status));
  return "apa: unknown status";
}

int apa_attn_pool_check(const ApaDesc* desc) { return check_desc(desc); }

size_t apa_attn_pool_workspace(const ApaDesc* desc, int want_attn, int want_topdown, int want_pose) {
  if (check_desc(desc) != APA_OK) return 0;
  return carve(desc, want_attn, want_topdown, want_pose).total;
}

int apa_attn_pool_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
  if (workspace == nullptr) return APA_ERR_NULL;
  if (workspace_bytes < kCtrlBytes || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return APA_ERR_WORKSPACE;
  // every counter of the kernels lives in the workspace and starts from zero
  APA_CUDA(cudaMemsetAsync(workspace, 0, workspace_bytes, static_cast<cudaStream_t>(stream)));
  return APA_OK;
}

int apa_attn_pool_plan(const ApaDesc* desc, const ApaTuning* tuning, int want_attn, int want_topdown,
                       int want_pose, int32_t plan[8]) {
  int st = check_desc(desc);
  if (st != APA_OK) return st;
  st = check_tuning(tuning);
  if (st != APA_OK) return st;
  if (plan == nullptr) return APA_ERR_NULL;
  for (int i = 0; i < 8; ++i) plan[i] = 0;
  const Carve c = carve(desc, want_attn, want_topdown, want_pose);
  plan[0] = c.path;
  const bool single = (desc->flags & APA_FLAG_SINGLE_LAYER_ATT) != 0;
  int launches = 0;
  {  // the PoseLogits scope, as fwd_impl runs it
    const int rows = desc->batch * desc->hw;
    const bool fuse_a0 = !single && c.path == 0 && !want_pose &&
                         dense_gemm_tc_eligible(rows, desc->pose_channels, desc->channels) &&
                         dense_gemm_tc_dot_parts(desc->pose_channels) <= 8 && !dense_force_simt();
    if (fuse_a0) launches += 1;                                   // pose GEMM with the row-dot epilogue
    else if (!single || want_pose) launches += 1 + (want_pose ? 1 : 0) + ((!single && c.path == 0) ? 1 : 0);
  }
  if (c.path == 0) {
    if (use_cluster(desc, nullptr, tuning)) {
      int li = 4;
      size_t lsm = 0;
      cluster_plan(desc->batch, desc->channels, desc->rank, desc->num_classes, tuning, &li, &lsm);
      const int ncl = desc->batch / li < 9 ? desc->batch / li : 9;  // 148 SMs / 16
      plan[1] = APA_FORM_CLUSTER;
      plan[2] = 16 * ncl;
      plan[4] = int32_t(lsm);
      launches += 1;
    } else {
      plan[1] = APA_FORM_FLOW;
      plan[2] = flow_stream_ctas(desc->batch, desc->hw, tuning);
      plan[4] = int32_t(flow_smem(desc->channels, tuning));
      plan[5] = flow_contract_grid(desc->num_classes);
      launches += 4;  // STREAM, REDUCE, CONTRACT, FINISH
    }
    if (want_attn) launches += 1;
  } else {
    const int R = desc->rank;
    launches += 2 * R + 1;  // attention convs, top-down convs, pool
    if (desc->flags & (APA_FLAG_SOFTMAX_ATT | APA_FLAG_RELU_ATT)) launches += 1;
  }
  plan[3] = launches;
  return APA_OK;
}

}  // extern "C"

namespace {
int fwd_impl(const ApaDesc* d, const float* X, const ApaWeights* w, float* logits, float* attn_out,
             float* topdown_out, float* pose_out, void* workspace, size_t workspace_bytes,
             const ApaTuning* tuning, void* stream_, const ApaPeerGather* gw);
}

extern "C" {

int apa_attn_pool_fwd(const ApaDesc* d, const float* X, const ApaWeights* w, float* logits,
                      float* attn_out, float* topdown_out, float* pose_out, void* workspace,
                      size_t workspace_bytes, const ApaTuning* tuning, void* stream_) {
  if (logits == nullptr) return APA_ERR_NULL;
  return fwd_impl(d, X, w, logits, attn_out, topdown_out, pose_out, workspace, workspace_bytes, tuning, stream_,
                  nullptr);
}

int apa_attn_pool_fwd_gather(const ApaDesc* d, const float* X, const ApaWeights* w, const ApaPeerGather* gw,
                             void* workspace, size_t workspace_bytes, const ApaTuning* tuning, void* stream_) {
  if (gw == nullptr) return APA_ERR_NULL;
  if (gw->world < 1 || gw->world > APA_MAX_PEERS || gw->rank < 0 || gw->rank >= gw->world) return APA_ERR_SHAPE;
  if (gw->depth < 1 || gw->depth > 4 || gw->reserved0 != 0) return APA_ERR_SHAPE;
  if (gw->counters == nullptr) return APA_ERR_NULL;
  for (int q = 0; q < gw->world; ++q)
    if (gw->peer_logits[q] == nullptr || gw->peer_flags[q] == nullptr) return APA_ERR_NULL;
  int st = check_desc(d);
  if (st != APA_OK) return st;
  if (!use_fast_path(d, 0)) return APA_ERR_UNSUPPORTED;  // the fused gather lives in the single-map kernels
  return fwd_impl(d, X, w, nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes, tuning, stream_, gw);
}

int apa_peer_gather_wait(const ApaPeerGather* gw, int32_t lag, void* stream) {
  if (gw == nullptr || gw->counters == nullptr) return APA_ERR_NULL;
  if (lag < 0) return APA_ERR_SHAPE;
  if (gw->world < 1 || gw->world > APA_MAX_PEERS || gw->rank < 0 || gw->rank >= gw->world) return APA_ERR_SHAPE;
  for (int q = 0; q < gw->world; ++q)
    if (gw->peer_flags[q] == nullptr) return APA_ERR_NULL;
  int st = check_device();
  if (st != APA_OK) return st;
  APA_CUDA(peer_gather_wait_launch(gw, lag, false, static_cast<cudaStream_t>(stream)));
  return APA_OK;
}

}  // extern "C"

namespace {

int fwd_impl(const ApaDesc* d, const float* X, const ApaWeights* w, float* logits, float* attn_out,
             float* topdown_out, float* pose_out, void* workspace, size_t workspace_bytes,
             const ApaTuning* tuning, void* stream_, const ApaPeerGather* gw) {
  int st = check_desc(d);
  if (st != APA_OK) return st;
  st = check_tuning(tuning);
  if (st != APA_OK) return st;
  if (X == nullptr || w == nullptr || workspace == nullptr) return APA_ERR_NULL;
  const int B = d->batch, HW = d->hw, C = d->channels, K = d->num_classes, R = d->rank, F = d->frames;
  const bool single = (d->flags & APA_FLAG_SINGLE_LAYER_ATT) != 0;
  const bool per_class = (d->flags & APA_FLAG_PER_CLASS) != 0;
  const int M = per_class ? K : 1;
  const int mode = mode_of(d->flags);
  for (int r = 0; r < R; ++r)
    if (!w->attn_w[r] || !w->attn_b[r] || !w->topdown_w[r] || !w->topdown_b[r]) return APA_ERR_NULL;
  const bool need_pose_pre = !single || pose_out != nullptr;
  if (need_pose_pre && (!w->pose_w || !w->pose_b)) return APA_ERR_NULL;
  if (pose_out && (!w->pose_logits_w || !w->pose_logits_b || d->num_pose_keypoints <= 0)) return APA_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return APA_ERR_WORKSPACE;
  const Carve c = carve(d, attn_out != nullptr, topdown_out != nullptr, pose_out != nullptr);
  // vector (16-byte) accesses: X tiles move by bulk copies, w_0 is read as float4 on the fast path
  if (reinterpret_cast<uintptr_t>(X) & 15u) return APA_ERR_ALIGN;
  if (c.path == 0 && single && (reinterpret_cast<uintptr_t>(w->attn_w[0]) & 15u)) return APA_ERR_ALIGN;
  if (workspace_bytes < c.total) return APA_ERR_WORKSPACE;
  st = check_device();
  if (st != APA_OK) return st;

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto wsf = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  const int rows = B * HW;

  // ---- PoseLogits scope (nets_factory.py:147-160): P = relu(X Wp + bp); PoseLogits = P Wq + bq.
  // When P only feeds the 1-wide attention conv (config 003, fast path, no PoseLogits requested) it is
  // never written: the tensor-core GEMM's epilogue emits the partial dots relu(P).w_0 directly.
  float* P = nullptr;
  const bool fuse_a0 = !single && c.path == 0 && pose_out == nullptr &&
                       dense_gemm_tc_eligible(rows, d->pose_channels, C) &&
                       dense_gemm_tc_dot_parts(d->pose_channels) <= 8 && !dense_force_simt();
  if (need_pose_pre && !fuse_a0) {
    P = wsf(c.pose_pre);
    APA_CUDA(dense_gemm(X, C, 1, w->pose_w, d->pose_channels, w->pose_b, P, d->pose_channels, 1, 0, rows,
                        d->pose_channels, C, /*relu=*/1, stream));
    if (pose_out)
      APA_CUDA(dense_gemm(P, d->pose_channels, 1, w->pose_logits_w, d->num_pose_keypoints, w->pose_logits_b,
                          pose_out, d->num_pose_keypoints, 1, 0, rows, d->num_pose_keypoints,
                          d->pose_channels, 0, stream));
  }
  const float* S = single ? X : P;       // attention source (:247-250)
  const int Cs = single ? C : d->pose_channels;

  if (c.path == 0) {
    // ---- single-map fast path
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.X = X;
    if (!single) {
      float* a0 = wsf(c.a0);
      if (fuse_a0) {
        APA_CUDA(dense_gemm_tc(X, C, 1, w->pose_w, d->pose_channels, w->pose_b, nullptr, 0, 1, 0, rows,
                               d->pose_channels, C, /*relu=*/1, stream, w->attn_w[0], a0));
        p.attn_in_parts = dense_gemm_tc_dot_parts(d->pose_channels);
        p.attn_in_bias = 1;
      } else {
        APA_CUDA(dense_gemm(S, Cs, 1, w->attn_w[0], 1, w->attn_b[0], a0, 1, 1, 0, rows, 1, Cs, 0, stream));
        p.attn_in_parts = 1;
        p.attn_in_bias = 0;
      }
      p.attn_in = a0;
    }
    for (int r = 0; r < R; ++r) {
      p.attn_w[r] = w->attn_w[r];
      p.attn_b[r] = w->attn_b[r];
      p.td_w[r] = w->topdown_w[r];
      p.td_b[r] = w->topdown_b[r];
    }
    p.logits = logits;
    if (gw != nullptr) {
      p.gw_world = gw->world;
      p.gw_rank = gw->rank;
      for (int q = 0; q < gw->world; ++q) {
        p.gw_logits[q] = gw->peer_logits[q];
        p.gw_flags[q] = gw->peer_flags[q];
      }
      p.gw_done = gw->counters;
      p.gw_depth = gw->depth;
    }
    p.attn_raw = attn_out ? wsf(c.attn_raw) : nullptr;
    p.smx = wsf(c.smx);
    p.trigger_lead = tuning ? tuning->pdl_trigger_lead : 0;
    p.inputs_stable = (tuning && tuning->inputs_stable) ? 1 : 0;
    p.ts = (tuning && tuning->debug_timestamps) ? reinterpret_cast<unsigned long long*>(ws + c.ts) : nullptr;
    p.B = B; p.HW = HW; p.C = C; p.K = K; p.R = R; p.F = F;
    p.inv_scale = 1.0f / (float(HW) * float(F));
    bool launched = false;
    if (use_cluster(d, w, tuning)) {
      // this kernel has no flow control of its own for the gathered ring: a 1-warp kernel in front of it does it
      if (gw != nullptr) APA_CUDA(peer_gather_begin_launch(gw, stream));
      int grid = 0;
      cudaError_t le = cluster_launch(p, mode, tuning, stream, &grid);
      if (le == cudaSuccess) launched = true;
      else if (gw != nullptr) return int(le);   // the begin kernel has already announced the step
      else (void)cudaGetLastError();            // e.g. 16-CTA clusters not launchable here: the flow form
    }
    if (!launched) {
      p.gpart = wsf(c.gpart);
      p.gstat = wsf(c.gstat);
      p.g = wsf(c.g);
      p.gsum = wsf(c.gsum);
      p.cpart = wsf(c.cpart);
      p.fl_ctr = reinterpret_cast<unsigned long long*>(ws + kFlowCtrOff);
      int g1 = 0, g2 = 0;
      APA_CUDA(flow_launch(p, mode, tuning, stream, &g1, &g2));
    }
    if (gw != nullptr) {
      // completeness: by default this call's step; lag >= 1 pipelines the consumer behind the producers
      const int lag = tuning ? tuning->gather_wait_lag : 0;
      if (lag >= 0) APA_CUDA(peer_gather_wait_launch(gw, lag, !(tuning && tuning->no_pdl), stream));
    }
    if (attn_out)
      APA_CUDA(attn_finalize_m1(p.attn_raw, w->attn_w, w->attn_b, p.smx, attn_out, B, HW, R, mode, stream));
    return APA_OK;
  }

  // ---- dense path: materialise attention [rows, M, R] and top-down [rows, K, R], then pool
  float* A = attn_out ? attn_out : wsf(c.A);
  float* T = topdown_out ? topdown_out : wsf(c.T);
  for (int r = 0; r < R; ++r) {
    if (r == 0)
      APA_CUDA(dense_gemm(S, Cs, 1, w->attn_w[0], M, w->attn_b[0], A, M * R, R, 0, rows, M, Cs, 0, stream));
    else  // chained conv on the previous map (quirk Q1): input is column r-1 of the stacked buffer
      APA_CUDA(dense_gemm(A + (r - 1), M * R, R, w->attn_w[r], M, w->attn_b[r], A, M * R, R, r, rows, M, M, 0,
                          stream));
  }
  if (mode == MODE_SOFTMAX) APA_CUDA(dense_softmax_hw(A, B, HW, M, stream));
  else if (mode == MODE_RELU) APA_CUDA(dense_relu(A, size_t(rows) * M * R, stream));
  for (int r = 0; r < R; ++r)
    APA_CUDA(dense_gemm(X, C, 1, w->topdown_w[r], K, w->topdown_b[r], T, K * R, R, r, rows, K, C, 0, stream));
  APA_CUDA(dense_pool(A, T, logits, B, HW, K, M, R, F, stream));
  return APA_OK;
}

}  // namespace

extern "C" {

int apa_attn_pool_fwd_host(const ApaDesc* d, const float* X_host, const ApaWeights* w, float* logits_host,
                           float* x_staging, float* logits_staging, void* workspace,
                           size_t workspace_bytes, const ApaTuning* tuning, void* stream_) {
  int st = check_desc(d);
  if (st != APA_OK) return st;
  if (!X_host || !logits_host || !x_staging || !logits_staging) return APA_ERR_NULL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t xbytes = size_t(d->batch) * d->hw * d->channels * sizeof(float);
  const size_t lbytes = size_t(d->batch / d->frames) * d->num_classes * sizeof(float);
  APA_CUDA(cudaMemcpyAsync(x_staging, X_host, xbytes, cudaMemcpyHostToDevice, stream));
  st = apa_attn_pool_fwd(d, x_staging, w, logits_staging, nullptr, nullptr, nullptr, workspace,
                         workspace_bytes, tuning, stream_);
  if (st != APA_OK) return st;
  APA_CUDA(cudaMemcpyAsync(logits_host, logits_staging, lbytes, cudaMemcpyDeviceToHost, stream));
  APA_CUDA(cudaStreamSynchronize(stream));
  return APA_OK;
}

int apa_gram_fwd(int32_t batch, int32_t hw, int32_t channels, const float* X, float* G, void* stream) {
  if (!X || !G) return APA_ERR_NULL;
  if (batch <= 0 || hw <= 0 || channels <= 0) return APA_ERR_SHAPE;
  int st = check_device();
  if (st != APA_OK) return st;
  APA_CUDA(gram_launch(batch, hw, channels, X, G, static_cast<cudaStream_t>(stream)));
  return APA_OK;
}

}  // extern "C"

/*
 * apa.h — C ABI of the B200-native attentional-pooling head.
 *
 * This is the drop-in boundary for ONE path of rohitgirdhar/AttentionalPoolingAction:
 * the head that turns conv5 features X[B,H,W,C] into action logits[B,K]
 * (reference: models/slim/nets/nets_factory.py:119-125 frame fold, :143-160 pose
 * branch, :242-328 PosePrelogitsBasedAttention, :350-352 squeeze, :354-374 frame mean).
 *
 * The reference has no C ABI for this path (it is a TF-1.1 Python graph; its native
 * ops use the TF C++ OpKernel interface, src/custom_ops/pose_to_heatmap.cc:26-104).
 * The conventions below restate those OpKernel conventions as plain C:
 *   - inputs are borrowed (context->input(i), pose_to_heatmap.cc:37-44)       -> const pointers
 *   - outputs are framework-allocated (allocate_output, :55-69)               -> caller-owned buffers
 *   - errors propagate as a status (OP_REQUIRES_OK, :27-32)                    -> int return, 0 == OK
 *   - attrs are fixed at construction (:26-33)                                 -> ApaDesc
 *   - Compute() is called concurrently from TF's inter-op pool                 -> re-entrant, no global
 *     (thread-safety note pose_utils.hpp:7-14)                                    mutable state; all work
 *                                                                                 is enqueued on `stream`.
 *
 * No torch types appear here. All device pointers are plain `float*` in the
 * address space of the current CUDA device; `stream` is a cudaStream_t passed
 * as void*.  The library never allocates or frees device memory on the
 * device-pointer entry points; it never synchronises the stream there either.
 *
 * Layouts (all fp32, row-major, "NHWC with H*W flattened"):
 *   X            [batch, hw, channels]                    (nets_factory.py:137 last_conv)
 *   conv weights [Cin, Cout]  == TF HWIO [1,1,Cin,Cout] with the two leading 1s dropped
 *   logits       [batch/frames, num_classes]              (nets_factory.py:350-352, :373)
 *   attn_out     [batch, hw, M]      if rank==1 else [batch, hw, M, rank]   (:271-274, :287)
 *   topdown_out  [batch, hw, K]      if rank==1 else [batch, hw, K, rank]   (:305-309)
 *   pose_out     [batch, hw, num_pose_keypoints]          (:158-160)
 *   with M = num_classes if APA_FLAG_PER_CLASS else 1     (:258)
 */
#ifndef APA_H_
#define APA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APA_ABI_VERSION 2
#define APA_MAX_RANK 4

/* cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION_* flags (src/config.py:182-209). */
enum {
  APA_FLAG_SINGLE_LAYER_ATT = 1 << 0, /* attention reads X, not the 768-ch pose pre-logits (:247-250) */
  APA_FLAG_SOFTMAX_ATT      = 1 << 1, /* spatial softmax over H*W per map (:276-284)                  */
  APA_FLAG_RELU_ATT         = 1 << 2, /* relu on the (stacked) attention (:285-286)                   */
  APA_FLAG_PER_CLASS        = 1 << 3  /* one attention map per class, M = K (:258)                    */
};

/* Status codes. 0 == OK; negatives are library errors; positives are cudaError_t passthrough. */
enum {
  APA_OK                   = 0,
  APA_ERR_NULL             = -1, /* required pointer is NULL                                         */
  APA_ERR_SHAPE            = -2, /* bad batch/hw/channels/classes/frames                             */
  APA_ERR_RANK             = -3, /* rank outside [1, APA_MAX_RANK]                                    */
  APA_ERR_SOFTMAX_RANK     = -4, /* SOFTMAX_ATT with RANK>1: graph-build error in the reference       */
                                 /* (nets_factory.py:278 applies a 4-perm transpose to a rank-5       */
                                 /*  tensor) -- rejected, not given invented semantics                */
  APA_ERR_WORKSPACE        = -5, /* workspace too small or misaligned                                 */
  APA_ERR_UNSUPPORTED      = -6, /* valid in the reference but not built here         */
  APA_ERR_ALIGN            = -7, /* X or a weight matrix is not 16-byte aligned                       */
  APA_ERR_DEVICE           = -8  /* not an sm_100 device                                              */
};

/* Construction-time attributes of the head (what the reference fixes at graph build). */
typedef struct ApaDesc {
  int32_t abi_version;        /* = APA_ABI_VERSION                                                   */
  int32_t batch;              /* number of feature maps B*F (frames already folded, :121-125)        */
  int32_t hw;                 /* H*W spatial locations (196 synthetic, 225 in the 450-px pipeline)    */
  int32_t channels;           /* C of last_conv (2048 for resnet_v1_101/block4)                       */
  int32_t num_classes;        /* K (393 MPII, 600 HICO)                                               */
  int32_t frames;             /* F frames/crops per video; logits are averaged over F (:354-374)      */
  int32_t rank;               /* USE_POSE_PRELOGITS_BASED_ATTENTION_RANK (chained, see quirk Q1)      */
  int32_t flags;              /* APA_FLAG_*                                                           */
  int32_t pose_channels;      /* width of ExtraConv2d_1x1 (768, :151-157); used iff !SINGLE_LAYER_ATT  */
  int32_t num_pose_keypoints; /* width of Conv2d_1c_1x1 (16, :158-159); used iff pose_out requested   */
  int32_t reserved[6];        /* must be 0                                                            */
} ApaDesc;

/* Borrowed device pointers to the head's variables, addressed as the reference names them
 * (variable scopes nets_factory.py:147,243,260-269,300-303).                               */
typedef struct ApaWeights {
  /* PosePrelogitsBasedAttention/Conv2d_PrePose_Attn{,1,2,..}/{weights,biases}
   *   r == 0: [Cs, M], [M]  with Cs = channels (SINGLE_LAYER_ATT) or pose_channels
   *   r >= 1: [M, M],  [M]  (the conv is applied to the previous map: quirk Q1, :263-270)  */
  const float* attn_w[APA_MAX_RANK];
  const float* attn_b[APA_MAX_RANK];
  /* PosePrelogitsBasedAttention/Conv{,_1,_2,..}/{weights,biases}: [channels, K], [K] (:300-303) */
  const float* topdown_w[APA_MAX_RANK];
  const float* topdown_b[APA_MAX_RANK];
  /* PoseLogits/ExtraConv2d_1x1/{weights,biases}: [channels, pose_channels], [pose_channels] */
  const float* pose_w;
  const float* pose_b;
  /* PoseLogits/Conv2d_1c_1x1/{weights,biases}: [pose_channels, num_pose_keypoints], [..]    */
  const float* pose_logits_w;
  const float* pose_logits_b;
} ApaWeights;

/* Which kernel form runs the single-map (M == 1) head (both compute the same function). */
enum {
  APA_FORM_AUTO    = 0, /* the flow form wherever it applies                                                   */
  APA_FORM_FLOW    = 1, /* row-balanced STREAM kernel + three small tail kernels (apa_pool_flow.cu)            */
  APA_FORM_CLUSTER = 2  /* one kernel of 16-CTA clusters, the tail stays in shared memory (apa_pool_cluster.cu): */
                        /* batch % 4 == 0, frames == 1, channels == 2048, classes <= 768                        */
};

/* Launch/tuning knobs; pass NULL for defaults. Does not change results beyond fp32 rounding. */
typedef struct ApaTuning {
  int32_t form;         /* APA_FORM_*                                                                          */
  int32_t stream_ctas;  /* flow form: STREAM CTAs per call. 0 = 148, one per SM: the shortest single call.      */
                        /*   With several independent batches in flight (one stream each) fewer CTAs per call   */
                        /*   (37..49) pack the SMs better and raise throughput.                                 */
  int32_t stream_shape; /* flow form, A/B runs: 0 = one STREAM CTA per SM (default), 1 = two, 2 = three          */
  int32_t cluster_shape;/* cluster form: 0 = widest shape that fits, 1..3 = shapes 0..2 (apa_pool_cluster.cu)    */
  int32_t no_pdl;       /* 1 = do not launch as programmatic dependents of the previous kernel in the stream    */
  int32_t inputs_stable;/* 1 = the caller guarantees X and the weights were completely written before the       */
                        /*   PREVIOUS kernel in `stream` was launched, so the head may read them before that     */
                        /*   kernel has finished (back-to-back calls on one stream then overlap through         */
                        /*   programmatic dependent launch).  0 (default) = wait for the previous kernel first.  */
  int32_t pdl_trigger_lead; /* when a streaming CTA lets the call's next kernel be scheduled: this many tiles     */
                        /*   before its last one (0 = at the last tiles; < 0 = as soon as the CTA starts)        */
  int32_t gather_wait_lag; /* apa_attn_pool_fwd_gather: after the call the gathered buffer of step               */
                        /*   (step - lag) is complete in stream order; 0 = this step (strict), < 0 = do not      */
                        /*   wait at all (call apa_peer_gather_wait before consuming)                            */
  int32_t debug_timestamps; /* 1 = kernels write a per-CTA clock64 / globaltimer timeline into the workspace     */
  int32_t debug_cut_tail;   /* measurement only (results are then wrong): flow form launches only its first 1, 2 */
                        /*   or 3 kernels (STREAM / + REDUCE / + CONTRACT)                                      */
  int32_t tail_smem_kb; /* flow form: pad the shared memory of the REDUCE / CONTRACT / FINISH CTAs to this many KB so   */
                        /*   that fewer of them share an SM (they are placed SM by SM and an SM that holds three or    */
                        /*   more cannot take the next call's STREAM CTA); 0 = library default                        */
  int32_t reserved[5];  /* must be 0                                                                           */
} ApaTuning;

/* Library/ABI version (APA_ABI_VERSION of the build). */
int apa_abi_version(void);

/* Human-readable text for a status returned by any function below. Never NULL. */
const char* apa_strerror(int status);

/* Validate a descriptor exactly as apa_attn_pool_fwd would (no device access). */
int apa_attn_pool_check(const ApaDesc* desc);

/* Bytes of device workspace apa_attn_pool_fwd needs for `desc` (and whether the optional
 * outputs are requested). 0 on invalid desc. The workspace must be 256-byte aligned and must have
 * been zeroed once (apa_attn_pool_workspace_init) before first use (it holds the kernels' counters);
 * calls that share a workspace must be issued on one stream (they may overlap there: the per-call
 * buffers exist twice), two streams need two workspaces.                                     */
size_t apa_attn_pool_workspace(const ApaDesc* desc, int want_attn, int want_topdown, int want_pose);

/* Enqueue zeroing of the workspace on `stream`. */
int apa_attn_pool_workspace_init(void* workspace, size_t workspace_bytes, void* stream);

/* The head, forward, device buffers.  Replaces the graph built by
 * nets_factory.py:143-160 + :242-328 + :350-352 + :354-374 for a given last_conv X.
 *   logits       required, [batch/frames, K]
 *   attn_out     optional (NULL to skip): end_points['PosePrelogitsBasedAttention'] (:287)
 *   topdown_out  optional: end_points['TopDownAttention'] (:309) -- forces the dense path
 *   pose_out     optional: end_points['PoseLogits'] (:160); needs pose weights
 * Everything is enqueued on `stream`; no host synchronisation.                              */
int apa_attn_pool_fwd(const ApaDesc* desc, const float* X, const ApaWeights* weights,
                      float* logits, float* attn_out, float* topdown_out, float* pose_out,
                      void* workspace, size_t workspace_bytes, const ApaTuning* tuning,
                      void* stream);

/* Same call with HOST buffers for X and logits (weights stay device-resident, as TF variables
 * do): copies X host->device into `x_staging` (device, batch*hw*channels floats), runs the
 * head, copies logits device->host, and waits for the result on `stream`.  This is the
 * plugin-style entry a sess.run()-like caller uses (eval.py:258-264 feeds host arrays).     */
int apa_attn_pool_fwd_host(const ApaDesc* desc, const float* X_host, const ApaWeights* weights,
                           float* logits_host, float* x_staging, float* logits_staging,
                           void* workspace, size_t workspace_bytes, const ApaTuning* tuning,
                           void* stream);

/* The single collective of the multi-GPU path, fused into the kernel: instead of writing `logits`, every rank
 * stores its [batch/frames, K] block straight into a gathered buffer on every peer (NVLink peer stores from
 * the kernel's last step) at row offset rank * batch/frames, then publishes a per-rank step flag.  The gathered
 * buffers form a ring of `depth` buffers on every rank: call number `step` (1, 2, 3, ... counted on the device in
 * `counters`, so CUDA-graph replays keep counting; the same on every rank) lands in buffer step % depth.
 *   completeness  After the call, in stream order, buffer (step - lag) % depth holds the blocks of ALL ranks for
 *                 step - lag, where lag = ApaTuning.gather_wait_lag (0 by default: this step, every call then
 *                 waits for the slowest rank; lag >= 1 pipelines the consumer behind the producers; lag < 0:
 *                 nothing is waited for -- call apa_peer_gather_wait before consuming).
 *   buffer reuse  The caller must have consumed buffer step % depth (the result of call step - depth) in stream
 *                 order before it enqueues call step - depth + 1 on that rank; a rank only overwrites a peer's
 *                 buffer once that peer has started call step - depth + 1, so with depth >= 2 a fast rank
 *                 runs up to depth - 1 calls ahead of a slow one without waiting.
 * peer_* are device pointers valid on THIS GPU (symmetric memory / IPC mappings); each flag array has 2 * world
 * words and is zeroed once by the caller, like `counters`; every rank must issue the same calls.  A rank that
 * never arrives makes the waiting kernels trap after ~60 s (APA_PEER_TIMEOUT_CYCLES) instead of hanging. */
#define APA_MAX_PEERS 8
typedef struct ApaPeerGather {
  int32_t world;                                     /* 1..APA_MAX_PEERS                                   */
  int32_t rank;                                      /* this rank                                          */
  int32_t depth;                                     /* gathered buffers in the ring, 1..4                 */
  int32_t reserved0;                                 /* must be 0                                          */
  float* peer_logits[APA_MAX_PEERS];                 /* peer q's ring [depth][world * batch/frames][K]      */
  unsigned long long* peer_flags[APA_MAX_PEERS];     /* peer q's flag array [2 * world]                     */
  unsigned long long* counters;                      /* 4 local words: [0] scratch, [1] calls completed,    */
                                                     /*                [2] steps drawn, [3] reserved        */
} ApaPeerGather;

/* apa_attn_pool_fwd for the batch-sharded multi-GPU path (single-map head only): logits go to the peers'
 * gathered buffers as described above.  Replaces kernel + ncclAllGather; the reference has no collective
 * at all (in-graph towers, models/slim/deployment/model_deploy.py:145).                                */
int apa_attn_pool_fwd_gather(const ApaDesc* desc, const float* X, const ApaWeights* weights,
                             const ApaPeerGather* gather, void* workspace, size_t workspace_bytes,
                             const ApaTuning* tuning, void* stream);

/* Enqueue on `stream` a wait until the blocks of every rank for call (latest - lag) have landed in this rank's
 * ring, `latest` being the last gather call enqueued on `stream` before this one (one thread polling this rank's
 * own flag array; returns at once when they already have).                                              */
int apa_peer_gather_wait(const ApaPeerGather* gather, int32_t lag, void* stream);

/* ---- training mode: train.py:399-420 builds the same graph with is_training=True
 * and train.py:509-513 differentiates it.  Built for the single-map head that reads X directly
 * (APA_FLAG_SINGLE_LAYER_ATT set, APA_FLAG_PER_CLASS clear; any rank, relu / softmax / plain attention); other
 * descriptors return APA_ERR_UNSUPPORTED.                                                                   */
typedef struct ApaTrain {
  float dropout_keep;     /* keep probability of slim.dropout(last_conv) (nets_factory.py:296): 0.2 when            */
                          /* cfg.NET.DROPOUT < 0 (:143-146, config.py:231), else 1 - DROPOUT; 1 = no dropout.        */
                          /* Only the top-down convs read the dropped map; the attention convs read X itself.        */
  int32_t reserved0;      /* must be 0                                                                               */
  uint64_t dropout_seed;  /* the mask is a pure function of (seed, element index): 0 or 1/keep, see apa_dropout_mask; */
                          /* pass the same ApaTrain to the forward and to the backward of a step                      */
} ApaTrain;

/* Caller-owned device buffers for the gradients; every one is overwritten (not accumulated into), shaped like the
 * variable it belongs to (ApaWeights).  dX may be NULL when the backbone is frozen.                        */
typedef struct ApaGrads {
  float* dX;                        /* [batch, hw, channels]                                                 */
  float* attn_w[APA_MAX_RANK];      /* r == 0: [channels, 1]; r >= 1: [1, 1]                                  */
  float* attn_b[APA_MAX_RANK];      /* [1]                                                                   */
  float* topdown_w[APA_MAX_RANK];   /* [channels, K]                                                         */
  float* topdown_b[APA_MAX_RANK];   /* [K]                                                                   */
} ApaGrads;

/* Bytes of the buffer the training forward leaves for the backward (attention pre-activations [batch, hw] and the
 * pooled features per video), and of the scratch workspace both need (no initialisation required; 256-byte
 * aligned; calls sharing either must be on one stream).  0 for descriptors the training path does not serve. */
size_t apa_attn_pool_train_saved_bytes(const ApaDesc* desc);
size_t apa_attn_pool_train_workspace(const ApaDesc* desc);

/* Forward in training mode: logits of the graph of nets_factory.py:242-374 with the dropout of :296 applied
 * (train == NULL or dropout_keep == 1: the eval graph).  Two passes over X; writes `saved`.            */
int apa_attn_pool_fwd_train(const ApaDesc* desc, const float* X, const ApaWeights* weights, const ApaTrain* train,
                            float* logits, void* saved, size_t saved_bytes, void* workspace,
                            size_t workspace_bytes, void* stream);

/* Backward: given dlogits [batch/frames, K] = dLoss/dlogits, the gradients with respect to X and to every variable
 * on the path (what optimizer.compute_gradients returns for them, train.py:509-513).  X, weights, train and
 * `saved` must be those of the forward call.  Two passes over X and one write of dX; bit-reproducible.    */
int apa_attn_pool_bwd(const ApaDesc* desc, const float* X, const ApaWeights* weights, const ApaTrain* train,
                      const float* dlogits, const void* saved, size_t saved_bytes, const ApaGrads* grads,
                      void* workspace, size_t workspace_bytes, void* stream);

/* The dropout mask the two calls above apply, materialised: mask[i] for the first `count` elements of X in
 * row-major order (count % 4 == 0, count / 4 < 2^32).  With key = splitmix64(seed), g = i / 4,
 * a = fmix32(g ^ low32(key)), b = fmix32(g ^ high32(key)) (fmix32 = murmur3's 32-bit finaliser), the four elements
 * of group g take the 16-bit fields a & 0xFFFF, a >> 16, b & 0xFFFF, b >> 16; an element is kept iff its field is
 * below round(keep * 65536); kept elements carry 1 / keep.                                                  */
int apa_dropout_mask(const ApaTrain* train, size_t count, float* mask, void* stream);

/* ---- the losses behind the head in training (train.py:404-411 -> loss.py).
 * apa_softmax_xent_loss: loss.py:72-79, tf.losses.softmax_cross_entropy(onehot(labels), logits, weights=weight):
 *   loss[0] = weight / rows * sum_i (logsumexp(logits[i]) - logits[i, labels[i]]);  dlogits (optional) = d loss / d logits
 *   -- what apa_attn_pool_bwd takes.  row_loss [rows] is a caller scratch that receives the per-row losses.
 * apa_pose_l2_loss: loss.py:29-70 with cfg.TRAIN.LOSS_FN_POSE_SAMPLED off: pose_logits / pose_labels
 *   [batch, hw, keypoints] (end_points['PoseLogits'], nets_factory.py:158-160, and the heat-map labels of the same size),
 *   valid [batch, keypoints] bytes (labels_pose_valid);  loss[0] = weight * sum_c mean_b ( valid[b,c] * 0.5 *
 *   sum_hw (logit - label)^2 / (batch * hw) ) -- the reference divides by reduce_sum(ones([B, H, W])), so the batch
 *   size enters twice (loss.py:55-57,60-64); dpose (optional) = d loss / d pose_logits.  row_loss [batch] scratch.
 * Both add the per-row values in index order (bit-reproducible).                                              */
int apa_softmax_xent_loss(int32_t rows, int32_t num_classes, const float* logits, const int32_t* labels, float weight,
                          float* row_loss, float* loss, float* dlogits, void* stream);
int apa_pose_l2_loss(int32_t batch, int32_t hw, int32_t keypoints, const float* pose_logits, const float* pose_labels,
                     const uint8_t* valid, float weight, float* row_loss, float* loss, float* dpose, void* stream);

/* ---- evaluation tail: what eval.py does with the logits.
 * apa_eval_tail: per row, predictions = argmax(logits) (first maximum, eval.py:193) and probs = softmax(logits)
 * (eval.py:197) or sigmoid(logits) for the multi-label losses (:195); with labels, counts[0] += hits and
 * counts[1] += rows (the streaming accuracy of :200-203; zero the two words before the first batch).
 * probs / predictions / labels+counts are each optional (NULL).                                            */
enum { APA_EVAL_SOFTMAX = 0, APA_EVAL_SIGMOID = 1 };
int apa_eval_tail(int32_t rows, int32_t num_classes, const float* logits, int32_t activation, float* probs,
                  int32_t* predictions, const int32_t* labels, unsigned long long* counts, void* stream);

/* Mean average precision over all collected scores (eval.py:303-306 -> eval/utils.py:4-16 compute_map ->
 * eval/cap_eval_utils.py:55-109): scores [samples, K], labels [samples] (class index); class_ap [K] receives the
 * one-vs-rest VOC AP of each class in float64 (-1 for a class without positives, which the mean skips);
 * mean_ap[0] = mAP, mean_ap[1] = number of classes averaged.  samples <= APA_AP_MAX_SAMPLES (the scores of one
 * class are sorted in shared memory; MPII's validation split has 6988).                                     */
#define APA_AP_MAX_SAMPLES 16384
int apa_mean_ap(int32_t samples, int32_t num_classes, const float* scores, const int32_t* labels, double* class_ap,
                double* mean_ap, void* stream);

/* Full second-order pooling G[b] = X[b]^T X[b]  ([batch, C, C], symmetric).  Not present in the
 * reference code; defined from the paper's second-order pooling equation (BASELINE config 5). */
int apa_gram_fwd(int32_t batch, int32_t hw, int32_t channels, const float* X, float* G,
                 void* stream);

/* Introspection for tests/bench: how `desc` would be executed.
 * Fills plan[0] = path (0 single-map fast path, 1 dense), [1] = form (APA_FORM_FLOW / APA_FORM_CLUSTER, 0 on
 * the dense path), [2] = CTAs of the streaming kernel, [3] = kernel launches per call, [4] = shared memory
 * bytes of the streaming kernel, [5] = CTAs of the flow form's CONTRACT kernel, [6..7] = 0.     */
int apa_attn_pool_plan(const ApaDesc* desc, const ApaTuning* tuning, int want_attn,
                       int want_topdown, int want_pose, int32_t plan[8]);

#ifdef __cplusplus
}
#endif
#endif /* APA_H_ */

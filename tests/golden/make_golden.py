#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the reference's own nets_factory.py (see oracle/ref_graph.py).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Each fixture holds the inputs (X, variables, flags) and what the reference graph returned for them
(logits + end-points), all computed in float64 by the NumPy stand-in for the TF primitives.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import attn_pool_oracle as orc  # noqa: E402  (only for the synthetic input generators)
from oracle import ref_graph  # noqa: E402

S = dict(SINGLE_LAYER_ATT=True)
# name, B*F, (H,W), C, K, kp, F, flags
CASES = [
    ("cfg002_single_layer", 4, (3, 3), 32, 7, 5, 1, S),
    ("cfg003_pose_attention", 4, (3, 3), 32, 7, 5, 1, {}),
    ("cfg003_rank2_chain", 4, (3, 3), 32, 7, 5, 1, dict(RANK=2)),
    ("single_rank3", 3, (2, 3), 32, 7, 5, 1, dict(S, RANK=3)),
    ("single_softmax", 4, (3, 3), 32, 7, 5, 1, dict(S, SOFTMAX_ATT=True)),
    ("single_relu", 4, (3, 3), 32, 7, 5, 1, dict(S, RELU_ATT=True)),
    ("single_softmax_relu", 4, (3, 3), 32, 7, 5, 1, dict(S, SOFTMAX_ATT=True, RELU_ATT=True)),
    ("single_rank2_relu", 4, (3, 3), 32, 7, 5, 1, dict(S, RANK=2, RELU_ATT=True)),
    ("per_class", 4, (3, 3), 32, 7, 5, 1, dict(S, PER_CLASS=True)),
    ("per_class_softmax", 4, (3, 3), 32, 7, 5, 1, dict(S, PER_CLASS=True, SOFTMAX_ATT=True)),
    ("per_class_rank2_relu", 4, (3, 3), 32, 7, 5, 1, dict(S, PER_CLASS=True, RANK=2, RELU_ATT=True)),
    ("per_class_pose_src", 4, (3, 3), 32, 7, 5, 1, dict(PER_CLASS=True)),
    ("frames3_single", 6, (3, 3), 32, 7, 5, 3, S),
    ("frames2_rank2_pose", 4, (2, 2), 32, 7, 5, 2, dict(RANK=2)),
    # sizes the fused CUDA kernel accepts (C = 1024): kept tiny in B/HW/K so the fixture stays small
    ("fused_c1024_single", 3, (2, 3), 1024, 5, 3, 1, S),
    ("fused_c1024_softmax", 2, (2, 3), 1024, 5, 3, 1, dict(S, SOFTMAX_ATT=True)),
    ("fused_c1024_rank2_relu_frames", 4, (2, 2), 1024, 5, 3, 2, dict(S, RANK=2, RELU_ATT=True)),
    # C = 2048, the shapes every form of the single-map head accepts (the headline kernels see a fixture directly;
    # tests/test_gpu_golden.py runs these through each form)
    ("c2048_b8_single", 8, (3, 3), 2048, 5, 3, 1, S),
    ("c2048_b8_softmax", 8, (3, 3), 2048, 5, 3, 1, dict(S, SOFTMAX_ATT=True)),
    ("c2048_frames2_rank2_relu", 8, (3, 3), 2048, 5, 3, 2, dict(S, RANK=2, RELU_ATT=True)),
]
ERROR_CASES = [("softmax_rank2_build_error", dict(S, SOFTMAX_ATT=True, RANK=2))]


def main():
    index = {}
    for i, (name, BF, (H, W), C, K, kp, F, fl) in enumerate(CASES):
        flags = dict(orc.DEFAULT_FLAGS, **fl)
        variables = orc.make_params(K, C, 768, kp, flags=flags, seed=100 + i, dtype=np.float32)
        if flags["SINGLE_LAYER_ATT"] and C >= 1024:
            for k in list(variables):          # keep the big-C fixtures small: no pose branch
                if k.startswith("PoseLogits/"):
                    del variables[k]
        X = orc.make_features(BF, H * W, C, seed=100 + i, dtype=np.float32)
        if any(k.startswith("PoseLogits/") for k in variables):
            logits, eps, touched = ref_graph.run_reference_head(X, variables, flags, K, kp, F, (H, W))
        else:
            # the reference always builds the PoseLogits convs; give it throw-away variables
            v2 = dict(variables)
            v2.update({k: v for k, v in orc.make_params(K, C, 768, kp, flags=flags, seed=1).items()
                       if k.startswith("PoseLogits/")})
            logits, eps, touched = ref_graph.run_reference_head(X, v2, flags, K, kp, F, (H, W))
            eps.pop("PoseLogits", None)
            touched = [t for t in touched if not t.startswith("PoseLogits/")]
        out = {"X": X, "logits": logits}
        for k, v in variables.items():
            out["var:" + k] = v
        for k, v in eps.items():
            if k != "resnet_v1_101/block4":
                out["ep:" + k] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        index[name] = {"flags": flags, "num_classes": K, "num_pose_keypoints": kp, "frames": F, "hw": [H, W],
                       "channels": C, "variables_in_creation_order": touched}
        print("%-32s logits %s  end-points %s" % (name, logits.shape, sorted(eps)))
    for name, fl in ERROR_CASES:
        flags = dict(orc.DEFAULT_FLAGS, **fl)
        variables = orc.make_params(7, 32, 768, 5, flags=flags, seed=1)
        try:
            ref_graph.run_reference_head(orc.make_features(2, 9, 32), variables, flags, 7, 5)
            raise SystemExit("expected the reference graph build to fail for %s" % name)
        except ValueError as e:
            index[name] = {"flags": flags, "raises": "ValueError", "message": str(e)}
            print("%-32s raises ValueError: %s" % (name, e))
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

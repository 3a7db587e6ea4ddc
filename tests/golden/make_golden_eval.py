#!/usr/bin/env python
"""Generates tests/golden/eval_tail.npz by running the REFERENCE's own evaluation functions
(/root/reference/src/eval/utils.py compute_map, cap_eval_utils.py calc_pr_ovr_noref / voc_ap) on seeded inputs.
Run in the build container only (the GPU box has no /root/reference); the fixture is committed.
The modules are Python 2 (print statements), so the three functions are exec'd from the reference source at run time."""
import builtins, os, sys, types
import numpy as np

REF = "/root/reference/src"
builtins.xrange = range


def ref_functions(path, names):
    """exec the named top-level `def`s of a reference module, verbatim, read from /root/reference at run time (the
    module as a whole is Python 2 -- print statements -- and cannot be imported)."""
    src = open(path).read().split("\n")
    ns = {"np": np}
    for name in names:
        start = next(i for i, l in enumerate(src) if l.startswith("def %s(" % name))
        end = next((i for i in range(start + 1, len(src)) if src[i] and not src[i][0].isspace()), len(src))
        exec(compile("\n".join(src[start:end]), path, "exec"), ns)
    return ns


cap = ref_functions(os.path.join(REF, "eval/cap_eval_utils.py"), ["calc_pr_ovr_noref", "voc_ap"])
calc_pr_ovr_noref, voc_ap = cap["calc_pr_ovr_noref"], cap["voc_ap"]
ut = ref_functions(os.path.join(REF, "eval/utils.py"), ["compute_map"])
ut["calc_pr_ovr_noref"] = calc_pr_ovr_noref
ref_utils = types.SimpleNamespace(compute_map=ut["compute_map"])

rng = np.random.RandomState(20240924)
out = {}
# case A: 600 samples, 24 classes (two of them without positives), continuous scores (no ties)
n, K = 600, 24
labels = rng.randint(0, K - 2, size=n).astype(np.int32)
scores = rng.randn(n, K).astype(np.float32)
scores[np.arange(n), labels] += 1.0            # make the task learnable so that precision varies
mAP, APs = ref_utils.compute_map(scores, labels)
full = np.full(K, -1.0)
full[[c for c in range(K) if (labels == c).sum() > 0]] = np.asarray(APs, dtype=np.float64).reshape(-1)
out.update(a_scores=scores, a_labels=labels, a_map=np.float64(mAP), a_aps=full,
           a_acc=np.float64(np.mean(scores.argmax(axis=1) == labels)))
# case B: MPII-sized: 6988 samples, 393 classes.  Only the seed and the reference's answers are stored; the test draws
# the same scores again (numpy's legacy RandomState stream is frozen) -- see tests/test_eval_tail.py case_b_inputs
def case_b_inputs(seed=77):
    r = np.random.RandomState(seed)
    n, K = 6988, 393
    labels = r.randint(0, K, size=n).astype(np.int32)
    scores = (r.randn(n, K) * 2).astype(np.float32)
    scores[np.arange(n), labels] += 5.0
    return scores, labels


scores, labels = case_b_inputs()
mAP, APs = ref_utils.compute_map(scores, labels)
full = np.full(393, -1.0)
full[[c for c in range(393) if (labels == c).sum() > 0]] = np.asarray(APs, dtype=np.float64).reshape(-1)
out.update(b_seed=np.int64(77), b_map=np.float64(mAP), b_aps=full, b_acc=np.float64(np.mean(scores.argmax(axis=1) == labels)),
           b_scores_head=scores[:4, :8].copy())
# one raw call of the two inner functions
cid = int(out["a_labels"][0])
lab = (out["a_labels"] == cid).astype(np.float32); sc = out["a_scores"][:, cid].astype(np.float32)
P, R, score, ap = calc_pr_ovr_noref(lab, sc)
out.update(c_cid=np.int64(cid), c_P=P, c_R=R, c_ap=np.float64(np.asarray(ap).reshape(-1)[0]),
           c_voc=np.float64(np.asarray(voc_ap(R, P)).reshape(-1)[0]))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "eval_tail.npz"), **out)
print("mAP A %.6f  mAP B %.6f  acc A %.4f" % (out["a_map"], out["b_map"], out["a_acc"]))

"""The slice of the reference's global config the head reads (src/config.py:160-231).

Same key names and defaults; attribute- and item-access like EasyDict, strict about unknown
keys like `_merge_a_into_b` (src/config.py:287-307).  Only NET.* keys that reach the
attentional-pooling head are defined; everything else in the reference's cfg is out of scope.
"""
from __future__ import annotations


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def default_cfg():
    cfg = AttrDict()
    cfg.NET = AttrDict()
    N = cfg.NET
    N.USE_POSE_PRELOGITS_BASED_ATTENTION = False                  # config.py:182
    N.USE_POSE_PRELOGITS_BASED_ATTENTION_SOFTMAX_ATT = False      # :194
    N.USE_POSE_PRELOGITS_BASED_ATTENTION_RELU_ATT = False         # :196
    N.USE_POSE_PRELOGITS_BASED_ATTENTION_PER_CLASS = False        # :201
    N.USE_POSE_PRELOGITS_BASED_ATTENTION_SINGLE_LAYER_ATT = False  # :203
    N.USE_POSE_PRELOGITS_BASED_ATTENTION_WITH_POSE_FEAT = False   # :205 (unsupported: raises)
    N.USE_POSE_PRELOGITS_BASED_ATTENTION_WITH_POSE_FEAT_2LAYER = False  # :207
    N.USE_POSE_PRELOGITS_BASED_ATTENTION_RANK = 1                 # :209
    N.USE_TEMPORAL_ATT = False                                    # :212 (unsupported: raises)
    N.DROPOUT = -1.0                                              # :231
    N.TRAIN_TOP_BN = False                                        # :227
    cfg.RNG_SEED = 42                                             # :238
    return cfg


def cfg_from_dict(overrides, base=None):
    """Merge {'NET': {...}} into the defaults; unknown keys / type changes raise (config.py:294-307)."""
    cfg = base if base is not None else default_cfg()

    def merge(a, b, path):
        for k, v in a.items():
            if k not in b:
                raise KeyError("%s%s is not a valid config key" % (path, k))
            if isinstance(v, dict):
                merge(v, b[k], path + k + ".")
            else:
                if type(b[k]) is not type(v) and not (isinstance(b[k], float) and isinstance(v, int)):
                    raise ValueError("Type mismatch (%s vs. %s) for config key: %s%s" %
                                     (type(b[k]), type(v), path, k))
                b[k] = type(b[k])(v)
    merge(overrides, cfg, "")
    return cfg


def cfg_from_file(path, base=None):
    """YAML experiment files of the reference (experiments/*.yaml); keys outside NET are ignored."""
    import yaml
    with open(path, "r") as f:
        y = yaml.safe_load(f)
    known = default_cfg().NET
    net = {k: v for k, v in (y.get("NET") or {}).items() if k in known}
    return cfg_from_dict({"NET": net}, base)

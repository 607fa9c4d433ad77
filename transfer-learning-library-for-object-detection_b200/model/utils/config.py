"""The subset of the reference's global ``cfg`` the hot path reads
(lib/model/utils/config.py:11-305; keys listed in SURVEY.md section 5), with the
reference's defaults.  The layers read this object at CALL time, never cache it
(lib/ATF/faster_rcnn.py:260 mutates TEST.RPN_POST_NMS_TOP_N at run time).

When the reference's lib/ is on sys.path behind this package (see model/__init__.py), the
reference's own ``model/utils/config.py`` is loaded in this module's place -- its ``cfg`` object,
``cfg_from_file``, ``cfg_from_list`` and ``get_output_dir`` become this module's -- so there is ONE
global cfg, with every key the reference's scripts expect (cfg.POOLING_MODE, cfg.RESNET, ...).
"""


def _adopt_reference_config():
    """Execute the reference's config.py (the next `model/utils/config.py` on the package path)
    as THIS module.  Returns False if there is none or it cannot be imported (e.g. no easydict)."""
    import os
    import sys
    pkg = sys.modules.get(__name__.rsplit(".", 1)[0])
    here = os.path.dirname(os.path.abspath(__file__))
    for d in list(getattr(pkg, "__path__", []))[1:]:
        cand = os.path.join(d, "config.py")
        if os.path.abspath(d) != here and os.path.exists(cand):
            try:
                with open(cand) as f:
                    code = compile(f.read(), cand, "exec")
                g = globals()
                saved = dict(g)
                g["__file__"] = cand
                exec(code, g)
                return "cfg" in g
            except Exception:  # noqa: BLE001 -- fall back to the built-in subset
                g.clear()
                g.update(saved)
                return False
    return False



class AttrDict(dict):
    """easydict-style attribute access (easydict itself is not a dependency)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


_ADOPTED = _adopt_reference_config()
if not _ADOPTED:
    __C = AttrDict()
    cfg = __C

    __C.TRAIN = AttrDict()
    __C.TRAIN.RPN_POSITIVE_OVERLAP = 0.7      # config.py:131
    __C.TRAIN.RPN_NEGATIVE_OVERLAP = 0.3      # :134
    __C.TRAIN.RPN_CLOBBER_POSITIVES = False   # :136
    __C.TRAIN.RPN_FG_FRACTION = 0.5           # :138
    __C.TRAIN.RPN_BATCHSIZE = 256             # :140
    __C.TRAIN.RPN_NMS_THRESH = 0.7            # :142
    __C.TRAIN.RPN_PRE_NMS_TOP_N = 12000       # :144
    __C.TRAIN.RPN_POST_NMS_TOP_N = 2000       # :146
    __C.TRAIN.RPN_MIN_SIZE = 8                # :148 (read, unused: proposal_layer.py:75,113)
    __C.TRAIN.RPN_BBOX_INSIDE_WEIGHTS = (1.0, 1.0, 1.0, 1.0)  # :150
    __C.TRAIN.RPN_POSITIVE_WEIGHT = -1.0      # :154

    __C.TRAIN.BATCH_SIZE = 128                # config.py:76 (cfgs/vgg16.yml, res101.yml: 256 / 128)
    __C.TRAIN.FG_FRACTION = 0.25              # :79
    __C.TRAIN.FG_THRESH = 0.5                 # :82
    __C.TRAIN.BG_THRESH_HI = 0.5              # :86
    __C.TRAIN.BG_THRESH_LO = 0.1              # :87 (cfgs set 0.0)
    __C.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED = True  # :117
    __C.TRAIN.BBOX_NORMALIZE_MEANS = (0.0, 0.0, 0.0, 0.0)  # :118
    __C.TRAIN.BBOX_NORMALIZE_STDS = (0.1, 0.1, 0.2, 0.2)   # :119
    __C.TRAIN.BBOX_INSIDE_WEIGHTS = (1.0, 1.0, 1.0, 1.0)   # :114

    __C.TEST = AttrDict()
    __C.TEST.NMS = 0.3
    __C.TEST.RPN_NMS_THRESH = 0.7             # :193
    __C.TEST.RPN_PRE_NMS_TOP_N = 6000         # :195
    __C.TEST.RPN_POST_NMS_TOP_N = 300         # :198
    __C.TEST.RPN_MIN_SIZE = 16                # :201

    __C.RNG_SEED = 3                          # :262
    __C.USE_GPU_NMS = True                    # :281
    __C.POOLING_MODE = 'align'
    __C.POOLING_SIZE = 7                      # :289
    __C.MAX_NUM_GT_BOXES = 20                 # :292
    __C.ANCHOR_SCALES = [4, 8, 16, 32]        # :295 (scripts force this for cityscape)
    __C.ANCHOR_RATIOS = [0.5, 1, 2]           # :298
    __C.FEAT_STRIDE = [16, ]                  # :301

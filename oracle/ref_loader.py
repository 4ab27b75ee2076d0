"""Loader for the UNMODIFIED reference decoder (test infrastructure only).

Only usable where /root/reference exists (the build container).  It is used by
tests/golden/make_golden.py to produce the committed fixtures and by the
CPU-side tests that pin oracle/parq_oracle.py against the real reference.
Nothing in the product path (parq_b200/) may import this module.

The reference needs two sys.modules stubs to import under torch>=2
(SURVEY.md App. C): `torch._six` (utils/wrappers.py:31) and
`pytorch_lightning.utilities.rank_zero_only` (model/parq_decoder.py:6); the
`model` package is registered empty so model/__init__.py (which needs real
Lightning) is bypassed.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PARQ_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "parq_decoder.py"))


class _NS(dict):
    __getattr__ = dict.__getitem__


def decoder_cfg(num_queries=256, dec_layers=8, for_vis=False):
    """Attribute namespace with the MODEL.DECODER fields of config/eval.yaml:37-56."""
    return _NS(
        DIM_IN=1024, NUM_QUERIES=num_queries, NUM_SEMCLS=9, LOSS_WEIGHT=[5.0, 5.0, 5.0, 1.0],
        FOR_VIS=for_vis, TRACK_SCALE=[-1.5, 1.5, -2, 1, 0, 2], SHARE_MLP_HEADS=True,
        MEAN_SIZE_PATH=os.path.join(REFERENCE_ROOT, "data", "average_scan2cad.txt"),
        EVAL_TYPE="f1", CONF_THRESH=0.8, ENABLE_NMS=True,
        TRANSFORMER=_NS(DEC_DIM=1024, QUERIES_DIM=1024, DEC_HEADS=4, DEC_LAYERS=dec_layers,
                        DEC_FFN_DIM=768, DROPOUT_RATE=0.1,
                        SCALE=[-3, 3, -2, 0.5, 0.25, 5.25], SHARE_WEIGHTS=True))


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's PARQDecoder, project, Pose, Camera."""
    if _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    six = types.ModuleType("torch._six")
    six.string_classes = (str, bytes)
    sys.modules.setdefault("torch._six", six)
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        plu = types.ModuleType("pytorch_lightning.utilities")
        plu.rank_zero_only = lambda f: f
        pl.utilities = plu
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.utilities"] = plu
    pkg = types.ModuleType("model")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "model")]
    sys.modules["model"] = pkg

    def _load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    import utils as ref_utils  # reference utils/wrappers.py
    dec = _load("model.parq_decoder", os.path.join(REFERENCE_ROOT, "model", "parq_decoder.py"))
    tp = sys.modules["model.transformer_parq"]
    ns = types.SimpleNamespace(PARQDecoder=dec.PARQDecoder, project=tp.project, transformer_parq=tp,
                               Pose=ref_utils.Pose, Camera=ref_utils.Camera, Obb3D=ref_utils.Obb3D, decoder_module=dec)
    _loaded["ns"] = ns
    return ns

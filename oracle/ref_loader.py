"""Loader for the UNMODIFIED reference decoder (test infrastructure only).

Looks for the reference in /root/reference (sources, the build container) and, when that is absent
(the GPU box), in oracle/_ref -- the sourceless bytecode tree oracle/build_ref.py compiles from the same
files.  It is used by tests/golden/make_golden.py to produce the committed fixtures, by the tests that pin
oracle/parq_oracle.py and the CUDA path against the real reference, and by bench.py's reference arm /
cpu_baseline / gpu_torch_baseline legs.  Nothing in the product path (parq_b200/) may import this module.

The reference needs two sys.modules stubs to import under torch>=2
(SURVEY.md App. C): `torch._six` (utils/wrappers.py:31) and
`pytorch_lightning.utilities.rank_zero_only` (model/parq_decoder.py:6); the
`model` package is registered empty so model/__init__.py (which needs real
Lightning) is bypassed.
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
BYTECODE_EXT = ".pybc"          # oracle/build_ref.py: sourceless bytecode under a name the GPU-box snapshot keeps


def _find_root():
    for cand in (os.environ.get("PARQ_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and (os.path.isfile(os.path.join(cand, "model", "parq_decoder.py")) or
                     os.path.isfile(os.path.join(cand, "model", "parq_decoder" + BYTECODE_EXT))):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()
# "source": the reference tree itself; "bytecode": oracle/_ref (same modules, compiled, unmodified)
REFERENCE_KIND = "source" if os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "parq_decoder.py")) else "bytecode"


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "parq_decoder.py")) or \
        os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "parq_decoder" + BYTECODE_EXT))


class _NS(dict):
    __getattr__ = dict.__getitem__


def decoder_cfg(num_queries=256, dec_layers=8, for_vis=False, share_weights=True):
    """Attribute namespace with the MODEL.DECODER fields of config/eval.yaml:37-56."""
    return _NS(
        DIM_IN=1024, NUM_QUERIES=num_queries, NUM_SEMCLS=9, LOSS_WEIGHT=[5.0, 5.0, 5.0, 1.0],
        FOR_VIS=for_vis, TRACK_SCALE=[-1.5, 1.5, -2, 1, 0, 2], SHARE_MLP_HEADS=True,
        MEAN_SIZE_PATH=os.path.join(REFERENCE_ROOT, "data", "average_scan2cad.txt"),
        EVAL_TYPE="f1", CONF_THRESH=0.8, ENABLE_NMS=True,
        TRANSFORMER=_NS(DEC_DIM=1024, QUERIES_DIM=1024, DEC_HEADS=4, DEC_LAYERS=dec_layers,
                        DEC_FFN_DIM=768, DROPOUT_RATE=0.1,
                        SCALE=[-3, 3, -2, 0.5, 0.25, 5.25], SHARE_WEIGHTS=share_weights))


_loaded = {}


class _BytecodeFinder(importlib.abc.MetaPathFinder):
    """Resolves `utils`, `utils.*` and `model.*` to the sourceless bytecode files of oracle/_ref."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, name, path=None, target=None):
        parts = name.split(".")
        if parts[0] not in ("utils", "model"):
            return None
        base = os.path.join(self.root, *parts)
        init = os.path.join(base, "__init__" + BYTECODE_EXT)
        if os.path.isfile(init):
            return importlib.util.spec_from_file_location(name, init, loader=importlib.machinery.SourcelessFileLoader(name, init),
                                                          submodule_search_locations=[base])
        if os.path.isfile(base + BYTECODE_EXT):
            return importlib.util.spec_from_file_location(name, base + BYTECODE_EXT,
                                                          loader=importlib.machinery.SourcelessFileLoader(name, base + BYTECODE_EXT))
        return None


def load_module(name):
    """Execute reference module `name` (e.g. "model.resnet_fpn") from its .py, or from its .pyc in oracle/_ref."""
    if name in sys.modules and name in _loaded:
        return sys.modules[name]
    base = os.path.join(REFERENCE_ROOT, *name.split("."))
    if os.path.isfile(base + ".py"):
        spec = importlib.util.spec_from_file_location(name, base + ".py")
    else:
        spec = _BytecodeFinder(REFERENCE_ROOT).find_spec(name)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    _loaded[name] = m
    return m


def load_reference():
    """Returns a namespace with the reference's PARQDecoder, project, Pose, Camera."""
    if "ns" in _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_KIND == "source":
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
    elif not any(isinstance(f, _BytecodeFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _BytecodeFinder(REFERENCE_ROOT))
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

    import utils as ref_utils  # reference utils/wrappers.py
    dec = load_module("model.parq_decoder")
    tp = sys.modules["model.transformer_parq"]
    ns = types.SimpleNamespace(PARQDecoder=dec.PARQDecoder, project=tp.project, transformer_parq=tp,
                               Pose=ref_utils.Pose, Camera=ref_utils.Camera, Obb3D=ref_utils.Obb3D, decoder_module=dec)
    _loaded["ns"] = ns
    return ns


def load_add_ray_pe():
    """The reference's AddRayPE class (model/ray_positional_encoding.py:29)."""
    load_reference()
    return load_module("model.ray_positional_encoding").AddRayPE


def load_resnet_fpn():
    """The reference's model/resnet_fpn.py module (ResnetFPN :16; needs torchvision)."""
    load_reference()
    return load_module("model.resnet_fpn")


def build_decoder(sd, num_queries=256, dec_layers=8, for_vis=False, device="cpu"):
    """An instance of the unmodified PARQDecoder (model/parq_decoder.py:30) in eval mode with the given state dict
    (SHARE_WEIGHTS False when the state dict carries more than one decoder layer)."""
    ns = load_reference()
    shared = "parq_module.decoder.layers.1.norm1.weight" not in sd
    m = ns.PARQDecoder(decoder_cfg(num_queries, dec_layers, for_vis, share_weights=shared)).eval()
    m.load_state_dict(sd, strict=True)
    return m.to(device)

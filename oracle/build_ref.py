"""Byte-compile the UNMODIFIED reference into oracle/_ref/ (test / baseline infrastructure only).

The reference (/root/reference, ymingxie/PARQ) is pure Python and does not exist on the GPU box.  The
recipe for a compiled reference applies: compile it from the sources WHERE THEY LIE, outputs only into
``oracle/_ref/`` (git-ignored, but it travels to the GPU box like our own built .so).  For Python the
compiled artefact is CPython bytecode: every module below is compiled with ``py_compile`` straight from
/root/reference into a sourceless bytecode tree -- no reference source file is copied into the repository.
The files carry the extension ``.pybc`` (the snapshot that ships the repository to the GPU box drops ``*.pyc``);
oracle/ref_loader.py imports them through a small meta-path finder.  The build container and the GPU box run the
same image (same CPython), so the bytecode loads there.

    python oracle/build_ref.py          # run in the build container; __graft_entry__.build() calls it

What the tree is used for (oracle/ref_loader.py imports it when /root/reference is absent):
  * bench.py --impl reference / cpu_baseline: the reference's own PARQDecoder.forward on the host cores
    (``kind: "reference"``), and the same module on the B200 through stock PyTorch (``gpu_torch_baseline``);
  * tests: accelerate() patched onto the real PARQDecoder class on the GPU, full-size parity against the
    reference's own modules run in fp32 on the GPU (TF32 off).
Nothing under parq_b200/ imports it.
"""
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PARQ_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
EXT = ".pybc"

# modules on (or next to) the hot path; model/__init__.py and parq_lightning.py need real Lightning and are
# bypassed by the loader exactly as in the build container
MODULES = [
    "model/generic_mlp.py", "model/parq_decoder.py", "model/transformer_parq.py", "model/ray_positional_encoding.py",
    "model/resnet_fpn.py",
    "utils/__init__.py", "utils/wrappers.py", "utils/parq_utils.py", "utils/nms.py", "utils/f1_eval.py", "utils/matcher.py",
    "utils/ortho6d_transforms.py", "utils/encoding_utils.py",
]
DATA = ["data/average_scan2cad.txt"]     # BoxProcessor's mean-size table (utils/parq_utils.py:45-88): data, not code


def available():
    return os.path.isfile(os.path.join(SRC, "model", "parq_decoder.py"))


def stale():
    for rel in MODULES:
        out = os.path.join(DST, rel[:-3] + EXT)
        if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(os.path.join(SRC, rel)):
            return True
    return not all(os.path.exists(os.path.join(DST, d)) for d in DATA)


def build(force=False):
    """Returns the path of the bytecode tree, or None when the reference sources are not present here
    (the GPU box: it only uses the prebuilt tree)."""
    if not available():
        return DST if os.path.isdir(DST) else None
    if not force and not stale():
        return DST
    for rel in MODULES:
        out = os.path.join(DST, rel[:-3] + EXT)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        py_compile.compile(os.path.join(SRC, rel), cfile=out, dfile="<reference>/" + rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    for rel in DATA:
        os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), os.path.join(DST, rel))
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("CPython %d.%d bytecode of the unmodified reference (ymingxie/PARQ), compiled by oracle/build_ref.py from %s.\n"
                "Build artefact: git-ignored, never edited, no source files.\n" % (sys.version_info[0], sys.version_info[1], SRC))
    return DST


if __name__ == "__main__":
    print(build(force=True))

"""In-tree build of libparq_b200.so with nvcc for sm_100a (no JIT cache, no torch extension):
the .so is loaded with ctypes and travels with the source tree."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libparq_b200.so")
SOURCES = ["parq_api.cu"]
HEADERS = ["ptx.cuh", "gemm_tc.cuh", "gemm2_tc.cuh", "gemm_sk.cuh", "chain_tc.cuh", "attn2_tc.cuh", "attn3_tc.cuh", "attn_tc.cuh", "project_sample.cuh", "rowwise.cuh",
           "parse_pred.cuh", "raype.cuh", "fpn.cuh", os.path.join("..", "..", "include", "parq_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libparq_b200.so")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> parq_b200/libparq_b200.so (sm_100a).  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))

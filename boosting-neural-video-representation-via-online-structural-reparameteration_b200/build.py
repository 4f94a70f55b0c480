"""In-tree build of liborepnerv.so (all CUDA kernels + the C ABI) and the device self-test binary.

nvcc cross-compiles for sm_100a without a GPU; the outputs land next to this file so that they travel
with the source tree (they are git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "liborepnerv.so")
SELFTEST_PATH = os.path.join(HERE, "onr_selftest")
STAMP = os.path.join(HERE, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
# No --use_fast_math: the positional encoding needs full-range sinf/cosf (arguments reach ~2e4 rad) and
# the fold / Adam kernels must track the fp32 reference; fast intrinsics are used explicitly where safe.

LIB_SOURCES = [s for s in [
    "onr_api.cu", "conv_igemm.cu", "wgrad_igemm.cu", "fold.cu", "fold_tc.cu", "fold_branches.cu", "stem.cu", "head.cu",
    "loss_ssim.cu", "adam.cu", "evalops.cu", "layout.cu",
] if os.path.exists(os.path.join(CSRC, s))]
# test infrastructure (SIMT cross-check kernels, MMA issue microbenchmark): part of the self-test binary only
SELFTEST_SOURCES = ["selftest.cu", "conv_simt.cu", "mma_bench.cu"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _source_hash():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    with open(os.path.join(HERE, "..", "include", "orepnerv.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("build failed:\n" + " ".join(cmd) + "\n" + res.stdout)
    if verbose and res.stdout.strip():
        print(res.stdout)


def build_all(force=False, verbose=False):
    """Compile the shared library and the self-test binary if sources changed. Returns LIB_PATH."""
    digest = _source_hash()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in LIB_SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print("+", " ".join(cmd), flush=True)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("build failed:\n" + " ".join(cmd) + "\n" + out)
        if verbose and out.strip():
            print(out)
    _run([nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                      "--cudart", "static"], verbose)
    _run([nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SELFTEST_SOURCES] +
         ["-o", SELFTEST_PATH, "-L" + HERE, "-lorepnerv", "-Xlinker", "-rpath=$ORIGIN"], verbose)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built", LIB_PATH)

"""In-tree build of libb200dq.so (nvcc, sm_100a only).  Cross-compiles without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["vq.cu", "tapgemm.cu", "pconv.cu", "mmgemm.cu", "norm.cu", "norm_fused.cu", "elementwise.cu", "entropy.cu", "permuter.cu", "lpips.cu", "oplevel.cu"]
OUT = os.path.join(HERE, "libb200dq.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


KERNEL_SOURCES = {"pconv": "pconv.cu", "pconv_taps": "pconv.cu", "vq": "vq.cu", "wgrad": "mmgemm.cu", "tapgemm": "tapgemm.cu", "gn": "norm_fused.cu"}


def source_sha(kernel):
    """sha256[:16] of the sources a kernel family is built from (its .cu + the shared headers): profiles/ncu_summary.json
    records it next to the DRAM traffic it measured, bench.py reports that traffic only for the same sources."""
    import hashlib
    h = hashlib.sha256()
    for f in (KERNEL_SOURCES[kernel], "common.cuh", "tmap.h"):
        h.update(open(os.path.join(CSRC, f), "rb").read())
    return h.hexdigest()[:16]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libb200dq.so")
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

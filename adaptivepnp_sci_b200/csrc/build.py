"""Build libsci_b200.so (sm_100a only) with nvcc, in-tree.

    python adaptivepnp_sci_b200/csrc/build.py [--force] [--verbose]

Elementwise / stencil files are compiled with --fmad=false (their fp32
arithmetic mirrors the reference's separate ATen/numpy ops); the tensor-core
convolution files keep FMA contraction.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libsci_b200.so")
OBJ_DIR = os.path.join(HERE, "_obj")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
          "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", HERE]
SOURCES = [
    ("sci_ops.cu", ["--fmad=false"]),
    ("sci_tv.cu", ["--fmad=false"]),
    ("sci_conv_ref.cu", []),
    ("sci_conv_tc.cu", []),
    ("sci_train.cu", ["--fmad=false"]),
    ("sci_ddnet.cu", ["--fmad=false"]),
    ("sci_host_rng.cu", ["-Xcompiler", "-ffp-contract=off"]),
    ("sci_p2p.cu", []),
]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    srcs = [(os.path.join(HERE, s), f) for s, f in SOURCES if os.path.exists(os.path.join(HERE, s))]
    deps = [s for s, _ in srcs] + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "sci_b200.h"))
    stamp = os.path.join(OBJ_DIR, "stamp")
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = []
    procs = []
    for src, flags in srcs:
        obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        cmd = [NVCC] + COMMON + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lpthread"]
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""Builds libbg_b200.so (the C-ABI library of sm_100a kernels) in-tree with nvcc.

Usage: python byo-gan_b200/build.py [--force]
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbg_b200.so")
STAMP = os.path.join(HERE, "build", "stamp.txt")
SOURCES = ["runtime.cu", "conv_fprop.cu", "conv_splitk.cu", "conv_halo.cu", "conv_wgrad.cu", "conv_wgrad_halo.cu", "aux_kernels.cu", "head_kernels.cu", "api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--cudart", "static",
    "-Xptxas", "-v",
]


def _digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    with open(os.path.join(HERE, "..", "include", "bg_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return OUT
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [NVCC] + FLAGS + ["-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            failed = True
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see byo-gan_b200/build/ptxas.log")
    cmd = [NVCC, "-shared", "-o", OUT, "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"] + objs
    subprocess.run(cmd + ["-lpthread", "-ldl", "-lrt"], check=True)
    with open(STAMP, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

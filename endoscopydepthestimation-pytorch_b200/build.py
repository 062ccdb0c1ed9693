"""Build libendo_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree).

    python -m endo_b200.build           (or __graft_entry__.build())
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libendo_b200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"]
if os.environ.get("ENDO_BUILD_TRACE") == "1":      # clock64 trace points in the persistent kernels (tools/trace_*.py)
    COMMON.append("-DENDO_TRACE_BUILD")
# per-file extra flags: the geometric kernels follow the reference's fp32 operation order
# (bit-exact thresholded masks), so FMA contraction is switched off there; they are HBM-bound.
PER_FILE = {"geometry.cu": ["-fmad=false"], "losses.cu": ["-fmad=false"], "export.cu": ["-fmad=false"]}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(ARCH + COMMON).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libendo_b200.so cannot be built here")


def build(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    nvcc = nvcc_path()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, "-c", os.path.join(CSRC, src), "-o", obj] + ARCH + COMMON + PER_FILE.get(src, [])
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(objdir, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if failed or verbose:
        print("\n".join(log), file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed, see build/build.log")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-lcudart"]
    subprocess.check_call(cmd)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Builds libmucon_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m mucon_b200.build [--force]

The library is plain CUDA runtime + C ABI (include/mucon_b200.h); it links cudart statically so
it loads next to any torch build and shares torch's primary context.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libmucon_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "mucon_b200.h")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# per-translation-unit flags: the Viterbi TU must not contract mul+add into FMA
SOURCES = {
    "api.cu": [],
    "viterbi.cu": ["-fmad=false"],
    "masks.cu": ["-fmad=false"],
    "backbone.cu": [],
    "metrics.cu": [],
    "shead.cu": [],
}


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    nvcc = nvcc_path()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [HEADER]
    jobs = []
    objs = []
    # developer shortcut: MUCON_DEV_FAST=1 compiles only the J = 66 Viterbi instantiations (seconds
    # instead of minutes) into a separate object; never used by __graft_entry__.build()
    fast = os.environ.get("MUCON_DEV_FAST") == "1"
    for src, extra in SOURCES.items():
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if src == "backbone.cu" and os.environ.get("MUCON_LAYER_TRACE") == "1":
            obj = os.path.join(OBJ_DIR, "backbone_trace.o")  # developer build with clock stamps (scripts/trace_layer.py)
            extra = list(extra) + ["-DMUCON_LAYER_TRACE"]
            fast = True  # mark the library as a developer build
        if fast and src == "viterbi.cu":
            obj = os.path.join(OBJ_DIR, "viterbi_fast.o")
            extra = list(extra) + ["-DMUCON_ONLY_SL9"]
        objs.append(obj)
        if force or _stale(obj, [sp] + headers):
            jobs.append((src, [nvcc] + ARCH + COMMON + extra + ["-c", sp, "-o", obj]))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OBJ_DIR, src + ".log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return src

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    if jobs or force or fast or _stale(LIB, objs) or os.path.exists(LIB + '.fast'):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        marker = LIB + ".fast"  # a fast-built library is never mistaken for a complete one
        if fast:
            open(marker, "w").close()
        elif os.path.exists(marker):
            os.remove(marker)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

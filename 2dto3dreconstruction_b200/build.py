"""Build recipe for the C-ABI library: nvcc, sm_100a only, in-tree output.

    python -m 2dto3dreconstruction_b200.build      (or __graft_entry__.build())

Produces 2dto3dreconstruction_b200/lib/libr3d_b200.so from csrc/*.cu.  Objects are rebuilt only
when a source or header is newer.  No torch headers are involved: the boundary is a plain C ABI
(include/r3d_b200.h) and PyTorch is used from Python for device memory and streams only.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
# R3D_NN_STATS=1 builds a diagnostic twin (search counters compiled in) next to the product library
# R3D_BUILD_TAG=<tag> (with any of the R3D_* compile-time knobs) builds an experimental twin for tools/ sweeps
_STATS = bool(os.environ.get("R3D_NN_STATS"))
_TAG = os.environ.get("R3D_BUILD_TAG") or ("stats" if _STATS else "")
OBJDIR = os.path.join(HERE, "build_" + _TAG if _TAG else "build")
LIB = os.path.join(LIBDIR, "libr3d_b200_%s.so" % _TAG if _TAG else "libr3d_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
] + (["-DR3D_NN_MIN_CTAS=" + os.environ["R3D_NN_MIN_CTAS"]] if os.environ.get("R3D_NN_MIN_CTAS") else []) \
  + (["-DR3D_LEAN_MIN_CTAS=" + os.environ["R3D_LEAN_MIN_CTAS"]] if os.environ.get("R3D_LEAN_MIN_CTAS") else []) \
  + (["-DR3D_NN_STATS"] if os.environ.get("R3D_NN_STATS") else []) \
  + ["-D%s=%s" % (k, os.environ[k]) for k in ("R3D_LEAN_FLUSH", "R3D_SEL_SMALL_MAX", "R3D_LEAN_RCAP0", "R3D_LEAN_TMA", "R3D_LEAN_WIN", "R3D_LEAN_ROW_SEGS_MAX", "R3D_LEAN_BLOCKS_MAX", "R3D_NN_MID_PAIR", "R3D_BP_V8", "R3D_BP_ONEPASS", "R3D_LOAD256", "R3D_STORE256", "R3D_BP_MIN_CTAS", "R3D_NN_SMALL_QPT", "R3D_NN_PREFETCH", "R3D_RS3_TILE", "R3D_RS3_MINB") if os.environ.get(k)]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(INCLUDE, "r3d_b200.h"))
    return hs


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    newest_header = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header)
        if stale:
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", src, "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True)
        log = p.stdout + p.stderr
        with open(obj + ".log", "w") as f:
            f.write(log)
        return src, p.returncode, log

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, rc, log in ex.map(compile_one, jobs):
                if verbose or rc != 0:
                    print(log)
                if rc != 0:
                    raise RuntimeError("nvcc failed on %s" % src)
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            print(p.stdout + p.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""In-tree build of libfgnn_b200.so with plain nvcc (no torch C++ extension, no JIT cache).

    python factor-graph-neural-network_b200/build.py [--force] [--verbose]

The library is compiled for sm_100a only (`-gencode arch=compute_100a,code=sm_100a`), with
-lineinfo so ncu's source page maps to these files, and lands next to this file so it travels
to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfgnn_b200.so")
SOURCES = ["api.cu", "mp_simt.cu", "mp_tc.cu", "mp_src.cu", "exchange.cu", "emodel.cu", "backward.cu", "plan.cu"]
HEADERS = ["common.cuh", "tc_common.cuh", os.path.join("..", "..", "include", "fgnn_b200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".cu", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, trace=False, debug=False):
    """Compile the CUDA library if it is missing or older than its sources.  Returns its path.
    trace=True builds libfgnn_b200_trace.so with -DFGNN_TC_TRACE (per-item pipeline timestamps,
    tools/tc_trace.py); the product library never carries that code."""
    if trace:      # FGNN_TRACE_DEFS="-DX -DY" adds experiment switches, FGNN_TRACE_TAG names the library
        extra = os.environ.get("FGNN_TRACE_DEFS", "").split()
        tag = os.environ.get("FGNN_TRACE_TAG", "")
        return _compile(LIB.replace(".so", "_trace%s.so" % tag), verbose, ["-DFGNN_TC_TRACE"] + extra)
    if debug:       # watchdog time-outs print the barrier they were waiting on before trapping
        return _compile(LIB.replace(".so", "_debug.so"), verbose, ["-DFGNN_TC_DEBUG"])
    extra = os.environ.get("FGNN_BUILD_DEFS", "").split()          # experiment switches (-DFGNN_...) for A/B builds
    if extra:
        return _compile(LIB.replace(".so", os.environ.get("FGNN_BUILD_TAG", "_exp") + ".so"), verbose, extra)
    if not force and not _stale():
        return LIB
    return _compile(LIB, verbose, [])


def _compile(out, verbose, extra):
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
           "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false"]
    cmd = [c for c in cmd if c != "--use_fast_math=false"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    # the image exports CC=/opt/gcc/bin/gcc; nvcc wants the system g++
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    cmd += extra + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libfgnn_b200.so:\n" + res.stdout[-4000:])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, trace="--trace" in sys.argv,
                debug="--debug" in sys.argv))

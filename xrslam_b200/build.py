"""Builds libxrvio.so (every CUDA kernel for sm_100a + the C ABI) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libxrvio.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-diag-suppress", "550"]
if os.environ.get("XRVIO_DEBUG_ASSUME") == "1":   # diagnostic build: address-space assumptions of the solve kernels become assertions
    COMMON.append("-DXRVIO_DEBUG_ASSUME")
if os.environ.get("XRVIO_KPROF") == "1":      # diagnostic build: clock64 sections inside the solve kernel (tools/phase_probe.py)
    COMMON.append("-DXRVIO_KPROF")
# per-source extra flags: the front end mirrors OpenCV's non-FMA float arithmetic
SOURCES = {
    "frontend.cu": ["-fmad=false"],
    "backend_preint.cu": [],
    "backend_solve.cu": [],
    "backend_marg.cu": [],
    "capi_ops.cu": [],
    "capi_session.cu": [],
    "xrslam_shim.cu": [],
}


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp", ".inl"))]
    headers.append(os.path.join(HERE, "..", "include", "xrvio.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        op = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(op)
        if force or _newer(sp, op) or hdr_time > os.path.getmtime(op):
            cmd = ["nvcc"] + ARCH + COMMON + extra + ["-c", sp, "-o", op]
            if verbose:
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out.decode()))
    if failed:
        raise RuntimeError("libxrvio.so build failed")
    if force or procs or not os.path.exists(LIB):
        cmd = ["nvcc"] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    return LIB


PLAYER_DIR = os.path.join(HERE, "..", "player")
PLAYER = os.path.join(PLAYER_DIR, "xrvio_player")


def build_player(force=False):
    """Headless player (player/main.cpp) linked against libxrvio.so through the XRSLAM C ABI."""
    build()
    srcs = [os.path.join(PLAYER_DIR, f) for f in os.listdir(PLAYER_DIR) if f.endswith((".cpp", ".hpp"))]
    srcs.append(os.path.join(HERE, "..", "include", "xrslam_compat.h"))
    if force or not os.path.exists(PLAYER) or max(os.path.getmtime(s) for s in srcs + [LIB]) > os.path.getmtime(PLAYER):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", PLAYER, os.path.join(PLAYER_DIR, "main.cpp"), "-L" + LIBDIR, "-lxrvio", "-lz",
                               "-Wl,-rpath,$ORIGIN/../xrslam_b200/lib"])
    return os.path.abspath(PLAYER)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_player(force="--force" in sys.argv))

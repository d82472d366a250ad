"""In-tree build of libuwcv.so with nvcc for sm_100a (no torch extension machinery:
the library is plain CUDA behind a C ABI and is loaded with ctypes)."""
from __future__ import annotations

import os
import subprocess

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libuwcv.so")
# the same sources with -DUWCV_TUNING: sweep / elimination knobs read from the environment
# (tools/*.sh, tools/*probe*.py).  Never loaded by the product path; see _lib.use_tuning_library.
TUNING_LIB_PATH = os.path.join(LIB_DIR, "libuwcv_tuning.so")
# -DUWCV_CHECK: device-side bounds traps on every tile / plane / scratch index (memory-safety
# substitute for compute-sanitizer, which is closed on the GPU pool)
CHECK_LIB_PATH = os.path.join(LIB_DIR, "libuwcv_check.so")
SOURCES = ["uwcv_capi.cu", "paste_measure.cu", "plane_fill.cu", "planes_alloc.cu", "contour.cu", "union.cu", "nms.cu", "unpack.cu", "cleanup.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",              # bit-exact parity: every FMA in the sources is explicit
    "-Xcompiler", "-fPIC", "-shared",
]


def needs_build(path: str = LIB_PATH) -> bool:
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(PKG_DIR), "include", "uwcv.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "") -> str:
    """variant: "" (release), "tuning" (-DUWCV_TUNING) or "check" (-DUWCV_CHECK); development
    sweeps may append compile-time overrides, e.g. "tuning,CONTOUR_MINBLOCKS=12" ->
    lib/libuwcv_tuning_CONTOUR_MINBLOCKS_12.so built with -DUWCV_TUNING -DUWCV_CONTOUR_MINBLOCKS=12."""
    parts = [v for v in variant.split(",") if v]
    base = parts[0] if parts else ""
    out = {"": LIB_PATH, "tuning": TUNING_LIB_PATH, "check": CHECK_LIB_PATH}[base]
    if len(parts) > 1:
        out = out[:-3] + "_" + "_".join(p.replace("=", "_") for p in parts[1:]) + ".so"
    if not force and not needs_build(out):
        return out
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = (["-DUWCV_" + base.upper()] if base else []) + ["-DUWCV_" + p for p in parts[1:]]
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose=True, variant=sys.argv[1] if len(sys.argv) > 1 else ""))

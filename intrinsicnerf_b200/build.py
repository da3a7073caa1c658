"""Build libinrf.so (sm_100a) in-tree with nvcc.  No torch dependency in the library."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libinrf.so")
SOURCES = ["pack.cu", "stages.cu", "mlp_fp32.cu", "mlp_bwd_fp32.cu", "mlp_tc.cu", "train_tc.cu", "cluster.cu", "loss.cu", "frame.cu", "status.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "inrf.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("INRF_NVCC_EXTRA", "").split() + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see intrinsicnerf_b200/csrc/build.log")
    subprocess.check_call([_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart"])
    return LIB


def build_timeline():
    """Development variant csrc/libinrf_tl.so: mlp_tc.cu compiled with -DINRF_TC_TIMELINE (clock64 stamps of the issuer, two
    epilogue warps and a back-end warp of CTA 0, printed by the launcher), linked with the production objects.
    Selected with INRF_LIB=<path> (see _lib.py); never loaded otherwise."""
    build()
    obj = os.path.join(CSRC, "mlp_tc_tl.o")
    flags = [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    subprocess.check_call([_nvcc()] + flags + ["-DINRF_TC_TIMELINE", "-c", os.path.join(CSRC, "mlp_tc.cu"), "-o", obj])
    objs = [obj if s == "mlp_tc.cu" else os.path.join(CSRC, s.replace(".cu", ".o")) for s in SOURCES]
    out = os.path.join(CSRC, "libinrf_tl.so")
    subprocess.check_call([_nvcc(), "-shared", "-o", out] + objs + ["-lcudart"])
    return out


if __name__ == "__main__":
    if "--timeline" in sys.argv:
        print(build_timeline())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

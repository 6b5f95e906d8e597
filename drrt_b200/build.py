"""Builds drrt_b200/libdrrt_b200.so in-tree with nvcc for sm_100a.

    python -m drrt_b200.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/drrt_gpu.h); nothing here
depends on torch.  nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdrrt_b200.so")
OBJDIR = os.path.join(HERE, "build")

SOURCES = ["capi.cu", "store.cu", "scan.cu", "range.cu", "knn.cu", "edgecheck.cu", "comm.cu", "sweep.cu", "timed.cu"]
HEADERS = ["common.cuh", "store.cuh", "gridscan.cuh", "geom.cuh", os.path.join("..", "..", "include", "drrt_gpu.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the reference's x86-64 build has no FMA: never contract a*b+c (bit parity)
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xcompiler", "-ffp-contract=off",
    "--expt-extended-lambda",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (drrt_b200 has no prebuilt or CPU fallback)")


def _stamp():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(name, defines):
    """A/B builds for kernel experiments: the same sources with extra -D flags into
    drrt_b200/variants/<name>/libdrrt_b200.so (select it with DRRT_B200_LIB=<path>)."""
    out_dir = os.path.join(HERE, "variants", name)
    os.makedirs(out_dir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(out_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc] + NVCC_FLAGS + list(defines) + ["-c", os.path.join(CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
    so = os.path.join(out_dir, "libdrrt_b200.so")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs +
                          ["-Xcompiler", "-fPIC", "-cudart", "static"])
    for o in objs:
        os.remove(o)
    return so


def build(force=False, verbose=False):
    stamp_file = os.path.join(OBJDIR, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(OUT) and os.path.exists(stamp_file):
        if open(stamp_file).read().strip() == stamp:
            return OUT
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                            env=env, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("drrt_b200: nvcc build failed")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + \
          ["-Xcompiler", "-fPIC", "-cudart", "static"]
    subprocess.check_call(cmd, env=env)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return OUT


if __name__ == "__main__":
    if "--variant" in sys.argv:   # python -m drrt_b200.build --variant NAME -DX=1 -DY=2
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

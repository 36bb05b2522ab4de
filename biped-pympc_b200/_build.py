"""In-tree build of the CUDA library (sm_100a only) with plain nvcc.

``build()`` is what ``__graft_entry__.build()`` calls; it is also run lazily by
``_native.load()`` when the shared object is missing *and* nvcc is available.  The built
``.so`` files stay in ``biped-pympc_b200/lib/`` (git-ignored, shipped to the GPU box)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsrbd_mpc.so")
COMPAT_LIBS = {
    "libqp_former.so": "qp_former_evaluate.c",
    "libsparse_pdipm_multiple_iterations.so": "sparse_pdipm_evaluate.c",
}
CU_SOURCES = ("srbd_abi.cu", "srbd_qp_form.cu", "srbd_pdipm.cu", "srbd_misc.cu", "srbd_lowlevel.cu", "srbd_t1.cu")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas=-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libsrbd_mpc.so")
    os.makedirs(LIBDIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "srbd_mpc.h"))
    if force or _stale(LIB, deps):
        objs = []
        for src in CU_SOURCES:
            obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
            cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed on " + src)
            with open(obj + ".ptxas.log", "w") as f:
                f.write(r.stderr)
            objs.append(obj)
        r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", LIB],
                           capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    for name, src in COMPAT_LIBS.items():
        out = os.path.join(LIBDIR, name)
        srcp = os.path.join(CSRC, "compat", src)
        if force or _stale(out, [srcp, LIB]):
            subprocess.run(["gcc", "-shared", "-fPIC", "-O2", srcp, "-o", out, "-L" + LIBDIR, "-lsrbd_mpc",
                            "-Wl,-rpath,$ORIGIN"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

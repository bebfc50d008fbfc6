"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE ONLY).

* ``build_oracle()`` compiles ``oracle/csrc/oracle_ops.c`` (the C restatement)
  into ``oracle/_build/liboracle.so`` with gcc.
* ``build_ref()`` compiles the reference's own two CUDA sources *where they
  lie* under ``/root/reference`` (never copied) into
  ``oracle/_ref/libref_ops.so`` so that GPU tests can pin the restatement to
  the reference's own arithmetic (`nms_cuda_compute`, `ROIAlignForwardLaucher`,
  `ROIAlignBackwardLaucher`).  Skipped silently when the checkout is absent
  (e.g. on the GPU box, which receives the prebuilt .so).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("STEREO_REFERENCE", "/root/reference")
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_ops.so")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(force=False):
    src = os.path.join(HERE, "csrc", "oracle_ops.c")
    if not force and _newer(ORACLE_SO, [src]):
        return ORACLE_SO
    os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-mfma", "-mavx2",
           "-ffp-contract=off", "-fno-fast-math", "-std=gnu11", src, "-o", ORACLE_SO, "-lm"]
    subprocess.check_call(cmd)
    return ORACLE_SO


def build_ref(force=False):
    nms_cu = os.path.join(REF, "lib/model/nms/src/nms_cuda_kernel.cu")
    roi_cu = os.path.join(REF, "lib/model/roi_align/src/roi_align_kernel.cu")
    if not (os.path.exists(nms_cu) and os.path.exists(roi_cu)):
        return REF_SO if os.path.exists(REF_SO) else None
    if not force and _newer(REF_SO, [nms_cu, roi_cu]):
        return REF_SO
    os.makedirs(os.path.dirname(REF_SO), exist_ok=True)
    objs = []
    for i, cu in enumerate((nms_cu, roi_cu)):
        obj = os.path.join(os.path.dirname(REF_SO), "ref%d.o" % i)
        subprocess.check_call([
            "nvcc", "-x", "cu", "-c", cu, "-o", obj, "-O3",
            "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-I", os.path.dirname(cu), "-w"])
        objs.append(obj)
    subprocess.check_call(["nvcc", "-shared", "-o", REF_SO] + objs +
                          ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return REF_SO


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv))

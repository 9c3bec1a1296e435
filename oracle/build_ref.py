#!/usr/bin/env python
"""Build recipe for oracle/_ref: the REFERENCE's own CUDA code, compiled unmodified.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package; only tests/, __graft_entry__.smoke() and bench.py's reference /
cpu_baseline legs may use it.

What it does
------------
Compiles, from the sources WHERE THEY LIE under /root/reference (nothing is
copied into this repository), two torch extensions into oracle/_ref/:

  ref_dgr_C.so  <- submodules/diff-gaussian-rasterization/{cuda_rasterizer/
                   rasterizer_impl.cu, forward.cu, backward.cu,
                   rasterize_points.cu, ext.cpp}
                   exports rasterize_gaussians / rasterize_gaussians_backward /
                   mark_visible (ext.cpp:15-19)
  ref_knn_C.so  <- submodules/simple-knn/{simple_knn.cu, spatial.cu, ext.cpp}
                   exports distCUDA2 (ext.cpp:15-17)

Flags follow the reference's own setup.py (DGR/setup.py:29: only -I glm, i.e.
nvcc defaults -O3-ish, -fmad=true, no fast-math) plus the two pre-includes that
gcc 13 needs: `--pre-include cstdint` for
rasterizer_impl.h:40-41 and `--pre-include cfloat` for simple_knn.cu:90,154.
The module name is changed with -DTORCH_EXTENSION_NAME only (no source edit).

It also STAGES (plain file copies, into the git-ignored oracle/_ref/py/ only - never into the tracked tree)
the reference's Python modules that the contract tests drive unmodified on the GPU box, where /root/reference
does not exist: gaussian_renderer/__init__.py (render()), scene/gaussian_model.py (GaussianModel and the
deformation networks), scene/rigid_body.py, utils/*.py and the reference's own Python wrapper
diff_gaussian_rasterization/__init__.py (bound to ref_dgr_C.so by oracle/ref_py.py).

/root/reference exists only in the build container.  On the GPU box the
prebuilt .so files (git-ignored, but shipped by gpurun) are used as they are.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("GSR_REFERENCE_ROOT", "/root/reference")
DGR = os.path.join(REF, "submodules", "diff-gaussian-rasterization")
KNN = os.path.join(REF, "submodules", "simple-knn")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100,code=sm_100"]


def _torch_flags():
    import torch  # noqa: F401
    from torch.utils.cpp_extension import include_paths, library_paths

    inc = []
    for p in include_paths("cuda") + [sysconfig.get_paths()["include"]]:
        inc += ["-I", p]
    try:
        import pybind11

        inc += ["-I", pybind11.get_include()]
    except Exception:
        pass
    libdirs = library_paths("cuda")
    link = []
    for p in libdirs:
        link += ["-L", p, "-Xlinker", "-rpath=" + p]
    link += ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    return inc, link


def _compile(src, obj, name, extra):
    inc, _ = _torch_flags()
    cmd = [NVCC, "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-w", *ARCH,
           "-DTORCH_EXTENSION_NAME=" + name, "-DTORCH_API_INCLUDE_EXTENSION_H",
           *extra, *inc, "-c", src, "-o", obj]
    subprocess.run(cmd, check=True)
    return obj


def _build(name, srcs, extra, objdir):
    os.makedirs(objdir, exist_ok=True)
    _, link = _torch_flags()
    objs = [os.path.join(objdir, os.path.basename(s) + ".o") for s in srcs]
    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        list(ex.map(lambda so: _compile(so[0], so[1], name, extra), zip(srcs, objs)))
    so = os.path.join(OUT, name + ".so")
    subprocess.run([NVCC, "-shared", *objs, "-o", so, *link], check=True)
    return so


STAGED_PY = [
    ("gaussian_renderer/__init__.py", "gaussian_renderer/__init__.py"),
    ("scene/gaussian_model.py", "scene/gaussian_model.py"),            # no scene/__init__.py: namespace package, so the
    ("scene/rigid_body.py", "scene/rigid_body.py"),                    # dataset readers are never imported
    ("utils/general_utils.py", "utils/general_utils.py"),
    ("utils/graphics_utils.py", "utils/graphics_utils.py"),
    ("utils/sh_utils.py", "utils/sh_utils.py"),
    ("utils/system_utils.py", "utils/system_utils.py"),
    ("utils/loss_utils.py", "utils/loss_utils.py"),
    ("submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py",
     "ref_dgr/diff_gaussian_rasterization/__init__.py"),
]


def stage_python():
    """Copy the reference's Python modules needed by the contract tests into oracle/_ref/py (git-ignored)."""
    import shutil
    for src, dst in STAGED_PY:
        d = os.path.join(OUT, "py", dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(REF, src), d)
    print("[oracle/_ref] staged %d reference Python modules under %s" % (len(STAGED_PY), os.path.join(OUT, "py")))


def build(force=False):
    """Build both reference extensions if the reference tree is present."""
    if not os.path.isdir(DGR):
        print("[oracle/_ref] reference tree not present (%s); using prebuilt files" % REF)
        return False
    os.makedirs(OUT, exist_ok=True)
    stage_python()
    dgr_so = os.path.join(OUT, "ref_dgr_C.so")
    knn_so = os.path.join(OUT, "ref_knn_C.so")
    if force or not os.path.exists(dgr_so):
        _build("ref_dgr_C",
               [os.path.join(DGR, "cuda_rasterizer", "rasterizer_impl.cu"),
                os.path.join(DGR, "cuda_rasterizer", "forward.cu"),
                os.path.join(DGR, "cuda_rasterizer", "backward.cu"),
                os.path.join(DGR, "rasterize_points.cu"),
                os.path.join(DGR, "ext.cpp")],
               ["-I", os.path.join(DGR, "third_party", "glm"), "--pre-include", "cstdint"],
               os.path.join(OUT, "obj_dgr"))
        print("[oracle/_ref] built", dgr_so)
    if force or not os.path.exists(knn_so):
        _build("ref_knn_C",
               [os.path.join(KNN, "simple_knn.cu"),
                os.path.join(KNN, "spatial.cu"),
                os.path.join(KNN, "ext.cpp")],
               ["--pre-include", "cfloat"],
               os.path.join(OUT, "obj_knn"))
        print("[oracle/_ref] built", knn_so)
    return True


if __name__ == "__main__":
    build(force="--force" in sys.argv)

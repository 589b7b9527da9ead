"""Build recipe for `oracle/_ref/` -- the UNMODIFIED reference CUDA extensions.

TEST INFRASTRUCTURE ONLY.  Nothing under `oracle/` is imported by the product
package `cnc_b200/`; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
reference / cpu_baseline legs may load what this script produces.

What it does: compiles the reference's own sources *where they lie* under
/root/reference (nothing is copied into this repo) into torch extensions whose
shared objects land in `oracle/_ref/` (git-ignored, NOT gpurun-ignored, so the
`.so` files travel to the GPU box next to our own library):

  _gridencoder     <- gridencoder/src/{gridencoder.cu,bindings.cpp}
  pack_and_align   <- my_cuda_backen/{aligner_kernel.cu,aligner.cpp}
  nerfacc_csrc     <- nerfacc/cuda/csrc/{grid,scan,pdf,camera}.cu + nerfacc.cpp

The only deviation from the reference's setup.py files is `-std=c++17`
(torch 2.11 headers need it; the reference hard-codes c++14, see
gridencoder/setup.py:9,14 and my_cuda_backen/setup.py:21-22) and an explicit
`-gencode arch=compute_100a,code=sm_100a`.  The reference build system itself
is not run.

The reference's hot-path *Python* (ngp.py, utils_bpp_acc.py, utils.py, the vendored nerfacc package)
is built the same way: `build_py()` byte-compiles every file of PY_MODULES from where it lies into
`oracle/_ref/py/<module path>.pyc.bin` (CPython 3.12 bytecode -- a binary like the `.so` files, git-ignored,
no source text in the repo).  `oracle/ref_py.py` imports those `.pyc` files on the GPU box with the
third-party modules the reference needs (torchac, tinycudann) shimmed by the oracle, so the UNMODIFIED
reference classes (CNC_context_models, NGPRadianceField_mygrid_2D3D, OccGridEstimator, rendering, ...)
run on the UNMODIFIED reference kernels next to ours.

On the GPU box /root/reference does not exist; the tests only *load* the
prebuilt `.so` / `.pyc` files through `oracle.ref_ext.load(name)` / `oracle.ref_py.load()`.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("CNC_REFERENCE_ROOT", "/root/reference")

TARGETS = {
    "_gridencoder": {
        "sources": ["gridencoder/src/gridencoder.cu", "gridencoder/src/bindings.cpp"],
        "include": ["gridencoder/src"],
    },
    "pack_and_align": {
        "sources": ["my_cuda_backen/aligner_kernel.cu", "my_cuda_backen/aligner.cpp"],
        "include": ["my_cuda_backen/include"],
    },
    "nerfacc_csrc": {
        "sources": [
            "nerfacc/cuda/csrc/grid.cu",
            "nerfacc/cuda/csrc/scan.cu",
            "nerfacc/cuda/csrc/pdf.cu",
            "nerfacc/cuda/csrc/camera.cu",
            "nerfacc/cuda/csrc/nerfacc.cpp",
        ],
        "include": ["nerfacc/cuda/csrc/include"],
    },
}


# module name the reference's own imports use -> file under /root/reference
PY_MODULES = {
    "utils": "examples/utils.py",
    "utils_bpp_acc": "examples/utils_bpp_acc.py",
    "radiance_fields": "examples/radiance_fields/__init__.py",
    "radiance_fields.ngp": "examples/radiance_fields/ngp.py",
    "datasets": None,   # (examples/datasets/__init__.py pulls in the image loaders: an empty package stands in)
    "datasets.utils": "examples/datasets/utils.py",
    "nerfacc": "nerfacc/__init__.py",
    "nerfacc.version": "nerfacc/version.py",
    "nerfacc.data_specs": "nerfacc/data_specs.py",
    "nerfacc.grid": "nerfacc/grid.py",
    "nerfacc.pack": "nerfacc/pack.py",
    "nerfacc.pdf": "nerfacc/pdf.py",
    "nerfacc.scan": "nerfacc/scan.py",
    "nerfacc.volrend": "nerfacc/volrend.py",
    "nerfacc.cameras": "nerfacc/cameras.py",
    "nerfacc.cuda": "nerfacc/cuda/__init__.py",
    "nerfacc.cuda._backend": "nerfacc/cuda/_backend.py",
    "nerfacc.estimators": "nerfacc/estimators/__init__.py",
    "nerfacc.estimators.base": "nerfacc/estimators/base.py",
    "nerfacc.estimators.occ_grid": "nerfacc/estimators/occ_grid.py",
    "nerfacc.estimators.prop_net": "nerfacc/estimators/prop_net.py",
}
PY_OUT = os.path.join(OUT, "py")


def pyc_path(module: str) -> str:
    return os.path.join(PY_OUT, module + ".pyc.bin")


def build_py(force: bool = False) -> None:
    """byte-compile the reference's hot-path python into oracle/_ref/py (outputs only; sources stay where they are)"""
    import py_compile

    if not os.path.isdir(REF):
        raise RuntimeError(f"reference sources not present at {REF}")
    os.makedirs(PY_OUT, exist_ok=True)
    for mod, rel in PY_MODULES.items():
        if rel is None:
            continue
        out = pyc_path(mod)
        src = os.path.join(REF, rel)
        if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            # dfile: what tracebacks show (a citation of the reference file, not a path that exists on the GPU box)
            py_compile.compile(src, cfile=out, dfile="<reference>/" + rel, doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)


def built(name: str) -> bool:
    return os.path.exists(os.path.join(OUT, name, f"{name}.so"))


def build(names=None, verbose=False) -> None:
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference sources not present at {REF}")
    from torch.utils.cpp_extension import load

    build_py()
    for name in names or TARGETS:
        spec = TARGETS[name]
        if built(name):
            continue
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        load(
            name=name,
            sources=[os.path.join(REF, s) for s in spec["sources"]],
            extra_include_paths=[os.path.join(REF, i) for i in spec["include"]],
            extra_cflags=["-O3", "-std=c++17"],
            extra_cuda_cflags=[
                "-O3",
                "-std=c++17",
                "-gencode",
                "arch=compute_100a,code=sm_100a",
                "-U__CUDA_NO_HALF_OPERATORS__",
                "-U__CUDA_NO_HALF_CONVERSIONS__",
                "-U__CUDA_NO_HALF2_OPERATORS__",
            ],
            build_directory=bdir,
            verbose=verbose,
            is_python_module=False,
        )


if __name__ == "__main__":
    build(sys.argv[1:] or None, verbose=True)

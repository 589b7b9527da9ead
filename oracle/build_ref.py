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

On the GPU box /root/reference does not exist; the tests only *load* the
prebuilt `.so` files through `oracle.ref_ext.load(name)`.
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


def built(name: str) -> bool:
    return os.path.exists(os.path.join(OUT, name, f"{name}.so"))


def build(names=None, verbose=False) -> None:
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference sources not present at {REF}")
    from torch.utils.cpp_extension import load

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

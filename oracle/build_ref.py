"""TEST INFRASTRUCTURE ONLY — builds the reference's own `vren` CUDA extension for sm_100
as `oracle/_ref/vren_ref.so`, compiling the sources WHERE THEY LIE under
/root/reference/models/csrc (nothing is copied into this repo; `oracle/ref_compat.h` is
force-included to bridge the torch-2.11 / CCCL-2.8 API drift).

The result is the GPU oracle for SURVEY.md §8 rows a1-a9: parity tests (`-m gpu`) call
`vren_ref.<fn>` and our C-ABI on identical tensors, and `tests/golden/make_golden.py`
uses it to mint the committed golden vectors.

    python oracle/build_ref.py          # no-op if /root/reference is absent
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/models/csrc"
OUT_DIR = os.path.join(HERE, "_ref")


def build(verbose: bool = False) -> str | None:
    """Returns the path of vren_ref.so, or None when the reference tree is absent."""
    so = os.path.join(OUT_DIR, "vren_ref.so")
    if not os.path.isdir(REF_CSRC):
        return so if os.path.exists(so) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "6")
    from torch.utils import cpp_extension

    srcs = [os.path.join(REF_CSRC, f) for f in
            ("binding.cpp", "intersection.cu", "raymarching.cu", "volumerendering.cu",
             "ref_loss.cu", "losses.cu")]
    compat = os.path.join(HERE, "ref_compat.h")
    cpp_extension.load(
        name="vren_ref",
        sources=srcs,
        extra_include_paths=[os.path.join(REF_CSRC, "include")],
        # same optimisation level as the reference's setup.py:26-27 (-O2, default -fmad)
        extra_cflags=["-O2", "-include", compat],
        extra_cuda_cflags=["-O2", "-include", compat, "-lineinfo"],
        build_directory=OUT_DIR,
        is_python_module=False,
        verbose=verbose,
    )
    return so


def load():
    """Import the prebuilt oracle extension (needs torch + a CUDA device to *run* it)."""
    so = os.path.join(OUT_DIR, "vren_ref.so")
    if not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("vren_ref", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("vren_ref:", p)

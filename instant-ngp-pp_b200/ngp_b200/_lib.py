"""ctypes binding of libngp_b200.so — the only way the Python host side reaches the CUDA path.

The prototypes are parsed from include/ngp_b200.h at import time, so the header is the single
source of truth for the C ABI.  There is no fallback: a missing library raises ImportError, a
non-zero status raises RuntimeError (the reference's TORCH_CHECK behaviour, models/csrc/include/
utils.h:4-6), and a non-CUDA / non-contiguous tensor raises RuntimeError before any launch.
"""
import ctypes
import os
import re

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)                      # instant-ngp-pp_b200/
LIB_PATH = os.environ.get("NGP_B200_LIB") or os.path.join(_ROOT, "libngp_b200.so")      # env: A/B runs of two builds only
HEADER_PATH = os.path.join(os.path.dirname(_ROOT), "include", "ngp_b200.h")

_CT = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
       "uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32, "uint64_t": ctypes.c_uint64}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes], [argnames])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#") and "extern" not in l)
    protos = {}
    for m in re.finditer(r"((?:const\s+char\s*\*|int64_t|int)\s+)(ngp_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("const"):
            restype = ctypes.c_char_p
        else:
            restype = _CT[ret]
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                nm = re.findall(r"(\w+)\s*$", a)[0]
                ty = a[: a.rfind(nm)].strip()
                argtypes.append(ctypes.c_void_p if "*" in ty else _CT[ty.replace("const", "").strip()])
                argnames.append(nm)
        protos[name] = (restype, argtypes, argnames)
    return protos


PROTOTYPES = parse_header()

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(make -C instant-ngp-pp_b200/csrc).  ngp_b200 has no CPU or PyTorch fallback.")

lib = ctypes.CDLL(LIB_PATH)
for _name, (_res, _args, _) in PROTOTYPES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.ngp_last_error().decode()


# count + block sums + block scan (+ coarse lattice + cull pre-pass when the single-cascade culling applies, march.cu);
# every other entry point = 1 launch
_KERNELS_PER_CALL = {"raymarching_train/count": 3, "raymarching_train/count+cull": 5,
                     "occupancy_sample": 3, "occupancy_sample/warmup": 1, "occupancy_update": 4}
_launches = 0


def lib_calls() -> int:
    """Number of libngp_b200 kernels launched by this process so far (bench.py's gpu_launches)."""
    return _launches


def check(status: int, what: str = ""):
    global _launches
    _launches += _KERNELS_PER_CALL.get(what, 1)
    if status != 0:
        raise RuntimeError(f"ngp_b200 {what}: {last_error()} (status {status})")


def ptr(t):
    """device pointer of a tensor (None -> NULL); enforces the reference's CHECK_INPUT contract."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("ngp_b200: tensor must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError("ngp_b200: tensor must be contiguous")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


_device_ok = False


def require_device():
    global _device_ok
    if not _device_ok:
        if not torch.cuda.is_available():
            raise RuntimeError("ngp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        check(lib.ngp_check_device(), "device check")
        _device_ok = True

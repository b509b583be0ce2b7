"""ctypes binding of libmmdfn_b200.so (the C ABI declared in include/mmdfn_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing, or a tensor is not
a contiguous CUDA tensor of the expected dtype, the call raises."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libmmdfn_b200.so")


class MMDFNError(RuntimeError):
    pass


_T = {"i": ctypes.c_int, "l": ctypes.c_longlong, "f": ctypes.c_float, "d": ctypes.c_double, "p": ctypes.c_void_p,
      "u": ctypes.c_ulonglong}

HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mmdfn_b200.h")


def parse_header(path=HEADER_PATH):
    """name -> (restype code, argument codes), read from the C prototypes of the public header so that the
    binding can never drift from include/mmdfn_b200.h."""
    import re
    text = re.sub(r"/\*.*?\*/", " ", open(path).read(), flags=re.S)
    sigs = {}
    for res, name, args in re.findall(r"\b(int|long long)\s+(mmdfn_\w+)\s*\(([^)]*)\)\s*;", text, flags=re.S):
        codes = ""
        for a in [x.strip() for x in args.split(",")]:
            if a in ("", "void"):
                continue
            if "*" in a:
                codes += "p"
            elif "unsigned long long" in a:
                codes += "u"
            elif "long long" in a:
                codes += "l"
            elif "double" in a:
                codes += "d"
            elif "float" in a:
                codes += "f"
            elif "int" in a:
                codes += "i"
            else:
                raise MMDFNError(f"cannot bind argument {a!r} of {name}")
        sigs[name] = ("i" if res == "int" else "l", codes)
    return sigs


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise MMDFNError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in parse_header().items():
            fn = getattr(L, name)
            fn.restype = _T[res]
            fn.argtypes = [_T[c] for c in args]
        if L.mmdfn_abi_version() != 2:
            raise MMDFNError("libmmdfn_b200.so ABI version mismatch")
        _lib = L
    return _lib


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise MMDFNError("mmdfn_b200 kernels need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise MMDFNError(f"expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise MMDFNError("expected a contiguous tensor")
    return t.data_ptr()


def ptr_table(tensors):
    """Host array of device pointers (const float* const*).  Keep the return value alive during the call."""
    arr = (ctypes.c_void_p * len(tensors))(*[ptr(t) for t in tensors])
    return arr


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        kind = "CUDA error" if rc > 0 else "argument error"
        raise MMDFNError(f"{name} failed: {kind} {rc}")


def query(name, *args):
    return int(getattr(lib(), name)(*args))

"""ctypes loader for the C-ABI library (libdsr_b200.so, built in-tree by __graft_entry__.build()).

The prototypes are parsed from ``include/dsr_b200.h`` so the Python side can never drift from the
header.  There is NO fallback: if the library is missing, or an op is called on a non-CUDA tensor,
the call raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
LIB_PATH = os.environ.get("DSR_B200_LIB") or os.path.join(_HERE, "libdsr_b200.so")     # (override: A/B of two builds on one box)
HEADER_PATH = os.path.join(_ROOT, "include", "dsr_b200.h")

_lib = None
_protos = None
LAUNCHES = 0          # number of library calls that enqueue kernels (bench.py reports it)

_CTYPES = {"int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float, "double": ctypes.c_double,
           "unsigned": ctypes.c_uint, "uint64_t": ctypes.c_uint64}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [(ctype, argname), ...])} for every function the header declares."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(dsr_\w+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        alist = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    alist.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    toks = a.replace("const", "").split()
                    alist.append((_CTYPES[toks[0]], toks[-1]))
        if "*" in ret:
            rt = ctypes.c_char_p
        else:
            rt = _CTYPES[ret.replace("const", "").split()[0]]
        protos[name] = (rt, alist)
    return protos


def load():
    """Load the shared library and bind every prototype; raises if it is not built."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"dsr_b200: CUDA extension not built ({LIB_PATH} missing). Run `python __graft_entry__.py` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (rt, alist) in _protos.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = rt
        fn.argtypes = [t for t, _ in alist]
    _lib = lib
    return lib


def exported_symbols():
    load()
    return sorted(_protos)


def last_error():
    return load().dsr_last_error_string().decode()


PROFILE = None        # when a list: every call appends (name, start_event, end_event, meta) - bench.py's
PROFILE_META = None   # per-kernel timing pass (CUDA events on the launching stream)


def call(name, *args):
    """Call an int-returning entry point; raise RuntimeError(dsr_last_error_string) on failure."""
    global LAUNCHES
    lib = load()
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        PROFILE.append((name, e0, e1, PROFILE_META))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error()}")
    LAUNCHES += 1
    return rc

"""ctypes binding of libspg_b200.so (the C-ABI declared in include/spg_b200.h).

The prototypes are read from the header itself, so the binding cannot drift from the
declared ABI.  There is deliberately no fallback: if the library is missing or a call
fails, the caller gets an exception — the product path never computes on the CPU.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libspg_b200.so")
HEADER_PATH = os.path.join(_ROOT, "include", "spg_b200.h")

_CTYPES = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "spg_stream_t": ctypes.c_void_p,
}

_lib = None
_protos = None


def parse_header(path=HEADER_PATH):
    """Returns {name: (restype_str, [(ctype_str, argname), ...])} for every spg_* prototype."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int64_t|int)\s+(spg_\w+)\s*\(([^;{]*?)\)\s*;", text, re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        params = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                params.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (" ".join(ret.split()), params)
    return protos


def _to_ctype(tstr):
    if "*" in tstr:
        return ctypes.c_void_p
    t = tstr.replace("const", "").strip()
    return _CTYPES[t]


def protos():
    global _protos
    if _protos is None:
        _protos = parse_header()
    return _protos


def lib():
    """Loads the shared library (after torch, so that libcudart resolves to torch's copy)."""
    global _lib
    if _lib is not None:
        return _lib
    import torch  # noqa: F401  (loads libcudart.so.12 into the process first)

    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libspg_b200.so is not built (%s). Run `python -m superpoint_graph_b200.build`; "
            "there is no CPU fallback." % LIB_PATH)
    dll = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (ret, params) in protos().items():
        fn = getattr(dll, name)  # AttributeError if the symbol is not exported
        fn.restype = ctypes.c_char_p if "char" in ret else _CTYPES[ret]
        fn.argtypes = [_to_ctype(t) for t, _ in params]
    _lib = dll
    return _lib


def error_string(code):
    return lib().spg_error_string(int(code)).decode()


_fns = {}
_Tensor = None


def call(name, *args):
    """Calls a C-ABI entry point: tensors become device pointers (plain ints), None becomes NULL.
    Kept lean on purpose: an eager training step makes ~250 of these calls."""
    global _Tensor
    fn = _fns.get(name)
    if fn is None:
        import torch

        _Tensor = torch.Tensor
        fn = _fns[name] = getattr(lib(), name)
    rc = fn(*[a.data_ptr() if isinstance(a, _Tensor) else a for a in args])
    if rc != 0:
        raise RuntimeError("%s failed: [%d] %s" % (name, rc, error_string(rc)))


_raw_stream = None


def current_stream():
    """Raw cudaStream_t (int) of torch's current stream on the current device."""
    global _raw_stream
    if _raw_stream is None:
        import torch

        getter = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        if getter is not None:
            _raw_stream = lambda: getter(torch.cuda.current_device())
        else:
            _raw_stream = lambda: torch.cuda.current_stream().cuda_stream
    return _raw_stream()

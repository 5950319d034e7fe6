"""ctypes binding of libjorldy_b200.so, generated from include/jorldy_b200.h at import time.

The product path has NO CPU fallback: if the CUDA library is missing or a symbol the header
declares is absent, this module raises at first use.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "jorldy_b200.h")
LIB_PATH = os.path.join(_HERE, "lib", "libjorldy_b200.so")

_SCALARS = {
    "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "long long": ctypes.c_longlong,
    "uint64_t": ctypes.c_uint64, "uint32_t": ctypes.c_uint32, "float": ctypes.c_float, "double": ctypes.c_double,
}


class JbError(RuntimeError):
    pass


def parse_header(path=HEADER):
    """Returns {name: [ctypes argtypes]} for every `JB_API int name(...)` declaration."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"JB_API\s+int\s+(\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2)
        types = []
        for a in [x.strip() for x in args.replace("\n", " ").split(",") if x.strip()]:
            if a == "void":
                continue
            if "*" in a:
                types.append(ctypes.c_void_p)
                continue
            a = a.replace("const ", "").strip()
            ty = " ".join(a.split()[:-1])
            if ty not in _SCALARS:
                raise JbError(f"unknown C type '{ty}' in {name}")
            types.append(_SCALARS[ty])
        decls[name] = types
    return decls


_lib = None
_decls = None


def load():
    global _lib, _decls
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JbError(
            f"{LIB_PATH} not found: build it with `python -m jorldy_b200.build` "
            "(__graft_entry__.build()). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    _decls = parse_header()
    for name, argtypes in _decls.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise JbError(f"symbol {name} declared in {HEADER} is missing from {LIB_PATH}") from e
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    _lib = lib
    return lib


def declared_symbols():
    return sorted(parse_header().keys())


class _Caller:
    """`C.jb_xxx(args...)` -> calls the C function, raises JbError on a negative status.
    Tensors are passed as `t.data_ptr()` by the call sites; None becomes NULL."""

    def __getattr__(self, name):
        lib = load()
        fn = getattr(lib, name)
        raw_ok = name in ("jb_grad_partials_count", "jb_ppo_fused_args_size", "jb_ppo_fused_max_ctas")

        def call(*args):
            rc = fn(*args)
            if rc < 0 and not raw_ok:
                raise JbError(f"{name} failed with status {rc}")
            return rc

        setattr(self, name, call)
        return call


C = _Caller()

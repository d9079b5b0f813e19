"""ctypes binding of libssunet_b200.so (include/ssunet_b200.h).

There is no CPU fallback: importing the package without the built library, or calling an op
without a CUDA device, raises.  Build with ``python ssunet-gan_b200/csrc/build.py`` (or
``__graft_entry__.build()``).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssunet_b200.so")

SSG_F32, SSG_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
W_RSCK, W_RSKC, W_RSCK_FLIP = 0, 1, 2

_T = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_longlong, "f": ctypes.c_float, "d": ctypes.c_double}

def _parse_header():
    """Derive every entry point's ctypes signature from include/ssunet_b200.h (single source of truth)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "ssunet_b200.h")
    if not os.path.exists(hdr):
        raise SsgError("missing C-ABI header %s" % hdr)
    text = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)
    sigs = {}
    for m in re.finditer(r"\bint\s+(ssg_\w+)\s*\(([^)]*)\)\s*;", text):
        name, params = m.group(1), m.group(2).strip()
        codes = ""
        if params and params != "void":
            for prm in params.split(","):
                prm = prm.strip()
                if "*" in prm or "ssg_stream_t" in prm:
                    codes += "p"
                elif "long long" in prm:
                    codes += "l"
                elif "double" in prm:
                    codes += "d"
                elif "float" in prm:
                    codes += "f"
                elif "int" in prm:
                    codes += "i"
                else:
                    raise SsgError("cannot parse parameter %r of %s" % (prm, name))
        sigs[name] = codes
    return sigs


class SsgError(RuntimeError):
    pass


_SIGS = {k: v for k, v in _parse_header().items() if k not in ("ssg_version",)}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SsgError("libssunet_b200.so is not built (%s); run `python ssunet-gan_b200/csrc/build.py`. "
                           "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, sig in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = [_T[c] for c in sig]
            fn.restype = ctypes.c_int
        L.ssg_version.restype = ctypes.c_int
        L.ssg_last_error.restype = ctypes.c_char_p
        L.ssg_pairwise_leaves_host.argtypes = [ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong]
        L.ssg_pairwise_leaves_host.restype = ctypes.c_longlong
        L.ssg_pairwise_combine_host.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
        L.ssg_pairwise_combine_host.restype = ctypes.c_float
        L.ssg_p2p_buffer_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
        L.ssg_p2p_buffer_bytes.restype = ctypes.c_longlong
        _lib = L
    return _lib


def exported_symbols():
    """Every symbol include/ssunet_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(list(_SIGS) + ["ssg_version", "ssg_last_error", "ssg_pairwise_leaves_host", "ssg_pairwise_combine_host",
                                  "ssg_p2p_buffer_bytes"])


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return t.data_ptr()
    return t


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


launch_count = 0   # number of C-ABI compute calls issued (bench.py reports it as gpu_launches)


_prof = None


def profile_reset(names):
    """Start timing every call to the named entry points with CUDA events on the launching stream."""
    global _prof
    _prof = {"names": set(names), "events": [], "flops": 0.0, "by_name": {}}


def profile_collect():
    """Stop profiling; returns {"ms": total device time, "flops": algorithmic FLOPs, "n": launches}."""
    global _prof
    if _prof is None:
        return None
    torch.cuda.synchronize()
    p, _prof = _prof, None
    by = {}
    for name, lst in p["by_name"].items():
        by[name] = {"ms": sum(a.elapsed_time(b) for a, b, _ in lst), "flops": sum(f for _, _, f in lst), "n": len(lst)}
    return {"ms": sum(a.elapsed_time(b) for a, b in p["events"]), "flops": p["flops"], "n": len(p["events"]), "by_name": by}


def call(name, *args, flops=0.0):
    """Invoke a C-ABI entry point on the current torch CUDA stream; tensors are passed as raw pointers."""
    global launch_count
    L = lib()
    if not torch.cuda.is_available():
        raise SsgError("%s: no CUDA device; ssunet-gan_b200 has no CPU path" % name)
    conv = [_ptr(a) for a in args]
    timed = _prof is not None and name in _prof["names"] and not torch.cuda.is_current_stream_capturing()
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(L, name)(*conv, stream_ptr())
    if timed:
        e1.record()
        _prof["events"].append((e0, e1))
        _prof["flops"] += flops
        _prof["by_name"].setdefault(name, []).append((e0, e1, flops))
    launch_count += 1
    if rc != 0:
        raise SsgError("%s failed (%d): %s" % (name, rc, L.ssg_last_error().decode()))


def dtype_code(dt):
    if dt == torch.float32:
        return SSG_F32
    if dt == torch.bfloat16:
        return SSG_BF16
    raise SsgError("unsupported activation dtype %s" % dt)

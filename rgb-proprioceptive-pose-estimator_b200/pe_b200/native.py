"""ctypes binding of libpe_b200.so (the C ABI declared in include/pe_b200.h).

The prototypes are parsed from the header itself so the Python side can never drift from the C
side.  There is deliberately NO fallback: if the shared library is missing or a call fails, the
caller gets an exception -- the product path never routes around the CUDA kernels.
"""
import ctypes
import os
import re
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_ROOT = os.path.dirname(_HERE)                    # rgb-proprioceptive-pose-estimator_b200/
_REPO_ROOT = os.path.dirname(_PKG_ROOT)
HEADER = os.path.join(_REPO_ROOT, "include", "pe_b200.h")
CSRC = os.path.join(_PKG_ROOT, "csrc")
LIB_PATH = os.path.join(_HERE, "libpe_b200.so")
SOURCES = ["pe_tapgemm.cu", "pe_gemm_api.cu", "pe_elementwise.cu", "pe_head.cu", "pe_fused_head.cu", "pe_lstm_seq.cu"]

_CTYPE = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "void": None,
}


def parse_header(path=HEADER):
    """Return {name: (restype, [argtypes])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int|void)\s+(pe_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = ctypes.c_char_p if "char" in ret else _CTYPE[ret.strip()]
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = re.sub(r"\b\w+$", "", a).replace("const", "").replace("unsigned", "").strip()
                    argtypes.append(_CTYPE[base])
        protos[name] = (restype, argtypes)
    return protos


def nvcc_command(out=LIB_PATH):
    return (["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
             "-Xcompiler", "-fPIC", "-shared", "-o", out] + [os.path.join(CSRC, s) for s in SOURCES])


def build(force=False, verbose=False):
    """Compile the CUDA sources into libpe_b200.so next to this file (sm_100a only)."""
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps):
            return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


# entry points whose int return value is a VALUE, not a status code
_VALUE_RETURNING = ("pe_version", "pe_device_error", "pe_pack_block_elems", "pe_head_desc_size",
                    "pe_lstm_seq_supported", "pe_debug_fast_div")


class PeError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise PeError(
                "libpe_b200.so is missing (%s). Build it with `python __graft_entry__.py build` or "
                "pe_b200.native.build(); there is no CPU / eager fallback for this path." % LIB_PATH)
        self._dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self._dll, name)
            fn.restype = restype
            fn.argtypes = argtypes
        for name, (restype, _) in self.protos.items():
            raw = getattr(self._dll, name)
            if restype is ctypes.c_int and name not in _VALUE_RETURNING:
                setattr(self, name, self._checked(name, raw))
            else:
                setattr(self, name, raw)

    def _checked(self, name, raw):
        last_error = self._dll.pe_last_error

        def call(*args):
            _STATE["calls"] += 1
            if _STATE["timing"]:
                import torch
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = raw(*args)
                e1.record()
                _STATE["events"].append((name, e0, e1))
            else:
                rc = raw(*args)
            if rc != 0:
                raise PeError("%s failed (%d): %s" % (name, rc, (last_error() or b"").decode()))
            return 0

        call.__name__ = name
        return call

    def check_device(self):
        code = self.pe_device_error()
        if code != 0:
            raise PeError("device-side pipeline error flag = %d (tcgen05/TMA pipeline timed out)" % code)


_lib = None
_STATE = {"calls": 0, "timing": False, "events": [], "bytes": {}}


def call_count():
    """Number of C-ABI kernel calls issued so far (each launches at least one CUDA kernel)."""
    return _STATE["calls"]


def enable_timing(on):
    """Bracket every C-ABI call with CUDA events on the current stream (profiling aid for bench.py)."""
    _STATE["timing"] = bool(on)
    if on:
        _STATE["events"] = []
        _STATE["bytes"] = {}


def account(name, nbytes):
    """Algorithmic bytes (operands read once + results written once, from the call's own shapes) of one C-ABI call;
    recorded only while enable_timing is on.  bench.py divides them by the live CUDA-event time of the same calls."""
    if _STATE["timing"]:
        _STATE["bytes"][name] = _STATE["bytes"].get(name, 0) + int(nbytes)


def bytes_summary():
    return dict(_STATE["bytes"])


def timing_summary():
    """{entry point: total device milliseconds} for the calls recorded since enable_timing(True)."""
    import torch
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in _STATE["events"]:
        out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
    return out


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


class HeadDesc(ctypes.Structure):
    """Mirror of `pe_head_desc` (include/pe_b200.h); tests/test_abi_cpu.py checks the size against the C side."""
    _fields_ = [
        ("x", ctypes.c_void_p), ("ldx", ctypes.c_int), ("n_rows", ctypes.c_int),
        ("inj", ctypes.c_void_p), ("ld_inj", ctypes.c_int), ("inj_col", ctypes.c_int),
        ("k_x", ctypes.c_int), ("w_x", ctypes.c_void_p),
        ("k_h", ctypes.c_int), ("w_h", ctypes.c_void_p), ("h_prev", ctypes.c_void_p),
        ("b1", ctypes.c_void_p), ("b2", ctypes.c_void_p),
        ("j_a", ctypes.c_int), ("relu_a", ctypes.c_int),
        ("out_a", ctypes.c_void_p), ("counter", ctypes.c_void_p),
        ("lstm_hidden", ctypes.c_int), ("c_prev", ctypes.c_void_p), ("c_out", ctypes.c_void_p),
        ("h_out", ctypes.c_void_p),
        ("n_tail", ctypes.c_int), ("tail_w", ctypes.c_void_p * 3), ("tail_b", ctypes.c_void_p * 3),
        ("tail_j", ctypes.c_int * 3), ("tail_relu", ctypes.c_int * 3),
        ("out", ctypes.c_void_p), ("ld_out", ctypes.c_int),
        ("meas", ctypes.c_void_p), ("ld_meas", ctypes.c_int),
        ("diff", ctypes.c_void_p), ("ld_diff", ctypes.c_int), ("diff_col", ctypes.c_int),
    ]


def fused_head(x, ldx, n_rows, k_x, w_x, j_a, out_a, counter, out, ld_out, inj=None, ld_inj=0, inj_col=0, k_h=0,
               w_h=None, h_prev=None, b1=None, b2=None, relu_a=False, lstm_hidden=0, c_prev=None, c_out=None,
               h_out=None, tail=(), meas=None, ld_meas=0, diff=None, ld_diff=0, diff_col=0):
    """Launch pe_fused_head.  Tensor arguments are torch CUDA tensors (or None); `tail` is a sequence of
    (weight, bias, out_features, relu) for the small layers that follow phase A."""
    d = HeadDesc()
    d.x, d.ldx, d.n_rows = ptr(x), ldx, n_rows
    d.inj, d.ld_inj, d.inj_col = ptr(inj), ld_inj, inj_col
    d.k_x, d.w_x = k_x, ptr(w_x)
    d.k_h, d.w_h, d.h_prev = k_h, ptr(w_h), ptr(h_prev)
    d.b1, d.b2 = ptr(b1), ptr(b2)
    d.j_a, d.relu_a = j_a, int(bool(relu_a))
    d.out_a, d.counter = ptr(out_a), ptr(counter)
    d.lstm_hidden, d.c_prev, d.c_out, d.h_out = lstm_hidden, ptr(c_prev), ptr(c_out), ptr(h_out)
    d.n_tail = len(tail)
    for i, (w, b, j, relu) in enumerate(tail):
        d.tail_w[i], d.tail_b[i], d.tail_j[i], d.tail_relu[i] = ptr(w), ptr(b), j, int(bool(relu))
    d.out, d.ld_out = ptr(out), ld_out
    d.meas, d.ld_meas = ptr(meas), ld_meas
    d.diff, d.ld_diff, d.diff_col = ptr(diff), ld_diff, diff_col
    return lib().pe_fused_head(ctypes.byref(d), stream_ptr())


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    if not torch.cuda.is_available():
        raise PeError("no CUDA device visible: the B200 pose-estimator path has no CPU fallback")
    return torch.cuda.current_stream().cuda_stream

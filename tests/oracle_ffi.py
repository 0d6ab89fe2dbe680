"""ctypes bindings for the CPU oracle (oracle/liboracle.so) -- test infrastructure only."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")


class Buf(ctypes.Structure):
    _fields_ = [("data", ctypes.POINTER(ctypes.c_uint8)), ("size", ctypes.c_size_t), ("cap", ctypes.c_size_t)]


class Cfg(ctypes.Structure):
    _fields_ = [("variant", ctypes.c_int), ("block_size", ctypes.c_uint32), ("filt", ctypes.c_int),
                ("prec", ctypes.c_int), ("flexible", ctypes.c_int), ("match_limit", ctypes.c_uint32)]


class Token(ctypes.Structure):
    _fields_ = [("pos", ctypes.c_uint32), ("len", ctypes.c_uint32), ("idx", ctypes.c_uint32)]


class Event(ctypes.Structure):
    _fields_ = [("ctx", ctypes.c_uint32), ("sym", ctypes.c_uint32)]


class Triple(ctypes.Structure):
    _fields_ = [("cum", ctypes.c_uint32), ("frq", ctypes.c_uint32), ("sum", ctypes.c_uint32), ("stream", ctypes.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "cr_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
        L = ctypes.CDLL(so)
        L.cro_new.restype = ctypes.c_void_p
        for f in ("cro_free", "cro_reset_models", "cro_set_flexible", "cro_trace_enable", "cro_trace_clear", "cro_filter_inplace",
                  "cro_dictionary_load", "cro_dictionary_encode", "cro_dictionary_decode", "cro_lzencode", "cro_lzdecode",
                  "cro_trace_tokens", "cro_trace_events", "cro_trace_triples"):
            getattr(L, f).argtypes = None
        L.cro_trace_tokens.restype = ctypes.c_size_t
        L.cro_trace_events.restype = ctypes.c_size_t
        L.cro_trace_triples.restype = ctypes.c_size_t
        L.cro_rolz_parse.restype = ctypes.c_size_t
        L.cro_lzp_parse.restype = ctypes.c_size_t
        _lib = L
    return _lib


def _take(b):
    r = ctypes.string_at(b.data, b.size) if b.size else b""
    lib().cro_buf_free(ctypes.byref(b))
    return r


def compress(data, variant=0, block_size=16 << 20, filt=0, prec=0, flexible=0, match_limit=0):
    b = Buf()
    cfg = Cfg(variant, block_size, filt, prec, flexible, match_limit)
    lib().cro_compress(ctypes.byref(cfg), data, ctypes.c_size_t(len(data)), ctypes.byref(b))
    return _take(b)


def decompress(data, variant=0):
    b = Buf()
    rc = lib().cro_decompress(variant, data, ctypes.c_size_t(len(data)), ctypes.byref(b))
    if rc != 0:
        raise ValueError("bad magic")
    return _take(b)


def dicpick(data):
    b = Buf()
    lib().cro_dicpick(data, ctypes.c_size_t(len(data)), ctypes.byref(b))
    return _take(b)


def lcp_encode(text):
    b = Buf()
    L = lib()
    tmp = ctypes.create_string_buffer(text, len(text))
    # hand the oracle a malloc'ed copy it may free
    libc = ctypes.CDLL(None)
    libc.malloc.restype = ctypes.c_void_p
    p = libc.malloc(len(text) + 1)
    ctypes.memmove(p, tmp, len(text))
    b.data = ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8))
    b.size = len(text)
    b.cap = len(text) + 1
    L.cro_dic_lcp_encode(ctypes.byref(b))
    return _take(b)


class Oracle:
    """One re-entrant oracle context = one reference process (models, dictionary, filter state)."""

    def __init__(self, variant=0):
        self.L = lib()
        self.c = ctypes.c_void_p(self.L.cro_new(variant))
        self.variant = variant

    def close(self):
        if self.c:
            self.L.cro_free(self.c)
            self.c = None

    def __del__(self):
        self.close()

    def reset_models(self):
        self.L.cro_reset_models(self.c)

    def trace(self, on=True):
        self.L.cro_trace_enable(self.c, int(on))
        self.L.cro_trace_clear(self.c)

    def dictionary_load(self, text, init_trie=1):
        return self.L.cro_dictionary_load(self.c, ctypes.c_char_p(text), init_trie)

    def dictionary_encode(self, data):
        b = Buf()
        self.L.cro_dictionary_encode(self.c, data, ctypes.c_uint32(len(data)), ctypes.byref(b))
        return _take(b)

    def dictionary_decode(self, data):
        b = Buf()
        self.L.cro_dictionary_decode(self.c, data, ctypes.c_uint32(len(data)), ctypes.byref(b))
        return _take(b)

    def lzdecode(self, data):
        b = Buf()
        self.L.cro_lzdecode(self.c, data, ctypes.c_uint32(len(data)), ctypes.byref(b))
        return _take(b)

    def lzencode(self, data):
        b = Buf()
        self.L.cro_lzencode(self.c, data, ctypes.c_uint32(len(data)), ctypes.byref(b))
        return _take(b)

    def filter_inplace(self, data, en_de=0):
        buf = ctypes.create_string_buffer(bytes(data) + bytes(64), len(data) + 64)
        filt = self.L.cro_filter_inplace(self.c, buf, ctypes.c_uint32(len(data)), en_de)
        return filt, buf.raw[:len(data)]

    def tokens(self):
        p = ctypes.POINTER(Token)()
        n = self.L.cro_trace_tokens(self.c, ctypes.byref(p))
        return [(p[i].pos, p[i].len, p[i].idx) for i in range(n)]

    def events(self):
        import numpy as np
        p = ctypes.POINTER(Event)()
        n = self.L.cro_trace_events(self.c, ctypes.byref(p))
        if n == 0:
            return np.zeros((0, 2), dtype=np.uint32)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint32)), shape=(n, 2)).copy()

    def triples(self):
        import numpy as np
        p = ctypes.POINTER(Triple)()
        n = self.L.cro_trace_triples(self.c, ctypes.byref(p))
        if n == 0:
            return np.zeros((0, 4), dtype=np.uint32)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint32)), shape=(n, 4)).copy()


def rolz_parse(data, flexible=0):
    p = ctypes.POINTER(Token)()
    n = lib().cro_rolz_parse(data, ctypes.c_uint32(len(data)), flexible, ctypes.byref(p))
    out = [(p[i].pos, p[i].len, p[i].idx) for i in range(n)]
    ctypes.CDLL(None).free(p)
    return out


def ref_binary(name):
    """Path of an unmodified reference binary under oracle/_ref, or None where the reference was never built (a CPU-only checkout
    without /root/reference).  On a GPU box oracle/_ref must have travelled with the repo: its absence there would silently
    drop every comparison with the reference CLI, so it is an error, not a downgrade."""
    path = os.path.join(REF_DIR, name)
    if os.path.exists(path):
        return path
    if os.path.exists("/dev/nvidiactl") or os.path.isdir("/root/reference/src"):
        raise RuntimeError("oracle/_ref/%s is missing: build it here with `make -C oracle ref` (it is git-ignored but travels to the "
                           "GPU box with gpurun); the reference-CLI parity checks must not be skipped" % name)
    return None


def ref_compress(data, binary, flags=(), tmpdir="/tmp"):
    """Runs the unmodified reference CLI (oracle/_ref) on `data`; returns the container bytes."""
    exe = ref_binary(binary)
    if exe is None:
        return None
    src = os.path.join(tmpdir, "crref_%d.in" % os.getpid())
    dst = src + ".out"
    with open(src, "wb") as f:
        f.write(data)
    subprocess.run([exe, "-q", *flags, "e", src, dst], check=True)
    with open(dst, "rb") as f:
        out = f.read()
    os.remove(src)
    os.remove(dst)
    return out


def ref_decompress(container, binary, tmpdir="/tmp"):
    exe = ref_binary(binary)
    if exe is None:
        return None
    src = os.path.join(tmpdir, "crref_%d.cin" % os.getpid())
    dst = src + ".out"
    with open(src, "wb") as f:
        f.write(container)
    subprocess.run([exe, "-q", "d", src, dst], check=True)
    with open(dst, "rb") as f:
        out = f.read()
    os.remove(src)
    os.remove(dst)
    return out

"""Python mirror of the C ABI in include/crgpu.h (ctypes over libcrgpu.so).

This is plumbing for tests and bench.py, not a second implementation: every call goes straight into the CUDA
library.  If libcrgpu.so is missing or no CUDA device is present the calls raise -- there is no CPU path.
"""
from __future__ import annotations

import ctypes
import os
import time

# see crgpu_api.cu: cr_more_work_queues (only counts if no CUDA context exists yet in this process)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcrgpu.so")

ROLZ, LZP, LZ77 = 0, 1, 2
OWN_STREAM = 1   # pass as `stream`: private stream per handle (several handles then overlap on one GPU)

_lib = None
last_batch_call_s = 0.0


class CrgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("crgpu error %d: %s" % (code, msg))
        self.code = code


def load(path: str | None = None):
    """Loads libcrgpu.so (once).  Raises FileNotFoundError if the CUDA extension has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(p + " is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(comprox_b200 has no CPU fallback)")
    L = ctypes.CDLL(p)
    L.crgpu_strerror.restype = ctypes.c_char_p
    L.crgpu_debug_fetch.restype = ctypes.c_int64
    if path is None:
        _lib = L
    return L


def _check(L, rc):
    if rc != 0:
        raise CrgpuError(rc, L.crgpu_strerror(rc).decode())


def decompress_batch(handles, containers, out_caps):
    """crgpu_decompress_batch: containers[i] is decoded with handles[i]; all in flight at once (one warp each)."""
    k = len(handles)
    assert k == len(containers) == len(out_caps) and k > 0
    L = handles[0].L
    outs = [ctypes.create_string_buffer(max(int(c), 1)) for c in out_caps]
    hs = (ctypes.c_void_p * k)(*[h.h for h in handles])
    ins = (ctypes.c_char_p * k)(*containers)
    in_lens = (ctypes.c_uint64 * k)(*[len(c) for c in containers])
    outp = (ctypes.c_void_p * k)(*[ctypes.addressof(o) for o in outs])
    caps = (ctypes.c_uint64 * k)(*[int(c) for c in out_caps])
    lens = (ctypes.c_uint64 * k)()
    global last_batch_call_s
    t0 = time.perf_counter()
    rc = L.crgpu_decompress_batch(hs, ctypes.c_uint32(k), ins, in_lens, outp, caps, lens)
    last_batch_call_s = time.perf_counter() - t0           # the C ABI call alone
    _check(L, rc)
    return [outs[i].raw[:lens[i]] for i in range(k)]


def compress_batch(handles, datas, block_size: int = 16 << 20, filt: bool = False, prec: bool = False, flexible: bool = False):
    """crgpu_compress_batch: independent containers side by side, dealt to whichever handle is idle (one host thread per handle inside
    the C call).  Returns the containers in input order."""
    k, m = len(handles), len(datas)
    assert k > 0
    L = handles[0].L
    L.crgpu_compress_bound.restype = ctypes.c_uint64
    cfg = Config(block_size, int(filt), int(prec), int(flexible), 0)
    caps = [int(L.crgpu_compress_bound(ctypes.c_uint64(len(d)), ctypes.c_uint32(block_size))) for d in datas]
    outs = [ctypes.create_string_buffer(c) for c in caps]
    hs = (ctypes.c_void_p * k)(*[h.h for h in handles])
    ins = (ctypes.c_char_p * max(m, 1))(*[bytes(d) for d in datas])
    in_lens = (ctypes.c_uint64 * max(m, 1))(*[len(d) for d in datas])
    outp = (ctypes.c_void_p * max(m, 1))(*[ctypes.addressof(o) for o in outs])
    ocaps = (ctypes.c_uint64 * max(m, 1))(*caps)
    lens = (ctypes.c_uint64 * max(m, 1))()
    _check(L, L.crgpu_compress_batch(hs, ctypes.c_uint32(k), ctypes.byref(cfg), ctypes.c_uint32(m), ins, in_lens, outp, ocaps, lens))
    return [outs[i].raw[:lens[i]] for i in range(m)]


def lcp_encode(text: bytes, lib=None) -> bytes:
    """dic_lcp_encode(): front coding of the NUL-terminated dictionary text (host side)."""
    L = lib or load()
    out = ctypes.create_string_buffer(len(text) + 64)
    n = ctypes.c_uint64()
    _check(L, L.crgpu_dic_lcp_encode(bytes(text), ctypes.c_uint64(len(text)), out, ctypes.c_uint64(len(text) + 64), ctypes.byref(n)))
    return out.raw[:n.value]


def lcp_decode(data: bytes, lib=None, cap: int = 25000 * 24 + 64) -> bytes:
    """dic_lcp_decode(): the dictionary text with its final NUL."""
    L = lib or load()
    out = ctypes.create_string_buffer(cap)
    n = ctypes.c_uint64()
    _check(L, L.crgpu_dic_lcp_decode(bytes(data), ctypes.c_uint64(len(data)), out, ctypes.c_uint64(cap), ctypes.byref(n)))
    return out.raw[:n.value]


class Config(ctypes.Structure):
    _fields_ = [("block_size", ctypes.c_uint32), ("filt", ctypes.c_int32), ("prec", ctypes.c_int32), ("flexible", ctypes.c_int32),
                ("window_bytes", ctypes.c_uint64)]


class Handle:
    """One crgpu_handle: the state a reference *process* holds in globals (models, dictionary, filter state)."""

    def __init__(self, variant: int = ROLZ, device: int = 0, stream: int | None = None, lib=None):
        self.L = lib or load()
        self.h = ctypes.c_void_p()
        _check(self.L, self.L.crgpu_create(ctypes.byref(self.h), variant, device, ctypes.c_void_p(stream or 0)))
        self.variant = variant

    def close(self):
        if getattr(self, "h", None):
            self.L.crgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def reset_models(self):
        _check(self.L, self.L.crgpu_reset_models(self.h))

    def lzencode(self, blocks, chain_ends: bool = True):
        """lzencode() of consecutive blocks of one model chain; returns the list of payloads."""
        blocks = [bytes(b) for b in blocks]
        n = len(blocks)
        data = b"".join(blocks)
        sizes = (ctypes.c_uint32 * max(n, 1))(*[len(b) for b in blocks])
        osz = (ctypes.c_uint32 * max(n, 1))()
        cap = len(data) + 64 * n + 1024
        out = ctypes.create_string_buffer(cap)
        _check(self.L, self.L.crgpu_lzencode(self.h, data, sizes, n, int(chain_ends), out, ctypes.c_uint64(cap), osz))
        res, p = [], 0
        for i in range(n):
            res.append(out.raw[p:p + osz[i]])
            p += osz[i]
        return res

    def compress(self, data: bytes, block_size: int = 16 << 20, filt: bool = False, prec: bool = False, flexible: bool = False,
                 window_bytes: int = 0) -> bytes:
        """The container the reference CLI writes for `comprolz/comprop [-bN] [-F] [-p] e`."""
        cfg = Config(block_size, int(filt), int(prec), int(flexible), window_bytes)
        self.L.crgpu_compress_bound.restype = ctypes.c_uint64
        cap = self.L.crgpu_compress_bound(ctypes.c_uint64(len(data)), ctypes.c_uint32(block_size))
        out = ctypes.create_string_buffer(cap)
        n = ctypes.c_uint64()
        t0 = time.perf_counter()
        rc = self.L.crgpu_compress(self.h, ctypes.byref(cfg), data, ctypes.c_uint64(len(data)), out, ctypes.c_uint64(cap), ctypes.byref(n))
        self.last_call_s = time.perf_counter() - t0          # the C ABI call alone (pageable host buffers), without the Python copies
        _check(self.L, rc)
        return ctypes.string_at(out, n.value)

    def dicpick(self, data: bytes) -> bytes:
        """dicpick(): the dictionary text (with the final NUL) built from the whole input."""
        cap = 25000 * 24 + 64
        out = ctypes.create_string_buffer(cap)
        n = ctypes.c_uint64()
        _check(self.L, self.L.crgpu_dicpick(self.h, data, ctypes.c_uint64(len(data)), out, ctypes.c_uint64(cap), ctypes.byref(n)))
        return out.raw[:n.value]

    # ---- stage-level entry points (one block per call; argument meaning of the reference's cr-* functions)
    def filter_inplace(self, data: bytes, en_de: int = 0):
        """filter_inplace(): returns (fired, transformed bytes); filter state carries over between calls."""
        buf = ctypes.create_string_buffer(bytes(data), max(len(data), 1))
        rc = self.L.crgpu_filter_inplace(self.h, buf, ctypes.c_uint32(len(data)), int(en_de))
        if rc < 0:
            _check(self.L, rc)
        return rc, buf.raw[:len(data)]

    def dictionary_load(self, text: bytes, init_trie: int = 1) -> int:
        rc = self.L.crgpu_dictionary_load(self.h, ctypes.c_char_p(text), int(init_trie))
        if rc < 0:
            _check(self.L, rc)
        return rc

    def _block_call(self, fn, data, cap):
        out = ctypes.create_string_buffer(max(cap, 1))
        n = ctypes.c_uint32()
        _check(self.L, fn(self.h, bytes(data), ctypes.c_uint32(len(data)), out, ctypes.c_uint64(cap), ctypes.byref(n)))
        return out.raw[:n.value]

    def dictionary_encode(self, data: bytes) -> bytes:
        return self._block_call(self.L.crgpu_dictionary_encode, data, len(data) + 1)

    def dictionary_decode(self, data: bytes, out_cap: int) -> bytes:
        return self._block_call(self.L.crgpu_dictionary_decode, data, out_cap)

    def lzdecode(self, payload: bytes) -> bytes:
        self.L.crgpu_lzdecode_size.restype = ctypes.c_int64
        size = self.L.crgpu_lzdecode_size(self.variant, bytes(payload), ctypes.c_uint32(len(payload)))
        if size < 0:
            _check(self.L, int(size))
        return self._block_call(self.L.crgpu_lzdecode, payload, int(size))

    def decompress(self, container: bytes, out_cap: int, grow: bool = False) -> bytes:
        """The bytes `comprolz/comprop d` would write for this container.  grow=True retries once with the size the library reports when
        out_cap turns out too small (the container does not store its decoded size)."""
        out = ctypes.create_string_buffer(max(out_cap, 1))
        n = ctypes.c_uint64()
        t0 = time.perf_counter()
        rc = self.L.crgpu_decompress(self.h, container, ctypes.c_uint64(len(container)), out, ctypes.c_uint64(out_cap), ctypes.byref(n))
        if rc == -3 and n.value > out_cap and grow:
            # out_cap was too small: *out_n holds the size needed and the second call resumes behind the entropy stage
            out_cap = n.value
            out = ctypes.create_string_buffer(out_cap)
            rc = self.L.crgpu_decompress(self.h, container, ctypes.c_uint64(len(container)), out, ctypes.c_uint64(out_cap), ctypes.byref(n))
        self.last_call_s = time.perf_counter() - t0
        _check(self.L, rc)
        return ctypes.string_at(out, n.value)

    def get_stat(self, name: str) -> int:
        """crgpu_get_stat: "cut_blocks" / "last_cut_blocks" (mid-chain blocks stored raw, see include/crgpu.h)."""
        self.L.crgpu_get_stat.restype = ctypes.c_int64
        v = self.L.crgpu_get_stat(self.h, name.encode())
        if v < 0:
            _check(self.L, int(v))
        return int(v)

    def stage_input(self, data: bytes):
        _check(self.L, self.L.crgpu_stage_input(self.h, data, ctypes.c_uint64(len(data))))

    def set_option(self, name, value):
        _check(self.L, self.L.crgpu_set_option(self.h, name.encode(), ctypes.c_int64(int(value))))

    def profile(self, enable=True):
        _check(self.L, self.L.crgpu_profile(self.h, int(enable)))

    def profile_report(self):
        buf = ctypes.create_string_buffer(8192)
        self.L.crgpu_profile_report(self.h, buf, ctypes.c_uint64(8192))
        return {k: float(v) for k, v in (ln.split() for ln in buf.value.decode().splitlines())}

    def debug_sort(self, keys, vals, begin_bit=0, end_bit=None):
        """Stable device radix sort of (keys, vals) on key bits [begin_bit, end_bit) (cr_sort.cuh); returns (keys, vals)."""
        import numpy as np
        keys = np.ascontiguousarray(keys); vals = np.ascontiguousarray(vals, dtype=np.uint32)
        kb = keys.dtype.itemsize
        assert kb in (4, 8) and len(keys) == len(vals)
        ko = np.empty_like(keys); vo = np.empty_like(vals)
        _check(self.L, self.L.crgpu_debug_sort(self.h, keys.ctypes.data_as(ctypes.c_void_p), vals.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(len(keys)),
                                               kb, int(begin_bit), int(kb * 8 if end_bit is None else end_bit),
                                               ko.ctypes.data_as(ctypes.c_void_p), vo.ctypes.data_as(ctypes.c_void_p)))
        return ko, vo

    def debug_scan(self, vals):
        """Exclusive prefix sum (mod 2^32) of uint32 values through the device scan (cr_sort.cuh)."""
        import numpy as np
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        vo = np.empty_like(vals)
        _check(self.L, self.L.crgpu_debug_sort(self.h, None, vals.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(len(vals)), 0, 0, 0,
                                               None, vo.ctypes.data_as(ctypes.c_void_p)))
        return vo

    def debug_fetch(self, what: str, dtype="uint8"):
        import numpy as np
        n = self.L.crgpu_debug_fetch(self.h, what.encode(), None, ctypes.c_uint64(0))
        if n < 0:
            _check(self.L, int(n))
        buf = np.empty(int(n), dtype=np.uint8)
        if n:
            self.L.crgpu_debug_fetch(self.h, what.encode(), buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(int(n)))
        return buf.view(dtype)

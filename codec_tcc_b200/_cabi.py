"""ctypes binding of libpeeb200.so (the C ABI declared in include/peeb200.h).

There is no CPU fallback: if the shared library is missing, cannot be built, or
no sm_100 device is visible, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import weakref

import numpy as np

from . import build as _build

PEEB_OK = 0
PEEB_E_CUDA = -1
PEEB_E_CAPACITY = -2
PEEB_E_INVALID = -3
PEEB_E_UNSUPPORTED = -4
MOMENTS = 12
INFO = 8
INFO_KEYS = ("T", "n_bits", "capacity", "cap0", "cap1", "n_flagged", "sse", "status")

_vp, _i32, _i64, _sz, _f64 = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double

# name -> (restype, argtypes); every symbol include/peeb200.h declares
SIGNATURES = {
    "peeb_abi_version": (_i32, []),
    "peeb_last_error": (C.c_char_p, []),
    "peeb_device_count": (_i32, [_vp]),
    "peeb_ws_create": (_i32, [_i32, _vp]),
    "peeb_ws_destroy": (_i32, [_vp]),
    "peeb_ws_sync": (_i32, [_vp]),
    "peeb_ws_stream": (_vp, [_vp]),
    "peeb_ws_set_option": (_i32, [_vp, _i32, _i32]),
    "peeb_host_alloc": (_i32, [_sz, _vp]),
    "peeb_host_free": (_i32, [_vp]),
    "peeb_dev_alloc": (_i32, [_vp, _sz, _vp]),
    "peeb_dev_free": (_i32, [_vp, _vp]),
    "peeb_memcpy_h2d": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "peeb_memcpy_d2h": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "peeb_prof_enable": (_i32, [_vp, _i32]),
    "peeb_prof_get": (_i32, [_vp, _i32, _vp, _vp]),
    "peeb_prof_name": (C.c_char_p, [_i32]),
    "peeb_pee_step_counters": (_i32, [_vp, _i32, _vp]),
    "peeb_moments_batch": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i64, _i64, _vp, _vp]),
    "peeb_moments_h": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "peeb_sse_batch": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i64, _i64, _vp, _vp]),
    "peeb_sse_h": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "peeb_hist_planes": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "peeb_hist_planes_h": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "peeb_planes_unpack": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "peeb_planes_unpack_h": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "peeb_planes_pack": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "peeb_planes_pack_h": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "peeb_tile_moments": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "peeb_tile_moments_h": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "peeb_lsb_embed": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "peeb_lsb_embed_h": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "peeb_compact_bits": (_i32, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp]),
    "peeb_compact_bits_h": (_i32, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _vp]),
    "peeb_lsb_recover": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "peeb_lsb_recover_h": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "peeb_lsb_extract": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp]),
    "peeb_lsb_extract_h": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "peeb_payload_bytes": (_sz, [_i64]),
    "peeb_pee_embed_batch": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64,
                                    _vp, _i64, _vp, _vp]),
    "peeb_pee_extract_batch": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64,
                                      _vp, _i64, _vp, _vp]),
    "peeb_pee_hist_batch": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "peeb_pee_embed_h": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "peeb_pee_extract_h": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "peeb_pee_hist_h": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "peeb_pee_med_embed_batch": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64,
                                        _vp, _i64, _vp, _vp]),
    "peeb_pee_med_extract_batch": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64,
                                          _vp, _i64, _vp, _vp]),
    "peeb_pee_med_embed_h": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "peeb_pee_med_extract_h": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "peeb_moments_f64": (_i32, [_vp, _vp, _vp, _i64, _i32, _f64, _f64, _f64, _f64, _f64, _f64, _vp, _vp]),
    "peeb_moments_f64_h": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "peeb_bitmap_blob_bound": (_sz, [_i64]),
    "peeb_bitmap_encode": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _vp, _vp]),
    "peeb_bitmap_decode": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp]),
    "peeb_bitmap_encode_h": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _vp]),
    "peeb_bitmap_decode_h": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32]),
    "peeb_debug_bounds": (_i32, [_vp, _i32]),
}

_lib = None
_lib_lock = threading.Lock()


class PeebError(RuntimeError):
    """A libpeeb200 call failed (CUDA error, unsupported shape, ...)."""


def library_path() -> str:
    """The in-tree build; PEEB_LIBRARY names another build of the same sources (the bounds-checked one,
    tests/test_gpu_bounds_build.py) -- it must exist, nothing is substituted for it."""
    alt = os.environ.get("PEEB_LIBRARY")
    if alt:
        if not os.path.exists(alt):
            raise PeebError(f"PEEB_LIBRARY={alt} does not exist")
        return alt
    return _build.LIBPATH


def debug_bounds(reset: bool = False) -> dict:
    """Counters of a bounds-checked build (see peeb_debug_bounds in include/peeb200.h)."""
    out = (C.c_ulonglong * 6)()
    check(lib().peeb_debug_bounds(out, 1 if reset else 0), "peeb_debug_bounds")
    return {"violations": int(out[0]), "first_site": int(out[1]), "first_offset": int(C.c_longlong(out[2]).value),
            "first_limit": int(out[3]), "items_checked": int(out[4]), "checking": bool(out[5])}


def lib():
    """The loaded shared library.  Builds it when it is missing and nvcc is
    available; otherwise raises -- there is no other implementation to fall
    back to."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            try:
                _build.build()
            except Exception as exc:  # noqa: BLE001
                raise PeebError(
                    f"libpeeb200.so is missing at {path} and could not be built ({exc}); "
                    "run `python -m codec_tcc_b200.build`") from exc
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if L.peeb_abi_version() != 1:
            raise PeebError("libpeeb200.so ABI version mismatch; rebuild with `python -m codec_tcc_b200.build --force`")
        _lib = L
    return _lib


def last_error() -> str:
    return (lib().peeb_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc == PEEB_OK:
        return
    msg = last_error()
    if rc == PEEB_E_INVALID:
        raise ValueError(f"{what}: {msg}" if what else msg)
    raise PeebError(f"{what}: {msg} (code {rc})" if what else f"{msg} (code {rc})")


# ---------------------------------------------------------------- workspaces
class Workspace:
    """One opaque peeb_ws: device scratch, streams, events.  Thread compatible."""

    def __init__(self, device: int):
        self.device = device
        h = _vp()
        check(lib().peeb_ws_create(device, C.byref(h)), "peeb_ws_create")
        self.handle = h
        self._finalizer = weakref.finalize(self, lib().peeb_ws_destroy, h)

    def sync(self):
        check(lib().peeb_ws_sync(self.handle), "peeb_ws_sync")

    @property
    def stream(self) -> int:
        return lib().peeb_ws_stream(self.handle) or 0

    def set_option(self, name: str, value: bool):
        """'bulk': TMA bulk copies for band staging (default on)."""
        opt = {"bulk": 0}[name]
        check(lib().peeb_ws_set_option(self.handle, opt, 1 if value else 0), "peeb_ws_set_option")

    # profiling counters (bench.py's roofline leg)
    def prof_enable(self, on: bool):
        check(lib().peeb_prof_enable(self.handle, 1 if on else 0))

    def step_counters(self, on: bool):
        """{steps, edge, redone}: warp-steps of the PEE embed kernel counted since the last call with on=True."""
        out = (C.c_uint64 * 3)()
        check(lib().peeb_pee_step_counters(self.handle, 1 if on else 0, out), "peeb_pee_step_counters")
        return {"steps": int(out[0]), "edge": int(out[1]), "redone": int(out[2])}

    def prof_report(self):
        out = {}
        for slot in range(20):
            ms, calls = C.c_double(0), C.c_longlong(0)
            check(lib().peeb_prof_get(self.handle, slot, C.byref(ms), C.byref(calls)))
            if calls.value:
                out[lib().peeb_prof_name(slot).decode()] = (ms.value, calls.value)
        return out


_tls = threading.local()


def default_device() -> int:
    env = os.environ.get("PEEB_DEVICE")
    if env is not None:
        return int(env)
    lr = os.environ.get("LOCAL_RANK")
    if lr is not None:
        n = device_count()
        return int(lr) % max(n, 1)
    return 0


def device_count() -> int:
    n = _i32(0)
    check(lib().peeb_device_count(C.byref(n)), "peeb_device_count")
    return n.value


def workspace(device: int | None = None) -> Workspace:
    """Per-thread, per-device workspace cache."""
    if device is None:
        device = default_device()
    cache = getattr(_tls, "ws", None)
    if cache is None:
        cache = _tls.ws = {}
    ws = cache.get(device)
    if ws is None:
        ws = cache[device] = Workspace(device)
    return ws


class DeviceBuffer:
    """Plain device memory owned by Python (cudaMalloc through the C ABI): for
    device-resident chains of calls in the numpy API that should not bounce
    intermediates through the host."""

    def __init__(self, ws: Workspace, nbytes: int):
        self.ws, self.nbytes = ws, int(nbytes)
        p = _vp()
        check(lib().peeb_dev_alloc(ws.handle, max(self.nbytes, 1), C.byref(p)), "peeb_dev_alloc")
        self.ptr = p.value
        self._finalizer = weakref.finalize(self, lib().peeb_dev_free, ws.handle, p)

    def upload(self, arr: np.ndarray, offset: int = 0):
        arr = np.ascontiguousarray(arr)
        check(lib().peeb_memcpy_h2d(self.ws.handle, self.ptr + offset, arr.ctypes.data, arr.nbytes, self.ws.stream),
              "peeb_memcpy_h2d")

    def download(self, arr: np.ndarray, offset: int = 0):
        """Copies into ``arr`` (C-contiguous) and waits for the stream."""
        check(lib().peeb_memcpy_d2h(self.ws.handle, arr.ctypes.data, self.ptr + offset, arr.nbytes, self.ws.stream),
              "peeb_memcpy_d2h")
        return arr

    def free(self):
        self._finalizer()


# ---------------------------------------------------------------- helpers
def ptr(a) -> int:
    """Address of a numpy array's data (or pass an int through)."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array in page-locked host memory (full-speed, truly asynchronous
    PCIe copies).  Freed when the array (and every view of it) is gone."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = _vp()
    check(lib().peeb_host_alloc(max(n, 1), C.byref(p)), "peeb_host_alloc")
    buf = (C.c_ubyte * max(n, 1)).from_address(p.value)
    weakref.finalize(buf, lib().peeb_host_free, p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr


# Large result arrays of the numpy API live in page-locked memory taken from a small pool: a device->host copy into a
# fresh pageable array runs at a few GB/s (staged by the driver, plus the page faults of first touch), page-locking
# itself costs milliseconds per call -- a pooled pinned block costs neither.  A block goes back to the pool when the
# array (and every view of it) is gone.  PEEB_PINNED_POOL_MB caps what the pool keeps (default 2048; 0: plain numpy).
_pool_lock = threading.Lock()
_pool_free: dict = {}
_pool_bytes = 0
_POOL_MIN = 4 << 20
_POOL_GRAIN = 2 << 20


def _pool_cap() -> int:
    return int(os.environ.get("PEEB_PINNED_POOL_MB", "2048")) << 20


def _pool_release(addr: int, nbytes: int) -> None:
    global _pool_bytes
    with _pool_lock:
        if _pool_bytes + nbytes <= _pool_cap():
            _pool_free.setdefault(nbytes, []).append(addr)
            _pool_bytes += nbytes
            return
    try:
        lib().peeb_host_free(_vp(addr))
    except Exception:  # noqa: BLE001 -- interpreter shutdown
        pass


def out_empty(shape, dtype) -> np.ndarray:
    """An uninitialised result array: pooled page-locked memory for large results, ``np.empty`` for small ones."""
    global _pool_bytes
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    n = count * dtype.itemsize
    if n < _POOL_MIN or _pool_cap() == 0:
        return np.empty(shape, dtype)
    bucket = (n + _POOL_GRAIN - 1) // _POOL_GRAIN * _POOL_GRAIN
    addr = None
    with _pool_lock:
        free = _pool_free.get(bucket)
        if free:
            addr = free.pop()
            _pool_bytes -= bucket
    if addr is None:
        p = _vp()
        rc = lib().peeb_host_alloc(bucket, C.byref(p))
        if rc != PEEB_OK:      # the host would not pin more memory: an ordinary array works, only slower
            return np.empty(shape, dtype)
        addr = p.value
    buf = (C.c_ubyte * bucket).from_address(addr)
    weakref.finalize(buf, _pool_release, addr, bucket)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def payload_bytes(n_bits: int) -> int:
    return int(lib().peeb_payload_bytes(int(n_bits)))


def as_image(a, name="image") -> np.ndarray:
    """C-contiguous uint8/uint16 view/copy; anything else is rejected the way
    the reference rejects it (src/codec.py:34-37)."""
    a = np.asarray(a)
    if a.dtype not in (np.uint8, np.uint16):
        raise ValueError(f"{name} must be uint8 or uint16, got {a.dtype}")
    return np.ascontiguousarray(a)

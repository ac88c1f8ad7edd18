"""Device-resident variants of the PEE and metric entry points: inputs and
outputs are torch CUDA tensors (PyTorch is only the allocator and the stream
here); the work is enqueued on torch's current stream through the same C ABI
the numpy API uses.  No host<->device copy of pixel data happens in here."""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi
from ._cabi import INFO, MOMENTS, check, lib, workspace


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _raw(t: torch.Tensor, name: str) -> int:
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous CUDA tensor")
    return t.data_ptr()


def _pixels(t: torch.Tensor, name: str):
    if t.element_size() not in (1, 2) or t.is_floating_point():
        raise ValueError(f"{name} must hold 8- or 16-bit integer pixels")
    return t.element_size()


def payload_stride(max_bits: int) -> int:
    """Bytes per unit a device payload buffer needs (4-byte aligned rows with the
    slack the kernels may read, see peeb_payload_bytes)."""
    return (_cabi.payload_bytes(max_bits) + 15) // 16 * 16


def pee_embed_device(imgs, payloads, n_bits, T, bit_depth, marked=None, lm=None, info=None, shared_cover=False,
                     predictor="rhombus"):
    """imgs (n,h,w) [or (h,w) with shared_cover]; payloads (n, stride) uint8 with
    stride >= payload_stride(max n_bits); n_bits / T host int arrays; T=None: the smallest threshold that holds
    each unit's payload is chosen on the device (see info[:, 0]; the call then synchronises the stream).
    -> (marked, lm, info) tensors; enqueued, not synchronised."""
    dev = imgs.device
    item = _pixels(imgs, "imgs")
    nb = np.ascontiguousarray(n_bits, dtype=np.int64).reshape(-1)
    n = nb.size
    if shared_cover:
        h, w = imgs.shape
        src_stride = 0
    else:
        if imgs.shape[0] != n:
            raise ValueError("n_bits must have one entry per image")
        _, h, w = imgs.shape
        src_stride = h * w * item
    if T is None and shared_cover:
        raise ValueError("T=None needs one cover per unit")
    Ts = None if T is None else np.ascontiguousarray(np.broadcast_to(np.asarray(T, dtype=np.int32), (n,)))
    if payloads.dtype != torch.uint8 or payloads.dim() != 2 or payloads.shape[0] != n:
        raise ValueError("payloads must be (n, stride) uint8")
    if payloads.shape[1] % 4 or payloads.shape[1] < _cabi.payload_bytes(int(nb.max()) if n else 0):
        raise ValueError("payload stride too small (see payload_stride())")
    lmw = (w + 7) // 8
    if marked is None:
        marked = torch.empty((n, h, w), dtype=imgs.dtype, device=dev)
    if lm is None:
        lm = torch.empty((n, h, lmw), dtype=torch.uint8, device=dev)
    if info is None:
        info = torch.empty((n, INFO), dtype=torch.int64, device=dev)
    ws = workspace(dev.index if dev.index is not None else torch.cuda.current_device())
    if predictor == "med" and shared_cover:
        raise ValueError("shared_cover is not supported with predictor='med'")
    fn = lib().peeb_pee_med_embed_batch if predictor == "med" else lib().peeb_pee_embed_batch
    check(fn(
        ws.handle, _raw(imgs, "imgs"), src_stride, n, h, w, item, int(bit_depth),
        Ts.ctypes.data if Ts is not None else None, nb.ctypes.data,
        _raw(payloads, "payloads"), payloads.shape[1], _raw(marked, "marked") if marked is not False else None,
        h * w * item, _raw(lm, "lm") if lm is not False else None, h * lmw, _raw(info, "info"), _stream(dev)),
        "peeb_pee_embed_batch")
    return marked, lm, info


def pee_extract_device(marked, lm, T, n_bits, bit_depth, payload_out=None, recovered=None, info=None,
                       predictor="rhombus"):
    """-> (payload_out (n, stride) uint8, recovered, info); enqueued only."""
    dev = marked.device
    item = _pixels(marked, "marked")
    n, h, w = marked.shape
    nb = np.ascontiguousarray(n_bits, dtype=np.int64).reshape(-1)
    Ts = np.ascontiguousarray(np.broadcast_to(np.asarray(T, dtype=np.int32), (n,)))
    lmw = (w + 7) // 8
    if payload_out is None:
        payload_out = torch.empty((n, payload_stride(int(nb.max()) if n else 0)), dtype=torch.uint8, device=dev)
    if recovered is None:
        recovered = torch.empty_like(marked)
    if info is None:
        info = torch.empty((n, INFO), dtype=torch.int64, device=dev)
    ws = workspace(dev.index if dev.index is not None else torch.cuda.current_device())
    fn = lib().peeb_pee_med_extract_batch if predictor == "med" else lib().peeb_pee_extract_batch
    check(fn(
        ws.handle, _raw(marked, "marked"), h * w * item, n, h, w, item, int(bit_depth), Ts.ctypes.data,
        nb.ctypes.data, _raw(lm, "lm"), h * lmw, _raw(payload_out, "payload_out"), payload_out.shape[1],
        _raw(recovered, "recovered") if recovered is not False else None, h * w * item, _raw(info, "info"),
        _stream(dev)), "peeb_pee_extract_batch")
    return payload_out, recovered, info


def moments_device(a, b, out=None, full=True):
    """Per-image integer moments of two (n, ...) pixel tensors -> (n, 12) int64.
    ``full=False``: the MSE-only subset (SSE, maxima, n), HBM bound."""
    dev = a.device
    item = _pixels(a, "a")
    if a.shape != b.shape or a.element_size() != b.element_size():
        raise ValueError("a and b must have the same shape and item size")
    n = a.shape[0]
    per = a[0].numel()
    if out is None:
        out = torch.empty((n, MOMENTS), dtype=torch.int64, device=dev)
    ws = workspace(dev.index if dev.index is not None else torch.cuda.current_device())
    fn = lib().peeb_moments_batch if full else lib().peeb_sse_batch
    check(fn(ws.handle, _raw(a, "a"), _raw(b, "b"), per, item, n, per, per, _raw(out, "out"), _stream(dev)),
          "peeb_moments_batch")
    return out


def bitmap_encode_device(elements, n=None, packed=False, blob=None):
    """Side bitmap (uint8 CUDA tensor, non-zero = 1; or np.packbits bytes holding ``n`` elements with
    ``packed``) -> ("PBR1" blob as a uint8 CUDA tensor view of the right length).  Synchronises the
    stream once to learn the blob's size."""
    import ctypes as C

    dev = elements.device
    if elements.dtype != torch.uint8:
        raise ValueError("elements must be a uint8 tensor")
    if packed:
        n = elements.numel() * 8 if n is None else int(n)
        if (n + 7) // 8 != elements.numel():
            raise ValueError("n does not match the packed tensor's length")
    else:
        n = elements.numel()
    cap = int(lib().peeb_bitmap_blob_bound(n))
    if blob is None:
        blob = torch.empty(cap, dtype=torch.uint8, device=dev)
    elif blob.numel() < cap:
        raise ValueError("blob is smaller than peeb_bitmap_blob_bound(n)")
    ws = workspace(dev.index if dev.index is not None else torch.cuda.current_device())
    got = C.c_int64(0)
    check(lib().peeb_bitmap_encode(ws.handle, _raw(elements, "elements") if n else None, n, 1 if packed else 0,
                                   _raw(blob, "blob"), blob.numel(), C.byref(got), _stream(dev)), "peeb_bitmap_encode")
    return blob[:got.value]


def bitmap_decode_device(blob, n, packed=False, out=None):
    """"PBR1" blob (uint8 CUDA tensor) -> n bytes of 0/1, or the packed bits.  Synchronises the stream."""
    dev = blob.device
    n = int(n)
    size = (n + 7) // 8 if packed else n
    if out is None:
        out = torch.empty(size, dtype=torch.uint8, device=dev)
    elif out.numel() != size or out.dtype != torch.uint8:
        raise ValueError("out must be a uint8 tensor of the decoded size")
    ws = workspace(dev.index if dev.index is not None else torch.cuda.current_device())
    check(lib().peeb_bitmap_decode(ws.handle, _raw(blob, "blob"), blob.numel(), _raw(out, "out") if size else None, n,
                                   1 if packed else 0, _stream(dev)), "peeb_bitmap_decode")
    return out

"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* reference modules ``/root/reference/src/codec.py`` and
``/root/reference/src/mse.py`` so that (a) the numpy restatements in this
directory can be validated against the real thing and (b) golden vectors can be
generated (``tests/golden/make_golden.py``).

The reference imports ``pydicom`` (and a pylibjpeg handler) at module top
(src/codec.py:4,10-16; src/mse.py:7).  pydicom is not installed here and none
of the pixel-array functions on the hot path touch it, so empty stand-in
modules are registered before the files are executed.  Only the DICOM / codec
I/O functions (out of scope) become
unusable.

``/root/reference`` exists only in the build container; on the GPU box
``available()`` returns False and every caller must skip.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CODEC_TCC_REFERENCE", "/root/reference")

_cache: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "codec.py"))


def _install_pydicom_stub() -> None:
    if "pydicom" in sys.modules and not getattr(sys.modules["pydicom"], "_peeb_stub", False):
        return  # a real pydicom is importable: leave it alone

    def mod(name: str, **attrs):
        m = types.ModuleType(name)
        m._peeb_stub = True
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Nothing:  # placeholder for FileDataset & friends
        def __init__(self, *a, **k):
            raise RuntimeError("pydicom is stubbed: DICOM I/O is out of scope")

    root = mod("pydicom", dcmread=_Nothing)
    root.dataset = mod("pydicom.dataset", FileDataset=_Nothing, FileMetaDataset=_Nothing)
    root.uid = mod(
        "pydicom.uid",
        ExplicitVRLittleEndian="1.2.840.10008.1.2.1",
        JPEGLSLossless="1.2.840.10008.1.2.4.80",
        JPEG2000Lossless="1.2.840.10008.1.2.4.90",
        DeflatedExplicitVRLittleEndian="1.2.840.10008.1.2.1.99",
        PYDICOM_IMPLEMENTATION_UID="0.0",
        generate_uid=lambda *a, **k: "0.0",
    )
    root.encaps = mod("pydicom.encaps", encapsulate=_Nothing)
    root.config = mod("pydicom.config", image_handlers=[])
    root.pixel_data_handlers = mod("pydicom.pixel_data_handlers")
    root.pixel_data_handlers.pylibjpeg_handler = mod(
        "pydicom.pixel_data_handlers.pylibjpeg_handler"
    )


def _load(name: str):
    if name in _cache:
        return _cache[name]
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    _install_pydicom_stub()
    path = os.path.join(REFERENCE_ROOT, "src", f"{name}.py")
    spec = importlib.util.spec_from_file_location(f"_codec_tcc_reference_{name}", path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    _cache[name] = module
    return module


def codec():
    """The reference's ``src/codec.py`` as a module object."""
    return _load("codec")


def mse():
    """The reference's ``src/mse.py`` as a module object."""
    return _load("mse")


@contextlib.contextmanager
def quiet():
    """The reference prints from inside its numerics (src/codec.py:568,577-578;
    src/mse.py:102).  Swallow that while generating vectors."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf


def read_fixture_pixels(name: str):
    """Raw pixel data of the two committed reference images, without pydicom
    (offsets from SURVEY.md section 2.3): pe.dcm = 512x512 u16 at byte 7010,
    torax.dcm = 512x512 u8 at byte 888."""
    import numpy as np

    table = {"pe": ("pe.dcm", "<u2", 7010), "torax": ("torax.dcm", "u1", 888)}
    fname, dtype, off = table[name]
    with open(os.path.join(REFERENCE_ROOT, "images", fname), "rb") as f:
        raw = f.read()
    arr = np.frombuffer(raw, dtype=dtype, count=512 * 512, offset=off).reshape(512, 512)
    return np.ascontiguousarray(arr)

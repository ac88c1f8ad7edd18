"""codec_tcc_b200 -- B200 (sm_100a) implementation of the pixel-array hot path
of wesleyfn/codec-tcc: reversible data hiding in 8/12/16-bit medical images and
its distortion metrics, behind the reference's numpy-in / numpy-out API.

    from codec_tcc_b200 import codec, mse, pee

``codec``  mirrors src/codec.py (bit-plane LSB embedding, rows a5-a9),
``mse``    mirrors src/mse.py's ``AnalisadorMSE`` (rows a1-a4),
``pee``    is the Prediction-Error-Expansion pipeline (row a10),
``shard``  partitions image batches over the GPUs of one box.

All per-pixel work runs in hand-written CUDA kernels (csrc/) loaded through a
C ABI (include/peeb200.h); importing the package does not touch the GPU, the
first compute call does, and raises if no sm_100 device or no built library
is available.
"""
from . import synth  # noqa: F401  (numpy only)

__all__ = ["codec", "mse", "pee", "shard", "synth", "build_library"]
__version__ = "0.1.0"


def build_library(force: bool = False, verbose: bool = False) -> str:
    from . import build

    return build.build(force=force, verbose=verbose)


def __getattr__(name):
    if name in ("codec", "mse", "pee", "shard", "_cabi"):
        import importlib

        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)

"""swiftover_b200 — B200-native liftover engine for swiftover's hot path.

Only what the path needs: `csrc/` (CUDA kernels + the C ABI of
include/swiftover_cuda.h, built into libswiftover_cuda.so), `host/` (C++ host
mirroring the reference CLI) and this thin ctypes mirror of the reference's
`ChainFile` / `liftBED` interface used by the tests and bench.py.
"""
from .chain import ChainFile, LiftedInterval, SwiftoverError  # noqa: F401
from .bed import liftBED  # noqa: F401
from .vcf import liftVCF  # noqa: F401
from .maf import liftMAF  # noqa: F401

__all__ = ["ChainFile", "LiftedInterval", "SwiftoverError", "liftBED", "liftVCF", "liftMAF"]

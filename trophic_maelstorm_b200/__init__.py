"""trophic_maelstorm_b200 -- B200 (sm_100a) Dornic lattice integrator behind a C ABI.

The product is the shared library ``libtm_lattice.so`` (CUDA kernels + ``extern "C"`` layer,
``include/tm_lattice.h``).  This package is only the thin ctypes view of that ABI used by the
tests and the benchmark; the reference-facing host layer is C++ (``host/``).

There is no CPU path: importing works anywhere (so the symbol table can be checked), but every
compute call fails loudly without the CUDA library and a GPU.
"""
from .lattice import (  # noqa: F401
    Lattice, Params, Stats, Spread, TmError, lib, lib_path, rietkerk_params, dp_params, dp_noise_off,
    MODEL_DP, MODEL_DP_CUBIC, MODEL_RIETKERK, MODEL_RIETKERK_DET,
)
from .fke import FkeMover, FkeTerms, FKE_ABI_SYMBOLS  # noqa: F401,E402
from .consumer import (ConsumerOps, ConsParams, GrazParams, CONS_ABI_SYMBOLS,   # noqa: F401,E402
                       ConsumerOps3, Cons3Params, CONS3_ABI_SYMBOLS)

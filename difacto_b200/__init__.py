"""difacto-b200: B200-native FM-SGD engine behind dmlc/difacto's Learner/Loss/Updater/Store API.

The product is the C-ABI shared library difacto_b200/lib/libdifacto_b200.so (hand-written
sm_100a CUDA, see include/difacto_b200.h) plus the host-side C++ mirror of the reference
interfaces under difacto_b200/host/.  This Python package is a thin ctypes binding used by
the tests, bench.py and the torch.distributed sharded store; it never falls back to a CPU
implementation -- importing ``difacto_b200.capi`` raises if the CUDA library is missing.
"""
__all__ = ["capi"]

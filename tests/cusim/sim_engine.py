"""Test-only: an ``Engine`` bound to the CPU-simulated build of the kernels (tests/cusim)."""
from audiotools_b200 import _lib
from audiotools_b200.engine import Engine
from tests.cusim import build_sim

_SIM = None


def sim_engine() -> Engine:
    global _SIM
    if _SIM is None:
        _SIM = Engine(_lib.B2ALibrary(build_sim.build()), require_cuda=False)
    return _SIM

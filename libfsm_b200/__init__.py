"""libfsm_b200 -- B200 (sm_100a) engine for the fsm_exec / fsm_determinise hot path of
katef/libfsm, behind a C ABI (include/fsm_b200.h).  Python here is plumbing only: the
ctypes binding, the flat automaton description and the multi-GPU sharding helpers.

The native library is loaded on first use of an engine name (``libfsm_b200.Dfa``, ``.plan``,
``.LIB_PATH`` ...), not at package import: ``libfsm_b200.desc`` and ``libfsm_b200.workloads`` are
plain numpy and are shared with the test oracle and with ``bench.py --impl reference``, which must
not pull the engine into its process.  There is still no CPU fallback -- the first engine name
raises ImportError when the library has not been built."""
from .desc import FlatFsm, RESULT_DTYPE

_ENGINE = ("Dfa", "plan", "determinise", "determinise_stats", "minimise", "minimise_stats", "device_count",
           "set_exec_variant", "launch_count", "results_from_torch", "stream_map_arrays", "StateLimitReached", "VARIANTS", "load_dfavm")
_NATIVE = ("FsmB200Error", "LIB_PATH", "ABI_SYMBOLS")

__all__ = ["FlatFsm", "RESULT_DTYPE", *_ENGINE, *_NATIVE]


def __getattr__(name):
    if name in _ENGINE:
        from . import engine
        return getattr(engine, name)
    if name in _NATIVE:
        from . import _native
        return getattr(_native, name)
    raise AttributeError(f"module 'libfsm_b200' has no attribute {name!r}")

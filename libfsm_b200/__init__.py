"""libfsm_b200 -- B200 (sm_100a) engine for the fsm_exec / fsm_determinise hot path of
katef/libfsm, behind a C ABI (include/fsm_b200.h).  Python here is plumbing only: the
ctypes binding, the flat automaton description and the multi-GPU sharding helpers."""
from .desc import FlatFsm, RESULT_DTYPE
from .engine import (Dfa, plan, determinise, determinise_stats, minimise, minimise_stats, device_count, set_exec_variant,
                     launch_count, results_from_torch, StateLimitReached, VARIANTS)
from ._native import FsmB200Error, LIB_PATH, ABI_SYMBOLS

__all__ = ["FlatFsm", "RESULT_DTYPE", "Dfa", "plan", "determinise", "determinise_stats", "minimise", "minimise_stats", "device_count",
           "set_exec_variant", "launch_count", "results_from_torch", "StateLimitReached",
           "VARIANTS", "FsmB200Error", "LIB_PATH", "ABI_SYMBOLS"]

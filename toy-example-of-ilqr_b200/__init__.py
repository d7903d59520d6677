"""cilqr_b200 — batched CILQR solver native to NVIDIA B200 (sm_100a).

The product is `libcilqr_b200.so` (csrc/, C ABI in include/cilqr_b200.h); this package is the
host-side mirror used by tests and bench.py:  binding (ctypes), scenario (YAML templates ->
arrays, synthetic batches), templates (the reference's four scenarios).
The directory name carries a hyphen, so import it through the top-level `cilqr_b200` module.
"""
from . import scenario, shard, templates  # noqa: F401
from .binding import (  # noqa: F401
    BatchSolver, CILQRSolver, CilqrError, CilqrParams, SolveResult, EXPORTS, LIB_PATH, STATUS_NAMES,
    EXIT_NAMES, load_library,
)
from .scenario import (BatchProblem, Scenario, TemplateData, synthetic_batch, single_problem, get_scenario,  # noqa: F401
                       synth_spec, generate_host)

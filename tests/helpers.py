"""Shared test helpers (no reference needed): re-export of efficient_slowfast_b200/workloads.py."""
from efficient_slowfast_b200.workloads import (GOLDEN_DIR, case_cfg, case_inputs, case_model_and_weights,  # noqa: F401
                                               load_golden, rel_err)

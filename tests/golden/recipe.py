"""Re-export of the package's synthetic-workload recipe (efficient_slowfast_b200/workloads.py) under the name the
golden generator and the tests import."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from efficient_slowfast_b200.workloads import (CASES, _gen, pack_pathway_output, sample_indices, seeded_clip,  # noqa: E402,F401
                                               seeded_state_dict)

#!/usr/bin/env python
"""Generates tests/golden/*.npz: outputs of the CPU restatement oracle (oracle/oracle.cpp) on small seeded inputs, plus
tests/golden/reference_kats.json: the known-answer values the reference's own tests hold for this path (file:line).

The reference cannot be built or imported here (dune-common / dune-grid / dune-xt are absent), so
the fixtures are oracle outputs, and the oracle itself is pinned against reference_kats.json by
tests/test_oracle_golden.py.  The fixtures freeze the oracle: a change in oracle.cpp that moves any value shows up in
tests/test_golden_fixtures.py on the CPU, and the CUDA path is compared against the same frozen numbers on the GPU.

  python tools/make_golden.py      # rewrites tests/golden/
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from dune_gdt_b200 import descriptors as D  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 20251017


sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_cases  # noqa: E402  (tests/golden_cases.py: the seeded cases)


def main():
    os.makedirs(OUT, exist_ok=True)
    oracle.build()
    for name, case in golden_cases.CASES.items():
        arrays = golden_cases.run_oracle(oracle, case)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        print(name, {k: v.shape for k, v in arrays.items()})
    kats = {
        "_comment": "known-answer values of the reference's own tests for this path (relative to /root/reference)",
        "q2_stiffness_9x2": {"source": "dune/gdt/test/integrands/integrands_laplace.cc:131-133", "grid": "[0,3]x[0,1], 9x2 Yasp cells",
                             "min": -1.896296296296300, "max": 6.162962962962970, "sum_sq": 1704.099039780521, "tol": [1e-13, 1e-13, 5e-12]},
        "q2_mass_9x2": {"source": "dune/gdt/test/integrands/integrands_product.cc:122-124", "grid": "[0,3]x[0,1], 9x2 Yasp cells",
                        "min": -0.002962962962963, "max": 0.047407407407407, "sum_sq": 0.066475994513031, "tol": [1e-13, 1e-13, 1e-13]},
        "esv2007_swipdg_h1_semi_error": {"source": "dune/gdt/test/stationary-heat-equation/stationary_heat_equation__ESV2007__table_1.mini:31-36",
                                         "n": [8, 16, 32], "values": [2.52e-01, 1.26e-01, 6.30e-02]},
        "linear_transport_1d_fv": {"source": "dune/gdt/test/linear-transport/linear_transport__1d__explicit__fv.mini:8-14", "n": [16, 32, 64],
                                   "L_infty_L_2": [1.77e-01, 1.25e-01, 8.84e-02], "num_timesteps": [18, 34, 66], "CFL": [2.0, 2.0, 2.0],
                                   "rel_mass_conserv_error": [0, 0, 0]},
        "burgers_1d_fv_upwind": {"source": "dune/gdt/test/burgers/burgers__1d__explicit__fv.mini:3-15", "n": [16, 32],
                                 "use_fixed_dt": 0.0096815612792968738, "dt_factor": 0.99, "num_timesteps": [107, 107], "CFL": [2.31e-01, 4.80e-01]},
        "burgers_1d_fv_lax_friedrichs": {"source": "dune/gdt/test/burgers/burgers__1d__explicit__fv.mini:18-30", "n": [16, 32],
                                         "use_fixed_dt": 0.009193328857421872, "dt_factor": 0.99, "num_timesteps": [112, 112], "CFL": [2.20e-01, 4.56e-01]},
    }
    with open(os.path.join(OUT, "reference_kats.json"), "w") as f:
        json.dump(kats, f, indent=1)


if __name__ == "__main__":
    main()

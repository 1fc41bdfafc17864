"""Generate tests/golden/codegen_kat.json by RUNNING THE REFERENCE'S OWN generated closed-form
dynamics (compiled from /root/reference/src/Codegen by oracle/Makefile into oracle/_ref/).
These are the known-answer vectors that pin the oracle (reference test:
UnitTests/testReflectedInertiaAlgos.cpp:144-222, tolerance 1e-5 there).
Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding  # noqa: E402

CASES = [("revolute_chain_with_rotor_2", "RevWithRotors2Dof", 2),
         ("revolute_chain_with_rotor_4", "RevWithRotors4Dof", 4),
         ("revolute_pair_chain_with_rotor_2", "RevPairWithRotors2Dof", 2),
         ("revolute_pair_chain_with_rotor_4", "RevPairWithRotors4Dof", 4)]


def main():
    binding.build()
    assert binding.reference_codegen_available(), "oracle/_ref was not built (no /root/reference?)"
    rng = np.random.default_rng(20261017)
    out = {"source": "/root/reference/src/Codegen/rev_{w,pair_w}_rotor_{2,4}dof_{FD,ID}.cpp (CasADi 3.6.3 generated)",
           "cases": []}
    # Appendix-D style fixed vectors first, then random states in the reference's ranges [-1, 1]
    fixed = {2: ([0.3, -0.2], [0.1, 0.4], [1.0, -0.5]),
             4: ([0.3, -0.2, 0.5, -0.7], [0.1, 0.4, -0.3, 0.2], [1.0, -0.5, 0.25, 0.75])}
    for robot, prefix, n in CASES:
        states = [tuple(np.array(x) for x in fixed[n])]
        for _ in range(24):
            states.append((rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)))
        recs = []
        for y, yd, tau in states:
            ydd = binding.reference_codegen(prefix + "FwdDyn", [y, yd, tau], n)
            tau_back = binding.reference_codegen(prefix + "InvDyn", [y, yd, ydd], n)
            recs.append({"y": y.tolist(), "yd": yd.tolist(), "tau": tau.tolist(), "ydd_fwd": ydd.tolist(),
                         "tau_inv_of_ydd": tau_back.tolist()})
        out["cases"].append({"robot": robot, "function_prefix": prefix, "n": n, "states": recs})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "codegen_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()

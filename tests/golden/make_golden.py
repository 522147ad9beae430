"""Regenerates the fixtures in this directory.  Run from the repo root: python tests/golden/make_golden.py

kat_expected.json   the analytic optima the reference's own test prints (test/test.cpp:34, 57, 82,
                    108, 134, 159, 183) -- copied by hand from SURVEY.md §4, NOT produced by any code
                    of this repo; case 5 uses the true optimum (2/3, 5/3, 8/3), the reference only
                    prints "near (1,2,2) or other feasible".
patterns.json       dims + SHA-256 of the CCS index arrays of the three benchmark problems as the
                    oracle (CPU restatement of SQPOptimizationSolver.cpp:12-92) assembles them.
oracle_sqp.npz      oracle SQP iterates for 2 seeded quadrotor instances (regression pin of the
                    restatement itself; the reference cannot be run here -- parity unpinned).
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import _oracle  # noqa: E402

KAT = {
    "1": {"ref": "test/test.cpp:13-36", "x": [0.5, 0.5]},
    "2": {"ref": "test/test.cpp:38-59", "x": [3.0, -2.0]},
    "3": {"ref": "test/test.cpp:61-84", "x": [2.0, 3.0]},
    "4": {"ref": "test/test.cpp:86-110", "x": [1.0, 2.0]},
    "5": {"ref": "test/test.cpp:112-136", "x": [2.0 / 3.0, 5.0 / 3.0, 8.0 / 3.0]},
    "6": {"ref": "test/test.cpp:138-161", "x": [5.0, 0.0]},
    "7": {"ref": "test/test.cpp:163-185", "x": [2.0, 3.0]},
}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.int32).tobytes()).hexdigest()


def main():
    (HERE / "kat_expected.json").write_text(json.dumps(KAT, indent=1) + "\n")
    pats = {}
    for name in ("quadrotor", "cartpole", "centroidal"):
        o = _oracle.OracleProblem(name)
        pats[name] = dict(np=o.np_, nf=o.nf, horizon=o.horizon, ng=o.ng, n=o.n, m=o.m, nnz_h=o.nnz_h, nnz_a=o.nnz_a,
                          h_colptr=digest(o.h_colptr), h_rowidx=digest(o.h_rowidx), a_colptr=digest(o.a_colptr),
                          a_rowidx=digest(o.a_rowidx))
    (HERE / "patterns.json").write_text(json.dumps(pats, indent=1) + "\n")
    o = _oracle.OracleProblem("quadrotor", alpha=0.1, step_num=10)
    frames, refs = o.sample_inputs(2, 0xB200)
    x, f, st = o.solve_batch(frames, refs)
    o.set_schedule(5, 1.0)
    x1, f1, st1 = o.solve_batch(frames, refs)
    np.savez_compressed(HERE / "oracle_sqp.npz", frames=frames, refs=refs, x_a01=x, f_a01=f, st_a01=st, x_a1=x1, f_a1=f1,
                        st_a1=st1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- numpy fp64 restatement of the reference's SQP / Schur / GBD-PCG path.

This package is the *checker*, never the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product package
(`trajoptmpcreference_b200`) never imports it and raises if its CUDA library is missing.

Every function cites the reference file:line (under /root/reference) whose arithmetic it restates.  The
restatement is block-structured (it never forms the dense KKT / Schur matrices the reference builds) and is
vectorised over knot points; `oracle.dense` re-assembles the reference's dense formulation for cross-checks.

Parity status: PINNED.  The oracle is checked (tests/test_oracle_golden.py) against
  * golden vectors recorded by the reference's authors (`data/3`, `data/4`, re-packed into tests/golden/ref_data*.npz),
  * outputs of the unmodified reference imported in the build container (tests/golden/make_golden.py -> *.npz),
for: rigid-body dynamics (rnea, minv, rnea_grad, EE kinematics), integrators 0-3 (2 / 3 literally as the reference computes them; 4 raises like it), QuadraticCost, UrdfCost (n=2),
torque box limits (1-DoF, QUADRATIC_PENALTY / AUGMENTED_LAGRANGIAN, and the hard ACTIVE_SET mode of the exact methods N / S,
where the literal dense form is bit-identical to the reference), KKT blocks, Schur complement, J/BJ/SS preconditioners, PCG
traces, and complete SQP solves (iteration counts, alpha sequences, J, c, x, u) within the reference's own measured 1-ulp floor
(tests/golden/floor.json, scripts/parity_floor.py).
UNPINNED (no working reference code; the oracle itself is the spec): box limits with more than one
constrained coordinate, joint / velocity limits, the n-link end-effector cost and its exact Hessian (hess_mode 1), iLQR, MPC.
"""

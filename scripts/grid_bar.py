#!/usr/bin/env python
"""The GRiD bar: the reference's emitted forward_dynamics_gradient_kernel<double> (baseline/_ref/grid_bench, built by
baseline/make_grid.py for sm_100a) next to our k_fd + k_fd_grad on the same box, 8192 x 64 knot points, fp64 (and fp32).

  python scripts/grid_bar.py            ->  gpurun_out/grid_bar.json + a table

Ours is timed through the C ABI: b2t_stage_dynamics with CUDA events around the two launches (kernel families `fd`, `fd_grad`);
it produces MORE than the GRiD kernel: x+ (integrator), q'', M^-1 and the full dq''/d(q, q', u) (n x 3n, incl. M^-1) against GRiD's
dq''/d(q, q') (n x 2n).  Only N-1 of the N knots of an instance carry dynamics, so ours evaluates 8192 x 63 knot points."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trajoptmpcreference_b200 as t            # noqa: E402

B, N = 8192, 64
out = {}
for dtype, binary in (("f64", "grid_bench"), ("f32", "grid_bench_f32")):
    exe = os.path.join(ROOT, "baseline", "_ref", binary)
    res = {}
    for blocks in (0, 148 * 8):               # one block per knot point (the emitted wrapper's usage) and a persistent grid-stride launch
        r = subprocess.run([exe, str(B * N), str(blocks)], capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        res["blocks_%d" % blocks] = json.loads(line[-1]) if line else {"error": (r.stdout + r.stderr)[-300:]}
    plant = t.URDFPlant(options={"path_to_urdf": "arm6"})
    cost = t.QuadraticCost(np.eye(12), 100.0 * np.eye(12), 0.1 * np.eye(6), np.zeros(12))
    s = t.BatchSolver(plant, cost, None, N=N, dt=0.1, batch=B, dtype=dtype)
    rng = np.random.default_rng(1337)
    s.set_trajectory(rng.uniform(-0.8, 0.8, (B, 12, N)), rng.uniform(-0.8, 0.8, (B, 6, N - 1)))
    best = None
    for rep in range(6):
        s.set_profiling(True)
        s.stage_dynamics()
        kt = s.kernel_times()
        cur = (kt["fd"][0] + kt["fd_grad"][0], kt["fd"][0], kt["fd_grad"][0])
        if rep >= 2 and (best is None or cur[0] < best[0]):
            best = cur
    s.set_profiling(False)
    knots = B * (N - 1)
    ours = {"ms_total": 1e3 * best[0], "ms_k_fd": 1e3 * best[1], "ms_k_fd_grad": 1e3 * best[2], "knot_points": knots, "ns_per_knot": 1e9 * best[0] / knots}
    g = min((v for v in res.values() if "ms_best" in v), key=lambda v: v["ms_best"])
    out[dtype] = {"grid": res, "ours": ours, "speedup_per_knot": g["ns_per_knot"] / ours["ns_per_knot"]}
    print("%s: GRiD forward_dynamics_gradient_kernel %.3f ms for %d knot points (%.2f ns / knot; one block per knot %.3f ms, persistent %.3f ms)  |  ours k_fd %.3f + k_fd_grad %.3f = %.3f ms for %d knot points (%.2f ns / knot)  ->  %.1fx" %
          (dtype, g["ms_best"], g["knots"], g["ns_per_knot"], res["blocks_0"].get("ms_best", float("nan")), res["blocks_%d" % (148 * 8)].get("ms_best", float("nan")),
           ours["ms_k_fd"], ours["ms_k_fd_grad"], ours["ms_total"], knots, ours["ns_per_knot"], out[dtype]["speedup_per_knot"]))
    s.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "grid_bar.json"), "w") as f:
    json.dump(out, f, indent=1)

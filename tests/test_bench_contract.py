"""CPU checks of bench.py's bookkeeping (the measurement contract): both arms describe the same config, the roofline's DRAM-traffic
figure comes from the newest committed ncu capture of the kernel that actually runs, the CPU pool measures throughput of a stream
(not the slowest member of a small group), and the flop model matches SURVEY.md section 8d."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    a = argparse.Namespace(gpus=1, batch=8192, limits=1, workload="c4", steps=3, warmup=3)
    a.__dict__.update(kw)
    return a


def test_both_arms_report_the_same_config():
    a = _args()
    c = bench.config_dict(a)
    assert c == bench.config_dict(_args())                        # no per-arm keys
    assert set(c) == {"workload", "batch_per_gpu", "knots", "method", "limits", "l2"}
    assert "C4" in c["workload"] and "N=64" in c["workload"] and c["knots"] == 64
    assert "C5" in bench.config_dict(_args(gpus=8))["workload"]


def test_traffic_comes_from_newest_capture_of_the_running_kernel():
    t, src = bench.ncu_traffic("pcg", 8192, 472, 71.3, "k_pcg3")
    assert t is not None and "r02_v9_ncu_full_k_pcg3.csv" in src
    per_inst = t / (8192 * 71.3 / 472)
    assert 1.4e5 < per_inst < 1.8e5                               # ~154 KB algorithmic per instance
    t2, src2 = bench.ncu_traffic("pcg", 8192, 472, 71.3, "k_pcg_no_such_kernel")
    assert t2 is None and "no ncu capture" in src2
    assert bench.ncu_traffic("trial_fd", 8192, 472, 71.3)[0] is None


def _sleep(x):
    time.sleep(x)
    return x


def test_pool_measures_stream_throughput():
    # 16 tasks of 0.05 s on 4 workers: ~0.2 s of wall for the 12 timed ones after 4 warm-up completions
    sec, cnt = bench.pool_throughput(_sleep, [0.05] * 16, 4, 4)
    assert cnt == 12 and 0.1 < sec < 0.6


def test_flop_model_matches_survey():
    fm = bench.flop_model(6, 64)
    assert fm["pcg_iter"] == 12 * 144 + 120 == 1848                # F_pcg_iter (SS), SURVEY.md 8d
    assert fm["trial_fd"] == 415 * 6 + 1051 * 6 + 92 * 6 * 7 + 2 * 36
    fam = bench.flops_of(fm, qp=6, pcg=563, trials=16, B=1)        # the anchor instance of SURVEY.md 8d
    assert 1.0e8 < sum(fam.values()) < 1.3e8                       # "~115 Mflop / solve"

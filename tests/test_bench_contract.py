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
    # the persistent tensor-memory kernel launches one CTA per SM: its capture carries the instance count of the launch (b2t__instances,
    # from the executed count of the work-queue ATOMG), and the per-instance traffic must come out the same as k_pcg3's
    t3, src3 = bench.ncu_traffic("pcg", 8192, 472, 71.3, "k_pcg_tm")
    assert t3 is not None and "_ncu_full_k_pcg_tm.csv" in src3
    assert 1.4e5 < t3 / (8192 * 71.3 / 472) < 1.8e5
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


def test_ticket_instances_of_a_persistent_launch(tmp_path):
    """scripts/profile_summary.ticket_instances: instances of a captured k_pcg_tm launch = work-queue tickets drawn (executed ATOMG)
    minus the terminating ticket of every half."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import profile_summary
    page = tmp_path / "sass.csv"
    page.write_text('"Kernel Name","k_pcg_tm",\n"Address","Source","# Samples","Instructions Executed"\n'
                    '"0x10","      S2R R8, SR_TID.X","0","2368"\n'
                    '"0x20","@P0   ATOMG.E.ADD.STRONG.GPU PT, R3, desc[UR8][R2.64], R9","2","1411"\n'
                    '"0x30","      DFMA R4, R4, R4, R4","9","798704"\n')
    assert profile_summary.ticket_instances(str(page), 148) == 1411 - 2 * 148

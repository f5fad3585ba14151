"""Turn the raw gpurun_out artefacts of a round into the tracked summaries under profiles/:
   python scripts/profile_summary.py <tag> <launches.csv> <full.ncu-rep | raw.csv> <kernel> [<sass page .csv.gz of a persistent kernel>]
writes profiles/<tag>_ncu_launch_summary.csv and profiles/<tag>_ncu_full_<kernel>.csv (needs ncu on PATH for the .ncu-rep)."""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum']


def launch_summary(tag, path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    hdr = rows[0]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = re.sub(r'<.*', '', r[ik]).replace('void b2t::', '').replace('void ', '')
        v = float(r[iv].replace(',', ''))
        v = {'ns': v / 1e3, 'nsecond': v / 1e3, 'us': v, 'usecond': v, 'ms': v * 1e3, 'msecond': v * 1e3}.get(r[iu], v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v for _, v in agg.values())
    out = os.path.join(ROOT, 'profiles', tag + '_ncu_launch_summary.csv')
    with open(out, 'w') as f:
        f.write('# aggregated from %s (ncu gpu__time_duration.sum)\nkernel,launches,total_us,share\n' % os.path.basename(path))
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%s,%d,%.1f,%.4f\n' % (k, n, v, v / tot))
    return out


def ticket_instances(sass_page, grid, halves_per_cta=2):
    """Persistent kernels (k_pcg_tm) launch one CTA per SM, not one per instance: the number of instances a captured launch processed
    is the number of work-queue tickets drawn (executed count of the ATOMG of the exported SASS page) minus the one terminating
    ticket per half."""
    import gzip
    op = gzip.open if sass_page.endswith('.gz') else open
    rows = list(csv.reader(l for l in op(sass_page, 'rt') if not l.startswith('==')))
    hdr = rows[1]
    isrc, ie = hdr.index('Source'), hdr.index('Instructions Executed')
    tickets = sum(float(r[ie]) for r in rows[2:] if len(r) > ie and 'ATOMG' in r[isrc])
    return int(tickets - halves_per_cta * grid)


def full_summary(tag, rep, kernel, sass_page=None):
    """rep: a .ncu-rep (read with ncu) or the raw page already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv > x_raw.csv`;
    the reports themselves exceed gpurun's 64 MiB return limit).  One column per captured launch."""
    if rep.endswith('.csv'):
        raw = open(rep).read()
    else:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(l for l in raw.splitlines() if not l.startswith('==')) if r]
    hdr, units, launches = rows[0], rows[1], rows[2:]
    keys = KEEP + [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    out = os.path.join(ROOT, 'profiles', '%s_ncu_full_%s.csv' % (tag, kernel))
    ik = hdr.index('Kernel Name')
    with open(out, 'w') as f:
        f.write('# ncu --set full --clock-control none, captured launches of %s in %s\n' % (kernel, os.path.basename(rep)))
        f.write('metric,unit,%s\n' % ','.join(re.sub(r'<.*', '', r[ik]).replace('void b2t::', '') for r in launches))
        for k in keys:
            if k in hdr:
                f.write('%s,%s,%s\n' % (k, units[hdr.index(k)], ','.join(r[hdr.index(k)].replace(',', '') for r in launches)))
        if sass_page:
            grid = float(launches[0][hdr.index('launch__grid_size')])
            f.write('b2t__instances,instance,%d\n' % ticket_instances(sass_page, grid))
    return out


if __name__ == '__main__':
    tag, launches, rep, kernel = sys.argv[1:5]
    print(launch_summary(tag, launches))
    print(full_summary(tag, rep, kernel, sys.argv[5] if len(sys.argv) > 5 else None))

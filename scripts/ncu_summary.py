"""profiles/r01_full_summary.csv (selected metrics of every kernel in the report) and profiles/fwd_kernel_traffic.json
(DRAM bytes of one forward launch, read by bench.py for roofline.traffic) from an `ncu --set full` report."""
import csv, io, json, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum'] + \
       ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
        ('long_scoreboard', 'short_scoreboard', 'wait', 'branch_resolving', 'not_selected', 'math_pipe_throttle',
         'mio_throttle', 'lg_throttle')]
OUT = sys.argv[3] if len(sys.argv) > 3 else "r02_full_summary.csv"   # e.g. r01_shade_summary.csv for the shading kernels
with open(os.path.join(ROOT, "profiles", OUT), "w") as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit'] + [r[hdr.index('Kernel Name')].split('(')[0] for r in rows[2:]])
    for m in want:
        if m in hdr:
            i = hdr.index(m)
            w.writerow([m, units[i]] + [r[i] for r in rows[2:]])

def val(r, m):
    i = hdr.index(m)
    return float(r[i].replace(',', '')) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[units[i]]

for r in rows[2:]:
    if 'trace_forward' in r[hdr.index('Kernel Name')]:
        json.dump({"kernel": "trace_forward_kernel<0,0>", "rays_per_launch": int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24,
                   "dram_bytes_read": val(r, 'dram__bytes_read.sum'), "dram_bytes_write": val(r, 'dram__bytes_write.sum'),
                   "gpu_time_ms_under_ncu": float(r[hdr.index('gpu__time_duration.sum')]),
                   "source": "profiles/" + OUT + "  (ncu --set full --clock-control none, bench.py C3 step, one forward launch)"},
                  open(os.path.join(ROOT, "profiles", "fwd_kernel_traffic.json"), "w"), indent=1)
        break
print(open(os.path.join(ROOT, "profiles", OUT)).read())

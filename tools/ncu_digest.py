"""Digest of one .ncu-rep: headline metrics, executed instructions by opcode, stall samples, hot SASS regions.
usage: ncu_digest.py <report.ncu-rep> [min_million_for_listing]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            print("%-70s %s" % (w, r[hdr.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = srows[1]
ia, isrc, isamp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
body = [r for r in srows[2:] if len(r) == len(h)]
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in body:
    s = r[isrc].strip().split()
    if not s:
        continue
    op = s[0] if not s[0].startswith('@') else s[1]
    op = '.'.join(op.split('.')[:2]) if op.startswith(('LDS', 'STS', 'STG', 'LDG')) else op.split('.')[0]
    n = int(r[ia])
    ops[op] += n
    tot += n
    samp[op] += int(r[isamp])
print("total warp instructions %.1fM" % (tot / 1e6))
for op, n in ops.most_common(18):
    print("  %-14s %8.1fM %5.1f%%  samples %d" % (op, n / 1e6, 100 * n / tot, samp[op]))
st = collections.Counter()
for r in body:
    for i, k in enumerate(h):
        if k.startswith('stall_') and 'Not' not in k:
            st[k] += int(r[i] or 0)
print("stalls:", st.most_common(8))
for s0 in range(0, len(body), 80):
    c = sum(int(r[ia]) for r in body[s0:s0 + 80])
    sm = sum(int(r[isamp]) for r in body[s0:s0 + 80])
    if c > tot * 0.02:
        print("  region %5d: %7.1fM %5.1f%% samples %d" % (s0, c / 1e6, 100 * c / tot, sm))
if len(sys.argv) > 2:
    lim = float(sys.argv[2]) * 1e6
    for i, r in enumerate(body):
        if int(r[ia]) >= lim:
            print(i, "%7.2fM" % (int(r[ia]) / 1e6), r[isamp].rjust(5), r[isrc].strip()[:110])

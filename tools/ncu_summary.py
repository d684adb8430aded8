"""Summarise an .ncu-rep: per-kernel headline metrics, top stall reasons and hottest SASS lines. usage: ncu_summary.py rep [kernel-regex]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('=' * 100)
    print(r[hdr.index('Kernel Name')][:110])
    for w in want:
        if w in hdr:
            print(f"  {w:85s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
    st = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                st.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    st.sort(reverse=True)
    print('  stalls/issue:', ', '.join(f"{n}={v:.2f}" for v, n in st[:7]))
if len(sys.argv) > 2:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + sys.argv[2]], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    iS, iSrc, iI = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    data = [(int(r[iS]), int(r[iI]), n, r[iSrc].strip()) for n, r in enumerate(rows[2:]) if len(r) > iS and r[iS].isdigit()]
    seen, uniq = set(), []
    for d in data:  # the page repeats per launch: keep the first
        if d[2] >= len(data) // max(1, sum(1 for x in rows if x and x[0] == 'Kernel Name')):
            break
        uniq.append(d)
    tot = sum(d[0] for d in uniq) or 1
    print('-' * 100, '\nhottest SASS (first launch), total samples', tot, 'instructions', len(uniq))
    for s_, i_, n_, src_ in sorted(uniq, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
        print(f"{s_:7d} {100 * s_ / tot:5.1f}% #{n_:5d} exec={i_:9d} {src_[:100]}")

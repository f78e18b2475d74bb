"""Key raw metrics + opcode mix of one ncu report.  usage: ncu_summary.py rep nsolves"""
import csv, collections, re, subprocess, sys
rep = sys.argv[1]; ns = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.per_cycle_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__sass_inst_executed_op_local_ld.sum',
        'smsp__sass_inst_executed_op_local_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size']
for i, h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h):
        print(f"{h},{units[i]},{vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]
iexec, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
ops = collections.Counter(); samp = collections.Counter(); tot = 0; ts = 0
for r in rows[2:]:
    try: e = int(r[iexec]); s = int(r[isamp])
    except (ValueError, IndexError): continue
    m = re.match(r'\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', r[isrc]); op = m.group(1) if m else '?'
    ops[op] += e; samp[op] += s; tot += e; ts += s
print(f"warp instructions per solve: {tot/ns:.0f}")
for op, e in ops.most_common(22): print(f"  {op:9s} {e/ns:9.0f} /solve {100*e/tot:5.1f}%  samples {100*samp[op]/max(ts,1):5.1f}%")

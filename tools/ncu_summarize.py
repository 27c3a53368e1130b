"""Condense `ncu --page raw --csv` exports (tools/evidence.sh) into the per-kernel summary kept under profiles/."""
import csv, sys, collections
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__thread_inst_executed_per_inst_executed.ratio']
def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    return [dict(name=r[idx['Kernel Name']], id=r[idx['ID']], **{k: (r[idx[k]], units[idx[k]]) for k in KEYS if k in idx}) for r in rows[2:]]
def show(r, out):
    out.write("%s  (launch %s)\n" % (r['name'][:110], r['id']))
    for k in KEYS:
        if k in r: out.write("    %-86s %s %s\n" % (k, r[k][0], r[k][1]))
def dur(r):
    v, u = r['gpu__time_duration.sum']; v = float(v.replace(',', ''))
    return v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(u, 1.0)
if __name__ == "__main__":
    out = sys.stdout
    for path, pick in zip(sys.argv[1::2], sys.argv[2::2]):
        rs = load(path)
        out.write("==== %s: %d launches captured; the %s longest\n" % (path, len(rs), pick))
        for r in sorted(rs, key=dur, reverse=True)[:int(pick)]: show(r, out)

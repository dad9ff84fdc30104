"""Summarise an .ncu-rep (raw page + source page hot lines) into text for profiles/."""
import csv, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'smsp__cycles_active.avg', 'sm__inst_executed_pipe_lsu.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'sm__maximum_warps_per_active_cycle_pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']


def main(rep, top=25):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    ki = h.index('Kernel Name')
    for r in rows[2:]:
        print('== kernel', r[ki][:60])
        for i, n in enumerate(h):
            if n in WANT:
                print(f'   {n} [{rows[1][i]}] = {r[i]}')
        stalls = [(float(r[i] or 0), n) for i, n in enumerate(h) if n.startswith('smsp__average_warps_issue_stalled') and n.endswith('_per_issue_active.ratio')]
        for v, n in sorted(stalls, reverse=True)[:6]:
            print(f'   stall {n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")}: {v:.2f}')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = None; agg = {}; cur = None; fn = None
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
        if len(r) >= 2 and r[0] == 'Function Name': fn = r[1][:40]; continue
        if len(r) > 5 and r[0] == 'Line No': hdr = r; continue
        if hdr and len(r) == len(hdr):
            try:
                ln = int(r[0]); ie = float(r[hdr.index('Instructions Executed')] or 0); te = float(r[hdr.index('Thread Instructions Executed')] or 0)
                ss = float(r[hdr.index('# Samples')] or 0)
            except Exception:
                continue
            a = agg.setdefault((fn, cur, ln, r[1][:100]), [0, 0, 0]); a[0] += ie; a[1] += te; a[2] += ss
    for f in sorted(set(k[0] for k in agg)):
        sub = {k: v for k, v in agg.items() if k[0] == f}
        tot = sum(v[0] for v in sub.values()); tots = sum(v[2] for v in sub.values())
        print(f'== hot lines of {f}: total warp inst {tot:.3g}, samples {tots:.0f}')
        for k, v in sorted(sub.items(), key=lambda kv: -kv[1][2])[:top]:
            print(f'  samples {v[2] / max(tots, 1) * 100:5.1f}%  inst {v[0] / max(tot, 1) * 100:5.1f}%  lanes {v[1] / max(v[0], 1):5.1f}  {k[1]}:{k[2]}  {k[3]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)

"""Key metrics per captured kernel from `ncu -i X.ncu-rep --page raw --csv`:  python tools/ncu_raw_summary.py file.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
WANT = [('gpu__time_duration.sum', 'duration'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__cluster_size', 'cluster'), ('launch__registers_per_thread', 'regs'),
        ('launch__shared_mem_per_block_dynamic', 'dyn smem'), ('dram__bytes_read.sum', 'dram read'),
        ('dram__bytes_write.sum', 'dram write'), ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
        ('lts__t_bytes.sum', 'L2 bytes'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
        ('sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'tensor pipe active cycles (avg over SMs)'),
        ('sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active', 'TMEM pipe %'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem wavefronts'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem wavefronts % of peak'),
        ('sm__cycles_elapsed.max', 'cycles'), ('smsp__inst_executed.sum', 'warp instructions')]
for r in data:
    print('==', r[ix['Kernel Name']][:70])
    for key, label in WANT:
        if key in ix:
            print(f'   {label:44s} {r[ix[key]]} {units[ix[key]]}')
    # issue-stall breakdown (warps per issue-active cycle), largest first
    st = [(h, r[ix[h]]) for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    def _f(x):
        try:
            return float(x.replace(',', ''))
        except Exception:
            return 0.0
    for h, v in sorted(st, key=lambda kv: -_f(kv[1]))[:7]:
        print(f"   stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:38s} {v}")
    for key in ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed',
                'smsp__inst_executed.avg.per_cycle_active', 'l1tex__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
                'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
                'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):
        if key in ix:
            print(f'   {key:60s} {r[ix[key]]} {units[ix[key]]}')

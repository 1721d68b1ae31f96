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

"""Print selected metrics of an ncu report's raw page: python tools/ncu_keys.py report.ncu-rep [substr ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
keys = sys.argv[2:] or ['gpu__time_duration.sum', 'sm__pipe_tensor', 'l1tex__data_pipe', 'l1tex__t_sector_hit_rate', 'smsp__inst_executed.sum',
                        'smsp__issue_active.avg.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum ', 'l1tex__data_bank',
                        'launch__registers', 'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'sm__cycles_elapsed.avg ', 'l1tex__f_wavefronts',
                        'smsp__inst_executed_op_shared', 'shared_op', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld']
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('==', r[4][:80])
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in keys):
            print(f'  {h} [{u}] = {v}')

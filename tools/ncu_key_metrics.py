import csv,sys,subprocess,io
keys=['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
'smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active']
for f in sys.argv[1:]:
    out=subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(io.StringIO(out)))
    h=rows[0]; u=rows[1]; r=rows[-1]
    print('==',f)
    d=dict(zip(h,zip(r,u)))
    for k in keys:
        if k in d: print(f'  {k:75s} {d[k][0]:>16s} {d[k][1]}')
    for k,(v,un) in d.items():
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
            try:
                if float(v)>0.15: print(f'  {k:75s} {v:>16s}')
            except: pass

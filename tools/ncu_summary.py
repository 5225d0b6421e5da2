import csv, sys, subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum','sm__cycles_elapsed.max','sm__inst_executed.avg.per_cycle_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_sectors.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sector_hit_rate.pct','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('----', d.get('Kernel Name')[:60])
    for k in keys:
        if k in d: print('  ',k, d[k], units[hdr.index(k)])
    st={k:d[k] for k in hdr if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')}
    for k,v in sorted(st.items(), key=lambda kv:-float(kv[1]))[:8]: print('   stall',k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v)

import csv,io,subprocess,sys
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out))); hdr=rows[0]; units=rows[1]; data=rows[2:]
pats=sys.argv[2:] or ['gpu__time_duration.sum','dram__bytes_read.sum ','dram__bytes_write.sum ','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__throughput.avg.pct','sm__warps_active.avg.pct','launch__registers_per_thread','smsp__issue_active.avg.pct','smsp__inst_executed.sum ','smsp__thread_inst_executed_per_inst_executed.ratio','issue_stalled','launch__occupancy','l1tex__data_bank_conflicts_pipe_lsu_mem_shared','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ','smsp__inst_executed_pipe_fp64','sm__inst_executed_pipe_fp64','pipe_fma','pipe_alu','pipe_xu','local']
for i,h in enumerate(hdr):
    hh=h+' '
    if any(p in hh for p in pats):
        print(h, '|', units[i], '|', [r[i] for r in data])

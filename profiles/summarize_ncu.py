import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; 
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print(d.get('Kernel Name','')[:90])
    want=['gpu__time_duration.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum',
     'lts__t_bytes.sum','l1tex__t_bytes.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__inst_executed.sum','smsp__inst_executed.sum','sm__cycles_elapsed.max','launch__occupancy_limit_registers','launch__waves_per_multiprocessor','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed']
    for k in want:
        if k in d: print('  %-70s %s'%(k,d[k]))
    st=[(k,float(d[k])) for k in hdr if 'smsp__average_warps_issue_stalled' in k and k.endswith('_per_issue_active.ratio') and d[k] not in ('','n/a')] if False else []
    st=[(k,float(d[k].replace(',',''))) for k in hdr if k.startswith('smsp__average_warp') and 'issue_stalled' in k and d[k] not in ('','n/a')]
    for k,v in sorted(st,key=lambda kv:-kv[1])[:8]: print('  STALL %-64s %.2f'%(k.replace('smsp__average_warps_issue_stalled_','').replace('smsp__average_warp_latency_issue_stalled_',''),v))

import csv,sys,subprocess
def summarize(path):
    out=subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    r=list(csv.reader(out.splitlines()))
    hdr,units,row=r[0],r[1],r[2]
    d={h:(row[i],units[i]) for i,h in enumerate(hdr)}
    keys=['Kernel Name','launch__grid_size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__warps_active.avg.per_cycle_active','smsp__warps_eligible.avg.per_cycle_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum']
    for k in keys:
        if k in d: print('  %-82s %s %s'%(k,d[k][0],d[k][1]))
    st=[(float(v[0]),h) for h,v in d.items() if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
    st.sort(reverse=True)
    print('  stalls:',', '.join('%s=%.2f'%(h.split('issue_stalled_')[1].split('_per_issue')[0],v) for v,h in st[:7]))
for p in sys.argv[1:]:
    print(p); summarize(p)

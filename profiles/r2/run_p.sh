mkdir -p gpurun_out /tmp/prof
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 300 $NCU -k 'regex:ExpandMmaBody' -s 3 -c 1 -o /tmp/prof/expandmma12 python bench.py --kernels-only --steps 1 --warmup 0 2>&1 | tail -5
python profiles/ncu_summary.py /tmp/prof/expandmma12.ncu-rep 2>&1 | tee gpurun_out/ncu_expandmma12.txt
ncu -i /tmp/prof/expandmma12.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h,u,row=r[0],r[1],r[2]
for i,k in enumerate(h):
    if any(x in k for x in ('achieved_occupancy','warps_active.avg.pct','pipe_fp64','pipe_tensor','lts__t_sectors_op_write','lts__t_bytes','l1tex__t_sectors_pipe_lsu_mem_global_op_st','sm__throughput','dram__throughput','lts__throughput','l1tex__throughput','cycles_elapsed.avg ','sm__cycles_active.avg','smsp__cycles_active.avg','launch__waves','theoretical_occupancy')): print(k, row[i], u[i])
"

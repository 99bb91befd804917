# Round-2 evidence run (1 GPU): tests, smoke, ncu captures of the kernel types of every configuration,
# DRAM traffic table, launch list, bench (both arms), breakdown of the other paths, SASS excerpt.
mkdir -p gpurun_out /tmp/prof
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/pytest_gpu_r2.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke\|engine"
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
# cap NAME SCRIPT... : one `ncu --set full` capture of the launch selected by the remaining ncu options
cap() { name=$1; script=$2; shift 2; timeout 300 $NCU "$@" -o /tmp/prof/$name $script > /dev/null 2>&1; python profiles/ncu_summary.py /tmp/prof/$name.ncu-rep > gpurun_out/ncu_r2_$name.txt 2>&1; }
C2="python bench.py --kernels-only --steps 1 --warmup 0"
cap expandmma12 "$C2" -k 'regex:ExpandMmaBody.*int.12' -s 1 -c 1
cap expandmma16 "$C2" -k 'regex:ExpandMmaBody.*int.16' -s 1 -c 1
cap passA_dense "$C2" -k 'regex:PassABody.*int.1024.*int.0.*int.1' -s 2 -c 1
# the W-writing launches of the second kernel (the coarse transforms of the expansion rows share the name)
CWTB_EXPAND_EPS=0 cap passB1024 "$C2" -k 'regex:PassBBody.*int.1.*int.1024' -s 1 -c 1
C4="python profiles/micro/config_kernels.py 4"
cap wct_prep "$C4" -k 'regex:WctPrepBody' -s 2 -c 1
cap wct_final "$C4" -k 'regex:WctFinalBody' -s 2 -c 1
cap smooth_fwd_passB "$C4" -k 'regex:PassBBody.*int.-1.*int.1024' -s 8 -c 1
cap smooth_passA "$C4" -k 'regex:PassABody.*int.256.*int.3.*int.1>' -s 8 -c 1
cap expand_f32 "python profiles/micro/config_kernels.py 5" -k 'regex:ExpandBody.float.*int.8' -s 1 -c 1
python profiles/ncu_traffic.py /tmp/prof/expandmma12.ncu-rep /tmp/prof/expandmma16.ncu-rep /tmp/prof/passB1024.ncu-rep /tmp/prof/passA_dense.ncu-rep > /dev/null 2>&1; cp profiles/traffic.json gpurun_out/traffic.json
cp /tmp/prof/expandmma12.ncu-rep gpurun_out/prof_r2_expandmma12.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv $C2 > /dev/null 2>&1
# bench after the traffic table exists (roofline.traffic is read from it)
timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_r2.err | tail -1 > gpurun_out/bench_r2.json; cut -c1-600 gpurun_out/bench_r2.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r2_reference.json; cut -c1-300 gpurun_out/bench_r2_reference.json
timeout 600 python profiles/misc_breakdown.py 2>&1 | grep -v Warning > gpurun_out/misc_breakdown_r2.txt; tail -8 gpurun_out/misc_breakdown_r2.txt
cuobjdump -sass pycwt_b200/libcwtb200.so | grep -E "UBLKCP|SYNCS|UTMA|MEMBAR|REDG|DMMA|STG.E.EF.ENL2.256" | awk '{print $2}' | sort | uniq -c | sort -rn | head -20 > gpurun_out/sass_mnemonics_r2.txt

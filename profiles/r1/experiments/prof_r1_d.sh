mkdir -p gpurun_out /tmp/prof
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_cwt.py tests/test_gpu_xwt_wct.py -x -q 2>&1 | tail -2
run() { echo "== $*"; env "$@" timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1; }
run CWTB_GROUP=4 CWTB_L2_PERSIST=1
run CWTB_GROUP=4 CWTB_L2_PERSIST=0
run CWTB_GROUP=4 CWTB_L2_PERSIST=1 CWTB_DIRECT_MAX=10
run CWTB_GROUP=2 CWTB_L2_PERSIST=1
run CWTB_GROUP=8 CWTB_L2_PERSIST=1
run CWTB_GROUP=16 CWTB_L2_PERSIST=0
run CWTB_GROUP=32 CWTB_L2_PERSIST=0
run CWTB_GROUP=4 CWTB_L2_PERSIST=1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv,noheader
CWTB_GROUP=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_d.csv python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_d0.log 2>&1
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
cap() { name=$1; shift; CWTB_GROUP=4 timeout 300 $NCU "$@" -o /tmp/prof/$name python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_$name.log 2>&1; python profiles/ncu_summary.py /tmp/prof/$name.ncu-rep > gpurun_out/sum_$name.txt 2>&1; }
cap passB -k regex:PassBBody -s 30 -c 1
cap single256 -k 'regex:SingleBody.*int.256' -s 1 -c 1
cap single1024 -k 'regex:SingleBody.*int.1024' -s 1 -c 1
cap passA_dense -k 'regex:PassABody.*int.1024.*int.0.*int.1' -s 1 -c 1
cap passA_band128 -k 'regex:PassABody.*int.128.*int.1.*int.1' -s 1 -c 1
cap direct4 -k 'regex:DirectBody.*int.4' -s 1 -c 1
cap direct8 -k 'regex:DirectBody.*int.8' -s 1 -c 1
cp /tmp/prof/passB.ncu-rep gpurun_out/prof_d_passB.ncu-rep
cp /tmp/prof/passA_dense.ncu-rep gpurun_out/prof_d_passA_dense.ncu-rep
ls -la gpurun_out/ /tmp/prof

mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cwt.py tests/test_gpu_xwt_wct.py -x -q 2>&1 | tail -4
for G in 4 8 16 32; do echo "GROUP $G"; CWTB_GROUP=$G timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1; done
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 300 $NCU -k regex:PassBBody -s 30 -c 1 -o gpurun_out/prof_passB_tma python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_b1.log 2>&1
timeout 300 $NCU -k regex:SingleBody.*256 -s 1 -c 1 -o gpurun_out/prof_single256 python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_b2.log 2>&1
timeout 300 $NCU -k regex:SingleBody.*1024 -s 1 -c 1 -o gpurun_out/prof_single1024 python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_b3.log 2>&1
timeout 300 $NCU -k regex:PassABody.*1024,.0 -s 1 -c 1 -o gpurun_out/prof_passA_dense python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_b4.log 2>&1
timeout 300 $NCU -k regex:PassABody.*128,.1 -s 1 -c 1 -o gpurun_out/prof_passA_band128 python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_b5.log 2>&1
tail -2 gpurun_out/ncu_b2.log gpurun_out/ncu_b4.log
ls -la gpurun_out/

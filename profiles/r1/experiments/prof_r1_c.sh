mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cwt.py tests/test_gpu_xwt_wct.py -x -q 2>&1 | tail -2
for G in 4 16 32; do echo "GROUP $G"; CWTB_GROUP=$G timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_c.csv python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_c0.log 2>&1
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 300 $NCU -k regex:PassBBody -s 30 -c 1 -o gpurun_out/prof_c_passB python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_c1.log 2>&1
timeout 300 $NCU -k regex:SingleBody.*256 -s 1 -c 1 -o gpurun_out/prof_c_single256 python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_c2.log 2>&1
timeout 300 $NCU -k regex:SingleBody.*1024 -s 1 -c 1 -o gpurun_out/prof_c_single1024 python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_c3.log 2>&1
timeout 300 $NCU -k 'regex:PassABody.*1024.*int.0.*int.1' -s 1 -c 1 -o gpurun_out/prof_c_passA_dense python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_c4.log 2>&1
timeout 300 $NCU -k 'regex:PassABody.*int.128.*int.1.*int.1' -s 1 -c 1 -o gpurun_out/prof_c_passA_band128 python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_c5.log 2>&1
timeout 300 $NCU -k 'regex:PassABody.*int.4,.*int.1.*int.1' -s 1 -c 1 -o gpurun_out/prof_c_passA_band4 python bench.py --kernels-only --steps 1 --warmup 0 > gpurun_out/ncu_c6.log 2>&1
ls gpurun_out/

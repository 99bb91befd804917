mkdir -p gpurun_out
for G in 1 2 4 8 16; do echo "GROUP $G"; CWTB_GROUP=$G timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1; done
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 300 $NCU -k regex:PassBBody -s 30 -c 1 -o gpurun_out/prof_passB python bench.py --kernels-only --steps 1 --warmup 0 > /dev/null 2>&1
timeout 300 $NCU -k "regex:SingleBody<double, 256>" -s 1 -c 1 -o gpurun_out/prof_single256 python bench.py --kernels-only --steps 1 --warmup 0 > /dev/null 2>&1
timeout 300 $NCU -k "regex:SingleBody<double, 1024>" -s 1 -c 1 -o gpurun_out/prof_single1024 python bench.py --kernels-only --steps 1 --warmup 0 > /dev/null 2>&1
timeout 300 $NCU -k "regex:PassABody<double, 1024, 0" -s 1 -c 1 -o gpurun_out/prof_passA_dense python bench.py --kernels-only --steps 1 --warmup 0 > /dev/null 2>&1
timeout 300 $NCU -k "regex:PassABody<double, 128, 1" -s 1 -c 1 -o gpurun_out/prof_passA_band128 python bench.py --kernels-only --steps 1 --warmup 0 > /dev/null 2>&1
ls -la gpurun_out/

mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python profiles/micro/unpadded_timing.py 2>&1 | grep -v Warn | tail -5 | tee gpurun_out/unpadded_timing_r1.txt
timeout 300 python bench.py --kernels-only --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-200

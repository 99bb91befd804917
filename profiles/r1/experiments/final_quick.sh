mkdir -p gpurun_out
timeout 100 python -m pytest tests -x -q -m gpu 2>&1 | tail -1
timeout 100 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1.json; cut -c1-330 gpurun_out/bench_r1.json

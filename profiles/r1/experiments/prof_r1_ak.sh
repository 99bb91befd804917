mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1.json; cut -c1-900 gpurun_out/bench_r1.json
timeout 600 python profiles/other_configs.py 2>&1 | grep -v Warning | tail -40 | tee gpurun_out/other_configs_r1.txt

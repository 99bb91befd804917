mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_a.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke\|engine"
timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_a.err | tail -1 > gpurun_out/bench_a.json; cut -c1-400 gpurun_out/bench_a.json

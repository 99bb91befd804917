mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_f.txt
for r in 3 2; do echo "#### CWTB_EXPAND_MIN_R=$r"; CWTB_EXPAND_MIN_R=$r SPLIT_PARTS=full timeout 300 python profiles/micro/split_timing.py 2>&1 | grep -v "=="; done | tee gpurun_out/minr_f.txt
timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_f.err | tail -1 > gpurun_out/bench_f.json; cut -c1-300 gpurun_out/bench_f.json

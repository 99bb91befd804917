mkdir -p gpurun_out /tmp/prof
timeout 600 python -m pytest tests/test_gpu_cwt.py -x -q -m gpu 2>&1 | tail -2
run() { echo "== $*"; env "$@" timeout 300 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-2500; }
run CWTB_NONE=1
run CWTB_EXPAND_MIN_R=2
run CWTB_STREAMS=1
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 300 $NCU -k 'regex:ExpandBody.*12' -s 2 -c 1 -o /tmp/prof/expand12 python bench.py --kernels-only --steps 1 --warmup 0 > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/prof/expand12.ncu-rep > gpurun_out/ncu_r2_expand12_v2.txt 2>&1
cp /tmp/prof/expand12.ncu-rep gpurun_out/
cat gpurun_out/ncu_r2_expand12_v2.txt | head -30

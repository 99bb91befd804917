# Round-2 call 2: GPU tests with the expansion path, kernel timings with / without it, ncu of ExpandBody.
mkdir -p gpurun_out /tmp/prof
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
run() { echo "== $*"; env "$@" timeout 300 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-3000; }
run CWTB_NONE=1
run CWTB_EXPAND_EPS=0
run CWTB_EXPAND_EPS=5e-15
run CWTB_STREAMS=1
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 300 $NCU -k 'regex:ExpandBody.*14' -s 1 -c 1 -o /tmp/prof/expand14 python bench.py --kernels-only --steps 1 --warmup 0 > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/prof/expand14.ncu-rep > gpurun_out/ncu_r2_expand14_v1.txt 2>&1
cp /tmp/prof/expand14.ncu-rep gpurun_out/
cat gpurun_out/ncu_r2_expand14_v1.txt | head -40

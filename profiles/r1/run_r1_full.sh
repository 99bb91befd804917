# Round-1 evidence run (1 GPU): tests, smoke, bench (both arms), launch list, ncu summaries.
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke\|engine"
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1.json; cut -c1-400 gpurun_out/bench_r1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r1_reference.json; cut -c1-400 gpurun_out/bench_r1_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --kernels-only --steps 1 --warmup 0 > /dev/null 2>&1
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
cap() { name=$1; shift; timeout 300 $NCU "$@" -o /tmp/prof/$name python bench.py --kernels-only --steps 1 --warmup 0 > /dev/null 2>&1; python profiles/ncu_summary.py /tmp/prof/$name.ncu-rep > gpurun_out/ncu_r1_$name.txt 2>&1; }
cap passB512 -k 'regex:PassBBody.*int.1.*int.512' -s 1 -c 1
cap passB1024 -k 'regex:PassBBody.*int.1.*int.1024' -s 1 -c 1
cap passA_dense -k 'regex:PassABody.*int.1024.*int.0.*int.1' -s 1 -c 1
cap passA_band128 -k 'regex:PassABody.*int.128.*int.1.*int.1' -s 1 -c 1
cap single1024 -k 'regex:SingleBody.*int.1024' -s 1 -c 1
cap direct8 -k 'regex:DirectBody.*int.8' -s 1 -c 1
cp /tmp/prof/passB512.ncu-rep gpurun_out/prof_r1_passB512.ncu-rep
# other configs, for the record (not bench lines)
timeout 600 python profiles/other_configs.py 2>&1 | grep -v Warning | tail -40 | tee gpurun_out/other_configs_r1.txt
timeout 300 python profiles/micro/resident_breakdown.py 2>&1 | grep -v Warn | tail -8 | tee gpurun_out/resident_breakdown_r1.txt

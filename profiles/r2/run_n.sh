mkdir -p gpurun_out /tmp/prof
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
cap() { name=$1; shift; timeout 300 $NCU "$@" -o /tmp/prof/$name python bench.py --kernels-only --steps 1 --warmup 0 > /dev/null 2>&1; python profiles/ncu_summary.py /tmp/prof/$name.ncu-rep > gpurun_out/ncu_$name.txt 2>&1; }
cap passA_dense_v2 -k 'regex:PassABody.*int.1024.*int.0.*int.1' -s 2 -c 1
cp /tmp/prof/passA_dense_v2.ncu-rep gpurun_out/
cat gpurun_out/ncu_passA_dense_v2.txt

# Round-2 first GPU call: memory-system micro-benchmark, baseline bench, env sweeps of the
# two-kernel scheduling (rows per chunk x concurrent chains, fused persistent kernel).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem --format=csv,noheader
./profiles/micro/l2_bw 2>&1 | tee gpurun_out/l2_bw_r2.txt
run() { echo "== $*"; env "$@" timeout 300 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-2000; }
run CWTB_NONE=1
run CWTB_GROUP=1 CWTB_CHAINS=4
run CWTB_GROUP=2 CWTB_CHAINS=4
run CWTB_GROUP=2 CWTB_CHAINS=2
run CWTB_GROUP=3 CWTB_CHAINS=2
run CWTB_GROUP=4 CWTB_CHAINS=2
run CWTB_GROUP=4 CWTB_CHAINS=1
run CWTB_GROUP=8 CWTB_CHAINS=2
run CWTB_FUSED=1 CWTB_RING=3
run CWTB_FUSED=1 CWTB_RING=4
run CWTB_STREAMS=1

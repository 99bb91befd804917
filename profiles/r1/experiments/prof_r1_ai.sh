# K2=512 second pass for band scales + chunked skewed TMA rows: A/B against K2=1024
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
for k2 in 9 10; do
  echo "CWTB_K2_BAND=$k2"
  CWTB_K2_BAND=$k2 timeout 300 python bench.py --kernels-only --steps 20 --warmup 3 2>&1 | tail -1
done
done

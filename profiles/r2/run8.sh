mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_xwt_wct.py -x -q -m gpu 2>&1 | tail -2
for mb in 64 32 128 256; do echo "#### CWTB_ROWS_CHUNK_MB=$mb"; CWTB_ROWS_CHUNK_MB=$mb timeout 600 python profiles/misc_breakdown.py 2>&1 | grep -v Warning | grep "==\|WctFinal\|200 pairs\|resident\|PassABody<double, 256, 3\|PassBBody<double, -1" | head -24; done

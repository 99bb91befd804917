mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tail -5
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1_baseline.json; cat gpurun_out/bench_r1_baseline.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r1_reference.json; cat gpurun_out/bench_r1_reference.json | cut -c1-600
# write-only / copy bandwidth calibration with the CUDA runtime
python - <<'PY'
import ctypes, time
rt = ctypes.CDLL("libcudart.so")
n = 4 << 30
a = ctypes.c_void_p(); b = ctypes.c_void_p()
rt.cudaMalloc(ctypes.byref(a), ctypes.c_size_t(n)); rt.cudaMalloc(ctypes.byref(b), ctypes.c_size_t(n))
def t(f, reps=5):
    f(); rt.cudaDeviceSynchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    rt.cudaDeviceSynchronize(); return (time.perf_counter() - t0) / reps
tm = t(lambda: rt.cudaMemset(a, 0, ctypes.c_size_t(n)))
tc = t(lambda: rt.cudaMemcpy(b, a, ctypes.c_size_t(n), 3))
print("memset 4GiB: %.3f ms -> %.2f TB/s write-only" % (tm*1e3, n/tm/1e12))
print("memcpy d2d 4GiB: %.3f ms -> %.2f TB/s read+write" % (tc*1e3, 2*n/tc/1e12))
PY

import os, sys, time, ctypes, numpy as np
sys.path.insert(0, '.')
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
ncpu = os.cpu_count()
words = (ncpu + 63) // 64
aff = pynvml.nvmlDeviceGetCpuAffinity(h, words)
cpus = [i for i in range(ncpu) if (aff[i // 64] >> (i % 64)) & 1]
print("GPU0 local cpus:", cpus[:4], "...", cpus[-4:], len(cpus), "of", ncpu)
print(os.popen("nvidia-smi topo -m | head -4").read())
print(os.popen("lscpu | grep -i 'numa\\|model name\\|socket'").read())
from pycwt_b200 import _engine
def bw(tag):
    eng = _engine.Engine(0)
    n = 1 << 30
    arr = eng.result_array((n // 16,), np.complex128)
    d = eng.dev_alloc(n)
    P = ctypes.c_void_p
    for _ in range(2):
        eng.lib.cwtb_memcpy_d2h(eng.h, arr.ctypes.data_as(P), d, n)
    t0 = time.perf_counter()
    for _ in range(4):
        eng.lib.cwtb_memcpy_d2h(eng.h, arr.ctypes.data_as(P), d, n)
    dt = (time.perf_counter() - t0) / 4
    print(tag, "D2H 1 GiB: %.1f GB/s" % (n / dt / 1e9), "running on cpu", os.sched_getaffinity(0).__len__())
bw("default affinity")
other = [c for c in range(ncpu) if c not in cpus]
if other:
    os.sched_setaffinity(0, other); bw("remote node")
os.sched_setaffinity(0, cpus); bw("GPU-local node")

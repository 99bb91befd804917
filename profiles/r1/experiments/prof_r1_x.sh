timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 120 python bench.py --kernels-only --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-110
python - <<'PY'
import numpy as np, time, sys
sys.path.insert(0, '.')
import pycwt_b200 as pycwt
eng = pycwt.default_engine()
n0 = 2 ** 22
t = np.arange(n0) / n0
x = np.sin(2 * np.pi * (50 * t + (n0 / 8) * t ** 2))
sj = 2.0 * 2 ** (np.arange(64) / 4.0)
d = eng.dev_alloc(x.nbytes); eng.h2d(d, x)
eng.cwt_dev(d, 0, n0, 1.0, sj, 0, 6.0, 0)
ms = eng.bench_last(5)
print("N=2^22, 64 scales fp64: %.2f ms -> %.3e scale-points/s (plan %s)" % (ms, 64 * n0 / ms * 1e3, sorted(set(eng.last_plan(64)))))
PY

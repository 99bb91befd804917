python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from conftest import load_golden, relerr, golden_cwt_kwargs
import pycwt_b200 as pycwt
eng = pycwt.default_engine()
for eps in (1e-20, 1e-16, 1e-13):
    eng.set_band_eps(eps)
    errs = []
    for name in ("nino3_morlet_tutorial", "chirp4000_morlet", "chirp32k_morlet", "chirp4000_dog", "chirp4000_paul"):
        g = load_golden(name)
        cls = {"morlet": pycwt.Morlet, "paul": pycwt.Paul, "dog": pycwt.DOG}[str(g["wavelet"])]
        W = pycwt.cwt(g["x"], float(g["dt"]), wavelet=cls(int(g["param"])), **golden_cwt_kwargs(g))[0]
        errs.append(relerr(W[:, ::int(g["stride"])], g["W"]))
    print("band_eps %.0e: max rel. error vs reference fixtures %.2e" % (eps, max(errs)))
PY

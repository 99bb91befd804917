import numpy as np
from expansion_error_study import test
Np = 1 << 15
print("rows: xi_max; cols: W = 4 6 8 10 12 14 16   (beta = pi W (1 - xi_max))")
for num, den in ((1,4),(7,32),(3,16),(5,32),(1,8),(3,32),(1,16),(1,32)):
    xi = num/den
    Nc = 1 << 10
    Kb = int(round(2*xi*Nc))
    row = []
    for W in (4,6,8,10,12,14,16):
        e = max(test(Np, Nc, Kb, W, np.pi*W*(1-xi)) for _ in range(2))
        row.append("%.1e" % e)
    print("%5.3f" % xi, " ".join(row))

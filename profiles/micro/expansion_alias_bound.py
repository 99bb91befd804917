import numpy as np
from expansion_error_study import kb_hat
def bound(xi_b, W):
    beta = np.pi * W * (1 - xi_b)
    xi = np.linspace(0, xi_b, 65)
    num = 0
    for l in (1, 2, 3, 4):
        num = num + np.abs(kb_hat(xi + l, W, beta)) + np.abs(kb_hat(xi - l, W, beta))
    return (num / np.abs(kb_hat(xi, W, beta))).max()
print("bound: rows xi_b; cols W=4..16")
for xi in (0.25, 7/32, 3/16, 5/32, 1/8, 3/32, 1/16, 1/32):
    print("%5.3f" % xi, " ".join("%.1e" % bound(xi, W) for W in (4,6,8,10,12,14,16)))

"""Numerical model of the band-limited expansion path: deconvolved band spectrum ->
coarse inverse FFT -> polyphase Kaiser-Bessel interpolation -> re-modulation, against the exact
full-length inverse FFT.  Used to choose tap counts before any CUDA was written."""
import numpy as np
from scipy.special import i0
rng = np.random.default_rng(0)

def kb_weights(R, W, beta):
    # h[rho][t] = phi(rho/R - (t - W/2 + 1)),  phi(x) = I0(beta sqrt(1-(2x/W)^2))/I0(beta), |x|<=W/2
    rho = np.arange(R)[:, None] / R
    t = np.arange(W)[None, :] - (W // 2 - 1)
    x = rho - t
    arg = 1 - (2 * x / W) ** 2
    return np.where(arg >= 0, i0(beta * np.sqrt(np.maximum(arg, 0))) / i0(beta), 0.0)

def kb_hat(xi, W, beta):
    # FT of the truncated KB window: W/I0(beta) * sinh(sqrt(beta^2-(pi W xi)^2))/sqrt(...)
    z = beta ** 2 - (np.pi * W * xi) ** 2
    s = np.sqrt(np.abs(z))
    val = np.where(z > 0, np.sinh(s) / np.where(s == 0, 1, s), np.sin(s) / np.where(s == 0, 1, s))
    return W / i0(beta) * val

def test(Np, Nc, Kb, W, beta, shape="flat"):
    R = Np // Nc
    kc = 12345 % Np
    k = kc + np.arange(-(Kb // 2), Kb - Kb // 2)
    B = rng.standard_normal(Kb) + 1j * rng.standard_normal(Kb)
    if shape == "gauss":
        B *= np.exp(-0.5 * ((np.arange(Kb) - Kb / 2) / (Kb / 17.16)) ** 2)   # edges at 8.58 sigma
    full = np.zeros(Np, complex); full[k % Np] = B
    exact = np.fft.ifft(full) * Np
    xi = (k - kc) / Nc
    coarse_spec = np.zeros(Nc, complex)
    coarse_spec[(k - kc) % Nc] = B / kb_hat(xi, W, beta)
    c = np.fft.ifft(coarse_spec) * Nc
    h = kb_weights(R, W, beta)
    n = np.arange(Np); m = n // R; rho = n % R
    acc = np.zeros(Np, complex)
    for t in range(W):
        acc += c[(m + t - (W // 2 - 1)) % Nc] * h[rho, t]
    out = acc * np.exp(2j * np.pi * ((kc * n) % Np) / Np)
    return np.abs(out - exact).max() / np.abs(exact).max()

if __name__ == "__main__":
    Np = 1 << 14
    for sigma_name, Nc, Kb in (("xi_max=1/4", 1 << 10, 512), ("xi_max=1/8", 1 << 10, 256), ("xi_max=3/16", 1<<10, 384)):
        for W in (8, 10, 12, 14, 16):
            best = None
            for bf in np.arange(1.8, 3.2, 0.05):
                e = test(Np, Nc, Kb, W, bf * W)
                if best is None or e < best[0]: best = (e, bf)
            eg = test(Np, Nc, Kb, W, best[1] * W, "gauss")
            print(sigma_name, "W", W, "best beta/W %.2f" % best[1], "err flat %.2e" % best[0], "gauss-shaped %.2e" % eg)

"""Known-answer vectors for the D-SSIM term of the trainer loss (reference LossFunction.py:4,31 calls the third-party
``pytorch_msssim.ssim``, which is NOT installed in this image, so parity against that package stays UNPINNED; what this file
pins is the published SSIM definition itself -- Wang, Bovik, Sheikh, Simoncelli 2004, eq. 13 with the package's documented
defaults: 11x11 Gaussian window sigma 1.5 normalised to 1, VALID windows, K1 = 0.01, K2 = 0.03, data range 1, biased window
statistics, mean over the map and the channels).

Two kinds of vectors:
  * ANALYTIC (hand-derived, no code involved): identical images -> 1; two constant images a, b ->
    (2ab + C1) / (a^2 + b^2 + C1); a constant image against any image whose every window has the same mean ...
  * COMPUTED by an independent restatement: float64 numpy, explicit 2-D window, one Python loop per output pixel -- no
    convolution routine, no separable filtering, no torch (the code paths under test use separable float32 filters).
Run:  python tests/golden/make_ssim_known_answers.py   -> tests/golden/ssim_known_answers.npz
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
C1, C2 = 0.01 ** 2, 0.03 ** 2


def window():
    x = np.arange(11, dtype=np.float64) - 5
    g = np.exp(-x ** 2 / (2 * 1.5 ** 2))
    g /= g.sum()
    return np.outer(g, g)


def ssim_textbook(x, y):
    """x, y: (3, H, W) float64.  Mean SSIM over all VALID 11x11 windows and the three channels."""
    w = window()
    C, H, W = x.shape
    total, n = 0.0, 0
    for c in range(C):
        for i in range(H - 10):
            for j in range(W - 10):
                px, py = x[c, i:i + 11, j:j + 11], y[c, i:i + 11, j:j + 11]
                mx, my = (w * px).sum(), (w * py).sum()
                vx, vy = (w * px * px).sum() - mx * mx, (w * py * py).sum() - my * my
                cxy = (w * px * py).sum() - mx * my
                total += ((2 * mx * my + C1) * (2 * cxy + C2)) / ((mx * mx + my * my + C1) * (vx + vy + C2))
                n += 1
    return total / n


def main():
    rng = np.random.default_rng(20240923)
    H, W = 24, 28
    out = {}
    cases = {}
    noise = rng.random((3, H, W))
    yy, xx = np.mgrid[0:H, 0:W]
    ramp = np.stack([xx / (W - 1), yy / (H - 1), (xx + yy) / (H + W - 2)]).astype(np.float64)
    cases["noise_vs_noise2"] = (noise, rng.random((3, H, W)))
    cases["ramp_vs_shifted_ramp"] = (ramp, np.clip(ramp + 0.1, 0, 1))
    cases["ramp_vs_noisy_ramp"] = (ramp, np.clip(ramp + 0.05 * rng.standard_normal((3, H, W)), 0, 1))
    cases["noise_vs_blurred"] = (noise, 0.25 * (noise + np.roll(noise, 1, 1) + np.roll(noise, 1, 2) + np.roll(noise, (1, 1), (1, 2))))
    for name, (x, y) in cases.items():
        x32, y32 = x.astype(np.float32), y.astype(np.float32)  # the inputs the float32 code paths will see
        out[name + "/x"], out[name + "/y"] = x32, y32
        out[name + "/ssim"] = np.float64(ssim_textbook(x32.astype(np.float64), y32.astype(np.float64)))
    # analytic cases (values derived by hand, see the module docstring)
    for k, (a, b) in enumerate([(0.5, 0.25), (0.9, 0.1), (0.0, 0.0), (1.0, 0.0)]):
        out[f"constant_{k}/a"], out[f"constant_{k}/b"] = np.float64(a), np.float64(b)
        out[f"constant_{k}/ssim"] = np.float64((2 * a * b + C1) / (a * a + b * b + C1))
    np.savez_compressed(os.path.join(HERE, "ssim_known_answers.npz"), **out)
    for k, v in out.items():
        if k.endswith("/ssim"):
            print(k, float(v))


if __name__ == "__main__":
    main()

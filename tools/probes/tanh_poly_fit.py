"""CPU study for the policy kernel: how cheap can tanh be on the FMA pipe?
Fits odd polynomials x * P(x^2) on |x| <= c (output clamped to +-1 outside) by Chebyshev-weighted least squares and
reports the maximum absolute error in float32 arithmetic (Horner with FMAs), next to the instruction count.
MUFU.TANH itself is accurate to ~5e-4 relative; the policy parity tests allow 2e-3 absolute on the network outputs."""
import numpy as np


def fit(c, terms, n=20001):
    k = np.arange(n)
    x = c * np.cos(np.pi * (k + 0.5) / n)                 # Chebyshev nodes on [-c, c]
    A = np.stack([x ** (2 * j + 1) for j in range(terms)], axis=1)
    coef, *_ = np.linalg.lstsq(A, np.tanh(x), rcond=None)
    return coef


def eval_f32(coef, c, x):
    x = np.clip(x.astype(np.float32), np.float32(-c), np.float32(c))
    x2 = x * x
    p = np.float32(coef[-1])
    for a in coef[-2::-1]:
        p = p * x2 + np.float32(a)                        # one FMA per term
    return np.clip(p * x, np.float32(-1), np.float32(1))


if __name__ == "__main__":
    xs = np.linspace(-8, 8, 400001)
    for c in (3.0, 3.5, 4.0, 4.5):
        for terms in (4, 5, 6, 7):
            coef = fit(c, terms)
            err = np.abs(eval_f32(coef, c, xs).astype(np.float64) - np.tanh(xs)).max()
            instr = 2 + 1 + (terms - 1) + 1 + 2               # clamp, x^2, Horner FMAs, * x, clamp
            print(f"|x| <= {c}: {terms} terms (degree {2 * terms - 1}), {instr} FP instructions: max abs error {err:.2e}")

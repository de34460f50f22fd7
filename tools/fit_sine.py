"""Derive the polynomial coefficients used by tuun_b200/csrc/render.cu for sin(pi/2 * x),
x in [-1, 1], as x * P(x^2): Chebyshev interpolation (near-minimax) in 60-digit arithmetic.
Prints C arrays and the measured max error of the rounded (f64 / f32) coefficients."""
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def fit(ncoef):
    # g(z) = sin(pi/2 sqrt(z))/sqrt(z) on z in [0,1]; interpolate at Chebyshev nodes of degree ncoef-1
    n = ncoef
    nodes = [(mp.cos(mp.pi * (2 * k + 1) / (2 * n)) + 1) / 2 for k in range(n)]
    def g(z):
        s = mp.sqrt(z)
        return mp.sin(mp.pi / 2 * s) / s if s != 0 else mp.pi / 2
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, z in enumerate(nodes):
        for j in range(n):
            A[i, j] = z ** j
        b[i] = g(z)
    c = mp.lu_solve(A, b)
    return [c[j] for j in range(n)]


def err(coefs, dtype):
    xs = np.linspace(-1, 1, 20001)
    worst = 0
    cs = [dtype(float(c)) for c in coefs]
    for x in xs[::7]:
        xx = dtype(x)
        z = xx * xx
        p = cs[-1]
        for c in reversed(cs[:-1]):
            p = dtype(p * z + c)
        v = float(xx * p)
        worst = max(worst, abs(v - float(mp.sin(mp.pi / 2 * mp.mpf(float(xx))))))
    return worst


for n, dt, name in ((7, np.float64, "SIN_D"), (8, np.float64, "SIN_D8"), (4, np.float32, "SIN_F"), (5, np.float32, "SIN_F5")):
    c = fit(n)
    print(f"// {name}: {n} coefficients, max abs err {err(c, dt):.3e}")
    if dt is np.float64:
        print("{" + ", ".join(float(x).hex() for x in c) + "}")
        print("{" + ", ".join(repr(float(x)) for x in c) + "}")
    else:
        print("{" + ", ".join(repr(float(np.float32(float(x)))) + "f" for x in c) + "}")

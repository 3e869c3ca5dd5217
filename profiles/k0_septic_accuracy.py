"""How well ONE degree-7 polynomial per ray reproduces the PROJ-form geodetic height h(t) along the ray (K0, k_ray_layers).

CPU / NumPy only (run anywhere): emulates the construction of k0_layers.cuh::ray_layers_septic -- eight exact heights at
t = i L / 7, coefficients through the exactly inverted Vandermonde matrix applied to the differences from the ground height,
Horner evaluation -- in double precision, and compares with the exact height (oracle.geodesy.ecef2height, PROJ's `cart`
inverse) at 4001 points of every ray.  Also prints the same for the piecewise forms the first round used (cubic / 6 km).

    python profiles/k0_septic_accuracy.py      ->  profiles/r02_k0_septic_accuracy.txt
"""
import sys
from fractions import Fraction
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import geodesy  # noqa: E402

DEG = 7
xs = [Fraction(-1) + Fraction(2 * i, DEG) for i in range(DEG + 1)]
n = DEG + 1
A = [[x ** k for k in range(n)] + [Fraction(int(i == j)) for j in range(n)] for i, x in enumerate(xs)]
for i in range(n):
    p = next(r for r in range(i, n) if A[r][i] != 0)
    A[i], A[p] = A[p], A[i]
    piv = A[i][i]
    A[i] = [a / piv for a in A[i]]
    for r in range(n):
        if r != i and A[r][i] != 0:
            f = A[r][i]
            A[r] = [a - f * b for a, b in zip(A[r], A[i])]
MINV = np.array([[float(A[i][n + j]) for j in range(n)] for i in range(n)])


def ray(lat, lon, inc, hd, ht):
    g = np.array(geodesy.lla2ecef(lat, lon, ht))
    enu = geodesy.inc_hd_to_enu(np.float64(inc), np.float64(hd))
    return g, geodesy.enu2ecef(enu[0], enu[1], enu[2], lat, lon, ht)


def h_of(g, u, t):
    p = g[None, :] + t[:, None] * u[None, :]
    return geodesy.ecef2height(p[:, 0], p[:, 1], p[:, 2])


def horner(c, x):
    r = np.full_like(x, c[-1])
    for k in range(len(c) - 2, -1, -1):
        r = r * x + c[k]
    return r


def septic_error(g, u, L):
    nodes = (np.array([float(x) for x in xs]) + 1) / 2 * L
    hn = h_of(g, u, nodes)
    c = MINV @ (hn - hn[0])
    s = np.linspace(0, 1, 4001)
    return np.abs(horner(c, 2 * s - 1) + hn[0] - h_of(g, u, L * s)).max()


def cubic_error(g, u, L, span=6000.0):
    worst, t0 = 0.0, 0.0
    while t0 < L:
        f = h_of(g, u, t0 + span * np.arange(4) / 3.0)
        d1, d2, d3 = f[1] - f[0], f[2] - f[0], f[3] - f[0]
        c1, c2, c3 = 9 * d1 - 4.5 * d2 + d3, -22.5 * d1 + 18 * d2 - 4.5 * d3, 13.5 * d1 - 13.5 * d2 + 4.5 * d3
        s = np.linspace(0, 1, 201)
        worst = max(worst, np.abs(f[0] + s * (c1 + s * (c2 + s * c3)) - h_of(g, u, t0 + span * s)).max())
        t0 += span
    return worst


lines = ['incidence  L_km   septic(whole ray)  cubic(6 km spans)   [max |h_poly - h_exact| in m over 4 latitudes x 3 headings x 2 heights]']
for inc in (0, 30, 45, 60, 70, 75, 80):
    ws = wc = 0.0
    for lat in (0, 34, 60, 80):
        for hd in (-168, 12, 90):
            for ht in (0.0, 3000.0):
                g, u = ray(lat, -118.0, inc, hd, ht)
                L = 1.05 * (82e3 - ht) / np.cos(np.radians(inc)) + 100.0
                ws = max(ws, septic_error(g, u, L))
                wc = max(wc, cubic_error(g, u, L))
    lines.append(f'{inc:9d} {1.05 * 82 / np.cos(np.radians(inc)):6.0f}   {ws:.2e}           {wc:.2e}')
out = '\n'.join(lines)
print(out)
(Path(__file__).resolve().parent / 'r02_k0_septic_accuracy.txt').write_text(out + '\n')

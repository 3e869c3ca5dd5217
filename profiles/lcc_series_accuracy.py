"""Accuracy of the local-series spherical Lambert forward used for the span nodes of Lambert (HRRR) cubes (fastpath.cuh,
node_eval<true>): rho = rho_g exp(-n (psi - psi_g)) with psi - psi_g from the 8-term Taylor series of the integral of sec about the
ray's ground point, against the direct PROJ formula in extended precision (numpy longdouble), over the whole small-angle window
(|dphi|, |dlam| <= 0.02 rad) and the latitudes of the HRRR domain.  CPU only.

    python profiles/lcc_series_accuracy.py   ->  profiles/r02_lcc_series_accuracy.txt
"""
import math
import sys
from pathlib import Path

import numpy as np
from numpy.polynomial import polynomial as P

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import geodesy  # noqa: E402

ld = np.longdouble
L = geodesy.LambertConformalSphere()
# psi^(k) = sec(phi) p_k(tan phi): p_1 = 1, p_{k+1} = t p_k + (1 + t^2) p_k'
p = [None, np.array([1.0])]
for k in range(1, 9):
    p.append(P.polyadd(P.polymul([0, 1], p[k]), P.polymul([1, 0, 1], P.polyder(p[k]))))


def direct(lat_deg, lon_deg):
    phi, lam = ld(lat_deg) * ld(np.pi) / ld(180), ld(lon_deg) * ld(np.pi) / ld(180) - ld(L.lam0)
    rho = ld(L.c) * np.tan(ld(np.pi) / 4 + phi / 2) ** (-ld(L.n))
    return ld(L.R) * rho * np.sin(lam * ld(L.n)), ld(L.R) * (ld(L.rho0) - rho * np.cos(lam * ld(L.n)))


def series(lat0, lon0, dphi, dlam, nterms):
    phi0 = math.radians(lat0)
    s, t = 1 / math.cos(phi0), math.tan(phi0)
    rho_g = L.c * (math.cos(phi0) / (1 + math.sin(phi0))) ** L.n
    th0 = (math.radians(lon0) - L.lam0) * L.n
    acc = 0.0
    for k in range(nterms, 0, -1):
        acc = acc * dphi + float(np.polyval(p[k][::-1], t)) / math.factorial(k)
    rho = rho_g * math.exp(-L.n * (s * dphi) * acc)
    th = th0 + L.n * dlam
    return L.R * rho * math.sin(th), L.R * (L.rho0 - rho * math.cos(th))


rng = np.random.default_rng(0)
lines = ['terms  window_rad  worst |dx|,|dy| in m (20000 random ground points 21..53 N, offsets uniform in the window)']
for nterms in (6, 8):
    for win in (0.008, 0.02):
        worst = 0.0
        for _ in range(20000):
            lat0, lon0 = rng.uniform(21, 53), rng.uniform(-135, -60)
            dphi, dlam = rng.uniform(-win, win), rng.uniform(-win, win)
            xf, yf = series(lat0, lon0, dphi, dlam, nterms)
            xd, yd = direct(lat0 + math.degrees(dphi), lon0 + math.degrees(dlam))
            worst = max(worst, abs(float(xd) - xf), abs(float(yd) - yf))
        lines.append(f'{nterms:5d}  {win:9.3f}  {worst:.2e}')
out = '\n'.join(lines)
print(out)
(Path(__file__).resolve().parent / 'r02_lcc_series_accuracy.txt').write_text(out + '\n')

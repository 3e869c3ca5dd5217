"""How well ONE degree-7 polynomial in the level height reproduces the layer tops the reference stores (K0, k_ray_layers).

The reference's layer top is the third iterate of its fixed-slope Newton scheme started at t = z (losreader.py:720-733 with
factor = the first layer's cos factor), T(z) = g_z(g_z(g_z(z))), g_z(t) = t + (z - h(t)) / factor.  CPU / NumPy only: T is
evaluated with the exact (PROJ-form) height at every level of the reference's 145-node table (read from the committed ERA-5
fixture) and compared with the interpolant through its values at 8 (9) Chebyshev nodes of [top of layer 1, top of the last layer].

    python profiles/k0_tfit_accuracy.py      ->  profiles/r02_k0_tfit_accuracy.txt
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import geodesy  # noqa: E402


def ray(lat, lon, inc, hd, ht):
    g = np.array(geodesy.lla2ecef(lat, lon, ht))
    enu = geodesy.inc_hd_to_enu(np.float64(inc), np.float64(hd))
    return g, np.array(geodesy.enu2ecef(enu[0], enu[1], enu[2], lat, lon, ht))


def height(g, u, t):
    p = g[None, :] + np.atleast_1d(t)[:, None] * u[None, :]
    return geodesy.ecef2height(p[:, 0], p[:, 1], p[:, 2])


def iterates(g, u, z, rfactor, iters=3):
    z = np.atleast_1d(np.asarray(z, float))
    t = z.copy()
    for _ in range(iters):
        t = t + (z - height(g, u, t)) * rfactor
    return t


def main():
    zs = np.load(ROOT / 'tests' / 'golden' / 'era5_slant_ref.npz')['z']
    lines = [f'145-node table of the reference ERA-5 fixture: {zs[0]:.0f} .. {zs[-1]:.2f} m; worst |fit - T| over all layer tops [m]',
             'incidence  ' + '  '.join(f'deg{d} {kind:>5s}' for d in (7, 8) for kind in ('cheb', 'equi'))]
    for inc in (0, 20, 30, 45, 60, 70):
        row = []
        for deg in (7, 8):
            for equi in (False, True):
                worst = 0.0
                for lat in (0, 33, 60, 80):
                    for hd in (-168, 12, 90):
                        for ht in (0.0, 1000.0):
                            g, u = ray(lat, -117.0, inc, hd, ht)
                            lev = zs[zs > ht + 1].copy()
                            lev[-1] -= 0.01
                            tlo, thi = iterates(g, u, [ht], 1.0, 10)[0], iterates(g, u, [lev[0]], 1.0, 10)[0]
                            rc = abs(thi - tlo) / (lev[0] - ht)
                            want = iterates(g, u, lev[1:], rc)
                            j = np.arange(deg + 1)
                            xn = (-1 + 2 * j / deg) if equi else np.cos(np.pi * (2 * j + 1) / (2 * (deg + 1)))
                            zn = lev[1] + (xn + 1) / 2 * (lev[-1] - lev[1])
                            c = np.polynomial.chebyshev.chebfit(xn, iterates(g, u, zn, rc), deg)
                            xk = 2 * (lev[1:] - lev[1]) / (lev[-1] - lev[1]) - 1
                            worst = max(worst, np.abs(np.polynomial.chebyshev.chebval(xk, c) - want).max())
                row.append(worst)
        lines.append(f'{inc:9d}  ' + '  '.join(f'{w:10.2e}' for w in row))
    out = '\n'.join(lines)
    print(out)
    (ROOT / 'profiles' / 'r02_k0_tfit_accuracy.txt').write_text(out + '\n')


if __name__ == '__main__':
    main()

// Hot-loop geometry + sampler of the ray tracer in the *meridian frame of the ray's ground point* (sm_100a).
//
// Reference arithmetic this stands for: tools/RAiDER/delay.py:292-298 (sub-step points, ECEF -> model CRS through PROJ) and
// scipy's trilinear evaluation (delay.py:319) for every sample; tools/RAiDER/losreader.py:706-733 for the Newton heights.
//
// A ray is g + t u.  With e_r = (cos lon0, sin lon0, 0), e_e = (-sin lon0, cos lon0, 0), e_z the unit vectors of the ground
// point's meridian plane, every point of the ray has coordinates
//     A = A0 + t uA   (horizontal distance from the polar axis, along the ground point's meridian plane)
//     B =      t uB   (east of that plane)
//     Z = Z0 + t uZ
// which is a rotation of ECEF about the polar axis: lengths are unchanged, p = sqrt(A^2 + B^2), sin(lon - lon0) = B / p, and the
// longitude of the ground point itself never has to be turned into sin/cos.  Latitude and height follow PROJ's `cart` inverse
// (Bowring 1976, single step; height p / cos(phi) - N), and the latitude / longitude are taken as small differences from the ground
// point:  sin(phi - phi0) = (y_phi cos phi0 - x_phi sin phi0) / |(x_phi, y_phi)|,  arcsine by a 5-term odd series that is
// exact to < 1 ulp for |sin| <= 0.02 (1.15 degrees, i.e. 127 km of horizontal travel).  Every reciprocal square root is one
// MUFU.RSQ64H seed (2^-20, measured: profiles/micro/seed_accuracy.cu) and one cubic (Halley-type) correction -> 2.7e-16.
//
// Rays that leave the small-angle window, start within ~9 degrees of a pole, leave the cube (horizontally, below the first or
// above the last z node) or touch the last node of a horizontal axis exactly are *flagged*, not approximated: the caller
// re-integrates them with the PROJ-form code path (k_ray_integrate in list mode), which also owns every NaN rule.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "geodesy.cuh"
#include "sampler.cuh"

namespace rdr {

// y ~ 1/sqrt(a): seed e0 ~ 2^-20, one cubic step y (1 + e/2 + 3 e^2 / 8), e = 1 - a y^2  ->  ~5/16 e0^3
__device__ __forceinline__ double rsqrt3(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double e = fma(-a * y, y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}

// y ~ 1/a: seed, one cubic step y (1 + e + e^2), e = 1 - a y
__device__ __forceinline__ double rcp3(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double e = fma(-a, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

constexpr double WGS84_A_DIV_B = 1.0 / (1.0 - WGS84_F);
constexpr double FAST_SIN_WINDOW = 0.02;       // bound on |sin| of the latitude / longitude difference: 1.15 deg, 127 km of travel
constexpr double FAST_MIN_AXIS_DIST = 1.0e6;   // ground points closer to the polar axis than this take the PROJ-form path

// The hot loop's fp64 constants live in constant memory: DFMA / DMUL take a c[bank][offset] operand for free, whereas a 64-bit
// immediate costs two UMOVs every time it is re-materialised (23 per sample in the first version of this kernel).
struct FastConst {
    double a_div_b, e2s_b, neg_es_a, neg_es, neg_a;  // Bowring
    double sin_window2;
    double asin_c[4];                                // 1/6, 3/40, 15/336, 105/3456
    double floor_magic;                              // 2^52 + 2^51
};
__constant__ FastConst c_fast = {WGS84_A_DIV_B, WGS84_E2S * WGS84_B, -WGS84_ES * WGS84_A, -WGS84_ES, -WGS84_A,
                                 FAST_SIN_WINDOW * FAST_SIN_WINDOW,
                                 {1.0 / 6.0, 3.0 / 40.0, 15.0 / 336.0, 105.0 / 3456.0},
                                 6755399441055744.0};

// asin(s) for |s| <= FAST_SIN_WINDOW: s + s^3/6 + 3 s^5/40 + 15 s^7/336 + 105 s^9/3456 (next term 0.022 s^11: 2.3e-19 relative
// at the bound)
__device__ __forceinline__ double asin_small(double s, double s2) {
    double r = fma(s2, c_fast.asin_c[3], c_fast.asin_c[2]);
    r = fma(s2, r, c_fast.asin_c[1]);
    r = fma(s2, r, c_fast.asin_c[0]);
    return fma(s * s2, r, s);
}

struct RayFrame {
    double A0, Z0;        // ground point (B0 = 0)
    double uA, uB, uZ;    // look vector
    double lat0_deg, lon0_deg;
    double slat, clat;    // of the ground latitude
    bool fast_ok;         // far enough from the polar axis for the division-free formulas
};

// ground point + look vector of ray r in its meridian frame (lla2ecef of utilFcns.py:77-82; enu2ecef :91-121; zenith vectors of
// losreader.py:312-314).  lat / lon in degrees.
__device__ __forceinline__ void frame_setup(double lat, double lon, double ht, int los_kind, const double *__restrict__ los, int64_t r, double e,
                                            double n, double up, RayFrame &F) {
    double slat, clat;
    sincos(lat * DEG_TO_RAD, &slat, &clat);
    const double N = WGS84_A / sqrt(1.0 - WGS84_ES * slat * slat);
    F.A0 = (N + ht) * clat;
    F.Z0 = (N * (1.0 - WGS84_ES) + ht) * slat;
    F.lat0_deg = lat;
    F.lon0_deg = lon;
    F.slat = slat;
    F.clat = clat;
    F.fast_ok = F.A0 > FAST_MIN_AXIS_DIST;
    if (los_kind == 0 /* RDR_LOS_ARRAY */) {
        double slon, clon;
        sincos(lon * DEG_TO_RAD, &slon, &clon);
        const double ux = __ldg(los + 3 * r), uy = __ldg(los + 3 * r + 1);
        F.uA = fma(ux, clon, uy * slon);
        F.uB = fma(uy, clon, -ux * slon);
        F.uZ = __ldg(los + 3 * r + 2);
    } else if (los_kind == 1 /* RDR_LOS_ENU_CONST */) {
        F.uA = clat * up - slat * n;
        F.uB = e;
        F.uZ = slat * up + clat * n;
    } else {
        F.uA = clat;
        F.uB = 0.0;
        F.uZ = slat;
    }
}

// Bowring's single step in the frame: everything the height and the latitude difference need
struct FrameBowring {
    double p, rp;          // sqrt(A^2 + B^2) and its reciprocal
    double x_phi, y_phi;   // tan(phi) = y_phi / x_phi
    double q2, rq;         // x_phi^2 + y_phi^2 and 1 / |(x_phi, y_phi)|
    double sphi;
};

__device__ __forceinline__ FrameBowring frame_bowring(double A, double B, double Z) {
    FrameBowring o;
    const double p2 = fma(A, A, B * B);
    o.rp = rsqrt3(p2);
    o.p = p2 * o.rp;
    const double yt = Z * c_fast.a_div_b;  // tan(theta) = (Z a) / (p b)
    const double rn = rsqrt3(fma(yt, yt, p2));
    const double ct = o.p * rn, st = yt * rn;
    o.y_phi = fma(c_fast.e2s_b * st, st * st, Z);
    o.x_phi = fma(c_fast.neg_es_a * ct, ct * ct, o.p);
    o.q2 = fma(o.y_phi, o.y_phi, o.x_phi * o.x_phi);
    o.rq = rsqrt3(o.q2);
    o.sphi = o.y_phi * o.rq;
    return o;
}

// PROJ's height: p / cos(phi) - N(phi), with 1 / cos(phi) = |(x_phi, y_phi)| / x_phi.  (The division-free form
// p cos(phi) + Z sin(phi) - a W is *more* accurate -- it is insensitive to the 1e-12 rad truncation error of Bowring's single
// step, which PROJ's form turns into ~1e-6 m at 48 km -- but parity is with the reference's arithmetic, so PROJ's form it is.)
__device__ __forceinline__ double frame_height(const FrameBowring &o) {
    const double rw = rsqrt3(fma(c_fast.neg_es * o.sphi, o.sphi, 1.0));
    return fma(o.p * (o.q2 * o.rq), rcp3(o.x_phi), c_fast.neg_a * rw);
}

__device__ __forceinline__ double frame_height(double A, double B, double Z) { return frame_height(frame_bowring(A, B, Z)); }

// getTopOfAtmosphere (losreader.py:706-733) in the frame: pos += look * (toa - h(pos)) / factor, ITERS times; returns the
// accumulated position and the along-ray distance.  EXACT = PROJ-form height (rays near the polar axis); the frame is a
// rotation of ECEF about that axis, so the height formulas take frame coordinates as they are.
template <int ITERS, bool EXACT>
__device__ __forceinline__ void frame_top_of_atmosphere(const RayFrame &F, double toa, double rfactor, double &A, double &B, double &Z, double &t) {
    A = fma(toa, F.uA, F.A0);
    B = toa * F.uB;
    Z = fma(toa, F.uZ, F.Z0);
    t = toa;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const double h = EXACT ? ecef2height(Vec3{A, B, Z}) : frame_height(A, B, Z);
        const double d = (toa - h) * rfactor;
        A = fma(F.uA, d, A);
        B = fma(F.uB, d, B);
        Z = fma(F.uZ, d, Z);
        t += d;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// the cube as the fast integrator reads it
// ------------------------------------------------------------------------------------------------------------------
// One 128-byte record (= one cache line) per cube *cell*: the trilinear interpolant of cell (iy, ix, iz) as the coefficients of
// its multilinear polynomial, both fields interleaved,
//     f(ty, tx, tz) = (a0 + a1 tz) + tx (a2 + a3 tz) + ty ((a4 + a5 tz) + tx (a6 + a7 tz))
// a0 = f000, a1 = dz f at (y0, x0), a2 / a3 = their differences along x, a4 / a5 along y, a6 / a7 the mixed differences
// (index order y, x, z).  The fields are fp32 at rest, so every one of these differences is exact in fp64, and a sample of
// both fields is 8 LDG.128 at immediate offsets from one address + 14 DFMA (the corner-lerp form needs 20 DP operations).
// Neighbouring rays of a warp sit in the same cell and share the line.
struct LerpCell {
    double4 q0, q1, q2, q3;  // {wet a0, hydro a0, wet a1, hydro a1}, {a2, a3}, {a4, a5}, {a6, a7}
};

struct FastCube {
    const LerpCell *cells;  // [(iy * (nx-1) + ix) * nzc + iz]
    int ny, nx, nzc;        // nodes along y, x; cells along z
    // uniform horizontal axes (degrees, or metres of the model projection): cell coordinate of v is fma(v, inv_d, c0) = (v - g_first) / d
    double y_inv, y_c0, x_inv, x_c0;
    int crs_kind;           // RDR_CRS_GEOGRAPHIC or RDR_CRS_LCC_SPHERE (the latter only through the polynomial integrator)
    LccParams lcc;
};

__device__ __forceinline__ void trilinear_cell(const LerpCell *q, double ty, double tx, double tz, double &vw, double &vh) {
    const double4 q0 = ld_cell(&q->q0), q1 = ld_cell(&q->q1), q2 = ld_cell(&q->q2), q3 = ld_cell(&q->q3);
    const double w0 = fma(tz, q0.z, q0.x), h0 = fma(tz, q0.w, q0.y);
    const double w1 = fma(tz, q1.z, q1.x), h1 = fma(tz, q1.w, q1.y);
    const double w2 = fma(tz, q2.z, q2.x), h2 = fma(tz, q2.w, q2.y);
    const double w3 = fma(tz, q3.z, q3.x), h3 = fma(tz, q3.w, q3.y);
    vw = fma(ty, fma(tx, w3, w2), fma(tx, w1, w0));
    vh = fma(ty, fma(tx, h3, h2), fma(tx, h1, h0));
}

struct LayerRec {
    double z_lo, neg_zlo_inv, inv_dz;  // the layer's own cell: node below, -z_lo / dz, 1 / dz
    double h_lo, h_hi;                 // a sample with h_lo <= h < h_hi is interpolated in that cell (see LAYER_TOL)
    double step;                       // 1 / (np - 1): np.linspace(0, 1, np) = j * step (delay.py:287)
    int np, iz;
};

// A sample within LAYER_TOL of its layer's own cell is interpolated (extrapolated by < 0.1 mm) in that cell instead of the
// neighbour scipy would pick: the two piecewise-linear branches differ there by |slope change| * 1e-4 m < 1e-5 N-units, i.e.
// < 1e-14 m of delay per sample.  The first / last node of the model keep their exact rule (below / above -> NaN).
constexpr double LAYER_TOL = 1.0e-4;
// Layer quadrature (k_ray_integrate_poly): how far the two end points of a layer may lie outside the layer's own cell.  They
// are handled exactly (end corrections), the bound only has to keep the *second* sample inside: layers under quadrature are
// >= 450 m of ray with >= 3 intervals, i.e. >= 100 m between samples.
constexpr double LAYER_QUAD_TOL = 2.0;

// z nodes + reciprocal cell thicknesses in shared memory: lookup for samples whose height is not inside their layer's own cell
// (the reference's fixed-point iteration leaves the layer tops of oblique rays metres away from the nominal height)
struct ZTable {
    const double *z;    // [nz]
    const double *inv;  // [nz - 1]
    int nz;
};

__device__ __forceinline__ void z_lookup(const ZTable &T, double h, int &iz, double &tz, bool &bad) {
    const int last = T.nz - 2;
    if (!(h >= T.z[0] && h <= T.z[T.nz - 1])) {  // below / above the model or NaN: scipy gives NaN, the PROJ-form path owns that
        bad = true;
        return;
    }
    int i = iz;
    while (i > 0 && h < T.z[i]) --i;
    while (i < last && h >= T.z[i + 1]) ++i;
    iz = i;
    tz = (h - T.z[i]) * T.inv[i];
}

// cell index + fraction along a uniform axis; `bad` is raised when the coordinate is outside [first, last) (the exact last
// node included: the PROJ-form path owns that rule).  floor(u) is the low word of RD(u + 2^52 + 2^51).
__device__ __forceinline__ double cell_coord(double u, int n_nodes, int &i, bool &bad) {
    const double s = __dadd_rd(u, c_fast.floor_magic);
    const int raw = __double2loint(s);
    const double fl = s - c_fast.floor_magic;
    bad |= (unsigned)raw > (unsigned)(n_nodes - 2);
    i = min(max(raw, 0), n_nodes - 2);
    return u - fl;
}

// the same without the range flag, for callers that have established the range otherwise
__device__ __forceinline__ double cell_coord_clamped(double u, int n_nodes, int &i) {
    const double s = __dadd_rd(u, c_fast.floor_magic);
    const int raw = __double2loint(s);
    const double fl = s - c_fast.floor_magic;
    i = min(max(raw, 0), n_nodes - 2);
    return u - fl;
}

// per-ray constants of the sampler: the ground point's own cell coordinates and d(cell coordinate) / d(radian);
// for a Lambert cube: {uy0, ux0, ky, kx} = {sec phi_g, tan phi_g, rho_g, theta_g} of the ground point (node_eval<true>)
struct RayCell {
    double uy0, ux0, ky, kx;
};

// R of a ray over a spherical-Lambert cube (PROJ lcc): rho_g = c tan(pi/4 + phi_g/2)^-n = c (cos phi_g / (1 + sin phi_g))^n,
// theta_g = n (lam_g - lam_0) with the longitude difference wrapped into [-pi, pi]
__device__ __forceinline__ RayCell ray_cell_lcc(const LccParams &P, double slat, double clat, double lon_deg) {
    double lam = lon_deg * DEG_TO_RAD - P.lam0;
    if (fabs(lam) > PI) lam -= 2.0 * PI * rint(lam / (2.0 * PI));
    return RayCell{1.0 / clat, slat / clat, P.c * pow(clat / (1.0 + slat), P.n), lam * P.n};
}

// One sample of both fields at frame point (A, B, Z) of a ray: geodetic latitude / longitude / height, cell lookup, trilinear
// value in lerp form.  Raises `bad` instead of handling any edge rule.
__device__ __forceinline__ void sample_fast(const FastCube &c, const RayFrame &F, const RayCell &R, const LayerRec &L, const ZTable &T, double A,
                                            double B, double Z, bool clamp_to, double clamp_h, double &h_out, double &vw, double &vh, bool &bad) {
    const FrameBowring o = frame_bowring(A, B, Z);
    double h = frame_height(o);
    h_out = h;
    if (clamp_to) h = clamp_h;
    const double slat = fma(o.y_phi, F.clat, -o.x_phi * F.slat) * o.rq;  // sin(phi - phi0)
    const double slon = B * o.rp;                                        // sin(lam - lam0)
    const double slat2 = slat * slat, slon2 = slon * slon;
    bad |= !(slat2 <= c_fast.sin_window2) | !(slon2 <= c_fast.sin_window2);
    // cell coordinates: (lat0 + dlat * RAD_TO_DEG - first) / d = uy0 + dlat * (RAD_TO_DEG / d)
    const double uy = fma(asin_small(slat, slat2), R.ky, R.uy0);
    const double ux = fma(asin_small(slon, slon2), R.kx, R.ux0);
    int iy, ix, iz = L.iz;
    const double ty = cell_coord(uy, c.ny, iy, bad);
    const double tx = cell_coord(ux, c.nx, ix, bad);
    double tz = fma(h, L.inv_dz, L.neg_zlo_inv);
    if (!(h >= L.h_lo && h < L.h_hi)) z_lookup(T, h, iz, tz, bad);
    trilinear_cell(c.cells + ((unsigned)(iy * (c.nx - 1) + ix) * (unsigned)c.nzc + (unsigned)iz), ty, tx, tz, vw, vh);
}

// ------------------------------------------------------------------------------------------------------------------
// The ray in cube coordinates as a piecewise cubic (k_ray_integrate_poly)
// ------------------------------------------------------------------------------------------------------------------
// Along a straight ray the cube coordinates (uy, ux) and the geodetic height h are smooth functions of the along-ray distance
// t: their k-th derivative scales like |r|^(1-k) (r ~ 6.4e6 m).  Over a span T the cubic through four exact evaluations at
// t_a + {0, 1/3, 2/3, 1} T misses h by < T^4 / (30 r^3): 2e-8 m for T = 8 km (measured against the PROJ-form arithmetic for
// 30 .. 75 deg incidence at 34 .. 70 deg latitude: h 2e-8 m, lat / lon 5e-8 m -- the PROJ-form height itself carries ~3e-9 m of
// rounding noise), i.e. < 1e-11 m of delay.  The exact evaluations cost ~80 DP instructions each (Bowring + two arcsines, or
// the PROJ-form + Lambert forward for projected cubes); a polynomial sample costs 3 x 3 DFMA.
struct Cubic {
    double c0, c1, c2, c3;  // f(s) = c0 + s (c1 + s (c2 + s c3)),  s in [0, 1]
};

__device__ __forceinline__ Cubic cubic_through(double f0, double f1, double f2, double f3) {
    const double d1 = f1 - f0, d2 = f2 - f0, d3 = f3 - f0;  // Lagrange on s = 0, 1/3, 2/3, 1 in difference form
    Cubic p;
    p.c0 = f0;
    p.c1 = fma(9.0, d1, fma(-4.5, d2, d3));
    p.c2 = fma(-22.5, d1, fma(18.0, d2, -4.5 * d3));
    p.c3 = fma(13.5, d1, fma(-13.5, d2, 4.5 * d3));
    return p;
}

__device__ __forceinline__ double cubic_eval(const Cubic &p, double s) { return fma(s, fma(s, fma(s, p.c3, p.c2), p.c1), p.c0); }

struct RayNode {
    double uy, ux, h;  // cell coordinates along y / x, geodetic height
};

constexpr double NODE_MARGIN = 1.0e-3;  // cells: nodes this close to the cube's outer faces hand the ray to the PROJ-form path

// exact evaluation of the ray at along-ray distance t.  LCC = false: the meridian-frame arithmetic of sample_fast (geographic
// cube); LCC = true: PROJ-form inverse in the frame (a rotation of ECEF about the polar axis, so only the longitude needs the
// ground point's added back) followed by the spherical Lambert forward of the model CRS.
template <bool LCC>
__device__ __forceinline__ RayNode node_eval(const FastCube &c, const RayFrame &F, const RayCell &R, double t, bool &bad) {
    const double A = fma(t, F.uA, F.A0), B = t * F.uB, Z = fma(t, F.uZ, F.Z0);
    RayNode n;
    if (LCC) {
        // Spherical Lambert forward (PROJ lcc, the HRRR grid) of a point a small angle (dphi, dlam) away from the ray's ground point,
        // without pow / tan per node: with psi = ln tan(pi/4 + phi/2) the projection is rho = rho_g exp(-n (psi - psi_g)),
        // theta = theta_g + n dlam, and psi - psi_g is the Taylor series of the integral of sec: its k-th derivative is
        // sec(phi_g) p_k(tan phi_g), p_{k+1} = t p_k + (1 + t^2) p_k'.  Eight terms are exact to < 1e-8 m over the whole small-angle
        // window (0.02 rad; profiles/lcc_series_accuracy.py), i.e. 4e-12 of a 3 km cell.  R = {sec phi_g, tan phi_g, rho_g, theta_g}.
        const FrameBowring o = frame_bowring(A, B, Z);
        n.h = frame_height(o);
        const double slat = fma(o.y_phi, F.clat, -o.x_phi * F.slat) * o.rq;  // sin(phi - phi0)
        const double slon = B * o.rp;                                        // sin(lam - lam0)
        const double slat2 = slat * slat, slon2 = slon * slon;
        bad |= !(slat2 <= c_fast.sin_window2) | !(slon2 <= c_fast.sin_window2);
        const double d = asin_small(slat, slat2), dlam = asin_small(slon, slon2);
        const double t = R.ux0, t2 = t * t;
        const double c8 = t * fma(t2, fma(t2, fma(t2, 5040.0, 10920.0), 7266.0), 1385.0) * (1.0 / 40320.0);
        const double c7 = fma(t2, fma(t2, fma(t2, 720.0, 1320.0), 662.0), 61.0) * (1.0 / 5040.0);
        const double c6 = t * fma(t2, fma(t2, 120.0, 180.0), 61.0) * (1.0 / 720.0);
        const double c5 = fma(t2, fma(t2, 24.0, 28.0), 5.0) * (1.0 / 120.0);
        const double c4 = t * fma(t2, 6.0, 5.0) * (1.0 / 24.0);
        const double c3 = fma(t2, 2.0, 1.0) * (1.0 / 6.0);
        const double c2 = 0.5 * t;
        double q = fma(d, c8, c7);
        q = fma(d, q, c6);
        q = fma(d, q, c5);
        q = fma(d, q, c4);
        q = fma(d, q, c3);
        q = fma(d, q, c2);
        q = fma(d, q, 1.0);
        const double dpsi = (R.uy0 * d) * q;
        const double rho = R.ky * exp(-c.lcc.n * dpsi);
        double sn, cs;
        sincos(fma(c.lcc.n, dlam, R.kx), &sn, &cs);
        const double x = fma(c.lcc.R, rho * sn, c.lcc.x0), y = fma(c.lcc.R, c.lcc.rho0 - rho * cs, c.lcc.y0);
        n.uy = fma(y, c.y_inv, c.y_c0);
        n.ux = fma(x, c.x_inv, c.x_c0);
    } else {
        const FrameBowring o = frame_bowring(A, B, Z);
        n.h = frame_height(o);
        const double slat = fma(o.y_phi, F.clat, -o.x_phi * F.slat) * o.rq;  // sin(phi - phi0)
        const double slon = B * o.rp;                                        // sin(lam - lam0)
        const double slat2 = slat * slat, slon2 = slon * slon;
        bad |= !(slat2 <= c_fast.sin_window2) | !(slon2 <= c_fast.sin_window2);
        n.uy = fma(asin_small(slat, slat2), R.ky, R.uy0);
        n.ux = fma(asin_small(slon, slon2), R.kx, R.ux0);
    }
    // (NaN coordinates fail the comparisons and flag the ray as well)
    bad |= !(n.uy >= NODE_MARGIN && n.uy <= (double)(c.ny - 1) - NODE_MARGIN) | !(n.ux >= NODE_MARGIN && n.ux <= (double)(c.nx - 1) - NODE_MARGIN);
    return n;
}

// phase-split form of sample_cell for the software-pipelined integrator: locate the cell (no cube access), load its record,
// evaluate -- so that the record of sample i + 1 is in flight while sample i is evaluated
struct CellRef {
    const LerpCell *q;
    double ty, tx, tz, w;  // fractions inside the cell and the trapezoid weight of the sample
};

struct CellData {
    double4 q0, q1, q2, q3;
};

__device__ __forceinline__ void locate_cell(const FastCube &c, int liz, double inv_dz, double neg_zlo_inv, double h_lo, double h_hi, const ZTable &T,
                                            double uy, double ux, double h, CellRef &o, bool &bad) {
    int iy, ix, iz = liz;
    o.ty = cell_coord(uy, c.ny, iy, bad);
    o.tx = cell_coord(ux, c.nx, ix, bad);
    double tz = fma(h, inv_dz, neg_zlo_inv);
    if (!(h >= h_lo && h < h_hi)) z_lookup(T, h, iz, tz, bad);
    o.tz = tz;
    o.q = c.cells + ((unsigned)(iy * (c.nx - 1) + ix) * (unsigned)c.nzc + (unsigned)iz);
}

__device__ __forceinline__ CellData load_cell(const LerpCell *q) { return {ld_cell(&q->q0), ld_cell(&q->q1), ld_cell(&q->q2), ld_cell(&q->q3)}; }

__device__ __forceinline__ void eval_cell(const CellData &d, double ty, double tx, double tz, double &vw, double &vh) {
    const double w0 = fma(tz, d.q0.z, d.q0.x), h0 = fma(tz, d.q0.w, d.q0.y);
    const double w1 = fma(tz, d.q1.z, d.q1.x), h1 = fma(tz, d.q1.w, d.q1.y);
    const double w2 = fma(tz, d.q2.z, d.q2.x), h2 = fma(tz, d.q2.w, d.q2.y);
    const double w3 = fma(tz, d.q3.z, d.q3.x), h3 = fma(tz, d.q3.w, d.q3.y);
    vw = fma(ty, fma(tx, w3, w2), fma(tx, w1, w0));
    vh = fma(ty, fma(tx, h3, h2), fma(tx, h1, h0));
}

// one sample of both fields at cube coordinates (uy, ux, h): cell lookup + trilinear value
__device__ __forceinline__ void sample_cell(const FastCube &c, const LayerRec &L, const ZTable &T, double uy, double ux, double h, double &vw,
                                            double &vh, bool &bad) {
    int iy, ix, iz = L.iz;
    const double ty = cell_coord(uy, c.ny, iy, bad);
    const double tx = cell_coord(ux, c.nx, ix, bad);
    double tz = fma(h, L.inv_dz, L.neg_zlo_inv);
    if (!(h >= L.h_lo && h < L.h_hi)) z_lookup(T, h, iz, tz, bad);
    trilinear_cell(c.cells + ((unsigned)(iy * (c.nx - 1) + ix) * (unsigned)c.nzc + (unsigned)iz), ty, tx, tz, vw, vh);
}

}  // namespace rdr

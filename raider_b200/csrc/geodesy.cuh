// WGS-84 geodesy device functions for the ray tracer (sm_100a).
//
// The reference reaches PROJ for every one of these (tools/RAiDER/utilFcns.py:77-88 via pyproj; call sites
// tools/RAiDER/delay.py:267,295 and tools/RAiDER/losreader.py:730).  PROJ is not available on the device, so
// the kernels carry PROJ's published `cart` algorithm: closed-form forward, Bowring (1976) single-step
// inverse with normalised parametric-latitude terms, height = p/cos(phi) - N (polar branch |z| - r_geocentric).
// All arithmetic is fp64 (the 1e-6 m tier).
//
// Two implementations of the inverse live here:
//   * ecef2lla / ecef2height        -- the formulas as written in PROJ (IEEE sqrt, divide, atan, atan2); used by the
//                                      API-parity entry points (rdr_ecef2lla) and as the fallback of the fast path;
//   * ecef2height_fast / ecef2lla_fast -- the same quantities for the hot loops: every sqrt/divide pair is folded into
//                                      one MUFU-seeded reciprocal (square root) with two Newton steps (no slow-path
//                                      branches), and the two inverse tangents are taken *relative to the ray's ground
//                                      point* (|delta| < ~2 deg along a tropospheric ray), where a 7-term odd series
//                                      is exact to < 1 ulp of the full angle.  Results agree with the PROJ-form code to a
//                                      few ulp (~1e-9 m in position, ~1e-15 relative in the delays); rays that leave the
//                                      small-angle window or approach the polar axis fall back to the exact code.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace rdr {

constexpr double WGS84_A = 6378137.0;
constexpr double WGS84_F = 1.0 / 298.257223563;
constexpr double WGS84_B = WGS84_A * (1.0 - WGS84_F);
constexpr double WGS84_ES = 2.0 * WGS84_F - WGS84_F * WGS84_F;
constexpr double WGS84_E2S = WGS84_ES / (1.0 - WGS84_ES);
constexpr double WGS84_B_DIV_A_SQ = (1.0 - WGS84_F) * (1.0 - WGS84_F);
constexpr double RAD_TO_DEG = 57.295779513082321;
constexpr double DEG_TO_RAD = 0.017453292519943296;
constexpr double PI = 3.14159265358979323846;

struct Vec3 {
    double x, y, z;
};

__device__ __forceinline__ Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }

// ---------------------------------------------------------------------------------------------------------------
// branch-free reciprocal / reciprocal square root for well-scaled positive arguments (no denormals, no inf/nan care
// beyond propagation): MUFU.RCP64H / MUFU.RSQ64H seed (~2^-20) + two Newton steps -> ~1 ulp.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    e = fma(-a, y, 1.0);
    return fma(y, e, y);
}

__device__ __forceinline__ double fast_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double ha = 0.5 * a;
    double e = fma(-ha * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-ha * y, y, 0.5);
    return fma(y, e, y);
}

__device__ __forceinline__ double norm3(Vec3 a) {
    const double s = fma(a.x, a.x, fma(a.y, a.y, a.z * a.z));
    return s > 0.0 ? s * fast_rsqrt(s) : sqrt(s);  // sqrt(s) keeps 0 and NaN behaviour
}

// point on the ray at along-ray distance t: g + t*u, one fused rounding per component (pinned with explicit fma
// so that every kernel reconstructs bit-identical positions from the stored distances)
__device__ __forceinline__ Vec3 ray_point(Vec3 g, Vec3 u, double t) { return {fma(t, u.x, g.x), fma(t, u.y, g.y), fma(t, u.z, g.z)}; }

// geodetic (deg, deg, m) -> ECEF; also returns sin/cos of lat and lon for the ENU rotation
__device__ __forceinline__ Vec3 lla2ecef(double lat_deg, double lon_deg, double h, double &slat, double &clat, double &slon, double &clon) {
    sincos(lat_deg * DEG_TO_RAD, &slat, &clat);
    sincos(lon_deg * DEG_TO_RAD, &slon, &clon);
    const double N = WGS84_A / sqrt(1.0 - WGS84_ES * slat * slat);
    Vec3 r;
    r.x = (N + h) * clat * clon;
    r.y = (N + h) * clat * slon;
    r.z = (N * (1.0 - WGS84_ES) + h) * slat;
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// PROJ-form inverse (exact-division code path)
// ---------------------------------------------------------------------------------------------------------------
struct Bowring {
    double p, x_phi, y_phi, cosphi, sinphi;
};

__device__ __noinline__ Bowring bowring(Vec3 c) {
    Bowring o;
    o.p = sqrt(c.x * c.x + c.y * c.y);
    const double y_theta = c.z * WGS84_A;
    const double x_theta = o.p * WGS84_B;
    const double norm = sqrt(y_theta * y_theta + x_theta * x_theta);
    const double ct = norm == 0.0 ? 1.0 : x_theta / norm;
    const double st = norm == 0.0 ? 0.0 : y_theta / norm;
    o.y_phi = c.z + WGS84_E2S * WGS84_B * st * st * st;
    o.x_phi = o.p - WGS84_ES * WGS84_A * ct * ct * ct;
    const double norm_phi = sqrt(o.y_phi * o.y_phi + o.x_phi * o.x_phi);
    o.cosphi = norm_phi == 0.0 ? 1.0 : o.x_phi / norm_phi;
    o.sinphi = norm_phi == 0.0 ? 0.0 : o.y_phi / norm_phi;
    if (o.x_phi <= 0.0) {  // degenerate / polar axis: PROJ clamps to +-90 deg
        o.cosphi = 0.0;
        o.sinphi = c.z >= 0.0 ? 1.0 : -1.0;
    }
    return o;
}

__device__ __forceinline__ double bowring_height(const Bowring &o, double z) {
    if (o.cosphi < 1e-6) {
        const double c2 = o.cosphi * o.cosphi, s2 = o.sinphi * o.sinphi;
        const double bs2 = WGS84_B_DIV_A_SQ * s2;
        const double r = WGS84_A * sqrt((c2 + WGS84_B_DIV_A_SQ * bs2) / (c2 + bs2));
        return fabs(z) - r;
    }
    return o.p / o.cosphi - WGS84_A / sqrt(1.0 - WGS84_ES * o.sinphi * o.sinphi);
}

__device__ __forceinline__ double ecef2height(Vec3 c) {
    const Bowring o = bowring(c);
    return bowring_height(o, c.z);
}

// full inverse: lon/lat in degrees + height (utilFcns.py:84-88 -> (lon, lat, h) with always_xy)
__device__ __forceinline__ void ecef2lla(Vec3 c, double &lon_deg, double &lat_deg, double &h) {
    const Bowring o = bowring(c);
    const double phi = o.x_phi <= 0.0 ? (c.z >= 0.0 ? 0.5 * PI : -0.5 * PI) : atan(o.y_phi / o.x_phi);
    lat_deg = phi * RAD_TO_DEG;
    lon_deg = atan2(c.y, c.x) * RAD_TO_DEG;
    h = bowring_height(o, c.z);
}

// ---------------------------------------------------------------------------------------------------------------
// hot-loop inverse
// ---------------------------------------------------------------------------------------------------------------
struct BowringFast {
    double p, x_phi, y_phi, q2, rq;  // q2 = x_phi^2 + y_phi^2, rq = 1/sqrt(q2)
};

__device__ __forceinline__ BowringFast bowring_fast(Vec3 c) {
    BowringFast o;
    const double p2 = fma(c.x, c.x, c.y * c.y);
    o.p = p2 * fast_rsqrt(p2);
    const double yt = c.z * WGS84_A, xt = o.p * WGS84_B;
    const double rn = fast_rsqrt(fma(yt, yt, xt * xt));
    const double ct = xt * rn, st = yt * rn;
    o.y_phi = fma(WGS84_E2S * WGS84_B * st, st * st, c.z);
    o.x_phi = fma(-WGS84_ES * WGS84_A * ct, ct * ct, o.p);
    o.q2 = fma(o.y_phi, o.y_phi, o.x_phi * o.x_phi);
    o.rq = fast_rsqrt(o.q2);
    return o;
}

// valid when x_phi > 0 and cos(phi) >= 1e-6 (caller checks via `regular`)
__device__ __forceinline__ bool regular(const BowringFast &o) { return o.x_phi * o.rq >= 1e-6; }  // also false for NaN / p == 0

__device__ __forceinline__ double height_fast(const BowringFast &o, double rx /* 1/x_phi */) {
    const double sinphi = o.y_phi * o.rq;
    const double rw = fast_rsqrt(fma(-WGS84_ES * sinphi, sinphi, 1.0));
    // p / cos(phi) - N,  1/cos(phi) = sqrt(q2)/x_phi = q2 * rq * rx
    return fma(o.p * (o.q2 * o.rq), rx, -WGS84_A * rw);
}

__device__ __forceinline__ double ecef2height_fast(Vec3 c) {
    const BowringFast o = bowring_fast(c);
    if (!regular(o)) return ecef2height(c);
    return height_fast(o, fast_rcp(o.x_phi));
}

// atan(u) for |u| <= ATAN_SMALL: odd Taylor series through u^15 (truncation < 2e-23 relative at the bound)
constexpr double ATAN_SMALL = 0.04;
__device__ __forceinline__ double atan_small(double u) {
    const double s = u * u;
    double r = fma(s, -1.0 / 15.0, 1.0 / 13.0);
    r = fma(s, r, -1.0 / 11.0);
    r = fma(s, r, 1.0 / 9.0);
    r = fma(s, r, -1.0 / 7.0);
    r = fma(s, r, 1.0 / 5.0);
    r = fma(s, r, -1.0 / 3.0);
    return fma(u * s, r, u);
}

// per-ray reference direction: the ground point's geodetic latitude / longitude (known exactly from the inputs)
struct RayRef {
    double lat0_rad, lon0_rad;
    double slat, clat, slon, clon;
};

__device__ __forceinline__ void ecef2lla_fast(Vec3 c, const RayRef &R, double &lon_deg, double &lat_deg, double &h) {
    const BowringFast o = bowring_fast(c);
    // tan(phi - phi0) and tan(lam - lam0) by the angle-difference identity; both denominators are ~|r| > 0 near the reference
    const double nphi = fma(o.y_phi, R.clat, -o.x_phi * R.slat), dphi = fma(o.x_phi, R.clat, o.y_phi * R.slat);
    const double nlam = fma(c.y, R.clon, -c.x * R.slon), dlam = fma(c.x, R.clon, c.y * R.slon);
    const bool ok = regular(o) && fabs(nphi) <= ATAN_SMALL * dphi && fabs(nlam) <= ATAN_SMALL * dlam;
    if (!ok) {
        ecef2lla(c, lon_deg, lat_deg, h);
        return;
    }
    h = height_fast(o, fast_rcp(o.x_phi));
    lat_deg = (R.lat0_rad + atan_small(nphi * fast_rcp(dphi))) * RAD_TO_DEG;
    lon_deg = (R.lon0_rad + atan_small(nlam * fast_rcp(dlam))) * RAD_TO_DEG;
}

// two samples at once (independent dependency chains for the FP64 pipe); one shared, rarely taken fallback branch
__device__ __forceinline__ void ecef2lla_fast2(Vec3 ca, Vec3 cb, const RayRef &R, double &lona, double &lata, double &ha, double &lonb,
                                               double &latb, double &hb) {
    const BowringFast oa = bowring_fast(ca), ob = bowring_fast(cb);
    const double npa = fma(oa.y_phi, R.clat, -oa.x_phi * R.slat), dpa = fma(oa.x_phi, R.clat, oa.y_phi * R.slat);
    const double npb = fma(ob.y_phi, R.clat, -ob.x_phi * R.slat), dpb = fma(ob.x_phi, R.clat, ob.y_phi * R.slat);
    const double nla = fma(ca.y, R.clon, -ca.x * R.slon), dla = fma(ca.x, R.clon, ca.y * R.slon);
    const double nlb = fma(cb.y, R.clon, -cb.x * R.slon), dlb = fma(cb.x, R.clon, cb.y * R.slon);
    const bool ok = regular(oa) && regular(ob) && fabs(npa) <= ATAN_SMALL * dpa && fabs(nla) <= ATAN_SMALL * dla &&
                    fabs(npb) <= ATAN_SMALL * dpb && fabs(nlb) <= ATAN_SMALL * dlb;
    if (!ok) {
        ecef2lla_fast(ca, R, lona, lata, ha);
        ecef2lla_fast(cb, R, lonb, latb, hb);
        return;
    }
    ha = height_fast(oa, fast_rcp(oa.x_phi));
    hb = height_fast(ob, fast_rcp(ob.x_phi));
    lata = (R.lat0_rad + atan_small(npa * fast_rcp(dpa))) * RAD_TO_DEG;
    latb = (R.lat0_rad + atan_small(npb * fast_rcp(dpb))) * RAD_TO_DEG;
    lona = (R.lon0_rad + atan_small(nla * fast_rcp(dla))) * RAD_TO_DEG;
    lonb = (R.lon0_rad + atan_small(nlb * fast_rcp(dlb))) * RAD_TO_DEG;
}

// getTopOfAtmosphere (losreader.py:706-733): Newton-Raphson along the ray to geodetic height `toa`.
// Returns the position accumulated exactly like the reference (pos += look * delta) and the along-ray distance.
template <int ITERS>
__device__ __forceinline__ Vec3 top_of_atmosphere(Vec3 g, Vec3 u, double toa, double rfactor /* 1/factor */, double &t) {
    Vec3 pos = {fma(toa, u.x, g.x), fma(toa, u.y, g.y), fma(toa, u.z, g.z)};
    t = toa;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const double d = (toa - ecef2height_fast(pos)) * rfactor;
        pos.x = fma(u.x, d, pos.x);
        pos.y = fma(u.y, d, pos.y);
        pos.z = fma(u.z, d, pos.z);
        t += d;
    }
    return pos;
}

// enu2ecef (utilFcns.py:91-121) with the trig of the pixel already in hand
__device__ __forceinline__ Vec3 enu2ecef(double e, double n, double up, double slat, double clat, double slon, double clon) {
    const double t = clat * up - slat * n;
    const double w = slat * up + clat * n;
    return {clon * t - slon * e, slon * t + clon * e, w};
}

// Lambert conformal conic, spherical form (PROJ lcc.cpp forward, e == 0 branch); P = {n, c, rho0, lam0, R, x0, y0}
struct LccParams {
    double n, c, rho0, lam0, R, x0, y0;
};

__device__ __noinline__ double2 lcc_forward(LccParams P, double lon_deg, double lat_deg) {
    double lam = lon_deg * DEG_TO_RAD - P.lam0;
    if (fabs(lam) > PI) lam -= 2.0 * PI * rint(lam / (2.0 * PI));
    const double phi = lat_deg * DEG_TO_RAD;
    const double rho = P.c * pow(tan(0.25 * PI + 0.5 * phi), -P.n);
    double s, c;
    sincos(lam * P.n, &s, &c);
    return make_double2(P.R * (rho * s) + P.x0, P.R * (P.rho0 - rho * c) + P.y0);
}

}  // namespace rdr

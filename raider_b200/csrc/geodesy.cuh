// WGS-84 geodesy device functions for the ray tracer (sm_100a).
//
// The reference reaches PROJ for every one of these (tools/RAiDER/utilFcns.py:77-88 via pyproj; call sites
// tools/RAiDER/delay.py:267,295 and tools/RAiDER/losreader.py:730).  PROJ is not available on the device, so
// the kernel carries PROJ's published `cart` algorithm: closed-form forward, Bowring (1976) single-step
// inverse with normalised parametric-latitude terms, height = p/cos(phi) - N (polar branch |z| - r_geocentric).
// All arithmetic is fp64 (the 1e-6 m tier).
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
__device__ __forceinline__ double norm3(Vec3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

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

struct Bowring {
    double p, x_phi, y_phi, cosphi, sinphi;
};

__device__ __forceinline__ Bowring bowring(Vec3 c) {
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

// height only: everything getTopOfAtmosphere needs (losreader.py:730-731) -- no atan at all
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

// getTopOfAtmosphere (losreader.py:706-733): Newton-Raphson along the ray to geodetic height `toa`.
// Returns the position accumulated exactly like the reference (pos += look * delta) and the along-ray distance.
template <int ITERS>
__device__ __forceinline__ Vec3 top_of_atmosphere(Vec3 g, Vec3 u, double toa, double factor, double &t) {
    Vec3 pos = {g.x + toa * u.x, g.y + toa * u.y, g.z + toa * u.z};
    t = toa;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const double d = (toa - ecef2height(pos)) / factor;
        pos.x += u.x * d;
        pos.y += u.y * d;
        pos.z += u.z * d;
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
__device__ __forceinline__ void lcc_forward(const double *P, double lon_deg, double lat_deg, double &X, double &Y) {
    double lam = lon_deg * DEG_TO_RAD - P[3];
    if (fabs(lam) > PI) lam -= 2.0 * PI * rint(lam / (2.0 * PI));
    const double phi = lat_deg * DEG_TO_RAD;
    const double rho = P[1] * pow(tan(0.25 * PI + 0.5 * phi), -P[0]);
    double s, c;
    sincos(lam * P[0], &s, &c);
    X = P[4] * (rho * s) + P[5];
    Y = P[4] * (P[2] - rho * c) + P[6];
}

}  // namespace rdr

// Trilinear sampler of the staged weather-model cube (wet + hydro in one pass).
//
// Reproduces scipy.interpolate.RegularGridInterpolator(method='linear', fill_value=nan, bounds_error=False) as the
// reference configures it (tools/RAiDER/delayFcns.py:55-56): interval grid[i] <= x < grid[i+1] with the last node
// inclusive, normalised distances, 8-corner sum in scipy's vertex order with separate (unfused) multiplies and adds
// so the result is bit-identical to scipy's for identical coordinates; fp32 values promoted to fp64.
// The RAiDER.interpolate interval rules (tools/bindings/interpolate/src/interpolate.cpp:106-135) are selectable.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rdr {

struct Axis {
    const double *g;  // nodes (device), strictly ascending
    int n;
    int uniform;      // 1: i = floor((v - g0) * inv_d) is a valid first guess
    double g0, inv_d;
};

struct CubeView {
    // cells[(iy*nx + ix)*(nz-1) + iz] = {wet[iz], hydro[iz], wet[iz+1], hydro[iz+1]} of column (iy, ix): 16-byte aligned pair
    const float4 *cells;
    Axis ay, ax, az;
    int crs_kind;
    double crs[7];
};

// largest i in [0, n-2] with g[i] <= v (0 if v < g[0]): scipy find_interval_ascending with extrapolate=True
__device__ __forceinline__ int cell_scipy(const Axis &a, double v, int guess) {
    const int last = a.n - 2;
    int i;
    if (guess >= 0) {
        i = guess;
    } else if (a.uniform) {
        const double f = floor((v - a.g0) * a.inv_d);
        i = f < 0.0 ? 0 : (f > (double)last ? last : (int)f);  // NaN -> comparisons false -> (int)NaN = 0
    } else {
        int lo = 0, hi = a.n - 1;  // invariant: g[lo] <= v (or lo == 0), v < g[hi] (or hi == n-1)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (v >= __ldg(a.g + mid)) lo = mid; else hi = mid;
        }
        return lo;
    }
    i = i < 0 ? 0 : (i > last ? last : i);
    while (i > 0 && v < __ldg(a.g + i)) --i;
    while (i < last && v >= __ldg(a.g + i + 1)) ++i;
    return i;
}

// bisect_left of interpolate.h:23-38 (first i with v < g[i]); returns hi in [0, n]
__device__ __forceinline__ int bisect_left(const double *g, int n, double v) {
    int left = 0, right = n;
    while (right != left) {
        const int mid = (left + right) >> 1;
        if (v < __ldg(g + mid)) right = mid; else left = mid + 1;
    }
    return right;
}

// scipy's _evaluate_linear for ndim = 3 and two fields; vertex order (y,x,z) = 000,001,010,011,100,101,110,111
__device__ __forceinline__ void trilinear_scipy(float4 c00, float4 c01, float4 c10, float4 c11, double ty, double tx, double tz,
                                                double &vw, double &vh) {
    const double uy = 1.0 - ty, ux = 1.0 - tx, uz = 1.0 - tz;
    const double w00 = __dmul_rn(uy, ux), w01 = __dmul_rn(uy, tx), w10 = __dmul_rn(ty, ux), w11 = __dmul_rn(ty, tx);
    double w, a, b;
    w = __dmul_rn(w00, uz); a = __dmul_rn((double)c00.x, w);                 b = __dmul_rn((double)c00.y, w);
    w = __dmul_rn(w00, tz); a = __dadd_rn(a, __dmul_rn((double)c00.z, w));   b = __dadd_rn(b, __dmul_rn((double)c00.w, w));
    w = __dmul_rn(w01, uz); a = __dadd_rn(a, __dmul_rn((double)c01.x, w));   b = __dadd_rn(b, __dmul_rn((double)c01.y, w));
    w = __dmul_rn(w01, tz); a = __dadd_rn(a, __dmul_rn((double)c01.z, w));   b = __dadd_rn(b, __dmul_rn((double)c01.w, w));
    w = __dmul_rn(w10, uz); a = __dadd_rn(a, __dmul_rn((double)c10.x, w));   b = __dadd_rn(b, __dmul_rn((double)c10.y, w));
    w = __dmul_rn(w10, tz); a = __dadd_rn(a, __dmul_rn((double)c10.z, w));   b = __dadd_rn(b, __dmul_rn((double)c10.w, w));
    w = __dmul_rn(w11, uz); a = __dadd_rn(a, __dmul_rn((double)c11.x, w));   b = __dadd_rn(b, __dmul_rn((double)c11.y, w));
    w = __dmul_rn(w11, tz); a = __dadd_rn(a, __dmul_rn((double)c11.z, w));   b = __dadd_rn(b, __dmul_rn((double)c11.w, w));
    vw = a;
    vh = b;
}

// One scipy-semantics sample of both fields at cube coordinates (y, x, z).  zguess >= 0: cell index hint for z.
__device__ __forceinline__ void sample_scipy(const CubeView &c, double y, double x, double z, int zguess, double &vw, double &vh) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const Axis &ay = c.ay, &ax = c.ax, &az = c.az;
    // out of bounds (strictly outside [g0, g_last]) or NaN coordinate -> NaN  (_find_out_of_bounds + nans mask)
    const bool inb = (y >= __ldg(ay.g)) && (y <= __ldg(ay.g + ay.n - 1)) && (x >= __ldg(ax.g)) && (x <= __ldg(ax.g + ax.n - 1)) &&
                     (z >= __ldg(az.g)) && (z <= __ldg(az.g + az.n - 1));
    if (!inb) {
        vw = qnan;
        vh = qnan;
        return;
    }
    const int iy = cell_scipy(ay, y, -1), ix = cell_scipy(ax, x, -1), iz = cell_scipy(az, z, zguess);
    const double y0 = __ldg(ay.g + iy), x0 = __ldg(ax.g + ix), z0 = __ldg(az.g + iz);
    const double ty = (y - y0) / (__ldg(ay.g + iy + 1) - y0);
    const double tx = (x - x0) / (__ldg(ax.g + ix + 1) - x0);
    const double tz = (z - z0) / (__ldg(az.g + iz + 1) - z0);
    const int nzc = az.n - 1;
    const float4 *p = c.cells + ((size_t)iy * ax.n + ix) * nzc + iz;
    const float4 c00 = __ldg(p), c01 = __ldg(p + nzc), c10 = __ldg(p + (size_t)ax.n * nzc), c11 = __ldg(p + (size_t)ax.n * nzc + nzc);
    trilinear_scipy(c00, c01, c10, c11, ty, tx, tz, vw, vh);
}

// RAiDER.interpolate 3-D formula (interpolate.cpp:155-174): un-normalised distances, one divide by dx*dy*dz.
// Grid order there is (x, y, z) = our (y, x, z) axes 0,1,2.
__device__ __forceinline__ double trilinear_raider(double w000, double w001, double w010, double w011, double w100, double w101,
                                                   double w110, double w111, double d0lo, double d0hi, double d1lo, double d1hi,
                                                   double d2lo, double d2hi, double vol) {
    // d?lo = x - x0 ("dist_x0"), d?hi = x1 - x ("dist_x1")
    const double a = __dadd_rn(__dmul_rn(d2hi, w000), __dmul_rn(d2lo, w001));
    const double b = __dadd_rn(__dmul_rn(d2hi, w010), __dmul_rn(d2lo, w011));
    const double c = __dadd_rn(__dmul_rn(d2hi, w100), __dmul_rn(d2lo, w101));
    const double d = __dadd_rn(__dmul_rn(d2hi, w110), __dmul_rn(d2lo, w111));
    const double lo = __dadd_rn(__dmul_rn(d1hi, a), __dmul_rn(d1lo, b));
    const double hi = __dadd_rn(__dmul_rn(d1hi, c), __dmul_rn(d1lo, d));
    return __ddiv_rn(__dadd_rn(__dmul_rn(d0hi, lo), __dmul_rn(d0lo, hi)), vol);
}

}  // namespace rdr

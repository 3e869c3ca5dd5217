// Trilinear sampler of the staged weather-model cube (wet + hydro in one pass).
//
// Reproduces scipy.interpolate.RegularGridInterpolator(method='linear', fill_value=nan, bounds_error=False) as the
// reference configures it (tools/RAiDER/delayFcns.py:55-56): interval grid[i] <= x < grid[i+1] with the last node
// inclusive, normalised distances t = (x - g[i]) / (g[i+1] - g[i]), 8-corner sum in scipy's vertex order with separate
// (unfused) multiplies and adds so the result is bit-identical to scipy's for identical coordinates; fp32 values
// promoted to fp64.  The RAiDER.interpolate interval rules (tools/bindings/interpolate/src/interpolate.cpp:106-135) are
// selectable in the K2 entry point.
//
// Per-axis acceleration tables (built at rdr_set_cube):
//   cell[i] = {g[i], g[i+1], d = fl(g[i+1] - g[i]), inv = fl(1/d)}   one 32-byte record per interval
//   bin[b]  = interval containing the start of uniform bin b          first guess for irregular axes (model z levels)
// The interval found from the guess is always verified against the real nodes, and the division is the two-step
// Markstein refinement of n * inv (correctly rounded n / d), so none of this changes a single bit of the result.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "geodesy.cuh"

namespace rdr {

struct Axis {
    const double *g;           // nodes (device), strictly ascending
    const double4 *cell;       // n-1 interval records
    const unsigned short *bin; // nbin first-guess table
    int n, nbin;
    int uniform;               // nodes are (numerically) equally spaced: first guess = (v - g_first) * inv_d, no bin lookup
    double inv_d;              // (n - 1) / (g_last - g_first)
    double g_first, g_last;    // g[0], g[n-1]: bounds tests read these from the kernel-parameter bank
    double inv_bw;             // nbin / (g_last - g_first)
};

struct CubeView {
    // cells[(iy*nx + ix)*(nz-1) + iz] = {wet[iz], hydro[iz], wet[iz+1], hydro[iz+1]} of column (iy, ix), promoted to fp64 once
    // at staging (the promotion float -> double is exact, and it keeps 8 F2F conversions per sample off the quarter-rate XU
    // pipe): one 32-byte record feeds the z-pair of both fields
    const double4 *cells;
    Axis ay, ax, az;
    int crs_kind;
    LccParams lcc;
};

// correctly rounded n / d given inv = RN(1/d): q0 = n*inv, two residual corrections (Markstein)
__device__ __forceinline__ double div_exact(double n, double d, double inv) {
    double q = n * inv;
    double r = fma(-d, q, n);
    q = fma(r, inv, q);
    r = fma(-d, q, n);
    return fma(r, inv, q);
}

__device__ __forceinline__ double4 ld_cell(const double4 *p) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// interval of an in-bounds coordinate v: largest i in [0, n-2] with g[i] <= v, found from `guess` (or the bin table when
// guess < 0) and verified against the nodes; returns t = (v - g[i]) / (g[i+1] - g[i])
__device__ __forceinline__ double locate(const Axis &a, double v, int &i) {
    const int last = a.n - 2;
    if (i < 0) {
        if (a.uniform) {
            i = (int)((v - a.g_first) * a.inv_d);
            i = i < 0 ? 0 : (i > last ? last : i);
        } else {
            int b = (int)((v - a.g_first) * a.inv_bw);
            b = b < 0 ? 0 : (b >= a.nbin ? a.nbin - 1 : b);
            i = __ldg(a.bin + b);
        }
    }
    double4 c = ld_cell(a.cell + i);
    if (v < c.x || v >= c.y) {  // first guess off (node hit, rounding of the guess, or the inclusive last node): walk to the interval
        while (v < c.x && i > 0) c = ld_cell(a.cell + --i);
        while (v >= c.y && i < last) c = ld_cell(a.cell + ++i);
    }
    return div_exact(v - c.x, c.z, c.w);
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
__device__ __forceinline__ void trilinear_scipy(double4 c00, double4 c01, double4 c10, double4 c11, double ty, double tx, double tz,
                                                double &vw, double &vh) {
    const double uy = 1.0 - ty, ux = 1.0 - tx, uz = 1.0 - tz;
    const double w00 = __dmul_rn(uy, ux), w01 = __dmul_rn(uy, tx), w10 = __dmul_rn(ty, ux), w11 = __dmul_rn(ty, tx);
    double w, a, b;
    w = __dmul_rn(w00, uz); a = __dmul_rn(c00.x, w);                 b = __dmul_rn(c00.y, w);
    w = __dmul_rn(w00, tz); a = __dadd_rn(a, __dmul_rn(c00.z, w));   b = __dadd_rn(b, __dmul_rn(c00.w, w));
    w = __dmul_rn(w01, uz); a = __dadd_rn(a, __dmul_rn(c01.x, w));   b = __dadd_rn(b, __dmul_rn(c01.y, w));
    w = __dmul_rn(w01, tz); a = __dadd_rn(a, __dmul_rn(c01.z, w));   b = __dadd_rn(b, __dmul_rn(c01.w, w));
    w = __dmul_rn(w10, uz); a = __dadd_rn(a, __dmul_rn(c10.x, w));   b = __dadd_rn(b, __dmul_rn(c10.y, w));
    w = __dmul_rn(w10, tz); a = __dadd_rn(a, __dmul_rn(c10.z, w));   b = __dadd_rn(b, __dmul_rn(c10.w, w));
    w = __dmul_rn(w11, uz); a = __dadd_rn(a, __dmul_rn(c11.x, w));   b = __dadd_rn(b, __dmul_rn(c11.y, w));
    w = __dmul_rn(w11, tz); a = __dadd_rn(a, __dmul_rn(c11.z, w));   b = __dadd_rn(b, __dmul_rn(c11.w, w));
    vw = a;
    vh = b;
}

// One scipy-semantics sample of both fields at cube coordinates (y, x, z).  iy/ix/iz: interval hints in (negative = none),
// intervals used out -- a ray marching through the cube hands each sample's cells to the next one.
__device__ __forceinline__ void sample_scipy(const CubeView &c, double y, double x, double z, int &iy, int &ix, int &iz, double &vw,
                                             double &vh) {
    // out of bounds (strictly outside [g0, g_last]) or NaN coordinate -> NaN  (_find_out_of_bounds + nans mask)
    const bool inb = (y >= c.ay.g_first) && (y <= c.ay.g_last) && (x >= c.ax.g_first) && (x <= c.ax.g_last) && (z >= c.az.g_first) &&
                     (z <= c.az.g_last);
    if (!inb) {
        vw = vh = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    const double ty = locate(c.ay, y, iy), tx = locate(c.ax, x, ix), tz = locate(c.az, z, iz);
    const int nzc = c.az.n - 1;
    const unsigned row = (unsigned)c.ax.n * (unsigned)nzc;  // cells per y-row; the whole cube has < 2^31 cells (checked at staging)
    const double4 *p = c.cells + ((unsigned)iy * row + (unsigned)ix * (unsigned)nzc + (unsigned)iz);
    const double4 c00 = ld_cell(p), c01 = ld_cell(p + nzc), c10 = ld_cell(p + row), c11 = ld_cell(p + row + nzc);
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

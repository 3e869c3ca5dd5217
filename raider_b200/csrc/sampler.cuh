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
    // "exact-uniform" axis: every node is bit-for-bit fma(i, d, g_first) with d = g[1] - g[0] (regular grids whose origin and
    // spacing are short binary fractions: 0.25, 0.3125, 0.5, 0.625 degrees ...).  Then the nodes of an interval are recomputed
    // in registers instead of loaded, every interval has the same width d, and only inv_dx = RN(1 / d) is needed.
    int exact_uniform;
    double d, inv_dx;
};

// fp32 view of an axis for the fp32 tier of the streaming sampler (k_sample_stream_f32): per interval one 16-byte record
// {lo_hi, lo_lo, hi, inv} with g[i] = lo_hi + lo_lo to fp32 x fp32 precision (so (v - lo_hi) - lo_lo is the exact difference
// rounded once, with the exact sign), hi = RU32(g[i+1]) for the walk, inv = RN32(1 / (g[i+1] - g[i])).  first_cmp / last_cmp are RU32(g[0]) /
// RD32(g[n-1]): for an fp32 coordinate v, v >= first_cmp <=> v >= g[0] and v <= last_cmp <=> v <= g[n-1] *exactly*, so the
// out-of-bounds rule of scipy (NaN outside the closed box) is reproduced bit for bit on fp32 inputs.
struct Axis32 {
    const float4 *rec;
    const unsigned short *bin;
    int n, nbin, uniform;
    float g_first, inv_d, inv_bw, first_cmp, last_cmp;
    // exact32: every node is exactly g_first + i d in fp32 (0.25 / 0.3125 / 0.5 degree grids ...): the interval record is rebuilt
    // in registers (one FFMA) instead of loaded
    int exact32;
    float d;
};

struct CubeView {
    Axis32 fy, fx, fz;
    // cells32[(iy*nx + ix)*(nz-1) + iz] = the same z-pair record in fp32 (16 bytes): what the streaming sampler K2 loads -- its
    // limiter is L1 wavefronts (bytes delivered per lane), and 8 F2F conversions on the XU pipe are cheaper than 64 more bytes
    const float4 *cells32;
    // cells[(iy*nx + ix)*(nz-1) + iz] = {wet[iz], hydro[iz], wet[iz+1], hydro[iz+1]} of column (iy, ix), promoted to fp64 once
    // at staging (the promotion float -> double is exact, and it keeps 8 F2F conversions per sample off the quarter-rate XU
    // pipe): one 32-byte record feeds the z-pair of both fields
    const double4 *cells;
    Axis ay, ax, az;
    int crs_kind;
    LccParams lcc;
};

// correctly rounded n / d given inv = RN(1/d): q0 = n*inv, then residual corrections (Markstein).  One correction is
// already correctly rounded when inv is the correctly rounded reciprocal and q0 is within 1 ulp (Markstein's theorem);
// rdr_selftest_div counts mismatches of both forms against IEEE division on the device (tests/test_gpu_parity.py).
__device__ __forceinline__ double div_exact2(double n, double d, double inv) {
    double q = n * inv;
    double r = fma(-d, q, n);
    q = fma(r, inv, q);
    r = fma(-d, q, n);
    return fma(r, inv, q);
}
__device__ __forceinline__ double div_exact1(double n, double d, double inv) {
    const double q = n * inv;
    const double r = fma(-d, q, n);
    return fma(r, inv, q);
}
#ifndef RDR_DIV_STEPS
#define RDR_DIV_STEPS 1
#endif
__device__ __forceinline__ double div_exact(double n, double d, double inv) {
#if RDR_DIV_STEPS == 1
    return div_exact1(n, d, inv);
#else
    return div_exact2(n, d, inv);
#endif
}

__device__ __forceinline__ double4 ld_cell(const double4 *p) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// first guess of the interval of v when no hint is available
enum { GUESS_HINT = 0, GUESS_UNIFORM = 1, GUESS_BINS = 2, GUESS_EXACT_UNIFORM = 3 };

constexpr double SAMPLER_FLOOR_MAGIC = 6755399441055744.0;  // 2^52 + 2^51

// interval record {g[i], g[i+1], d, inv} of an exact-uniform axis, rebuilt in registers: floor((v - g0) / d) by directed rounding
// against 2^52 + 2^51 (its low word is the integer, the difference is the same integer as a double -- no F2I / I2F)
__device__ __forceinline__ double4 exact_uniform_record(const Axis &a, double v, int &i) {
    const double u = (v - a.g_first) * a.inv_d;
    const double s = __dadd_rd(u, SAMPLER_FLOOR_MAGIC);
    const int raw = __double2loint(s);
    i = min(max(raw, 0), a.n - 2);
    const double fi = (raw == i) ? s - SAMPLER_FLOOR_MAGIC : (double)i;  // clamped (out of bounds / NaN): rare
    const double lo = fma(fi, a.d, a.g_first);
    return make_double4(lo, lo + a.d, a.d, a.inv_dx);
}

__device__ __forceinline__ void fix_interval_exact_uniform(const Axis &a, double v, int &i, double4 &r, bool &inb) {
    if (v < r.x || v >= r.y) {
        const int last = a.n - 2;
        while (v < r.x && i > 0) {
            --i;
            r.x = fma((double)i, a.d, a.g_first);
            r.y = r.x + a.d;
        }
        while (v >= r.y && i < last) {
            ++i;
            r.x = fma((double)i, a.d, a.g_first);
            r.y = r.x + a.d;
        }
        inb &= (v >= r.x) && (v <= r.y);
    }
}

template <int MODE>
__device__ __forceinline__ int guess_interval(const Axis &a, double v, int hint) {
    if (MODE == GUESS_HINT) return hint;
    const int last = a.n - 2;
    if (MODE == GUESS_UNIFORM) {
        const int i = (int)((v - a.g_first) * a.inv_d);
        return min(max(i, 0), last);
    }
    const int b = (int)((v - a.g_first) * a.inv_bw);
    return __ldg(a.bin + min(max(b, 0), a.nbin - 1));
}

// interval of an in-bounds coordinate v: largest i in [0, n-2] with g[i] <= v, starting from the guess and verified against
// the nodes; returns t = (v - g[i]) / (g[i+1] - g[i]) and leaves the interval in i
// `inb` is cleared when v lies outside [g[0], g[n-1]]: only the walk can find that out, so the common case (guess right, v
// inside its interval) pays no bounds test at all.  A NaN coordinate fails both walk conditions, keeps `inb`, and poisons t --
// which is scipy's answer for it (NaN in, NaN out).
template <int MODE>
__device__ __forceinline__ double locate(const Axis &a, double v, int &i, bool &inb) {
    const int last = a.n - 2;
    i = guess_interval<MODE>(a, v, i);
    double4 c = ld_cell(a.cell + i);
    if (v < c.x || v >= c.y) {  // guess off (node hit, rounding of the guess, the inclusive last node) or out of bounds: walk
        while (v < c.x && i > 0) c = ld_cell(a.cell + --i);
        while (v >= c.y && i < last) c = ld_cell(a.cell + ++i);
        inb &= (v >= c.x) && (v <= c.y);
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

// One scipy-semantics sample of both fields at cube coordinates (y, x, z).  MXY / MZ: how the first guess of the y,x / z
// interval is made (GUESS_HINT: iy/ix/iz carry the previous sample's intervals in; they always carry the intervals used out).
// Out-of-bounds or NaN coordinates give NaN (_find_out_of_bounds + nans mask of scipy) without a divergent early exit: the
// lookup runs on a safe stand-in coordinate and the result is replaced at the end.
template <int MXY, int MZ>
__device__ __forceinline__ void sample_scipy(const CubeView &c, double y, double x, double z, int &iy, int &ix, int &iz, double &vw,
                                             double &vh) {
    bool inb = true;
    int jy = iy, jx = ix, jz = iz;  // (OOB / NaN coordinates cannot make the walks run away: see sample_prepare)
    const double ty = locate<MXY>(c.ay, y, jy, inb), tx = locate<MXY>(c.ax, x, jx, inb), tz = locate<MZ>(c.az, z, jz, inb);
    const int nzc = c.az.n - 1;
    const unsigned row = (unsigned)c.ax.n * (unsigned)nzc;  // cells per y-row; the whole cube has < 2^31 cells (checked at staging)
    const double4 *p = c.cells + ((unsigned)jy * row + (unsigned)jx * (unsigned)nzc + (unsigned)jz);
    const double4 c00 = ld_cell(p), c01 = ld_cell(p + nzc), c10 = ld_cell(p + row), c11 = ld_cell(p + row + nzc);
    trilinear_scipy(c00, c01, c10, c11, ty, tx, tz, vw, vh);
    if (inb) {
        iy = jy;
        ix = jx;
        iz = jz;
    } else {
        vw = vh = __longlong_as_double(0x7ff8000000000000LL);
    }
}

// ---- phase-split form of the scipy sampler for kernels that keep several points in flight per thread (K2): all guesses,
// then all interval-record loads, then the (rare) fix-ups, then all cell loads, then the arithmetic -- so that the loads of the
// points overlap instead of forming one dependent chain per point.
struct SamplePrep {
    double y, x, z;   // coordinates used for the lookup (stand-ins when out of bounds)
    int iy, ix, iz;
    bool inb;
};

template <int MXY, int MZ>
__device__ __forceinline__ void sample_prepare(const CubeView &c, double y, double x, double z, SamplePrep &s) {
    s.inb = true;  // cleared by fix_interval when a walk ends outside the grid
    // out-of-bounds / NaN coordinates need no stand-in: the guesses are clamped into the table, a coordinate below the first or
    // above the last node stops the walk of fix_interval at the end interval (and the garbage t is replaced by NaN at the end), a
    // NaN fails every comparison and poisons t by itself
    s.y = y;
    s.x = x;
    s.z = z;
    s.iy = guess_interval<MXY>(c.ay, y, s.iy);
    s.ix = guess_interval<MXY>(c.ax, x, s.ix);
    s.iz = guess_interval<MZ>(c.az, z, s.iz);
}

__device__ __forceinline__ void fix_interval(const Axis &a, double v, int &i, double4 &r, bool &inb) {
    if (v < r.x || v >= r.y) {
        const int last = a.n - 2;
        while (v < r.x && i > 0) r = ld_cell(a.cell + --i);
        while (v >= r.y && i < last) r = ld_cell(a.cell + ++i);
        inb &= (v >= r.x) && (v <= r.y);
    }
}

__device__ __forceinline__ double4 ld_cell32(const float4 *p) {
    const float4 a = __ldg(p);
    return make_double4((double)a.x, (double)a.y, (double)a.z, (double)a.w);  // exact promotions
}

template <int NPT, int MXY, int MZ>
__device__ __forceinline__ void sample_scipy_batch(const CubeView &c, const double (&y)[NPT], const double (&x)[NPT], const double (&z)[NPT],
                                                   double (&vw)[NPT], double (&vh)[NPT]) {
    SamplePrep s[NPT];
    double4 ry[NPT], rx[NPT], rz[NPT];
    constexpr bool XY_EXACT = MXY == GUESS_EXACT_UNIFORM;
#pragma unroll
    for (int p = 0; p < NPT; ++p) {
        s[p].iy = s[p].ix = s[p].iz = 0;
        if (XY_EXACT) {
            s[p].inb = true;
            s[p].y = y[p];
            s[p].x = x[p];
            s[p].z = z[p];
            ry[p] = exact_uniform_record(c.ay, y[p], s[p].iy);
            rx[p] = exact_uniform_record(c.ax, x[p], s[p].ix);
            s[p].iz = guess_interval<MZ>(c.az, z[p], 0);
        } else {
            sample_prepare<MXY, MZ>(c, y[p], x[p], z[p], s[p]);
        }
    }
#pragma unroll
    for (int p = 0; p < NPT; ++p) {
        if (!XY_EXACT) {
            ry[p] = ld_cell(c.ay.cell + s[p].iy);
            rx[p] = ld_cell(c.ax.cell + s[p].ix);
        }
        rz[p] = ld_cell(c.az.cell + s[p].iz);
    }
#pragma unroll
    for (int p = 0; p < NPT; ++p) {
        if (XY_EXACT) {
            fix_interval_exact_uniform(c.ay, s[p].y, s[p].iy, ry[p], s[p].inb);
            fix_interval_exact_uniform(c.ax, s[p].x, s[p].ix, rx[p], s[p].inb);
        } else {
            fix_interval(c.ay, s[p].y, s[p].iy, ry[p], s[p].inb);
            fix_interval(c.ax, s[p].x, s[p].ix, rx[p], s[p].inb);
        }
        fix_interval(c.az, s[p].z, s[p].iz, rz[p], s[p].inb);
    }
    const int nzc = c.az.n - 1;
    const unsigned row = (unsigned)c.ax.n * (unsigned)nzc;
    double4 c00[NPT], c01[NPT], c10[NPT], c11[NPT];
#pragma unroll
    for (int p = 0; p < NPT; ++p) {
        const float4 *q = c.cells32 + ((unsigned)s[p].iy * row + (unsigned)s[p].ix * (unsigned)nzc + (unsigned)s[p].iz);
        c00[p] = ld_cell32(q);
        c01[p] = ld_cell32(q + nzc);
        c10[p] = ld_cell32(q + row);
        c11[p] = ld_cell32(q + row + nzc);
    }
#pragma unroll
    for (int p = 0; p < NPT; ++p) {
        const double ty = div_exact(s[p].y - ry[p].x, ry[p].z, ry[p].w);
        const double tx = div_exact(s[p].x - rx[p].x, rx[p].z, rx[p].w);
        const double tz = div_exact(s[p].z - rz[p].x, rz[p].z, rz[p].w);
        trilinear_scipy(c00[p], c01[p], c10[p], c11[p], ty, tx, tz, vw[p], vh[p]);
        if (!s[p].inb) vw[p] = vh[p] = __longlong_as_double(0x7ff8000000000000LL);
    }
}

// two samples of one ray, both starting from the same interval hints (the previous sample's cells); hints leave as the
// second sample's cells when it was in bounds
__device__ __forceinline__ void sample_scipy_pair_hinted(const CubeView &c, const double (&y)[2], const double (&x)[2], const double (&z)[2],
                                                         int &iy, int &ix, int &iz, double (&vw)[2], double (&vh)[2]) {
    SamplePrep s[2];
    double4 ry[2], rx[2], rz[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        s[p].iy = iy;
        s[p].ix = ix;
        s[p].iz = iz;
        sample_prepare<GUESS_HINT, GUESS_HINT>(c, y[p], x[p], z[p], s[p]);
    }
    ry[0] = ry[1] = ld_cell(c.ay.cell + iy);
    rx[0] = rx[1] = ld_cell(c.ax.cell + ix);
    rz[0] = rz[1] = ld_cell(c.az.cell + iz);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        fix_interval(c.ay, s[p].y, s[p].iy, ry[p], s[p].inb);
        fix_interval(c.ax, s[p].x, s[p].ix, rx[p], s[p].inb);
        fix_interval(c.az, s[p].z, s[p].iz, rz[p], s[p].inb);
    }
    const int nzc = c.az.n - 1;
    const unsigned row = (unsigned)c.ax.n * (unsigned)nzc;
    double4 c00[2], c01[2], c10[2], c11[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const double4 *q = c.cells + ((unsigned)s[p].iy * row + (unsigned)s[p].ix * (unsigned)nzc + (unsigned)s[p].iz);
        c00[p] = ld_cell(q);
        c01[p] = ld_cell(q + nzc);
        c10[p] = ld_cell(q + row);
        c11[p] = ld_cell(q + row + nzc);
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const double ty = div_exact(s[p].y - ry[p].x, ry[p].z, ry[p].w);
        const double tx = div_exact(s[p].x - rx[p].x, rx[p].z, rx[p].w);
        const double tz = div_exact(s[p].z - rz[p].x, rz[p].z, rz[p].w);
        trilinear_scipy(c00[p], c01[p], c10[p], c11[p], ty, tx, tz, vw[p], vh[p]);
        if (!s[p].inb) vw[p] = vh[p] = __longlong_as_double(0x7ff8000000000000LL);
    }
    const int last = s[1].inb ? 1 : 0;
    if (s[last].inb) {
        iy = s[last].iy;
        ix = s[last].ix;
        iz = s[last].iz;
    }
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

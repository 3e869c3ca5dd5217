// k6_orbit.cuh -- K6: look vectors from orbit state vectors.
// A fragment of libraider_b200.so: included by raider_b200.cu INSIDE its anonymous namespace, in this order (the translation unit
// is the one file it used to be; see the kernel inventory at the top of raider_b200.cu).  Not a stand-alone header.
// ------------------------------------------------------------------------------------------------
// K6: look vectors from orbit state vectors -- replaces the per-pixel Python loop over isce3.geometry.geo2rdr +
// Orbit.interpolate of Raytracing.getLookVectors (losreader.py:219-255).  One thread per target: Newton iteration on the
// zero-Doppler condition (dr . v = 0) with the 4-point Hermite orbit interpolator (isce3's defaults; algorithm restated in
// oracle/orbit.py), threshold 1e-7 m on the slant range, at most 30 iterations, start at the orbit's mid time; a target that
// does not converge or leaves the orbit's time span gets a NaN vector, as the reference's try/except does.
// ------------------------------------------------------------------------------------------------
struct OrbitView {
    const double *t;    // [n] uniformly spaced
    const double *pos;  // [n][3]
    const double *vel;  // [n][3]
    int n;
    double inv_dt;
};

// ROI_PAC / ISCE orbitHermite on state vectors idx .. idx+3
__device__ __forceinline__ void orbit_hermite(const OrbitView &O, int idx, double time, Vec3 &p, Vec3 &v) {
    double t[4], f0[4], f1[4], h[4], hdot[4], g0[4], g1[4], isum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = __ldg(O.t + idx + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f1[i] = time - t[i];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j != i) s += 1.0 / (t[i] - t[j]);
        isum[i] = s;
        f0[i] = 1.0 - 2.0 * (time - t[i]) * s;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double product = 1.0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k != i) product *= (time - t[k]) / (t[i] - t[k]);
        h[i] = product;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double pr = 1.0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k != i && k != j) pr *= (time - t[k]) / (t[i] - t[k]);
            if (j != i) s += 1.0 / (t[i] - t[j]) * pr;
        }
        hdot[i] = s;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        g1[i] = h[i] + 2.0 * (time - t[i]) * hdot[i];
        g0[i] = 2.0 * (f0[i] * hdot[i] - h[i] * isum[i]);
    }
    p = {0.0, 0.0, 0.0};
    v = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double *x = O.pos + 3 * (idx + i), *w = O.vel + 3 * (idx + i);
        const double hh = h[i] * h[i];
        p.x += (__ldg(x) * f0[i] + __ldg(w) * f1[i]) * hh;
        p.y += (__ldg(x + 1) * f0[i] + __ldg(w + 1) * f1[i]) * hh;
        p.z += (__ldg(x + 2) * f0[i] + __ldg(w + 2) * f1[i]) * hh;
        v.x += (__ldg(x) * g0[i] + __ldg(w) * g1[i]) * h[i];
        v.y += (__ldg(x + 1) * g0[i] + __ldg(w + 1) * g1[i]) * h[i];
        v.z += (__ldg(x + 2) * g0[i] + __ldg(w + 2) * g1[i]) * h[i];
    }
}

// Orbit.interpolate with FillNaN borders; false outside [t[0], t[n-1]]
__device__ __forceinline__ bool orbit_interpolate(const OrbitView &O, double time, Vec3 &p, Vec3 &v) {
    const double t0 = __ldg(O.t), t1 = __ldg(O.t + O.n - 1);
    if (!(time >= t0 && time <= t1)) return false;
    // first state vector with t[i] >= time: guess from the spacing, settle on the stored times
    int i = (int)ceil((time - t0) * O.inv_dt);
    i = min(max(i, 0), O.n - 1);
    while (i > 0 && __ldg(O.t + i - 1) >= time) --i;
    while (i < O.n - 1 && __ldg(O.t + i) < time) ++i;
    const int idx = min(max(i - 2, 0), O.n - 4);
    orbit_hermite(O, idx, time, p, v);
    return true;
}

__global__ void k_orbit_los(const OrbitView O, int geom_kind, const double *__restrict__ gx, const double *__restrict__ gy,
                            const double *__restrict__ hgt, double ht, int nx, int64_t n, double threshold, int maxiter,
                            double *__restrict__ los, double *__restrict__ slant_out, double *__restrict__ aztime_out) {
    const double qn = qnan();
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        double lat, lon;
        if (geom_kind == RDR_GEOM_GRID) {
            lon = __ldg(gx + (r % nx));
            lat = __ldg(gy + (r / nx));
        } else {
            lon = __ldg(gx + r);
            lat = __ldg(gy + r);
        }
        const double h = hgt ? __ldg(hgt + r) : ht;
        double a, b, c2, d;
        const Vec3 g = lla2ecef(lat, lon, h, a, b, c2, d);
        double aztime = __ldg(O.t) + 0.5 * (__ldg(O.t + O.n - 1) - __ldg(O.t));
        double slant = 0.0, slant_old = 0.0;
        bool converged = false;
        Vec3 p, v;
        for (int it = 0; it < maxiter; ++it) {
            if (!orbit_interpolate(O, aztime, p, v)) break;  // NaN position: no comparison ever succeeds
            const Vec3 dr = g - p;
            slant = sqrt(dr.x * dr.x + dr.y * dr.y + dr.z * dr.z);
            if (fabs(slant - slant_old) < threshold) {
                converged = true;
                break;
            }
            slant_old = slant;
            const double fn = dr.x * v.x + dr.y * v.y + dr.z * v.z;
            const double fnprime = -(v.x * v.x + v.y * v.y + v.z * v.z);
            aztime -= fn / fnprime;
        }
        // losreader.py:252-253: sat_xyz, _ = orbit.interpolate(aztime); los = (sat_xyz - inp_xyz) / slant_range
        if (converged && (lat == lat) && (lon == lon) && (h == h)) {
            los[3 * r] = (p.x - g.x) / slant;
            los[3 * r + 1] = (p.y - g.y) / slant;
            los[3 * r + 2] = (p.z - g.z) / slant;
        } else {
            los[3 * r] = los[3 * r + 1] = los[3 * r + 2] = qn;
            slant = aztime = qn;
        }
        if (slant_out) slant_out[r] = slant;
        if (aztime_out) aztime_out[r] = aztime;
    }
}


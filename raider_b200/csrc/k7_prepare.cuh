// k7_prepare.cuh -- K7: weather-model column processing.
// A fragment of libraider_b200.so: included by raider_b200.cu INSIDE its anonymous namespace, in this order (the translation unit
// is the one file it used to be; see the kernel inventory at the top of raider_b200.cu).  Not a stand-alone header.
// ------------------------------------------------------------------------------------------------
// K7: weather-model processing, the step before the path -- WeatherModel.load after load_weather
// (models/weatherModel.py:252-260): _find_e, _uniform_in_z (3 x interpolate_along_axis, fill NaN, cast fp32), _checkForNans
// (fillna3D), refractivities, _adjust_grid (one level at zmin), _getZTD (cumulative trapezoid).
// One WARP per model column; z is the fastest axis of every array ((y, x, z) like the reference's), so the lanes of a warp
// read and write consecutive levels.  Per-warp shared memory: e[nl] f64 | p, t, e [nzo] f32 | wet, hydro [nzo] f32.
// ------------------------------------------------------------------------------------------------
struct PrepParams {
    int nl, nz_out, pad;      // native levels, target levels, 1 when a level at zmin is prepended
    int hum_is_rh;
    double k1, k2, k3, R_v, R_d, zmin;
};

__device__ __forceinline__ float find_svp_f32(double t) {  // weatherModel.py:750-780 (float64 arithmetic, float32 result)
    const double t1 = 273.15, t2 = 250.15;
    const double tref = t - t1, wgt = (t - t2) / (t1 - t2);
    const double svpw = 6.1121 * exp((17.502 * tref) / (240.97 + tref));
    const double svpi = 6.1121 * exp((22.587 * tref) / (273.86 + tref));
    double svp = svpi + (svpw - svpi) * (wgt * wgt);
    if (t > t1) svp = svpw;
    if (t < t2) svp = svpi;
    return (float)(svp * 100.0);
}

// fillna3D (interpolator.py:110-130) on one column held in shared memory: leading NaNs <- first valid value, interior NaNs <-
// linear in the level index between the valid neighbours, trailing NaNs <- fill
__device__ __forceinline__ void fill_column(float *v, int n, float fill, int lane) {
    int first = n, last = -1;
    for (int l = lane; l < n; l += 32)
        if (v[l] == v[l]) {
            first = min(first, l);
            last = max(last, l);
        }
    first = __reduce_min_sync(0xffffffffu, first);
    last = __reduce_max_sync(0xffffffffu, last);
    float nv[8];  // n <= 256
    int cnt = 0;
    for (int l = lane; l < n; l += 32, ++cnt) {
        float x = v[l];
        if (!(x == x)) {
            if (last < 0 || l > last) x = fill;
            else if (l < first) x = v[first];
            else {
                int a = l - 1, b = l + 1;
                while (!(v[a] == v[a])) --a;
                while (!(v[b] == v[b])) ++b;
                // np.interp in float64 on the index axis, stored back in the array's float32
                const double slope = ((double)v[b] - (double)v[a]) / (double)(b - a);
                x = (float)(slope * (double)(l - a) + (double)v[a]);
            }
        }
        nv[cnt] = x;
    }
    __syncwarp();
    cnt = 0;
    for (int l = lane; l < n; l += 32, ++cnt) v[l] = nv[cnt];
    __syncwarp();
}

__global__ void __launch_bounds__(128) k_prepare_columns(const PrepParams P, int64_t ncol, const double *__restrict__ zs, const double *__restrict__ p_in,
                                                         const double *__restrict__ t_in, const double *__restrict__ hum,
                                                         const double *__restrict__ zlev, float *__restrict__ out_wet,
                                                         float *__restrict__ out_hydro, float *__restrict__ out_wet_total,
                                                         float *__restrict__ out_hydro_total, float *__restrict__ out_p,
                                                         float *__restrict__ out_t, float *__restrict__ out_e) {
    extern __shared__ __align__(16) unsigned char prep_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nzo = P.nz_out + P.pad;
    const size_t per_warp = (size_t)P.nl * sizeof(double) + (size_t)5 * nzo * sizeof(float) + 16;
    unsigned char *base = prep_smem + (size_t)wib * ((per_warp + 15) / 16 * 16);
    double *s_e = reinterpret_cast<double *>(base);
    float *s_p = reinterpret_cast<float *>(s_e + P.nl), *s_t = s_p + nzo, *s_ee = s_t + nzo, *s_w = s_ee + nzo, *s_h = s_w + nzo;
    const float qn = __int_as_float(0x7fc00000);
    for (int64_t col = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; col < ncol; col += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const double *cz = zs + col * P.nl, *cp = p_in + col * P.nl, *ct = t_in + col * P.nl, *ch = hum + col * P.nl;
        // _find_e (weatherModel.py:333-354)
        for (int i = lane; i < P.nl; i += 32) {
            const double svp = (double)find_svp_f32(ct[i]);
            double e;
            if (P.hum_is_rh) e = ch[i] / 100.0 * svp;
            else {
                const double w = ch[i] / (1.0 - ch[i]);
                e = w * P.R_v * (cp[i] - svp) / P.R_d;
            }
            s_e[i] = e;
        }
        __syncwarp();
        // _uniform_in_z: interpolate_along_axis(zs, v, new_zs, fill_value=nan).astype(float32) (interpolate.h:78-118 per column)
        float *o_p = s_p + P.pad, *o_t = s_t + P.pad, *o_e = s_ee + P.pad;
        for (int l = lane; l < P.nz_out; l += 32) {
            const double v = __ldg(zlev + l);
            const int hi = bisect_left(cz, P.nl, v);
            float rp = qn, rt = qn, re = qn;
            if (hi >= 1 && hi <= P.nl - 1) {
                const double x0 = cz[hi - 1], x1 = cz[hi], dx = v - x0;
                rp = (float)__dadd_rn(cp[hi - 1], __dmul_rn(__ddiv_rn(cp[hi] - cp[hi - 1], x1 - x0), dx));
                rt = (float)__dadd_rn(ct[hi - 1], __dmul_rn(__ddiv_rn(ct[hi] - ct[hi - 1], x1 - x0), dx));
                re = (float)__dadd_rn(s_e[hi - 1], __dmul_rn(__ddiv_rn(s_e[hi] - s_e[hi - 1], x1 - x0), dx));
            }
            o_p[l] = rp;
            o_t[l] = rt;
            o_e[l] = re;
        }
        __syncwarp();
        // _checkForNans
        fill_column(o_p, P.nz_out, 0.0f, lane);
        fill_column(o_t, P.nz_out, 1e16f, lane);
        fill_column(o_e, P.nz_out, 0.0f, lane);
        // refractivities in float32, operation by operation as numpy evaluates k2 * e / t + k3 * e / t**2 and k1 * p / t
        const float k1 = (float)P.k1, k2 = (float)P.k2, k3 = (float)P.k3;
        float *o_w = s_w + P.pad, *o_h = s_h + P.pad;
        for (int l = lane; l < P.nz_out; l += 32) {
            const float e = o_e[l], t = o_t[l];
            o_w[l] = __fadd_rn(__fdiv_rn(__fmul_rn(k2, e), t), __fdiv_rn(__fmul_rn(k3, e), __fmul_rn(t, t)));
            o_h[l] = __fdiv_rn(__fmul_rn(k1, o_p[l]), t);
        }
        __syncwarp();
        if (P.pad && lane == 0) {  // _adjust_grid: the new lowest level repeats the first valid value (no NaNs are left)
            s_p[0] = o_p[0]; s_t[0] = o_t[0]; s_ee[0] = o_e[0]; s_w[0] = o_w[0]; s_h[0] = o_h[0];
        }
        __syncwarp();
        // _getZTD: total[l] = 1e-6 * sum_{m >= l} (z[m+1] - z[m]) * (f[m] + f[m+1]) / 2, the pair sum in float32 as np.trapz does
        auto zat = [&](int l) { return (P.pad && l == 0) ? P.zmin : __ldg(zlev + l - P.pad); };
        float *ow = out_wet + col * nzo, *oh = out_hydro + col * nzo, *owt = out_wet_total + col * nzo, *oht = out_hydro_total + col * nzo;
        for (int l = lane; l < nzo; l += 32) {
            double tw = 0.0, th = 0.0;
            for (int m = l; m + 1 < nzo; ++m) {
                const double d = zat(m + 1) - zat(m);
                tw += d * (double)__fadd_rn(s_w[m + 1], s_w[m]) / 2.0;
                th += d * (double)__fadd_rn(s_h[m + 1], s_h[m]) / 2.0;
            }
            ow[l] = s_w[l];
            oh[l] = s_h[l];
            owt[l] = (float)(1e-6 * tw);
            oht[l] = (float)(1e-6 * th);
            if (out_p) {
                out_p[col * nzo + l] = s_p[l];
                out_t[col * nzo + l] = s_t[l];
                out_e[col * nzo + l] = s_ee[l];
            }
        }
        __syncwarp();
    }
}


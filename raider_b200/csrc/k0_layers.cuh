// k0_layers.cuh -- ray geometry shared by K0 and K3, and K0: layer intersections of every ray + per-layer maxima + the whole-raster predicate counters.
// A fragment of libraider_b200.so: included by raider_b200.cu INSIDE its anonymous namespace, in this order (the translation unit
// is the one file it used to be; see the kernel inventory at the top of raider_b200.cu).  Not a stand-alone header.
// ------------------------------------------------------------------------------------------------
// ray geometry shared by K0 and K3
// ------------------------------------------------------------------------------------------------
struct RayGeom {
    int geom_kind, los_kind;
    const double *gx, *gy;  // GRID: xpts[nx], ypts[ny];  POINTS: lon[n], lat[n]
    const double *los;      // ARRAY: [n][3]
    double e, n, u;         // ENU_CONST
    double ht;
    int nx;
};

// Extra destinations of the integrator's results: the same row block of the delay maps in the HBM of the other GPUs of the node
// (peer-mapped symmetric memory, NVLink / NVSwitch).  The integration kernel stores every ray's two results to all of them as it
// finishes the ray -- the all-gather of SURVEY section 8(e) fused into K3 as posted peer writes: 16 B per ray and peer spread
// over the whole integration, instead of a collective after it.

template <typename OUT>
__device__ __forceinline__ void store_result(OUT *__restrict__ out_wet, OUT *__restrict__ out_hydro, const PeerOut &peers, int64_t r, double acc_w,
                                             double acc_h, int accumulate) {
    if (accumulate) {
        out_wet[r] = (OUT)((double)out_wet[r] + acc_w);
        out_hydro[r] = (OUT)((double)out_hydro[r] + acc_h);
        return;
    }
    __stcs(out_wet + r, (OUT)acc_w);
    __stcs(out_hydro + r, (OUT)acc_h);
    if (peers.multicast) {
        // one store into the multicast mapping of the symmetric maps: the NVSwitch replicates it into every GPU's copy
        if (sizeof(OUT) == 8) {
            asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(static_cast<OUT *>(peers.wet[0]) + r), "d"((double)acc_w) : "memory");
            asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(static_cast<OUT *>(peers.hydro[0]) + r), "d"((double)acc_h) : "memory");
        } else {
            asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(static_cast<OUT *>(peers.wet[0]) + r), "f"((float)acc_w) : "memory");
            asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(static_cast<OUT *>(peers.hydro[0]) + r), "f"((float)acc_h) : "memory");
        }
        return;
    }
    for (int p = 0; p < peers.n; ++p) {
        static_cast<OUT *>(peers.wet[p])[r] = (OUT)acc_w;
        static_cast<OUT *>(peers.hydro[p])[r] = (OUT)acc_h;
    }
}

__device__ __forceinline__ void ray_setup(const RayGeom &G, int64_t r, Vec3 &g, Vec3 &u, RayRef &R) {
    double lat, lon;
    if (G.geom_kind == RDR_GEOM_GRID) {
        lon = __ldg(G.gx + (r % G.nx));
        lat = __ldg(G.gy + (r / G.nx));
    } else {
        lon = __ldg(G.gx + r);
        lat = __ldg(G.gy + r);
    }
    double slat, clat, slon, clon;
    g = lla2ecef(lat, lon, G.ht, slat, clat, slon, clon);
    R.lat0_rad = lat * DEG_TO_RAD; R.lon0_rad = lon * DEG_TO_RAD;
    R.slat = slat; R.clat = clat; R.slon = slon; R.clon = clon;
    if (G.los_kind == RDR_LOS_ARRAY) {
        u = {__ldg(G.los + 3 * r), __ldg(G.los + 3 * r + 1), __ldg(G.los + 3 * r + 2)};
    } else if (G.los_kind == RDR_LOS_ENU_CONST) {
        u = enu2ecef(G.e, G.n, G.u, slat, clat, slon, clon);
    } else {  // zenith: getZenithLookVecs (losreader.py:312-314)
        u = {clat * clon, clat * slon, slat};
    }
}

// warp max of non-negative doubles via two 32-bit REDUX ops on the IEEE bit pattern (monotone for x >= 0)
__device__ __forceinline__ unsigned long long warp_max_bits(unsigned long long bits) {
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return ((unsigned long long)mhi << 32) | mlo;
}

// ------------------------------------------------------------------------------------------------
// K0: layer intersections for every ray + per-layer max length + first-sample-below counter
//   t_out[0][r]   = along-ray distance of the bottom of the first contributing layer
//   t_out[k+1][r] = along-ray distance of the top of contributing layer k
//   red[k]        = bits of max_r |P_hi - P_lo| (atomicMax on the bit pattern), red[K] = #NaN rays, red[K+1] = #first sample below zmin
// The ray lives in the meridian frame of its ground point (fastpath.cuh): 3 FMAs per Newton update, no longitude trig.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ray_latlon(const RayGeom &G, int64_t r, double &lat, double &lon) {
    if (G.geom_kind == RDR_GEOM_GRID) {
        lon = __ldg(G.gx + (r % G.nx));
        lat = __ldg(G.gy + (r / G.nx));
    } else {
        lon = __ldg(G.gx + r);
        lat = __ldg(G.gy + r);
    }
}

template <bool EXACT>
__device__ __forceinline__ void ray_layers_one(const RayFrame &F, int K, const double *__restrict__ plan, double *__restrict__ t_out, int64_t n_rays,
                                               int64_t r, bool valid, int lane, double zmin, double zmax, unsigned long long *srow,
                                               bool &any_nan) {
    double Alo, Blo, Zlo, Ahi = 0.0, Bhi = 0.0, Zhi = 0.0, rcosf = 1.0, t;
    for (int k = 0; k < K; ++k) {
        const double a = __ldg(plan + k), b = __ldg(plan + K + k);
        if (k == 0) {
            frame_top_of_atmosphere<10, EXACT>(F, a, 1.0, Alo, Blo, Zlo, t);
            if (valid) __stcs(t_out + r, t);
            // hint for the whole-raster clamp of delay.py:306-307: height of the very first sample, evaluated on the
            // same reconstructed point K3 will use (K3 re-evaluates the predicate itself and has the last word)
            const double A1 = fma(t, F.uA, F.A0), B1 = t * F.uB, Z1 = fma(t, F.uZ, F.Z0);
            const double h0 = EXACT ? ecef2height(Vec3{A1, B1, Z1}) : frame_height(A1, B1, Z1);
            const unsigned below = __ballot_sync(0xffffffffu, valid && (h0 < zmin));
            if (lane == 0 && below) srow[K + 1] += (unsigned long long)__popc(below);
            frame_top_of_atmosphere<10, EXACT>(F, b, 1.0, Ahi, Bhi, Zhi, t);
        } else {
            Alo = Ahi; Blo = Bhi; Zlo = Zhi;
            frame_top_of_atmosphere<3, EXACT>(F, b, rcosf, Ahi, Bhi, Zhi, t);
        }
        const double len = norm3(Vec3{Ahi - Alo, Bhi - Blo, Zhi - Zlo});
        if (k == 0) rcosf = len / (b - a);  // 1 / cos_factor of losreader.py:824-825
        if (valid) __stcs(t_out + (int64_t)(k + 1) * n_rays + r, t);
        const bool isn = !(len == len);
        any_nan |= isn;
        const unsigned long long bits = (valid && !isn) ? (unsigned long long)__double_as_longlong(len) : 0ull;
        const unsigned long long m = warp_max_bits(bits);
        if (lane == 0 && m > srow[k]) srow[k] = m;
    }
    // hint for the whole-raster upper clamp (delay.py:310-311): height of the very last sample (the top of the top layer)
    const double hK = EXACT ? ecef2height(Vec3{Ahi, Bhi, Zhi}) : frame_height(Ahi, Bhi, Zhi);
    const unsigned above = __ballot_sync(0xffffffffu, valid && (hK > zmax));
    if (lane == 0 && above) srow[K + 2] += (unsigned long long)__popc(above);
}

// K0 with the height along the ray as ONE polynomial.  h(t) along a straight ray is so smooth (k-th derivative ~ r^(1-k)) that the
// degree-7 interpolant through eight exact (PROJ-form) heights at t = i L / 7, L = the length of the whole ray, misses the exact
// height by < 1e-8 m for every incidence up to 80 deg (L = 500 km) -- which is the rounding noise of the PROJ-form height itself
// (p / cos(phi) - N at |h| ~ 1e5 m; measured 5 .. 8e-9 m against the oracle for 0 .. 80 deg incidence, 0 .. 80 deg latitude,
// three headings, two output heights: profiles/k0_septic_accuracy.py).  Every Newton iterate of getTopOfAtmosphere
// (losreader.py:720-733) is then 8 DFMA instead of a Bowring inversion (~50 DP instructions), and a ray needs 8 exact heights
// in all: instead of 3 per layer (119 on C2, 452 on the 145-node tables), and instead of the 3 per 6-km span of the first form of
// this idea (30 / 48), whose span tables lived in thread-local memory (1.6 GB of DRAM write-backs per 4e6 rays on the 145-node
// table).  The iteration itself -- start at t = toa, three (ten) updates divided by the cos factor -- is the reference's.
//   coefficient k of x^k, x = 2 t / L - 1, from the node values:  c = V^-1 f,  V^-1 exact rationals rounded once
__constant__ double c_septic_inv[8][8] = {
    {-5.0 / 2048.0, 49.0 / 2048.0, -245.0 / 2048.0, 1225.0 / 2048.0, 1225.0 / 2048.0, -245.0 / 2048.0, 49.0 / 2048.0, -5.0 / 2048.0},
    {5.0 / 2048.0, -343.0 / 10240.0, 1715.0 / 6144.0, -8575.0 / 2048.0, 8575.0 / 2048.0, -1715.0 / 6144.0, 343.0 / 10240.0, -5.0 / 2048.0},
    {12691.0 / 92160.0, -24451.0 / 18432.0, 63651.0 / 10240.0, -92659.0 / 18432.0, -92659.0 / 18432.0, 63651.0 / 10240.0, -24451.0 / 18432.0, 12691.0 / 92160.0},
    {-12691.0 / 92160.0, 171157.0 / 92160.0, -148519.0 / 10240.0, 648613.0 / 18432.0, -648613.0 / 18432.0, 148519.0 / 10240.0, -171157.0 / 92160.0, 12691.0 / 92160.0},
    {-16807.0 / 18432.0, 141659.0 / 18432.0, -36015.0 / 2048.0, 199283.0 / 18432.0, 199283.0 / 18432.0, -36015.0 / 2048.0, 141659.0 / 18432.0, -16807.0 / 18432.0},
    {16807.0 / 18432.0, -991613.0 / 92160.0, 84035.0 / 2048.0, -1394981.0 / 18432.0, 1394981.0 / 18432.0, -84035.0 / 2048.0, 991613.0 / 92160.0, -16807.0 / 18432.0},
    {117649.0 / 92160.0, -117649.0 / 18432.0, 117649.0 / 10240.0, -117649.0 / 18432.0, -117649.0 / 18432.0, 117649.0 / 10240.0, -117649.0 / 18432.0, 117649.0 / 92160.0},
    {-117649.0 / 92160.0, 823543.0 / 92160.0, -823543.0 / 30720.0, 823543.0 / 18432.0, -823543.0 / 18432.0, 823543.0 / 30720.0, -823543.0 / 92160.0, 117649.0 / 92160.0},
};
constexpr double K0_MAX_RAY = 3.0e5;  // rays longer than this (incidence beyond ~73 deg through an 80 km model) take the exact form: the error of a layer top in t is the height error over cos(incidence)

struct Septic {
    double c[8];
    double two_over_L;
};

__device__ __forceinline__ double septic_height(const Septic &S, double t) {
    const double x = fma(t, S.two_over_L, -1.0);
    double r = fma(x, S.c[7], S.c[6]);
#pragma unroll
    for (int k = 5; k >= 0; --k) r = fma(x, r, S.c[k]);
    return r;
}

template <int ITERS>
__device__ __forceinline__ double septic_top_of_atmosphere(const Septic &S, double toa, double rfactor) {
    double t = toa;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) t = fma(toa - septic_height(S, t), rfactor, t);
    return t;
}

// The layer tops as ONE polynomial in the level height.  What the reference stores for a layer top is not the root of h(t) = z but
// the third iterate of its fixed-slope Newton scheme started at t = z (losreader.py:720-733 with factor = the first layer's cos
// factor): T(z) = g_z(g_z(g_z(z))), g_z(t) = t + (z - h(t)) / factor.  For one ray that is a smooth function of z alone (h(t) is
// the septic above, the factor is fixed once the first layer is done), and the degree-7 interpolant through its values at the
// eight Chebyshev nodes of [top of layer 1, top of layer K - 1] misses it by <= 1.6e-8 m up to 70 deg incidence through the
// 80 km of the 145-node tables (<= 1e-8 m up to 60 deg; profiles/k0_tfit_accuracy.py) -- the size of the rounding noise of the
// PROJ-form height that the septic itself carries.  A layer top is then 7 DFMA (two layers interleaved: no dependent chain
// between them) instead of three dependent Horner evaluations (30 DFMA): per ray 8 x 30 for the nodes + 56 for the coefficients
// + 7 K, i.e. 1270 instead of 4170 DFMA on the 145-node tables.  The first layer (ten iterations at factor 1, which defines the
// factor) is evaluated as before.  Layer x positions are ray independent: s_x[k], computed once per CTA.
__constant__ double c_tfit_u[8] = {  // (x_j + 1) / 2, x_j = cos(pi (2 j + 1) / 16)
    0.9903926402016152, 0.9157348061512726, 0.7777851165098011, 0.5975451610080641,
    0.40245483899193585, 0.22221488349019886, 0.08426519384872738, 0.009607359798384785};
__constant__ double c_tfit_inv[8][8] = {  // inverse Vandermonde matrix of the Chebyshev nodes (monomials in x), 50-digit arithmetic rounded once
    {-0.02486404592245725, 0.08352232973991236, -0.18707572033318612, 0.628417436515731, 0.628417436515731, -0.18707572033318612, 0.08352232973991236, -0.02486404592245725},
    {-0.025351161379823003, 0.10045145186799834, -0.3367274004519704, 3.2211615113525687, -3.2211615113525687, 0.3367274004519704, -0.10045145186799834, 0.025351161379823003},
    {0.7698016495254523, -2.5519026177451503, 5.380329742491341, -3.5982287742716426, -3.5982287742716426, 5.380329742491341, -2.5519026177451503, 0.7698016495254523},
    {0.7848829554303298, -3.069147182274407, 9.684337681751762, -18.443912220177555, 18.443912220177555, -9.684337681751762, 3.069147182274407, -0.7848829554303298},
    {-3.1779876260079822, 9.672340827762346, -12.500767952508536, 6.006414750754172, 6.006414750754172, -12.500767952508536, 9.672340827762346, -3.1779876260079822},
    {-3.2402480843731825, 11.632825402935941, -22.500787856406752, 30.787866300500635, -30.787866300500635, 22.500787856406752, -11.632825402935941, 3.2402480843731825},
    {3.0614674589207183, -7.391036260090294, 7.391036260090294, -3.0614674589207183, -3.0614674589207183, 7.391036260090294, -7.391036260090294, 3.0614674589207183},
    {3.1214451522580524, -8.889123728313635, 13.303513796840724, -15.692564486451687, 15.692564486451687, -13.303513796840724, 8.889123728313635, -3.1214451522580524},
};
constexpr int K0_TFIT_MIN = 16;  // fewest layers for which the fit pays (8 node solves = 8 layers' worth of iterations)

// returns false (nothing stored or counted) when a ray of the warp is too long for the polynomial: the caller redoes the warp exactly.
// s_plan: low[K] | high[K] | x[K] (fit coordinate of the layer tops, TFIT only) in shared memory.
template <bool TFIT>
__device__ __forceinline__ bool ray_layers_septic(const RayFrame &F, double ht, int K, const double *__restrict__ s_plan,
                                                  double *__restrict__ t_out, int64_t n_rays, int64_t rr, bool valid, int lane, double zmin,
                                                  double zmax, unsigned long long *srow, bool &any_nan) {
    const double unorm = norm3(Vec3{F.uA, F.uB, F.uZ});
    // length of the whole ray from the incidence at the ground point: cos = look . ellipsoid normal (curvature only shortens it)
    const double cos0 = fma(F.uA, F.clat, F.uZ * F.slat) / unorm;
    const double L = fma(1.05, (s_plan[2 * K - 1] - fmin(ht, s_plan[0])) / cos0, 100.0);
    const bool too_long = !(L > 0.0 && L < K0_MAX_RAY);  // (NaN look vectors land here too: the exact form propagates the NaN)
    if (__any_sync(0xffffffffu, too_long)) return false;
    Septic S;
    {
        double f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double ti = L * ((double)i / 7.0);
            f[i] = frame_height(fma(ti, F.uA, F.A0), ti * F.uB, fma(ti, F.uZ, F.Z0));
        }
#pragma unroll
        for (int i = 1; i < 8; ++i) f[i] -= f[0];  // differences from the ground height: the products below stay at the size of the variation
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double a = c_septic_inv[k][1] * f[1];
#pragma unroll
            for (int i = 2; i < 8; ++i) a = fma(c_septic_inv[k][i], f[i], a);
            S.c[k] = a;
        }
        S.c[0] += f[0];
        S.two_over_L = 2.0 / L;
    }
    const double a0 = s_plan[0], b0 = s_plan[K];
    double t_lo = septic_top_of_atmosphere<10>(S, a0, 1.0);
    double t_hi = septic_top_of_atmosphere<10>(S, b0, 1.0);
    double len = fabs(t_hi - t_lo) * unorm;  // |P_hi - P_lo| (losreader.py:821): the points are g + t u
    const double rcosf = len / (b0 - a0);    // 1 / cos_factor of losreader.py:824-825
    {
        // hint for the whole-raster clamp of delay.py:306-307: height of the very first sample, evaluated exactly on the point K3
        // will reconstruct (K3 re-evaluates the predicate itself and has the last word)
        const double h0 = frame_height(fma(t_lo, F.uA, F.A0), t_lo * F.uB, fma(t_lo, F.uZ, F.Z0));
        const unsigned below = __ballot_sync(0xffffffffu, valid && (h0 < zmin));
        if (lane == 0 && below) srow[K + 1] += (unsigned long long)__popc(below);
    }
    // The lanes past the end of the raster (last warp only) carry a copy of the last ray (rr = n_rays - 1): they compute and store
    // the same values to the same addresses and cannot change a maximum, so the layer loop needs no `valid` predicate.
    double *tp = t_out + rr;
    __stcs(tp, t_lo);
    const uint32_t srow_s = smem_u32(srow);
    // top of layer k at distance t_top, the layer's chord length: store, NaN flag, warp maximum (this warp's row: no atomics)
    auto emit = [&](int k, double t_top, double length) {
        tp += n_rays;
        __stcs(tp, t_top);
        const unsigned hi = (unsigned)__double2hiint(length), lo = (unsigned)__double2loint(length);
        const bool isn = hi > 0x7ff00000u || (hi == 0x7ff00000u && lo != 0u);  // length >= 0 (fabs): NaN by its bit pattern
        any_nan |= isn;
        const unsigned h1 = isn ? 0u : hi;
        const unsigned mhi = __reduce_max_sync(0xffffffffu, h1);
        const unsigned mlo = __reduce_max_sync(0xffffffffu, h1 == mhi ? lo : 0u);
        const unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
        // lane 0 alone reads and updates the warp's row, by predicate (no branch).  Letting every lane read the row (a broadcast whose
        // value only lane 0 uses) is 9 % faster for K0, but it is a read / write pair between lanes without a barrier in between,
        // which racecheck reports; this form is clean (profiles/r02w_racecheck.txt: 0 hazards).
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .u64 cur;\n"
            "setp.eq.u32 p, %2, 0;\n"
            "@p ld.shared.u64 cur, [%0];\n"
            "@p setp.gt.u64 p, %1, cur;\n"
            "@p st.shared.u64 [%0], %1;\n"
            "}\n" ::"r"(srow_s + 8u * (unsigned)k),
            "l"(m), "r"(lane));
    };
    emit(0, t_hi, len);
    if (TFIT) {
        double c[8];
        {
            const double zA = s_plan[K + 1], dz = s_plan[2 * K - 1] - zA;
            double T[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) T[j] = septic_top_of_atmosphere<3>(S, fma(c_tfit_u[j], dz, zA), rcosf);  // eight independent chains
#pragma unroll
            for (int j = 1; j < 8; ++j) T[j] -= T[0];  // (row sums of the inverse: 1 for k = 0, 0 above -- the constant goes back into c0)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                double a = c_tfit_inv[k][1] * T[1];
#pragma unroll
                for (int j = 2; j < 8; ++j) a = fma(c_tfit_inv[k][j], T[j], a);
                c[k] = a;
            }
            c[0] += T[0];
        }
        const double *s_x = s_plan + 2 * K;
        int k = 1;
        for (; k + 1 < K; k += 2) {  // two layers at a time: independent Horner chains
            const double x0 = s_x[k], x1 = s_x[k + 1];
            double r0 = fma(x0, c[7], c[6]), r1 = fma(x1, c[7], c[6]);
#pragma unroll
            for (int i = 5; i >= 0; --i) {
                r0 = fma(x0, r0, c[i]);
                r1 = fma(x1, r1, c[i]);
            }
            emit(k, r0, fabs(r0 - t_hi) * unorm);
            emit(k + 1, r1, fabs(r1 - r0) * unorm);
            t_hi = r1;
        }
        if (k < K) {
            const double x0 = s_x[k];
            double r0 = fma(x0, c[7], c[6]);
#pragma unroll
            for (int i = 5; i >= 0; --i) r0 = fma(x0, r0, c[i]);
            emit(k, r0, fabs(r0 - t_hi) * unorm);
            t_hi = r0;
        }
    } else {
        for (int k = 1; k < K; ++k) {
            t_lo = t_hi;
            t_hi = septic_top_of_atmosphere<3>(S, s_plan[K + k], rcosf);
            emit(k, t_hi, fabs(t_hi - t_lo) * unorm);
        }
    }
    {
        // hint for the whole-raster upper clamp (delay.py:310-311): height of the very last sample (the top of the top layer),
        // evaluated exactly on the point K3 will reconstruct.  With zref at its default (1 m below the model top) the reference's
        // three iterates overshoot the top by more than that metre from ~58 deg incidence on (80 km tables).
        const double hK = frame_height(fma(t_hi, F.uA, F.A0), t_hi * F.uB, fma(t_hi, F.uZ, F.Z0));
        const unsigned above = __ballot_sync(0xffffffffu, valid && (hK > zmax));
        if (lane == 0 && above) srow[K + 2] += (unsigned long long)__popc(above);
    }
    return true;
}

template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_ray_layers(const RayGeom G, int64_t n_rays, int K, const double *__restrict__ plan,
                                                      double *__restrict__ t_out, unsigned long long *__restrict__ red, double zmin,
                                                      double zmax, int use_poly) {
    extern __shared__ unsigned long long smax[];  // [BLOCK / 32][K + 3] maxima / counters (#NaN, #first below, #last above) per warp | low[K] | high[K] | x[K]
    constexpr int NW = BLOCK / 32;
    double *s_plan = reinterpret_cast<double *>(smax + NW * (K + 3));
    const bool tfit = use_poly == 2 && K >= K0_TFIT_MIN;
    for (int i = threadIdx.x; i < NW * (K + 3); i += BLOCK) smax[i] = 0ull;
    for (int i = threadIdx.x; i < 2 * K; i += BLOCK) s_plan[i] = plan[i];
    if (tfit) {  // fit coordinate of every layer top: x = 2 (z - zA) / (zB - zA) - 1 on [top of layer 1, top of layer K - 1]
        const double zA = plan[K + 1], two_inv = 2.0 / (plan[2 * K - 1] - zA);
        for (int i = threadIdx.x; i < K; i += BLOCK) s_plan[2 * K + i] = fma(plan[K + i] - zA, two_inv, -1.0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned long long *srow = smax + (threadIdx.x >> 5) * (K + 3);  // this warp's maxima / counters (lane 0 writes: no atomics)
    const int64_t n_pad = (n_rays + 31) / 32 * 32;
    for (int64_t r = blockIdx.x * (int64_t)BLOCK + threadIdx.x; r < n_pad; r += (int64_t)gridDim.x * BLOCK) {
        const bool valid = r < n_rays;
        const int64_t rr = valid ? r : n_rays - 1;
        double lat, lon;
        ray_latlon(G, rr, lat, lon);
        RayFrame F;
        frame_setup(lat, lon, G.ht, G.los_kind, G.los, rr, G.e, G.n, G.u, F);
        bool any_nan = false;
        // the branch is taken per warp (all lanes vote): the ballots / REDUX inside need the full warp
        if (__all_sync(0xffffffffu, F.fast_ok)) {
            // (a warp with a ray too long for the polynomial bails out of that form before storing or counting anything)
            const bool done = !use_poly ? false
                              : tfit    ? ray_layers_septic<true>(F, G.ht, K, s_plan, t_out, n_rays, rr, valid, lane, zmin, zmax, srow, any_nan)
                                        : ray_layers_septic<false>(F, G.ht, K, s_plan, t_out, n_rays, rr, valid, lane, zmin, zmax, srow, any_nan);
            if (!done) ray_layers_one<false>(F, K, plan, t_out, n_rays, r, valid, lane, zmin, zmax, srow, any_nan);
        } else {
            ray_layers_one<true>(F, K, plan, t_out, n_rays, r, valid, lane, zmin, zmax, srow, any_nan);
        }
        const unsigned nn = __ballot_sync(0xffffffffu, valid && any_nan);
        if (lane == 0 && nn) srow[K] += (unsigned long long)__popc(nn);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K + 3; i += BLOCK) {
        unsigned long long v = smax[i];
        for (int w = 1; w < NW; ++w) {
            const unsigned long long u = smax[w * (K + 3) + i];
            v = i < K ? max(v, u) : v + u;
        }
        if (v) {
            // red: maxima [K] | #NaN rays | #first sample below | (#rays) | (K3's #first below) | #last sample above
            if (i < K) atomicMax(red + i, v); else atomicAdd(red + (i == K + 2 ? K + 4 : i), v);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) red[K + 2] = (unsigned long long)n_rays;  // the slot carries the call's ray count (k_plan)
}


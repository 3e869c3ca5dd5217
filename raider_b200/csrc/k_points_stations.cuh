// k_points_stations.cuh -- K1b sample points of the rays; K5 station (point) mode.
// A fragment of libraider_b200.so: included by raider_b200.cu INSIDE its anonymous namespace, in this order (the translation unit
// is the one file it used to be; see the kernel inventory at the top of raider_b200.cu).  Not a stand-alone header.
// ------------------------------------------------------------------------------------------------
// K1b: materialise the model-coordinate sample points of the rays (the per-sub-step `pts` arrays of delay.py:292-298)
// for the unfused pipeline / the K2 roofline measurement: pts[(slot - slot0) * n_rays + r] = (y, x, z), slots counted
// over the unique samples in layer-then-step order.  Each warp writes 32 x 24 contiguous bytes per slot.
// ------------------------------------------------------------------------------------------------
template <typename T, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_ray_points(const CubeView c, const RayGeom G, int64_t n_rays, int K, const double *__restrict__ t_in,
                                                      const int *__restrict__ nparts, int slot0, int nslots, T *__restrict__ pts) {
    for (int64_t r = blockIdx.x * (int64_t)BLOCK + threadIdx.x; r < n_rays; r += (int64_t)gridDim.x * BLOCK) {
        Vec3 g, u;
        RayRef R;
        ray_setup(G, r, g, u, R);
        Vec3 lo = ray_point(g, u, __ldg(t_in + r));
        int slot = 0;
        for (int k = 0; k < K && slot < slot0 + nslots; ++k) {
            const Vec3 hi = ray_point(g, u, __ldg(t_in + (int64_t)(k + 1) * n_rays + r));
            const Vec3 d = hi - lo;
            const int np = __ldg(nparts + k);
            const double step = 1.0 / (double)(np - 1);
            for (int j = (k == 0 ? 0 : 1); j < np; ++j, ++slot) {
                if (slot < slot0) continue;
                if (slot >= slot0 + nslots) break;
                const double ff = (j == np - 1) ? 1.0 : (double)j * step;
                double lon, lat, h;
                ecef2lla_fast({fma(ff, d.x, lo.x), fma(ff, d.y, lo.y), fma(ff, d.z, lo.z)}, R, lon, lat, h);
                double X = lon, Y = lat;
                if (c.crs_kind == RDR_CRS_LCC_SPHERE) {
                    const double2 xy = lcc_forward(c.lcc, lon, lat);
                    X = xy.x;
                    Y = xy.y;
                }
                T *o = pts + ((int64_t)(slot - slot0) * n_rays + r) * 3;
                o[0] = (T)Y;
                o[1] = (T)X;
                o[2] = (T)h;
            }
            lo = hi;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K5: station (point) mode -- one WARP per ray.  Every station is its own 1 x 1 raster with its own height (BASELINE C4: GNSS
// stations), i.e. _build_cube_ray(xpts=[lon], ypts=[lat], zpts=[h]) per station: the layer plan (losreader.py:785-809), the
// step counts nParts = ceil(L_k / S) + 1 (delay.py:283, the "raster maximum" is the ray's own length) and the clamps
// (delay.py:306-311: `.all()` over one pixel) are per ray, so K0, the reduction and K3 collapse into one kernel:
//   lane 0 .. 31 own the contributing layers k = lane, lane + 32, ...: every layer top is an independent Newton solve from
//   g + b u (losreader.py:727), only the cos factor of the first layer is shared (shuffle); each lane then walks the sub-steps
//   of its layers with the PROJ-form sampler, and the per-lane partial sums meet in a warp-shuffle reduction.
// 10 000 stations = 10 000 warps: the raster kernels would leave 3/4 of the machine idle on this shape.
// ------------------------------------------------------------------------------------------------
struct StationGeom {
    const double *lon, *lat, *hgt;  // [n] degrees, degrees, metres
    const double *los;              // [n][3]: ECEF (RDR_LOS_ARRAY) or local ENU (RDR_LOS_ENU_ARRAY) unit vectors ground -> sensor
    int los_kind;                   // RDR_LOS_ARRAY, RDR_LOS_ENU_ARRAY, RDR_LOS_ENU_CONST (e, n, u below) or RDR_LOS_ZENITH
    double e, n, u;
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_ray_stations(const CubeView c, const StationGeom S, int64_t n_rays, double zref, double max_seg,
                                                        double *__restrict__ out_wet, double *__restrict__ out_hydro,
                                                        int *__restrict__ out_nsamples) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)BLOCK + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * BLOCK) >> 5;
    const int nz = c.az.n;
    const double zmin = c.az.g_first, zmax = c.az.g_last;
    for (int64_t r = warp0; r < n_rays; r += nwarps) {
        const double lat = __ldg(S.lat + r), lon = __ldg(S.lon + r), ht = __ldg(S.hgt + r);
        double slat, clat, slon, clon;
        const Vec3 g = lla2ecef(lat, lon, ht, slat, clat, slon, clon);
        RayRef R;
        R.lat0_rad = lat * DEG_TO_RAD; R.lon0_rad = lon * DEG_TO_RAD;
        R.slat = slat; R.clat = clat; R.slon = slon; R.clon = clon;
        Vec3 u;
        if (S.los_kind == RDR_LOS_ARRAY) u = {__ldg(S.los + 3 * r), __ldg(S.los + 3 * r + 1), __ldg(S.los + 3 * r + 2)};
        else if (S.los_kind == RDR_LOS_ENU_ARRAY) u = enu2ecef(__ldg(S.los + 3 * r), __ldg(S.los + 3 * r + 1), __ldg(S.los + 3 * r + 2), slat, clat, slon, clon);
        else if (S.los_kind == RDR_LOS_ENU_CONST) u = enu2ecef(S.e, S.n, S.u, slat, clat, slon, clon);
        else u = {clat * clon, clat * slon, slat};
        // layer plan of this station: contributing model layers in order (scalar rules of losreader.py:785-809)
        auto plan = [&](int zz, double &lo_h, double &hi_h) -> bool {
            lo_h = __ldg(c.az.g + zz);
            hi_h = __ldg(c.az.g + zz + 1);
            if (hi_h == zmax) hi_h -= 0.01;
            if (hi_h < ht || lo_h >= zref) return false;
            if (lo_h < ht) lo_h = ht;
            if (hi_h > zref) hi_h = zref;
            return !(fabs(hi_h - lo_h) < 1.0);
        };
        int first = -1, count = 0;
        for (int zz = 0; zz < nz - 1; ++zz) {
            double a, b;
            if (plan(zz, a, b)) {
                if (first < 0) first = zz;
                ++count;
            }
        }
        double acc_w = 0.0, acc_h = 0.0;
        int nsamp = 0;
        if (count > 0) {
            // first contributing layer: 10 + 10 fixed-point iterations with factor 1 -> cos factor for every other layer
            double a0, b0, t;
            plan(first, a0, b0);
            const Vec3 lo0 = top_of_atmosphere<10>(g, u, a0, 1.0, t);
            const Vec3 hi0 = top_of_atmosphere<10>(g, u, b0, 1.0, t);
            const double len0 = norm3(hi0 - lo0);
            const double rcosf = len0 / (b0 - a0);
            // contributing layers are contiguous in zz except for sub-metre layers; walk them in order, lane-strided
            int k = 0;
            Vec3 prev_hi = hi0;  // top of the previous contributing layer (recomputed per lane: 3 iterations, no exchange needed)
            for (int zz = first; zz < nz - 1; ++zz) {
                double a, b;
                if (!plan(zz, a, b)) continue;
                const bool mine = (k & 31) == lane;
                if (mine || ((k + 1) & 31) == lane) {  // this lane needs the top of layer k either as its `hi` or as the next one's `lo`
                    Vec3 hi = hi0;
                    if (k > 0) hi = top_of_atmosphere<3>(g, u, b, rcosf, t);
                    if (mine) {
                        const Vec3 lo = k == 0 ? lo0 : prev_hi;
                        const Vec3 d = hi - lo;
                        const double len = norm3(d);
                        const double q = ceil(len / max_seg);
                        const int np = (q == q && q < 1e7) ? max(2, (int)q + 1) : 2;
                        const double step = 1.0 / (double)(np - 1), wt_full = (len * 1.0e-6) / ((double)np - 1.0);
                        int iy = -1, ix = -1, iz = zz;
                        for (int j = 0; j < np; ++j) {
                            const double ff = (j == np - 1) ? 1.0 : (double)j * step;
                            double lo_deg, la_deg, h, X, Y, vw, vh;
                            ecef2lla_fast({fma(ff, d.x, lo.x), fma(ff, d.y, lo.y), fma(ff, d.z, lo.z)}, R, lo_deg, la_deg, h);
                            X = lo_deg;
                            Y = la_deg;
                            if (c.crs_kind == RDR_CRS_LCC_SPHERE) {
                                const double2 xy = lcc_forward(c.lcc, lo_deg, la_deg);
                                X = xy.x;
                                Y = xy.y;
                            }
                            if (h < zmin) h = zmin;  // delay.py:306-311 with a one-pixel raster: `.all()` is the pixel itself
                            if (h > zmax) h = zmax;
                            if (iy < 0) sample_scipy<GUESS_BINS, GUESS_HINT>(c, Y, X, h, iy, ix, iz, vw, vh);
                            else sample_scipy<GUESS_HINT, GUESS_HINT>(c, Y, X, h, iy, ix, iz, vw, vh);
                            if (iy < 0) iy = ix = -1;
                            const double wt = (j == 0 || j == np - 1) ? 0.5 * wt_full : wt_full;
                            acc_w = fma(wt, vw, acc_w);
                            acc_h = fma(wt, vh, acc_h);
                        }
                        nsamp += np;
                    }
                    prev_hi = hi;
                }
                ++k;
            }
        }
        // warp-shuffle accumulator: partial integrals of the lanes' layers -> lane 0
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            acc_w += __shfl_down_sync(0xffffffffu, acc_w, off);
            acc_h += __shfl_down_sync(0xffffffffu, acc_h, off);
            nsamp += __shfl_down_sync(0xffffffffu, nsamp, off);
        }
        if (lane == 0) {
            out_wet[r] = acc_w;
            out_hydro[r] = acc_h;
            if (out_nsamples) out_nsamples[r] = nsamp;
        }
    }
}


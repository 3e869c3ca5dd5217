// k3_general.cuh -- K3, PROJ-form arithmetic per sample: the fallback integrator (flagged rays in list mode, any CRS).
// A fragment of libraider_b200.so: included by raider_b200.cu INSIDE its anonymous namespace, in this order (the translation unit
// is the one file it used to be; see the kernel inventory at the top of raider_b200.cu).  Not a stand-alone header.
// ------------------------------------------------------------------------------------------------
// K3: fused integrate.  One thread per ray; all lanes of a warp walk the same (layer, step) sequence because the
// step counts are global (delay.py:283), so there is no divergence and neighbouring rays hit the same cube cells.
// The sample at a layer interface is evaluated once and used with both layers' end weights (the reference evaluates
// the same point twice, delay.py:290-323).
// ------------------------------------------------------------------------------------------------
template <typename OUT, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_ray_integrate(const CubeView c, const RayGeom G, int64_t n_rays, int K,
                                                         const double *__restrict__ t_in, const DevPlan *__restrict__ P, double zmin, double zmax,
                                                         OUT *__restrict__ out_wet, OUT *__restrict__ out_hydro, int accumulate, const PeerOut peers,
                                                         unsigned long long *__restrict__ counters, const int *__restrict__ list,
                                                         const unsigned long long *__restrict__ list_count) {
    // list mode (list != nullptr): only the rays the fast integrator flagged, *list_count of them (read on the device, so the
    // launch needs no host round trip and is a no-op when nothing was flagged); their first samples were already counted
    const int lane = threadIdx.x & 31;
    const int64_t n_items = list ? (int64_t)*list_count : n_rays;
    if (n_items == 0 || P->blocked) return;
    const int *__restrict__ nparts = P->nparts;
    const int *__restrict__ layer_cell = P->layer_cell;
    const int clamp_low_first = P->clamp_low_first, clamp_high_last = P->clamp_high_last;
    const int64_t n_pad = (n_items + 31) / 32 * 32;
    unsigned n_below = 0, n_above = 0, n_first_below = 0;
    for (int64_t idx = blockIdx.x * (int64_t)BLOCK + threadIdx.x; idx < n_pad; idx += (int64_t)gridDim.x * BLOCK) {
        const bool valid = idx < n_items;
        const int64_t r = list ? (int64_t)__ldg(list + (valid ? idx : n_items - 1)) : idx;
        const int64_t rr = list ? r : (valid ? r : n_rays - 1);
        Vec3 g, u;
        RayRef R;
        ray_setup(G, rr, g, u, R);
        double acc_w = 0.0, acc_h = 0.0;
        Vec3 lo = ray_point(g, u, __ldcs(t_in + rr));
        Vec3 hi = ray_point(g, u, __ldcs(t_in + n_rays + rr));
        double len = norm3(hi - lo);
        double vw = 0.0, vh = 0.0;
        double gx0 = R.lon0_rad * RAD_TO_DEG, gy0 = R.lat0_rad * RAD_TO_DEG;
        if (c.crs_kind == RDR_CRS_LCC_SPHERE) {
            const double2 xy = lcc_forward(c.lcc, gx0, gy0);
            gx0 = xy.x;
            gy0 = xy.y;
        }
        // interval hints for the march: the ground point's own cell (clamped into the grid when the pixel hangs outside)
        int iy = guess_interval<GUESS_BINS>(c.ay, fmin(fmax(gy0, c.ay.g_first), c.ay.g_last), 0);
        int ix = guess_interval<GUESS_BINS>(c.ax, fmin(fmax(gx0, c.ax.g_first), c.ax.g_last), 0);
        // model-CRS coordinates + height of a sample, with the whole-raster bookkeeping of delay.py:306-311
        auto to_model = [&](double lon, double lat, double &X, double &Y) {
            X = lon;
            Y = lat;
            if (c.crs_kind == RDR_CRS_LCC_SPHERE) {
                const double2 xy = lcc_forward(c.lcc, lon, lat);
                X = xy.x;
                Y = xy.y;
            }
        };
        auto count_oob = [&](double h) {
            if (!(h >= zmin && h <= zmax)) {  // rare: counts feed the whole-raster predicate checks on the host
                n_below += valid && (h < zmin);
                n_above += valid && (h > zmax);
            }
        };
        for (int k = 0; k < K; ++k) {
            const Vec3 d = hi - lo;
            const int np = __ldg(nparts + k);
            int iz = __ldg(layer_cell + k);
            const double step = 1.0 / (double)(np - 1);                 // np.linspace(0, 1, np): j * step, last = 1.0
            const double wt_full = (len * 1.0e-6) / ((double)np - 1.0);  // delay.py:315
            const double wt_half = 0.5 * wt_full;
            int j = 1;
            if (k == 0) {  // very first sample of the ray (ff = 0)
                double lon, lat, h, X, Y;
                ecef2lla_fast(lo, R, lon, lat, h);
                to_model(lon, lat, X, Y);
                const unsigned b = __ballot_sync(0xffffffffu, valid && (h < zmin));
                n_first_below += __popc(b);
                if (clamp_low_first) h = zmin;  // all pixels below min(z): delay.py:306-307
                count_oob(h);
                sample_scipy<GUESS_HINT, GUESS_HINT>(c, Y, X, h, iy, ix, iz, vw, vh);
            }
            // first sample of this layer == last sample of the previous one (evaluated once, used with both end weights)
            acc_w = __dadd_rn(acc_w, __dmul_rn(wt_half, vw));
            acc_h = __dadd_rn(acc_h, __dmul_rn(wt_half, vh));
            for (; j + 1 < np; j += 2) {  // two interior / end samples per trip: independent chains keep the FP64 pipe busy
                const double fa = (double)j * step, fb = (j + 1 == np - 1) ? 1.0 : (double)(j + 1) * step;
                const Vec3 pa = {fma(fa, d.x, lo.x), fma(fa, d.y, lo.y), fma(fa, d.z, lo.z)};  // delay.py:292
                const Vec3 pb = {fma(fb, d.x, lo.x), fma(fb, d.y, lo.y), fma(fb, d.z, lo.z)};
                double lon[2], lat[2], hh[2], X[2], Y[2], sw[2], sh[2];
                ecef2lla_fast2(pa, pb, R, lon[0], lat[0], hh[0], lon[1], lat[1], hh[1]);
                to_model(lon[0], lat[0], X[0], Y[0]);
                to_model(lon[1], lat[1], X[1], Y[1]);
                if (clamp_high_last && k == K - 1 && j + 1 == np - 1) hh[1] = zmax;  // all pixels above max(z): delay.py:310-311
                count_oob(hh[0]);
                count_oob(hh[1]);
                sample_scipy_pair_hinted(c, Y, X, hh, iy, ix, iz, sw, sh);
                const double wb = (j + 1 == np - 1) ? wt_half : wt_full;
                acc_w = __dadd_rn(acc_w, __dmul_rn(wt_full, sw[0]));
                acc_h = __dadd_rn(acc_h, __dmul_rn(wt_full, sh[0]));
                acc_w = __dadd_rn(acc_w, __dmul_rn(wb, sw[1]));
                acc_h = __dadd_rn(acc_h, __dmul_rn(wb, sh[1]));
                vw = sw[1];
                vh = sh[1];
            }
            if (j < np) {  // odd one out: always the layer's last sample (ff = 1)
                const Vec3 p = {fma(1.0, d.x, lo.x), fma(1.0, d.y, lo.y), fma(1.0, d.z, lo.z)};
                double lon, lat, h, X, Y;
                ecef2lla_fast(p, R, lon, lat, h);
                to_model(lon, lat, X, Y);
                if (clamp_high_last && k == K - 1) h = zmax;  // all pixels above max(z): delay.py:310-311
                count_oob(h);
                sample_scipy<GUESS_HINT, GUESS_HINT>(c, Y, X, h, iy, ix, iz, vw, vh);
                acc_w = __dadd_rn(acc_w, __dmul_rn(wt_half, vw));
                acc_h = __dadd_rn(acc_h, __dmul_rn(wt_half, vh));
            }
            lo = hi;
            if (k + 1 < K) {
                hi = ray_point(g, u, __ldcs(t_in + (int64_t)(k + 2) * n_rays + rr));
                len = norm3(hi - lo);
            }
        }
        if (valid) store_result(out_wet, out_hydro, peers, r, acc_w, acc_h, accumulate);
    }
    // per-thread OOB counters -> warp sums -> three atomics per warp at most
    n_below = __reduce_add_sync(0xffffffffu, n_below);
    n_above = __reduce_add_sync(0xffffffffu, n_above);
    if (lane == 0) {
        if (n_first_below && !list) atomicAdd(counters + 0, (unsigned long long)n_first_below);
        if (n_below) atomicAdd(counters + 1, (unsigned long long)n_below);
        if (n_above) atomicAdd(counters + 2, (unsigned long long)n_above);
    }
}


// k3_poly.cuh -- K3, polynomial form: span cubics + closed-form layer sums -- the production integrator of thick layers.
// A fragment of libraider_b200.so: included by raider_b200.cu INSIDE its anonymous namespace, in this order (the translation unit
// is the one file it used to be; see the kernel inventory at the top of raider_b200.cu).  Not a stand-alone header.
// ------------------------------------------------------------------------------------------------
// K3 (polynomial form): the production integrator.  The ray is cut into *spans* of whole layers (host plan: greedy, span
// length <= RDR_K3_SPAN metres of the longest ray); per span the cube coordinates (uy, ux) and the height h are evaluated
// exactly at four points (three new ones, the first is the previous span's last) and carried as cubics in the normalised
// along-ray coordinate s (fastpath.cuh: < 2e-8 m in h, 5e-8 m horizontally for 8 km spans).  Every sample of delay.py:287-323
// is then 9 DFMA of geometry + cell lookup + 14 DFMA of trilinear value instead of a Bowring inversion and two arcsines:
// ~40 DP instructions per sample instead of ~100, and the model CRS (geographic or Lambert) only matters at the span nodes.
// Sample positions, step counts (nParts) and trapezoid weights are the reference's; flagged rays go to k_ray_integrate in
// list mode exactly as for k_ray_integrate_fast.
// Dynamic shared memory: LayerRec[K] | z nodes [nz] | 1/dz [nz-1] | span ends int[nspan].
// ------------------------------------------------------------------------------------------------
template <typename OUT, int BLOCK, int MINB, bool LCC, bool CACHE, bool FROM0>
__global__ void __launch_bounds__(BLOCK, MINB) k_ray_integrate_poly(const FastCube c, const RayGeom G, int64_t n_rays,
                                                              const double *__restrict__ t_in, const DevPlan *__restrict__ P,
                                                              const double *__restrict__ znodes, int nz, double zmin, OUT *__restrict__ out_wet,
                                                              OUT *__restrict__ out_hydro, int accumulate, const PeerOut peers,
                                                              unsigned long long *__restrict__ counters, int *__restrict__ fix_list, int quad,
                                                              int tile_map, double *__restrict__ part) {
    // layers [k0, K) / spans [sp0, nspan) of the device plan are this kernel's; a non-empty thin part [0, k0) is integrated by
    // k_ray_integrate_thin, which runs after this kernel and adds the partial sums left in `part`
    // Two instantiations are launched back to back and the plan picks one: FROM0 (no thin part: the whole ray, k0 = sp0 = 0 known at
    // compile time -- the C2-type case, where the registers the two variables would take are spills) or the upper part only.
    if (P->blocked) return;
    const int K = P->K, nspan = P->nspan;
    if (FROM0 != (P->k_split == 0)) return;
    const int k0 = FROM0 ? 0 : P->k_split, sp0 = FROM0 ? 0 : P->span_split;
    if (k0 >= K) return;
    const int clamp_low_first = P->clamp_low_first;
    extern __shared__ __align__(16) unsigned char fast_smem[];
    LayerRec *s_layers = reinterpret_cast<LayerRec *>(fast_smem);
    double *s_z = reinterpret_cast<double *>(fast_smem + (size_t)K * sizeof(LayerRec));
    double *s_inv = s_z + nz;
    int *s_span = reinterpret_cast<int *>(s_inv + (nz - 1));
    for (int i = threadIdx.x; i < K; i += BLOCK) s_layers[i] = P->layers[i];
    for (int i = threadIdx.x; i < nz; i += BLOCK) s_z[i] = znodes[i];
    for (int i = threadIdx.x; i < nz - 1; i += BLOCK) s_inv[i] = 1.0 / (znodes[i + 1] - znodes[i]);
    for (int i = threadIdx.x; i < nspan; i += BLOCK) s_span[i] = P->span_end[i];
    __syncthreads();
    const ZTable T = {s_z, s_inv, nz};
    const double ky = RAD_TO_DEG * c.y_inv, kx = RAD_TO_DEG * c.x_inv;
    const int64_t n_pad = (n_rays + 31) / 32 * 32;
    unsigned n_first_below = 0;
    for (int64_t q = blockIdx.x * (int64_t)BLOCK + threadIdx.x; q < n_pad; q += (int64_t)gridDim.x * BLOCK) {
        // tile_map: a warp takes a compact tile of the raster (8 x 4 pixels) instead of 32 pixels of one row.  The layers in which the rays of a
        // warp cross a horizontal cell face are summed sample by sample (per thread, the others wait): a compact tile crosses
        // a face within fewer layers than a 32-pixel row does.  (The along-ray distances are indexed by ray, not by thread.)
        int64_t r = q;
        if (tile_map) {  // tile_map = log2(tile width): 2^tile_map x 2^(5 - tile_map) pixels
            const int64_t tile = q >> 5, per_band = G.nx >> tile_map, band = tile / per_band;
            const int lane = (int)(q & 31);
            r = ((band << (5 - tile_map)) + (lane >> tile_map)) * G.nx + ((tile - band * per_band) << tile_map) + (lane & ((1 << tile_map) - 1));
        }
        const bool valid = r < n_rays;
        const int64_t rr = valid ? r : n_rays - 1;
        double lat, lon;
        ray_latlon(G, rr, lat, lon);
        RayFrame F;
        frame_setup(lat, lon, G.ht, G.los_kind, G.los, rr, G.e, G.n, G.u, F);
        const RayCell R = LCC ? ray_cell_lcc(c.lcc, F.slat, F.clat, lon) : RayCell{fma(lat, c.y_inv, c.y_c0), fma(lon, c.x_inv, c.x_c0), ky, kx};
        const double unorm = norm3(Vec3{F.uA, F.uB, F.uZ});  // |P_hi - P_lo| = |t_hi - t_lo| |u|  (losreader.py:821)
        bool bad = !F.fast_ok;   // (too close to the polar axis for the small-angle formulas: the PROJ-form kernel takes the ray)
        double acc_w = 0.0, acc_h = 0.0, vw, vh;
        double t_a = __ldcs(t_in + (int64_t)k0 * n_rays + rr), t_lo = t_a;
        // the along-ray distances stream from HBM: the top of the next layer and the end of the next span are requested one
        // layer / one span ahead of their use
        double t_next = __ldcs(t_in + (int64_t)(k0 + 1) * n_rays + rr);
        double tb_next = __ldcs(t_in + (int64_t)s_span[sp0] * n_rays + rr);
        RayNode n0 = node_eval<LCC>(c, F, R, t_a, bad);
        const bool clamp_first = (k0 == 0) && clamp_low_first;
        if (k0 == 0) {
            // very first sample of the ray (ff = 0 of the first layer); all pixels below min(z) -> clamp (delay.py:306-307)
            n_first_below += __popc(__ballot_sync(0xffffffffu, valid && (n0.h < zmin)));
        }
        sample_cell(c, s_layers[k0], T, n0.uy, n0.ux, clamp_first ? zmin : n0.h, vw, vh, bad);
        // CACHE: the 128-byte record of the cell the previous sample fell into stays in registers.  The samples of a layer share
        // their z cell and a ray crosses a horizontal cell face only every few km, so most samples reuse it: the gather drops
        // from 8 LDG.128 per sample (32 L1 wavefront cycles per warp: the limiter of the uncached kernel) to 8 per cell entered.
        // The cell is identified by a packed key (iy | ix << 10 | iz << 20; the host checks ny, nx <= 1024, nz <= 2048), so the
        // common case costs one compare; the address arithmetic and the loads only run when a new cell is entered.
        unsigned held = 0xffffffffu;
        CellData Q;
        Cubic py, px, ph;
        // horizontal cell (iy | ix << 10) and height of the last sample evaluated: the start of the next layer
        unsigned last_hkey;
        double last_h = clamp_first ? zmin : n0.h, last_ty, last_tx;
        {
            int iy0, ix0;
            last_ty = cell_coord_clamped(n0.uy, c.ny, iy0);
            last_tx = cell_coord_clamped(n0.ux, c.nx, ix0);
            last_hkey = (unsigned)iy0 | ((unsigned)ix0 << 10);
        }
        auto sample_cached = [&](const LayerRec &L, double s, double &w_out, double &h_out) {
            const double s2 = s * s;  // Estrin: two dependent levels after s instead of Horner's three
            const double uy = fma(s2, fma(s, py.c3, py.c2), fma(s, py.c1, py.c0));
            const double ux = fma(s2, fma(s, px.c3, px.c2), fma(s, px.c1, px.c0));
            const double h = fma(s2, fma(s, ph.c3, ph.c2), fma(s, ph.c1, ph.c0));
            int iy, ix, iz = L.iz;
            // (the span nodes keep NODE_MARGIN cells away from the cube's outer faces and the coordinates are monotone to well below
            // that margin in between, so the per-sample indices need clamping for memory safety only)
            const double ty = cell_coord_clamped(uy, c.ny, iy), tx = cell_coord_clamped(ux, c.nx, ix);
            double tz = fma(h, L.inv_dz, L.neg_zlo_inv);
            if (!(h >= L.h_lo && h < L.h_hi)) z_lookup(T, h, iz, tz, bad);
            const unsigned key = (unsigned)iy | ((unsigned)ix << 10) | ((unsigned)iz << 20);
            if (key != held) {
                Q = load_cell(c.cells + ((unsigned)(iy * (c.nx - 1) + ix) * (unsigned)c.nzc + (unsigned)iz));
                held = key;
            }
            eval_cell(Q, ty, tx, tz, w_out, h_out);
            last_hkey = key & 0xfffffu;
            last_h = h;
            last_ty = ty;
            last_tx = tx;
        };
        int k = k0;
        for (int sp = sp0; sp < nspan; ++sp) {
            const int k1 = s_span[sp];
            const double t_b = tb_next;
            if (sp + 1 < nspan) tb_next = __ldcs(t_in + (int64_t)s_span[sp + 1] * n_rays + rr);
            const double span = t_b - t_a;
            bad |= !(span > 0.0);
            const RayNode n1 = node_eval<LCC>(c, F, R, fma(span, 1.0 / 3.0, t_a), bad);
            const RayNode n2 = node_eval<LCC>(c, F, R, fma(span, 2.0 / 3.0, t_a), bad);
            const RayNode n3 = node_eval<LCC>(c, F, R, t_b, bad);
            py = cubic_through(n0.uy, n1.uy, n2.uy, n3.uy);
            px = cubic_through(n0.ux, n1.ux, n2.ux, n3.ux);
            ph = cubic_through(n0.h, n1.h, n2.h, n3.h);
            const double inv_span = rcp3(span);
            for (; k < k1; ++k) {
                const LayerRec L = s_layers[k];
                const double t_hi = t_next;
                if (k + 2 <= K) t_next = __ldcs(t_in + (int64_t)(k + 2) * n_rays + rr);
                const double dt = t_hi - t_lo;
                const double len = fabs(dt) * unorm;
                const double wt_full = (len * 1.0e-6) * L.step;   // delay.py:315 (L.step = RN(1 / (np - 1)): 1 ulp from the division)
                const double wt_half = 0.5 * wt_full;
                // sample j sits at t_lo + (j step) dt (delay.py:287,292), i.e. at s = s_lo + j (step ds) of the span
                const double s_lo = (t_lo - t_a) * inv_span, ds = dt * inv_span, sstep = L.step * ds;
                double fj = 1.0;
                int j = 1;
                bool layer_done = false;
                double end_w = 0.0, end_h = 0.0;
                if (CACHE && quad && L.np >= 4) {
                    // Layer quadrature.  Inside ONE cube cell the interpolant is a cubic p along the (straight) segment, up to the
                    // ~1e-5 curvature of the coordinates; for a cubic the composite trapezoid sum over n intervals is *exactly*
                    //     T_n[p] = (p(0) + 4 p(1/2) + p(1)) / 6 + (p(0) - 2 p(1/2) + p(1)) / (3 n^2)
                    // (Euler-Maclaurin stops after the h^2 term, p'(1) - p'(0) = 4 x the second central difference, Simpson is
                    // exact), so the n - 1 interior samples of delay.py:287-323 are replaced by the one in the middle of the layer:
                    // the sum the reference forms, to ~1e-15 m per layer (the quartic remainder).
                    // The layer's two END samples need not lie in the cell: Newton leaves the layer tops mm .. m off their nodes
                    // (losreader.py:720-733).  p(0), p(1) are then the cell's own polynomial continued to the end points, and
                    // the sum gets the two end corrections (f - p) / (2 n) with f the interpolant's value in the cell the end
                    // point really lies in -- exact as long as only the end samples are outside (LAYER_QUAD_TOL << sample spacing).
                    // A layer that crosses a horizontal cell face is summed sample by sample below.
                    const double sm = fma(0.5, ds, s_lo), se = s_lo + ds;
                    const double sm2 = sm * sm, se2 = se * se;
                    const double uym = fma(sm2, fma(sm, py.c3, py.c2), fma(sm, py.c1, py.c0)), uye = fma(se2, fma(se, py.c3, py.c2), fma(se, py.c1, py.c0));
                    const double uxm = fma(sm2, fma(sm, px.c3, px.c2), fma(sm, px.c1, px.c0)), uxe = fma(se2, fma(se, px.c3, px.c2), fma(se, px.c1, px.c0));
                    const double h_m = fma(sm2, fma(sm, ph.c3, ph.c2), fma(sm, ph.c1, ph.c0)), h_e = fma(se2, fma(se, ph.c3, ph.c2), fma(se, ph.c1, ph.c0));
                    int iym, ixm, iye, ixe;
                    const double tym = cell_coord_clamped(uym, c.ny, iym), txm = cell_coord_clamped(uxm, c.nx, ixm);
                    const double tye = cell_coord_clamped(uye, c.ny, iye), txe = cell_coord_clamped(uxe, c.nx, ixe);
                    const unsigned hkm = (unsigned)iym | ((unsigned)ixm << 10), hke = (unsigned)iye | ((unsigned)ixe << 10);
                    const double z_hi = T.z[L.iz + 1];
                    const bool top_cell = L.iz + 2 >= T.nz;  // nothing above: the end point must be inside (it is: zref < max(z))
                    const bool one_cell = (hkm == hke) & (hkm == last_hkey) & (last_h >= L.z_lo - LAYER_QUAD_TOL) & (h_m >= L.z_lo) & (h_m < z_hi) &
                                          (h_e >= L.z_lo) & (top_cell ? (h_e <= z_hi) : (h_e < z_hi + LAYER_QUAD_TOL));
                    if (one_cell) {
                        const unsigned key = hkm | ((unsigned)L.iz << 20);
                        if (key != held) {
                            Q = load_cell(c.cells + ((unsigned)(iym * (c.nx - 1) + ixm) * (unsigned)c.nzc + (unsigned)L.iz));
                            held = key;
                        }
                        double p0w = vw, p0h = vh, mw, mh, p1w, p1h;
                        const double tz0 = fma(last_h, L.inv_dz, L.neg_zlo_inv), tzm = fma(h_m, L.inv_dz, L.neg_zlo_inv), tze = fma(h_e, L.inv_dz, L.neg_zlo_inv);
                        if (last_h < L.z_lo) eval_cell(Q, last_ty, last_tx, tz0, p0w, p0h);  // start point below the cell
                        eval_cell(Q, tym, txm, tzm, mw, mh);
                        eval_cell(Q, tye, txe, tze, p1w, p1h);
                        end_w = p1w;
                        end_h = p1h;
                        // The one term beyond a cubic that matters: the fractions are quadratics b u + q u^2 (q ~ 1e-4: curvature of
                        // latitude / longitude / height along the chord), so the triple product a7 ty tx tz carries
                        // a7 (qy bx bz + by qx bz + by bx qz) u^4, and T_n[u^4] differs from the three-point formula by
                        // kappa_n = -1/120 + 1/(24 n^2) - 1/(30 n^4).  (1e-11 m per thick layer on a cube with O(1) mixed differences;
                        // everything of higher order is < 1e-13 m.)
                        const double qy = 2.0 * ((last_ty + tye) - 2.0 * tym), by = (tye - last_ty) - qy;
                        const double qx = 2.0 * ((last_tx + txe) - 2.0 * txm), bx = (txe - last_tx) - qx;
                        const double qz = 2.0 * ((tz0 + tze) - 2.0 * tzm), bz = (tze - tz0) - qz;
                        const double st2 = L.step * L.step;
                        const double g4 = fma(qy, bx * bz, by * fma(qx, bz, bx * qz)) * fma(st2, fma(st2, -1.0 / 30.0, 1.0 / 24.0), -1.0 / 120.0);
                        const double e4w = Q.q3.z * g4, e4h = Q.q3.w * g4;
                        if (!top_cell && h_e >= z_hi) {  // end point above the cell: its value in the cell it lies in (the next layer's)
                            Q = load_cell(c.cells + ((unsigned)(iym * (c.nx - 1) + ixm) * (unsigned)c.nzc + (unsigned)(L.iz + 1)));
                            held = hkm | ((unsigned)(L.iz + 1) << 20);
                            eval_cell(Q, tye, txe, (h_e - z_hi) * T.inv[L.iz + 1], end_w, end_h);
                        }
                        const double W = len * 1.0e-6, cn = st2 * (1.0 / 3.0), hn = 0.5 * L.step;
                        double tw = fma(fma(-2.0, mw, p0w + p1w), cn, fma(fma(4.0, mw, p0w + p1w), 1.0 / 6.0, e4w));
                        double th = fma(fma(-2.0, mh, p0h + p1h), cn, fma(fma(4.0, mh, p0h + p1h), 1.0 / 6.0, e4h));
                        tw = fma((vw - p0w) + (end_w - p1w), hn, tw);
                        th = fma((vh - p0h) + (end_h - p1h), hn, th);
                        acc_w = fma(W, tw, acc_w);
                        acc_h = fma(W, th, acc_h);
                        last_hkey = hke;
                        last_h = h_e;
                        last_ty = tye;
                        last_tx = txe;
                        vw = end_w;
                        vh = end_h;
                        layer_done = true;
                    }
                }
                if (!layer_done) {
                // first sample of this layer == last sample of the previous one (evaluated once, used with both end weights)
                acc_w = fma(wt_half, vw, acc_w);
                acc_h = fma(wt_half, vh, acc_h);
                if (CACHE) {
                    if (k + 2 < K && held != 0xffffffffu) {
                        // the record two layers up in the column the ray is in now: requested into L1 a layer or more before its first use
                        const LerpCell *nx2 = c.cells + ((unsigned)((int)(held & 1023u) * (c.nx - 1) + (int)((held >> 10) & 1023u)) * (unsigned)c.nzc +
                                                         (unsigned)s_layers[k + 2].iz);
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(nx2));
                    }
                    for (; j + 1 < L.np - 1; j += 2) {
                        // two interior samples as one straight-line block (two independent dependency chains: the loop is latency
                        // bound otherwise).  Both are taken to lie in the layer's own z cell and in one horizontal cell, which is
                        // the case for all but a few per ray; the exceptions are redone one at a time.
                        const double sa = fma(fj, sstep, s_lo), sb = fma(fj + 1.0, sstep, s_lo);
                        fj += 2.0;
                        const double sa2 = sa * sa, sb2 = sb * sb;
                        const double uya = fma(sa2, fma(sa, py.c3, py.c2), fma(sa, py.c1, py.c0)), uyb = fma(sb2, fma(sb, py.c3, py.c2), fma(sb, py.c1, py.c0));
                        const double uxa = fma(sa2, fma(sa, px.c3, px.c2), fma(sa, px.c1, px.c0)), uxb = fma(sb2, fma(sb, px.c3, px.c2), fma(sb, px.c1, px.c0));
                        const double h_a = fma(sa2, fma(sa, ph.c3, ph.c2), fma(sa, ph.c1, ph.c0)), h_b = fma(sb2, fma(sb, ph.c3, ph.c2), fma(sb, ph.c1, ph.c0));
                        int iya, ixa, iyb, ixb;
                        const double tya = cell_coord_clamped(uya, c.ny, iya), txa = cell_coord_clamped(uxa, c.nx, ixa);
                        const double tyb = cell_coord_clamped(uyb, c.ny, iyb), txb = cell_coord_clamped(uxb, c.nx, ixb);
                        const double tza = fma(h_a, L.inv_dz, L.neg_zlo_inv), tzb = fma(h_b, L.inv_dz, L.neg_zlo_inv);
                        const unsigned keya = (unsigned)iya | ((unsigned)ixa << 10) | ((unsigned)L.iz << 20);
                        const unsigned keyb = (unsigned)iyb | ((unsigned)ixb << 10) | ((unsigned)L.iz << 20);
                        const bool regular = (keya == keyb) & (h_a >= L.h_lo) & (h_a < L.h_hi) & (h_b >= L.h_lo) & (h_b < L.h_hi);
                        double wa, ha, wb, hb;
                        if (regular) {
                            if (keya != held) {
                                Q = load_cell(c.cells + ((unsigned)(iya * (c.nx - 1) + ixa) * (unsigned)c.nzc + (unsigned)L.iz));
                                held = keya;
                            }
                            eval_cell(Q, tya, txa, tza, wa, ha);
                            eval_cell(Q, tyb, txb, tzb, wb, hb);
                        } else {
                            sample_cached(L, sa, wa, ha);
                            sample_cached(L, sb, wb, hb);
                        }
                        acc_w = fma(wt_full, wa, acc_w);
                        acc_h = fma(wt_full, ha, acc_h);
                        acc_w = fma(wt_full, wb, acc_w);
                        acc_h = fma(wt_full, hb, acc_h);
                    }
                    if (j < L.np - 1) {
                        double wa, ha;
                        sample_cached(L, fma(fj, sstep, s_lo), wa, ha);
                        acc_w = fma(wt_full, wa, acc_w);
                        acc_h = fma(wt_full, ha, acc_h);
                    }
                    sample_cached(L, s_lo + ds, vw, vh);  // the layer's last sample (ff = 1)
                } else {
                    for (; j + 1 < L.np - 1; j += 2) {  // two interior samples per trip: independent chains for the FP64 pipe
                        const double sa = fma(fj, sstep, s_lo), sb = fma(fj + 1.0, sstep, s_lo);
                        fj += 2.0;
                        double wa, ha, wb, hb;
                        sample_cell(c, L, T, cubic_eval(py, sa), cubic_eval(px, sa), cubic_eval(ph, sa), wa, ha, bad);
                        sample_cell(c, L, T, cubic_eval(py, sb), cubic_eval(px, sb), cubic_eval(ph, sb), wb, hb, bad);
                        acc_w = fma(wt_full, wa, acc_w);
                        acc_h = fma(wt_full, ha, acc_h);
                        acc_w = fma(wt_full, wb, acc_w);
                        acc_h = fma(wt_full, hb, acc_h);
                    }
                    if (j < L.np - 1) {
                        const double sa = fma(fj, sstep, s_lo);
                        double wa, ha;
                        sample_cell(c, L, T, cubic_eval(py, sa), cubic_eval(px, sa), cubic_eval(ph, sa), wa, ha, bad);
                        acc_w = fma(wt_full, wa, acc_w);
                        acc_h = fma(wt_full, ha, acc_h);
                    }
                    {   // the layer's last sample (ff = 1)
                        const double se = s_lo + ds;
                        sample_cell(c, L, T, cubic_eval(py, se), cubic_eval(px, se), cubic_eval(ph, se), vw, vh, bad);
                    }
                }
                acc_w = fma(wt_half, vw, acc_w);
                acc_h = fma(wt_half, vh, acc_h);
                }  // !layer_done
                t_lo = t_hi;
            }
            t_a = t_b;
            n0 = n3;
        }
        if (valid) {
            if (bad) {
                fix_list[atomicAdd(counters + 3, 1ull)] = (int)r;
                if (k0 > 0) __stcs(part + r, __longlong_as_double(PART_FLAGGED));  // the ray is on the fix list: the thin kernel leaves it alone
            } else if (k0 > 0) {  // the thin-layer kernel finishes the ray
                __stcs(part + r, acc_w);
                __stcs(part + n_rays + r, acc_h);
            } else {
                store_result(out_wet, out_hydro, peers, r, acc_w, acc_h, accumulate);
            }
        }
    }
    if ((threadIdx.x & 31) == 0 && n_first_below) atomicAdd(counters + 0, (unsigned long long)n_first_below);
}


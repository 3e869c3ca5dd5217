// k3_thin.cuh -- K3, thin-layer form: TMA-staged cell records + cp.async distance ring -- the production integrator of the 145-node tables.
// A fragment of libraider_b200.so: included by raider_b200.cu INSIDE its anonymous namespace, in this order (the translation unit
// is the one file it used to be; see the kernel inventory at the top of raider_b200.cu).  Not a stand-alone header.
// ------------------------------------------------------------------------------------------------
// K3 (thin-layer form): layers [0, k_split) of the device plan -- the run of layers with <= 3 samples each that the real
// processed cubes consist of below ~20 km (145-node tables at the reference's 1000 m segments: ~115 of 139 layers hold 2
// samples, models/model_levels.py:12, delay.py:283).  There the quadrature kernel has nothing to sum in closed form and its
// per-layer bookkeeping is the cost; this kernel is the lean loop: one sample per layer (the layer top; the interface sample is
// shared by both layers as everywhere), geometry from the span cubics, and a cell lookup that costs two subtractions and two
// integer compares while the ray stays in its horizontal cell (the floor values are held; a cell is ~25 km wide, a thin layer
// moves the ray ~100 m).  Every sample needs a new 128-byte record (the z cell changes with every layer), so the records of the
// layers ahead are requested into L1 `pf_cells` layers early (they are consecutive lines: z is the fastest axis of the record
// array) and the along-ray distances into L2 `pf_t` layers early -- the uncached polynomial kernel spent 53 % of its stall
// samples on the long scoreboard here (profiles/r01f ml145).  Sample positions, step counts and weights are the reference's.
// Runs after k_ray_integrate_poly (which leaves the partial sums of layers [k_split, K) in `part`) and stores the results.
// ------------------------------------------------------------------------------------------------
// both fields of a cell record held in shared memory (staged columns), by shared-space address: 8 LDS.128
__device__ __forceinline__ void trilinear_cell_s(uint32_t rec, double ty, double tx, double tz, double &vw, double &vh) {
    double q[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q[2 * i]), "=d"(q[2 * i + 1]) : "r"(rec + 16u * i));
    // record layout (LerpCell): q0 = {w a0, h a0, w a1, h a1}, q1 = {a2, a3}, q2 = {a4, a5}, q3 = {a6, a7}
    const double w0 = fma(tz, q[2], q[0]), h0 = fma(tz, q[3], q[1]);
    const double w1 = fma(tz, q[6], q[4]), h1 = fma(tz, q[7], q[5]);
    const double w2 = fma(tz, q[10], q[8]), h2 = fma(tz, q[11], q[9]);
    const double w3 = fma(tz, q[14], q[12]), h3 = fma(tz, q[15], q[13]);
    vw = fma(ty, fma(tx, w3, w2), fma(tx, w1, w0));
    vh = fma(ty, fma(tx, h3, h2), fma(tx, h1, h0));
}

// STAGE: the north_star form -- the CTA's footprint of the cube is staged in shared memory by the TMA engine, span by span.
// Within one span of the polynomial geometry (<= 24 km of ray) the 128 rays of a CTA pass (a 32 x 4 pixel tile, ~3 km wide)
// drift a few km: they sit in 1-4 horizontal cells of a 0.25 deg cube, ~6 of a 3 km one.  The bounding box of those cells comes
// for free from the span's end nodes (which the cubics need anyway); the record columns of the box, restricted to the z cells of
// the span's layers -- contiguous in memory, z fastest -- are copied with one cp.async.bulk each (UBLKCP) onto an mbarrier, as
// long as they fit the `rec_cap` records of shared memory left beside MINB CTAs per SM.  A sample then reads its record with
// 8 LDS.128 at ~30 cycles instead of 8 LDG.128 from L2 at ~600 under load (profiles/r02a: 59 % of the stall samples of the
// unstaged kernel sit on the first use of those loads).  Samples whose cell is not staged (box too large for the capacity, z cell
// off the span's range) read the record from global memory as before.
template <typename OUT, int BLOCK, int MINB, bool LCC, bool STAGE, bool QUAD>
__global__ void __launch_bounds__(BLOCK, MINB) k_ray_integrate_thin(const FastCube c, const RayGeom G, int64_t n_rays,
                                                              const double *__restrict__ t_in, const DevPlan *__restrict__ P,
                                                              const double *__restrict__ znodes, int nz, double zmin, OUT *__restrict__ out_wet,
                                                              OUT *__restrict__ out_hydro, int accumulate, const PeerOut peers,
                                                              unsigned long long *__restrict__ counters, int *__restrict__ fix_list, int tile_map,
                                                              const double *__restrict__ part, int pf_cells, int quad, int rec_cap,
                                                              unsigned long long *__restrict__ stage_stats) {
    if (P->blocked) return;
    const int K = P->K, k_end = P->k_split, nspan = P->span_split;
    if (k_end == 0) return;
    const int clamp_low_first = P->clamp_low_first;
    extern __shared__ __align__(128) unsigned char fast_smem[];
    LayerRec *s_layers = reinterpret_cast<LayerRec *>(fast_smem);
    double *s_z = reinterpret_cast<double *>(fast_smem + (size_t)K * sizeof(LayerRec));
    double *s_inv = s_z + nz;
    int *s_span = reinterpret_cast<int *>(s_inv + (nz - 1));
    // staging area: mbarrier | per-warp bounding boxes | record columns (128-byte aligned)
    const size_t stage_off = (((size_t)K * sizeof(LayerRec) + (2 * (size_t)nz - 1) * sizeof(double) + (size_t)K * sizeof(int)) + 127) / 128 * 128;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(fast_smem + stage_off);
    int *s_wbox = reinterpret_cast<int *>(fast_smem + stage_off + 16);            // [BLOCK / 32][4]
    // ring of along-ray distances: THIN_TD rows in flight per thread (cp.async), slot d of thread i at [d][i]
    const uint32_t s_ring = smem_u32(fast_smem + stage_off + 128) + 8u * threadIdx.x;
    LerpCell *s_cols = reinterpret_cast<LerpCell *>(fast_smem + stage_off + 128 + THIN_TD * BLOCK * sizeof(double));
    const uint32_t s_cols_addr = smem_u32(s_cols);
    for (int i = threadIdx.x; i < k_end; i += BLOCK) s_layers[i] = P->layers[i];
    for (int i = threadIdx.x; i < nz; i += BLOCK) s_z[i] = znodes[i];
    for (int i = threadIdx.x; i < nz - 1; i += BLOCK) s_inv[i] = 1.0 / (znodes[i + 1] - znodes[i]);
    for (int i = threadIdx.x; i < nspan; i += BLOCK) s_span[i] = P->span_end[i];
    if (STAGE && threadIdx.x == 0) mbar_init(s_bar, 1);
    __syncthreads();
    const ZTable T = {s_z, s_inv, nz};
    const double ky = RAD_TO_DEG * c.y_inv, kx = RAD_TO_DEG * c.x_inv;
    const int64_t n_pad = (n_rays + BLOCK - 1) / BLOCK * BLOCK;  // whole CTAs walk the loop together (barriers inside)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned n_first_below = 0, phase = 0;
    unsigned long long n_staged = 0, n_unstaged = 0;
    for (int64_t q0 = blockIdx.x * (int64_t)BLOCK; q0 < n_pad; q0 += (int64_t)gridDim.x * BLOCK) {
        const int64_t q = q0 + threadIdx.x;
        int64_t r = q;
        if (tile_map) {  // as in k_ray_integrate_poly
            const int64_t tile = q >> 5, per_band = G.nx >> tile_map, band = tile / per_band;
            r = ((band << (5 - tile_map)) + (lane >> tile_map)) * G.nx + ((tile - band * per_band) << tile_map) + (lane & ((1 << tile_map) - 1));
        }
        const bool valid = r < n_rays;
        const int64_t rr = valid ? r : n_rays - 1;
        double lat, lon;
        ray_latlon(G, rr, lat, lon);
        RayFrame F;
        frame_setup(lat, lon, G.ht, G.los_kind, G.los, rr, G.e, G.n, G.u, F);
        const RayCell R = LCC ? ray_cell_lcc(c.lcc, F.slat, F.clat, lon) : RayCell{fma(lat, c.y_inv, c.y_c0), fma(lon, c.x_inv, c.x_c0), ky, kx};
        const double u6 = norm3(Vec3{F.uA, F.uB, F.uZ}) * 1.0e-6;  // |P_hi - P_lo| 1e-6 = |t_hi - t_lo| |u| 1e-6  (losreader.py:821, delay.py:315)
        bool bad = !F.fast_ok;   // (too close to the polar axis for the small-angle formulas: the PROJ-form kernel takes the ray)
        double acc_w = 0.0, acc_h = 0.0, vw, vh;
        // the along-ray distances stream from HBM: a thin layer is ~200 cycles of work, a load from HBM takes 600 .. 900, so the
        // rows k + 1 .. k + THIN_TD are kept in flight as asynchronous copies (LDGSTS) into a per-thread ring in shared memory --
        // registers would have to be rotated by moves, and a move waits for its load
        const double *tp = t_in + rr;  // row k of the distances: bottom of layer k
        double t_a = __ldcs(tp), t_lo = t_a;
#pragma unroll
        for (int d = 0; d < THIN_TD; ++d) {
            tp += n_rays;  // row d + 1
            if (d + 1 <= K) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s_ring + (uint32_t)(d * BLOCK * 8)), "l"(tp) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        uint32_t slot = 0;  // ring slot of row k + 1
        double tb_next = __ldcs(t_in + (int64_t)s_span[0] * n_rays + rr);
        RayNode n0 = node_eval<LCC>(c, F, R, t_a, bad);
        // very first sample of the ray (ff = 0 of the first layer); all pixels below min(z) -> clamp (delay.py:306-307)
        n_first_below += __popc(__ballot_sync(0xffffffffu, valid && (n0.h < zmin)));
        sample_cell(c, s_layers[0], T, n0.uy, n0.ux, clamp_low_first ? zmin : n0.h, vw, vh, bad);
        // the horizontal cell the ray is in: index, floor values of the cell coordinates, record column in global memory and -- when the
        // cell is inside the staged box -- in shared memory (shared-space address of its record for z cell 0)
        int iy, ix, bx0 = 0, by0 = 0, nbx = 0, nby = 0, lev0 = 0, nlev = 0;   // staged box: origin, extent (cells), first z cell, z cells
        double fy, fx;
        const LerpCell *col;
        uint32_t col_s = 0;
        bool in_smem = false;
        auto enter_cell = [&](double uy, double ux) {
            const double sy = __dadd_rd(uy, c_fast.floor_magic), sx = __dadd_rd(ux, c_fast.floor_magic);
            iy = min(max(__double2loint(sy), 0), c.ny - 2);
            ix = min(max(__double2loint(sx), 0), c.nx - 2);
            fy = sy - c_fast.floor_magic;
            fx = sx - c_fast.floor_magic;
            col = c.cells + (unsigned)(iy * (c.nx - 1) + ix) * (unsigned)c.nzc;
            if (STAGE) {
                const int cy = iy - by0, cx = ix - bx0;
                in_smem = ((unsigned)cy < (unsigned)nby) & ((unsigned)cx < (unsigned)nbx);
                col_s = s_cols_addr + (uint32_t)(((cy * nbx + cx) * nlev - lev0) * (int)sizeof(LerpCell));
            }
        };
        enter_cell(n0.uy, n0.ux);
        // the last sample evaluated (the start of the next layer): height and fractions in the held horizontal cell
        double last_h = clamp_low_first ? zmin : n0.h, last_ty = n0.uy - fy, last_tx = n0.ux - fx;
        Cubic py, px, ph;
        auto sample = [&](const LayerRec &L, double s, double &w_out, double &h_out) {
            const double s2 = s * s;  // Estrin, as in k_ray_integrate_poly (same rounding)
            const double uy = fma(s2, fma(s, py.c3, py.c2), fma(s, py.c1, py.c0));
            const double ux = fma(s2, fma(s, px.c3, px.c2), fma(s, px.c1, px.c0));
            const double h = fma(s2, fma(s, ph.c3, ph.c2), fma(s, ph.c1, ph.c0));
            double ty = uy - fy, tx = ux - fx;
            // 0 <= t < 1  <=>  the high word of t, as an unsigned, is below that of 1.0 (negative and NaN have larger high words)
            if (((unsigned)__double2hiint(ty) >= 0x3ff00000u) | ((unsigned)__double2hiint(tx) >= 0x3ff00000u)) {
                enter_cell(uy, ux);
                ty = uy - fy;
                tx = ux - fx;
            }
            int iz = L.iz;
            double tz = fma(h, L.inv_dz, L.neg_zlo_inv);
            const bool own_cell = (h >= L.h_lo) & (h < L.h_hi);
            last_h = h;
            last_ty = ty;
            last_tx = tx;
            if (STAGE && in_smem && own_cell) {
                trilinear_cell_s(col_s + (uint32_t)iz * (uint32_t)sizeof(LerpCell), ty, tx, tz, w_out, h_out);
            } else {
                if (!own_cell) z_lookup(T, h, iz, tz, bad);  // (rare) not in the layer's own cell
                const LerpCell *rec = col + iz;
                if (!STAGE && pf_cells) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + pf_cells));  // (the record array is padded at its end)
                trilinear_cell(rec, ty, tx, tz, w_out, h_out);
            }
        };
        int k = 0;
        for (int sp = 0; sp < nspan; ++sp) {
            const int k1 = s_span[sp];
            const double t_b = tb_next;
            if (sp + 1 < nspan) tb_next = __ldcs(t_in + (int64_t)s_span[sp + 1] * n_rays + rr);
            const double span = t_b - t_a;
            bad |= !(span > 0.0);
            const RayNode n1 = node_eval<LCC>(c, F, R, fma(span, 1.0 / 3.0, t_a), bad);
            const RayNode n2 = node_eval<LCC>(c, F, R, fma(span, 2.0 / 3.0, t_a), bad);
            const RayNode n3 = node_eval<LCC>(c, F, R, t_b, bad);
            if (STAGE) {
                // ---- footprint of this span: bounding box of the horizontal cells at its two ends, over the CTA -> staged columns
                int ia, ib, ja, jb;
                (void)cell_coord_clamped(n0.uy, c.ny, ia);
                (void)cell_coord_clamped(n3.uy, c.ny, ib);
                (void)cell_coord_clamped(n0.ux, c.nx, ja);
                (void)cell_coord_clamped(n3.ux, c.nx, jb);
                const int w_ylo = __reduce_min_sync(0xffffffffu, min(ia, ib)), w_yhi = __reduce_max_sync(0xffffffffu, max(ia, ib));
                const int w_xlo = __reduce_min_sync(0xffffffffu, min(ja, jb)), w_xhi = __reduce_max_sync(0xffffffffu, max(ja, jb));
                __syncthreads();  // the CTA is done with the boxes and the columns of the previous span
                if (lane == 0) {
                    s_wbox[4 * warp] = w_ylo; s_wbox[4 * warp + 1] = w_yhi; s_wbox[4 * warp + 2] = w_xlo; s_wbox[4 * warp + 3] = w_xhi;
                }
                __syncthreads();
                int ylo = s_wbox[0], yhi = s_wbox[1], xlo = s_wbox[2], xhi = s_wbox[3];
#pragma unroll
                for (int w = 1; w < BLOCK / 32; ++w) {
                    ylo = min(ylo, s_wbox[4 * w]); yhi = max(yhi, s_wbox[4 * w + 1]);
                    xlo = min(xlo, s_wbox[4 * w + 2]); xhi = max(xhi, s_wbox[4 * w + 3]);
                }
                // z cells of the span's layers and one neighbour each way (layer tops sit mm .. m off their nodes)
                lev0 = max(s_layers[k].iz - 1, 0);
                nlev = min(s_layers[k1 - 1].iz + 1, c.nzc - 1) - lev0 + 1;
                const int ncols = (yhi - ylo + 1) * (xhi - xlo + 1);
                if (ncols * nlev <= rec_cap) {  // (CTA-uniform)
                    by0 = ylo; bx0 = xlo; nby = yhi - ylo + 1; nbx = xhi - xlo + 1;
                    const uint32_t col_bytes = (uint32_t)nlev * (uint32_t)sizeof(LerpCell);
                    if (threadIdx.x == 0) {
                        mbar_expect_tx(s_bar, (uint32_t)ncols * col_bytes);
                        for (int j = 0; j < ncols; ++j) {
                            const int cy = by0 + j / nbx, cx = bx0 + j % nbx;
                            tma_load_1d(s_cols + (size_t)j * nlev, c.cells + ((size_t)(cy * (c.nx - 1) + cx) * c.nzc + lev0), col_bytes, s_bar);
                        }
                    }
                    mbar_wait(s_bar, phase);
                    phase ^= 1u;
                    n_staged += threadIdx.x == 0;
                } else {
                    nby = nbx = 0;
                    n_unstaged += threadIdx.x == 0;
                }
                enter_cell(n0.uy, n0.ux);  // (the staging changed: refresh the column addresses of the cell the ray is in)
            }
            py = cubic_through(n0.uy, n1.uy, n2.uy, n3.uy);
            px = cubic_through(n0.ux, n1.ux, n2.ux, n3.ux);
            ph = cubic_through(n0.h, n1.h, n2.h, n3.h);
            const double inv_span = rcp3(span);
            for (; k < k1; ++k) {
                const LayerRec L = s_layers[k];
                double t_hi;
                asm volatile("cp.async.wait_group %0;" ::"n"(THIN_TD - 1) : "memory");  // row k + 1 has landed
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t_hi) : "r"(s_ring + slot * (uint32_t)(BLOCK * 8)) : "memory");
                tp += n_rays;  // row k + 1 + THIN_TD goes into the slot just read
                if (k + 1 + THIN_TD <= K) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s_ring + slot * (uint32_t)(BLOCK * 8)), "l"(tp) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                slot = (slot + 1) & (THIN_TD - 1);
                const double dt = t_hi - t_lo;
                const double wt_full = (fabs(dt) * u6) * L.step;  // delay.py:315
                const double wt_half = 0.5 * wt_full;
                double ew, eh;
                bool layer_done = false;
                if (STAGE && QUAD && quad && L.np >= 4 && in_smem) {
                    // Layer quadrature exactly as in k_ray_integrate_poly (see there): the composite trapezoid sum of a layer whose
                    // samples share one cube cell, in closed form from the values at its start, middle and end -- here with the
                    // cell's record read from the staged column in shared memory instead of a register-held copy.
                    const double s_lo = (t_lo - t_a) * inv_span, ds = dt * inv_span;
                    const double sm = fma(0.5, ds, s_lo), se = s_lo + ds;
                    const double sm2 = sm * sm, se2 = se * se;
                    const double uym = fma(sm2, fma(sm, py.c3, py.c2), fma(sm, py.c1, py.c0)), uye = fma(se2, fma(se, py.c3, py.c2), fma(se, py.c1, py.c0));
                    const double uxm = fma(sm2, fma(sm, px.c3, px.c2), fma(sm, px.c1, px.c0)), uxe = fma(se2, fma(se, px.c3, px.c2), fma(se, px.c1, px.c0));
                    const double h_m = fma(sm2, fma(sm, ph.c3, ph.c2), fma(sm, ph.c1, ph.c0)), h_e = fma(se2, fma(se, ph.c3, ph.c2), fma(se, ph.c1, ph.c0));
                    const double tym = uym - fy, txm = uxm - fx, tye = uye - fy, txe = uxe - fx;
                    const unsigned hi_max = max(max((unsigned)__double2hiint(tym), (unsigned)__double2hiint(txm)),
                                                max((unsigned)__double2hiint(tye), (unsigned)__double2hiint(txe)));
                    const double z_hi = T.z[L.iz + 1];
                    const bool top_cell = L.iz + 2 >= T.nz;  // nothing above: the end point must be inside (it is: zref < max(z))
                    const bool one_cell = (hi_max < 0x3ff00000u) & (last_h >= L.z_lo - LAYER_QUAD_TOL) & (h_m >= L.z_lo) & (h_m < z_hi) &
                                          (h_e >= L.z_lo) & (top_cell ? (h_e <= z_hi) : (h_e < z_hi + LAYER_QUAD_TOL));
                    if (one_cell) {
                        const uint32_t rec = col_s + (uint32_t)L.iz * (uint32_t)sizeof(LerpCell);
                        double p0w = vw, p0h = vh, mw, mh, p1w, p1h;
                        const double tz0 = fma(last_h, L.inv_dz, L.neg_zlo_inv), tzm = fma(h_m, L.inv_dz, L.neg_zlo_inv), tze = fma(h_e, L.inv_dz, L.neg_zlo_inv);
                        if (last_h < L.z_lo) trilinear_cell_s(rec, last_ty, last_tx, tz0, p0w, p0h);  // start point below the cell
                        trilinear_cell_s(rec, tym, txm, tzm, mw, mh);
                        trilinear_cell_s(rec, tye, txe, tze, p1w, p1h);
                        ew = p1w;
                        eh = p1h;
                        // the quartic term of the curved chord (see k_ray_integrate_poly): a7 (qy bx bz + by qx bz + by bx qz) kappa_n
                        const double qy = 2.0 * ((last_ty + tye) - 2.0 * tym), by = (tye - last_ty) - qy;
                        const double qx = 2.0 * ((last_tx + txe) - 2.0 * txm), bx = (txe - last_tx) - qx;
                        const double qz = 2.0 * ((tz0 + tze) - 2.0 * tzm), bz = (tze - tz0) - qz;
                        const double st2 = L.step * L.step;
                        const double g4 = fma(qy, bx * bz, by * fma(qx, bz, bx * qz)) * fma(st2, fma(st2, -1.0 / 30.0, 1.0 / 24.0), -1.0 / 120.0);
                        double a7w, a7h;
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a7w), "=d"(a7h) : "r"(rec + 112u));  // q3.z, q3.w
                        const double e4w = a7w * g4, e4h = a7h * g4;
                        if (!top_cell && h_e >= z_hi)  // end point above the cell: its value in the cell it lies in (the next layer's; staged: lev0 .. + 1)
                            trilinear_cell_s(rec + (uint32_t)sizeof(LerpCell), tye, txe, (h_e - z_hi) * T.inv[L.iz + 1], ew, eh);
                        const double W = (fabs(dt) * u6), cn = st2 * (1.0 / 3.0), hn = 0.5 * L.step;
                        double tw = fma(fma(-2.0, mw, p0w + p1w), cn, fma(fma(4.0, mw, p0w + p1w), 1.0 / 6.0, e4w));
                        double th = fma(fma(-2.0, mh, p0h + p1h), cn, fma(fma(4.0, mh, p0h + p1h), 1.0 / 6.0, e4h));
                        tw = fma((vw - p0w) + (ew - p1w), hn, tw);
                        th = fma((vh - p0h) + (eh - p1h), hn, th);
                        acc_w = fma(W, tw, acc_w);
                        acc_h = fma(W, th, acc_h);
                        last_h = h_e;
                        last_ty = tye;
                        last_tx = txe;
                        layer_done = true;
                    }
                }
                if (layer_done) {
                } else if (L.np == 2) {
                    // one interval: 0.5 w (f(lo) + f(hi)); the sample at the layer top sits at t_hi
                    sample(L, (t_hi - t_a) * inv_span, ew, eh);
                    acc_w = fma(wt_half, vw + ew, acc_w);
                    acc_h = fma(wt_half, vh + eh, acc_h);
                } else {
                    const double s_lo = (t_lo - t_a) * inv_span, ds = dt * inv_span, sstep = L.step * ds;
                    acc_w = fma(wt_half, vw, acc_w);
                    acc_h = fma(wt_half, vh, acc_h);
                    double fj = 1.0;
                    for (int j = 1; j < L.np - 1; ++j, fj += 1.0) {
                        double wa, ha;
                        sample(L, fma(fj, sstep, s_lo), wa, ha);
                        acc_w = fma(wt_full, wa, acc_w);
                        acc_h = fma(wt_full, ha, acc_h);
                    }
                    sample(L, s_lo + ds, ew, eh);  // the layer's last sample (ff = 1)
                    acc_w = fma(wt_half, ew, acc_w);
                    acc_h = fma(wt_half, eh, acc_h);
                }
                vw = ew;
                vh = eh;
                t_lo = t_hi;
            }
            t_a = t_b;
            n0 = n3;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");  // (rows beyond the thin part that were still in flight)
        if (valid) {
            const double pw = k_end < K ? __ldcs(part + r) : 0.0;  // the layers above were summed by k_ray_integrate_poly
            if (k_end < K && __double_as_longlong(pw) == PART_FLAGGED) {
                // already on the fix list
            } else if (bad) {
                fix_list[atomicAdd(counters + 3, 1ull)] = (int)r;
            } else {
                if (k_end < K) {
                    acc_w += pw;
                    acc_h += __ldcs(part + n_rays + r);
                }
                store_result(out_wet, out_hydro, peers, r, acc_w, acc_h, accumulate);
            }
        }
    }
    if (lane == 0 && n_first_below) atomicAdd(counters + 0, (unsigned long long)n_first_below);
    if (STAGE && stage_stats && threadIdx.x == 0) {
        if (n_staged) atomicAdd(stage_stats, n_staged);
        if (n_unstaged) atomicAdd(stage_stats + 1, n_unstaged);
    }
}


// k3_fast.cuh -- K3, per-sample Bowring form (tests / comparisons).
// A fragment of libraider_b200.so: included by raider_b200.cu INSIDE its anonymous namespace, in this order (the translation unit
// is the one file it used to be; see the kernel inventory at the top of raider_b200.cu).  Not a stand-alone header.
// ------------------------------------------------------------------------------------------------
// K3 (fast form): the same integral with the per-sample arithmetic of fastpath.cuh -- meridian-frame geometry, cubic-step
// reciprocal square roots, small-angle latitude / longitude differences, floor-by-rounding cell lookup on uniform horizontal
// axes, trilinear value in lerp form on {f[z], f[z+1]-f[z]} cells: ~100 DP instructions per sample instead of ~200.
// It integrates what it can prove regular and *flags* every other ray (polar, outside the small-angle window, leaving the
// cube, on the last node) into `fix_list`; k_ray_integrate re-does exactly those rays in list mode, with all the NaN rules.
// Dynamic shared memory: LayerRec[K] | z nodes [nz] | 1/dz [nz-1].
// ------------------------------------------------------------------------------------------------
template <typename OUT, int BLOCK, int MINB, int NPT>
__global__ void __launch_bounds__(BLOCK, MINB) k_ray_integrate_fast(const FastCube c, const RayGeom G, int64_t n_rays, int K,
                                                              const double *__restrict__ t_in, const DevPlan *__restrict__ P,
                                                              const double *__restrict__ znodes, int nz, double zmin,
                                                              OUT *__restrict__ out_wet, OUT *__restrict__ out_hydro, int accumulate, const PeerOut peers,
                                                              unsigned long long *__restrict__ counters, int *__restrict__ fix_list) {
    if (P->blocked) return;
    const LayerRec *__restrict__ layers = P->layers;
    const int clamp_low_first = P->clamp_low_first;
    extern __shared__ __align__(16) unsigned char fast_smem[];
    LayerRec *s_layers = reinterpret_cast<LayerRec *>(fast_smem);
    double *s_z = reinterpret_cast<double *>(fast_smem + (size_t)K * sizeof(LayerRec));
    double *s_inv = s_z + nz;
    for (int i = threadIdx.x; i < K; i += BLOCK) s_layers[i] = layers[i];
    for (int i = threadIdx.x; i < nz; i += BLOCK) s_z[i] = znodes[i];
    for (int i = threadIdx.x; i < nz - 1; i += BLOCK) s_inv[i] = 1.0 / (znodes[i + 1] - znodes[i]);
    __syncthreads();
    const ZTable T = {s_z, s_inv, nz};
    const double ky = RAD_TO_DEG * c.y_inv, kx = RAD_TO_DEG * c.x_inv;
    const int64_t n_pad = (n_rays + 31) / 32 * 32;
    unsigned n_first_below = 0;
    for (int64_t r = blockIdx.x * (int64_t)BLOCK + threadIdx.x; r < n_pad; r += (int64_t)gridDim.x * BLOCK) {
        const bool valid = r < n_rays;
        const int64_t rr = valid ? r : n_rays - 1;
        double lat, lon;
        ray_latlon(G, rr, lat, lon);
        RayFrame F;
        frame_setup(lat, lon, G.ht, G.los_kind, G.los, rr, G.e, G.n, G.u, F);
        const RayCell R = {fma(lat, c.y_inv, c.y_c0), fma(lon, c.x_inv, c.x_c0), ky, kx};
        const double unorm = norm3(Vec3{F.uA, F.uB, F.uZ});  // |P_hi - P_lo| = |t_hi - t_lo| |u|  (losreader.py:821)
        bool bad = !F.fast_ok;
        double acc_w = 0.0, acc_h = 0.0, vw, vh, h;
        // a sample is the point g + t u of the frame, t = t_lo + ff (t_hi - t_lo): the reference's low + ff (high - low) (delay.py:292)
        auto sample_at = [&](const LayerRec &L, double t, bool clamp, double &w_out, double &h_out) {
            sample_fast(c, F, R, L, T, fma(t, F.uA, F.A0), t * F.uB, fma(t, F.uZ, F.Z0), clamp, zmin, h, w_out, h_out, bad);
        };
        double t_lo = __ldcs(t_in + rr);
        // very first sample of the ray (ff = 0 of the first layer); all pixels below min(z) -> clamp (delay.py:306-307)
        sample_at(s_layers[0], t_lo, clamp_low_first != 0, vw, vh);
        n_first_below += __popc(__ballot_sync(0xffffffffu, valid && (h < zmin)));
        for (int k = 0; k < K; ++k) {
            const LayerRec L = s_layers[k];
            const double t_hi = __ldcs(t_in + (int64_t)(k + 1) * n_rays + rr);
            const double dt = t_hi - t_lo;
            const double len = fabs(dt) * unorm;
            const double wt_full = (len * 1.0e-6) / ((double)L.np - 1.0);   // delay.py:315
            const double wt_half = 0.5 * wt_full;
            // first sample of this layer == last sample of the previous one (evaluated once, used with both end weights)
            acc_w = fma(wt_half, vw, acc_w);
            acc_h = fma(wt_half, vh, acc_h);
            int j = 1;
            if (NPT == 2) {
                for (; j + 1 < L.np - 1; j += 2) {  // two interior samples per trip: independent chains for the FP64 pipe
                    double wa, ha, wb, hb;
                    sample_at(L, fma((double)j * L.step, dt, t_lo), false, wa, ha);
                    sample_at(L, fma((double)(j + 1) * L.step, dt, t_lo), false, wb, hb);
                    acc_w = fma(wt_full, wa, acc_w);
                    acc_h = fma(wt_full, ha, acc_h);
                    acc_w = fma(wt_full, wb, acc_w);
                    acc_h = fma(wt_full, hb, acc_h);
                }
            }
            for (; j < L.np - 1; ++j) {
                double wa, ha;
                sample_at(L, fma((double)j * L.step, dt, t_lo), false, wa, ha);
                acc_w = fma(wt_full, wa, acc_w);
                acc_h = fma(wt_full, ha, acc_h);
            }
            sample_at(L, t_hi, false, vw, vh);  // the layer's last sample (ff = 1)
            acc_w = fma(wt_half, vw, acc_w);
            acc_h = fma(wt_half, vh, acc_h);
            t_lo = t_hi;
        }
        if (valid) {
            if (bad) {
                fix_list[atomicAdd(counters + 3, 1ull)] = (int)r;
            } else {
                store_result(out_wet, out_hydro, peers, r, acc_w, acc_h, accumulate);
            }
        }
    }
    if ((threadIdx.x & 31) == 0 && n_first_below) atomicAdd(counters + 0, (unsigned long long)n_first_below);
}


/*
 * raider_b200.h -- C ABI of libraider_b200.so: the B200 (sm_100a) slant / zenith tropospheric-delay hot path.
 *
 * This is the drop-in boundary.  The reference (dbekaert/RAiDER @ e38c4eb4) has no C ABI of its own for this
 * path: it exposes CPython extension modules (setup.py:20-40) and Python functions.  Each entry point below
 * therefore names the reference *Python-level* interface it sits under (path:line relative to the reference
 * root); the Python shims in raider_b200/ keep those names/signatures and call these symbols through ctypes.
 * INTEGRATION.md shows the binding a RAiDER maintainer would add.
 *
 * Conventions
 *   - plain pointers + int64 sizes, no C++/torch types; every function returns an rdr_status (0 = OK).
 *   - `mem` says where the *bulk* arrays of that call live: RDR_MEM_HOST (library stages H2D/D2H itself)
 *     or RDR_MEM_DEVICE (pointers are CUDA device pointers of the handle's device; no copies).
 *     Small parameter vectors (axes, layer tables, maxlen) are always host pointers.
 *   - caller allocates every output; the library never frees caller memory; scratch lives in the handle.
 *   - one handle = one device + one stream; calls on a handle are serialised by the caller.
 *   - RDR_MEM_HOST calls are synchronous (outputs are complete on return).  RDR_MEM_DEVICE calls that return nothing to the
 *     host (rdr_sample, rdr_ray_points, ...) are enqueued on the handle's stream and return at once: order later work by
 *     rdr_set_stream (run on the consumer's stream) or rdr_synchronize.
 *   - no exception crosses the boundary; rdr_last_error() gives the message for the last failing call.
 *   - there is no CPU fallback: without a CUDA device rdr_create() fails with RDR_ERR_CUDA.
 */
#ifndef RAIDER_B200_H
#define RAIDER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDR_ABI_VERSION 3

typedef struct rdr_handle_s *rdr_handle_t;

typedef enum {
    RDR_OK = 0,
    RDR_ERR_INVALID = 1,   /* bad argument / shape  -> TypeError or ValueError in the shims (module.cpp:36-63) */
    RDR_ERR_CUDA = 2,      /* CUDA runtime failure   -> RuntimeError */
    RDR_ERR_STATE = 3,     /* call order (no cube set, integrate before layers, ...) -> RuntimeError */
    RDR_ERR_NO_LAYERS = 4, /* no model layer contributes (losreader.py:832-833 returns (None, None, None)) */
    RDR_ERR_ALL_NAN = 5    /* every ray length is NaN (delay.py:279-280 raises ValueError) */
} rdr_status;

enum { RDR_MEM_HOST = 0, RDR_MEM_DEVICE = 1 };
enum { RDR_F64 = 0, RDR_F32 = 1 };

/* cube field layout handed to rdr_set_cube */
enum {
    RDR_LAYOUT_ZYX = 0, /* (nz, ny, nx) C-order: the on-disk order of the processed weather model (weatherModel.py:676-724) */
    RDR_LAYOUT_YXZ = 1  /* (ny, nx, nz) C-order: the order getInterpolators hands to scipy (delayFcns.py:40-41) */
};

/* model CRS of the cube's x/y axes (delay.py:253 ecef_to_model) */
enum {
    RDR_CRS_GEOGRAPHIC = 0, /* x = lon deg, y = lat deg (EPSG:4326: ERA5/GMAO/HRES/MERRA2...) */
    RDR_CRS_LCC_SPHERE = 1  /* Lambert conformal conic on a sphere (HRRR, models/hrrr.py:255-260);
                               crs_params = {n, c, rho0, lam0_rad, R, x_0, y_0} */
};

/* where the ground (target) points of a ray-tracing call come from (delay.py:242,262-267) */
enum {
    RDR_GEOM_GRID = 0,  /* regular raster: gx = xpts[nx] (lon deg), gy = ypts[ny] (lat deg), one height `ht`; ray r = j*nx + i */
    RDR_GEOM_POINTS = 1 /* explicit points: gx = lon[n], gy = lat[n] (deg), n = ny*nx, all at height `ht` */
};

/* line-of-sight source (delay.py:270 los.getLookVectors) */
enum {
    RDR_LOS_ARRAY = 0,  /* los = [n][3] ECEF unit vectors ground->sensor (Raytracing.getLookVectors, losreader.py:219-255) */
    RDR_LOS_ENU_CONST = 1, /* los = {east, north, up}: constant local ENU vector, rotated to ECEF per pixel
                              (inc_hd_to_enu losreader.py:374-396 + enu2ecef utilFcns.py:91-121) */
    RDR_LOS_ZENITH = 2, /* los = NULL: local zenith (getZenithLookVecs, losreader.py:302-316) */
    RDR_LOS_ORBIT = 3,  /* los = {n_sv, n_sv rows of (t [s], x, y, z, vx, vy, vz)} on the host: per-ray zero-Doppler look vectors from
                           orbit state vectors, computed on the device (K6; replaces the per-pixel isce3 geo2rdr loop of
                           Raytracing.getLookVectors, losreader.py:219-255) */
    RDR_LOS_ENU_ARRAY = 4 /* los = [n][3] local ENU unit vectors, one per ray (inc_hd_to_enu per station, losreader.py:374-396), rotated
                           to ECEF on the device (enu2ecef); point mode only (rdr_ray_stations) */
};

/* interval semantics of the sampler */
enum {
    RDR_SEM_SCIPY = 0,       /* scipy RegularGridInterpolator(fill_value=nan, bounds_error=False): last node inclusive (delayFcns.py:55-56) */
    RDR_SEM_RAIDER_FILL = 1, /* RAiDER.interpolate with fill_value: bisect_left, upper edge exclusive (interpolate.cpp:116-125) */
    RDR_SEM_RAIDER_CLAMP = 2 /* RAiDER.interpolate without fill_value: clamp -> linear extrapolation (interpolate.cpp:127-131) */
};

/* ---------------------------------------------------------------- lifecycle ---------------------------- */
int rdr_abi_version(void);
/* number of CUDA devices visible to the library (0 when there is none); never fails */
int rdr_device_count(void);
int rdr_create(int device, rdr_handle_t *out);
int rdr_destroy(rdr_handle_t h);
/* message of the last failing call on `h` (or of the last failing handle-less call when h == NULL) */
const char *rdr_last_error(rdr_handle_t h);
/* run subsequent calls on an existing cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL = own stream */
int rdr_set_stream(rdr_handle_t h, void *cuda_stream);
int rdr_synchronize(rdr_handle_t h);
/* Page-locked host memory for result arrays (pooled: freed blocks are reused).  rdr_ray_integrate writes its outputs straight
 * into such arrays from the kernel (posted PCIe writes that overlap the integration) instead of staging + copying; any other
 * host pointer still works, through a staged copy.  The reference allocates its outputs with np.zeros (delay.py:248). */
int rdr_host_alloc(int64_t bytes, void **out);
int rdr_host_free(void *p);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches claim) */
int64_t rdr_launch_count(rdr_handle_t h);
/* rays the last rdr_ray_integrate handed from the fast integrator to the PROJ-form one (polar, very oblique, leaving the
 * cube, on a last node); -1 when the fast integrator was not used or the counters were not read back */
int64_t rdr_last_fix_count(rdr_handle_t h);

/* ---------------------------------------------------------------- cube --------------------------------- */
/* Replaces delayFcns.getInterpolators (delayFcns.py:23-58): stage one weather-model cube (two float32 fields on
 * one grid: wet+hydro for ray tracing, wet_total+hydro_total for zenith) into HBM, z fastest, the two fields and
 * the z/z+1 neighbours interleaved so one 16-byte load feeds a bilinear column pair.  Axes are host f64 arrays;
 * descending axes are flipped on the way in (scipy does the same at construction).  Fields are host or device
 * by `mem`. */
int rdr_set_cube(rdr_handle_t h, const double *ys, int64_t ny, const double *xs, int64_t nx, const double *zs, int64_t nz,
                 const float *wet, const float *hydro, int layout, int crs_kind, const double *crs_params, int mem);
/* Second epoch for temporal interpolation, blended as cube = w0*cube0 + w1*cube1 (cli/raider.py:817-819)
 * at staging time, in fp64 then rounded to fp32 exactly like the reference's xarray arithmetic. */
int rdr_blend_cube(rdr_handle_t h, const float *wet1, const float *hydro1, int layout, double w0, double w1, int mem);

/* ---------------------------------------------------------------- K2: trilinear sample ------------------ */
/* Replaces scipy RGI __call__ as configured at delayFcns.py:55-56 (call sites delay.py:120-121,214,319):
 * pts = [n][3] (y, x, z) in cube coordinates, dtype f64 or f32; out_wet/out_hydro = [n] same dtype. */
int rdr_sample(rdr_handle_t h, const void *pts, int64_t n, void *out_wet, void *out_hydro, int dtype, int semantics, int mem);
/* Replaces _build_cube (delay.py:196-216) for one height: sample both fields at (ypts[j], xpts[i], ht); out = [ny][nx] f64.
 * Query axes are in the cube's own CRS (model_crs == pts_crs branch, delay.py:210-211). */
int rdr_sample_grid(rdr_handle_t h, const double *xpts, int64_t nx, const double *ypts, int64_t ny, double ht,
                    double *out_wet, double *out_hydro, int mem);
/* The whole loop of _build_cube (delay.py:205-214) in one launch: every output height zpts[k]; out = [nh][ny][nx] f64. */
int rdr_sample_grid_levels(rdr_handle_t h, const double *xpts, int64_t nx, const double *ypts, int64_t ny, const double *zpts, int64_t nh,
                           double *out_wet, double *out_hydro, int mem);

/* ---------------------------------------------------------------- K0 + K3: ray tracing ------------------ */
/* Scalar layer decisions of build_ray (losreader.py:785-809) for the staged cube: writes up to nz-1 (low, high)
 * height pairs and their count.  Pure host logic, exposed for tests and the shims. */
int rdr_ray_plan(rdr_handle_t h, double ht, double zref, int64_t *n_layers, double *low_ht, double *high_ht);

/* K0 -- replaces build_ray + getTopOfAtmosphere (losreader.py:706-733,772-835) over a whole raster: one thread per
 * ray runs the 10-then-3 Newton schedule for every contributing layer, keeps the along-ray distance of each layer
 * top in handle scratch, and reduces max_over_raster(ray_length[k]) (delay.py:283) with warp shuffles + atomics.
 *   maxlen_out[n_layers]  host, this call's (this GPU's) per-layer maxima  -> all-reduce(MAX) across GPUs
 *   counts_out[5]         host: {n_rays, n_rays_with_nan_length, n_first_sample_below_zmin, n_layers, n_last_sample_above_zmax}
 *                         (ABI 3: five entries; the last one feeds the upper `.all()` clamp of delay.py:310-311)
 */
int rdr_ray_layers(rdr_handle_t h, int geom_kind, const double *gx, const double *gy, int64_t ny, int64_t nx,
                   int los_kind, const double *los, double ht, double zref,
                   double *maxlen_out, int64_t *counts_out, int mem);

/* K3 -- replaces the integration loops of _build_cube_ray (delay.py:283-323) for the rays of the last
 * rdr_ray_layers call: nParts from the (globally reduced) maxlen, sub-step points, ECEF -> model CRS, trilinear
 * wet+hydro sample, trapezoid weights, fp64 accumulation in layer-then-step order.
 *   maxlen[n_layers]      host: global per-layer maxima (delay.py:283)
 *   clamp                 bit 0: *all* pixels of the very first sample are below min(z) globally (delay.py:306-307: the sample is
 *                         taken at min(z)); bit 1 (ABI 3): all pixels of the very last sample -- the top of the top layer -- are above
 *                         max(z) globally (delay.py:310-311: taken at max(z); happens at the default zref = top - 1 m from ~58 deg
 *                         incidence on, where the three Newton iterates of losreader.py:720-733 overshoot by more than that metre)
 *   out_wet/out_hydro     [n_rays] f64 (or f32 when out_dtype == RDR_F32); accumulate != 0 -> out += (delay.py:245-248,323)
 *   nparts_out[n_layers]  host, optional: the integer step counts used (bit-exact contract)
 *   oob_out[2]            host, optional: {samples below min(z), samples above max(z)} that became NaN
 */
int rdr_ray_integrate(rdr_handle_t h, const double *maxlen, double max_segment_length, int clamp,
                      void *out_wet, void *out_hydro, int out_dtype, int accumulate,
                      int64_t *nparts_out, int64_t *oob_out, int mem);

/* ---------------------------------------------------------------- the fused step: K0 -> plan -> K3 without the host ---- */
/* status bits of the device-side step plan (rdr_trace_result info[0] / info[1]) */
enum {
    RDR_PLAN_ABSURD = 1,        /* a per-layer maximum is NaN / absurd: ceil(max / MAX_SEGMENT_LENGTH) is not a step count (delay.py:283) */
    RDR_PLAN_ALL_NAN = 2,       /* np.isnan(ray_lengths).all() over the WHOLE raster, all ranks (delay.py:279-280 raises ValueError) */
    RDR_PLAN_KNIFE_EDGE = 4,    /* some max / MAX_SEGMENT_LENGTH lies within 1e-6 of an integer: nParts (delay.py:283) is decided by the
                                   last bits of the maximum; when K0 ran in its default form the step is NOT integrated and the caller
                                   redoes it with RDR_TRACE_EXACT_K0 */
    RDR_PLAN_SPAN_TOO_LONG = 8  /* one layer is longer than the span cubics of the polynomial integrators allow: redo with mode 1 */
};
/* flags of rdr_trace_begin */
enum {
    RDR_TRACE_EXACT_K0 = 1,      /* Newton iterates of getTopOfAtmosphere (losreader.py:720-733) on PROJ-form (Bowring) heights instead of
                                    on the per-ray polynomials of K0 (h(t) as a septic, the layer tops as a degree-7 polynomial in z): layer maxima to
                                    ~1e-9 m of the reference's instead of ~1e-8 m */
    RDR_TRACE_NO_KNIFE_GUARD = 2 /* integrate even on an nParts knife edge (tuning runs) */
};
/* The same three stages as rdr_ray_layers / rdr_ray_integrate, but nothing returns to the host in between: K0's per-layer maxima
 * and predicate counters stay in HBM, a one-CTA kernel turns them into the step plan (nParts of delay.py:283, layer records, spans,
 * the `.all()` clamp of delay.py:306-307, the all-NaN check of delay.py:279), and the integration kernels read the plan from
 * device memory.  rdr_trace_begin and rdr_trace_finish only enqueue work on the handle's stream (device-resident or page-locked
 * outputs are complete once the stream reaches that point; other host outputs once rdr_trace_result returns).
 *
 * Across GPUs (rdr_set_exchange): rdr_trace_begin also stores this rank's K + 5 words into its slot of every peer's exchange
 * buffer; the caller puts ONE barrier on the stream (all ranks' rdr_trace_begin work done) before rdr_trace_finish, whose plan
 * kernel takes MAX / SUM over the slots -- the all-reduce of SURVEY 8(e) without NCCL or the host -- and a second barrier after it
 * (which also publishes the delay maps written through rdr_set_peer_outputs).
 *   force_clamp   -1: the plan decides both `.all()` clamps (delay.py:306-307 first sample below min(z), :310-311 last sample above
 *                 max(z)) from K0's global counts; >= 0: bit 0 = the lower clamp's value (redo after a cross-check miss), bit 1 / bit 2 =
 *                 upper clamp forced on / off (neither: from the count)
 *   mode          0: polynomial integrators (quadrature + thin-layer kernels); 1: per-sample Bowring form; 2: PROJ-form for every sample
 * rdr_trace_result synchronises the stream and reports the step:
 *   maxlen_out[n_layers], nparts_out[n_layers]  global maxima and the integer step counts used (may be NULL)
 *   info_out[20] = {status, blocked, n_layers, n_rays, n_nan_rays, K0's #first samples below min(z), K3's own count of the same,
 *                   clamp used, #samples below min(z), #samples above max(z), #rays redone in PROJ form, knife-edge layer or -1,
 *                   k_split (layers handled by the thin-layer kernel), n_spans, K0 ran its polynomial form, K3 ran the polynomial form,
 *                   #CTA passes of the thin-layer kernel with TMA-staged record columns, #passes without,
 *                   K0's #last samples above max(z), upper clamp used};
 *                   counts 3..6 and 18 are global (all ranks), 8..10 and 16..17 this rank's.  blocked != 0: nothing was integrated.
 *                   Rays whose last sample lies above max(z) are integrated by the PROJ-form kernel (list mode), which applies the
 *                   upper clamp; the polynomial kernels hand them over. */
int rdr_trace_begin(rdr_handle_t h, int geom_kind, const double *gx, const double *gy, int64_t ny, int64_t nx, int los_kind,
                    const double *los, double ht, double zref, int flags, int mem);
int rdr_trace_finish(rdr_handle_t h, double max_segment_length, int force_clamp, int mode, void *out_wet, void *out_hydro,
                     int out_dtype, int accumulate, int mem);
int rdr_trace_result(rdr_handle_t h, double *maxlen_out, int64_t *nparts_out, int64_t *info_out);
/* bufs[q] = device address of rank q's exchange buffer (rdr_exchange_bytes(world) bytes each, zero-initialised, peer-mapped: torch
 * symmetric memory / cudaIpc), bufs[rank] being this rank's own.  world = 0 detaches. */
int rdr_set_exchange(rdr_handle_t h, int rank, int world, void *const *bufs);
int64_t rdr_exchange_bytes(int world);

/* Multi-GPU: mirror the results of the following rdr_ray_integrate calls into up to 8 more device buffers --
 * the same row block of the full delay maps in the HBM of the node's other GPUs, peer-mapped (torch symmetric memory /
 * cudaIpc) by the caller.  wet[i] / hydro[i] point at the first ray of this rank's block inside peer i's maps and have the
 * element type of the call's out_dtype.  The integration kernel stores each ray to every destination as the ray finishes
 * (posted NVLink writes, overlapped with the integration): the all-gather that reassembles the output map, without a
 * collective.  The caller orders the kernels of all ranks (barrier) before reading the maps.  n = 0 clears the list;
 * accumulate != 0 calls ignore it.  The reference has no parallel delay path (delay.py:178-185 raises for nproc > 1). */
int rdr_set_peer_outputs(rdr_handle_t h, int n, void *const *wet, void *const *hydro);
/* The same through NVLink-SHARP multicast: wet_mc / hydro_mc are the addresses of this rank's row block inside the MULTICAST mapping
 * of the symmetric maps (torch symmetric memory `multicast_ptr`, cuMulticast*): the integration kernel issues ONE multimem.st per
 * value and the NVSwitch replicates it into every GPU's copy (the local one included).  NULL, NULL detaches. */
int rdr_set_multicast_outputs(rdr_handle_t h, void *wet_mc, void *hydro_mc);

/* K1b -- north_star (a): generate the 3-D sample points along each ray of the last rdr_ray_layers call, in model
 * coordinates, i.e. the per-sub-step `pts` arrays of delay.py:292-298 (replaces the role of tools/bindings makePoints3D in
 * the unfused dataflow).  Unique samples are numbered in layer-then-step order ("slots"); pts = [nslots][n_rays][3] (y, x, z)
 * for slots slot0 .. slot0+nslots-1.  pts == NULL only reports total_slots. */
int rdr_ray_points(rdr_handle_t h, const double *maxlen, double max_segment_length, int64_t slot0, int64_t nslots, void *pts, int dtype,
                   int64_t *total_slots, int mem);

/* API-parity pieces of losreader (small problems, tests): */
/* getTopOfAtmosphere(xyz, look_vecs, toaheight, factor=None) losreader.py:706-733; factor == NULL -> 10 iterations */
int rdr_top_of_atmosphere(const double *xyz, const double *look, int64_t n, double toaheight, const double *factor,
                          double *out_xyz, int device);
/* build_ray(model_zs, ht, xyz, LOS, MAX_TROPO_HEIGHT) losreader.py:772-835: outputs [K][n], [K][n][3], [K][n][3]; K via rdr_ray_plan */
int rdr_build_ray(const double *model_zs, int64_t nz, double ht, const double *xyz, const double *look, int64_t n, double zref,
                  int64_t *n_layers, double *ray_lengths, double *low_xyzs, double *high_xyzs, int device);
/* lla2ecef / ecef2lla (utilFcns.py:77-88); n points, SoA in / SoA out, host pointers */
int rdr_lla2ecef(const double *lat, const double *lon, const double *hgt, int64_t n, double *x, double *y, double *z, int device);
int rdr_ecef2lla(const double *x, const double *y, const double *z, int64_t n, double *lon, double *lat, double *hgt, int device);

/* ---------------------------------------------------------------- K1: makePoints ------------------------ */
/* Npts rule of makePoints.pyx:130-134 */
int rdr_make_points_count(double max_len, double step, int64_t *npts);
/* makePoints{0,1,2,3}D (makePoints.pyx:15-148): sp/slv = [n_rays][3]; out = [n_rays][3][npts], out[r][c][k] = sp[r][c] + (k*step)*slv[r][c] */
int rdr_make_points(double max_len, const double *sp, const double *slv, int64_t n_rays, double step, double *out, int64_t npts,
                    int device, int mem);

/* ---------------------------------------------------------------- RAiDER.interpolate -------------------- */
/* interpolate(points, values, interp_points, fill_value, assume_sorted, max_threads) module.cpp:26-294:
 * ndim grids (host), values f64 C-order, pts = [n][ndim], out = [n].  ndim <= 8. */
int rdr_interpolate(int ndim, const double *const *grids, const int64_t *sizes, const double *values, const double *pts, int64_t n,
                    int has_fill, double fill_value, double *out, int device, int mem);
/* interpolate_along_axis (module.cpp:296-493 -> interpolate.cpp:260-332) on columns made contiguous by the shim:
 * x, y = [ncol][nin]; xnew, out = [ncol][nout]. */
int rdr_interp_along_axis(const double *x, const double *y, const double *xnew, int64_t ncol, int64_t nin, int64_t nout,
                          int has_fill, double fill_value, double *out, int device, int mem);

/* K5 -- station (point) mode, BASELINE config C4: n rays with their own longitude / latitude / height, each traced as the reference
 * would trace a 1 x 1 raster at that height (_build_cube_ray(xpts=[lon], ypts=[lat], zpts=[h]), delay.py:219-326): the layer
 * plan, nParts = ceil(own length / max_segment_length) + 1 and the `.all()` clamps are per ray.  One warp per ray, layers spread
 * over the lanes, warp-shuffle reduction.  los_kind: RDR_LOS_ARRAY (ECEF [n][3]), RDR_LOS_ENU_ARRAY ([n][3]), RDR_LOS_ENU_CONST,
 * RDR_LOS_ZENITH, RDR_LOS_ORBIT (host payload as above).  out_nsamples (may be NULL): sum of nParts of each ray. */
int rdr_ray_stations(rdr_handle_t h, const double *lon, const double *lat, const double *hgt, int64_t n, int los_kind, const double *los,
                     double zref, double max_segment_length, double *out_wet, double *out_hydro, int32_t *out_nsamples, int mem);

/* K7 -- the processing step that produces the cube (WeatherModel.load after load_weather, models/weatherModel.py:252-260):
 * _find_e (specific humidity q, or relative humidity when hum_is_rh), _uniform_in_z (interpolate_along_axis x 3 onto `zlevels`,
 * NaN fill, float32), _checkForNans (fillna3D), k2 e/T + k3 e/T^2 and k1 P/T, _adjust_grid (a level at zmin when zmin <
 * zlevels[0]) and _getZTD.  Inputs: ncol columns of nl native levels, z fastest ((y, x, z) like the reference's arrays), heights
 * ascending per column.  Outputs: [ncol][*nz_written] float32, z fastest (stage them with RDR_LAYOUT_YXZ); p/t/e outputs are
 * optional (all three or none).  mem applies to the bulk arrays; zlevels is a host parameter vector. */
int rdr_prepare_cube(int64_t ncol, int64_t nl, const double *zs, const double *p, const double *t, const double *hum, int hum_is_rh,
                     const double *zlevels, int64_t nz_out, double k1, double k2, double k3, double zmin, float *wet, float *hydro,
                     float *wet_total, float *hydro_total, float *p_out, float *t_out, float *e_out, int64_t *nz_written, int device, int mem);

/* K6 on its own -- Raytracing.getLookVectors (losreader.py:219-255): ECEF unit vectors ground -> sensor at the zero-Doppler time
 * of every target, from n_sv uniformly spaced state-vector rows (t, x, y, z, vx, vy, vz).  Targets: a raster (RDR_GEOM_GRID:
 * gx[nx] lon, gy[ny] lat, height ht) or points (RDR_GEOM_POINTS: gx/gy/[hgt] of length ny*nx).  isce3's defaults at the
 * reference call site: threshold 1e-7 m, maxiter 30.  Targets that do not converge / leave the orbit span get NaN vectors.
 * out_slant / out_aztime may be NULL.  Host pointers. */
int rdr_orbit_los(const double *sv_rows, int64_t n_sv, int geom_kind, const double *gx, const double *gy, const double *hgt, double ht,
                  int64_t ny, int64_t nx, double threshold, int maxiter, double *out_los, double *out_slant, double *out_aztime, int device);

/* ---------------------------------------------------------------- test hook ----------------------------- */
/* Counts, over n pseudo-random (numerator, cell width) pairs, how often the sampler's table-driven division (n * RN(1/d) with
 * one / two Markstein corrections) differs from IEEE n / d on the device.  The sampler uses the one-step form (0 mismatches in 2e8 trials; two-step kept for reference). */
int rdr_selftest_div(int64_t n, uint64_t seed, int64_t *mismatch_1step, int64_t *mismatch_2step, int device);

#ifdef __cplusplus
}
#endif
#endif /* RAIDER_B200_H */

// k_parity.cuh -- small API-parity kernels (top of atmosphere, build_ray, geodesy, makePoints, interpolate).
// A fragment of libraider_b200.so: included by raider_b200.cu INSIDE its anonymous namespace, in this order (the translation unit
// is the one file it used to be; see the kernel inventory at the top of raider_b200.cu).  Not a stand-alone header.
// ------------------------------------------------------------------------------------------------
// small API-parity kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_top_of_atmosphere(const double *__restrict__ xyz, const double *__restrict__ look, int64_t n, double toa,
                                    const double *__restrict__ factor, double *__restrict__ out) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const Vec3 g = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]}, u = {look[3 * r], look[3 * r + 1], look[3 * r + 2]};
    double t;
    const Vec3 p = factor ? top_of_atmosphere<3>(g, u, toa, 1.0 / factor[r], t) : top_of_atmosphere<10>(g, u, toa, 1.0, t);
    out[3 * r] = p.x;
    out[3 * r + 1] = p.y;
    out[3 * r + 2] = p.z;
}

__global__ void k_build_ray(const double *__restrict__ xyz, const double *__restrict__ look, int64_t n, int K, const double *__restrict__ plan,
                            double *__restrict__ lens, double *__restrict__ lows, double *__restrict__ highs) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const Vec3 g = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]}, u = {look[3 * r], look[3 * r + 1], look[3 * r + 2]};
    Vec3 lo, hi;
    double rcosf = 1.0, t;
    for (int k = 0; k < K; ++k) {
        const double a = plan[k], b = plan[K + k];
        if (k == 0) {
            lo = top_of_atmosphere<10>(g, u, a, 1.0, t);
            hi = top_of_atmosphere<10>(g, u, b, 1.0, t);
        } else {
            lo = hi;
            hi = top_of_atmosphere<3>(g, u, b, rcosf, t);
        }
        const double len = norm3(hi - lo);
        if (k == 0) rcosf = len / (b - a);
        const int64_t o = (int64_t)k * n + r;
        lens[o] = len;
        lows[3 * o] = lo.x; lows[3 * o + 1] = lo.y; lows[3 * o + 2] = lo.z;
        highs[3 * o] = hi.x; highs[3 * o + 1] = hi.y; highs[3 * o + 2] = hi.z;
    }
}

__global__ void k_lla2ecef(const double *lat, const double *lon, const double *hgt, int64_t n, double *x, double *y, double *z) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    double a, b, c2, d;
    const Vec3 p = lla2ecef(lat[r], lon[r], hgt[r], a, b, c2, d);
    x[r] = p.x; y[r] = p.y; z[r] = p.z;
}

__global__ void k_ecef2lla(const double *x, const double *y, const double *z, int64_t n, double *lon, double *lat, double *hgt) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    double lo, la, h;
    ecef2lla({x[r], y[r], z[r]}, lo, la, h);
    lon[r] = lo; lat[r] = la; hgt[r] = h;
}

// K1: makePoints (makePoints.pyx:142-147): out[r][c][k] = sp[r][c] + (k*step)*slv[r][c]; separate multiply and add, no FMA,
// because the reference is built without FMA contraction (setup.py:31-37) -- bit-exact against test_result_makePoints3D.txt
__global__ void k_make_points(const double *__restrict__ sp, const double *__restrict__ slv, int64_t n_rays, double step, int64_t npts,
                              double *__restrict__ out) {
    const int64_t total = n_rays * 3 * npts;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i % npts, rc = i / npts;
        const double base = __dmul_rn((double)k, step);  // np.arange(0, L+step, step)[k]
        __stcs(out + i, __dadd_rn(__ldg(sp + rc), __dmul_rn(base, __ldg(slv + rc))));
    }
}

// K4: interpolate_along_axis (interpolate.h:78-118 per column): one thread per output element
__global__ void k_interp_axis(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ xnew, int64_t ncol,
                              int nin, int nout, int has_fill, double fill, double *__restrict__ out) {
    const int64_t total = ncol * nout;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t col = i / nout;
        const double *gx = x + col * nin, *gy = y + col * nin;
        const double v = xnew[i];
        int hi = bisect_left(gx, nin, v);
        if (has_fill) {
            if (hi < 1 || hi > nin - 1) {
                out[i] = fill;
                continue;
            }
        } else {
            hi = hi < 1 ? 1 : (hi > nin - 1 ? nin - 1 : hi);
        }
        const double x0 = gx[hi - 1], x1 = gx[hi], y0 = gy[hi - 1], y1 = gy[hi];
        const double slope = __ddiv_rn(y1 - y0, x1 - x0);
        out[i] = __dadd_rn(y0, __dmul_rn(slope, v - x0));
    }
}

// RAiDER.interpolate.interpolate for ndim = 1, 2, 3 (dedicated formulas) and N-D (corner bitmask walk)
struct NdGrid {
    const double *g[8];
    int n[8];
    int ndim;
};

__global__ void k_interp_nd(const NdGrid G, const double *__restrict__ values, const double *__restrict__ pts, int64_t n, int has_fill,
                            double fill, double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int nd = G.ndim;
        int hi[8];
        double dlo[8], dhi[8], span[8];
        bool filled = false;
        for (int d = 0; d < nd; ++d) {
            const double v = pts[i * nd + d];
            int k = bisect_left(G.g[d], G.n[d], v);
            if (has_fill) {
                if (k < 1 || k > G.n[d] - 1) {
                    filled = true;
                    break;
                }
            } else {
                k = k < 1 ? 1 : (k > G.n[d] - 1 ? G.n[d] - 1 : k);
            }
            hi[d] = k;
            const double g0 = G.g[d][k - 1], g1 = G.g[d][k];
            dlo[d] = v - g0;
            dhi[d] = g1 - v;
            span[d] = g1 - g0;
        }
        if (filled) {
            out[i] = fill;
            continue;
        }
        if (nd == 1) {  // interpolate.h:109-116
            const double y0 = values[hi[0] - 1], y1 = values[hi[0]];
            const double slope = __ddiv_rn(y1 - y0, span[0]);
            out[i] = __dadd_rn(y0, __dmul_rn(slope, dlo[0]));
        } else if (nd == 2) {  // interpolate.cpp:61-81
            const int64_t n1 = G.n[1];
            const double z00 = values[(hi[0] - 1) * n1 + hi[1] - 1], z01 = values[(hi[0] - 1) * n1 + hi[1]];
            const double z10 = values[hi[0] * n1 + hi[1] - 1], z11 = values[hi[0] * n1 + hi[1]];
            const double a = __dadd_rn(__dmul_rn(z00, dhi[1]), __dmul_rn(z01, dlo[1]));
            const double b = __dadd_rn(__dmul_rn(z10, dhi[1]), __dmul_rn(z11, dlo[1]));
            out[i] = __ddiv_rn(__dadd_rn(__dmul_rn(dhi[0], a), __dmul_rn(dlo[0], b)), __dmul_rn(span[0], span[1]));
        } else if (nd == 3) {  // interpolate.cpp:138-174
            const int64_t n1 = G.n[1], n2 = G.n[2];
            const int64_t l0 = (hi[0] - 1) * n1 * n2, h0 = hi[0] * n1 * n2, l1 = (hi[1] - 1) * n2, h1 = hi[1] * n2, l2 = hi[2] - 1, h2 = hi[2];
            out[i] = trilinear_raider(values[l0 + l1 + l2], values[l0 + l1 + h2], values[l0 + h1 + l2], values[l0 + h1 + h2],
                                      values[h0 + l1 + l2], values[h0 + l1 + h2], values[h0 + h1 + l2], values[h0 + h1 + h2], dlo[0], dhi[0],
                                      dlo[1], dhi[1], dlo[2], dhi[2], __dmul_rn(__dmul_rn(span[0], span[1]), span[2]));
        } else {  // interpolate.cpp:204-256
            double vol = 1.0;
            for (int d = 0; d < nd; ++d) vol = __dmul_rn(vol, span[d]);
            double acc = 0.0;
            for (unsigned j = 0; j < (1u << nd); ++j) {
                int64_t index = 0;
                for (int d = 0; d < nd; ++d) {
                    index += ((j >> d) & 1) ? hi[d] : hi[d] - 1;
                    index *= (d + 1 < nd) ? G.n[d + 1] : 1;
                }
                double term = values[index];
                for (int d = 0; d < nd; ++d) term = __dmul_rn(term, ((j >> d) & 1) ? dlo[d] : dhi[d]);
                acc = __dadd_rn(acc, term);
            }
            out[i] = __ddiv_rn(acc, vol);
        }
    }
}

// self-test of the table-driven exact division: random cell widths d (any mantissa, exponents 2^-8 .. 2^16) and numerators
// n = u * d, u in [0, 1]; counts results that differ from IEEE n / d
__global__ void k_selftest_div(int64_t n, unsigned long long seed, unsigned long long *mis) {
    unsigned long long m1 = 0, m2 = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
        auto next = [&x]() {
            x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
            return x * 0x2545F4914F6CDD1Dull;
        };
        const unsigned long long a = next(), b = next();
        const int e = (int)(next() % 25) - 8;
        const double d = ldexp(1.0 + (double)(a >> 12) * 0x1p-52, e);
        double u = (double)(b >> 11) * 0x1p-53;
        if ((b & 1023) == 0) u = 1.0;  // t == 1 happens (inclusive last node)
        const double num = u * d;
        const double inv = 1.0 / d, want = num / d;
        m1 += div_exact1(num, d, inv) != want;
        m2 += div_exact2(num, d, inv) != want;
    }
    if (m1) atomicAdd(mis, m1);
    if (m2) atomicAdd(mis + 1, m2);
}

// host-side restatement of the scalar layer decisions of build_ray (losreader.py:785-809)
void layer_plan(const std::vector<double> &zs, double ht, double zref, std::vector<double> &low, std::vector<double> &high,
                std::vector<int> &cell) {
    low.clear();
    high.clear();
    cell.clear();
    const size_t nz = zs.size();
    for (size_t zz = 0; zz + 1 < nz; ++zz) {
        double low_ht = zs[zz], high_ht = zs[zz + 1];
        if (high_ht == zs[nz - 1]) high_ht -= 0.01;
        if (high_ht < ht || low_ht >= zref) continue;
        if (low_ht < ht) low_ht = ht;
        if (high_ht > zref) high_ht = zref;
        if (fabs(high_ht - low_ht) < 1.0) continue;
        low.push_back(low_ht);
        high.push_back(high_ht);
        cell.push_back((int)zz);
    }
}

// The Npts rule of makePoints.pyx:130-134 as Cython compiles it for C doubles: `a // b` is floor(a / b) and `a % b` is
// fmod with Python's sign convention (__Pyx_mod_double).  Pinned against the compiled reference (tests/golden/makepoints.npz).
int64_t make_points_npts(double max_len, double step) {
    double r = fmod(max_len, step);
    if (r != 0.0 && ((r < 0.0) != (step < 0.0))) r += step;
    int64_t n = (int64_t)floor(max_len / step);
    if (r != 0.0) n += 1;
    return n;
}

// rows of (t, x, y, z, vx, vy, vz) -> t[n] | pos[n][3] | vel[n][3]; isce3.core.Orbit needs >= 4 uniformly spaced, increasing times
int split_orbit(rdr_handle_t h, const double *rows, int64_t n_sv, std::vector<double> &blob) {
    CHECK_ARG(h, rows != nullptr && n_sv >= 4 && n_sv < (1 << 20), "orbit: at least 4 state vectors are required for Hermite interpolation");
    blob.resize((size_t)n_sv * 7);
    for (int64_t i = 0; i < n_sv; ++i) {
        blob[i] = rows[7 * i];
        for (int c = 0; c < 3; ++c) {
            blob[n_sv + 3 * i + c] = rows[7 * i + 1 + c];
            blob[4 * n_sv + 3 * i + c] = rows[7 * i + 4 + c];
        }
    }
    const double dt = (blob[n_sv - 1] - blob[0]) / (double)(n_sv - 1);
    CHECK_ARG(h, dt > 0, "orbit: state-vector times must increase");
    for (int64_t i = 1; i < n_sv; ++i)
        CHECK_ARG(h, fabs((blob[i] - blob[i - 1]) - dt) <= 1e-6 * dt, "orbit: state vectors must be uniformly spaced in time");
    return RDR_OK;
}

struct ScopedDevice {
    int prev = -1;
    explicit ScopedDevice(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~ScopedDevice() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

// stage a host array to device scratch (or pass a device pointer through)
template <typename T>
int stage_in(rdr_handle_t h, DevBuf &buf, const T *src, size_t count, int mem, const T **out) {
    if (mem == RDR_MEM_DEVICE) {
        *out = src;
        return RDR_OK;
    }
    CUDA_TRY(h, buf.reserve(std::max<size_t>(count * sizeof(T), 16)));
    CUDA_TRY(h, cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    *out = buf.as<T>();
    return RDR_OK;
}


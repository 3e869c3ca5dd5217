// libraider_b200.so -- hand-written sm_100a kernels + C ABI for the RAiDER slant/zenith delay hot path.
// See include/raider_b200.h for the boundary and DESIGN.md for the kernel inventory:
//   K0 k_ray_layers          build_ray/getTopOfAtmosphere over a raster (h(t) as one septic per ray, the layer tops as one polynomial in z) + global
//                            per-layer max length
//   K3 k_ray_integrate_poly  the production integrator: span cubics of the cube coordinates, closed-form layer sums, register-held
//                            cell record; also stores the results into the other GPUs' maps (rdr_set_peer_outputs)
//      k_ray_integrate_thin  runs of layers with <= 3 samples (the 145-node production tables): cell records staged in shared memory
//                            per span by TMA bulk copies, along-ray distances through a cp.async ring
//      k_plan / k_publish    the step plan (nParts, layer records, spans, predicates) built on the device from K0's maxima -- of
//                            all ranks, exchanged through peer-mapped symmetric memory -- between K0 and K3: no host round trip
//      k_ray_integrate_fast  per-sample Bowring form (tests / comparisons);  k_ray_integrate  PROJ-form per sample (flagged rays, any CRS)
//   K2 k_sample_stream*      unfused trilinear sampler (points streamed from HBM through a TMA-bulk ring) -- the HBM-roofline kernel;
//                            k_sample_stream_f32 its fp32 tier
//   K1 k_make_points         makePoints{0..3}D;  K1b k_ray_points  the sample points of the rays
//   K4 k_interp_axis         interpolate_along_axis;  k_interp_nd  RAiDER.interpolate.interpolate
//   K5 k_ray_stations        station / point mode (one warp per ray);  K6 k_orbit_los  look vectors from orbit state vectors
//   K7 k_prepare_columns     weather-model processing (find_e, uniform_in_z, fillna, refractivity, ZTD)
// No CPU fallback lives here: every entry point needs a CUDA device.
#include "../../include/raider_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "geodesy.cuh"
#include "sampler.cuh"
#include "fastpath.cuh"

using namespace rdr;

#define RDR_API extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------------------------------------
// handle + error plumbing
// ------------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_last_error;

// Freed device blocks are parked in a small per-process pool instead of going back to the driver: the reference-shaped
// API creates a fresh cube (handle) per weather-model file, and re-allocating GB-sized ray scratch with cudaMalloc/cudaFree
// on every call would cost more than the kernels.
struct BlockPool {
    struct Block {
        int device;
        void *p;
        size_t cap;
    };
    std::mutex mu;
    std::vector<Block> free_blocks;
    size_t cached = 0;
    static constexpr size_t MAX_CACHED = 48ull << 30;

    void *take(int device, size_t bytes, size_t &cap) {
        std::lock_guard<std::mutex> lk(mu);
        int best = -1;
        for (int i = 0; i < (int)free_blocks.size(); ++i) {
            const Block &b = free_blocks[i];
            if (b.device == device && b.cap >= bytes && b.cap <= 2 * bytes + (1 << 20) && (best < 0 || b.cap < free_blocks[best].cap)) best = i;
        }
        if (best < 0) return nullptr;
        Block b = free_blocks[best];
        free_blocks.erase(free_blocks.begin() + best);
        cached -= b.cap;
        cap = b.cap;
        return b.p;
    }
    void give(int device, void *p, size_t cap) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (cached + cap <= MAX_CACHED) {
                free_blocks.push_back({device, p, cap});
                cached += cap;
                return;
            }
        }
        cudaFree(p);
    }
    void trim(int device) {  // out of memory: hand everything cached on this device back to the driver
        std::lock_guard<std::mutex> lk(mu);
        for (int i = (int)free_blocks.size() - 1; i >= 0; --i)
            if (free_blocks[i].device == device) {
                cudaFree(free_blocks[i].p);
                cached -= free_blocks[i].cap;
                free_blocks.erase(free_blocks.begin() + i);
            }
    }
};
BlockPool g_pool;

// Page-locked host blocks for result arrays (rdr_host_alloc): cudaHostAlloc costs ~0.3 ms per MB, so freed blocks are parked
// and handed out again.  Kernels write results straight into such blocks (they are device-mapped under UVA), which turns the
// device->host copy of the delay maps into posted PCIe writes that overlap the integration.
struct PinnedPool {
    struct Block {
        void *p;
        size_t cap;
    };
    std::mutex mu;
    std::vector<Block> free_blocks, live;
    size_t cached = 0;
    static constexpr size_t MAX_CACHED = 8ull << 30;

    void *take(size_t bytes) {
        std::lock_guard<std::mutex> lk(mu);
        int best = -1;
        for (int i = 0; i < (int)free_blocks.size(); ++i) {
            const Block &b = free_blocks[i];
            if (b.cap >= bytes && b.cap <= 2 * bytes + (1 << 16) && (best < 0 || b.cap < free_blocks[best].cap)) best = i;
        }
        if (best < 0) return nullptr;
        Block b = free_blocks[best];
        free_blocks.erase(free_blocks.begin() + best);
        cached -= b.cap;
        live.push_back(b);
        return b.p;
    }
    void add_live(void *p, size_t cap) {
        std::lock_guard<std::mutex> lk(mu);
        live.push_back({p, cap});
    }
    bool give(void *p) {  // false: not one of ours
        Block b{nullptr, 0};
        {
            std::lock_guard<std::mutex> lk(mu);
            for (size_t i = 0; i < live.size(); ++i)
                if (live[i].p == p) {
                    b = live[i];
                    live.erase(live.begin() + i);
                    break;
                }
            if (!b.p) return false;
            if (cached + b.cap <= MAX_CACHED) {
                free_blocks.push_back(b);
                cached += b.cap;
                return true;
            }
        }
        cudaFreeHost(b.p);
        return true;
    }
};
PinnedPool g_pinned;

// is `p` page-locked host memory the device can address directly (UVA)?  Returns the device-side alias or nullptr.
void *device_alias_of_host(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int device = -1;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        cudaGetDevice(&device);
        if ((p = g_pool.take(device, bytes, cap))) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            g_pool.trim(device);
            e = cudaMalloc(&p, bytes);
        }
        if (e == cudaSuccess) cap = bytes; else p = nullptr;
        return e;
    }
    void release() {
        if (p) g_pool.give(device, p, cap);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return static_cast<T *>(p); }
};

constexpr int MAX_LAYERS = 1024;

// extra (peer-GPU) destinations of the integrator's results: see store_result
constexpr int RDR_MAX_PEERS = 8;
struct PeerOut {
    void *wet[RDR_MAX_PEERS];
    void *hydro[RDR_MAX_PEERS];
    int n;
    int multicast;  // wet[0] / hydro[0] are NVLink-SHARP multicast addresses: ONE multimem.st reaches every GPU of the group
};

}  // namespace

struct rdr_handle_s {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int64_t launches = 0;

    // cube
    bool has_cube = false;
    int64_t ny = 0, nx = 0, nz = 0;
    std::vector<double> ys, xs, zs;  // ascending host copies
    bool flip_y = false, flip_x = false, flip_z = false;
    int crs_kind = RDR_CRS_GEOGRAPHIC;
    double crs[7] = {0, 0, 0, 0, 0, 0, 0};
    DevBuf d_axes;   // ys | xs | zs (nodes)
    DevBuf d_tabs;   // per axis: interval records (double4) then first-guess bins (uint16)
    size_t tab_cell_off[3] = {0, 0, 0}, tab_bin_off[3] = {0, 0, 0}, tab_rec32_off[3] = {0, 0, 0};
    int tab_nbin[3] = {0, 0, 0};
    DevBuf d_cells;  // double4 [ny][nx][nz-1]  {wet[z], hydro[z], wet[z+1], hydro[z+1]}
    DevBuf d_cells32;  // float4 [ny][nx][nz-1]: the same records in fp32 for the streaming sampler K2
    DevBuf d_lerp;   // LerpCell [ny-1][nx-1][nz-1]: 128-byte cell records of the fast integrator (fastpath.cuh)
    DevBuf d_stage;  // staging for field uploads (and packed fp32 pairs for blending)
    DevBuf d_fields; // float2 [ny][nx][nz] (wet, hydro) kept for blending

    // ray state (between rdr_ray_layers and rdr_ray_integrate)
    bool has_rays = false;
    int64_t n_rays = 0, ray_ny = 0, ray_nx = 0;
    int geom_kind = 0, los_kind = 0;
    double ht = 0, zref = 0;
    int n_layers = 0;
    std::vector<double> low_ht, high_ht;
    std::vector<int> layer_cell;  // model layer index of each contributing layer (z-cell hint)
    DevBuf d_gx, d_gy, d_los;     // staged geometry when the caller's arrays are on the host (d_los: also the orbit-derived vectors)
    DevBuf d_orbit;               // t[n] | pos[n][3] | vel[n][3] of RDR_LOS_ORBIT
    const double *p_gx = nullptr, *p_gy = nullptr, *p_los = nullptr;
    double los_const[3] = {0, 0, 1};
    DevBuf d_plan;    // low[K] | high[K]
    DevBuf d_t;       // [K+1][n_rays] along-ray distances: row 0 = bottom of first layer, row k+1 = top of layer k
    DevBuf d_red;     // maxlen bits [K] | counters
    DevBuf d_nparts;  // int [K] + int cell [K]
    DevBuf d_layers;  // LayerRec [K] of the fast integrator
    DevBuf d_spans;   // int [nspan]: one-past-last layer of every span of the polynomial integrator
    DevBuf d_fix;     // int [n_rays]: rays the fast integrator handed to the PROJ-form path
    int64_t last_fix_count = -1;  // how many rays that was in the last rdr_ray_integrate (-1: fast path not used / not read back)
    PeerOut peers = {};  // extra (peer-GPU) destinations of the next rdr_ray_integrate: rdr_set_peer_outputs
    DevBuf d_out;     // staging for host outputs
    DevBuf d_in;      // staging for host inputs of K2
    DevBuf d_devplan; // DevPlan: the step plan k_plan builds on the device
    DevBuf d_part;    // double [2][n_rays]: partial sums the quadrature kernel hands to the thin-layer kernel
    bool thin_ok = true;        // the record array carries the prefetch padding (always, kept for clarity)
    bool k0_was_cubic = true;   // last K0 ran in its polynomial form (default) rather than on Bowring heights
    bool last_k3_poly = false;
    int trace_flags = 0;
    double last_max_seg = 0;
    // cross-GPU exchange of the K0 words (rdr_set_exchange): peer-mapped buffers of 2 x world x XCHG_STRIDE words per rank
    int xchg_world = 0, xchg_rank = 0, xchg_parity = 0;
    void *xchg_bufs[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {

int fail(rdr_handle_t h, int code, const std::string &msg) {
    g_last_error = msg;
    if (h) h->err = msg;
    return code;
}

#define CUDA_TRY(h, expr)                                                                                         \
    do {                                                                                                          \
        cudaError_t _e = (expr);                                                                                  \
        if (_e != cudaSuccess)                                                                                    \
            return fail(h, RDR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                                             std::to_string(__LINE__) + ")");                                     \
    } while (0)

#define CHECK_ARG(h, cond, msg) \
    do {                        \
        if (!(cond)) return fail(h, RDR_ERR_INVALID, msg); \
    } while (0)

// occupancy variant of the ray kernels: __launch_bounds__(128, minb); overridable for tuning runs
inline int tune_minb(const char *env, int dflt) {
    const char *v = getenv(env);
    const int m = v ? atoi(v) : dflt;
    return (m == 2 || m == 3 || m == 4 || m == 5 || m == 6 || m == 8) ? m : dflt;
}

inline int grid_for(int64_t n, int block, int sm_count, int per_sm) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)sm_count * per_sm;
    return (int)std::max<int64_t>(1, std::min(need, cap));
}

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// ------------------------------------------------------------------------------------------------
// cube staging
// ------------------------------------------------------------------------------------------------
// src fields in caller layout -> float2 (wet, hydro) [ny][nx][nz] with ascending axes
__global__ void k_gather_fields(const float *__restrict__ wet, const float *__restrict__ hydro, float2 *__restrict__ dst, int ny,
                                int nx, int nz, int layout, int flip_y, int flip_x, int flip_z) {
    const int64_t total = (int64_t)ny * nx * nz;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int iz = (int)(i % nz);
        const int ix = (int)((i / nz) % nx);
        const int iy = (int)(i / ((int64_t)nz * nx));
        const int sy = flip_y ? ny - 1 - iy : iy, sx = flip_x ? nx - 1 - ix : ix, sz = flip_z ? nz - 1 - iz : iz;
        const int64_t s = layout == RDR_LAYOUT_ZYX ? ((int64_t)sz * ny + sy) * nx + sx : ((int64_t)sy * nx + sx) * nz + sz;
        dst[i] = make_float2(wet[s], hydro[s]);
    }
}

// temporal blend exactly as the reference's float32 xarray arithmetic: fl32(w0)*a + fl32(w1)*b (cli/raider.py:817-819)
__global__ void k_blend_fields(float2 *__restrict__ a, const float2 *__restrict__ b, int64_t total, float w0, float w1) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 p = a[i], q = b[i];
        a[i] = make_float2(__fadd_rn(__fmul_rn(w0, p.x), __fmul_rn(w1, q.x)), __fadd_rn(__fmul_rn(w0, p.y), __fmul_rn(w1, q.y)));
    }
}

// float2 [ny][nx][nz] -> double4 cells [ny][nx][nz-1] = {f[z], f[z+1]} (exact promotion)
__global__ void k_pack_cells(const float2 *__restrict__ f, double4 *__restrict__ cells, float4 *__restrict__ cells32, int64_t ncol, int nz) {
    const int64_t total = ncol * (nz - 1);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t col = i / (nz - 1);
        const int iz = (int)(i % (nz - 1));
        const float2 a = f[col * nz + iz], b = f[col * nz + iz + 1];
        cells[i] = make_double4((double)a.x, (double)a.y, (double)b.x, (double)b.y);
        cells32[i] = make_float4(a.x, a.y, b.x, b.y);
    }
}

// float2 [ny][nx][nz] -> LerpCell [ny-1][nx-1][nz-1] for the fast integrators (fastpath.cuh): one thread per 32-byte quarter of
// a cell record = one pair (a_2q, a_2q+1) of multilinear coefficients of both fields
__global__ void k_pack_lerp(const float2 *__restrict__ f, double4 *__restrict__ out, int ny, int nx, int nz) {
    const int nzc = nz - 1;
    const int64_t total = (int64_t)(ny - 1) * (nx - 1) * nzc * 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int quarter = (int)(i & 3);
        const int64_t cell = i >> 2;
        const int iz = (int)(cell % nzc);
        const int ix = (int)((cell / nzc) % (nx - 1));
        const int iy = (int)(cell / ((int64_t)nzc * (nx - 1)));
        // corner columns (y, x) = 00, 01, 10, 11: value at z and its difference along z (exact: fp32 data in fp64)
        double w[4], dw[4], hh[4], dh[4];
#pragma unroll
        for (int cnr = 0; cnr < 4; ++cnr) {
            const int64_t col = (int64_t)(iy + (cnr >> 1)) * nx + (ix + (cnr & 1));
            const float2 a = f[col * nz + iz], b = f[col * nz + iz + 1];
            w[cnr] = (double)a.x; dw[cnr] = (double)b.x - (double)a.x;
            hh[cnr] = (double)a.y; dh[cnr] = (double)b.y - (double)a.y;
        }
        double4 q;
        if (quarter == 0) q = make_double4(w[0], hh[0], dw[0], dh[0]);
        else if (quarter == 1) q = make_double4(w[1] - w[0], hh[1] - hh[0], dw[1] - dw[0], dh[1] - dh[0]);
        else if (quarter == 2) q = make_double4(w[2] - w[0], hh[2] - hh[0], dw[2] - dw[0], dh[2] - dh[0]);
        else q = make_double4((w[3] - w[2]) - (w[1] - w[0]), (hh[3] - hh[2]) - (hh[1] - hh[0]), (dw[3] - dw[2]) - (dw[1] - dw[0]),
                              (dh[3] - dh[2]) - (dh[1] - dh[0]));
        out[i] = q;
    }
}

CubeView make_view(rdr_handle_t h) {
    CubeView c;
    c.cells = h->d_cells.as<double4>();
    c.cells32 = h->d_cells32.as<float4>();
    const double *nodes = h->d_axes.as<double>();
    const char *tabs = h->d_tabs.as<char>();
    const std::vector<double> *v[3] = {&h->ys, &h->xs, &h->zs};
    Axis *ax[3] = {&c.ay, &c.ax, &c.az};
    size_t off = 0;
    for (int d = 0; d < 3; ++d) {
        Axis &a = *ax[d];
        a.g = nodes + off;
        off += v[d]->size();
        a.cell = reinterpret_cast<const double4 *>(tabs + h->tab_cell_off[d]);
        a.bin = reinterpret_cast<const unsigned short *>(tabs + h->tab_bin_off[d]);
        a.n = (int)v[d]->size();
        a.nbin = h->tab_nbin[d];
        a.g_first = v[d]->front();
        a.g_last = v[d]->back();
        a.inv_bw = (double)a.nbin / (a.g_last - a.g_first);
        const double dmean = (a.g_last - a.g_first) / (double)(a.n - 1);
        a.inv_d = 1.0 / dmean;
        a.uniform = 1;
        for (size_t i = 0; i < v[d]->size(); ++i)  // every node within a quarter cell of its uniform position: the guess is off by <= 1
            if (fabs((*v[d])[i] - (a.g_first + dmean * (double)i)) > 0.25 * dmean) a.uniform = 0;
        a.d = (*v[d])[1] - (*v[d])[0];
        a.inv_dx = 1.0 / a.d;
        a.exact_uniform = a.uniform;
        for (size_t i = 0; i < v[d]->size() && a.exact_uniform; ++i)  // bit-for-bit: the kernel recomputes the nodes with this very fma
            if (fma((double)i, a.d, a.g_first) != (*v[d])[i]) a.exact_uniform = 0;
        for (size_t i = 0; i + 1 < v[d]->size() && a.exact_uniform; ++i)
            if ((*v[d])[i] + a.d != (*v[d])[i + 1] || (*v[d])[i + 1] - (*v[d])[i] != a.d) a.exact_uniform = 0;
    }
    Axis32 *a32[3] = {&c.fy, &c.fx, &c.fz};
    for (int d = 0; d < 3; ++d) {
        const Axis &a = *ax[d];
        Axis32 &f = *a32[d];
        f.rec = reinterpret_cast<const float4 *>(tabs + h->tab_rec32_off[d]);
        f.bin = a.bin;
        f.n = a.n;
        f.nbin = a.nbin;
        f.uniform = a.uniform;
        f.g_first = (float)a.g_first;
        f.inv_d = (float)a.inv_d;
        f.inv_bw = (float)a.inv_bw;
        f.first_cmp = (float)a.g_first;
        if ((double)f.first_cmp < a.g_first) f.first_cmp = nextafterf(f.first_cmp, INFINITY);   // RU32(g[0])
        f.last_cmp = (float)a.g_last;
        if ((double)f.last_cmp > a.g_last) f.last_cmp = nextafterf(f.last_cmp, -INFINITY);      // RD32(g[n-1])
        f.d = (float)a.d;
        f.exact32 = a.exact_uniform && (double)f.d == a.d;
        for (size_t i = 0; i < v[d]->size() && f.exact32; ++i)
            if ((double)fmaf((float)i, f.d, f.g_first) != (*v[d])[i]) f.exact32 = 0;
    }
    c.crs_kind = h->crs_kind;
    c.lcc = {h->crs[0], h->crs[1], h->crs[2], h->crs[3], h->crs[4], h->crs[5], h->crs[6]};
    return c;
}

// the fast integrator's view: needs a geographic cube whose horizontal axes are uniform to 1e-9 of a cell, so that the cell
// coordinate (v - first) / d stands for the node search (t differs from the node-based one by < 1e-9: micrometres on the ground)
bool make_fast_cube(rdr_handle_t h, FastCube &c) {
    if (h->crs_kind != RDR_CRS_GEOGRAPHIC && h->crs_kind != RDR_CRS_LCC_SPHERE) return false;
    const std::vector<double> *v[2] = {&h->ys, &h->xs};
    double inv[2], c0[2];
    for (int d = 0; d < 2; ++d) {
        const std::vector<double> &g = *v[d];
        const double dd = (g.back() - g.front()) / (double)(g.size() - 1);
        for (size_t i = 0; i < g.size(); ++i)
            if (!(fabs(g[i] - (g.front() + dd * (double)i)) <= 1e-9 * dd)) return false;
        inv[d] = 1.0 / dd;
        c0[d] = -g.front() * inv[d];
    }
    c.cells = h->d_lerp.as<LerpCell>();
    c.ny = (int)h->ny; c.nx = (int)h->nx; c.nzc = (int)h->nz - 1;
    c.y_inv = inv[0]; c.y_c0 = c0[0]; c.x_inv = inv[1]; c.x_c0 = c0[1];
    c.crs_kind = h->crs_kind;
    c.lcc = {h->crs[0], h->crs[1], h->crs[2], h->crs[3], h->crs[4], h->crs[5], h->crs[6]};
    return true;
}

// interval records + first-guess bins for the three axes (see sampler.cuh)
int build_axis_tables(rdr_handle_t h) {
    const std::vector<double> *v[3] = {&h->ys, &h->xs, &h->zs};
    std::vector<char> blob;
    for (int d = 0; d < 3; ++d) {
        const std::vector<double> &g = *v[d];
        const int n = (int)g.size();
        h->tab_cell_off[d] = blob.size();
        std::vector<double> rec(4 * (size_t)(n - 1));
        for (int i = 0; i + 1 < n; ++i) {
            const double dd = g[i + 1] - g[i];
            rec[4 * i] = g[i]; rec[4 * i + 1] = g[i + 1]; rec[4 * i + 2] = dd; rec[4 * i + 3] = 1.0 / dd;
        }
        blob.insert(blob.end(), reinterpret_cast<char *>(rec.data()), reinterpret_cast<char *>(rec.data() + rec.size()));
        double dmin = g[1] - g[0];
        for (int i = 1; i + 1 < n; ++i) dmin = std::min(dmin, g[i + 1] - g[i]);
        // bin width <= the narrowest interval: a bin contains at most one node, so the guess is at most one step short
        const int nbin = (int)std::min(65536.0, std::max(64.0, ceil((g.back() - g.front()) / dmin) + 1.0));
        h->tab_nbin[d] = nbin;
        h->tab_bin_off[d] = blob.size();
        std::vector<unsigned short> bins(nbin);
        const double bw = (g.back() - g.front()) / nbin;
        int i = 0;
        for (int b = 0; b < nbin; ++b) {
            const double start = g.front() + b * bw;
            while (i < n - 2 && g[i + 1] <= start) ++i;
            bins[b] = (unsigned short)i;  // interval holding the start of bin b; the device verifies against the nodes either way
        }
        blob.insert(blob.end(), reinterpret_cast<char *>(bins.data()), reinterpret_cast<char *>(bins.data() + bins.size()));
        while (blob.size() % 32) blob.push_back(0);
        // fp32 interval records of the fp32 sampler tier (sampler.cuh: Axis32)
        h->tab_rec32_off[d] = blob.size();
        std::vector<float> rec32(4 * (size_t)(n - 1));
        for (int k = 0; k + 1 < n; ++k) {
            const float lo_hi = (float)g[k];
            rec32[4 * k] = lo_hi;
            rec32[4 * k + 1] = (float)(g[k] - (double)lo_hi);
            float hi = (float)g[k + 1];
            if ((double)hi < g[k + 1]) hi = nextafterf(hi, INFINITY);  // RU32: v >= hi <=> v >= g[k+1] for every fp32 v
            rec32[4 * k + 2] = hi;
            rec32[4 * k + 3] = (float)(1.0 / (g[k + 1] - g[k]));
        }
        blob.insert(blob.end(), reinterpret_cast<char *>(rec32.data()), reinterpret_cast<char *>(rec32.data() + rec32.size()));
        while (blob.size() % 32) blob.push_back(0);
    }
    CUDA_TRY(h, h->d_tabs.reserve(blob.size()));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_tabs.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return RDR_OK;
}

// ------------------------------------------------------------------------------------------------
// K2: unfused trilinear sampler.  One thread per point; the [n][3] AoS points are read with coalesced
// 16-byte loads through shared memory (3 x 16 B per 2 points), the two outputs are written as plain
// coalesced fp64/fp32 stores.  Algorithmic traffic: 40 B/point (f64) or 20 B/point (f32).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void sample_any(const CubeView &c, int semantics, double y, double x, double z, double &vw, double &vh) {
    if (semantics == RDR_SEM_SCIPY) {
        int iy = -1, ix = -1, iz = -1;
        sample_scipy<GUESS_BINS, GUESS_BINS>(c, y, x, z, iy, ix, iz, vw, vh);
        return;
    }
    // RAiDER.interpolate rules on the staged fp32 cube (values promoted to fp64)
    const Axis *ax[3] = {&c.ay, &c.ax, &c.az};
    const double v[3] = {y, x, z};
    int hi[3];
    for (int d = 0; d < 3; ++d) {
        int k = bisect_left(ax[d]->g, ax[d]->n, v[d]);
        if (semantics == RDR_SEM_RAIDER_FILL) {
            if (k < 1 || k > ax[d]->n - 1) {
                vw = vh = qnan();
                return;
            }
        } else {
            k = k < 1 ? 1 : (k > ax[d]->n - 1 ? ax[d]->n - 1 : k);
        }
        hi[d] = k;
    }
    double lo_d[3], hi_d[3], vol = 1.0;
    for (int d = 0; d < 3; ++d) {
        const double g0 = __ldg(ax[d]->g + hi[d] - 1), g1 = __ldg(ax[d]->g + hi[d]);
        lo_d[d] = v[d] - g0;
        hi_d[d] = g1 - v[d];
        vol = d == 0 ? (g1 - g0) : __dmul_rn(vol, g1 - g0);
    }
    const int nzc = c.az.n - 1;
    const double4 *p = c.cells + ((size_t)(hi[0] - 1) * c.ax.n + (hi[1] - 1)) * nzc + (hi[2] - 1);
    const double4 c00 = ld_cell(p), c01 = ld_cell(p + nzc), c10 = ld_cell(p + (size_t)c.ax.n * nzc), c11 = ld_cell(p + (size_t)c.ax.n * nzc + nzc);
    vw = trilinear_raider(c00.x, c00.z, c01.x, c01.z, c10.x, c10.z, c11.x, c11.z, lo_d[0], hi_d[0], lo_d[1], hi_d[1], lo_d[2], hi_d[2], vol);
    vh = trilinear_raider(c00.y, c00.w, c01.y, c01.w, c10.y, c10.w, c11.y, c11.w, lo_d[0], hi_d[0], lo_d[1], hi_d[1], lo_d[2], hi_d[2], vol);
}

// ---- mbarrier / TMA-bulk helpers (sm_90+ PTX; on sm_100a these become SYNCS.* and UBLKCP) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on `bar` (cp.async.bulk = the TMA engine without a tensor map)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// K2, scipy semantics, streaming form: the [n][3] point stream is pulled into a 3-deep shared-memory ring by TMA bulk copies
// (one elected thread issues, an mbarrier per stage counts the bytes), so the HBM reads of tile i+2 overlap the arithmetic of
// tile i; each thread samples two points of a tile (two independent dependency chains), outputs are plain coalesced stores.
constexpr int K2_THREADS = 128, K2_STAGES = 3;

template <typename T, int MXY, int K2_PPT>
__global__ void __launch_bounds__(K2_THREADS) k_sample_stream(const CubeView c, const T *__restrict__ pts, int64_t n, T *__restrict__ out_wet,
                                                            T *__restrict__ out_hydro) {
    constexpr int K2_TILE = K2_THREADS * K2_PPT;
    constexpr uint32_t TILE_BYTES = K2_TILE * 3 * sizeof(T);
    extern __shared__ __align__(128) unsigned char k2_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(k2_smem + K2_STAGES * TILE_BYTES);
    const int64_t ntiles = n / K2_TILE;  // full tiles go through the ring; the ragged tail is handled below with plain loads
    if (threadIdx.x == 0) {
        for (int s = 0; s < K2_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < K2_STAGES; ++s) {
            const int64_t tile = blockIdx.x + (int64_t)s * gridDim.x;
            if (tile < ntiles) {
                mbar_expect_tx(&full[s], TILE_BYTES);
                tma_load_1d(k2_smem + s * TILE_BYTES, pts + tile * K2_TILE * 3, TILE_BYTES, &full[s]);
            }
        }
    }
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % K2_STAGES;
        mbar_wait(&full[s], (uint32_t)(it / K2_STAGES) & 1u);
        const T *tp = reinterpret_cast<const T *>(k2_smem + s * TILE_BYTES);
        double y[K2_PPT], x[K2_PPT], z[K2_PPT], w[K2_PPT], hh[K2_PPT];
#pragma unroll
        for (int p = 0; p < K2_PPT; ++p) {
            const int q = threadIdx.x + p * K2_THREADS;
            y[p] = (double)tp[3 * q];
            x[p] = (double)tp[3 * q + 1];
            z[p] = (double)tp[3 * q + 2];
        }
        sample_scipy_batch<K2_PPT, MXY, GUESS_BINS>(c, y, x, z, w, hh);
        const int64_t base = tile * K2_TILE;
#pragma unroll
        for (int p = 0; p < K2_PPT; ++p) {
            __stcs(out_wet + base + threadIdx.x + p * K2_THREADS, (T)w[p]);
            __stcs(out_hydro + base + threadIdx.x + p * K2_THREADS, (T)hh[p]);
        }
        __syncthreads();  // every thread has read stage s: it can be refilled
        if (threadIdx.x == 0) {
            const int64_t next = tile + (int64_t)K2_STAGES * gridDim.x;
            if (next < ntiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&full[s], TILE_BYTES);
                tma_load_1d(k2_smem + s * TILE_BYTES, pts + next * K2_TILE * 3, TILE_BYTES, &full[s]);
            }
        }
    }
    // ragged tail (< K2_TILE points): plain loads, first block only
    if (blockIdx.x == 0) {
        for (int64_t i = ntiles * K2_TILE + threadIdx.x; i < n; i += K2_THREADS) {
            double w, hh;
            int iy = -1, ix = -1, iz = -1;
            sample_scipy<MXY, GUESS_BINS>(c, (double)pts[3 * i], (double)pts[3 * i + 1], (double)pts[3 * i + 2], iy, ix, iz, w, hh);
            out_wet[i] = (T)w;
            out_hydro[i] = (T)hh;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2, fp32 tier: the same point stream (TMA-bulk ring) with fp32 coordinates in, fp32 values out and fp32 ARITHMETIC -- the
// 1e-3 m tier of north_star.  20 B per point: at the HBM roofline a warp of 32 points has ~110 issue slots, which the fp64
// arithmetic of k_sample_stream (188 instructions per point, half-rate pipe) cannot meet; this form needs ~85 fp32 / integer
// instructions.  Semantics are scipy's: NaN outside the closed box (decided exactly on the fp32 inputs, see Axis32), NaN in ->
// NaN out, last node inclusive, NaN corners poison; values agree with scipy evaluated at the same fp32 points to ~1e-6 of
// the field's range (fp32 rounding of t and of the lerps), far inside the tier's tolerance.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int guess32(const Axis32 &a, float v) {
    if (a.uniform) {  // floor((v - g0) / d) by directed rounding against 2^23 + 2^22: no F2I
        const float s = __fadd_rd((v - a.g_first) * a.inv_d, 12582912.0f);
        return min(max(__float_as_int(s) - 0x4b400000, 0), a.n - 2);
    }
    const int b = (int)((v - a.g_first) * a.inv_bw);
    return (int)__ldg(a.bin + min(max(b, 0), a.nbin - 1));
}

__device__ __forceinline__ float locate32(const Axis32 &a, float v, int &i, float4 r) {
    // num = v - g[i] rounded once; its sign is exact (v - lo_hi is exact and a multiple of the ulp, |lo_lo| < ulp / 2), and
    // v >= hi <=> v >= g[i+1] exactly (hi = RU32(g[i+1])): the interval is scipy's, not a neighbour within rounding of a node
    float num = (v - r.x) - r.y;
    if (num < 0.0f || v >= r.z) {  // guess one off (rounding of the guess, node hit, the inclusive last node); clamped for OOB / NaN
        const int last = a.n - 2;
        while (num < 0.0f && i > 0) {
            r = __ldg(a.rec + --i);
            num = (v - r.x) - r.y;
        }
        while (v >= r.z && i < last) {
            r = __ldg(a.rec + ++i);
            num = (v - r.x) - r.y;
        }
    }
    return num * r.w;
}

// exact32 axis: interval and fraction without a table -- floor by directed rounding, node = fmaf(i, d, g0) exactly
__device__ __forceinline__ float locate32_exact(const Axis32 &a, float v, int &i) {
    const float s = __fadd_rd((v - a.g_first) * a.inv_d, 12582912.0f);
    const int raw = __float_as_int(s) - 0x4b400000;
    i = min(max(raw, 0), a.n - 2);
    float lo = fmaf((float)i, a.d, a.g_first);
    if (v < lo || v >= lo + a.d) {  // the product rounded across a node, the inclusive last node, out of bounds
        const int last = a.n - 2;
        while (v < lo && i > 0) lo = fmaf((float)(--i), a.d, a.g_first);
        while (v >= lo + a.d && i < last) lo = fmaf((float)(++i), a.d, a.g_first);
    }
    return (v - lo) * a.inv_d;
}

template <int PPT, bool XY_EXACT>
__global__ void __launch_bounds__(K2_THREADS) k_sample_stream_f32(const CubeView c, const float *__restrict__ pts, int64_t n,
                                                                float *__restrict__ out_wet, float *__restrict__ out_hydro) {
    constexpr int TILE = K2_THREADS * PPT;
    constexpr uint32_t TILE_BYTES = TILE * 3 * sizeof(float);
    extern __shared__ __align__(128) unsigned char k2_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(k2_smem + K2_STAGES * TILE_BYTES);
    const int64_t ntiles = n / TILE;
    if (threadIdx.x == 0) {
        for (int s = 0; s < K2_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < K2_STAGES; ++s) {
            const int64_t tile = blockIdx.x + (int64_t)s * gridDim.x;
            if (tile < ntiles) {
                mbar_expect_tx(&full[s], TILE_BYTES);
                tma_load_1d(k2_smem + s * TILE_BYTES, pts + tile * TILE * 3, TILE_BYTES, &full[s]);
            }
        }
    }
    const int nzc = c.fz.n - 1;
    const unsigned row = (unsigned)c.fx.n * (unsigned)nzc;
    const float qnanf = __int_as_float(0x7fc00000);
    auto sample = [&](const float (&y)[PPT], const float (&x)[PPT], const float (&z)[PPT], float (&vw)[PPT], float (&vh)[PPT]) {
        int iy[PPT], ix[PPT], iz[PPT];
        float4 ry[PPT], rx[PPT], rz[PPT];
#pragma unroll
        for (int p = 0; p < PPT; ++p) {  // guesses, then all interval records in flight
            if (!XY_EXACT) {
                iy[p] = guess32(c.fy, y[p]);
                ix[p] = guess32(c.fx, x[p]);
            }
            iz[p] = guess32(c.fz, z[p]);
        }
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            if (!XY_EXACT) {
                ry[p] = __ldg(c.fy.rec + iy[p]);
                rx[p] = __ldg(c.fx.rec + ix[p]);
            }
            rz[p] = __ldg(c.fz.rec + iz[p]);
        }
        float ty[PPT], tx[PPT], tz[PPT];
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            ty[p] = XY_EXACT ? locate32_exact(c.fy, y[p], iy[p]) : locate32(c.fy, y[p], iy[p], ry[p]);
            tx[p] = XY_EXACT ? locate32_exact(c.fx, x[p], ix[p]) : locate32(c.fx, x[p], ix[p], rx[p]);
            tz[p] = locate32(c.fz, z[p], iz[p], rz[p]);
        }
        float4 c00[PPT], c01[PPT], c10[PPT], c11[PPT];
#pragma unroll
        for (int p = 0; p < PPT; ++p) {  // 4 x LDG.128: the z-pair of both fields at the four corner columns
            const float4 *q = c.cells32 + ((unsigned)iy[p] * row + (unsigned)ix[p] * (unsigned)nzc + (unsigned)iz[p]);
            c00[p] = __ldg(q);
            c01[p] = __ldg(q + nzc);
            c10[p] = __ldg(q + row);
            c11[p] = __ldg(q + row + nzc);
        }
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const float w00 = fmaf(tz[p], c00[p].z - c00[p].x, c00[p].x), h00 = fmaf(tz[p], c00[p].w - c00[p].y, c00[p].y);
            const float w01 = fmaf(tz[p], c01[p].z - c01[p].x, c01[p].x), h01 = fmaf(tz[p], c01[p].w - c01[p].y, c01[p].y);
            const float w10 = fmaf(tz[p], c10[p].z - c10[p].x, c10[p].x), h10 = fmaf(tz[p], c10[p].w - c10[p].y, c10[p].y);
            const float w11 = fmaf(tz[p], c11[p].z - c11[p].x, c11[p].x), h11 = fmaf(tz[p], c11[p].w - c11[p].y, c11[p].y);
            const float w0 = fmaf(tx[p], w01 - w00, w00), h0 = fmaf(tx[p], h01 - h00, h00);
            const float w1 = fmaf(tx[p], w11 - w10, w10), h1 = fmaf(tx[p], h11 - h10, h10);
            const bool inb = (y[p] >= c.fy.first_cmp) & (y[p] <= c.fy.last_cmp) & (x[p] >= c.fx.first_cmp) & (x[p] <= c.fx.last_cmp) &
                             (z[p] >= c.fz.first_cmp) & (z[p] <= c.fz.last_cmp);  // false for NaN coordinates too
            vw[p] = inb ? fmaf(ty[p], w1 - w0, w0) : qnanf;
            vh[p] = inb ? fmaf(ty[p], h1 - h0, h0) : qnanf;
        }
    };
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % K2_STAGES;
        mbar_wait(&full[s], (uint32_t)(it / K2_STAGES) & 1u);
        const float *tp = reinterpret_cast<const float *>(k2_smem + s * TILE_BYTES);
        float y[PPT], x[PPT], z[PPT], w[PPT], hh[PPT];
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int q = threadIdx.x + p * K2_THREADS;
            y[p] = tp[3 * q];
            x[p] = tp[3 * q + 1];
            z[p] = tp[3 * q + 2];
        }
        sample(y, x, z, w, hh);
        const int64_t base = tile * TILE;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            __stcs(out_wet + base + threadIdx.x + p * K2_THREADS, w[p]);
            __stcs(out_hydro + base + threadIdx.x + p * K2_THREADS, hh[p]);
        }
        __syncthreads();  // every thread has read stage s: it can be refilled
        if (threadIdx.x == 0) {
            const int64_t next = tile + (int64_t)K2_STAGES * gridDim.x;
            if (next < ntiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&full[s], TILE_BYTES);
                tma_load_1d(k2_smem + s * TILE_BYTES, pts + next * TILE * 3, TILE_BYTES, &full[s]);
            }
        }
    }
    if (blockIdx.x == 0) {  // ragged tail (< TILE points): the same arithmetic on plain loads, one point at a time
        for (int64_t i = ntiles * TILE + threadIdx.x; i < n; i += K2_THREADS) {
            float y[PPT], x[PPT], z[PPT], w[PPT], hh[PPT];
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                y[p] = pts[3 * i];
                x[p] = pts[3 * i + 1];
                z[p] = pts[3 * i + 2];
            }
            sample(y, x, z, w, hh);
            out_wet[i] = w[0];
            out_hydro[i] = hh[0];
        }
    }
}

template <typename T, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_sample_points(const CubeView c, const T *__restrict__ pts, int64_t n, T *__restrict__ out_wet,
                                                         T *__restrict__ out_hydro, int semantics) {
    __shared__ __align__(16) T tile[BLOCK * 3];
    const int64_t ntiles = (n + BLOCK - 1) / BLOCK;
    for (int64_t tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
        const int64_t base = tile_i * BLOCK;
        const int cnt = (int)min((int64_t)BLOCK, n - base);
        // coalesced 16-byte loads of this tile's cnt*3 scalars
        constexpr int VEC = 16 / sizeof(T);
        const T *src = pts + base * 3;
        const int nscal = cnt * 3;
        if (cnt == BLOCK && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
            float4 *d4 = reinterpret_cast<float4 *>(tile);
            for (int i = threadIdx.x; i < BLOCK * 3 / VEC; i += BLOCK) d4[i] = __ldcs(s4 + i);
        } else {
            for (int i = threadIdx.x; i < nscal; i += BLOCK) tile[i] = src[i];
        }
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            const double y = (double)tile[threadIdx.x * 3 + 0], x = (double)tile[threadIdx.x * 3 + 1], z = (double)tile[threadIdx.x * 3 + 2];
            double vw, vh;
            sample_any<T>(c, semantics, y, x, z, vw, vh);
            __stcs(out_wet + base + threadIdx.x, (T)vw);
            __stcs(out_hydro + base + threadIdx.x, (T)vh);
        }
        __syncthreads();
    }
}

// _build_cube for one height: points generated on device from the query axes (delay.py:211)
// (zpts[nh]: all output heights of _build_cube in one launch, out[nh][ny][nx])
__global__ void k_sample_grid(const CubeView c, const double *__restrict__ xpts, int nx, const double *__restrict__ ypts, int ny,
                              const double *__restrict__ zpts, int nh, double *__restrict__ out_wet, double *__restrict__ out_hydro) {
    const int64_t plane = (int64_t)ny * nx, n = plane * nh;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = r % plane;
        const int j = (int)(q / nx), i = (int)(q % nx);
        double vw, vh;
        int iy = -1, ix = -1, iz = -1;
        sample_scipy<GUESS_BINS, GUESS_BINS>(c, __ldg(ypts + j), __ldg(xpts + i), __ldg(zpts + r / plane), iy, ix, iz, vw, vh);
        out_wet[r] = vw;
        out_hydro[r] = vh;
    }
}

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

// ------------------------------------------------------------------------------------------------
// The step plan, built ON THE DEVICE between K0 and K3 (k_plan, one CTA): everything the host used to derive from K0's maxima
// -- nParts = ceil(max / MAX_SEGMENT_LENGTH) + 1 (delay.py:283), the per-layer records, the spans of the polynomial
// integrator, the whole-raster clamp predicate (delay.py:306-307), the all-NaN check (delay.py:279) -- so that K0 -> K3 needs
// no host round trip (no cudaStreamSynchronize, no D2H, no collective through the host).  Across GPUs every rank stores its
// K maxima + 3 counters into a slot of every peer's exchange buffer (k_publish: peer-mapped symmetric memory over NVLink),
// the caller orders the ranks with one signal-pad barrier on the stream, and k_plan takes MAX / SUM over the slots: the
// all-reduce of SURVEY 8(e), without NCCL and without the host.  The host reads the plan back after the step
// (rdr_trace_result), when it synchronises for the results anyway.
// ------------------------------------------------------------------------------------------------
constexpr int THIN_TD = 8;   // along-ray distances in flight per thread in k_ray_integrate_thin (ring depth, power of two)
constexpr int LERP_PAD = 8;  // records of padding behind the cell-record array (prefetch distance bound of k_ray_integrate_thin)
constexpr int XCHG_STRIDE = MAX_LAYERS + 8;  // words per rank slot: maxima bits [K] | #NaN rays | #first sample below | #rays | K3's #first sample below | #last sample above
// |maxlen / S - nearest integer| below which nParts is declared a knife edge: the default K0 reproduces the reference's maxima
// to ~1e-8 m (polynomials of h(t) and of the layer tops), the exact form to ~1e-9 m; 1e-6 of a segment is 1 mm at the default 1000 m
constexpr double KNIFE_EPS = 1.0e-6;

// written into `part` by the quadrature kernel for a ray it put on the fix list (a NaN no arithmetic produces)
constexpr long long PART_FLAGGED = 0x7ff8dead00000001LL;

struct DevPlan {
    int status;           // RDR_PLAN_* bits seen
    int blocked;          // status & block_mask: non-zero -> the integration kernels do nothing (the host redoes / raises)
    int K, nspan;
    int k_split;          // layers [0, k_split): thin-layer kernel, [k_split, K): quadrature kernel
    int span_split;       // spans  [0, span_split) belong to the thin part
    int clamp_low_first;  // delay.py:306-307 decided from K0's global count
    int clamp_high_last;  // delay.py:310-311 for the very last sample (top of the top layer), decided from K0's global count
    int knife_layer;      // a layer whose maxlen / S is within KNIFE_EPS of an integer (-1: none)
    long long n_rays, n_nan, n_below, n_above;  // global counters
    double longest_span;
    double maxlen[MAX_LAYERS];
    int nparts[MAX_LAYERS];
    int layer_cell[MAX_LAYERS];
    int span_end[MAX_LAYERS];
    LayerRec layers[MAX_LAYERS];
};

__global__ void __launch_bounds__(256) k_plan(const unsigned long long *__restrict__ slots, int world, int stride, int K,
                                              const int *__restrict__ layer_cell, const double *__restrict__ zs, int nz, double max_seg,
                                              double span_max, int thin_min, int thin_absorb, int force_clamp, int block_mask,
                                              DevPlan *__restrict__ P, unsigned long long *__restrict__ k3_counters) {
    __shared__ int s_status, s_knife;
    __shared__ int s_np[MAX_LAYERS];
    __shared__ double s_len[MAX_LAYERS];
    if (threadIdx.x == 0) {
        s_status = 0;
        s_knife = -1;
    }
    if (threadIdx.x < 6) k3_counters[threadIdx.x] = 0ull;  // 4 integration counters + staged / unstaged CTA passes of the thin kernel
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        unsigned long long m = 0ull;
        for (int q = 0; q < world; ++q) m = max(m, slots[(size_t)q * stride + k]);  // MAX over ranks on the IEEE bits (lengths >= 0)
        const double len = __longlong_as_double((long long)m);
        const double x = len / max_seg;
        const double qn = ceil(x);
        int st = 0, np = 2;
        if (qn == qn && qn < 1.0e7) {
            np = (int)qn + 1;  // nParts = ceil(max / MAX_SEGMENT_LENGTH).astype(int) + 1   (delay.py:283)
            if (np < 2) np = 2;  // a zero-length layer would divide by zero in the reference (np.linspace(0, 1, 1)); keep 2
        } else {
            st |= RDR_PLAN_ABSURD;
        }
        const double fr = x - floor(x);
        if (len > 0.0 && (fr < KNIFE_EPS || fr > 1.0 - KNIFE_EPS)) {
            st |= RDR_PLAN_KNIFE_EDGE;
            atomicMax(&s_knife, k);
        }
        const int iz = layer_cell[k];
        const double z_lo = zs[iz], z_hi = zs[iz + 1];
        LayerRec r;
        r.z_lo = z_lo;
        r.inv_dz = 1.0 / (z_hi - z_lo);
        r.neg_zlo_inv = -z_lo * r.inv_dz;
        r.h_lo = iz == 0 ? z_lo : z_lo - LAYER_TOL;                                             // below the first node: NaN rule
        r.h_hi = iz == nz - 2 ? __longlong_as_double(__double_as_longlong(z_hi) + (z_hi >= 0 ? 1 : -1)) : z_hi + LAYER_TOL;  // the last node is inclusive
        r.step = 1.0 / (double)(np - 1);
        r.np = np;
        r.iz = iz;
        P->layers[k] = r;
        P->maxlen[k] = len;
        P->nparts[k] = np;
        P->layer_cell[k] = iz;
        s_np[k] = np;
        s_len[k] = len;
        if (st) atomicOr(&s_status, st);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long n_nan = 0, n_below = 0, n_rays = 0, n_above = 0;
        for (int q = 0; q < world; ++q) {
            n_nan += (long long)slots[(size_t)q * stride + K];
            n_below += (long long)slots[(size_t)q * stride + K + 1];
            n_above += (long long)slots[(size_t)q * stride + K + 4];
            n_rays += (long long)slots[(size_t)q * stride + K + 2];
        }
        int st = s_status;
        if (n_nan == n_rays) st |= RDR_PLAN_ALL_NAN;  // np.isnan(ray_lengths).all() over the WHOLE raster (delay.py:279)
        // thin-layer part: the leading run of layers with <= 3 samples (the 145-node tables at 1000 m: ~115 of 139 layers)
        int n_thin = 0, last_thin = -1;
        for (int k = 0; k < K; ++k)
            if (s_np[k] <= 3) {
                ++n_thin;
                last_thin = k;
            }
        int k_split = 0;
        if (thin_min < 0) {
            k_split = K;  // unified mode: the staged kernel takes every layer (closed-form sums for the thick ones included)
        } else if (thin_min > 0 && n_thin >= thin_min) {
            // cut where the thin layers stop dominating: the longest prefix in which >= 3/4 of the layers are thin
            int seen = 0;
            for (int k = 0; k <= last_thin; ++k) {
                seen += s_np[k] <= 3;
                if (s_np[k] <= 3 && 4 * seen >= 3 * (k + 1)) k_split = k + 1;
            }
            if (k_split < thin_min) k_split = 0;
            // a short thick tail (the top of the 145-node tables: 15 layers of 4 .. 7 samples) is cheaper sample by sample in the
            // thin-layer kernel than as a second pass of every ray through the quadrature kernel (per-ray set-up, partial sums
            // through HBM): absorb it when it holds at most `thin_absorb` samples beyond its layer tops
            if (k_split > 0 && k_split < K) {
                int extra = 0;
                for (int k = k_split; k < K; ++k) extra += s_np[k] - 1;
                if (extra <= thin_absorb) k_split = K;
            }
        }
        // spans of the polynomial integrators: whole layers, greedy, <= span_max metres of the longest ray, cut at k_split
        int nspan = 0, span_split = 0;
        double acc = 0.0, longest = 0.0;
        for (int k = 0; k < K; ++k) {
            if (k > 0 && (acc + s_len[k] > span_max || k == k_split)) {
                P->span_end[nspan++] = k;
                longest = fmax(longest, acc);
                acc = 0.0;
                if (k == k_split) span_split = nspan;
            }
            acc += s_len[k];
        }
        P->span_end[nspan++] = K;
        longest = fmax(longest, acc);
        if (k_split == K) span_split = nspan;
        // a single layer longer than 2 spans (48 km at the default) would stretch the cubic's error bound (T^4) by > 16
        if (longest > 2.0 * span_max) st |= RDR_PLAN_SPAN_TOO_LONG;
        P->status = st;
        P->blocked = st & block_mask;
        P->K = K;
        P->nspan = nspan;
        P->k_split = k_split;
        P->span_split = span_split;
        // force_clamp < 0: both predicates from K0's global counts; otherwise bit 0 = the lower clamp's value, bit 1 = upper clamp forced on,
        // bit 2 = upper clamp forced off (neither: from the count)
        P->clamp_low_first = force_clamp >= 0 ? (force_clamp & 1) : (n_below == n_rays);
        P->clamp_high_last = (force_clamp >= 0 && (force_clamp & 2)) ? 1 : (force_clamp >= 0 && (force_clamp & 4)) ? 0 : (n_above == n_rays);
        P->knife_layer = s_knife;
        P->n_rays = n_rays;
        P->n_nan = n_nan;
        P->n_below = n_below;
        P->n_above = n_above;
        P->longest_span = longest;
    }
}

// every rank's K0 words -> slot `rank` of every peer's exchange buffer (and of its own)
__global__ void k_publish(const unsigned long long *__restrict__ src, int nwords, int dst_off, const PeerOut dst) {
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) {
        const unsigned long long v = src[i];
        for (int p = 0; p < dst.n; ++p) static_cast<unsigned long long *>(dst.wet[p])[dst_off + i] = v;
    }
}

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

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
RDR_API int rdr_abi_version(void) { return RDR_ABI_VERSION; }

RDR_API int rdr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

RDR_API const char *rdr_last_error(rdr_handle_t h) { return h ? h->err.c_str() : g_last_error.c_str(); }

RDR_API int rdr_create(int device, rdr_handle_t *out) {
    if (!out) return fail(nullptr, RDR_ERR_INVALID, "rdr_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, RDR_ERR_CUDA, std::string("rdr_create: no CUDA device available (") + cudaGetErrorString(e) +
                                               "); libraider_b200 has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(nullptr, RDR_ERR_INVALID, "rdr_create: device index out of range");
    rdr_handle_t h = new rdr_handle_s();
    h->device = device;
    ScopedDevice sd(device);
    if ((e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) {
        delete h;
        return fail(nullptr, RDR_ERR_CUDA, std::string("cudaDeviceGetAttribute: ") + cudaGetErrorString(e));
    }
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete h;
        return fail(nullptr, RDR_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    h->own_stream = true;
    *out = h;
    return RDR_OK;
}

RDR_API int rdr_destroy(rdr_handle_t h) {
    if (!h) return RDR_OK;
    ScopedDevice sd(h->device);
    cudaStreamSynchronize(h->stream);
    for (DevBuf *b : {&h->d_axes, &h->d_tabs, &h->d_cells, &h->d_stage, &h->d_fields, &h->d_gx, &h->d_gy, &h->d_los, &h->d_plan, &h->d_t, &h->d_red,
                      &h->d_nparts, &h->d_out, &h->d_in, &h->d_lerp, &h->d_layers, &h->d_spans, &h->d_fix, &h->d_cells32, &h->d_orbit, &h->d_devplan,
                      &h->d_part})
        b->release();
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return RDR_OK;
}

RDR_API int rdr_set_stream(rdr_handle_t h, void *cuda_stream) {
    CHECK_ARG(h, h != nullptr, "rdr_set_stream: NULL handle");
    ScopedDevice sd(h->device);
    cudaStreamSynchronize(h->stream);
    if (cuda_stream == nullptr) {
        if (!h->own_stream) {
            CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
            h->own_stream = true;
        }
    } else {
        if (h->own_stream) cudaStreamDestroy(h->stream);
        h->stream = static_cast<cudaStream_t>(cuda_stream);
        h->own_stream = false;
    }
    return RDR_OK;
}

RDR_API int rdr_synchronize(rdr_handle_t h) {
    CHECK_ARG(h, h != nullptr, "rdr_synchronize: NULL handle");
    ScopedDevice sd(h->device);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return RDR_OK;
}

RDR_API int rdr_host_alloc(int64_t bytes, void **out) {
    CHECK_ARG(nullptr, out != nullptr && bytes >= 0, "rdr_host_alloc: bad arguments");
    *out = nullptr;
    if (rdr_device_count() == 0) return fail(nullptr, RDR_ERR_CUDA, "rdr_host_alloc: no CUDA device available; libraider_b200 has no CPU fallback");
    const size_t want = std::max<size_t>((size_t)bytes, 64);
    if ((*out = g_pinned.take(want))) return RDR_OK;
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, RDR_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    }
    g_pinned.add_live(p, want);
    *out = p;
    return RDR_OK;
}

RDR_API int rdr_host_free(void *p) {
    if (!p) return RDR_OK;
    if (!g_pinned.give(p)) return fail(nullptr, RDR_ERR_INVALID, "rdr_host_free: pointer was not returned by rdr_host_alloc");
    return RDR_OK;
}

RDR_API int64_t rdr_launch_count(rdr_handle_t h) { return h ? h->launches : 0; }

RDR_API int64_t rdr_last_fix_count(rdr_handle_t h) { return h ? h->last_fix_count : -1; }

// ------------------------------------------------------------------------------------------------
static int check_axis(rdr_handle_t h, const double *a, int64_t n, const char *name, std::vector<double> &out, bool &flipped) {
    CHECK_ARG(h, a != nullptr && n >= 2, std::string("rdr_set_cube: axis ") + name + " needs at least 2 nodes");
    out.assign(a, a + n);
    bool asc = true, desc = true;
    for (int64_t i = 1; i < n; ++i) {
        if (!(a[i] > a[i - 1])) asc = false;
        if (!(a[i] < a[i - 1])) desc = false;
    }
    CHECK_ARG(h, asc || desc, std::string("rdr_set_cube: axis ") + name + " must be strictly ascending or descending");
    flipped = !asc;
    if (flipped) std::reverse(out.begin(), out.end());
    return RDR_OK;
}

static int upload_fields(rdr_handle_t h, const float *wet, const float *hydro, int layout, int mem, float2 *dst) {
    const int64_t total = h->ny * h->nx * h->nz;
    const float *dw = wet, *dh = hydro;
    if (mem == RDR_MEM_HOST) {
        CUDA_TRY(h, h->d_stage.reserve(2 * total * sizeof(float)));
        float *s = h->d_stage.as<float>();
        CUDA_TRY(h, cudaMemcpyAsync(s, wet, total * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(s + total, hydro, total * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        dw = s;
        dh = s + total;
    }
    k_gather_fields<<<grid_for(total, 256, h->sm_count, 16), 256, 0, h->stream>>>(dw, dh, dst, (int)h->ny, (int)h->nx, (int)h->nz, layout,
                                                                                  h->flip_y, h->flip_x, h->flip_z);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return RDR_OK;
}

static int pack_cells(rdr_handle_t h) {
    const int64_t ncol = h->ny * h->nx;
    CUDA_TRY(h, h->d_cells.reserve(ncol * (h->nz - 1) * sizeof(double4)));
    CUDA_TRY(h, h->d_cells32.reserve(ncol * (h->nz - 1) * sizeof(float4)));
    k_pack_cells<<<grid_for(ncol * (h->nz - 1), 256, h->sm_count, 16), 256, 0, h->stream>>>(h->d_fields.as<float2>(), h->d_cells.as<double4>(),
                                                                                            h->d_cells32.as<float4>(), ncol, (int)h->nz);
    h->launches++;
    const int64_t nlerp = (h->ny - 1) * (h->nx - 1) * (h->nz - 1) * 4;
    // (+ LERP_PAD records: the thin-layer kernel prefetches up to that many records past the one it reads)
    CUDA_TRY(h, h->d_lerp.reserve((nlerp + 4 * LERP_PAD) * sizeof(double4)));
    k_pack_lerp<<<grid_for(nlerp, 256, h->sm_count, 16), 256, 0, h->stream>>>(h->d_fields.as<float2>(), h->d_lerp.as<double4>(), (int)h->ny,
                                                                              (int)h->nx, (int)h->nz);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return RDR_OK;
}

RDR_API int rdr_set_cube(rdr_handle_t h, const double *ys, int64_t ny, const double *xs, int64_t nx, const double *zs, int64_t nz,
                         const float *wet, const float *hydro, int layout, int crs_kind, const double *crs_params, int mem) {
    CHECK_ARG(h, h != nullptr, "rdr_set_cube: NULL handle");
    CHECK_ARG(h, wet && hydro, "rdr_set_cube: NULL field pointer");
    CHECK_ARG(h, layout == RDR_LAYOUT_ZYX || layout == RDR_LAYOUT_YXZ, "rdr_set_cube: unknown layout");
    CHECK_ARG(h, crs_kind == RDR_CRS_GEOGRAPHIC || crs_kind == RDR_CRS_LCC_SPHERE, "rdr_set_cube: unknown crs_kind");
    CHECK_ARG(h, crs_kind == RDR_CRS_GEOGRAPHIC || crs_params, "rdr_set_cube: LCC needs crs_params");
    CHECK_ARG(h, ny < 65536 && nx < 65536 && nz <= MAX_LAYERS && ny * nx * nz < (1ll << 31),
              "rdr_set_cube: cube too large (axes are limited to 65535 nodes, z to 1024, 2^31 cells in total)");
    ScopedDevice sd(h->device);
    h->has_cube = false;
    h->has_rays = false;
    int rc;
    if ((rc = check_axis(h, ys, ny, "y", h->ys, h->flip_y))) return rc;
    if ((rc = check_axis(h, xs, nx, "x", h->xs, h->flip_x))) return rc;
    if ((rc = check_axis(h, zs, nz, "z", h->zs, h->flip_z))) return rc;
    h->ny = ny; h->nx = nx; h->nz = nz;
    h->crs_kind = crs_kind;
    for (int i = 0; i < 7; ++i) h->crs[i] = crs_params ? crs_params[i] : 0.0;
    std::vector<double> axes;
    axes.insert(axes.end(), h->ys.begin(), h->ys.end());
    axes.insert(axes.end(), h->xs.begin(), h->xs.end());
    axes.insert(axes.end(), h->zs.begin(), h->zs.end());
    CUDA_TRY(h, h->d_axes.reserve(axes.size() * sizeof(double)));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_axes.p, axes.data(), axes.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if ((rc = build_axis_tables(h))) return rc;
    CUDA_TRY(h, h->d_fields.reserve(ny * nx * nz * sizeof(float2)));
    if ((rc = upload_fields(h, wet, hydro, layout, mem, h->d_fields.as<float2>()))) return rc;
    if ((rc = pack_cells(h))) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // host staging vectors go out of scope
    h->has_cube = true;
    return RDR_OK;
}

RDR_API int rdr_blend_cube(rdr_handle_t h, const float *wet1, const float *hydro1, int layout, double w0, double w1, int mem) {
    CHECK_ARG(h, h != nullptr, "rdr_blend_cube: NULL handle");
    if (!h->has_cube) return fail(h, RDR_ERR_STATE, "rdr_blend_cube: no cube staged (call rdr_set_cube first)");
    CHECK_ARG(h, wet1 && hydro1, "rdr_blend_cube: NULL field pointer");
    ScopedDevice sd(h->device);
    const int64_t total = h->ny * h->nx * h->nz;
    // second epoch goes behind the (possibly host-staged) raw fields in d_stage
    DevBuf second;
    CUDA_TRY(h, second.reserve(total * sizeof(float2)));
    int rc = upload_fields(h, wet1, hydro1, layout, mem, second.as<float2>());
    if (rc == RDR_OK) {
        k_blend_fields<<<grid_for(total, 256, h->sm_count, 16), 256, 0, h->stream>>>(h->d_fields.as<float2>(), second.as<float2>(), total,
                                                                                     (float)w0, (float)w1);
        h->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = fail(h, RDR_ERR_CUDA, "k_blend_fields launch failed");
    }
    if (rc == RDR_OK) rc = pack_cells(h);
    cudaStreamSynchronize(h->stream);
    second.release();
    h->has_rays = false;
    return rc;
}

// ------------------------------------------------------------------------------------------------
RDR_API int rdr_sample(rdr_handle_t h, const void *pts, int64_t n, void *out_wet, void *out_hydro, int dtype, int semantics, int mem) {
    CHECK_ARG(h, h != nullptr, "rdr_sample: NULL handle");
    if (!h->has_cube) return fail(h, RDR_ERR_STATE, "rdr_sample: no cube staged");
    CHECK_ARG(h, n >= 0 && (n == 0 || (pts && out_wet && out_hydro)), "rdr_sample: NULL pointer");
    CHECK_ARG(h, dtype == RDR_F64 || dtype == RDR_F32, "rdr_sample: dtype must be RDR_F64 or RDR_F32");
    CHECK_ARG(h, semantics >= RDR_SEM_SCIPY && semantics <= RDR_SEM_RAIDER_CLAMP, "rdr_sample: unknown semantics");
    if (n == 0) return RDR_OK;
    ScopedDevice sd(h->device);
    const size_t es = dtype == RDR_F64 ? 8 : 4;
    const void *dpts = pts;
    void *dw = out_wet, *dh = out_hydro;
    if (mem == RDR_MEM_HOST) {
        CUDA_TRY(h, h->d_in.reserve(n * 3 * es));
        CUDA_TRY(h, h->d_out.reserve(n * 2 * es));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_in.p, pts, n * 3 * es, cudaMemcpyHostToDevice, h->stream));
        dpts = h->d_in.p;
        dw = h->d_out.p;
        dh = static_cast<char *>(h->d_out.p) + n * es;
    }
    const CubeView c = make_view(h);
    if (semantics == RDR_SEM_SCIPY && (reinterpret_cast<uintptr_t>(dpts) & 15) == 0) {
        const bool uni = c.ay.uniform && c.ax.uniform;
        const char *noexact = getenv("RDR_K2_NO_EXACT_UNIFORM");
        const bool exact_uni = c.ay.exact_uniform && c.ax.exact_uniform && !(noexact && atoi(noexact) != 0);
        const char *ppt_env = getenv("RDR_K2_PPT");
        const int ppt = ppt_env ? atoi(ppt_env) : 2;
#define RDR_LAUNCH_K2(T, M, P)                                                                                                   \
    do {                                                                                                                         \
        const size_t smem = K2_STAGES * (K2_THREADS * P) * 3 * es + K2_STAGES * sizeof(uint64_t);                                \
        const int64_t ntiles = std::max<int64_t>(1, n / (K2_THREADS * P));                                                       \
        CUDA_TRY(h, cudaFuncSetAttribute(k_sample_stream<T, M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        int occ = 1;                                                                                                             \
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sample_stream<T, M, P>, K2_THREADS, smem));            \
        const int g = (int)std::min<int64_t>(ntiles, (int64_t)h->sm_count * std::max(1, occ)); /* one resident wave, persistent */ \
        k_sample_stream<T, M, P><<<g, K2_THREADS, smem, h->stream>>>(c, static_cast<const T *>(dpts), n, static_cast<T *>(dw),   \
                                                                     static_cast<T *>(dh));                                      \
    } while (0)
#define RDR_LAUNCH_K2_P(T, M)                      \
    do {                                           \
        if (ppt == 4) RDR_LAUNCH_K2(T, M, 4);      \
        else if (ppt == 3) RDR_LAUNCH_K2(T, M, 3); \
        else if (ppt == 1) RDR_LAUNCH_K2(T, M, 1); \
        else RDR_LAUNCH_K2(T, M, 2);               \
    } while (0)
        if (dtype == RDR_F64) {
            if (exact_uni) RDR_LAUNCH_K2_P(double, GUESS_EXACT_UNIFORM); else if (uni) RDR_LAUNCH_K2_P(double, GUESS_UNIFORM); else RDR_LAUNCH_K2_P(double, GUESS_BINS);
        } else {
            // fp32 tier: fp32 arithmetic (k_sample_stream_f32) unless RDR_K2_F32_ARITH=0 asks for the fp64 arithmetic on fp32 I/O
            const char *f32_env = getenv("RDR_K2_F32_ARITH");
            if (!(f32_env && atoi(f32_env) == 0)) {
#define RDR_LAUNCH_K2F(P, E)                                                                                                        \
    do {                                                                                                                         \
        const size_t smem = K2_STAGES * (K2_THREADS * P) * 3 * es + K2_STAGES * sizeof(uint64_t);                                \
        const int64_t ntiles = std::max<int64_t>(1, n / (K2_THREADS * P));                                                       \
        CUDA_TRY(h, cudaFuncSetAttribute(k_sample_stream_f32<P, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        int occ = 1;                                                                                                             \
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sample_stream_f32<P, E>, K2_THREADS, smem));           \
        const int g = (int)std::min<int64_t>(ntiles, (int64_t)h->sm_count * std::max(1, occ));                                   \
        k_sample_stream_f32<P, E><<<g, K2_THREADS, smem, h->stream>>>(c, static_cast<const float *>(dpts), n,                    \
                                                                      static_cast<float *>(dw), static_cast<float *>(dh));       \
    } while (0)
                const char *ppt32_env = getenv("RDR_K2_PPT32");
                const int ppt32 = ppt32_env ? atoi(ppt32_env) : 4;
                const bool xy32 = c.fy.exact32 && c.fx.exact32 && !(noexact && atoi(noexact) != 0);
                if (xy32) {
                    if (ppt32 == 1) RDR_LAUNCH_K2F(1, true); else if (ppt32 == 2) RDR_LAUNCH_K2F(2, true); else RDR_LAUNCH_K2F(4, true);
                } else {
                    if (ppt32 == 1) RDR_LAUNCH_K2F(1, false); else if (ppt32 == 2) RDR_LAUNCH_K2F(2, false); else RDR_LAUNCH_K2F(4, false);
                }
#undef RDR_LAUNCH_K2F
            } else if (exact_uni) RDR_LAUNCH_K2_P(float, GUESS_EXACT_UNIFORM); else if (uni) RDR_LAUNCH_K2_P(float, GUESS_UNIFORM); else RDR_LAUNCH_K2_P(float, GUESS_BINS);
        }
#undef RDR_LAUNCH_K2_P
#undef RDR_LAUNCH_K2
    } else {
        constexpr int BLOCK = 256;
        const int grid = grid_for(n, BLOCK, h->sm_count, 8);
        if (dtype == RDR_F64)
            k_sample_points<double, BLOCK><<<grid, BLOCK, 0, h->stream>>>(c, static_cast<const double *>(dpts), n, static_cast<double *>(dw),
                                                                          static_cast<double *>(dh), semantics);
        else
            k_sample_points<float, BLOCK><<<grid, BLOCK, 0, h->stream>>>(c, static_cast<const float *>(dpts), n, static_cast<float *>(dw),
                                                                         static_cast<float *>(dh), semantics);
    }
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    if (mem == RDR_MEM_HOST) {
        CUDA_TRY(h, cudaMemcpyAsync(out_wet, dw, n * es, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(out_hydro, dh, n * es, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    return RDR_OK;
}

RDR_API int rdr_sample_grid_levels(rdr_handle_t h, const double *xpts, int64_t nx, const double *ypts, int64_t ny, const double *zpts, int64_t nh,
                                   double *out_wet, double *out_hydro, int mem) {
    CHECK_ARG(h, h != nullptr, "rdr_sample_grid_levels: NULL handle");
    if (!h->has_cube) return fail(h, RDR_ERR_STATE, "rdr_sample_grid_levels: no cube staged");
    CHECK_ARG(h, xpts && ypts && zpts && out_wet && out_hydro && nx > 0 && ny > 0 && nh > 0, "rdr_sample_grid_levels: bad arguments");
    ScopedDevice sd(h->device);
    const int64_t n = nx * ny * nh;
    const double *dx, *dy, *dz;
    int rc;
    // query axes are small parameter vectors: always host
    if ((rc = stage_in(h, h->d_gx, xpts, nx, RDR_MEM_HOST, &dx))) return rc;
    if ((rc = stage_in(h, h->d_gy, ypts, ny, RDR_MEM_HOST, &dy))) return rc;
    if ((rc = stage_in(h, h->d_plan, zpts, nh, RDR_MEM_HOST, &dz))) return rc;
    double *dw = out_wet, *dh = out_hydro;
    if (mem == RDR_MEM_HOST) {
        CUDA_TRY(h, h->d_out.reserve(n * 2 * sizeof(double)));
        dw = h->d_out.as<double>();
        dh = dw + n;
    }
    k_sample_grid<<<grid_for(n, 256, h->sm_count, 8), 256, 0, h->stream>>>(make_view(h), dx, (int)nx, dy, (int)ny, dz, (int)nh, dw, dh);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    if (mem == RDR_MEM_HOST) {
        CUDA_TRY(h, cudaMemcpyAsync(out_wet, dw, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(out_hydro, dh, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->has_rays = false;  // d_gx / d_gy / d_plan were reused
    return RDR_OK;
}

RDR_API int rdr_sample_grid(rdr_handle_t h, const double *xpts, int64_t nx, const double *ypts, int64_t ny, double ht, double *out_wet,
                            double *out_hydro, int mem) {
    return rdr_sample_grid_levels(h, xpts, nx, ypts, ny, &ht, 1, out_wet, out_hydro, mem);
}

// ------------------------------------------------------------------------------------------------
RDR_API int rdr_ray_plan(rdr_handle_t h, double ht, double zref, int64_t *n_layers, double *low_ht, double *high_ht) {
    CHECK_ARG(h, h != nullptr, "rdr_ray_plan: NULL handle");
    if (!h->has_cube) return fail(h, RDR_ERR_STATE, "rdr_ray_plan: no cube staged");
    CHECK_ARG(h, n_layers != nullptr, "rdr_ray_plan: n_layers is NULL");
    std::vector<double> lo, hi;
    std::vector<int> cell;
    layer_plan(h->zs, ht, zref, lo, hi, cell);
    *n_layers = (int64_t)lo.size();
    if (low_ht) std::copy(lo.begin(), lo.end(), low_ht);
    if (high_ht) std::copy(hi.begin(), hi.end(), high_ht);
    return RDR_OK;
}

static RayGeom make_geom(rdr_handle_t h) {
    RayGeom G;
    G.geom_kind = h->geom_kind;
    G.los_kind = h->los_kind;
    G.gx = h->p_gx;
    G.gy = h->p_gy;
    G.los = h->p_los;
    G.e = h->los_const[0];
    G.n = h->los_const[1];
    G.u = h->los_const[2];
    G.ht = h->ht;
    G.nx = (int)h->ray_nx;
    return G;
}

// ---- K0: everything up to (and including) the launch; nothing is read back ---------------------------------------------------
static int k0_enqueue(rdr_handle_t h, int geom_kind, const double *gx, const double *gy, int64_t ny, int64_t nx, int los_kind,
                      const double *los, double ht, double zref, int exact_k0, int mem, const char *who) {
    CHECK_ARG(h, h != nullptr, std::string(who) + ": NULL handle");
    if (!h->has_cube) return fail(h, RDR_ERR_STATE, std::string(who) + ": no cube staged");
    CHECK_ARG(h, geom_kind == RDR_GEOM_GRID || geom_kind == RDR_GEOM_POINTS, std::string(who) + ": unknown geom_kind");
    CHECK_ARG(h, los_kind >= RDR_LOS_ARRAY && los_kind <= RDR_LOS_ORBIT, std::string(who) + ": unknown los_kind");
    CHECK_ARG(h, gx && gy && ny > 0 && nx > 0, std::string(who) + ": bad geometry arguments");
    CHECK_ARG(h, los_kind == RDR_LOS_ZENITH || los != nullptr, std::string(who) + ": los is NULL");
    h->has_rays = false;
    const int64_t n = ny * nx;
    h->n_rays = n; h->ray_ny = ny; h->ray_nx = nx;
    h->geom_kind = geom_kind; h->los_kind = los_kind;
    h->ht = ht; h->zref = zref;
    layer_plan(h->zs, ht, zref, h->low_ht, h->high_ht, h->layer_cell);
    const int K = (int)h->low_ht.size();
    h->n_layers = K;
    if (K == 0) return fail(h, RDR_ERR_NO_LAYERS, std::string(who) + ": no model layer contributes between ht and zref");
    CHECK_ARG(h, K <= MAX_LAYERS, std::string(who) + ": more than 1024 contributing layers");
    int rc;
    // geometry: grid axes are parameter vectors (host); point lists and LOS arrays are bulk (per `mem`)
    if (geom_kind == RDR_GEOM_GRID) {
        if ((rc = stage_in(h, h->d_gx, gx, nx, RDR_MEM_HOST, &h->p_gx))) return rc;
        if ((rc = stage_in(h, h->d_gy, gy, ny, RDR_MEM_HOST, &h->p_gy))) return rc;
    } else {
        if ((rc = stage_in(h, h->d_gx, gx, n, mem, &h->p_gx))) return rc;
        if ((rc = stage_in(h, h->d_gy, gy, n, mem, &h->p_gy))) return rc;
    }
    h->p_los = nullptr;
    if (los_kind == RDR_LOS_ARRAY) {
        if ((rc = stage_in(h, h->d_los, los, n * 3, mem, &h->p_los))) return rc;
    } else if (los_kind == RDR_LOS_ENU_CONST) {
        h->los_const[0] = los[0]; h->los_const[1] = los[1]; h->los_const[2] = los[2];
    } else if (los_kind == RDR_LOS_ORBIT) {
        // los = {n_sv, then n_sv rows of (t, x, y, z, vx, vy, vz)} on the host: K6 turns it into per-ray ECEF vectors on the device
        const int64_t n_sv = (int64_t)los[0];
        std::vector<double> blob;
        if ((rc = split_orbit(h, los + 1, n_sv, blob))) return rc;
        CUDA_TRY(h, h->d_orbit.reserve(blob.size() * sizeof(double)));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_orbit.p, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, h->d_los.reserve((size_t)n * 3 * sizeof(double)));
        const double *ob = h->d_orbit.as<double>();
        const OrbitView O = {ob, ob + n_sv, ob + 4 * n_sv, (int)n_sv, (double)(n_sv - 1) / (blob[n_sv - 1] - blob[0])};
        k_orbit_los<<<grid_for(n, 128, h->sm_count, 16), 128, 0, h->stream>>>(O, geom_kind, h->p_gx, h->p_gy, nullptr, ht, (int)nx, n, 1.0e-7, 30,
                                                                             h->d_los.as<double>(), nullptr, nullptr);
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // blob goes out of scope
        h->p_los = h->d_los.as<double>();
        h->los_kind = RDR_LOS_ARRAY;
    }
    // low[K] | high[K] | layer cell int[K] (read by k_plan)
    std::vector<double> plan(h->low_ht);
    plan.insert(plan.end(), h->high_ht.begin(), h->high_ht.end());
    plan.resize(2 * (size_t)K + ((size_t)K + 1) / 2);
    memcpy(plan.data() + 2 * (size_t)K, h->layer_cell.data(), (size_t)K * sizeof(int));
    CUDA_TRY(h, h->d_plan.reserve(plan.size() * sizeof(double)));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_plan.p, plan.data(), plan.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, h->d_t.reserve((size_t)(K + 1) * n * sizeof(double)));
    CUDA_TRY(h, h->d_red.reserve((XCHG_STRIDE + 8) * sizeof(unsigned long long)));
    CUDA_TRY(h, cudaMemsetAsync(h->d_red.p, 0, (K + 8) * sizeof(unsigned long long), h->stream));
    constexpr int BLOCK = 128;
    const int minb = tune_minb("RDR_K0_MINB", 6);
    const int grid = grid_for(n, BLOCK, h->sm_count, 4 * minb);
    const size_t smem = (BLOCK / 32) * (size_t)(K + 3) * sizeof(unsigned long long) + 3 * (size_t)K * sizeof(double);
    // RDR_K0_MODE: poly (default: septic h(t) + the layer tops as a polynomial in z) | iter (septic h(t), three iterates per layer) |
    // exact (the reference's iterates on Bowring heights)
    const char *k0_env = getenv("RDR_K0_MODE");
    const int use_cubic = (exact_k0 || (k0_env && !strcmp(k0_env, "exact"))) ? 0 : (k0_env && !strcmp(k0_env, "iter")) ? 1 : 2;
    h->k0_was_cubic = use_cubic != 0;
#define RDR_LAUNCH_K0(M)                                                                                                              \
    k_ray_layers<BLOCK, M><<<grid, BLOCK, smem, h->stream>>>(make_geom(h), n, K, h->d_plan.as<double>(), h->d_t.as<double>(),          \
                                                            h->d_red.as<unsigned long long>(), h->zs.front(), h->zs.back(), use_cubic)
    switch (minb) {
        case 4: RDR_LAUNCH_K0(4); break;
        case 5: RDR_LAUNCH_K0(5); break;
        case 6: RDR_LAUNCH_K0(6); break;
        default: RDR_LAUNCH_K0(8); break;
    }
#undef RDR_LAUNCH_K0
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    h->has_rays = true;
    return RDR_OK;
}

// Default span length of the polynomial integrators [m of the longest ray].  24 km keeps the delays within 3e-10 m of the PROJ-form
// arithmetic and costs the fewest span set-ups (profiles/runs/r02p_span.sh); on km-scale grids (HRRR: 3 km cells) the footprint
// of a 24 km span no longer fits the staged-record capacity of the thin-layer kernel (10 columns x 42 z cells), and 12 km is
// faster (K3 2.06 vs 2.31 ms on the 57-node Lambert variant of bench.py).  RDR_K3_SPAN overrides.
static double default_span_max(rdr_handle_t h) {
    const char *span_env = getenv("RDR_K3_SPAN");
    if (span_env && atof(span_env) > 0) return atof(span_env);
    const double dy = (h->ys.back() - h->ys.front()) / (double)std::max<size_t>(h->ys.size() - 1, 1);
    const double dx = (h->xs.back() - h->xs.front()) / (double)std::max<size_t>(h->xs.size() - 1, 1);
    double cell_m;
    if (h->crs_kind == RDR_CRS_LCC_SPHERE) {
        cell_m = std::min(fabs(dy), fabs(dx));
    } else {
        const double mid = 0.5 * (h->ys.back() + h->ys.front()) * (M_PI / 180.0);
        cell_m = std::min(fabs(dy), fabs(dx) * std::max(cos(mid), 0.05)) * 111.0e3;
    }
    return cell_m < 8000.0 ? 12000.0 : 24000.0;
}

// ---- k_plan: MAX / SUM over the slots, nParts, layer records, spans, predicates -> the device plan ------------------------------
static int plan_enqueue(rdr_handle_t h, const unsigned long long *slots, int world, double max_segment_length, int force_clamp, int block_mask) {
    const int K = h->n_layers;
    CUDA_TRY(h, h->d_devplan.reserve(sizeof(DevPlan)));
    const double span_max = default_span_max(h);
    const char *thin_env = getenv("RDR_K3_THIN_MIN");  // fewest thin layers (<= 3 samples) that are worth the thin-layer kernel; 0: never
    const char *uni_env = getenv("RDR_K3_UNIFIED");   // 1: every layer goes to the staged kernel (K >= 4 rows of distances are preloaded)
    int thin_min = thin_env ? (atoi(thin_env) > 0 ? std::max(atoi(thin_env), 4) : 0) : 16;
    const char *absorb_env = getenv("RDR_K3_THIN_ABSORB");  // most samples of a thick tail that the thin-layer kernel takes over
    const int thin_absorb = absorb_env ? std::max(atoi(absorb_env), 0) : 128;
    if (uni_env && atoi(uni_env) != 0 && K >= 8) thin_min = -1;
    unsigned long long *counters = h->d_red.as<unsigned long long>() + XCHG_STRIDE;
    const int *d_cell = reinterpret_cast<const int *>(h->d_plan.as<double>() + 2 * (size_t)K);
    const double *znodes = h->d_axes.as<double>() + h->ny + h->nx;
    k_plan<<<1, 256, 0, h->stream>>>(slots, world, XCHG_STRIDE, K, d_cell, znodes, (int)h->nz, max_segment_length, span_max,
                                     h->thin_ok ? thin_min : 0, thin_absorb, force_clamp, block_mask, h->d_devplan.as<DevPlan>(), counters);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    h->last_max_seg = max_segment_length;
    return RDR_OK;
}

template <typename KERNEL>
static cudaError_t allow_smem(KERNEL k, size_t smem) {
    // > 48 KB of dynamic shared memory (models with more than ~600 levels) needs the opt-in
    return smem > 48 * 1024 ? cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) : cudaSuccess;
}

// ---- K3: the integration kernels on the device plan; `mode`: 0 auto (poly + thin), 1 fast, 2 general -----------------------------
static int k3_enqueue(rdr_handle_t h, void *out_wet, void *out_hydro, int out_dtype, int accumulate, int mem, int mode, bool *staged) {
    const int K = h->n_layers;
    const int64_t n = h->n_rays;
    const size_t es = out_dtype == RDR_F64 ? 8 : 4;
    unsigned long long *counters = h->d_red.as<unsigned long long>() + XCHG_STRIDE;
    void *dw = out_wet, *dh = out_hydro;
    *staged = false;
    if (mem == RDR_MEM_HOST) {
        // page-locked result arrays (rdr_host_alloc) are written by the kernel itself: 16 B per ray of posted PCIe writes spread
        // over the whole integration instead of a 2 x n x 8 B copy after it
        void *aw = accumulate ? nullptr : device_alias_of_host(out_wet), *ah = accumulate ? nullptr : device_alias_of_host(out_hydro);
        if (aw && ah) {
            dw = aw;
            dh = ah;
        } else {
            *staged = true;
            CUDA_TRY(h, h->d_out.reserve(2 * n * es));
            dw = h->d_out.p;
            dh = static_cast<char *>(h->d_out.p) + n * es;
            if (accumulate) {
                CUDA_TRY(h, cudaMemcpyAsync(dw, out_wet, n * es, cudaMemcpyHostToDevice, h->stream));
                CUDA_TRY(h, cudaMemcpyAsync(dh, out_hydro, n * es, cudaMemcpyHostToDevice, h->stream));
            }
        }
    }
    constexpr int BLOCK = 128;
    const CubeView c = make_view(h);
    const RayGeom G = make_geom(h);
    const DevPlan *P = h->d_devplan.as<DevPlan>();
    const double *t_in = h->d_t.as<double>();
    PeerOut peers = h->peers;
    if (accumulate) peers.n = peers.multicast = 0;  // += has no meaning across replicas: peers only mirror freshly written maps
    FastCube fc;
    const char *force_general = getenv("RDR_K3_GENERAL");
    // integrator: poly (default; geographic or Lambert cube with uniform horizontal axes), fast (per-sample Bowring; geographic
    // only), general (PROJ-form arithmetic for every sample).  RDR_K3_MODE = poly | fast | general overrides for tests / tuning.
    const char *mode_env = getenv("RDR_K3_MODE");
    const bool want_general = mode == 2 || (force_general && atoi(force_general) != 0) || (mode_env && !strcmp(mode_env, "general"));
    const bool fast_cube = make_fast_cube(h, fc) && !want_general && n < (1ll << 31);
    const bool poly = fast_cube && mode != 1 && !(mode_env && !strcmp(mode_env, "fast"));
    const bool fast = fast_cube && (poly || fc.crs_kind == RDR_CRS_GEOGRAPHIC);
    h->last_k3_poly = poly;
    h->last_fix_count = -1;
#define RDR_LAUNCH_K3(T, M, LIST, COUNT)                                                                                                   \
    k_ray_integrate<T, BLOCK, M><<<grid, BLOCK, 0, h->stream>>>(c, G, n, K, t_in, P, h->zs.front(), h->zs.back(), static_cast<T *>(dw),   \
                                                                static_cast<T *>(dh), accumulate, peers, counters, LIST, COUNT)
    if (fast) {
        CUDA_TRY(h, h->d_fix.reserve(std::max<size_t>(2 * n * sizeof(int), 16)));
        const size_t smem = K * sizeof(LayerRec) + (2 * (size_t)h->nz - 1) * sizeof(double);
        const double *znodes = h->d_axes.as<double>() + h->ny + h->nx;
        if (poly) {
            const size_t smem_p = smem + (size_t)K * sizeof(int);
            const bool lcc = fc.crs_kind == RDR_CRS_LCC_SPHERE;
            // cell-record cache (CACHE = true) + layer quadrature: the thick layers.  Needs the packed cell key.
            const char *cache_env = getenv("RDR_K3_CACHE");
            const bool key_ok = h->ny <= 1024 && h->nx <= 1024 && h->nz <= 2048;
            const bool split = key_ok && (cache_env ? atoi(cache_env) != 0 : true);
            const char *quad_env = getenv("RDR_K3_QUAD");  // layer quadrature (closed-form trapezoid sum per one-cell layer): on unless 0
            const int quad = !(quad_env && atoi(quad_env) == 0);
            const bool tiles_ok = h->geom_kind == RDR_GEOM_GRID;
            auto tile_for = [&](const char *env, int dflt) {
                // log2 of the pixel-tile width a warp takes: 3 -> 8 x 4 pixels; 0 = 32 pixels of a row
                const char *v = getenv(env);
                int t = v ? atoi(v) : dflt;
                if (t < 0 || t > 4 || !tiles_ok || h->ray_nx % (1 << t) || h->ray_ny % (32 >> t)) t = 0;
                return t;
            };
            // tiles pay where divergence is the cost (the cached / quadrature path)
            const int tile_map = tile_for("RDR_K3_TILE", split ? 3 : 0);
            const int tile_thin = tile_for("RDR_K3_THIN_TILE", 3);
            const int minb_p = tune_minb("RDR_K3_MINB", 4);
            const int grid_p = grid_for(n, BLOCK, h->sm_count, 4 * minb_p);
            CUDA_TRY(h, h->d_part.reserve(std::max<size_t>(2 * n * sizeof(double), 16)));
            double *part = h->d_part.as<double>();
#define RDR_LAUNCH_K3P1(T, M, L, S, F0)                                                                                                          \
    do {                                                                                                                                         \
        CUDA_TRY(h, allow_smem(k_ray_integrate_poly<T, BLOCK, M, L, S, F0>, smem_p));                                                            \
        k_ray_integrate_poly<T, BLOCK, M, L, S, F0><<<grid_p, BLOCK, smem_p, h->stream>>>(fc, G, n, t_in, P, znodes, (int)h->nz, h->zs.front(),  \
                                                                                       static_cast<T *>(dw), static_cast<T *>(dh), accumulate,  \
                                                                                       peers, counters, h->d_fix.as<int>(), quad, tile_map,     \
                                                                                       part);                                                   \
        h->launches++;                                                                                                                           \
    } while (0)
#define RDR_LAUNCH_K3P(T, M, L, S)                       \
    do {                                                 \
        RDR_LAUNCH_K3P1(T, M, L, S, true);               \
        if (h->thin_ok) RDR_LAUNCH_K3P1(T, M, L, S, false); \
    } while (0)
#define RDR_LAUNCH_K3P_M(T, L)                                                     \
    switch (minb_p) {                                                              \
        case 3: if (split) RDR_LAUNCH_K3P(T, 3, L, true); else RDR_LAUNCH_K3P(T, 3, L, false); break;   \
        default: if (split) RDR_LAUNCH_K3P(T, 4, L, true); else RDR_LAUNCH_K3P(T, 4, L, false); break;  \
    }
            if (out_dtype == RDR_F64) {
                if (lcc) { RDR_LAUNCH_K3P_M(double, true) } else { RDR_LAUNCH_K3P_M(double, false) }
            } else {
                if (lcc) RDR_LAUNCH_K3P(float, 4, true, false); else RDR_LAUNCH_K3P(float, 4, false, false);
            }
#undef RDR_LAUNCH_K3P_M
#undef RDR_LAUNCH_K3P
#undef RDR_LAUNCH_K3P1
            CUDA_TRY(h, cudaGetLastError());
            if (h->thin_ok) {
                // the thin-layer part of the plan (no-op when the plan has none); runs second and stores the results
                const char *pfc_env = getenv("RDR_K3_THIN_PF"), *st_env = getenv("RDR_K3_THIN_STAGE");
                const int pf_cells = pfc_env ? std::min(std::max(atoi(pfc_env), 0), LERP_PAD) : 0;
                const int minb_t = tune_minb("RDR_K3_THIN_MINB", 4);
                const int quad_t = !(quad_env && atoi(quad_env) == 0);
                const int grid_t = grid_for(n, BLOCK, h->sm_count, 4 * minb_t);
                // staged record columns (north_star: cube staged into shared memory via TMA): the records that fit beside minb_t CTAs per
                // SM (227 KB per SM, 1 KB reserved per CTA), at most 48 KB worth
                const size_t base_t = (smem_p + 127) / 128 * 128 + 128 + THIN_TD * BLOCK * sizeof(double);
                const size_t budget = (size_t)227 * 1024 / minb_t - 1024;
                int rec_cap = budget > base_t ? (int)std::min<size_t>((budget - base_t) / sizeof(LerpCell), 384) : 0;
                if (st_env) rec_cap = std::min(rec_cap, std::max(atoi(st_env), 0));
                const bool stage = rec_cap > 0;
                const size_t smem_t = base_t + (size_t)rec_cap * sizeof(LerpCell);
                unsigned long long *stage_stats = h->d_red.as<unsigned long long>() + XCHG_STRIDE + 4;
#define RDR_LAUNCH_K3T(T, M, L, ST, Q)                                                                                                     \
    do {                                                                                                                                    \
        CUDA_TRY(h, allow_smem(k_ray_integrate_thin<T, BLOCK, M, L, ST, Q>, smem_t));                                                       \
        if (ST) CUDA_TRY(h, cudaFuncSetAttribute(k_ray_integrate_thin<T, BLOCK, M, L, ST, Q>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)); \
        k_ray_integrate_thin<T, BLOCK, M, L, ST, Q><<<grid_t, BLOCK, smem_t, h->stream>>>(fc, G, n, t_in, P, znodes, (int)h->nz, h->zs.front(), \
                                                                                       static_cast<T *>(dw), static_cast<T *>(dh), accumulate, \
                                                                                       peers, counters, h->d_fix.as<int>(), tile_thin, part,  \
                                                                                       pf_cells, quad_t, rec_cap, stage_stats);               \
    } while (0)
    // QUAD (closed-form sums of thick layers inside the staged kernel) is only built for the unified mode (RDR_K3_UNIFIED=1), the A/B of
    // "cell records in shared memory" against k_ray_integrate_poly's register-held record: it costs the thin-layer loop registers
#define RDR_LAUNCH_K3T_S(T, M, L) \
    do { if (stage) RDR_LAUNCH_K3T(T, M, L, true, false); else RDR_LAUNCH_K3T(T, M, L, false, false); } while (0)
#define RDR_LAUNCH_K3T_M(T, L)                          \
    switch (minb_t) {                                   \
        case 3: RDR_LAUNCH_K3T_S(T, 3, L); break;       \
        case 5: RDR_LAUNCH_K3T_S(T, 5, L); break;       \
        default: if (unified && stage) RDR_LAUNCH_K3T(T, 4, L, true, true); else RDR_LAUNCH_K3T_S(T, 4, L); break; \
    }
                const bool unified = getenv("RDR_K3_UNIFIED") && atoi(getenv("RDR_K3_UNIFIED")) != 0;
                if (out_dtype == RDR_F64) {
                    if (lcc) { RDR_LAUNCH_K3T_M(double, true) } else { RDR_LAUNCH_K3T_M(double, false) }
                } else {
                    if (lcc) RDR_LAUNCH_K3T_S(float, 4, true); else RDR_LAUNCH_K3T_S(float, 4, false);
                }
#undef RDR_LAUNCH_K3T_M
#undef RDR_LAUNCH_K3T_S
#undef RDR_LAUNCH_K3T
                h->launches++;
                CUDA_TRY(h, cudaGetLastError());
            }
        } else {
            // (the per-sample Bowring form is kept for tests / comparisons: one occupancy variant, one or two samples per trip)
            const int grid = grid_for(n, BLOCK, h->sm_count, 4 * 5);
            const char *npt_env = getenv("RDR_K3_NPT");
            const int npt = npt_env ? atoi(npt_env) : 1;
#define RDR_LAUNCH_K3F(T, M, NP)                                                                                                           \
    do {                                                                                                                                    \
        CUDA_TRY(h, allow_smem(k_ray_integrate_fast<T, BLOCK, M, NP>, smem));                                                               \
        k_ray_integrate_fast<T, BLOCK, M, NP><<<grid, BLOCK, smem, h->stream>>>(fc, G, n, K, t_in, P, znodes, (int)h->nz, h->zs.front(),   \
                                                                               static_cast<T *>(dw), static_cast<T *>(dh), accumulate,     \
                                                                               peers, counters, h->d_fix.as<int>());                       \
    } while (0)
            if (out_dtype == RDR_F64) {
                if (npt == 1) RDR_LAUNCH_K3F(double, 5, 1); else RDR_LAUNCH_K3F(double, 5, 2);
            } else {
                RDR_LAUNCH_K3F(float, 5, 1);
            }
#undef RDR_LAUNCH_K3F
            h->launches++;
            CUDA_TRY(h, cudaGetLastError());
        }
        // flagged rays -> PROJ-form integrator in list mode; the count stays on the device (no-op launch when it is zero)
        {
            const int grid = grid_for(n, BLOCK, h->sm_count, 4 * 4);
            const int *list = h->d_fix.as<int>();
            const unsigned long long *count = counters + 3;
            if (out_dtype == RDR_F64) RDR_LAUNCH_K3(double, 4, list, count); else RDR_LAUNCH_K3(float, 4, list, count);
            h->launches++;
            CUDA_TRY(h, cudaGetLastError());
        }
    } else {
        const int minb = tune_minb("RDR_K3_MINB", 4);
        const int grid = grid_for(n, BLOCK, h->sm_count, 4 * minb);
        if (out_dtype == RDR_F64) {
            switch (minb) {
                case 2: RDR_LAUNCH_K3(double, 2, nullptr, nullptr); break;
                case 3: RDR_LAUNCH_K3(double, 3, nullptr, nullptr); break;
                case 5: RDR_LAUNCH_K3(double, 5, nullptr, nullptr); break;
                case 6: RDR_LAUNCH_K3(double, 6, nullptr, nullptr); break;
                case 8: RDR_LAUNCH_K3(double, 8, nullptr, nullptr); break;
                default: RDR_LAUNCH_K3(double, 4, nullptr, nullptr); break;
            }
        } else {
            RDR_LAUNCH_K3(float, 4, nullptr, nullptr);
        }
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
    }
#undef RDR_LAUNCH_K3
    if (*staged) {
        CUDA_TRY(h, cudaMemcpyAsync(out_wet, dw, n * es, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(out_hydro, dh, n * es, cudaMemcpyDeviceToHost, h->stream));
    }
    return RDR_OK;
}

RDR_API int rdr_ray_layers(rdr_handle_t h, int geom_kind, const double *gx, const double *gy, int64_t ny, int64_t nx, int los_kind,
                           const double *los, double ht, double zref, double *maxlen_out, int64_t *counts_out, int mem) {
    CHECK_ARG(h, h != nullptr, "rdr_ray_layers: NULL handle");
    ScopedDevice sd(h->device);
    if (counts_out) {
        counts_out[0] = ny * nx; counts_out[1] = 0; counts_out[2] = 0; counts_out[3] = 0; counts_out[4] = 0;
    }
    int rc = k0_enqueue(h, geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref, 0, mem, "rdr_ray_layers");
    if (rc) return rc;
    const int K = h->n_layers;
    std::vector<unsigned long long> red(K + 5);
    CUDA_TRY(h, cudaMemcpyAsync(red.data(), h->d_red.p, (K + 5) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (maxlen_out)
        for (int k = 0; k < K; ++k) memcpy(&maxlen_out[k], &red[k], sizeof(double));
    if (counts_out) {
        counts_out[1] = (int64_t)red[K];
        counts_out[2] = (int64_t)red[K + 1];
        counts_out[3] = K;
        counts_out[4] = (int64_t)red[K + 4];
    }
    // (every ray of THIS call being NaN is not an error here: the reference's np.isnan(ray_lengths).all() (delay.py:279) is over the
    // whole raster, a call may be one row tile or one rank's block -- the caller decides on the summed counts)
    return RDR_OK;
}

RDR_API int rdr_ray_integrate(rdr_handle_t h, const double *maxlen, double max_segment_length, int clamp, void *out_wet,
                              void *out_hydro, int out_dtype, int accumulate, int64_t *nparts_out, int64_t *oob_out, int mem) {
    CHECK_ARG(h, h != nullptr, "rdr_ray_integrate: NULL handle");
    if (!h->has_cube || !h->has_rays) return fail(h, RDR_ERR_STATE, "rdr_ray_integrate: call rdr_ray_layers first");
    CHECK_ARG(h, maxlen && out_wet && out_hydro, "rdr_ray_integrate: NULL pointer");
    CHECK_ARG(h, max_segment_length > 0, "rdr_ray_integrate: max_segment_length must be positive");
    CHECK_ARG(h, out_dtype == RDR_F64 || out_dtype == RDR_F32, "rdr_ray_integrate: out_dtype must be RDR_F64 or RDR_F32");
    ScopedDevice sd(h->device);
    const int K = h->n_layers;
    // nParts = ceil(max / MAX_SEGMENT_LENGTH).astype(int) + 1   (delay.py:283) -- the bit-exact integer contract.  The host
    // restates it for its caller; the kernels take it from the device plan (k_plan), which gets the caller's maxima as its slot
    std::vector<unsigned long long> slot(K + 3, 0ull);
    double acc = 0.0, longest = 0.0;
    const double span_max = default_span_max(h);
    for (int k = 0; k < K; ++k) {
        const double q = ceil(maxlen[k] / max_segment_length);
        CHECK_ARG(h, q == q && q < 1e7 && maxlen[k] >= 0, "rdr_ray_integrate: per-layer max length is NaN or absurd");
        int np = (int)q + 1;
        if (np < 2) np = 2;
        if (nparts_out) nparts_out[k] = np;
        memcpy(&slot[k], &maxlen[k], sizeof(double));
        if (k > 0 && acc + maxlen[k] > span_max) {
            longest = std::max(longest, acc);
            acc = 0.0;
        }
        acc += maxlen[k];
    }
    longest = std::max(longest, acc);
    slot[K + 2] = (unsigned long long)h->n_rays;
    unsigned long long *d_slot = h->d_red.as<unsigned long long>();
    CUDA_TRY(h, cudaMemcpyAsync(d_slot, slot.data(), slot.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream));
    int rc;
    // clamp: bit 0 = the first sample is below min(z) on every pixel of the raster, bit 1 = the last sample is above max(z) on every pixel
    if ((rc = plan_enqueue(h, d_slot, 1, max_segment_length, (clamp & 1) | ((clamp & 2) ? 2 : 4), RDR_PLAN_ABSURD))) return rc;
    bool staged = false;
    // a single layer longer than 2 spans would stretch the cubic's error bound (T^4) by > 16: leave those calls to `fast`
    if ((rc = k3_enqueue(h, out_wet, out_hydro, out_dtype, accumulate, mem, longest <= 2.0 * span_max ? 0 : 1, &staged))) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // `slot` goes out of scope; host outputs are complete on return
    if (oob_out || mem == RDR_MEM_HOST) {
        unsigned long long cnt[4] = {0, 0, 0, 0};
        CUDA_TRY(h, cudaMemcpy(cnt, h->d_red.as<unsigned long long>() + XCHG_STRIDE, sizeof(cnt), cudaMemcpyDeviceToHost));
        h->last_fix_count = (int64_t)cnt[3];
        if (oob_out) {
            oob_out[0] = (int64_t)cnt[0];  // first sample below min(z) (pre-clamp)
            oob_out[1] = (int64_t)cnt[1];  // samples below min(z) after the clamp decision
            oob_out[2] = (int64_t)cnt[2];  // samples above max(z)
        }
    }
    return RDR_OK;
}

// ---- the fused step: K0 -> [exchange] -> k_plan -> K3 with no host synchronisation in between ---------------------------------
RDR_API int rdr_set_exchange(rdr_handle_t h, int rank, int world, void *const *bufs) {
    CHECK_ARG(h, h != nullptr, "rdr_set_exchange: NULL handle");
    CHECK_ARG(h, world >= 0 && world <= RDR_MAX_PEERS, "rdr_set_exchange: at most 8 ranks");
    CHECK_ARG(h, world == 0 || (bufs && rank >= 0 && rank < world), "rdr_set_exchange: bad rank / buffer list");
    h->xchg_world = world;
    h->xchg_rank = rank;
    h->xchg_parity = 0;
    for (int i = 0; i < world; ++i) {
        CHECK_ARG(h, bufs[i] != nullptr, "rdr_set_exchange: NULL buffer");
        h->xchg_bufs[i] = bufs[i];
    }
    return RDR_OK;
}

RDR_API int64_t rdr_exchange_bytes(int world) { return (int64_t)2 * world * XCHG_STRIDE * (int64_t)sizeof(unsigned long long); }

static int publish(rdr_handle_t h, const unsigned long long *src, int nwords, int word_off) {
    PeerOut dst = {};
    dst.n = h->xchg_world;
    for (int i = 0; i < h->xchg_world; ++i) dst.wet[i] = h->xchg_bufs[i];
    const int off = (h->xchg_parity * h->xchg_world + h->xchg_rank) * XCHG_STRIDE + word_off;
    k_publish<<<1, 256, 0, h->stream>>>(src, nwords, off, dst);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return RDR_OK;
}

RDR_API int rdr_trace_begin(rdr_handle_t h, int geom_kind, const double *gx, const double *gy, int64_t ny, int64_t nx, int los_kind,
                            const double *los, double ht, double zref, int flags, int mem) {
    CHECK_ARG(h, h != nullptr, "rdr_trace_begin: NULL handle");
    ScopedDevice sd(h->device);
    int rc = k0_enqueue(h, geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref, (flags & RDR_TRACE_EXACT_K0) != 0, mem, "rdr_trace_begin");
    if (rc) return rc;
    h->trace_flags = flags;
    if (h->xchg_world > 0) {
        h->xchg_parity ^= 1;
        return publish(h, h->d_red.as<unsigned long long>(), h->n_layers + 5, 0);
    }
    return RDR_OK;
}

RDR_API int rdr_trace_finish(rdr_handle_t h, double max_segment_length, int force_clamp, int mode, void *out_wet, void *out_hydro,
                             int out_dtype, int accumulate, int mem) {
    CHECK_ARG(h, h != nullptr, "rdr_trace_finish: NULL handle");
    if (!h->has_cube || !h->has_rays) return fail(h, RDR_ERR_STATE, "rdr_trace_finish: call rdr_trace_begin first");
    CHECK_ARG(h, out_wet && out_hydro, "rdr_trace_finish: NULL pointer");
    CHECK_ARG(h, max_segment_length > 0, "rdr_trace_finish: max_segment_length must be positive");
    CHECK_ARG(h, out_dtype == RDR_F64 || out_dtype == RDR_F32, "rdr_trace_finish: out_dtype must be RDR_F64 or RDR_F32");
    CHECK_ARG(h, mode >= 0 && mode <= 2, "rdr_trace_finish: mode must be 0 (auto), 1 (fast) or 2 (general)");
    ScopedDevice sd(h->device);
    const unsigned long long *slots = h->d_red.as<unsigned long long>();
    int world = 1;
    if (h->xchg_world > 0) {
        slots = static_cast<const unsigned long long *>(h->xchg_bufs[h->xchg_rank]) + (size_t)h->xchg_parity * h->xchg_world * XCHG_STRIDE;
        world = h->xchg_world;
    }
    // what stops the integration kernels (the host reads the status back and redoes the step / raises):
    //   absurd maxima, every ray NaN, a single layer too long for the span cubics (mode 0 only), and -- when K0 ran in its
    //   default (span-cubic) form -- an nParts knife edge, which is redone with the exact K0 (detect, don't hide)
    int block = RDR_PLAN_ABSURD | RDR_PLAN_ALL_NAN;
    if (mode == 0) block |= RDR_PLAN_SPAN_TOO_LONG;
    if (h->k0_was_cubic && !(h->trace_flags & RDR_TRACE_NO_KNIFE_GUARD)) block |= RDR_PLAN_KNIFE_EDGE;
    int rc;
    if ((rc = plan_enqueue(h, slots, world, max_segment_length, force_clamp, block))) return rc;
    bool staged = false;
    if ((rc = k3_enqueue(h, out_wet, out_hydro, out_dtype, accumulate, mem, mode, &staged))) return rc;
    if (h->xchg_world > 0)  // K3's own count of first samples below min(z): the cross-check of the clamp predicate, summed in rdr_trace_result
        return publish(h, h->d_red.as<unsigned long long>() + XCHG_STRIDE, 1, h->n_layers + 3);
    return RDR_OK;
}

RDR_API int rdr_trace_result(rdr_handle_t h, double *maxlen_out, int64_t *nparts_out, int64_t *info_out) {
    CHECK_ARG(h, h != nullptr, "rdr_trace_result: NULL handle");
    if (!h->has_rays || !h->d_devplan.p) return fail(h, RDR_ERR_STATE, "rdr_trace_result: no step to report");
    ScopedDevice sd(h->device);
    const int K = h->n_layers;
    std::vector<unsigned char> buf(offsetof(DevPlan, nparts) + (size_t)K * sizeof(int));
    unsigned long long cnt[6];
    CUDA_TRY(h, cudaMemcpyAsync(buf.data(), h->d_devplan.p, buf.size(), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(cnt, h->d_red.as<unsigned long long>() + XCHG_STRIDE, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream));
    std::vector<unsigned long long> k3_below(std::max(h->xchg_world, 1), 0ull);
    if (h->xchg_world > 0) {
        const unsigned long long *base = static_cast<const unsigned long long *>(h->xchg_bufs[h->xchg_rank]) + (size_t)h->xchg_parity * h->xchg_world * XCHG_STRIDE;
        for (int q = 0; q < h->xchg_world; ++q)
            CUDA_TRY(h, cudaMemcpyAsync(&k3_below[q], base + (size_t)q * XCHG_STRIDE + K + 3, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const DevPlan *P = reinterpret_cast<const DevPlan *>(buf.data());
    if (maxlen_out) memcpy(maxlen_out, buf.data() + offsetof(DevPlan, maxlen), (size_t)K * sizeof(double));
    if (nparts_out) {
        const int *np = reinterpret_cast<const int *>(buf.data() + offsetof(DevPlan, nparts));
        for (int k = 0; k < K; ++k) nparts_out[k] = np[k];
    }
    h->last_fix_count = (int64_t)cnt[3];
    if (info_out) {
        unsigned long long below3 = cnt[0];
        if (h->xchg_world > 0) {
            below3 = 0;
            for (int q = 0; q < h->xchg_world; ++q) below3 += k3_below[q];
        }
        info_out[0] = P->status;
        info_out[1] = P->blocked;
        info_out[2] = K;
        info_out[3] = P->n_rays;
        info_out[4] = P->n_nan;
        info_out[5] = P->n_below;           // K0's global count of first samples below min(z)
        info_out[6] = (int64_t)below3;      // K3's own (global) count of the same: the cross-check
        info_out[7] = P->clamp_low_first;
        info_out[8] = (int64_t)cnt[1];      // samples below min(z) after the clamp decision (this rank)
        info_out[9] = (int64_t)cnt[2];      // samples above max(z) (this rank)
        info_out[10] = (int64_t)cnt[3];     // rays handed to the PROJ-form integrator (this rank)
        info_out[11] = P->knife_layer;
        info_out[12] = P->k_split;
        info_out[13] = P->nspan;
        info_out[14] = h->k0_was_cubic ? 1 : 0;
        info_out[15] = h->last_k3_poly ? 1 : 0;
        info_out[16] = (int64_t)cnt[4];     // CTA passes of the thin-layer kernel whose record columns were staged in shared memory
        info_out[17] = (int64_t)cnt[5];     // ... and those that read the records from global memory (footprint larger than the slots)
        info_out[18] = P->n_above;          // K0's global count of last samples above max(z)
        info_out[19] = P->clamp_high_last;
    }
    return RDR_OK;
}

RDR_API int rdr_set_peer_outputs(rdr_handle_t h, int n, void *const *wet, void *const *hydro) {
    CHECK_ARG(h, h != nullptr, "rdr_set_peer_outputs: NULL handle");
    CHECK_ARG(h, n >= 0 && n <= RDR_MAX_PEERS, "rdr_set_peer_outputs: at most 8 peer destinations");
    CHECK_ARG(h, n == 0 || (wet && hydro), "rdr_set_peer_outputs: NULL pointer list");
    h->peers.n = n;
    h->peers.multicast = 0;
    for (int i = 0; i < n; ++i) {
        CHECK_ARG(h, wet[i] && hydro[i], "rdr_set_peer_outputs: NULL destination");
        h->peers.wet[i] = wet[i];
        h->peers.hydro[i] = hydro[i];
    }
    return RDR_OK;
}

RDR_API int rdr_set_multicast_outputs(rdr_handle_t h, void *wet_mc, void *hydro_mc) {
    CHECK_ARG(h, h != nullptr, "rdr_set_multicast_outputs: NULL handle");
    CHECK_ARG(h, (wet_mc == nullptr) == (hydro_mc == nullptr), "rdr_set_multicast_outputs: both or neither");
    h->peers.n = 0;
    h->peers.multicast = wet_mc ? 1 : 0;
    h->peers.wet[0] = wet_mc;
    h->peers.hydro[0] = hydro_mc;
    return RDR_OK;
}

RDR_API int rdr_ray_stations(rdr_handle_t h, const double *lon, const double *lat, const double *hgt, int64_t n, int los_kind, const double *los,
                             double zref, double max_segment_length, double *out_wet, double *out_hydro, int32_t *out_nsamples, int mem) {
    CHECK_ARG(h, h != nullptr, "rdr_ray_stations: NULL handle");
    if (!h->has_cube) return fail(h, RDR_ERR_STATE, "rdr_ray_stations: no cube staged");
    CHECK_ARG(h, n >= 0 && (n == 0 || (lon && lat && hgt && out_wet && out_hydro)), "rdr_ray_stations: NULL pointer");
    CHECK_ARG(h, los_kind == RDR_LOS_ARRAY || los_kind == RDR_LOS_ENU_ARRAY || los_kind == RDR_LOS_ENU_CONST || los_kind == RDR_LOS_ZENITH ||
                     los_kind == RDR_LOS_ORBIT, "rdr_ray_stations: unknown los_kind");
    CHECK_ARG(h, los_kind == RDR_LOS_ZENITH || los != nullptr, "rdr_ray_stations: los is NULL");
    CHECK_ARG(h, max_segment_length > 0, "rdr_ray_stations: max_segment_length must be positive");
    if (n == 0) return RDR_OK;
    ScopedDevice sd(h->device);
    h->has_rays = false;  // the geometry scratch of the raster path is reused
    StationGeom S;
    int rc;
    DevBuf d_h;
    if ((rc = stage_in(h, h->d_gx, lon, n, mem, &S.lon))) return rc;
    if ((rc = stage_in(h, h->d_gy, lat, n, mem, &S.lat))) return rc;
    if ((rc = stage_in(h, d_h, hgt, n, mem, &S.hgt))) return rc;
    S.los = nullptr;
    S.los_kind = los_kind;
    S.e = S.n = 0.0;
    S.u = 1.0;
    if (los_kind == RDR_LOS_ARRAY || los_kind == RDR_LOS_ENU_ARRAY) {
        if ((rc = stage_in(h, h->d_los, los, n * 3, mem, &S.los))) return rc;
    } else if (los_kind == RDR_LOS_ENU_CONST) {
        S.e = los[0]; S.n = los[1]; S.u = los[2];
    } else if (los_kind == RDR_LOS_ORBIT) {
        const int64_t n_sv = (int64_t)los[0];
        std::vector<double> blob;
        if ((rc = split_orbit(h, los + 1, n_sv, blob))) return rc;
        CUDA_TRY(h, h->d_orbit.reserve(blob.size() * sizeof(double)));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_orbit.p, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, h->d_los.reserve((size_t)n * 3 * sizeof(double)));
        const double *ob = h->d_orbit.as<double>();
        const OrbitView O = {ob, ob + n_sv, ob + 4 * n_sv, (int)n_sv, (double)(n_sv - 1) / (blob[n_sv - 1] - blob[0])};
        k_orbit_los<<<grid_for(n, 128, h->sm_count, 16), 128, 0, h->stream>>>(O, RDR_GEOM_POINTS, S.lon, S.lat, S.hgt, 0.0, 1, n, 1.0e-7, 30,
                                                                             h->d_los.as<double>(), nullptr, nullptr);
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        S.los = h->d_los.as<double>();
        S.los_kind = RDR_LOS_ARRAY;
    }
    double *dw = out_wet, *dh = out_hydro;
    int32_t *dn = out_nsamples;
    if (mem == RDR_MEM_HOST) {
        CUDA_TRY(h, h->d_out.reserve(n * (2 * sizeof(double) + sizeof(int32_t))));
        dw = h->d_out.as<double>();
        dh = dw + n;
        dn = out_nsamples ? reinterpret_cast<int32_t *>(dh + n) : nullptr;
    }
    constexpr int BLOCK = 128;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n * 32 + BLOCK - 1) / BLOCK, (int64_t)h->sm_count * 16));
    k_ray_stations<BLOCK><<<grid, BLOCK, 0, h->stream>>>(make_view(h), S, n, zref, max_segment_length, dw, dh, dn);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    if (mem == RDR_MEM_HOST) {
        CUDA_TRY(h, cudaMemcpyAsync(out_wet, dw, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(out_hydro, dh, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (out_nsamples) CUDA_TRY(h, cudaMemcpyAsync(out_nsamples, dn, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    d_h.release();
    return RDR_OK;
}

RDR_API int rdr_ray_points(rdr_handle_t h, const double *maxlen, double max_segment_length, int64_t slot0, int64_t nslots, void *pts, int dtype,
                           int64_t *total_slots, int mem) {
    CHECK_ARG(h, h != nullptr, "rdr_ray_points: NULL handle");
    if (!h->has_cube || !h->has_rays) return fail(h, RDR_ERR_STATE, "rdr_ray_points: call rdr_ray_layers first");
    CHECK_ARG(h, maxlen && max_segment_length > 0, "rdr_ray_points: bad arguments");
    CHECK_ARG(h, dtype == RDR_F64 || dtype == RDR_F32, "rdr_ray_points: dtype must be RDR_F64 or RDR_F32");
    ScopedDevice sd(h->device);
    const int K = h->n_layers;
    const int64_t n = h->n_rays;
    std::vector<int> np(K);
    int64_t total = 1;
    for (int k = 0; k < K; ++k) {
        const double q = ceil(maxlen[k] / max_segment_length);
        CHECK_ARG(h, q == q && q < 1e7, "rdr_ray_points: per-layer max length is NaN or absurd");
        np[k] = std::max(2, (int)q + 1);
        total += np[k] - 1;
    }
    if (total_slots) *total_slots = total;
    if (!pts || nslots <= 0) return RDR_OK;  // count query
    CHECK_ARG(h, slot0 >= 0 && slot0 + nslots <= total, "rdr_ray_points: slot range out of bounds");
    CUDA_TRY(h, h->d_nparts.reserve(2 * K * sizeof(int)));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_nparts.p, np.data(), K * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    const size_t es = dtype == RDR_F64 ? 8 : 4;
    void *dp = pts;
    if (mem == RDR_MEM_HOST) {
        CUDA_TRY(h, h->d_in.reserve((size_t)nslots * n * 3 * es));
        dp = h->d_in.p;
    }
    constexpr int BLOCK = 128;
    const int grid = grid_for(n, BLOCK, h->sm_count, 16);
    if (dtype == RDR_F64)
        k_ray_points<double, BLOCK><<<grid, BLOCK, 0, h->stream>>>(make_view(h), make_geom(h), n, K, h->d_t.as<double>(), h->d_nparts.as<int>(),
                                                                   (int)slot0, (int)nslots, static_cast<double *>(dp));
    else
        k_ray_points<float, BLOCK><<<grid, BLOCK, 0, h->stream>>>(make_view(h), make_geom(h), n, K, h->d_t.as<double>(), h->d_nparts.as<int>(),
                                                                  (int)slot0, (int)nslots, static_cast<float *>(dp));
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    if (mem == RDR_MEM_HOST) CUDA_TRY(h, cudaMemcpyAsync(pts, dp, (size_t)nslots * n * 3 * es, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return RDR_OK;
}

// ------------------------------------------------------------------------------------------------
// handle-less helpers: create a transient context on `device`
// ------------------------------------------------------------------------------------------------
namespace {
struct Transient {
    int device;
    ScopedDevice sd;
    std::vector<void *> bufs;
    explicit Transient(int dev) : device(dev), sd(dev) {}
    ~Transient() {
        for (void *p : bufs) cudaFree(p);
    }
    template <typename T>
    cudaError_t in(const T *src, size_t count, int mem, const T **out) {
        if (mem == RDR_MEM_DEVICE) {
            *out = src;
            return cudaSuccess;
        }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16));
        if (e != cudaSuccess) return e;
        bufs.push_back(p);
        *out = static_cast<const T *>(p);
        return cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice);
    }
    template <typename T>
    cudaError_t out(T *dst, size_t count, int mem, T **dev) {
        if (mem == RDR_MEM_DEVICE) {
            *dev = dst;
            return cudaSuccess;
        }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16));
        if (e != cudaSuccess) return e;
        bufs.push_back(p);
        *dev = static_cast<T *>(p);
        return cudaSuccess;
    }
};

int need_device(int device) {
    const int n = rdr_device_count();
    if (n == 0) return fail(nullptr, RDR_ERR_CUDA, "no CUDA device available; libraider_b200 has no CPU fallback");
    if (device < 0 || device >= n) return fail(nullptr, RDR_ERR_INVALID, "device index out of range");
    return RDR_OK;
}
}  // namespace

#define T_TRY(expr) CUDA_TRY(nullptr, expr)

RDR_API int rdr_top_of_atmosphere(const double *xyz, const double *look, int64_t n, double toaheight, const double *factor, double *out_xyz,
                                  int device) {
    CHECK_ARG(nullptr, xyz && look && out_xyz && n >= 0, "rdr_top_of_atmosphere: bad arguments");
    int rc = need_device(device);
    if (rc || n == 0) return rc;
    Transient T(device);
    const double *dx, *dl, *df = nullptr;
    double *dout;
    T_TRY(T.in(xyz, 3 * n, RDR_MEM_HOST, &dx));
    T_TRY(T.in(look, 3 * n, RDR_MEM_HOST, &dl));
    if (factor) T_TRY(T.in(factor, n, RDR_MEM_HOST, &df));
    T_TRY(T.out(out_xyz, 3 * n, RDR_MEM_HOST, &dout));
    k_top_of_atmosphere<<<(unsigned)((n + 127) / 128), 128>>>(dx, dl, n, toaheight, df, dout);
    T_TRY(cudaGetLastError());
    T_TRY(cudaMemcpy(out_xyz, dout, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
    return RDR_OK;
}

RDR_API int rdr_build_ray(const double *model_zs, int64_t nz, double ht, const double *xyz, const double *look, int64_t n, double zref,
                          int64_t *n_layers, double *ray_lengths, double *low_xyzs, double *high_xyzs, int device) {
    CHECK_ARG(nullptr, model_zs && nz >= 2 && n_layers, "rdr_build_ray: bad arguments");
    std::vector<double> zs(model_zs, model_zs + nz), lo, hi;
    std::vector<int> cell;
    layer_plan(zs, ht, zref, lo, hi, cell);
    const int K = (int)lo.size();
    *n_layers = K;
    if (K == 0) return RDR_ERR_NO_LAYERS;
    if (!ray_lengths) return RDR_OK;  // count query
    CHECK_ARG(nullptr, xyz && look && low_xyzs && high_xyzs && n >= 0, "rdr_build_ray: bad arguments");
    int rc = need_device(device);
    if (rc || n == 0) return rc;
    Transient T(device);
    std::vector<double> plan(lo);
    plan.insert(plan.end(), hi.begin(), hi.end());
    const double *dx, *dl, *dp;
    double *dlen, *dlo, *dhi;
    T_TRY(T.in(xyz, 3 * n, RDR_MEM_HOST, &dx));
    T_TRY(T.in(look, 3 * n, RDR_MEM_HOST, &dl));
    T_TRY(T.in(plan.data(), plan.size(), RDR_MEM_HOST, &dp));
    T_TRY(T.out(ray_lengths, (size_t)K * n, RDR_MEM_HOST, &dlen));
    T_TRY(T.out(low_xyzs, (size_t)K * n * 3, RDR_MEM_HOST, &dlo));
    T_TRY(T.out(high_xyzs, (size_t)K * n * 3, RDR_MEM_HOST, &dhi));
    k_build_ray<<<(unsigned)((n + 127) / 128), 128>>>(dx, dl, n, K, dp, dlen, dlo, dhi);
    T_TRY(cudaGetLastError());
    T_TRY(cudaMemcpy(ray_lengths, dlen, (size_t)K * n * sizeof(double), cudaMemcpyDeviceToHost));
    T_TRY(cudaMemcpy(low_xyzs, dlo, (size_t)K * n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    T_TRY(cudaMemcpy(high_xyzs, dhi, (size_t)K * n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return RDR_OK;
}

RDR_API int rdr_lla2ecef(const double *lat, const double *lon, const double *hgt, int64_t n, double *x, double *y, double *z, int device) {
    CHECK_ARG(nullptr, lat && lon && hgt && x && y && z && n >= 0, "rdr_lla2ecef: bad arguments");
    int rc = need_device(device);
    if (rc || n == 0) return rc;
    Transient T(device);
    const double *a, *b, *c;
    double *dx, *dy, *dz;
    T_TRY(T.in(lat, n, 0, &a)); T_TRY(T.in(lon, n, 0, &b)); T_TRY(T.in(hgt, n, 0, &c));
    T_TRY(T.out(x, n, 0, &dx)); T_TRY(T.out(y, n, 0, &dy)); T_TRY(T.out(z, n, 0, &dz));
    k_lla2ecef<<<(unsigned)((n + 255) / 256), 256>>>(a, b, c, n, dx, dy, dz);
    T_TRY(cudaGetLastError());
    T_TRY(cudaMemcpy(x, dx, n * 8, cudaMemcpyDeviceToHost)); T_TRY(cudaMemcpy(y, dy, n * 8, cudaMemcpyDeviceToHost));
    T_TRY(cudaMemcpy(z, dz, n * 8, cudaMemcpyDeviceToHost));
    return RDR_OK;
}

RDR_API int rdr_ecef2lla(const double *x, const double *y, const double *z, int64_t n, double *lon, double *lat, double *hgt, int device) {
    CHECK_ARG(nullptr, lat && lon && hgt && x && y && z && n >= 0, "rdr_ecef2lla: bad arguments");
    int rc = need_device(device);
    if (rc || n == 0) return rc;
    Transient T(device);
    const double *a, *b, *c;
    double *dlo, *dla, *dh;
    T_TRY(T.in(x, n, 0, &a)); T_TRY(T.in(y, n, 0, &b)); T_TRY(T.in(z, n, 0, &c));
    T_TRY(T.out(lon, n, 0, &dlo)); T_TRY(T.out(lat, n, 0, &dla)); T_TRY(T.out(hgt, n, 0, &dh));
    k_ecef2lla<<<(unsigned)((n + 255) / 256), 256>>>(a, b, c, n, dlo, dla, dh);
    T_TRY(cudaGetLastError());
    T_TRY(cudaMemcpy(lon, dlo, n * 8, cudaMemcpyDeviceToHost)); T_TRY(cudaMemcpy(lat, dla, n * 8, cudaMemcpyDeviceToHost));
    T_TRY(cudaMemcpy(hgt, dh, n * 8, cudaMemcpyDeviceToHost));
    return RDR_OK;
}

RDR_API int rdr_prepare_cube(int64_t ncol, int64_t nl, const double *zs, const double *p, const double *t, const double *hum, int hum_is_rh,
                             const double *zlevels, int64_t nz_out, double k1, double k2, double k3, double zmin, float *wet, float *hydro,
                             float *wet_total, float *hydro_total, float *p_out, float *t_out, float *e_out, int64_t *nz_written,
                             int device, int mem) {
    CHECK_ARG(nullptr, zs && p && t && hum && zlevels && wet && hydro && wet_total && hydro_total, "rdr_prepare_cube: NULL pointer");
    CHECK_ARG(nullptr, ncol >= 1 && nl >= 2 && nl <= 4096 && nz_out >= 2 && nz_out <= 255, "rdr_prepare_cube: 2..4096 native levels, 2..255 target levels");
    CHECK_ARG(nullptr, (p_out == nullptr) == (t_out == nullptr) && (p_out == nullptr) == (e_out == nullptr), "rdr_prepare_cube: p/t/e outputs go together");
    for (int64_t i = 1; i < nz_out; ++i) CHECK_ARG(nullptr, zlevels[i] > zlevels[i - 1], "rdr_prepare_cube: target levels must be strictly ascending");
    int rc = need_device(device);
    if (rc) return rc;
    PrepParams P;
    P.nl = (int)nl; P.nz_out = (int)nz_out;
    P.pad = zmin < zlevels[0] ? 1 : 0;  // weatherModel.py:377
    P.hum_is_rh = hum_is_rh;
    P.k1 = k1; P.k2 = k2; P.k3 = k3; P.R_v = 461.524; P.R_d = 287.06; P.zmin = zmin;  // weatherModel.py:75-76
    const int64_t nzo = nz_out + P.pad;
    if (nz_written) *nz_written = nzo;
    Transient T(device);
    const double *dz, *dp, *dt, *dh, *dl;
    float *dw, *dhy, *dwt, *dht, *dpo = nullptr, *dto = nullptr, *deo = nullptr;
    T_TRY(T.in(zs, (size_t)ncol * nl, mem, &dz));
    T_TRY(T.in(p, (size_t)ncol * nl, mem, &dp));
    T_TRY(T.in(t, (size_t)ncol * nl, mem, &dt));
    T_TRY(T.in(hum, (size_t)ncol * nl, mem, &dh));
    T_TRY(T.in(zlevels, (size_t)nz_out, RDR_MEM_HOST, &dl));
    T_TRY(T.out(wet, (size_t)ncol * nzo, mem, &dw));
    T_TRY(T.out(hydro, (size_t)ncol * nzo, mem, &dhy));
    T_TRY(T.out(wet_total, (size_t)ncol * nzo, mem, &dwt));
    T_TRY(T.out(hydro_total, (size_t)ncol * nzo, mem, &dht));
    if (p_out) {
        T_TRY(T.out(p_out, (size_t)ncol * nzo, mem, &dpo));
        T_TRY(T.out(t_out, (size_t)ncol * nzo, mem, &dto));
        T_TRY(T.out(e_out, (size_t)ncol * nzo, mem, &deo));
    }
    const size_t per_warp = ((size_t)nl * sizeof(double) + (size_t)5 * nzo * sizeof(float) + 16 + 15) / 16 * 16;
    const size_t smem = 4 * per_warp;
    T_TRY(cudaFuncSetAttribute(k_prepare_columns, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ncol + 3) / 4, 148 * 8));
    k_prepare_columns<<<grid, 128, smem>>>(P, ncol, dz, dp, dt, dh, dl, dw, dhy, dwt, dht, dpo, dto, deo);
    T_TRY(cudaGetLastError());
    if (mem == RDR_MEM_HOST) {
        const size_t nb = (size_t)ncol * nzo * sizeof(float);
        T_TRY(cudaMemcpy(wet, dw, nb, cudaMemcpyDeviceToHost));
        T_TRY(cudaMemcpy(hydro, dhy, nb, cudaMemcpyDeviceToHost));
        T_TRY(cudaMemcpy(wet_total, dwt, nb, cudaMemcpyDeviceToHost));
        T_TRY(cudaMemcpy(hydro_total, dht, nb, cudaMemcpyDeviceToHost));
        if (p_out) {
            T_TRY(cudaMemcpy(p_out, dpo, nb, cudaMemcpyDeviceToHost));
            T_TRY(cudaMemcpy(t_out, dto, nb, cudaMemcpyDeviceToHost));
            T_TRY(cudaMemcpy(e_out, deo, nb, cudaMemcpyDeviceToHost));
        }
    } else {
        T_TRY(cudaDeviceSynchronize());
    }
    return RDR_OK;
}

RDR_API int rdr_orbit_los(const double *sv_rows, int64_t n_sv, int geom_kind, const double *gx, const double *gy, const double *hgt, double ht,
                          int64_t ny, int64_t nx, double threshold, int maxiter, double *out_los, double *out_slant, double *out_aztime, int device) {
    CHECK_ARG(nullptr, gx && gy && out_los && ny > 0 && nx > 0 && maxiter > 0 && threshold > 0, "rdr_orbit_los: bad arguments");
    CHECK_ARG(nullptr, geom_kind == RDR_GEOM_GRID || geom_kind == RDR_GEOM_POINTS, "rdr_orbit_los: unknown geom_kind");
    CHECK_ARG(nullptr, geom_kind == RDR_GEOM_POINTS || hgt == nullptr, "rdr_orbit_los: per-point heights need RDR_GEOM_POINTS");
    std::vector<double> blob;
    int rc = split_orbit(nullptr, sv_rows, n_sv, blob);
    if (rc) return rc;
    if ((rc = need_device(device))) return rc;
    Transient T(device);
    const int64_t n = ny * nx;
    const double *dob, *dx, *dy, *dh = nullptr;
    double *dlos, *dsl = nullptr, *daz = nullptr;
    T_TRY(T.in(blob.data(), blob.size(), RDR_MEM_HOST, &dob));
    T_TRY(T.in(gx, geom_kind == RDR_GEOM_GRID ? nx : n, RDR_MEM_HOST, &dx));
    T_TRY(T.in(gy, geom_kind == RDR_GEOM_GRID ? ny : n, RDR_MEM_HOST, &dy));
    if (hgt) T_TRY(T.in(hgt, n, RDR_MEM_HOST, &dh));
    T_TRY(T.out(out_los, 3 * n, RDR_MEM_HOST, &dlos));
    if (out_slant) T_TRY(T.out(out_slant, n, RDR_MEM_HOST, &dsl));
    if (out_aztime) T_TRY(T.out(out_aztime, n, RDR_MEM_HOST, &daz));
    const OrbitView O = {dob, dob + n_sv, dob + 4 * n_sv, (int)n_sv, (double)(n_sv - 1) / (blob[n_sv - 1] - blob[0])};
    k_orbit_los<<<(unsigned)std::min<int64_t>((n + 127) / 128, 148 * 16), 128>>>(O, geom_kind, dx, dy, dh, ht, (int)nx, n, threshold, maxiter, dlos, dsl, daz);
    T_TRY(cudaGetLastError());
    T_TRY(cudaMemcpy(out_los, dlos, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
    if (out_slant) T_TRY(cudaMemcpy(out_slant, dsl, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (out_aztime) T_TRY(cudaMemcpy(out_aztime, daz, n * sizeof(double), cudaMemcpyDeviceToHost));
    return RDR_OK;
}

RDR_API int rdr_selftest_div(int64_t n, uint64_t seed, int64_t *mismatch_1step, int64_t *mismatch_2step, int device) {
    CHECK_ARG(nullptr, mismatch_1step && mismatch_2step && n >= 0, "rdr_selftest_div: bad arguments");
    int rc = need_device(device);
    if (rc) return rc;
    Transient T(device);
    unsigned long long *dm = nullptr, hm[2] = {0, 0};
    T_TRY(cudaMalloc(&dm, sizeof(hm)));
    T.bufs.push_back(dm);
    T_TRY(cudaMemset(dm, 0, sizeof(hm)));
    k_selftest_div<<<148 * 8, 256>>>(n, seed, dm);
    T_TRY(cudaGetLastError());
    T_TRY(cudaMemcpy(hm, dm, sizeof(hm), cudaMemcpyDeviceToHost));
    *mismatch_1step = (int64_t)hm[0];
    *mismatch_2step = (int64_t)hm[1];
    return RDR_OK;
}

RDR_API int rdr_make_points_count(double max_len, double step, int64_t *npts) {
    CHECK_ARG(nullptr, npts != nullptr, "rdr_make_points_count: npts is NULL");
    CHECK_ARG(nullptr, step != 0.0, "float modulo");  // Python raises ZeroDivisionError('float modulo')
    *npts = make_points_npts(max_len, step);
    return RDR_OK;
}

RDR_API int rdr_make_points(double max_len, const double *sp, const double *slv, int64_t n_rays, double step, double *out, int64_t npts,
                            int device, int mem) {
    CHECK_ARG(nullptr, sp && slv && out && n_rays >= 0 && npts >= 0, "rdr_make_points: bad arguments");
    int rc = need_device(device);
    if (rc || n_rays == 0 || npts == 0) return rc;
    Transient T(device);
    const double *dsp, *dslv;
    double *dout;
    T_TRY(T.in(sp, 3 * n_rays, mem, &dsp));
    T_TRY(T.in(slv, 3 * n_rays, mem, &dslv));
    T_TRY(T.out(out, (size_t)3 * n_rays * npts, mem, &dout));
    const int64_t total = 3 * n_rays * npts;
    k_make_points<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 32), 256>>>(dsp, dslv, n_rays, step, npts, dout);
    T_TRY(cudaGetLastError());
    if (mem == RDR_MEM_HOST) T_TRY(cudaMemcpy(out, dout, total * sizeof(double), cudaMemcpyDeviceToHost));
    else T_TRY(cudaDeviceSynchronize());
    (void)max_len;
    return RDR_OK;
}

RDR_API int rdr_interpolate(int ndim, const double *const *grids, const int64_t *sizes, const double *values, const double *pts, int64_t n,
                            int has_fill, double fill_value, double *out, int device, int mem) {
    CHECK_ARG(nullptr, ndim >= 1 && ndim <= 8, "rdr_interpolate: 1 <= ndim <= 8 supported");
    CHECK_ARG(nullptr, grids && sizes && values && pts && out && n >= 0, "rdr_interpolate: NULL pointer");
    int rc = need_device(device);
    if (rc || n == 0) return rc;
    Transient T(device);
    NdGrid G;
    G.ndim = ndim;
    size_t nval = 1;
    for (int d = 0; d < ndim; ++d) {
        CHECK_ARG(nullptr, sizes[d] >= 1 && sizes[d] < (1ll << 31), "rdr_interpolate: bad grid size");
        T_TRY(T.in(grids[d], sizes[d], RDR_MEM_HOST, &G.g[d]));
        G.n[d] = (int)sizes[d];
        nval *= (size_t)sizes[d];
    }
    const double *dv, *dp;
    double *dout;
    T_TRY(T.in(values, nval, mem, &dv));
    T_TRY(T.in(pts, (size_t)n * ndim, mem, &dp));
    T_TRY(T.out(out, n, mem, &dout));
    k_interp_nd<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256>>>(G, dv, dp, n, has_fill, fill_value, dout);
    T_TRY(cudaGetLastError());
    if (mem == RDR_MEM_HOST) T_TRY(cudaMemcpy(out, dout, n * sizeof(double), cudaMemcpyDeviceToHost));
    else T_TRY(cudaDeviceSynchronize());
    return RDR_OK;
}

RDR_API int rdr_interp_along_axis(const double *x, const double *y, const double *xnew, int64_t ncol, int64_t nin, int64_t nout, int has_fill,
                                  double fill_value, double *out, int device, int mem) {
    CHECK_ARG(nullptr, x && y && xnew && out && ncol >= 0 && nin >= 1 && nout >= 0, "rdr_interp_along_axis: bad arguments");
    CHECK_ARG(nullptr, nin < (1ll << 31) && nout < (1ll << 31), "rdr_interp_along_axis: axis too long");
    int rc = need_device(device);
    if (rc || ncol == 0 || nout == 0) return rc;
    Transient T(device);
    const double *dx, *dy, *dq;
    double *dout;
    T_TRY(T.in(x, (size_t)ncol * nin, mem, &dx));
    T_TRY(T.in(y, (size_t)ncol * nin, mem, &dy));
    T_TRY(T.in(xnew, (size_t)ncol * nout, mem, &dq));
    T_TRY(T.out(out, (size_t)ncol * nout, mem, &dout));
    const int64_t total = ncol * nout;
    k_interp_axis<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16), 256>>>(dx, dy, dq, ncol, (int)nin, (int)nout, has_fill, fill_value,
                                                                                      dout);
    T_TRY(cudaGetLastError());
    if (mem == RDR_MEM_HOST) T_TRY(cudaMemcpy(out, dout, total * sizeof(double), cudaMemcpyDeviceToHost));
    else T_TRY(cudaDeviceSynchronize());
    return RDR_OK;
}

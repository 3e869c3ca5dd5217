// libraider_b200.so -- hand-written sm_100a kernels + C ABI for the RAiDER slant/zenith delay hot path.
// See include/raider_b200.h for the boundary and DESIGN.md for the kernel inventory (kernels live in the k*.cuh / plan.cuh fragments
// included below; this file keeps the handle, cube staging and the C ABI):
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

// ---- the device code, one fragment per kernel family (textually included here, inside the anonymous namespace, in dependency
// order: the translation unit -- and the SASS -- is what it was as a single 4300-line file) -------------------------------------
#include "k2_sampler.cuh"        // K2 k_sample_stream / _f32, mbarrier + TMA-bulk helpers
#include "k0_layers.cuh"         // ray geometry, K0 k_ray_layers
#include "plan.cuh"              // DevPlan, k_plan, k_publish
#include "k3_general.cuh"       // K3 k_ray_integrate (PROJ-form, list mode)
#include "k3_fast.cuh"          // K3 k_ray_integrate_fast
#include "k3_poly.cuh"          // K3 k_ray_integrate_poly (quadrature)
#include "k3_thin.cuh"          // K3 k_ray_integrate_thin (TMA-staged)
#include "k_points_stations.cuh"  // K1b k_ray_points, K5 k_ray_stations
#include "k7_prepare.cuh"       // K7 k_prepare_columns
#include "k6_orbit.cuh"         // K6 k_orbit_los
#include "k_parity.cuh"         // K1 / K4 and the small API-parity kernels
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

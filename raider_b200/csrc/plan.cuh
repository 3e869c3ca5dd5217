// plan.cuh -- the step plan built on the device between K0 and K3 (k_plan, k_publish).
// A fragment of libraider_b200.so: included by raider_b200.cu INSIDE its anonymous namespace, in this order (the translation unit
// is the one file it used to be; see the kernel inventory at the top of raider_b200.cu).  Not a stand-alone header.
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


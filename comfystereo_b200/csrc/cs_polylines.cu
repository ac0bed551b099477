// cs_polylines.cu -- P: apply_stereo_divergence_polylines (SIG:1912-1992), soft and sharp.
//
// One CTA per (row, frame, eye); the whole row lives in shared memory.  Two kernels:
//
// k_polylines<PER>  the fast path (rows up to ~4090 px sharp / ~8190 px soft)
//   A points   thread per source column: coord_d in FP64, x = col + 0.5 + coord_d + sep rounded to float32,
//              closeness |coord_d| float32; sharp emits x -+ 0.45.  Sentinels (-W, 0) and (2W, 0).   SIG:1919-1936
//   B,C sort   the reference's stable insertion sort by x (SIG:1941-1946) without sorting: shifts are bounded, so the
//              table is nearly sorted.  Two CTA scans (prefix max / suffix min in source order) tell every point
//              whether anything before it is larger or anything after it smaller; if not, its rank is its source
//              index.  The others go to a work list and count their inversions in the window the scans bound.
//   D,D2 sets  the reference's "active list" at a centre is the SET of segments with x0 < ctr <= x1.  A prefix max
//              of segment ends (sorted order) bounds the search.  Per sorted interval the set is constant; intervals
//              with one candidate, or two whose order cannot change inside the interval, are resolved here.
//   E sweep    thread per output column; the sub-intervals of column c are (pred, b0), (b0, b1), ... of the sorted
//              points in bucket c (SIG:1955-1961).  Selection and colour accumulation reproduce the reference's
//              FP64 / float32 rounding sequence.  Choices that depend on the ORDER of the reference's append /
//              swap-remove list (exact ties, no valid candidate -- quirk Q7) are replayed from the nearest visit
//              with a single active segment; if that does not fit its budget the row is redone by
//              sequential_row(), one thread replaying the reference's sweep bit for bit.
//
// k_polylines_exact  counting sort + one-thread sequential sweep for every row: serves rows too wide for the fast
//              kernel's tables and is the independent cross-check of the fast path in the tests.
//
// Bytes per pixel and eye: depth 4 B + RGBX8 4 B read (L2-resident scratch), RGBX8 4 B written.
#include "cs_internal.cuh"

namespace cs {

namespace {

constexpr int kPolyThreads = 512;
constexpr double kEps = 1e-7;

struct RowCtx {
    int w, npts, nsg;   // npts = points incl. both sentinels, nsg = npts - 1 segments
    bool sharp;
};

// source point index -> source column (the "s" field of the reference's point table)
__device__ __forceinline__ int pt_col(int i, const RowCtx& c) {
    if (i <= 0) return 0;
    if (i >= c.npts - 1) return c.w - 1;
    return c.sharp ? ((i - 1) >> 1) : (i - 1);
}
__device__ __forceinline__ float pt_clo(int i, const RowCtx& c, const float* clo) {
    if (i <= 0 || i >= c.npts - 1) return 0.0f;
    return clo[c.sharp ? ((i - 1) >> 1) : (i - 1)];
}
// branch-free variants on a table padded with the two sentinels: clo2[0] = 0, clo2[1 + col], clo2[w + 1] = 0
__device__ __forceinline__ int pt_slot(int i, bool sharp) { return sharp ? ((i + 1) >> 1) : i; }
__device__ __forceinline__ int slot_col(int slot, int w) { return min(max(slot - 1, 0), w - 1); }

__device__ __forceinline__ Normalizer pl_normalizer(const WarpArgs& a, int eye, int frame, float* scale_out) {
    const FrameStats st = a.stats[frame];
    float scale = 1.0f;
    int lo, hi;
    if (a.use_blur_stats) {
        lo = eye ? st.r_min : st.l_min;
        hi = eye ? st.r_max : st.l_max;
    } else {
        lo = st.gray_min; hi = st.gray_max;
        if (a.scale_by_stats && ord2f(st.gray_max) <= 1.0f) scale = 255.0f;
    }
    *scale_out = scale;
    return make_normalizer(lo, hi, scale, a.conv);
}

// CTA-wide exclusive scan of cnt[0..n) in place (n arbitrary), returns nothing; blockDim = kPolyThreads.
__device__ void cta_exclusive_scan(int* cnt, int n, int* s_warp) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + kPolyThreads - 1) / kPolyThreads;
    const int b0 = tid * per, b1 = min(b0 + per, n);
    int sum = 0;
    for (int i = b0; i < b1; ++i) sum += cnt[i];
    int inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int v = (lane < kPolyThreads / 32) ? s_warp[lane] : 0;
        int vi = v;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += t;
        }
        if (lane < kPolyThreads / 32) s_warp[lane] = vi - v;
    }
    __syncthreads();
    int run = s_warp[wid] + inc - sum;
    for (int i = b0; i < b1; ++i) { int c = cnt[i]; cnt[i] = run; run += c; }
    __syncthreads();
}

// Builds the point table and its stable sort.  On return (after a barrier):
//   px[i]     float32 x of source point i                         [npts]
//   clo[col]  float32 closeness of source column col              [w]
//   sidx[k]   source point index of the k-th point in sorted order [npts]
//   start[b]  rank of the first point of bucket b, b = floor(x)+1 clamped to [0, w+1];  start[w+2] = npts
// tmp is a scratch array of npts uint16.
__device__ void build_sorted_points(const WarpArgs& a, int eye, int frame, int y, const RowCtx& c,
                                    float* px, float* clo, unsigned short* sidx, unsigned short* tmp,
                                    unsigned short* rnk, int* start, int* s_warp) {
    const int w = c.w, npts = c.npts;
    const double div_px = a.eye[eye].div_px, sep_px = a.eye[eye].sep_px;
    float scale;
    const Normalizer norm = pl_normalizer(a, eye, frame, &scale);
    const float* dep = a.depth[eye] + (int64_t)frame * a.h * w + (int64_t)y * w;
    const int nb = w + 2;
    for (int b = threadIdx.x; b <= nb; b += blockDim.x) start[b] = 0;
    if (threadIdx.x == 0) {
        px[0] = (float)(-1.0 * w);
        px[npts - 1] = (float)(2.0 * w);
    }
    for (int col = threadIdx.x; col < w; col += blockDim.x) {
        float d = dep[col];
        if (scale != 1.0f) d = d * scale;
        double cd = signed_pow_offset(norm(d), a.expo, div_px);
        double cx = ((double)col + 0.5) + cd;
        cx = cx + sep_px;
        clo[col] = (float)fabs(cd);
        if (c.sharp) {
            px[1 + 2 * col] = (float)(cx - 0.45);
            px[2 + 2 * col] = (float)(cx + 0.45);
        } else {
            px[1 + col] = (float)cx;
        }
    }
    __syncthreads();
    // histogram; slot order inside a bucket is arbitrary here
    for (int i = threadIdx.x; i < npts; i += blockDim.x) {
        float fl = floorf(px[i]);
        int b = (fl < 0.0f) ? 0 : ((fl >= (float)w) ? w + 1 : (int)fl + 1);
        rnk[i] = (unsigned short)atomicAdd(&start[b], 1);
    }
    __syncthreads();
    cta_exclusive_scan(start, nb + 1, s_warp);  // start[nb] = npts
    for (int i = threadIdx.x; i < npts; i += blockDim.x) {
        float fl = floorf(px[i]);
        int b = (fl < 0.0f) ? 0 : ((fl >= (float)w) ? w + 1 : (int)fl + 1);
        tmp[start[b] + rnk[i]] = (unsigned short)i;
    }
    __syncthreads();
    // rank inside the bucket by (x, source index): equals the reference's stable insertion sort
    for (int i = threadIdx.x; i < npts; i += blockDim.x) {
        float xi = px[i];
        float fl = floorf(xi);
        int b = (fl < 0.0f) ? 0 : ((fl >= (float)w) ? w + 1 : (int)fl + 1);
        int s0 = start[b], s1 = start[b + 1];
        int r = s0;
        for (int q = s0; q < s1; ++q) {
            int j = tmp[q];
            float xj = px[j];
            r += (xj < xi || (xj == xi && j < i)) ? 1 : 0;
        }
        rnk[i] = (unsigned short)r;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npts; i += blockDim.x) sidx[rnk[i]] = (unsigned short)i;
    __syncthreads();
}

// colour accumulation of one sub-interval, SIG:1981-1989 (float32 accumulator, float64 terms)
__device__ __forceinline__ void accumulate(float* color, uint32_t pl, uint32_t pr, bool same, double ip,
                                           double sig) {
    if (same) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            double term = (double)((pl >> (8 * ch)) & 255u) * sig;
            color[ch] = (float)((double)color[ch] + term);
        }
    } else {
        double om = 1.0 - ip;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            double t0 = (double)((pl >> (8 * ch)) & 255u) * om, t1 = (double)((pr >> (8 * ch)) & 255u) * ip;
            double mix = t0 + t1;
            double term = mix * sig;
            color[ch] = (float)((double)color[ch] + term);
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// exact sweep: one thread replays the reference's list semantics
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPolyThreads) k_polylines_exact(const WarpArgs a, int sharp, int act_cap,
                                                                  const int* __restrict__ row_flags,
                                                                  int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int w = a.w, y = blockIdx.x, frame = blockIdx.y, eye = blockIdx.z;
    if (a.eye[eye].passthrough) return;
    if (row_flags && !row_flags[((int64_t)frame * 2 + eye) * a.h + y]) return;
    RowCtx c;
    c.w = w; c.sharp = sharp != 0; c.npts = (sharp ? 2 * w : w) + 2; c.nsg = c.npts - 1;
    const int npts = c.npts, nsg = c.nsg;
    float* px = reinterpret_cast<float*>(smem_raw);
    float* clo = px + npts;
    int* start = reinterpret_cast<int*>(clo + w);
    unsigned short* sidx = reinterpret_cast<unsigned short*>(start + (w + 4));
    unsigned short* tmp = sidx + (npts + (npts & 1));
    unsigned short* rnk = tmp + (npts + (npts & 1));
    unsigned short* act = rnk + (npts + (npts & 1));   // [act_cap] source point indices of active segments
    __shared__ int s_warp[32];
    build_sorted_points(a, eye, frame, y, c, px, clo, sidx, tmp, rnk, start, s_warp);

    if (threadIdx.x != 0) return;
    const int64_t row_off = (int64_t)frame * a.h * w + (int64_t)y * w;
    const uint32_t* img = a.image_u8 + row_off;
    uint32_t* out = a.out[eye] + row_off;
    int nact = 0, sgp = 0, pi = 0;
    bool overflow = false;
    for (int col = 0; col < w; ++col) {
        float color[3] = {0.5f, 0.5f, 0.5f};
        while ((double)px[sidx[pi]] < (double)col) ++pi;
        --pi;
        while ((double)px[sidx[pi]] < (double)(col + 1)) {
            double pa = (double)px[sidx[pi]], pb = (double)px[sidx[pi + 1]];
            double from = fmax((double)col, pa) + kEps;
            double to = fmin((double)(col + 1), pb) - kEps;
            double sig = to - from;
            double ctr = from + 0.5 * sig;
            while (sgp < nsg && (double)px[sidx[sgp]] < ctr) {
                if (nact < act_cap) act[nact++] = sidx[sgp];
                else overflow = true;
                ++sgp;
            }
            for (int i = 0; i < nact;) {
                if ((double)px[act[i] + 1] < ctr) { act[i] = act[nact - 1]; --nact; }
                else ++i;
            }
            int best = 0;
            if (nact != 1) {
                double bestc = -kEps;
                for (int i = 0; i < nact; ++i) {
                    int sp = act[i];
                    float x0 = px[sp], x1 = px[sp + 1];
                    float den = x1 - x0;
                    double ip = (ctr - (double)x0) / (double)den;
                    double t0 = (1.0 - ip) * (double)pt_clo(sp, c, clo), t1 = ip * (double)pt_clo(sp + 1, c, clo);
                    double cl = t0 + t1;
                    if (bestc < cl && 0.0 < ip && ip < 1.0) { bestc = cl; best = i; }
                }
            }
            if (nact > 0) {
                int sp = act[best];
                int cl = pt_col(sp, c), cr = pt_col(sp + 1, c);
                double ip = 0.0;
                if (cl != cr) {
                    float den = px[sp + 1] - px[sp];
                    ip = (ctr - (double)px[sp]) / (double)den;
                }
                accumulate(color, img[cl], img[cr], cl == cr, ip, sig);
            }
            ++pi;
        }
        out[col] = pack_rgbx((int)color[0], (int)color[1], (int)color[2]);
    }
    if (overflow) atomicOr(status, 1);
}

// ------------------------------------------------------------------------------------------
// fast sweep
// ------------------------------------------------------------------------------------------
// Conversions between float32 and float64 (F2F / I2F) issue on the 16-lane XU pipe and were the
// bottleneck of the first version of this kernel (ncu: xu pipe saturated, fp64 pipe 12 % busy).
// The hot loop therefore never converts: the sorted coordinates are widened once per point, uint8
// colours are widened with a 2^52 bias trick, and float64 sums are rounded to float32 precision on
// the FP64 pipe.

// Round-to-nearest of a float64 to 24 significant bits, result kept as float64: equals
// (double)(float)x whenever (float)x is a normal float32.  Veltkamp / Dekker split with 2^29 + 1
// (three FP64 operations).  It differs from IEEE round-half-even only on exact ties (low 29 bits
// = 100...0), which an accumulated colour sum hits with probability 2^-29 per operation.
__device__ __forceinline__ double round24_fp(double x) {
    const double g = x * 536870913.0;
    const double d = x - g;
    return g + d;
}
// The same rounding with exact IEEE ties-to-even, in integer arithmetic.  Used where ties are NOT improbable: the
// float32 subtraction x1 - x0 of two float32 coordinates is a short exact binary number, and when it needs 25 bits
// (small coordinates, first columns of a row) it is an exact tie half of the time.
__device__ __forceinline__ double round24_even(double x) {
    uint32_t hi = (uint32_t)__double2hiint(x), lo = (uint32_t)__double2loint(x);
    uint32_t nlo = lo + 0x0FFFFFFFu + ((lo >> 29) & 1u);
    hi += (nlo < lo) ? 1u : 0u;
    return __hiloint2double((int)hi, (int)(nlo & 0xE0000000u));
}
// uint8 -> float64 without I2F: 2^52 + v is exact, subtracting 2^52 leaves v
__device__ __forceinline__ double u8_to_f64(uint32_t v) {
    return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}

struct PolyCtx {
    const float* px;        // [npts] source-order x (float32)
    const double* sxd;      // [npts] sorted x, widened
    const unsigned short* sidx;   // [npts] sorted rank -> source point index
    const float* reach;     // [npts] prefix max (sorted order) of segment ends
    const float* clo;       // [w + 2] padded: clo[pt_slot(i)] is the closeness of source point i
    const int* start;       // [tw+3] first sorted rank of bucket b = floor(x) - t0 + 1
    int t0;                 // first output column of the tile (0 when the CTA owns the whole row)
    RowCtx row;             // describes the tile's SOURCE window: w = window width, points / segments of the window
};
// cand[k] (k = sorted rank, uint16): bits 0-1 number of segments active anywhere strictly inside the interval
// (sorted point k, sorted point k+1), 3 = "three or more / does not fit"; bits 2-8 and 9-15: how many ranks back the
// first / second active segment starts.  Kept apart from sidx[] so that the threads resolving intervals (writers of
// cand) never touch words other threads are reading (sidx) -- racecheck-clean.
constexpr int kOff1Shift = 2, kOff2Shift = 9;

// `col` is the output column relative to the tile
__device__ __forceinline__ double visit_ctr(const PolyCtx& c, int col, int k, double* sig_out) {
    double pa = c.sxd[k], pb = c.sxd[k + 1];
    double from = fmax((double)(col + c.t0), pa) + kEps;
    double to = fmin((double)(col + c.t0 + 1), pb) - kEps;
    double sig = to - from;
    *sig_out = sig;
    return from + 0.5 * sig;
}

// number of active segments at ctr (interval k); *which = sorted index of the last one found
__device__ __forceinline__ int active_count(const PolyCtx& c, int k, double ctr, int* which) {
    int n = 0;
    for (int j = k; j >= 0 && !((double)c.reach[j] < ctr); --j) {
        int sp = (int)c.sidx[j];
        if (!(c.sxd[j] < ctr) || ((double)c.px[sp + 1] < ctr)) continue;
        ++n;
        *which = j;
    }
    return n;
}

// The reference's selection when the result depends on the ORDER of its active list (quirk Q7):
// find the nearest earlier visit with exactly one active segment (there the list is [that segment],
// whatever happened before), replay the append / swap-remove list from there to the target visit,
// then choose as the reference does.  Returns the source point index of the chosen segment, or -1
// when the replay does not fit its budget (the whole row is then redone sequentially).
__device__ __noinline__ int replay_choice(const PolyCtx& c, int col, int k) {
    constexpr int kCap = 64, kBudget = 6000;
    unsigned short lst[kCap];
    const int nsg = c.row.nsg;
    // ---- backward: locate the reset visit
    int rc = col, rk = k, sgp = 0, steps = 0;
    bool from_row_start = false;
    while (true) {
        // previous visit
        if (rk > c.start[rc + 1] - 1) --rk;
        else if (rc > 0) { --rc; rk = c.start[rc + 2] - 1; }
        else if (c.t0 == 0) { from_row_start = true; break; }
        else return -1;   // the history continues left of this tile: the whole row is replayed instead
        if (++steps > kBudget) return -1;
        double sig;
        double ctr = visit_ctr(c, rc, rk, &sig);
        int which = 0;
        if (active_count(c, rk, ctr, &which) == 1) { sgp = which; break; }
    }
    int n = 0;
    if (from_row_start) { rc = 0; rk = c.start[1] - 1; sgp = 0; }
    // ---- forward: replay list maintenance up to and including the target visit
    while (true) {
        double sig;
        double ctr = visit_ctr(c, rc, rk, &sig);
        while (sgp < nsg && c.sxd[sgp] < ctr) {
            if (n >= kCap) return -1;
            lst[n++] = c.sidx[sgp];
            ++sgp;
        }
        for (int i = 0; i < n;) {
            if ((double)c.px[lst[i] + 1] < ctr) { lst[i] = lst[n - 1]; --n; }
            else ++i;
        }
        if (rc == col && rk == k) {
            if (n == 0) return -1;
            int best = 0;
            if (n != 1) {
                double bestc = -kEps;
                for (int i = 0; i < n; ++i) {
                    int sp = lst[i];
                    float x0 = c.px[sp], x1 = c.px[sp + 1];
                    float den = x1 - x0;
                    double ip = (ctr - (double)x0) / (double)den;
                    double t0 = (1.0 - ip) * (double)c.clo[pt_slot(sp, c.row.sharp)], t1 = ip * (double)c.clo[pt_slot(sp + 1, c.row.sharp)];
                    double cl = t0 + t1;
                    if (bestc < cl && 0.0 < ip && ip < 1.0) { bestc = cl; best = i; }
                }
            }
            return lst[best];
        }
        // next visit
        if (rk < c.start[rc + 2] - 1) ++rk;
        else { ++rc; rk = c.start[rc + 1] - 1; }
    }
}

// Any visit that is not "exactly one active segment, the one starting at the interval's left point":
// builds the active set, selects in FP64 like the reference, and resolves order-dependent choices by replay.
// Returns the source point index of the chosen segment (-1: nothing active), -2: give up (row flagged).
__device__ __noinline__ int general_visit(const PolyCtx& c, int col, int k, double ctr) {
    int nact = 0, best = -1, only = -1, nbest = 0;
    double bestc = -kEps;
    for (int j = k; j >= 0 && !((double)c.reach[j] < ctr); --j) {
        int sp = (int)c.sidx[j];
        float x0 = c.px[sp], x1 = c.px[sp + 1];
        if (!((double)x0 < ctr) || ((double)x1 < ctr)) continue;
        ++nact;
        only = sp;
        float den = x1 - x0;
        double ip = (ctr - (double)x0) / (double)den;
        double t0 = (1.0 - ip) * (double)c.clo[pt_slot(sp, c.row.sharp)], t1 = ip * (double)c.clo[pt_slot(sp + 1, c.row.sharp)];
        double cl = t0 + t1;
        if (0.0 < ip && ip < 1.0) {
            if (bestc < cl) { bestc = cl; best = sp; nbest = 1; }
            else if (bestc == cl) ++nbest;
        }
    }
    if (nact == 0) return -1;
    if (nact == 1) return only;
    if (best >= 0 && nbest == 1) return best;
    int r = replay_choice(c, col, k);
    return r < 0 ? -2 : r;
}

// The reference's sequential sweep (SIG:1948-1991) over one row whose sorted point table is in shared
// memory; run by ONE thread.  act: scratch for the active list (capacity act_cap).
__device__ __noinline__ bool sequential_row(const PolyCtx& c, const uint32_t* img, uint32_t* out,
                                            unsigned short* act, int act_cap) {
    const int w = c.row.w, nsg = c.row.nsg;
    int nact = 0, sgp = 0, pi = 0;
    bool overflow = false;
    for (int col = 0; col < w; ++col) {
        float color[3] = {0.5f, 0.5f, 0.5f};
        while (c.sxd[pi] < (double)col) ++pi;
        --pi;
        while (c.sxd[pi] < (double)(col + 1)) {
            double sig;
            double ctr = visit_ctr(c, col, pi, &sig);
            while (sgp < nsg && c.sxd[sgp] < ctr) {
                if (nact < act_cap) act[nact++] = c.sidx[sgp];
                else overflow = true;
                ++sgp;
            }
            for (int i = 0; i < nact;) {
                if ((double)c.px[act[i] + 1] < ctr) { act[i] = act[nact - 1]; --nact; }
                else ++i;
            }
            int best = 0;
            if (nact != 1) {
                double bestc = -kEps;
                for (int i = 0; i < nact; ++i) {
                    int sp = act[i];
                    float x0 = c.px[sp], x1 = c.px[sp + 1];
                    float den = x1 - x0;
                    double ip = (ctr - (double)x0) / (double)den;
                    double t0 = (1.0 - ip) * (double)c.clo[pt_slot(sp, c.row.sharp)], t1 = ip * (double)c.clo[pt_slot(sp + 1, c.row.sharp)];
                    double cl = t0 + t1;
                    if (bestc < cl && 0.0 < ip && ip < 1.0) { bestc = cl; best = i; }
                }
            }
            if (nact > 0) {
                int sp = act[best];
                int cl = pt_col(sp, c.row), cr = pt_col(sp + 1, c.row);
                double ip = 0.0;
                if (cl != cr) {
                    float den = c.px[sp + 1] - c.px[sp];
                    ip = (ctr - (double)c.px[sp]) / (double)den;
                }
                accumulate(color, img[cl], img[cr], cl == cr, ip, sig);
            }
            ++pi;
        }
        out[col] = pack_rgbx((int)color[0], (int)color[1], (int)color[2]);
    }
    return overflow;
}

// One CTA per (row, frame, eye), 512 threads, PER points per thread (512 * PER >= points of the row).
// tile_w == 0: the CTA owns the whole row.  tile_w > 0 (rows too wide for the shared-memory tables): blockIdx.z also
// enumerates tiles of tile_w output columns; the CTA builds the point table of the SOURCE window that can reach its
// tile (|shift| <= reach_px[eye], plus a guard band) and sweeps only its own output columns.  Everything that decides
// a centre inside the tile -- the segments active there, their sorted order -- lies inside the window, so the result is
// the whole-row result; the artificial sentinel segments at the window's ends only cover columns outside the tile.
template <int PER, bool SHARP, bool TILED>   // compile-time: polylines_sharp (two points per source pixel); tiles of a wide row
__global__ void __launch_bounds__(kPolyThreads, 2) k_polylines(const WarpArgs a, int* __restrict__ row_flags,
                                                               int* __restrict__ status, int tile_w, int tile_ext,
                                                               double reach0, double reach1) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NP = kPolyThreads * PER;
    const int W = a.w, y = blockIdx.x, frame = blockIdx.y, eye = blockIdx.z & 1, tile = blockIdx.z >> 1;
    if (a.eye[eye].passthrough) return;
    // t0 / tw: origin and width of the column range the point buckets cover; the sweep writes its last `own` columns.
    // In tile mode the buckets start tile_ext columns left of the tile so that a list replay (quirk Q7) can walk back
    // through a whole fold to the nearest visit with a single active segment.
    int t0 = 0, tw = W, own = W, s0 = 0, w = W;
    if (!TILED) tile_w = 0;
    if (TILED) {
        const int o0 = tile * tile_w;
        own = min(tile_w, W - o0);
        t0 = max(o0 - tile_ext, 0);
        tw = o0 + own - t0;
        const double sep = a.eye[eye].sep_px, rch = (eye ? reach1 : reach0) + 4.0;
        const double lo = floor((double)t0 - sep - rch), hi = ceil((double)(t0 + tw) - sep + rch);
        s0 = (int)fmax(lo, 0.0);
        const int s1 = (int)fmin(hi, (double)W);
        w = max(s1 - s0, 1);
        if (s0 + w > W) s0 = W - w;
    }
    RowCtx c;
    c.w = w; c.sharp = SHARP; c.npts = (SHARP ? 2 * w : w) + 2; c.nsg = c.npts - 1;
    const int npts = c.npts, nsg = c.nsg;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float* px = reinterpret_cast<float*>(smem_raw);                 // [NP]
    double* sxd = reinterpret_cast<double*>(px + NP);               // [NP]   (aliases pm / sm during the sort)
    float* pm = reinterpret_cast<float*>(sxd);                      // [NP] inclusive prefix max of px
    float* sm = pm + NP;                                            // [NP] inclusive suffix min of px
    unsigned short* sidx = reinterpret_cast<unsigned short*>(sxd + NP);   // [NP] sorted rank -> source point
    unsigned short* cand = sidx + NP;                                     // [NP] per-interval candidate code
    float* reach = reinterpret_cast<float*>(cand + NP);             // [NP]
    float* clo = reach + NP;                                        // [w + 2] padded closeness table
    int* start = reinterpret_cast<int*>(clo + (w + 4));             // [w + 4]
    uint32_t* simg = reinterpret_cast<uint32_t*>(start + (w + 4));  // [w]
    unsigned short* tlist = reinterpret_cast<unsigned short*>(simg + w);  // [NP] sorted intervals with two active segments
    __shared__ float s_wa[16], s_wb[16];
    __shared__ int s_flag, s_ndirty, s_ntwo, s_next;
    if (tid == 0) { s_flag = 0; s_ndirty = 0; s_ntwo = 0; s_next = 0; }
    // during the sort the reach[] region holds the list of out-of-order points and their ranks (uint16 each)
    unsigned short* dlist = reinterpret_cast<unsigned short*>(reach);
    unsigned short* drank = dlist + NP;

    // ---- A: image row, points
    const int64_t row_off = (int64_t)frame * a.h * W + (int64_t)y * W;
    {
        const uint32_t* img = a.image_u8 + row_off + s0;
        const double div_px = a.eye[eye].div_px, sep_px = a.eye[eye].sep_px;
        float scale;
        const Normalizer norm = pl_normalizer(a, eye, frame, &scale);
        const float* dep = a.depth[eye] + row_off + s0;
        constexpr int kMaxIter = 4;
        for (int cbase = 0; cbase < w; cbase += kMaxIter * kPolyThreads) {
        float dreg[kMaxIter];
        uint32_t ireg[kMaxIter];
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const int col = cbase + tid + it * kPolyThreads;
            if (col < w) { dreg[it] = dep[col]; ireg[it] = img[col]; }
        }
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const int col = cbase + tid + it * kPolyThreads;
            if (col >= w) break;
            simg[col] = ireg[it];
            float d = dreg[it];
            if (scale != 1.0f) d = d * scale;
            float nd = norm(d);
            double an = (double)fabsf(nd);
            double p;
            if (a.expo == 2.0) p = an * an;
            else if (a.expo == 1.0) p = an;
            else p = pow(an, a.expo);
            double sp = (nd >= 0.0f) ? p : -p;
            double cd = sp * div_px;
            double cx = ((double)(col + s0) + 0.5) + cd;
            cx = cx + sep_px;
            clo[col + 1] = (float)fabs(cd);
            if (SHARP) {
                px[1 + 2 * col] = (float)(cx - 0.45);
                px[2 + 2 * col] = (float)(cx + 0.45);
            } else {
                px[1 + col] = (float)cx;
            }
        }
        }
        if (tid == 0) { px[0] = (float)(-1.0 * W); clo[0] = 0.0f; clo[w + 1] = 0.0f; }
        for (int i = npts - 1 + tid; i < NP; i += kPolyThreads) px[i] = (i == npts - 1) ? (float)(2.0 * W) : INFINITY;
        for (int b = tid; b < tw + 4; b += kPolyThreads) start[b] = 0;
    }
    __syncthreads();

    // ---- B: inclusive prefix max / suffix min of px in source order (thread t owns points t*PER .. t*PER+PER-1)
    float v[PER], lmax[PER], lmin[PER];
    const int i0 = tid * PER;
    {
#pragma unroll
        for (int q = 0; q < PER / 4; ++q) {
            float4 t = reinterpret_cast<const float4*>(px + i0)[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
        float m = -INFINITY;
#pragma unroll
        for (int e = 0; e < PER; ++e) { m = fmaxf(m, v[e]); lmax[e] = m; }
        float n = INFINITY;
#pragma unroll
        for (int e = PER - 1; e >= 0; --e) { n = fminf(n, v[e]); lmin[e] = n; }
        // warp-level exclusive scans of the per-thread totals
        float im = m, in_ = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            float t = __shfl_up_sync(0xffffffffu, im, o);
            if (lane >= o) im = fmaxf(im, t);
            float u = __shfl_down_sync(0xffffffffu, in_, o);
            if (lane + o < 32) in_ = fminf(in_, u);
        }
        if (lane == 31) s_wa[wid] = im;
        if (lane == 0) s_wb[wid] = in_;
        float em = __shfl_up_sync(0xffffffffu, im, 1);    // max over earlier lanes of the warp
        float en = __shfl_down_sync(0xffffffffu, in_, 1); // min over later lanes
        if (lane == 0) em = -INFINITY;
        if (lane == 31) en = INFINITY;
        __syncthreads();
        for (int q = 0; q < wid; ++q) em = fmaxf(em, s_wa[q]);
        for (int q = wid + 1; q < kPolyThreads / 32; ++q) en = fminf(en, s_wb[q]);
#pragma unroll
        for (int e = 0; e < PER; ++e) { lmax[e] = fmaxf(lmax[e], em); lmin[e] = fminf(lmin[e], en); }
#pragma unroll
        for (int q = 0; q < PER / 4; ++q) {
            reinterpret_cast<float4*>(pm + i0)[q] = make_float4(lmax[4 * q], lmax[4 * q + 1], lmax[4 * q + 2], lmax[4 * q + 3]);
            reinterpret_cast<float4*>(sm + i0)[q] = make_float4(lmin[4 * q], lmin[4 * q + 1], lmin[4 * q + 2], lmin[4 * q + 3]);
        }
        // lmax[e] / lmin[e] now hold the INCLUSIVE scans; em / en the exclusive values of the thread's first / last point
        __syncthreads();
        // ---- C: stable rank of every point = i - #(earlier, larger) + #(later, smaller).  A point with nothing larger
        // before it and nothing smaller after it keeps its source index.  The others (folds, jitter) are put on a
        // work list and counted by the whole CTA, neighbours in a fold going to neighbouring lanes.
        uint32_t dirty_bits = 0;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const float x = v[e];
            const float pe = (e == 0) ? em : lmax[e - 1];          // max of all earlier points
            const float se = (e == PER - 1) ? en : lmin[e + 1];    // min of all later points
            if (!((pe <= x) && (se >= x))) dirty_bits |= 1u << e;
        }
        if (dirty_bits) {
            int base = atomicAdd(&s_ndirty, __popc(dirty_bits));
#pragma unroll
            for (int e = 0; e < PER; ++e)
                if (dirty_bits & (1u << e)) dlist[base++] = (unsigned short)(i0 + e);
        }
        __syncthreads();
        const int ndirty = s_ndirty;
        for (int q = tid; q < ndirty; q += kPolyThreads) {
            const int i = dlist[q];
            const float x = px[i];
            int r = i;
            for (int j = i - 1; j >= 0 && pm[j] > x; --j) r -= (px[j] > x) ? 1 : 0;
            for (int j = i + 1; j < npts && sm[j] < x; ++j) r += (px[j] < x) ? 1 : 0;
            drank[i] = (unsigned short)r;
        }
        __syncthreads();   // pm / sm are dead from here on: sxd overwrites them
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int i = i0 + e;
            if (i < npts) {
                const int r = (dirty_bits & (1u << e)) ? (int)drank[i] : i;
                sxd[r] = (double)v[e];
                sidx[r] = (unsigned short)i;
            }
        }
    }
    __syncthreads();

    // ---- D: reach = prefix max over sorted segments of their end x1; bucket starts
    {
        float m = -INFINITY;
        float loc[PER], avv[PER], x1v[PER];
        int spv[PER];
        int bprev;
        {
            int kp = i0 - 1;
            if (kp < 0) bprev = -1;
            else if (kp >= npts) bprev = tw + 1;
            else bprev = min(max(__float2int_rd(px[sidx[kp]]) - t0 + 1, 0), tw + 1);
        }
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int k = i0 + e;
            float x1 = -INFINITY;
            avv[e] = 0.0f; spv[e] = 0;
            if (k < npts) {
                const int sp = (int)sidx[k];
                if (k < nsg) x1 = px[sp + 1];
                const float pv = px[sp];
                avv[e] = pv; spv[e] = sp;
                // bucket = floor(x) - t0 + 1 clamped to [0, tw + 1] (the float -> int conversion saturates)
                const int b = min(max(__float2int_rd(pv) - t0 + 1, 0), tw + 1);
                if (b > bprev) {
                    start[b] = k;
                    for (int q = bprev + 1; q < b; ++q) start[q] = k;   // empty buckets in between (holes)
                    bprev = b;
                }
                if (k == npts - 1) start[tw + 2] = npts;
            }
            x1v[e] = x1;
            m = fmaxf(m, x1);
            loc[e] = m;
        }
        float im = m;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            float t = __shfl_up_sync(0xffffffffu, im, o);
            if (lane >= o) im = fmaxf(im, t);
        }
        if (lane == 31) s_wa[wid] = im;
        float em = __shfl_up_sync(0xffffffffu, im, 1);
        if (lane == 0) em = -INFINITY;
        __syncthreads();
        for (int q = 0; q < wid; ++q) em = fmaxf(em, s_wa[q]);
#pragma unroll
        for (int q = 0; q < PER / 4; ++q)
            reinterpret_cast<float4*>(reach + i0)[q] = make_float4(fmaxf(loc[4 * q], em), fmaxf(loc[4 * q + 1], em),
                                                                   fmaxf(loc[4 * q + 2], em), fmaxf(loc[4 * q + 3], em));
        // ---- D2 (first half): an interval whose left point no EARLIER segment reaches past has exactly its own
        // segment as candidate (or none, if that one does not go forward).  Everything else goes to a work list.
        uint32_t hard = 0;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int k = i0 + e;
            if (k >= nsg) continue;
            const float rprev = (e == 0) ? em : fmaxf(loc[e - 1], em);   // reach[k - 1]
            if (rprev > avv[e]) {
                hard |= 1u << e;
            } else {
                const uint32_t code = (x1v[e] > avv[e]) ? 1u : 0u;
                cand[k] = (unsigned short)code;
            }
        }
        if (hard) {
            int base = atomicAdd(&s_ntwo, __popc(hard));
#pragma unroll
            for (int e = 0; e < PER; ++e)
                if (hard & (1u << e)) tlist[base++] = (unsigned short)(i0 + e);
        }
    }
    __syncthreads();

    // ---- D2 (second half): active set of the listed intervals (a, b) = (point k, point k+1).  For a centre strictly
    // inside, a segment j is active iff it starts at or before a (j <= k) and ends beyond a; nothing between a and b
    // is a point, so all of this is float32 comparisons.  Up to two active segments are encoded in cand[k].
    {
        const int nlist = s_ntwo;
        for (int q = tid; q < nlist; q += kPolyThreads) {
            const int k = tlist[q];
            const uint32_t me = sidx[k];
            const float av = px[me];
            int cnt = 0, o1 = 0, o2 = 0;
            for (int j = k; j >= 0 && reach[j] > av; --j) {
                const int sp = (int)sidx[j];
                if (px[sp + 1] > av) {
                    if (cnt == 0) o1 = k - j; else if (cnt == 1) o2 = k - j;
                    ++cnt;
                }
            }
            uint32_t code = (uint32_t)min(cnt, 3);
            if (o1 > 127 || o2 > 127) code = 3u;
            if (code == 2u) {
                // Interpolated closeness is linear in the centre, so if one candidate leads at both ends of the interval
                // by more than any rounding could matter, it leads at every centre inside: the interval becomes a
                // one-candidate interval.  The reference also requires 0 < ip < 1; ip > 0 always holds for an active
                // segment, and ip < 1 can only fail (float32 rounding of x1 - x0) for long segments that end at or just
                // beyond this interval's right point -- those stay two-candidate and are decided per visit in FP64.
                const float bv = px[sidx[k + 1]];
                const int spA = (int)sidx[k - o1], spB = (int)sidx[k - o2];
                const float ax0 = px[spA], ax1 = px[spA + 1], bx0 = px[spB], bx1 = px[spB + 1];
                const float aq0 = clo[pt_slot(spA, SHARP)], aq1 = clo[pt_slot(spA + 1, SHARP)];
                const float bq0 = clo[pt_slot(spB, SHARP)], bq1 = clo[pt_slot(spB + 1, SHARP)];
                const float ad = ax1 - ax0, bd = bx1 - bx0;
                // ip < 1 is guaranteed when the centre stays more than half a float32 ulp of (x1 - x0) below x1:
                // always for lengths < 2 (the centre is >= 1e-7 below the interval's end), else when x1 is far enough
                // beyond the interval
                const bool safe = (ad < 2.0f || (ax1 - bv) > ad * 1.2e-7f) && (bd < 2.0f || (bx1 - bv) > bd * 1.2e-7f);
                // closeness of both candidates at the two ends (float32 is plenty: the margin below is 1e-3)
                const float ta = __fdividef(av - ax0, ad), tb = __fdividef(bv - ax0, ad);
                const float ua = __fdividef(av - bx0, bd), ub = __fdividef(bv - bx0, bd);
                const float a_lo = aq0 + ta * (aq1 - aq0), a_hi = aq0 + tb * (aq1 - aq0);
                const float b_lo = bq0 + ua * (bq1 - bq0), b_hi = bq0 + ub * (bq1 - bq0);
                const float margin = 1e-3f + 1e-4f * fmaxf(fmaxf(aq0, aq1), fmaxf(bq0, bq1));
                if (safe && a_lo > b_lo + margin && a_hi > b_hi + margin) code = 1u;                     // first candidate
                else if (safe && b_lo > a_lo + margin && b_hi > a_hi + margin) { code = 1u; o1 = o2; }  // second candidate
            }
            cand[k] = (unsigned short)(code | ((uint32_t)(o1 & 127) << kOff1Shift) | ((uint32_t)(o2 & 127) << kOff2Shift));
        }
    }
    __syncthreads();

    // ---- E: sweep, one thread per output column
    PolyCtx ctx;
    ctx.px = px; ctx.sxd = sxd; ctx.sidx = sidx; ctx.reach = reach; ctx.clo = clo; ctx.start = start; ctx.row = c;
    ctx.t0 = t0;
    uint32_t* out = a.out[eye] + row_off + t0;
    bool give_up = false;
    constexpr bool shp = SHARP;
    // warps take 32-column blocks from a shared counter: blocks inside folds cost several times more than smooth ones
    const int nblk = (own + 31) >> 5, first = tw - own;   // the tile's own columns are the last `own` bucket columns
    for (;;) {
        int blk = 0;
        if (lane == 0) blk = atomicAdd(&s_next, 1);
        blk = __shfl_sync(0xffffffffu, blk, 0);
        if (blk >= nblk) break;
        const int col = first + (blk << 5) + lane;     // output column relative to the bucket origin t0
        if (col >= tw) continue;
        double c0 = 0.5, c1 = 0.5, c2 = 0.5;   // float32-valued accumulators kept in float64 registers
        const int k0 = start[col + 1] - 1, k1 = start[col + 2] - 1;
        const double cold = u8_to_f64((uint32_t)(col + t0)), col1d = cold + 1.0;
        double pa = sxd[k0];
        for (int k = k0; k <= k1; ++k) {
            const double pb = sxd[k + 1];
            const double from = ((pa > cold) ? pa : cold) + kEps;     // no NaNs here: plain compare-select
            const double to = ((pb < col1d) ? pb : col1d) - kEps;
            const double sig = to - from;
            const double ctr = from + 0.5 * sig;
            const uint32_t inf = cand[k];
            const uint32_t code = inf & 3u;
            int sp;
            bool resolved = false, off_is_k = false;
            if (code == 1u && sig > 0.0) {
                const int off1 = (int)((inf >> kOff1Shift) & 127u);
                sp = (int)sidx[k - off1];
                off_is_k = (off1 == 0);   // the segment starts at this interval's left point: x0 = pa
                resolved = true;
            }
            if (!resolved) {
                off_is_k = false;
                sp = general_visit(ctx, col, k, ctr);
                if (sp == -2) { give_up = true; sp = -1; }
                if (sp < 0) { pa = pb; continue; }
            }
            const int sl = pt_slot(sp, shp), sr = pt_slot(sp + 1, shp);
            const int cl = slot_col(sl, w), cr = slot_col(sr, w);
            const uint32_t pl = simg[cl];
            double v0 = u8_to_f64(pl & 255u), v1 = u8_to_f64((pl >> 8) & 255u), v2 = u8_to_f64((pl >> 16) & 255u);
            if (cl != cr) {
                // ip = (ctr - x0) / (x1 - x0) with the reference's float32 subtraction in the denominator
                const double x0 = off_is_k ? pa : (double)px[sp];
                const double x1 = (double)px[sp + 1];
                const double den = round24_even(x1 - x0);
                const double ip = (ctr - x0) / den;
                const uint32_t pr = simg[cr];
                const double om = 1.0 - ip;
                double t0 = v0 * om, t1 = u8_to_f64(pr & 255u) * ip;
                v0 = t0 + t1;
                t0 = v1 * om; t1 = u8_to_f64((pr >> 8) & 255u) * ip;
                v1 = t0 + t1;
                t0 = v2 * om; t1 = u8_to_f64((pr >> 16) & 255u) * ip;
                v2 = t0 + t1;
            }
            c0 = round24_fp(c0 + v0 * sig);
            c1 = round24_fp(c1 + v1 * sig);
            c2 = round24_fp(c2 + v2 * sig);
            pa = pb;
        }
        out[col] = pack_rgbx(__double2int_rz(c0), __double2int_rz(c1), __double2int_rz(c2));
    }
    if (give_up) s_flag = 1;
    __syncthreads();
    const int flagged = s_flag;
    if (tid == 0 && tile_w > 0) {
        // a tile cannot replay the whole row: flag it for k_polylines_exact, which runs right after this kernel
        if (flagged) atomicOr(&row_flags[((int64_t)frame * 2 + eye) * a.h + y], 1);
    } else if (tid == 0) {
        if (row_flags) row_flags[((int64_t)frame * 2 + eye) * a.h + y] = flagged;
        if (flagged) {
            // the list replay did not fit its budget somewhere in this row: redo the row sequentially
            // (reach[] is dead now and serves as the active list)
            bool ovf = sequential_row(ctx, simg, out, reinterpret_cast<unsigned short*>(reach), 2 * NP);
            if (ovf) atomicOr(status, 1);
        }
    }
}

template <int PER>
static size_t fast_smem_per(int w) {
    return (size_t)kPolyThreads * PER * (4 + 8 + 4 + 4 + 2) + (size_t)(w + 4) * 4 + (size_t)(w + 4) * 4 + (size_t)w * 4;
}

static size_t exact_smem(int w, int sharp, int act_cap) {
    size_t npts = (size_t)(sharp ? 2 * w : w) + 2;
    size_t np2 = npts + (npts & 1);
    return npts * 4 + (size_t)w * 4 + (size_t)(w + 4) * 4 + np2 * 2 * 3 + (size_t)act_cap * 2;
}
size_t polylines_scratch_bytes(int n, int h) { return ((size_t)n * 2 * h + 16) * sizeof(int); }

template <int PER>
static cudaError_t launch_fast(const WarpArgs& a, int sharp, int* flags, int* status, int wmax, int tile_w, int tile_ext,
                               int ntiles, double reach0, double reach1, cudaStream_t s) {
    const size_t fs = fast_smem_per<PER>(wmax);
    const void* fn = tile_w > 0 ? (sharp ? (const void*)k_polylines<PER, true, true> : (const void*)k_polylines<PER, false, true>)
                                : (sharp ? (const void*)k_polylines<PER, true, false> : (const void*)k_polylines<PER, false, false>);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs);
    if (e != cudaSuccess) return e;
    prof_begin(K_POLY_FAST, s);
    const dim3 grid(a.h, a.n, 2 * ntiles);
#define CS_POLY_LAUNCH(SH, TL) k_polylines<PER, SH, TL><<<grid, kPolyThreads, fs, s>>>(a, flags, status, tile_w, tile_ext, reach0, reach1)
    if (tile_w > 0) { if (sharp) CS_POLY_LAUNCH(true, true); else CS_POLY_LAUNCH(false, true); }
    else { if (sharp) CS_POLY_LAUNCH(true, false); else CS_POLY_LAUNCH(false, false); }
#undef CS_POLY_LAUNCH
    prof_end(K_POLY_FAST, s);
    count_launch();
    return cudaGetLastError();
}

static cudaError_t launch_exact(const WarpArgs& a, int sharp, const int* flags, int* status, cudaStream_t s) {
    const size_t kMaxSmem = 227 * 1024;
    double dmax = fmax(fabs(a.eye[0].div_px), fabs(a.eye[1].div_px));
    long long cap_ref = 5ll * (long long)dmax + 25;    // the reference's own list capacity, SIG:1947
    size_t base = exact_smem(a.w, sharp, 0);
    if (base + 64 > kMaxSmem) return cudaErrorInvalidValue;
    long long cap_fit = (long long)((kMaxSmem - base) / 2);
    int act_cap = (int)(cap_ref < cap_fit ? cap_ref : cap_fit);
    size_t es = exact_smem(a.w, sharp, act_cap);
    if (es > 48 * 1024) cudaFuncSetAttribute(k_polylines_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)es);
    prof_begin(K_POLY_EXACT, s);
    k_polylines_exact<<<dim3(a.h, a.n, 2), kPolyThreads, es, s>>>(a, sharp, act_cap, flags, status);
    prof_end(K_POLY_EXACT, s);
    count_launch();
    return cudaGetLastError();
}

// scratch: [n*2*h] row flags + [1] status word.
// flags bit 0 = replay every row with the sequential kernel, bit 2 = force 64-column tiles (tests).
cudaError_t launch_polylines(const WarpArgs& a, cudaStream_t s) {
    const int sharp = a.fill == CS_FILL_POLYLINES_SHARP;
    const int w = a.w;
    if (a.scratch_bytes < polylines_scratch_bytes(a.n, a.h)) return cudaErrorInvalidValue;
    const int npts = (sharp ? 2 * w : w) + 2;
    if (npts > 65535) return cudaErrorInvalidValue;
    int* flags = reinterpret_cast<int*>(a.scratch);
    int* status = flags + (size_t)a.n * 2 * a.h;
    const size_t kMaxSmem = 227 * 1024;
    const bool force_exact = (a.flags & 1) != 0, force_tiles = (a.flags & 4) != 0;
    if (!force_exact) {
        // whole row per CTA while two CTAs still fit on an SM
        if (!force_tiles) {
            if (npts <= kPolyThreads * 4 && fast_smem_per<4>(w) <= kMaxSmem) return launch_fast<4>(a, sharp, flags, status, w, 0, 0, 1, 0.0, 0.0, s);
            if (npts <= kPolyThreads * 8 && fast_smem_per<8>(w) <= kMaxSmem) return launch_fast<8>(a, sharp, flags, status, w, 0, 0, 1, 0.0, 0.0, s);
        }
        // wider rows: tiles of output columns, each with the source window that can reach it.  |shift| is bounded by
        // |div_px| * max(conv, 1 - conv)^expo because the normalised depth lies in [-conv, 1 - conv].
        const double span = pow(fmax((double)a.conv, 1.0 - (double)a.conv), a.expo);
        const double reach0 = fabs(a.eye[0].div_px) * span + 1.0, reach1 = fabs(a.eye[1].div_px) * span + 1.0;
        const int rmax = (int)ceil(fmax(reach0, reach1));
        const int tile_ext = 2 * rmax + 8;                      // a fold is at most 2 * reach wide
        const int guard = 2 * (rmax + 5) + 2 + tile_ext;
        const int cap_cols = force_tiles ? (kPolyThreads * 4 - 2) / (sharp ? 2 : 1) : (kPolyThreads * 8 - 2) / (sharp ? 2 : 1);
        int tile_w = force_tiles ? 64 : ((cap_cols - guard) / 32) * 32;
        if (tile_w >= 64 && tile_w + guard <= cap_cols) {
            const int wmax = (tile_w + guard < w) ? tile_w + guard : w;
            const int ntiles = (w + tile_w - 1) / tile_w;
            cudaError_t e = cudaMemsetAsync(flags, 0, (size_t)a.n * 2 * a.h * sizeof(int), s);
            if (e != cudaSuccess) return e;
            e = force_tiles ? launch_fast<4>(a, sharp, flags, status, wmax, tile_w, tile_ext, ntiles, reach0, reach1, s)
                            : launch_fast<8>(a, sharp, flags, status, wmax, tile_w, tile_ext, ntiles, reach0, reach1, s);
            if (e != cudaSuccess) return e;
            return launch_exact(a, sharp, flags, status, s);   // rows a tile could not finish (usually none)
        }
    }
    // the test hook, or a disparity range so large that no useful tile fits: sequential kernel for every row
    return launch_exact(a, sharp, nullptr, status, s);
}

}  // namespace cs

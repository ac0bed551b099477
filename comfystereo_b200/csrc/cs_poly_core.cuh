// cs_poly_core.cuh -- per-interval and per-column arithmetic of the polylines sweep (SIG:1948-1991), shared by the
// CUDA kernel (cs_polylines.cu) and by the host model that checks its certification logic on the CPU
// (tools/poly_model.cpp).  Everything here works on the sorted point tables of ONE row (or one tile of a row).
//
// Two evaluations of an output column:
//   exact_column   the reference's arithmetic operation by operation (float64 centres, float32 accumulator rounded after
//                  every sub-interval, list-order dependent choices by replay -- quirk Q7).  Always right, slow.
//   fast_column    float32 only.  It does not try to reproduce the reference's roundings; it proves instead that they
//                  cannot matter: the result R of the reference lies within E of the value F computed here (E bounds
//                  every rounding on both sides, derivation at fast_column), so trunc(R) == trunc(F) unless an integer
//                  lies within E of F.  Columns that cannot be certified (about 1 in 1000 on natural input), or whose
//                  intervals have no pre-resolved winner, return false and are redone by exact_column.
//
// The arithmetic contract of the library (-fmad=false / -ffp-contract=off) holds here too: a * b + c rounds twice unless
// it is written fmaf().
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define CS_HD __host__ __device__ __forceinline__
#define CS_HDN __host__ __device__ __noinline__
#else
#define CS_HD inline
#define CS_HDN
#endif

namespace cs {
namespace poly {

constexpr double kEps = 1e-7;

// Tables of one row / tile, all in sorted order unless noted (k = sorted rank, i = source point index).
//   point i: 0 = left sentinel (-W), 1 .. npts-2 the row's points in source order, npts-1 = right sentinel (2W)
//   segment i connects point i to point i + 1 (nsg = npts - 1 of them)
struct Tab {
    const float* X;          // [npts]   x of source point i
    const float* SX;         // [npts]   x of sorted point k
    const uint32_t* ER;      // [npts]   low 16: rank of the END point of the segment that starts at sorted point k
                             //          high 16: inclusive prefix maximum of that (how far any segment up to k reaches)
    const uint16_t* SID;     // [npts]   source index of sorted point k
    const uint16_t* WSP;     // [npts]   interval (k, k+1): source index of the segment that wins at every centre strictly
                             //          inside, or kUnresolved | own index when that is not known (none, several, ties)
    const float* Q;          // [w + 2]  closeness, padded: Q[pt_slot(i)] belongs to source point i (both sentinels 0)
    const uint32_t* IMGP;    // [npts]   RGBX of source point i's pixel (both sentinels repeat the edge pixels): same index as X,
                             //          so the sweep reaches x and colour of a segment's two ends from one address
    const uint16_t* START;   // [tw + 3] first sorted rank of bucket b = floor(x) - t0 + 1 (0 = left of t0, tw + 1 = right)
    int w, npts, nsg;        // source window width, points, segments
    int t0;                  // absolute output column of bucket 1
};
constexpr uint32_t kUnresolved = 0x8000u;

CS_HD int imin_(int a, int b) { return a < b ? a : b; }
CS_HD int imax_(int a, int b) { return a > b ? a : b; }
template <bool SHARP> CS_HD int pt_slot(int i) { return SHARP ? ((i + 1) >> 1) : i; }
CS_HD int slot_col(int slot, int w) { return imin_(imax_(slot - 1, 0), w - 1); }
CS_HD uint32_t pack3(int r, int g, int b) { return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16); }

// ------------------------------------------------------------------ bit-level helpers of the exact path
CS_HD double hilo2double(uint32_t hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double((int)hi, (int)lo);
#else
    uint64_t b = ((uint64_t)hi << 32) | lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
CS_HD void double2hilo(double x, uint32_t* hi, uint32_t* lo) {
#ifdef __CUDA_ARCH__
    *hi = (uint32_t)__double2hiint(x);
    *lo = (uint32_t)__double2loint(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    *hi = (uint32_t)(b >> 32);
    *lo = (uint32_t)b;
#endif
}
// Round-to-nearest of a float64 to 24 significant bits, kept as float64: equals (double)(float)x for normal results.
// Veltkamp split with 2^29 + 1 (three FP64 operations, no F2F conversion).  Differs from ties-to-even only on exact ties,
// which an accumulated colour sum hits with probability 2^-29 per operation.
CS_HD double round24_fp(double x) {
    const double g = x * 536870913.0;
    const double d = x - g;
    return g + d;
}
// The same rounding with exact ties-to-even in integer arithmetic, for the float32 subtraction x1 - x0 of two float32
// coordinates (a short exact binary number: exact ties are common there).
CS_HD double round24_even(double x) {
    uint32_t hi, lo;
    double2hilo(x, &hi, &lo);
    const uint32_t nlo = lo + 0x0FFFFFFFu + ((lo >> 29) & 1u);
    hi += (nlo < lo) ? 1u : 0u;
    return hilo2double(hi, nlo & 0xE0000000u);
}
CS_HD double u8_to_f64(uint32_t v) { return hilo2double(0x43300000u, v) - 4503599627370496.0; }   // 2^52 + v - 2^52

CS_HD double visit_ctr(const Tab& c, int col, int k, double* sig_out) {
    const double pa = (double)c.SX[k], pb = (double)c.SX[k + 1];
    const double from = fmax((double)(col + c.t0), pa) + kEps;
    const double to = fmin((double)(col + c.t0 + 1), pb) - kEps;
    const double sig = to - from;
    *sig_out = sig;
    return from + 0.5 * sig;
}
// furthest end (as x) of any segment starting at sorted points 0 .. j
CS_HD double reach_x(const Tab& c, int j) { return (double)c.SX[c.ER[j] >> 16]; }

// number of active segments at ctr (interval k); *which = sorted index of the last one found
CS_HD int active_count(const Tab& c, int k, double ctr, int* which) {
    int n = 0;
    for (int j = k; j >= 0 && !(reach_x(c, j) < ctr); --j) {
        const int sp = (int)c.SID[j];
        if (!((double)c.SX[j] < ctr) || ((double)c.X[sp + 1] < ctr)) continue;
        ++n;
        *which = j;
    }
    return n;
}

// The reference's selection when the result depends on the ORDER of its active list (quirk Q7): find the nearest
// earlier visit with exactly one active segment (there the list is [that segment], whatever happened before), replay
// the append / swap-remove list from there to the target visit, then choose as the reference does.  Returns the source
// point index of the chosen segment, or -1 when the replay does not fit its budget or its history starts left of the
// tile (the whole row is then redone sequentially by k_polylines_exact).
template <bool SHARP>
CS_HDN int replay_choice(const Tab& c, int col, int k) {
    constexpr int kCap = 64, kBudget = 6000;
    unsigned short lst[kCap];
    const int nsg = c.nsg;
    int rc = col, rk = k, sgp = 0, steps = 0;
    bool from_row_start = false;
    while (true) {
        if (rk > (int)c.START[rc + 1] - 1) --rk;
        else if (rc > 0) { --rc; rk = (int)c.START[rc + 2] - 1; }
        else if (c.t0 == 0) { from_row_start = true; break; }
        else return -1;
        if (++steps > kBudget) return -1;
        double sig;
        const double ctr = visit_ctr(c, rc, rk, &sig);
        int which = 0;
        if (active_count(c, rk, ctr, &which) == 1) { sgp = which; break; }
    }
    int n = 0;
    if (from_row_start) { rc = 0; rk = (int)c.START[1] - 1; sgp = 0; }
    while (true) {
        double sig;
        const double ctr = visit_ctr(c, rc, rk, &sig);
        while (sgp < nsg && (double)c.SX[sgp] < ctr) {
            if (n >= kCap) return -1;
            lst[n++] = c.SID[sgp];
            ++sgp;
        }
        for (int i = 0; i < n;) {
            if ((double)c.X[lst[i] + 1] < ctr) { lst[i] = lst[n - 1]; --n; }
            else ++i;
        }
        if (rc == col && rk == k) {
            if (n == 0) return -1;
            int best = 0;
            if (n != 1) {
                double bestc = -kEps;
                for (int i = 0; i < n; ++i) {
                    const int sp = lst[i];
                    const float x0 = c.X[sp], x1 = c.X[sp + 1];
                    const float den = x1 - x0;
                    const double ip = (ctr - (double)x0) / (double)den;
                    const double t0 = (1.0 - ip) * (double)c.Q[pt_slot<SHARP>(sp)], t1 = ip * (double)c.Q[pt_slot<SHARP>(sp + 1)];
                    const double cl = t0 + t1;
                    if (bestc < cl && 0.0 < ip && ip < 1.0) { bestc = cl; best = i; }
                }
            }
            return lst[best];
        }
        if (rk < (int)c.START[rc + 2] - 1) ++rk;
        else { ++rc; rk = (int)c.START[rc + 1] - 1; }
    }
}

// Any visit that is not pre-resolved: builds the active set at ctr, selects in FP64 like the reference, and resolves
// order-dependent choices by replay.  Returns the source point index of the chosen segment, -1: nothing active,
// -2: give up (the row is flagged for the sequential kernel).
template <bool SHARP>
CS_HDN int general_visit(const Tab& c, int col, int k, double ctr) {
    int nact = 0, best = -1, only = -1, nbest = 0;
    double bestc = -kEps;
    for (int j = k; j >= 0 && !(reach_x(c, j) < ctr); --j) {
        const int sp = (int)c.SID[j];
        const float x0 = c.SX[j], x1 = c.X[sp + 1];
        if (!((double)x0 < ctr) || ((double)x1 < ctr)) continue;
        ++nact;
        only = sp;
        const float den = x1 - x0;
        const double ip = (ctr - (double)x0) / (double)den;
        const double t0 = (1.0 - ip) * (double)c.Q[pt_slot<SHARP>(sp)], t1 = ip * (double)c.Q[pt_slot<SHARP>(sp + 1)];
        const double cl = t0 + t1;
        if (0.0 < ip && ip < 1.0) {
            if (bestc < cl) { bestc = cl; best = sp; nbest = 1; }
            else if (bestc == cl) ++nbest;
        }
    }
    if (nact == 0) return -1;
    if (nact == 1) return only;
    if (best >= 0 && nbest == 1) return best;
    const int r = replay_choice<SHARP>(c, col, k);
    return r < 0 ? -2 : r;
}

// One output column exactly as the reference computes it (SIG:1955-1991).  `col` is relative to the bucket origin t0.
// Returns the RGBX pixel; bit 31 is set when a list replay gave up (the pixel is then not final).
constexpr uint32_t kGaveUp = 0x80000000u;
template <bool SHARP>
CS_HDN uint32_t exact_column(const Tab& c, int col) {
    double c0 = 0.5, c1 = 0.5, c2 = 0.5;   // float32-valued accumulators kept in float64 registers
    const int k0 = (int)c.START[col + 1] - 1, k1 = (int)c.START[col + 2] - 1;
    const double cold = (double)(col + c.t0), col1d = cold + 1.0;
    double pa = (double)c.SX[k0];
    bool ok = true;
    for (int k = k0; k <= k1; ++k) {
        const double pb = (double)c.SX[k + 1];
        const double from = ((pa > cold) ? pa : cold) + kEps;
        const double to = ((pb < col1d) ? pb : col1d) - kEps;
        const double sig = to - from;
        const double ctr = from + 0.5 * sig;
        const uint32_t inf = c.WSP[k];
        int sp;
        if (!(inf & kUnresolved) && sig > 0.0) {
            sp = (int)inf;
        } else {
            sp = general_visit<SHARP>(c, col, k, ctr);
            if (sp == -2) { ok = false; sp = -1; }
            if (sp < 0) { pa = pb; continue; }
        }
        const int cl = slot_col(pt_slot<SHARP>(sp), c.w), cr = slot_col(pt_slot<SHARP>(sp + 1), c.w);
        const uint32_t pl = c.IMGP[sp];
        double v0 = u8_to_f64(pl & 255u), v1 = u8_to_f64((pl >> 8) & 255u), v2 = u8_to_f64((pl >> 16) & 255u);
        if (cl != cr) {
            // ip = (ctr - x0) / (x1 - x0) with the reference's float32 subtraction in the denominator
            const double x0 = (double)c.X[sp];
            const double x1 = (double)c.X[sp + 1];
            const double den = round24_even(x1 - x0);
            const double ip = (ctr - x0) / den;
            const uint32_t pr = c.IMGP[sp + 1];
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
    return pack3((int)c0, (int)c1, (int)c2) | (ok ? 0u : kGaveUp);
}

#ifdef __CUDACC__
// exact_column by a whole warp: all 32 lanes call it with the same column.  The sub-intervals of a column are independent
// up to the colour accumulation (centre, winner, interpolation parameter, the three products colour * significance), so
// each lane prepares one sub-interval; only the float32-rounded running sums are then formed in order.  Same operations
// in the same order as exact_column, a third of its latency -- the CTA waits for these columns with everything else done.
__device__ __forceinline__ double shfl_f64(double v, int src) {
    const int lo = __shfl_sync(0xffffffffu, __double2loint(v), src), hi = __shfl_sync(0xffffffffu, __double2hiint(v), src);
    return __hiloint2double(hi, lo);
}
template <bool SHARP>
__device__ __noinline__ uint32_t exact_column_warp(const Tab& c, int col) {
    const int lane = threadIdx.x & 31;
    double c0 = 0.5, c1 = 0.5, c2 = 0.5;
    const int k0 = (int)c.START[col + 1] - 1, k1 = (int)c.START[col + 2] - 1;
    const double cold = (double)(col + c.t0), col1d = cold + 1.0;
    bool ok = true;
    for (int kb = k0; kb <= k1; kb += 32) {
        const int k = kb + lane;
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
        bool have = false, bad = false;
        if (k <= k1) {
            const double pa = (double)c.SX[k], pb = (double)c.SX[k + 1];
            const double from = ((pa > cold) ? pa : cold) + kEps;
            const double to = ((pb < col1d) ? pb : col1d) - kEps;
            const double sig = to - from;
            const double ctr = from + 0.5 * sig;
            const uint32_t inf = c.WSP[k];
            int sp;
            if (!(inf & kUnresolved) && sig > 0.0) {
                sp = (int)inf;
            } else {
                sp = general_visit<SHARP>(c, col, k, ctr);
                if (sp == -2) { bad = true; sp = -1; }
            }
            if (sp >= 0) {
                const int cl = slot_col(pt_slot<SHARP>(sp), c.w), cr = slot_col(pt_slot<SHARP>(sp + 1), c.w);
                const uint32_t pl = c.IMGP[sp];
                double v0 = u8_to_f64(pl & 255u), v1 = u8_to_f64((pl >> 8) & 255u), v2 = u8_to_f64((pl >> 16) & 255u);
                if (cl != cr) {
                    const double x0 = (double)c.X[sp];
                    const double x1 = (double)c.X[sp + 1];
                    const double den = round24_even(x1 - x0);
                    const double ip = (ctr - x0) / den;
                    const uint32_t pr = c.IMGP[sp + 1];
                    const double om = 1.0 - ip;
                    double a = v0 * om, b = u8_to_f64(pr & 255u) * ip;
                    v0 = a + b;
                    a = v1 * om; b = u8_to_f64((pr >> 8) & 255u) * ip;
                    v1 = a + b;
                    a = v2 * om; b = u8_to_f64((pr >> 16) & 255u) * ip;
                    v2 = a + b;
                }
                t0 = v0 * sig; t1 = v1 * sig; t2 = v2 * sig;
                have = true;
            }
        }
        if (__any_sync(0xffffffffu, bad)) ok = false;
        const uint32_t hv = __ballot_sync(0xffffffffu, have);
        const int nv = imin_(32, k1 - kb + 1);
        for (int j = 0; j < nv; ++j) {
            const double u0 = shfl_f64(t0, j), u1 = shfl_f64(t1, j), u2 = shfl_f64(t2, j);
            if (hv & (1u << j)) {
                c0 = round24_fp(c0 + u0);
                c1 = round24_fp(c1 + u1);
                c2 = round24_fp(c2 + u2);
            }
        }
    }
    return pack3((int)c0, (int)c1, (int)c2) | (ok ? 0u : kGaveUp);
}

// Up to four uncertified columns at once, eight lanes each (a column has 2..6 sub-intervals almost always): the FP64 work of
// a row's few uncertified columns is one pass of one warp instead of one pass per column.  `cols[g]` is the column of lane
// group g (lanes 8g..8g+7), g < ncols.  Returns the column's RGBX | kGaveUp to every lane of its group; groups whose
// column has more than eight sub-intervals are reported in *big (bit g) and left to exact_column_warp.
template <bool SHARP>
__device__ __noinline__ uint32_t exact_columns_quad(const Tab& c, const uint16_t* cols, int ncols, uint32_t* big) {
    const int lane = threadIdx.x & 31, g = lane >> 3, j = lane & 7;
    const int col = (g < ncols) ? (int)cols[g] : -1;
    int k0 = 0, cnt = 0;
    if (col >= 0) { k0 = (int)c.START[col + 1] - 1; cnt = (int)c.START[col + 2] - 1 - k0 + 1; }
    const bool is_big = cnt > 8;
    const uint32_t bigbits = __ballot_sync(0xffffffffu, is_big);
    *big = ((bigbits & 0x1u) ? 1u : 0u) | ((bigbits & 0x100u) ? 2u : 0u) | ((bigbits & 0x10000u) ? 4u : 0u) | ((bigbits & 0x1000000u) ? 8u : 0u);
    if (is_big) cnt = 0;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    bool have = false, bad = false;
    if (j < cnt) {
        const int k = k0 + j;
        const double cold = (double)(col + c.t0), col1d = cold + 1.0;
        const double pa = (double)c.SX[k], pb = (double)c.SX[k + 1];
        const double from = ((pa > cold) ? pa : cold) + kEps;
        const double to = ((pb < col1d) ? pb : col1d) - kEps;
        const double sig = to - from;
        const double ctr = from + 0.5 * sig;
        const uint32_t inf = c.WSP[k];
        int sp;
        if (!(inf & kUnresolved) && sig > 0.0) {
            sp = (int)inf;
        } else {
            sp = general_visit<SHARP>(c, col, k, ctr);
            if (sp == -2) { bad = true; sp = -1; }
        }
        if (sp >= 0) {
            const int cl = slot_col(pt_slot<SHARP>(sp), c.w), cr = slot_col(pt_slot<SHARP>(sp + 1), c.w);
            const uint32_t pl = c.IMGP[sp];
            double v0 = u8_to_f64(pl & 255u), v1 = u8_to_f64((pl >> 8) & 255u), v2 = u8_to_f64((pl >> 16) & 255u);
            if (cl != cr) {
                const double x0 = (double)c.X[sp];
                const double x1 = (double)c.X[sp + 1];
                const double den = round24_even(x1 - x0);
                const double ip = (ctr - x0) / den;
                const uint32_t pr = c.IMGP[sp + 1];
                const double om = 1.0 - ip;
                double a = v0 * om, b = u8_to_f64(pr & 255u) * ip;
                v0 = a + b;
                a = v1 * om; b = u8_to_f64((pr >> 8) & 255u) * ip;
                v1 = a + b;
                a = v2 * om; b = u8_to_f64((pr >> 16) & 255u) * ip;
                v2 = a + b;
            }
            t0 = v0 * sig; t1 = v1 * sig; t2 = v2 * sig;
            have = true;
        }
    }
    const uint32_t badbits = __ballot_sync(0xffffffffu, bad), hv = __ballot_sync(0xffffffffu, have);
    const int maxcnt = __reduce_max_sync(0xffffffffu, cnt);
    double c0 = 0.5, c1 = 0.5, c2 = 0.5;
    for (int q = 0; q < maxcnt; ++q) {      // the float32-rounded running sums, in sub-interval order, per group
        const int src = (lane & ~7) + q;
        const double u0 = shfl_f64(t0, src), u1 = shfl_f64(t1, src), u2 = shfl_f64(t2, src);
        if (q < cnt && ((hv >> src) & 1u)) {
            c0 = round24_fp(c0 + u0);
            c1 = round24_fp(c1 + u1);
            c2 = round24_fp(c2 + u2);
        }
    }
    const bool gbad = ((badbits >> (lane & ~7)) & 0xFFu) != 0u;
    return pack3((int)c0, (int)c1, (int)c2) | (gbad ? kGaveUp : 0u);
}
#endif

// ------------------------------------------------------------------ float32 path
// 2^23 + byte `ch` of p, as a float: one PRMT on the device
CS_HD float u8f_biased(uint32_t p, int ch) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__byte_perm(p, 0x4B000000u, 0x7440u | (uint32_t)ch));
#else
    return 8388608.0f + (float)((p >> (8 * ch)) & 255u);
#endif
}
CS_HD float u8f(uint32_t p, int ch) {
#ifdef __CUDA_ARCH__
    // 0x4B0000vv is 2^23 + v: one PRMT and one FADD, no I2F
    return __uint_as_float(__byte_perm(p, 0x4B000000u, 0x7440u | (uint32_t)ch)) - 8388608.0f;
#else
    return (float)((p >> (8 * ch)) & 255u);
#endif
}

// Error budget of fast_column against the reference's float32 accumulator, per channel, colour units (0..255):
//   per sub-interval   2^-17   rounding of this accumulator (values < 256: half an ulp)
//                      2^-17   rounding of the reference's accumulator
//                      2^-17   significance: d = to' - from' is exact (both inside one pixel), d - 2e-7f rounds once
//                              (<= 2^-25), times a colour <= 255
//                      2^-17   the interpolated colour (one fmaf, values < 256), times a significance <= 1
//   per column         255 * 3.0e-7 = 7.7e-5   interpolation parameter: two float32 roundings of the numerator (the
//                              centre itself is from' + d / 2 exactly: the two epsilons cancel; 2 * 2^-24), the approximate
//                              reciprocal (rcp.approx: 2^-23) and the product (2^-24); its weight is the significance,
//                              which sums to <= 1 per column
// kErrVisit = 4 * 2^-17 and kErrColumn rounded up.
constexpr float kErrVisit = 3.2e-5f, kErrColumn = 9.0e-5f;

CS_HD float fast_rcp(float a) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return 1.0f / a;
#endif
}

// `col` relative to t0.  Returns false when the column has to be redone by exact_column.  Branch-free per sub-interval:
// every winner is interpolated (a segment between two points of the same source pixel interpolates a colour with itself).
template <bool SHARP>
CS_HD bool fast_column(const Tab& c, int col, uint32_t* out_px) {
    const int k0 = (int)c.START[col + 1] - 1, k1 = (int)c.START[col + 2] - 1;
    const float cf = (float)(col + c.t0), cf1 = cf + 1.0f;
    float a0 = 0.5f, a1 = 0.5f, a2 = 0.5f;
    float pa = c.SX[k0], dmin = 1.0f;
    uint32_t bad = 0;
    // (not unrolled: the trip count differs from lane to lane, and an unrolled body plus a remainder loop makes the warp
    // execute both for the longest lane; fetching sub-interval k + 1's tables while k is accumulated -- software
    // pipelining by hand -- costs registers the 64-register budget does not have: measured 3 % slower)
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int k = k0; k <= k1; ++k) {
        const float pb = c.SX[k + 1];
        const uint32_t inf = c.WSP[k];
        const float fromp = fmaxf(pa, cf), top = fminf(pb, cf1);
        const float d = top - fromp;
        // no pre-resolved winner (bit 15), or -- checked once, on the minimum -- an interval too short to have a centre
        // strictly inside it
        bad |= inf;
        dmin = fminf(dmin, d);
        const float sig = d - 2e-7f;
        const int sp = (int)(inf & (kUnresolved - 1u));
        const float x0 = c.X[sp], x1 = c.X[sp + 1];
        const uint32_t pl = c.IMGP[sp], pr = c.IMGP[sp + 1];
        const float den = x1 - x0;                       // the reference's float32 subtraction, bit for bit
        const float num = fmaf(0.5f, d, fromp - x0);     // centre - x0
        const float ip = num * fast_rcp(den);
        // (2^23 + r) - (2^23 + l) is r - l exactly: the right pixel's bias never has to be removed
        const float L0 = u8f_biased(pl, 0), L1 = u8f_biased(pl, 1), L2 = u8f_biased(pl, 2);
        const float l0 = L0 - 8388608.0f, l1 = L1 - 8388608.0f, l2 = L2 - 8388608.0f;
        const float v0 = fmaf(ip, u8f_biased(pr, 0) - L0, l0);
        const float v1 = fmaf(ip, u8f_biased(pr, 1) - L1, l1);
        const float v2 = fmaf(ip, u8f_biased(pr, 2) - L2, l2);
        a0 = fmaf(v0, sig, a0);
        a1 = fmaf(v1, sig, a1);
        a2 = fmaf(v2, sig, a2);
        pa = pb;
    }
    const float E = fmaf((float)(k1 - k0 + 1), kErrVisit, kErrColumn);
    const float f0 = floorf(a0), f1 = floorf(a1), f2 = floorf(a2);
    const float lo = fminf(fminf(a0 - f0, a1 - f1), a2 - f2), hi = fmaxf(fmaxf(a0 - f0, a1 - f1), a2 - f2);
    *out_px = pack3((int)f0, (int)f1, (int)f2);
    // (NaN sums -- a degenerate interval divided by zero -- fail the comparison and go to the exact path)
    return !(bad & kUnresolved) && dmin >= 1e-6f && lo >= E && hi <= 1.0f - E;
}

// ------------------------------------------------------------------ interval classification
// Candidates of interval (k, k+1) whose left point an earlier segment reaches past: every segment j <= k (sorted) whose
// end point ranks above k is active at every centre strictly inside the interval (rank comparisons: points between the
// two do not exist).  One candidate wins outright; two are reduced to one when one of them leads at both ends of the
// interval by more than any rounding could matter.  Returns the WSP entry.
template <bool SHARP>
CS_HD uint32_t classify_interval(const Tab& c, int k) {
    constexpr int kMaxCand = 4;
    // Interpolated closeness is linear in the centre, so a candidate that leads every other one at both ends of the
    // interval leads at every centre inside.  The reference also requires 0 < ip < 1; ip > 0 always holds for an active
    // segment, and ip < 1 can only fail (float32 rounding of x1 - x0) for long segments that end at or just beyond this
    // interval's right point -- those leave the interval unresolved, to be decided per visit in FP64.
    // One pass over the candidates as the walk finds them, with running statistics instead of a table: the leader at the
    // left end (largest closeness there), the runner-up's value, and the two largest values at the right end.
    const float av = c.SX[k], bv = c.SX[k + 1];
    int cnt = 0, obest = -1, ohi1 = -1;
    float lo1 = -1.0f, lo2 = -1.0f, hib = -1.0f, hi1 = -1.0f, hi2 = -1.0f;     // closeness is >= 0: -1 is "none"
    uint32_t spb = 0;
    bool safe = true;
    float qmax = 0.0f;
    for (int j = k; j >= 0; --j) {
        const uint32_t er = c.ER[j];
        if ((int)(er >> 16) <= k) break;            // nothing at or before j reaches past point k
        if ((int)(er & 0xFFFFu) > k) {
            const int sp = (int)c.SID[j];
            const float q0 = c.Q[pt_slot<SHARP>(sp)], q1 = c.Q[pt_slot<SHARP>(sp + 1)];
            const float x0 = c.SX[j], x1 = c.X[sp + 1];
            const float d = x1 - x0;
            safe = safe && d > 0.0f && (d < 2.0f || (x1 - bv) > d * 1.2e-7f);
            float l = q0, h = q0;
            if (q0 != q1) {     // (segments of constant closeness -- the two points of one pixel -- need none of this)
                const float r = fast_rcp(d);
                l = q0 + (av - x0) * r * (q1 - q0);
                h = q0 + (bv - x0) * r * (q1 - q0);
            }
            if (l > lo1) { lo2 = lo1; lo1 = l; hib = h; spb = (uint32_t)sp; obest = cnt; }
            else if (l > lo2) lo2 = l;
            if (h > hi1) { hi2 = hi1; hi1 = h; ohi1 = cnt; }
            else if (h > hi2) hi2 = h;
            qmax = fmaxf(qmax, fmaxf(q0, q1));
            ++cnt;
        }
    }
    const uint32_t own = (uint32_t)c.SID[k];
    if (cnt == 1) return (obest == 0) ? spb : (own | kUnresolved);   // (a NaN closeness -- zero-length segment -- is nobody's win)
    if (cnt < 1 || cnt > kMaxCand) return own | kUnresolved;
    // the leader at the left end must lead every other candidate by the margin at both ends
    const float margin = 1e-3f + 1e-4f * qmax;
    const float others_hi = (ohi1 == obest) ? hi2 : hi1;
    const bool lead = safe && (lo1 > lo2 + margin) && (hib > others_hi + margin);
    return lead ? spb : (own | kUnresolved);
}

}  // namespace poly
}  // namespace cs

// cs_polylines.cu -- P: apply_stereo_divergence_polylines (SIG:1912-1992), soft and sharp.
//
// k_polylines<NW, SHARP>   one CTA of NW warps per (tile of output columns, row, frame, eye).  A tile's CTA builds the
//   sorted point table of the SOURCE window that can reach its columns and sweeps only those columns; a row narrow
//   enough is one tile.  Thread t owns 8 consecutive points (4 source columns sharp / 8 soft) in registers from the
//   depth load to the sorted scatter:
//   A points   coord_d in FP64, x = col + 0.5 + coord_d + sep rounded to float32, closeness |coord_d|; sharp emits
//              x -+ 0.45; sentinels (-W, 0) and (2W, 0)                                                   SIG:1919-1936
//   B,C sort   the reference's stable insertion sort (SIG:1941-1946) without sorting: shifts are bounded, so the table is
//              nearly sorted.  Prefix-max / suffix-min scans tell every point whether anything before it is larger or
//              anything after it smaller; if not, its rank is its index.  Inversions within two positions (depth jitter
//              swaps neighbours) are counted in registers; deeper ones (folds) go to a work list.
//   D sets     rank space: END[k] = rank of the end point of the segment starting at sorted point k, REACH = its prefix
//              maximum.  A segment j <= k is active strictly inside interval (k, k+1) iff END[j] > k -- integer
//              compares.  Intervals nobody reaches into have their own segment as the only candidate; the others go to
//              a work list and are resolved to one or two candidates (cs_poly_core.cuh classify_interval).
//   E sweep    thread per output column, float32 with a certified error bound (fast_column); the ~0.3 % of columns it
//              cannot certify are redone at once with the reference's exact FP64 / float32 rounding sequence and list
//              replay (exact_column).  Rows whose replay gives up are listed for k_polylines_exact.
//
// k_polylines_exact   counting sort + one-thread sequential sweep per listed row (or every row: test hook, and
//   disparity ranges too large for a useful tile).
//
// Bytes per pixel and eye: depth 4 B + RGBX8 4 B read (L2-resident scratch), RGBX8 4 B written.
#include <cstdlib>

#include "cs_internal.cuh"

#include <atomic>
#include "cs_poly_core.cuh"

namespace cs {

namespace {

using poly::Tab;

constexpr int kExactThreads = 512;
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

// exponents other than 1 and 2: kept out of line so that the unrolled point loop stays small
__device__ __noinline__ double pow_general(double a, double e) { return pow(a, e); }

// CTA-wide exclusive scan of cnt[0..n) in place (n arbitrary); blockDim = kExactThreads.
__device__ void cta_exclusive_scan(int* cnt, int n, int* s_warp) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + kExactThreads - 1) / kExactThreads;
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
        int v = (lane < kExactThreads / 32) ? s_warp[lane] : 0;
        int vi = v;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += t;
        }
        if (lane < kExactThreads / 32) s_warp[lane] = vi - v;
    }
    __syncthreads();
    int run = s_warp[wid] + inc - sum;
    for (int i = b0; i < b1; ++i) { int c = cnt[i]; cnt[i] = run; run += c; }
    __syncthreads();
}

// Builds the point table of a whole row and its stable sort (counting sort by floor(x), then rank inside the bucket).
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
__device__ __forceinline__ void accumulate(float* color, uint32_t pl, uint32_t pr, bool same, double ip, double sig) {
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
// sequential sweep: one thread replays the reference's list semantics for a whole row
// ------------------------------------------------------------------------------------------
// row_list == nullptr: every (frame, eye, row); else the rows k_polylines listed (row id = (frame * 2 + eye) * h + y).
// gtables != nullptr: rows too wide for a CTA's shared memory keep their tables in global scratch (one slice per CTA).
__global__ void __launch_bounds__(kExactThreads) k_polylines_exact(const WarpArgs a, int sharp, int act_cap,
                                                                   const int* __restrict__ row_list,
                                                                   const int* __restrict__ row_count,
                                                                   int* __restrict__ status,
                                                                   unsigned char* __restrict__ gtables, size_t gtable_bytes) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned char* smem_raw = gtables ? gtables + (size_t)blockIdx.x * gtable_bytes : smem_dyn;
    __shared__ int s_warp[32];
    const int w = a.w;
    const int total = row_list ? *row_count : a.n * 2 * a.h;
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
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
        const int rowid = row_list ? row_list[idx] : idx;
        const int y = rowid % a.h, fe = rowid / a.h, eye = fe & 1, frame = fe >> 1;
        if (a.eye[eye].passthrough) continue;
        build_sorted_points(a, eye, frame, y, c, px, clo, sidx, tmp, rnk, start, s_warp);
        if (threadIdx.x == 0) {
            const int64_t row_off = (int64_t)frame * a.h * w + (int64_t)y * w;
            const uint32_t* img = a.image_u8 + row_off;
            uint32_t* out = a.fused_stereo ? nullptr : a.out[eye] + row_off;
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
                if (a.fused_stereo) {
                    const int64_t o = fused_index(a, eye, frame, y, col);
                    const int r = (int)color[0], gch = (int)color[1], b = (int)color[2];
                    a.fused_stereo[o * 3 + 0] = (float)r / 255.0f;
                    a.fused_stereo[o * 3 + 1] = (float)gch / 255.0f;
                    a.fused_stereo[o * 3 + 2] = (float)b / 255.0f;
                    a.fused_mask[o] = (r + gch + b == 0) ? 1.0f : 0.0f;
                } else {
                    out[col] = pack_rgbx((int)color[0], (int)color[1], (int)color[2]);
                }
            }
            if (overflow) atomicOr(status, 1);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// tile kernel
// ------------------------------------------------------------------------------------------
#ifdef CS_POLY_TIMING
__device__ unsigned long long g_poly_ticks[16];
__device__ __forceinline__ long long cs_clock() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) :: "memory"); return c; }
#define CS_TICK(n) do { if (threadIdx.x == 0) { long long now_ = cs_clock(); atomicAdd(&g_poly_ticks[n], (unsigned long long)(now_ - tick_)); tick_ = now_; } } while (0)
#else
#define CS_TICK(n) do { } while (0)
#endif

struct PolyGeom {             // per eye
    int tile_w[2];            // output columns per tile (>= W: the row is one tile and the window is the whole row)
    int ext[2];               // buckets start this many columns left of the tile (list replays walk back through a fold)
    int ntiles[2];
    int lo_off[2], hi_off[2]; // source window of bucket columns [t0, t0 + tw): [t0 + lo_off, t0 + tw + hi_off) clipped to the row
};

template <int NW, bool SHARP>
struct PolyLayout {
    static constexpr int NT = NW * 32, NP = NT * 8, CPT = SHARP ? 4 : 8, SCAP = NT * CPT;
    static constexpr size_t kBytes = 4 * (size_t)(NP + 8) + 4 * (size_t)NP + 4 * (size_t)NP + 4 * (size_t)(SCAP + 8) + 4 * (size_t)(NP + 8) +
                                     8 * (size_t)NT + 2 * (size_t)NP + 2 * (size_t)(NP + 16) + 2 * (size_t)NP +
                                     2 * (size_t)(SCAP + 8);
    // source columns a CTA can take: every point incl. both sentinels fits the NP slots (4 columns of slack for the
    // 16-byte alignment of the window's first column)
    static constexpr int kCapCols = (NP - 2) / (SHARP ? 2 : 1) - 4;
};

template <int NW, bool SHARP, int TPS>   // TPS: resident threads per SM the register budget is set for
__global__ void __launch_bounds__(NW * 32, TPS / (NW * 32))
k_polylines(const WarpArgs a, const PolyGeom g, int* __restrict__ row_flags, int* __restrict__ row_list,
            int* __restrict__ counters, int mode_flags) {
    using L = PolyLayout<NW, SHARP>;
    constexpr int NT = L::NT, NP = L::NP, CPT = L::CPT, SCAP = L::SCAP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = a.w, tile = blockIdx.x, y = blockIdx.y, frame = blockIdx.z >> 1, eye = blockIdx.z & 1;
    if (a.eye[eye].passthrough || tile >= g.ntiles[eye]) return;
#ifdef CS_POLY_TIMING
    long long tick_ = cs_clock();
#endif

    // ---- geometry: own output columns [o0, o0 + own), buckets [t0, t0 + tw), source window [s0, s0 + w)
    int t0 = 0, tw = W, own = W, s0 = 0, w = W;
    if (g.tile_w[eye] < W) {
        const int o0 = tile * g.tile_w[eye];
        own = min(g.tile_w[eye], W - o0);
        t0 = max(o0 - g.ext[eye], 0);
        tw = o0 + own - t0;
        s0 = min(max(t0 + g.lo_off[eye], 0), W - 1) & ~3;
        w = min(max(t0 + tw + g.hi_off[eye], s0 + 1), W) - s0;
    }
    const int npts = (SHARP ? 2 * w : w) + 2, nsg = npts - 1;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;

    float* XA = reinterpret_cast<float*>(smem_raw);             // [NP + 8]  x in source order; X[1] is 16-byte aligned
    float* X = XA + 3;
    float* SX = XA + NP + 8;                                     // [NP]      x in sorted order
    uint32_t* ER = reinterpret_cast<uint32_t*>(SX + NP);         // [NP]      END | REACH << 16
    float* QA = reinterpret_cast<float*>(ER + NP);               // [SCAP + 8] closeness, padded by one on either side
    float* Q = QA + 3;
    uint32_t* IMGA = reinterpret_cast<uint32_t*>(QA + SCAP + 8); // [NP + 8]  RGBX per source point (laid out like X)
    uint32_t* IMGP = IMGA + 3;
    float* TMX = reinterpret_cast<float*>(IMGA + NP + 8);        // [NT] max of every point up to the end of thread t's
    float* TMN = TMX + NT;                                       // [NT] min of every point from the start of thread t's
    uint16_t* SID = reinterpret_cast<uint16_t*>(TMN + NT);       // [NP]
    uint16_t* WSP = SID + NP;                                    // [NP + 16] (holds the source-order ranks during the sort)
    uint16_t* RNK = WSP + 7;                                     //           RNK[1] is 16-byte aligned
    uint16_t* LIST = WSP + NP + 16;                              // [NP] work lists: deep inversions, then hard intervals
    uint16_t* START = LIST + NP;                                 // [SCAP + 8]
    __shared__ float s_wa[32], s_wb[32];
    __shared__ int s_wr[32];
    __shared__ int s_nslow, s_nhard, s_nflag;
    __shared__ float s_q255[256];     // k / 255.0f (GS:365-378), for the fused composed output
    __shared__ Tab s_tab;   // the exact path is a real function call and takes the tables by reference
    for (int k = t; k < 256; k += NT) s_q255[k] = kQ255[k];
    if (t == 0) {
        s_nslow = 0; s_nhard = 0; s_nflag = 0;
        s_tab.X = X; s_tab.SX = SX; s_tab.ER = ER; s_tab.SID = SID; s_tab.WSP = WSP; s_tab.Q = Q; s_tab.IMGP = IMGP;
        s_tab.START = START; s_tab.w = w; s_tab.npts = npts; s_tab.nsg = nsg; s_tab.t0 = t0;
    }

    // ---- A: depth + image of this thread's CPT source columns -> 8 points in registers
    const int64_t row_off = ((int64_t)frame * a.h + y) * W;
    const int c0 = t * CPT, i0 = 1 + 8 * t;
    float v[8];
    {
        const float* dep = a.depth[eye] + row_off + s0;
        const uint32_t* img = a.image_u8 + row_off + s0;
        float dv[CPT];
        uint32_t iv[CPT];
        const bool vec = ((reinterpret_cast<uintptr_t>(dep) | reinterpret_cast<uintptr_t>(img)) & 15) == 0;
        if (vec && c0 + CPT <= w) {
#pragma unroll
            for (int q = 0; q < CPT / 4; ++q) {
                const float4 d4 = __ldg(reinterpret_cast<const float4*>(dep + c0) + q);
                const uint4 i4 = __ldg(reinterpret_cast<const uint4*>(img + c0) + q);
                dv[4 * q] = d4.x; dv[4 * q + 1] = d4.y; dv[4 * q + 2] = d4.z; dv[4 * q + 3] = d4.w;
                iv[4 * q] = i4.x; iv[4 * q + 1] = i4.y; iv[4 * q + 2] = i4.z; iv[4 * q + 3] = i4.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int cc = min(c0 + j, w - 1);   // the column after the last repeats it (padding of IMGP)
                const bool in = c0 + j <= w;
                dv[j] = in ? __ldg(dep + cc) : 0.0f;
                iv[j] = in ? __ldg(img + cc) : 0u;
            }
        }
        // (after the row's loads are in flight: the frame statistics are another dependent global load)
        float scale;
        const Normalizer norm = pl_normalizer(a, eye, frame, &scale);
        const double div_px = a.eye[eye].div_px, sep_px = a.eye[eye].sep_px;
        const double base = (double)(c0 + s0) + 0.5;
        float qv[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const bool in = c0 + j < w;
            float d = dv[j];
            if (scale != 1.0f) d = d * scale;
            const float nd = norm(d);
            const double an = (double)fabsf(nd);
            double p;
            if (a.expo == 2.0) p = an * an;
            else if (a.expo == 1.0) p = an;
            else p = pow_general(an, a.expo);
            const double sp = (nd >= 0.0f) ? p : -p;
            const double cd = sp * div_px;
            double cx = (base + (double)j) + cd;
            cx = cx + sep_px;
            qv[j] = in ? (float)fabs(cd) : 0.0f;
            if (SHARP) {
                v[2 * j] = in ? (float)(cx - 0.45) : INFINITY;
                v[2 * j + 1] = in ? (float)(cx + 0.45) : INFINITY;
            } else {
                v[j] = in ? (float)cx : INFINITY;
            }
        }
        const int se = npts - 1 - i0;   // slot of the right sentinel, if it is one of this thread's
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (e == se) v[e] = (float)(2.0 * W);
        reinterpret_cast<float4*>(X + i0)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(X + i0)[1] = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
        for (int q = 0; q < CPT / 4; ++q)
            reinterpret_cast<float4*>(Q + 1 + c0)[q] = make_float4(qv[4 * q], qv[4 * q + 1], qv[4 * q + 2], qv[4 * q + 3]);
        // colour per POINT (sharp: both points of a pixel), indexed like X: the column after the last repeats it, which is
        // the right sentinel's colour
        if (SHARP) {
            reinterpret_cast<uint4*>(IMGP + i0)[0] = make_uint4(iv[0], iv[0], iv[1], iv[1]);
            reinterpret_cast<uint4*>(IMGP + i0)[1] = make_uint4(iv[2], iv[2], iv[3], iv[3]);
        } else {
#pragma unroll
            for (int q = 0; q < CPT / 4; ++q)
                reinterpret_cast<uint4*>(IMGP + i0)[q] = make_uint4(iv[4 * q], iv[4 * q + 1], iv[4 * q + 2], iv[4 * q + 3]);
        }
        if (t == 0) { X[0] = (float)(-1.0 * W); Q[0] = 0.0f; IMGP[0] = iv[0]; }
        // vector path: the pad after the last column (the scalar path loaded it in place)
        if (c0 + CPT == w && vec) IMGP[npts - 1] = iv[CPT - 1];
    }

    CS_TICK(0);
    // ---- B: prefix max / suffix min of x in source order (per-thread totals here, per-point values in C)
    float em, en;
    {
        const float m = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
        const float n = fminf(fminf(fminf(v[0], v[1]), fminf(v[2], v[3])), fminf(fminf(v[4], v[5]), fminf(v[6], v[7])));
        float im = m, in_ = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float u = __shfl_up_sync(0xffffffffu, im, o);
            if (lane >= o) im = fmaxf(im, u);
            const float d = __shfl_down_sync(0xffffffffu, in_, o);
            if (lane + o < 32) in_ = fminf(in_, d);
        }
        em = __shfl_up_sync(0xffffffffu, im, 1);     // max over the earlier lanes of the warp
        en = __shfl_down_sync(0xffffffffu, in_, 1);  // min over the later lanes
        if (lane == 0) em = -INFINITY;
        if (lane == 31) en = INFINITY;
        if (lane == 31) s_wa[wid] = im;
        if (lane == 0) s_wb[wid] = in_;
        __syncthreads();
        CS_TICK(1);
        {   // totals of the other warps: one scan over the NW per-warp values
            float wa = (lane < NW) ? s_wa[lane] : -INFINITY, wb = (lane < NW) ? s_wb[lane] : INFINITY;
#pragma unroll
            for (int o = 1; o < NW; o <<= 1) {
                const float u = __shfl_up_sync(0xffffffffu, wa, o);
                if (lane >= o) wa = fmaxf(wa, u);
                const float d = __shfl_down_sync(0xffffffffu, wb, o);
                if (lane + o < 32) wb = fminf(wb, d);
            }
            const float ea = __shfl_sync(0xffffffffu, wa, (wid + 31) & 31), eb = __shfl_sync(0xffffffffu, wb, (wid + 1) & 31);
            if (wid > 0) em = fmaxf(em, ea);
            if (wid < NW - 1) en = fminf(en, eb);
        }
        TMX[t] = fmaxf(em, m);
        TMN[t] = fminf(en, n);
    }

    // ---- C: stable rank of every point = i - #(earlier, larger) + #(later, smaller).  Counted over two neighbours on
    // either side, which is the whole answer when nothing further away is out of order with the point (depth jitter
    // swaps neighbours); the scans certify that.  Deeper inversions (folds) are counted from shared memory.
    int r[8];
    {
        // P[e] = max of everything before point e, S[e] = min of everything after it
        float P[8], S[8];
        P[0] = em;
#pragma unroll
        for (int e = 1; e < 8; ++e) P[e] = fmaxf(P[e - 1], v[e - 1]);
        S[7] = en;
#pragma unroll
        for (int e = 6; e >= 0; --e) S[e] = fminf(S[e + 1], v[e + 1]);
        // the neighbouring threads' edge points and what bounds them; across a warp boundary the bounds are the
        // conservative ones and nothing is counted
        float pv6 = __shfl_up_sync(0xffffffffu, v[6], 1), pv7 = __shfl_up_sync(0xffffffffu, v[7], 1);
        float pb6 = __shfl_up_sync(0xffffffffu, P[6], 1), pb7 = __shfl_up_sync(0xffffffffu, P[7], 1);
        float nv0 = __shfl_down_sync(0xffffffffu, v[0], 1), nv1 = __shfl_down_sync(0xffffffffu, v[1], 1);
        float na0 = __shfl_down_sync(0xffffffffu, S[0], 1), na1 = __shfl_down_sync(0xffffffffu, S[1], 1);
        if (lane == 0) { pv6 = -INFINITY; pv7 = -INFINITY; pb6 = em; pb7 = em; }
        if (lane == 31) { nv0 = INFINITY; nv1 = INFINITY; na0 = en; na1 = en; }
        uint32_t slow = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float x = v[e];
            const float p1 = (e >= 1) ? v[e >= 1 ? e - 1 : 0] : pv7;
            const float p2 = (e >= 2) ? v[e >= 2 ? e - 2 : 0] : (e == 1 ? pv7 : pv6);
            const float b3 = (e >= 2) ? P[e >= 2 ? e - 2 : 0] : (e == 1 ? pb7 : pb6);     // everything before p2
            const float n1 = (e <= 6) ? v[e <= 6 ? e + 1 : 7] : nv0;
            const float n2 = (e <= 5) ? v[e <= 5 ? e + 2 : 7] : (e == 6 ? nv0 : nv1);
            const float a3 = (e <= 5) ? S[e <= 5 ? e + 2 : 7] : (e == 6 ? na0 : na1);     // everything after n2
            r[e] = i0 + e + ((n1 < x) ? 1 : 0) + ((n2 < x) ? 1 : 0) - ((p1 > x) ? 1 : 0) - ((p2 > x) ? 1 : 0);
            if (b3 > x || a3 < x) slow |= 1u << e;
        }
        *reinterpret_cast<uint4*>(RNK + i0) = make_uint4((uint32_t)r[0] | ((uint32_t)r[1] << 16), (uint32_t)r[2] | ((uint32_t)r[3] << 16),
                                                         (uint32_t)r[4] | ((uint32_t)r[5] << 16), (uint32_t)r[6] | ((uint32_t)r[7] << 16));
        if (slow) {
            int base = atomicAdd(&s_nslow, __popc(slow));
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (slow & (1u << e)) LIST[base++] = (uint16_t)(i0 + e);
        }
        __syncthreads();
        CS_TICK(2);
        const int nslow = s_nslow;
        if (nslow) {
            // thread blocks of 8 points outwards from the point's own; a block is skipped, and the walk ends, as soon as
            // its running bound says that nothing at or beyond it can be out of order with x
            for (int q = t; q < nslow; q += NT) {
                const int i = LIST[q];
                const float x = X[i];
                const int tb0 = (i - 1) >> 3;
                int rr = i;
                for (int tb = tb0; tb >= 0; --tb) {
                    if (tb != tb0 && !(TMX[tb] > x)) break;
                    const float4 u0 = reinterpret_cast<const float4*>(X + 1 + 8 * tb)[0], u1 = reinterpret_cast<const float4*>(X + 1 + 8 * tb)[1];
                    const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                    const int lim = (tb == tb0) ? (i - 1) - 8 * tb : 8;   // own block: only the points before i
#pragma unroll
                    for (int e = 0; e < 8; ++e) rr -= (e < lim && u[e] > x) ? 1 : 0;
                }
                for (int tb = tb0; tb < NT; ++tb) {
                    if (tb != tb0 && !(TMN[tb] < x)) break;
                    const float4 u0 = reinterpret_cast<const float4*>(X + 1 + 8 * tb)[0], u1 = reinterpret_cast<const float4*>(X + 1 + 8 * tb)[1];
                    const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                    const int lim = (tb == tb0) ? (i - 1) - 8 * tb : -1;  // own block: only the points after i
#pragma unroll
                    for (int e = 0; e < 8; ++e) rr += (e > lim && u[e] < x) ? 1 : 0;
                }
                RNK[i] = (uint16_t)rr;
            }
            __syncthreads();
            CS_TICK(3);
            const uint4 pk = *reinterpret_cast<const uint4*>(RNK + i0);
            r[0] = pk.x & 0xFFFF; r[1] = pk.x >> 16; r[2] = pk.y & 0xFFFF; r[3] = pk.y >> 16;
            r[4] = pk.z & 0xFFFF; r[5] = pk.z >> 16; r[6] = pk.w & 0xFFFF; r[7] = pk.w >> 16;
        }
        const int rn = RNK[i0 + 8];   // rank of the next thread's first point
        uint16_t* END16 = reinterpret_cast<uint16_t*>(ER);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int i = i0 + e;
            if (i < npts) {
                const int k = r[e];
                SX[k] = v[e];
                SID[k] = (uint16_t)i;
                END16[2 * k] = (uint16_t)((i < npts - 1) ? (e < 7 ? r[e < 7 ? e + 1 : 7] : rn) : 0);
            }
        }
        if (t == 0) { SX[0] = (float)(-1.0 * W); SID[0] = 0; END16[0] = (uint16_t)r[0]; }
    }
    __syncthreads();
    CS_TICK(4);

    // ---- D: REACH = prefix max of END in sorted order; bucket starts; intervals with a single candidate
    {
        const int k0 = 8 * t;
        uint32_t er[8], sid[8];
        float sx[8];
        {
            const uint4 a0 = reinterpret_cast<const uint4*>(ER + k0)[0], a1 = reinterpret_cast<const uint4*>(ER + k0)[1];
            er[0] = a0.x; er[1] = a0.y; er[2] = a0.z; er[3] = a0.w; er[4] = a1.x; er[5] = a1.y; er[6] = a1.z; er[7] = a1.w;
            const float4 b0 = reinterpret_cast<const float4*>(SX + k0)[0], b1 = reinterpret_cast<const float4*>(SX + k0)[1];
            sx[0] = b0.x; sx[1] = b0.y; sx[2] = b0.z; sx[3] = b0.w; sx[4] = b1.x; sx[5] = b1.y; sx[6] = b1.z; sx[7] = b1.w;
            const uint4 c4 = *reinterpret_cast<const uint4*>(SID + k0);
            sid[0] = c4.x & 0xFFFF; sid[1] = c4.x >> 16; sid[2] = c4.y & 0xFFFF; sid[3] = c4.y >> 16;
            sid[4] = c4.z & 0xFFFF; sid[5] = c4.z >> 16; sid[6] = c4.w & 0xFFFF; sid[7] = c4.w >> 16;
        }
        const float sxp = (t > 0) ? SX[k0 - 1] : 0.0f;
        int end[8], lr[8];
        int m = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            end[e] = (k0 + e < npts) ? (int)(er[e] & 0xFFFFu) : 0;
            m = max(m, end[e]);
            lr[e] = m;
        }
        int im = m;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, im, o);
            if (lane >= o) im = max(im, u);
        }
        int emr = __shfl_up_sync(0xffffffffu, im, 1);
        if (lane == 0) emr = 0;
        if (lane == 31) s_wr[wid] = im;
        __syncthreads();
        CS_TICK(5);
        {
            int wr = (lane < NW) ? s_wr[lane] : 0;
#pragma unroll
            for (int o = 1; o < NW; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, wr, o);
                if (lane >= o) wr = max(wr, u);
            }
            const int er0 = __shfl_sync(0xffffffffu, wr, (wid + 31) & 31);
            if (wid > 0) emr = max(emr, er0);
        }
        uint32_t wsp[8];
        uint32_t hard = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = k0 + e;
            const int rprev = (e == 0) ? emr : max(emr, lr[e > 0 ? e - 1 : 0]);   // REACH[k - 1]
            er[e] = (uint32_t)end[e] | ((uint32_t)max(emr, lr[e]) << 16);
            // nobody reaches past this interval's left point: its own segment wins, if that one goes forward
            wsp[e] = sid[e] | ((end[e] > k) ? 0u : poly::kUnresolved);
            if (k < nsg && rprev > k) hard |= 1u << e;
        }
        reinterpret_cast<uint4*>(ER + k0)[0] = make_uint4(er[0], er[1], er[2], er[3]);
        reinterpret_cast<uint4*>(ER + k0)[1] = make_uint4(er[4], er[5], er[6], er[7]);
        *reinterpret_cast<uint4*>(WSP + k0) = make_uint4(wsp[0] | (wsp[1] << 16), wsp[2] | (wsp[3] << 16),
                                                         wsp[4] | (wsp[5] << 16), wsp[6] | (wsp[7] << 16));
        if (hard) {
            int base = atomicAdd(&s_nhard, __popc(hard));
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (hard & (1u << e)) LIST[base++] = (uint16_t)(k0 + e);
        }
        // bucket b = floor(x) - t0 + 1 clamped to [0, tw + 1] (the float -> int conversion saturates)
        int bprev = -1;
        if (t > 0) bprev = (k0 - 1 < npts) ? min(max(__float2int_rd(sxp) - t0 + 1, 0), tw + 1) : tw + 1;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = k0 + e;
            const int b = (k < npts) ? min(max(__float2int_rd(sx[e]) - t0 + 1, 0), tw + 1) : bprev;
            if (b != bprev) {
                START[b] = (uint16_t)k;
                if (b > bprev + 1)                                              // empty buckets in between (gaps)
                    for (int q = bprev + 1; q < b; ++q) START[q] = (uint16_t)k;
                bprev = b;
            }
        }
        if (k0 <= npts - 1 && npts - 1 < k0 + 8) START[tw + 2] = (uint16_t)npts;
    }
    __syncthreads();
    CS_TICK(6);

    Tab tab;   // register copy for the inlined float32 path
    tab.X = X; tab.SX = SX; tab.ER = ER; tab.SID = SID; tab.WSP = WSP; tab.Q = Q; tab.IMGP = IMGP; tab.START = START;
    tab.w = w; tab.npts = npts; tab.nsg = nsg; tab.t0 = t0;
    {
        const int nhard = s_nhard;
        for (int q = t; q < nhard; q += NT) {
            const int k = LIST[q];
            WSP[k] = (uint16_t)poly::classify_interval<SHARP>(tab, k);
        }
    }
    __syncthreads();
    CS_TICK(7);

    // ---- E: sweep.  Warps take the 32-column blocks round-robin (a shared counter balanced the blocks inside folds, which
    // cost several times more, but its atomic + shuffle was 6 % of the kernel's instructions and the balance buys nothing:
    // the other CTA of the SM fills the gaps).  Columns the float32 path cannot certify are listed and redone afterwards,
    // four per warp pass: inside the sweep each of them would stall its whole warp for longer than a block takes.
    uint32_t* out = a.fused_stereo ? nullptr : a.out[eye] + row_off + t0;
    const uint64_t pol = policy_evict_first();
    // one finished pixel: the RGBX8 eye image, or (fused) its place in the composed float32 tensor and the black-pixel mask
    const int64_t o_row = a.fused_stereo ? fused_index(a, eye, frame, y, t0) : 0;      // the row's place in the composed tensors
    float* const dst_row = a.fused_stereo + o_row * 3;
    float* const msk_row = a.fused_mask + o_row;
    auto emit = [&](int col, uint32_t px) {
        if (out) { out[col] = px; return; }
        float* dst = dst_row + col * 3;
        st_stream_f1(dst, s_q255[px & 255u], pol);
        st_stream_f1(dst + 1, s_q255[(px >> 8) & 255u], pol);
        st_stream_f1(dst + 2, s_q255[(px >> 16) & 255u], pol);
        st_stream_f1(msk_row + col, ((px & 0x00FFFFFFu) == 0u) ? 1.0f : 0.0f, pol);
    };
    const int nblk = (own + 31) >> 5, first = tw - own;   // the tile's own columns are the last `own` bucket columns
    const bool all_exact = (mode_flags & 8) != 0;
    for (int blk = wid; blk < nblk; blk += NW) {
        const int col = first + (blk << 5) + lane;
        const bool in = col < tw;
        uint32_t px = 0;
        const bool ok = in && poly::fast_column<SHARP>(tab, col, &px) && !all_exact;
        if (in && !ok) LIST[atomicAdd(&s_nflag, 1)] = (uint16_t)col;
        // (pixel by pixel: the 128-bit warp-collective form used by the row kernels measured 2 % slower here)
        if (ok) emit(col, px);
    }
    CS_TICK(8);
    __syncthreads();
    CS_TICK(9);
    bool give_up = false;
    {
        const int nflag = s_nflag;
        // (from the last warp down: with round-robin blocks the first warps have swept one block more)
        for (int base = 4 * (NW - 1 - wid); base < nflag; base += 4 * NW) {     // four columns per warp at a time, eight lanes each
            const int ncols = min(4, nflag - base);
            uint32_t big = 0;
            uint32_t px = poly::exact_columns_quad<SHARP>(s_tab, LIST + base, ncols, &big);
            const int gq = lane >> 3;
            if (gq < ncols && !((big >> gq) & 1u)) {
                if (px & poly::kGaveUp) { give_up = true; px &= ~poly::kGaveUp; }
                if ((lane & 7) == 0) emit(LIST[base + gq], px);
            }
            while (big) {                                         // columns with more than eight sub-intervals: a warp each
                const int gb = __ffs(big) - 1;
                big &= big - 1;
                const int col = LIST[base + gb];
                uint32_t pb = poly::exact_column_warp<SHARP>(s_tab, col);
                if (pb & poly::kGaveUp) { give_up = true; pb &= ~poly::kGaveUp; }
                if (lane == 0) emit(col, pb);
            }
        }
#ifdef CS_POLY_TIMING
        if (t == 0) { atomicAdd(&g_poly_ticks[12], (unsigned long long)nflag); atomicAdd(&g_poly_ticks[13], 1ull); }
#endif
    }
    CS_TICK(10);
    if (give_up) {
        // a list replay did not fit its budget (or its history starts left of this tile): the row is redone sequentially
        const int rowid = (frame * 2 + eye) * a.h + y;
        if (atomicOr(&row_flags[rowid], 1) == 0) row_list[atomicAdd(&counters[1], 1)] = rowid;
    }
}

static size_t exact_smem(int w, int sharp, int act_cap) {
    size_t npts = (size_t)(sharp ? 2 * w : w) + 2;
    size_t np2 = npts + (npts & 1);
    return npts * 4 + (size_t)w * 4 + (size_t)(w + 4) * 4 + np2 * 2 * 3 + (size_t)act_cap * 2;
}
constexpr int kExactGlobalGrid = 64;     // CTAs of the sequential kernel when its tables live in global scratch
constexpr size_t kMaxSmem = 226 * 1024;   // dynamic shared memory a CTA may ask for here (1 KB left for static arrays)
static long long exact_act_cap_ref(const WarpArgs& a) {
    const double dmax = fmax(fabs(a.eye[0].div_px), fabs(a.eye[1].div_px));
    return 5ll * (long long)dmax + 25;    // the reference's own list capacity, SIG:1947
}
static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }
// scratch: [16] counters (0: status bits, 1: listed rows) + [n*2*h] row flags + [n*2*h] row list
//          (+ kExactGlobalGrid table slices for rows whose sequential-sweep tables exceed a CTA's shared memory)
static size_t poly_header_bytes(int n, int h) { return align256(((size_t)n * 2 * h * 2 + 16) * sizeof(int)); }
static size_t exact_gtable_bytes(int w) {
    // sharp tables (the larger), active list of the largest capacity the widget ranges can ask for: |div_px| <= 0.3 w
    return align256(exact_smem(w, 1, 0) + 2 * (size_t)(5 * (long long)(0.3 * w + 1) + 25));
}
size_t polylines_scratch_bytes(int n, int h, int w) {
    size_t b = poly_header_bytes(n, h);
    if (exact_smem(w, 1, 64) + 64 > kMaxSmem) b += (size_t)kExactGlobalGrid * exact_gtable_bytes(w);
    return b;
}

static cudaError_t launch_exact(const WarpArgs& a, int sharp, bool listed, int* counters, int* list, cudaStream_t s) {
    const long long cap_ref = exact_act_cap_ref(a);
    const size_t base = exact_smem(a.w, sharp, 0);
    const int rows = a.n * 2 * a.h;
    unsigned char* gtables = nullptr;
    size_t gbytes = 0, es = 0;
    int act_cap, grid;
    if (base + 64 <= kMaxSmem) {
        const long long cap_fit = (long long)((kMaxSmem - base) / 2);
        act_cap = (int)(cap_ref < cap_fit ? cap_ref : cap_fit);
        es = exact_smem(a.w, sharp, act_cap);
        if (es > 48 * 1024) cudaFuncSetAttribute(k_polylines_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)es);
        // listed rows are rare (usually none): a small grid that strides over the list
        grid = listed ? (rows < 296 ? rows : 296) : rows;
    } else {
        // very wide rows: tables in global scratch, behind the header
        gbytes = exact_gtable_bytes(a.w);
        const size_t hdr = poly_header_bytes(a.n, a.h);
        if (a.scratch_bytes < hdr + (size_t)kExactGlobalGrid * gbytes) return cudaErrorInvalidValue;
        gtables = reinterpret_cast<unsigned char*>(a.scratch) + hdr;
        const long long cap_fit = (long long)((gbytes - exact_smem(a.w, sharp, 0)) / 2);
        act_cap = (int)(cap_ref < cap_fit ? cap_ref : cap_fit);
        grid = rows < kExactGlobalGrid ? rows : kExactGlobalGrid;
    }
    prof_begin(K_POLY_EXACT, s);
    k_polylines_exact<<<grid, kExactThreads, es, s>>>(a, sharp, act_cap, listed ? list : nullptr, counters + 1, counters,
                                                      gtables, gbytes);
    prof_end(K_POLY_EXACT, s);
    count_launch();
    return cudaGetLastError();
}

struct PolyPlan {
    int nw;
    PolyGeom g;
    int max_tiles;
    double cost;
};

// Tile geometry for CTAs of nw warps; false when no useful tile fits.  cost ~ relative work per output column.
static bool plan_for(int nw, const WarpArgs& a, bool sharp, bool force_tiles, PolyPlan* p) {
    const int ppc = sharp ? 2 : 1;
    const int np = nw * 256, cap_cols = (np - 2) / ppc - 4;
    const double span = pow(fmax((double)a.conv, 1.0 - (double)a.conv), a.expo);
    p->nw = nw; p->max_tiles = 0; p->cost = 0.0;
    for (int eye = 0; eye < 2; ++eye) {
        p->g.tile_w[eye] = a.w; p->g.ext[eye] = 0; p->g.ntiles[eye] = 0; p->g.lo_off[eye] = 0; p->g.hi_off[eye] = 0;
        if (a.eye[eye].passthrough) continue;
        // |shift| <= |div_px| * max(conv, 1 - conv)^expo because the normalised depth lies in [-conv, 1 - conv]
        const double reach = fabs(a.eye[eye].div_px) * span + 1.0;
        if (reach + fabs(a.eye[eye].sep_px) > 1e6) return false;
        // a source column c lands within reach + 1 of c + sep: the window of bucket columns [t0, t0 + tw)
        p->g.lo_off[eye] = (int)floor(-a.eye[eye].sep_px - (reach + 4.0));
        p->g.hi_off[eye] = (int)ceil(-a.eye[eye].sep_px + (reach + 4.0));
        if (!force_tiles && a.w <= cap_cols) {
            p->g.ntiles[eye] = 1;
            // whole rows have no window overlap, and large CTAs keep the warps of an SM sub-partition in the same code
            // (measured: 16-warp whole rows 2.63 ms vs 4-warp tiles 2.85 ms per 16 frames of 1080p sharp)
            p->cost += 0.40 + 0.55 * (double)(np / ppc) / a.w;
        } else {
            const int rmax = (int)ceil(reach);
            const int ext = 2 * rmax + 8;                       // a fold is at most 2 * reach wide
            const int guard = 2 * (rmax + 5) + 2 + ext + 4;
            int tile_w = force_tiles ? 64 : ((cap_cols - guard) / 32) * 32;
            if (tile_w < 64 || tile_w + guard > cap_cols) return false;
            int ntiles = (a.w + tile_w - 1) / tile_w;
            if (!force_tiles) tile_w = (((a.w + ntiles - 1) / ntiles + 31) / 32) * 32;   // even tiles
            ntiles = (a.w + tile_w - 1) / tile_w;
            p->g.tile_w[eye] = tile_w; p->g.ext[eye] = ext; p->g.ntiles[eye] = ntiles;
            p->cost += 0.45 + 0.55 * (double)(np / ppc) * ntiles / a.w;
        }
        if (p->g.ntiles[eye] > p->max_tiles) p->max_tiles = p->g.ntiles[eye];
    }
    return p->max_tiles > 0;
}

template <int NW, bool SHARP, int TPS>
static cudaError_t launch_tiles_occ(const WarpArgs& a, const PolyPlan& p, int* counters, int* flags, int* list, cudaStream_t s) {
    using L = PolyLayout<NW, SHARP>;
    // the opt-in to more than 48 KB of dynamic shared memory is a per-device attribute of the function: once per device
    // (benign race: setting it twice is harmless)
    static std::atomic<unsigned long long> attr_done{0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !((attr_done.load(std::memory_order_relaxed) >> dev) & 1ull)) {
        e = cudaFuncSetAttribute(k_polylines<NW, SHARP, TPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kBytes);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done.fetch_or(1ull << dev, std::memory_order_relaxed);
    }
    prof_begin(K_POLY_FAST, s);
    k_polylines<NW, SHARP, TPS><<<dim3(p.max_tiles, a.h, 2 * a.n), NW * 32, L::kBytes, s>>>(a, p.g, flags, list, counters, a.flags);
    prof_end(K_POLY_FAST, s);
    count_launch();
    return cudaGetLastError();
}
template <int NW, bool SHARP>
static cudaError_t launch_tiles(const WarpArgs& a, const PolyPlan& p, int* counters, int* flags, int* list, cudaStream_t s) {
    // 64 registers per thread (1024 resident threads per SM): 80- and 92-register builds have no spills but measured
    // slower -- the occupancy they cost outweighs the 48 bytes of spilled loop invariants
    return launch_tiles_occ<NW, SHARP, 1024>(a, p, counters, flags, list, s);
}

#ifdef CS_POLY_TIMING
extern "C" __attribute__((visibility("default"))) void cs_poly_ticks(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, g_poly_ticks, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_poly_ticks, z, sizeof(z)); }
}
#endif

// flags bit 0 = replay every row with the sequential kernel, bit 2 = force 64-column tiles, bit 3 = every column through
// exact_column (tests); bits 8-15 = CTA size in warps (0 = choose)
cudaError_t launch_polylines(const WarpArgs& a, cudaStream_t s) {
    const bool sharp = a.fill == CS_FILL_POLYLINES_SHARP;
    if (a.scratch_bytes < polylines_scratch_bytes(a.n, a.h, a.w)) return cudaErrorInvalidValue;
    if ((sharp ? 2 * (long long)a.w : (long long)a.w) + 2 > 65535) return cudaErrorInvalidValue;
    int* counters = reinterpret_cast<int*>(a.scratch);
    int* flags = counters + 16;
    int* list = flags + (size_t)a.n * 2 * a.h;
    cudaError_t e = cudaMemsetAsync(counters, 0, (16 + (size_t)a.n * 2 * a.h) * sizeof(int), s);
    if (e != cudaSuccess) return e;
    const bool force_exact = (a.flags & 1) != 0, force_tiles = (a.flags & 4) != 0;
    int want_nw = (a.flags >> 8) & 0xFF;
    if (!want_nw) {
        static int env_nw = -1;
        if (env_nw < 0) { const char* v = getenv("COMFYSTEREO_POLY_NW"); env_nw = v ? atoi(v) : 0; }
        want_nw = env_nw;
    }
    if (!force_exact) {
        static const int kSizes[] = {4, 8, 16};
        PolyPlan best;
        bool have = false;
        for (int nw : kSizes) {
            if (want_nw && nw != want_nw) continue;
            PolyPlan p;
            if (!plan_for(nw, a, sharp, force_tiles, &p)) continue;
            if (force_tiles) { best = p; have = true; break; }
            if (!have || p.cost < best.cost - 1e-9) { best = p; have = true; }
        }
        if (have) {
#define CS_POLY_CASE(NWV) case NWV: e = sharp ? launch_tiles<NWV, true>(a, best, counters, flags, list, s) \
                                              : launch_tiles<NWV, false>(a, best, counters, flags, list, s); break;
            switch (best.nw) {
                CS_POLY_CASE(4)
                CS_POLY_CASE(8)
                CS_POLY_CASE(16)
                default: e = cudaErrorInvalidValue;
            }
#undef CS_POLY_CASE
            if (e != cudaSuccess) return e;
            return launch_exact(a, sharp, true, counters, list, s);   // rows a tile could not finish (usually none)
        }
    }
    // the test hook, or a disparity range so large that no useful tile fits: sequential kernel for every row
    return launch_exact(a, sharp, false, counters, list, s);
}

}  // namespace cs

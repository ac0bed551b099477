// cs_polylines.cu -- P: apply_stereo_divergence_polylines (SIG:1912-1992), soft and sharp.
//
// One CTA per (row, frame, eye); the whole row lives in shared memory.
//
//   points    thread per source column: coord_d in FP64, x = col + 0.5 + coord_d + sep rounded to
//             float32, closeness |coord_d| float32; sharp emits x -+ 0.45 (two points per column).
//             Sentinels (-W, 0, col 0) and (2W, 0, col W-1) bracket the row.          SIG:1919-1936
//   sort      the reference's stable insertion sort by x (segments ride along, SIG:1941-1946) is
//             replaced by a counting sort: bucket = floor(x) clamped to [-1, W], shared-memory
//             histogram + CTA scan, then every point ranks itself inside its (tiny) bucket by
//             (x, source index).  Stable, deterministic, O(points).
//   cover     output column c owns the sorted points of bucket c; the sub-intervals the reference
//             visits for c are (pred, b0), (b0, b1), ..., (b_last, succ)               SIG:1955-1961
//   fast sweep (k_polylines) thread per output column.  The reference's "active list" at a centre
//             ctr is the SET of segments with x0 < ctr <= x1; the thread finds it by walking back
//             from the interval's left point while a prefix maximum of segment ends still reaches
//             ctr.  Selection (max interpolated closeness with 0 < ip < 1) is done in FP64 exactly
//             as the reference does.  The set is enough unless the choice depends on the ORDER of
//             the reference's append / swap-remove list -- an exact closeness tie or "no valid
//             candidate" (Q7).  Such a row raises a flag and is redone by
//   exact sweep (k_polylines_exact) the same points/sort, then ONE thread replays the reference's
//             sequential sweep with its list semantics, bit for bit.  Slow, rare on real depth.
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
// fast sweep: thread per output column, set-based selection, flags order-dependent rows
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPolyThreads) k_polylines(const WarpArgs a, int sharp, int* __restrict__ row_flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int w = a.w, y = blockIdx.x, frame = blockIdx.y, eye = blockIdx.z;
    if (a.eye[eye].passthrough) return;
    RowCtx c;
    c.w = w; c.sharp = sharp != 0; c.npts = (sharp ? 2 * w : w) + 2; c.nsg = c.npts - 1;
    const int npts = c.npts, nsg = c.nsg;
    float* px = reinterpret_cast<float*>(smem_raw);
    float* clo = px + npts;
    float* sx = clo + w;            // [npts] sorted x
    float* reach = sx + npts;       // [npts] prefix max (sorted order) of segment end x1
    int* start = reinterpret_cast<int*>(reach + npts);
    uint32_t* simg = reinterpret_cast<uint32_t*>(start + (w + 4));   // [w] RGBX8 row
    unsigned short* sidx = reinterpret_cast<unsigned short*>(simg + w);
    unsigned short* tmp = sidx + (npts + (npts & 1));
    unsigned short* rnk = tmp + (npts + (npts & 1));
    __shared__ int s_warp[32];
    __shared__ float s_wmax[32];
    __shared__ int s_flag;
    if (threadIdx.x == 0) s_flag = 0;
    const int64_t row_off = (int64_t)frame * a.h * w + (int64_t)y * w;
    const uint32_t* img = a.image_u8 + row_off;
    for (int x = threadIdx.x; x < w; x += blockDim.x) simg[x] = img[x];
    build_sorted_points(a, eye, frame, y, c, px, clo, sidx, tmp, rnk, start, s_warp);

    // sorted x and the running maximum of segment ends (segment k = sorted point k -> its source successor)
    {
        const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
        const int per = (npts + kPolyThreads - 1) / kPolyThreads;
        const int b0 = tid * per, b1 = min(b0 + per, npts);
        float m = -INFINITY;
        for (int k = b0; k < b1; ++k) {
            int sp = sidx[k];
            sx[k] = px[sp];
            float x1 = (k < nsg) ? px[sp + 1] : -INFINITY;   // the last sorted point (sentinel 2W) starts no segment
            m = fmaxf(m, x1);
            reach[k] = m;
        }
        float inc = m;
        for (int o = 1; o < 32; o <<= 1) {
            float t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = fmaxf(inc, t);
        }
        if (lane == 31) s_wmax[wid] = inc;
        __syncthreads();
        float before = -INFINITY;   // max over all earlier threads
        for (int q = 0; q < wid; ++q) before = fmaxf(before, s_wmax[q]);
        float prev = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane > 0) before = fmaxf(before, prev);
        for (int k = b0; k < b1; ++k) reach[k] = fmaxf(reach[k], before);
        __syncthreads();
    }

    uint32_t* out = a.out[eye] + row_off;
    bool need_exact = false;
    for (int col = threadIdx.x; col < w; col += blockDim.x) {
        float color[3] = {0.5f, 0.5f, 0.5f};
        const int k0 = start[col + 1] - 1, k1 = start[col + 2] - 1;   // intervals k0..k1 (left point rank)
        for (int k = k0; k <= k1; ++k) {
            double pa = (double)sx[k], pb = (double)sx[k + 1];
            double from = fmax((double)col, pa) + kEps;
            double to = fmin((double)(col + 1), pb) - kEps;
            double sig = to - from;
            double ctr = from + 0.5 * sig;
            // active set: segments j <= k with x0 < ctr and not x1 < ctr
            int nact = 0, best = -1, only = -1, nbest = 0;
            double bestc = -kEps, best_ip = 0.0;
            for (int j = k; j >= 0 && !((double)reach[j] < ctr); --j) {
                int sp = sidx[j];
                float x0 = sx[j], x1 = px[sp + 1];
                if (!((double)x0 < ctr) || ((double)x1 < ctr)) continue;
                ++nact;
                only = sp;
                float den = x1 - x0;
                double ip = (ctr - (double)x0) / (double)den;
                double t0 = (1.0 - ip) * (double)pt_clo(sp, c, clo), t1 = ip * (double)pt_clo(sp + 1, c, clo);
                double cl = t0 + t1;
                if (0.0 < ip && ip < 1.0) {
                    if (bestc < cl) { bestc = cl; best = sp; best_ip = ip; nbest = 1; }
                    else if (bestc == cl) ++nbest;
                }
            }
            int sp;
            double ip;
            if (nact == 1) {
                sp = only;
                float den = px[sp + 1] - px[sp];
                ip = (ctr - (double)px[sp]) / (double)den;
            } else if (nact == 0) {
                continue;  // cannot happen inside the sentinels; the reference would read a stale slot
            } else {
                if (best < 0 || nbest > 1) { need_exact = true; best = (best < 0) ? only : best; }
                sp = best;
                ip = best_ip;
                if (best_ip == 0.0) {
                    float den = px[sp + 1] - px[sp];
                    ip = (ctr - (double)px[sp]) / (double)den;
                }
            }
            int cl = pt_col(sp, c), cr = pt_col(sp + 1, c);
            accumulate(color, simg[cl], simg[cr], cl == cr, ip, sig);
        }
        out[col] = pack_rgbx((int)color[0], (int)color[1], (int)color[2]);
    }
    if (need_exact) s_flag = 1;
    __syncthreads();
    if (threadIdx.x == 0) row_flags[((int64_t)frame * 2 + eye) * a.h + y] = s_flag;
}

static size_t exact_smem(int w, int sharp, int act_cap) {
    size_t npts = (size_t)(sharp ? 2 * w : w) + 2;
    size_t np2 = npts + (npts & 1);
    return npts * 4 + (size_t)w * 4 + (size_t)(w + 4) * 4 + np2 * 2 * 3 + (size_t)act_cap * 2;
}
static size_t fast_smem(int w, int sharp) {
    size_t npts = (size_t)(sharp ? 2 * w : w) + 2;
    size_t np2 = npts + (npts & 1);
    return npts * 4 * 3 + (size_t)w * 4 + (size_t)(w + 4) * 4 + (size_t)w * 4 + np2 * 2 * 3;
}

size_t polylines_scratch_bytes(int n, int h) { return ((size_t)n * 2 * h + 16) * sizeof(int); }

// scratch: [n*2*h] row flags + [1] status word.  force_exact = 1 skips the fast sweep (tests).
cudaError_t launch_polylines(const WarpArgs& a, cudaStream_t s) {
    const int sharp = a.fill == CS_FILL_POLYLINES_SHARP;
    const int w = a.w;
    if (a.scratch_bytes < polylines_scratch_bytes(a.n, a.h)) return cudaErrorInvalidValue;
    if ((sharp ? 2 * w : w) + 2 > 65535) return cudaErrorInvalidValue;
    int* flags = reinterpret_cast<int*>(a.scratch);
    int* status = flags + (size_t)a.n * 2 * a.h;
    double dmax = fmax(fabs(a.eye[0].div_px), fabs(a.eye[1].div_px));
    long long cap_ref = 5ll * (long long)dmax + 25;    // the reference's own list capacity, SIG:1947
    const size_t kMaxSmem = 227 * 1024;
    size_t base = exact_smem(w, sharp, 0);
    if (base + 64 > kMaxSmem) return cudaErrorInvalidValue;
    long long cap_fit = (long long)((kMaxSmem - base) / 2);
    int act_cap = (int)(cap_ref < cap_fit ? cap_ref : cap_fit);
    dim3 grid(a.h, a.n, 2);
    const bool force_exact = (a.flags & 1) != 0;
    const size_t fs = fast_smem(w, sharp);
    const bool use_fast = !force_exact && fs <= kMaxSmem;
    cudaError_t e;
    if (use_fast) {
        if (fs > 48 * 1024) cudaFuncSetAttribute(k_polylines, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs);
        prof_begin(K_POLY_FAST, s);
        k_polylines<<<grid, kPolyThreads, fs, s>>>(a, sharp, flags);
        prof_end(K_POLY_FAST, s);
        count_launch();
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    size_t es = exact_smem(w, sharp, act_cap);
    if (es > 48 * 1024) cudaFuncSetAttribute(k_polylines_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)es);
    prof_begin(K_POLY_EXACT, s);
    k_polylines_exact<<<grid, kPolyThreads, es, s>>>(a, sharp, act_cap, use_fast ? flags : nullptr, status);
    prof_end(K_POLY_EXACT, s);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cs

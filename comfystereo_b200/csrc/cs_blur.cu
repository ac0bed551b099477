// cs_blur.cu -- B1: the edge-aware directional depth blur (SIG:1171-1251, SIG:1131-1168).
//
// Two kernels per chunk of frames:
//   k_edge_dist   one CTA per image row: 3x3 Sobel-x (zero padded), the two threshold masks as
//                 bit words in shared memory (warp ballots), then per pixel the distance to the
//                 nearest mask bit left/right (clz/ffs over at most R/32+2 words), saturated at
//                 R+1 and stored as uint8 -> 2 B/px of scratch instead of two float weight maps.
//   k_blur_blend  persistent CTAs over 256-pixel row segments: vertical (2v+1)-tap mean of the
//                 LUT weights, bs-tap box mean of the depth row staged in shared memory, blend,
//                 per-frame min/max of both results, and (CPU techniques) the depth outputs.
//
// Summation order is fixed (ascending tap, fmaf) and identical to oracle/stereo_oracle.c, so
// CUDA and oracle agree bit for bit; torch's own conv2d order is unspecified.
//
// Bytes per pixel: k_edge_dist reads gray 4 B (x3 rows, L1/L2 hits) and writes 2 B;
// k_blur_blend reads gray 4 B + (2v+1) x 2 B of distances (L1/L2 hits), writes 8 B of blurred
// depth scratch and, for CPU techniques, 24 B of depth outputs (streamed, evict-first).
#include "cs_internal.cuh"

#include <atomic>
#include <cmath>

namespace cs {

struct BlurLut {
    float w[256];  // weight for distance k = 0..R+1
};

__device__ __forceinline__ float frame_scale(const FrameStats* st, int frame, int scale_mode,
                                             int group, int n) {
    if (scale_mode == 0) return 1.0f;
    float gm;
    if (scale_mode == 1) gm = ord2f(st[frame].gray_max);
    else {
        int g0 = (frame / group) * group, g1 = min(g0 + group, n);
        gm = -INFINITY;
        for (int f = g0; f < g1; ++f) gm = fmaxf(gm, ord2f(st[f].gray_max));
    }
    return (gm <= 1.0f) ? 255.0f : 1.0f;
}

__device__ __forceinline__ float scaled(float g, float scale) {
    return scale == 1.0f ? g : g * scale;
}

// ------------------------------------------------------------------------------------ dist
// NP = true: scipy.ndimage.sobel as the reference's numpy blur calls it (SIG:1377): borders reflected (d c b a | a b c d),
// derivative pass rounded to float32, then the [1 2 1] smoothing pass summed in float64 (centre first) and rounded.
__device__ __forceinline__ float sobel_np(const float* __restrict__ base, int h, int w, int y, int x) {
    const int xm = max(x - 1, 0), xp = min(x + 1, w - 1);
    const int ym = max(y - 1, 0), yp = min(y + 1, h - 1);
    const float* r0 = base + (int64_t)ym * w;
    const float* r1 = base + (int64_t)y * w;
    const float* r2 = base + (int64_t)yp * w;
    const float g0 = (float)((double)r0[xp] - (double)r0[xm]);
    const float g1 = (float)((double)r1[xp] - (double)r1[xm]);
    const float g2 = (float)((double)r2[xp] - (double)r2[xm]);
    double t = (double)g1 * 2.0;
    t = t + ((double)g0 + (double)g2) * 1.0;
    return (float)t;
}

template <bool NP>
__global__ void __launch_bounds__(256) k_edge_dist(const float* __restrict__ gray,
                                                   const FrameStats* __restrict__ st, int scale_mode,
                                                   int group, int n, int h, int w, float edge_div,
                                                   int radius, uint8_t* __restrict__ dist_l,
                                                   uint8_t* __restrict__ dist_r) {
    extern __shared__ uint32_t s_bits[];  // [2][nwords]
    const int y = blockIdx.x, frame = blockIdx.y;
    const int nwords = (w + 31) >> 5;
    uint32_t* bl = s_bits;
    uint32_t* br = s_bits + nwords;
    const float scale = frame_scale(st, frame, scale_mode, group, n);
    const float* base = gray + (int64_t)frame * h * w;
    const float* r0 = (y > 0) ? base + (int64_t)(y - 1) * w : nullptr;
    const float* r1 = base + (int64_t)y * w;
    const float* r2 = (y + 1 < h) ? base + (int64_t)(y + 1) * w : nullptr;

    const int wpad = nwords << 5;
    for (int x = threadIdx.x; x < wpad; x += blockDim.x) {
        bool ml = false, mr = false;
        if (NP) {
            if (x < w) {
                const float g = sobel_np(base, h, w, y, x);
                float e = fabsf(g) / edge_div;
                e = fminf(fmaxf(e, 0.0f), 1.0f);
                ml = (g > 0.0f) && (e > 0.5f);
                mr = (g < 0.0f) && (e > 0.5f);
            }
        } else if (x < w) {
            float g = 0.0f;
            const bool hasl = x > 0, hasr = x + 1 < w;
            if (r0) {
                float l = hasl ? scaled(r0[x - 1], scale) : 0.0f, r = hasr ? scaled(r0[x + 1], scale) : 0.0f;
                g = fmaf(-1.0f, l, g);
                g = fmaf(1.0f, r, g);
            }
            {
                float l = hasl ? scaled(r1[x - 1], scale) : 0.0f, r = hasr ? scaled(r1[x + 1], scale) : 0.0f;
                g = fmaf(-2.0f, l, g);
                g = fmaf(2.0f, r, g);
            }
            if (r2) {
                float l = hasl ? scaled(r2[x - 1], scale) : 0.0f, r = hasr ? scaled(r2[x + 1], scale) : 0.0f;
                g = fmaf(-1.0f, l, g);
                g = fmaf(1.0f, r, g);
            }
            float e = fabsf(g) / edge_div;          // IEEE float32 division, SIG:1217
            e = fminf(fmaxf(e, 0.0f), 1.0f);
            ml = (g > 0.0f) && (e > 0.5f);
            mr = (g < 0.0f) && (e > 0.5f);
        }
        uint32_t wl = __ballot_sync(0xffffffffu, ml), wr = __ballot_sync(0xffffffffu, mr);
        if ((threadIdx.x & 31) == 0) { bl[x >> 5] = wl; br[x >> 5] = wr; }
    }
    __syncthreads();

    // Per 32-pixel word: position of the nearest set bit in an earlier / later word (within reach of the radius),
    // so that the per-pixel lookup below is loop-free.
    const int far = radius + 1;
    int* prevpos = reinterpret_cast<int*>(s_bits + 2 * nwords);   // [2][nwords]
    int* nextpos = prevpos + 2 * nwords;                          // [2][nwords]
    const int kwords = (radius >> 5) + 1;
    for (int t = threadIdx.x; t < 2 * nwords; t += blockDim.x) {
        const int pass = t >= nwords, wi = pass ? t - nwords : t;
        const uint32_t* bits = pass ? br : bl;
        int pp = -(1 << 28), np = 1 << 28;
        for (int k = 1; k <= kwords && wi - k >= 0; ++k) {
            uint32_t q = bits[wi - k];
            if (q) { pp = ((wi - k) << 5) + 31 - __clz(q); break; }
        }
        for (int k = 1; k <= kwords && wi + k < nwords; ++k) {
            uint32_t q = bits[wi + k];
            if (q) { np = ((wi + k) << 5) + __ffs(q) - 1; break; }
        }
        prevpos[t] = pp;
        nextpos[t] = np;
    }
    __syncthreads();
    uint8_t* ol = dist_l + ((int64_t)frame * h + y) * w;
    uint8_t* orr = dist_r + ((int64_t)frame * h + y) * w;
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
        const int wi = x >> 5, b = x & 31;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const uint32_t word = (pass ? br : bl)[wi];
            const uint32_t ml = word & (0xffffffffu >> (31 - b));    // bits at or left of x
            const uint32_t mr = word & (0xffffffffu << b);           // bits at or right of x
            const int pl = ml ? (wi << 5) + 31 - __clz(ml) : prevpos[pass * nwords + wi];
            const int pn = mr ? (wi << 5) + __ffs(mr) - 1 : nextpos[pass * nwords + wi];
            const int d = min(min(x - pl, pn - x), far);
            (pass ? orr : ol)[x] = (uint8_t)d;
        }
    }
}

// The same for rows whose width is a multiple of 4 (every video format): four pixels per thread from 128-bit loads, the
// IEEE division of SIG:1217 replaced by the threshold it is equivalent to (division by a positive constant is monotone:
// fl(|g| / edge_div) > 0.5  <=>  |g| >= thr, thr found on the host), the two masks assembled from nibbles with three
// shuffles, four distances per 32-bit store.  ~45 instructions per pixel instead of ~160.
__global__ void __launch_bounds__(256) k_edge_dist4(const float* __restrict__ gray, const FrameStats* __restrict__ st,
                                                    int scale_mode, int group, int n, int h, int w, float thr, int radius,
                                                    uint8_t* __restrict__ dist_l, uint8_t* __restrict__ dist_r) {
    extern __shared__ uint32_t s_bits[];  // [2][nwords] masks, [2][nwords] prevpos, [2][nwords] nextpos
    const int y = blockIdx.x, frame = blockIdx.y;
    const int nwords = (w + 31) >> 5;
    uint32_t* bl = s_bits;
    uint32_t* br = s_bits + nwords;
    const float scale = frame_scale(st, frame, scale_mode, group, n);
    const float* base = gray + (int64_t)frame * h * w;
    const float* rows[3] = {(y > 0) ? base + (int64_t)(y - 1) * w : nullptr, base + (int64_t)y * w,
                            (y + 1 < h) ? base + (int64_t)(y + 1) * w : nullptr};
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    for (int xb = wid * 128; xb < w; xb += 4 * blockDim.x) {
        const int x4 = xb + 4 * lane;
        const bool act = x4 < w;
        float g[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (!rows[r]) continue;                         // zero padding above / below: no terms (block-uniform)
            float4 c = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            float l = 0.0f, rr = 0.0f;
            if (act) {
                c = *reinterpret_cast<const float4*>(rows[r] + x4);
                if (x4 > 0) l = rows[r][x4 - 1];
                if (x4 + 4 < w) rr = rows[r][x4 + 4];
                if (scale != 1.0f) { c.x *= scale; c.y *= scale; c.z *= scale; c.w *= scale; l *= scale; rr *= scale; }
            }
            const float k = (r == 1) ? 2.0f : 1.0f;
            const float v[6] = {l, c.x, c.y, c.z, c.w, rr};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                g[j] = fmaf(-k, v[j], g[j]);
                g[j] = fmaf(k, v[j + 2], g[j]);
            }
        }
        uint32_t nl = 0, nr = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool strong = fabsf(g[j]) >= thr;
            nl |= (uint32_t)(strong && g[j] > 0.0f) << j;
            nr |= (uint32_t)(strong && g[j] < 0.0f) << j;
        }
        const int sh = 4 * (lane & 7);
        uint32_t vl = nl << sh, vr = nr << sh;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            vl |= __shfl_xor_sync(0xffffffffu, vl, o);
            vr |= __shfl_xor_sync(0xffffffffu, vr, o);
        }
        const int wi = (xb >> 5) + (lane >> 3);
        if ((lane & 7) == 0 && wi < nwords) { bl[wi] = vl; br[wi] = vr; }
    }
    __syncthreads();

    const int far = radius + 1;
    int* prevpos = reinterpret_cast<int*>(s_bits + 2 * nwords);   // [2][nwords]
    int* nextpos = prevpos + 2 * nwords;                          // [2][nwords]
    const int kwords = (radius >> 5) + 1;
    for (int t = threadIdx.x; t < 2 * nwords; t += blockDim.x) {
        const int pass = t >= nwords, wi = pass ? t - nwords : t;
        const uint32_t* bits = pass ? br : bl;
        int pp = -(1 << 28), np = 1 << 28;
        for (int k = 1; k <= kwords && wi - k >= 0; ++k) {
            uint32_t q = bits[wi - k];
            if (q) { pp = ((wi - k) << 5) + 31 - __clz(q); break; }
        }
        for (int k = 1; k <= kwords && wi + k < nwords; ++k) {
            uint32_t q = bits[wi + k];
            if (q) { np = ((wi + k) << 5) + __ffs(q) - 1; break; }
        }
        prevpos[t] = pp;
        nextpos[t] = np;
    }
    __syncthreads();
    uint32_t* ol = reinterpret_cast<uint32_t*>(dist_l + ((int64_t)frame * h + y) * w);
    uint32_t* orr = reinterpret_cast<uint32_t*>(dist_r + ((int64_t)frame * h + y) * w);
    for (int x4 = 4 * threadIdx.x; x4 < w; x4 += 4 * blockDim.x) {
        const int wi = x4 >> 5, b0 = x4 & 31, wbase = wi << 5;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const uint32_t word = (pass ? br : bl)[wi];
            const int pp = prevpos[pass * nwords + wi], np = nextpos[pass * nwords + wi];
            uint32_t pack = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = b0 + j, x = x4 + j;
                const uint32_t ml = word & (0xffffffffu >> (31 - b));
                const uint32_t mr = word & (0xffffffffu << b);
                const int pl = ml ? wbase + 31 - __clz(ml) : pp;
                const int pn = mr ? wbase + __ffs(mr) - 1 : np;
                pack |= (uint32_t)min(min(x - pl, pn - x), far) << (8 * j);
            }
            (pass ? orr : ol)[x4 >> 2] = pack;
        }
    }
}

// trunc(t) mod 256 as numpy's float32 -> int64 -> uint8 cast chain does it (wrap quirk Q1).  Beyond the int32 range a
// float32 is a multiple of 256 (24-bit significand), and NaN / overflow convert to INT64_MIN on x86: all of them give 0.
__device__ __forceinline__ int wrap_u8(float t) {
    return (fabsf(t) < 2147483520.0f) ? (__float2int_rz(t) & 255) : 0;
}

// ------------------------------------------------------------------------------------ blend
constexpr int kSeg = 256;  // pixels per row segment = threads per CTA
constexpr int kTileY = 8;  // rows per work item
constexpr int kRowPad = 16;

// Work item = kTileY rows x kSeg columns:
//   pass 1  thread = column: loads the kTileY + 2v edge distances of its column, looks their weights up and feeds all
//           kTileY vertical sums from registers (each weight is used by up to 2v+1 accumulators; every accumulator still
//           sees its taps in ascending order); the sums go to shared memory for pass 2
//   pass 2  thread = (row, 8 consecutive columns): the bs-tap horizontal box for 8 pixels from a register window over
//           the depth rows staged with their halo (15 loads per 64 fused multiply-adds), blend, min/max, depth outputs
// Rows outside the image hold weight 0: fmaf(0, wv, acc) == acc, which is the reference's zero padding.
// Tap order (ascending, fmaf) is the oracle's, so the result is bit-identical to it.
template <int V>   // vertical smoothing radius (0..15): compile-time so that the column walk unrolls without predicates
__global__ void __launch_bounds__(kSeg, 4) k_blur_blend(
    const float* __restrict__ gray, FrameStats* __restrict__ st, int scale_mode, int group, int n,
    int h, int w, int bs, int radius, const __grid_constant__ BlurLut lut,
    const uint8_t* __restrict__ dist_l, const uint8_t* __restrict__ dist_r, float* __restrict__ blur_l,
    float* __restrict__ blur_r, float* __restrict__ out_l, float* __restrict__ out_r, int items_per_frame,
    int segs_per_row) {
    extern __shared__ __align__(16) float s_dyn[];
    constexpr int v = V;
    constexpr int wrows = kTileY + 2 * V;
    const int rw = (kSeg + bs + kRowPad + 3) & ~3;     // row pitch, multiple of 4 floats
    float* s_wl = s_dyn;                         // [kTileY][kSeg] vertical sums of the weights, then the depth outputs
    float* s_wr = s_wl + kTileY * kSeg;          // [kTileY][kSeg]
    float* s_row = s_wr + kTileY * kSeg;         // [kTileY][rw]
    __shared__ float s_red[4][kSeg / 32];
    __shared__ float s_lut[257];
    __shared__ float s_q[256];                   // k / 255.0f, the depth outputs' dequantisation (GS:365)
    const int tid = threadIdx.x;
    s_lut[tid] = lut.w[tid];
    s_q[tid] = kQ255[tid];
    if (tid == 0) s_lut[256] = 0.0f;
    const int frame = blockIdx.y;
    const float scale = frame_scale(st, frame, scale_mode, group, n);
    const float wv = 1.0f / (float)(2 * v + 1);
    const float wb = 1.0f / (float)bs;
    const int lo = bs / 2;
    const uint64_t pol = policy_evict_first();
    const float* base = gray + (int64_t)frame * h * w;
    const uint8_t* dl = dist_l + (int64_t)frame * h * w;
    const uint8_t* dr = dist_r + (int64_t)frame * h * w;
    float* bl = blur_l + (int64_t)frame * h * w;
    float* br = blur_r + (int64_t)frame * h * w;
    const bool vec_out = out_l && (w % 4 == 0) && ((uintptr_t)out_l % 16 == 0) && ((uintptr_t)out_r % 16 == 0);
    const bool vec_blur = (w % 4 == 0) && ((uintptr_t)blur_l % 16 == 0) && ((uintptr_t)blur_r % 16 == 0);
    const int pr = tid >> 5, pg = tid & 31;      // pass 2: row of the tile, group of 8 columns
    __syncthreads();

    float mnl = INFINITY, mxl = -INFINITY, mnr = INFINITY, mxr = -INFINITY;
    for (int item = blockIdx.x; item < items_per_frame; item += gridDim.x) {
        const int ty = item / segs_per_row, x0 = (item - ty * segs_per_row) * kSeg;
        const int y0 = ty * kTileY;
        const int x = x0 + tid;
        // ---- pass 1: vertical sums of column tid for all kTileY rows, straight from registers: the thread that
        // loads a column's distances (rows y0 - v .. y0 + kTileY - 1 + v, all byte loads first) is the one that sums them
        float al[kTileY], ar[kTileY];
        {
            uint32_t da[wrows], db[wrows];
            const bool xin = x < w;
            if (y0 - v >= 0 && y0 + kTileY + v <= h) {      // interior tile (CTA-uniform): no per-row tests
                const uint8_t* pa = dl + (int64_t)(y0 - v) * w + (xin ? x : 0);
                const uint8_t* pb = dr + (int64_t)(y0 - v) * w + (xin ? x : 0);
#pragma unroll
                for (int rr = 0; rr < wrows; ++rr) {
                    da[rr] = *pa; db[rr] = *pb;
                    pa += w; pb += w;
                }
                if (!xin) {
#pragma unroll
                    for (int rr = 0; rr < wrows; ++rr) { da[rr] = 256u; db[rr] = 256u; }
                }
            } else {
#pragma unroll
                for (int rr = 0; rr < wrows; ++rr) {
                    const int yy = y0 - v + rr;
                    const bool in = xin && yy >= 0 && yy < h;
                    const int off = in ? yy * w + x : 0;
                    da[rr] = in ? (uint32_t)dl[off] : 256u;     // s_lut[256] = 0: rows outside the image weigh nothing
                    db[rr] = in ? (uint32_t)dr[off] : 256u;
                }
            }
#pragma unroll
            for (int r = 0; r < kTileY; ++r) { al[r] = 0.0f; ar[r] = 0.0f; }
#pragma unroll
            for (int t = 0; t < wrows; ++t) {
                const float a = s_lut[da[t]], b = s_lut[db[t]];
                if (v > 0) {
#pragma unroll
                    for (int r = 0; r < kTileY; ++r) {
                        const int tap = t - r;
                        if (tap >= 0 && tap <= 2 * v) { al[r] = fmaf(a, wv, al[r]); ar[r] = fmaf(b, wv, ar[r]); }
                    }
                } else {
                    al[t] = a; ar[t] = b;
                }
            }
        }
        // ---- stage depth[y][x0 - lo .. ) for the tile's rows (zeros outside the row)
        {
            const int xx = x0 - lo + tid;                 // body: thread = column, all kTileY rows
            const bool xin2 = xx >= 0 && xx < w;
            float tmp[kTileY];
#pragma unroll
            for (int r = 0; r < kTileY; ++r) {
                const int y = y0 + r;
                tmp[r] = (xin2 && y < h) ? base[(int64_t)y * w + xx] : 0.0f;
            }
#pragma unroll
            for (int r = 0; r < kTileY; ++r) s_row[r * rw + tid] = scaled(tmp[r], scale);
            // halo (the bs - 1 columns past the segment that the box reaches): warp = row, lane = column
            const int y = y0 + pr;
            for (int i = kSeg + pg; i < kSeg + bs - 1; i += 32) {
                const int xh = x0 - lo + i;
                s_row[pr * rw + i] = (xh >= 0 && xh < w && y < h) ? scaled(base[(int64_t)y * w + xh], scale) : 0.0f;
            }
        }
#pragma unroll
        for (int r = 0; r < kTileY; ++r) { s_wl[r * kSeg + tid] = al[r]; s_wr[r * kSeg + tid] = ar[r]; }
        __syncthreads();
        // ---- pass 2: horizontal box for 8 consecutive pixels of row pr
        {
            const int y = y0 + pr;
            const int c0 = pg * 8;
            const float* prow = s_row + pr * rw + c0;
            float bsum[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bsum[j] = 0.0f;
            for (int k0 = 0; k0 < bs; k0 += 8) {
                float win[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t4 = *reinterpret_cast<const float4*>(prow + k0 + 4 * q);
                    win[4 * q] = t4.x; win[4 * q + 1] = t4.y; win[4 * q + 2] = t4.z; win[4 * q + 3] = t4.w;
                }
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    if (k0 + kk < bs) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) bsum[j] = fmaf(win[j + kk], wb, bsum[j]);
                    }
                }
            }
            float vl[8], vr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = prow[lo + j];
                const float wl = s_wl[pr * kSeg + c0 + j], wr = s_wr[pr * kSeg + c0 + j];
                float t0 = wl * bsum[j], t1 = (1.0f - wl) * d;
                vl[j] = t0 + t1;
                t0 = wr * bsum[j]; t1 = (1.0f - wr) * d;
                vr[j] = t0 + t1;
            }
            const int xg = x0 + c0;
            if (y < h) {
                if (vec_blur && xg + 8 <= w) {
                    float4* pl4 = reinterpret_cast<float4*>(bl + (int64_t)y * w + xg);
                    float4* pr4 = reinterpret_cast<float4*>(br + (int64_t)y * w + xg);
                    pl4[0] = make_float4(vl[0], vl[1], vl[2], vl[3]); pl4[1] = make_float4(vl[4], vl[5], vl[6], vl[7]);
                    pr4[0] = make_float4(vr[0], vr[1], vr[2], vr[3]); pr4[1] = make_float4(vr[4], vr[5], vr[6], vr[7]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (xg + j < w) { bl[(int64_t)y * w + xg + j] = vl[j]; br[(int64_t)y * w + xg + j] = vr[j]; }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (xg + j < w) {
                        mnl = fminf(mnl, vl[j]); mxl = fmaxf(mxl, vl[j]);
                        mnr = fminf(mnr, vr[j]); mxr = fmaxf(mxr, vr[j]);
                    }
            }
            if (out_l) {  // CPU-technique depth outputs: u8 = trunc(v*255) mod 256, /255 (Q1); staged in place
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s_wl[pr * kSeg + c0 + j] = s_q[wrap_u8(vl[j] * 255.0f)];
                    s_wr[pr * kSeg + c0 + j] = s_q[wrap_u8(vr[j] * 255.0f)];
                }
            }
        }
        if (out_l) {
            __syncthreads();
            const int npx = min(kSeg, w - x0);
            const int rows = min(kTileY, h - y0);
            float* pl = out_l + (((int64_t)frame * h + y0) * w + x0) * 3;
            float* pr_ = out_r + (((int64_t)frame * h + y0) * w + x0) * 3;
            if (vec_out) {
                // float4 m of a row holds floats 4m..4m+3: channel k.. of pixel p0, then pixel p0 + 1 (npx % 4 == 0 here,
                // so 3 * npx / 4 <= 192 float4 per row: one per thread)
                const int nvec = (npx * 3) >> 2;
                if (tid < nvec) {
                    const int f0 = 4 * tid;
                    const int p0 = f0 / 3, k = f0 - 3 * p0;
                    const int p1 = p0 + 1 < kSeg ? p0 + 1 : p0;
                    float4* ql = reinterpret_cast<float4*>(pl) + tid;
                    float4* qr = reinterpret_cast<float4*>(pr_) + tid;
                    const int pitch4 = (w * 3) >> 2;      // w % 4 == 0
                    for (int r = 0; r < rows; ++r) {
                        const float a0 = s_wl[r * kSeg + p0], a1 = s_wl[r * kSeg + p1];
                        const float c0 = s_wr[r * kSeg + p0], c1 = s_wr[r * kSeg + p1];
                        st_stream_f4(ql, make_float4(a0, (k < 2) ? a0 : a1, (k < 1) ? a0 : a1, a1), pol);
                        st_stream_f4(qr, make_float4(c0, (k < 2) ? c0 : c1, (k < 1) ? c0 : c1, c1), pol);
                        ql += pitch4; qr += pitch4;
                    }
                }
            } else {
                for (int r = 0; r < rows; ++r) {
                    const float* sl = s_wl + r * kSeg;
                    const float* sr = s_wr + r * kSeg;
                    for (int f = tid; f < npx * 3; f += kSeg) { pl[f] = sl[f / 3]; pr_[f] = sr[f / 3]; }
                    pl += (int64_t)w * 3; pr_ += (int64_t)w * 3;
                }
            }
        }
        __syncthreads();
    }
    mnl = warp_min(mnl); mxl = warp_max(mxl); mnr = warp_min(mnr); mxr = warp_max(mxr);
    const int wid = tid >> 5, lane = tid & 31;
    if (lane == 0) { s_red[0][wid] = mnl; s_red[1][wid] = mxl; s_red[2][wid] = mnr; s_red[3][wid] = mxr; }
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < kSeg / 32; ++k) {
            mnl = fminf(mnl, s_red[0][k]); mxl = fmaxf(mxl, s_red[1][k]);
            mnr = fminf(mnr, s_red[2][k]); mxr = fmaxf(mxr, s_red[3][k]);
        }
        if (mnl <= mxl) {
            atomicMin(&st[frame].l_min, f2ord(mnl)); atomicMax(&st[frame].l_max, f2ord(mxl));
            atomicMin(&st[frame].r_min, f2ord(mnr)); atomicMax(&st[frame].r_max, f2ord(mxr));
        }
    }
}

// The numpy / scipy blur (SIG:1397-1419), one thread per pixel.  Not on the node's path (non-tensor inputs of
// create_stereoimages only), so it is written for fidelity, not speed:
//   weights  LUT(dist), then scipy's convolve1d over rows with nearest borders: float64, centre first, pairs from the
//            outside in, rounded to float32, clipped to [0, 1]
//   box      convolve1d along the row, nearest borders; an even box reaches bs/2 - 1 samples left and bs/2 right of the
//            pixel; odd boxes are summed in the paired order, even ones last tap first and then ascending (NI_Correlate1D)
__global__ void __launch_bounds__(256) k_blur_blend_np(const float* __restrict__ gray, FrameStats* __restrict__ st, int h, int w,
                                                       int bs, int v, const __grid_constant__ BlurLut lut,
                                                       const uint8_t* __restrict__ dist_l, const uint8_t* __restrict__ dist_r,
                                                       float* __restrict__ blur_l, float* __restrict__ blur_r) {
    const int frame = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    float vl = 0.0f, vr = 0.0f;
    const bool in = x < w;
    if (in) {
        const float* base = gray + (int64_t)frame * h * w;
        const uint8_t* dl = dist_l + (int64_t)frame * h * w;
        const uint8_t* dr = dist_r + (int64_t)frame * h * w;
        float wl, wr;
        if (v > 0) {
            const double k = 1.0 / (double)(2 * v + 1);
            double al = (double)lut.w[dl[(int64_t)y * w + x]] * k, ar = (double)lut.w[dr[(int64_t)y * w + x]] * k;
            for (int q = v; q >= 1; --q) {
                const int ya = max(y - q, 0), yb = min(y + q, h - 1);
                al = al + ((double)lut.w[dl[(int64_t)ya * w + x]] + (double)lut.w[dl[(int64_t)yb * w + x]]) * k;
                ar = ar + ((double)lut.w[dr[(int64_t)ya * w + x]] + (double)lut.w[dr[(int64_t)yb * w + x]]) * k;
            }
            wl = fminf(fmaxf((float)al, 0.0f), 1.0f);
            wr = fminf(fmaxf((float)ar, 0.0f), 1.0f);
        } else {
            wl = lut.w[dl[(int64_t)y * w + x]];
            wr = lut.w[dr[(int64_t)y * w + x]];
        }
        const float* row = base + (int64_t)y * w;
        const double kb = 1.0 / (double)bs;
        double acc;
        if (bs & 1) {
            const int half = bs / 2;
            acc = (double)row[x] * kb;
            for (int q = half; q >= 1; --q)
                acc = acc + ((double)row[max(x - q, 0)] + (double)row[min(x + q, w - 1)]) * kb;
        } else {
            const int lo = bs / 2 - 1;     // taps x - lo .. x + bs / 2
            acc = (double)row[min(x + bs / 2, w - 1)] * kb;
            for (int i = 0; i < bs - 1; ++i)
                acc = acc + (double)row[min(max(x - lo + i, 0), w - 1)] * kb;
        }
        const float b = (float)acc, d = row[x];
        float t0 = wl * b, t1 = (1.0f - wl) * d;
        vl = t0 + t1;
        t0 = wr * b; t1 = (1.0f - wr) * d;
        vr = t0 + t1;
        blur_l[((int64_t)frame * h + y) * w + x] = vl;
        blur_r[((int64_t)frame * h + y) * w + x] = vr;
    }
    float mnl = in ? vl : INFINITY, mxl = in ? vl : -INFINITY, mnr = in ? vr : INFINITY, mxr = in ? vr : -INFINITY;
    mnl = warp_min(mnl); mxl = warp_max(mxl); mnr = warp_min(mnr); mxr = warp_max(mxr);
    if ((threadIdx.x & 31) == 0 && mnl <= mxl) {
        atomicMin(&st[frame].l_min, f2ord(mnl)); atomicMax(&st[frame].l_max, f2ord(mxl));
        atomicMin(&st[frame].r_min, f2ord(mnr)); atomicMax(&st[frame].r_max, f2ord(mxr));
    }
}

// weight(dist) = clamp(1 - dist/R, 0, 1) ** falloff, float32 (SIG:1168).  Built on the host once
// per call (<= 256 entries): torch.pow special-cases exponents 1, 2, 3, 0.5; otherwise powf.
template <int V, typename... Args>
static void launch_blend_one(dim3 grid, size_t smem, cudaStream_t s, Args... args) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_blur_blend<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_blur_blend<V><<<grid, kSeg, smem, s>>>(args...);
}
template <typename... Args>
static void launch_blend_v(int v, dim3 grid, size_t smem, cudaStream_t s, Args... args) {
    switch (v) {
        case 0: launch_blend_one<0>(grid, smem, s, args...); break;
        case 1: launch_blend_one<1>(grid, smem, s, args...); break;
        case 2: launch_blend_one<2>(grid, smem, s, args...); break;
        case 3: launch_blend_one<3>(grid, smem, s, args...); break;
        case 4: launch_blend_one<4>(grid, smem, s, args...); break;
        case 5: launch_blend_one<5>(grid, smem, s, args...); break;
        case 6: launch_blend_one<6>(grid, smem, s, args...); break;
        case 7: launch_blend_one<7>(grid, smem, s, args...); break;
        case 8: launch_blend_one<8>(grid, smem, s, args...); break;
        case 9: launch_blend_one<9>(grid, smem, s, args...); break;
        case 10: launch_blend_one<10>(grid, smem, s, args...); break;
        case 11: launch_blend_one<11>(grid, smem, s, args...); break;
        case 12: launch_blend_one<12>(grid, smem, s, args...); break;
        case 13: launch_blend_one<13>(grid, smem, s, args...); break;
        case 14: launch_blend_one<14>(grid, smem, s, args...); break;
        default: launch_blend_one<15>(grid, smem, s, args...); break;
    }
}

static void build_lut(BlurLut& lut, int radius, float falloff, bool numpy_pow = false) {
    for (int k = 0; k < 256; ++k) {
        float q = (float)k / (float)radius;   // radius 0 -> NaN at k = 0, inf elsewhere (quirk Q11)
        float wgt = 1.0f - q;
        if (wgt == wgt) {
            if (wgt < 0.0f) wgt = 0.0f;
            if (wgt > 1.0f) wgt = 1.0f;
            if (falloff == 1.0f) {}
            else if (falloff == 2.0f) wgt = wgt * wgt;
            else if (falloff == 3.0f && !numpy_pow) wgt = (wgt * wgt) * wgt;   // torch.pow only; numpy calls powf
            else if (falloff == 0.5f) wgt = sqrtf(wgt);
            else wgt = powf(wgt, falloff);
        }
        lut.w[k] = wgt;
    }
}

static std::atomic<int> g_blur_test_flags{0};   // bit 0: the scalar k_edge_dist even where the 4-pixel form applies (tests)
void set_blur_test_flags(int flags) { g_blur_test_flags.store(flags); }

cudaError_t launch_blur(const float* gray, FrameStats* stats, int scale_mode, int group, int n, int h,
                        int w, const cs_params& p, float* blur_l, float* blur_r, uint8_t* dist,
                        float* depth_l_out, float* depth_r_out, cudaStream_t s) {
    const int bs = p.blur_box, radius = p.blur_radius, v = p.blur_vert_smooth;
    if (bs < 1 || radius < 0 || radius > kMaxBlurRadius || v < 0 || v > 15) return cudaErrorInvalidValue;
    uint8_t* dist_l = dist;
    uint8_t* dist_r = dist + (int64_t)n * h * w;
    const float edge_div = (float)(10.0 * p.blur_edge_threshold);  // python float 10*thr -> float32 scalar
    const int nwords = (w + 31) >> 5;
    const bool np_flavor = p.blur_flavor == 1;
    if (np_flavor && (scale_mode != 0 || depth_l_out || depth_r_out)) return cudaErrorInvalidValue;   // function-level only
    prof_begin(K_EDGE_DIST, s);
    if (np_flavor)
        k_edge_dist<true><<<dim3(h, n), 256, 6 * nwords * sizeof(uint32_t), s>>>(
            gray, stats, scale_mode, group < 1 ? 1 : group, n, h, w, edge_div, radius, dist_l, dist_r);
    else if (w % 4 == 0 && edge_div > 0.0f && std::isfinite(edge_div) && !(g_blur_test_flags.load() & 1)) {
        // smallest float32 whose quotient by edge_div rounds above 0.5 (the division is monotone in its numerator)
        float thr = 0.5f * edge_div;
        while (thr > 0.0f && thr / edge_div > 0.5f) thr = nextafterf(thr, 0.0f);
        while (!(thr / edge_div > 0.5f)) thr = nextafterf(thr, INFINITY);
        // (64- to 256-thread CTAs measure the same; wider ones wait longer at the two barriers)
        k_edge_dist4<<<dim3(h, n), 256, 6 * nwords * sizeof(uint32_t), s>>>(
            gray, stats, scale_mode, group < 1 ? 1 : group, n, h, w, thr, radius, dist_l, dist_r);
    } else
        k_edge_dist<false><<<dim3(h, n), 256, 6 * nwords * sizeof(uint32_t), s>>>(
            gray, stats, scale_mode, group < 1 ? 1 : group, n, h, w, edge_div, radius, dist_l, dist_r);
    prof_end(K_EDGE_DIST, s);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;

    BlurLut lut;
    build_lut(lut, radius, (float)p.blur_falloff, np_flavor);
    if (np_flavor) {
        prof_begin(K_BLUR_BLEND, s);
        k_blur_blend_np<<<dim3((w + 255) / 256, h, n), 256, 0, s>>>(gray, stats, h, w, bs, v, lut, dist_l, dist_r, blur_l, blur_r);
        prof_end(K_BLUR_BLEND, s);
        count_launch();
        return cudaGetLastError();
    }
    const int segs = (w + kSeg - 1) / kSeg;
    const int tiles_y = (h + kTileY - 1) / kTileY;
    const int items = segs * tiles_y;
    // few CTAs per frame so the 4 min/max atomics per CTA stay cheap; >= 4 waves in total
    int per_frame = (sm_count() * 5 * 4 + n - 1) / n;
    if (per_frame > items) per_frame = items;
    if (per_frame < 1) per_frame = 1;
    const int rw = (kSeg + bs + kRowPad + 3) & ~3;
    size_t smem = ((size_t)2 * kTileY * kSeg + (size_t)kTileY * rw) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    prof_begin(K_BLUR_BLEND, s);
    launch_blend_v(v, dim3(per_frame, n), smem, s, gray, stats, scale_mode, group < 1 ? 1 : group, n, h, w, bs, radius, lut,
                   dist_l, dist_r, blur_l, blur_r, depth_l_out, depth_r_out, items, segs);
    prof_end(K_BLUR_BLEND, s);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cs

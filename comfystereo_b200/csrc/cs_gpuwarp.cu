// cs_gpuwarp.cu -- G1 + C2 + M2: "GPU Warp (Fast)" = forward_warp_gpu (SIG:277-450) for both
// eyes, composed straight into the final stereo tensor (SIG:1093-1122) with the unfilled mask
// (SIG:1073-1090).
//
// One CTA per (row, frame); the row's whole state (normalised depth, offsets, destination x,
// z-buffer, source map) stays in shared memory for all 8 scatter rounds, the gap interpolation
// and the bilinear resample, so HBM sees one read of the depth row and image row and one write
// of the output row per eye -- the reference makes ~60 full-tensor passes.
//
// Scatter semantics (Q8): in round k every pixel pair writes BOTH buffers at its clamped column
// (its candidate if it is valid and passes the z-test against the state before the round, else the
// old value), and on the CPU the last -- highest-index -- writer of a column wins.  So per round
// and column only the highest pair index targeting it matters: an atomicMax on the pair index
// selects it, then one thread per column replays that single pair.  Everything is float32, one
// rounded operation at a time (the library is built with -fmad=false).
#include "cs_internal.cuh"

namespace cs {

__device__ __forceinline__ float gw_group_max(const FrameStats* st, int frame, int group, int n, int which) {
    int g0 = (frame / group) * group, g1 = min(g0 + group, n);
    float m = -INFINITY;
    for (int f = g0; f < g1; ++f) {
        int o = which == 0 ? st[f].gray_max : (which == 1 ? st[f].l_max : st[f].r_max);
        m = fmaxf(m, ord2f(o));
    }
    return m;
}

// torch.linspace(-1, 1, n) float32 (CPU kernel: first half counts up, second half counts down)
__device__ __forceinline__ float linspace_m1_1(int i, int n) {
    if (n == 1) return -1.0f;
    float step = (1.0f - (-1.0f)) / (float)(n - 1);
    int half = n / 2;
    if (i < half) return -1.0f + step * (float)i;
    return 1.0f - step * (float)(n - i - 1);
}

__device__ __forceinline__ float gw_pow(float ab, float expo) {  // torch.pow special cases
    if (expo == 2.0f) return ab * ab;
    if (expo == 1.0f) return ab;
    if (expo == 0.5f) return sqrtf(ab);
    if (expo == 3.0f) return (ab * ab) * ab;
    return powf(ab, expo);
}

// depth scale chain of one (frame, eye): x255 if the sub-batch max <= 1 (SIG:1045), blur, /255 if any frame of the
// sub-batch > 1 (SIG:314-316, 487-489), then the frame's own min/max normalisation (SIG:317-324, 490-497)
struct DepthNorm { float pre, dmin, rng; bool div255, flat; };
__device__ __forceinline__ DepthNorm gw_depth_norm(const GpuWarpArgs& a, int frame, int eye) {
    const FrameStats st = a.stats[frame];
    DepthNorm nm;
    nm.pre = 1.0f;
    int omin, omax;
    if (a.use_blur_stats) {
        omin = eye ? st.r_min : st.l_min; omax = eye ? st.r_max : st.l_max;
        nm.div255 = gw_group_max(a.stats, frame, a.group, a.n, eye ? 2 : 1) > 1.0f;
    } else {
        omin = st.gray_min; omax = st.gray_max;
        float gm = gw_group_max(a.stats, frame, a.group, a.n, 0);
        nm.pre = (a.prescale && gm <= 1.0f) ? 255.0f : 1.0f;
        nm.div255 = gm * nm.pre > 1.0f;
    }
    float dmin = ord2f(omin), dmax = ord2f(omax);
    if (nm.pre != 1.0f) { dmin = dmin * nm.pre; dmax = dmax * nm.pre; }
    if (nm.div255) { dmin = dmin / 255.0f; dmax = dmax / 255.0f; }
    const float range = dmax - dmin;
    nm.flat = !(range > 1e-6f);
    nm.rng = range < 1e-6f ? 1e-6f : range;
    nm.dmin = dmin;
    return nm;
}

// pixel offset of one depth sample (SIG:326-331, 499-504); *nd = the normalised depth
__device__ __forceinline__ float gw_offset(const DepthNorm& nm, float dv, float conv, float expo, float div_px,
                                           float sep_px, float* nd) {
    if (nm.pre != 1.0f) dv = dv * nm.pre;
    if (nm.div255) dv = dv / 255.0f;
    const float n = nm.flat ? 0.0f : (dv - nm.dmin) / nm.rng;
    *nd = n;
    const float sh = n - conv;
    const float sg = (sh > 0.0f) ? 1.0f : ((sh < 0.0f) ? -1.0f : 0.0f);
    const float od = sg * gw_pow(fabsf(sh), expo);
    const float m = od * div_px;
    return m + sep_px;
}

// One (row, frame) of the scatter warp; `smem_f` is the row's state: shared memory, or -- for rows too wide for it -- the
// CTA's slice of a global scratch buffer (same code: the atomics and barriers work on either).
__device__ __forceinline__ void gpuwarp_row(const GpuWarpArgs& a, const int y, const int frame, float* smem_f) {
    const int w = a.w, h = a.h;
    const int nwords = (w + 31) >> 5;
    float* ndv = smem_f;
    float* po = ndv + w;
    float* src = po + w;        // (dest[x] = x + po[x] is recomputed where it is needed: 21 instead of 25 bytes per column
                                //  lets five CTAs instead of four share an SM at 1080p)
    float* zb = src + w;
    int* win = reinterpret_cast<int*>(zb + w);
    uint32_t* fbits = reinterpret_cast<uint32_t*>(win + w + 16);   // filled (src >= 0) bitmap
    uint32_t* ubits = fbits + nwords;                         // unfilled in either eye (mask)
    unsigned char* vm = reinterpret_cast<unsigned char*>(ubits + nwords);   // [w] per pair: rounds in which it can be valid
    __shared__ int s_last;

    for (int i = threadIdx.x; i < nwords; i += blockDim.x) ubits[i] = 0u;

    // output geometry (SIG:1093-1118)
    const int mode = a.mode;
    int wo = w, ho = h;
    if (mode == CS_MODE_LEFT_RIGHT || mode == CS_MODE_RIGHT_LEFT) wo = 2 * w;
    if (mode == CS_MODE_TOP_BOTTOM || mode == CS_MODE_BOTTOM_TOP) ho = 2 * h;
    float* outf = a.stereo + (int64_t)frame * ho * wo * 3;
    const float* img = a.image + (int64_t)frame * h * w * 3;

    // bilinear row weights, identical for both eyes (grid_sample, align_corners=True, border)
    const float sx = (float)(w - 1) / 2.0f, sy = (float)(h - 1) / 2.0f;
    float iy = (linspace_m1_1(y, h) + 1.0f) * sy;
    iy = fminf(fmaxf(iy, 0.0f), (float)(h - 1));
    const float fy = floorf(iy);
    const int y0 = (int)fy, y1 = y0 + 1;
    const float wy1 = iy - fy, wy0 = 1.0f - wy1;
    const float* r0 = img + (int64_t)y0 * w * 3;
    const float* r1 = (y1 < h) ? img + (int64_t)y1 * w * 3 : nullptr;

    for (int eye = 0; eye < 2; ++eye) {
        // where this eye's pixels go and which channels it contributes
        int oy = y, ox = 0, ch_lo = 0, ch_hi = 3;
        bool emit = true;
        switch (mode) {
            case CS_MODE_LEFT_RIGHT: ox = eye ? w : 0; break;
            case CS_MODE_RIGHT_LEFT: ox = eye ? 0 : w; break;
            case CS_MODE_TOP_BOTTOM: oy = eye ? y + h : y; break;
            case CS_MODE_BOTTOM_TOP: oy = eye ? y : y + h; break;
            case CS_MODE_RED_CYAN: if (eye == 0) { ch_lo = 0; ch_hi = 1; } else { ch_lo = 1; ch_hi = 3; } break;
            case CS_MODE_CYAN_RED: if (eye == 1) { ch_lo = 0; ch_hi = 1; } else { ch_lo = 1; ch_hi = 3; } break;
            case CS_MODE_LEFT_ONLY: emit = (eye == 0); break;
            default: emit = (eye == 1); break;
        }
        float* orow = outf + ((int64_t)oy * wo + ox) * 3;

        if (a.eye[eye].passthrough) {  // SIG:1076 / 1083: the eye is the input image, mask all false
            if (emit)
                for (int x = threadIdx.x; x < w; x += blockDim.x)
                    for (int ch = ch_lo; ch < ch_hi; ++ch) orow[x * 3 + ch] = img[((int64_t)y * w + x) * 3 + ch];
            continue;
        }
        const float div_px = (float)a.eye[eye].div_px, sep_px = (float)a.eye[eye].sep_px;
        const DepthNorm nm = gw_depth_norm(a, frame, eye);
        const float* dep = a.depth[eye] + (int64_t)frame * h * w + (int64_t)y * w;

        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            float n;
            const float p = gw_offset(nm, dep[x], a.conv, a.expo, div_px, sep_px, &n);
            ndv[x] = n;
            po[x] = p;
        }
        __syncthreads();

        // The 8 scatter rounds.  In round k pair i targets column clamp(cb_i + k), cb_i = floor(min(dest_i, dest_i+1)),
        // and only the highest-index pair targeting a column decides it (Q8).  So one pass builds
        //     M[c] = max { i : cb_i == c }        (cb clamped into [-8, w+7]),
        // and the decisive pair of column x in round k is M[x - k] (for the two border columns, the maximum over
        // everything the clamp folds onto them).  A pair can only be VALID (0 <= frac < 1) while its column stays
        // inside [dest_i, dest_i+1), which is shorter than 2.5 for connected pairs: rounds k >= 4 never change
        // anything (an invalid decisive pair writes back the value it read), so rounds 0..4 reproduce all 8.
        int* M = win;   // [w + 16], index c + 8
        for (int x = threadIdx.x; x < w + 16; x += blockDim.x) M[x] = -1;
        __syncthreads();
        for (int i = threadIdx.x; i + 1 < w; i += blockDim.x) {
            const float dl = (float)i + po[i], dr = (float)(i + 1) + po[i + 1];
            const float fl = floorf(fminf(dl, dr));
            const int cbc = (fl < -8.0f) ? -8 : ((fl > (float)(w + 7)) ? w + 7 : (int)fl);
            atomicMax(&M[cbc + 8], i);
            // rounds in which this pair CAN be valid: 0 <= frac < 1 needs (c - dl) between 0 and safe (correctly rounded
            // division is monotone, so this pre-test only removes pairs the full test below would reject as well)
            uint32_t vmask = 0;
            if (fabsf(po[i + 1] - po[i]) < 1.5f && fl >= -8.0f && fl <= (float)(w + 7)) {
                const float sw = dr - dl;
                const float safe = (fabsf(sw) < 1e-4f) ? 1.0f : sw;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int c = (int)fl + k;
                    const float num = (float)c - dl;
                    const bool pass = (safe > 0.0f) ? (num >= 0.0f && num < safe) : (num <= 0.0f && num > safe);
                    if (c >= 0 && c < w && pass) vmask |= 1u << k;
                }
            }
            vm[i] = (unsigned char)vmask;
        }
        __syncthreads();
        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            float z = -1.0f, sv = -1.0f;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                int i;
                if (x == 0) {
                    i = -1;
                    for (int c = -8; c <= -k; ++c) i = max(i, M[c + 8]);
                    if (w == 1) i = -1;
                } else if (x == w - 1) {
                    i = -1;
                    for (int c = w - 1 - k; c <= w + 7; ++c) i = max(i, M[c + 8]);
                } else {
                    i = M[x - k + 8];
                }
                if (i < 0) continue;
                if (!((vm[i] >> k) & 1u)) continue;
                const float dl = (float)i + po[i], dr = (float)(i + 1) + po[i + 1];
                const bool connected = fabsf(po[i + 1] - po[i]) < 1.5f;
                const float dm = fminf(dl, dr);
                const long long c = (long long)floorf(dm) + k;
                const float sw = dr - dl;
                const float safe = (fabsf(sw) < 1e-4f) ? 1.0f : sw;
                const float frac = ((float)c - dl) / safe;
                const bool valid = connected && c >= 0 && c < w && frac >= 0.0f && frac < 1.0f;
                if (!valid) continue;
                const float a0 = ndv[i] * (1.0f - frac), a1 = ndv[i + 1] * frac;
                const float zi = a0 + a1;
                if (zi > z + 1e-6f) { z = zi; sv = (float)i + frac; }
            }
            zb[x] = z;
            src[x] = sv;
        }
        __syncthreads();

        // filled bitmap, row-wide right-most filled column (SIG:404-410 quirk), unfilled mask
        if (threadIdx.x == 0) s_last = -1;
        __syncthreads();
        {
            int last = -1;
            const int wpad = nwords << 5;
            for (int x = threadIdx.x; x < wpad; x += blockDim.x) {
                bool f = (x < w) && !(src[x] < 0.0f);
                uint32_t b = __ballot_sync(0xffffffffu, f);
                if ((threadIdx.x & 31) == 0) {
                    fbits[x >> 5] = b;
                    uint32_t valid = (x + 32 <= w) ? 0xffffffffu : ((1u << (w - x)) - 1u);
                    ubits[x >> 5] |= (~b) & valid;
                }
                if (f) last = x;
            }
            for (int o = 16; o; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
            if ((threadIdx.x & 31) == 0 && last >= 0) atomicMax(&s_last, last);
        }
        __syncthreads();
        const int lastf = s_last;

        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            float s = src[x];
            if (s < 0.0f) {
                // nearest filled column at or left of x
                int ln = -1;
                {
                    int wi = x >> 5;
                    uint32_t m = fbits[wi] & (0xffffffffu >> (31 - (x & 31)));
                    while (true) {
                        if (m) { ln = (wi << 5) + 31 - __clz(m); break; }
                        if (--wi < 0) break;
                        m = fbits[wi];
                    }
                }
                int rn = (lastf >= x) ? lastf : -1;
                bool hl = ln >= 0, hr = rn >= 0;
                if (hl || hr) {
                    int li = hl ? ln : 0, ri = hr ? rn : 0;
                    float ls = src[li], rs = src[ri], lz = zb[li], rz = zb[ri];
                    float ld = (float)(x - ln), rd = (float)(rn - x);
                    float tot = ld + rd;
                    if (tot < 1.0f) tot = 1.0f;
                    float t = ld / tot;
                    if (!hl) t = 1.0f;
                    if (!hr) t = 0.0f;
                    float tb = (lz < rz) ? sqrtf(t) : 1.0f - sqrtf(1.0f - t);
                    float g0 = ls * (1.0f - tb), g1 = rs * tb;
                    s = g0 + g1;
                }
            }
            s = s < 0.0f ? 0.0f : (s > (float)(w - 1) ? (float)(w - 1) : s);
            if (!emit) continue;
            // grid_sample: gx = s*2/(W-1) - 1, ix = (gx+1) * (W-1)/2, clipped
            float t2 = s * 2.0f;
            float gx = t2 / (float)(w - 1) - 1.0f;
            float ix = (gx + 1.0f) * sx;
            ix = fminf(fmaxf(ix, 0.0f), (float)(w - 1));
            float fx = floorf(ix);
            int x0 = (int)fx, x1 = x0 + 1;
            float wx1 = ix - fx, wx0 = 1.0f - wx1;
            float w_nw = wx0 * wy0, w_ne = wx1 * wy0, w_sw = wx0 * wy1, w_se = wx1 * wy1;
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
                float p_nw = r0[x0 * 3 + ch];
                float p_ne = (x1 < w) ? r0[x1 * 3 + ch] : 0.0f;
                float p_sw = r1 ? r1[x0 * 3 + ch] : 0.0f;
                float p_se = (r1 && x1 < w) ? r1[x1 * 3 + ch] : 0.0f;
                float acc = p_nw * w_nw;
                acc = fmaf(p_ne, w_ne, acc);
                acc = fmaf(p_sw, w_sw, acc);
                acc = fmaf(p_se, w_se, acc);
                orow[x * 3 + ch] = acc;
            }
        }
        __syncthreads();
    }
    // M2: mask = unfilled_left | unfilled_right, [n][h][w]
    float* mrow = a.mask + ((int64_t)frame * h + y) * w;
    for (int x = threadIdx.x; x < w; x += blockDim.x) mrow[x] = ((ubits[x >> 5] >> (x & 31)) & 1u) ? 1.0f : 0.0f;
}

template <int TPB>   // CTA size the kernel is compiled for (register budget): 256 (five CTAs per SM at 1080p), or 512 for rows that leave room for two CTAs per SM
__global__ void __launch_bounds__(TPB, TPB == 256 ? 5 : 2) k_gpuwarp(const GpuWarpArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    gpuwarp_row(a, blockIdx.x, blockIdx.y, smem_f);
}

// Rows whose state does not fit a CTA's shared memory (wider than ~9000 px: 16K panoramas): a few CTAs per SM walk the
// (row, frame) items with the state in their slice of a global scratch buffer.  Slow next to the shared-memory kernel, but
// the reference has no width limit either.
__global__ void __launch_bounds__(512) k_gpuwarp_wide(const GpuWarpArgs a) {
    float* mem = reinterpret_cast<float*>(reinterpret_cast<char*>(a.row_scratch) + (size_t)blockIdx.x * a.row_scratch_stride);
    for (int item = blockIdx.x; item < a.h * a.n; item += gridDim.x) {
        gpuwarp_row(a, item % a.h, item / a.h, mem);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Mesh warp: forward_warp_mesh (SIG:453-689), what 'GPU Warp (Fast)' runs when the host has ModernGL (SIG:1068-1071).
//
// The reference hands the per-pixel vertex grid to OpenGL; which fragments a GPU's rasteriser produces for it, and
// with what interpolation rounding, is implementation-defined, so there is no bit pattern to be identical to
// ("parity unpinned": DESIGN.md section 9).  This is a software rasteriser for the same mesh with a fixed, documented
// rule set -- the one oracle/stereo_oracle.c:orc_mesh_raster restates -- and it IS bit-exact against that:
//   * vertex (r, c) sits at window X = (clip_x + 1) * W/2, Y = H - (clip_y + 1) * H/2 with the reference's clip
//     coordinates (SIG:545-551); vertex rows are H/(H-1) apart, so output row j (centre j + 0.5) lies in exactly one
//     strip r of quads, at fy = (j + 0.5 - Y_r) / (Y_{r+1} - Y_r);
//   * triangle a = (v00, v10, v01) spans, on that scanline, from the edge v00-v01 to the edge v10-v01; triangle
//     b = (v11, v10, v01) from v10-v01 to v10-v11 (SIG:513-520); a pixel centre i + 0.5 is covered when
//     lo <= i + 0.5 < hi (top-left rule for shared edges);
//   * the fragment's depth is the normalised depth interpolated along the two edges and then across; the larger
//     (nearer) one wins -- the reference's '<' test on clip_z = (1 - nd) * 1.98 - 0.99 -- and equal depths go to the
//     triangle drawn first (all a, then all b, SIG:520; each in raster order);
//   * colours interpolate the same way; uncovered pixels take the nearest covered pixel on the eye's fill side
//     (SIG:655-683) and the mask is the coverage before that fill (SIG:639).
// Triangles are culled with the reference's rule (SIG:522-537): kept when max pairwise offset difference < 1.5 in ANY
// frame of the sub-batch -- one topology per sub-batch (k_mesh_keep), which is why a triangle can be badly stretched
// in the other frames and why the z-test is needed at all.

__device__ __forceinline__ float mesh_edge(float xa, float xb, float fy) { float t = xb - xa; t = fy * t; return xa + t; }
__device__ __forceinline__ float mesh_lerp(float a, float b, float t) { float d = b - a; d = t * d; return a + d; }
__device__ __forceinline__ float mesh_window_y(int q, int h) {
    const float cy = -(((float)q / (float)(h - 1)) * 2.0f - 1.0f);
    return (float)h - (cy + 1.0f) * ((float)h / 2.0f);
}
__device__ __forceinline__ float mesh_window_x(float dest, int w) {
    const float cx = (dest / (float)(w - 1)) * 2.0f - 1.0f;
    return (cx + 1.0f) * ((float)w / 2.0f);
}

// keep[eye][sub-batch][strip r][quad x]: bit 0 = triangle a kept, bit 1 = triangle b kept.  One CTA per
// (strip, sub-batch, eye) walks the sub-batch's frames with the two vertex rows' offsets in shared memory.
__global__ void __launch_bounds__(256) k_mesh_keep(const GpuWarpArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    const int w = a.w, h = a.h, r = blockIdx.x, g = blockIdx.y, eye = blockIdx.z;
    if (a.eye[eye].passthrough) return;
    float* p0 = smem_f;
    float* p1 = p0 + w;
    unsigned char* kb = reinterpret_cast<unsigned char*>(p1 + w);
    for (int x = threadIdx.x; x < w; x += blockDim.x) kb[x] = 0;
    const float div_px = (float)a.eye[eye].div_px, sep_px = (float)a.eye[eye].sep_px;
    const int f0 = g * a.group, f1 = min(f0 + a.group, a.n);
    for (int f = f0; f < f1; ++f) {
        const DepthNorm nm = gw_depth_norm(a, f, eye);
        const float* d0 = a.depth[eye] + ((int64_t)f * h + r) * w;
        __syncthreads();
        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            float n;
            p0[x] = gw_offset(nm, d0[x], a.conv, a.expo, div_px, sep_px, &n);
            p1[x] = gw_offset(nm, d0[w + x], a.conv, a.expo, div_px, sep_px, &n);
        }
        __syncthreads();
        for (int x = threadIdx.x; x + 1 < w; x += blockDim.x) {
            const float o00 = p0[x], o10 = p0[x + 1], o01 = p1[x], o11 = p1[x + 1];
            const float thr = 1.5f, dd = fabsf(o10 - o01);   // the diagonal both triangles share
            const bool ka = fabsf(o00 - o10) < thr && fabsf(o00 - o01) < thr && dd < thr;
            const bool kbb = fabsf(o11 - o10) < thr && fabsf(o11 - o01) < thr && dd < thr;
            kb[x] |= (unsigned char)((ka ? 1 : 0) | (kbb ? 2 : 0));
        }
    }
    const int ng = (a.n + a.group - 1) / a.group;
    unsigned char* out = a.keep + (((int64_t)eye * ng + g) * (h - 1) + r) * (int64_t)(w - 1);
    for (int x = threadIdx.x; x + 1 < w; x += blockDim.x) out[x] = kb[x];
}

// One (output row, frame), both eyes in turn, composed into the final stereo layout like k_gpuwarp.
__device__ __forceinline__ void meshwarp_row(const GpuWarpArgs& a, const int y, const int frame, float* smem_f) {
    const int w = a.w, h = a.h;
    const int nwords = (w + 31) >> 5;
    unsigned long long* key = reinterpret_cast<unsigned long long*>(smem_f);   // z-buffer: ordered depth << 32 | ~draw order
    float* X0 = reinterpret_cast<float*>(key + w);
    float* X1 = X0 + w;
    float* N0 = X1 + w;
    float* N1 = N0 + w;
    uint32_t* fbits = reinterpret_cast<uint32_t*>(N1 + w);    // covered
    uint32_t* ubits = fbits + nwords;                          // uncovered in either eye (mask)

    for (int i = threadIdx.x; i < nwords; i += blockDim.x) ubits[i] = 0u;

    const int mode = a.mode;
    int wo = w, ho = h;
    if (mode == CS_MODE_LEFT_RIGHT || mode == CS_MODE_RIGHT_LEFT) wo = 2 * w;
    if (mode == CS_MODE_TOP_BOTTOM || mode == CS_MODE_BOTTOM_TOP) ho = 2 * h;
    float* outf = a.stereo + (int64_t)frame * ho * wo * 3;
    const float* img = a.image + (int64_t)frame * h * w * 3;
    const int ng = (a.n + a.group - 1) / a.group;

    // the strip of quads this row's pixel centres fall in: largest r <= h - 2 with Y(r) <= y + 0.5
    const float ys = (float)y + 0.5f;
    int r = 0;
    float fy = 0.0f;
    if (h >= 2) {
        r = (int)(((double)y + 0.5) * (double)(h - 1) / (double)h);
        r = max(0, min(r, h - 2));
        while (r > 0 && mesh_window_y(r, h) > ys) --r;
        while (r < h - 2 && mesh_window_y(r + 1, h) <= ys) ++r;
        const float yr = mesh_window_y(r, h), yn = mesh_window_y(r + 1, h);
        fy = (ys - yr) / (yn - yr);
    }
    const float* r0 = img + (int64_t)r * w * 3;
    const float* r1 = r0 + (int64_t)w * 3;

    for (int eye = 0; eye < 2; ++eye) {
        int oy = y, ox = 0, ch_lo = 0, ch_hi = 3;
        bool emit = true;
        switch (mode) {
            case CS_MODE_LEFT_RIGHT: ox = eye ? w : 0; break;
            case CS_MODE_RIGHT_LEFT: ox = eye ? 0 : w; break;
            case CS_MODE_TOP_BOTTOM: oy = eye ? y + h : y; break;
            case CS_MODE_BOTTOM_TOP: oy = eye ? y : y + h; break;
            case CS_MODE_RED_CYAN: if (eye == 0) { ch_lo = 0; ch_hi = 1; } else { ch_lo = 1; ch_hi = 3; } break;
            case CS_MODE_CYAN_RED: if (eye == 1) { ch_lo = 0; ch_hi = 1; } else { ch_lo = 1; ch_hi = 3; } break;
            case CS_MODE_LEFT_ONLY: emit = (eye == 0); break;
            default: emit = (eye == 1); break;
        }
        float* orow = outf + ((int64_t)oy * wo + ox) * 3;

        if (a.eye[eye].passthrough) {
            if (emit)
                for (int x = threadIdx.x; x < w; x += blockDim.x)
                    for (int ch = ch_lo; ch < ch_hi; ++ch) orow[x * 3 + ch] = img[((int64_t)y * w + x) * 3 + ch];
            continue;
        }
        __syncthreads();   // the previous eye's readers are done with key / X / N / fbits
        const bool degenerate = h < 2 || w < 2;   // no triangles at all: nothing is covered
        const float div_px = (float)a.eye[eye].div_px, sep_px = (float)a.eye[eye].sep_px;
        if (!degenerate) {
            const DepthNorm nm = gw_depth_norm(a, frame, eye);
            const float* d0 = a.depth[eye] + ((int64_t)frame * h + r) * w;
            for (int x = threadIdx.x; x < w; x += blockDim.x) {
                float n0, n1;
                const float q0 = gw_offset(nm, d0[x], a.conv, a.expo, div_px, sep_px, &n0);
                const float q1 = gw_offset(nm, d0[w + x], a.conv, a.expo, div_px, sep_px, &n1);
                N0[x] = n0; N1[x] = n1;
                X0[x] = mesh_window_x((float)x + q0, w);
                X1[x] = mesh_window_x((float)x + q1, w);
                key[x] = 0ull;
            }
        } else {
            for (int x = threadIdx.x; x < w; x += blockDim.x) key[x] = 0ull;
        }
        __syncthreads();

        // the two edges and edge depths of triangle `pass` of quad x on this scanline
        auto tri = [&](int x, int pass, float& e1, float& e2, float& n1, float& n2) {
            const float ed = mesh_edge(X0[x + 1], X1[x], fy), nd = mesh_lerp(N0[x + 1], N1[x], fy);   // the diagonal v10-v01
            if (pass == 0) { e1 = mesh_edge(X0[x], X1[x], fy); n1 = mesh_lerp(N0[x], N1[x], fy); e2 = ed; n2 = nd; }
            else { e1 = ed; n1 = nd; e2 = mesh_edge(X0[x + 1], X1[x + 1], fy); n2 = mesh_lerp(N0[x + 1], N1[x + 1], fy); }
        };

        if (!degenerate) {
            const unsigned char* kp = a.keep + (((int64_t)eye * ng + frame / a.group) * (h - 1) + r) * (int64_t)(w - 1);
            for (int x = threadIdx.x; x + 1 < w; x += blockDim.x) {
                const int kept = kp[x];
                for (int pass = 0; pass < 2; ++pass) {
                    if (!((kept >> pass) & 1)) continue;
                    float e1, e2, n1, n2;
                    tri(x, pass, e1, e2, n1, n2);
                    const float lo = fminf(e1, e2), hi = fmaxf(e1, e2);
                    if (!(lo < hi)) continue;
                    const float c0 = ceilf(lo - 0.5f), c1 = ceilf(hi - 0.5f) - 1.0f;
                    if (!(c0 <= (float)(w - 1)) || !(c1 >= 0.0f)) continue;
                    const int i0 = c0 < 0.0f ? 0 : (int)c0, i1 = c1 > (float)(w - 1) ? w - 1 : (int)c1;
                    const uint32_t tie = 0xffffffffu - (uint32_t)(pass * (w - 1) + x);
                    for (int i = i0; i <= i1; ++i) {
                        const float xs = (float)i + 0.5f;
                        if (!(lo <= xs && xs < hi)) continue;
                        const float t = (xs - e1) / (e2 - e1);
                        const float z = mesh_lerp(n1, n2, t);
                        uint32_t zo = __float_as_uint(z);
                        zo = (zo & 0x80000000u) ? ~zo : (zo | 0x80000000u);
                        atomicMax(&key[i], ((unsigned long long)zo << 32) | tie);
                    }
                }
            }
        }
        __syncthreads();

        {   // coverage bitmap, mask
            const int wpad = nwords << 5;
            for (int x = threadIdx.x; x < wpad; x += blockDim.x) {
                const bool f = (x < w) && key[x] != 0ull;
                const uint32_t b = __ballot_sync(0xffffffffu, f);
                if ((threadIdx.x & 31) == 0) {
                    fbits[x >> 5] = b;
                    const uint32_t valid = (x + 32 <= w) ? 0xffffffffu : ((1u << (w - x)) - 1u);
                    ubits[x >> 5] |= (~b) & valid;
                }
            }
        }
        __syncthreads();
        if (!emit) continue;

        const bool from_left = div_px >= 0.0f;   // SIG:663
        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            int sc = x;
            if (key[x] == 0ull) {     // gap: nearest covered column on the fill side, if there is one
                sc = -1;
                int wi = x >> 5;
                if (from_left) {
                    uint32_t m = fbits[wi] & (0xffffffffu >> (31 - (x & 31)));
                    while (true) {
                        if (m) { sc = (wi << 5) + 31 - __clz(m); break; }
                        if (--wi < 0) break;
                        m = fbits[wi];
                    }
                } else {
                    uint32_t m = fbits[wi] & (0xffffffffu << (x & 31));
                    while (true) {
                        if (m) { sc = (wi << 5) + __ffs(m) - 1; break; }
                        if (++wi >= nwords) break;
                        m = fbits[wi];
                    }
                }
            }
            if (sc < 0) {
                for (int ch = ch_lo; ch < ch_hi; ++ch) orow[x * 3 + ch] = 0.0f;    // the cleared colour buffer
                continue;
            }
            const uint32_t order = 0xffffffffu - (uint32_t)(key[sc] & 0xffffffffull);
            const int pass = order >= (uint32_t)(w - 1) ? 1 : 0, q = (int)order - pass * (w - 1);
            float e1, e2, n1, n2;
            tri(q, pass, e1, e2, n1, n2);
            const float t = (((float)sc + 0.5f) - e1) / (e2 - e1);
            // vertex pairs of the two edges: a = (v00-v01, v10-v01), b = (v10-v01, v10-v11)
            const int ta = pass ? q + 1 : q, ba = q, tb = q + 1, bb = pass ? q + 1 : q;
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
                const float a1 = mesh_lerp(r0[ta * 3 + ch], r1[ba * 3 + ch], fy);
                const float a2 = mesh_lerp(r0[tb * 3 + ch], r1[bb * 3 + ch], fy);
                orow[x * 3 + ch] = mesh_lerp(a1, a2, t);
            }
        }
    }
    __syncthreads();
    float* mrow = a.mask + ((int64_t)frame * h + y) * w;
    for (int x = threadIdx.x; x < w; x += blockDim.x) mrow[x] = ((ubits[x >> 5] >> (x & 31)) & 1u) ? 1.0f : 0.0f;
}

__global__ void __launch_bounds__(512) k_meshwarp(const GpuWarpArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    meshwarp_row(a, blockIdx.x, blockIdx.y, smem_f);
}
__global__ void __launch_bounds__(512) k_meshwarp_wide(const GpuWarpArgs a) {      // see k_gpuwarp_wide
    float* mem = reinterpret_cast<float*>(reinterpret_cast<char*>(a.row_scratch) + (size_t)blockIdx.x * a.row_scratch_stride);
    for (int item = blockIdx.x; item < a.h * a.n; item += gridDim.x) {
        meshwarp_row(a, item % a.h, item / a.h, mem);
        __syncthreads();
    }
}

// global row scratch for rows too wide for shared memory: bytes per CTA and CTAs (both warps share the geometry)
constexpr size_t kRowSmemLimit = 227 * 1024;
static size_t gw_row_bytes(int w) {
    const size_t nwords = (size_t)(w + 31) >> 5;
    const size_t b = (size_t)w * 24 + 64 + nwords * 8 + (size_t)w + 16;     // k_gpuwarp's layout, the larger of the two
    return (b + 255) & ~(size_t)255;
}
static int gw_wide_ctas() { return 2 * sm_count(); }
size_t gpuwarp_row_scratch_stride(int w) { return gw_row_bytes(w); }
size_t gpuwarp_row_scratch_bytes(int w) { return gw_row_bytes(w) > kRowSmemLimit ? gw_row_bytes(w) * (size_t)gw_wide_ctas() : 0; }

size_t mesh_keep_bytes(int n, int h, int w) { return (size_t)2 * n * (h > 1 ? h - 1 : 1) * (w > 1 ? w - 1 : 1); }

cudaError_t launch_meshwarp(const GpuWarpArgs& a, cudaStream_t s) {
    const int nwords = (a.w + 31) >> 5;
    if (!a.keep || a.group < 1) return cudaErrorInvalidValue;
    const int ng = (a.n + a.group - 1) / a.group;
    if (a.h >= 2 && a.w >= 2) {
        const size_t ksm = (size_t)a.w * 9;
        if (ksm > 227 * 1024) return cudaErrorInvalidValue;
        if (ksm > 48 * 1024) cudaFuncSetAttribute(k_mesh_keep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksm);
        prof_begin(K_GPUWARP, s);
        k_mesh_keep<<<dim3(a.h - 1, ng, 2), 256, ksm, s>>>(a);
        prof_end(K_GPUWARP, s);
        count_launch();
    }
    const size_t smem = (size_t)a.w * 24 + (size_t)nwords * 8;
    if (gw_row_bytes(a.w) > kRowSmemLimit) {     // (the same threshold as the scatter warp: one scratch geometry for both)
        if (!a.row_scratch || a.row_scratch_stride < gw_row_bytes(a.w)) return cudaErrorInvalidValue;
        prof_begin(K_GPUWARP, s);
        k_meshwarp_wide<<<gw_wide_ctas(), 512, 0, s>>>(a);
        prof_end(K_GPUWARP, s);
        count_launch();
        return cudaGetLastError();
    }
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_meshwarp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // a wide row's shared memory (24 B per column) leaves room for two CTAs per SM: make them 512 threads, like k_gpuwarp
    const int tpb = smem > 56 * 1024 ? 512 : 256;
    prof_begin(K_GPUWARP, s);
    k_meshwarp<<<dim3(a.h, a.n), tpb, smem, s>>>(a);
    prof_end(K_GPUWARP, s);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_gpuwarp(const GpuWarpArgs& a, cudaStream_t s) {
    const int nwords = (a.w + 31) >> 5;
    size_t smem = (size_t)a.w * 20 + 64 + (size_t)nwords * 8 + (size_t)a.w + 16;
    if (gw_row_bytes(a.w) > kRowSmemLimit) {
        if (!a.row_scratch || a.row_scratch_stride < gw_row_bytes(a.w)) return cudaErrorInvalidValue;
        prof_begin(K_GPUWARP, s);
        k_gpuwarp_wide<<<gw_wide_ctas(), 512, 0, s>>>(a);
        prof_end(K_GPUWARP, s);
        count_launch();
        return cudaGetLastError();
    }
    // a row's shared memory (21 B per column) limits the CTAs per SM: keep ~32 warps resident by widening the CTA
    const bool wide = smem > 56 * 1024;
    if (smem > 48 * 1024) {
        if (wide) cudaFuncSetAttribute(k_gpuwarp<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        else cudaFuncSetAttribute(k_gpuwarp<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    prof_begin(K_GPUWARP, s);
    if (wide) k_gpuwarp<512><<<dim3(a.h, a.n), 512, smem, s>>>(a);
    else k_gpuwarp<256><<<dim3(a.h, a.n), 256, smem, s>>>(a);
    prof_end(K_GPUWARP, s);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cs

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

template <int TPB>   // CTA size the kernel is compiled for (register budget): 256, or 512 for rows that leave room for two CTAs per SM
__global__ void __launch_bounds__(TPB) k_gpuwarp(const GpuWarpArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    const int w = a.w, h = a.h, y = blockIdx.x, frame = blockIdx.y;
    const int nwords = (w + 31) >> 5;
    float* ndv = smem_f;
    float* po = ndv + w;
    float* dest = po + w;
    float* src = dest + w;
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
        // depth scale chain: x255 if the sub-batch max <= 1 (SIG:1045), blur, /255 if any frame > 1 (SIG:314-316)
        const FrameStats st = a.stats[frame];
        float pre = 1.0f;
        int omin, omax;
        bool div255;
        if (a.use_blur_stats) {
            omin = eye ? st.r_min : st.l_min; omax = eye ? st.r_max : st.l_max;
            div255 = gw_group_max(a.stats, frame, a.group, a.n, eye ? 2 : 1) > 1.0f;
        } else {
            omin = st.gray_min; omax = st.gray_max;
            float gm = gw_group_max(a.stats, frame, a.group, a.n, 0);
            pre = (a.prescale && gm <= 1.0f) ? 255.0f : 1.0f;
            div255 = gm * pre > 1.0f;
        }
        float dmin = ord2f(omin), dmax = ord2f(omax);
        if (pre != 1.0f) { dmin = dmin * pre; dmax = dmax * pre; }
        if (div255) { dmin = dmin / 255.0f; dmax = dmax / 255.0f; }
        const float range = dmax - dmin;
        const bool flat = !(range > 1e-6f);
        const float rng = range < 1e-6f ? 1e-6f : range;
        const float* dep = a.depth[eye] + (int64_t)frame * h * w + (int64_t)y * w;

        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            float dv = dep[x];
            if (pre != 1.0f) dv = dv * pre;
            if (div255) dv = dv / 255.0f;
            float n = flat ? 0.0f : (dv - dmin) / rng;
            ndv[x] = n;
            float sh = n - a.conv;
            float sg = (sh > 0.0f) ? 1.0f : ((sh < 0.0f) ? -1.0f : 0.0f);
            float od = sg * gw_pow(fabsf(sh), a.expo);
            float m = od * div_px;
            float p = m + sep_px;
            po[x] = p;
            dest[x] = (float)x + p;
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
            const float dl = dest[i], dr = dest[i + 1];
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
                const float dl = dest[i], dr = dest[i + 1];
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

cudaError_t launch_gpuwarp(const GpuWarpArgs& a, cudaStream_t s) {
    const int nwords = (a.w + 31) >> 5;
    size_t smem = (size_t)a.w * 24 + 64 + (size_t)nwords * 8 + (size_t)a.w + 16;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    // a row's shared memory (25 B per column) limits the CTAs per SM: keep ~32 warps resident by widening the CTA
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

// cs_prep.cu -- N1/L1/O1: depth RGB->gray + per-frame min/max, image quantisation to RGBX8,
// the no-blur depth outputs and the debug export of the integer shift indices.
//
// HBM traffic (per input pixel): reads image 12 B + depth 12 B (streamed once, evict-first),
// writes gray 4 B + RGBX8 4 B scratch that the later kernels re-read from L2.
#include "cs_internal.cuh"

namespace cs {

__global__ void k_init_stats(FrameStats* st, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        FrameStats s;
        s.gray_min = s.l_min = s.r_min = 0x7FFFFFFF;
        s.gray_max = s.l_max = s.r_max = (int)0x80000000;
        s.pad0 = s.pad1 = 0;
        st[i] = s;
    }
}

cudaError_t launch_init_stats(FrameStats* stats, int n, cudaStream_t s) {
    prof_begin(K_MISC, s);
    k_init_stats<<<(n + 127) / 128, 128, 0, s>>>(stats, n);
    prof_end(K_MISC, s);
    count_launch();
    return cudaGetLastError();
}

// gray = 0.2989*R + 0.5870*G + 0.1140*B, float32, left to right (GS:206-207 / GS:134-135)
__device__ __forceinline__ float gray_of(float r, float g, float b) {
    float a = 0.2989f * r;
    float c = 0.5870f * g;
    float e = 0.1140f * b;
    float s = a + c;
    return s + e;
}

// ---- N1 depth resize: gray + F.interpolate(..., mode='bilinear', align_corners=False) (GS:141-148, GS:214-220) ----
// torch's CPU kernel in strict float32 (no contraction; this file is built with -fmad=false):
//   src = max(scale*(dst+0.5) - 0.5, 0), i0 = min(floor(src), in-1), i1 = i0 + (i0 < in-1), l1 = clamp(src - i0, 0, 1),
//   l0 = 1 - l1; an axis whose size does not change maps to itself with weights (1, 0).
//   out = (a00*lx0 + a01*lx1)*ly0 + (a10*lx0 + a11*lx1)*ly1, or torch's direct form for small outputs
//   (h + w <= 128): ((ly0*lx0)*a00 + (ly0*lx1)*a01 + (ly1*lx0)*a10) + (ly1*lx1)*a11, summed left to right.
struct ResizeAxis { int i0, i1; float l0, l1; };
__device__ __forceinline__ ResizeAxis resize_axis(int d, int in, int out, float scale) {
    ResizeAxis r;
    if (in == out) { r.i0 = r.i1 = d; r.l0 = 1.0f; r.l1 = 0.0f; return r; }
    float src = fmaxf(scale * ((float)d + 0.5f) - 0.5f, 0.0f);
    int a = min((int)floorf(src), in - 1);
    float lam = fminf(fmaxf(src - (float)a, 0.0f), 1.0f);
    r.i0 = a;
    r.i1 = a + (a < in - 1 ? 1 : 0);
    r.l1 = lam;
    r.l0 = 1.0f - lam;
    return r;
}

template <int C>
__device__ __forceinline__ float depth_tap(const float* __restrict__ p, int64_t i, int c_any) {
    if (C == 3) return gray_of(__ldg(p + i * 3), __ldg(p + i * 3 + 1), __ldg(p + i * 3 + 2));
    if (C == 1) return __ldg(p + i);
    return __ldg(p + i * c_any);   // other channel counts: channel 0 (GS:138)
}

// One thread = one output pixel; the source frame (smaller or similar size) is served by L1/L2.
template <int C>
__global__ void __launch_bounds__(256) k_resize_gray(const float* __restrict__ depth, int c_any, int dh, int dw,
                                                     int h, int w, float sy, float sx, int direct,
                                                     float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int cc = (C > 0) ? C : c_any;
    if (x >= w) return;
    const float* src = depth + (int64_t)blockIdx.z * dh * dw * cc;
    const ResizeAxis ay = resize_axis(y, dh, h, sy), ax = resize_axis(x, dw, w, sx);
    const float a00 = depth_tap<C>(src, (int64_t)ay.i0 * dw + ax.i0, c_any);
    const float a01 = depth_tap<C>(src, (int64_t)ay.i0 * dw + ax.i1, c_any);
    const float a10 = depth_tap<C>(src, (int64_t)ay.i1 * dw + ax.i0, c_any);
    const float a11 = depth_tap<C>(src, (int64_t)ay.i1 * dw + ax.i1, c_any);
    float v;
    if (direct) {
        const float w00 = ay.l0 * ax.l0, w01 = ay.l0 * ax.l1, w10 = ay.l1 * ax.l0, w11 = ay.l1 * ax.l1;
        float acc = w00 * a00 + w01 * a01;
        acc = acc + w10 * a10;
        v = acc + w11 * a11;
    } else {
        const float top = a00 * ax.l0 + a01 * ax.l1;
        const float bot = a10 * ax.l0 + a11 * ax.l1;
        v = top * ay.l0 + bot * ay.l1;
    }
    out[((int64_t)blockIdx.z * h + y) * w + x] = v;
}

cudaError_t launch_resize_gray(const float* depth, int n, int dh, int dw, int c, int h, int w, float* out,
                               cudaStream_t s) {
    const float sy = (float)dh / (float)h, sx = (float)dw / (float)w;   // area_pixel_compute_scale, float32
    const int direct = (h + w <= 128) ? 1 : 0;
    dim3 grid((w + 255) / 256, h, n);
    prof_begin(K_RESIZE, s);
    if (c == 3) k_resize_gray<3><<<grid, 256, 0, s>>>(depth, c, dh, dw, h, w, sy, sx, direct, out);
    else if (c == 1) k_resize_gray<1><<<grid, 256, 0, s>>>(depth, c, dh, dw, h, w, sy, sx, direct, out);
    else k_resize_gray<0><<<grid, 256, 0, s>>>(depth, c, dh, dw, h, w, sy, sx, direct, out);
    prof_end(K_RESIZE, s);
    count_launch();
    return cudaGetLastError();
}

// clip(x*255, 0, 255).astype(uint8): truncation, NaN -> 0 (SIG:1508)
__device__ __forceinline__ int quant_u8(float x) {
    float v = x * 255.0f;
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    return __float2int_rz(v);
}

// One thread = 4 consecutive pixels of one frame: 3 x 128-bit loads per input tensor.
template <int C, bool QUANT>
__global__ void __launch_bounds__(256) k_prepare(const float* __restrict__ image,
                                                 const float* __restrict__ depth, int c_any,
                                                 int64_t npx, int vec_ok, float* __restrict__ gray,
                                                 uint32_t* __restrict__ image_u8,
                                                 FrameStats* __restrict__ stats) {
    const int frame = blockIdx.y;
    const float* img = image ? image + (int64_t)frame * npx * 3 : nullptr;
    const int cc = (C > 0) ? C : c_any;
    const float* dep = depth + (int64_t)frame * npx * cc;
    float* gr = gray + (int64_t)frame * npx;
    uint32_t* iq = QUANT ? image_u8 + (int64_t)frame * npx : nullptr;

    const uint64_t pol = policy_evict_first();
    float lo = INFINITY, hi = -INFINITY;
    const int64_t nquad = vec_ok ? (npx >> 2) : 0;  // vector path needs npx % 4 == 0 and 16 B bases
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquad;
         q += (int64_t)gridDim.x * blockDim.x) {
        float g[4];
        if (C == 3) {
            const float4* p = reinterpret_cast<const float4*>(dep) + q * 3;
            float4 a = ld_stream_f4(p, pol), b = ld_stream_f4(p + 1, pol), c = ld_stream_f4(p + 2, pol);
            g[0] = gray_of(a.x, a.y, a.z);
            g[1] = gray_of(a.w, b.x, b.y);
            g[2] = gray_of(b.z, b.w, c.x);
            g[3] = gray_of(c.y, c.z, c.w);
        } else if (C == 1) {
            float4 a = ld_stream_f4(reinterpret_cast<const float4*>(dep) + q, pol);
            g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) g[k] = ld_stream_f1(dep + (q * 4 + k) * cc, pol);
        }
        reinterpret_cast<float4*>(gr)[q] = make_float4(g[0], g[1], g[2], g[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) { lo = fminf(lo, g[k]); hi = fmaxf(hi, g[k]); }
        if (QUANT) {
            const float4* p = reinterpret_cast<const float4*>(img) + q * 3;
            float4 a = ld_stream_f4(p, pol), b = ld_stream_f4(p + 1, pol), c = ld_stream_f4(p + 2, pol);
            uint4 o;
            o.x = pack_rgbx(quant_u8(a.x), quant_u8(a.y), quant_u8(a.z));
            o.y = pack_rgbx(quant_u8(a.w), quant_u8(b.x), quant_u8(b.y));
            o.z = pack_rgbx(quant_u8(b.z), quant_u8(b.w), quant_u8(c.x));
            o.w = pack_rgbx(quant_u8(c.y), quant_u8(c.z), quant_u8(c.w));
            reinterpret_cast<uint4*>(iq)[q] = o;
        }
    }
    // scalar path: everything when the frame is not vectorisable (odd sizes)
    {
        for (int64_t i = (nquad << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx;
             i += (int64_t)gridDim.x * blockDim.x) {
            float gv;
            if (cc == 3) gv = gray_of(dep[i * 3], dep[i * 3 + 1], dep[i * 3 + 2]);
            else gv = dep[i * cc];
            gr[i] = gv;
            lo = fminf(lo, gv); hi = fmaxf(hi, gv);
            if (QUANT) iq[i] = pack_rgbx(quant_u8(img[i * 3]), quant_u8(img[i * 3 + 1]), quant_u8(img[i * 3 + 2]));
        }
    }
    lo = warp_min(lo); hi = warp_max(hi);
    __shared__ float s_lo[8], s_hi[8];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_lo[wid] = lo; s_hi[wid] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { lo = fminf(lo, s_lo[k]); hi = fmaxf(hi, s_hi[k]); }
        if (lo <= hi) {
            atomicMin(&stats[frame].gray_min, f2ord(lo));
            atomicMax(&stats[frame].gray_max, f2ord(hi));
        }
    }
}

cudaError_t launch_prepare(const float* image, const float* depth, int n, int h, int w, int c,
                           float* gray, uint32_t* image_u8, FrameStats* stats, cudaStream_t s) {
    const int64_t npx = (int64_t)h * w;
    int64_t nquad = npx >> 2;
    int bx = (int)((nquad + 255) / 256);
    // a few waves of SMs x 8 resident CTAs; grid-stride the rest
    const int cap = sm_count() * 8 * 2;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid(bx, n);
    const int vec = ((uintptr_t)depth % 16 == 0) && (image == nullptr || (uintptr_t)image % 16 == 0) &&
                    ((uintptr_t)gray % 16 == 0) && (image_u8 == nullptr || (uintptr_t)image_u8 % 16 == 0) &&
                    (npx % 4 == 0);
    const bool quant = image_u8 != nullptr;
    prof_begin(K_PREPARE, s);
    if (c == 3) {
        if (quant) k_prepare<3, true><<<grid, 256, 0, s>>>(image, depth, c, npx, vec, gray, image_u8, stats);
        else k_prepare<3, false><<<grid, 256, 0, s>>>(image, depth, c, npx, vec, gray, image_u8, stats);
    } else if (c == 1) {
        if (quant) k_prepare<1, true><<<grid, 256, 0, s>>>(image, depth, c, npx, vec, gray, image_u8, stats);
        else k_prepare<1, false><<<grid, 256, 0, s>>>(image, depth, c, npx, vec, gray, image_u8, stats);
    } else {  // GPU Warp with c not in {1,3}: channel 0 (GS:138-139)
        if (quant) k_prepare<0, true><<<grid, 256, 0, s>>>(image, depth, c, npx, vec, gray, image_u8, stats);
        else k_prepare<0, false><<<grid, 256, 0, s>>>(image, depth, c, npx, vec, gray, image_u8, stats);
    }
    prof_end(K_PREPARE, s);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------- min/max only (stage API)
__global__ void __launch_bounds__(256) k_minmax(const float* __restrict__ src, int64_t npx,
                                                FrameStats* __restrict__ stats) {
    const int frame = blockIdx.y;
    const float* p = src + (int64_t)frame * npx;
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (int64_t)gridDim.x * blockDim.x) {
        float v = p[i];
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(&stats[frame].gray_min, f2ord(lo));
        atomicMax(&stats[frame].gray_max, f2ord(hi));
    }
}

cudaError_t launch_minmax(const float* src, int n, int64_t npx, FrameStats* stats, cudaStream_t s) {
    int bx = (int)((npx + 1023) / 1024);
    if (bx > sm_count() * 4) bx = sm_count() * 4;
    if (bx < 1) bx = 1;
    prof_begin(K_MISC, s);
    k_minmax<<<dim3(bx, n), 256, 0, s>>>(src, npx, stats);
    prof_end(K_MISC, s);
    count_launch();
    return cudaGetLastError();
}

// which 0: {gray_min, gray_max}; 1: {l_min, l_max, r_min, r_max}
__global__ void k_export_stats(const FrameStats* __restrict__ st, int n, int which, float* __restrict__ out, int stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (which == 0) {
        out[i * stride + 0] = ord2f(st[i].gray_min);
        out[i * stride + 1] = ord2f(st[i].gray_max);
    } else {
        out[i * stride + 0] = ord2f(st[i].l_min);
        out[i * stride + 1] = ord2f(st[i].l_max);
        out[i * stride + 2] = ord2f(st[i].r_min);
        out[i * stride + 3] = ord2f(st[i].r_max);
    }
}

cudaError_t launch_export_stats(const FrameStats* stats, int n, int which, float* out, int stride, cudaStream_t s) {
    prof_begin(K_MISC, s);
    k_export_stats<<<(n + 127) / 128, 128, 0, s>>>(stats, n, which, out, stride);
    prof_end(K_MISC, s);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------- depth outputs
// out_kind 1 (CPU techniques, quirk Q1): u8 = trunc(d255 * 255) mod 256, value = u8 / 255
//          2 (GPU Warp, SIG:1125 + GS:165): d255 / 255 if the sub-batch max > 1, clamp to [0,1]
// Every output pixel is three identical floats; a thread takes 4 pixels = three float4 per output.
__device__ __forceinline__ float depth_out_value(float d255, int out_kind, bool div255) {
    if (out_kind == 1) {
        float v = d255 * 255.0f;
        // numpy float->uint8 goes through a wide signed integer and keeps the low byte
        long long iv = (long long)v;
        return (float)(int)(iv & 255) / 255.0f;
    }
    float v = div255 ? d255 / 255.0f : d255;
    return fminf(fmaxf(v, 0.0f), 1.0f);
}

__device__ inline float group_max(const FrameStats* st, int frame, int group, int n, int which) {
    int g0 = (frame / group) * group, g1 = min(g0 + group, n);
    float m = -INFINITY;
    for (int f = g0; f < g1; ++f) {
        int o = which == 0 ? st[f].gray_max : (which == 1 ? st[f].l_max : st[f].r_max);
        m = fmaxf(m, ord2f(o));
    }
    return m;
}

// src_kind 0: gray (apply the x255 decision: per frame for out_kind 1, per group for 2)
//          1: already on the 0..255 scale (blurred L / R)
__global__ void __launch_bounds__(256) k_depth_out(const float* __restrict__ src_l,
                                                   const float* __restrict__ src_r,
                                                   const FrameStats* __restrict__ st, int n, int64_t npx,
                                                   int src_kind, int out_kind, int group, int vec_ok,
                                                   float* __restrict__ out_l, float* __restrict__ out_r) {
    const int frame = blockIdx.y;
    const uint64_t pol = policy_evict_first();
    float scale = 1.0f;
    bool div_l = false, div_r = false;
    if (src_kind == 0) {
        float gm = out_kind == 1 ? ord2f(st[frame].gray_max) : group_max(st, frame, group, n, 0);
        scale = (gm <= 1.0f) ? 255.0f : 1.0f;
        if (out_kind == 2) div_l = div_r = (gm * scale > 1.0f);
    } else if (out_kind == 2) {
        div_l = group_max(st, frame, group, n, 1) > 1.0f;
        div_r = group_max(st, frame, group, n, 2) > 1.0f;
    }
    const float* sl = src_l + (int64_t)frame * npx;
    const float* sr = src_r + (int64_t)frame * npx;
    float4* ol = reinterpret_cast<float4*>(out_l + (int64_t)frame * npx * 3);
    float4* orr = reinterpret_cast<float4*>(out_r + (int64_t)frame * npx * 3);
    // vector path: 4 pixels per thread, one 128-bit load per source and three 128-bit streaming stores per output
    const int64_t nquad = vec_ok ? (npx >> 2) : 0;
    const int64_t nvec = nquad * 3;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += (int64_t)gridDim.x * blockDim.x) {
        const float4 l = reinterpret_cast<const float4*>(sl)[q], r = reinterpret_cast<const float4*>(sr)[q];
        const float a0 = depth_out_value(l.x * scale, out_kind, div_l), a1 = depth_out_value(l.y * scale, out_kind, div_l);
        const float a2 = depth_out_value(l.z * scale, out_kind, div_l), a3 = depth_out_value(l.w * scale, out_kind, div_l);
        const float b0 = depth_out_value(r.x * scale, out_kind, div_r), b1 = depth_out_value(r.y * scale, out_kind, div_r);
        const float b2 = depth_out_value(r.z * scale, out_kind, div_r), b3 = depth_out_value(r.w * scale, out_kind, div_r);
        st_stream_f4(ol + 3 * q, make_float4(a0, a0, a0, a1), pol);
        st_stream_f4(ol + 3 * q + 1, make_float4(a1, a1, a2, a2), pol);
        st_stream_f4(ol + 3 * q + 2, make_float4(a2, a3, a3, a3), pol);
        st_stream_f4(orr + 3 * q, make_float4(b0, b0, b0, b1), pol);
        st_stream_f4(orr + 3 * q + 1, make_float4(b1, b1, b2, b2), pol);
        st_stream_f4(orr + 3 * q + 2, make_float4(b2, b3, b3, b3), pol);
    }
    {   // scalar path for frames that are not vectorisable
        for (int64_t f = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < npx * 3;
             f += (int64_t)gridDim.x * blockDim.x) {
            int64_t p = f / 3;
            out_l[(int64_t)frame * npx * 3 + f] = depth_out_value(sl[p] * scale, out_kind, div_l);
            out_r[(int64_t)frame * npx * 3 + f] = depth_out_value(sr[p] * scale, out_kind, div_r);
        }
    }
}

cudaError_t launch_depth_out(const float* src_l, const float* src_r, const FrameStats* stats, int n,
                             int h, int w, int src_kind, int out_kind, int group, float* out_l,
                             float* out_r, cudaStream_t s) {
    const int64_t npx = (int64_t)h * w;
    int64_t nquad = (npx + 3) >> 2;
    int bx = (int)((nquad + 255) / 256);
    const int cap = sm_count() * 8 * 2;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    const int vec = (npx % 4 == 0) && ((uintptr_t)out_l % 16 == 0) && ((uintptr_t)out_r % 16 == 0) &&
                    ((uintptr_t)src_l % 16 == 0) && ((uintptr_t)src_r % 16 == 0);
    prof_begin(K_DEPTH_OUT, s);
    k_depth_out<<<dim3(bx, n), 256, 0, s>>>(src_l, src_r, stats, n, npx, src_kind, out_kind,
                                            group < 1 ? 1 : group, vec, out_l, out_r);
    prof_end(K_DEPTH_OUT, s);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------- quantise only (stage API)
__global__ void k_quantize(const float* __restrict__ image, int64_t total_px, uint32_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_px;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = pack_rgbx(quant_u8(image[i * 3]), quant_u8(image[i * 3 + 1]), quant_u8(image[i * 3 + 2]));
}

cudaError_t launch_quantize(const float* image, int64_t total_px, uint32_t* out, cudaStream_t s) {
    int bx = (int)((total_px + 255) / 256);
    if (bx > sm_count() * 16) bx = sm_count() * 16;
    if (bx < 1) bx = 1;
    prof_begin(K_MISC, s);
    k_quantize<<<bx, 256, 0, s>>>(image, total_px, out);
    prof_end(K_MISC, s);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------- shift indices (debug export)
__global__ void k_shift_indices(const float* __restrict__ nd, int64_t total, int w, double div_px,
                                double sep_px, double expo, int kind, int32_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int x = (int)(i % w);
        double off = signed_pow_offset(nd[i], expo, div_px);
        int r;
        if (kind == 0) {
            double t = off + sep_px;
            r = x + (int)t;
        } else {
            double dx = ((double)x + 0.5) + off;
            dx = dx + sep_px;
            r = (int)floor(dx);
        }
        out[i] = r;
    }
}

cudaError_t launch_shift_indices(const float* nd, int n, int h, int w, double div_px, double sep_px,
                                 double expo, int kind, int32_t* out, cudaStream_t s) {
    int64_t total = (int64_t)n * h * w;
    int bx = (int)((total + 255) / 256);
    if (bx > sm_count() * 16) bx = sm_count() * 16;
    if (bx < 1) bx = 1;
    prof_begin(K_MISC, s);
    k_shift_indices<<<bx, 256, 0, s>>>(nd, total, w, div_px, sep_px, expo, kind, out);
    prof_end(K_MISC, s);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cs

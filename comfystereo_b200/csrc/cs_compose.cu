// cs_compose.cu -- C1 + M1 + O1 for the CPU techniques: SBS / top-bottom / anaglyph composition
// of the two uint8 eyes (SIG:1543-1562, overlap_red_cyan SIG:1996-2010), conversion to the
// float32 NHWC tensor the node returns (u8 / 255, GS:365-378) and the "pure black pixel" mask
// (GS:355-361, quirk Q6 -- the mask has the composed image's shape).
//
// Pure streaming kernel: a thread turns 4 output pixels (one 128-bit RGBX8 load per source eye)
// into three 128-bit colour stores and one 128-bit mask store, all with evict-first hints.
// Bytes per OUTPUT pixel: 4 B read (L2-resident scratch), 12 B + 4 B written to HBM.
#include "cs_internal.cuh"

namespace cs {

__device__ __forceinline__ uint32_t compose_px(uint32_t l, uint32_t r, int mode) {
    // anaglyph: R <- first eye, G,B <- second eye
    if (mode == CS_MODE_RED_CYAN) return (l & 0x000000FFu) | (r & 0x00FFFF00u);
    return (r & 0x000000FFu) | (l & 0x00FFFF00u);  // cyan-red reverse
}

__global__ void __launch_bounds__(256) k_compose(const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                                                 int h, int w, int mode, int ho, int wo, int vec_ok,
                                                 float* __restrict__ stereo, float* __restrict__ mask) {
    const int frame = blockIdx.y;
    const int64_t npx_in = (int64_t)h * w, npx_out = (int64_t)ho * wo;
    const uint32_t* L = left + (int64_t)frame * npx_in;
    const uint32_t* R = right + (int64_t)frame * npx_in;
    float* so = stereo + (int64_t)frame * npx_out * 3;
    float* mo = mask + (int64_t)frame * npx_out;
    const uint64_t pol = policy_evict_first();
    const int step = vec_ok ? 4 : 1;
    const int64_t nitem = npx_out / step;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < nitem;
         it += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = it * step;
        const int oy = (int)(o / wo), ox = (int)(o - (int64_t)oy * wo);
        // source eye(s) and source pixel of output pixel (oy, ox)
        const uint32_t* A = L;
        int sy = oy, sx = ox;
        bool both = false;
        switch (mode) {
            case CS_MODE_LEFT_RIGHT: if (ox >= w) { A = R; sx = ox - w; } break;
            case CS_MODE_RIGHT_LEFT: if (ox >= w) { sx = ox - w; } else { A = R; } break;
            case CS_MODE_TOP_BOTTOM: if (oy >= h) { A = R; sy = oy - h; } break;
            case CS_MODE_BOTTOM_TOP: if (oy >= h) { sy = oy - h; } else { A = R; } break;
            case CS_MODE_RED_CYAN: case CS_MODE_CYAN_RED: both = true; break;
            case CS_MODE_LEFT_ONLY: break;
            default: A = R; break;  // only-right
        }
        const int64_t si = (int64_t)sy * w + sx;
        uint32_t p[4];
        if (vec_ok) {
            if (both) {
                uint4 a = *reinterpret_cast<const uint4*>(L + si), b = *reinterpret_cast<const uint4*>(R + si);
                p[0] = compose_px(a.x, b.x, mode); p[1] = compose_px(a.y, b.y, mode);
                p[2] = compose_px(a.z, b.z, mode); p[3] = compose_px(a.w, b.w, mode);
            } else {
                uint4 a = *reinterpret_cast<const uint4*>(A + si);
                p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w;
            }
            float f[12], m[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t r = p[k] & 255u, g = (p[k] >> 8) & 255u, b = (p[k] >> 16) & 255u;
                f[3 * k + 0] = (float)r / 255.0f;
                f[3 * k + 1] = (float)g / 255.0f;
                f[3 * k + 2] = (float)b / 255.0f;
                m[k] = (r + g + b == 0u) ? 1.0f : 0.0f;
            }
            float4* dst = reinterpret_cast<float4*>(so + o * 3);
            st_stream_f4(dst + 0, make_float4(f[0], f[1], f[2], f[3]), pol);
            st_stream_f4(dst + 1, make_float4(f[4], f[5], f[6], f[7]), pol);
            st_stream_f4(dst + 2, make_float4(f[8], f[9], f[10], f[11]), pol);
            st_stream_f4(reinterpret_cast<float4*>(mo + o), make_float4(m[0], m[1], m[2], m[3]), pol);
        } else {
            uint32_t q = both ? compose_px(L[si], R[si], mode) : A[si];
            uint32_t r = q & 255u, g = (q >> 8) & 255u, b = (q >> 16) & 255u;
            so[o * 3 + 0] = (float)r / 255.0f;
            so[o * 3 + 1] = (float)g / 255.0f;
            so[o * 3 + 2] = (float)b / 255.0f;
            mo[o] = (r + g + b == 0u) ? 1.0f : 0.0f;
        }
    }
}

cudaError_t launch_compose(const uint32_t* left, const uint32_t* right, int n, int h, int w, int mode,
                           float* stereo, float* mask, cudaStream_t s) {
    int ho = h, wo = w;
    if (mode == CS_MODE_LEFT_RIGHT || mode == CS_MODE_RIGHT_LEFT) wo = 2 * w;
    if (mode == CS_MODE_TOP_BOTTOM || mode == CS_MODE_BOTTOM_TOP) ho = 2 * h;
    const int64_t npx_out = (int64_t)ho * wo;
    // 4-pixel groups never straddle an eye boundary when w % 4 == 0
    const int vec = (w % 4 == 0) && ((uintptr_t)left % 16 == 0) && ((uintptr_t)right % 16 == 0) &&
                    ((uintptr_t)stereo % 16 == 0) && ((uintptr_t)mask % 16 == 0);
    int64_t nitem = npx_out / (vec ? 4 : 1);
    int bx = (int)((nitem + 255) / 256);
    const int cap = sm_count() * 8 * 2;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    prof_begin(K_COMPOSE, s);
    k_compose<<<dim3(bx, n), 256, 0, s>>>(left, right, h, w, mode, ho, wo, vec, stereo, mask);
    prof_end(K_COMPOSE, s);
    count_launch();
    return cudaGetLastError();
}

// Host transport (cs_host.cu): the depth outputs are three identical channels and the mask is 0/1, so only one
// channel of each depth output and one byte per mask pixel cross PCIe; the host side re-expands them.
template <bool DEPTH_U8>   // CPU techniques: the depth outputs are exactly k / 255, so k (one byte) is all that has to travel
__global__ void __launch_bounds__(256) k_compact_outputs(const float* __restrict__ dl3, const float* __restrict__ dr3,
                                                         const float* __restrict__ mask, int64_t npx, int64_t nmask,
                                                         void* __restrict__ cdl, void* __restrict__ cdr,
                                                         uint8_t* __restrict__ cmask) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (DEPTH_U8) {
        // (k / 255.0f) * 255.0f == k exactly for k = 0..255 (checked in tests), 4 pixels per thread
        uint8_t* bl = (uint8_t*)cdl;
        uint8_t* br = (uint8_t*)cdr;
        const int64_t n4 = npx >> 2;
        for (int64_t i = t0; i < n4; i += stride) {
            uint32_t a = 0, b = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a |= (uint32_t)(__float2int_rn(dl3[3 * (4 * i + j)] * 255.0f) & 255) << (8 * j);
                b |= (uint32_t)(__float2int_rn(dr3[3 * (4 * i + j)] * 255.0f) & 255) << (8 * j);
            }
            reinterpret_cast<uint32_t*>(bl)[i] = a;
            reinterpret_cast<uint32_t*>(br)[i] = b;
        }
        for (int64_t i = (n4 << 2) + t0; i < npx; i += stride) {
            bl[i] = (uint8_t)__float2int_rn(dl3[3 * i] * 255.0f);
            br[i] = (uint8_t)__float2int_rn(dr3[3 * i] * 255.0f);
        }
    } else {
        float* fl = (float*)cdl;
        float* fr = (float*)cdr;
        for (int64_t i = t0; i < npx; i += stride) { fl[i] = dl3[3 * i]; fr[i] = dr3[3 * i]; }
    }
    // 4 mask pixels per thread, tail by the first threads
    const int64_t n4 = nmask >> 2;
    const bool al = ((uintptr_t)mask % 16 == 0) && ((uintptr_t)cmask % 4 == 0);
    if (al) {
        for (int64_t i = t0; i < n4; i += stride) {
            const float4 m = reinterpret_cast<const float4*>(mask)[i];
            const uint32_t b = (m.x != 0.0f ? 1u : 0u) | (m.y != 0.0f ? 1u << 8 : 0u) | (m.z != 0.0f ? 1u << 16 : 0u) |
                               (m.w != 0.0f ? 1u << 24 : 0u);
            reinterpret_cast<uint32_t*>(cmask)[i] = b;
        }
        for (int64_t i = (n4 << 2) + t0; i < nmask; i += stride) cmask[i] = mask[i] != 0.0f ? 1 : 0;
    } else {
        for (int64_t i = t0; i < nmask; i += stride) cmask[i] = mask[i] != 0.0f ? 1 : 0;
    }
}

cudaError_t launch_compact_outputs(const float* dl3, const float* dr3, const float* mask, int64_t npx, int64_t nmask,
                                   int depth_u8, void* cdl, void* cdr, uint8_t* cmask, cudaStream_t s) {
    prof_begin(K_MISC, s);
    if (depth_u8) k_compact_outputs<true><<<sm_count() * 8, 256, 0, s>>>(dl3, dr3, mask, npx, nmask, cdl, cdr, cmask);
    else k_compact_outputs<false><<<sm_count() * 8, 256, 0, s>>>(dl3, dr3, mask, npx, nmask, cdl, cdr, cmask);
    prof_end(K_MISC, s);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cs

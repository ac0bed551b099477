// cs_internal.cuh -- shared device helpers and launcher declarations (sm_100a only).
//
// Arithmetic contract: the library is compiled with -fmad=false, so every float/double
// multiply and add below rounds separately exactly like the reference's numba / numpy /
// torch-CPU elementwise code.  Where a fused multiply-add is part of the defined arithmetic
// (the three blur convolutions, the bilinear sample) it is written explicitly as fmaf().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/comfystereo_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "comfystereo_b200 targets sm_100a (B200) only"
#endif

namespace cs {

constexpr int kMaxBlurRadius = 254;  // dist scratch is uint8: R + 1 must fit

// ---------------------------------------------------------------- per-frame statistics
// Monotone float <-> int encoding so that atomicMin/atomicMax on ints order floats.
__host__ __device__ inline int f2ord(float f) {
#ifdef __CUDA_ARCH__
    int b = __float_as_int(f);
#else
    int b; memcpy(&b, &f, 4);
#endif
    return b >= 0 ? b : (b ^ 0x7FFFFFFF);
}
__host__ __device__ inline float ord2f(int o) {
    int b = o >= 0 ? o : (o ^ 0x7FFFFFFF);
#ifdef __CUDA_ARCH__
    return __int_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

struct FrameStats {     // one per frame in the chunk, lives in the workspace
    int gray_min, gray_max;  // of the gray depth as given (before the x255 decision)
    int l_min, l_max;        // of the depth the LEFT eye warps with (0..255 scale)
    int r_min, r_max;        // same, right eye
    int pad0, pad1;
};

// ---------------------------------------------------------------- streaming memory access
// Inputs/outputs are touched exactly once: keep them from displacing the L2-resident scratch.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld_stream_f4(const float4* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float ld_stream_f1(const float* p, uint64_t pol) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream_f1(float* p, float v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(p), "f"(v), "l"(pol) : "memory");
}

// ---------------------------------------------------------------- bulk asynchronous copies (TMA, 1-D) + mbarrier
// One thread arms the barrier with the byte count and issues the copies; everybody waits on the barrier's phase.
// Source, destination and size must be multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_global), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_min(float v) {
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------- shared arithmetic
// sign_d * (abs(d) ** exponent) * divergence_px in float64, sign(0) = +1 (e.g. SIG:1863-1865).
// pow(x, 2) and pow(x, 1) are exact in libm; they are special-cased so that the two common
// exponents are bit-identical to the reference; other exponents use CUDA's pow (<= 2 ulp).
__device__ __forceinline__ double signed_pow_offset(float nd, double expo, double div_px) {
    double a = (double)fabsf(nd);
    double p;
    if (expo == 2.0) p = a * a;
    else if (expo == 1.0) p = a;
    else p = pow(a, expo);
    double sp = (nd >= 0.0f) ? p : -p;
    return sp * div_px;
}

// nd = (d - min) / (max - min) - conv  (numpy float32, SIG:1591-1600); flat frame -> 0 - conv.
struct Normalizer {
    float mn, range, conv;
    int flat;
    __device__ __forceinline__ float operator()(float d) const {
        float q = flat ? 0.0f : (d - mn) / range;
        return q - conv;
    }
};
__device__ __forceinline__ Normalizer make_normalizer(int ord_min, int ord_max, float scale, float conv) {
    // min/max are tracked on the unscaled values; x*255 is monotone so scaling commutes.
    Normalizer n;
    float lo = ord2f(ord_min) * scale, hi = ord2f(ord_max) * scale;
    n.mn = lo; n.range = hi - lo; n.conv = conv; n.flat = (hi == lo);
    return n;
}

// uint8 pixel packed RGBX in one 32-bit word (R lowest byte).
__device__ __forceinline__ uint32_t pack_rgbx(int r, int g, int b, int x = 0) {
    return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | ((uint32_t)x << 24);
}

// ---------------------------------------------------------------- launchers (host)

struct EyeSpec {          // one eye of one call
    double div_px, sep_px; // signed, in pixels (SIG:1602-1603)
    int passthrough;       // divergence*(1 +- balance) < 0.001 -> the eye is the input image
};

struct WarpArgs {
    const uint32_t* image_u8;   // [n][h][w] RGBX
    const float* depth[2];      // per eye [n][h][w]; unscaled gray or blurred (see scale[])
    const FrameStats* stats;    // [n]
    int use_blur_stats;         // 1: l_/r_ min/max, 0: gray min/max
    int scale_by_stats;         // 1: multiply depth by 255 when the frame's gray max <= 1 (blur off path)
    uint32_t* out[2];           // per eye [n][h][w] RGBX (X = "filled" flag where meaningful)
    int n, h, w;
    int fill;
    EyeSpec eye[2];
    double expo;
    float conv;
    void* scratch;              // technique-specific device scratch
    size_t scratch_bytes;
    int flags;                  // bit 0: polylines -- skip the fast sweep, replay every row exactly (tests)
    // Polylines only, side-by-side / top-bottom modes: the sweep writes the composed float32 tensors itself (C1 + M1 + O1,
    // SIG:1543-1552, GS:355-378) instead of the RGBX8 eye image that k_compose would read back.  nullptr: classic output.
    float* fused_stereo;        // [n][ho][wo][3]
    float* fused_mask;          // [n][ho][wo]
    int fused_mode;             // CS_MODE_LEFT_RIGHT .. CS_MODE_BOTTOM_TOP
};

// composed position of eye pixel (frame, y, x): index of the pixel in the [n][ho][wo] output
__device__ __forceinline__ int64_t fused_index(const WarpArgs& a, int eye, int frame, int y, int x) {
    const bool sbs = a.fused_mode == CS_MODE_LEFT_RIGHT || a.fused_mode == CS_MODE_RIGHT_LEFT;
    const bool second = (a.fused_mode == CS_MODE_LEFT_RIGHT || a.fused_mode == CS_MODE_TOP_BOTTOM) ? (eye == 1) : (eye == 0);
    const int wo = sbs ? 2 * a.w : a.w, ho = sbs ? a.h : 2 * a.h;
    const int ox = x + ((sbs && second) ? a.w : 0), oy = y + ((!sbs && second) ? a.h : 0);
    return ((int64_t)frame * ho + oy) * wo + ox;
}

// k / 255.0f for k = 0..255, the IEEE float32 quotients (the dequantisation of GS:365-378): kernels copy the table into
// shared memory instead of dividing 256 times per CTA (the division was 4 % of k_warp_rows' instructions).
static __device__ const float kQ255[256] = {
    0x0.0p+0f, 0x1.0101020000000p-8f, 0x1.0101020000000p-7f, 0x1.8181820000000p-7f, 0x1.0101020000000p-6f, 0x1.4141420000000p-6f, 0x1.8181820000000p-6f, 0x1.c1c1c20000000p-6f,
    0x1.0101020000000p-5f, 0x1.2121220000000p-5f, 0x1.4141420000000p-5f, 0x1.6161620000000p-5f, 0x1.8181820000000p-5f, 0x1.a1a1a20000000p-5f, 0x1.c1c1c20000000p-5f, 0x1.e1e1e20000000p-5f,
    0x1.0101020000000p-4f, 0x1.1111120000000p-4f, 0x1.2121220000000p-4f, 0x1.3131320000000p-4f, 0x1.4141420000000p-4f, 0x1.5151520000000p-4f, 0x1.6161620000000p-4f, 0x1.7171720000000p-4f,
    0x1.8181820000000p-4f, 0x1.9191920000000p-4f, 0x1.a1a1a20000000p-4f, 0x1.b1b1b20000000p-4f, 0x1.c1c1c20000000p-4f, 0x1.d1d1d20000000p-4f, 0x1.e1e1e20000000p-4f, 0x1.f1f1f20000000p-4f,
    0x1.0101020000000p-3f, 0x1.09090a0000000p-3f, 0x1.1111120000000p-3f, 0x1.19191a0000000p-3f, 0x1.2121220000000p-3f, 0x1.29292a0000000p-3f, 0x1.3131320000000p-3f, 0x1.39393a0000000p-3f,
    0x1.4141420000000p-3f, 0x1.49494a0000000p-3f, 0x1.5151520000000p-3f, 0x1.59595a0000000p-3f, 0x1.6161620000000p-3f, 0x1.69696a0000000p-3f, 0x1.7171720000000p-3f, 0x1.79797a0000000p-3f,
    0x1.8181820000000p-3f, 0x1.89898a0000000p-3f, 0x1.9191920000000p-3f, 0x1.99999a0000000p-3f, 0x1.a1a1a20000000p-3f, 0x1.a9a9aa0000000p-3f, 0x1.b1b1b20000000p-3f, 0x1.b9b9ba0000000p-3f,
    0x1.c1c1c20000000p-3f, 0x1.c9c9ca0000000p-3f, 0x1.d1d1d20000000p-3f, 0x1.d9d9da0000000p-3f, 0x1.e1e1e20000000p-3f, 0x1.e9e9ea0000000p-3f, 0x1.f1f1f20000000p-3f, 0x1.f9f9fa0000000p-3f,
    0x1.0101020000000p-2f, 0x1.0505060000000p-2f, 0x1.09090a0000000p-2f, 0x1.0d0d0e0000000p-2f, 0x1.1111120000000p-2f, 0x1.1515160000000p-2f, 0x1.19191a0000000p-2f, 0x1.1d1d1e0000000p-2f,
    0x1.2121220000000p-2f, 0x1.2525260000000p-2f, 0x1.29292a0000000p-2f, 0x1.2d2d2e0000000p-2f, 0x1.3131320000000p-2f, 0x1.3535360000000p-2f, 0x1.39393a0000000p-2f, 0x1.3d3d3e0000000p-2f,
    0x1.4141420000000p-2f, 0x1.4545460000000p-2f, 0x1.49494a0000000p-2f, 0x1.4d4d4e0000000p-2f, 0x1.5151520000000p-2f, 0x1.5555560000000p-2f, 0x1.59595a0000000p-2f, 0x1.5d5d5e0000000p-2f,
    0x1.6161620000000p-2f, 0x1.6565660000000p-2f, 0x1.69696a0000000p-2f, 0x1.6d6d6e0000000p-2f, 0x1.7171720000000p-2f, 0x1.7575760000000p-2f, 0x1.79797a0000000p-2f, 0x1.7d7d7e0000000p-2f,
    0x1.8181820000000p-2f, 0x1.8585860000000p-2f, 0x1.89898a0000000p-2f, 0x1.8d8d8e0000000p-2f, 0x1.9191920000000p-2f, 0x1.9595960000000p-2f, 0x1.99999a0000000p-2f, 0x1.9d9d9e0000000p-2f,
    0x1.a1a1a20000000p-2f, 0x1.a5a5a60000000p-2f, 0x1.a9a9aa0000000p-2f, 0x1.adadae0000000p-2f, 0x1.b1b1b20000000p-2f, 0x1.b5b5b60000000p-2f, 0x1.b9b9ba0000000p-2f, 0x1.bdbdbe0000000p-2f,
    0x1.c1c1c20000000p-2f, 0x1.c5c5c60000000p-2f, 0x1.c9c9ca0000000p-2f, 0x1.cdcdce0000000p-2f, 0x1.d1d1d20000000p-2f, 0x1.d5d5d60000000p-2f, 0x1.d9d9da0000000p-2f, 0x1.ddddde0000000p-2f,
    0x1.e1e1e20000000p-2f, 0x1.e5e5e60000000p-2f, 0x1.e9e9ea0000000p-2f, 0x1.ededee0000000p-2f, 0x1.f1f1f20000000p-2f, 0x1.f5f5f60000000p-2f, 0x1.f9f9fa0000000p-2f, 0x1.fdfdfe0000000p-2f,
    0x1.0101020000000p-1f, 0x1.0303040000000p-1f, 0x1.0505060000000p-1f, 0x1.0707080000000p-1f, 0x1.09090a0000000p-1f, 0x1.0b0b0c0000000p-1f, 0x1.0d0d0e0000000p-1f, 0x1.0f0f100000000p-1f,
    0x1.1111120000000p-1f, 0x1.1313140000000p-1f, 0x1.1515160000000p-1f, 0x1.1717180000000p-1f, 0x1.19191a0000000p-1f, 0x1.1b1b1c0000000p-1f, 0x1.1d1d1e0000000p-1f, 0x1.1f1f200000000p-1f,
    0x1.2121220000000p-1f, 0x1.2323240000000p-1f, 0x1.2525260000000p-1f, 0x1.2727280000000p-1f, 0x1.29292a0000000p-1f, 0x1.2b2b2c0000000p-1f, 0x1.2d2d2e0000000p-1f, 0x1.2f2f300000000p-1f,
    0x1.3131320000000p-1f, 0x1.3333340000000p-1f, 0x1.3535360000000p-1f, 0x1.3737380000000p-1f, 0x1.39393a0000000p-1f, 0x1.3b3b3c0000000p-1f, 0x1.3d3d3e0000000p-1f, 0x1.3f3f400000000p-1f,
    0x1.4141420000000p-1f, 0x1.4343440000000p-1f, 0x1.4545460000000p-1f, 0x1.4747480000000p-1f, 0x1.49494a0000000p-1f, 0x1.4b4b4c0000000p-1f, 0x1.4d4d4e0000000p-1f, 0x1.4f4f500000000p-1f,
    0x1.5151520000000p-1f, 0x1.5353540000000p-1f, 0x1.5555560000000p-1f, 0x1.5757580000000p-1f, 0x1.59595a0000000p-1f, 0x1.5b5b5c0000000p-1f, 0x1.5d5d5e0000000p-1f, 0x1.5f5f600000000p-1f,
    0x1.6161620000000p-1f, 0x1.6363640000000p-1f, 0x1.6565660000000p-1f, 0x1.6767680000000p-1f, 0x1.69696a0000000p-1f, 0x1.6b6b6c0000000p-1f, 0x1.6d6d6e0000000p-1f, 0x1.6f6f700000000p-1f,
    0x1.7171720000000p-1f, 0x1.7373740000000p-1f, 0x1.7575760000000p-1f, 0x1.7777780000000p-1f, 0x1.79797a0000000p-1f, 0x1.7b7b7c0000000p-1f, 0x1.7d7d7e0000000p-1f, 0x1.7f7f800000000p-1f,
    0x1.8181820000000p-1f, 0x1.8383840000000p-1f, 0x1.8585860000000p-1f, 0x1.8787880000000p-1f, 0x1.89898a0000000p-1f, 0x1.8b8b8c0000000p-1f, 0x1.8d8d8e0000000p-1f, 0x1.8f8f900000000p-1f,
    0x1.9191920000000p-1f, 0x1.9393940000000p-1f, 0x1.9595960000000p-1f, 0x1.9797980000000p-1f, 0x1.99999a0000000p-1f, 0x1.9b9b9c0000000p-1f, 0x1.9d9d9e0000000p-1f, 0x1.9f9fa00000000p-1f,
    0x1.a1a1a20000000p-1f, 0x1.a3a3a40000000p-1f, 0x1.a5a5a60000000p-1f, 0x1.a7a7a80000000p-1f, 0x1.a9a9aa0000000p-1f, 0x1.ababac0000000p-1f, 0x1.adadae0000000p-1f, 0x1.afafb00000000p-1f,
    0x1.b1b1b20000000p-1f, 0x1.b3b3b40000000p-1f, 0x1.b5b5b60000000p-1f, 0x1.b7b7b80000000p-1f, 0x1.b9b9ba0000000p-1f, 0x1.bbbbbc0000000p-1f, 0x1.bdbdbe0000000p-1f, 0x1.bfbfc00000000p-1f,
    0x1.c1c1c20000000p-1f, 0x1.c3c3c40000000p-1f, 0x1.c5c5c60000000p-1f, 0x1.c7c7c80000000p-1f, 0x1.c9c9ca0000000p-1f, 0x1.cbcbcc0000000p-1f, 0x1.cdcdce0000000p-1f, 0x1.cfcfd00000000p-1f,
    0x1.d1d1d20000000p-1f, 0x1.d3d3d40000000p-1f, 0x1.d5d5d60000000p-1f, 0x1.d7d7d80000000p-1f, 0x1.d9d9da0000000p-1f, 0x1.dbdbdc0000000p-1f, 0x1.ddddde0000000p-1f, 0x1.dfdfe00000000p-1f,
    0x1.e1e1e20000000p-1f, 0x1.e3e3e40000000p-1f, 0x1.e5e5e60000000p-1f, 0x1.e7e7e80000000p-1f, 0x1.e9e9ea0000000p-1f, 0x1.ebebec0000000p-1f, 0x1.ededee0000000p-1f, 0x1.efeff00000000p-1f,
    0x1.f1f1f20000000p-1f, 0x1.f3f3f40000000p-1f, 0x1.f5f5f60000000p-1f, 0x1.f7f7f80000000p-1f, 0x1.f9f9fa0000000p-1f, 0x1.fbfbfc0000000p-1f, 0x1.fdfdfe0000000p-1f, 0x1.0000000000000p+0f
};

void count_launch();
int sm_count();           // SMs of the current device (148 on a B200), for grid sizing
void release_graphs();   // cs_api.cu: drops the cached CUDA graphs of small repeated cs_stereo_batch calls
// Optional per-kernel timing (bench.py): CUDA events recorded on the launch stream around every kernel.
enum KernelId { K_PREPARE = 0, K_EDGE_DIST, K_BLUR_BLEND, K_DEPTH_OUT, K_WARP_ROWS, K_POLY_FAST, K_POLY_EXACT,
                K_HYBRID_SPLAT, K_HYBRID_GAPFILL, K_GPUWARP, K_COMPOSE, K_MISC, K_RESIZE, K_COUNT };
void prof_begin(int id, cudaStream_t s);
void prof_end(int id, cudaStream_t s);
int fail(int code, const char* fmt, ...);   // records the thread-local cs_last_error() text, returns code
cudaError_t launch_init_stats(FrameStats* stats, int n, cudaStream_t s);
// N1: gray + bilinear resize of a [n][dh][dw][c] depth to [n][h][w] (torch CPU arithmetic, strict float32)
cudaError_t launch_resize_gray(const float* depth, int n, int dh, int dw, int c, int h, int w, float* out,
                               cudaStream_t s);
cudaError_t launch_prepare(const float* image, const float* depth, int n, int h, int w, int c,
                           float* gray, uint32_t* image_u8, FrameStats* stats, cudaStream_t s);
// scale_mode 0: input already 0..255; 1: x255 when the frame's gray max <= 1 (SIG:1475);
//            2: same test over the frame's sub-batch of `group` frames (SIG:1045)
// depth_l_out/depth_r_out (optional): the CPU-technique depth outputs (wrap quirk Q1), [n][h][w][3]
void set_blur_test_flags(int flags);   // bit 0: scalar k_edge_dist (tests)
cudaError_t launch_blur(const float* gray, FrameStats* stats, int scale_mode, int group,
                        int n, int h, int w, const cs_params& p, float* blur_l, float* blur_r,
                        uint8_t* dist, float* depth_l_out, float* depth_r_out, cudaStream_t s);
cudaError_t launch_depth_out(const float* src_l, const float* src_r, const FrameStats* stats, int n,
                             int h, int w, int src_kind, int out_kind, int group, float* out_l,
                             float* out_r, cudaStream_t s);
cudaError_t launch_quantize(const float* image, int64_t total_px, uint32_t* out, cudaStream_t s);
cudaError_t launch_shift_indices(const float* nd, int n, int h, int w, double div_px, double sep_px,
                                 double expo, int kind, int32_t* out, cudaStream_t s);
cudaError_t launch_warp_rows(const WarpArgs& a, cudaStream_t s);       // none / naive / interp / inverse
cudaError_t launch_polylines(const WarpArgs& a, cudaStream_t s);       // soft / sharp
cudaError_t launch_hybrid(const WarpArgs& a, cudaStream_t s);          // hybrid_edge (2 kernels)
size_t polylines_scratch_bytes(int n, int h, int w);
size_t hybrid_plus_scratch_bytes(int n, int h, int w);
cudaError_t launch_hybrid_plus(const WarpArgs& a, cudaStream_t s);     // hybrid_edge_plus (hybrid + polylines_soft + merge)
cudaError_t launch_minmax(const float* src, int n, int64_t npx, FrameStats* stats, cudaStream_t s);
cudaError_t launch_export_stats(const FrameStats* stats, int n, int which, float* out, int stride, cudaStream_t s);
cudaError_t launch_compose(const uint32_t* left, const uint32_t* right, int n, int h, int w, int mode,
                           float* stereo, float* mask, cudaStream_t s);

struct GpuWarpArgs {
    const float* image;        // [n][h][w][3] float32
    const float* depth[2];     // per eye [n][h][w], 0..255 scale or as given
    const FrameStats* stats;
    int use_blur_stats;        // 1: depth[] are the blurred maps (0..255 scale), stats l_/r_ fields
    int prescale;              // use_blur_stats == 0 only.  1: depth *= 255 when the sub-batch max <= 1 (SIG:1045)
                               //                            0: depth used as given (function-level call)
    int group;                 // sub-batch size for the coupled range tests (Q9)
    int n, h, w, mode;
    EyeSpec eye[2];
    float expo, conv;
    float* stereo;             // final layout
    float* mask;               // [n][h][w]
    uint8_t* keep;             // mesh warp only: mesh_keep_bytes(n, h, w) of scratch for the culled topology
    void* row_scratch;         // rows too wide for shared memory only: gpuwarp_row_scratch_bytes(w) of global scratch ...
    size_t row_scratch_stride; // ... and the bytes of one CTA's slice (gpuwarp_row_scratch_stride(w))
};
size_t gpuwarp_row_scratch_bytes(int w);      // 0 while a row's state fits a CTA's shared memory (up to ~9200 columns)
size_t gpuwarp_row_scratch_stride(int w);
cudaError_t launch_gpuwarp(const GpuWarpArgs& a, cudaStream_t s);
// forward_warp_mesh (SIG:453-689): same arguments, plus the keep scratch
size_t mesh_keep_bytes(int n, int h, int w);
cudaError_t launch_meshwarp(const GpuWarpArgs& a, cudaStream_t s);
// one channel of each depth output + one byte per mask pixel, for the host transport (cs_host.cu)
// (depth_u8: the depth outputs are k / 255 -- every CPU technique -- and travel as the byte k)
cudaError_t launch_compact_outputs(const float* dl3, const float* dr3, const float* mask, int64_t npx, int64_t nmask,
                                   int depth_u8, void* cdl, void* cdr, uint8_t* cmask, cudaStream_t s);

}  // namespace cs


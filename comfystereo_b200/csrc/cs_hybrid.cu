// cs_hybrid.cu -- "Imperfect fill - Hybrid Edge":
//   H1  enhanced_inverse_mapping_with_mask (SIG:1622-1661): 3-tap Gaussian forward splat
//   H2  rgb2gray (SIG:1740-1742) + edge_aware_gap_fill (SIG:1745-1774): 3x3 joint-bilateral fill
//
// H1, one CTA per (row, frame, eye).  The reference scatters in ascending source order into
// float32 accumulators, so the float32 rounding sequence of every destination column is defined
// by that order.  The kernel turns the scatter into a gather: each destination thread walks the
// (bounded) window of source columns that can reach it, in ascending order, and applies exactly
// the contributions the reference would have applied -- same order, same float64 products, same
// float32 re-rounding -- without atomics and therefore deterministically.
// H2 is a per-pixel stencil on the H1 result and runs in place: it only writes pixels whose mask
// is 0 and only reads colours of pixels whose mask is 1.
//
// Bytes per pixel and eye: H1 reads depth 4 B + RGBX8 4 B, writes RGBX8 4 B; H2 re-reads the
// 4 B (+ the 3x3 neighbourhood from L1/L2) and rewrites only hole pixels.
#include "cs_internal.cuh"

namespace cs {

__device__ __forceinline__ Normalizer hy_normalizer(const WarpArgs& a, int eye, int frame, float* scale_out) {
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

// float64 -> nearest float32 value, kept as float64 (Veltkamp split, 2^29 + 1), and uint8 -> float64 through the
// 2^52 bias: the float32 accumulators of the reference are carried in FP64 registers without F2F / I2F conversions,
// which issue on the 16-lane XU pipe (52 % busy in this kernel before the change).  See cs_polylines.cu.
__device__ __forceinline__ double hy_round24(double x) {
    const double g = x * 536870913.0;
    const double d = x - g;
    return g + d;
}
__device__ __forceinline__ double hy_u8(uint32_t v) {
    return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}

// exp(-t) for t in [0, 2] -- all the Gaussian splat ever asks for (|diff| < 2, below): t = i/64 + r with |r| <= 1/128,
// exp(-i/64) from a table of correctly rounded values, exp(-r) from its degree-6 Taylor polynomial (truncation 3.5e-19).
// About 1 ulp like the library exp, in a quarter of its instructions (exp was a quarter of this kernel's instruction
// count -- profiles/r02_ncu_kernels.md -- and the kernel issues 87 percent of its cycles).
__device__ const double kExpNegTab[129] = {
    0x1.0000000000000p+0, 0x1.f80feabfeefa5p-1, 0x1.f03f56a88b5d8p-1, 0x1.e88dc6afecfc0p-1,
    0x1.e0fabfbc702a4p-1, 0x1.d985c89d041a3p-1, 0x1.d22e6a0197c03p-1, 0x1.caf42e73a4c7ep-1,
    0x1.c3d6a24ed8222p-1, 0x1.bcd553b9d7b62p-1, 0x1.b5efd29f24c26p-1, 0x1.af25b0a61a7b5p-1,
    0x1.a876812c0877cp-1, 0x1.a1e1d93d687d0p-1, 0x1.9b674f8f2f3d8p-1, 0x1.95067c78379f2p-1,
    0x1.8ebef9eac820bp-1, 0x1.8890636e31f54p-1, 0x1.827a561889716p-1, 0x1.7c7c70887763cp-1,
    0x1.769652df22f7ep-1, 0x1.70c79eba33c07p-1, 0x1.6b0ff72deb89dp-1, 0x1.656f00bf5796ap-1,
    0x1.5fe4615e98e8fp-1, 0x1.5a6fc061433c8p-1, 0x1.5510c67cd2591p-1, 0x1.4fc71dc135627p-1,
    0x1.4a9271936fd09p-1, 0x1.45726ea84fb88p-1, 0x1.4066c2ff39127p-1, 0x1.3b6f1ddd05a92p-1,
    0x1.368b2fc6f960ap-1, 0x1.31baaa7dca843p-1, 0x1.2cfd40f8bdcaep-1, 0x1.2852a760d5c59p-1,
    0x1.23ba930c1568bp-1, 0x1.1f34ba78d5666p-1, 0x1.1ac0d5492c0dcp-1, 0x1.165e9c3e67663p-1,
    0x1.120dc934993e8p-1, 0x1.0dce171e34e7fp-1, 0x1.099f41ffbe580p-1, 0x1.058106eb8a6aap-1,
    0x1.017323fd90020p-1, 0x1.faeab0ae9381dp-2, 0x1.f30ec8375038bp-2, 0x1.eb5210d6270b7p-2,
    0x1.e3b40ebefcd7ep-2, 0x1.dc3448110daaep-2, 0x1.d4d244cf4ea9ep-2, 0x1.cd8d8ed8ee386p-2,
    0x1.c665b1e1f1e0dp-2, 0x1.bf5a3b6bf18b7p-2, 0x1.b86ababeef8dfp-2, 0x1.b196c0e24d229p-2,
    0x1.aadde095dad4bp-2, 0x1.a43fae4b0474ep-2, 0x1.9dbbc01e18268p-2, 0x1.9751adcfa81c1p-2,
    0x1.910110be06976p-2, 0x1.8ac983dedbc65p-2, 0x1.84aaa3b8d514ep-2, 0x1.7ea40e5d6d8fap-2,
    0x1.78b56362cef38p-2, 0x1.72de43ddcb07fp-2, 0x1.6d1e525bece49p-2, 0x1.677532dda1c1cp-2,
    0x1.61e28ad078f7fp-2, 0x1.5c6601097ad0fp-2, 0x1.56ff3dbf95d14p-2, 0x1.51adea86221f4p-2,
    0x1.4c71b2477ab20p-2, 0x1.474a413fabef5p-2, 0x1.423744f737661p-2, 0x1.3d386c3dec4f1p-2,
    0x1.384d6725d4833p-2, 0x1.3375e6fe3595dp-2, 0x1.2eb19e4ea5c20p-2, 0x1.2a0040d2345ddp-2,
    0x1.25618372a584fp-2, 0x1.20d51c43c0ae6p-2, 0x1.1c5ac27eb1e31p-2, 0x1.17f22e7d7d4a6p-2,
    0x1.139b19b684c48p-2, 0x1.0f553eb81f4abp-2, 0x1.0b20592441cecp-2, 0x1.06fc25ac3954ep-2,
    0x1.02e8620c7602cp-2, 0x1.fdc99a10cdc21p-3, 0x1.f5e24ccccc19ap-3, 0x1.ee1a5dd76a300p-3,
    0x1.e67150b112b02p-3, 0x1.dee6aac84fc87p-3, 0x1.d779f37222093p-3, 0x1.d02ab3e275aabp-3,
    0x1.c8f87724b5c1dp-3, 0x1.c1e2ca147ced4p-3, 0x1.bae93b5663055p-3, 0x1.b40b5b50e75c2p-3,
    0x1.ad48bc25771c7p-3, 0x1.a6a0f1a98f570p-3, 0x1.a013915ffa516p-3, 0x1.99a0327227aa0p-3,
    0x1.93466da99ee64p-3, 0x1.8d05dd698c022p-3, 0x1.86de1da8659adp-3, 0x1.80cecbe9ac4dbp-3,
    0x1.7ad78737c2e82p-3, 0x1.74f7f01ddf059p-3, 0x1.6f2fa8a211badp-3, 0x1.697e543f67ef8p-3,
    0x1.63e397e022073p-3, 0x1.5e5f19d8027d6p-3, 0x1.58f081deb31aep-3, 0x1.5397790a40685p-3,
    0x1.4e53a9c9ab086p-3, 0x1.4924bfdf8ea01p-3, 0x1.440a685cddfa0p-3, 0x1.3f04519bb40e3p-3,
    0x1.3a122b3a399d5p-3, 0x1.3533a6159f0c4p-3, 0x1.306874452a30bp-3, 0x1.2bb0491557bf0p-3,
    0x1.270ad903100b9p-3, 0x1.2277d9b6eed30p-3, 0x1.1df702009dbdcp-3, 0x1.198809d241548p-3,
    0x1.152aaa3bf81ccp-3
};
__device__ __forceinline__ double hy_exp_neg(double t, const double* __restrict__ tab) {
    const double magic = 6755399441055744.0;                 // 2^52 + 2^51: the sum's low word is round(64 t)
    const double kf = fma(t, 64.0, magic);
    int i = __double2loint(kf);
    i = min(max(i, 0), 128);                                  // NaN / out-of-range input: any entry, the NaN propagates below
    const double s = fma(kf - magic, 0.015625, -t);           // -(t - i/64) = -r
    double q = fma(s, 1.0 / 720.0, 1.0 / 120.0);
    q = fma(q, s, 1.0 / 24.0);
    q = fma(q, s, 1.0 / 6.0);
    q = fma(q, s, 0.5);
    q = fma(q, s, 1.0);
    q = q * s;                                                // exp(-r) - 1
    const double T = tab[i];
    return fma(T, q, T);
}

template <int TPB>   // CTA size the kernel is compiled for: 256, or 512 for very wide rows (see launch_warp_rows)
__global__ void __launch_bounds__(TPB) k_hybrid_splat(const WarpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int w = a.w, y = blockIdx.x, frame = blockIdx.y, eye = blockIdx.z;
    if (a.eye[eye].passthrough) return;
    double* dxs = reinterpret_cast<double*>(smem_raw);          // [w] destination x of each source column
    int* jcs = reinterpret_cast<int*>(dxs + w);                 // [w] floor(dest_x)
    const int nblk = (w + 31) >> 5;
    int* bmin = jcs + w;                                        // [nblk] min / max of (jc - x) per 32 source columns
    int* bmax = bmin + nblk;
    __shared__ int s_omin, s_omax;
    __shared__ double s_exp[129];
    if (threadIdx.x < 129) s_exp[threadIdx.x] = kExpNegTab[threadIdx.x];
    if (threadIdx.x == 0) { s_omin = 0x7FFFFFFF; s_omax = (int)0x80000000; }
    const double div_px = a.eye[eye].div_px, sep_px = a.eye[eye].sep_px;
    float scale;
    const Normalizer norm = hy_normalizer(a, eye, frame, &scale);
    const int64_t row_off = (int64_t)frame * a.h * w + (int64_t)y * w;
    const float* dep = a.depth[eye] + row_off;
    const uint32_t* img = a.image_u8 + row_off;
    uint32_t* out = a.out[eye] + row_off;
    __syncthreads();
    int omin = 0x7FFFFFFF, omax = (int)0x80000000;
    const int wpad = nblk << 5;
    for (int x = threadIdx.x; x < wpad; x += blockDim.x) {   // padded to whole warps: the block ranges use full-mask shuffles
        int lo = 0x7FFFFFFF, hi = (int)0x80000000;
        if (x < w) {
            float d = dep[x];
            if (scale != 1.0f) d = d * scale;
            double off = signed_pow_offset(norm(d), a.expo, div_px);
            double dx = ((double)x + 0.5) + off;
            dx = dx + sep_px;
            double fl = floor(dx);
            // keep the index sane for absurd parameters; such columns can never be on screen
            int jc = (fl < -1.0e9) ? -1000000000 : ((fl > 1.0e9) ? 1000000000 : (int)fl);
            dxs[x] = dx;
            jcs[x] = jc;
            int o = jc - x;
            omin = min(omin, o); omax = max(omax, o);
            lo = o; hi = o;
        }
        // a warp covers 32 consecutive source columns here: their offset range, for the per-destination windows below
        for (int q = 16; q; q >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, q));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, q));
        }
        if ((threadIdx.x & 31) == 0) { bmin[x >> 5] = lo; bmax[x >> 5] = hi; }
    }
    for (int o = 16; o; o >>= 1) {
        omin = min(omin, __shfl_xor_sync(0xffffffffu, omin, o));
        omax = max(omax, __shfl_xor_sync(0xffffffffu, omax, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_omin, omin); atomicMax(&s_omax, omax); }
    __syncthreads();
    omin = s_omin; omax = s_omax;
    for (int j = threadIdx.x; j < w; j += blockDim.x) {
        // sources with jc in {j-1, j, j+1}:  x = jc - (jc - x)  lies in [j-1-omax, j+1-omin]
        long long lo = (long long)j - 1 - omax, hi = (long long)j + 1 - omin;
        int x0 = (int)max(lo, 0ll), x1 = (int)min(hi, (long long)w - 1);
        if (x0 <= x1) {
            // tighten with the offset range of the 32-column blocks the row-wide window touches
            int lmin = 0x7FFFFFFF, lmax = (int)0x80000000;
            for (int b = x0 >> 5; b <= (x1 >> 5); ++b) { lmin = min(lmin, bmin[b]); lmax = max(lmax, bmax[b]); }
            x0 = max(x0, (int)max((long long)j - 1 - lmax, 0ll));
            x1 = min(x1, (int)min((long long)j + 1 - lmin, (long long)w - 1));
        }
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, ws = 0.0;   // float32-valued
        bool hit = false;
        for (int x = x0; x <= x1; ++x) {
            int dj = j - jcs[x];
            if (dj < -1 || dj > 1) continue;
            double diff = dxs[x] - (double)j;
            double wg = hy_exp_neg((diff * diff) / 2.0, s_exp);
            uint32_t p = img[x];
            acc0 = hy_round24(acc0 + hy_u8(p & 255u) * wg);
            acc1 = hy_round24(acc1 + hy_u8((p >> 8) & 255u) * wg);
            acc2 = hy_round24(acc2 + hy_u8((p >> 16) & 255u) * wg);
            ws = hy_round24(ws + wg);
            hit = true;
        }
        uint32_t px = 0;
        if (ws > 0.0) {
            const float wsf = (float)ws;
            float v0 = fminf(fmaxf((float)acc0 / wsf, 0.0f), 255.0f);
            float v1 = fminf(fmaxf((float)acc1 / wsf, 0.0f), 255.0f);
            float v2 = fminf(fmaxf((float)acc2 / wsf, 0.0f), 255.0f);
            px = pack_rgbx((int)v0, (int)v1, (int)v2);
        }
        out[j] = px | (hit ? 0x01000000u : 0u);
    }
}

__device__ __forceinline__ double guidance(uint32_t p) {  // rgb2gray, float64, SIG:1740-1742
    double a = 0.299 * (double)(p & 255u), b = 0.587 * (double)((p >> 8) & 255u), c = 0.114 * (double)((p >> 16) & 255u);
    double s = a + b;
    return s + c;
}

// the 3x3 joint-bilateral fill of one hole pixel (x, y); `self` when no neighbour is filled
__device__ __forceinline__ uint32_t hy_fill_pixel(const uint32_t* __restrict__ img, const uint32_t* __restrict__ orig,
                                                  int w, int h, int x, int y, uint32_t self, double ws1, double ws2) {
    const double g0 = guidance(orig[(int64_t)y * w + x]);
    float nv0 = 0.0f, nv1 = 0.0f, nv2 = 0.0f;
    double wt = 0.0;
    for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
            int ny = y + di, nx = x + dj;
            if (ny < 0 || ny >= h || nx < 0 || nx >= w) continue;
            uint32_t q = img[(int64_t)ny * w + nx];
            if (!(q >> 24)) continue;
            int dsq = di * di + dj * dj;
            double w_s = (dsq == 1) ? ws1 : ws2;  // exp(-dsq/2), evaluated on the host
            double diff = g0 - guidance(orig[(int64_t)ny * w + nx]);
            double w_r = exp(-(diff * diff) / 200.0);
            double wg = w_s * w_r;
            float wf = (float)wg;
            nv0 = nv0 + (float)(q & 255u) * wf;
            nv1 = nv1 + (float)((q >> 8) & 255u) * wf;
            nv2 = nv2 + (float)((q >> 16) & 255u) * wf;
            wt += wg;
        }
    if (!(wt > 0.0)) return self;
    float wtf = (float)wt;
    float v0 = fminf(fmaxf(nv0 / wtf, 0.0f), 255.0f);
    float v1 = fminf(fmaxf(nv1 / wtf, 0.0f), 255.0f);
    float v2 = fminf(fmaxf(nv2 / wtf, 0.0f), 255.0f);
    return pack_rgbx((int)v0, (int)v1, (int)v2);  // X stays 0: still "mask == 0" for the neighbours
}

__global__ void __launch_bounds__(256) k_hybrid_gapfill(const WarpArgs a, double ws1, double ws2) {
    const int w = a.w, h = a.h, frame = blockIdx.y, eye = blockIdx.z;
    if (a.eye[eye].passthrough) return;
    const int64_t base_off = (int64_t)frame * h * w;
    const uint32_t* orig = a.image_u8 + base_off;
    uint32_t* img = a.out[eye] + base_off;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)h * w;
         i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t self = img[i];
        if (self >> 24) continue;  // mask != 0: keep
        const int y = (int)(i / w), x = (int)(i - (int64_t)y * w);
        const uint32_t px = hy_fill_pixel(img, orig, w, h, x, y, self, ws1, ws2);
        if (px != self) img[i] = px;
    }
}

// The same fill for the side-by-side / top-bottom modes, written straight into the composed float32 tensor and the mask
// (SIG:1543-1552, GS:355-378) instead of back into the eye image: four pixels per thread, 128-bit stores, no k_compose pass.
// Holes only READ pixels whose mask is set, which this kernel never writes, so nothing is updated in place.  w % 4 == 0.
__global__ void __launch_bounds__(256) k_hybrid_gapfill_fused(const WarpArgs a, double ws1, double ws2) {
    const int w = a.w, h = a.h, frame = blockIdx.y, eye = blockIdx.z;
    __shared__ float s_q255[256];
    s_q255[threadIdx.x] = kQ255[threadIdx.x];
    __syncthreads();
    const int64_t base_off = (int64_t)frame * h * w;
    const uint32_t* orig = a.image_u8 + base_off;
    const uint32_t* img = a.out[eye] + base_off;
    const uint64_t pol = policy_evict_first();
    const int wq = w >> 2;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < (int64_t)h * wq; q += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(q / wq), x = (int)(q - (int64_t)y * wq) << 2;
        const uint4 v = *reinterpret_cast<const uint4*>(img + (int64_t)y * w + x);
        uint32_t px[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (!(px[j] >> 24)) px[j] = hy_fill_pixel(img, orig, w, h, x + j, y, px[j], ws1, ws2);
        float f[12], m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t r = px[j] & 255u, g = (px[j] >> 8) & 255u, b = (px[j] >> 16) & 255u;
            f[3 * j] = s_q255[r]; f[3 * j + 1] = s_q255[g]; f[3 * j + 2] = s_q255[b];
            m[j] = (r + g + b == 0u) ? 1.0f : 0.0f;
        }
        const int64_t o = fused_index(a, eye, frame, y, x);
        float4* dst = reinterpret_cast<float4*>(a.fused_stereo + o * 3);
        st_stream_f4(dst, make_float4(f[0], f[1], f[2], f[3]), pol);
        st_stream_f4(dst + 1, make_float4(f[4], f[5], f[6], f[7]), pol);
        st_stream_f4(dst + 2, make_float4(f[8], f[9], f[10], f[11]), pol);
        st_stream_f4(reinterpret_cast<float4*>(a.fused_mask + o), make_float4(m[0], m[1], m[2], m[3]), pol);
    }
}

cudaError_t launch_hybrid(const WarpArgs& a, cudaStream_t s) {
    size_t smem = (size_t)a.w * 12 + (size_t)((a.w + 31) / 32) * 8 + 16;
    const bool wide = smem > 56 * 1024;
    if (smem > 48 * 1024) {
        if (wide) cudaFuncSetAttribute(k_hybrid_splat<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        else cudaFuncSetAttribute(k_hybrid_splat<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    prof_begin(K_HYBRID_SPLAT, s);
    if (wide) k_hybrid_splat<512><<<dim3(a.h, a.n, 2), 512, smem, s>>>(a);
    else k_hybrid_splat<256><<<dim3(a.h, a.n, 2), 256, smem, s>>>(a);
    prof_end(K_HYBRID_SPLAT, s);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int64_t npx = (int64_t)a.h * a.w;
    int bx = (int)((npx + 255) / 256);
    if (bx > sm_count() * 8) bx = sm_count() * 8;
    prof_begin(K_HYBRID_GAPFILL, s);
    if (a.fused_stereo) {
        if (a.w % 4) return cudaErrorInvalidValue;      // (the caller only fuses rows of whole pixel quads)
        bx = (int)((npx / 4 + 255) / 256);
        if (bx > sm_count() * 16) bx = sm_count() * 16;
        k_hybrid_gapfill_fused<<<dim3(bx, a.n, 2), 256, 0, s>>>(a, exp(-0.5), exp(-1.0));
    } else {
        k_hybrid_gapfill<<<dim3(bx, a.n, 2), 256, 0, s>>>(a, exp(-0.5), exp(-1.0));
    }
    prof_end(K_HYBRID_GAPFILL, s);
    count_launch();
    return cudaGetLastError();
}

// hybrid_edge_plus (SIG:1778-1802): pixels the hybrid result leaves black take the polylines_soft pixel.
__global__ void __launch_bounds__(256) k_merge_black(uint32_t* __restrict__ primary, const uint32_t* __restrict__ fallback,
                                                     int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t p = primary[i];
        if ((p & 0x00FFFFFFu) == 0u) primary[i] = fallback[i] & 0x00FFFFFFu;
    }
}

static inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
size_t hybrid_plus_scratch_bytes(int n, int h, int w) {
    return 2 * up256((size_t)n * h * w * 4) + up256(polylines_scratch_bytes(n, h, w));
}

cudaError_t launch_hybrid_plus(const WarpArgs& a, cudaStream_t s) {
    if (a.scratch_bytes < hybrid_plus_scratch_bytes(a.n, a.h, a.w)) return cudaErrorInvalidValue;
    cudaError_t e = launch_hybrid(a, s);
    if (e != cudaSuccess) return e;
    const size_t eye_bytes = up256((size_t)a.n * a.h * a.w * 4);
    WarpArgs p = a;
    p.fill = CS_FILL_POLYLINES_SOFT;
    p.out[0] = reinterpret_cast<uint32_t*>(a.scratch);
    p.out[1] = reinterpret_cast<uint32_t*>((char*)a.scratch + eye_bytes);
    p.scratch = (char*)a.scratch + 2 * eye_bytes;
    p.scratch_bytes = polylines_scratch_bytes(a.n, a.h, a.w);
    if ((e = cudaMemsetAsync((char*)p.scratch + p.scratch_bytes - 16 * sizeof(int), 0, 16 * sizeof(int), s)) != cudaSuccess) return e;
    if ((e = launch_polylines(p, s)) != cudaSuccess) return e;
    const int64_t total = (int64_t)a.n * a.h * a.w;
    int bx = (int)((total + 255) / 256);
    if (bx > sm_count() * 16) bx = sm_count() * 16;
    for (int eye = 0; eye < 2; ++eye) {
        if (a.eye[eye].passthrough || !a.out[eye]) continue;
        prof_begin(K_MISC, s);
        k_merge_black<<<bx, 256, 0, s>>>(a.out[eye], p.out[eye], total);
        prof_end(K_MISC, s);
        count_launch();
    }
    return cudaGetLastError();
}

}  // namespace cs

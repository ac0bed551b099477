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
            double wg = exp(-(diff * diff) / 2.0);
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
        const double g0 = guidance(orig[i]);
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
        if (wt > 0.0) {
            float wtf = (float)wt;
            float v0 = fminf(fmaxf(nv0 / wtf, 0.0f), 255.0f);
            float v1 = fminf(fmaxf(nv1 / wtf, 0.0f), 255.0f);
            float v2 = fminf(fmaxf(nv2 / wtf, 0.0f), 255.0f);
            img[i] = pack_rgbx((int)v0, (int)v1, (int)v2);  // X stays 0: still "mask == 0" for the neighbours
        }
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
    k_hybrid_gapfill<<<dim3(bx, a.n, 2), 256, 0, s>>>(a, exp(-0.5), exp(-1.0));
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

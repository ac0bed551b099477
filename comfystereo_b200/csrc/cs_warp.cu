// cs_warp.cu -- W1 + F0/F1/F2 (apply_stereo_divergence_naive, SIG:1850-1910) and
// I (apply_stereo_divergence_inverse, SIG:1715-1737).
//
// One CTA per (image row, frame, eye).  Disparity is purely horizontal, so a row is a closed
// problem that lives in shared memory:
//   sweep     every source pixel computes its destination column in FP64 and claims it with a
//             shared-memory atomic.  The reference's sequential sweep ("last writer wins",
//             right->left for div_px >= 0, left->right otherwise) is exactly "smallest source
//             column wins" / "largest source column wins" (SURVEY.md 8a, W1); its z-buffer
//             variant (inverse) is "largest nd wins, ties to the smallest source column",
//             a 64-bit packed key.
//   resolve   every destination pixel gathers its winner's colour from the RGBX8 row.
//   fill      'naive': nearest filled pixel right-then-left within |int(div_px)|+1, found with
//             clz/ffs over a bit map of the filled flags (reads the pre-fill row, SIG:1893-1908);
//             'naive_interpolating': the reference's left-to-right gap walk has a true chain
//             dependency through uint8 wrap-around (Q5), so one thread replays it on the row in
//             shared memory (rows still run in parallel across CTAs).
//
// Bytes per pixel and eye: depth 4 B + RGBX8 4 B read (L2-resident scratch), RGBX8 4 B written.
#include "cs_internal.cuh"

namespace cs {

__device__ __forceinline__ float eye_depth(const WarpArgs& a, int eye, int frame, int64_t off, float scale) {
    float d = a.depth[eye][(int64_t)frame * a.h * a.w + off];
    return scale == 1.0f ? d : d * scale;
}

// Per-(frame, eye) normaliser from the statistics block (D1, SIG:1586-1600).
__device__ __forceinline__ Normalizer eye_normalizer(const WarpArgs& a, int eye, int frame, float* scale_out) {
    const FrameStats st = a.stats[frame];
    float scale = 1.0f;
    int lo, hi;
    if (a.use_blur_stats) {
        lo = eye ? st.r_min : st.l_min;
        hi = eye ? st.r_max : st.l_max;
    } else {
        lo = st.gray_min; hi = st.gray_max;
        if (a.scale_by_stats && ord2f(st.gray_max) <= 1.0f) scale = 255.0f;   // SIG:1475-1476
    }
    *scale_out = scale;
    return make_normalizer(lo, hi, scale, a.conv);
}

// distance (> 0) to the nearest set bit strictly right / left of x, or a large number
__device__ __forceinline__ int bit_dist_right(const uint32_t* bits, int nwords, int x, int limit) {
    int p = x + 1;
    int wi = p >> 5;
    if (wi >= nwords) return 1 << 30;
    uint32_t m = bits[wi] & (0xffffffffu << (p & 31));
    while (true) {
        if (m) return (wi << 5) + (__ffs(m) - 1) - x;
        ++wi;
        if (wi >= nwords || (wi << 5) - x > limit) return 1 << 30;
        m = bits[wi];
    }
}
__device__ __forceinline__ int bit_dist_left(const uint32_t* bits, int x, int limit) {
    int p = x - 1;
    if (p < 0) return 1 << 30;
    int wi = p >> 5;
    uint32_t m = bits[wi] & (0xffffffffu >> (31 - (p & 31)));
    while (true) {
        if (m) return x - ((wi << 5) + 31 - __clz(m));
        --wi;
        if (wi < 0 || x - ((wi << 5) + 31) > limit) return 1 << 30;
        m = bits[wi];
    }
}

__device__ __forceinline__ int px_sum(uint32_t p) { return (int)(p & 255) + (int)((p >> 8) & 255) + (int)((p >> 16) & 255); }

// np.interp row fill of the *_post variants (SIG:1804-1833): left of the first valid column its value, right of the
// last one its value, between two valid columns  slope = (y1 - y0) / (x1 - x0);  slope * (x - x0) + y0  in float64,
// stored as float32, truncated to uint8.  `xl`, `xr` are the nearest valid columns (-1 / w when there is none).
__device__ __forceinline__ uint32_t interp_px(int x, int xl, int xr, int w, uint32_t pl, uint32_t pr) {
    if (xl < 0) return pr & 0x00FFFFFFu;
    if (xr >= w) return pl & 0x00FFFFFFu;
    uint32_t o = 0;
    const double dx = (double)xr - (double)xl, t = (double)x - (double)xl;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const double y0 = (double)((pl >> (8 * ch)) & 255u), y1 = (double)((pr >> (8 * ch)) & 255u);
        const double slope = (y1 - y0) / dx;
        const double v = slope * t;
        const float rf = (float)(v + y0);
        o |= ((uint32_t)(int)rf & 255u) << (8 * ch);
    }
    return o;
}

// Register budget per variant (measured): the sweep-only variants run fastest compiled for up to 512 threads (40
// registers), the ones with a second pass over the row (interpolating fill, z-buffer) with the 256-thread budget.
template <int FILL>
constexpr int rows_max_threads() {
    return (FILL == CS_FILL_NAIVE_INTERP || FILL == CS_FILL_INVERSE || FILL == CS_FILL_INVERSE_POST) ? 256 : 512;
}

template <int FILL>  // CS_FILL_NONE / NAIVE / NAIVE_INTERP / INVERSE / NONE_POST / INVERSE_POST
__global__ void __launch_bounds__(rows_max_threads<FILL>()) k_warp_rows(const WarpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int w = a.w, y = blockIdx.x, frame = blockIdx.y, eye = blockIdx.z;
    if (a.eye[eye].passthrough) return;
    const int nwords = (w + 31) >> 5;
    const double div_px = a.eye[eye].div_px, sep_px = a.eye[eye].sep_px;
    float scale;
    const Normalizer norm = eye_normalizer(a, eye, frame, &scale);
    const int64_t row_off = (int64_t)y * w;
    uint32_t* out = a.fused_stereo ? nullptr : a.out[eye] + (int64_t)frame * a.h * w + row_off;
    // one finished pixel: the RGBX8 eye image, or (fused: side-by-side / top-bottom modes) its place in the composed
    // float32 tensor and the black-pixel mask (C1 + M1 + O1, SIG:1543-1552, GS:355-378) -- no k_compose pass then
    __shared__ float s_q255[256];
    if (a.fused_stereo) for (int k = threadIdx.x; k < 256; k += blockDim.x) s_q255[k] = kQ255[k];   // visible after the first barrier
    const uint64_t pol = policy_evict_first();
    // (pixel by pixel: three 4-byte stores at a 12-byte stride per warp merge in L2; a warp-collective version that
    // shuffles the pixels into 128-bit stores measured 10 % slower)
    auto emit_w = [&](int x, uint32_t px, bool in) {
        if (!in) return;
        if (out) { out[x] = px; return; }
        const int64_t o = fused_index(a, eye, frame, y, x);
        const uint32_t r = px & 255u, g = (px >> 8) & 255u, b = (px >> 16) & 255u;
        float* dst = a.fused_stereo + o * 3;
        st_stream_f1(dst, s_q255[r], pol);
        st_stream_f1(dst + 1, s_q255[g], pol);
        st_stream_f1(dst + 2, s_q255[b], pol);
        st_stream_f1(a.fused_mask + o, (r + g + b == 0u) ? 1.0f : 0.0f, pol);
    };

    // the RGBX8 row is gathered at shifted columns: stage it in shared memory with one coalesced pass (the scattered
    // global reads were 31 % of this kernel's stall samples), and fetch depth four columns ahead of the FP64 chain
    uint32_t* simg = reinterpret_cast<uint32_t*>(smem_raw);
    const size_t simg_bytes = ((size_t)w * 4 + 15) & ~(size_t)15;
    unsigned char* smem_rest = smem_raw + simg_bytes;
    // one bulk asynchronous copy (TMA) when the row is 16-byte aligned: it lands while the winners are being resolved
    // and is waited for right before the first gather; otherwise a coalesced loop
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t* gimg = a.image_u8 + (int64_t)frame * a.h * w + row_off;
    const bool bulk = ((reinterpret_cast<uintptr_t>(gimg) & 15) == 0) && ((w & 3) == 0);
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            mbar_expect_tx(&s_bar, (uint32_t)w * 4u);
            bulk_g2s(simg, gimg, (uint32_t)w * 4u, &s_bar);
        }
    } else {
        for (int x = threadIdx.x; x < w; x += blockDim.x) simg[x] = gimg[x];
    }
    const uint32_t* img = simg;
    const float* dep = a.depth[eye] + (int64_t)frame * a.h * w + row_off;

    if (FILL == CS_FILL_INVERSE || FILL == CS_FILL_INVERSE_POST) {
        unsigned long long* key = reinterpret_cast<unsigned long long*>(smem_rest);
        for (int x = threadIdx.x; x < w; x += blockDim.x) key[x] = 0ull;
        __syncthreads();
        for (int xb = threadIdx.x; xb < w; xb += 4 * blockDim.x) {
          float dv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { const int x = xb + u * blockDim.x; dv[u] = (x < w) ? dep[x] : 0.0f; }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int x = xb + u * blockDim.x;
            if (x >= w) break;
            float nd = norm(scale == 1.0f ? dv[u] : dv[u] * scale) + 0.0f;
            if (!(nd > -1.0f)) continue;  // z-buffer starts at -1 and the test is strict (SIG:1722, 1731)
            double off = signed_pow_offset(nd, a.expo, div_px);
            double dx = ((double)x + 0.5) + off;
            dx = dx + sep_px;
            int j = (int)floor(dx);
            // larger nd wins, ties go to the smaller source column (ascending sweep + strict '>')
            unsigned long long k = ((unsigned long long)((uint32_t)f2ord(nd) ^ 0x80000000u) << 32) |
                                   (unsigned long long)(0xFFFFFFFFu - (uint32_t)x);
            if (j >= 0 && j < w) atomicMax(&key[j], k);
            if (j + 1 >= 0 && j + 1 < w) atomicMax(&key[j + 1], k);
          }
        }
        __syncthreads();
        if (bulk) mbar_wait(&s_bar, 0);   // the image row has landed
        if (FILL == CS_FILL_INVERSE_POST) {
            uint32_t* bits = reinterpret_cast<uint32_t*>(key + w);
            const int wpad = nwords << 5;
            for (int x = threadIdx.x; x < wpad; x += blockDim.x) {
                bool f = (x < w) && (key[x] != 0ull);
                uint32_t b = __ballot_sync(0xffffffffu, f);
                if ((threadIdx.x & 31) == 0) bits[x >> 5] = b;
            }
            __syncthreads();
            for (int x = threadIdx.x; x - (int)(threadIdx.x & 31) < w; x += blockDim.x) {
                uint32_t px_ = 0;
                const bool in_ = x < w;
                if (in_) {
                unsigned long long k = key[x];
                uint32_t px;
                if (k) px = (img[0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull)] & 0x00FFFFFFu) | 0x01000000u;
                else {
                    const int dr = bit_dist_right(bits, nwords, x, 1 << 29), dl = bit_dist_left(bits, x, 1 << 29);
                    const int xl = (dl < (1 << 29)) ? x - dl : -1, xr = (dr < (1 << 29)) ? x + dr : w;
                    if (xl < 0 && xr >= w) px = 0u;   // no valid column in the row: it stays as mapped (black)
                    else {
                        const uint32_t pl = xl >= 0 ? img[0xFFFFFFFFu - (uint32_t)(key[xl] & 0xFFFFFFFFull)] : 0u;
                        const uint32_t pr = xr < w ? img[0xFFFFFFFFu - (uint32_t)(key[xr] & 0xFFFFFFFFull)] : 0u;
                        px = interp_px(x, xl, xr, w, pl, pr);
                    }
                }
                px_ = px;
                }
                emit_w(x, px_, in_);
            }
            return;
        }
        for (int x = threadIdx.x; x - (int)(threadIdx.x & 31) < w; x += blockDim.x) {
            uint32_t px_ = 0;
            const bool in_ = x < w;
            if (in_) {
            unsigned long long k = key[x];
            uint32_t px = 0;
            if (k) px = (img[0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull)] & 0x00FFFFFFu) | 0x01000000u;
            px_ = px;
            }
            emit_w(x, px_, in_);
        }
        return;
    } else {
        int* win = reinterpret_cast<int*>(smem_rest);
        uint32_t* bits = reinterpret_cast<uint32_t*>(win + w);
        uint32_t* row = bits + nwords;  // only FILL == NAIVE_INTERP
        const bool take_min = !(div_px < 0);
        const int empty = take_min ? 0x7FFFFFFF : -1;
        for (int x = threadIdx.x; x < w; x += blockDim.x) win[x] = empty;
        __syncthreads();
        for (int xb = threadIdx.x; xb < w; xb += 4 * blockDim.x) {
          float dv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { const int x = xb + u * blockDim.x; dv[u] = (x < w) ? dep[x] : 0.0f; }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int x = xb + u * blockDim.x;
            if (x >= w) break;
            float nd = norm(scale == 1.0f ? dv[u] : dv[u] * scale);
            double off = signed_pow_offset(nd, a.expo, div_px);
            double t = off + sep_px;
            int cd = x + (int)t;  // truncation toward zero, SIG:1865
            if (cd >= 0 && cd < w) {
                if (take_min) atomicMin(&win[cd], x); else atomicMax(&win[cd], x);
            }
          }
        }
        __syncthreads();
        const int wpad = nwords << 5;
        for (int x = threadIdx.x; x < wpad; x += blockDim.x) {
            bool f = (x < w) && (win[x] != empty);
            uint32_t b = __ballot_sync(0xffffffffu, f);
            if ((threadIdx.x & 31) == 0) bits[x >> 5] = b;
        }
        __syncthreads();
        if (bulk) mbar_wait(&s_bar, 0);   // the image row has landed
        if (FILL == CS_FILL_NONE) {
            for (int x = threadIdx.x; x - (int)(threadIdx.x & 31) < w; x += blockDim.x) {
                uint32_t px_ = 0;
                const bool in_ = x < w;
                if (in_) {
                int s = win[x];
                px_ = (s != empty) ? ((img[s] & 0x00FFFFFFu) | 0x01000000u) : 0u;
                }
                emit_w(x, px_, in_);
            }
        } else if (FILL == CS_FILL_NONE_POST) {
            for (int x = threadIdx.x; x - (int)(threadIdx.x & 31) < w; x += blockDim.x) {
                uint32_t px_ = 0;
                const bool in_ = x < w;
                if (in_) {
                int s = win[x];
                uint32_t px;
                if (s != empty) px = (img[s] & 0x00FFFFFFu) | 0x01000000u;
                else {
                    const int dr = bit_dist_right(bits, nwords, x, 1 << 29), dl = bit_dist_left(bits, x, 1 << 29);
                    const int xl = (dl < (1 << 29)) ? x - dl : -1, xr = (dr < (1 << 29)) ? x + dr : w;
                    if (xl < 0 && xr >= w) px = 0u;
                    else px = interp_px(x, xl, xr, w, xl >= 0 ? img[win[xl]] : 0u, xr < w ? img[win[xr]] : 0u);
                }
                px_ = px;
                }
                emit_w(x, px_, in_);
            }
        } else if (FILL == CS_FILL_NAIVE) {
            double adiv = fabs(div_px);
            const int reach = (int)adiv + 1;  // range(1, abs(int(div_px)) + 2), SIG:1896
            for (int x = threadIdx.x; x - (int)(threadIdx.x & 31) < w; x += blockDim.x) {
                uint32_t px_ = 0;
                const bool in_ = x < w;
                if (in_) {
                int s = win[x];
                uint32_t px = 0;
                if (s != empty) px = (img[s] & 0x00FFFFFFu) | 0x01000000u;
                else {
                    int dr = bit_dist_right(bits, nwords, x, reach), dl = bit_dist_left(bits, x, reach);
                    if (dr <= reach && dr <= dl) px = img[win[x + dr]] & 0x00FFFFFFu;
                    else if (dl <= reach) px = img[win[x - dl]] & 0x00FFFFFFu;
                }
                px_ = px;
                }
                emit_w(x, px_, in_);
            }
        } else {  // naive_interpolating, SIG:1871-1892
            // An "anchor" is a pixel that stops the reference's right-border scan: non-black and filled.  A fill started
            // at l writes [l, r) where r is the next anchor, and reads row[l - 1] >= the previous anchor, so everything the
            // sequential loop does between two neighbouring anchors stays between them and anchors never change: the
            // intervals are independent.  One thread replays the reference's loop on each interval that holds a gap.
            uint32_t* abits = row + w;
            for (int x = threadIdx.x; x < wpad; x += blockDim.x) {
                bool anc = false;
                if (x < w) {
                    const int s = win[x];
                    const uint32_t v = (s != empty) ? (img[s] & 0x00FFFFFFu) : 0u;
                    row[x] = v;
                    anc = (s != empty) && px_sum(v) != 0;
                }
                const uint32_t b = __ballot_sync(0xffffffffu, anc);
                if ((threadIdx.x & 31) == 0) abits[x >> 5] = b;
            }
            __syncthreads();
            for (int x = (int)threadIdx.x - 1; x < w; x += blockDim.x) {    // x = -1: the interval left of the first anchor
                if (x >= 0 && !((abits[x >> 5] >> (x & 31)) & 1u)) continue;
                const int dn = bit_dist_right(abits, nwords, x, 1 << 29);
                const int nx = (dn < (1 << 29)) ? x + dn : w;               // next anchor, or the end of the row
                // only unfilled pixels can start a gap: walk the zero bits of the filled map inside (x, nx)
                for (int wi = (x + 1) >> 5; wi <= ((nx - 1) >> 5) && x + 1 < nx; ++wi) {
                  uint32_t um = ~bits[wi];
                  if ((wi << 5) < x + 1) um &= 0xffffffffu << ((x + 1) & 31);
                  if (((wi + 1) << 5) > nx) um &= (nx & 31) ? ((1u << (nx & 31)) - 1u) : 0xffffffffu;
                  while (um) {
                    const int l = (wi << 5) + __ffs(um) - 1;
                    um &= um - 1;
                    if (px_sum(row[l]) != 0) continue;
                    uint32_t lb = (l > 0) ? row[l - 1] : 0u, rb = 0u;
                    int r = l + 1;
                    while (r < w) {
                        if (px_sum(row[r]) != 0 && ((bits[r >> 5] >> (r & 31)) & 1u)) { rb = row[r]; break; }
                        ++r;
                    }
                    if (px_sum(lb) == 0) lb = rb;
                    else if (px_sum(rb) == 0) rb = lb;
                    const int total = 1 + r - l;
                    float stepv[3];   // float32 array / int64 scalar stays float32 in numba
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        float df = (float)((rb >> (8 * ch)) & 255u) - (float)((lb >> (8 * ch)) & 255u);
                        stepv[ch] = df / (float)total;
                    }
                    for (int c = l; c < r; ++c) {
                        uint32_t px = 0;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            float prod = stepv[ch] * (float)(c - l + 1);
                            uint32_t inc = (uint32_t)__float2int_rz(prod) & 255u;   // float -> uint8 wraps (Q5)
                            uint32_t v = (((lb >> (8 * ch)) & 255u) + inc) & 255u;
                            px |= v << (8 * ch);
                        }
                        row[c] = px;
                    }
                  }
                }
            }
            __syncthreads();
            for (int x = threadIdx.x; x - (int)(threadIdx.x & 31) < w; x += blockDim.x) {
                uint32_t px_ = 0;
                const bool in_ = x < w;
                if (in_) {
                px_ = row[x] | (((bits[x >> 5] >> (x & 31)) & 1u) << 24);
                }
                emit_w(x, px_, in_);
            }
        }
    }
}

// a wide row's shared memory limits the CTAs per SM: keep the SM's warp slots busy with wider CTAs then
static int row_threads(size_t smem) { return smem > 56 * 1024 ? 512 : 256; }

template <int FILL>
static void launch_rows_as(const WarpArgs& a, dim3 grid, size_t smem, cudaStream_t s) {
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(k_warp_rows<FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int threads = row_threads(smem) < rows_max_threads<FILL>() ? row_threads(smem) : rows_max_threads<FILL>();
    k_warp_rows<FILL><<<grid, threads, smem, s>>>(a);
}

cudaError_t launch_warp_rows(const WarpArgs& a, cudaStream_t s) {
    const int nwords = (a.w + 31) >> 5;
    dim3 grid(a.h, a.n, 2);
    const size_t simg = ((size_t)a.w * 4 + 15) & ~(size_t)15;   // staged RGBX8 row
    const size_t base = simg + (size_t)a.w * 4 + nwords * 4;     // + winners + filled bitmap
    if (base + (size_t)a.w * 4 + nwords * 4 > 227 * 1024) return cudaErrorInvalidValue;
    prof_begin(K_WARP_ROWS, s);
    switch (a.fill) {
        case CS_FILL_NONE: launch_rows_as<CS_FILL_NONE>(a, grid, base, s); break;
        case CS_FILL_NAIVE: launch_rows_as<CS_FILL_NAIVE>(a, grid, base, s); break;
        case CS_FILL_NAIVE_INTERP: launch_rows_as<CS_FILL_NAIVE_INTERP>(a, grid, base + (size_t)a.w * 4 + nwords * 4, s); break;
        case CS_FILL_NONE_POST: launch_rows_as<CS_FILL_NONE_POST>(a, grid, base, s); break;
        case CS_FILL_INVERSE_POST: launch_rows_as<CS_FILL_INVERSE_POST>(a, grid, simg + (size_t)a.w * 8 + nwords * 4, s); break;
        case CS_FILL_INVERSE: launch_rows_as<CS_FILL_INVERSE>(a, grid, simg + (size_t)a.w * 8, s); break;
        default:
            prof_end(K_WARP_ROWS, s);
            return cudaErrorInvalidValue;
    }
    prof_end(K_WARP_ROWS, s);
    count_launch();
    return cudaGetLastError();
}

}  // namespace cs

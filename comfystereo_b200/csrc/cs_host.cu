// cs_host.cu -- cs_stereo_batch_host: the hot path with HOST buffers, the shape in which ComfyUI
// hands IMAGE tensors to the node and expects them back (GS:126, GS:161-171, GS:298-307).
//
// Frames stream through kSlots device slots: while chunk i runs its kernels on the compute stream,
// chunk i+1 is uploading on the copy-in stream and chunk i-1 is downloading on the copy-out stream.
//
// Page-locked caller buffers are copied directly (cudaMemcpyAsync is truly asynchronous on them).
// Pageable buffers -- what ComfyUI passes in, and what a large result tensor has to be -- are not
// handed to the driver (measured on B200: 2.2 GB/s into fresh pageable memory): the library owns
// page-locked bounce buffers (one per slot and direction) and moves data between them and the caller's memory
// with a team of host threads, overlapped with the GPU work of the neighbouring chunks.
// The depth outputs are three identical channels and the mask is 0/1: when the host has threads to spare, only
// one channel of each depth output (the byte k of k/255 for the CPU techniques, one float for GPU Warp) and one
// byte per mask pixel cross PCIe (a small kernel compacts them), and a consumer thread's team expands them into
// the caller's tensors with non-temporal stores while later chunks are in flight -- 58 MB instead of 116 MB per
// 1080p side-by-side frame.  Whatever needs host work after the download (bounce copy, expansion) is done by that
// consumer thread, in chunk order, so the producer only ever waits for a free slot.
// Device buffers, bounce buffers and streams are cached per device; cs_host_release() frees them.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "cs_internal.cuh"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__linux__)
#include <sys/mman.h>
#endif

namespace cs {

constexpr int kSlots = 4;   // chunks in flight: upload, kernels, download and the host-side drain each hold one

struct HostCtx {
    int device = -1;
    size_t in_bytes = 0, out_bytes = 0, ws_bytes = 0;
    size_t bounce_in_bytes = 0, bounce_out_bytes = 0, cmp_bytes = 0;
    char* d_cmp[kSlots] = {};    // compact depth / mask outputs (device) ...
    char* h_cmp[kSlots] = {};    // ... and their page-locked landing buffers
    char* d_in[kSlots] = {};
    char* d_out[kSlots] = {};
    char* h_in[kSlots] = {};     // page-locked bounce buffers (only when the caller's memory is pageable)
    char* h_out[kSlots] = {};
    char* d_ws = nullptr;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[kSlots] = {}, ev_run[kSlots] = {}, ev_out[kSlots] = {};
    bool ready = false;
};

constexpr int kMaxDevices = 64;          // device ordinals this process can address (CUDA's own limit per process)
static std::mutex g_mu[kMaxDevices];     // one pipeline per device; devices run concurrently
static HostCtx g_ctx[kMaxDevices];

static void ctx_free(HostCtx& c) {
    if (!c.ready) return;
    cudaSetDevice(c.device);
    for (int i = 0; i < kSlots; ++i) {
        if (c.d_in[i]) cudaFree(c.d_in[i]);
        if (c.d_out[i]) cudaFree(c.d_out[i]);
        if (c.h_in[i]) cudaFreeHost(c.h_in[i]);
        if (c.h_out[i]) cudaFreeHost(c.h_out[i]);
        if (c.d_cmp[i]) cudaFree(c.d_cmp[i]);
        if (c.h_cmp[i]) cudaFreeHost(c.h_cmp[i]);
        if (c.ev_in[i]) cudaEventDestroy(c.ev_in[i]);
        if (c.ev_run[i]) cudaEventDestroy(c.ev_run[i]);
        if (c.ev_out[i]) cudaEventDestroy(c.ev_out[i]);
    }
    if (c.d_ws) cudaFree(c.d_ws);
    if (c.s_in) cudaStreamDestroy(c.s_in);
    if (c.s_run) cudaStreamDestroy(c.s_run);
    if (c.s_out) cudaStreamDestroy(c.s_out);
    c = HostCtx();
}

static cudaError_t ctx_ensure(HostCtx& c, int device, size_t in_bytes, size_t out_bytes, size_t ws_bytes,
                              size_t cmp_bytes, bool bounce_in, bool bounce_out) {
    cudaError_t e;
    const bool fits = c.ready && c.in_bytes >= in_bytes && c.out_bytes >= out_bytes && c.ws_bytes >= ws_bytes &&
                      c.cmp_bytes >= cmp_bytes &&
                      (!bounce_in || c.bounce_in_bytes >= in_bytes) && (!bounce_out || c.bounce_out_bytes >= out_bytes);
    if (fits) return cudaSuccess;
    const bool keep_in = c.ready && c.bounce_in_bytes > 0, keep_out = c.ready && c.bounce_out_bytes > 0;
    if (c.ready && c.cmp_bytes > cmp_bytes) cmp_bytes = c.cmp_bytes;
    ctx_free(c);
    c.device = device;
    c.ready = true;
    for (int i = 0; i < kSlots; ++i) {
        if ((e = cudaMalloc((void**)&c.d_in[i], in_bytes)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void**)&c.d_out[i], out_bytes)) != cudaSuccess) return e;
        if (bounce_in || keep_in)
            if ((e = cudaHostAlloc((void**)&c.h_in[i], in_bytes, cudaHostAllocDefault)) != cudaSuccess) return e;
        if (bounce_out || keep_out)
            if ((e = cudaHostAlloc((void**)&c.h_out[i], out_bytes, cudaHostAllocDefault)) != cudaSuccess) return e;
        if (cmp_bytes) {
            if ((e = cudaMalloc((void**)&c.d_cmp[i], cmp_bytes)) != cudaSuccess) return e;
            if ((e = cudaHostAlloc((void**)&c.h_cmp[i], cmp_bytes, cudaHostAllocDefault)) != cudaSuccess) return e;
        }
        if ((e = cudaEventCreateWithFlags(&c.ev_in[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&c.ev_run[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&c.ev_out[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    if ((e = cudaMalloc((void**)&c.d_ws, ws_bytes)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&c.s_in, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&c.s_run, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&c.s_out, cudaStreamNonBlocking)) != cudaSuccess) return e;
    c.in_bytes = in_bytes; c.out_bytes = out_bytes; c.ws_bytes = ws_bytes; c.cmp_bytes = cmp_bytes;
    c.bounce_in_bytes = (bounce_in || keep_in) ? in_bytes : 0;
    c.bounce_out_bytes = (bounce_out || keep_out) ? out_bytes : 0;
    return cudaSuccess;
}

// Fresh pageable result tensors are first touched by this library: ask for huge pages so that the first touch costs one
// fault per 2 MB instead of one per 4 KB (a hint; ignored when transparent huge pages are off).
static void hint_huge_pages(void* p, size_t bytes) {
#if defined(__linux__) && defined(MADV_HUGEPAGE)
    const uintptr_t a = ((uintptr_t)p + 4095) & ~(uintptr_t)4095, b = ((uintptr_t)p + bytes) & ~(uintptr_t)4095;
    if (b > a + (4u << 20)) (void)madvise((void*)a, b - a, MADV_HUGEPAGE);
#else
    (void)p; (void)bytes;
#endif
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

struct Span { char* dst; const char* src; size_t bytes; };

// memcpy a list of spans with a team of threads (first touch of fresh pageable pages is the expensive part;
// it parallelises well)
static std::atomic<int> g_active{0};   // cs_stereo_batch_host calls in flight in this process

// Host threads this call may use: the machine's threads shared between the ranks of a one-process-per-GPU job
// (LOCAL_WORLD_SIZE, set by torchrun) or the pipelines this process is running side by side.
static int host_share() {
    unsigned hw = std::thread::hardware_concurrency();
    if (!hw) hw = 8;
    int peers = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) peers = std::max(1, atoi(e));
    peers = std::max(peers, g_active.load());
    return std::max(1, (int)hw / peers);
}

// compact transport of the depth outputs and the mask when the host has the threads to expand them
static bool compact_transport(int team) {
    if (const char* e = getenv("COMFYSTEREO_COMPACT_D2H")) return atoi(e) != 0;
    // Measured with four ranks on a 32-thread host (8 threads each): 783 fps compact vs 633 fps plain end to end -- the
    // smaller DMA writes matter more than the expansion's threads even when several pipelines share the host.  Only a
    // pipeline left with a single thread keeps the plain transport.
    return team >= 2;
}

static void parallel_copy(const std::vector<Span>& spans, int team) {
    size_t total = 0;
    for (const auto& s : spans) total += s.bytes;
    int nt = (int)std::min<size_t>((size_t)std::max(1, std::min(team, 16)), total / (4u << 20) + 1);
    if (nt <= 1) {
        for (const auto& s : spans) memcpy(s.dst, s.src, s.bytes);
        return;
    }
    const size_t per = (total + nt - 1) / nt;
    auto work = [&](int t) {
        size_t lo = (size_t)t * per, hi = std::min(total, lo + per), pos = 0;
        for (const auto& s : spans) {
            const size_t a = std::max(lo, pos), b = std::min(hi, pos + s.bytes);
            if (a < b) memcpy(s.dst + (a - pos), s.src + (a - pos), b - a);
            pos += s.bytes;
            if (pos >= hi) break;
        }
    };
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
}

// depth: one channel -> three identical ones; mask: bytes -> floats (the inverse of k_compact_outputs).
// The destination is written once and not read back here: non-temporal stores skip the read-for-ownership of every
// cache line, which is what bounds this loop while the DMA engines use the same memory.
static void expand3(const float* src, float* dst, size_t p0, size_t p1) {
    size_t i = p0;
#if defined(__SSE2__)
    for (; i < p1 && ((uintptr_t)(dst + 3 * i) & 15); ++i) { const float v = src[i]; dst[3 * i] = v; dst[3 * i + 1] = v; dst[3 * i + 2] = v; }
    for (; i + 4 <= p1; i += 4) {
        const __m128 v = _mm_loadu_ps(src + i);                       // a b c d
        float* d = dst + 3 * i;
        _mm_stream_ps(d, _mm_shuffle_ps(v, v, _MM_SHUFFLE(1, 0, 0, 0)));      // a a a b
        _mm_stream_ps(d + 4, _mm_shuffle_ps(v, v, _MM_SHUFFLE(2, 2, 1, 1)));  // b b c c
        _mm_stream_ps(d + 8, _mm_shuffle_ps(v, v, _MM_SHUFFLE(3, 3, 3, 2)));  // c d d d
    }
#endif
    for (; i < p1; ++i) { const float v = src[i]; dst[3 * i] = v; dst[3 * i + 1] = v; dst[3 * i + 2] = v; }
}

// byte k -> k / 255.0f three times (the CPU techniques' depth outputs); lut[k] is the same IEEE division the kernels do
static void expand3_u8(const uint8_t* src, float* dst, const float* lut, size_t p0, size_t p1) {
    size_t i = p0;
#if defined(__SSE2__)
    for (; i < p1 && ((uintptr_t)(dst + 3 * i) & 15); ++i) { const float v = lut[src[i]]; dst[3 * i] = v; dst[3 * i + 1] = v; dst[3 * i + 2] = v; }
    for (; i + 4 <= p1; i += 4) {
        const __m128 v = _mm_set_ps(lut[src[i + 3]], lut[src[i + 2]], lut[src[i + 1]], lut[src[i]]);   // a b c d
        float* d = dst + 3 * i;
        _mm_stream_ps(d, _mm_shuffle_ps(v, v, _MM_SHUFFLE(1, 0, 0, 0)));
        _mm_stream_ps(d + 4, _mm_shuffle_ps(v, v, _MM_SHUFFLE(2, 2, 1, 1)));
        _mm_stream_ps(d + 8, _mm_shuffle_ps(v, v, _MM_SHUFFLE(3, 3, 3, 2)));
    }
#endif
    for (; i < p1; ++i) { const float v = lut[src[i]]; dst[3 * i] = v; dst[3 * i + 1] = v; dst[3 * i + 2] = v; }
}

static void expand_mask(const uint8_t* src, float* dst, size_t m0, size_t m1) {
    size_t i = m0;
#if defined(__SSE2__)
    for (; i < m1 && ((uintptr_t)(dst + i) & 15); ++i) dst[i] = (float)src[i];
    const __m128i zero = _mm_setzero_si128();
    for (; i + 16 <= m1; i += 16) {
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
        const __m128i lo = _mm_unpacklo_epi8(b, zero), hi = _mm_unpackhi_epi8(b, zero);
        _mm_stream_ps(dst + i, _mm_cvtepi32_ps(_mm_unpacklo_epi16(lo, zero)));
        _mm_stream_ps(dst + i + 4, _mm_cvtepi32_ps(_mm_unpackhi_epi16(lo, zero)));
        _mm_stream_ps(dst + i + 8, _mm_cvtepi32_ps(_mm_unpacklo_epi16(hi, zero)));
        _mm_stream_ps(dst + i + 12, _mm_cvtepi32_ps(_mm_unpackhi_epi16(hi, zero)));
    }
#endif
    for (; i < m1; ++i) dst[i] = (float)src[i];
}

static void expand_slice(const void* cdl, const void* cdr, const uint8_t* cm, const float* lut, float* dl, float* dr,
                         float* mask, size_t p0, size_t p1, size_t m0, size_t m1) {
    if (lut) {
        expand3_u8((const uint8_t*)cdl, dl, lut, p0, p1);
        expand3_u8((const uint8_t*)cdr, dr, lut, p0, p1);
    } else {
        expand3((const float*)cdl, dl, p0, p1);
        expand3((const float*)cdr, dr, p0, p1);
    }
    expand_mask(cm, mask, m0, m1);
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

static void parallel_expand(const void* cdl, const void* cdr, const uint8_t* cm, const float* lut, float* dl, float* dr,
                            float* mask, size_t npx, size_t nmask, int team) {
    int nt = (int)std::min<size_t>((size_t)std::max(1, std::min(team, 16)), (npx * 24 + nmask * 4) / (4u << 20) + 1);
    if (nt <= 1) { expand_slice(cdl, cdr, cm, lut, dl, dr, mask, 0, npx, 0, nmask); return; }
    auto work = [&](int t) {
        expand_slice(cdl, cdr, cm, lut, dl, dr, mask, npx * t / nt, npx * (t + 1) / nt, nmask * t / nt,
                     nmask * (t + 1) / nt);
    };
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
}

}  // namespace cs

using namespace cs;

extern "C" {

int cs_host_compact_enabled(void) { return compact_transport(host_share()) ? 1 : 0; }

// Streaming-copy bandwidth of this host (GB/s, read + write bytes): `threads` threads (0 = what one pipeline may use)
// copy `bytes` from one buffer to another with the non-temporal stores of the expansion loops, best of three.
// bench.py relates the end-to-end rate to it (e2e.host_frac).
double cs_host_stream_bandwidth(size_t bytes, int threads) {
    if (bytes < (1u << 20)) bytes = 1u << 20;
    bytes &= ~(size_t)63;
    int nt = threads > 0 ? threads : host_share();
    if (nt > 64) nt = 64;
    float* a = (float*)aligned_alloc(64, bytes);
    float* b = (float*)aligned_alloc(64, bytes);
    if (!a || !b) { free(a); free(b); return 0.0; }
    const size_t nf = bytes / 4;
    auto work = [&](int t, int pass) {
        const size_t lo = (nf * t / nt) & ~(size_t)15, hi = (t == nt - 1) ? nf : ((nf * (t + 1) / nt) & ~(size_t)15);
        if (pass == 0) { for (size_t i = lo; i < hi; ++i) { a[i] = (float)i; b[i] = 0.0f; } return; }   // first touch
        size_t i = lo;
#if defined(__SSE2__)
        for (; i + 4 <= hi; i += 4) _mm_stream_ps(b + i, _mm_load_ps(a + i));
        _mm_sfence();
#endif
        for (; i < hi; ++i) b[i] = a[i];
    };
    double best = 0.0;
    for (int pass = 0; pass < 4; ++pass) {
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work, t, pass);
        work(0, pass);
        for (auto& t : th) t.join();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (pass > 0 && dt > 0.0) best = std::max(best, 2.0 * (double)bytes / dt / 1e9);
    }
    free(a); free(b);
    return best;
}

void cs_host_release(void) {
    for (int i = 0; i < kMaxDevices; ++i) {
        std::lock_guard<std::mutex> lk(g_mu[i]);
        ctx_free(g_ctx[i]);
    }
    cs::release_graphs();
}

int cs_stereo_batch_host(const cs_params* p, const float* image, const float* depth, int n, int h, int w, int c,
                         float* stereo, float* depth_l, float* depth_r, float* mask, int device) {
    return cs_stereo_batch_host_progress(p, image, depth, n, h, w, c, stereo, depth_l, depth_r, mask, device, nullptr, nullptr);
}

int cs_stereo_batch_host_progress(const cs_params* p, const float* image, const float* depth, int n, int h, int w, int c,
                                  float* stereo, float* depth_l, float* depth_r, float* mask, int device,
                                  cs_progress_fn progress, void* user) {
#define HOST_FAIL(code, ...) return cs::fail(code, __VA_ARGS__)
#define HOST_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cs::fail(CS_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)
    if (!p || !image || !depth || !stereo || !depth_l || !depth_r || !mask) HOST_FAIL(CS_ERR_ARG, "cs_stereo_batch_host: NULL pointer");
    if (n < 1 || h < 1 || w < 2 || c < 1) HOST_FAIL(CS_ERR_ARG, "cs_stereo_batch_host: bad size");
    if (device < 0 || device >= cs::kMaxDevices) HOST_FAIL(CS_ERR_ARG, "cs_stereo_batch_host: bad device %d", device);
    int ho, wo, hm, wm;
    int rc = cs_output_dims(p, h, w, &ho, &wo, &hm, &wm);
    if (rc) return rc;
    int prev_device = -1;
    HOST_CUDA(cudaGetDevice(&prev_device));
    HOST_CUDA(cudaSetDevice(device));
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_device};   // leave the caller's device as found

    const size_t px = (size_t)h * w;
    const bool resize = p->depth_h > 0 && p->depth_w > 0 && (p->depth_h != h || p->depth_w != w);
    const size_t dpx = resize ? (size_t)p->depth_h * p->depth_w : px;   // depth frames may have their own size (N1)
    const size_t b_img = px * 3 * 4, b_dep = dpx * c * 4, b_st = (size_t)ho * wo * 3 * 4, b_d = px * 3 * 4, b_m = (size_t)hm * wm * 4;
    const size_t in_frame = b_img + b_dep, out_frame = b_st + 2 * b_d + b_m;
    const bool gpu_warp = p->fill == CS_FILL_GPU_WARP || p->fill == CS_FILL_GPU_WARP_MESH;
    // chunk: whole GPU-Warp sub-batches (Q9); otherwise ~256 MB of I/O per slot
    int group = (gpu_warp && p->group_size > 0) ? (p->group_size < n ? p->group_size : n) : 1;
    int chunk = (int)((size_t)256 * 1024 * 1024 / (in_frame + out_frame));
    if (chunk < 1) chunk = 1;
    chunk = (chunk / group) * group;
    if (chunk < group) chunk = group;
    if (chunk > n) chunk = n;
    const size_t ws_bytes = cs_workspace_bytes(p, chunk, h, w);
    struct Active { Active() { g_active.fetch_add(1); } ~Active() { g_active.fetch_sub(1); } } active;
    const int team = host_share();
    const bool compact = compact_transport(team);
    const bool bounce_in = !(is_pinned(image) && is_pinned(depth));
    const bool stereo_pinned = is_pinned(stereo);
    const bool bounce_out = compact ? !stereo_pinned
                                    : !(stereo_pinned && is_pinned(depth_l) && is_pinned(depth_r) && is_pinned(mask));
    if (!stereo_pinned) hint_huge_pages(stereo, (size_t)n * b_st);
    if (!is_pinned(depth_l)) hint_huge_pages(depth_l, (size_t)n * b_d);
    if (!is_pinned(depth_r)) hint_huge_pages(depth_r, (size_t)n * b_d);
    if (!is_pinned(mask)) hint_huge_pages(mask, (size_t)n * b_m);
    auto al256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // every CPU technique's depth outputs are exactly k / 255 (wrap quirk Q1): one byte per pixel is enough
    const bool depth_u8 = compact && !gpu_warp;
    const size_t dsz = depth_u8 ? 1 : 4;
    float lut[256];
    for (int k = 0; k < 256; ++k) lut[k] = (float)k / 255.0f;
    const size_t cmp_depth = al256((size_t)chunk * px * dsz), cmp_mask = al256((size_t)chunk * hm * wm);
    const size_t cmp_bytes = compact ? 2 * cmp_depth + cmp_mask : 0;

    std::lock_guard<std::mutex> lk(g_mu[device]);
    HostCtx& cx = g_ctx[device];
    cudaError_t e = ctx_ensure(cx, device, (size_t)chunk * in_frame, (size_t)chunk * out_frame, ws_bytes, cmp_bytes,
                               bounce_in, bounce_out);
    if (e != cudaSuccess) { ctx_free(cx); HOST_FAIL(CS_ERR_CUDA, "device buffers: %s", cudaGetErrorString(e)); }

    // spans of chunk `it` between the caller's tensors and a contiguous slot
    auto in_spans = [&](int f0, int m, char* slot, bool to_slot) {
        std::vector<Span> v;
        char* a = (char*)(image + (size_t)f0 * px * 3);
        char* b = (char*)(depth + (size_t)f0 * dpx * c);
        if (to_slot) { v.push_back({slot, a, (size_t)m * b_img}); v.push_back({slot + (size_t)m * b_img, b, (size_t)m * b_dep}); }
        return v;
    };
    auto out_spans = [&](int f0, int m, const char* slot) {
        std::vector<Span> v;
        v.push_back({(char*)(stereo + (size_t)f0 * ho * wo * 3), slot, (size_t)m * b_st});
        if (compact) return v;    // depth outputs and mask come through the compact buffers
        v.push_back({(char*)(depth_l + (size_t)f0 * px * 3), slot + (size_t)m * b_st, (size_t)m * b_d});
        v.push_back({(char*)(depth_r + (size_t)f0 * px * 3), slot + (size_t)m * (b_st + b_d), (size_t)m * b_d});
        v.push_back({(char*)(mask + (size_t)f0 * hm * wm), slot + (size_t)m * (b_st + 2 * b_d), (size_t)m * b_m});
        return v;
    };

    const int nchunks = (n + chunk - 1) / chunk;
    // Producer (this thread) enqueues chunk after chunk; when results need host work (bounce copy and / or expansion)
    // a consumer thread drains them in order, so that the uploads and kernels of later chunks are never held up by it.
    const bool drain = bounce_out || compact;
    std::mutex mu;
    std::condition_variable cv;
    int enqueued = 0, drained = 0;      // chunks whose download has been enqueued / whose results reached the caller
    bool stop = false;
    int drain_rc = CS_OK;
    double t_event = 0.0, t_host = 0.0, t_slot = 0.0;   // COMFYSTEREO_HOST_TRACE=1: where the consumer / producer waited
    const auto call0 = std::chrono::steady_clock::now();
    auto consumer = [&]() {
        cudaSetDevice(device);
        for (int pit = 0; pit < nchunks; ++pit) {
            {
                std::unique_lock<std::mutex> g(mu);
                cv.wait(g, [&] { return enqueued > pit || stop; });
                if (enqueued <= pit) return;
            }
            const int f0 = pit * chunk, m = (n - f0 < chunk) ? n - f0 : chunk, sl = pit % kSlots;
            const auto c0 = std::chrono::steady_clock::now();
            if (cudaEventSynchronize(cx.ev_out[sl]) != cudaSuccess) { std::lock_guard<std::mutex> g(mu); drain_rc = CS_ERR_CUDA; }
            const auto c1 = std::chrono::steady_clock::now();
            if (bounce_out) parallel_copy(out_spans(f0, m, cx.h_out[sl]), team);
            if (compact)
                parallel_expand(cx.h_cmp[sl], cx.h_cmp[sl] + cmp_depth, (const uint8_t*)(cx.h_cmp[sl] + 2 * cmp_depth),
                                depth_u8 ? lut : nullptr, depth_l + (size_t)f0 * px * 3,
                                depth_r + (size_t)f0 * px * 3, mask + (size_t)f0 * hm * wm, (size_t)m * px,
                                (size_t)m * hm * wm, team);
            const auto c2 = std::chrono::steady_clock::now();
            t_event += std::chrono::duration<double>(c1 - c0).count();
            t_host += std::chrono::duration<double>(c2 - c1).count();
            { std::lock_guard<std::mutex> g(mu); drained = pit + 1; }
            cv.notify_all();
        }
    };
    std::thread drainer;
    if (drain) drainer = std::thread(consumer);
    // every exit path stops and joins the consumer
    struct Joiner {
        std::thread& t; std::mutex& mu; std::condition_variable& cv; bool& stop;
        ~Joiner() { if (t.joinable()) { { std::lock_guard<std::mutex> g(mu); stop = true; } cv.notify_all(); t.join(); } }
    } joiner{drainer, mu, cv, stop};
    // an early error return must not leave copies into the caller's buffers in flight (they may be freed right after)
    struct SyncOnError { bool ok = false; ~SyncOnError() { if (!ok) { cudaDeviceSynchronize(); (void)cudaGetLastError(); } } } sync_guard;

    // Progress (GS:173, GS:262: the reference updates its bar per sub-batch / per frame): chunks are reported in order, on
    // the calling thread, as soon as their results are complete in the caller's memory.
    int reported = 0;
    auto chunk_frames = [&](int ci) { const int f0 = ci * chunk; return (n - f0 < chunk) ? n - f0 : chunk; };
    auto report_done = [&](int upto, bool wait) {    // chunks [reported, upto) -- only those already complete unless `wait`
        while (reported < upto) {
            bool done;
            if (drain) {
                std::unique_lock<std::mutex> g(mu);
                if (wait) cv.wait(g, [&] { return drained > reported || drain_rc != CS_OK; });
                done = drained > reported;
            } else {
                cudaEvent_t ev = cx.ev_out[reported % kSlots];
                done = wait ? (cudaEventSynchronize(ev) == cudaSuccess) : (cudaEventQuery(ev) == cudaSuccess);
            }
            if (!done) break;
            if (progress) progress(chunk_frames(reported), user);
            ++reported;
        }
        (void)cudaGetLastError();   // cudaEventQuery's cudaErrorNotReady is not an error
    };

    for (int it = 0; it < nchunks; ++it) {
        {
            const int f0 = it * chunk, m = (n - f0 < chunk) ? n - f0 : chunk, sl = it % kSlots;
            if (progress && it >= kSlots) report_done(it - kSlots + 1, false);
            if (drain && it >= kSlots) {   // the slot's host landing buffers must have been drained (chunk it - kSlots)
                const auto w0 = std::chrono::steady_clock::now();
                std::unique_lock<std::mutex> g(mu);
                cv.wait(g, [&] { return drained >= it - kSlots + 1; });
                t_slot += std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
            }
            char* din = cx.d_in[sl];
            char* dout = cx.d_out[sl];
            float* d_img = (float*)din;
            float* d_dep = (float*)(din + (size_t)m * b_img);
            float* d_st = (float*)dout;
            float* d_dl = (float*)(dout + (size_t)m * b_st);
            float* d_dr = (float*)(dout + (size_t)m * (b_st + b_d));
            float* d_mk = (float*)(dout + (size_t)m * (b_st + 2 * b_d));
            // upload: the slot's device inputs are free once the kernels of chunk it - kSlots are done; its bounce buffer
            // once the upload of chunk it - kSlots is done
            if (it >= kSlots) HOST_CUDA(cudaStreamWaitEvent(cx.s_in, cx.ev_run[sl], 0));
            if (bounce_in) {
                if (it >= kSlots) HOST_CUDA(cudaEventSynchronize(cx.ev_in[sl]));
                parallel_copy(in_spans(f0, m, cx.h_in[sl], true), team);
                HOST_CUDA(cudaMemcpyAsync(din, cx.h_in[sl], (size_t)m * in_frame, cudaMemcpyHostToDevice, cx.s_in));
            } else {
                HOST_CUDA(cudaMemcpyAsync(d_img, image + (size_t)f0 * px * 3, (size_t)m * b_img, cudaMemcpyHostToDevice, cx.s_in));
                HOST_CUDA(cudaMemcpyAsync(d_dep, depth + (size_t)f0 * dpx * c, (size_t)m * b_dep, cudaMemcpyHostToDevice, cx.s_in));
            }
            HOST_CUDA(cudaEventRecord(cx.ev_in[sl], cx.s_in));
            // kernels: need the upload, and the slot's device outputs must have left (chunk it - kSlots)
            HOST_CUDA(cudaStreamWaitEvent(cx.s_run, cx.ev_in[sl], 0));
            if (it >= kSlots) HOST_CUDA(cudaStreamWaitEvent(cx.s_run, cx.ev_out[sl], 0));
            rc = cs_stereo_batch(p, d_img, d_dep, m, h, w, c, d_st, d_dl, d_dr, d_mk, cx.d_ws, cx.ws_bytes, cx.s_run);
            if (rc) { cudaDeviceSynchronize(); return rc; }
            if (compact) {
                e = launch_compact_outputs(d_dl, d_dr, d_mk, (int64_t)m * px, (int64_t)m * hm * wm, depth_u8 ? 1 : 0,
                                           cx.d_cmp[sl], cx.d_cmp[sl] + cmp_depth,
                                           (uint8_t*)(cx.d_cmp[sl] + 2 * cmp_depth), cx.s_run);
                if (e != cudaSuccess) { cudaDeviceSynchronize(); HOST_FAIL(CS_ERR_CUDA, "compact: %s", cudaGetErrorString(e)); }
            }
            HOST_CUDA(cudaEventRecord(cx.ev_run[sl], cx.s_run));
            // download (the bounce buffers of this slot were drained by the host one iteration ago)
            HOST_CUDA(cudaStreamWaitEvent(cx.s_out, cx.ev_run[sl], 0));
            if (compact) {
                float* dst = stereo_pinned ? stereo + (size_t)f0 * ho * wo * 3 : (float*)cx.h_out[sl];
                HOST_CUDA(cudaMemcpyAsync(dst, d_st, (size_t)m * b_st, cudaMemcpyDeviceToHost, cx.s_out));
                HOST_CUDA(cudaMemcpyAsync(cx.h_cmp[sl], cx.d_cmp[sl], (size_t)m * px * dsz, cudaMemcpyDeviceToHost, cx.s_out));
                HOST_CUDA(cudaMemcpyAsync(cx.h_cmp[sl] + cmp_depth, cx.d_cmp[sl] + cmp_depth, (size_t)m * px * dsz,
                                          cudaMemcpyDeviceToHost, cx.s_out));
                HOST_CUDA(cudaMemcpyAsync(cx.h_cmp[sl] + 2 * cmp_depth, cx.d_cmp[sl] + 2 * cmp_depth, (size_t)m * hm * wm,
                                          cudaMemcpyDeviceToHost, cx.s_out));
            } else if (bounce_out) {
                HOST_CUDA(cudaMemcpyAsync(cx.h_out[sl], dout, (size_t)m * out_frame, cudaMemcpyDeviceToHost, cx.s_out));
            } else {
                HOST_CUDA(cudaMemcpyAsync(stereo + (size_t)f0 * ho * wo * 3, d_st, (size_t)m * b_st, cudaMemcpyDeviceToHost, cx.s_out));
                HOST_CUDA(cudaMemcpyAsync(depth_l + (size_t)f0 * px * 3, d_dl, (size_t)m * b_d, cudaMemcpyDeviceToHost, cx.s_out));
                HOST_CUDA(cudaMemcpyAsync(depth_r + (size_t)f0 * px * 3, d_dr, (size_t)m * b_d, cudaMemcpyDeviceToHost, cx.s_out));
                HOST_CUDA(cudaMemcpyAsync(mask + (size_t)f0 * hm * wm, d_mk, (size_t)m * b_m, cudaMemcpyDeviceToHost, cx.s_out));
            }
            HOST_CUDA(cudaEventRecord(cx.ev_out[sl], cx.s_out));
            if (drain) { { std::lock_guard<std::mutex> g(mu); enqueued = it + 1; } cv.notify_all(); }
        }
    }
    if (progress) report_done(nchunks, true);
    if (drain) {
        std::unique_lock<std::mutex> g(mu);
        cv.wait(g, [&] { return drained >= nchunks || drain_rc != CS_OK; });
        if (drain_rc) HOST_FAIL(CS_ERR_CUDA, "pipeline: download failed");
    }
    if (const char* tr = getenv("COMFYSTEREO_HOST_TRACE"))
        if (atoi(tr))
            fprintf(stderr, "[cs_host] %d frames, %d chunk(s) of %d, team %d, compact %d, bounce in/out %d/%d: total %.2f ms; "
                    "consumer waited %.2f ms for downloads, %.2f ms host copy/expand; producer waited %.2f ms for slots\n",
                    n, nchunks, chunk, team, (int)compact, (int)bounce_in, (int)bounce_out,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - call0).count() * 1e3,
                    t_event * 1e3, t_host * 1e3, t_slot * 1e3);
    HOST_CUDA(cudaStreamSynchronize(cx.s_in));
    HOST_CUDA(cudaStreamSynchronize(cx.s_run));
    HOST_CUDA(cudaStreamSynchronize(cx.s_out));
    e = cudaGetLastError();
    if (e != cudaSuccess) HOST_FAIL(CS_ERR_CUDA, "pipeline: %s", cudaGetErrorString(e));
    sync_guard.ok = true;
    return CS_OK;
#undef HOST_FAIL
#undef HOST_CUDA
}

}  // extern "C"

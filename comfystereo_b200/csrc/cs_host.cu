// cs_host.cu -- cs_stereo_batch_host: the hot path with HOST buffers, the shape in which ComfyUI
// hands IMAGE tensors to the node and expects them back (GS:126, GS:161-171, GS:298-307).
//
// Frames stream through two device slots: while chunk i runs its kernels on the compute stream,
// chunk i+1 is uploading on the copy-in stream and chunk i-1 is downloading on the copy-out stream.
//
// Page-locked caller buffers are copied directly (cudaMemcpyAsync is truly asynchronous on them).
// Pageable buffers -- what ComfyUI passes in, and what a large result tensor has to be -- are not
// handed to the driver (measured on B200: 2.2 GB/s into fresh pageable memory): the library owns
// two page-locked bounce buffers per direction and moves data between them and the caller's memory
// with a team of host threads, overlapped with the GPU work of the neighbouring chunks.
// Device buffers, bounce buffers and streams are cached per device; cs_host_release() frees them.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "cs_internal.cuh"

namespace cs {

struct HostCtx {
    int device = -1;
    size_t in_bytes = 0, out_bytes = 0, ws_bytes = 0;
    size_t bounce_in_bytes = 0, bounce_out_bytes = 0;
    char* d_in[2] = {nullptr, nullptr};
    char* d_out[2] = {nullptr, nullptr};
    char* h_in[2] = {nullptr, nullptr};     // page-locked bounce buffers (only when the caller's memory is pageable)
    char* h_out[2] = {nullptr, nullptr};
    char* d_ws = nullptr;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_run[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    bool ready = false;
};

static std::mutex g_mu[16];   // one pipeline per device; devices run concurrently
static HostCtx g_ctx[16];

static void ctx_free(HostCtx& c) {
    if (!c.ready) return;
    cudaSetDevice(c.device);
    for (int i = 0; i < 2; ++i) {
        if (c.d_in[i]) cudaFree(c.d_in[i]);
        if (c.d_out[i]) cudaFree(c.d_out[i]);
        if (c.h_in[i]) cudaFreeHost(c.h_in[i]);
        if (c.h_out[i]) cudaFreeHost(c.h_out[i]);
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
                              bool bounce_in, bool bounce_out) {
    cudaError_t e;
    const bool fits = c.ready && c.in_bytes >= in_bytes && c.out_bytes >= out_bytes && c.ws_bytes >= ws_bytes &&
                      (!bounce_in || c.bounce_in_bytes >= in_bytes) && (!bounce_out || c.bounce_out_bytes >= out_bytes);
    if (fits) return cudaSuccess;
    const bool keep_in = c.ready && c.bounce_in_bytes > 0, keep_out = c.ready && c.bounce_out_bytes > 0;
    ctx_free(c);
    c.device = device;
    c.ready = true;
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaMalloc((void**)&c.d_in[i], in_bytes)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void**)&c.d_out[i], out_bytes)) != cudaSuccess) return e;
        if (bounce_in || keep_in)
            if ((e = cudaHostAlloc((void**)&c.h_in[i], in_bytes, cudaHostAllocDefault)) != cudaSuccess) return e;
        if (bounce_out || keep_out)
            if ((e = cudaHostAlloc((void**)&c.h_out[i], out_bytes, cudaHostAllocDefault)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&c.ev_in[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&c.ev_run[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&c.ev_out[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    if ((e = cudaMalloc((void**)&c.d_ws, ws_bytes)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&c.s_in, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&c.s_run, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&c.s_out, cudaStreamNonBlocking)) != cudaSuccess) return e;
    c.in_bytes = in_bytes; c.out_bytes = out_bytes; c.ws_bytes = ws_bytes;
    c.bounce_in_bytes = (bounce_in || keep_in) ? in_bytes : 0;
    c.bounce_out_bytes = (bounce_out || keep_out) ? out_bytes : 0;
    return cudaSuccess;
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

struct Span { char* dst; const char* src; size_t bytes; };

// memcpy a list of spans with a team of threads (first touch of fresh pageable pages is the expensive part;
// it parallelises well)
static void parallel_copy(const std::vector<Span>& spans) {
    size_t total = 0;
    for (const auto& s : spans) total += s.bytes;
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<size_t>(std::max(1u, std::min(hw ? hw : 8u, 16u)), total / (4u << 20) + 1);
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

}  // namespace cs

using namespace cs;

extern "C" {

void cs_host_release(void) {
    for (int i = 0; i < 16; ++i) {
        std::lock_guard<std::mutex> lk(g_mu[i]);
        ctx_free(g_ctx[i]);
    }
}

int cs_stereo_batch_host(const cs_params* p, const float* image, const float* depth, int n, int h, int w, int c,
                         float* stereo, float* depth_l, float* depth_r, float* mask, int device) {
#define HOST_FAIL(code, ...) return cs::fail(code, __VA_ARGS__)
#define HOST_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cs::fail(CS_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)
    if (!p || !image || !depth || !stereo || !depth_l || !depth_r || !mask) HOST_FAIL(CS_ERR_ARG, "cs_stereo_batch_host: NULL pointer");
    if (n < 1 || h < 1 || w < 2 || c < 1) HOST_FAIL(CS_ERR_ARG, "cs_stereo_batch_host: bad size");
    if (device < 0 || device >= 16) HOST_FAIL(CS_ERR_ARG, "cs_stereo_batch_host: bad device %d", device);
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
    // chunk: whole GPU-Warp sub-batches (Q9); otherwise ~256 MB of I/O per slot
    int group = (p->fill == CS_FILL_GPU_WARP && p->group_size > 0) ? (p->group_size < n ? p->group_size : n) : 1;
    int chunk = (int)((size_t)256 * 1024 * 1024 / (in_frame + out_frame));
    if (chunk < 1) chunk = 1;
    chunk = (chunk / group) * group;
    if (chunk < group) chunk = group;
    if (chunk > n) chunk = n;
    const size_t ws_bytes = cs_workspace_bytes(p, chunk, h, w);
    const bool bounce_in = !(is_pinned(image) && is_pinned(depth));
    const bool bounce_out = !(is_pinned(stereo) && is_pinned(depth_l) && is_pinned(depth_r) && is_pinned(mask));

    std::lock_guard<std::mutex> lk(g_mu[device]);
    HostCtx& cx = g_ctx[device];
    cudaError_t e = ctx_ensure(cx, device, (size_t)chunk * in_frame, (size_t)chunk * out_frame, ws_bytes, bounce_in, bounce_out);
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
        v.push_back({(char*)(depth_l + (size_t)f0 * px * 3), slot + (size_t)m * b_st, (size_t)m * b_d});
        v.push_back({(char*)(depth_r + (size_t)f0 * px * 3), slot + (size_t)m * (b_st + b_d), (size_t)m * b_d});
        v.push_back({(char*)(mask + (size_t)f0 * hm * wm), slot + (size_t)m * (b_st + 2 * b_d), (size_t)m * b_m});
        return v;
    };

    const int nchunks = (n + chunk - 1) / chunk;
    // iteration `it` enqueues chunk `it` and, meanwhile on the host, drains the results of chunk `it - 1`
    for (int it = 0; it <= nchunks; ++it) {
        if (it < nchunks) {
            const int f0 = it * chunk, m = (n - f0 < chunk) ? n - f0 : chunk, sl = it & 1;
            char* din = cx.d_in[sl];
            char* dout = cx.d_out[sl];
            float* d_img = (float*)din;
            float* d_dep = (float*)(din + (size_t)m * b_img);
            float* d_st = (float*)dout;
            float* d_dl = (float*)(dout + (size_t)m * b_st);
            float* d_dr = (float*)(dout + (size_t)m * (b_st + b_d));
            float* d_mk = (float*)(dout + (size_t)m * (b_st + 2 * b_d));
            // upload: the slot's device inputs are free once the kernels of chunk it-2 are done; its bounce buffer
            // once the upload of chunk it-2 is done
            if (it >= 2) HOST_CUDA(cudaStreamWaitEvent(cx.s_in, cx.ev_run[sl], 0));
            if (bounce_in) {
                if (it >= 2) HOST_CUDA(cudaEventSynchronize(cx.ev_in[sl]));
                parallel_copy(in_spans(f0, m, cx.h_in[sl], true));
                HOST_CUDA(cudaMemcpyAsync(din, cx.h_in[sl], (size_t)m * in_frame, cudaMemcpyHostToDevice, cx.s_in));
            } else {
                HOST_CUDA(cudaMemcpyAsync(d_img, image + (size_t)f0 * px * 3, (size_t)m * b_img, cudaMemcpyHostToDevice, cx.s_in));
                HOST_CUDA(cudaMemcpyAsync(d_dep, depth + (size_t)f0 * dpx * c, (size_t)m * b_dep, cudaMemcpyHostToDevice, cx.s_in));
            }
            HOST_CUDA(cudaEventRecord(cx.ev_in[sl], cx.s_in));
            // kernels: need the upload, and the slot's device outputs must have left (chunk it-2)
            HOST_CUDA(cudaStreamWaitEvent(cx.s_run, cx.ev_in[sl], 0));
            if (it >= 2) HOST_CUDA(cudaStreamWaitEvent(cx.s_run, cx.ev_out[sl], 0));
            rc = cs_stereo_batch(p, d_img, d_dep, m, h, w, c, d_st, d_dl, d_dr, d_mk, cx.d_ws, cx.ws_bytes, cx.s_run);
            if (rc) { cudaDeviceSynchronize(); return rc; }
            HOST_CUDA(cudaEventRecord(cx.ev_run[sl], cx.s_run));
            // download (the bounce buffer of this slot was drained by the host one iteration ago)
            HOST_CUDA(cudaStreamWaitEvent(cx.s_out, cx.ev_run[sl], 0));
            if (bounce_out) {
                HOST_CUDA(cudaMemcpyAsync(cx.h_out[sl], dout, (size_t)m * out_frame, cudaMemcpyDeviceToHost, cx.s_out));
            } else {
                HOST_CUDA(cudaMemcpyAsync(stereo + (size_t)f0 * ho * wo * 3, d_st, (size_t)m * b_st, cudaMemcpyDeviceToHost, cx.s_out));
                HOST_CUDA(cudaMemcpyAsync(depth_l + (size_t)f0 * px * 3, d_dl, (size_t)m * b_d, cudaMemcpyDeviceToHost, cx.s_out));
                HOST_CUDA(cudaMemcpyAsync(depth_r + (size_t)f0 * px * 3, d_dr, (size_t)m * b_d, cudaMemcpyDeviceToHost, cx.s_out));
                HOST_CUDA(cudaMemcpyAsync(mask + (size_t)f0 * hm * wm, d_mk, (size_t)m * b_m, cudaMemcpyDeviceToHost, cx.s_out));
            }
            HOST_CUDA(cudaEventRecord(cx.ev_out[sl], cx.s_out));
        }
        if (bounce_out && it >= 1) {
            const int pit = it - 1, f0 = pit * chunk, m = (n - f0 < chunk) ? n - f0 : chunk, sl = pit & 1;
            HOST_CUDA(cudaEventSynchronize(cx.ev_out[sl]));
            parallel_copy(out_spans(f0, m, cx.h_out[sl]));
        }
    }
    HOST_CUDA(cudaStreamSynchronize(cx.s_in));
    HOST_CUDA(cudaStreamSynchronize(cx.s_run));
    HOST_CUDA(cudaStreamSynchronize(cx.s_out));
    e = cudaGetLastError();
    if (e != cudaSuccess) HOST_FAIL(CS_ERR_CUDA, "pipeline: %s", cudaGetErrorString(e));
    return CS_OK;
#undef HOST_FAIL
#undef HOST_CUDA
}

}  // extern "C"

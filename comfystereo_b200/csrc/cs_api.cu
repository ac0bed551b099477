// cs_api.cu -- the C ABI (include/comfystereo_b200.h): argument checks, workspace carving, the
// per-chunk kernel sequence of the hot path, and the host-buffer pipeline.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "cs_internal.cuh"

namespace cs {

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
    static std::atomic<int> cached[64];      // per device ordinal; 0 = not asked yet
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = 148;
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// ---- per-kernel timing -------------------------------------------------------------------
struct ProfRec { int id; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::atomic<int> g_prof_on{0};
static const char* const kKernelNames[K_COUNT] = {
    "k_prepare", "k_edge_dist", "k_blur_blend", "k_depth_out", "k_warp_rows", "k_polylines", "k_polylines_exact",
    "k_hybrid_splat", "k_hybrid_gapfill", "k_gpuwarp", "k_compose", "misc", "k_resize_gray"};
void prof_begin(int id, cudaStream_t s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec r;
    r.id = id;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
}
void prof_end(int id, cudaStream_t s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (size_t i = g_prof.size(); i-- > 0;)
        if (g_prof[i].id == id) { cudaEventRecord(g_prof[i].b, s); break; }
}

static thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
    return fail(CS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CS_CUDA(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, what); } while (0)

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

static bool is_cpu_technique(int fill) {
    return (fill >= CS_FILL_NONE && fill <= CS_FILL_HYBRID_EDGE) || (fill >= CS_FILL_NONE_POST && fill <= CS_FILL_HYBRID_EDGE_PLUS);
}

static bool is_gpu_warp(int fill) { return fill == CS_FILL_GPU_WARP || fill == CS_FILL_GPU_WARP_MESH; }

static int check_params(const cs_params* p) {
    if (!p) return fail(CS_ERR_ARG, "params is NULL");
    if (p->fill < CS_FILL_NONE || p->fill > CS_FILL_GPU_WARP_MESH) return fail(CS_ERR_ARG, "unknown fill %d", p->fill);
    if (p->mode < CS_MODE_LEFT_RIGHT || p->mode > CS_MODE_CYAN_RED) return fail(CS_ERR_MODE, "Unknown mode");
    if (p->depth_h < 0 || p->depth_w < 0 || (p->depth_h > 0) != (p->depth_w > 0))
        return fail(CS_ERR_ARG, "depth_h/depth_w must both be 0 or both be positive");
    if (p->blur_flavor != 0 && p->blur_flavor != 1) return fail(CS_ERR_ARG, "unknown blur_flavor %d", p->blur_flavor);
    if (p->blur_enabled) {
        if (p->blur_box < 1) return fail(CS_ERR_UNSUPPORTED, "kernel size should be greater than zero");
        if (p->blur_radius < 0 || p->blur_radius > kMaxBlurRadius)
            return fail(CS_ERR_UNSUPPORTED, "blur radius %d outside 0..%d", p->blur_radius, kMaxBlurRadius);
        if (p->blur_vert_smooth < 0) return fail(CS_ERR_ARG, "negative vert_smooth");
        if (p->blur_vert_smooth > 15)
            return fail(CS_ERR_UNSUPPORTED, "depth_blur_vert_smooth %d: this library supports 0..15 (the widget's range)", p->blur_vert_smooth);
        if (p->blur_box > 4096) return fail(CS_ERR_UNSUPPORTED, "blur box %d wider than 4096 px is not supported", p->blur_box);
    }
    return CS_OK;
}

static void out_dims(int mode, int h, int w, int* ho, int* wo) {
    *ho = h; *wo = w;
    if (mode == CS_MODE_LEFT_RIGHT || mode == CS_MODE_RIGHT_LEFT) *wo = 2 * w;
    if (mode == CS_MODE_TOP_BOTTOM || mode == CS_MODE_BOTTOM_TOP) *ho = 2 * h;
}

// Per-eye signed divergence / separation in pixels (SIG:1533-1541, SIG:1602-1603, SIG:1060-1065).
static void eye_specs(const cs_params* p, int w, EyeSpec eye[2]) {
    const double ldiv = p->divergence * (1 + p->stereo_balance);
    const double rdiv = p->divergence * (1 - p->stereo_balance);
    const double sep = p->separation;
    eye[0].passthrough = ldiv < 0.001;
    eye[1].passthrough = rdiv < 0.001;
    if (is_gpu_warp(p->fill)) {
        const double ldiv_px = (ldiv / 100.0) * w, rdiv_px = (rdiv / 100.0) * w, sep_px = (sep / 100.0) * w;
        eye[0].div_px = +ldiv_px; eye[0].sep_px = -sep_px;
        eye[1].div_px = -rdiv_px; eye[1].sep_px = sep_px;
    } else {
        eye[0].div_px = ((+1 * ldiv) / 100.0) * w; eye[0].sep_px = ((-1 * sep) / 100.0) * w;
        eye[1].div_px = ((-1 * rdiv) / 100.0) * w; eye[1].sep_px = (sep / 100.0) * w;
    }
}

static bool needs_resize(const cs_params* p, int h, int w) {
    return p->depth_h > 0 && p->depth_w > 0 && (p->depth_h != h || p->depth_w != w);
}

struct Workspace {
    FrameStats* stats;
    float* resized;      // N1: the gray depth at image size, only when the depth frames have another size
    float* gray;
    uint32_t* image_u8;
    float* blur_l;
    float* blur_r;
    uint8_t* dist;
    uint32_t* eye_out[2];
    void* warp_scratch;
    size_t warp_scratch_bytes;
    void* row_scratch;       // GPU Warp rows too wide for shared memory (gpuwarp_row_scratch_bytes)
    size_t total;
};

static Workspace carve(const cs_params* p, int chunk, int h, int w, void* base) {
    Workspace ws;
    memset(&ws, 0, sizeof(ws));
    const size_t px = (size_t)chunk * h * w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return (char*)base + o; };
    ws.stats = (FrameStats*)take((size_t)chunk * sizeof(FrameStats));
    if (needs_resize(p, h, w)) ws.resized = (float*)take(px * 4);
    ws.gray = (float*)take(px * 4);
    const bool cpu = is_cpu_technique(p->fill);
    if (cpu) ws.image_u8 = (uint32_t*)take(px * 4);
    if (p->blur_enabled) {
        ws.blur_l = (float*)take(px * 4);
        ws.blur_r = (float*)take(px * 4);
        ws.dist = (uint8_t*)take(px * 2);
    }
    if (cpu) {
        ws.eye_out[0] = (uint32_t*)take(px * 4);
        ws.eye_out[1] = (uint32_t*)take(px * 4);
        if (p->fill == CS_FILL_POLYLINES_SOFT || p->fill == CS_FILL_POLYLINES_SHARP) {
            ws.warp_scratch_bytes = polylines_scratch_bytes(chunk, h, w);
            ws.warp_scratch = take(ws.warp_scratch_bytes);
        } else if (p->fill == CS_FILL_HYBRID_EDGE_PLUS) {
            ws.warp_scratch_bytes = hybrid_plus_scratch_bytes(chunk, h, w);
            ws.warp_scratch = take(ws.warp_scratch_bytes);
        }
    } else if (p->fill == CS_FILL_GPU_WARP_MESH) {
        ws.warp_scratch_bytes = mesh_keep_bytes(chunk, h, w);   // the culled topology, one per sub-batch and eye
        ws.warp_scratch = take(ws.warp_scratch_bytes);
    }
    if (!cpu && gpuwarp_row_scratch_bytes(w)) ws.row_scratch = take(gpuwarp_row_scratch_bytes(w));
    ws.total = off;
    return ws;
}

static cudaError_t launch_fill(const WarpArgs& a, cudaStream_t s) {
    switch (a.fill) {
        case CS_FILL_NONE: case CS_FILL_NAIVE: case CS_FILL_NAIVE_INTERP: case CS_FILL_INVERSE:
        case CS_FILL_NONE_POST: case CS_FILL_INVERSE_POST:
            return launch_warp_rows(a, s);
        case CS_FILL_HYBRID_EDGE_PLUS:
            return launch_hybrid_plus(a, s);
        case CS_FILL_POLYLINES_SOFT: case CS_FILL_POLYLINES_SHARP:
            return launch_polylines(a, s);
        case CS_FILL_HYBRID_EDGE:
            return launch_hybrid(a, s);
        default:
            return cudaErrorInvalidValue;
    }
}

// One chunk of frames, all on `s`.  Pointers are already offset to the chunk's first frame.
static int run_chunk(const cs_params* p, const float* image, const float* depth, int n, int h, int w, int c,
                     float* stereo, float* depth_l, float* depth_r, float* mask, const Workspace& ws,
                     int flags, cudaStream_t s) {
    const bool cpu = is_cpu_technique(p->fill);
    const int group = p->group_size > 0 ? p->group_size : n;
    EyeSpec eye[2];
    eye_specs(p, w, eye);
    CS_CUDA(launch_init_stats(ws.stats, n, s), "init_stats");
    // N1 resize: depth frames of another size become a 1-channel gray depth at image size first
    if (needs_resize(p, h, w)) {
        CS_CUDA(launch_resize_gray(depth, n, p->depth_h, p->depth_w, c, h, w, ws.resized, s), "resize");
        depth = ws.resized;
        c = 1;
    }
    // N1 + L1 (+ O1 input side for the CPU techniques)
    CS_CUDA(launch_prepare(cpu ? image : nullptr, depth, n, h, w, c, ws.gray, cpu ? ws.image_u8 : nullptr, ws.stats, s),
            "prepare");
    const float* dl = ws.gray;
    const float* dr = ws.gray;
    if (p->blur_enabled) {
        // B1; CPU techniques decide x255 per frame (SIG:1475), GPU Warp per sub-batch (SIG:1045)
        CS_CUDA(launch_blur(ws.gray, ws.stats, cpu ? 1 : 2, group, n, h, w, *p, ws.blur_l, ws.blur_r, ws.dist,
                            cpu ? depth_l : nullptr, cpu ? depth_r : nullptr, s), "blur");
        dl = ws.blur_l; dr = ws.blur_r;
        if (!cpu) CS_CUDA(launch_depth_out(dl, dr, ws.stats, n, h, w, 1, 2, group, depth_l, depth_r, s), "depth_out");
    } else {
        CS_CUDA(launch_depth_out(dl, dr, ws.stats, n, h, w, 0, cpu ? 1 : 2, group, depth_l, depth_r, s), "depth_out");
    }
    if (cpu) {
        WarpArgs a;
        memset(&a, 0, sizeof(a));
        a.image_u8 = ws.image_u8;
        a.depth[0] = dl; a.depth[1] = dr;
        a.stats = ws.stats;
        a.use_blur_stats = p->blur_enabled ? 1 : 0;
        a.scale_by_stats = p->blur_enabled ? 0 : 1;
        a.out[0] = ws.eye_out[0]; a.out[1] = ws.eye_out[1];
        a.n = n; a.h = h; a.w = w;
        a.fill = p->fill;
        a.eye[0] = eye[0]; a.eye[1] = eye[1];
        a.expo = p->stereo_offset_exponent;
        a.conv = (float)p->convergence_point;
        a.scratch = ws.warp_scratch; a.scratch_bytes = ws.warp_scratch_bytes;
        a.flags = flags;
        // Polylines, the row techniques and Hybrid Edge (its gap-fill pass) write the composed tensors themselves in the
        // side-by-side / top-bottom modes (when both eyes are warped)
        const bool rows = p->fill == CS_FILL_NONE || p->fill == CS_FILL_NAIVE || p->fill == CS_FILL_NAIVE_INTERP ||
                          p->fill == CS_FILL_INVERSE || p->fill == CS_FILL_NONE_POST || p->fill == CS_FILL_INVERSE_POST;
        const bool fused = (rows || p->fill == CS_FILL_POLYLINES_SOFT || p->fill == CS_FILL_POLYLINES_SHARP ||
                            (p->fill == CS_FILL_HYBRID_EDGE && w % 4 == 0)) &&
                           p->mode >= CS_MODE_LEFT_RIGHT && p->mode <= CS_MODE_BOTTOM_TOP &&
                           !eye[0].passthrough && !eye[1].passthrough && !(flags & 16);
        if (fused) { a.fused_stereo = stereo; a.fused_mask = mask; a.fused_mode = p->mode; }
        if (!(eye[0].passthrough && eye[1].passthrough)) CS_CUDA(launch_fill(a, s), "warp/fill");
        if (!fused) {
            // an eye whose divergence is < 0.001 is the quantised input itself (SIG:1536, 1539)
            const uint32_t* L = eye[0].passthrough ? ws.image_u8 : ws.eye_out[0];
            const uint32_t* R = eye[1].passthrough ? ws.image_u8 : ws.eye_out[1];
            CS_CUDA(launch_compose(L, R, n, h, w, p->mode, stereo, mask, s), "compose");
        }
    } else {
        GpuWarpArgs g;
        memset(&g, 0, sizeof(g));
        g.image = image;
        g.depth[0] = dl; g.depth[1] = dr;
        g.stats = ws.stats;
        g.use_blur_stats = p->blur_enabled ? 1 : 0;
        g.prescale = 1;
        g.group = group;
        g.n = n; g.h = h; g.w = w; g.mode = p->mode;
        g.eye[0] = eye[0]; g.eye[1] = eye[1];
        g.expo = (float)p->stereo_offset_exponent;
        g.conv = (float)p->convergence_point;
        g.stereo = stereo; g.mask = mask;
        g.row_scratch = ws.row_scratch; g.row_scratch_stride = gpuwarp_row_scratch_stride(w);
        if (p->fill == CS_FILL_GPU_WARP_MESH) {
            g.keep = (uint8_t*)ws.warp_scratch;
            CS_CUDA(launch_meshwarp(g, s), "meshwarp");
        } else {
            CS_CUDA(launch_gpuwarp(g, s), "gpuwarp");
        }
    }
    return CS_OK;
}

static std::atomic<int> g_test_flags{0};

// ---- CUDA-graph replay of small repeated jobs ------------------------------------------------------
struct GraphKey {
    cs_params params;
    const void* ptr[7];
    size_t workspace_bytes;
    int n, h, w, c, flags, device;
};
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec;
    long long launches;
    bool failed;
};
constexpr size_t kGraphMaxPixels = (size_t)2 * 1920 * 1080;   // beyond two HD frames the kernels dwarf the launches
constexpr size_t kGraphEntries = 16;
static std::mutex g_graph_mu;
static std::vector<GraphEntry> g_graphs;       // most recently used last
static std::vector<cudaStream_t> g_cap_streams;
static bool graphs_enabled() {
    static const bool on = [] { const char* e = getenv("COMFYSTEREO_GRAPHS"); return !(e && e[0] == '0'); }();
    return on;
}
static GraphEntry* graph_lookup(const GraphKey& k) {
    for (size_t i = g_graphs.size(); i-- > 0;)
        if (memcmp(&g_graphs[i].key, &k, sizeof(k)) == 0) {
            if (i + 1 != g_graphs.size()) { GraphEntry e = g_graphs[i]; g_graphs.erase(g_graphs.begin() + i); g_graphs.push_back(e); }
            return &g_graphs.back();
        }
    return nullptr;
}
static void graph_insert(const GraphKey& k) {
    if (g_graphs.size() >= kGraphEntries) {
        if (g_graphs.front().exec) cudaGraphExecDestroy(g_graphs.front().exec);
        g_graphs.erase(g_graphs.begin());
    }
    GraphEntry e;
    e.key = k; e.exec = nullptr; e.launches = 0; e.failed = false;
    g_graphs.push_back(e);
}
static cudaStream_t graph_capture_stream(int device) {   // capture cannot start on the legacy default stream: a private one per device
    if (device < 0) return nullptr;
    if ((size_t)device >= g_cap_streams.size()) g_cap_streams.resize(device + 1, nullptr);
    if (!g_cap_streams[device] && cudaStreamCreateWithFlags(&g_cap_streams[device], cudaStreamNonBlocking) != cudaSuccess)
        g_cap_streams[device] = nullptr;
    return g_cap_streams[device];
}
void release_graphs() {
    std::lock_guard<std::mutex> lk(g_graph_mu);
    for (auto& e : g_graphs)
        if (e.exec) cudaGraphExecDestroy(e.exec);
    g_graphs.clear();
}


}  // namespace cs

using namespace cs;

extern "C" {

int cs_abi_version(void) { return CS_ABI_VERSION; }
const char* cs_last_error(void) { return g_err; }

int cs_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(CS_ERR_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return fail(CS_ERR_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10) return fail(CS_ERR_DEVICE, "device %d is sm_%d%d; this library is sm_100a only", dev, prop.major, prop.minor);
    return CS_OK;
}

int cs_output_dims(const cs_params* p, int h, int w, int* ho, int* wo, int* hm, int* wm) {
    int rc = check_params(p);
    if (rc) return rc;
    if (h < 1 || w < 1 || !ho || !wo || !hm || !wm) return fail(CS_ERR_ARG, "bad dims");
    out_dims(p->mode, h, w, ho, wo);
    if (is_gpu_warp(p->fill)) { *hm = h; *wm = w; }   // M2: single-eye shape even for SBS
    else { *hm = *ho; *wm = *wo; }                           // M1: shape of the composed image
    return CS_OK;
}

size_t cs_workspace_bytes(const cs_params* p, int chunk, int h, int w) {
    if (!p || chunk < 1 || h < 1 || w < 1) return 0;
    return carve(p, chunk, h, w, nullptr).total;
}

void cs_set_test_flags(int flags) { g_test_flags.store(flags); set_blur_test_flags((flags >> 5) & 1); }

int cs_profile_kernel_count(void) { return K_COUNT; }
const char* cs_profile_kernel_name(int id) { return (id >= 0 && id < K_COUNT) ? kKernelNames[id] : ""; }
void cs_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }
int cs_profile_collect(double* ms, long long* launches) {
    // Call after synchronising the streams that were profiled.  Accumulates into ms[K_COUNT] / launches[K_COUNT].
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 0; i < K_COUNT; ++i) { ms[i] = 0.0; launches[i] = 0; }
    for (auto& r : g_prof) {
        float t = 0.0f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            ms[r.id] += t;
            launches[r.id] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    (void)cudaGetLastError();
    return CS_OK;
}

long long cs_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

int cs_depth_prepare(const float* depth, int n, int h, int w, int c, float* gray, float* minmax, void* stream) {
    if (!depth || !gray || n < 1 || h < 1 || w < 1 || c < 1) return fail(CS_ERR_ARG, "cs_depth_prepare: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    FrameStats* st = nullptr;
    CS_CUDA(cudaMallocAsync((void**)&st, (size_t)n * sizeof(FrameStats), s), "cudaMallocAsync");
    CS_CUDA(launch_init_stats(st, n, s), "init_stats");
    CS_CUDA(launch_prepare(nullptr, depth, n, h, w, c, gray, nullptr, st, s), "prepare");
    if (minmax) CS_CUDA(launch_export_stats(st, n, 0, minmax, 2, s), "export_stats");
    CS_CUDA(cudaFreeAsync(st, s), "cudaFreeAsync");
    return CS_OK;
}

int cs_depth_resize(const float* depth, int n, int dh, int dw, int c, int h, int w, float* gray, void* stream) {
    if (!depth || !gray || n < 1 || dh < 1 || dw < 1 || c < 1 || h < 1 || w < 1)
        return fail(CS_ERR_ARG, "cs_depth_resize: bad argument");
    CS_CUDA(launch_resize_gray(depth, n, dh, dw, c, h, w, gray, (cudaStream_t)stream), "resize");
    return CS_OK;
}

int cs_blur(const float* depth255, int n, int h, int w, const cs_params* p, float* blur_l, float* blur_r,
            float* minmax, uint8_t* dist_scratch, void* stream) {
    if (!depth255 || !blur_l || !blur_r || !dist_scratch || !p || n < 1 || h < 1 || w < 1)
        return fail(CS_ERR_ARG, "cs_blur: bad argument");
    if (p->blur_box < 1) return fail(CS_ERR_UNSUPPORTED, "kernel size should be greater than zero");
    if (p->blur_radius < 0 || p->blur_radius > kMaxBlurRadius) return fail(CS_ERR_UNSUPPORTED, "blur radius out of range");
    cudaStream_t s = (cudaStream_t)stream;
    FrameStats* st = nullptr;
    CS_CUDA(cudaMallocAsync((void**)&st, (size_t)n * sizeof(FrameStats), s), "cudaMallocAsync");
    CS_CUDA(launch_init_stats(st, n, s), "init_stats");
    CS_CUDA(launch_blur(depth255, st, 0, 1, n, h, w, *p, blur_l, blur_r, dist_scratch, nullptr, nullptr, s), "blur");
    if (minmax) CS_CUDA(launch_export_stats(st, n, 1, minmax, 4, s), "export_stats");
    CS_CUDA(cudaFreeAsync(st, s), "cudaFreeAsync");
    return CS_OK;
}

int cs_shift_indices(const float* nd, int n, int h, int w, double div_px, double sep_px, double exponent,
                     int kind, int32_t* out, void* stream) {
    if (!nd || !out || n < 1 || h < 1 || w < 1 || kind < 0 || kind > 1) return fail(CS_ERR_ARG, "cs_shift_indices: bad argument");
    CS_CUDA(launch_shift_indices(nd, n, h, w, div_px, sep_px, exponent, kind, out, (cudaStream_t)stream), "shift_indices");
    return CS_OK;
}

size_t cs_warp_fill_scratch_bytes(int n, int h, int w) {
    (void)w;
    return align_up((size_t)n * sizeof(FrameStats)) + align_up(polylines_scratch_bytes(n, h, w)) +
           align_up(hybrid_plus_scratch_bytes(n, h, w));
}

int cs_warp_fill(const uint8_t* image_u8, const float* depth, int n, int h, int w, int fill, double divergence,
                 double separation, double exponent, double convergence, uint8_t* out_u8, void* scratch,
                 size_t scratch_bytes, void* stream) {
    if (!image_u8 || !depth || !out_u8 || !scratch || n < 1 || h < 1 || w < 1)
        return fail(CS_ERR_ARG, "cs_warp_fill: bad argument");
    if (!is_cpu_technique(fill)) return fail(CS_ERR_ARG, "cs_warp_fill: fill %d is not a CPU technique", fill);
    if (scratch_bytes < cs_warp_fill_scratch_bytes(n, h, w)) return fail(CS_ERR_WORKSPACE, "cs_warp_fill: scratch too small");
    cudaStream_t s = (cudaStream_t)stream;
    FrameStats* st = (FrameStats*)scratch;
    char* rest = (char*)scratch + align_up((size_t)n * sizeof(FrameStats));
    CS_CUDA(launch_init_stats(st, n, s), "init_stats");
    CS_CUDA(launch_minmax(depth, n, (int64_t)h * w, st, s), "minmax");
    CS_CUDA(cudaMemsetAsync(rest, 0, 64, s), "memset");
    WarpArgs a;
    memset(&a, 0, sizeof(a));
    a.image_u8 = (const uint32_t*)image_u8;
    a.depth[0] = depth; a.depth[1] = depth;
    a.stats = st;
    a.use_blur_stats = 0; a.scale_by_stats = 0;   // apply_stereo_divergence takes the depth as given
    a.out[0] = (uint32_t*)out_u8; a.out[1] = nullptr;
    a.n = n; a.h = h; a.w = w; a.fill = fill;
    a.eye[0].div_px = (divergence / 100.0) * w;
    a.eye[0].sep_px = (separation / 100.0) * w;
    a.eye[0].passthrough = 0;
    a.eye[1].passthrough = 1;
    a.expo = exponent;
    a.conv = (float)convergence;
    a.scratch = rest; a.scratch_bytes = polylines_scratch_bytes(n, h, w);
    if (fill == CS_FILL_HYBRID_EDGE_PLUS) {
        a.scratch = rest + align_up(polylines_scratch_bytes(n, h, w));
        a.scratch_bytes = hybrid_plus_scratch_bytes(n, h, w);
    }
    a.flags = g_test_flags.load();
    cudaError_t e = launch_fill(a, s);
    if (e != cudaSuccess) return cuda_fail(e, "warp/fill");
    return CS_OK;
}

size_t cs_forward_warp_scratch_bytes(int n, int h, int w) {
    if (n < 1 || h < 1 || w < 1) return 0;
    return align_up((size_t)n * sizeof(FrameStats)) + align_up(gpuwarp_row_scratch_bytes(w));
}

int cs_forward_warp(const float* image, const float* depth, int n, int h, int w, double div_px, double sep_px,
                    double exponent, double convergence, float* warped, float* mask, void* scratch,
                    size_t scratch_bytes, void* stream) {
    if (!image || !depth || !warped || !mask || !scratch || n < 1 || h < 1 || w < 1)
        return fail(CS_ERR_ARG, "cs_forward_warp: bad argument");
    if (scratch_bytes < cs_forward_warp_scratch_bytes(n, h, w)) return fail(CS_ERR_WORKSPACE, "cs_forward_warp: scratch too small");
    if (w > 24000) return fail(CS_ERR_UNSUPPORTED, "cs_forward_warp: width %d exceeds the 24000-pixel row capacity", w);
    cudaStream_t s = (cudaStream_t)stream;
    FrameStats* st = (FrameStats*)scratch;
    CS_CUDA(launch_init_stats(st, n, s), "init_stats");
    CS_CUDA(launch_minmax(depth, n, (int64_t)h * w, st, s), "minmax");
    GpuWarpArgs g;
    memset(&g, 0, sizeof(g));
    g.image = image;
    g.depth[0] = depth; g.depth[1] = depth;
    g.stats = st;
    g.use_blur_stats = 0; g.prescale = 0; g.group = n;   // SIG:314-316: "/255 if ANY frame of the batch has max > 1"
    g.n = n; g.h = h; g.w = w; g.mode = CS_MODE_LEFT_ONLY;
    g.eye[0].div_px = div_px; g.eye[0].sep_px = sep_px; g.eye[0].passthrough = 0;
    g.eye[1].passthrough = 1;
    g.expo = (float)exponent; g.conv = (float)convergence;
    g.stereo = warped; g.mask = mask;
    g.row_scratch = (char*)scratch + align_up((size_t)n * sizeof(FrameStats));
    g.row_scratch_stride = gpuwarp_row_scratch_stride(w);
    CS_CUDA(launch_gpuwarp(g, s), "gpuwarp");
    return CS_OK;
}

size_t cs_forward_warp_mesh_scratch_bytes(int n, int h, int w) {
    if (n < 1 || h < 1 || w < 1) return 0;
    return align_up((size_t)n * sizeof(FrameStats)) + align_up(mesh_keep_bytes(n, h, w)) + align_up(gpuwarp_row_scratch_bytes(w));
}

int cs_forward_warp_mesh(const float* image, const float* depth, int n, int h, int w, double div_px, double sep_px,
                    double exponent, double convergence, float* warped, float* mask, void* scratch,
                    size_t scratch_bytes, void* stream) {
    if (!image || !depth || !warped || !mask || !scratch || n < 1 || h < 1 || w < 1)
        return fail(CS_ERR_ARG, "cs_forward_warp_mesh: bad argument");
    if (scratch_bytes < cs_forward_warp_mesh_scratch_bytes(n, h, w))
        return fail(CS_ERR_WORKSPACE, "cs_forward_warp_mesh: scratch too small");
    if (w > 24000) return fail(CS_ERR_UNSUPPORTED, "cs_forward_warp_mesh: width %d exceeds the 24000-pixel row capacity", w);
    cudaStream_t s = (cudaStream_t)stream;
    FrameStats* st = (FrameStats*)scratch;
    CS_CUDA(launch_init_stats(st, n, s), "init_stats");
    CS_CUDA(launch_minmax(depth, n, (int64_t)h * w, st, s), "minmax");
    GpuWarpArgs g;
    memset(&g, 0, sizeof(g));
    g.image = image;
    g.depth[0] = depth; g.depth[1] = depth;
    g.stats = st;
    g.use_blur_stats = 0; g.prescale = 0; g.group = n;   // SIG:314-316: "/255 if ANY frame of the batch has max > 1"
    g.n = n; g.h = h; g.w = w; g.mode = CS_MODE_LEFT_ONLY;
    g.eye[0].div_px = div_px; g.eye[0].sep_px = sep_px; g.eye[0].passthrough = 0;
    g.eye[1].passthrough = 1;
    g.expo = (float)exponent; g.conv = (float)convergence;
    g.stereo = warped; g.mask = mask;
    g.keep = (uint8_t*)scratch + align_up((size_t)n * sizeof(FrameStats));
    g.row_scratch = (char*)g.keep + align_up(mesh_keep_bytes(n, h, w));
    g.row_scratch_stride = gpuwarp_row_scratch_stride(w);
    CS_CUDA(launch_meshwarp(g, s), "meshwarp");
    return CS_OK;
}

int cs_quantize_image(const float* image, int n, int h, int w, uint8_t* image_u8, void* stream) {
    if (!image || !image_u8 || n < 1 || h < 1 || w < 1) return fail(CS_ERR_ARG, "cs_quantize_image: bad argument");
    CS_CUDA(launch_quantize(image, (int64_t)n * h * w, (uint32_t*)image_u8, (cudaStream_t)stream), "quantize");
    return CS_OK;
}

int cs_compose(const uint8_t* left_u8, const uint8_t* right_u8, int n, int h, int w, int mode, float* stereo,
               float* mask, void* stream) {
    if (!left_u8 || !right_u8 || !stereo || !mask || n < 1 || h < 1 || w < 1) return fail(CS_ERR_ARG, "cs_compose: bad argument");
    if (mode < CS_MODE_LEFT_RIGHT || mode > CS_MODE_CYAN_RED) return fail(CS_ERR_MODE, "Unknown mode");
    CS_CUDA(launch_compose((const uint32_t*)left_u8, (const uint32_t*)right_u8, n, h, w, mode, stereo, mask,
                           (cudaStream_t)stream), "compose");
    return CS_OK;
}

int cs_stereo_batch(const cs_params* p, const float* image, const float* depth, int n, int h, int w, int c,
                    float* stereo, float* depth_l, float* depth_r, float* mask, void* workspace,
                    size_t workspace_bytes, void* stream) {
    int rc = check_params(p);
    if (rc) return rc;
    if (!image || !depth || !stereo || !depth_l || !depth_r || !mask || !workspace)
        return fail(CS_ERR_ARG, "cs_stereo_batch: NULL pointer");
    if (n < 1 || h < 1 || w < 2 || c < 1) return fail(CS_ERR_ARG, "cs_stereo_batch: bad size n=%d h=%d w=%d c=%d", n, h, w, c);
    if (p->blur_flavor != 0)
        return fail(CS_ERR_UNSUPPORTED, "cs_stereo_batch: blur_flavor 1 (the scipy blur of non-tensor inputs) is a cs_blur option; "
                    "the node's path always uses the torch blur");
    if ((uintptr_t)workspace % 256) return fail(CS_ERR_ARG, "cs_stereo_batch: workspace must be 256-byte aligned");
    const bool cpu = is_cpu_technique(p->fill);
    if (cpu && c != 1 && c != 3) return fail(CS_ERR_ARG, "cs_stereo_batch: depth must have 1 or 3 channels");
    // one image row lives in shared memory (227 KB per CTA): say so up front instead of failing at launch
    {
        const bool sharp = p->fill == CS_FILL_POLYLINES_SHARP;
        const bool soft = p->fill == CS_FILL_POLYLINES_SOFT || p->fill == CS_FILL_HYBRID_EDGE_PLUS;
        // Polylines: rows of any width are tiled (only the 16-bit point indices of the sequential fallback bound them);
        // the other techniques keep one row per CTA in shared memory
        const bool plus = p->fill == CS_FILL_HYBRID_EDGE_PLUS;
        const int wmax = sharp ? 32766 : (soft && !plus ? 65533 : (is_gpu_warp(p->fill) ? 24000 : 18000));
        if (w > wmax)
            return fail(CS_ERR_UNSUPPORTED, "cs_stereo_batch: width %d exceeds the %d-pixel row capacity of this technique "
                        "(one row per CTA in shared memory)", w, wmax);
    }
    // largest chunk that fits; GPU Warp couples frames inside a sub-batch (Q9), so chunks are whole sub-batches
    const int group = (!cpu && p->group_size > 0) ? (p->group_size < n ? p->group_size : n) : 1;
    const size_t per = cs_workspace_bytes(p, group, h, w);
    if (per == 0 || workspace_bytes < per)
        return fail(CS_ERR_WORKSPACE, "cs_stereo_batch: workspace %zu B < %zu B needed for %d frame(s)", workspace_bytes, per, group);
    int chunk = group;
    // (frames are a grid dimension of every kernel: stay well inside its 65535 limit)
    while (chunk + group <= n && chunk + group <= 16384 && cs_workspace_bytes(p, chunk + group, h, w) <= workspace_bytes)
        chunk += group;
    int ho, wo, hm, wm;
    cs_output_dims(p, h, w, &ho, &wo, &hm, &wm);
    cudaStream_t s = (cudaStream_t)stream;
    auto enqueue = [&](cudaStream_t q) -> int {
        for (int f0 = 0; f0 < n; f0 += chunk) {
            const int m = (n - f0 < chunk) ? n - f0 : chunk;
            Workspace ws = carve(p, m, h, w, workspace);
            const size_t px = (size_t)h * w;
            const size_t dpx = needs_resize(p, h, w) ? (size_t)p->depth_h * p->depth_w : px;
            int r = run_chunk(p, image + (size_t)f0 * px * 3, depth + (size_t)f0 * dpx * c, m, h, w, c,
                              stereo + (size_t)f0 * ho * wo * 3, depth_l + (size_t)f0 * px * 3, depth_r + (size_t)f0 * px * 3,
                              mask + (size_t)f0 * hm * wm, ws, g_test_flags.load(), q);
            if (r) return r;
        }
        return CS_OK;
    };
    // Small jobs (a frame or two) are bound by launch latency, not by the kernels: the second identical call -- same
    // buffers, same parameters, what a streaming caller that reuses its tensors makes -- is captured into a CUDA graph
    // and every later one replays it with a single launch.
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;     // a caller building its own graph gets plain launches
    if (cudaStreamIsCapturing(s, &capturing) != cudaSuccess) { (void)cudaGetLastError(); capturing = cudaStreamCaptureStatusActive; }
    if (graphs_enabled() && (size_t)n * h * w <= kGraphMaxPixels && !g_prof_on.load(std::memory_order_relaxed) &&
        capturing == cudaStreamCaptureStatusNone) {
        GraphKey key;
        memset(&key, 0, sizeof(key));
        key.params = *p;
        key.ptr[0] = image; key.ptr[1] = depth; key.ptr[2] = stereo; key.ptr[3] = depth_l; key.ptr[4] = depth_r;
        key.ptr[5] = mask; key.ptr[6] = workspace;
        key.workspace_bytes = workspace_bytes;
        key.n = n; key.h = h; key.w = w; key.c = c; key.flags = g_test_flags.load();
        cudaGetDevice(&key.device);
        std::lock_guard<std::mutex> lk(g_graph_mu);
        GraphEntry* e = graph_lookup(key);
        if (e && e->exec) {
            CS_CUDA(cudaGraphLaunch(e->exec, s), "cudaGraphLaunch");
            g_launches.fetch_add(e->launches, std::memory_order_relaxed);
            return CS_OK;
        }
        if (e && !e->failed) {
            cudaStream_t cap = graph_capture_stream(key.device);
            cudaGraph_t graph = nullptr;
            if (cap && cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                const long long before = g_launches.load();
                const int r = enqueue(cap);
                const cudaError_t ce = cudaStreamEndCapture(cap, &graph);
                const long long captured = g_launches.exchange(before) - before;   // nothing has run yet
                if (r == CS_OK && ce == cudaSuccess && graph &&
                    cudaGraphInstantiate(&e->exec, graph, 0) == cudaSuccess) {
                    e->launches = captured;
                    cudaGraphDestroy(graph);
                    CS_CUDA(cudaGraphLaunch(e->exec, s), "cudaGraphLaunch");
                    g_launches.fetch_add(e->launches, std::memory_order_relaxed);
                    return CS_OK;
                }
                if (graph) cudaGraphDestroy(graph);
                e->exec = nullptr;
                (void)cudaGetLastError();
            }
            e->failed = true;        // this configuration does not capture: launch it directly from now on
        } else if (!e) {
            graph_insert(key);
        }
    }
    return enqueue(s);
}

int cs_polylines_status(const cs_params* p, int chunk, int h, int w, const void* workspace, int* status_out, int* flagged_rows) {
    // Debug/telemetry: reads back the status word and the number of rows that needed the exact replay
    // in the most recent chunk.  Synchronous.
    if (!p || !workspace) return fail(CS_ERR_ARG, "cs_polylines_status: bad argument");
    Workspace ws = carve(p, chunk, h, w, const_cast<void*>(workspace));
    if (!ws.warp_scratch || (p->fill != CS_FILL_POLYLINES_SOFT && p->fill != CS_FILL_POLYLINES_SHARP))
        return fail(CS_ERR_ARG, "cs_polylines_status: not a polylines configuration");
    // scratch layout (cs_polylines.cu): [0] status bits, [1] rows listed for the sequential kernel, ...
    int host[2] = {0, 0};
    cudaError_t e = cudaMemcpy(host, ws.warp_scratch, sizeof(host), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
    if (status_out) *status_out = host[0];
    if (flagged_rows) *flagged_rows = host[1];
    return CS_OK;
}

}  // extern "C"

// context.cu — context, device memory and host<->device plumbing of the resvg_b200 C ABI.
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>

#include "common.cuh"
#include <algorithm>

#include "rb_internal.h"

int rb_filters_init(rb_ctx *ctx);

int rb_fail(rb_ctx *ctx, int code, const char *what)
{
    if (ctx) ctx->err = what ? what : "";
    return code;
}

int rb_cuda_fail(rb_ctx *ctx, cudaError_t e, const char *what)
{
    if (ctx) {
        ctx->err = std::string(what ? what : "") + ": " + cudaGetErrorString(e);
        fprintf(stderr, "[resvg_b200] CUDA error: %s\n", ctx->err.c_str());
    }
    return e == cudaErrorMemoryAllocation ? RB_ERR_OOM : RB_ERR_CUDA;
}

rb_host_prof g_rb_prof = {{0}, {0}, false};
static double rb_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
rb_prof_scope::rb_prof_scope(int kind) : k(kind), t0(g_rb_prof.on ? rb_now() : 0.0) {}
rb_prof_scope::~rb_prof_scope() { if (g_rb_prof.on) { g_rb_prof.t[k] += rb_now() - t0; g_rb_prof.n[k]++; } }

int rb_check_flags(rb_ctx *ctx)
{
    if (!ctx || !ctx->h_flags || !ctx->h_flags[0]) return RB_OK;
    ctx->h_flags[0] = 0;
    return rb_fail(ctx, RB_ERR_CUDA, "tile-row edge list overflow: edges were dropped by an earlier batch");
}

int rb_scratch(rb_ctx *ctx, size_t bytes, void **out)
{
    if (bytes > ctx->scratch_bytes) {
        // The old block may still be in use by enqueued kernels.
        RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) RB_CUDA(ctx, cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        size_t want = bytes + bytes / 8 + 4096;
        RB_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return RB_OK;
}

int rb_staging(rb_ctx *ctx, size_t bytes, void **out)
{
    if (bytes <= rb_ctx::kStageSmall) {
        rb_ctx::StageSlot &s = ctx->stage_ring[ctx->stage_next];
        ctx->stage_next = (ctx->stage_next + 1) % rb_ctx::kStageSlots;
        if (s.in_flight) {
            RB_CUDA(ctx, cudaEventSynchronize(s.ev)); // recorded kStageSlots uploads ago: normally long done
            s.in_flight = false;
        }
        if (!s.p) {
            RB_CUDA(ctx, cudaHostAlloc(&s.p, rb_ctx::kStageSmall, cudaHostAllocDefault));
            s.bytes = rb_ctx::kStageSmall;
            RB_CUDA(ctx, cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
        }
        ctx->staging_cur = &s;
        *out = s.p;
        return RB_OK;
    }
    ctx->staging_cur = nullptr;
    if (ctx->staging_in_flight) {
        RB_CUDA(ctx, cudaEventSynchronize(ctx->staging_ev));
        ctx->staging_in_flight = false;
    }
    if (bytes > ctx->staging_bytes) {
        if (ctx->staging) RB_CUDA(ctx, cudaFreeHost(ctx->staging));
        ctx->staging = nullptr;
        ctx->staging_bytes = 0;
        size_t want = bytes + bytes / 4 + (1u << 20);
        RB_CUDA(ctx, cudaHostAlloc(&ctx->staging, want, cudaHostAllocDefault));
        ctx->staging_bytes = want;
    }
    *out = ctx->staging;
    return RB_OK;
}

int rb_staging_mark(rb_ctx *ctx, cudaStream_t stream)
{
    if (!stream) stream = ctx->stream;
    if (ctx->staging_cur) {
        RB_CUDA(ctx, cudaEventRecord(ctx->staging_cur->ev, stream));
        ctx->staging_cur->in_flight = true;
    } else {
        RB_CUDA(ctx, cudaEventRecord(ctx->staging_ev, stream));
        ctx->staging_in_flight = true;
    }
    return RB_OK;
}

void rb_ctx_retain(rb_ctx *ctx) { ctx->refs.fetch_add(1, std::memory_order_relaxed); }

void rb_ctx_release(rb_ctx *ctx)
{
    if (ctx->refs.fetch_sub(1, std::memory_order_acq_rel) != 1) return;
    if (g_rb_prof.on) {
        static const char *names[6] = {"host build", "alloc + upload", "run (launches)", "layer create/destroy", "record", "composite/mask/filter calls"};
        for (int i = 0; i < 6; i++)
            fprintf(stderr, "[rb profile] %-28s %9.3f ms  %8llu calls\n", names[i], g_rb_prof.t[i] * 1e3, (unsigned long long)g_rb_prof.n[i]);
    }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->px_tables) cudaFree(ctx->px_tables);
    if (ctx->staging) cudaFreeHost(ctx->staging);
    if (ctx->h_flags) cudaFreeHost((void *)ctx->h_flags);
    if (ctx->geo_pinned) cudaFreeHost(ctx->geo_pinned);
    for (auto &st : ctx->geo_streams) if (st) cudaStreamDestroy(st);
    for (auto &ev : ctx->geo_events) if (ev) cudaEventDestroy(ev);
    if (ctx->staging_ev) cudaEventDestroy(ctx->staging_ev);
    for (auto &s : ctx->stage_ring) {
        if (s.p) cudaFreeHost(s.p);
        if (s.ev) cudaEventDestroy(s.ev);
    }
    for (auto &e : ctx->ev_run) if (e) cudaEventDestroy(e);
    if (ctx->ev_band) cudaEventDestroy(ctx->ev_band);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int rb_ctx_create(int device, rb_ctx **out)
{
    if (!out) return RB_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        fprintf(stderr, "[resvg_b200] no usable CUDA device (%s); there is no CPU fallback\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return RB_ERR_CUDA;
    }
    if (device < 0 || device >= count) return RB_ERR_INVALID;
    g_rb_prof.on = getenv("RB_PROFILE") != nullptr;
    rb_ctx *ctx = new rb_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return RB_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return RB_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return RB_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return RB_ERR_CUDA; }
    {
        void *hf = nullptr;
        if (cudaHostAlloc(&hf, 64, cudaHostAllocMapped) != cudaSuccess) { delete ctx; return RB_ERR_CUDA; }
        memset(hf, 0, 64);
        ctx->h_flags = (volatile unsigned int *)hf;
        if (cudaHostGetDevicePointer((void **)&ctx->d_flags, hf, 0) != cudaSuccess) { delete ctx; return RB_ERR_CUDA; }
    }
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaEventCreateWithFlags(&ctx->staging_ev, cudaEventDisableTiming);
    for (auto &e : ctx->ev_run) cudaEventCreate(&e);
    cudaEventCreateWithFlags(&ctx->ev_band, cudaEventDisableTiming);
    // Keep freed layer memory in the stream-ordered pool: isolated groups allocate one layer each
    // (render.rs:108), so layer create/destroy must not hit the driver.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thresh = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    int st = rb_filters_init(ctx);
    if (st != RB_OK) { rb_ctx_destroy(ctx); return st; }
    *out = ctx;
    return RB_OK;
}

// Drops the owner's reference; the context is torn down when the last layer / mask / batch created from it is gone.
extern "C" void rb_ctx_destroy(rb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    rb_ctx_release(ctx);
}

extern "C" int rb_ctx_synchronize(rb_ctx *ctx)
{
    rb_enter(ctx);
    if (!ctx) return RB_ERR_INVALID;
    while (!ctx->dirty.empty()) { // pending immediate draws are work the caller has issued
        int st = rb_layer_flush(ctx->dirty.back());
        if (st != RB_OK) return st;
    }
    RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return rb_check_flags(ctx);
}

extern "C" const char *rb_last_error(rb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void *rb_ctx_stream(rb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int rb_ctx_device(rb_ctx *ctx) { return ctx ? ctx->device : -1; }
extern "C" uint64_t rb_ctx_launch_count(rb_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" uint64_t rb_ctx_h2d_bytes(rb_ctx *ctx) { return ctx ? ctx->h2d_bytes : 0; }

extern "C" int rb_timer_begin(rb_ctx *ctx)
{
    rb_enter(ctx);
    if (!ctx) return RB_ERR_INVALID;
    RB_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return RB_OK;
}

extern "C" int rb_timer_end(rb_ctx *ctx, float *ms)
{
    rb_enter(ctx);
    if (!ctx || !ms) return RB_ERR_INVALID;
    while (!ctx->dirty.empty()) { // the timed region covers the immediate draws issued inside it
        int st = rb_layer_flush(ctx->dirty.back());
        if (st != RB_OK) return st;
    }
    RB_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    RB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    RB_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return rb_check_flags(ctx);
}

extern "C" int rb_host_alloc(size_t bytes, void **out)
{
    if (!out) return RB_ERR_INVALID;
    return cudaMallocHost(out, bytes) == cudaSuccess ? RB_OK : RB_ERR_OOM;
}
extern "C" void rb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ---- layers -------------------------------------------------------------------------------------

extern "C" int rb_layer_create(rb_ctx *ctx, uint32_t w, uint32_t h, rb_layer **out)
{
    rb_prof_scope prof__(RB_T_LAYER);
    rb_enter(ctx);
    if (!ctx || !out) return RB_ERR_INVALID;
    *out = nullptr;
    // tiny-skia Pixmap::new: zero size or a row wider than i32::MAX/4 fails.
    if (w == 0 || h == 0 || w > 0x1fffffffu) return rb_fail(ctx, RB_ERR_INVALID, "invalid layer size");
    size_t bytes = (size_t)w * h * 4;
    void *d = nullptr;
    cudaSetDevice(ctx->device);
    RB_CUDA(ctx, cudaMallocAsync(&d, bytes, ctx->stream));
    RB_CUDA(ctx, cudaMemsetAsync(d, 0, bytes, ctx->stream));
    rb_layer *l = new rb_layer();
    l->ctx = ctx; l->w = w; l->h = h; l->d = (uint8_t *)d;
    rb_ctx_retain(ctx);
    *out = l;
    return RB_OK;
}

extern "C" void rb_layer_destroy(rb_layer *l)
{
    rb_prof_scope prof__(RB_T_LAYER);
    if (!l) return;
    if (l->pending) { // immediate draws nobody looked at: drop them
        rb_batch_destroy(l->pending);
        l->pending = nullptr;
        auto &dv = l->ctx->dirty;
        dv.erase(std::remove(dv.begin(), dv.end(), l), dv.end());
    }
    if (l->dl_pending) cudaEventSynchronize(l->dl_done);
    if (l->dl_ready) cudaEventDestroy(l->dl_ready);
    if (l->dl_done) cudaEventDestroy(l->dl_done);
    cudaFreeAsync(l->d, l->ctx->stream);
    rb_ctx_release(l->ctx);
    delete l;
}

extern "C" uint32_t rb_layer_width(const rb_layer *l) { return l ? l->w : 0; }
extern "C" uint32_t rb_layer_height(const rb_layer *l) { return l ? l->h : 0; }
extern "C" void *rb_layer_device_ptr(rb_layer *l)
{
    if (!l) return nullptr;
    if (l->pending && rb_layer_flush(l) != RB_OK) return nullptr;
    return l->d;
}

extern "C" int rb_layer_upload(rb_layer *l, const uint8_t *host)
{
    rb_enter(l ? l->ctx : nullptr);
    RB_SYNC_LAYER(l);
    if (!l || !host) return RB_ERR_INVALID;
    RB_CUDA(l->ctx, cudaMemcpyAsync(l->d, host, (size_t)l->w * l->h * 4, cudaMemcpyHostToDevice, l->ctx->stream));
    l->ctx->h2d_bytes += (size_t)l->w * l->h * 4;
    return RB_OK;
}

extern "C" int rb_layer_download(rb_layer *l, uint8_t *host)
{
    rb_enter(l ? l->ctx : nullptr);
    RB_SYNC_LAYER(l);
    if (!l || !host) return RB_ERR_INVALID;
    RB_CUDA(l->ctx, cudaMemcpyAsync(host, l->d, (size_t)l->w * l->h * 4, cudaMemcpyDeviceToHost, l->ctx->stream));
    RB_CUDA(l->ctx, cudaStreamSynchronize(l->ctx->stream));
    return rb_check_flags(l->ctx);
}

// Asynchronous download: the copy is ordered after everything enqueued on the layer so far and runs on the context's
// copy stream, so the kernels of the NEXT render (into another layer) overlap it.  `host` should be pinned (rb_host_alloc).
// The layer must not be written again, and `host` not read, before rb_layer_download_end has returned.
extern "C" int rb_layer_download_begin(rb_layer *l, uint8_t *host)
{
    rb_enter(l ? l->ctx : nullptr);
    RB_SYNC_LAYER(l);
    if (!l || !host) return RB_ERR_INVALID;
    rb_ctx *ctx = l->ctx;
    if (l->dl_pending) RB_CUDA(ctx, cudaEventSynchronize(l->dl_done));
    if (!l->dl_ready) {
        RB_CUDA(ctx, cudaEventCreateWithFlags(&l->dl_ready, cudaEventDisableTiming));
        RB_CUDA(ctx, cudaEventCreateWithFlags(&l->dl_done, cudaEventDisableTiming));
    }
    RB_CUDA(ctx, cudaEventRecord(l->dl_ready, ctx->stream));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, l->dl_ready, 0));
    RB_CUDA(ctx, cudaMemcpyAsync(host, l->d, (size_t)l->w * l->h * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
    RB_CUDA(ctx, cudaEventRecord(l->dl_done, ctx->copy_stream));
    l->dl_pending = true;
    return RB_OK;
}

extern "C" int rb_layer_download_end(rb_layer *l)
{
    rb_enter(l ? l->ctx : nullptr);
    RB_SYNC_LAYER(l);
    if (!l) return RB_ERR_INVALID;
    if (l->dl_pending) {
        RB_CUDA(l->ctx, cudaEventSynchronize(l->dl_done));
        l->dl_pending = false;
    }
    return RB_OK;
}

__global__ void __launch_bounds__(256) k_fill_u32(uint32_t *__restrict__ px, size_t n, uint32_t v)
{
    size_t n4 = n >> 2;
    uint4 *p4 = reinterpret_cast<uint4 *>(px);
    uint4 vv = make_uint4(v, v, v, v);
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) p4[i] = vv;
    size_t tail = n & 3;
    if (blockIdx.x == 0 && threadIdx.x < tail) px[(n4 << 2) + threadIdx.x] = v;
}

extern "C" int rb_layer_fill(rb_layer *l, uint8_t r, uint8_t g, uint8_t b, uint8_t a)
{
    rb_enter(l ? l->ctx : nullptr);
    RB_SYNC_LAYER(l);
    if (!l) return RB_ERR_INVALID;
    rb_ctx *ctx = l->ctx;
    size_t n = (size_t)l->w * l->h;
    uint32_t v = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | ((uint32_t)a << 24);
    k_fill_u32<<<rb_grid_1d(ctx, (n + 3) / 4, 256), 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t *>(l->d), n, v);
    RB_LAUNCHED(ctx, "fill");
    return RB_OK;
}

extern "C" int rb_layer_copy(rb_layer *dst, const rb_layer *src)
{
    rb_enter(dst ? dst->ctx : nullptr);
    RB_SYNC_LAYER(dst);
    RB_SYNC_LAYER(src);
    if (!dst || !src || dst->w != src->w || dst->h != src->h) return RB_ERR_INVALID;
    RB_CUDA(dst->ctx, cudaMemcpyAsync(dst->d, src->d, (size_t)src->w * src->h * 4, cudaMemcpyDeviceToDevice,
                                      dst->ctx->stream));
    return RB_OK;
}

// Pixmap::clone_rect (tiny-skia pixmap.rs) as filter/mod.rs:104-108 copy_region uses it: the part of the rectangle that
// lies inside the layer becomes a new layer; RB_ERR_INVALID when they do not intersect (the reference's None).
extern "C" int rb_layer_clone_rect(const rb_layer *src, int32_t x, int32_t y, uint32_t w, uint32_t h, rb_layer **out)
{
    rb_enter(src ? src->ctx : nullptr);
    RB_SYNC_LAYER(src);
    if (!src || !out || w == 0 || h == 0) return RB_ERR_INVALID;
    *out = nullptr;
    const int64_t x0 = std::max<int64_t>(x, 0), y0 = std::max<int64_t>(y, 0);
    const int64_t x1 = std::min<int64_t>((int64_t)x + w, src->w), y1 = std::min<int64_t>((int64_t)y + h, src->h);
    if (x1 <= x0 || y1 <= y0) return RB_ERR_INVALID;
    rb_layer *l = nullptr;
    int st = rb_layer_create(src->ctx, (uint32_t)(x1 - x0), (uint32_t)(y1 - y0), &l);
    if (st != RB_OK) return st;
    cudaError_t e = cudaMemcpy2DAsync(l->d, (size_t)l->w * 4, src->d + ((size_t)y0 * src->w + (size_t)x0) * 4, (size_t)src->w * 4,
                                      (size_t)l->w * 4, l->h, cudaMemcpyDeviceToDevice, src->ctx->stream);
    if (e != cudaSuccess) { rb_layer_destroy(l); return rb_cuda_fail(src->ctx, e, "clone_rect"); }
    *out = l;
    return RB_OK;
}

// Test / tuning hook: clears the RB_PROFILE phase timers (e.g. after warm-up) and prints them on demand.
extern "C" void rb_debug_profile(int reset_only)
{
    if (!reset_only && g_rb_prof.on) {
        static const char *names[6] = {"host build", "alloc + upload", "run (launches)", "layer create/destroy", "record", "composite/mask/filter calls"};
        for (int i = 0; i < 6; i++)
            fprintf(stderr, "[rb profile] %-28s %9.3f ms  %8llu calls\n", names[i], g_rb_prof.t[i] * 1e3, (unsigned long long)g_rb_prof.n[i]);
    }
    for (int i = 0; i < 8; i++) { g_rb_prof.t[i] = 0.0; g_rb_prof.n[i] = 0; }
}

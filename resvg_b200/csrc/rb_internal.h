// rb_internal.h — shared declarations of the resvg_b200 CUDA library (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/resvg_b200.h"

struct rb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; // device-to-host downloads that overlap the next render (rb_layer_download_begin)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_band = nullptr; // a band of the last raster launch is complete (rb_batch_submit_download)
    cudaEvent_t ev_run[3] = {nullptr, nullptr, nullptr}; // last batch run: start, after the pre-pass, after the raster kernel
    std::string err;
    uint64_t launches = 0;
    uint64_t h2d_bytes = 0; // bytes uploaded by batches and layer / mask uploads (rb_ctx_h2d_bytes)
    int sm_count = 148;
    // which opt-in kernel attributes (dynamic shared memory > 48 KB) have been set on THIS context's device: function
    // attributes are per device, so a process-wide flag would leave a second GPU without them
    uint32_t attr_bits = 0;
    // sticky error flags written by kernels straight into mapped pinned host memory ([0]: a tile-row edge list entry was
    // dropped because the host's capacity bound was short); read by rb_check_flags after a stream synchronisation
    volatile unsigned int *h_flags = nullptr;
    unsigned int *d_flags = nullptr;
    // scratch reused by multi-pass filters (grown on demand, freed with the context)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    void *geo_pinned = nullptr;   // 4 KB of pinned memory the geometry kernels' totals are copied back into (geo.cu)
    cudaStream_t geo_streams[4] = {nullptr, nullptr, nullptr, nullptr}; // [0]: the geometry kernels run beside the previous part's raster kernel; [1..3]: their independent chains (geo.cu)
    cudaEvent_t geo_events[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t *px_tables = nullptr; // 3 x 64 KB: demultiply, into_linear_rgb, into_srgb as functions of (alpha, channel)
    // pinned host staging for batch uploads (grown on demand); staging_ev marks the last copy that read it
    void *staging = nullptr;
    size_t staging_bytes = 0;
    cudaEvent_t staging_ev = nullptr;
    bool staging_in_flight = false;
    // Small uploads (the batches of a tree traversal: a few draws per layer) rotate through a ring of pinned slots, so that
    // recording batch k + 1 does not wait for the GPU to have consumed batch k; `staging_cur` is the slot rb_staging handed
    // out last, rb_staging_mark records the copy that reads it.
    struct StageSlot { void *p = nullptr; size_t bytes = 0; cudaEvent_t ev = nullptr; bool in_flight = false; };
    static constexpr int kStageSlots = 16;
    static constexpr size_t kStageSmall = 1u << 20;
    StageSlot stage_ring[kStageSlots];
    int stage_next = 0;
    StageSlot *staging_cur = nullptr; // nullptr: the large block
    // owner + every live layer / mask / batch: the context outlives them whatever the destruction order
    std::atomic<int> refs{1};
    std::vector<struct rb_layer *> dirty; // layers holding pending immediate draws (rb_fill_path), flushed by rb_ctx_synchronize
};

// Every entry point that allocates or launches makes the context's device current first: the host (torch, a second
// context) may have switched devices since the last call.
static inline void rb_enter(const rb_ctx *ctx)
{
    if (ctx) cudaSetDevice(ctx->device);
}
enum { RB_ATTR_BOX = 1, RB_ATTR_MORPH = 2, RB_ATTR_TURB = 4, RB_ATTR_WIDE = 8, RB_ATTR_PXTABLE = 16, RB_ATTR_IIR = 32, RB_ATTR_GEO = 64 };

void rb_ctx_retain(rb_ctx *ctx);
void rb_ctx_release(rb_ctx *ctx);
// Pinned staging block of at least `bytes`, safe to overwrite (waits for the previous upload out of it).
int rb_staging(rb_ctx *ctx, size_t bytes, void **out);
// After enqueuing the copy that reads the block rb_staging returned last: records when it may be overwritten.
int rb_staging_mark(rb_ctx *ctx, cudaStream_t stream = nullptr); // stream: the one the copy was enqueued on (default: the context's)

struct rb_layer {
    rb_ctx *ctx;
    uint32_t w, h;
    uint8_t *d; // w*h*4 bytes, premultiplied RGBA8
    cudaEvent_t dl_ready = nullptr, dl_done = nullptr; // rb_layer_download_begin / _end
    bool dl_pending = false;
    // Draws issued with the immediate calls (rb_fill_path) are collected here and executed as ONE batch the next time
    // anything else looks at or changes the layer (rb_layer_flush): consecutive fill_path calls of a traversal cost one
    // tile-kernel launch instead of one each.
    struct rb_batch *pending = nullptr;
    uint32_t pending_n = 0;
    // Canvas strips (rb_render_strip): the layer is a window of a larger pixmap placed at (vp_x, vp_y) relative to it; immediate
    // draws are recorded with that viewport (rb_batch_set_viewport), i.e. built against the whole pixmap.  vp_w == 0: none.
    int32_t vp_x = 0, vp_y = 0, vp_w = 0, vp_h = 0;
};

// Executes the layer's pending immediate draws, if any.  Every entry point that reads or writes a layer calls it first.
int rb_layer_flush(rb_layer *l);
#define RB_SYNC_LAYER(l)                                                        \
    do {                                                                        \
        rb_layer *l__ = const_cast<rb_layer *>(l);                              \
        if (l__ && l__->pending) {                                              \
            int st__ = rb_layer_flush(l__);                                     \
            if (st__ != RB_OK) return st__;                                     \
        }                                                                       \
    } while (0)

struct rb_mask {
    rb_ctx *ctx;
    uint32_t w, h;
    uint8_t *d; // w*h bytes
    int32_t vp_x = 0, vp_y = 0, vp_w = 0, vp_h = 0; // a window of a larger mask, like rb_layer's (canvas strips)
};

// Host-side phase timers (RB_PROFILE=1 prints them when a context is destroyed): where a traversal's wall time goes.
struct rb_host_prof { double t[8]; uint64_t n[8]; bool on; };
extern rb_host_prof g_rb_prof;
enum { RB_T_BUILD = 0, RB_T_UPLOAD = 1, RB_T_RUN = 2, RB_T_LAYER = 3, RB_T_RECORD = 4, RB_T_COMPOSITE = 5 };
struct rb_prof_scope {
    int k; double t0;
    explicit rb_prof_scope(int kind);
    ~rb_prof_scope();
};
int rb_fail(rb_ctx *ctx, int code, const char *what);
// After a synchronisation: RB_ERR_CUDA (and rb_last_error) if a kernel raised a sticky flag since the last check.
int rb_check_flags(rb_ctx *ctx);
int rb_cuda_fail(rb_ctx *ctx, cudaError_t e, const char *what);
// Ensures ctx->scratch holds at least `bytes`; returns RB_OK or an error status.
int rb_scratch(rb_ctx *ctx, size_t bytes, void **out);

#define RB_CUDA(ctx, call)                                              \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return rb_cuda_fail((ctx), e__, #call); \
    } while (0)

// Call after every kernel launch: counts the launch and surfaces launch-configuration errors.
#define RB_LAUNCHED(ctx, name)                                           \
    do {                                                                 \
        (ctx)->launches++;                                               \
        cudaError_t e__ = cudaGetLastError();                            \
        if (e__ != cudaSuccess) return rb_cuda_fail((ctx), e__, (name)); \
    } while (0)

static inline int rb_grid_1d(const rb_ctx *ctx, size_t work_items, int block, int max_waves = 8)
{
    size_t blocks = (work_items + block - 1) / block;
    size_t cap = (size_t)ctx->sm_count * (2048 / block) * max_waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

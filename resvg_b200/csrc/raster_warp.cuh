// raster_warp.cuh — the fill-path kernel proper (included by raster.cu after the pipeline code).
//
// One WARP owns one 32x8-pixel tile of the layer for a whole batch: its 256 destination pixels live in registers
// (8 per lane), it walks the tile's draw list in painter's order, and no block-level barrier exists anywhere.
// Per (draw, tile):
//   scatter   lanes = edges of the draw's list for this tile row (built on the device by k_row_lists); every
//             crossing x(y) is added into a packed per-position counter (one 32-bit word per sub-sample position,
//             byte s = net crossings on sub-row s of the pixel row) and marked in a per-sub-row bit mask;
//   scan      lanes = the 32 sub-scanlines; a lane visits only the marked positions of its sub-row (ffs), keeps the
//             running winding in a full int, records the positions where inside/outside toggles, and a 128-bit
//             prefix-xor turns those into the inside mask of the sub-row — work proportional to crossings, not to
//             sub-samples;
//   coverage  lanes = 8-pixel groups; nibble popcounts of the four sub-row masks give tiny-skia's coverage
//             (16 per sub-sample, 64/64/64/63 for full pixels, abutting-span exception from the sub-row-3 marks);
//   blend     the reference's pipeline per covered pixel (pipeline code in raster.cu).
//
// Only per-position NET crossings are packed into bytes, so the limit is |net crossings at one sub-sample| <= 127;
// the host routes draws that could exceed it to k_raster_tiles_wide.
#pragma once

constexpr int WT_W = 32;           // warp-tile width  (pixels)
constexpr int WT_H = 8;            // warp-tile height (pixels)
constexpr int WT_POS = WT_W * 4;   // sub-sample positions per row
constexpr int WT_SUB = WT_H * 4;   // sub-scanlines per tile
// Warps per CTA.  The warps of a CTA never cooperate, so a CTA is one warp: with 4 warps per CTA a slot stayed occupied
// until the slowest of the four tiles was finished (measured on the C2 scene: 21.9 ms with 4 warps x 6 CTAs, 18.5 ms with
// 1 warp x up to 25 CTAs per SM at the same 80 registers).
constexpr int WT_WARPS = 1;

struct WarpTileSmem {
    int wsum[WT_H * WT_POS];        // packed net crossings: [pixel row][position], byte s = sub-row s
    uint32_t tmask[WT_SUB][4];      // positions with at least one crossing, per sub-row
    uint32_t dmask[WT_H][2][4];     // sub-row 3 of every pixel row: positions with downward [0] / upward [1] crossings
    uint32_t inside[WT_SUB][4];     // scan result: inside mask per sub-row
    uint32_t flip[WT_H][4];         // sub-row 3: positions where the winding passes through zero between two non-zero values
    int bd[WT_SUB + 4];             // difference array of the backdrop: edges wholly left of the tile add +-1 over their rows
    uint32_t px[WT_W * WT_H];       // blend_tile_gradient: the tile's pixels while the covered ones are dealt out to the lanes
};

// ---- rare path ---------------------------------------------------------------------------------------------------
// Crossings of both directions share one sub-sample position strictly inside a fully covered pixel on the 4th
// sub-row, so whether the span breaks there depends on the scanline walker's list order (tiny-skia scan/path.rs
// walk_edges).  One lane replays that position from the tile row's edge list.  `list` entries carry their own slot
// (index in the draw's edge array) in meta >> 4; `edges` is the draw's edge array (meta >> 4 = previous segment).
struct WalkEdge { DevEdge e; uint32_t slot; uint32_t prev_meta; };

// insert_new_edges on the scanline where a (non-continuation) edge joins the walker's list: the first new edge of
// that scanline (smallest x among the edges starting there) goes after equal-x actives, every later one before them.
// Those edges need not reach this tile row, so the whole draw is scanned — only when two crossings really tie.
__device__ __noinline__ bool inserted_before(const DevEdge *__restrict__ edges, uint32_t n_edges, const DevEdge &E)
{
    const uint32_t fy = E.ypack & 0xffffu;
    for (uint32_t k = 0; k < n_edges; k++) {
        const DevEdge O = edges[k];
        if ((O.ypack & 0xffffu) != fy || (O.meta & 2u) || (O.ypack >> 16) < fy) continue;
        if (O.x < E.x) return true;
    }
    return false;
}

__device__ __forceinline__ bool sorted_before(const DevEdge &A, uint32_t sa, const DevEdge &B, uint32_t sb)
{
    // order of the walker's initial sort: (first_y, x), ties in builder order
    const int fa = (int)(A.ypack & 0xffffu), fb = (int)(B.ypack & 0xffffu);
    if (fa != fb) return fa < fb;
    if (A.x != B.x) return A.x < B.x;
    return sa < sb;
}

// kind: 0 = was active on y-1 ("survivor", keyed by its previous x), 1 = new edge placed after equal-x survivors,
// 2 = new edge placed before them (insert_new_edges stops at the first active edge with x >= its x for every new
// edge but the first of its scanline batch).
__device__ bool walker_less_l(const DevEdge *__restrict__ edges, uint32_t n_edges, const WalkEdge &A, const WalkEdge &B, int y)
{
    for (int depth = 0; depth < 2; depth++) {
        const int xa = edge_x_at(A.e, y), xb = edge_x_at(B.e, y);
        if (xa != xb) return xa < xb;
        const int fya = (int)(A.e.ypack & 0xffffu), fyb = (int)(B.e.ypack & 0xffffu);
        int ka, kb, pa = 0, pb = 0;
        if (y > fya) { ka = 0; pa = (int)((uint32_t)xa - (uint32_t)A.e.dx); }
        else if (A.prev_meta & 2u) { const DevEdge P = edges[A.prev_meta >> 4]; ka = 0; pa = edge_x_at(P, (int)(P.ypack >> 16)); }
        else ka = inserted_before(edges, n_edges, A.e) ? 2 : 1;
        if (y > fyb) { kb = 0; pb = (int)((uint32_t)xb - (uint32_t)B.e.dx); }
        else if (B.prev_meta & 2u) { const DevEdge P = edges[B.prev_meta >> 4]; kb = 0; pb = edge_x_at(P, (int)(P.ypack >> 16)); }
        else kb = inserted_before(edges, n_edges, B.e) ? 2 : 1;
        if (ka == 0 && kb == 0) {
            if (pa != pb) return pa < pb;
            // coincident lines: their order was fixed on the scanline where the later one joined the list
            if (y > fya && y > fyb) { y = max(fya, fyb); continue; }
            return sorted_before(A.e, A.slot, B.e, B.slot);
        }
        if (ka == 0) return kb == 1;
        if (kb == 0) return ka == 2;
        return sorted_before(A.e, A.slot, B.e, B.slot);
    }
    return sorted_before(A.e, A.slot, B.e, B.slot);
}

__device__ __noinline__ bool exact_span_break_list(const DevEdge *__restrict__ all_edges, const DevDraw *__restrict__ draws, uint32_t draw,
                                                   const DevEdge *__restrict__ list, uint32_t n, int y, int target_r, int lo_r)
{
    const DevEdge *edges = all_edges + draws[draw].edge_off;
    const uint32_t n_edges = draws[draw].edge_cnt;
    // winding before the position (crossings left of it; positions are clamped to lo_r on the left) and the
    // candidates at it
    WalkEdge c[12];
    int cnt = 0, w = 0;
    for (uint32_t i = 0; i < n; i++) {
        const DevEdge E = list[i];
        const int fy = (int)(E.ypack & 0xffffu), ly = (int)(E.ypack >> 16);
        if (fy > y || ly < y) continue;
        const int r = max((int)((uint32_t)edge_x_at(E, y) + 0x8000u) >> 16, lo_r);
        if (r < target_r) { w += edge_winding(E.meta); continue; }
        if (r != target_r) continue;
        if (cnt == 12) return true;
        WalkEdge we;
        we.e = E;
        we.slot = E.meta >> 4;
        we.prev_meta = edges[we.slot].meta;
        int j = cnt++;
        while (j > 0 && walker_less_l(edges, n_edges, we, c[j - 1], y)) { c[j] = c[j - 1]; j--; }
        c[j] = we;
    }
    for (int i = 0; i < cnt; i++) {
        w += edge_winding(c[i].e.meta);
        if (w == 0) return true;
    }
    return false;
}

// ---- per-draw tile-row edge lists ----------------------------------------------------------------------------------
// One CTA per draw.  Tile row r of a draw = layer pixel rows [8 (r0 + r), 8 (r0 + r) + 8); an edge is copied into the
// list of every tile row its sub-scanline range touches (meta: bit 0 upward, bit 1 continuation, bits 4.. own slot).
struct DrawBox { uint32_t rows, row_base; }; // r0 | r1 << 16 (inclusive warp-tile rows); where its row_cols entries start
constexpr int RL_THREADS = 128;
constexpr int RL_MAX_ROWS = 1026;
constexpr int RL_LONG = 32, RL_LONG_Q = 32; // edges crossing >= RL_LONG tile rows are walked by the whole CTA (up to RL_LONG_Q per draw)

__device__ __forceinline__ int draw_row_of(const DevDraw &D, int suby)
{
    const int r = (((suby >> D.shift) + D.oy) >> 3) - D.r0;
    return min(max(r, 0), (int)D.n_rows - 1);
}

// Item mode: the draw's edge array is first materialised from the uploaded items — line edges are copied to their
// slots, every recorded curve is forward-differenced by one thread into the line edges tiny-skia's update() calls
// would produce (edge_math.h, shared with the host builder), unused slots are marked empty (first_y > last_y).
__global__ void __launch_bounds__(RL_THREADS)
k_row_lists(const DevDraw *__restrict__ draws, const DevEdge *__restrict__ lines, const rbh::CurveRec *__restrict__ curves,
            DevEdge *__restrict__ edges, uint32_t *__restrict__ row_off, DevEdge *__restrict__ row_edges, uint32_t *__restrict__ row_cols,
            int items, unsigned int *__restrict__ overflow, int wtiles_x, DrawBox *__restrict__ boxes, uint32_t *__restrict__ row_cnt,
            uint32_t *__restrict__ tile_cnt)
{
    __shared__ uint32_t cnt[RL_MAX_ROWS + 2];
    __shared__ int xlo[RL_MAX_ROWS + 2], xhi[RL_MAX_ROWS + 2]; // extent of the crossings of each tile row, in pixels
    __shared__ uint32_t warp_tot[RL_THREADS / 32];
    __shared__ uint32_t long_q[RL_LONG_Q], n_long;
    const DevDraw D = draws[blockIdx.x];
    const int tid = threadIdx.x, nr = (int)D.n_rows;
    if (nr == 0) {
        // an empty draw (the geometry kernels keep one DevDraw per task: culled paths, strokes that come out empty, the
        // reserved draws of a dashed hairline beyond its dashes): no rows, no lists, never binned
        if (tid == 0) {
            if (boxes) boxes[blockIdx.x] = DrawBox{D.r0 | ((D.r0 + D.n_rows - 1) << 16), D.row_base};
            row_off[D.row_base] = 0;
        }
        return;
    }
    DevEdge *E0 = edges + D.edge_off;
    for (int i = tid; i <= nr; i += RL_THREADS) { cnt[i] = 0; xlo[i] = INT_MAX; xhi[i] = INT_MIN; }
    if (tid == 0 && boxes) boxes[blockIdx.x] = DrawBox{D.r0 | ((D.r0 + D.n_rows - 1) << 16), D.row_base};
    // Binning counts (how many draws touch each warp-tile row / each warp tile) are taken here, by the threads that have
    // just computed a row's column extent: a draw spanning the whole layer has hundreds of rows, which a thread-per-draw
    // counting kernel walked serially (124 us per launch for full-layer draws on a 4096 x 4096 layer).
    auto count_row = [&](int i, uint32_t cols) {
        const uint32_t c0 = cols & 0xffffu, c1 = cols >> 16;
        if (c0 > c1 || !row_cnt) return; // direct mode (<= 32 draws): no bin tables
        const uint32_t r = D.r0 + (uint32_t)i;
        atomicAdd(&row_cnt[r], 1u);
        uint32_t *t = tile_cnt + (size_t)r * wtiles_x;
        for (uint32_t c = c0; c <= c1; c++) atomicAdd(&t[c], 1u);
    };
    if (D.rule == 2) {
        // A hairline stroke: `lines` holds its ordered blits (x | y << 16 in layer pixels, coverage, rank inside its
        // warp-tile cell, cell index).  The draw's list table has one entry per CELL of its bounding box (row-major,
        // curve_cnt cells per row) instead of one per tile row; blits are copied to cell start + rank, i.e. in the
        // order the walker produced them.
        const uint32_t n_cells = D.n_rows * D.curve_cnt;
        uint32_t *cell_off = row_off + D.row_base;
        for (uint32_t i = tid; i <= n_cells; i += RL_THREADS) cell_off[i] = 0;
        __syncthreads();
        const DevEdge *B0 = lines + D.line_off;
        for (uint32_t i = tid; i < D.line_cnt; i += RL_THREADS) {
            const DevEdge B = B0[i];
            const int r = (int)(((uint32_t)B.x >> 16) >> 3) - (int)D.r0, x = (int)((uint32_t)B.x & 0xffffu);
            atomicAdd(&cell_off[B.meta], 1u);
            atomicMin(&xlo[r], x);
            atomicMax(&xhi[r], x);
        }
        __syncthreads();
        {
            // exclusive scan of the cell counts (global memory; contiguous chunks per thread + warp shuffles)
            const uint32_t per = (n_cells + RL_THREADS - 1) / RL_THREADS;
            const uint32_t lo = min((uint32_t)tid * per, n_cells), hi = min(lo + per, n_cells);
            uint32_t sum = 0;
            for (uint32_t i = lo; i < hi; i++) sum += cell_off[i];
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if ((tid & 31) >= d) incl += v;
            }
            if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
            __syncthreads();
            uint32_t base = incl - sum;
            for (int w = 0; w < (tid >> 5); w++) base += warp_tot[w];
            for (uint32_t i = lo; i < hi; i++) { const uint32_t c = cell_off[i]; cell_off[i] = base; base += c; }
            if (tid == RL_THREADS - 1) cell_off[n_cells] = base;
        }
        __syncthreads();
        for (int i = tid; i < nr; i += RL_THREADS) {
            const uint32_t cols = xlo[i] <= xhi[i] ? ((uint32_t)(xlo[i] / WT_W) | ((uint32_t)(xhi[i] / WT_W) << 16)) : 1u;
            row_cols[D.row_base + i] = cols;
            count_row(i, cols);
        }
        DevEdge *out = row_edges + D.list_off;
        for (uint32_t i = tid; i < D.line_cnt; i += RL_THREADS) {
            const DevEdge B = B0[i];
            const uint32_t at = cell_off[B.meta] + B.ypack;
            if (at < D.list_cap) out[at] = B;
            else *(volatile unsigned int *)overflow = 1u; // host-mapped sticky flag, checked at the next sync point
        }
        return;
    }
    if (items) {
        for (uint32_t i = tid; i < D.line_cnt; i += RL_THREADS) {
            DevEdge L = lines[D.line_off + i];
            const uint32_t slot = L.meta >> 4;
            L.meta &= 1u;
            E0[slot] = L;
        }
        for (uint32_t i = tid; i < D.curve_cnt; i += RL_THREADS) {
            const rbh::CurveRec C = curves[D.curve_off + i];
            const int sh = (int)((C.info >> 4) & 0xfu);
            const uint32_t up = (C.info >> 8) & 1u, first = C.item, n_slots = 1u << sh;
            uint32_t k = 0;
            auto emit = [&](const rbe::RawEdge &r) {
                DevEdge e;
                e.x = r.x;
                e.dx = r.dx;
                e.ypack = ((uint32_t)r.first_y & 0xffffu) | ((uint32_t)r.last_y << 16);
                e.meta = up | (k ? (2u | ((first + k - 1) << 4)) : 0u);
                if (k < n_slots) E0[first + k] = e;
                k++;
            };
            if (C.info & 1u) rbe::cubic_expand(C.p[0], C.p[1], C.p[2], C.p[3], C.p[4], C.p[5], C.p[6], C.p[7], sh, emit);
            else rbe::quad_expand(C.p[0], C.p[1], C.p[2], C.p[3], C.p[4], C.p[5], sh, emit);
            DevEdge none;
            none.x = 0; none.dx = 0; none.ypack = 0xffffu; none.meta = 0;
            for (; k < n_slots; k++) E0[first + k] = none;
        }
    }
    __syncthreads();
    {
        // Per tile row: how many edges touch it, and between which pixel columns they cross it.  Outside that extent the
        // winding is zero (contours are closed), so only the warp tiles inside it are paired with this draw.
        const int sub_lo = D.sy << D.shift, sub_hi = ((D.sy + D.sh) << D.shift) - 1; // sub-scanlines the blitter may touch
        auto touch = [&](const DevEdge &E, int fy, int ly, int r) {
            atomicAdd(&cnt[r], 1u);
            // sub-scanlines of tile row r: layer pixel rows [8 (r0 + r), +8) in the draw's units
            const int top = max(max((((int)(D.r0 + r) << 3) - D.oy) << D.shift, sub_lo), fy);
            const int bot = min(min((((((int)(D.r0 + r) + 1) << 3) - D.oy) << D.shift) - 1, sub_hi), ly);
            if (top > bot) return;
            const int xa = (int)((uint32_t)E.x + (uint32_t)(top - fy) * (uint32_t)E.dx), xb = (int)((uint32_t)E.x + (uint32_t)(bot - fy) * (uint32_t)E.dx);
            const int pa = (((int)((uint32_t)xa + 0x8000u) >> 16) >> D.shift), pb = (((int)((uint32_t)xb + 0x8000u) >> 16) >> D.shift);
            atomicMin(&xlo[r], min(pa, pb));
            atomicMax(&xhi[r], max(pa, pb));
        };
        // An edge that crosses many tile rows (the sides of a layer-sized rectangle: 512 rows on a 4096 px layer) is
        // queued and walked by the whole CTA, rows strided over the threads, instead of by the one thread that owns it.
        if (tid == 0) n_long = 0;
        __syncthreads();
        for (uint32_t e = tid; e < D.edge_cnt; e += RL_THREADS) {
            const DevEdge E = E0[e];
            const int fy = (int)(E.ypack & 0xffffu), ly = (int)(E.ypack >> 16);
            if (fy > ly) continue; // empty slot
            const int ra = draw_row_of(D, fy), rb = draw_row_of(D, ly);
            if (rb - ra >= RL_LONG) {
                const uint32_t q = atomicAdd(&n_long, 1u);
                if (q < RL_LONG_Q) { long_q[q] = e; continue; }
            }
            for (int r = ra; r <= rb; r++) touch(E, fy, ly, r);
        }
        __syncthreads();
        for (uint32_t q = 0; q < min(n_long, (uint32_t)RL_LONG_Q); q++) {
            const DevEdge E = E0[long_q[q]];
            const int fy = (int)(E.ypack & 0xffffu), ly = (int)(E.ypack >> 16);
            const int ra = draw_row_of(D, fy), rb = draw_row_of(D, ly);
            for (int r = ra + tid; r <= rb; r += RL_THREADS) touch(E, fy, ly, r);
        }
    }
    __syncthreads();
    for (int i = tid; i < nr; i += RL_THREADS) {
        // warp-tile columns [c0, c1] of row i that can receive coverage, clipped to the blitter rectangle; c0 > c1: none
        uint32_t cols = 1u;
        if (xlo[i] <= xhi[i]) {
            const int px0 = max(xlo[i], D.sx), px1 = min(xhi[i], D.sx + D.sw - 1);
            if (px0 <= px1) cols = (uint32_t)((D.ox + px0) / WT_W) | ((uint32_t)((D.ox + px1) / WT_W) << 16);
        }
        row_cols[D.row_base + i] = cols;
        count_row(i, cols);
    }
    // exclusive scan of cnt[0..nr) -> cnt; contiguous chunks per thread + warp shuffles
    {
        const int per = (nr + RL_THREADS - 1) / RL_THREADS;
        const int lo = tid * per, hi = min(lo + per, nr);
        uint32_t s = 0;
        for (int i = lo; i < hi; i++) s += cnt[i];
        uint32_t incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31) >= d) incl += v;
        }
        if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
        __syncthreads();
        uint32_t base = incl - s;
        for (int w = 0; w < (tid >> 5); w++) base += warp_tot[w];
        for (int i = lo; i < hi; i++) { uint32_t c = cnt[i]; cnt[i] = base; base += c; }
        if (hi == nr && lo < nr) cnt[nr] = base;
        if (nr == 0 && tid == 0) cnt[0] = 0;
    }
    __syncthreads();
    for (int i = tid; i <= nr; i += RL_THREADS) row_off[D.row_base + i] = cnt[i];
    __syncthreads();
    DevEdge *out = row_edges + D.list_off;
    auto put = [&](const DevEdge &E, int r) {
        const uint32_t at = atomicAdd(&cnt[r], 1u);
        if (at < D.list_cap) out[at] = E;
        else *(volatile unsigned int *)overflow = 1u; // cannot happen: the host's bound covers every segment
    };
    for (uint32_t e = tid; e < D.edge_cnt; e += RL_THREADS) {
        DevEdge E = E0[e];
        const int fy = (int)(E.ypack & 0xffffu), ly = (int)(E.ypack >> 16);
        if (fy > ly) continue;
        E.meta = (E.meta & 3u) | (e << 4);
        const int ra = draw_row_of(D, fy), rb = draw_row_of(D, ly);
        if (rb - ra >= RL_LONG) {
            bool queued = false;
            for (uint32_t q = 0; q < min(n_long, (uint32_t)RL_LONG_Q); q++) queued = queued || long_q[q] == e;
            if (queued) continue; // walked by the whole CTA below
        }
        for (int r = ra; r <= rb; r++) put(E, r);
    }
    for (uint32_t q = 0; q < min(n_long, (uint32_t)RL_LONG_Q); q++) {
        const uint32_t e = long_q[q];
        DevEdge E = E0[e];
        const int fy = (int)(E.ypack & 0xffffu), ly = (int)(E.ypack >> 16);
        E.meta = (E.meta & 3u) | (e << 4);
        const int ra = draw_row_of(D, fy), rb = draw_row_of(D, ly);
        for (int r = ra + tid; r <= rb; r += RL_THREADS) put(E, r);
    }
}

// ---- binning draws into warp tiles (painter's order kept without sorting) ------------------------------------------

// In-place exclusive scan of a[0..n), a[n] = total, in three small launches: every CTA scans 4096 consecutive
// elements (coalesced 16-byte loads, 4 per thread) and publishes its sum; one CTA scans the sums; every CTA adds its
// offset.  (The tile table has 262 144 entries for an 8192 x 8192 layer: a single CTA took 0.45 ms.)
constexpr int SCAN_THREADS = 1024, SCAN_PER_CTA = SCAN_THREADS * 4;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_tot, uint32_t *total)
{
    const uint32_t tid = threadIdx.x;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
        if ((tid & 31u) >= (uint32_t)d) incl += u;
    }
    if ((tid & 31u) == 31u) warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        uint32_t w = warp_tot[tid], t = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t u = __shfl_up_sync(0xffffffffu, t, d);
            if (tid >= (uint32_t)d) t += u;
        }
        warp_tot[tid] = t - w;
        if (tid == 31) *total = t;
    }
    __syncthreads();
    return incl - v + warp_tot[tid >> 5];
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_local(uint32_t *__restrict__ a, uint32_t n, uint32_t *__restrict__ block_sums)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t total;
    const uint32_t base = blockIdx.x * SCAN_PER_CTA + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = base + k < n ? a[base + k] : 0u;
    const uint32_t sum = v[0] + v[1] + v[2] + v[3];
    uint32_t off = block_exclusive_scan(sum, warp_tot, &total);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < n) a[base + k] = off;
        off += v[k];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// scans the block sums (at most SCAN_PER_CTA of them, i.e. n <= 16 M entries) and writes the grand total to a[n]
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_sums(uint32_t *__restrict__ block_sums, uint32_t n_blocks, uint32_t *__restrict__ a, uint32_t n)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t total;
    const uint32_t base = threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = base + k < n_blocks ? block_sums[base + k] : 0u;
    uint32_t off = block_exclusive_scan(v[0] + v[1] + v[2] + v[3], warp_tot, &total);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < n_blocks) block_sums[base + k] = off;
        off += v[k];
    }
    if (threadIdx.x == 0) a[n] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_add(uint32_t *__restrict__ a, uint32_t n, const uint32_t *__restrict__ block_sums)
{
    const uint32_t add = block_sums[blockIdx.x];
    const uint32_t base = blockIdx.x * SCAN_PER_CTA + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) if (base + k < n) a[base + k] += add;
}

struct RowEnt { uint32_t draw, cols; };

// One CTA per warp-tile row: the draws touching the row, in draw order.
__global__ void __launch_bounds__(256)
k_bin_rows(const DrawBox *__restrict__ boxes, uint32_t n_draws, const uint32_t *__restrict__ row_cols, const uint32_t *__restrict__ row_off,
           RowEnt *__restrict__ row_draws)
{
    __shared__ uint32_t warp_cnt[8];
    const uint32_t R = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint32_t out = row_off[R];
    if (row_off[R + 1] == out) return;
    for (uint32_t base = 0; base < n_draws; base += 256) {
        const uint32_t d = base + tid;
        uint32_t cols = 1u;
        bool ok = false;
        if (d < n_draws) {
            const DrawBox b = boxes[d];
            if ((b.rows & 0xffffu) <= R && R <= (b.rows >> 16)) {
                cols = row_cols[b.row_base + (R - (b.rows & 0xffffu))];
                ok = (cols & 0xffffu) <= (cols >> 16);
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_cnt[wid] = __popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (uint32_t w = 0; w < 8; w++) {
            const uint32_t c = warp_cnt[w];
            if (w < wid) before += c;
            total += c;
        }
        if (ok) row_draws[out + before + __popc(m & ((1u << lane) - 1u))] = RowEnt{d, cols};
        out += total;
        __syncthreads();
    }
}

// One warp per warp tile: filters its row's draw list by column range, keeping the order.
__global__ void __launch_bounds__(256)
k_bin_tiles(const RowEnt *__restrict__ row_draws, const uint32_t *__restrict__ row_off, const uint32_t *__restrict__ tile_off,
            int wtiles_x, uint32_t n_wtiles, uint32_t *__restrict__ tile_pairs)
{
    const uint32_t t = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (t >= n_wtiles) return;
    uint32_t out = tile_off[t];
    if (tile_off[t + 1] == out) return;
    const uint32_t R = t / (uint32_t)wtiles_x, Cc = t % (uint32_t)wtiles_x;
    const uint32_t lo = row_off[R], hi = row_off[R + 1];
    for (uint32_t i = lo; i < hi; i += 32) {
        bool ok = false;
        RowEnt e = RowEnt{0, 0};
        if (i + lane < hi) {
            e = row_draws[i + lane];
            ok = (e.cols & 0xffffu) <= Cc && Cc <= (e.cols >> 16);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (ok) tile_pairs[out + __popc(m & ((1u << lane) - 1u))] = e.draw;
        out += __popc(m);
    }
}

// ---- gradients: the covered pixels of a pair, dealt out evenly -------------------------------------------------------------
// A gradient costs ~200 instructions per pixel, and a lane that walks its own eight pixels keeps the whole warp busy for as
// many steps as the fullest lane has covered pixels (7 of 8 on the 100 000-path scene, with 40 % of the lanes covered).  Here
// the warp stages the tile's pixels and coverages in shared memory, lists the covered pixels, and every lane takes every
// 32nd entry of the list: ceil(covered / 32) steps.  Source / SourceOver in the u16 or f32 pipeline; what the pixels share
// (the paint's fields) is read once per call.  Called by the whole warp.  (Choosing per pair between this and a
// lane-by-lane loop — better for fully covered tiles — was measured: two hot gradient functions fall out of the instruction
// cache, 18.6 ms instead of 14.6.)
__device__ __noinline__ void blend_tile_gradient(WarpTileSmem &S, const DevPaint &P, const DevStop *__restrict__ stops, Px8 &px, uint32_t c0,
                                                 uint32_t c1, uint32_t dec, int tlx, int tly)
{
    const int lane = threadIdx.x & 31;
    // coverage bytes: c = min(16 * count - dec, 255), four pixels per word
    auto cov4 = [](uint32_t cnt, uint32_t dnib) { // dnib: bit 4k = pixel k counts 63 on its last sub-row
        const uint32_t full = (cnt >> 4) & 0x01010101u; // count == 16
        uint32_t d = dnib & 0x1111u;                    // bits 0, 4, 8, 12 -> 0, 8, 16, 24
        d = (d | (d << 8)) & 0x00ff00ffu;
        d = (d | (d << 4)) & 0x01010101u;
        return ((((cnt & 0x0f0f0f0fu) << 4) - (d & ~full)) | (full * 255u));
    };
    const uint32_t cv0 = cov4(c0, dec), cv1 = cov4(c1, dec >> 16);
    uint8_t *list = reinterpret_cast<uint8_t *>(S.inside);       // [256] covered pixels, p = lane * 8 + q
    uint32_t *cvw = reinterpret_cast<uint32_t *>(S.inside) + 64; // [64] coverage bytes, indexed like the pixels
    uint32_t m8 = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if ((cv0 >> (8 * q)) & 0xffu) m8 |= 1u << q;
        if ((cv1 >> (8 * q)) & 0xffu) m8 |= 16u << q;
    }
    const int n = __popc(m8);
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += u;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    *reinterpret_cast<uint4 *>(&S.px[lane * 8]) = make_uint4(px.v[0], px.v[1], px.v[2], px.v[3]);
    *reinterpret_cast<uint4 *>(&S.px[lane * 8 + 4]) = make_uint4(px.v[4], px.v[5], px.v[6], px.v[7]);
    *reinterpret_cast<uint2 *>(&cvw[lane * 2]) = make_uint2(cv0, cv1);
    {
        int at = incl - n;
        uint32_t m = m8;
        while (m) {
            const int q = __ffs(m) - 1;
            m &= m - 1;
            list[at++] = (uint8_t)(lane * 8 + q);
        }
    }
    __syncwarp();
    // everything the pixels share is read here, once
    const bool lowp = P.lowp != 0, src_over = P.blend == 3, memset_ok = P.has_memset != 0, has_ts = P.has_ts != 0;
    const bool premul_after = P.premul_after != 0;
    const uint32_t memset_color = P.memset_color;
    const GradGeom G = grad_geom(P);
    const float *__restrict__ t0s = P.t0s;
    const DevStop *__restrict__ st = stops + P.stop_off;
    const int len = P.two_stop ? 1 : P.len;
    const float ts0 = P.ts[0], ts1 = P.ts[1], ts2 = P.ts[2], ts3 = P.ts[3], ts4 = P.ts[4], ts5 = P.ts[5];
    const uint8_t *cvb = reinterpret_cast<const uint8_t *>(cvw);
#pragma unroll 1
    for (int i = lane; i < total; i += 32) {
        const uint32_t p = list[i];
        const uint32_t c = cvb[p];
        if (c == 255 && memset_ok) { S.px[p] = memset_color; continue; }
        const uint32_t d = S.px[p];
        float x = (float)(tlx + (int)(((p >> 3) & 3u) * 8u + (p & 7u))) + 0.5f, y = (float)(tly + (int)(p >> 5)) + 0.5f;
        if (has_ts) {
            const float nx = mad(x, ts0, mad(y, ts2, ts4)), ny = mad(x, ts1, mad(y, ts3, ts5));
            x = nx; y = ny;
        }
        bool masked;
        const float t = gradient_t_at(G, x, y, masked);
        PF sc = gradient_color_at(t0s, st, len, t);
        if (lowp) {
            uint32_t sr = __float2uint_rz(clamp01(sc.r) * 255.0f + 0.5f), sg = __float2uint_rz(clamp01(sc.g) * 255.0f + 0.5f);
            uint32_t sb = __float2uint_rz(clamp01(sc.b) * 255.0f + 0.5f), sa = __float2uint_rz(clamp01(sc.a) * 255.0f + 0.5f);
            if (premul_after) { sr = div255(sr * sa); sg = div255(sg * sa); sb = div255(sb * sa); }
            // two channels per multiply, as in the solid-colour code below
            const uint32_t s_rb = sr | (sb << 16), s_ag = sg | (sa << 16);
            const uint32_t d_rb = d & 0x00ff00ffu, d_ag = (d >> 8) & 0x00ff00ffu;
            uint32_t o_rb, o_ag;
            if (src_over) { // scale_1_float (coverage folded into the source), then source_over
                const uint32_t p_rb = c == 255 ? s_rb : (((s_rb * c + 0x00ff00ffu) >> 8) & 0x00ff00ffu);
                const uint32_t p_ag = c == 255 ? s_ag : (((s_ag * c + 0x00ff00ffu) >> 8) & 0x00ff00ffu);
                const uint32_t ia = 255 - (p_ag >> 16);
                o_rb = p_rb + (((d_rb * ia + 0x00ff00ffu) >> 8) & 0x00ff00ffu);
                o_ag = p_ag + (((d_ag * ia + 0x00ff00ffu) >> 8) & 0x00ff00ffu);
            } else {        // Source: lerp_1_float(dst, src, coverage)
                const uint32_t ic = 255 - c;
                o_rb = (d_rb * ic + s_rb * c + 0x00ff00ffu) >> 8;
                o_ag = (d_ag * ic + s_ag * c + 0x00ff00ffu) >> 8;
            }
            S.px[p] = (o_rb & 0x00ff00ffu) | ((o_ag & 0x00ff00ffu) << 8); // the store truncates every lane to u8
        } else {
            if (premul_after) { sc.r *= sc.a; sc.g *= sc.a; sc.b *= sc.a; }
            if (masked) sc.r = sc.g = sc.b = sc.a = 0.0f;
            const PF dd = load_pf(d);
            PF o;
            const float cf = (float)c * (1.0f / 255.0f);
            if (src_over) { // scale_1_float, then source_over: d * (1 - sa) + s
                if (c != 255) { sc.r *= cf; sc.g *= cf; sc.b *= cf; sc.a *= cf; }
                const float ia = 1.0f - sc.a;
                o.r = mad(dd.r, ia, sc.r); o.g = mad(dd.g, ia, sc.g); o.b = mad(dd.b, ia, sc.b); o.a = mad(dd.a, ia, sc.a);
            } else if (c == 255) {
                o = sc;
            } else {        // Source: lerp_1_float(dst, src, coverage)
                o.r = mad(sc.r - dd.r, cf, dd.r); o.g = mad(sc.g - dd.g, cf, dd.g);
                o.b = mad(sc.b - dd.b, cf, dd.b); o.a = mad(sc.a - dd.a, cf, dd.a);
            }
            S.px[p] = store_pf(o);
        }
    }
    __syncwarp();
    const uint4 a = *reinterpret_cast<const uint4 *>(&S.px[lane * 8]), b = *reinterpret_cast<const uint4 *>(&S.px[lane * 8 + 4]);
    px.v[0] = a.x; px.v[1] = a.y; px.v[2] = a.z; px.v[3] = a.w; px.v[4] = b.x; px.v[5] = b.y; px.v[6] = b.z; px.v[7] = b.w;
    __syncwarp(); // the next pair's scan writes S.inside
}

// ---- the tile kernel ---------------------------------------------------------------------------------------------------
// The per-pair path is executed once per (draw, tile) by warps that are all at different points of it, so its code
// has to stay within the 32 KB L1.5 instruction cache: loops are kept rolled (the 8 pixels of a lane rotate through
// one copy of the blend code), non-anti-aliased draws reuse the anti-aliased machinery (a crossing on pixel row r at
// pixel p is entered on the four sub-rows of r at position 4p, which yields exactly 0 / 255), and rare work lives in
// __noinline__ functions.
struct WarpDraw { // what one lane holds about one upcoming (draw, tile) pair
    int tlx, tly;          // tile origin in the draw's DrawTiler-tile-local pixels
    uint32_t bounds;       // py0 | py1 << 8 | pxa << 16 | pxb << 24 (tile-local pixel bounds of the blitter rectangle)
    uint32_t flags;        // shift | rule << 4 | valid << 8
    uint32_t list_begin, n_list, paint;
};

// minimum resident CTAs per SM the compiler plans for (caps the registers at 80; 24 / 28 cost more spills, 18 / 16 lose
// occupancy — all measured)
#ifndef RW_MIN_CTAS
#define RW_MIN_CTAS 21
#endif
#ifndef RW_BLEND_UNROLL
#define RW_BLEND_UNROLL 2
#endif
constexpr int kBlendUnroll = RW_BLEND_UNROLL;
// HAIR: the batch holds hairline strokes (a second instantiation, so that batches without them run the leaner code).
template <bool MASK, bool HAIR, bool direct>
__device__ __forceinline__ void
raster_warp_tile(WarpTileSmem &S, const uint32_t tile, const uint32_t direct_mask, void *__restrict__ target, int W, int H,
                 int wtiles_x, const uint32_t *__restrict__ tile_off, const uint32_t *__restrict__ tile_pairs, const DevDraw *__restrict__ draws,
                 const uint32_t *__restrict__ row_off, const DevEdge *__restrict__ row_edges, const DevEdge *__restrict__ edges,
                 const DevPaint *__restrict__ paints, const DevStop *__restrict__ stops, unsigned long long *__restrict__ px_stats, uint32_t n_direct)
{
    const int lane = threadIdx.x & 31;
    const uint32_t d_begin = direct ? 0u : tile_off[tile], d_end = direct ? n_direct : tile_off[tile + 1];
    if (d_begin == d_end) return; // nothing touches this tile (most tiles of a sparse atlas)
    {
        uint4 *z = reinterpret_cast<uint4 *>(&S);
        for (int i = lane; i < (int)(sizeof(WarpTileSmem) / 16); i += 32) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    const int X0 = (int)(tile % (uint32_t)wtiles_x) * WT_W, Y0 = (int)(tile / (uint32_t)wtiles_x) * WT_H;
    const int prow = lane >> 2, pj = lane & 3; // this lane's pixels: row prow, columns 8 pj .. 8 pj + 7

    uint32_t dst0 = 0, dst1 = 0, dst2 = 0, dst3 = 0, dst4 = 0, dst5 = 0, dst6 = 0, dst7 = 0;
    {
        const int gy = Y0 + prow, gx = X0 + 8 * pj;
        if (gy < H) {
            const size_t o = (size_t)gy * W + gx;
            if (MASK) {
                const uint8_t *t = reinterpret_cast<const uint8_t *>(target);
                if (gx + 0 < W) dst0 = t[o + 0];
                if (gx + 1 < W) dst1 = t[o + 1];
                if (gx + 2 < W) dst2 = t[o + 2];
                if (gx + 3 < W) dst3 = t[o + 3];
                if (gx + 4 < W) dst4 = t[o + 4];
                if (gx + 5 < W) dst5 = t[o + 5];
                if (gx + 6 < W) dst6 = t[o + 6];
                if (gx + 7 < W) dst7 = t[o + 7];
            } else {
                const uint32_t *t = reinterpret_cast<const uint32_t *>(target);
                if (gx + 8 <= W && (W & 3) == 0) {
                    const uint4 a = *reinterpret_cast<const uint4 *>(t + o), b = *reinterpret_cast<const uint4 *>(t + o + 4);
                    dst0 = a.x; dst1 = a.y; dst2 = a.z; dst3 = a.w; dst4 = b.x; dst5 = b.y; dst6 = b.z; dst7 = b.w;
                } else {
                    if (gx + 0 < W) dst0 = t[o + 0];
                    if (gx + 1 < W) dst1 = t[o + 1];
                    if (gx + 2 < W) dst2 = t[o + 2];
                    if (gx + 3 < W) dst3 = t[o + 3];
                    if (gx + 4 < W) dst4 = t[o + 4];
                    if (gx + 5 < W) dst5 = t[o + 5];
                    if (gx + 6 < W) dst6 = t[o + 6];
                    if (gx + 7 < W) dst7 = t[o + 7];
                }
            }
        }
    }
    __syncwarp();

    uint32_t n_partial = 0, n_full = 0;
#pragma unroll 1
    for (uint32_t gbase = d_begin; gbase < d_end; gbase += 32) {
        // ---- every lane prepares one upcoming pair ---------------------------------------------------------------
        WarpDraw mine;
        mine.flags = 0; mine.n_list = 0; mine.list_begin = 0; mine.tlx = 0; mine.tly = 0; mine.bounds = 0; mine.paint = 0;
        const int n_group = (int)min(32u, d_end - gbase);
        if (lane < n_group && (!direct || ((direct_mask >> lane) & 1u))) {
            const DevDraw D = draws[direct ? (uint32_t)lane : tile_pairs[gbase + lane]];
            const int tlx = X0 - D.ox, tly = Y0 - D.oy;
            const int py0 = max(0, D.sy - tly), py1 = min(WT_H, D.sy + D.sh - tly);
            const int pxa = max(0, D.sx - tlx), pxb = min(WT_W, D.sx + D.sw - tlx);
            mine.tlx = tlx; mine.tly = tly;
            mine.bounds = (uint32_t)py0 | ((uint32_t)py1 << 8) | ((uint32_t)pxa << 16) | ((uint32_t)pxb << 24);
            const bool valid = py0 < py1 && pxa < pxb;
            mine.flags = (uint32_t)D.shift | ((uint32_t)(D.rule & 1) << 4) | (valid ? 0x100u : 0u) | (D.rule == 2 ? 0x200u : 0u);
            mine.paint = D.paint;
            if (valid) {
                uint32_t r = (uint32_t)(Y0 >> 3) - D.r0;
                if (D.rule == 2) r = r * D.curve_cnt + ((uint32_t)(X0 >> 5) - D.curve_off); // hairline: its cell of the bounding box
                const uint32_t lb = row_off[D.row_base + r], le = row_off[D.row_base + r + 1];
                mine.list_begin = D.list_off + lb;
                mine.n_list = le - lb;
            }
        }
        // first edges of the group's first pair
        DevEdge pre;
        pre.x = 0; pre.dx = 0; pre.ypack = 0xffffu; pre.meta = 0; // fy > ly: empty
        {
            const uint32_t lb0 = __shfl_sync(0xffffffffu, mine.list_begin, 0), n0 = __shfl_sync(0xffffffffu, mine.n_list, 0);
            if ((uint32_t)lane < n0) pre = row_edges[lb0 + lane];
        }
#pragma unroll 1
        for (int k = 0; k < n_group; k++) {
            const uint32_t flags = __shfl_sync(0xffffffffu, mine.flags, k);
            const uint32_t list_begin = __shfl_sync(0xffffffffu, mine.list_begin, k);
            const uint32_t n_list = __shfl_sync(0xffffffffu, mine.n_list, k);
            const uint32_t paint_idx = __shfl_sync(0xffffffffu, mine.paint, k);
            DevEdge E = pre;
            // prefetch the first edges of the next pair while this one is processed
            pre.ypack = 0xffffu;
            if (k + 1 < n_group) {
                const uint32_t lbn = __shfl_sync(0xffffffffu, mine.list_begin, k + 1), nn = __shfl_sync(0xffffffffu, mine.n_list, k + 1);
                if ((uint32_t)lane < nn) pre = row_edges[lbn + lane];
            }
            if (px_stats && lane == 0) atomicAdd(px_stats + 2, 1ull);
            if (!(flags & 0x100u) || n_list == 0) { if (px_stats && lane == 0) atomicAdd(px_stats + 3, 1ull); continue; }
            if (px_stats && lane == 0) atomicAdd(px_stats + 6, (unsigned long long)n_list);
            const int tlx = __shfl_sync(0xffffffffu, mine.tlx, k), tly = __shfl_sync(0xffffffffu, mine.tly, k);
            if (HAIR && (flags & 0x200u)) {
                // ---- hairline stroke: apply the row's blits that fall into this tile, one after the other -----------------
                if (!MASK) {
                    // Blits of one pixel must be applied in list order, different pixels are independent: every chunk of
                    // 32 blits is dealt into per-owner-lane queues (order kept by ranking the lanes that share an owner),
                    // then all owner lanes work through their queues at once.  The queues live in the (idle) wsum array.
                    const DevPaint &P = paints[paint_idx];
                    const bool memset_ok = P.has_memset != 0;
                    const uint32_t memset_color = P.memset_color;
                    const bool solid_so = P.kind == 0 && P.lowp && P.blend == 3; // solid colour, SourceOver, u16 pipeline
                    const uint32_t sr = P.solid16[0], sg = P.solid16[1], sb = P.solid16[2], sa = P.solid16[3];
                    uint32_t *queue = reinterpret_cast<uint32_t *>(S.wsum); // [owner lane][32]
                    int *qcnt = S.bd;
#pragma unroll 1
                    for (uint32_t cb = 0; cb < n_list; cb += 32) {
                        if (cb) {
                            E.x = -1;
                            if (cb + lane < n_list) E = row_edges[list_begin + cb + lane];
                        }
                        const bool have = cb + lane < n_list;
                        const int bx = (int)((uint32_t)E.x & 0xffffu) - X0, by = (int)((uint32_t)E.x >> 16) - Y0;
                        const bool hit = have && bx >= 0 && bx < WT_W && by >= 0 && by < WT_H;
                        const uint32_t owner = hit ? (uint32_t)(by << 2 | bx >> 3) : 32u + (uint32_t)lane;
                        const uint32_t same = __match_any_sync(0xffffffffu, owner);
                        qcnt[lane] = 0;
                        __syncwarp();
                        if (hit) {
                            const uint32_t rank = __popc(same & ((1u << lane) - 1u));
                            queue[owner * 32 + rank] = (uint32_t)(bx & 7) | ((uint32_t)E.dx << 8);
                            if (rank == 0) qcnt[owner] = __popc(same);
                        }
                        __syncwarp();
                        const int mine_cnt = qcnt[lane];
                        for (int i = 0; i < mine_cnt; i++) {
                            const uint32_t v = queue[lane * 32 + i];
                            queue[lane * 32 + i] = 0;
                            const uint32_t qi = v & 7u, c = v >> 8;
                            uint32_t d;
                            switch (qi) {
                            case 0: d = dst0; break; case 1: d = dst1; break; case 2: d = dst2; break; case 3: d = dst3; break;
                            case 4: d = dst4; break; case 5: d = dst5; break; case 6: d = dst6; break; default: d = dst7; break;
                            }
                            if (c == 255 && memset_ok) { d = memset_color; n_full++; }
                            else if (solid_so) {
                                n_partial++;
                                const uint32_t pr = c == 255 ? sr : div255(sr * c), pg = c == 255 ? sg : div255(sg * c);
                                const uint32_t pb = c == 255 ? sb : div255(sb * c), pa = c == 255 ? sa : div255(sa * c);
                                const uint32_t ia = 255 - pa;
                                d = rb_pack((pr + div255(RB_R(d) * ia)) & 0xffu, (pg + div255(RB_G(d) * ia)) & 0xffu,
                                            (pb + div255(RB_B(d) * ia)) & 0xffu, (pa + div255(RB_A(d) * ia)) & 0xffu);
                            } else {
                                n_partial++;
                                d = blend_pixel(P, stops, d, c, tlx + 8 * pj + (int)qi, tly + prow);
                            }
                            switch (qi) {
                            case 0: dst0 = d; break; case 1: dst1 = d; break; case 2: dst2 = d; break; case 3: dst3 = d; break;
                            case 4: dst4 = d; break; case 5: dst5 = d; break; case 6: dst6 = d; break; default: dst7 = d; break;
                            }
                        }
                        __syncwarp();
                    }
                    qcnt[lane] = 0;
                    __syncwarp();
                }
                continue;
            }
            const uint32_t bounds = __shfl_sync(0xffffffffu, mine.bounds, k);
            const int py0 = (int)(bounds & 0xffu), py1 = (int)((bounds >> 8) & 0xffu);
            const int pxa = (int)((bounds >> 16) & 0xffu), pxb = (int)(bounds >> 24);
            const int sh = (int)(flags & 0xfu); // 2: draw units are quarter pixels; 0: whole pixels
            const int up4 = 2 - sh;             // draw units -> the tile's quarter-pixel units
            const bool evenodd = (flags & 0x10u) != 0;
            const int lo_pos = pxa << sh, hi_pos = pxb << sh;  // draw units
            const int lo4 = pxa << 2, hi4 = pxb << 2;          // quarter pixels
            const int sub_top = (tly + py0) << sh, sub_bot = (tly + py1) << sh;
            const int row0 = tly << sh, col0 = tlx << sh;

            // ---- scatter ---------------------------------------------------------------------------------------------
            // An edge's rounded x is monotonic in y, so its two end crossings inside the tile classify it: wholly right of
            // the blitter range -> no effect; wholly left -> it only shifts the winding the rows start with (two adds into
            // a difference array instead of one crossing per sub-row); otherwise its crossings are scattered, the
            // (edge, sub-row) pairs of all such edges spread evenly over the lanes.
            bool did = false, any_left = false;
#pragma unroll 1
            for (uint32_t cb = 0; cb < n_list; cb += 32) {
                if (cb) {
                    E.ypack = 0xffffu;
                    if (cb + lane < n_list) E = row_edges[list_begin + cb + lane];
                }
                const int fy = (int)(E.ypack & 0xffffu), ly = (int)(E.ypack >> 16);
                const int ys = max(fy, sub_top), ye = min(ly, sub_bot - 1);
                int n = max(ye - ys + 1, 0); // 0: nothing (also empty lanes: fy = 0xffff, ly = 0)
                const uint32_t upbit = E.meta & 1u;
                const uint32_t xs = (uint32_t)E.x + (uint32_t)(ys - fy) * (uint32_t)E.dx;
                if (n > 0) {
                    const uint32_t xe = xs + (uint32_t)(n - 1) * (uint32_t)E.dx;
                    const int ra = ((int)(xs + 0x8000u) >> 16) - col0, rb = ((int)(xe + 0x8000u) >> 16) - col0;
                    if (min(ra, rb) >= hi_pos) n = 0;
                    else if (max(ra, rb) <= lo_pos) {
                        atomicAdd(&S.bd[(ys - row0) << up4], upbit ? -1 : 1);
                        atomicAdd(&S.bd[(ye + 1 - row0) << up4], upbit ? 1 : -1);
                        any_left = true;
                        n = 0;
                    }
                }
                if (!__any_sync(0xffffffffu, n > 0)) continue; // interior / exterior rows: nothing to scatter from this chunk
                int incl = n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += u;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                did = true;
                if (px_stats && lane == 0) atomicAdd(px_stats + 7, (unsigned long long)total);
                const uint32_t ysup = (uint32_t)ys | (upbit << 16);
#pragma unroll 1
                for (int base = 0; base < total; base += 32) {
                    const int i = base + lane;
                    int e = 0; // first lane whose inclusive count exceeds i
#pragma unroll
                    for (int step = 16; step; step >>= 1) {
                        const int v = __shfl_sync(0xffffffffu, incl, e + step - 1);
                        if (v <= i) e += step;
                    }
                    e = min(e, 31);
                    const uint32_t exs = __shfl_sync(0xffffffffu, xs, e), edx = __shfl_sync(0xffffffffu, (uint32_t)E.dx, e);
                    const uint32_t eys = __shfl_sync(0xffffffffu, ysup, e);
                    const int eexcl = __shfl_sync(0xffffffffu, incl - n, e);
                    if (i < total) {
                        const int kk = i - eexcl, y = (int)(eys & 0xffffu) + kk;
                        const bool eup = (eys >> 16) != 0;
                        const uint32_t x = exs + (uint32_t)kk * edx;
                        const int r = (int)(x + 0x8000u) >> 16;
                        const int pos = max(r - col0, lo_pos);
                        if (pos < hi_pos) {
                            const int rel = y - row0;
                            if (sh == 2) {
                                const int sr = rel & 3, pr = rel >> 2;
                                const int one = 1 << (8 * sr);
                                const uint32_t bit = 1u << (pos & 31);
                                atomicAdd(&S.wsum[pr * WT_POS + pos], eup ? -one : one);
                                atomicOr(&S.tmask[rel][pos >> 5], bit);
                                if (sr == 3) atomicOr(&S.dmask[pr][eup ? 1 : 0][pos >> 5], bit);
                            } else {
                                const int p4 = pos << 2;
                                const uint32_t bit = 1u << (p4 & 31);
                                atomicAdd(&S.wsum[rel * WT_POS + p4], eup ? -0x01010101 : 0x01010101);
                                atomicOr(&S.tmask[4 * rel + 0][p4 >> 5], bit);
                                atomicOr(&S.tmask[4 * rel + 1][p4 >> 5], bit);
                                atomicOr(&S.tmask[4 * rel + 2][p4 >> 5], bit);
                                atomicOr(&S.tmask[4 * rel + 3][p4 >> 5], bit);
                            }
                        }
                    }
                }
            }
            any_left = __any_sync(0xffffffffu, any_left);
            if (!did && !any_left) { if (px_stats && lane == 0) atomicAdd(px_stats + 4, 1ull); continue; } // bounds overlap the tile but no span does
            __syncwarp();
            // winding every sub-scanline starts with at the left end of the blitter range
            int backdrop = 0;
            if (any_left) {
                const int v = S.bd[lane];
                int incl = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += u;
                }
                backdrop = incl;
                if (v) S.bd[lane] = 0;
                if (lane == 0) S.bd[32] = 0;
                if (!did && !__any_sync(0xffffffffu, backdrop != 0)) { if (px_stats && lane == 0) atomicAdd(px_stats + 4, 1ull); continue; } // the edges left of the tile cancel out
            }

            // ---- scan: lane = sub-scanline ------------------------------------------------------------------------------
            {
                const int sr = lane & 3, dshift = 8 * sr;
                int w = backdrop;
                uint32_t carry = 0;
                const bool in0 = evenodd ? (w & 1) : (w != 0);
#pragma unroll 1
                for (int q = 0; q < 4; q++) {
                    uint32_t m = S.tmask[lane][q];
                    uint32_t T = (in0 && q == (lo4 >> 5)) ? (1u << (lo4 & 31)) : 0u, F = 0;
                    const int *wrow = &S.wsum[prow * WT_POS + 32 * q]; // prow == lane >> 2
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        const int d = (int)((((uint32_t)wrow[b] + 0x80808080u) >> dshift) & 0xffu) - 128;
                        const int wn = w + d;
                        const bool ib = evenodd ? (w & 1) : (w != 0), ia = evenodd ? (wn & 1) : (wn != 0);
                        if (ib != ia) T ^= 1u << b;
                        if (w != 0 && wn != 0 && (w ^ wn) < 0) F |= 1u << b;
                        w = wn;
                    }
                    // inside mask = prefix xor of the toggles, restricted to [lo4, hi4)
                    uint32_t x = T;
                    x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;
                    x ^= carry;
                    carry = (x >> 31) ? 0xffffffffu : 0u;
                    const int a = hi4 - 32 * q;
                    const uint32_t keep = a >= 32 ? 0xffffffffu : (a <= 0 ? 0u : ((1u << a) - 1u));
                    S.inside[lane][q] = x & keep;
                    if (sr == 3) S.flip[prow][q] = F;
                }
            }
            __syncwarp();

            // ---- coverage: lane = 8 pixels of one row ------------------------------------------------------------------------
            uint32_t c0 = 0, c1 = 0, dec = 0; // sample counts (0..16) of pixels 0..3 / 4..7, one per byte; 63-instead-of-64 flags
            {
                const uint32_t I0 = S.inside[4 * prow + 0][pj], I1 = S.inside[4 * prow + 1][pj];
                const uint32_t I2 = S.inside[4 * prow + 2][pj], I3 = S.inside[4 * prow + 3][pj];
                if (I0 | I1 | I2 | I3) {
                    auto nib = [](uint32_t v) {
                        v = v - ((v >> 1) & 0x55555555u);
                        return (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
                    };
                    const uint32_t n0 = nib(I0), n1 = nib(I1), n2 = nib(I2), n3 = nib(I3);
                    const uint32_t lo = (n0 & 0x0f0f0f0fu) + (n1 & 0x0f0f0f0fu) + (n2 & 0x0f0f0f0fu) + (n3 & 0x0f0f0f0fu);
                    const uint32_t hi = ((n0 >> 4) & 0x0f0f0f0fu) + ((n1 >> 4) & 0x0f0f0f0fu) + ((n2 >> 4) & 0x0f0f0f0fu) + ((n3 >> 4) & 0x0f0f0f0fu);
                    c0 = __byte_perm(lo, hi, 0x5140); // pixels 0, 1, 2, 3
                    c1 = __byte_perm(lo, hi, 0x7362); // pixels 4, 5, 6, 7
                    const uint32_t full3 = I3 & (I3 >> 1) & (I3 >> 2) & (I3 >> 3) & 0x11111111u;
                    uint32_t brk = 0; // bit 4k: the span breaks inside pixel k on sub-row 3
                    if (full3) {
                        const uint32_t dn = S.dmask[prow][0][pj], upm = S.dmask[prow][1][pj];
                        const uint32_t inner = (dn | upm) & 0xeeeeeeeeu; // crossings strictly inside a pixel
                        if (inner) {
                            if (evenodd) {
                                brk = ((inner >> 1) | (inner >> 2) | (inner >> 3)) & 0x11111111u;
                            } else {
                                const uint32_t mixed = dn & upm & 0xeeeeeeeeu;
                                const uint32_t f2 = S.flip[prow][pj] & 0xeeeeeeeeu & ~mixed;
                                brk = ((f2 >> 1) | (f2 >> 2) | (f2 >> 3)) & 0x11111111u;
                                uint32_t mx = mixed;
                                while (mx) { // both directions at one position: the walker's order decides
                                    const int b = __ffs(mx) - 1;
                                    mx &= mx - 1;
                                    const uint32_t pixbit = 1u << (b & ~3);
                                    if (!(full3 & pixbit) || (brk & pixbit)) continue;
                                    if (exact_span_break_list(edges, draws, direct ? (uint32_t)k : tile_pairs[gbase + k], row_edges + list_begin, n_list,
                                                              row0 + prow * 4 + 3, col0 + 32 * pj + b, col0 + lo_pos))
                                        brk |= pixbit;
                                }
                            }
                        }
                    }
                    dec = full3 & ~brk;
                }
            }
            __syncwarp();

            // ---- clear the marks this draw left -------------------------------------------------------------------------------
#pragma unroll 1
            for (int q = 0; q < 4; q++) {
                uint32_t m = S.tmask[lane][q];
                if (!m) continue;
                S.tmask[lane][q] = 0;
                int *wrow = &S.wsum[prow * WT_POS + 32 * q];
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    wrow[b] = 0;
                }
            }
            if (lane < 16) reinterpret_cast<uint4 *>(S.dmask)[lane] = make_uint4(0, 0, 0, 0);
            __syncwarp();

            // ---- blend ----------------------------------------------------------------------------------------------------------
            if (px_stats && __any_sync(0xffffffffu, (c0 | c1) != 0) && lane == 0) atomicAdd(px_stats + 5, 1ull);
            const DevPaint &P = paints[paint_idx];
            const bool memset_ok = !MASK && P.has_memset != 0;
            const bool plain = !MASK && P.kind != 2 && (P.blend == 1 || P.blend == 3); // Source / SourceOver, no pattern
            if (px_stats && !MASK && (c0 | c1)) { // the counters of rb_batch_run_counting (kept out of the blend code proper)
                uint32_t a0 = c0, a1 = c1, dd = dec;
                for (int q = 0; q < 8; q++) {
                    const uint32_t c = min(16u * (a0 & 0xffu) - (dd & 1u), 255u);
                    a0 = __funnelshift_r(a0, a1, 8); a1 >>= 8; dd >>= 4;
                    if (c == 255 && memset_ok) n_full++;
                    else if (c) n_partial++;
                }
            }
            if (plain && P.kind == 1) {
                // a gradient, u16 or f32 pipeline: the whole warp shares out the covered pixels of the tile
                if (__any_sync(0xffffffffu, (c0 | c1) != 0)) {
                    Px8 t;
                    t.v[0] = dst0; t.v[1] = dst1; t.v[2] = dst2; t.v[3] = dst3; t.v[4] = dst4; t.v[5] = dst5; t.v[6] = dst6; t.v[7] = dst7;
                    blend_tile_gradient(S, P, stops, t, c0, c1, dec, tlx, tly);
                    dst0 = t.v[0]; dst1 = t.v[1]; dst2 = t.v[2]; dst3 = t.v[3]; dst4 = t.v[4]; dst5 = t.v[5]; dst6 = t.v[6]; dst7 = t.v[7];
                }
            } else if (c0 | c1) {
                const uint32_t memset_color = P.memset_color;
                const bool src_over = P.blend == 3;
                {
                    // solid colours in the u16 pipeline inline, masks inline, everything else through blend_pixel; the lane's
                    // 8 pixels rotate through one copy of the code
                    const bool solid16 = plain && P.kind == 0 && P.lowp;
                    // Two channels per multiply: R | B << 16 and G | A << 16 hold two 16-bit lanes whose products with an
                    // 8-bit factor stay below 2^16, so (x * k + 0x00ff00ff) >> 8 & 0x00ff00ff IS div255 on both lanes
                    // (bit-identical to the per-channel u16 pipeline, half the instructions).
                    const uint32_t s_rb = P.solid16[0] | (P.solid16[2] << 16), s_ag = P.solid16[1] | (P.solid16[3] << 16);
#pragma unroll kBlendUnroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t c = min(16u * (c0 & 0xffu) - (dec & 1u), 255u);
                        c0 = __funnelshift_r(c0, c1, 8);
                        c1 >>= 8;
                        dec >>= 4;
                        uint32_t d = dst0;
                        if (c) {
                            if (MASK) {
                                d = c == 255 ? 255u : div255(d * (255 - c) + 255u * c);
                            } else if (c == 255 && memset_ok) {
                                d = memset_color;
                            } else if (solid16) {
                                const uint32_t d_rb = d & 0x00ff00ffu, d_ag = (d >> 8) & 0x00ff00ffu;
                                uint32_t o_rb, o_ag;
                                if (src_over) { // scale_1_float (coverage folded into the source), then source_over
                                    const uint32_t p_rb = c == 255 ? s_rb : (((s_rb * c + 0x00ff00ffu) >> 8) & 0x00ff00ffu);
                                    const uint32_t p_ag = c == 255 ? s_ag : (((s_ag * c + 0x00ff00ffu) >> 8) & 0x00ff00ffu);
                                    const uint32_t ia = 255 - (p_ag >> 16);
                                    o_rb = p_rb + (((d_rb * ia + 0x00ff00ffu) >> 8) & 0x00ff00ffu);
                                    o_ag = p_ag + (((d_ag * ia + 0x00ff00ffu) >> 8) & 0x00ff00ffu);
                                } else {        // Source: lerp_1_float(dst, src, coverage)
                                    const uint32_t ic = 255 - c;
                                    o_rb = (d_rb * ic + s_rb * c + 0x00ff00ffu) >> 8;
                                    o_ag = (d_ag * ic + s_ag * c + 0x00ff00ffu) >> 8;
                                }
                                d = (o_rb & 0x00ff00ffu) | ((o_ag & 0x00ff00ffu) << 8); // the store truncates every lane to u8
                            } else {
                                d = blend_pixel(P, stops, d, c, tlx + 8 * pj + q, tly + prow);
                            }
                        }
                        dst0 = dst1; dst1 = dst2; dst2 = dst3; dst3 = dst4; dst4 = dst5; dst5 = dst6; dst6 = dst7; dst7 = d;
                    }
                }
            }
        }
    }

    if (px_stats) {
        n_partial = __reduce_add_sync(0xffffffffu, n_partial);
        n_full = __reduce_add_sync(0xffffffffu, n_full);
        if (lane == 0) {
            atomicAdd(px_stats, (unsigned long long)n_partial);
            atomicAdd(px_stats + 1, (unsigned long long)n_full);
        }
    }
    {
        const int gy = Y0 + prow, gx = X0 + 8 * pj;
        if (gy < H) {
            const size_t o = (size_t)gy * W + gx;
            if (MASK) {
                uint8_t *t = reinterpret_cast<uint8_t *>(target);
                if (gx + 0 < W) t[o + 0] = (uint8_t)dst0;
                if (gx + 1 < W) t[o + 1] = (uint8_t)dst1;
                if (gx + 2 < W) t[o + 2] = (uint8_t)dst2;
                if (gx + 3 < W) t[o + 3] = (uint8_t)dst3;
                if (gx + 4 < W) t[o + 4] = (uint8_t)dst4;
                if (gx + 5 < W) t[o + 5] = (uint8_t)dst5;
                if (gx + 6 < W) t[o + 6] = (uint8_t)dst6;
                if (gx + 7 < W) t[o + 7] = (uint8_t)dst7;
            } else {
                uint32_t *t = reinterpret_cast<uint32_t *>(target);
                if (gx + 8 <= W && (W & 3) == 0) {
                    *reinterpret_cast<uint4 *>(t + o) = make_uint4(dst0, dst1, dst2, dst3);
                    *reinterpret_cast<uint4 *>(t + o + 4) = make_uint4(dst4, dst5, dst6, dst7);
                } else {
                    if (gx + 0 < W) t[o + 0] = dst0;
                    if (gx + 1 < W) t[o + 1] = dst1;
                    if (gx + 2 < W) t[o + 2] = dst2;
                    if (gx + 3 < W) t[o + 3] = dst3;
                    if (gx + 4 < W) t[o + 4] = dst4;
                    if (gx + 5 < W) t[o + 5] = dst5;
                    if (gx + 6 < W) t[o + 6] = dst6;
                    if (gx + 7 < W) t[o + 7] = dst7;
                }
            }
        }
    }
}

// One warp per 32x8 tile.  Binned mode: one tile per warp (the grid covers the tile table).  Direct mode (tile_off ==
// nullptr; batches of at most 32 draws — the per-layer batches of a tree traversal): no bin tables were built, a tile's list
// is every draw of the batch, each kept or dropped by its own row / column extent; the grid is a few CTAs per SM and every
// warp strides over the tiles, so that the (many) tiles nothing touches cost one load instead of a CTA launch.
template <bool MASK, bool HAIR, bool DIRECT>
__global__ void __launch_bounds__(WT_WARPS * 32, RW_MIN_CTAS)
k_raster_warp(void *__restrict__ target, int W, int H, int wtiles_x, uint32_t n_wtiles, const uint32_t *__restrict__ tile_off,
              const uint32_t *__restrict__ tile_pairs, const DevDraw *__restrict__ draws, const uint32_t *__restrict__ row_off,
              const DevEdge *__restrict__ row_edges, const DevEdge *__restrict__ edges, const DevPaint *__restrict__ paints,
              const DevStop *__restrict__ stops, unsigned long long *__restrict__ px_stats, const uint32_t *__restrict__ row_cols,
              uint32_t n_direct, uint32_t tile0)
{
    __shared__ WarpTileSmem s_all[WT_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpTileSmem &S = s_all[wid];
    if (!DIRECT) {
        const uint32_t tile = tile0 + blockIdx.x * WT_WARPS + wid; // tile0: first tile of this launch (banded runs)
        if (tile < n_wtiles)
            raster_warp_tile<MASK, HAIR, false>(S, tile, 0u, target, W, H, wtiles_x, tile_off, tile_pairs, draws, row_off, row_edges, edges, paints,
                                                stops, px_stats, 0u);
        return;
    }
    // direct mode: lane d keeps draw d's tile-row range
    uint32_t my_r0 = 0, my_nr = 0, my_base = 0;
    if ((uint32_t)lane < n_direct) { my_r0 = draws[lane].r0; my_nr = draws[lane].n_rows; my_base = draws[lane].row_base; }
    for (uint32_t tile = blockIdx.x * WT_WARPS + wid; tile < n_wtiles; tile += gridDim.x * WT_WARPS) {
        const uint32_t r = tile / (uint32_t)wtiles_x - my_r0, c = tile % (uint32_t)wtiles_x;
        bool hit = false;
        if (r < my_nr) {
            const uint32_t cols = row_cols[my_base + r];
            hit = (cols & 0xffffu) <= c && c <= (cols >> 16);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, hit);
        if (mask)
            raster_warp_tile<MASK, HAIR, true>(S, tile, mask, target, W, H, wtiles_x, nullptr, nullptr, draws, row_off, row_edges, edges, paints,
                                               stops, px_stats, n_direct);
        __syncwarp();
    }
}

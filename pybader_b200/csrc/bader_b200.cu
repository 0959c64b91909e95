// bader_b200.cu -- host side of libbader_b200.so: the C ABI declared in
// include/bader_b200.h on top of the kernels in kernels.cuh.
#include "kernels.cuh"
#include "seed.cuh"
#include "edge.cuh"
#include "comm.cuh"
#include "parse.cuh"
#include "format.h"

#include <chrono>
#include <cstdlib>
#include <mutex>
#include <thread>

namespace bdr {
thread_local std::string g_err;

constexpr int TX = 8, TY = 8, TZ = 32;   // edge-pass tile (z fastest)
constexpr int SX = 15;                   // stencil tile depth along x (the marched axis)
constexpr int EDGE_CX = 32;              // planes one CTA of the edge pass streams through

static dim3 tile_grid(const Grid &g, int tx = TX) {
    return dim3((g.nz + TZ - 1) / TZ, (g.ny + TY - 1) / TY, (g.nx + tx - 1) / tx);
}
constexpr size_t stencil_smem() {
    return (size_t)(SX + 2) * (TY + 2) * (TZ + 2) * sizeof(double) +
           (size_t)SX * TY * TZ * sizeof(int32_t);
}
static unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

static Weights make_weights(const double *dist_mat) {
    Weights W;
    for (int ix = -1; ix <= 1; ++ix)
        for (int iy = -1; iy <= 1; ++iy)
            for (int iz = -1; iz <= 1; ++iz)
                W.w[(ix + 1) * 9 + (iy + 1) * 3 + (iz + 1)] =
                    dist_mat[((ix + 3) % 3) * 9 + ((iy + 3) % 3) * 3 + ((iz + 3) % 3)];
    return W;
}
// the stencil kernel keeps 13 weights: w(-d) must equal w(d) bit for bit
static bool weights_symmetric(const Weights &W) {
    for (int k = 0; k < 13; ++k)
        if (memcmp(&W.w[k], &W.w[26 - k], sizeof(double)) != 0) return false;
    return true;
}
static HalfWeights half_weights(const Weights &W) {
    HalfWeights h;
    for (int k = 0; k < 14; ++k) h.w[k] = W.w[k];
    return h;
}
// fp32 step weights of the neargrid seed (seed.cuh); false when the weights are
// outside the range its "clearly uphill" argument covers (the exact kernel runs then)
static bool seed_weights(const Weights &W, SeedWeights *out) {
    double wmax = 0.0, wmin = 1e300;
    for (int k = 0; k < 13; ++k) {
        out->w[k] = (float)W.w[k];
        wmax = std::max(wmax, W.w[k]);
        wmin = std::min(wmin, W.w[k]);
    }
    out->c1 = (float)(std::ldexp(1.0, -21) * wmax * 1.0001);
    out->floor_ = (float)(std::ldexp(1.0, -96) * wmax);
    out->tag_mask = 0xfffffff0u;
    return wmin > 1e-6 && wmax < 1e6 && wmax / wmin < 1e3;
}
static TGrad make_tgrad(const double *T) {
    TGrad t;
    for (int i = 0; i < 9; ++i) t.t[i] = T[i];
    return t;
}

static Window window_of(const bdr_ctx *c) {
    Window w;
    w.own_lo = (int)c->own_lo;
    w.own_hi = (int)c->own_hi;
    if (c->halo > 0) {
        // classification needs labels two planes out, so positions are trusted
        // on planes [2, nx-3] of the window
        w.xlo = 2;
        w.xhi = c->g.nx - 3;
    } else {
        w.xlo = -2147483647;
        w.xhi = 2147483647;
    }
    return w;
}

static double *rho_ptr(bdr_ctx *c, int which) {
    int w = which;
    for (int hop = 0; hop < 3 && c->rho_alias[w] >= 0 && c->rho[w] == nullptr; ++hop)
        w = c->rho_alias[w];
    return c->rho[w];
}

static int check(bdr_ctx *c) {
    if (!c) return fail_msg("null handle");
    CU(cudaSetDevice(c->device));
    return 0;
}

static int ensure_labels(bdr_ctx *c, int which) {
    if (c->labels[which]) return 0;
    CU(cudaMalloc((void **)&c->labels[which], (size_t)c->N * sizeof(int32_t)));
    CU(cudaMemsetAsync(c->labels[which], 0, (size_t)c->N * sizeof(int32_t), c->stream));
    if (which == BDR_LABELS_BADER) c->vac_mode = 0;
    return 0;
}
static int ensure_known(bdr_ctx *c) {
    if (c->known) return 0;
    CU(cudaMalloc((void **)&c->known, (size_t)c->N));
    return 0;
}
static int ensure_bits(bdr_ctx *c) {
    if (c->ebits) return 0;
    c->nzw = (c->g.nz + 31) / 32;
    const size_t words = (size_t)c->g.nx * c->g.ny * c->nzw;
    CU(cudaMalloc((void **)&c->ebits, 7 * words * sizeof(uint32_t)));
    c->vbits = c->ebits + words;
    c->cbits = c->vbits + words;
    c->sbits = c->cbits + words;
    c->eqz = c->sbits + words;
    c->eqy = c->eqz + words;
    c->eqx = c->eqy + words;
    return 0;
}
// the label set's equality bits no longer describe it (labels rewritten wholesale)
static void eq_invalidate(bdr_ctx *c, int which) {
    if (which < 0 || c->eq_which == which) {
        c->eq_valid = false;
        c->eq_pending_n = 0;
    }
}
// voxels a trace relabelled (c->list2[0, n)) join the list k_eq_update will patch in
static int eq_note_changed(bdr_ctx *c, int which, int64_t n, bool have_list) {
    if (!c->eq_valid || c->eq_which != which || n == 0) return 0;
    if (!have_list || c->eq_pending_n + n > c->N / 32) {   // cheaper to recompute than to patch
        eq_invalidate(c, which);
        return 0;
    }
    if (c->eq_pending_n + n > c->eq_pending_cap) {
        const int64_t cap = (c->eq_pending_n + n) * 2 + 4096;
        int32_t *grown = nullptr;
        CU(cudaMalloc((void **)&grown, (size_t)cap * sizeof(int32_t)));
        if (c->eq_pending_n)
            CU(cudaMemcpyAsync(grown, c->eq_pending, (size_t)c->eq_pending_n * sizeof(int32_t),
                               cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (c->eq_pending) cudaFree(c->eq_pending);
        c->eq_pending = grown;
        c->eq_pending_cap = cap;
    }
    CU(cudaMemcpyAsync(c->eq_pending + c->eq_pending_n, c->list2, (size_t)n * sizeof(int32_t),
                       cudaMemcpyDeviceToDevice, c->stream));
    c->eq_pending_n += n;
    return 0;
}

static int ensure_rho(bdr_ctx *c, int which) {
    if (c->rho[which]) return 0;
    CU(cudaMalloc((void **)&c->rho[which], (size_t)c->N * sizeof(double)));
    c->rho_alias[which] = -1;
    return 0;
}
static int ensure_sums(bdr_ctx *c, int64_t n) {
    if (c->d_sums_cap >= n) return 0;
    if (c->d_sums) cudaFree(c->d_sums);
    c->d_sums = nullptr;
    c->d_sums_cap = 0;
    const int64_t cap = std::max<int64_t>(n + n / 4, 64);
    CU(cudaMalloc((void **)&c->d_sums, (size_t)cap * sizeof(double)));
    c->d_sums_cap = cap;
    return 0;
}
static int ensure_slots(bdr_ctx *c, int64_t n) {
    if (c->slots_cap >= n) return 0;
    c->maxima_fresh[0] = c->maxima_fresh[1] = false;   // c->roots goes away
    for (int32_t **p : {&c->roots, &c->minidx, &c->rank}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    c->slots_cap = 0;
    const int64_t cap = std::max<int64_t>(n + n / 4, 4096);
    CU(cudaMalloc((void **)&c->roots, (size_t)cap * sizeof(int32_t)));
    CU(cudaMalloc((void **)&c->minidx, (size_t)cap * sizeof(int32_t)));
    CU(cudaMalloc((void **)&c->rank, (size_t)cap * sizeof(int32_t)));
    c->slots_cap = cap;
    return 0;
}

// ---------------------------------------------------------------------------
// numbering: volume numbers ascend with the first voxel (C order) carrying them
// ---------------------------------------------------------------------------
// `keys` = first voxel per current id, `n` ids.  Produces rank (id -> number)
// on the device; returns in `order` the ids sorted by key.
static int rank_from_first(bdr_ctx *c, int64_t n, std::vector<int32_t> &order, bool *identity) {
    std::vector<int32_t> keys((size_t)n);
    CU(cudaMemcpyAsync(keys.data(), c->minidx, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
    order.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) order[(size_t)i] = (int32_t)i;
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        return keys[(size_t)a] != keys[(size_t)b] ? keys[(size_t)a] < keys[(size_t)b] : a < b;
    });
    std::vector<int32_t> rank((size_t)n);
    bool ident = true;
    for (int64_t r = 0; r < n; ++r) {
        rank[(size_t)order[(size_t)r]] = (int32_t)r;
        ident &= (order[(size_t)r] == r);
    }
    if (identity) *identity = ident;
    CU(cudaMemcpyAsync(c->rank, rank.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------------------
// ongrid: stencil -> pointer codes -> resolve -> numbering
// ---------------------------------------------------------------------------
// slot numbering of a slab window: [0, 2*ny*nz) are the exit-plane voxels,
// real maxima follow
static int exit_base_of(const bdr_ctx *c) { return c->halo > 0 ? 2 * c->g.ny * c->g.nz : 0; }

// tiles of the last stencil pass (the resolve pass walks the same tiling)
static int stencil_tx(const bdr_ctx *c) { return c->seed_f32 ? FX : SX; }
static int stencil_ty(const bdr_ctx *c) { return c->seed_f32 ? FY : TY; }
static int stencil_tz(const bdr_ctx *c) { return c->seed_f32 ? FZ : TZ; }
static dim3 stencil_grid(const bdr_ctx *c) {
    const int tx = stencil_tx(c), ty = stencil_ty(c), tz = stencil_tz(c);
    return dim3((c->g.nz + tz - 1) / tz, (c->g.ny + ty - 1) / ty, (c->g.nx + tx - 1) / tx);
}
static int ensure_tiles(bdr_ctx *c) {
    const dim3 gr = stencil_grid(c);
    const int64_t n = (int64_t)gr.x * gr.y * gr.z;
    if (!c->tile_hist) CU(cudaMalloc((void **)&c->tile_hist, 65536 * sizeof(unsigned)));
    if (c->tiles_cap >= n) return 0;
    if (c->tile_keys) cudaFree(c->tile_keys);
    if (c->tile_order) cudaFree(c->tile_order);
    c->tile_keys = nullptr;
    c->tile_order = nullptr;
    c->tiles_cap = 0;
    CU(cudaMalloc((void **)&c->tile_keys, (size_t)n * sizeof(uint32_t)));
    CU(cudaMalloc((void **)&c->tile_order, (size_t)n * sizeof(int32_t)));
    c->tiles_cap = n;
    return 0;
}

// the stencil pass over planes [x_begin, x_end) (x_begin a multiple of the tile depth)
static int stencil_launch(bdr_ctx *c, const Weights &W_full, int x_begin, int x_end) {
    int32_t *code = c->labels[BDR_LABELS_BADER];
    const double *rho = rho_ptr(c, BDR_RHO_REFERENCE);
    const int xb = exit_base_of(c);
    dim3 grid = stencil_grid(c);
    const int tx = stencil_tx(c);
    grid.z = (x_end - x_begin + tx - 1) / tx;
    // vacuum comes from the fused tolerance test when the labels were made
    // by bdr_vacuum_assign / bdr_clear_labels on this handle, else from
    // the label array itself (-1 entries)
    if (c->seed_f32) {
        SeedWeights Wf;
        seed_weights(W_full, &Wf);
        const size_t smem = seed_smem();
        const double *W_dev = c->d_seedw;
        using SeedKernel = void (*)(const double *, int32_t *, Grid, SeedWeights, const double *, double,
                                    unsigned long long *, int32_t *, int64_t, int, int, uint32_t *);
        const SeedKernel table[3][2] = {
            {k_seed_pointers<VAC_NONE, false>, k_seed_pointers<VAC_NONE, true>},
            {k_seed_pointers<VAC_TOL, false>, k_seed_pointers<VAC_TOL, true>},
            {k_seed_pointers<VAC_LABELS, false>, k_seed_pointers<VAC_LABELS, true>}};
        const SeedKernel kern = table[c->vac_mode][xb > 0 ? 1 : 0];
        LAUNCH(c, BDR_K_STENCIL, kern, grid, 256, smem, rho, code, c->g, Wf, W_dev,
               c->vac_mode == VAC_TOL ? c->vac_tol : 0.0, c->d_cnt + CNT_ROOTS, c->roots, c->slots_cap,
               xb, x_begin, c->tile_keys);
        return 0;
    }
    const HalfWeights W = half_weights(W_full);
    const size_t smem = stencil_smem();
    if (c->vac_mode == VAC_NONE)
        LAUNCH(c, BDR_K_STENCIL, (k_ongrid_pointers<SX, TY, TZ, VAC_NONE>), grid, 256, smem, rho, code,
               c->g, W, 0.0, c->d_cnt + CNT_ROOTS, c->roots, c->slots_cap, xb, x_begin, c->tile_keys);
    else if (c->vac_mode == VAC_TOL)
        LAUNCH(c, BDR_K_STENCIL, (k_ongrid_pointers<SX, TY, TZ, VAC_TOL>), grid, 256, smem, rho, code,
               c->g, W, c->vac_tol, c->d_cnt + CNT_ROOTS, c->roots, c->slots_cap, xb, x_begin,
               c->tile_keys);
    else
        LAUNCH(c, BDR_K_STENCIL, (k_ongrid_pointers<SX, TY, TZ, VAC_LABELS>), grid, 256, smem, rho,
               code, c->g, W, 0.0, c->d_cnt + CNT_ROOTS, c->roots, c->slots_cap, xb, x_begin,
               c->tile_keys);
    return 0;
}

// which stencil kernel seeds this call: the fp32-ranked one for 'neargrid'
// (seed.cuh), the bit-exact fp64 one for 'ongrid'
static int choose_seed(bdr_ctx *c, int method, const Weights &W) {
    SeedWeights Wf;
    c->seed_f32 = method == BDR_METHOD_NEARGRID && seed_weights(W, &Wf) && !getenv("BDR_SEED_EXACT");
    if (c->seed_f32) {
        // the exact fallback of the seed kernel reads the fp64 weights from global memory
        if (!c->d_seedw) CU(cudaMalloc((void **)&c->d_seedw, sizeof(Weights)));
        CU(cudaMemcpyAsync(c->d_seedw, W.w, sizeof(Weights), cudaMemcpyHostToDevice, c->stream));
    }
    return 0;
}

// pointer codes -> terminal codes, tile by tile in order of decreasing density
// (seed.cuh K2); with minidx the first voxel of every slot is recorded on the way
static int resolve_dev(bdr_ctx *c, int32_t *minidx) {
    int32_t *code = c->labels[BDR_LABELS_BADER];
    if (getenv("BDR_RESOLVE_OLD")) {
        LAUNCH(c, BDR_K_RESOLVE, k_resolve, blocks_for(c->N, 1024), 256, 0, code, c->N, minidx, 3);
        return 0;
    }
    const dim3 gr = stencil_grid(c);
    const int n = (int)(gr.x * gr.y * gr.z);
    CU(cudaMemsetAsync(c->tile_hist, 0, 65536 * sizeof(unsigned), c->stream));
    LAUNCH(c, BDR_K_RESOLVE, k_tile_hist, blocks_for(n, 256), 256, 0, c->tile_keys, n, c->tile_hist);
    LAUNCH(c, BDR_K_RESOLVE, k_tile_scan, 1, 1024, 0, c->tile_hist);
    LAUNCH(c, BDR_K_RESOLVE, k_tile_scatter, blocks_for(n, 256), 256, 0, c->tile_keys, n, c->tile_hist,
           c->tile_order);
    const bool vec = (c->g.nz & 3) == 0;
    if (c->seed_f32) {
        if (vec)
            LAUNCH(c, BDR_K_RESOLVE, (k_resolve_tiles<FX, FY, FZ, true>), n, 256, 0, code, c->g,
                   c->tile_order, (int)gr.x, (int)gr.y, minidx);
        else
            LAUNCH(c, BDR_K_RESOLVE, (k_resolve_tiles<FX, FY, FZ, false>), n, 256, 0, code, c->g,
                   c->tile_order, (int)gr.x, (int)gr.y, minidx);
    } else {
        if (vec)
            LAUNCH(c, BDR_K_RESOLVE, (k_resolve_tiles<SX, TY, TZ, true>), n, 256, 0, code, c->g,
                   c->tile_order, (int)gr.x, (int)gr.y, minidx);
        else
            LAUNCH(c, BDR_K_RESOLVE, (k_resolve_tiles<SX, TY, TZ, false>), n, 256, 0, code, c->g,
                   c->tile_order, (int)gr.x, (int)gr.y, minidx);
    }
    return 0;
}

static int stencil_dev(bdr_ctx *c, const Weights &W_full, int64_t *n_real) {
    TRY(ensure_labels(c, BDR_LABELS_BADER));
    TRY(ensure_slots(c, 4096));
    if (!weights_symmetric(W_full))
        return fail_msg("bader_calc: dist_mat[-d] != dist_mat[d]; not a step-length table");
    TRY(ensure_tiles(c));
    const Weights &W = W_full;
    for (int attempt = 0; attempt < 2; ++attempt) {
        TRY(zero_counter(c, CNT_ROOTS));
        TRY(stencil_launch(c, W, 0, c->g.nx));
        TRY(read_counters(c));
        const int64_t n = (int64_t)c->h_cnt[CNT_ROOTS];
        if (n <= c->slots_cap) break;
        if (attempt == 1) return fail_msg("maxima list overflow");
        // codes are garbage beyond capacity: restore vacuum/unassigned and redo
        // (only voxel classes -1 / not -1 matter to the stencil pass)
        TRY(ensure_slots(c, n));
    }
    *n_real = (int64_t)c->h_cnt[CNT_ROOTS];
    return 0;
}

// bdr_run: the density arrives in x chunks on a copy stream and the stencil
// pass follows it chunk by chunk on the compute stream, so the first kernel of
// the pipeline hides under the PCIe transfer (which is ~2.5x longer than all
// kernels together at 1024^3).  A chunk's stencil needs one plane of the next
// chunk, and chunk 0 needs the last plane (periodic), so it runs last.
static int upload_and_stencil_dev(bdr_ctx *c, const double *host, const Weights &W_full,
                                  int64_t *n_real) {
    TRY(ensure_rho(c, BDR_RHO_REFERENCE));
    TRY(ensure_labels(c, BDR_LABELS_BADER));
    TRY(ensure_slots(c, 4096));
    if (!weights_symmetric(W_full))
        return fail_msg("bader_calc: dist_mat[-d] != dist_mat[d]; not a step-length table");
    TRY(ensure_tiles(c));
    const Weights &W = W_full;
    const int CH = 4 * stencil_tx(c);
    const int nchunks = (c->g.nx + CH - 1) / CH;
    if (!c->copy_stream) CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    while ((int)c->chunk_events.size() < nchunks) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->chunk_events.push_back(e);
    }
    const int64_t plane = (int64_t)c->g.ny * c->g.nz;
    double *rho = c->rho[BDR_RHO_REFERENCE];
    CU(cudaStreamSynchronize(c->stream));  // nothing may still read the old density
    const auto t_up = std::chrono::steady_clock::now();
    for (int i = 0; i < nchunks; ++i) {
        const int xa = i * CH, xe = std::min(c->g.nx, xa + CH);
        CU(cudaMemcpyAsync(rho + xa * plane, host + xa * plane, (size_t)(xe - xa) * plane * sizeof(double),
                           cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->chunk_events[(size_t)i], c->copy_stream));
    }
    TRY(zero_counter(c, CNT_ROOTS));
    for (int i = 1; i < nchunks; ++i) {
        const int xa = i * CH, xe = std::min(c->g.nx, xa + CH);
        CU(cudaStreamWaitEvent(c->stream, c->chunk_events[(size_t)std::min(i + 1, nchunks - 1)], 0));
        TRY(stencil_launch(c, W, xa, xe));
    }
    CU(cudaStreamWaitEvent(c->stream, c->chunk_events[(size_t)nchunks - 1], 0));
    TRY(stencil_launch(c, W, 0, std::min(c->g.nx, CH)));
    TRY(read_counters(c));
    c->dbg_upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_up).count();
    int64_t n = (int64_t)c->h_cnt[CNT_ROOTS];
    if (n > c->slots_cap) {  // more maxima than slots: redo on the resident density
        TRY(ensure_slots(c, n));
        TRY(stencil_dev(c, W_full, &n));
    }
    *n_real = n;
    return 0;
}

// slot codes -> volume numbers: volumes are numbered by their first voxel in C
// order, which is the reference's order of discovery.  With
// have_first the resolve pass has already filled c->minidx.
static int number_slots_dev(bdr_ctx *c, bool have_first) {
    const int64_t n = c->n_max;
    if (n == 0) return 0;
    int32_t *code = c->labels[BDR_LABELS_BADER];
    if (!have_first) {
        CU(cudaMemsetAsync(c->minidx, 0x7f, (size_t)n * sizeof(int32_t), c->stream));
        LAUNCH(c, BDR_K_FIRST, k_first_voxel_slots, blocks_for(c->N, 1024), 256, 0, code, 0, (int)c->N,
               c->minidx);
    }
    std::vector<int32_t> order;
    TRY(rank_from_first(c, n, order, nullptr));
    std::vector<int32_t> roots((size_t)n);
    CU(cudaMemcpyAsync(roots.data(), c->roots, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const int64_t nyz = (int64_t)c->g.ny * c->g.nz;
    c->maxima.assign((size_t)n * 3, 0);
    for (int64_t r = 0; r < n; ++r) {
        const int64_t v = roots[(size_t)order[(size_t)r]];
        c->maxima[(size_t)r * 3 + 0] = v / nyz;
        c->maxima[(size_t)r * 3 + 1] = (v / c->g.nz) % c->g.ny;
        c->maxima[(size_t)r * 3 + 2] = v % c->g.nz;
    }
    LAUNCH(c, BDR_K_RELABEL, k_relabel_slots, blocks_for(c->N, 1024), 256, 0, code, c->N, c->rank);
    return 0;
}

// stencil -> pointer codes -> resolved slot codes (-2 - slot; vacuum -1).  With
// `number` the codes are then turned into volume numbers; bader_calc('neargrid')
// keeps the codes as labels through its rounds and numbers once at the end.
static int ongrid_dev(bdr_ctx *c, const Weights &W_full, int64_t seeded = -1, bool number = true) {
    if (c->halo > 0) return fail_msg("bader_calc: slab windows are driven through the bdr_slab_* entry points");
    int64_t n = seeded;  // >= 0: the stencil pass already ran (bdr_run's pipelined upload)
    if (n < 0) TRY(stencil_dev(c, W_full, &n));
    c->n_max = n;
    c->maxima.assign((size_t)n * 3, 0);
    if (n == 0) return 0;
    if (number) CU(cudaMemsetAsync(c->minidx, 0x7f, (size_t)n * sizeof(int32_t), c->stream));
    TRY(resolve_dev(c, number ? c->minidx : (int32_t *)nullptr));
    if (number) TRY(number_slots_dev(c, true));
    return 0;
}

// ---------------------------------------------------------------------------
// refinement pieces
// ---------------------------------------------------------------------------
static int edge_find_dev(bdr_ctx *c, int which, int64_t *edges, int64_t n_changed = -1,
                         int sticky_mode = 0) {
    TRY(ensure_known(c));
    TRY(ensure_bits(c));
    if (!c->labels[which]) return fail_msg("edge_find: label set is empty");
    TRY(ensure(&c->list, &c->list_cap, std::max<int64_t>(c->N / 16, 1024)));
    const bool old_bits = getenv("BDR_EDGE_OLD") != nullptr;
    if (!old_bits) TRY(ensure(&c->defer, &c->defer_cap, std::max<int64_t>(c->N / 64, 4096)));
    int64_t n = 0;
    for (int pass = 0;; ++pass) {
        if (old_bits) {
            const dim3 grid((c->g.nz + 127) / 128, (c->g.ny + 7) / 8, (c->g.nx + EDGE_CX - 1) / EDGE_CX);
            if ((c->g.nz & 3) == 0)
                LAUNCH(c, BDR_K_EDGE_FLAG, (k_edge_bits<EDGE_CX, true>), grid, 256, 0, c->labels[which],
                       c->g, c->ebits, c->vbits, c->nzw);
            else
                LAUNCH(c, BDR_K_EDGE_FLAG, (k_edge_bits<EDGE_CX, false>), grid, 256, 0, c->labels[which],
                       c->g, c->ebits, c->vbits, c->nzw);
        } else {
            // label equality bits -> candidate bits (edge.cuh); voxels next to vacuum
            // are classified exactly from a list (checked for overflow below)
            if (c->eq_valid && c->eq_which == which && !getenv("BDR_EQ_RECOMPUTE")) {
                // the equality bits of these labels exist; patch in what was relabelled since
                // (the vacuum-seen flag of the pass that made them stays: a superset is fine)
                TRY(zero_counter(c, CNT_DEFER));
                if (c->eq_pending_n > 0)
                    LAUNCH(c, BDR_K_EDGE_FLAG, k_eq_update, blocks_for(c->eq_pending_n, 128), 128, 0,
                           c->labels[which], c->g, c->nzw, c->eqz, c->eqy, c->eqx, c->vbits, c->eq_pending,
                           c->eq_pending_n);
                c->eq_pending_n = 0;
                if (c->halo > 0) {
                    // a slab's halo labels come from its neighbours (halo exchanges): the bit planes
                    // that compare with them -- the halos and the owned plane next to each -- are
                    // recomputed, 2 * (halo + 1) planes instead of the whole window
                    const int H = c->halo, Wp = c->g.nx;
                    const int lo_end = std::min(H + 1, Wp), hi_begin = std::max(Wp - H - 1, lo_end);
                    const dim3 gl((c->nzw + 3) / 4, (c->g.ny + 7) / 8, (lo_end + EDGE_CX - 1) / EDGE_CX);
                    LAUNCH(c, BDR_K_EDGE_EQ, (k_label_eq_bits<4, EDGE_CX>), gl, 256, 0, c->labels[which], c->g,
                           c->nzw, c->eqz, c->eqy, c->eqx, c->vbits, c->d_cnt + CNT_VACSEEN, 0, lo_end);
                    if (hi_begin < Wp) {
                        const dim3 gh((c->nzw + 3) / 4, (c->g.ny + 7) / 8, (Wp - hi_begin + EDGE_CX - 1) / EDGE_CX);
                        LAUNCH(c, BDR_K_EDGE_EQ, (k_label_eq_bits<4, EDGE_CX>), gh, 256, 0, c->labels[which],
                               c->g, c->nzw, c->eqz, c->eqy, c->eqx, c->vbits, c->d_cnt + CNT_VACSEEN,
                               hi_begin, Wp);
                    }
                }
            } else {
            CU(cudaMemsetAsync(c->d_cnt + CNT_DEFER, 0, 2 * sizeof(unsigned long long), c->stream));
            const dim3 ga((c->nzw + 3) / 4, (c->g.ny + 7) / 8, (c->g.nx + EDGE_CX - 1) / EDGE_CX);
            LAUNCH(c, BDR_K_EDGE_EQ, (k_label_eq_bits<4, EDGE_CX>), ga, 256, 0, c->labels[which], c->g,
                   c->nzw, c->eqz, c->eqy, c->eqx, c->vbits, c->d_cnt + CNT_VACSEEN, 0, c->g.nx);
            c->eq_valid = true;
            c->eq_which = which;
            c->eq_pending_n = 0;
            }
            const dim3 gb((c->g.ny * c->nzw + 255) / 256, (c->g.nx + 15) / 16);
            LAUNCH(c, BDR_K_EDGE_FLAG, (k_edge_from_eq<16>), gb, 256, 0, c->eqz, c->eqy, c->eqx, c->vbits, c->g,
                   c->nzw, c->ebits, c->d_cnt + CNT_VACSEEN, c->d_cnt + CNT_DEFER, c->defer,
                   c->defer_cap);
            LAUNCH(c, BDR_K_EDGE_FLAG, k_edge_deferred, 148 * 8, 128, 0, c->labels[which], c->g, c->nzw,
                   c->ebits, c->defer, c->d_cnt + CNT_DEFER, c->defer_cap);
        }
        // masked pass (n_changed >= 0): list only the edges in the 27-neighbourhood
        // of the voxels in c->list2 (the ones the last trace relabelled)
        const uint32_t *mask = nullptr;
        if (n_changed >= 0) {
            const size_t words = (size_t)c->g.nx * c->g.ny * c->nzw;
            CU(cudaMemsetAsync(c->cbits, 0, words * sizeof(uint32_t), c->stream));
            if (n_changed > 0)
                LAUNCH(c, BDR_K_EDGE_CHECK, k_bits_from_list, blocks_for(n_changed, 256), 256, 0, c->cbits,
                       c->g, c->nzw, c->list2, n_changed);
            mask = c->cbits;
        }
        for (int attempt = 0; attempt < 2; ++attempt) {
            TRY(zero_counter(c, CNT_EDGES));
            LAUNCH(c, BDR_K_EDGE_DILATE, k_edge_known,
                   dim3((c->nzw + 3) / 4, (c->g.ny + 7) / 8, (c->g.nx + 7) / 8), 256, 0, c->ebits, c->vbits,
                   c->known, c->g, c->nzw, c->d_cnt + CNT_EDGES, c->list, c->list_cap, mask, c->sbits,
                   sticky_mode);
            TRY(read_counters(c));
            n = (int64_t)c->h_cnt[CNT_EDGES];
            if (n <= c->list_cap) break;
            TRY(ensure(&c->list, &c->list_cap, n));  // the list overflowed: grow it and redo the cheap half
        }
        const int64_t nd = old_bits ? 0 : (int64_t)c->h_cnt[CNT_DEFER];
        if (nd <= c->defer_cap) break;
        if (pass == 1) return fail_msg("edge_find: deferred list overflow");
        TRY(ensure(&c->defer, &c->defer_cap, nd));  // redo the pass with room for every deferred voxel
    }
    c->list_n = n;  // every candidate of the window; the trace kernel skips what it does not own
    c->window_fresh = n_changed < 0;   // a full pass over the whole window (exact or conservative)
    c->halo_known_current = false;
    *edges = 0;
    if (n == 0) return 0;
    if (sticky_mode != 0) {
        // conservative passes inside bader_calc('neargrid') skip the density half:
        // a candidate that is a maximum stays listed; the maxima of the stencil pass
        // (c->roots) are marked interior instead and the trace kernels skip them
        if (c->n_max > 0 && c->roots)
            LAUNCH(c, BDR_K_EDGE_CONFIRM, k_mark_interior, blocks_for(c->n_max, 128), 128, 0, c->known,
                   c->roots, c->n_max);
        *edges = n;
        return 0;
    }
    // density half of the classification: drop the candidates that are maxima
    if (c->maxima_fresh[which] && c->roots && !getenv("BDR_CONFIRM_ALL")) {
        // the maxima are known (k_edge_confirm_roots): test those, not every candidate
        int64_t nf = 0;
        if (c->n_max > 0) {
            TRY(ensure(&c->list3, &c->list3_cap, c->n_max));
            TRY(zero_counter(c, CNT_CENTRES));
            LAUNCH(c, BDR_K_EDGE_CONFIRM, k_edge_confirm_roots, blocks_for(c->n_max, 128), 128, 0,
                   rho_ptr(c, BDR_RHO_REFERENCE), c->labels[which], c->g, c->ebits, c->nzw, c->roots,
                   c->n_max, c->d_cnt + CNT_CENTRES, c->list3, c->list3_cap);
            TRY(read_counters(c));
            nf = (int64_t)c->h_cnt[CNT_CENTRES];
        }
        if (nf > 0) {
            LAUNCH(c, BDR_K_EDGE_CONFIRM, k_edge_fix_clear, blocks_for(nf, 128), 128, 0, c->ebits, c->g,
                   c->nzw, (int32_t *)nullptr, c->list3, nf);
            LAUNCH(c, BDR_K_EDGE_CONFIRM, k_edge_fix_known, blocks_for(nf * 27, 128), 128, 0, c->ebits,
                   c->vbits, c->known, c->g, c->nzw, (int32_t *)nullptr, c->list3, nf);
        }
        if (c->halo == 0) {
            *edges = n - nf;
        } else {
            // a slab reports the edges it owns: the edge bits left on its own planes
            const int64_t row_words = (int64_t)c->g.ny * c->nzw;
            const int64_t lo = (int64_t)c->halo * row_words, cnt = (int64_t)(c->g.nx - 2 * c->halo) * row_words;
            TRY(zero_counter(c, CNT_NEWEDGE));
            LAUNCH(c, BDR_K_EDGE_CONFIRM, k_popcount_words, 148 * 8, 256, 0, c->ebits + lo, cnt,
                   c->d_cnt + CNT_NEWEDGE);
            TRY(read_counters(c));
            *edges = (int64_t)c->h_cnt[CNT_NEWEDGE];
        }
        return 0;   // the maxima stay in the list; its consumers skip entries with known != -2
    }
    TRY(ensure(&c->list3, &c->list3_cap, 4096));
    for (int attempt = 0; attempt < 2; ++attempt) {
        TRY(zero_counter(c, CNT_NEWEDGE));
        TRY(zero_counter(c, CNT_CENTRES));
        LAUNCH(c, BDR_K_EDGE_CONFIRM, k_edge_confirm, blocks_for(n, 128), 128, 0,
               rho_ptr(c, BDR_RHO_REFERENCE), c->labels[which], c->g, window_of(c), c->list, n,
               c->d_cnt + CNT_NEWEDGE, c->d_cnt + CNT_CENTRES, c->list3, c->list3_cap);
        TRY(read_counters(c));
        if ((int64_t)c->h_cnt[CNT_CENTRES] <= c->list3_cap) break;
        TRY(ensure(&c->list3, &c->list3_cap, (int64_t)c->h_cnt[CNT_CENTRES]));
    }
    *edges = (int64_t)c->h_cnt[CNT_NEWEDGE];
    const int64_t nf = (int64_t)c->h_cnt[CNT_CENTRES];
    if (nf > 0) {
        LAUNCH(c, BDR_K_EDGE_CONFIRM, k_edge_fix_clear, blocks_for(nf, 128), 128, 0, c->ebits, c->g,
               c->nzw, c->list, c->list3, nf);
        LAUNCH(c, BDR_K_EDGE_CONFIRM, k_edge_fix_known, blocks_for(nf * 27, 128), 128, 0, c->ebits,
               c->vbits, c->known, c->g, c->nzw, c->list, c->list3, nf);
        LAUNCH(c, BDR_K_EDGE_CONFIRM, k_edge_fix_tomb, blocks_for(nf, 128), 128, 0, c->list, c->list3,
               nf);
    }
    return 0;
}

// incremental edge update between full passes (kernels.cuh K5'): consumes the
// changed list (c->list2), produces the next trace list (c->list)
static int incremental_dev(bdr_ctx *c, int which, int64_t n_changed, int64_t *queued) {
    *queued = 0;
    c->list_n = 0;
    c->window_fresh = false;
    c->halo_known_current = false;
    if (n_changed == 0) return 0;
    const int64_t cap = std::min<int64_t>(n_changed * 27, c->N);
    TRY(ensure(&c->list3, &c->list3_cap, cap));
    TRY(zero_counter(c, CNT_CENTRES));
    if (n_changed * 64 > c->N) {
        // a large round: marking with plain stores and one streaming compaction
        // of the known array beats millions of contended byte exchanges
        LAUNCH(c, BDR_K_EDGE_CHECK, k_inc_mark, blocks_for(n_changed * 27, 128), 128, 0,
               c->labels[which], c->known, c->g, c->list2, n_changed);
        LAUNCH(c, BDR_K_EDGE_CHECK, k_compact_known, blocks_for(c->N, 256), 256, 0, c->known, c->N,
               (int8_t)-6, c->d_cnt + CNT_CENTRES, c->list3, c->list3_cap);
    } else {
        LAUNCH(c, BDR_K_EDGE_CHECK, k_inc_collect, blocks_for(n_changed * 27, 128), 128, 0,
               c->labels[which], c->known, c->g, c->list2, n_changed, c->d_cnt + CNT_CENTRES,
               c->list3, c->list3_cap);
    }
    TRY(read_counters(c));
    const int64_t nc = (int64_t)c->h_cnt[CNT_CENTRES];
    if (nc > c->list3_cap) return fail_msg("incremental candidate list overflow");
    if (nc == 0) return 0;
    TRY(ensure(&c->list, &c->list_cap, nc));
    TRY(zero_counter(c, CNT_NEWEDGE));
    LAUNCH(c, BDR_K_EDGE_CHECK, k_inc_classify, blocks_for(nc, 128), 128, 0,
           rho_ptr(c, BDR_RHO_REFERENCE), c->labels[which], c->known, c->g, c->list3, nc,
           c->d_cnt + CNT_NEWEDGE, c->list, c->list_cap);
    TRY(read_counters(c));
    const int64_t nq = (int64_t)c->h_cnt[CNT_NEWEDGE];
    if (nq > 0)
        LAUNCH(c, BDR_K_EDGE_CHECK, k_inc_dilate, blocks_for(nq * 27, 128), 128, 0, c->known, c->g,
               c->list, nq);
    // maxima stay interior through the conservative rounds (see k_mark_interior)
    if (c->n_max > 0 && c->roots)
        LAUNCH(c, BDR_K_EDGE_CHECK, k_mark_interior, blocks_for(c->n_max, 128), 128, 0, c->known, c->roots,
               c->n_max);
    c->list_n = nq;
    *queued = nq;
    return 0;
}

// one Jacobi iteration of the trajectory re-trace over c->list[0..list_n)
static int trace_dev(bdr_ctx *c, int which, const Weights &W, const TGrad &T, int64_t *changed,
                     bool want_changed_list) {
    const int64_t n = c->list_n;
    *changed = 0;
    if (n == 0) return 0;
    if (want_changed_list) TRY(ensure(&c->list2, &c->list2_cap, n));
    TRY(ensure(&c->list3, &c->list3_cap, n));  // overflow list can never overflow
    const int step_cap = 1 << 20;
    CU(cudaMemsetAsync(c->d_cnt + CNT_CHANGED, 0, sizeof(unsigned long long) * 2, c->stream));
    CU(cudaMemsetAsync(c->d_cnt + CNT_OVERFLOW, 0, sizeof(unsigned long long) * 2, c->stream));
    TRY(zero_counter(c, CNT_STEPS));
    TRY(zero_counter(c, CNT_ESCAPED));
    // list entries per warp: lanes refill from their warp's chunk, so longer chunks
    // hide the spread of trajectory lengths, but the voxels in flight should stay a
    // compact region that L2 can hold (measured best at 1024^3: 128); short lists
    // keep every SM busy instead
    int chunk = 32 * (int)std::min<int64_t>(4, std::max<int64_t>(1, n / (32 * 8192)));
    if (getenv("BDR_TRACE_CHUNK")) chunk = std::min(chunk, atoi(getenv("BDR_TRACE_CHUNK")));
    const int64_t n_warps = (n + chunk - 1) / chunk;
    const PeerView *pv = static_cast<const PeerView *>(c->peer_view);
    int32_t *chg = want_changed_list ? c->list2 : (int32_t *)nullptr;
    int32_t *term = c->use_term ? c->term : (int32_t *)nullptr;
    constexpr int SLOW_CAP = 4096;
    const int64_t batch = 16384;
    // stage 1: the local kernel.  On a slab window it may only read classification and
    // labels on the planes this rank owns; walks that step off them (and over-long ones)
    // come back in c->list3
    Window win = window_of(c);
    const bool local_first = pv && c->halo >= 3 && !getenv("BDR_TRACE_PEER_ONLY");
    if (local_first) {
        win.xlo = c->halo;
        win.xhi = c->g.nx - c->halo - 1;
        // an exact pass follows a full edge pass over the whole window on freshly exchanged
        // labels: the classification of planes [2, W-3] equals the owners', and the labels read
        // there belong to interior voxels or maxima, which no rank writes during the pass
        // (the conservative first pass of bader_calc('neargrid') qualifies as well: trajectory
        // ends off the owned planes are simply not cached, see k_trace)
        if (c->window_fresh) {
            win.xlo = 2;
            win.xhi = c->g.nx - 3;
        }
        // the library's own round loops copy the owners' known planes into the halos after
        // every classification step: every plane whose stencil fits the window is trusted
        if (c->halo_known_current) {
            win.xlo = 1;
            win.xhi = c->g.nx - 2;
        }
    }
    // scratch of the SLOW variant: a path buffer in global memory, in batches so that it
    // stays bounded (SLOW_CAP entries per walk)
    auto slow_pass = [&](const int32_t *in, int64_t ov, int32_t *esc_list, int64_t esc_cap) -> int {
        TRY(ensure_stage(c, (size_t)(std::min(ov, batch) + 128) * SLOW_CAP * sizeof(long long)));
        for (int64_t o = 0; o < ov; o += batch) {
            const int64_t m = std::min(batch, ov - o);
            LAUNCH(c, BDR_K_TRACE, (k_trace<SLOW_CAP, true>), blocks_for(m, 128), 128, 0,
                   rho_ptr(c, BDR_RHO_REFERENCE), c->labels[which], c->known, c->g, win, W, T, in + o, m, 32,
                   (int32_t *)c->stage, c->d_cnt, chg, c->list2_cap, (int32_t *)nullptr, (int64_t)0, step_cap,
                   term, esc_list, esc_cap, c->halo_known_current ? 1 : 0);
        }
        TRY(read_counters(c));
        if (c->h_cnt[CNT_ERROR]) return fail_msg("trace: trajectory longer than 4096 voxels");
        return 0;
    };
    TRY(zero_counter(c, CNT_ESCLIST));
    if (!pv || local_first) {
        // walks that leave the trusted planes are listed for the peer kernel (slab windows),
        // walks that outgrow the register-file path buffer for the SLOW variant
        int32_t *esc_list = nullptr;
        if (local_first) {
            TRY(ensure(&c->list4, &c->list4_cap, n));
            esc_list = c->list4;
        }
        LAUNCH(c, BDR_K_TRACE, (k_trace<PATH_FAST, false>), blocks_for(n_warps * 32, 128), 128, 0,
               rho_ptr(c, BDR_RHO_REFERENCE), c->labels[which], c->known, c->g, win, W, T, c->list, n,
               chunk, (int32_t *)nullptr, c->d_cnt, chg, c->list2_cap, c->list3, c->list3_cap, step_cap,
               term, esc_list, c->list4_cap, c->halo_known_current ? 1 : 0);
        TRY(read_counters(c));
        if (c->h_cnt[CNT_ERROR]) return fail_msg("trace: trajectory exceeded the step cap");
        const int64_t ov = (int64_t)c->h_cnt[CNT_OVERFLOW];
        if (ov > 0) {
            // long paths stay local as long as they stay on trusted planes
            CU(cudaMemsetAsync(c->d_cnt + CNT_OVERFLOW, 0, sizeof(unsigned long long) * 2, c->stream));
            TRY(slow_pass(c->list3, ov, esc_list, c->list4_cap));
        }
    }
    if (pv) {
        // stage 2 (slab windows): the peer kernel continues on the neighbours' memory
        // (K4p) -- over the walks stage 1 handed back, or over the whole list
        const int32_t *in = c->list;
        int64_t m_in = n;
        int32_t *pov = c->list4;            // over-long walks of the peer kernel
        if (local_first) {
            in = c->list4;
            m_in = (int64_t)c->h_cnt[CNT_ESCLIST];
            pov = c->list3;                 // free again: the local SLOW pass is done
        } else {
            TRY(ensure(&c->list4, &c->list4_cap, m_in));
            pov = c->list4;
        }
        if (m_in > 0) {
            const int64_t pov_cap = local_first ? c->list3_cap : c->list4_cap;
            CU(cudaMemsetAsync(c->d_cnt + CNT_OVERFLOW, 0, sizeof(unsigned long long) * 2, c->stream));
            // (the escape list of a big pass is long enough for lanes to refill from a chunk)
            const int pchunk = local_first ? (m_in > (1 << 20) ? 128 : (m_in > (1 << 17) ? 64 : 32)) : chunk;
            PeerView pvl = *pv;   // what the local kernel trusted is read locally here as well
            pvl.tlo = local_first ? win.xlo : c->halo;
            pvl.thi = local_first ? win.xhi : c->g.nx - c->halo - 1;
            pvl.clo = c->halo_known_current ? pvl.tlo : c->halo;
            pvl.chi = c->halo_known_current ? pvl.thi : c->g.nx - c->halo - 1;
            pv = &pvl;
            LAUNCH(c, BDR_K_TRACE_PEER, (k_trace_peer<PATH_FAST, false>),
                   blocks_for((m_in + pchunk - 1) / pchunk * 32, 128), 128, 0, *pv, c->labels[which],
                   c->known, c->g, window_of(c), W, T, in, m_in, pchunk, (long long *)nullptr, c->d_cnt,
                   chg, c->list2_cap, pov, pov_cap, step_cap, term);
            TRY(read_counters(c));
            if (c->h_cnt[CNT_ERROR]) return fail_msg("trace: trajectory exceeded the step cap");
            const int64_t ov = (int64_t)c->h_cnt[CNT_OVERFLOW];
            if (ov > 0) {
                TRY(ensure_stage(c, (size_t)(std::min(ov, batch) + 128) * SLOW_CAP * sizeof(long long)));
                CU(cudaMemsetAsync(c->d_cnt + CNT_OVERFLOW, 0, sizeof(unsigned long long) * 2, c->stream));
                for (int64_t o = 0; o < ov; o += batch) {
                    const int64_t m = std::min(batch, ov - o);
                    LAUNCH(c, BDR_K_TRACE_PEER, (k_trace_peer<SLOW_CAP, true>), blocks_for(m, 128), 128, 0, *pv,
                           c->labels[which], c->known, c->g, window_of(c), W, T, pov + o, m, 32,
                           (long long *)c->stage, c->d_cnt, chg, c->list2_cap, (int32_t *)nullptr,
                           (int64_t)0, step_cap, term);
                }
                TRY(read_counters(c));
                if (c->h_cnt[CNT_ERROR]) return fail_msg("trace: trajectory longer than 4096 voxels");
            }
        }
    }
    *changed = (int64_t)c->h_cnt[CNT_CHANGED];
    TRY(eq_note_changed(c, which, *changed, want_changed_list));
    c->escaped = (int64_t)c->h_cnt[CNT_ESCAPED];
    if (c->escaped && c->halo == 0) return fail_msg("trace: internal error (escape on a periodic grid)");
    c->trace_steps += (int64_t)c->h_cnt[CNT_STEPS];
    c->trace_voxels += n;
    return 0;
}

// refinement.edge_check restated (see kernels.cuh K5).  Input: c->list2 holds
// the voxels changed by the last trace (known == -2).  Output: c->list holds
// the voxels to trace next (known == -2).  Three phases, so that the ranks of a
// sharded run can exchange the known planes between them (bdr_slab_ec_*):
//   begin   class-2 changed voxels are centres at once
//   round   one step of the centre selection (repeat until nothing is undecided)
//   finish  re-classify the 27-neighbourhoods of the centres, dilate, queue
static int ec_begin_dev(bdr_ctx *c, int which, int64_t n_changed) {
    if (n_changed == 0) return 0;
    LAUNCH(c, BDR_K_EDGE_CHECK, k_ec_init, blocks_for(n_changed, 128), 128, 0,
           rho_ptr(c, BDR_RHO_REFERENCE), c->labels[which], c->known, c->g, c->list2, n_changed);
    return 0;
}
static int ec_round_dev(bdr_ctx *c, int64_t n_changed, int64_t *undecided) {
    *undecided = 0;
    if (n_changed == 0) return 0;
    const PeerView *pv = static_cast<const PeerView *>(c->peer_view);
    TRY(zero_counter(c, CNT_UNDECIDED));
    LAUNCH(c, BDR_K_EDGE_CHECK, k_ec_round, blocks_for(n_changed, 128), 128, 0, c->known, c->g,
           c->list2, n_changed, c->d_cnt + CNT_UNDECIDED, pv ? pv->x0w : 0, pv ? pv->NX : c->g.nx);
    TRY(read_counters(c));
    *undecided = (int64_t)c->h_cnt[CNT_UNDECIDED];
    return 0;
}
static int ec_finish_dev(bdr_ctx *c, int which, int64_t n_changed, int64_t *edges) {
    *edges = 0;
    c->list_n = 0;
    c->window_fresh = false;   // only the neighbourhoods of the centres were re-classified
    c->halo_known_current = false;
    const double *rho = rho_ptr(c, BDR_RHO_REFERENCE);
    const int32_t *lab = c->labels[which];
    const int64_t plane = (int64_t)c->g.ny * c->g.nz;
    const int64_t halo_cells = c->halo > 0 ? 2 * (int64_t)(c->halo - 1) * plane : 0;
    if (n_changed == 0 && halo_cells == 0) return 0;
    TRY(ensure(&c->list3, &c->list3_cap, n_changed + halo_cells + 1));
    TRY(zero_counter(c, CNT_CENTRES));
    if (n_changed > 0)
        LAUNCH(c, BDR_K_EDGE_CHECK, k_ec_collect_centres, blocks_for(n_changed, 128), 128, 0, c->known,
               c->list2, n_changed, c->d_cnt + CNT_CENTRES, c->list3);
    if (halo_cells > 0)
        LAUNCH(c, BDR_K_EDGE_CHECK, k_ec_collect_halo, blocks_for(halo_cells, 256), 256, 0, c->known,
               (int)plane, c->g.nx, c->halo, c->d_cnt + CNT_CENTRES, c->list3, c->list3_cap);
    TRY(read_counters(c));
    const int64_t nc = (int64_t)c->h_cnt[CNT_CENTRES];
    if (nc == 0) return 0;
    TRY(ensure(&c->list, &c->list_cap, nc * 27 + nc));
    TRY(zero_counter(c, CNT_NEWEDGE));
    LAUNCH(c, BDR_K_EDGE_CHECK, k_ec_classify, blocks_for(nc * 27, 128), 128, 0, rho, lab, c->known,
           c->g, c->list3, nc, c->d_cnt + CNT_NEWEDGE, c->list, c->list_cap);
    TRY(read_counters(c));
    const int64_t ne = (int64_t)c->h_cnt[CNT_NEWEDGE];
    if (ne > 0)
        LAUNCH(c, BDR_K_EDGE_CHECK, k_ec_dilate, blocks_for(ne * 27, 128), 128, 0, c->known, c->g,
               c->list, ne);
    TRY(zero_counter(c, CNT_EDGES));
    LAUNCH(c, BDR_K_EDGE_CHECK, k_ec_finish, blocks_for(ne + nc, 128), 128, 0, c->known, c->list, ne,
           c->list3, nc, c->d_cnt + CNT_NEWEDGE, c->list_cap, (int)c->own_lo, (int)c->own_hi,
           c->d_cnt + CNT_EDGES);
    TRY(read_counters(c));
    c->list_n = (int64_t)c->h_cnt[CNT_NEWEDGE];
    *edges = (int64_t)c->h_cnt[CNT_EDGES];
    return 0;
}
static int edge_check_dev(bdr_ctx *c, int which, int64_t n_changed, int64_t *edges) {
    *edges = 0;
    c->list_n = 0;
    if (n_changed == 0) return 0;
    TRY(ec_begin_dev(c, which, n_changed));
    for (int round = 0;; ++round) {
        int64_t undecided = 0;
        TRY(ec_round_dev(c, n_changed, &undecided));
        if (undecided == 0) break;
        if (round > 1 << 20) return fail_msg("edge_check: centre selection did not converge");
    }
    return ec_finish_dev(c, which, n_changed, edges);
}

// thread_handlers.refine (thread_handlers.py:144-236)
static int refine_dev(bdr_ctx *c, int which, int mode, int64_t iters, const Weights &W,
                      const TGrad &T, int64_t *iters_run, int64_t *history, int64_t hist_cap) {
    int64_t run = 0;
    auto record = [&](int64_t e, int64_t ch) {
        if (history && run < hist_cap) {
            history[2 * run] = e;
            history[2 * run + 1] = ch;
        }
        ++run;
    };
    if (iters_run) *iters_run = 0;
    if (iters == 0) return 0;
    int64_t edges = 0, changed = 0;
    TRY(edge_find_dev(c, which, &edges));
    if (edges == 0) return 0;
    const bool chg_mode = (mode == BDR_MODE_CHANGED);
    TRY(trace_dev(c, which, W, T, &changed, chg_mode));
    record(edges, changed);
    for (int64_t it = 2; iters < 0 || it <= iters; ++it) {
        if (chg_mode) {
            TRY(edge_check_dev(c, which, changed, &edges));
        } else {
            // a fresh edge_find on unchanged labels reproduces the previous
            // iteration exactly, so it would again change nothing
            if (changed == 0) {
                record(edges, 0);
                break;
            }
            TRY(edge_find_dev(c, which, &edges));
        }
        TRY(trace_dev(c, which, W, T, &changed, chg_mode));
        record(edges, changed);
        if (changed == 0) break;
    }
    if (iters_run) *iters_run = run;
    return 0;
}

// bader_calc('neargrid'): drive labels to the fixed point of the reference's
// order-free refinement iteration ("every edge voxel carries the label its own
// trajectory ends in").  One full edge pass + trace, then cheap incremental
// rounds around the voxels that changed, then a full pass to confirm; repeat
// until a full pass changes nothing.
// drop queue entries whose cached trajectory end is still interior (c->list -> c->list)
static int filter_cached_dev(bdr_ctx *c) {
    const int64_t n = c->list_n;
    if (n == 0) return 0;
    TRY(ensure(&c->list3, &c->list3_cap, n));
    TRY(zero_counter(c, CNT_CENTRES));
    LAUNCH(c, BDR_K_EDGE_CHECK, k_filter_cached, blocks_for(n, 256), 256, 0, c->list, n, c->term,
           c->known, c->d_cnt + CNT_CENTRES, c->list3);
    TRY(read_counters(c));
    std::swap(c->list, c->list3);
    std::swap(c->list_cap, c->list3_cap);
    c->list_n = (int64_t)c->h_cnt[CNT_CENTRES];
    return 0;
}

static int converge_rounds(bdr_ctx *c, int which, const Weights &W, const TGrad &T, bool dbg,
                           int outer) {
    int64_t edges = 0, changed = 0;
    // conservative full pass: starts the sticky "never interior again" bits
    TRY(edge_find_dev(c, which, &edges, -1, 1));
    if (edges == 0) return 0;
    CU(cudaMemsetAsync(c->term, 0xff, (size_t)c->N * sizeof(int32_t), c->stream));
    TRY(trace_dev(c, which, W, T, &changed, true));
    if (dbg) fprintf(stderr, "[bdr] full pass %d: edges %lld changed %lld\n", outer, (long long)edges,
                     (long long)changed);
    for (int inner = 0; inner < 4096 && changed > 0; ++inner) {
        int64_t queued = 0;
        if (changed * 2048 > c->N) {
            // many voxels moved: a full (streaming) edge pass that lists only
            // the edges next to them beats gathering their neighbourhoods
            TRY(edge_find_dev(c, which, &queued, changed, 2));
        } else {
            TRY(incremental_dev(c, which, changed, &queued));
        }
        queued = c->list_n;
        TRY(filter_cached_dev(c));
        const int64_t kept = c->list_n;
        TRY(trace_dev(c, which, W, T, &changed, true));
        if (dbg) fprintf(stderr, "[bdr]   incremental %d: queued %lld traced %lld changed %lld\n", inner,
                         (long long)queued, (long long)kept, (long long)changed);
    }
    return changed == 0 ? 0 : fail_msg("bader_calc(neargrid): incremental rounds did not settle");
}

// bader_calc('neargrid'): drive labels to the fixed point of the reference's
// order-free refinement iteration ("every edge voxel carries the label its own
// trajectory ends in").  One full edge pass + trace, then rounds around the
// voxels that changed (a masked streaming pass while they are many, list-based
// gathers when they are few).  Within these rounds the interior set only
// shrinks, so a voxel whose recorded trajectory end is still interior is not
// traced again.  With BDR_OPT_VERIFY_FIXED_POINT the whole thing repeats until
// an exact full pass changes nothing.
static int converge_dev(bdr_ctx *c, int which, const Weights &W, const TGrad &T) {
    const bool dbg = getenv("BDR_DEBUG") != nullptr;
    if (!c->term) CU(cudaMalloc((void **)&c->term, (size_t)c->N * sizeof(int32_t)));
    TRY(ensure_known(c));
    TRY(ensure_bits(c));
    c->use_term = true;
    int rc = converge_rounds(c, which, W, T, dbg, 0);
    c->use_term = false;
    if (rc || !c->verify_fixed_point) return rc;
    for (int outer = 1; outer < 64; ++outer) {
        // exact full pass (what the caller's refine() would run): if it moves
        // nothing the labels are a fixed point of the reference iteration
        int64_t edges = 0, changed = 0;
        TRY(edge_find_dev(c, which, &edges));
        if (edges == 0) return 0;
        TRY(trace_dev(c, which, W, T, &changed, true));
        if (dbg) fprintf(stderr, "[bdr] verify pass %d: edges %lld changed %lld\n", outer,
                         (long long)edges, (long long)changed);
        if (changed == 0) return 0;
        c->use_term = true;
        rc = converge_rounds(c, which, W, T, dbg, outer);
        c->use_term = false;
        if (rc) return rc;
    }
    return fail_msg("bader_calc(neargrid): refinement did not reach a fixed point");
}

template <typename T>
static int download_cast(bdr_ctx *c, const int32_t *src, void *host) {
    TRY(ensure_stage(c, (size_t)c->N * sizeof(T)));
    LAUNCH(c, BDR_K_NARROW, k_narrow<T>, blocks_for(c->N, 256), 256, 0, src, (T *)c->stage, c->N);
    CU(cudaMemcpyAsync(host, c->stage, (size_t)c->N * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
template <typename T>
static int upload_cast(bdr_ctx *c, int32_t *dst, const void *host) {
    TRY(ensure_stage(c, (size_t)c->N * sizeof(T)));
    CU(cudaMemcpyAsync(c->stage, host, (size_t)c->N * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, BDR_K_NARROW, k_widen<T>, blocks_for(c->N, 256), 256, 0, (const T *)c->stage, dst,
           c->N);
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

static int charge_sum_dev(bdr_ctx *c, int which_labels, int which_density, double dV, int64_t n,
                          double *charge, double *volume) {
    if (!c->labels[which_labels]) return fail_msg("charge_sum: label set is empty");
    const double *dens = rho_ptr(c, which_density);
    if (!dens) return fail_msg("charge_sum: density slot is empty");
    if (n <= 0) return 0;
    TRY(ensure_sums(c, 2 * n));
    CU(cudaMemsetAsync(c->d_sums, 0, (size_t)(2 * n) * sizeof(double), c->stream));
    double *q = c->d_sums;
    unsigned long long *cnt = reinterpret_cast<unsigned long long *>(c->d_sums + n);
    const int64_t per_block = 256 * 64;
    const int64_t n_own = c->own_hi - c->own_lo;  // a slab sums its own voxels only
    const unsigned nb = blocks_for(n_own, (int)per_block);
    const int32_t *lab = c->labels[which_labels] + c->own_lo;
    dens += c->own_lo;
    TRY(zero_counter(c, CNT_ERROR));
    if (n <= SUM_BINS)
        LAUNCH(c, BDR_K_CHARGE_SUM, k_charge_sum<true>, nb, 256, 0, dens, lab, n_own, (int)n, q, cnt,
               per_block, c->d_cnt + CNT_ERROR);
    else
        LAUNCH(c, BDR_K_CHARGE_SUM, k_charge_sum<false>, nb, 256, 0, dens, lab, n_own, (int)n, q, cnt,
               per_block, c->d_cnt + CNT_ERROR);
    std::vector<double> hq((size_t)n);
    std::vector<unsigned long long> hc((size_t)n);
    CU(cudaMemcpyAsync(hq.data(), q, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(hc.data(), cnt, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    TRY(read_counters(c));
    if (c->h_cnt[CNT_ERROR]) {
        const unsigned long long bad = c->h_cnt[CNT_ERROR];
        TRY(zero_counter(c, CNT_ERROR));
        return fail_msg("charge_sum: " + std::to_string(bad) + " voxels carry a label >= " + std::to_string(n) +
                        " (the charge / volume arrays have " + std::to_string(n) + " entries)");
    }
    for (int64_t i = 0; i < n; ++i) {
        // the reference adds into caller-provided (zeroed) arrays, then scales
        if (charge) charge[i] = (charge[i] + hq[(size_t)i]) * dV;
        if (volume) volume[i] += (double)hc[(size_t)i] * dV;
    }
    return 0;
}

}  // namespace bdr

using namespace bdr;

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char *bdr_last_error(void) { return g_err.c_str(); }
int bdr_version(void) { return 100; }

int bdr_device_count(int *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail("cudaGetDeviceCount", __FILE__, __LINE__, e);
    }
    *count = n;
    return 0;
}

int bdr_create(int device, int64_t nx, int64_t ny, int64_t nz, bdr_ctx **out) {
    if (!out) return fail_msg("bdr_create: null out");
    *out = nullptr;
    if (nx < 1 || ny < 1 || nz < 1) return fail_msg("bdr_create: empty grid");
    const int64_t N = nx * ny * nz;
    if (N > (int64_t)2147483647 - 65536)
        return fail_msg("bdr_create: grid too large for one device handle (N must be < 2^31 - 2^16); shard it");
    int ndev = 0;
    TRY(bdr_device_count(&ndev));
    if (ndev == 0) return fail_msg("bdr_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail_msg("bdr_create: bad device index");
    CU(cudaSetDevice(device));
    bdr_ctx *c = new bdr_ctx();
    c->device = device;
    c->g = make_grid((int)nx, (int)ny, (int)nz);
    c->N = N;
    c->own_lo = 0;
    c->own_hi = N;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaMalloc((void **)&c->d_cnt, sizeof(unsigned long long) * CNT_NUM));
    CU(cudaMemsetAsync(c->d_cnt, 0, sizeof(unsigned long long) * CNT_NUM, c->stream));
    CU(cudaMallocHost((void **)&c->h_cnt, sizeof(unsigned long long) * CNT_NUM));
    CU(cudaFuncSetAttribute(k_ongrid_pointers<SX, TY, TZ, VAC_NONE>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil_smem()));
    CU(cudaFuncSetAttribute(k_ongrid_pointers<SX, TY, TZ, VAC_TOL>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil_smem()));
    CU(cudaFuncSetAttribute(k_ongrid_pointers<SX, TY, TZ, VAC_LABELS>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil_smem()));
    CU(cudaFuncSetAttribute(k_seed_pointers<VAC_NONE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem()));
    CU(cudaFuncSetAttribute(k_seed_pointers<VAC_TOL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem()));
    CU(cudaFuncSetAttribute(k_seed_pointers<VAC_LABELS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem()));
    CU(cudaFuncSetAttribute(k_seed_pointers<VAC_NONE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem()));
    CU(cudaFuncSetAttribute(k_seed_pointers<VAC_TOL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem()));
    CU(cudaFuncSetAttribute(k_seed_pointers<VAC_LABELS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem()));
    *out = c;
    return 0;
}

int bdr_slab_create(int device, int64_t nx_window, int64_t ny, int64_t nz, int halo, bdr_ctx **out) {
    if (halo < 3) return fail_msg("bdr_slab_create: halo must be at least 3 planes");
    if (nx_window <= 2 * (int64_t)halo) return fail_msg("bdr_slab_create: window thinner than its halos");
    TRY(bdr_create(device, nx_window, ny, nz, out));
    bdr_ctx *c = *out;
    c->halo = halo;
    c->own_lo = (int64_t)halo * ny * nz;
    c->own_hi = (nx_window - halo) * ny * nz;
    // the arrays other ranks map (bdr_slab_ipc_export) exist from the start
    TRY(ensure_rho(c, BDR_RHO_REFERENCE));
    TRY(ensure_labels(c, BDR_LABELS_BADER));
    TRY(ensure_known(c));
    CU(cudaMemsetAsync(c->known, 0, (size_t)c->N, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_slab_ipc_export(bdr_ctx *c, void *handles) {
    TRY(check(c));
    if (c->halo == 0) return fail_msg("bdr_slab_ipc_export: not a slab handle");
    cudaIpcMemHandle_t *h = static_cast<cudaIpcMemHandle_t *>(handles);
    CU(cudaIpcGetMemHandle(&h[0], c->rho[BDR_RHO_REFERENCE]));
    CU(cudaIpcGetMemHandle(&h[1], c->labels[BDR_LABELS_BADER]));
    CU(cudaIpcGetMemHandle(&h[2], c->known));
    return 0;
}

int bdr_slab_ipc_attach(bdr_ctx *c, int world, int rank, const void *all_handles,
                        const int64_t *bounds, int64_t nx_global) {
    TRY(check(c));
    if (c->halo == 0) return fail_msg("bdr_slab_ipc_attach: not a slab handle");
    if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world)
        return fail_msg("bdr_slab_ipc_attach: bad world / rank");
    PeerView *pv = new PeerView();
    memset(pv, 0, sizeof *pv);
    const cudaIpcMemHandle_t *h = static_cast<const cudaIpcMemHandle_t *>(all_handles);
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            pv->rho[r] = c->rho[BDR_RHO_REFERENCE];
            pv->lab[r] = c->labels[BDR_LABELS_BADER];
            pv->known[r] = c->known;
            continue;
        }
        void *p[3];
        for (int k = 0; k < 3; ++k) {
            cudaIpcMemHandle_t hk = h[r * 3 + k];
            CU(cudaIpcOpenMemHandle(&p[k], hk, cudaIpcMemLazyEnablePeerAccess));
            c->ipc_opened.push_back(p[k]);
        }
        pv->rho[r] = static_cast<const double *>(p[0]);
        pv->lab[r] = static_cast<const int32_t *>(p[1]);
        pv->known[r] = static_cast<const int8_t *>(p[2]);
    }
    for (int r = 0; r <= world; ++r) pv->bound[r] = (int)bounds[r];
    pv->world = world;
    pv->rank = rank;
    pv->halo = c->halo;
    pv->NX = (int)nx_global;
    pv->x0w = (int)bounds[rank] - c->halo;
    pv->W = c->g.nx;
    if (bounds[rank + 1] - bounds[rank] + 2 * c->halo != c->g.nx)
        return fail_msg("bdr_slab_ipc_attach: slab bounds do not match the window");
    c->peer_view = pv;
    return 0;
}

int bdr_slab_seed(bdr_ctx *c, const double *dist_mat, int64_t *n_real, int64_t *exit_base) {
    TRY(check(c));
    if (c->halo == 0) return fail_msg("bdr_slab_seed: not a slab handle");
    if (!rho_ptr(c, BDR_RHO_REFERENCE)) return fail_msg("bdr_slab_seed: density not set");
    const Weights W = make_weights(dist_mat);
    int64_t n = 0;
    const int vac_mode_at_entry = c->vac_mode;
    c->maxima_fresh[0] = c->maxima_fresh[1] = false;
    eq_invalidate(c, BDR_LABELS_BADER);
    TRY(choose_seed(c, c->slab_seed_method, W));
    TRY(stencil_dev(c, W, &n));
    TRY(resolve_dev(c, nullptr));
    CU(cudaStreamSynchronize(c->stream));
    c->n_max = n;
    // c->roots now lists every maximum of the window's density off the two exit planes
    // (whose classification no pass trusts): the exact edge passes test only those
    c->maxima_fresh[BDR_LABELS_BADER] = vac_mode_at_entry != VAC_LABELS;
    if (n_real) *n_real = n;
    if (exit_base) *exit_base = exit_base_of(c);
    return 0;
}

int bdr_slab_first_voxel(bdr_ctx *c, int64_t n_slots, int32_t *dev_out) {
    TRY(check(c));
    CU(cudaMemsetAsync(dev_out, 0x7f, (size_t)n_slots * sizeof(int32_t), c->stream));
    LAUNCH(c, BDR_K_FIRST, k_first_voxel_slots, blocks_for(c->own_hi - c->own_lo, 1024), 256, 0,
           c->labels[BDR_LABELS_BADER], (int)c->own_lo, (int)c->own_hi, dev_out);
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_slab_apply_rank(bdr_ctx *c, const int32_t *dev_rank) {
    TRY(check(c));
    eq_invalidate(c, BDR_LABELS_BADER);   // exit slots and maxima slots may share a volume number
    LAUNCH(c, BDR_K_RELABEL, k_relabel_slots, blocks_for(c->N, 1024), 256, 0,
           c->labels[BDR_LABELS_BADER], c->N, dev_rank);
    c->vac_mode = VAC_LABELS;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_slab_first_voxel_labels(bdr_ctx *c, int64_t n_labels, int32_t *dev_out) {
    TRY(check(c));
    if (n_labels <= 0) return 0;
    CU(cudaMemsetAsync(dev_out, 0x7f, (size_t)n_labels * sizeof(int32_t), c->stream));
    LAUNCH(c, BDR_K_FIRST, k_first_voxel_labels, blocks_for(c->own_hi - c->own_lo, 1024), 256, 0,
           c->labels[BDR_LABELS_BADER], (int)c->own_lo, (int)c->own_hi, dev_out, (int)n_labels);
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_slab_relabel(bdr_ctx *c, int which, const int32_t *dev_lut) {
    TRY(check(c));
    if (which < 0 || which > 1 || !c->labels[which]) return fail_msg("bdr_slab_relabel: bad label set");
    LAUNCH(c, BDR_K_RELABEL, k_relabel_lut, blocks_for(c->N, 1024), 256, 0, c->labels[which],
           c->labels[which], c->N, dev_lut);
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_slab_roots(bdr_ctx *c, int32_t *host_out, int64_t cap) {
    TRY(check(c));
    if (cap < c->n_max) return fail_msg("bdr_slab_roots: buffer too small");
    if (c->n_max)
        CU(cudaMemcpy(host_out, c->roots, (size_t)c->n_max * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

int bdr_edge_pass(bdr_ctx *c, int which, int64_t *edges) { return bdr_edge_find(c, which, edges); }

// sharded bader_calc('neargrid'): the same conservative rounds as converge_dev,
// cut where the ranks have to meet (sharded.py drives them)
int bdr_slab_first_pass(bdr_ctx *c, int which, int64_t *edges) {
    TRY(check(c));
    if (which < 0 || which > 1 || !c->labels[which]) return fail_msg("bdr_slab_first_pass: bad label set");
    int64_t e = 0;
    TRY(edge_find_dev(c, which, &e, -1, 1));
    c->last_changed = 0;
    // trajectory-end cache for the rounds that follow (see converge_dev)
    if (!c->term) CU(cudaMalloc((void **)&c->term, (size_t)c->N * sizeof(int32_t)));
    CU(cudaMemsetAsync(c->term, 0xff, (size_t)c->N * sizeof(int32_t), c->stream));
    c->use_term = true;
    CU(cudaStreamSynchronize(c->stream));
    if (edges) *edges = e;
    return 0;
}

int bdr_slab_trace(bdr_ctx *c, int which, const double *dist_mat, const double *T_grad,
                   int want_list, int64_t *changed) {
    TRY(check(c));
    if (which < 0 || which > 1 || !c->labels[which]) return fail_msg("bdr_slab_trace: bad label set");
    if (!c->known) return fail_msg("bdr_slab_trace: no edge pass has run");
    const Weights W = make_weights(dist_mat);
    const TGrad T = make_tgrad(T_grad);
    int64_t ch = 0;
    TRY(trace_dev(c, which, W, T, &ch, want_list != 0));
    c->last_changed = want_list ? ch : 0;
    CU(cudaStreamSynchronize(c->stream));
    if (changed) *changed = ch;
    return 0;
}

int bdr_slab_requeue(bdr_ctx *c, int which, const int32_t *dev_extra, int64_t n_extra,
                     int64_t *queued) {
    TRY(check(c));
    if (which < 0 || which > 1 || !c->labels[which]) return fail_msg("bdr_slab_requeue: bad label set");
    const int64_t n = c->last_changed + n_extra;
    if (n_extra > 0) {
        // voxels a neighbour relabelled on the planes next to this slab count
        // as changed here too: their 27-neighbourhoods reach owned voxels
        if (c->last_changed + n_extra > c->list2_cap) {
            int32_t *grown = nullptr;
            const int64_t cap = n + n / 4 + 1024;
            CU(cudaMalloc((void **)&grown, (size_t)cap * sizeof(int32_t)));
            if (c->last_changed)
                CU(cudaMemcpyAsync(grown, c->list2, (size_t)c->last_changed * sizeof(int32_t),
                                   cudaMemcpyDeviceToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            if (c->list2) cudaFree(c->list2);
            c->list2 = grown;
            c->list2_cap = cap;
        }
        CU(cudaMemcpyAsync(c->list2 + c->last_changed, dev_extra, (size_t)n_extra * sizeof(int32_t),
                           cudaMemcpyDeviceToDevice, c->stream));
    }
    int64_t q = 0;
    if (n * 2048 > c->N) {
        TRY(edge_find_dev(c, which, &q, n, 2));
    } else {
        TRY(incremental_dev(c, which, n, &q));
    }
    if (c->use_term) TRY(filter_cached_dev(c));
    q = c->list_n;
    CU(cudaStreamSynchronize(c->stream));
    if (queued) *queued = q;
    return 0;
}

// 'changed'-mode refinement across slabs (thread_handlers.py:201-205 ->
// refinement.edge_check): the phases of edge_check_dev, cut where the ranks exchange the
// halo planes of the known array
int bdr_slab_ec_begin(bdr_ctx *c, int which) {
    TRY(check(c));
    if (which < 0 || which > 1 || !c->labels[which] || !c->known) return fail_msg("bdr_slab_ec_begin: bad state");
    TRY(ec_begin_dev(c, which, c->last_changed));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
int bdr_slab_ec_round(bdr_ctx *c, int64_t *undecided) {
    TRY(check(c));
    int64_t u = 0;
    TRY(ec_round_dev(c, c->last_changed, &u));
    if (undecided) *undecided = u;
    return 0;
}
int bdr_slab_ec_finish(bdr_ctx *c, int which, int64_t *edges_owned) {
    TRY(check(c));
    if (which < 0 || which > 1 || !c->labels[which]) return fail_msg("bdr_slab_ec_finish: bad label set");
    int64_t e = 0;
    TRY(ec_finish_dev(c, which, c->last_changed, &e));
    CU(cudaStreamSynchronize(c->stream));
    if (edges_owned) *edges_owned = e;
    return 0;
}

// ---- the sharded protocol's round loops, driven from the library (comm.cuh) --------
int bdr_slab_comm_id(void *id_out) {
    NcclApi *api = nccl_api();
    if (!api) return fail_msg("bdr_slab_comm_id: libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    NC(api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return 0;
}

int bdr_slab_comm_init(bdr_ctx *c, int world, int rank, const void *id_bytes) {
    TRY(check(c));
    if (c->halo == 0) return fail_msg("bdr_slab_comm_init: not a slab handle");
    if (world < 1 || rank < 0 || rank >= world) return fail_msg("bdr_slab_comm_init: bad world / rank");
    if (c->slab_comm) return fail_msg("bdr_slab_comm_init: already initialised");
    SlabComm *sc = new SlabComm();
    sc->world = world;
    sc->rank = rank;
    sc->prev = (rank + world - 1) % world;
    sc->next = (rank + 1) % world;
    if (world > 1) {
        NcclApi *api = nccl_api();
        if (!api) {
            delete sc;
            return fail_msg("bdr_slab_comm_init: libnccl.so.2 could not be loaded");
        }
        ncclUniqueId id;
        memcpy(&id, id_bytes, sizeof id);
        ncclResult_t r = api->CommInitRank(&sc->comm, world, id, rank);
        if (r != ncclSuccess) {
            delete sc;
            return fail_msg(std::string("ncclCommInitRank failed: ") + api->GetErrorString(r));
        }
    }
    const int64_t plane = (int64_t)c->g.ny * c->g.nz;
    CU(cudaMalloc((void **)&sc->d_red, 8 * sizeof(unsigned long long)));
    CU(cudaMallocHost((void **)&sc->h_red, 8 * sizeof(unsigned long long)));
    CU(cudaMalloc((void **)&sc->plane_lo, (size_t)plane * sizeof(int32_t)));
    CU(cudaMalloc((void **)&sc->plane_hi, (size_t)plane * sizeof(int32_t)));
    c->slab_comm = sc;
    return 0;
}

int bdr_slab_exchange(bdr_ctx *c, int what) {
    TRY(check(c));
    SlabComm *sc = static_cast<SlabComm *>(c->slab_comm);
    if (!sc) return fail_msg("bdr_slab_exchange: bdr_slab_comm_init has not run");
    if (what == 0 || what == 1) {
        if (!c->labels[what]) return fail_msg("bdr_slab_exchange: label set is empty");
        return comm_halo_exchange(c, sc, c->labels[what], 4);
    }
    if (what == 2) {
        if (!c->known) return fail_msg("bdr_slab_exchange: no edge pass has run");
        return comm_halo_exchange(c, sc, c->known, 1);
    }
    return fail_msg("bdr_slab_exchange: bad selector");
}

// every rank's classification of its own planes goes to the neighbours' halos: walks that
// leave the slab by less than `halo` planes then run on local copies of exactly the bytes
// the owner holds, and only deeper ones take the peer kernel's remote loads
static int slab_publish_known(bdr_ctx *c, SlabComm *sc) {
    if (getenv("BDR_NO_KNOWN_EXCHANGE")) return 0;
    TRY(comm_halo_exchange(c, sc, c->known, 1));
    c->halo_known_current = true;
    return 0;
}

// relabelled voxels of the last trace (c->list2) within 2*halo planes of either end of the
// owned slab: zero on every rank means no halo copy anywhere went stale
static int slab_zone_changed(bdr_ctx *c, int64_t n_changed, int64_t *count) {
    *count = 0;
    if (n_changed == 0) return 0;
    const int64_t plane = (int64_t)c->g.ny * c->g.nz;
    const int64_t zone = 2 * (int64_t)c->halo;
    TRY(zero_counter(c, CNT_CENTRES));
    LAUNCH(c, BDR_K_EDGE_CHECK, k_count_zone, blocks_for(n_changed, 256), 256, 0, c->list2, n_changed,
           (c->halo + zone) * plane, (c->g.nx - c->halo - zone) * plane, c->d_cnt + CNT_CENTRES);
    TRY(read_counters(c));
    *count = (int64_t)c->h_cnt[CNT_CENTRES];
    return 0;
}

// bader_calc('neargrid') of a sharded run after the seed is numbered: the conservative
// rounds of converge_rounds, the ranks meeting at one all-reduce per decision
int bdr_slab_rounds(bdr_ctx *c, int which, const double *dist_mat, const double *T_grad,
                    int64_t max_passes, int64_t *history, int64_t hist_cap, int64_t *n_hist,
                    int *settled) {
    TRY(check(c));
    SlabComm *sc = static_cast<SlabComm *>(c->slab_comm);
    if (!sc) return fail_msg("bdr_slab_rounds: bdr_slab_comm_init has not run");
    if (which < 0 || which > 1 || !c->labels[which]) return fail_msg("bdr_slab_rounds: bad label set");
    const Weights W = make_weights(dist_mat);
    const TGrad T = make_tgrad(T_grad);
    const bool dbg = getenv("BDR_DEBUG") != nullptr && sc->rank == 0;
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [&]() {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    };
    int64_t run = 0;
    auto record = [&](int64_t a, int64_t b) {
        if (history && run < hist_cap) {
            history[2 * run] = a;
            history[2 * run + 1] = b;
        }
        ++run;
    };
    int32_t *lab = c->labels[which];
    const int64_t plane = (int64_t)c->g.ny * c->g.nz;
    const int H = c->halo, Wp = c->g.nx;
    TRY(comm_halo_exchange(c, sc, lab, 4));
    int64_t e = 0;
    TRY(edge_find_dev(c, which, &e, -1, 1));
    TRY(slab_publish_known(c, sc));
    c->last_changed = 0;
    if (!c->term) CU(cudaMalloc((void **)&c->term, (size_t)c->N * sizeof(int32_t)));
    CU(cudaMemsetAsync(c->term, 0xff, (size_t)c->N * sizeof(int32_t), c->stream));
    c->use_term = true;
    long long v[2] = {(long long)e, 0};
    TRY(comm_allreduce(c, sc, 1, v));     // also the barrier before the first remote reads
    const int64_t edges = v[0];
    int64_t changed = 0, zone_changed = 0;
    if (edges > 0) {
        int64_t ch = 0, zc = 0;
        TRY(trace_dev(c, which, W, T, &ch, true));
        c->last_changed = ch;
        TRY(slab_zone_changed(c, ch, &zc));
        v[0] = ch;
        v[1] = zc;
        TRY(comm_allreduce(c, sc, 2, v));
        changed = v[0];
        zone_changed = v[1];
    }
    record(edges, changed);
    if (dbg) fprintf(stderr, "[bdr slab] first pass: edges %lld changed %lld, %lld walks left the window  (t = %.2f ms)\n", (long long)edges, (long long)changed, (long long)c->h_cnt[CNT_ESCLIST], ms_since());
    while (changed > 0 && run < max_passes) {
        // No rank relabelled anything within 2*halo planes of a slab boundary: every halo copy
        // (labels and known) is still what its owner holds, and the re-classification below
        // stays 2*halo - 2 planes away from the halos.  Nothing to exchange this round.
        const bool quiet_boundaries = zone_changed == 0 && !getenv("BDR_ALWAYS_EXCHANGE");
        int64_t n_extra = 0;
        if (!quiet_boundaries) {
        // the planes next to the owned slab, before and after the exchange: voxels a
        // neighbour relabelled there count as changed here too
        CU(cudaMemcpyAsync(sc->plane_lo, lab + (int64_t)(H - 1) * plane, (size_t)plane * 4, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(sc->plane_hi, lab + (int64_t)(Wp - H) * plane, (size_t)plane * 4, cudaMemcpyDeviceToDevice, c->stream));
        TRY(comm_halo_exchange(c, sc, lab, 4));
        const int64_t need = c->last_changed + 2 * plane;
        if (need > c->list2_cap) {
            int32_t *grown = nullptr;
            const int64_t cap = need + need / 4 + 1024;
            CU(cudaMalloc((void **)&grown, (size_t)cap * sizeof(int32_t)));
            if (c->last_changed)
                CU(cudaMemcpyAsync(grown, c->list2, (size_t)c->last_changed * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            if (c->list2) cudaFree(c->list2);
            c->list2 = grown;
            c->list2_cap = cap;
        }
        TRY(zero_counter(c, CNT_CENTRES));
        LAUNCH(c, BDR_K_EDGE_CHECK, k_plane_diff, blocks_for(plane, 256), 256, 0, lab + (int64_t)(H - 1) * plane,
               sc->plane_lo, (int)plane, (int)((H - 1) * plane), c->d_cnt + CNT_CENTRES, c->list2, c->last_changed,
               c->list2_cap);
        LAUNCH(c, BDR_K_EDGE_CHECK, k_plane_diff, blocks_for(plane, 256), 256, 0, lab + (int64_t)(Wp - H) * plane,
               sc->plane_hi, (int)plane, (int)((Wp - H) * plane), c->d_cnt + CNT_CENTRES, c->list2, c->last_changed,
               c->list2_cap);
        TRY(read_counters(c));
        n_extra = (int64_t)c->h_cnt[CNT_CENTRES];
        }
        const int64_t n = c->last_changed + n_extra;
        int64_t q = 0;
        // (a quiet round takes the list-based update whatever its size: that one stays away
        // from the halo planes of known, which therefore remain the owners' copies -- and the
        // choice between the two updates may differ from rank to rank, an exchange may not)
        if (n * 2048 > c->N && !quiet_boundaries) {
            TRY(edge_find_dev(c, which, &q, n, 2));
        } else {
            TRY(incremental_dev(c, which, n, &q));
        }
        if (quiet_boundaries) c->halo_known_current = true;
        else TRY(slab_publish_known(c, sc));
        TRY(filter_cached_dev(c));
        v[0] = c->list_n;
        TRY(comm_allreduce(c, sc, 1, v));
        const int64_t queued = v[0];
        int64_t ch = 0, zc = 0;
        TRY(trace_dev(c, which, W, T, &ch, true));
        c->last_changed = ch;
        TRY(slab_zone_changed(c, ch, &zc));
        v[0] = ch;
        v[1] = zc;
        TRY(comm_allreduce(c, sc, 2, v));
        changed = v[0];
        zone_changed = v[1];
        record(queued, changed);
        if (dbg) fprintf(stderr, "[bdr slab] round %lld: queued %lld changed %lld  (t = %.2f ms)\n", (long long)run - 1, (long long)queued, (long long)changed, ms_since());
    }
    c->use_term = false;
    c->last_changed = 0;
    TRY(comm_halo_exchange(c, sc, lab, 4));
    CU(cudaStreamSynchronize(c->stream));
    if (n_hist) *n_hist = run;
    if (settled) *settled = changed == 0;
    return 0;
}

// thread_handlers.refine (thread_handlers.py:128-236) across the slabs: refine_dev with the
// ranks meeting where the reference's serial code looks at the whole grid
int bdr_slab_refine(bdr_ctx *c, int which, int mode, int64_t iters, const double *dist_mat,
                    const double *T_grad, int64_t *iters_run, int64_t *history, int64_t hist_cap) {
    TRY(check(c));
    SlabComm *sc = static_cast<SlabComm *>(c->slab_comm);
    if (!sc) return fail_msg("bdr_slab_refine: bdr_slab_comm_init has not run");
    if (which < 0 || which > 1 || !c->labels[which]) return fail_msg("bdr_slab_refine: bad label set");
    if (mode != BDR_MODE_ALL && mode != BDR_MODE_CHANGED) return fail_msg("bdr_slab_refine: bad mode");
    const Weights W = make_weights(dist_mat);
    const TGrad T = make_tgrad(T_grad);
    const bool chg_mode = mode == BDR_MODE_CHANGED;
    const bool dbg = getenv("BDR_DEBUG") != nullptr && sc->rank == 0;
    int64_t run = 0;
    auto record = [&](int64_t a, int64_t b) {
        if (history && run < hist_cap) {
            history[2 * run] = a;
            history[2 * run + 1] = b;
        }
        ++run;
    };
    if (iters_run) *iters_run = 0;
    if (iters == 0) return 0;
    int32_t *lab = c->labels[which];
    c->use_term = false;
    int64_t edges = 0, changed = 0;
    for (int64_t it = 0; iters < 0 || it < iters; ++it) {
        TRY(comm_halo_exchange(c, sc, lab, 4));
        long long v[2] = {0, 0};
        if (it == 0 || !chg_mode) {
            int64_t e = 0;
            TRY(edge_find_dev(c, which, &e));
            TRY(slab_publish_known(c, sc));
            v[0] = e;
            TRY(comm_allreduce(c, sc, 1, v));      // the barrier before the remote reads, too
            edges = v[0];
            if (edges == 0) break;
        } else {
            TRY(ec_begin_dev(c, which, c->last_changed));
            for (int round = 0;; ++round) {
                TRY(comm_halo_exchange(c, sc, c->known, 1));
                int64_t u = 0;
                TRY(ec_round_dev(c, c->last_changed, &u));
                v[0] = u;
                TRY(comm_allreduce(c, sc, 1, v));
                if (v[0] == 0) break;
                if (round > 1 << 20) return fail_msg("edge_check: centre selection did not converge");
            }
            TRY(comm_halo_exchange(c, sc, c->known, 1));
            int64_t e = 0;
            TRY(ec_finish_dev(c, which, c->last_changed, &e));
            TRY(slab_publish_known(c, sc));
            v[0] = e;
            TRY(comm_allreduce(c, sc, 1, v));
            edges = v[0];
        }
        int64_t ch = 0;
        TRY(trace_dev(c, which, W, T, &ch, chg_mode));
        c->last_changed = chg_mode ? ch : 0;
        v[0] = ch;
        v[1] = c->escaped;
        TRY(comm_allreduce(c, sc, 2, v));
        changed = v[0];
        if (v[1]) return fail_msg("bdr_slab_refine: a trajectory left the slab halo: raise the halo");
        record(edges, changed);
        if (dbg) fprintf(stderr, "[bdr slab] refine pass %lld: edges %lld changed %lld\n", (long long)it, (long long)edges, (long long)changed);
        if (changed == 0) {
            if (it == 0 && (iters < 0 || iters >= 2)) record(chg_mode ? 0 : edges, 0);
            break;
        }
    }
    c->last_changed = 0;
    TRY(comm_halo_exchange(c, sc, lab, 4));
    CU(cudaStreamSynchronize(c->stream));
    if (iters_run) *iters_run = run;
    return 0;
}

int bdr_set_stream(bdr_ctx *c, void *stream) {
    TRY(check(c));
    CU(cudaStreamSynchronize(c->stream));
    if (c->owns_stream) cudaStreamDestroy(c->stream);
    c->stream = static_cast<cudaStream_t>(stream);
    c->owns_stream = false;
    return 0;
}

int bdr_trace_pass(bdr_ctx *c, int which, const double *dist_mat, const double *T_grad,
                   int64_t *changed, int64_t *escaped) {
    return bdr_trace_pass_list(c, which, dist_mat, T_grad, 0, changed, escaped);
}

int bdr_trace_pass_list(bdr_ctx *c, int which, const double *dist_mat, const double *T_grad,
                        int want_list, int64_t *changed, int64_t *escaped) {
    TRY(check(c));
    if (which < 0 || which > 1 || !c->labels[which]) return fail_msg("bdr_trace_pass: bad label set");
    if (!c->known) return fail_msg("bdr_trace_pass: run bdr_edge_pass first");
    c->use_term = false;  // an exact pass: nothing cached applies
    const Weights W = make_weights(dist_mat);
    const TGrad T = make_tgrad(T_grad);
    int64_t ch = 0;
    TRY(trace_dev(c, which, W, T, &ch, want_list != 0));
    c->last_changed = want_list ? ch : 0;
    CU(cudaStreamSynchronize(c->stream));
    if (changed) *changed = ch;
    if (escaped) *escaped = c->escaped;
    return 0;
}

int bdr_destroy(bdr_ctx *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 3; ++i)
        if (c->rho[i]) cudaFree(c->rho[i]);
    for (int i = 0; i < 2; ++i)
        if (c->labels[i]) cudaFree(c->labels[i]);
    for (void *p : {(void *)c->known, (void *)c->list, (void *)c->list2, (void *)c->list3,
                    (void *)c->roots, (void *)c->minidx, (void *)c->rank, (void *)c->d_cnt,
                    (void *)c->d_sums, c->stage, (void *)c->ebits, (void *)c->term,
                    (void *)c->tile_keys, (void *)c->tile_order, (void *)c->tile_hist,
                    (void *)c->d_seedw, (void *)c->defer, (void *)c->list4, (void *)c->eq_pending})
        if (p) cudaFree(p);
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    if (c->pinned) cudaFreeHost(c->pinned);
    for (auto &r : c->recs) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    delete static_cast<PeerView *>(c->peer_view);
    if (c->slab_comm) {
        SlabComm *sc = static_cast<SlabComm *>(c->slab_comm);
        if (sc->comm && nccl_api()) nccl_api()->CommDestroy(sc->comm);
        if (sc->d_red) cudaFree(sc->d_red);
        if (sc->h_red) cudaFreeHost(sc->h_red);
        if (sc->plane_lo) cudaFree(sc->plane_lo);
        if (sc->plane_hi) cudaFree(sc->plane_hi);
        delete sc;
    }
    for (auto e : c->pool) cudaEventDestroy(e);
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    for (auto e : c->chunk_events) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->owns_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int bdr_synchronize(bdr_ctx *c) {
    TRY(check(c));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_upload_density(bdr_ctx *c, int which, const double *host) {
    TRY(check(c));
    if (which < 0 || which > 2 || !host) return fail_msg("bdr_upload_density: bad argument");
    if (which == BDR_RHO_REFERENCE) c->maxima_fresh[0] = c->maxima_fresh[1] = false;
    TRY(ensure_rho(c, which));
    CU(cudaMemcpyAsync(c->rho[which], host, (size_t)c->N * sizeof(double), cudaMemcpyHostToDevice,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_download_density(bdr_ctx *c, int which, double *host) {
    TRY(check(c));
    if (which < 0 || which > 2 || !host) return fail_msg("bdr_download_density: bad argument");
    const double *p = rho_ptr(c, which);
    if (!p) return fail_msg("bdr_download_density: slot is empty");
    CU(cudaMemcpyAsync(host, p, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_alias_density(bdr_ctx *c, int which, int of) {
    TRY(check(c));
    if (which < 0 || which > 2 || of < 0 || of > 2 || which == of)
        return fail_msg("bdr_alias_density: bad argument");
    if (which == BDR_RHO_REFERENCE) c->maxima_fresh[0] = c->maxima_fresh[1] = false;
    if (c->rho[which]) {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(c->rho[which]);
        c->rho[which] = nullptr;
    }
    c->rho_alias[which] = of;
    return 0;
}

int bdr_copy_density(bdr_ctx *c, int dst, int src) {
    TRY(check(c));
    if (dst < 0 || dst > 2 || src < 0 || src > 2) return fail_msg("bdr_copy_density: bad argument");
    const double *from = rho_ptr(c, src);
    if (!from) return fail_msg("bdr_copy_density: source slot is empty");
    if (c->rho[dst] == from) return 0;
    if (dst == BDR_RHO_REFERENCE) c->maxima_fresh[0] = c->maxima_fresh[1] = false;
    // (a dst that merely aliased src gets storage of its own here)
    TRY(ensure_rho(c, dst));
    CU(cudaMemcpyAsync(c->rho[dst], from, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_clear_labels(bdr_ctx *c, int which) {
    TRY(check(c));
    if (which < 0 || which > 1) return fail_msg("bdr_clear_labels: bad argument");
    c->maxima_fresh[which] = false;
    eq_invalidate(c, which);
    TRY(ensure_labels(c, which));
    CU(cudaMemsetAsync(c->labels[which], 0, (size_t)c->N * sizeof(int32_t), c->stream));
    if (which == BDR_LABELS_BADER) c->vac_mode = VAC_NONE;
    return 0;
}

int bdr_upload_labels(bdr_ctx *c, int which, const void *host, int elem_size) {
    TRY(check(c));
    if (which < 0 || which > 1 || !host) return fail_msg("bdr_upload_labels: bad argument");
    c->maxima_fresh[which] = false;
    eq_invalidate(c, which);
    TRY(ensure_labels(c, which));
    if (which == BDR_LABELS_BADER) c->vac_mode = VAC_LABELS;
    switch (elem_size) {
        case 1: return upload_cast<int8_t>(c, c->labels[which], host);
        case 2: return upload_cast<int16_t>(c, c->labels[which], host);
        case 4:
            CU(cudaMemcpyAsync(c->labels[which], host, (size_t)c->N * 4, cudaMemcpyHostToDevice,
                               c->stream));
            CU(cudaStreamSynchronize(c->stream));
            return 0;
        case 8: return upload_cast<int64_t>(c, c->labels[which], host);
    }
    return fail_msg("bdr_upload_labels: elem_size must be 1, 2, 4 or 8");
}

int bdr_download_labels(bdr_ctx *c, int which, void *host, int elem_size) {
    TRY(check(c));
    if (which < 0 || which > 1 || !host) return fail_msg("bdr_download_labels: bad argument");
    if (!c->labels[which]) return fail_msg("bdr_download_labels: label set is empty");
    switch (elem_size) {
        case 1: return download_cast<int8_t>(c, c->labels[which], host);
        case 2: return download_cast<int16_t>(c, c->labels[which], host);
        case 4:
            CU(cudaMemcpyAsync(host, c->labels[which], (size_t)c->N * 4, cudaMemcpyDeviceToHost,
                               c->stream));
            CU(cudaStreamSynchronize(c->stream));
            return 0;
        case 8: return download_cast<int64_t>(c, c->labels[which], host);
    }
    return fail_msg("bdr_download_labels: elem_size must be 1, 2, 4 or 8");
}

int bdr_download_known(bdr_ctx *c, int8_t *host) {
    TRY(check(c));
    if (!c->known) return fail_msg("bdr_download_known: no edge pass has run");
    CU(cudaMemcpyAsync(host, c->known, (size_t)c->N, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_vacuum_assign(bdr_ctx *c, double vac_tol, double voxel_volume, int which_density,
                      double *vac_charge, double *vac_volume) {
    TRY(check(c));
    const double *ref = rho_ptr(c, BDR_RHO_REFERENCE);
    const double *dens = rho_ptr(c, which_density);
    if (!ref || !dens) return fail_msg("bdr_vacuum_assign: density not uploaded");
    c->maxima_fresh[0] = c->maxima_fresh[1] = false;
    eq_invalidate(c, BDR_LABELS_BADER);
    TRY(ensure_labels(c, BDR_LABELS_BADER));
    TRY(ensure_sums(c, 2));
    CU(cudaMemsetAsync(c->d_sums, 0, sizeof(double), c->stream));
    TRY(zero_counter(c, CNT_VACUUM));
    const unsigned nb = std::min<unsigned>(blocks_for(c->N, 256 * 8), 148 * 16);
    LAUNCH(c, BDR_K_VACUUM, k_vacuum, nb, 256, 0, ref, dens, c->labels[BDR_LABELS_BADER], c->N,
           vac_tol, c->d_sums, c->d_cnt + CNT_VACUUM, c->own_lo, c->own_hi);
    double s = 0;
    CU(cudaMemcpyAsync(&s, c->d_sums, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    TRY(read_counters(c));
    // the stencil pass may re-derive the mask from the tolerance only if the
    // labels held nothing but zeros before this call
    c->vac_mode = (c->vac_mode == VAC_NONE) ? VAC_TOL : VAC_LABELS;
    c->vac_tol = vac_tol;
    if (vac_charge) *vac_charge = s * voxel_volume;
    if (vac_volume) *vac_volume = (double)c->h_cnt[CNT_VACUUM] * voxel_volume;
    c->vac_count = (int64_t)c->h_cnt[CNT_VACUUM];
    return 0;
}

int bdr_vacuum_count(bdr_ctx *c, int64_t *count) {
    TRY(check(c));
    if (!count) return fail_msg("bdr_vacuum_count: null out");
    *count = c->vac_count;
    return 0;
}

static int bader_calc_dev(bdr_ctx *c, int method, const double *dist_mat, const double *T_grad,
                          int64_t *n_maxima, const double *host_density) {
    if (!dist_mat) return fail_msg("bdr_bader_calc: dist_mat is null");
    if (method != BDR_METHOD_ONGRID && method != BDR_METHOD_NEARGRID)
        return fail_msg("bdr_bader_calc: unknown method");
    if (method == BDR_METHOD_NEARGRID && !T_grad) return fail_msg("bdr_bader_calc: T_grad is null");
    const Weights W = make_weights(dist_mat);
    const int vac_mode_at_entry = c->vac_mode;
    c->maxima_fresh[0] = c->maxima_fresh[1] = false;
    eq_invalidate(c, BDR_LABELS_BADER);
    TRY(choose_seed(c, method, W));
    int64_t seeded = -1;
    if (host_density) TRY(upload_and_stencil_dev(c, host_density, W, &seeded));
    if (!rho_ptr(c, BDR_RHO_REFERENCE)) return fail_msg("bdr_bader_calc: reference density not uploaded");
    TRY(ongrid_dev(c, W, seeded, method == BDR_METHOD_ONGRID));
    if (method == BDR_METHOD_NEARGRID) {
        const TGrad T = make_tgrad(T_grad);
        // seeded with the pointer-jumpable ongrid field (still as slot codes, which
        // serve as labels); now drive the order-free refinement iteration to its
        // fixed point, then number the volumes once
        TRY(converge_dev(c, BDR_LABELS_BADER, W, T));
        TRY(number_slots_dev(c, false));
    }
    // the stencil pass's maxima stay valid for the edge passes that follow as long as the
    // reference density and the vacuum mask (a threshold of that density, or none) do
    c->maxima_fresh[BDR_LABELS_BADER] = vac_mode_at_entry != VAC_LABELS;
    c->maxima_fresh[BDR_LABELS_ATOMS] = false;
    c->vac_mode = VAC_LABELS;
    if (n_maxima) *n_maxima = c->n_max;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_bader_calc(bdr_ctx *c, int method, const double *dist_mat, const double *T_grad,
                   int64_t *n_maxima) {
    TRY(check(c));
    return bader_calc_dev(c, method, dist_mat, T_grad, n_maxima, nullptr);
}

int bdr_get_maxima(bdr_ctx *c, int64_t *out, int64_t cap) {
    TRY(check(c));
    if (cap < c->n_max) return fail_msg("bdr_get_maxima: buffer too small");
    if (c->n_max) memcpy(out, c->maxima.data(), (size_t)c->n_max * 3 * sizeof(int64_t));
    return 0;
}

int bdr_refine(bdr_ctx *c, int which, int mode, int64_t iters, const double *dist_mat,
               const double *T_grad, int64_t *iters_run, int64_t *history, int64_t hist_cap) {
    TRY(check(c));
    if (which < 0 || which > 1) return fail_msg("bdr_refine: bad label set");
    if (!c->labels[which]) return fail_msg("bdr_refine: label set is empty");
    if (!rho_ptr(c, BDR_RHO_REFERENCE)) return fail_msg("bdr_refine: reference density not uploaded");
    if (mode != BDR_MODE_ALL && mode != BDR_MODE_CHANGED) return fail_msg("bdr_refine: bad mode");
    const Weights W = make_weights(dist_mat);
    const TGrad T = make_tgrad(T_grad);
    TRY(refine_dev(c, which, mode, iters, W, T, iters_run, history, hist_cap));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_edge_find(bdr_ctx *c, int which, int64_t *edges) {
    TRY(check(c));
    if (which < 0 || which > 1) return fail_msg("bdr_edge_find: bad label set");
    int64_t e = 0;
    TRY(edge_find_dev(c, which, &e));
    CU(cudaStreamSynchronize(c->stream));
    if (edges) *edges = e;
    return 0;
}

int bdr_charge_sum(bdr_ctx *c, int which_labels, int which_density, double voxel_volume, int64_t n,
                   double *charge, double *volume) {
    TRY(check(c));
    if (which_labels < 0 || which_labels > 1 || which_density < 0 || which_density > 2)
        return fail_msg("bdr_charge_sum: bad argument");
    return charge_sum_dev(c, which_labels, which_density, voxel_volume, n, charge, volume);
}

int bdr_assign_atoms(bdr_ctx *c, const double *maxima_cart, int64_t n_max, const double *atoms_cart,
                     int64_t n_atoms, const double *lattice, int64_t *bader_atoms,
                     double *bader_distance) {
    TRY(check(c));
    if (n_atoms < 1) return fail_msg("bdr_assign_atoms: no atoms");
    if (!c->labels[BDR_LABELS_BADER]) return fail_msg("bdr_assign_atoms: label set is empty");
    TRY(ensure_labels(c, BDR_LABELS_ATOMS));
    const int64_t nd = 3 * n_max + 3 * n_atoms + 9 + 2 * n_max;
    TRY(ensure_sums(c, nd));
    double *d_max = c->d_sums, *d_atoms = d_max + 3 * n_max, *d_lat = d_atoms + 3 * n_atoms;
    long long *d_who = reinterpret_cast<long long *>(d_lat + 9);
    double *d_dist = reinterpret_cast<double *>(d_who + n_max);
    std::vector<long long> who((size_t)n_max);
    if (n_max > 0) {
        CU(cudaMemcpyAsync(d_max, maxima_cart, (size_t)3 * n_max * sizeof(double),
                           cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_atoms, atoms_cart, (size_t)3 * n_atoms * sizeof(double),
                           cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_lat, lattice, 9 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        LAUNCH(c, BDR_K_ASSIGN, k_atom_assign, blocks_for(n_max, 128), 128, 0, d_max, n_max, d_atoms,
               n_atoms, d_lat, d_who, d_dist);
        CU(cudaMemcpyAsync(who.data(), d_who, (size_t)n_max * sizeof(long long),
                           cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(bader_distance, d_dist, (size_t)n_max * sizeof(double),
                           cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (int64_t i = 0; i < n_max; ++i) bader_atoms[i] = who[(size_t)i];
    }
    // utils.volume_assign: atoms labels = LUT(bader labels)
    TRY(ensure_slots(c, std::max<int64_t>(n_max, 1)));
    std::vector<int32_t> lut((size_t)std::max<int64_t>(n_max, 1), 0);
    for (int64_t i = 0; i < n_max; ++i) lut[(size_t)i] = (int32_t)who[(size_t)i];
    CU(cudaMemcpyAsync(c->rank, lut.data(), lut.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                       c->stream));
    LAUNCH(c, BDR_K_ASSIGN, k_relabel_lut, blocks_for(c->N, 1024), 256, 0, c->labels[BDR_LABELS_BADER],
           c->labels[BDR_LABELS_ATOMS], c->N, c->rank);
    // the atom labels are the LUT image of the Bader labels: same vacuum voxels, same maxima
    c->maxima_fresh[BDR_LABELS_ATOMS] = c->maxima_fresh[BDR_LABELS_BADER];
    eq_invalidate(c, BDR_LABELS_ATOMS);
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// edge_find on label set `which`, then the smallest squared distance from every atom to an
// (owned) edge voxel of its own volume over the 27 lattice images (utils.py:321-379)
static int surface_distance_dev(bdr_ctx *c, int which, const double *lattice, const double *atoms_cart,
                                int64_t n_atoms, std::vector<double> &best, std::vector<unsigned long long> &seen,
                                int64_t *edges) {
    TRY(edge_find_dev(c, which, edges));
    // utils.py:341-343: the running minimum starts at nx^2+ny^2+nz^2
    const PeerView *pv = static_cast<const PeerView *>(c->peer_view);
    const int64_t NX = pv ? pv->NX : c->g.nx;
    const double init = (double)(NX * NX + (int64_t)c->g.ny * c->g.ny + (int64_t)c->g.nz * c->g.nz);
    best.assign((size_t)n_atoms, init);
    seen.assign((size_t)n_atoms, 0);
    if (c->list_n == 0) return 0;
    TRY(ensure_sums(c, 5 * n_atoms + 9));
    double *d_atoms = c->d_sums, *d_lat = d_atoms + 3 * n_atoms;
    unsigned long long *d_best = reinterpret_cast<unsigned long long *>(d_lat + 9);
    unsigned long long *d_seen = d_best + n_atoms;
    CU(cudaMemcpyAsync(d_atoms, atoms_cart, (size_t)3 * n_atoms * sizeof(double),
                       cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_lat, lattice, 9 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_best, best.data(), (size_t)n_atoms * sizeof(double), cudaMemcpyHostToDevice,
                       c->stream));
    CU(cudaMemsetAsync(d_seen, 0, (size_t)n_atoms * sizeof(unsigned long long), c->stream));
    LAUNCH(c, BDR_K_SURFACE, k_surface_dist, blocks_for(c->list_n, 128), 128, 0, c->labels[which],
           c->known, c->g, c->list, c->list_n, d_lat, d_atoms, d_best, d_seen, (int)n_atoms,
           (int)c->own_lo, (int)c->own_hi, pv ? pv->x0w : 0, (int)NX);
    CU(cudaMemcpyAsync(best.data(), d_best, (size_t)n_atoms * sizeof(double), cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaMemcpyAsync(seen.data(), d_seen, (size_t)n_atoms * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_surface_distance(bdr_ctx *c, int which, const double *lattice, const double *atoms_cart,
                         int64_t n_atoms, double *distance, int *found) {
    TRY(check(c));
    if (which < 0 || which > 1) return fail_msg("bdr_surface_distance: bad label set");
    if (c->halo > 0) return fail_msg("bdr_surface_distance: slab windows go through bdr_slab_surface_distance");
    int64_t edges = 0;
    std::vector<double> best;
    std::vector<unsigned long long> seen;
    TRY(surface_distance_dev(c, which, lattice, atoms_cart, n_atoms, best, seen, &edges));
    if (found) *found = edges > 0;
    // atoms that own no edge voxel keep 0 (thread_handlers.py:289-297)
    for (int64_t a = 0; a < n_atoms; ++a)
        distance[a] = (edges > 0 && seen[(size_t)a]) ? std::sqrt(best[(size_t)a]) : 0.0;
    return 0;
}

// one rank's share of thread_handlers.surface_distance: the smallest SQUARED distance per
// atom over the edge voxels this slab owns (the caller takes the minimum over the ranks and
// the square root), whether the atom's volume has an owned edge voxel, and the owned edges
int bdr_slab_surface_distance(bdr_ctx *c, int which, const double *lattice, const double *atoms_cart,
                              int64_t n_atoms, double *best_sq, int64_t *seen_out, int64_t *edges_owned) {
    TRY(check(c));
    if (which < 0 || which > 1) return fail_msg("bdr_slab_surface_distance: bad label set");
    if (c->halo == 0) return fail_msg("bdr_slab_surface_distance: not a slab handle");
    int64_t edges = 0;
    std::vector<double> best;
    std::vector<unsigned long long> seen;
    TRY(surface_distance_dev(c, which, lattice, atoms_cart, n_atoms, best, seen, &edges));
    for (int64_t a = 0; a < n_atoms; ++a) {
        best_sq[a] = best[(size_t)a];
        seen_out[a] = (int64_t)seen[(size_t)a];
    }
    if (edges_owned) *edges_owned = edges;
    return 0;
}

int bdr_volume_mask(bdr_ctx *c, int which_labels, int which_density, int64_t vol_num,
                    double *host_out) {
    TRY(check(c));
    if (which_labels < 0 || which_labels > 1 || which_density < 0 || which_density > 2 || !host_out)
        return fail_msg("bdr_volume_mask: bad argument");
    if (!c->labels[which_labels]) return fail_msg("bdr_volume_mask: label set is empty");
    const double *dens = rho_ptr(c, which_density);
    if (!dens) return fail_msg("bdr_volume_mask: density slot is empty");
    TRY(ensure_stage(c, (size_t)c->N * sizeof(double)));
    LAUNCH(c, BDR_K_NARROW, k_volume_mask, blocks_for(c->N, 256), 256, 0, c->labels[which_labels],
           dens, (double *)c->stage, c->N, (int32_t)vol_num);
    CU(cudaMemcpyAsync(host_out, c->stage, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int bdr_run(bdr_ctx *c, const double *host_density, double vac_tol, double voxel_volume, int method,
            int refine_mode, int64_t refine_iters, const double *dist_mat, const double *T_grad,
            void *host_labels, int label_elem_size, int64_t *n_maxima, int64_t *maxima,
            int64_t max_cap, double *charge, double *volume) {
    TRY(check(c));
    if (!host_density) return fail_msg("bdr_run: host_density is null");
    // volumes_init: labels start at 0; the vacuum mask (reference <= tol) is
    // applied by the stencil pass itself
    TRY(ensure_labels(c, BDR_LABELS_BADER));
    c->vac_mode = (vac_tol == vac_tol) ? VAC_TOL : VAC_NONE;
    c->vac_tol = vac_tol;
    int64_t n = 0;
    const bool dbg = getenv("BDR_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(
                        std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    TRY(bader_calc_dev(c, method, dist_mat, T_grad, &n, host_density));
    const double t1 = now();
    if (refine_iters != 0) {
        int64_t run = 0;
        TRY(bdr_refine(c, BDR_LABELS_BADER, refine_mode, refine_iters, dist_mat, T_grad, &run,
                       nullptr, 0));
    }
    const double t2 = now();
    if (n_maxima) *n_maxima = n;
    if ((maxima || charge || volume) && n > max_cap)
        return fail_msg("bdr_run: " + std::to_string(n) + " maxima do not fit max_cap = " + std::to_string(max_cap));
    if (maxima) TRY(bdr_get_maxima(c, maxima, max_cap));
    if (charge || volume) {
        for (int64_t i = 0; i < n; ++i) {
            if (charge) charge[i] = 0;
            if (volume) volume[i] = 0;
        }
        TRY(charge_sum_dev(c, BDR_LABELS_BADER, BDR_RHO_REFERENCE, voxel_volume, n, charge, volume));
    }
    if (host_labels) TRY(bdr_download_labels(c, BDR_LABELS_BADER, host_labels, label_elem_size));
    if (dbg)
        fprintf(stderr, "[bdr] run: upload+bader_calc %.1f ms (upload+stencil %.1f), refine %.1f ms, sums+download %.1f ms\n",
                t1 - t0, c->dbg_upload_ms, t2 - t1, now() - t2);
    return 0;
}

// ---- text -> grid --------------------------------------
struct ParsePool {
    void *p[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaError_t reserve(int i, size_t bytes) {
        if (cap[i] >= bytes) return cudaSuccess;
        if (p[i]) cudaFree(p[i]);
        p[i] = nullptr;
        cap[i] = 0;
        const cudaError_t e = cudaMalloc(&p[i], bytes);
        if (e == cudaSuccess) cap[i] = bytes;
        return e;
    }
    void release() {
        for (int i = 0; i < 8; ++i) {
            if (p[i]) cudaFree(p[i]);
            p[i] = nullptr;
            cap[i] = 0;
        }
    }
};
static ParsePool &parse_pool(int device) {
    static ParsePool pools[64];
    return pools[device & 63];
}

int bdr_parse_release(int device) {
    CU(cudaSetDevice(device));
    parse_pool(device).release();
    return 0;
}

// page-locked host memory for the readers (text in, grid out): PCIe at full rate
void *bdr_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (bytes <= 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
        fail_msg("bdr_host_alloc: cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}
int bdr_host_free(void *p) {
    if (p) CU(cudaFreeHost(p));
    return 0;
}

// Content fingerprint of a host array (the Python session keys device residency on
// it: an in-place edit anywhere in the array must change the key).  All host threads
// stream the bytes once; 64-bit words go through four independent multiply-xorshift
// lanes per 4 KiB block, block hashes are combined with their position.  Also
// reports whether every byte is zero (a fresh label array needs no upload).
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
int bdr_host_hash(const void *data, int64_t nbytes, int threads, uint64_t *hash, int *all_zero) {
    if ((!data && nbytes > 0) || nbytes < 0 || !hash) return fail_msg("bdr_host_hash: bad argument");
    const unsigned char *p = static_cast<const unsigned char *>(data);
    constexpr int64_t BLOCK = 1 << 16;
    const int64_t nblocks = (nbytes + BLOCK - 1) / BLOCK;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(nt, 64), (nblocks + 63) / 64));
    std::vector<uint64_t> acc((size_t)nt, 0), orr((size_t)nt, 0);
    auto work = [&](int t) {
        uint64_t a = 0, nz = 0;
        for (int64_t b = t; b < nblocks; b += nt) {
            const int64_t lo = b * BLOCK, hi = std::min(nbytes, lo + BLOCK);
            uint64_t h0 = 0x9E3779B97F4A7C15ULL, h1 = 0xC2B2AE3D27D4EB4FULL, h2 = 0x165667B19E3779F9ULL,
                     h3 = 0x27D4EB2F165667C5ULL, o = 0;
            int64_t i = lo;
            for (; i + 32 <= hi; i += 32) {
                uint64_t w[4];
                memcpy(w, p + i, 32);
                o |= w[0] | w[1] | w[2] | w[3];
                h0 = (h0 ^ w[0]) * 0xFF51AFD7ED558CCDULL; h0 ^= h0 >> 29;
                h1 = (h1 ^ w[1]) * 0xC4CEB9FE1A85EC53ULL; h1 ^= h1 >> 31;
                h2 = (h2 ^ w[2]) * 0x9FB21C651E98DF25ULL; h2 ^= h2 >> 28;
                h3 = (h3 ^ w[3]) * 0xD6E8FEB86659FD93ULL; h3 ^= h3 >> 32;
            }
            for (; i < hi; ++i) {
                o |= p[i];
                h0 = (h0 ^ p[i]) * 0xFF51AFD7ED558CCDULL; h0 ^= h0 >> 29;
            }
            nz |= o;
            a += mix64(mix64(h0) + 3 * mix64(h1) + 5 * mix64(h2) + 7 * mix64(h3) + (uint64_t)b * 0x9E3779B97F4A7C15ULL);
        }
        acc[(size_t)t] = a;
        orr[(size_t)t] = nz;
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t) pool.emplace_back(work, t);
        for (auto &th : pool) th.join();
    }
    uint64_t h = mix64((uint64_t)nbytes), nz = 0;
    for (int t = 0; t < nt; ++t) {
        h += acc[(size_t)t];   // block hashes carry their position: the sum is order-free
        nz |= orr[(size_t)t];
    }
    *hash = mix64(h);
    if (all_zero) *all_zero = nz == 0;
    return 0;
}

int bdr_parse_token_host(const char *token, int64_t len, double *out) {
    if (!token || !out || len <= 0) return fail_msg("bdr_parse_token_host: bad argument");
    int l = 0;
    const int st = parse_token(token, len, out, &l);
    return (st == 0 && l == len) ? 0 : 2;   // 2: not handled exactly here, ask the host's strtod
}

int bdr_parse_text(int device, const char *text, int64_t nbytes, int64_t n_values, int64_t nx,
                   int64_t ny, int64_t nz, int x_fastest, int op, double operand, double *out,
                   int64_t *tokens_found, int64_t *bytes_consumed, int64_t *n_fallback,
                   int64_t *fallback, int64_t fallback_cap) {
    if (!text || !out || nbytes < 0 || n_values <= 0 || nx * ny * nz != n_values)
        return fail_msg("bdr_parse_text: bad argument");
    if (n_values >= (1LL << 31)) return fail_msg("bdr_parse_text: more than 2^31 values");
    CU(cudaSetDevice(device));
    cudaStream_t st;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const bool dbg = getenv("BDR_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(
                        std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    const int64_t CHUNK = 256LL << 20;
    const int64_t cbytes = std::min<int64_t>(std::max<int64_t>(nbytes, 1), CHUNK);
    const int64_t max_blocks = (cbytes + TOK_BLOCK - 1) / TOK_BLOCK;
    // device buffers are kept between calls (a file is several blocks; cudaMalloc / cudaFree
    // of gigabytes cost more than the conversion); bdr_parse_release frees them
    static std::mutex pool_mutex;
    std::lock_guard<std::mutex> pool_lock(pool_mutex);
    ParsePool &pool = parse_pool(device);
    const int64_t fb_cap = std::max<int64_t>(fallback_cap, 0);
    auto cleanup = [&]() { cudaStreamDestroy(st); };
#define PCU(x)                                                    \
    do {                                                          \
        cudaError_t e_ = (x);                                     \
        if (e_ != cudaSuccess) {                                  \
            cleanup();                                            \
            return bdr::fail(#x, __FILE__, __LINE__, e_);         \
        }                                                         \
    } while (0)
    PCU(pool.reserve(0, (size_t)cbytes + 64));
    PCU(pool.reserve(1, (size_t)n_values * sizeof(double)));
    PCU(pool.reserve(2, (size_t)n_values * sizeof(double)));
    PCU(pool.reserve(3, (size_t)max_blocks * sizeof(unsigned)));
    PCU(pool.reserve(4, 1024 * sizeof(unsigned)));
    PCU(pool.reserve(5, 1024 * sizeof(unsigned)));
    PCU(pool.reserve(6, sizeof(ParseOut)));
    PCU(pool.reserve(7, (size_t)std::max<int64_t>(fb_cap, 1) * 3 * sizeof(int64_t)));
    char *d_text = (char *)pool.p[0];
    double *d_vals = (double *)pool.p[1], *d_out = (double *)pool.p[2];
    unsigned *d_counts = (unsigned *)pool.p[3], *d_tot1 = (unsigned *)pool.p[4], *d_tot2 = (unsigned *)pool.p[5];
    ParseOut *d_po = (ParseOut *)pool.p[6];
    int64_t *d_fb = (int64_t *)pool.p[7];
    PCU(cudaMemsetAsync(d_po, 0, sizeof(ParseOut), st));
    // values no token reaches read as NaN
    PCU(cudaMemsetAsync(d_vals, 0xff, (size_t)n_values * sizeof(double), st));
    int64_t pos = 0, tokens = 0;
    const double t_alloc = now();
    while (pos < nbytes && tokens < n_values) {
        // a chunk ends on whitespace, so no token straddles two chunks
        int64_t end = std::min(nbytes, pos + CHUNK);
        if (end < nbytes) {
            int64_t e = end;
            while (e > pos && !is_space((unsigned char)text[e - 1])) --e;
            if (e == pos) {
                cleanup();
                return fail_msg("bdr_parse_text: a token longer than the chunk size");
            }
            end = e;
        }
        const int64_t n = end - pos;
        const int64_t nb = (n + TOK_BLOCK - 1) / TOK_BLOCK, n1 = (nb + 1023) / 1024;
        PCU(cudaMemcpyAsync(d_text, text + pos, (size_t)n, cudaMemcpyHostToDevice, st));
        k_tok_count<<<(unsigned)nb, 256, 0, st>>>(d_text, n, d_counts);
        k_scan_local<<<(unsigned)n1, 1024, 0, st>>>(d_counts, nb, d_tot1);
        k_scan_local<<<1, 1024, 0, st>>>(d_tot1, n1, d_tot2);
        k_scan_add<<<(unsigned)n1, 1024, 0, st>>>(d_counts, nb, d_tot1);
        k_tok_parse<<<(unsigned)nb, 256, 0, st>>>(d_text, n, d_counts, tokens, pos, d_vals, n_values, d_po,
                                                  d_fb, fb_cap);
        unsigned chunk_tokens = 0;
        PCU(cudaMemcpyAsync(&chunk_tokens, d_tot2, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        PCU(cudaStreamSynchronize(st));
        PCU(cudaGetLastError());
        tokens += chunk_tokens;
        pos = end;
    }
    ParseOut po;
    PCU(cudaMemcpyAsync(&po, d_po, sizeof(ParseOut), cudaMemcpyDeviceToHost, st));
    PCU(cudaStreamSynchronize(st));
    const double t_parsed = now();
    const int64_t found = std::min(tokens, n_values);
    if (tokens_found) *tokens_found = found;
    if (bytes_consumed) *bytes_consumed = tokens >= n_values ? (int64_t)po.end_of_last : nbytes;
    if (n_fallback) *n_fallback = (int64_t)po.n_fallback;
    if (fallback && fb_cap > 0 && po.n_fallback > 0)
        PCU(cudaMemcpyAsync(fallback, d_fb,
                            (size_t)std::min<int64_t>((int64_t)po.n_fallback, fb_cap) * 3 * sizeof(int64_t),
                            cudaMemcpyDeviceToHost, st));
    if (x_fastest) {
        const dim3 grid((unsigned)(((nx + 31) / 32) * ((nz + 31) / 32)), (unsigned)ny);
        k_grid_finish<<<grid, 256, 0, st>>>(d_vals, d_out, (int)nx, (int)ny, (int)nz, 1, op, operand);
    } else {
        k_grid_finish<<<148 * 8, 256, 0, st>>>(d_vals, d_out, (int)nx, (int)ny, (int)nz, 0, op, operand);
    }
    PCU(cudaGetLastError());
    PCU(cudaStreamSynchronize(st));
    const double t_fin = now();
    PCU(cudaMemcpyAsync(out, d_out, (size_t)n_values * sizeof(double), cudaMemcpyDeviceToHost, st));
    PCU(cudaStreamSynchronize(st));
    const double t_d2h = now();
#undef PCU
    cleanup();
    if (dbg)
        fprintf(stderr, "[bdr] parse_text: %lld values from %.1f MB: alloc %.1f ms, H2D + tokenise + convert %.1f ms, "
                "layout %.2f ms, D2H %.1f ms, free %.1f ms; %lld tokens for the host\n",
                (long long)n_values, nbytes / 1e6, t_alloc - t_begin, t_parsed - t_alloc, t_fin - t_parsed,
                t_d2h - t_fin, now() - t_d2h, (long long)po.n_fallback);
    return 0;
}

int bdr_format_grid(const char *path, const double *data, int64_t nx, int64_t ny, int64_t nz,
                    int x_fastest, int64_t row_len, int per_line, int prec, int sign_space) {
    if (!path || !data) return fail_msg("bdr_format_grid: bad argument");
    std::string err;
    if (format_grid_append(path, data, nx, ny, nz, x_fastest, row_len, per_line, prec, sign_space, &err))
        return fail_msg(err);
    return 0;
}

int bdr_profile_enable(bdr_ctx *c, int on) {
    TRY(check(c));
    prof_collect(c);
    c->prof = on != 0;
    return 0;
}
int bdr_profile_reset(bdr_ctx *c) {
    TRY(check(c));
    prof_collect(c);
    for (int i = 0; i < BDR_K_COUNT; ++i) {
        c->prof_ms[i] = 0;
        c->prof_n[i] = 0;
    }
    c->trace_steps = 0;
    c->trace_voxels = 0;
    return 0;
}
int bdr_profile_get(bdr_ctx *c, int family, double *ms, int64_t *launches) {
    TRY(check(c));
    if (family < 0 || family >= BDR_K_COUNT) return fail_msg("bdr_profile_get: bad family");
    prof_collect(c);
    if (ms) *ms = c->prof_ms[family];
    if (launches) *launches = c->prof_n[family];
    return 0;
}
int bdr_launch_count(bdr_ctx *c, int64_t *launches) {
    TRY(check(c));
    *launches = c->launches;
    return 0;
}
int bdr_sync_count(bdr_ctx *c, int64_t *syncs) {
    TRY(check(c));
    *syncs = c->syncs;
    return 0;
}
int bdr_timer_start(bdr_ctx *c) {
    TRY(check(c));
    if (!c->t0) {
        CU(cudaEventCreate(&c->t0));
        CU(cudaEventCreate(&c->t1));
    }
    CU(cudaEventRecord(c->t0, c->stream));
    return 0;
}
int bdr_timer_stop(bdr_ctx *c, double *ms) {
    TRY(check(c));
    if (!c->t0) return fail_msg("bdr_timer_stop: timer was not started");
    CU(cudaEventRecord(c->t1, c->stream));
    CU(cudaEventSynchronize(c->t1));
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, c->t0, c->t1));
    *ms = f;
    return 0;
}
int bdr_trace_steps(bdr_ctx *c, int64_t *steps, int64_t *voxels) {
    TRY(check(c));
    if (steps) *steps = c->trace_steps;
    if (voxels) *voxels = c->trace_voxels;
    return 0;
}

int bdr_synth_separable(bdr_ctx *c, int which, const double *tx, const double *ty, const double *tz,
                        int64_t n_atoms) {
    TRY(check(c));
    if (which < 0 || which > 2) return fail_msg("bdr_synth_separable: bad slot");
    if (which == BDR_RHO_REFERENCE) c->maxima_fresh[0] = c->maxima_fresh[1] = false;
    TRY(ensure_rho(c, which));
    const int64_t nt = n_atoms * ((int64_t)c->g.nx + c->g.ny + c->g.nz);
    double *d_t = nullptr;
    CU(cudaMalloc((void **)&d_t, (size_t)nt * sizeof(double)));
    double *d_tx = d_t, *d_ty = d_tx + n_atoms * c->g.nx, *d_tz = d_ty + n_atoms * c->g.ny;
    CU(cudaMemcpyAsync(d_tx, tx, (size_t)n_atoms * c->g.nx * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_ty, ty, (size_t)n_atoms * c->g.ny * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_tz, tz, (size_t)n_atoms * c->g.nz * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    dim3 grid((c->g.nz + 255) / 256, c->g.ny, c->g.nx);
    LAUNCH(c, BDR_K_SYNTH, k_synth_separable, grid, 256, 0, c->rho[which], c->g, d_tx, d_ty, d_tz,
           (int)n_atoms);
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d_t);
    return 0;
}

int bdr_synth_general(bdr_ctx *c, int which, const double *lattice, const double *frac_atoms,
                      const double *amps, const double *sigmas, int64_t n_atoms) {
    TRY(check(c));
    if (which < 0 || which > 2) return fail_msg("bdr_synth_general: bad slot");
    if (which == BDR_RHO_REFERENCE) c->maxima_fresh[0] = c->maxima_fresh[1] = false;
    TRY(ensure_rho(c, which));
    double *d = nullptr;
    CU(cudaMalloc((void **)&d, (size_t)(9 + 5 * n_atoms) * sizeof(double)));
    double *d_lat = d, *d_frac = d + 9, *d_amp = d_frac + 3 * n_atoms, *d_sig = d_amp + n_atoms;
    CU(cudaMemcpyAsync(d_lat, lattice, 9 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_frac, frac_atoms, (size_t)3 * n_atoms * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_amp, amps, (size_t)n_atoms * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_sig, sigmas, (size_t)n_atoms * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    dim3 grid((c->g.nz + 255) / 256, c->g.ny, c->g.nx);
    LAUNCH(c, BDR_K_SYNTH, k_synth_general, grid, 256, 0, c->rho[which], c->g, d_lat, d_frac, d_amp,
           d_sig, (int)n_atoms);
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    return 0;
}

int bdr_selftest_div(bdr_ctx *c, int64_t n, uint64_t seed, int64_t *mismatches) {
    TRY(check(c));
    TRY(zero_counter(c, CNT_ERROR));
    LAUNCH(c, BDR_K_TRACE, k_selftest_div, blocks_for(n, 256), 256, 0, (unsigned long long)seed, n,
           c->d_cnt + CNT_ERROR);
    TRY(read_counters(c));
    *mismatches = (int64_t)c->h_cnt[CNT_ERROR];
    TRY(zero_counter(c, CNT_ERROR));
    return 0;
}

int bdr_set_option(bdr_ctx *c, int option, int64_t value) {
    TRY(check(c));
    switch (option) {
        case BDR_OPT_VERIFY_FIXED_POINT: c->verify_fixed_point = value != 0; return 0;
        case BDR_OPT_SLAB_SEED_METHOD:
            if (value != BDR_METHOD_ONGRID && value != BDR_METHOD_NEARGRID)
                return fail_msg("bdr_set_option: unknown method");
            c->slab_seed_method = (int)value;
            return 0;
    }
    return fail_msg("bdr_set_option: unknown option");
}

int bdr_device_ptr(bdr_ctx *c, int what, void **ptr) {
    TRY(check(c));
    switch (what) {
        case 0: *ptr = rho_ptr(c, 0); return 0;
        case 1: *ptr = rho_ptr(c, 1); return 0;
        case 2: *ptr = rho_ptr(c, 2); return 0;
        case 3: *ptr = c->labels[0]; return 0;
        case 4: *ptr = c->labels[1]; return 0;
        case 5: *ptr = c->known; return 0;
    }
    return fail_msg("bdr_device_ptr: bad selector");
}

}  // extern "C"

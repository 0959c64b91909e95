// common.cuh -- handle layout, error plumbing and launch/profiling helpers of
// libbader_b200.so.  Compiled only for sm_100a (see build.py); no CPU path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bader_b200.h"

namespace bdr {

extern thread_local std::string g_err;

inline int fail(const char *what, const char *file, int line, cudaError_t e) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    g_err = buf;
    return 1;
}
inline int fail_msg(const std::string &m) {
    g_err = m;
    return 1;
}

#define CU(x)                                                               \
    do {                                                                    \
        cudaError_t e_ = (x);                                               \
        if (e_ != cudaSuccess) return bdr::fail(#x, __FILE__, __LINE__, e_); \
    } while (0)
#define TRY(x)                 \
    do {                       \
        int r_ = (x);          \
        if (r_ != 0) return r_; \
    } while (0)

// grid geometry as the kernels see it (32-bit coordinates; N < 2^31 - 2^16)
struct Grid {
    int nx, ny, nz;
    // division of a linear voxel index (0 <= v < 2^31) by nz and ny as one 32x32 -> 64 bit
    // multiply and a shift: floor(v / d) == (v * m) >> s with m = ceil(2^s / d), s = 31 +
    // ceil(log2 d) (Granlund & Montgomery); unlin3 runs at every refill of the trace kernel
    unsigned m_nz, m_ny;
    int s_nz, s_ny;
};
inline void grid_magic(int d, unsigned *m, int *s) {
    int l = 0;
    while ((1ll << l) < d) ++l;
    *s = 31 + l;
    *m = (unsigned)(((1ull << *s) + (unsigned long long)d - 1) / (unsigned long long)d);
}
inline Grid make_grid(int nx, int ny, int nz) {
    Grid g{nx, ny, nz, 0u, 0u, 0, 0};
    grid_magic(nz, &g.m_nz, &g.s_nz);
    grid_magic(ny, &g.m_ny, &g.s_ny);
    return g;
}
__host__ __device__ inline int64_t gsize(const Grid &g) {
    return (int64_t)g.nx * g.ny * g.nz;
}

// step weights 1/|step| ordered by offset k = (ix+1)*9 + (iy+1)*3 + (iz+1)
struct Weights {
    double w[27];
};
struct TGrad {
    double t[9];
};

// device-side counters (one small block of managed-by-hand words)
enum {
    CNT_ROOTS = 0,     // number of maxima found by the stencil pass
    CNT_EDGES = 1,     // edge voxels appended to the work list
    CNT_CHANGED = 2,   // voxels relabelled by the trace kernel
    CNT_CHANGED_LIST = 3,
    CNT_UNDECIDED = 4, // edge_check centre selection: still undecided
    CNT_CENTRES = 5,
    CNT_NEWEDGE = 6,
    CNT_OVERFLOW = 7,  // trace paths that outgrew the register-file path buffer
    CNT_ERROR = 8,     // trace step cap exceeded
    CNT_VACUUM = 9,    // vacuum voxel count
    CNT_ESCAPED = 11,  // trajectories that left the trusted planes of a slab window
    CNT_STEPS = 10,    // trajectory steps taken by the trace kernel (accounting)
    CNT_DEFER = 12,    // edge pass: voxels next to vacuum, classified from a list (edge.cuh)
    CNT_VACSEEN = 13,  // edge pass: nonzero when the label volume holds vacuum voxels
    CNT_ESCLIST = 14,  // slab windows: walks that left the trusted planes, listed for the peer kernel
    CNT_NUM = 16
};

struct ProfRec {
    int fam;
    cudaEvent_t a, b;
};

}  // namespace bdr

struct bdr_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;                 // false after bdr_set_stream
    cudaStream_t copy_stream = nullptr;      // host -> device chunks of bdr_run
    std::vector<cudaEvent_t> chunk_events;
    bdr::Grid g{0, 0, 0, 0u, 0u, 0, 0};
    int64_t N = 0;
    int halo = 0;               // slab windows: extra x planes on each side (0 = periodic grid)
    int64_t own_lo = 0, own_hi = 0;  // owned linear index range
    bool halo_known_current = false;  // the halo planes of known are copies of the owners' planes (native slab loops)
    bool window_fresh = false;  // known holds a full exact classification of the whole window (slab passes)
    int64_t escaped = 0;        // trajectories that left the trusted planes in the last trace

    double *rho[3] = {nullptr, nullptr, nullptr};
    int rho_alias[3] = {-1, 0, 0};  // -1: owns storage (or empty); k: alias of slot k

    int32_t *labels[2] = {nullptr, nullptr};
    int vac_mode = 0;      // how the stencil pass learns the vacuum mask (VAC_* in kernels.cuh)
    double vac_tol = 0.0;
    int64_t vac_count = 0;  // voxels the last bdr_vacuum_assign marked
    bool verify_fixed_point = false;  // BDR_OPT_VERIFY_FIXED_POINT
    int8_t *known = nullptr;
    uint32_t *ebits = nullptr;  // edge pass: 1 bit per voxel, nzw words per (x,y) row
    uint32_t *vbits = nullptr;  // vacuum bits, same layout (second half of the ebits allocation)
    uint32_t *cbits = nullptr;  // scratch bit volume (voxels relabelled by the last trace)
    uint32_t *sbits = nullptr;  // sticky "was ever an edge or next to one" bits (conservative passes)
    uint32_t *eqz = nullptr, *eqy = nullptr, *eqx = nullptr;  // label equality bits (edge.cuh)
    // the equality bits describe label set eq_which as of the last k_label_eq_bits pass; voxels
    // relabelled since then wait in eq_pending and are patched in by k_eq_update (edge.cuh)
    bool eq_valid = false;
    int eq_which = -1;
    int32_t *eq_pending = nullptr;
    int64_t eq_pending_n = 0, eq_pending_cap = 0;
    int32_t *defer = nullptr;   // edge pass: voxels next to vacuum
    int64_t defer_cap = 0;
    int32_t *term = nullptr;    // where each traced voxel's trajectory ended (bader_calc('neargrid') only)
    bool use_term = false;
    int64_t last_changed = 0;   // entries of list2 written by the last trace (slab rounds)
    int nzw = 0;

    int32_t *list = nullptr;   // work list (edge voxels to trace)
    int64_t list_cap = 0;
    int64_t list_n = 0;
    int32_t *list2 = nullptr;  // changed voxels / centres
    int64_t list2_cap = 0;
    int32_t *list3 = nullptr;  // centres (edge_check)
    int64_t list3_cap = 0;
    int32_t *list4 = nullptr;  // slab windows: over-long walks of the peer trace kernel
    int64_t list4_cap = 0;

    int32_t *roots = nullptr;  // voxel index of each maximum, by slot
    int32_t *minidx = nullptr; // first voxel (C order) of each slot's volume
    int32_t *rank = nullptr;   // slot -> volume number
    int64_t slots_cap = 0;
    bool maxima_fresh[2] = {false, false};  // c->roots lists every density maximum the label set's edge pass can meet
    bool seed_f32 = false;         // the last stencil pass was the fp32-ranked seed (seed.cuh)
    int slab_seed_method = 0;      // BDR_OPT_SLAB_SEED_METHOD
    uint32_t *tile_keys = nullptr; // largest density of every stencil tile (resolve order)
    int32_t *tile_order = nullptr;
    unsigned *tile_hist = nullptr;
    double *d_seedw = nullptr;     // fp64 step weights for the seed kernel's exact fallback
    int64_t tiles_cap = 0;
    std::vector<int64_t> maxima;  // [n][3], in volume-number order
    int64_t n_max = 0;

    unsigned long long *d_cnt = nullptr;  // device counters [CNT_NUM]
    unsigned long long *h_cnt = nullptr;  // pinned mirror
    double *d_sums = nullptr;             // device scratch doubles
    int64_t d_sums_cap = 0;

    void *stage = nullptr;  // device staging for narrowed labels / masks
    size_t stage_bytes = 0;
    void *pinned = nullptr;  // pinned host bounce buffer
    size_t pinned_bytes = 0;

    // sharded runs: device pointers of every rank's arrays (CUDA IPC), see kernels.cuh K4p
    void *peer_view = nullptr;      // bdr::PeerView, host copy
    void *slab_comm = nullptr;      // bdr::SlabComm (comm.cuh): the library's own NCCL communicator
    std::vector<void *> ipc_opened;  // pointers to close on destroy

    bool prof = false;
    std::vector<bdr::ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    double prof_ms[BDR_K_COUNT] = {0};
    int64_t prof_n[BDR_K_COUNT] = {0};
    int64_t launches = 0;
    int64_t syncs = 0;            // counter read-backs (host decisions between data-dependent launches)
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    int64_t trace_steps = 0, trace_voxels = 0;
    double dbg_upload_ms = 0;  // BDR_DEBUG: host->device upload + stencil of the last bdr_run
};

namespace bdr {

inline void prof_begin(bdr_ctx *c, int fam) {
    if (!c->prof) return;
    ProfRec r;
    r.fam = fam;
    auto get = [&]() {
        cudaEvent_t e;
        if (!c->pool.empty()) {
            e = c->pool.back();
            c->pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    };
    r.a = get();
    r.b = get();
    cudaEventRecord(r.a, c->stream);
    c->recs.push_back(r);
}
inline void prof_end(bdr_ctx *c, int fam) {
    (void)fam;
    if (!c->prof) return;
    cudaEventRecord(c->recs.back().b, c->stream);
}
// fold finished event pairs into the per-family totals
inline void prof_collect(bdr_ctx *c) {
    if (c->recs.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto &r : c->recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        c->prof_ms[r.fam] += ms;
        c->prof_n[r.fam] += 1;
        c->pool.push_back(r.a);
        c->pool.push_back(r.b);
    }
    c->recs.clear();
}

#define LAUNCH(ctx, fam, kern, grid, block, smem, ...)                     \
    do {                                                                   \
        bdr::prof_begin(ctx, fam);                                         \
        kern<<<grid, block, smem, (ctx)->stream>>>(__VA_ARGS__);           \
        bdr::prof_end(ctx, fam);                                           \
        (ctx)->launches++;                                                 \
        CU(cudaGetLastError());                                            \
    } while (0)

inline int ensure(int32_t **p, int64_t *cap, int64_t want) {
    if (*cap >= want) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    int64_t n = std::max<int64_t>(want + want / 4, 1024);
    CU(cudaMalloc((void **)p, (size_t)n * sizeof(int32_t)));
    *cap = n;
    return 0;
}

inline int ensure_stage(bdr_ctx *c, size_t bytes) {
    if (c->stage_bytes >= bytes) return 0;
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr;
    c->stage_bytes = 0;
    CU(cudaMalloc(&c->stage, bytes));
    c->stage_bytes = bytes;
    return 0;
}

// read the device counters (synchronises the stream)
inline int read_counters(bdr_ctx *c) {
    c->syncs++;
    CU(cudaMemcpyAsync(c->h_cnt, c->d_cnt, sizeof(unsigned long long) * CNT_NUM,
                       cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
inline int zero_counter(bdr_ctx *c, int which) {
    CU(cudaMemsetAsync(c->d_cnt + which, 0, sizeof(unsigned long long), c->stream));
    return 0;
}

}  // namespace bdr

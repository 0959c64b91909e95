/*
 * bader_b200.h -- C ABI of the B200-native Bader hot path (libbader_b200.so).
 *
 * This is the drop-in boundary: every entry point replaces one call that the
 * reference's `Bader` object (pybader/interface.py) makes into its numba layer
 * (pybader/thread_handlers.py, pybader/utils.py, pybader/methods.py,
 * pybader/refinement.py).  The reference file:line each one stands in for is
 * cited on the declaration.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; the message of
 *    the last failure on the calling thread is bdr_last_error().
 *  - there is NO CPU fallback: without a CUDA device every compute entry fails.
 *  - arrays are C-contiguous [x][y][z], z fastest (io/vasp.py:102-103).
 *  - host buffers are borrowed for the duration of the call only.
 *  - labels on the device are int32: -1 vacuum, >= 0 volume number.
 *  - dist_mat is the reference's 3x3x3 table verbatim (interface.py:242-259;
 *    index 2 on an axis is the step -1), T_grad its 3x3 (interface.py:285-290).
 *  - one handle = one device = one stream; calls on a handle must not overlap.
 *  - the file entry points at the end (bdr_parse_text, bdr_format_grid and their
 *    helpers) stand in for the number blocks of the reference's readers and
 *    writers (pybader/io/vasp.py, pybader/io/cube.py) and take no handle;
 *    bdr_format_grid, bdr_parse_token_host and bdr_host_free need no device.
 */
#ifndef BADER_B200_H
#define BADER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bdr_ctx bdr_ctx;

enum { BDR_METHOD_ONGRID = 0, BDR_METHOD_NEARGRID = 1 };   /* methods.py:12 */
enum { BDR_MODE_ALL = 0, BDR_MODE_CHANGED = 1 };           /* thread_handlers.py:201 */
enum { BDR_RHO_REFERENCE = 0, BDR_RHO_CHARGE = 1, BDR_RHO_SPIN = 2 };
enum { BDR_LABELS_BADER = 0, BDR_LABELS_ATOMS = 1 };

/* kernel families for bdr_profile_get (CUDA-event time on the handle's stream) */
enum {
    BDR_K_VACUUM = 0,      /* vacuum mask + sums                      */
    BDR_K_STENCIL = 1,     /* 27-point stencil -> ascent pointers     */
    BDR_K_RESOLVE = 2,     /* pointer jumping to root codes           */
    BDR_K_RELABEL = 3,     /* slot -> volume-number LUT pass          */
    BDR_K_EDGE_FLAG = 4,   /* edge candidates from the equality bits  */
    BDR_K_EDGE_DILATE = 5, /* near-edge dilation + compaction         */
    BDR_K_TRACE = 6,       /* neargrid trajectory re-trace of edges   */
    BDR_K_EDGE_CHECK = 7,  /* 'changed'-mode incremental reclassify   */
    BDR_K_CHARGE_SUM = 8,  /* per-volume charge / voxel-count sums    */
    BDR_K_ASSIGN = 9,      /* maxima -> atom, label -> atom LUT pass  */
    BDR_K_SURFACE = 10,    /* min distance atom -> its surface        */
    BDR_K_NARROW = 11,     /* int32 -> int8/16/64 staging for D2H     */
    BDR_K_SYNTH = 12,      /* synthetic density generator             */
    BDR_K_FIRST = 13,      /* first-voxel (numbering) pass            */
    BDR_K_EDGE_CONFIRM = 14, /* edge candidates -> edges (density test) */
    BDR_K_TRACE_PEER = 15, /* sharded runs: walks continued on other ranks' memory */
    BDR_K_EDGE_EQ = 16,    /* label equality bits: the streaming R 4 B/voxel half of the edge pass */
    BDR_K_COUNT = 17
};

const char *bdr_last_error(void);
int bdr_version(void);
int bdr_device_count(int *count);

/* ---- lifecycle ---------------------------------------------------------- */
/* Allocates device state for an nx*ny*nz periodic grid on `device`.         */
int bdr_create(int device, int64_t nx, int64_t ny, int64_t nz, bdr_ctx **out);
int bdr_destroy(bdr_ctx *ctx);
int bdr_synchronize(bdr_ctx *ctx);

/* ---- data movement (host <-> device) ------------------------------------ */
/* which = BDR_RHO_*.  Bader.reference / .density / .spin (interface.py:136-137,
 * 203-213).  Slots never uploaded alias BDR_RHO_REFERENCE.                  */
int bdr_upload_density(bdr_ctx *ctx, int which, const double *host);
int bdr_download_density(bdr_ctx *ctx, int which, double *host);
/* Make slot `which` an alias of slot `of` (e.g. density is reference).      */
int bdr_alias_density(bdr_ctx *ctx, int which, int of);
/* Device-to-device copy of slot `src` into slot `dst` (Bader.reference = another
 * density that is already resident, interface.py:136-137 and the -ref flow of
 * entry_points.py:184-194).                                                  */
int bdr_copy_density(bdr_ctx *ctx, int dst, int src);
/* Labels in/out in any of the reference's label dtypes (jits.py:9: int8/16/
 * 32/64); elem_size in bytes.  Narrowing happens on the device
 * (utils.dtype_change, utils.py:256-259).                                   */
int bdr_upload_labels(bdr_ctx *ctx, int which, const void *host, int elem_size);
int bdr_download_labels(bdr_ctx *ctx, int which, void *host, int elem_size);
int bdr_download_known(bdr_ctx *ctx, int8_t *host);
int bdr_clear_labels(bdr_ctx *ctx, int which);

/* ---- hot path ----------------------------------------------------------- */
/* utils.vacuum_assign (utils.py:383-401), called by Bader.volumes_init
 * (interface.py:449-469): reference <= tol -> label -1; returns
 * charge = (sum density)*voxel_volume and volume = count*voxel_volume.      */
int bdr_vacuum_assign(bdr_ctx *ctx, double vac_tol, double voxel_volume,
                      int which_density, double *vac_charge, double *vac_volume);
/* number of voxels the last bdr_vacuum_assign on this handle labelled -1 (0: the
 * host copy of the labels is still current and needs no download)             */
int bdr_vacuum_count(bdr_ctx *ctx, int64_t *count);

/* thread_handlers.bader_calc (thread_handlers.py:15-75) -> methods.ongrid /
 * methods.neargrid (methods.py:15-219, 222-611).  Consumes the BADER labels
 * (0 = to do, -1 = vacuum) and leaves 0-based volume numbers, numbered by the
 * first voxel (C order) of each volume like the reference's discovery order.
 * neargrid: see DESIGN.md -- labels are the refinement fixed point.         */
int bdr_bader_calc(bdr_ctx *ctx, int method, const double *dist_mat,
                   const double *T_grad, int64_t *n_maxima);
/* voxel indices of the maxima, int64[n][3] (the reference's bader_max).     */
int bdr_get_maxima(bdr_ctx *ctx, int64_t *out, int64_t cap);

/* thread_handlers.refine (thread_handlers.py:128-236) -> refinement.edge_find
 * / refinement.neargrid / refinement.edge_check (refinement.py:326, 17, 409)
 * on label set `which`.  iters < 0 = until nothing changes.  history receives
 * (edges, changed) per iteration, up to hist_cap pairs; iters_run the count. */
int bdr_refine(bdr_ctx *ctx, int which, int mode, int64_t iters,
               const double *dist_mat, const double *T_grad,
               int64_t *iters_run, int64_t *history, int64_t hist_cap);

/* refinement.edge_find (refinement.py:326-405): fills the known array from
 * label set `which`; returns the number of edge voxels.                     */
int bdr_edge_find(bdr_ctx *ctx, int which, int64_t *edges);

/* utils.charge_sum (utils.py:236-252), called by Bader.sum_volumes
 * (interface.py:492-525).  charge[l] = voxel_volume * sum density,
 * volume[l] = voxel_volume * count, for labels 0 <= l < n.                  */
int bdr_charge_sum(bdr_ctx *ctx, int which_labels, int which_density,
                   double voxel_volume, int64_t n, double *charge, double *volume);

/* thread_handlers.assign_to_atoms (thread_handlers.py:78-125) ->
 * utils.atom_assign (utils.py:186-232) + utils.volume_assign (utils.py:405-421):
 * nearest atom over 27 images per maximum, then ATOMS labels = LUT(BADER).   */
int bdr_assign_atoms(bdr_ctx *ctx, const double *maxima_cart, int64_t n_max,
                     const double *atoms_cart, int64_t n_atoms, const double *lattice,
                     int64_t *bader_atoms, double *bader_distance);

/* thread_handlers.surface_distance (thread_handlers.py:239-297) ->
 * refinement.edge_find + utils.surface_dist (utils.py:321-379) on label set
 * `which` (the reference passes atoms_volumes).  found = 0 when there is no
 * edge (the reference returns None).                                        */
int bdr_surface_distance(bdr_ctx *ctx, int which, const double *lattice,
                         const double *atoms_cart, int64_t n_atoms,
                         double *distance, int *found);

/* utils.volume_mask (utils.py:462-476): out = density where label == vol_num */
int bdr_volume_mask(bdr_ctx *ctx, int which_labels, int which_density,
                    int64_t vol_num, double *host_out);

/* ---- one-shot, host buffers in / host buffers out ----------------------- */
/* Bader.__call__ stages volumes_init -> bader_calc -> refine_volumes ->
 * sum_volumes(bader=True) (interface.py:405-410) in one call: uploads the
 * density, runs the device pipeline, downloads labels narrowed to
 * label_elem_size.  vac_tol = NaN means no vacuum.  charge/volume may be
 * NULL; maxima receives up to max_cap int64[3] rows.                         */
int bdr_run(bdr_ctx *ctx, const double *host_density, double vac_tol,
            double voxel_volume, int method, int refine_mode, int64_t refine_iters,
            const double *dist_mat, const double *T_grad, void *host_labels,
            int label_elem_size, int64_t *n_maxima, int64_t *maxima, int64_t max_cap,
            double *charge, double *volume);

/* ---- measurement --------------------------------------------------------- */
int bdr_profile_enable(bdr_ctx *ctx, int on);
int bdr_profile_reset(bdr_ctx *ctx);
/* total CUDA-event milliseconds and launch count of one kernel family       */
int bdr_profile_get(bdr_ctx *ctx, int family, double *ms, int64_t *launches);
/* number of kernels this library launched on the handle since creation      */
int bdr_launch_count(bdr_ctx *ctx, int64_t *launches);
/* number of device-counter read-backs (host decisions between data-dependent launches,
 * each a stream synchronisation) on the handle since creation                       */
int bdr_sync_count(bdr_ctx *ctx, int64_t *syncs);
/* CUDA-event stopwatch on the handle's stream (the stream every kernel of
 * this library is launched on): start records an event, stop records a second
 * one, synchronises and returns the elapsed milliseconds.                   */
int bdr_timer_start(bdr_ctx *ctx);
int bdr_timer_stop(bdr_ctx *ctx, double *ms);
/* trajectory steps taken by the trace kernel since the last profile reset   */
int bdr_trace_steps(bdr_ctx *ctx, int64_t *steps, int64_t *voxels);

/* ---- synthetic inputs (bench / tests) ------------------------------------ */
/* separable Gaussian superposition for orthorhombic cells: rho[i][j][k] =
 * sum_a tx[a][i]*ty[a][j]*tz[a][k]; tables are host float64, amplitude folded
 * into tx.  Writes slot `which` on the device.                              */
int bdr_synth_separable(bdr_ctx *ctx, int which, const double *tx, const double *ty,
                        const double *tz, int64_t n_atoms);
/* general cell: sum over atoms and 27 images of amp*exp(-r^2/(2 sigma^2))    */
int bdr_synth_general(bdr_ctx *ctx, int which, const double *lattice,
                      const double *frac_atoms, const double *amps,
                      const double *sigmas, int64_t n_atoms);

/* ---- sharded runs: one slab handle per rank ---------- */
/* A slab handle holds nx_window = owned + 2*halo x planes (y, z stay periodic).
 * It stands for the reference's brick decomposition (thread_handlers.py:27-47):
 * trajectories leaving the brick become provisional "exit" labels
 * (methods.py:170-199) that the host resolves across ranks (utils.edge_assign,
 * utils.py:263-280) -- here over NCCL.  All pointers below marked dev_ are
 * device pointers (e.g. torch tensors' data_ptr()).                         */
int bdr_slab_create(int device, int64_t nx_window, int64_t ny, int64_t nz, int halo,
                    bdr_ctx **out);
/* stencil + local pointer jumping on the window; leaves slot codes -2-s in the
 * BADER labels: s < exit_base are exit-plane voxels (plane 0: s = y*nz+z,
 * plane nx_window-1: s = ny*nz + y*nz+z), s >= exit_base are local maxima.  */
int bdr_slab_seed(bdr_ctx *ctx, const double *dist_mat, int64_t *n_real, int64_t *exit_base);
/* window-linear voxel index of each local maximum, int32[n_real]            */
int bdr_slab_roots(bdr_ctx *ctx, int32_t *host_out, int64_t cap);
/* first owned voxel (window-linear index, 0x7f7f7f7f if none) of every slot  */
int bdr_slab_first_voxel(bdr_ctx *ctx, int64_t n_slots, int32_t *dev_out);
/* slot codes -> global volume numbers through dev_rank[slot]                 */
int bdr_slab_apply_rank(bdr_ctx *ctx, const int32_t *dev_rank);
/* renumbering once the labels are final (volumes are numbered by their first voxel in
 * C order, the reference's discovery order; utils.volume_offset utils.py:497-510):
 * first owned voxel (window-linear, 0x7f7f7f7f if none) of every volume number, and
 * labels >= 0 -> dev_lut[label] in place over the whole window                  */
int bdr_slab_first_voxel_labels(bdr_ctx *ctx, int64_t n_labels, int32_t *dev_out);
int bdr_slab_relabel(bdr_ctx *ctx, int which, const int32_t *dev_lut);
/* one full edge pass / one Jacobi trace launch over the owned edge voxels;
 * escaped counts trajectories that left the trusted planes of the window    */
int bdr_edge_pass(bdr_ctx *ctx, int which, int64_t *edges);
int bdr_trace_pass(bdr_ctx *ctx, int which, const double *dist_mat, const double *T_grad,
                   int64_t *changed, int64_t *escaped);
/* the same pass keeping the list of relabelled voxels for bdr_slab_ec_* (want_list != 0) */
int bdr_trace_pass_list(bdr_ctx *ctx, int which, const double *dist_mat, const double *T_grad,
                        int want_list, int64_t *changed, int64_t *escaped);
/* 'changed'-mode refinement across slabs: refinement.edge_check (refinement.py:409-508) cut
 * into the phases between which the ranks exchange the halo planes of the known array
 * (device pointer: bdr_device_ptr(ctx, 5)).  begin: the relabelled voxels of the last
 * bdr_trace_pass_list that are still edges and maxima are centres at once.  round: one
 * step of the centre selection in GLOBAL scan order over the owned relabelled voxels;
 * repeat (exchanging known halos) until no rank reports undecided voxels.  finish (after
 * a last exchange): re-classify the 27-neighbourhoods of all centres of the window,
 * dilate, queue the new edges for the next trace; returns the new edges this rank owns. */
int bdr_slab_ec_begin(bdr_ctx *ctx, int which);
int bdr_slab_ec_round(bdr_ctx *ctx, int64_t *undecided);
int bdr_slab_ec_finish(bdr_ctx *ctx, int which, int64_t *edges_owned);
/* The round loops of a sharded run inside the library (one host decision per round instead
 * of a Python protocol step per kernel): the library's own NCCL communicator over the slab
 * ring (libnccl.so.2 is bound at run time).  comm_id: 128 bytes from ncclGetUniqueId on one
 * rank, to be broadcast by the caller; comm_init: collective over all ranks.
 * exchange: halo planes of label set 0 / 1 or of the known array (2) from their owners.
 * rounds: bader_calc('neargrid') after the seed is numbered -- conservative first pass,
 * then rounds around the relabelled voxels (own and the neighbours' on the adjacent halo
 * planes) until nothing changes anywhere; history gets (edges or queued, changed) per round
 * as global counts.  refine: thread_handlers.refine (thread_handlers.py:128-236) in mode
 * BDR_MODE_ALL / BDR_MODE_CHANGED with the same history as bdr_refine on one GPU.       */
int bdr_slab_comm_id(void *id_out_128_bytes);
int bdr_slab_comm_init(bdr_ctx *ctx, int world, int rank, const void *id_128_bytes);
int bdr_slab_exchange(bdr_ctx *ctx, int what);
int bdr_slab_rounds(bdr_ctx *ctx, int which, const double *dist_mat, const double *T_grad,
                    int64_t max_passes, int64_t *history, int64_t hist_cap, int64_t *n_hist,
                    int *settled);
int bdr_slab_refine(bdr_ctx *ctx, int which, int mode, int64_t iters, const double *dist_mat,
                    const double *T_grad, int64_t *iters_run, int64_t *history, int64_t hist_cap);
/* run every kernel and copy of this handle on `stream` (a cudaStream_t, e.g. the stream
 * the caller's NCCL plumbing is ordered on) instead of the handle's own stream          */
int bdr_set_stream(bdr_ctx *ctx, void *stream);

/* bader_calc('neargrid') of a sharded run, cut where the ranks have to meet:
 * first_pass = full edge classification (starts the conservative interior
 * bits); trace = one Jacobi trace of the queued voxels (want_list keeps the
 * relabelled voxels for the next requeue); requeue = queue the edges next to
 * the voxels relabelled by the last trace plus `dev_extra` (window-linear
 * indices of halo voxels a neighbour relabelled), a masked streaming pass when
 * they are many, list-based gathers when they are few.                      */
int bdr_slab_first_pass(bdr_ctx *ctx, int which, int64_t *edges);
int bdr_slab_trace(bdr_ctx *ctx, int which, const double *dist_mat, const double *T_grad,
                   int want_list, int64_t *changed);
int bdr_slab_requeue(bdr_ctx *ctx, int which, const int32_t *dev_extra, int64_t n_extra,
                     int64_t *queued);

/* One rank's share of thread_handlers.surface_distance (thread_handlers.py:239-297): exact
 * edge pass on label set `which` of the window (halo labels must be current), then per atom
 * the smallest SQUARED distance to an owned edge voxel of its volume (voxel positions are
 * global), seen[a] != 0 iff the rank owns such a voxel, and the number of owned edges.  The
 * caller reduces: min over ranks, sqrt; no edge anywhere -> the reference returns None.     */
int bdr_slab_surface_distance(bdr_ctx *ctx, int which, const double *lattice, const double *atoms_cart,
                              int64_t n_atoms, double *best_sq, int64_t *seen, int64_t *edges_owned);

/* Trajectories that leave a rank's window continue on the owning rank's memory
 * (CUDA IPC mappings over NVLink / NVSwitch) instead of needing deep halos:
 * export writes three 64-byte IPC handles (density, labels, known); attach
 * receives all ranks' handles (world * 3 * 64 bytes, rank major), the global
 * slab bounds (world + 1 first planes) and the global nx.  Every rank must have
 * finished its edge pass before any rank starts a trace pass.              */
int bdr_slab_ipc_export(bdr_ctx *ctx, void *handles);
int bdr_slab_ipc_attach(bdr_ctx *ctx, int world, int rank, const void *all_handles,
                        const int64_t *bounds, int64_t nx_global);

/* ---- text -> grid: the numeric block of a CHGCAR / cube file ---
 * Replaces the token-by-token numpy conversion of io/vasp.py:90-137 (charge and
 * spin blocks; values / cell volume; file order x fastest -> [x][y][z]) and
 * io/cube.py:99-113 (values * bohr^-3, file order already [x][y][z]).
 * `text` holds whitespace-separated decimal tokens; the first n_values = nx*ny*nz
 * of them are converted (correctly rounded, like Python's float) on the GPU and
 * written to `out` in C order [nx][ny][nz].  x_fastest: token t is voxel
 * (t % nx, (t / nx) % ny, t / (nx * ny)).  op: 0 none, 1 out = v / operand,
 * 2 out = v * operand.  Tokens the device cannot convert exactly (more than 19
 * digits, '****', nan, exponents beyond its tables, subnormal results) leave NaN
 * in `out` and are reported as (token index, byte offset, byte length) triples:
 * the caller converts those with its own strtod / float() and applies `op`
 * (reference behaviour, including its ValueError on junk).                    */
int bdr_parse_text(int device, const char *text, int64_t nbytes, int64_t n_values, int64_t nx,
                   int64_t ny, int64_t nz, int x_fastest, int op, double operand, double *out,
                   int64_t *tokens_found, int64_t *bytes_consumed, int64_t *n_fallback,
                   int64_t *fallback, int64_t fallback_cap);
/* grid -> text: appends the numeric block of a CHGCAR (io/vasp.py:245-258) or of a
 * cube file (io/cube.py:215-222) to `path`, formatted like utils.python_format
 * (utils.py:85-94): every value " %.{prec}E" (sign_space: " % .{prec}E"), per_line
 * values per line, a line break at the end of every row of row_len values.
 * data is C order [nx][ny][nz]; x_fastest writes it in CHGCAR order (x fastest) as one
 * row of nx*ny*nz values.  Host code on all host threads (no device needed).      */
int bdr_format_grid(const char *path, const double *data, int64_t nx, int64_t ny, int64_t nz,
                    int x_fastest, int64_t row_len, int per_line, int prec, int sign_space);
/* bdr_parse_text keeps its device buffers between calls; this frees them     */
int bdr_parse_release(int device);
/* page-locked host memory for the readers' text and result buffers (numpy's
 * pageable arrays cross PCIe at a fifth of the rate); NULL on failure         */
void *bdr_host_alloc(int64_t bytes);
int bdr_host_free(void *ptr);
/* 64-bit content hash of a host buffer on `threads` host threads (<= 0: all), and
 * whether every byte is zero.  The reference passes numpy arrays between its stages
 * (interface.py:449-534) and always reads the array it is given; the Python session
 * keeps device copies and keys them on this hash of the WHOLE array, so an in-place
 * edit between stages re-uploads instead of computing on stale device data.  No
 * device needed.                                                               */
int bdr_host_hash(const void *data, int64_t nbytes, int threads, uint64_t *hash, int *all_zero);
/* the same conversion for one token on the CPU (tests; no device needed):
 * 0 converted, 2 not handled exactly (ask strtod)                             */
int bdr_parse_token_host(const char *token, int64_t len, double *out);

/* ---- options ------------------------------------------------------------- */
/* BDR_OPT_VERIFY_FIXED_POINT (default 0): bader_calc('neargrid') drives the
 * labels to quiescence with one full edge pass plus incremental rounds; with
 * 1 it additionally repeats full passes until one changes nothing, so the
 * labels are certified to be a fixed point of the reference's refinement
 * iteration before any refine() call.                                       */
/* BDR_OPT_SLAB_SEED_METHOD (default BDR_METHOD_ONGRID): which method the next
 * bdr_slab_seed serves.  'ongrid' needs the bit-exact fp64 argmax of
 * methods.py:87-117; 'neargrid' only needs an ascending pointer field with the
 * same maxima and takes the cheaper fp32-ranked stencil (as bdr_bader_calc
 * does on one GPU, so sharded and single-GPU runs seed identically).        */
enum { BDR_OPT_VERIFY_FIXED_POINT = 0, BDR_OPT_SLAB_SEED_METHOD = 1 };
int bdr_set_option(bdr_ctx *ctx, int option, int64_t value);

/* self test: the trace kernel divides three gradient components by one
 * maximum through a shared reciprocal; this checks that sequence against the
 * hardware IEEE division on n pseudo-random operand pairs                    */
int bdr_selftest_div(bdr_ctx *ctx, int64_t n, uint64_t seed, int64_t *mismatches);

/* device pointers, for torch.distributed halo plumbing in the sharded path  */
int bdr_device_ptr(bdr_ctx *ctx, int what, void **ptr);

#ifdef __cplusplus
}
#endif
#endif

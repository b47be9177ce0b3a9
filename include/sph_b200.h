/* sph_b200.h — C ABI of libsphb200.so: the WCSPH per-timestep hot path of
 * colonelwatch/pi-sph-fluid (pi_sph_fluid.c) on NVIDIA B200 (sm_100a).
 *
 * Plain C: opaque handle, plain pointers and sizes, int return codes (0 = ok, negative =
 * SPHB_E_*), no exceptions, no hidden host threads, one CUDA stream per handle.  There is
 * NO CPU fallback: every compute entry point returns SPHB_E_CUDA when no sm_100 device is
 * usable.
 *
 * Two tiers:
 *   1. resident tier  sphb_*      state lives in HBM between calls; what a host loop uses.
 *   2. compat tier    the seven reference-named operators with the reference's exact
 *                     signatures (upload -> kernels -> download per call), so the
 *                     reference's main() links against this library unchanged.
 *
 * Each entry point cites the reference lines (pi_sph_fluid.c:NNN) it replaces.
 */
#ifndef SPH_B200_H
#define SPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

#define SPHB_VERSION 1

enum {
    SPHB_OK = 0,
    SPHB_E_ARG = -1,      /* bad argument / call order                      */
    SPHB_E_CUDA = -2,     /* CUDA runtime error or no usable sm_100 device  */
    SPHB_E_NOMEM = -3,
    SPHB_E_STATE = -4,    /* e.g. step before upload                        */
    SPHB_E_COMM = -5      /* multi-GPU transport error                      */
};

/* Host-side particle record == `struct particle`, pi_sph_fluid.c:26-31 (7 x f32, 28 B). */
typedef struct sphb_particle {
    float x, y, u, v;
    float m;
    float rho;
    float p;
} sphb_particle;

/* The reference's compile-time constants (pi_sph_fluid.c:10-21) as runtime fields, plus
 * the grid geometry its main() passes to alloc_neighbors_context (:595-597). */
typedef struct sphb_params {
    float R;            /* :11  initial spacing                                   */
    float H;            /* :12  smoothing length, R*1.3f                          */
    float width;        /* :13                                                    */
    float height;       /* :14                                                    */
    float rho0;         /* :15  1000                                              */
    float c0;           /* :16  400                                               */
    float g;            /* :17  9.81 (used only by sphb_gravity_from_raw)         */
    float dt;           /* :19  1.0f*H/C                                          */
    float vol;          /* :20  0.57f*H*H                                         */
    float x_min, x_max, y_min, y_max;   /* :595                                   */
    float cell_length;  /* :596 2*H; also the search radius (:144)                */
    int deterministic;  /* 1: inside a cell particles are kept in ascending original
                              index, i.e. the reference's list order (:110-123), so sums
                              run in the reference's order and results are run-to-run
                              reproducible.  0: in-cell order is whatever the atomic
                              counting sort produced.                              */
    int device;         /* CUDA device ordinal                                    */
    int fast_force;     /* 0 (default): calculate_accelerations (:303-373) in the reference's own
                              arithmetic — float/double types, operation order and roundings of
                              :317-337, :52-62, :219-228 — so du_dt, dv_dt are bit-identical with
                              the reference's source semantics (deterministic = 1).
                           1: single precision throughout with approximate rsqrt / reciprocal
                              and folded constants: ~1.8x faster force pass, acceleration within
                              1e-4*max(|a|, G) of the reference up to ~1M particles only (the
                              artificial-pressure pair terms grow like 1/H and so does their
                              rounding noise: 3.6e-4*G at 4M particles).                  */
    int reserved[5];
} sphb_params;

typedef struct sphb_stats {
    double mass;            /* sum m                                              */
    double mom_x, mom_y;    /* sum m*u, sum m*v                                   */
    double kinetic;         /* 0.5 * sum m*(u^2+v^2)                              */
    float max_speed;        /* :667-671                                           */
    float max_rho_err;      /* max(rho - rho0) — the quantity :657-659 means to compute */
    float last_rho_err_ref; /* what :657-659 actually computes     */
    float min_rho, max_rho;
    unsigned int n_escaped;         /* particles binned by clamping (reference: UB, :111-116) */
    unsigned int max_cell_count;    /* largest cell population at the last build    */
    unsigned int n_fluid, n_boundary; /* on a slab context: particles this rank owns / keeps */
    unsigned long long steps;
    unsigned int n_lost;            /* slabs: particles that moved > 2 cell columns in one step  */
    unsigned int n_overflow;        /* slabs: halo message or slot capacity exceeded (fatal);    */
                                    /*   bit 30: a neighbour's message did not arrive in 20 s —  */
                                    /*   sphb_synchronize, sphb_get_stats, sphb_step_stats[_end] */
                                    /*   and sphb_mg_download then return SPHB_E_COMM            */
    /* n_escaped and max_cell_count describe the LAST grid build of the fluid (not a running total).   */
} sphb_stats;

typedef struct sphb_ctx sphb_ctx;

/* ---- resident tier -------------------------------------------------------------- */

/* Fills *prm exactly as the reference's #defines evaluate (:11-21, :595-596) for the
 * given spacing and tank; deterministic = 1, device = 0. */
int sphb_default_params(sphb_params *prm, float R, float width, float height);

int sphb_create(const sphb_params *prm, sphb_ctx **out);
int sphb_destroy(sphb_ctx *ctx);

/* Replaces the malloc'd host arrays of :491-493, :519 as the simulation state.  Copies
 * both arrays to HBM (AoS -> SoA).  boundary may be NULL / n_boundary 0. */
int sphb_upload(sphb_ctx *ctx, const sphb_particle *fluid, int n_fluid,
                const sphb_particle *boundary, int n_boundary);

/* :600-601  update_neighbors_context(ctx_boundary) + calculate_boundary_pseudomass */
int sphb_init_boundary(sphb_ctx *ctx);

/* :604-607  update_neighbors_context(ctx_fluid) + calculate_density +
 * calculate_particle_pressure + calculate_accelerations on the state as it is */
int sphb_compute_accel(sphb_ctx *ctx, float gravity_x, float gravity_y);

/* Restores the caller-owned du_dt/dv_dt arrays (:492-493) of a saved state (original order),
 * so a run can continue from a checkpoint exactly where :612 would: the next kick uses them. */
int sphb_upload_accel(sphb_ctx *ctx, const float *du_dt, const float *dv_dt);

/* :612-641  nsteps iterations of the leapfrog body (kick, drift, grid rebuild, density,
 * pressure, accelerations, kick) with constant gravity.  Asynchronous on the handle's
 * stream; sphb_download / sphb_stats / sphb_synchronize wait for it. */
int sphb_step(sphb_ctx *ctx, float gravity_x, float gravity_y, int nsteps);

/* Same, reading one (gx, gy) pair per step from host memory — the per-step value the
 * reference's OpenMP threads read from `g` at :632. */
int sphb_step_trace(sphb_ctx *ctx, const float *gravity_xy, int nsteps);

/* One iteration of the reference's loop body INCLUDING its statistics scans (:612-675), nsteps
 * times: like sphb_step_trace, and the statistics of the state after the last step are returned in
 * *out.  They are reduced by the force pass of that step itself and delivered through mapped pinned
 * host memory, so this call returns when the step is done without any further copy or kernel;
 * the values are those sphb_get_stats would return.  nsteps >= 1. */
int sphb_step_stats(sphb_ctx *ctx, const float *gravity_xy, int nsteps, sphb_stats *out);

/* The same in two halves, for a host loop that wants the GPU to start step s + 1 while the statistics
 * of step s are still on their way (the reference hands each frame to a display thread in the same
 * spirit, :558-579): _begin launches the steps and returns a ticket at once, _end waits for that
 * ticket's statistics.  At most two tickets may be outstanding; they are collected in order. */
int sphb_step_stats_begin(sphb_ctx *ctx, const float *gravity_xy, int nsteps, unsigned long long *ticket_out);
int sphb_step_stats_end(sphb_ctx *ctx, unsigned long long ticket, sphb_stats *out);

/* Copies the state back in ORIGINAL particle order.  Any pointer may be NULL. */
int sphb_download(sphb_ctx *ctx, sphb_particle *fluid_out, float *du_dt, float *dv_dt);
int sphb_download_boundary(sphb_ctx *ctx, sphb_particle *boundary_out);

/* :649 / :380-411  draw_metaballs into the SSD1306 page layout (1024 bytes). */
int sphb_render(sphb_ctx *ctx, unsigned char *draw_buffer);

/* The frame for LARGE particle counts.  draw_metaballs (:380-411) is tied to the reference's scale: its condition
 * divides by W(px_width/2) (:401), and once the pixel (WIDTH/128 = 31 mm) is wider than the kernel support that is
 * W far outside 2H, where the reference's W — no cut-off, :45-50 — is a growing polynomial.  For such scenes
 * (px_width > 4H, i.e. R below ~6 mm; BASELINE configs[1]-[4]) the frame is a splat: a pixel is lit when the fluid
 * volume inside it, (particles in the pixel) * V (:20), covers at least half of the pixel; same 1 KiB SSD1306 page
 * layout (:407-408).  sphb_render_counts gives the per-pixel counts of the particles this context owns (a slab
 * run adds the ranks' counts, e.g. with an all-reduce, before sphb_splat_frame). */
int sphb_render_splat(sphb_ctx *ctx, unsigned char *draw_buffer);
int sphb_render_counts(sphb_ctx *ctx, unsigned int *counts /* 64 x 128, row 0 = top */);
int sphb_splat_frame(const sphb_params *prm, const unsigned int *counts, unsigned char *draw_buffer /* 1024 */);

/* :656-675 as device reductions, plus conservation sums. */
int sphb_get_stats(sphb_ctx *ctx, sphb_stats *out);

int sphb_synchronize(sphb_ctx *ctx);

/* State file (the reference has none: its arrays live in malloc'd memory until exit): positions,
 * velocities, du_dt/dv_dt, the boundary and the parameters.  sphb_load_state returns a context
 * that continues the run bit-identically (device < 0: the device recorded in the file). */
int sphb_save_state(sphb_ctx *ctx, const char *path);
int sphb_load_state(const char *path, int device, sphb_ctx **out);

/* :439-440  raw MPU6050 counts -> gravity vector */
int sphb_gravity_from_raw(const sphb_params *prm, int accel_x_raw, int accel_y_raw,
                          float *gravity_x, float *gravity_y);

/* ---- parity / inspection helpers (not on the hot path) ---------------------------- */

/* cell index row*m_cells+col (:111-113) of every fluid particle, original order */
int sphb_cell_ids(sphb_ctx *ctx, int *cell_out);
int sphb_grid_shape(sphb_ctx *ctx, int *n_cells_rows, int *m_cells_cols);

/* Neighbour sets as find_neighbors (:126-153) would return them for the current state.
 * which = 0: fluid-fluid, 1: fluid-boundary, 2: boundary-boundary.  For particle i
 * (original index) counts[i] neighbours are written to lists[i*cap ...] as ORIGINAL
 * indices in this library's visiting order (== the reference's order when
 * deterministic = 1).  Returns the number of particles whose count exceeded cap. */
int sphb_neighbor_lists(sphb_ctx *ctx, int which, int cap, int *counts, int *lists);

/* The fluid-fluid neighbour lists of the hot path itself: what the density pass of the last
 * sphb_compute_accel / sphb_step found (find_neighbors, :126-153) and handed to the force pass, decoded from
 * the device representation (tile offsets per chunk of 128 sorted particles + the chunk's staging plan) back
 * to ORIGINAL indices in visiting order.  counts[i] = -1 where no list was handed over for particle i (the
 * force pass searches again for it: list flushed or its part of the chunk not staged).  *n_whole_chunks =
 * chunks whose plan travelled in the chunk record.  Single-GPU contexts. */
int sphb_handover_lists(sphb_ctx *ctx, int cap, int *counts, int *lists, unsigned int *n_whole_chunks);

/* The pair term m_j*temp_ij*grad_a W_ij of calculate_accelerations (:317-337, :52-62, :226) for n
 * caller-given pairs, one device thread per pair — the device instruction sequences of the force pass
 * made testable one pair at a time.  pairs: 12 floats each (x_i y_i x_j y_j | u_i v_i u_j v_j |
 * rho_i p_i/rho_i^2 rho_j p_j/rho_j^2), m_j = the uniform fluid mass; out: (tx, ty) per pair.
 * variant 0: the hot loop's form (packed, exact-division shortcuts), 1: general IEEE divisions,
 * 2: scalar form with the shortcuts, 3: the hot loop's two-neighbours-at-once form (rows 2t and 2t+1 must
 * carry the same particle i; n even); +4 on variants 0-2: j is a boundary particle (:346-365).  *exact_shortcuts says
 * whether the host verified the shortcuts for this context's H (variants 0, 2 are only then exact). */
int sphb_probe_force_pair(sphb_ctx *ctx, int n, const float *pairs, int variant, float *out_txy, int *exact_shortcuts);

/* Grid builds of the fluid set, so far, whose deterministic reorder took the in-cell order of unchanged cells
 * from the previous build (cell marks) instead of ranking every particle by id.  Sets below
 * 2^20 slots do not use the marks unless SPHB_TOUCH_MIN_SLOTS is set in the environment when the set is
 * uploaded (the parity tests run small scenes with SPHB_TOUCH_MIN_SLOTS=0). */
int sphb_reorder_marks(sphb_ctx *ctx, unsigned long long *builds_with_marks);

/* ---- measurement hooks --------------------------------------------------------------- */

enum { SPHB_K_ADVECT_BIN = 0, SPHB_K_SCAN, SPHB_K_REORDER, SPHB_K_DENSITY, SPHB_K_FORCE,
       SPHB_K_OTHER, SPHB_K_COUNT };

/* mode 0: off.  mode 1: CUDA events around every kernel of sphb_step (accumulated). */
int sphb_profile(sphb_ctx *ctx, int mode);
/* Accumulated device milliseconds and launch counts per kernel class since the last
 * reset; pair statistics of the last density pass when collected. */
int sphb_profile_read(sphb_ctx *ctx, double ms[SPHB_K_COUNT], unsigned long long launches[SPHB_K_COUNT], int reset);
/* candidate / accepted fluid-fluid pairs per particle for the current state (SURVEY.md
 * §8d: C and P, "measured by the run") */
int sphb_pair_stats(sphb_ctx *ctx, double *candidates_per_particle, double *accepted_per_particle);
/* Evicts L2 between timed iterations: overwrites a 256 MiB scratch buffer (> the 126 MB L2)
 * on the handle's stream.  Not part of any step. */
int sphb_flush_l2(sphb_ctx *ctx);
/* the cudaStream_t the handle launches on, for event timing by the caller */
void *sphb_stream(sphb_ctx *ctx);
/* number of kernels launched by this handle so far */
unsigned long long sphb_launch_count(sphb_ctx *ctx);
const char *sphb_last_error(void);
const char *sphb_build_info(void);

/* ---- multi-GPU: x-slabs of whole cell columns (new design, the reference is
 * single-process) ---------------------------------------------------------------------------
 *
 * Rank r of `world` owns the global cell columns [cuts[r], cuts[r+1]) of the reference's grid
 * (:93-94, cell = 2H) and keeps two ghost columns either side.  Per step each rank sends ONE
 * message to each neighbour (halo + migrants, 20 B per particle).  Call order per rank:
 *
 *   sphb_create -> sphb_mg_configure -> sphb_mg_connect_nccl [-> sphb_mg_ipc_handle, sphb_mg_connect_ipc]
 *                                      | sphb_mg_connect_local
 *   -> sphb_mg_upload (this rank's particles, global ids) -> sphb_init_boundary
 *   -> sphb_compute_accel / sphb_step / sphb_step_trace        (NCCL or peer-store transport, one process per GPU)
 *    | sphb_mg_group_compute_accel / sphb_mg_group_step       (in-process transport, one host thread)
 *   -> sphb_get_stats (this rank's owned particles) [+ sphb_mg_allreduce_stats | sphb_mg_merge_stats]
 *   -> sphb_mg_download (owned particles with their global ids)
 *
 * In deterministic mode the result is bit-identical to the single-GPU run for any cuts.
 */
#define SPHB_NCCL_ID_BYTES 128

typedef struct sphb_mg_info_t {
    int rank, world;
    int col_lo, col_hi;             /* owned global columns                         */
    int window_lo, window_hi;       /* columns held, ghosts included                */
    int halo_capacity;              /* entries per message                          */
    int particle_capacity;          /* particle slots on this rank                  */
    int transport;                  /* 1 NCCL, 2 in-process peer stores, 3 peer stores across processes (CUDA IPC) */
    unsigned long long message_bytes, bytes_sent, exchanges;
} sphb_mg_info_t;

/* host-side planning (no GPU work): grid shape (:93-94), the global column of x (:112, clamped),
 * a per-column particle histogram, and cuts at its quantiles (every slab >= min_width columns). */
int sphb_grid_columns(const sphb_params *prm, int *rows, int *cols);
int sphb_column_of(const sphb_params *prm, float x);
int sphb_column_histogram(const sphb_params *prm, const sphb_particle *particles, int n, unsigned long long *hist);
int sphb_mg_plan_cuts(const unsigned long long *hist, int cols, int world, int min_width, int *cuts /* world+1 */);
/* the same with a cost per column on top of its particles, in particle units: what the (possibly empty)
 * cells of a column cost a rank per step — the prefix scan walks every cell of the rank's window, measured
 * 0.034 particle-equivalents per cell, i.e. column_cost = 0.034 * rows — so that a dam break's dry half is
 * not one rank's burden.  column_cost = 0: plain particle-count quantiles. */
#define SPHB_CELL_COST 0.034
int sphb_mg_plan_cuts_cost(const unsigned long long *hist, int cols, int world, int min_width, double column_cost,
                           int *cuts /* world+1 */);

int sphb_mg_configure(sphb_ctx *ctx, int rank, int world, int col_lo, int col_hi,
                      int particle_capacity /* 0: 1.25 n + slack */, int halo_capacity /* 0: 65536 */);
/* NCCL transport: rank 0 makes an id, the host program broadcasts it (MPI, torch.distributed,
 * a file...), every rank connects.  libnccl.so.2 is loaded on first use. */
int sphb_mg_unique_id(char id_out[SPHB_NCCL_ID_BYTES]);
int sphb_mg_connect_nccl(sphb_ctx *ctx, const char id[SPHB_NCCL_ID_BYTES]);
/* Peer-store transport across processes (one process per GPU on one NVLink/NVSwitch box): every rank
 * exports its receive block (sphb_mg_ipc_handle), the host program hands each rank its neighbours'
 * handles (NULL where there is no neighbour), and from then on the advect+bin kernel stores the
 * message entries straight into the neighbour's receive buffer over NVLink; a count + epoch word
 * written with release semantics after the kernel completes the message and the neighbour's binning
 * kernel waits for it on the device.  No NCCL call is left in the step (an NCCL connection made
 * before is kept for sphb_mg_allreduce_stats).  sphb_mg_disconnect_ipc unmaps the neighbours'
 * blocks again: call it on every rank, then synchronise the ranks, before sphb_destroy. */
#define SPHB_IPC_HANDLE_BYTES 64
int sphb_mg_ipc_handle(sphb_ctx *ctx, unsigned char handle_out[SPHB_IPC_HANDLE_BYTES]);
int sphb_mg_connect_ipc(sphb_ctx *ctx, const unsigned char *left_handle, const unsigned char *right_handle);
int sphb_mg_disconnect_ipc(sphb_ctx *ctx);
/* in-process transport: all ranks' contexts live in this process (same or different GPUs) */
int sphb_mg_connect_local(sphb_ctx **ctxs, int n);

/* this rank's particles; ids[i] (or id_base + i when ids is NULL) is the particle's global
 * original index.  The boundary is replicated: pass all of it on every rank. */
int sphb_mg_upload(sphb_ctx *ctx, const sphb_particle *fluid, const uint32_t *ids, uint32_t id_base, int n_fluid,
                   const sphb_particle *boundary, int n_boundary);
/* du_dt/dv_dt of the particles just uploaded (same order): a restart / re-cut continues exactly,
 * without sphb_compute_accel.  Directly after sphb_mg_upload (and sphb_init_boundary in any order). */
int sphb_mg_upload_accel(sphb_ctx *ctx, const float *du_dt, const float *dv_dt);
/* the particles this rank owns now (any order) with their global ids; *n_out = how many */
int sphb_mg_download(sphb_ctx *ctx, int cap, sphb_particle *fluid_out, uint32_t *ids_out, float *du_dt,
                     float *dv_dt, int *n_out);

/* Re-cut of a RUNNING multi-process simulation (a dam break drains the left slabs, so cuts made
 * at t = 0 go stale).  Collective: every rank calls it after the same step.  The ranks' per-column particle
 * counts are summed, every rank plans the same new cuts (quantiles of particles + column_cost per column, as
 * sphb_mg_plan_cuts_cost) and every particle — position, velocity, du_dt/dv_dt, global id — moves to the rank
 * that owns its column now, in one grouped exchange; windows, slots and the windowed boundary are rebuilt and
 * the run continues bit-identically (the peer-store halo transport, if connected, stays as it is).
 * min_imbalance > 0: nothing happens while max/mean of the ranks' particle counts is <= it (e.g. 1.05).
 * *changed_out: 1 when the cuts moved.
 *   sphb_mg_rebalance       moves the bytes with ncclAllReduce / ncclSend / ncclRecv (sphb_mg_connect_nccl first);
 *   sphb_mg_rebalance_host  lets the host program carry them (MPI, gloo, ...): `allreduce` sums `count` 64-bit
 *                           words in place over all ranks; `alltoallv` sends send_bytes[r] bytes at send + send_off[r]
 *                           to rank r and receives recv_bytes[r] bytes from rank r at recv + recv_off[r] (its own
 *                           segment included); both return 0 on success.  Buffers are host memory. */
typedef int (*sphb_mg_allreduce_u64_fn)(void *user, unsigned long long *inout, int count);
typedef int (*sphb_mg_alltoallv_fn)(void *user, const void *send, const unsigned long long *send_bytes,
                                    const unsigned long long *send_off, void *recv, const unsigned long long *recv_bytes,
                                    const unsigned long long *recv_off);
int sphb_mg_rebalance(sphb_ctx *ctx, int min_width, double column_cost, double min_imbalance, int *changed_out);
int sphb_mg_rebalance_host(sphb_ctx *ctx, int min_width, double column_cost, double min_imbalance,
                           sphb_mg_allreduce_u64_fn allreduce, sphb_mg_alltoallv_fn alltoallv, void *user, int *changed_out);

/* State files of a slab run: every rank writes ONE part (its owned particles with their global ids, du_dt/dv_dt,
 * the parameters, the step count, and the whole tank's boundary).  sphb_mg_load_state gives a freshly
 * configured slab context (sphb_create + sphb_mg_configure, any rank count, any cuts) the particles of ALL the
 * parts that fall into its columns, restores their accelerations and the boundary; after sphb_init_boundary the
 * run continues bit-identically — on a different number of ranks if wanted (1 rank: one slab over all columns).
 * Layout: 64-byte header (version 2) | sphb_params | ids[n] | struct particle[n] | du_dt[n] | dv_dt[n] |
 * struct particle boundary[nb] (the CPU checker under tests reads both versions). */
int sphb_mg_save_state(sphb_ctx *ctx, const char *path);
int sphb_mg_load_state(sphb_ctx *ctx, const char *const *paths, int n_paths);

int sphb_mg_group_compute_accel(sphb_ctx **ctxs, int n, float gravity_x, float gravity_y);
/* gravity_xy: NULL (constant gravity_x/y) or nsteps (gx, gy) pairs */
int sphb_mg_group_step(sphb_ctx **ctxs, int n, float gravity_x, float gravity_y, const float *gravity_xy, int nsteps);
int sphb_mg_group_synchronize(sphb_ctx **ctxs, int n);

int sphb_mg_merge_stats(const sphb_stats *per_rank, int n, sphb_stats *out);
int sphb_mg_allreduce_stats(sphb_ctx *ctx, sphb_stats *inout);
int sphb_mg_info(sphb_ctx *ctx, sphb_mg_info_t *out);

/* ---- compat tier: the reference's operator signatures ------------------------------- */

struct particle;            /* == sphb_particle; the reference's name, :26 */
struct neighbors_context;   /* opaque here; the reference's struct is :73-80 */

struct neighbors_context *alloc_neighbors_context(int n_particles, float x_min, float x_max,
                                                  float y_min, float y_max, float cell_length);   /* :82  */
void update_neighbors_context(struct neighbors_context *ctx, struct particle *particles);          /* :104 */
void calculate_boundary_pseudomass(struct particle *boundary, struct neighbors_context *ctx_boundary); /* :242 */
void calculate_density(struct particle *fluid, struct particle *boundary,
                       struct neighbors_context *ctx_fluid, struct neighbors_context *ctx_boundary); /* :263 */
void calculate_particle_pressure(struct particle *particles, int n_particles);                     /* :294 */
void calculate_accelerations(float *du_dt_fluid, float *dv_dt_fluid, struct particle *fluid,
                             struct particle *boundary, struct neighbors_context *ctx_fluid,
                             struct neighbors_context *ctx_boundary, float gravity_x, float gravity_y); /* :303 */
void draw_metaballs(unsigned char *draw_buffer, struct particle *pixel_pseudoparticles,
                    struct particle *fluid, struct neighbors_context *ctx_fluid);                    /* :380 */

/* The reference fixes R/H/RHO_0/C as macros; the compat operators take them from here
 * (default: sphb_default_params(0.075, 4, 2), H re-derived as cell_length/2 per context). */
int sphb_compat_set_params(const sphb_params *prm);
void sphb_compat_free_context(struct neighbors_context *ctx);
/* releases the private context calculate_particle_pressure keeps (its signature carries none); also done at exit */
void sphb_compat_shutdown(void);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* SPH_B200_H */

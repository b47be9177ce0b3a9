/* sph_oracle.h — CPU ORACLE for the WCSPH per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the algorithm in
 * colonelwatch/pi-sph-fluid's pi_sph_fluid.c (reference @ dcf1c71), used as the checker
 * for the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product library (libsphb200.so) never links,
 * loads or calls anything in oracle/.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this
 * oracle is pinned against the reference's OWN translation unit compiled here into
 * oracle/_ref/ (see oracle/Makefile): tests/test_oracle_vs_reference.py requires
 * bit-for-bit agreement on the reference's default scene, and tests/golden/ holds
 * vectors produced by that reference build.
 *
 * Differences from the reference, all deliberate and none changing the arithmetic:
 *   - scene/physics constants are runtime fields (reference: #defines, :10-21);
 *   - particle indices are uint32 with UINT32_MAX as the list terminator
 *     (reference: unsigned short / USHRT_MAX, :78-79, :107);
 *   - the per-particle neighbour buffer is sized by params.max_neighbors and an
 *     overflow is counted instead of smashing the stack (reference: fixed 48, :21,
 *     no check at :144-147);
 *   - out-of-grid cell indices are clamped and counted (reference: UB, :111-116).
 * Build strict (-O2 -fno-fast-math -ffp-contract=off) for parity.
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* pi_sph_fluid.c:26-31 — identical layout, 7 x f32 = 28 bytes */
typedef struct {
    float x, y, u, v;
    float m;
    float rho;
    float p;
} oracle_particle;

/* pi_sph_fluid.c:10-21 made runtime.  Derived fields are computed by
 * oracle_make_params() with the reference's exact expression order and types. */
typedef struct {
    float R;        /* :11 initial spacing                    */
    float H;        /* :12 (R*1.3f)                           */
    float width;    /* :13                                    */
    float height;   /* :14                                    */
    float rho0;     /* :15                                    */
    float c0;       /* :16                                    */
    float g;        /* :17                                    */
    float dt;       /* :19 (1.0f*H/C)                         */
    float vol;      /* :20 (0.57f*H*H)                        */
    float mass;     /* :502 RHO_0*V                           */
    int   max_neighbors;   /* :21, default 48                 */
} oracle_params;

/* pi_sph_fluid.c:73-80 with uint32 links */
typedef struct {
    float x_min, x_max, y_min, y_max;
    float cell_length;
    int n_cells, m_cells;       /* rows, columns — names as in the reference */
    int n_particles;
    uint32_t *cells_head, *cells_tail, *particles_next;
    long long n_clamped;        /* particles whose cell index had to be clamped */
} oracle_grid;

typedef struct {
    long long neighbor_overflows;   /* queries that hit max_neighbors */
    int       max_neighbors_seen;
} oracle_counters;

void oracle_make_params(oracle_params *prm, float R, float width, float height);

/* kernel primitives, :40-62 */
float oracle_euclid_dist(float xi, float yi, float xj, float yj);
float oracle_W(const oracle_params *prm, float xi, float yi, float xj, float yj);
void  oracle_grad_W(const oracle_params *prm, float xi, float yi, float xj, float yj,
                    float *gx, float *gy);

/* neighbour grid, :82-153 */
oracle_grid *oracle_grid_alloc(int n_particles, float x_min, float x_max, float y_min,
                               float y_max, float cell_length);
void oracle_grid_free(oracle_grid *g);
void oracle_grid_update(oracle_grid *g, const oracle_particle *particles);
int  oracle_find_neighbors(const oracle_params *prm, int *j_out, int cap,
                           const oracle_particle *a, const oracle_particle *b, int same_array,
                           int i, const oracle_grid *grid_b, oracle_counters *ctr);
/* cell id of every particle (row*m_cells+col, :111-113), for exact comparison */
void oracle_cell_ids(const oracle_grid *g, const oracle_particle *p, int n, int *cell_out);
/* neighbour set of particle i in reference visiting order; returns count */
int  oracle_neighbor_list(const oracle_params *prm, const oracle_particle *a,
                          const oracle_particle *b, int same_array, int i,
                          const oracle_grid *grid_b, int *j_out, int cap);

/* operators, :242-373 */
void oracle_boundary_pseudomass(const oracle_params *prm, oracle_particle *boundary,
                                const oracle_grid *gb, oracle_counters *ctr);
void oracle_density(const oracle_params *prm, oracle_particle *fluid,
                    const oracle_particle *boundary, const oracle_grid *gf,
                    const oracle_grid *gb, oracle_counters *ctr);
void oracle_pressure(const oracle_params *prm, oracle_particle *particles, int n);
void oracle_accelerations(const oracle_params *prm, float *du_dt, float *dv_dt,
                          const oracle_particle *fluid, const oracle_particle *boundary,
                          const oracle_grid *gf, const oracle_grid *gb, float gx, float gy,
                          oracle_counters *ctr);

/* leapfrog pieces, :615-624 and :637-640 */
void oracle_kick(const oracle_params *prm, oracle_particle *fluid, int n,
                 const float *du_dt, const float *dv_dt);
void oracle_drift(const oracle_params *prm, oracle_particle *fluid, int n);

/* :604-607 — accelerations for the state as given (no advection) */
void oracle_compute_accel(const oracle_params *prm, oracle_particle *fluid, int n_fluid,
                          const oracle_particle *boundary, oracle_grid *gf,
                          const oracle_grid *gb, float gx, float gy, float *du_dt,
                          float *dv_dt, oracle_counters *ctr);
/* :612-641 — nsteps leapfrog steps with constant gravity; gxy may instead give one
 * (gx,gy) pair per step when non-NULL */
void oracle_step(const oracle_params *prm, oracle_particle *fluid, int n_fluid,
                 const oracle_particle *boundary, oracle_grid *gf, const oracle_grid *gb,
                 float gx, float gy, const float *gxy_per_step, int nsteps, float *du_dt,
                 float *dv_dt, oracle_counters *ctr);

/* render, :380-411 and pixel centres :570-577 */
void oracle_pixel_pseudoparticles(const oracle_params *prm, oracle_particle *pixels /*8192*/);
void oracle_draw_metaballs(const oracle_params *prm, unsigned char *draw_buffer /*1024*/,
                           const oracle_particle *pixels, const oracle_particle *fluid,
                           const oracle_grid *gf, oracle_counters *ctr);

/* scenes.  drop = the reference's own (:484-540).  The others are builder-defined
 *  but use the reference's lattice idiom (float accumulation). */
int  oracle_scene_count_drop(const oracle_params *prm);
void oracle_scene_fill_drop(const oracle_params *prm, oracle_particle *fluid);
int  oracle_scene_count_block(const oracle_params *prm, float x0, float x1, float y0, float y1);
void oracle_scene_fill_block(const oracle_params *prm, oracle_particle *fluid, float x0,
                             float x1, float y0, float y1);
int  oracle_scene_count_boundary(const oracle_params *prm);
void oracle_scene_fill_boundary(const oracle_params *prm, oracle_particle *boundary);

/* gravity mapping of the MPU6050 reader, :439-440 */
void oracle_gravity_from_raw(const oracle_params *prm, int accel_x_raw, int accel_y_raw,
                             float *gx, float *gy);

/* FNV-1a-64 over (cell*65536+idx) walking cells in order and each list head->tail
 * (SURVEY.md Appendix B known answer) */
uint64_t oracle_grid_fnv(const oracle_grid *g);

int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif

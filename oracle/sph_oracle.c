/* sph_oracle.c — CPU ORACLE (test infrastructure only; see sph_oracle.h).
 *
 * A restatement, with runtime parameters, of the per-timestep WCSPH path of
 * colonelwatch/pi-sph-fluid, pi_sph_fluid.c.  Every function cites the reference lines
 * whose arithmetic (operand types, evaluation order, summation order) it follows.  The
 * reference writes unsuffixed literals, so several sub-expressions are evaluated in
 * double and rounded to float on assignment; those sites are spelled
 * out with explicit casts here.
 *
 * Compile with -O2 -fno-fast-math -ffp-contract=off for parity work.
 */
#include "sph_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_NIL UINT32_MAX

/* powf with the small integer exponents the reference uses (:49 ^4, :56 ^3, :298 ^7, :325 ^4).
 * Default: glibc powf, as the IEEE-strict build of the reference calls it — this flavour is
 * the one pinned bit-for-bit against oracle/_ref/libpisph_ref_strict.so.
 * -DORACLE_POW_CHAIN: the multiply chains gcc emits for those calls under the reference's
 * own shipped flags (Makefile:2 -Ofast; checked with `gcc -Ofast -S`: x^4=(x*x)*(x*x),
 * x^3=(x*x)*x, x^7=((x*x)*x)*((x*x)*(x*x))).  SURVEY.md §8c allows either choice and asks to
 * say which; the chain flavour is what the GPU kernels can match bit-for-bit, because powf's
 * last bit is a libm implementation detail. */
#ifdef ORACLE_POW_CHAIN
static inline float pow3_(float x) { float x2 = x * x; return x2 * x; }
static inline float pow4_(float x) { float x2 = x * x; return x2 * x2; }
static inline float pow7_(float x) { float x2 = x * x; float x4 = x2 * x2; float x3 = x2 * x; return x3 * x4; }
#else
static inline float pow3_(float x) { return powf(x, 3); }
static inline float pow4_(float x) { return powf(x, 4); }
static inline float pow7_(float x) { return powf(x, 7); }
#endif

/* ------------------------------------------------------------------ parameters */

/* pi_sph_fluid.c:11-21, :502.  All float arithmetic, left to right. */
void oracle_make_params(oracle_params *prm, float R, float width, float height)
{
    prm->R = R;
    prm->H = R * 1.3f;                    /* :12 */
    prm->width = width;
    prm->height = height;
    prm->rho0 = 1000.0f;                  /* :15 */
    prm->c0 = 400.0f;                     /* :16 */
    prm->g = 9.81f;                       /* :17 */
    prm->dt = 1.0f * prm->H / prm->c0;    /* :19 */
    prm->vol = 0.57f * prm->H * prm->H;   /* :20 */
    prm->mass = prm->rho0 * prm->vol;     /* :502 */
    prm->max_neighbors = 48;              /* :21 */
}

/* :46 / :53 — 7/(4*M_PI*H*H): int/double chain, rounded to float once */
static float norm_factor(const oracle_params *prm)
{
    double h = (double)prm->H;
    return (float)(7 / (4 * M_PI * h * h));
}

/* ------------------------------------------------------------------ kernel */

/* :40-43 */
float oracle_euclid_dist(float xi, float yi, float xj, float yj)
{
    float dx = xi - xj, dy = yi - yj;
    return sqrtf(dx * dx + dy * dy);
}

/* :45-50.  No support cutoff here — the neighbour search is the only cutoff. */
float oracle_W(const oracle_params *prm, float xi, float yi, float xj, float yj)
{
    const float nf = norm_factor(prm);
    float q = oracle_euclid_dist(xi, yi, xj, yj) / prm->H;
    float a = 1 - 0.5f * q;
    float b = 1 + 2 * q;
    return nf * pow4_(a) * b;
}

/* :52-62.  r == 0 gives 0/0 = NaN exactly as the reference does. */
void oracle_grad_W(const oracle_params *prm, float xi, float yi, float xj, float yj,
                   float *gx, float *gy)
{
    const float nf = norm_factor(prm);
    float r = oracle_euclid_dist(xi, yi, xj, yj);
    float q = r / prm->H;
    float a = 1 - 0.5f * q;
    float dW_dq = nf * (-5) * q * pow3_(a);
    float dq_dx = (xi - xj) / r / prm->H;
    float dq_dy = (yi - yj) / r / prm->H;
    *gx = dW_dq * dq_dx;
    *gy = dW_dq * dq_dy;
}

/* ------------------------------------------------------------------ grid */

/* :82-102 */
oracle_grid *oracle_grid_alloc(int n_particles, float x_min, float x_max, float y_min,
                               float y_max, float cell_length)
{
    oracle_grid *g = (oracle_grid *)calloc(1, sizeof *g);
    g->x_min = x_min; g->x_max = x_max; g->y_min = y_min; g->y_max = y_max;
    g->cell_length = cell_length;
    g->n_cells = (int)((y_max - y_min) / cell_length) + 1;   /* :93 rows */
    g->m_cells = (int)((x_max - x_min) / cell_length) + 1;   /* :94 cols */
    size_t nc = (size_t)g->n_cells * (size_t)g->m_cells;
    g->cells_head = (uint32_t *)malloc(nc * sizeof(uint32_t));
    g->cells_tail = (uint32_t *)malloc(nc * sizeof(uint32_t));
    g->n_particles = n_particles;
    g->particles_next = (uint32_t *)malloc((size_t)(n_particles > 0 ? n_particles : 1) * sizeof(uint32_t));
    return g;
}

void oracle_grid_free(oracle_grid *g)
{
    if (!g) return;
    free(g->cells_head); free(g->cells_tail); free(g->particles_next); free(g);
}

/* :111-112 — float subtract, float divide, truncate toward zero.  Clamp is ours. */
static inline void cell_of(const oracle_grid *g, float x, float y, int *row, int *col, int *clamped)
{
    int r = (int)((y - g->y_min) / g->cell_length);
    int c = (int)((x - g->x_min) / g->cell_length);
    int bad = 0;
    if (r < 0) { r = 0; bad = 1; } else if (r >= g->n_cells) { r = g->n_cells - 1; bad = 1; }
    if (c < 0) { c = 0; bad = 1; } else if (c >= g->m_cells) { c = g->m_cells - 1; bad = 1; }
    *row = r; *col = c;
    if (clamped) *clamped = bad;
}

/* :104-124 — serial tail-append, so each cell's list is in ascending particle index */
void oracle_grid_update(oracle_grid *g, const oracle_particle *particles)
{
    size_t nc = (size_t)g->n_cells * (size_t)g->m_cells;
    for (size_t c = 0; c < nc; c++) g->cells_head[c] = g->cells_tail[c] = ORACLE_NIL;
    g->n_clamped = 0;
    for (int i = 0; i < g->n_particles; i++) {
        int row, col, bad;
        cell_of(g, particles[i].x, particles[i].y, &row, &col, &bad);
        g->n_clamped += bad;
        size_t cell = (size_t)row * g->m_cells + col;       /* :113 */
        if (g->cells_head[cell] == ORACLE_NIL) {
            g->cells_head[cell] = g->cells_tail[cell] = (uint32_t)i;
        } else {
            g->particles_next[g->cells_tail[cell]] = (uint32_t)i;
            g->cells_tail[cell] = (uint32_t)i;
        }
        g->particles_next[i] = ORACLE_NIL;
    }
}

void oracle_cell_ids(const oracle_grid *g, const oracle_particle *p, int n, int *cell_out)
{
    for (int i = 0; i < n; i++) {
        int row, col;
        cell_of(g, p[i].x, p[i].y, &row, &col, NULL);
        cell_out[i] = row * g->m_cells + col;
    }
}

/* :126-153.  Rows outer, columns inner, list order inside a cell; accept iff
 * sqrtf(d2) < 2*H (float compare) and, for same-array queries, j != i (:130, :144;
 * the reference's flag is named backwards). */
int oracle_find_neighbors(const oracle_params *prm, int *j_out, int cap,
                          const oracle_particle *a, const oracle_particle *b, int same_array,
                          int i, const oracle_grid *gb, oracle_counters *ctr)
{
    const float support = 2 * prm->H;
    int count = 0, overflow = 0;
    int row_c = (int)((a[i].y - gb->y_min) / gb->cell_length);
    int col_c = (int)((a[i].x - gb->x_min) / gb->cell_length);
    for (int row = row_c - 1; row <= row_c + 1; row++) {
        for (int col = col_c - 1; col <= col_c + 1; col++) {
            if (row < 0 || row >= gb->n_cells || col < 0 || col >= gb->m_cells) continue;
            size_t cell = (size_t)row * gb->m_cells + col;
            for (uint32_t j = gb->cells_head[cell]; j != ORACLE_NIL; j = gb->particles_next[j]) {
                float d = oracle_euclid_dist(a[i].x, a[i].y, b[j].x, b[j].y);
                if (d < support && (!same_array || (uint32_t)i != j)) {
                    if (count < cap) j_out[count++] = (int)j;
                    else overflow = 1;
                }
            }
        }
    }
    if (ctr) {
        if (overflow) {
#pragma omp atomic
            ctr->neighbor_overflows++;
        }
        if (count > ctr->max_neighbors_seen) {
#pragma omp critical(oracle_maxnb)
            if (count > ctr->max_neighbors_seen) ctr->max_neighbors_seen = count;
        }
    }
    return count;
}

int oracle_neighbor_list(const oracle_params *prm, const oracle_particle *a,
                         const oracle_particle *b, int same_array, int i,
                         const oracle_grid *gb, int *j_out, int cap)
{
    return oracle_find_neighbors(prm, j_out, cap, a, b, same_array, i, gb, NULL);
}

/* ------------------------------------------------------------------ operators */

/* :242-261 — psi_i = rho_i / sum_{j != i} W_ij, sequential float sum from 0 */
void oracle_boundary_pseudomass(const oracle_params *prm, oracle_particle *boundary,
                                const oracle_grid *gb, oracle_counters *ctr)
{
    const int cap = prm->max_neighbors;
#pragma omp parallel
    {
        int *nb = (int *)malloc((size_t)cap * sizeof(int));
#pragma omp for schedule(static)
        for (int i = 0; i < gb->n_particles; i++) {
            int n = oracle_find_neighbors(prm, nb, cap, boundary, boundary, 1, i, gb, ctr);
            float recip_volume = 0;
            for (int k = 0; k < n; k++) {
                const oracle_particle *bj = &boundary[nb[k]];
                recip_volume += oracle_W(prm, boundary[i].x, boundary[i].y, bj->x, bj->y);
            }
            boundary[i].m = boundary[i].rho / recip_volume;
        }
        free(nb);
    }
}

/* :200-214 with MASS and quantity == 1: sum += (m_j * 1.0f) * W_ij, sequential */
static float sum_mass_W(const oracle_params *prm, const oracle_particle *pi,
                        const oracle_particle *set, const int *nb, int n)
{
    float s = 0;
    for (int k = 0; k < n; k++) {
        const oracle_particle *pj = &set[nb[k]];
        float w = oracle_W(prm, pi->x, pi->y, pj->x, pj->y);
        s += pj->m * 1.0f * w;
    }
    return s;
}

/* :263-289 — rho_i = (m_i*W(0) + ff) + fb */
void oracle_density(const oracle_params *prm, oracle_particle *fluid,
                    const oracle_particle *boundary, const oracle_grid *gf,
                    const oracle_grid *gb, oracle_counters *ctr)
{
    const int cap = prm->max_neighbors;
    const float W_ii = oracle_W(prm, 0, 0, 0, 0);     /* :274 */
#pragma omp parallel
    {
        int *nb = (int *)malloc((size_t)cap * sizeof(int));
#pragma omp for schedule(static)
        for (int i = 0; i < gf->n_particles; i++) {
            float self_density = fluid[i].m * W_ii;
            int n = oracle_find_neighbors(prm, nb, cap, fluid, fluid, 1, i, gf, ctr);
            float ff = sum_mass_W(prm, &fluid[i], fluid, nb, n);
            n = oracle_find_neighbors(prm, nb, cap, fluid, boundary, 0, i, gb, ctr);
            float fb = sum_mass_W(prm, &fluid[i], boundary, nb, n);
            fluid[i].rho = self_density + ff + fb;      /* :287 */
        }
        free(nb);
    }
}

/* :294-301 — Tait, clamped at zero */
void oracle_pressure(const oracle_params *prm, oracle_particle *particles, int n)
{
    const float B = prm->c0 * prm->c0 * prm->rho0 / 7;   /* :297 */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        float p = B * (pow7_(particles[i].rho / prm->rho0) - 1);
        particles[i].p = (p > 0) ? p : 0;
    }
}

/* :303-373 */
void oracle_accelerations(const oracle_params *prm, float *du_dt, float *dv_dt,
                          const oracle_particle *fluid, const oracle_particle *boundary,
                          const oracle_grid *gf, const oracle_grid *gb, float gx, float gy,
                          oracle_counters *ctr)
{
    const int cap = prm->max_neighbors;
    const float H = prm->H, C = prm->c0;
    const double Hd = (double)H;
    const float W_ref = oracle_W(prm, (float)(0.2 * Hd), 0, 0, 0);   /* :325 W(0.2*H,0,0,0) */
    const double eps_h2 = 0.01 * Hd * Hd;                             /* :332 0.01*H*H */
    const double visc_c = -0.01 * (double)C;                          /* :334 -0.01*C */
#pragma omp parallel
    {
        int *nb = (int *)malloc((size_t)cap * sizeof(int));
        float *temp = (float *)malloc((size_t)cap * sizeof(float));
#pragma omp for schedule(static)
        for (int i = 0; i < gf->n_particles; i++) {
            const oracle_particle fi = fluid[i];

            /* fluid neighbours, :314-340 */
            int n = oracle_find_neighbors(prm, nb, cap, fluid, fluid, 1, i, gf, ctr);
            for (int k = 0; k < n; k++) {
                const oracle_particle *fj = &fluid[nb[k]];
                float pressure_ij = fi.p / (fi.rho * fi.rho) + fj->p / (fj->rho * fj->rho);
                float W_ij = oracle_W(prm, fi.x, fi.y, fj->x, fj->y);
                float art_ij = (float)(0.1 * (double)pow4_(W_ij / W_ref));
                float u_ij = fi.u - fj->u, v_ij = fi.v - fj->v;
                float x_ij = fi.x - fj->x, y_ij = fi.y - fj->y;
                float xu = x_ij * u_ij + y_ij * v_ij;
                float xx = x_ij * x_ij + y_ij * y_ij;
                float mu_ij = (float)((double)(H * xu) / ((double)xx + eps_h2));
                float mean_rho = (fi.rho + fj->rho) / 2;
                float visc_ij = (xu < 0) ? (float)(visc_c * (double)mu_ij / (double)mean_rho) : 0.0f;
                temp[k] = pressure_ij + art_ij + visc_ij;
            }
            float sx = 0, sy = 0;       /* :216-231 */
            for (int k = 0; k < n; k++) {
                const oracle_particle *fj = &fluid[nb[k]];
                float gwx, gwy;
                oracle_grad_W(prm, fi.x, fi.y, fj->x, fj->y, &gwx, &gwy);
                sx += fj->m * temp[k] * gwx;
                sy += fj->m * temp[k] * gwy;
            }

            /* boundary neighbours, :343-368 */
            n = oracle_find_neighbors(prm, nb, cap, fluid, boundary, 0, i, gb, ctr);
            for (int k = 0; k < n; k++) {
                const oracle_particle *bj = &boundary[nb[k]];
                float pressure_ij = fi.p / (fi.rho * fi.rho);
                float W_ij = oracle_W(prm, fi.x, fi.y, bj->x, bj->y);
                float art_ij = (float)(0.1 * (double)pow4_(W_ij / W_ref));
                float u_ij = fi.u - bj->u, v_ij = fi.v - bj->v;
                float x_ij = fi.x - bj->x, y_ij = fi.y - bj->y;
                float xu = x_ij * u_ij + y_ij * v_ij;
                float xx = x_ij * x_ij + y_ij * y_ij;
                float mu_ij = (float)((double)(H * xu) / ((double)xx + eps_h2));
                float visc_ij = (xu < 0) ? (float)(visc_c * (double)mu_ij / (double)fi.rho) : 0.0f;
                temp[k] = pressure_ij + art_ij + visc_ij;
            }
            float bx = 0, by = 0;
            for (int k = 0; k < n; k++) {
                const oracle_particle *bj = &boundary[nb[k]];
                float gwx, gwy;
                oracle_grad_W(prm, fi.x, fi.y, bj->x, bj->y, &gwx, &gwy);
                bx += bj->m * temp[k] * gwx;
                by += bj->m * temp[k] * gwy;
            }

            du_dt[i] = gx - sx - bx;    /* :370 */
            dv_dt[i] = gy - sy - by;    /* :371 */
        }
        free(nb); free(temp);
    }
}

/* ------------------------------------------------------------------ leapfrog */

/* :615-618 / :637-640 — u += 0.5*DT*a with the product and the add in double */
void oracle_kick(const oracle_params *prm, oracle_particle *fluid, int n,
                 const float *du_dt, const float *dv_dt)
{
    const double half_dt = 0.5 * (double)prm->dt;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        fluid[i].u = (float)((double)fluid[i].u + half_dt * (double)du_dt[i]);
        fluid[i].v = (float)((double)fluid[i].v + half_dt * (double)dv_dt[i]);
    }
}

/* :621-624 — float multiply then float add */
void oracle_drift(const oracle_params *prm, oracle_particle *fluid, int n)
{
    const float dt = prm->dt;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        fluid[i].x += dt * fluid[i].u;
        fluid[i].y += dt * fluid[i].v;
    }
}

/* :604-607 */
void oracle_compute_accel(const oracle_params *prm, oracle_particle *fluid, int n_fluid,
                          const oracle_particle *boundary, oracle_grid *gf,
                          const oracle_grid *gb, float gx, float gy, float *du_dt,
                          float *dv_dt, oracle_counters *ctr)
{
    oracle_grid_update(gf, fluid);
    oracle_density(prm, fluid, boundary, gf, gb, ctr);
    oracle_pressure(prm, fluid, n_fluid);
    oracle_accelerations(prm, du_dt, dv_dt, fluid, boundary, gf, gb, gx, gy, ctr);
}

/* :612-641 */
void oracle_step(const oracle_params *prm, oracle_particle *fluid, int n_fluid,
                 const oracle_particle *boundary, oracle_grid *gf, const oracle_grid *gb,
                 float gx, float gy, const float *gxy_per_step, int nsteps, float *du_dt,
                 float *dv_dt, oracle_counters *ctr)
{
    for (int s = 0; s < nsteps; s++) {
        if (gxy_per_step) { gx = gxy_per_step[2 * s]; gy = gxy_per_step[2 * s + 1]; }
        oracle_kick(prm, fluid, n_fluid, du_dt, dv_dt);
        oracle_drift(prm, fluid, n_fluid);
        oracle_compute_accel(prm, fluid, n_fluid, boundary, gf, gb, gx, gy, du_dt, dv_dt, ctr);
        oracle_kick(prm, fluid, n_fluid, du_dt, dv_dt);
    }
}

/* ------------------------------------------------------------------ render */

/* :570-577 — (j+0.5)*WIDTH/128 and (64-(i+0.5))*HEIGHT/64 in double, then float */
void oracle_pixel_pseudoparticles(const oracle_params *prm, oracle_particle *pixels)
{
    memset(pixels, 0, 64 * 128 * sizeof *pixels);
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 128; j++) {
            pixels[i * 128 + j].x = (float)((j + 0.5) * (double)prm->width / 128);
            pixels[i * 128 + j].y = (float)((64 - (i + 0.5)) * (double)prm->height / 64);
        }
}

/* :380-411 */
void oracle_draw_metaballs(const oracle_params *prm, unsigned char *draw_buffer,
                           const oracle_particle *pixels, const oracle_particle *fluid,
                           const oracle_grid *gf, oracle_counters *ctr)
{
    const int cap = prm->max_neighbors;
    const float px_width = prm->width / 128;                       /* :399 */
    const float W_px = oracle_W(prm, px_width / 2, 0, 0, 0);       /* :401 */
    int *nb = (int *)malloc((size_t)cap * sizeof(int));
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 128; j++) {
            int ij = i * 128 + j;
            int n = oracle_find_neighbors(prm, nb, cap, pixels, fluid, 0, ij, gf, ctr);
            float cond = 0;
            for (int k = 0; k < n; k++) {
                const oracle_particle *fj = &fluid[nb[k]];
                float w = oracle_W(prm, pixels[ij].x, pixels[ij].y, fj->x, fj->y);
                cond += w / W_px;
                if (cond >= 1) break;
            }
            if (cond >= 1) draw_buffer[i / 8 * 128 + j] |= (unsigned char)(1 << (i % 8));
            else draw_buffer[i / 8 * 128 + j] &= (unsigned char)~(1 << (i % 8));
        }
    free(nb);
}

/* ------------------------------------------------------------------ scenes */

/* :238-240 — float distance compared against the double literal 0.70 */
static int in_drop(const oracle_params *prm, float x, float y)
{
    return (double)oracle_euclid_dist(x, y, prm->width / 2, prm->height / 2) < 0.70;
}

/* :485-488 */
int oracle_scene_count_drop(const oracle_params *prm)
{
    int n = 0;
    for (float x0 = 0; x0 < prm->width; x0 += prm->R)
        for (float y0 = 0; y0 < prm->height; y0 += prm->R)
            if (in_drop(prm, x0, y0)) n++;
    return n;
}

/* :496-506 */
void oracle_scene_fill_drop(const oracle_params *prm, oracle_particle *fluid)
{
    int k = 0;
    for (float x0 = 0; x0 < prm->width; x0 += prm->R)
        for (float y0 = 0; y0 < prm->height; y0 += prm->R)
            if (in_drop(prm, x0, y0)) {
                oracle_particle p = { x0, y0, 0, 0, prm->mass, prm->rho0, 0 };
                fluid[k++] = p;
            }
}

/* builder-defined rectangular block on the same accumulated lattice: keeps lattice
 * points with x0 <= x < x1 and y0 <= y < y1 */
int oracle_scene_count_block(const oracle_params *prm, float x0, float x1, float y0, float y1)
{
    int n = 0;
    for (float x = 0; x < prm->width; x += prm->R) {
        if (x < x0 || !(x < x1)) continue;
        for (float y = 0; y < prm->height; y += prm->R)
            if (y >= y0 && y < y1) n++;
    }
    return n;
}

void oracle_scene_fill_block(const oracle_params *prm, oracle_particle *fluid, float x0,
                             float x1, float y0, float y1)
{
    int k = 0;
    for (float x = 0; x < prm->width; x += prm->R) {
        if (x < x0 || !(x < x1)) continue;
        for (float y = 0; y < prm->height; y += prm->R)
            if (y >= y0 && y < y1) {
                oracle_particle p = { x, y, 0, 0, prm->mass, prm->rho0, 0 };
                fluid[k++] = p;
            }
    }
}

/* :514-516 */
int oracle_scene_count_boundary(const oracle_params *prm)
{
    int n = 0;
    for (float x0 = 0; x0 < prm->width; x0 += prm->R) n += 2;
    for (float y0 = 0; y0 < prm->height; y0 += prm->R) n += 2;
    return n;
}

/* :522-540 — bottom/top pairs along x, then left/right pairs along y */
void oracle_scene_fill_boundary(const oracle_params *prm, oracle_particle *boundary)
{
    int k = 0;
    for (float x0 = 0; x0 < prm->width; x0 += prm->R) {
        oracle_particle lo = { x0, 0, 0, 0, 0, prm->rho0, 0 };
        oracle_particle hi = { x0, prm->height, 0, 0, 0, prm->rho0, 0 };
        boundary[k] = lo; boundary[k + 1] = hi; k += 2;
    }
    for (float y0 = 0; y0 < prm->height; y0 += prm->R) {
        oracle_particle lf = { 0, y0, 0, 0, 0, prm->rho0, 0 };
        oracle_particle rt = { prm->width, y0, 0, 0, 0, prm->rho0, 0 };
        boundary[k] = lf; boundary[k + 1] = rt; k += 2;
    }
}

/* :439-440 */
void oracle_gravity_from_raw(const oracle_params *prm, int accel_x_raw, int accel_y_raw,
                             float *gx, float *gy)
{
    *gx = (float)accel_y_raw / (1 << 14) * prm->g;
    *gy = -(float)accel_x_raw / (1 << 14) * prm->g;
}

/* ------------------------------------------------------------------ misc */

uint64_t oracle_grid_fnv(const oracle_grid *g)
{
    uint64_t h = 0xcbf29ce484222325ULL;
    size_t nc = (size_t)g->n_cells * (size_t)g->m_cells;
    for (size_t c = 0; c < nc; c++)
        for (uint32_t j = g->cells_head[c]; j != ORACLE_NIL; j = g->particles_next[j]) {
            uint64_t v = (uint64_t)c * 65536u + j;
            for (int b = 0; b < 8; b++) {
                h ^= (v >> (8 * b)) & 0xff;
                h *= 0x100000001b3ULL;
            }
        }
    return h;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* scene.c — host-side scene builders (see include/sph_b200_scene.h). */
#include <math.h>
#include <stddef.h>

#include "../../include/sph_b200_scene.h"

/* pi_sph_fluid.c:238-240 — float distance to the tank centre against the double 0.70 */
static int inside_drop(const sphb_params *prm, float x, float y)
{
    const float cx = prm->width / 2, cy = prm->height / 2;
    const float dx = x - cx, dy = y - cy;
    const float dist = sqrtf(dx * dx + dy * dy);
    return (double)dist < 0.70;
}

static sphb_particle fluid_particle(const sphb_params *prm, float x, float y)
{
    /* :500-502 — at rest, m = RHO_0*V, rho = RHO_0 */
    sphb_particle p = { x, y, 0.0f, 0.0f, prm->rho0 * prm->vol, prm->rho0, 0.0f };
    return p;
}

int sphb_scene_count_drop(const sphb_params *prm)
{
    if (!prm || !(prm->R > 0)) return SPHB_E_ARG;
    int n = 0;
    for (float x = 0; x < prm->width; x += prm->R)
        for (float y = 0; y < prm->height; y += prm->R)
            n += inside_drop(prm, x, y);
    return n;
}

int sphb_scene_fill_drop(const sphb_params *prm, sphb_particle *out)
{
    if (!prm || !out || !(prm->R > 0)) return SPHB_E_ARG;
    int n = 0;
    for (float x = 0; x < prm->width; x += prm->R)
        for (float y = 0; y < prm->height; y += prm->R)
            if (inside_drop(prm, x, y)) out[n++] = fluid_particle(prm, x, y);
    return n;
}

int sphb_scene_count_block(const sphb_params *prm, float x0, float x1, float y0, float y1)
{
    if (!prm || !(prm->R > 0)) return SPHB_E_ARG;
    long long nx = 0, ny = 0;
    for (float x = 0; x < prm->width; x += prm->R) nx += (x >= x0 && x < x1);
    for (float y = 0; y < prm->height; y += prm->R) ny += (y >= y0 && y < y1);
    const long long n = nx * ny;
    return n > 2000000000LL ? SPHB_E_ARG : (int)n;
}

int sphb_scene_fill_block(const sphb_params *prm, float x0, float x1, float y0, float y1, sphb_particle *out)
{
    if (!prm || !out || !(prm->R > 0)) return SPHB_E_ARG;
    int n = 0;
    for (float x = 0; x < prm->width; x += prm->R) {
        if (!(x >= x0 && x < x1)) continue;
        for (float y = 0; y < prm->height; y += prm->R)
            if (y >= y0 && y < y1) out[n++] = fluid_particle(prm, x, y);
    }
    return n;
}

int sphb_scene_count_boundary(const sphb_params *prm)
{
    if (!prm || !(prm->R > 0)) return SPHB_E_ARG;
    int n = 0;
    for (float x = 0; x < prm->width; x += prm->R) n += 2;     /* :515 */
    for (float y = 0; y < prm->height; y += prm->R) n += 2;    /* :516 */
    return n;
}

int sphb_scene_fill_boundary(const sphb_params *prm, sphb_particle *out)
{
    if (!prm || !out || !(prm->R > 0)) return SPHB_E_ARG;
    int n = 0;
    /* :523-531 — floor and ceiling; .m is left 0 (computed by sphb_init_boundary) */
    for (float x = 0; x < prm->width; x += prm->R) {
        sphb_particle lo = { x, 0.0f, 0, 0, 0, prm->rho0, 0 };
        sphb_particle hi = { x, prm->height, 0, 0, 0, prm->rho0, 0 };
        out[n++] = lo;
        out[n++] = hi;
    }
    /* :532-540 — left and right walls (the (0,0) corner is emitted twice, as in the reference) */
    for (float y = 0; y < prm->height; y += prm->R) {
        sphb_particle lf = { 0.0f, y, 0, 0, 0, prm->rho0, 0 };
        sphb_particle rt = { prm->width, y, 0, 0, 0, prm->rho0, 0 };
        out[n++] = lf;
        out[n++] = rt;
    }
    return n;
}

int sphb_gravity_trace_tilt(const sphb_params *prm, float amplitude_deg, int period_steps, int hold_steps,
                            int nsteps, float *out)
{
    if (!prm || !out || period_steps <= 0 || hold_steps <= 0 || nsteps < 0) return SPHB_E_ARG;
    const double two_pi = 6.283185307179586476925286766559;
    float gx = 0, gy = 0;
    for (int s = 0; s < nsteps; s++) {
        if (s % hold_steps == 0) {
            const double theta = (double)amplitude_deg * two_pi / 360.0 * sin(two_pi * (double)s / period_steps);
            const int ax_raw = (int)lround(16384.0 * cos(theta));
            const int ay_raw = (int)lround(16384.0 * sin(theta));
            sphb_gravity_from_raw(prm, ax_raw, ay_raw, &gx, &gy);     /* :439-440 */
        }
        out[2 * s] = gx;
        out[2 * s + 1] = gy;
    }
    return SPHB_OK;
}

float sphb_spacing_for_count(double area, double n_target)
{
    return (float)sqrt(area / n_target);
}

/* ---- slab planning (multi-GPU), host side ---------------------------------------------------- */

int sphb_grid_columns(const sphb_params *prm, int *rows, int *cols)
{
    if (!prm || !(prm->cell_length > 0)) return SPHB_E_ARG;
    if (rows) *rows = (int)((prm->y_max - prm->y_min) / prm->cell_length) + 1;     /* :93 */
    if (cols) *cols = (int)((prm->x_max - prm->x_min) / prm->cell_length) + 1;     /* :94 */
    return SPHB_OK;
}

/* :112 — j_cell = (int)((x - x_min)/cell_length), clamped into the grid like the kernels do */
int sphb_column_of(const sphb_params *prm, float x)
{
    int cols = 0;
    if (sphb_grid_columns(prm, NULL, &cols)) return SPHB_E_ARG;
    const float d = x - prm->x_min;
    int c = (int)(d / prm->cell_length);
    if (c < 0) c = 0;
    if (c >= cols) c = cols - 1;
    return c;
}

int sphb_column_histogram(const sphb_params *prm, const sphb_particle *particles, int n, unsigned long long *hist)
{
    int cols = 0;
    if (!hist || n < 0 || (n > 0 && !particles) || sphb_grid_columns(prm, NULL, &cols)) return SPHB_E_ARG;
    for (int i = 0; i < n; i++) hist[sphb_column_of(prm, particles[i].x)]++;
    return SPHB_OK;
}

/* cuts[r] = first column of rank r: the columns are split at the particle-count quantiles, then
 * widened so that every slab has at least min_width columns */
int sphb_mg_plan_cuts_cost(const unsigned long long *hist, int cols, int world, int min_width, double column_cost, int *cuts)
{
    if (!hist || !cuts || world < 1 || cols < 1 || !(column_cost >= 0.0)) return SPHB_E_ARG;
    if (min_width < 1) min_width = 1;
    if ((long long)world * min_width > cols) return SPHB_E_ARG;
    /* cost of a column = its particles + what its (possibly empty) cells cost the scan, in particle units */
    long double total = 0;
    for (int c = 0; c < cols; c++) total += (long double)hist[c] + column_cost;
    cuts[0] = 0;
    cuts[world] = cols;
    long double acc = 0;
    int c = 0;
    for (int r = 1; r < world; r++) {
        /* smallest cut with at least r/world of the cost to its left */
        const long double want = total * r / world;
        while (c < cols && acc < want) { acc += (long double)hist[c] + column_cost; c++; }
        cuts[r] = c;
    }
    for (int r = 1; r < world; r++)          /* forward: minimum width of slab r-1 */
        if (cuts[r] < cuts[r - 1] + min_width) cuts[r] = cuts[r - 1] + min_width;
    for (int r = world - 1; r >= 1; r--)     /* backward: minimum width of slab r */
        if (cuts[r] > cuts[r + 1] - min_width) cuts[r] = cuts[r + 1] - min_width;
    return SPHB_OK;
}

int sphb_mg_plan_cuts(const unsigned long long *hist, int cols, int world, int min_width, int *cuts)
{
    return sphb_mg_plan_cuts_cost(hist, cols, world, min_width, 0.0, cuts);
}

/* block scene restricted to the lattice columns whose x falls into cell columns [col_lo, col_hi):
 * because the lattice is filled x-outer (like :496-506) that subset is one contiguous range of
 * the full scene's particle indices; *id_base is its first index. */
int sphb_scene_fill_block_slab(const sphb_params *prm, float x0, float x1, float y0, float y1, int col_lo,
                               int col_hi, sphb_particle *out, unsigned int *id_base)
{
    if (!prm || !(prm->R > 0)) return SPHB_E_ARG;
    long long ny = 0;
    for (float y = 0; y < prm->height; y += prm->R) ny += (y >= y0 && y < y1);
    long long before = 0, n = 0;
    for (float x = 0; x < prm->width; x += prm->R) {
        if (!(x >= x0 && x < x1)) continue;
        const int col = sphb_column_of(prm, x);
        if (col < col_lo) { before += ny; continue; }
        if (col >= col_hi) break;
        if (out) {
            for (float y = 0; y < prm->height; y += prm->R)
                if (y >= y0 && y < y1) out[n++] = fluid_particle(prm, x, y);
        } else {
            n += ny;
        }
    }
    if (id_base) *id_base = (unsigned int)before;
    return n > 2000000000LL ? SPHB_E_ARG : (int)n;
}

/* per-column particle counts of the block scene without building it */
int sphb_scene_block_column_hist(const sphb_params *prm, float x0, float x1, float y0, float y1,
                                 unsigned long long *hist)
{
    if (!prm || !hist || !(prm->R > 0)) return SPHB_E_ARG;
    unsigned long long ny = 0;
    for (float y = 0; y < prm->height; y += prm->R) ny += (y >= y0 && y < y1);
    for (float x = 0; x < prm->width; x += prm->R)
        if (x >= x0 && x < x1) hist[sphb_column_of(prm, x)] += ny;
    return SPHB_OK;
}

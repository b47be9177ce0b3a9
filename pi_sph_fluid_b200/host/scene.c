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

/* ref_driver.c — drives a build of the REAL reference translation unit
 * (oracle/_ref/libpisph_ref_*.so, compiled by oracle/Makefile straight from
 * /root/reference/pi_sph_fluid.c) through the same call sequence its main() uses.
 *
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline); never part of the product path.
 *
 * Why a driver: the reference's operators contain orphaned `#pragma omp for` loops
 * (pi_sph_fluid.c:246,272,295,311) that only run multi-threaded when called by every
 * thread of an enclosing parallel region, and its main() (:475-704) never returns.  This
 * file opens the reference .so, and re-creates the enclosing region and the step
 * sequence of :600-607 and :610-641.  The three kick/drift loops are inline in the
 * reference's main() (:615-624, :637-640), so they are restated here (same types:
 * product and add in double for the kick, float for the drift).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

typedef struct { float x, y, u, v, m, rho, p; } ref_particle;   /* pi_sph_fluid.c:26-31 */

typedef void *(*fn_alloc_ctx)(int, float, float, float, float, float);                 /* :82  */
typedef void (*fn_update_ctx)(void *, ref_particle *);                                 /* :104 */
typedef int (*fn_find_neighbors)(int *, ref_particle *, ref_particle *, int, void *);  /* :126 */
typedef void (*fn_pseudomass)(ref_particle *, void *);                                 /* :242 */
typedef void (*fn_density)(ref_particle *, ref_particle *, void *, void *);            /* :263 */
typedef void (*fn_pressure)(ref_particle *, int);                                      /* :294 */
typedef void (*fn_accel)(float *, float *, ref_particle *, ref_particle *, void *, void *,
                         float, float);                                                /* :303 */
typedef void (*fn_metaballs)(unsigned char *, ref_particle *, ref_particle *, void *); /* :380 */

typedef struct {
    void *dl;
    fn_alloc_ctx alloc_ctx;
    fn_update_ctx update_ctx;
    fn_find_neighbors find_neighbors;
    fn_pseudomass pseudomass;
    fn_density density;
    fn_pressure pressure;
    fn_accel accel;
    fn_metaballs metaballs;
} refdrv;

static void *need(void *dl, const char *name)
{
    void *s = dlsym(dl, name);
    if (!s) { fprintf(stderr, "ref_driver: missing symbol %s\n", name); abort(); }
    return s;
}

refdrv *refdrv_open(const char *so_path)
{
    void *dl = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!dl) { fprintf(stderr, "ref_driver: %s\n", dlerror()); return NULL; }
    refdrv *r = (refdrv *)calloc(1, sizeof *r);
    r->dl = dl;
    r->alloc_ctx = (fn_alloc_ctx)need(dl, "alloc_neighbors_context");
    r->update_ctx = (fn_update_ctx)need(dl, "update_neighbors_context");
    r->find_neighbors = (fn_find_neighbors)need(dl, "find_neighbors");
    r->pseudomass = (fn_pseudomass)need(dl, "calculate_boundary_pseudomass");
    r->density = (fn_density)need(dl, "calculate_density");
    r->pressure = (fn_pressure)need(dl, "calculate_particle_pressure");
    r->accel = (fn_accel)need(dl, "calculate_accelerations");
    r->metaballs = (fn_metaballs)need(dl, "draw_metaballs");
    return r;
}

void refdrv_close(refdrv *r)
{
    if (!r) return;
    dlclose(r->dl);
    free(r);
}

void *refdrv_alloc_ctx(refdrv *r, int n, float x_min, float x_max, float y_min, float y_max,
                       float cell_length)
{
    return r->alloc_ctx(n, x_min, x_max, y_min, y_max, cell_length);
}

void refdrv_update_ctx(refdrv *r, void *ctx, ref_particle *particles) { r->update_ctx(ctx, particles); }

int refdrv_find_neighbors(refdrv *r, int *j_out, ref_particle *a, ref_particle *b, int i, void *ctx_b)
{
    return r->find_neighbors(j_out, a, b, i, ctx_b);
}

/* :600-601 (serial, outside any parallel region, as in the reference) */
void refdrv_init_boundary(refdrv *r, ref_particle *boundary, void *ctx_b)
{
    r->update_ctx(ctx_b, boundary);
    r->pseudomass(boundary, ctx_b);
}

/* :604-607 */
void refdrv_compute_accel(refdrv *r, ref_particle *fluid, int n_fluid, ref_particle *boundary,
                          void *ctx_f, void *ctx_b, float gx, float gy, float *du_dt, float *dv_dt)
{
    r->update_ctx(ctx_f, fluid);
    r->density(fluid, boundary, ctx_f, ctx_b);
    r->pressure(fluid, n_fluid);
    r->accel(du_dt, dv_dt, fluid, boundary, ctx_f, ctx_b, gx, gy);
}

void refdrv_draw_metaballs(refdrv *r, unsigned char *draw_buffer, ref_particle *pixels,
                           ref_particle *fluid, void *ctx_f)
{
    r->metaballs(draw_buffer, pixels, fluid, ctx_f);
}

static double now_s(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* :610-641 — nsteps iterations of the reference's while(1) body (draw/stats/realtime
 * governor excluded, as the benchmark protocol in BASELINE.md §3 says).  gxy_per_step,
 * when non-NULL, supplies the (gx,gy) each step reads at :632.  Returns wall seconds. */
double refdrv_step(refdrv *r, ref_particle *fluid, int n_fluid, ref_particle *boundary,
                   void *ctx_f, void *ctx_b, float dt, float gx, float gy,
                   const float *gxy_per_step, int nsteps, int nthreads, float *du_dt, float *dv_dt)
{
    const double half_dt = 0.5 * (double)dt;
    if (nthreads <= 0) nthreads = 4;   /* :610 num_threads(4) */
    double t0 = now_s();
#pragma omp parallel num_threads(nthreads)
    for (int s = 0; s < nsteps; s++) {
        float sgx = gxy_per_step ? gxy_per_step[2 * s] : gx;
        float sgy = gxy_per_step ? gxy_per_step[2 * s + 1] : gy;
#pragma omp single
        {
            for (int i = 0; i < n_fluid; i++) {                                  /* :615-618 */
                fluid[i].u = (float)((double)fluid[i].u + half_dt * (double)du_dt[i]);
                fluid[i].v = (float)((double)fluid[i].v + half_dt * (double)dv_dt[i]);
            }
            for (int i = 0; i < n_fluid; i++) {                                  /* :621-624 */
                fluid[i].x += dt * fluid[i].u;
                fluid[i].y += dt * fluid[i].v;
            }
            r->update_ctx(ctx_f, fluid);                                         /* :626 */
        }
        r->density(fluid, boundary, ctx_f, ctx_b);                               /* :630 */
        r->pressure(fluid, n_fluid);                                             /* :631 */
        r->accel(du_dt, dv_dt, fluid, boundary, ctx_f, ctx_b, sgx, sgy);         /* :632 */
#pragma omp single
        {
            for (int i = 0; i < n_fluid; i++) {                                  /* :637-640 */
                fluid[i].u = (float)((double)fluid[i].u + half_dt * (double)du_dt[i]);
                fluid[i].v = (float)((double)fluid[i].v + half_dt * (double)dv_dt[i]);
            }
        }
    }
    return now_s() - t0;
}

int refdrv_max_threads(void) { return omp_get_max_threads(); }

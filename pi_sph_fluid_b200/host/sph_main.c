/* sph_main.c — plain-C host driver over the C ABI of libsphb200.so.
 *
 * Mirrors the structure of the reference's main() (pi_sph_fluid.c:475-704): build the scene
 * (:484-540), print dt / particle counts (:543-545), initialise the boundary pseudo-mass and
 * the zero-th accelerations (:600-607), then loop { step; draw at 60 Hz; statistics every
 * 0.1 s of simulated time } (:610-691).  What differs: the state lives in HBM behind
 * sphb_ctx, the OLED is replaced by an optional ASCII dump of the same 1 KiB SSD1306 frame,
 * the MPU6050 by an optional synthetic tilt trace, and the run is bounded (--steps).
 *
 *   sph_b200_main [--scene drop|dam|tank] [--R 0.075] [--steps 4000] [--chunk 50]
 *                 [--tilt DEG] [--render] [--nondeterministic] [--device 0]
 *                 [--slabs N [--devices 0,1,...]]   multi-GPU: N x-slabs driven by this one host thread
 *                                                   (in-process transport; slab r runs on devices[r % count])
 *                 [--save FILE] [--load FILE]       state files (single GPU)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>

#include "sph_b200.h"
#include "sph_b200_scene.h"

static double now_s(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc__ = (call);                                                           \
        if (rc__ < 0) {                                                              \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, sphb_last_error()); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

/* the SSD1306 page layout of :407-408: byte (i/8)*128+j, bit i%8 */
static void print_frame(const unsigned char *buf)
{
    for (int i = 0; i < 64; i += 2) {
        char line[129];
        for (int j = 0; j < 128; j++) {
            int top = (buf[i / 8 * 128 + j] >> (i % 8)) & 1;
            int bot = (buf[(i + 1) / 8 * 128 + j] >> ((i + 1) % 8)) & 1;
            line[j] = top && bot ? '#' : (top ? '"' : (bot ? '_' : ' '));
        }
        line[128] = 0;
        puts(line);
    }
}

#define MAX_SLABS 64

/* Multi-GPU run: the same loop as below over sphb_mg_group_* (one host thread, N slab contexts).
 * The scene is split by the global cell column of each particle (sphb_column_of == :112); the cuts
 * sit at the particle-count quantiles.  Returns 0 on success. */
static int run_slabs(const sphb_params *prm0, int n_slabs, const int *devices, int n_devices,
                     const sphb_particle *fluid, int n_fluid, const sphb_particle *boundary, int n_boundary,
                     long steps, int chunk, float gx, float gy, int render)
{
    int cols = 0, cuts[MAX_SLABS + 1];
    sphb_grid_columns(prm0, NULL, &cols);
    unsigned long long *hist = (unsigned long long *)calloc((size_t)cols, sizeof *hist);
    CHECK(sphb_column_histogram(prm0, fluid, n_fluid, hist));
    CHECK(sphb_mg_plan_cuts(hist, cols, n_slabs, 4, cuts));
    free(hist);

    sphb_ctx *ctx[MAX_SLABS];
    sphb_particle *part = (sphb_particle *)malloc(sizeof *part * (size_t)(n_fluid > 0 ? n_fluid : 1));
    uint32_t *ids = (uint32_t *)malloc(sizeof *ids * (size_t)(n_fluid > 0 ? n_fluid : 1));
    int halo = 0;
    for (int i = 0; i < n_fluid; i++) {                       /* generous: 4 columns' worth around the busiest cut */
        const int c = sphb_column_of(prm0, fluid[i].x);
        for (int r = 1; r < n_slabs; r++) halo += (c >= cuts[r] - 2 && c < cuts[r] + 2);
    }
    halo = 2 * halo / (n_slabs > 1 ? n_slabs - 1 : 1) + 4096;
    for (int r = 0; r < n_slabs; r++) {
        sphb_params prm = *prm0;
        prm.device = devices[r % n_devices];
        CHECK(sphb_create(&prm, &ctx[r]));
        CHECK(sphb_mg_configure(ctx[r], r, n_slabs, cuts[r], cuts[r + 1], 0, halo));
    }
    CHECK(sphb_mg_connect_local(ctx, n_slabs));
    for (int r = 0; r < n_slabs; r++) {
        int n = 0;
        for (int i = 0; i < n_fluid; i++) {
            const int c = sphb_column_of(prm0, fluid[i].x);
            if (c >= cuts[r] && c < cuts[r + 1]) { part[n] = fluid[i]; ids[n++] = (uint32_t)i; }
        }
        printf("slab %d: device %d, columns [%d, %d), %d particles\n", r, devices[r % n_devices], cuts[r], cuts[r + 1], n);
        CHECK(sphb_mg_upload(ctx[r], part, ids, 0, n, boundary, n_boundary));
        CHECK(sphb_init_boundary(ctx[r]));                                         /* :600-601 */
    }
    CHECK(sphb_mg_group_compute_accel(ctx, n_slabs, gx, gy));                       /* :604-607 */
    const double t0 = now_s();
    unsigned char frame[1024], part_frame[1024];
    for (long done = 0; done < steps;) {
        const int n = (int)((steps - done) < chunk ? (steps - done) : chunk);
        CHECK(sphb_mg_group_step(ctx, n_slabs, gx, gy, NULL, n));                   /* :612-641 */
        done += n;
        if (done % (20L * chunk) == 0 || done == steps) {
            sphb_stats per[MAX_SLABS], all;
            for (int r = 0; r < n_slabs; r++) CHECK(sphb_get_stats(ctx[r], &per[r]));
            CHECK(sphb_mg_merge_stats(per, n_slabs, &all));
            printf("step %ld: %u particles, max rho error %.3f%%, max speed %.2f m/s, lost %u, overflow %u\n", done,
                   all.n_fluid, all.max_rho_err / prm0->rho0 * 100, all.max_speed, all.n_lost, all.n_overflow);
        }
    }
    CHECK(sphb_mg_group_synchronize(ctx, n_slabs));
    const double wall = now_s() - t0;
    printf("%ld steps of %d particles on %d slabs in %.3f s: %.3e particle-updates/s\n", steps, n_fluid, n_slabs, wall,
           (double)steps * n_fluid / wall);
    if (render) {
        /* every byte of the SSD1306 frame covers one pixel column, i.e. one x: OR-ing the slabs' frames
         * assembles the picture (a slab sees all neighbours of the pixels in its owned columns) */
        memset(frame, 0, sizeof frame);
        for (int r = 0; r < n_slabs; r++) {
            CHECK(sphb_render(ctx[r], part_frame));
            for (int i = 0; i < 1024; i++) frame[i] |= part_frame[i];
        }
        print_frame(frame);
    }
    for (int r = 0; r < n_slabs; r++) sphb_destroy(ctx[r]);
    free(part); free(ids);
    return 0;
}

int main(int argc, char **argv)
{
    const char *scene = "drop";
    float R = 0.0750f;                /* :11 */
    const float WIDTH = 4.0f, HEIGHT = 2.0f;   /* :13-14 */
    long steps = 4000;
    int chunk = 50, render = 0, deterministic = 1, device = 0;
    float tilt_deg = 0.0f;
    int n_slabs = 0, devices[MAX_SLABS], n_devices = 0;
    const char *save_path = NULL, *load_path = NULL;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--scene") && i + 1 < argc) scene = argv[++i];
        else if (!strcmp(argv[i], "--R") && i + 1 < argc) R = (float)atof(argv[++i]);
        else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atol(argv[++i]);
        else if (!strcmp(argv[i], "--chunk") && i + 1 < argc) chunk = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--tilt") && i + 1 < argc) tilt_deg = (float)atof(argv[++i]);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--slabs") && i + 1 < argc) n_slabs = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--devices") && i + 1 < argc) {
            char *list = argv[++i];
            for (char *tok = strtok(list, ","); tok && n_devices < MAX_SLABS; tok = strtok(NULL, ",")) devices[n_devices++] = atoi(tok);
        }
        else if (!strcmp(argv[i], "--save") && i + 1 < argc) save_path = argv[++i];
        else if (!strcmp(argv[i], "--load") && i + 1 < argc) load_path = argv[++i];
        else if (!strcmp(argv[i], "--render")) render = 1;
        else if (!strcmp(argv[i], "--nondeterministic")) deterministic = 0;
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    if (chunk < 1) chunk = 1;

    sphb_params prm;
    CHECK(sphb_default_params(&prm, R, WIDTH, HEIGHT));
    prm.deterministic = deterministic;
    prm.device = device;

    /* scene, :484-540 */
    int n_fluid, n_boundary;
    sphb_particle *fluid, *boundary;
    if (!strcmp(scene, "drop")) {
        n_fluid = sphb_scene_count_drop(&prm);
        fluid = (sphb_particle *)malloc(sizeof *fluid * (size_t)(n_fluid > 0 ? n_fluid : 1));
        sphb_scene_fill_drop(&prm, fluid);
    } else {
        /* dam: column x in [2R,2), y in [2R,1);  tank: filled to y = 1 across the width.  The
         * blocks start 2R off the walls: one lattice step away the Akinci single-layer wall
         * (psi ~ 4.2 m) would put rho at 1700 and p at 9e8 Pa at t = 0. */
        const float x1 = !strcmp(scene, "dam") ? 2.0f : WIDTH - 1.5f * R;
        n_fluid = sphb_scene_count_block(&prm, 2 * R, x1, 2 * R, 1.0f);
        fluid = (sphb_particle *)malloc(sizeof *fluid * (size_t)(n_fluid > 0 ? n_fluid : 1));
        sphb_scene_fill_block(&prm, 2 * R, x1, 2 * R, 1.0f, fluid);
    }
    n_boundary = sphb_scene_count_boundary(&prm);
    boundary = (sphb_particle *)malloc(sizeof *boundary * (size_t)n_boundary);
    sphb_scene_fill_boundary(&prm, boundary);
    float *du_dt = (float *)malloc(sizeof(float) * (size_t)n_fluid);
    float *dv_dt = (float *)malloc(sizeof(float) * (size_t)n_fluid);

    printf("dt = %f    (expected ticks/s) %d\n", prm.dt, (int)(1 / prm.dt));     /* :543 */
    printf("n_fluid = %d\n", n_fluid);                                            /* :544 */
    printf("n_boundary = %d\n", n_boundary);                                      /* :545 */

    float gx = 0.0f, gy = -prm.g;                                                 /* :442-443 */
    if (n_slabs > MAX_SLABS) { fprintf(stderr, "at most %d slabs\n", MAX_SLABS); return 2; }
    if (n_slabs > 0) {
        if (n_devices == 0) devices[n_devices++] = device;
        return run_slabs(&prm, n_slabs, devices, n_devices, fluid, n_fluid, boundary, n_boundary, steps, chunk, gx, gy, render);
    }

    sphb_ctx *ctx = NULL;
    if (load_path) {
        CHECK(sphb_load_state(load_path, device, &ctx));                          /* continues a saved run */
        sphb_stats st0;
        CHECK(sphb_get_stats(ctx, &st0));
        n_fluid = (int)st0.n_fluid;
        printf("loaded %s: %d particles at step %llu\n", load_path, n_fluid, st0.steps);
        free(fluid); free(du_dt); free(dv_dt);
        fluid = (sphb_particle *)malloc(sizeof *fluid * (size_t)(n_fluid > 0 ? n_fluid : 1));
        du_dt = (float *)malloc(sizeof(float) * (size_t)(n_fluid > 0 ? n_fluid : 1));
        dv_dt = (float *)malloc(sizeof(float) * (size_t)(n_fluid > 0 ? n_fluid : 1));
    } else {
        CHECK(sphb_create(&prm, &ctx));
        CHECK(sphb_upload(ctx, fluid, n_fluid, boundary, n_boundary));
        CHECK(sphb_init_boundary(ctx));                                           /* :600-601 */
        CHECK(sphb_compute_accel(ctx, gx, gy));                                   /* :604-607 */
    }

    float *trace = NULL;
    if (tilt_deg != 0.0f) {
        trace = (float *)malloc(sizeof(float) * 2 * (size_t)chunk);
    }
    unsigned char frame[1024];
    memset(frame, 0, sizeof frame);                                               /* :563 */
    float worst_max_rho_error_pct = 0, max_max_speed = 0;                         /* :583 */
    double t = 0, last_t = 0;
    double last_reported = now_s(), last_drew = last_reported, t_begin = last_reported;

    for (long done = 0; done < steps;) {
        const int n = (int)((steps - done) < chunk ? (steps - done) : chunk);
        if (trace) {
            /* one sample per 410 steps ~ the reference's 10 Hz poll at dt = 2.44e-4 (:454-463) */
            float *full = (float *)malloc(sizeof(float) * 2 * (size_t)(done + n));
            sphb_gravity_trace_tilt(&prm, tilt_deg, 4 * 4102, 410, (int)(done + n), full);
            memcpy(trace, full + 2 * done, sizeof(float) * 2 * (size_t)n);
            free(full);
            CHECK(sphb_step_trace(ctx, trace, n));
        } else {
            CHECK(sphb_step(ctx, gx, gy, n));                                     /* :612-641 */
        }
        done += n;
        t += (double)n * prm.dt;                                                  /* :678 */

        const double now = now_s();
        if (render && now - last_drew > 1.0 / 60) {                               /* :648 */
            CHECK(sphb_render(ctx, frame));                                       /* :649 */
            print_frame(frame);
            last_drew = now;
        }
        if (t - last_t > 0.1 || done == steps) {                                  /* :679 */
            sphb_stats st;
            CHECK(sphb_get_stats(ctx, &st));
            const double wall = now_s() - last_reported;
            const float max_rho_error_pct = st.max_rho_err / prm.rho0 * 100;      /* :660, bug fixed */
            if (max_rho_error_pct > worst_max_rho_error_pct) worst_max_rho_error_pct = max_rho_error_pct;
            if (st.max_speed > max_max_speed) max_max_speed = st.max_speed;
            printf("sim time: %.2f, ticks/s: %d, max rho error: %.3f%% (worst) %.3f%%, "
                   "max speed: %.1f m/s (worst) %.1f m/s, escaped: %u\n",
                   t, (int)(((t - last_t) / prm.dt) / wall), max_rho_error_pct, worst_max_rho_error_pct,
                   st.max_speed, max_max_speed, st.n_escaped);                    /* :683-687 */
            last_t = t;
            last_reported = now_s();
        }
    }
    CHECK(sphb_download(ctx, fluid, du_dt, dv_dt));
    if (save_path) CHECK(sphb_save_state(ctx, save_path));
    const double wall = now_s() - t_begin;
    printf("%ld steps of %d particles in %.3f s: %.3e particle-updates/s\n", steps, n_fluid, wall,
           (double)steps * n_fluid / wall);
    if (render) { CHECK(sphb_render(ctx, frame)); print_frame(frame); }
    sphb_destroy(ctx);
    free(fluid); free(boundary); free(du_dt); free(dv_dt); free(trace);
    return 0;
}

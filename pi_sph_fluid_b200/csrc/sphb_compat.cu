// sphb_compat.cu — compat tier: the reference's seven operator entry points with the
// reference's exact signatures (pi_sph_fluid.c:82, :104, :242, :263, :294, :303, :380), each
// implemented as  host AoS -> HBM -> sm_100a kernels -> host AoS.  With these the reference's
// main() (display/gravity threads aside) links against libsphb200.so unchanged; INTEGRATION.md
// shows the link line.  They are for drop-in and parity use: the resident tier (sphb_step)
// is what avoids the per-call copies.
//
// Semantics kept from the reference:
//   * a context indexes the array handed to update_neighbors_context; operators read the
//     *current* field values of the arrays they are given (positions included) but the cell
//     lists of the last update, exactly like the linked lists at :142;
//   * only the fields the reference operator writes are written back (m at :259, rho at :287,
//     p at :299, du/dv at :370-371, the frame bits at :407-408);
//   * all functions return void; failures (no GPU, out of memory) abort with a message —
//     there is no CPU fallback to fall through to.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "sphb_internal.cuh"
#include "sph_consts.h"

using namespace sphb;

struct neighbors_context {
    sphb_ctx *core;        // core->fluid is "the particle set this context indexes"
    int n_particles;
    bool built;
    sphb_particle *h_tmp;  // host bounce buffer for partial write-back
    float2 *h_pixels;      // draw_metaballs: the pixel centres uploaded last (the reference builds them once, :570-577)
};

namespace {

sphb_params g_compat_prm;
bool g_compat_prm_set = false;

[[noreturn]] void die(const char *where)
{
    fprintf(stderr, "libsphb200 (%s): %s\n", where, sphb_last_error());
    abort();
}
#define COMPAT_CUDA(call, where)                                                      \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) { cuda_fail(e__, #call, __FILE__, __LINE__); die(where); } \
    } while (0)

const sphb_params &compat_params()
{
    if (!g_compat_prm_set) {
        sphb_default_params(&g_compat_prm, 0.075f, 4.0f, 2.0f);     // :11-14
        g_compat_prm_set = true;
    }
    return g_compat_prm;
}

// copy `aos` into the context's set keeping its permutation; returns #positions that moved
unsigned int refresh(neighbors_context *ctx, const sphb_particle *aos, const char *where)
{
    sphb_ctx *c = ctx->core;
    COMPAT_CUDA(cudaSetDevice(c->device), where);
    const size_t bytes = (size_t)ctx->n_particles * sizeof(sphb_particle);
    if (ensure_stage(c, bytes + 64)) die(where);
    unsigned int *d_moved = &c->d_counters->tiles_unstaged;      // scratch counter
    COMPAT_CUDA(cudaMemsetAsync(d_moved, 0, sizeof(unsigned int), c->stream), where);
    COMPAT_CUDA(cudaMemcpyAsync(c->d_stage, aos, bytes, cudaMemcpyHostToDevice, c->stream), where);
    c->launches += launch_refresh(c->stream, static_cast<const sphb_particle *>(c->d_stage), c->fluid, d_moved);
    unsigned int moved = 0;
    COMPAT_CUDA(cudaMemcpyAsync(&moved, d_moved, sizeof moved, cudaMemcpyDeviceToHost, c->stream), where);
    COMPAT_CUDA(cudaStreamSynchronize(c->stream), where);
    return moved;
}

struct PressureScratch {
    std::mutex mu;
    sphb_ctx *ctx = nullptr;
    sphb_particle *h_tmp = nullptr;
    int h_cap = 0;
    bool at_exit = false;
} g_pressure;

void pressure_release()
{
    std::lock_guard<std::mutex> lock(g_pressure.mu);
    if (g_pressure.ctx) sphb_destroy(g_pressure.ctx);
    free(g_pressure.h_tmp);
    g_pressure.ctx = nullptr;
    g_pressure.h_tmp = nullptr;
    g_pressure.h_cap = 0;
}

void require_built(neighbors_context *ctx, const char *where)
{
    if (!ctx || !ctx->core || !ctx->built) {
        set_error("neighbors_context used before update_neighbors_context");
        die(where);
    }
}

}  // namespace

extern "C" {

void sphb_compat_shutdown(void) { pressure_release(); }

int sphb_compat_set_params(const sphb_params *prm)
{
    if (!prm) return SPHB_E_ARG;
    g_compat_prm = *prm;
    g_compat_prm_set = true;
    return SPHB_OK;
}

// :82-102
struct neighbors_context *alloc_neighbors_context(int n_particles, float x_min, float x_max, float y_min,
                                                  float y_max, float cell_length)
{
    sphb_params prm = compat_params();
    prm.x_min = x_min; prm.x_max = x_max; prm.y_min = y_min; prm.y_max = y_max;
    prm.cell_length = cell_length;
    // the reference passes cell_length = 2*H (:596); if the caller's cell differs from the
    // configured one, re-derive the kernel scale from it
    if (cell_length != 2 * prm.H) {
        prm.H = cell_length / 2;
        prm.dt = 1.0f * prm.H / prm.c0;
        prm.vol = 0.57f * prm.H * prm.H;
    }
    neighbors_context *ctx = static_cast<neighbors_context *>(calloc(1, sizeof *ctx));
    if (!ctx) { set_error("out of memory"); die("alloc_neighbors_context"); }
    if (sphb_create(&prm, &ctx->core)) die("alloc_neighbors_context");
    ctx->n_particles = n_particles;
    ctx->h_tmp = static_cast<sphb_particle *>(malloc(sizeof(sphb_particle) * (size_t)(n_particles > 0 ? n_particles : 1)));
    return ctx;
}

void sphb_compat_free_context(struct neighbors_context *ctx)
{
    if (!ctx) return;
    sphb_destroy(ctx->core);
    free(ctx->h_tmp);
    free(ctx->h_pixels);
    free(ctx);
}

// :104-124 — counting-sort cell build of `particles`
void update_neighbors_context(struct neighbors_context *ctx, struct particle *particles)
{
    const char *where = "update_neighbors_context";
    if (!ctx || !ctx->core) { set_error("null context"); die(where); }
    sphb_ctx *c = ctx->core;
    const sphb_particle *p = reinterpret_cast<const sphb_particle *>(particles);
    COMPAT_CUDA(cudaSetDevice(c->device), where);
    // every context carries per-particle m and the rho field: it may index fluid or boundary
    if (alloc_set(c->fluid, ctx->n_particles, c->k.ncells, false, true)) die(where);
    if (!c->fluid.aux[0]) {
        for (int i = 0; i < 2; i++)
            COMPAT_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->fluid.aux[i]), sizeof(float) * ((size_t)ctx->n_particles + 4)), where);
    }
    c->fluid.uniform_mass = false;
    const size_t bytes = (size_t)ctx->n_particles * sizeof(sphb_particle);
    if (ensure_stage(c, bytes + 64)) die(where);
    COMPAT_CUDA(cudaMemcpyAsync(c->d_stage, p, bytes, cudaMemcpyHostToDevice, c->stream), where);
    c->launches += launch_aos_to_soa(c->stream, static_cast<const sphb_particle *>(c->d_stage), c->fluid, false);
    // aos_to_soa(fluid flavour) fills rho_prr/p/acc; also keep rho in aux for boundary use
    COMPAT_CUDA(cudaStreamSynchronize(c->stream), where);
    build_grid(c, c->fluid, false);
    COMPAT_CUDA(cudaStreamSynchronize(c->stream), where);
    COMPAT_CUDA(cudaGetLastError(), where);
    ctx->built = true;
}

// :242-261 — writes boundary[i].m only
void calculate_boundary_pseudomass(struct particle *boundary, struct neighbors_context *ctx_boundary)
{
    const char *where = "calculate_boundary_pseudomass";
    require_built(ctx_boundary, where);
    sphb_ctx *c = ctx_boundary->core;
    sphb_particle *b = reinterpret_cast<sphb_particle *>(boundary);
    refresh(ctx_boundary, b, where);
    c->launches += launch_pseudomass(c->stream, c->k, c->fluid);
    c->launches += launch_soa_to_aos(c->stream, c->fluid, static_cast<sphb_particle *>(c->d_stage), nullptr, nullptr, true);
    COMPAT_CUDA(cudaMemcpyAsync(ctx_boundary->h_tmp, c->d_stage, (size_t)ctx_boundary->n_particles * sizeof(sphb_particle),
                                cudaMemcpyDeviceToHost, c->stream), where);
    COMPAT_CUDA(cudaStreamSynchronize(c->stream), where);
    COMPAT_CUDA(cudaGetLastError(), where);
    for (int i = 0; i < ctx_boundary->n_particles; i++) b[i].m = ctx_boundary->h_tmp[i].m;     // :259
}

// :263-289 — writes fluid[i].rho only
void calculate_density(struct particle *fluid, struct particle *boundary, struct neighbors_context *ctx_fluid,
                       struct neighbors_context *ctx_boundary)
{
    const char *where = "calculate_density";
    require_built(ctx_fluid, where);
    require_built(ctx_boundary, where);
    sphb_ctx *c = ctx_fluid->core;
    sphb_particle *f = reinterpret_cast<sphb_particle *>(fluid);
    const unsigned int moved = refresh(ctx_fluid, f, where);
    refresh(ctx_boundary, reinterpret_cast<sphb_particle *>(boundary), where);
    c->launches += launch_density(c->stream, c->k, c->fluid, ctx_boundary->core->fluid, c->d_counters, false, moved == 0);
    c->launches += launch_soa_to_aos(c->stream, c->fluid, static_cast<sphb_particle *>(c->d_stage), nullptr, nullptr, false);
    COMPAT_CUDA(cudaMemcpyAsync(ctx_fluid->h_tmp, c->d_stage, (size_t)ctx_fluid->n_particles * sizeof(sphb_particle),
                                cudaMemcpyDeviceToHost, c->stream), where);
    COMPAT_CUDA(cudaStreamSynchronize(c->stream), where);
    COMPAT_CUDA(cudaGetLastError(), where);
    for (int i = 0; i < ctx_fluid->n_particles; i++) f[i].rho = ctx_fluid->h_tmp[i].rho;       // :287
}

// :294-301 — writes particles[i].p only.  No context in the signature: uses a private one.
void calculate_particle_pressure(struct particle *particles, int n_particles)
{
    const char *where = "calculate_particle_pressure";
    if (n_particles <= 0) return;
    // the signature carries no context: one private context for the process, serialised (the reference calls
    // this from every thread of its OpenMP team, :631) and released by sphb_compat_shutdown / at exit
    std::lock_guard<std::mutex> lock(g_pressure.mu);
    sphb_ctx *&scratch = g_pressure.ctx;
    sphb_particle *&h_tmp = g_pressure.h_tmp;
    int &h_cap = g_pressure.h_cap;
    if (!scratch) {
        if (sphb_create(&compat_params(), &scratch)) die(where);
        if (!g_pressure.at_exit) { atexit(pressure_release); g_pressure.at_exit = true; }
    }
    sphb_ctx *c = scratch;
    sphb_particle *p = reinterpret_cast<sphb_particle *>(particles);
    COMPAT_CUDA(cudaSetDevice(c->device), where);
    const size_t bytes = (size_t)n_particles * sizeof(sphb_particle);
    if (ensure_stage(c, bytes + 64)) die(where);
    if (h_cap < n_particles) {
        free(h_tmp);
        h_tmp = static_cast<sphb_particle *>(malloc(bytes));
        h_cap = n_particles;
    }
    Consts k = make_consts(compat_params(), 0.0f);
    COMPAT_CUDA(cudaMemcpyAsync(c->d_stage, p, bytes, cudaMemcpyHostToDevice, c->stream), where);
    c->launches += launch_tait_aos(c->stream, k, n_particles, static_cast<sphb_particle *>(c->d_stage));
    COMPAT_CUDA(cudaMemcpyAsync(h_tmp, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream), where);
    COMPAT_CUDA(cudaStreamSynchronize(c->stream), where);
    COMPAT_CUDA(cudaGetLastError(), where);
    for (int i = 0; i < n_particles; i++) p[i].p = h_tmp[i].p;                                  // :299
}

// :303-373 — writes du_dt_fluid / dv_dt_fluid only
void calculate_accelerations(float *du_dt_fluid, float *dv_dt_fluid, struct particle *fluid, struct particle *boundary,
                             struct neighbors_context *ctx_fluid, struct neighbors_context *ctx_boundary,
                             float gravity_x, float gravity_y)
{
    const char *where = "calculate_accelerations";
    require_built(ctx_fluid, where);
    require_built(ctx_boundary, where);
    sphb_ctx *c = ctx_fluid->core;
    const unsigned int moved = refresh(ctx_fluid, reinterpret_cast<sphb_particle *>(fluid), where);
    refresh(ctx_boundary, reinterpret_cast<sphb_particle *>(boundary), where);
    c->launches += launch_force(c->stream, c->k, c->fluid, ctx_boundary->core->fluid, gravity_x, gravity_y, nullptr,
                                false, c->d_counters, moved == 0, nullptr, c->prm.fast_force != 0);
    const int n = ctx_fluid->n_particles;
    if (ensure_stage(c, (size_t)n * (sizeof(sphb_particle) + 2 * sizeof(float)) + 64)) die(where);
    float *d_du = static_cast<float *>(c->d_stage), *d_dv = d_du + n;
    c->launches += launch_soa_to_aos(c->stream, c->fluid, nullptr, d_du, d_dv, false);
    COMPAT_CUDA(cudaMemcpyAsync(du_dt_fluid, d_du, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream), where);
    COMPAT_CUDA(cudaMemcpyAsync(dv_dt_fluid, d_dv, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream), where);
    COMPAT_CUDA(cudaStreamSynchronize(c->stream), where);
    COMPAT_CUDA(cudaGetLastError(), where);
}

// :380-411 — every pixel's bit of the 1 KiB frame is (re)written
void draw_metaballs(unsigned char *draw_buffer, struct particle *pixel_pseudoparticles, struct particle *fluid,
                    struct neighbors_context *ctx_fluid)
{
    const char *where = "draw_metaballs";
    require_built(ctx_fluid, where);
    sphb_ctx *c = ctx_fluid->core;
    refresh(ctx_fluid, reinterpret_cast<sphb_particle *>(fluid), where);
    const sphb_particle *px = reinterpret_cast<const sphb_particle *>(pixel_pseudoparticles);
    if (!c->d_pixels) {
        COMPAT_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->d_pixels), sizeof(float2) * 64 * 128), where);
        COMPAT_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->d_frame), 1024), where);
    }
    // the pixel centres travel to HBM only when they differ from the ones uploaded last (the reference builds
    // them once, :570-577, and hands the same array to every frame)
    bool fresh = ctx_fluid->h_pixels == nullptr;
    if (fresh) {
        ctx_fluid->h_pixels = static_cast<float2 *>(malloc(sizeof(float2) * 64 * 128));
        if (!ctx_fluid->h_pixels) { set_error("out of memory"); die(where); }
    }
    for (int i = 0; i < 64 * 128; i++) {
        const float2 v = make_float2(px[i].x, px[i].y);
        if (fresh || memcmp(&v, &ctx_fluid->h_pixels[i], sizeof v) != 0) { ctx_fluid->h_pixels[i] = v; fresh = true; }
    }
    if (fresh)
        COMPAT_CUDA(cudaMemcpyAsync(c->d_pixels, ctx_fluid->h_pixels, sizeof(float2) * 64 * 128, cudaMemcpyHostToDevice, c->stream), where);
    const float px_width = c->prm.width / 128;                  // :399
    const float W_px = host_W(c->prm.H, px_width / 2);           // :401
    c->launches += launch_render(c->stream, c->k, c->fluid, c->d_pixels, W_px, c->d_frame);
    COMPAT_CUDA(cudaMemcpyAsync(draw_buffer, c->d_frame, 1024, cudaMemcpyDeviceToHost, c->stream), where);
    COMPAT_CUDA(cudaStreamSynchronize(c->stream), where);
    COMPAT_CUDA(cudaGetLastError(), where);
}

}  // extern "C"

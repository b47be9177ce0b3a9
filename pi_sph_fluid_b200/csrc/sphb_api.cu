// sphb_api.cu — resident tier of the C ABI (include/sph_b200.h): owns the device state and
// sequences the kernels of one leapfrog step exactly as the reference's main loop does
// (pi_sph_fluid.c:600-607 init, :612-641 step).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "sphb_internal.cuh"
#include "sph_consts.h"

namespace sphb {

int density_force(sphb_ctx *c, float gx, float gy, const float2 *g_dev, bool kick2, const StepStats *ss = nullptr);

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return SPHB_E_CUDA;
}

static void free_set(ParticleSet &ps)
{
    for (int i = 0; i < 2; i++) {
        cudaFree(ps.pos[i]); cudaFree(ps.vel[i]); cudaFree(ps.id[i]); cudaFree(ps.mass[i]); cudaFree(ps.aux[i]);
    }
    cudaFree(ps.acc); cudaFree(ps.rho_prr); cudaFree(ps.p); cudaFree(ps.key); cudaFree(ps.rank);
    cudaFree(ps.ids_tmp); cudaFree(ps.cell_count); cudaFree(ps.cell_start); cudaFree(ps.cellkey);
    cudaFree(ps.cell_start_prev); cudaFree(ps.cell_touch);
    cudaFree(ps.nbr_list); cudaFree(ps.nbr_count); cudaFree(ps.chunk_rec); cudaFree(ps.chunk_queue);
    ps = ParticleSet();
}

template <class T>
static cudaError_t dmalloc(T **p, size_t count)
{
    return cudaMalloc(reinterpret_cast<void **>(p), (count ? count : 1) * sizeof(T));
}

// (re)allocate a set for n particles; arrays carry 2 padding slots so 16-byte-aligned tile
// copies may touch one element past either end of a run
int alloc_set(ParticleSet &ps, int n, int ncells, bool is_boundary, bool need_mass)
{
    if (ps.cap >= n && ps.cell_count && (need_mass ? ps.mass[0] != nullptr : true) &&
        (is_boundary ? ps.aux[0] != nullptr : ps.acc != nullptr)) {
        ps.n = n;
        ps.touch_min_slots = touch_min_slots();
        return SPHB_OK;
    }
    free_set(ps);
    ps.touch_min_slots = touch_min_slots();
    const size_t m = (size_t)n + 4;
    for (int i = 0; i < 2; i++) {
        SPHB_CUDA(dmalloc(&ps.pos[i], m));
        SPHB_CUDA(dmalloc(&ps.vel[i], m));
        SPHB_CUDA(dmalloc(&ps.id[i], m));
        SPHB_CUDA(cudaMemset(ps.pos[i], 0, m * sizeof(float2)));
        SPHB_CUDA(cudaMemset(ps.vel[i], 0, m * sizeof(float2)));
        if (need_mass) { SPHB_CUDA(dmalloc(&ps.mass[i], m)); SPHB_CUDA(cudaMemset(ps.mass[i], 0, m * sizeof(float))); }
        if (is_boundary) SPHB_CUDA(dmalloc(&ps.aux[i], m));
    }
    if (!is_boundary) {
        SPHB_CUDA(dmalloc(&ps.acc, m));
        SPHB_CUDA(dmalloc(&ps.rho_prr, m));
        SPHB_CUDA(dmalloc(&ps.p, m));
        SPHB_CUDA(cudaMemset(ps.rho_prr, 0, m * sizeof(float2)));
        const size_t ctas = ((size_t)n + kPairThreads - 1) / kPairThreads + 1;
        SPHB_CUDA(dmalloc(&ps.nbr_list, ctas * kListCap * kPairThreads));
        SPHB_CUDA(dmalloc(&ps.nbr_count, ctas * kPairThreads));
        SPHB_CUDA(dmalloc(&ps.chunk_rec, ctas * (size_t)kChunkRecWords));
        SPHB_CUDA(dmalloc(&ps.chunk_queue, 2));
        SPHB_CUDA(cudaMemset(ps.chunk_queue, 0, 2 * sizeof(unsigned long long)));
    }
    SPHB_CUDA(dmalloc(&ps.key, m));
    SPHB_CUDA(dmalloc(&ps.rank, m));
    SPHB_CUDA(dmalloc(&ps.ids_tmp, m));
    SPHB_CUDA(dmalloc(&ps.cellkey, m));
    SPHB_CUDA(dmalloc(&ps.cell_count, (size_t)ncells + 8));
    SPHB_CUDA(dmalloc(&ps.cell_start, (size_t)ncells + 8));
    SPHB_CUDA(cudaMemset(ps.cell_count, 0, ((size_t)ncells + 8) * sizeof(uint32_t)));
    SPHB_CUDA(cudaMemset(ps.cell_start, 0, ((size_t)ncells + 8) * sizeof(uint32_t)));
    SPHB_CUDA(dmalloc(&ps.cell_start_prev, (size_t)ncells + 8));
    SPHB_CUDA(dmalloc(&ps.cell_touch, (size_t)ncells + 8));
    SPHB_CUDA(cudaMemset(ps.cell_start_prev, 0, ((size_t)ncells + 8) * sizeof(uint32_t)));
    SPHB_CUDA(cudaMemset(ps.cell_touch, 0, (size_t)ncells + 8));
    // the memsets above ran on the legacy default stream, which does not order against the
    // handle's non-blocking stream
    SPHB_CUDA(cudaDeviceSynchronize());
    ps.n = n;
    ps.cap = n;
    return SPHB_OK;
}

int ensure_stage(sphb_ctx *c, size_t bytes)
{
    if (c->stage_bytes >= bytes) return SPHB_OK;
    cudaFree(c->d_stage);
    c->d_stage = nullptr;
    c->stage_bytes = 0;
    SPHB_CUDA(cudaMalloc(&c->d_stage, bytes));
    c->stage_bytes = bytes;
    return SPHB_OK;
}

bool pdl_enabled()
{
    static const bool on = [] { const char *e = getenv("SPHB_NO_PDL"); return !(e && *e && *e != '0'); }();
    return on;
}

// ---- profiling ------------------------------------------------------------------------

static int prof_drain(sphb_ctx *c)
{
    if (c->ev_used == 0) return SPHB_OK;
    SPHB_CUDA(cudaEventSynchronize(c->ev[2 * (c->ev_used - 1) + 1]));
    for (int i = 0; i < c->ev_used; i++) {
        float ms = 0.0f;
        SPHB_CUDA(cudaEventElapsedTime(&ms, c->ev[2 * i], c->ev[2 * i + 1]));
        c->prof_ms[c->ev_kind[i]] += (double)ms;
        c->prof_launches[c->ev_kind[i]] += 1;
    }
    c->ev_used = 0;
    return SPHB_OK;
}

struct ProfScope {
    sphb_ctx *c;
    int slot;
    ProfScope(sphb_ctx *ctx, int kind) : c(ctx), slot(-1)
    {
        if (!c->profile_mode) return;
        if (c->ev_used == 64) prof_drain(c);
        slot = c->ev_used++;
        c->ev_kind[slot] = kind;
        cudaEventRecord(c->ev[2 * slot], c->stream);
    }
    ~ProfScope()
    {
        if (slot >= 0) cudaEventRecord(c->ev[2 * slot + 1], c->stream);
    }
};

// ---- step pieces ----------------------------------------------------------------------

// update_neighbors_context (:104-124) for one set, optionally fused with kick+drift
int build_grid(sphb_ctx *c, ParticleSet &ps, bool advect, const Consts *kk, const StepStats *deliver)
{
    const Consts &k = kk ? *kk : c->k;
    // the boundary's builds count into their own block, so the fluid's per-build counters stay the fluid's
    DeviceCounters *ctr = (&ps == &c->boundary) ? c->d_counters + 1 : c->d_counters;
    { ProfScope p(c, SPHB_K_ADVECT_BIN); c->launches += launch_advect_bin(c->stream, k, ps, advect, ctr, nullptr, deliver); }
    { ProfScope p(c, SPHB_K_SCAN); c->launches += launch_scan(c->stream, k, ps, c->scan, ctr); }
    { ProfScope p(c, SPHB_K_REORDER); c->launches += launch_reorder(c->stream, k, ps, c->prm.deterministic != 0); }
    return SPHB_OK;
}

// One step in two halves so that a slab exchange fits between them (sphb_mg.cu):
//   phase A  kick + drift + binning of the resident particles (:615-626); on a slab it also
//            appends the halo / migration messages for both neighbours
//   phase B  [slab: bin the received entries] scan, reorder, density+pressure, accelerations, kick
int step_phase_a(sphb_ctx *c, bool advect, const StepStats *deliver)
{
    ProfScope p(c, SPHB_K_ADVECT_BIN);
    if (c->mg.on) {
        const SlabIO io = mg_slab_io(c);
        c->launches += launch_advect_bin(c->stream, c->k, c->fluid, advect, c->d_counters, &io, deliver);
    } else {
        c->launches += launch_advect_bin(c->stream, c->k, c->fluid, advect, c->d_counters, nullptr, deliver);
    }
    return SPHB_OK;
}

int step_phase_b(sphb_ctx *c, float gx, float gy, bool kick2, const StepStats *ss)
{
    if (c->mg.on) {
        ProfScope p(c, SPHB_K_OTHER);
        const SlabIO io = mg_slab_io(c);
        c->launches += launch_bin_recv(c->stream, c->k, c->fluid, io, c->d_counters);
        c->mg.exchanges++;
    }
    { ProfScope p(c, SPHB_K_SCAN); c->launches += launch_scan(c->stream, c->k, c->fluid, c->scan, c->d_counters); }
    { ProfScope p(c, SPHB_K_REORDER); c->launches += launch_reorder(c->stream, c->k, c->fluid, c->prm.deterministic != 0); }
    return density_force(c, gx, gy, nullptr, kick2, ss);
}

int free_set_public(ParticleSet &ps) { free_set(ps); return SPHB_OK; }

int density_force(sphb_ctx *c, float gx, float gy, const float2 *g_dev, bool kick2, const StepStats *ss_in)
{
    StepStats ss_here;
    const StepStats *ss = nullptr;
    if (ss_in) {
        ss_here = *ss_in;
        ss_here.id = c->fluid.id[c->fluid.ic];      // the sorted order this step's reorder produced
        ss = &ss_here;
    }
    { ProfScope p(c, SPHB_K_DENSITY); c->launches += launch_density(c->stream, c->k, c->fluid, c->boundary, c->d_counters, false, true, ss ? ss->block : nullptr, ss ? ss->flags : nullptr); }
    { ProfScope p(c, SPHB_K_FORCE); c->launches += launch_force(c->stream, c->k, c->fluid, c->boundary, gx, gy, g_dev, kick2, c->d_counters, true, ss, c->prm.fast_force != 0); }
    return SPHB_OK;
}

static int check_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); libsphb200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return SPHB_E_CUDA;
    }
    if (device < 0 || device >= count) { set_error("device %d out of range (%d devices)", device, count); return SPHB_E_ARG; }
    cudaDeviceProp prop;
    SPHB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10 || prop.minor != 0) {
        // sm_100a code is not forward compatible: an sm_103 part has no loadable image of these kernels
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return SPHB_E_CUDA;
    }
    return SPHB_OK;
}

}  // namespace sphb

using namespace sphb;

#define SPHB_ENTER(ctx)                                               \
    do {                                                              \
        if (!(ctx)) { set_error("null context"); return SPHB_E_ARG; } \
        SPHB_CUDA(cudaSetDevice((ctx)->device));                      \
    } while (0)

extern "C" {

const char *sphb_last_error(void) { return g_err; }

const char *sphb_build_info(void)
{
#define SPHB_STR2(x) #x
#define SPHB_STR(x) SPHB_STR2(x)
    return "libsphb200 v1: sm_100a, nvcc " __DATE__ ", pair CTA " SPHB_STR(SPHB_PT) " thr, tile " SPHB_STR(SPHB_TILE_CAP)
           ", list rows " SPHB_STR(SPHB_LIST_CAP) ", window " SPHB_STR(SPHB_WIN_CAP) ", persistent " SPHB_STR(SPHB_PERSISTENT)
           ", minb " SPHB_STR(SPHB_MINB_D) "/" SPHB_STR(SPHB_MINB_F);
}

int sphb_default_params(sphb_params *prm, float R, float width, float height)
{
    if (!prm) return SPHB_E_ARG;
    memset(prm, 0, sizeof *prm);
    prm->R = R;
    prm->H = R * 1.3f;                      // :12
    prm->width = width;
    prm->height = height;
    prm->rho0 = 1000.0f;
    prm->c0 = 400.0f;
    prm->g = 9.81f;
    prm->dt = 1.0f * prm->H / prm->c0;      // :19
    prm->vol = 0.57f * prm->H * prm->H;     // :20
    prm->x_min = 0; prm->x_max = width; prm->y_min = 0; prm->y_max = height;   // :595
    prm->cell_length = 2 * prm->H;          // :596
    prm->deterministic = 1;
    prm->device = 0;
    return SPHB_OK;
}

int sphb_gravity_from_raw(const sphb_params *prm, int accel_x_raw, int accel_y_raw, float *gx, float *gy)
{
    if (!prm || !gx || !gy) return SPHB_E_ARG;
    *gx = (float)accel_y_raw / (1 << 14) * prm->g;      // :439
    *gy = -(float)accel_x_raw / (1 << 14) * prm->g;     // :440
    return SPHB_OK;
}

static int create_fill(sphb_ctx *c, const sphb_params *prm)
{
    c->prm = *prm;
    c->device = prm->device;
    c->k = make_consts(*prm, prm->rho0 * prm->vol);
    cudaDeviceProp prop;
    SPHB_CUDA(cudaGetDeviceProperties(&prop, c->device));
    c->sm_count = prop.multiProcessorCount;
    SPHB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->d_counters), 2 * sizeof(DeviceCounters)));
    SPHB_CUDA(cudaMemset(c->d_counters, 0, 2 * sizeof(DeviceCounters)));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->d_gravity), sizeof(float2) * 4096));
    c->scan.n_tiles = (c->k.ncells + kScanTile - 1) / kScanTile;
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->scan.tile_state), sizeof(unsigned long long) * (c->scan.n_tiles + 1)));
    SPHB_CUDA(cudaMemset(c->scan.tile_state, 0, sizeof(unsigned long long) * (c->scan.n_tiles + 1)));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->scan.tile_counter), sizeof(unsigned long long)));
    SPHB_CUDA(cudaMemset(c->scan.tile_counter, 0, sizeof(unsigned long long)));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->d_stats), 128));
    SPHB_CUDA(cudaMemset(c->d_stats, 0, 128));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->d_step_slots), (size_t)2 * kStatsSlots * 128));      // two sets, by sequence parity
    SPHB_CUDA(cudaMemset(c->d_step_slots, 0, (size_t)2 * kStatsSlots * 128));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->d_stats_done), sizeof(unsigned int)));
    SPHB_CUDA(cudaMemset(c->d_stats_done, 0, sizeof(unsigned int)));
    SPHB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&c->h_stats), 512, cudaHostAllocMapped));     // two slots
    memset(c->h_stats, 0, 512);
    for (int i = 0; i < 128; i++) SPHB_CUDA(cudaEventCreate(&c->ev[i]));
    SPHB_CUDA(cudaDeviceSynchronize());      // memsets above vs. the non-blocking stream
    return SPHB_OK;
}

int sphb_create(const sphb_params *prm, sphb_ctx **out)
{
    if (!prm || !out) { set_error("null argument"); return SPHB_E_ARG; }
    *out = nullptr;
    if (!(prm->H > 0) || !(prm->cell_length > 0) || !(prm->x_max > prm->x_min) || !(prm->y_max > prm->y_min)) {
        set_error("bad geometry (H=%g cell=%g)", prm->H, prm->cell_length);
        return SPHB_E_ARG;
    }
    int rc = check_device(prm->device);
    if (rc) return rc;
    SPHB_CUDA(cudaSetDevice(prm->device));
    const double rows = floor(((double)prm->y_max - prm->y_min) / prm->cell_length) + 1;
    const double cols = floor(((double)prm->x_max - prm->x_min) / prm->cell_length) + 1;
    if (rows * cols > 2.0e9) { set_error("grid of %.0f x %.0f cells exceeds 2^31", rows, cols); return SPHB_E_ARG; }
    if (rows >= 65536.0 || cols >= 65536.0) { set_error("grid of %.0f x %.0f cells: a dimension exceeds 65535", rows, cols); return SPHB_E_ARG; }

    sphb_ctx *c = new (std::nothrow) sphb_ctx();
    if (!c) return SPHB_E_NOMEM;
    for (int i = 0; i < 128; i++) c->ev[i] = nullptr;
    rc = create_fill(c, prm);
    if (rc) {                 // nothing leaks on a failed create: sphb_destroy tolerates the members still null
        sphb_destroy(c);
        return rc;
    }
    *out = c;
    return SPHB_OK;
}

int sphb_destroy(sphb_ctx *c)
{
    if (!c) return SPHB_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    mg_free(c);
    free_set(c->fluid);
    free_set(c->boundary);
    cudaFree(c->d_counters); cudaFree(c->d_gravity); cudaFree(c->d_stage);
    cudaFree(c->d_pixels); cudaFree(c->d_frame); cudaFree(c->d_stats); cudaFree(c->d_l2_scratch);
    cudaFree(c->scan.tile_state); cudaFree(c->scan.tile_counter); cudaFree(c->d_stats_done); cudaFree(c->d_step_slots);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_stats) cudaFreeHost(c->h_stats);
    for (int i = 0; i < 128; i++)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
    return SPHB_OK;
}

int sphb_upload(sphb_ctx *c, const sphb_particle *fluid, int n_fluid, const sphb_particle *boundary, int n_boundary)
{
    SPHB_ENTER(c);
    if (c->mg.on) { set_error("this context is a slab of a multi-GPU run: use sphb_mg_upload"); return SPHB_E_STATE; }
    if (n_fluid < 0 || n_boundary < 0 || (n_fluid > 0 && !fluid) || (n_boundary > 0 && !boundary)) {
        set_error("bad particle arrays"); return SPHB_E_ARG;
    }
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    const size_t fb = (size_t)n_fluid * sizeof(sphb_particle), bb = (size_t)n_boundary * sizeof(sphb_particle);
    int rc = ensure_stage(c, (fb > bb ? fb : bb) + 16);
    if (rc) return rc;
    if (n_fluid > 0) SPHB_CUDA(cudaMemcpyAsync(c->d_stage, fluid, fb, cudaMemcpyHostToDevice, c->stream));
    // uniform fluid mass (the reference sets every m = RHO_0*V, :502) selects the kernels that keep the
    // mass in the constant bank; the conversion kernel itself reports whether the masses differ
    bool uniform = true;
    rc = alloc_set(c->fluid, n_fluid, c->k.ncells, false, c->fluid.mass[0] != nullptr);
    if (rc) return rc;
    if (n_fluid > 0) {
        uint32_t m0_bits;
        memcpy(&m0_bits, &fluid[0].m, sizeof m0_bits);
        unsigned int *d_differs = &c->d_counters[1].list_flushes;      // scratch word of the boundary's counter block
        unsigned int differs = 0;
        SPHB_CUDA(cudaMemsetAsync(d_differs, 0, sizeof(unsigned int), c->stream));
        c->launches += launch_aos_to_soa(c->stream, reinterpret_cast<const sphb_particle *>(c->d_stage), c->fluid, false, -1,
                                         nullptr, 0, m0_bits, d_differs);
        SPHB_CUDA(cudaMemcpyAsync(&differs, d_differs, sizeof differs, cudaMemcpyDeviceToHost, c->stream));
        SPHB_CUDA(cudaStreamSynchronize(c->stream));     // d_stage is reused below
        uniform = differs == 0;
        if (!uniform && c->fluid.mass[0] == nullptr) {
            // first set with per-particle masses: the arrays that carry them, and the conversion again
            rc = alloc_set(c->fluid, n_fluid, c->k.ncells, false, true);
            if (rc) return rc;
            c->launches += launch_aos_to_soa(c->stream, reinterpret_cast<const sphb_particle *>(c->d_stage), c->fluid, false);
            SPHB_CUDA(cudaStreamSynchronize(c->stream));
        }
    }
    c->fluid.uniform_mass = uniform;
    c->fluid.uniform_mass_value = n_fluid > 0 ? fluid[0].m : c->prm.rho0 * c->prm.vol;
    c->k.mass = c->fluid.uniform_mass_value;
    if (n_boundary > 0) {
        rc = alloc_set(c->boundary, n_boundary, c->k.ncells, true, true);
        if (rc) return rc;
        c->boundary.uniform_mass = false;
        SPHB_CUDA(cudaMemcpyAsync(c->d_stage, boundary, bb, cudaMemcpyHostToDevice, c->stream));
        c->launches += launch_aos_to_soa(c->stream, reinterpret_cast<const sphb_particle *>(c->d_stage), c->boundary, true);
        SPHB_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        c->boundary.n = 0;
        c->boundary.sorted = false;
    }
    c->boundary_ready = false;
    c->accel_ready = false;
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

int sphb_init_boundary(sphb_ctx *c)
{
    SPHB_ENTER(c);
    if (c->mg.on) return mg_init_boundary(c);
    if (c->boundary.n > 0) {
        build_grid(c, c->boundary, false);                                        // :600
        ProfScope p(c, SPHB_K_OTHER);
        c->launches += launch_pseudomass(c->stream, c->k, c->boundary);            // :601
    }
    c->boundary_ready = true;
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

int sphb_compute_accel(sphb_ctx *c, float gx, float gy)
{
    SPHB_ENTER(c);
    if (c->fluid.n <= 0) { set_error("no fluid uploaded"); return SPHB_E_STATE; }
    if (c->boundary.n > 0 && !c->boundary_ready) { set_error("sphb_init_boundary not called"); return SPHB_E_STATE; }
    if (c->mg.on) {
        if (c->mg.transport != 1 && c->mg.transport != 3) { set_error("slab not connected over NCCL or peer stores: use sphb_mg_group_compute_accel"); return SPHB_E_STATE; }
        step_phase_a(c, false);
        int rc = mg_exchange(c);
        if (rc) return rc;
        step_phase_b(c, gx, gy, false);
    } else {
        build_grid(c, c->fluid, false);                      // :604
        density_force(c, gx, gy, nullptr, false);            // :605-607
    }
    c->accel_ready = true;
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

int sphb_upload_accel(sphb_ctx *c, const float *du_dt, const float *dv_dt)
{
    SPHB_ENTER(c);
    if (c->mg.on) { set_error("not available on a slab context"); return SPHB_E_STATE; }
    const int n = c->fluid.n;
    if (n <= 0) { set_error("no fluid uploaded"); return SPHB_E_STATE; }
    if (!du_dt || !dv_dt) return SPHB_E_ARG;
    const size_t db = (size_t)n * sizeof(float);
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    int rc = ensure_stage(c, 2 * db + 64);
    if (rc) return rc;
    float *d_du = static_cast<float *>(c->d_stage), *d_dv = d_du + n;
    SPHB_CUDA(cudaMemcpyAsync(d_du, du_dt, db, cudaMemcpyHostToDevice, c->stream));
    SPHB_CUDA(cudaMemcpyAsync(d_dv, dv_dt, db, cudaMemcpyHostToDevice, c->stream));
    c->launches += launch_set_accel(c->stream, c->fluid, d_du, d_dv);
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaGetLastError());
    c->accel_ready = true;
    return SPHB_OK;
}

struct StatsBlock { double d[4]; unsigned int u[16]; };     // the 128-byte block k_stats / k_force<STATS> fill
static void decode_stats(const sphb_ctx *c, const StatsBlock *h, sphb_stats *out);

// The force pass of the LAST step also reduces the step statistics when the caller asks for them and its
// last CTA stores them into mapped host memory; the host waits for the sequence word there (no copy, no
// stream synchronisation).  Two host slots alternate by sequence number, so the statistics of step s can
// still be on their way while step s + 1 is launched (sphb_step_stats_begin / _end).
constexpr int kStatsSlotWords = 32;      // 16 block words, the sequence word, padding: 256 bytes per slot

// defer: the force pass only sums (sphb_step_stats_begin); folding the slots and handing the block to the
// host is left to CTA 0 of the next advect+bin kernel on the stream, or to stats_collect when none follows.
static int step_launch(sphb_ctx *c, float gx, float gy, const float *trace, int nsteps, bool want_stats,
                       unsigned long long *ticket_out, bool defer = false)
{
    SPHB_ENTER(c);
    if (nsteps < 0) return SPHB_E_ARG;
    if (c->fluid.n <= 0) { set_error("no fluid uploaded"); return SPHB_E_STATE; }
    if (!c->accel_ready) { set_error("sphb_compute_accel must run before sphb_step (:604-607 precede :610)"); return SPHB_E_STATE; }
    if (c->mg.on && c->mg.transport != 1 && c->mg.transport != 3) { set_error("slab not connected over NCCL or peer stores: use sphb_mg_group_step"); return SPHB_E_STATE; }
    StepStats ss;
    want_stats = want_stats && nsteps > 0;
    if (want_stats) {
        if (c->stats_seq - c->stats_collected >= 2) {
            set_error("two sets of step statistics are outstanding: collect one with sphb_step_stats_end first");
            return SPHB_E_STATE;
        }
        ss.ctr = c->d_counters;
        ss.flags = c->mg.on ? c->mg.d_flags : nullptr;
        ss.last_id = c->fluid.windowed ? 0xffffffffu : (uint32_t)(c->fluid.n - 1);
        ss.done = defer ? nullptr : c->d_stats_done;
        ss.seq = ++c->stats_seq;
        // device slots and host slot alternate with the sequence number: request s + 2 reuses what request
        // s used, and s has been collected by then (two outstanding at most)
        ss.block = c->d_step_slots + (size_t)(ss.seq & 1ULL) * kStatsSlots * 16;
        ss.host = c->h_stats + kStatsSlotWords * (ss.seq & 1ULL);
        if (ticket_out) *ticket_out = ss.seq;
    }
    for (int s = 0; s < nsteps; s++) {
        if (trace) { gx = trace[2 * s]; gy = trace[2 * s + 1]; }    // the value every thread reads at :632
        const bool with_stats = want_stats && s == nsteps - 1;
        // statistics a previous sphb_step_stats_begin left undelivered ride on this step's first kernel
        StepStats pending;
        const StepStats *deliver = nullptr;
        if (c->stats_has_pending) { pending = c->stats_pending; deliver = &pending; c->stats_has_pending = false; }
        if (c->mg.on) {
            step_phase_a(c, true, deliver);
            int rc = mg_exchange(c);
            if (rc) {
                // the ticket of this call is never delivered: take it back so the statistics path stays usable
                // (no force pass has added to its slots yet: only the last step's does)
                if (want_stats) { c->stats_seq--; if (ticket_out) *ticket_out = 0; }
                return rc;
            }
            step_phase_b(c, gx, gy, true, with_stats ? &ss : nullptr);
        } else {
            build_grid(c, c->fluid, true, nullptr, deliver);                   // :615-626
            density_force(c, gx, gy, nullptr, true, with_stats ? &ss : nullptr);         // :630-640
        }
        c->steps++;
    }
    if (want_stats) c->stats_steps[ss.seq & 1ULL] = c->steps;
    if (want_stats && defer) { c->stats_pending = ss; c->stats_has_pending = true; }
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

// waits for the statistics with sequence number `ticket` (the oldest outstanding one) and decodes them
static int stats_collect(sphb_ctx *c, unsigned long long ticket, sphb_stats *stats_out)
{
    if (ticket != c->stats_collected + 1 || ticket > c->stats_seq) {
        set_error("step statistics are collected in the order they were requested (next: %llu, asked: %llu)",
                  c->stats_collected + 1, ticket);
        return SPHB_E_STATE;
    }
    if (c->stats_has_pending && c->stats_pending.seq == ticket) {
        // no step followed the request: deliver with a kernel of its own
        c->launches += launch_stats_deliver(c->stream, c->stats_pending);
        c->stats_has_pending = false;
        SPHB_CUDA(cudaGetLastError());
    }
    const unsigned long long *slot = c->h_stats + kStatsSlotWords * (ticket & 1ULL);
    volatile const unsigned long long *seq = slot + 16;
    unsigned long long spins = 0;
    while (*seq != ticket) {
        if ((++spins & 0x3ffULL) == 0) {
            const cudaError_t q = cudaStreamQuery(c->stream);
            if (q == cudaSuccess) {
                if (*seq == ticket) break;
                set_error("step statistics were not delivered");
                return SPHB_E_STATE;
            }
            if (q != cudaErrorNotReady) return cuda_fail(q, "cudaStreamQuery", __FILE__, __LINE__);
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    StatsBlock h;
    memcpy(&h, slot, sizeof h);
    decode_stats(c, &h, stats_out);
    stats_out->steps = c->stats_steps[ticket & 1ULL];     // the step count when these were requested
    c->stats_collected = ticket;
    if (c->mg.on && (stats_out->n_overflow & (1u << 30))) { c->mg.comm_failed = true; return mg_health(c); }
    return SPHB_OK;
}

static int step_impl(sphb_ctx *c, float gx, float gy, const float *trace, int nsteps, sphb_stats *stats_out = nullptr)
{
    if (stats_out && c && c->stats_seq != c->stats_collected) {
        set_error("step statistics requested with sphb_step_stats_begin are still outstanding");
        return SPHB_E_STATE;
    }
    unsigned long long ticket = 0;
    int rc = step_launch(c, gx, gy, trace, nsteps, stats_out != nullptr, &ticket);
    if (rc || !stats_out || nsteps <= 0) return rc;
    return stats_collect(c, ticket, stats_out);
}

int sphb_step(sphb_ctx *c, float gx, float gy, int nsteps) { return step_impl(c, gx, gy, nullptr, nsteps); }

int sphb_step_trace(sphb_ctx *c, const float *gravity_xy, int nsteps)
{
    if (!gravity_xy && nsteps > 0) return SPHB_E_ARG;
    return step_impl(c, 0.0f, 0.0f, gravity_xy, nsteps);
}

int sphb_step_stats(sphb_ctx *c, const float *gravity_xy, int nsteps, sphb_stats *out)
{
    if (!out || nsteps < 1 || !gravity_xy) { set_error("sphb_step_stats: needs nsteps >= 1, a gravity sample per step and an output"); return SPHB_E_ARG; }
    return step_impl(c, 0.0f, 0.0f, gravity_xy, nsteps, out);
}

int sphb_step_stats_begin(sphb_ctx *c, const float *gravity_xy, int nsteps, unsigned long long *ticket_out)
{
    if (!ticket_out || nsteps < 1 || !gravity_xy) { set_error("sphb_step_stats_begin: needs nsteps >= 1, a gravity sample per step and a ticket"); return SPHB_E_ARG; }
    return step_launch(c, 0.0f, 0.0f, gravity_xy, nsteps, true, ticket_out, true);
}

int sphb_step_stats_end(sphb_ctx *c, unsigned long long ticket, sphb_stats *out)
{
    SPHB_ENTER(c);
    if (!out) return SPHB_E_ARG;
    return stats_collect(c, ticket, out);
}

int sphb_synchronize(sphb_ctx *c)
{
    SPHB_ENTER(c);
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaGetLastError());
    return mg_health(c);
}

int sphb_download(sphb_ctx *c, sphb_particle *fluid_out, float *du_dt, float *dv_dt)
{
    SPHB_ENTER(c);
    if (c->mg.on) { set_error("this context is a slab of a multi-GPU run: use sphb_mg_download"); return SPHB_E_STATE; }
    const int n = c->fluid.n;
    if (n <= 0) return SPHB_OK;
    if ((du_dt == nullptr) != (dv_dt == nullptr)) { set_error("du_dt and dv_dt go together"); return SPHB_E_ARG; }
    const size_t ab = (size_t)n * sizeof(sphb_particle), db = (size_t)n * sizeof(float);
    int rc = ensure_stage(c, ab + 2 * db + 64);
    if (rc) return rc;
    char *base = static_cast<char *>(c->d_stage);
    sphb_particle *d_aos = reinterpret_cast<sphb_particle *>(base);
    float *d_du = reinterpret_cast<float *>(base + ((ab + 15) & ~(size_t)15));
    float *d_dv = d_du + n;
    c->launches += launch_soa_to_aos(c->stream, c->fluid, fluid_out ? d_aos : nullptr, du_dt ? d_du : nullptr,
                                     du_dt ? d_dv : nullptr, false);
    if (fluid_out) SPHB_CUDA(cudaMemcpyAsync(fluid_out, d_aos, ab, cudaMemcpyDeviceToHost, c->stream));
    if (du_dt) {
        SPHB_CUDA(cudaMemcpyAsync(du_dt, d_du, db, cudaMemcpyDeviceToHost, c->stream));
        SPHB_CUDA(cudaMemcpyAsync(dv_dt, d_dv, db, cudaMemcpyDeviceToHost, c->stream));
    }
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

int sphb_download_boundary(sphb_ctx *c, sphb_particle *boundary_out)
{
    SPHB_ENTER(c);
    const int n = c->boundary.n;
    if (n <= 0 || !boundary_out) return SPHB_OK;
    const size_t ab = (size_t)n * sizeof(sphb_particle);
    int rc = ensure_stage(c, ab + 64);
    if (rc) return rc;
    c->launches += launch_soa_to_aos(c->stream, c->boundary, reinterpret_cast<sphb_particle *>(c->d_stage), nullptr, nullptr, true);
    SPHB_CUDA(cudaMemcpyAsync(boundary_out, c->d_stage, ab, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

int sphb_grid_shape(sphb_ctx *c, int *rows, int *cols)
{
    if (!c) return SPHB_E_ARG;
    if (rows) *rows = c->k.rows;
    if (cols) *cols = c->k.cols;
    return SPHB_OK;
}

int sphb_cell_ids(sphb_ctx *c, int *cell_out)
{
    SPHB_ENTER(c);
    const int n = c->fluid.n;
    if (n <= 0 || !cell_out) return SPHB_E_ARG;
    int rc = ensure_stage(c, (size_t)n * sizeof(int) + 64);
    if (rc) return rc;
    c->launches += launch_cell_ids(c->stream, c->k, c->fluid, static_cast<int *>(c->d_stage));
    SPHB_CUDA(cudaMemcpyAsync(cell_out, c->d_stage, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    return SPHB_OK;
}

int sphb_neighbor_lists(sphb_ctx *c, int which, int cap, int *counts, int *lists)
{
    SPHB_ENTER(c);
    if (which < 0 || which > 2 || cap <= 0 || !counts || !lists) return SPHB_E_ARG;
    const ParticleSet &a = which == 2 ? c->boundary : c->fluid;
    const ParticleSet &b = which == 0 ? c->fluid : c->boundary;
    if (a.n <= 0) return SPHB_E_STATE;
    if (!b.sorted || (which == 0 && !a.sorted)) { set_error("grid not built yet"); return SPHB_E_STATE; }
    const size_t cb = (size_t)a.n * sizeof(int), lb = (size_t)a.n * cap * sizeof(int);
    int rc = ensure_stage(c, cb + lb + 64);
    if (rc) return rc;
    int *d_counts = static_cast<int *>(c->d_stage);
    int *d_lists = d_counts + (((size_t)a.n + 3) & ~(size_t)3);
    unsigned int *d_over = &c->d_counters->list_flushes;   // borrowed as a scratch counter
    SPHB_CUDA(cudaMemsetAsync(d_over, 0, sizeof(unsigned int), c->stream));
    SPHB_CUDA(cudaMemsetAsync(d_lists, 0xff, lb, c->stream));
    c->launches += launch_neighbor_lists(c->stream, c->k, a, b, which != 1, cap, d_counts, d_lists, d_over);
    unsigned int over = 0;
    SPHB_CUDA(cudaMemcpyAsync(counts, d_counts, cb, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaMemcpyAsync(lists, d_lists, lb, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaMemcpyAsync(&over, d_over, sizeof over, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaMemsetAsync(d_over, 0, sizeof(unsigned int), c->stream));
    return (int)over;
}

int sphb_handover_lists(sphb_ctx *c, int cap, int *counts, int *lists, unsigned int *n_whole_chunks)
{
    SPHB_ENTER(c);
    if (cap <= 0 || !counts || !lists) return SPHB_E_ARG;
    const ParticleSet &f = c->fluid;
    if (f.n <= 0 || c->mg.on) { set_error("sphb_handover_lists: needs an uploaded single-GPU context"); return SPHB_E_STATE; }
    if (!f.sorted || !f.lists_valid) { set_error("no handed-over lists: run sphb_compute_accel or sphb_step first"); return SPHB_E_STATE; }
    const size_t cb = (((size_t)f.n + 3) & ~(size_t)3) * sizeof(int), lb = (size_t)f.n * cap * sizeof(int);
    int rc = ensure_stage(c, cb + lb + 64);
    if (rc) return rc;
    int *d_counts = static_cast<int *>(c->d_stage);
    int *d_lists = d_counts + (((size_t)f.n + 3) & ~(size_t)3);
    unsigned int *d_fast = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(c->d_stage) + cb + lb);
    SPHB_CUDA(cudaMemsetAsync(d_lists, 0xff, lb, c->stream));
    SPHB_CUDA(cudaMemsetAsync(d_fast, 0, sizeof(unsigned int), c->stream));
    c->launches += launch_decode_handover(c->stream, c->k, f, cap, d_counts, d_lists, d_fast);
    unsigned int fast = 0;
    SPHB_CUDA(cudaMemcpyAsync(counts, d_counts, (size_t)f.n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaMemcpyAsync(lists, d_lists, lb, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaMemcpyAsync(&fast, d_fast, sizeof fast, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaGetLastError());
    if (n_whole_chunks) *n_whole_chunks = fast;
    return SPHB_OK;
}

int sphb_probe_force_pair(sphb_ctx *c, int n, const float *pairs, int variant, float *out_txy, int *exact_shortcuts)
{
    SPHB_ENTER(c);
    if (n < 0 || (n > 0 && (!pairs || !out_txy)) || variant < 0 || variant > 6) return SPHB_E_ARG;
    if (exact_shortcuts) *exact_shortcuts = (c->k.div_exact && c->k.wref_div_exact && c->k.visc_pow2) ? 1 : 0;
    if (n == 0) return SPHB_OK;
    const size_t ib = (size_t)n * 12 * sizeof(float), ob = (size_t)n * 2 * sizeof(float);
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    int rc = ensure_stage(c, ib + ob + 64);
    if (rc) return rc;
    float *d_in = static_cast<float *>(c->d_stage), *d_out = d_in + (size_t)n * 12;
    SPHB_CUDA(cudaMemcpyAsync(d_in, pairs, ib, cudaMemcpyHostToDevice, c->stream));
    Consts k = c->k;
    if (c->fluid.n == 0) k.mass = c->prm.rho0 * c->prm.vol;
    c->launches += launch_probe_force_pair(c->stream, k, n, d_in, variant, d_out);
    SPHB_CUDA(cudaMemcpyAsync(out_txy, d_out, ob, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

int sphb_pair_stats(sphb_ctx *c, double *cand, double *acc)
{
    SPHB_ENTER(c);
    if (c->fluid.n <= 0 || !c->fluid.sorted) return SPHB_E_STATE;
    DeviceCounters h;
    SPHB_CUDA(cudaMemsetAsync(&c->d_counters->pair_candidates, 0, 2 * sizeof(unsigned long long), c->stream));
    // a counting density pass writes the same rho/p it would anyway
    c->launches += launch_density(c->stream, c->k, c->fluid, c->boundary, c->d_counters, true);
    SPHB_CUDA(cudaMemcpyAsync(&h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    int n = c->fluid.n;
    if (c->fluid.d_n_cur) SPHB_CUDA(cudaMemcpyAsync(&n, c->fluid.d_n_cur, sizeof n, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    if (n < 1) n = 1;
    if (cand) *cand = (double)h.pair_candidates / n;
    if (acc) *acc = (double)h.pair_accepted / n;
    return SPHB_OK;
}

int sphb_profile(sphb_ctx *c, int mode)
{
    SPHB_ENTER(c);
    int rc = prof_drain(c);
    c->profile_mode = mode;
    return rc;
}

int sphb_profile_read(sphb_ctx *c, double ms[SPHB_K_COUNT], unsigned long long launches[SPHB_K_COUNT], int reset)
{
    SPHB_ENTER(c);
    int rc = prof_drain(c);
    if (rc) return rc;
    for (int i = 0; i < SPHB_K_COUNT; i++) {
        if (ms) ms[i] = c->prof_ms[i];
        if (launches) launches[i] = c->prof_launches[i];
        if (reset) { c->prof_ms[i] = 0; c->prof_launches[i] = 0; }
    }
    return SPHB_OK;
}

int sphb_render(sphb_ctx *c, unsigned char *draw_buffer)
{
    SPHB_ENTER(c);
    if (!draw_buffer) return SPHB_E_ARG;
    if (c->fluid.n <= 0 || !c->fluid.sorted) { set_error("no sorted fluid state to render"); return SPHB_E_STATE; }
    if (!c->d_pixels) {
        // pixel-centre pseudo-particles, :570-577 (double expression, rounded to float)
        float2 *h = static_cast<float2 *>(malloc(sizeof(float2) * 64 * 128));
        if (!h) return SPHB_E_NOMEM;
        for (int i = 0; i < 64; i++)
            for (int j = 0; j < 128; j++) {
                h[i * 128 + j].x = (float)((j + 0.5) * (double)c->prm.width / 128);
                h[i * 128 + j].y = (float)((64 - (i + 0.5)) * (double)c->prm.height / 64);
            }
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&c->d_pixels), sizeof(float2) * 64 * 128);
        if (e == cudaSuccess) e = cudaMemcpy(c->d_pixels, h, sizeof(float2) * 64 * 128, cudaMemcpyHostToDevice);
        free(h);
        SPHB_CUDA(e);
        SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->d_frame), 1024));
    }
    const float px_width = c->prm.width / 128;                 // :399
    const float W_px = host_W(c->prm.H, px_width / 2);          // :401
    c->launches += launch_render(c->stream, c->k, c->fluid, c->d_pixels, W_px, c->d_frame);
    SPHB_CUDA(cudaMemcpyAsync(draw_buffer, c->d_frame, 1024, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

int sphb_render_counts(sphb_ctx *c, unsigned int *counts)
{
    SPHB_ENTER(c);
    if (!counts) return SPHB_E_ARG;
    if (c->fluid.n <= 0) { set_error("no fluid uploaded"); return SPHB_E_STATE; }
    int rc = ensure_stage(c, 64 * 128 * sizeof(unsigned int) + 64);
    if (rc) return rc;
    unsigned int *d = static_cast<unsigned int *>(c->d_stage);
    SPHB_CUDA(cudaMemsetAsync(d, 0, 64 * 128 * sizeof(unsigned int), c->stream));
    c->launches += launch_pixel_counts(c->stream, c->k, c->fluid, c->prm.x_max - c->prm.x_min, c->prm.y_max - c->prm.y_min, d);
    SPHB_CUDA(cudaMemcpyAsync(counts, d, 64 * 128 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

int sphb_splat_frame(const sphb_params *prm, const unsigned int *counts, unsigned char *draw_buffer)
{
    if (!prm || !counts || !draw_buffer) return SPHB_E_ARG;
    // lit when the fluid volume inside the pixel, count * V (:20), covers at least half of the pixel
    const double pixel = ((double)prm->x_max - prm->x_min) / 128 * (((double)prm->y_max - prm->y_min) / 64);
    const double need = 0.5 * pixel / (double)prm->vol;
    memset(draw_buffer, 0, 1024);
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 128; j++)
            if ((double)counts[i * 128 + j] >= need) draw_buffer[(i / 8) * 128 + j] |= (unsigned char)(1u << (i % 8));      // :407
    return SPHB_OK;
}

int sphb_render_splat(sphb_ctx *c, unsigned char *draw_buffer)
{
    if (!draw_buffer) return SPHB_E_ARG;
    unsigned int *counts = static_cast<unsigned int *>(malloc(64 * 128 * sizeof(unsigned int)));
    if (!counts) return SPHB_E_NOMEM;
    int rc = sphb_render_counts(c, counts);
    if (!rc) rc = sphb_splat_frame(&c->prm, counts, draw_buffer);
    free(counts);
    return rc;
}

static float key_to_float(unsigned int key)
{
    const unsigned int u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    float f;
    memcpy(&f, &u, sizeof f);
    return f;
}

static void decode_stats(const sphb_ctx *c, const StatsBlock *h, sphb_stats *out)
{
    memset(out, 0, sizeof *out);
    out->mass = h->d[0]; out->mom_x = h->d[1]; out->mom_y = h->d[2]; out->kinetic = h->d[3];
    memcpy(&out->max_speed, &h->u[0], sizeof(float));
    out->n_fluid = c->mg.on ? h->u[4] : (unsigned int)c->fluid.n;
    if (out->n_fluid > 0) {
        out->max_rho = key_to_float(h->u[1]);
        out->min_rho = key_to_float(~h->u[2]);
        out->max_rho_err = out->max_rho - c->prm.rho0;
        float last;
        memcpy(&last, &h->u[3], sizeof last);
        out->last_rho_err_ref = last - c->prm.rho0;      // :657-659 as written
    }
    out->n_escaped = h->u[5];
    out->max_cell_count = h->u[6];
    out->n_lost = h->u[7];
    out->n_overflow = h->u[8];
    out->n_boundary = (unsigned int)c->boundary.n;
    out->steps = c->steps;
}

int sphb_get_stats(sphb_ctx *c, sphb_stats *out)
{
    SPHB_ENTER(c);
    if (!out) return SPHB_E_ARG;
    memset(out, 0, sizeof *out);
    // one 128-byte block: 4 doubles | 16 words; read back through pinned memory in one copy
    if (!c->h_pinned) { SPHB_CUDA(cudaMallocHost(&c->h_pinned, 256)); c->pinned_bytes = 256; }
    StatsBlock *h = static_cast<StatsBlock *>(c->h_pinned);
    memset(h, 0, sizeof *h);
    SPHB_CUDA(cudaMemsetAsync(c->d_stats, 0, sizeof(StatsBlock), c->stream));
    if (c->fluid.n > 0)
        c->launches += launch_stats(c->stream, c->k, c->fluid, c->d_stats, reinterpret_cast<float *>(c->d_stats + 4),
                                    c->d_counters, c->mg.on ? c->mg.d_flags : nullptr);
    SPHB_CUDA(cudaMemcpyAsync(h, c->d_stats, sizeof(StatsBlock), cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    decode_stats(c, h, out);
    if (c->mg.on && (out->n_overflow & (1u << 30))) { c->mg.comm_failed = true; return mg_health(c); }
    return SPHB_OK;
}

int sphb_reorder_marks(sphb_ctx *c, unsigned long long *builds_with_marks)
{
    SPHB_ENTER(c);
    if (!builds_with_marks) { set_error("null argument"); return SPHB_E_ARG; }
    *builds_with_marks = c->fluid.touch_builds;
    return SPHB_OK;
}

int sphb_flush_l2(sphb_ctx *c)
{
    SPHB_ENTER(c);
    const size_t bytes = (size_t)256 << 20;
    if (!c->d_l2_scratch) SPHB_CUDA(cudaMalloc(&c->d_l2_scratch, bytes));
    c->l2_flush_value ^= 0x5a;
    SPHB_CUDA(cudaMemsetAsync(c->d_l2_scratch, c->l2_flush_value, bytes, c->stream));
    return SPHB_OK;
}

// ---- state files (the reference keeps its state in malloc'd arrays only) -----
// Layout: 64-byte header | sphb_params | fluid[n] (struct particle) | du_dt[n] | dv_dt[n] | boundary[nb].
// Everything a run needs to continue bit-identically: positions, velocities and the accelerations the
// next kick uses (:616); rho/p/psi are recomputed by the first step / sphb_init_boundary.
struct StateHeader {
    char magic[8];                  // "SPHB200\0"
    uint32_t version, params_bytes;
    uint32_t n_fluid, n_boundary;
    unsigned long long steps;
    uint32_t has_accel, reserved[7];
};
static_assert(sizeof(StateHeader) == 64, "header layout");

int sphb_save_state(sphb_ctx *c, const char *path)
{
    SPHB_ENTER(c);
    if (!path) return SPHB_E_ARG;
    if (c->mg.on) { set_error("save each slab through sphb_mg_download"); return SPHB_E_STATE; }
    const int n = c->fluid.n, nb = c->boundary.n;
    sphb_particle *f = static_cast<sphb_particle *>(malloc(sizeof(sphb_particle) * (size_t)(n > 0 ? n : 1)));
    sphb_particle *b = static_cast<sphb_particle *>(malloc(sizeof(sphb_particle) * (size_t)(nb > 0 ? nb : 1)));
    float *du = static_cast<float *>(malloc(sizeof(float) * (size_t)(n > 0 ? n : 1)));
    float *dv = static_cast<float *>(malloc(sizeof(float) * (size_t)(n > 0 ? n : 1)));
    int rc = (f && b && du && dv) ? SPHB_OK : SPHB_E_NOMEM;
    if (!rc) rc = sphb_download(c, f, du, dv);
    if (!rc) rc = sphb_download_boundary(c, b);
    if (!rc) {
        StateHeader h;
        memset(&h, 0, sizeof h);
        memcpy(h.magic, "SPHB200", 8);
        h.version = 1; h.params_bytes = (uint32_t)sizeof(sphb_params);
        h.n_fluid = (uint32_t)n; h.n_boundary = (uint32_t)nb; h.steps = c->steps; h.has_accel = c->accel_ready ? 1u : 0u;
        FILE *fp = fopen(path, "wb");
        bool ok = fp != nullptr;
        ok = ok && fwrite(&h, sizeof h, 1, fp) == 1 && fwrite(&c->prm, sizeof c->prm, 1, fp) == 1;
        ok = ok && (n == 0 || (fwrite(f, sizeof *f, n, fp) == (size_t)n && fwrite(du, 4, n, fp) == (size_t)n && fwrite(dv, 4, n, fp) == (size_t)n));
        ok = ok && (nb == 0 || fwrite(b, sizeof *b, nb, fp) == (size_t)nb);
        if (fp) ok = (fclose(fp) == 0) && ok;
        if (!ok) { set_error("cannot write state file %s", path); rc = SPHB_E_ARG; }
    }
    free(f); free(b); free(du); free(dv);
    return rc;
}

int sphb_load_state(const char *path, int device, sphb_ctx **out)
{
    if (!path || !out) { set_error("null argument"); return SPHB_E_ARG; }
    *out = nullptr;
    FILE *fp = fopen(path, "rb");
    if (!fp) { set_error("cannot open state file %s", path); return SPHB_E_ARG; }
    StateHeader h;
    sphb_params prm;
    sphb_particle *f = nullptr, *b = nullptr;
    float *du = nullptr, *dv = nullptr;
    sphb_ctx *c = nullptr;
    int rc = SPHB_OK;
    if (fread(&h, sizeof h, 1, fp) != 1 || memcmp(h.magic, "SPHB200", 8) != 0 || h.version != 1 ||
        h.params_bytes != sizeof(sphb_params) || fread(&prm, sizeof prm, 1, fp) != 1) {
        set_error("%s is not a version-1 state file", path);
        rc = SPHB_E_ARG;
    }
    if (!rc) {
        const size_t n = h.n_fluid, nb = h.n_boundary;
        f = static_cast<sphb_particle *>(malloc(sizeof(sphb_particle) * (n ? n : 1)));
        b = static_cast<sphb_particle *>(malloc(sizeof(sphb_particle) * (nb ? nb : 1)));
        du = static_cast<float *>(malloc(4 * (n ? n : 1)));
        dv = static_cast<float *>(malloc(4 * (n ? n : 1)));
        if (!f || !b || !du || !dv) rc = SPHB_E_NOMEM;
        else if ((n && (fread(f, sizeof *f, n, fp) != n || fread(du, 4, n, fp) != n || fread(dv, 4, n, fp) != n)) ||
                 (nb && fread(b, sizeof *b, nb, fp) != nb)) {
            set_error("%s is truncated", path);
            rc = SPHB_E_ARG;
        }
    }
    fclose(fp);
    if (!rc) {
        if (device >= 0) prm.device = device;
        rc = sphb_create(&prm, &c);
    }
    if (!rc) rc = sphb_upload(c, f, (int)h.n_fluid, b, (int)h.n_boundary);
    // psi is recomputed from the wall particles' rho (:242-261), which the file carries unchanged
    if (!rc) rc = sphb_init_boundary(c);
    if (!rc && h.has_accel && h.n_fluid) rc = sphb_upload_accel(c, du, dv);
    if (!rc) { c->steps = h.steps; *out = c; }
    else if (c) sphb_destroy(c);
    free(f); free(b); free(du); free(dv);
    return rc;
}

void *sphb_stream(sphb_ctx *c) { return c ? (void *)c->stream : nullptr; }
unsigned long long sphb_launch_count(sphb_ctx *c) { return c ? c->launches : 0ULL; }

}  // extern "C"

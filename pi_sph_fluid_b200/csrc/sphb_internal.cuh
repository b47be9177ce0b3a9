// sphb_internal.cuh — device-state layout and kernel entry points shared by the .cu files.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sph_b200.h"
#include "sph_math.cuh"

namespace sphb {

// ---- tunables ----------------------------------------------------------------------------
constexpr int kStreamThreads = 256;   // advect/bin, reorder, gather/scatter kernels
constexpr int kPairThreads = 128;     // density / force CTAs: one thread per particle
constexpr int kListCap = 48;          // per-thread accepted list entries (u16 tile offsets) before a flush
constexpr int kDensityListCap = 40;   // same for the density pass's list of f32 squared distances
constexpr int kTileCap = 576;         // staged neighbourhood entries per CTA (x 8 B must stay < 64 KiB)
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// A sorted particle set: SoA in HBM, permanently ordered by cell (row-major, the
// reference's ij_cell = i_cell*m_cells + j_cell, :113).
struct ParticleSet {
    int n = 0;
    int cap = 0;                 // allocated slots
    float2 *pos[2] = {nullptr, nullptr};
    float2 *vel[2] = {nullptr, nullptr};
    uint32_t *id[2] = {nullptr, nullptr};     // original index of the particle in each slot
    float *mass[2] = {nullptr, nullptr};      // fluid: only when masses differ; boundary: psi
    int pc = 0, vc = 0, ic = 0, mc = 0, xc = 0; // which buffer is current
    float2 *acc = nullptr;                    // du_dt, dv_dt (fluid)
    float2 *rho_prr = nullptr;                // rho, p/rho^2 (fluid)
    float *p = nullptr;                       // pressure (fluid)
    float *aux[2] = {nullptr, nullptr};       // boundary: the rho field handed in (:526-538)
    float uniform_mass_value = 0.0f;
    uint32_t *key = nullptr, *rank = nullptr; // counting-sort scratch
    uint32_t *cellkey = nullptr;              // (row << 16 | col) of the particle in each sorted slot
    uint32_t *ids_tmp = nullptr;              // deterministic-rank scratch
    uint32_t *cell_count = nullptr;           // ncells, zero between builds
    uint32_t *cell_start = nullptr;           // ncells + 1
    bool sorted = false;
    bool uniform_mass = true;
};

struct ScanState {
    unsigned long long *tile_state = nullptr;   // (epoch,flag | value) per tile
    unsigned long long *tile_counter = nullptr; // dynamic tile ids
    int n_tiles = 0;
    unsigned int epoch = 0;
    unsigned long long launches = 0;
};

struct DeviceCounters {        // lives in HBM, read back by sphb_get_stats
    unsigned int n_escaped;
    unsigned int max_cell_count;
    unsigned int list_flushes;     // times a thread's accepted list filled up
    unsigned int tiles_unstaged;   // CTAs that fell back to global reads
    unsigned long long pair_candidates;
    unsigned long long pair_accepted;
};

}  // namespace sphb

struct sphb_ctx {
    sphb_params prm;
    sphb::Consts k;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    sphb::ParticleSet fluid, boundary;
    sphb::ScanState scan;
    sphb::DeviceCounters *d_counters = nullptr;
    float2 *d_gravity = nullptr;          // device copy of per-step gravity (trace mode)
    void *d_stage = nullptr;              // AoS staging for upload/download
    size_t stage_bytes = 0;
    void *h_pinned = nullptr;             // pinned bounce buffer
    size_t pinned_bytes = 0;
    float2 *d_pixels = nullptr;           // pixel-centre pseudo-particles (:570-577)
    unsigned char *d_frame = nullptr;     // 1 KiB SSD1306 frame
    double *d_stats = nullptr;            // 4 doubles + 4 words
    void *d_l2_scratch = nullptr;         // sphb_flush_l2
    int l2_flush_value = 0;
    bool boundary_ready = false;
    bool accel_ready = false;
    unsigned long long steps = 0;
    unsigned long long launches = 0;
    // profiling
    int profile_mode = 0;
    cudaEvent_t ev[2 * 64];
    int ev_kind[64];
    int ev_used = 0;
    double prof_ms[SPHB_K_COUNT] = {0};
    unsigned long long prof_launches[SPHB_K_COUNT] = {0};
};

namespace sphb {

// error plumbing (sphb_api.cu)
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define SPHB_CUDA(call)                                                         \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return sphb::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// ---- kernel launchers (kernels_build.cu) ---------------------------------------------------
// All launch on `st`; return the number of kernels launched.
int launch_advect_bin(cudaStream_t st, const Consts &k, ParticleSet &ps, bool advect, DeviceCounters *ctr);
int launch_scan(cudaStream_t st, const Consts &k, ParticleSet &ps, ScanState &sc, DeviceCounters *ctr);
int launch_reorder(cudaStream_t st, const Consts &k, ParticleSet &ps, bool deterministic);
int launch_aos_to_soa(cudaStream_t st, const sphb_particle *aos, ParticleSet &ps, bool is_boundary);
int launch_soa_to_aos(cudaStream_t st, const ParticleSet &ps, sphb_particle *aos, float *du, float *dv, bool is_boundary);
int launch_set_accel(cudaStream_t st, ParticleSet &ps, const float *du, const float *dv);
int launch_cell_ids(cudaStream_t st, const Consts &k, const ParticleSet &ps, int *cell_out);

// ---- kernel launchers (kernels_pair.cu) ----------------------------------------------------
int launch_pseudomass(cudaStream_t st, const Consts &k, ParticleSet &boundary);
int launch_density(cudaStream_t st, const Consts &k, ParticleSet &fluid, const ParticleSet &boundary,
                   DeviceCounters *ctr, bool count_pairs, bool allow_stage = true);
int launch_force(cudaStream_t st, const Consts &k, ParticleSet &fluid, const ParticleSet &boundary,
                 float gx, float gy, const float2 *g_dev, bool kick2, DeviceCounters *ctr,
                 bool allow_stage = true);
int launch_neighbor_lists(cudaStream_t st, const Consts &k, const ParticleSet &a, const ParticleSet &b,
                          bool same, int cap, int *counts, int *lists, unsigned int *overflow);

// ---- kernel launchers (kernels_aux.cu) -----------------------------------------------------
int launch_render(cudaStream_t st, const Consts &k, const ParticleSet &fluid, const float2 *pixels,
                  float W_px, unsigned char *frame);
int launch_stats(cudaStream_t st, const Consts &k, const ParticleSet &fluid, double *out_d /*4*/,
                 float *out_u /*4 x 32-bit*/);
int launch_refresh(cudaStream_t st, const sphb_particle *aos, ParticleSet &ps, unsigned int *moved);
int launch_tait_aos(cudaStream_t st, const Consts &k, int n, sphb_particle *aos);

// ---- host-side helpers (sphb_api.cu) ---------------------------------------------------------
int alloc_set(ParticleSet &ps, int n, int ncells, bool is_boundary, bool need_mass);
int ensure_stage(sphb_ctx *c, size_t bytes);
int build_grid(sphb_ctx *c, ParticleSet &ps, bool advect);

}  // namespace sphb

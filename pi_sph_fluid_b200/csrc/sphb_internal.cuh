// sphb_internal.cuh — device-state layout and kernel entry points shared by the .cu files.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sph_b200.h"
#include "sph_math.cuh"

namespace sphb {

// ---- tunables ----------------------------------------------------------------------------
#ifndef SPHB_STREAM_THREADS
#define SPHB_STREAM_THREADS 256
#endif
constexpr int kStreamThreads = SPHB_STREAM_THREADS;   // advect/bin, reorder, gather/scatter kernels
// (the SPHB_* macros exist so that scripts/tune_pair.sh can build variants; defaults are the tuned values)
#ifndef SPHB_PT
#define SPHB_PT 128
#endif
#ifndef SPHB_LIST_CAP
#define SPHB_LIST_CAP 44
#endif
#ifndef SPHB_TILE_CAP
#define SPHB_TILE_CAP 624
#endif
#ifndef SPHB_WIN_CAP
#define SPHB_WIN_CAP 96
#endif
#ifndef SPHB_PERSISTENT
#define SPHB_PERSISTENT 0
#endif
#ifndef SPHB_MINB_D
#define SPHB_MINB_D 10
#endif
#ifndef SPHB_MINB_F
#define SPHB_MINB_F 8
#endif
#ifndef SPHB_MINB_FS
#define SPHB_MINB_FS 8      // the reference-arithmetic force pass (k_force MODE 1 / 2)
#endif
constexpr int kPairThreads = SPHB_PT;        // density / force CTAs: one thread per particle
constexpr int kStatsSlots = 64;              // copies of the step-statistics block the force pass spreads its atomics over
constexpr int kChunkRecWords = 16;           // the record k_density leaves per chunk for k_force (64 bytes)
constexpr int kListCap = SPHB_LIST_CAP;      // per-thread accepted list entries (u16 tile offsets) before a flush
constexpr int kTileCap = SPHB_TILE_CAP;      // staged neighbourhood entries per CTA (x 8 B must stay < 64 KiB)
constexpr int kWinCap = SPHB_WIN_CAP;        // staged cell_start words per neighbour row of a chunk (multiple of 4)
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

constexpr uint32_t kTrashKey = 0xffffffffu;   // build key of a slot that is dropped by the sort

// ---- programmatic dependent launch ---------------------------------------------------------------
// The kernels of a step form one chain on one stream.  Each of them lets its successor's CTAs be
// scheduled as soon as all of its own CTAs have started (pdl_trigger, first instruction) and waits for
// its predecessor's completion and memory (pdl_wait) before it touches global memory, so launch
// latency, CTA ramp-up and shared-memory set-up of kernel n+1 overlap the tail of kernel n.  Waiting
// CTAs only ever occupy resources the predecessor no longer needs (all its CTAs are resident by then).
// SPHB_NO_PDL=1 in the environment launches the same kernels plainly (the wait is then a no-op).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();      // sphb_api.cu
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(cudaStream_t st, dim3 grid, dim3 block, void (*kern)(KArgs...), Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at = {};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = pdl_enabled() ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<Args &&>(args)...);
}

// Particle count of a launch: `n` sizes the grid (an upper bound when `dev` is set); the kernels
// use *dev when it is non-null (multi-GPU slabs: counts change on the device every step).
struct Count {
    int n;
    const int *dev;
};
__device__ __forceinline__ int count_of(const Count &c)
{
    if (!c.dev) return c.n;
    const int v = *c.dev;
    return v < c.n ? v : c.n;
}

// One halo/migration message (sphb_mg.cu): [count, pad x3 | pos[cap] | vel[cap] | id[cap]]
struct HaloBuf {
    unsigned char *base = nullptr;
    int cap = 0;
    __host__ __device__ uint32_t *hdr() const { return reinterpret_cast<uint32_t *>(base); }
    __host__ __device__ float2 *pos() const { return reinterpret_cast<float2 *>(base + 16); }
    __host__ __device__ float2 *vel() const { return pos() + cap; }
    __host__ __device__ uint32_t *id() const { return reinterpret_cast<uint32_t *>(vel() + cap); }
    __host__ __device__ size_t bytes() const { return 16 + (size_t)cap * 20; }
};

// What the slab variant of the build kernels needs (all device pointers; side 0 = left, 1 = right)
struct SlabIO {
    uint32_t *send_cnt[2] = {nullptr, nullptr};   // running count of entries appended to send[side]
    HaloBuf send[2];                              // local send buffer (NCCL) or the peer's recv buffer (direct stores)
    HaloBuf recv[2];
    int has[2] = {0, 0};                          // neighbour present on that side
    unsigned int *lost = nullptr;                 // particles that left the window without a taker
    unsigned int *overflow = nullptr;             // message or slot capacity exceeded
    int capacity = 0;                             // allocated particle slots
    // peer stores across processes: the binning kernel waits until hdr()[1] of each receive buffer
    // carries this epoch (0: the message is complete by stream order, nothing to wait for)
    uint32_t wait_epoch = 0;
};

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
    // Deterministic reorder without the id pass (kernels_build.cu): the scan writes into the other of two
    // cell_start buffers, so the reorder still sees where every cell began in the PREVIOUS sorted order, and
    // the binning kernels mark (with the build's epoch) every cell a particle left or entered.  A cell that
    // is not marked holds exactly the particles it held before, in the same ascending-id order.
    uint32_t *cell_start_prev = nullptr;      // ncells + 1: cell_start of the previous build
    unsigned char *cell_touch = nullptr;      // ncells: epoch of the last build that changed the cell's population
    unsigned int touch_epoch = 0;             // 1..255 (0 = never; the array is cleared when the epoch wraps)
    bool touch_ok = false;                    // this build's input is the previous sorted order: marks are complete
    unsigned long long touch_builds = 0;      // builds that used the marks (sphb_reorder_marks)
    int touch_min_slots = 1 << 20;            // sets smaller than this keep the id pass for every cell (touch_min_slots())
    bool sorted = false;
    bool counters_dirty = false;              // a grid build ran since the density pass last published the build counters
    bool uniform_mass = true;
    // accepted-neighbour lists handed from the density pass to the force pass of the same step
    unsigned short *nbr_list = nullptr;       // [CTA][entry < kListCap][thread] tile byte offsets
    unsigned short *nbr_count = nullptr;      // per sorted slot; 0xffff = search again
    unsigned int *chunk_rec = nullptr;        // per chunk, kChunkRecWords words: [S0 S1 S2 | n0 n1 n2 | w0 w1 w2 |
                                              //   wn0 wn1 wn2 | flags (1 whole-chunk plan valid, 2 wall near,
                                              //   4 cell_start windows staged) | rows of its list block in use]
    bool lists_valid = false;
    // chunk tickets of the density [0] and force [1] kernels (kernels_pair.cu: ChunkQueue)
    unsigned long long *chunk_queue = nullptr;
    unsigned int queue_epoch = 0;
    // multi-GPU slabs: counts live on the device, `n` is only the launch bound
    int *d_n_cur = nullptr;                   // valid sorted slots (owned + ghost)
    int *d_n_in = nullptr;                    // slots feeding the current build (previous + received)
    bool windowed = false;                    // the grid is a slab window: the sort may drop particles
    Count cur() const { return Count{n, d_n_cur}; }
    Count in() const { return Count{n, d_n_in}; }
};

struct ScanState {
    unsigned long long *tile_state = nullptr;   // (epoch,flag | value) per tile
    unsigned long long *tile_counter = nullptr; // dynamic tile ids
    int n_tiles = 0;
    unsigned int epoch = 0;
    unsigned long long launches = 0;
};

struct DeviceCounters {        // lives in HBM, read back by sphb_get_stats ([0] fluid builds, [1] boundary builds)
    unsigned int n_escaped;        // particles binned by clamping at the LAST build    } published by the density pass
    unsigned int max_cell_count;   // largest cell population at the last build          } that follows the build
    unsigned int escaped_acc;      // running count the binning kernel adds to
    unsigned int escaped_prev;     //   ... and its value at the previous publication
    unsigned int max_cell_acc;     // running maximum of the scan since the previous publication
    unsigned int list_flushes;     // times a thread's accepted list filled up
    unsigned int tiles_unstaged;   // CTAs that fell back to global reads
    unsigned long long pair_candidates;
    unsigned long long pair_accepted;
};

// Per-step statistics gathered by the force pass itself (sphb_step_stats): the 128-byte block of
// sphb_get_stats (4 doubles | 16 words) is accumulated in HBM by one set of atomics per CTA, and the
// last CTA to finish stores it, followed by a sequence word, into mapped pinned host memory — the
// host polls that word, so reading a step's statistics costs no further API call.
struct StepStats {
    unsigned long long *block = nullptr;   // device: kStatsSlots x 16 x 8 bytes, all zero between steps (the last
                                           //   CTA folds and clears them); chunk c adds into slot c % kStatsSlots
    const DeviceCounters *ctr = nullptr;   // snapshotted into slot 0 by the density pass of the step
    const unsigned int *flags = nullptr;   // slabs: [0] lost, [1] overflow (likewise)
    const uint32_t *id = nullptr;          // original index per sorted slot
    uint32_t last_id = 0xffffffffu;        // :657-659 as written reports the LAST particle's error
    unsigned int *done = nullptr;          // CTAs that have finished (left at zero by the last one); nullptr: the
                                           //   force pass does not deliver, a later kernel does (stats_fold_deliver)
    unsigned long long *host = nullptr;    // mapped pinned: [0..15] the block, [16] sequence word
    unsigned long long seq = 0;
};

// Folds the kStatsSlots partial blocks of a step's statistics into one, clears the slots for their next
// use, and hands the block and then the sequence word to the host through mapped pinned memory.  Run by
// ONE whole CTA (>= kStatsSlots threads) once every chunk of the force pass has added its sums: the last
// CTA of the force pass itself (blocking sphb_step_stats), CTA 0 of the next step's advect+bin kernel
// (sphb_step_stats_begin: the delivery then hides inside that kernel instead of delaying it), or
// k_stats_deliver when no step follows.  Slot 0 also carries [5 hi] the last particle's rho (:657-659)
// and, in words 9 and 10, the build / slab counters the density pass of the step snapshotted.
#if defined(__CUDACC__)
__device__ __forceinline__ void stats_fold_deliver(const StepStats &ss)
{
    __shared__ double s_fd[kStatsSlots / 32][4];
    __shared__ unsigned int s_fu[kStatsSlots / 32][4];
    __shared__ unsigned long long s_out[16];
    __shared__ unsigned long long s_extra[3];
    const int tid = threadIdx.x;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
    unsigned int u0 = 0u, u1 = 0u, u2 = 0u, u4 = 0u;
    if (tid < kStatsSlots) {
        unsigned long long *sl = ss.block + (size_t)tid * 16;
        d0 = __longlong_as_double((long long)__ldcg(sl + 0)); d1 = __longlong_as_double((long long)__ldcg(sl + 1));
        d2 = __longlong_as_double((long long)__ldcg(sl + 2)); d3 = __longlong_as_double((long long)__ldcg(sl + 3));
        const unsigned long long w4 = __ldcg(sl + 4), w5 = __ldcg(sl + 5), w6 = __ldcg(sl + 6);
        u0 = (unsigned int)w4; u1 = (unsigned int)(w4 >> 32);
        u2 = (unsigned int)w5;
        u4 = (unsigned int)w6;
        if (tid == 0) { s_extra[0] = w5 >> 32; s_extra[1] = __ldcg(sl + 9); s_extra[2] = __ldcg(sl + 10); }
#pragma unroll
        for (int i = 0; i < 16; i++) sl[i] = 0ULL;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            d0 += __shfl_xor_sync(0xffffffffu, d0, d); d1 += __shfl_xor_sync(0xffffffffu, d1, d);
            d2 += __shfl_xor_sync(0xffffffffu, d2, d); d3 += __shfl_xor_sync(0xffffffffu, d3, d);
        }
        u0 = __reduce_max_sync(0xffffffffu, u0); u1 = __reduce_max_sync(0xffffffffu, u1);
        u2 = __reduce_max_sync(0xffffffffu, u2); u4 = __reduce_add_sync(0xffffffffu, u4);
        if ((tid & 31) == 0) {
            const int w = tid >> 5;
            s_fd[w][0] = d0; s_fd[w][1] = d1; s_fd[w][2] = d2; s_fd[w][3] = d3;
            s_fu[w][0] = u0; s_fu[w][1] = u1; s_fu[w][2] = u2; s_fu[w][3] = u4;
        }
    }
    __syncthreads();
    if (tid == 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        unsigned int m0 = 0u, m1 = 0u, m2 = 0u, c4 = 0u;
#pragma unroll
        for (int w = 0; w < kStatsSlots / 32; w++) {
            a0 += s_fd[w][0]; a1 += s_fd[w][1]; a2 += s_fd[w][2]; a3 += s_fd[w][3];
            m0 = s_fu[w][0] > m0 ? s_fu[w][0] : m0; m1 = s_fu[w][1] > m1 ? s_fu[w][1] : m1;
            m2 = s_fu[w][2] > m2 ? s_fu[w][2] : m2; c4 += s_fu[w][3];
        }
        // the 128-byte block of sphb_get_stats: 4 doubles | u32 [0] max speed [1] key(max rho) [2] ~key(min rho)
        // [3] last particle's rho [4] owned [5] escaped [6] max cell population [7] lost [8] overflow
        s_out[0] = (unsigned long long)__double_as_longlong(a0); s_out[1] = (unsigned long long)__double_as_longlong(a1);
        s_out[2] = (unsigned long long)__double_as_longlong(a2); s_out[3] = (unsigned long long)__double_as_longlong(a3);
        s_out[4] = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
        s_out[5] = (unsigned long long)m2 | (s_extra[0] << 32);
        s_out[6] = (unsigned long long)c4 | ((s_extra[1] & 0xffffffffULL) << 32);
        s_out[7] = (s_extra[1] >> 32) | ((s_extra[2] & 0xffffffffULL) << 32);
        s_out[8] = s_extra[2] >> 32;
#pragma unroll
        for (int i = 9; i < 16; i++) s_out[i] = 0ULL;
    }
    __syncthreads();
    if (tid < 16) {
        ss.host[tid] = s_out[tid];
        __threadfence_system();
    }
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile unsigned long long *>(ss.host + 16) = ss.seq;
}
#endif

// Multi-GPU slab state of one rank (sphb_mg.cu)
struct MgState {
    bool on = false;
    int rank = 0, world = 1;
    int col_lo = 0, col_hi = 0;          // owned global columns [col_lo, col_hi)
    int transport = 0;                   // 0: not connected, 1: NCCL (one process per GPU), 2: in-process peers,
                                         // 3: peer stores across processes (CUDA IPC)
    void *nccl_comm = nullptr;           // ncclComm_t
    int halo_cap = 0;                    // entries per message
    int capacity = 0;                    // particle slots
    Consts k_global;                     // the unwindowed grid (boundary pseudo-mass pass)
    unsigned char *d_send[2] = {nullptr, nullptr};
    unsigned char *d_recv[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [side][step parity], inside d_recv_block
    unsigned char *d_recv_block = nullptr;   // one allocation (one IPC handle): 4 messages, recv_stride apart
    size_t recv_stride = 0;
    unsigned char *ipc_peer_block[2] = {nullptr, nullptr};    // the neighbours' receive blocks mapped here
    uint32_t *d_send_cnt = nullptr;      // 2 words (in-process transport; NCCL counts in the message header)
    unsigned int *d_flags = nullptr;     // [0] lost, [1] overflow
    int *d_counts = nullptr;             // [0] n_cur, [1] n_in (fluid), [2] boundary n
    sphb_ctx *peer[2] = {nullptr, nullptr};
    cudaEvent_t ev_sent = nullptr;       // phase A of the current step is complete on this rank's stream
    unsigned long long exchanges = 0;    // step parity for the double-buffered receive side
    unsigned long long halo_bytes = 0;   // bytes sent so far
    int n_uploaded = 0;                  // particles of the last sphb_mg_upload (slot order until the first sort)
    bool comm_failed = false;            // a neighbour's message timed out (mg_health): every later call fails
    int bnd_n_global = 0;                // wall particles of the whole tank: they stay, psi computed, in the buffers the
                                         //   windowed sort read from, so a re-cut can window them again (sphb_mg_rebalance)
};

}  // namespace sphb

struct sphb_ctx {
    sphb::MgState mg;
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
    double *d_stats = nullptr;            // 4 doubles + 16 words (sphb_get_stats)
    unsigned long long *d_step_slots = nullptr;   // StepStats::block
    unsigned int *d_stats_done = nullptr; // StepStats::done
    unsigned long long *h_stats = nullptr;   // StepStats::host (mapped pinned): two slots of 32 words, [0..15] block, [16] sequence
    unsigned long long stats_seq = 0;        // step statistics requested so far (the sequence word delivered with them)
    unsigned long long stats_collected = 0;  // ... and read by the host, in order
    unsigned long long stats_steps[2] = {0, 0};   // sphb_stats::steps of the outstanding requests, by slot
    sphb::StepStats stats_pending;            // sphb_step_stats_begin: statistics summed on the device but not yet
    bool stats_has_pending = false;           //   folded and delivered (the next advect+bin kernel or the collect does it)
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
int launch_advect_bin(cudaStream_t st, const Consts &k, ParticleSet &ps, bool advect, DeviceCounters *ctr,
                      const SlabIO *slab = nullptr, const StepStats *deliver = nullptr);
int launch_stats_deliver(cudaStream_t st, const StepStats &ss);
int launch_bin_recv(cudaStream_t st, const Consts &k, ParticleSet &ps, const SlabIO &slab, DeviceCounters *ctr);
int launch_scan(cudaStream_t st, const Consts &k, ParticleSet &ps, ScanState &sc, DeviceCounters *ctr);
int launch_reorder(cudaStream_t st, const Consts &k, ParticleSet &ps, bool deterministic);
int touch_min_slots();
#ifndef SPHB_TOUCH
#define SPHB_TOUCH 1      // 0: every particle goes through the id pass of the deterministic reorder (round-1 behaviour)
#endif
int launch_aos_to_soa(cudaStream_t st, const sphb_particle *aos, ParticleSet &ps, bool is_boundary, int n = -1,
                      const uint32_t *ids = nullptr, uint32_t id_base = 0, uint32_t m0_bits = 0,
                      unsigned int *mass_differs = nullptr);
int launch_pack_owned(cudaStream_t st, const Consts &k, const ParticleSet &ps, int cap, sphb_particle *aos,
                      uint32_t *ids_out, float *du, float *dv, unsigned int *n_out);
int launch_soa_to_aos(cudaStream_t st, const ParticleSet &ps, sphb_particle *aos, float *du, float *dv, bool is_boundary);
int launch_set_accel(cudaStream_t st, ParticleSet &ps, const float *du, const float *dv, int n_slots = -1);
int launch_cell_ids(cudaStream_t st, const Consts &k, const ParticleSet &ps, int *cell_out);
// re-cut of the slabs (sphb_mg_rebalance): per-global-column counts of the owned particles; the owned particles
// as 32-byte records grouped by destination rank; records back into an unsorted set
struct MoveRec { float2 pos, vel, acc; uint32_t id, pad; };
int launch_column_hist(cudaStream_t st, const Consts &k, const ParticleSet &ps, unsigned long long *hist);
int launch_pack_by_dest(cudaStream_t st, const Consts &k, const ParticleSet &ps, const int *cuts_dev, int world,
                        const unsigned long long *seg_off_dev, unsigned long long *cursor_dev, MoveRec *out);
int launch_unpack_moved(cudaStream_t st, ParticleSet &ps, const MoveRec *in, int n);

// ---- kernel launchers (kernels_pair.cu) ----------------------------------------------------
int launch_pseudomass(cudaStream_t st, const Consts &k, ParticleSet &boundary);
int launch_density(cudaStream_t st, const Consts &k, ParticleSet &fluid, const ParticleSet &boundary,
                   DeviceCounters *ctr, bool count_pairs, bool allow_stage = true,
                   unsigned long long *stats_zero = nullptr, const unsigned int *stats_flags = nullptr);
int launch_force(cudaStream_t st, const Consts &k, ParticleSet &fluid, const ParticleSet &boundary,
                 float gx, float gy, const float2 *g_dev, bool kick2, DeviceCounters *ctr,
                 bool allow_stage = true, const StepStats *stats = nullptr, bool fast_force = false);
int launch_decode_handover(cudaStream_t st, const Consts &k, const ParticleSet &f, int cap, int *counts, int *lists,
                           unsigned int *n_fast_chunks);
int launch_probe_force_pair(cudaStream_t st, const Consts &k, int n, const float *in, int variant, float *out);
int launch_neighbor_lists(cudaStream_t st, const Consts &k, const ParticleSet &a, const ParticleSet &b,
                          bool same, int cap, int *counts, int *lists, unsigned int *overflow);

// ---- kernel launchers (kernels_aux.cu) -----------------------------------------------------
int launch_render(cudaStream_t st, const Consts &k, const ParticleSet &fluid, const float2 *pixels,
                  float W_px, unsigned char *frame);
int launch_pixel_counts(cudaStream_t st, const Consts &k, const ParticleSet &f, float width, float height, unsigned int *counts);
int launch_stats(cudaStream_t st, const Consts &k, const ParticleSet &fluid, double *out_d /*4*/,
                 float *out_u /*16 x 32-bit*/, const DeviceCounters *ctr, const unsigned int *flags);
int launch_refresh(cudaStream_t st, const sphb_particle *aos, ParticleSet &ps, unsigned int *moved);
int launch_tait_aos(cudaStream_t st, const Consts &k, int n, sphb_particle *aos);

// ---- host-side helpers (sphb_api.cu) ---------------------------------------------------------
int alloc_set(ParticleSet &ps, int n, int ncells, bool is_boundary, bool need_mass);
int ensure_stage(sphb_ctx *c, size_t bytes);
int build_grid(sphb_ctx *c, ParticleSet &ps, bool advect, const Consts *kk = nullptr, const StepStats *deliver = nullptr);
int step_phase_a(sphb_ctx *c, bool advect, const StepStats *deliver = nullptr);
int step_phase_b(sphb_ctx *c, float gx, float gy, bool kick2, const StepStats *ss = nullptr);
int free_set_public(ParticleSet &ps);

// sphb_mg.cu
SlabIO mg_slab_io(const sphb_ctx *c);
int mg_exchange_nccl(sphb_ctx *c);
int mg_exchange(sphb_ctx *c);            // the transport's part of a step between phase A and phase B
int launch_halo_signal(cudaStream_t st, const SlabIO &io, uint32_t epoch);
int mg_init_boundary(sphb_ctx *c);
int mg_health(sphb_ctx *c);              // SPHB_E_COMM once a halo wait has timed out; call after a stream synchronise
void mg_free(sphb_ctx *c);

}  // namespace sphb

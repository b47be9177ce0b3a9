// kernels_build.cu — the counting-sort cell build that replaces the reference's serial
// linked-list rebuild (update_neighbors_context, pi_sph_fluid.c:104-124), fused with the
// leapfrog kick + drift loops that precede it in the step (:615-624), plus the AoS<->SoA
// converters at the host boundary.
//
//   k_advect_bin   u,v += 0.5*DT*a (double) ; x,y += DT*u ; cell = (int)((y-ymin)/cell)*m + ...
//                  ; rank = atomicAdd(count[cell])                       HBM-bound, 44 B/particle
//   k_scan         exclusive prefix sum of count[] -> start[] (single pass, decoupled
//                  look-back, warp-shuffle scans), re-zeroing count[]    HBM-bound, 8 B/cell
//   k_scatter_ids  (deterministic mode) ids into their cell's range in arrival order — only for
//                  cells whose population changed in this build (cell_touch)
//   k_reorder      dst = start[cell] + rank, rank = #ids in the cell below mine
//                  (deterministic: the reference's ascending-index list order, :110-123)
//                  or the atomic's arrival rank; moves pos/vel/id        HBM-bound, 40-56 B/particle
//                  A cell nobody left or entered keeps its particles in their previous (ascending-id)
//                  order: rank = slot - cell_start_prev[cell], no id lookups.
#include "sphb_internal.cuh"
#include <cstdlib>

namespace sphb {

// ------------------------------------------------------------------------ advect + bin

#ifndef SPHB_TOUCH_MIN_SLOTS
#define SPHB_TOUCH_MIN_SLOTS (1 << 20)
#endif
// smallest set (in slots) whose deterministic reorder uses the cell marks; SPHB_TOUCH_MIN_SLOTS in the
// environment overrides it when a set is allocated (the parity tests run small scenes with 0)
int touch_min_slots()
{
    const char *e = getenv("SPHB_TOUCH_MIN_SLOTS");
    return (e && *e) ? atoi(e) : SPHB_TOUCH_MIN_SLOTS;
}

// Appends (p, v, id) of the lanes with `want` to a halo message: one atomic per warp, the lanes
// take consecutive entries.  Every lane of the warp must call this.
__device__ __forceinline__ void slab_append(bool want, uint32_t *cnt, const HaloBuf &b, float2 p, float2 v,
                                            uint32_t id, unsigned int *overflow)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cnt, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) {
        const uint32_t i = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
        if (i < (uint32_t)b.cap) {
            b.pos()[i] = p;
            b.vel()[i] = v;
            b.id()[i] = id;
        } else {
            atomicAdd(overflow, 1u);
        }
    }
}

// SLAB (multi-GPU): slots in ghost columns are dropped (their owner sends them again); owned
// particles within two columns of a cut — on either side of it, after the drift — are appended to
// the message for that neighbour: it covers the neighbour's ghost columns and the particles that
// migrate to it.  Particles whose cell left this rank's window get the trash key.
template <bool ADVECT, bool SLAB>
__global__ void __launch_bounds__(kStreamThreads)
k_advect_bin(const Consts k, const Count cnt, float2 *__restrict__ pos, float2 *__restrict__ vel,
             const float2 *__restrict__ acc, const uint32_t *__restrict__ id, const uint32_t *__restrict__ cellkey,
             uint32_t *__restrict__ key, uint32_t *__restrict__ rank, uint32_t *__restrict__ cell_count,
             DeviceCounters *__restrict__ ctr, const SlabIO io, const StepStats deliver,
             unsigned char *__restrict__ touch, const unsigned char epoch)
{
    pdl_trigger();
    pdl_wait();
    // sphb_step_stats_begin: the previous step's statistics are summed but not yet with the host; CTA 0
    // folds and delivers them while the rest of the grid is already advecting
    if (deliver.block != nullptr && blockIdx.x == 0) stats_fold_deliver(deliver);
    const int n = count_of(cnt);
    const int s = blockIdx.x * kStreamThreads + threadIdx.x;
    bool live = s < n;
    if (SLAB && live && cellkey) live = owned_col(k, (int)(cellkey[s] & 0xffffu));
    if (SLAB && s < n && !live) key[s] = kTrashKey;
    float2 p = make_float2(0.0f, 0.0f), v = make_float2(0.0f, 0.0f);
    int row = 0, col = 0;
    bool clamped = false, outside = false;
    if (live) {
        p = pos[s];
        if (ADVECT || SLAB) v = vel[s];
        float2 a = make_float2(0.0f, 0.0f);
        if (ADVECT) a = acc[s];            // all three loads in flight before anything waits for one of them
        // the cell this slot was sorted into by the previous build: the same function of the same position
        int row_old = 0, col_old = 0;
        if (ADVECT && touch != nullptr) {
            bool c0, o0;
            cell_of_window(k, p.x, p.y, row_old, col_old, c0, o0);
        }
        if (ADVECT) {
            v.x = kick(k, v.x, a.x);          // :616
            v.y = kick(k, v.y, a.y);          // :617
            p.x = drift(k, p.x, v.x);         // :622
            p.y = drift(k, p.y, v.y);         // :623
            vel[s] = v;
            pos[s] = p;
        }
        cell_of_window(k, p.x, p.y, row, col, clamped, outside);          // :111-112
        if (clamped) atomicAdd(&ctr->escaped_acc, 1u);
        if (outside) {
            key[s] = kTrashKey;
        } else {
            const uint32_t c = (uint32_t)(row * k.cols + col);   // :113
            key[s] = c;
            rank[s] = atomicAdd(&cell_count[c], 1u);
        }
        // a particle that changed cell (or left the window) marks both cells: their populations differ
        // from the previous build's, so the reorder ranks their particles by id again
        if (ADVECT && touch != nullptr && (outside || row != row_old || col != col_old)) {
            touch[row_old * k.cols + col_old] = epoch;
            if (!outside) touch[row * k.cols + col] = epoch;
        }
    }
    if (SLAB) {
        const bool to_left = live && io.has[0] && col < k.own_lo + 2;
        const bool to_right = live && io.has[1] && col >= k.own_hi - 2;
        const uint32_t my = (to_left || to_right) ? id[s] : 0u;
        slab_append(to_left, io.send_cnt[0], io.send[0], p, v, my, io.overflow);
        slab_append(to_right, io.send_cnt[1], io.send[1], p, v, my, io.overflow);
        if (live && outside && !to_left && !to_right) atomicAdd(io.lost, 1u);
    }
}

int launch_advect_bin(cudaStream_t st, const Consts &k, ParticleSet &ps, bool advect, DeviceCounters *ctr,
                      const SlabIO *slab, const StepStats *deliver)
{
    const StepStats dl = deliver ? *deliver : StepStats{};
    if (ps.n == 0) return deliver ? launch_stats_deliver(st, dl) : 0;
    const int grid = (ps.n + kStreamThreads - 1) / kStreamThreads;
    const uint32_t *keys = ps.sorted ? ps.cellkey : nullptr;
    // a new build starts: its epoch for the cell marks, and whether the marks will be complete (the input
    // is the previous sorted order, advanced in place)
    // (below ~1M slots the build kernels are latency-bound and the marks' extra dependent load costs more
    // than the id traffic it saves: 65.8 vs 66.7 us/step at 262k particles, 9.88 vs 9.75 ms at 64M)
    ps.touch_ok = SPHB_TOUCH && ps.sorted && ps.cell_touch != nullptr && ps.n >= ps.touch_min_slots;
    if (++ps.touch_epoch > 255u) {
        ps.touch_epoch = 1u;
        if (ps.cell_touch) cudaMemsetAsync(ps.cell_touch, 0, (size_t)k.ncells, st);
    }
    unsigned char *touch = ps.touch_ok ? ps.cell_touch : nullptr;
#define SPHB_ADV(A, S, IO)                                                                                  \
    launch_pdl(st, grid, kStreamThreads, k_advect_bin<A, S>, k, ps.cur(), ps.pos[ps.pc], ps.vel[ps.vc], ps.acc, \
               ps.id[ps.ic], keys, ps.key, ps.rank, ps.cell_count, ctr, IO, dl, touch, (unsigned char)ps.touch_epoch)
    if (slab) { if (advect) SPHB_ADV(true, true, *slab); else SPHB_ADV(false, true, *slab); }
    else { if (advect) SPHB_ADV(true, false, SlabIO()); else SPHB_ADV(false, false, SlabIO()); }
#undef SPHB_ADV
    return 1;
}

// ---- peer-store transport across processes: message completion on the device -------------------
// The advect+bin kernel has stored this rank's message entries into the neighbours' receive buffers
// (peer memory over NVLink).  This one-warp kernel follows it on the stream: lane `side` publishes
// (epoch << 32 | count) into the neighbour's message header with a system-scope release store, so the
// entries are visible there before the word is.  The neighbour's k_bin_recv waits for the epoch.
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(32)
k_halo_signal(const SlabIO io, const uint32_t epoch)
{
    pdl_trigger();
    pdl_wait();
    const int side = threadIdx.x;
    if (side < 2 && io.has[side]) {
        const uint32_t cnt = *io.send_cnt[side];
        __threadfence_system();
        st_release_sys_u64(reinterpret_cast<unsigned long long *>(io.send[side].hdr()),
                           ((unsigned long long)epoch << 32) | cnt);
    }
}

int launch_halo_signal(cudaStream_t st, const SlabIO &io, uint32_t epoch)
{
    launch_pdl(st, 1, 32, k_halo_signal, io, epoch);
    return 1;
}

// how long k_bin_recv waits for a neighbour's message before it gives up (a dead neighbour must not
// hang this GPU): the slab is then flagged (overflow word, bit 30) and the step goes on without it
constexpr unsigned long long kHaloWaitNs = 20ULL * 1000ULL * 1000ULL * 1000ULL;
constexpr unsigned int kHaloTimeoutFlag = 1u << 30;

// Received halo / migrant entries become slots n_cur .. n_cur + count_left + count_right - 1 of the
// build input and are binned like the rest.
__global__ void __launch_bounds__(kStreamThreads)
k_bin_recv(const Consts k, const SlabIO io, const int *__restrict__ n_cur, int *__restrict__ n_in,
           float2 *__restrict__ pos, float2 *__restrict__ vel, uint32_t *__restrict__ id,
           uint32_t *__restrict__ key, uint32_t *__restrict__ rank, uint32_t *__restrict__ cell_count,
           DeviceCounters *__restrict__ ctr, unsigned char *__restrict__ touch, const unsigned char epoch)
{
    __shared__ uint32_t s_cnt[2];
    pdl_trigger();
    pdl_wait();
    const uint32_t cap = (uint32_t)io.recv[0].cap;
    if (io.wait_epoch) {
        // peer-store transport: the neighbour's signal kernel publishes (count, epoch) when its message
        // is complete; one thread per side polls this rank's own memory
        if (threadIdx.x < 2) {
            const int side = threadIdx.x;
            uint32_t c = 0u;
            if (io.has[side]) {
                const uint32_t *h = io.recv[side].hdr();
                const unsigned long long t0 = global_timer_ns();
                // a slab that has timed out once stays flagged and does not wait again (the run is lost
                // either way; it should end quickly rather than after one time-out per step)
                const bool given_up = (*reinterpret_cast<volatile unsigned int *>(io.overflow) & kHaloTimeoutFlag) != 0u;
                bool ok;
                while (!(ok = ld_acquire_sys_u32(h + 1) == io.wait_epoch)) {
                    if (given_up || global_timer_ns() - t0 > kHaloWaitNs) break;
                    __nanosleep(100);
                }
                if (ok) c = ld_acquire_sys_u32(h);
                else if (blockIdx.x == 0) atomicOr(io.overflow, kHaloTimeoutFlag);
            }
            s_cnt[side] = c;
        }
        __syncthreads();
    } else {
        if (threadIdx.x < 2) s_cnt[threadIdx.x] = io.has[threadIdx.x] ? io.recv[threadIdx.x].hdr()[0] : 0u;
        __syncthreads();
    }
    uint32_t cl = s_cnt[0];
    uint32_t cr = s_cnt[1];
    cl = cl < cap ? cl : cap;
    cr = cr < cap ? cr : cap;
    const int n0 = *n_cur;
    const uint32_t t = blockIdx.x * kStreamThreads + threadIdx.x;
    if (t == 0) {
        const long long tot = (long long)n0 + cl + cr;
        *n_in = tot < io.capacity ? (int)tot : io.capacity;
        // the messages of this step are out (stream order): restart the send counters
        if (io.send_cnt[0]) *io.send_cnt[0] = 0u;
        if (io.send_cnt[1]) *io.send_cnt[1] = 0u;
    }
    if (t >= cl + cr) return;
    const HaloBuf &b = t < cl ? io.recv[0] : io.recv[1];
    const uint32_t e = t < cl ? t : t - cl;
    const long long slot = (long long)n0 + t;
    if (slot >= io.capacity) { atomicAdd(io.overflow, 1u); return; }
    // through L2 (the peer-store transport's entries arrive there from another GPU)
    const float2 p = __ldcg(b.pos() + e);
    pos[slot] = p;
    vel[slot] = __ldcg(b.vel() + e);
    id[slot] = __ldcg(b.id() + e);
    int row, col;
    bool clamped, outside;
    cell_of_window(k, p.x, p.y, row, col, clamped, outside);
    if (outside) {            // moved more than two columns in one step: nobody keeps it
        key[slot] = kTrashKey;
        atomicAdd(io.lost, 1u);
        return;
    }
    const uint32_t c = (uint32_t)(row * k.cols + col);
    key[slot] = c;
    rank[slot] = atomicAdd(&cell_count[c], 1u);
    if (touch != nullptr) touch[c] = epoch;      // an arrival (ghost or migrant): the cell is ranked by id again
}

int launch_bin_recv(cudaStream_t st, const Consts &k, ParticleSet &ps, const SlabIO &slab, DeviceCounters *ctr)
{
    const int threads = 2 * slab.recv[0].cap > 0 ? 2 * slab.recv[0].cap : 1;
    const int grid = (threads + kStreamThreads - 1) / kStreamThreads;
    launch_pdl(st, grid, kStreamThreads, k_bin_recv, k, slab, ps.d_n_cur, ps.d_n_in, ps.pos[ps.pc], ps.vel[ps.vc],
               ps.id[ps.ic], ps.key, ps.rank, ps.cell_count, ctr, ps.touch_ok ? ps.cell_touch : nullptr,
               (unsigned char)ps.touch_epoch);
    return 1;
}

// ------------------------------------------------------------------------ scan

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr unsigned int kFlagAggregate = 1u, kFlagPrefix = 2u;
__device__ __forceinline__ unsigned long long pack_state(unsigned int epoch, unsigned int flag, uint32_t value)
{
    return ((unsigned long long)((epoch << 2) | flag) << 32) | value;
}

// One tile = kScanTile counters.  Tile ids are handed out by an atomic so a tile only ever
// waits on tiles that already started (forward progress without relying on block order).
__global__ void __launch_bounds__(kScanThreads)
k_scan(uint32_t *__restrict__ count, uint32_t *__restrict__ start, const int n,
       unsigned long long *__restrict__ tile_state, unsigned long long *__restrict__ tile_counter,
       const unsigned long long counter_base, const unsigned int epoch, const int n_tiles,
       DeviceCounters *__restrict__ ctr, int *__restrict__ n_out)
{
    __shared__ uint32_t s_warp_sum[kScanThreads / 32];
    __shared__ uint32_t s_warp_max[kScanThreads / 32];
    __shared__ uint32_t s_tile, s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    pdl_wait();
    if (tid == 0) s_tile = (uint32_t)(atomicAdd(tile_counter, 1ULL) - counter_base);
    __syncthreads();
    const uint32_t tile = s_tile;
    const long long base = (long long)tile * kScanTile + (long long)tid * kScanItems;

    uint32_t v[kScanItems];
    if (base + kScanItems <= n) {
        const uint4 a = *reinterpret_cast<const uint4 *>(count + base);
        const uint4 b = *reinterpret_cast<const uint4 *>(count + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; i++) v[i] = (base + i < n) ? count[base + i] : 0u;
    }
    uint32_t tsum = 0, tmax = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        const uint32_t c = v[i];
        v[i] = tsum;                 // exclusive within the thread
        tsum += c;
        tmax = c > tmax ? c : tmax;
    }
    // warp inclusive scan of thread sums
    uint32_t incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, tmax, d);
        tmax = o > tmax ? o : tmax;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    if (lane == 0) s_warp_max[warp] = tmax;
    __syncthreads();
    uint32_t warp_excl = 0, aggregate = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        const uint32_t ws = s_warp_sum[w];
        if (w < warp) warp_excl += ws;
        aggregate += ws;
    }

    if (warp == 0) {
        uint32_t excl = 0;
        if (tile == 0) {
            if (lane == 0) st_volatile_u64(&tile_state[0], pack_state(epoch, kFlagPrefix, aggregate));
        } else {
            if (lane == 0) st_volatile_u64(&tile_state[tile], pack_state(epoch, kFlagAggregate, aggregate));
            int look = (int)tile - 1;
            while (true) {
                const int idx = look - lane;
                unsigned long long st;
                bool invalid;
                do {
                    st = idx >= 0 ? ld_volatile_u64(&tile_state[idx]) : pack_state(epoch, kFlagPrefix, 0u);
                    const unsigned int hi = (unsigned int)(st >> 32);
                    invalid = ((hi >> 2) != epoch) || ((hi & 3u) == 0u);
                } while (__any_sync(0xffffffffu, invalid));
                const bool is_prefix = (((unsigned int)(st >> 32)) & 3u) == kFlagPrefix;
                const unsigned int mask = __ballot_sync(0xffffffffu, is_prefix);
                const int first = mask ? (__ffs(mask) - 1) : 32;
                uint32_t contrib = (lane <= first) ? (uint32_t)st : 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
                excl += contrib;
                if (mask) break;
                look -= 32;
            }
            if (lane == 0) st_volatile_u64(&tile_state[tile], pack_state(epoch, kFlagPrefix, excl + aggregate));
        }
        if (lane == 0) {
            s_prefix = excl;
            uint32_t m = 0;
#pragma unroll
            for (int w = 0; w < kScanThreads / 32; w++) m = s_warp_max[w] > m ? s_warp_max[w] : m;
            if (m) atomicMax(&ctr->max_cell_acc, m);
            if ((int)tile == n_tiles - 1) {
                start[n] = excl + aggregate;
                if (n_out) *n_out = (int)(excl + aggregate);     // slabs: particles the sort keeps
            }
        }
    }
    __syncthreads();
    const uint32_t off = s_prefix + warp_excl + (incl - tsum);
    if (base + kScanItems <= n) {
        *reinterpret_cast<uint4 *>(start + base) = make_uint4(off + v[0], off + v[1], off + v[2], off + v[3]);
        *reinterpret_cast<uint4 *>(start + base + 4) = make_uint4(off + v[4], off + v[5], off + v[6], off + v[7]);
        *reinterpret_cast<uint4 *>(count + base) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(count + base + 4) = make_uint4(0, 0, 0, 0);
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; i++)
            if (base + i < n) { start[base + i] = off + v[i]; count[base + i] = 0u; }
    }
}

int launch_scan(cudaStream_t st, const Consts &k, ParticleSet &ps, ScanState &sc, DeviceCounters *ctr)
{
    const int n_tiles = (k.ncells + kScanTile - 1) / kScanTile;
    // the previous build's cell_start stays readable for the reorder (ParticleSet::cell_start_prev)
    if (ps.cell_start_prev) { uint32_t *t = ps.cell_start; ps.cell_start = ps.cell_start_prev; ps.cell_start_prev = t; }
    sc.epoch = (sc.epoch + 1) & 0x3fffffffu;
    if (sc.epoch == 0) sc.epoch = 1;   // 0 is the memset state ("never written")
    const unsigned long long base = sc.launches;   // tiles handed out so far
    launch_pdl(st, n_tiles, kScanThreads, k_scan, ps.cell_count, ps.cell_start, k.ncells, sc.tile_state,
               sc.tile_counter, base, sc.epoch, n_tiles, ctr, ps.d_n_cur);
    sc.launches += (unsigned long long)n_tiles;
    return 1;
}

// ------------------------------------------------------------------------ reorder

// kScatterItems slots per thread, their loads issued side by side: with the marks most threads only read a
// key and a mark and leave, and a CTA that lives for two dependent loads is bound by its own latency (one slot
// per thread: 0.26 ms for 64M slots at 1 TB/s)
constexpr int kScatterItems = 4;
__global__ void __launch_bounds__(kStreamThreads)
k_scatter_ids(const Count cnt, const uint32_t *__restrict__ key, const uint32_t *__restrict__ rank,
              const uint32_t *__restrict__ id_in, const uint32_t *__restrict__ start,
              uint32_t *__restrict__ ids_tmp, const unsigned char *__restrict__ touch, const unsigned char epoch)
{
    pdl_trigger();
    pdl_wait();
    const int n = count_of(cnt);
    const int s0 = blockIdx.x * (kStreamThreads * kScatterItems) + threadIdx.x;
    uint32_t c[kScatterItems];
    bool go[kScatterItems];
#pragma unroll
    for (int j = 0; j < kScatterItems; j++) {
        const int s = s0 + j * kStreamThreads;
        c[j] = s < n ? key[s] : kTrashKey;
    }
#pragma unroll
    for (int j = 0; j < kScatterItems; j++) {
        go[j] = c[j] != kTrashKey;
        // an unmarked cell keeps its previous order (k_reorder)
        if (go[j] && touch != nullptr) go[j] = touch[c[j]] == epoch;
    }
#pragma unroll
    for (int j = 0; j < kScatterItems; j++) {
        const int s = s0 + j * kStreamThreads;
        if (go[j]) ids_tmp[start[c[j]] + rank[s]] = id_in[s];
    }
}

template <bool DET, bool MASS, bool AUX>
__global__ void __launch_bounds__(kStreamThreads)
k_reorder(const Count cnt, const uint32_t *__restrict__ key, const uint32_t *__restrict__ rank,
          const uint32_t *__restrict__ start, const uint32_t *__restrict__ ids_tmp,
          const float2 *__restrict__ pos_in, const float2 *__restrict__ vel_in,
          const uint32_t *__restrict__ id_in, const float *__restrict__ mass_in,
          const float *__restrict__ aux_in, float2 *__restrict__ pos_out, float2 *__restrict__ vel_out,
          uint32_t *__restrict__ id_out, float *__restrict__ mass_out, float *__restrict__ aux_out,
          uint32_t *__restrict__ cellkey_out, const int cols, const uint32_t *__restrict__ start_prev,
          const unsigned char *__restrict__ touch, const unsigned char epoch)
{
    pdl_trigger();
    pdl_wait();
    const int s = blockIdx.x * kStreamThreads + threadIdx.x;
    if (s >= count_of(cnt)) return;
    const uint32_t c = key[s];
    if (c == kTrashKey) return;      // left this rank's window (slabs) — not carried over
    const uint32_t my = id_in[s];
    const uint32_t b = start[c];
    // (the three per-cell words are loaded side by side, not one after the other's branch)
    const bool keep = DET && touch != nullptr;
    const unsigned char mark = keep ? touch[c] : epoch;
    const uint32_t b_prev = keep ? start_prev[c] : 0u;
    uint32_t dst;
    if (mark != epoch) {
        // nobody left or entered this cell since the previous build: its particles are the slots
        // [start_prev[c], start_prev[c+1]) of the input, already in ascending-id order
        dst = b + ((uint32_t)s - b_prev);
    } else if (DET) {
        // rank = number of ids in my cell smaller than mine -> ascending original index,
        // the order the reference's tail-append produces (:110-123)
        const uint32_t e = start[c + 1];
        uint32_t r = 0;
        for (uint32_t j = b; j < e; j++) r += (ids_tmp[j] < my) ? 1u : 0u;
        dst = b + r;
    } else {
        dst = b + rank[s];
    }
    pos_out[dst] = pos_in[s];
    vel_out[dst] = vel_in[s];
    id_out[dst] = my;
    cellkey_out[dst] = ((c / (uint32_t)cols) << 16) | (c % (uint32_t)cols);     // row | col, < 65536 each
    if (MASS) mass_out[dst] = mass_in[s];
    if (AUX) aux_out[dst] = aux_in[s];
}

int launch_reorder(cudaStream_t st, const Consts &k, ParticleSet &ps, bool deterministic)
{
    if (ps.n == 0) return 0;
    const int grid = (ps.n + kStreamThreads - 1) / kStreamThreads;
    const Count in = ps.d_n_in ? ps.in() : ps.cur();
    int launches = 0;
    const unsigned char *touch = (deterministic && ps.touch_ok) ? ps.cell_touch : nullptr;
    const unsigned char epoch = (unsigned char)ps.touch_epoch;
    if (touch) ps.touch_builds++;
    if (deterministic) {
        const int per_cta = kStreamThreads * kScatterItems;
        launch_pdl(st, (ps.n + per_cta - 1) / per_cta, kStreamThreads, k_scatter_ids, in, ps.key, ps.rank, ps.id[ps.ic],
                   ps.cell_start, ps.ids_tmp, touch, epoch);
        launches++;
    }
    const bool has_mass = ps.mass[0] != nullptr;
    const bool has_aux = ps.aux[0] != nullptr;
    const float *mass_in = has_mass ? ps.mass[ps.mc] : nullptr;
    float *mass_out = has_mass ? ps.mass[ps.mc ^ 1] : nullptr;
    const float *aux_in = has_aux ? ps.aux[ps.xc] : nullptr;
    float *aux_out = has_aux ? ps.aux[ps.xc ^ 1] : nullptr;
#define SPHB_REORDER(D, M, A)                                                                           \
    launch_pdl(st, grid, kStreamThreads, k_reorder<D, M, A>, in, ps.key, ps.rank, ps.cell_start, ps.ids_tmp, \
        ps.pos[ps.pc], ps.vel[ps.vc], ps.id[ps.ic], mass_in, aux_in, ps.pos[ps.pc ^ 1],                  \
        ps.vel[ps.vc ^ 1], ps.id[ps.ic ^ 1], mass_out, aux_out, ps.cellkey, k.cols, ps.cell_start_prev, touch, epoch)
    if (deterministic) {
        if (has_mass && has_aux) SPHB_REORDER(true, true, true);
        else if (has_mass) SPHB_REORDER(true, true, false);
        else if (has_aux) SPHB_REORDER(true, false, true);
        else SPHB_REORDER(true, false, false);
    } else {
        if (has_mass && has_aux) SPHB_REORDER(false, true, true);
        else if (has_mass) SPHB_REORDER(false, true, false);
        else if (has_aux) SPHB_REORDER(false, false, true);
        else SPHB_REORDER(false, false, false);
    }
#undef SPHB_REORDER
    launches++;
    ps.pc ^= 1; ps.vc ^= 1; ps.ic ^= 1;
    if (has_mass) ps.mc ^= 1;
    if (has_aux) ps.xc ^= 1;
    ps.sorted = true;
    ps.counters_dirty = true;
    ps.lists_valid = false;          // the handed-over neighbour lists describe the previous order
    return launches;
}

// ------------------------------------------------------------------------ host boundary

// AoS `struct particle` (:26-31) -> SoA, identity order
__global__ void __launch_bounds__(kStreamThreads)
k_aos_to_soa(const int n, const float *__restrict__ aos, const uint32_t *__restrict__ ids, const uint32_t id_base,
             float2 *__restrict__ pos, float2 *__restrict__ vel,
             uint32_t *__restrict__ id, float *__restrict__ mass, float *__restrict__ aux,
             float2 *__restrict__ rho_prr, float *__restrict__ p, float2 *__restrict__ acc, const uint32_t m0_bits,
             unsigned int *__restrict__ mass_differs)
{
    const int i = blockIdx.x * kStreamThreads + threadIdx.x;
    if (i >= n) return;
    const float *r = aos + (size_t)i * 7;
    // are all masses the first particle's (the reference's m = RHO_0*V, :502)?  Checked here, on the fly,
    // instead of by a host loop over the caller's array (64M particles: 0.1 s)
    if (mass_differs && __float_as_uint(r[4]) != m0_bits) *mass_differs = 1u;
    pos[i] = make_float2(r[0], r[1]);
    vel[i] = make_float2(r[2], r[3]);
    id[i] = ids ? ids[i] : id_base + (uint32_t)i;
    if (mass) mass[i] = r[4];
    if (aux) aux[i] = r[5];
    if (rho_prr) rho_prr[i] = make_float2(r[5], 0.0f);
    if (p) p[i] = r[6];
    if (acc) acc[i] = make_float2(0.0f, 0.0f);
}

int launch_aos_to_soa(cudaStream_t st, const sphb_particle *aos, ParticleSet &ps, bool is_boundary, int n,
                      const uint32_t *ids, uint32_t id_base, uint32_t m0_bits, unsigned int *mass_differs)
{
    if (n < 0) n = ps.n;
    ps.pc = ps.vc = ps.ic = ps.mc = ps.xc = 0;
    ps.sorted = false;
    ps.lists_valid = false;
    if (n == 0) return 0;
    const int grid = (n + kStreamThreads - 1) / kStreamThreads;
    k_aos_to_soa<<<grid, kStreamThreads, 0, st>>>(n, reinterpret_cast<const float *>(aos), ids, id_base, ps.pos[0], ps.vel[0],
                                                  ps.id[0], ps.mass[0], is_boundary ? ps.aux[0] : nullptr,
                                                  is_boundary ? nullptr : ps.rho_prr, is_boundary ? nullptr : ps.p,
                                                  is_boundary ? nullptr : ps.acc, m0_bits, mass_differs);
    ps.sorted = false;
    return 1;
}

// SoA (sorted) -> AoS in ORIGINAL order via id[]
__global__ void __launch_bounds__(kStreamThreads)
k_soa_to_aos(const int n, const float2 *__restrict__ pos, const float2 *__restrict__ vel,
             const uint32_t *__restrict__ id, const float *__restrict__ mass, const float uniform_mass,
             const float *__restrict__ aux, const float2 *__restrict__ rho_prr, const float *__restrict__ p,
             const float2 *__restrict__ acc, float *__restrict__ aos, float *__restrict__ du,
             float *__restrict__ dv)
{
    const int s = blockIdx.x * kStreamThreads + threadIdx.x;
    if (s >= n) return;
    const uint32_t i = id[s];
    if (aos) {
        float *r = aos + (size_t)i * 7;
        const float2 x = pos[s], v = vel[s];
        r[0] = x.x; r[1] = x.y; r[2] = v.x; r[3] = v.y;
        r[4] = mass ? mass[s] : uniform_mass;
        r[5] = rho_prr ? rho_prr[s].x : (aux ? aux[s] : 0.0f);
        r[6] = p ? p[s] : 0.0f;
    }
    if (du) { const float2 a = acc[s]; du[i] = a.x; dv[i] = a.y; }
}

int launch_soa_to_aos(cudaStream_t st, const ParticleSet &ps, sphb_particle *aos, float *du, float *dv, bool is_boundary)
{
    if (ps.n == 0) return 0;
    const int grid = (ps.n + kStreamThreads - 1) / kStreamThreads;
    k_soa_to_aos<<<grid, kStreamThreads, 0, st>>>(
        ps.n, ps.pos[ps.pc], ps.vel[ps.vc], ps.id[ps.ic], ps.mass[0] ? ps.mass[ps.mc] : nullptr,
        ps.uniform_mass_value, is_boundary ? ps.aux[ps.xc] : nullptr, is_boundary ? nullptr : ps.rho_prr,
        is_boundary ? nullptr : ps.p, ps.acc, reinterpret_cast<float *>(aos), is_boundary ? nullptr : du,
        is_boundary ? nullptr : dv);
    return 1;
}

// slabs: the owned particles of this rank, compacted (arrival order), with their global ids
__global__ void __launch_bounds__(kStreamThreads)
k_pack_owned(const Consts k, const Count cnt, const float2 *__restrict__ pos, const float2 *__restrict__ vel,
             const uint32_t *__restrict__ id, const uint32_t *__restrict__ cellkey, const float uniform_mass,
             const float2 *__restrict__ rho_prr, const float *__restrict__ p, const float2 *__restrict__ acc,
             const int cap, float *__restrict__ aos, uint32_t *__restrict__ ids_out, float *__restrict__ du,
             float *__restrict__ dv, unsigned int *__restrict__ n_out)
{
    const int s = blockIdx.x * kStreamThreads + threadIdx.x;
    const bool mine = s < count_of(cnt) && owned_col(k, (int)(cellkey[s] & 0xffffu));
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(n_out, (unsigned int)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!mine) return;
    const unsigned int i = base + (unsigned int)__popc(m & ((1u << lane) - 1u));
    if (i >= (unsigned int)cap) return;
    float *r = aos + (size_t)i * 7;
    const float2 x = pos[s], v = vel[s];
    r[0] = x.x; r[1] = x.y; r[2] = v.x; r[3] = v.y;
    r[4] = uniform_mass;
    r[5] = rho_prr[s].x;
    r[6] = p[s];
    ids_out[i] = id[s];
    if (du) { const float2 a = acc[s]; du[i] = a.x; dv[i] = a.y; }
}

int launch_pack_owned(cudaStream_t st, const Consts &k, const ParticleSet &ps, int cap, sphb_particle *aos,
                      uint32_t *ids_out, float *du, float *dv, unsigned int *n_out)
{
    if (ps.n == 0) return 0;
    const int grid = (ps.n + kStreamThreads - 1) / kStreamThreads;
    k_pack_owned<<<grid, kStreamThreads, 0, st>>>(k, ps.cur(), ps.pos[ps.pc], ps.vel[ps.vc], ps.id[ps.ic], ps.cellkey,
                                                  ps.uniform_mass_value, ps.rho_prr, ps.p, ps.acc, cap,
                                                  reinterpret_cast<float *>(aos), ids_out, du, dv, n_out);
    return 1;
}

// ---- re-cut of the slabs (sphb_mg_rebalance) ------------------------------------------------------

__global__ void __launch_bounds__(kStreamThreads)
k_column_hist(const Consts k, const Count cnt, const uint32_t *__restrict__ cellkey, unsigned long long *__restrict__ hist)
{
    const int s = blockIdx.x * kStreamThreads + threadIdx.x;
    if (s >= count_of(cnt)) return;
    const int col = (int)(cellkey[s] & 0xffffu);
    if (owned_col(k, col)) atomicAdd(&hist[col + k.col_off], 1ULL);
}

int launch_column_hist(cudaStream_t st, const Consts &k, const ParticleSet &ps, unsigned long long *hist)
{
    if (ps.n == 0) return 0;
    k_column_hist<<<(ps.n + kStreamThreads - 1) / kStreamThreads, kStreamThreads, 0, st>>>(k, ps.cur(), ps.cellkey, hist);
    return 1;
}

// every owned particle as one 32-byte record in the segment of the rank that owns its column under the NEW
// cuts (order inside a segment is arrival order: the next grid build restores ascending global id per cell)
__global__ void __launch_bounds__(kStreamThreads)
k_pack_by_dest(const Consts k, const Count cnt, const float2 *__restrict__ pos, const float2 *__restrict__ vel,
               const float2 *__restrict__ acc, const uint32_t *__restrict__ id, const uint32_t *__restrict__ cellkey,
               const int *__restrict__ cuts, const int world, const unsigned long long *__restrict__ seg_off,
               unsigned long long *__restrict__ cursor, MoveRec *__restrict__ out)
{
    const int s = blockIdx.x * kStreamThreads + threadIdx.x;
    if (s >= count_of(cnt)) return;
    const int col = (int)(cellkey[s] & 0xffffu);
    if (!owned_col(k, col)) return;
    const int gcol = col + k.col_off;
    int dest = 0;
    while (dest + 1 < world && gcol >= cuts[dest + 1]) ++dest;
    const unsigned long long i = seg_off[dest] + atomicAdd(&cursor[dest], 1ULL);
    MoveRec r;
    r.pos = pos[s]; r.vel = vel[s]; r.acc = acc[s]; r.id = id[s]; r.pad = 0u;
    out[i] = r;
}

int launch_pack_by_dest(cudaStream_t st, const Consts &k, const ParticleSet &ps, const int *cuts_dev, int world,
                        const unsigned long long *seg_off_dev, unsigned long long *cursor_dev, MoveRec *out)
{
    if (ps.n == 0) return 0;
    k_pack_by_dest<<<(ps.n + kStreamThreads - 1) / kStreamThreads, kStreamThreads, 0, st>>>(
        k, ps.cur(), ps.pos[ps.pc], ps.vel[ps.vc], ps.acc, ps.id[ps.ic], ps.cellkey, cuts_dev, world, seg_off_dev, cursor_dev, out);
    return 1;
}

__global__ void __launch_bounds__(kStreamThreads)
k_unpack_moved(const int n, const MoveRec *__restrict__ in, float2 *__restrict__ pos, float2 *__restrict__ vel,
               float2 *__restrict__ acc, uint32_t *__restrict__ id)
{
    const int i = blockIdx.x * kStreamThreads + threadIdx.x;
    if (i >= n) return;
    const MoveRec r = in[i];
    pos[i] = r.pos; vel[i] = r.vel; acc[i] = r.acc; id[i] = r.id;
}

// the set becomes `n` unsorted particles (slot order = record order), as after an upload
int launch_unpack_moved(cudaStream_t st, ParticleSet &ps, const MoveRec *in, int n)
{
    ps.pc = ps.vc = ps.ic = ps.mc = ps.xc = 0;
    ps.sorted = false;
    ps.lists_valid = false;
    if (n == 0) return 0;
    k_unpack_moved<<<(n + kStreamThreads - 1) / kStreamThreads, kStreamThreads, 0, st>>>(n, in, ps.pos[0], ps.vel[0], ps.acc, ps.id[0]);
    return 1;
}

// du_dt/dv_dt in original order -> acc[] in the set's current order
__global__ void __launch_bounds__(kStreamThreads)
k_set_accel(const int n, const uint32_t *__restrict__ id, const float *__restrict__ du,
            const float *__restrict__ dv, float2 *__restrict__ acc)
{
    const int s = blockIdx.x * kStreamThreads + threadIdx.x;
    if (s >= n) return;
    const uint32_t i = id ? id[s] : (uint32_t)s;      // id == nullptr: du/dv are in slot order
    acc[s] = make_float2(du[i], dv[i]);
}

int launch_set_accel(cudaStream_t st, ParticleSet &ps, const float *du, const float *dv, int n_slots)
{
    const int n = n_slots >= 0 ? n_slots : ps.n;
    if (n == 0) return 0;
    k_set_accel<<<(n + kStreamThreads - 1) / kStreamThreads, kStreamThreads, 0, st>>>(n, n_slots >= 0 ? nullptr : ps.id[ps.ic],
                                                                                      du, dv, ps.acc);
    return 1;
}

__global__ void __launch_bounds__(kStreamThreads)
k_cell_ids(const Consts k, const int n, const float2 *__restrict__ pos, const uint32_t *__restrict__ id,
           int *__restrict__ cell_out)
{
    const int s = blockIdx.x * kStreamThreads + threadIdx.x;
    if (s >= n) return;
    const float2 p = pos[s];
    int row, col;
    bool clamped;
    cell_of(k, p.x, p.y, row, col, clamped);
    cell_out[id[s]] = row * k.cols + col;
}

int launch_cell_ids(cudaStream_t st, const Consts &k, const ParticleSet &ps, int *cell_out)
{
    if (ps.n == 0) return 0;
    const int grid = (ps.n + kStreamThreads - 1) / kStreamThreads;
    k_cell_ids<<<grid, kStreamThreads, 0, st>>>(k, ps.n, ps.pos[ps.pc], ps.id[ps.ic], cell_out);
    return 1;
}

}  // namespace sphb

// sphb_mg.cu — multi-GPU: the tank is cut into x-slabs of whole cell columns, one slab per GPU
// (the reference is single-process, so this layer is new design, not a port).
//
// Each rank keeps the particles of its owned columns [col_lo, col_hi) plus two ghost columns on
// either side, sorted on a grid that covers just that window (Consts::col_off / cols), so every
// kernel of the single-GPU step runs unchanged on it.  Per step there is ONE message per
// neighbour, built inside the advect+bin kernel (kernels_build.cu, k_advect_bin<.., SLAB>):
// every owned particle that — after the drift — lies within two columns of a cut, on either side
// of it, is appended to the message for that neighbour.  The receiver sorts the entries into its
// window: those in its owned columns are particles that migrated to it, the others are its
// ghosts.  Ghost slots are dropped at the start of the next step (their owner sends them again),
// so halo exchange and migration are the same 20-byte-per-particle message.  Two ghost columns
// make the density of the inner ghost column complete locally, which the force pass needs, so
// there is no second exchange for rho/p.  In deterministic mode the in-cell order is by global
// id on every rank, so an N-GPU run is bit-identical to the 1-GPU run.
//
// Transports:
//   NCCL        one process per GPU (torchrun / MPI style): ncclSend/ncclRecv of the two messages
//               inside one group on the rank's stream.  libnccl is dlopen'ed on first use.
//   in-process  one host thread drives several contexts (sphb_mg_group_*): the advect+bin kernel
//               stores message entries straight into the neighbour's receive buffer (peer memory
//               over NVLink when the slabs sit on different GPUs), ordered by CUDA events.
//   peer/IPC    one process per GPU, the same peer stores: every rank maps its neighbours' receive
//               blocks (cudaIpc*), a one-warp kernel publishes (count, epoch) with a system-scope
//               release store once the advect+bin kernel is complete, and the neighbour's binning
//               kernel waits for that word on the device.  No NCCL call in the step.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <nccl.h>      // types only; the entry points are resolved with dlsym

#include "sphb_internal.cuh"
#include "sph_consts.h"

namespace sphb {

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load()
{
    if (g_nccl.lib) return SPHB_OK;
    // RTLD_NOLOAD first: a host that already carries libnccl (e.g. through PyTorch) keeps its copy
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("cannot load libnccl.so.2: %s", dlerror()); return SPHB_E_COMM; }
#define SPHB_SYM(field, name)                                                      \
    *reinterpret_cast<void **>(&g_nccl.field) = dlsym(h, name);                    \
    if (!g_nccl.field) { set_error("libnccl lacks %s", name); return SPHB_E_COMM; }
    SPHB_SYM(GetUniqueId, "ncclGetUniqueId")
    SPHB_SYM(CommInitRank, "ncclCommInitRank")
    SPHB_SYM(CommDestroy, "ncclCommDestroy")
    SPHB_SYM(GroupStart, "ncclGroupStart")
    SPHB_SYM(GroupEnd, "ncclGroupEnd")
    SPHB_SYM(Send, "ncclSend")
    SPHB_SYM(Recv, "ncclRecv")
    SPHB_SYM(AllReduce, "ncclAllReduce")
    SPHB_SYM(GetErrorString, "ncclGetErrorString")
#undef SPHB_SYM
    g_nccl.lib = h;
    return SPHB_OK;
}

#define SPHB_NCCL(call)                                                                          \
    do {                                                                                         \
        ncclResult_t r__ = (call);                                                               \
        if (r__ != ncclSuccess) {                                                                \
            set_error("NCCL error %d (%s) in %s", (int)r__, g_nccl.GetErrorString(r__), #call);  \
            return SPHB_E_COMM;                                                                  \
        }                                                                                        \
    } while (0)

size_t msg_bytes(int cap) { return 16 + (size_t)cap * 20; }

}  // namespace

SlabIO mg_slab_io(const sphb_ctx *c)
{
    const MgState &m = c->mg;
    SlabIO io;
    const int q = (int)(m.exchanges & 1ULL);
    io.has[0] = m.rank > 0;
    io.has[1] = m.rank < m.world - 1;
    for (int side = 0; side < 2; side++) {
        if (!io.has[side]) continue;
        if (m.transport == 2) {
            // store straight into the neighbour's receive buffer for the side that faces us
            io.send[side].base = m.peer[side] ? m.peer[side]->mg.d_recv[1 - side][q] : nullptr;
            io.send_cnt[side] = m.d_send_cnt + side;
            io.recv[side].base = m.d_recv[side][q];
        } else if (m.transport == 3) {
            // the same across processes: the neighbour's block is mapped here, laid out like ours
            io.send[side].base = m.ipc_peer_block[side] + (size_t)((1 - side) * 2 + q) * m.recv_stride;
            io.send_cnt[side] = m.d_send_cnt + side;
            io.recv[side].base = m.d_recv[side][q];
        } else {
            io.send[side].base = m.d_send[side];
            io.send_cnt[side] = reinterpret_cast<uint32_t *>(m.d_send[side]);
            io.recv[side].base = m.d_recv[side][0];
        }
        io.send[side].cap = m.halo_cap;
        io.recv[side].cap = m.halo_cap;
    }
    io.recv[0].cap = io.recv[1].cap = m.halo_cap;
    io.lost = m.d_flags;
    io.overflow = m.d_flags + 1;
    io.capacity = m.capacity;
    // every rank makes the same number of exchanges, so the count doubles as the message's epoch
    io.wait_epoch = m.transport == 3 ? (uint32_t)(m.exchanges + 1ULL) : 0u;
    return io;
}

// both messages of this step, full capacity (the count travels in the header): sizes must be
// known to both ends without a host round trip
int mg_exchange_nccl(sphb_ctx *c)
{
    MgState &m = c->mg;
    if (m.world == 1) return SPHB_OK;
    ncclComm_t comm = static_cast<ncclComm_t>(m.nccl_comm);
    const size_t bytes = msg_bytes(m.halo_cap);
    SPHB_NCCL(g_nccl.GroupStart());
    if (m.rank > 0) {
        SPHB_NCCL(g_nccl.Send(m.d_send[0], bytes, ncclUint8, m.rank - 1, comm, c->stream));
        SPHB_NCCL(g_nccl.Recv(m.d_recv[0][0], bytes, ncclUint8, m.rank - 1, comm, c->stream));
        m.halo_bytes += bytes;
    }
    if (m.rank < m.world - 1) {
        SPHB_NCCL(g_nccl.Send(m.d_send[1], bytes, ncclUint8, m.rank + 1, comm, c->stream));
        SPHB_NCCL(g_nccl.Recv(m.d_recv[1][0], bytes, ncclUint8, m.rank + 1, comm, c->stream));
        m.halo_bytes += bytes;
    }
    SPHB_NCCL(g_nccl.GroupEnd());
    return SPHB_OK;
}

// peer stores across processes: the entries are already in the neighbours' buffers; publish the
// counts behind them (k_halo_signal).  The waiting side is k_bin_recv (SlabIO::wait_epoch).
static int mg_exchange_ipc(sphb_ctx *c)
{
    MgState &m = c->mg;
    if (m.world == 1) return SPHB_OK;
    const SlabIO io = mg_slab_io(c);
    c->launches += launch_halo_signal(c->stream, io, io.wait_epoch);
    for (int side = 0; side < 2; side++)
        if (io.has[side]) m.halo_bytes += 8;      // the entries themselves are counted on the device only
    return SPHB_OK;
}

int mg_exchange(sphb_ctx *c)
{
    return c->mg.transport == 3 ? mg_exchange_ipc(c) : mg_exchange_nccl(c);
}

// :600-601 on a slab.  psi needs every boundary neighbour, so it is computed once on the whole
// tank's grid (the boundary is replicated and small); then the boundary is re-sorted on the
// rank's window, which drops the wall particles no owned or ghost cell can see.
int mg_init_boundary(sphb_ctx *c)
{
    MgState &m = c->mg;
    ParticleSet &b = c->boundary;
    if (b.n > 0) {
        b.windowed = false;
        b.d_n_cur = nullptr;
        build_grid(c, b, false, &m.k_global);
        c->launches += launch_pseudomass(c->stream, m.k_global, b);
        int n = b.n;
        const int both[2] = {n, n};
        SPHB_CUDA(cudaMemcpyAsync(m.d_counts + 2, both, sizeof both, cudaMemcpyHostToDevice, c->stream));
        b.windowed = true;
        b.d_n_cur = m.d_counts + 2;         // the scan replaces it by the number kept
        b.d_n_in = m.d_counts + 3;          // the reorder still walks all n inputs
        b.sorted = false;                       // keys of the global grid do not apply to the window
        build_grid(c, b, false, &c->k);
        SPHB_CUDA(cudaMemcpyAsync(&n, m.d_counts + 2, sizeof n, cudaMemcpyDeviceToHost, c->stream));
        SPHB_CUDA(cudaStreamSynchronize(c->stream));
        b.n = n;
        b.d_n_cur = nullptr;
        b.d_n_in = nullptr;
        if (n == 0) b.sorted = false;
    }
    c->boundary_ready = true;
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

// A neighbour's message that never arrived (k_bin_recv gave up after its device-side time-out) makes the
// slab's state wrong from that step on: fatal.  Called where the host has just synchronised the stream.
int mg_health(sphb_ctx *c)
{
    MgState &m = c->mg;
    if (!m.on || m.transport != 3 || !m.d_flags) return SPHB_OK;
    if (!m.comm_failed) {
        unsigned int flags[2] = {0u, 0u};
        SPHB_CUDA(cudaMemcpyAsync(flags, m.d_flags, sizeof flags, cudaMemcpyDeviceToHost, c->stream));
        SPHB_CUDA(cudaStreamSynchronize(c->stream));
        m.comm_failed = (flags[1] & (1u << 30)) != 0u;
    }
    if (m.comm_failed) {
        set_error("rank %d of %d: a neighbour's halo/migration message did not arrive within the device-side time-out "
                  "(peer-store transport); the state of this slab is invalid from that step on", m.rank, m.world);
        return SPHB_E_COMM;
    }
    return SPHB_OK;
}

void mg_free(sphb_ctx *c)
{
    MgState &m = c->mg;
    if (!m.on) return;
    if (m.nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<ncclComm_t>(m.nccl_comm));
    for (int s = 0; s < 2; s++) {
        if (m.ipc_peer_block[s]) cudaIpcCloseMemHandle(m.ipc_peer_block[s]);
        cudaFree(m.d_send[s]);
    }
    cudaFree(m.d_recv_block);
    cudaFree(m.d_send_cnt); cudaFree(m.d_flags); cudaFree(m.d_counts);
    if (m.ev_sent) cudaEventDestroy(m.ev_sent);
    m = MgState();
}

}  // namespace sphb

using namespace sphb;

#define SPHB_ENTER(ctx)                                               \
    do {                                                              \
        if (!(ctx)) { set_error("null context"); return SPHB_E_ARG; } \
        SPHB_CUDA(cudaSetDevice((ctx)->device));                      \
    } while (0)

extern "C" {

int sphb_mg_configure(sphb_ctx *c, int rank, int world, int col_lo, int col_hi, int particle_capacity, int halo_capacity)
{
    SPHB_ENTER(c);
    if (c->mg.on) { set_error("already configured"); return SPHB_E_STATE; }
    if (c->fluid.n > 0 || c->boundary.n > 0) { set_error("configure the slab before uploading"); return SPHB_E_STATE; }
    const Consts kg = make_consts(c->prm, c->prm.rho0 * c->prm.vol);
    if (world < 1 || rank < 0 || rank >= world) { set_error("bad rank %d of %d", rank, world); return SPHB_E_ARG; }
    if (col_lo < 0 || col_hi > kg.cols || col_hi <= col_lo) { set_error("bad column range [%d,%d) of %d", col_lo, col_hi, kg.cols); return SPHB_E_ARG; }
    if ((rank == 0) != (col_lo == 0) || (rank == world - 1) != (col_hi == kg.cols)) {
        set_error("rank %d of %d cannot own columns [%d,%d) of %d: the slabs must tile the tank in rank order", rank, world,
                  col_lo, col_hi, kg.cols);
        return SPHB_E_ARG;
    }
    if (world > 1 && col_hi - col_lo < 4) { set_error("a slab must be at least 4 cell columns wide"); return SPHB_E_ARG; }
    MgState &m = c->mg;
    m.rank = rank; m.world = world; m.col_lo = col_lo; m.col_hi = col_hi;
    m.k_global = kg;
    m.halo_cap = halo_capacity > 0 ? ((halo_capacity + 1) & ~1) : 65536;
    m.capacity = particle_capacity;       // 0: decided at upload
    const int win_lo = rank > 0 ? col_lo - 2 : 0;
    const int win_hi = rank < world - 1 ? col_hi + 2 : kg.cols;
    set_window(c->k, win_lo, win_hi, col_lo, col_hi);

    // the scan also serves the boundary's pass on the whole tank's grid
    cudaFree(c->scan.tile_state);
    c->scan.n_tiles = (kg.ncells + kScanTile - 1) / kScanTile;
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c->scan.tile_state), sizeof(unsigned long long) * (c->scan.n_tiles + 1)));
    SPHB_CUDA(cudaMemset(c->scan.tile_state, 0, sizeof(unsigned long long) * (c->scan.n_tiles + 1)));

    const size_t bytes = msg_bytes(m.halo_cap);
    // the four receive buffers [side][parity] share one allocation, so one IPC handle exports them
    m.recv_stride = (bytes + 255) & ~(size_t)255;
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m.d_recv_block), 4 * m.recv_stride));
    for (int s = 0; s < 2; s++) {
        SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m.d_send[s]), bytes));
        SPHB_CUDA(cudaMemset(m.d_send[s], 0, 16));
        for (int q = 0; q < 2; q++) {
            m.d_recv[s][q] = m.d_recv_block + (size_t)(s * 2 + q) * m.recv_stride;
            SPHB_CUDA(cudaMemset(m.d_recv[s][q], 0, 16));
        }
    }
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m.d_send_cnt), 2 * sizeof(uint32_t)));
    SPHB_CUDA(cudaMemset(m.d_send_cnt, 0, 2 * sizeof(uint32_t)));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m.d_flags), 2 * sizeof(unsigned int)));
    SPHB_CUDA(cudaMemset(m.d_flags, 0, 2 * sizeof(unsigned int)));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m.d_counts), 4 * sizeof(int)));
    SPHB_CUDA(cudaMemset(m.d_counts, 0, 4 * sizeof(int)));
    SPHB_CUDA(cudaEventCreateWithFlags(&m.ev_sent, cudaEventDisableTiming));
    SPHB_CUDA(cudaDeviceSynchronize());
    m.on = true;
    m.transport = world == 1 ? 1 : 0;
    return SPHB_OK;
}

int sphb_mg_unique_id(char *id_out)
{
    if (!id_out) return SPHB_E_ARG;
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    SPHB_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof id <= SPHB_NCCL_ID_BYTES, "ncclUniqueId grew");
    memset(id_out, 0, SPHB_NCCL_ID_BYTES);
    memcpy(id_out, &id, sizeof id);
    return SPHB_OK;
}

int sphb_mg_connect_nccl(sphb_ctx *c, const char *id_in)
{
    SPHB_ENTER(c);
    if (!c->mg.on) { set_error("sphb_mg_configure first"); return SPHB_E_STATE; }
    if (!id_in) return SPHB_E_ARG;
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, id_in, sizeof id);
    ncclComm_t comm;
    SPHB_NCCL(g_nccl.CommInitRank(&comm, c->mg.world, id, c->mg.rank));
    c->mg.nccl_comm = comm;
    c->mg.transport = 1;
    return SPHB_OK;
}

int sphb_mg_ipc_handle(sphb_ctx *c, unsigned char *handle_out)
{
    SPHB_ENTER(c);
    if (!c->mg.on) { set_error("sphb_mg_configure first"); return SPHB_E_STATE; }
    if (!handle_out) return SPHB_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) <= SPHB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t grew");
    cudaIpcMemHandle_t h;
    SPHB_CUDA(cudaIpcGetMemHandle(&h, c->mg.d_recv_block));
    memset(handle_out, 0, SPHB_IPC_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof h);
    return SPHB_OK;
}

int sphb_mg_connect_ipc(sphb_ctx *c, const unsigned char *left_handle, const unsigned char *right_handle)
{
    SPHB_ENTER(c);
    MgState &m = c->mg;
    if (!m.on) { set_error("sphb_mg_configure first"); return SPHB_E_STATE; }
    if (m.transport == 2 || m.transport == 3) { set_error("already connected (transport %d)", m.transport); return SPHB_E_STATE; }
    if (m.exchanges != 0) { set_error("connect before the first step: the exchange count is the message epoch"); return SPHB_E_STATE; }
    const unsigned char *hs[2] = {left_handle, right_handle};
    const bool need[2] = {m.rank > 0, m.rank < m.world - 1};
    for (int side = 0; side < 2; side++)
        if (need[side] != (hs[side] != nullptr)) {
            set_error("rank %d of %d: %s handle %s", m.rank, m.world, side ? "right" : "left", need[side] ? "missing" : "given without a neighbour");
            return SPHB_E_ARG;
        }
    for (int side = 0; side < 2; side++) {
        if (!need[side]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs[side], sizeof h);
        void *p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int s2 = 0; s2 < side; s2++)
                if (m.ipc_peer_block[s2]) { cudaIpcCloseMemHandle(m.ipc_peer_block[s2]); m.ipc_peer_block[s2] = nullptr; }
            set_error("cudaIpcOpenMemHandle (%s neighbour): %s", side ? "right" : "left", cudaGetErrorString(e));
            return SPHB_E_COMM;
        }
        m.ipc_peer_block[side] = static_cast<unsigned char *>(p);
    }
    m.transport = 3;
    return SPHB_OK;
}

int sphb_mg_disconnect_ipc(sphb_ctx *c)
{
    SPHB_ENTER(c);
    MgState &m = c->mg;
    if (m.transport != 3) return SPHB_OK;
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    for (int side = 0; side < 2; side++)
        if (m.ipc_peer_block[side]) { SPHB_CUDA(cudaIpcCloseMemHandle(m.ipc_peer_block[side])); m.ipc_peer_block[side] = nullptr; }
    m.transport = (m.nccl_comm || m.world == 1) ? 1 : 0;      // back to what sphb_mg_configure / _connect_nccl left
    return SPHB_OK;
}

int sphb_mg_connect_local(sphb_ctx **ctxs, int n)
{
    if (!ctxs || n < 1) return SPHB_E_ARG;
    for (int r = 0; r < n; r++) {
        if (!ctxs[r] || !ctxs[r]->mg.on || ctxs[r]->mg.rank != r || ctxs[r]->mg.world != n) {
            set_error("context %d is not configured as rank %d of %d", r, r, n);
            return SPHB_E_ARG;
        }
        if (ctxs[r]->mg.halo_cap != ctxs[0]->mg.halo_cap) { set_error("halo capacities differ"); return SPHB_E_ARG; }
        if (r > 0 && ctxs[r]->mg.col_lo != ctxs[r - 1]->mg.col_hi) { set_error("slabs %d and %d do not abut", r - 1, r); return SPHB_E_ARG; }
    }
    for (int r = 0; r < n; r++) {
        sphb_ctx *c = ctxs[r];
        c->mg.peer[0] = r > 0 ? ctxs[r - 1] : nullptr;
        c->mg.peer[1] = r < n - 1 ? ctxs[r + 1] : nullptr;
        c->mg.transport = 2;
        SPHB_CUDA(cudaSetDevice(c->device));
        for (int side = 0; side < 2; side++) {
            sphb_ctx *p = c->mg.peer[side];
            if (!p || p->device == c->device) continue;
            int can = 0;
            SPHB_CUDA(cudaDeviceCanAccessPeer(&can, c->device, p->device));
            if (!can) { set_error("device %d cannot access device %d", c->device, p->device); return SPHB_E_COMM; }
            cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SPHB_CUDA(e);
            cudaGetLastError();
        }
    }
    return SPHB_OK;
}

int sphb_mg_upload(sphb_ctx *c, const sphb_particle *fluid, const uint32_t *ids, uint32_t id_base, int n_fluid,
                   const sphb_particle *boundary, int n_boundary)
{
    SPHB_ENTER(c);
    MgState &m = c->mg;
    if (!m.on) { set_error("sphb_mg_configure first"); return SPHB_E_STATE; }
    if (n_fluid < 0 || n_boundary < 0 || (n_fluid > 0 && !fluid) || (n_boundary > 0 && !boundary)) {
        set_error("bad particle arrays"); return SPHB_E_ARG;
    }
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    if (m.capacity <= 0) m.capacity = n_fluid + n_fluid / 4 + 4 * m.halo_cap + 1024;
    if (m.capacity < n_fluid + 2 * m.halo_cap) { set_error("particle capacity %d too small for %d particles + two messages", m.capacity, n_fluid); return SPHB_E_ARG; }
    int rc = alloc_set(c->fluid, m.capacity, c->k.ncells, false, false);
    if (rc) return rc;
    ParticleSet &f = c->fluid;
    f.n = m.capacity;                       // launch bound; the live counts are on the device
    f.d_n_cur = m.d_counts;
    f.d_n_in = m.d_counts + 1;
    f.windowed = true;
    f.uniform_mass = true;
    f.uniform_mass_value = n_fluid > 0 ? fluid[0].m : c->prm.rho0 * c->prm.vol;
    c->k.mass = f.uniform_mass_value;
    const size_t fb = (size_t)n_fluid * sizeof(sphb_particle), ib = ids ? (size_t)n_fluid * sizeof(uint32_t) : 0;
    const size_t bb = (size_t)n_boundary * sizeof(sphb_particle);
    rc = ensure_stage(c, (fb + ib > bb ? fb + ib : bb) + 64);
    if (rc) return rc;
    int counts[2] = {n_fluid, n_fluid};
    SPHB_CUDA(cudaMemcpyAsync(m.d_counts, counts, sizeof counts, cudaMemcpyHostToDevice, c->stream));
    if (n_fluid > 0) {
        char *base = static_cast<char *>(c->d_stage);
        uint32_t *d_ids = ids ? reinterpret_cast<uint32_t *>(base + ((fb + 15) & ~(size_t)15)) : nullptr;
        SPHB_CUDA(cudaMemcpyAsync(base, fluid, fb, cudaMemcpyHostToDevice, c->stream));
        if (ids) SPHB_CUDA(cudaMemcpyAsync(d_ids, ids, ib, cudaMemcpyHostToDevice, c->stream));
        // while the copies are on their way the host checks the masses
        for (int i = 1; i < n_fluid; i++)
            if (memcmp(&fluid[i].m, &fluid[0].m, sizeof(float)) != 0) {
                cudaStreamSynchronize(c->stream);
                f.n = 0;
                set_error("slab contexts need a uniform fluid mass (the reference's m = RHO_0*V, :502)");
                return SPHB_E_ARG;
            }
        c->launches += launch_aos_to_soa(c->stream, reinterpret_cast<const sphb_particle *>(base), f, false, n_fluid, d_ids, id_base);
    } else {
        f.pc = f.vc = f.ic = f.mc = f.xc = 0;
        f.sorted = false;
    }
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    if (n_boundary > 0) {
        rc = alloc_set(c->boundary, n_boundary, m.k_global.ncells, true, true);
        if (rc) return rc;
        c->boundary.uniform_mass = false;
        c->boundary.windowed = false;
        c->boundary.d_n_cur = nullptr;
        SPHB_CUDA(cudaMemcpyAsync(c->d_stage, boundary, bb, cudaMemcpyHostToDevice, c->stream));
        c->launches += launch_aos_to_soa(c->stream, reinterpret_cast<const sphb_particle *>(c->d_stage), c->boundary, true);
        SPHB_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        c->boundary.n = 0;
        c->boundary.sorted = false;
    }
    c->boundary_ready = false;
    c->accel_ready = false;
    m.n_uploaded = n_fluid;
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

// Restores du_dt/dv_dt (:492-493) of the particles just uploaded, in the same order, so that a
// run continues exactly where it was (checkpoint restart, re-cut of the slabs): the next kick uses
// them.  Must follow sphb_mg_upload directly.
int sphb_mg_upload_accel(sphb_ctx *c, const float *du_dt, const float *dv_dt)
{
    SPHB_ENTER(c);
    MgState &m = c->mg;
    if (!m.on) { set_error("not a slab context"); return SPHB_E_STATE; }
    if (c->fluid.sorted) { set_error("sphb_mg_upload_accel must directly follow sphb_mg_upload"); return SPHB_E_STATE; }
    const int n = m.n_uploaded;
    if (n > 0 && (!du_dt || !dv_dt)) return SPHB_E_ARG;
    if (n > 0) {
        const size_t db = (size_t)n * sizeof(float);
        SPHB_CUDA(cudaStreamSynchronize(c->stream));
        int rc = ensure_stage(c, 2 * db + 64);
        if (rc) return rc;
        float *d_du = static_cast<float *>(c->d_stage), *d_dv = d_du + n;
        SPHB_CUDA(cudaMemcpyAsync(d_du, du_dt, db, cudaMemcpyHostToDevice, c->stream));
        SPHB_CUDA(cudaMemcpyAsync(d_dv, dv_dt, db, cudaMemcpyHostToDevice, c->stream));
        c->launches += launch_set_accel(c->stream, c->fluid, d_du, d_dv, n);
        SPHB_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->accel_ready = true;
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

int sphb_mg_download(sphb_ctx *c, int cap, sphb_particle *fluid_out, uint32_t *ids_out, float *du_dt, float *dv_dt, int *n_out)
{
    SPHB_ENTER(c);
    if (!c->mg.on) { set_error("not a slab context"); return SPHB_E_STATE; }
    if (cap < 0 || !n_out || (cap > 0 && (!fluid_out || !ids_out)) || ((du_dt == nullptr) != (dv_dt == nullptr))) return SPHB_E_ARG;
    *n_out = 0;
    if (!c->fluid.sorted) { set_error("no sorted state yet (sphb_compute_accel first)"); return SPHB_E_STATE; }
    const size_t ab = ((size_t)cap * sizeof(sphb_particle) + 15) & ~(size_t)15, ib = ((size_t)cap * 4 + 15) & ~(size_t)15;
    int rc = ensure_stage(c, ab + 3 * ib + 64);
    if (rc) return rc;
    char *base = static_cast<char *>(c->d_stage);
    sphb_particle *d_aos = reinterpret_cast<sphb_particle *>(base);
    uint32_t *d_ids = reinterpret_cast<uint32_t *>(base + ab);
    float *d_du = reinterpret_cast<float *>(base + ab + ib), *d_dv = reinterpret_cast<float *>(base + ab + 2 * ib);
    unsigned int *d_n = reinterpret_cast<unsigned int *>(base + ab + 3 * ib);
    SPHB_CUDA(cudaMemsetAsync(d_n, 0, sizeof(unsigned int), c->stream));
    c->launches += launch_pack_owned(c->stream, c->k, c->fluid, cap, d_aos, d_ids, du_dt ? d_du : nullptr,
                                     du_dt ? d_dv : nullptr, d_n);
    unsigned int n = 0;
    SPHB_CUDA(cudaMemcpyAsync(&n, d_n, sizeof n, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    *n_out = (int)n;
    rc = mg_health(c);
    if (rc) return rc;
    if ((int)n > cap) { set_error("%u owned particles exceed the caller's capacity %d", n, cap); return SPHB_E_ARG; }
    if (n > 0) {
        SPHB_CUDA(cudaMemcpyAsync(fluid_out, d_aos, (size_t)n * sizeof(sphb_particle), cudaMemcpyDeviceToHost, c->stream));
        SPHB_CUDA(cudaMemcpyAsync(ids_out, d_ids, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (du_dt) {
            SPHB_CUDA(cudaMemcpyAsync(du_dt, d_du, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
            SPHB_CUDA(cudaMemcpyAsync(dv_dt, d_dv, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        }
        SPHB_CUDA(cudaStreamSynchronize(c->stream));
    }
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

// ---- in-process group: one host thread, several slabs ----------------------------------------

static int group_check(sphb_ctx **ctxs, int n)
{
    if (!ctxs || n < 1) { set_error("bad group"); return SPHB_E_ARG; }
    for (int r = 0; r < n; r++)
        if (!ctxs[r] || !ctxs[r]->mg.on || ctxs[r]->mg.transport != 2) {
            set_error("context %d is not part of a connected in-process group (sphb_mg_connect_local)", r);
            return SPHB_E_STATE;
        }
    return SPHB_OK;
}

static int group_pass(sphb_ctx **ctxs, int n, bool advect, float gx, float gy)
{
    // phase A everywhere: the kernels store the messages into the neighbours' receive buffers;
    // then the 4-byte counts follow on the same stream and an event marks "sent"
    for (int r = 0; r < n; r++) {
        sphb_ctx *c = ctxs[r];
        SPHB_CUDA(cudaSetDevice(c->device));
        step_phase_a(c, advect);
        const int q = (int)(c->mg.exchanges & 1ULL);
        for (int side = 0; side < 2; side++) {
            sphb_ctx *p = c->mg.peer[side];
            if (!p) continue;
            SPHB_CUDA(cudaMemcpyAsync(p->mg.d_recv[1 - side][q], c->mg.d_send_cnt + side, sizeof(uint32_t),
                                      cudaMemcpyDeviceToDevice, c->stream));
            c->mg.halo_bytes += 4;
        }
        SPHB_CUDA(cudaEventRecord(c->mg.ev_sent, c->stream));
    }
    // phase B everywhere, after both neighbours' messages are complete
    for (int r = 0; r < n; r++) {
        sphb_ctx *c = ctxs[r];
        SPHB_CUDA(cudaSetDevice(c->device));
        for (int side = 0; side < 2; side++)
            if (c->mg.peer[side]) SPHB_CUDA(cudaStreamWaitEvent(c->stream, c->mg.peer[side]->mg.ev_sent, 0));
        step_phase_b(c, gx, gy, advect);
    }
    return SPHB_OK;
}

int sphb_mg_group_compute_accel(sphb_ctx **ctxs, int n, float gx, float gy)
{
    int rc = group_check(ctxs, n);
    if (rc) return rc;
    for (int r = 0; r < n; r++)
        if (ctxs[r]->boundary.n > 0 && !ctxs[r]->boundary_ready) { set_error("sphb_init_boundary not called on rank %d", r); return SPHB_E_STATE; }
    rc = group_pass(ctxs, n, false, gx, gy);
    if (rc) return rc;
    for (int r = 0; r < n; r++) ctxs[r]->accel_ready = true;
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

int sphb_mg_group_step(sphb_ctx **ctxs, int n, float gx, float gy, const float *gravity_xy, int nsteps)
{
    int rc = group_check(ctxs, n);
    if (rc) return rc;
    if (nsteps < 0) return SPHB_E_ARG;
    for (int r = 0; r < n; r++)
        if (!ctxs[r]->accel_ready) { set_error("sphb_mg_group_compute_accel must run first"); return SPHB_E_STATE; }
    for (int s = 0; s < nsteps; s++) {
        if (gravity_xy) { gx = gravity_xy[2 * s]; gy = gravity_xy[2 * s + 1]; }
        rc = group_pass(ctxs, n, true, gx, gy);
        if (rc) return rc;
        for (int r = 0; r < n; r++) ctxs[r]->steps++;
    }
    SPHB_CUDA(cudaGetLastError());
    return SPHB_OK;
}

int sphb_mg_group_synchronize(sphb_ctx **ctxs, int n)
{
    if (!ctxs) return SPHB_E_ARG;
    for (int r = 0; r < n; r++) {
        int rc = sphb_synchronize(ctxs[r]);
        if (rc) return rc;
    }
    return SPHB_OK;
}

// ---- statistics across slabs ------------------------------------------------------------------

int sphb_mg_merge_stats(const sphb_stats *in, int n, sphb_stats *out)
{
    if (!in || !out || n < 1) return SPHB_E_ARG;
    sphb_stats o = in[0];
    for (int r = 1; r < n; r++) {
        const sphb_stats &s = in[r];
        o.mass += s.mass; o.mom_x += s.mom_x; o.mom_y += s.mom_y; o.kinetic += s.kinetic;
        if (s.n_fluid > 0) {
            if (o.n_fluid == 0 || s.max_rho > o.max_rho) o.max_rho = s.max_rho;
            if (o.n_fluid == 0 || s.min_rho < o.min_rho) o.min_rho = s.min_rho;
            if (o.n_fluid == 0 || s.max_rho_err > o.max_rho_err) o.max_rho_err = s.max_rho_err;
        }
        if (s.max_speed > o.max_speed) o.max_speed = s.max_speed;
        if (s.max_cell_count > o.max_cell_count) o.max_cell_count = s.max_cell_count;
        o.n_escaped += s.n_escaped; o.n_fluid += s.n_fluid;
        o.n_lost += s.n_lost; o.n_overflow += s.n_overflow;
        if (s.n_boundary > o.n_boundary) o.n_boundary = s.n_boundary;
    }
    o.last_rho_err_ref = 0.0f;      // the reference's buggy scan (:657-659) has no slab meaning
    *out = o;
    return SPHB_OK;
}

// NCCL transport: every rank gets the merged statistics (two tiny all-reduces: sums and maxima)
int sphb_mg_allreduce_stats(sphb_ctx *c, sphb_stats *inout)
{
    SPHB_ENTER(c);
    if (!inout) return SPHB_E_ARG;
    MgState &m = c->mg;
    if (!m.on || (!m.nccl_comm && m.world > 1)) { set_error("not an NCCL slab context (sphb_mg_connect_nccl)"); return SPHB_E_STATE; }
    if (m.world == 1) return SPHB_OK;
    int rc = ensure_stage(c, 256);
    if (rc) return rc;
    double h[16];
    memset(h, 0, sizeof h);
    const bool has = inout->n_fluid > 0;
    h[0] = inout->mass; h[1] = inout->mom_x; h[2] = inout->mom_y; h[3] = inout->kinetic;
    h[4] = inout->n_escaped; h[5] = inout->n_fluid; h[6] = inout->n_lost; h[7] = inout->n_overflow;
    h[8] = inout->max_speed; h[9] = has ? inout->max_rho : -INFINITY; h[10] = has ? -(double)inout->min_rho : -INFINITY;
    h[11] = inout->max_cell_count; h[12] = inout->n_boundary;
    double *d = static_cast<double *>(c->d_stage);
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaMemcpyAsync(d, h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    ncclComm_t comm = static_cast<ncclComm_t>(m.nccl_comm);
    SPHB_NCCL(g_nccl.GroupStart());
    SPHB_NCCL(g_nccl.AllReduce(d, d, 8, ncclDouble, ncclSum, comm, c->stream));
    SPHB_NCCL(g_nccl.AllReduce(d + 8, d + 8, 8, ncclDouble, ncclMax, comm, c->stream));
    SPHB_NCCL(g_nccl.GroupEnd());
    SPHB_CUDA(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    inout->mass = h[0]; inout->mom_x = h[1]; inout->mom_y = h[2]; inout->kinetic = h[3];
    inout->n_escaped = (unsigned int)h[4]; inout->n_fluid = (unsigned int)h[5];
    inout->n_lost = (unsigned int)h[6]; inout->n_overflow = (unsigned int)h[7];
    inout->max_speed = (float)h[8]; inout->max_rho = (float)h[9]; inout->min_rho = (float)-h[10];
    inout->max_rho_err = inout->max_rho - c->prm.rho0;
    inout->max_cell_count = (unsigned int)h[11]; inout->n_boundary = (unsigned int)h[12];
    inout->last_rho_err_ref = 0.0f;
    return SPHB_OK;
}

int sphb_mg_info(sphb_ctx *c, sphb_mg_info_t *out)
{
    if (!c || !out) return SPHB_E_ARG;
    memset(out, 0, sizeof *out);
    const MgState &m = c->mg;
    out->rank = m.rank; out->world = m.world; out->col_lo = m.col_lo; out->col_hi = m.col_hi;
    out->window_lo = c->k.col_off; out->window_hi = c->k.col_off + c->k.cols;
    out->halo_capacity = m.halo_cap; out->particle_capacity = m.capacity; out->transport = m.transport;
    out->message_bytes = (unsigned long long)msg_bytes(m.halo_cap);
    out->bytes_sent = m.halo_bytes;
    out->exchanges = m.exchanges;
    return SPHB_OK;
}

}  // extern "C"

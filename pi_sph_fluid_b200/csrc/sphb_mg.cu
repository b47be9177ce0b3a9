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
// the boundary set as it stands (the whole tank's wall particles, psi computed, sorted on the tank's grid)
// re-sorted on this rank's window; the sort reads those buffers and writes the other pair, so the whole set
// survives there for the next re-cut
static int mg_window_boundary(sphb_ctx *c)
{
    MgState &m = c->mg;
    ParticleSet &b = c->boundary;
    int n = b.n = m.bnd_n_global;
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
    return SPHB_OK;
}

int mg_init_boundary(sphb_ctx *c)
{
    MgState &m = c->mg;
    ParticleSet &b = c->boundary;
    if (b.n > 0) {
        b.windowed = false;
        b.d_n_cur = nullptr;
        build_grid(c, b, false, &m.k_global);
        c->launches += launch_pseudomass(c->stream, m.k_global, b);
        m.bnd_n_global = b.n;
        int rc = mg_window_boundary(c);
        if (rc) return rc;
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

// ---- re-cut across processes ("re-cut every K steps") ----------------------------------------
// A dam break drains the left slabs: cuts made at t = 0 go stale.  sphb_mg_rebalance is collective — every
// rank calls it after the same step: the per-column counts of all ranks are summed (one all-reduce; the old
// cuts ride in the same buffer), every rank plans the same new cuts from them and, from the same histogram,
// knows how many particles go from every rank to every other, so the particles (position, velocity,
// du_dt/dv_dt, global id: 32 bytes each) travel in ONE grouped exchange with no count handshake.  The rank
// then holds its new particles as an unsorted set — the state sphb_mg_upload + sphb_mg_upload_accel leave —
// on the window of its new columns, and the run continues bit-identically.

namespace sphb {
namespace {

struct Exchange {                        // how bytes travel between the ranks of this run
    bool nccl = true;
    sphb_mg_allreduce_u64_fn ar = nullptr;
    sphb_mg_alltoallv_fn a2a = nullptr;
    void *user = nullptr;
};

// d_buf: this rank's values -> h_buf: the sums over all ranks
int xch_allreduce(sphb_ctx *c, const Exchange &x, unsigned long long *d_buf, unsigned long long *h_buf, int count)
{
    if (x.nccl) {
        SPHB_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)count, ncclUint64, ncclSum, static_cast<ncclComm_t>(c->mg.nccl_comm), c->stream));
        SPHB_CUDA(cudaMemcpyAsync(h_buf, d_buf, (size_t)count * 8, cudaMemcpyDeviceToHost, c->stream));
        SPHB_CUDA(cudaStreamSynchronize(c->stream));
        return SPHB_OK;
    }
    SPHB_CUDA(cudaMemcpyAsync(h_buf, d_buf, (size_t)count * 8, cudaMemcpyDeviceToHost, c->stream));
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    if (x.ar(x.user, h_buf, count) != 0) { set_error("sphb_mg_rebalance_host: the all-reduce callback failed"); return SPHB_E_COMM; }
    return SPHB_OK;
}

// counts and offsets in records; segment r of d_send goes to rank r, segment r of d_recv comes from rank r
int xch_alltoallv(sphb_ctx *c, const Exchange &x, const MoveRec *d_send, const unsigned long long *scnt,
                  const unsigned long long *soff, MoveRec *d_recv, const unsigned long long *rcnt, const unsigned long long *roff)
{
    const MgState &m = c->mg;
    if (x.nccl) {
        ncclComm_t comm = static_cast<ncclComm_t>(m.nccl_comm);
        if (scnt[m.rank])
            SPHB_CUDA(cudaMemcpyAsync(d_recv + roff[m.rank], d_send + soff[m.rank], (size_t)scnt[m.rank] * sizeof(MoveRec),
                                      cudaMemcpyDeviceToDevice, c->stream));
        SPHB_NCCL(g_nccl.GroupStart());
        for (int r = 0; r < m.world; r++) {
            if (r == m.rank) continue;
            if (scnt[r]) SPHB_NCCL(g_nccl.Send(d_send + soff[r], (size_t)scnt[r] * sizeof(MoveRec), ncclUint8, r, comm, c->stream));
            if (rcnt[r]) SPHB_NCCL(g_nccl.Recv(d_recv + roff[r], (size_t)rcnt[r] * sizeof(MoveRec), ncclUint8, r, comm, c->stream));
        }
        SPHB_NCCL(g_nccl.GroupEnd());
        SPHB_CUDA(cudaStreamSynchronize(c->stream));
        return SPHB_OK;
    }
    // host-carried: stage both sides through pinned memory and hand byte counts to the caller's exchange
    unsigned long long ns = 0, nr = 0;
    for (int r = 0; r < m.world; r++) { ns += scnt[r]; nr += rcnt[r]; }
    void *h_send = nullptr, *h_recv = nullptr;
    SPHB_CUDA(cudaMallocHost(&h_send, (size_t)(ns ? ns : 1) * sizeof(MoveRec)));
    cudaError_t e = cudaMallocHost(&h_recv, (size_t)(nr ? nr : 1) * sizeof(MoveRec));
    if (e != cudaSuccess) { cudaFreeHost(h_send); SPHB_CUDA(e); }
    int rc = SPHB_OK;
    unsigned long long *bytes = static_cast<unsigned long long *>(malloc(sizeof(unsigned long long) * 4 * (size_t)m.world));
    if (!bytes) rc = SPHB_E_NOMEM;
    if (!rc) {
        for (int r = 0; r < m.world; r++) {
            bytes[r] = scnt[r] * sizeof(MoveRec); bytes[m.world + r] = soff[r] * sizeof(MoveRec);
            bytes[2 * m.world + r] = rcnt[r] * sizeof(MoveRec); bytes[3 * m.world + r] = roff[r] * sizeof(MoveRec);
        }
        e = cudaMemcpyAsync(h_send, d_send, (size_t)ns * sizeof(MoveRec), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "staging the send buffer", __FILE__, __LINE__);
    }
    if (!rc && x.a2a(x.user, h_send, bytes, bytes + m.world, h_recv, bytes + 2 * m.world, bytes + 3 * m.world) != 0) {
        set_error("sphb_mg_rebalance_host: the all-to-all callback failed");
        rc = SPHB_E_COMM;
    }
    if (!rc) {
        e = cudaMemcpyAsync(d_recv, h_recv, (size_t)nr * sizeof(MoveRec), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "staging the receive buffer", __FILE__, __LINE__);
    }
    free(bytes);
    cudaFreeHost(h_send); cudaFreeHost(h_recv);
    return rc;
}

struct DevBufs {                         // scratch of one re-cut, freed on every path
    unsigned long long *hist = nullptr, *segoff = nullptr, *cursor = nullptr;
    int *cuts = nullptr;
    MoveRec *send = nullptr, *recv = nullptr;
    ~DevBufs() { cudaFree(hist); cudaFree(segoff); cudaFree(cursor); cudaFree(cuts); cudaFree(send); cudaFree(recv); }
};

int rebalance_impl(sphb_ctx *c, int min_width, double column_cost, double min_imbalance, const Exchange &x, int *changed_out)
{
    MgState &m = c->mg;
    if (changed_out) *changed_out = 0;
    if (!m.on) { set_error("not a slab context"); return SPHB_E_STATE; }
    if (!c->fluid.sorted || !c->accel_ready) { set_error("re-cut needs a stepped state (sphb_compute_accel / sphb_step first)"); return SPHB_E_STATE; }
    if (m.world == 1) return SPHB_OK;
    if (x.nccl && !m.nccl_comm) { set_error("sphb_mg_rebalance moves particles over NCCL: sphb_mg_connect_nccl first (or use sphb_mg_rebalance_host)"); return SPHB_E_STATE; }
    if (m.transport == 2) { set_error("in-process groups re-cut through sphb_mg_download / sphb_mg_upload"); return SPHB_E_STATE; }
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    int rc = mg_health(c);
    if (rc) return rc;
    const int world = m.world, gcols = m.k_global.cols, count = gcols + world;
    DevBufs d;
    unsigned long long *h = static_cast<unsigned long long *>(calloc((size_t)count + 6 * (size_t)world + 8, sizeof(unsigned long long)));
    int *cuts_old = static_cast<int *>(malloc(sizeof(int) * 2 * ((size_t)world + 1)));
    if (!h || !cuts_old) { free(h); free(cuts_old); return SPHB_E_NOMEM; }
    int *cuts_new = cuts_old + world + 1;
    unsigned long long *scnt = h + count, *soff = scnt + world, *rcnt = soff + world, *roff = rcnt + world, *own = roff + world;
    struct Free { void *a, *b; ~Free() { free(a); free(b); } } fr{h, cuts_old};

    // 1. per-column counts of every rank, summed; the old cuts ride behind them
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d.hist), (size_t)count * 8));
    SPHB_CUDA(cudaMemsetAsync(d.hist, 0, (size_t)count * 8, c->stream));
    c->launches += launch_column_hist(c->stream, c->k, c->fluid, d.hist);
    const unsigned long long my_lo = (unsigned long long)m.col_lo;
    SPHB_CUDA(cudaMemcpyAsync(d.hist + gcols + m.rank, &my_lo, 8, cudaMemcpyHostToDevice, c->stream));
    rc = xch_allreduce(c, x, d.hist, h, count);
    if (rc) return rc;
    for (int r = 0; r < world; r++) cuts_old[r] = (int)h[gcols + r];
    cuts_old[world] = gcols;
    if (cuts_old[m.rank] != m.col_lo || cuts_old[m.rank + 1] != m.col_hi) {
        set_error("rank %d: the ranks disagree about the current cuts (was sphb_mg_rebalance called on every rank after the same step?)", m.rank);
        return SPHB_E_COMM;
    }
    unsigned long long total = 0, most = 0;
    for (int r = 0; r < world; r++) {
        own[r] = 0;
        for (int col = cuts_old[r]; col < cuts_old[r + 1]; col++) own[r] += h[col];
        total += own[r];
        most = own[r] > most ? own[r] : most;
    }
    if (min_imbalance > 0.0 && total > 0 && (double)most * world <= min_imbalance * (double)total) return SPHB_OK;

    // 2. the new cuts (every rank computes the same)
    rc = sphb_mg_plan_cuts_cost(h, gcols, world, min_width < 4 ? 4 : min_width, column_cost, cuts_new);
    if (rc) { set_error("cannot cut %d columns into %d slabs", gcols, world); return rc; }
    bool same = true;
    for (int r = 0; r <= world; r++) same = same && cuts_new[r] == cuts_old[r];
    if (same) return SPHB_OK;

    // 3. who sends how much to whom follows from the histogram: no count handshake
    unsigned long long n_send = 0, n_new = 0;
    for (int r = 0; r < world; r++) {
        scnt[r] = rcnt[r] = 0;
        const int slo = cuts_new[r] > m.col_lo ? cuts_new[r] : m.col_lo, shi = cuts_new[r + 1] < m.col_hi ? cuts_new[r + 1] : m.col_hi;
        for (int col = slo; col < shi; col++) scnt[r] += h[col];
        const int rlo = cuts_old[r] > cuts_new[m.rank] ? cuts_old[r] : cuts_new[m.rank];
        const int rhi = cuts_old[r + 1] < cuts_new[m.rank + 1] ? cuts_old[r + 1] : cuts_new[m.rank + 1];
        for (int col = rlo; col < rhi; col++) rcnt[r] += h[col];
        soff[r] = n_send; roff[r] = n_new;
        n_send += scnt[r]; n_new += rcnt[r];
    }
    if (n_send != own[m.rank]) { set_error("rank %d: histogram and owned count disagree", m.rank); return SPHB_E_STATE; }
    if (n_new > 2000000000ULL) { set_error("rank %d would own %llu particles", m.rank, n_new); return SPHB_E_ARG; }

    // 4. pack by destination, exchange
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d.send), (size_t)(n_send ? n_send : 1) * sizeof(MoveRec)));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d.recv), (size_t)(n_new ? n_new : 1) * sizeof(MoveRec)));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d.cuts), sizeof(int) * ((size_t)world + 1)));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d.segoff), 8 * (size_t)world));
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d.cursor), 8 * (size_t)world));
    SPHB_CUDA(cudaMemcpyAsync(d.cuts, cuts_new, sizeof(int) * ((size_t)world + 1), cudaMemcpyHostToDevice, c->stream));
    SPHB_CUDA(cudaMemcpyAsync(d.segoff, soff, 8 * (size_t)world, cudaMemcpyHostToDevice, c->stream));
    SPHB_CUDA(cudaMemsetAsync(d.cursor, 0, 8 * (size_t)world, c->stream));
    c->launches += launch_pack_by_dest(c->stream, c->k, c->fluid, d.cuts, world, d.segoff, d.cursor, d.send);
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    rc = xch_alltoallv(c, x, d.send, scnt, soff, d.recv, rcnt, roff);
    if (rc) return rc;

    // 5. this rank's new slab: window, slots, particles, boundary
    m.col_lo = cuts_new[m.rank];
    m.col_hi = cuts_new[m.rank + 1];
    const int win_lo = m.rank > 0 ? m.col_lo - 2 : 0, win_hi = m.rank < world - 1 ? m.col_hi + 2 : gcols;
    Consts k2 = m.k_global;
    k2.mass = c->k.mass;
    set_window(k2, win_lo, win_hi, m.col_lo, m.col_hi);
    c->k = k2;
    const long long need = (long long)n_new + 2LL * m.halo_cap;
    if (need > m.capacity) m.capacity = (int)(n_new + n_new / 4 + 4ULL * m.halo_cap + 1024ULL);
    ParticleSet &f = c->fluid;
    const float mass_value = f.uniform_mass_value;
    free_set_public(f);
    rc = alloc_set(f, m.capacity, c->k.ncells, false, false);
    if (rc) return rc;
    f.n = m.capacity;
    f.d_n_cur = m.d_counts;
    f.d_n_in = m.d_counts + 1;
    f.windowed = true;
    f.uniform_mass = true;
    f.uniform_mass_value = mass_value;
    const int counts[2] = {(int)n_new, (int)n_new};
    SPHB_CUDA(cudaMemcpyAsync(m.d_counts, counts, sizeof counts, cudaMemcpyHostToDevice, c->stream));
    c->launches += launch_unpack_moved(c->stream, f, d.recv, (int)n_new);
    m.n_uploaded = (int)n_new;
    if (m.bnd_n_global > 0) {
        ParticleSet &b = c->boundary;
        b.pc ^= 1; b.vc ^= 1; b.ic ^= 1; b.mc ^= 1; b.xc ^= 1;      // back to the buffers that hold the whole tank's walls
        rc = mg_window_boundary(c);
        if (rc) return rc;
    }
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaGetLastError());
    if (changed_out) *changed_out = 1;
    return SPHB_OK;
}

}  // namespace
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
    SPHB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m.d_flags), 4 * sizeof(unsigned int)));      // [2]: scratch of the upload
    SPHB_CUDA(cudaMemset(m.d_flags, 0, 4 * sizeof(unsigned int)));
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
        // slabs need the reference's uniform fluid mass (:502): the conversion kernel checks it on the fly
        uint32_t m0_bits;
        memcpy(&m0_bits, &fluid[0].m, sizeof m0_bits);
        unsigned int *d_differs = m.d_flags + 2;
        unsigned int differs = 0;
        SPHB_CUDA(cudaMemsetAsync(d_differs, 0, sizeof(unsigned int), c->stream));
        c->launches += launch_aos_to_soa(c->stream, reinterpret_cast<const sphb_particle *>(base), f, false, n_fluid, d_ids, id_base,
                                         m0_bits, d_differs);
        SPHB_CUDA(cudaMemcpyAsync(&differs, d_differs, sizeof differs, cudaMemcpyDeviceToHost, c->stream));
        SPHB_CUDA(cudaStreamSynchronize(c->stream));
        if (differs) {
            f.n = 0;
            set_error("slab contexts need a uniform fluid mass (the reference's m = RHO_0*V, :502)");
            return SPHB_E_ARG;
        }
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

// ---- state files of a slab run ---------------------------------------------------------------
// One PART file per rank: 64-byte header (version 2: rank, world, owned columns) | sphb_params | global ids[n] |
// struct particle fluid[n] | du_dt[n] | dv_dt[n] | the whole tank's boundary[nb] (every part carries it: it is
// small and makes each part self-describing).  sphb_mg_load_state gives a freshly configured slab context the
// particles of ALL parts that fall into ITS columns — so a run saved on N ranks continues on M ranks (M = 1:
// one slab over all columns), bit-identically.
struct PartHeader {
    char magic[8];                  // "SPHB200\0"
    uint32_t version, params_bytes;
    uint32_t n_fluid, n_boundary;
    unsigned long long steps;
    uint32_t has_accel, rank, world, col_lo, col_hi, reserved[3];
};
static_assert(sizeof(PartHeader) == 64, "header layout");

int sphb_mg_save_state(sphb_ctx *c, const char *path)
{
    SPHB_ENTER(c);
    MgState &m = c->mg;
    if (!path) return SPHB_E_ARG;
    if (!m.on) { set_error("not a slab context (single GPU: sphb_save_state)"); return SPHB_E_STATE; }
    if (!c->fluid.sorted) { set_error("no sorted state yet (sphb_compute_accel first)"); return SPHB_E_STATE; }
    int n_cur = 0;
    SPHB_CUDA(cudaStreamSynchronize(c->stream));
    SPHB_CUDA(cudaMemcpy(&n_cur, m.d_counts, sizeof n_cur, cudaMemcpyDeviceToHost));
    const int cap = n_cur > 0 ? n_cur : 1, nb = m.bnd_n_global;
    sphb_particle *f = static_cast<sphb_particle *>(malloc(sizeof(sphb_particle) * (size_t)cap));
    uint32_t *ids = static_cast<uint32_t *>(malloc(4 * (size_t)cap));
    float *du = static_cast<float *>(malloc(4 * (size_t)cap)), *dv = static_cast<float *>(malloc(4 * (size_t)cap));
    sphb_particle *b = static_cast<sphb_particle *>(malloc(sizeof(sphb_particle) * (size_t)(nb > 0 ? nb : 1)));
    int rc = (f && ids && du && dv && b) ? SPHB_OK : SPHB_E_NOMEM, n = 0;
    if (!rc) rc = sphb_mg_download(c, cap, f, ids, du, dv, &n);
    if (!rc && nb > 0) {
        // the whole tank's walls live in the buffers the windowed sort read from (mg_window_boundary)
        ParticleSet g = c->boundary;
        g.pc ^= 1; g.vc ^= 1; g.ic ^= 1; g.mc ^= 1; g.xc ^= 1;
        g.n = nb;
        rc = ensure_stage(c, (size_t)nb * sizeof(sphb_particle) + 64);
        if (!rc) {
            c->launches += launch_soa_to_aos(c->stream, g, static_cast<sphb_particle *>(c->d_stage), nullptr, nullptr, true);
            cudaError_t e = cudaMemcpyAsync(b, c->d_stage, (size_t)nb * sizeof(sphb_particle), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "boundary download", __FILE__, __LINE__);
        }
    }
    if (!rc) {
        PartHeader h;
        memset(&h, 0, sizeof h);
        memcpy(h.magic, "SPHB200", 8);
        h.version = 2; h.params_bytes = (uint32_t)sizeof(sphb_params);
        h.n_fluid = (uint32_t)n; h.n_boundary = (uint32_t)nb; h.steps = c->steps; h.has_accel = c->accel_ready ? 1u : 0u;
        h.rank = (uint32_t)m.rank; h.world = (uint32_t)m.world; h.col_lo = (uint32_t)m.col_lo; h.col_hi = (uint32_t)m.col_hi;
        FILE *fp = fopen(path, "wb");
        bool ok = fp != nullptr;
        ok = ok && fwrite(&h, sizeof h, 1, fp) == 1 && fwrite(&c->prm, sizeof c->prm, 1, fp) == 1;
        ok = ok && (n == 0 || (fwrite(ids, 4, n, fp) == (size_t)n && fwrite(f, sizeof *f, n, fp) == (size_t)n &&
                               fwrite(du, 4, n, fp) == (size_t)n && fwrite(dv, 4, n, fp) == (size_t)n));
        ok = ok && (nb == 0 || fwrite(b, sizeof *b, nb, fp) == (size_t)nb);
        if (fp) ok = (fclose(fp) == 0) && ok;
        if (!ok) { set_error("cannot write state file %s", path); rc = SPHB_E_ARG; }
    }
    free(f); free(ids); free(du); free(dv); free(b);
    return rc;
}

int sphb_mg_load_state(sphb_ctx *c, const char *const *paths, int n_paths)
{
    SPHB_ENTER(c);
    MgState &m = c->mg;
    if (!paths || n_paths < 1) return SPHB_E_ARG;
    if (!m.on) { set_error("sphb_mg_configure the context first (its columns decide which particles it takes)"); return SPHB_E_STATE; }
    sphb_particle *f = nullptr, *b = nullptr;
    uint32_t *ids = nullptr;
    float *du = nullptr, *dv = nullptr;
    size_t n = 0, cap = 0, nb = 0;
    unsigned long long steps = 0;
    bool has_accel = true;
    int rc = SPHB_OK;
    for (int p = 0; p < n_paths && !rc; p++) {
        FILE *fp = fopen(paths[p], "rb");
        if (!fp) { set_error("cannot open state file %s", paths[p]); rc = SPHB_E_ARG; break; }
        PartHeader h;
        sphb_params prm;
        if (fread(&h, sizeof h, 1, fp) != 1 || memcmp(h.magic, "SPHB200", 8) != 0 || h.version != 2 ||
            h.params_bytes != sizeof(sphb_params) || fread(&prm, sizeof prm, 1, fp) != 1) {
            set_error("%s is not a version-2 (slab part) state file", paths[p]);
            rc = SPHB_E_ARG;
        } else if (prm.R != c->prm.R || prm.H != c->prm.H || prm.cell_length != c->prm.cell_length || prm.x_min != c->prm.x_min ||
                   prm.x_max != c->prm.x_max || prm.y_min != c->prm.y_min || prm.y_max != c->prm.y_max) {
            set_error("%s was written for another scene (R, H or the tank differ from this context's parameters)", paths[p]);
            rc = SPHB_E_ARG;
        }
        if (!rc) {
            if (p == 0) steps = h.steps;
            else if (h.steps != steps) { set_error("%s is from step %llu, the first part from step %llu", paths[p], h.steps, steps); rc = SPHB_E_ARG; }
            has_accel = has_accel && h.has_accel != 0;
        }
        const size_t np_ = h.n_fluid;
        uint32_t *pid = nullptr; sphb_particle *pf = nullptr; float *pdu = nullptr, *pdv = nullptr;
        if (!rc && np_ > 0) {
            pid = static_cast<uint32_t *>(malloc(4 * np_)); pf = static_cast<sphb_particle *>(malloc(sizeof(sphb_particle) * np_));
            pdu = static_cast<float *>(malloc(4 * np_)); pdv = static_cast<float *>(malloc(4 * np_));
            if (!pid || !pf || !pdu || !pdv) rc = SPHB_E_NOMEM;
            else if (fread(pid, 4, np_, fp) != np_ || fread(pf, sizeof *pf, np_, fp) != np_ || fread(pdu, 4, np_, fp) != np_ ||
                     fread(pdv, 4, np_, fp) != np_) { set_error("%s is truncated", paths[p]); rc = SPHB_E_ARG; }
        }
        if (!rc && p == 0 && h.n_boundary > 0) {
            nb = h.n_boundary;
            b = static_cast<sphb_particle *>(malloc(sizeof(sphb_particle) * nb));
            if (!b) rc = SPHB_E_NOMEM;
            else if (fread(b, sizeof *b, nb, fp) != nb) { set_error("%s is truncated", paths[p]); rc = SPHB_E_ARG; }
        }
        fclose(fp);
        // keep what falls into this rank's columns (:112 decides the column, as everywhere)
        for (size_t i = 0; i < np_ && !rc; i++) {
            const int col = sphb_column_of(&c->prm, pf[i].x);
            if (col < m.col_lo || col >= m.col_hi) continue;
            if (n == cap) {
                cap = cap ? cap * 2 : (np_ > 1024 ? np_ : 1024);
                f = static_cast<sphb_particle *>(realloc(f, sizeof(sphb_particle) * cap)); ids = static_cast<uint32_t *>(realloc(ids, 4 * cap));
                du = static_cast<float *>(realloc(du, 4 * cap)); dv = static_cast<float *>(realloc(dv, 4 * cap));
                if (!f || !ids || !du || !dv) { rc = SPHB_E_NOMEM; break; }
            }
            f[n] = pf[i]; ids[n] = pid[i]; du[n] = pdu[i]; dv[n] = pdv[i];
            n++;
        }
        free(pid); free(pf); free(pdu); free(pdv);
    }
    if (!rc && n > 2000000000ULL) { set_error("too many particles for one rank"); rc = SPHB_E_ARG; }
    if (!rc) rc = sphb_mg_upload(c, f, ids, 0, (int)n, b, (int)nb);
    if (!rc && has_accel) rc = sphb_mg_upload_accel(c, du, dv);
    if (!rc) c->steps = steps;
    free(f); free(ids); free(du); free(dv); free(b);
    return rc;
}

int sphb_mg_rebalance(sphb_ctx *c, int min_width, double column_cost, double min_imbalance, int *changed_out)
{
    SPHB_ENTER(c);
    Exchange x;
    x.nccl = true;
    return rebalance_impl(c, min_width, column_cost, min_imbalance, x, changed_out);
}

int sphb_mg_rebalance_host(sphb_ctx *c, int min_width, double column_cost, double min_imbalance,
                           sphb_mg_allreduce_u64_fn allreduce, sphb_mg_alltoallv_fn alltoallv, void *user, int *changed_out)
{
    SPHB_ENTER(c);
    if (!allreduce || !alltoallv) { set_error("sphb_mg_rebalance_host: both callbacks are needed"); return SPHB_E_ARG; }
    Exchange x;
    x.nccl = false; x.ar = allreduce; x.a2a = alltoallv; x.user = user;
    return rebalance_impl(c, min_width, column_cost, min_imbalance, x, changed_out);
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

// kernels_aux.cu — the callers either side of the hot path (SURVEY.md §8f "next" rows):
//   k_render   draw_metaballs (pi_sph_fluid.c:380-411) -> SSD1306 page-packed 1 KiB frame
//   k_stats    the per-step statistics scans (:656-675) + conservation sums, as reductions
//   k_refresh / k_tait  helpers of the compat tier (sphb_compat.cu)
#include "sphb_internal.cuh"

namespace sphb {

// ------------------------------------------------------------------------------ render

// One thread per output byte: byte (page, j) packs rows i = 8*page .. 8*page+7 of column j,
// bit i%8 (:407-408).  Each pixel is a find_neighbors query of the pixel-centre
// pseudo-particle against the fluid grid (:391) followed by the metaball sum with its early
// exit (:394-404).  8192 queries: latency-bound and tiny, so no staging.
__global__ void __launch_bounds__(128)
k_render(const Consts k, const float2 *__restrict__ pixels, const float2 *__restrict__ pos,
         const uint32_t *__restrict__ start, const float W_px, unsigned char *__restrict__ frame)
{
    const int b = blockIdx.x * 128 + threadIdx.x;     // 0..1023
    if (b >= 1024) return;
    const int page = b >> 7, j = b & 127;
    unsigned int bits = 0;
    for (int bit = 0; bit < 8; bit++) {
        const int i = page * 8 + bit;
        const float2 px = pixels[i * 128 + j];
        // :134-135 — the reference does not clamp the centre cell; pixel centres are inside
        // the tank, so the clamp in cell_of never triggers here
        int row, col;
        bool esc;
        cell_of(k, px.x, px.y, row, col, esc);
        const int c0 = col > 0 ? col - 1 : 0, c1 = col < k.cols - 1 ? col + 1 : k.cols - 1;
        float cond = 0.0f;     // :394
        for (int rr = row - 1; rr <= row + 1 && cond < 1.0f; rr++) {
            if (rr < 0 || rr >= k.rows) continue;
            const int a = (int)start[rr * k.cols + c0], e = (int)start[rr * k.cols + c1 + 1];
            for (int q = a; q < e; q++) {
                const float2 pj = pos[q];
                const float d2 = dist2(f_sub(px.x, pj.x), f_sub(px.y, pj.y));
                if (within_support(k, d2)) {
                    cond = f_add(cond, f_div(W_strict(k, d2), W_px));     // :400-401
                    if (cond >= 1.0f) break;                              // :403
                }
            }
        }
        if (cond >= 1.0f) bits |= 1u << bit;     // :407
    }
    frame[b] = (unsigned char)bits;
}

int launch_render(cudaStream_t st, const Consts &k, const ParticleSet &fluid, const float2 *pixels, float W_px,
                  unsigned char *frame)
{
    k_render<<<8, 128, 0, st>>>(k, pixels, fluid.pos[fluid.pc], fluid.cell_start, W_px, frame);
    return 1;
}

// ------------------------------------------------------------------------------ splat (large N)

// draw_metaballs (:380-411) is meaningless once a pixel (WIDTH/128 = 31 mm) is much wider than the kernel
// support: its condition divides by W(px/2), i.e. W far outside 2H, where the reference's W (no cut-off, :45-50)
// is a growing polynomial.  For such scenes the frame is a SPLAT: the particles inside each pixel are counted
// here (shared-memory histogram per CTA, one global atomic per touched pixel), and a pixel is lit when the
// fluid volume inside it, count * V, covers at least half of the pixel (sphb_splat_frame, host side).
// Pixel (i, j) covers x in [j, j+1) * WIDTH/128 and y in (64-i-1, 64-i] * HEIGHT/64 — the cells whose centres
// the reference uses as pixel pseudo-particles (:573-574).
__global__ void __launch_bounds__(256)
k_pixel_counts(const Consts k, const Count cnt, const float2 *__restrict__ pos, const uint32_t *__restrict__ cellkey,
               const float sx, const float sy, unsigned int *__restrict__ counts)
{
    __shared__ unsigned int s_c[64 * 128];
    for (int i = threadIdx.x; i < 64 * 128; i += 256) s_c[i] = 0u;
    __syncthreads();
    const int n = count_of(cnt);
    for (int s = blockIdx.x * 256 + threadIdx.x; s < n; s += gridDim.x * 256) {
        if (cellkey && !owned_col(k, (int)(cellkey[s] & 0xffffu))) continue;      // slabs: ghosts are the neighbour's
        const float2 p = pos[s];
        int j = (int)floorf((p.x - k.x_min) * sx), i = 63 - (int)floorf((p.y - k.y_min) * sy);
        j = j < 0 ? 0 : (j > 127 ? 127 : j);
        i = i < 0 ? 0 : (i > 63 ? 63 : i);
        atomicAdd(&s_c[i * 128 + j], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 128; i += 256)
        if (s_c[i]) atomicAdd(&counts[i], s_c[i]);
}

int launch_pixel_counts(cudaStream_t st, const Consts &k, const ParticleSet &f, float width, float height, unsigned int *counts)
{
    if (f.n == 0) return 0;
    int grid = (f.n + 256 * 64 - 1) / (256 * 64);
    grid = grid < 1 ? 1 : (grid > 148 * 4 ? 148 * 4 : grid);
    k_pixel_counts<<<grid, 256, 0, st>>>(k, f.cur(), f.pos[f.pc], (f.windowed && f.sorted) ? f.cellkey : nullptr,
                                         128.0f / width, 64.0f / height, counts);
    return 1;
}

// ------------------------------------------------------------------------------ stats

__device__ __forceinline__ unsigned int float_order_key(float f)
{
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // monotone map float -> uint
}

// out_d: [0] sum m, [1] sum m*u, [2] sum m*v, [3] 0.5*sum m*(u^2+v^2)
// out_u: [0] max speed (float bits, >= 0), [1] key(max rho), [2] key(min rho) (stored inverted),
//        [3] rho of the particle with the highest original index (what :657-659 reports),
//        [4] particles counted (owned by this rank)
__global__ void __launch_bounds__(256)
k_stats(const Consts k, const Count cnt, const uint32_t last_id, const float2 *__restrict__ vel,
        const float2 *__restrict__ rho_prr, const float *__restrict__ mass, const float uniform_mass,
        const uint32_t *__restrict__ id, const uint32_t *__restrict__ cellkey, double *__restrict__ out_d,
        unsigned int *__restrict__ out_u, const DeviceCounters *__restrict__ ctr,
        const unsigned int *__restrict__ flags)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // counters of the build / slab kernels: [5] escaped, [6] max cell population, [7] lost, [8] overflow
        out_u[5] = ctr->n_escaped;
        out_u[6] = ctr->max_cell_count;
        out_u[7] = flags ? flags[0] : 0u;
        out_u[8] = flags ? flags[1] : 0u;
    }
    double m_sum = 0, mx = 0, my = 0, ke = 0;
    float vmax = 0.0f;
    unsigned int rmax = 0u, rmin_inv = 0u, owned = 0u;
    const int n = count_of(cnt);
    for (int s = blockIdx.x * 256 + threadIdx.x; s < n; s += gridDim.x * 256) {
        // slabs: ghost slots belong to (and are counted by) the neighbouring rank
        if (cellkey && !owned_col(k, (int)(cellkey[s] & 0xffffu))) continue;
        ++owned;
        const float2 v = vel[s];
        const float rho = rho_prr[s].x;
        const double m = mass ? (double)mass[s] : (double)uniform_mass;
        m_sum += m;
        mx += m * (double)v.x;
        my += m * (double)v.y;
        ke += 0.5 * m * ((double)v.x * v.x + (double)v.y * v.y);
        const float sp = f_sqrt(f_add(f_mul(v.x, v.x), f_mul(v.y, v.y)));     // :669
        vmax = sp > vmax ? sp : vmax;
        const unsigned int key = float_order_key(rho);
        rmax = key > rmax ? key : rmax;
        rmin_inv = ~key > rmin_inv ? ~key : rmin_inv;
        if (id[s] == last_id) out_u[3] = __float_as_uint(rho);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        m_sum += __shfl_xor_sync(0xffffffffu, m_sum, d);
        mx += __shfl_xor_sync(0xffffffffu, mx, d);
        my += __shfl_xor_sync(0xffffffffu, my, d);
        ke += __shfl_xor_sync(0xffffffffu, ke, d);
        const float ov = __shfl_xor_sync(0xffffffffu, vmax, d);
        vmax = ov > vmax ? ov : vmax;
        const unsigned int o1 = __shfl_xor_sync(0xffffffffu, rmax, d);
        rmax = o1 > rmax ? o1 : rmax;
        const unsigned int o2 = __shfl_xor_sync(0xffffffffu, rmin_inv, d);
        rmin_inv = o2 > rmin_inv ? o2 : rmin_inv;
        owned += __shfl_xor_sync(0xffffffffu, owned, d);
    }
    // block level, then ONE set of atomics per CTA: same-address double atomics serialise in L2,
    // and with one set per warp they — not the reads — set the kernel's duration
    __shared__ double s_d[8][4];
    __shared__ unsigned int s_u[8][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        s_d[warp][0] = m_sum; s_d[warp][1] = mx; s_d[warp][2] = my; s_d[warp][3] = ke;
        s_u[warp][0] = __float_as_uint(vmax); s_u[warp][1] = rmax; s_u[warp][2] = rmin_inv; s_u[warp][3] = owned;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double acc = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) acc += s_d[w][threadIdx.x];
        atomicAdd(&out_d[threadIdx.x], acc);
    } else if (threadIdx.x < 8) {
        const int j = threadIdx.x - 4;
        unsigned int acc = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) acc = j == 3 ? acc + s_u[w][j] : (s_u[w][j] > acc ? s_u[w][j] : acc);
        if (j == 3) { if (acc) atomicAdd(&out_u[4], acc); }
        else atomicMax(&out_u[j], acc);
    }
}

int launch_stats(cudaStream_t st, const Consts &k, const ParticleSet &f, double *out_d, float *out_u,
                 const DeviceCounters *ctr, const unsigned int *flags)
{
    if (f.n == 0) return 0;
    int grid = (f.n + 255) / 256;
    if (grid > 148 * 4) grid = 148 * 4;
    const uint32_t last_id = f.windowed ? 0xffffffffu : (uint32_t)(f.n - 1);
    k_stats<<<grid, 256, 0, st>>>(k, f.cur(), last_id, f.vel[f.vc], f.rho_prr, f.uniform_mass ? nullptr : f.mass[f.mc],
                                  f.uniform_mass_value, f.id[f.ic],
                                  // an unsorted slab set (just uploaded or re-cut) holds owned particles only: no ghost slots yet
                                  (f.windowed && f.sorted) ? f.cellkey : nullptr, out_d,
                                  reinterpret_cast<unsigned int *>(out_u), ctr, flags);
    return 1;
}

// ------------------------------------------------------------------------------ compat helpers

// Re-read every field of an already-sorted set from a fresh AoS copy (original order),
// keeping the permutation and the grid.  *moved is raised if a position differs from the
// one the grid was built for (stale context -> callers disable tile staging).
__global__ void __launch_bounds__(kStreamThreads)
k_refresh(const int n, const float *__restrict__ aos, const uint32_t *__restrict__ id, float2 *__restrict__ pos,
          float2 *__restrict__ vel, float *__restrict__ mass, float *__restrict__ aux,
          float2 *__restrict__ rho_prr, float *__restrict__ p, unsigned int *__restrict__ moved)
{
    const int s = blockIdx.x * kStreamThreads + threadIdx.x;
    if (s >= n) return;
    const float *r = aos + (size_t)id[s] * 7;
    const float2 old = pos[s];
    const float2 now = make_float2(r[0], r[1]);
    if (__float_as_uint(old.x) != __float_as_uint(now.x) || __float_as_uint(old.y) != __float_as_uint(now.y))
        atomicAdd(moved, 1u);
    pos[s] = now;
    vel[s] = make_float2(r[2], r[3]);
    if (mass) mass[s] = r[4];
    if (aux) aux[s] = r[5];
    if (rho_prr) rho_prr[s] = make_float2(r[5], p_over_rho2(r[6], r[5]));     // :321
    if (p) p[s] = r[6];
}

int launch_refresh(cudaStream_t st, const sphb_particle *aos, ParticleSet &ps, unsigned int *moved)
{
    ps.lists_valid = false;
    if (ps.n == 0) return 0;
    const int grid = (ps.n + kStreamThreads - 1) / kStreamThreads;
    k_refresh<<<grid, kStreamThreads, 0, st>>>(ps.n, reinterpret_cast<const float *>(aos), ps.id[ps.ic],
                                               ps.pos[ps.pc], ps.vel[ps.vc], ps.mass[0] ? ps.mass[ps.mc] : nullptr,
                                               ps.aux[0] ? ps.aux[ps.xc] : nullptr, ps.rho_prr, ps.p, moved);
    return 1;
}

// calculate_particle_pressure (:294-301) on an AoS array in place
__global__ void __launch_bounds__(kStreamThreads)
k_tait_aos(const Consts k, const int n, float *__restrict__ aos)
{
    const int i = blockIdx.x * kStreamThreads + threadIdx.x;
    if (i >= n) return;
    float *r = aos + (size_t)i * 7;
    r[6] = tait_pressure(k, r[5]);
}

int launch_tait_aos(cudaStream_t st, const Consts &k, int n, sphb_particle *aos)
{
    if (n == 0) return 0;
    k_tait_aos<<<(n + kStreamThreads - 1) / kStreamThreads, kStreamThreads, 0, st>>>(k, n, reinterpret_cast<float *>(aos));
    return 1;
}

}  // namespace sphb

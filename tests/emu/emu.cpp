// emu.cpp — HOST EMULATION of the device arithmetic, test infrastructure only.
//
// Compiles the exact expressions the CUDA kernels use (pi_sph_fluid_b200/csrc/sph_math.cuh,
// sph_consts.h) for the host and runs them with the kernels' visiting order: particles sorted
// by (cell, original index), three contiguous runs per neighbour row, sequential sums.  It
// lets the CPU-only test suite check the formulas and the run logic against the oracle
// without a GPU.  It is NOT a fallback: nothing in the product links or loads this file.
// Build: g++ -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (tests/test_device_math.py).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../pi_sph_fluid_b200/csrc/sph_consts.h"

using namespace sphb;

namespace {

struct Sorted {
    std::vector<int> order;        // sorted slot -> original index
    std::vector<int> start;        // cell -> first slot, ncells+1 entries
};

Sorted sort_by_cell(const Consts &k, const sphb_particle *p, int n)
{
    Sorted s;
    std::vector<int> cell(n);
    for (int i = 0; i < n; i++) {
        int r, c; bool e;
        cell_of(k, p[i].x, p[i].y, r, c, e);
        cell[i] = r * k.cols + c;
    }
    s.order.resize(n);
    for (int i = 0; i < n; i++) s.order[i] = i;
    std::stable_sort(s.order.begin(), s.order.end(), [&](int a, int b) { return cell[a] < cell[b]; });
    s.start.assign(k.ncells + 1, 0);
    for (int i = 0; i < n; i++) s.start[cell[i] + 1]++;
    for (int c = 0; c < k.ncells; c++) s.start[c + 1] += s.start[c];
    return s;
}

template <class F>
void for_each_candidate(const Consts &k, float x, float y, const Sorted &sb, F &&f)
{
    int row, col; bool e;
    cell_of(k, x, y, row, col, e);
    const int c0 = col > 0 ? col - 1 : 0, c1 = col < k.cols - 1 ? col + 1 : k.cols - 1;
    for (int rr = row - 1; rr <= row + 1; rr++) {
        if (rr < 0 || rr >= k.rows) continue;
        for (int q = sb.start[rr * k.cols + c0]; q < sb.start[rr * k.cols + c1 + 1]; q++) f(sb.order[q]);
    }
}

}  // namespace

extern "C" {

void emu_cell_ids(const sphb_params *prm, const sphb_particle *p, int n, int *out)
{
    const Consts k = make_consts(*prm, 0.0f);
    for (int i = 0; i < n; i++) {
        int r, c; bool e;
        cell_of(k, p[i].x, p[i].y, r, c, e);
        out[i] = r * k.cols + c;
    }
}

// k_pseudomass
void emu_pseudomass(const sphb_params *prm, sphb_particle *b, int nb)
{
    const Consts k = make_consts(*prm, 0.0f);
    const Sorted sb = sort_by_cell(k, b, nb);
    std::vector<float> psi(nb);
    for (int i = 0; i < nb; i++) {
        float recip = 0.0f;
        for_each_candidate(k, b[i].x, b[i].y, sb, [&](int j) {
            const float d2 = dist2(f_sub(b[i].x, b[j].x), f_sub(b[i].y, b[j].y));
            if (within_support(k, d2) && j != i) recip = f_add(recip, W_strict(k, d2));
        });
        psi[i] = f_div(b[i].rho, recip);
    }
    for (int i = 0; i < nb; i++) b[i].m = psi[i];
}

// k_density + k_force on the state as given; returns neighbour lists when asked
void emu_compute_accel(const sphb_params *prm, sphb_particle *f, int n, const sphb_particle *b, int nb,
                       float gx, float gy, float *du, float *dv, int *nb_counts, int *nb_lists, int cap)
{
    const Consts k = make_consts(*prm, n > 0 ? f[0].m : 0.0f);
    const Sorted sf = sort_by_cell(k, f, n);
    const Sorted sb = sort_by_cell(k, b, nb);
    std::vector<float> prr(n);
    for (int i = 0; i < n; i++) {
        float sum_ff = 0.0f, sum_fb = 0.0f;
        int cnt = 0;
        for_each_candidate(k, f[i].x, f[i].y, sf, [&](int j) {
            const float d2 = dist2(f_sub(f[i].x, f[j].x), f_sub(f[i].y, f[j].y));
            if (within_support(k, d2) && j != i) {
                sum_ff = f_add(sum_ff, f_mul(f[j].m, W_strict(k, d2)));
                if (nb_lists && cnt < cap) nb_lists[(size_t)i * cap + cnt] = j;
                cnt++;
            }
        });
        if (nb_counts) nb_counts[i] = cnt;
        if (nb > 0)
            for_each_candidate(k, f[i].x, f[i].y, sb, [&](int j) {
                const float d2 = dist2(f_sub(f[i].x, b[j].x), f_sub(f[i].y, b[j].y));
                if (within_support(k, d2)) sum_fb = f_add(sum_fb, f_mul(b[j].m, W_strict(k, d2)));
            });
        const float rho = f_add(f_add(f_mul(f[i].m, k.nf), sum_ff), sum_fb);
        f[i].rho = rho;
        f[i].p = tait_pressure(k, rho);
        prr[i] = p_over_rho2(f[i].p, rho);
    }
    const bool fast = prm->fast_force != 0;
    for (int i = 0; i < n; i++) {
        float sx = 0, sy = 0, bx = 0, by = 0;
        for_each_candidate(k, f[i].x, f[i].y, sf, [&](int j) {
            const float dx = f_sub(f[i].x, f[j].x), dy = f_sub(f[i].y, f[j].y);
            const float d2 = dist2(dx, dy);
            if (within_support(k, d2) && j != i) {
                if (!fast) {      // k_force MODE 1 / 2: the reference's arithmetic (:317-337, :52-62, :219-228)
                    const float xu = f_add(f_mul(dx, f_sub(f[i].u, f[j].u)), f_mul(dy, f_sub(f[i].v, f[j].v)));
                    const PairStrict o = force_pair_strict<true, false, false>(k, dx, dy, d2, xu, prr[i], prr[j], f[i].rho,
                                                                               f[j].rho, f[j].m);
                    sx = f_add(sx, o.tx);
                    sy = f_add(sy, o.ty);
                    return;
                }
                const float xu = dx * (f[i].u - f[j].u) + dy * (f[i].v - f[j].v);
                const float tg = f[j].m * force_pair(k, d2, xu, prr[i] + prr[j], f[i].rho + f[j].rho);
                sx += tg * dx;
                sy += tg * dy;
            }
        });
        if (nb > 0)
            for_each_candidate(k, f[i].x, f[i].y, sb, [&](int j) {
                const float dx = f_sub(f[i].x, b[j].x), dy = f_sub(f[i].y, b[j].y);
                const float d2 = dist2(dx, dy);
                if (within_support(k, d2)) {
                    if (!fast) {
                        const float xu = f_add(f_mul(dx, f_sub(f[i].u, b[j].u)), f_mul(dy, f_sub(f[i].v, b[j].v)));
                        const PairStrict o = force_pair_strict<false, false, false>(k, dx, dy, d2, xu, prr[i], 0.0f, f[i].rho,
                                                                                    0.0f, b[j].m);
                        bx = f_add(bx, o.tx);
                        by = f_add(by, o.ty);
                        return;
                    }
                    const float xu = dx * (f[i].u - b[j].u) + dy * (f[i].v - b[j].v);
                    const float tg = b[j].m * force_pair(k, d2, xu, prr[i], f[i].rho + f[i].rho);
                    bx += tg * dx;
                    by += tg * dy;
                }
            });
        if (!fast) {
            du[i] = f_sub(f_sub(gx, sx), bx);      // :370
            dv[i] = f_sub(f_sub(gy, sy), by);      // :371
        } else {
            du[i] = (gx - k.grad_c * sx) - k.grad_c * bx;
            dv[i] = (gy - k.grad_c * sy) - k.grad_c * by;
        }
    }
}

// k_advect_bin's arithmetic + the rest of a step
void emu_step(const sphb_params *prm, sphb_particle *f, int n, const sphb_particle *b, int nb, float gx,
              float gy, float *du, float *dv, int nsteps)
{
    const Consts k = make_consts(*prm, n > 0 ? f[0].m : 0.0f);
    for (int s = 0; s < nsteps; s++) {
        for (int i = 0; i < n; i++) {
            f[i].u = kick(k, f[i].u, du[i]);
            f[i].v = kick(k, f[i].v, dv[i]);
            f[i].x = drift(k, f[i].x, f[i].u);
            f[i].y = drift(k, f[i].y, f[i].v);
        }
        emu_compute_accel(prm, f, n, b, nb, gx, gy, du, dv, nullptr, nullptr, 0);
        for (int i = 0; i < n; i++) {
            f[i].u = kick(k, f[i].u, du[i]);
            f[i].v = kick(k, f[i].v, dv[i]);
        }
    }
}

// the pair term of the force pass for caller-given pairs (layout of sphb_probe_force_pair)
void emu_force_pair(const sphb_params *prm, int n, const float *in, int boundary, float *out)
{
    const Consts k = make_consts(*prm, prm->rho0 * prm->vol);
    for (int i = 0; i < n; i++) {
        const float *a = in + (size_t)i * 12;
        const float dx = f_sub(a[0], a[2]), dy = f_sub(a[1], a[3]);
        const float d2 = dist2(dx, dy);
        const float xu = f_add(f_mul(dx, f_sub(a[4], a[6])), f_mul(dy, f_sub(a[5], a[7])));
        const PairStrict o = boundary ? force_pair_strict<false, false, false>(k, dx, dy, d2, xu, a[9], a[11], a[8], a[10], k.mass)
                                      : force_pair_strict<true, false, false>(k, dx, dy, d2, xu, a[9], a[11], a[8], a[10], k.mass);
        out[2 * i] = o.tx;
        out[2 * i + 1] = o.ty;
    }
}

void emu_consts(const sphb_params *prm, float *out /*8*/)
{
    const Consts k = make_consts(*prm, 0.0f);
    out[0] = k.d2max; out[1] = k.nf; out[2] = k.grad_c; out[3] = k.inv_W_ref;
    out[4] = k.B; out[5] = k.eps_h2; out[6] = k.visc_cH; out[7] = k.support;
}

}  // extern "C"

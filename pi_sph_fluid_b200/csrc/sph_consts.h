// sph_consts.h — host-side derivation of the kernel constants (sphb::Consts) from
// sphb_params, with the reference's expression types (pi_sph_fluid.c:11-21, :46, :144, :297,
// :325, :332-334, :616).  Header-only so the host emulation test (tests/emu) uses the very
// same numbers as the library.
#pragma once

#include <math.h>
#include <string.h>

#include "../../include/sph_b200.h"
#include "sph_math.cuh"

namespace sphb {

// W(r,0,0,0) on the host with the reference's expression (:45-50); used only for constants.
inline float host_W(float H, float r)
{
    const float nf = (float)(7 / (4 * M_PI * (double)H * (double)H));
    const float q = r / H;
    const float a = 1 - 0.5f * q, b = 1 + 2 * q;
    return nf * powf(a, 4) * b;
}

// the same with powf(a, 4) as the multiply chain gcc emits under the reference's shipped -Ofast — the flavour
// the kernels follow; W_ref of the STRICT force pass
inline float host_W_chain(float H, float r)
{
    const float nf = (float)(7 / (4 * M_PI * (double)H * (double)H));
    const float q = r / H;
    const float a = 1 - 0.5f * q, b = 1 + 2 * q;
    const float a2 = a * a;
    return nf * (a2 * a2) * b;
}

// Is  q0 = r*y, q = q0 + fma(-q0, H, r)*y  (y = RN(1/H)) the correctly rounded r/H for every r?
// The expression is invariant under scaling r by 2, so all 2^23 mantissas of one binade decide it.
// (Serves any constant divisor: H and W_ref.)
inline bool markstein_div_exact(float H)
{
    static float cached_H[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    static bool cached[4] = {false, false, false, false};
    static int next = 0;
    for (int i = 0; i < 4; i++)
        if (H == cached_H[i]) return cached[i];
    const float y = 1.0f / H;
    bool ok = true;
    for (unsigned m = 0; m < (1u << 23) && ok; m++) {
        const unsigned bits = 0x3f800000u | m;
        float r;
        memcpy(&r, &bits, sizeof r);
        const float q0 = r * y;
        const float q = fmaf(fmaf(-q0, H, r), y, q0);
        ok = (q == r / H);
    }
    cached_H[next] = H;
    cached[next] = ok;
    next = (next + 1) & 3;
    return ok;
}

inline Consts make_consts(const sphb_params &p, float uniform_mass)
{
    Consts k;
    memset(&k, 0, sizeof k);
    k.x_min = p.x_min;
    k.y_min = p.y_min;
    k.cell = p.cell_length;
    k.rows = (int)((p.y_max - p.y_min) / p.cell_length) + 1;     // :93
    k.cols = (int)((p.x_max - p.x_min) / p.cell_length) + 1;     // :94
    k.ncells = k.rows * k.cols;
    k.gcols = k.cols;
    k.col_off = 0;
    k.own_lo = 0;
    k.own_hi = k.cols;
    k.H = p.H;
    k.inv_H = 1.0f / p.H;
    k.support = 2 * p.H;                                          // :144
    // largest float t with sqrtf(t) < 2*H: the test `sqrtf(d2) < 2*H` is then `d2 <= t`
    float t = k.support * k.support;
    while (sqrtf(t) < k.support) t = nextafterf(t, INFINITY);
    while (!(sqrtf(t) < k.support)) t = nextafterf(t, -INFINITY);
    k.d2max = t;
    const double Hd = (double)p.H;
    k.nf = (float)(7 / (4 * M_PI * Hd * Hd));                     // :46
    k.grad_c = (float)(-5.0 * (double)k.nf / (Hd * Hd));
    k.inv_W_ref = 1.0f / host_W(p.H, (float)(0.2 * Hd));          // :325
    k.a_c = (float)(-0.5 / Hd);
    k.b_c = (float)(2.0 / Hd);
    k.art_c = (float)(pow(0.1, 0.25) * (double)k.nf * (double)k.inv_W_ref);
    k.div_exact = markstein_div_exact(p.H) ? 1 : 0;
    k.W_ref = host_W_chain(p.H, (float)(0.2 * Hd));               // :325, W(0.2*H, 0, 0, 0)
    k.inv_W_ref_c = 1.0f / k.W_ref;
    k.wref_div_exact = markstein_div_exact(k.W_ref) ? 1 : 0;
    k.nf_m5 = k.nf * (-5);                                        // :56
    k.eps_h2_d = 0.01 * Hd * Hd;                                  // :332
    k.visc_c_d = -0.01 * (double)p.c0;                            // :334
    {
        int e;
        const double mant = frexp(fabs(k.visc_c_d), &e);
        k.visc_pow2 = (mant == 0.5) ? 1 : 0;
        k.visc_c_f = (float)k.visc_c_d;
    }
    {   // corner-cell culling radius: 2H plus 8 ulp of the largest coordinate (covers the rounding
        // of cell edges and offsets), squared, rounded up
        const float big = fmaxf(fmaxf(fabsf(p.x_min), fabsf(p.x_max)), fmaxf(fabsf(p.y_min), fabsf(p.y_max)));
        const float slack = 8.0f * (nextafterf(big, INFINITY) - big);
        const float rc = k.support + slack;
        k.cull2 = rc * rc * 1.00001f;
    }
    k.rho0 = p.rho0;
    k.inv_rho0 = 1.0f / p.rho0;
    k.B = p.c0 * p.c0 * p.rho0 / 7;                               // :297
    k.eps_h2 = (float)(0.01 * Hd * Hd);                           // :332
    k.visc_cH = (float)(-0.01 * (double)p.c0 * Hd);               // :332, :334
    k.visc2_cH = (float)(-0.02 * (double)p.c0 * Hd);
    k.mass = uniform_mass;
    k.dt = p.dt;
    k.half_dt = 0.5 * (double)p.dt;                               // :616
    return k;
}

// Restrict the grid to the window of global columns [win_lo, win_hi) of which [own_lo, own_hi)
// are owned by this rank (multi-GPU slabs).
inline void set_window(Consts &k, int win_lo, int win_hi, int own_lo, int own_hi)
{
    k.col_off = win_lo;
    k.cols = win_hi - win_lo;
    k.ncells = k.rows * k.cols;
    k.own_lo = own_lo - win_lo;
    k.own_hi = own_hi - win_lo;
}

}  // namespace sphb

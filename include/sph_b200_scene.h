/* sph_b200_scene.h — host-side scene builders of libsphb200.so (plain C, no GPU work).
 *
 * sphb_scene_*_drop / *_boundary reproduce the reference's own initial condition
 * (pi_sph_fluid.c:484-540) including its float-accumulated lattice `x_0 += R`, so the same
 * byte-identical arrays can be fed to the reference and to this library.  The block scenes
 * (dam break, filled tank) and the synthetic tilt trace are NOT in the reference; they are
 * the builder-defined benchmark scenes of SURVEY.md §8(d), on the same lattice idiom.
 */
#ifndef SPH_B200_SCENE_H
#define SPH_B200_SCENE_H

#include "sph_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

/* :485-488 / :496-506 — lattice points inside the disc of radius 0.70 at the tank centre */
int sphb_scene_count_drop(const sphb_params *prm);
int sphb_scene_fill_drop(const sphb_params *prm, sphb_particle *fluid_out);

/* lattice points with x0 <= x < x1 and y0 <= y < y1 (dam break / filled tank) */
int sphb_scene_count_block(const sphb_params *prm, float x0, float x1, float y0, float y1);
int sphb_scene_fill_block(const sphb_params *prm, float x0, float x1, float y0, float y1,
                          sphb_particle *fluid_out);

/* :514-516 / :522-540 — one layer of wall particles at spacing R on the four walls */
int sphb_scene_count_boundary(const sphb_params *prm);
int sphb_scene_fill_boundary(const sphb_params *prm, sphb_particle *boundary_out);

/* Synthetic MPU6050 trace for a tank rocking by +-amplitude_deg with the given period:
 * raw counts ax = round(16384*cos(theta)), ay = round(16384*sin(theta)),
 * theta(t) = amplitude*sin(2*pi*step/period_steps), sampled every hold_steps steps and held
 * (the reference polls at 10 Hz, :454-463), mapped through :439-440.  Writes nsteps (gx,gy)
 * pairs. */
int sphb_gravity_trace_tilt(const sphb_params *prm, float amplitude_deg, int period_steps,
                            int hold_steps, int nsteps, float *gravity_xy_out);

/* slab-wise construction of the block scene (multi-GPU): the particles whose x lies in cell
 * columns [col_lo, col_hi) — a contiguous index range [*id_base, *id_base + n) of the full scene.
 * out may be NULL to count only. */
int sphb_scene_fill_block_slab(const sphb_params *prm, float x0, float x1, float y0, float y1, int col_lo,
                               int col_hi, sphb_particle *out, unsigned int *id_base);
/* adds the block scene's per-column particle counts to hist[cols] */
int sphb_scene_block_column_hist(const sphb_params *prm, float x0, float x1, float y0, float y1,
                                 unsigned long long *hist);

/* spacing R such that a block of the given area holds about n_target lattice particles */
float sphb_spacing_for_count(double area, double n_target);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif

/* libus3d — profiling and tuning hooks.  NOT part of the drop-in surface (include/us3d.h): nothing in the reference binds
 * these; bench.py's roofline leg (per-launch CUDA events) and the tuning scripts under scripts/ use them.
 * Every exported symbol of libus3d.so is declared either in us3d.h or here (tests/test_abi.py checks both directions). */
#ifndef US3D_DEBUG_H
#define US3D_DEBUG_H

#ifdef __cplusplus
extern "C" {
#endif

/* Programmatic dependent launch of the convolution / BatchNorm kernels (a kernel's prologue overlaps its predecessor's tail):
 * on by default, US3D_PDL=0 in the environment or on = 0 here launches the same kernels without the attribute. */
void us3d_debug_set_pdl(int on);

/* Per-launch timing of the convolution kernels: between start and stop every convolution entry point brackets its launch
 * with CUDA events recorded on the launch stream.  stop() synchronises the events and fills, per launch in launch order,
 * meta[7] = (kind 0 fwd/dgrad | 1 wgrad as tagged, n_in, n_rows, kvol, cin, cout, tag) and ms; returns the launch count. */
void us3d_debug_profile_start(void);
void us3d_debug_profile_tag(int kind);
int us3d_debug_profile_stop(int *meta, float *ms, int cap);

/* us3d_spconv_gather_mt: per-CTA counters of the MMA-issuing thread, 8 x int64 per CTA (total cycles, cycles waiting for the
 * accumulator / a weight slab / a gathered tile, tiles multiplied, slabs consumed, cycles until the first tile landed);
 * buf = NULL switches the counters off. */
void us3d_debug_set_prof(void *buf);

/* us3d_spconv_wgrad_planes: per-CTA cycle counters, 16 x int64 per CTA: [0] producer loop, [1] / [2] of it waiting for a free dY /
 * X ring slot, [3] epilogue; [4] MMA warp's loop, [5] / [6] of it waiting for dY / gathered X, [7] fence + issue + commit,
 * [8] ring slots consumed.  buf = NULL switches the profiling instantiation off. */
void us3d_debug_set_prof_wgrad(void *buf);

/* Overrides of the launcher's choices for us3d_spconv_gather_mt (0 = launcher's choice): ring slots, producer completion
 * (1..3 = cp.async.wait_group look-ahead, 9 = cp.async.mbarrier.arrive.noinc), tiles per weight slab, fused [W_hi | W_lo]
 * operand (1 = off, 2 = on where eligible). */
void us3d_debug_set_tuning4(int a_slots, int lag, int T, int fuse);
void us3d_debug_set_tuning(int a_slots, int lag, int T);
/* accumulator sets of us3d_spconv_gather_mt: 0 = launcher's choice, 1 = one, 2 = two (tile groups shrink to fit the 512 TMEM columns) */
void us3d_debug_set_tuning_acc(int sets);

#ifdef __cplusplus
}
#endif
#endif

/*
 * pb200_lbl.h -- C ABI of the B200-native line-by-line opacity engine.
 *
 * This is the drop-in boundary for the one hot path of pcubillos/pyratbay that the
 * engine replaces (paths relative to the reference tree):
 *
 *   pb200_voigt_grid            <->  lib.vprofile.grid            src_c/vprofile.c:42-114
 *                                    called at pyratbay/pyrat/voigt.py:145-149
 *   pb200_engine_* + pb200_extinction_batch
 *                               <->  lib._extcoeff.extinction     src_c/_extcoeff.c:87-345
 *                                    called at pyratbay/pyrat/extinction.py:197-208 under
 *                                    the fork pools of extinction.py:109-119 and
 *                                    line_by_line.py:231-246 (one call = all (T,p) units)
 *   pb200_interp_ec             <->  lib._extcoeff.interp_ec      src_c/_extcoeff.c:367-418
 *   pb200_interp_ec_per_mol     <->  lib._extcoeff.interp_ec_per_mol  :421-472
 *                                    called at pyratbay/opacity/line_sampling.py:379-384,451-456
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no Python, NumPy or torch types.
 *   - Unless a parameter is documented as a device pointer, pointers are HOST memory,
 *     C-contiguous, float64 / int64 exactly like the NumPy arrays the reference passes.
 *   - Every function returns 0 on success and a negative PB200_E* code on failure;
 *     pb200_last_error() returns a thread-local human-readable message.
 *   - There is no CPU fallback: every entry point fails with PB200_ENODEVICE when no
 *     CUDA device is usable.
 *   - The engine handle owns device copies of the static inputs (spectral grids, Voigt
 *     table, line list); a batch call takes only the per-(T,p) quantities.  This replaces
 *     the reference's fork()-per-CPU orchestration, which cannot coexist with CUDA.
 */
#ifndef PB200_LBL_H
#define PB200_LBL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_OK 0
#define PB200_EINVAL (-1)    /* bad argument / inconsistent sizes            */
#define PB200_ENODEVICE (-2) /* no usable CUDA device                        */
#define PB200_ECUDA (-3)     /* CUDA runtime error (message has the detail)  */
#define PB200_ESTATE (-4)    /* call sequence error (e.g. lines before grid) */
#define PB200_ENOMEM (-5)

typedef struct pb200_engine pb200_engine;

/* Library / device ------------------------------------------------------------------- */
const char *pb200_last_error(void);
const char *pb200_version(void);
/* Number of visible CUDA devices (0 when none; never fails). */
int pb200_device_count(void);

/* Voigt grid (stand-alone form of vprofile.grid) ----------------------------------------
 * psize[nlor*ndop] in/out: half-sizes; entries equal to 0 are aliased to the previous
 *   Doppler sample exactly as the reference does (vprofile.c:99-105).
 * pindex[nlor*ndop] out: start index of every profile.
 * profile[profile_len] out (host): concatenated profiles; profile_len must be at least
 *   sum(2*psize+1) over the computed entries.  Unused tail entries are left untouched.
 * dwn: fine wavenumber step (spec.ownstep).
 * Profiles with 2*size+1 > 99999 use point sampling (VOIGT_QUICK, voigt.h:124,276-279). */
int pb200_voigt_grid(int device, int nlor, int ndop, const double *lorentz,
                     const double *doppler, double dwn, int64_t *psize, int64_t *pindex,
                     double *profile, int64_t profile_len);

/* Engine life cycle ---------------------------------------------------------------------- */
int pb200_engine_create(int device, pb200_engine **out);
void pb200_engine_destroy(pb200_engine *e);

/* Spectral grids (spec.wn, spec.own, spec.odivisors; pyrat/spectrum.py:215-228). */
int pb200_engine_set_grid(pb200_engine *e, const double *wn, int64_t nwave,
                          const double *own, int64_t onwn, const int64_t *divisors,
                          int ndivs);

/* Voigt table computed on the device and kept there (Voigt.__init__, pyrat/voigt.py:105-149).
 * psize in/out and pindex out as in pb200_voigt_grid.  cutoff in cm-1 (<=0: none). */
int pb200_engine_build_voigt(pb200_engine *e, int nlor, int ndop, const double *lorentz,
                             const double *doppler, double dwn, int64_t *psize,
                             int64_t *pindex, double cutoff);
/* Alternative: adopt a host profile table computed elsewhere (exact ec.extinction
 * semantics for an arbitrary caller-provided table). */
int pb200_engine_set_voigt(pb200_engine *e, int nlor, int ndop, const double *lorentz,
                           const double *doppler, const int64_t *psize,
                           const int64_t *pindex, const double *profile,
                           int64_t profile_len, double cutoff);
int64_t pb200_engine_profile_len(const pb200_engine *e);
/* Copy the device Voigt table back to host (Voigt.profile). */
int pb200_engine_get_profile(pb200_engine *e, double *profile, int64_t profile_len);

/* Species / isotope static data (atm.mol_radius [cm], atm.mol_mass, lbl.iso_atm_index,
 * lbl.iso_mass, lbl.iso_ratio). */
int pb200_engine_set_species(pb200_engine *e, int nmol, const double *mol_radius,
                             const double *mol_mass, int niso, const int64_t *iso_imol,
                             const double *iso_mass, const double *iso_ratio);
/* Optional partition-function tables Z_i(T) [niso, ntemp] on a common temperature grid;
 * used when a batch call passes unit_isoz == NULL (piecewise-linear in T, the same
 * interpolant as line_by_line.py:156-158). */
int pb200_engine_set_partition(pb200_engine *e, int ntemp, const double *temp,
                               const double *z);

/* Line list (lbl.wn, lbl.elow, lbl.gf, lbl.isoid).  Requires set_grid first.  Lines must
 * be ascending in wavenumber within each isotope (the TLI guarantee, lread.py:187-203);
 * otherwise PB200_EINVAL.  Performs the (T,p)-independent pre-processing: window filter
 * (_extcoeff.c:215,239), nearest fine-grid index (:243-245) and the greedy co-add
 * grouping (:249-262). */
int pb200_engine_set_lines(pb200_engine *e, int64_t nlines, const double *wn,
                           const double *elow, const double *gf, const int64_t *iso_id);

/* Static facts after set_lines: [0]=lines in window, [1]=co-add groups, [2]=nadd
 * (lines absorbed into a preceding head line; _extcoeff.c:256). */
int pb200_engine_line_stats(const pb200_engine *e, int64_t stats[3]);

/* Batched extinction ------------------------------------------------------------------------
 * One call evaluates n_units independent (T,p) units; unit u is exactly one
 * ec.extinction(...) call of the reference with
 *     temp = unit_temp[u], moldensity = unit_density[u,:], isoz = unit_isoz[u,:]
 * and the engine's static data.  Rows: nrows = 1 if add else nextinct.
 *   unit_temp     [n_units]
 *   unit_density  [n_units, nmol]   (molecules cm-3)
 *   unit_isoz     [n_units, niso] or NULL (then set_partition tables are used)
 *   iso_iext      [niso]  output row per isotope, <0 skips the isotope (skip_mol)
 *   ethresh       line-strength threshold factor
 *   add           1: extinction coefficient (cm-1), 0: cross section per species (cm2)
 *   resolution    1: 2-point interpolation onto wn (utils.h:139-163), 0: resample (:119-135)
 *   out           [n_units, nrows, nwave]; OVERWRITTEN with what the reference would
 *                 leave in a zero-initialised `ext` (every reference call site zeroes it:
 *                 pyrat/extinction.py:195).
 *   counters      NULL or [n_units, 6] int64: nadd, nskip, neval (the reference's verbose
 *                 counters, _extcoeff.c:311-318), the dynamic-grid samples the reference
 *                 accumulates for this unit (sum of maxj-minj, :304-307), the profile
 *                 samples this engine gathers for it, and the bytes of distinct Voigt-table
 *                 samples the unit's lines can select (first-touch HBM traffic).
 * The *_host form copies inputs/outputs itself; the *_dev form takes `out` as a device
 * pointer (e.g. a torch tensor's data_ptr) and leaves the result on the device. */
int pb200_extinction_batch_host(pb200_engine *e, int n_units, const double *unit_temp,
                                const double *unit_density, const double *unit_isoz,
                                const int64_t *iso_iext, int nextinct, double ethresh,
                                int add, int resolution, double *out, int64_t *counters);
int pb200_extinction_batch_dev(pb200_engine *e, int n_units, const double *unit_temp,
                               const double *unit_density, const double *unit_isoz,
                               const int64_t *iso_iext, int nextinct, double ethresh,
                               int add, int resolution, double *out_dev,
                               int64_t *counters, void *cuda_stream);

/* Time (ms, CUDA events on the engine's stream) spent by the most recent batch call in:
 * [0] strengths kernel(s), [1] accumulate kernel(s), [2] H2D, [3] D2H, [4] whole call. */
int pb200_engine_last_timing(const pb200_engine *e, double ms[5]);
/* Kernel launches issued by this engine since creation. */
int64_t pb200_engine_launch_count(const pb200_engine *e);

/* (unit, isotope) pairs of the last batch that took the dense-convolution accumulate path
 * (isotopes whose co-add groups fill >= 25 % of the fine grid, constant-step output grids;
 * csrc/dense_kernels.cu).  0: every isotope went through the gather kernels.  Same results
 * either way (_extcoeff.c:229-309); diagnostic only. */
int64_t pb200_engine_dense_units(const pb200_engine *e);
/* Device time (ms) the dense-convolution kernels of the last batch took (part of the
 * accumulate time of pb200_engine_last_timing). */
double pb200_engine_dense_ms(const pb200_engine *e);
/* The engine's CUDA stream (cudaStream_t), so a caller can record its own events on it. */
void *pb200_engine_stream(const pb200_engine *e);

/* Device ceilings for roofline statements (bench.py only): fp64 FMA throughput in TFLOP/s
 * and read bandwidth (GB/s) of an L2-resident buffer of `mbytes` MiB; best of `reps`. */
int pb200_bench_fp64(int device, int reps, double *tflops);
int pb200_bench_l2(int device, int mbytes, int reps, double *gbs);

/* Exactness aids of the accumulate kernel (tests only).
 * pb200_nearest_thresholds (host, no device needed): thr[j] = smallest double v whose nearest
 * sample of the strictly increasing positive grid[n] is >= j (thr[0] = 0), the table that replaces
 * the nearest-index search of _extcoeff.c:278 (pyramidsearch, utils.h:44-72); returns
 * PB200_EINVAL when the grid is not strictly increasing (the engine then keeps the search).
 * pb200_selftest_exact (device): draws n operands; mismatches[0] counts quotients a/b for which
 * the FMA-only form used for idwn (_extcoeff.c:275) differs from the IEEE division,
 * mismatches[1] the widths for which the threshold table and the bisection disagree. */
int pb200_nearest_thresholds(const double *grid, int n, double *thr);
int pb200_selftest_exact(int device, int64_t n, uint64_t seed, const double *steps, int nsteps,
                         const double *grid, const double *thr, int ngrid,
                         uint64_t mismatches[2]);

/* Cross-section table interpolation in temperature ------------------------------------------
 * ext[nlayers,nwave] (or [nspec,nlayers,nwave] for per_mol) is ACCUMULATED (+=) like the
 * reference.  etable[nspec,ntemp,nlayers,nwave], ttable[ntemp], temperature[nlayers],
 * density[nlayers,nspec].  Host pointers; *_dev forms take etable/ext on the device. */
int pb200_interp_ec(int device, double *ext, const double *etable, const double *ttable,
                    const double *temperature, const double *density, int nspec,
                    int ntemp, int nlayers, int nwave, int lay1, int lay2);
int pb200_interp_ec_per_mol(int device, double *ext, const double *etable,
                            const double *ttable, const double *temperature,
                            const double *density, int nspec, int ntemp, int nlayers,
                            int nwave, int lay1, int lay2);
int pb200_interp_ec_dev(int device, double *ext_dev, const double *etable_dev,
                        const double *ttable, const double *temperature,
                        const double *density, int nspec, int ntemp, int nlayers,
                        int nwave, int lay1, int lay2, int per_mol, void *cuda_stream);

/* Device-resident table consumer ------------------------------------
 * A handle over a table [nspec, ntemp, nlayers, nwave] that already lives in HBM (borrowed
 * pointer: the caller keeps it alive), e.g. the rows pb200_extinction_batch_dev left there.
 * pb200_table_interp is interp_ec / interp_ec_per_mol (src_c/_extcoeff.c:367-472; call sites
 * pyratbay/opacity/line_sampling.py:366-391,440-463) without per-call allocation, lock or
 * stream synchronisation: the per-layer scalars travel through a ring of pinned staging slots
 * and the call returns once the kernel is queued on `cuda_stream` (sync != 0: waits).
 *   overwrite != 0: ext_dev is treated as zero on entry (the reference's callers pass a
 *                   zeroed array), so a persistent output buffer needs no memset;
 *   overwrite == 0: accumulate (+=) like the reference's C function. */
typedef struct pb200_table pb200_table;
int pb200_table_create(int device, const double *etable_dev, const double *ttable, int nspec,
                       int ntemp, int nlayers, int nwave, pb200_table **out);
void pb200_table_destroy(pb200_table *t);
int pb200_table_interp(pb200_table *t, const double *temperature, const double *density,
                       int lay1, int lay2, int per_mol, double *ext_dev, int overwrite,
                       void *cuda_stream, int sync);
int64_t pb200_table_launch_count(const pb200_table *t);

/* p/T re-gridding of a table on the device: pyratbay/tools/tools.py:1026-1107
 * (interpolate_opacity; call site opacity/line_sampling.py:243-250).  out[t, p, w] =
 * exp(lerp over T of lerp over log p of log table), the log of non-positive entries floored at
 * -230, or a plain copy when take_log == 0 (same grids).  The caller supplies the brackets:
 * output temperature i mixes table rows t_lo[i], t_hi[i] with weight t_f[i] on the upper one
 * (same for pressure; lo == hi with weight 0 for values outside the table or an axis that is
 * not resampled); wave_idx[nwave_out] selects the wavenumber samples (NULL: the first nwave_out).
 * accumulate != 0 adds to out (several files of one species).  Synchronises `cuda_stream`. */
int pb200_regrid_table_dev(int device, const double *table_dev, int ntemp, int nlayers, int nwave,
                           const int *t_lo, const int *t_hi, const double *t_f, int ntemp_out,
                           const int *p_lo, const int *p_hi, const double *p_f, int nlayers_out,
                           const int *wave_idx, int nwave_out, int take_log, double *out_dev,
                           int accumulate, void *cuda_stream);

/* Optical depth (next-tier row: the step after the extinction in Pyrat.run) ------------------
 * Replaces lib._trapezoid.plane_parallel_optical_depth (src_c/_trapezoid.c:147-211) and the
 * lib._trapezoid.optdepth loop of pyratbay/opacity/optic_depth.py:104-111 (transit geometry).
 *   transit     0: plane-parallel (emission/eclipse paths), 1: grazing rays (transit)
 *   depth       [nlayers, nwave] out (rows the reference leaves at zero are written as zero)
 *   ideep       [nwave] int32 out: layer where each channel reached maxdepth (or the bottom)
 *   extinction  [nlayers, nwave] (cm-1)
 *   geometry    plane-parallel: intervals [nlayers-1] (cm); transit: [nlayers, nlayers]
 *               zero-padded matrix whose row r holds raypath[r] (r - itop entries)
 * The *_dev form takes depth/ideep/extinction as device pointers (extinction straight from
 * pb200_extinction_batch_dev, no host round trip); geometry is always a host array. */
int pb200_optical_depth(int device, int transit, double *depth, int32_t *ideep,
                        const double *extinction, const double *geometry, double maxdepth,
                        int itop, int ibottom, int nlayers, int nwave);
int pb200_optical_depth_dev(int device, int transit, double *depth_dev, int32_t *ideep_dev,
                            const double *extinction_dev, const double *geometry,
                            double maxdepth, int itop, int ibottom, int nlayers, int nwave,
                            void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* PB200_LBL_H */

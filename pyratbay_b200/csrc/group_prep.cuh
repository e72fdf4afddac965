// group_prep.cuh -- the per-group part of _extcoeff.c:264-299 as device functions shared by the
// gather kernels (lbl_kernels.cu) and the dense-convolution kernel (dense_kernels.cu): both
// take every discrete decision (skip, Doppler profile, index ranges) with this one code.
#pragma once
#include "engine.cuh"

namespace pb200 {

// Accumulation modes of kernel 3.
//   kStrided    : constant-step output, samples gathered from the reference-layout table with
//                 stride ofactor*scale (generic; any scale)
//   kLinterp    : arbitrary output grid, two dynamic samples per output point (utils.h:139-163)
//   kTransposed : constant-step output, samples gathered from the output-stride copy of the
//                 table: consecutive output points read consecutive addresses (coalesced)
enum AccMode { kStrided = 0, kLinterp = 1, kTransposed = 2 };

struct Prep {
    double k;
    long long base;  // sample for coordinate x is table[base + mult*x]
    int lo, hi;      // coordinate range [lo, hi): dynamic index j, or output index m (kTransposed)
};

// Nearest Doppler sample of width v.  `s_dop` is the shared-memory copy of the threshold
// table (V.dop_thr) when the grid has one, of the grid itself otherwise.
__device__ __forceinline__ int doppler_index(const StaticView &V, const double *s_dop, double v) {
    return V.dop_thr ? nearest_index_thr(s_dop, V.ndop, v, V.dop_hi0, V.dop_inv_step)
                     : nearest_index_log(s_dop, V.ndop, v, V.dop_hi0, V.dop_inv_step);
}

__device__ __forceinline__ ProfileSlot load_slot(const ProfileSlot *p) {
    const int4 raw = __ldg(reinterpret_cast<const int4 *>(p));
    ProfileSlot s;
    s.base = (long long)(((unsigned long long)(unsigned)raw.y << 32) | (unsigned)raw.x);
    s.half = raw.z;
    s.rowlen = raw.w;
    return s;
}

// Dynamic-sample range [jlo, jhi) of a group with nearest fine index `iown` and dynamic index
// `idwn` whose profile has half-size `half` (_extcoeff.c:281-299), C integer semantics.
__device__ __forceinline__ void dynamic_range(const UnitParams &U, double cutoff, int half,
                                              int iown, int idwn, int *jlo_out, int *jhi_out) {
    const int sub = iown - idwn * U.ofactor;                                // :281
    int jlo = idwn - U.fd_ofactor.div_trunc(half - sub);                    // :286
    int jhi = idwn + U.fd_ofactor.div_trunc(half + sub);                    // :287
    if (jlo < 0) jlo = 0;
    if (jhi > U.dnwn) jhi = U.dnwn;
    if (cutoff > 0.0) {                                                     // :294-299
        const int lo_cut = (int)dsub((double)idwn, U.cut_steps);
        const int hi_cut = (int)dadd((double)idwn, U.cut_steps);
        if (lo_cut > jlo) jlo = lo_cut;
        if (hi_cut < jhi) jhi = hi_cut;
    }
    *jlo_out = jlo;
    *jhi_out = jhi;
}

// Output samples [mlo, mhi) that keep a dynamic sample of [jlo, jhi): outputs m with
// scale*m in [jlo, jhi), m < mcount (resample, utils.h:130-133).
__device__ __forceinline__ void output_range(const UnitParams &U, int jlo, int jhi, int *mlo,
                                             int *mhi) {
    *mlo = U.fd_scale.div_ceil(jlo);
    int hi = jhi > 0 ? U.fd_scale.div_ceil(jhi) : 0;
    if (hi > U.mcount) hi = U.mcount;
    *mhi = hi;
}

// Dynamic index of a line at wavenumber w (:275): trunc((w - own0)/dwnstep), the quotient
// exactly rounded.
__device__ __forceinline__ int dynamic_index(const StaticView &V, const UnitParams &U, double w) {
    return (int)quotient_rn(dsub(w, V.own0), U.dwnstep, U.inv_dwnstep);
}

// The per-group part of _extcoeff.c:264-299, identical integer/floating-point decisions.
template <int MODE>
__device__ __forceinline__ bool prepare_group(const StaticView &V, const UnitParams &U,
                                              const IsoUnit &I, const double *s_dop,
                                              double kthr, double cutoff, double w, int iown,
                                              double k, Prep *out, int *idop_out = nullptr) {
    if (k < kthr) return false;  // :265 skip weak lines
    k = dmul(k, I.dens);         // :271-272 (dens == 1 when add == 0)
    const int idwn = dynamic_index(V, U, w);                                         // :275
    const int idop = (V.ndop >= 2) ? doppler_index(V, s_dop, dmul(I.adop, w)) : 0;  // :278
    if (idop_out) *idop_out = idop;
    const int at = I.ilor * V.ndop + idop;
    const ProfileSlot ps = load_slot((MODE == kTransposed ? V.tslot : V.pslot) + at);
    const int half = ps.half;
    int jlo, jhi;
    dynamic_range(U, cutoff, half, iown, idwn, &jlo, &jhi);
    out->k = k;
    if (MODE == kTransposed) {
        int mlo, mhi;
        output_range(U, jlo, jhi, &mlo, &mhi);
        // profile sample of output m: half - iown + tstride*m = q*tstride + r
        const int d = half - iown;
        const int q0 = V.fd_tstride.div_floor(d);
        const int r = d - q0 * V.tstride;
        out->base = ps.base + (long long)r * ps.rowlen + q0;
        out->lo = mlo;
        out->hi = mhi;
    } else {
        out->base = ps.base + (long long)half - (long long)iown;            // :283,303
        out->lo = jlo;
        out->hi = jhi;
    }
    return true;
}

// True when a group of a MERGED minor isotope (IsoUnit::merged) is evaluated by the dense
// kernel: it sits at or above dense_from and selects the Doppler sample the main isotope's
// lines select on its cell (`main_bounds`: the main isotope's segments of this strengths pass).
__device__ __forceinline__ bool in_dense_plane(const IsoUnit &I, const int *__restrict__ main_bounds,
                                               int iown, int idop) {
    return iown >= I.dense_from && main_bounds[idop] <= iown && iown < main_bounds[idop + 1];
}

}  // namespace pb200

// engine.cuh -- data structures shared by the engine host code and its kernels.
#pragma once
#include <vector>

#include "common.cuh"

namespace pb200 {

constexpr int kTileOutputs = 256;  // output samples owned by one CTA of the output-owned kernel
#ifndef PB200_CHUNK_TILE
#define PB200_CHUNK_TILE 512
#endif
constexpr int kChunkTile = PB200_CHUNK_TILE;  // ... of the chunk-owned kernel (multiple of 256)
constexpr int kMaxIso = 256;

// Exact division of a non-negative 31-bit integer by a launch-invariant divisor with one
// multiply-high and one shift (Granlund-Montgomery round-up magic numbers).
struct FastDiv {
    unsigned mul = 0, shr = 0;
    int d = 1;
    int one = 1;  // 1 when d == 1 (then mul == 0 and the quotient is n itself)
    void set(int denom) {
        d = denom;
        one = 0;
        if (denom <= 1) { mul = 0; shr = 0; d = 1; one = 1; return; }
        unsigned lg = 0;
        while ((1ull << lg) < (unsigned long long)denom) lg++;  // ceil(log2(denom))
        const unsigned p = 31 + lg;
        mul = (unsigned)(((1ull << p) + (unsigned)denom - 1) / (unsigned)denom);
        shr = p - 32;
    }
#ifdef __CUDACC__
    // All four are branch-free (the sign is folded in with masks).
    __device__ __forceinline__ int div(int n) const {  // 0 <= n < 2^31
        return (int)(__umulhi((unsigned)n, mul) >> shr) + n * one;
    }
    __device__ __forceinline__ int div_trunc(int n) const {  // C semantics, any sign
        const int s = n >> 31;
        return (div((n ^ s) - s) ^ s) - s;
    }
    __device__ __forceinline__ int div_ceil(int n) const {  // n >= 0
        return div(n + d - 1);
    }
    __device__ __forceinline__ int div_floor(int n) const {  // any sign
        const int s = n >> 31;
        return (div(((n ^ s) - s) + (s & (d - 1))) ^ s) - s;
    }
#endif
};

// Per-profile descriptor read with one 16-byte load.
struct __align__(16) ProfileSlot {
    long long base;  // start of the profile (reference layout) or of its transposed block
    int half;        // half size (aliases resolved)
    int rowlen;      // transposed block: row length Q; reference layout: unused
};

// Device view of everything that does not depend on (T,p).  Passed to kernels by value.
struct StaticView {
    // spectral grids
    const double *wn;  // [nwave] output grid
    int nwave;
    long long onwn;    // fine-grid samples
    double own0, ownstep, own_last, wn0;
    // Voigt table
    const double *profile;
    const int *psize;          // [nlor*ndop] half sizes (aliases resolved)
    const long long *pindex;   // [nlor*ndop] start index
    const int *pmaxrow;        // [nlor*ndop] running maximum of psize along the Doppler axis
    int cut_fine;              // cutoff in fine samples (+1), INT_MAX when there is no cutoff
    const double *doppler;     // [ndop]
    // dop_thr[j] = smallest double v whose nearest Doppler sample is >= j (thr[0] = 0): the
    // nearest index is a step function of v, so the O(log n) search becomes one or two
    // comparisons.  NULL when the grid is not strictly increasing (generic search then).
    const double *dop_thr;
    int nlor, ndop;
    int dop_hi0;               // high word of doppler[0]            (nearest_index_log)
    float dop_inv_step;        // (ndop-1)/log2(doppler[-1]/doppler[0])/2^20
    // Output-stride ("transposed") copy of the Voigt table for constant-step output grids:
    // profile p is stored as rows of every tstride-th sample, T[r][q] = profile[q*tstride+r],
    // so the samples one line contributes to consecutive output points are contiguous.
    const double *tprofile;
    const long long *tbase;    // [nlor*ndop] start of the profile's transposed block
    const int *trow;           // [nlor*ndop] row length Q = ceil((2*size+1)/tstride)
    const ProfileSlot *pslot;  // [nlor*ndop] {pindex, half, -} : one 16-byte load per group
    const ProfileSlot *tslot;  // [nlor*ndop] {tbase, half, trow}
    int tstride;               // fine samples per output sample (0: no transposed copy)
    FastDiv fd_tstride;
    // co-add groups (sorted by isotope, then wavenumber)
    const double *l_wn, *l_elow, *l_gf;   // in-window lines, member order
    const double *g_wn;                   // head-line wavenumber
    const int *g_iown;                    // head-line nearest fine index
    const unsigned int *g_start;          // [ngroups+1] first member line
    const unsigned short *g_iso;          // isotope of the group
    long long ngroups;
    // per-isotope coarse index: gbin[iso*(nbins+1)+b] = first group with iown >= b*binw
    const int *gbin;
    int nbins, binw;
    FastDiv fd_binw;
    int niso;
    const double *iso_ratio;  // [niso]
};

// Per-(T,p) unit quantities computed on the host exactly as _extcoeff.c:138-200 does.
struct UnitParams {
    double dwnstep;    // ownstep*ofactor                 (:194)
    double inv_dwnstep;  // RN(1/dwnstep): (w-own0)/dwnstep is formed from it exactly (quotient_rn)
    double cut_steps;  // cutoff/dwnstep                  (:295,297)
    int tpass;         // strengths pass (distinct T, Z) within the current chunk
    int ofactor;       // dynamic oversampling divisor    (:193)
    int scale;         // round(wnstep/ownstep/ofactor)   (:330)
    int dnwn;          // dynamic samples                 (:195)
    int mcount;        // resampled outputs written       (utils.h:130)
    int out_index;     // position of this unit in the caller's batch
    int aslot;         // dense path: slot of this ofactor's anomaly bitmask (dense_kernels.cu)
    FastDiv fd_ofactor, fd_scale;
};

struct IsoUnit {
    double adop;  // Doppler HWHM per unit wavenumber   (:170)
    double dens;  // number density of the isotope's species if add else 1 (:271-272)
    int ilor;     // nearest Lorentz grid index          (:183)
    int reach;    // fine samples a line of this isotope can reach from its centre
    // Split between the two accumulate paths: groups of this isotope on fine cells below
    // `dense_from` go through the gather kernels, the others through the dense convolution
    // (INT_MAX: all gathered, the default; dense_kernels.cu).
    int dense_from;
    // 1: a minor isotope merged into the dense plane of the main one (dense_kernels.cu): its
    // groups at or above dense_from that select the SAME Doppler sample as the main isotope
    // does on their cell are in that plane; the gather kernels take the others.
    int merged;
};

}  // namespace pb200

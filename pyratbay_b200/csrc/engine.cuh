// engine.cuh -- data structures shared by the engine host code and its kernels.
#pragma once
#include <vector>

#include "common.cuh"

namespace pb200 {

constexpr int kTileOutputs = 256;  // output samples owned by one CTA of the accumulate kernel
constexpr int kMaxIso = 256;

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
    const double *doppler;     // [ndop]
    int nlor, ndop;
    // Output-stride ("transposed") copy of the Voigt table for constant-step output grids:
    // profile p is stored as rows of every tstride-th sample, T[r][q] = profile[q*tstride+r],
    // so the samples one line contributes to consecutive output points are contiguous.
    const double *tprofile;
    const long long *tbase;    // [nlor*ndop] start of the profile's transposed block
    const int *trow;           // [nlor*ndop] row length Q = ceil((2*size+1)/tstride)
    int tstride;               // fine samples per output sample (0: no transposed copy)
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
    int niso;
    const double *iso_ratio;  // [niso]
};

// Per-(T,p) unit quantities computed on the host exactly as _extcoeff.c:138-200 does.
struct UnitParams {
    double dwnstep;    // ownstep*ofactor                 (:194)
    double cut_steps;  // cutoff/dwnstep                  (:295,297)
    int tpass;         // strengths pass (distinct T, Z) within the current chunk
    int ofactor;       // dynamic oversampling divisor    (:193)
    int scale;         // round(wnstep/ownstep/ofactor)   (:330)
    int dnwn;          // dynamic samples                 (:195)
    int mcount;        // resampled outputs written       (utils.h:130)
    int out_index;     // position of this unit in the caller's batch
};

struct IsoUnit {
    double adop;  // Doppler HWHM per unit wavenumber   (:170)
    double dens;  // number density of the isotope's species if add else 1 (:271-272)
    int ilor;     // nearest Lorentz grid index          (:183)
    int reach;    // fine samples a line of this isotope can reach from its centre
};

}  // namespace pb200

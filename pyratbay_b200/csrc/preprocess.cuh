// preprocess.cuh -- device-side static line pre-processing (preprocess.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace pb200 {

struct GroupInput {
    long long nlines = 0;
    const double *wn = nullptr, *elow = nullptr, *gf = nullptr;  // host, TLI order
    const unsigned short *iso16 = nullptr;                       // host, isotope per line
    const double *own = nullptr;                                 // host fine grid
    long long onwn = 0;
    int niso = 0;
    int nblocks = 0;                       // isotope blocks in file order
    const long long *block_start = nullptr;
    const int *block_iso = nullptr;
    int nbins = 0, binw = 1;               // coarse index geometry
};

struct GroupOutput {
    // The callback allocates the engine's device buffers for n_inwin lines and ngroups groups
    // (+1 for g_start, niso*(nbins+1) for gbin) and fills the pointers below.
    void *ctx = nullptr;
    int (*alloc)(void *ctx, long long n_inwin, long long ngroups) = nullptr;
    double *l_wn = nullptr, *l_elow = nullptr, *l_gf = nullptr, *g_wn = nullptr;
    int *g_iown = nullptr, *gbin = nullptr;
    unsigned int *g_start = nullptr;
    unsigned short *g_iso = nullptr;
    long long n_inwin = 0, ngroups = 0;
    int max_segment = 0;                   // longest independent segment walked by one thread
    std::vector<int> iso_gbeg, iso_gend;   // group range of every isotope
    std::vector<long long> iso_nadd;       // absorbed lines per isotope
};

// Window filter, nearest fine index, greedy co-add grouping, compaction and coarse index on
// the device.  Synchronises `st`.  Returns 0 or a PB200_E* code.
int device_group_lines(cudaStream_t st, const GroupInput &in, GroupOutput *out);

}  // namespace pb200

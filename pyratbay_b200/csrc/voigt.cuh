// voigt.cuh -- host-side interface of the Voigt grid kernel (voigt.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace pb200 {

constexpr int kVoigtSeriesLen = 61;          // MAXCONV, voigt.h:59
constexpr int kVoigtQuickElements = 99999;   // _voigt_maxelements, voigt.h:124

// The profiles that are actually computed (psize != 0 on input), in table order.
struct VoigtPlan {
    std::vector<int64_t> start;  // first bin of each computed profile
    std::vector<int> half;       // half-size
    std::vector<int> ilor, idop;
    int64_t total = 0;           // bins in the concatenated table
};

// Resolve aliases and start indices exactly like vprofile.c:67-108 (psize/pindex updated
// in place).  Returns 0 on success.
int voigt_plan(int nlor, int ndop, int64_t *psize, int64_t *pindex, VoigtPlan *plan);

// Fill d_profile[0:plan.total] on `stream`; synchronises the stream before returning.
int voigt_launch(cudaStream_t stream, const VoigtPlan &plan, const double *lorentz,
                 const double *doppler, double dwn, double *d_profile, int64_t *launches);

}  // namespace pb200

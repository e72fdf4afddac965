// voigt.cu -- Voigt profile grid on the device (kernel 2 of the engine).
//
// Replaces vprofile.grid (reference src_c/vprofile.c:42-114) and the per-profile sampling
// of voigtn/voigtxy (src_c/include/voigt.h:147-295).  One thread produces one output bin of
// one profile: it evaluates the (oversampled) Voigt function at the bin's sub-samples and
// forms the same trapezoid / Simpson bin mean as the reference, or the point sample in
// VOIGT_QUICK mode.  All profiles of the (Lorentz x Doppler) grid are covered by a single
// launch over the flattened bin index; the table never leaves HBM unless the host asks.
#include "voigt.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace pb200 {

__constant__ double c_series[kVoigtSeriesLen];  // 1/(n!(2n+1)), voigt.h:61-123
__constant__ double c_regA[6] = {0.46131350, 0.19016350, 0.09999216,
                                 1.78449270, 0.002883894, 5.52534370};  // constants.h:28-33
__constant__ double c_regB[4] = {0.51242424, 0.27525510, 0.05176536, 2.72474500};  // :35-38

// Re[w(x+iy)] * sqrt(ln2/pi)/alphaD in the three regions of voigt.h:147-217.
// `pre` = SQRTLN2PI / alphaD.
__device__ double voigt_point(double x, double y, double pre) {
    const double re2 = dsub(dmul(x, x), dmul(y, y));
    const double im2 = dmul(dmul(2.0, x), y);
    if (x < 3.0 && y < 1.8) {
        // Region I: Taylor series of erf(-iz).  The reference carries these sums in x87
        // extended precision; plain fp64 differs by <1e-13 relative (kernel 2).
        const int nterms = (x < 1.0 ? 15 : (int)(dadd(dmul(6.842, x), 8.0))) + 1;
        double pr = y, pi = -x, sr = y, si = -x;
        for (int i = 1; i <= nterms; i++) {
            const double ti = pr * im2 + pi * re2;
            const double tr = pr * re2 - pi * im2;
            si += ti * c_series[i];
            sr += tr * c_series[i];
            pi = ti;
            pr = tr;
        }
        double s, c;
        sincos(im2, &s, &c);
        return pre * exp(-re2) *
               (c * (1.0 - sr * kTwoOverSqrtPi) - s * si * kTwoOverSqrtPi);
    }
    const double q = im2 * im2;
    const double p = im2 * x;
    if (x < 5.0 && y < 5.0) {  // Region II
        const double d1 = re2 - c_regA[1], d2 = re2 - c_regA[3], d3 = re2 - c_regA[5];
        return pre * (c_regA[0] * ((p - d1 * y) / (d1 * d1 + q)) +
                      c_regA[2] * ((p - d2 * y) / (d2 * d2 + q)) +
                      c_regA[4] * ((p - d3 * y) / (d3 * d3 + q)));
    }
    const double d1 = re2 - c_regB[1], d2 = re2 - c_regB[3];  // Region III
    return pre * (c_regB[0] * ((p - d1 * y) / (d1 * d1 + q)) +
                  c_regB[2] * ((p - d2 * y) / (d2 * d2 + q)));
}

struct ProfileDesc {
    long long start;  // first bin in the concatenated table
    int nbin;         // 2*half+1
    int over;         // sub-samples per bin minus one (1: trapezoid of two, even: Simpson)
    int quick;        // point sampling
    double half_width;  // dwn * half
    double fine_step;
    double y;         // sqrt(ln2) aL/aD
    double alpha_d;
};

__global__ void __launch_bounds__(256)
voigt_grid_kernel(const ProfileDesc *__restrict__ desc, int nprof, long long total_bins,
                  double *__restrict__ profile) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total_bins;
         gid += stride) {
        // profile that owns this bin: last desc with start <= gid
        int lo = 0, hi = nprof;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (desc[mid].start <= gid) lo = mid; else hi = mid;
        }
        const ProfileDesc d = desc[lo];
        const int b = (int)(gid - d.start);
        const double pre = ddiv(kSqrtLn2OverPi, d.alpha_d);
        auto sample = [&](int i) {
            // x = SQRTLN2 * |dint*i - dwn| / alphaD   (voigt.h:271)
            const double off = fabs(dsub(dmul(d.fine_step, (double)i), d.half_width));
            const double x = ddiv(dmul(kSqrtLn2, off), d.alpha_d);
            return voigt_point(x, d.y, pre);
        };
        double out;
        if (d.quick) {
            out = sample(b);
        } else if (d.over == 1) {  // meanintegTrap with two points per bin (voigt.h:339-359)
            out = dadd(sample(b), sample(b + 1)) / 2.0;
        } else {  // meanintegSimp (voigt.h:300-334), same summation order
            const int base = b * d.over;
            double acc = 0.0;
            for (int i = 1; i < d.over; i += 2) acc = dadd(acc, sample(base + i));
            acc = dmul(acc, 2.0);
            for (int i = 2; i < d.over; i += 2) acc = dadd(acc, sample(base + i));
            acc = dmul(acc, 2.0);
            acc = dadd(acc, dadd(sample(base), sample(base + d.over)));
            out = ddiv(acc, dmul((double)d.over, 3.0));
        }
        profile[gid] = out;
    }
}

static void fill_series(double *coef) {
    // voigt.h:62-122 tabulates 1/(n!(2n+1)) to 21 digits; regenerate and round once.
    long double fact = 1.0L;
    for (int n = 0; n < kVoigtSeriesLen; n++) {
        if (n > 0) fact *= (long double)n;
        coef[n] = (double)(1.0L / (fact * (long double)(2 * n + 1)));
    }
}

int voigt_plan(int nlor, int ndop, int64_t *psize, int64_t *pindex, VoigtPlan *plan) {
    plan->start.clear();
    plan->half.clear();
    plan->ilor.clear();
    plan->idop.clear();
    int64_t cursor = 0;
    for (int m = 0; m < nlor; m++) {
        for (int n = 0; n < ndop; n++) {
            const size_t at = (size_t)m * ndop + n;
            if (psize[at] != 0) {
                if (psize[at] < 0 || psize[at] > (int64_t)0x3fffffff) {
                    set_error("voigt grid: profile half-size out of range");
                    return -1;
                }
                pindex[at] = cursor;
                plan->start.push_back(cursor);
                plan->half.push_back((int)psize[at]);
                plan->ilor.push_back(m);
                plan->idop.push_back(n);
                cursor += 2 * psize[at] + 1;
            } else {
                if (n == 0) {
                    set_error("voigt grid: psize[m,0] must be non-zero (vprofile.c:101 "
                              "aliases to the previous Doppler sample)");
                    return -1;
                }
                pindex[at] = pindex[at - 1];  // vprofile.c:101-103
                psize[at] = psize[at - 1];
            }
        }
    }
    plan->total = cursor;
    return 0;
}

int voigt_launch(cudaStream_t stream, const VoigtPlan &plan, const double *lorentz,
                 const double *doppler, double dwn, double *d_profile, int64_t *launches) {
    const int nprof = (int)plan.start.size();
    if (nprof == 0 || plan.total == 0) return 0;
    double coef[kVoigtSeriesLen];
    fill_series(coef);
    PB_CUDA(cudaMemcpyToSymbolAsync(c_series, coef, sizeof(coef), 0, cudaMemcpyHostToDevice,
                                    stream));

    std::vector<ProfileDesc> desc(nprof);
    for (int p = 0; p < nprof; p++) {
        ProfileDesc &d = desc[p];
        const int nbin = 2 * plan.half[p] + 1;
        const double al = lorentz[plan.ilor[p]], ad = doppler[plan.idop[p]];
        d.start = plan.start[p];
        d.nbin = nbin;
        d.half_width = dwn * (long)(nbin / 2);  // vprofile.c:82
        d.alpha_d = ad;
        d.y = kSqrtLn2 * al / ad;  // voigt.h:234
        // Sampling mode, voigt.h:236-261
        const double bin_step = 2.0 * d.half_width / (nbin - 1);
        double fine_step = ad / (50 - 1);
        d.quick = nbin > kVoigtQuickElements ? 1 : 0;
        if (bin_step < fine_step || d.quick) {
            d.over = 1;
            fine_step = bin_step;
        } else {
            int over = (int)(bin_step / fine_step) + 1;
            if (over & 1) over++;
            const long npts = (long)nbin * over + 1;
            if (npts > 0x7fffffffL) {
                set_error("voigt grid: oversampled profile exceeds 2^31 samples");
                return -1;
            }
            d.over = over;
            fine_step = 2.0 * d.half_width / (double)(npts - 1);
        }
        d.fine_step = fine_step;
    }
    ProfileDesc *d_desc = nullptr;
    PB_CUDA(cudaMallocAsync((void **)&d_desc, sizeof(ProfileDesc) * nprof, stream));
    PB_CUDA(cudaMemcpyAsync(d_desc, desc.data(), sizeof(ProfileDesc) * nprof,
                            cudaMemcpyHostToDevice, stream));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (plan.total + 255) / 256;
    const int blocks = (int)std::min<long long>(want, (long long)sms * 32);
    voigt_grid_kernel<<<blocks, 256, 0, stream>>>(d_desc, nprof, plan.total, d_profile);
    PB_CUDA(cudaGetLastError());
    if (launches) (*launches)++;
    // desc must outlive the async copy: synchronise before the vector is destroyed.
    PB_CUDA(cudaStreamSynchronize(stream));
    PB_CUDA(cudaFreeAsync(d_desc, stream));
    return 0;
}

}  // namespace pb200

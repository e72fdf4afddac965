// engine.cu -- host side of the pb200 engine and its C ABI (include/pb200_lbl.h).
//
// The handle owns device copies of every (T,p)-independent input of the reference's
// ec.extinction call (spectral grids, Voigt table, line list) plus the static line
// pre-processing; a batch call evaluates any number of (T,p) units with two kernels per
// chunk (strengths, accumulate).  See DESIGN.md for the data layout.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pb200_lbl.h"
#include "dense_kernels.cuh"
#include "lbl_kernels.cuh"
#include "preprocess.cuh"
#include "voigt.cuh"

namespace pb200 {

static thread_local std::string g_error;

void set_error(const std::string &msg) { g_error = msg; }

int cuda_fail(cudaError_t err, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in %s", (int)err,
             cudaGetErrorString(err), file, line, what);
    g_error = buf;
    if (err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver) return PB200_ENODEVICE;
    if (err == cudaErrorMemoryAllocation) return PB200_ENOMEM;
    return PB200_ECUDA;
}

static int fail(int code, const std::string &msg) {
    g_error = msg;
    return code;
}

static int select_device(int device) {
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(PB200_ENODEVICE,
                    "no usable CUDA device: this engine has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(PB200_EINVAL, "device index out of range");
    PB_CUDA(cudaSetDevice(device));
    return 0;
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        if (count <= n && p) return 0;
        release();
        if (count == 0) return 0;
        PB_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
        n = count;
        return 0;
    }
    int upload(const T *src, size_t count, cudaStream_t st) {
        int rc = alloc(count);
        if (rc) return rc;
        if (count) PB_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
        return 0;
    }
};

// binsearchapprox (src_c/include/utils.h:75-89): bisection keeping a[lo] <= v < a[hi],
// then the closer end, ties to the lower index.
static int nearest_bisect(const double *a, double v, int lo, int hi) {
    while (hi - lo > 1) {
        const int mid = (hi + lo) / 2;
        if (a[mid] > v) hi = mid; else lo = mid;
    }
    return (std::fabs(a[hi] - v) < std::fabs(a[lo] - v)) ? hi : lo;
}

// Host replica of the device's nearest_index (common.cuh): clamp, bisection on a[mid] < v,
// closer end with ties to the lower index.
static int nearest_index_host(const double *a, int n, double v) {
    if (v < a[0]) return 0;
    if (a[n - 1] < v) return n - 1;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (hi + lo) >> 1;
        if (a[mid] < v) lo = mid; else hi = mid;
    }
    return (std::fabs(a[hi] - v) < std::fabs(a[lo] - v)) ? hi : lo;
}

// thr[j] = smallest double v with nearest_index(a, n, v) >= j, found by bisection on the bit
// pattern (positive doubles order like their bits; the predicate is monotonic in v because
// RN(a[hi]-v) and RN(v-a[lo]) are).  Empty when the grid is not strictly increasing/positive.
static std::vector<double> nearest_thresholds(const double *a, int n) {
    std::vector<double> thr;
    if (n < 2 || !(a[0] > 0.0)) return thr;
    for (int j = 1; j < n; j++)
        if (!(a[j] > a[j - 1]) || !std::isfinite(a[j])) return thr;
    thr.assign(n, 0.0);
    for (int j = 1; j < n; j++) {
        uint64_t lo, hi;  // nearest(lo) < j <= nearest(hi)
        std::memcpy(&lo, &a[j - 1], 8);
        std::memcpy(&hi, &a[j], 8);
        while (hi - lo > 1) {
            const uint64_t mid = lo + (hi - lo) / 2;
            double v;
            std::memcpy(&v, &mid, 8);
            if (nearest_index_host(a, n, v) >= j) hi = mid; else lo = mid;
        }
        std::memcpy(&thr[j], &hi, 8);
    }
    return thr;
}

}  // namespace pb200

using namespace pb200;

struct pb200_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // D2H of the first half of a batch while the second runs
    cudaEvent_t ev_half = nullptr;
    cudaEvent_t ev[6] = {};
    int64_t launches = 0;
    double timing[5] = {0, 0, 0, 0, 0};

    // grids
    bool has_grid = false;
    int64_t nwave = 0, onwn = 0;
    std::vector<double> wn, own;
    std::vector<int64_t> divisors;
    DevBuf<double> d_wn;

    // voigt
    bool has_voigt = false;
    int nlor = 0, ndop = 0;
    double cutoff = 0.0;
    std::vector<double> lorentz, doppler;
    std::vector<int> psize;
    std::vector<long long> pindex;
    int64_t profile_len = 0;
    DevBuf<double> d_profile, d_doppler;
    DevBuf<int> d_psize, d_pmaxrow;
    DevBuf<long long> d_pindex;
    // output-stride copy of the table (built on first use for constant-step output grids)
    int tstride = 0;
    DevBuf<double> d_tprofile;
    DevBuf<long long> d_tbase;
    DevBuf<int> d_trow;
    DevBuf<ProfileSlot> d_pslot, d_tslot;
    DevBuf<double> d_dop_thr;
    std::vector<double> dop_thr_h;
    bool has_dop_thr = false;

    // species
    bool has_species = false;
    int nmol = 0, niso = 0;
    std::vector<double> mol_radius, mol_mass, iso_mass, iso_ratio;
    std::vector<int> iso_imol;
    DevBuf<double> d_iso_ratio;

    // partition tables
    int pf_ntemp = 0;
    std::vector<double> pf_temp, pf_z;

    // lines
    bool has_lines = false;
    int64_t n_inwin = 0, ngroups = 0, nadd = 0;
    std::vector<int64_t> iso_nadd;  // absorbed lines per isotope
    DevBuf<double> d_lwn, d_lelow, d_lgf, d_gwn;
    DevBuf<int> d_giown, d_gbin;
    DevBuf<unsigned int> d_gstart;
    DevBuf<unsigned short> d_giso;
    DevBuf<int> d_lgroup;    // co-add group of every in-window line (strengths kernel)
    DevBuf<unsigned short> d_liso;
    int nbins = 0, binw = 1;

    // dense-convolution accumulate path (dense_kernels.cu)
    std::vector<long long> iso_gbeg, iso_gend;   // group range of every isotope
    DevBuf<double> d_kd;                         // dense strengths of one (T, isotope) [onwn]
    DevBuf<double> d_kd_pool;                    // sum plane + one array per merged minor isotope
    DevBuf<int> d_bounds, d_dense_err;           // Doppler segments [ndop+1]; error flag
    DevBuf<int> d_bounds_main;                   // main isotope's segments of every pass of a chunk
    DevBuf<int> d_bounds_minor;                  // merged minor isotopes' segments [pass][minor][ndop+1]
    struct DenseIso {
        DevBuf<unsigned> abits;                  // [ndivs][words] anomaly bitmask per ofactor
        std::vector<char> built;                 // per divisor slot
    };
    std::map<int, DenseIso> dense_iso;
    bool dense_off = false;                      // set when a line list violates the path's premise
    int64_t dense_unit_isos = 0;                 // (unit, isotope) pairs of the last batch on it
    double dense_ms = 0.0;                       // device time of the dense kernels, last batch
    cudaEvent_t ev_dense[2] = {};

    // per-batch scratch (grown on demand)
    DevBuf<double> d_ksum, d_out, d_partial;
    DevBuf<double> d_dyn;   // dynamic-grid spectra of constant-R units (kModeDynGrid)
    int sm_count = 0;
    DevBuf<unsigned long long> d_kmax, d_counters;
    // per-batch scalars (iso_row, 1/T, 1/Z, UnitParams, IsoUnit) travel in ONE copy from a
    // pinned staging block
    DevBuf<char> d_stage;
    char *h_stage = nullptr;
    size_t h_stage_bytes = 0;

    StaticView view() const {
        StaticView V;
        V.wn = d_wn.p;
        V.nwave = (int)nwave;
        V.onwn = onwn;
        V.own0 = own.empty() ? 0.0 : own[0];
        V.ownstep = own.size() > 1 ? own[1] - own[0] : 0.0;
        V.own_last = own.empty() ? 0.0 : own[onwn - 1];
        V.wn0 = wn.empty() ? 0.0 : wn[0];
        V.profile = d_profile.p;
        V.psize = d_psize.p;
        V.pindex = d_pindex.p;
        V.pmaxrow = d_pmaxrow.p;
        V.cut_fine = 0x7fffffff;
        if (cutoff > 0.0 && own.size() > 1) {
            const double c = cutoff / (own[1] - own[0]) + 1.0;
            if (c < 2.0e9) V.cut_fine = (int)c;
        }
        V.doppler = d_doppler.p;
        V.dop_thr = has_dop_thr ? d_dop_thr.p : nullptr;
        V.pslot = d_pslot.p;
        V.tslot = d_tslot.p;
        V.nlor = nlor;
        V.ndop = ndop;
        V.dop_hi0 = 0;
        V.dop_inv_step = 0.f;
        if (ndop >= 2 && doppler[0] > 0.0 && doppler[ndop - 1] > doppler[0]) {
            long long bits;
            std::memcpy(&bits, &doppler[0], sizeof(bits));
            V.dop_hi0 = (int)(bits >> 32);
            V.dop_inv_step = (float)((ndop - 1) / std::log2(doppler[ndop - 1] / doppler[0]) /
                                     1048576.0);
        }
        V.tprofile = d_tprofile.p;
        V.tbase = d_tbase.p;
        V.trow = d_trow.p;
        V.tstride = tstride;
        V.fd_tstride.set(tstride > 0 ? tstride : 1);
        V.l_wn = d_lwn.p;
        V.l_elow = d_lelow.p;
        V.l_gf = d_lgf.p;
        V.g_wn = d_gwn.p;
        V.g_iown = d_giown.p;
        V.g_start = d_gstart.p;
        V.g_iso = d_giso.p;
        V.ngroups = ngroups;
        V.gbin = d_gbin.p;
        V.nbins = nbins;
        V.binw = binw;
        V.fd_binw.set(binw);
        V.niso = niso;
        V.iso_ratio = d_iso_ratio.p;
        return V;
    }
};

extern "C" {

const char *pb200_last_error(void) { return g_error.c_str(); }
const char *pb200_version(void) { return "pb200-lbl 0.1 (sm_100a)"; }

int pb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ----------------------------------------------------------------------------------------
int pb200_voigt_grid(int device, int nlor, int ndop, const double *lorentz,
                     const double *doppler, double dwn, int64_t *psize, int64_t *pindex,
                     double *profile, int64_t profile_len) {
    if (nlor <= 0 || ndop <= 0 || !lorentz || !doppler || !psize || !pindex || !profile)
        return fail(PB200_EINVAL, "pb200_voigt_grid: null or empty argument");
    int rc = select_device(device);
    if (rc) return rc;
    VoigtPlan plan;
    if (voigt_plan(nlor, ndop, psize, pindex, &plan)) return PB200_EINVAL;
    if (plan.total > profile_len)
        return fail(PB200_EINVAL, "pb200_voigt_grid: profile array shorter than sum(2*size+1)");
    DevBuf<double> d_prof;
    rc = d_prof.alloc((size_t)plan.total);
    if (rc) return rc;
    cudaStream_t st;
    PB_CUDA(cudaStreamCreate(&st));
    rc = voigt_launch(st, plan, lorentz, doppler, dwn, d_prof.p, nullptr);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(profile, d_prof.p, sizeof(double) * plan.total,
                                        cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "copy profile", __FILE__, __LINE__);
    }
    cudaStreamDestroy(st);
    return rc;
}

// ----------------------------------------------------------------------------------------
int pb200_engine_create(int device, pb200_engine **out) {
    if (!out) return fail(PB200_EINVAL, "pb200_engine_create: null output");
    *out = nullptr;
    int rc = select_device(device);
    if (rc) return rc;
    pb200_engine *e = new pb200_engine();
    e->device = device;
    cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && err == cudaSuccess; i++) err = cudaEventCreate(&e->ev[i]);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_half, cudaEventDisableTiming);
    for (int i = 0; i < 2 && err == cudaSuccess; i++) err = cudaEventCreate(&e->ev_dense[i]);
    if (err != cudaSuccess) {
        delete e;
        return cuda_fail(err, "engine create", __FILE__, __LINE__);
    }
    *out = e;
    return 0;
}

void pb200_engine_destroy(pb200_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) {
        cudaStreamSynchronize(e->stream);
        cudaStreamDestroy(e->stream);
        if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
        if (e->h_stage) cudaFreeHost(e->h_stage);
        if (e->ev_half) cudaEventDestroy(e->ev_half);
    }
    for (int i = 0; i < 6; i++)
        if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    for (int i = 0; i < 2; i++)
        if (e->ev_dense[i]) cudaEventDestroy(e->ev_dense[i]);
    delete e;
}

int pb200_engine_set_grid(pb200_engine *e, const double *wn, int64_t nwave, const double *own,
                          int64_t onwn, const int64_t *divisors, int ndivs) {
    if (!e || !wn || !own || !divisors || nwave < 2 || onwn < 2 || ndivs < 1)
        return fail(PB200_EINVAL, "pb200_engine_set_grid: need nwave>=2, onwn>=2, ndivs>=1");
    if (onwn > 0x7ffffff0LL || nwave > 0x7ffffff0LL)
        return fail(PB200_EINVAL, "pb200_engine_set_grid: grids limited to 2^31 samples "
                                  "(the reference indexes them with C int)");
    PB_CUDA(cudaSetDevice(e->device));
    e->wn.assign(wn, wn + nwave);
    e->own.assign(own, own + onwn);
    e->divisors.assign(divisors, divisors + ndivs);
    e->nwave = nwave;
    e->onwn = onwn;
    int rc = e->d_wn.upload(wn, (size_t)nwave, e->stream);
    if (rc) return rc;
    PB_CUDA(cudaStreamSynchronize(e->stream));
    e->has_grid = true;
    e->has_lines = false;  // grouping depends on the fine grid
    return 0;
}

// l_group[line] for the line-parallel strengths kernel; group range of every isotope.
static int build_line_groups(pb200_engine *e) {
    e->dense_iso.clear();
    e->dense_off = false;
    e->iso_gbeg.assign(e->niso, 0);
    e->iso_gend.assign(e->niso, 0);
    if (e->ngroups == 0 || e->n_inwin == 0) return 0;
    for (int i = 0; i < e->niso; i++) {
        int ends[2] = {0, 0};
        const int *gb = e->d_gbin.p + (size_t)i * (e->nbins + 1);
        PB_CUDA(cudaMemcpyAsync(&ends[0], gb, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        PB_CUDA(cudaMemcpyAsync(&ends[1], gb + e->nbins, sizeof(int), cudaMemcpyDeviceToHost,
                                e->stream));
        PB_CUDA(cudaStreamSynchronize(e->stream));
        e->iso_gbeg[i] = ends[0];
        e->iso_gend[i] = ends[1];
    }
    int rc = e->d_lgroup.alloc((size_t)e->n_inwin);
    if (rc) return rc;
    if (!rc) rc = e->d_liso.alloc((size_t)e->n_inwin);
    if (rc) return rc;
    rc = launch_line_groups(e->stream, e->d_gstart.p, e->d_giso.p, e->ngroups, e->d_lgroup.p,
                            e->d_liso.p);
    if (rc) return rc;
    PB_CUDA(cudaStreamSynchronize(e->stream));
    e->launches++;
    return 0;
}

static int adopt_voigt_tables(pb200_engine *e, int nlor, int ndop, const double *lorentz,
                              const double *doppler, const int64_t *psize,
                              const int64_t *pindex, double cutoff) {
    e->nlor = nlor;
    e->ndop = ndop;
    e->cutoff = cutoff;
    e->tstride = 0;  // the output-stride copy is rebuilt on demand
    e->d_tprofile.release();
    e->lorentz.assign(lorentz, lorentz + nlor);
    e->doppler.assign(doppler, doppler + ndop);
    e->psize.resize((size_t)nlor * ndop);
    e->pindex.resize((size_t)nlor * ndop);
    for (size_t i = 0; i < (size_t)nlor * ndop; i++) {
        if (psize[i] < 0 || psize[i] > 0x3fffffff)
            return fail(PB200_EINVAL, "voigt: half-size out of range");
        e->psize[i] = (int)psize[i];
        e->pindex[i] = (long long)pindex[i];
    }
    std::vector<int> pmaxrow(e->psize.size());
    for (int m = 0; m < nlor; m++) {
        int best = 0;
        for (int n = 0; n < ndop; n++) {
            best = std::max(best, e->psize[(size_t)m * ndop + n]);
            pmaxrow[(size_t)m * ndop + n] = best;
        }
    }
    int rc = e->d_psize.upload(e->psize.data(), e->psize.size(), e->stream);
    if (!rc) rc = e->d_pmaxrow.upload(pmaxrow.data(), pmaxrow.size(), e->stream);
    if (!rc) rc = e->d_pindex.upload(e->pindex.data(), e->pindex.size(), e->stream);
    if (!rc) rc = e->d_doppler.upload(doppler, (size_t)ndop, e->stream);
    std::vector<ProfileSlot> pslot(e->psize.size());
    for (size_t i = 0; i < pslot.size(); i++) pslot[i] = ProfileSlot{e->pindex[i], e->psize[i], 0};
    if (!rc) rc = e->d_pslot.upload(pslot.data(), pslot.size(), e->stream);
    const std::vector<double> thr = nearest_thresholds(doppler, ndop);
    e->dop_thr_h = thr;
    e->has_dop_thr = !thr.empty();
    if (!rc && e->has_dop_thr) rc = e->d_dop_thr.upload(thr.data(), thr.size(), e->stream);
    if (rc) return rc;
    PB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

int pb200_engine_build_voigt(pb200_engine *e, int nlor, int ndop, const double *lorentz,
                             const double *doppler, double dwn, int64_t *psize,
                             int64_t *pindex, double cutoff) {
    if (!e || nlor <= 0 || ndop <= 0 || !lorentz || !doppler || !psize || !pindex)
        return fail(PB200_EINVAL, "pb200_engine_build_voigt: null or empty argument");
    PB_CUDA(cudaSetDevice(e->device));
    // The table length follows the reference: sum(2*size+1) over the *input* sizes, where
    // skipped entries count one sample (pyrat/voigt.py:142), leaving zero padding at the end.
    int64_t alloc_len = 0;
    for (size_t i = 0; i < (size_t)nlor * ndop; i++) alloc_len += 2 * psize[i] + 1;
    VoigtPlan plan;
    if (voigt_plan(nlor, ndop, psize, pindex, &plan)) return PB200_EINVAL;
    if (alloc_len < plan.total) alloc_len = plan.total;
    e->d_profile.release();
    int rc = e->d_profile.alloc((size_t)alloc_len);
    if (rc) return rc;
    PB_CUDA(cudaMemsetAsync(e->d_profile.p, 0, sizeof(double) * alloc_len, e->stream));
    rc = voigt_launch(e->stream, plan, lorentz, doppler, dwn, e->d_profile.p, &e->launches);
    if (rc) return rc;
    e->profile_len = alloc_len;
    rc = adopt_voigt_tables(e, nlor, ndop, lorentz, doppler, psize, pindex, cutoff);
    if (rc) return rc;
    e->has_voigt = true;
    return 0;
}

int pb200_engine_set_voigt(pb200_engine *e, int nlor, int ndop, const double *lorentz,
                           const double *doppler, const int64_t *psize, const int64_t *pindex,
                           const double *profile, int64_t profile_len, double cutoff) {
    if (!e || nlor <= 0 || ndop <= 0 || !lorentz || !doppler || !psize || !pindex || !profile ||
        profile_len <= 0)
        return fail(PB200_EINVAL, "pb200_engine_set_voigt: null or empty argument");
    PB_CUDA(cudaSetDevice(e->device));
    for (size_t i = 0; i < (size_t)nlor * ndop; i++)
        if (pindex[i] < 0 || pindex[i] + 2 * psize[i] + 1 > profile_len)
            return fail(PB200_EINVAL, "pb200_engine_set_voigt: profile index out of bounds");
    e->d_profile.release();
    int rc = e->d_profile.upload(profile, (size_t)profile_len, e->stream);
    if (rc) return rc;
    e->profile_len = profile_len;
    rc = adopt_voigt_tables(e, nlor, ndop, lorentz, doppler, psize, pindex, cutoff);
    if (rc) return rc;
    e->has_voigt = true;
    return 0;
}

int64_t pb200_engine_profile_len(const pb200_engine *e) { return e ? e->profile_len : 0; }

int pb200_engine_get_profile(pb200_engine *e, double *profile, int64_t profile_len) {
    if (!e || !profile) return fail(PB200_EINVAL, "pb200_engine_get_profile: null argument");
    if (!e->has_voigt) return fail(PB200_ESTATE, "pb200_engine_get_profile: no Voigt table");
    if (profile_len < e->profile_len)
        return fail(PB200_EINVAL, "pb200_engine_get_profile: destination too short");
    PB_CUDA(cudaSetDevice(e->device));
    PB_CUDA(cudaMemcpyAsync(profile, e->d_profile.p, sizeof(double) * e->profile_len,
                            cudaMemcpyDeviceToHost, e->stream));
    PB_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

int pb200_engine_set_species(pb200_engine *e, int nmol, const double *mol_radius,
                             const double *mol_mass, int niso, const int64_t *iso_imol,
                             const double *iso_mass, const double *iso_ratio) {
    if (!e || nmol <= 0 || niso <= 0 || !mol_radius || !mol_mass || !iso_imol || !iso_mass ||
        !iso_ratio)
        return fail(PB200_EINVAL, "pb200_engine_set_species: null or empty argument");
    if (niso > kMaxIso) return fail(PB200_EINVAL, "pb200_engine_set_species: too many isotopes");
    for (int i = 0; i < niso; i++)
        if (iso_imol[i] < 0 || iso_imol[i] >= nmol)
            return fail(PB200_EINVAL, "pb200_engine_set_species: iso_imol out of range");
    PB_CUDA(cudaSetDevice(e->device));
    e->nmol = nmol;
    e->niso = niso;
    e->mol_radius.assign(mol_radius, mol_radius + nmol);
    e->mol_mass.assign(mol_mass, mol_mass + nmol);
    e->iso_mass.assign(iso_mass, iso_mass + niso);
    e->iso_ratio.assign(iso_ratio, iso_ratio + niso);
    e->iso_imol.resize(niso);
    for (int i = 0; i < niso; i++) e->iso_imol[i] = (int)iso_imol[i];
    int rc = e->d_iso_ratio.upload(iso_ratio, (size_t)niso, e->stream);
    if (rc) return rc;
    PB_CUDA(cudaStreamSynchronize(e->stream));
    e->has_species = true;
    e->has_lines = false;
    return 0;
}

int pb200_engine_set_partition(pb200_engine *e, int ntemp, const double *temp, const double *z) {
    if (!e || ntemp < 2 || !temp || !z)
        return fail(PB200_EINVAL, "pb200_engine_set_partition: need ntemp>=2");
    if (!e->has_species) return fail(PB200_ESTATE, "pb200_engine_set_partition: set_species first");
    for (int i = 1; i < ntemp; i++)
        if (!(temp[i] > temp[i - 1]))
            return fail(PB200_EINVAL, "pb200_engine_set_partition: temperatures must increase");
    e->pf_ntemp = ntemp;
    e->pf_temp.assign(temp, temp + ntemp);
    e->pf_z.assign(z, z + (size_t)ntemp * e->niso);
    return 0;
}

// Static line pre-processing: everything in _extcoeff.c:229-262 that does not depend on (T,p).
int pb200_engine_set_lines(pb200_engine *e, int64_t nlines, const double *wn,
                           const double *elow, const double *gf, const int64_t *iso_id) {
    if (!e || nlines < 0 || (nlines > 0 && (!wn || !elow || !gf || !iso_id)))
        return fail(PB200_EINVAL, "pb200_engine_set_lines: null argument");
    if (!e->has_grid || !e->has_species)
        return fail(PB200_ESTATE, "pb200_engine_set_lines: set_grid and set_species first");
    if (nlines > 0x7ffffff0LL)
        return fail(PB200_EINVAL, "pb200_engine_set_lines: more than 2^31 lines "
                                  "(TLI n_transitions is int32, lread.py:299)");
    PB_CUDA(cudaSetDevice(e->device));
    const int niso = e->niso;
    const double *own = e->own.data();
    const int64_t onwn = e->onwn;
    const double own0 = own[0], own_last = own[onwn - 1];
    const double ownstep = own[1] - own[0];

    // Isotopes must come in contiguous blocks, ascending wavenumber inside (TLI layout).
    std::vector<char> seen(niso, 0);
    std::vector<long long> block_start;
    std::vector<int> block_iso;
    std::vector<unsigned short> iso16((size_t)nlines);
    for (int64_t ln = 0; ln < nlines; ln++) {
        const int64_t i = iso_id[ln];
        if (i < 0 || i >= niso)
            return fail(PB200_EINVAL, "pb200_engine_set_lines: isotope id out of range");
        iso16[ln] = (unsigned short)i;
        if (ln == 0 || iso_id[ln - 1] != i) {
            if (seen[i])
                return fail(PB200_EINVAL, "pb200_engine_set_lines: an isotope appears in more "
                                          "than one block; lines must be grouped by isotope");
            seen[i] = 1;
            block_start.push_back(ln);
            block_iso.push_back((int)i);
        } else if (wn[ln] < wn[ln - 1]) {
            return fail(PB200_EINVAL, "pb200_engine_set_lines: wavenumbers must ascend within "
                                      "each isotope block");
        }
    }

    // Coarse per-isotope index geometry over the fine grid.
    const int binw_all = (int)std::max<int64_t>(64, onwn >> 16);
    const int nbins_all = (int)((onwn + binw_all - 1) / binw_all);

    // Device path (default for large lists; PB200_SETLINES=host|device overrides): the
    // grouping, compaction and coarse index are built on the GPU (preprocess.cu).
    {
        const char *mode = std::getenv("PB200_SETLINES");
        const bool on_device = mode ? std::strcmp(mode, "device") == 0 : nlines >= 200000;
        if (on_device && nlines > 0) {
            GroupInput gi;
            gi.nlines = nlines; gi.wn = wn; gi.elow = elow; gi.gf = gf; gi.iso16 = iso16.data();
            gi.own = own; gi.onwn = onwn; gi.niso = niso;
            gi.nblocks = (int)block_start.size();
            gi.block_start = block_start.data(); gi.block_iso = block_iso.data();
            gi.nbins = nbins_all; gi.binw = binw_all;
            GroupOutput go;
            struct Ctx { pb200_engine *e; GroupOutput *go; int niso, nbins; } ctx{e, &go, niso, nbins_all};
            go.ctx = &ctx;
            go.alloc = [](void *c, long long n_inwin, long long ngroups) -> int {
                Ctx *x = (Ctx *)c;
                pb200_engine *en = x->e;
                int r = en->d_lwn.alloc((size_t)std::max<long long>(n_inwin, 1));
                if (!r) r = en->d_lelow.alloc((size_t)std::max<long long>(n_inwin, 1));
                if (!r) r = en->d_lgf.alloc((size_t)std::max<long long>(n_inwin, 1));
                if (!r) r = en->d_gwn.alloc((size_t)std::max<long long>(ngroups, 1));
                if (!r) r = en->d_giown.alloc((size_t)std::max<long long>(ngroups, 1));
                if (!r) r = en->d_gstart.alloc((size_t)ngroups + 1);
                if (!r) r = en->d_giso.alloc((size_t)std::max<long long>(ngroups, 1));
                if (!r) r = en->d_gbin.alloc((size_t)x->niso * (x->nbins + 1));
                if (r) return r;
                x->go->l_wn = en->d_lwn.p; x->go->l_elow = en->d_lelow.p; x->go->l_gf = en->d_lgf.p;
                x->go->g_wn = en->d_gwn.p; x->go->g_iown = en->d_giown.p;
                x->go->g_start = en->d_gstart.p; x->go->g_iso = en->d_giso.p;
                x->go->gbin = en->d_gbin.p;
                return 0;
            };
            e->has_lines = false;
            int rc = device_group_lines(e->stream, gi, &go);
            if (rc) return rc;
            e->launches += 9;
            e->iso_nadd.assign(niso, 0);
            int64_t nadd_dev = 0;
            for (int i = 0; i < niso; i++) {
                e->iso_nadd[i] = go.iso_nadd.empty() ? 0 : go.iso_nadd[i];
                nadd_dev += e->iso_nadd[i];
            }
            e->n_inwin = go.n_inwin;
            e->ngroups = go.ngroups;
            e->nadd = nadd_dev;
            e->nbins = nbins_all;
            e->binw = binw_all;
            rc = build_line_groups(e);
            if (rc) return rc;
            e->has_lines = true;
            return 0;
        }
    }

    std::vector<double> l_wn, l_elow, l_gf, g_wn;
    std::vector<int> g_iown;
    std::vector<unsigned int> g_start;
    std::vector<unsigned short> g_iso;
    l_wn.reserve(nlines);
    l_elow.reserve(nlines);
    l_gf.reserve(nlines);
    e->iso_nadd.assign(niso, 0);
    int64_t nadd = 0;
    for (int64_t ln = 0; ln < nlines; ln++) {
        const double w = wn[ln];
        if (w < own0 || w > own_last) continue;  // :215,239
        const int iso = (int)iso_id[ln];
        int iown = (int)((w - own0) / ownstep);  // :243
        if (iown + 1 < onwn && std::fabs(w - own[iown + 1]) < std::fabs(w - own[iown])) iown++;
        g_wn.push_back(w);
        g_iown.push_back(iown);
        g_iso.push_back((unsigned short)iso);
        g_start.push_back((unsigned int)l_wn.size());
        l_wn.push_back(w);
        l_elow.push_back(elow[ln]);
        l_gf.push_back(gf[ln]);
        // :249-262 absorb the following lines that fall on the same fine sample
        while (ln + 1 != nlines && iso_id[ln + 1] == iso && wn[ln + 1] <= own_last) {
            if (std::fabs(wn[ln + 1] - own[iown]) < ownstep) {
                ln++;
                nadd++;
                e->iso_nadd[iso]++;
                l_wn.push_back(wn[ln]);
                l_elow.push_back(elow[ln]);
                l_gf.push_back(gf[ln]);
            } else {
                break;
            }
        }
    }
    g_start.push_back((unsigned int)l_wn.size());
    const int64_t ngroups = (int64_t)g_wn.size();

    // Coarse per-isotope index over the fine grid.
    const int binw = binw_all, nbins = nbins_all;
    std::vector<int> gbin((size_t)niso * (nbins + 1), 0);
    {
        int64_t g = 0;
        // groups are in line order: isotope blocks, ascending iown inside each block
        std::vector<int64_t> beg(niso, -1), end(niso, -1);
        for (g = 0; g < ngroups; g++) {
            const int iso = g_iso[g];
            if (beg[iso] < 0) beg[iso] = g;
            end[iso] = g + 1;
        }
        for (int iso = 0; iso < niso; iso++) {
            int *gb = gbin.data() + (size_t)iso * (nbins + 1);
            if (beg[iso] < 0) {
                for (int b = 0; b <= nbins; b++) gb[b] = 0;
                continue;
            }
            int64_t cur = beg[iso];
            for (int b = 0; b <= nbins; b++) {
                const int64_t edge = (int64_t)b * binw;
                while (cur < end[iso] && g_iown[cur] < edge) cur++;
                gb[b] = (int)cur;
            }
            gb[nbins] = (int)end[iso];
        }
    }

    cudaStream_t st = e->stream;
    int rc = 0;
    if (!rc) rc = e->d_lwn.upload(l_wn.data(), l_wn.size(), st);
    if (!rc) rc = e->d_lelow.upload(l_elow.data(), l_elow.size(), st);
    if (!rc) rc = e->d_lgf.upload(l_gf.data(), l_gf.size(), st);
    if (!rc) rc = e->d_gwn.upload(g_wn.data(), g_wn.size(), st);
    if (!rc) rc = e->d_giown.upload(g_iown.data(), g_iown.size(), st);
    if (!rc) rc = e->d_gstart.upload(g_start.data(), g_start.size(), st);
    if (!rc) rc = e->d_giso.upload(g_iso.data(), g_iso.size(), st);
    if (!rc) rc = e->d_gbin.upload(gbin.data(), gbin.size(), st);
    if (rc) return rc;
    PB_CUDA(cudaStreamSynchronize(st));
    e->n_inwin = (int64_t)l_wn.size();
    e->ngroups = ngroups;
    e->nadd = nadd;
    e->nbins = nbins;
    e->binw = binw;
    rc = build_line_groups(e);
    if (rc) return rc;
    e->has_lines = true;
    return 0;
}

int pb200_nearest_thresholds(const double *grid, int n, double *thr) {
    if (!grid || !thr || n < 1) return fail(PB200_EINVAL, "pb200_nearest_thresholds: null argument");
    const std::vector<double> t = nearest_thresholds(grid, n);
    if (t.empty()) return fail(PB200_EINVAL, "pb200_nearest_thresholds: grid must be positive and "
                                             "strictly increasing with at least two samples");
    std::copy(t.begin(), t.end(), thr);
    return 0;
}

int pb200_engine_line_stats(const pb200_engine *e, int64_t stats[3]) {
    if (!e || !stats) return fail(PB200_EINVAL, "pb200_engine_line_stats: null argument");
    if (!e->has_lines) return fail(PB200_ESTATE, "pb200_engine_line_stats: no lines loaded");
    stats[0] = e->n_inwin;
    stats[1] = e->ngroups;
    stats[2] = e->nadd;
    return 0;
}

// Build (once per Voigt table and stride) the output-stride copy used by the coalesced
// accumulate mode: every distinct profile becomes a [stride, Q] block.
static int ensure_transposed(pb200_engine *e, int stride) {
    if (e->tstride == stride && e->d_tprofile.p) return 0;
    const size_t nslot = (size_t)e->nlor * e->ndop;
    std::vector<long long> tbase(nslot), src, dst;
    std::vector<int> trow(nslot), nbin, rowlen;
    std::map<long long, size_t> seen;  // reference-layout start -> block id
    long long total = 0;
    for (size_t at = 0; at < nslot; at++) {
        auto it = seen.find(e->pindex[at]);
        if (it == seen.end()) {
            const int nb = 2 * e->psize[at] + 1;
            const int q = (nb + stride - 1) / stride;
            seen.emplace(e->pindex[at], src.size());
            src.push_back(e->pindex[at]);
            dst.push_back(total);
            nbin.push_back(nb);
            rowlen.push_back(q);
            tbase[at] = total;
            trow[at] = q;
            total += (long long)stride * q;
        } else {
            tbase[at] = dst[it->second];
            trow[at] = rowlen[it->second];
        }
    }
    e->tstride = 0;
    // 256 zero samples of padding: the slot-major accumulate path reads (and multiplies by 0)
    // up to 8*32 samples from the table start for the empty tail slots of a chunk.
    int rc = e->d_tprofile.alloc((size_t)total + 256);
    if (!rc) PB_CUDA(cudaMemsetAsync(e->d_tprofile.p + total, 0, 256 * sizeof(double), e->stream));
    if (!rc) rc = e->d_tbase.upload(tbase.data(), nslot, e->stream);
    if (!rc) rc = e->d_trow.upload(trow.data(), nslot, e->stream);
    std::vector<ProfileSlot> tslot(nslot);
    for (size_t at = 0; at < nslot; at++) tslot[at] = ProfileSlot{tbase[at], e->psize[at], trow[at]};
    if (!rc) rc = e->d_tslot.upload(tslot.data(), nslot, e->stream);
    if (rc) return rc;
    rc = launch_transpose(e->stream, (int)src.size(), src.data(), dst.data(), nbin.data(),
                          rowlen.data(), total, stride, e->d_profile.p, e->d_tprofile.p);
    if (rc) return rc;
    e->launches++;
    e->tstride = stride;
    return 0;
}

// ----------------------------------------------------------------------------------------
static int run_batch(pb200_engine *e, int n_units, const double *unit_temp,
                     const double *unit_density, const double *unit_isoz,
                     const int64_t *iso_iext, int nextinct, double ethresh, int add,
                     int resolution, double *out_host, double *out_dev, int64_t *counters,
                     cudaStream_t user_stream) {
    if (!e || n_units < 0 || !unit_temp || !unit_density || !iso_iext || nextinct < 1)
        return fail(PB200_EINVAL, "pb200_extinction_batch: null or empty argument");
    if (!e->has_grid || !e->has_voigt || !e->has_species || !e->has_lines)
        return fail(PB200_ESTATE, "pb200_extinction_batch: engine needs grid, voigt, species "
                                  "and lines before a batch call");
    if (!unit_isoz && e->pf_ntemp == 0)
        return fail(PB200_EINVAL, "pb200_extinction_batch: unit_isoz is NULL and no partition "
                                  "tables were set");
    if (!out_host && !out_dev) return fail(PB200_EINVAL, "pb200_extinction_batch: null output");
    PB_CUDA(cudaSetDevice(e->device));
    if (n_units == 0) return 0;
    cudaStream_t st = e->stream;
    const int niso = e->niso, nmol = e->nmol, nlor = e->nlor, ndop = e->ndop;
    const int nrows = add ? 1 : nextinct;
    const int64_t nwave = e->nwave, onwn = e->onwn;
    const double own0 = e->own[0];
    const double ownstep = e->own[1] - e->own[0];  // :186
    const double wnstep = e->wn[1] - e->wn[0];     // :185
    const double cutoff = e->cutoff;

    // Output row of every isotope (iso_iext, :209-212,234-237).
    std::vector<int> iso_row(niso);
    for (int i = 0; i < niso; i++) {
        if (iso_iext[i] >= nextinct)
            return fail(PB200_EINVAL, "pb200_extinction_batch: iso_iext >= nextinct");
        iso_row[i] = iso_iext[i] < 0 ? -1 : (add ? 0 : (int)iso_iext[i]);
    }

    // Partition functions when the caller leaves them to the engine.
    std::vector<double> isoz_local;
    if (!unit_isoz) {
        isoz_local.resize((size_t)n_units * niso);
        const double *tg = e->pf_temp.data();
        const int nt = e->pf_ntemp;
        for (int u = 0; u < n_units; u++) {
            const double t = unit_temp[u];
            if (!(t >= tg[0] && t <= tg[nt - 1]))
                return fail(PB200_EINVAL, "pb200_extinction_batch: temperature outside the "
                                          "partition-function table");
            int lo = (int)(std::upper_bound(tg, tg + nt, t) - tg) - 1;
            if (lo > nt - 2) lo = nt - 2;
            const double f = (t - tg[lo]) / (tg[lo + 1] - tg[lo]);
            for (int i = 0; i < niso; i++) {
                const double *z = e->pf_z.data() + (size_t)i * nt;
                isoz_local[(size_t)u * niso + i] = z[lo] + (z[lo + 1] - z[lo]) * f;
            }
        }
        unit_isoz = isoz_local.data();
    }

    // Per-unit scalars, exactly as _extcoeff.c:138-200.
    std::vector<UnitParams> units(n_units);
    std::vector<IsoUnit> iso_units((size_t)n_units * niso);
    std::map<std::vector<double>, int> tp_index;
    std::vector<double> tp_temp, tp_isoz;
    std::vector<int> unit_tp(n_units);
    std::vector<long long> unit_table_bytes(n_units, 0);
    for (int u = 0; u < n_units; u++) {
        const double temp = unit_temp[u];
        const double *dens = unit_density + (size_t)u * nmol;
        const double fdop = std::sqrt(2 * kBoltzmann * temp / kAmu) * kSqrtLn2 / kLightSpeed;
        const double flor = std::sqrt(2 * kBoltzmann * temp / kPi / kAmu) / kLightSpeed;
        double minwidth = 1e5;
        for (int i = 0; i < niso; i++) {
            const int imol = e->iso_imol[i];
            double al = 0.0;
            for (int j = 0; j < nmol; j++) {
                const double cd = e->mol_radius[imol] + e->mol_radius[j];
                al += dens[j] * cd * cd * std::sqrt(1 / e->iso_mass[i] + 1 / e->mol_mass[j]);
            }
            al *= flor;
            const double ad = fdop / std::sqrt(e->iso_mass[i]);
            const double dw = ad * own0;
            const double vw = 0.5346 * al + std::sqrt(al * al * 0.2166 + dw * dw);
            minwidth = std::fmin(minwidth, vw);
            IsoUnit &I = iso_units[(size_t)u * niso + i];
            I.adop = ad;
            I.dens = add ? dens[imol] : 1.0;
            I.ilor = nearest_bisect(e->lorentz.data(), al, 0, nlor - 1);
        }
        int d;
        const int ndivs = (int)e->divisors.size();
        for (d = 1; d < ndivs; d++)
            if (e->divisors[d] * ownstep >= 0.5 * minwidth) break;
        const int ofactor = (int)e->divisors[d - 1];
        UnitParams &U = units[u];
        U.ofactor = ofactor;
        U.dwnstep = ownstep * ofactor;
        U.inv_dwnstep = 1.0 / U.dwnstep;
        U.dnwn = (int)(1 + (onwn - 1) / ofactor);
        U.cut_steps = cutoff / U.dwnstep;
        U.scale = (int)std::round(wnstep / ownstep / ofactor);
        if (U.scale < 1) U.scale = 1;
        U.mcount = 1 + (U.dnwn - 1) / U.scale;
        if (U.mcount > nwave) U.mcount = (int)nwave;
        U.out_index = u;
        U.aslot = d - 1;
        U.fd_ofactor.set(U.ofactor);
        U.fd_scale.set(U.scale);
        for (int i = 0; i < niso; i++) {
            IsoUnit &I = iso_units[(size_t)u * niso + i];
            // largest half-size among the Doppler samples lines inside the window can select
            // (aD*wn is monotonic in wn; one extra sample either side for rounding safety)
            const double *dg = e->doppler.data();
            int n0 = nearest_bisect(dg, I.adop * own0, 0, ndop - 1) - 1;
            int n1 = nearest_bisect(dg, I.adop * e->own[onwn - 1], 0, ndop - 1) + 1;
            n0 = std::max(n0, 0);
            n1 = std::min(n1, ndop - 1);
            int pmax = 0;
            long long last_index = -1;
            for (int n = n0; n <= n1; n++) {
                const size_t at = (size_t)I.ilor * ndop + n;
                pmax = std::max(pmax, e->psize[at]);
                // distinct table samples within the cutoff window that this unit's lines of
                // this isotope can read (roofline accounting only)
                if (iso_row[i] >= 0 && e->pindex[at] != last_index) {
                    long long span = 2LL * e->psize[at] + 1;
                    if (cutoff > 0.0)
                        span = std::min<long long>(span, 2LL * (long long)(cutoff / ownstep) + 1);
                    unit_table_bytes[u] += 8 * span;
                    last_index = e->pindex[at];
                }
            }
            long long reach = pmax;
            if (cutoff > 0.0) reach = std::min<long long>(reach, (long long)(cutoff / ownstep) + 1);
            reach += 2LL * ofactor + 2;
            I.reach = (int)std::min<long long>(reach, 0x7fffffffLL);
            I.dense_from = 0x7fffffff;
            I.merged = 0;
        }
        // distinct (T, Z) -> strengths pass
        std::vector<double> key(1 + niso);
        key[0] = temp;
        for (int i = 0; i < niso; i++) key[1 + i] = unit_isoz[(size_t)u * niso + i];
        auto it = tp_index.find(key);
        if (it == tp_index.end()) {
            const int id = (int)tp_temp.size();
            tp_index.emplace(key, id);
            tp_temp.push_back(temp);
            tp_isoz.insert(tp_isoz.end(), key.begin() + 1, key.end());
            unit_tp[u] = id;
        } else {
            unit_tp[u] = it->second;
        }
    }
    const int ntp = (int)tp_temp.size();
    // the strengths kernel forms the exactly rounded quotients x/T and x/Z from {T, RN(1/T)}
    // and {Z, RN(1/Z)} (quotient_rn); the reciprocals are taken here, once
    std::vector<double2> tp_t(tp_temp.size()), tp_z(tp_isoz.size());
    for (size_t i = 0; i < tp_temp.size(); i++) tp_t[i] = make_double2(tp_temp[i], 1.0 / tp_temp[i]);
    for (size_t i = 0; i < tp_isoz.size(); i++) tp_z[i] = make_double2(tp_isoz[i], 1.0 / tp_isoz[i]);

    // Accumulate mode per unit: constant-step outputs read the output-stride table when the
    // unit's dynamic stride ofactor*scale equals the table's stride (always true when
    // wnstep/ownstep is the integer wnosamp); anything else uses the generic strided gather.
    std::vector<int> unit_mode(n_units, resolution ? kModeLinterp : kModeStrided);
    // Constant-R / constant-wavelength grids: a unit whose dynamic grid (step ofactor*ownstep) is
    // not much finer than the output grid is evaluated ON the dynamic grid with the chunk kernel
    // (a constant-step problem with stride ofactor: coalesced gathers from the output-stride
    // table built for that stride) and then interpolated (utils.h:139-163), which is what the
    // reference does; the others gather their two samples per output point directly
    // (kModeLinterp).  A line covers 2*cutoff/dwnstep dynamic samples but up to thousands of
    // output points at high pressure and low wavenumber.  PB200_DYN_FACTOR: use the dynamic
    // grid when dwnstep >= factor * mean output spacing (default 0.1; 0 = never).
    if (resolution) {
        double factor = 0.1;
        if (const char *env = std::getenv("PB200_DYN_FACTOR")) factor = std::atof(env);
        const double mean_step = (e->wn[nwave - 1] - e->wn[0]) / (double)std::max<int64_t>(nwave - 1, 1);
        for (int u = 0; u < n_units; u++) {
            const bool small_table =
                (long long)e->profile_len + (long long)nlor * ndop * units[u].ofactor +
                units[u].dnwn + 512 < 0x7fffffffLL;
            if (factor > 0.0 && small_table && units[u].dwnstep >= factor * mean_step &&
                units[u].dnwn >= 2 * kChunkTile)
                unit_mode[u] = kModeDynGrid;
        }
    }
    // PB200_ACC_MODE=strided forces the generic gather (used by the tests to cover it).
    const char *force = std::getenv("PB200_ACC_MODE");
    const bool allow_transposed = !(force && std::strcmp(force, "strided") == 0);
    if (!resolution && allow_transposed) {
        const int stride = (int)std::llround(wnstep / ownstep);
        // The output-owned kernel packs the table offset into 32 bits and the chunk kernel into
        // 31: a larger output-stride table (> 34 GB) falls back to the strided gather, which
        // uses 64-bit addresses.  PB200_PACK_LIMIT lowers the limit (tests force the fallback).
        long long pack_limit = 0xffffffffLL;
        {
            const char *env = std::getenv("PB200_PACK_LIMIT");
            if (env && std::atoll(env) > 0) pack_limit = std::min(pack_limit, std::atoll(env));
        }
        // the transposed table is at most stride-1 samples per profile longer than the original
        const long long tlen_bound =
            e->profile_len + (long long)e->nlor * e->ndop * (long long)std::max(stride, 1) + 256;
        const bool fits = tlen_bound + nwave + 64 < pack_limit;
        bool any = false;
        for (int u = 0; u < n_units; u++)
            if (fits && stride >= 1 && (long long)units[u].ofactor * units[u].scale == stride) {
                unit_mode[u] = kModeTransposed;
                any = true;
            }
        if (any) {
            int rc_t = ensure_transposed(e, stride);
            if (rc_t) return rc_t;
        }
    }

    // Dense-convolution path (dense_kernels.cu) for isotopes whose groups fill a large share of
    // the fine grid: the gather kernels leave out the cells the dense kernel owns (per unit and
    // isotope everything at or above `dense_from`), the dense kernel adds them afterwards.
    // Only where its premise holds: constant-step grid on the output-stride table, windows that
    // are translation invariant (cutoff/dwnstep not within 1e-6 of an integer, so that the
    // reference's (int)(idwn +- cutoff/dwnstep) is idwn +- a constant), footprints that fit
    // the shared tiles, grid ends far from the reference cell.  PB200_DENSE=0 disables it,
    // PB200_DENSE_MIN_OCC sets the occupancy threshold (groups / fine samples, default 0.25:
    // the gather path delivers ~2.6e12 samples/s, the dense one ~1.4e13 MAC/s over all cells).
    int rc = 0;
    std::vector<int> dense_isos, minor_isos;
    int main_iso = -1;
    std::vector<char> unit_dense(n_units, 0), unit_merged(n_units, 0);
    e->dense_unit_isos = 0;
    e->dense_ms = 0.0;
    {
        const char *env = std::getenv("PB200_DENSE");
        const bool allow = !(env && std::strcmp(env, "0") == 0) && !e->dense_off;
        double min_occ = 0.25;
        if (const char *mo = std::getenv("PB200_DENSE_MIN_OCC")) min_occ = std::atof(mo);
        const int stride = e->tstride;
        if (allow && !resolution && stride >= 1 && stride <= kDenseMaxStride &&
            nwave >= 4 * kDenseTile && onwn < 0x7fffffffLL) {
            for (int i = 0; i < niso; i++)
                if (iso_row[i] >= 0 &&
                    (double)(e->iso_gend[i] - e->iso_gbeg[i]) >= min_occ * (double)onwn)
                    dense_isos.push_back(i);
        }
        if (!dense_isos.empty()) {
            // A footprint narrower than ~16 outputs leaves the dense kernel's register tile
            // mostly idle (a half-warp walks 15 + span cells whatever the span), and footprints
            // grow with wavenumber (Doppler width): per unit and isotope the cells below
            // `dense_from` (narrow Doppler samples) stay with the gather kernels.  Measured on
            // the bench table: 8 / 16 / 24 / 32 outputs -> flat within 1.5 % around 16.
            int min_span = 16;
            if (const char *ms = std::getenv("PB200_DENSE_MIN_SPAN")) min_span = std::atoi(ms);
            const int cut_fine = cutoff > 0.0 ? (int)std::min(cutoff / ownstep + 1.0, 2.0e9)
                                              : 0x7fffffff;
            bool any = false;
            for (int u = 0; u < n_units; u++) {
                bool ok = unit_mode[u] == kModeTransposed && e->has_dop_thr;
                if (cutoff > 0.0) {
                    const double cs = units[u].cut_steps, f = cs - std::floor(cs);
                    if (f != 0.0 && std::min(f, 1.0 - f) < 1e-6) ok = false;
                }
                for (int di : dense_isos) {
                    const long long r = iso_units[(size_t)u * niso + di].reach / stride + 2;
                    if (2 * r + 2 > kDenseSpanMax || units[u].mcount < 8 * r + 2 * kDenseTile)
                        ok = false;
                }
                bool some = false;
                for (int di : dense_isos) {
                    if (!ok) break;
                    IsoUnit &I = iso_units[(size_t)u * niso + di];
                    for (int j = 0; j < ndop; j++) {
                        const long long half =
                            std::min<long long>(e->psize[(size_t)I.ilor * ndop + j], cut_fine);
                        if (2 * half / stride + 1 < min_span) continue;
                        // first cell whose line can select Doppler sample j (approximate: any
                        // split point is correct, both kernels test the same number)
                        double cell = j == 0 ? 0.0
                                             : (e->dop_thr_h[j] / I.adop - own0) / ownstep - 1.0;
                        if (!(cell > 0.0)) cell = 0.0;
                        if (cell < (double)onwn) {
                            I.dense_from = (int)cell;
                            some = true;
                        }
                        break;
                    }
                }
                unit_dense[u] = ok && some;
                any = any || unit_dense[u];
            }
            if (!any) dense_isos.clear();
        }
        // Minor isotopes of the main isotope's species are merged into its dense plane where
        // they select the same profile (same Lorentz sample for the unit, same Doppler sample
        // on the cell): the convolution costs the same whatever number of isotopes it sums,
        // and the gather kernels are left with the cells where the Doppler samples differ.
        // PB200_DENSE_MERGE=0 keeps every minor isotope on the gather path.
        if (!dense_isos.empty()) {
            const char *me = std::getenv("PB200_DENSE_MERGE");
            main_iso = dense_isos[0];
            for (int di : dense_isos)
                if (e->iso_gend[di] - e->iso_gbeg[di] > e->iso_gend[main_iso] - e->iso_gbeg[main_iso])
                    main_iso = di;
            if (!(me && std::strcmp(me, "0") == 0))
                for (int i = 0; i < niso && (int)minor_isos.size() < std::min(kMaxMerge - 1, kMaxMinor) &&
                                niso + ((int)minor_isos.size() + 1) * 2 * ndop <= kMaxEntries;
                     i++) {
                    if (std::find(dense_isos.begin(), dense_isos.end(), i) != dense_isos.end())
                        continue;
                    if (iso_row[i] == iso_row[main_iso] && e->iso_imol[i] == e->iso_imol[main_iso] &&
                        e->iso_gend[i] > e->iso_gbeg[i])
                        minor_isos.push_back(i);
                }
            bool any_merged = false;
            for (int u = 0; u < n_units && !minor_isos.empty(); u++) {
                const IsoUnit &M = iso_units[(size_t)u * niso + main_iso];
                if (!unit_dense[u] || M.dense_from == 0x7fffffff) continue;
                bool same = true;
                for (int mi : minor_isos)
                    same = same && iso_units[(size_t)u * niso + mi].ilor == M.ilor;
                if (!same) continue;
                for (int mi : minor_isos) {
                    IsoUnit &I = iso_units[(size_t)u * niso + mi];
                    I.dense_from = M.dense_from;
                    I.merged = 1;
                }
                unit_merged[u] = 1;
                any_merged = true;
            }
            if (!any_merged) minor_isos.clear();
        }
        if (!dense_isos.empty()) {
            // anomaly bitmasks of the ofactors in use (static per line list: built once)
            const long long words = (onwn >> 5) + 2;
            const int ndivs = (int)e->divisors.size();
            rc = e->d_dense_err.alloc(1);
            if (rc) return rc;
            PB_CUDA(cudaMemsetAsync(e->d_dense_err.p, 0, sizeof(int), st));
            const StaticView V0 = e->view();
            bool built_any = false;
            std::vector<int> bit_isos(dense_isos);
            bit_isos.insert(bit_isos.end(), minor_isos.begin(), minor_isos.end());
            for (int di : bit_isos) {
                pb200_engine::DenseIso &D = e->dense_iso[di];
                if (D.built.empty()) {
                    rc = D.abits.alloc((size_t)ndivs * (size_t)words);
                    if (rc) return rc;
                    D.built.assign(ndivs, 0);
                }
                for (int u = 0; u < n_units; u++) {
                    const int slot = units[u].aslot;
                    if (!unit_dense[u] || D.built[slot]) continue;
                    unsigned *bits = D.abits.p + (size_t)slot * (size_t)words;
                    PB_CUDA(cudaMemsetAsync(bits, 0, sizeof(unsigned) * (size_t)words, st));
                    rc = launch_anomaly_bits(st, V0, e->iso_gbeg[di], e->iso_gend[di], units[u],
                                             bits, e->d_dense_err.p);
                    if (rc) return rc;
                    e->launches++;
                    D.built[slot] = 1;
                    built_any = true;
                }
            }
            if (built_any) {
                int flag = 0;
                PB_CUDA(cudaMemcpyAsync(&flag, e->d_dense_err.p, sizeof(int),
                                        cudaMemcpyDeviceToHost, st));
                PB_CUDA(cudaStreamSynchronize(st));
                if (flag) {  // a line whose dynamic index is not cell/ofactor or one below
                    e->dense_off = true;
                    e->dense_iso.clear();
                    dense_isos.clear();
                    minor_isos.clear();
                    std::fill(unit_dense.begin(), unit_dense.end(), 0);
                    std::fill(unit_merged.begin(), unit_merged.end(), 0);
                    for (IsoUnit &I : iso_units) {
                        I.dense_from = 0x7fffffff;
                        I.merged = 0;
                    }
                }
            }
        }
        if (!dense_isos.empty()) {
            rc = e->d_kd.alloc((size_t)onwn);
            if (!rc) rc = e->d_bounds.alloc((size_t)ndop + 1);
            if (!rc && !minor_isos.empty())
                rc = e->d_kd_pool.alloc((1 + minor_isos.size()) * (size_t)onwn);
            if (rc) return rc;
            for (int u = 0; u < n_units; u++)
                if (unit_merged[u]) units[u].aslot |= kMergedUnitBit;
        }
    }

    // Output buffer first, then chunk the strengths passes so that ksum[ntp_chunk, ngroups]
    // fits in 60% of what is left.
    const size_t out_bytes = sizeof(double) * (size_t)n_units * nrows * (size_t)nwave;
    double *d_out = out_dev;
    if (!out_dev) {
        rc = e->d_out.alloc((size_t)n_units * nrows * (size_t)nwave);
        if (rc) return rc;
        d_out = e->d_out.p;
    }
    const size_t per_tp = sizeof(double) * (size_t)std::max<int64_t>(e->ngroups, 1);
    int tp_chunk = ntp;
    if (e->d_ksum.n * sizeof(double) < per_tp * (size_t)ntp) {
        // the strengths scratch has to grow: size it against what is free right now
        size_t free_b = 0, total_b = 0;
        PB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t budget = (size_t)(0.6 * (double)(free_b + sizeof(double) * e->d_ksum.n));
        tp_chunk = (int)std::min<size_t>((size_t)ntp, std::max<size_t>(1, budget / per_tp));
    }
    if (tp_chunk > 65535) tp_chunk = 65535;
    {
        // PB200_TP_CHUNK=<n> forces the strengths passes into chunks of n (the GPU tests compare
        // a chunked batch with an unchunked one bit for bit)
        const char *env = std::getenv("PB200_TP_CHUNK");
        if (env && std::atoi(env) >= 1) tp_chunk = std::min(tp_chunk, std::atoi(env));
    }

    rc = e->d_ksum.alloc((size_t)tp_chunk * (size_t)std::max<int64_t>(e->ngroups, 1));
    if (!rc) rc = e->d_kmax.alloc((size_t)tp_chunk * nrows);
    // staging block layout (every section 16-byte aligned)
    auto align16 = [](size_t n) { return (n + 15) & ~(size_t)15; };
    const size_t off_row = 0;
    const size_t off_invt = off_row + align16(sizeof(int) * niso);
    const size_t off_invz = off_invt + align16(sizeof(double2) * tp_chunk);
    const size_t off_units = off_invz + align16(sizeof(double2) * (size_t)tp_chunk * niso);
    const size_t off_iso = off_units + align16(sizeof(UnitParams) * (size_t)n_units);
    const size_t stage_bytes = off_iso + align16(sizeof(IsoUnit) * (size_t)n_units * niso);
    if (!rc) rc = e->d_stage.alloc(stage_bytes);
    if (!rc && e->h_stage_bytes < stage_bytes) {
        if (e->h_stage) cudaFreeHost(e->h_stage);
        e->h_stage = nullptr;
        e->h_stage_bytes = 0;
        PB_CUDA(cudaHostAlloc((void **)&e->h_stage, stage_bytes, cudaHostAllocDefault));
        e->h_stage_bytes = stage_bytes;
    }
    if (!rc && counters) rc = e->d_counters.alloc((size_t)n_units * 4);
    if (rc) return rc;

    if (user_stream) {
        // order after work already queued on the caller's stream
        PB_CUDA(cudaEventRecord(e->ev[5], user_stream));
        PB_CUDA(cudaStreamWaitEvent(st, e->ev[5], 0));
    }
    PB_CUDA(cudaEventRecord(e->ev[0], st));
    if (counters) PB_CUDA(cudaMemsetAsync(e->d_counters.p, 0, sizeof(unsigned long long) * n_units * 4, st));

    const StaticView V = e->view();
    // Tile split: when a unit has far fewer tiles than there are resident CTAs (4 per SM),
    // several CTAs share a tile (each takes every ksplit-th chunk of the candidate lists) so
    // that fewer units are in flight at once and their Voigt profiles stay longer in L2.
    // Measured at configs[1] with the chunk kernel and 512-wide tiles (36 tiles): ksplit
    // 2/4/8/16 -> 2.68/2.31/2.26/2.45 ms.
    if (e->sm_count == 0) cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, e->device);
    int ksplit = 1;
    {
        const long long ntiles = (nwave + kChunkTile - 1) / kChunkTile;
        const long long resident = 4LL * std::max(e->sm_count, 1);
        ksplit = (int)std::max<long long>(1, std::min<long long>(8, resident / std::max<long long>(2 * ntiles * nrows, 1)));
        // ...but only while every warp keeps >= 4 chunks of 32 groups (sparse lists: the fixed
        // cost per CTA dominates; 1e5 lines, 512-wide tiles: ksplit 1/2/4 -> 0.55/0.50/0.51 ms)
        const long long chunks_per_tile = e->ngroups / std::max<long long>(ntiles, 1) / 32;
        ksplit = (int)std::max<long long>(1, std::min<long long>(ksplit, chunks_per_tile / 32));
        const char *env = std::getenv("PB200_KSPLIT");
        if (env && std::atoi(env) >= 1) ksplit = std::min(64, std::atoi(env));
    }
    // Constant-step grids run the chunk-owned kernel (32-bit table offsets: the output-stride
    // table plus the output grid must stay below 2^31 samples); PB200_ACC_KERNEL=owner keeps
    // the output-owned kernel (the GPU tests compare the two).
    int chunked = ((long long)e->d_tprofile.n + nwave < 0x7fffffffLL) ? 1 : 0;
    {
        const char *env = std::getenv("PB200_ACC_KERNEL");
        if (env && std::strcmp(env, "owner") == 0) chunked = 0;
    }
    // Units per accumulate launch: grid.y is limited to 65535, and the partial spectra of a
    // split launch (ksplit > 1) to 1 GiB; the scratch is sized ONCE per batch for the largest
    // launch (growing it between launches would free a buffer that queued kernels still use).
    size_t max_nu = 65535;
    if (ksplit > 1) {
        const size_t per_unit = sizeof(double) * (size_t)nrows * ksplit * (size_t)nwave;
        max_nu = std::max<size_t>(1, std::min<size_t>(max_nu, ((size_t)1 << 30) / per_unit));
        rc = e->d_partial.alloc(std::min<size_t>((size_t)n_units, max_nu) * nrows * ksplit *
                                (size_t)nwave);
        if (rc) return rc;
    }
    {
        const char *env = std::getenv("PB200_MAX_UNITS_PER_LAUNCH");  // tests: force launch splits
        if (env && std::atoi(env) >= 1) max_nu = std::min<size_t>(max_nu, (size_t)std::atoi(env));
    }
    // The half-batch D2H overlap below needs page-locked memory: cudaMemcpyAsync into pageable
    // memory blocks the host until the copy is done, which would delay the second half.
    bool out_pinned = false;
    if (out_host) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, out_host) == cudaSuccess)
            out_pinned = attr.type == cudaMemoryTypeHost;
        else
            cudaGetLastError();
    }
    float ms_strengths = 0.f, ms_accum = 0.f;
    size_t copied_rows = 0, split_after = 0;  // rows already sent to the host by the copy stream
    // order units by strengths pass
    std::vector<int> order(n_units);
    for (int u = 0; u < n_units; u++) order[u] = u;
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return unit_tp[a] < unit_tp[b]; });
    size_t pos = 0;
    for (int tp0 = 0; tp0 < ntp; tp0 += tp_chunk) {
        const int ntc = std::min(tp_chunk, ntp - tp0);
        std::vector<UnitParams> cu;
        std::vector<IsoUnit> ci;
        std::vector<int> cmode;   // accumulate mode | 16 for units on the dense path
        {
            std::vector<int> members;
            while (pos < order.size() && unit_tp[order[pos]] < tp0 + ntc) members.push_back(order[pos++]);
            // units of a launch share mode, dense flag and (dynamic-grid units) the ofactor
            auto key = [&](int u) {
                return unit_mode[u] | (unit_dense[u] ? 16 : 0) |
                       (unit_mode[u] == kModeDynGrid ? units[u].ofactor << 8 : 0);
            };
            std::stable_sort(members.begin(), members.end(),
                             [&](int a, int b) { return key(a) < key(b); });
            for (int u : members) {
                UnitParams U = units[u];
                U.tpass = unit_tp[u] - tp0;
                if (unit_mode[u] == kModeDynGrid) {
                    // the unit as a constant-step problem on its own dynamic grid
                    U.scale = 1;
                    U.fd_scale.set(1);
                    U.mcount = U.dnwn;
                    U.aslot = U.out_index;      // real output row, for linterp_rows_kernel
                }
                cu.push_back(U);
                cmode.push_back(key(u));
                ci.insert(ci.end(), iso_units.begin() + (size_t)u * niso,
                          iso_units.begin() + (size_t)(u + 1) * niso);
            }
        }
        std::memcpy(e->h_stage + off_row, iso_row.data(), sizeof(int) * niso);
        std::memcpy(e->h_stage + off_invt, tp_t.data() + tp0, sizeof(double2) * ntc);
        std::memcpy(e->h_stage + off_invz, tp_z.data() + (size_t)tp0 * niso,
                    sizeof(double2) * ntc * niso);
        std::memcpy(e->h_stage + off_units, cu.data(), sizeof(UnitParams) * cu.size());
        std::memcpy(e->h_stage + off_iso, ci.data(), sizeof(IsoUnit) * ci.size());
        PB_CUDA(cudaMemcpyAsync(e->d_stage.p, e->h_stage, stage_bytes, cudaMemcpyHostToDevice, st));
        const int *p_iso_row = reinterpret_cast<const int *>(e->d_stage.p + off_row);
        const double2 *p_inv_t = reinterpret_cast<const double2 *>(e->d_stage.p + off_invt);
        const double2 *p_inv_z = reinterpret_cast<const double2 *>(e->d_stage.p + off_invz);
        const UnitParams *p_units = reinterpret_cast<const UnitParams *>(e->d_stage.p + off_units);
        const IsoUnit *p_iso_units = reinterpret_cast<const IsoUnit *>(e->d_stage.p + off_iso);
        PB_CUDA(cudaMemsetAsync(e->d_kmax.p, 0, sizeof(unsigned long long) * ntc * nrows, st));
        PB_CUDA(cudaEventRecord(e->ev[1], st));
        rc = launch_strengths(st, V, ntc, p_inv_t, p_inv_z, p_iso_row, nrows,
                              e->d_ksum.p, e->d_kmax.p, e->d_lgroup.p, e->d_liso.p, e->n_inwin);
        if (rc) return rc;
        if (V.ngroups > 0) e->launches++;
        PB_CUDA(cudaEventRecord(e->ev[2], st));
        // merged minor isotopes: the gather kernels need the main isotope's Doppler segments
        // of every strengths pass of this chunk
        const int *p_main_bounds = nullptr;
        MergeView merge;
        if (!minor_isos.empty()) {
            const size_t nminor = minor_isos.size();
            rc = e->d_bounds_main.alloc((size_t)ntc * (ndop + 1));
            if (!rc) rc = e->d_bounds_minor.alloc((size_t)ntc * nminor * (ndop + 1));
            if (rc) return rc;
            std::vector<int> pass_unit(ntc, -1);      // any unit of the pass (adop depends on T only)
            for (size_t a = 0; a < cu.size(); a++) pass_unit[cu[a].tpass] = (int)a;
            for (int t = 0; t < ntc; t++) {
                if (pass_unit[t] < 0) continue;
                const IsoUnit *iu = ci.data() + (size_t)pass_unit[t] * niso;
                rc = launch_segment_bounds(st, V, e->iso_gbeg[main_iso], e->iso_gend[main_iso],
                                           iu[main_iso].adop,
                                           e->d_bounds_main.p + (size_t)t * (ndop + 1));
                for (size_t j = 0; j < nminor && !rc; j++) {
                    const int mi = minor_isos[j];
                    rc = launch_segment_bounds(st, V, e->iso_gbeg[mi], e->iso_gend[mi], iu[mi].adop,
                                               e->d_bounds_minor.p + ((size_t)t * nminor + j) * (ndop + 1));
                }
                if (rc) return rc;
                e->launches += 1 + (int64_t)nminor;
            }
            p_main_bounds = e->d_bounds_main.p;
            merge.nminor = (int)nminor;
            for (size_t j = 0; j < nminor; j++) merge.iso[j] = minor_isos[j];
            merge.main_bounds = e->d_bounds_main.p;
            merge.minor_bounds = e->d_bounds_minor.p;
        }
        // one launch per run of equal mode; grid.y is limited to 65535 units per launch
        for (size_t u0 = 0; u0 < cu.size();) {
            size_t u1 = u0;
            while (u1 < cu.size() && cmode[u1] == cmode[u0] && u1 - u0 < max_nu) u1++;
            // Host output: run the first half of the batch on its own, so that its rows travel
            // to the host (copy stream) while the second half is computed.  Only when the
            // first half is exactly the rows [0, half) of the caller's array.
            if (out_host && out_pinned && !counters && dense_isos.empty() && u0 == 0 && tp0 == 0 &&
                ntc == ntp &&
                u1 == cu.size() &&
                copied_rows == 0 && u1 >= 16) {
                const size_t half = u1 / 2;
                bool prefix = true;
                for (size_t u = 0; u < u1 && prefix; u++)
                    prefix = (u < half) == (cu[u].out_index < (int)half);
                if (prefix) {
                    u1 = half;
                    split_after = half;
                }
            }
            if ((cmode[u0] & 15) == kModeDynGrid) {
                // dynamic-grid units of one ofactor: output-stride table for that stride, chunk
                // kernel on the dynamic grid into scratch rows, 2-point interpolation to the output
                const int of = cu[u0].ofactor, dn = cu[u0].dnwn;
                rc = ensure_transposed(e, of);
                if (rc) return rc;
                StaticView V2 = e->view();
                V2.nwave = dn;
                if (counters) {   // first: the scratch launches renumber out_index on the device
                    rc = launch_counters(st, V, (int)(u1 - u0), p_units + u0, p_iso_units + u0 * niso,
                                         p_iso_row, e->d_ksum.p, e->d_kmax.p, nrows, ethresh,
                                         cutoff, 1, e->d_counters.p);
                    if (rc) return rc;
                    if (V.ngroups > 0) e->launches++;
                }
                const size_t row_doubles = (size_t)nrows * (size_t)dn;
                const size_t per_launch = std::max<size_t>(1, ((size_t)1 << 28) / row_doubles);  // 2 GiB
                for (size_t a0 = u0; a0 < u1; a0 += per_launch) {
                    const size_t a1 = std::min(u1, a0 + per_launch);
                    const int nd = (int)(a1 - a0);
                    rc = e->d_dyn.alloc((size_t)nd * row_doubles);
                    if (rc) return rc;
                    // rows of the scratch are numbered within the launch
                    for (size_t a = a0; a < a1; a++) cu[a].out_index = (int)(a - a0);
                    PB_CUDA(cudaMemcpyAsync(e->d_stage.p + off_units + sizeof(UnitParams) * a0,
                                            cu.data() + a0, sizeof(UnitParams) * nd,
                                            cudaMemcpyHostToDevice, st));
                    rc = launch_accumulate(st, V2, nd, p_units + a0, p_iso_units + a0 * niso,
                                           p_iso_row, e->d_ksum.p, e->d_kmax.p, nrows, ethresh,
                                           cutoff, kModeTransposed, e->d_dyn.p, 1, nullptr, 1);
                    if (!rc)
                        rc = launch_linterp_rows(st, V, nd, p_units + a0, e->d_dyn.p, dn, nrows,
                                                 d_out);
                    if (rc) return rc;
                    e->launches += 2;
                }
                u0 = u1;
                continue;
            }
            const int nu = (int)(u1 - u0);
            rc = launch_accumulate(st, V, nu, p_units + u0, p_iso_units + u0 * niso,
                                   p_iso_row, e->d_ksum.p,
                                   e->d_kmax.p, nrows, ethresh, cutoff, cmode[u0] & 15, d_out,
                                   ksplit, e->d_partial.p, chunked, merge);
            if (rc) return rc;
            e->launches += ksplit > 1 ? 2 : 1;
            if (counters) {
                rc = launch_counters(st, V, nu, p_units + u0, p_iso_units + u0 * niso,
                                     p_iso_row, e->d_ksum.p, e->d_kmax.p, nrows, ethresh,
                                     cutoff, resolution ? 1 : 0, e->d_counters.p);
                if (rc) return rc;
                if (V.ngroups > 0) e->launches++;
            }
            if (split_after && u1 == split_after && copied_rows == 0) {
                PB_CUDA(cudaEventRecord(e->ev_half, st));
                PB_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_half, 0));
                copied_rows = split_after;
                PB_CUDA(cudaMemcpyAsync(out_host, d_out,
                                        sizeof(double) * copied_rows * nrows * (size_t)nwave,
                                        cudaMemcpyDeviceToHost, e->copy_stream));
            }
            u0 = u1;
        }
        // Dense path: per strengths pass and dense isotope, scatter the group strengths onto the
        // fine grid, find the Doppler segments, and convolve for all units of the pass at once.
        if (!dense_isos.empty()) PB_CUDA(cudaEventRecord(e->ev_dense[0], st));
        for (size_t a = 0; a < cu.size() && !dense_isos.empty();) {
            if (!(cmode[a] & 16)) {
                a++;
                continue;
            }
            size_t b = a;
            while (b < cu.size() && (cmode[b] & 16) && cu[b].tpass == cu[a].tpass) b++;
            const long long words = (onwn >> 5) + 2;
            const int tpass = cu[a].tpass;
            const double *ks_t = e->d_ksum.p + (size_t)tpass * (size_t)e->ngroups;
            for (int di : dense_isos) {
                const int row = iso_row[di];
                const unsigned long long *kmax_t = e->d_kmax.p + (size_t)tpass * nrows + row;
                PB_CUDA(cudaMemsetAsync(e->d_kd.p, 0, sizeof(double) * (size_t)onwn, st));
                rc = launch_densify(st, V, e->iso_gbeg[di], e->iso_gend[di], ks_t, kmax_t, ethresh,
                                    e->d_kd.p);
                if (rc) return rc;
                DenseSet set;
                set.n = 1;
                set.iso[0] = di;
                set.kd[0] = e->d_kd.p;
                set.abits[0] = e->dense_iso[di].abits.p;
                set.kd_all = e->d_kd.p;
                set.bounds = e->d_bounds.p;
                if (di == main_iso && !minor_isos.empty()) {
                    // sum plane = main plane + the minor isotopes' groups that share its profile
                    const int *mb = p_main_bounds + (size_t)tpass * (ndop + 1);
                    double *pool = e->d_kd_pool.p;
                    PB_CUDA(cudaMemcpyAsync(pool, e->d_kd.p, sizeof(double) * (size_t)onwn,
                                            cudaMemcpyDeviceToDevice, st));
                    PB_CUDA(cudaMemsetAsync(pool + onwn, 0,
                                            sizeof(double) * minor_isos.size() * (size_t)onwn, st));
                    for (size_t j = 0; j < minor_isos.size(); j++) {
                        const int mi = minor_isos[j];
                        rc = launch_merge_minor(st, V, e->iso_gbeg[mi], e->iso_gend[mi], ks_t, kmax_t,
                                                ethresh, ci[a * niso + mi].adop, mb,
                                                pool + (j + 1) * (size_t)onwn, pool);
                        if (rc) return rc;
                        e->launches++;
                        set.iso[set.n] = mi;
                        set.kd[set.n] = pool + (j + 1) * (size_t)onwn;
                        set.abits[set.n] = e->dense_iso[mi].abits.p;
                        set.n++;
                    }
                    set.kd_all = pool;
                    set.bounds = mb;
                } else {
                    rc = launch_segment_bounds(st, V, e->iso_gbeg[di], e->iso_gend[di],
                                               ci[a * niso + di].adop, e->d_bounds.p);
                    if (rc) return rc;
                    e->launches++;
                }
                rc = launch_accumulate_dense(st, V, (int)(b - a), p_units + a,
                                             p_iso_units + a * niso, set, row, nrows, words,
                                             cutoff, d_out, e->d_dense_err.p);
                if (rc) return rc;
                e->launches += 2;
                e->dense_unit_isos += (int64_t)(b - a) * set.n;
            }
            a = b;
        }
        if (!dense_isos.empty()) PB_CUDA(cudaEventRecord(e->ev_dense[1], st));
        PB_CUDA(cudaEventRecord(e->ev[3], st));
        // the host vectors cu/ci must stay alive until their copies have completed
        PB_CUDA(cudaStreamSynchronize(st));
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, e->ev[1], e->ev[2]);
        cudaEventElapsedTime(&b, e->ev[2], e->ev[3]);
        ms_strengths += a;
        ms_accum += b;
        if (!dense_isos.empty()) {
            float d = 0.f;
            cudaEventElapsedTime(&d, e->ev_dense[0], e->ev_dense[1]);
            e->dense_ms += d;
        }
    }
    PB_CUDA(cudaEventRecord(e->ev[3], st));
    if (out_host) {
        const size_t done = sizeof(double) * copied_rows * nrows * (size_t)nwave;
        PB_CUDA(cudaMemcpyAsync((char *)out_host + done, (const char *)d_out + done,
                                out_bytes - done, cudaMemcpyDeviceToHost, st));
    }
    std::vector<unsigned long long> cnt;
    if (counters) {
        cnt.resize((size_t)n_units * 4);
        PB_CUDA(cudaMemcpyAsync(cnt.data(), e->d_counters.p, sizeof(unsigned long long) * cnt.size(),
                                cudaMemcpyDeviceToHost, st));
    }
    PB_CUDA(cudaEventRecord(e->ev[4], st));
    if (user_stream && out_dev) {
        PB_CUDA(cudaStreamWaitEvent(user_stream, e->ev[4], 0));
    }
    int dense_flag = 0;
    if (!dense_isos.empty())
        PB_CUDA(cudaMemcpyAsync(&dense_flag, e->d_dense_err.p, sizeof(int), cudaMemcpyDeviceToHost,
                                st));
    PB_CUDA(cudaStreamSynchronize(st));
    if (copied_rows) PB_CUDA(cudaStreamSynchronize(e->copy_stream));
    if (dense_flag)
        return fail(PB200_ECUDA, "dense accumulate path: a line footprint exceeded the shared "
                                 "tiles (internal sizing error; rerun with PB200_DENSE=0)");
    if (counters) {
        // nadd is static per isotope: lines absorbed into a head line of a processed isotope
        int64_t nadd = 0;
        for (int i = 0; i < niso; i++)
            if (iso_row[i] >= 0) nadd += e->iso_nadd[i];
        for (int u = 0; u < n_units; u++) {
            counters[(size_t)u * 6 + 0] = nadd;
            counters[(size_t)u * 6 + 1] = (int64_t)cnt[(size_t)u * 4 + 0];
            counters[(size_t)u * 6 + 2] = (int64_t)cnt[(size_t)u * 4 + 1];
            counters[(size_t)u * 6 + 3] = (int64_t)cnt[(size_t)u * 4 + 2];
            counters[(size_t)u * 6 + 4] = (int64_t)cnt[(size_t)u * 4 + 3];
            counters[(size_t)u * 6 + 5] = (int64_t)unit_table_bytes[u];
        }
    }
    float t_all = 0.f, t_d2h = 0.f;
    cudaEventElapsedTime(&t_all, e->ev[0], e->ev[4]);
    cudaEventElapsedTime(&t_d2h, e->ev[3], e->ev[4]);
    e->timing[0] = ms_strengths;
    e->timing[1] = ms_accum;
    e->timing[2] = t_all - ms_strengths - ms_accum - t_d2h;
    e->timing[3] = t_d2h;
    e->timing[4] = t_all;
    return 0;
}

int pb200_extinction_batch_host(pb200_engine *e, int n_units, const double *unit_temp,
                                const double *unit_density, const double *unit_isoz,
                                const int64_t *iso_iext, int nextinct, double ethresh, int add,
                                int resolution, double *out, int64_t *counters) {
    if (!out) return fail(PB200_EINVAL, "pb200_extinction_batch_host: null output");
    return run_batch(e, n_units, unit_temp, unit_density, unit_isoz, iso_iext, nextinct,
                     ethresh, add, resolution, out, nullptr, counters, nullptr);
}

int pb200_extinction_batch_dev(pb200_engine *e, int n_units, const double *unit_temp,
                               const double *unit_density, const double *unit_isoz,
                               const int64_t *iso_iext, int nextinct, double ethresh, int add,
                               int resolution, double *out_dev, int64_t *counters,
                               void *cuda_stream) {
    if (!out_dev) return fail(PB200_EINVAL, "pb200_extinction_batch_dev: null output");
    return run_batch(e, n_units, unit_temp, unit_density, unit_isoz, iso_iext, nextinct,
                     ethresh, add, resolution, nullptr, out_dev, counters,
                     (cudaStream_t)cuda_stream);
}

int pb200_engine_last_timing(const pb200_engine *e, double ms[5]) {
    if (!e || !ms) return fail(PB200_EINVAL, "pb200_engine_last_timing: null argument");
    for (int i = 0; i < 5; i++) ms[i] = e->timing[i];
    return 0;
}

int64_t pb200_engine_launch_count(const pb200_engine *e) { return e ? e->launches : 0; }

int64_t pb200_engine_dense_units(const pb200_engine *e) { return e ? e->dense_unit_isos : 0; }

double pb200_engine_dense_ms(const pb200_engine *e) { return e ? e->dense_ms : 0.0; }

void *pb200_engine_stream(const pb200_engine *e) { return e ? (void *)e->stream : nullptr; }

// ----------------------------------------------------------------------------------------
static int interp_impl(int device, double *ext, const double *etable, bool on_device,
                       const double *ttable, const double *temperature, const double *density,
                       int nspec, int ntemp, int nlayers, int nwave, int lay1, int lay2,
                       int per_mol, cudaStream_t user_stream) {
    if (!ext || !etable || !ttable || !temperature || !density || nspec < 1 || ntemp < 2 ||
        nlayers < 1 || nwave < 1)
        return fail(PB200_EINVAL, "pb200_interp_ec: null argument or ntemp<2");
    int rc = select_device(device);
    if (rc) return rc;
    if (lay2 > nlayers) lay2 = nlayers;  // :389
    if (lay1 < 0) return fail(PB200_EINVAL, "pb200_interp_ec: lay1 < 0");
    if (lay2 <= lay1) return 0;
    // Host: bracketing temperature and weights per layer (:392-402)
    std::vector<int> tlo(nlayers, 0);
    std::vector<double> w_lo(nlayers, 0.0), w_hi(nlayers, 0.0);
    for (int k = lay1; k < lay2; k++) {
        const double t = temperature[k];
        int lo = nearest_bisect(ttable, t, 0, ntemp - 1);
        if (t < ttable[lo] || lo == ntemp - 1) lo--;
        if (lo < 0) lo = 0;  // the reference would index ttable[-1]; clamp instead
        const int hi = lo + 1;
        tlo[k] = lo;
        w_lo[k] = (ttable[hi] - t) / (ttable[hi] - ttable[lo]);
        w_hi[k] = (t - ttable[lo]) / (ttable[hi] - ttable[lo]);
    }
    cudaStream_t st = user_stream;
    bool own_stream = false;
    if (!st) {
        PB_CUDA(cudaStreamCreate(&st));
        own_stream = true;
    }
    // Per-layer scalars in ONE stream-ordered allocation (no device-wide sync on the
    // device-resident path): [w_lo | w_hi | density | tlo].
    DevBuf<double> d_tab, d_ext;
    const size_t tab_n = (size_t)nspec * ntemp * nlayers * (size_t)nwave;
    const size_t ext_n = (size_t)(per_mol ? nspec : 1) * nlayers * (size_t)nwave;
    const size_t ndbl = (size_t)nlayers * (2 + nspec);
    std::vector<double> packed(ndbl + ((size_t)nlayers + 1) / 2);
    std::copy(w_lo.begin(), w_lo.end(), packed.begin());
    std::copy(w_hi.begin(), w_hi.end(), packed.begin() + nlayers);
    std::copy(density, density + (size_t)nlayers * nspec, packed.begin() + 2 * (size_t)nlayers);
    std::memcpy(packed.data() + ndbl, tlo.data(), sizeof(int) * nlayers);
    // (a persistent per-device scratch block: a stream-ordered allocation freed before the
    // synchronize returned its memory to the OS on every call, 0.5 ms for a 69 us kernel)
    double *d_small = nullptr;
    static std::mutex mu;  // held until the call has synchronised: the block is shared
    std::lock_guard<std::mutex> lock(mu);
    {
        static std::map<int, std::pair<double *, size_t>> scratch;  // device -> (block, doubles)
        auto &slot = scratch[device];
        cudaError_t ea = cudaSuccess;
        if (slot.second < packed.size()) {
            if (slot.first) cudaFree(slot.first);
            slot = {nullptr, 0};
            ea = cudaMalloc((void **)&slot.first, sizeof(double) * packed.size());
            if (ea == cudaSuccess) slot.second = packed.size();
        }
        d_small = slot.first;
        if (ea == cudaSuccess)
            ea = cudaMemcpyAsync(d_small, packed.data(), sizeof(double) * packed.size(),
                                 cudaMemcpyHostToDevice, st);
        if (ea != cudaSuccess) rc = cuda_fail(ea, "interp scalars", __FILE__, __LINE__);
    }
    const double *tab = etable;
    double *dext = ext;
    if (!on_device) {
        if (!rc) rc = d_tab.upload(etable, tab_n, st);
        if (!rc) rc = d_ext.upload(ext, ext_n, st);
        tab = d_tab.p;
        dext = d_ext.p;
    }
    if (!rc)
        rc = launch_interp_ec(st, dext, tab, (const int *)(d_small + ndbl), d_small,
                              d_small + nlayers, d_small + 2 * (size_t)nlayers, nspec, ntemp,
                              nlayers, nwave, lay1, lay2, per_mol);
    if (!rc && !on_device) {
        cudaError_t e2 = cudaMemcpyAsync(ext, dext, sizeof(double) * ext_n,
                                         cudaMemcpyDeviceToHost, st);
        if (e2 != cudaSuccess) rc = cuda_fail(e2, "interp d2h", __FILE__, __LINE__);
    }
    cudaError_t e3 = cudaStreamSynchronize(st);
    if (!rc && e3 != cudaSuccess) rc = cuda_fail(e3, "interp sync", __FILE__, __LINE__);
    if (own_stream) cudaStreamDestroy(st);
    return rc;
}

int pb200_interp_ec(int device, double *ext, const double *etable, const double *ttable,
                    const double *temperature, const double *density, int nspec, int ntemp,
                    int nlayers, int nwave, int lay1, int lay2) {
    return interp_impl(device, ext, etable, false, ttable, temperature, density, nspec, ntemp,
                       nlayers, nwave, lay1, lay2, 0, nullptr);
}

int pb200_interp_ec_per_mol(int device, double *ext, const double *etable,
                            const double *ttable, const double *temperature,
                            const double *density, int nspec, int ntemp, int nlayers, int nwave,
                            int lay1, int lay2) {
    return interp_impl(device, ext, etable, false, ttable, temperature, density, nspec, ntemp,
                       nlayers, nwave, lay1, lay2, 1, nullptr);
}

int pb200_interp_ec_dev(int device, double *ext_dev, const double *etable_dev,
                        const double *ttable, const double *temperature, const double *density,
                        int nspec, int ntemp, int nlayers, int nwave, int lay1, int lay2,
                        int per_mol, void *cuda_stream) {
    return interp_impl(device, ext_dev, etable_dev, true, ttable, temperature, density, nspec,
                       ntemp, nlayers, nwave, lay1, lay2, per_mol, (cudaStream_t)cuda_stream);
}

}  // extern "C"

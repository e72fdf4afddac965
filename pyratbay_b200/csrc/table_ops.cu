// table_ops.cu -- device-resident consumers of a cross-section table (the Line_Sample side):
//
//  * pb200_table_*: a handle over a table [nspec, ntemp, nlayers, nwave] that lives in HBM;
//    pb200_table_interp evaluates _extcoeff.c:367-472 (interp_ec / interp_ec_per_mol) with the
//    interp_ec kernel of lbl_kernels.cu WITHOUT any per-call allocation, lock or stream
//    synchronisation: the per-layer scalars travel through a ring of pinned staging slots and
//    the call returns as soon as the kernel is queued.
//  * pb200_regrid_table_dev: the p/T re-gridding of tools/tools.py:1026-1107
//    (interpolate_opacity): piecewise-linear in log(cross section) over log p, then over T,
//    log of non-positive values floored at -230, edge values outside the tabulated range.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/pb200_lbl.h"
#include "lbl_kernels.cuh"

namespace pb200 {

// out[t, p, w] (+)= exp(lerp_T(lerp_p(log table))) ; one thread per output sample.
__global__ void __launch_bounds__(256)
regrid_table_kernel(const double *__restrict__ table, int nlayers, int nwave,
                    const int *__restrict__ t_lo, const int *__restrict__ t_hi,
                    const double *__restrict__ t_f, const int *__restrict__ p_lo,
                    const int *__restrict__ p_hi, const double *__restrict__ p_f,
                    int nlayers_out, const int *__restrict__ wave_idx, int nwave_out,
                    int take_log, double *__restrict__ out, int accumulate) {
    const int iw = blockIdx.x * blockDim.x + threadIdx.x;
    if (iw >= nwave_out) return;
    const int ip = blockIdx.y, it = blockIdx.z;
    const int src = wave_idx ? wave_idx[iw] : iw;
    const int tl = t_lo[it], th = t_hi[it], pl = p_lo[ip], ph = p_hi[ip];
    const double ft = t_f[it], fp = p_f[ip];
    auto at = [&](int t, int p) {
        const double v = table[((size_t)t * nlayers + p) * (size_t)nwave + src];
        if (!take_log) return v;
        const double l = log(v);
        return isfinite(l) ? l : -230.0;                      // tools.py:1079-1080
    };
    double res;
    if (take_log) {
        const double a0 = at(tl, pl), a1 = at(tl, ph);
        const double a = dadd(a0, dmul(dsub(a1, a0), fp));    // over log p at the lower T
        double b = a;
        if (th != tl) {
            const double b0 = at(th, pl), b1 = at(th, ph);
            b = dadd(b0, dmul(dsub(b1, b0), fp));
        }
        res = exp(dadd(a, dmul(dsub(b, a), ft)));             // over T, back to linear
    } else {
        res = at(tl, pl);                                     // same grids: plain copy
    }
    double *dst = out + ((size_t)it * nlayers_out + ip) * (size_t)nwave_out + iw;
    *dst = accumulate ? dadd(*dst, res) : res;
}

}  // namespace pb200

using namespace pb200;

struct pb200_table {
    int device = 0;
    int nspec = 0, ntemp = 0, nlayers = 0, nwave = 0;
    const double *etable = nullptr;   // borrowed device pointer
    std::vector<double> ttable;
    static constexpr int kSlots = 8;
    size_t slot_doubles = 0;
    double *h_stage = nullptr, *d_stage = nullptr;
    cudaEvent_t ev[kSlots] = {};
    bool ev_used[kSlots] = {};
    int next = 0;
    int64_t launches = 0;
};

static int table_fail(int code, const char *msg) {
    set_error(msg);
    return code;
}

extern "C" {

int pb200_table_create(int device, const double *etable_dev, const double *ttable, int nspec,
                       int ntemp, int nlayers, int nwave, pb200_table **out) {
    if (!out) return table_fail(PB200_EINVAL, "pb200_table_create: null output");
    *out = nullptr;
    if (!etable_dev || !ttable || nspec < 1 || ntemp < 2 || nlayers < 1 || nwave < 1)
        return table_fail(PB200_EINVAL, "pb200_table_create: null argument or ntemp<2");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return table_fail(PB200_ENODEVICE, "no usable CUDA device: this engine has no CPU fallback");
    }
    if (device < 0 || device >= n) return table_fail(PB200_EINVAL, "device index out of range");
    PB_CUDA(cudaSetDevice(device));
    pb200_table *t = new pb200_table();
    t->device = device;
    t->nspec = nspec; t->ntemp = ntemp; t->nlayers = nlayers; t->nwave = nwave;
    t->etable = etable_dev;
    t->ttable.assign(ttable, ttable + ntemp);
    // slot: [w_lo | w_hi | density[nlayers, nspec] | tlo (ints)] , 16-byte multiple
    t->slot_doubles = ((size_t)nlayers * (2 + nspec) + ((size_t)nlayers + 1) / 2 + 1) & ~(size_t)1;
    cudaError_t e = cudaHostAlloc((void **)&t->h_stage,
                                  sizeof(double) * t->slot_doubles * pb200_table::kSlots,
                                  cudaHostAllocDefault);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&t->d_stage, sizeof(double) * t->slot_doubles * pb200_table::kSlots);
    for (int i = 0; i < pb200_table::kSlots && e == cudaSuccess; i++)
        e = cudaEventCreateWithFlags(&t->ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        pb200_table_destroy(t);
        return cuda_fail(e, "pb200_table_create", __FILE__, __LINE__);
    }
    *out = t;
    return 0;
}

void pb200_table_destroy(pb200_table *t) {
    if (!t) return;
    cudaSetDevice(t->device);
    for (int i = 0; i < pb200_table::kSlots; i++)
        if (t->ev[i]) {
            if (t->ev_used[i]) cudaEventSynchronize(t->ev[i]);
            cudaEventDestroy(t->ev[i]);
        }
    if (t->h_stage) cudaFreeHost(t->h_stage);
    if (t->d_stage) cudaFree(t->d_stage);
    delete t;
}

int pb200_table_interp(pb200_table *t, const double *temperature, const double *density,
                       int lay1, int lay2, int per_mol, double *ext_dev, int overwrite,
                       void *cuda_stream, int sync) {
    if (!t || !temperature || !density || !ext_dev)
        return table_fail(PB200_EINVAL, "pb200_table_interp: null argument");
    const int nlayers = t->nlayers, nspec = t->nspec, ntemp = t->ntemp;
    if (lay2 > nlayers) lay2 = nlayers;  // :389
    if (lay1 < 0) return table_fail(PB200_EINVAL, "pb200_table_interp: lay1 < 0");
    if (lay2 <= lay1) return 0;
    PB_CUDA(cudaSetDevice(t->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // staging slot: wait only if the copy that last used it has not run yet (8 calls ago)
    const int slot = t->next;
    t->next = (t->next + 1) % pb200_table::kSlots;
    if (t->ev_used[slot]) PB_CUDA(cudaEventSynchronize(t->ev[slot]));
    double *h = t->h_stage + (size_t)slot * t->slot_doubles;
    double *d = t->d_stage + (size_t)slot * t->slot_doubles;
    double *w_lo = h, *w_hi = h + nlayers, *dens = h + 2 * (size_t)nlayers;
    int *tlo = reinterpret_cast<int *>(h + (size_t)nlayers * (2 + nspec));
    const double *tt = t->ttable.data();
    for (int k = 0; k < nlayers; k++) {
        w_lo[k] = w_hi[k] = 0.0;
        tlo[k] = 0;
    }
    for (int k = lay1; k < lay2; k++) {  // bracketing temperatures and weights (:392-402)
        const double tk = temperature[k];
        int lo = 0, hi = ntemp - 1;      // binsearchapprox (utils.h:75-89)
        while (hi - lo > 1) {
            const int mid = (hi + lo) / 2;
            if (tt[mid] > tk) hi = mid; else lo = mid;
        }
        int near = (std::fabs(tt[hi] - tk) < std::fabs(tt[lo] - tk)) ? hi : lo;
        if (tk < tt[near] || near == ntemp - 1) near--;
        if (near < 0) near = 0;          // the reference would read ttable[-1]; clamped
        tlo[k] = near;
        w_lo[k] = (tt[near + 1] - tk) / (tt[near + 1] - tt[near]);
        w_hi[k] = (tk - tt[near]) / (tt[near + 1] - tt[near]);
    }
    std::memcpy(dens, density, sizeof(double) * (size_t)nlayers * nspec);
    PB_CUDA(cudaMemcpyAsync(d, h, sizeof(double) * t->slot_doubles, cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaEventRecord(t->ev[slot], st));
    t->ev_used[slot] = true;
    if (overwrite && (lay1 > 0 || lay2 < nlayers)) {
        // rows outside [lay1, lay2) are what the reference leaves in its zeroed array
        PB_CUDA(cudaMemsetAsync(ext_dev, 0,
                                sizeof(double) * (size_t)(per_mol ? nspec : 1) * nlayers *
                                    (size_t)t->nwave, st));
    }
    int rc = launch_interp_ec(st, ext_dev, t->etable,
                              reinterpret_cast<const int *>(d + (size_t)nlayers * (2 + nspec)), d,
                              d + nlayers, d + 2 * (size_t)nlayers, nspec, ntemp, nlayers,
                              t->nwave, lay1, lay2, per_mol, overwrite);
    if (rc) return rc;
    t->launches++;
    if (sync) PB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int64_t pb200_table_launch_count(const pb200_table *t) { return t ? t->launches : 0; }

int pb200_regrid_table_dev(int device, const double *table_dev, int ntemp, int nlayers, int nwave,
                           const int *t_lo, const int *t_hi, const double *t_f, int ntemp_out,
                           const int *p_lo, const int *p_hi, const double *p_f, int nlayers_out,
                           const int *wave_idx, int nwave_out, int take_log, double *out_dev,
                           int accumulate, void *cuda_stream) {
    if (!table_dev || !out_dev || !t_lo || !t_hi || !t_f || !p_lo || !p_hi || !p_f ||
        ntemp < 1 || nlayers < 1 || nwave < 1 || ntemp_out < 1 || nlayers_out < 1 || nwave_out < 1)
        return table_fail(PB200_EINVAL, "pb200_regrid_table_dev: null or empty argument");
    if (ntemp_out > 65535 || nlayers_out > 65535)
        return table_fail(PB200_EINVAL, "pb200_regrid_table_dev: more than 65535 output T or p");
    for (int i = 0; i < ntemp_out; i++)
        if (t_lo[i] < 0 || t_hi[i] >= ntemp || t_lo[i] > t_hi[i])
            return table_fail(PB200_EINVAL, "pb200_regrid_table_dev: temperature index out of range");
    for (int i = 0; i < nlayers_out; i++)
        if (p_lo[i] < 0 || p_hi[i] >= nlayers || p_lo[i] > p_hi[i])
            return table_fail(PB200_EINVAL, "pb200_regrid_table_dev: pressure index out of range");
    if (wave_idx)
        for (int i = 0; i < nwave_out; i++)
            if (wave_idx[i] < 0 || wave_idx[i] >= nwave)
                return table_fail(PB200_EINVAL, "pb200_regrid_table_dev: wavenumber index out of range");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return table_fail(PB200_ENODEVICE, "no usable CUDA device: this engine has no CPU fallback");
    }
    PB_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // index/weight arrays in one stream-ordered scratch block
    const size_t ni = 2 * ((size_t)ntemp_out + nlayers_out) + (wave_idx ? (size_t)nwave_out : 0);
    const size_t nd = (size_t)ntemp_out + nlayers_out;
    std::vector<double> hd(nd + (ni + 1) / 2);
    std::memcpy(hd.data(), t_f, sizeof(double) * ntemp_out);
    std::memcpy(hd.data() + ntemp_out, p_f, sizeof(double) * nlayers_out);
    int *hi = reinterpret_cast<int *>(hd.data() + nd);
    std::memcpy(hi, t_lo, sizeof(int) * ntemp_out);
    std::memcpy(hi + ntemp_out, t_hi, sizeof(int) * ntemp_out);
    std::memcpy(hi + 2 * ntemp_out, p_lo, sizeof(int) * nlayers_out);
    std::memcpy(hi + 2 * ntemp_out + nlayers_out, p_hi, sizeof(int) * nlayers_out);
    if (wave_idx)
        std::memcpy(hi + 2 * ((size_t)ntemp_out + nlayers_out), wave_idx, sizeof(int) * nwave_out);
    double *dd = nullptr;
    PB_CUDA(cudaMalloc((void **)&dd, sizeof(double) * hd.size()));
    cudaError_t e = cudaMemcpyAsync(dd, hd.data(), sizeof(double) * hd.size(),
                                    cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        const int *di = reinterpret_cast<const int *>(dd + nd);
        dim3 grid((unsigned)((nwave_out + 255) / 256), (unsigned)nlayers_out, (unsigned)ntemp_out);
        regrid_table_kernel<<<grid, 256, 0, st>>>(
            table_dev, nlayers, nwave, di, di + ntemp_out, dd, di + 2 * ntemp_out,
            di + 2 * ntemp_out + nlayers_out, dd + ntemp_out, nlayers_out,
            wave_idx ? di + 2 * ((size_t)ntemp_out + nlayers_out) : nullptr, nwave_out, take_log,
            out_dev, accumulate);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // hd/dd are released below
    cudaFree(dd);
    if (e != cudaSuccess) return cuda_fail(e, "pb200_regrid_table_dev", __FILE__, __LINE__);
    return 0;
}

}  // extern "C"

// preprocess.cu -- the (T,p)-independent line pre-processing on the device.
//
// What pb200_engine_set_lines has to produce from a TLI-ordered line list (reference
// src_c/_extcoeff.c:229-262, everything that does not depend on T or p):
//   * window filter            wn < own[0] || wn > own[-1] -> ignored           (:215,239)
//   * nearest fine-grid index  iown = trunc((wn-own0)/ownstep), +1 if closer    (:243-245)
//   * greedy co-add grouping   a head line absorbs the following lines of its isotope while
//                              |wn_next - own[iown_head]| < ownstep             (:249-262)
// The grouping is a sequential chain, but a line that lies >= 1.5 fine steps above its
// predecessor can never be absorbed (own[iown_head] <= wn_head + step/2), so such lines cut the
// list into independent segments; one thread walks one segment.  For TLI-scale lists (about
// one line per fine cell) segments are a handful of lines long.  Very dense lists produce long
// segments; the caller falls back to the host walk when the longest segment is excessive.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <vector>

#include "preprocess.cuh"

namespace pb200 {

__global__ void __launch_bounds__(256)
line_flags_kernel(long long nlines, const double *__restrict__ wn,
                  const unsigned short *__restrict__ iso, const double *__restrict__ own,
                  long long onwn, double own0, double own_last, double ownstep,
                  int *__restrict__ inwin, int *__restrict__ iown, int *__restrict__ certain) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const double w = wn[l];
    const bool in = !(w < own0 || w > own_last);
    int idx = 0;
    if (in) {
        idx = (int)ddiv(dsub(w, own0), ownstep);
        if (idx + 1 < onwn && fabs(dsub(w, own[idx + 1])) < fabs(dsub(w, own[idx]))) idx++;
    }
    inwin[l] = in ? 1 : 0;
    iown[l] = idx;
    bool head = false;
    if (in) {
        if (l == 0 || iso[l - 1] != iso[l]) head = true;              // first line of a block
        else {
            const double prev = wn[l - 1];
            if (prev < own0) head = true;                              // first one in the window
            else if (dsub(w, prev) >= dmul(1.6, ownstep)) head = true; // cannot be absorbed
        }
    }
    certain[l] = head ? 1 : 0;
}

__global__ void __launch_bounds__(256)
segment_starts_kernel(long long nlines, const int *__restrict__ certain,
                      const int *__restrict__ segid, long long *__restrict__ seg_start) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l < nlines && certain[l]) seg_start[segid[l]] = l;
}

// One thread per segment: the reference's sequential co-add walk inside the segment.
__global__ void __launch_bounds__(128)
segment_walk_kernel(long long nseg, long long nlines, const long long *__restrict__ seg_start,
                    const double *__restrict__ wn, const int *__restrict__ inwin,
                    const int *__restrict__ iown, const double *__restrict__ own, double ownstep,
                    int *__restrict__ head_flag, int *__restrict__ max_len) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const long long l0 = seg_start[s];
    const long long l1 = (s + 1 < nseg) ? seg_start[s + 1] : nlines;
    double anchor = own[iown[l0]];
    head_flag[l0] = 1;
    for (long long l = l0 + 1; l < l1; l++) {
        int flag = 0;
        if (inwin[l]) {
            if (!(fabs(dsub(wn[l], anchor)) < ownstep)) {  // not absorbed: a new head
                flag = 1;
                anchor = own[iown[l]];
            }
        }
        head_flag[l] = flag;
    }
    atomicMax(max_len, (int)min(l1 - l0, (long long)0x7fffffff));
}

__global__ void __launch_bounds__(256)
compact_kernel(long long nlines, const double *__restrict__ wn, const double *__restrict__ elow,
               const double *__restrict__ gf, const unsigned short *__restrict__ iso,
               const int *__restrict__ inwin, const int *__restrict__ cidx,
               const int *__restrict__ iown, const int *__restrict__ head_flag,
               const int *__restrict__ gidx, double *__restrict__ l_wn,
               double *__restrict__ l_elow, double *__restrict__ l_gf, double *__restrict__ g_wn,
               int *__restrict__ g_iown, unsigned int *__restrict__ g_start,
               unsigned short *__restrict__ g_iso) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines || !inwin[l]) return;
    const int c = cidx[l];
    l_wn[c] = wn[l];
    l_elow[c] = elow[l];
    l_gf[c] = gf[l];
    if (head_flag[l]) {
        const int g = gidx[l];
        g_wn[g] = wn[l];
        g_iown[g] = iown[l];
        g_start[g] = (unsigned int)c;
        g_iso[g] = iso[l];
    }
}

// gbin[iso][b] = first group of the isotope's block with iown >= b*binw.
__global__ void __launch_bounds__(256)
gbin_kernel(int niso, int nbins, int binw, const int *__restrict__ gbeg,
            const int *__restrict__ gend, const int *__restrict__ g_iown,
            int *__restrict__ gbin) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int iso = blockIdx.y;
    if (b > nbins || iso >= niso) return;
    int lo = gbeg[iso], hi = gend[iso];
    if (b == nbins) {
        gbin[(size_t)iso * (nbins + 1) + b] = hi;
        return;
    }
    const long long edge = (long long)b * binw;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (g_iown[mid] < edge) lo = mid + 1; else hi = mid;
    }
    gbin[(size_t)iso * (nbins + 1) + b] = lo;
}

template <typename T>
static int dev_alloc(T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    PB_CUDA(cudaMalloc((void **)p, n * sizeof(T)));
    return 0;
}

static int exclusive_scan(cudaStream_t st, const int *in, int *out, long long n, void **tmp,
                          size_t *tmp_bytes) {
    size_t need = 0;
    PB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n, st));
    if (need > *tmp_bytes) {
        if (*tmp) cudaFree(*tmp);
        PB_CUDA(cudaMalloc(tmp, need));
        *tmp_bytes = need;
    }
    PB_CUDA(cub::DeviceScan::ExclusiveSum(*tmp, need, in, out, (int)n, st));
    return 0;
}

int device_group_lines(cudaStream_t st, const GroupInput &in, GroupOutput *out) {
    const long long n = in.nlines;
    out->n_inwin = out->ngroups = 0;
    out->max_segment = 0;
    // everything allocated here is released on every exit path
    double *d_wn = nullptr, *d_elow = nullptr, *d_gf = nullptr, *d_own = nullptr;
    unsigned short *d_iso = nullptr;
    int *d_inwin = nullptr, *d_iown = nullptr, *d_flag = nullptr, *d_cidx = nullptr,
        *d_scan = nullptr, *d_maxlen = nullptr;
    long long *d_segstart = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(d_wn); cudaFree(d_elow); cudaFree(d_gf); cudaFree(d_own); cudaFree(d_iso);
        cudaFree(d_inwin); cudaFree(d_iown); cudaFree(d_flag); cudaFree(d_cidx);
        cudaFree(d_scan); cudaFree(d_maxlen); cudaFree(d_segstart); cudaFree(d_tmp);
    };
#define PB_TRY(expr)                \
    do {                            \
        rc = (expr);                \
        if (rc) { cleanup(); return rc; } \
    } while (0)
#define PB_TRY_CUDA(call)                                                         \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) {                                                  \
            rc = cuda_fail(_e, #call, __FILE__, __LINE__);                        \
            cleanup();                                                            \
            return rc;                                                            \
        }                                                                         \
    } while (0)

    PB_TRY(dev_alloc(&d_wn, (size_t)n));
    PB_TRY(dev_alloc(&d_elow, (size_t)n));
    PB_TRY(dev_alloc(&d_gf, (size_t)n));
    PB_TRY(dev_alloc(&d_iso, (size_t)n));
    PB_TRY(dev_alloc(&d_own, (size_t)in.onwn));
    PB_TRY(dev_alloc(&d_inwin, (size_t)n));
    PB_TRY(dev_alloc(&d_iown, (size_t)n));
    PB_TRY(dev_alloc(&d_flag, (size_t)n));
    PB_TRY(dev_alloc(&d_cidx, (size_t)n + 1));
    PB_TRY(dev_alloc(&d_scan, (size_t)n + 1));
    PB_TRY(dev_alloc(&d_maxlen, 1));
    PB_TRY_CUDA(cudaMemcpyAsync(d_wn, in.wn, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    PB_TRY_CUDA(cudaMemcpyAsync(d_elow, in.elow, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    PB_TRY_CUDA(cudaMemcpyAsync(d_gf, in.gf, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    PB_TRY_CUDA(cudaMemcpyAsync(d_iso, in.iso16, sizeof(unsigned short) * n,
                                cudaMemcpyHostToDevice, st));
    PB_TRY_CUDA(cudaMemcpyAsync(d_own, in.own, sizeof(double) * in.onwn, cudaMemcpyHostToDevice,
                                st));
    PB_TRY_CUDA(cudaMemsetAsync(d_maxlen, 0, sizeof(int), st));

    const unsigned blocks = (unsigned)((n + 255) / 256);
    const double own0 = in.own[0], own_last = in.own[in.onwn - 1];
    const double ownstep = in.own[1] - in.own[0];
    line_flags_kernel<<<blocks, 256, 0, st>>>(n, d_wn, d_iso, d_own, in.onwn, own0, own_last,
                                              ownstep, d_inwin, d_iown, d_flag);
    PB_TRY_CUDA(cudaGetLastError());

    // compacted line index, segment index
    PB_TRY(exclusive_scan(st, d_inwin, d_cidx, n, &d_tmp, &tmp_bytes));
    PB_TRY(exclusive_scan(st, d_flag, d_scan, n, &d_tmp, &tmp_bytes));
    int last_in = 0, last_c = 0, last_cert = 0, last_seg = 0;
    PB_TRY_CUDA(cudaMemcpyAsync(&last_in, d_inwin + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    PB_TRY_CUDA(cudaMemcpyAsync(&last_c, d_cidx + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    PB_TRY_CUDA(cudaMemcpyAsync(&last_cert, d_flag + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    PB_TRY_CUDA(cudaMemcpyAsync(&last_seg, d_scan + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    PB_TRY_CUDA(cudaStreamSynchronize(st));
    const long long n_inwin = (long long)last_c + last_in;
    const long long nseg = (long long)last_seg + last_cert;
    out->n_inwin = n_inwin;

    // block boundaries in compacted-line space (for the per-isotope absorbed-line counts)
    std::vector<int> blk_c(in.nblocks + 1, 0);
    for (int b = 0; b < in.nblocks; b++)
        PB_TRY_CUDA(cudaMemcpyAsync(&blk_c[b], d_cidx + in.block_start[b], sizeof(int),
                                    cudaMemcpyDeviceToHost, st));
    blk_c[in.nblocks] = (int)n_inwin;

    if (nseg > 0) {
        PB_TRY(dev_alloc(&d_segstart, (size_t)nseg));
        segment_starts_kernel<<<blocks, 256, 0, st>>>(n, d_flag, d_scan, d_segstart);
        PB_TRY_CUDA(cudaGetLastError());
        // head flags (reuse d_flag: certain heads are heads; the walk fills the rest)
        PB_TRY_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int) * n, st));
        segment_walk_kernel<<<(unsigned)((nseg + 127) / 128), 128, 0, st>>>(
            nseg, n, d_segstart, d_wn, d_inwin, d_iown, d_own, ownstep, d_flag, d_maxlen);
        PB_TRY_CUDA(cudaGetLastError());
    } else {
        PB_TRY_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int) * n, st));
    }
    PB_TRY(exclusive_scan(st, d_flag, d_scan, n, &d_tmp, &tmp_bytes));
    int last_head = 0, last_g = 0;
    PB_TRY_CUDA(cudaMemcpyAsync(&last_head, d_flag + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    PB_TRY_CUDA(cudaMemcpyAsync(&last_g, d_scan + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    PB_TRY_CUDA(cudaMemcpyAsync(&out->max_segment, d_maxlen, sizeof(int), cudaMemcpyDeviceToHost, st));
    std::vector<int> blk_g(in.nblocks + 1, 0);
    for (int b = 0; b < in.nblocks; b++)
        PB_TRY_CUDA(cudaMemcpyAsync(&blk_g[b], d_scan + in.block_start[b], sizeof(int),
                                    cudaMemcpyDeviceToHost, st));
    PB_TRY_CUDA(cudaStreamSynchronize(st));
    const long long ngroups = (long long)last_g + last_head;
    blk_g[in.nblocks] = (int)ngroups;
    out->ngroups = ngroups;

    // outputs owned by the caller (engine buffers)
    PB_TRY(out->alloc(out->ctx, n_inwin, ngroups));
    compact_kernel<<<blocks, 256, 0, st>>>(n, d_wn, d_elow, d_gf, d_iso, d_inwin, d_cidx, d_iown,
                                           d_flag, d_scan, out->l_wn, out->l_elow, out->l_gf,
                                           out->g_wn, out->g_iown, out->g_start, out->g_iso);
    PB_TRY_CUDA(cudaGetLastError());
    const unsigned int end_mark = (unsigned int)n_inwin;
    PB_TRY_CUDA(cudaMemcpyAsync(out->g_start + ngroups, &end_mark, sizeof(unsigned int),
                                cudaMemcpyHostToDevice, st));

    // per-isotope group ranges and absorbed-line counts
    out->iso_gbeg.assign(in.niso, 0);
    out->iso_gend.assign(in.niso, 0);
    out->iso_nadd.assign(in.niso, 0);
    for (int b = 0; b < in.nblocks; b++) {
        const int iso = in.block_iso[b];
        out->iso_gbeg[iso] = blk_g[b];
        out->iso_gend[iso] = blk_g[b + 1];
        out->iso_nadd[iso] = (long long)(blk_c[b + 1] - blk_c[b]) - (blk_g[b + 1] - blk_g[b]);
    }
    int *d_gbeg = nullptr, *d_gend = nullptr;
    PB_TRY(dev_alloc(&d_gbeg, (size_t)in.niso));
    rc = dev_alloc(&d_gend, (size_t)in.niso);
    if (!rc) {
        cudaMemcpyAsync(d_gbeg, out->iso_gbeg.data(), sizeof(int) * in.niso, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_gend, out->iso_gend.data(), sizeof(int) * in.niso, cudaMemcpyHostToDevice, st);
        dim3 grid((unsigned)((in.nbins + 1 + 255) / 256), (unsigned)in.niso);
        gbin_kernel<<<grid, 256, 0, st>>>(in.niso, in.nbins, in.binw, d_gbeg, d_gend, out->g_iown,
                                          out->gbin);
        cudaError_t e2 = cudaGetLastError();
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);
        if (e2 != cudaSuccess) rc = cuda_fail(e2, "gbin_kernel", __FILE__, __LINE__);
    }
    cudaFree(d_gbeg);
    cudaFree(d_gend);
    cleanup();
    return rc;
#undef PB_TRY
#undef PB_TRY_CUDA
}

}  // namespace pb200

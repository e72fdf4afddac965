// optical_depth.cu -- optical depth from the extinction coefficient (next-tier row, SURVEY.md
// section 8f.3): the step right after the LBL extinction in Pyrat.run (pyrat_obj.py:209).
//
//   plane-parallel  src_c/_trapezoid.c:147-211  cumulative trapezoid down the layers, stops at
//                   maxdepth / ibottom and records the layer (ideep)
//   transit         pyratbay/opacity/optic_depth.py:104-111 + src_c/_trapezoid.c:214-262:
//                   slant depth along the grazing ray of every impact layer, only while the
//                   channel is not yet deeper than maxdepth
// One thread owns one wavenumber channel (channels are independent); a warp reads 32
// consecutive channels of a layer: coalesced, each extinction value read once (plane-parallel)
// -> HBM-bound streaming kernels.  Operation order follows the reference: results bit-exact.
#include <vector>

#include "../../include/pb200_lbl.h"
#include "common.cuh"

namespace pb200 {

__global__ void __launch_bounds__(256)
plane_parallel_depth_kernel(double *__restrict__ depth, int *__restrict__ ideep,
                            const double *__restrict__ ext, const double *__restrict__ dz,
                            double maxdepth, int itop, int ibottom, int nlayers, int nwave) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwave) return;
    double sum = 0.0;
    double prev = (itop >= 0 && itop < nlayers) ? ext[(size_t)itop * nwave + i] : 0.0;
    int k;
    bool stopped = false;
    for (k = 0; k < nlayers; k++) {
        if (k <= itop || stopped) {
            depth[(size_t)k * nwave + i] = 0.0;   // the reference leaves np.zeros here
            continue;
        }
        const double cur = ext[(size_t)k * nwave + i];
        // sum += 0.5*dz[k-1] * (ext[k] + ext[k-1])                      (_trapezoid.c:200-201)
        sum = dadd(sum, dmul(dmul(0.5, dz[k - 1]), dadd(cur, prev)));
        prev = cur;
        depth[(size_t)k * nwave + i] = sum;
        if (sum >= maxdepth || k == ibottom || k == nlayers - 1) {
            ideep[i] = k;
            stopped = true;
        }
    }
    if (!stopped) ideep[i] = nlayers;  // loop ran off the end (only when nlayers-1 <= itop)
}

__global__ void __launch_bounds__(256)
transit_depth_kernel(double *__restrict__ depth, int *__restrict__ ideep,
                     const double *__restrict__ ext, const double *__restrict__ paths,
                     double maxdepth, int itop, int ibottom, int nlayers, int nwave) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nwave) return;
    int deep = -1;
    for (int r = 0; r < nlayers; r++) {
        double tau = 0.0;
        if (r >= itop && r < ibottom && deep < 0) {
            const double *row = paths + (size_t)r * nlayers;
            double prev = ext[(size_t)itop * nwave + j];
            for (int i = 0; i < r - itop; i++) {
                const double cur = ext[(size_t)(itop + i + 1) * nwave + j];
                // tau += path[i] * (ext[i+1] + ext[i])                 (_trapezoid.c:246-249)
                tau = dadd(tau, dmul(row[i], dadd(cur, prev)));
                prev = cur;
            }
            if (tau > maxdepth) deep = r;
        }
        depth[(size_t)r * nwave + j] = tau;
    }
    if (deep < 0) deep = ibottom > itop ? ibottom - 1 : itop;  // optic_depth.py:111
    ideep[j] = deep;
}

static int run_depth(int device, int transit, double *depth, int *ideep, const double *ext,
                     const double *geom, double maxdepth, int itop, int ibottom, int nlayers,
                     int nwave, bool on_device, cudaStream_t user) {
    if (!depth || !ideep || !ext || !geom || nlayers < 1 || nwave < 1 || itop < 0) {
        set_error("pb200_optical_depth: null argument or empty grid");
        return PB200_EINVAL;
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device: this engine has no CPU fallback");
        return PB200_ENODEVICE;
    }
    PB_CUDA(cudaSetDevice(device));
    cudaStream_t st = user;
    bool own = false;
    if (!st) {
        PB_CUDA(cudaStreamCreate(&st));
        own = true;
    }
    const size_t cells = (size_t)nlayers * nwave;
    const size_t ngeom = transit ? (size_t)nlayers * nlayers : (size_t)(nlayers > 1 ? nlayers - 1 : 1);
    double *d_depth = depth, *d_geom = nullptr;
    const double *d_ext = ext;
    int *d_ideep = ideep;
    double *tmp_ext = nullptr;
    int rc = 0;
    auto fail_cuda = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == 0) rc = cuda_fail(e, what, __FILE__, __LINE__);
    };
    fail_cuda(cudaMalloc((void **)&d_geom, sizeof(double) * ngeom), "malloc geom");
    if (!rc) fail_cuda(cudaMemcpyAsync(d_geom, geom, sizeof(double) * ngeom,
                                       cudaMemcpyHostToDevice, st), "h2d geom");
    if (!on_device && !rc) {
        d_depth = nullptr;
        d_ideep = nullptr;
        fail_cuda(cudaMalloc((void **)&d_depth, sizeof(double) * cells), "malloc depth");
        if (!rc) fail_cuda(cudaMalloc((void **)&tmp_ext, sizeof(double) * cells), "malloc ext");
        if (!rc) fail_cuda(cudaMalloc((void **)&d_ideep, sizeof(int) * nwave), "malloc ideep");
        if (!rc) fail_cuda(cudaMemcpyAsync(tmp_ext, ext, sizeof(double) * cells,
                                           cudaMemcpyHostToDevice, st), "h2d ext");
        d_ext = tmp_ext;
    }
    if (!rc) {
        const int blocks = (nwave + 255) / 256;
        if (transit)
            transit_depth_kernel<<<blocks, 256, 0, st>>>(d_depth, d_ideep, d_ext, d_geom, maxdepth,
                                                         itop, ibottom, nlayers, nwave);
        else
            plane_parallel_depth_kernel<<<blocks, 256, 0, st>>>(d_depth, d_ideep, d_ext, d_geom,
                                                                maxdepth, itop, ibottom, nlayers,
                                                                nwave);
        fail_cuda(cudaGetLastError(), "depth kernel");
    }
    if (!on_device && !rc) {
        fail_cuda(cudaMemcpyAsync(depth, d_depth, sizeof(double) * cells, cudaMemcpyDeviceToHost,
                                  st), "d2h depth");
        fail_cuda(cudaMemcpyAsync(ideep, d_ideep, sizeof(int) * nwave, cudaMemcpyDeviceToHost, st),
                  "d2h ideep");
    }
    fail_cuda(cudaStreamSynchronize(st), "sync");
    if (d_geom) cudaFree(d_geom);
    if (!on_device) {
        if (d_depth) cudaFree(d_depth);
        if (tmp_ext) cudaFree(tmp_ext);
        if (d_ideep) cudaFree(d_ideep);
    }
    if (own) cudaStreamDestroy(st);
    return rc;
}

}  // namespace pb200

extern "C" {

int pb200_optical_depth(int device, int transit, double *depth, int32_t *ideep,
                        const double *extinction, const double *geometry, double maxdepth,
                        int itop, int ibottom, int nlayers, int nwave) {
    return pb200::run_depth(device, transit, depth, ideep, extinction, geometry, maxdepth, itop,
                            ibottom, nlayers, nwave, false, nullptr);
}

int pb200_optical_depth_dev(int device, int transit, double *depth_dev, int32_t *ideep_dev,
                            const double *extinction_dev, const double *geometry,
                            double maxdepth, int itop, int ibottom, int nlayers, int nwave,
                            void *cuda_stream) {
    return pb200::run_depth(device, transit, depth_dev, ideep_dev, extinction_dev, geometry,
                            maxdepth, itop, ibottom, nlayers, nwave, true,
                            (cudaStream_t)cuda_stream);
}

}  // extern "C"

// dense_kernels.cu -- accumulate as a dense strided convolution (constant-step output grids).
//
// The gather kernels (lbl_kernels.cu) deliver one 8-byte profile sample from L1/L2 per FMA and
// are bound by that delivery (L2->SM at 97 % in table mode, 11 % of the fp64 peak).  For an
// isotope whose co-add groups occupy a large share of the fine grid the same sums are a dense
// 1-D convolution with operand reuse in registers:
//
//   a group sits on ONE fine cell f = S*c + r (c: output cell, r: sub-cell offset, S = fine
//   samples per output sample), no two groups of an isotope share a cell (_extcoeff.c:249-262),
//   and it adds  k * P[half - r + S*(m - c)]  to output m for m - c in a window [DL(r), DH(r))
//   that depends on r only (for one unit, isotope and Doppler sample).  With the strengths
//   scattered to a dense array K[f] (zeros where there is no group):
//
//       out[m] = sum_r sum_c K[S*c + r] * W_r[m - c],   W_r[d] = P[half - r + S*d] inside the
//                                                       window, 0 outside.
//
// A lane owns one r, a warp 32 consecutive r and kDenseJ consecutive outputs in registers; it
// walks the cells c that reach its outputs, loading ONE K value and ONE new W value per cell
// for kDenseJ FMAs (the W window slides through a statically rotated register file).  K and W
// tiles are staged in shared memory per (32 r) block and shared by the CTA's 16 warps, the lanes
// are summed with shuffles at the very end, one CTA owns its outputs: no atomics, fixed order.
//
// Discrete decisions stay those of the reference.  The window of a cell comes from the SAME
// device functions the gather kernels use (group_prep.cuh: dynamic_range, output_range),
// evaluated at a reference cell in the middle of the grid (the window is translation invariant
// away from the grid ends; clipping at the ends equals "outputs that exist").  Two things depend
// on the line's exact wavenumber and not only on its cell:
//   * the Doppler sample (:278): groups ascend in wavenumber, so the cells of one Doppler
//     sample are a contiguous segment; segment_bounds_kernel finds the boundaries with the
//     reference's nearest search and the tiles of K are masked per segment;
//   * the dynamic index idwn = trunc((w - own0)/dwnstep) (:275), which is (cell / ofactor) - 1
//     instead of cell / ofactor for a line below the centre of a cell divisible by ofactor
//     ("anomalous" cells; static per ofactor, anomaly_bits_kernel).  Their window is shifted by
//     one dynamic sample; the at most two output samples by which it differs are added /
//     removed explicitly per anomalous cell.
// The host uses this path only when the windows are provably translation invariant
// (engine.cu: frac(cutoff/dwnstep) not within 1e-6 of an integer, grid ends far away).
#include "dense_kernels.cuh"
#include "group_prep.cuh"

#include <climits>
#include <cstdlib>

namespace pb200 {

namespace {
constexpr int J = kDenseJ;
constexpr int kKRows = kDenseTile + kDenseSpanMax;           // K tile rows (cells)
constexpr int kWRows = kDenseSpanMax + 2 * J;                // W tile rows (output offsets)
}  // namespace

// 8-byte asynchronous global->shared copy; `on == false` writes zeros without reading.
__device__ __forceinline__ void async_copy8(double *dst, const double *src, bool on) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = on ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(saddr), "l"(src), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void async_copy8s(unsigned saddr, const double *src, bool on) {
    const int bytes = on ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(saddr), "l"(src), "r"(bytes)
                 : "memory");
}

// 16-byte form (source and destination 16-byte aligned).
__device__ __forceinline__ void async_copy16s(unsigned saddr, const double *src, bool on) {
    const int bytes = on ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(src), "r"(bytes)
                 : "memory");
}

// min(rows, max(0, ceil(x / S))): the first row t of a tile with S*t >= x, without a 64-bit
// division (fd divides non-negative 31-bit numbers by S exactly).
__device__ __forceinline__ int first_row(long long x, int rows, int S, const FastDiv &fd) {
    if (x <= 0) return 0;
    if (x >= (long long)rows * S) return rows;
    return fd.div_ceil((int)x);
}

__device__ __forceinline__ void async_copy_wait() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

size_t dense_smem_bytes() {
    return sizeof(double) * kDenseLanes * (kKRows + kWRows) +
           sizeof(unsigned short) * kKRows * kMaxMerge + sizeof(short4) * kDenseMaxStride;
}

__global__ void __launch_bounds__(256)
anomaly_bits_kernel(StaticView V, long long gbeg, long long gend, UnitParams U,
                    unsigned *__restrict__ bits, int *__restrict__ err) {
    const long long g = gbeg + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= gend) return;
    const int iown = V.g_iown[g];
    const int idwn = dynamic_index(V, U, V.g_wn[g]);
    const int idwn0 = U.fd_ofactor.div(iown);
    if (idwn == idwn0) return;
    if (idwn == idwn0 - 1 && iown - idwn0 * U.ofactor == 0)
        atomicOr(&bits[iown >> 5], 1u << (iown & 31));
    else
        atomicExch(err, 2);
}

__global__ void __launch_bounds__(256)
densify_kernel(StaticView V, long long gbeg, long long gend, const double *__restrict__ ks,
               const unsigned long long *__restrict__ kmax_entry, double ethresh,
               double *__restrict__ kd) {
    const long long g = gbeg + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= gend) return;
    const double kthr = dmul(ethresh, __longlong_as_double((long long)*kmax_entry));
    const double k = ks[g];
    if (k < kthr) return;  // :265
    kd[V.g_iown[g]] = k;
}

__global__ void __launch_bounds__(256)
merge_minor_kernel(StaticView V, long long gbeg, long long gend, const double *__restrict__ ks,
                   const unsigned long long *__restrict__ kmax_entry, double ethresh, double adop,
                   const int *__restrict__ main_bounds, double *__restrict__ kd,
                   double *__restrict__ kd_all) {
    extern __shared__ double s_dop[];
    for (int i = threadIdx.x; i < V.ndop; i += blockDim.x)
        s_dop[i] = V.dop_thr ? V.dop_thr[i] : V.doppler[i];
    __syncthreads();
    const long long g = gbeg + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= gend) return;
    const double kthr = dmul(ethresh, __longlong_as_double((long long)*kmax_entry));
    const double k = ks[g];
    if (k < kthr) return;  // :265
    const int idop = V.ndop >= 2 ? doppler_index(V, s_dop, dmul(adop, V.g_wn[g])) : 0;  // :278
    const int cell = V.g_iown[g];
    if (main_bounds[idop] <= cell && cell < main_bounds[idop + 1]) {
        kd[cell] = k;
        kd_all[cell] = dadd(kd_all[cell], k);   // one group per cell and isotope: no race
    }
}

__global__ void __launch_bounds__(256)
segment_bounds_kernel(StaticView V, long long gbeg, long long gend, double adop,
                      int *__restrict__ bounds) {
    extern __shared__ double s_dop[];
    for (int i = threadIdx.x; i < V.ndop; i += blockDim.x)
        s_dop[i] = V.dop_thr ? V.dop_thr[i] : V.doppler[i];
    __syncthreads();
    const int last = (int)(V.onwn < 0x7fffffffLL ? V.onwn : 0x7fffffffLL);
    for (int j = threadIdx.x; j <= V.ndop; j += blockDim.x) {
        int b = 0;
        if (j == V.ndop) {
            b = last;
        } else if (j > 0) {
            // first group whose nearest Doppler sample (:278) is >= j (monotonic in wavenumber)
            long long lo = gbeg, hi = gend;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                const int idop =
                    V.ndop >= 2 ? doppler_index(V, s_dop, dmul(adop, V.g_wn[mid])) : 0;
                if (idop >= j) hi = mid; else lo = mid + 1;
            }
            b = lo < gend ? V.g_iown[lo] : last;
        }
        bounds[j] = b;
    }
}

// One sub-step of the convolution: cell `krow + I` of the warp's walk.  The window registers
// rotate statically: output x reads w[(x - I) mod J], the register freed by output J-1 takes
// the new offset.
template <int I>
__device__ __forceinline__ void dense_step(const double *__restrict__ kp,
                                           const double *__restrict__ wp, double (&w)[kDenseJ],
                                           double (&acc)[kDenseJ]) {
    constexpr int J = kDenseJ;
    const double kv = kp[I * kDenseLanes];
    w[(J - I) % J] = wp[-I * kDenseLanes];
#pragma unroll
    for (int x = 0; x < J; x++) acc[x] = fma(kv, w[(x - I + J) % J], acc[x]);
}

// The first J cells of a walk: at its n-th cell only the outputs x <= n are inside the widest
// window (offset (dmax-1) + x - n < dmax), the others would multiply zero padding.
template <int I>
__device__ __forceinline__ void dense_steps_first(const double *__restrict__ kp,
                                                  const double *__restrict__ wp,
                                                  double (&w)[kDenseJ], double (&acc)[kDenseJ]) {
    constexpr int J = kDenseJ;
    if constexpr (I < J) {
        const double kv = kp[I * kDenseLanes];
        w[(J - I) % J] = wp[-I * kDenseLanes];
#pragma unroll
        for (int x = 0; x <= I; x++) acc[x] = fma(kv, w[(x - I + J) % J], acc[x]);
        dense_steps_first<I + 1>(kp, wp, w, acc);
    }
}

// The last J-1 cells of a walk, entered at rotation phase P (= the number of steps the partial
// iteration before them took): at the k-th of them only the outputs x >= k are still inside the
// widest window.  kp / wp point at the phase-0 rows of the current iteration.
template <int P, int K>
__device__ __forceinline__ void dense_steps_last(const double *__restrict__ kp,
                                                 const double *__restrict__ wp,
                                                 double (&w)[kDenseJ], double (&acc)[kDenseJ]) {
    constexpr int J = kDenseJ;
    if constexpr (K < J) {
        constexpr int N = P + K - 1;        // steps since the iteration's phase 0
        constexpr int I = N % J;            // rotation phase of this step
        const double kv = kp[N * kDenseLanes];
        w[(J - I) % J] = wp[-N * kDenseLanes];
#pragma unroll
        for (int x = K; x < J; x++) acc[x] = fma(kv, w[(x - I + J) % J], acc[x]);
        dense_steps_last<P, K + 1>(kp, wp, w, acc);
    }
}

template <int P>
__device__ __forceinline__ void dense_last_dispatch(int phase, const double *__restrict__ kp,
                                                    const double *__restrict__ wp,
                                                    double (&w)[kDenseJ], double (&acc)[kDenseJ]) {
    if constexpr (P < kDenseJ) {
        if (phase == P)   // warp-uniform
            dense_steps_last<P, 1>(kp, wp, w, acc);
        else
            dense_last_dispatch<P + 1>(phase, kp, wp, w, acc);
    }
}

template <int I>
__device__ __forceinline__ void dense_steps(const double *__restrict__ kp,
                                            const double *__restrict__ wp, double (&w)[kDenseJ],
                                            double (&acc)[kDenseJ], int count) {
    if constexpr (I < kDenseJ) {
        if (I < count) {   // warp-uniform
            dense_step<I>(kp, wp, w, acc);
            dense_steps<I + 1>(kp, wp, w, acc, count);
        }
    }
}

// The walk of one half-warp over the cells that reach its J outputs: step n is cell
// c_start + n (tile row orow + n), at which output x needs offset (dmax-1) + x - n.
// [it_lo, it_hi) are the iterations (of J cells each) that can hold a non-zero K, warp-uniform:
// a tile that straddles a Doppler-segment boundary walks only its side of it in each pass.
__device__ __forceinline__ void dense_walk(const double (*Ks)[kDenseLanes],
                                           const double (*Ws)[kDenseLanes], int orow, int l16,
                                           int dmax, int w_lo, int nsteps, int it_lo, int it_hi,
                                           double (&acc)[kDenseJ]) {
    constexpr int J = kDenseJ, L = kDenseLanes;
    const int niter = (nsteps + J - 1) / J;
    if (it_hi <= it_lo) return;
    double w[J];
    if (it_lo == 0 && it_hi >= niter) {
        // the whole walk.  cells 0..J-1: growing triangle; cells J..span-1: all outputs; the
        // last J-1 cells: shrinking triangle (span = nsteps - (J-1) >= J, else the plain walk)
#pragma unroll
        for (int x = 0; x < J; x++) w[x] = 0.0;   // offsets >= dmax: outside all windows
        const double *kp = &Ks[orow][l16];
        const double *wp = &Ws[(dmax - 1) - w_lo][l16];
        const int span = nsteps - (J - 1);
        dense_steps_first<0>(kp, wp, w, acc);     // nsteps >= J always
        kp += J * L;
        wp -= J * L;
        int n = J;
        const int full_end = span >= J ? span : nsteps;
        for (; n + J <= full_end; n += J) {
            dense_steps<0>(kp, wp, w, acc, J);
            kp += J * L;
            wp -= J * L;
        }
        dense_steps<0>(kp, wp, w, acc, full_end - n);
        if (span >= J) dense_last_dispatch<0>(full_end - n, kp, wp, w, acc);
        return;
    }
    // part of the walk: the window registers as the iteration before it_lo would have left them
    const int n0 = it_lo * J;
    w[0] = 0.0;
#pragma unroll
    for (int x = 1; x < J; x++) w[x] = Ws[(dmax - 1 + x - n0) - w_lo][l16];
    const double *kp = &Ks[orow + n0][l16];
    const double *wp = &Ws[(dmax - 1) - w_lo - n0][l16];
    for (int it = it_lo; it < it_hi; it++) {
        dense_steps<0>(kp, wp, w, acc, min(J, nsteps - it * J));
        kp += J * L;
        wp -= J * L;
    }
}

// Thread layout: a warp owns 2 x J consecutive outputs and 16 sub-cell offsets: lanes 0-15 hold
// offsets r0..r0+15 for the first J outputs, lanes 16-31 the same offsets for the next J.  Tile
// rows are 16 doubles (128 B); the two half-warps read two different rows per load, which is
// the same two shared-memory wavefronts as one 256-byte row.  Half-width rows keep the tiles
// under 100 KB, so two CTAs share an SM: one stages while the other computes.
__global__ void __launch_bounds__(kDenseWarps * 32, 2)
accumulate_dense_kernel(StaticView V, const UnitParams *__restrict__ units,
                        const IsoUnit *__restrict__ iso_units, DenseSet D, int row, int nrows,
                        long long abits_words, double cutoff, double *__restrict__ out,
                        int *__restrict__ err) {
    constexpr int L = kDenseLanes;
    extern __shared__ double s_dyn[];
    double (*Ks)[L] = reinterpret_cast<double (*)[L]>(s_dyn);     // [kKRows]: cell c_lo + t
    double (*Ws)[L] = Ks + kKRows;                                // [kWRows]: offset w_lo + t
    // anomaly bits of the K rows, one plane of kKRows entries per isotope of the set
    unsigned short *As = reinterpret_cast<unsigned short *>(Ws + kWRows);
    short4 *s_win = reinterpret_cast<short4 *>(As + kMaxMerge * kKRows);   // [S] windows
    __shared__ int s_dmin, s_dmax;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 4, l16 = lane & (L - 1);
    // blockIdx.x = unit (fastest): the CTAs in flight share a tile of K in L2
    const UnitParams U = units[blockIdx.x];
    const int iso = D.iso[0];
    const IsoUnit I = iso_units[(size_t)blockIdx.x * V.niso + iso];
    // a merged unit convolves the sum plane and corrects every isotope's anomalous cells
    const bool merged_unit = (U.aslot & kMergedUnitBit) != 0;
    const int nset = merged_unit ? D.n : 1;
    const double *__restrict__ kd = merged_unit ? D.kd_all : D.kd[0];
    const int *__restrict__ bounds = D.bounds;
    const long long slot_off = (long long)(U.aslot & ~kMergedUnitBit) * abits_words;
    const int S = V.tstride;
    const int ms = blockIdx.y * kDenseTile;
    const int m_end = min(ms + kDenseTile, min(V.nwave, U.mcount));   // exclusive
    if (ms >= m_end) return;
    const int orow = (warp * 2 + grp) * J;       // first output of this half-warp within the tile

    double acc[J];
#pragma unroll
    for (int x = 0; x < J; x++) acc[x] = 0.0;

    // fine cells that can reach this tile (conservative: the unit-wide reach of the isotope)
    const long long reach_cells = I.reach / S + 2;
    const long long flo = max(0LL, ((long long)ms - reach_cells) * S);
    const long long fhi = min(V.onwn, ((long long)m_end + reach_cells) * S);
    const int c_ref = U.mcount >> 1;   // reference cell for the windows, far from both grid ends

    for (int seg = 0; seg < V.ndop; seg++) {
        // cells of Doppler sample `seg` that this path owns (below dense_from: gather kernels)
        const long long sa = max((long long)bounds[seg], (long long)I.dense_from);
        const long long sb = bounds[seg + 1];
        if (sb <= sa || sb <= flo || sa >= fhi) continue;         // CTA-uniform
        const ProfileSlot ps = load_slot(V.pslot + I.ilor * V.ndop + seg);
        const int half = ps.half;
        const double *__restrict__ prof = V.profile + ps.base;

        // (1) window of every sub-cell offset: outputs m - c in [x, y) for a regular cell,
        //     [z, w) for an anomalous one (only cells divisible by ofactor can be)
        __syncthreads();
        if (threadIdx.x == 0) {
            s_dmin = INT_MAX;
            s_dmax = INT_MIN;
        }
        __syncthreads();
        for (int r = threadIdx.x; r < S; r += blockDim.x) {
            const int iown = c_ref * S + r;
            const int idwn0 = U.fd_ofactor.div(iown);
            int jlo, jhi, mlo, mhi, alo, ahi;
            dynamic_range(U, cutoff, half, iown, idwn0, &jlo, &jhi);
            output_range(U, jlo, jhi, &mlo, &mhi);
            alo = mlo;
            ahi = mhi;
            if (iown - idwn0 * U.ofactor == 0) {
                dynamic_range(U, cutoff, half, iown, idwn0 - 1, &jlo, &jhi);
                output_range(U, jlo, jhi, &alo, &ahi);
            }
            if (mhi <= mlo) mlo = mhi = c_ref;   // empty windows: keep the offsets small
            if (ahi <= alo) alo = ahi = c_ref;
            s_win[r] = make_short4((short)(mlo - c_ref), (short)(mhi - c_ref),
                                   (short)(alo - c_ref), (short)(ahi - c_ref));
            if (mhi > mlo || ahi > alo) {
                const int lo = (mhi > mlo ? (ahi > alo ? min(mlo, alo) : mlo) : alo) - c_ref;
                const int hi = (mhi > mlo ? (ahi > alo ? max(mhi, ahi) : mhi) : ahi) - c_ref;
                atomicMin(&s_dmin, lo);
                atomicMax(&s_dmax, hi);
            }
        }
        __syncthreads();
        const int dmin = s_dmin, dmax = s_dmax;   // a cell c reaches outputs c+dmin .. c+dmax-1
        if (dmax <= dmin) continue;
        if (dmax - dmin > kDenseSpanMax || dmin < -30000 || dmax > 30000) {
            if (threadIdx.x == 0) atomicExch(err, 1);   // the host sizes batches so this cannot happen
            continue;
        }

        const int c_lo = ms - (dmax - 1);             // first cell that reaches the tile
        const int w_lo = dmin - J;                    // first offset held in Ws
        const int nsteps = J + (dmax - dmin) - 1;     // cells that reach one half-warp's outputs
        const int krows = kDenseTile - J + nsteps;
        const int wrows = dmax + J - w_lo;
        const int nrb = (S + L - 1) / L;
        // iterations of this half-warp's walk whose cells can lie in the segment (warp-uniform)
        int it_lo, it_hi;
        {
            const int seg_c_lo = V.fd_tstride.div((int)sa), seg_c_hi = V.fd_tstride.div((int)(sb - 1));
            const int c_start = c_lo + orow;
            const int n_lo = max(0, seg_c_lo - c_start), n_hi = min(nsteps, seg_c_hi + 1 - c_start);
            const bool any = n_hi > n_lo;
            it_lo = __reduce_min_sync(0xffffffffu, any ? n_lo / J : 0x7fffffff);
            it_hi = __reduce_max_sync(0xffffffffu, any ? (n_hi + J - 1) / J : 0);
        }

        for (int rb = 0; rb < nrb; rb++) {
            const int r = rb * L + l16;
            const bool active = r < S;
            const short4 win = active ? s_win[r] : make_short4(0, 0, 0, 0);
            __syncthreads();   // the previous block's tiles are fully consumed

            // (2)+(3) stage the tiles with asynchronous copies (LDGSTS): all rows of a warp are
            // in flight at once, a masked element is zero-filled (source size 0).  The masks are
            // row ranges computed once per block, the addresses advance by a constant.
            //   W: the unit's profile on the sub-cell offsets of this block, zero outside each
            //      offset's window (reference layout: consecutive r are consecutive samples);
            {
                const int t0 = warp * 2 + grp;
                const int wa = win.x - w_lo, wn_rows = active ? win.y - win.x : 0;   // rows on
                const double *src = prof + ((long long)half - r + (long long)S * (w_lo + t0));
                const long long step = (long long)S * (kDenseWarps * 2);
                unsigned dst = (unsigned)__cvta_generic_to_shared(&Ws[t0][l16]);
                for (int t = t0; t < wrows; t += kDenseWarps * 2) {
                    const bool on = (unsigned)(t - wa) < (unsigned)wn_rows;
                    async_copy8s(dst, on ? src : prof, on);
                    src += step;
                    dst += kDenseWarps * 2 * L * 8;
                }
            }
            //   K: the dense strengths, masked to the Doppler segment.  Two sub-cell offsets per
            //      thread (16-byte copies) when the rows are 16-byte aligned (S even).
            if ((S & 1) == 0) {
                const int pr = lane & 7;                       // pair of offsets 2*pr, 2*pr + 1
                const int t0 = warp * 4 + (lane >> 3);         // 4 rows per warp instruction
                const int r2 = rb * L + 2 * pr;
                // rows t with sa <= (c_lo + t)*S + r < sb and c_lo + t >= 0, for r2 and r2 + 1
                const long long base = (long long)c_lo * S + r2;
                int lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
                if (r2 < S) {
                    const long long a = max(sa, 0LL);
                    lo0 = first_row(a - base, krows, S, V.fd_tstride);
                    hi0 = first_row(sb - base, krows, S, V.fd_tstride);
                    lo1 = first_row(a - base - 1, krows, S, V.fd_tstride);
                    hi1 = first_row(sb - base - 1, krows, S, V.fd_tstride);
                    if (c_lo < 0) {
                        lo0 = max(lo0, -c_lo);
                        lo1 = max(lo1, -c_lo);
                    }
                }
                const double *src = kd + (base + (long long)S * t0);
                const long long step = (long long)S * (kDenseWarps * 4);
                unsigned dst = (unsigned)__cvta_generic_to_shared(&Ks[t0][2 * pr]);
                for (int t = t0; t < krows; t += kDenseWarps * 4) {
                    const bool on0 = t >= lo0 && t < hi0, on1 = t >= lo1 && t < hi1;
                    if (on0 == on1) {
                        async_copy16s(dst, on0 ? src : kd, on0);
                    } else {
                        async_copy8s(dst, on0 ? src : kd, on0);
                        async_copy8s(dst + 8, on1 ? src + 1 : kd, on1);
                    }
                    src += step;
                    dst += kDenseWarps * 4 * L * 8;
                }
            } else {
                for (int t = warp * 2 + grp; t < krows; t += kDenseWarps * 2) {
                    const int c = c_lo + t;
                    const long long cell = (long long)c * S + r;
                    const bool on = active && c >= 0 && cell >= sa && cell < sb;
                    async_copy8(&Ks[t][l16], on ? kd + cell : kd, on);
                }
            }
            // anomaly bits of the K rows (only blocks with an offset whose anomalous window
            // differs need them): thread i of the CTA takes rows i, i + 256, ...
            const bool fix = __syncthreads_or(active && (win.z != win.x || win.w != win.y));
            if (fix) {
                for (int t = threadIdx.x; t < krows; t += blockDim.x) {
                    const int c = c_lo + t;
                    const long long cell0 = (long long)c * S + rb * L;
                    const bool in = c >= 0 && cell0 < V.onwn;
                    const long long wi = cell0 >> 5;
                    for (int q = 0; q < nset; q++) {
                        unsigned bits = 0u;
                        if (in) {
                            const unsigned *ab = D.abits[q] + slot_off;
                            bits = __funnelshift_r(ab[wi], ab[wi + 1], (unsigned)(cell0 & 31));
                        }
                        As[q * kKRows + t] = (unsigned short)(bits & 0xffffu);
                    }
                }
            }
            async_copy_wait();
            __syncthreads();

            // (4) the convolution.  At its n-th cell, output x of the half-warp needs offset
            //     (dmax-1) + x - n.
            dense_walk(Ks, Ws, orow, l16, dmax, w_lo, nsteps, it_lo, it_hi, acc);

            // (5) anomalous cells of this block: add the samples their window has and the
            //     regular one lacks, remove the opposite (at most one output at either end).
            if (fix && active && (win.z != win.x || win.w != win.y)) {
#pragma unroll 1
                for (int part = 0; part < 2; part++) {
                    const int a = part ? min(win.y, win.w) : min(win.x, win.z);
                    const int b = part ? max(win.y, win.w) : max(win.x, win.z);
#pragma unroll 1
                    for (int d = a; d < b; d++) {
                        const bool in_n = d >= win.x && d < win.y;
                        const bool in_a = d >= win.z && d < win.w;
                        if (in_n == in_a) continue;
                        const long long pi = (long long)half - r + (long long)S * d;
                        if (pi < 0 || pi > 2LL * half) continue;
                        double pv = prof[pi];
                        if (in_n) pv = -pv;
                        const int t0 = orow + (dmax - 1) - d;   // K row of output 0's cell
                        // the strength of the anomalous GROUP (the tile holds the sum over the
                        // merged isotopes): from the isotope's own array, masked like the tile
                        for (int q = 0; q < nset; q++) {
                            const unsigned short *aq = As + q * kKRows + t0;
                            const double *__restrict__ kq = D.kd[q];
#pragma unroll
                            for (int x = 0; x < J; x++)
                                if ((aq[x] >> l16) & 1u) {
                                    const long long cell = (long long)(c_lo + t0 + x) * S + r;
                                    if (cell >= sa && cell < sb) acc[x] = fma(kq[cell], pv, acc[x]);
                                }
                        }
                    }
                }
            }
        }
    }

    // (6) sum the 16 sub-cell offsets of a half-warp and add to the output row (the gather
    //     kernels have written it: other isotopes and narrow footprints, or zeros)
    double mine = 0.0;
#pragma unroll
    for (int x = 0; x < J; x++) {
        double v = acc[x];
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (l16 == x) mine = v;
    }
    mine = dmul(mine, I.dens);   // :271-272 (1 unless add): applied to the sum, not per line
    if (ms + orow + l16 < m_end) {
        double *dst = out + ((size_t)U.out_index * nrows + row) * (size_t)V.nwave;
        dst[ms + orow + l16] += mine;
    }
}

// Warp-specialised form of the same kernel: ONE CTA per SM with two tile buffers; two producer
// warps stage block rb+1 (LDGSTS) while the eight consumer warps convolve block rb, so the
// FMA pipe does not wait for tile staging.  Hand-off through named barriers: full[b] (the
// producers arrive, the consumers wait), empty[b] (the other way round).  Everything else --
// windows, masks, walk, corrections, summation order -- is the code of accumulate_dense_kernel.
constexpr int kWsConsumers = kDenseWarps * 32;          // 256 threads
constexpr int kWsProducers = 64;                        // two warps
constexpr int kWsThreads = kWsConsumers + kWsProducers;

__device__ __forceinline__ void named_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

size_t dense_ws_smem_bytes() {
    return 2 * (sizeof(double) * kDenseLanes * (kKRows + kWRows) +
                sizeof(unsigned short) * kKRows * kMaxMerge) +
           sizeof(short4) * kDenseMaxStride;
}

__global__ void __launch_bounds__(kWsThreads, 1)
accumulate_dense_ws_kernel(StaticView V, const UnitParams *__restrict__ units,
                           const IsoUnit *__restrict__ iso_units, DenseSet D, int row, int nrows,
                           long long abits_words, double cutoff, double *__restrict__ out,
                           int *__restrict__ err) {
    constexpr int L = kDenseLanes;
    extern __shared__ double s_dyn[];
    // two buffers of {K tile, W tile, anomaly bits}, then the window table
    constexpr size_t kBufDoubles = (size_t)L * (kKRows + kWRows) + (kKRows * kMaxMerge) / 4;
    short4 *s_win = reinterpret_cast<short4 *>(s_dyn + 2 * kBufDoubles);
    __shared__ int s_dmin, s_dmax;
    __shared__ int s_fix[2];

    const bool producer = threadIdx.x >= kWsConsumers;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 4, l16 = lane & (L - 1);
    const UnitParams U = units[blockIdx.x];
    const int iso = D.iso[0];
    const IsoUnit I = iso_units[(size_t)blockIdx.x * V.niso + iso];
    const bool merged_unit = (U.aslot & kMergedUnitBit) != 0;
    const int nset = merged_unit ? D.n : 1;
    const double *__restrict__ kd = merged_unit ? D.kd_all : D.kd[0];
    const int *__restrict__ bounds = D.bounds;
    const long long slot_off = (long long)(U.aslot & ~kMergedUnitBit) * abits_words;
    const int S = V.tstride;
    const int ms = blockIdx.y * kDenseTile;
    const int m_end = min(ms + kDenseTile, min(V.nwave, U.mcount));   // exclusive
    if (ms >= m_end) return;
    const int orow = (warp * 2 + grp) * J;       // consumers: first output of this half-warp

    double acc[J];
#pragma unroll
    for (int x = 0; x < J; x++) acc[x] = 0.0;

    const long long reach_cells = I.reach / S + 2;
    const long long flo = max(0LL, ((long long)ms - reach_cells) * S);
    const long long fhi = min(V.onwn, ((long long)m_end + reach_cells) * S);
    const int c_ref = U.mcount >> 1;

    for (int seg = 0; seg < V.ndop; seg++) {
        const long long sa = max((long long)bounds[seg], (long long)I.dense_from);
        const long long sb = bounds[seg + 1];
        if (sb <= sa || sb <= flo || sa >= fhi) continue;         // CTA-uniform
        const ProfileSlot ps = load_slot(V.pslot + I.ilor * V.ndop + seg);
        const int half = ps.half;
        const double *__restrict__ prof = V.profile + ps.base;

        // (1) windows of every sub-cell offset (all threads)
        __syncthreads();
        if (threadIdx.x == 0) {
            s_dmin = INT_MAX;
            s_dmax = INT_MIN;
        }
        __syncthreads();
        for (int r = threadIdx.x; r < S; r += blockDim.x) {
            const int iown = c_ref * S + r;
            const int idwn0 = U.fd_ofactor.div(iown);
            int jlo, jhi, mlo, mhi, alo, ahi;
            dynamic_range(U, cutoff, half, iown, idwn0, &jlo, &jhi);
            output_range(U, jlo, jhi, &mlo, &mhi);
            alo = mlo;
            ahi = mhi;
            if (iown - idwn0 * U.ofactor == 0) {
                dynamic_range(U, cutoff, half, iown, idwn0 - 1, &jlo, &jhi);
                output_range(U, jlo, jhi, &alo, &ahi);
            }
            if (mhi <= mlo) mlo = mhi = c_ref;
            if (ahi <= alo) alo = ahi = c_ref;
            s_win[r] = make_short4((short)(mlo - c_ref), (short)(mhi - c_ref),
                                   (short)(alo - c_ref), (short)(ahi - c_ref));
            if (mhi > mlo || ahi > alo) {
                const int lo = (mhi > mlo ? (ahi > alo ? min(mlo, alo) : mlo) : alo) - c_ref;
                const int hi = (mhi > mlo ? (ahi > alo ? max(mhi, ahi) : mhi) : ahi) - c_ref;
                atomicMin(&s_dmin, lo);
                atomicMax(&s_dmax, hi);
            }
        }
        __syncthreads();
        const int dmin = s_dmin, dmax = s_dmax;
        if (dmax <= dmin) continue;
        if (dmax - dmin > kDenseSpanMax || dmin < -30000 || dmax > 30000) {
            if (threadIdx.x == 0) atomicExch(err, 1);
            continue;
        }

        const int c_lo = ms - (dmax - 1);
        const int w_lo = dmin - J;
        const int nsteps = J + (dmax - dmin) - 1;
        const int krows = kDenseTile - J + nsteps;
        const int wrows = dmax + J - w_lo;
        const int nrb = (S + L - 1) / L;
        // iterations of this half-warp's walk whose cells can lie in the segment (warp-uniform)
        int it_lo, it_hi;
        {
            const int seg_c_lo = V.fd_tstride.div((int)sa), seg_c_hi = V.fd_tstride.div((int)(sb - 1));
            const int c_start = c_lo + orow;
            const int n_lo = max(0, seg_c_lo - c_start), n_hi = min(nsteps, seg_c_hi + 1 - c_start);
            const bool any = n_hi > n_lo;
            it_lo = __reduce_min_sync(0xffffffffu, any ? n_lo / J : 0x7fffffff);
            it_hi = __reduce_max_sync(0xffffffffu, any ? (n_hi + J - 1) / J : 0);
        }

        if (producer) {
            const int pt = threadIdx.x - kWsConsumers;             // 0..63
            for (int rb = 0; rb < nrb; rb++) {
                const int b = rb & 1;
                double (*Ks)[L] = reinterpret_cast<double (*)[L]>(s_dyn + b * kBufDoubles);
                double (*Ws)[L] = Ks + kKRows;
                unsigned short *As = reinterpret_cast<unsigned short *>(Ws + kWRows);
                if (rb >= 2) named_sync(3 + b, kWsThreads);       // buffer b released
                // W tile: 4 rows per pass of the 64 threads
                {
                    const int r = rb * L + (pt & 15);
                    const bool active = r < S;
                    const short4 win = active ? s_win[r] : make_short4(0, 0, 0, 0);
                    const bool fix =
                        __any_sync(0xffffffffu, active && (win.z != win.x || win.w != win.y));
                    if (pt == 0) s_fix[b] = fix;
                    const int t0 = pt >> 4;
                    const int wa = win.x - w_lo, wn_rows = active ? win.y - win.x : 0;
                    const double *src = prof + ((long long)half - r + (long long)S * (w_lo + t0));
                    const long long step = (long long)S * 4;
                    unsigned dst = (unsigned)__cvta_generic_to_shared(&Ws[t0][pt & 15]);
                    for (int t = t0; t < wrows; t += 4) {
                        const bool on = (unsigned)(t - wa) < (unsigned)wn_rows;
                        async_copy8s(dst, on ? src : prof, on);
                        src += step;
                        dst += 4 * L * 8;
                    }
                    // anomaly bits of the K rows (only blocks with a differing window)
                    if (fix) {
                        for (int t = pt; t < krows; t += kWsProducers) {
                            const int c = c_lo + t;
                            const long long cell0 = (long long)c * S + rb * L;
                            const bool in = c >= 0 && cell0 < V.onwn;
                            const long long wi = cell0 >> 5;
                            for (int q = 0; q < nset; q++) {
                                unsigned bits = 0u;
                                if (in) {
                                    const unsigned *ab = D.abits[q] + slot_off;
                                    bits = __funnelshift_r(ab[wi], ab[wi + 1], (unsigned)(cell0 & 31));
                                }
                                As[q * kKRows + t] = (unsigned short)(bits & 0xffffu);
                            }
                        }
                    }
                }
                // K tile: 8 rows per pass (16-byte copies), or 4 rows with 8-byte copies
                if ((S & 1) == 0) {
                    const int pr = pt & 7;
                    const int t0 = pt >> 3;
                    const int r2 = rb * L + 2 * pr;
                    const long long base = (long long)c_lo * S + r2;
                    int lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
                    if (r2 < S) {
                        const long long a = max(sa, 0LL);
                        lo0 = first_row(a - base, krows, S, V.fd_tstride);
                        hi0 = first_row(sb - base, krows, S, V.fd_tstride);
                        lo1 = first_row(a - base - 1, krows, S, V.fd_tstride);
                        hi1 = first_row(sb - base - 1, krows, S, V.fd_tstride);
                        if (c_lo < 0) {
                            lo0 = max(lo0, -c_lo);
                            lo1 = max(lo1, -c_lo);
                        }
                    }
                    const double *src = kd + (base + (long long)S * t0);
                    const long long step = (long long)S * 8;
                    unsigned dst = (unsigned)__cvta_generic_to_shared(&Ks[t0][2 * pr]);
                    for (int t = t0; t < krows; t += 8) {
                        const bool on0 = t >= lo0 && t < hi0, on1 = t >= lo1 && t < hi1;
                        if (on0 == on1) {
                            async_copy16s(dst, on0 ? src : kd, on0);
                        } else {
                            async_copy8s(dst, on0 ? src : kd, on0);
                            async_copy8s(dst + 8, on1 ? src + 1 : kd, on1);
                        }
                        src += step;
                        dst += 8 * L * 8;
                    }
                } else {
                    const int r = rb * L + (pt & 15);
                    for (int t = pt >> 4; t < krows; t += 4) {
                        const int c = c_lo + t;
                        const long long cell = (long long)c * S + r;
                        const bool on = r < S && c >= 0 && cell >= sa && cell < sb;
                        async_copy8(&Ks[t][pt & 15], on ? kd + cell : kd, on);
                    }
                }
                async_copy_wait();
                __threadfence_block();
                named_arrive(1 + b, kWsThreads);                   // buffer b is full
            }
        } else {
            for (int rb = 0; rb < nrb; rb++) {
                const int b = rb & 1;
                double (*Ks)[L] = reinterpret_cast<double (*)[L]>(s_dyn + b * kBufDoubles);
                double (*Ws)[L] = Ks + kKRows;
                const unsigned short *As = reinterpret_cast<const unsigned short *>(Ws + kWRows);
                const int r = rb * L + l16;
                const bool active = r < S;
                const short4 win = active ? s_win[r] : make_short4(0, 0, 0, 0);
                named_sync(1 + b, kWsThreads);                     // wait until buffer b is full
                const bool fix = s_fix[b] != 0;

                // (4) the convolution
                dense_walk(Ks, Ws, orow, l16, dmax, w_lo, nsteps, it_lo, it_hi, acc);

                // (5) anomalous cells of this block
                if (fix && active && (win.z != win.x || win.w != win.y)) {
#pragma unroll 1
                    for (int part = 0; part < 2; part++) {
                        const int a = part ? min(win.y, win.w) : min(win.x, win.z);
                        const int b2 = part ? max(win.y, win.w) : max(win.x, win.z);
#pragma unroll 1
                        for (int d = a; d < b2; d++) {
                            const bool in_n = d >= win.x && d < win.y;
                            const bool in_a = d >= win.z && d < win.w;
                            if (in_n == in_a) continue;
                            const long long pi = (long long)half - r + (long long)S * d;
                            if (pi < 0 || pi > 2LL * half) continue;
                            double pv = prof[pi];
                            if (in_n) pv = -pv;
                            const int t0 = orow + (dmax - 1) - d;
                            for (int q = 0; q < nset; q++) {
                                const unsigned short *aq = As + q * kKRows + t0;
                                const double *__restrict__ kq = D.kd[q];
#pragma unroll
                                for (int x = 0; x < J; x++)
                                    if ((aq[x] >> l16) & 1u) {
                                        const long long cell = (long long)(c_lo + t0 + x) * S + r;
                                        if (cell >= sa && cell < sb)
                                            acc[x] = fma(kq[cell], pv, acc[x]);
                                    }
                            }
                        }
                    }
                }
                if (rb + 2 < nrb) named_arrive(3 + b, kWsThreads);  // buffer b may be refilled
            }
        }
    }

    if (producer) return;
    // (6) sum the 16 sub-cell offsets of a half-warp and add to the output row
    double mine = 0.0;
#pragma unroll
    for (int x = 0; x < J; x++) {
        double v = acc[x];
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (l16 == x) mine = v;
    }
    mine = dmul(mine, I.dens);
    if (ms + orow + l16 < m_end) {
        double *dst = out + ((size_t)U.out_index * nrows + row) * (size_t)V.nwave;
        dst[ms + orow + l16] += mine;
    }
}

// ---------------------------------------------------------------------------------------
int launch_anomaly_bits(cudaStream_t st, const StaticView &V, long long gbeg, long long gend,
                        const UnitParams &U, unsigned *bits, int *err) {
    if (gend <= gbeg) return 0;
    const unsigned blocks = (unsigned)((gend - gbeg + 255) / 256);
    anomaly_bits_kernel<<<blocks, 256, 0, st>>>(V, gbeg, gend, U, bits, err);
    PB_CUDA(cudaGetLastError());
    return 0;
}

int launch_densify(cudaStream_t st, const StaticView &V, long long gbeg, long long gend,
                   const double *ksum_tp, const unsigned long long *kmax_entry, double ethresh,
                   double *kd) {
    if (gend <= gbeg) return 0;
    const unsigned blocks = (unsigned)((gend - gbeg + 255) / 256);
    densify_kernel<<<blocks, 256, 0, st>>>(V, gbeg, gend, ksum_tp, kmax_entry, ethresh, kd);
    PB_CUDA(cudaGetLastError());
    return 0;
}

int launch_segment_bounds(cudaStream_t st, const StaticView &V, long long gbeg, long long gend,
                          double adop, int *bounds) {
    segment_bounds_kernel<<<1, 256, sizeof(double) * V.ndop, st>>>(V, gbeg, gend, adop, bounds);
    PB_CUDA(cudaGetLastError());
    return 0;
}

int launch_merge_minor(cudaStream_t st, const StaticView &V, long long gbeg, long long gend,
                       const double *ksum_tp, const unsigned long long *kmax_entry,
                       double ethresh, double adop, const int *main_bounds, double *kd,
                       double *kd_all) {
    if (gend <= gbeg) return 0;
    const unsigned blocks = (unsigned)((gend - gbeg + 255) / 256);
    merge_minor_kernel<<<blocks, 256, sizeof(double) * V.ndop, st>>>(
        V, gbeg, gend, ksum_tp, kmax_entry, ethresh, adop, main_bounds, kd, kd_all);
    PB_CUDA(cudaGetLastError());
    return 0;
}

int launch_accumulate_dense(cudaStream_t st, const StaticView &V, int nunits,
                            const UnitParams *units, const IsoUnit *iso_units,
                            const DenseSet &set, int row, int nrows, long long abits_words,
                            double cutoff, double *out, int *err) {
    if (nunits == 0 || V.nwave == 0) return 0;
    // Default: the two-CTAs-per-SM form (one stages while the other computes).  PB200_DENSE_WS=1
    // selects the warp-specialised one (producer warps + double-buffered tiles), measured 5 %
    // slower on the bench table (1001 vs 953 ms): with staging fully hidden the FMA-pipe share
    // does not rise, i.e. it is set by the walk itself.
    const char *env = std::getenv("PB200_DENSE_WS");
    const bool ws = env && env[0] == '1';
    const size_t smem = ws ? dense_ws_smem_bytes() : dense_smem_bytes();
    if (ws)
        PB_CUDA(cudaFuncSetAttribute(accumulate_dense_ws_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
        PB_CUDA(cudaFuncSetAttribute(accumulate_dense_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = (V.nwave + kDenseTile - 1) / kDenseTile;
    for (int u0 = 0; u0 < nunits; u0 += 65535) {
        const int nu = nunits - u0 < 65535 ? nunits - u0 : 65535;
        dim3 grid((unsigned)nu, (unsigned)ntiles);
        if (ws)
            accumulate_dense_ws_kernel<<<grid, kWsThreads, smem, st>>>(
                V, units + u0, iso_units + (size_t)u0 * V.niso, set, row, nrows, abits_words,
                cutoff, out, err);
        else
            accumulate_dense_kernel<<<grid, kDenseWarps * 32, smem, st>>>(
                V, units + u0, iso_units + (size_t)u0 * V.niso, set, row, nrows, abits_words,
                cutoff, out, err);
        PB_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace pb200

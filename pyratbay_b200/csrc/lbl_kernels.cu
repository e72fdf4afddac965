// lbl_kernels.cu -- strengths, accumulate, counters and table-interpolation kernels.
//
// Kernel 1 (strengths_kernel)   _extcoeff.c:203-226 + the group sums of :248-262
// Kernel 3 (accumulate_kernel)  _extcoeff.c:229-309 + output stage :320-332 (utils.h:119-163)
// counters_kernel               the verbose counters of _extcoeff.c:311-318
// interp_ec_kernel              _extcoeff.c:367-472
//
// Design (DESIGN.md has the long form): the reference scatters every line onto a dynamic
// fine grid and then keeps 1 of `scale` samples (resample) or 2 per output point (linterp).
// Here each thread OWNS one output sample, visits the co-add groups whose footprint covers
// it and gathers exactly the profile samples the reference would have left in that output:
// no atomics, no scratch spectrum, accumulation in a register.  Groups are prepared
// (ethresh test, nearest Doppler profile, index range) 32 at a time, one per lane, staged in
// shared memory and then broadcast to the warp.
#include "lbl_kernels.cuh"
#include "group_prep.cuh"


#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstdint>
#include <vector>

namespace pb200 {

// ---------------------------------------------------------------------------------------
// Kernel 1: line strengths per (T, Z) pass, summed per co-add group, and the per-row maximum.
// The divisions by T and Z are correctly rounded IEEE quotients (quotient_rn: host-rounded
// reciprocal + two FMA residual corrections, bit-identical to a/b), so the strength that
// feeds the `k < ethresh*kmax` predicate equals the strict-IEEE evaluation of the reference's
// expression by construction, not only within an ulp.
#ifndef PB200_STR_MINBLOCKS
#define PB200_STR_MINBLOCKS 4
#endif
// SIGCTE*ratio*gf * exp(-EXPCTE*elow/T) * (1-exp(-EXPCTE*wn/T)) / Z   (_extcoeff.c:219-224)
__device__ __forceinline__ double line_strength(double w, double elow, double gf, double pref,
                                                double t, double inv_t, double z, double inv_z) {
    const double pop = exp(quotient_rn(dmul(-kExpCte, elow), t, inv_t));
    const double ind = dsub(1.0, exp(quotient_rn(dmul(-kExpCte, w), t, inv_t)));
    return quotient_rn(dmul(dmul(dmul(pref, gf), pop), ind), z, inv_z);
}

// Kernel 1.  One thread per in-window LINE, kStrTemps temperature passes per thread:
//  * every line strength is evaluated exactly once whatever the co-add group sizes are (a
//    per-group loop makes all 32 lanes of a warp redo the two exponentials as often as its
//    longest group has members);
//  * a line's wn/elow/gf, group code and isotope are independent coalesced loads issued once
//    and reused for all the thread's temperatures (30 B per line instead of per line x T);
//  * no block barrier: the members of a co-add group are adjacent lines, so the head lane
//    collects them with shuffles in the reference's order k = k_head + k_1 + k_2 ... (:248,
//    258); members that fall into the next warp are evaluated by the head lane itself.
// l_group[line] = group id for a head line, ~group id (negative) for an absorbed member.
// The maximum is taken over every single line (:225): warp REDUX of the IEEE bit pattern
// (kprop >= 0 orders like its bits), one atomic per warp and pass, skipped when the running
// maximum is already larger.
#ifndef PB200_STR_TEMPS
#define PB200_STR_TEMPS 32   // <= 32: lane j keeps the maximum of pass j
#endif
constexpr int kStrTemps = PB200_STR_TEMPS;

__global__ void __launch_bounds__(256, PB200_STR_MINBLOCKS)
strengths_kernel(StaticView V, const int *__restrict__ l_group,
                 const unsigned short *__restrict__ l_iso, long long nlines, int ntp,
                 const double2 *__restrict__ tp_t, const double2 *__restrict__ tp_z,
                 const int *__restrict__ iso_row, int nrows, double *__restrict__ ksum,
                 unsigned long long *__restrict__ kmax) {
    const long long ln = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const long long warp_end = ln - lane + 32;  // first line of the next warp
    double w = 0.0, elow = 0.0, gf = 0.0, pref = 0.0;
    int row = -1, code = INT_MIN, iso = 0;      // INT_MIN: past the end (acts like a head)
    if (ln < nlines) {
        w = V.l_wn[ln];
        elow = V.l_elow[ln];
        gf = V.l_gf[ln];
        code = l_group[ln];
        iso = l_iso[ln];
        row = iso_row[iso];
        if (row >= 0) pref = dmul(kSigCte, V.iso_ratio[iso]);
    }
    // static group geometry inside the warp
    const bool head = code >= 0;
    const unsigned heads = __ballot_sync(0xffffffffu, code >= 0 || code == INT_MIN);
    int nmem = 0;            // members of this head that sit in the following lanes
    unsigned int ext_end = 0;  // one past the group's last line if it runs into the next warp
    if (head) {
        const unsigned above = lane == 31 ? 0u : (heads >> (lane + 1));
        nmem = above ? __ffs(above) - 1 : 31 - lane;
        if (!above && row >= 0) ext_end = V.g_start[code + 1];
    }
    const int maxm = __reduce_max_sync(0xffffffffu, (head && row >= 0) ? nmem : 0);

    const int t0 = blockIdx.y * kStrTemps, t1 = min(ntp, t0 + kStrTemps);
    unsigned long long mine = 0ull;  // lane j keeps the warp maximum of pass t0 + j
    // {T, RN(1/T)} and {Z, RN(1/Z)} of a pass travel as one 16-byte word each
    double2 nx_t = tp_t[t0], nx_z = tp_z[(size_t)t0 * V.niso + iso];
#pragma unroll 1
    for (int tp = t0; tp < t1; tp++) {
        const double2 ct = nx_t, cz = nx_z;
        if (tp + 1 < t1) {  // next pass's scalars, requested one pass ahead
            nx_t = tp_t[tp + 1];
            nx_z = tp_z[(size_t)(tp + 1) * V.niso + iso];
        }
        const double kl =
            row >= 0 ? line_strength(w, elow, gf, pref, ct.x, ct.y, cz.x, cz.y) : 0.0;
        double k = kl;
        for (int jm = 1; jm <= maxm; jm++) {
            const double v = __shfl_down_sync(0xffffffffu, kl, jm);
            if (jm <= nmem) k = dadd(k, v);
        }
        if (head) {
            for (long long m = warp_end; m < (long long)ext_end; m++)
                k = dadd(k, line_strength(V.l_wn[m], V.l_elow[m], V.l_gf[m], pref, ct.x, ct.y,
                                          cz.x, cz.y));
            ksum[(size_t)tp * V.ngroups + code] = k;
        }
        if (nrows == 1) {
            const unsigned hi = (unsigned)__double2hiint(kl), lo = (unsigned)__double2loint(kl);
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
            if (lane == tp - t0) mine = ((unsigned long long)mh << 32) | ml;
        } else if (row >= 0 && kl > 0.0) {
            atomicMax(&kmax[(size_t)tp * nrows + row],
                      (unsigned long long)__double_as_longlong(kl));
        }
    }
    // one atomic per warp and pass, skipped when the running maximum is already larger; lane j
    // handles pass t0 + j, after the arithmetic, so nothing in the loop waits on these reads
    if (nrows == 1 && lane < t1 - t0) {
        if (mine > *(volatile unsigned long long *)&kmax[t0 + lane])
            atomicMax(&kmax[t0 + lane], mine);
    }
}

// l_group / l_iso of every in-window line (static; set_lines).
__global__ void __launch_bounds__(256)
line_group_kernel(const unsigned int *__restrict__ g_start,
                  const unsigned short *__restrict__ g_iso, long long ngroups,
                  int *__restrict__ l_group, unsigned short *__restrict__ l_iso) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    const unsigned int s = g_start[g], e = g_start[g + 1];
    const unsigned short iso = g_iso[g];
    for (unsigned int ln = s; ln < e; ln++) {
        l_group[ln] = ln == s ? (int)g : ~(int)g;
        l_iso[ln] = iso;
    }
}

// ---------------------------------------------------------------------------------------
// Kernel 3: output-owned accumulation.  grid = (tiles, units, rows), 256 threads.
#ifndef PB200_UNROLL
#define PB200_UNROLL 8
#endif
#ifndef PB200_PACKED
#define PB200_PACKED 1   // 16-byte packed staging for constant-step grids (needs tlen < 2^32)
#endif
constexpr unsigned kPackBias = 64;
#ifdef PB200_MINBLOCKS
#define PB200_ACC_BOUNDS __launch_bounds__(256, PB200_MINBLOCKS)
#else
#define PB200_ACC_BOUNDS __launch_bounds__(256)  // 56 registers -> 4 CTAs/SM (measured best)
#endif
#define PB200_STR2(x) #x
#define PB200_STR(x) PB200_STR2(x)
#define PB200_PRAGMA_UNROLL _Pragma(PB200_STR(unroll PB200_UNROLL))
template <int MODE>
__global__ void PB200_ACC_BOUNDS
accumulate_kernel(StaticView V, const UnitParams *__restrict__ units,
                  const IsoUnit *__restrict__ iso_units, const int *__restrict__ iso_row,
                  const double *__restrict__ ksum,
                  const unsigned long long *__restrict__ kmax, int nrows, double ethresh,
                  double cutoff, double *__restrict__ out, int ksplit,
                  double *__restrict__ partial, const int *__restrict__ dense_bounds) {
    extern __shared__ double s_doppler[];  // [ndop]
    // One staged group per lane: {k, byte address of its sample for coordinate 0} and
    // {first coordinate, number of coordinates}.
    __shared__ double2 s_kp[8][32];
    __shared__ int2 s_r[8][32];

    for (int i = threadIdx.x; i < V.ndop; i += blockDim.x)
        s_doppler[i] = V.dop_thr ? V.dop_thr[i] : V.doppler[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const UnitParams U = units[blockIdx.y];
    const int row = blockIdx.z;
    // ksplit CTAs share one tile: CTA `split` takes every ksplit-th chunk of each warp's
    // candidate list and writes a partial sum (reduced by reduce_partials_kernel).  With
    // tiles*ksplit close to the number of resident CTAs, all CTAs in flight work on ONE unit,
    // so that unit's Voigt profiles stay L2-resident while they are gathered.
    const int tile = blockIdx.x / ksplit, split = blockIdx.x - tile * ksplit;
    const int m = tile * kTileOutputs + threadIdx.x;
    const bool in_grid = m < V.nwave;
    const double *__restrict__ table = (MODE == kTransposed) ? V.tprofile : V.profile;
    const int mult = (MODE == kTransposed) ? 1 : U.ofactor;            // address stride per x
    const int fstride = (MODE == kTransposed) ? V.tstride : U.ofactor; // fine samples per x

    // Coordinates this output needs: its own index (kTransposed), the dynamic sample
    // scale*m (kStrided, utils.h:132) or the bracketing pair ilo, ilo+1 (kLinterp, :153-160).
    int x0 = -1, x1 = -1;
    double wn_i = 0.0;
    if (MODE == kLinterp) {
        if (in_grid) {
            wn_i = V.wn[m];
            x0 = (int)ddiv(dsub(wn_i, V.wn0), U.dwnstep);
            x1 = x0 + 1;
        }
    } else if (in_grid && m < U.mcount) {
        x0 = (MODE == kTransposed) ? m : U.scale * m;
    }
    int xmin = (x0 >= 0) ? x0 : INT_MAX;
    int xmax = (x0 >= 0) ? (MODE == kLinterp ? x1 : x0) : INT_MIN;
    xmin = __reduce_min_sync(0xffffffffu, xmin);
    xmax = __reduce_max_sync(0xffffffffu, xmax);
    // byte offset of this lane's sample(s) from a staged group's base address
    const long long off0 = (long long)mult * x0 * 8, off1 = (long long)mult * x1 * 8;
#if PB200_PACKED
    // packed staging (kTransposed): lane l of an active warp owns output xmin + l
    const unsigned lanebit = (MODE == kTransposed && x0 >= 0) ? (1u << (x0 - xmin)) : 0u;
    const double *__restrict__ lane_ptr =
        V.tprofile + (x0 >= 0 ? x0 - xmin : 0) - (long long)kPackBias;
#endif

    double acc0 = 0.0, acc1 = 0.0;
    if (xmax >= xmin) {
        const double *__restrict__ ks = ksum + (size_t)U.tpass * V.ngroups;
        for (int iso = 0; iso < V.niso; iso++) {
            if (iso_row[iso] != row) continue;
            const IsoUnit I = iso_units[(size_t)blockIdx.y * V.niso + iso];
            // Candidate window: first with the unit-wide reach, then tightened with the
            // largest profile any line up to the window's upper end can select (the Doppler
            // index grows with wavenumber; pmaxrow is the running maximum along that axis).
            long long fhi = (long long)xmax * fstride + I.reach;
            int reach = I.reach;
            if (V.ndop >= 2) {
                const double wn_hi = dadd(V.own0, dmul((double)min(fhi, V.onwn - 1), V.ownstep));
                const int nhi = min(V.ndop - 1, 1 + doppler_index(V, s_doppler, dmul(I.adop, wn_hi)));
                reach = min(reach, min(V.pmaxrow[I.ilor * V.ndop + nhi], V.cut_fine) +
                                       2 * U.ofactor + 2);
            }
            long long flo = (long long)xmin * fstride - reach;
            fhi = (long long)xmax * fstride + reach;
            if (fhi < 0 || flo > V.onwn - 1) continue;
            if (flo < 0) flo = 0;
            if (fhi > V.onwn - 1) fhi = V.onwn - 1;
            const int *gb = V.gbin + (size_t)iso * (V.nbins + 1);
            const int glo = gb[V.fd_binw.div((int)flo)];
            const int ghi = gb[V.fd_binw.div((int)fhi) + 1];
            const double kthr =
                dmul(ethresh, __longlong_as_double((long long)kmax[(size_t)U.tpass * nrows + row]));
            // inputs of the first chunk; each later chunk's are requested one chunk ahead so
            // that their latency overlaps the slot loop
            double nx_w = 0.0, nx_k = 0.0;
            int nx_iown = 0;
            const int cstep = 32 * ksplit;
            if (glo + 32 * split + lane < ghi) {
                nx_w = V.g_wn[glo + 32 * split + lane];
                nx_iown = V.g_iown[glo + 32 * split + lane];
                nx_k = ks[glo + 32 * split + lane];
            }
            for (int c = glo + 32 * split; c < ghi; c += cstep) {
                const int g = c + lane;
                const double cur_w = nx_w, cur_k = nx_k;
                const int cur_iown = nx_iown;
                if (g + cstep < ghi) {
                    nx_w = V.g_wn[g + cstep];
                    nx_iown = V.g_iown[g + cstep];
                    nx_k = ks[g + cstep];
                }
                Prep p;
                p.k = 0.0; p.base = 0; p.lo = 0; p.hi = 0;
                // cells at or above dense_from belong to the dense kernel, except the groups
                // of a merged isotope that select another Doppler sample than the main one
                if (g < ghi && (cur_iown < I.dense_from || I.merged)) {
                    int idop = 0;
                    const bool ok = prepare_group<MODE>(V, U, I, s_doppler, kthr, cutoff, cur_w,
                                                        cur_iown, cur_k, &p, &idop);
                    if (ok && I.merged &&
                        in_dense_plane(I, dense_bounds + (size_t)U.tpass * (V.ndop + 1), cur_iown, idop)) {
                        p.k = 0.0; p.lo = 0; p.hi = 0;
                    }
                }
                // clip to the coordinates this warp owns so that empty slots cost nothing more
                const int lo = max(p.lo, xmin), hi = min(p.hi, xmax + 1);
                const int n = min(32, ghi - c);
#if PB200_PACKED
                if (MODE == kTransposed) {
                    // Packed staging (constant-step grids: the warp's outputs are consecutive,
                    // lane l owns output xmin+l): one 16-byte word per group
                    //   {k, table offset of the sample of output xmin (+bias), lane mask}
                    // read with a single broadcast LDS.128; a lane tests its mask bit.
                    unsigned mask = 0u;
                    if (hi > lo) mask = (0xffffffffu >> (32 - (hi - lo))) << (lo - xmin);
                    const unsigned off = (unsigned)(p.base + xmin + (long long)kPackBias);
                    s_kp[warp][lane] = make_double2(
                        p.k, __longlong_as_double((long long)(((unsigned long long)mask << 32) | off)));
                    __syncwarp();
PB200_PRAGMA_UNROLL
                    for (int t = 0; t < n; t++) {
                        const double2 sl = s_kp[warp][t];
                        const unsigned long long bits =
                            (unsigned long long)__double_as_longlong(sl.y);
                        if ((unsigned)(bits >> 32) & lanebit)
                            acc0 = fma(sl.x, __ldg(lane_ptr + (unsigned)bits), acc0);
                    }
                    __syncwarp();
                    continue;
                }
#endif
                s_kp[warp][lane] = make_double2(
                    p.k, __longlong_as_double((long long)(table + p.base)));
                s_r[warp][lane] = make_int2(lo, hi > lo ? hi - lo : 0);
                __syncwarp();
PB200_PRAGMA_UNROLL
                for (int t = 0; t < n; t++) {
                    const int2 r = s_r[warp][t];
                    if ((unsigned)(x0 - r.x) < (unsigned)r.y) {
                        const double2 kp = s_kp[warp][t];
                        const double *src =
                            (const double *)(__double_as_longlong(kp.y) + off0);
                        acc0 = fma(kp.x, __ldg(src), acc0);
                    }
                    if (MODE == kLinterp && (unsigned)(x1 - r.x) < (unsigned)r.y) {
                        const double2 kp = s_kp[warp][t];
                        const double *src =
                            (const double *)(__double_as_longlong(kp.y) + off1);
                        acc1 = fma(kp.x, __ldg(src), acc1);
                    }
                }
                __syncwarp();
            }
        }
    }
    if (in_grid) {
        double *dst = ksplit > 1
            ? partial + (((size_t)blockIdx.y * nrows + row) * ksplit + split) * (size_t)V.nwave
            : out + ((size_t)U.out_index * nrows + row) * (size_t)V.nwave;
        if (MODE == kLinterp) {
            const double wlo = dadd(V.wn0, dmul(U.dwnstep, (double)x0));
            dst[m] = (acc0 * (wlo + U.dwnstep - wn_i) + acc1 * (wn_i - wlo)) / U.dwnstep;
        } else {
            dst[m] = acc0;  // 0 beyond mcount
        }
    }
}

// ---------------------------------------------------------------------------------------
// Kernel 3b: chunk-owned accumulation for constant-step grids (the output-stride table).
//
// The output-owned kernel above spends most of its slots on lanes that lie outside a
// group's footprint: a warp that owns 32 fixed outputs visits every chunk of 32 groups whose
// footprint [a, a+F) overlaps them, i.e. (32+F)/32 warp visits per chunk for F useful lanes.
// Consecutive groups (sorted by wavenumber) have almost the same footprint, so here a warp
// owns a CHUNK of 32 groups instead and places its lanes on the chunk's own footprint
// [lo_min, hi_max): ceil(F/32) passes per chunk, every pass with (nearly) all lanes inside.
// Chunks whose footprint is <= 16 (<= 8) outputs wide run two (four) groups per instruction
// on half (quarter) warps.  A pass accumulates its 32 slots in a register and adds the sum
// once to the warp's PRIVATE copy of the CTA's tile (kChunkTile = 512 outputs) in shared
// memory; the eight copies are summed in a fixed order at the end (no atomics, deterministic).
// The chunk list of a tile (all isotopes) is split evenly over the 8*ksplit warps working
// on the tile, so the work of a CTA is balanced whatever the line density.
// Staged slot (16 bytes, one broadcast LDS.128): {k, table offset of the sample of output
// lo_min (int32), lane mask of the current pass}.  Needs tlen + nwave < 2^31.
#ifndef PB200_CHUNK_UNROLL
#define PB200_CHUNK_UNROLL 8
#endif
#ifndef PB200_CHUNK_MINBLOCKS
#define PB200_CHUNK_MINBLOCKS 4
#endif

// 16-byte load of a staged slot through a 32-bit shared-memory address (a generic pointer makes
// the compiler rebuild the shared window base inside the slot loops).
__device__ __forceinline__ double2 load_slot16(unsigned saddr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(saddr));
    return v;
}

// 8-byte gather of a table sample.  PB200_GATHER_NOALLOC=1 (ld.global.nc.L1::no_allocate) was
// measured and rejected: 2.35 -> 3.48 ms at configs[1], 0.59 -> 1.04 s for the 1e7-line table
// (the L1 hit rate is only 8 %, but the allocating path merges the sectors of a request).
#ifndef PB200_GATHER_NOALLOC
#define PB200_GATHER_NOALLOC 0
#endif
__device__ __forceinline__ double gather_sample(const double *p) {
#if PB200_GATHER_NOALLOC == 1
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#elif PB200_GATHER_NOALLOC == 2
    double v;
    asm("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#elif PB200_GATHER_NOALLOC == 3
    double v;
    asm("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// One empty asm that "modifies" all N values: everything that produces them (the gathers) is
// scheduled before it, everything that consumes them after it.
template <int N>
__device__ __forceinline__ void issue_fence(double (&v)[N]) {
    static_assert(N % 2 == 0 && N <= 16, "unsupported unroll");
    if constexpr (N == 2) {
        asm volatile("" : "+d"(v[0]), "+d"(v[1]));
    } else if constexpr (N == 4) {
        asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]));
    } else if constexpr (N == 6) {
        asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]), "+d"(v[5]));
    } else if constexpr (N == 8) {
        asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]),
                          "+d"(v[5]), "+d"(v[6]), "+d"(v[7]));
    } else if constexpr (N == 10) {
        asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]),
                          "+d"(v[5]), "+d"(v[6]), "+d"(v[7]), "+d"(v[8]), "+d"(v[9]));
    } else if constexpr (N == 12) {
        asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]),
                          "+d"(v[5]), "+d"(v[6]), "+d"(v[7]), "+d"(v[8]), "+d"(v[9]),
                          "+d"(v[10]), "+d"(v[11]));
    } else if constexpr (N == 14) {
        asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]),
                          "+d"(v[5]), "+d"(v[6]), "+d"(v[7]), "+d"(v[8]), "+d"(v[9]),
                          "+d"(v[10]), "+d"(v[11]), "+d"(v[12]), "+d"(v[13]));
    } else {
        static_assert(N == 16 || N <= 14, "unsupported unroll");
        asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]),
                          "+d"(v[5]), "+d"(v[6]), "+d"(v[7]), "+d"(v[8]), "+d"(v[9]),
                          "+d"(v[10]), "+d"(v[11]), "+d"(v[12]), "+d"(v[13]), "+d"(v[14]),
                          "+d"(v[15]));
    }
}

// Run the staged slots of one pass: W lanes per slot, 32/W slots per instruction.
template <int W, int UNR>
__device__ __forceinline__ double run_slots(const double2 *__restrict__ slots, int nslots,
                                            const double *lane_ptr, int lane) {
    constexpr int R = 32 / W;                      // slots per instruction
    constexpr int U = UNR < 32 / R ? UNR : 32 / R;
    unsigned bit = 1u << (lane & (W - 1));
    asm volatile("" : "+r"(bit));  // keep the mask test one LOP3 (not shift + and + compare)
    unsigned mine = (unsigned)__cvta_generic_to_shared(slots + (R > 1 ? lane / W : 0));
    asm volatile("" : "+r"(mine));
    // keep the lane's table pointer in a register pair: one IMAD.WIDE per slot address
    asm volatile("" : "+l"(lane_ptr));
    // slots beyond nslots carry an empty mask, so the trip count is rounded up to the unroll
    const int niter = (((nslots + R - 1) / R) + U - 1) / U * U;
    double acc0 = 0.0, acc1 = 0.0;
    for (int i = 0; i < niter; i += U) {
        // phase 1: issue all U gathers (lanes outside a slot's mask keep a zero sample)...
        double kk[U], vv[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            const double2 s = load_slot16(mine + 16u * (unsigned)((i + j) * R));
            kk[j] = s.x;
            vv[j] = 0.0;
            if ((unsigned)__double2hiint(s.y) & bit) vv[j] = gather_sample(lane_ptr + __double2loint(s.y));
        }
        // ...before the first sample is consumed (U loads in flight per warp; without this
        // fence ptxas interleaves each FMA right behind its load)
        issue_fence(vv);
        // phase 2
#pragma unroll
        for (int j = 0; j < U; j++) {
            if (j & 1) acc1 = fma(kk[j], vv[j], acc1);
            else acc0 = fma(kk[j], vv[j], acc0);
        }
    }
    double acc = acc0 + acc1;
    if (W <= 16) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (W <= 8) acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    return acc;
}

// Slot-major form for chunks that need P passes of 32 outputs (lane l owns outputs
// lo_min + 32p + l, p < P, in registers): ONE broadcast per slot serves all its passes, the
// P gathers of a slot share one address (immediate offsets 256p bytes) and only the first and
// the last pass are predicated -- every staged group starts inside pass 0 and ends inside
// pass P-1 (checked by the caller), so the passes in between are covered by all of them.
// Staged word 3 = a | bL << 8: first lane of pass 0, one-past-last lane of pass P-1.
// The shared-memory/L1 data pipe is what bounds this kernel (2 cycles per broadcast, 2 per
// 256-byte gather): P passes cost 2 + 2P instead of 4P cycles.
template <int P, int UNR>
__device__ __forceinline__ void run_multi(const double2 *__restrict__ slots, int nslots,
                                          const double *lane_ptr, int lane, double (&acc)[P]) {
    constexpr int U = P == 1 ? UNR : (P == 2 ? 8 : (P <= 4 ? 4 : 2));
    asm volatile("" : "+l"(lane_ptr));
    unsigned sbase = (unsigned)__cvta_generic_to_shared(slots);
    asm volatile("" : "+r"(sbase));  // keep it in a register (else rebuilt in every iteration)
    const int niter = (nslots + U - 1) / U * U;  // tail slots: k = 0, offset 0 (valid address)
#pragma unroll
    for (int q = 0; q < P; q++) acc[q] = 0.0;
    double odd = 0.0;  // second chain for P == 1
    for (int i = 0; i < niter; i += U) {
        double kk[U], vv[U * P];
#pragma unroll
        for (int j = 0; j < U; j++) {
            const double2 s = load_slot16(sbase + 16u * (unsigned)(i + j));
            kk[j] = s.x;
            const double *src = lane_ptr + __double2loint(s.y);
            const int w = __double2hiint(s.y);
            const int a = w & 0xff, bl = w >> 8;
#pragma unroll
            for (int q = 0; q < P; q++) {
                bool on = true;
                if (q == 0) on = lane >= a;
                if (q == P - 1) on = on && lane < bl;
                if (q == 0 || q == P - 1) {
                    vv[j * P + q] = 0.0;
                    if (on) vv[j * P + q] = gather_sample(src + 32 * q);
                } else {
                    vv[j * P + q] = gather_sample(src + 32 * q);
                }
            }
        }
        issue_fence(vv);
#pragma unroll
        for (int j = 0; j < U; j++)
#pragma unroll
            for (int q = 0; q < P; q++) {
                if (P == 1 && (j & 1)) odd = fma(kk[j], vv[j * P + q], odd);
                else acc[q] = fma(kk[j], vv[j * P + q], acc[q]);
            }
    }
    if (P == 1) acc[0] += odd;
}

template <int P, int UNR>
__device__ __forceinline__ void run_multi_store(const double2 *slots, int nslots,
                                                const double *lane_ptr, int lane,
                                                double *acc_rel, int span) {
    double acc[P];
    run_multi<P, UNR>(slots, nslots, lane_ptr, lane, acc);
#pragma unroll
    for (int q = 0; q < P; q++)
        if (32 * q + lane < span) acc_rel[32 * q + lane] += acc[q];
}

// Fine cells whose groups of one isotope can reach the outputs [xmin, xmax] of a unit (the unit-
// wide reach, tightened with the running profile maximum along the Doppler axis).
__device__ __forceinline__ bool tile_fine_range(const StaticView &V, const UnitParams &U,
                                                const IsoUnit &I, const double *s_doppler,
                                                int xmin, int xmax, long long *flo_out,
                                                long long *fhi_out) {
    long long fhi = (long long)xmax * V.tstride + I.reach;
    int reach = I.reach;
    if (V.ndop >= 2) {
        const double wn_hi = dadd(V.own0, dmul((double)min(fhi, V.onwn - 1), V.ownstep));
        const int nhi = min(V.ndop - 1, 1 + doppler_index(V, s_doppler, dmul(I.adop, wn_hi)));
        reach = min(reach, min(V.pmaxrow[I.ilor * V.ndop + nhi], V.cut_fine) + 2 * U.ofactor + 2);
    }
    long long flo = (long long)xmin * V.tstride - reach;
    fhi = (long long)xmax * V.tstride + reach;
    if (fhi < 0 || flo > V.onwn - 1) return false;
    *flo_out = flo < 0 ? 0 : flo;
    *fhi_out = fhi > V.onwn - 1 ? V.onwn - 1 : fhi;
    return true;
}

// Groups of isotope `iso` with a fine cell in [clo, chi] (inclusive), at the granularity of the
// coarse index (the caller tests the exact cell).
__device__ __forceinline__ int2 groups_of_cells(const StaticView &V, int iso, long long clo,
                                                long long chi) {
    if (chi < clo) return make_int2(0, 0);
    const int *gb = V.gbin + (size_t)iso * (V.nbins + 1);
    const int glo = gb[V.fd_binw.div((int)clo)], ghi = gb[V.fd_binw.div((int)chi) + 1];
    return ghi > glo ? make_int2(glo, ghi) : make_int2(0, 0);
}

// Two instantiations: <4 CTAs/SM, gather batch 8> for footprints of one or two passes
// (forward models, 1 cm-1 steps) and <3 CTAs/SM, batch 16> for wide footprints (table grids with
// >= 3 passes per line: 517 vs 539 ms at 1e7 lines, but 7 % slower at configs[1]).
template <int MINB, int UNR>
__global__ void __launch_bounds__(256, MINB)
accumulate_chunks_kernel(StaticView V, const UnitParams *__restrict__ units,
                         const IsoUnit *__restrict__ iso_units, const int *__restrict__ iso_row,
                         const double *__restrict__ ksum,
                         const unsigned long long *__restrict__ kmax, int nrows, double ethresh,
                         double cutoff, double *__restrict__ out, int ksplit,
                         double *__restrict__ partial, MergeView M) {
    // dynamic shared memory: [8][kChunkTile] per-warp private copies of the tile, then the
    // Doppler thresholds [ndop]
    extern __shared__ double s_dyn[];
    double (*s_acc)[kChunkTile] = reinterpret_cast<double (*)[kChunkTile]>(s_dyn);
    double *s_doppler = s_dyn + 8 * kChunkTile;
    __shared__ double2 s_slot[8][32];
    // Work list of the tile: entry e < niso = isotope e, its groups below dense_from (all of them
    // for an isotope that is not on the dense path); entries >= niso = (merged minor isotope,
    // Doppler segment j of the MAIN isotope, side): the cells of that segment on which the minor
    // isotope's lines still select a lower sample (side 0: below the start of its own segment j)
    // or already a higher one (side 1) -- the only groups of a merged isotope left to gather
    // (dense_kernels.cu); disjoint by construction.
    __shared__ int2 s_range[kMaxEntries];           // candidate groups [glo, ghi)
    __shared__ int2 s_cells[kMaxEntries];           // exact fine cells [clo, chi) of the entry

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < V.ndop; i += blockDim.x)
        s_doppler[i] = V.dop_thr ? V.dop_thr[i] : V.doppler[i];
#pragma unroll
    for (int i = 0; i < kChunkTile / 32; i++) s_acc[warp][i * 32 + lane] = 0.0;
    __syncthreads();

    const UnitParams U = units[blockIdx.y];
    const int row = blockIdx.z;
    const int tile = blockIdx.x / ksplit, split = blockIdx.x - tile * ksplit;
    const int m0 = tile * kChunkTile;
    const int tile_hi = min(m0 + kChunkTile, min(V.nwave, U.mcount));  // exclusive
    const double *__restrict__ ks = ksum + (size_t)U.tpass * V.ngroups;
    double *acc_tile = s_acc[warp];
    double2 *slots = s_slot[warp];

    // work list, one thread per entry
    const int nzone = M.nminor * 2 * V.ndop;
    const int nent = V.niso + nzone;
    for (int e = threadIdx.x; e < nent; e += blockDim.x) {
        int2 rng = make_int2(0, 0), cells = make_int2(0, 0);
        const int iso = e < V.niso ? e : M.iso[(e - V.niso) / (2 * V.ndop)];
        const IsoUnit I = iso_units[(size_t)blockIdx.y * V.niso + iso];
        long long flo, fhi;
        if (tile_hi > m0 && iso_row[iso] == row &&
            tile_fine_range(V, U, I, s_doppler, m0, tile_hi - 1, &flo, &fhi)) {
            if (e < V.niso) {
                // cells at or above dense_from belong to the dense kernel (or to a zone entry)
                const long long chi = min(fhi, (long long)I.dense_from - 1);
                rng = groups_of_cells(V, iso, flo, chi);
                cells = make_int2((int)flo, (int)min(chi + 1, 0x7fffffffLL));
            } else if (I.merged) {
                const int mi = (e - V.niso) / (2 * V.ndop), rem = (e - V.niso) % (2 * V.ndop);
                const int j = rem >> 1, side = rem & 1;
                const int *bm = M.main_bounds + (size_t)U.tpass * (V.ndop + 1);
                const int *bi = M.minor_bounds + ((size_t)U.tpass * M.nminor + mi) * (V.ndop + 1);
                const int a = side ? max(bm[j], bi[j + 1]) : bm[j];
                const int b = side ? bm[j + 1] : min(bm[j + 1], bi[j]);               // exclusive
                const long long zlo = max((long long)max(a, I.dense_from), flo);
                const long long zhi = min((long long)b - 1, fhi);                     // inclusive
                rng = groups_of_cells(V, iso, zlo, zhi);
                cells = make_int2((int)zlo, (int)(zhi + 1));
            }
        }
        s_range[e] = rng;
        s_cells[e] = cells;
    }
    __syncthreads();

    if (tile_hi > m0) {
        // this warp's share [cb, ce) of the tile's chunk list (all isotopes, concatenated)
        long long total = 0;
        for (int e = 0; e < nent; e++) {
            const int2 r = s_range[e];
            total += (r.y - r.x + 31) >> 5;
        }
        const int nworkers = 8 * ksplit, wid = split * 8 + warp;
        const long long cb = total * wid / nworkers, ce = total * (wid + 1) / nworkers;
        const double kmax_row =
            __longlong_as_double((long long)kmax[(size_t)U.tpass * nrows + row]);
        const double kthr = dmul(ethresh, kmax_row);
        const unsigned lt = (1u << lane) - 1u;

        long long cpos = 0;
        for (int e = 0; e < nent && cpos < ce; e++) {
            const int glo = s_range[e].x, ghi = s_range[e].y;
            const long long nchunk = (ghi - glo + 31) >> 5;
            const long long c0 = max(cb, cpos) - cpos, c1 = min(ce, cpos + nchunk) - cpos;
            cpos += nchunk;
            if (c1 <= c0) continue;
            const int iso = e < V.niso ? e : M.iso[(e - V.niso) / (2 * V.ndop)];
            const int2 cells = s_cells[e];
            const IsoUnit I = iso_units[(size_t)blockIdx.y * V.niso + iso];
            const int gbeg = glo + 32 * (int)c0, gend = min(ghi, glo + 32 * (int)c1);

            // inputs of the first chunk; each later chunk's are requested one chunk ahead
            double nx_w = 0.0, nx_k = 0.0;
            int nx_iown = 0;
            if (gbeg + lane < gend) {
                nx_w = V.g_wn[gbeg + lane];
                nx_iown = V.g_iown[gbeg + lane];
                nx_k = ks[gbeg + lane];
            }
            for (int c = gbeg; c < gend; c += 32) {
                const int g = c + lane;
                const double cur_w = nx_w, cur_k = nx_k;
                const int cur_iown = nx_iown;
                if (g + 32 < gend) {
                    nx_w = V.g_wn[g + 32];
                    nx_iown = V.g_iown[g + 32];
                    nx_k = ks[g + 32];
                }
                Prep p;
                p.k = 0.0; p.base = 0; p.lo = 0; p.hi = 0;
                bool valid = false;
                // only the cells of this entry (the others: dense kernel or another entry)
                if (g < gend && cur_iown >= cells.x && cur_iown < cells.y)
                    valid = prepare_group<kTransposed>(V, U, I, s_doppler, kthr, cutoff, cur_w,
                                                       cur_iown, cur_k, &p);
                const int lo = max(p.lo, m0), hi = min(p.hi, tile_hi);
                valid = valid && hi > lo;
                const unsigned vb = __ballot_sync(0xffffffffu, valid);
                if (vb == 0u) continue;
                const int nval = __popc(vb);
                const int lo_min = __reduce_min_sync(0xffffffffu, valid ? lo : INT_MAX);
                const int hi_max = __reduce_max_sync(0xffffffffu, valid ? hi : INT_MIN);
                const int span = hi_max - lo_min;
                // survivors are compacted to the front; the rest fill the tail with empty masks
                const int pos = valid ? __popc(vb & lt) : nval + __popc(~vb & lt);
                const int a0 = valid ? lo - lo_min : 0, b0 = valid ? hi - lo_min : 0;
                const int off = valid ? (int)(p.base + lo_min) : 0;
                const double kval = valid ? p.k : 0.0;
                double *acc_rel = acc_tile + (lo_min - m0);
                const int npass = (span + 31) >> 5;
                // slot-major form: every group starts in pass 0 and ends in the last pass
                const int bl = b0 - 32 * (npass - 1);
#ifndef PB200_MULTI_MIN_SPAN
#define PB200_MULTI_MIN_SPAN 32   // single-pass chunks: one LOP3 mask test (6 instr/slot vs 9)
#endif
                const bool multi = span > PB200_MULTI_MIN_SPAN && npass <= 8 &&
                                   __all_sync(0xffffffffu, !valid || (a0 < 32 && bl >= 1));
                if (multi) {
                    slots[pos] = make_double2(
                        kval, __hiloint2double(valid ? (a0 | (bl << 8)) : 32, off));
                    __syncwarp();
                    const double *lp = V.tprofile + lane;
                    switch (npass) {
                    case 1: run_multi_store<1, UNR>(slots, nval, lp, lane, acc_rel, span); break;
                    case 2: run_multi_store<2, UNR>(slots, nval, lp, lane, acc_rel, span); break;
                    case 3: run_multi_store<3, UNR>(slots, nval, lp, lane, acc_rel, span); break;
                    case 4: run_multi_store<4, UNR>(slots, nval, lp, lane, acc_rel, span); break;
                    case 5: run_multi_store<5, UNR>(slots, nval, lp, lane, acc_rel, span); break;
                    case 6: run_multi_store<6, UNR>(slots, nval, lp, lane, acc_rel, span); break;
                    case 7: run_multi_store<7, UNR>(slots, nval, lp, lane, acc_rel, span); break;
                    default: run_multi_store<8, UNR>(slots, nval, lp, lane, acc_rel, span); break;
                    }
                } else {
                    unsigned mask = 0u;
                    {
                        const int b = min(b0, 32);
                        if (b > a0) mask = (0xffffffffu >> (32 - (b - a0))) << a0;
                    }
                    slots[pos] = make_double2(kval, __hiloint2double((int)mask, off));
                    __syncwarp();
                    if (span <= 8) {
                        const double acc =
                            run_slots<8, UNR>(slots, nval, V.tprofile + (lane & 7), lane);
                        if (lane < 8 && lane < span) acc_rel[lane] += acc;
                    } else if (span <= 16) {
                        const double acc =
                            run_slots<16, UNR>(slots, nval, V.tprofile + (lane & 15), lane);
                        if (lane < 16 && lane < span) acc_rel[lane] += acc;
                    } else {
                        // generic: one pass of 32 outputs at a time with restaged lane masks
                        for (int x0 = 0; x0 < span; x0 += 32) {
                            if (x0 > 0) {
                                const int a = max(a0, x0) - x0, b = min(b0, x0 + 32) - x0;
                                mask = 0u;
                                if (b > a) mask = (0xffffffffu >> (32 - (b - a))) << a;
                                __syncwarp();
                                reinterpret_cast<int *>(&slots[pos])[3] = (int)mask;
                                __syncwarp();
                                if (!__any_sync(0xffffffffu, mask != 0u)) continue;
                            }
                            const double acc =
                                run_slots<32, UNR>(slots, nval, V.tprofile + x0 + lane, lane);
                            if (x0 + lane < span) acc_rel[x0 + lane] += acc;
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    double *dst = ksplit > 1
        ? partial + (((size_t)blockIdx.y * nrows + row) * ksplit + split) * (size_t)V.nwave
        : out + ((size_t)U.out_index * nrows + row) * (size_t)V.nwave;
#pragma unroll
    for (int t = threadIdx.x; t < kChunkTile; t += 256) {
        const int m = m0 + t;
        if (m < V.nwave) {
            double sum = s_acc[0][t];
#pragma unroll
            for (int w = 1; w < 8; w++) sum += s_acc[w][t];
            dst[m] = sum;  // 0 beyond mcount
        }
    }
}

// Output stage of a constant-R / constant-wavelength grid from a unit's DYNAMIC-grid spectrum
// (utils.h:139-163 linterp, _extcoeff.c:320-332): ext[i] = 2-point interpolation of ktmp at
// wn[i].  ktmp rows [nunits, nrows, dn] come from the chunk kernel run on the dynamic grid
// (engine.cu: units whose dynamic grid is not much finer than the output grid); the unit's
// real output index travels in UnitParams::aslot.
__global__ void __launch_bounds__(256)
linterp_rows_kernel(StaticView V, const UnitParams *__restrict__ units,
                    const double *__restrict__ ktmp, int dn, int nrows, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.nwave) return;
    const UnitParams U = units[blockIdx.y];
    const int row = blockIdx.z;
    const double wn_i = V.wn[i];
    const int x0 = (int)ddiv(dsub(wn_i, V.wn0), U.dwnstep);
    const double *src = ktmp + ((size_t)U.out_index * nrows + row) * (size_t)dn;
    const double k0 = (x0 >= 0 && x0 < dn) ? src[x0] : 0.0;
    const double k1 = (x0 + 1 >= 0 && x0 + 1 < dn) ? src[x0 + 1] : 0.0;
    const double wlo = dadd(V.wn0, dmul(U.dwnstep, (double)x0));
    out[((size_t)U.aslot * nrows + row) * (size_t)V.nwave + i] =
        (k0 * (wlo + U.dwnstep - wn_i) + k1 * (wn_i - wlo)) / U.dwnstep;
}

// Sum the ksplit partial spectra of every (unit, row) in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const UnitParams *__restrict__ units, const double *__restrict__ partial,
                       int nrows, int ksplit, int nwave, double *__restrict__ out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nwave) return;
    const int row = blockIdx.z;
    const double *src = partial + (((size_t)blockIdx.y * nrows + row) * ksplit) * (size_t)nwave + m;
    double acc = src[0];
    for (int s = 1; s < ksplit; s++) acc += src[(size_t)s * nwave];
    out[((size_t)units[blockIdx.y].out_index * nrows + row) * (size_t)nwave + m] = acc;
}

// One-time re-layout of the Voigt table: block of profile p is T[r][q] = P[q*stride + r].
struct TransposeDesc {
    long long src;    // start of the profile in the reference-layout table
    long long dst;    // start of its transposed block
    int nbin;         // 2*half+1
    int rowlen;       // Q
};

__global__ void __launch_bounds__(256)
transpose_kernel(const TransposeDesc *__restrict__ desc, int nprof, long long total, int stride,
                 const double *__restrict__ profile, double *__restrict__ tprofile) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
         gid += step) {
        int lo = 0, hi = nprof;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (desc[mid].dst <= gid) lo = mid; else hi = mid;
        }
        const TransposeDesc d = desc[lo];
        const long long local = gid - d.dst;
        const int r = (int)(local / d.rowlen), q = (int)(local % d.rowlen);
        const long long s = (long long)q * stride + r;
        tprofile[gid] = (s < d.nbin) ? profile[d.src + s] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------
// Counters: nskip, neval, dynamic samples (reference-equivalent work) and samples gathered.
__global__ void __launch_bounds__(256)
counters_kernel(StaticView V, const UnitParams *__restrict__ units,
                const IsoUnit *__restrict__ iso_units, const int *__restrict__ iso_row,
                const double *__restrict__ ksum, const unsigned long long *__restrict__ kmax,
                int nrows, double ethresh, double cutoff, int linterp,
                unsigned long long *__restrict__ counters /* [units,4] */) {
    extern __shared__ double s_doppler[];
    for (int i = threadIdx.x; i < V.ndop; i += blockDim.x)
        s_doppler[i] = V.dop_thr ? V.dop_thr[i] : V.doppler[i];
    __syncthreads();
    const UnitParams U = units[blockIdx.y];
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long skip = 0, eval = 0, dyn = 0, used = 0;
    if (g < V.ngroups) {
        const int iso = V.g_iso[g];
        const int row = iso_row[iso];
        if (row >= 0) {
            const IsoUnit I = iso_units[(size_t)blockIdx.y * V.niso + iso];
            const double kthr =
                dmul(ethresh, __longlong_as_double((long long)kmax[(size_t)U.tpass * nrows + row]));
            Prep p;
            if (prepare_group<kStrided>(V, U, I, s_doppler, kthr, cutoff, V.g_wn[g],
                                        V.g_iown[g], ksum[(size_t)U.tpass * V.ngroups + g],
                                        &p)) {
                eval = 1;
                if (p.hi > p.lo) {
                    dyn = (unsigned long long)(p.hi - p.lo);
                    if (!linterp) {
                        // outputs m < mcount with scale*m in [jlo, jhi)
                        long long mlo = ((long long)p.lo + U.scale - 1) / U.scale;
                        long long mhi = ((long long)p.hi + U.scale - 1) / U.scale;
                        if (mhi > U.mcount) mhi = U.mcount;
                        if (mhi > V.nwave) mhi = V.nwave;
                        if (mhi > mlo) used = (unsigned long long)(mhi - mlo);
                    } else {
                        // outputs whose bracketing pair (ilo, ilo+1) touches [jlo, jhi): counted
                        // on the output grid by bisection; each touching output gathers <= 2.
                        const double a = dadd(V.wn0, dmul(U.dwnstep, (double)(p.lo - 1)));
                        const double b = dadd(V.wn0, dmul(U.dwnstep, (double)p.hi));
                        int lo = 0, hi = V.nwave;
                        while (lo < hi) { int mid = (lo + hi) >> 1; if (V.wn[mid] < a) lo = mid + 1; else hi = mid; }
                        const int first = lo;
                        hi = V.nwave;
                        while (lo < hi) { int mid = (lo + hi) >> 1; if (V.wn[mid] < b) lo = mid + 1; else hi = mid; }
                        used = 2ull * (unsigned long long)(lo - first);
                    }
                }
            } else {
                skip = 1;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        skip += __shfl_xor_sync(0xffffffffu, skip, o);
        eval += __shfl_xor_sync(0xffffffffu, eval, o);
        dyn += __shfl_xor_sync(0xffffffffu, dyn, o);
        used += __shfl_xor_sync(0xffffffffu, used, o);
    }
    if ((threadIdx.x & 31) == 0) {
        unsigned long long *c = counters + (size_t)U.out_index * 4;
        if (skip) atomicAdd(&c[0], skip);
        if (eval) atomicAdd(&c[1], eval);
        if (dyn) atomicAdd(&c[2], dyn);
        if (used) atomicAdd(&c[3], used);
    }
}

// ---------------------------------------------------------------------------------------
// Table interpolation in temperature.  grid = (wave chunks, layers); bit-exact with the
// reference's operation order: ext += (tab_lo*e1 + tab_hi*e2) for each species in turn.
template <int VEC>
__global__ void __launch_bounds__(256)
interp_ec_kernel(double *__restrict__ ext, const double *__restrict__ table,
                 const int *__restrict__ tlo, const double *__restrict__ w_lo,
                 const double *__restrict__ w_hi, const double *__restrict__ density,
                 int nspec, int ntemp, int nlayers, int nwave, int lay1, int per_mol,
                 int overwrite) {
    // VEC = 2: two adjacent samples per thread through 16-byte loads (nwave even, so every row
    // of the table and of ext is 16-byte aligned); VEC = 1: generic.
    const int k = lay1 + blockIdx.y;
    const int lo = tlo[k];
    const double wl = w_lo[k], wh = w_hi[k];
    const size_t tstride = (size_t)nlayers * nwave;  // table[j][t+1] - table[j][t]
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < nwave;
         i += gridDim.x * blockDim.x * VEC) {
        double acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[v] = 0.0;
        // overwrite: the destination counts as zero (what the reference's caller passes in),
        // so no separate memset of a persistent output buffer is needed
        if (!per_mol && !overwrite) {
            if (VEC == 2) {
                const double2 a = *reinterpret_cast<const double2 *>(ext + (size_t)k * nwave + i);
                acc[0] = a.x;
                acc[VEC - 1] = a.y;
            } else {
                acc[0] = ext[(size_t)k * nwave + i];
            }
        }
        for (int j = 0; j < nspec; j++) {
            const double d = density[(size_t)k * nspec + j];
            const double e1 = dmul(wl, d), e2 = dmul(wh, d);
            const double *r = table + (((size_t)j * ntemp + lo) * nlayers + k) * (size_t)nwave + i;
            double a[VEC], b[VEC];
            if (VEC == 2) {
                const double2 a2 = *reinterpret_cast<const double2 *>(r);
                const double2 b2 = *reinterpret_cast<const double2 *>(r + tstride);
                a[0] = a2.x; a[VEC - 1] = a2.y;
                b[0] = b2.x; b[VEC - 1] = b2.y;
            } else {
                a[0] = r[0];
                b[0] = r[tstride];
            }
            double *dst = ext + ((size_t)j * nlayers + k) * (size_t)nwave + i;
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                const double val = dadd(dmul(a[v], e1), dmul(b[v], e2));
                if (per_mol) dst[v] = dadd(overwrite ? 0.0 : dst[v], val);
                else acc[v] = dadd(acc[v], val);
            }
        }
        if (!per_mol) {
            if (VEC == 2)
                *reinterpret_cast<double2 *>(ext + (size_t)k * nwave + i) =
                    make_double2(acc[0], acc[VEC - 1]);
            else
                ext[(size_t)k * nwave + i] = acc[0];
        }
    }
}

// ---------------------------------------------------------------------------------------
// Launch wrappers
int launch_strengths(cudaStream_t st, const StaticView &V, int ntp, const double2 *tp_t,
                     const double2 *tp_z, const int *iso_row, int nrows, double *ksum,
                     unsigned long long *kmax, const int *l_group,
                     const unsigned short *l_iso, long long nlines) {
    if (V.ngroups == 0 || ntp == 0 || nlines == 0) return 0;
    dim3 grid((unsigned)((nlines + 255) / 256), (unsigned)((ntp + kStrTemps - 1) / kStrTemps));
    strengths_kernel<<<grid, 256, 0, st>>>(V, l_group, l_iso, nlines, ntp, tp_t, tp_z, iso_row,
                                           nrows, ksum, kmax);
    PB_CUDA(cudaGetLastError());
    return 0;
}

int launch_line_groups(cudaStream_t st, const unsigned int *g_start,
                       const unsigned short *g_iso, long long ngroups, int *l_group,
                       unsigned short *l_iso) {
    if (ngroups == 0) return 0;
    line_group_kernel<<<(unsigned)((ngroups + 255) / 256), 256, 0, st>>>(g_start, g_iso, ngroups,
                                                                         l_group, l_iso);
    PB_CUDA(cudaGetLastError());
    return 0;
}

int launch_accumulate(cudaStream_t st, const StaticView &V, int nunits,
                      const UnitParams *units, const IsoUnit *iso_units, const int *iso_row,
                      const double *ksum, const unsigned long long *kmax, int nrows,
                      double ethresh, double cutoff, int mode, double *out, int ksplit,
                      double *partial, int chunked, const MergeView &M) {
    if (nunits == 0 || V.nwave == 0) return 0;
    if (ksplit < 1 || !partial) ksplit = 1;
    const int tile_w = (mode == kTransposed && chunked) ? kChunkTile : kTileOutputs;
    const int ntiles = (V.nwave + tile_w - 1) / tile_w;
    dim3 grid((unsigned)(ntiles * ksplit), (unsigned)nunits, (unsigned)nrows);
    const size_t smem = sizeof(double) * V.ndop;
    if (mode == kTransposed && chunked) {
        const size_t csmem = smem + sizeof(double) * 8 * kChunkTile;
        // more than 48 KB of dynamic shared memory needs the per-device opt-in (not the case
        // for 512-output tiles unless the Doppler grid is huge)
        // outputs a line can cover: 2*cutoff/wnstep (+1); three or more passes of 32 -> wide variant
        bool wide = V.tstride > 0 && V.cut_fine != 0x7fffffff &&
                    2LL * V.cut_fine / V.tstride >= 96;
        if (const char *force = std::getenv("PB200_CHUNK_FORM"))   // "narrow" / "wide" (tuning)
            wide = force[0] == 'w';
        auto narrow_k = accumulate_chunks_kernel<PB200_CHUNK_MINBLOCKS, PB200_CHUNK_UNROLL>;
        auto wide_k = accumulate_chunks_kernel<3, 16>;
        if (csmem > 48 * 1024)
            PB_CUDA(cudaFuncSetAttribute(wide ? wide_k : narrow_k,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
        (wide ? wide_k : narrow_k)<<<grid, 256, csmem, st>>>(
            V, units, iso_units, iso_row, ksum, kmax, nrows, ethresh, cutoff, out, ksplit, partial,
            M);
    } else if (mode == kLinterp)
        accumulate_kernel<kLinterp><<<grid, 256, smem, st>>>(
            V, units, iso_units, iso_row, ksum, kmax, nrows, ethresh, cutoff, out, ksplit, partial,
            M.main_bounds);
    else if (mode == kTransposed)
        accumulate_kernel<kTransposed><<<grid, 256, smem, st>>>(
            V, units, iso_units, iso_row, ksum, kmax, nrows, ethresh, cutoff, out, ksplit, partial,
            M.main_bounds);
    else
        accumulate_kernel<kStrided><<<grid, 256, smem, st>>>(
            V, units, iso_units, iso_row, ksum, kmax, nrows, ethresh, cutoff, out, ksplit, partial,
            M.main_bounds);
    PB_CUDA(cudaGetLastError());
    if (ksplit > 1) {
        dim3 rgrid((unsigned)((V.nwave + 255) / 256), (unsigned)nunits, (unsigned)nrows);
        reduce_partials_kernel<<<rgrid, 256, 0, st>>>(units, partial, nrows, ksplit, V.nwave, out);
        PB_CUDA(cudaGetLastError());
    }
    return 0;
}

int launch_linterp_rows(cudaStream_t st, const StaticView &V, int nunits, const UnitParams *units,
                        const double *ktmp, int dn, int nrows, double *out) {
    if (nunits == 0 || V.nwave == 0) return 0;
    dim3 grid((unsigned)((V.nwave + 255) / 256), (unsigned)nunits, (unsigned)nrows);
    linterp_rows_kernel<<<grid, 256, 0, st>>>(V, units, ktmp, dn, nrows, out);
    PB_CUDA(cudaGetLastError());
    return 0;
}

int launch_transpose(cudaStream_t st, int nprof, const long long *src, const long long *dst,
                     const int *nbin, const int *rowlen, long long total, int stride,
                     const double *profile, double *tprofile) {
    if (nprof == 0 || total == 0) return 0;
    std::vector<TransposeDesc> desc(nprof);
    for (int p = 0; p < nprof; p++) desc[p] = TransposeDesc{src[p], dst[p], nbin[p], rowlen[p]};
    TransposeDesc *d_desc = nullptr;
    PB_CUDA(cudaMalloc((void **)&d_desc, sizeof(TransposeDesc) * nprof));
    PB_CUDA(cudaMemcpyAsync(d_desc, desc.data(), sizeof(TransposeDesc) * nprof,
                            cudaMemcpyHostToDevice, st));
    const long long want = (total + 255) / 256;
    const int blocks = (int)std::min<long long>(want, 148LL * 32);
    transpose_kernel<<<blocks, 256, 0, st>>>(d_desc, nprof, total, stride, profile, tprofile);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    cudaFree(d_desc);
    if (err != cudaSuccess) return cuda_fail(err, "transpose_kernel", __FILE__, __LINE__);
    return 0;
}

int launch_counters(cudaStream_t st, const StaticView &V, int nunits, const UnitParams *units,
                    const IsoUnit *iso_units, const int *iso_row, const double *ksum,
                    const unsigned long long *kmax, int nrows, double ethresh, double cutoff,
                    int linterp, unsigned long long *counters) {
    if (nunits == 0 || V.ngroups == 0) return 0;
    dim3 grid((unsigned)((V.ngroups + 255) / 256), (unsigned)nunits);
    counters_kernel<<<grid, 256, sizeof(double) * V.ndop, st>>>(
        V, units, iso_units, iso_row, ksum, kmax, nrows, ethresh, cutoff, linterp, counters);
    PB_CUDA(cudaGetLastError());
    return 0;
}

int launch_interp_ec(cudaStream_t st, double *ext, const double *table, const int *tlo,
                     const double *w_lo, const double *w_hi, const double *density, int nspec,
                     int ntemp, int nlayers, int nwave, int lay1, int lay2, int per_mol,
                     int overwrite) {
    if (lay2 <= lay1 || nwave == 0) return 0;
    const bool vec2 = (nwave % 2 == 0) && ((uintptr_t)ext % 16 == 0) && ((uintptr_t)table % 16 == 0);
    const int per_thread = vec2 ? 2 : 1;
    int bx = (nwave / per_thread + 255) / 256;
    if (bx > 1024) bx = 1024;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)(lay2 - lay1));
    if (vec2)
        interp_ec_kernel<2><<<grid, 256, 0, st>>>(ext, table, tlo, w_lo, w_hi, density, nspec,
                                                  ntemp, nlayers, nwave, lay1, per_mol, overwrite);
    else
        interp_ec_kernel<1><<<grid, 256, 0, st>>>(ext, table, tlo, w_lo, w_hi, density, nspec,
                                                  ntemp, nlayers, nwave, lay1, per_mol, overwrite);
    PB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pb200

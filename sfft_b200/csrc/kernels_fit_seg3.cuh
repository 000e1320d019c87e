// kernels_fit_seg3.cuh -- segmented fit column pass, warp-specialised (KerPolyOrder <= 2 in one launch, 3 in three).
//
// Same mathematics, inputs and outputs as fit_seg_kernel (kernels_fit_seg.cuh).  The CTA has 16 warps:
//   * warps 8..15 ("transform warps", 104 registers after setmaxnreg.dec): one 256-point forward FFT per warp at a
//     time on the 8-values-per-thread engine (fft_vpt.cuh; no CTA barrier inside a transform), jobs
//     (segment, plane-role) handed out round-robin, spectra written into a two-slot ring;
//   * warps 0..7 ("product warps", 152 registers after setmaxnreg.inc): thread = frequency bin, keeps the
//     Fij (Fij + 1) / 2 + Fij complex cross-spectrum accumulators in registers for the whole column, consumes a ring
//     slot as soon as its 2 Fij + 1 spectra are complete, issues the cp.async prefetch of the row windows two
//     segments ahead, and accumulates the column moments for the background cross terms.
// The two groups are decoupled by mbarriers (slot full / slot empty / window landed), so transforms of segment s + 1
// overlap the products of segment s and no warp waits on a CTA-wide barrier inside the segment loop.  One inverse
// transform per pair per column (all 16 warps) yields the lags.
#pragma once
#include <stdio.h>
#include "fft_vpt.cuh"
#include "kernels_fit_seg.cuh"

#define FS3_NT 512
#define FS3_M 256
#define FS3_PITCH 288
// window ring: 8 buffers for fp32 spectra, 4 for fp64 (shared-memory budget); prefetch distance = depth - 2
template <typename TSt> struct Fs3Ring { static const int depth = sizeof(TSt) == 8 ? 8 : 4; };
#ifndef FS3_SLEEP_NS
#define FS3_SLEEP_NS 200
#endif
#define FS3_NMT 64           // product threads per stored plane that also accumulate the column moments (DK <= 2)
// KerPolyOrder = 3 has 5 stored planes and needs the shared memory for the window ring: 32 moment threads per plane
template <int DK> struct Fs3Mom { static const int nmt = DK == 3 ? 32 : 64, npl = DK == 3 ? 5 : 4; };
// the launches ("passes") that cover all plane pairs: planes [A0, A1) in the "A role" against all planes >= A0.
// KerPolyOrder <= 2: one pass; 3 (Fij = 10, 65 accumulators): three passes of 21 / 24 / 20 accumulators
#define FS3_DK3_PASSES {0, 2, 5, 10}

__device__ __forceinline__ unsigned fs3_saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fs3_mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fs3_saddr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void fs3_mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fs3_saddr(b)) : "memory");
}
__device__ __forceinline__ void fs3_mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned done = 0;
    int spins = 0;
    while (true) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(fs3_saddr(b)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(FS3_SLEEP_NS);                  // idle waiters must not eat the issue slots of the transform warps
        if (++spins > (1 << 22)) __trap();          // watchdog: a protocol error must abort, not hang the GPU
    }
}
__device__ __forceinline__ void fs3_cp_async_arrive(unsigned long long* b) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(fs3_saddr(b)) : "memory");
}
// "slot consumed" hand-shake product warps -> transform warps.  NOT an mbarrier: a parity wait is only unambiguous within
// one phase of the barrier, and with few spectra per segment (NP < 8) the consecutive jobs of a transform warp lie more
// than two segments apart (it would have to skip phases).  Every product warp publishes the number of segments it has
// consumed (monotonic, global segment counter); a transform warp may overwrite the slot of segment gs once all eight
// counters have reached gs - 1.
__device__ __forceinline__ void fs3_publish(unsigned* cons_w, unsigned v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(fs3_saddr(cons_w)), "r"(v) : "memory");
}
__device__ __forceinline__ void fs3_wait_consumed(const unsigned* cons, unsigned need) {
    int spins = 0;
    while (true) {
        unsigned a0, a1, a2, a3, b0, b1, b2, b3;
        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(fs3_saddr(cons)) : "memory");
        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(fs3_saddr(cons + 4)) : "memory");
        const unsigned m = min(min(min(a0, a1), min(a2, a3)), min(min(b0, b1), min(b2, b3)));
        if (m >= need) break;
        __nanosleep(FS3_SLEEP_NS);
        if (++spins > (1 << 22)) __trap();
    }
    __threadfence_block();
}
__device__ __forceinline__ void fs3_bar0() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ void fs3_barP() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// column_poly_rows for an explicit thread subset
template <typename TSt>
__device__ void column_poly_rows_sub(const SegFitArgs& fa, const TSt* __restrict__ gI, int k1, const cd* mom, cd* __restrict__ kaprow,
                                     int tid, int nthr, bool jonly = false)
{
    const ColArgs& a = fa.c;
    const double inv0 = 1.0 / (double)a.N0;
    const int np = a.DB + 1;
    for (int idx = tid; idx < (jonly ? 0 : a.Fij * np * a.nlj0); idx += nthr) {
        const int ia = idx % a.nlj0;
        const int p = (idx / a.nlj0) % np;
        const int A = idx / (a.nlj0 * np);
        const int i = a.pl_i[A], j = a.pl_j[A];
        const int sh = ia - a.w0;
        const double beta = sh * inv0;
        cd v = cmake(0, 0);
        for (int e = 0; e <= p; ++e) {
            const double c = binom_small(p, e) * ipow(beta, p - e);
            const cd m = mom[j * SFFTB_MAXE + i + e];
            v.x += c * m.x; v.y += c * m.y;
        }
        if (p > 0 && sh != 0) {
            const TSt* col = gI + ((size_t)j * a.NH + k1) * a.N0;
            const int rbeg = sh > 0 ? a.N0 - sh : 0;
            const int rend = sh > 0 ? a.N0 : -sh;
            const double wrap = sh > 0 ? -1.0 : 1.0;
            for (int r = rbeg; r < rend; ++r) {
                const double cx = (r + 1) * inv0;
                const double corr = ipow(cx + beta + wrap, p) - ipow(cx + beta, p);
                const double c = ipow(cx, i) * corr;
                const cd g = load_c(col + r);
                v.x += c * g.x; v.y += c * g.y;
            }
        }
        for (int q = 0; q + p <= a.DB; ++q) {
            const int pq = fa.pq_of[p][q];
            kaprow[fa.nK + (A * a.Fpq + pq) * a.nlj0 + ia] = cmulcj(v, fa.Q[(size_t)q * a.NH + k1]);
        }
    }
    for (int p = tid; p < np; p += nthr)
        for (int q = 0; q + p <= a.DB; ++q)
            kaprow[fa.nK + fa.nLT + fa.pq_of[p][q]] = cmulcj(mom[a.nj * SFFTB_MAXE + p], fa.Q[(size_t)q * a.NH + k1]);
}

// inverse transform of one accumulated cross spectrum (plane `pl` of the ring) by one warp; keeps the lags of pair `job`
template <int NPAIR>
__device__ __forceinline__ void fs3_inverse_job(const SegFitArgs& fa, const VTabs& vt, cd* plane, int job, int lane, cd* __restrict__ kaprow)
{
    const ColArgs& a = fa.c;
    cd v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = plane[VPAD(lane + 32 * q)];
    __syncwarp();
    vfft<FS3_M>(v, plane, lane, vt, +1.0, 0);
#pragma unroll
    for (int q = 0; q < 8; ++q) plane[VPAD(lane + 32 * q)] = v[q];
    __syncwarp();
    const bool om = job < NPAIR;
    const int lim = om ? 2 * a.w0 : a.w0;
    const int rowbase = om ? job * a.nl0 : fa.nOm + (job - NPAIR) * a.nlj0;
    const double invM = 1.0 / (double)FS3_M;
    for (int l = lane; l <= 2 * lim; l += 32) {
        const int m0 = l - lim;
        kaprow[rowbase + l] = cscale(plane[VPAD(m0 & (FS3_M - 1))], invM);
    }
}

// JONLY (shared-template tiles after the first): the template is unchanged, so only the cross spectra with J and the
// moments of J are recomputed -- Fij "A role" transforms + one of J per segment, Fij accumulators; the rows of the other
// pairs and of the I x T terms stay in `kap` from the first tile of the batch.
// A0 / A1: this launch accumulates the pairs (A, B >= A) and (A, J) for the planes A in [A0, A1) (register budget of the
// product threads); the launch with A0 == 0 also accumulates the column moments and writes the background rows.
template <typename TSt, int DK, bool JONLY = false, int A0 = 0, int A1 = (DK + 1) * (DK + 2) / 2>
__global__ void __launch_bounds__(FS3_NT, 1) fit_seg3_kernel(SegFitArgs fa, VTabs vt_g, const TSt* __restrict__ gI, const TSt* __restrict__ gJ,
                                                             cd* __restrict__ kap)
{
    constexpr int Fij = (DK + 1) * (DK + 2) / 2;
    constexpr int NPAIR = Fij * (Fij + 1) / 2;
    constexpr int NA = A1 - A0;                           // planes transformed in the A role (zero-padded segment)
    constexpr int NB = JONLY ? 0 : Fij - A0;              // planes transformed in the B role (segment + halo)
    constexpr int NPR = JONLY ? 0 : NA * (Fij - A0) - NA * (NA - 1) / 2;   // pairs (A, B >= A) of this launch
    constexpr int NACC = NPR + NA;                        // accumulators kept per product thread
    constexpr int QJ = NPR;                               // first (A, J) accumulator
    constexpr int NP = NA + NB + 1;                       // spectra per segment: A roles | B roles | J
    constexpr int PJ = NP - 1;                            // ring plane of the spectrum of J
    constexpr bool DO_MOM = A0 == 0;
    constexpr int NMT = Fs3Mom<DK>::nmt, NMPL = Fs3Mom<DK>::npl, NMS = NMPL * SFFTB_MAXE;
    static_assert(DK + 2 <= NMPL && NMPL * NMT <= 256, "moment threads");
    // lag-row job of accumulator q: pairs are enumerated `for A for B >= A` over all planes, then the Fij (A, J) rows
    auto job_of = [](int q) -> int {
        if (q >= QJ) return NPAIR + A0 + (q - QJ);
        int ai = 0;
        while (q >= NB - ai) { q -= NB - ai; ++ai; }
        const int A = A0 + ai;
        return A * Fij - A * (A - 1) / 2 + q;
    };
    constexpr int NSRC = DK + 2;
    constexpr int NPL = 2 * NP;                       // planes in the ring (two slots)
    constexpr int NSTG = Fs3Ring<TSt>::depth, LOG2STG = NSTG == 8 ? 3 : 2, PFD = NSTG - 2;
    static_assert(NPL >= 16 || NACC <= NPL, "ring too small for the inverse batches");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ColArgs& a = fa.c;
    cd* spec = reinterpret_cast<cd*>(smem_raw);                                   // NPL (>= 16) planes
    constexpr int NPLA = NPL > 16 ? NPL : 16;
    cd* mom = spec + NPLA * FS3_PITCH;
    cd* macc = mom + NMS;
    cd* tw8 = macc + NMS * NMT;                     // 56 entries  (Ns = 8,  R = 8)
    cd* tw64 = tw8 + 56;                            // 192 entries (Ns = 64, R = 4)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tw64 + 192);
    unsigned* cons = reinterpret_cast<unsigned*>(bars + 12);      // [8] segments consumed by product warp w (fs3_publish)
    TSt* stage = reinterpret_cast<TSt*>(bars + 16);
    unsigned long long* full = bars;          // [2]  count NP   (one arrive per transform job)
    unsigned long long* landed = bars + 4;    // [NSTG] count 256  (cp.async arrivals of the product threads)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double inv0 = 1.0 / (double)a.N0;
    const int h = fa.h, S = fa.S, nseg = fa.nseg;

    if (tid == 0) {
        fs3_mbar_init(full + 0, NP); fs3_mbar_init(full + 1, NP);
        for (int b = 0; b < NSTG; ++b) fs3_mbar_init(landed + b, 256);
    }
    if (tid < 8) cons[tid] = 0u;
    // the engine's twiddle tables live in shared memory: with ~200 KB of shared memory carved out there is next to
    // no L1 left, and a table miss costs an L2 round trip in the middle of a transform
    for (int i = tid; i < 56; i += FS3_NT) tw8[i] = vt_g.t8_8[i];
    for (int i = tid; i < 192; i += FS3_NT) tw64[i] = vt_g.t64_4[i];
    VTabs vt = vt_g;
    vt.t8_8 = tw8; vt.t64_4 = tw64;
    __syncthreads();
    int g = 0;                                // global segment counter (ring phases continue across columns)

    if (warp < 8) {
        // ======================================= product warps =======================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        for (int k1 = blockIdx.x; k1 < a.NH; k1 += gridDim.x, g += nseg) {
            cd acc[NACC];
#pragma unroll
            for (int q = 0; q < NACC; ++q) acc[q] = cmake(0.0, 0.0);
            cd* kaprow = kap + (size_t)k1 * fa.nrows;
            if (DO_MOM) for (int e = tid; e < NMS * NMT; e += 256) macc[e] = cmake(0.0, 0.0);
            // window prefetch: element tid of every stored plane, two segments ahead
            auto issue = [&](int s) {
                const int buf = (g + s) & (NSTG - 1);
                const int r = wrap_row(s * S - h + tid, a.N0);
#pragma unroll
                for (int jj = 0; jj < NSRC; ++jj) {
                    const TSt* col = (jj == DK + 1) ? gJ + (size_t)k1 * a.N0 : gI + ((size_t)jj * a.NH + k1) * a.N0;
                    cp_async_elem(stage + ((size_t)buf * NSRC + jj) * FS3_M + tid, col + r);
                }
                fs3_cp_async_arrive(landed + buf);
            };
            for (int s = 0; s < PFD && s < nseg; ++s) issue(s);
#ifdef FS3_DEBUG
            long long dWaitFull = 0, dProd = 0, dMom = 0, dT0 = clock64();
#endif
            for (int s = 0; s < nseg; ++s) {
                const int gs = g + s, slot = gs & 1;
#ifdef FS3_DEBUG
                long long q0 = clock64();
#endif
                fs3_mbar_wait(full + slot, (gs >> 1) & 1);
#ifdef FS3_DEBUG
                long long q1 = clock64(); dWaitFull += q1 - q0;
#endif
                {
                    const cd* sp = spec + (size_t)slot * NP * FS3_PITCH + VPAD(tid);
                    cd fA[NA];
#pragma unroll
                    for (int A = 0; A < NA; ++A) fA[A] = sp[A * FS3_PITCH];
                    const cd fJ = sp[PJ * FS3_PITCH];
                    if constexpr (!JONLY) {
                        cd fB[NB];                    // fB[bi] is plane A0 + bi, fA[ai] plane A0 + ai: B >= A <=> bi >= ai
#pragma unroll
                        for (int B = 0; B < NB; ++B) fB[B] = sp[(NA + B) * FS3_PITCH];
                        int q = 0;
#pragma unroll
                        for (int A = 0; A < NA; ++A)
#pragma unroll
                            for (int B = A; B < NB; ++B) {
                                acc[q].x = fma(fA[A].x, fB[B].x, acc[q].x); acc[q].x = fma(fA[A].y, fB[B].y, acc[q].x);
                                acc[q].y = fma(fA[A].x, fB[B].y, acc[q].y); acc[q].y = fma(-fA[A].y, fB[B].x, acc[q].y);
                                ++q;
                            }
                    }
#pragma unroll
                    for (int A = 0; A < NA; ++A) {
                        acc[QJ + A].x = fma(fA[A].x, fJ.x, acc[QJ + A].x); acc[QJ + A].x = fma(fA[A].y, fJ.y, acc[QJ + A].x);
                        acc[QJ + A].y = fma(fA[A].x, fJ.y, acc[QJ + A].y); acc[QJ + A].y = fma(-fA[A].y, fJ.x, acc[QJ + A].y);
                    }
                }
#ifdef FS3_DEBUG
                long long q2 = clock64(); dProd += q2 - q1;
#endif
                // column moments of this segment's core rows: 64 threads per stored plane, <= 4 rows each; the slots of
                // a thread are loaded once, updated in registers and stored back (no read-modify-write chains)
                if constexpr (DO_MOM) {
                    const int jj = tid / NMT, mt = tid - jj * NMT;
                    if (jj < NSRC && (!JONLY || jj == DK + 1)) {
                        fs3_mbar_wait(landed + (gs & (NSTG - 1)), (gs >> LOG2STG) & 1);
                        const TSt* st = stage + ((size_t)(gs & (NSTG - 1)) * NSRC + jj) * FS3_M;
                        const int c0 = s * S, Sc = min(S, a.N0 - c0);
                        const int ne = (jj == DK + 1) ? a.DB + 1 : DK - jj + a.DB + 1;
                        cd ma[SFFTB_MAXE];
#pragma unroll
                        for (int e = 0; e < SFFTB_MAXE; ++e) ma[e] = (e < ne) ? macc[(jj * SFFTB_MAXE + e) * NMT + mt] : cmake(0.0, 0.0);
#pragma unroll
                        for (int rr = 0; rr < FS3_M / NMT; ++rr) {
                            // rows are independent; powers of cx first so that only one FMA level depends on the load
                            const int n = h + mt + rr * NMT;
                            const bool live = n < h + Sc;
                            const double cx = (c0 + (n - h) + 1) * inv0;
                            const double cx2 = cx * cx, cx3 = cx2 * cx, cx4 = cx2 * cx2, cx5 = cx4 * cx;
                            const cd gg = live ? load_c(st + (live ? n : 0)) : cmake(0.0, 0.0);
                            const double pw[SFFTB_MAXE] = {1.0, cx, cx2, cx3, cx4, cx5, cx3 * cx3};
#pragma unroll
                            for (int e = 0; e < SFFTB_MAXE; ++e) {
                                if (e < ne) { ma[e].x = fma(gg.x, pw[e], ma[e].x); ma[e].y = fma(gg.y, pw[e], ma[e].y); }
                            }
                        }
#pragma unroll
                        for (int e = 0; e < SFFTB_MAXE; ++e)
                            if (e < ne) macc[(jj * SFFTB_MAXE + e) * NMT + mt] = ma[e];
                    }
                }
#ifdef FS3_DEBUG
                dMom += clock64() - q2;
#endif
                __syncwarp();
                if (lane == 0) fs3_publish(cons + warp, (unsigned)(gs + 1));
                if (s + PFD < nseg) issue(s + PFD);
            }
#ifdef FS3_DEBUG
            if (blockIdx.x == 0 && k1 == blockIdx.x && (tid == 0 || tid == 128))
                printf("P tid %d: loop %lld cycles, wait_full %lld, product %lld, moments %lld (nseg %d)\n", tid, clock64() - dT0, dWaitFull, dProd, dMom, nseg);
#endif
            // ---- column moments -> background cross-term rows (product warps only) ----
            if constexpr (DO_MOM) {
                fs3_barP();
                if (tid < NMS) {
                    cd sm = cmake(0.0, 0.0);
                    for (int t = 0; t < NMT; ++t) sm = cadd(sm, macc[tid * NMT + t]);
                    mom[tid] = sm;
                }
                fs3_barP();
                column_poly_rows_sub(fa, gI, k1, mom, kaprow, tid, 256, JONLY);
            }
            fs3_bar0();                                    // (A) all transforms and products of the column are done
#pragma unroll
            for (int b0 = 0; b0 < NACC; b0 += 16) {
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (b0 + q < NACC) spec[q * FS3_PITCH + VPAD(tid)] = acc[b0 + q];
                fs3_bar0();
                const int job = b0 + warp;
                if (job < NACC) fs3_inverse_job<NPAIR>(fa, vt, spec + warp * FS3_PITCH, job_of(job), lane, kaprow);
                fs3_bar0();
            }
        }
    } else {
        // ====================================== transform warps ======================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
        const int fw = warp - 8;
        for (int k1 = blockIdx.x; k1 < a.NH; k1 += gridDim.x, g += nseg) {
            cd* kaprow = kap + (size_t)k1 * fa.nrows;
#ifdef FS3_DEBUG
            long long fWaitL = 0, fWaitE = 0, fWork = 0, fT0 = clock64(); int fJobs = 0;
#endif
            for (int id = fw; id < nseg * NP; id += 8) {
                const int s = id / NP, p = id - s * NP;
                const int gs = g + s, slot = gs & 1;
                const bool roleA = p < NA, isJ = p == PJ;
                const int pl = roleA ? A0 + p : (isJ ? 0 : A0 + (p - NA));
                const int my_i = isJ ? 0 : a.pl_i[pl];
                const int my_src = isJ ? DK + 1 : a.pl_j[pl];
                const int c0 = s * S, Sc = min(S, a.N0 - c0);
#ifdef FS3_DEBUG
                long long f0 = clock64();
#endif
                fs3_mbar_wait(landed + (gs & (NSTG - 1)), (gs >> LOG2STG) & 1);
#ifdef FS3_DEBUG
                long long f1 = clock64(); fWaitL += f1 - f0;
#endif
                if (gs >= 2) fs3_wait_consumed(cons, (unsigned)(gs - 1));      // segment gs - 2 (same slot) consumed by every product warp
                const TSt* src = stage + ((size_t)(gs & (NSTG - 1)) * NSRC + my_src) * FS3_M;
                cd* plane = spec + ((size_t)slot * NP + p) * FS3_PITCH;
                cd v[8];
                // cx of window position n = lane + 32 q: one int->double conversion per job, the rest by FMA; rows that
                // wrapped around the column ends (first / last segments only) are shifted by one period
                const int row_l = c0 - h + lane;
                const double cx_l = (double)(row_l + 1) * inv0;
                const bool simple = (a.N0 >= 2 * FS3_M);          // at most one wrap per window
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int n = lane + 32 * q;
                    cd gg = load_c(src + n);
                    if (my_i > 0) {
                        double cx;
                        if (simple) {
                            const int row = row_l + 32 * q;
                            cx = fma((double)(32 * q), inv0, cx_l);
                            cx += (row < 0) ? 1.0 : ((row >= a.N0) ? -1.0 : 0.0);
                        } else {
                            cx = (wrap_row(row_l + 32 * q, a.N0) + 1) * inv0;
                        }
                        gg = cscale(gg, my_i == 1 ? cx : (my_i == 2 ? cx * cx : cx * cx * cx));
                    }
                    const bool keep = !roleA || (n >= h && n < h + Sc);
                    v[q] = keep ? gg : cmake(0.0, 0.0);
                }
                vfft<FS3_M>(v, plane, lane, vt, -1.0, 0);
#pragma unroll
                for (int q = 0; q < 8; ++q) plane[VPAD(lane + 32 * q)] = v[q];
                __syncwarp();
                if (lane == 0) fs3_mbar_arrive(full + slot);
#ifdef FS3_DEBUG
                fWork += clock64() - f2; ++fJobs;
#endif
            }
#ifdef FS3_DEBUG
            if (blockIdx.x == 0 && k1 == blockIdx.x && lane == 0 && (fw == 0 || fw == 7))
                printf("F warp %d: loop %lld cycles, %d jobs, wait_landed %lld, wait_empty %lld, work %lld\n", fw, clock64() - fT0, fJobs, fWaitL, fWaitE, fWork);
#endif
            fs3_bar0();                                    // (A)
#pragma unroll
            for (int b0 = 0; b0 < NACC; b0 += 16) {
                fs3_bar0();
                const int job = b0 + warp;
                if (job < NACC) fs3_inverse_job<NPAIR>(fa, vt, spec + warp * FS3_PITCH, job_of(job), lane, kaprow);
                fs3_bar0();
            }
        }
    }
}

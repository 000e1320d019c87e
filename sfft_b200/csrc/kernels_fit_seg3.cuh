// kernels_fit_seg3.cuh -- shared pieces of the warp-specialised segmented fit kernels (ring geometry, mbarrier / counter
// hand-shakes, background cross-term rows).  The kernels themselves are fit_seg4_kernel (kernels_fit_seg4.cuh) and
// fit_gen4_kernel (kernels_gen.cuh), both on the half-warp FFT engine; the round-1 kernel on the warp-wide 8-value engine that
// this header used to hold (fit_seg3_kernel) is gone.  Its design notes are kept because the ring protocol is unchanged:
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

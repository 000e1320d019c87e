// kernels_fit_seg4.cuh -- segmented fit column pass, warp-specialised, on the half-warp 16-value FFT engine (fft_h16.cuh).
//
// Same mathematics, inputs, outputs, ring protocol and shared-memory layout as fit_seg3_kernel (kernels_fit_seg3.cuh, which
// stays as the engine of the general-basis path).  What changed is the transform side: fit_seg3 is bound by the shared-memory
// exchanges of its warp-wide 8-value transforms (two exchanges per transform, 250 cycles per transform per SM); here a
// transform belongs to HALF a warp (16 lanes x 16 values, ONE exchange, twiddles in registers: 126 cycles per transform per
// SM with eight half-warp pairs in flight, scripts/micro/hfft_bench.cu), so the CTA has
//   * warps 0..7  ("product warps"): thread = frequency bin, cross-spectrum accumulators in registers, window prefetch
//     (cp.async ring), column moments -- unchanged;
//   * warps 8..11 ("transform warps"): 8 half-warp workers, jobs (segment, plane-role) handed out round-robin; the two halves
//     of a warp run their transforms in lockstep (__syncwarp only).
// 384 threads, 168 registers each (no setmaxnreg).  The inverse transforms at the end of a column run on all 24 half warps,
// and the lags are written straight from registers.
#pragma once
#include "fft_h16.cuh"
#include "kernels_fit_seg3.cuh"

#ifndef FS4_NTW
#define FS4_NTW 8
#endif
#define FS4_NT (256 + 32 * FS4_NTW)
#ifndef FS4_REGT
#define FS4_REGT 120
#endif
#define FS4_REGP (256 - FS4_REGT)
#define FS4_NTW_UNUSED                 // transform warps
#define FS4_NWK (2 * FS4_NTW)     // half-warp transform workers
#define FS4_NHW (2 * (FS4_NT / 32))   // half warps of the CTA (inverse phase)
#define FS4_PITCH H16_SCRATCH         // elements per spectrum plane
#ifndef FS4_INV_ALL
#define FS4_INV_ALL 0                 // 1: the inverse phase of a column uses all 32 half warps (one round)
#endif
#ifndef FS4_BULK
#define FS4_BULK 1                    // window ring filled by cp.async.bulk (0: per-element cp.async of the product threads)
#endif
#ifndef FS4_NSTG
#define FS4_NSTG 8                    // window ring depth for fp32-stored spectra (prefetch distance NSTG - 2 segments); fp64: 4
#endif
#ifndef FS4_NSLOT32
#define FS4_NSLOT32 2
#endif
// ring geometry: fp32-stored spectra with KerPolyOrder <= 2 take THREE spectrum slots (the transforms of segment s + 2 start under the
// products of s) and six windows (219 KB; with four windows the third slot measured slower, with six 0.73 against 0.75 ms at C2); every
// other case two slots, fp32 eight / fp64 four windows (their planes or windows are larger)
#ifndef FS4_RING3
#define FS4_RING3 1
#endif
template <typename TSt, int DK = 3> struct Fs4Ring {
    static const bool three = FS4_RING3 && sizeof(TSt) == 8 && DK <= 2;
    static const int nslot = three ? 3 : (sizeof(TSt) == 8 ? FS4_NSLOT32 : 2);
    static const int depth = sizeof(TSt) == 8 ? (three ? 6 : FS4_NSTG) : 4;
};
// dynamic shared memory of fit_seg4_kernel for the largest instantiation of a plan (host side)
static inline size_t fs4_smem_bytes(int DK, bool f32) {
    const int Fij = (DK + 1) * (DK + 2) / 2;
    const int NP = DK == 3 ? 13 : 2 * Fij + 1, NACC = DK == 3 ? 24 : Fij * (Fij + 1) / 2 + Fij;
    const bool three = FS4_RING3 && f32 && DK <= 2;
    const int planes = std::max((three ? 3 : (f32 ? FS4_NSLOT32 : 2)) * NP, NACC);
    return sizeof(cd) * ((size_t)planes * FS4_PITCH + (DK == 3 ? 5 : 4) * SFFTB_MAXE) + 128 +
           (f32 ? sizeof(float2) * (three ? 6 : FS4_NSTG) : sizeof(double2) * 4) * (size_t)(DK + 2) * FS3_M;
}

// TMA-style bulk copies (cp.async.bulk, the 1-D form of the tensor memory accelerator's copy engine): a 256-row window of a
// stored plane is one contiguous 2 KB (fp32) / 4 KB (fp64) piece of a column, so ONE elected thread requests the whole window
// ring slot -- (DK + 2) planes, two pieces where the window wraps around the column ends -- and the bytes are counted on the
// slot's mbarrier (expect_tx / complete_tx).  The product warps issue no per-element cp.async any more.
__device__ __forceinline__ void fs4_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fs3_saddr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fs4_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(fs3_saddr(dst)), "l"(src), "r"(bytes), "r"(fs3_saddr(b)) : "memory");
}

// Wait until window `seg` (global segment number) of the cp.async ring has landed.  The mbarrier answers "has the phase
// of this parity completed", which is ambiguous once the barrier is two phases away from the question: a worker whose
// consecutive jobs lie more than a ring length apart (few spectra per segment) may ask about a window that landed two
// phases ago and would be told "not yet" until the phase after next.  The product warps' monotonic consumed counters
// settle that case: a consumed segment has certainly landed.  (The other direction -- a stale "yes" -- cannot happen
// because every worker walks through the windows in order.)
__device__ __forceinline__ void fs4_wait_landed(unsigned long long* b, unsigned parity, const unsigned* cons, unsigned seg) {
    int spins = 0;
    while (true) {
        unsigned done = 0;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(fs3_saddr(b)), "r"(parity) : "memory");
        if (done) break;
        unsigned a0, a1, a2, a3, b0, b1, b2, b3;
        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(fs3_saddr(cons)) : "memory");
        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(fs3_saddr(cons + 4)) : "memory");
        if (min(min(min(a0, a1), min(a2, a3)), min(min(b0, b1), min(b2, b3))) > seg) break;
        __nanosleep(FS3_SLEEP_NS);
        if (++spins > (1 << 22)) __trap();
    }
}

// inverse transform of one accumulated cross spectrum (plane of the ring) by one half warp; keeps the lags of pair `job`
template <int NPAIR>
__device__ __forceinline__ void fs4_inverse_job(const SegFitArgs& fa, const H16Tw& tw, cd* plane, int job, int hl, bool active,
                                                cd* __restrict__ kaprow)
{
    const ColArgs& a = fa.c;
    cd v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = active ? plane[HPAD(hl + 16 * q)] : cmake(0.0, 0.0);
    __syncwarp();
    hfft256(v, plane, hl, tw, +1.0, active);
    if (!active) return;
    const bool om = job < NPAIR;
    const int lim = om ? 2 * a.w0 : a.w0;
    const int rowbase = om ? job * a.nl0 : fa.nOm + (job - NPAIR) * a.nlj0;
    const double invM = 1.0 / (double)FS3_M;
    // v[q] = X[hl + 16 q]; lag m0 lives at index m0 & 255, |m0| <= lim < 128
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int idx = hl + 16 * q;
        const int m0 = idx < FS3_M / 2 ? idx : idx - FS3_M;
        if (m0 >= -lim && m0 <= lim) kaprow[rowbase + m0 + lim] = cscale(v[q], invM);
    }
}

// Column moments  mom[k1][jj][e] = sum_r cx(r)^e g_jj[k1][r]  of the stored planes (jj <= DK: planes of I, jj = DK + 1: J) for
// the cross terms with the background basis (column_poly_rows_sub): one warp per (column, plane), coalesced reads, shuffle
// reduction.  fit_seg3_kernel sums them from the staged windows inside its product warps; here they are a kernel of their own
// (0.05 ms at 4096^2), which takes 40 % of the instructions, 42 registers and the read-modify-write traffic out of the product
// loop.  nms = planes-per-column stride * SFFTB_MAXE of the fit kernel's layout; jonly: the moments of J only.
template <int NE, typename TSt>
__device__ __forceinline__ void col_moments_sum(const TSt* __restrict__ col, int N0, int lane, cd (&m)[SFFTB_MAXE]) {
    const double inv0 = 1.0 / (double)N0;
    double rd = (double)(lane + 1);                      // r + 1 as a double: exact, and no int -> double conversion per row
#pragma unroll 4
    for (int r = lane; r < N0; r += 32, rd += 32.0) {
        const cd g = load_c(col + r);
        const double cx = rd * inv0;
        double pw = 1.0;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            m[e].x = fma(g.x, pw, m[e].x); m[e].y = fma(g.y, pw, m[e].y);
            if (e + 1 < NE) pw *= cx;
        }
    }
}

template <typename TSt>
__global__ void __launch_bounds__(256) col_moments_kernel(int N0, int NH, int DK, int DB, int nms, int jonly, const TSt* __restrict__ gI,
                                                          const TSt* __restrict__ gJ, cd* __restrict__ momg)
{
    const int nsrc = DK + 2;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= NH * nsrc) return;
    const int k1 = gw / nsrc, jj = gw - k1 * nsrc;
    if (jonly && jj != DK + 1) return;
    const TSt* col = (jj == DK + 1) ? gJ + (size_t)k1 * N0 : gI + ((size_t)jj * NH + k1) * N0;
    const int ne = (jj == DK + 1) ? DB + 1 : DK - jj + DB + 1;      // uniform per warp: the row loop is compiled per count
    cd m[SFFTB_MAXE];
#pragma unroll
    for (int e = 0; e < SFFTB_MAXE; ++e) m[e] = cmake(0.0, 0.0);
    switch (ne) {
        case 1: col_moments_sum<1>(col, N0, lane, m); break;
        case 2: col_moments_sum<2>(col, N0, lane, m); break;
        case 3: col_moments_sum<3>(col, N0, lane, m); break;
        case 4: col_moments_sum<4>(col, N0, lane, m); break;
        case 5: col_moments_sum<5>(col, N0, lane, m); break;
        case 6: col_moments_sum<6>(col, N0, lane, m); break;
        default: col_moments_sum<7>(col, N0, lane, m); break;
    }
#pragma unroll
    for (int e = 0; e < SFFTB_MAXE; ++e) {
        if (e < ne) {
            const double sx = warp_sum(m[e].x), sy = warp_sum(m[e].y);
            if (lane == 0) momg[(size_t)k1 * nms + jj * SFFTB_MAXE + e] = cmake(sx, sy);
        }
    }
}

// Background cross-term rows (I x T, J x T) of every column from the column moments: one CTA per column.  Inside fit_seg4_kernel this
// work sat at the start of every column on the product warps (wrap-row corrections read the column ends from global memory, an L2
// round trip per row, while the transform warps had filled their two ring slots and waited); as a kernel of its own it costs the
// product warps nothing.
template <typename TSt, int DK>
__global__ void __launch_bounds__(256) col_poly_rows_kernel(SegFitArgs fa, const TSt* __restrict__ gI, cd* __restrict__ kap, int jonly)
{
    constexpr int NMS = Fs3Mom<DK>::npl * SFFTB_MAXE;
    __shared__ cd mom[NMS];
    const int tid = threadIdx.x;
    for (int k1 = blockIdx.x; k1 < fa.c.NH; k1 += gridDim.x) {
        if (tid < NMS) mom[tid] = fa.momg[(size_t)k1 * NMS + tid];
        __syncthreads();
        column_poly_rows_sub(fa, gI, k1, mom, kap + (size_t)k1 * fa.nrows, tid, 256, jonly != 0);
        __syncthreads();
    }
}

// JONLY (shared-template tiles after the first): the template is unchanged, so only the cross spectra with J and the
// moments of J are recomputed -- Fij "A role" transforms + one of J per segment, Fij accumulators; the rows of the other
// pairs and of the I x T terms stay in `kap` from the first tile of the batch.
// A0 / A1: this launch accumulates the pairs (A, B >= A) and (A, J) for the planes A in [A0, A1) (register budget of the
// product threads); the launch with A0 == 0 also accumulates the column moments and writes the background rows.
template <typename TSt, int DK, bool JONLY = false, int A0 = 0, int A1 = (DK + 1) * (DK + 2) / 2>
__global__ void __launch_bounds__(FS4_NT, 1) fit_seg4_kernel(SegFitArgs fa, VTabs vt_g, const TSt* __restrict__ gI, const TSt* __restrict__ gJ,
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
    constexpr int NSLOT = Fs4Ring<TSt, DK>::nslot;
    constexpr int NPL = NSLOT * NP;                   // planes in the spectrum ring
    constexpr int NSTG = Fs4Ring<TSt, DK>::depth, PFD = NSTG - 2;   // any depth, not only powers of two
    static_assert(NACC <= FS4_NHW, "one half warp per accumulator in the inverse phase");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ColArgs& a = fa.c;
    cd* spec = reinterpret_cast<cd*>(smem_raw);                                   // NPL (>= 16) planes
    constexpr int NPLA = NPL > NACC ? NPL : NACC;     // every accumulator gets a plane: ONE inverse batch per column
    cd* mom = spec + NPLA * FS4_PITCH;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(mom + NMS);     // (less than the fit_seg3 layout the host sizes)
    unsigned* cons = reinterpret_cast<unsigned*>(bars + 12);      // [8] segments consumed by product warp w (fs3_publish)
    TSt* stage = reinterpret_cast<TSt*>(bars + 16);
    unsigned long long* full = bars;          // [NSLOT] count NP (one arrive per transform job)
    unsigned long long* landed = bars + 4;    // [NSTG] count 256  (cp.async arrivals of the product threads)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double inv0 = 1.0 / (double)a.N0;
    const int h = fa.h, S = fa.S, nseg = fa.nseg;

    // bulk window copies need 16-byte aligned pieces (even N0 for 8-byte elements) and at most one wrap per window
    const bool bulk = a.N0 >= FS3_M && (sizeof(TSt) == 16 || a.N0 % 2 == 0) && FS4_BULK;
    if (tid == 0) {
        for (int b = 0; b < NSLOT; ++b) fs3_mbar_init(full + b, NP);
        for (int b = 0; b < NSTG; ++b) fs3_mbar_init(landed + b, bulk ? 1 : 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");       // the copy engine (async proxy) completes bytes on them
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < 8) cons[tid] = 0u;
    // the engine's twiddle tables live in shared memory: with ~200 KB of shared memory carved out there is next to
    // no L1 left, and a table miss costs an L2 round trip in the middle of a transform
    (void)vt_g;
    const int half = lane >> 4, hl = lane & 15;
    __syncthreads();
    int g = 0;                                // global segment counter (ring phases continue across columns)

    if (warp < 8) {
        // ======================================= product warps =======================================
#if FS4_NTW == 8
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FS4_REGP));
#endif
        int kprev = -1;
        for (int k1 = blockIdx.x; k1 < a.NH; k1 += gridDim.x, g += nseg) {
            cd acc[NACC];
#pragma unroll
            for (int q = 0; q < NACC; ++q) acc[q] = cmake(0.0, 0.0);
            cd* kaprow = kap + (size_t)k1 * fa.nrows;
            // window prefetch: element tid of every stored plane, PF segments ahead; the ring runs on across the column
            // boundary (the first windows of the CTA's next column are requested during the last segments of this one)
            auto issue = [&](int kx, int gx, int s) {
                const int buf = (int)((unsigned)(gx + s) % (unsigned)NSTG);
                if (bulk) {
                    if (tid == 0) {
                        const int r0 = s * S - h;
                        fs4_expect_tx(landed + buf, (unsigned)(NSRC * FS3_M * sizeof(TSt)));
#pragma unroll
                        for (int jj = 0; jj < NSRC; ++jj) {
                            const TSt* col = (jj == DK + 1) ? gJ + (size_t)kx * a.N0 : gI + ((size_t)jj * a.NH + kx) * a.N0;
                            TSt* dst = stage + ((size_t)buf * NSRC + jj) * FS3_M;
                            if (r0 < 0) {
                                fs4_bulk_g2s(dst, col + a.N0 + r0, (unsigned)(-r0 * (int)sizeof(TSt)), landed + buf);
                                fs4_bulk_g2s(dst - r0, col, (unsigned)((FS3_M + r0) * (int)sizeof(TSt)), landed + buf);
                            } else if (r0 + FS3_M > a.N0) {
                                const int n1 = a.N0 - r0;
                                fs4_bulk_g2s(dst, col + r0, (unsigned)(n1 * (int)sizeof(TSt)), landed + buf);
                                fs4_bulk_g2s(dst + n1, col, (unsigned)((FS3_M - n1) * (int)sizeof(TSt)), landed + buf);
                            } else {
                                fs4_bulk_g2s(dst, col + r0, (unsigned)(FS3_M * sizeof(TSt)), landed + buf);
                            }
                        }
                    }
                    return;
                }
                const int r = wrap_row(s * S - h + tid, a.N0);
#pragma unroll
                for (int jj = 0; jj < NSRC; ++jj) {
                    const TSt* col = (jj == DK + 1) ? gJ + (size_t)kx * a.N0 : gI + ((size_t)jj * a.NH + kx) * a.N0;
                    cp_async_elem(stage + ((size_t)buf * NSRC + jj) * FS3_M + tid, col + r);
                }
                fs3_cp_async_arrive(landed + buf);
            };
            const int PF = min(PFD, nseg);
            if (k1 == (int)blockIdx.x) for (int s = 0; s < PF; ++s) issue(k1, g, s);
            // background cross-term rows of the PREVIOUS column: off the critical path (the transform warps are busy with the
            // first segments of this column)
            if constexpr (DO_MOM) {
                if (kprev >= 0 && !fa.mom_external) {
                    if (tid < NMS) mom[tid] = fa.momg[(size_t)kprev * NMS + tid];
                    fs3_barP();
                    column_poly_rows_sub(fa, gI, kprev, mom, kap + (size_t)kprev * fa.nrows, tid, 256, JONLY);
                }
            }
#ifdef FS4_DEBUG
            long long dWaitFull = 0, dProd = 0, dMom = 0, dT0 = clock64();
#endif
            for (int s = 0; s < nseg; ++s) {
                const int gs = g + s, slot = gs % NSLOT;
#ifdef FS4_DEBUG
                long long q0 = clock64();
#endif
                fs3_mbar_wait(full + slot, (gs / NSLOT) & 1);
#ifdef FS4_DEBUG
                long long q1 = clock64(); dWaitFull += q1 - q0;
#endif
                {
                    const cd* sp = spec + (size_t)slot * NP * FS4_PITCH + HPAD(tid);
                    cd fA[NA];
#pragma unroll
                    for (int A = 0; A < NA; ++A) fA[A] = sp[A * FS4_PITCH];
                    const cd fJ = sp[PJ * FS4_PITCH];
                    if constexpr (!JONLY) {
                        cd fB[NB];                    // fB[bi] is plane A0 + bi, fA[ai] plane A0 + ai: B >= A <=> bi >= ai
#pragma unroll
                        for (int B = 0; B < NB; ++B) fB[B] = sp[(NA + B) * FS4_PITCH];
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
#ifdef FS4_DEBUG
                dProd += clock64() - q1;
#endif
                // the spectra of this slot are in registers / accumulated: release the slot BEFORE the moments, so that the
                // transform warps start on segment gs + 2 while the moments of gs are still being summed
                __syncwarp();
                if (lane == 0) fs3_publish(cons + warp, (unsigned)(gs + 1));
                if (s + PF < nseg) issue(k1, g, s + PF);
                else if (k1 + (int)gridDim.x < a.NH) issue(k1 + gridDim.x, g + nseg, s + PF - nseg);
            }
#ifdef FS4_DEBUG
            if (blockIdx.x == 0 && k1 == blockIdx.x + gridDim.x && (tid == 0 || tid == 128))
                printf("P tid %d: loop %lld cycles, wait_full %lld, product %lld, moments %lld (nseg %d)\n", tid, clock64() - dT0, dWaitFull, dProd, dMom, nseg);
#endif
            // ---- column moments -> background cross-term rows (product warps only) ----
            fs3_bar0();                                    // (A) all transforms and products of the column are done
#pragma unroll
            for (int q = 0; q < NACC; ++q) spec[q * FS4_PITCH + HPAD(tid)] = acc[q];
            fs3_bar0();
#if FS4_INV_ALL
            {
                H16Tw htw;                                 // (reloaded per column: 16 registers the product loop cannot spare)
                h16_load(htw, fa.tabA, hl);
                const int wk = 2 * warp + half;
                const bool act = wk < NACC;
                if (__any_sync(0xffffffffu, act))
                    fs4_inverse_job<NPAIR>(fa, htw, spec + (act ? wk : 0) * FS4_PITCH, job_of(act ? wk : 0), hl, act, kaprow);
            }
#endif
            fs3_bar0();
            kprev = k1;
        }
        if constexpr (DO_MOM) {
            // (the column moments come from col_moments_kernel, one warp per column and stored plane, before this launch)
            if (kprev >= 0 && !fa.mom_external) {
                if (tid < NMS) mom[tid] = fa.momg[(size_t)kprev * NMS + tid];
                fs3_barP();
                column_poly_rows_sub(fa, gI, kprev, mom, kap + (size_t)kprev * fa.nrows, tid, 256, JONLY);
            }
        }
    } else {
        // ====================================== transform warps ======================================
#if FS4_NTW == 8
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FS4_REGT));
#endif
        const int fw = warp - 8;
        const int njobs = nseg * NP;
        H16Tw htw;
        h16_load(htw, fa.tabA, hl);
        const bool wrap1 = (a.N0 >= FS3_M);               // at most one wrap per window (uniform)
        int seenL = 0;                                    // window-ring phases (global segment numbers) this warp has seen land
        for (int k1 = blockIdx.x; k1 < a.NH; k1 += gridDim.x, g += nseg) {
            cd* kaprow = kap + (size_t)k1 * fa.nrows;
            // this warp's half h works on job id0 + h; the halves run in lockstep, so the warp waits for the windows and the
            // ring slots of both jobs (they belong to the same or to consecutive segments)
#ifdef FS4_DEBUG
            long long fWaitL = 0, fWaitC = 0, fWork = 0, fT0 = clock64(); int fJobs = 0;
#endif
            for (int id0 = 2 * fw; id0 < njobs; id0 += FS4_NWK) {
                const bool active = id0 + half < njobs;
#ifdef FS4_DEBUG
                long long f0 = clock64();
#endif
                const int id = active ? id0 + half : id0;
                const int s = id / NP, p = id - s * NP;
                const int gs = g + s, slot = gs % NSLOT;
                const int gsB = g + min(id0 + 1, njobs - 1) / NP;
                const bool roleA = p < NA, isJ = p == PJ;
                const int pl = roleA ? A0 + p : (isJ ? 0 : A0 + (p - NA));
                const int my_i = isJ ? 0 : a.pl_i[pl];
                const int my_src = isJ ? DK + 1 : a.pl_j[pl];
                const int c0 = s * S, Sc = min(S, a.N0 - c0);
                // every phase of the window ring is observed in order: a parity wait is only unambiguous within one phase, and
                // with few spectra per segment the consecutive jobs of a warp lie more than a ring length apart
                while (seenL <= gsB) { fs4_wait_landed(landed + (unsigned)seenL % (unsigned)NSTG, ((unsigned)seenL / (unsigned)NSTG) & 1u, cons, (unsigned)seenL); ++seenL; }
#ifdef FS4_DEBUG
                long long f1 = clock64(); fWaitL += f1 - f0;
#endif
                if (gsB >= NSLOT) fs3_wait_consumed(cons, (unsigned)(gsB - NSLOT + 1));    // segment gs - NSLOT (same slot) consumed by every product warp
#ifdef FS4_DEBUG
                long long f2 = clock64(); fWaitC += f2 - f1;
#endif
                const TSt* src = stage + ((size_t)((unsigned)gs % (unsigned)NSTG) * NSRC + my_src) * FS3_M;
                cd* plane = spec + ((size_t)slot * NP + p) * FS4_PITCH;
                cd v[16];
                // Branch-free load (the two halves of the warp differ in role and power, and a divergent branch per element
                // costs an instruction fetch each): v = g * s with s = keep * cx^i, cx^i by Horner on the indicator of i.
                // cx of window position n = hl + 16 q: one int->double conversion per job, the rest by FMA; rows that wrapped
                // around the column ends (first / last segments only) are shifted by one period.
                const int row_l = c0 - h + hl;
                const double cx_l = (double)(row_l + 1) * inv0;
                const double e0 = my_i == 0 ? 1.0 : 0.0, e1 = my_i == 1 ? 1.0 : 0.0, e2 = my_i == 2 ? 1.0 : 0.0, e3 = my_i == 3 ? 1.0 : 0.0;
                const int klo = (roleA ? h : 0) - hl, khi = (active ? (roleA ? h + Sc : FS3_M) : 0) - hl;   // keep <=> klo <= 16 q < khi
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const cd gg = load_c(src + hl + 16 * q);
                    const int row = row_l + 16 * q;
                    double cx = fma((double)(16 * q), inv0, cx_l);
                    if (wrap1) cx += (row < 0) ? 1.0 : ((row >= a.N0) ? -1.0 : 0.0);
                    else cx -= floor(fma(-0.5, inv0, cx));       // columns shorter than the window: any number of wraps, no integer division
                                                                 // ((row + 0.5) / N0 is never within 0.5 / N0 of an integer)
                    double sc = fma(cx, fma(cx, fma(cx, e3, e2), e1), e0);
                    sc = (16 * q >= klo && 16 * q < khi) ? sc : 0.0;
                    v[q] = cmake(gg.x * sc, gg.y * sc);
                }
                hfft256(v, plane, hl, htw, -1.0, active);
                if (active) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) plane[HPAD(hl + 16 * q)] = v[q];
                }
                __syncwarp();
                if (hl == 0 && active) fs3_mbar_arrive(full + slot);
#ifdef FS4_DEBUG
                fWork += clock64() - f2; ++fJobs;
#endif
            }
#ifdef FS4_DEBUG
            if (blockIdx.x == 0 && k1 == blockIdx.x + gridDim.x && lane == 0)
                printf("F warp %d: loop %lld cycles, %d job pairs, wait_landed %lld, wait_consumed %lld, work %lld\n", fw, clock64() - fT0, fJobs, fWaitL, fWaitC, fWork);
            long long fI0 = clock64();
#endif
            fs3_bar0();                                    // (A)
            fs3_bar0();
#if FS4_INV_ALL
            {
                const int wk = 2 * warp + half;
                const bool act = wk < NACC;
                if (__any_sync(0xffffffffu, act))
                    fs4_inverse_job<NPAIR>(fa, htw, spec + (act ? wk : 0) * FS4_PITCH, job_of(act ? wk : 0), hl, act, kaprow);
            }
#else
            // the inverse transforms run on the 16 half-warp workers of the transform warps only (two rounds for 27 pairs): the
            // product warps carry no copy of the transform code, and the cold code executed once per column is halved
#pragma unroll 1
            for (int jb = 2 * fw + half; jb - half < NACC; jb += FS4_NWK) {
                const bool act = jb < NACC;
                fs4_inverse_job<NPAIR>(fa, htw, spec + (act ? jb : 0) * FS4_PITCH, job_of(act ? jb : 0), hl, act, kaprow);
            }
#endif
            fs3_bar0();
#ifdef FS4_DEBUG
            if (blockIdx.x == 0 && k1 == blockIdx.x + gridDim.x && lane == 0 && fw == 0) printf("F warp 0: inverse phase %lld cycles\n", clock64() - fI0);
#endif
        }
    }
}

// kernels_gen.cuh -- the table-driven ("general basis") column kernels: B-spline or polynomial spatial variation of any
// degree for the kernel, the photometric scaling and the background (sfft/BSplineSFFT.py).
//
// Reference: every basis image is a tensor product U_i(r) V_j(c) of 1-D functions (B-splines from
// Create_BSplineBasis, BSplineSFFT.py:2624-2634, or monomials; KerSpatial / ScaSpatial / BkgSpatial :276-458); the
// reference materialises Fij + ScaFij + Fpq full planes, FFTs them, forms Fij^2 + ... cross-spectrum planes and FFTs
// those (:463-1222, 1348-2004).  Here, exactly like the polynomial path (kernels_fit_seg3.cuh):
//   * the row pass stores one transposed row-spectrum plane per distinct V function (V_j multiplied into the load);
//   * a "column plane" A = (U table index, stored plane) is formed on the fly, G_A[r; k1] = U_A(r) g_vs(A)[r; k1];
//   * the lag rows kappa_AB[m0; k1] = sum_r conj(G_A[r]) G_B[(r + m0) % N0] come from segmented overlap-save
//     correlation with 256-point FFTs, cross spectra accumulated over segments in registers.
// The pair set is too large for one launch (Fij = 25: 351 pairs), so a launch ("pass") covers a block of up to
// GEN_NA x GEN_NB pairs: GEN_NA planes in the A role against GEN_NB entries of the B universe
//   B universe = column planes | J | background row functions P_p
// The background planes T_pq = P_p(r) Q_q(c) are never transformed: the B-role spectrum of the real sequence P_p is the
// same for every column and comes from a table computed at plan creation; the factor DFT(Q_q)[k1] multiplies the lags.
#pragma once
#include "kernels_fit_seg4.cuh"

#define GEN_NA 5
#define GEN_NB 5
#define GEN_MAXSRC 4
#define GEN_MAXQ 12          // background functions sharing one row function P_p
#define GEN_MAXP 12          // distinct background row functions

struct GenPass {
    int na, nb, nbt;                 // A slots, B slots, B slots that are transformed (types 0 / 1; they come first)
    int nsrc;                        // stored planes staged per segment
    int src_plane[GEN_MAXSRC];       // index of the stored plane (row spectra of I x V_vs), -1 = the plane of J
    short a_u[GEN_NA], a_src[GEN_NA];            // U table row, staged-source slot
    short b_type[GEN_NB];            // 0 = column plane, 1 = J, 2 = background row function
    short b_u[GEN_NB], b_src[GEN_NB];            // type 0: U table row and source slot; 1: source slot; 2: b_u = p
    int rowbase[GEN_NA * GEN_NB];    // first kap row of pair (a, b), -1 = pair not needed
    int ninv;                        // number of needed pairs
    unsigned char inv_q[GEN_NA * GEN_NB];        // their accumulator indices a * GEN_NB + b, ascending
    // Local support of the 1-D functions (B-splines vanish outside a few knot intervals; BSplineSFFT.py:2624-2634): a window
    // whose U factor is identically zero has a zero spectrum, so its transform and every product with it are skipped.
    //   jobs[]    : the transforms that are really run, (segment << 4) | spectrum index p (A roles, then transformed B roles),
    //               ascending; every segment keeps at least one (the ring protocol walks through all segments);
    //   seginfo[] : per segment, bits 0..4 = A slots with a non-zero window, bits 5..9 = B slots to multiply with (transformed
    //               ones with a non-zero window, background row functions always), bits 16.. = number of jobs of the segment.
    const unsigned short* jobs;
    const unsigned* seginfo;
    int njobs;
};

struct GenFitArgs {
    int N0, NH, w0;
    int S, nseg, h;                  // core rows per segment, number of segments, halo = 2 w0
    int nrows;                       // rows per column of kap
    const double* U;                 // [nU][N0] 1-D functions along axis 0 (kernel | scaling | sum)
    const cd* TF;                    // [nseg][Fp][256] B-role spectra of the background row functions
    int Fp;
    const cd* Q;                     // [Fq][NH] DFT along axis 1 of the background column functions
    int tq_n[GEN_MAXP];              // background functions f = (p, q) per p ...
    unsigned char tq_q[GEN_MAXP][GEN_MAXQ];      // ... their q, in the order their rows are laid out
    size_t plane_stride;             // elements between stored planes
};

// ---- one pass of the general fit column kernel on the half-warp FFT engine (fft_h16.cuh), structure of fit_seg4_kernel ----------------------------------------
// 8 product warps (thread = frequency bin, GEN_NA x GEN_NB accumulators) + 8 transform warps = 16 half-warp workers; one
// shared-memory exchange per 256-point transform instead of two; the window load is branch-free (the halves of a warp differ in
// role and table); every needed accumulator gets a plane, so a column ends with ONE inverse batch on the 32 half warps.
// spectrum ring slots (the transforms of segment s + 2, s + 3 start while the products of s still run) and window ring depth (any
// number, not only powers of two).  fp32-stored spectra: GEN_NSLOT32 slots, GEN_NSTG32 windows of 4 sources; fp64: 3 slots, 4 windows.
#ifndef GEN_NSLOT32
#define GEN_NSLOT32 3
#endif
#ifndef GEN_NSTG32
#define GEN_NSTG32 8
#endif
template <typename TSt> struct GenRing {
    static const int depth = sizeof(TSt) == 8 ? GEN_NSTG32 : 4;
    static const int nslot = sizeof(TSt) == 8 ? GEN_NSLOT32 : 3;
    static const int nplanes = nslot * (GEN_NA + GEN_NB) > 26 ? nslot * (GEN_NA + GEN_NB) : 26;     // ring planes, and >= GEN_NA GEN_NB inverse planes
};
static_assert(26 >= GEN_NA * GEN_NB, "planes of the general fit kernel");
static inline size_t gen_fit4_smem_bytes(bool f32) {
    const int np = f32 ? GenRing<float2>::nplanes : GenRing<double2>::nplanes;
    return sizeof(cd) * (size_t)np * FS3_PITCH + 128 + (f32 ? sizeof(float2) * GEN_NSTG32 : sizeof(double2) * 4) * (size_t)GEN_MAXSRC * FS3_M;
}

__device__ __forceinline__ void gen4_inverse_job(const GenFitArgs& fa, const GenPass& ps, const H16Tw& tw, cd* plane, int jb, int hl, bool active,
                                                 int k1, cd* __restrict__ kaprow)
{
    cd v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = active ? plane[HPAD(hl + 16 * q)] : cmake(0.0, 0.0);
    __syncwarp();
    hfft256(v, plane, hl, tw, +1.0, active);
    if (!active) return;
    const int qq = ps.inv_q[jb];
    const int b = qq % GEN_NB;
    const int bt = ps.b_type[b];
    const int lim = bt == 0 ? 2 * fa.w0 : fa.w0;
    const int rowb = ps.rowbase[qq];
    const double invM = 1.0 / (double)FS3_M;
    const int p = ps.b_u[b], nlj0 = 2 * fa.w0 + 1;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int idx = hl + 16 * q;
        const int m0 = idx < FS3_M / 2 ? idx : idx - FS3_M;
        if (m0 >= -lim && m0 <= lim) {
            const cd lam = cscale(v[q], invM);
            if (bt != 2) kaprow[rowb + m0 + lim] = lam;
            else
                for (int t = 0; t < fa.tq_n[p]; ++t)
                    kaprow[rowb + t * nlj0 + m0 + lim] = cmul(lam, fa.Q[(size_t)fa.tq_q[p][t] * fa.NH + k1]);
        }
    }
}

template <typename TSt>
__global__ void __launch_bounds__(FS4_NT, 1) fit_gen4_kernel(GenFitArgs fa, GenPass ps, const cd* __restrict__ tabA, const TSt* __restrict__ gP,
                                                             const TSt* __restrict__ gJ, cd* __restrict__ kap)
{
    constexpr int NA = GEN_NA, NB = GEN_NB, NACC = NA * NB;
    constexpr int NPMAX = NA + NB;
    constexpr int NSTG = GenRing<TSt>::depth, PFD = NSTG - 2, GEN_NSLOT = GenRing<TSt>::nslot, GEN_NPLANES = GenRing<TSt>::nplanes;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* spec = reinterpret_cast<cd*>(smem_raw);                                   // GEN_NPLANES planes
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(spec + GEN_NPLANES * FS3_PITCH);
    unsigned* cons = reinterpret_cast<unsigned*>(bars + 12);      // [8] segments consumed by product warp w (fs3_publish)
    TSt* stage = reinterpret_cast<TSt*>(bars + 16);
    unsigned long long* full = bars;          // [GEN_NSLOT]  count NP   (one arrive per transform job, skipped ones included)
    unsigned long long* landed = bars + 4;    // [NSTG] count 256  (cp.async arrivals of the product threads)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, half = lane >> 4, hl = lane & 15;
    const int h = fa.h, S = fa.S, nseg = fa.nseg, N0 = fa.N0;
    const int NP = ps.na + ps.nbt;            // spectra per segment: A roles | transformed B roles
    const int nsrc = ps.nsrc;

    if (tid == 0) {
        for (int b = 0; b < GEN_NSLOT; ++b) fs3_mbar_init(full + b, NP);
        for (int b = 0; b < NSTG; ++b) fs3_mbar_init(landed + b, 256);
    }
    if (tid < 8) cons[tid] = 0u;
    for (int i = tid; i < GEN_NPLANES * FS3_PITCH; i += FS4_NT) spec[i] = cmake(0.0, 0.0);   // unused slots must stay finite
    __syncthreads();
    int g = 0;                                // global segment counter (ring phases continue across columns)

    if (warp < 8) {
        // ======================================= product warps =======================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FS4_REGP));
        for (int k1 = blockIdx.x; k1 < fa.NH; k1 += gridDim.x, g += nseg) {
            cd acc[NACC];
#pragma unroll
            for (int q = 0; q < NACC; ++q) acc[q] = cmake(0.0, 0.0);
            cd* kaprow = kap + (size_t)k1 * fa.nrows;
            auto issue = [&](int s) {
                const int buf = (int)((unsigned)(g + s) % (unsigned)NSTG);
                const int r = wrap_row(s * S - h + tid, N0);
                for (int jj = 0; jj < nsrc; ++jj) {
                    const int pl = ps.src_plane[jj];
                    const TSt* col = (pl < 0) ? gJ + (size_t)k1 * N0 : gP + (size_t)pl * fa.plane_stride + (size_t)k1 * N0;
                    cp_async_elem(stage + ((size_t)buf * GEN_MAXSRC + jj) * FS3_M + tid, col + r);
                }
                fs3_cp_async_arrive(landed + buf);
            };
            for (int s = 0; s < PFD && s < nseg; ++s) issue(s);
            for (int s = 0; s < nseg; ++s) {
                const int gs = g + s, gq = gs / GEN_NSLOT, slot = gs - gq * GEN_NSLOT;
                // background B slots: their spectrum is a table, independent of the column (loaded before the wait)
                const unsigned sm = __ldg(ps.seginfo + s);                  // non-zero A slots | B slots (uniform over the CTA)
                cd fB[NB];
#pragma unroll
                for (int b = 0; b < NB; ++b)
                    fB[b] = (b < ps.nb && ps.b_type[b] == 2) ? fa.TF[((size_t)s * fa.Fp + ps.b_u[b]) * FS3_M + tid] : cmake(0.0, 0.0);
                fs3_mbar_wait(full + slot, gq & 1);
                {
                    const cd* sp = spec + (size_t)slot * NPMAX * FS3_PITCH + HPAD(tid);
                    // slots that are unused or skipped hold stale spectra / inverse-phase scratch and are never read: a skipped B
                    // slot enters the products as an exact zero, a skipped A slot skips its row of products
#pragma unroll
                    for (int b = 0; b < NB; ++b)
                        if (b < ps.nbt) fB[b] = ((sm >> (5 + b)) & 1u) ? sp[(NA + b) * FS3_PITCH] : cmake(0.0, 0.0);
#pragma unroll
                    for (int A = 0; A < NA; ++A) {
                        if (!((sm >> A) & 1u)) continue;
                        const cd fA = sp[A * FS3_PITCH];
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            cd& c = acc[A * NB + b];
                            c.x = fma(fA.x, fB[b].x, c.x); c.x = fma(fA.y, fB[b].y, c.x);
                            c.y = fma(fA.x, fB[b].y, c.y); c.y = fma(-fA.y, fB[b].x, c.y);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) fs3_publish(cons + warp, (unsigned)(gs + 1));
                if (s + PFD < nseg) issue(s + PFD);
            }
            fs3_bar0();                                    // (A) all transforms and products of the column are done
            {
                int pos = 0;
#pragma unroll
                for (int q = 0; q < NACC; ++q) {
                    if (ps.rowbase[q] >= 0) { spec[pos * FS3_PITCH + HPAD(tid)] = acc[q]; ++pos; }
                }
            }
            fs3_bar0();
            {
                H16Tw htw;
                h16_load(htw, tabA, hl);
                const int wk = 2 * warp + half;
                const bool act = wk < ps.ninv;
                if (__any_sync(0xffffffffu, act)) gen4_inverse_job(fa, ps, htw, spec + (act ? wk : 0) * FS3_PITCH, act ? wk : 0, hl, act, k1, kaprow);
            }
            fs3_bar0();
        }
    } else {
        // ====================================== transform warps ======================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FS4_REGT));
        const int fw = warp - 8;
        const int njobs = ps.njobs;               // the transforms with a non-zero window
        H16Tw htw;
        h16_load(htw, tabA, hl);
        int seenL = 0;
        for (int k1 = blockIdx.x; k1 < fa.NH; k1 += gridDim.x, g += nseg) {
            cd* kaprow = kap + (size_t)k1 * fa.nrows;
            for (int id0 = 2 * fw; id0 < njobs; id0 += FS4_NWK) {
                const bool active = id0 + half < njobs;
                const int id = active ? id0 + half : id0;
                const unsigned jb = __ldg(ps.jobs + id);
                const int s = (int)(jb >> 4), p = (int)(jb & 15u);
                const int gs = g + s, slot = gs % GEN_NSLOT;
                const int gsB = g + (int)(__ldg(ps.jobs + min(id0 + 1, njobs - 1)) >> 4);
                // the first job of a segment also stands in for the skipped ones on the slot's barrier
                const bool first = id == 0 || (int)(__ldg(ps.jobs + id - 1) >> 4) != s;
                const unsigned narrive = 1u + (first ? (unsigned)NP - (__ldg(ps.seginfo + s) >> 16) : 0u);
                const bool roleA = p < ps.na;
                const int bs = p - ps.na;
                const bool isJ = !roleA && ps.b_type[bs] == 1;
                const int my_u = roleA ? ps.a_u[p] : ps.b_u[bs];
                const int my_src = roleA ? ps.a_src[p] : ps.b_src[bs];
                const int c0 = s * S, Sc = min(S, N0 - c0);
                // the U factors of the window do not depend on the staged data: request them before the waits, so that their
                // latency (L2: the tables of a pass do not fit the small L1 left beside 185 KB of shared memory) is hidden
                const double* urow = fa.U + (size_t)(isJ ? 0 : my_u) * N0;
                const int klo = (roleA ? h : 0) - hl, khi = (active ? (roleA ? h + Sc : FS3_M) : 0) - hl;   // keep <=> klo <= 16 q < khi
                double uu[16];
                {
                    int r = wrap_row(c0 - h + hl, N0);
                    const int step = 16 % N0;
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        uu[q] = (16 * q >= klo && 16 * q < khi) ? (isJ ? 1.0 : __ldg(urow + r)) : 0.0;
                        r += step;
                        if (r >= N0) r -= N0;
                    }
                }
                while (seenL <= gsB) { fs4_wait_landed(landed + (unsigned)seenL % (unsigned)NSTG, ((unsigned)seenL / (unsigned)NSTG) & 1u, cons, (unsigned)seenL); ++seenL; }
                if (gsB >= GEN_NSLOT) fs3_wait_consumed(cons, (unsigned)(gsB - (GEN_NSLOT - 1)));
                const TSt* src = stage + ((size_t)((unsigned)gs % (unsigned)NSTG) * GEN_MAXSRC + my_src) * FS3_M;
                cd* plane = spec + ((size_t)slot * NPMAX + (roleA ? p : NA + bs)) * FS3_PITCH;
                cd v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const cd gg = load_c(src + hl + 16 * q);
                    v[q] = cmake(gg.x * uu[q], gg.y * uu[q]);
                }
                hfft256(v, plane, hl, htw, -1.0, active);
                if (active) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) plane[HPAD(hl + 16 * q)] = v[q];
                }
                __syncwarp();
                if (hl == 0 && active)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(fs3_saddr(full + slot)), "r"(narrive) : "memory");
            }
            fs3_bar0();                                    // (A)
            fs3_bar0();
            {
                const int wk = 2 * warp + half;
                const bool act = wk < ps.ninv;
                if (__any_sync(0xffffffffu, act)) gen4_inverse_job(fa, ps, htw, spec + (act ? wk : 0) * FS3_PITCH, act ? wk : 0, hl, act, k1, kaprow);
            }
            fs3_bar0();
        }
    }
}

// ---- lag tables: sum of the k1 chunks of lag_reduce2_kernel, every row keeps all 4 w1 + 1 axis-1 lags ------------------
#ifdef SFFTB_TU_GEN
__global__ void gen_lag_finish_kernel(int nrows, int nl1, int ksplit, const double* __restrict__ part, double* __restrict__ Rall)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)nrows * nl1) return;
    double s = 0.0;
    for (int ks = 0; ks < ksplit; ++ks) s += part[(size_t)ks * nrows * nl1 + idx];
    Rall[idx] = s;
}
#endif

// ---- J x T terms in real space: RJT[f] = sum_{r,c} J[r,c] P_fu(r) Q_fv(c) ------------------------------------------------
struct GenBkg {
    int N0, N1, Fp, Fq, Fpq;
    const double* P;                 // [Fp][N0]
    const double* Qr;                // [Fq][N1] (real-space column functions)
    const int* fu; const int* fv;    // [Fpq]
};

// grid: CTAs stride over rows; smem: Fq doubles per warp partial + Fp*Fq accumulators.  Every CTA writes its partial sums to
// PQpart[blockIdx.x][Fp * Fq]; gen_rjt_finish_kernel adds them in CTA order (no atomics: the right-hand side -- and with it Solution
// and DIFF -- repeats bit for bit from run to run, like the reference's).
template <typename TIn>
__global__ void __launch_bounds__(256) gen_rjt_kernel(GenBkg a, const TIn* __restrict__ J, double* __restrict__ PQpart)
{
    __shared__ double sq[8][GEN_MAXQ];
    __shared__ double accs[GEN_MAXP * GEN_MAXQ];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < a.Fp * a.Fq; i += 256) accs[i] = 0.0;
    __syncthreads();
    for (int r = blockIdx.x; r < a.N0; r += gridDim.x) {
        double s[GEN_MAXQ];
#pragma unroll
        for (int q = 0; q < GEN_MAXQ; ++q) s[q] = 0.0;
        const TIn* row = J + (size_t)r * a.N1;
        for (int c = tid; c < a.N1; c += 256) {
            const double v = (double)row[c];
#pragma unroll
            for (int q = 0; q < GEN_MAXQ; ++q)
                if (q < a.Fq) s[q] = fma(v, a.Qr[(size_t)q * a.N1 + c], s[q]);
        }
#pragma unroll
        for (int q = 0; q < GEN_MAXQ; ++q) {
            if (q < a.Fq) {
                const double t = warp_sum(s[q]);
                if (lane == 0) sq[warp][q] = t;
            }
        }
        __syncthreads();
        if (tid < a.Fp * a.Fq) {
            const int p = tid / a.Fq, q = tid - p * a.Fq;
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += sq[w][q];
            accs[tid] = fma(t, a.P[(size_t)p * a.N0 + r], accs[tid]);
        }
        __syncthreads();
    }
    if (tid < a.Fp * a.Fq) PQpart[(size_t)blockIdx.x * (a.Fp * a.Fq) + tid] = accs[tid];
}

#ifdef SFFTB_TU_GEN
__global__ void gen_rjt_finish_kernel(int nparts, int n, const double* __restrict__ PQpart, double* __restrict__ PQ)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int b = 0; b < nparts; ++b) s += PQpart[(size_t)b * n + i];
    PQ[i] = s;
}
#endif

// ---- normal-equation fill for arbitrary unknown descriptors ------------------------------------------------------------
// Unknown u of the solved (tweaked) system: a design-matrix column SCALE * (roll(X_plane, (a, b)) - mod * X_plane) with
// X_plane = I * U * V of column plane `plane` (kernel, scaling or the sum plane), or the background function T_f
// (plane = -1 - f).  refs: the indices of the reference's (untweaked) Solution layout this unknown stands for -- one for
// ordinary unknowns, Fij for the summed centre-tap unknown of a B-spline kernel with a constant scaling
// (TweakLS 'sum' variant, BSplineSFFT.py:2202-2272; Restore_Solution :3764-3771).
struct GenFillArgs {
    int n, NEQ, Fijab, Fab, Fij, Fpq, L0, L1, w0, w1, nl1, P;
    const int* u_plane; const signed char* u_a; const signed char* u_b; const signed char* u_mod;
    const int* u_ref0; const int* u_nref; const int* refs;
    const int* pairrow;              // [P][P] first Rall row of the pair with plane A in the A role, -1 = stored as (B, A)
    const int* rowJ;                 // [P]
    const int* rowT;                 // [P][Fpq]
    const double* Rall; const double* PQ; const double* PHI;
    const int* fu; const int* fv; int Fq;
    double invN, invN2, invN3;
    const double* SST; const double* iREG; const double* CSST; const double* DSST; double regw;
};

#ifdef SFFTB_TU_GEN
__device__ __forceinline__ double gen_R(const GenFillArgs& f, int A, int B, int m0, int m1) {
    // the pair was computed with one of its planes in the A role (whichever the pass order put first): R_BA[m] = R_AB[-m]
    int rb = f.pairrow[A * f.P + B];
    if (rb < 0) { rb = f.pairrow[B * f.P + A]; m0 = -m0; m1 = -m1; }
    return f.Rall[((size_t)rb + (m0 + 2 * f.w0)) * f.nl1 + (m1 + 2 * f.w1)];
}

__device__ double gen_reg(const GenFillArgs& f, int u, int v) {
    double s = 0.0;
    const int c0 = f.w0 * f.L1 + f.w1;
    for (int x = 0; x < f.u_nref[u]; ++x) {
        const int k = f.refs[f.u_ref0[u] + x];
        const int K = k / f.Fab, c = k - K * f.Fab;
        for (int y = 0; y < f.u_nref[v]; ++y) {
            const int l = f.refs[f.u_ref0[v] + y];
            const int K2 = l / f.Fab, c2 = l - K2 * f.Fab;
            double g = f.SST[K * f.Fij + K2];
            if (f.CSST) {
                if (c == c0 && c2 == c0) g = f.DSST[K * f.Fij + K2];
                else if (c2 == c0) g = f.CSST[K * f.Fij + K2];      // row tap off-centre, column tap = centre
                else if (c == c0) g = f.CSST[K2 * f.Fij + K];
            }
            s = fma(g, f.iREG[(size_t)c * f.Fab + c2], s);
        }
    }
    return s * f.regw;
}

__device__ double gen_lh_entry(const GenFillArgs& f, int u, int v) {
    const int A = f.u_plane[u], B = f.u_plane[v];
    if (A >= 0 && B >= 0) {
        const int a8 = f.u_a[u], b8 = f.u_b[u], a0 = f.u_a[v], b0 = f.u_b[v];
        const bool nz8 = f.u_mod[u] != 0, nz = f.u_mod[v] != 0;
        double x = gen_R(f, A, B, a8 - a0, b8 - b0);
        if (nz) x -= gen_R(f, A, B, a8, b8);
        if (nz8) x -= gen_R(f, A, B, -a0, -b0);
        if (nz && nz8) x += gen_R(f, A, B, 0, 0);
        x *= f.invN3;
        if (f.SST) x += gen_reg(f, u, v);
        return x;
    }
    if (A < 0 && B < 0) return f.PHI[(-1 - A) * f.Fpq + (-1 - B)] * f.invN;
    const int uk = A >= 0 ? u : v;                       // the kernel-type unknown
    const int pl = A >= 0 ? A : B, fb = A >= 0 ? -1 - B : -1 - A;
    const int rb = f.rowT[pl * f.Fpq + fb];
    double x = f.Rall[((size_t)rb + (f.u_a[uk] + f.w0)) * f.nl1 + (f.u_b[uk] + 2 * f.w1)];
    if (f.u_mod[uk]) x -= f.Rall[((size_t)rb + f.w0) * f.nl1 + 2 * f.w1];
    return x * f.invN2;
}

__device__ double gen_rhs_entry(const GenFillArgs& f, int u) {
    const int A = f.u_plane[u];
    if (A < 0) { const int fb = -1 - A; return f.PQ[f.fu[fb] * f.Fq + f.fv[fb]] * f.invN; }
    const int rb = f.rowJ[A];
    double x = f.Rall[((size_t)rb + (f.u_a[u] + f.w0)) * f.nl1 + (f.u_b[u] + 2 * f.w1)];
    if (f.u_mod[u]) x -= f.Rall[((size_t)rb + f.w0) * f.nl1 + 2 * f.w1];
    return x * f.invN2;
}

__global__ void gen_fill_diag_kernel(GenFillArgs f, double* __restrict__ sc, int* __restrict__ info)
{
    const int rr = blockIdx.x * blockDim.x + threadIdx.x;
    if (rr >= f.n) return;
    const double d = gen_lh_entry(f, rr, rr);
    if (!isfinite(d)) { atomicExch(&info[1], 1); sc[rr] = 1.0; }
    else if (!(d > 0.0)) { atomicExch(&info[0], rr + 1); sc[rr] = 1.0; }
    else sc[rr] = rsqrt(d);
}

// Aug is (n+1) x ld row-major as in fill_matrix_kernel; sc == nullptr -> unscaled (export hook)
__global__ void gen_fill_matrix_kernel(GenFillArgs f, const double* __restrict__ sc, double* __restrict__ Aug, int ld, int* __restrict__ info)
{
    const int cc = blockIdx.x * blockDim.x + threadIdx.x;
    const int rr = blockIdx.y * blockDim.y + threadIdx.y;
    const int n = f.n;
    if (rr > n || cc > n) return;
    double v;
    if (rr < n && cc < n) {
        v = gen_lh_entry(f, rr, cc);
        if (sc) v *= sc[rr] * sc[cc];
    } else if (rr == n && cc == n) {
        v = 0.0;
    } else {
        const int k = rr == n ? cc : rr;
        v = gen_rhs_entry(f, k);
        if (sc) v *= sc[k];
    }
    if (!isfinite(v)) atomicExch(&info[1], 1);
    Aug[(size_t)rr * ld + cc] = v;
}

// Restore_Solution (BSplineSFFT.py:3704-3783): the solver has written unknown u to sol[refs[u_ref0[u]]]; copy it to the
// other indices it stands for (the tied centre taps).  Indices without an unknown stay 0.
__global__ void gen_restore_kernel(GenFillArgs f, double* __restrict__ sol)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= f.n) return;
    const int nr = f.u_nref[u];
    if (nr <= 1) return;
    const double v = sol[f.refs[f.u_ref0[u]]];
    for (int x = 1; x < nr; ++x) sol[f.refs[f.u_ref0[u] + x]] = v;
}
#endif  // SFFTB_TU_GEN

// ---- subtract column pass (Construct_FDIFF, BSplineSFFT.py:2343-2528) as a FIR over the stored planes ------------------
//   d[r; k1] = g_J[r] + sum_A c_A U_A(r) g_A[r] - sum_A sum_a h_A[a; k1] U_A(r - a) g_A[r - a]
// apply planes: the Fij kernel planes (taps from the Solution, centre tap unmodified, c_A = sum of the off-centre taps)
// and, for a varying scaling, the ScaFij scaling planes (centre tap only).  Planes are ordered by stored plane:
// vs_first[j] .. vs_first[j + 1] - 1 are the apply planes read from stored plane vs_id[j].
#define GFIR_NT 256
#define GFIR_R 2
#define GFIR_CH (GFIR_NT * GFIR_R)
#define GEN_MAXVS 24
struct GenFirArgs {
    int N0, N1, NH, w0, w1, L0;
    int nap;                         // apply planes
    int nvs;                         // stored planes they read
    int nU;                          // rows of the staged U window
    short vs_id[GEN_MAXVS], vs_first[GEN_MAXVS + 1];
    const short* ap_u;               // [nap] U table row (index into the staged window)
    const int* ap_sol;               // [nap] kernel planes: first Solution index of the plane; scaling planes: -1 - index of its centre tap
    const int* ap_centre;            // [nap] Solution index of the plane's own centre-tap coefficient, -1 = none (varying scaling)
    const double* U;                 // [nU][N0]
    const cd* tw1;
    size_t plane_stride;
};

#ifdef SFFTB_TU_GEN
// taps[k1][A][ia] = (1/N) sum_b coef(A, ia, b) e^{-2 pi i b k1 / N1};  cA[A] = (1/N) sum_{ab != 00} coef
__global__ void __launch_bounds__(128) gen_taps_kernel(GenFirArgs a, const double* __restrict__ sol, cd* __restrict__ taps, double* __restrict__ cAout)
{
    const int k1 = blockIdx.x, tid = threadIdx.x;
    const int L0 = a.L0, L1 = 2 * a.w1 + 1, Fab = L0 * L1;
    const double invN = 1.0 / ((double)a.N0 * (double)a.N1);
    for (int idx = tid; idx < a.nap * L0; idx += blockDim.x) {
        const int A = idx / L0, ia = idx - A * L0;
        cd acc = cmake(0.0, 0.0);
        const int s0 = a.ap_sol[A];
        if (s0 >= 0) {
            const double* s = sol + (size_t)s0 + (size_t)ia * L1;
            for (int ib = 0; ib < L1; ++ib) {
                const int b = ib - a.w1;
                double co = s[ib];
                if (ia == a.w0 && b == 0) co = a.ap_centre[A] >= 0 ? sol[a.ap_centre[A]] : 0.0;
                const cd w = a.tw1[imod((int)(((long long)b * k1) % a.N1), a.N1)];
                acc.x = fma(co, w.x, acc.x);
                acc.y = fma(co, w.y, acc.y);
            }
        } else if (ia == a.w0) {
            acc.x = sol[-1 - s0];
        }
        taps[((size_t)k1 * a.nap + A) * L0 + ia] = cscale(acc, invN);
    }
    if (k1 == 0) {
        for (int A = tid; A < a.nap; A += blockDim.x) {
            double t = 0.0;
            const int s0 = a.ap_sol[A];
            if (s0 >= 0) {
                const double* s = sol + (size_t)s0;
                for (int ab = 0; ab < Fab; ++ab) if (ab != a.w0 * L1 + a.w1) t += s[ab];
            }
            cAout[A] = t * invN;
        }
    }
}
#endif

// smem: h[nap][L0] cd | cA[nap] double | uw[nU][W] double | st[nvs][W] cd | uflag[nU] | jfirst[nvs + 1] | alist[nap],  W = GFIR_CH + 2 w0
// Local support: a B-spline row function that vanishes on the whole staged window of this chunk contributes nothing; the CTA
// compacts the apply planes with a non-zero window into alist (per stored plane j: entries jfirst[j] .. jfirst[j + 1] - 1) and
// loops over those only.  The GFIR_R output rows of a thread share every tap load.
static inline size_t gen_fir_smem_bytes(int nap, int L0, int nU, int nvs, int w0) {
    const size_t W = GFIR_CH + 2 * (size_t)w0, nuw = (size_t)nU * W;
    return sizeof(cd) * (size_t)nap * L0 + sizeof(double) * (((size_t)nap + 1) & ~(size_t)1) + sizeof(double) * (nuw + (nuw & 1)) +
           sizeof(cd) * (size_t)nvs * W + sizeof(int) * (size_t)(((nU + 3) & ~3) + ((nvs + 1 + 3) & ~3)) + sizeof(int4) * (size_t)nap;
}

template <typename TSt>
__global__ void __launch_bounds__(GFIR_NT) gen_fir_kernel(GenFirArgs a, const TSt* __restrict__ gP, const TSt* gJ,
                                                          const cd* __restrict__ taps, const double* __restrict__ cAin, TSt* outD)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L0 = a.L0, W = GFIR_CH + 2 * a.w0;
    cd* h = reinterpret_cast<cd*>(smem_raw);
    double* cA = reinterpret_cast<double*>(h + (size_t)a.nap * L0);
    double* uw = cA + ((a.nap + 1) & ~1);
    cd* st = reinterpret_cast<cd*>(uw + (size_t)a.nU * W + (((size_t)a.nU * W) & 1));
    int* uflag = reinterpret_cast<int*>(st + (size_t)a.nvs * W);
    int* jfirst = uflag + ((a.nU + 3) & ~3);
    int4* alist = reinterpret_cast<int4*>(jfirst + ((a.nvs + 1 + 3) & ~3));         // {U row offset, tap offset, apply plane, -}
    const int tid = threadIdx.x, k1 = blockIdx.x, rbeg = blockIdx.y * GFIR_CH;

    for (int u = tid; u < a.nU; u += GFIR_NT) uflag[u] = 0;
    __syncthreads();
    for (int idx = tid; idx < W; idx += GFIR_NT) {
        int r = (rbeg - a.w0 + idx) % a.N0; if (r < 0) r += a.N0;
        for (int u = 0; u < a.nU; ++u) {
            const double v = a.U[(size_t)u * a.N0 + r];
            uw[(size_t)u * W + idx] = v;
            if (v != 0.0) uflag[u] = 1;
        }
        for (int j = 0; j < a.nvs; ++j) st[(size_t)j * W + idx] = load_c(gP + (size_t)a.vs_id[j] * a.plane_stride + (size_t)k1 * a.N0 + r);
    }
    for (int idx = tid; idx < a.nap * L0; idx += GFIR_NT) h[idx] = taps[(size_t)k1 * a.nap * L0 + idx];
    for (int idx = tid; idx < a.nap; idx += GFIR_NT) cA[idx] = cAin[idx];
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int j = 0; j < a.nvs; ++j) {
            jfirst[j] = n;
            for (int A = a.vs_first[j]; A < a.vs_first[j + 1]; ++A) {
                const int u = a.ap_u[A];
                if (uflag[u]) alist[n++] = make_int4(u * W, A * L0, A, 0);
            }
        }
        jfirst[a.nvs] = n;
    }
    __syncthreads();

    cd acc[GFIR_R];
    int nc[GFIR_R];                                        // staged index of the output row itself
#pragma unroll
    for (int o = 0; o < GFIR_R; ++o) {
        const int lo = tid + o * GFIR_NT, r = rbeg + lo;
        nc[o] = lo + a.w0;
        acc[o] = r < a.N0 ? load_c(gJ + (size_t)k1 * a.N0 + r) : cmake(0.0, 0.0);
    }
    for (int j = 0; j < a.nvs; ++j) {
        const cd* sj = st + (size_t)j * W;
        const int x0 = jfirst[j], x1 = jfirst[j + 1];
        if (x0 == x1) continue;
        {   // + c_A U_A(r) g[r]
            double t[GFIR_R];
#pragma unroll
            for (int o = 0; o < GFIR_R; ++o) t[o] = 0.0;
            for (int x = x0; x < x1; ++x) {
                const int4 e = alist[x];
                const double c = cA[e.z];
#pragma unroll
                for (int o = 0; o < GFIR_R; ++o) t[o] = fma(c, uw[e.x + nc[o]], t[o]);
            }
#pragma unroll
            for (int o = 0; o < GFIR_R; ++o) {
                const cd g = sj[nc[o]];
                acc[o].x = fma(t[o], g.x, acc[o].x); acc[o].y = fma(t[o], g.y, acc[o].y);
            }
        }
        for (int ia = 0; ia < L0; ++ia) {
            const int sh = ia - a.w0;                      // source row r - a
            cd t[GFIR_R];
#pragma unroll
            for (int o = 0; o < GFIR_R; ++o) t[o] = cmake(0.0, 0.0);
            for (int x = x0; x < x1; ++x) {
                const int4 e = alist[x];
                const cd hh = h[e.y + ia];
#pragma unroll
                for (int o = 0; o < GFIR_R; ++o) {
                    const double uu = uw[e.x + nc[o] - sh];
                    t[o].x = fma(hh.x, uu, t[o].x); t[o].y = fma(hh.y, uu, t[o].y);
                }
            }
#pragma unroll
            for (int o = 0; o < GFIR_R; ++o) {
                const cd g = sj[nc[o] - sh];
                acc[o].x = fma(-t[o].x, g.x, acc[o].x); acc[o].x = fma(t[o].y, g.y, acc[o].x);
                acc[o].y = fma(-t[o].x, g.y, acc[o].y); acc[o].y = fma(-t[o].y, g.x, acc[o].y);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < GFIR_R; ++o) {
        const int r = rbeg + tid + o * GFIR_NT;
        if (r < a.N0) store_c(outD + (size_t)k1 * a.N0 + r, acc[o]);
    }
}

// ---- background subtraction in real space: DIFF[r, c] -= sum_f b_f P_fu(r) Q_fv(c) ---------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256) gen_bkg_subtract_kernel(GenBkg a, const double* __restrict__ bf, TOut* __restrict__ D)
{
    __shared__ double cq[GEN_MAXQ];
    const int r = blockIdx.y, tid = threadIdx.x;
    if (tid < a.Fq) {
        double s = 0.0;
        for (int f = 0; f < a.Fpq; ++f)
            if (a.fv[f] == tid) s = fma(bf[f], a.P[(size_t)a.fu[f] * a.N0 + r], s);
        cq[tid] = s;
    }
    __syncthreads();
    const int c = blockIdx.x * 256 + tid;
    if (c >= a.N1) return;
    double s = 0.0;
    for (int q = 0; q < a.Fq; ++q) s = fma(cq[q], a.Qr[(size_t)q * a.N1 + c], s);
    TOut* d = D + (size_t)r * a.N1 + c;
    *d = (TOut)((double)*d - s);
}

// sfft_b200.cu -- plan management and the C ABI of libsfft_b200.so (see include/sfft_b200.h).
#define SFFTB_TU_MAIN
#include "plan.h"

// ---------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

int sfftb_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// ---------------------------------------------------------------------------------------------------------------
static bool make_fft_desc(int n, FftDesc* fd) {
    static const int radices[] = {8, 4, 2, 3, 5, 7, 11, 13};
    memset(fd, 0, sizeof *fd);
    fd->n = n;
    int m = n;
    for (int R : radices) {
        while (m % R == 0 && m > 1) {
            if (fd->ns >= SFFTB_MAX_STAGES) return false;
            fd->radix[fd->ns++] = R;
            m /= R;
        }
    }
    return m == 1;
}

static bool fft_fits_threads(const FftDesc& fd, int nthr) {
    for (int s = 0; s < fd.ns; ++s) {
        const int R = fd.radix[s];
        const int mb = (16 / R) > 0 ? (16 / R) : 1;
        if (fd.n / R > mb * nthr) return false;
    }
    return true;
}

int upload_twiddles(int n, cd** out) {
    std::vector<cd> h((size_t)n);
    const long double tp = 6.283185307179586476925286766559005768L;
    for (int e = 0; e < n; ++e) {
        const long double ang = tp * (long double)e / (long double)n;
        h[e].x = (double)cosl(ang);
        h[e].y = (double)(-sinl(ang));
    }
    CK(cudaMalloc(out, sizeof(cd) * (size_t)n));
    CK(cudaMemcpy(*out, h.data(), sizeof(cd) * (size_t)n, cudaMemcpyHostToDevice));
    return 0;
}

// engine table: tab[(r-1)*Ns + k] = exp(-2 pi i r k / (Ns R)),  r = 1..R-1, k = 0..Ns-1
int upload_engine_table(int Ns, int R, cd** out) {
    std::vector<cd> h((size_t)(R - 1) * Ns);
    const long double tp = 6.283185307179586476925286766559005768L;
    for (int r = 1; r < R; ++r)
        for (int k = 0; k < Ns; ++k) {
            const long double ang = tp * (long double)r * (long double)k / ((long double)Ns * (long double)R);
            h[(size_t)(r - 1) * Ns + k].x = (double)cosl(ang);
            h[(size_t)(r - 1) * Ns + k].y = (double)(-sinl(ang));
        }
    CK(cudaMalloc(out, sizeof(cd) * h.size()));
    CK(cudaMemcpy(*out, h.data(), sizeof(cd) * h.size(), cudaMemcpyHostToDevice));
    return 0;
}

// Bluestein tables for length H through a power-of-two M >= 2 H - 1: chirp c[n] = exp(-pi i n^2 / H) with n^2 reduced
// mod 2 H in integers, and B = FFT_M(conj(c) wrapped) / M (radix-2 in long double on the host; once per plan).
int upload_bluestein(int H, int M, cd** outC, cd** outB) {
    typedef long double ld;
    const ld pi = 3.141592653589793238462643383279502884L;
    std::vector<cd> c((size_t)H);
    std::vector<ld> br((size_t)M, 0.0L), bi((size_t)M, 0.0L);
    for (int n = 0; n < H; ++n) {
        const long long q = ((long long)n * n) % (2LL * H);
        const ld ang = pi * (ld)q / (ld)H;
        const ld cr = cosl(ang), ci = -sinl(ang);
        c[n].x = (double)cr; c[n].y = (double)ci;
        br[n] = cr; bi[n] = -ci;                                   // conj(c[n])
        if (n > 0) { br[M - n] = cr; bi[M - n] = -ci; }
    }
    // iterative radix-2 decimation-in-time FFT, sign -1
    for (int i = 1, j = 0; i < M; ++i) {
        int bit = M >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(br[i], br[j]); std::swap(bi[i], bi[j]); }
    }
    for (int len = 2; len <= M; len <<= 1) {
        for (int k = 0; k < len / 2; ++k) {
            const ld ang = 2.0L * pi * (ld)k / (ld)len;
            const ld wr = cosl(ang), wi = -sinl(ang);
            for (int i = k; i < M; i += len) {
                const int j2 = i + len / 2;
                const ld xr = br[j2] * wr - bi[j2] * wi, xi = br[j2] * wi + bi[j2] * wr;
                br[j2] = br[i] - xr; bi[j2] = bi[i] - xi;
                br[i] += xr; bi[i] += xi;
            }
        }
    }
    std::vector<cd> B((size_t)M);
    for (int m = 0; m < M; ++m) { B[m].x = (double)(br[m] / (ld)M); B[m].y = (double)(bi[m] / (ld)M); }
    CK(cudaMalloc(outC, sizeof(cd) * (size_t)H));
    CK(cudaMemcpy(*outC, c.data(), sizeof(cd) * (size_t)H, cudaMemcpyHostToDevice));
    CK(cudaMalloc(outB, sizeof(cd) * (size_t)M));
    CK(cudaMemcpy(*outB, B.data(), sizeof(cd) * (size_t)M, cudaMemcpyHostToDevice));
    return 0;
}

// Q[q][k1] = sum_c cy(c)^q exp(-2 pi i k1 c / N1): one warp per output
__global__ void qtable_kernel(int N1, int NH, int nq, const cd* __restrict__ tw1, cd* __restrict__ Q) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= nq * NH) return;
    const int q = gw / NH, k1 = gw - q * NH;
    double sx = 0.0, sy = 0.0;
    for (int c = lane; c < N1; c += 32) {
        const cd w = tw1[(int)(((long long)k1 * c) % N1)];
        const double v = ipow((c + 1) / (double)N1, q);
        sx = fma(v, w.x, sx);
        sy = fma(v, w.y, sy);
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    if (lane == 0) Q[(size_t)q * NH + k1] = cmake(sx, sy);
}

__global__ void __launch_bounds__(512) dbg_fft_kernel(FftDesc fd, const cd* __restrict__ tw, const cd* __restrict__ in, cd* __restrict__ out,
                               int nbatch, int ppc, int pitch, double sgn) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* buf = reinterpret_cast<cd*>(smem_raw);
    const int b0 = blockIdx.x * ppc;
    const int np = min(ppc, nbatch - b0);
    for (int idx = threadIdx.x; idx < np * fd.n; idx += blockDim.x) {
        const int p = idx / fd.n, e = idx - p * fd.n;
        buf[(size_t)p * pitch + e] = in[(size_t)(b0 + p) * fd.n + e];
    }
    __syncthreads();
    fft_planes(buf, pitch, np, fd, tw, sgn);
    for (int idx = threadIdx.x; idx < np * fd.n; idx += blockDim.x) {
        const int p = idx / fd.n, e = idx - p * fd.n;
        out[(size_t)(b0 + p) * fd.n + e] = buf[(size_t)p * pitch + e];
    }
}

// ---------------------------------------------------------------------------------------------------------------
static void fill_col_common(ColArgs& c, const sfftb_plan* p) {
    const sfftb_dims& d = p->d;
    memset(&c, 0, sizeof c);
    c.N0 = d.N0; c.N1 = d.N1; c.NH = d.N1 / 2 + 1;
    c.DK = d.DK; c.DB = d.DB; c.Fij = d.Fij; c.Fpq = d.Fpq; c.nj = d.DK + 1;
    c.w0 = d.w0; c.w1 = d.w1;
    c.npairs = d.Fij * (d.Fij + 1) / 2;
    c.nl0 = 4 * d.w0 + 1; c.nlj0 = 2 * d.w0 + 1;
    memset(c.plane_of, 0xff, sizeof c.plane_of);
    int A = 0;
    for (int i = 0; i <= d.DK; ++i)
        for (int j = 0; j <= d.DK - i; ++j) { c.pl_i[A] = (unsigned char)i; c.pl_j[A] = (unsigned char)j; c.plane_of[i][j] = (unsigned char)A; ++A; }
    int q = 0;
    for (int a = 0; a < d.Fij; ++a)
        for (int b = a; b < d.Fij; ++b) { c.pairA[q] = (unsigned char)a; c.pairB[q] = (unsigned char)b; ++q; }
    c.tw0 = p->tw0;
}

static size_t fit_smem_bytes(const ColArgs& c, int PB) {
    const size_t nacc = (size_t)c.npairs * c.nl0 + (size_t)c.Fij * c.nlj0;
    return sizeof(cd) * ((size_t)(c.Fij + 1 + PB) * c.pitch + nacc + (size_t)(c.nj + 1) * SFFTB_MAXE + 16 * SFFTB_MAXE);
}

static int plan_free(sfftb_plan* p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    gen_free(p);
    void* ptrs[] = {p->vt8_8, p->vt64_8, p->vt64_4, p->vt256_4, p->vt512_4, p->tabA, p->tabB_row, p->tabC_row, p->tabC_row32, p->tw0, p->tw1, p->twMf, p->twH, p->Q, p->PHI, p->idxmap, p->ident, p->gI, p->gJ, p->stA, p->stB,
                    p->kap, p->lam, p->nuJ, p->kap2, p->momg, p->part, p->R, p->RJ, p->RT, p->RJT, p->Aug, p->sc, p->diagU, p->sol, p->exportbuf, p->info, p->cholW, p->cholY, p->cholX, p->cholBar, p->substFlags, p->substMsg, p->solEff, p->regC, p->regD, p->regSST, p->regI, p->bluTw, p->bluC, p->bluB, p->firTaps, p->firCA, p->tstate, p->stC, p->stD, p->deltaIdx, p->deltaVal, p->aspec, p->bluC16, p->bluB16, p->bluBp16, p->bluTwP16};
    for (void* q : ptrs) if (q) cudaFree(q);
    if (p->gIa && p->gIa != p->gI) cudaFree(p->gIa);
    if (p->gJa && p->gJa != p->gJ) cudaFree(p->gJa);
    if (p->info_h) cudaFreeHost(p->info_h);
    for (int k = 0; k < EV_COUNT; ++k) if (p->ev[k]) cudaEventDestroy(p->ev[k]);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    if (p->stream2) cudaStreamDestroy(p->stream2);
    if (p->evFork) cudaEventDestroy(p->evFork);
    for (int k = 0; k < 4; ++k) if (p->evCopy[k]) cudaEventDestroy(p->evCopy[k]);
    if (p->evStart) cudaEventDestroy(p->evStart);
    if (p->evDone) cudaEventDestroy(p->evDone);
    if (p->evJoin) cudaEventDestroy(p->evJoin);
    delete p;
    return 0;
}

extern "C" int sfftb_version(void) { return SFFTB_VERSION; }
extern "C" const char* sfftb_last_error(void) { return g_err.c_str(); }

// Streams, events, twiddles, the row-pass geometry and the image-sized buffers every kind of plan needs.
int plan_init_common(sfftb_plan* p, const sfftb_config* cfg) {
    p->cfg = *cfg;
    p->device = cfg->device;
    CK(cudaSetDevice(p->device));
    int v = 0;
    CK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, p->device)); p->nsm = v;
    CK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device)); p->max_smem = (size_t)v;
    CK(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
    p->stream = p->own_stream;
    CK(cudaStreamCreateWithFlags(&p->stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&p->evFork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p->evJoin, cudaEventDisableTiming));
    for (int k = 0; k < 4; ++k) CK(cudaEventCreateWithFlags(&p->evCopy[k], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p->evStart, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p->evDone, cudaEventDisableTiming));
    p->overlap = env_int("SFFTB_OVERLAP", 1);
    p->solver_sms = std::max(0, std::min(p->nsm - 1, env_int("SFFTB_SOLVER_SMS", 0)));
    for (int k = 0; k < EV_COUNT; ++k) CK(cudaEventCreate(&p->ev[k]));
    if (init_generic_radix_tables()) return SFFTB_ECUDA;

    sfftb_dims& d = p->d;
    d.N0 = cfg->N0; d.N1 = cfg->N1; d.w0 = cfg->w0; d.w1 = cfg->w1; d.DK = cfg->DK; d.DB = cfg->DB;
    d.L0 = 2 * d.w0 + 1; d.L1 = 2 * d.w1 + 1; d.Fab = d.L0 * d.L1;
    const int N0 = d.N0, N1 = d.N1, NH = N1 / 2 + 1;
    const size_t csz = cfg->storage == SFFTB_STORE_F32 ? sizeof(float2) : sizeof(double2);
    // ---- twiddle tables ----
    if (upload_twiddles(N0, &p->tw0)) return SFFTB_ECUDA;
    if (upload_twiddles(N1, &p->tw1)) return SFFTB_ECUDA;

    // ---- row pass geometry ----
    RowArgs& r = p->row;
    memset(&r, 0, sizeof r);
    r.N0 = N0; r.N1 = N1; r.NH = NH;
    r.packed = (N1 % 2 == 0) ? 1 : 0;
    r.H = r.packed ? N1 / 2 : N1;
    r.pitch = r.H + 1;
    if (!make_fft_desc(r.H, &r.fd) || !fft_fits_threads(r.fd, 512) || env_int("SFFTB_ROW_BLUESTEIN", 0)) {
        // a prime factor > 13 (e.g. a trimmed 4094-pixel DECam row: 2047 = 23 * 89): chirp-z through a power of two
        int M = 1;
        while (M < 2 * r.H - 1) M <<= 1;
        if (!make_fft_desc(M, &r.blu_fd) || !fft_fits_threads(r.blu_fd, 512) || (size_t)(M + 1) * sizeof(cd) > p->max_smem)
            return fail(SFFTB_EINVAL, "unsupported image width N1=%d: row transform length %d has a prime factor > 13 and its "
                                      "chirp-z length %d does not fit shared memory", N1, r.H, M);
        r.blu_M = M;
        r.pitch = M + 1;
        if (upload_twiddles(M, &p->bluTw) || upload_bluestein(r.H, M, &p->bluC, &p->bluB)) return SFFTB_ECUDA;
        r.blu_tw = p->bluTw; r.blu_c = p->bluC; r.blu_B = p->bluB;
        memset(&r.fd, 0, sizeof r.fd);
    }
    int RB = env_int("SFFTB_RB", 8);
    while (RB > 1 && ((size_t)RB * r.pitch * sizeof(cd) > p->max_smem || RB > N0)) RB >>= 1;
    if ((size_t)RB * r.pitch * sizeof(cd) > p->max_smem)
        return fail(SFFTB_EINVAL, "image width N1=%d too large for the shared-memory row transform", N1);
    r.RB = RB;
    p->smem_row = (size_t)RB * r.pitch * sizeof(cd);
    if (!r.blu_M && upload_twiddles(r.H, &p->twH)) return SFFTB_ECUDA;
    r.twH = p->twH; r.tw1 = p->tw1;
    p->rinv.r = r;
    p->rinv.scale = (r.packed ? 2.0 : 1.0) / (double)N1;      // the FIR column pass already carries 1/N0
    CK(cudaMalloc(&p->gJ, csz * (size_t)NH * N0));
    if (cfg->storage == SFFTB_STORE_F32) CK(cudaMalloc(&p->gJa, sizeof(double2) * (size_t)NH * N0));
    else p->gJa = p->gJ;
    CK(cudaMalloc(&p->stA, sizeof(double) * (size_t)N0 * N1));
    CK(cudaMalloc(&p->stB, sizeof(double) * (size_t)N0 * N1));
    CK(cudaMalloc(&p->info, sizeof(int) * 4));
    CK(cudaMallocHost(&p->info_h, sizeof(int) * 4));
    memset(p->info_h, 0, sizeof(int) * 4);
    return 0;
}

static int plan_create_impl(sfftb_plan* p, const sfftb_config* cfg) {
    int rc0 = plan_init_common(p, cfg);
    if (rc0) return rc0;
    sfftb_dims& d = p->d;
    RowArgs& r = p->row;
    d.Fij = (d.DK + 1) * (d.DK + 2) / 2; d.Fpq = (d.DB + 1) * (d.DB + 2) / 2;
    d.Fijab = d.Fij * d.Fab; d.NEQ = d.Fijab + d.Fpq;
    // sca_degree: with const_phot_ratio set, 0 ties the centre taps to ONE constant (stripes dropped) and
    // DS > 0 gives them their own polynomial of degree DS <= DK (SEPARATE-VARYING, sfft/BSplineSFFT.py:77-86, 176-190)
    const int DS = cfg->const_phot_ratio ? cfg->sca_degree : 0;
    if (DS < 0 || DS > d.DK) return fail(SFFTB_EINVAL, "scaling degree %d must lie in 0..KerPolyOrder", DS);
    p->sca_n = DS > 0 ? (DS + 1) * (DS + 2) / 2 : 0;
    d.NEQ_FSfree = cfg->const_phot_ratio ? d.NEQ - (d.Fij - (p->sca_n ? p->sca_n : 1)) : d.NEQ;
    const int N0 = d.N0, N1 = d.N1, NH = N1 / 2 + 1;
    const size_t csz = cfg->storage == SFFTB_STORE_F32 ? sizeof(float2) : sizeof(double2);
    p->rinv.DB = d.DB; p->rinv.Fpq = d.Fpq;
    {
        int k = 0;
        for (int pp = 0; pp <= d.DB; ++pp)
            for (int q = 0; q <= d.DB - pp; ++q) { p->rinv.p_of[k] = (unsigned char)pp; p->rinv.q_of[k] = (unsigned char)q; ++k; }
    }

    // ---- column pass geometry: fold factor V, slice length M ----
    fill_col_common(p->cfit, p);
    const int Mmax = env_int("SFFTB_MMAX", 512);
    const int ntot = p->cfit.npairs + d.Fij;
    bool okf = false;
    for (int V = 1; V <= N0 && !okf; ++V) {
        if (N0 % V) continue;
        const int M = N0 / V;
        if (cfg->fold > 0 && V != cfg->fold) continue;
        if (cfg->fold <= 0 && M > Mmax) continue;
        FftDesc fd;
        if (!make_fft_desc(M, &fd) || !fft_fits_threads(fd, NT_COL)) continue;
        if (!okf) {
            ColArgs& c = p->cfit;
            c.V = V; c.M = M; c.pitch = M + 1; c.fd = fd;
            int PB = env_int("SFFTB_PB", ntot);
            if (PB > ntot) PB = ntot;
            while (PB > 1 && fit_smem_bytes(c, PB) > p->max_smem) --PB;
            if (fit_smem_bytes(c, PB) <= p->max_smem) { c.PB = PB; okf = true; }
        }
    }
    // KerPolyOrder = 3 exists only as the warp-specialised kernel (three launches); it needs the folded path as fallback
    const bool want_seg = (d.DK <= 2 || okf) && 4 * d.w0 + 32 <= FS3_M && cfg->fold <= 0 && !env_int("SFFTB_FIT_NOSEG", 0);
    if (!okf && !want_seg)
        return fail(SFFTB_EINVAL, "unsupported image height N0=%d (fold=%d): no slice length with prime factors <= 13 fits shared memory",
                    N0, cfg->fold);
    if (okf) {
        d.fold = p->cfit.V; d.sub_len = p->cfit.M;
        if (upload_twiddles(p->cfit.M, &p->twMf)) return SFFTB_ECUDA;
        p->cfit.twM = p->twMf;
        p->smem_fit = fit_smem_bytes(p->cfit, p->cfit.PB);
    } else {
        p->cfit.V = 1; p->cfit.M = N0; p->cfit.pitch = N0 + 1; p->cfit.PB = 1;
        p->smem_fit = 0;
    }
    {
        FirArgs& fa = p->fir;
        memset(&fa, 0, sizeof fa);
        fa.N0 = N0; fa.N1 = N1; fa.NH = NH; fa.DK = d.DK; fa.Fij = d.Fij; fa.nj = d.DK + 1; fa.w0 = d.w0; fa.w1 = d.w1;
        memcpy(fa.plane_of, p->cfit.plane_of, sizeof fa.plane_of);
        fa.tw1 = p->tw1;
        p->smem_fir = sizeof(cd) * (size_t)d.Fij * d.L0 + sizeof(double) * (size_t)d.Fij;
        const size_t WP3 = (FIR3_CH + 2 * (size_t)d.w0 + FIR3_R - 1) / FIR3_R;
        p->smem_fir3 = sizeof(cd) * (size_t)d.Fij * d.L0 + 128 + sizeof(cd) * (size_t)(d.DK + 1) * FIR3_R * WP3 + sizeof(double) * FIR3_R * WP3;
        CK(cudaMalloc(&p->firTaps, sizeof(cd) * (size_t)NH * d.Fij * d.L0));
        CK(cudaMalloc(&p->firCA, sizeof(double) * 16));
    }

    // ---- workspaces ----
    CK(cudaMalloc(&p->gI, csz * (size_t)(d.DK + 1) * NH * N0));
    if (cfg->storage == SFFTB_STORE_F32) CK(cudaMalloc(&p->gIa, sizeof(double2) * (size_t)(d.DK + 1) * NH * N0));
    else p->gIa = p->gI;
    p->nrowsK = p->cfit.npairs * p->cfit.nl0 + d.Fij * p->cfit.nlj0;
    p->nrowsL = d.Fij * (d.DB + 1) * p->cfit.nlj0;
    CK(cudaMalloc(&p->kap, sizeof(cd) * (size_t)p->nrowsK * NH));
    CK(cudaMalloc(&p->lam, sizeof(cd) * (size_t)p->nrowsL * NH));
    CK(cudaMalloc(&p->nuJ, sizeof(cd) * (size_t)(d.DB + 1) * NH));
    const int nl0 = 4 * d.w0 + 1, nl1 = 4 * d.w1 + 1, nlj0 = 2 * d.w0 + 1, nlj1 = 2 * d.w1 + 1;
    CK(cudaMalloc(&p->R, sizeof(double) * (size_t)p->cfit.npairs * nl0 * nl1));
    CK(cudaMalloc(&p->RJ, sizeof(double) * (size_t)d.Fij * nlj0 * nlj1));
    CK(cudaMalloc(&p->RT, sizeof(double) * (size_t)d.Fij * d.Fpq * nlj0 * nlj1));
    CK(cudaMalloc(&p->RJT, sizeof(double) * (size_t)d.Fpq));
    p->nsolve = d.NEQ_FSfree;
    p->ld = p->nsolve + 1;
    CK(cudaMalloc(&p->Aug, sizeof(double) * (size_t)(p->nsolve + 1) * p->ld));
    CK(cudaMalloc(&p->sc, sizeof(double) * (size_t)p->nsolve));
    CK(cudaMalloc(&p->diagU, sizeof(double) * (size_t)p->nsolve));
    CK(cudaMalloc(&p->sol, sizeof(double) * (size_t)d.NEQ));
    if (chol_setup(p)) return SFFTB_ECUDA;

    // ---- index maps (forbidden stripes: SFFTSubtract.py:82-90) ----
    {
        std::vector<int> ident(d.NEQ), idx;
        for (int k = 0; k < d.NEQ; ++k) ident[k] = k;
        for (int k = 0; k < d.NEQ; ++k) {
            bool forbidden = false;
            if (cfg->const_phot_ratio && k < d.Fijab) {
                const int A = k / d.Fab, ab = k % d.Fab;
                forbidden = (A >= (p->sca_n ? p->sca_n : 1)) && (ab == d.w0 * d.L1 + d.w1);
            }
            if (!forbidden) idx.push_back(k);
        }
        if ((int)idx.size() != p->nsolve) return fail(SFFTB_EINVAL, "internal: index map size mismatch");
        CK(cudaMalloc(&p->idxmap, sizeof(int) * idx.size()));
        CK(cudaMalloc(&p->ident, sizeof(int) * ident.size()));
        CK(cudaMemcpy(p->idxmap, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(p->ident, ident.data(), sizeof(int) * ident.size(), cudaMemcpyHostToDevice));
    }

    // ---- Phi block: sum_x T_p'q' T_pq from power sums ----
    {
        std::vector<long double> sx(2 * d.DB + 1, 0.0L), sy(2 * d.DB + 1, 0.0L);
        for (int e = 0; e <= 2 * d.DB; ++e) {
            for (int rr = 0; rr < N0; ++rr) sx[e] += powl((long double)(rr + 1) / N0, e);
            for (int c = 0; c < N1; ++c) sy[e] += powl((long double)(c + 1) / N1, e);
        }
        std::vector<double> phi((size_t)d.Fpq * d.Fpq);
        for (int a = 0; a < d.Fpq; ++a)
            for (int b = 0; b < d.Fpq; ++b)
                phi[(size_t)a * d.Fpq + b] = (double)(sx[p->rinv.p_of[a] + p->rinv.p_of[b]] * sy[p->rinv.q_of[a] + p->rinv.q_of[b]]);
        CK(cudaMalloc(&p->PHI, sizeof(double) * phi.size()));
        CK(cudaMemcpy(p->PHI, phi.data(), sizeof(double) * phi.size(), cudaMemcpyHostToDevice));
    }

    // ---- Q table ----
    CK(cudaMalloc(&p->Q, sizeof(cd) * (size_t)(d.DB + 1) * NH));
    {
        const int nwarps = (d.DB + 1) * NH;
        qtable_kernel<<<(nwarps * 32 + 255) / 256, 256, 0, p->stream>>>(N1, NH, d.DB + 1, p->tw1, p->Q);
        CKL(p);
    }

    // ---- reduction / fill arguments ----
    ReduceArgs& ra = p->red;
    ra.N1 = N1; ra.NH = NH; ra.w1 = d.w1; ra.nOm = p->cfit.npairs * nl0; ra.nrows = p->nrowsK; ra.tw1 = p->tw1;
    PolyReduceArgs& pa = p->pred;
    memset(&pa, 0, sizeof pa);
    pa.N1 = N1; pa.NH = NH; pa.w1 = d.w1; pa.DB = d.DB; pa.Fpq = d.Fpq; pa.Fij = d.Fij;
    pa.nlj0 = nlj0; pa.nlj1 = nlj1; pa.nrowsL = p->nrowsL; pa.tw1 = p->tw1; pa.Q = p->Q;
    memset(pa.pq_of, 0xff, sizeof pa.pq_of);
    for (int k = 0; k < d.Fpq; ++k) pa.pq_of[p->rinv.p_of[k]][p->rinv.q_of[k]] = (signed char)k;
    FillArgs& f = p->fill;
    f.Fij = d.Fij; f.Fpq = d.Fpq; f.Fab = d.Fab; f.Fijab = d.Fijab; f.L0 = d.L0; f.L1 = d.L1; f.w0 = d.w0; f.w1 = d.w1;
    f.nl0 = nl0; f.nl1 = nl1; f.nlj0 = nlj0; f.nlj1 = nlj1;
    const double N = (double)N0 * (double)N1;
    f.invN = 1.0 / N; f.invN2 = 1.0 / (N * N); f.invN3 = 1.0 / (N * N * N);
    f.R = p->R; f.RJ = p->RJ; f.RT = p->RT; f.RJT = p->RJT; f.PHI = p->PHI;
    f.sca_on = p->sca_n > 0 ? 1 : 0;
    memset(f.sca, 0xff, sizeof f.sca);
    if (p->sca_n) {
        // k-th scaling basis function (i, j), enumerated like the kernel's but up to degree DS (ScaREF_ij, :2779-2786)
        int k = 0;
        for (int i = 0; i <= DS; ++i)
            for (int j = 0; j <= DS - i; ++j) f.sca[k++] = (signed char)p->cfit.plane_of[i][j];
        CK(cudaMalloc(&p->solEff, sizeof(double) * (size_t)d.NEQ));
    }

    const bool f32 = cfg->storage == SFFTB_STORE_F32;
    if (apply_setup(p) || rows_setup(p)) return SFFTB_ECUDA;

    // ---- segmented fit path (KerPolyOrder <= 2, halo 2 w0 well inside a 256-point window) ----
    p->fit_seg = 0;
    if (want_seg) {
        SegFitArgs& sf = p->sfit;
        memset(&sf, 0, sizeof sf);
        sf.c = p->cfit;
        sf.h = 2 * d.w0;
        const int Smax = (FS3_M - 2 * sf.h) & ~1;
        sf.nseg = (N0 + Smax - 1) / Smax;
        sf.S = (((N0 + sf.nseg - 1) / sf.nseg) + 1) & ~1;       // balanced, even
        if (sf.S > Smax) sf.S = Smax;
        sf.nseg = (N0 + sf.S - 1) / sf.S;
        sf.nOm = p->cfit.npairs * nl0;
        sf.nK = sf.nOm + d.Fij * nlj0;
        sf.nLT = d.Fij * d.Fpq * nlj0;
        sf.nrows = sf.nK + sf.nLT + d.Fpq;
        sf.tabA = p->tabA;
        sf.Q = p->Q;
        memcpy(sf.pq_of, pa.pq_of, sizeof sf.pq_of);
        CK(cudaMalloc(&p->kap2, sizeof(cd) * (size_t)NH * sf.nrows));
        CK(cudaMalloc(&p->momg, sizeof(cd) * (size_t)NH * 5 * SFFTB_MAXE));
        sf.momg = p->momg; sf.mom_external = env_int("SFFTB_MOM_EXTERNAL", 1);
        LagReduce2Args& r2 = p->red2;
        r2.N1 = N1; r2.NH = NH; r2.nrows = sf.nrows; r2.w1 = d.w1; r2.tw1 = p->tw1; r2.rb0 = 0;
        const int rowblocks = (sf.nrows + 15) / 16;
        r2.ksplit = (NH + LR2_KC - 1) / LR2_KC;
        (void)rowblocks;
        CK(cudaMalloc(&p->part, sizeof(double) * (size_t)r2.ksplit * sf.nrows * nl1));
        LagFinishArgs& f2 = p->fin2;
        f2.nrows = sf.nrows; f2.nOm = sf.nOm; f2.nK = sf.nK; f2.nLT = sf.nLT; f2.w1 = d.w1; f2.ksplit = r2.ksplit;
        f2.R = p->R; f2.RJ = p->RJ; f2.RT = p->RT; f2.RJT = p->RJT;
        p->grid_sfit = std::min(NH, p->nsm);
        {
            p->smem_sfit3 = fs4_smem_bytes(d.DK, f32);
            if (p->smem_sfit3 <= p->max_smem) p->fit_seg = 2;
        }
        if (p->fit_seg) { d.fold = sf.nseg; d.sub_len = sf.S; }
    }
    if (!p->fit_seg && !okf)
        return fail(SFFTB_EINVAL, "unsupported image height N0=%d (fold=%d): no slice length with prime factors <= 13 fits shared memory",
                    N0, cfg->fold);
    p->fit_generic_ok = okf ? 1 : 0;
    if (fit_setup(p)) return SFFTB_ECUDA;
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}


extern "C" int sfftb_plan_create(sfftb_plan** out, const sfftb_config* cfg) {
    if (!out || !cfg) return fail(SFFTB_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->N0 < 2 || cfg->N1 < 2) return fail(SFFTB_EINVAL, "image shape (%d, %d) too small", cfg->N0, cfg->N1);
    if (cfg->DK < 0 || cfg->DK > 3) return fail(SFFTB_EINVAL, "Input KerPolyOrder should be 0/1/2/3!");
    if (cfg->DB < 0 || cfg->DB > 3) return fail(SFFTB_EINVAL, "Input BGPolyOrder should be 0/1/2/3!");
    if (cfg->w0 < 0 || cfg->w1 < 0 || 2 * cfg->w0 + 1 > cfg->N0 || 2 * cfg->w1 + 1 > cfg->N1)
        return fail(SFFTB_EINVAL, "kernel half width (%d, %d) does not fit the image (%d, %d)", cfg->w0, cfg->w1, cfg->N0, cfg->N1);
    if (cfg->storage != SFFTB_STORE_F64 && cfg->storage != SFFTB_STORE_F32) return fail(SFFTB_EINVAL, "bad storage precision");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(SFFTB_EINVAL, "CUDA device %d not available (%d visible)", cfg->device, ndev);
    sfftb_plan* p = new sfftb_plan();
    memset((void*)p, 0, sizeof *p);
    const int rc = plan_create_impl(p, cfg);
    if (rc) { std::string keep = g_err; plan_free(p); g_err = keep; return rc; }
    *out = p;
    return 0;
}

extern "C" int sfftb_plan_create_general(sfftb_plan** out, const sfftb_config* cfg, const sfftb_basis* ker, const sfftb_basis* sca,
                                         const sfftb_basis* bkg, int scaling_mode) {
    if (!out || !cfg) return fail(SFFTB_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->N0 < 2 || cfg->N1 < 2) return fail(SFFTB_EINVAL, "image shape (%d, %d) too small", cfg->N0, cfg->N1);
    if (cfg->w0 < 0 || cfg->w1 < 0 || 2 * cfg->w0 + 1 > cfg->N0 || 2 * cfg->w1 + 1 > cfg->N1)
        return fail(SFFTB_EINVAL, "kernel half width (%d, %d) does not fit the image (%d, %d)", cfg->w0, cfg->w1, cfg->N0, cfg->N1);
    if (cfg->storage != SFFTB_STORE_F64 && cfg->storage != SFFTB_STORE_F32) return fail(SFFTB_EINVAL, "bad storage precision");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(SFFTB_EINVAL, "CUDA device %d not available (%d visible)", cfg->device, ndev);
    sfftb_plan* p = new sfftb_plan();
    memset((void*)p, 0, sizeof *p);
    const int rc = gen_plan_create(p, cfg, ker, sca, bkg, scaling_mode);
    if (rc) { std::string keep = g_err; plan_free(p); g_err = keep; return rc; }
    *out = p;
    return 0;
}

extern "C" int sfftb_plan_destroy(sfftb_plan* p) { return plan_free(p); }

extern "C" int sfftb_plan_dims(const sfftb_plan* p, sfftb_dims* out) {
    if (!p || !out) return fail(SFFTB_EINVAL, "null argument");
    *out = p->d;
    return 0;
}

extern "C" int sfftb_plan_set_stream(sfftb_plan* p, void* s) {
    if (!p) return fail(SFFTB_EINVAL, "null plan");
    p->stream = s ? (cudaStream_t)s : p->own_stream;
    return 0;
}

extern "C" int sfftb_plan_set_partition(sfftb_plan* p, int solver_sms) {
    if (!p) return fail(SFFTB_EINVAL, "null plan");
    if (solver_sms < 0 || solver_sms >= p->nsm) return fail(SFFTB_EINVAL, "solver partition of %d SMs does not fit a device with %d", solver_sms, p->nsm);
    p->solver_sms = solver_sms;
    return 0;
}

extern "C" int sfftb_plan_sync(sfftb_plan* p) {
    if (!p) return fail(SFFTB_EINVAL, "null plan");
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" int sfftb_plan_set_timing(sfftb_plan* p, int enable) {
    if (!p) return fail(SFFTB_EINVAL, "null plan");
    p->timing = enable;
    return 0;
}

extern "C" int sfftb_gen_info(const sfftb_plan* p, int* out8) {
    if (!p || !out8) return fail(SFFTB_EINVAL, "null argument");
    if (!p->gen) return fail(SFFTB_EINVAL, "not a general-basis plan");
    gen_info(p, out8);
    return 0;
}
extern "C" long long sfftb_launch_count(const sfftb_plan* p) { return p ? p->launches : 0; }
extern "C" int sfftb_last_solver(const sfftb_plan* p) { return p ? p->last_solver : 0; }


// ---------------------------------------------------------------------------------------------------------------
static int stage_in(sfftb_plan* p, const void* src, int memkind, int dtype, void* staging, const void** dev) {
    if (!src) return fail(SFFTB_EINVAL, "null image pointer");
    if (dtype != SFFTB_F64 && dtype != SFFTB_F32) return fail(SFFTB_EINVAL, "bad image dtype");
    if (memkind == SFFTB_MEM_DEVICE) { *dev = src; return 0; }
    if (memkind != SFFTB_MEM_HOST) return fail(SFFTB_EINVAL, "bad memkind");
    const size_t bytes = (size_t)p->d.N0 * p->d.N1 * (dtype == SFFTB_F64 ? 8 : 4);
    CK(cudaMemcpyAsync(staging, src, bytes, cudaMemcpyHostToDevice, p->stream));
    *dev = staging;
    return 0;
}


static int fill_system(sfftb_plan* p) {
    const int n = p->nsolve;
    fill_diag_kernel<<<(n + 127) / 128, 128, 0, p->stream>>>(p->fill, p->idxmap, n, p->sc, p->info);
    CKL(p);
    dim3 blk(32, 8), grd((n + 1 + 31) / 32, (n + 1 + 7) / 8);
    fill_matrix_kernel<<<grd, blk, 0, p->stream>>>(p->fill, p->idxmap, n, p->sc, p->Aug, p->ld, p->info);
    CKL(p);
    return 0;
}


// ovI / ovJ != NULL (device images of the apply step): their forward row pass is issued on the side stream right
// after the normal equations are assembled, on half of the SMs, while the Cholesky solve runs on the other half
// (both kernels need a whole SM per CTA, so they cannot share one; the solve is latency bound and hardly notices).
// The row spectra buffers gI / gJ are free at that point.  apply_device is then called with rows_done = true.
template <typename TSt>
static int fit_device(sfftb_plan* p, const void* dI, const void* dJ, int dtype, const void* tI = nullptr,
                      const void* ovI = nullptr, const void* ovJ = nullptr) {
    const sfftb_dims& d = p->d;
    EVREC(p, EV_START);
    if (p->gen) {
        // general-basis plan: stored planes through the column tables, block passes, descriptor-driven fill
        if (launch_row_fwd<TSt>(p, dI, dtype, (TSt*)gen_planes(p), gen_nvs(p), p->vtab)) return SFFTB_ECUDA;
        if (launch_row_fwd<TSt>(p, dJ, dtype, (TSt*)p->gJ, 1)) return SFFTB_ECUDA;
        if (gen_rjt(p, dJ, dtype)) return SFFTB_ECUDA;
        EVREC(p, EV_ROWS);
        if (gen_fit_cols<TSt>(p)) return SFFTB_ECUDA;
        CK(cudaMemsetAsync(p->info, 0, sizeof(int) * 4, p->stream));
        if (gen_fill_system(p)) return SFFTB_ECUDA;
        EVREC(p, EV_RED);
        if (run_cholesky(p)) return SFFTB_ECUDA;
        if (gen_restore(p)) return SFFTB_ECUDA;
        EVREC(p, EV_SOLVE);
        CK(cudaMemcpyAsync(p->info_h, p->info, sizeof(int) * 4, cudaMemcpyDeviceToHost, p->stream));
        p->have_fit = 1;
        p->last_solver = 1;
        return 0;
    }
    const TSt* gIsrc = tI ? (const TSt*)tI : (const TSt*)p->gI;
    if (p->pendI) { CK(cudaStreamWaitEvent(p->stream, p->pendI, 0)); p->pendI = nullptr; }
    if (!tI && launch_row_fwd<TSt>(p, dI, dtype, (TSt*)p->gI, d.DK + 1)) return SFFTB_ECUDA;
    if (p->pendJ) { CK(cudaStreamWaitEvent(p->stream, p->pendJ, 0)); p->pendJ = nullptr; }
    if (launch_row_fwd<TSt>(p, dJ, dtype, (TSt*)p->gJ, 1)) return SFFTB_ECUDA;
    EVREC(p, EV_ROWS);
    const bool jonly = tI && p->factor_cached && p->chol_coop && p->fit_seg == 2 && d.DK <= 2 && !env_int("SFFTB_TEMPLATE_FULLFIT", 0);
    if (launch_fit_cols<TSt>(p, gIsrc, jonly)) return SFFTB_ECUDA;
    CK(cudaMemsetAsync(p->info, 0, sizeof(int) * 4, p->stream));
    bool used_cache = false;
    if (tI && p->factor_cached && p->chol_coop) {
        // LHMAT depends on the (masked) template only: reuse its factor, assemble just the right-hand side
        fill_rhs_kernel<<<(p->nsolve + 127) / 128, 128, 0, p->stream>>>(p->fill, p->idxmap, p->nsolve, p->sc, p->cholY, p->info);
        CKL(p);
        EVREC(p, EV_RED);
        if (run_cholesky(p, 1)) return SFFTB_ECUDA;
        p->resolves++;
        used_cache = true;
    } else {
        p->factor_cached = 0;            // Aug is about to be overwritten
        if (fill_system(p)) return SFFTB_ECUDA;
        EVREC(p, EV_RED);
        const bool ov = ovI && ovJ && (p->row_v8 || p->row_h16 || p->row_g16 || (p->row_blu && p->row_blu <= 16)) && p->chol_coop && p->nsm >= 8;
        if (ov) {
            CK(cudaEventRecord(p->evFork, p->stream));
            CK(cudaStreamWaitEvent(p->stream2, p->evFork, 0));
            cudaStream_t main = p->stream;
            p->stream = p->stream2;
            const int rowsm = p->solver_sms > 0 ? work_sms(p)
                                                : std::max(1, std::min(p->nsm - 1, p->nsm * env_int("SFFTB_OVERLAP_ROWS_PCT", 50) / 100));
            p->row_grid_limit = rowsm;
            int rc2 = launch_row_fwd<double2>(p, ovI, dtype, (double2*)p->gIa, d.DK + 1);
            if (!rc2) rc2 = launch_row_fwd<double2>(p, ovJ, dtype, (double2*)p->gJa, 1);
            p->row_grid_limit = 0;
            p->stream = main;
            if (rc2) return SFFTB_ECUDA;
            CK(cudaEventRecord(p->evJoin, p->stream2));
            p->chol_grid_limit = p->nsm - rowsm;
        }
        if (p->solver_sms > 0) p->chol_grid_limit = p->solver_sms;
        const int rcc = run_cholesky(p);
        p->chol_grid_limit = 0;
        if (rcc) return SFFTB_ECUDA;
        if (ov) CK(cudaStreamWaitEvent(p->stream, p->evJoin, 0));
    }
    EVREC(p, EV_SOLVE);
    CK(cudaMemcpyAsync(p->info_h, p->info, sizeof(int) * 4, cudaMemcpyDeviceToHost, p->stream));
    p->have_fit = 1;
    p->last_solver = used_cache ? 3 : 1;
    return 0;
}

// After a stream sync: inspect the solver flags; run the pivoted-LU fallback if Cholesky broke down.
// Returns 0 (ok), 1 (fallback was run, solution replaced) or a negative error.
static int check_solver(sfftb_plan* p) {
    if (p->info_h[1]) return fail(SFFTB_ENOTFINITE, "array must not contain infs or NaNs (normal equations)");
    if (p->info_h[0] == 0) return 0;
    // Cholesky pivot not positive: refill and solve by LU with partial pivoting
    CK(cudaMemsetAsync(p->info, 0, sizeof(int) * 4, p->stream));
    if (p->gen ? gen_fill_system(p) : fill_system(p)) return SFFTB_ECUDA;
    if (run_lu(p)) return SFFTB_ECUDA;
    if (p->gen && gen_restore(p)) return SFFTB_ECUDA;
    CK(cudaMemcpyAsync(p->info_h, p->info, sizeof(int) * 4, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    p->last_solver = 2;
    if (p->info_h[1]) return fail(SFFTB_ENOTFINITE, "array must not contain infs or NaNs (normal equations)");
    if (p->info_h[2]) return fail(SFFTB_ESINGULAR, "Singular matrix (zero pivot at column %d of the normal equations)", p->info_h[2] - 1);
    return 1;
}

// SEPARATE-VARYING scaling: the centre-tap coefficient of plane k multiplies the image times the k-th scaling basis
// function, which is the unshifted plane sca[k] of the kernel's plane set.  Moving those coefficients onto the centre
// taps of their planes turns the model into the standard form the FIR apply evaluates (Construct_FDIFF, :2487-2507).
__global__ void sca_remap_kernel(FillArgs f, int NEQ, const double* __restrict__ sol, double* __restrict__ eff) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= NEQ) return;
    double v = sol[k];
    const int c0 = f.w0 * f.L1 + f.w1;
    if (k < f.Fijab && k % f.Fab == c0) {
        const int A = k / f.Fab;
        v = 0.0;
        for (int q = 0; q < f.Fij; ++q)
            if (f.sca[q] == A) v += sol[q * f.Fab + c0];
    }
    eff[k] = v;
}

template <typename TSt>
static int apply_device(sfftb_plan* p, const void* dI, const void* dJ, int dtype, const double* dsol, void* ddiff, int diff_dtype,
                        const void* tI = nullptr, bool rows_done = false, void* hdiff = nullptr) {
    const sfftb_dims& d = p->d;
    EVREC(p, EV_A0);
    if (p->gen) {
        if (!rows_done) {
            if (launch_row_fwd<double2>(p, dI, dtype, (double2*)gen_planes_apply(p), gen_nvs(p), p->vtab)) return SFFTB_ECUDA;
            if (launch_row_fwd<double2>(p, dJ, dtype, (double2*)p->gJa, 1)) return SFFTB_ECUDA;
        }
        EVREC(p, EV_AROWS);
        if (gen_fir(p, dsol)) return SFFTB_ECUDA;
        EVREC(p, EV_ACOL);
        if (launch_row_inv(p, dsol + d.Fijab, ddiff, diff_dtype, nullptr)) return SFFTB_ECUDA;
        if (gen_bkg_subtract(p, dsol + d.Fijab, ddiff, diff_dtype)) return SFFTB_ECUDA;
        EVREC(p, EV_AINV);
        return 0;
    }
    if (p->sca_n) {
        sca_remap_kernel<<<(d.NEQ + 255) / 256, 256, 0, p->stream>>>(p->fill, d.NEQ, dsol, p->solEff);
        CKL(p);
        dsol = p->solEff;
    }
    // (the apply step works on fp64 spectra whatever the storage of the fit spectra: TSt is not used here)
    const double2* gIsrc = tI ? (const double2*)tI : (const double2*)p->gIa;
    if (!rows_done) {
        if (p->pendI) { CK(cudaStreamWaitEvent(p->stream, p->pendI, 0)); p->pendI = nullptr; }
        if (!tI && launch_row_fwd<double2>(p, dI, dtype, (double2*)p->gIa, d.DK + 1)) return SFFTB_ECUDA;
        if (p->pendJ) { CK(cudaStreamWaitEvent(p->stream, p->pendJ, 0)); p->pendJ = nullptr; }
        if (launch_row_fwd<double2>(p, dJ, dtype, (double2*)p->gJa, 1)) return SFFTB_ECUDA;
    }
    EVREC(p, EV_AROWS);
    if (launch_fir(p, gIsrc, dsol)) return SFFTB_ECUDA;
    EVREC(p, EV_ACOL);
    const double* bpq = dsol + d.Fijab;
    if (launch_row_inv(p, bpq, ddiff, diff_dtype, hdiff)) return SFFTB_ECUDA;
    EVREC(p, EV_AINV);
    return 0;
}

static int collect_timings(sfftb_plan* p, bool fit, bool app) {
    if (!p->timing) return 0;
    if (fit) {
        CK(cudaEventElapsedTime(&p->ms[0], p->ev[EV_START], p->ev[EV_ROWS]));
        CK(cudaEventElapsedTime(&p->ms[1], p->ev[EV_ROWS], p->ev[EV_COL]));
        CK(cudaEventElapsedTime(&p->ms[2], p->ev[EV_COL], p->ev[EV_RED]));
        CK(cudaEventElapsedTime(&p->ms[3], p->ev[EV_RED], p->ev[EV_SOLVE]));
        if (p->fit_seg && !p->gen) CK(cudaEventElapsedTime(&p->ms[7], p->ev[EV_KFIT0], p->ev[EV_KFIT1]));
    }
    if (app) {
        CK(cudaEventElapsedTime(&p->ms[4], p->ev[EV_A0], p->ev[EV_AROWS]));
        CK(cudaEventElapsedTime(&p->ms[5], p->ev[EV_AROWS], p->ev[EV_ACOL]));
        CK(cudaEventElapsedTime(&p->ms[6], p->ev[EV_ACOL], p->ev[EV_AINV]));
    }
    return 0;
}

extern "C" int sfftb_timings(sfftb_plan* p, float* ms, int n) {
    if (!p || !ms) return fail(SFFTB_EINVAL, "null argument");
    for (int k = 0; k < n && k < 8; ++k) ms[k] = p->ms[k];
    return 0;
}

static int copy_out(sfftb_plan* p, void* dst, int memkind, const void* src_dev, size_t bytes) {
    if (!dst) return 0;
    CK(cudaMemcpyAsync(dst, src_dev, bytes, memkind == SFFTB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, p->stream));
    return 0;
}

extern "C" int sfftb_fit(sfftb_plan* p, const void* I, const void* J, int memkind, int dtype, double* solution, int sol_memkind) {
    if (!p) return fail(SFFTB_EINVAL, "null plan");
    CK(cudaSetDevice(p->device));
    const void *dI, *dJ;
    int rc;
    if ((rc = stage_in(p, I, memkind, dtype, p->stA, &dI))) return rc;
    if ((rc = stage_in(p, J, memkind, dtype, p->stB, &dJ))) return rc;
    rc = p->cfg.storage == SFFTB_STORE_F32 ? fit_device<float2>(p, dI, dJ, dtype) : fit_device<double2>(p, dI, dJ, dtype);
    if (rc) return rc;
    CK(cudaStreamSynchronize(p->stream));
    if ((rc = check_solver(p)) < 0) return rc;
    if ((rc = collect_timings(p, true, false))) return rc;
    if ((rc = copy_out(p, solution, sol_memkind, p->sol, sizeof(double) * p->d.NEQ))) return rc;
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" int sfftb_apply(sfftb_plan* p, const void* I, const void* J, int memkind, int dtype, const double* solution, int sol_memkind,
                           void* diff, int diff_memkind, int diff_dtype) {
    if (!p || !solution || !diff) return fail(SFFTB_EINVAL, "null argument");
    if (diff_dtype != SFFTB_F64 && diff_dtype != SFFTB_F32) return fail(SFFTB_EINVAL, "bad diff dtype");
    CK(cudaSetDevice(p->device));
    const void *dI, *dJ;
    int rc;
    if ((rc = stage_in(p, I, memkind, dtype, p->stA, &dI))) return rc;
    if ((rc = stage_in(p, J, memkind, dtype, p->stB, &dJ))) return rc;
    const double* dsol = solution;
    if (sol_memkind == SFFTB_MEM_HOST) {
        CK(cudaMemcpyAsync(p->sol, solution, sizeof(double) * p->d.NEQ, cudaMemcpyHostToDevice, p->stream));
        dsol = p->sol;
    }
    void* ddiff = diff_memkind == SFFTB_MEM_DEVICE ? diff : p->stA;   // stA is free again once the row pass has consumed it
    rc = p->cfg.storage == SFFTB_STORE_F32 ? apply_device<float2>(p, dI, dJ, dtype, dsol, ddiff, diff_dtype)
                                           : apply_device<double2>(p, dI, dJ, dtype, dsol, ddiff, diff_dtype);
    if (rc) return rc;
    if (diff_memkind == SFFTB_MEM_HOST) {
        const size_t bytes = (size_t)p->d.N0 * p->d.N1 * (diff_dtype == SFFTB_F64 ? 8 : 4);
        CK(cudaMemcpyAsync(diff, p->stA, bytes, cudaMemcpyDeviceToHost, p->stream));
    }
    CK(cudaStreamSynchronize(p->stream));
    return collect_timings(p, false, true);
}

extern "C" int sfftb_gss(sfftb_plan* p, const void* I, const void* J, const void* mI, const void* mJ, int memkind, int dtype,
                         double* solution, int sol_memkind, void* diff, int diff_memkind, int diff_dtype) {
    if (!p || !diff) return fail(SFFTB_EINVAL, "null argument");
    if (diff_dtype != SFFTB_F64 && diff_dtype != SFFTB_F32) return fail(SFFTB_EINVAL, "bad diff dtype");
    CK(cudaSetDevice(p->device));
    const bool f32 = p->cfg.storage == SFFTB_STORE_F32;
    const void *dI, *dJ;
    int rc;
    // host images: all four H2D copies are queued on the side stream up front, so the copies of the apply pair run
    // under the fit and every row pass starts as soon as its own image has landed
    const bool hostpipe = memkind == SFFTB_MEM_HOST && I && J && mI && mJ && (dtype == SFFTB_F64 || dtype == SFFTB_F32) &&
                          !p->gen && !env_int("SFFTB_NO_HOSTPIPE", 0);
    if (hostpipe) {
        const size_t bytes = (size_t)p->d.N0 * p->d.N1 * (dtype == SFFTB_F64 ? 8 : 4);
        if (!p->stC) { CK(cudaMalloc(&p->stC, sizeof(double) * (size_t)p->d.N0 * p->d.N1)); CK(cudaMalloc(&p->stD, sizeof(double) * (size_t)p->d.N0 * p->d.N1)); }
        CK(cudaEventRecord(p->evStart, p->stream));
        CK(cudaStreamWaitEvent(p->stream2, p->evStart, 0));
        const void* srcs[4] = {mI, mJ, I, J};
        void* dsts[4] = {p->stA, p->stB, p->stC, p->stD};
        for (int k = 0; k < 4; ++k) {
            CK(cudaMemcpyAsync(dsts[k], srcs[k], bytes, cudaMemcpyHostToDevice, p->stream2));
            CK(cudaEventRecord(p->evCopy[k], p->stream2));
        }
        p->pendI = p->evCopy[0]; p->pendJ = p->evCopy[1];
        rc = f32 ? fit_device<float2>(p, p->stA, p->stB, dtype) : fit_device<double2>(p, p->stA, p->stB, dtype);
        if (rc) return rc;
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (attempt == 0) { p->pendI = p->evCopy[2]; p->pendJ = p->evCopy[3]; }
            void* hd = (diff_memkind == SFFTB_MEM_HOST && p->row_fast) ? diff : nullptr;
            rc = f32 ? apply_device<float2>(p, p->stC, p->stD, dtype, p->sol, p->stA, diff_dtype, nullptr, false, hd)
                     : apply_device<double2>(p, p->stC, p->stD, dtype, p->sol, p->stA, diff_dtype, nullptr, false, hd);
            if (rc) return rc;
            CK(cudaStreamSynchronize(p->stream));
            if (attempt == 1) break;
            rc = check_solver(p);
            if (rc < 0) return rc;
            if (rc == 0) break;
        }
        if ((rc = collect_timings(p, true, true))) return rc;
        if (diff_memkind == SFFTB_MEM_HOST) {
            if (!p->row_fast) {
                const size_t ob = (size_t)p->d.N0 * p->d.N1 * (diff_dtype == SFFTB_F64 ? 8 : 4);
                CK(cudaMemcpyAsync(diff, p->stA, ob, cudaMemcpyDeviceToHost, p->stream));
            }
        } else {
            const size_t ob = (size_t)p->d.N0 * p->d.N1 * (diff_dtype == SFFTB_F64 ? 8 : 4);
            CK(cudaMemcpyAsync(diff, p->stA, ob, cudaMemcpyDeviceToDevice, p->stream));
        }
        if ((rc = copy_out(p, solution, sol_memkind, p->sol, sizeof(double) * p->d.NEQ))) return rc;
        CK(cudaStreamSynchronize(p->stream));
        return 0;
    }
    if ((rc = stage_in(p, mI, memkind, dtype, p->stA, &dI))) return rc;
    if ((rc = stage_in(p, mJ, memkind, dtype, p->stB, &dJ))) return rc;
    const bool ov = p->overlap && memkind == SFFTB_MEM_DEVICE && I && J && (p->row_v8 || p->row_h16 || p->row_g16 || (p->row_blu && p->row_blu <= 16)) && p->chol_coop && p->nsm >= 8 && !p->gen;
    rc = f32 ? fit_device<float2>(p, dI, dJ, dtype, nullptr, ov ? I : nullptr, ov ? J : nullptr)
             : fit_device<double2>(p, dI, dJ, dtype, nullptr, ov ? I : nullptr, ov ? J : nullptr);
    if (rc) return rc;
    void* ddiff = diff_memkind == SFFTB_MEM_DEVICE ? diff : p->stA;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if ((rc = stage_in(p, I, memkind, dtype, p->stA, &dI))) return rc;
        if ((rc = stage_in(p, J, memkind, dtype, p->stB, &dJ))) return rc;
        const bool rows_done = ov && attempt == 0;
        rc = f32 ? apply_device<float2>(p, dI, dJ, dtype, p->sol, ddiff, diff_dtype, nullptr, rows_done)
                 : apply_device<double2>(p, dI, dJ, dtype, p->sol, ddiff, diff_dtype, nullptr, rows_done);
        if (rc) return rc;
        CK(cudaStreamSynchronize(p->stream));
        if (attempt == 1) break;
        rc = check_solver(p);
        if (rc < 0) return rc;
        if (rc == 0) break;          // Cholesky was fine; rc == 1: solution replaced by the LU fallback -> apply again
    }
    if ((rc = collect_timings(p, true, true))) return rc;
    if (diff_memkind == SFFTB_MEM_HOST) {
        const size_t bytes = (size_t)p->d.N0 * p->d.N1 * (diff_dtype == SFFTB_F64 ? 8 : 4);
        CK(cudaMemcpyAsync(diff, p->stA, bytes, cudaMemcpyDeviceToHost, p->stream));
    }
    if ((rc = copy_out(p, solution, sol_memkind, p->sol, sizeof(double) * p->d.NEQ))) return rc;
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

// ---- asynchronous host-buffer GSS: submit queues the H2D copies (side stream), the fit, the apply and the D2H of the
// difference image and returns; finish waits for THIS plan's work only.  Two plans that share one compute stream
// (sfftb_plan_set_stream) and are driven alternately overlap the H2D copies of pair k + 1 with the kernels and the D2H
// of pair k (PCIe is full duplex), which is what bounds a stream of pairs coming from host memory.
// mI = I, mI[idx[k]] = val[k]: the masked images of the packets differ from the unmasked ones only inside the masked
// stamps (sfft/CustomizedPacket.py:114-162 fills / zeroes regions of the same arrays), so a host caller can send them as
// sparse deltas and halve the host-to-device traffic of a pair.
template <typename T>
__global__ void delta_scatter_kernel(T* __restrict__ img, const long long* __restrict__ idx, const T* __restrict__ val, long long n, long long npix) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const long long i = idx[k];
    if (i >= 0 && i < npix) img[i] = val[k];
}

struct DeltaSpec { long long nI, nJ; const long long *idxI, *idxJ; const void *valI, *valJ; };

// shared body of the asynchronous pair submissions.
//   memkind HOST  : the H2D copies run on the plan's copy stream; with `dl` the masked pair is rebuilt on the device
//                   from I, J and the sparse deltas (two image copies instead of four);
//   memkind DEVICE: the four images (and the outputs) are device buffers used in place; the forward row pass of the
//                   apply step overlaps the Cholesky like in sfftb_gss.
static int gss_submit_impl(sfftb_plan* p, const void* I, const void* J, const void* mI, const void* mJ, const DeltaSpec* dl, int memkind,
                           int dtype, double* solution, void* diff, int diff_dtype) {
    if ((dtype != SFFTB_F64 && dtype != SFFTB_F32) || (diff_dtype != SFFTB_F64 && diff_dtype != SFFTB_F32)) return fail(SFFTB_EINVAL, "bad dtype");
    if (p->gen) return fail(SFFTB_EINVAL, "the asynchronous pair submission is not available for general-basis plans");
    if (p->pending) return fail(SFFTB_ESTATE, "sfftb_gss_submit: the previous submission of this plan has not been finished");
    CK(cudaSetDevice(p->device));
    const bool f32 = p->cfg.storage == SFFTB_STORE_F32;
    const size_t npix = (size_t)p->d.N0 * p->d.N1, esz = dtype == SFFTB_F64 ? 8 : 4;
    const size_t bytes = npix * esz;
    int rc;
    p->pend_mem = memkind;
    if (memkind == SFFTB_MEM_DEVICE) {
        const bool ov = p->overlap && (p->row_v8 || p->row_h16 || p->row_g16 || (p->row_blu && p->row_blu <= 16)) && p->chol_coop && p->nsm >= 8;
        p->pendI = nullptr; p->pendJ = nullptr;
        rc = f32 ? fit_device<float2>(p, mI, mJ, dtype, nullptr, ov ? I : nullptr, ov ? J : nullptr)
                 : fit_device<double2>(p, mI, mJ, dtype, nullptr, ov ? I : nullptr, ov ? J : nullptr);
        if (rc) return rc;
        rc = f32 ? apply_device<float2>(p, I, J, dtype, p->sol, diff, diff_dtype, nullptr, ov)
                 : apply_device<double2>(p, I, J, dtype, p->sol, diff, diff_dtype, nullptr, ov);
        if (rc) return rc;
        if (solution) CK(cudaMemcpyAsync(solution, p->sol, sizeof(double) * p->d.NEQ, cudaMemcpyDeviceToDevice, p->stream));
        CK(cudaEventRecord(p->evDone, p->stream));
        p->pending = 1; p->pend_diff = diff; p->pend_sol = solution; p->pend_dtype = dtype; p->pend_diff_dtype = diff_dtype;
        p->pend_mode = 1; p->pend_I = I; p->pend_J = J;
        return 0;
    }
    if (!p->stC) { CK(cudaMalloc(&p->stC, sizeof(double) * npix)); CK(cudaMalloc(&p->stD, sizeof(double) * npix)); }
    // no wait on the compute stream here: the staging buffers are free (the previous submission was finished), and the
    // copies must not queue behind another plan's kernels on a shared compute stream
    if (!dl) {
        const void* srcs[4] = {mI, mJ, I, J};
        void* dsts[4] = {p->stA, p->stB, p->stC, p->stD};
        for (int k = 0; k < 4; ++k) {
            CK(cudaMemcpyAsync(dsts[k], srcs[k], bytes, cudaMemcpyHostToDevice, p->stream2));
            CK(cudaEventRecord(p->evCopy[k], p->stream2));
        }
    } else {
        const long long nd = dl->nI + dl->nJ;
        if ((size_t)nd > p->delta_cap) {
            if (p->deltaIdx) { CK(cudaFree(p->deltaIdx)); CK(cudaFree(p->deltaVal)); p->deltaIdx = nullptr; p->deltaVal = nullptr; }
            p->delta_cap = (size_t)nd + (size_t)nd / 4 + 1024;
            CK(cudaMalloc(&p->deltaIdx, sizeof(long long) * p->delta_cap));
            CK(cudaMalloc(&p->deltaVal, sizeof(double) * p->delta_cap));
        }
        char* dval = (char*)p->deltaVal;
        // copy stream: the unmasked pair and the deltas.  The masked pair is rebuilt on the COMPUTE stream (a device copy and a
        // scatter per image, ~0.07 ms per 4096^2 pair): kernels queued on the copy stream would have to wait for a free SM
        // behind the persistent compute kernels of the pair in flight.
        CK(cudaMemcpyAsync(p->stC, I, bytes, cudaMemcpyHostToDevice, p->stream2));
        if (dl->nI) {
            CK(cudaMemcpyAsync(p->deltaIdx, dl->idxI, sizeof(long long) * dl->nI, cudaMemcpyHostToDevice, p->stream2));
            CK(cudaMemcpyAsync(dval, dl->valI, esz * dl->nI, cudaMemcpyHostToDevice, p->stream2));
        }
        CK(cudaEventRecord(p->evCopy[0], p->stream2));
        CK(cudaMemcpyAsync(p->stD, J, bytes, cudaMemcpyHostToDevice, p->stream2));
        if (dl->nJ) {
            CK(cudaMemcpyAsync(p->deltaIdx + dl->nI, dl->idxJ, sizeof(long long) * dl->nJ, cudaMemcpyHostToDevice, p->stream2));
            CK(cudaMemcpyAsync(dval + esz * dl->nI, dl->valJ, esz * dl->nJ, cudaMemcpyHostToDevice, p->stream2));
        }
        CK(cudaEventRecord(p->evCopy[1], p->stream2));
        CK(cudaStreamWaitEvent(p->stream, p->evCopy[0], 0));
        CK(cudaMemcpyAsync(p->stA, p->stC, bytes, cudaMemcpyDeviceToDevice, p->stream));
        if (dl->nI) {
            const unsigned grid = (unsigned)((dl->nI + 255) / 256);
            if (dtype == SFFTB_F64) delta_scatter_kernel<double><<<grid, 256, 0, p->stream>>>((double*)p->stA, p->deltaIdx, (const double*)dval, dl->nI, (long long)npix);
            else delta_scatter_kernel<float><<<grid, 256, 0, p->stream>>>((float*)p->stA, p->deltaIdx, (const float*)dval, dl->nI, (long long)npix);
            CKL(p);
        }
        CK(cudaStreamWaitEvent(p->stream, p->evCopy[1], 0));
        CK(cudaMemcpyAsync(p->stB, p->stD, bytes, cudaMemcpyDeviceToDevice, p->stream));
        if (dl->nJ) {
            const unsigned grid = (unsigned)((dl->nJ + 255) / 256);
            if (dtype == SFFTB_F64) delta_scatter_kernel<double><<<grid, 256, 0, p->stream>>>((double*)p->stB, p->deltaIdx + dl->nI, (const double*)(dval + esz * dl->nI), dl->nJ, (long long)npix);
            else delta_scatter_kernel<float><<<grid, 256, 0, p->stream>>>((float*)p->stB, p->deltaIdx + dl->nI, (const float*)(dval + esz * dl->nI), dl->nJ, (long long)npix);
            CKL(p);
        }
    }
    if (dl) {
        p->pendI = nullptr; p->pendJ = nullptr;          // the compute stream already waits for both images
    } else {
        // evCopy[2..3] are re-recorded by the chunked D2H of the apply step, so the apply pair gets its own wait now
        p->pendI = p->evCopy[0]; p->pendJ = p->evCopy[1];
    }
    // the forward row pass of the apply pair runs on the copy stream (behind the copies of I and J, which were queued there
    // first) while the Cholesky runs on the compute stream, like in sfftb_gss
    const bool ovh = p->overlap && (p->row_v8 || p->row_h16 || p->row_g16 || (p->row_blu && p->row_blu <= 16)) && p->chol_coop && p->nsm >= 8;
    rc = f32 ? fit_device<float2>(p, p->stA, p->stB, dtype, nullptr, ovh ? p->stC : nullptr, ovh ? p->stD : nullptr)
             : fit_device<double2>(p, p->stA, p->stB, dtype, nullptr, ovh ? p->stC : nullptr, ovh ? p->stD : nullptr);
    if (rc) return rc;
    if (dl || ovh) { p->pendI = nullptr; p->pendJ = nullptr; } else { p->pendI = p->evCopy[2]; p->pendJ = p->evCopy[3]; }
    void* hd = p->row_fast ? diff : nullptr;
    p->defer_join = 1;
    rc = f32 ? apply_device<float2>(p, p->stC, p->stD, dtype, p->sol, p->stA, diff_dtype, nullptr, ovh, hd)
             : apply_device<double2>(p, p->stC, p->stD, dtype, p->sol, p->stA, diff_dtype, nullptr, ovh, hd);
    p->defer_join = 0;
    if (rc) return rc;
    if (!p->row_fast) {
        const size_t ob = npix * (diff_dtype == SFFTB_F64 ? 8 : 4);
        CK(cudaMemcpyAsync(diff, p->stA, ob, cudaMemcpyDeviceToHost, p->stream));
    }
    // completion = compute stream done AND the chunked copies of the difference image (copy stream) done; the compute stream
    // itself does not wait for those copies, so the next pair's kernels start while this image is still going home.  The
    // Solution goes home on the copy stream too: a device-to-host copy queued on the compute stream would sit in the copy
    // engine's queue behind the image chunks and stall the compute stream just the same.
    CK(cudaEventRecord(p->evFork, p->stream));
    CK(cudaStreamWaitEvent(p->stream2, p->evFork, 0));
    if (solution) CK(cudaMemcpyAsync(solution, p->sol, sizeof(double) * p->d.NEQ, cudaMemcpyDeviceToHost, p->stream2));
    CK(cudaEventRecord(p->evDone, p->stream2));
    p->pending = 1; p->pend_diff = diff; p->pend_sol = solution; p->pend_dtype = dtype; p->pend_diff_dtype = diff_dtype;
    p->pend_mode = 1;
    return 0;
}

extern "C" int sfftb_gss_submit(sfftb_plan* p, const void* I, const void* J, const void* mI, const void* mJ, int dtype,
                                double* solution, void* diff, int diff_dtype) {
    if (!p || !I || !J || !mI || !mJ || !diff) return fail(SFFTB_EINVAL, "null argument");
    return gss_submit_impl(p, I, J, mI, mJ, nullptr, SFFTB_MEM_HOST, dtype, solution, diff, diff_dtype);
}

extern "C" int sfftb_gss_submit_device(sfftb_plan* p, const void* I, const void* J, const void* mI, const void* mJ, int dtype,
                                       double* solution, void* diff, int diff_dtype) {
    if (!p || !I || !J || !mI || !mJ || !diff) return fail(SFFTB_EINVAL, "null argument");
    return gss_submit_impl(p, I, J, mI, mJ, nullptr, SFFTB_MEM_DEVICE, dtype, solution, diff, diff_dtype);
}

extern "C" int sfftb_gss_submit_delta(sfftb_plan* p, const void* I, const void* J, long long nI, const long long* idxI, const void* valI,
                                      long long nJ, const long long* idxJ, const void* valJ, int dtype, double* solution, void* diff,
                                      int diff_dtype) {
    if (!p || !I || !J || !diff) return fail(SFFTB_EINVAL, "null argument");
    if (nI < 0 || nJ < 0 || (nI && (!idxI || !valI)) || (nJ && (!idxJ || !valJ))) return fail(SFFTB_EINVAL, "bad delta arguments");
    DeltaSpec dl = {nI, nJ, idxI, idxJ, valI, valJ};
    return gss_submit_impl(p, I, J, nullptr, nullptr, &dl, SFFTB_MEM_HOST, dtype, solution, diff, diff_dtype);
}

extern "C" int sfftb_gss_template(sfftb_plan* p, const void* J, const void* mJ, int memkind, int dtype,
                                  double* solution, int sol_memkind, void* diff, int diff_memkind, int diff_dtype);

// Shared-template tile from HOST buffers, asynchronous (see sfftb_gss_submit).  Until the plan holds the cached Cholesky
// factor of the template (first tile) the call completes synchronously; afterwards the two H2D copies run on the copy
// stream, and sfftb_gss_finish waits for this plan's work only.
extern "C" int sfftb_gss_template_submit(sfftb_plan* p, const void* J, const void* mJ, int memkind, int dtype, double* solution, void* diff,
                                         int diff_dtype) {
    if (!p || !J || !mJ || !diff) return fail(SFFTB_EINVAL, "null argument");
    if (memkind != SFFTB_MEM_HOST && memkind != SFFTB_MEM_DEVICE) return fail(SFFTB_EINVAL, "bad memkind");
    if ((dtype != SFFTB_F64 && dtype != SFFTB_F32) || (diff_dtype != SFFTB_F64 && diff_dtype != SFFTB_F32)) return fail(SFFTB_EINVAL, "bad dtype");
    if (p->gen) return fail(SFFTB_EINVAL, "the shared-template path is not available for general-basis plans");
    if (!p->have_template) return fail(SFFTB_ESTATE, "no template has been prepared on this plan");
    if (p->pending) return fail(SFFTB_ESTATE, "sfftb_gss_template_submit: the previous submission of this plan has not been finished");
    CK(cudaSetDevice(p->device));
    p->pend_diff = diff; p->pend_sol = solution; p->pend_dtype = dtype; p->pend_diff_dtype = diff_dtype; p->pend_J = J; p->pend_mJ = mJ;
    p->pend_memkind = memkind;
    if (!p->factor_cached || !p->chol_coop) {
        int rc = sfftb_gss_template(p, J, mJ, memkind, dtype, solution, memkind, diff, memkind, diff_dtype);
        if (rc) return rc;
        p->pending = 1; p->pend_mode = 3;
        return 0;
    }
    if (memkind == SFFTB_MEM_DEVICE) {
        // images and outputs already on the device: nothing to copy, the tile is queued behind whatever the compute
        // stream holds (back-to-back tiles leave no launch gaps)
        const bool f32d = p->cfg.storage == SFFTB_STORE_F32;
        const void* tfitd = p->tstate;
        const void* tappd = (const char*)p->tstate + p->tstate_fit_bytes;
        p->pendI = nullptr; p->pendJ = nullptr;
        int rcd = f32d ? fit_device<float2>(p, nullptr, mJ, dtype, tfitd) : fit_device<double2>(p, nullptr, mJ, dtype, tfitd);
        if (rcd) return rcd;
        rcd = f32d ? apply_device<float2>(p, nullptr, J, dtype, p->sol, diff, diff_dtype, tappd)
                   : apply_device<double2>(p, nullptr, J, dtype, p->sol, diff, diff_dtype, tappd);
        if (rcd) return rcd;
        if (solution) CK(cudaMemcpyAsync(solution, p->sol, sizeof(double) * p->d.NEQ, cudaMemcpyDeviceToDevice, p->stream));
        CK(cudaEventRecord(p->evDone, p->stream));
        p->pending = 1; p->pend_mode = 2;
        return 0;
    }
    const bool f32 = p->cfg.storage == SFFTB_STORE_F32;
    const size_t bytes = (size_t)p->d.N0 * p->d.N1 * (dtype == SFFTB_F64 ? 8 : 4);
    if (!p->stC) { CK(cudaMalloc(&p->stC, sizeof(double) * (size_t)p->d.N0 * p->d.N1)); CK(cudaMalloc(&p->stD, sizeof(double) * (size_t)p->d.N0 * p->d.N1)); }
    CK(cudaMemcpyAsync(p->stB, mJ, bytes, cudaMemcpyHostToDevice, p->stream2));
    CK(cudaEventRecord(p->evCopy[0], p->stream2));
    CK(cudaMemcpyAsync(p->stD, J, bytes, cudaMemcpyHostToDevice, p->stream2));
    CK(cudaEventRecord(p->evCopy[1], p->stream2));
    const void* tfit = p->tstate;
    const void* tapp = (const char*)p->tstate + p->tstate_fit_bytes;
    p->pendI = nullptr; p->pendJ = p->evCopy[0];
    int rc = f32 ? fit_device<float2>(p, nullptr, p->stB, dtype, tfit) : fit_device<double2>(p, nullptr, p->stB, dtype, tfit);
    if (rc) return rc;
    p->pendI = nullptr; p->pendJ = p->evCopy[1];
    void* hd = p->row_fast ? diff : nullptr;
    p->defer_join = 1;
    rc = f32 ? apply_device<float2>(p, nullptr, p->stD, dtype, p->sol, p->stA, diff_dtype, tapp, false, hd)
             : apply_device<double2>(p, nullptr, p->stD, dtype, p->sol, p->stA, diff_dtype, tapp, false, hd);
    p->defer_join = 0;
    if (rc) return rc;
    if (!p->row_fast) {
        const size_t ob = (size_t)p->d.N0 * p->d.N1 * (diff_dtype == SFFTB_F64 ? 8 : 4);
        CK(cudaMemcpyAsync(diff, p->stA, ob, cudaMemcpyDeviceToHost, p->stream));
    }
    CK(cudaEventRecord(p->evFork, p->stream));
    CK(cudaStreamWaitEvent(p->stream2, p->evFork, 0));
    if (solution) CK(cudaMemcpyAsync(solution, p->sol, sizeof(double) * p->d.NEQ, cudaMemcpyDeviceToHost, p->stream2));
    CK(cudaEventRecord(p->evDone, p->stream2));
    p->pending = 1; p->pend_mode = 2;
    return 0;
}

extern "C" int sfftb_gss_finish(sfftb_plan* p) {
    if (!p) return fail(SFFTB_EINVAL, "null plan");
    if (!p->pending) return fail(SFFTB_ESTATE, "sfftb_gss_finish: nothing was submitted");
    CK(cudaSetDevice(p->device));
    p->pending = 0;
    if (p->pend_mode == 3) return 0;                       // completed inside the submit call
    CK(cudaEventSynchronize(p->evDone));
    int rc = check_solver(p);
    if (rc < 0) { if (p->pend_mode == 2) p->factor_cached = 0; return rc; }
    if (rc == 1 && p->pend_mode == 2) {
        // cannot happen with a cached factor (no factorisation ran); be safe: drop the cache and redo the tile synchronously
        p->factor_cached = 0;
        return sfftb_gss_template(p, p->pend_J, p->pend_mJ, p->pend_memkind, p->pend_dtype, p->pend_sol, p->pend_memkind,
                                  p->pend_diff, p->pend_memkind, p->pend_diff_dtype);
    }
    if (rc == 1) {
        // the Cholesky broke down and the LU fallback replaced the solution: apply again (rare; synchronous)
        const bool f32 = p->cfg.storage == SFFTB_STORE_F32;
        if (p->pend_mem == SFFTB_MEM_DEVICE) {
            rc = f32 ? apply_device<float2>(p, p->pend_I, p->pend_J, p->pend_dtype, p->sol, p->pend_diff, p->pend_diff_dtype)
                     : apply_device<double2>(p, p->pend_I, p->pend_J, p->pend_dtype, p->sol, p->pend_diff, p->pend_diff_dtype);
            if (rc) return rc;
            if (p->pend_sol) CK(cudaMemcpyAsync(p->pend_sol, p->sol, sizeof(double) * p->d.NEQ, cudaMemcpyDeviceToDevice, p->stream));
            CK(cudaStreamSynchronize(p->stream));
            return collect_timings(p, true, true);
        }
        rc = f32 ? apply_device<float2>(p, p->stC, p->stD, p->pend_dtype, p->sol, p->stA, p->pend_diff_dtype, nullptr, false, nullptr)
                 : apply_device<double2>(p, p->stC, p->stD, p->pend_dtype, p->sol, p->stA, p->pend_diff_dtype, nullptr, false, nullptr);
        if (rc) return rc;
        const size_t ob = (size_t)p->d.N0 * p->d.N1 * (p->pend_diff_dtype == SFFTB_F64 ? 8 : 4);
        CK(cudaMemcpyAsync(p->pend_diff, p->stA, ob, cudaMemcpyDeviceToHost, p->stream));
        if (p->pend_sol) CK(cudaMemcpyAsync(p->pend_sol, p->sol, sizeof(double) * p->d.NEQ, cudaMemcpyDeviceToHost, p->stream));
        CK(cudaStreamSynchronize(p->stream));
    }
    return collect_timings(p, true, true);
}

// ---- shared-template batch path (the reference re-transforms the template for every pair) ------------
static int template_alloc(sfftb_plan* p) {
    if (p->gen) return fail(SFFTB_EINVAL, "the shared-template path is not available for general-basis plans");
    if (p->tstate) return 0;
    const size_t csz = p->cfg.storage == SFFTB_STORE_F32 ? sizeof(float2) : sizeof(double2);
    const size_t nel = (size_t)(p->d.DK + 1) * (p->d.N1 / 2 + 1) * p->d.N0;
    p->tstate_fit_bytes = csz * nel;
    p->tstate_bytes = p->tstate_fit_bytes + sizeof(double2) * nel;
    CK(cudaMalloc(&p->tstate, p->tstate_bytes));
    return 0;
}

extern "C" int sfftb_template_prepare(sfftb_plan* p, const void* I, const void* mI, int memkind, int dtype) {
    if (!p || !I || !mI) return fail(SFFTB_EINVAL, "null argument");
    CK(cudaSetDevice(p->device));
    int rc;
    if ((rc = template_alloc(p))) return rc;
    const void *dI, *dmI;
    if ((rc = stage_in(p, mI, memkind, dtype, p->stA, &dmI))) return rc;
    if ((rc = stage_in(p, I, memkind, dtype, p->stB, &dI))) return rc;
    char* half = (char*)p->tstate + p->tstate_fit_bytes;
    if (p->cfg.storage == SFFTB_STORE_F32) {
        if (launch_row_fwd<float2>(p, dmI, dtype, (float2*)p->tstate, p->d.DK + 1)) return SFFTB_ECUDA;
    } else {
        if (launch_row_fwd<double2>(p, dmI, dtype, (double2*)p->tstate, p->d.DK + 1)) return SFFTB_ECUDA;
    }
    if (launch_row_fwd<double2>(p, dI, dtype, (double2*)half, p->d.DK + 1)) return SFFTB_ECUDA;
    CK(cudaStreamSynchronize(p->stream));
    p->have_template = 1;
    p->factor_cached = 0;
    p->tmpl_epoch++;
    return 0;
}

extern "C" int sfftb_template_state(sfftb_plan* p, void** dptr, size_t* bytes) {
    if (!p || !dptr || !bytes) return fail(SFFTB_EINVAL, "null argument");
    CK(cudaSetDevice(p->device));
    int rc;
    if ((rc = template_alloc(p))) return rc;
    *dptr = p->tstate; *bytes = p->tstate_bytes;
    return 0;
}

// Copy the complete shared-template state of `src` into `dst` (same configuration, same device): the template row spectra and, when
// `src` already holds them, the Cholesky factor of the template's normal equations, its lag rows and the cached segment spectra --
// device-to-device copies (about 0.3 GB at 2048^2, 0.1 ms) instead of one factorisation and one cache build per plan of a pipeline.
extern "C" int sfftb_template_clone(sfftb_plan* dst, sfftb_plan* src) {
    if (!dst || !src || dst == src) return fail(SFFTB_EINVAL, "sfftb_template_clone: two different plans are needed");
    if (dst->gen || src->gen) return fail(SFFTB_EINVAL, "the shared-template path is not available for general-basis plans");
    const sfftb_dims &a = dst->d, &b = src->d;
    if (dst->device != src->device || a.N0 != b.N0 || a.N1 != b.N1 || a.w0 != b.w0 || a.w1 != b.w1 || a.DK != b.DK || a.DB != b.DB ||
        dst->cfg.storage != src->cfg.storage || dst->nsolve != src->nsolve || dst->ld != src->ld || dst->fit_seg != src->fit_seg ||
        dst->chol_coop != src->chol_coop)
        return fail(SFFTB_EINVAL, "sfftb_template_clone: the two plans differ in configuration or device");
    if (!src->have_template || !src->tstate) return fail(SFFTB_ESTATE, "no template has been prepared on the source plan");
    if (dst->pending || src->pending) return fail(SFFTB_ESTATE, "sfftb_template_clone: a submission is still in flight");
    CK(cudaSetDevice(dst->device));
    int rc;
    if ((rc = template_alloc(dst))) return rc;
    CK(cudaStreamSynchronize(src->stream));
    cudaStream_t st = dst->stream;
    CK(cudaMemcpyAsync(dst->tstate, src->tstate, src->tstate_bytes, cudaMemcpyDeviceToDevice, st));
    dst->have_template = 1;
    dst->factor_cached = 0;
    dst->tmpl_epoch++;
    if (src->factor_cached && src->chol_coop && src->fit_seg == 2) {
        const int n = src->nsolve, nblk = (n + CC_NB - 1) / CC_NB;
        CK(cudaMemcpyAsync(dst->Aug, src->Aug, sizeof(double) * (size_t)(n + 1) * src->ld, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(dst->sc, src->sc, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(dst->cholW, src->cholW, sizeof(double) * (size_t)nblk * CC_NB * CC_NB, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(dst->kap2, src->kap2, sizeof(cd) * (size_t)(a.N1 / 2 + 1) * src->sfit.nrows, cudaMemcpyDeviceToDevice, st));
        dst->factor_cached = 1;
        if (src->aspec && src->aspec_epoch == src->tmpl_epoch && !dst->aspec_off && dst->grid_jc) {
            if (!dst->aspec || dst->aspec_elems < src->aspec_elems) {
                if (dst->aspec) { cudaFree(dst->aspec); dst->aspec = nullptr; }
                if (cudaMalloc(&dst->aspec, sizeof(cd) * src->aspec_elems) != cudaSuccess) { cudaGetLastError(); dst->aspec = nullptr; dst->aspec_elems = 0; }
                else dst->aspec_elems = src->aspec_elems;
            }
            if (dst->aspec) {
                CK(cudaMemcpyAsync(dst->aspec, src->aspec, sizeof(cd) * src->aspec_elems, cudaMemcpyDeviceToDevice, st));
                dst->aspec_epoch = dst->tmpl_epoch;
            }
        }
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int sfftb_template_mark_ready(sfftb_plan* p) {
    if (!p || !p->tstate) return fail(SFFTB_ESTATE, "template state was never allocated");
    p->have_template = 1;
    p->factor_cached = 0;
    p->tmpl_epoch++;
    return 0;
}

extern "C" int sfftb_gss_template(sfftb_plan* p, const void* J, const void* mJ, int memkind, int dtype,
                                  double* solution, int sol_memkind, void* diff, int diff_memkind, int diff_dtype) {
    if (!p || !diff || !J || !mJ) return fail(SFFTB_EINVAL, "null argument");
    if (!p->have_template) return fail(SFFTB_ESTATE, "no template has been prepared on this plan");
    if (diff_dtype != SFFTB_F64 && diff_dtype != SFFTB_F32) return fail(SFFTB_EINVAL, "bad diff dtype");
    CK(cudaSetDevice(p->device));
    const bool f32 = p->cfg.storage == SFFTB_STORE_F32;
    const void* tfit = p->tstate;
    const void* tapp = (const char*)p->tstate + p->tstate_fit_bytes;
    const void* dJ;
    int rc;
    if ((rc = stage_in(p, mJ, memkind, dtype, p->stB, &dJ))) return rc;
    rc = f32 ? fit_device<float2>(p, nullptr, dJ, dtype, tfit) : fit_device<double2>(p, nullptr, dJ, dtype, tfit);
    if (rc) return rc;
    void* ddiff = diff_memkind == SFFTB_MEM_DEVICE ? diff : p->stA;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if ((rc = stage_in(p, J, memkind, dtype, p->stB, &dJ))) return rc;
        rc = f32 ? apply_device<float2>(p, nullptr, dJ, dtype, p->sol, ddiff, diff_dtype, tapp)
                 : apply_device<double2>(p, nullptr, dJ, dtype, p->sol, ddiff, diff_dtype, tapp);
        if (rc) return rc;
        CK(cudaStreamSynchronize(p->stream));
        if (attempt == 1) break;
        rc = check_solver(p);
        if (rc < 0) { p->factor_cached = 0; return rc; }
        if (rc == 0) { p->factor_cached = (p->chol_coop && !env_int("SFFTB_NO_FACTOR_CACHE", 0)) ? 1 : 0; break; }
        p->factor_cached = 0;            // the LU fallback ran: Aug no longer holds a Cholesky factor
    }
    if ((rc = collect_timings(p, true, true))) return rc;
    if (diff_memkind == SFFTB_MEM_HOST) {
        const size_t bytes = (size_t)p->d.N0 * p->d.N1 * (diff_dtype == SFFTB_F64 ? 8 : 4);
        CK(cudaMemcpyAsync(diff, p->stA, bytes, cudaMemcpyDeviceToHost, p->stream));
    }
    if ((rc = copy_out(p, solution, sol_memkind, p->sol, sizeof(double) * p->d.NEQ))) return rc;
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

// ---- packet I/O edge on the device; all pointers are DEVICE pointers, work is queued on `stream` -----
static int fio_check(int device, int n1, int n2) {
    if (n1 <= 0 || n2 <= 0) return fail(SFFTB_EINVAL, "bad image shape");
    CK(cudaSetDevice(device));
    return 0;
}
extern "C" int sfftb_fits_decode(int device, void* stream, const void* raw, int bitpix, int naxis1, int naxis2, double bscale, double bzero,
                                 void* out, int out_dtype) {
    if (!raw || !out) return fail(SFFTB_EINVAL, "null argument");
    if (bitpix != 8 && bitpix != 16 && bitpix != 32 && bitpix != 64 && bitpix != -32 && bitpix != -64) return fail(SFFTB_EINVAL, "unsupported BITPIX %d", bitpix);
    int rc = fio_check(device, naxis1, naxis2);
    if (rc) return rc;
    dim3 blk(32, 8), grd((naxis1 + 31) / 32, (naxis2 + 31) / 32);
    if (out_dtype == SFFTB_F64) fits_decode_T_kernel<double><<<grd, blk, 0, (cudaStream_t)stream>>>((const unsigned char*)raw, bitpix, naxis1, naxis2, bscale, bzero, (double*)out);
    else if (out_dtype == SFFTB_F32) fits_decode_T_kernel<float><<<grd, blk, 0, (cudaStream_t)stream>>>((const unsigned char*)raw, bitpix, naxis1, naxis2, bscale, bzero, (float*)out);
    else return fail(SFFTB_EINVAL, "bad dtype");
    CK(cudaGetLastError());
    return 0;
}
extern "C" int sfftb_fits_encode(int device, void* stream, const void* img, int img_dtype, int naxis1, int naxis2, int bitpix, void* raw) {
    if (!raw || !img) return fail(SFFTB_EINVAL, "null argument");
    if (bitpix != -32 && bitpix != -64) return fail(SFFTB_EINVAL, "the difference image is written as BITPIX -32 or -64");
    int rc = fio_check(device, naxis1, naxis2);
    if (rc) return rc;
    dim3 blk(32, 8), grd((naxis1 + 31) / 32, (naxis2 + 31) / 32);
    if (img_dtype == SFFTB_F64) fits_encode_T_kernel<double><<<grd, blk, 0, (cudaStream_t)stream>>>((const double*)img, bitpix, naxis1, naxis2, (unsigned char*)raw);
    else if (img_dtype == SFFTB_F32) fits_encode_T_kernel<float><<<grd, blk, 0, (cudaStream_t)stream>>>((const float*)img, bitpix, naxis1, naxis2, (unsigned char*)raw);
    else return fail(SFFTB_EINVAL, "bad dtype");
    CK(cudaGetLastError());
    return 0;
}
// flags[0] = a NaN was found in the unmasked pair, flags[1] = a NaN in the masked pair (the packets assert there is none)
extern "C" int sfftb_nan_union_fill(int device, void* stream, void* A, void* B, const void* mA, const void* mB, int dtype, size_t n,
                                    unsigned char* mask, int* flags) {
    if (!A || !B || !mA || !mB || !mask || !flags) return fail(SFFTB_EINVAL, "null argument");
    CK(cudaSetDevice(device));
    CK(cudaMemsetAsync(flags, 0, 2 * sizeof(int), (cudaStream_t)stream));
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (dtype == SFFTB_F64) nan_union_fill_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double*)A, (double*)B, (const double*)mA, (const double*)mB, n, mask, flags);
    else if (dtype == SFFTB_F32) nan_union_fill_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)A, (float*)B, (const float*)mA, (const float*)mB, n, mask, flags);
    else return fail(SFFTB_EINVAL, "bad dtype");
    CK(cudaGetLastError());
    return 0;
}
extern "C" int sfftb_nan_mask_apply(int device, void* stream, void* D, int dtype, const unsigned char* mask, size_t n, double sign) {
    if (!D) return fail(SFFTB_EINVAL, "null argument");
    CK(cudaSetDevice(device));
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (dtype == SFFTB_F64) nan_mask_apply_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double*)D, mask, n, sign);
    else if (dtype == SFFTB_F32) nan_mask_apply_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)D, mask, n, sign);
    else return fail(SFFTB_EINVAL, "bad dtype");
    CK(cudaGetLastError());
    return 0;
}

// Kernel regularisation of sfft/BSplineSFFT.py:3570-3700: LHMAT += LAMBDA * REGMAT with the Kronecker structure
// REGMAT[(k,c),(k',c')] = SCALE^2 * SST[k,k'] * iREG[c,c'] (fill_regmat, :2091-2119).  The two small factors are kept on
// the device and added inside the matrix fill.  SST == NULL switches it off.
extern "C" int sfftb_set_regularizer_varying(sfftb_plan* p, const double* CSST, const double* DSST);
extern "C" int sfftb_set_regularizer(sfftb_plan* p, const double* SST, const double* iREG, double lambda) {
    if (!p) return fail(SFFTB_EINVAL, "null plan");
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    p->factor_cached = 0;
    p->fill.CSST = nullptr; p->fill.DSST = nullptr;
    if (!SST || !iREG) { p->fill.SST = nullptr; p->fill.iREG = nullptr; p->fill.regw = 0.0; if (p->gen) gen_set_regularizer(p); return 0; }
    if (!(lambda >= 0.0)) return fail(SFFTB_EINVAL, "LAMBDA_REGULARIZE must be >= 0");
    const size_t nS = (size_t)p->d.Fij * p->d.Fij, nI = (size_t)p->d.Fab * p->d.Fab;
    if (!p->regSST) { CK(cudaMalloc(&p->regSST, sizeof(double) * nS)); CK(cudaMalloc(&p->regI, sizeof(double) * nI)); }
    CK(cudaMemcpy(p->regSST, SST, sizeof(double) * nS, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->regI, iREG, sizeof(double) * nI, cudaMemcpyHostToDevice));
    const double N = (double)p->d.N0 * (double)p->d.N1;
    p->fill.SST = p->regSST; p->fill.iREG = p->regI; p->fill.regw = lambda / (N * N);
    if (p->gen) gen_set_regularizer(p);
    return 0;
}

// SEPARATE-VARYING plans: the Gram matrices that replace SST where a centre tap is involved (fill_regmat, :2122-2166).
// Call after sfftb_set_regularizer; (Fij x Fij) host arrays, rows / columns beyond ScaFij zero (the placeholder basis).
extern "C" int sfftb_set_regularizer_varying(sfftb_plan* p, const double* CSST, const double* DSST) {
    if (!p || !CSST || !DSST) return fail(SFFTB_EINVAL, "null argument");
    if (!p->sca_n && !p->gen) return fail(SFFTB_ESTATE, "the plan was not created with a varying scaling (sca_degree > 0)");
    if (!p->fill.SST) return fail(SFFTB_ESTATE, "call sfftb_set_regularizer first");
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    p->factor_cached = 0;
    const size_t nS = (size_t)p->d.Fij * p->d.Fij;
    if (!p->regC) { CK(cudaMalloc(&p->regC, sizeof(double) * nS)); CK(cudaMalloc(&p->regD, sizeof(double) * nS)); }
    CK(cudaMemcpy(p->regC, CSST, sizeof(double) * nS, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->regD, DSST, sizeof(double) * nS, cudaMemcpyHostToDevice));
    p->fill.CSST = p->regC; p->fill.DSST = p->regD;
    if (p->gen) gen_set_regularizer(p);
    return 0;
}

// Realize_MatchingKernel / Realize_FluxScaling (sfft/utils/SFFTSolutionReader.py:116-196) on the device.
extern "C" int sfftb_realize(sfftb_plan* p, const double* solution, int sol_memkind, const double* xy, int xy_memkind, int nq,
                             double* kerstack, double* fscal, int out_memkind) {
    if (!p || !solution || !xy) return fail(SFFTB_EINVAL, "null argument");
    if (p->gen) return fail(SFFTB_EINVAL, "general-basis plans realise kernels through sfftb_realize_general");
    if (nq <= 0) return fail(SFFTB_EINVAL, "no coordinates requested");
    if (!kerstack && !fscal) return 0;
    CK(cudaSetDevice(p->device));
    const int Fab = p->d.Fab;
    double *dsol = nullptr, *dxy = nullptr, *dks = nullptr, *dfs = nullptr;
    std::vector<void*> tmp;
    auto cleanup = [&]() { for (void* q : tmp) cudaFree(q); };
    auto dev_in = [&](const double* src, int kind, size_t n, double** out) -> int {
        if (kind == SFFTB_MEM_DEVICE) { *out = const_cast<double*>(src); return 0; }
        CK(cudaMalloc(out, sizeof(double) * n));
        tmp.push_back(*out);
        CK(cudaMemcpyAsync(*out, src, sizeof(double) * n, cudaMemcpyHostToDevice, p->stream));
        return 0;
    };
    int rc;
    if ((rc = dev_in(solution, sol_memkind, p->d.NEQ, &dsol)) || (rc = dev_in(xy, xy_memkind, 2 * (size_t)nq, &dxy))) { cleanup(); return rc; }
    if (out_memkind == SFFTB_MEM_DEVICE) { dks = kerstack; dfs = fscal; }
    else {
        if (kerstack) { if (cudaMalloc(&dks, sizeof(double) * (size_t)nq * Fab) != cudaSuccess) { cleanup(); return fail(SFFTB_ECUDA, "cudaMalloc failed"); } tmp.push_back(dks); }
        if (fscal) { if (cudaMalloc(&dfs, sizeof(double) * (size_t)nq) != cudaSuccess) { cleanup(); return fail(SFFTB_ECUDA, "cudaMalloc failed"); } tmp.push_back(dfs); }
    }
    ReaderArgs ra;
    ra.sol = dsol; ra.xy = dxy; ra.nq = nq; ra.N0 = p->d.N0; ra.N1 = p->d.N1; ra.L0 = p->d.L0; ra.L1 = p->d.L1;
    ra.DK = p->d.DK; ra.Fij = p->d.Fij; ra.Fab = Fab; ra.kerstack = dks; ra.fscal = dfs;
    realize_kernel<<<nq, 256, 0, p->stream>>>(ra);
    p->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && out_memkind == SFFTB_MEM_HOST) {
        if (kerstack) e = cudaMemcpyAsync(kerstack, dks, sizeof(double) * (size_t)nq * Fab, cudaMemcpyDeviceToHost, p->stream);
        if (e == cudaSuccess && fscal) e = cudaMemcpyAsync(fscal, dfs, sizeof(double) * (size_t)nq, cudaMemcpyDeviceToHost, p->stream);
    }
    if (e == cudaSuccess && !tmp.empty()) e = cudaStreamSynchronize(p->stream);
    cleanup();
    if (e != cudaSuccess) return fail(SFFTB_ECUDA, "sfftb_realize: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int sfftb_export_solved_system(sfftb_plan* p, double* L, double* b) {
    if (!p) return fail(SFFTB_EINVAL, "null plan");
    if (!p->have_fit) return fail(SFFTB_ESTATE, "no fit has been run on this plan");
    CK(cudaSetDevice(p->device));
    const int n = p->nsolve, ld = n + 1;
    double* buf = nullptr;
    CK(cudaMalloc(&buf, sizeof(double) * (size_t)(n + 1) * ld));
    int rc = 0;
    if (p->gen) rc = gen_export(p, buf);
    else {
        dim3 blk(32, 8), grd((n + 1 + 31) / 32, (n + 1 + 7) / 8);
        fill_matrix_kernel<<<grd, blk, 0, p->stream>>>(p->fill, p->idxmap, n, nullptr, buf, ld, p->info + 3);
        p->launches++;
    }
    cudaError_t e = rc ? cudaErrorUnknown : cudaGetLastError();
    if (e == cudaSuccess && L) e = cudaMemcpy2DAsync(L, sizeof(double) * n, buf, sizeof(double) * ld, sizeof(double) * n, n, cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess && b) e = cudaMemcpyAsync(b, buf + (size_t)n * ld, sizeof(double) * n, cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(SFFTB_ECUDA, "sfftb_export_solved_system: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int sfftb_export_normal_eq(sfftb_plan* p, double* LHMAT, double* RHb) {
    if (!p) return fail(SFFTB_EINVAL, "null plan");
    if (p->gen) return fail(SFFTB_EINVAL, "general-basis plans export the solved system (sfftb_export_solved_system)");
    if (!p->have_fit) return fail(SFFTB_ESTATE, "no fit has been run on this plan");
    CK(cudaSetDevice(p->device));
    const int n = p->d.NEQ, ld = n + 1;
    if (!p->exportbuf) CK(cudaMalloc(&p->exportbuf, sizeof(double) * (size_t)(n + 1) * ld));
    dim3 blk(32, 8), grd((n + 1 + 31) / 32, (n + 1 + 7) / 8);
    fill_matrix_kernel<<<grd, blk, 0, p->stream>>>(p->fill, p->ident, n, nullptr, p->exportbuf, ld, p->info + 3);
    CKL(p);
    if (LHMAT) CK(cudaMemcpy2DAsync(LHMAT, sizeof(double) * n, p->exportbuf, sizeof(double) * ld, sizeof(double) * n, n, cudaMemcpyDeviceToHost, p->stream));
    if (RHb) CK(cudaMemcpyAsync(RHb, p->exportbuf + (size_t)n * ld, sizeof(double) * n, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" int sfftb_dbg_lag_tables(sfftb_plan* p, double* R, double* RJ, double* RT, double* RJT) {
    if (!p) return fail(SFFTB_EINVAL, "null plan");
    if (p->gen) return fail(SFFTB_EINVAL, "not available for general-basis plans");
    if (!p->have_fit) return fail(SFFTB_ESTATE, "no fit has been run on this plan");
    CK(cudaSetDevice(p->device));
    const sfftb_dims& d = p->d;
    const size_t nl0 = 4 * d.w0 + 1, nl1 = 4 * d.w1 + 1, nlj0 = 2 * d.w0 + 1, nlj1 = 2 * d.w1 + 1;
    CK(cudaStreamSynchronize(p->stream));
    if (R) CK(cudaMemcpy(R, p->R, sizeof(double) * p->cfit.npairs * nl0 * nl1, cudaMemcpyDeviceToHost));
    if (RJ) CK(cudaMemcpy(RJ, p->RJ, sizeof(double) * d.Fij * nlj0 * nlj1, cudaMemcpyDeviceToHost));
    if (RT) CK(cudaMemcpy(RT, p->RT, sizeof(double) * d.Fij * d.Fpq * nlj0 * nlj1, cudaMemcpyDeviceToHost));
    if (RJT) CK(cudaMemcpy(RJT, p->RJT, sizeof(double) * d.Fpq, cudaMemcpyDeviceToHost));
    return 0;
}

template <typename TSt>
__global__ void widen_kernel(const TSt* __restrict__ in, cd* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = load_c(in + i);
}

extern "C" int sfftb_dbg_row_spectra(sfftb_plan* p, int which, double* out) {
    if (!p || !out) return fail(SFFTB_EINVAL, "null argument");
    if (p->gen) return fail(SFFTB_EINVAL, "not available for general-basis plans");
    CK(cudaSetDevice(p->device));
    const sfftb_dims& d = p->d;
    const size_t n = (size_t)(which == 0 ? d.DK + 1 : 1) * (d.N1 / 2 + 1) * d.N0;
    cd* tmp = nullptr;
    CK(cudaMalloc(&tmp, sizeof(cd) * n));
    const void* src = which == 0 ? p->gI : p->gJ;
    if (p->cfg.storage == SFFTB_STORE_F32) widen_kernel<float2><<<1024, 256, 0, p->stream>>>((const float2*)src, tmp, n);
    else widen_kernel<double2><<<1024, 256, 0, p->stream>>>((const double2*)src, tmp, n);
    cudaError_t e = cudaMemcpyAsync(out, tmp, sizeof(cd) * n, cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(SFFTB_ECUDA, "CUDA error %s in dbg_row_spectra", cudaGetErrorString(e));
    return 0;
}

extern "C" int sfftb_dbg_fft1d(int device, int n, int nbatch, int sign, const double* in, double* out) {
    if (!in || !out || n < 1 || nbatch < 1) return fail(SFFTB_EINVAL, "bad argument");
    CK(cudaSetDevice(device));
    FftDesc fd;
    if (!make_fft_desc(n, &fd) || !fft_fits_threads(fd, 512)) return fail(SFFTB_EINVAL, "unsupported FFT length %d", n);
    if (init_generic_radix_tables()) return SFFTB_ECUDA;
    int v = 0;
    CK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const int pitch = n + 1;
    int ppc = 3;
    while (ppc > 1 && sizeof(cd) * (size_t)ppc * pitch > (size_t)v) --ppc;
    const size_t smem = sizeof(cd) * (size_t)ppc * pitch;
    if (smem > (size_t)v) return fail(SFFTB_EINVAL, "FFT length %d does not fit shared memory", n);
    cd *tw = nullptr, *din = nullptr, *dout = nullptr;
    int rc = upload_twiddles(n, &tw);
    if (rc) return rc;
    const size_t bytes = sizeof(cd) * (size_t)n * nbatch;
    cudaError_t e = cudaMalloc(&din, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&dout, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(din, in, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && set_smem(dbg_fft_kernel, smem)) e = cudaErrorInvalidValue;
    if (e == cudaSuccess) {
        dbg_fft_kernel<<<(nbatch + ppc - 1) / ppc, 512, smem>>>(fd, tw, din, dout, nbatch, ppc, pitch, sign < 0 ? -1.0 : 1.0);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost);
    cudaFree(tw); cudaFree(din); cudaFree(dout);
    if (e != cudaSuccess) return fail(SFFTB_ECUDA, "CUDA error %s in dbg_fft1d", cudaGetErrorString(e));
    return 0;
}

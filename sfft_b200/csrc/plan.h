// plan.h -- the plan structure and the helpers shared by the translation units of libsfft_b200.so.
//
// The library is built from several .cu files (one per kernel family) so that they compile in parallel; every file
// includes this header.  Kernel TEMPLATES are instantiated only where they are launched; the non-template kernels are
// compiled in exactly one translation unit (guards SFFTB_TU_*).
#pragma once
#include "../../include/sfft_b200.h"
#include "common.cuh"
#include "fft_smem.cuh"
#include "kernels_row.cuh"
#include "kernels_fit.cuh"
#include "kernels_solve.cuh"
#include "kernels_chol.cuh"
#include "kernels_apply.cuh"
#include "kernels_row_fast.cuh"
#include "kernels_row_v8.cuh"
#include "kernels_row_h16.cuh"
#include "kernels_row_g16.cuh"
#include "kernels_row_blu.cuh"
#include "kernels_fit_seg.cuh"
#include "kernels_fit_seg3.cuh"
#include "kernels_fit_seg4.cuh"
#include "kernels_fit_jcache.cuh"
#include "kernels_reader.cuh"
#include "kernels_fitsio.cuh"
#include "kernels_gen.cuh"

#include <math.h>
#include <stdarg.h>
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

// thread-local error message (sfft_b200.cu); returns `code`
int sfftb_fail(int code, const char* fmt, ...);
#define fail sfftb_fail

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(SFFTB_ECUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
    } while (0)

#define CKL(p)                                                                                           \
    do {                                                                                                 \
        (p)->launches++;                                                                                 \
        cudaError_t e_ = cudaGetLastError();                                                             \
        if (e_ != cudaSuccess)                                                                           \
            return fail(SFFTB_ECUDA, "kernel launch failed: %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

enum { EV_START = 0, EV_ROWS, EV_COL, EV_RED, EV_SOLVE, EV_A0, EV_AROWS, EV_ACOL, EV_AINV, EV_KFIT0, EV_KFIT1, EV_COUNT };

struct sfftb_plan {
    sfftb_config cfg;
    sfftb_dims d;
    int device;
    int nsm;
    size_t max_smem;
    cudaStream_t stream, own_stream;
    cudaStream_t stream2;        // side stream: forward row pass of the apply step overlapped with the solve
    cudaEvent_t evFork, evJoin;
    int overlap;                 // 1: sfftb_gss overlaps the apply row pass with the Cholesky solve (device images)
    int row_grid_limit;          // > 0: cap of the persistent row-kernel grid (half the SMs while overlapped)
    int chol_grid_limit;
    int solver_sms;              // > 0: SM partition (sfftb_plan_set_partition): the Cholesky runs on this many SMs, every persistent
                                 // throughput kernel on the others, so pairs in flight on different streams overlap
    // tables
    cd *tw0, *tw1, *twMf, *twH, *Q;
    cd *tabA, *tabB_row, *tabC_row;   // register-engine twiddle tables
    cd* tabC_row32;                   // exp(-2 pi i r k / 8192), [(r-1) 256 + k]: pass-A twiddles of the 32 x 256 row pass
    cd *vt8_8, *vt64_8, *vt64_4, *vt256_4, *vt512_4;   // 8-values-per-thread engine tables
    RowV8Args rowv;
    VTabs vtabs;
    size_t smem_sfit3;
    int row_v8;                  // 0 or the engine length H
    int row_h16;                 // 0 or H: R x 256 forward row pass on the half-warp engine (kernels_row_h16.cuh)
    RowH16Args rowh;
    int row_blu;                 // 0 or R: chirp-z row passes on the half-warp engine for any other even N1 <= 4096 (kernels_row_blu.cuh)
    RowBluArgs rowb;
    cd *bluC16, *bluB16, *bluBp16, *bluTwP16;
    int row_g16;                 // 0 or R: N1 = 512 R with R in {3, 5, 6, 10, 12} (kernels_row_g16.cuh), forward and inverse
    size_t smem_rowv;
    double* PHI;
    int *idxmap, *ident;
    // workspaces
    void *gI, *gJ;               // transposed row spectra of the FIT step (storage type)
    void *gIa, *gJa;             // row spectra of the APPLY step, always fp64 (alias gI / gJ for fp64 storage); gJa doubles as the
                                 // FDIFF column buffer.  fp32 storage rounds only the spectra the normal equations are built
                                 // from: rounding the spectra of the subtraction itself is amplified by the kernel gain
                                 // (2e-5 on a deconvolution-direction pair), see DESIGN.md
    void *stA, *stB;             // device staging for host images / host diff
    void *stC, *stD;             // second staging pair (host GSS: the apply images are copied while the fit computes)
    cudaEvent_t evCopy[4], evStart;
    cudaEvent_t evDone;          // end of the work queued by sfftb_gss_submit
    int pending;                 // a submitted GSS has not been finished yet
    int defer_join;              // asynchronous submissions: the chunked D2H of the difference image is not joined into the compute stream
    void* pend_diff; double* pend_sol; int pend_dtype, pend_diff_dtype;
    int pend_mode;               // 1 = pair (sfftb_gss_submit), 2 = shared-template tile, 3 = already completed synchronously
    const void *pend_J, *pend_mJ; int pend_memkind;
    const void* pend_I; int pend_mem;            // pair submissions: memory kind of the images (device pairs are re-applied in place)
    long long* deltaIdx; void* deltaVal; size_t delta_cap;   // sfftb_gss_submit_delta: staged sparse deltas of the masked pair
    cudaEvent_t pendI, pendJ;    // events the next row pass of I / J has to wait for (host pipeline), or NULL
    cd *kap, *lam, *nuJ;
    double *R, *RJ, *RT, *RJT;
    double *Aug, *sc, *diagU, *sol;
    double *cholW, *cholY, *cholX;   // cooperative Cholesky: inverse diagonal blocks, back-substitution vectors
    unsigned* cholBar;
    unsigned* substFlags;        // 2 * nblk epoch tags of the dataflow substitution kernel
    unsigned substEpoch;
    int subst_ok;
    int sca_n;                   // SEPARATE-VARYING polynomial scaling: number of scaling basis functions (0 = off)
    double* solEff;              // solution with the centre taps moved onto the planes of the scaling basis (apply step)
    double *regSST, *regI, *regC, *regD;       // kernel regulariser factors (sfftb_set_regularizer)
    cd *bluTw, *bluC, *bluB;     // Bluestein tables of the generic row pass (row lengths with a prime factor > 13)
    ulonglong2* substMsg;        // 2 * nblk * 64 {value | epoch} messages of chol_subst2_kernel
    int chol_coop;
    double* exportbuf;
    int ld, nsolve;
    int* info;                   // device: [0] cholesky pivot, [1] non-finite, [2] lu pivot
    int* info_h;                 // pinned
    // kernel arguments
    ColArgs cfit;
    FirArgs fir;
    RowArgs row;
    RowInvArgs rinv;
    RowFastArgs rowf;
    RowInvFastArgs rinvf;
    int row_fast;                // 0 or the engine length H
    // shared-template tiles: A-role segment spectra of the template cached in HBM (kernels_fit_jcache.cuh)
    cd* aspec; size_t aspec_elems;   // [NH][nseg][Fij][256]
    long long aspec_epoch;           // template epoch the cache was built from (-1 = none)
    long long tmpl_epoch;            // bumped whenever the template state changes
    int aspec_off;                   // 1: allocation failed or switched off -> JONLY instantiation of fit_seg4_kernel
    int grid_jc;                     // CTAs per SM of fit_jonly_cached_kernel
    // segmented fit path
    SegFitArgs sfit;
    int fit_seg;                 // 2: fit_seg3_kernel + lag_reduce2 path, 0: folded-slice generic kernel
    int fit_generic_ok;          // the folded-slice kernel has a valid geometry for this shape
    int grid_sfit;
    cd* kap2;                    // [NH][nrows]
    cd* momg;                    // [NH][5 * SFFTB_MAXE] column moments of the stored planes
    double* part;                // [ksplit][nrows][4 w1 + 1]
    LagReduce2Args red2;
    LagFinishArgs fin2;
    ReduceArgs red;
    PolyReduceArgs pred;
    FillArgs fill;
    size_t smem_fit, smem_fir, smem_row, smem_fir3;
    cd* firTaps; double* firCA;
    void* tstate;                // cached template row spectra: [fit: mI planes | apply: I planes], storage type
    size_t tstate_bytes, tstate_fit_bytes;   // [fit half: mI planes, storage type | apply half: I planes, fp64]
    int have_template;
    int factor_cached;           // template path: Aug / cholW hold the Cholesky factor of the (tile independent) LHMAT
    int resolves;                // number of solves served from the cached factor (diagnostics)
    int grid_fit;
    int nrowsK, nrowsL;
    // state
    cudaEvent_t ev[EV_COUNT];
    int timing;
    float ms[8];                 // [7]: the fit column kernel alone (the launches between EV_KFIT0 and EV_KFIT1)
    long long launches;
    int last_solver;
    int have_fit;
    void* gen;                   // general-basis plans (tu_gen.cu): GenState*, NULL for the polynomial sfftcore plans
    const double* vtab;          // general-basis plans: column tables multiplied into the row pass (device, [nVs][N1])
};

static inline int init_generic_radix_tables() {
    const int rad[4] = {5, 7, 11, 13};
    double2 h[4][16];
    memset(h, 0, sizeof h);
    const long double tp = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < 4; ++k)
        for (int q = 0; q < rad[k]; ++q) {
            const long double ang = tp * q / rad[k];
            h[k][q].x = (double)cosl(ang);
            h[k][q].y = (double)(-sinl(ang));
        }
    CK(cudaMemcpyToSymbol(c_wgen, h, sizeof h));
    return 0;
}

// The dynamic shared-memory limit of a kernel is a property of the FUNCTION (per device), not of a plan: plans of
// different shapes live side by side (and are created from different host threads), so the limit is only ever raised.
template <typename T>
static int set_smem(T kernel, size_t bytes) {
    static std::mutex mtx;
    static std::map<std::pair<int, const void*>, size_t> cur;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mtx);
    size_t& c = cur[std::make_pair(dev, (const void*)kernel)];
    if (bytes > c) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        c = bytes;
    }
    return 0;
}

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// SMs the persistent throughput kernels (row passes, fit column pass) may occupy
static inline int work_sms(const sfftb_plan* p) { return p->solver_sms > 0 ? std::max(1, p->nsm - p->solver_sms) : p->nsm; }

#define EVREC(p, k) do { if ((p)->timing) CK(cudaEventRecord((p)->ev[k], (p)->stream)); } while (0)

// ---- family entry points (one translation unit each) -----------------------------------------------------------------
// tu_rows.cu
int rows_setup(sfftb_plan* p);
template <typename TSt> int launch_row_fwd(sfftb_plan* p, const void* img, int dtype, TSt* out, int nj, const double* vtab = nullptr);
int launch_row_inv(sfftb_plan* p, const double* bpq, void* ddiff, int diff_dtype, void* hdiff);
// tu_fit.cu
int fit_setup(sfftb_plan* p);
template <typename TSt> int launch_fit_cols(sfftb_plan* p, const TSt* gIsrc, bool jonly);
// tu_chol.cu
int chol_setup(sfftb_plan* p);
int run_cholesky(sfftb_plan* p, int resolve = 0);
int run_lu(sfftb_plan* p);
// tu_apply.cu
int apply_setup(sfftb_plan* p);
int launch_fir(sfftb_plan* p, const double2* gIsrc, const double* dsol);
// sfft_b200.cu
int upload_twiddles(int n, cd** out);
int upload_engine_table(int Ns, int R, cd** out);
int upload_bluestein(int H, int M, cd** outC, cd** outB);
int plan_init_common(sfftb_plan* p, const sfftb_config* cfg);
// tu_fit.cu (shared with the general-basis path)
int lag_reduce2_setup();
int launch_lag_reduce2(sfftb_plan* p, const LagReduce2Args& a, const cd* kap, double* part);
// tu_gen.cu
int gen_plan_create(sfftb_plan* p, const sfftb_config* cfg, const sfftb_basis* ker, const sfftb_basis* sca, const sfftb_basis* bkg, int mode);
void gen_free(sfftb_plan* p);
int gen_nvs(const sfftb_plan* p);
void gen_info(const sfftb_plan* p, int* out);
void* gen_planes(const sfftb_plan* p);
int gen_set_regularizer(sfftb_plan* p);
int gen_rjt(sfftb_plan* p, const void* dJ, int dtype);
template <typename TSt> int gen_fit_cols(sfftb_plan* p);
int gen_fill_system(sfftb_plan* p);
int gen_restore(sfftb_plan* p);
int gen_export(sfftb_plan* p, double* buf);
int gen_fir(sfftb_plan* p, const double* dsol);
void* gen_planes_apply(const sfftb_plan* p);
int gen_bkg_subtract(sfftb_plan* p, const double* bf, void* ddiff, int diff_dtype);

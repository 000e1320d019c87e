/*
 * sfft_b200.h -- C ABI of libsfft_b200.so, the B200 (sm_100a) SFFT subtraction core.
 *
 * The reference (thomasvrussell/sfft @ fa820e8) has no FFI for this path: its boundary is
 * Python and the backend is picked by the string BACKEND_4SUBTRACT in {'Cupy','Numpy'}
 * (sfft/sfftcore/SFFTConfigure.py:1387-1395, sfft/sfftcore/SFFTSubtract.py:828-835).
 * These entry points are what a third backend branch in those two dispatchers binds with
 * ctypes.  Plain pointers and sizes only; no torch / cupy types.
 *
 * Conventions
 *   - images are C-contiguous (N0, N1) arrays, N1 fastest, exactly the arrays the reference
 *     hands to ElementalSFFTSubtract.ESS (i.e. FITS data after the `.T`,
 *     sfft/CustomizedPacket.py:93-96);
 *   - `memkind` says where a caller buffer lives, `dtype` what it holds;
 *   - every function returns 0 on success and a negative SFFTB_E* code on failure;
 *     sfftb_last_error() returns a thread-local message the Python shim re-raises as
 *     Exception('MeLOn ERROR: ...') like the reference does;
 *   - a plan is bound to one device and one stream and is not re-entrant; different plans
 *     may be driven concurrently from different host threads (the one-thread-per-GPU model
 *     of sfft/MultiEasySparsePacket.py:510-514, 937-940).  Every entry point sets the
 *     plan's device itself.
 */
#ifndef SFFT_B200_H
#define SFFT_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFFTB_VERSION 100

#define SFFTB_OK            0
#define SFFTB_EINVAL       -1   /* bad argument / unsupported configuration */
#define SFFTB_ECUDA        -2   /* CUDA runtime error */
#define SFFTB_ESINGULAR    -3   /* normal matrix not factorisable (numpy.linalg.LinAlgError analogue) */
#define SFFTB_ENOTFINITE   -4   /* non-finite value reached the solver (lu_factor check_finite analogue,
                                   sfft/sfftcore/SFFTSubtract.py:15-23) */
#define SFFTB_ESTATE       -5   /* call sequence error (e.g. export before any fit) */

#define SFFTB_MEM_HOST      0
#define SFFTB_MEM_DEVICE    1

#define SFFTB_F64           0
#define SFFTB_F32           1

/* storage precision of the intermediate spectra in HBM (all arithmetic is fp64 either way) */
#define SFFTB_STORE_F64     0
#define SFFTB_STORE_F32     1

typedef struct sfftb_plan sfftb_plan;

/* Plays the role of the arguments of SingleSFFTConfigure.SSC
 * (sfft/sfftcore/SFFTConfigure.py:1371-1395: NX, NY, KerHW, KerPolyOrder, BGPolyOrder, ConstPhotRatio). */
typedef struct sfftb_config {
    int device;            /* CUDA ordinal (CUDA_DEVICE_4SUBTRACT, sfft/CustomizedPacket.py:128-131) */
    int N0, N1;            /* image shape (NX, NY) */
    int w0, w1;            /* kernel half widths (KerHW on both axes in sfftcore) */
    int DK, DB;            /* KerPolyOrder, BGPolyOrder: 0..3 (SFFTConfigure.py:19-28) */
    int const_phot_ratio;  /* ConstPhotRatio */
    int storage;           /* SFFTB_STORE_F64 | SFFTB_STORE_F32 */
    int fold;              /* column-pass fold factor V (0 = choose automatically) */
    int sca_degree;        /* with const_phot_ratio: 0 = one constant photometric scaling (SEPARATE-CONSTANT of
                              sfft/BSplineSFFT.py:77-86; sfftcore's ConstPhotRatio=True), DS > 0 = the scaling varies as a
                              polynomial of degree DS <= DK (SEPARATE-VARYING, :176-190, :3733-3747) */
    int reserved[5];       /* must be zero */
} sfftb_config;

/* Sizes derived exactly as SFFTParam_dict (SFFTConfigure.py:35-75). */
typedef struct sfftb_dims {
    int N0, N1, w0, w1, DK, DB, L0, L1, Fab, Fij, Fpq, Fijab, NEQ, NEQ_FSfree;
    int fold, sub_len;     /* V and M = N0 / V actually used by the column pass */
} sfftb_dims;

int  sfftb_version(void);
const char* sfftb_last_error(void);

/* SSC: build the plan (tables, workspaces).  Replaces SingleSFFTConfigure.SSC's kernel JIT
 * (SFFTConfigure.py:9-817). */
int  sfftb_plan_create(sfftb_plan** out, const sfftb_config* cfg);
int  sfftb_plan_destroy(sfftb_plan* plan);
int  sfftb_plan_dims(const sfftb_plan* plan, sfftb_dims* out);
/* Run all work of this plan on `cuda_stream` (a cudaStream_t; NULL = the plan's own stream, which is created
 * cudaStreamNonBlocking and therefore does NOT synchronise with the legacy default stream: a caller that produces its
 * inputs on the default stream passes cudaStreamLegacy, (cudaStream_t)0x1, or cudaStreamPerThread explicitly). */
int  sfftb_plan_set_stream(sfftb_plan* plan, void* cuda_stream);
int  sfftb_plan_sync(sfftb_plan* plan);
/* SM partition for several pairs in flight on one GPU (plans on different streams; the reference keeps a queue of pairs per
 * GPU, sfft/MultiEasySparsePacket.py:568-649): the latency-bound Cholesky of LSSolver (SFFTSubtract.py:15-23) is launched on
 * `solver_sms` SMs and every persistent throughput kernel of this plan (row passes, fit column pass) on the remaining ones,
 * so the solve of pair k runs beside the transforms of pair k + 1 instead of idling most of the GPU.  0 = off (default:
 * one pair at a time owns the whole GPU).  Results do not depend on the partition. */
int  sfftb_plan_set_partition(sfftb_plan* plan, int solver_sms);

/* ESS(SFFTSolution=None, Subtract=False): fit the kernel + background coefficients
 * (SFFTSubtract.py:10-412 / 479-752).  `solution` receives NEQ doubles. */
int  sfftb_fit(sfftb_plan* plan, const void* PixA_I, const void* PixA_J, int img_memkind, int img_dtype,
               double* solution, int sol_memkind);

/* ESS(SFFTSolution=solution, Subtract=True): Fourier-space kernel apply + inverse transform
 * (SFFTSubtract.py:429-461 / 754-807).  `diff` receives N0*N1 values of `diff_dtype`. */
int  sfftb_apply(sfftb_plan* plan, const void* PixA_I, const void* PixA_J, int img_memkind, int img_dtype,
                 const double* solution, int sol_memkind, void* diff, int diff_memkind, int diff_dtype);

/* GeneralSFFTSubtract.GSS without the contamination branch (SFFTSubtract.py:841-904 / 1373-1430):
 * fit on (mI, mJ), apply to (I, J), no host synchronisation in between. */
int  sfftb_gss(sfftb_plan* plan, const void* PixA_I, const void* PixA_J, const void* PixA_mI, const void* PixA_mJ,
               int img_memkind, int img_dtype, double* solution, int sol_memkind,
               void* diff, int diff_memkind, int diff_dtype);

/* Asynchronous form of sfftb_gss for HOST buffers (pinned memory recommended): submit queues the four H2D copies on
 * the plan's copy stream, the fit, the apply and the D2H of `diff` / `solution`, and returns; finish waits for this
 * plan's work only, runs the LU fallback if the Cholesky broke down, and reports errors like sfftb_gss.  One submission
 * per plan may be in flight.  Two plans bound to ONE compute stream (sfftb_plan_set_stream) and driven alternately keep
 * the PCIe link busy in both directions while the kernels of the other pair run -- the way a queue of pairs coming
 * from host memory (sfft/MultiEasySparsePacket.py:568-649 feeds one GPU from a host-side task queue) should be fed. */
int  sfftb_gss_submit(sfftb_plan* plan, const void* PixA_I, const void* PixA_J, const void* PixA_mI, const void* PixA_mJ,
                      int img_dtype, double* solution, void* diff, int diff_dtype);
/* sfftb_gss_submit for a pair whose four images (and both outputs) already live on the DEVICE: nothing is copied, the fit
 * and the apply are queued on the plan's stream without any host synchronisation (the forward row pass of the apply step
 * overlaps the Cholesky like in sfftb_gss), sfftb_gss_finish waits and reports.  A queue of device-resident pairs driven
 * through two plans alternately never drains the GPU between pairs. */
int  sfftb_gss_submit_device(sfftb_plan* plan, const void* PixA_I, const void* PixA_J, const void* PixA_mI, const void* PixA_mJ,
                             int img_dtype, double* solution, void* diff, int diff_dtype);
/* sfftb_gss_submit with the masked pair given as SPARSE DELTAS against the unmasked pair: mI = I except mI[idxI[k]] = valI[k]
 * (flat C-order pixel indices; values of the image dtype), likewise mJ.  The packets build their masked images exactly
 * that way -- stamps of the same arrays set to zero or to a fill value (sfft/CustomizedPacket.py:114-162,
 * sfft/EasySparsePacket.py) -- so two images instead of four cross the PCIe link per pair.  HOST pointers. */
int  sfftb_gss_submit_delta(sfftb_plan* plan, const void* PixA_I, const void* PixA_J, long long nI, const long long* idxI, const void* valI,
                            long long nJ, const long long* idxJ, const void* valJ, int img_dtype, double* solution, void* diff,
                            int diff_dtype);
/* The same for one science tile against the cached template; completes inside the call until the plan holds the
 * template's Cholesky factor (first tile).  `memkind` covers the images and both outputs: host buffers are copied on
 * the plan's copy stream, device buffers are used in place (tiles queued back to back leave no launch gaps). */
int  sfftb_gss_template_submit(sfftb_plan* plan, const void* PixA_J, const void* PixA_mJ, int memkind, int img_dtype,
                               double* solution, void* diff, int diff_dtype);
int  sfftb_gss_finish(sfftb_plan* plan);

/* Shared-template batch path (BASELINE config 4).  The reference re-transforms the template for every
 * pair (ESS is called per pair, sfft/MultiEasySparsePacket.py:568-649); here the row spectra of the convolved image
 * (I = template when ForceConv='REF', sfft/CustomizedPacket.py:148-162) are computed once:
 *   sfftb_template_prepare : row spectra of (PixA_I, PixA_mI) into the plan's template state;
 *   sfftb_template_state   : device pointer + size of that state (allocated on first use) -- the buffer a caller
 *                            broadcasts to the other GPUs with ONE ncclBroadcast / torch.distributed.broadcast;
 *   sfftb_template_mark_ready : on a receiving rank, after the broadcast landed in the state buffer (every change of the
 *                            state buffer must be followed by prepare or mark_ready: they invalidate the cached factor and
 *                            the cached segment spectra below);
 *   sfftb_gss_template     : GSS for one science image (PixA_J, PixA_mJ) against the cached template.  LHMAT = D^T D / N
 *                            involves the masked template only, so its Cholesky factor is kept from the first tile
 *                            and later tiles assemble the right-hand side and run the two substitutions; the segment
 *                            spectra of the template that the right-hand side needs are kept in device memory as well
 *                            (NH * segments * Fij * 4 KB, 252 MB for a 2048^2 template with KerPolyOrder 2; built by the
 *                            second tile, SFFTB_NO_ASPEC_CACHE=1 switches it off), so a tile transforms only J. */
int  sfftb_template_prepare(sfftb_plan* plan, const void* PixA_I, const void* PixA_mI, int img_memkind, int img_dtype);
int  sfftb_template_state(sfftb_plan* plan, void** device_ptr, size_t* bytes);
int  sfftb_template_mark_ready(sfftb_plan* plan);
/* Copy the whole shared-template state of `src` (row spectra; Cholesky factor, lag rows and cached segment spectra once `src` has
 * processed its first tiles) into `dst`, a plan of the same configuration on the same device: the plans of a multi-stream pipeline
 * (batch.TemplatePipeline) then share ONE factorisation per template instead of one per plan.  The reference has no counterpart
 * (it refits every pair from scratch, sfft/MultiEasySparsePacket.py:568-649). */
int  sfftb_template_clone(sfftb_plan* dst, sfftb_plan* src);
int  sfftb_gss_template(sfftb_plan* plan, const void* PixA_J, const void* PixA_mJ, int img_memkind, int img_dtype,
                        double* solution, int sol_memkind, void* diff, int diff_memkind, int diff_dtype);

/* Consumers of the Solution that every Easy*Packet calls right after GSS (sfft/EasySparsePacket.py:417-436):
 * Realize_MatchingKernel(XY_q).FromArray and Realize_FluxScaling(XY_q).FromArray
 * (sfft/utils/SFFTSolutionReader.py:116-196).  `xy` holds nq (x, y) pairs in FortranCoor (pixel centre (r, c) ->
 * (r + 1, c + 1)); `kerstack` receives (nq, L0, L1) kernels in the Cartesian-delta basis, `fscal` the nq flux
 * scalings; either output may be NULL.  With device pointers throughout nothing is copied and nothing synchronises. */
int  sfftb_realize(sfftb_plan* plan, const double* solution, int sol_memkind, const double* xy, int xy_memkind, int nq,
                   double* kerstack, double* fscal, int out_memkind);

/* The I/O edge of the packets on the device (sfft/CustomizedPacket.py:93-126, 183-203): the raw big-endian FITS data
 * block (NAXIS2 x NAXIS1, as in the file) is decoded -- byte swap, BITPIX conversion, BSCALE / BZERO -- straight into the
 * transposed (NAXIS1, NAXIS2) array the plan reads (the reference's `fits.getdata(..).T` + float64 copy on the host), the
 * difference image is encoded back; NaN union fill and NaN restore / sign flip of the packets as elementwise kernels.
 * All pointers are DEVICE pointers; work is queued on `cuda_stream`; nothing synchronises. */
int  sfftb_fits_decode(int device, void* cuda_stream, const void* raw, int bitpix, int naxis1, int naxis2, double bscale, double bzero,
                       void* out, int out_dtype);
int  sfftb_fits_encode(int device, void* cuda_stream, const void* img, int img_dtype, int naxis1, int naxis2, int bitpix, void* raw);
int  sfftb_nan_union_fill(int device, void* cuda_stream, void* A, void* B, const void* mA, const void* mB, int dtype, size_t n,
                          unsigned char* mask, int* flags);
int  sfftb_nan_mask_apply(int device, void* cuda_stream, void* D, int dtype, const unsigned char* mask, size_t n, double sign);

/* Noise decorrelation, the step right after the subtraction in the Roman / JWST pipelines:
 * PureCupy_DeCorrelation_Calculator.PCDC (sfft/utils/PureCupyDeCorrelationCalculator.py:46-125), DeCorrelation_Calculator.DCC
 * (sfft/utils/DeCorrelationCalculator.py:11-103) and BSpline_DeCorrelation.BDC (sfft/BSplineSFFT.py:4757-4868).
 * `nker` match kernels are packed in `kdata` (row-major (L0, L1) stamps one after the other, odd sizes), kshape = {L0, L1} per
 * kernel, role = 0 (J queue) / 1 (I queue) / 2 (the match kernel of the subtraction; at most one), sig = background sigma per
 * kernel (unused for role 2); a missing kernel (None in the reference) is passed as the 3 x 3 delta by the caller.  All HOST.
 *   DeNo = sum_J sig^2 |FK_J|^2 / NJ^2 + |FMK|^2 sum_I sig^2 |FK_I|^2 / NI^2 on the (N0, N1) Fourier grid, clipped from below at
 *   max(DeNo) / clip_ratio when clip_ratio > 0 (BDC), FKDECO = 1 / sqrt(DeNo).
 * out_mode 0: `out` receives FKDECO, N0 * N1 doubles (normalize: divided by its [0, 0] value, PCDC REAL_OUTPUT=False);
 * out_mode 1: `out` receives the real-space kernel of shape (LO0, LO1): ifft2(FKDECO).real, inverse circular shift and tail
 *             truncation (KERNEL_CSZ_INV / iCSZ; normalize: unit sum); *lost_weight = 1 - sum|K| / sum|ifft2(FKDECO)| for grids
 *             of at most 65536 points (the sizes DCC / BDC use), NaN otherwise.
 * The kernels' spectra are evaluated in closed form from their taps: no image-sized transform is computed. */
int  sfftb_decorr(int device, void* cuda_stream, int N0, int N1, int nker, const double* kdata, const int* kshape, const int* role,
                  const double* sig, double clip_ratio, int out_mode, int LO0, int LO1, int normalize, double* out, int out_memkind,
                  double* lost_weight);
/* PureCupy_FFTKits.FFT_CONVOLVE (sfft/utils/PureCupyFFTKits.py:71-105), the convolution that applies a decorrelation kernel:
 * out = img (*) kernel with the image extended by `pad_fill` and NaN samples replaced by `nan_fill` (fill_nan = 0 keeps them,
 * NAN_FILL_VALUE=None), evaluated directly in real space (same result as the zero-padded FFT product, cropped).  `kernel`
 * (L0, L1) odd, HOST; img / out: `memkind`, `dtype`. */
int  sfftb_convolve(int device, void* cuda_stream, const void* img, int dtype, int N0, int N1, const double* kernel, int L0, int L1,
                    double pad_fill, double nan_fill, int fill_nan, int normalize_kernel, void* out, int memkind);

/* BSpline_GridConvolve.GSVC_GPU / GSVC_CPU (sfft/BSplineSFFT.py:4870-5010): grid-wise spatially varying convolution.  `labels`
 * (N0, N1) int32 gives every pixel the index of its grid cell (AllocatedL), `kerstack` (nseg, L0, L1) the kernel of every cell
 * (HOST, odd sizes; normalize_kernel divides each by its sum); out[r, c] = image convolved with the kernel of the pixel's cell,
 * zero outside the image, NaN samples replaced by `nan_fill`.  img / labels / out: `memkind`; img / out: `dtype`. */
int  sfftb_convolve_grid(int device, void* cuda_stream, const void* img, int dtype, int N0, int N1, const int* labels, int nseg,
                         const double* kerstack, int L0, int L1, double nan_fill, int normalize_kernel, void* out, int memkind);

/* Kernel regularisation (sfft/BSplineSFFT.py:3570-3700, REGULARIZE_KERNEL / LAMBDA_REGULARIZE): every following fit solves
 * (LHMAT + lambda * REGMAT) x = RHb with REGMAT[(k,c),(k',c')] = SCALE^2 * SST[k,k'] * iREG[c,c'] (fill_regmat, :2091-2119).
 * SST is the (Fij x Fij) Gram matrix of the kernel basis at the regularisation coordinates, iREG the (Fab x Fab)
 * Laplacian penalty in the modified-delta basis; host pointers, row-major.  SST == NULL switches it off. */
int  sfftb_set_regularizer(sfftb_plan* plan, const double* SST, const double* iREG, double lambda);
/* Plans with sca_degree > 0: where a centre tap is involved REGMAT uses the scaling basis instead (fill_regmat of the
 * SEPARATE-VARYING mode, :2122-2166): CSST = kernel x scaling Gram matrix, DSST = scaling x scaling, both (Fij x Fij) with
 * zero rows / columns beyond ScaFij.  Call after sfftb_set_regularizer. */
int  sfftb_set_regularizer_varying(sfftb_plan* plan, const double* CSST, const double* DSST);

/* ---- general spatial bases: sfft/BSplineSFFT.py (B-spline or polynomial variation of the kernel, the photometric scaling and
 * the background, any degree) --------------------------------------------------------------------------------------------
 * Every 2-D basis function of the reference is a tensor product of two 1-D functions of the scaled pixel-centre
 * coordinates (Create_BSplineBasis :2624-2634; KerSpatial / ScaSpatial / BkgSpatial :276-458), so a basis is handed over as
 * its 1-D tables; the caller evaluates them (B-splines with any knot vector, monomials, ...).  Host pointers, copied. */
typedef struct sfftb_basis {
    int nu, nv;            /* number of 1-D functions along axis 0 (rows, x) / axis 1 (columns, y) */
    int nf;                /* number of 2-D basis functions: Fij, ScaFij or Fpq */
    const double* U;       /* (nu, N0) row-major: U_i at the pixel centres (r + 1) / N0 */
    const double* V;       /* (nv, N1) row-major: V_j at (c + 1) / N1 */
    const int* fu;         /* (nf): function k = U[fu[k]] (x) V[fv[k]], in the reference's ij / pq order */
    const int* fv;
} sfftb_basis;

#define SFFTB_SCALING_ENTANGLED       0   /* SEPARATE_SCALING=False (:77-86) */
#define SFFTB_SCALING_CONSTANT_DROP   1   /* SEPARATE-CONSTANT, polynomial kernel: the stripes (ij > 0, 00) are dropped (TweakLS :2204-2233) */
#define SFFTB_SCALING_CONSTANT_SUM    2   /* SEPARATE-CONSTANT, B-spline kernel: the stripes are summed (:2235-2272) and the value is copied
                                             back to every (ij, 00) (Restore_Solution :3764-3771) */
#define SFFTB_SCALING_VARYING         3   /* SEPARATE-VARYING: the (ij, 00) unknown of ij < ScaFij scales I x ScaBasis_ij (:2487-2495), the
                                             others are dropped (:3733-3747); `sca` gives that basis */

/* BSplineSFFT.SingleSFFTConfigure.SSC (:2538-2611).  cfg->DK / DB are informational (the degrees), cfg->const_phot_ratio,
 * sca_degree and fold are ignored; the Solution layout is the reference's: NEQ = ker->nf * Fab + bkg->nf.  `sca` may be NULL
 * unless scaling_mode is SFFTB_SCALING_VARYING.  Such a plan serves sfftb_fit / sfftb_apply / sfftb_gss, the regulariser and
 * the export hooks; the shared-template and asynchronous entry points return SFFTB_EINVAL for it. */
int  sfftb_plan_create_general(sfftb_plan** out, const sfftb_config* cfg, const sfftb_basis* ker, const sfftb_basis* sca,
                               const sfftb_basis* bkg, int scaling_mode);

/* Parity hook for every kind of plan: the system that was actually solved by the last fit -- after the stripe tweak, in the
 * order of the reference's tweaked system (TweakLS; what the reference hands to its solver): L (n x n, row-major) and b (n),
 * n = NEQ_FSfree of sfftb_plan_dims.  Host pointers; either may be NULL. */
int  sfftb_export_solved_system(sfftb_plan* plan, double* L, double* b);

/* Parity hook: the full (NEQ x NEQ) LHMAT and (NEQ) RHb of the last fit, before stripe removal,
 * in the reference's layout (what FillLS_* produce, SFFTSubtract.py:244-380).  Host pointers. */
int  sfftb_export_normal_eq(sfftb_plan* plan, double* LHMAT, double* RHb);

/* Device-event stage timings of the last fit/apply, milliseconds.
 * ms[0] row spectra (fit), [1] column pass (fit), [2] lag reductions + fill, [3] solve,
 * [4] row spectra (apply), [5] column pass (apply), [6] inverse rows, [7] the fit column kernel alone (segmented path; [1] also
 * holds the column-moment kernel).  Enabled by sfftb_plan_set_timing. */
int  sfftb_plan_set_timing(sfftb_plan* plan, int enable);
int  sfftb_timings(sfftb_plan* plan, float* ms, int n);

/* Which factorisation the last fit used: 1 = Cholesky, 2 = pivoted LU fallback, 3 = substitutions with the cached
 * Cholesky factor (template path: LHMAT depends on the masked template only, so tiles after the first reuse it). */
int  sfftb_last_solver(const sfftb_plan* plan);

/* General-basis plans: out8 = { passes of the fit column kernel, stored planes staged summed over the passes, lag rows per
 * column, stored planes, column planes, unknowns of the solved system, apply planes, stored planes read by the FIR }
 * (bench.py's algorithmic-byte count). */
int  sfftb_gen_info(const sfftb_plan* plan, int* out8);

/* Number of kernel launches issued by this plan since creation (bench.py's gpu_launches). */
long long sfftb_launch_count(const sfftb_plan* plan);

/* Debug / unit-test hooks (used by tests/ only). */
/* batched 1-D complex FFT of `nbatch` rows of length n through the in-shared-memory engine;
 * host pointers, interleaved re/im doubles; sign = -1 forward, +1 unnormalised inverse. */
int  sfftb_dbg_fft1d(int device, int n, int nbatch, int sign, const double* in, double* out);
/* row spectra of the last call as stored in HBM: out[(j*NH + k1)*N0 + r] complex128 (host), j = 0..DK for
 * which = 0 (I planes), single plane for which = 1 (J). */
int  sfftb_dbg_row_spectra(sfftb_plan* plan, int which, double* out);
/* lag tables of the last fit: R (npairs, 4w0+1, 4w1+1), RJ (Fij, 2w0+1, 2w1+1), RT (Fij, Fpq, 2w0+1, 2w1+1),
 * RJT (Fpq); any pointer may be NULL. */
int  sfftb_dbg_lag_tables(sfftb_plan* plan, double* R, double* RJ, double* RT, double* RJT);

#ifdef __cplusplus
}
#endif
#endif

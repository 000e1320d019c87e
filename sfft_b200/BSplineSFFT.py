"""
sfft/BSplineSFFT.py on the B200 core.

`BSplineSFFT.py` generalises sfftcore: the kernel, the scaling and the background may vary as total-degree
polynomials or as tensor-product B-splines, the photometric scaling can be entangled with the kernel or separate
(constant / varying), and the kernel can be regularised by a Laplacian penalty (:3570-3700).  This module keeps its
call signatures (`SingleSFFTConfigure.SSC` :2538, `ElementalSFFTSubtract.ESS`, `GeneralSFFTSubtract.GSS` :3882,
`BSpline_Packet.BSP` :3969) and serves:

  * KerSpType = BkgSpType = 'Polynomial', degrees 0..3;
  * SEPARATE_SCALING=False (ENTANGLED == sfftcore ConstPhotRatio=False), SEPARATE_SCALING=True with
    ScaSpDegree=0 (SEPARATE-CONSTANT == ConstPhotRatio=True: for a polynomial kernel TweakLS drops the stripes, :2204-2233)
    and with a polynomial ScaSpDegree in 1..KerSpDegree (SEPARATE-VARYING: the centre-tap unknown of plane k scales the
    image times the k-th scaling basis function, :2487-2495; planes beyond ScaFij lose that unknown, :3733-3747);
  * REGULARIZE_KERNEL with XY_REGULARIZE / WEIGHT_REGULARIZE / LAMBDA_REGULARIZE / IGNORE_LAPLACIAN_KERCENT:
    the two Kronecker factors of REGMAT are built here and added inside the native matrix fill (sfftb_set_regularizer).

  * KerSpType / ScaSpType / BkgSpType = 'B-Spline' with any internal knots, and polynomial degrees above 3, through the
    general-basis plan (sfftb_plan_create_general): the 1-D basis tables are evaluated here exactly as Create_BSplineBasis
    (:2624-2634) does and handed to the CUDA library, which treats every basis image as U_i(r) V_j(c)
    (csrc/kernels_gen.cuh).  SEPARATE-CONSTANT sums the stripes for a B-spline kernel and drops them for a polynomial
    one (TweakLS :2202-2272), SEPARATE-VARYING takes its own scaling basis, Restore_Solution (:3704-3783) is mirrored.

Pure-polynomial configurations of degree <= 3 keep the specialised sfftcore kernels (same results, one fit launch).
"""
import os.path as pa
import time
import numpy as np

from . import fitsio
from .plan import Plan
from .sfftcore.SFFTConfigure import _current_device
from .sfftcore.SFFTSubtract import ElementalSFFTSubtract as _ESS, GeneralSFFTSubtract as _GSS

__all__ = ['SingleSFFTConfigure', 'ElementalSFFTSubtract', 'GeneralSFFTSubtract', 'BSpline_Packet', 'regularizer_factors',
           'Read_SFFTSolution', 'BSpline_MatchingKernel', 'ConvKernel_Convertion', 'BSpline_DeCorrelation', 'BSpline_GridConvolve']


def _laplacian_penalty(w0, w1, IGNORE_LAPLACIAN_KERCENT):
    """iREGMAT (Fab, Fab) of fill_iregmat (:2051-2086): 2 M^T (Lap^T Lap) M with Lap the 5-point Laplacian on the
    (L0, L1) kernel stamp (neighbour count on the diagonal, -1 on the 4-neighbours, zero-filled border, :3641-3666) and
    M the map from modified-delta coefficients to kernel pixels (every non-centre tap also subtracts from the centre)."""
    L0, L1 = 2 * w0 + 1, 2 * w1 + 1
    Fab = L0 * L1
    rr, cc = np.divmod(np.arange(Fab), L1)
    adj = ((np.abs(rr[:, None] - rr[None, :]) + np.abs(cc[:, None] - cc[None, :])) == 1)
    LAP = np.diag(adj.sum(axis=1).astype(float)) - adj.astype(float)
    c0 = w0 * L1 + w1
    if IGNORE_LAPLACIAN_KERCENT:                                                       # :3670-3676
        LAP[[c0 - L1, c0 - 1, c0, c0 + 1, c0 + L1], :] = 0.0
    M = np.eye(Fab)
    M[c0, :] = -1.0
    M[c0, c0] = 1.0
    return 2.0 * (M.T @ (LAP.T @ LAP) @ M)


def _bspline_basis(N, IntKnot, Degree, ReqCoord=None):
    """Create_BSplineBasis / Create_BSplineBasis_Req (:2624-2646): clamped knots [0.5]*(k+1) ++ IntKnot ++ [N+0.5]*(k+1),
    scaled by 1/N, evaluated at the pixel centres (1 + arange(N)) / N or at the requested scaled coordinates."""
    from scipy.interpolate import BSpline
    coord = (1.0 + np.arange(N)) / N if ReqCoord is None else np.asarray(ReqCoord, float)
    knot = np.concatenate(([0.5] * (Degree + 1), np.asarray(IntKnot, float), [N + 0.5] * (Degree + 1))) / N
    Nc = len(IntKnot) + Degree + 1
    return np.array([BSpline(t=knot, c=(np.arange(Nc) == idx).astype(float), k=Degree, extrapolate=False)(coord)
                     for idx in range(Nc)])


def _basis_tables(SpType, Degree, KnotX, KnotY, N0, N1, CX=None, CY=None):
    """(U, V, fu, fv): 1-D tables along x / y and the (i, j) of every 2-D basis function in the reference's order
    (REF_ij / REF_pq, :2764-2772).  CX / CY: evaluate at requested scaled coordinates instead of the pixel grid."""
    if SpType == 'Polynomial':
        cx = (1.0 + np.arange(N0)) / N0 if CX is None else np.asarray(CX, float)
        cy = (1.0 + np.arange(N1)) / N1 if CY is None else np.asarray(CY, float)
        U = np.array([cx ** i for i in range(Degree + 1)])
        V = np.array([cy ** j for j in range(Degree + 1)])
        ij = [(i, j) for i in range(Degree + 1) for j in range(Degree + 1 - i)]
    else:
        U = _bspline_basis(N0, KnotX, Degree, CX)
        V = _bspline_basis(N1, KnotY, Degree, CY)
        ij = [(i, j) for i in range(U.shape[0]) for j in range(V.shape[0])]
    return U, V, np.array([a for a, _ in ij], np.int32), np.array([b for _, b in ij], np.int32)


def _gram_factors(P, N0, N1, w0, w1, XY_REGULARIZE, WEIGHT_REGULARIZE, IGNORE_LAPLACIAN_KERCENT):
    """(SST, iREG[, CSST, DSST]) of fill_regmat (:2091-2166) for any kernel / scaling basis."""
    XY = np.asarray(XY_REGULARIZE, float)
    if XY.ndim != 2 or XY.shape[1] != 2 or XY.shape[0] < 1:
        raise Exception('MeLOn ERROR: XY_REGULARIZE must have shape (N_points, 2)')
    CX, CY = XY[:, 0] / N0, XY[:, 1] / N1
    U, V, fu, fv = _basis_tables(P['KerSpType'], P['DK'], P['KerIntKnotX'], P['KerIntKnotY'], N0, N1, CX, CY)
    SP = U[fu] * V[fv]                                                                # (Fij, NREG), :3576-3592
    if WEIGHT_REGULARIZE is None:
        Wd = np.full(XY.shape[0], 1.0 / XY.shape[0])
    else:
        Wd = np.asarray(WEIGHT_REGULARIZE, float)
        if Wd.shape != (XY.shape[0],):
            raise Exception('MeLOn ERROR: WEIGHT_REGULARIZE must have shape (N_points,)')
        Wd = Wd / Wd.sum()
    SST, iREG = (SP * Wd) @ SP.T, _laplacian_penalty(w0, w1, IGNORE_LAPLACIAN_KERCENT)
    if P['SCALING_MODE'] != 'SEPARATE-VARYING':
        return SST, iREG
    U, V, fu, fv = _basis_tables(P['ScaSpType'], P['DS'], P['ScaIntKnotX'], P['ScaIntKnotY'], N0, N1, CX, CY)
    Sca = U[fu] * V[fv]
    Sca = np.concatenate([Sca, np.zeros((SP.shape[0] - Sca.shape[0], XY.shape[0]))], axis=0)   # placeholder rows :3614-3620
    return SST, iREG, (SP * Wd) @ Sca.T, (Sca * Wd) @ Sca.T


def regularizer_factors(N0, N1, w0, w1, DK, XY_REGULARIZE, WEIGHT_REGULARIZE=None, IGNORE_LAPLACIAN_KERCENT=True, DS=None):
    """(SST, iREG) with REGMAT = SCALE^2 * kron(SST, iREG) on the kernel block (fill_regmat, :2091-2119), polynomial
    kernel basis: SST = SPMAT W SPMAT^T, SPMAT[k, n] = cx_n^i cy_n^j at the requested coordinates (:3576-3582, 3624-3633).
    With DS (SEPARATE-VARYING) also (CSST, DSST): the Gram matrices with the scaling basis, zero-padded to Fij (:3594-3637)."""
    XY = np.asarray(XY_REGULARIZE, float)
    if XY.ndim != 2 or XY.shape[1] != 2 or XY.shape[0] < 1:
        raise Exception('MeLOn ERROR: XY_REGULARIZE must have shape (N_points, 2)')
    cx, cy = XY[:, 0] / N0, XY[:, 1] / N1
    SP = np.array([cx ** i * cy ** j for i in range(DK + 1) for j in range(DK + 1 - i)])
    if WEIGHT_REGULARIZE is None:
        Wd = np.full(XY.shape[0], 1.0 / XY.shape[0])
    else:
        Wd = np.asarray(WEIGHT_REGULARIZE, float)
        if Wd.shape != (XY.shape[0],):
            raise Exception('MeLOn ERROR: WEIGHT_REGULARIZE must have shape (N_points,)')
        Wd = Wd / Wd.sum()
    SST, iREG = (SP * Wd) @ SP.T, _laplacian_penalty(w0, w1, IGNORE_LAPLACIAN_KERCENT)
    if DS is None:
        return SST, iREG
    Sca = np.array([cx ** i * cy ** j for i in range(DS + 1) for j in range(DS + 1 - i)])
    Sca = np.concatenate([Sca, np.zeros((SP.shape[0] - Sca.shape[0], XY.shape[0]))], axis=0)
    return SST, iREG, (SP * Wd) @ Sca.T, (Sca * Wd) @ Sca.T


class SingleSFFTConfigure:
    @staticmethod
    def SSC(NX, NY, KerHW=8, KerSpType='Polynomial', KerSpDegree=2, KerIntKnotX=[], KerIntKnotY=[],
            SEPARATE_SCALING=True, ScaSpType='Polynomial', ScaSpDegree=0, ScaIntKnotX=[], ScaIntKnotY=[],
            BkgSpType='Polynomial', BkgSpDegree=2, BkgIntKnotX=[], BkgIntKnotY=[],
            REGULARIZE_KERNEL=False, IGNORE_LAPLACIAN_KERCENT=True, XY_REGULARIZE=None, WEIGHT_REGULARIZE=None,
            LAMBDA_REGULARIZE=1e-6, BACKEND_4SUBTRACT='B200', MAX_THREADS_PER_BLOCK=8,
            MINIMIZE_GPU_MEMORY_USAGE=False, NUM_CPU_THREADS_4SUBTRACT=8, VERBOSE_LEVEL=2,
            CUDA_DEVICE=None, STORAGE='fp64', FORCE_GENERAL_PLAN=False):
        """Arguments as BSplineSFFT.SingleSFFTConfigure.SSC (:2538-2545); MAX_THREADS_PER_BLOCK, MINIMIZE_GPU_MEMORY_USAGE and
        NUM_CPU_THREADS_4SUBTRACT are accepted and ignored; FORCE_GENERAL_PLAN runs a polynomial configuration through the
        table-driven kernels as well (test hook).  Returns (SFFTParam_dict, SFFTModule_dict) with the keys of
        :204-273."""
        if BACKEND_4SUBTRACT not in ('B200', 'Cupy'):
            raise Exception("MeLOn ERROR: BACKEND_4SUBTRACT=%r is not available in sfft_b200 (use 'B200')" % (BACKEND_4SUBTRACT,))
        N0, N1, w0, w1 = int(NX), int(NY), int(KerHW), int(KerHW)
        DK, DB = int(KerSpDegree), int(BkgSpDegree)
        assert DK >= 0 and DB >= 0                                                       # :33, :79
        assert KerSpType in ['Polynomial', 'B-Spline'] and BkgSpType in ['Polynomial', 'B-Spline']
        if SEPARATE_SCALING:
            DS = int(ScaSpDegree)
            assert DS >= 0 and ScaSpType in ['Polynomial', 'B-Spline']
        if not SEPARATE_SCALING:
            SCALING_MODE = 'ENTANGLED'
        elif int(ScaSpDegree) == 0:
            SCALING_MODE = 'SEPARATE-CONSTANT'
        else:
            SCALING_MODE = 'SEPARATE-VARYING'
        if KerSpType == 'B-Spline' and DK == 0:
            assert len(KerIntKnotX) == 0 and len(KerIntKnotY) == 0                      # :36-38
        if BkgSpType == 'B-Spline' and DB == 0:
            assert len(BkgIntKnotX) == 0 and len(BkgIntKnotY) == 0                      # :79-81

        def dof(SpType, Degree, KX, KY):
            if SpType == 'Polynomial':
                return -1, -1, ((Degree + 1) * (Degree + 2)) // 2
            return len(KX) + Degree + 1, len(KY) + Degree + 1, (len(KX) + Degree + 1) * (len(KY) + Degree + 1)
        Fi, Fj, Fij = dof(KerSpType, DK, KerIntKnotX, KerIntKnotY)
        Fp, Fq, Fpq = dof(BkgSpType, DB, BkgIntKnotX, BkgIntKnotY)
        ScaFi = ScaFj = ScaFij = None
        if SCALING_MODE == 'SEPARATE-VARYING':
            if ScaSpType == 'B-Spline' and DS == 0:
                assert len(ScaIntKnotX) == 0 and len(ScaIntKnotY) == 0
            ScaFi, ScaFj, ScaFij = dof(ScaSpType, DS, ScaIntKnotX, ScaIntKnotY)
            assert ScaFij <= Fij                                                        # :190
        general = (FORCE_GENERAL_PLAN or KerSpType != 'Polynomial' or BkgSpType != 'Polynomial' or DK > 3 or DB > 3 or
                   (SCALING_MODE == 'SEPARATE-VARYING' and (ScaSpType != 'Polynomial' or DS > DK)))
        if VERBOSE_LEVEL in [1, 2]:
            print('\n --//--//--//--//-- TRIGGER SFFT COMPILATION --//--//--//--//-- ')
            print('\n ---//--- %s Kernel | KerSpDegree %d | KerHW %d ---//---' % (KerSpType, DK, w0))
            print('\n ---//--- [%s] Scaling ---//---' % SCALING_MODE)
            print('\n ---//--- %s Background | BkgSpDegree %d ---//---' % (BkgSpType, DB))
        L0, L1 = 2 * w0 + 1, 2 * w1 + 1
        Fab = L0 * L1
        Fijab, NEQ = Fij * Fab, Fij * Fab + Fpq
        NEQt = NEQ - Fij + 1 if SCALING_MODE == 'SEPARATE-CONSTANT' else NEQ            # :199-200
        if SCALING_MODE == 'SEPARATE-VARYING':
            NEQt = NEQ - Fij + ScaFij                                                   # :201-202
        SCALE = np.float64(1 / (N0 * N1))
        P = dict(KerHW=KerHW, KerSpType=KerSpType, KerSpDegree=KerSpDegree, KerIntKnotX=KerIntKnotX, KerIntKnotY=KerIntKnotY,
                 SEPARATE_SCALING=SEPARATE_SCALING, BkgSpType=BkgSpType, BkgSpDegree=BkgSpDegree, BkgIntKnotX=BkgIntKnotX,
                 BkgIntKnotY=BkgIntKnotY, REGULARIZE_KERNEL=REGULARIZE_KERNEL, IGNORE_LAPLACIAN_KERCENT=IGNORE_LAPLACIAN_KERCENT,
                 XY_REGULARIZE=XY_REGULARIZE, WEIGHT_REGULARIZE=WEIGHT_REGULARIZE, LAMBDA_REGULARIZE=LAMBDA_REGULARIZE,
                 MAX_THREADS_PER_BLOCK=MAX_THREADS_PER_BLOCK, MINIMIZE_GPU_MEMORY_USAGE=MINIMIZE_GPU_MEMORY_USAGE,
                 N0=N0, N1=N1, w0=w0, w1=w1, DK=DK, DB=DB, SCALE=SCALE, SCALE_L=np.float64(1 / SCALE), L0=L0, L1=L1, Fab=Fab,
                 Fi=Fi, Fj=Fj, Fij=Fij, Fp=Fp, Fq=Fq, Fpq=Fpq, Fijab=Fijab, FOMG=Fij ** 2, FGAM=Fij * Fpq, FTHE=Fij,
                 FPSI=Fpq * Fij, FPHI=Fpq ** 2, FDEL=Fpq, NEQ=NEQ, NEQt=NEQt, SCALING_MODE=SCALING_MODE)
        if SEPARATE_SCALING:
            P.update(ScaSpType=ScaSpType, ScaSpDegree=ScaSpDegree, ScaIntKnotX=ScaIntKnotX, ScaIntKnotY=ScaIntKnotY, DS=DS)
        if SCALING_MODE == 'SEPARATE-VARYING':
            P.update(ScaFi=ScaFi, ScaFj=ScaFj, ScaFij=ScaFij)
        device = _current_device() if CUDA_DEVICE is None else int(CUDA_DEVICE)
        if general:
            from . import _lib as B
            ker = _basis_tables(KerSpType, DK, KerIntKnotX, KerIntKnotY, N0, N1)
            bkg = _basis_tables(BkgSpType, DB, BkgIntKnotX, BkgIntKnotY, N0, N1)
            sca = None
            if SCALING_MODE == 'ENTANGLED':
                mode = B.SCALING_ENTANGLED
            elif SCALING_MODE == 'SEPARATE-CONSTANT':
                mode = B.SCALING_CONSTANT_SUM if KerSpType == 'B-Spline' else B.SCALING_CONSTANT_DROP    # TweakLS :2204-2272
            else:
                mode = B.SCALING_VARYING
                sca = _basis_tables(ScaSpType, DS, ScaIntKnotX, ScaIntKnotY, N0, N1)
            plan = Plan.general(N0, N1, w0, w1, ker, bkg, mode, sca=sca, device=device, storage=STORAGE, DK=DK, DB=DB)
            assert plan.dims['NEQ'] == NEQ and plan.dims['NEQ_FSfree'] == NEQt
        else:
            plan = Plan(N0, N1, w0, w1, DK, DB, SCALING_MODE != 'ENTANGLED', device=device, storage=STORAGE,
                        sca_degree=DS if SCALING_MODE == 'SEPARATE-VARYING' else 0)
        if REGULARIZE_KERNEL:
            if XY_REGULARIZE is None:
                raise Exception('MeLOn ERROR: REGULARIZE_KERNEL needs XY_REGULARIZE')
            fac = _gram_factors(P, N0, N1, w0, w1, XY_REGULARIZE, WEIGHT_REGULARIZE, IGNORE_LAPLACIAN_KERCENT)
            plan.set_regularizer(fac[0], fac[1], float(LAMBDA_REGULARIZE), *(fac[2:]))
        if VERBOSE_LEVEL in [1, 2]:
            print('\n --//--//--//--//-- EXIT SFFT COMPILATION --//--//--//--//-- ')
        return (P, {'BACKEND': 'B200', 'plan': plan, 'device': device, 'storage': STORAGE})


class ElementalSFFTSubtract:
    ESS = staticmethod(_ESS.ESS)            # same contract as sfftcore's ESS once the plan exists (:3855-3877)


class GeneralSFFTSubtract:
    GSS = staticmethod(_GSS.GSS)            # :3882-3964 (fit on the masked pair, subtract, optional contamination mask)


def _read_T(path):
    return np.ascontiguousarray(fitsio.getdata(path).T, np.float64)


class BSpline_Packet:
    @staticmethod
    def BSP(FITS_REF, FITS_SCI, FITS_mREF, FITS_mSCI, FITS_DIFF=None, FITS_Solution=None, ForceConv='REF', GKerHW=8,
            KerSpType='Polynomial', KerSpDegree=2, KerIntKnotX=[], KerIntKnotY=[], SEPARATE_SCALING=True,
            ScaSpType='Polynomial', ScaSpDegree=0, ScaIntKnotX=[], ScaIntKnotY=[], BkgSpType='Polynomial', BkgSpDegree=2,
            BkgIntKnotX=[], BkgIntKnotY=[], REGULARIZE_KERNEL=False, IGNORE_LAPLACIAN_KERCENT=True, XY_REGULARIZE=None,
            WEIGHT_REGULARIZE=None, LAMBDA_REGULARIZE=1e-6, BACKEND_4SUBTRACT='B200', CUDA_DEVICE_4SUBTRACT='0',
            MAX_THREADS_PER_BLOCK=8, MINIMIZE_GPU_MEMORY_USAGE=False, NUM_CPU_THREADS_4SUBTRACT=8, VERBOSE_LEVEL=2,
            STORAGE='fp64'):
        """FITS in -> FITS out, as BSpline_Packet.BSP (:3969-4260)."""
        arrays = [_read_T(f) for f in (FITS_REF, FITS_SCI, FITS_mREF, FITS_mSCI)]
        Solution, PixA_DIFF, SFFTConfig = BSpline_Packet.BSP_arrays(
            *arrays, ForceConv=ForceConv, GKerHW=GKerHW, KerSpType=KerSpType, KerSpDegree=KerSpDegree,
            KerIntKnotX=KerIntKnotX, KerIntKnotY=KerIntKnotY, SEPARATE_SCALING=SEPARATE_SCALING, ScaSpType=ScaSpType,
            ScaSpDegree=ScaSpDegree, ScaIntKnotX=ScaIntKnotX, ScaIntKnotY=ScaIntKnotY, BkgSpType=BkgSpType,
            BkgSpDegree=BkgSpDegree, BkgIntKnotX=BkgIntKnotX, BkgIntKnotY=BkgIntKnotY, REGULARIZE_KERNEL=REGULARIZE_KERNEL,
            IGNORE_LAPLACIAN_KERCENT=IGNORE_LAPLACIAN_KERCENT, XY_REGULARIZE=XY_REGULARIZE, WEIGHT_REGULARIZE=WEIGHT_REGULARIZE,
            LAMBDA_REGULARIZE=LAMBDA_REGULARIZE, BACKEND_4SUBTRACT=BACKEND_4SUBTRACT, CUDA_DEVICE_4SUBTRACT=CUDA_DEVICE_4SUBTRACT,
            VERBOSE_LEVEL=VERBOSE_LEVEL, STORAGE=STORAGE, _return_config=True)
        if FITS_DIFF is not None:                                                      # :4218-4236
            cards, _ = fitsio.read_header(FITS_SCI)
            fitsio.writeto(FITS_DIFF, PixA_DIFF.T, base_cards=cards, updates=[
                ('NAME_REF', pa.basename(FITS_REF), 'MeLOn: SFFT'), ('NAME_SCI', pa.basename(FITS_SCI), 'MeLOn: SFFT'),
                ('KERHW', GKerHW, 'MeLOn: SFFT'), ('CONVD', ForceConv, 'MeLOn: SFFT'),
                ('KSPTYPE', str(KerSpType), 'MeLOn: SFFT'), ('KSPDEG', KerSpDegree, 'MeLOn: SFFT'),
                ('BSPTYPE', str(BkgSpType), 'MeLOn: SFFT'), ('BSPDEG', BkgSpDegree, 'MeLOn: SFFT'),
                ('SEPSCA', str(SEPARATE_SCALING), 'MeLOn: SFFT'), ('REGKER', str(REGULARIZE_KERNEL), 'MeLOn: SFFT')])
        if FITS_Solution is not None:                                                  # :4277-4354 (same header keys)
            P = SFFTConfig[0]
            C = 'SFFT'
            ups = [('NAME_REF', pa.basename(FITS_REF), C), ('NAME_SCI', pa.basename(FITS_SCI), C), ('BEND4SUB', BACKEND_4SUBTRACT, C),
                   ('CONVD', ForceConv, C), ('KERHW', GKerHW, C), ('KSPTYPE', str(KerSpType), C), ('KSPDEG', KerSpDegree, C),
                   ('NKIKX', len(KerIntKnotX), C)]
            ups += [('KIKX%d' % i, float(k), C) for i, k in enumerate(KerIntKnotX)] + [('NKIKY', len(KerIntKnotY), C)]
            ups += [('KIKY%d' % i, float(k), C) for i, k in enumerate(KerIntKnotY)] + [('SEPSCA', str(SEPARATE_SCALING), C)]
            if SEPARATE_SCALING:
                ups += [('SSPTYPE', str(ScaSpType), C), ('SSPDEG', ScaSpDegree, C), ('NSIKX', len(ScaIntKnotX), C)]
                ups += [('SIKX%d' % i, float(k), C) for i, k in enumerate(ScaIntKnotX)] + [('NSIKY', len(ScaIntKnotY), C)]
                ups += [('SIKY%d' % i, float(k), C) for i, k in enumerate(ScaIntKnotY)]
            ups += [('BSPTYPE', str(BkgSpType), C), ('BSPDEG', BkgSpDegree, C), ('NBIKX', len(BkgIntKnotX), C)]
            ups += [('BIKX%d' % i, float(k), C) for i, k in enumerate(BkgIntKnotX)] + [('NBIKY', len(BkgIntKnotY), C)]
            ups += [('BIKY%d' % i, float(k), C) for i, k in enumerate(BkgIntKnotY)]
            ups += [('REGKER', str(REGULARIZE_KERNEL), C), ('ILKC', str(IGNORE_LAPLACIAN_KERCENT), C),
                    ('NREG', -1 if XY_REGULARIZE is None else int(np.asarray(XY_REGULARIZE).shape[0]), C),
                    ('REGW', 'UNIFORM' if WEIGHT_REGULARIZE is None else 'SPECIFIED', C), ('REGLAMB', LAMBDA_REGULARIZE, C)]
            ups += [(k, P[v], C) for k, v in (('N0', 'N0'), ('N1', 'N1'), ('W0', 'w0'), ('W1', 'w1'), ('DK', 'DK'), ('DB', 'DB'))]
            if SEPARATE_SCALING:
                ups += [('DS', P['DS'], C)]
            ups += [(k, P[v], C) for k, v in (('L0', 'L0'), ('L1', 'L1'), ('FAB', 'Fab'), ('FI', 'Fi'), ('FJ', 'Fj'), ('FIJ', 'Fij'),
                                              ('FP', 'Fp'), ('FQ', 'Fq'), ('FPQ', 'Fpq'))]
            if SEPARATE_SCALING and ScaSpDegree > 0:
                ups += [('SCAFI', P['ScaFi'], C), ('SCAFJ', P['ScaFj'], C), ('SCAFIJ', P['ScaFij'], C)]
            ups += [('FIJAB', P['Fijab'], C), ('NEQ', P['NEQ'], C), ('NEQT', P['NEQt'], C)]
            fitsio.writeto(FITS_Solution, Solution.reshape((-1, 1)).T, base_cards=None, updates=ups)
        return Solution, PixA_DIFF

    @staticmethod
    def BSP_arrays(PixA_REF, PixA_SCI, PixA_mREF, PixA_mSCI, ForceConv='REF', GKerHW=8, CUDA_DEVICE_4SUBTRACT='0',
                   BACKEND_4SUBTRACT='B200', VERBOSE_LEVEL=2, STORAGE='fp64', _return_config=False, **ssc_kwargs):
        """The array-level body of BSP (:4100-4216): NaN-union fill from the masked images, role swap, sign flip."""
        PixA_REF, PixA_SCI = np.asarray(PixA_REF, np.float64), np.asarray(PixA_SCI, np.float64)
        PixA_mREF, PixA_mSCI = np.asarray(PixA_mREF, np.float64), np.asarray(PixA_mSCI, np.float64)
        NaNmask_U = None
        NaNmask_REF, NaNmask_SCI = np.isnan(PixA_REF), np.isnan(PixA_SCI)
        if NaNmask_REF.any() or NaNmask_SCI.any():
            NaNmask_U = np.logical_or(NaNmask_REF, NaNmask_SCI)
        assert np.sum(np.isnan(PixA_mREF)) == 0
        assert np.sum(np.isnan(PixA_mSCI)) == 0
        assert ForceConv in ['REF', 'SCI']
        t0 = time.time()
        SFFTConfig = SingleSFFTConfigure.SSC(NX=PixA_REF.shape[0], NY=PixA_REF.shape[1], KerHW=GKerHW,
                                             BACKEND_4SUBTRACT=BACKEND_4SUBTRACT, VERBOSE_LEVEL=VERBOSE_LEVEL,
                                             CUDA_DEVICE=int(CUDA_DEVICE_4SUBTRACT), STORAGE=STORAGE, **ssc_kwargs)
        if VERBOSE_LEVEL in [1, 2]:
            print('\nMeLOn Report: FUNCTION COMPILATIONS OF SFFT-SUBTRACTION TAKES [%.3f s] \n' % (time.time() - t0))
        if ForceConv == 'REF':
            PixA_mI, PixA_mJ, PixA_I, PixA_J = PixA_mREF, PixA_mSCI, PixA_REF, PixA_SCI
        else:
            PixA_mI, PixA_mJ, PixA_I, PixA_J = PixA_mSCI, PixA_mREF, PixA_SCI, PixA_REF
        if NaNmask_U is not None:
            PixA_I, PixA_J = PixA_I.copy(), PixA_J.copy()
            PixA_I[NaNmask_U] = PixA_mI[NaNmask_U]
            PixA_J[NaNmask_U] = PixA_mJ[NaNmask_U]
        Solution, PixA_DIFF = GeneralSFFTSubtract.GSS(PixA_I=PixA_I, PixA_J=PixA_J, PixA_mI=PixA_mI, PixA_mJ=PixA_mJ,
                                                      SFFTConfig=SFFTConfig, ContamMask_I=None,
                                                      BACKEND_4SUBTRACT=BACKEND_4SUBTRACT, VERBOSE_LEVEL=VERBOSE_LEVEL)[:2]
        if NaNmask_U is not None:
            PixA_DIFF[NaNmask_U] = np.nan
        if ForceConv == 'SCI':
            PixA_DIFF = -PixA_DIFF
        if _return_config:
            return Solution, PixA_DIFF, SFFTConfig
        return Solution, PixA_DIFF


# ---- consumers of the Solution (:4358-4723) and the decorrelation step (:4725-4868) -------------------------------------------
def _solution_header(FITS_Solution):
    cards, _ = fitsio.read_header(FITS_Solution)
    h = fitsio.header_dict(cards)
    sep = str(h['SEPSCA']).strip() == 'True'
    d = dict(KerHW=int(h['KERHW']), KerSpType=str(h['KSPTYPE']).strip(),
             KerIntKnotX=[float(h['KIKX%d' % i]) for i in range(int(h['NKIKX']))],
             KerIntKnotY=[float(h['KIKY%d' % i]) for i in range(int(h['NKIKY']))],
             N0=int(h['N0']), N1=int(h['N1']), DK=int(h['DK']), L0=int(h['L0']), L1=int(h['L1']), Fi=int(h['FI']), Fj=int(h['FJ']),
             Fpq=int(h['FPQ']), SEPARATE_SCALING=sep, ScaSpType=None, DS=None, ScaIntKnotX=None, ScaIntKnotY=None, ScaFi=None, ScaFj=None)
    if sep:
        d.update(ScaSpType=str(h['SSPTYPE']).strip(), DS=int(h['SSPDEG']),
                 ScaIntKnotX=[float(h['SIKX%d' % i]) for i in range(int(h['NSIKX']))],
                 ScaIntKnotY=[float(h['SIKY%d' % i]) for i in range(int(h['NSIKY']))])
        if d['DS'] > 0:
            d.update(ScaFi=int(h['SCAFI']), ScaFj=int(h['SCAFJ']))
    Solution = np.asarray(fitsio.getdata(FITS_Solution), np.float64)[0]
    return Solution, d


class Read_SFFTSolution:
    """(SfftKerDict, SfftScaDict) of a BSplineSFFT Solution (:4358-4553): SfftKerDict[(i, j)][a + w0, b + w1] = ac_ijab in the
    modified-delta basis (ac = a / (N0 N1)); in SEPARATE-VARYING mode the centre taps are NaN there and SfftScaDict[(i, j)]
    holds the coefficients of the scaling basis."""

    def FromArray(self, Solution, KerSpType, N0, N1, DK, L0, L1, Fi, Fj, Fpq, SEPARATE_SCALING, ScaSpType, DS, ScaFi, ScaFj):
        varying = bool(SEPARATE_SCALING) and DS != 0
        w0, w1 = (L0 - 1) // 2, (L1 - 1) // 2
        if KerSpType == 'Polynomial':
            REF_ij = [(i, j) for i in range(DK + 1) for j in range(DK + 1 - i)]
        else:
            REF_ij = [(i, j) for i in range(Fi) for j in range(Fj)]
        Fij = len(REF_ij)
        ac = (np.asarray(Solution, np.float64)[:-Fpq] / (N0 * N1)).reshape(Fij, L0, L1)
        SfftKerDict = {ij: ac[k].copy() for k, ij in enumerate(REF_ij)}
        SfftScaDict = None
        if varying:
            if ScaSpType == 'Polynomial':
                ScaREF_ij = [(i, j) for i in range(DS + 1) for j in range(DS + 1 - i)]
            else:
                ScaREF_ij = [(i, j) for i in range(ScaFi) for j in range(ScaFj)]
            # (the reference initialises the dictionary from KerSpType, :4478-4488; the keys it then fills are these)
            SfftScaDict = {ij: 0.0 for ij in ScaREF_ij}
            for k, ij in enumerate(REF_ij):
                if k < len(ScaREF_ij):
                    SfftScaDict[ScaREF_ij[k]] = float(ac[k, w0, w1])
                SfftKerDict[ij][w0, w1] = np.nan
        return SfftKerDict, SfftScaDict

    def FromFITS(self, FITS_Solution):
        Solution, d = _solution_header(FITS_Solution)
        return self.FromArray(Solution=Solution, KerSpType=d['KerSpType'], N0=d['N0'], N1=d['N1'], DK=d['DK'], L0=d['L0'], L1=d['L1'],
                              Fi=d['Fi'], Fj=d['Fj'], Fpq=d['Fpq'], SEPARATE_SCALING=d['SEPARATE_SCALING'], ScaSpType=d['ScaSpType'],
                              DS=d['DS'], ScaFi=d['ScaFi'], ScaFj=d['ScaFj'])


class BSpline_MatchingKernel:
    """Matching kernels realised at the requested FortranCoor positions XY_q, shape (NPOINT, L0, L1), Cartesian-delta basis
    (:4555-4723)."""

    def __init__(self, XY_q, VERBOSE_LEVEL=2):
        self.XY_q = XY_q
        self.VERBOSE_LEVEL = VERBOSE_LEVEL

    def FromArray(self, Solution, KerSpType, KerIntKnotX, KerIntKnotY, N0, N1, DK, L0, L1, Fi, Fj, Fpq,
                  SEPARATE_SCALING, ScaSpType, ScaIntKnotX, ScaIntKnotY, DS, ScaFi, ScaFj):
        sXY_q = np.asarray(self.XY_q).astype(float)
        sXY_q[:, 0] /= N0
        sXY_q[:, 1] /= N1
        SfftKerDict, SfftScaDict = Read_SFFTSolution().FromArray(
            Solution=Solution, KerSpType=KerSpType, N0=N0, N1=N1, DK=DK, L0=L0, L1=L1, Fi=Fi, Fj=Fj, Fpq=Fpq,
            SEPARATE_SCALING=SEPARATE_SCALING, ScaSpType=ScaSpType, DS=DS, ScaFi=ScaFi, ScaFj=ScaFj)
        w0, w1 = (L0 - 1) // 2, (L1 - 1) // 2
        U, V, fu, fv = _basis_tables(KerSpType, DK, KerIntKnotX, KerIntKnotY, N0, N1, sXY_q[:, 0], sXY_q[:, 1])
        KerBASE = U[fu] * V[fv]                                                        # (Fij, NPOINT)
        KerCOEFF = np.array([SfftKerDict[(int(i), int(j))] for i, j in zip(fu, fv)])   # (Fij, L0, L1)
        KerStack = np.tensordot(KerBASE, KerCOEFF, (0, 0))                             # (NPOINT, L0, L1)
        # modified-delta -> kernel pixels: every non-centre tap also subtracts from the centre (:4619-4660)
        if SfftScaDict is None:
            KerCENT = KerStack[:, w0, w1].copy()
            KerCENT -= np.sum(KerStack, axis=(1, 2)) - KerStack[:, w0, w1]
            KerStack[:, w0, w1] = KerCENT
        else:
            U, V, fu, fv = _basis_tables(ScaSpType, DS, ScaIntKnotX, ScaIntKnotY, N0, N1, sXY_q[:, 0], sXY_q[:, 1])
            ScaBASE = U[fu] * V[fv]
            ScaCOEFF = np.array([SfftScaDict[(int(i), int(j))] for i, j in zip(fu, fv)])
            KerCENT = np.matmul(ScaCOEFF.reshape((1, -1)), ScaBASE)[0]
            KerCENT -= np.nansum(KerStack, axis=(1, 2))
            KerStack[:, w0, w1] = KerCENT
        return KerStack

    def FromFITS(self, FITS_Solution):
        Solution, d = _solution_header(FITS_Solution)
        return self.FromArray(Solution=Solution, KerSpType=d['KerSpType'], KerIntKnotX=d['KerIntKnotX'], KerIntKnotY=d['KerIntKnotY'],
                              N0=d['N0'], N1=d['N1'], DK=d['DK'], L0=d['L0'], L1=d['L1'], Fi=d['Fi'], Fj=d['Fj'], Fpq=d['Fpq'],
                              SEPARATE_SCALING=d['SEPARATE_SCALING'], ScaSpType=d['ScaSpType'], ScaIntKnotX=d['ScaIntKnotX'],
                              ScaIntKnotY=d['ScaIntKnotY'], DS=d['DS'], ScaFi=d['ScaFi'], ScaFj=d['ScaFj'])


class ConvKernel_Convertion:
    """CSZ / iCSZ of :4725-4753 (iCSZ returns the lost weight here, unlike sfft/utils/ConvKernelConvertion.py)."""

    def CSZ(ConvKernel, N0, N1):
        L0, L1 = ConvKernel.shape
        w0, w1 = (L0 - 1) // 2, (L1 - 1) // 2
        TailZP = np.pad(ConvKernel, ((0, N0 - L0), (0, N1 - L1)), 'constant', constant_values=(0, 0))
        return np.roll(np.roll(TailZP, -w0, axis=0), -w1, axis=1)

    def iCSZ(KIMG, L0, L1):
        w0, w1 = (L0 - 1) // 2, (L1 - 1) // 2
        KIMG_iCSZ = np.roll(np.roll(KIMG, w1, axis=1), w0, axis=0)
        ConvKernel = KIMG_iCSZ[:L0, :L1]
        return ConvKernel, 1.0 - np.sum(np.abs(ConvKernel)) / np.sum(np.abs(KIMG_iCSZ))


class BSpline_DeCorrelation:
    @staticmethod
    def BDC(MK_JLst, SkySig_JLst, MK_ILst=[], SkySig_ILst=[], MK_Fin=None, KERatio=2.0, DENO_CLIP_RATIO=100000.0, VERBOSE_LEVEL=2,
            CUDA_DEVICE='0'):
        """DeCorrelation_Calculator.DCC with the denominator clipped from below at max / DENO_CLIP_RATIO (:4757-4868)."""
        from .utils.DeCorrelationCalculator import DeCorrelation_Calculator
        return DeCorrelation_Calculator.DCC(MK_JLst, SkySig_JLst, MK_ILst=MK_ILst, SkySig_ILst=SkySig_ILst, MK_Fin=MK_Fin, KERatio=KERatio,
                                            VERBOSE_LEVEL=VERBOSE_LEVEL, CUDA_DEVICE=CUDA_DEVICE, _CLIP_RATIO=float(DENO_CLIP_RATIO))


class BSpline_GridConvolve:
    """Grid-wise spatially varying convolution (:4870-5010): every pixel is convolved with the kernel of its grid cell
    (`AllocatedL` labels, `KerStack[label]`), zero outside the image, NaN samples replaced by `nan_fill_value`.  The reference
    convolves a mini image per cell on the CPU (astropy) or the GPU (cupyx convolve2d / fftconvolve, `use_fft`); here one CUDA
    kernel evaluates the same sums for all cells (sfftb_convolve_grid) -- `use_fft` is accepted and changes nothing but rounding
    in the reference.  Cells are assumed to be what the reference builds (rectangular tiles, :4876-4893): it pastes bounding boxes."""

    def __init__(self, PixA_obj, AllocatedL, KerStack, nan_fill_value=0.0, use_fft=False, normalize_kernel=True):
        self.PixA_in = np.ascontiguousarray(PixA_obj, np.float64)
        self.nan_fill_value = float(nan_fill_value)
        self.AllocatedL = np.ascontiguousarray(AllocatedL, np.int32)
        self.KerStack = np.ascontiguousarray(KerStack, np.float64)
        self.use_fft = use_fft
        self.normalize_kernel = normalize_kernel
        if self.AllocatedL.shape != self.PixA_in.shape or self.KerStack.ndim != 3:
            raise Exception('MeLOn ERROR: AllocatedL must have the image shape and KerStack the shape (Nseg, L0, L1)')

    def GSVC_GPU(self, CUDA_DEVICE='0', CLEAN_GPU_MEMORY=False, nproc=32):
        from . import _lib as B
        N0, N1 = self.PixA_in.shape
        Nseg, L0, L1 = self.KerStack.shape
        out = np.empty((N0, N1), np.float64)
        B.check(B.lib().sfftb_convolve_grid(int(CUDA_DEVICE), 0, self.PixA_in.ctypes.data, B.F64, N0, N1, self.AllocatedL.ctypes.data, Nseg,
                                            self.KerStack.ctypes.data, L0, L1, self.nan_fill_value, int(bool(self.normalize_kernel)),
                                            out.ctypes.data, B.MEM_HOST))
        return out

    def GSVC_CPU(self, nproc=32):
        """The reference's multiprocessing CPU variant (:4912-4949); this package has no CPU path, the same CUDA routine serves it."""
        return self.GSVC_GPU()

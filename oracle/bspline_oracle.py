"""
CPU oracle of the B-spline SFFT variant (sfft/BSplineSFFT.py) -- TEST INFRASTRUCTURE, not product.

The reference's B-spline subtraction (`BSpline_Packet.BSP` -> `SingleSFFTConfigure.SSC` :2538 ->
`ElementalSFFTSubtract_Cupy.ESSC` :2613) has no CPU backend (:2605-2607), so it cannot run without CuPy.  This
module restates WHAT it computes in the implementation-independent form of SURVEY.md 8c-4, extended to the richer
configuration:

    D = [ SCALE * (roll(I_ij, (a, b)) - [ab != 00] I_ij)  for ij, ab  |  T_pq ]          (N0*N1, NEQ)
    LHMAT = D^T D / N  (+ LAMBDA * REGMAT),   RHb = D^T J / N,   DIFF = J - D @ Solution

with I_ij = I * KerBasis_ij, T_pq = BkgBasis_pq, the bases being tensor products of 1-D B-spline (or total-degree
polynomial) functions of the scaled pixel-centre coordinates, exactly as built by `Create_BSplineBasis`
(:2624-2634) / the `KerSpatial`, `BkgSpatial`, `ScaSpatial` kernels (:276-458).  Scaling modes (:77-86):

  * ENTANGLED          -- nothing else;
  * SEPARATE-CONSTANT  -- the Fij columns (ij, 00) are tied to one unknown: polynomial kernel drops (ij>0, 00)
                          (sfftcore's Remove_LSFStripes), B-spline kernel SUMS the stripes (TweakLS :2202-2272,
                          partition of unity) and copies the value back (:3764-3766);
  * SEPARATE-VARYING   -- the (ij, 00) column of plane ij < ScaFij is SCALE * I * ScaBasis_ij instead
                          (Construct_FDIFF :2487-2495), the planes ij >= ScaFij have no (00) unknown (:3733-3747).

Regulariser (:3570-3700): REGMAT[(k,c),(k',c')] = SCALE^2 * SST[k,k'] * iREG[c,c'] with SST the (weighted) Gram
matrix of the kernel basis at the requested coordinates and iREG = 2 M^T (Lap^T Lap) M, M the change of basis from
the modified-delta coefficients to kernel pixels; SEPARATE-VARYING swaps in the scaling basis on the centre tap.

Pinning: ENTANGLED and SEPARATE-CONSTANT, polynomial and B-spline kernels / backgrounds, are checked against the
executable NumPy backend of the reference's own development copy (misc/beta4spline/new_version_sfftcore, run by
tests/golden/make_golden_bspline.py; fixtures tests/golden/bspline_*.npz).  SEPARATE-VARYING and the regulariser
have no executable reference in this container: for them parity is UNPINNED (restated from the cited lines only).

Dense D: intended for sizes up to ~256^2.
"""
import numpy as np
from scipy.interpolate import BSpline
from scipy import signal

__all__ = ['create_bspline_basis', 'ssc_params', 'design_matrix', 'regularizer', 'ess', 'gss']


def create_bspline_basis(N, IntKnot, Degree, ReqCoord=None):
    """Create_BSplineBasis / Create_BSplineBasis_Req (BSplineSFFT.py:2624-2646): clamped knot vector
    [0.5]*(k+1) ++ IntKnot ++ [N+0.5]*(k+1), divided by N, evaluated at the pixel centres (1 + arange(N)) / N."""
    coord = (1.0 + np.arange(N)) / N if ReqCoord is None else np.asarray(ReqCoord, float)
    knot = np.concatenate(([0.5] * (Degree + 1), IntKnot, [N + 0.5] * (Degree + 1))) / N
    Nc = len(IntKnot) + Degree + 1
    out = []
    for idx in range(Nc):
        c = (np.arange(Nc) == idx).astype(float)
        out.append(BSpline(t=knot, c=c, k=Degree, extrapolate=False)(coord))
    return np.array(out)


def _dof(SpType, Degree, KnotX, KnotY):
    if SpType == 'Polynomial':
        return -1, -1, ((Degree + 1) * (Degree + 2)) // 2
    Fi, Fj = len(KnotX) + Degree + 1, len(KnotY) + Degree + 1
    return Fi, Fj, Fi * Fj


def ssc_params(NX, NY, KerHW=8, KerSpType='Polynomial', KerSpDegree=2, KerIntKnotX=(), KerIntKnotY=(),
               SEPARATE_SCALING=True, ScaSpType='Polynomial', ScaSpDegree=0, ScaIntKnotX=(), ScaIntKnotY=(),
               BkgSpType='Polynomial', BkgSpDegree=2, BkgIntKnotX=(), BkgIntKnotY=(),
               REGULARIZE_KERNEL=False, IGNORE_LAPLACIAN_KERCENT=True, XY_REGULARIZE=None, WEIGHT_REGULARIZE=None,
               LAMBDA_REGULARIZE=1e-6):
    """SFFTParam_dict of SingleSFFTConfigure_Cupy.SSCC (BSplineSFFT.py:26-273)."""
    N0, N1, w0, w1 = int(NX), int(NY), int(KerHW), int(KerHW)
    DK, DB = int(KerSpDegree), int(BkgSpDegree)
    assert DK >= 0 and DB >= 0 and KerSpType in ('Polynomial', 'B-Spline') and BkgSpType in ('Polynomial', 'B-Spline')
    if KerSpType == 'B-Spline' and DK == 0:
        assert len(KerIntKnotX) == 0 and len(KerIntKnotY) == 0                         # :36-38
    if BkgSpType == 'B-Spline' and DB == 0:
        assert len(BkgIntKnotX) == 0 and len(BkgIntKnotY) == 0
    if not SEPARATE_SCALING:
        MODE = 'ENTANGLED'
    elif int(ScaSpDegree) == 0:
        MODE = 'SEPARATE-CONSTANT'
    else:
        MODE = 'SEPARATE-VARYING'
    L0, L1 = 2 * w0 + 1, 2 * w1 + 1
    Fab = L0 * L1
    Fi, Fj, Fij = _dof(KerSpType, DK, KerIntKnotX, KerIntKnotY)
    Fp, Fq, Fpq = _dof(BkgSpType, DB, BkgIntKnotX, BkgIntKnotY)
    P = dict(N0=N0, N1=N1, w0=w0, w1=w1, DK=DK, DB=DB, KerHW=KerHW, L0=L0, L1=L1, Fab=Fab, Fi=Fi, Fj=Fj, Fij=Fij,
             Fp=Fp, Fq=Fq, Fpq=Fpq, Fijab=Fij * Fab, NEQ=Fij * Fab + Fpq, SCALE=1.0 / (N0 * N1), SCALE_L=float(N0 * N1),
             KerSpType=KerSpType, KerSpDegree=DK, KerIntKnotX=list(KerIntKnotX), KerIntKnotY=list(KerIntKnotY),
             BkgSpType=BkgSpType, BkgSpDegree=DB, BkgIntKnotX=list(BkgIntKnotX), BkgIntKnotY=list(BkgIntKnotY),
             SEPARATE_SCALING=SEPARATE_SCALING, SCALING_MODE=MODE, REGULARIZE_KERNEL=REGULARIZE_KERNEL,
             IGNORE_LAPLACIAN_KERCENT=IGNORE_LAPLACIAN_KERCENT, XY_REGULARIZE=XY_REGULARIZE,
             WEIGHT_REGULARIZE=WEIGHT_REGULARIZE, LAMBDA_REGULARIZE=LAMBDA_REGULARIZE)
    P['NEQt'] = P['NEQ']
    if MODE == 'SEPARATE-CONSTANT':
        P['NEQt'] = P['NEQ'] - Fij + 1                                                  # :199-200
    if SEPARATE_SCALING:
        DS = int(ScaSpDegree)
        P.update(DS=DS, ScaSpType=ScaSpType, ScaSpDegree=DS, ScaIntKnotX=list(ScaIntKnotX), ScaIntKnotY=list(ScaIntKnotY))
    if MODE == 'SEPARATE-VARYING':
        ScaFi, ScaFj, ScaFij = _dof(ScaSpType, DS, ScaIntKnotX, ScaIntKnotY)
        assert ScaFij <= Fij                                                            # :190
        P.update(ScaFi=ScaFi, ScaFj=ScaFj, ScaFij=ScaFij)
        P['NEQt'] = P['NEQ'] - Fij + ScaFij                                             # :201-202
    return P


def _basis_planes(SpType, Degree, KnotX, KnotY, N0, N1, CX=None, CY=None):
    """(F, n0, n1) tensor-product basis: on the pixel grid (CX = CY = None) or, for the regulariser, the values at the
    requested scaled coordinates (returned as (F, NREG))."""
    grid = CX is None
    if SpType == 'Polynomial':
        cx = (1.0 + np.arange(N0)) / N0 if grid else CX
        cy = (1.0 + np.arange(N1)) / N1 if grid else CY
        ij = [(i, j) for i in range(Degree + 1) for j in range(Degree + 1 - i)]
        if grid:
            return np.array([np.outer(cx ** i, cy ** j) for (i, j) in ij])
        return np.array([cx ** i * cy ** j for (i, j) in ij])
    BX = create_bspline_basis(N0, KnotX, Degree, None if grid else CX)
    BY = create_bspline_basis(N1, KnotY, Degree, None if grid else CY)
    ij = [(i, j) for i in range(BX.shape[0]) for j in range(BY.shape[0])]
    if grid:
        return np.array([np.outer(BX[i], BY[j]) for (i, j) in ij])
    return np.array([BX[i] * BY[j] for (i, j) in ij])


def _ij00(P):
    return np.arange(P['w0'] * P['L1'] + P['w1'], P['Fijab'], P['Fab'])                # :2805


def design_matrix(PixA_I, P):
    """D (N0*N1, NEQ): the columns the Greeks of ESSC are inner products of (:3013-3565), see the module docstring."""
    N0, N1, w0, w1, L1, Fab = P['N0'], P['N1'], P['w0'], P['w1'], P['L1'], P['Fab']
    I = np.asarray(PixA_I, np.float64)
    KB = _basis_planes(P['KerSpType'], P['DK'], P['KerIntKnotX'], P['KerIntKnotY'], N0, N1)
    TB = _basis_planes(P['BkgSpType'], P['DB'], P['BkgIntKnotX'], P['BkgIntKnotY'], N0, N1)
    SB = None
    if P['SCALING_MODE'] == 'SEPARATE-VARYING':
        SB = _basis_planes(P['ScaSpType'], P['DS'], P['ScaIntKnotX'], P['ScaIntKnotY'], N0, N1)
    SCALE = P['SCALE']
    D = np.empty((N0 * N1, P['NEQ']))
    for ij in range(P['Fij']):
        Iij = I * KB[ij]
        for a in range(-w0, w0 + 1):
            for b in range(-w1, w1 + 1):
                col = ij * Fab + (a + w0) * L1 + (b + w1)
                if a == 0 and b == 0:
                    if SB is None:
                        c = Iij
                    else:
                        c = I * SB[ij] if ij < P['ScaFij'] else np.zeros_like(I)       # ScaREF (-1,-1) placeholder :2783
                else:
                    c = np.roll(Iij, (a, b), axis=(0, 1)) - Iij
                D[:, col] = SCALE * c.ravel()
    for pq in range(P['Fpq']):
        D[:, P['Fijab'] + pq] = TB[pq].ravel()
    return D


def regularizer(P):
    """REGMAT (NEQ, NEQ) of :3570-3697 (LHMAT += LAMBDA_REGULARIZE * REGMAT)."""
    N0, N1, w0, w1, L0, L1, Fab, Fij = P['N0'], P['N1'], P['w0'], P['w1'], P['L0'], P['L1'], P['Fab'], P['Fij']
    XY = np.asarray(P['XY_REGULARIZE'], float)
    NREG = XY.shape[0]
    CX, CY = XY[:, 0] / N0, XY[:, 1] / N1
    SP = _basis_planes(P['KerSpType'], P['DK'], P['KerIntKnotX'], P['KerIntKnotY'], N0, N1, CX, CY)   # (Fij, NREG)
    if P['WEIGHT_REGULARIZE'] is None:
        Wd = np.full(NREG, 1.0 / NREG)
    else:
        Wd = np.asarray(P['WEIGHT_REGULARIZE'], float) / np.sum(P['WEIGHT_REGULARIZE'])
    SST = (SP * Wd) @ SP.T
    varying = P['SCALING_MODE'] == 'SEPARATE-VARYING'
    if varying:
        ScaSP = _basis_planes(P['ScaSpType'], P['DS'], P['ScaIntKnotX'], P['ScaIntKnotY'], N0, N1, CX, CY)
        ScaSP = np.concatenate([ScaSP, np.zeros((Fij - P['ScaFij'], NREG))], axis=0)   # placeholder rows :3614-3620
        CSST = (SP * Wd) @ ScaSP.T
        DSST = (ScaSP * Wd) @ ScaSP.T
    # Laplacian on the kernel stamp: neighbour count on the diagonal, -1 for the 4-neighbours (:3641-3666)
    LAP = np.zeros((Fab, Fab))
    ad = signal.correlate2d(np.ones((L0, L1)), np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]]), mode='same',
                            boundary='fill', fillvalue=0)
    LAP[np.arange(Fab), np.arange(Fab)] = ad.ravel()
    for r in range(L0):
        for c in range(L1):
            for (dr, dc) in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                r2, c2 = r + dr, c + dc
                if 0 <= r2 < L0 and 0 <= c2 < L1:
                    LAP[r * L1 + c, r2 * L1 + c2] = -1.0
    c0 = w0 * L1 + w1
    if P['IGNORE_LAPLACIAN_KERCENT']:
        for row in ((w0 - 1) * L1 + w1, w0 * L1 + w1 - 1, c0, w0 * L1 + w1 + 1, (w0 + 1) * L1 + w1):   # :3670-3676
            LAP[row, :] = 0.0
    LTL = LAP.T @ LAP
    # iREGMAT (:2051-2086) = 2 M^T LTL M, M: modified-delta coefficients -> kernel pixels
    M = np.eye(Fab)
    M[c0, :] = -1.0
    M[c0, c0] = 1.0
    iREG = 2.0 * M.T @ LTL @ M
    REG = np.zeros((P['NEQ'], P['NEQ']))
    S2 = P['SCALE'] ** 2
    K = np.kron(SST, iREG)
    if varying:                                                                        # :2122-2166
        cen = np.zeros(Fab, bool)
        cen[c0] = True
        mC = np.kron(np.ones((Fij, Fij)), np.outer(~cen, cen)).astype(bool)            # c != c0, c' == c0
        mR = np.kron(np.ones((Fij, Fij)), np.outer(cen, ~cen)).astype(bool)
        mD = np.kron(np.ones((Fij, Fij)), np.outer(cen, cen)).astype(bool)
        K = np.where(mC, np.kron(CSST, iREG), K)
        K = np.where(mR, np.kron(CSST.T, iREG), K)
        K = np.where(mD, np.kron(DSST, iREG), K)
    REG[:P['Fijab'], :P['Fijab']] = S2 * K
    return REG


def _tweak(P, L, b):
    """Tweak of the linear system for the scaling mode (:3702-3747); returns (L_t, b_t, restore)."""
    MODE, NEQ, Fij = P['SCALING_MODE'], P['NEQ'], P['Fij']
    ij00 = _ij00(P)
    if MODE == 'ENTANGLED' or (MODE == 'SEPARATE-VARYING' and P['NEQt'] == NEQ):
        return L, b, lambda x: x
    if MODE == 'SEPARATE-CONSTANT':
        pres = np.setdiff1d(np.arange(NEQ), ij00[1:])
        if P['KerSpType'] == 'Polynomial':
            def restore(xt):
                x = np.zeros(NEQ)
                x[pres] = xt
                return x
            return L[np.ix_(pres, pres)], b[pres], restore
        # B-spline kernel: sum the stripes into ij00[0] (rows and columns)
        T = np.zeros((NEQ, len(pres)))
        T[pres, np.arange(len(pres))] = 1.0
        key = int(np.searchsorted(pres, ij00[0]))
        T[ij00[1:], key] = 1.0

        def restore(xt):
            return T @ xt                                                              # :3764-3771
        return T.T @ L @ T, T.T @ b, restore
    pres = np.setdiff1d(np.arange(NEQ), ij00[P['ScaFij']:])                            # SEPARATE-VARYING, NEQt < NEQ

    def restore(xt):
        x = np.zeros(NEQ)
        x[pres] = xt
        return x
    return L[np.ix_(pres, pres)], b[pres], restore


def ess(PixA_I, PixA_J, P, SFFTSolution=None, Subtract=False, export=None):
    """ElementalSFFTSubtract_Cupy.ESSC (:2613-3850) in design-matrix form."""
    I, J = np.asarray(PixA_I, np.float64), np.asarray(PixA_J, np.float64)
    assert I.shape == (P['N0'], P['N1']) and J.shape == I.shape
    N = P['N0'] * P['N1']
    D = design_matrix(I, P)
    if SFFTSolution is not None:
        Solution = np.asarray(SFFTSolution, np.float64)
    else:
        L = D.T @ D / N
        b = D.T @ J.ravel() / N
        if P['REGULARIZE_KERNEL']:
            L = L + P['LAMBDA_REGULARIZE'] * regularizer(P)
        Lt, bt, restore = _tweak(P, L, b)
        if export is not None:
            export.update(LHMAT=L, RHb=b, LHMAT_tweaked=Lt, RHb_tweaked=bt)
        Solution = restore(np.linalg.solve(Lt, bt))
    DIFF = None
    if Subtract:
        DIFF = J - (D @ Solution).reshape(J.shape)
    return Solution, DIFF


def gss(PixA_I, PixA_J, PixA_mI, PixA_mJ, P):
    """GeneralSFFTSubtract.GSS (BSplineSFFT.py:3882-3966): fit on the masked pair, subtract the unmasked pair."""
    Solution = ess(PixA_mI, PixA_mJ, P, None, False)[0]
    DIFF = ess(PixA_I, PixA_J, P, Solution, True)[1]
    return Solution, DIFF

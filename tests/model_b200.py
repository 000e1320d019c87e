"""
NumPy model of the B200 algorithm (TEST INFRASTRUCTURE, not product).

It restates, in vectorised NumPy, exactly the arithmetic the CUDA kernels in
sfft_b200/csrc perform (row R2C spectra stored transposed, DIF-folded column
slices, cross-spectrum -> pruned inverse -> lag tables, polynomial moments with
explicit wrap corrections, normal-equation fill, kernel-spectrum apply), so that
every formula can be pinned against oracle/sfft_oracle.py on the CPU before it is
written in CUDA.  tests/test_model_vs_oracle.py runs it.
"""
import numpy as np
from math import comb


def index_tables(DK, DB, w0, w1):
    REF_ij = [(i, j) for i in range(DK + 1) for j in range(DK + 1 - i)]
    REF_pq = [(p, q) for p in range(DB + 1) for q in range(DB + 1 - p)]
    REF_ab = [(a, b) for a in range(-w0, w0 + 1) for b in range(-w1, w1 + 1)]
    return REF_ij, REF_pq, REF_ab


def row_spectra(img, DK):
    """g_j[r, k1] = sum_c img[r, c] * cy(c)^j * exp(-2 pi i k1 c / N1), j = 0..DK; half spectrum."""
    N0, N1 = img.shape
    cy = (np.arange(N1) + 1.0) / N1
    return np.stack([np.fft.rfft(img * cy[None, :] ** j, axis=1) for j in range(DK + 1)])  # (DK+1, N0, NH)


def fold_slices(col, V):
    """DIF fold of one column (length N0) into V slices of length M = N0 / V:
    z_t[n] = exp(-2 pi i t n / N0) * sum_v exp(-2 pi i t v / V) col[n + M v];  FFT_M(z_t)[u] = F[V u + t]."""
    N0 = col.shape[-1]
    M = N0 // V
    c = col.reshape(col.shape[:-1] + (V, M))                       # [v, n]
    t = np.arange(V)
    Wv = np.exp(-2j * np.pi * np.outer(t, np.arange(V)) / V)       # [t, v]
    z = np.einsum('tv,...vn->...tn', Wv, c)
    z = z * np.exp(-2j * np.pi * np.outer(t, np.arange(M)) / N0)
    return z                                                       # [..., t, n]


def kappa_column(gA, gB, lags, V):
    """kappa[m0] = sum_r conj(gA[r]) gB[(r + m0) % N0] via folded FFT slices (what the fit kernel does)."""
    N0 = gA.shape[0]
    M = N0 // V
    FA = np.fft.fft(fold_slices(gA, V), axis=-1)                   # [t, u] = F[V u + t]
    FB = np.fft.fft(fold_slices(gB, V), axis=-1)
    X = np.conj(FA) * FB
    y = np.fft.ifft(X, axis=-1) * M                                # unnormalised inverse: sum_u X e^{+2 pi i u m / M}
    out = np.zeros(len(lags), complex)
    for t in range(V):
        out += np.exp(2j * np.pi * t * lags / N0) * y[t, np.mod(lags, M)]
    return out / N0


def lag_tables(gI, gJ, REF_ij, w0, w1, V):
    """R_AB[m0, m1] for all unordered I-plane pairs (|m| <= 2w) and R_AJ (|m| <= w).
    gI: (DK+1, N0, NH) row spectra of I*cy^j;  gJ: (N0, NH)."""
    _, N0, NH = gI.shape
    N1 = 2 * (NH - 1) if True else None
    cx = (np.arange(N0) + 1.0) / N0
    Fij = len(REF_ij)
    l0 = np.arange(-2 * w0, 2 * w0 + 1)
    l1 = np.arange(-2 * w1, 2 * w1 + 1)
    pairs = [(A, B) for A in range(Fij) for B in range(A, Fij)]
    kap = np.zeros((len(pairs), len(l0), NH), complex)
    kapJ = np.zeros((Fij, 2 * w0 + 1, NH), complex)
    lj = np.arange(-w0, w0 + 1)
    for k1 in range(NH):
        cols = [cx ** i * gI[j, :, k1] for (i, j) in REF_ij]
        for p, (A, B) in enumerate(pairs):
            kap[p, :, k1] = kappa_column(cols[A], cols[B], l0, V)
        for A in range(Fij):
            kapJ[A, :, k1] = kappa_column(cols[A], gJ[:, k1], lj, V)
    return pairs, kap, kapJ


def axis1_reduce(kap, N1, l1):
    """R[m0, m1] = (1/N1) sum_{k1 half} wt(k1) Re(kap[m0, k1] e^{+2 pi i k1 m1 / N1})."""
    NH = kap.shape[-1]
    wt = np.full(NH, 2.0)
    wt[0] = 1.0
    if N1 % 2 == 0:
        wt[-1] = 1.0
    E = np.exp(2j * np.pi * np.outer(np.arange(NH), l1) / N1)      # [k1, m1]
    return np.real(np.einsum('...k,km->...m', kap * wt, E)) / N1


def poly_moment_tables(gI, gJ, REF_ij, REF_pq, DK, DB, w0, w1, N1):
    """R_{A,T}[m0, m1] = sum_x I_A[x] T_pq[x + m]  (|m| <= w) and  sum_x J T_pq, from the row spectra:
    column moments nu_{e,j}[k1] = sum_r cx^e g_j[r, k1] + explicit wrap rows, then an axis-1 reduction
    against Q_q[k1] = DFT(cy^q)."""
    _, N0, NH = gI.shape
    cx = (np.arange(N0) + 1.0) / N0
    cy = (np.arange(N1) + 1.0) / N1
    Q = np.stack([np.fft.rfft(cy ** q) for q in range(DB + 1)])    # (DB+1, NH)
    a_l = np.arange(-w0, w0 + 1)
    b_l = np.arange(-w1, w1 + 1)
    Fij, Fpq = len(REF_ij), len(REF_pq)
    # lam[(i,j), p, a][k1] = sum_r cx(r)^i cx((r+a)%N0)^p g_j[r,k1]
    nu = np.stack([[(cx[:, None] ** e * gI[j]).sum(0) for e in range(DK + DB + 1)] for j in range(DK + 1)])  # [j, e, k1]
    lam = np.zeros((Fij, DB + 1, len(a_l), NH), complex)
    for A, (i, j) in enumerate(REF_ij):
        for p in range(DB + 1):
            for ia, a in enumerate(a_l):
                beta = a / N0
                v = sum(comb(p, e) * beta ** (p - e) * nu[j, i + e] for e in range(p + 1))
                if a > 0:
                    rows = np.arange(N0 - a, N0)
                    corr = (cx[rows] + beta - 1.0) ** p - (cx[rows] + beta) ** p
                elif a < 0:
                    rows = np.arange(0, -a)
                    corr = (cx[rows] + beta + 1.0) ** p - (cx[rows] + beta) ** p
                else:
                    rows = np.arange(0)
                    corr = np.zeros(0)
                v = v + ((cx[rows] ** i * corr)[:, None] * gI[j][rows]).sum(0)
                lam[A, p, ia] = v
    wt = np.full(NH, 2.0)
    wt[0] = 1.0
    if N1 % 2 == 0:
        wt[-1] = 1.0
    k1 = np.arange(NH)
    RT = np.zeros((Fij, Fpq, len(a_l), len(b_l)))
    for pq, (p, q) in enumerate(REF_pq):
        for ib, b in enumerate(b_l):
            ph = np.conj(Q[q]) * np.exp(-2j * np.pi * k1 * b / N1) * wt
            RT[:, pq, :, ib] = np.real(lam[:, p] * ph).sum(-1) / N1
    # Delta: sum_x J T_pq
    nuJ = np.stack([(cx[:, None] ** e * gJ).sum(0) for e in range(DB + 1)])
    RJT = np.array([np.real(nuJ[p] * np.conj(Q[q]) * wt).sum() / N1 for (p, q) in REF_pq])
    # Phi: sum_x T_p'q' T_pq (analytic power sums)
    PHI = np.array([[(cx ** (p8 + p)).sum() * (cy ** (q8 + q)).sum() for (p, q) in REF_pq] for (p8, q8) in REF_pq])
    return RT, RJT, PHI


def fill_normal_eq(pairs, R, RJ, RT, RJT, PHI, REF_ij, REF_pq, REF_ab, w0, w1, N):
    """LHMAT, RHb from the lag tables (restating FillLS_*: SFFTConfigure.py:198-275, 329-377, 431-479,
    532-560, 590-634, 665-688, through the identity LHMAT = D^T D / N)."""
    Fij, Fpq, Fab = len(REF_ij), len(REF_pq), len(REF_ab)
    Fijab = Fij * Fab
    NEQ = Fijab + Fpq
    L = np.zeros((NEQ, NEQ))
    b = np.zeros(NEQ)
    pidx = {pr: k for k, pr in enumerate(pairs)}
    ab = np.array(REF_ab)
    nz = (ab[:, 0] != 0) | (ab[:, 1] != 0)

    def Rab(A, B, m0, m1):
        if A <= B:
            return R[pidx[(A, B)]][m0 + 2 * w0, m1 + 2 * w1]
        return R[pidx[(B, A)]][-m0 + 2 * w0, -m1 + 2 * w1]
    a8, b8 = np.meshgrid(ab[:, 0], ab[:, 1], indexing='ij')  # unused helper
    A0 = ab[:, 0][:, None]
    B0 = ab[:, 1][:, None]
    A1 = ab[:, 0][None, :]
    B1 = ab[:, 1][None, :]
    d8 = nz[:, None].astype(float)
    d = nz[None, :].astype(float)
    for A in range(Fij):
        for B in range(Fij):
            blk = (Rab(A, B, A0 - A1, B0 - B1) - d * Rab(A, B, A0 + 0 * A1, B0 + 0 * B1)
                   - d8 * Rab(A, B, -A1 + 0 * A0, -B1 + 0 * B0) + d8 * d * Rab(A, B, 0, 0))
            L[A * Fab:(A + 1) * Fab, B * Fab:(B + 1) * Fab] = blk / N ** 3
        for pq in range(Fpq):
            S = RT[A, pq][ab[:, 0] + w0, ab[:, 1] + w1]
            S0 = RT[A, pq][w0, w1]
            col = (S - nz * S0) / N ** 2
            L[A * Fab:(A + 1) * Fab, Fijab + pq] = col
            L[Fijab + pq, A * Fab:(A + 1) * Fab] = col
        T = RJ[A][ab[:, 0] + w0, ab[:, 1] + w1]
        b[A * Fab:(A + 1) * Fab] = (T - nz * RJ[A][w0, w1]) / N ** 2
    L[Fijab:, Fijab:] = PHI / N
    b[Fijab:] = RJT / N
    return L, b


def apply_solution(I, J, sol, DK, DB, w0, w1, V):
    """DIFF = J - sum_ij conv(I_ij, K_ij) - sum_pq b_pq T_pq with the column pass done in folded slices."""
    N0, N1 = I.shape
    NH = N1 // 2 + 1
    N = N0 * N1
    REF_ij, REF_pq, REF_ab = index_tables(DK, DB, w0, w1)
    Fij, Fab = len(REF_ij), len(REF_ab)
    L0, L1 = 2 * w0 + 1, 2 * w1 + 1
    M = N0 // V
    cx = (np.arange(N0) + 1.0) / N0
    cy = (np.arange(N1) + 1.0) / N1
    gI = row_spectra(I, DK)
    gJ = np.fft.rfft(J, axis=1)
    a_l = np.arange(-w0, w0 + 1)
    b_l = np.arange(-w1, w1 + 1)
    Dt = np.zeros((N0, NH), complex)
    for k1 in range(NH):
        acc = np.zeros((V, M), complex)                     # FDIFF slices [t, u] -> k0 = V u + t
        FJ = np.fft.fft(fold_slices(gJ[:, k1], V), axis=-1)
        acc += FJ
        for A, (i, j) in enumerate(REF_ij):
            a2 = sol[A * Fab:(A + 1) * Fab].reshape(L0, L1)
            h = (a2 * np.exp(-2j * np.pi * k1 * b_l / N1)[None, :]).sum(1)      # h[a]
            const = a2.sum() - a2[w0, w1]
            FA = np.fft.fft(fold_slices(cx ** i * gI[j, :, k1], V), axis=-1)
            # kernel spectrum on slice t: Kf[V u + t] = sum_a h[a] e^{-2 pi i a (V u + t) / N0}
            sp = np.zeros((V, M), complex)
            for t in range(V):
                buf = np.zeros(M, complex)
                np.add.at(buf, np.mod(a_l, M), h * np.exp(-2j * np.pi * a_l * t / N0))
                sp[t] = np.fft.fft(buf)
            acc -= FA * (sp - const) / N
        # inverse column transform of the folded slices (DIT unfold)
        e = np.fft.ifft(acc, axis=-1) * M                                         # e_t[n]
        n = np.arange(M)
        for v in range(V):
            r = n + M * v
            Dt[r, k1] = sum(np.exp(2j * np.pi * t * r / N0) * e[t] for t in range(V))
    diff = np.fft.irfft(Dt, n=N1, axis=1) * N1 / N
    bpq = sol[Fij * Fab:]
    for pq, (p, q) in enumerate(REF_pq):
        diff -= bpq[pq] * cx[:, None] ** p * cy[None, :] ** q
    return diff


def fit_normal_eq(I, J, DK, DB, w0, w1, V):
    N0, N1 = I.shape
    REF_ij, REF_pq, REF_ab = index_tables(DK, DB, w0, w1)
    gI = row_spectra(I, DK)
    gJ = np.fft.rfft(J, axis=1)
    pairs, kap, kapJ = lag_tables(gI, gJ, REF_ij, w0, w1, V)
    R = axis1_reduce(kap, N1, np.arange(-2 * w1, 2 * w1 + 1))
    RJ = axis1_reduce(kapJ, N1, np.arange(-w1, w1 + 1))
    RT, RJT, PHI = poly_moment_tables(gI, gJ, REF_ij, REF_pq, DK, DB, w0, w1, N1)
    return fill_normal_eq(pairs, R, RJ, RT, RJT, PHI, REF_ij, REF_pq, REF_ab, w0, w1, N0 * N1)

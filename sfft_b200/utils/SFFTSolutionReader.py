"""
Read_SFFTSolution / SVKDict_ST2SFFT / SVKDict_SFFT2ST / Realize_MatchingKernel / Realize_FluxScaling -- the consumers
of the Solution vector that every Easy*Packet calls right after GSS (sfft/EasySparsePacket.py:417-436), with the
signatures of sfft/utils/SFFTSolutionReader.py:9-196.

Layout of the Solution (SFFTSubtract.py:61-90): a_ijab at [ij * Fab + (a + w0) * L1 + (b + w1)] with (i, j) enumerated
as `for i in 0..DK for j in 0..DK-i`, followed by the Fpq background coefficients.  The SFFT dictionary holds
ac = a / (N0 N1) in the modified-delta basis; the 'standard' dictionary is the same kernel in the Cartesian-delta
basis, which only changes the centre tap: s[0,0] = 2 ac[0,0] - sum_ab ac[a,b]  (:102-114).

`FromArray` / `FromFITS` are host (NumPy) routines for host Solutions, as in the reference.  `FromDevice` evaluates the
same quantities with the CUDA kernel `realize_kernel` through `sfftb_realize` on a Solution that is still on the GPU
(the PureCupy-style path), returning torch CUDA tensors without a device-to-host copy.
"""
import numpy as np

from .. import fitsio

__all__ = ['Read_SFFTSolution', 'SVKDict_ST2SFFT', 'SVKDict_SFFT2ST', 'Realize_MatchingKernel', 'Realize_FluxScaling']


def _ij_list(DK):
    return [(i, j) for i in range(DK + 1) for j in range(DK + 1 - i)]


def _blocks(Solution, N0, N1, L0, L1, DK, Fpq):
    """(Fij, L0, L1) array of ac_ijab."""
    Fij = len(_ij_list(DK))
    a = np.asarray(Solution, np.float64)[:-Fpq]
    if a.size != Fij * L0 * L1:
        raise Exception('MeLOn ERROR: Solution of length %d does not match Fij*L0*L1 + Fpq = %d' % (
            np.asarray(Solution).size, Fij * L0 * L1 + Fpq))
    return a.reshape(Fij, L0, L1) / (N0 * N1)


def _header_of(FITS_Solution):
    cards, _ = fitsio.read_header(FITS_Solution)
    h = fitsio.header_dict(cards)
    keys = dict(N0=int(h['N0']), N1=int(h['N1']), L0=int(h['L0']), L1=int(h['L1']), DK=int(h['DK']), Fpq=int(h['FPQ']))
    Solution = np.asarray(fitsio.getdata(FITS_Solution), np.float64)[0]
    return Solution, keys


def _scaled(XY_q, N0, N1):
    s = np.array(XY_q, dtype=float)             # FortranCoor in, ScaledFortranCoor out (copy: the request is kept)
    s[:, 0] /= N0
    s[:, 1] /= N1
    return s


class Read_SFFTSolution:
    def FromArray(self, Solution, N0, N1, L0, L1, DK, Fpq):
        blk = _blocks(Solution, N0, N1, L0, L1, DK, Fpq)
        return {ij: blk[k].copy() for k, ij in enumerate(_ij_list(DK))}

    def FromFITS(self, FITS_Solution):
        Solution, k = _header_of(FITS_Solution)
        return self.FromArray(Solution=Solution, **k)


class SVKDict_ST2SFFT:
    @staticmethod
    def convert(DKx, DKy, Standard_dict):
        out = {k: np.array(v, dtype=float) for k, v in Standard_dict.items()}
        for (i, j) in [(i, j) for i in range(DKx + 1) for j in range(DKy + 1 - i)]:
            L0, L1 = out[(i, j)].shape
            out[(i, j)][(L0 - 1) // 2, (L1 - 1) // 2] = np.sum(Standard_dict[(i, j)])
        return out


class SVKDict_SFFT2ST:
    @staticmethod
    def convert(DKx, DKy, Sfft_dict):
        out = {k: np.array(v, dtype=float) for k, v in Sfft_dict.items()}
        for (i, j) in [(i, j) for i in range(DKx + 1) for j in range(DKy + 1 - i)]:
            L0, L1 = out[(i, j)].shape
            c = ((L0 - 1) // 2, (L1 - 1) // 2)
            out[(i, j)][c] = 2.0 * Sfft_dict[(i, j)][c] - np.sum(Sfft_dict[(i, j)])
        return out


def _device_realize(XY_q, Solution_GPU, SFFTConfig, want_kernel, want_fscal):
    import torch
    plan = SFFTConfig[1]['plan']
    from .. import _lib as B
    dev = torch.device('cuda', plan.device)
    sol = Solution_GPU if isinstance(Solution_GPU, torch.Tensor) else torch.as_tensor(Solution_GPU, device=dev)
    assert sol.is_cuda and sol.dtype == torch.float64 and sol.is_contiguous() and sol.numel() == plan.NEQ
    xy = torch.as_tensor(np.ascontiguousarray(np.asarray(XY_q, np.float64)) if not isinstance(XY_q, torch.Tensor) else XY_q)
    xy = xy.to(device=dev, dtype=torch.float64).contiguous()
    nq = int(xy.shape[0])
    d = plan.dims
    ker = torch.empty((nq, d['L0'], d['L1']), dtype=torch.float64, device=dev) if want_kernel else None
    fs = torch.empty(nq, dtype=torch.float64, device=dev) if want_fscal else None
    plan.bind_torch_stream(torch.cuda.current_stream(dev))
    B.check(B.lib().sfftb_realize(plan._h, sol.data_ptr(), B.MEM_DEVICE, xy.data_ptr(), B.MEM_DEVICE, nq,
                                  ker.data_ptr() if want_kernel else None, fs.data_ptr() if want_fscal else None,
                                  B.MEM_DEVICE))
    return ker, fs


class Realize_MatchingKernel:
    def __init__(self, XY_q):
        self.XY_q = XY_q

    def FromArray(self, Solution, N0, N1, L0, L1, DK, Fpq):
        """(Num_request, L0, L1) matching kernels at the requested FortranCoor coordinates (:121-141)."""
        s = _scaled(self.XY_q, N0, N1)
        blk = _blocks(Solution, N0, N1, L0, L1, DK, Fpq)
        std = blk.copy()
        w0, w1 = (L0 - 1) // 2, (L1 - 1) // 2
        std[:, w0, w1] = 2.0 * blk[:, w0, w1] - blk.sum(axis=(1, 2))
        Bq = np.array([s[:, 0] ** i * s[:, 1] ** j for (i, j) in _ij_list(DK)])
        return np.tensordot(Bq, std, (0, 0))

    def FromFITS(self, FITS_Solution):
        Solution, k = _header_of(FITS_Solution)
        return self.FromArray(Solution=Solution, **k)

    def FromDevice(self, Solution_GPU, SFFTConfig):
        """Same as FromArray for a Solution resident on the GPU; returns a torch CUDA tensor (no D2H copy)."""
        return _device_realize(self.XY_q, Solution_GPU, SFFTConfig, True, False)[0]


class Realize_FluxScaling:
    def __init__(self, XY_q):
        self.XY_q = XY_q

    def FromArray(self, Solution, N0, N1, L0, L1, DK, Fpq):
        """Flux scaling sum_ij ac_ij00 x^i y^j at the requested coordinates (:160-181)."""
        s = _scaled(self.XY_q, N0, N1)
        blk = _blocks(Solution, N0, N1, L0, L1, DK, Fpq)
        w0, w1 = (L0 - 1) // 2, (L1 - 1) // 2
        out = np.zeros(s.shape[0])
        for k, (i, j) in enumerate(_ij_list(DK)):
            out += blk[k, w0, w1] * s[:, 0] ** i * s[:, 1] ** j
        return out

    def FromFITS(self, FITS_Solution):
        Solution, k = _header_of(FITS_Solution)
        return self.FromArray(Solution=Solution, **k)

    def FromDevice(self, Solution_GPU, SFFTConfig):
        return _device_realize(self.XY_q, Solution_GPU, SFFTConfig, False, True)[1]

"""
Customized_Packet.CP -- mirror of sfft/CustomizedPacket.py:14-223 on the B200 core.

FITS in -> FITS out, with the reference's conventions: images are the transposed FITS arrays, the union of
NaN pixels of REF / SCI is filled from the masked images before the subtraction and set back to NaN afterwards,
`ForceConv` selects which image is convolved, and the sign is flipped when the science image is the one
convolved so that transients on SCI stay positive.  astropy is not required: a minimal FITS reader / writer
(sfft_b200/fitsio.py) covers the primary-HDU image access the packet needs.
"""
import os.path as pa
import time
import numpy as np

from . import fitsio
from .sfftcore.SFFTConfigure import SingleSFFTConfigure
from .sfftcore.SFFTSubtract import GeneralSFFTSubtract

__all__ = ['Customized_Packet']


def _read_T(path):
    return np.ascontiguousarray(fitsio.getdata(path).T, np.float64)                    # CP :93-112


def _cp_on_device(FITS_REF, FITS_SCI, FITS_mREF, FITS_mSCI, ForceConv, GKerHW, KerPolyOrder, BGPolyOrder, ConstPhotRatio,
                  BACKEND_4SUBTRACT, CUDA_DEVICE_4SUBTRACT, VERBOSE_LEVEL, STORAGE, want_raw_diff):
    """CP with the I/O edge on the device: the raw FITS data blocks are uploaded as they are in the
    files and decoded / transposed by sfftb_fits_decode; the NaN union fill, the subtraction, the NaN restore, the sign
    flip and the encoding of the difference image all stay on the GPU (CustomizedPacket.py:93-203 without its host-side
    `.T` + float64 copies).  Returns (Solution, PixA_DIFF, SFFTConfig, raw difference block or None)."""
    import torch
    from . import _lib as B
    L = B.lib()
    dev_i = int(CUDA_DEVICE_4SUBTRACT)
    dev = torch.device('cuda', dev_i)
    imgs, shape = {}, None
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        for name, path in (('REF', FITS_REF), ('SCI', FITS_SCI), ('mREF', FITS_mREF), ('mSCI', FITS_mSCI)):
            cards, raw, bp, n1, n2, bscale, bzero = fitsio.read_raw(path)
            if shape is None:
                shape = (n1, n2)
            elif shape != (n1, n2):
                raise Exception('MeLOn ERROR: Input images should have same size!')
            raw_d = torch.from_numpy(raw).to(dev, non_blocking=True)
            out = torch.empty(shape, dtype=torch.float64, device=dev)
            B.check(L.sfftb_fits_decode(dev_i, stream.cuda_stream, raw_d.data_ptr(), bp, n1, n2, bscale, bzero, out.data_ptr(), B.F64))
            imgs[name] = out
            del raw_d
        n = shape[0] * shape[1]
        mask = torch.empty(n, dtype=torch.uint8, device=dev)
        flags = torch.zeros(2, dtype=torch.int32, device=dev)
        B.check(L.sfftb_nan_union_fill(dev_i, stream.cuda_stream, imgs['REF'].data_ptr(), imgs['SCI'].data_ptr(),
                                       imgs['mREF'].data_ptr(), imgs['mSCI'].data_ptr(), B.F64, n, mask.data_ptr(), flags.data_ptr()))
        fl = flags.cpu().numpy()
        assert fl[1] == 0, 'masked images must be NaN-free'                               # CP :118-119
        assert ForceConv in ['REF', 'SCI']
        SFFTConfig = SingleSFFTConfigure.SSC(NX=shape[0], NY=shape[1], KerHW=GKerHW, KerPolyOrder=KerPolyOrder,
                                             BGPolyOrder=BGPolyOrder, ConstPhotRatio=ConstPhotRatio,
                                             BACKEND_4SUBTRACT=BACKEND_4SUBTRACT, VERBOSE_LEVEL=VERBOSE_LEVEL,
                                             CUDA_DEVICE=dev_i, STORAGE=STORAGE)
        plan = SFFTConfig[1]['plan']
        if ForceConv == 'REF':
            I, J, mI, mJ = imgs['REF'], imgs['SCI'], imgs['mREF'], imgs['mSCI']
        else:
            I, J, mI, mJ = imgs['SCI'], imgs['REF'], imgs['mSCI'], imgs['mREF']
        sol_d = torch.empty(plan.NEQ, dtype=torch.float64, device=dev)
        diff_d = torch.empty(shape, dtype=torch.float64, device=dev)
        plan.bind_torch_stream(stream)
        plan.gss_device(I.data_ptr(), J.data_ptr(), mI.data_ptr(), mJ.data_ptr(), B.F64, sol_d.data_ptr(), diff_d.data_ptr(), B.F64)
        B.check(L.sfftb_nan_mask_apply(dev_i, stream.cuda_stream, diff_d.data_ptr(), B.F64, mask.data_ptr() if fl[0] else None, n,
                                       -1.0 if ForceConv == 'SCI' else 1.0))
        raw_out = None
        if want_raw_diff:
            raw_dd = torch.empty(n * 8, dtype=torch.uint8, device=dev)
            B.check(L.sfftb_fits_encode(dev_i, stream.cuda_stream, diff_d.data_ptr(), B.F64, shape[0], shape[1], -64, raw_dd.data_ptr()))
            raw_out = raw_dd.cpu().numpy()
        return sol_d.cpu().numpy(), diff_d.cpu().numpy(), SFFTConfig, raw_out


class Customized_Packet:
    @staticmethod
    def CP(FITS_REF, FITS_SCI, FITS_mREF, FITS_mSCI, ForceConv, GKerHW, FITS_DIFF=None, FITS_Solution=None,
           KerPolyOrder=2, BGPolyOrder=2, ConstPhotRatio=True, BACKEND_4SUBTRACT='B200',
           CUDA_DEVICE_4SUBTRACT='0', NUM_CPU_THREADS_4SUBTRACT=8, NUMBA_CACHE=True, VERBOSE_LEVEL=2,
           STORAGE='fp64', FITS_ON_DEVICE=False):
        """FITS_ON_DEVICE=True decodes the FITS data blocks and encodes the difference image on the GPU (no host-side
        transpose / float64 copies; needs torch for the device buffers); the default keeps the host route."""
        if FITS_ON_DEVICE:
            Solution, PixA_DIFF, SFFTConfig, raw_diff = _cp_on_device(
                FITS_REF, FITS_SCI, FITS_mREF, FITS_mSCI, ForceConv, GKerHW, KerPolyOrder, BGPolyOrder, ConstPhotRatio,
                BACKEND_4SUBTRACT, CUDA_DEVICE_4SUBTRACT, VERBOSE_LEVEL, STORAGE, FITS_DIFF is not None)
            if FITS_DIFF is not None:
                cards, _ = fitsio.read_header(FITS_SCI)
                fitsio.writeto(FITS_DIFF, (raw_diff, -64, PixA_DIFF.shape[0], PixA_DIFF.shape[1]), base_cards=cards, updates=[
                    ('NAME_REF', pa.basename(FITS_REF), 'MeLOn: SFFT'), ('NAME_SCI', pa.basename(FITS_SCI), 'MeLOn: SFFT'),
                    ('KERORDER', KerPolyOrder, 'MeLOn: SFFT'), ('BGORDER', BGPolyOrder, 'MeLOn: SFFT'),
                    ('CPHOTR', str(ConstPhotRatio), 'MeLOn: SFFT'), ('KERHW', GKerHW, 'MeLOn: SFFT'),
                    ('CONVD', ForceConv, 'MeLOn: SFFT')])
            if FITS_Solution is not None:
                P = SFFTConfig[0]
                ups = [(k, P[v], 'MeLOn: SFFT') for k, v in (('N0', 'N0'), ('N1', 'N1'), ('DK', 'DK'), ('DB', 'DB'),
                       ('L0', 'L0'), ('L1', 'L1'), ('FIJ', 'Fij'), ('FAB', 'Fab'), ('FPQ', 'Fpq'), ('FIJAB', 'Fijab'))]
                fitsio.writeto(FITS_Solution, Solution.reshape((-1, 1)).T, base_cards=None, updates=ups)
            return Solution, PixA_DIFF
        PixA_REF, PixA_SCI = _read_T(FITS_REF), _read_T(FITS_SCI)
        PixA_mREF, PixA_mSCI = _read_T(FITS_mREF), _read_T(FITS_mSCI)
        Solution, PixA_DIFF, SFFTConfig = Customized_Packet.CP_arrays(
            PixA_REF, PixA_SCI, PixA_mREF, PixA_mSCI, ForceConv, GKerHW, KerPolyOrder=KerPolyOrder,
            BGPolyOrder=BGPolyOrder, ConstPhotRatio=ConstPhotRatio, BACKEND_4SUBTRACT=BACKEND_4SUBTRACT,
            CUDA_DEVICE_4SUBTRACT=CUDA_DEVICE_4SUBTRACT, VERBOSE_LEVEL=VERBOSE_LEVEL, STORAGE=STORAGE,
            _return_config=True)
        if FITS_DIFF is not None:                                                      # CP :191-203
            cards, _ = fitsio.read_header(FITS_SCI)
            fitsio.writeto(FITS_DIFF, PixA_DIFF.T, base_cards=cards, updates=[
                ('NAME_REF', pa.basename(FITS_REF), 'MeLOn: SFFT'), ('NAME_SCI', pa.basename(FITS_SCI), 'MeLOn: SFFT'),
                ('KERORDER', KerPolyOrder, 'MeLOn: SFFT'), ('BGORDER', BGPolyOrder, 'MeLOn: SFFT'),
                ('CPHOTR', str(ConstPhotRatio), 'MeLOn: SFFT'), ('KERHW', GKerHW, 'MeLOn: SFFT'),
                ('CONVD', ForceConv, 'MeLOn: SFFT')])
        if FITS_Solution is not None:                                                  # CP :205-221
            P = SFFTConfig[0]
            ups = [(k, P[v], 'MeLOn: SFFT') for k, v in (('N0', 'N0'), ('N1', 'N1'), ('DK', 'DK'), ('DB', 'DB'),
                   ('L0', 'L0'), ('L1', 'L1'), ('FIJ', 'Fij'), ('FAB', 'Fab'), ('FPQ', 'Fpq'), ('FIJAB', 'Fijab'))]
            fitsio.writeto(FITS_Solution, Solution.reshape((-1, 1)).T, base_cards=None, updates=ups)
        return Solution, PixA_DIFF

    @staticmethod
    def CP_arrays(PixA_REF, PixA_SCI, PixA_mREF, PixA_mSCI, ForceConv, GKerHW, KerPolyOrder=2, BGPolyOrder=2,
                  ConstPhotRatio=True, BACKEND_4SUBTRACT='B200', CUDA_DEVICE_4SUBTRACT='0', VERBOSE_LEVEL=2,
                  STORAGE='fp64', _return_config=False):
        """The array-level body of CP (sfft/CustomizedPacket.py:114-188), usable without FITS files."""
        PixA_REF, PixA_SCI = np.asarray(PixA_REF, np.float64), np.asarray(PixA_SCI, np.float64)
        PixA_mREF, PixA_mSCI = np.asarray(PixA_mREF, np.float64), np.asarray(PixA_mSCI, np.float64)
        NaNmask_U = None
        NaNmask_REF, NaNmask_SCI = np.isnan(PixA_REF), np.isnan(PixA_SCI)
        if NaNmask_REF.any() or NaNmask_SCI.any():
            NaNmask_U = np.logical_or(NaNmask_REF, NaNmask_SCI)
        assert np.sum(np.isnan(PixA_mREF)) == 0
        assert np.sum(np.isnan(PixA_mSCI)) == 0
        assert ForceConv in ['REF', 'SCI']
        ConvdSide, KerHW = ForceConv, GKerHW

        if VERBOSE_LEVEL in [0, 1, 2]:
            print('MeLOn CheckPoint: TRIGGER Function Compilations of SFFT-SUBTRACTION!')
        t0 = time.time()
        SFFTConfig = SingleSFFTConfigure.SSC(NX=PixA_REF.shape[0], NY=PixA_REF.shape[1], KerHW=KerHW,
                                             KerPolyOrder=KerPolyOrder, BGPolyOrder=BGPolyOrder,
                                             ConstPhotRatio=ConstPhotRatio, BACKEND_4SUBTRACT=BACKEND_4SUBTRACT,
                                             VERBOSE_LEVEL=VERBOSE_LEVEL, CUDA_DEVICE=int(CUDA_DEVICE_4SUBTRACT),
                                             STORAGE=STORAGE)
        if VERBOSE_LEVEL in [1, 2]:
            print('\nMeLOn Report: Function Compilations of SFFT-SUBTRACTION TAKES [%.3f s]' % (time.time() - t0))

        if ConvdSide == 'REF':
            PixA_mI, PixA_mJ, PixA_I, PixA_J = PixA_mREF, PixA_mSCI, PixA_REF, PixA_SCI
        else:
            PixA_mI, PixA_mJ, PixA_I, PixA_J = PixA_mSCI, PixA_mREF, PixA_SCI, PixA_REF
        if NaNmask_U is not None:
            PixA_I, PixA_J = PixA_I.copy(), PixA_J.copy()
            PixA_I[NaNmask_U] = PixA_mI[NaNmask_U]
            PixA_J[NaNmask_U] = PixA_mJ[NaNmask_U]

        if VERBOSE_LEVEL in [0, 1, 2]:
            print('MeLOn CheckPoint: TRIGGER SFFT-SUBTRACTION!')
        t0 = time.time()
        Solution, PixA_DIFF = GeneralSFFTSubtract.GSS(PixA_I=PixA_I, PixA_J=PixA_J, PixA_mI=PixA_mI, PixA_mJ=PixA_mJ,
                                                      SFFTConfig=SFFTConfig, ContamMask_I=None,
                                                      BACKEND_4SUBTRACT=BACKEND_4SUBTRACT,
                                                      VERBOSE_LEVEL=VERBOSE_LEVEL)[:2]
        if VERBOSE_LEVEL in [1, 2]:
            print('\nMeLOn Report: SFFT-SUBTRACTION TAKES [%.3f s]' % (time.time() - t0))
        if NaNmask_U is not None:
            PixA_DIFF[NaNmask_U] = np.nan
        if ConvdSide == 'SCI':
            PixA_DIFF = -PixA_DIFF
        if _return_config:
            return Solution, PixA_DIFF, SFFTConfig
        return Solution, PixA_DIFF

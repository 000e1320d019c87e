"""Shared host glue of the decorrelation mirrors: packs the match kernels for sfftb_decorr (include/sfft_b200.h)."""
import ctypes as C
import numpy as np

from .. import _lib as B

_UMK = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=np.float64)


def _host(k):
    """Kernel stamp as a C-contiguous float64 host array (torch CUDA tensors are copied back: they are a few hundred taps)."""
    if k is None:
        return _UMK
    if hasattr(k, 'detach'):
        k = k.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(k, dtype=np.float64))


def decorr(N0, N1, KJ, SJ, KI, SI, KM, clip_ratio=0.0, real_size=None, normalize=True, device=0, stream=0, out_device_ptr=None):
    """Returns (array, lost_weight).  real_size=None: FKDECO on the (N0, N1) grid (host array, or written to `out_device_ptr`);
    real_size=(L0, L1): the truncated real-space kernel (host array)."""
    ks = [_host(k) for k in KJ] + [_host(k) for k in KI] + [_host(KM)]
    role = [0] * len(KJ) + [1] * len(KI) + [2]
    sig = [float(s) for s in SJ] + [float(s) for s in SI] + [0.0]
    for k in ks:
        if k.ndim != 2 or k.shape[0] % 2 == 0 or k.shape[1] % 2 == 0:
            raise Exception('MeLOn ERROR: only odd-sized 2-D kernels are supported, got %s' % (k.shape,))
    kdata = np.concatenate([k.ravel() for k in ks])
    kshape = np.array([s for k in ks for s in k.shape], dtype=np.int32)
    role = np.array(role, dtype=np.int32)
    sig = np.array(sig, dtype=np.float64)
    lost = C.c_double(float('nan'))
    L = B.lib()
    if real_size is None:
        if out_device_ptr is not None:
            B.check(L.sfftb_decorr(device, stream, N0, N1, len(ks), kdata.ctypes.data, kshape.ctypes.data, role.ctypes.data, sig.ctypes.data,
                                   float(clip_ratio), 0, 0, 0, int(bool(normalize)), out_device_ptr, B.MEM_DEVICE, C.addressof(lost)))
            return None, lost.value
        out = np.empty((N0, N1), np.float64)
        B.check(L.sfftb_decorr(device, stream, N0, N1, len(ks), kdata.ctypes.data, kshape.ctypes.data, role.ctypes.data, sig.ctypes.data,
                               float(clip_ratio), 0, 0, 0, int(bool(normalize)), out.ctypes.data, B.MEM_HOST, C.addressof(lost)))
        return out, lost.value
    L0, L1 = int(real_size[0]), int(real_size[1])
    out = np.empty((L0, L1), np.float64)
    B.check(L.sfftb_decorr(device, stream, N0, N1, len(ks), kdata.ctypes.data, kshape.ctypes.data, role.ctypes.data, sig.ctypes.data,
                           float(clip_ratio), 1, L0, L1, int(bool(normalize)), out.ctypes.data, B.MEM_HOST, C.addressof(lost)))
    return out, lost.value


def check_modes(MK_JLst, MK_ILst, MK_Fin, names=('Image-Stacking Mode requires at least 2 J-images!',
                                                 'Image-Stacking Mode requires at least 1 not-None J-kernel!',
                                                 'Image-Subtraction Mode requires at least 1 I-image & 1 J-image!',
                                                 'Image-Subtraction Mode requires at least 1 not-None J/I/Fin-kernel!')):
    """The mode checks of DCC / BDC / PCDC with the reference's messages; returns True for the subtraction mode."""
    NumI, NumJ = len(MK_ILst), len(MK_JLst)
    if NumI == 0:
        if NumJ < 2:
            raise Exception('MeLOn ERROR: %s' % names[0])
        if sum(m is not None for m in MK_JLst) == 0:
            raise Exception('MeLOn ERROR: %s' % names[1])
        return False
    if NumJ == 0:
        raise Exception('MeLOn ERROR: %s' % names[2])
    if sum(m is not None for m in list(MK_JLst) + list(MK_ILst) + [MK_Fin]) == 0:
        raise Exception('MeLOn ERROR: %s' % names[3])
    return True

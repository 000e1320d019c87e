"""
PureCupy_Customized_Packet.PCCP -- mirror of sfft/PureCupyCustomizedPacket.py:41-187: device arrays in,
device arrays out, no FITS, no host round trip.  "Cupy" in the name is the reference's; arrays here are torch
CUDA tensors or anything exposing __cuda_array_interface__.
"""
from . import _lib as B
from .sfftcore.SFFTConfigure import SingleSFFTConfigure
from .sfftcore.SFFTSubtract import GeneralSFFTSubtract_PureCupy

__all__ = ['PureCupy_Customized_Packet']


class PureCupy_Customized_Packet:
    @staticmethod
    def PCCP(PixA_REF_GPU, PixA_SCI_GPU, PixA_mREF_GPU, PixA_mSCI_GPU, ForceConv, GKerHW,
             KerPolyOrder=2, BGPolyOrder=2, ConstPhotRatio=True, CUDA_DEVICE_4SUBTRACT='0', VERBOSE_LEVEL=2,
             STORAGE='fp64', SFFTConfig=None):
        import torch
        arrs = (PixA_REF_GPU, PixA_SCI_GPU, PixA_mREF_GPU, PixA_mSCI_GPU)
        for a in arrs:                                                                  # :105-116
            assert len(a.shape) == 2
            cai = a.__cuda_array_interface__
            assert cai['typestr'] in ('<f8', '<f4')
            assert cai.get('strides') is None
        for m in (PixA_mREF_GPU, PixA_mSCI_GPU):
            assert not bool(torch.isnan(torch.as_tensor(m)).any())
        assert ForceConv in ['REF', 'SCI']
        ConvdSide, KerHW = ForceConv, GKerHW
        if SFFTConfig is None:
            SFFTConfig = SingleSFFTConfigure.SSC(NX=PixA_REF_GPU.shape[0], NY=PixA_REF_GPU.shape[1], KerHW=KerHW,
                                                 KerPolyOrder=KerPolyOrder, BGPolyOrder=BGPolyOrder,
                                                 ConstPhotRatio=ConstPhotRatio, BACKEND_4SUBTRACT='B200',
                                                 VERBOSE_LEVEL=VERBOSE_LEVEL, CUDA_DEVICE=int(CUDA_DEVICE_4SUBTRACT),
                                                 STORAGE=STORAGE)
        # union NaN mask of the unmasked pair (:124-131): those pixels are filled from the masked images before the
        # subtraction (:146-160) and set back to NaN in the difference image (:178-180).  One elementwise kernel builds
        # the mask and fills COPIES of REF / SCI (the caller's arrays are not modified); without a NaN the originals are used.
        plan = SFFTConfig[1]['plan']
        dev = torch.device('cuda', plan.device)
        tREF, tSCI = torch.as_tensor(PixA_REF_GPU, device=dev), torch.as_tensor(PixA_SCI_GPU, device=dev)
        tmREF, tmSCI = torch.as_tensor(PixA_mREF_GPU, device=dev), torch.as_tensor(PixA_mSCI_GPU, device=dev)
        NaNmask_GPU = None
        if bool(torch.isnan(tREF).any()) or bool(torch.isnan(tSCI).any()):
            L = B.lib()
            code = B.F64 if tREF.dtype == torch.float64 else B.F32
            tREF, tSCI = tREF.clone(), tSCI.clone()
            NaNmask_GPU = torch.empty(tREF.numel(), dtype=torch.uint8, device=dev)
            flags = torch.zeros(2, dtype=torch.int32, device=dev)
            sptr = torch.cuda.current_stream(dev).cuda_stream
            B.check(L.sfftb_nan_union_fill(plan.device, sptr, tREF.data_ptr(), tSCI.data_ptr(), tmREF.data_ptr(), tmSCI.data_ptr(),
                                           code, tREF.numel(), NaNmask_GPU.data_ptr(), flags.data_ptr()))
        if ConvdSide == 'REF':
            I, J, mI, mJ = tREF, tSCI, tmREF, tmSCI
        else:
            I, J, mI, mJ = tSCI, tREF, tmSCI, tmREF
        Solution_GPU, PixA_DIFF_GPU, _ = GeneralSFFTSubtract_PureCupy.GSS(
            PixA_I_GPU=I, PixA_J_GPU=J, PixA_mI_GPU=mI, PixA_mJ_GPU=mJ, SFFTConfig=SFFTConfig,
            ContamMask_I_GPU=None, VERBOSE_LEVEL=VERBOSE_LEVEL)
        if NaNmask_GPU is not None or ConvdSide == 'SCI':
            # NaN restore and the sign flip of a convolved science image (:178-185) in one pass
            code = B.F64 if PixA_DIFF_GPU.dtype == torch.float64 else B.F32
            B.check(B.lib().sfftb_nan_mask_apply(plan.device, torch.cuda.current_stream(dev).cuda_stream, PixA_DIFF_GPU.data_ptr(), code,
                                                 NaNmask_GPU.data_ptr() if NaNmask_GPU is not None else None,
                                                 PixA_DIFF_GPU.numel(), -1.0 if ConvdSide == 'SCI' else 1.0))
        return Solution_GPU, PixA_DIFF_GPU

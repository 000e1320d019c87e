"""
ElementalSFFTSubtract.ESS / GeneralSFFTSubtract.GSS / GeneralSFFTSubtract_PureCupy.GSS -- host-side mirrors of
sfft/sfftcore/SFFTSubtract.py:825-923 and :1373-1450 on top of the native plan.

Not reproduced on purpose: the reference's size check compares mI twice (:892)
-- here all four shapes are compared; its PureCupy contamination branch calls cp.zeros_like on a shape tuple
(:1442) -- here it works.
"""
import numpy as np

__all__ = ['ElementalSFFTSubtract', 'GeneralSFFTSubtract', 'GeneralSFFTSubtract_PureCupy']

_BACKENDS = ('B200', 'Cupy')


def _plan_of(SFFTConfig):
    mod = SFFTConfig[1]
    if not isinstance(mod, dict) or 'plan' not in mod:
        raise Exception('MeLOn ERROR: SFFTConfig was not produced by sfft_b200 SingleSFFTConfigure.SSC')
    return mod['plan']


def _check_backend(b):
    if b not in _BACKENDS:
        raise Exception("MeLOn ERROR: BACKEND_4SUBTRACT=%r is not available in sfft_b200 (use 'B200')" % (b,))


class ElementalSFFTSubtract:
    @staticmethod
    def ESS(PixA_I, PixA_J, SFFTConfig, SFFTSolution=None, Subtract=False,
            BACKEND_4SUBTRACT='B200', NUM_CPU_THREADS_4SUBTRACT=8, VERBOSE_LEVEL=2):
        """(Solution, PixA_DIFF) exactly as ElementalSFFTSubtract.ESS (SFFTSubtract.py:825-837):
        SFFTSolution=None -> fit on (I, J); Subtract=True -> DIFF = J - I (x) K - background."""
        _check_backend(BACKEND_4SUBTRACT)
        plan = _plan_of(SFFTConfig)
        N0, N1 = plan.shape
        if PixA_I.shape != (N0, N1) or PixA_J.shape != (N0, N1):                       # :37-39
            raise Exception('MeLOn ERROR: INCONSISTENT shape of input images I & J, [%d, %d] required!' % (N0, N1))
        if SFFTSolution is not None:
            Solution = np.asarray(SFFTSolution).astype(np.float64)                     # :189-193
        else:
            Solution = plan.fit(PixA_I, PixA_J)
        PixA_DIFF = None
        if Subtract:
            PixA_DIFF = plan.apply(PixA_I, PixA_J, Solution)
        return Solution, PixA_DIFF


class GeneralSFFTSubtract:
    @staticmethod
    def GSS(PixA_I, PixA_J, PixA_mI, PixA_mJ, SFFTConfig, ContamMask_I=None,
            BACKEND_4SUBTRACT='B200', NUM_CPU_THREADS_4SUBTRACT=8, VERBOSE_LEVEL=2):
        """Fit on the masked pair, subtract the unmasked pair, optionally propagate a contamination
        mask through the convolution (SFFTSubtract.py:841-923).  Host arrays in, host arrays out."""
        _check_backend(BACKEND_4SUBTRACT)
        plan = _plan_of(SFFTConfig)
        if len({tuple(PixA_I.shape), tuple(PixA_J.shape), tuple(PixA_mI.shape), tuple(PixA_mJ.shape)}) > 1:
            raise Exception('MeLOn ERROR: Input images should have same size!')        # :891-894
        Solution, PixA_DIFF = plan.gss(PixA_I, PixA_J, PixA_mI, PixA_mJ)               # :897-904, fused
        ContamMask_CI = None
        if ContamMask_I is not None:                                                   # :907-921
            tSolution = Solution.copy()
            Fpq = SFFTConfig[0]['Fpq']
            tSolution[-Fpq:] = 0.0
            _tmpI = np.asarray(ContamMask_I).astype(np.float64)
            _tmpJ = np.zeros(PixA_J.shape, np.float64)
            _tmpD = plan.apply(_tmpI, _tmpJ, tSolution)
            FTHRESH = -0.001
            ContamMask_CI = _tmpD < FTHRESH
        return Solution, PixA_DIFF, ContamMask_CI


class GeneralSFFTSubtract_PureCupy:
    @staticmethod
    def GSS(PixA_I_GPU, PixA_J_GPU, PixA_mI_GPU, PixA_mJ_GPU, SFFTConfig, ContamMask_I_GPU=None, VERBOSE_LEVEL=2):
        """Device arrays in, device arrays out (SFFTSubtract.py:1373-1450).  Inputs: anything exposing
        __cuda_array_interface__ (torch CUDA tensors, CuPy arrays), C-contiguous float64 (float32 also accepted);
        outputs are torch CUDA tensors (float64 Solution, DIFF in the input dtype)."""
        import torch
        from .. import _lib as B
        plan = _plan_of(SFFTConfig)
        imgs = (PixA_I_GPU, PixA_J_GPU, PixA_mI_GPU, PixA_mJ_GPU)
        if len({tuple(a.shape) for a in imgs}) > 1:
            raise Exception('MeLOn ERROR: Input images should have same size!')
        if tuple(PixA_I_GPU.shape) != plan.shape:
            raise Exception('MeLOn ERROR: INCONSISTENT shape of input images I & J, [%d, %d] required!' % plan.shape)
        cais = [a.__cuda_array_interface__ for a in imgs]
        for c in cais:
            assert c.get('strides') is None, 'inputs must be C-contiguous'             # :1003-1004
        ts = {c['typestr'] for c in cais}
        assert len(ts) == 1 and ts <= {'<f8', '<f4'}
        f64 = ts == {'<f8'}
        dev = torch.device('cuda', plan.device)
        tdt = torch.float64 if f64 else torch.float32
        code = B.F64 if f64 else B.F32
        Solution_GPU = torch.empty(plan.NEQ, dtype=torch.float64, device=dev)
        PixA_DIFF_GPU = torch.empty(plan.shape, dtype=tdt, device=dev)
        stream = torch.cuda.current_stream(dev)
        plan.bind_torch_stream(stream)
        plan.gss_device(cais[0]['data'][0], cais[1]['data'][0], cais[2]['data'][0], cais[3]['data'][0], code,
                        Solution_GPU.data_ptr(), PixA_DIFF_GPU.data_ptr(), code)
        ContamMask_CI_GPU = None
        if ContamMask_I_GPU is not None:
            tSolution = Solution_GPU.clone()
            tSolution[-SFFTConfig[0]['Fpq']:] = 0.0
            _tmpI = torch.as_tensor(ContamMask_I_GPU, device=dev).to(tdt).contiguous()
            _tmpJ = torch.zeros(plan.shape, dtype=tdt, device=dev)
            _tmpD = torch.empty(plan.shape, dtype=tdt, device=dev)
            plan.apply_device(_tmpI.data_ptr(), _tmpJ.data_ptr(), code, tSolution.data_ptr(), _tmpD.data_ptr(), code)
            ContamMask_CI_GPU = _tmpD < -0.001
        return Solution_GPU, PixA_DIFF_GPU, ContamMask_CI_GPU

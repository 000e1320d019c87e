"""
Golden fixtures of the noise-decorrelation path: the UNMODIFIED reference DeCorrelation_Calculator.DCC
(sfft/utils/DeCorrelationCalculator.py) and SkyLevel_Estimator.SLE (sfft/utils/SkyLevelEstimator.py), loaded by path from
/root/reference in the build container, run on the reference's own test inputs (test/difference_noise_decorrelation), next
to the known answer the reference ships (4check/DeCorrKernel.fits).

    python tests/golden/make_golden_decorr.py      ->  tests/golden/decorr_cases.npz
"""
import os
import sys
import types
import importlib.util
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFROOT = os.environ.get('SFFT_REFERENCE', '/root/reference')
from sfft_b200 import fitsio                       # noqa: E402


def load(name, path):
    s = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(s)
    s.loader.exec_module(m)
    return m


def main():
    ckc = load('sfft.utils.ConvKernelConvertion', os.path.join(REFROOT, 'sfft/utils/ConvKernelConvertion.py'))
    pkg, utils = types.ModuleType('sfft'), types.ModuleType('sfft.utils')
    sys.modules.update({'sfft': pkg, 'sfft.utils': utils, 'sfft.utils.ConvKernelConvertion': ckc})
    dcm = load('ref_dcc', os.path.join(REFROOT, 'sfft/utils/DeCorrelationCalculator.py'))
    sle = load('ref_sle', os.path.join(REFROOT, 'sfft/utils/SkyLevelEstimator.py'))
    T = os.path.join(REFROOT, 'test/difference_noise_decorrelation')
    out = {}
    sig, mks = {}, {}
    for grp in ('04', '18'):
        for tag in 'abcde':
            img = fitsio.getdata(os.path.join(T, 'input_data/DEC-OBS%s%s.mini.fits' % (grp, tag))).T.astype(float)
            sig[grp + tag] = float(sle.SkyLevel_Estimator.SLE(PixA_obj=img)[1])
            p = os.path.join(T, 'input_data/DEC-OBS%s%s.MatchKernel.fits' % (grp, tag))
            mks[grp + tag] = fitsio.getdata(p).T.astype(float) if tag != 'a' else None
    fin = fitsio.getdata(os.path.join(T, 'input_data/FinalMatchKernel.fits')).T.astype(float)
    S = [mks['04' + t] for t in 'abcde']
    R = [mks['18' + t] for t in 'abcde']
    sS = [sig['04' + t] for t in 'abcde']
    sR = [sig['18' + t] for t in 'abcde']
    k = dcm.DeCorrelation_Calculator.DCC(MK_JLst=S, SkySig_JLst=sS, MK_ILst=R, SkySig_ILst=sR, MK_Fin=fin, KERatio=2.0, VERBOSE_LEVEL=0)
    shipped = fitsio.getdata(os.path.join(T, '4check/DeCorrKernel.fits')).T.astype(float)
    print('reference DCC vs shipped 4check/DeCorrKernel.fits: max abs diff %.3e (kernel peak %.3e)' % (np.abs(k - shipped).max(), np.abs(shipped).max()))
    kstack = dcm.DeCorrelation_Calculator.DCC(MK_JLst=S, SkySig_JLst=sS, KERatio=2.0, VERBOSE_LEVEL=0)
    for t in 'bcde':
        out['mk04' + t] = mks['04' + t]
        out['mk18' + t] = mks['18' + t]
    out.update(mkfin=fin, sig04=np.array(sS), sig18=np.array(sR), dcc_sub=k, dcc_stack=kstack, shipped=shipped)
    np.savez_compressed(os.path.join(HERE, 'decorr_cases.npz'), **out)
    print('wrote decorr_cases.npz', {k_: (v.shape if hasattr(v, 'shape') else v) for k_, v in out.items()})


if __name__ == '__main__':
    main()

"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the host mirror of the
reference interface validates its arguments like the reference does, FITS round trip.  No compute calls."""
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from sfft_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'sfft_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(sfftb_[a-z0-9_]+)\s*\(', hdr)))
    assert len(declared) >= 15
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(declared) == sorted(_lib.EXPORTS)
    assert L.sfftb_version() == 100


def test_struct_layout_matches_header():
    import ctypes as C
    from sfft_b200 import _lib
    assert C.sizeof(_lib.Config) == 16 * 4
    assert C.sizeof(_lib.Dims) == 16 * 4


def test_ssc_argument_checks_match_reference():
    import sfft_b200 as sb
    with pytest.raises(Exception, match='KerPolyOrder should be 0/1/2/3'):
        sb.SingleSFFTConfigure.SSC(64, 64, 2, KerPolyOrder=4, VERBOSE_LEVEL=0)
    with pytest.raises(Exception, match='BGPolyOrder should be 0/1/2/3'):
        sb.SingleSFFTConfigure.SSC(64, 64, 2, BGPolyOrder=-1, VERBOSE_LEVEL=0)
    with pytest.raises(Exception, match='dramatically small size'):
        sb.SingleSFFTConfigure.SSC(4, 64, 1, VERBOSE_LEVEL=0)
    with pytest.raises(Exception, match='no CPU backend'):
        sb.SingleSFFTConfigure.SSC(64, 64, 2, BACKEND_4SUBTRACT='Numpy', VERBOSE_LEVEL=0)


def test_no_gpu_fails_loudly():
    """Without a CUDA device plan creation must raise (no silent CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import sfft_b200 as sb
    with pytest.raises(Exception, match='MeLOn ERROR'):
        sb.SingleSFFTConfigure.SSC(64, 64, 2, VERBOSE_LEVEL=0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'sfft_b200')
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, fn)).read()
                assert 'oracle' not in txt.replace('no CPU', ''), fn


def test_fits_roundtrip(tmp_path):
    from sfft_b200 import fitsio
    a = np.random.default_rng(0).normal(size=(7, 5))
    p = str(tmp_path / 'a.fits')
    fitsio.writeto(p, a, updates=[('KERHW', 4, 'MeLOn: SFFT'), ('CONVD', 'REF', 'MeLOn: SFFT')])
    b = fitsio.getdata(p)
    assert np.array_equal(a, b)
    h = fitsio.header_dict(fitsio.read_header(p)[0])
    assert h['KERHW'] == 4 and h['CONVD'] == 'REF'

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


def test_reference_fits_fixture_rebuilds_bit_exactly(tmp_path):
    """tests/golden/ztf_fits_headers.npz + ztf1024.npz reproduce the reference's known-answer FITS inputs (SHA-256 of the
    whole files), and the minimal FITS reader decodes them to the arrays of the fixture."""
    import hashlib
    import os
    import numpy as np
    from goldenio import load_case
    from sfft_b200 import fitsio
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ztf_fits_headers.npz'))
    case = load_case('ztf1024')
    for key in ('REF', 'SCI', 'mREF', 'mSCI'):
        data = np.ascontiguousarray(case[key].T).astype('>f8').tobytes()
        blob = bytes(z[key + '_header']) + data + b'\0' * ((-len(data)) % 2880)
        assert hashlib.sha256(blob).hexdigest() == str(z[key + '_sha256'])
        path = str(tmp_path / (key + '.fits'))
        open(path, 'wb').write(blob)
        assert np.array_equal(fitsio.getdata(path).T, case[key], equal_nan=True)
        cards, raw, bp, n1, n2, bs, bz = fitsio.read_raw(path)
        assert (bp, n1, n2, bs, bz) == (-64, 1024, 1024, 1.0, 0.0) and raw.size == 8 * 1024 * 1024
        open(path, 'wb').write(blob[:len(blob) // 2])
        with __import__('pytest').raises(IOError):
            fitsio.read_raw(path)

"""
Fixture for the FITS edge of the packets: the header blocks of the reference's own known-answer test files
(test/subtract_test_customized/input_data/*.fits, 4check/sfft_diff4check.fits), byte for byte, plus the SHA-256 of every
whole file.  Together with the pixel data already in ztf1024.npz the tests rebuild the reference-written FITS files
bit-exactly on the GPU box (where /root/reference does not exist) and check the hashes.

    python tests/golden/make_golden_fits.py   ->  tests/golden/ztf_fits_headers.npz
"""
import hashlib
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from goldenio import load_case                       # noqa: E402
from sfft_b200 import fitsio                         # noqa: E402

REFROOT = os.environ.get('SFFT_REFERENCE', '/root/reference')
D = os.path.join(REFROOT, 'test/subtract_test_customized')
FILES = {'REF': 'input_data/ztf_001735_zg_c01_q2_refimg.resampled.mini.fits',
         'mREF': 'input_data/ztf_001735_zg_c01_q2_refimg.resampled.mini.masked.fits',
         'SCI': 'input_data/ztf_20180705481609_001735_zg_c01_o_q2_sciimg.mini.fits',
         'mSCI': 'input_data/ztf_20180705481609_001735_zg_c01_o_q2_sciimg.mini.masked.fits',
         'GOLD': '4check/sfft_diff4check.fits'}

if __name__ == '__main__':
    case = load_case('ztf1024')
    out = {}
    for key, rel in FILES.items():
        path = os.path.join(D, rel)
        raw = open(path, 'rb').read()
        cards, off = fitsio.read_header(path)
        out[key + '_header'] = np.frombuffer(raw[:off], np.uint8).copy()
        out[key + '_name'] = np.array(os.path.basename(rel))
        out[key + '_sha256'] = np.array(hashlib.sha256(raw).hexdigest())
        out[key + '_size'] = np.array(len(raw))
        if key != 'GOLD':
            # the pixel data of the fixture reproduces the file's data block bit for bit (FITS stores the transpose)
            data = np.ascontiguousarray(case[key].T).astype('>f8').tobytes()
            pad = (-len(data)) % 2880
            assert raw[off:] == data + b'\0' * pad, key
    np.savez_compressed(os.path.join(HERE, 'ztf_fits_headers.npz'), **out)
    print('wrote ztf_fits_headers.npz', {k: int(v) for k, v in out.items() if k.endswith('_size')})

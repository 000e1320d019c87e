"""
Minimal FITS primary-HDU image reader / writer.

The reference's packets go through `astropy.io.fits` (sfft/CustomizedPacket.py:93-96,
191-221); astropy is not a dependency of this package, and the only FITS features the
subtraction path touches are: read a 2-D primary image (any BITPIX, optional
BSCALE/BZERO), copy a header, update a handful of cards, write a float image.
"""
import numpy as np

_BLOCK = 2880
_DTYPES = {8: '>u1', 16: '>i2', 32: '>i4', 64: '>i8', -32: '>f4', -64: '>f8'}


def _parse_value(raw):
    s = raw.strip()
    if s.startswith("'"):
        end = s.find("'", 1)
        while end != -1 and end + 1 < len(s) and s[end + 1] == "'":
            end = s.find("'", end + 2)
        return s[1:end].replace("''", "'").rstrip()
    s = s.split('/')[0].strip()
    if s == 'T':
        return True
    if s == 'F':
        return False
    try:
        return int(s)
    except ValueError:
        try:
            return float(s.replace('D', 'E'))
        except ValueError:
            return s


def read_header(path):
    """Return (cards, data_offset). cards is a list of raw 80-char strings up to, not including, END."""
    cards = []
    with open(path, 'rb') as f:
        off = 0
        while True:
            blk = f.read(_BLOCK)
            if len(blk) < _BLOCK:
                raise IOError('truncated FITS header: %s' % path)
            off += _BLOCK
            for i in range(0, _BLOCK, 80):
                c = blk[i:i + 80].decode('ascii', errors='replace')
                if c.startswith('END') and c[3:].strip() == '':
                    return cards, off
                cards.append(c)


def header_dict(cards):
    d = {}
    for c in cards:
        if c[8:10] == '= ':
            d[c[:8].strip()] = _parse_value(c[10:])
    return d


def getdata(path):
    """Primary-HDU image as a numpy array of shape (NAXIS2, NAXIS1), like fits.getdata(path, ext=0)."""
    cards, off = read_header(path)
    h = header_dict(cards)
    if h.get('NAXIS') != 2:
        raise ValueError('only 2-D primary images are supported: %s' % path)
    n1, n2, bp = int(h['NAXIS1']), int(h['NAXIS2']), int(h['BITPIX'])
    a = np.fromfile(path, dtype=_DTYPES[bp], count=n1 * n2, offset=off).reshape(n2, n1)
    bscale, bzero = h.get('BSCALE', 1), h.get('BZERO', 0)
    if bscale != 1 or bzero != 0:
        a = a.astype(np.float64) * bscale + bzero
        return a
    return a.astype(a.dtype.newbyteorder('='))


def read_raw(path):
    """The primary image as it is in the file: (cards, raw uint8 array of the data block, bitpix, naxis1, naxis2, bscale,
    bzero).  No byte swap, no conversion -- the device decodes it (sfftb_fits_decode)."""
    cards, off = read_header(path)
    h = header_dict(cards)
    if h.get('NAXIS') != 2:
        raise ValueError('only 2-D primary images are supported: %s' % path)
    n1, n2, bp = int(h['NAXIS1']), int(h['NAXIS2']), int(h['BITPIX'])
    nbytes = n1 * n2 * abs(bp) // 8
    raw = np.fromfile(path, dtype=np.uint8, count=nbytes, offset=off)
    if raw.size != nbytes:                 # np.fromfile returns a short array for a truncated file; the device decode would
        raise IOError('truncated FITS data block: %s (%d of %d bytes)' % (path, raw.size, nbytes))   # read out of bounds
    return cards, raw, bp, n1, n2, float(h.get('BSCALE', 1)), float(h.get('BZERO', 0))


def _card(key, value, comment=None):
    if isinstance(value, bool):
        v = '%20s' % ('T' if value else 'F')
    elif isinstance(value, (int, np.integer)):
        v = '%20d' % value
    elif isinstance(value, (float, np.floating)):
        v = '%20s' % repr(float(value)).upper()
    else:
        # a string value must keep its closing quote inside the 80 columns (no CONTINUE cards are written:
        # longer values are truncated to the 68 characters a single card holds)
        body = str(value).replace("'", "''")[:68]
        if body.endswith("'") and (len(body) - len(body.rstrip("'"))) % 2 == 1:
            body = body[:-1]               # do not cut an escaped quote pair in half
        v = "'%-8s'" % body
    s = '%-8s= %s' % (key[:8].upper(), v)
    if comment and len(s) + 3 < 80:
        s = (s + ' / ' + comment)[:80]
    return '%-80s' % s


def _pad(b, fill):
    r = (-len(b)) % _BLOCK
    return b + fill * r


def writeto(path, data, base_cards=None, updates=None):
    """Write `data` (2-D; FITS axis order, i.e. shape (NAXIS2, NAXIS1)) as a primary HDU.

    base_cards: raw cards of a header to carry over (structural keywords are regenerated).
    updates   : list of (key, value, comment) appended/replaced after the carried-over cards.

    Differences from the reference's astropy writer (sfft/CustomizedPacket.py:191-203): only the primary HDU is written
    (extensions of the source file are not carried over), the packets write the difference image as float64 whatever
    the science image's BITPIX was, and non-ASCII bytes of a carried-over header are replaced by '?'.
    """
    raw_block = None
    if isinstance(data, tuple):          # (raw big-endian data block, bitpix, naxis1, naxis2): already encoded on the device
        raw_block, bitpix, n1_, n2_ = data

        class _Shape:
            shape = (n2_, n1_)
        data = _Shape()
    else:
        data = np.asarray(data)
        if data.ndim != 2:
            raise ValueError('2-D image required')
        bitpix = {np.dtype('float64'): -64, np.dtype('float32'): -32,
                  np.dtype('int32'): 32, np.dtype('int16'): 16, np.dtype('uint8'): 8}[np.dtype(data.dtype.name)]
    struct = ('SIMPLE', 'BITPIX', 'NAXIS', 'NAXIS1', 'NAXIS2', 'EXTEND', 'BSCALE', 'BZERO')
    cards = [_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', bitpix, 'array data type'),
             _card('NAXIS', 2, 'number of array dimensions'), _card('NAXIS1', data.shape[1]),
             _card('NAXIS2', data.shape[0])]
    upd = list(updates or [])
    upd_keys = {k[:8].upper() for k, _, _ in upd}
    for c in (base_cards or []):
        k = c[:8].strip()
        if k in struct or k in upd_keys:
            continue
        cards.append('%-80s' % c[:80])
    for k, v, cm in upd:
        cards.append(_card(k, v, cm))
    cards.append('%-80s' % 'END')
    hb = _pad(''.join(cards).encode('ascii', errors='replace'), b' ')
    if raw_block is not None:
        db = _pad(np.asarray(raw_block, np.uint8).tobytes(), b'\0')
    else:
        db = _pad(np.ascontiguousarray(data, dtype=_DTYPES[bitpix]).tobytes(), b'\0')
    with open(path, 'wb') as f:
        f.write(hb)
        f.write(db)

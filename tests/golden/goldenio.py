"""Compact storage of golden cases: images that are exactly representable in float32 are
stored as float32; masked images are stored as sparse deltas against the unmasked ones."""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _pack_img(a):
    a = np.asarray(a, np.float64)
    a32 = a.astype(np.float32)
    ok = np.array_equal(a32.astype(np.float64), a, equal_nan=True)
    return a32 if ok else a


def pack_case(REF, SCI, mREF, mSCI, **extra):
    d = dict(REF=_pack_img(REF), SCI=_pack_img(SCI))
    for name, m, u in (('mREF', mREF, REF), ('mSCI', mSCI, SCI)):
        m = np.asarray(m, np.float64)
        u = np.asarray(u, np.float64)
        diff = ~((m == u) | (np.isnan(m) & np.isnan(u)))
        idx = np.flatnonzero(diff.ravel()).astype(np.int64)
        d[name + '_idx'] = idx
        d[name + '_val'] = _pack_img(m.ravel()[idx])
    d.update(extra)
    return d


def save_case(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrays)


def load_case(name):
    z = np.load(os.path.join(HERE, name + '.npz'), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    REF = d['REF'].astype(np.float64)
    SCI = d['SCI'].astype(np.float64)
    out = dict(REF=REF, SCI=SCI)
    for name_, u in (('mREF', REF), ('mSCI', SCI)):
        m = u.copy()
        m.ravel()[d[name_ + '_idx']] = d[name_ + '_val'].astype(np.float64)
        out[name_] = m
    for k, v in d.items():
        if k not in ('REF', 'SCI') and not k.startswith(('mREF_', 'mSCI_')):
            out[k] = v
    return out

"""Plan object: owns one native sfftb_plan (the role SFFTModule_dict plays in the reference,
sfft/sfftcore/SFFTConfigure.py:50-75, :1371-1395)."""
import ctypes as C
import numpy as np

from . import _lib as B


def _ptr_of(a):
    """(pointer, memkind, dtype_code, keepalive) of a host numpy array or a device array
    (anything with __cuda_array_interface__, e.g. a torch CUDA tensor or a CuPy array)."""
    if hasattr(a, '__cuda_array_interface__'):
        cai = a.__cuda_array_interface__
        if cai.get('strides') is not None:
            raise Exception('MeLOn ERROR: device arrays must be C-contiguous')
        ts = cai['typestr']
        if ts not in ('<f8', '<f4'):
            raise Exception('MeLOn ERROR: device arrays must be float64 or float32')
        return cai['data'][0], B.MEM_DEVICE, (B.F64 if ts == '<f8' else B.F32), a
    arr = np.asarray(a)
    if arr.dtype != np.float32:
        arr = arr.astype(np.float64, copy=False)
    arr = np.ascontiguousarray(arr)
    return arr.ctypes.data, B.MEM_HOST, (B.F64 if arr.dtype == np.float64 else B.F32), arr


class Plan:
    def __init__(self, N0, N1, w0, w1, DK, DB, ConstPhotRatio, device=0, storage='fp64', fold=0, sca_degree=0):
        self._h = C.c_void_p()
        self._L = B.lib()
        cfg = B.Config()
        cfg.device, cfg.N0, cfg.N1, cfg.w0, cfg.w1 = int(device), int(N0), int(N1), int(w0), int(w1)
        cfg.DK, cfg.DB, cfg.const_phot_ratio = int(DK), int(DB), int(bool(ConstPhotRatio))
        cfg.storage = {'fp64': B.STORE_F64, 'fp32': B.STORE_F32}[storage]
        cfg.fold = int(fold)
        cfg.sca_degree = int(sca_degree)
        B.check(self._L.sfftb_plan_create(C.byref(self._h), C.byref(cfg)))
        d = B.Dims()
        B.check(self._L.sfftb_plan_dims(self._h, C.byref(d)))
        self.dims = {k: getattr(d, k) for k, _ in B.Dims._fields_}
        self.device, self.storage = int(device), storage
        self.shape = (int(N0), int(N1))
        self.NEQ = self.dims['NEQ']

    @classmethod
    def general(cls, N0, N1, w0, w1, ker, bkg, scaling_mode, sca=None, device=0, storage='fp64', DK=0, DB=0):
        """General-basis plan (sfftb_plan_create_general): `ker`, `bkg`, `sca` are (U, V, fu, fv) tuples -- the 1-D tables
        U (nu, N0), V (nv, N1) and, per 2-D basis function, the rows it is the product of."""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        self._L = B.lib()
        cfg = B.Config()
        cfg.device, cfg.N0, cfg.N1, cfg.w0, cfg.w1 = int(device), int(N0), int(N1), int(w0), int(w1)
        cfg.DK, cfg.DB = int(DK), int(DB)
        cfg.storage = {'fp64': B.STORE_F64, 'fp32': B.STORE_F32}[storage]
        keep = []

        def mk(b):
            if b is None:
                return None
            U = np.ascontiguousarray(b[0], np.float64)
            V = np.ascontiguousarray(b[1], np.float64)
            fu = np.ascontiguousarray(b[2], np.int32)
            fv = np.ascontiguousarray(b[3], np.int32)
            if U.ndim != 2 or V.ndim != 2 or U.shape[1] != N0 or V.shape[1] != N1 or fu.shape != fv.shape or fu.ndim != 1:
                raise Exception('MeLOn ERROR: basis tables must have shapes (nu, N0), (nv, N1) and index vectors of one length')
            keep.extend([U, V, fu, fv])
            s = B.Basis()
            s.nu, s.nv, s.nf = U.shape[0], V.shape[0], fu.shape[0]
            s.U, s.V, s.fu, s.fv = U.ctypes.data, V.ctypes.data, fu.ctypes.data, fv.ctypes.data
            keep.append(s)
            return s
        bk, bs, bb = mk(ker), mk(sca), mk(bkg)
        B.check(self._L.sfftb_plan_create_general(C.byref(self._h), C.byref(cfg), C.byref(bk), C.byref(bs) if bs is not None else None,
                                                  C.byref(bb), int(scaling_mode)))
        d = B.Dims()
        B.check(self._L.sfftb_plan_dims(self._h, C.byref(d)))
        self.dims = {k: getattr(d, k) for k, _ in B.Dims._fields_}
        self.device, self.storage = int(device), storage
        self.shape = (int(N0), int(N1))
        self.NEQ = self.dims['NEQ']
        return self

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._L.sfftb_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing --------------------------------------------------------------------------
    def set_stream(self, stream_ptr):
        """Raw cudaStream_t; 0 / None = the plan's own (non-blocking) stream."""
        B.check(self._L.sfftb_plan_set_stream(self._h, C.c_void_p(stream_ptr or 0)))

    def bind_torch_stream(self, stream):
        """Run the plan's work on a torch stream so that it is ordered after whatever the caller has queued there.
        torch's default stream has the handle 0, which sfftb_plan_set_stream reads as "the plan's own stream" -- a
        non-blocking stream that does NOT synchronise with the legacy default stream; bind cudaStreamLegacy (0x1) then."""
        self.set_stream(stream.cuda_stream or 0x1)

    def set_partition(self, solver_sms):
        """SM partition for several pairs in flight on one GPU (sfftb_plan_set_partition): the Cholesky on `solver_sms` SMs,
        the persistent throughput kernels on the others; 0 = off."""
        B.check(self._L.sfftb_plan_set_partition(self._h, int(solver_sms)))

    def sync(self):
        B.check(self._L.sfftb_plan_sync(self._h))

    def set_timing(self, on=True):
        B.check(self._L.sfftb_plan_set_timing(self._h, int(bool(on))))

    def timings(self):
        ms = (C.c_float * 8)()
        B.check(self._L.sfftb_timings(self._h, ms, 8))
        # 'fit_cols' = column moments + fit column kernel(s); 'fit_cols_kernel' = the fit column kernel(s) alone (segmented path)
        keys = ('fit_rows', 'fit_cols', 'fit_reduce_fill', 'fit_solve', 'apply_rows', 'apply_cols', 'apply_inv_rows', 'fit_cols_kernel')
        return dict(zip(keys, [float(v) for v in ms]))

    @property
    def launch_count(self):
        return int(self._L.sfftb_launch_count(self._h))

    @property
    def last_solver(self):
        return {0: None, 1: 'cholesky', 2: 'lu', 3: 'cholesky-cached'}[int(self._L.sfftb_last_solver(self._h))]

    def _check_pair(self, *imgs):
        for a in imgs:
            shp = tuple(a.shape)
            if shp != self.shape:
                raise Exception('MeLOn ERROR: INCONSISTENT shape of input images I & J, [%d, %d] required!' % self.shape)

    # ---- host-array entry points (numpy in / numpy out) -----------------------------------------
    def fit(self, PixA_I, PixA_J):
        self._check_pair(PixA_I, PixA_J)
        pI, mk, dt, kI = _ptr_of(PixA_I)
        pJ, mk2, dt2, kJ = _ptr_of(PixA_J)
        if (mk, dt) != (mk2, dt2):
            raise Exception('MeLOn ERROR: I and J must live in the same memory and share a dtype')
        sol = np.empty(self.NEQ, np.float64)
        B.check(self._L.sfftb_fit(self._h, pI, pJ, mk, dt, sol.ctypes.data, B.MEM_HOST))
        return sol

    def apply(self, PixA_I, PixA_J, Solution, out_dtype=np.float64):
        self._check_pair(PixA_I, PixA_J)
        pI, mk, dt, kI = _ptr_of(PixA_I)
        pJ, mk2, dt2, kJ = _ptr_of(PixA_J)
        if (mk, dt) != (mk2, dt2):
            raise Exception('MeLOn ERROR: I and J must live in the same memory and share a dtype')
        sol = np.ascontiguousarray(np.asarray(Solution, np.float64))
        if sol.shape != (self.NEQ,):
            raise Exception('MeLOn ERROR: SFFTSolution must have shape (%d,)' % self.NEQ)
        diff = np.empty(self.shape, out_dtype)
        B.check(self._L.sfftb_apply(self._h, pI, pJ, mk, dt, sol.ctypes.data, B.MEM_HOST, diff.ctypes.data, B.MEM_HOST,
                                    B.F64 if diff.dtype == np.float64 else B.F32))
        return diff

    def gss(self, PixA_I, PixA_J, PixA_mI, PixA_mJ, out_dtype=np.float64):
        self._check_pair(PixA_I, PixA_J, PixA_mI, PixA_mJ)
        ptrs = [_ptr_of(a) for a in (PixA_I, PixA_J, PixA_mI, PixA_mJ)]
        if len({(p[1], p[2]) for p in ptrs}) != 1:
            raise Exception('MeLOn ERROR: all four images must live in the same memory and share a dtype')
        sol = np.empty(self.NEQ, np.float64)
        diff = np.empty(self.shape, out_dtype)
        B.check(self._L.sfftb_gss(self._h, ptrs[0][0], ptrs[1][0], ptrs[2][0], ptrs[3][0], ptrs[0][1], ptrs[0][2],
                                  sol.ctypes.data, B.MEM_HOST, diff.ctypes.data, B.MEM_HOST,
                                  B.F64 if diff.dtype == np.float64 else B.F32))
        return sol, diff

    def set_regularizer(self, SST=None, iREG=None, lam=0.0, CSST=None, DSST=None):
        """Kernel regularisation (BSplineSFFT.py:3570-3700): LHMAT += lam * SCALE^2 * kron(SST, iREG) on the kernel block."""
        if SST is None:
            B.check(self._L.sfftb_set_regularizer(self._h, None, None, 0.0))
            return
        S = np.ascontiguousarray(SST, np.float64)
        R = np.ascontiguousarray(iREG, np.float64)
        d = self.dims
        if S.shape != (d['Fij'], d['Fij']) or R.shape != (d['Fab'], d['Fab']):
            raise Exception('MeLOn ERROR: regulariser factors must have shapes (Fij, Fij) and (Fab, Fab)')
        B.check(self._L.sfftb_set_regularizer(self._h, S.ctypes.data, R.ctypes.data, float(lam)))
        if CSST is not None:           # SEPARATE-VARYING: scaling-basis Gram matrices for the centre taps
            Cm = np.ascontiguousarray(CSST, np.float64)
            Dm = np.ascontiguousarray(DSST, np.float64)
            if Cm.shape != S.shape or Dm.shape != S.shape:
                raise Exception('MeLOn ERROR: CSST / DSST must have shape (Fij, Fij)')
            B.check(self._L.sfftb_set_regularizer_varying(self._h, Cm.ctypes.data, Dm.ctypes.data))

    # ---- asynchronous host-buffer GSS (sfftb_gss_submit / sfftb_gss_finish) -------------------------
    def gss_submit(self, PixA_I, PixA_J, PixA_mI, PixA_mJ, out_dtype=np.float64, Solution_out=None, DIFF_out=None):
        """Queue one GSS on host arrays and return at once; gss_finish() waits for it and returns (Solution, DIFF).
        Arrays are used in place when they are C-contiguous float64 / float32 (pinned memory makes the copies
        asynchronous); optional preallocated (pinned) outputs may be passed."""
        self._check_pair(PixA_I, PixA_J, PixA_mI, PixA_mJ)
        ptrs = [_ptr_of(a) for a in (PixA_I, PixA_J, PixA_mI, PixA_mJ)]
        if len({(q[1], q[2]) for q in ptrs}) != 1 or ptrs[0][1] != B.MEM_HOST:
            raise Exception('MeLOn ERROR: gss_submit takes four host arrays of one dtype')
        sol = np.empty(self.NEQ, np.float64) if Solution_out is None else Solution_out
        diff = np.empty(self.shape, out_dtype) if DIFF_out is None else DIFF_out
        dptr = diff.ctypes.data if isinstance(diff, np.ndarray) else diff.data_ptr()
        sptr = sol.ctypes.data if isinstance(sol, np.ndarray) else sol.data_ptr()
        ddt = B.F64 if str(diff.dtype).endswith('float64') else B.F32
        B.check(self._L.sfftb_gss_submit(self._h, ptrs[0][0], ptrs[1][0], ptrs[2][0], ptrs[3][0], ptrs[0][2], sptr, dptr, ddt))
        self._inflight = (sol, diff, [q[3] for q in ptrs])          # keep the buffers alive until finish

    def gss_submit_delta(self, PixA_I, PixA_J, delta_I, delta_J, out_dtype=np.float64, Solution_out=None, DIFF_out=None):
        """gss_submit with the masked pair as sparse deltas: delta_I = (idx, val) with mI = I except mI.flat[idx] = val
        (sfftb_gss_submit_delta; see batch.sparse_delta).  Two images instead of four are copied to the device."""
        self._check_pair(PixA_I, PixA_J)
        ptrs = [_ptr_of(a) for a in (PixA_I, PixA_J)]
        if len({(q[1], q[2]) for q in ptrs}) != 1 or ptrs[0][1] != B.MEM_HOST:
            raise Exception('MeLOn ERROR: gss_submit_delta takes two host arrays of one dtype')
        vdt = np.float64 if ptrs[0][2] == B.F64 else np.float32
        ds = []
        for (idx, val) in (delta_I, delta_J):
            idx = np.ascontiguousarray(idx, np.int64)
            val = np.ascontiguousarray(val, vdt)
            if idx.shape != val.shape or idx.ndim != 1:
                raise Exception('MeLOn ERROR: a delta is a pair of 1-D arrays (flat pixel indices, values)')
            ds.append((idx, val))
        sol = np.empty(self.NEQ, np.float64) if Solution_out is None else Solution_out
        diff = np.empty(self.shape, out_dtype) if DIFF_out is None else DIFF_out
        dptr = diff.ctypes.data if isinstance(diff, np.ndarray) else diff.data_ptr()
        sptr = sol.ctypes.data if isinstance(sol, np.ndarray) else sol.data_ptr()
        ddt = B.F64 if str(diff.dtype).endswith('float64') else B.F32
        B.check(self._L.sfftb_gss_submit_delta(self._h, ptrs[0][0], ptrs[1][0], ds[0][0].size, ds[0][0].ctypes.data, ds[0][1].ctypes.data,
                                               ds[1][0].size, ds[1][0].ctypes.data, ds[1][1].ctypes.data, ptrs[0][2], sptr, dptr, ddt))
        self._inflight = (sol, diff, [q[3] for q in ptrs] + ds)

    def gss_submit_device(self, pI, pJ, pmI, pmJ, img_dtype, psol, pdiff, diff_dtype):
        """Device pointers in / out, queued on the plan's stream without a host synchronisation; gss_finish() waits."""
        B.check(self._L.sfftb_gss_submit_device(self._h, pI, pJ, pmI, pmJ, img_dtype, psol, pdiff, diff_dtype))
        self._inflight = (None, None, None)

    def gen_info(self):
        out = (C.c_int * 8)()
        B.check(self._L.sfftb_gen_info(self._h, out))
        keys = ('passes', 'staged_planes_total', 'lag_rows', 'stored_planes', 'column_planes', 'unknowns', 'apply_planes', 'fir_planes')
        return dict(zip(keys, [int(v) for v in out]))

    def gss_template_submit(self, PixA_J, PixA_mJ, out_dtype=np.float64, Solution_out=None, DIFF_out=None):
        """Queue one science tile against the cached template (host arrays); gss_finish() returns (Solution, DIFF)."""
        self._check_pair(PixA_J, PixA_mJ)
        ptrs = [_ptr_of(a) for a in (PixA_J, PixA_mJ)]
        if len({(q[1], q[2]) for q in ptrs}) != 1 or ptrs[0][1] != B.MEM_HOST:
            raise Exception('MeLOn ERROR: gss_template_submit takes two host arrays of one dtype')
        sol = np.empty(self.NEQ, np.float64) if Solution_out is None else Solution_out
        diff = np.empty(self.shape, out_dtype) if DIFF_out is None else DIFF_out
        dptr = diff.ctypes.data if isinstance(diff, np.ndarray) else diff.data_ptr()
        sptr = sol.ctypes.data if isinstance(sol, np.ndarray) else sol.data_ptr()
        ddt = B.F64 if str(diff.dtype).endswith('float64') else B.F32
        B.check(self._L.sfftb_gss_template_submit(self._h, ptrs[0][0], ptrs[1][0], B.MEM_HOST, ptrs[0][2], sptr, dptr, ddt))
        self._inflight = (sol, diff, [q[3] for q in ptrs])

    def gss_template_submit_device(self, pJ, pmJ, img_dtype, psol, pdiff, diff_dtype):
        """Device pointers in / out, queued on the plan's stream without a host synchronisation; gss_finish() waits."""
        B.check(self._L.sfftb_gss_template_submit(self._h, pJ, pmJ, B.MEM_DEVICE, img_dtype, psol, pdiff, diff_dtype))
        self._inflight = (None, None, None)

    def gss_finish(self):
        B.check(self._L.sfftb_gss_finish(self._h))
        sol, diff, _ = self._inflight
        self._inflight = None
        return sol, diff

    # ---- shared-template batch path (BASELINE config 4) ---------------------------
    def template_prepare(self, PixA_I, PixA_mI):
        """Row spectra of the convolved image (the template when ForceConv='REF') and of its masked version,
        computed once and reused by every gss_template call."""
        self._check_pair(PixA_I, PixA_mI)
        pI, mk, dt, kI = _ptr_of(PixA_I)
        pm, mk2, dt2, km = _ptr_of(PixA_mI)
        if (mk, dt) != (mk2, dt2):
            raise Exception('MeLOn ERROR: I and mI must live in the same memory and share a dtype')
        B.check(self._L.sfftb_template_prepare(self._h, pI, pm, mk, dt))

    def template_state(self):
        """(device pointer, bytes) of the template state buffer -- what gets broadcast to the other GPUs."""
        ptr, n = C.c_void_p(), C.c_size_t()
        B.check(self._L.sfftb_template_state(self._h, C.byref(ptr), C.byref(n)))
        return int(ptr.value), int(n.value)

    def template_state_tensor(self):
        """The template state as a torch uint8 CUDA tensor aliasing the plan's buffer (for torch.distributed)."""
        import torch
        ptr, n = self.template_state()

        class _Alias:
            __cuda_array_interface__ = {'shape': (n,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2, 'strides': None}
        with torch.cuda.device(self.device):
            return torch.as_tensor(_Alias(), device=torch.device('cuda', self.device))

    def template_mark_ready(self):
        B.check(self._L.sfftb_template_mark_ready(self._h))

    def template_clone(self, src):
        """Copy the shared-template state of `src` (spectra, and -- once src has them -- Cholesky factor, lag rows, cached segment
        spectra) into this plan: one factorisation per template for all plans of a pipeline."""
        B.check(self._L.sfftb_template_clone(self._h, src._h))

    def gss_template(self, PixA_J, PixA_mJ, out_dtype=np.float64):
        self._check_pair(PixA_J, PixA_mJ)
        pJ, mk, dt, kJ = _ptr_of(PixA_J)
        pm, mk2, dt2, km = _ptr_of(PixA_mJ)
        if (mk, dt) != (mk2, dt2):
            raise Exception('MeLOn ERROR: J and mJ must live in the same memory and share a dtype')
        sol = np.empty(self.NEQ, np.float64)
        diff = np.empty(self.shape, out_dtype)
        B.check(self._L.sfftb_gss_template(self._h, pJ, pm, mk, dt, sol.ctypes.data, B.MEM_HOST, diff.ctypes.data,
                                           B.MEM_HOST, B.F64 if diff.dtype == np.float64 else B.F32))
        return sol, diff

    def gss_template_device(self, pJ, pmJ, img_dtype, psol, pdiff, diff_dtype):
        B.check(self._L.sfftb_gss_template(self._h, pJ, pmJ, B.MEM_DEVICE, img_dtype, psol, B.MEM_DEVICE,
                                           pdiff, B.MEM_DEVICE, diff_dtype))

    # ---- device-array entry points (pointers in / pointers out; used by the PureCupy-style API) ----
    def gss_device(self, pI, pJ, pmI, pmJ, img_dtype, psol, pdiff, diff_dtype):
        B.check(self._L.sfftb_gss(self._h, pI, pJ, pmI, pmJ, B.MEM_DEVICE, img_dtype, psol, B.MEM_DEVICE,
                                  pdiff, B.MEM_DEVICE, diff_dtype))

    def fit_device(self, pI, pJ, img_dtype, psol):
        B.check(self._L.sfftb_fit(self._h, pI, pJ, B.MEM_DEVICE, img_dtype, psol, B.MEM_DEVICE))

    def apply_device(self, pI, pJ, img_dtype, psol, pdiff, diff_dtype):
        B.check(self._L.sfftb_apply(self._h, pI, pJ, B.MEM_DEVICE, img_dtype, psol, B.MEM_DEVICE,
                                    pdiff, B.MEM_DEVICE, diff_dtype))

    # ---- parity hooks -----------------------------------------------------------------------------
    def export_normal_eq(self):
        n = self.NEQ
        L = np.empty((n, n), np.float64)
        b = np.empty(n, np.float64)
        B.check(self._L.sfftb_export_normal_eq(self._h, L.ctypes.data, b.ctypes.data))
        return L, b

    def export_solved_system(self):
        """(L, b) of the system the last fit solved, after the stripe tweak (sfftb_export_solved_system)."""
        n = self.dims['NEQ_FSfree']
        L = np.empty((n, n), np.float64)
        b = np.empty(n, np.float64)
        B.check(self._L.sfftb_export_solved_system(self._h, L.ctypes.data, b.ctypes.data))
        return L, b

    def dbg_row_spectra(self, which):
        N0, N1 = self.shape
        nj = self.dims['DK'] + 1 if which == 0 else 1
        out = np.empty((nj, N1 // 2 + 1, N0), np.complex128)
        B.check(self._L.sfftb_dbg_row_spectra(self._h, which, out.ctypes.data))
        return out

    def dbg_lag_tables(self):
        d = self.dims
        npairs = d['Fij'] * (d['Fij'] + 1) // 2
        R = np.empty((npairs, 4 * d['w0'] + 1, 4 * d['w1'] + 1))
        RJ = np.empty((d['Fij'], 2 * d['w0'] + 1, 2 * d['w1'] + 1))
        RT = np.empty((d['Fij'], d['Fpq'], 2 * d['w0'] + 1, 2 * d['w1'] + 1))
        RJT = np.empty(d['Fpq'])
        B.check(self._L.sfftb_dbg_lag_tables(self._h, R.ctypes.data, RJ.ctypes.data, RT.ctypes.data, RJT.ctypes.data))
        return R, RJ, RT, RJT


def dbg_fft1d(x, sign=-1, device=0):
    """Batched 1-D complex FFT through the shared-memory engine (unit-test hook)."""
    x = np.ascontiguousarray(np.asarray(x, np.complex128))
    x2 = x.reshape(-1, x.shape[-1])
    out = np.empty_like(x2)
    B.check(B.lib().sfftb_dbg_fft1d(int(device), x2.shape[1], x2.shape[0], int(sign), x2.ctypes.data, out.ctypes.data))
    return out.reshape(x.shape)

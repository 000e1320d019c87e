// tu_rows.cu -- row passes: forward (real rows -> transposed half spectra) and inverse (FDIFF rows -> difference image).
#define SFFTB_TU_ROWS
#include "plan.h"

// Bp[c * 256 + d] = B[c + R d]: the chirp filter in the plane layout of the R x 256 decomposition
__global__ void blu_permute_kernel(int R, const cd* __restrict__ B, cd* __restrict__ Bp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 256 * R) return;
    const int c = i >> 8, d = i & 255;
    Bp[i] = B[c + R * d];
}

int rows_setup(sfftb_plan* p) {
    const sfftb_dims& d = p->d;
    const RowArgs& r = p->row;
    const bool f32 = p->cfg.storage == SFFTB_STORE_F32;
    if (init_generic_radix_tables()) return SFFTB_ECUDA;
    // forward: fp64 spectra always (apply step), fp32 spectra in addition for fp32 storage (fit step); inverse: fp64 only
    if (set_smem(row_fwd_kernel<float, double2>, p->smem_row) || set_smem(row_fwd_kernel<double, double2>, p->smem_row)) return SFFTB_ECUDA;
    if (f32 && (set_smem(row_fwd_kernel<float, float2>, p->smem_row) || set_smem(row_fwd_kernel<double, float2>, p->smem_row))) return SFFTB_ECUDA;
    if (set_smem(row_inv_kernel<double2, float>, p->smem_row) || set_smem(row_inv_kernel<double2, double>, p->smem_row)) return SFFTB_ECUDA;
    // ---- fast paths on the register FFT engines ----
    if (upload_engine_table(16, 16, &p->tabA)) return SFFTB_ECUDA;
    p->row_fast = 0;
    if (r.packed && !env_int("SFFTB_ROW_GENERIC", 0) &&
        (r.H == 512 || r.H == 1024 || r.H == 2048 || r.H == 4096 || r.H == 8192)) {
        const int R3 = reg_fft_tail_radix(r.H);
        if (upload_engine_table(256, R3, &p->tabB_row)) return SFFTB_ECUDA;
        if (r.H == 8192 && upload_engine_table(4096, 2, &p->tabC_row)) return SFFTB_ECUDA;
        RowFastArgs& rf = p->rowf;
        rf.N0 = d.N0; rf.N1 = d.N1; rf.NH = d.N1 / 2 + 1; rf.H = r.H;
        rf.tabA = p->tabA; rf.tabB = p->tabB_row; rf.tabC = p->tabC_row; rf.tw1 = p->tw1;
        p->rinvf.r = rf; p->rinvf.scale = p->rinv.scale; p->rinvf.Fpq = d.Fpq;
        memcpy(p->rinvf.p_of, p->rinv.p_of, 16); memcpy(p->rinvf.q_of, p->rinv.q_of, 16);
        p->row_fast = r.H;
    }
    p->row_v8 = 0;
    if (upload_engine_table(8, 8, &p->vt8_8) || upload_engine_table(64, 8, &p->vt64_8) || upload_engine_table(64, 4, &p->vt64_4) ||
        upload_engine_table(256, 4, &p->vt256_4) || upload_engine_table(512, 4, &p->vt512_4)) return SFFTB_ECUDA;
    p->vtabs.t8_8 = p->vt8_8; p->vtabs.t64_8 = p->vt64_8; p->vtabs.t64_4 = p->vt64_4; p->vtabs.t256_4 = p->vt256_4; p->vtabs.t512_4 = p->vt512_4;
    if (r.packed && !env_int("SFFTB_ROW_NOV8", 0) && (r.H == 256 || r.H == 512 || r.H == 1024 || r.H == 2048)) {
        RowV8Args& rv = p->rowv;
        rv.N0 = d.N0; rv.N1 = d.N1; rv.NH = d.N1 / 2 + 1; rv.H = r.H;
        rv.nit = std::max(1, env_int("SFFTB_ROW_NIT", 2));
        rv.tabs = p->vtabs;
        rv.tw1 = p->tw1;
        const int RBI = ROWV_NT / (r.H / 8);
        p->smem_rowv = sizeof(cd) * ((size_t)RBI * (r.H + r.H / 8 + 8) + 3000 + r.H / 2 + 1);
#define SET_ROWV(HH)                                                                                              \
        if (r.H == HH) {                                                                                              \
            if (f32 && (set_smem(row_fwd_v8_kernel<float, float2, HH>, p->smem_rowv) || set_smem(row_fwd_v8_kernel<double, float2, HH>, p->smem_rowv))) return SFFTB_ECUDA; \
            if (set_smem(row_fwd_v8_kernel<float, double2, HH>, p->smem_rowv) || set_smem(row_fwd_v8_kernel<double, double2, HH>, p->smem_rowv)) return SFFTB_ECUDA; \
        }
        SET_ROWV(256) SET_ROWV(512) SET_ROWV(1024) SET_ROWV(2048)
#undef SET_ROWV
        p->row_v8 = r.H;
    }
    p->row_h16 = 0;
    if (p->row_fast && !env_int("SFFTB_ROW_NOH16", 0) && r.H == 8192 && rowh32_smem_bytes() <= p->max_smem) {
        RowH16Args& rh = p->rowh;
        rh.N0 = d.N0; rh.N1 = d.N1; rh.NH = d.N1 / 2 + 1; rh.H = r.H;
        if (upload_engine_table(256, 32, &p->tabC_row32)) return SFFTB_ECUDA;
        rh.tabA = p->tabA; rh.twP = p->tabC_row32; rh.tw1 = p->tw1; rh.vtab = nullptr;
        const size_t smh = rowh32_smem_bytes();
        if (f32 && (set_smem(row_fwd_h16x32_kernel<float, float2>, smh) || set_smem(row_fwd_h16x32_kernel<double, float2>, smh))) return SFFTB_ECUDA;
        if (set_smem(row_fwd_h16x32_kernel<float, double2>, smh) || set_smem(row_fwd_h16x32_kernel<double, double2>, smh)) return SFFTB_ECUDA;
        p->row_h16 = r.H;
    }
    if (p->row_fast && !env_int("SFFTB_ROW_NOH16", 0) && (r.H == 1024 || r.H == 2048 || r.H == 4096) && rowh_smem_bytes(r.H) <= p->max_smem) {
        RowH16Args& rh = p->rowh;
        rh.N0 = d.N0; rh.N1 = d.N1; rh.NH = d.N1 / 2 + 1; rh.H = r.H;
        rh.tabA = p->tabA; rh.twP = p->tabB_row; rh.tw1 = p->tw1; rh.vtab = nullptr;
        const size_t smh = rowh_smem_bytes(r.H), smh64 = rowh_smem_bytes(r.H, rowh_threads(r.H, true));
#define SET_ROWH(RR)                                                                                              \
        if (r.H == 256 * RR) {                                                                                        \
            if (f32 && (set_smem(row_fwd_h16_kernel<float, float2, RR>, smh) || set_smem(row_fwd_h16_kernel<double, float2, RR>, smh))) return SFFTB_ECUDA; \
            if (set_smem(row_fwd_h16_kernel<float, double2, RR>, smh64) || set_smem(row_fwd_h16_kernel<double, double2, RR>, smh64)) return SFFTB_ECUDA; \
        }
        SET_ROWH(4) SET_ROWH(8) SET_ROWH(16)
#undef SET_ROWH
        p->row_h16 = r.H;
    }
    p->row_g16 = 0;
    if (r.packed && !p->row_fast && !env_int("SFFTB_ROW_GENERIC", 0) && !env_int("SFFTB_ROW_NOG16", 0) && r.H % 256 == 0 &&
        rowg_supported(r.H / 256) && rowg_smem_bytes(r.H / 256, p->vtab != nullptr) <= p->max_smem) {
        const int R = r.H / 256;
        if (upload_engine_table(256, R, &p->tabB_row)) return SFFTB_ECUDA;
        RowH16Args& rh = p->rowh;
        rh.N0 = d.N0; rh.N1 = d.N1; rh.NH = d.N1 / 2 + 1; rh.H = r.H;
        rh.tabA = p->tabA; rh.twP = p->tabB_row; rh.tw1 = p->tw1; rh.vtab = nullptr;
        RowFastArgs& rf = p->rowf;
        rf.N0 = d.N0; rf.N1 = d.N1; rf.NH = d.N1 / 2 + 1; rf.H = r.H;
        rf.tabA = p->tabA; rf.tabB = p->tabB_row; rf.tabC = nullptr; rf.tw1 = p->tw1;
        p->rinvf.r = rf; p->rinvf.scale = p->rinv.scale; p->rinvf.Fpq = d.Fpq; p->rinvf.row0 = 0;
        memcpy(p->rinvf.p_of, p->rinv.p_of, 16); memcpy(p->rinvf.q_of, p->rinv.q_of, 16);
        const size_t smg = rowg_smem_bytes(R, p->vtab != nullptr);
#define SET_ROWG(RR)                                                                                              \
        if (R == RR) {                                                                                                \
            if (f32 && (set_smem(row_fwd_g16_kernel<float, float2, RR>, smg) || set_smem(row_fwd_g16_kernel<double, float2, RR>, smg))) return SFFTB_ECUDA; \
            if (set_smem(row_fwd_g16_kernel<float, double2, RR>, smg) || set_smem(row_fwd_g16_kernel<double, double2, RR>, smg)) return SFFTB_ECUDA; \
            if (set_smem(row_inv_g16_kernel<double2, float, RR>, smg) || set_smem(row_inv_g16_kernel<double2, double, RR>, smg)) return SFFTB_ECUDA; \
        }
        SET_ROWG(3) SET_ROWG(5) SET_ROWG(6) SET_ROWG(10) SET_ROWG(12)
#undef SET_ROWG
        p->row_g16 = R;
    }
    p->row_blu = 0;
    if (r.packed && !p->row_fast && !p->row_g16 && !p->row_v8 && r.H > 128 && blu_radix(r.H) && !env_int("SFFTB_ROW_GENERIC", 0) &&
        !env_int("SFFTB_ROW_NOBLU16", 0) && blu_smem_bytes(blu_radix(r.H)) <= p->max_smem) {
        const int R = blu_radix(r.H), M = 256 * R;
        if (upload_engine_table(256, R, &p->bluTwP16) || upload_bluestein(r.H, M, &p->bluC16, &p->bluB16)) return SFFTB_ECUDA;
        CK(cudaMalloc(&p->bluBp16, sizeof(cd) * M));
        blu_permute_kernel<<<(M + 255) / 256, 256, 0, p->stream>>>(R, p->bluB16, p->bluBp16);
        CKL(p);
        RowBluArgs& rb = p->rowb;
        rb.N0 = d.N0; rb.N1 = d.N1; rb.NH = d.N1 / 2 + 1; rb.H = r.H;
        rb.tabA = p->tabA; rb.twP = p->bluTwP16; rb.tw1 = p->tw1; rb.chirp = p->bluC16; rb.Bp = p->bluBp16; rb.vtab = nullptr;
        RowFastArgs& rf = p->rowf;
        rf.N0 = d.N0; rf.N1 = d.N1; rf.NH = d.N1 / 2 + 1; rf.H = r.H;
        rf.tabA = p->tabA; rf.tabB = nullptr; rf.tabC = nullptr; rf.tw1 = p->tw1;
        p->rinvf.r = rf; p->rinvf.scale = p->rinv.scale; p->rinvf.Fpq = d.Fpq; p->rinvf.row0 = 0;
        memcpy(p->rinvf.p_of, p->rinv.p_of, 16); memcpy(p->rinvf.q_of, p->rinv.q_of, 16);
        const size_t smb = blu_smem_bytes(R);
#define SET_BLU(RR)                                                                                               \
        if (R == RR) {                                                                                                \
            if (f32 && (set_smem(row_fwd_blu_kernel<float, float2, RR>, smb) || set_smem(row_fwd_blu_kernel<double, float2, RR>, smb))) return SFFTB_ECUDA; \
            if (set_smem(row_fwd_blu_kernel<float, double2, RR>, smb) || set_smem(row_fwd_blu_kernel<double, double2, RR>, smb)) return SFFTB_ECUDA; \
            if (set_smem(row_inv_blu_kernel<double2, float, RR>, smb) || set_smem(row_inv_blu_kernel<double2, double, RR>, smb)) return SFFTB_ECUDA; \
        }
        SET_BLU(4) SET_BLU(8) SET_BLU(16)
#undef SET_BLU
        if (R == 32) {
            if (f32 && (set_smem(row_fwd_blu32_kernel<float, float2>, smb) || set_smem(row_fwd_blu32_kernel<double, float2>, smb))) return SFFTB_ECUDA;
            if (set_smem(row_fwd_blu32_kernel<float, double2>, smb) || set_smem(row_fwd_blu32_kernel<double, double2>, smb)) return SFFTB_ECUDA;
            if (set_smem(row_inv_blu32_kernel<double2, float>, smb) || set_smem(row_inv_blu32_kernel<double2, double>, smb)) return SFFTB_ECUDA;
        }
        p->row_blu = R;
    }
    if (p->row_fast) {
        const size_t sm = sizeof(cd) * (size_t)(ROWF_NT / (r.H / 16)) * (r.H + r.H / 16);
#define SET_ROWF(HH)                                                                                              \
        if (r.H == HH) {                                                                                              \
            if (set_smem(row_inv_fast_kernel<double2, float, HH>, sm) || set_smem(row_inv_fast_kernel<double2, double, HH>, sm)) return SFFTB_ECUDA; \
        }
        SET_ROWF(512) SET_ROWF(1024) SET_ROWF(2048) SET_ROWF(4096) SET_ROWF(8192)
#undef SET_ROWF
    }
    return 0;
}

template <typename TSt>
int launch_row_fwd(sfftb_plan* p, const void* img, int dtype, TSt* out, int nj, const double* vtab) {
    // vtab != NULL (general-basis plans, the planes of I only): plane j = row x vtab[j][c] instead of cy^j
    RowV8Args rowv = p->rowv; rowv.vtab = vtab;
    RowFastArgs rowf = p->rowf; rowf.vtab = vtab;
    RowArgs rowg = p->row; rowg.vtab = vtab;
    const size_t esz2 = dtype == SFFTB_F64 ? 16 : 8;
    if (p->row_blu && ((uintptr_t)img % esz2) == 0) {
        RowBluArgs rowb = p->rowb; rowb.vtab = vtab;
        const int R = p->row_blu, RBI = R == 32 ? 1 : BLU_NT / (16 * R);
        const int ngroups = (p->d.N0 + RBI - 1) / RBI;
        const int grid = std::min(ngroups, p->row_grid_limit > 0 ? p->row_grid_limit : work_sms(p));
        const size_t smb = blu_smem_bytes(R);
        if (R == 32) {
            if (dtype == SFFTB_F64) row_fwd_blu32_kernel<double, TSt><<<grid, BLU32_NT, smb, p->stream>>>(rowb, (const double*)img, out, nj);
            else row_fwd_blu32_kernel<float, TSt><<<grid, BLU32_NT, smb, p->stream>>>(rowb, (const float*)img, out, nj);
        }
#define RUN_BLU(RR)                                                                                                    \
        if (R == RR) {                                                                                                 \
            if (dtype == SFFTB_F64) row_fwd_blu_kernel<double, TSt, RR><<<grid, BLU_NT, smb, p->stream>>>(rowb, (const double*)img, out, nj); \
            else row_fwd_blu_kernel<float, TSt, RR><<<grid, BLU_NT, smb, p->stream>>>(rowb, (const float*)img, out, nj);                     \
        }
        RUN_BLU(4) RUN_BLU(8) RUN_BLU(16)
#undef RUN_BLU
        CKL(p);
        return 0;
    }
    if (p->row_g16 && ((uintptr_t)img % esz2) == 0) {
        RowH16Args rowh = p->rowh; rowh.vtab = vtab;
        const int R = p->row_g16, RBI = rowg_rbi(R);
        const int ngroups = (p->d.N0 + RBI - 1) / RBI;
        const int grid = std::min(ngroups, p->row_grid_limit > 0 ? p->row_grid_limit : work_sms(p));
        const size_t smg = rowg_smem_bytes(R, p->vtab != nullptr);
#define RUN_ROWG(RR)                                                                                                   \
        if (R == RR) {                                                                                                 \
            if (dtype == SFFTB_F64) row_fwd_g16_kernel<double, TSt, RR><<<grid, RowgCfg<RR>::nt, smg, p->stream>>>(rowh, (const double*)img, out, nj); \
            else row_fwd_g16_kernel<float, TSt, RR><<<grid, RowgCfg<RR>::nt, smg, p->stream>>>(rowh, (const float*)img, out, nj);                     \
        }
        RUN_ROWG(3) RUN_ROWG(5) RUN_ROWG(6) RUN_ROWG(10) RUN_ROWG(12)
#undef RUN_ROWG
        CKL(p);
        return 0;
    }
    if (p->row_h16 == 8192 && ((uintptr_t)img % esz2) == 0) {
        RowH16Args rowh = p->rowh; rowh.vtab = vtab;
        const int grid = std::min(p->d.N0, (p->row_grid_limit > 0 ? p->row_grid_limit : work_sms(p)));
        const size_t smh = rowh32_smem_bytes();
        if (dtype == SFFTB_F64) row_fwd_h16x32_kernel<double, TSt><<<grid, ROWH_NT, smh, p->stream>>>(rowh, (const double*)img, out, nj);
        else row_fwd_h16x32_kernel<float, TSt><<<grid, ROWH_NT, smh, p->stream>>>(rowh, (const float*)img, out, nj);
        CKL(p);
        return 0;
    }
    if (p->row_h16 && ((uintptr_t)img % esz2) == 0) {
        RowH16Args rowh = p->rowh; rowh.vtab = vtab;
        const int H = p->row_h16, NTH = rowh_threads(H, sizeof(TSt) == 16), RBI = NTH / (H / 16);
        const int ngroups = (p->d.N0 + RBI - 1) / RBI;
        const int grid = std::min(ngroups, (p->row_grid_limit > 0 ? p->row_grid_limit : work_sms(p)) * (ROWH_NT / NTH));
        const size_t smh = rowh_smem_bytes(H, NTH);
#define RUN_ROWH(RR)                                                                                                   \
        if (H == 256 * RR) {                                                                                           \
            if (dtype == SFFTB_F64) row_fwd_h16_kernel<double, TSt, RR><<<grid, NTH, smh, p->stream>>>(rowh, (const double*)img, out, nj); \
            else row_fwd_h16_kernel<float, TSt, RR><<<grid, NTH, smh, p->stream>>>(rowh, (const float*)img, out, nj);                     \
        }
        RUN_ROWH(4) RUN_ROWH(8) RUN_ROWH(16)
#undef RUN_ROWH
        CKL(p);
        return 0;
    }
    if (p->row_v8 && ((uintptr_t)img % esz2) == 0) {
        const int H = p->row_v8, RBI = ROWV_NT / (H / 8);
        const int ngroups = (p->d.N0 + RBI - 1) / RBI;
        const int nbatch = (ngroups + rowv.nit - 1) / rowv.nit;
        const int grid = std::min(nbatch, p->row_grid_limit > 0 ? p->row_grid_limit : work_sms(p));
#define RUN_ROWV(HH)                                                                                                   \
        if (H == HH) {                                                                                                 \
            if (dtype == SFFTB_F64) row_fwd_v8_kernel<double, TSt, HH><<<grid, ROWV_NT, p->smem_rowv, p->stream>>>(rowv, (const double*)img, out, nj); \
            else row_fwd_v8_kernel<float, TSt, HH><<<grid, ROWV_NT, p->smem_rowv, p->stream>>>(rowv, (const float*)img, out, nj);                     \
        }
        RUN_ROWV(256) RUN_ROWV(512) RUN_ROWV(1024) RUN_ROWV(2048)
#undef RUN_ROWV
        CKL(p);
        return 0;
    }
    const int grid = (p->d.N0 + p->row.RB - 1) / p->row.RB;
    if (dtype == SFFTB_F64)
        row_fwd_kernel<double, TSt><<<grid, 512, p->smem_row, p->stream>>>(rowg, (const double*)img, out, nj);
    else
        row_fwd_kernel<float, TSt><<<grid, 512, p->smem_row, p->stream>>>(rowg, (const float*)img, out, nj);
    CKL(p);
    return 0;
}

// Inverse row pass (C2R, scaling, background subtraction in real space).  hdiff != NULL (host GSS): the rows are
// produced in chunks and every finished chunk is copied to the host on the side stream while the next one is computed.
int launch_row_inv(sfftb_plan* p, const double* bpq, void* ddiff, int diff_dtype, void* hdiff) {
    typedef double2 TSt;                           // the apply step always works on fp64 spectra
    const sfftb_dims& d = p->d;
    const size_t osz2 = diff_dtype == SFFTB_F64 ? 16 : 8;
    if (p->row_fast && ((uintptr_t)ddiff % osz2) == 0) {
        const int H = p->row_fast, RB = ROWF_NT / (H / 16);
        const size_t sm = sizeof(cd) * (size_t)RB * (H + H / 16);
        const int nchunk = (hdiff && d.N0 >= 8 * RB) ? 4 : 1;
        const int rows_per = ((d.N0 + nchunk - 1) / nchunk + RB - 1) / RB * RB;
        const size_t esz = diff_dtype == SFFTB_F64 ? 8 : 4;
        for (int c = 0; c < nchunk; ++c) {
            const int row0 = c * rows_per, nrow = std::min(rows_per, d.N0 - row0);
            if (nrow <= 0) break;
            const int grid = (nrow + RB - 1) / RB;
            p->rinvf.row0 = row0;
#define RUN_RINVF(HH)                                                                                                  \
            if (H == HH) {                                                                                             \
                if (diff_dtype == SFFTB_F64) row_inv_fast_kernel<TSt, double, HH><<<grid, ROWF_NT, sm, p->stream>>>(p->rinvf, (const TSt*)p->gJa, bpq, (double*)ddiff); \
                else row_inv_fast_kernel<TSt, float, HH><<<grid, ROWF_NT, sm, p->stream>>>(p->rinvf, (const TSt*)p->gJa, bpq, (float*)ddiff);                            \
            }
            RUN_RINVF(512) RUN_RINVF(1024) RUN_RINVF(2048) RUN_RINVF(4096) RUN_RINVF(8192)
#undef RUN_RINVF
            CKL(p);
            if (hdiff) {
                CK(cudaEventRecord(p->evCopy[c & 3], p->stream));
                CK(cudaStreamWaitEvent(p->stream2, p->evCopy[c & 3], 0));
                CK(cudaMemcpyAsync((char*)hdiff + (size_t)row0 * d.N1 * esz, (const char*)ddiff + (size_t)row0 * d.N1 * esz,
                                   (size_t)nrow * d.N1 * esz, cudaMemcpyDeviceToHost, p->stream2));
            }
        }
        p->rinvf.row0 = 0;
        if (hdiff) {
            // blocking calls join the copy stream back; the asynchronous submissions (defer_join) leave the device-to-host
            // copies off the compute stream's critical path and record their completion event on the copy stream instead
            CK(cudaEventRecord(p->evJoin, p->stream2));
            if (!p->defer_join) CK(cudaStreamWaitEvent(p->stream, p->evJoin, 0));
        }
        return 0;
    }
    if (p->row_blu && ((uintptr_t)ddiff % osz2) == 0) {
        const int R = p->row_blu, RBI = R == 32 ? 1 : BLU_NT / (16 * R);
        const int ngroups = (d.N0 + RBI - 1) / RBI;
        const int grid = std::min(ngroups, p->row_grid_limit > 0 ? p->row_grid_limit : work_sms(p));
        const size_t smb = blu_smem_bytes(R);
        p->rinvf.Fpq = p->rinv.Fpq;
        if (R == 32) {
            if (diff_dtype == SFFTB_F64) row_inv_blu32_kernel<TSt, double><<<grid, BLU32_NT, smb, p->stream>>>(p->rowb, p->rinvf, (const TSt*)p->gJa, bpq, (double*)ddiff);
            else row_inv_blu32_kernel<TSt, float><<<grid, BLU32_NT, smb, p->stream>>>(p->rowb, p->rinvf, (const TSt*)p->gJa, bpq, (float*)ddiff);
        }
#define RUN_RINVB(RR)                                                                                                  \
        if (R == RR) {                                                                                                 \
            if (diff_dtype == SFFTB_F64) row_inv_blu_kernel<TSt, double, RR><<<grid, BLU_NT, smb, p->stream>>>(p->rowb, p->rinvf, (const TSt*)p->gJa, bpq, (double*)ddiff); \
            else row_inv_blu_kernel<TSt, float, RR><<<grid, BLU_NT, smb, p->stream>>>(p->rowb, p->rinvf, (const TSt*)p->gJa, bpq, (float*)ddiff);                            \
        }
        RUN_RINVB(4) RUN_RINVB(8) RUN_RINVB(16)
#undef RUN_RINVB
        CKL(p);
        return 0;
    }
    if (p->row_g16 && ((uintptr_t)ddiff % osz2) == 0) {
        const int R = p->row_g16, RBI = rowg_rbi(R);
        const int ngroups = (d.N0 + RBI - 1) / RBI;
        const int grid = std::min(ngroups, p->row_grid_limit > 0 ? p->row_grid_limit : work_sms(p));
        const size_t smg = rowg_smem_bytes(R, p->vtab != nullptr);
        p->rinvf.Fpq = p->rinv.Fpq;
#define RUN_RINVG(RR)                                                                                                  \
        if (R == RR) {                                                                                                 \
            if (diff_dtype == SFFTB_F64) row_inv_g16_kernel<TSt, double, RR><<<grid, RowgCfg<RR>::nt, smg, p->stream>>>(p->rinvf, p->tabB_row, (const TSt*)p->gJa, bpq, (double*)ddiff); \
            else row_inv_g16_kernel<TSt, float, RR><<<grid, RowgCfg<RR>::nt, smg, p->stream>>>(p->rinvf, p->tabB_row, (const TSt*)p->gJa, bpq, (float*)ddiff);                            \
        }
        RUN_RINVG(3) RUN_RINVG(5) RUN_RINVG(6) RUN_RINVG(10) RUN_RINVG(12)
#undef RUN_RINVG
        CKL(p);
        return 0;
    }
    const int grid = (d.N0 + p->row.RB - 1) / p->row.RB;
    if (diff_dtype == SFFTB_F64)
        row_inv_kernel<TSt, double><<<grid, 512, p->smem_row, p->stream>>>(p->rinv, (const TSt*)p->gJa, bpq, (double*)ddiff);
    else
        row_inv_kernel<TSt, float><<<grid, 512, p->smem_row, p->stream>>>(p->rinv, (const TSt*)p->gJa, bpq, (float*)ddiff);
    CKL(p);
    return 0;
}

template int launch_row_fwd<float2>(sfftb_plan*, const void*, int, float2*, int, const double*);
template int launch_row_fwd<double2>(sfftb_plan*, const void*, int, double2*, int, const double*);

// tu_fit.cu -- fit column pass (segmented warp-specialised kernel; folded-slice generic fallback) and the lag reductions.
#define SFFTB_TU_FIT
#include "plan.h"

int fit_setup(sfftb_plan* p) {
    const sfftb_dims& d = p->d;
    const bool f32 = p->cfg.storage == SFFTB_STORE_F32;
    if (init_generic_radix_tables()) return SFFTB_ECUDA;
    int occ = 1;
    if (p->fit_generic_ok) {
        if (f32) { if (set_smem(fit_col_kernel<float2>, p->smem_fit)) return SFFTB_ECUDA;
                   CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fit_col_kernel<float2>, NT_COL, p->smem_fit)); }
        else     { if (set_smem(fit_col_kernel<double2>, p->smem_fit)) return SFFTB_ECUDA;
                   CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fit_col_kernel<double2>, NT_COL, p->smem_fit)); }
    }
    p->grid_fit = std::min(d.N1 / 2 + 1, std::max(1, occ) * p->nsm);
    const size_t red_smem = sizeof(cd) * (size_t)(d.N1 / 2 + 1);
    if (set_smem(lag_reduce_kernel, red_smem) || set_smem(poly_reduce_kernel, red_smem)) return SFFTB_ECUDA;
    if (p->fit_seg) {
        if (set_smem(lag_reduce2_kernel, sizeof(cd) * LR2_KC * LR2_LB)) return SFFTB_ECUDA;
        if (d.DK == 3) {
            const bool bad = f32 ? (set_smem(fit_seg4_kernel<float2, 3, false, 0, 2>, p->smem_sfit3) || set_smem(fit_seg4_kernel<float2, 3, false, 2, 5>, p->smem_sfit3) ||
                                    set_smem(fit_seg4_kernel<float2, 3, false, 5, 10>, p->smem_sfit3))
                                 : (set_smem(fit_seg4_kernel<double2, 3, false, 0, 2>, p->smem_sfit3) || set_smem(fit_seg4_kernel<double2, 3, false, 2, 5>, p->smem_sfit3) ||
                                    set_smem(fit_seg4_kernel<double2, 3, false, 5, 10>, p->smem_sfit3));
            if (bad) return SFFTB_ECUDA;
        }
#define SET_SFIT3(DKK)                                                                                            \
        if (d.DK == DKK) {                                                                                            \
            if (f32) { if (set_smem(fit_seg4_kernel<float2, DKK>, p->smem_sfit3) || set_smem(fit_seg4_kernel<float2, DKK, true>, p->smem_sfit3)) return SFFTB_ECUDA; }       \
            else     { if (set_smem(fit_seg4_kernel<double2, DKK>, p->smem_sfit3) || set_smem(fit_seg4_kernel<double2, DKK, true>, p->smem_sfit3)) return SFFTB_ECUDA; }      \
        }
        SET_SFIT3(0) SET_SFIT3(1) SET_SFIT3(2)
#undef SET_SFIT3
        // shared-template tiles with cached template spectra (kernels_fit_jcache.cuh)
        p->grid_jc = 0;
        if (d.DK <= 2 && !env_int("SFFTB_NO_ASPEC_CACHE", 0) && jcache_smem_bytes() <= p->max_smem) {
            const size_t smj = jcache_smem_bytes();
            int occ = 0;
#define SET_JC(DKK)                                                                                               \
            if (d.DK == DKK) {                                                                                        \
                if (f32) { if (set_smem(aspec_cache_kernel<float2, DKK>, smj) || set_smem(fit_jonly_cached_kernel<float2, DKK>, smj)) return SFFTB_ECUDA;      \
                           CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fit_jonly_cached_kernel<float2, DKK>, JC_NT, smj)); }                     \
                else     { if (set_smem(aspec_cache_kernel<double2, DKK>, smj) || set_smem(fit_jonly_cached_kernel<double2, DKK>, smj)) return SFFTB_ECUDA;    \
                           CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fit_jonly_cached_kernel<double2, DKK>, JC_NT, smj)); }                    \
            }
            SET_JC(0) SET_JC(1) SET_JC(2)
#undef SET_JC
            p->grid_jc = std::max(1, occ);
        }
    }
    return 0;
}

// Cached A-role spectra of the template for the J-only column pass; returns false when the cache cannot be used (the caller
// falls back to fit_seg4_kernel<JONLY>).
template <typename TSt>
static bool jcache_ready(sfftb_plan* p, const TSt* gIsrc) {
    if (!p->grid_jc || p->aspec_off) return false;
    const sfftb_dims& d = p->d;
    const int NH = d.N1 / 2 + 1, Fij = d.Fij;
    const size_t need = (size_t)NH * p->sfit.nseg * Fij * FS3_M;
    // the cache is worth its memory for tile-sized templates (252 MB at 2048^2, 1 GB at 4096^2); beyond SFFTB_ASPEC_MAX_MB (4096) the
    // recomputing kernel stays
    if (sizeof(cd) * need > (size_t)env_int("SFFTB_ASPEC_MAX_MB", 4096) * 1048576ull) { p->aspec_off = 1; return false; }
    if (!p->aspec || p->aspec_elems < need) {
        if (p->aspec) { cudaFree(p->aspec); p->aspec = nullptr; }
        if (cudaMalloc(&p->aspec, sizeof(cd) * need) != cudaSuccess) { cudaGetLastError(); p->aspec = nullptr; p->aspec_off = 1; return false; }
        p->aspec_elems = need;
        p->aspec_epoch = -1;
    }
    if (p->aspec_epoch != p->tmpl_epoch) {
        const long long njobs = (long long)NH * p->sfit.nseg * Fij;
        const int grid = (int)std::min<long long>((njobs + JC_NHW - 1) / JC_NHW, (long long)work_sms(p) * p->grid_jc);
        const size_t smj = jcache_smem_bytes();
        if (d.DK == 0) aspec_cache_kernel<TSt, 0><<<grid, JC_NT, smj, p->stream>>>(p->sfit, gIsrc, p->aspec);
        else if (d.DK == 1) aspec_cache_kernel<TSt, 1><<<grid, JC_NT, smj, p->stream>>>(p->sfit, gIsrc, p->aspec);
        else aspec_cache_kernel<TSt, 2><<<grid, JC_NT, smj, p->stream>>>(p->sfit, gIsrc, p->aspec);
        if (cudaGetLastError() != cudaSuccess) { p->aspec_off = 1; return false; }
        p->launches++;
        p->aspec_epoch = p->tmpl_epoch;
    }
    return true;
}

// Column pass + contraction over k1 into the lag tables R, RJ, RT, RJT.
// jonly: tiles after the first of a shared-template batch -- the rows of the template-only pairs are already in kap2.
template <typename TSt>
int launch_fit_cols(sfftb_plan* p, const TSt* gIsrc, bool jonly) {
    const sfftb_dims& d = p->d;
    const int DK = d.DK;
    const int grid_sfit = std::min(p->grid_sfit, work_sms(p));
    if (p->fit_seg) {
        const int NH = d.N1 / 2 + 1, nms = (DK == 3 ? 5 : 4) * SFFTB_MAXE;
        const int nwarps = NH * (DK + 2);
        col_moments_kernel<TSt><<<(nwarps + 7) / 8, 256, 0, p->stream>>>(d.N0, NH, DK, d.DB, nms, jonly ? 1 : 0, gIsrc, (const TSt*)p->gJ, p->momg);
        CKL(p);
    }
    if (p->fit_seg) EVREC(p, EV_KFIT0);
    const bool jc = jonly && jcache_ready<TSt>(p, gIsrc);
    if (jc) {
        const int NHj = d.N1 / 2 + 1;
        const int grid = std::min(NHj, work_sms(p) * p->grid_jc);
        const size_t smj = jcache_smem_bytes();
        if (DK == 0) fit_jonly_cached_kernel<TSt, 0><<<grid, JC_NT, smj, p->stream>>>(p->sfit, gIsrc, (const TSt*)p->gJ, p->aspec, p->kap2);
        else if (DK == 1) fit_jonly_cached_kernel<TSt, 1><<<grid, JC_NT, smj, p->stream>>>(p->sfit, gIsrc, (const TSt*)p->gJ, p->aspec, p->kap2);
        else fit_jonly_cached_kernel<TSt, 2><<<grid, JC_NT, smj, p->stream>>>(p->sfit, gIsrc, (const TSt*)p->gJ, p->aspec, p->kap2);
    } else if (jonly) {
        if (DK == 0) fit_seg4_kernel<TSt, 0, true><<<grid_sfit, FS4_NT, p->smem_sfit3, p->stream>>>(p->sfit, p->vtabs, gIsrc, (const TSt*)p->gJ, p->kap2);
        else if (DK == 1) fit_seg4_kernel<TSt, 1, true><<<grid_sfit, FS4_NT, p->smem_sfit3, p->stream>>>(p->sfit, p->vtabs, gIsrc, (const TSt*)p->gJ, p->kap2);
        else fit_seg4_kernel<TSt, 2, true><<<grid_sfit, FS4_NT, p->smem_sfit3, p->stream>>>(p->sfit, p->vtabs, gIsrc, (const TSt*)p->gJ, p->kap2);
    } else if (p->fit_seg) {
        if (DK == 0) fit_seg4_kernel<TSt, 0><<<grid_sfit, FS4_NT, p->smem_sfit3, p->stream>>>(p->sfit, p->vtabs, gIsrc, (const TSt*)p->gJ, p->kap2);
        else if (DK == 1) fit_seg4_kernel<TSt, 1><<<grid_sfit, FS4_NT, p->smem_sfit3, p->stream>>>(p->sfit, p->vtabs, gIsrc, (const TSt*)p->gJ, p->kap2);
        else if (DK == 2) fit_seg4_kernel<TSt, 2><<<grid_sfit, FS4_NT, p->smem_sfit3, p->stream>>>(p->sfit, p->vtabs, gIsrc, (const TSt*)p->gJ, p->kap2);
        else {
            // KerPolyOrder = 3: 65 accumulators do not fit the product threads' registers; three launches over plane ranges
            fit_seg4_kernel<TSt, 3, false, 0, 2><<<grid_sfit, FS4_NT, p->smem_sfit3, p->stream>>>(p->sfit, p->vtabs, gIsrc, (const TSt*)p->gJ, p->kap2);
            CKL(p);
            fit_seg4_kernel<TSt, 3, false, 2, 5><<<grid_sfit, FS4_NT, p->smem_sfit3, p->stream>>>(p->sfit, p->vtabs, gIsrc, (const TSt*)p->gJ, p->kap2);
            CKL(p);
            fit_seg4_kernel<TSt, 3, false, 5, 10><<<grid_sfit, FS4_NT, p->smem_sfit3, p->stream>>>(p->sfit, p->vtabs, gIsrc, (const TSt*)p->gJ, p->kap2);
        }
    } else
        fit_col_kernel<TSt><<<p->grid_fit, NT_COL, p->smem_fit, p->stream>>>(p->cfit, gIsrc, (const TSt*)p->gJ, p->kap, p->lam, p->nuJ);
    CKL(p);
    if (p->fit_seg && p->sfit.mom_external && !(jonly && jc)) {
        const int NHm = d.N1 / 2 + 1;
        const int gridm = std::min(NHm, 8 * work_sms(p));
        if (DK == 0) col_poly_rows_kernel<TSt, 0><<<gridm, 256, 0, p->stream>>>(p->sfit, gIsrc, p->kap2, jonly ? 1 : 0);
        else if (DK == 1) col_poly_rows_kernel<TSt, 1><<<gridm, 256, 0, p->stream>>>(p->sfit, gIsrc, p->kap2, jonly ? 1 : 0);
        else if (DK == 2) col_poly_rows_kernel<TSt, 2><<<gridm, 256, 0, p->stream>>>(p->sfit, gIsrc, p->kap2, jonly ? 1 : 0);
        else col_poly_rows_kernel<TSt, 3><<<gridm, 256, 0, p->stream>>>(p->sfit, gIsrc, p->kap2, jonly ? 1 : 0);
        CKL(p);
    }
    if (p->fit_seg) EVREC(p, EV_KFIT1);
    EVREC(p, EV_COL);
    if (p->fit_seg) {
        // (one launch over all rows also for shared-template tiles: the reduction is a single latency-bound wave, and
        //  restricting it to the rows that changed measured slower: 0.082 vs 0.054 ms at 2048^2)
        dim3 grd((p->sfit.nrows + 15) / 16, p->red2.ksplit);
        lag_reduce2_kernel<<<grd, 256, sizeof(cd) * LR2_KC * LR2_LB, p->stream>>>(p->red2, p->kap2, p->part);
        CKL(p);
        const int tot = p->sfit.nrows * (4 * d.w1 + 1);
        lag_finish_kernel<<<(tot + 255) / 256, 256, 0, p->stream>>>(p->fin2, p->part);
        CKL(p);
    } else {
        const size_t red_smem = sizeof(cd) * (size_t)(d.N1 / 2 + 1);
        lag_reduce_kernel<<<p->nrowsK, 256, red_smem, p->stream>>>(p->red, p->kap, p->R, p->RJ);
        CKL(p);
        poly_reduce_kernel<<<p->nrowsL + d.DB + 1, 256, red_smem, p->stream>>>(p->pred, p->lam, p->nuJ, p->RT, p->RJT);
        CKL(p);
    }
    return 0;
}

int lag_reduce2_setup() { return set_smem(lag_reduce2_kernel, sizeof(cd) * LR2_KC * LR2_LB); }

int launch_lag_reduce2(sfftb_plan* p, const LagReduce2Args& a, const cd* kap, double* part) {
    dim3 grd((a.nrows + 15) / 16, a.ksplit);
    lag_reduce2_kernel<<<grd, 256, sizeof(cd) * LR2_KC * LR2_LB, p->stream>>>(a, kap, part);
    CKL(p);
    return 0;
}

template int launch_fit_cols<float2>(sfftb_plan*, const float2*, bool);
template int launch_fit_cols<double2>(sfftb_plan*, const double2*, bool);

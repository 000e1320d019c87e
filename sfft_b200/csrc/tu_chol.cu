// tu_chol.cu -- dense fp64 solve of the normal equations: cooperative Cholesky, cached-factor substitutions, LU fallback.
#define SFFTB_TU_CHOL
#include "plan.h"

int chol_setup(sfftb_plan* p) {
    {
        const int nblk = (p->nsolve + CC_NB - 1) / CC_NB;
        CK(cudaMalloc(&p->cholW, sizeof(double) * (size_t)nblk * CC_NB * CC_NB));
        CK(cudaMalloc(&p->cholY, sizeof(double) * (size_t)p->nsolve));
        CK(cudaMalloc(&p->cholX, sizeof(double) * (size_t)p->nsolve));
        CK(cudaMalloc(&p->cholBar, sizeof(unsigned) * 4));
        const size_t csm = sizeof(double) * (4 * CC_NB * CC_PITCH);
        CK(cudaFuncSetAttribute(chol_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
        int occ = 0, coop = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, chol_coop_kernel, CC_NT, csm));
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device));
        {
            CK(cudaMalloc(&p->substFlags, sizeof(unsigned) * 2 * (size_t)nblk));
            CK(cudaMemset(p->substFlags, 0, sizeof(unsigned) * 2 * (size_t)nblk));
            const size_t ssm = sizeof(double) * (3 * CC_NB * CC_DP + 64 + 64 + 256);
            CK(cudaFuncSetAttribute(chol_subst_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
            int occs = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occs, chol_subst_kernel, CC_NT, ssm));
            p->subst_ok = (occs >= 1 && coop && !env_int("SFFTB_RESOLVE_V1", 0)) ? 1 : 0;
            CK(cudaMalloc(&p->substMsg, sizeof(ulonglong2) * 2 * (size_t)nblk * CC_NB));
            CK(cudaMemset(p->substMsg, 0, sizeof(ulonglong2) * 2 * (size_t)nblk * CC_NB));
            const size_t ssm2 = sizeof(double) * (3 * CC_NB * CC_DP + 64 + 128);
            CK(cudaFuncSetAttribute(chol_subst2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm2));
            if (p->subst_ok && !env_int("SFFTB_RESOLVE_V2", 0)) p->subst_ok = 2;
        }
        p->chol_coop = (occ >= 1 && coop && !env_int("SFFTB_CHOL_LEGACY", 0)) ? std::min(occ, std::max(1, env_int("SFFTB_CHOL_CTAS", CC_CTAS_PER_SM))) : 0;
    }
    const sfftb_dims& d = p->d;
    const size_t bs_smem = sizeof(double) * ((size_t)p->nsolve + CH_NB + CH_NB * (CH_NB + 1));
    if (bs_smem > p->max_smem) return fail(SFFTB_EINVAL, "NEQ=%d too large for the back-substitution kernel", d.NEQ);
    if (set_smem(chol_backsolve_kernel, bs_smem)) return SFFTB_ECUDA;
    if (set_smem(lu_solve_kernel, sizeof(double) * (size_t)p->nsolve)) return SFFTB_ECUDA;
    return 0;
}

int run_cholesky(sfftb_plan* p, int resolve) {
    const int n = p->nsolve, ntot = n + 1;
    if (p->chol_coop && resolve && p->subst_ok == 2) {
        SubstArgs2 sa;
        sa.A = p->Aug; sa.ld = p->ld; sa.n = n; sa.W = p->cholW; sa.yv = p->cholY; sa.xs = p->cholX;
        sa.msg = p->substMsg; sa.epoch = ++p->substEpoch;
        sa.sc = p->sc; sa.idx = p->idxmap; sa.sol = p->sol; sa.NEQ = p->d.NEQ;
        const int nblk = (n + CC_NB - 1) / CC_NB;
        void* args[] = {&sa};
        // at most half of the SMs: two plans (TemplatePipeline) may run their substitutions at the same time, and two
        // cooperative grids must be able to be resident together; blocks beyond the grid are owned cyclically
        CK(cudaLaunchCooperativeKernel((void*)chol_subst2_kernel, dim3(std::min(nblk, std::max(1, p->nsm / 2))), dim3(CC_NT), args,
                                       sizeof(double) * (3 * CC_NB * CC_DP + 64 + 128), p->stream));
        p->launches++;
        return 0;
    }
    if (p->chol_coop && resolve && p->subst_ok) {
        SubstArgs sa;
        sa.A = p->Aug; sa.ld = p->ld; sa.n = n; sa.W = p->cholW; sa.yv = p->cholY; sa.xs = p->cholX;
        sa.flags = p->substFlags; sa.epoch = ++p->substEpoch;
        sa.sc = p->sc; sa.idx = p->idxmap; sa.sol = p->sol; sa.NEQ = p->d.NEQ;
        const int nblk = (n + CC_NB - 1) / CC_NB;
        void* args[] = {&sa};
        CK(cudaLaunchCooperativeKernel((void*)chol_subst_kernel, dim3(std::min(nblk, p->nsm)), dim3(CC_NT), args,
                                       sizeof(double) * (3 * CC_NB * CC_DP + 64 + 64 + 256), p->stream));
        p->launches++;
        return 0;
    }
    if (p->chol_coop) {
        CholArgs ca;
        ca.resolve = resolve;
        ca.A = p->Aug; ca.ld = p->ld; ca.n = n; ca.ntot = ntot; ca.W = p->cholW; ca.yv = p->cholY; ca.xs = p->cholX;
        ca.bar = p->cholBar; ca.info = p->info; ca.sc = p->sc; ca.idx = p->idxmap; ca.sol = p->sol; ca.NEQ = p->d.NEQ;
        CK(cudaMemsetAsync(p->cholBar, 0, sizeof(unsigned) * 4, p->stream));
        ca.dbg = nullptr;
        static unsigned long long* dbgbuf = nullptr;
        const bool dbg = env_int("SFFTB_CHOL_DBG", 0) != 0;
        if (dbg) {
            if (!dbgbuf) CK(cudaMalloc(&dbgbuf, sizeof(unsigned long long) * 2048));
            CK(cudaMemsetAsync(dbgbuf, 0, sizeof(unsigned long long) * 2048, p->stream));
            ca.dbg = dbgbuf;
        }
        void* args[] = {&ca};
        CK(cudaLaunchCooperativeKernel((void*)chol_coop_kernel, dim3(p->chol_grid_limit > 0 ? p->chol_grid_limit : p->nsm * p->chol_coop), dim3(CC_NT), args,
                                       sizeof(double) * (4 * CC_NB * CC_PITCH), p->stream));
        p->launches++;
        if (dbg) {
            std::vector<unsigned long long> hst(2048);
            CK(cudaStreamSynchronize(p->stream));
            CK(cudaMemcpy(hst.data(), dbgbuf, sizeof(unsigned long long) * 2048, cudaMemcpyDeviceToHost));
            const int nblk = (n + CC_NB - 1) / CC_NB;
            fprintf(stderr, "chol dbg (us): k  trsm  bar1  dsg_tile  dsg_potrf  others_tiles  bar2_end\n");
            for (int k = 0; k < nblk && k < 32; ++k) {
                const unsigned long long* t = &hst[8 * k];
                auto us = [&](int i) { return t[i] ? (double)(t[i] - t[0]) * 1e-3 : -1.0; };
                const unsigned long long* u = &hst[1024 + 4 * (k + 1)];
                fprintf(stderr, "  %2d  %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f | potrf(k+1): loaded %6.1f loop_end %6.1f\n", k, us(1), us(2), us(3), us(4), us(5), us(6),
                        u[0] ? (double)(u[0] - t[0]) * 1e-3 : -1.0, u[1] ? (double)(u[1] - t[0]) * 1e-3 : -1.0);
            }
        }
        return 0;
    }
    for (int k0 = 0; k0 < n; k0 += CH_NB) {
        const int kb = std::min(CH_NB, n - k0);
        const int below = ntot - (k0 + kb);
        const int gp = std::max(1, (below + 127) / 128);
        chol_panel_kernel<<<gp, 128, 0, p->stream>>>(p->Aug, p->ld, ntot, n, k0, p->info);
        CKL(p);
        if (below > 0 && k0 + kb < n) {
            const int nt = (below + 63) / 64;
            chol_update_kernel<<<dim3(nt, nt), 256, 0, p->stream>>>(p->Aug, p->ld, ntot, n, k0);
            CKL(p);
        }
    }
    const size_t bs_smem = sizeof(double) * ((size_t)n + CH_NB + CH_NB * (CH_NB + 1));
    chol_backsolve_kernel<<<1, 1024, bs_smem, p->stream>>>(p->Aug, p->ld, n, p->sc, p->idxmap, p->sol, p->d.NEQ);
    CKL(p);
    return 0;
}

int run_lu(sfftb_plan* p) {
    const int n = p->nsolve;
    int occ = 0;
    const size_t smem = sizeof(double) * (size_t)n;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lu_solve_kernel, 512, smem));
    if (occ < 1) return fail(SFFTB_ECUDA, "LU fallback kernel cannot be made resident");
    int grid = p->nsm;
    double* A = p->Aug; int ld = p->ld; int nn = n; double* du = p->diagU; const double* sc = p->sc; const int* idx = p->idxmap;
    double* sol = p->sol; int NEQ = p->d.NEQ; int* info = p->info;
    void* args[] = {&A, &ld, &nn, &du, &sc, &idx, &sol, &NEQ, &info};
    CK(cudaLaunchCooperativeKernel((void*)lu_solve_kernel, dim3(grid), dim3(512), args, smem, p->stream));
    p->launches++;
    return 0;
}

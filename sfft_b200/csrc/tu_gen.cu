// tu_gen.cu -- general-basis plans (B-spline / polynomial spatial variation of any degree, sfft/BSplineSFFT.py): host
// side of kernels_gen.cuh.  The row passes, the lag contraction (lag_reduce2_kernel lives in tu_fit.cu, wrapped by
// launch_lag_reduce2) and the dense solver are shared with the polynomial path.
#define SFFTB_TU_GEN
#include "plan.h"

struct GenState {
    int mode;
    int Fij, nsca, P;                // kernel planes, scaling (or sum) planes, column planes = Fij + nsca
    int nU, nVs;                     // rows of the U table, stored planes
    int Fp, Fq, Fpq;
    std::vector<int> cp_u, cp_vs;    // column plane -> (U row, stored plane)
    double *dU, *dV, *dP, *dQr;
    cd *dQ, *dTF;
    int *d_fu, *d_fv;
    std::vector<GenPass> passes;
    long long jobs_run, jobs_dense;  // transforms per column over all passes: really run / without the local-support skip
    GenFitArgs fit;
    size_t smem_fit;
    int nrows;
    cd* kap; double* part; double* Rall; double* PQ; double* PQpart;
    LagReduce2Args red2;
    GenFillArgs fill;
    std::vector<void*> owned;        // device allocations released with the plan
    GenFirArgs fir; size_t smem_fir; cd* taps; double* cA;
    GenBkg bkg;
    void* gP;                        // stored planes of the fit step [nVs][NH][N0] (storage type)
    void* gPa;                       // stored planes of the apply step, fp64 (alias of gP for fp64 storage)
};

template <typename T>
static int dev_upload(GenState* g, const std::vector<T>& h, T** out) {
    *out = nullptr;
    if (h.empty()) return 0;
    CK(cudaMalloc(out, sizeof(T) * h.size()));
    g->owned.push_back(*out);
    CK(cudaMemcpy(*out, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return 0;
}
template <typename T>
static int dev_alloc(GenState* g, size_t n, T** out) {
    CK(cudaMalloc(out, sizeof(T) * std::max<size_t>(n, 1)));
    g->owned.push_back(*out);
    return 0;
}

// Q[q][k1] = sum_c Qr[q][c] exp(-2 pi i k1 c / N1): one warp per output
__global__ void gen_qtable_kernel(int N1, int NH, int nq, const cd* __restrict__ tw1, const double* __restrict__ Qr, cd* __restrict__ Q) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= nq * NH) return;
    const int q = gw / NH, k1 = gw - q * NH;
    double sx = 0.0, sy = 0.0;
    for (int c = lane; c < N1; c += 32) {
        const cd w = tw1[(int)(((long long)k1 * c) % N1)];
        const double v = Qr[(size_t)q * N1 + c];
        sx = fma(v, w.x, sx);
        sy = fma(v, w.y, sy);
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    if (lane == 0) Q[(size_t)q * NH + k1] = cmake(sx, sy);
}

static int host_wrap(int r, int n) { r %= n; return r < 0 ? r + n : r; }

void gen_free(sfftb_plan* p) {
    GenState* g = (GenState*)p->gen;
    if (!g) return;
    for (void* q : g->owned) cudaFree(q);
    delete g;
    p->gen = nullptr;
}

static int check_basis(const sfftb_basis* b, const char* what) {
    if (!b || b->nu < 1 || b->nv < 1 || b->nf < 1 || !b->U || !b->V || !b->fu || !b->fv) return fail(SFFTB_EINVAL, "%s basis: null or empty", what);
    for (int k = 0; k < b->nf; ++k)
        if (b->fu[k] < 0 || b->fu[k] >= b->nu || b->fv[k] < 0 || b->fv[k] >= b->nv) return fail(SFFTB_EINVAL, "%s basis: function %d refers to a missing table row", what, k);
    return 0;
}

int gen_plan_create(sfftb_plan* p, const sfftb_config* cfg, const sfftb_basis* ker, const sfftb_basis* sca, const sfftb_basis* bkg, int mode) {
    int rc;
    if ((rc = check_basis(ker, "kernel")) || (rc = check_basis(bkg, "background"))) return rc;
    if (mode < 0 || mode > 3) return fail(SFFTB_EINVAL, "bad scaling mode %d", mode);
    if (mode == 3 && (rc = check_basis(sca, "scaling"))) return rc;
    if (mode == 3 && sca->nf > ker->nf) return fail(SFFTB_EINVAL, "the scaling basis must not have more functions than the kernel basis (ScaFij <= Fij)");
    if (bkg->nu > GEN_MAXP || bkg->nv > GEN_MAXQ) return fail(SFFTB_EINVAL, "background basis too large (%d x %d 1-D functions; at most %d x %d)", bkg->nu, bkg->nv, GEN_MAXP, GEN_MAXQ);
    if (4 * cfg->w0 + 32 > FS3_M) return fail(SFFTB_EINVAL, "kernel half width %d too large for the segmented column pass", cfg->w0);
    if ((rc = plan_init_common(p, cfg))) return rc;
    GenState* g = new GenState();
    p->gen = g;
    sfftb_dims& d = p->d;
    const int N0 = d.N0, N1 = d.N1, NH = N1 / 2 + 1, w0 = d.w0, w1 = d.w1;
    const bool f32 = cfg->storage == SFFTB_STORE_F32;
    const size_t csz = f32 ? sizeof(float2) : sizeof(double2);
    g->mode = mode;
    g->Fij = ker->nf; g->Fp = bkg->nu; g->Fq = bkg->nv; g->Fpq = bkg->nf;
    d.Fij = ker->nf; d.Fpq = bkg->nf; d.Fijab = d.Fij * d.Fab; d.NEQ = d.Fijab + d.Fpq;
    const int Fij = d.Fij, Fab = d.Fab, Fpq = d.Fpq, L1 = d.L1, c0 = w0 * L1 + w1;

    // ---- tables: U rows = kernel | scaling | sum; stored planes (V rows) likewise ----
    g->nsca = mode == 3 ? sca->nf : (mode == 2 ? 1 : 0);
    g->P = Fij + g->nsca;
    std::vector<double> hU, hV;
    hU.insert(hU.end(), ker->U, ker->U + (size_t)ker->nu * N0);
    hV.insert(hV.end(), ker->V, ker->V + (size_t)ker->nv * N1);
    g->nU = ker->nu; g->nVs = ker->nv;
    for (int k = 0; k < Fij; ++k) { g->cp_u.push_back(ker->fu[k]); g->cp_vs.push_back(ker->fv[k]); }
    if (mode == 3) {
        hU.insert(hU.end(), sca->U, sca->U + (size_t)sca->nu * N0);
        hV.insert(hV.end(), sca->V, sca->V + (size_t)sca->nv * N1);
        for (int k = 0; k < sca->nf; ++k) { g->cp_u.push_back(g->nU + sca->fu[k]); g->cp_vs.push_back(g->nVs + sca->fv[k]); }
        g->nU += sca->nu; g->nVs += sca->nv;
    } else if (mode == 2) {
        // the summed centre-tap column of TweakLS (:2202-2272): sum_ij I U_i V_j = I (sum_i U_i)(sum_j V_j) when the kernel
        // basis is the full tensor product; otherwise the sum is not separable and the mode is refused
        if (ker->nf != ker->nu * ker->nv) return fail(SFFTB_EINVAL, "the 'sum' stripe tweak needs a full tensor-product kernel basis");
        std::vector<double> su((size_t)N0, 0.0), sv((size_t)N1, 0.0);
        for (int i = 0; i < ker->nu; ++i) for (int r = 0; r < N0; ++r) su[r] += ker->U[(size_t)i * N0 + r];
        for (int j = 0; j < ker->nv; ++j) for (int c = 0; c < N1; ++c) sv[c] += ker->V[(size_t)j * N1 + c];
        hU.insert(hU.end(), su.begin(), su.end());
        hV.insert(hV.end(), sv.begin(), sv.end());
        g->cp_u.push_back(g->nU); g->cp_vs.push_back(g->nVs);
        g->nU += 1; g->nVs += 1;
    }
    if (g->nVs > GEN_MAXVS) return fail(SFFTB_EINVAL, "too many distinct column functions (%d; at most %d)", g->nVs, GEN_MAXVS);
    if (dev_upload(g, hU, &g->dU) || dev_upload(g, hV, &g->dV)) return SFFTB_ECUDA;
    std::vector<double> hP(bkg->U, bkg->U + (size_t)bkg->nu * N0), hQr(bkg->V, bkg->V + (size_t)bkg->nv * N1);
    std::vector<int> hfu(bkg->fu, bkg->fu + Fpq), hfv(bkg->fv, bkg->fv + Fpq);
    if (dev_upload(g, hP, &g->dP) || dev_upload(g, hQr, &g->dQr) || dev_upload(g, hfu, &g->d_fu) || dev_upload(g, hfv, &g->d_fv)) return SFFTB_ECUDA;
    if (dev_alloc(g, (size_t)g->Fq * NH, &g->dQ)) return SFFTB_ECUDA;
    {
        const int nwarps = g->Fq * NH;
        gen_qtable_kernel<<<(nwarps * 32 + 255) / 256, 256, 0, p->stream>>>(N1, NH, g->Fq, p->tw1, g->dQr, g->dQ);
        CKL(p);
    }
    GenBkg& bk = g->bkg;
    bk.N0 = N0; bk.N1 = N1; bk.Fp = g->Fp; bk.Fq = g->Fq; bk.Fpq = Fpq; bk.P = g->dP; bk.Qr = g->dQr; bk.fu = g->d_fu; bk.fv = g->d_fv;

    // ---- row passes: stored planes through the column tables, no background in the inverse row pass ----
    p->vtab = g->dV;
    p->rinv.DB = 0; p->rinv.Fpq = 0;
    if (rows_setup(p)) return SFFTB_ECUDA;
    if (dev_alloc(g, csz * (size_t)g->nVs * NH * N0 / sizeof(char), (char**)&g->gP)) return SFFTB_ECUDA;
    g->gPa = g->gP;
    if (f32 && dev_alloc(g, sizeof(double2) * (size_t)g->nVs * NH * N0, (char**)&g->gPa)) return SFFTB_ECUDA;

    // ---- segment geometry (as the polynomial segmented kernel) ----
    GenFitArgs& fa = g->fit;
    memset(&fa, 0, sizeof fa);
    fa.N0 = N0; fa.NH = NH; fa.w0 = w0; fa.h = 2 * w0;
    {
        const int Smax = (FS3_M - 2 * fa.h) & ~1;
        fa.nseg = (N0 + Smax - 1) / Smax;
        fa.S = (((N0 + fa.nseg - 1) / fa.nseg) + 1) & ~1;
        if (fa.S > Smax) fa.S = Smax;
        fa.nseg = (N0 + fa.S - 1) / fa.S;
    }
    d.fold = fa.nseg; d.sub_len = fa.S;
    fa.U = g->dU; fa.Fp = g->Fp; fa.Q = g->dQ; fa.plane_stride = (size_t)NH * N0;
    std::vector<std::vector<int>> f_of_p(g->Fp);
    for (int f = 0; f < Fpq; ++f) f_of_p[hfu[f]].push_back(f);
    std::vector<int> tpos(Fpq);
    for (int pp = 0; pp < g->Fp; ++pp) {
        fa.tq_n[pp] = (int)f_of_p[pp].size();
        for (size_t t = 0; t < f_of_p[pp].size(); ++t) { fa.tq_q[pp][t] = (unsigned char)hfv[f_of_p[pp][t]]; tpos[f_of_p[pp][t]] = (int)t; }
    }
    // B-role spectra of the background row functions (plain O(M^2) DFT on the host, once per plan)
    {
        const int M = FS3_M;
        std::vector<long double> cw(M), sw(M);
        const long double tp = 6.283185307179586476925286766559005768L;
        for (int e = 0; e < M; ++e) { cw[e] = cosl(tp * e / M); sw[e] = -sinl(tp * e / M); }
        std::vector<cd> hTF((size_t)fa.nseg * g->Fp * M);
        std::vector<double> win(M);
        for (int s = 0; s < fa.nseg; ++s)
            for (int pp = 0; pp < g->Fp; ++pp) {
                for (int n = 0; n < M; ++n) win[n] = hP[(size_t)pp * N0 + host_wrap(s * fa.S - fa.h + n, N0)];
                for (int k = 0; k < M; ++k) {
                    long double sx = 0.0L, sy = 0.0L;
                    for (int n = 0; n < M; ++n) { const int e = (k * n) & (M - 1); sx += win[n] * cw[e]; sy += win[n] * sw[e]; }
                    hTF[((size_t)s * g->Fp + pp) * M + k] = cmake((double)sx, (double)sy);
                }
            }
        if (dev_upload(g, hTF, &g->dTF)) return SFFTB_ECUDA;
        fa.TF = g->dTF;
    }

    // ---- passes: blocks of A slots x B slots over the pairs the fill needs ----
    const int P = g->P, nl0 = 4 * w0 + 1, nlj0 = 2 * w0 + 1, nl1 = 4 * w1 + 1;
    std::vector<int> order(P);
    for (int k = 0; k < P; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return g->cp_vs[x] != g->cp_vs[y] ? g->cp_vs[x] < g->cp_vs[y] : g->cp_u[x] < g->cp_u[y]; });
    struct Ent { int type, id; };                  // type 0: plane id, 1: J, 2: background row function id
    std::vector<Ent> uni;
    for (int k : order) uni.push_back({0, k});
    uni.push_back({1, 0});
    for (int pp = 0; pp < g->Fp; ++pp) uni.push_back({2, pp});
    auto src_of = [&](const Ent& e) { return e.type == 0 ? g->cp_vs[e.id] : (e.type == 1 ? -1 : -2); };
    auto chunk = [&](const std::vector<Ent>& v, int cap) {
        std::vector<std::vector<Ent>> out;
        std::vector<Ent> cur; std::vector<int> srcs;
        for (const Ent& e : v) {
            const int s = src_of(e);
            const bool newsrc = s != -2 && std::find(srcs.begin(), srcs.end(), s) == srcs.end();
            if ((int)cur.size() == cap || (newsrc && (int)srcs.size() == 2)) { out.push_back(cur); cur.clear(); srcs.clear(); }
            if (s != -2 && std::find(srcs.begin(), srcs.end(), s) == srcs.end()) srcs.push_back(s);
            cur.push_back(e);
        }
        if (!cur.empty()) out.push_back(cur);
        return out;
    };
    std::vector<Ent> aents;
    for (int k : order) aents.push_back({0, k});
    const auto agroups = chunk(aents, GEN_NA);
    const auto bgroups = chunk(uni, GEN_NB);
    std::vector<int> pairrow((size_t)P * P, -1), rowJ(P, -1), rowT((size_t)P * Fpq, -1);
    std::vector<int> pos(P);
    for (int k = 0; k < P; ++k) pos[order[k]] = k;
    int nrows = 0;
    for (const auto& ga : agroups)
        for (const auto& gb : bgroups) {
            GenPass ps;
            memset(&ps, 0, sizeof ps);
            for (int q = 0; q < GEN_NA * GEN_NB; ++q) ps.rowbase[q] = -1;
            std::vector<int> srcs;
            auto slot_of = [&](int s) {
                for (size_t i = 0; i < srcs.size(); ++i) if (srcs[i] == s) return (int)i;
                srcs.push_back(s);
                return (int)srcs.size() - 1;
            };
            ps.na = (int)ga.size();
            for (int a = 0; a < ps.na; ++a) { ps.a_u[a] = (short)g->cp_u[ga[a].id]; ps.a_src[a] = (short)slot_of(g->cp_vs[ga[a].id]); }
            // transformed B entries first
            std::vector<Ent> bs;
            for (const Ent& e : gb) if (e.type != 2) bs.push_back(e);
            ps.nbt = (int)bs.size();
            for (const Ent& e : gb) if (e.type == 2) bs.push_back(e);
            ps.nb = (int)bs.size();
            bool any = false;
            for (int b = 0; b < ps.nb; ++b) {
                const Ent& e = bs[b];
                ps.b_type[b] = (short)e.type;
                if (e.type == 0) { ps.b_u[b] = (short)g->cp_u[e.id]; ps.b_src[b] = (short)slot_of(g->cp_vs[e.id]); }
                else if (e.type == 1) { ps.b_u[b] = 0; ps.b_src[b] = (short)slot_of(-1); }
                else { ps.b_u[b] = (short)e.id; ps.b_src[b] = 0; }
                for (int a = 0; a < ps.na; ++a) {
                    const int A = ga[a].id;
                    int need = 0;
                    if (e.type == 0) {
                        const int B = e.id;
                        if (pos[A] <= pos[B] && pairrow[(size_t)A * P + B] < 0 && pairrow[(size_t)B * P + A] < 0) { pairrow[(size_t)A * P + B] = nrows; need = nl0; }
                    } else if (e.type == 1) {
                        if (rowJ[A] < 0) { rowJ[A] = nrows; need = nlj0; }
                    } else {
                        const int pp = e.id;
                        if (fa.tq_n[pp] > 0 && rowT[(size_t)A * Fpq + f_of_p[pp][0]] < 0) {
                            for (size_t t = 0; t < f_of_p[pp].size(); ++t) rowT[(size_t)A * Fpq + f_of_p[pp][t]] = nrows + (int)t * nlj0;
                            need = fa.tq_n[pp] * nlj0;
                        }
                    }
                    if (need) { ps.rowbase[a * GEN_NB + b] = nrows; nrows += need; any = true; }
                }
            }
            if (!any) continue;
            if ((int)srcs.size() > GEN_MAXSRC) return fail(SFFTB_EINVAL, "internal: pass needs %d staged planes", (int)srcs.size());
            ps.nsrc = (int)srcs.size();
            for (int i = 0; i < ps.nsrc; ++i) ps.src_plane[i] = srcs[i];
            ps.ninv = 0;
            for (int q = 0; q < GEN_NA * GEN_NB; ++q) if (ps.rowbase[q] >= 0) ps.inv_q[ps.ninv++] = (unsigned char)q;
            g->passes.push_back(ps);
        }
    // ---- local support: per pass, the transforms that are really run and the per-segment slot masks (GenPass::jobs) ----
    {
        const int M = FS3_M, nseg = fa.nseg, S = fa.S, h = fa.h;
        const int nUall = (int)(hU.size() / (size_t)N0);
        // zA[u][s]: U_u is non-zero somewhere in the core rows of segment s; zB[u][s]: ... in its 256-row window (halo included)
        const bool dense = env_int("SFFTB_GEN_DENSE", 0) != 0;       // parity hook: no skipping
        std::vector<char> zA((size_t)nUall * nseg, dense ? 1 : 0), zB((size_t)nUall * nseg, dense ? 1 : 0);
        for (int u = 0; u < nUall && !dense; ++u)
            for (int sg = 0; sg < nseg; ++sg) {
                const int c0 = sg * S, Sc = std::min(S, N0 - c0);
                for (int n = 0; n < M; ++n) {
                    if (hU[(size_t)u * N0 + host_wrap(c0 - h + n, N0)] == 0.0) continue;
                    zB[(size_t)u * nseg + sg] = 1;
                    if (n >= h && n < h + Sc) zA[(size_t)u * nseg + sg] = 1;
                }
            }
        std::vector<unsigned short> hjobs;
        std::vector<unsigned> hseg;
        std::vector<size_t> joff, soff;
        size_t njobs_all = 0, njobs_dense = 0;
        if (nseg > 4095) return fail(SFFTB_EINVAL, "too many column segments (%d)", nseg);
        for (GenPass& ps : g->passes) {
            joff.push_back(hjobs.size()); soff.push_back(hseg.size());
            const int NP = ps.na + ps.nbt;
            int nj = 0;
            for (int sg = 0; sg < nseg; ++sg) {
                unsigned am = 0, bm = 0;
                for (int a = 0; a < ps.na; ++a) if (zA[(size_t)ps.a_u[a] * nseg + sg]) am |= 1u << a;
                for (int b = 0; b < ps.nb; ++b) {
                    const bool nz = ps.b_type[b] == 0 ? zB[(size_t)ps.b_u[b] * nseg + sg] != 0 : true;     // J and the background tables: always
                    if (nz) bm |= 1u << b;
                }
                if (am == 0 || bm == 0) { am = 0; bm = 0; }               // nothing to accumulate in this segment
                unsigned cnt = 0;
                for (int pp = 0; pp < NP; ++pp) {
                    const bool run = pp < ps.na ? ((am >> pp) & 1u) : ((bm >> (pp - ps.na)) & 1u);
                    if (run) { hjobs.push_back((unsigned short)((sg << 4) | pp)); ++cnt; }
                }
                if (cnt == 0) { hjobs.push_back((unsigned short)(sg << 4)); cnt = 1; }    // keeps the ring protocol walking
                hseg.push_back(am | (bm << 5) | (cnt << 16));
                nj += (int)cnt;
            }
            ps.njobs = nj;
            njobs_all += (size_t)nj; njobs_dense += (size_t)NP * nseg;
        }
        unsigned short* djobs = nullptr; unsigned* dseg = nullptr;
        if (dev_upload(g, hjobs, &djobs) || dev_upload(g, hseg, &dseg)) return SFFTB_ECUDA;
        for (size_t i = 0; i < g->passes.size(); ++i) { g->passes[i].jobs = djobs + joff[i]; g->passes[i].seginfo = dseg + soff[i]; }
        g->jobs_run = (long long)njobs_all; g->jobs_dense = (long long)njobs_dense;
    }
    g->nrows = nrows;
    fa.nrows = nrows;
    if (dev_alloc(g, (size_t)NH * nrows, &g->kap)) return SFFTB_ECUDA;
    LagReduce2Args& r2 = g->red2;
    r2.N1 = N1; r2.NH = NH; r2.nrows = nrows; r2.w1 = w1; r2.tw1 = p->tw1; r2.rb0 = 0;
    r2.ksplit = (NH + LR2_KC - 1) / LR2_KC;
    if (dev_alloc(g, (size_t)r2.ksplit * nrows * nl1, &g->part) || dev_alloc(g, (size_t)nrows * nl1, &g->Rall) ||
        dev_alloc(g, (size_t)GEN_MAXP * GEN_MAXQ, &g->PQ) || dev_alloc(g, (size_t)GEN_MAXP * GEN_MAXQ * 4 * (size_t)p->nsm, &g->PQpart)) return SFFTB_ECUDA;
    g->smem_fit = gen_fit4_smem_bytes(f32);
    if (g->smem_fit > p->max_smem) return fail(SFFTB_EINVAL, "the general fit kernel needs %zu bytes of shared memory", g->smem_fit);
    if (f32) { if (set_smem(fit_gen4_kernel<float2>, g->smem_fit)) return SFFTB_ECUDA; }
    else     { if (set_smem(fit_gen4_kernel<double2>, g->smem_fit)) return SFFTB_ECUDA; }
    if (lag_reduce2_setup()) return SFFTB_ECUDA;

    // ---- unknowns of the solved system (TweakLS, :2171-2338, 3702-3747) ----
    std::vector<int> u_plane, u_ref0, u_nref, refs, first;
    std::vector<signed char> u_a, u_b, u_mod;
    for (int k = 0; k < d.Fijab; ++k) {
        const int A = k / Fab, ab = k - A * Fab;
        const int a = ab / L1 - w0, b = ab % L1 - w1;
        const bool centre = ab == c0;
        int plane = A, nref = 1;
        if (centre && mode == 1 && A > 0) continue;
        if (centre && mode == 2) { if (A > 0) continue; plane = Fij; nref = Fij; }
        if (centre && mode == 3) { if (A >= g->nsca) continue; plane = Fij + A; }
        u_plane.push_back(plane); u_a.push_back((signed char)a); u_b.push_back((signed char)b); u_mod.push_back(centre ? 0 : 1);
        u_ref0.push_back((int)refs.size()); u_nref.push_back(nref);
        if (nref == 1) refs.push_back(k);
        else for (int A2 = 0; A2 < Fij; ++A2) refs.push_back(A2 * Fab + c0);
        first.push_back(k);
    }
    for (int f = 0; f < Fpq; ++f) {
        u_plane.push_back(-1 - f); u_a.push_back(0); u_b.push_back(0); u_mod.push_back(0);
        u_ref0.push_back((int)refs.size()); u_nref.push_back(1); refs.push_back(d.Fijab + f); first.push_back(d.Fijab + f);
    }
    const int n = (int)u_plane.size();
    d.NEQ_FSfree = n;
    p->nsolve = n; p->ld = n + 1;
    GenFillArgs& gf = g->fill;
    memset(&gf, 0, sizeof gf);
    gf.n = n; gf.NEQ = d.NEQ; gf.Fijab = d.Fijab; gf.Fab = Fab; gf.Fij = Fij; gf.Fpq = Fpq; gf.L0 = d.L0; gf.L1 = L1; gf.w0 = w0; gf.w1 = w1; gf.nl1 = nl1; gf.P = P;
    {
        int *a1, *a2, *a3, *a4, *a5, *a6, *a7; signed char *c1, *c2, *c3;
        if (dev_upload(g, u_plane, &a1) || dev_upload(g, u_ref0, &a2) || dev_upload(g, u_nref, &a3) || dev_upload(g, refs, &a4) ||
            dev_upload(g, pairrow, &a5) || dev_upload(g, rowJ, &a6) || dev_upload(g, rowT, &a7) ||
            dev_upload(g, u_a, &c1) || dev_upload(g, u_b, &c2) || dev_upload(g, u_mod, &c3)) return SFFTB_ECUDA;
        gf.u_plane = a1; gf.u_ref0 = a2; gf.u_nref = a3; gf.refs = a4; gf.pairrow = a5; gf.rowJ = a6; gf.rowT = a7;
        gf.u_a = c1; gf.u_b = c2; gf.u_mod = c3;
    }
    gf.Rall = g->Rall; gf.PQ = g->PQ; gf.fu = g->d_fu; gf.fv = g->d_fv; gf.Fq = g->Fq;
    const double N = (double)N0 * (double)N1;
    gf.invN = 1.0 / N; gf.invN2 = 1.0 / (N * N); gf.invN3 = 1.0 / (N * N * N);
    {
        std::vector<long double> gu((size_t)g->Fp * g->Fp, 0.0L), gv((size_t)g->Fq * g->Fq, 0.0L);
        for (int a = 0; a < g->Fp; ++a) for (int b = 0; b < g->Fp; ++b) { long double s = 0; for (int r = 0; r < N0; ++r) s += (long double)hP[(size_t)a * N0 + r] * hP[(size_t)b * N0 + r]; gu[(size_t)a * g->Fp + b] = s; }
        for (int a = 0; a < g->Fq; ++a) for (int b = 0; b < g->Fq; ++b) { long double s = 0; for (int c = 0; c < N1; ++c) s += (long double)hQr[(size_t)a * N1 + c] * hQr[(size_t)b * N1 + c]; gv[(size_t)a * g->Fq + b] = s; }
        std::vector<double> phi((size_t)Fpq * Fpq);
        for (int a = 0; a < Fpq; ++a) for (int b = 0; b < Fpq; ++b)
            phi[(size_t)a * Fpq + b] = (double)(gu[(size_t)hfu[a] * g->Fp + hfu[b]] * gv[(size_t)hfv[a] * g->Fq + hfv[b]]);
        double* dphi;
        if (dev_upload(g, phi, &dphi)) return SFFTB_ECUDA;
        gf.PHI = dphi;
    }
    // solver workspaces (shared with the polynomial path)
    CK(cudaMalloc(&p->Aug, sizeof(double) * (size_t)(n + 1) * p->ld));
    CK(cudaMalloc(&p->sc, sizeof(double) * (size_t)n));
    CK(cudaMalloc(&p->diagU, sizeof(double) * (size_t)n));
    CK(cudaMalloc(&p->sol, sizeof(double) * (size_t)d.NEQ));
    CK(cudaMalloc(&p->idxmap, sizeof(int) * (size_t)n));
    CK(cudaMemcpy(p->idxmap, first.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
    if (chol_setup(p)) return SFFTB_ECUDA;

    // ---- apply planes, ordered by stored plane ----
    GenFirArgs& fr = g->fir;
    memset(&fr, 0, sizeof fr);
    fr.N0 = N0; fr.N1 = N1; fr.NH = NH; fr.w0 = w0; fr.w1 = w1; fr.L0 = d.L0; fr.nU = g->nU; fr.U = g->dU; fr.tw1 = p->tw1; fr.plane_stride = (size_t)NH * N0;
    {
        std::vector<int> ap;                       // column planes used by the apply step
        for (int k = 0; k < Fij; ++k) ap.push_back(k);
        if (mode == 3) for (int k = 0; k < g->nsca; ++k) ap.push_back(Fij + k);
        std::stable_sort(ap.begin(), ap.end(), [&](int x, int y) { return g->cp_vs[x] < g->cp_vs[y]; });
        std::vector<short> apu; std::vector<int> apsol, apcen;
        int nvs = 0, last = -1;
        for (size_t i = 0; i < ap.size(); ++i) {
            const int k = ap[i];
            if (g->cp_vs[k] != last) { fr.vs_id[nvs] = (short)g->cp_vs[k]; fr.vs_first[nvs] = (short)i; ++nvs; last = g->cp_vs[k]; }
            apu.push_back((short)g->cp_u[k]);
            if (k < Fij) { apsol.push_back(k * Fab); apcen.push_back(mode == 3 ? -1 : k * Fab + c0); }
            else { apsol.push_back(-1 - ((k - Fij) * Fab + c0)); apcen.push_back(-1); }
        }
        fr.vs_first[nvs] = (short)ap.size();
        fr.nvs = nvs; fr.nap = (int)ap.size();
        short* d1; int *d2, *d3;
        if (dev_upload(g, apu, &d1) || dev_upload(g, apsol, &d2) || dev_upload(g, apcen, &d3)) return SFFTB_ECUDA;
        fr.ap_u = d1; fr.ap_sol = d2; fr.ap_centre = d3;
    }
    if (dev_alloc(g, (size_t)NH * fr.nap * d.L0, &g->taps) || dev_alloc(g, (size_t)fr.nap + 1, &g->cA)) return SFFTB_ECUDA;
    {
        g->smem_fir = gen_fir_smem_bytes(fr.nap, d.L0, fr.nU, fr.nvs, w0);
        if (g->smem_fir > p->max_smem) return fail(SFFTB_EINVAL, "the general FIR kernel needs %zu bytes of shared memory", g->smem_fir);
        if (set_smem(gen_fir_kernel<double2>, g->smem_fir)) return SFFTB_ECUDA;
    }
    p->grid_sfit = std::min(NH, p->nsm);
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

// [0] passes of the fit column kernel, [1] stored planes staged summed over the passes, [2] lag rows per column,
// [3] stored planes, [4] column planes, [5] unknowns of the solved system, [6] apply planes, [7] stored planes read by the FIR
void gen_info(const sfftb_plan* p, int* out) {
    const GenState* g = (const GenState*)p->gen;
    int ns = 0;
    for (const GenPass& ps : g->passes) ns += ps.nsrc;
    out[0] = (int)g->passes.size(); out[1] = ns; out[2] = g->nrows; out[3] = g->nVs; out[4] = g->P; out[5] = p->nsolve;
    out[6] = g->fir.nap; out[7] = g->fir.nvs;
}
int gen_nvs(const sfftb_plan* p) { return ((const GenState*)p->gen)->nVs; }
void* gen_planes(const sfftb_plan* p) { return ((const GenState*)p->gen)->gP; }
void* gen_planes_apply(const sfftb_plan* p) { return ((const GenState*)p->gen)->gPa; }

int gen_set_regularizer(sfftb_plan* p) {
    GenState* g = (GenState*)p->gen;
    g->fill.SST = p->fill.SST; g->fill.iREG = p->fill.iREG; g->fill.CSST = p->fill.CSST; g->fill.DSST = p->fill.DSST; g->fill.regw = p->fill.regw;
    return 0;
}

// J x T moments in real space (the image of J is still on the device when the fit runs)
int gen_rjt(sfftb_plan* p, const void* dJ, int dtype) {
    GenState* g = (GenState*)p->gen;
    const int grid = std::min(p->d.N0, 4 * p->nsm), n = g->Fp * g->Fq;
    if (dtype == SFFTB_F64) gen_rjt_kernel<double><<<grid, 256, 0, p->stream>>>(g->bkg, (const double*)dJ, g->PQpart);
    else gen_rjt_kernel<float><<<grid, 256, 0, p->stream>>>(g->bkg, (const float*)dJ, g->PQpart);
    CKL(p);
    gen_rjt_finish_kernel<<<(n + 127) / 128, 128, 0, p->stream>>>(grid, n, g->PQpart, g->PQ);
    CKL(p);
    return 0;
}

// all passes of the column kernel + contraction over k1 into Rall
template <typename TSt>
int gen_fit_cols(sfftb_plan* p) {
    GenState* g = (GenState*)p->gen;
    for (const GenPass& ps : g->passes) {
        fit_gen4_kernel<TSt><<<p->grid_sfit, FS4_NT, g->smem_fit, p->stream>>>(g->fit, ps, p->tabA, (const TSt*)g->gP, (const TSt*)p->gJ, g->kap);
        CKL(p);
    }
    EVREC(p, EV_COL);
    if (launch_lag_reduce2(p, g->red2, g->kap, g->part)) return SFFTB_ECUDA;
    const size_t tot = (size_t)g->nrows * g->fill.nl1;
    gen_lag_finish_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, p->stream>>>(g->nrows, g->fill.nl1, g->red2.ksplit, g->part, g->Rall);
    CKL(p);
    return 0;
}
template int gen_fit_cols<float2>(sfftb_plan*);
template int gen_fit_cols<double2>(sfftb_plan*);

int gen_fill_system(sfftb_plan* p) {
    GenState* g = (GenState*)p->gen;
    const int n = p->nsolve;
    gen_fill_diag_kernel<<<(n + 127) / 128, 128, 0, p->stream>>>(g->fill, p->sc, p->info);
    CKL(p);
    dim3 blk(32, 8), grd((n + 1 + 31) / 32, (n + 1 + 7) / 8);
    gen_fill_matrix_kernel<<<grd, blk, 0, p->stream>>>(g->fill, p->sc, p->Aug, p->ld, p->info);
    CKL(p);
    return 0;
}

int gen_restore(sfftb_plan* p) {
    GenState* g = (GenState*)p->gen;
    if (g->mode != 2) return 0;
    gen_restore_kernel<<<(p->nsolve + 255) / 256, 256, 0, p->stream>>>(g->fill, p->sol);
    CKL(p);
    return 0;
}

// unscaled tweaked system for the parity hook: L (n x n) and b (n) into a device buffer of (n + 1) x (n + 1) doubles
int gen_export(sfftb_plan* p, double* buf) {
    GenState* g = (GenState*)p->gen;
    const int n = p->nsolve;
    dim3 blk(32, 8), grd((n + 1 + 31) / 32, (n + 1 + 7) / 8);
    gen_fill_matrix_kernel<<<grd, blk, 0, p->stream>>>(g->fill, nullptr, buf, n + 1, p->info + 3);
    CKL(p);
    return 0;
}

int gen_fir(sfftb_plan* p, const double* dsol) {
    typedef double2 TSt;                           // the apply step always works on fp64 spectra
    GenState* g = (GenState*)p->gen;
    const sfftb_dims& d = p->d;
    const int NH = d.N1 / 2 + 1;
    gen_taps_kernel<<<NH, 128, 0, p->stream>>>(g->fir, dsol, g->taps, g->cA);
    CKL(p);
    dim3 grd(NH, (d.N0 + GFIR_CH - 1) / GFIR_CH);
    gen_fir_kernel<TSt><<<grd, GFIR_NT, g->smem_fir, p->stream>>>(g->fir, (const TSt*)g->gPa, (const TSt*)p->gJa, g->taps, g->cA, (TSt*)p->gJa);
    CKL(p);
    return 0;
}

int gen_bkg_subtract(sfftb_plan* p, const double* bf, void* ddiff, int diff_dtype) {
    GenState* g = (GenState*)p->gen;
    dim3 grd((p->d.N1 + 255) / 256, p->d.N0);
    if (diff_dtype == SFFTB_F64) gen_bkg_subtract_kernel<double><<<grd, 256, 0, p->stream>>>(g->bkg, bf, (double*)ddiff);
    else gen_bkg_subtract_kernel<float><<<grd, 256, 0, p->stream>>>(g->bkg, bf, (float*)ddiff);
    CKL(p);
    return 0;
}

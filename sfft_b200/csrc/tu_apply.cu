// tu_apply.cu -- the subtract column pass: FIR in the row-spectrum domain (kernels_apply.cuh).
#define SFFTB_TU_APPLY
#include "plan.h"

int apply_setup(sfftb_plan* p) {
    const sfftb_dims& d = p->d;
    // the apply step always works on fp64 spectra (plan.h: gIa / gJa)
    if (set_smem(apply_fir_kernel<double2>, p->smem_fir)) return SFFTB_ECUDA;
    if (p->smem_fir3 <= p->max_smem) {
#define SET_FIR3(DKK)                                                                                             \
        if (d.DK == DKK) {                                                                                            \
            if (set_smem(apply_fir3_kernel<double2, DKK>, p->smem_fir3)) return SFFTB_ECUDA;                          \
        }
        SET_FIR3(0) SET_FIR3(1) SET_FIR3(2) SET_FIR3(3)
#undef SET_FIR3
    }
    return 0;
}

int launch_fir(sfftb_plan* p, const double2* gIsrc, const double* dsol) {
    typedef double2 TSt;
    const sfftb_dims& d = p->d;
    if (p->smem_fir3 <= p->max_smem && !env_int("SFFTB_FIR_V1", 0)) {
        fir_taps_kernel<<<d.N1 / 2 + 1, 128, 0, p->stream>>>(p->fir, dsol, p->firTaps, p->firCA);
        CKL(p);
        dim3 grd(d.N1 / 2 + 1, (d.N0 + FIR3_CH - 1) / FIR3_CH);
#define RUN_FIR3(DKK) if (d.DK == DKK) apply_fir3_kernel<TSt, DKK><<<grd, FIR3_NT, p->smem_fir3, p->stream>>>(p->fir, gIsrc, (const TSt*)p->gJa, p->firTaps, p->firCA, (TSt*)p->gJa);
        RUN_FIR3(0) RUN_FIR3(1) RUN_FIR3(2) RUN_FIR3(3)
#undef RUN_FIR3
    } else {
        dim3 grd(d.N1 / 2 + 1, (d.N0 + FIR_CHUNK - 1) / FIR_CHUNK);
        apply_fir_kernel<TSt><<<grd, FIR_NT, p->smem_fir, p->stream>>>(p->fir, gIsrc, (const TSt*)p->gJa, dsol, (TSt*)p->gJa);
    }
    CKL(p);
    return 0;
}


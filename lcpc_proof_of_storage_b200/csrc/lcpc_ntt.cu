// Field dispatch of the Ligero encoder; the kernels live in lcpc_ntt_impl.cuh and are instantiated per field in
// lcpc_ntt_f0.cu .. lcpc_ntt_f4.cu.
#include "lcpc_field.cuh"
#include "lcpc_kernels.h"

namespace lcpc {

template <int FID>
cudaError_t plan_build_t(NttPlan &plan, int log_n, const uint64_t *root_mont, const Launch &lc);
template <int FID>
cudaError_t encode_t(const NttPlan &plan, const uint64_t *src, size_t src_stride, size_t src_valid, uint64_t *dst,
                     size_t n_rows, const Launch &lc, const ScatterDst *scatter);

template <int FID>
cudaError_t decode_t(const NttPlan &plan, uint64_t *data, size_t n_rows, const Launch &lc);

#define LCPC_FIELD_SWITCH(fid, CALL)                  \
    switch (fid) {                                    \
    case FT63: return CALL(FT63);                     \
    case FT127: return CALL(FT127);                   \
    case FT191: return CALL(FT191);                   \
    case FT255: return CALL(FT255);                   \
    case FT253_192: return CALL(FT253_192);           \
    default: return cudaErrorInvalidValue;            \
    }

cudaError_t ntt_plan_build(NttPlan &plan, int fid, int log_n, const uint64_t *root_mont, const Launch &lc) {
#define CALL(F) plan_build_t<F>(plan, log_n, root_mont, lc)
    LCPC_FIELD_SWITCH(fid, CALL)
#undef CALL
}

void ntt_plan_free(NttPlan &plan) {
    if (plan.d_tw) cudaFree(plan.d_tw);
    plan.d_tw = nullptr;
}

cudaError_t ntt_encode(const NttPlan &plan, const uint64_t *src, size_t src_stride, size_t src_valid,
                       uint64_t *dst, size_t n_rows, const Launch &lc, const ScatterDst *scatter) {
#define CALL(F) encode_t<F>(plan, src, src_stride, src_valid, dst, n_rows, lc, scatter)
    LCPC_FIELD_SWITCH(plan.fid, CALL)
#undef CALL
}

cudaError_t ntt_decode(const NttPlan &plan, uint64_t *data, size_t n_rows, const Launch &lc) {
#define CALL(F) decode_t<F>(plan, data, n_rows, lc)
    LCPC_FIELD_SWITCH(plan.fid, CALL)
#undef CALL
}

}  // namespace lcpc

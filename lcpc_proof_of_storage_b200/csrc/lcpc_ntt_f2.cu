// NTT kernels of the FT191 field (explicit instantiation of lcpc_ntt_impl.cuh).
#include "lcpc_ntt_impl.cuh"

namespace lcpc {
template cudaError_t plan_build_t<FT191>(NttPlan &, int, const uint64_t *, const Launch &);
template cudaError_t encode_t<FT191>(const NttPlan &, const uint64_t *, size_t, size_t, uint64_t *, size_t, const Launch &,
                                     const ScatterDst *);
template cudaError_t decode_t<FT191>(const NttPlan &, uint64_t *, size_t, const Launch &);
}  // namespace lcpc

// Column-kernel instantiations of the fused EvalMult(ct,ct) pipeline, part c (see fused_mul.cuh).
#include "fused_mul.cuh"

namespace psi {

cudaError_t dispatch_cols_c(PSI_COLS_ARGS) {
    PSI_COLS_CASE_BIG(4, 5)
    PSI_COLS_CASE_BIG(5, 5)
    PSI_COLS_CASE_BIG(5, 6)
    return cudaErrorInvalidValue;
}

}  // namespace psi

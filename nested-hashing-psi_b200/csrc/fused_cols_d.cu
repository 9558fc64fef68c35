// Column-kernel instantiations of the fused EvalMult(ct,ct) pipeline, part d (see fused_mul.cuh).
#include "fused_mul.cuh"

namespace psi {

cudaError_t dispatch_cols_d(PSI_COLS_ARGS) {
    PSI_COLS_CASE_BIG(6, 6)
    PSI_COLS_CASE_BIG(6, 7)
    PSI_COLS_CASE_BIG(7, 7)
    return cudaErrorInvalidValue;
}

}  // namespace psi

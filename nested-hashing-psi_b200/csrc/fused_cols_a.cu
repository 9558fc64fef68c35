// Column-kernel instantiations of the fused EvalMult(ct,ct) pipeline, part a (see fused_mul.cuh).
#include "fused_mul.cuh"

namespace psi {

cudaError_t dispatch_cols_a(PSI_COLS_ARGS) {
    PSI_COLS_CASE(1, 1)
    PSI_COLS_CASE(1, 2)
    PSI_COLS_CASE(2, 2)
    PSI_COLS_CASE(2, 3)
    PSI_COLS_CASE(3, 3)
    return cudaErrorInvalidValue;
}

}  // namespace psi

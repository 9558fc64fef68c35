// Column-kernel instantiations of the fused EvalMult(ct,ct) pipeline, part b (see fused_mul.cuh).
#include "fused_mul.cuh"

namespace psi {

cudaError_t dispatch_cols_b(PSI_COLS_ARGS) {
    PSI_COLS_CASE(3, 4)
    PSI_COLS_CASE(4, 4)
    return cudaErrorInvalidValue;
}

}  // namespace psi

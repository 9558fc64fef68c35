// Internal helpers shared by the host-side translation units of libpsi_b200.so.
#pragma once
#include <cstdint>
#include <string>

#include "psi_b200.h"

namespace psi {

// Records the message returned by psi_last_error() and passes the status through.
int set_error(int status, const std::string& msg);

bool is_prime_u64(uint64_t n);
uint64_t min_primitive_root(uint64_t m, uint64_t q);
int generate_params(uint32_t N, uint64_t t, uint32_t depth, uint32_t L_override, psi_params* out, uint32_t mult_technique,
                    uint32_t ks_technique, uint32_t fp_contract);

}  // namespace psi

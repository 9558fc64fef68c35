// Helpers shared by the reference-side adapters (BatchedFHEHIPPIE_b200.cpp, FHEHIPPIE_b200.cpp): status -> exception,
// limb vectors of lbcrypto objects as plain word pointers, psi_params of a BFVrns context.
#pragma once
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "openfhe.h"
#include "psi_b200.h"

#ifndef FHEEncType
#define FHEEncType lbcrypto::DCRTPoly
#endif

namespace psi_adapter {
using namespace lbcrypto;

inline void ck(int rc) {
    if (rc == PSI_OK) return;
    if (rc == PSI_ERR_INVALID) throw std::invalid_argument(psi_last_error());
    throw std::runtime_error(psi_last_error());
}

// NativeVector stores N machine words back to back (NativeInteger is one uint64_t): the limb vector itself is what
// the GPU library reads / fills, no per-coefficient ConvertToInt() loop
inline const uint64_t* words_of(const NativeVector& v) { return reinterpret_cast<const uint64_t*>(&v[0]); }
inline uint64_t* words_of(NativeVector& v) { return reinterpret_cast<uint64_t*>(&v[0]); }

// psi_params of the context.  Moduli and roots are read from the element parameters; the integer tables are a
// function of the moduli (psi_params_from_moduli); the three double tables are copied from OpenFHE's own (so that
// they are bit-identical to what the host library rounds with) unless PSI_B200_TABLES=derived.
// need_mult: the caller multiplies ciphertexts (the batched PIE); the non-batched PIE only key-switches
inline psi_params params_from(const CryptoContext<FHEEncType>& cc, bool need_mult = true) {
    const auto cp = std::dynamic_pointer_cast<CryptoParametersBFVRNS>(cc->GetCryptoParameters());
    if (!cp) throw std::invalid_argument("PIE adapter (B200): the context is not BFVrns");
    if (need_mult && cp->GetMultiplicationTechnique() != HPSPOVERQ)
        throw std::invalid_argument("PIE adapter (B200): only MultiplicationTechnique HPSPOVERQ is implemented");
    if (cp->GetKeySwitchTechnique() != BV || cp->GetDigitSize() != 0)
        throw std::invalid_argument("PIE adapter (B200): only BV key switching with digit size 0 is implemented");
    const auto& tq = cp->GetElementParams()->GetParams();
    const auto& tp = cp->GetParamsRl(0)->GetParams();
    if (tq.size() > PSI_MAX_LIMBS || tp.size() > PSI_MAX_LIMBS) throw std::invalid_argument("PIE adapter (B200): too many RNS limbs");
    uint64_t q[PSI_MAX_LIMBS], psiq[PSI_MAX_LIMBS], p[PSI_MAX_LIMBS], psip[PSI_MAX_LIMBS];
    for (size_t i = 0; i < tq.size(); i++) {
        q[i] = tq[i]->GetModulus().ConvertToInt();
        psiq[i] = tq[i]->GetRootOfUnity().ConvertToInt();
    }
    for (size_t j = 0; j < tp.size(); j++) {
        p[j] = tp[j]->GetModulus().ConvertToInt();
        psip[j] = tp[j]->GetRootOfUnity().ConvertToInt();
    }
    psi_params P;
    ck(psi_params_from_moduli(cc->GetRingDimension(), cp->GetPlaintextModulus(), (uint32_t)tq.size(), q, psiq, (uint32_t)tp.size(), p,
                              psip, cp->GetEncodingParams()->GetPlaintextRootOfUnity(), &P));
    const char* mode = std::getenv("PSI_B200_TABLES");
    if (need_mult && (!mode || std::strcmp(mode, "derived") != 0)) {  // the double tables only matter for EvalMult(ct, ct)
        const auto& qInv = cp->GetqInv();
        const auto& rInv = cp->GetrInv();
        const auto& frac = cp->GettQlSlHatInvModsDivsFrac(0);
        if (qInv.size() < tq.size() || rInv.size() < tp.size() || frac.size() < tp.size())
            throw std::runtime_error("PIE adapter (B200): HPS double tables missing in the context");
        for (size_t i = 0; i < tq.size(); i++) P.qInv[i] = qInv[i];
        for (size_t j = 0; j < tp.size(); j++) {
            P.pInv[j] = rInv[j];
            P.tQSHatInvModsDivsFrac[j] = frac[j];
        }
    }
    return P;
}

// limb pointers of one ciphertext in [comp][limb] order; checks what run() relies on
inline void limbs_of(const Ciphertext<FHEEncType>& ct, size_t L, size_t N, const uint64_t** out) {
    if (!ct) throw std::invalid_argument("PIE adapter (B200): null ciphertext in the query");
    const auto& cv = ct->GetElements();
    if (cv.size() != 2) throw std::invalid_argument("PIE adapter (B200): query ciphertexts must have two components");
    for (size_t c = 0; c < 2; c++) {
        if (cv[c].GetFormat() != Format::EVALUATION || cv[c].GetNumOfElements() != L)
            throw std::invalid_argument("PIE adapter (B200): query ciphertexts must be fresh (EVALUATION, all limbs)");
        for (size_t l = 0; l < L; l++) {
            const NativeVector& v = cv[c].GetElementAtIndex((usint)l).GetValues();
            if (v.GetLength() != N) throw std::invalid_argument("PIE adapter (B200): limb length differs from the ring dimension");
            out[c * L + l] = words_of(v);
        }
    }
}

}  // namespace psi_adapter

// Drop-in mirror of the reference's PIE operator for the batched FHE path:
//   class BatchedFHEHIPPIE   /root/reference/src/Common/Crypto/PrivateIndexedEqualityCheck/BatchedFHEHIPPIE.hpp:18-48
// Same five entry points, same argument meaning, same exceptions (std::invalid_argument for a stash
// or combined tables, BatchedFHEHIPPIE.cpp:13-21).  What differs is underneath: the object owns no
// OpenFHE plaintexts/ciphertexts; the constructor encodes the transposed nested cuckoo table on the
// GPU (psi_db_encode_slots), setIndex / setMinusCompareElement upload limbs, run() enqueues the
// sm_100a kernels and getResultList() reads the b result ciphertexts back.
//
// OpenFHE handle types are replaced by limb containers with the layout a DCRTPoly exposes through
// GetElementAtIndex(l).GetValues(): Ciphertext = [2][L][N] u64, EVALUATION format (INTEGRATION.md
// shows the adapter that converts in both directions when OpenFHE is present).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "hashing.hpp"
#include "psi_b200.h"

namespace psi {

// Stands in for lbcrypto::CryptoContext<DCRTPoly> as far as the PIE uses it: the BFV parameters
// (GetCryptoParameters()->GetPlaintextModulus(), BatchedFHEHIPPIE.cpp:43) and the evaluator.
struct CryptoContext {
    psi_params params;
    psi_ctx* device_ctx;           // owned by the caller (the server object owns the context in the reference too)
    psi_multi* multi = nullptr;    // when set, the evaluator is a single-process multi-device one (a device list);
                                   // device_ctx is then unused
    uint64_t GetPlaintextModulus() const { return params.t; }
};

// pK is stored but never used by run() (BatchedFHEHIPPIE.hpp:22); kept for signature parity.
struct PublicKey {};

typedef std::shared_ptr<std::vector<uint64_t>> Ciphertext;  // [2][L][N] limbs, EVALUATION

class BatchedFHEHIPPIE {
   protected:
    CryptoContext& cryptoContext;
    PublicKey& pK;
    std::vector<Ciphertext> resultList;
    std::vector<std::vector<Ciphertext>> indexMatrix;  // chfN x chfIndex
    Ciphertext minusCompareElement;
    uint32_t K, b, E, nslots;
    bool uploaded = false;

   public:
    // The bin shuffle and the masks are SECURITY-CRITICAL randomness: a client who knows the mask r can divide it
    // out of r * prod(y_hf - x) and learn about non-matching server items.  Like the reference
    // (BatchedFHEHIPPIE.cpp:25-26) the default draws both seeds from std::random_device (PSI_SEED_RANDOM).
    // Explicit seeds exist for TESTS ONLY (they make the database reproducible).  keepSlots retains the slot
    // vectors the constructor built (tests compare them with the oracle's transposition).
    BatchedFHEHIPPIE(CryptoContext& cryptor, PublicKey& pK, HierarchicalCuckooHashTable& hct,
                     uint64_t shuffleSeed = PSI_SEED_RANDOM, uint64_t maskSeed = PSI_SEED_RANDOM, bool keepSlots = false);

    void run();

    std::vector<Ciphertext>& getResultList();

    void setIndex(std::vector<std::vector<Ciphertext>>&& indexMatrix) {
        this->indexMatrix = indexMatrix;
        uploaded = false;
    }

    void setMinusCompareElement(Ciphertext minusCompareElement) {
        this->minusCompareElement = minusCompareElement;
        uploaded = false;
    }

    // flat-buffer variants used by the C ABI (no per-ciphertext allocation on the way in / out)
    void setQueryFlat(const uint64_t* idx, const uint64_t* minus);
    void getResultFlat(uint64_t* out);

    uint32_t numberOfCuckooHashFunctions() const { return K; }
    uint32_t binSize() const { return b; }
    uint32_t cuckooTableSize() const { return E; }
    uint32_t batchSize() const { return nslots; }
    std::vector<int64_t> keptSlots, keptMaskSlots;  // [K][b][E][nslots], [b][nslots] when keepSlots

   private:
    bool resultsFetched = false;
    void upload();
};

}  // namespace psi

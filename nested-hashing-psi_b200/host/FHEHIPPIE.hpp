// Drop-in mirror of the reference's NON-batched PIE operator and its collection:
//   class FHEHIPPIE            /root/reference/src/Common/Crypto/PrivateIndexedEqualityCheck/FHEHIPPIE.hpp:18-50
//   class FHEHIPPIECollection  /root/reference/src/Common/Crypto/PrivateIndexedEqualityCheck/PIECollection.hpp
// Same entry points (ctor from a CuckooHashTable, setIndex, run, getResultList; addPIE, runAll, myPIEs), same
// std::invalid_argument cases and messages (FHEHIPPIE.cpp:13-20).  Underneath, the PIEs of a collection share ONE
// device database (psi_nb_db_encode_slots) and runAll() evaluates all of them in lock step (psi_nb_run); a PIE
// constructed on its own is a collection of one.  The automorphism keys the reference's context holds after
// DeserializeEvalSumKey / DeserializeEvalAutomorphismKey (SimpleFHEPSIServer.cpp:45-62) are installed on the device
// context with psi_nb_set_automorphism_keys before run().
// Header-only: everything below is host-side bookkeeping over the C ABI.
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <memory>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "BatchedFHEHIPPIE.hpp"  // CryptoContext, PublicKey, Ciphertext

namespace psi {

class FHEHIPPIECollection;

class FHEHIPPIE {
   protected:
    CryptoContext& cryptor;
    PublicKey& pK;
    std::vector<Ciphertext> shuffledResultList;
    std::vector<unsigned> permutationVector;
    std::vector<Ciphertext> indexMatrix;
    unsigned numberOfResultElements;
    // what the reference keeps as vectorizedCT / preCalcRandomMask, before encoding
    uint32_t K, b;
    std::vector<int64_t> slots;      // [K][b][b + 1]
    std::vector<int64_t> maskSlots;  // [K][b]
    FHEHIPPIECollection* owner = nullptr;
    uint32_t number = 0;  // position inside the owner's database
    friend class FHEHIPPIECollection;

    static std::vector<unsigned> createPermutationVector(unsigned n, std::mt19937_64& mt) {
        std::vector<unsigned> v(n);
        std::iota(v.begin(), v.end(), 0u);
        std::shuffle(v.begin(), v.end(), mt);
        return v;
    }

   public:
    // seed: PSI_SEED_RANDOM (default) draws the bin permutation, the masks and the result permutation from
    // std::random_device like the reference (FHEHIPPIE.cpp:31); explicit seeds are for tests only - the masks are
    // what keeps non-matching server items hidden from the client.
    FHEHIPPIE(CryptoContext& cryptor, PublicKey& pK, CuckooHashTable& ct, uint64_t seed = PSI_SEED_RANDOM)
        : cryptor(cryptor), pK(pK), numberOfResultElements(ct.getNumberOfHashFunctions()) {
        if (ct.getBinSize() != ct.getEachTableSize())
            throw std::invalid_argument(
                "Error, for FHE PIE the size of a cuckoo bin has to be equal than the number of bins per hash function.");
        if (ct.stash.size() != 0) throw std::invalid_argument("Error, FHE PIE does not support a stash (yet).");
        std::mt19937_64 mt(seed == PSI_SEED_RANDOM ? ((uint64_t)std::random_device{}() << 32) ^ std::random_device{}() : seed);
        permutationVector = createPermutationVector(numberOfResultElements, mt);
        shuffledResultList.assign(numberOfResultElements, nullptr);
        K = ct.getNumberOfHashFunctions();
        b = (uint32_t)ct.getBinSize();
        const uint32_t E = (uint32_t)ct.getEachTableSize(), ns = E + 1;
        const auto permVec2 = createPermutationVector(b, mt);  // hides the correct bin index (:29)
        const uint64_t t = cryptor.GetPlaintextModulus();
        slots.assign((size_t)K * b * ns, 0);
        maskSlots.assign((size_t)K * b, 0);
        for (uint32_t hf = 0; hf < K; hf++)
            for (uint32_t bin = 0; bin < b; bin++) {
                int64_t* plainVec = &slots[((size_t)hf * b + permVec2[bin]) * ns];
                for (uint32_t pos = 0; pos < E; pos++) plainVec[pos] = (int64_t)ct.cell(hf, bin, pos);
                plainVec[E] = 1;                                               // exponent of the "minus client" element (:49)
                maskSlots[(size_t)hf * b + bin] = (int64_t)(mt() % (t - 1) + 1);  // without 0 (:54)
            }
    }

    inline void run();

    std::vector<Ciphertext>& getResultList() { return shuffledResultList; }

    void setIndex(std::vector<Ciphertext>&& indexMatrix) { this->indexMatrix = indexMatrix; }
};

class FHEHIPPIECollection {
   protected:
    CryptoContext& cryptor;
    PublicKey& pK;
    bool encoded = false;
    uint64_t seed;

   public:
    std::deque<FHEHIPPIE> myPIEs;  // stable addresses: PIEs keep a back pointer

    FHEHIPPIECollection(CryptoContext& cryptor, PublicKey& pK, uint64_t seed = PSI_SEED_RANDOM) : cryptor(cryptor), pK(pK), seed(seed) {}

    void addPIE(CuckooHashTable& ct) {
        myPIEs.emplace_back(cryptor, pK, ct, seed == PSI_SEED_RANDOM ? seed : seed + myPIEs.size());
        FHEHIPPIE& p = myPIEs.back();
        if (myPIEs.size() > 1 && (p.K != myPIEs.front().K || p.b != myPIEs.front().b))
            throw std::invalid_argument("Error, all PIEs of a collection need the same table shape.");
        p.owner = this;
        p.number = (uint32_t)myPIEs.size() - 1;
        encoded = false;
    }

    void runAll() { runRange(0, (uint32_t)myPIEs.size()); }

    void runRange(uint32_t begin, uint32_t end) {
        if (begin >= end || end > myPIEs.size()) return;
        const uint32_t K = myPIEs.front().K, b = myPIEs.front().b;
        const size_t ctWords = (size_t)2 * cryptor.params.L * cryptor.params.N;
        if (!encoded) {  // MakePackedPlaintext of every plainVec and mask on the device (FHEHIPPIE.cpp:52,57)
            std::vector<int64_t> slots, masks;
            for (const FHEHIPPIE& p : myPIEs) {
                slots.insert(slots.end(), p.slots.begin(), p.slots.end());
                masks.insert(masks.end(), p.maskSlots.begin(), p.maskSlots.end());
            }
            check(psi_nb_db_encode_slots(cryptor.device_ctx, (uint32_t)myPIEs.size(), K, b, b + 1, slots.data(), masks.data()));
            encoded = true;
        }
        std::vector<uint64_t> idx((size_t)(end - begin) * K * ctWords), out(idx.size());
        for (uint32_t p = begin; p < end; p++) {
            const FHEHIPPIE& pie = myPIEs[p];
            if (pie.indexMatrix.size() != K) throw std::invalid_argument("Error, setIndex needs one ciphertext per hash function.");
            for (uint32_t hf = 0; hf < K; hf++) {
                if (!pie.indexMatrix[hf] || pie.indexMatrix[hf]->size() != ctWords)
                    throw std::invalid_argument("Error, an index ciphertext must hold [2][L][N] limbs.");
                std::copy(pie.indexMatrix[hf]->begin(), pie.indexMatrix[hf]->end(), idx.begin() + ((size_t)(p - begin) * K + hf) * ctWords);
            }
        }
        check(psi_nb_run(cryptor.device_ctx, begin, end, idx.data(), out.data(), nullptr));
        for (uint32_t p = begin; p < end; p++) {
            FHEHIPPIE& pie = myPIEs[p];
            for (uint32_t hf = 0; hf < K; hf++) {  // shuffledResultList[permutationVector[hfInd]] = result (:74)
                const uint64_t* src = out.data() + ((size_t)(p - begin) * K + hf) * ctWords;
                pie.shuffledResultList[pie.permutationVector[hf]] = std::make_shared<std::vector<uint64_t>>(src, src + ctWords);
            }
        }
    }

   private:
    static void check(int rc) {
        if (rc == PSI_ERR_INVALID) throw std::invalid_argument(psi_last_error());
        if (rc != PSI_OK) throw std::runtime_error(psi_last_error());  // OpenFHE exceptions propagate in the reference
    }
};

inline void FHEHIPPIE::run() {
    if (owner) {
        owner->runRange(number, number + 1);
        return;
    }
    // a PIE constructed on its own: a database of one
    const size_t ctWords = (size_t)2 * cryptor.params.L * cryptor.params.N;
    auto fail = [](int rc) {
        if (rc == PSI_ERR_INVALID) throw std::invalid_argument(psi_last_error());
        if (rc != PSI_OK) throw std::runtime_error(psi_last_error());
    };
    fail(psi_nb_db_encode_slots(cryptor.device_ctx, 1, K, b, b + 1, slots.data(), maskSlots.data()));
    if (indexMatrix.size() != K) throw std::invalid_argument("Error, setIndex needs one ciphertext per hash function.");
    std::vector<uint64_t> idx((size_t)K * ctWords), out(idx.size());
    for (uint32_t hf = 0; hf < K; hf++) {
        if (!indexMatrix[hf] || indexMatrix[hf]->size() != ctWords)
            throw std::invalid_argument("Error, an index ciphertext must hold [2][L][N] limbs.");
        std::copy(indexMatrix[hf]->begin(), indexMatrix[hf]->end(), idx.begin() + (size_t)hf * ctWords);
    }
    fail(psi_nb_run(cryptor.device_ctx, 0, 1, idx.data(), out.data(), nullptr));
    for (uint32_t hf = 0; hf < K; hf++)
        shuffledResultList[permutationVector[hf]] =
            std::make_shared<std::vector<uint64_t>>(out.begin() + (size_t)hf * ctWords, out.begin() + (size_t)(hf + 1) * ctWords);
}

}  // namespace psi

// C++ known-answer test of the NON-batched drop-in operator, shaped like the reference's own
//   /root/reference/tests/TestFHEPIE.cpp
// (same constants: t = 2^32 + 2^20 + 2^19 + 1, depth 3, batch size 100, 15000 random elements mod t, client element =
// #7500, 3 cuckoo hash functions, table size 100, bin size 100, the default TabulationHashing, EvalSum keys and rotation
// keys -1 .. -100) but against psi::FHEHIPPIE (libpsi_b200.so, B200 kernels) instead of OpenFHE.
// Client-side cryptography (KeyGen, EvalSumKeyGen / EvalRotateKeyGen, Encrypt, Decrypt) is not part of the product: it
// comes from the ORACLE (oracle/psi_oracle.c), which only tests may link.
// Expected output: exactly one line "Matches" (the element sits in one bin of one hash function's table).
// With the argument "check" the result limbs are also compared with the oracle's FHEHIPPIE::run restatement.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "FHEHIPPIE.hpp"

extern "C" {
struct orc_ctx;
orc_ctx* orc_create(const psi_params* p);
void orc_destroy(orc_ctx* c);
void orc_keygen(const orc_ctx* c, uint64_t seed, uint64_t* sk, uint64_t* evk_b, uint64_t* evk_a);
int orc_encrypt_sk(const orc_ctx* c, const uint64_t* sk, const int64_t* slots, int nslots, uint64_t seed, uint64_t* ct);
int orc_decrypt(const orc_ctx* c, const uint64_t* sk, const uint64_t* ct, int ncomp, int64_t* slots_out, int* ambiguous,
                double* noise_budget_bits);
int orc_eval_sum_indices(const orc_ctx* c, int batch_size, uint64_t* out);
uint64_t orc_find_automorphism_index(const orc_ctx* c, int64_t i);
void orc_auto_keygen(const orc_ctx* c, const uint64_t* sk, uint64_t seed, uint64_t g, uint64_t* key_b, uint64_t* key_a);
int orc_nb_run(const orc_ctx* c, int K, int b, const uint64_t* idx, const uint64_t* pt, const uint64_t* merge_pt,
               const uint64_t* mask, int n_keys, const uint64_t* key_index, const uint64_t* key_b, const uint64_t* key_a,
               uint64_t* out);
}

using namespace psi;

static void ck(int rc, const char* what) {
    if (rc != PSI_OK) {
        std::fprintf(stderr, "%s: %s\n", what, psi_last_error());
        std::exit(2);
    }
}

int main(int argc, char** argv) {
    const bool check = argc > 1 && std::strcmp(argv[1], "check") == 0;
    // Step 1 - crypto context (TestFHEPIE.cpp:13-31)
    const uint64_t n = (1ULL << 32) + (1ULL << 20) + (1ULL << 19) + 1;
    psi_params params;
    ck(psi_params_generate(16384, n, /*depth*/ 3, 0, &params), "GenCryptoContext");
    psi_ctx* dev = nullptr;
    ck(psi_ctx_create(&params, 0, &dev), "psi_ctx_create");
    CryptoContext cryptoContext{params, dev, nullptr};
    PublicKey publicKey;
    const size_t limb = params.N, poly = (size_t)params.L * limb, ctWords = 2 * poly, keyWords = (size_t)params.L * poly;

    // KeyGen, EvalSumKeyGen, EvalRotateKeyGen(-1 .. -100) (:37-50)
    orc_ctx* client = orc_create(&params);
    std::vector<uint64_t> sk(poly), evk_b(keyWords), evk_a(keyWords);
    orc_keygen(client, 2025, sk.data(), evk_b.data(), evk_a.data());
    std::vector<uint64_t> keyIndex(32);
    keyIndex.resize(orc_eval_sum_indices(client, 100, keyIndex.data()));  // parameters.SetBatchSize(100)
    for (int i = 0; i < 100; i++) {
        const uint64_t g = orc_find_automorphism_index(client, -(int64_t)(i + 1));
        if (std::find(keyIndex.begin(), keyIndex.end(), g) == keyIndex.end()) keyIndex.push_back(g);
    }
    std::vector<uint64_t> key_b(keyIndex.size() * keyWords), key_a(key_b.size());
#pragma omp parallel for
    for (size_t i = 0; i < keyIndex.size(); i++)
        orc_auto_keygen(client, sk.data(), 31337, keyIndex[i], key_b.data() + i * keyWords, key_a.data() + i * keyWords);
    ck(psi_nb_set_automorphism_keys(dev, (uint32_t)keyIndex.size(), keyIndex.data(), key_b.data(), key_a.data()),
       "DeserializeEvalAutomorphismKey");

    // 15000 random non-zero elements mod n (:52-68)
    std::mt19937 mt((uint32_t)122333444455555ULL);
    const int numberOfElem = 15000;
    std::vector<item_t> elemForCuckoo(numberOfElem);
    for (auto& e : elemForCuckoo) {
        item_t r = 0;
        while (r == 0) r = boost_uniform_u64(mt) % n;
        e = r;
    }
    const item_t clientElem = elemForCuckoo[numberOfElem / 2];
    const int64_t elem = (int64_t)clientElem;
    std::cout << "Element to compare: \t" << elem << std::endl;

    const unsigned numberOfCuckooHashFunctions = 3;
    const uint64_t cuckooHashTableSize = 100, eachBinSize = 100, stashSize = 0;
    TabulationHashing hashfu;
    CuckooHashTable cT(hashfu, cuckooHashTableSize, numberOfCuckooHashFunctions, 0, stashSize, true, eachBinSize);
    cT.insertAll(elemForCuckoo);

    // one index ciphertext per hash function: one-hot position + the minus element in slot 100 (:95-115)
    std::vector<Ciphertext> indexMatrix(numberOfCuckooHashFunctions);
    for (unsigned hfInd = 0; hfInd < numberOfCuckooHashFunctions; hfInd++) {
        std::vector<int64_t> plainIndexVec(cuckooHashTableSize + 1, 0);
        const uint64_t hashIndex = calculateHashIndex(hashfu, clientElem, hfInd, (uint32_t)cuckooHashTableSize);
        std::cout << "Hash index " << hfInd << ": " << hashIndex << std::endl;
        plainIndexVec[hashIndex] = 1;
        plainIndexVec[cuckooHashTableSize] = -elem;
        auto ct = std::make_shared<std::vector<uint64_t>>(ctWords);
        orc_encrypt_sk(client, sk.data(), plainIndexVec.data(), (int)plainIndexVec.size(), 500 + hfInd, ct->data());
        indexMatrix[hfInd] = ct;
    }
    std::vector<Ciphertext> indexCopy = indexMatrix;

    FHEHIPPIE pie(cryptoContext, publicKey, cT);
    pie.setIndex(std::move(indexMatrix));
    pie.run();

    int matches = 0, bad = 0;
    std::vector<int64_t> slots(params.N);
    for (auto& encryptedResult : pie.getResultList()) {
        int amb = 0;
        double budget = 0;
        orc_decrypt(client, sk.data(), encryptedResult->data(), 2, slots.data(), &amb, &budget);
        if (amb || budget < 5) bad++;
        for (uint64_t s = 0; s < eachBinSize; s++)  // plaintext->SetLength(eachBinSize)
            if (slots[s] == 0) {
                std::cout << "Matches" << std::endl;
                matches++;
                break;
            }
    }
    std::cout << "noise budget ok: " << (bad == 0 ? "yes" : "no") << std::endl;

    int differing = 0;
    if (check) {  // limb parity with the oracle on the database the device encoded
        const int K = (int)numberOfCuckooHashFunctions, b = (int)eachBinSize;
        std::vector<uint64_t> pt((size_t)K * b * poly), mask((size_t)K * poly), merge(poly), idx((size_t)K * ctWords),
            want((size_t)K * ctWords);
        ck(psi_nb_db_get_limbs(dev, pt.data(), mask.data(), merge.data()), "psi_nb_db_get_limbs");
        for (int hf = 0; hf < K; hf++) std::copy(indexCopy[hf]->begin(), indexCopy[hf]->end(), idx.begin() + (size_t)hf * ctWords);
        if (orc_nb_run(client, K, b, idx.data(), pt.data(), merge.data(), mask.data(), (int)keyIndex.size(), keyIndex.data(),
                       key_b.data(), key_a.data(), want.data()))
            differing = -1;
        else {
            // the PIE permutes its results (permutationVector): match every oracle result with one of the K outputs
            for (int hf = 0; hf < K; hf++) {
                bool found = false;
                for (auto& r : pie.getResultList())
                    found = found || std::memcmp(r->data(), want.data() + (size_t)hf * ctWords, ctWords * sizeof(uint64_t)) == 0;
                if (!found) differing++;
            }
        }
        std::cout << "limb parity with the oracle: " << (differing == 0 ? "identical" : "DIFFERENT") << std::endl;
    }
    orc_destroy(client);
    psi_ctx_destroy(dev);
    return (matches == 1 && bad == 0 && differing == 0) ? 0 : 1;
}

// C++ known-answer test of the drop-in operator, shaped like the reference's own
//   /root/reference/tests/TestBatchedFHEPIE.cpp
// (same constants: t = 2^32 + 2^20 + 2^19 + 1, depth 2, 100 random elements mod t, client element = #50,
// k = K = 2, e = 1, E = 10, b = 20, hash seed 12223222 with 4 hash functions, both slots filled with the
// element) but against psi::BatchedFHEHIPPIE (libpsi_b200.so, B200 kernels) instead of OpenFHE.
// Client-side cryptography (KeyGen, Encrypt with the secret key, Decrypt) is not part of the product: it
// comes from the ORACLE (oracle/psi_oracle.c), which only tests may link.
// Expected output: "Test should output matches twice" followed by exactly two lines "Matches".
// With an argument the evaluator is the single-process multi-device one (psi_multi_*, the server stays one
// process with one PIE object):  TestBatchedFHEPIE 0,1  runs on devices 0 and 1 (a device may repeat).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "BatchedFHEHIPPIE.hpp"

extern "C" {
struct orc_ctx;
orc_ctx* orc_create(const psi_params* p);
void orc_destroy(orc_ctx* c);
void orc_keygen(const orc_ctx* c, uint64_t seed, uint64_t* sk, uint64_t* evk_b, uint64_t* evk_a);
int orc_encrypt_sk(const orc_ctx* c, const uint64_t* sk, const int64_t* slots, int nslots, uint64_t seed, uint64_t* ct);
int orc_decrypt(const orc_ctx* c, const uint64_t* sk, const uint64_t* ct, int ncomp, int64_t* slots_out, int* ambiguous,
                double* noise_budget_bits);
}

using namespace psi;

static void ck(int rc, const char* what) {
    if (rc != PSI_OK) {
        std::fprintf(stderr, "%s: %s\n", what, psi_last_error());
        std::exit(2);
    }
}

int main(int argc, char** argv) {
    std::vector<int> devices;  // empty: one psi_ctx on device 0
    if (argc > 1) {
        std::stringstream ss(argv[1]);
        for (std::string tok; std::getline(ss, tok, ',');) devices.push_back(std::atoi(tok.c_str()));
    }
    // Step 1 - crypto context (TestBatchedFHEPIE.cpp:14-31): ring dimension and sizeQ as the library picks
    const uint64_t n = (1ULL << 32) + (1ULL << 20) + (1ULL << 19) + 1;
    psi_params params;
    ck(psi_params_generate(8192, n, /*depth*/ 2, 0, &params), "GenCryptoContext");
    psi_ctx* dev = nullptr;
    psi_multi* multi = nullptr;
    if (devices.empty())
        ck(psi_ctx_create(&params, 0, &dev), "psi_ctx_create");
    else
        ck(psi_multi_create(&params, devices.data(), (uint32_t)devices.size(), &multi), "psi_multi_create");
    CryptoContext cryptoContext{params, dev, multi};
    PublicKey publicKey;
    const size_t ctWords = (size_t)2 * params.L * params.N;

    // KeyGen + EvalMultKeysGen (:39-41)
    orc_ctx* client = orc_create(&params);
    std::vector<uint64_t> sk((size_t)params.L * params.N), evk_b((size_t)params.L * params.L * params.N), evk_a(evk_b.size());
    orc_keygen(client, 2024, sk.data(), evk_b.data(), evk_a.data());
    ck(multi ? psi_multi_set_relin_key(multi, evk_b.data(), evk_a.data()) : psi_set_relin_key(dev, evk_b.data(), evk_a.data()),
       "InsertEvalMultKey");

    // 100 random non-zero elements mod n (:54-70)
    std::mt19937 mt((uint32_t)122333444455555ULL);
    const int numberOfElem = 100;
    std::vector<item_t> elemForCuckoo(numberOfElem);
    for (auto& e : elemForCuckoo) {
        item_t r = 0;
        while (r == 0) r = boost_uniform_u64(mt) % n;
        e = r;
    }
    std::cout << "Test should output matches twice" << std::endl;
    const item_t clientElem = elemForCuckoo[numberOfElem / 2];
    const int64_t elem = (int64_t)clientElem;
    std::cout << "Element to compare: \t" << elem << std::endl;

    const unsigned numberOfSimpleHashFunctions = 2, numberOfCuckooHashFunctions = 2;
    const uint64_t eachSimpleTableSize = 1, cuckooHashTableSize = 10, eachBinSize = 20, stashSize = 0;
    TabulationHashing hashfu(12223222, numberOfSimpleHashFunctions + numberOfCuckooHashFunctions);
    HierarchicalCuckooHashTable hcT(hashfu, eachSimpleTableSize, cuckooHashTableSize, stashSize, numberOfSimpleHashFunctions,
                                    numberOfCuckooHashFunctions, true, true, eachBinSize);
    hcT.insertAll(elemForCuckoo);

    // encrypted one-hot index matrix (:101-124)
    std::vector<std::vector<Ciphertext>> indexMatrix(numberOfCuckooHashFunctions, std::vector<Ciphertext>(cuckooHashTableSize));
    uint64_t encSeed = 1000;
    for (unsigned hfInd = numberOfSimpleHashFunctions; hfInd < numberOfSimpleHashFunctions + numberOfCuckooHashFunctions; hfInd++) {
        const uint64_t hashIndex = calculateHashIndex(hashfu, clientElem, hfInd, (uint32_t)cuckooHashTableSize);
        std::cout << "Hash index " << hfInd << ": " << hashIndex << std::endl;
        for (uint64_t vectorIndex = 0; vectorIndex < cuckooHashTableSize; vectorIndex++) {
            std::vector<int64_t> plainIndexVec(2, vectorIndex == hashIndex ? 1 : 0);
            auto ct = std::make_shared<std::vector<uint64_t>>(ctWords);
            orc_encrypt_sk(client, sk.data(), plainIndexVec.data(), 2, encSeed++, ct->data());
            indexMatrix[hfInd - numberOfSimpleHashFunctions][vectorIndex] = ct;
        }
    }

    BatchedFHEHIPPIE pie(cryptoContext, publicKey, hcT);
    pie.setIndex(std::move(indexMatrix));

    // compare element (:131-137)
    std::vector<int64_t> plainMinusEl(2, -elem);
    auto minusComp = std::make_shared<std::vector<uint64_t>>(ctWords);
    orc_encrypt_sk(client, sk.data(), plainMinusEl.data(), 2, 999, minusComp->data());
    pie.setMinusCompareElement(minusComp);
    pie.run();

    int matches = 0;
    std::vector<int64_t> slots(params.N);
    for (auto& encryptedResult : pie.getResultList()) {
        int amb = 0;
        double budget = 0;
        orc_decrypt(client, sk.data(), encryptedResult->data(), 2, slots.data(), &amb, &budget);
        for (int s = 0; s < 2; s++)  // plaintext->SetLength(2)
            if (slots[s] == 0) {
                std::cout << "Matches" << std::endl;
                matches++;
            }
    }
    orc_destroy(client);
    psi_ctx_destroy(dev);
    psi_multi_destroy(multi);
    return matches == 2 ? 0 : 1;
}

// Runs adapter/BatchedFHEHIPPIE_b200.cpp — the reference's class (unmodified header) over the GPU library — on the
// scenario of the reference's own tests/TestBatchedFHEPIE.cpp: t = 2^32 + 2^20 + 2^19 + 1, depth 2, 100 random
// elements, client element #50, k = K = 2, e = 1, E = 10, b = 20.  OpenFHE is absent here: the lbcrypto objects are
// the shim's (adapter/shim/openfhe.h), filled with limbs produced by the ORACLE acting as the client (key
// generation, secret-key encryption, decryption).  Expected: "Matches" exactly twice, exit code 0.
// Usage: test_adapter [derived]   (derived: the double tables are computed from the moduli instead of being copied
//                                  from the context, PSI_B200_TABLES=derived)
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>

#include "BatchedFHEHIPPIE.hpp"  // the reference's header
#include "psi_b200.h"

extern "C" {
struct orc_ctx;
orc_ctx* orc_create(const psi_params* p);
void orc_destroy(orc_ctx* c);
void orc_keygen(const orc_ctx* c, uint64_t seed, uint64_t* sk, uint64_t* evk_b, uint64_t* evk_a);
int orc_encrypt_sk(const orc_ctx* c, const uint64_t* sk, const int64_t* slots, int nslots, uint64_t seed, uint64_t* ct);
int orc_decrypt(const orc_ctx* c, const uint64_t* sk, const uint64_t* ct, int ncomp, int64_t* slots_out, int* ambiguous,
                double* noise_budget_bits);
}

using namespace lbcrypto;

static std::shared_ptr<DCRTPoly::Params> g_paramsQ;

static DCRTPoly poly_from(const uint64_t* limbs, size_t L, size_t N) {
    DCRTPoly poly(g_paramsQ, Format::EVALUATION, true);
    for (size_t l = 0; l < L; l++) {
        NativeVector v((usint)N, g_paramsQ->GetParams()[l]->GetModulus());
        for (size_t n = 0; n < N; n++) v[n] = NativeInteger(limbs[l * N + n]);
        NativePoly limb = poly.GetElementAtIndex((usint)l);
        limb.SetValues(std::move(v), Format::EVALUATION);
        poly.SetElementAtIndex((usint)l, std::move(limb));
    }
    return poly;
}

static Ciphertext<DCRTPoly> ct_from(const CryptoContext<DCRTPoly>& cc, const uint64_t* limbs, size_t L, size_t N) {
    auto ct = std::make_shared<CiphertextImpl<DCRTPoly>>(cc, "client-key");
    std::vector<DCRTPoly> cv;
    cv.push_back(poly_from(limbs, L, N));
    cv.push_back(poly_from(limbs + L * N, L, N));
    ct->SetElements(std::move(cv));
    return ct;
}

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "derived") setenv("PSI_B200_TABLES", "derived", 1);
    // Step 1 - crypto context (TestBatchedFHEPIE.cpp:14-31)
    const uint64_t n = (1ULL << 32) + (1ULL << 20) + (1ULL << 19) + 1;
    psi_params P;
    if (psi_params_generate(8192, n, 2, 0, &P) != PSI_OK) return 3;
    const size_t L = P.L, N = P.N, ctWords = 2 * L * N;
    auto cp = std::make_shared<CryptoParametersBFVRNS>();
    std::vector<std::shared_ptr<ILNativeParams>> tq, tp;
    for (size_t i = 0; i < L; i++) tq.push_back(std::make_shared<ILNativeParams>(2 * N, P.q[i], P.psi_q[i]));
    for (size_t j = 0; j < P.Lp; j++) tp.push_back(std::make_shared<ILNativeParams>(2 * N, P.p[j], P.psi_p[j]));
    cp->elementParams = g_paramsQ = std::make_shared<DCRTPoly::Params>(2 * N, tq);
    cp->paramsRl = std::make_shared<DCRTPoly::Params>(2 * N, tp);
    cp->encodingParams = std::make_shared<EncodingParamsImpl>(P.t, P.psi_t);
    cp->qInv.assign(P.qInv, P.qInv + L);
    cp->rInv.assign(P.pInv, P.pInv + P.Lp);
    cp->tQlSlHatInvModsDivsFrac.push_back(std::vector<double>(P.tQSHatInvModsDivsFrac, P.tQSHatInvModsDivsFrac + P.Lp));
    CryptoContext<DCRTPoly> cryptoContext = std::make_shared<CryptoContextImpl<DCRTPoly>>(cp);

    // KeyGen + EvalMultKeysGen (:39-41): the oracle is the client
    orc_ctx* client = orc_create(&P);
    std::vector<uint64_t> sk(L * N), evk_b(L * L * N), evk_a(L * L * N);
    orc_keygen(client, 2024, sk.data(), evk_b.data(), evk_a.data());
    PublicKey<DCRTPoly> publicKey = std::make_shared<PublicKeyImpl<DCRTPoly>>("client-key");
    {
        auto key = std::make_shared<EvalKeyRelinImpl<DCRTPoly>>();
        std::vector<DCRTPoly> av, bv;
        for (size_t i = 0; i < L; i++) {
            bv.push_back(poly_from(&evk_b[i * L * N], L, N));
            av.push_back(poly_from(&evk_a[i * L * N], L, N));
        }
        key->SetAVector(std::move(av));
        key->SetBVector(std::move(bv));
        CryptoContextImpl<DCRTPoly>::InsertEvalMultKey({key}, "client-key");
    }

    // 100 random non-zero elements mod n (:54-70)
    std::mt19937 mt((uint32_t)122333444455555ULL);
    const int numberOfElem = 100;
    vector<biginteger> elemForCuckoo(numberOfElem);
    for (auto& e : elemForCuckoo) {
        biginteger r = 0;
        while (r == 0) r = psi::boost_uniform_u64(mt) % n;
        e = r;
    }
    std::cout << "Test should output matches twice" << std::endl;
    const biginteger clientElem = elemForCuckoo[numberOfElem / 2];
    const int64_t elem = (int64_t)clientElem;

    const uint numberOfSimpleHashFunctions = 2, numberOfCuckooHashFunctions = 2;
    const uint64_t eachSimpleTableSize = 1, cuckooHashTableSize = 10, eachBinSize = 20, stashSize = 0;
    TabulationHashing hashfu(12223222, numberOfSimpleHashFunctions + numberOfCuckooHashFunctions);
    HierarchicalCuckooHashTable hcT(hashfu, eachSimpleTableSize, cuckooHashTableSize, stashSize, numberOfSimpleHashFunctions,
                                    numberOfCuckooHashFunctions, true, true, eachBinSize);
    hcT.insertAll(elemForCuckoo);

    // encrypted one-hot index matrix (:101-124)
    vector<vector<Ciphertext<DCRTPoly>>> indexMatrix(numberOfCuckooHashFunctions, vector<Ciphertext<DCRTPoly>>(cuckooHashTableSize));
    std::vector<uint64_t> buf(ctWords);
    uint64_t encSeed = 1000;
    for (uint hfInd = numberOfSimpleHashFunctions; hfInd < numberOfSimpleHashFunctions + numberOfCuckooHashFunctions; hfInd++) {
        const uint64_t hashIndex = psi::calculateHashIndex(hashfu, clientElem, hfInd, (uint32_t)cuckooHashTableSize);
        for (uint64_t vectorIndex = 0; vectorIndex < cuckooHashTableSize; vectorIndex++) {
            std::vector<int64_t> plainIndexVec(2, vectorIndex == hashIndex ? 1 : 0);
            orc_encrypt_sk(client, sk.data(), plainIndexVec.data(), 2, encSeed++, buf.data());
            indexMatrix[hfInd - numberOfSimpleHashFunctions][vectorIndex] = ct_from(cryptoContext, buf.data(), L, N);
        }
    }

    int rc = 0;
    try {
        // error behaviour of the reference's constructor (BatchedFHEHIPPIE.cpp:13-21)
        bool threw = false;
        try {
            HierarchicalCuckooHashTable withStash(hashfu, 1, 10, /*stash*/ 3, 2, 2, true, true, 20);
            BatchedFHEHIPPIE bad(cryptoContext, publicKey, withStash);
        } catch (const std::invalid_argument& e) {
            threw = std::string(e.what()) == "Error, batched FHE PIE does not support a stash (yet).";
        }
        if (!threw) {
            std::cerr << "stash case did not throw the reference's invalid_argument" << std::endl;
            rc = 4;
        }

        BatchedFHEHIPPIE pie(cryptoContext, publicKey, hcT);
        pie.setIndex(std::move(indexMatrix));
        std::vector<int64_t> plainMinusEl(2, -elem);
        orc_encrypt_sk(client, sk.data(), plainMinusEl.data(), 2, 999, buf.data());
        pie.setMinusCompareElement(ct_from(cryptoContext, buf.data(), L, N));
        pie.run();

        int matches = 0, nonzero = 0;
        std::vector<int64_t> slots(N);
        auto& results = pie.getResultList();
        if (results.size() != eachBinSize) rc = 5;
        for (auto& encryptedResult : results) {
            // back to flat limbs for the oracle's decryption
            for (size_t c = 0; c < 2; c++)
                for (size_t l = 0; l < L; l++) {
                    const NativeVector& v = encryptedResult->GetElements()[c].GetElementAtIndex((usint)l).GetValues();
                    for (size_t i = 0; i < N; i++) buf[(c * L + l) * N + i] = v[i].ConvertToInt();
                }
            int amb = 0;
            double budget = 0;
            orc_decrypt(client, sk.data(), buf.data(), 2, slots.data(), &amb, &budget);
            for (int s = 0; s < 2; s++) {  // plaintext->SetLength(2)
                if (slots[s] == 0) {
                    std::cout << "Matches" << std::endl;
                    matches++;
                } else {
                    nonzero++;
                }
            }
        }
        if (matches != 2 || nonzero != 2 * (int)eachBinSize - 2) rc = rc ? rc : 1;
    } catch (const std::exception& e) {
        std::cerr << "exception: " << e.what() << std::endl;
        rc = 2;
    }
    orc_destroy(client);
    return rc;
}

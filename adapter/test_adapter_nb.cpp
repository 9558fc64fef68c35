// Runs adapter/FHEHIPPIE_b200.cpp - the reference's non-batched class (unmodified header) over the GPU library - on a
// scaled-down form of the reference's own tests/TestFHEPIE.cpp (same t, depth 3, 3 cuckoo hash functions, square inner
// table, EvalSum keys + rotation keys -1 .. -E; 10 x 10 cells and 150 elements instead of 100 x 100 and 15000 so that
// the shim's key objects stay small; the full-size scenario runs through tests/cpp/TestFHEPIE.cpp).  TWO PIEs share the
// context, as the reference's server builds them (SimpleFHEPSIServer.cpp:99-121): the first holds the client's
// element ("Matches" exactly once), the second does not (no match).  OpenFHE is absent: the lbcrypto objects are the
// shim's, filled with limbs produced by the ORACLE acting as the client.  Result limbs are also compared with the
// oracle's FHEHIPPIE::run restatement on the database the device encoded.  Exit code 0 on success.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>

#include "FHEHIPPIE.hpp"  // the reference's header
#include "psi_b200.h"

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
psi_ctx* psi_b200_nb_adapter_ctx(const void* cryptoContextImpl);
void psi_b200_nb_adapter_release(const void* cryptoContextImpl);
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
static void flat_of(const Ciphertext<DCRTPoly>& ct, size_t L, size_t N, uint64_t* out) {
    for (size_t c = 0; c < 2; c++)
        for (size_t l = 0; l < L; l++) {
            const NativeVector& v = ct->GetElements()[c].GetElementAtIndex((usint)l).GetValues();
            for (size_t i = 0; i < N; i++) out[(c * L + l) * N + i] = v[i].ConvertToInt();
        }
}

int main() {
    const uint64_t n = (1ULL << 32) + (1ULL << 20) + (1ULL << 19) + 1;
    psi_params P;
    if (psi_params_generate(16384, n, 3, 0, &P) != PSI_OK) return 3;
    const size_t L = P.L, N = P.N, poly = L * N, ctWords = 2 * poly, keyWords = L * poly;
    auto cp = std::make_shared<CryptoParametersBFVRNS>();
    std::vector<std::shared_ptr<ILNativeParams>> tq, tp;
    for (size_t i = 0; i < L; i++) tq.push_back(std::make_shared<ILNativeParams>(2 * N, P.q[i], P.psi_q[i]));
    for (size_t j = 0; j < P.Lp; j++) tp.push_back(std::make_shared<ILNativeParams>(2 * N, P.p[j], P.psi_p[j]));
    cp->elementParams = g_paramsQ = std::make_shared<DCRTPoly::Params>(2 * N, tq);
    cp->paramsRl = std::make_shared<DCRTPoly::Params>(2 * N, tp);
    cp->encodingParams = std::make_shared<EncodingParamsImpl>(P.t, P.psi_t);
    CryptoContext<DCRTPoly> cryptoContext = std::make_shared<CryptoContextImpl<DCRTPoly>>(cp);
    PublicKey<DCRTPoly> publicKey = std::make_shared<PublicKeyImpl<DCRTPoly>>("client-key");

    const uint numberOfCuckooHashFunctions = 3;
    const uint64_t cuckooHashTableSize = 10, eachBinSize = 10, stashSize = 0;
    const int K = (int)numberOfCuckooHashFunctions, b = (int)eachBinSize;

    // KeyGen, EvalSumKeyGen (batch size E + 1), EvalRotateKeyGen(-1 .. -E) (SimpleFHEPSIClient.cpp:79-90): the oracle is the client
    orc_ctx* client = orc_create(&P);
    std::vector<uint64_t> sk(poly), evk_b(keyWords), evk_a(keyWords);
    orc_keygen(client, 2026, sk.data(), evk_b.data(), evk_a.data());
    std::vector<uint64_t> keyIndex(32);
    keyIndex.resize(orc_eval_sum_indices(client, (int)cuckooHashTableSize + 1, keyIndex.data()));
    for (uint64_t i = 0; i < cuckooHashTableSize; i++) {
        const uint64_t g = orc_find_automorphism_index(client, -(int64_t)(i + 1));
        if (std::find(keyIndex.begin(), keyIndex.end(), g) == keyIndex.end()) keyIndex.push_back(g);
    }
    std::vector<uint64_t> key_b(keyIndex.size() * keyWords), key_a(key_b.size());
    auto keyMap = std::make_shared<std::map<usint, EvalKey<DCRTPoly>>>();
    for (size_t i = 0; i < keyIndex.size(); i++) {
        orc_auto_keygen(client, sk.data(), 4242, keyIndex[i], &key_b[i * keyWords], &key_a[i * keyWords]);
        auto key = std::make_shared<EvalKeyRelinImpl<DCRTPoly>>();
        std::vector<DCRTPoly> av, bv;
        for (size_t d = 0; d < L; d++) {
            bv.push_back(poly_from(&key_b[i * keyWords + d * poly], L, N));
            av.push_back(poly_from(&key_a[i * keyWords + d * poly], L, N));
        }
        key->SetAVector(std::move(av));
        key->SetBVector(std::move(bv));
        (*keyMap)[(usint)keyIndex[i]] = key;
    }
    CryptoContextImpl<DCRTPoly>::InsertEvalAutomorphismKey(keyMap, "client-key");

    // two inner tables of 150 random non-zero elements each (TestFHEPIE.cpp:52-68, scaled down)
    std::mt19937 mt((uint32_t)122333444455555ULL);
    const int numberOfElem = 150;
    vector<biginteger> elemA(numberOfElem), elemB(numberOfElem);
    for (auto* set : {&elemA, &elemB})
        for (auto& e : *set) {
            biginteger r = 0;
            while (r == 0) r = psi::boost_uniform_u64(mt) % n;
            e = r;
        }
    const biginteger clientElem = elemA[numberOfElem / 2];
    const int64_t elem = (int64_t)clientElem;
    TabulationHashing hashfu;
    CuckooHashTable cTA(hashfu, cuckooHashTableSize, numberOfCuckooHashFunctions, 0, stashSize, true, eachBinSize);
    CuckooHashTable cTB(hashfu, cuckooHashTableSize, numberOfCuckooHashFunctions, 0, stashSize, true, eachBinSize);
    cTA.insertAll(elemA);
    cTB.insertAll(elemB);

    int rc = 0;
    try {
        // error behaviour of the reference's constructor (FHEHIPPIE.cpp:13-16)
        bool threw = false;
        try {
            CuckooHashTable notSquare(hashfu, 10, 3, 0, 0, true, 7);
            notSquare.insertAll(elemB);
            FHEHIPPIE bad(cryptoContext, publicKey, notSquare);
        } catch (const std::invalid_argument& e) {
            threw = std::strstr(e.what(), "size of a cuckoo bin has to be equal") != nullptr;
        }
        if (!threw) {
            std::cerr << "non-square table did not throw the reference's invalid_argument" << std::endl;
            rc = 4;
        }

        FHEHIPPIE pieA(cryptoContext, publicKey, cTA), pieB(cryptoContext, publicKey, cTB);
        // the same query goes to both PIEs (one-hot position per hash function + the minus element in slot E, :95-115)
        std::vector<uint64_t> buf(ctWords), idxFlat((size_t)K * ctWords);
        for (FHEHIPPIE* pie : {&pieA, &pieB}) {
            vector<Ciphertext<DCRTPoly>> indexMatrix(numberOfCuckooHashFunctions);
            for (uint hfInd = 0; hfInd < numberOfCuckooHashFunctions; hfInd++) {
                vector<int64_t> plainIndexVec(cuckooHashTableSize + 1, 0);
                plainIndexVec[psi::calculateHashIndex(hashfu, clientElem, hfInd, (uint32_t)cuckooHashTableSize)] = 1;
                plainIndexVec[cuckooHashTableSize] = -elem;
                orc_encrypt_sk(client, sk.data(), plainIndexVec.data(), (int)plainIndexVec.size(), 500 + hfInd, buf.data());
                std::memcpy(&idxFlat[hfInd * ctWords], buf.data(), ctWords * sizeof(uint64_t));
                indexMatrix[hfInd] = ct_from(cryptoContext, buf.data(), L, N);
            }
            pie->setIndex(std::move(indexMatrix));
        }
        pieA.run();
        pieB.run();

        // limb parity with the oracle on the database the device encoded (results are permuted: match each with one output)
        psi_ctx* dev = psi_b200_nb_adapter_ctx(cryptoContext.get());
        std::vector<uint64_t> pt((size_t)2 * K * b * poly), mask((size_t)2 * K * poly), merge(poly), want((size_t)K * ctWords);
        if (!dev || psi_nb_db_get_limbs(dev, pt.data(), mask.data(), merge.data()) != PSI_OK) rc = rc ? rc : 6;
        int pieNumber = 0;
        for (FHEHIPPIE* pie : {&pieA, &pieB}) {
            int matches = 0;
            std::vector<int64_t> slots(N);
            auto& results = pie->getResultList();
            if (results.size() != numberOfCuckooHashFunctions) rc = rc ? rc : 5;
            orc_nb_run(client, K, b, idxFlat.data(), &pt[(size_t)pieNumber * K * b * poly], merge.data(), &mask[(size_t)pieNumber * K * poly],
                       (int)keyIndex.size(), keyIndex.data(), key_b.data(), key_a.data(), want.data());
            std::vector<bool> seen(K, false);
            for (auto& encryptedResult : results) {
                flat_of(encryptedResult, L, N, buf.data());
                for (int hf = 0; hf < K; hf++)
                    if (std::memcmp(buf.data(), &want[(size_t)hf * ctWords], ctWords * sizeof(uint64_t)) == 0) seen[hf] = true;
                int amb = 0;
                double budget = 0;
                orc_decrypt(client, sk.data(), buf.data(), 2, slots.data(), &amb, &budget);
                if (amb || budget < 5) rc = rc ? rc : 7;
                for (uint64_t s = 0; s < eachBinSize; s++)  // plaintext->SetLength(eachBinSize)
                    if (slots[s] == 0) {
                        std::cout << "Matches (PIE " << pieNumber << ")" << std::endl;
                        matches++;
                        break;
                    }
            }
            for (int hf = 0; hf < K; hf++)
                if (!seen[hf]) {
                    std::cerr << "PIE " << pieNumber << ": result of hash function " << hf << " differs from the oracle" << std::endl;
                    rc = rc ? rc : 8;
                }
            if (matches != (pieNumber == 0 ? 1 : 0)) rc = rc ? rc : 1;
            pieNumber++;
        }
        if (rc == 0) std::cout << "limb parity with the oracle: identical" << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "exception: " << e.what() << std::endl;
        rc = 2;
    }
    psi_b200_nb_adapter_release(cryptoContext.get());
    orc_destroy(client);
    return rc;
}

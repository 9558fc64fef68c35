// Drop-in replacement for the reference's
//   src/Common/Crypto/PrivateIndexedEqualityCheck/BatchedFHEHIPPIE.cpp
// It is compiled against the reference's UNMODIFIED header (BatchedFHEHIPPIE.hpp:18-48: same class, same five
// entry points, same members) and keeps the reference's behaviour at the boundary — the std::invalid_argument cases
// (BatchedFHEHIPPIE.cpp:13-21), the in-place bin shuffle of the caller's table (:25-35), the transposition
// (:48-66), random non-zero masks (:73-82), resultList of b ciphertexts — but every homomorphic operation of run()
// (:88-129) happens on the GPUs behind the C ABI of include/psi_b200.h (libpsi_b200.so).
//
// In the reference tree: replace BatchedFHEHIPPIE.cpp by this file in src/CMakeLists.txt and link psi_b200
// (INTEGRATION.md).  In THIS repository OpenFHE, libscapi and Boost are absent, so the file is compiled against
// adapter/shim/ (minimal lbcrypto / hashing declarations, only the members touched here) and run end to end by
// adapter/test_adapter.cpp; lbcrypto member names are those of OpenFHE 1.0.x as recalled and must be checked
// against the installed headers.
//
// The header is not modified, so the per-object GPU state lives in a side table keyed by `this`.
// Devices: environment variable PSI_B200_DEVICES ("0,1,2", default: every visible device, at most one per bin).
#include "BatchedFHEHIPPIE.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <sstream>
#include <unordered_map>

#include "adapter_common.hpp"
#include "psi_b200.h"

using namespace lbcrypto;
using namespace psi_adapter;

namespace {

struct State {
    psi_multi* multi = nullptr;
    size_t K = 0, b = 0, E = 0, batchSize = 0, L = 0, N = 0;
    ~State() { psi_multi_destroy(multi); }
};
std::mutex g_mutex;
std::unordered_map<const BatchedFHEHIPPIE*, std::unique_ptr<State>> g_state;

State& state_of(const BatchedFHEHIPPIE* self) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_state.find(self);
    if (it == g_state.end()) throw std::runtime_error("BatchedFHEHIPPIE (B200): object has no device state");
    return *it->second;
}

std::vector<int> device_list(size_t max_devices) {
    std::vector<int> devs;
    if (const char* env = std::getenv("PSI_B200_DEVICES")) {
        std::stringstream ss(env);
        for (std::string tok; std::getline(ss, tok, ',');)
            if (!tok.empty()) devs.push_back(std::atoi(tok.c_str()));
    } else {
        int n = 0;
        ck(psi_device_count(&n));
        for (int d = 0; d < n; d++) devs.push_back(d);
    }
    if (devs.empty()) throw std::runtime_error("BatchedFHEHIPPIE (B200): no CUDA device; the library has no CPU path");
    if (devs.size() > max_devices) devs.resize(max_devices);  // bins are the shard axis: at most one device per bin
    return devs;
}

}  // namespace

BatchedFHEHIPPIE::BatchedFHEHIPPIE(lbcrypto::CryptoContext<FHEEncType>& cryptoContext, lbcrypto::PublicKey<FHEEncType>& pK,
                                   HierarchicalCuckooHashTable& hct)
    : cryptoContext(cryptoContext), pK(pK) {
    if (hct.getServerStashSize() != 0) {
        throw invalid_argument("Error, batched FHE PIE does not support a stash (yet).");
    }
    if (!hct.hasSimpleMultiTables() || !hct.hasCuckooMultiTables()) {
        throw invalid_argument("Error, batched FHE PIE currently does not support combined tables.");
    }

    // Shuffle bins beforehand; mutates the caller's table exactly like the reference (fresh random_device seed).
    std::random_device rd;
    std::mt19937 mt(rd());
    for (auto& hctRow : hct.hierarchicalCuckooTable)
        for (auto& ct : hctRow)
            for (auto& ctRow : ct.cuckooTable) std::shuffle(std::begin(ctRow), std::end(ctRow), mt);

    std::unique_ptr<State> st(new State());
    st->K = hct.getNumberOfCuckooTables();
    st->b = hct.getEachBinSize();
    st->E = hct.getEachCuckooTableSize();
    const size_t k = hct.getNumberOfSimpleTables(), e = hct.getEachSimpleTableSize();
    st->batchSize = k * e;  // assumes simple multi table!
    const psi_params P = params_from(cryptoContext);
    st->L = P.L;
    st->N = P.N;
    const uint64_t plaintextModulus = cryptoContext->GetCryptoParameters()->GetPlaintextModulus();

    const std::vector<int> devices = device_list(st->b);
    ck(psi_multi_create(&P, devices.data(), (uint32_t)devices.size(), &st->multi));

    // relinearisation key (BV, digit size 0: one (b, a) pair of DCRTPolys per digit), limbs as they are
    {
        const auto& evk = cryptoContext->GetEvalMultKeyVector(pK->GetKeyTag());
        if (evk.empty()) throw std::runtime_error("BatchedFHEHIPPIE (B200): no relinearisation key");
        const auto& bv = evk[0]->GetBVector();
        const auto& av = evk[0]->GetAVector();
        if (bv.size() != st->L || av.size() != st->L) throw std::invalid_argument("BatchedFHEHIPPIE (B200): relinearisation key is not BV / digit size 0");
        std::vector<uint64_t> kb(st->L * st->L * st->N), ka(kb.size());
        for (size_t i = 0; i < st->L; i++)
            for (size_t l = 0; l < st->L; l++) {
                std::memcpy(&kb[(i * st->L + l) * st->N], words_of(bv[i].GetElementAtIndex((usint)l).GetValues()), st->N * sizeof(uint64_t));
                std::memcpy(&ka[(i * st->L + l) * st->N], words_of(av[i].GetElementAtIndex((usint)l).GetValues()), st->N * sizeof(uint64_t));
            }
        ck(psi_multi_set_relin_key(st->multi, kb.data(), ka.data()));
    }

    // transposition hct[outerHf][outerPos].cuckooTable[innerHf][bin][innerPos] -> slot vectors; encoded on the devices
    // (MakePackedPlaintext + SetFormat(EVALUATION)); vectorizedHCT / preCalcRandomMask stay empty: the database lives in HBM
    std::vector<int64_t> slots(st->K * st->b * st->E * st->batchSize);
    for (size_t innerHfInd = 0; innerHfInd < st->K; innerHfInd++)
        for (size_t binIndex = 0; binIndex < st->b; binIndex++)
            for (size_t innerhashPos = 0; innerhashPos < st->E; innerhashPos++) {
                int64_t* plainVec = &slots[((innerHfInd * st->b + binIndex) * st->E + innerhashPos) * st->batchSize];
                size_t batchIndex = 0;
                for (size_t outerHfInd = 0; outerHfInd < k; outerHfInd++)
                    for (size_t outerhashPos = 0; outerhashPos < e; outerhashPos++)
                        plainVec[batchIndex++] = (int64_t)hct.hierarchicalCuckooTable[outerHfInd][outerhashPos].cuckooTable[innerHfInd][binIndex][innerhashPos];
            }
    // random non-zero masks: security-critical, drawn from the same random_device-seeded generator
    std::vector<int64_t> maskSlots(st->b * st->batchSize);
    std::uniform_int_distribution<uint64_t> randGen;
    for (auto& randomMask : maskSlots) randomMask = (int64_t)(randGen(mt) % (plaintextModulus - 1) + 1);  // without 0
    ck(psi_multi_db_encode_slots(st->multi, (uint32_t)st->K, (uint32_t)st->b, (uint32_t)st->E, (uint32_t)st->batchSize, slots.data(),
                                 maskSlots.data()));

    resultList = vector<lbcrypto::Ciphertext<FHEEncType>>(st->b);
    std::lock_guard<std::mutex> lock(g_mutex);
    g_state[this] = std::move(st);
}

void BatchedFHEHIPPIE::run() {
    State& st = state_of(this);
    const size_t L = st.L, N = st.N;
    if (indexMatrix.size() != st.K) throw std::invalid_argument("BatchedFHEHIPPIE (B200): indexMatrix must have one row per cuckoo hash function");
    // the query as the limb vectors OpenFHE holds (K*E*2*L + 2*L vectors): gathered into pinned memory by the library
    std::vector<const uint64_t*> idxLimbs(st.K * st.E * 2 * L), minusLimbs(2 * L);
    for (size_t hf = 0; hf < st.K; hf++) {
        if (indexMatrix[hf].size() != st.E) throw std::invalid_argument("BatchedFHEHIPPIE (B200): indexMatrix row length must equal the cuckoo table size");
        for (size_t pos = 0; pos < st.E; pos++) limbs_of(indexMatrix[hf][pos], L, N, &idxLimbs[(hf * st.E + pos) * 2 * L]);
    }
    limbs_of(minusCompareElement, L, N, minusLimbs.data());
    // b result ciphertexts: the library scatters straight into the limb vectors that become the DCRTPolys
    const auto& params = minusCompareElement->GetElements()[0].GetParams();
    std::vector<NativeVector> vecs;
    vecs.reserve(st.b * 2 * L);
    std::vector<uint64_t*> outLimbs(st.b * 2 * L);
    for (size_t bin = 0; bin < st.b; bin++)
        for (size_t c = 0; c < 2; c++)
            for (size_t l = 0; l < L; l++) {
                vecs.emplace_back((usint)N, params->GetParams()[l]->GetModulus());
                outLimbs[(bin * 2 + c) * L + l] = words_of(vecs.back());
            }
    // setIndex + setMinusCompareElement + run + getResultList of the session as one call: on one device the host
    // gather, the upload slices, the evaluation, the download groups and the scatter overlap inside the query
    ck(psi_multi_query_run_limbs(st.multi, idxLimbs.data(), minusLimbs.data(), outLimbs.data()));
    for (size_t bin = 0; bin < st.b; bin++) {
        std::vector<FHEEncType> cv;
        for (size_t c = 0; c < 2; c++) {
            FHEEncType poly(params, Format::EVALUATION, true);
            for (size_t l = 0; l < L; l++) {
                NativePoly limb = poly.GetElementAtIndex((usint)l);
                limb.SetValues(std::move(vecs[(bin * 2 + c) * L + l]), Format::EVALUATION);
                poly.SetElementAtIndex((usint)l, std::move(limb));
            }
            cv.push_back(std::move(poly));
        }
        auto ct = minusCompareElement->CloneEmpty();
        ct->SetElements(std::move(cv));
        resultList[bin] = ct;
    }
}

// The reference class has no destructor; a server that builds more than one PIE per process may release the device
// state of a dead object explicitly.
extern "C" void psi_b200_adapter_release(const void* pie) {
    std::lock_guard<std::mutex> lock(g_mutex);
    g_state.erase(static_cast<const BatchedFHEHIPPIE*>(pie));
}

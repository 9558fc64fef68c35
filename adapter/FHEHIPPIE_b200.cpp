// Drop-in replacement for the reference's
//   src/Common/Crypto/PrivateIndexedEqualityCheck/FHEHIPPIE.cpp
// compiled against the reference's UNMODIFIED header (FHEHIPPIE.hpp:18-50: same class, same entry points, same
// members).  It keeps the reference's behaviour at the boundary - the std::invalid_argument cases (FHEHIPPIE.cpp:13-20),
// the bin permutation permVec2 (:29), plainVec with the trailing 1 (:44-50), random non-zero masks (:54),
// shuffledResultList[permutationVector[hf]] (:74) - but every homomorphic operation of run() (:61-77: EvalInnerProduct,
// EvalMerge, EvalMult) happens on the GPU behind the psi_nb_* calls of include/psi_b200.h.
//
// In the reference tree: list this file instead of FHEHIPPIE.cpp in src/CMakeLists.txt and link psi_b200.  Here it is
// compiled against adapter/shim/ and run end to end by adapter/test_adapter_nb.cpp (OpenFHE is absent from the image;
// lbcrypto member names as recalled from 1.0.x, to be checked against the installed headers).
//
// The header has no room for new members and the reference builds one FHEHIPPIE object per outer cell, all on one
// CryptoContext (SimpleFHEPSIServer.cpp:99-121).  The PIEs of a context therefore share ONE device context and ONE
// device database (a registry keyed by the context): constructors only record slot values; the first run() encodes the
// whole collection (MakePackedPlaintext of every plainVec / mask on the device) and installs the automorphism keys the
// context received (DeserializeEvalSumKey / DeserializeEvalAutomorphismKey, SimpleFHEPSIServer.cpp:45-62).
// run() on one object evaluates that PIE (FHEHIPPIECollection::runAll is header-only in the reference and keeps calling
// run() object by object); a caller that wants all PIEs in lock step uses psi_nb_run over the whole range, as
// nested-hashing-psi_b200/host/FHEHIPPIE.hpp's FHEHIPPIECollection::runAll does.
// Device: first entry of PSI_B200_DEVICES, default 0.
#include "FHEHIPPIE.hpp"

#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <unordered_map>

#include "adapter_common.hpp"

using namespace lbcrypto;
using namespace psi_adapter;

namespace {

struct Collection {
    psi_ctx* ctx = nullptr;
    size_t K = 0, b = 0, L = 0, N = 0;
    std::vector<int64_t> slots, masks;  // [n_pie][K][b][b+1], [n_pie][K][b]
    std::vector<FHEHIPPIE*> pies;
    bool encoded = false, keys = false;
    ~Collection() { psi_ctx_destroy(ctx); }
};
std::mutex g_mutex;
std::map<const void*, std::unique_ptr<Collection>> g_collections;       // by CryptoContextImpl*
std::unordered_map<const FHEHIPPIE*, std::pair<Collection*, uint32_t>> g_pie;  // object -> (collection, number)

int first_device() {
    const char* env = std::getenv("PSI_B200_DEVICES");
    return env && *env ? std::atoi(env) : 0;
}

// the automorphism keys of the context as [n][L][L][N] word arrays (BV, digit size 0: L digits of L limbs each)
void install_keys(Collection& c, const std::string& keyTag) {
    auto& keyMap = CryptoContextImpl<FHEEncType>::GetEvalAutomorphismKeyMap(keyTag);
    if (keyMap.empty()) throw std::runtime_error("FHEHIPPIE (B200): the context holds no automorphism keys");
    const size_t L = c.L, N = c.N, words = L * L * N;
    std::vector<uint64_t> index, kb(keyMap.size() * words), ka(kb.size());
    size_t n = 0;
    for (auto& [g, ek] : keyMap) {
        const auto& bv = ek->GetBVector();
        const auto& av = ek->GetAVector();
        if (bv.size() != L || av.size() != L) throw std::invalid_argument("FHEHIPPIE (B200): automorphism keys are not BV / digit size 0");
        for (size_t i = 0; i < L; i++)
            for (size_t l = 0; l < L; l++) {
                std::memcpy(&kb[n * words + (i * L + l) * N], words_of(bv[i].GetElementAtIndex((usint)l).GetValues()), N * sizeof(uint64_t));
                std::memcpy(&ka[n * words + (i * L + l) * N], words_of(av[i].GetElementAtIndex((usint)l).GetValues()), N * sizeof(uint64_t));
            }
        index.push_back(g);
        n++;
    }
    ck(psi_nb_set_automorphism_keys(c.ctx, (uint32_t)n, index.data(), kb.data(), ka.data()));
    c.keys = true;
}

// evaluates PIEs [begin, end) of a collection and fills their shuffledResultList
void run_range(Collection& c, uint32_t begin, uint32_t end, std::vector<Ciphertext<FHEEncType>>* const* results,
               const std::vector<Ciphertext<FHEEncType>>* const* index, const std::vector<uint>* const* perm, const std::string& keyTag) {
    const size_t K = c.K, L = c.L, N = c.N, ctWords = 2 * L * N;
    if (!c.keys) install_keys(c, keyTag);
    if (!c.encoded) {
        ck(psi_nb_db_encode_slots(c.ctx, (uint32_t)c.pies.size(), (uint32_t)K, (uint32_t)c.b, (uint32_t)c.b + 1, c.slots.data(), c.masks.data()));
        c.encoded = true;
    }
    std::vector<uint64_t> idx((size_t)(end - begin) * K * ctWords), out(idx.size());
    std::vector<const uint64_t*> limbs(2 * L);
    for (uint32_t p = begin; p < end; p++) {
        const auto& im = *index[p - begin];
        if (im.size() != K) throw std::invalid_argument("FHEHIPPIE (B200): setIndex needs one ciphertext per cuckoo hash function");
        for (size_t hf = 0; hf < K; hf++) {
            limbs_of(im[hf], L, N, limbs.data());
            for (size_t v = 0; v < 2 * L; v++) std::memcpy(&idx[((p - begin) * K + hf) * ctWords + v * N], limbs[v], N * sizeof(uint64_t));
        }
    }
    ck(psi_nb_run(c.ctx, begin, end, idx.data(), out.data(), nullptr));
    for (uint32_t p = begin; p < end; p++) {
        const auto& proto = (*index[p - begin])[0];
        const auto& params = proto->GetElements()[0].GetParams();
        for (size_t hf = 0; hf < K; hf++) {
            std::vector<FHEEncType> cv;
            for (size_t comp = 0; comp < 2; comp++) {
                FHEEncType poly(params, Format::EVALUATION, true);
                for (size_t l = 0; l < L; l++) {
                    NativeVector v((usint)N, params->GetParams()[l]->GetModulus());
                    std::memcpy(words_of(v), &out[((p - begin) * K + hf) * ctWords + (comp * L + l) * N], N * sizeof(uint64_t));
                    NativePoly limb = poly.GetElementAtIndex((usint)l);
                    limb.SetValues(std::move(v), Format::EVALUATION);
                    poly.SetElementAtIndex((usint)l, std::move(limb));
                }
                cv.push_back(std::move(poly));
            }
            auto ct = proto->CloneEmpty();
            ct->SetElements(std::move(cv));
            (*results[p - begin])[(*perm[p - begin])[hf]] = ct;  // shuffledResultList[permutationVector[hfInd]] = result
        }
    }
}

}  // namespace

FHEHIPPIE::FHEHIPPIE(lbcrypto::CryptoContext<FHEEncType>& cryptor, lbcrypto::PublicKey<FHEEncType>& pK, CuckooHashTable& ct)
    : cryptor(cryptor), pK(pK), numberOfResultElements(ct.getNumberOfHashFunctions()) {
    if (ct.getBinSize() != ct.getEachTableSize()) {
        throw invalid_argument("Error, for FHE PIE the size of a cuckoo bin has to be equal than the number of bins per hash function.");
    }
    if (ct.stash.size() != 0) {
        throw invalid_argument("Error, FHE PIE does not support a stash (yet).");
    }
    initPermutationVector(numberOfResultElements);
    // another Perm vector, hide correct bin index
    auto permVec2 = createPermutationVector(ct.getBinSize());
    auto plaintextModulus = cryptor->GetCryptoParameters()->GetPlaintextModulus();
    std::random_device rd;
    std::mt19937_64 mt(((uint64_t)rd() << 32) ^ rd());

    std::lock_guard<std::mutex> lock(g_mutex);
    auto& slot = g_collections[cryptor.get()];
    if (!slot) {
        slot.reset(new Collection());
        const psi_params P = params_from(cryptor, /*need_mult=*/false);
        slot->L = P.L;
        slot->N = P.N;
        slot->K = ct.getNumberOfHashFunctions();
        slot->b = ct.getBinSize();
        ck(psi_ctx_create(&P, first_device(), &slot->ctx));
    }
    Collection& c = *slot;
    if (c.K != ct.getNumberOfHashFunctions() || c.b != ct.getBinSize())
        throw invalid_argument("Error, all FHE PIEs of one context need the same table shape.");
    const size_t K = c.K, b = c.b, E = ct.getEachTableSize(), ns = E + 1, base = c.slots.size(), mbase = c.masks.size();
    c.slots.resize(base + K * b * ns, 0);
    c.masks.resize(mbase + K * b, 0);
    for (size_t hfInd = 0; hfInd < K; hfInd++)
        for (size_t binIndex = 0; binIndex < b; binIndex++) {
            // Add exponent for "minus client" element 1
            int64_t* plainVec = &c.slots[base + (hfInd * b + permVec2[binIndex]) * ns];
            for (size_t hashPos = 0; hashPos < E; hashPos++) plainVec[hashPos] = (int64_t)ct.cuckooTable[ct.getTableIndex(hfInd)][binIndex][hashPos];
            plainVec[E] = 1;
            c.masks[mbase + hfInd * b + binIndex] = (int64_t)(mt() % (plaintextModulus - 1) + 1);  // without 0
        }
    g_pie[this] = {&c, (uint32_t)c.pies.size()};
    c.pies.push_back(this);
    c.encoded = false;
}

void FHEHIPPIE::run() {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_pie.find(this);
    if (it == g_pie.end()) throw std::runtime_error("FHEHIPPIE (B200): object has no device state");
    Collection& c = *it->second.first;
    const uint32_t n = it->second.second;
    std::vector<Ciphertext<FHEEncType>>* res[1] = {&shuffledResultList};
    const std::vector<Ciphertext<FHEEncType>>* idx[1] = {&indexMatrix};
    const std::vector<uint>* perm[1] = {&permutationVector};
    run_range(c, n, n + 1, res, idx, perm, pK->GetKeyTag());
}

// The reference class has no destructor: release the device state of a context's collection explicitly.
extern "C" void psi_b200_nb_adapter_release(const void* cryptoContextImpl) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_collections.find(cryptoContextImpl);
    if (it == g_collections.end()) return;
    for (FHEHIPPIE* p : it->second->pies) g_pie.erase(p);
    g_collections.erase(it);
}
// debugging / tests: the device context behind a CryptoContext's collection (e.g. psi_nb_db_get_limbs)
extern "C" psi_ctx* psi_b200_nb_adapter_ctx(const void* cryptoContextImpl) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_collections.find(cryptoContextImpl);
    return it == g_collections.end() ? nullptr : it->second->ctx;
}

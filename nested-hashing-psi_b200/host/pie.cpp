// BatchedFHEHIPPIE (see BatchedFHEHIPPIE.hpp) and the C ABI of the host-side objects around it.
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>

#include "BatchedFHEHIPPIE.hpp"
#include "psi_host_internal.hpp"

namespace psi {

// A failed C-ABI call travelling through the C++ class: keeps the original status code.
struct AbiError : std::runtime_error {
    int rc;
    AbiError(int rc_, const std::string& m) : std::runtime_error(m), rc(rc_) {}
};
static void check(int rc, const char* what) {
    if (rc != PSI_OK) throw AbiError(rc, std::string(what) + ": " + psi_last_error());
}

BatchedFHEHIPPIE::BatchedFHEHIPPIE(CryptoContext& cryptor, PublicKey& pK_, HierarchicalCuckooHashTable& hct,
                                   uint64_t shuffleSeed, uint64_t maskSeed, bool keepSlots)
    : cryptoContext(cryptor), pK(pK_) {
    // BatchedFHEHIPPIE.cpp:13-21
    if (hct.getServerStashSize() != 0) throw std::invalid_argument("Error, batched FHE PIE does not support a stash (yet).");
    if (!hct.hasSimpleMultiTables() || !hct.hasCuckooMultiTables())
        throw std::invalid_argument("Error, batched FHE PIE currently does not support combined tables.");

    // Shuffle bins beforehand (BatchedFHEHIPPIE.cpp:25-35): the rows of every hash-function table of
    // every inner cuckoo table are permuted; this mutates the caller's table, as the reference does.
    std::mt19937 mt = seeded_mt19937(resolve_seed(shuffleSeed));
    for (auto& hctRow : hct.hierarchicalCuckooTable)
        for (auto& ct : hctRow) ct.shuffleBins(mt);

    K = hct.getNumberOfCuckooHashFunctions();
    b = (uint32_t)hct.getEachBinSize();
    E = (uint32_t)hct.getEachCuckooTableSize();
    const size_t k = hct.getNumberOfSimpleTables(), e = hct.getEachSimpleTableSize();
    const size_t batchSize = k * e;  // assumes simple multi table (BatchedFHEHIPPIE.cpp:41)
    if (batchSize > cryptoContext.params.N) throw std::invalid_argument("batch size exceeds the ring dimension");
    nslots = (uint32_t)batchSize;
    const uint64_t t = cryptoContext.GetPlaintextModulus();

    // Transposition hct[outerHf][outerPos].cuckooTable[innerHf][bin][innerPos] ->
    // plainVec[innerHf][bin][innerPos][outerHf * e + outerPos]   (BatchedFHEHIPPIE.cpp:48-66)
    std::vector<int64_t> slots((size_t)K * b * E * batchSize);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < (int64_t)batchSize; s++) {
        const CuckooHashTable& cur = hct.hierarchicalCuckooTable[s / e][s % e];
        for (uint32_t hf = 0; hf < K; hf++)
            for (uint32_t bin = 0; bin < b; bin++)
                for (uint32_t pos = 0; pos < E; pos++)
                    slots[(((size_t)hf * b + bin) * E + pos) * batchSize + s] = (int64_t)cur.cell(hf, bin, pos);
    }
    // Random non-zero masks r in [1, t-1] (BatchedFHEHIPPIE.cpp:73-82); generator documented in
    // DESIGN.md (the reference draws from a random_device-seeded boost::mt19937, so no sequence is pinned).
    std::vector<int64_t> maskSlots((size_t)b * batchSize);
    std::mt19937_64 mm(resolve_seed(maskSeed));
    for (auto& v : maskSlots) v = (int64_t)(mm() % (t - 1) + 1);

    if (cryptoContext.multi)
        check(psi_multi_db_encode_slots(cryptoContext.multi, K, b, E, nslots, slots.data(), maskSlots.data()),
              "MakePackedPlaintext on the devices");
    else
        check(psi_db_encode_slots(cryptoContext.device_ctx, K, b, E, nslots, slots.data(), maskSlots.data()),
              "MakePackedPlaintext on device");
    if (keepSlots) {
        keptSlots.swap(slots);
        keptMaskSlots.swap(maskSlots);
    }
    resultList = std::vector<Ciphertext>(b);
}

void BatchedFHEHIPPIE::upload() {
    if (uploaded) return;
    const size_t ct = (size_t)2 * cryptoContext.params.L * cryptoContext.params.N;
    if (indexMatrix.size() != K) throw std::invalid_argument("indexMatrix must have one row per cuckoo hash function");
    std::vector<uint64_t> flat((size_t)K * E * ct);
    for (uint32_t hf = 0; hf < K; hf++) {
        if (indexMatrix[hf].size() != E) throw std::invalid_argument("indexMatrix row length must equal the cuckoo table size");
        for (uint32_t pos = 0; pos < E; pos++) {
            const Ciphertext& c = indexMatrix[hf][pos];
            if (!c || c->size() != ct) throw std::invalid_argument("index ciphertext has the wrong number of limbs");
            std::memcpy(&flat[((size_t)hf * E + pos) * ct], c->data(), ct * sizeof(uint64_t));
        }
    }
    if (!minusCompareElement || minusCompareElement->size() != ct)
        throw std::invalid_argument("minusCompareElement is not set or has the wrong number of limbs");
    setQueryFlat(flat.data(), minusCompareElement->data());
}

void BatchedFHEHIPPIE::setQueryFlat(const uint64_t* idx, const uint64_t* minus) {
    if (cryptoContext.multi) {
        check(psi_multi_query_set(cryptoContext.multi, idx, minus), "setIndex / setMinusCompareElement");
        check(psi_multi_sync(cryptoContext.multi), "query upload");
    } else {
        check(psi_query_set(cryptoContext.device_ctx, idx, minus, nullptr), "setIndex / setMinusCompareElement");
        check(psi_stream_sync(nullptr), "query upload");
    }
    uploaded = true;
}

void BatchedFHEHIPPIE::run() {
    if (!uploaded && cryptoContext.multi) {
        // the ciphertexts the reference hands over are separate limb containers: no flat copy on the host, the
        // staging pool of psi_multi_query_set_limbs gathers them while the first pieces are already uploading
        const size_t L = cryptoContext.params.L, N = cryptoContext.params.N, ct = 2 * L * N;
        if (indexMatrix.size() != K) throw std::invalid_argument("indexMatrix must have one row per cuckoo hash function");
        std::vector<const uint64_t*> limbs((size_t)K * E * 2 * L), mlimbs(2 * L);
        for (uint32_t hf = 0; hf < K; hf++) {
            if (indexMatrix[hf].size() != E) throw std::invalid_argument("indexMatrix row length must equal the cuckoo table size");
            for (uint32_t pos = 0; pos < E; pos++) {
                const Ciphertext& c = indexMatrix[hf][pos];
                if (!c || c->size() != ct) throw std::invalid_argument("index ciphertext has the wrong number of limbs");
                for (size_t v = 0; v < 2 * L; v++) limbs[((size_t)hf * E + pos) * 2 * L + v] = c->data() + v * N;
            }
        }
        if (!minusCompareElement || minusCompareElement->size() != ct)
            throw std::invalid_argument("minusCompareElement is not set or has the wrong number of limbs");
        for (size_t v = 0; v < 2 * L; v++) mlimbs[v] = minusCompareElement->data() + v * N;
        check(psi_multi_query_set_limbs(cryptoContext.multi, limbs.data(), mlimbs.data()), "setIndex / setMinusCompareElement");
        uploaded = true;
    }
    upload();
    if (cryptoContext.multi)
        check(psi_multi_run(cryptoContext.multi), "run");
    else
        check(psi_run(cryptoContext.device_ctx, nullptr), "run");
    resultsFetched = false;
}

void BatchedFHEHIPPIE::getResultFlat(uint64_t* out) {
    if (cryptoContext.multi) {
        check(psi_multi_result_get(cryptoContext.multi, out), "getResultList");
        check(psi_multi_sync(cryptoContext.multi), "getResultList");
        return;
    }
    check(psi_result_get(cryptoContext.device_ctx, out, nullptr), "getResultList");
    check(psi_stream_sync(nullptr), "getResultList");
}

std::vector<Ciphertext>& BatchedFHEHIPPIE::getResultList() {
    if (!resultsFetched && cryptoContext.multi) {
        // straight into the b result containers (scatter from the pinned pool, no flat intermediate)
        const size_t L = cryptoContext.params.L, N = cryptoContext.params.N, ct = 2 * L * N;
        std::vector<uint64_t*> limbs((size_t)b * 2 * L);
        for (uint32_t bin = 0; bin < b; bin++) {
            resultList[bin] = std::make_shared<std::vector<uint64_t>>(ct);
            for (size_t v = 0; v < 2 * L; v++) limbs[(size_t)bin * 2 * L + v] = resultList[bin]->data() + v * N;
        }
        check(psi_multi_result_get_limbs(cryptoContext.multi, limbs.data()), "getResultList");
        resultsFetched = true;
    }
    if (!resultsFetched) {
        const size_t ct = (size_t)2 * cryptoContext.params.L * cryptoContext.params.N;
        std::vector<uint64_t> flat((size_t)b * ct);
        getResultFlat(flat.data());
        for (uint32_t bin = 0; bin < b; bin++)
            resultList[bin] = std::make_shared<std::vector<uint64_t>>(flat.begin() + (size_t)bin * ct,
                                                                      flat.begin() + (size_t)(bin + 1) * ct);
        resultsFetched = true;
    }
    return resultList;
}

}  // namespace psi

// ---------------------------------------------------------------------------------------------
// C ABI of the host objects (declared in include/psi_b200.h)
// ---------------------------------------------------------------------------------------------
struct psi_hct {
    psi::TabulationHashing hash;
    psi::HierarchicalCuckooHashTable table;
    psi_hct(uint64_t seed, unsigned k, uint64_t e, unsigned K, uint64_t E, uint64_t b, uint64_t stash, bool sm, bool cm,
            uint64_t evict)
        : hash(seed, k + K), table(hash, e, E, stash, k, K, sm, cm, b, evict) {}
};

struct psi_pie {
    psi::CryptoContext cc;
    psi::PublicKey pk;
    std::unique_ptr<psi::BatchedFHEHIPPIE> pie;
};

template <typename F>
static int guarded(F&& f) {
    try {
        f();
        return PSI_OK;
    } catch (const std::invalid_argument& e) {
        return psi::set_error(PSI_ERR_INVALID, e.what());
    } catch (const psi::AbiError& e) {
        return psi::set_error(e.rc, e.what());
    } catch (const std::bad_alloc&) {
        return psi::set_error(PSI_ERR_INVALID, "out of host memory");
    } catch (const std::exception& e) {
        // device failures keep the status the C layer recorded in the message; hashing failures
        // (std::runtime_error "(Blocked) Cuckoo hashing error", CuckooHashTable.cpp:113) map to STATE
        return psi::set_error(PSI_ERR_STATE, e.what());
    }
}

extern "C" {

int psi_hct_create(uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, uint64_t E, uint64_t b, uint64_t stash,
                   int simple_multi, int cuckoo_multi, uint64_t eviction_seed, psi_hct** out) {
    if (!out) return psi::set_error(PSI_ERR_INVALID, "null argument");
    *out = nullptr;
    if (k < 1 || e < 1 || E < 1) return psi::set_error(PSI_ERR_INVALID, "table sizes must be positive");
    return guarded([&] { *out = new psi_hct(hash_seed, k, e, K, E, b, stash, simple_multi != 0, cuckoo_multi != 0, eviction_seed); });
}

int psi_hct_insert_all(psi_hct* h, const uint64_t* items, size_t n) {
    if (!h || (!items && n)) return psi::set_error(PSI_ERR_INVALID, "null argument");
    return guarded([&] { h->table.insertAll(items, n); });
}

int psi_hct_get_cells(psi_hct* h, uint64_t* out) {
    if (!h || !out) return psi::set_error(PSI_ERR_INVALID, "null argument");
    auto& T = h->table;
    const size_t k = T.getNumberOfSimpleTables(), e = T.getEachSimpleTableSize(), K = T.getNumberOfCuckooHashFunctions(),
                 b = T.getEachBinSize(), E = T.getEachCuckooTableSize();
    for (size_t i = 0; i < k; i++)
        for (size_t j = 0; j < e; j++)
            for (size_t hf = 0; hf < K; hf++)
                for (size_t bin = 0; bin < b; bin++)
                    for (size_t pos = 0; pos < E; pos++)
                        out[((((i * e + j) * K + hf) * b + bin) * E) + pos] = T.hierarchicalCuckooTable[i][j].cell((unsigned)hf, bin, pos);
    return PSI_OK;
}

int psi_hct_destroy(psi_hct* h) {
    delete h;
    return PSI_OK;
}

int psi_hash_index(uint64_t hash_seed, uint32_t n_hash_functions, const uint64_t* items, size_t n, uint32_t hf,
                   uint32_t table_size, uint64_t* out) {
    if (!items || !out) return psi::set_error(PSI_ERR_INVALID, "null argument");
    if (hf >= n_hash_functions || table_size == 0) return psi::set_error(PSI_ERR_INVALID, "hash function index / table size");
    return guarded([&] {
        psi::TabulationHashing h(hash_seed, n_hash_functions);
        for (size_t i = 0; i < n; i++) out[i] = psi::calculateHashIndex(h, items[i], hf, table_size);
    });
}

int psi_client_table(uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, const uint64_t* items, size_t n,
                     uint64_t eviction_seed, uint64_t* cells) {
    // BatchedFHEPSIClient.cpp:97-99,109: the client's own cuckoo table, k tables x e positions x 1 item
    if (!items || !cells) return psi::set_error(PSI_ERR_INVALID, "null argument");
    return guarded([&] {
        psi::TabulationHashing h(hash_seed, k + K);
        psi::CuckooHashTable t(h, e, k, 0, 0, true, 1, eviction_seed);
        t.insertAll(items, n);
        for (uint32_t i = 0; i < k; i++)
            for (uint64_t j = 0; j < e; j++) cells[i * e + j] = t.cell(i, 0, j);
    });
}

int psi_random_data_input(size_t server_size, size_t client_size, size_t intersection_size, uint64_t seed,
                          uint64_t bit_size, uint64_t* server, uint64_t* client, uint64_t* intersection) {
    return guarded([&] {
        psi::RandomDataInput d(server_size, client_size, intersection_size, seed, bit_size);
        if (server) std::memcpy(server, d.serverSet.data(), server_size * sizeof(uint64_t));
        if (client) std::memcpy(client, d.clientSet.data(), client_size * sizeof(uint64_t));
        if (intersection) std::memcpy(intersection, d.intersectionSet.data(), intersection_size * sizeof(uint64_t));
    });
}

int psi_pie_create(psi_ctx* ctx, const psi_params* params, psi_hct* hct, uint64_t shuffle_seed, uint64_t mask_seed,
                   int keep_slots, psi_pie** out) {
    if (!ctx || !params || !hct || !out) return psi::set_error(PSI_ERR_INVALID, "null argument");
    *out = nullptr;
    psi_pie* p = new (std::nothrow) psi_pie();
    if (!p) return psi::set_error(PSI_ERR_INVALID, "out of host memory");
    p->cc.params = *params;
    p->cc.device_ctx = ctx;
    int rc = guarded([&] {
        p->pie.reset(new psi::BatchedFHEHIPPIE(p->cc, p->pk, hct->table, shuffle_seed, mask_seed, keep_slots != 0));
    });
    if (rc != PSI_OK) {
        delete p;
        return rc;
    }
    *out = p;
    return PSI_OK;
}

int psi_pie_create_multi(psi_multi* m, const psi_params* params, psi_hct* hct, uint64_t shuffle_seed, uint64_t mask_seed,
                         int keep_slots, psi_pie** out) {
    if (!m || !params || !hct || !out) return psi::set_error(PSI_ERR_INVALID, "null argument");
    *out = nullptr;
    psi_pie* p = new (std::nothrow) psi_pie();
    if (!p) return psi::set_error(PSI_ERR_INVALID, "out of host memory");
    p->cc.params = *params;
    p->cc.device_ctx = nullptr;
    p->cc.multi = m;
    int rc = guarded([&] {
        p->pie.reset(new psi::BatchedFHEHIPPIE(p->cc, p->pk, hct->table, shuffle_seed, mask_seed, keep_slots != 0));
    });
    if (rc != PSI_OK) {
        delete p;
        return rc;
    }
    *out = p;
    return PSI_OK;
}

int psi_pie_dims(psi_pie* p, uint32_t* K, uint32_t* b, uint32_t* E, uint32_t* nslots) {
    if (!p) return psi::set_error(PSI_ERR_INVALID, "null argument");
    if (K) *K = p->pie->numberOfCuckooHashFunctions();
    if (b) *b = p->pie->binSize();
    if (E) *E = p->pie->cuckooTableSize();
    if (nslots) *nslots = p->pie->batchSize();
    return PSI_OK;
}

int psi_pie_get_slots(psi_pie* p, int64_t* slots, int64_t* mask_slots) {
    if (!p) return psi::set_error(PSI_ERR_INVALID, "null argument");
    if (p->pie->keptSlots.empty()) return psi::set_error(PSI_ERR_STATE, "the PIE was created without keep_slots");
    if (slots) std::memcpy(slots, p->pie->keptSlots.data(), p->pie->keptSlots.size() * sizeof(int64_t));
    if (mask_slots) std::memcpy(mask_slots, p->pie->keptMaskSlots.data(), p->pie->keptMaskSlots.size() * sizeof(int64_t));
    return PSI_OK;
}

int psi_pie_set_query(psi_pie* p, const uint64_t* idx, const uint64_t* minus) {
    if (!p || !idx || !minus) return psi::set_error(PSI_ERR_INVALID, "null argument");
    return guarded([&] { p->pie->setQueryFlat(idx, minus); });
}

int psi_pie_run(psi_pie* p) {
    if (!p) return psi::set_error(PSI_ERR_INVALID, "null argument");
    return guarded([&] { p->pie->run(); });
}

int psi_pie_get_result_list(psi_pie* p, uint64_t* out) {
    if (!p || !out) return psi::set_error(PSI_ERR_INVALID, "null argument");
    return guarded([&] { p->pie->getResultFlat(out); });
}

int psi_pie_destroy(psi_pie* p) {
    delete p;
    return PSI_OK;
}

}  // extern "C"

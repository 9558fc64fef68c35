// See hashing.hpp for the reference line map.
#include "hashing.hpp"

#include <algorithm>
#include <numeric>

namespace psi {

TabulationHashing::TabulationHashing(uint64_t seed, size_t numberOfHashfunctions)
    : nHashfunctions(numberOfHashfunctions), tTable(numberOfHashfunctions * kChunks * (1u << kChunkBits)) {
    // TabulationHashing.cpp:22-36: std::mt19937 + std::uniform_int_distribution<uint64_t>, filled
    // in [hf][chunk][value] order.  libstdc++ semantics, same as the reference's build.
    std::mt19937 gen(seed);
    std::uniform_int_distribution<uint64_t> dis;
    for (auto& v : tTable) v = dis(gen);
}

uint64_t TabulationHashing::hashWithIndicator(item_t input, unsigned hfInd) const {
    // TabulationHashing.cpp:45-54: XOR of one table entry per input byte, 16 bytes consumed; a
    // 64-bit item contributes zero bytes for chunks 8..15.
    const uint64_t* t = &tTable[(size_t)hfInd * kChunks * 256];
    uint64_t res = 0;
    for (size_t i = 0; i < kChunks; i++) {
        unsigned byte = i < 8 ? (unsigned)((input >> (8 * i)) & 0xff) : 0u;
        res ^= t[i * 256 + byte];
    }
    return res;
}

CuckooHashTable::CuckooHashTable(const TabulationHashing& hashfunction, uint64_t eachTableSize,
                                 unsigned numberOfHashFunctions, unsigned startingHashId, uint64_t maxStashSize,
                                 bool multipleTables, uint64_t maxItemsPerPosition, uint64_t evictionSeed)
    : stash(maxStashSize, 0),
      hash(&hashfunction),
      tableSize(eachTableSize),
      nHf(numberOfHashFunctions),
      startId(startingHashId),
      binSize(maxItemsPerPosition),
      mt((uint32_t)evictionSeed) {
    if (numberOfHashFunctions < 2) throw std::invalid_argument("Cuckoo Table needs more than one hash function!");
    if (maxItemsPerPosition < 1) throw std::invalid_argument("Bin size needs to be at least of size one!");
    if (!multipleTables) throw std::invalid_argument("combined cuckoo tables are not supported by the batched FHE path");
    cells.assign((size_t)nHf * binSize * tableSize, 0);
}

bool CuckooHashTable::lookUp(item_t element) const {
    // CuckooHashTable.cpp:135-167
    for (unsigned hf = 0; hf < nHf; hf++) {
        uint64_t pos = calculateHashIndex(*hash, element, startId + hf, (uint32_t)tableSize);
        for (uint64_t bin = 0; bin < binSize; bin++) {
            item_t cur = cell(hf, bin, pos);
            if (cur == element) return true;
            if (cur == 0) break;
        }
    }
    for (item_t s : stash)
        if (s == element) return true;
    return false;
}

void CuckooHashTable::insert(item_t value) {
    // CuckooHashTable.cpp:72-114: first free bin at the hashed position of each function in turn,
    // otherwise evict a uniformly chosen bin and carry the evicted item to the next function.
    if (lookUp(value)) return;
    for (unsigned run = 0; run < kRetries; run++) {
        for (unsigned hf = 0; hf < nHf; hf++) {
            uint64_t pos = calculateHashIndex(*hash, value, startId + hf, (uint32_t)tableSize);
            for (uint64_t bin = 0; bin < binSize; bin++) {
                item_t& c = cell(hf, bin, pos);
                if (c == 0) {
                    c = value;
                    return;
                }
            }
            uint64_t victim = boost_uniform_u64(mt) % binSize;  // randomModRange, HashUtils.cpp:103-106
            std::swap(value, cell(hf, victim, pos));
        }
    }
    for (auto& s : stash) {
        if (s == 0) {
            s = value;
            return;
        }
    }
    throw std::runtime_error("(Blocked) Cuckoo hashing error");
}

void CuckooHashTable::insertAll(const item_t* elements, size_t n) {
    for (size_t i = 0; i < n; i++) insert(elements[i]);
}
void CuckooHashTable::insertAll(const std::vector<item_t>& elements) { insertAll(elements.data(), elements.size()); }

void CuckooHashTable::shuffleBins(std::mt19937& rng) {
    std::vector<uint32_t> perm(binSize);
    std::vector<item_t> tmp(binSize * tableSize);
    for (unsigned hf = 0; hf < nHf; hf++) {
        std::iota(perm.begin(), perm.end(), 0u);
        std::shuffle(perm.begin(), perm.end(), rng);
        item_t* base = &cells[(size_t)hf * binSize * tableSize];
        std::copy(base, base + binSize * tableSize, tmp.begin());
        for (uint64_t bin = 0; bin < binSize; bin++)
            std::copy(tmp.begin() + perm[bin] * tableSize, tmp.begin() + (perm[bin] + 1) * tableSize,
                      base + bin * tableSize);
    }
}

std::vector<uint16_t> makeBinShuffle(size_t nSimpleTables, uint64_t simpleSize, unsigned nCuckooHf, uint64_t binSize,
                                     std::mt19937& rng) {
    std::vector<uint16_t> out;
    out.reserve(nSimpleTables * simpleSize * nCuckooHf * binSize);
    std::vector<uint32_t> perm(binSize);
    for (size_t t = 0; t < nSimpleTables * simpleSize; t++)
        for (unsigned hf = 0; hf < nCuckooHf; hf++) {
            std::iota(perm.begin(), perm.end(), 0u);
            std::shuffle(perm.begin(), perm.end(), rng);  // same call, same order as CuckooHashTable::shuffleBins
            for (uint32_t v : perm) out.push_back((uint16_t)v);
        }
    return out;
}

HierarchicalCuckooHashTable::HierarchicalCuckooHashTable(const TabulationHashing& hashfunction,
                                                         uint64_t eachSimpleTableSize, uint64_t eachCuckooTableSize,
                                                         uint64_t serverStashSize, unsigned numberOfSimpleHashFunctions,
                                                         unsigned numberOfCuckooHashFunctions, bool simpleMultiTable,
                                                         bool cuckooMultiTable, uint64_t maxItemsPerPosition,
                                                         uint64_t evictionSeed)
    : hash(&hashfunction),
      simpleSize(eachSimpleTableSize),
      cuckooSize(eachCuckooTableSize),
      stashSize(serverStashSize),
      nSimpleHf(numberOfSimpleHashFunctions),
      nCuckooHf(numberOfCuckooHashFunctions),
      simpleMulti(simpleMultiTable),
      cuckooMulti(cuckooMultiTable),
      binSize(maxItemsPerPosition) {
    // HierarchicalCuckooHashTable.cpp:16-53: inner tables use hash ids k..k+K-1.  The reference seeds
    // every inner table's eviction RNG from std::random_device; here each gets a fixed seed so that
    // tables are reproducible and the OpenMP build below is deterministic.
    if (!simpleMulti || !cuckooMulti)
        throw std::invalid_argument("Error, batched FHE PIE currently does not support combined tables.");
    size_t nTables = getNumberOfSimpleTables();
    hierarchicalCuckooTable.resize(nTables);
    for (size_t i = 0; i < nTables; i++) {
        hierarchicalCuckooTable[i].reserve(simpleSize);
        for (uint64_t j = 0; j < simpleSize; j++)
            hierarchicalCuckooTable[i].emplace_back(hashfunction, cuckooSize, nCuckooHf, nSimpleHf, stashSize,
                                                    cuckooMulti, binSize, evictionSeed + 0x9e3779b9ull * (i * simpleSize + j + 1));
    }
}

void HierarchicalCuckooHashTable::insertAll(const item_t* elements, size_t n) {
    // HierarchicalCuckooHashTable.cpp:55-73: every element goes into ALL k simple tables; each simple
    // position is an independent inner cuckoo table (the reference's `#pragma omp parallel for`, :65).
    for (unsigned st = 0; st < nSimpleHf; st++) {
        // generateSimpleHashTable (HashUtils.cpp:48-60) as a counting sort: stable, so the insertion
        // order inside each bucket equals the reference's push_back order.
        std::vector<uint32_t> where(n);
        std::vector<size_t> start(simpleSize + 1, 0);
        for (size_t i = 0; i < n; i++) {
            where[i] = (uint32_t)calculateHashIndex(*hash, elements[i], st, (uint32_t)simpleSize);
            start[where[i] + 1]++;
        }
        for (uint64_t p = 0; p < simpleSize; p++) start[p + 1] += start[p];
        std::vector<item_t> bucketed(n);
        {
            std::vector<size_t> fill(start.begin(), start.end() - 1);
            for (size_t i = 0; i < n; i++) bucketed[fill[where[i]]++] = elements[i];
        }
        bool failed = false;
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t p = 0; p < (int64_t)simpleSize; p++) {
            try {
                hierarchicalCuckooTable[st][p].insertAll(bucketed.data() + start[p], start[p + 1] - start[p]);
            } catch (const std::runtime_error&) {
#pragma omp critical
                failed = true;
            }
        }
        if (failed) throw std::runtime_error("(Blocked) Cuckoo hashing error");
    }
}

// randomBiginteger (HashUtils.cpp:120-142) restricted to bitSize <= 64.
static item_t random_item(std::mt19937& mt, uint64_t bitSize) {
    uint64_t rest = bitSize % 64;
    if (rest != 0) return boost_uniform_u64(mt) % (1ull << rest);
    return boost_uniform_u64(mt);
}

RandomDataInput::RandomDataInput(size_t serverSetSize, size_t clientSetSize, size_t intersectionSetSize,
                                 uint64_t setGenerationSeed, uint64_t bitSize)
    : serverSet(serverSetSize), clientSet(clientSetSize), intersectionSet(intersectionSetSize) {
    if (clientSetSize > serverSetSize || intersectionSetSize > clientSetSize || bitSize == 0 || bitSize > 64)
        throw std::invalid_argument("RandomDataInput: inconsistent set sizes or bit size");
    const uint64_t serverSeedDiff = (1ull << 32) + (1ull << 16) + 1;  // RandomDataInput.hpp:22
    // boost::mt19937(uint64) seeds with the value narrowed to 32 bits, like std::mt19937
    std::mt19937 mtClient((uint32_t)setGenerationSeed), mtServer((uint32_t)(setGenerationSeed + serverSeedDiff));
    size_t onlyClient = clientSetSize - intersectionSetSize;
    for (size_t i = 0; i < onlyClient; i++) clientSet[i] = random_item(mtClient, bitSize);
    for (size_t i = 0; i < serverSetSize; i++) serverSet[i] = random_item(mtServer, bitSize);
    for (size_t i = 0; i < intersectionSetSize; i++) {
        intersectionSet[i] = serverSet[i];
        clientSet[onlyClient + i] = serverSet[i];
    }
}

}  // namespace psi

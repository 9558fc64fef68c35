// Host-side nested (hierarchical) cuckoo hashing for the BatchedFHEPIE path.
//
// The PIE constructor consumes a HierarchicalCuckooHashTable
// (/root/reference/src/Common/Crypto/PrivateIndexedEqualityCheck/BatchedFHEHIPPIE.hpp:30-31), so the
// drop-in class needs the same data structure.  The reference's own classes cannot be compiled here
// (they include libscapi's biginteger through src/PSIConfigs.h:8), therefore the semantics are
// restated with 64-bit items (the FHE path supports items of at most 48 bits,
// BatchedFHEPSIClient.cpp:23-41).  Offline phase only — nothing in here is on run()'s path.
//
//   TabulationHashing            src/Common/Hashing/TabulationHashing.cpp:16-54
//   calculateHashIndex           src/Common/Hashing/HashUtils.cpp:34-37
//   CuckooHashTable              src/Common/Hashing/CuckooHashTable.cpp:25-167
//   HierarchicalCuckooHashTable  src/Common/Hashing/HierarchicalCuckooHashTable.cpp:16-87
//   RandomDataInput              src/Common/DataInput/RandomDataInput.cpp:10-66, HashUtils.cpp:120-142
#pragma once
#include <cstdint>
#include <random>
#include <stdexcept>
#include <vector>

namespace psi {

typedef uint64_t item_t;  // 0 is the empty-cell marker (CuckooHashTable.cpp:90)

// One 64-bit draw of boost::random::uniform_int_distribution<uint64_t> over a 32-bit mt19937:
// Boost concatenates engine outputs LOW word first (restated from Boost.Random's
// generate_uniform_int; Boost itself is not available in this image).
inline uint64_t boost_uniform_u64(std::mt19937& mt) {
    uint64_t lo = mt();
    uint64_t hi = mt();
    return lo | (hi << 32);
}

// Generators of the security-critical randomness of the PIE constructor (bin shuffle, masks).  The reference
// seeds a 32-bit mt19937 from std::random_device (BatchedFHEHIPPIE.cpp:25-26); here all 64 bits of a seed are
// used (seed_seq over both halves), and PSI_SEED_RANDOM asks for a fresh std::random_device seed.
inline uint64_t resolve_seed(uint64_t seed) {
    if (seed != ~0ull) return seed;  // PSI_SEED_RANDOM
    std::random_device rd;
    return ((uint64_t)rd() << 32) | (uint64_t)rd();
}
inline std::mt19937 seeded_mt19937(uint64_t seed) {
    std::seed_seq sq{(uint32_t)seed, (uint32_t)(seed >> 32)};
    return std::mt19937(sq);
}

class TabulationHashing {
   public:
    static constexpr size_t kChunks = 16;  // tParam: bytes of input consumed
    static constexpr size_t kChunkBits = 8;  // rParam
    explicit TabulationHashing(uint64_t seed = 342797434736ull, size_t numberOfHashfunctions = 3);
    uint64_t hashWithIndicator(item_t input, unsigned hfInd) const;
    size_t numberOfHashFunctions() const { return nHashfunctions; }
    const std::vector<uint64_t>& tables() const { return tTable; }  // [hf][chunk][256], for the device build

   private:
    size_t nHashfunctions;
    std::vector<uint64_t> tTable;  // [hf][chunk][256]
};

inline uint64_t calculateHashIndex(const TabulationHashing& h, item_t value, unsigned hfInd, uint32_t tableSize) {
    return h.hashWithIndicator(value, hfInd) % tableSize;
}

// Blocked cuckoo table: cells[hf][bin][pos], `maxItemsPerPosition` bins per position.
class CuckooHashTable {
   public:
    CuckooHashTable(const TabulationHashing& hashfunction, uint64_t eachTableSize, unsigned numberOfHashFunctions = 2,
                    unsigned startingHashId = 0, uint64_t maxStashSize = 0, bool multipleTables = true,
                    uint64_t maxItemsPerPosition = 1, uint64_t evictionSeed = 0x5eedull);

    void insert(item_t value);
    void insertAll(const std::vector<item_t>& elements);
    void insertAll(const item_t* elements, size_t n);
    bool lookUp(item_t element) const;

    item_t& cell(unsigned hf, uint64_t bin, uint64_t pos) { return cells[(hf * binSize + bin) * tableSize + pos]; }
    item_t cell(unsigned hf, uint64_t bin, uint64_t pos) const { return cells[(hf * binSize + bin) * tableSize + pos]; }

    unsigned getNumberOfHashFunctions() const { return nHf; }
    unsigned getStartingHashIndex() const { return startId; }
    bool hasMultipleTables() const { return true; }
    uint64_t getBinSize() const { return binSize; }
    size_t getNumberOfTables() const { return nHf; }
    uint64_t getEachTableSize() const { return tableSize; }
    // permutes whole bin rows of every hash-function table (BatchedFHEHIPPIE.cpp:28-35)
    void shuffleBins(std::mt19937& mt);

    std::vector<item_t> stash;

   private:
    const TabulationHashing* hash;
    uint64_t tableSize;
    unsigned nHf;
    unsigned startId;
    uint64_t binSize;
    static constexpr unsigned kRetries = 1000;
    std::mt19937 mt;
    std::vector<item_t> cells;
};

class HierarchicalCuckooHashTable {
   public:
    HierarchicalCuckooHashTable(const TabulationHashing& hashfunction, uint64_t eachSimpleTableSize,
                                uint64_t eachCuckooTableSize, uint64_t serverStashSize = 0,
                                unsigned numberOfSimpleHashFunctions = 2, unsigned numberOfCuckooHashFunctions = 2,
                                bool simpleMultiTable = false, bool cuckooMultiTable = true,
                                uint64_t maxItemsPerPosition = 1, uint64_t evictionSeed = 0x5eedull);

    void insertAll(const std::vector<item_t>& elements) { insertAll(elements.data(), elements.size()); }
    void insertAll(const item_t* elements, size_t n);

    size_t getNumberOfSimpleTables() const { return simpleMulti ? nSimpleHf : 1; }
    uint64_t getEachSimpleTableSize() const { return simpleSize; }
    size_t getNumberOfCuckooTables() const { return cuckooMulti ? nCuckooHf : 1; }
    uint64_t getEachCuckooTableSize() const { return cuckooSize; }
    uint64_t getServerStashSize() const { return stashSize; }
    bool hasSimpleMultiTables() const { return simpleMulti; }
    bool hasCuckooMultiTables() const { return cuckooMulti; }
    uint64_t getEachBinSize() const { return binSize; }
    unsigned getNumberOfSimpleHashFunctions() const { return nSimpleHf; }
    unsigned getNumberOfCuckooHashFunctions() const { return nCuckooHf; }

    // hierarchicalCuckooTable[outerHf][outerPos]
    std::vector<std::vector<CuckooHashTable>> hierarchicalCuckooTable;

   private:
    const TabulationHashing* hash;
    uint64_t simpleSize, cuckooSize, stashSize;
    unsigned nSimpleHf, nCuckooHf;
    bool simpleMulti, cuckooMulti;
    uint64_t binSize;
};

// The bin shuffle of the PIE constructor (BatchedFHEHIPPIE.cpp:28-35) as explicit permutations, in the order
// shuffleBins() consumes the generator: [outerHf][outerPos][innerHf][bin] -> source bin.  The device-side
// constructor path applies them while transposing.
std::vector<uint16_t> makeBinShuffle(size_t nSimpleTables, uint64_t simpleSize, unsigned nCuckooHf, uint64_t binSize,
                                     std::mt19937& mt);

// Synthetic sets with a planted intersection (RandomDataInput): the first I draws of the server
// stream are the intersection; the client puts them in its last I positions.
struct RandomDataInput {
    RandomDataInput(size_t serverSetSize, size_t clientSetSize, size_t intersectionSetSize, uint64_t setGenerationSeed,
                    uint64_t bitSize);
    std::vector<item_t> serverSet, clientSet, intersectionSet;
};

}  // namespace psi

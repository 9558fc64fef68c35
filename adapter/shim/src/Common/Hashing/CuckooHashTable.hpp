// MINIMAL SHIM of the reference's src/Common/Hashing/CuckooHashTable.hpp (+ the one HashUtils function the PIE headers
// call), test infrastructure for the adapters.  The reference's own header pulls in libscapi (biginteger) and Boost,
// neither of which exists in this image.  Declared here: ONLY the members the PIE constructors touch
// (FHEHIPPIE.cpp:9-59, BatchedFHEHIPPIE.cpp:13-66) with the reference's names, over this repo's host table
// (host/hashing.hpp).  In the reference tree the adapters include the real header instead.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <random>
#include <vector>

#include "hashing.hpp"  // psi::CuckooHashTable, psi::TabulationHashing (nested-hashing-psi_b200/host)

using namespace std;  // the reference's headers rely on it (libscapi does the same)

typedef unsigned int uint;
typedef unsigned long long biginteger;  // libscapi: boost::multiprecision::cpp_int; the PIEs only cast cells to int64_t
typedef psi::TabulationHashing TabulationHashing;

// HashUtils.cpp:95-105
inline vector<uint> createPermutationVector(size_t size, int seed = 0) {
    std::random_device rd;
    std::mt19937 rng(rd());
    std::vector<uint> v(size);
    std::iota(std::begin(v), std::end(v), 0);
    std::shuffle(std::begin(v), std::end(v), rng);
    return v;
}

class CuckooHashTable {
    uint64_t eachTableSize = 0, maxItemsPerPosition = 1;
    uint numberOfHashFunctions = 0;
    TabulationHashing* hashfunction = nullptr;
    uint startingHashId = 0;

   public:
    vector<vector<vector<biginteger>>> cuckooTable;  // [hfInd] x [binIndex] x [index]
    vector<biginteger> stash;

    CuckooHashTable() = default;
    CuckooHashTable(TabulationHashing& hashfunction, uint64_t eachTableSize, uint numberOfHashFunctions = 2, uint startingHashId = 0,
                    uint64_t maxStashSize = 0, bool multipleTables = true, uint64_t maxItemsPerPosition = 1)
        : eachTableSize(eachTableSize), maxItemsPerPosition(maxItemsPerPosition), numberOfHashFunctions(numberOfHashFunctions),
          hashfunction(&hashfunction), startingHashId(startingHashId) {}
    void insertAll(vector<biginteger>& elements) {
        psi::CuckooHashTable impl(*hashfunction, eachTableSize, numberOfHashFunctions, startingHashId, 0, true, maxItemsPerPosition);
        std::vector<psi::item_t> items(elements.begin(), elements.end());
        impl.insertAll(items);
        cuckooTable.assign(numberOfHashFunctions,
                           vector<vector<biginteger>>(maxItemsPerPosition, vector<biginteger>(eachTableSize)));
        for (uint hf = 0; hf < numberOfHashFunctions; hf++)
            for (uint64_t bin = 0; bin < maxItemsPerPosition; bin++)
                for (uint64_t pos = 0; pos < eachTableSize; pos++) cuckooTable[hf][bin][pos] = impl.cell(hf, bin, pos);
        stash.assign(impl.stash.begin(), impl.stash.end());
    }
    uint getNumberOfHashFunctions() { return numberOfHashFunctions; }
    uint64_t getBinSize() { return maxItemsPerPosition; }
    size_t getNumberOfTables() { return cuckooTable.size(); }
    uint64_t getEachTableSize() { return eachTableSize; }
    uint64_t getTableIndex(uint hfInd) { return hfInd; }  // multipleTables: one table per hash function
};

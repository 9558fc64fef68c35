// MINIMAL SHIM of the reference's hashing header, test infrastructure for adapter/BatchedFHEHIPPIE_b200.cpp.
// The reference's own src/Common/Hashing/HierarchicalCuckooHashTable.hpp pulls in libscapi (biginteger) and Boost,
// neither of which exists in this image.  This file declares ONLY the members the PIE constructor touches
// (BatchedFHEHIPPIE.cpp:13-66) with the reference's names, over this repo's host table (host/hashing.hpp), so that
// the adapter compiles and runs here.  In the reference tree the adapter includes the real header instead.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

#include "CuckooHashTable.hpp"  // the shim of the inner table (cuckooTable [hfInd][binIndex][index]), biginteger, TabulationHashing

class HierarchicalCuckooHashTable {
    psi::HierarchicalCuckooHashTable impl;

   public:
    vector<vector<CuckooHashTable>> hierarchicalCuckooTable;
    HierarchicalCuckooHashTable(TabulationHashing& hashfunction, uint64_t eachSimpleTableSize, uint64_t eachCuckooTableSize,
                                uint64_t serverStashSize = 0, uint numberOfSimpleHashFunctions = 2,
                                uint numberOfCuckooHashFunctions = 2, bool simpleMultiTable = false, bool cuckooMultiTable = true,
                                uint64_t maxItemsPerPosition = 1)
        : impl(hashfunction, eachSimpleTableSize, eachCuckooTableSize, serverStashSize, numberOfSimpleHashFunctions,
               numberOfCuckooHashFunctions, simpleMultiTable, cuckooMultiTable, maxItemsPerPosition, 0x5eed) {}
    void insertAll(vector<biginteger>& elements) {
        std::vector<psi::item_t> items(elements.begin(), elements.end());
        impl.insertAll(items);
        const size_t k = impl.getNumberOfSimpleTables(), e = impl.getEachSimpleTableSize();
        const size_t K = impl.getNumberOfCuckooHashFunctions(), b = impl.getEachBinSize(), E = impl.getEachCuckooTableSize();
        hierarchicalCuckooTable.assign(k, vector<CuckooHashTable>(e));
        for (size_t i = 0; i < k; i++)
            for (size_t j = 0; j < e; j++) {
                auto& t = hierarchicalCuckooTable[i][j].cuckooTable;
                t.assign(K, vector<vector<biginteger>>(b, vector<biginteger>(E)));
                for (size_t hf = 0; hf < K; hf++)
                    for (size_t bin = 0; bin < b; bin++)
                        for (size_t pos = 0; pos < E; pos++) t[hf][bin][pos] = impl.hierarchicalCuckooTable[i][j].cell((unsigned)hf, bin, pos);
            }
    }
    size_t getNumberOfSimpleTables() { return hierarchicalCuckooTable.size(); }
    uint64_t getEachSimpleTableSize() { return impl.getEachSimpleTableSize(); }
    uint64_t getEachCuckooTableSize() { return impl.getEachCuckooTableSize(); }
    uint64_t getServerStashSize() { return impl.getServerStashSize(); }
    uint getNumberOfCuckooHashFunctions() { return impl.getNumberOfCuckooHashFunctions(); }
    uint getNumberOfCuckooTables() { return impl.getNumberOfCuckooHashFunctions(); }
    uint64_t getEachBinSize() { return impl.getEachBinSize(); }
    bool hasSimpleMultiTables() { return impl.hasSimpleMultiTables(); }
    bool hasCuckooMultiTables() { return impl.hasCuckooMultiTables(); }
};

// CPU-only check of the constructor bookkeeping of psi::FHEHIPPIE / psi::FHEHIPPIECollection (host/FHEHIPPIE.hpp), the
// mirror of the reference's FHEHIPPIE.cpp:9-59: the two std::invalid_argument cases with the reference's messages, the
// bin permutation (whole rows, the same for every hash function), the trailing 1 of every plainVec, masks in [1, t - 1],
// the result permutation.  No device is touched: run() is never called.  Exit code 0 on success.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "FHEHIPPIE.hpp"

using namespace psi;

struct Probe : FHEHIPPIE {  // exposes what the constructor recorded
    using FHEHIPPIE::FHEHIPPIE;
    const std::vector<int64_t>& s() const { return slots; }
    const std::vector<int64_t>& m() const { return maskSlots; }
    const std::vector<unsigned>& perm() const { return permutationVector; }
};

static int fail(const char* what) {
    std::fprintf(stderr, "FAILED: %s\n", what);
    return 1;
}

int main() {
    const uint64_t t = 4296540161ULL;
    psi_params params;
    if (psi_params_generate(1024, t, 3, 0, &params) != PSI_OK) return fail("psi_params_generate");
    CryptoContext cc{params, nullptr, nullptr};
    PublicKey pk;
    TabulationHashing hashfu;
    std::vector<item_t> elems(40);
    for (size_t i = 0; i < elems.size(); i++) elems[i] = 1000 + 7 * i;

    // error cases (FHEHIPPIE.cpp:13-20)
    try {
        CuckooHashTable notSquare(hashfu, 6, 2, 0, 0, true, 4);
        Probe p(cc, pk, notSquare, 1);
        return fail("non-square table accepted");
    } catch (const std::invalid_argument& e) {
        if (!std::strstr(e.what(), "size of a cuckoo bin has to be equal")) return fail("wrong message (bin size)");
    }
    try {
        CuckooHashTable withStash(hashfu, 5, 2, 0, 3, true, 5);
        withStash.stash.push_back(99);
        Probe p(cc, pk, withStash, 1);
        return fail("stash accepted");
    } catch (const std::invalid_argument& e) {
        if (!std::strstr(e.what(), "does not support a stash")) return fail("wrong message (stash)");
    }

    const unsigned K = 2;
    const uint64_t b = 5;
    CuckooHashTable ct(hashfu, b, K, 0, 0, true, b);
    ct.insertAll(elems);
    Probe pie(cc, pk, ct, /*seed*/ 42);
    const size_t ns = b + 1;
    if (pie.s().size() != K * b * ns || pie.m().size() != K * b) return fail("slot vector sizes");
    std::vector<size_t> order0;
    for (unsigned hf = 0; hf < K; hf++) {
        std::vector<size_t> order;
        for (uint64_t bin = 0; bin < b; bin++) {
            // find the row of the table that landed in slot row `bin`
            size_t found = b;
            for (uint64_t src = 0; src < b && found == b; src++) {
                bool same = true;
                for (uint64_t pos = 0; pos < b; pos++) same = same && pie.s()[(hf * b + bin) * ns + pos] == (int64_t)ct.cell(hf, src, pos);
                if (same) found = src;
            }
            if (pie.s()[(hf * b + bin) * ns + b] != 1) return fail("trailing 1 missing");
            order.push_back(found);
        }
        // rows may be indistinguishable when empty; compare only when all rows were identified uniquely
        std::set<size_t> uniq(order.begin(), order.end());
        if (uniq.size() == b && uniq.count(b) == 0) {
            if (order0.empty()) order0 = order;
            else if (order0 != order) return fail("bin permutation differs between hash functions");
        }
    }
    for (int64_t m : pie.m())
        if (m < 1 || (uint64_t)m >= t) return fail("mask out of [1, t - 1]");
    std::vector<unsigned> perm = pie.perm();
    std::sort(perm.begin(), perm.end());
    for (unsigned i = 0; i < K; i++)
        if (perm[i] != i) return fail("result permutation is not a permutation");

    // explicit seeds reproduce, different seeds differ (tests only: the default draws from std::random_device)
    Probe again(cc, pk, ct, 42), other(cc, pk, ct, 43);
    if (again.m() != pie.m() || again.s() != pie.s()) return fail("seeded construction is not reproducible");
    if (other.m() == pie.m()) return fail("different seeds gave the same masks");

    // a collection demands one table shape
    FHEHIPPIECollection coll(cc, pk, 7);
    coll.addPIE(ct);
    try {
        CuckooHashTable smaller(hashfu, 4, K, 0, 0, true, 4);
        coll.addPIE(smaller);
        return fail("collection accepted a second table shape");
    } catch (const std::invalid_argument&) {
    }
    std::puts("FHEHIPPIE constructor bookkeeping ok");
    return 0;
}

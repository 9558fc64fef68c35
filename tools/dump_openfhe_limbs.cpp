// Pins limb-level parity against a real OpenFHE (parity ladder step 4).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE: OpenFHE is absent here, so the getter names below are those of
// OpenFHE 1.0.x as recalled (scheme/bfvrns/bfvrns-cryptoparameters.h, schemerns/rns-cryptoparameters.h) and
// must be checked against the installed headers.  Build next to an OpenFHE install:
//
//   g++ -std=c++17 -O2 tools/dump_openfhe_limbs.cpp -Iinclude -I$OPENFHE/include/openfhe{,/core,/pke,/binfhe} \
//       -L$OPENFHE/lib -lOPENFHEpke -lOPENFHEcore -fopenmp -o dump_openfhe_limbs
//   ./dump_openfhe_limbs tests/golden/openfhe_fixture.bin
//
// It creates the BFV context exactly as the reference client does (BatchedFHEPSIClient.cpp:58-91), fills a
// psi_params from OpenFHE's own tables, runs the circuit of BatchedFHEHIPPIE::run (BatchedFHEHIPPIE.cpp:88-129)
// with OpenFHE on a small deterministic database and query, and writes every operand and the results as raw
// limbs.  tests/test_openfhe_fixture.py replays the file through the oracle and through the CUDA path and
// requires bit-identical limbs; with the file absent that test is skipped and parity stays "unpinned".
//
// File layout (little endian): "PSIOFHE1", u64 sizeof(psi_params), psi_params, u64 K, b, E, nslots, then u64
// arrays evk_b[L][L][N], evk_a[L][L][N], pt[K][b][E][L][N], mask[b][L][N], idx[K][E][2][L][N], minus[2][L][N],
// result[b][2][L][N], then int64 arrays slots[K][b][E][nslots], mask_slots[b][nslots].
// Optional trailing section for the NON-batched path (FHEHIPPIE.cpp:61-77; pins the order inside EvalAutomorphism, the
// EvalSum index sequence and FindAutomorphismIndex2n): "PSINB001", u64 n_keys, u64 batch, u64 n_rot, u64
// key_index[n_keys], key_b[n_keys][L][L][N], key_a[n_keys][L][L][N] (every automorphism key of the context, BV),
// ct_in[2][L][N], then per rotation: i64 rotation, u64 automorphism index, ct EvalAtIndex(ct_in, rotation)[2][L][N];
// finally u64 n_sum, u64 sum_index[n_sum] (what EvalSumKeyGen(batch) generated) and ct EvalSum(ct_in, batch)[2][L][N].
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "openfhe.h"
#include "psi_b200.h"

using namespace lbcrypto;

static void put(FILE* f, const void* p, size_t n) {
    if (fwrite(p, 1, n, f) != n) { perror("fwrite"); exit(1); }
}
static void put_u64(FILE* f, uint64_t v) { put(f, &v, 8); }

// [limb][N] of one DCRTPoly, in the format it is in
static void put_poly(FILE* f, const DCRTPoly& d) {
    for (size_t l = 0; l < d.GetNumOfElements(); l++) {
        const auto& v = d.GetElementAtIndex(l).GetValues();
        for (size_t n = 0; n < v.GetLength(); n++) put_u64(f, v[n].ConvertToInt());
    }
}
static void put_ct(FILE* f, const Ciphertext<DCRTPoly>& ct) {  // [2][L][N], EVALUATION
    for (const auto& e : ct->GetElements()) put_poly(f, e);
}
static void put_pt_eval(FILE* f, const Plaintext& pt) {  // what EvalMult(ct, pt) multiplies with
    DCRTPoly d = pt->GetElement<DCRTPoly>();
    d.SetFormat(Format::EVALUATION);
    put_poly(f, d);
}
template <class V>
static void copy_vec(uint64_t* dst, const V& v) {
    for (size_t i = 0; i < v.size(); i++) dst[i] = v[i].ConvertToInt();
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s out.bin\n", argv[0]); return 2; }
    const uint64_t t = 4296540161ULL;  // --bitSize 32 (BatchedFHEPSIClient.cpp:27-30)
    const uint32_t K = 2, b = 2, E = 3, k = 2, e = 40, nslots = k * e;
    CCParams<CryptoContextBFVRNS> parameters;  // BatchedFHEPSIClient.cpp:72-78
    parameters.SetRingDim(16384);
    parameters.SetPlaintextModulus(t);
    parameters.SetMultiplicativeDepth(3);      // E < 500 (:46-49)
    parameters.SetBatchSize(nslots);
    parameters.SetSecurityLevel(SecurityLevel::HEStd_128_classic);
    CryptoContext<DCRTPoly> cc = GenCryptoContext(parameters);
    cc->Enable(PKE);
    cc->Enable(KEYSWITCH);
    cc->Enable(LEVELEDSHE);
    auto kp = cc->KeyGen();
    cc->EvalMultKeyGen(kp.secretKey);

    // ---- psi_params from the library's own tables (level 0) ---------------------------------------------
    const auto cp = std::dynamic_pointer_cast<CryptoParametersBFVRNS>(cc->GetCryptoParameters());
    if (cp->GetMultiplicationTechnique() != HPSPOVERQ || cp->GetKeySwitchTechnique() != BV) {
        fprintf(stderr, "context is not HPSPOVERQ + BV: the device path implements those defaults only\n");
        return 1;
    }
    psi_params P;
    memset(&P, 0, sizeof P);
    const auto& pq = cp->GetElementParams()->GetParams();
    const auto& pr = cp->GetParamsRl(0)->GetParams();
    P.N = cc->GetRingDimension();
    P.L = pq.size();
    P.Lp = pr.size();
    P.mult_technique = PSI_MULT_HPSPOVERQ;
    P.ks_technique = PSI_KS_BV;
    P.t = t;
    for (uint32_t i = 0; i < P.L; i++) {
        P.q[i] = pq[i]->GetModulus().ConvertToInt();
        P.psi_q[i] = pq[i]->GetRootOfUnity().ConvertToInt();
    }
    for (uint32_t j = 0; j < P.Lp; j++) {
        P.p[j] = pr[j]->GetModulus().ConvertToInt();
        P.psi_p[j] = pr[j]->GetRootOfUnity().ConvertToInt();
    }
    P.psi_t = RootOfUnity<NativeInteger>(2 * P.N, NativeInteger(t)).ConvertToInt();  // as PackedEncoding::SetParams; if
    // the installed version caches a different generator, read it from PackedEncoding's tables instead
    copy_vec(P.QHatInvModq, cp->GetQlHatInvModq(0));
    for (uint32_t j = 0; j < P.Lp; j++) copy_vec(P.QHatModp[j], cp->GetQlHatModr(0)[j]);      // verify [j][i] orientation
    for (uint32_t a = 0; a <= P.L; a++) copy_vec(P.alphaQModp[a], cp->GetalphaQlModr(0)[a]);
    for (uint32_t i = 0; i < P.L; i++) P.qInv[i] = cp->GetqlInv(0)[i];
    copy_vec(P.negPQHatInvModq, cp->GetNegRlQHatInvModq(0));
    for (uint32_t i = 0; i < P.L; i++) copy_vec(P.qInvModp[i], cp->GetqInvModr()[i]);
    copy_vec(P.PHatInvModp, cp->GetRlHatInvModr(0));
    for (uint32_t i = 0; i < P.L; i++) copy_vec(P.PHatModq[i], cp->GetRlHatModq(0)[i]);
    for (uint32_t a = 0; a <= P.Lp; a++) copy_vec(P.alphaPModq[a], cp->GetalphaRlModq(0)[a]);
    for (uint32_t j = 0; j < P.Lp; j++) P.pInv[j] = cp->GetrlInv()[j];
    for (uint32_t i = 0; i < P.L; i++) copy_vec(P.tQSHatInvModsDivsModq[i], cp->GettQlSlHatInvModsDivsModq(0)[i]);
    for (uint32_t j = 0; j < P.Lp; j++) P.tQSHatInvModsDivsFrac[j] = cp->GettQlSlHatInvModsDivsFrac(0)[j];

    // ---- deterministic database and query -----------------------------------------------------------------
    uint64_t x = 88172645463325252ULL;  // xorshift64
    auto next = [&x]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    std::vector<std::vector<std::vector<std::vector<int64_t>>>> slots(K);
    std::vector<std::vector<int64_t>> mask_slots(b);
    std::vector<std::vector<std::vector<Plaintext>>> pt(K);
    std::vector<Plaintext> mask(b);
    for (uint32_t hf = 0; hf < K; hf++) {
        slots[hf].resize(b);
        pt[hf].resize(b);
        for (uint32_t bin = 0; bin < b; bin++)
            for (uint32_t pos = 0; pos < E; pos++) {
                std::vector<int64_t> v(nslots);
                for (auto& s : v) s = (int64_t)(next() % (1ULL << 32));          // server items (32-bit)
                slots[hf][bin].push_back(v);
                pt[hf][bin].push_back(cc->MakePackedPlaintext(v));               // BatchedFHEHIPPIE.cpp:68
            }
    }
    for (uint32_t bin = 0; bin < b; bin++) {
        mask_slots[bin].resize(nslots);
        for (auto& s : mask_slots[bin]) s = (int64_t)(next() % (t - 1) + 1);     // :77
        mask[bin] = cc->MakePackedPlaintext(mask_slots[bin]);                    // :81
    }
    // query: one-hot position per slot and -x (BatchedFHEPSIClient.cpp:128-148), secret-key encryption (:156,:166)
    std::vector<std::vector<Ciphertext<DCRTPoly>>> idx(K, std::vector<Ciphertext<DCRTPoly>>(E));
    std::vector<int64_t> minus_v(nslots);
    for (uint32_t s = 0; s < nslots; s++) minus_v[s] = -(int64_t)(next() % (1ULL << 32));
    for (uint32_t hf = 0; hf < K; hf++) {
        std::vector<uint32_t> hot(nslots);
        for (auto& h : hot) h = (uint32_t)(next() % E);
        for (uint32_t pos = 0; pos < E; pos++) {
            std::vector<int64_t> v(nslots);
            for (uint32_t s = 0; s < nslots; s++) v[s] = hot[s] == pos;
            idx[hf][pos] = cc->Encrypt(kp.secretKey, cc->MakePackedPlaintext(v));
        }
    }
    Ciphertext<DCRTPoly> minus = cc->Encrypt(kp.secretKey, cc->MakePackedPlaintext(minus_v));

    // ---- the reference circuit, verbatim (BatchedFHEHIPPIE.cpp:91-128) --------------------------------------
    std::vector<Ciphertext<DCRTPoly>> result(b);
    for (uint32_t bin = 0; bin < b; bin++) {
        Ciphertext<DCRTPoly> multipliedResult;
        for (uint32_t hf = 0; hf < K; hf++) {
            Ciphertext<DCRTPoly> innerProductResult;
            for (uint32_t pos = 0; pos < E; pos++) {
                if (pos == 0)
                    innerProductResult = cc->EvalMult(idx[hf][pos], pt[hf][bin][pos]);
                else
                    innerProductResult = cc->EvalAdd(innerProductResult, cc->EvalMult(idx[hf][pos], pt[hf][bin][pos]));
            }
            innerProductResult = cc->EvalAdd(innerProductResult, minus);
            multipliedResult = hf == 0 ? innerProductResult : cc->EvalMult(multipliedResult, innerProductResult);
        }
        result[bin] = cc->EvalMult(multipliedResult, mask[bin]);
    }

    // ---- dump -------------------------------------------------------------------------------------------------
    FILE* f = fopen(argv[1], "wb");
    if (!f) { perror(argv[1]); return 1; }
    put(f, "PSIOFHE1", 8);
    put_u64(f, sizeof(psi_params));
    put(f, &P, sizeof P);
    put_u64(f, K); put_u64(f, b); put_u64(f, E); put_u64(f, nslots);
    const auto& evk = cc->GetEvalMultKeyVector(kp.secretKey->GetKeyTag())[0];  // BV: L digits
    for (const auto& d : evk->GetBVector()) put_poly(f, d);
    for (const auto& d : evk->GetAVector()) put_poly(f, d);
    for (uint32_t hf = 0; hf < K; hf++)
        for (uint32_t bin = 0; bin < b; bin++)
            for (uint32_t pos = 0; pos < E; pos++) put_pt_eval(f, pt[hf][bin][pos]);
    for (uint32_t bin = 0; bin < b; bin++) put_pt_eval(f, mask[bin]);
    for (uint32_t hf = 0; hf < K; hf++)
        for (uint32_t pos = 0; pos < E; pos++) put_ct(f, idx[hf][pos]);
    put_ct(f, minus);
    for (uint32_t bin = 0; bin < b; bin++) put_ct(f, result[bin]);
    for (uint32_t hf = 0; hf < K; hf++)
        for (uint32_t bin = 0; bin < b; bin++)
            for (uint32_t pos = 0; pos < E; pos++) put(f, slots[hf][bin][pos].data(), 8 * nslots);
    for (uint32_t bin = 0; bin < b; bin++) put(f, mask_slots[bin].data(), 8 * nslots);
    // ---- non-batched section: EvalSum keys for `batch` slots, rotation keys -1 .. -3 (SimpleFHEPSIClient.cpp:79-90) ----
    {
        const uint32_t batch = 5;
        const std::vector<int32_t> rotations = {-1, -2, -3};
        // the indices EvalSumKeyGen adds are read off the key map: taken before the rotation keys are generated
        cc->EvalSumKeyGen(kp.secretKey, kp.publicKey);
        std::vector<usint> sumIndex;
        for (const auto& kv : cc->GetEvalAutomorphismKeyMap(kp.secretKey->GetKeyTag())) sumIndex.push_back(kv.first);
        cc->EvalRotateKeyGen(kp.secretKey, rotations, kp.publicKey);
        const auto& keyMap = cc->GetEvalAutomorphismKeyMap(kp.secretKey->GetKeyTag());
        std::vector<int64_t> v(nslots);
        for (uint32_t i = 0; i < nslots; i++) v[i] = (int64_t)(i + 1);
        auto ctIn = cc->Encrypt(kp.secretKey, cc->MakePackedPlaintext(v));
        put(f, "PSINB001", 8);
        put_u64(f, keyMap.size()); put_u64(f, batch); put_u64(f, rotations.size());
        for (const auto& kv : keyMap) put_u64(f, kv.first);
        for (const auto& kv : keyMap)
            for (const auto& d : kv.second->GetBVector()) put_poly(f, d);
        for (const auto& kv : keyMap)
            for (const auto& d : kv.second->GetAVector()) put_poly(f, d);
        put_ct(f, ctIn);
        for (int32_t r : rotations) {
            int64_t r64 = r;
            put(f, &r64, 8);
            put_u64(f, FindAutomorphismIndex2n(r, 2 * P.N));  // core/math/nbtheory.h
            put_ct(f, cc->EvalAtIndex(ctIn, r));
        }
        put_u64(f, sumIndex.size());
        for (usint g : sumIndex) put_u64(f, g);
        put_ct(f, cc->EvalSum(ctIn, batch));
    }
    fclose(f);
    printf("wrote %s: N=%u L=%u Lp=%u K=%u b=%u E=%u nslots=%u\n", argv[1], P.N, P.L, P.Lp, K, b, E, nslots);
    return 0;
}

/*
 * psi_b200.h — C ABI of the B200-native BatchedFHEPIE server evaluation.
 *
 * This is the drop-in boundary for ONE hot path of SAP/nested-hashing-psi:
 *   BatchedFHEHIPPIE::run()          src/Common/Crypto/PrivateIndexedEqualityCheck/BatchedFHEHIPPIE.cpp:88-129
 * and for the data the path consumes:
 *   BatchedFHEHIPPIE ctor            BatchedFHEHIPPIE.cpp:9-86   (plaintext DB + random masks)
 *   setIndex / setMinusCompareElement BatchedFHEHIPPIE.hpp:40-48 (the encrypted query)
 *   getResultList                    BatchedFHEHIPPIE.hpp:35-38  (b result ciphertexts)
 *
 * The reference performs every homomorphic operation through OpenFHE
 * (cryptoContext->EvalMult / EvalAdd, BatchedFHEHIPPIE.cpp:108,112,113,116,123,126).
 * This library replaces exactly those calls with sm_100a kernels.  All entry
 * points are extern "C", take plain pointers and sizes, return an int status
 * (0 = ok) and never throw across the ABI.  Host buffers are owned by the
 * caller, device buffers by the library.  There is no CPU fallback: every
 * compute entry point fails with PSI_ERR_NO_DEVICE when no CUDA device exists.
 *
 * Layouts (u64 = uint64_t, all residues canonical in [0, q_l)):
 *   poly        [L][N]              limb-major, N contiguous
 *   ciphertext  [2][L][N]           (c0, c1), Format::EVALUATION (negacyclic NTT,
 *                                   natural-order input -> bit-reversed output,
 *                                   psi = minimal primitive 2N-th root mod q_l)
 *   plaintext   [L][N]              EVALUATION
 *   index       [K][E][2][L][N]     indexMatrix[hf][pos]         (BatchedFHEHIPPIE.hpp:25)
 *   minus       [2][L][N]           minusCompareElement          (BatchedFHEHIPPIE.hpp:26)
 *   pt DB       [K][b][E][L][N]     vectorizedHCT[hf][bin][pos]  (BatchedFHEHIPPIE.hpp:23)
 *   masks       [b][L][N]           preCalcRandomMask[bin]       (BatchedFHEHIPPIE.hpp:27)
 *   results     [b][2][L][N]        resultList[bin]              (BatchedFHEHIPPIE.hpp:24)
 *   relin key   [L][L][N] x 2       BV digit i, limb k; "b" then "a" vector
 */
#ifndef PSI_B200_H
#define PSI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSI_MAX_LIMBS 8

enum psi_status {
    PSI_OK = 0,
    PSI_ERR_INVALID = 1,   /* bad argument (mirrors std::invalid_argument in the reference) */
    PSI_ERR_NO_DEVICE = 2, /* CUDA device / driver missing: the product has no CPU path */
    PSI_ERR_CUDA = 3,      /* a CUDA call failed; see psi_last_error() */
    PSI_ERR_STATE = 4      /* call order violated (e.g. run before query_set) */
};

/* OpenFHE MultiplicationTechnique / KeySwitchTechnique of the BFV context.  HPSPOVERQ + BV (digit size 0) are the
 * BFVrns defaults of the 1.0.x line as recalled and run through the fused kernels; HPS and HYBRID are the other
 * variants a context may carry (risk register) and run through the unfused kernels. */
enum { PSI_MULT_HPS = 0, PSI_MULT_HPSPOVERQ = 1 };
enum { PSI_KS_BV = 0, PSI_KS_HYBRID = 1 };
/* How the host library's compiler evaluates the double sums nu += x * inv of SwitchCRTBasis / ScaleAndRound:
 * separate multiply and add (x86-64 without -march flags: no FMA available; default) or fused multiply-add
 * (-march=native builds, aarch64: GCC contracts by default). */
enum { PSI_FP_SEPARATE = 0, PSI_FP_FMA = 1 };

/*
 * Everything the device needs to know about the BFV-RNS context.  When linked
 * next to a real OpenFHE the adapter fills this from CryptoParametersBFVRNS
 * getters (INTEGRATION.md lists each getter), so the double tables are
 * bit-identical to the host library's; psi_params_generate() builds the same
 * tables from (N, t, depth) for stand-alone use.
 * q = moduli of Q (ciphertext basis), p = moduli of the auxiliary basis P used
 * by EvalMult(ct,ct); S = Q*P.
 */
typedef struct psi_params {
    uint32_t N;              /* ring dimension, power of two, 1024..16384 */
    uint32_t L;              /* sizeQ */
    uint32_t Lp;             /* sizeP */
    uint32_t mult_technique; /* PSI_MULT_* */
    uint32_t ks_technique;   /* PSI_KS_*  (BV, digit size 0) */
    uint32_t reserved;
    uint64_t t;              /* plaintext modulus, t = 1 mod 2N */
    uint64_t q[PSI_MAX_LIMBS];
    uint64_t p[PSI_MAX_LIMBS];
    uint64_t psi_q[PSI_MAX_LIMBS]; /* primitive 2N-th root of unity mod q_i */
    uint64_t psi_p[PSI_MAX_LIMBS];
    uint64_t psi_t;                /* primitive 2N-th root mod t (packed encoding) */

    /* DCRTPoly::ExpandCRTBasis (exact Q -> P, first operand of EvalMult) */
    uint64_t QHatInvModq[PSI_MAX_LIMBS];               /* [(Q/q_i)^-1]_{q_i} */
    uint64_t QHatModp[PSI_MAX_LIMBS][PSI_MAX_LIMBS];   /* [j][i] = (Q/q_i) mod p_j */
    uint64_t alphaQModp[PSI_MAX_LIMBS + 1][PSI_MAX_LIMBS]; /* [a][j] = a*Q mod p_j */
    double qInv[PSI_MAX_LIMBS];                        /* 1.0 / (double) q_i */

    /* DCRTPoly::FastExpandCRTBasisPloverQ (second operand, HPSPOVERQ) */
    uint64_t negPQHatInvModq[PSI_MAX_LIMBS];           /* [-P (Q/q_i)^-1]_{q_i} */
    uint64_t qInvModp[PSI_MAX_LIMBS][PSI_MAX_LIMBS];   /* [i][j] = q_i^-1 mod p_j */
    uint64_t PHatInvModp[PSI_MAX_LIMBS];               /* [(P/p_j)^-1]_{p_j} */
    uint64_t PHatModq[PSI_MAX_LIMBS][PSI_MAX_LIMBS];   /* [i][j] = (P/p_j) mod q_i */
    uint64_t alphaPModq[PSI_MAX_LIMBS + 1][PSI_MAX_LIMBS]; /* [a][i] = a*P mod q_i */
    double pInv[PSI_MAX_LIMBS];                        /* 1.0 / (double) p_j */

    /* DCRTPoly::ScaleAndRound by t/P with output basis Q (HPSPOVERQ) */
    uint64_t tQSHatInvModsDivsModq[PSI_MAX_LIMBS][PSI_MAX_LIMBS + 1]; /* [j][i<Lp], [j][Lp] */
    double tQSHatInvModsDivsFrac[PSI_MAX_LIMBS];

    /* ---- appended in 0.2; the layout of everything above is unchanged ---- */
    uint32_t fp_contract;  /* PSI_FP_* */
    uint32_t ks_num_parts; /* HYBRID: numPartQ, the number of digits Q is partitioned into (0 for BV) */
    uint32_t Lk;           /* HYBRID: sizeP of the key-switching basis (0 for BV) */
    uint32_t reserved2;
    uint64_t pk[PSI_MAX_LIMBS];     /* HYBRID: special primes of the key-switching basis */
    uint64_t psi_pk[PSI_MAX_LIMBS]; /* their 2N-th roots */
    /* DCRTPoly::ScaleAndRound by t/Q with output basis P (PSI_MULT_HPS); S = Q*P */
    uint64_t tPSHatInvModsDivsModp[PSI_MAX_LIMBS][PSI_MAX_LIMBS + 1]; /* [j][i<L], [j][L] */
    double tPSHatInvModsDivsFrac[PSI_MAX_LIMBS];
} psi_params;

typedef struct psi_ctx psi_ctx; /* opaque, one per device */

/* Stand-alone context parameters, restating OpenFHE's BFVrns parameter
 * generation (client side: BatchedFHEPSIClient.cpp:22-78 picks t, depth, ring
 * dimension).  L_override = 0 derives sizeQ from the noise estimate. */
int psi_params_generate(uint32_t N, uint64_t t, uint32_t mult_depth, uint32_t L_override,
                        psi_params* out);
/* The same with the variant choices of the context: mult_technique PSI_MULT_*, ks_technique PSI_KS_* (HYBRID: the
 * number of digits follows OpenFHE's rule for the depth, the special primes continue below the other two bases),
 * fp_contract PSI_FP_*. */
int psi_params_generate_ex(uint32_t N, uint64_t t, uint32_t mult_depth, uint32_t L_override, uint32_t mult_technique,
                           uint32_t ks_technique, uint32_t fp_contract, psi_params* out);

/* The same tables for moduli and roots that come from the host library (the OpenFHE adapter reads them from
 * ILDCRTParams: q / psi_q = ciphertext basis Q, p / psi_p = auxiliary basis of EvalMult, psi_t = the packed
 * encoding's 2N-th root mod t); no choice of primes involved. */
int psi_params_from_moduli(uint32_t N, uint64_t t, uint32_t L, const uint64_t* q, const uint64_t* psi_q, uint32_t Lp,
                           const uint64_t* p, const uint64_t* psi_p, uint64_t psi_t, psi_params* out);

/* Replaces the server's context deserialisation result as far as run() needs it
 * (BatchedFHEPSIServer.cpp:28). */
int psi_ctx_create(const psi_params* p, int device, psi_ctx** out);
int psi_ctx_destroy(psi_ctx* ctx);

/* Replaces cryptoContext->DeserializeEvalMultKey (BatchedFHEPSIServer.cpp:49) as
 * consumed by EvalMult(ct,ct) (BatchedFHEHIPPIE.cpp:123).  evk_b, evk_a, EVALUATION:
 *   BV      [L][L][N] u64 each (digit i, limb k)
 *   HYBRID  [ks_num_parts][L + Lk][N] u64 each (digit j, limbs of Q then of the key-switching basis) */
int psi_set_relin_key(psi_ctx* ctx, const uint64_t* evk_b, const uint64_t* evk_a);

/* Replaces the ctor's vectorizedHCT / preCalcRandomMask members
 * (BatchedFHEHIPPIE.cpp:37-82) with already encoded EVALUATION limbs. */
int psi_db_load_limbs(psi_ctx* ctx, uint32_t K, uint32_t b, uint32_t E, const uint64_t* pt_limbs,
                      const uint64_t* mask_limbs);

/* Same members, from raw slot values: performs MakePackedPlaintext
 * (BatchedFHEHIPPIE.cpp:68,81) plus the first-use SetFormat(EVALUATION) on the
 * device.  slots: [K][b][E][nslots] int64, mask_slots: [b][nslots] int64,
 * |v| < t; nslots <= N, remaining slots are zero. */
int psi_db_encode_slots(psi_ctx* ctx, uint32_t K, uint32_t b, uint32_t E, uint32_t nslots,
                        const int64_t* slots, const int64_t* mask_slots);

/* How MakePackedPlaintext lifts the packed coefficients from [0, t) to the limbs mod q_l before the first-use
 * SetFormat(EVALUATION) (BatchedFHEHIPPIE.cpp:68,81 -> OpenFHE PackedEncoding::Encode):
 *   PSI_ENCODE_LIFT_PLAIN    the coefficient itself in every limb (OpenFHE >= 1.0 as recalled; default)
 *   PSI_ENCODE_LIFT_CENTRED  c > t/2 -> q_l - (t - c), the signed representative (SURVEY.md Appendix A's reading)
 * Both decrypt identically; only limb parity with the host library depends on it.  Set before the psi_db_* calls
 * that encode (risk register). */
enum { PSI_ENCODE_LIFT_PLAIN = 0, PSI_ENCODE_LIFT_CENTRED = 1 };
int psi_set_encode_lift(psi_ctx* ctx, uint32_t mode);
/* Slot order of the packed encoding (PackedEncoding::SetParams_2n): slots N/2.. sit at exponents cofactor * 5^i with
 *   PSI_PACK_COFACTOR_3     cofactor 3 (OpenFHE 1.0.x as recalled; default)
 *   PSI_PACK_COFACTOR_CONJ  cofactor 2N - 1 (conjugation; the other reading)
 * Matters whenever the server encodes the database itself while the client encodes the query with the host
 * library: both sides must use the same slot order (risk register). */
enum { PSI_PACK_COFACTOR_3 = 0, PSI_PACK_COFACTOR_CONJ = 1 };
int psi_set_packing_cofactor(psi_ctx* ctx, uint32_t mode);

/* Sharded variants (single-process multi-device evaluation, psi_multi_* below, and one-process-per-GPU hosts):
 * the arrays are the FULL database of b_total bins in the layouts of the unsharded calls; bins
 * [bin_begin, bin_end) become this context's resident database (b_local = bin_end - bin_begin).  A sharded
 * psi_db_build_from_items needs an explicit shuffle seed shared by all shards (PSI_SEED_RANDOM is rejected). */
int psi_db_load_limbs_shard(psi_ctx* ctx, uint32_t K, uint32_t b_total, uint32_t bin_begin, uint32_t bin_end, uint32_t E,
                            const uint64_t* pt_limbs, const uint64_t* mask_limbs);
int psi_db_encode_slots_shard(psi_ctx* ctx, uint32_t K, uint32_t b_total, uint32_t bin_begin, uint32_t bin_end, uint32_t E,
                              uint32_t nslots, const int64_t* slots, const int64_t* mask_slots);
int psi_db_build_from_items_shard(psi_ctx* ctx, uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, uint64_t E, uint64_t b,
                                  uint64_t eviction_seed, const uint64_t* items, size_t n, uint64_t shuffle_seed,
                                  uint64_t mask_seed, uint32_t bin_begin, uint32_t bin_end);
/* One resident bin back to the host: pt [K][E][L][N], mask [L][N] (tests, bench.py's in-run parity check). */
int psi_db_get_bin_limbs(psi_ctx* ctx, uint32_t bin, uint64_t* pt_limbs, uint64_t* mask_limbs);

/* Copies the encoded DB back (tests / caching): pt [K][b][E][L][N], mask [b][L][N]. */
int psi_db_get_limbs(psi_ctx* ctx, uint64_t* pt_limbs, uint64_t* mask_limbs);

/* setIndex + setMinusCompareElement (BatchedFHEHIPPIE.hpp:40-48): asynchronous
 * H2D on `stream` (a cudaStream_t, NULL = legacy default stream). */
int psi_query_set(psi_ctx* ctx, const uint64_t* idx, const uint64_t* minus, void* stream);

/* The same in two steps, so that a server with a stream of queries can overlap the PCIe upload of the
 * next query with the evaluation of the current one: psi_query_upload only copies host -> device landing
 * buffers (no kernel reads them), psi_query_commit makes the oldest uploaded query the active one (re-tiling
 * kernel, must be ordered after the previous psi_run by the caller's streams/events).  There are TWO landing
 * buffers, used in turn: query i+1 may be uploaded before query i is committed (the copy engine never waits for
 * the compute stream); a third upload before a commit is PSI_ERR_STATE.  The caller orders upload i+2 after
 * commit i on the device (stream/event), since they touch the same buffer. */
int psi_query_upload(psi_ctx* ctx, const uint64_t* idx, const uint64_t* minus, void* stream);
int psi_query_commit(psi_ctx* ctx, void* stream);

/* The same upload from SEPARATE limb vectors, the layout OpenFHE holds after receiveIndexMatrix /
 * receiveEncryptedMinusElements (BatchedFHEPSIServer.cpp:114-141): idx_limbs = K*E*2*L pointers in
 * [hf][pos][comp][limb] order, minus_limbs = 2*L pointers, each to N words (pageable memory is fine).  Host
 * threads (psi_set_host_threads, default 8) gather the vectors into a pinned pool piece by piece while the pieces
 * already gathered are uploaded; the vectors may be freed on return.  psi_query_commit follows as usual.
 * psi_result_get_limbs is the mirror for getResultList / sendResult (:143-152): b*2*L pointers in
 * [bin][comp][limb] order; synchronous (the vectors are filled when it returns). */
int psi_query_upload_limbs(psi_ctx* ctx, const uint64_t* const* idx_limbs, const uint64_t* const* minus_limbs, void* stream);
int psi_result_get_limbs(psi_ctx* ctx, uint64_t* const* out_limbs, void* stream);
int psi_set_host_threads(psi_ctx* ctx, int n);

/* Device addresses of landing buffer `which` (0 or 1; idx [K][E][2][L][N], minus [2][L][N], u64), for hosts
 * that distribute one query over several GPUs themselves: every GPU receives 1/G of the index ciphertexts
 * over its own PCIe link and the slices are exchanged with an all-gather over NVLink ("query
 * replication = ... sliced H2D + NVLink all-gather") instead of G full uploads.  psi_query_next_landing tells
 * which buffer the next query goes to; the caller fills it on its stream, declares it with psi_query_uploaded
 * and calls psi_query_commit on a stream ordered after the fill. */
int psi_query_landing_ptr(psi_ctx* ctx, uint32_t which, void** idx, size_t* idx_bytes, void** minus, size_t* minus_bytes);
int psi_query_next_landing(psi_ctx* ctx, uint32_t* which);
int psi_query_uploaded(psi_ctx* ctx, uint32_t which);

/* BatchedFHEHIPPIE::run (BatchedFHEHIPPIE.cpp:88-129): enqueues all kernels on
 * `stream`; does not synchronise. */
int psi_run(psi_ctx* ctx, void* stream);

/* The two halves of run(), separately launchable so that each kernel family can be timed with CUDA
 * events and profiled on its own (bench.py roofline):
 *   PSI_PHASE_INNER_PRODUCT  BatchedFHEHIPPIE.cpp:101-116  (ct x pt multiply-accumulate + minus)
 *   PSI_PHASE_MULTIPLY_MASK  BatchedFHEHIPPIE.cpp:119-127  (ct x ct + relinearise, mask)
 * psi_run(ctx, s) == psi_run_phases(ctx, PSI_PHASE_ALL, s).  Phase 2 alone reuses the inner
 * products of the previous phase-1 launch. */
enum { PSI_PHASE_INNER_PRODUCT = 1, PSI_PHASE_MULTIPLY_MASK = 2, PSI_PHASE_ALL = 3 };
int psi_run_phases(psi_ctx* ctx, uint32_t phases, void* stream);

/* setMinusCompareElement + setIndex + run + getResultList for ONE query with the PCIe transfers overlapped inside
 * the query (the reference's server evaluates exactly one query per session, BatchedFHEPSIServer.cpp:99-108): the
 * index ciphertexts are uploaded in slices (hash function, position range) and each slice's share of the inner
 * products is accumulated as soon as it has landed; the ct x ct chain runs over the bins in groups and the results
 * of a group are downloaded while the next group is evaluated.  idx [K][E][2][L][N], minus [2][L][N], out
 * [b][2][L][N] are host buffers (pinned for full speed).  Asynchronous: synchronise `stream` before reading out. */
int psi_query_run_streamed(psi_ctx* ctx, const uint64_t* idx, const uint64_t* minus, uint64_t* out, void* stream);
/* The same from / into separately allocated limb vectors (layouts of psi_query_upload_limbs / psi_result_get_limbs):
 * every upload slice is gathered into the pinned pool by the host threads while the copy engine moves the previous
 * one, every download group is scattered into the result vectors as soon as it has arrived.  Synchronous: the
 * vectors are filled on return.  Replaces receiveIndexMatrix .. sendResult of one session
 * (BatchedFHEPSIServer.cpp:99-108,124-152) for a query held the way OpenFHE holds it. */
int psi_query_run_streamed_limbs(psi_ctx* ctx, const uint64_t* const* idx_limbs, const uint64_t* const* minus_limbs,
                                 uint64_t* const* out_limbs, void* stream);

/* getResultList (BatchedFHEHIPPIE.hpp:35-38): asynchronous D2H of [b][2][L][N]
 * on `stream`; the caller synchronises the stream before reading `out`.  Results are double-buffered on
 * the device: the call reads the buffer of the most recently ENQUEUED psi_run, and the next psi_run writes
 * the other one, so this copy may overlap the next evaluation. */
int psi_result_get(psi_ctx* ctx, uint64_t* out, void* stream);

int psi_stream_sync(void* stream);

/* Number of kernels one psi_run() launches for the loaded DB (bench.py's
 * gpu_launches claim is computed from this). */
int psi_run_launch_count(psi_ctx* ctx, uint32_t* out);

/* Device pointer of the result buffer (multi-GPU gather over NCCL works on
 * device memory). */
int psi_result_device_ptr(psi_ctx* ctx, void** out, size_t* bytes);

/* Kernel-level entry points used by the parity tests (each one kernel family):
 *  psi_debug_ntt:      in-place forward (inverse=0) / inverse NTT of n_polys
 *                      limb-polys; modulus index m: 0..L-1 = q, L..L+Lp-1 = p,
 *                      L+Lp = t.  data: [n_polys][N] host, moduli: [n_polys].
 *  psi_debug_mul_ctct: one EvalMult(ct,ct)+relinearise on host ciphertexts
 *                      ([2][L][N] each) -> out [2][L][N]. */
int psi_debug_ntt(psi_ctx* ctx, uint64_t* data, const uint32_t* moduli, uint32_t n_polys,
                  int inverse);
int psi_debug_mul_ctct(psi_ctx* ctx, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out);

/* Tuning knobs for measurements (tools/tune_shapes.py): mac_variant 0 = bin-block width of the inner-product kernel
 * chosen from the resident bins (default), 1 / 2 = force 2 / 4 bins per CTA (process-wide); phase2_groups 0 = number
 * of concurrent bin groups of the ct x ct chain chosen from the resident bins (default), 1..4 = force; -1 = leave. */
int psi_debug_set_tuning(psi_ctx* ctx, int mac_variant, int phase2_groups);

/* Integer-pipe peak micro-benchmark (the IMAD roofline
 * denominator is measured, not assumed).  Returns 32x32->64 multiply-adds per
 * second over the whole chip. */
int psi_bench_imad_peak(int device, double* mads_per_second);
/* kind 0: IMAD.WIDE.U32 per second (same as above); kind 1: 64-bit Harvey/Shoup lazy butterflies per
 * second with operands in registers (the NTT's compute ceiling); kinds 2, 3: two other codings of the same butterfly.
 * kinds 4, 5: the exchange of 16 u64 per thread between two radix passes through padded shared memory / through
 * __shfl_xor butterflies, in exchanges per second.
 * kind 6: the same butterfly with the Shoup quotient formed on the FP64 pipe (a probe: measured slower);
 * kind 7: its exactness pass, per_second then holds the number of lazy products outside [0, 4q) (must be 0).
 * Bits 4..7 of kind: resident 256-thread blocks per SM (0 = 8), to read the rate against occupancy. */
int psi_bench_pipe_peak(int device, int kind, double* per_second);

/* ------------------------------------------------------------------------------------------
 * Host-side objects around the path (offline phase; nothing here is inside run()).
 * They exist so that the PIE operator can be constructed exactly as the reference constructs it:
 *   BatchedFHEHIPPIE(cryptoContext, pk, HierarchicalCuckooHashTable&)   BatchedFHEHIPPIE.cpp:9-86
 * and so that tests / the bench can build the same inputs the reference's drivers build.
 * ------------------------------------------------------------------------------------------ */
typedef struct psi_hct psi_hct; /* TabulationHashing + HierarchicalCuckooHashTable */
typedef struct psi_pie psi_pie; /* BatchedFHEHIPPIE */

/* HierarchicalCuckooHashTable ctor (HierarchicalCuckooHashTable.cpp:16-53) over
 * TabulationHashing(hash_seed, k + K) (TabulationHashing.cpp:16-36).  The reference seeds the
 * eviction RNG of every inner table from std::random_device (CuckooHashTable.cpp:51-52);
 * eviction_seed pins it. */
int psi_hct_create(uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, uint64_t E, uint64_t b, uint64_t stash,
                   int simple_multi, int cuckoo_multi, uint64_t eviction_seed, psi_hct** out);
/* insertAll (HierarchicalCuckooHashTable.cpp:55-73); a failed insertion reports PSI_ERR_STATE
 * with the reference's message "(Blocked) Cuckoo hashing error" (CuckooHashTable.cpp:113). */
int psi_hct_insert_all(psi_hct* h, const uint64_t* items, size_t n);
/* cells: [k][e][K][b][E] = hierarchicalCuckooTable[outerHf][outerPos].cuckooTable[innerHf][bin][pos] */
int psi_hct_get_cells(psi_hct* h, uint64_t* cells);
int psi_hct_destroy(psi_hct* h);

/* The same table built ON THE DEVICE of ctx (one warp per inner cuckoo table; SURVEY 8f "next" #3):
 * bit-identical cells to psi_hct_create + psi_hct_insert_all with the same seeds (multi-tables, no stash). */
int psi_hct_build_device(psi_ctx* ctx, uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, uint64_t E, uint64_t b,
                         uint64_t eviction_seed, const uint64_t* items, size_t n, uint64_t* cells);
/* Whole constructor data path on the device: table build, bin shuffle, transposition, MakePackedPlaintext +
 * SetFormat(EVALUATION) (BatchedFHEHIPPIE.cpp:25-82, HierarchicalCuckooHashTable.cpp:55-73); the table never
 * visits the host.  Leaves ctx with the same database psi_pie_create produces from a host-built table with
 * the same seeds. */
int psi_db_build_from_items(psi_ctx* ctx, uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, uint64_t E, uint64_t b,
                            uint64_t eviction_seed, const uint64_t* items, size_t n, uint64_t shuffle_seed,
                            uint64_t mask_seed);

/* calculateHashIndex (HashUtils.cpp:34-37) for n items with TabulationHashing(hash_seed, n_hash_functions). */
int psi_hash_index(uint64_t hash_seed, uint32_t n_hash_functions, const uint64_t* items, size_t n, uint32_t hf,
                   uint32_t table_size, uint64_t* out);
/* The client's own cuckoo table, k tables x e positions x 1 item (BatchedFHEPSIClient.cpp:97-99,109);
 * cells: [k][e], 0 = empty. */
int psi_client_table(uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, const uint64_t* items, size_t n,
                     uint64_t eviction_seed, uint64_t* cells);
/* RandomDataInput (RandomDataInput.cpp:10-66): any output pointer may be NULL. */
int psi_random_data_input(size_t server_size, size_t client_size, size_t intersection_size, uint64_t seed,
                          uint64_t bit_size, uint64_t* server, uint64_t* client, uint64_t* intersection);

/* Seed value that asks for a fresh std::random_device seed.  shuffle_seed / mask_seed drive the bin shuffle and
 * the random masks of the PIE constructor (BatchedFHEHIPPIE.cpp:25-35,73-82); the masks are security-critical
 * (a known mask lets the client divide it out of the response), so production callers pass PSI_SEED_RANDOM,
 * which is what the reference does.  Explicit seeds are for reproducible tests only. */
#define PSI_SEED_RANDOM 0xFFFFFFFFFFFFFFFFull

/* BatchedFHEHIPPIE ctor: validates (stash, combined tables -> PSI_ERR_INVALID with the reference's
 * messages), shuffles the bin rows of hct IN PLACE, transposes, encodes on the device. */
int psi_pie_create(psi_ctx* ctx, const psi_params* params, psi_hct* hct, uint64_t shuffle_seed, uint64_t mask_seed,
                   int keep_slots, psi_pie** out);
int psi_pie_dims(psi_pie* p, uint32_t* K, uint32_t* b, uint32_t* E, uint32_t* nslots);
/* slots [K][b][E][nslots], mask_slots [b][nslots]; needs keep_slots at creation. */
int psi_pie_get_slots(psi_pie* p, int64_t* slots, int64_t* mask_slots);
/* setIndex + setMinusCompareElement / run / getResultList (BatchedFHEHIPPIE.hpp:33-48), synchronous. */
int psi_pie_set_query(psi_pie* p, const uint64_t* idx, const uint64_t* minus);
int psi_pie_run(psi_pie* p);
int psi_pie_get_result_list(psi_pie* p, uint64_t* out);
int psi_pie_destroy(psi_pie* p);

/* ------------------------------------------------------------------------------------------
 * Single-process multi-device evaluation.  The reference's server is ONE process whose single PIE object is
 * called once per session (BatchedFHEPSIServer.cpp:86,101-108); psi_multi keeps that shape on several GPUs:
 * one host thread, one psi_ctx per device, bins [b*d/n, b*(d+1)/n) resident on device d (the loop over bins at
 * BatchedFHEHIPPIE.cpp:91 carries nothing from bin to bin).  The query crosses PCIe once (1/n per device) and
 * the slices are exchanged device-to-device (cudaMemcpyPeerAsync, NVLink when peer access is available); every
 * device copies its own result ciphertexts straight into the caller's single [b][2][L][N] buffer.  All calls are
 * asynchronous with respect to the devices; psi_multi_sync waits for everything enqueued.  A device may be
 * listed more than once (several contexts on one GPU; the tests use this on single-GPU boxes).
 * ------------------------------------------------------------------------------------------ */
typedef struct psi_multi psi_multi;
int psi_multi_create(const psi_params* p, const int* devices, uint32_t n_devices, psi_multi** out);
int psi_multi_destroy(psi_multi* m);
int psi_multi_device_count(psi_multi* m, uint32_t* n);
/* bins of device `index` (after a psi_multi_db_* call) and its context (read-only use: debugging, tests) */
int psi_multi_bin_range(psi_multi* m, uint32_t index, uint32_t* bin_begin, uint32_t* bin_end);
int psi_multi_ctx(psi_multi* m, uint32_t index, psi_ctx** ctx);
int psi_multi_set_encode_lift(psi_multi* m, uint32_t mode);
int psi_multi_set_host_threads(psi_multi* m, int n); /* threads of the *_limbs staging copies (default 8) */
/* DeserializeEvalMultKey (BatchedFHEPSIServer.cpp:49): replicated on every device */
int psi_multi_set_relin_key(psi_multi* m, const uint64_t* evk_b, const uint64_t* evk_a);
/* the ctor's vectorizedHCT / preCalcRandomMask (BatchedFHEHIPPIE.cpp:37-82); same arrays as the psi_db_* calls,
 * FULL database, sharded inside */
int psi_multi_db_load_limbs(psi_multi* m, uint32_t K, uint32_t b, uint32_t E, const uint64_t* pt_limbs,
                            const uint64_t* mask_limbs);
int psi_multi_db_encode_slots(psi_multi* m, uint32_t K, uint32_t b, uint32_t E, uint32_t nslots, const int64_t* slots,
                              const int64_t* mask_slots);
int psi_multi_db_build_from_items(psi_multi* m, uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, uint64_t E,
                                  uint64_t b, uint64_t eviction_seed, const uint64_t* items, size_t n,
                                  uint64_t shuffle_seed, uint64_t mask_seed);
/* setIndex + setMinusCompareElement (BatchedFHEHIPPIE.hpp:40-48): idx [K][E][2][L][N], minus [2][L][N]; the host
 * buffers must stay valid until the copies have run (psi_multi_sync, or the next psi_multi_result_get + sync) */
int psi_multi_query_set(psi_multi* m, const uint64_t* idx, const uint64_t* minus);
/* The same from SEPARATE limb vectors, the layout an OpenFHE ciphertext holds (receiveIndexMatrix,
 * BatchedFHEPSIServer.cpp:124-141, leaves K*E ciphertexts of 2 DCRTPolys of L NativeVectors): idx_limbs has
 * K*E*2*L pointers in [hf][pos][comp][limb] order, minus_limbs 2*L, each to N words in pageable memory.  Host
 * threads copy them into a pinned pool piece by piece while the pieces already staged are being uploaded; the
 * vectors may be freed when the call returns. */
int psi_multi_query_set_limbs(psi_multi* m, const uint64_t* const* idx_limbs, const uint64_t* const* minus_limbs);
/* run() (BatchedFHEHIPPIE.cpp:88-129) on every device */
int psi_multi_run(psi_multi* m);
/* getResultList (BatchedFHEHIPPIE.hpp:35-38): out [b][2][L][N]; read it after psi_multi_sync */
int psi_multi_result_get(psi_multi* m, uint64_t* out);
/* the same into b*2*L separate limb vectors ([bin][comp][limb] order, N words each; sendResult serialises
 * ciphertext by ciphertext, BatchedFHEPSIServer.cpp:143-152); synchronous: the vectors are filled on return */
int psi_multi_result_get_limbs(psi_multi* m, uint64_t* const* out_limbs);
/* setIndex + setMinusCompareElement + run + getResultList of one session in one synchronous call, limb vectors in and
 * out: one device takes the streamed path (psi_query_run_streamed_limbs), several devices the three calls above. */
int psi_multi_query_run_limbs(psi_multi* m, const uint64_t* const* idx_limbs, const uint64_t* const* minus_limbs,
                              uint64_t* const* out_limbs);
int psi_multi_sync(psi_multi* m);
int psi_multi_run_launch_count(psi_multi* m, uint32_t* out);
/* BatchedFHEHIPPIE over a psi_multi (same constructor semantics as psi_pie_create) */
int psi_pie_create_multi(psi_multi* m, const psi_params* params, psi_hct* hct, uint64_t shuffle_seed, uint64_t mask_seed,
                         int keep_slots, psi_pie** out);

/* ------------------------------------------------------------------------------------------
 * Non-batched FHEHIPPIE: one private indexed equality check per outer cell of the server's nested
 * cuckoo table.  Replaces FHEHIPPIE (FHEHIPPIE.hpp:18-50; ctor FHEHIPPIE.cpp:9-59, run :61-77) and the loop
 * FHEHIPPIECollection::runAll makes over it (SimpleFHEPSIServer.cpp:126-160): the whole collection of PIEs is one
 * database and psi_nb_run evaluates a contiguous range of them in lock step.  BFV; key switching as the context says:
 * BV with digit size 0 (fused kernels) or HYBRID (one kernel per operation).
 * Per PIE: K hash functions, b bins of b positions (the ctor demands a square inner table, FHEHIPPIE.cpp:13-16).
 * Results come back per hash function in natural order; permutationVector (FHEHIPPIE.cpp:74) is the caller's.
 * ------------------------------------------------------------------------------------------ */
/* the automorphism indices run() needs: EvalSum of `batch_size` slots (EvalSum_2n), and EvalAtIndex(i)
 * (FindAutomorphismIndex2n).  Pure host arithmetic, no device. */
int psi_nb_eval_sum_indices(uint32_t N, uint32_t batch_size, uint64_t* out /*[<= 32]*/, uint32_t* n);
int psi_nb_rotation_index(uint32_t N, int64_t i, uint64_t* out);
/* DeserializeEvalSumKey + DeserializeEvalAutomorphismKey (SimpleFHEPSIServer.cpp:45-62): one BV key per automorphism
 * index, key_b / key_a [n_keys][L][L][N] EVALUATION ((*evalKey->GetBVector())[digit] / GetAVector, limb by limb);
 * HYBRID contexts: [n_keys][numPartQ][L + Lk][N], the layout of psi_set_relin_key. */
int psi_nb_set_automorphism_keys(psi_ctx* c, uint32_t n_keys, const uint64_t* auto_index, const uint64_t* key_b,
                                 const uint64_t* key_a);
/* the ctor's vectorizedCT / preCalcRandomMask of n_pie PIEs (FHEHIPPIE.cpp:25-58) as EVALUATION limbs:
 * pt [n_pie][K][b][L][N], mask [n_pie][K][L][N]; merge [L][N] = MakePackedPlaintext({1, 0, ...}) of EvalMerge */
int psi_nb_db_load_limbs(psi_ctx* c, uint32_t n_pie, uint32_t K, uint32_t b, const uint64_t* pt_limbs,
                         const uint64_t* mask_limbs, const uint64_t* merge_limbs);
/* the same from slot values, encoded on the device (MakePackedPlaintext, FHEHIPPIE.cpp:52,57):
 * slots [n_pie][K][b][nslots] (plainVec: the b positions of a bin, then 1; nslots = b + 1), mask_slots [n_pie][K][b] */
int psi_nb_db_encode_slots(psi_ctx* c, uint32_t n_pie, uint32_t K, uint32_t b, uint32_t nslots, const int64_t* slots,
                           const int64_t* mask_slots);
int psi_nb_db_get_limbs(psi_ctx* c, uint64_t* pt_limbs, uint64_t* mask_limbs, uint64_t* merge_limbs); /* any may be null */
/* setIndex + run + getResultList (FHEHIPPIE.hpp:39-49, FHEHIPPIE.cpp:61-77) of PIEs [pie_begin, pie_end):
 * idx, out: HOST [pie_end - pie_begin][K][2][L][N]; synchronous.  A missing key is PSI_ERR_STATE (OpenFHE throws). */
int psi_nb_run(psi_ctx* c, uint32_t pie_begin, uint32_t pie_end, const uint64_t* idx, uint64_t* out, void* stream);
int psi_nb_launch_count(psi_ctx* c, uint32_t* out); /* kernels launched by the last psi_nb_run */
/* The same collection over the devices of a psi_multi: PIEs sharded in contiguous blocks (the reference runs its
 * FHEHIPPIECollections on parallel threads, SimpleFHEPSIServer.cpp:126-160), keys replicated, one host thread per
 * device inside psi_multi_nb_run; idx / out cover ALL PIEs, [n_pie][K][2][L][N]. */
int psi_multi_nb_set_automorphism_keys(psi_multi* m, uint32_t n_keys, const uint64_t* auto_index, const uint64_t* key_b,
                                       const uint64_t* key_a);
int psi_multi_nb_db_encode_slots(psi_multi* m, uint32_t n_pie, uint32_t K, uint32_t b, uint32_t nslots, const int64_t* slots,
                                 const int64_t* mask_slots);
int psi_multi_nb_db_load_limbs(psi_multi* m, uint32_t n_pie, uint32_t K, uint32_t b, const uint64_t* pt_limbs,
                               const uint64_t* mask_limbs, const uint64_t* merge_limbs);
int psi_multi_nb_pie_range(psi_multi* m, uint32_t index, uint32_t* pie_begin, uint32_t* pie_end);
int psi_multi_nb_run(psi_multi* m, const uint64_t* idx, uint64_t* out);

/* number of CUDA devices visible to the process (hosts that build their device list without the CUDA headers) */
int psi_device_count(int* n);

const char* psi_last_error(void);
const char* psi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PSI_B200_H */

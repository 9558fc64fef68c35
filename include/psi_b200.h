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

/* OpenFHE MultiplicationTechnique / KeySwitchTechnique of the BFV context. */
enum { PSI_MULT_HPS = 0, PSI_MULT_HPSPOVERQ = 1 };
enum { PSI_KS_BV = 0 };

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
} psi_params;

typedef struct psi_ctx psi_ctx; /* opaque, one per device */

/* Stand-alone context parameters, restating OpenFHE's BFVrns parameter
 * generation (client side: BatchedFHEPSIClient.cpp:22-78 picks t, depth, ring
 * dimension).  L_override = 0 derives sizeQ from the noise estimate. */
int psi_params_generate(uint32_t N, uint64_t t, uint32_t mult_depth, uint32_t L_override,
                        psi_params* out);

/* Replaces the server's context deserialisation result as far as run() needs it
 * (BatchedFHEPSIServer.cpp:28). */
int psi_ctx_create(const psi_params* p, int device, psi_ctx** out);
int psi_ctx_destroy(psi_ctx* ctx);

/* Replaces cryptoContext->DeserializeEvalMultKey (BatchedFHEPSIServer.cpp:49) as
 * consumed by EvalMult(ct,ct) (BatchedFHEHIPPIE.cpp:123).  evk_b, evk_a:
 * [L][L][N] u64 each, EVALUATION. */
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

/* Integer-pipe peak micro-benchmark (the IMAD roofline
 * denominator is measured, not assumed).  Returns 32x32->64 multiply-adds per
 * second over the whole chip. */
int psi_bench_imad_peak(int device, double* mads_per_second);
/* kind 0: IMAD.WIDE.U32 per second (same as above); kind 1: 64-bit Harvey/Shoup lazy butterflies per
 * second with operands in registers (the NTT's compute ceiling). */
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

const char* psi_last_error(void);
const char* psi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PSI_B200_H */

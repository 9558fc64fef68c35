// psi_multi — the BatchedFHEPIE server evaluation on several GPUs driven by ONE host thread of ONE process,
// which is how the reference's server runs (a single BatchedFHEPSIServer object calls setMinusCompareElement /
// setIndex / run() / getResultList() once per session, BatchedFHEPSIServer.cpp:86,101-108).
//
// Bins are independent (loop at BatchedFHEHIPPIE.cpp:91), so the b bins are split into contiguous balanced
// blocks, one psi_ctx per device.  Nothing crosses devices while evaluating.  Around the evaluation:
//   query in    every device needs all K*E + 1 query ciphertexts.  The index ciphertexts cross PCIe ONCE:
//               device d receives the d-th 1/n of them from the host over its own link and the slices are
//               exchanged device-to-device (cudaMemcpyPeerAsync: NVLink / NVSwitch when peer access is on),
//               chunked so that the exchange of a chunk overlaps the upload of the next one.
//   response    every device copies its own block of result ciphertexts straight into the caller's single
//               [b][2][L][N] buffer (the only "gather": the next step is host serialisation anyway,
//               BatchedFHEPSIServer.cpp:143-152).
// The *_limbs entry points take the query / return the results as SEPARATE limb vectors (what an OpenFHE
// DCRTPoly holds: K*E*2*L vectors of N words in pageable memory): a few host threads copy them into a pinned
// staging pool piece by piece while the copy engines upload the pieces already staged.
#include <cuda_runtime.h>

#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "psi_b200.h"
#include "host_copy.hpp"
#include "../host/hashing.hpp"
#include "../host/psi_host_internal.hpp"

namespace {

using psi::set_error;

int cuda_rc(cudaError_t e, const char* what) {
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
        return set_error(PSI_ERR_NO_DEVICE, std::string(what) + ": no usable CUDA device; this library has no CPU path");
    return set_error(PSI_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define MCK(call)                                         \
    do {                                                  \
        cudaError_t e__ = (call);                         \
        if (e__ != cudaSuccess) return cuda_rc(e__, #call); \
    } while (0)
#define PCK(call)                    \
    do {                             \
        int rc__ = (call);           \
        if (rc__ != PSI_OK) return rc__; \
    } while (0)

constexpr uint32_t kExchangeChunks = 4;  // pieces per device slice of the query (upload / exchange overlap)

struct Dev {
    int device = 0;
    psi_ctx* ctx = nullptr;
    // s_in: host -> device uploads; s_x: peer pulls of the other devices' slices; s_run: commit + kernels;
    // s_out: device -> host downloads.  Query q+1 is uploaded and exchanged while query q is evaluated, and the
    // results of run r are downloaded while run r+1 is evaluated (the contexts double-buffer both).
    cudaStream_t s_in = nullptr, s_x = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t ev_up[kExchangeChunks] = {}, ev_pulled = nullptr, ev_query = nullptr, ev_done = nullptr, ev_d2h[2] = {};
    uint32_t bin_begin = 0, bin_end = 0;
};

}  // namespace

struct psi_multi {
    psi_params P{};
    std::vector<Dev> devs;
    uint32_t K = 0, b = 0, E = 0;
    bool have_db = false, have_evk = false, ran = false;
    uint64_t n_runs = 0;       // run r writes result buffer r & 1 of every context
    int host_threads = 8;      // host threads that fill / drain the pinned pools of the *_limbs entry points
    // pinned staging pools of the *_limbs entry points (allocated on first use)
    uint64_t* stage_in = nullptr;
    size_t stage_in_words = 0;
    uint64_t* stage_out = nullptr;
    size_t stage_out_words = 0;
    // non-batched collection (psi_multi_nb_*): PIEs [nb_begin[d], nb_begin[d + 1]) live on device d
    std::vector<uint32_t> nb_begin;
    uint32_t nb_K = 0, nb_b = 0;
    size_t ct_words() const { return (size_t)2 * P.L * P.N; }
    size_t idx_words() const { return (size_t)K * E * ct_words(); }
};

namespace {

int set_dims(psi_multi* m, uint32_t K, uint32_t b, uint32_t E) {
    if (K < 1 || b < 1 || E < 1) return set_error(PSI_ERR_INVALID, "K, b, E must be positive");
    const uint32_t n = (uint32_t)m->devs.size();
    if (b < n) return set_error(PSI_ERR_INVALID, "fewer bins than devices: create the psi_multi with at most b devices");
    m->K = K;
    m->b = b;
    m->E = E;
    for (uint32_t d = 0; d < n; d++) {
        m->devs[d].bin_begin = (uint32_t)(((uint64_t)d * b) / n);
        m->devs[d].bin_end = (uint32_t)(((uint64_t)(d + 1) * b) / n);
    }
    m->have_db = false;
    m->ran = false;
    return PSI_OK;
}

// [begin, end) of the index words device d uploads itself, piece c of kExchangeChunks
void slice(const psi_multi* m, uint32_t d, uint32_t c, size_t* begin, size_t* end) {
    const size_t W = m->idx_words(), n = m->devs.size();
    const size_t s0 = W * d / n, s1 = W * (d + 1) / n;
    *begin = s0 + (s1 - s0) * c / kExchangeChunks;
    *end = s0 + (s1 - s0) * (c + 1) / kExchangeChunks;
}

int ensure_pinned(uint64_t** p, size_t* have, size_t words) {
    if (*p && *have >= words) return PSI_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *have = 0;
    MCK(cudaHostAlloc((void**)p, words * sizeof(uint64_t), cudaHostAllocPortable));
    *have = words;
    return PSI_OK;
}

// The query is in (pinned or pageable) host memory at idx / minus; `staged_piece`, when given, is called before
// the upload of piece (d, c) and makes that piece available (the *_limbs path fills the pinned pool there).
template <typename BeforePiece>
int distribute_query(psi_multi* m, const uint64_t* idx, const uint64_t* minus, BeforePiece before_piece) {
    const uint32_t n = (uint32_t)m->devs.size();
    const size_t ctw = m->ct_words();
    std::vector<uint64_t*> land(n);
    std::vector<uint32_t> which(n);
    for (uint32_t d = 0; d < n; d++) {
        Dev& D = m->devs[d];
        void *pi, *pm;
        size_t ni, nm;
        PCK(psi_query_next_landing(D.ctx, &which[d]));
        PCK(psi_query_landing_ptr(D.ctx, which[d], &pi, &ni, &pm, &nm));
        land[d] = (uint64_t*)pi;
        MCK(cudaSetDevice(D.device));
        // the landing buffer is free once the previous commit that read it has run (same stream order as s_run)
        MCK(cudaStreamWaitEvent(D.s_in, D.ev_query, 0));
        MCK(cudaMemcpyAsync(pm, minus, ctw * sizeof(uint64_t), cudaMemcpyHostToDevice, D.s_in));
    }
    // piece-major order: while piece c of every slice is exchanged over NVLink, piece c + 1 is uploaded over PCIe
    for (uint32_t c = 0; c < kExchangeChunks; c++) {
        for (uint32_t d = 0; d < n; d++) {
            Dev& D = m->devs[d];
            size_t b0, b1;
            slice(m, d, c, &b0, &b1);
            before_piece(d, c, b0, b1);
            MCK(cudaSetDevice(D.device));
            if (b1 > b0) MCK(cudaMemcpyAsync(land[d] + b0, idx + b0, (b1 - b0) * sizeof(uint64_t), cudaMemcpyHostToDevice, D.s_in));
            MCK(cudaEventRecord(D.ev_up[c], D.s_in));
        }
        for (uint32_t j = 0; j < n; j++) {  // destination j pulls piece c of every other slice
            Dev& J = m->devs[j];
            MCK(cudaSetDevice(J.device));
            for (uint32_t o = 1; o < n; o++) {
                const uint32_t d = (j + o) % n;  // staggered: at any moment the n pulls read n different sources
                size_t b0, b1;
                slice(m, d, c, &b0, &b1);
                if (b1 == b0) continue;
                MCK(cudaStreamWaitEvent(J.s_x, m->devs[d].ev_up[c], 0));
                MCK(cudaMemcpyPeerAsync(land[j] + b0, J.device, land[d] + b0, m->devs[d].device, (b1 - b0) * sizeof(uint64_t), J.s_x));
            }
        }
    }
    for (uint32_t d = 0; d < n; d++) {
        Dev& D = m->devs[d];
        MCK(cudaSetDevice(D.device));
        MCK(cudaEventRecord(D.ev_pulled, D.s_x));
        MCK(cudaStreamWaitEvent(D.s_run, D.ev_up[kExchangeChunks - 1], 0));  // own slice + minus (s_in order)
        MCK(cudaStreamWaitEvent(D.s_run, D.ev_pulled, 0));                    // the other devices' slices
        PCK(psi_query_uploaded(D.ctx, which[d]));
        PCK(psi_query_commit(D.ctx, D.s_run));
        MCK(cudaEventRecord(D.ev_query, D.s_run));
    }
    // a peer may still be reading this device's landing buffer when the NEXT query is uploaded into the other
    // one; the two landing buffers alternate, and a buffer is rewritten only two queries later, after every
    // device's commit of this query (ev_query, waited for on s_in above) — which follows all pulls of this query
    // on the pulling device's s_run.  Cross-device: wait for every device's ev_query before the next upload.
    for (uint32_t d = 0; d < n; d++) {
        MCK(cudaSetDevice(m->devs[d].device));
        for (uint32_t j = 0; j < n; j++)
            if (j != d) MCK(cudaStreamWaitEvent(m->devs[d].s_in, m->devs[j].ev_query, 0));
    }
    return PSI_OK;
}

}  // namespace

extern "C" {

int psi_multi_create(const psi_params* p, const int* devices, uint32_t n_devices, psi_multi** out) {
    if (!p || !devices || !out) return set_error(PSI_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64) return set_error(PSI_ERR_INVALID, "device count must be in [1, 64]");
    psi_multi* m = new (std::nothrow) psi_multi();
    if (!m) return set_error(PSI_ERR_INVALID, "out of host memory");
    m->P = *p;
    m->devs.resize(n_devices);
    int rc = PSI_OK;
    for (uint32_t d = 0; d < n_devices && rc == PSI_OK; d++) {
        Dev& D = m->devs[d];
        D.device = devices[d];
        rc = psi_ctx_create(p, D.device, &D.ctx);
        if (rc != PSI_OK) break;
        cudaError_t e = cudaSetDevice(D.device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&D.s_in, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&D.s_x, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&D.s_run, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&D.s_out, cudaStreamNonBlocking);
        for (uint32_t c = 0; c < kExchangeChunks && e == cudaSuccess; c++) e = cudaEventCreateWithFlags(&D.ev_up[c], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&D.ev_query, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&D.ev_pulled, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&D.ev_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&D.ev_d2h[0], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&D.ev_d2h[1], cudaEventDisableTiming);
        if (e != cudaSuccess) rc = cuda_rc(e, "psi_multi_create");
    }
    // peer access between every pair of distinct devices: the exchange then runs over NVLink / NVSwitch; without it
    // cudaMemcpyPeerAsync still works (staged by the driver)
    for (uint32_t d = 0; d < n_devices && rc == PSI_OK; d++)
        for (uint32_t j = 0; j < n_devices; j++) {
            if (m->devs[d].device == m->devs[j].device) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, m->devs[d].device, m->devs[j].device) == cudaSuccess && can) {
                cudaSetDevice(m->devs[d].device);
                cudaError_t e = cudaDeviceEnablePeerAccess(m->devs[j].device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            }
        }
    if (rc != PSI_OK) {
        const std::string msg = psi_last_error();
        psi_multi_destroy(m);
        return set_error(rc, msg);
    }
    *out = m;
    return PSI_OK;
}

int psi_multi_destroy(psi_multi* m) {
    if (!m) return PSI_OK;
    for (Dev& D : m->devs) {
        if (D.ctx || D.s_in) cudaSetDevice(D.device);
        cudaStream_t* streams[] = {&D.s_in, &D.s_x, &D.s_run, &D.s_out};
        for (auto* st : streams)
            if (*st) cudaStreamSynchronize(*st);
        for (auto& e : D.ev_up)
            if (e) cudaEventDestroy(e);
        cudaEvent_t* events[] = {&D.ev_pulled, &D.ev_query, &D.ev_done, &D.ev_d2h[0], &D.ev_d2h[1]};
        for (auto* e : events)
            if (*e) cudaEventDestroy(*e);
        for (auto* st : streams)
            if (*st) cudaStreamDestroy(*st);
        if (D.ctx) psi_ctx_destroy(D.ctx);
    }
    if (m->stage_in) cudaFreeHost(m->stage_in);
    if (m->stage_out) cudaFreeHost(m->stage_out);
    delete m;
    return PSI_OK;
}

int psi_multi_device_count(psi_multi* m, uint32_t* n) {
    if (!m || !n) return set_error(PSI_ERR_INVALID, "null argument");
    *n = (uint32_t)m->devs.size();
    return PSI_OK;
}

int psi_multi_bin_range(psi_multi* m, uint32_t index, uint32_t* bin_begin, uint32_t* bin_end) {
    if (!m || !bin_begin || !bin_end || index >= m->devs.size()) return set_error(PSI_ERR_INVALID, "bad argument");
    if (!m->have_db) return set_error(PSI_ERR_STATE, "no database loaded");
    *bin_begin = m->devs[index].bin_begin;
    *bin_end = m->devs[index].bin_end;
    return PSI_OK;
}

int psi_multi_ctx(psi_multi* m, uint32_t index, psi_ctx** ctx) {
    if (!m || !ctx || index >= m->devs.size()) return set_error(PSI_ERR_INVALID, "bad argument");
    *ctx = m->devs[index].ctx;
    return PSI_OK;
}

int psi_multi_set_encode_lift(psi_multi* m, uint32_t mode) {
    if (!m) return set_error(PSI_ERR_INVALID, "null argument");
    for (Dev& D : m->devs) PCK(psi_set_encode_lift(D.ctx, mode));
    return PSI_OK;
}

int psi_multi_set_relin_key(psi_multi* m, const uint64_t* evk_b, const uint64_t* evk_a) {
    if (!m || !evk_b || !evk_a) return set_error(PSI_ERR_INVALID, "null argument");
    for (Dev& D : m->devs) PCK(psi_set_relin_key(D.ctx, evk_b, evk_a));
    m->have_evk = true;
    return PSI_OK;
}

int psi_multi_db_load_limbs(psi_multi* m, uint32_t K, uint32_t b, uint32_t E, const uint64_t* pt_limbs,
                            const uint64_t* mask_limbs) {
    if (!m || !pt_limbs || !mask_limbs) return set_error(PSI_ERR_INVALID, "null argument");
    PCK(set_dims(m, K, b, E));
    for (Dev& D : m->devs) PCK(psi_db_load_limbs_shard(D.ctx, K, b, D.bin_begin, D.bin_end, E, pt_limbs, mask_limbs));
    m->have_db = true;
    return PSI_OK;
}

int psi_multi_db_encode_slots(psi_multi* m, uint32_t K, uint32_t b, uint32_t E, uint32_t nslots, const int64_t* slots,
                              const int64_t* mask_slots) {
    if (!m || !slots || !mask_slots) return set_error(PSI_ERR_INVALID, "null argument");
    PCK(set_dims(m, K, b, E));
    for (Dev& D : m->devs) PCK(psi_db_encode_slots_shard(D.ctx, K, b, D.bin_begin, D.bin_end, E, nslots, slots, mask_slots));
    m->have_db = true;
    return PSI_OK;
}

int psi_multi_db_build_from_items(psi_multi* m, uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, uint64_t E, uint64_t b,
                                  uint64_t eviction_seed, const uint64_t* items, size_t n, uint64_t shuffle_seed,
                                  uint64_t mask_seed) {
    if (!m || (!items && n)) return set_error(PSI_ERR_INVALID, "null argument");
    PCK(set_dims(m, K, (uint32_t)b, (uint32_t)E));
    // one draw of the random seeds for the whole database: every device builds the same table, the same shuffle
    // and the same masks and keeps its own bins
    const uint64_t sh = psi::resolve_seed(shuffle_seed), mk = psi::resolve_seed(mask_seed);
    for (Dev& D : m->devs)
        PCK(psi_db_build_from_items_shard(D.ctx, hash_seed, k, e, K, E, b, eviction_seed, items, n, sh, mk, D.bin_begin, D.bin_end));
    m->have_db = true;
    return PSI_OK;
}

int psi_multi_query_set(psi_multi* m, const uint64_t* idx, const uint64_t* minus) {
    if (!m || !idx || !minus) return set_error(PSI_ERR_INVALID, "null argument");
    if (!m->have_db) return set_error(PSI_ERR_STATE, "load the database before the query (it fixes K and E)");
    return distribute_query(m, idx, minus, [](uint32_t, uint32_t, size_t, size_t) {});
}

int psi_multi_query_set_limbs(psi_multi* m, const uint64_t* const* idx_limbs, const uint64_t* const* minus_limbs) {
    if (!m || !idx_limbs || !minus_limbs) return set_error(PSI_ERR_INVALID, "null argument");
    if (!m->have_db) return set_error(PSI_ERR_STATE, "load the database before the query (it fixes K and E)");
    const size_t N = m->P.N, ctw = m->ct_words(), W = m->idx_words();
    PCK(ensure_pinned(&m->stage_in, &m->stage_in_words, W + ctw));
    // the previous query's uploads read the pool: they have finished once every s_in is idle
    for (Dev& D : m->devs) {
        MCK(cudaSetDevice(D.device));
        MCK(cudaStreamSynchronize(D.s_in));
    }
    uint64_t* pool = m->stage_in;
    for (size_t v = 0; v < ctw / N; v++) std::memcpy(pool + W + v * N, minus_limbs[v], N * sizeof(uint64_t));
    // limb vector v covers pool words [v*N, (v+1)*N); a piece may start / end inside a vector
    auto fill = [&](uint32_t, uint32_t, size_t b0, size_t b1) {
        if (b1 <= b0) return;
        const long v0 = (long)(b0 / N), v1 = (long)((b1 + N - 1) / N);
        const int nt = m->host_threads;
        (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt) if (v1 - v0 >= 16)
        for (long v = v0; v < v1; v++) {
            const size_t lo = (size_t)v * N < b0 ? b0 : (size_t)v * N;
            const size_t hi = (size_t)(v + 1) * N > b1 ? b1 : (size_t)(v + 1) * N;
            psi::copy_limb_vector(pool + lo, idx_limbs[v] + (lo - (size_t)v * N), hi - lo);
        }
    };
    return distribute_query(m, pool, pool + W, fill);
}

int psi_multi_run(psi_multi* m) {
    if (!m) return set_error(PSI_ERR_INVALID, "null argument");
    if (!m->have_db) return set_error(PSI_ERR_STATE, "run() needs a database and a query");
    for (Dev& D : m->devs) {
        MCK(cudaSetDevice(D.device));
        // run r writes the result buffer that the download of run r-2 read
        MCK(cudaStreamWaitEvent(D.s_run, D.ev_d2h[m->n_runs & 1], 0));
        PCK(psi_run(D.ctx, D.s_run));
        MCK(cudaEventRecord(D.ev_done, D.s_run));
    }
    m->n_runs++;
    m->ran = true;
    return PSI_OK;
}

int psi_multi_result_get(psi_multi* m, uint64_t* out) {
    if (!m || !out) return set_error(PSI_ERR_INVALID, "null argument");
    if (!m->ran) return set_error(PSI_ERR_STATE, "getResultList() before run()");
    const size_t ctw = m->ct_words();
    for (Dev& D : m->devs) {
        MCK(cudaSetDevice(D.device));
        MCK(cudaStreamWaitEvent(D.s_out, D.ev_done, 0));
        PCK(psi_result_get(D.ctx, out + (size_t)D.bin_begin * ctw, D.s_out));
        MCK(cudaEventRecord(D.ev_d2h[(m->n_runs - 1) & 1], D.s_out));
    }
    return PSI_OK;
}

int psi_multi_result_get_limbs(psi_multi* m, uint64_t* const* out_limbs) {
    if (!m || !out_limbs) return set_error(PSI_ERR_INVALID, "null argument");
    if (!m->ran) return set_error(PSI_ERR_STATE, "getResultList() before run()");
    const size_t N = m->P.N, ctw = m->ct_words();
    PCK(ensure_pinned(&m->stage_out, &m->stage_out_words, (size_t)m->b * ctw));
    PCK(psi_multi_result_get(m, m->stage_out));
    // device after device: scatter the block that has arrived while the other devices are still copying
    for (Dev& D : m->devs) {
        MCK(cudaSetDevice(D.device));
        MCK(cudaStreamSynchronize(D.s_out));
        const long v0 = (long)((size_t)D.bin_begin * ctw / N), v1 = (long)((size_t)D.bin_end * ctw / N);
        const int nt = m->host_threads;
        (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt) if (v1 - v0 >= 16)
        for (long v = v0; v < v1; v++) psi::copy_limb_vector(out_limbs[v], m->stage_out + (size_t)v * N, N);
    }
    return PSI_OK;
}

// One query from limb vectors to limb vectors, synchronous.  On ONE device this is psi_query_run_streamed_limbs (host
// gather, upload slices, evaluation, download groups and scatter overlapped inside the query); on several devices the
// three calls above in sequence (every device already moves only its share of the query and of the results).
int psi_multi_query_run_limbs(psi_multi* m, const uint64_t* const* idx_limbs, const uint64_t* const* minus_limbs,
                              uint64_t* const* out_limbs) {
    if (!m || !idx_limbs || !minus_limbs || !out_limbs) return set_error(PSI_ERR_INVALID, "null argument");
    if (!m->have_db) return set_error(PSI_ERR_STATE, "run() needs a database and a query");
    if (m->devs.size() == 1) {
        Dev& D = m->devs[0];
        MCK(cudaSetDevice(D.device));
        // earlier asynchronous work of this object (uploads, downloads of a previous query) is ordered before
        MCK(cudaStreamSynchronize(D.s_in));
        MCK(cudaStreamSynchronize(D.s_out));
        PCK(psi_set_host_threads(D.ctx, m->host_threads));
        PCK(psi_query_run_streamed_limbs(D.ctx, idx_limbs, minus_limbs, out_limbs, D.s_run));
        MCK(cudaEventRecord(D.ev_done, D.s_run));
        MCK(cudaStreamSynchronize(D.s_run));
        m->n_runs++;
        m->ran = true;
        return PSI_OK;
    }
    PCK(psi_multi_query_set_limbs(m, idx_limbs, minus_limbs));
    PCK(psi_multi_run(m));
    PCK(psi_multi_result_get_limbs(m, out_limbs));
    return psi_multi_sync(m);
}

int psi_multi_sync(psi_multi* m) {
    if (!m) return set_error(PSI_ERR_INVALID, "null argument");
    for (Dev& D : m->devs) {
        MCK(cudaSetDevice(D.device));
        MCK(cudaStreamSynchronize(D.s_in));
        MCK(cudaStreamSynchronize(D.s_x));
        MCK(cudaStreamSynchronize(D.s_run));
        MCK(cudaStreamSynchronize(D.s_out));
    }
    return PSI_OK;
}

int psi_multi_set_host_threads(psi_multi* m, int n) {
    if (!m || n < 1 || n > 256) return set_error(PSI_ERR_INVALID, "host thread count must be in [1, 256]");
    m->host_threads = n;
    return PSI_OK;
}

// ---- non-batched FHEHIPPIE collection over the devices (FHEHIPPIE.cpp; the reference runs its collections on
// parallel threads, SimpleFHEPSIServer.cpp:126-160 - here the PIEs are sharded over the devices in contiguous blocks,
// the automorphism keys are replicated, and psi_multi_nb_run drives every device from its own host thread) ----------
int psi_multi_nb_set_automorphism_keys(psi_multi* m, uint32_t n_keys, const uint64_t* auto_index, const uint64_t* key_b,
                                       const uint64_t* key_a) {
    if (!m) return set_error(PSI_ERR_INVALID, "null argument");
    for (Dev& D : m->devs) PCK(psi_nb_set_automorphism_keys(D.ctx, n_keys, auto_index, key_b, key_a));
    return PSI_OK;
}

static int nb_shard(psi_multi* m, uint32_t n_pie, uint32_t K, uint32_t b) {
    const uint32_t n = (uint32_t)m->devs.size();
    if (n_pie < n) return set_error(PSI_ERR_INVALID, "fewer PIEs than devices: create the psi_multi with at most n_pie devices");
    m->nb_begin.resize(n + 1);
    for (uint32_t d = 0; d <= n; d++) m->nb_begin[d] = (uint32_t)(((uint64_t)d * n_pie) / n);
    m->nb_K = K;
    m->nb_b = b;
    return PSI_OK;
}

int psi_multi_nb_db_encode_slots(psi_multi* m, uint32_t n_pie, uint32_t K, uint32_t b, uint32_t nslots, const int64_t* slots,
                                 const int64_t* mask_slots) {
    if (!m || !slots || !mask_slots) return set_error(PSI_ERR_INVALID, "null argument");
    PCK(nb_shard(m, n_pie, K, b));
    for (size_t d = 0; d < m->devs.size(); d++) {
        const uint32_t p0 = m->nb_begin[d], p1 = m->nb_begin[d + 1];
        PCK(psi_nb_db_encode_slots(m->devs[d].ctx, p1 - p0, K, b, nslots, slots + (size_t)p0 * K * b * nslots,
                                   mask_slots + (size_t)p0 * K * b));
    }
    return PSI_OK;
}

int psi_multi_nb_db_load_limbs(psi_multi* m, uint32_t n_pie, uint32_t K, uint32_t b, const uint64_t* pt_limbs,
                               const uint64_t* mask_limbs, const uint64_t* merge_limbs) {
    if (!m || !pt_limbs || !mask_limbs || !merge_limbs) return set_error(PSI_ERR_INVALID, "null argument");
    PCK(nb_shard(m, n_pie, K, b));
    const size_t LN = (size_t)m->P.L * m->P.N;
    for (size_t d = 0; d < m->devs.size(); d++) {
        const uint32_t p0 = m->nb_begin[d], p1 = m->nb_begin[d + 1];
        PCK(psi_nb_db_load_limbs(m->devs[d].ctx, p1 - p0, K, b, pt_limbs + (size_t)p0 * K * b * LN, mask_limbs + (size_t)p0 * K * LN,
                                 merge_limbs));
    }
    return PSI_OK;
}

int psi_multi_nb_pie_range(psi_multi* m, uint32_t index, uint32_t* pie_begin, uint32_t* pie_end) {
    if (!m || !pie_begin || !pie_end || index + 1 >= m->nb_begin.size()) return set_error(PSI_ERR_INVALID, "bad argument");
    *pie_begin = m->nb_begin[index];
    *pie_end = m->nb_begin[index + 1];
    return PSI_OK;
}

int psi_multi_nb_run(psi_multi* m, const uint64_t* idx, uint64_t* out) {
    if (!m || !idx || !out) return set_error(PSI_ERR_INVALID, "null argument");
    if (m->nb_begin.size() != m->devs.size() + 1) return set_error(PSI_ERR_STATE, "no non-batched database loaded");
    const size_t n = m->devs.size(), per_pie = (size_t)m->nb_K * m->ct_words();
    std::vector<int> rc(n, PSI_OK);
    std::vector<std::string> msg(n);
    std::vector<std::thread> workers;
    for (size_t d = 0; d < n; d++)
        workers.emplace_back([&, d] {  // psi_nb_run is synchronous: one host thread per device keeps them all busy
            const uint32_t p0 = m->nb_begin[d], p1 = m->nb_begin[d + 1];
            rc[d] = psi_nb_run(m->devs[d].ctx, 0, p1 - p0, idx + p0 * per_pie, out + p0 * per_pie, m->devs[d].s_run);
            if (rc[d] != PSI_OK) msg[d] = psi_last_error();  // the message is thread-local
        });
    for (auto& w : workers) w.join();
    for (size_t d = 0; d < n; d++)
        if (rc[d] != PSI_OK) return set_error(rc[d], msg[d]);
    return PSI_OK;
}

int psi_multi_run_launch_count(psi_multi* m, uint32_t* out) {
    if (!m || !out) return set_error(PSI_ERR_INVALID, "null argument");
    uint32_t total = 0;
    for (Dev& D : m->devs) {
        uint32_t n = 0;
        PCK(psi_run_launch_count(D.ctx, &n));
        total += n;
    }
    *out = total;
    return PSI_OK;
}

}  // extern "C"

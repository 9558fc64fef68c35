// Internal definition of the context behind the opaque psi_ctx handle of include/psi_b200.h, shared by the
// translation units that orchestrate kernels on it (psi_api.cu: the batched PIE; psi_nonbatched.cu: FHEHIPPIE).
#pragma once
#include <cuda_runtime.h>

#include <new>
#include <string>
#include <vector>

#include "psi_b200.h"
#include "psi_kernels.cuh"
#include "../host/psi_host_internal.hpp"

struct psi_ctx;

namespace psi {

// maps a CUDA error to a status + psi_last_error() message (PSI_ERR_NO_DEVICE when there is no usable device)
int cuda_fail(cudaError_t e, const char* what);

#define CK(call)                                        \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return ::psi::cuda_fail(e__, #call); \
    } while (0)

constexpr size_t kDbChunk = 256;  // plaintexts staged per re-tiling / encode step
constexpr uint32_t kMaxStreamSlices = 32;  // upload slices of psi_query_run_streamed

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }  // locals are freed on every exit path (CK returns early)
};

struct NbState;  // non-batched FHEHIPPIE state (psi_nonbatched.cu)
void nb_release(psi_ctx* c);
}  // namespace psi

struct psi_ctx {
    using u64 = psi::u64;
    template <typename T> using DevBuf = psi::DevBuf<T>;
    using DevTables = psi::DevTables;
    using KCtx = psi::KCtx;
    int device = 0;
    psi_params P{};
    uint32_t N = 0, logN = 0, L = 0, Lp = 0;
    DevTables* d_tab = nullptr;
    DevBuf<u64> twiddles;       // [(L+Lp+1)][4][N]
    DevBuf<u64> twiddles2;      // [(L+Lp+1)][2][N][2]: {w, ws} and {iw, iws} interleaved
    DevBuf<u64> twiddles_rows;  // [(L+Lp)][2][N/1024][1016][2]: row-stage twiddles packed per row tile
    DevBuf<uint32_t> to_crt;    // packed-encoding permutation
    DevBuf<u64> evk_b, evk_a;   // [L][L][N]
    DevBuf<u64> evk_bR, evk_aR; // the same times R = 2^64 mod q_k (Montgomery form for the fused relinearisation)
    DevBuf<u64> maskR;          // masks times R
    bool have_evk = false;
    // database
    uint32_t K = 0, b = 0, E = 0;
    DevBuf<u64> pt, mask;
    bool have_db = false;
    uint32_t encode_lift = PSI_ENCODE_LIFT_PLAIN;
    // query
    DevBuf<u64> idx, idx_in, minus;  // idx: tiled split-30; idx_in: two H2D landing buffers [2][K][E][2][L][N]
    DevBuf<u64> stage;               // chunk staging for the DB re-tiling
    bool have_query = false;
    // work
    DevBuf<u64> acc, coef, e1, e2, ten, res, dig, prod, out, out2;
    DevBuf<u64> minus_in;  // two H2D landing buffers of minusCompareElement [2][2][L][N]
    // results are double-buffered: run() i writes out[i & 1], so the D2H of query i can overlap run() i+1
    uint32_t out_cur = 0;
    // landing buffer n & 1 receives the n-th uploaded query; commits consume them in the same order, so the upload
    // of query i+1 never has to wait for the commit of query i
    uint32_t n_uploaded = 0, n_committed = 0;
    size_t idx_words() const { return (size_t)K * E * 2 * L * N; }
    u64* out_buf(uint32_t which) { return which ? out2.p : out.p; }
    bool ran = false;
    uint32_t launches_per_run = 0;
    // pinned staging pools of the *_limbs entry points (separate limb vectors <-> one DMA-able buffer)
    u64* pool_in = nullptr;
    u64* pool_out = nullptr;
    size_t pool_in_words = 0, pool_out_words = 0;
    cudaEvent_t ev_pool_in = nullptr;  // last upload that read pool_in
    int host_threads = 8;
    // phase 2 in bin groups on concurrent streams (tails of one group's kernels overlap the next group's heads);
    // 0 = choose from the number of resident bins
    uint32_t p2_groups = 0;
    cudaStream_t aux[7] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[7] = {};
    // psi_query_run_streamed: copy-in / copy-out streams and the events that order slices and bin groups
    cudaStream_t sq_in = nullptr, sq_out = nullptr;
    cudaStream_t sq_grp[4] = {};  // bin groups of the streamed query, descending priority
    cudaEvent_t ev_dl[4] = {};    // download of a group complete (limb-vector form: the host scatters it then)
    cudaEvent_t ev_sq_fork = nullptr;
    cudaEvent_t ev_slice[psi::kMaxStreamSlices] = {}, ev_group[4] = {}, ev_sq = nullptr;

    // run() as a CUDA graph: the launch set of one evaluation (inner product, bin groups forked over the auxiliary
    // streams, their joins) is captured once per (phases, result buffer, grouping, buffer addresses) and replayed, so a
    // query costs one graph launch instead of 11-25 kernel launches and event operations on the host
    struct RunGraph {
        uint64_t key = 0;
        cudaGraphExec_t exec = nullptr;
        uint32_t launches = 0;
    };
    std::vector<RunGraph> graphs;
    bool use_graph = true;
    uint32_t Lk = 0, ks_parts = 0;  // HYBRID key switching
    bool hybrid = false, hps = false;
    psi::NbState* nb = nullptr;
    KCtx k(cudaStream_t s) const { return KCtx{d_tab, N, logN, L, Lp, s, Lk, !hybrid && !hps}; }
};

namespace psi {
int ensure_device(psi_ctx* c);
// MakePackedPlaintext + SetFormat(EVALUATION) of n_pt slot vectors on the device (psi_api.cu); tiled_E = 0: dst flat [n_pt][L][N]
int encode_into(psi_ctx* c, size_t n_pt, uint32_t nslots, const int64_t* slots, u64* dst, uint32_t tiled_E, size_t p_base = 0);
}  // namespace psi

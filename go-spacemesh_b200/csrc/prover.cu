// prover.cu — POST proof generation scan (include/b200post_prove.h).
//
// K6 prove_scan_kernel streams 16-byte labels (H2D from the postdata files, double-buffered) through one
// AES-128 cipher per nonce group and appends (nonce, index) hits to a small list; the host keeps the per-nonce
// hit lists and stops when a nonce owns K2 of them.  Bandwidth view: 16 B in per label, ~nothing out; the
// kernel is far faster than PCIe/NVMe can feed it, so the design goal is simply to keep copies and compute
// overlapped.  Conventions: post-rs Prover8_56 from memory (ASSUMED, unpinned).
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200post_prove.h"
#include "../../include/b200post_k2pow.h"
#include "aes_device.cuh"
#include "engine.h"
#include "metrics.h"
#include "proof_common.h"

namespace b200post {
namespace {

struct Hit { uint32_t nonce; uint32_t pad; uint64_t index; };

// K6a: rk = per nonce group 11 round keys.  Every (label, group) costs one AES; ciphertext bytes below the
// difficulty MSB are hits, bytes EQUAL to it (1 in 256) need the nonce's "lazy" cipher: those are queued as
// (label offset, nonce) candidates and resolved densely by K6b — evaluating them in place would run a whole
// AES with one or two active lanes for most warps.
__global__ void __launch_bounds__(256) prove_scan_kernel(const uint4 *__restrict__ labels, uint64_t first_index, uint32_t count,
                                                         const uint4 *__restrict__ rk, uint32_t n_groups, uint32_t diff_msb,
                                                         const AesTables *__restrict__ tables, Hit *__restrict__ hits,
                                                         uint32_t hit_cap, uint32_t *__restrict__ n_hits,
                                                         uint2 *__restrict__ cands, uint32_t cand_cap, uint32_t *__restrict__ n_cands) {
    extern __shared__ uint32_t aes_sm[];
    aes_load_smem(aes_sm, tables);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t *tl = aes_sm + lane;
    const uint32_t msb4 = diff_msb * 0x01010101u;
    const uint32_t stride = gridDim.x * blockDim.x;
    // whole warps stay in the loop together (the candidate compaction below is warp-collective)
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < count; base += stride) {
        const uint32_t i = base + lane;
        const bool live = i < count;
        const uint4 label = live ? labels[i] : make_uint4(0, 0, 0, 0);
        for (uint32_t g = 0; g < n_groups; g++) {
            const uint4 out = aes128_encrypt(tl, rk + 11 * g, label);
            // per byte: 0xff where ciphertext byte <= MSB / == MSB
            const uint32_t le[4] = {__vcmpleu4(out.x, msb4), __vcmpleu4(out.y, msb4), __vcmpleu4(out.z, msb4), __vcmpleu4(out.w, msb4)};
            const uint32_t eq[4] = {__vcmpeq4(out.x, msb4), __vcmpeq4(out.y, msb4), __vcmpeq4(out.z, msb4), __vcmpeq4(out.w, msb4)};
            const bool any_le = live && (le[0] | le[1] | le[2] | le[3]);
            if (!__any_sync(0xffffffffu, any_le)) continue;
            uint32_t n_eq = 0;
            if (any_le) {
#pragma unroll
                for (int w = 0; w < 4; w++) n_eq += __popc(eq[w]) >> 3;
            }
            // warp-aggregated reservation of candidate slots
            uint32_t incl = n_eq;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += t; }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t slot = 0;
            if (total) {
                if (lane == 31) slot = atomicAdd(n_cands, total);
                slot = __shfl_sync(0xffffffffu, slot, 31) + incl - n_eq;
            }
            if (any_le) {
#pragma unroll 1
                for (uint32_t b = 0; b < 16; b++) {
                    const uint32_t bit = 0xffu << (8 * (b & 3));
                    if (!(le[b >> 2] & bit)) continue;
                    const uint32_t nonce = g * 16 + b;
                    if (eq[b >> 2] & bit) {
                        if (slot < cand_cap) cands[slot] = make_uint2(i, nonce);
                        slot++;
                    } else {
                        const uint32_t pos = atomicAdd(n_hits, 1u);
                        if (pos < hit_cap) hits[pos] = Hit{nonce, 0, first_index + i};
                    }
                }
            }
        }
    }
}

// K6b: one thread per candidate: the nonce's lazy cipher decides with the low 56 bits.
__global__ void __launch_bounds__(256) prove_lazy_kernel(const uint4 *__restrict__ labels, uint64_t first_index,
                                                         const uint2 *__restrict__ cands, const uint32_t *__restrict__ n_cands,
                                                         uint32_t cand_cap, const uint4 *__restrict__ lazy_rk, uint64_t diff_lsb,
                                                         const AesTables *__restrict__ tables, Hit *__restrict__ hits,
                                                         uint32_t hit_cap, uint32_t *__restrict__ n_hits) {
    extern __shared__ uint32_t aes_sm[];
    aes_load_smem(aes_sm, tables);
    const uint32_t *tl = aes_sm + (threadIdx.x & 31);
    const uint32_t n = min(*n_cands, cand_cap);
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        const uint2 cd = cands[c];
        const uint4 lz = aes128_encrypt(tl, lazy_rk + 11 * cd.y, labels[cd.x]);
        const uint64_t lsb = ((uint64_t)lz.x | ((uint64_t)lz.y << 32)) & 0x00ffffffffffffffull;
        if (lsb >= diff_lsb) continue;
        const uint32_t pos = atomicAdd(n_hits, 1u);
        if (pos < hit_cap) hits[pos] = Hit{cd.y, 0, first_index + cd.x};
    }
}

#define P_TRY(expr)                                                                                                   \
    do {                                                                                                              \
        cudaError_t e__ = (expr);                                                                                     \
        if (e__ != cudaSuccess) {                                                                                     \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));                                           \
            return e__ == cudaErrorMemoryAllocation ? B200POST_ERR_OUT_OF_MEMORY : B200POST_ERR_CUDA;                 \
        }                                                                                                             \
    } while (0)

// Streaming scan state: device buffers, keys, per-nonce hit lists.
class Scanner {
public:
    ~Scanner() {
        if (dev_ >= 0) cudaSetDevice(dev_);
        cudaFree(d_cands_); cudaFree(d_ncands_);
        for (int b = 0; b < 2; b++) { cudaFree(d_labels_[b]); cudaFreeHost(h_labels_[b]); if (ev_[b]) cudaEventDestroy(ev_[b]); if (st_[b]) cudaStreamDestroy(st_[b]); cudaFree(d_hits_[b]); cudaFree(d_nhits_[b]); cudaFreeHost(h_hits_[b]); cudaFreeHost(h_nhits_[b]); cudaFreeHost(h_ncands_[b]); }
        cudaFree(d_rk_); cudaFree(d_lazy_); cudaFree(d_tables_);
    }
    int init(uint32_t provider, const uint8_t challenge[32], uint32_t nonces, const uint64_t *pows, uint32_t k1, uint32_t k2,
             uint64_t num_labels, uint64_t chunk) {
        DeviceEngine *e = engine_for(provider);
        if (!e) return provider == B200POST_CPU_PROVIDER_ID ? B200POST_ERR_UNSUPPORTED : B200POST_ERR_NO_DEVICE;
        if (nonces == 0 || nonces % 16 || nonces > 4096 || k1 == 0 || k2 == 0 || num_labels == 0 || chunk == 0 || chunk > (1u << 28)) {
            set_error("invalid proving parameters (nonces must be a positive multiple of 16, <= 4096)");
            return B200POST_ERR_INVALID_ARGUMENT;
        }
        dev_ = e->device(); nonces_ = nonces; k2_ = k2; chunk_ = chunk;
        const uint64_t diff = b200post_proving_difficulty(k1, num_labels);
        msb_ = (uint32_t)(diff >> 56); lsb_ = diff & 0x00ffffffffffffffull;
        // hits per chunk are ~ chunk * nonces * K1/numLabels; leave generous slack, cap the buffer at 64 MiB
        const double expect = (double)chunk * nonces * ((double)k1 / (double)num_labels);
        hit_cap_ = (uint32_t)std::min<double>(std::max<double>(4.0 * expect + 65536.0, 65536.0), 4.0 * 1024 * 1024);
        P_TRY(cudaSetDevice(dev_));
        std::vector<uint8_t> rk((size_t)(nonces / 16) * 176), lazy((size_t)nonces * 176);
        for (uint32_t g = 0; g < nonces / 16; g++) {
            uint8_t key[16];
            cipher_key(challenge, g, pows[g], nullptr, key);
            const Aes128 a(key);
            memcpy(rk.data() + (size_t)g * 176, a.rk, 176);
        }
        for (uint32_t n = 0; n < nonces; n++) {
            uint8_t key[16];
            cipher_key(challenge, n / 16, pows[n / 16], &n, key);
            const Aes128 a(key);
            memcpy(lazy.data() + (size_t)n * 176, a.rk, 176);
        }
        static AesTables host_tables;
        static std::once_flag once;
        std::call_once(once, [] { aes_build_tables(host_tables); });
        P_TRY(cudaMalloc(&d_rk_, rk.size()));
        P_TRY(cudaMalloc(&d_lazy_, lazy.size()));
        P_TRY(cudaMalloc(&d_tables_, sizeof(AesTables)));
        P_TRY(cudaMemcpy(d_rk_, rk.data(), rk.size(), cudaMemcpyHostToDevice));
        P_TRY(cudaMemcpy(d_lazy_, lazy.data(), lazy.size(), cudaMemcpyHostToDevice));
        P_TRY(cudaMemcpy(d_tables_, &host_tables, sizeof(AesTables), cudaMemcpyHostToDevice));
        for (int b = 0; b < 2; b++) {
            P_TRY(cudaStreamCreateWithFlags(&st_[b], cudaStreamNonBlocking));
            P_TRY(cudaEventCreateWithFlags(&ev_[b], cudaEventDisableTiming));
            P_TRY(cudaMalloc(&d_labels_[b], chunk * 16));
            P_TRY(cudaMallocHost(&h_labels_[b], chunk * 16));
            P_TRY(cudaMalloc(&d_hits_[b], (size_t)hit_cap_ * sizeof(Hit)));
            P_TRY(cudaMalloc(&d_nhits_[b], 4));
            P_TRY(cudaMallocHost(&h_hits_[b], (size_t)hit_cap_ * sizeof(Hit)));
            P_TRY(cudaMallocHost(&h_nhits_[b], 4));
            P_TRY(cudaMallocHost(&h_ncands_[b], 4));
        }
        // lazy-cipher candidates: one ciphertext byte in 256 equals the MSB; 2x slack, shared by both buffers
        // (chunks are processed in stream order on alternating streams, so the queue is fenced by events below)
        cand_cap_ = (uint32_t)std::min<uint64_t>((chunk * nonces) / 128 + 65536, 1u << 27);
        P_TRY(cudaMalloc(&d_cands_, (size_t)cand_cap_ * sizeof(uint2)));
        P_TRY(cudaMalloc(&d_ncands_, 8));
        cudaDeviceProp p;
        P_TRY(cudaGetDeviceProperties(&p, dev_));
        grid_ = (uint32_t)p.multiProcessorCount * 6;   // 6 CTAs x 32 KiB of lane-replicated AES table per SM
        return B200POST_OK;
    }
    uint8_t *staging(int b) { return h_labels_[b]; }
    // enqueue chunk in staging(b): labels [first, first+count)
    int submit(int b, uint64_t first, uint32_t count) {
        P_TRY(cudaMemcpyAsync(d_labels_[b], h_labels_[b], (size_t)count * 16, cudaMemcpyHostToDevice, st_[b]));
        P_TRY(cudaMemsetAsync(d_nhits_[b], 0, 4, st_[b]));
        // the single candidate queue is reused by consecutive chunks: wait for the other stream's lazy pass
        if (pending_[b ^ 1]) P_TRY(cudaStreamWaitEvent(st_[b], ev_[b ^ 1], 0));
        P_TRY(cudaMemsetAsync(d_ncands_, 0, 8, st_[b]));
        prove_scan_kernel<<<grid_, 256, AES_SMEM_BYTES, st_[b]>>>(reinterpret_cast<const uint4 *>(d_labels_[b]), first, count,
                                                                reinterpret_cast<const uint4 *>(d_rk_), nonces_ / 16, msb_, d_tables_,
                                                                d_hits_[b], hit_cap_, d_nhits_[b], d_cands_, cand_cap_, d_ncands_);
        prove_lazy_kernel<<<grid_, 256, AES_SMEM_BYTES, st_[b]>>>(reinterpret_cast<const uint4 *>(d_labels_[b]), first, d_cands_, d_ncands_,
                                                                cand_cap_, reinterpret_cast<const uint4 *>(d_lazy_), lsb_, d_tables_,
                                                                d_hits_[b], hit_cap_, d_nhits_[b]);
        g_launches += 2;
        P_TRY(cudaGetLastError());
        P_TRY(cudaMemcpyAsync(h_ncands_[b], d_ncands_, 4, cudaMemcpyDeviceToHost, st_[b]));
        P_TRY(cudaMemcpyAsync(h_nhits_[b], d_nhits_[b], 4, cudaMemcpyDeviceToHost, st_[b]));
        P_TRY(cudaMemcpyAsync(h_hits_[b], d_hits_[b], (size_t)hit_cap_ * sizeof(Hit), cudaMemcpyDeviceToHost, st_[b]));
        P_TRY(cudaEventRecord(ev_[b], st_[b]));
        pending_[b] = true; end_[b] = first + count;
        return B200POST_OK;
    }
    // wait for chunk b and fold its hits in; *found set when some nonce has K2 hits
    int collect(int b, bool *found) {
        if (!pending_[b]) return B200POST_OK;
        P_TRY(cudaEventSynchronize(ev_[b]));
        pending_[b] = false;
        const uint32_t n = *h_nhits_[b];
        if (n > hit_cap_ || *h_ncands_[b] > cand_cap_) { set_error("hit buffer overflow: K1 too large for this chunk size"); return B200POST_ERR_OUT_OF_MEMORY; }
        std::vector<Hit> v(h_hits_[b], h_hits_[b] + n);
        std::sort(v.begin(), v.end(), [](const Hit &x, const Hit &y) { return x.index != y.index ? x.index < y.index : x.nonce < y.nonce; });
        for (const Hit &h : v) {
            std::vector<uint64_t> &l = lists_[h.nonce];
            if (l.size() < k2_) l.push_back(h.index);
        }
        scanned_ = std::max(scanned_, end_[b]);
        for (auto &kv : lists_) if (kv.second.size() >= k2_) *found = true;
        return B200POST_OK;
    }
    // the winning nonce: lowest K2-th hit index, ties to the lower nonce
    bool winner(uint32_t *nonce, std::vector<uint64_t> *indices) const {
        bool have = false;
        for (const auto &kv : lists_) {
            if (kv.second.size() < k2_) continue;
            if (!have || kv.second[k2_ - 1] < (*indices)[k2_ - 1]) { *nonce = kv.first; *indices = kv.second; have = true; }
        }
        return have;
    }
    uint64_t scanned() const { return scanned_; }

private:
    int dev_ = -1;
    uint32_t nonces_ = 0, k2_ = 0, msb_ = 0, hit_cap_ = 0, grid_ = 0;
    uint64_t lsb_ = 0, chunk_ = 0, scanned_ = 0;
    uint8_t *d_rk_ = nullptr, *d_lazy_ = nullptr;
    AesTables *d_tables_ = nullptr;
    cudaStream_t st_[2] = {nullptr, nullptr};
    cudaEvent_t ev_[2] = {nullptr, nullptr};
    uint8_t *d_labels_[2] = {nullptr, nullptr}, *h_labels_[2] = {nullptr, nullptr};
    Hit *d_hits_[2] = {nullptr, nullptr}, *h_hits_[2] = {nullptr, nullptr};
    uint32_t *d_nhits_[2] = {nullptr, nullptr}, *h_nhits_[2] = {nullptr, nullptr}, *h_ncands_[2] = {nullptr, nullptr};
    uint2 *d_cands_ = nullptr;
    uint32_t *d_ncands_ = nullptr;
    uint32_t cand_cap_ = 0;
    bool pending_[2] = {false, false};
    uint64_t end_[2] = {0, 0};
    std::map<uint32_t, std::vector<uint64_t>> lists_;   // ordered: ties resolve to the lower nonce
};

int finish(const Scanner &sc, uint32_t nonces, const uint64_t *pows, uint64_t num_labels, b200post_proof_out *out) {
    uint32_t nonce = 0;
    std::vector<uint64_t> idx;
    if (!sc.winner(&nonce, &idx)) { set_error("no proof found: no nonce reached K2 qualifying labels"); return B200POST_ERR_INVALID_PROOF; }
    (void)nonces;
    metrics().prove_labels_scanned_total += sc.scanned(); metrics().proofs_generated_total++;
    memset(out, 0, sizeof *out);
    out->nonce = nonce; out->pow = pows[nonce / 16]; out->labels_scanned = sc.scanned();
    out->indices_len = b200post_pack_indices(idx.data(), idx.size(), b200post_bits_per_index(num_labels), out->indices, sizeof out->indices);
    if (out->indices_len == 0) { set_error("packed indices exceed the 800-byte wire cap"); return B200POST_ERR_INVALID_ARGUMENT; }
    return B200POST_OK;
}

}  // namespace
}  // namespace b200post

using namespace b200post;

namespace {
// pread of [off, off + bytes) into dst on up to 8 threads (pread is position-independent, so slices are independent)
bool parallel_pread(int fd, uint8_t *dst, size_t bytes, off_t off) {
    auto read_all = [fd](uint8_t *d, size_t n, off_t o) {
        while (n) {
            const ssize_t r = pread(fd, d, n, o);
            if (r <= 0) return false;
            d += r; n -= (size_t)r; o += r;
        }
        return true;
    };
    const size_t kMinSlice = (size_t)4 << 20;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nt = std::min<size_t>({(size_t)8, (size_t)hw, std::max<size_t>(1, bytes / kMinSlice)});
    if (nt <= 1) return read_all(dst, bytes, off);
    std::vector<std::thread> th;
    std::vector<char> ok(nt, 0);
    const size_t per = (bytes / nt + 15) & ~(size_t)15;
    for (size_t t = 0; t < nt; t++) {
        const size_t lo = std::min(bytes, t * per), hi = t + 1 == nt ? bytes : std::min(bytes, (t + 1) * per);
        th.emplace_back([&, t, lo, hi] { ok[t] = read_all(dst + lo, hi - lo, off + (off_t)lo); });
    }
    for (auto &x : th) x.join();
    for (char c : ok) if (!c) return false;
    return true;
}
// the same for labels already in (pageable) host memory
void parallel_copy(uint8_t *dst, const uint8_t *src, size_t bytes) {
    const size_t kMinSlice = (size_t)4 << 20;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nt = std::min<size_t>({(size_t)8, (size_t)hw, std::max<size_t>(1, bytes / kMinSlice)});
    if (nt <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = (bytes / nt + 63) & ~(size_t)63;
    for (size_t t = 0; t < nt; t++) {
        const size_t lo = std::min(bytes, t * per), hi = t + 1 == nt ? bytes : std::min(bytes, (t + 1) * per);
        th.emplace_back([=] { memcpy(dst + lo, src + lo, hi - lo); });
    }
    for (auto &x : th) x.join();
}
}  // namespace

extern "C" {

int b200post_prove_scan(uint32_t provider, const uint8_t *labels16, uint64_t first_index, uint64_t count, const uint8_t challenge[32],
                        uint32_t nonces, const uint64_t *pows, uint32_t k1, uint32_t k2, uint64_t num_labels, b200post_proof_out *out) {
    if (!labels16 || !challenge || !pows || !out) { set_error("invalid argument"); return B200POST_ERR_INVALID_ARGUMENT; }
    Scanner sc;
    const uint64_t chunk = std::min<uint64_t>(std::max<uint64_t>(count, 1), 1u << 22);
    int rc = sc.init(provider, challenge, nonces, pows, k1, k2, num_labels, chunk);
    if (rc) return rc;
    bool found = false;
    int b = 0;
    for (uint64_t off = 0; off < count && !found; off += chunk, b ^= 1) {
        if ((rc = sc.collect(b, &found))) return rc;
        if (found) break;
        const uint32_t n = (uint32_t)std::min<uint64_t>(chunk, count - off);
        parallel_copy(sc.staging(b), labels16 + off * 16, (size_t)n * 16);   // pageable -> pinned staging, the scan's host-side bound
        if ((rc = sc.submit(b, first_index + off, n))) return rc;
    }
    for (int k = 0; k < 2; k++) if ((rc = sc.collect(b ^ k, &found))) return rc;   // older chunk first
    return finish(sc, nonces, pows, num_labels, out);
}

int b200post_generate_proof(const char *data_dir, const uint8_t challenge[32], const b200post_post_config *cfg,
                            const b200post_prove_opts *opts, b200post_proof_out *out, b200post_proof_metadata *meta_out,
                            const volatile int *cancel) {
    if (!data_dir || !challenge || !cfg || !out) { set_error("invalid argument"); return B200POST_ERR_INVALID_ARGUMENT; }
    b200post_prove_opts o{};
    if (opts) o = *opts;
    if (o.nonces == 0) o.nonces = 16;
    if (o.chunk_labels == 0) o.chunk_labels = 1ull << 22;
    b200post_post_metadata md;
    int rc = b200post_load_metadata(data_dir, &md);
    if (rc) return rc;
    const uint64_t num_labels = (uint64_t)md.num_units * md.labels_per_unit;
    if (num_labels == 0 || o.nonces % 16 || o.nonces > 4096) { set_error("invalid metadata or nonce count"); return B200POST_ERR_INVALID_ARGUMENT; }
    // k2pow per nonce group (RandomX upstream) through the caller's hook
    std::vector<uint64_t> pows(o.nonces / 16, 0);
    if (o.pow_mode > B200POST_POW_SKIP || (o.pow_mode == B200POST_POW_CALLBACK && !o.pow_prove)) {
        set_error("pow_mode CALLBACK needs a pow_prove function; to prove without k2pow ask for B200POST_POW_SKIP explicitly");
        return B200POST_ERR_UNSUPPORTED;
    }
    if (o.pow_mode != B200POST_POW_SKIP) {
        uint8_t scaled[32];
        div256_u32(cfg->pow_difficulty, md.num_units, scaled);
        if (o.pow_mode == B200POST_POW_CALLBACK) {
            for (uint32_t g = 0; g < o.nonces / 16; g++)
                if (o.pow_prove(o.pow_ctx, (uint8_t)g, challenge, scaled, md.node_id, &pows[g]) != 0) { set_error("k2pow hook failed"); return B200POST_ERR_INVALID_ARGUMENT; }
        } else {
            // the k2pow step of NIPostBuilder.Proof (activation/nipost.go:171 -> post-service): RandomX nonce search on the device
            b200post_k2pow_params kp{};
            kp.cache_key = o.pow_cache_key; kp.cache_key_len = o.pow_cache_key_len;
            memcpy(kp.challenge8, challenge, 8);
            memcpy(kp.node_id, md.node_id, 32);
            memcpy(kp.difficulty, scaled, 32);
            if ((rc = b200post_k2pow_search_groups(o.provider, &kp, o.nonces / 16, 0, pows.data(), nullptr, cancel))) return rc;
            for (uint64_t v : pows) if (v == B200POST_K2POW_NOT_FOUND) { set_error("k2pow: nonce space exhausted"); return B200POST_ERR_INVALID_PROOF; }
        }
    }
    Scanner sc;
    const uint64_t chunk = std::min<uint64_t>(o.chunk_labels, num_labels);
    if ((rc = sc.init(o.provider, challenge, o.nonces, pows.data(), cfg->k1, cfg->k2, num_labels, chunk))) return rc;

    const uint64_t per_file = md.max_file_size / 16;
    if (per_file == 0) { set_error("corrupt metadata: MaxFileSize"); return B200POST_ERR_IO; }
    bool found = false;
    int b = 0, fd = -1;
    uint64_t open_file = ~0ull;
    for (uint64_t pos = 0; pos < num_labels && !found; b ^= 1) {
        if (cancel && *cancel) { if (fd >= 0) close(fd); set_error("cancelled"); return B200POST_ERR_CANCELLED; }
        if ((rc = sc.collect(b, &found))) { if (fd >= 0) close(fd); return rc; }
        if (found) break;
        // fill the staging buffer from the files (a chunk may span files)
        uint64_t n = 0;
        const uint64_t want = std::min<uint64_t>(chunk, num_labels - pos);
        while (n < want) {
            const uint64_t file = (pos + n) / per_file, in_file = (pos + n) % per_file;
            if (file != open_file) {
                if (fd >= 0) close(fd);
                const std::string path = std::string(data_dir) + "/postdata_" + std::to_string(file) + ".bin";
                fd = open(path.c_str(), O_RDONLY);
                if (fd < 0) { set_error("open " + path + ": " + strerror(errno)); return B200POST_ERR_IO; }
                open_file = file;
            }
            const uint64_t take = std::min<uint64_t>(want - n, per_file - in_file);
            // the read into the pinned staging buffer was the scan's bound (6.7-7.5 GB/s on one thread, r01): split it
            if (!parallel_pread(fd, sc.staging(b) + n * 16, (size_t)take * 16, (off_t)(in_file * 16))) { close(fd); set_error("POST data is incomplete (short read): initialisation not finished?"); return B200POST_ERR_IO; }
            n += take;
        }
        if ((rc = sc.submit(b, pos, (uint32_t)n))) { close(fd); return rc; }
        pos += n;
    }
    if (fd >= 0) close(fd);
    for (int k = 0; k < 2; k++) if ((rc = sc.collect(b ^ k, &found))) return rc;   // older chunk first
    if ((rc = finish(sc, o.nonces, pows.data(), num_labels, out))) return rc;
    if (meta_out) {
        memcpy(meta_out->node_id, md.node_id, 32);
        memcpy(meta_out->commitment_atx_id, md.commitment_atx_id, 32);
        memcpy(meta_out->challenge, challenge, 32);
        meta_out->num_units = md.num_units; meta_out->labels_per_unit = md.labels_per_unit;
    }
    return B200POST_OK;
}

}  // extern "C"

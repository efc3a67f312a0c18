// engine.cu — wave scheduler for the POST label kernels (see engine.h).
//
// Mirrors what the reference's initializer does around libpost's `initialize()` (activation/post.go:295:
// batches of ComputeBatchSize labels, cancellable, progress observable) but sized for a B200: a wave is
// every scratchpad that fits the SMs (and HBM) at once, waves are queued two deep on one stream, and the
// 16-byte labels of wave w are copied out while wave w+1 computes.
#include "engine.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

#include "../../include/b200post.h"

namespace b200post {

Options &options() { static Options o; return o; }
std::atomic<uint64_t> g_launches{0};

static thread_local std::string t_error;
void set_error(const std::string &msg) { t_error = msg; }
const char *last_error() { return t_error.c_str(); }

#define CU_TRY(expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));                              \
            return e__ == cudaErrorMemoryAllocation ? B200POST_ERR_OUT_OF_MEMORY : B200POST_ERR_CUDA;    \
        }                                                                                                \
    } while (0)

static inline uint32_t round_up(uint32_t x, uint32_t m) { return (x + m - 1) / m * m; }

DeviceEngine::DeviceEngine(int device) : dev_(device) {
    cudaGetDeviceProperties(&prop_, device);
}

DeviceEngine::~DeviceEngine() {
    cudaSetDevice(dev_);
    release();
}

void DeviceEngine::release() {
    if (stream_) cudaStreamSynchronize(stream_);
    cudaFree(V_); V_ = nullptr; v_bytes_ = 0;
    cudaFree(X_); X_ = nullptr;
    for (int b = 0; b < 2; b++) {
        cudaFree(d_out_[b]); d_out_[b] = nullptr;
        cudaFreeHost(h_out_[b]); h_out_[b] = nullptr;
        if (ev_done_[b]) cudaEventDestroy(ev_done_[b]);
        if (ev_k2a_[b]) cudaEventDestroy(ev_k2a_[b]);
        if (ev_k2b_[b]) cudaEventDestroy(ev_k2b_[b]);
        if (ev_call_[b]) cudaEventDestroy(ev_call_[b]);
        ev_done_[b] = ev_k2a_[b] = ev_k2b_[b] = ev_call_[b] = nullptr;
        k2_pending_[b] = false;
    }
    cudaFree(d_commit_); d_commit_ = nullptr;
    cudaFree(d_idx_); d_idx_ = nullptr;
    cudaFreeHost(h_commit_); h_commit_ = nullptr;
    cudaFreeHost(h_idx_); h_idx_ = nullptr;
    cudaFree(d_mid_); d_mid_ = nullptr;
    cudaFree(d_diff_); d_diff_ = nullptr;
    cudaFree(d_cta_cand_); d_cta_cand_ = nullptr;
    cudaFree(d_running_); d_running_ = nullptr;
    cudaFreeHost(h_running_); h_running_ = nullptr;
    if (stream_) cudaStreamDestroy(stream_);
    stream_ = nullptr;
    alloc_slots_ = 0;
    wave_slots_ = 0;
}

// Decide the wave size for scrypt-N and make sure scratch for min(wave, want_slots) slots exists.
int DeviceEngine::ensure(uint64_t N, uint64_t want_slots) {
    Options &o = options();
    const int variant = (int)o.romix_variant.load(), mw = (int)o.mulwide_mask.load(), tpb = (int)o.tpb.load();
    if (!stream_) {
        CU_TRY(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            CU_TRY(cudaEventCreateWithFlags(&ev_done_[b], cudaEventDisableTiming));
            CU_TRY(cudaEventCreate(&ev_k2a_[b]));
            CU_TRY(cudaEventCreate(&ev_k2b_[b]));
            CU_TRY(cudaEventCreate(&ev_call_[b]));
        }
        CU_TRY(cudaMalloc(&d_diff_, 32));
        CU_TRY(cudaMalloc(&d_running_, sizeof(VrfCandidate)));
        CU_TRY(cudaMallocHost(&h_running_, sizeof(VrfCandidate)));
    }
    variant_ = variant; mw_ = mw; tpb_ = tpb; policy_ = (int)o.mem_policy.load();
    int ctas = romix_max_ctas_per_sm(variant, mw, policy_, tpb);
    if (ctas <= 0) { set_error("romix kernel cannot be resident (bad variant/mask/tpb?)"); return B200POST_ERR_INVALID_ARGUMENT; }
    const int64_t want_ctas = o.ctas_per_sm.load();
    if (want_ctas > 0) ctas = std::min<int>(ctas, (int)want_ctas);

    const size_t per_slot = 128 * (size_t)N;
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = (size_t)((double)(free_b + v_bytes_) * 0.90);
    const int64_t cap_mib = o.max_scratch_mib.load();
    if (cap_mib > 0) budget = std::min(budget, (size_t)cap_mib << 20);
    const size_t per_cta_layer = per_slot * (size_t)tpb * (size_t)prop_.multiProcessorCount;
    while (ctas > 1 && per_cta_layer * (size_t)ctas > budget) ctas--;
    uint64_t wave = (uint64_t)prop_.multiProcessorCount * (uint64_t)ctas * (uint64_t)tpb;
    if (per_slot * wave > budget) {
        // not even one CTA per SM: shrink to what fits, in whole CTAs
        wave = budget / per_slot / (uint64_t)tpb * (uint64_t)tpb;
        if (wave == 0) { set_error("not enough HBM for one CTA of ROMix scratch"); return B200POST_ERR_OUT_OF_MEMORY; }
    }
    wave_slots_ = (uint32_t)wave;

    const uint32_t need = (uint32_t)std::min<uint64_t>(wave, round_up((uint32_t)std::min<uint64_t>(want_slots, wave), 32));
    const size_t need_v = per_slot * (size_t)need;
    if (need_v > v_bytes_) {
        CU_TRY(cudaStreamSynchronize(stream_));
        cudaFree(V_); V_ = nullptr; v_bytes_ = 0;
        CU_TRY(cudaMalloc(&V_, need_v));
        v_bytes_ = need_v;
    }
    if (need > alloc_slots_) {
        CU_TRY(cudaStreamSynchronize(stream_));
        cudaFree(X_); cudaFree(d_commit_); cudaFree(d_idx_); cudaFree(d_mid_); cudaFree(d_cta_cand_);
        cudaFreeHost(h_commit_); cudaFreeHost(h_idx_);
        X_ = nullptr; d_commit_ = nullptr; d_idx_ = nullptr; d_mid_ = nullptr; d_cta_cand_ = nullptr;
        h_commit_ = nullptr; h_idx_ = nullptr;
        for (int b = 0; b < 2; b++) { cudaFree(d_out_[b]); cudaFreeHost(h_out_[b]); d_out_[b] = nullptr; h_out_[b] = nullptr; }
        alloc_slots_ = 0;
        CU_TRY(cudaMalloc(&X_, (size_t)need * 128));
        CU_TRY(cudaMalloc(&d_commit_, (size_t)need * 32));
        CU_TRY(cudaMalloc(&d_idx_, (size_t)need * 8));
        CU_TRY(cudaMalloc(&d_mid_, (size_t)need * 64));
        CU_TRY(cudaMalloc(&d_cta_cand_, (size_t)pbkdf2_final_ctas(need) * sizeof(VrfCandidate)));
        CU_TRY(cudaMallocHost(&h_commit_, (size_t)need * 32));
        CU_TRY(cudaMallocHost(&h_idx_, (size_t)need * 8));
        for (int b = 0; b < 2; b++) {
            CU_TRY(cudaMalloc(&d_out_[b], (size_t)need * 16));
            CU_TRY(cudaMallocHost(&h_out_[b], (size_t)need * 16));
        }
        alloc_slots_ = need;
    }
    return B200POST_OK;
}

// collect the ROMix device time of the wave that last used buffer `buf` (its events have completed)
void DeviceEngine::harvest(int buf) {
    if (!k2_pending_[buf]) return;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ev_k2a_[buf], ev_k2b_[buf]) == cudaSuccess) { romix_ms_ += ms; romix_launches_++; }
    k2_pending_[buf] = false;
}

int DeviceEngine::run_wave(const LabelJob &job, uint32_t n_slots, uint64_t N, uint8_t *d_out, const uint32_t *d_diff, int buf) {
    CU_TRY(launch_pbkdf2_expand(job, X_, alloc_slots_, n_slots, stream_));
    RomixParams rp;
    rp.V = V_; rp.X = X_; rp.x_stride = alloc_slots_; rp.N = (uint32_t)N; rp.n_slots = n_slots;
    rp.flags = (uint32_t)options().debug_skip_phase.load();
    rp.rc = RotConsts{1u << 7, 1u << 9, 1u << 13, 1u << 18};
    CU_TRY(cudaEventRecord(ev_k2a_[buf], stream_));
    CU_TRY(launch_romix(variant_, mw_, policy_, tpb_, rp, stream_));
    CU_TRY(cudaEventRecord(ev_k2b_[buf], stream_));
    k2_pending_[buf] = true;
    CU_TRY(launch_pbkdf2_final(job, X_, alloc_slots_, n_slots, d_out, d_diff, d_cta_cand_, stream_));
    g_launches += 3;
    if (d_diff) {
        CU_TRY(launch_vrf_merge(d_cta_cand_, pbkdf2_final_ctas(n_slots), d_running_, stream_));
        g_launches += 1;
    }
    return B200POST_OK;
}

int DeviceEngine::labels_range(const uint8_t commitment[32], uint64_t N, uint64_t start, uint64_t count, uint8_t *out_host,
                               uint8_t *out_dev, const uint8_t *vrf_difficulty, VrfResult *vrf, const volatile int *cancel) {
    std::lock_guard<std::mutex> lk(mu_);
    CU_TRY(cudaSetDevice(dev_));
    if (vrf) *vrf = VrfResult{};
    if (count == 0) return B200POST_OK;
    int rc = ensure(N, count);
    if (rc) return rc;

    CU_TRY(cudaEventRecord(ev_call_[0], stream_));
    // per-call constants: commitment -> HMAC midstates (K0), VRF threshold, running candidate
    CU_TRY(cudaMemcpyAsync(d_commit_, commitment, 32, cudaMemcpyHostToDevice, stream_));
    CU_TRY(launch_hmac_midstates(d_commit_, 1, d_mid_, stream_));
    g_launches += 1;
    const uint32_t *d_diff = nullptr;
    if (vrf_difficulty) {
        uint32_t be[8];
        for (int k = 0; k < 8; k++)
            be[k] = ((uint32_t)vrf_difficulty[4 * k] << 24) | ((uint32_t)vrf_difficulty[4 * k + 1] << 16) |
                    ((uint32_t)vrf_difficulty[4 * k + 2] << 8) | vrf_difficulty[4 * k + 3];
        CU_TRY(cudaMemcpyAsync(d_diff_, be, 32, cudaMemcpyHostToDevice, stream_));
        CU_TRY(cudaMemsetAsync(d_running_, 0, sizeof(VrfCandidate), stream_));
        CU_TRY(cudaStreamSynchronize(stream_));   // `be` is a stack buffer
        d_diff = d_diff_;
    }

    const uint64_t wave = std::min<uint64_t>(wave_slots_, alloc_slots_);
    struct Pending { uint64_t off; uint32_t n; bool live; } pend[2] = {{0, 0, false}, {0, 0, false}};
    int status = B200POST_OK;
    uint64_t done = 0;
    int w = 0;
    auto retire = [&](int b) -> int {
        if (!pend[b].live) return B200POST_OK;
        CU_TRY(cudaEventSynchronize(ev_done_[b]));
        harvest(b);
        if (out_host) memcpy(out_host + pend[b].off * 16, h_out_[b], (size_t)pend[b].n * 16);
        pend[b].live = false;
        return B200POST_OK;
    };
    while (done < count) {
        if (cancel && *cancel) { status = B200POST_ERR_CANCELLED; break; }
        const int b = w & 1;
        if ((rc = retire(b))) return rc;   // buffer b (wave w-2) must be drained before reuse
        const uint32_t n_valid = (uint32_t)std::min<uint64_t>(wave, count - done);
        const uint32_t n_slots = round_up(n_valid, 32);
        LabelJob job{d_mid_, 0, nullptr, start + done, n_valid};
        uint8_t *d_out = out_dev ? out_dev + done * 16 : d_out_[b];
        if ((rc = run_wave(job, n_slots, N, d_out, d_diff, b))) return rc;
        if (out_host) CU_TRY(cudaMemcpyAsync(h_out_[b], d_out_[b], (size_t)n_valid * 16, cudaMemcpyDeviceToHost, stream_));
        CU_TRY(cudaEventRecord(ev_done_[b], stream_));
        pend[b] = Pending{done, n_valid, true};
        done += n_valid;
        w++;
    }
    for (int k = 0; k < 2; k++) if ((rc = retire((w + k) & 1))) return rc;
    if (status == B200POST_OK && vrf_difficulty && vrf) {
        CU_TRY(cudaMemcpyAsync(h_running_, d_running_, sizeof(VrfCandidate), cudaMemcpyDeviceToHost, stream_));
        CU_TRY(cudaStreamSynchronize(stream_));
        vrf->found = h_running_->found != 0;
        if (vrf->found) {
            vrf->index = h_running_->index;
            for (int k = 0; k < 8; k++) {
                const uint32_t v = h_running_->label_be[k];
                vrf->label32[4 * k] = (uint8_t)(v >> 24); vrf->label32[4 * k + 1] = (uint8_t)(v >> 16);
                vrf->label32[4 * k + 2] = (uint8_t)(v >> 8); vrf->label32[4 * k + 3] = (uint8_t)v;
            }
        }
    }
    CU_TRY(cudaEventRecord(ev_call_[1], stream_));
    CU_TRY(cudaStreamSynchronize(stream_));
    { float ms = 0; if (cudaEventElapsedTime(&ms, ev_call_[0], ev_call_[1]) == cudaSuccess) last_call_ms_ = ms; }
    if (status == B200POST_ERR_CANCELLED) set_error("cancelled");
    return status;
}

int DeviceEngine::labels_gather(size_t n_items, const uint8_t *commitments, const uint64_t *indices, uint64_t N, uint8_t *out_host) {
    std::lock_guard<std::mutex> lk(mu_);
    CU_TRY(cudaSetDevice(dev_));
    if (n_items == 0) return B200POST_OK;
    int rc = ensure(N, n_items);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(ev_call_[0], stream_));
    const uint64_t wave = std::min<uint64_t>(wave_slots_, alloc_slots_);
    uint64_t done = 0;
    // single-buffered inputs (they are consumed by K0/K1 at the head of the wave), double-buffered outputs
    struct Pending { uint64_t off; uint32_t n; bool live; } pend[2] = {{0, 0, false}, {0, 0, false}};
    auto retire = [&](int b) -> int {
        if (!pend[b].live) return B200POST_OK;
        CU_TRY(cudaEventSynchronize(ev_done_[b]));
        harvest(b);
        memcpy(out_host + pend[b].off * 16, h_out_[b], (size_t)pend[b].n * 16);
        pend[b].live = false;
        return B200POST_OK;
    };
    int w = 0;
    while (done < n_items) {
        const int b = w & 1;
        if ((rc = retire(b))) return rc;
        // the previous wave's H2D of the staging buffers must have been consumed: wave w-1's done event
        if (w > 0) CU_TRY(cudaEventSynchronize(ev_done_[(w - 1) & 1]));
        const uint32_t n_valid = (uint32_t)std::min<uint64_t>(wave, n_items - done);
        const uint32_t n_slots = round_up(n_valid, 32);
        memcpy(h_commit_, commitments + done * 32, (size_t)n_valid * 32);
        memcpy(h_idx_, indices + done, (size_t)n_valid * 8);
        CU_TRY(cudaMemcpyAsync(d_commit_, h_commit_, (size_t)n_valid * 32, cudaMemcpyHostToDevice, stream_));
        CU_TRY(cudaMemcpyAsync(d_idx_, h_idx_, (size_t)n_valid * 8, cudaMemcpyHostToDevice, stream_));
        CU_TRY(launch_hmac_midstates(d_commit_, n_valid, d_mid_, stream_));
        g_launches += 1;
        LabelJob job{d_mid_, 16, d_idx_, 0, n_valid};
        if ((rc = run_wave(job, n_slots, N, d_out_[b], nullptr, b))) return rc;
        CU_TRY(cudaMemcpyAsync(h_out_[b], d_out_[b], (size_t)n_valid * 16, cudaMemcpyDeviceToHost, stream_));
        CU_TRY(cudaEventRecord(ev_done_[b], stream_));
        pend[b] = Pending{done, n_valid, true};
        done += n_valid;
        w++;
    }
    for (int k = 0; k < 2; k++) if ((rc = retire((w + k) & 1))) return rc;
    CU_TRY(cudaEventRecord(ev_call_[1], stream_));
    CU_TRY(cudaStreamSynchronize(stream_));
    { float ms = 0; if (cudaEventElapsedTime(&ms, ev_call_[0], ev_call_[1]) == cudaSuccess) last_call_ms_ = ms; }
    return B200POST_OK;
}

uint32_t DeviceEngine::wave_slots(uint64_t N) {
    std::lock_guard<std::mutex> lk(mu_);
    if (cudaSetDevice(dev_) != cudaSuccess) return 0;
    if (ensure(N, 32) != B200POST_OK) return 0;
    return wave_slots_;
}

double DeviceEngine::last_call_ms() {
    std::lock_guard<std::mutex> lk(mu_);
    return last_call_ms_;
}

void DeviceEngine::romix_time(double *ms_total, uint64_t *launches, bool reset) {
    std::lock_guard<std::mutex> lk(mu_);
    if (ms_total) *ms_total = romix_ms_;
    if (launches) *launches = romix_launches_;
    if (reset) { romix_ms_ = 0; romix_launches_ = 0; }
}

// ------------------------------------------------------------------------------------------------ registry
static std::mutex g_reg_mu;
// heap-allocated and never destroyed at exit: static destructors may run after the CUDA runtime has
// torn down, where cudaFree is no longer legal.  b200post_shutdown() releases explicitly.
static std::map<int, std::unique_ptr<DeviceEngine>> &g_engines = *new std::map<int, std::unique_ptr<DeviceEngine>>();

int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

DeviceEngine *engine_for(uint32_t provider) {
    if (provider == B200POST_CPU_PROVIDER_ID) {
        set_error("provider 0xffffffff (CPU) is not served by libb200post: this library has no CPU path");
        return nullptr;
    }
    const int n = device_count();
    if ((int64_t)provider >= n) {
        set_error(n == 0 ? "no CUDA device available" : "unknown provider id");
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_engines.find((int)provider);
    if (it == g_engines.end()) it = g_engines.emplace((int)provider, std::make_unique<DeviceEngine>((int)provider)).first;
    return it->second.get();
}

void shutdown_all() {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    g_engines.clear();
}

}  // namespace b200post

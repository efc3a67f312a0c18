// randomx_engine.h — host runtime of the k2pow (RandomX) engine: per-device dataset residency, batch buffers, the launch
// sequence of one batch.  C++ for the same reason as engine.h (the reference's host side is compiled Go; no Go here).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>

#include "randomx_kernels.cuh"

namespace b200post {

class RandomxEngine {
public:
    explicit RandomxEngine(int device);
    ~RandomxEngine();
    RandomxEngine(const RandomxEngine &) = delete;

    int prepare(const std::string &key);
    int hash_inputs(const std::string &key, const uint8_t *inputs, size_t input_len, size_t n, uint8_t *out32);
    // hashes != nullptr: every hash of the range is returned (count x 32).  difficulty != nullptr: search semantics.
    int k2pow(const std::string &key, const rx::K2powTemplate &tmpl, const uint8_t *difficulty, uint64_t start, uint64_t count,
              uint8_t *hashes, uint64_t *found, uint64_t *done, const volatile int *cancel, uint64_t batch_stride = 0,
              const volatile int *peer_hit = nullptr);
    int batch_size(uint64_t *vms);
    void last_timing(double *total_ms, double *vm_ms, uint64_t *hashes, uint64_t *vm_launches);

private:
    int ensure_dataset(const std::string &key);
    int ensure_batch(uint32_t want);
    void release_batch();
    // runs seeds..finalize for the n VMs whose seeds are already written
    int run_chain(uint32_t n);
    uint32_t desired_batch() const;

    int dev_;
    cudaDeviceProp prop_{};
    std::mutex mu_;
    cudaStream_t stream_ = nullptr;
    bool tables_ = false;
    std::string key_;                       // key of the resident dataset ("" = none)
    uint64_t *d_dataset_ = nullptr;
    rx::BatchBuffers buf_;
    uint32_t cap_ = 0;
    uint8_t *d_inputs_ = nullptr; size_t inputs_cap_ = 0;
    uint8_t *d_diff_ = nullptr;
    uint32_t *d_found_ = nullptr;
    uint8_t *h_stage_ = nullptr; size_t stage_cap_ = 0;   // pinned: hashes coming back
    cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
    double total_ms_ = 0, vm_ms_ = 0;
    uint64_t hashes_ = 0, vm_launches_ = 0;
};

RandomxEngine *randomx_engine_for(uint32_t provider);
void randomx_shutdown_all();

}  // namespace b200post

// reference_label.cpp — ONE label on the host CPU, from the very header the kernels inline (post_device.cuh is
// host/device code).  It exists for the fault detector only: the reference's initializer cross-checks the provider's
// output against a CPU-computed label and reports ErrReferenceLabelMismatch (activation/post.go:299-312) — a check that
// compares the GPU against the GPU cannot see a device that miscomputes consistently.  It is NOT a compute path: nothing
// falls back to it, it computes one label per call on the calling thread (~3 ms at N = 8192), and it does not touch
// oracle/.
#include <cstring>
#include <vector>

#include "host_hash.h"
#include "post_device.cuh"

namespace b200post {

void reference_label32(const uint8_t commitment[32], uint64_t index, uint32_t n, uint8_t out[32]) {
    uint32_t c[8];
    memcpy(c, commitment, 32);
    uint32_t lo[16], hi[16];
    label_expand(c, index, lo, hi);
    std::vector<uint32_t> v((size_t)n * 32);
    for (uint32_t i = 0; i < n; i++) {
        memcpy(&v[(size_t)i * 32], lo, 64);
        memcpy(&v[(size_t)i * 32 + 16], hi, 64);
        blockmix_r1<0>(lo, hi);
    }
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t j = hi[0] & (n - 1);
        uint32_t vlo[16], vhi[16];
        memcpy(vlo, &v[(size_t)j * 32], 64);
        memcpy(vhi, &v[(size_t)j * 32 + 16], 64);
        blockmix_r1_xor<0>(lo, hi, vlo, vhi);
    }
    uint32_t lab[8];
    label_final(c, index, lo, hi, lab);
    for (int k = 0; k < 8; k++) { const uint32_t w = bswap32(lab[k]); memcpy(out + 4 * k, &w, 4); }
}

}  // namespace b200post

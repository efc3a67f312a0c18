// host_emul.cpp — compiles the product's per-thread device arithmetic (post_device.cuh) as plain C++
// so the exact code the kernels run can be checked against the oracle on a CPU-only box.
// Test infrastructure; built by tests/conftest.py with g++.
#include <cstring>
#include <vector>
#include "../go-spacemesh_b200/csrc/post_device.cuh"

using namespace b200post;

// one label the way a kernel thread computes it: label_expand (K1) -> ROMix (K2) -> label_final (K3)
extern "C" int emul_label32(const uint8_t commitment[32], uint64_t index, uint32_t N, uint8_t out[32]) {
    uint32_t c[8];
    memcpy(c, commitment, 32);   // 8 little-endian words, as the kernels load them
    uint32_t lo[16], hi[16];
    label_expand(c, index, lo, hi);
    std::vector<uint32_t> V((size_t)N * 32);
    for (uint32_t i = 0; i < N; i++) {
        memcpy(&V[(size_t)i * 32], lo, 64); memcpy(&V[(size_t)i * 32 + 16], hi, 64);
        blockmix_r1<0>(lo, hi);
    }
    for (uint32_t i = 0; i < N; i++) {
        const uint32_t j = hi[0] & (N - 1);
        uint32_t vlo[16], vhi[16];
        memcpy(vlo, &V[(size_t)j * 32], 64); memcpy(vhi, &V[(size_t)j * 32 + 16], 64);
        blockmix_r1_xor<0>(lo, hi, vlo, vhi);
    }
    uint32_t lab[8];
    label_final(c, index, lo, hi, lab);
    for (int k = 0; k < 8; k++) { uint32_t v = bswap32(lab[k]); memcpy(out + 4 * k, &v, 4); }
    return 0;
}

// the dual-label step (fill + mix interleaved) must equal the two single steps
extern "C" int emul_dual_step_matches(const uint32_t seed[64]) {
    uint32_t lo_f[16], hi_f[16], lo_m[16], hi_m[16], vlo[16], vhi[16];
    uint32_t a_lo[16], a_hi[16], b_lo[16], b_hi[16];
    for (int i = 0; i < 16; i++) {
        lo_f[i] = a_lo[i] = seed[i]; hi_f[i] = a_hi[i] = seed[16 + i];
        lo_m[i] = b_lo[i] = seed[32 + i]; hi_m[i] = b_hi[i] = seed[48 + i];
        vlo[i] = seed[i] * 2654435761u + 1; vhi[i] = seed[63 - i] ^ 0x9e3779b9u;
    }
    blockmix_r1_dual<0, 4>(lo_f, hi_f, lo_m, hi_m, vlo, vhi);
    blockmix_r1<0>(a_lo, a_hi);
    blockmix_r1_xor<0>(b_lo, b_hi, vlo, vhi);
    return !memcmp(lo_f, a_lo, 64) && !memcmp(hi_f, a_hi, 64) && !memcmp(lo_m, b_lo, 64) && !memcmp(hi_m, b_hi, 64);
}

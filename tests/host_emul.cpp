// host_emul.cpp — compiles the product's per-thread device arithmetic (post_device.cuh) as plain C++
// so the exact code the kernels run can be checked against the oracle on a CPU-only box.
// Test infrastructure; built by tests/conftest.py with g++.
#include <cstring>
#include <vector>
#include "../go-spacemesh_b200/csrc/post_device.cuh"

using namespace b200post;

extern "C" int emul_label32(const uint8_t commitment[32], uint64_t index, uint32_t N, uint8_t out[32]) {
    uint32_t key[8];
    for (int k = 0; k < 8; k++) { uint32_t c; memcpy(&c, commitment + 4 * k, 4); key[k] = bswap32(c); }
    HmacMid m;
    hmac_midstates(key, m);
    uint32_t lo[16], hi[16];
    pbkdf2_expand(m, index, lo, hi);
    RotConsts rc{1u << 7, 1u << 9, 1u << 13, 1u << 18};
    std::vector<uint32_t> V((size_t)N * 32);
    for (uint32_t i = 0; i < N; i++) {
        memcpy(&V[(size_t)i * 32], lo, 64); memcpy(&V[(size_t)i * 32 + 16], hi, 64);
        blockmix_r1<0>(lo, hi, rc);
    }
    for (uint32_t i = 0; i < N; i++) {
        const uint32_t j = hi[0] & (N - 1);
        uint32_t vlo[16], vhi[16];
        memcpy(vlo, &V[(size_t)j * 32], 64); memcpy(vhi, &V[(size_t)j * 32 + 16], 64);
        blockmix_r1_xor<0>(lo, hi, vlo, vhi, rc);
    }
    uint32_t lab[8];
    pbkdf2_final(m, lo, hi, lab);
    for (int k = 0; k < 8; k++) { uint32_t v = bswap32(lab[k]); memcpy(out + 4 * k, &v, 4); }
    return 0;
}

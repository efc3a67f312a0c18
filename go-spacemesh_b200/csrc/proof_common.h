// proof_common.h — host helpers shared by the verifier and the prover (ASSUMED post-rs conventions).
#pragma once
#include <immintrin.h>

#include <cstdint>
#include <cstring>

#include "host_hash.h"

namespace b200post {

// FIPS-197 AES-128 key schedule / single-block encryption with AES-NI (host side: round keys for the device
// kernels, and nothing label-sized — the per-label AES runs on the GPU).
struct Aes128 {
    __m128i rk[11];
    template <int RCON>
    static __m128i expand(__m128i k) {
        __m128i t = _mm_aeskeygenassist_si128(k, RCON);
        t = _mm_shuffle_epi32(t, 0xff);
        k = _mm_xor_si128(k, _mm_slli_si128(k, 4));
        k = _mm_xor_si128(k, _mm_slli_si128(k, 4));
        k = _mm_xor_si128(k, _mm_slli_si128(k, 4));
        return _mm_xor_si128(k, t);
    }
    explicit Aes128(const uint8_t key[16]) {
        rk[0] = _mm_loadu_si128(reinterpret_cast<const __m128i *>(key));
        rk[1] = expand<0x01>(rk[0]); rk[2] = expand<0x02>(rk[1]); rk[3] = expand<0x04>(rk[2]);
        rk[4] = expand<0x08>(rk[3]); rk[5] = expand<0x10>(rk[4]); rk[6] = expand<0x20>(rk[5]);
        rk[7] = expand<0x40>(rk[6]); rk[8] = expand<0x80>(rk[7]); rk[9] = expand<0x1b>(rk[8]);
        rk[10] = expand<0x36>(rk[9]);
    }
};

inline void put_le32(uint8_t *p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (8 * i)); }
inline void put_le64(uint8_t *p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i)); }

// key = blake3(challenge || LE32(nonce_group) || LE64(pow) [|| LE32(nonce)])[0:16]   (post-rs cipher.rs, ASSUMED)
inline void cipher_key(const uint8_t challenge[32], uint32_t nonce_group, uint64_t pow, const uint32_t *nonce, uint8_t key[16]) {
    uint8_t buf[48];
    memcpy(buf, challenge, 32);
    put_le32(buf + 32, nonce_group);
    put_le64(buf + 36, pow);
    size_t len = 44;
    if (nonce) { put_le32(buf + 44, *nonce); len = 48; }
    blake3_single_chunk(buf, len, key, 16);
}

// 256-bit big-endian value / 32-bit divisor (scale_pow_difficulty: difficulty / num_units, ASSUMED)
inline void div256_u32(const uint8_t in[32], uint32_t d, uint8_t out[32]) {
    uint64_t rem = 0;
    for (int i = 0; i < 32; i++) {
        const uint64_t cur = (rem << 8) | in[i];
        out[i] = (uint8_t)(cur / d);
        rem = cur % d;
    }
}

}  // namespace b200post

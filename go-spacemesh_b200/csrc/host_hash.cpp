// host_hash.cpp — see host_hash.h.  BLAKE3 per the BLAKE3 paper §2 (single-chunk case only).
#include "host_hash.h"

#include <array>
#include <cstring>

namespace b200post {
namespace {

constexpr std::array<uint32_t, 8> kIV = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                         0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
constexpr uint32_t kChunkStart = 1, kChunkEnd = 2, kRoot = 8;

inline uint32_t rotr(uint32_t v, unsigned n) { return (v >> n) | (v << (32 - n)); }

struct State {
    uint32_t v[16];
    void mix(int a, int b, int c, int d, uint32_t x, uint32_t y) {
        v[a] += v[b] + x; v[d] = rotr(v[d] ^ v[a], 16);
        v[c] += v[d];     v[b] = rotr(v[b] ^ v[c], 12);
        v[a] += v[b] + y; v[d] = rotr(v[d] ^ v[a], 8);
        v[c] += v[d];     v[b] = rotr(v[b] ^ v[c], 7);
    }
};

// message word schedule: round r uses m[sched[r][i]]
struct Schedule {
    uint8_t idx[7][16];
    constexpr Schedule() : idx{} {
        constexpr uint8_t perm[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
        for (int i = 0; i < 16; i++) idx[0][i] = (uint8_t)i;
        for (int r = 1; r < 7; r++)
            for (int i = 0; i < 16; i++) idx[r][i] = idx[r - 1][perm[i]];
    }
};
constexpr Schedule kSched{};

void compress(const uint32_t cv[8], const uint32_t m[16], uint64_t counter, uint32_t block_len, uint32_t flags,
              uint32_t out[16]) {
    State s;
    for (int i = 0; i < 8; i++) s.v[i] = cv[i];
    for (int i = 0; i < 4; i++) s.v[8 + i] = kIV[i];
    s.v[12] = (uint32_t)counter; s.v[13] = (uint32_t)(counter >> 32); s.v[14] = block_len; s.v[15] = flags;
    for (int r = 0; r < 7; r++) {
        const uint8_t *q = kSched.idx[r];
        s.mix(0, 4, 8, 12, m[q[0]], m[q[1]]);   s.mix(1, 5, 9, 13, m[q[2]], m[q[3]]);
        s.mix(2, 6, 10, 14, m[q[4]], m[q[5]]);  s.mix(3, 7, 11, 15, m[q[6]], m[q[7]]);
        s.mix(0, 5, 10, 15, m[q[8]], m[q[9]]);  s.mix(1, 6, 11, 12, m[q[10]], m[q[11]]);
        s.mix(2, 7, 8, 13, m[q[12]], m[q[13]]); s.mix(3, 4, 9, 14, m[q[14]], m[q[15]]);
    }
    for (int i = 0; i < 8; i++) { out[i] = s.v[i] ^ s.v[i + 8]; out[i + 8] = s.v[i + 8] ^ cv[i]; }
}

void load_block(const uint8_t *p, size_t n, uint32_t m[16]) {
    uint8_t tmp[64] = {0};
    memcpy(tmp, p, n);
    for (int i = 0; i < 16; i++)
        m[i] = (uint32_t)tmp[4 * i] | ((uint32_t)tmp[4 * i + 1] << 8) | ((uint32_t)tmp[4 * i + 2] << 16) | ((uint32_t)tmp[4 * i + 3] << 24);
}

}  // namespace

bool blake3_single_chunk(const uint8_t *msg, size_t len, uint8_t *out, size_t outlen) {
    if (len > 1024) return false;
    uint32_t cv[8], m[16], full[16];
    for (int i = 0; i < 8; i++) cv[i] = kIV[i];
    uint32_t flags = kChunkStart;
    size_t off = 0;
    while (len - off > 64) {
        load_block(msg + off, 64, m);
        compress(cv, m, 0, 64, flags, full);
        memcpy(cv, full, 32);
        flags = 0;
        off += 64;
    }
    const uint32_t last_len = (uint32_t)(len - off);
    load_block(msg + off, last_len, m);
    flags |= kChunkEnd | kRoot;
    for (uint64_t ctr = 0; outlen; ctr++) {
        compress(cv, m, ctr, last_len, flags, full);
        uint8_t bytes[64];
        for (int i = 0; i < 16; i++) {
            bytes[4 * i] = (uint8_t)full[i]; bytes[4 * i + 1] = (uint8_t)(full[i] >> 8);
            bytes[4 * i + 2] = (uint8_t)(full[i] >> 16); bytes[4 * i + 3] = (uint8_t)(full[i] >> 24);
        }
        const size_t take = outlen < 64 ? outlen : 64;
        memcpy(out, bytes, take);
        out += take; outlen -= take;
    }
    return true;
}

void commitment_bytes(const uint8_t node_id[32], const uint8_t commitment_atx[32], uint8_t out[32]) {
    uint8_t buf[64];
    memcpy(buf, node_id, 32);
    memcpy(buf + 32, commitment_atx, 32);
    blake3_single_chunk(buf, 64, out, 32);
}

void vrf_difficulty(uint64_t num_labels, uint8_t out[32]) {
    if (num_labels <= 1) { memset(out, 0xff, 32); return; }
    // schoolbook division of the 33-byte number 0x01 00..00 by a 64-bit divisor
    unsigned __int128 rem = 1;
    for (int i = 0; i < 32; i++) {
        rem <<= 8;
        out[i] = (uint8_t)(rem / num_labels);
        rem %= num_labels;
    }
}

}  // namespace b200post

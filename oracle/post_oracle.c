/*
 * post_oracle.c — CPU ORACLE (test infrastructure, see post_oracle.h for the parity status).
 *
 * Follows, by specification section rather than by any source file (the reference tree
 * contains none of this arithmetic — SURVEY.md "Three facts" #1):
 *   SHA-256 ............ FIPS 180-4 §6.2
 *   HMAC ............... FIPS 198-1 §4
 *   PBKDF2 ............. RFC 8018 §5.2
 *   Salsa20/8, BlockMix, ROMix, scrypt ... RFC 7914 §3-§6 (kept as a KAT-pinned building block)
 *   Keccak-f[1600], Keccak-512 ........... the Keccak submission (original 0x01 padding, not SHA-3's 0x06)
 *   ChaCha20/8 core ...................... D. J. Bernstein, "ChaCha, a variant of Salsa20" (8 rounds, no constants)
 *   scrypt-jane (ChaCha20/8 + Keccak-512) . floodyberry/scrypt-jane as libpost builds it: the POST LABEL FUNCTION.
 *       Pinned by real data: the 42 identities of the reference's checkpoint/checkpointdata.json carry VRF
 *       nonces that are the arg-min label of their POST under exactly this function (tests/golden/
 *       checkpoint_vrf.json, tools/pin_search.py) and under no RFC 7914 variant.
 *   BLAKE3 ............. BLAKE3 paper §2 (zeebo/blake3 v0.2.4 is what hash/hash.go:16-25 calls)
 *   AES-128 ............ FIPS-197
 * Call-site anchors in the reference: activation/post.go:295,355-361 (init),
 * activation/post_verifier.go:159 (verify), activation/validation.go:261-282 (VRF nonce).
 */
#define _POSIX_C_SOURCE 200809L
#include "post_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ======================================================================== SHA-256 */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

typedef struct {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t total;
} sha256_ctx;

static inline uint32_t ror32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline uint32_t rol32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
static inline uint32_t be32(const uint8_t *p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
static inline void put_be32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}
static inline uint32_t le32(const uint8_t *p) {
    return ((uint32_t)p[3] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0];
}
static inline void put_le32(uint8_t *p, uint32_t v) {
    p[3] = (uint8_t)(v >> 24); p[2] = (uint8_t)(v >> 16); p[1] = (uint8_t)(v >> 8); p[0] = (uint8_t)v;
}

static void sha256_compress(uint32_t h[8], const uint8_t blk[64]) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = be32(blk + 4 * i);
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = ror32(w[i - 15], 7) ^ ror32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = ror32(w[i - 2], 17) ^ ror32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = ror32(e, 6) ^ ror32(e, 11) ^ ror32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
        uint32_t S0 = ror32(a, 2) ^ ror32(a, 13) ^ ror32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

static void sha256_init(sha256_ctx *c) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                   0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(c->h, iv, sizeof iv);
    c->total = 0;
}
static void sha256_update(sha256_ctx *c, const uint8_t *p, size_t n) {
    size_t fill = (size_t)(c->total & 63);
    c->total += n;
    if (fill) {
        size_t take = 64 - fill;
        if (take > n) take = n;
        memcpy(c->buf + fill, p, take);
        p += take; n -= take; fill += take;
        if (fill < 64) return;
        sha256_compress(c->h, c->buf);
    }
    while (n >= 64) { sha256_compress(c->h, p); p += 64; n -= 64; }
    if (n) memcpy(c->buf, p, n);
}
static void sha256_final(sha256_ctx *c, uint8_t out[32]) {
    uint64_t bits = c->total * 8;
    uint8_t pad[72];
    size_t fill = (size_t)(c->total & 63);
    size_t padlen = (fill < 56) ? (56 - fill) : (120 - fill);
    memset(pad, 0, sizeof pad);
    pad[0] = 0x80;
    for (int i = 0; i < 8; i++) pad[padlen + i] = (uint8_t)(bits >> (56 - 8 * i));
    sha256_update(c, pad, padlen + 8);
    for (int i = 0; i < 8; i++) put_be32(out + 4 * i, c->h[i]);
}

void oracle_sha256(const uint8_t *msg, size_t len, uint8_t out[32]) {
    sha256_ctx c;
    sha256_init(&c);
    sha256_update(&c, msg, len);
    sha256_final(&c, out);
}

/* ======================================================================== HMAC / PBKDF2 */
typedef struct { sha256_ctx inner, outer; } hmac_ctx;

static void hmac_init(hmac_ctx *h, const uint8_t *key, size_t klen) {
    uint8_t k[64], pad[64];
    memset(k, 0, 64);
    if (klen > 64) oracle_sha256(key, klen, k); else memcpy(k, key, klen);
    for (int i = 0; i < 64; i++) pad[i] = k[i] ^ 0x36;
    sha256_init(&h->inner); sha256_update(&h->inner, pad, 64);
    for (int i = 0; i < 64; i++) pad[i] = k[i] ^ 0x5c;
    sha256_init(&h->outer); sha256_update(&h->outer, pad, 64);
}
static void hmac_final(hmac_ctx *h, uint8_t out[32]) {
    uint8_t d[32];
    sha256_final(&h->inner, d);
    sha256_update(&h->outer, d, 32);
    sha256_final(&h->outer, out);
}

void oracle_hmac_sha256(const uint8_t *key, size_t klen, const uint8_t *msg, size_t mlen, uint8_t out[32]) {
    hmac_ctx h;
    hmac_init(&h, key, klen);
    sha256_update(&h.inner, msg, mlen);
    hmac_final(&h, out);
}

void oracle_pbkdf2_sha256(const uint8_t *pw, size_t pwlen, const uint8_t *salt, size_t saltlen,
                          uint32_t iters, uint8_t *out, size_t dklen) {
    hmac_ctx base;
    hmac_init(&base, pw, pwlen);
    uint32_t blk = 1;
    while (dklen) {
        uint8_t u[32], t[32], ctr[4];
        hmac_ctx h = base;
        put_be32(ctr, blk);
        sha256_update(&h.inner, salt, saltlen);
        sha256_update(&h.inner, ctr, 4);
        hmac_final(&h, u);
        memcpy(t, u, 32);
        for (uint32_t i = 1; i < iters; i++) {
            h = base;
            sha256_update(&h.inner, u, 32);
            hmac_final(&h, u);
            for (int j = 0; j < 32; j++) t[j] ^= u[j];
        }
        size_t take = dklen < 32 ? dklen : 32;
        memcpy(out, t, take);
        out += take; dklen -= take; blk++;
    }
}

/* ======================================================================== scrypt */
void oracle_salsa20_8(uint32_t b[16]) {
    uint32_t x[16];
    memcpy(x, b, 64);
#define QR(a, b_, c, d) \
    x[b_] ^= rol32(x[a] + x[d], 7); x[c] ^= rol32(x[b_] + x[a], 9); \
    x[d] ^= rol32(x[c] + x[b_], 13); x[a] ^= rol32(x[d] + x[c], 18);
    for (int i = 0; i < 4; i++) {
        /* column round */
        QR(0, 4, 8, 12) QR(5, 9, 13, 1) QR(10, 14, 2, 6) QR(15, 3, 7, 11)
        /* row round */
        QR(0, 1, 2, 3) QR(5, 6, 7, 4) QR(10, 11, 8, 9) QR(15, 12, 13, 14)
    }
#undef QR
    for (int i = 0; i < 16; i++) b[i] += x[i];
}

/* ChaCha20/8 core as scrypt-jane uses it (scrypt-jane-mix_chacha.h chacha_core_basic): the 64-byte block is
 * the whole state (no constants / counter), 4 double rounds, feed-forward add. */
void oracle_chacha20_8(uint32_t b[16]) {
    uint32_t x[16];
    memcpy(x, b, 64);
#define QR(a, b_, c, d) \
    x[a] += x[b_]; x[d] = rol32(x[d] ^ x[a], 16); x[c] += x[d]; x[b_] = rol32(x[b_] ^ x[c], 12); \
    x[a] += x[b_]; x[d] = rol32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b_] = rol32(x[b_] ^ x[c], 7);
    for (int i = 0; i < 4; i++) {
        QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)     /* columns */
        QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)     /* diagonals */
    }
#undef QR
    for (int i = 0; i < 16; i++) b[i] += x[i];
}

typedef void (*mix_core_fn)(uint32_t *);

static void blockmix_with(uint32_t *b, uint32_t *y, uint32_t r, mix_core_fn core) {
    uint32_t x[16];
    memcpy(x, b + (2 * r - 1) * 16, 64);
    for (uint32_t i = 0; i < 2 * r; i++) {
        for (int k = 0; k < 16; k++) x[k] ^= b[i * 16 + k];
        core(x);
        /* even blocks to the first half, odd blocks to the second (RFC 7914 §4 step 3) */
        memcpy(y + ((i & 1) ? (r + i / 2) : (i / 2)) * 16, x, 64);
    }
    memcpy(b, y, 128 * (size_t)r);
}
void oracle_blockmix(uint32_t *b, uint32_t *y, uint32_t r) { blockmix_with(b, y, r, oracle_salsa20_8); }

static void romix(uint32_t *x, uint32_t *v, uint32_t *y, uint64_t N, uint32_t r, mix_core_fn core) {
    const size_t words = 32 * (size_t)r;
    for (uint64_t i = 0; i < N; i++) {
        memcpy(v + i * words, x, words * 4);
        blockmix_with(x, y, r, core);
    }
    for (uint64_t i = 0; i < N; i++) {
        /* Integerify: first 8 bytes of the last 64-byte sub-block, little-endian, mod N */
        uint64_t j = ((uint64_t)x[(2 * r - 1) * 16] | ((uint64_t)x[(2 * r - 1) * 16 + 1] << 32)) & (N - 1);
        const uint32_t *vj = v + j * words;
        for (size_t k = 0; k < words; k++) x[k] ^= vj[k];
        blockmix_with(x, y, r, core);
    }
}

int oracle_scrypt(const uint8_t *pw, size_t pwlen, const uint8_t *salt, size_t saltlen,
                  uint64_t N, uint32_t r, uint32_t p, uint8_t *out, size_t dklen) {
    if (N < 2 || (N & (N - 1)) || r == 0 || p == 0) return -1;
    const size_t blk = 128 * (size_t)r;
    uint8_t *B = (uint8_t *)malloc(blk * p);
    uint32_t *x = (uint32_t *)malloc(blk), *y = (uint32_t *)malloc(blk);
    uint32_t *v = (uint32_t *)malloc(blk * N);
    if (!B || !x || !y || !v) { free(B); free(x); free(y); free(v); return -1; }
    oracle_pbkdf2_sha256(pw, pwlen, salt, saltlen, 1, B, blk * p);
    for (uint32_t i = 0; i < p; i++) {
        for (size_t k = 0; k < blk / 4; k++) x[k] = le32(B + i * blk + 4 * k);
        romix(x, v, y, N, r, oracle_salsa20_8);
        for (size_t k = 0; k < blk / 4; k++) put_le32(B + i * blk + 4 * k, x[k]);
    }
    oracle_pbkdf2_sha256(pw, pwlen, B, blk * p, 1, out, dklen);
    free(B); free(x); free(y); free(v);
    return 0;
}

/* ======================================================================== Keccak-512 / scrypt-jane */
static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
    0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
static inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

/* Keccak-f[1600], lane (x, y) at s[x + 5 y]; step by step as in the specification (theta, rho, pi, chi, iota) */
static void keccak_f1600(uint64_t s[25]) {
    static const int RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int round = 0; round < 24; round++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) s[i] ^= d[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(s[x + 5 * y], RHO[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) s[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        s[0] ^= KECCAK_RC[round];
    }
}

#define K512_RATE 72
typedef struct { uint64_t s[25]; uint8_t buf[K512_RATE]; size_t fill; } k512_ctx;
static void k512_init(k512_ctx *c) { memset(c, 0, sizeof *c); }
static void k512_block(k512_ctx *c, const uint8_t *p) {
    for (int i = 0; i < K512_RATE / 8; i++) {
        uint64_t w = 0;
        for (int k = 0; k < 8; k++) w |= (uint64_t)p[8 * i + k] << (8 * k);
        c->s[i] ^= w;
    }
    keccak_f1600(c->s);
}
static void k512_update(k512_ctx *c, const uint8_t *p, size_t n) {
    while (n) {
        size_t take = K512_RATE - c->fill;
        if (take > n) take = n;
        memcpy(c->buf + c->fill, p, take);
        c->fill += take; p += take; n -= take;
        if (c->fill == K512_RATE) { k512_block(c, c->buf); c->fill = 0; }
    }
}
static void k512_final(k512_ctx *c, uint8_t pad, uint8_t out[64]) {
    memset(c->buf + c->fill, 0, K512_RATE - c->fill);
    c->buf[c->fill] ^= pad;
    c->buf[K512_RATE - 1] ^= 0x80;
    k512_block(c, c->buf);
    for (int i = 0; i < 8; i++)
        for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(c->s[i] >> (8 * k));
}
/* pad = 0x01: original Keccak (scrypt-jane's SCRYPT_KECCAK512); pad = 0x06 gives SHA3-512 (used to pin the
 * permutation against hashlib in tests) */
void oracle_keccak512(const uint8_t *msg, size_t len, uint8_t pad, uint8_t out[64]) {
    k512_ctx c;
    k512_init(&c);
    k512_update(&c, msg, len);
    k512_final(&c, pad, out);
}

/* HMAC over Keccak-512: block size = the sponge rate, 72 bytes (scrypt-jane-hash_keccak.h SCRYPT_HASH_BLOCK_SIZE) */
typedef struct { k512_ctx inner, outer; } khmac_ctx;
static void khmac_init(khmac_ctx *h, const uint8_t *key, size_t klen) {
    uint8_t k[K512_RATE], pad[K512_RATE], kd[64];
    memset(k, 0, sizeof k);
    if (klen > K512_RATE) { oracle_keccak512(key, klen, 0x01, kd); memcpy(k, kd, 64); } else memcpy(k, key, klen);
    for (int i = 0; i < K512_RATE; i++) pad[i] = k[i] ^ 0x36;
    k512_init(&h->inner); k512_update(&h->inner, pad, K512_RATE);
    for (int i = 0; i < K512_RATE; i++) pad[i] = k[i] ^ 0x5c;
    k512_init(&h->outer); k512_update(&h->outer, pad, K512_RATE);
}
static void khmac_final(khmac_ctx *h, uint8_t out[64]) {
    uint8_t d[64];
    k512_final(&h->inner, 0x01, d);
    k512_update(&h->outer, d, 64);
    k512_final(&h->outer, 0x01, out);
}
void oracle_hmac_keccak512(const uint8_t *key, size_t klen, const uint8_t *msg, size_t mlen, uint8_t out[64]) {
    khmac_ctx h;
    khmac_init(&h, key, klen);
    k512_update(&h.inner, msg, mlen);
    khmac_final(&h, out);
}
/* PBKDF2 (RFC 8018 §5.2) with that HMAC, one iteration: all scrypt needs */
void oracle_pbkdf2_keccak512(const uint8_t *pw, size_t pwlen, const uint8_t *salt, size_t saltlen, uint8_t *out, size_t dklen) {
    khmac_ctx base;
    khmac_init(&base, pw, pwlen);
    uint32_t blk = 1;
    while (dklen) {
        uint8_t t[64], ctr[4];
        khmac_ctx h = base;
        put_be32(ctr, blk);
        k512_update(&h.inner, salt, saltlen);
        k512_update(&h.inner, ctr, 4);
        khmac_final(&h, t);
        size_t take = dklen < 64 ? dklen : 64;
        memcpy(out, t, take);
        out += take; dklen -= take; blk++;
    }
}

/* scrypt-jane with SCRYPT_CHACHA + SCRYPT_KECCAK512: scrypt's structure (PBKDF2 -> ROMix per lane -> PBKDF2) with
 * ChaCha20/8 as the BlockMix core and HMAC-Keccak-512 inside PBKDF2. */
int oracle_scrypt_jane(const uint8_t *pw, size_t pwlen, const uint8_t *salt, size_t saltlen,
                       uint64_t N, uint32_t r, uint32_t p, uint8_t *out, size_t dklen) {
    if (N < 2 || (N & (N - 1)) || r == 0 || p == 0) return -1;
    const size_t blk = 128 * (size_t)r;
    uint8_t *B = (uint8_t *)malloc(blk * p);
    uint32_t *x = (uint32_t *)malloc(blk), *y = (uint32_t *)malloc(blk);
    uint32_t *v = (uint32_t *)malloc(blk * N);
    if (!B || !x || !y || !v) { free(B); free(x); free(y); free(v); return -1; }
    oracle_pbkdf2_keccak512(pw, pwlen, salt, saltlen, B, blk * p);
    for (uint32_t i = 0; i < p; i++) {
        for (size_t k = 0; k < blk / 4; k++) x[k] = le32(B + i * blk + 4 * k);
        romix(x, v, y, N, r, oracle_chacha20_8);
        for (size_t k = 0; k < blk / 4; k++) put_le32(B + i * blk + 4 * k, x[k]);
    }
    oracle_pbkdf2_keccak512(pw, pwlen, B, blk * p, out, dklen);
    free(B); free(x); free(y); free(v);
    return 0;
}

/* ======================================================================== BLAKE3 */
static const uint32_t B3_IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A,
                                  0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
static const uint8_t B3_PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_PARENT = 4, B3_ROOT = 8 };

static void b3_g(uint32_t *s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    s[a] = s[a] + s[b] + mx; s[d] = ror32(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];      s[b] = ror32(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my; s[d] = ror32(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];      s[b] = ror32(s[b] ^ s[c], 7);
}
/* full 16-word compression output */
static void b3_compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter,
                        uint32_t block_len, uint32_t flags, uint32_t out[16]) {
    uint32_t s[16], m[16], t[16];
    memcpy(s, cv, 32);
    memcpy(s + 8, B3_IV, 16);
    s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = block_len; s[15] = flags;
    memcpy(m, block, 64);
    for (int r = 0; r < 7; r++) {
        b3_g(s, 0, 4, 8, 12, m[0], m[1]);   b3_g(s, 1, 5, 9, 13, m[2], m[3]);
        b3_g(s, 2, 6, 10, 14, m[4], m[5]);  b3_g(s, 3, 7, 11, 15, m[6], m[7]);
        b3_g(s, 0, 5, 10, 15, m[8], m[9]);  b3_g(s, 1, 6, 11, 12, m[10], m[11]);
        b3_g(s, 2, 7, 8, 13, m[12], m[13]); b3_g(s, 3, 4, 9, 14, m[14], m[15]);
        for (int i = 0; i < 16; i++) t[i] = m[B3_PERM[i]];
        memcpy(m, t, 64);
    }
    for (int i = 0; i < 8; i++) { out[i] = s[i] ^ s[i + 8]; out[i + 8] = s[i + 8] ^ cv[i]; }
}

/* An "output" node: everything needed to finalise either as a chaining value or as root. */
typedef struct { uint32_t cv[8]; uint32_t block[16]; uint64_t counter; uint32_t block_len, flags; } b3_output;

static void b3_words(const uint8_t *p, size_t n, uint32_t w[16]) {
    uint8_t tmp[64];
    memset(tmp, 0, 64);
    memcpy(tmp, p, n);
    for (int i = 0; i < 16; i++) w[i] = le32(tmp + 4 * i);
}
/* hash one chunk (<= 1024 bytes) up to, but not including, its final compression */
static b3_output b3_chunk(const uint8_t *p, size_t n, uint64_t chunk_counter) {
    b3_output o;
    uint32_t cv[8], full[16];
    memcpy(cv, B3_IV, 32);
    uint32_t flags = B3_CHUNK_START;
    while (n > 64) {
        b3_words(p, 64, o.block);
        b3_compress(cv, o.block, chunk_counter, 64, flags, full);
        memcpy(cv, full, 32);
        flags = 0; p += 64; n -= 64;
    }
    memcpy(o.cv, cv, 32);
    b3_words(p, n, o.block);
    o.counter = chunk_counter; o.block_len = (uint32_t)n; o.flags = flags | B3_CHUNK_END;
    return o;
}
static void b3_output_cv(const b3_output *o, uint32_t cv[8]) {
    uint32_t full[16];
    b3_compress(o->cv, o->block, o->counter, o->block_len, o->flags, full);
    memcpy(cv, full, 32);
}
static b3_output b3_parent(const uint32_t l[8], const uint32_t r[8]) {
    b3_output o;
    memcpy(o.cv, B3_IV, 32);
    memcpy(o.block, l, 32); memcpy(o.block + 8, r, 32);
    o.counter = 0; o.block_len = 64; o.flags = B3_PARENT;
    return o;
}

void oracle_blake3_xof(const uint8_t *msg, size_t len, uint8_t *out, size_t outlen) {
    uint32_t stack[54][8];
    int sp = 0;
    uint64_t chunk = 0;
    b3_output cur;
    for (;;) {
        size_t take = len > 1024 ? 1024 : len;
        cur = b3_chunk(msg, take, chunk);
        msg += take; len -= take;
        if (len == 0) break;
        /* not the last chunk: push its CV, merging completed subtrees */
        uint32_t cv[8];
        b3_output_cv(&cur, cv);
        chunk++;
        uint64_t total = chunk;
        while ((total & 1) == 0) {
            b3_output par = b3_parent(stack[--sp], cv);
            b3_output_cv(&par, cv);
            total >>= 1;
        }
        memcpy(stack[sp++], cv, 32);
    }
    while (sp > 0) {
        uint32_t cv[8];
        b3_output_cv(&cur, cv);
        cur = b3_parent(stack[--sp], cv);
    }
    uint64_t ctr = 0;
    while (outlen) {
        uint32_t full[16];
        uint8_t bytes[64];
        b3_compress(cur.cv, cur.block, ctr++, cur.block_len, cur.flags | B3_ROOT, full);
        for (int i = 0; i < 16; i++) put_le32(bytes + 4 * i, full[i]);
        size_t take = outlen < 64 ? outlen : 64;
        memcpy(out, bytes, take);
        out += take; outlen -= take;
    }
}
void oracle_blake3_256(const uint8_t *msg, size_t len, uint8_t out[32]) { oracle_blake3_xof(msg, len, out, 32); }

/* ======================================================================== AES-128 */
static uint8_t AES_SBOX[256];
static int aes_ready;
static uint8_t gf_mul(uint8_t a, uint8_t b) {
    uint8_t p = 0;
    while (b) { if (b & 1) p ^= a; a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1b : 0)); b >>= 1; }
    return p;
}
static void aes_init_sbox(void) {
    /* S(x) = affine(x^-1) over GF(2^8)/0x11b (FIPS-197 §5.1.1) */
    for (int x = 0; x < 256; x++) {
        uint8_t inv = 0;
        if (x) for (int y = 1; y < 256; y++) if (gf_mul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
        uint8_t s = inv, r = inv;
        for (int i = 0; i < 4; i++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
        AES_SBOX[x] = s ^ 0x63;
    }
    aes_ready = 1;
}
void oracle_aes128_encrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16]) {
    if (!aes_ready) aes_init_sbox();
    uint8_t rk[176], s[16];
    memcpy(rk, key, 16);
    uint8_t rcon = 1;
    for (int i = 16; i < 176; i += 4) {
        uint8_t t[4] = {rk[i - 4], rk[i - 3], rk[i - 2], rk[i - 1]};
        if (i % 16 == 0) {
            uint8_t u = t[0];
            t[0] = AES_SBOX[t[1]] ^ rcon; t[1] = AES_SBOX[t[2]]; t[2] = AES_SBOX[t[3]]; t[3] = AES_SBOX[u];
            rcon = (uint8_t)((rcon << 1) ^ ((rcon & 0x80) ? 0x1b : 0));
        }
        for (int k = 0; k < 4; k++) rk[i + k] = rk[i - 16 + k] ^ t[k];
    }
    for (int i = 0; i < 16; i++) s[i] = in[i] ^ rk[i];
    for (int round = 1; round <= 10; round++) {
        uint8_t t[16];
        for (int i = 0; i < 16; i++) t[i] = AES_SBOX[s[i]];
        /* ShiftRows: state is column-major, byte index = 4*col + row */
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) s[4 * c + r] = t[4 * ((c + r) & 3) + r];
        if (round != 10) {
            for (int c = 0; c < 4; c++) {
                uint8_t a0 = s[4 * c], a1 = s[4 * c + 1], a2 = s[4 * c + 2], a3 = s[4 * c + 3];
                s[4 * c + 0] = (uint8_t)(gf_mul(a0, 2) ^ gf_mul(a1, 3) ^ a2 ^ a3);
                s[4 * c + 1] = (uint8_t)(a0 ^ gf_mul(a1, 2) ^ gf_mul(a2, 3) ^ a3);
                s[4 * c + 2] = (uint8_t)(a0 ^ a1 ^ gf_mul(a2, 2) ^ gf_mul(a3, 3));
                s[4 * c + 3] = (uint8_t)(gf_mul(a0, 3) ^ a1 ^ a2 ^ gf_mul(a3, 2));
            }
        }
        for (int i = 0; i < 16; i++) s[i] ^= rk[16 * round + i];
    }
    memcpy(out, s, 16);
}

/* ======================================================================== SSE2 / AVX2 ROMix (r = 1) */
/* SSE2 ROMix used for the CPU baseline (the reference's CPU path, scrypt-jane inside libpost, is SIMD code too).
 * ChaCha20/8 on four 128-bit rows in natural word order: the column round works on whole vectors, the diagonal
 * round needs three lane rotations.  Cross-checked against the scalar path in tests. */
#if defined(__SSE2__)
#include <emmintrin.h>
#define ORACLE_HAVE_SSE2 1
static inline void chacha20_8_sse(__m128i B[4]) {
    __m128i a = B[0], b = B[1], c = B[2], d = B[3], T;
#define ROTV(v, k) T = v; v = _mm_or_si128(_mm_slli_epi32(T, k), _mm_srli_epi32(T, 32 - (k)));
#define HALF \
    a = _mm_add_epi32(a, b); d = _mm_xor_si128(d, a); ROTV(d, 16) \
    c = _mm_add_epi32(c, d); b = _mm_xor_si128(b, c); ROTV(b, 12) \
    a = _mm_add_epi32(a, b); d = _mm_xor_si128(d, a); ROTV(d, 8)  \
    c = _mm_add_epi32(c, d); b = _mm_xor_si128(b, c); ROTV(b, 7)
    for (int i = 0; i < 4; i++) {
        HALF                                                                      /* columns */
        b = _mm_shuffle_epi32(b, 0x39); c = _mm_shuffle_epi32(c, 0x4E); d = _mm_shuffle_epi32(d, 0x93);
        HALF                                                                      /* diagonals */
        b = _mm_shuffle_epi32(b, 0x93); c = _mm_shuffle_epi32(c, 0x4E); d = _mm_shuffle_epi32(d, 0x39);
    }
#undef HALF
#undef ROTV
    B[0] = _mm_add_epi32(B[0], a); B[1] = _mm_add_epi32(B[1], b);
    B[2] = _mm_add_epi32(B[2], c); B[3] = _mm_add_epi32(B[3], d);
}
/* BlockMix for r = 1 on X = (lo, hi), 8 vectors, in place */
static inline void blockmix_r1_sse(__m128i X[8]) {
    __m128i T[4];
    for (int k = 0; k < 4; k++) T[k] = _mm_xor_si128(X[k], X[4 + k]);
    chacha20_8_sse(T);
    for (int k = 0; k < 4; k++) { X[k] = T[k]; T[k] = _mm_xor_si128(T[k], X[4 + k]); }
    chacha20_8_sse(T);
    for (int k = 0; k < 4; k++) X[4 + k] = T[k];
}
static void romix_r1_sse(uint32_t x[32], void *vmem, uint64_t N) {
    __m128i X[8];
    __m128i *V = (__m128i *)vmem;
    for (int k = 0; k < 8; k++) X[k] = _mm_loadu_si128((const __m128i *)(x + 4 * k));
    for (uint64_t i = 0; i < N; i++) {
        for (int k = 0; k < 8; k++) _mm_storeu_si128(V + 8 * i + k, X[k]);
        blockmix_r1_sse(X);
    }
    for (uint64_t i = 0; i < N; i++) {
        const uint64_t j = (uint32_t)_mm_cvtsi128_si32(X[4]) & (N - 1);
        for (int k = 0; k < 8; k++) X[k] = _mm_xor_si128(X[k], _mm_loadu_si128(V + 8 * j + k));
        blockmix_r1_sse(X);
    }
    for (int k = 0; k < 8; k++) _mm_storeu_si128((__m128i *)(x + 4 * k), X[k]);
}
#if defined(__AVX2__)
#include <immintrin.h>
#define ORACLE_HAVE_AVX2 1
/* Two labels per thread: the low 128-bit lane of every vector belongs to label A, the high lane to label B (all AVX2
 * integer ops used here work per lane).  16- and 8-bit rotates are one byte shuffle. */
static inline void chacha20_8_avx2(__m256i B[4]) {
    const __m256i R16 = _mm256_setr_epi8(2, 3, 0, 1, 6, 7, 4, 5, 10, 11, 8, 9, 14, 15, 12, 13, 2, 3, 0, 1, 6, 7, 4, 5, 10, 11, 8, 9, 14, 15, 12, 13);
    const __m256i R8 = _mm256_setr_epi8(3, 0, 1, 2, 7, 4, 5, 6, 11, 8, 9, 10, 15, 12, 13, 14, 3, 0, 1, 2, 7, 4, 5, 6, 11, 8, 9, 10, 15, 12, 13, 14);
    __m256i a = B[0], b = B[1], c = B[2], d = B[3];
#define ROTS(v, k) v = _mm256_or_si256(_mm256_slli_epi32(v, k), _mm256_srli_epi32(v, 32 - (k)));
#define HALF2 \
    a = _mm256_add_epi32(a, b); d = _mm256_shuffle_epi8(_mm256_xor_si256(d, a), R16); \
    c = _mm256_add_epi32(c, d); b = _mm256_xor_si256(b, c); ROTS(b, 12) \
    a = _mm256_add_epi32(a, b); d = _mm256_shuffle_epi8(_mm256_xor_si256(d, a), R8); \
    c = _mm256_add_epi32(c, d); b = _mm256_xor_si256(b, c); ROTS(b, 7)
    for (int i = 0; i < 4; i++) {
        HALF2
        b = _mm256_shuffle_epi32(b, 0x39); c = _mm256_shuffle_epi32(c, 0x4E); d = _mm256_shuffle_epi32(d, 0x93);
        HALF2
        b = _mm256_shuffle_epi32(b, 0x93); c = _mm256_shuffle_epi32(c, 0x4E); d = _mm256_shuffle_epi32(d, 0x39);
    }
#undef HALF2
#undef ROTS
    B[0] = _mm256_add_epi32(B[0], a); B[1] = _mm256_add_epi32(B[1], b);
    B[2] = _mm256_add_epi32(B[2], c); B[3] = _mm256_add_epi32(B[3], d);
}
static inline void blockmix_r1_avx2(__m256i X[8]) {
    __m256i T[4];
    for (int k = 0; k < 4; k++) T[k] = _mm256_xor_si256(X[k], X[4 + k]);
    chacha20_8_avx2(T);
    for (int k = 0; k < 4; k++) { X[k] = T[k]; T[k] = _mm256_xor_si256(T[k], X[4 + k]); }
    chacha20_8_avx2(T);
    for (int k = 0; k < 4; k++) X[4 + k] = T[k];
}
/* ROMix of labels A and B in lock-step; va / vb are their private 128*N-byte scratchpads */
static void romix_r1_avx2_x2(uint32_t xa[32], uint32_t xb[32], void *va, void *vb, uint64_t N) {
    __m256i X[8];
    __m128i *VA = (__m128i *)va, *VB = (__m128i *)vb;
    for (int k = 0; k < 8; k++)
        X[k] = _mm256_set_m128i(_mm_loadu_si128((const __m128i *)(xb + 4 * k)), _mm_loadu_si128((const __m128i *)(xa + 4 * k)));
    for (uint64_t i = 0; i < N; i++) {
        for (int k = 0; k < 8; k++) {
            _mm_storeu_si128(VA + 8 * i + k, _mm256_castsi256_si128(X[k]));
            _mm_storeu_si128(VB + 8 * i + k, _mm256_extracti128_si256(X[k], 1));
        }
        blockmix_r1_avx2(X);
    }
    for (uint64_t i = 0; i < N; i++) {
        const uint64_t ja = (uint32_t)_mm_cvtsi128_si32(_mm256_castsi256_si128(X[4])) & (N - 1);
        const uint64_t jb = (uint32_t)_mm_cvtsi128_si32(_mm256_extracti128_si256(X[4], 1)) & (N - 1);
        for (int k = 0; k < 8; k++)
            X[k] = _mm256_xor_si256(X[k], _mm256_set_m128i(_mm_loadu_si128(VB + 8 * jb + k), _mm_loadu_si128(VA + 8 * ja + k)));
        blockmix_r1_avx2(X);
    }
    for (int k = 0; k < 8; k++) {
        _mm_storeu_si128((__m128i *)(xa + 4 * k), _mm256_castsi256_si128(X[k]));
        _mm_storeu_si128((__m128i *)(xb + 4 * k), _mm256_extracti128_si256(X[k], 1));
    }
}

/* Four labels per thread in the four 128-bit lanes of AVX-512 registers (compiled per function, chosen at run time) */
#define T512 __attribute__((target("avx512f")))
T512 static inline void chacha20_8_avx512(__m512i B[4]) {
    __m512i a = B[0], b = B[1], c = B[2], d = B[3];
#define HALF4 \
    a = _mm512_add_epi32(a, b); d = _mm512_rol_epi32(_mm512_xor_si512(d, a), 16); \
    c = _mm512_add_epi32(c, d); b = _mm512_rol_epi32(_mm512_xor_si512(b, c), 12); \
    a = _mm512_add_epi32(a, b); d = _mm512_rol_epi32(_mm512_xor_si512(d, a), 8);  \
    c = _mm512_add_epi32(c, d); b = _mm512_rol_epi32(_mm512_xor_si512(b, c), 7);
    for (int i = 0; i < 4; i++) {
        HALF4
        b = _mm512_shuffle_epi32(b, 0x39); c = _mm512_shuffle_epi32(c, 0x4E); d = _mm512_shuffle_epi32(d, 0x93);
        HALF4
        b = _mm512_shuffle_epi32(b, 0x93); c = _mm512_shuffle_epi32(c, 0x4E); d = _mm512_shuffle_epi32(d, 0x39);
    }
#undef HALF4
    B[0] = _mm512_add_epi32(B[0], a); B[1] = _mm512_add_epi32(B[1], b);
    B[2] = _mm512_add_epi32(B[2], c); B[3] = _mm512_add_epi32(B[3], d);
}
T512 static inline void blockmix_r1_avx512(__m512i X[8]) {
    __m512i T[4];
    for (int k = 0; k < 4; k++) T[k] = _mm512_xor_si512(X[k], X[4 + k]);
    chacha20_8_avx512(T);
    for (int k = 0; k < 4; k++) { X[k] = T[k]; T[k] = _mm512_xor_si512(T[k], X[4 + k]); }
    chacha20_8_avx512(T);
    for (int k = 0; k < 4; k++) X[4 + k] = T[k];
}
T512 static inline __m512i gather4(const __m128i *p0, const __m128i *p1, const __m128i *p2, const __m128i *p3) {
    __m512i v = _mm512_castsi128_si512(_mm_loadu_si128(p0));
    v = _mm512_inserti32x4(v, _mm_loadu_si128(p1), 1);
    v = _mm512_inserti32x4(v, _mm_loadu_si128(p2), 2);
    return _mm512_inserti32x4(v, _mm_loadu_si128(p3), 3);
}
/* ROMix of four labels in lock-step; v[q] = label q's private 128*N-byte scratchpad */
T512 static void romix_r1_avx512_x4(uint32_t *x[4], void *v[4], uint64_t N) {
    __m512i X[8];
    __m128i *V0 = (__m128i *)v[0], *V1 = (__m128i *)v[1], *V2 = (__m128i *)v[2], *V3 = (__m128i *)v[3];
    for (int k = 0; k < 8; k++)
        X[k] = gather4((const __m128i *)(x[0] + 4 * k), (const __m128i *)(x[1] + 4 * k), (const __m128i *)(x[2] + 4 * k), (const __m128i *)(x[3] + 4 * k));
    for (uint64_t i = 0; i < N; i++) {
        for (int k = 0; k < 8; k++) {
            _mm_storeu_si128(V0 + 8 * i + k, _mm512_castsi512_si128(X[k]));
            _mm_storeu_si128(V1 + 8 * i + k, _mm512_extracti32x4_epi32(X[k], 1));
            _mm_storeu_si128(V2 + 8 * i + k, _mm512_extracti32x4_epi32(X[k], 2));
            _mm_storeu_si128(V3 + 8 * i + k, _mm512_extracti32x4_epi32(X[k], 3));
        }
        blockmix_r1_avx512(X);
    }
    for (uint64_t i = 0; i < N; i++) {
        const uint64_t j0 = (uint32_t)_mm_cvtsi128_si32(_mm512_castsi512_si128(X[4])) & (N - 1);
        const uint64_t j1 = (uint32_t)_mm_cvtsi128_si32(_mm512_extracti32x4_epi32(X[4], 1)) & (N - 1);
        const uint64_t j2 = (uint32_t)_mm_cvtsi128_si32(_mm512_extracti32x4_epi32(X[4], 2)) & (N - 1);
        const uint64_t j3 = (uint32_t)_mm_cvtsi128_si32(_mm512_extracti32x4_epi32(X[4], 3)) & (N - 1);
        for (int k = 0; k < 8; k++)
            X[k] = _mm512_xor_si512(X[k], gather4(V0 + 8 * j0 + k, V1 + 8 * j1 + k, V2 + 8 * j2 + k, V3 + 8 * j3 + k));
        blockmix_r1_avx512(X);
    }
    for (int k = 0; k < 8; k++) {
        _mm_storeu_si128((__m128i *)(x[0] + 4 * k), _mm512_castsi512_si128(X[k]));
        _mm_storeu_si128((__m128i *)(x[1] + 4 * k), _mm512_extracti32x4_epi32(X[k], 1));
        _mm_storeu_si128((__m128i *)(x[2] + 4 * k), _mm512_extracti32x4_epi32(X[k], 2));
        _mm_storeu_si128((__m128i *)(x[3] + 4 * k), _mm512_extracti32x4_epi32(X[k], 3));
    }
}
static int have_avx512(void) { return __builtin_cpu_supports("avx512f"); }
#else
#define ORACLE_HAVE_AVX2 0
static int have_avx512(void) { return 0; }
#endif
#else
#define ORACLE_HAVE_SSE2 0
#define ORACLE_HAVE_AVX2 0
#endif

static int g_oracle_impl = ORACLE_HAVE_AVX2 ? 2 : ORACLE_HAVE_SSE2;   /* 0 = scalar restatement, 1 = SSE2, 2 = AVX2 two labels at a time, 3 = AVX-512 four at a time (opt-in) */
int oracle_set_impl(int impl) {
    if (impl == 1 && !ORACLE_HAVE_SSE2) return -1;
    if (impl == 2 && !ORACLE_HAVE_AVX2) return -1;
    if (impl == 3 && !(ORACLE_HAVE_AVX2 && have_avx512())) return -1;
    if (impl < 0 || impl > 3) return -1;
    g_oracle_impl = impl;
    return 0;
}
int oracle_get_impl(void) { return g_oracle_impl; }

/* ======================================================================== label path */
void oracle_commitment(const uint8_t node_id[32], const uint8_t commitment_atx[32], uint8_t out[32]) {
    uint8_t buf[64];
    memcpy(buf, node_id, 32);
    memcpy(buf + 32, commitment_atx, 32);
    oracle_blake3_256(buf, 64, out);
}

/* The 72-byte scrypt password of label `index`: commitment || LE64(index) || 32 zero bytes (the slot of the
 * unused salt of the original gpu-post API); scrypt's own salt is empty.  Pinned by tests/golden/checkpoint_vrf.json. */
static void label_password(const uint8_t commitment[32], uint64_t index, uint8_t pw[72]) {
    memcpy(pw, commitment, 32);
    for (int i = 0; i < 8; i++) pw[32 + i] = (uint8_t)(index >> (8 * i));
    memset(pw + 40, 0, 32);
}

/* r = p = 1 fast path used for every label: one allocation-free ROMix over caller scratch. */
static void label32_r1(const uint8_t commitment[32], uint64_t index, uint64_t N, uint32_t *v, uint8_t out[32]) {
    uint8_t pw[72], B[128];
    uint32_t x[32], y[32];
    label_password(commitment, index, pw);
    oracle_pbkdf2_keccak512(pw, 72, NULL, 0, B, 128);
    for (int k = 0; k < 32; k++) x[k] = le32(B + 4 * k);
#if ORACLE_HAVE_SSE2
    if (g_oracle_impl >= 1) romix_r1_sse(x, v, N); else
#endif
    romix(x, v, y, N, 1, oracle_chacha20_8);
    for (int k = 0; k < 32; k++) put_le32(B + 4 * k, x[k]);
    oracle_pbkdf2_keccak512(pw, 72, B, 128, out, 32);
}

#if ORACLE_HAVE_AVX2
/* two labels at once (AVX2 lanes); va / vb = two scratchpads of 128*N bytes */
static void label32_r1_x2(const uint8_t ca[32], uint64_t ia, const uint8_t cb[32], uint64_t ib, uint64_t N,
                          uint32_t *va, uint32_t *vb, uint8_t outa[32], uint8_t outb[32]) {
    uint8_t pwa[72], pwb[72], Ba[128], Bb[128];
    uint32_t xa[32], xb[32];
    label_password(ca, ia, pwa); label_password(cb, ib, pwb);
    oracle_pbkdf2_keccak512(pwa, 72, NULL, 0, Ba, 128);
    oracle_pbkdf2_keccak512(pwb, 72, NULL, 0, Bb, 128);
    for (int k = 0; k < 32; k++) { xa[k] = le32(Ba + 4 * k); xb[k] = le32(Bb + 4 * k); }
    romix_r1_avx2_x2(xa, xb, va, vb, N);
    for (int k = 0; k < 32; k++) { put_le32(Ba + 4 * k, xa[k]); put_le32(Bb + 4 * k, xb[k]); }
    oracle_pbkdf2_keccak512(pwa, 72, Ba, 128, outa, 32);
    oracle_pbkdf2_keccak512(pwb, 72, Bb, 128, outb, 32);
}
#endif

#if ORACLE_HAVE_AVX2
T512 static void label32_r1_x4(const uint8_t *c[4], const uint64_t idx[4], uint64_t N, uint32_t *v, uint8_t out[4][32]) {
    uint8_t pw[4][72], B[4][128];
    uint32_t x[4][32];
    uint32_t *xp[4]; void *vp[4];
    for (int q = 0; q < 4; q++) {
        label_password(c[q], idx[q], pw[q]);
        oracle_pbkdf2_keccak512(pw[q], 72, NULL, 0, B[q], 128);
        for (int k = 0; k < 32; k++) x[q][k] = le32(B[q] + 4 * k);
        xp[q] = x[q]; vp[q] = v + (size_t)q * 32 * N;
    }
    romix_r1_avx512_x4(xp, vp, N);
    for (int q = 0; q < 4; q++) {
        for (int k = 0; k < 32; k++) put_le32(B[q] + 4 * k, x[q][k]);
        oracle_pbkdf2_keccak512(pw[q], 72, B[q], 128, out[q], 32);
    }
}
#endif

int oracle_label32(const uint8_t commitment[32], uint64_t index, uint64_t N, uint32_t r, uint32_t p,
                   uint8_t out[32]) {
    uint8_t pw[72];
    label_password(commitment, index, pw);
    return oracle_scrypt_jane(pw, 72, NULL, 0, N, r, p, out, 32);
}

void oracle_vrf_difficulty(uint64_t num_labels, uint8_t out[32]) {
    /* long division of 2^256 (33 bytes: 0x01 then 32 zero bytes) by num_labels, big-endian.
     * num_labels <= 1 would need 2^256 itself, which does not fit: saturate to 0xff..ff. */
    if (num_labels <= 1) { memset(out, 0xff, 32); return; }
    unsigned __int128 rem = 1; /* the leading 0x01 byte; its quotient digit is 0 for n > 1 */
    for (int i = 0; i < 32; i++) {
        rem <<= 8;
        out[i] = (uint8_t)(rem / num_labels);
        rem %= num_labels;
    }
}

typedef struct {
    const uint8_t *commitment;      /* range mode: one commitment; gather mode: n x 32 */
    const uint64_t *indices;        /* gather mode only */
    uint64_t N, start, count;
    uint8_t *out16;
    const uint8_t *vrf_difficulty;
    int tid, nthreads, gather;
    int found; uint64_t best_index; uint8_t best[32];
    int rc;
} worker_arg;

static void *worker(void *p) {
    worker_arg *a = (worker_arg *)p;
    uint32_t *v = (uint32_t *)malloc(4 * 128 * (size_t)a->N);   /* up to four scratchpads: the SIMD paths do 2 / 4 labels at a time */
    if (!v) { a->rc = -1; return NULL; }
    /* contiguous sub-range per thread so that "first index wins ties" is easy to merge */
    uint64_t per = (a->count + a->nthreads - 1) / a->nthreads;
    uint64_t lo = per * a->tid, hi = lo + per;
    if (hi > a->count) hi = a->count;
    a->found = 0;
    if (a->vrf_difficulty) memcpy(a->best, a->vrf_difficulty, 32);
    for (uint64_t k = lo; k < hi; k++) {
        uint8_t l32[4][32];
        int n = 1;
#if ORACLE_HAVE_AVX2
        if (g_oracle_impl == 3 && k + 3 < hi) {
            const uint8_t *c4[4]; uint64_t i4[4];
            for (int q = 0; q < 4; q++) {
                c4[q] = a->gather ? a->commitment + 32 * (k + q) : a->commitment;
                i4[q] = a->gather ? a->indices[k + q] : a->start + k + q;
            }
            n = 4;
            label32_r1_x4(c4, i4, a->N, v, l32);
        } else if (g_oracle_impl >= 2 && k + 1 < hi) {
            n = 2;
            if (a->gather) label32_r1_x2(a->commitment + 32 * k, a->indices[k], a->commitment + 32 * (k + 1), a->indices[k + 1], a->N,
                                         v, v + 32 * a->N, l32[0], l32[1]);
            else label32_r1_x2(a->commitment, a->start + k, a->commitment, a->start + k + 1, a->N, v, v + 32 * a->N, l32[0], l32[1]);
        } else
#endif
        if (a->gather) label32_r1(a->commitment + 32 * k, a->indices[k], a->N, v, l32[0]);
        else label32_r1(a->commitment, a->start + k, a->N, v, l32[0]);
        for (int q = 0; q < n; q++) {
            if (a->out16) memcpy(a->out16 + 16 * (k + q), l32[q], 16);
            if (a->vrf_difficulty && memcmp(l32[q], a->best, 32) < 0) {
                memcpy(a->best, l32[q], 32); a->best_index = a->start + k + q; a->found = 1;
            }
        }
        k += n - 1;
    }
    free(v);
    a->rc = 0;
    return NULL;
}

static int run_workers(worker_arg *proto, int threads, int *found, uint64_t *best_index, uint8_t best_label32[32]) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > proto->count && proto->count > 0) threads = (int)proto->count;
    worker_arg *args = (worker_arg *)calloc((size_t)threads, sizeof *args);
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof *th);
    if (!args || !th) { free(args); free(th); return -1; }
    for (int t = 0; t < threads; t++) {
        args[t] = *proto; args[t].tid = t; args[t].nthreads = threads;
        if (threads == 1) worker(&args[t]);
        else if (pthread_create(&th[t], NULL, worker, &args[t])) { args[t].rc = -2; worker(&args[t]); th[t] = 0; }
    }
    int rc = 0;
    for (int t = 0; t < threads; t++) {
        if (threads > 1 && th[t]) pthread_join(th[t], NULL);
        if (args[t].rc) rc = -1;
    }
    if (proto->vrf_difficulty && found) {
        uint8_t best[32];
        memcpy(best, proto->vrf_difficulty, 32);
        *found = 0;
        for (int t = 0; t < threads; t++) /* ascending index order + strict '<' keeps the first index on ties */
            if (args[t].found && memcmp(args[t].best, best, 32) < 0) {
                memcpy(best, args[t].best, 32); *found = 1;
                if (best_index) *best_index = args[t].best_index;
            }
        if (*found && best_label32) memcpy(best_label32, best, 32);
    }
    free(args); free(th);
    return rc;
}

int oracle_labels_range(const uint8_t commitment[32], uint64_t N, uint32_t r, uint32_t p,
                        uint64_t start, uint64_t count, uint8_t *out16,
                        const uint8_t *vrf_difficulty, int *found, uint64_t *best_index,
                        uint8_t best_label32[32], int threads) {
    if (N < 2 || (N & (N - 1)) || r != 1 || p != 1) return -1;
    if (found) *found = 0;
    if (count == 0) return 0;
    worker_arg a;
    memset(&a, 0, sizeof a);
    a.commitment = commitment; a.N = N; a.start = start; a.count = count; a.out16 = out16;
    a.vrf_difficulty = vrf_difficulty; a.gather = 0;
    return run_workers(&a, threads, found, best_index, best_label32);
}

int oracle_labels_gather(size_t n, const uint8_t *commitments, const uint64_t *indices,
                         uint64_t N, uint32_t r, uint32_t p, uint8_t *out16, int threads) {
    if (N < 2 || (N & (N - 1)) || r != 1 || p != 1) return -1;
    if (n == 0) return 0;
    worker_arg a;
    memset(&a, 0, sizeof a);
    a.commitment = commitments; a.indices = indices; a.N = N; a.count = n; a.out16 = out16; a.gather = 1;
    return run_workers(&a, threads, NULL, NULL, NULL);
}

double oracle_time_labels(const uint8_t commitment[32], uint64_t N, uint64_t start, uint64_t count,
                          int threads, uint8_t *out16) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    oracle_labels_range(commitment, N, 1, 1, start, count, out16, NULL, NULL, NULL, NULL, threads);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

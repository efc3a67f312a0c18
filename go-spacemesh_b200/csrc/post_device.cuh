// post_device.cuh — per-thread arithmetic of the POST label function for sm_100a.
//
// label32(i) = scrypt(P = commitment[32], S = LE64(i), N, r = 1, p = 1, dkLen = 32)
// (the function go-spacemesh reaches through activation/post.go:295 `mgr.init.Initialize` and
// activation/post_verifier.go:159 `ProofVerifier.Verify`; arithmetic per RFC 7914 / FIPS 180-4).
//
// Everything here is HD (host+device) so that tools/host_emul.cu can run the exact same
// per-thread code on the CPU against the oracle before any GPU time is spent.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PD_HD __host__ __device__ __forceinline__
#define PD_D __device__ __forceinline__
#else
#define PD_HD inline
#define PD_D inline
#endif

namespace b200post {

// ------------------------------------------------------------------------------------------------
// rotates.  On the device `rotl` is one SHF.L.W (alu pipe).  `rotl_mulwide` produces the same value
// with one IMAD.WIDE.U32 (fma pipe) whose two result halves are XOR-ed into the destination by the
// caller's LOP3 — used to move part of the Salsa rotate work off the alu pipe.
// ------------------------------------------------------------------------------------------------
PD_HD uint32_t rotl(uint32_t x, int k) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(x, x, k);
#else
    return (x << k) | (x >> (32 - k));
#endif
}
PD_HD uint32_t bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
#endif
}

// dst ^= rotl(s, k), with the rotate done as a 32x32->64 multiply by 2^k when MULWIDE.
// `pow2k` must hold 1u<<k in a register the compiler cannot constant-fold (a kernel parameter).
template <bool MULWIDE>
PD_HD void xor_rotl(uint32_t &dst, uint32_t s, int k, uint32_t pow2k) {
#if defined(__CUDA_ARCH__)
    if (MULWIDE) {
        uint32_t lo, hi;
        asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0,%1}, t;\n\t}"
            : "=r"(lo), "=r"(hi) : "r"(s), "r"(pow2k));
        dst = dst ^ lo ^ hi;   // one LOP3
        return;
    }
#endif
    (void)pow2k;
    dst ^= rotl(s, k);
}

// multipliers for the mul.wide rotate; filled by the host, passed by value as a kernel parameter so
// that they live in the constant bank (IMAD.WIDE takes a c[][] operand directly).
struct RotConsts { uint32_t p7, p9, p13, p18; };

// ------------------------------------------------------------------------------------------------
// Salsa20/8 (RFC 7914 §3): x <- x + rounds(x).
// MW is a 16-bit mask selecting which of the 16 rotates of a half-round use the mul.wide form
// (bit 4*q + r: quarter-round q = 0..3 of the half-round, rotate r = 0..3 for 7/9/13/18).  The same
// mask serves column and row half-rounds.  Measured on B200 (profiles/): SHF and LOP3 issue on the
// 16-lane alu pipe, IMAD.IADD / IMAD.WIDE on the 16-lane fmaheavy pipe (IMAD.WIDE at half rate), so
// moving ~1/3 of the rotates to IMAD.WIDE balances the two pipes; all-SHF is alu-bound.
// ------------------------------------------------------------------------------------------------
#define PD_QR(w, q, a, b, c, d)                                                    \
    xor_rotl<((MW >> (4 * (q) + 0)) & 1) != 0>(w[b], w[a] + w[d], 7, rc.p7);       \
    xor_rotl<((MW >> (4 * (q) + 1)) & 1) != 0>(w[c], w[b] + w[a], 9, rc.p9);       \
    xor_rotl<((MW >> (4 * (q) + 2)) & 1) != 0>(w[d], w[c] + w[b], 13, rc.p13);     \
    xor_rotl<((MW >> (4 * (q) + 3)) & 1) != 0>(w[a], w[d] + w[c], 18, rc.p18);
#define PD_DOUBLE_ROUND(w)                                                                               \
    PD_QR(w, 0, 0, 4, 8, 12) PD_QR(w, 1, 5, 9, 13, 1) PD_QR(w, 2, 10, 14, 2, 6) PD_QR(w, 3, 15, 3, 7, 11) \
    PD_QR(w, 0, 0, 1, 2, 3) PD_QR(w, 1, 5, 6, 7, 4) PD_QR(w, 2, 10, 11, 8, 9) PD_QR(w, 3, 15, 12, 13, 14)

template <int MW, int DR_UNROLL = 4>
PD_HD void salsa20_8(uint32_t (&x)[16], const RotConsts &rc) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = x[i];
#pragma unroll DR_UNROLL
    for (int r = 0; r < 4; r++) { PD_DOUBLE_ROUND(w) }
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] += w[i];
}

// Two independent Salsa20/8 cores in one instruction stream (8 independent dependency chains): used by
// the pipelined ROMix kernel, where every thread advances a filling and a mixing label together.
template <int MW, int DR_UNROLL>
PD_HD void salsa20_8_x2(uint32_t (&xa)[16], uint32_t (&xb)[16], const RotConsts &rc) {
    uint32_t wa[16], wb[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { wa[i] = xa[i]; wb[i] = xb[i]; }
#pragma unroll DR_UNROLL
    for (int r = 0; r < 4; r++) { PD_DOUBLE_ROUND(wa) PD_DOUBLE_ROUND(wb) }
#pragma unroll
    for (int i = 0; i < 16; i++) { xa[i] += wa[i]; xb[i] += wb[i]; }
}

// scryptBlockMix for r = 1 (RFC 7914 §4) on X = lo(16) || hi(16):
//   T = hi ^ lo; Y0 = Salsa(T); Y1 = Salsa(Y0 ^ hi); X = Y0 || Y1.
template <int MW, int DR_UNROLL = 4>
PD_HD void blockmix_r1(uint32_t (&lo)[16], uint32_t (&hi)[16], const RotConsts &rc) {
#pragma unroll
    for (int i = 0; i < 16; i++) lo[i] ^= hi[i];
    salsa20_8<MW, DR_UNROLL>(lo, rc);
#pragma unroll
    for (int i = 0; i < 16; i++) hi[i] ^= lo[i];
    salsa20_8<MW, DR_UNROLL>(hi, rc);
}
// Same, fused with the ROMix phase-2 "X ^= V[j]" so that the three-way XOR is one LOP3 per word.
template <int MW, int DR_UNROLL = 4>
PD_HD void blockmix_r1_xor(uint32_t (&lo)[16], uint32_t (&hi)[16], const uint32_t (&vlo)[16],
                           const uint32_t (&vhi)[16], const RotConsts &rc) {
#pragma unroll
    for (int i = 0; i < 16; i++) { hi[i] ^= vhi[i]; lo[i] = lo[i] ^ vlo[i] ^ hi[i]; }
    salsa20_8<MW, DR_UNROLL>(lo, rc);
#pragma unroll
    for (int i = 0; i < 16; i++) hi[i] ^= lo[i];
    salsa20_8<MW, DR_UNROLL>(hi, rc);
}

// One ROMix step of the pipelined kernel: BlockMix of the filling label (lo_f, hi_f) and
// BlockMix(X ^ V[j]) of the mixing label (lo_m, hi_m), interleaved.
template <int MW, int DR_UNROLL>
PD_HD void blockmix_r1_dual(uint32_t (&lo_f)[16], uint32_t (&hi_f)[16], uint32_t (&lo_m)[16], uint32_t (&hi_m)[16],
                            const uint32_t (&vlo)[16], const uint32_t (&vhi)[16], const RotConsts &rc) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
        lo_f[i] ^= hi_f[i];
        hi_m[i] ^= vhi[i]; lo_m[i] = lo_m[i] ^ vlo[i] ^ hi_m[i];
    }
    salsa20_8_x2<MW, DR_UNROLL>(lo_f, lo_m, rc);
#pragma unroll
    for (int i = 0; i < 16; i++) { hi_f[i] ^= lo_f[i]; hi_m[i] ^= lo_m[i]; }
    salsa20_8_x2<MW, DR_UNROLL>(hi_f, hi_m, rc);
}

// ------------------------------------------------------------------------------------------------
// SHA-256 (FIPS 180-4 §6.2).  One compression of a 16-word big-endian block `w` into `st`.
// The round loop is 4 x 16 with a register ring so that code size stays small; SHA is < 0.3 % of
// the work at N = 8192 (12 compressions vs 32768 Salsa20/8 cores per label).
// ------------------------------------------------------------------------------------------------
#define PD_K256_TABLE \
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, \
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, \
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, \
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, \
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, \
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, \
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, \
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2
static const uint32_t h_K256[64] = {PD_K256_TABLE};
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t c_K256[64] = {PD_K256_TABLE};
#endif

PD_HD uint32_t sha_k(int i) {
#if defined(__CUDA_ARCH__)
    return c_K256[i];
#else
    return h_K256[i];
#endif
}
PD_HD uint32_t rotr(uint32_t x, int k) { return rotl(x, 32 - k); }

PD_HD void sha256_compress(uint32_t (&st)[8], uint32_t (&w)[16]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
    for (int base = 0; base < 64; base += 16) {
#pragma unroll
        for (int t = 0; t < 16; t++) {
            if (base) {
                const uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
                const uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
                const uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
                w[t] = w[t] + s0 + w[(t + 9) & 15] + s1;
            }
            const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t t1 = h + S1 + ch + sha_k(base + t) + w[t];
            const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + S0 + mj;
        }
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

PD_HD void sha256_iv(uint32_t (&st)[8]) {
    st[0] = 0x6a09e667; st[1] = 0xbb67ae85; st[2] = 0x3c6ef372; st[3] = 0xa54ff53a;
    st[4] = 0x510e527f; st[5] = 0x9b05688c; st[6] = 0x1f83d9ab; st[7] = 0x5be0cd19;
}

// HMAC-SHA256 key schedule for a 32-byte key (the commitment): the SHA-256 states after absorbing
// K^ipad and K^opad (FIPS 198-1 §4).  `key_be` = the 8 big-endian words of the commitment.
struct HmacMid { uint32_t inner[8], outer[8]; };

PD_HD void hmac_midstates(const uint32_t (&key_be)[8], HmacMid &m) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = key_be[i] ^ 0x36363636u; w[i + 8] = 0x36363636u; }
    sha256_iv(m.inner);
    sha256_compress(m.inner, w);
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = key_be[i] ^ 0x5c5c5c5cu; w[i + 8] = 0x5c5c5c5cu; }
    sha256_iv(m.outer);
    sha256_compress(m.outer, w);
}

// PBKDF2-HMAC-SHA256(P = commitment, S = LE64(index), c = 1, dkLen = 128) -> X as 32 LE words
// (RFC 7914 §6 step 1).  8 compressions.
PD_HD void pbkdf2_expand(const HmacMid &m, uint64_t index, uint32_t (&lo)[16], uint32_t (&hi)[16]) {
    // salt bytes = LE64(index); SHA reads big-endian words => word0 = bswap(low 32), word1 = bswap(high 32)
    const uint32_t s0 = bswap32((uint32_t)index), s1 = bswap32((uint32_t)(index >> 32));
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        uint32_t st[8], w[16];
#pragma unroll
        for (int i = 0; i < 8; i++) st[i] = m.inner[i];
        w[0] = s0; w[1] = s1; w[2] = (uint32_t)(k + 1); w[3] = 0x80000000u;
#pragma unroll
        for (int i = 4; i < 15; i++) w[i] = 0;
        w[15] = (64 + 12) * 8;
        sha256_compress(st, w);
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = st[i];
        w[8] = 0x80000000u;
#pragma unroll
        for (int i = 9; i < 15; i++) w[i] = 0;
        w[15] = (64 + 32) * 8;
#pragma unroll
        for (int i = 0; i < 8; i++) st[i] = m.outer[i];
        sha256_compress(st, w);
        // output block k is bytes [32k, 32k+32) of B; X words are little-endian reads of B
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t v = bswap32(st[i]);
            // static indexing only: k is a runtime loop variable, so select with predicates
            if (k == 0) lo[i] = v; else if (k == 1) lo[8 + i] = v; else if (k == 2) hi[i] = v; else hi[8 + i] = v;
        }
    }
}

// PBKDF2-HMAC-SHA256(P = commitment, S = X (128 bytes), c = 1, dkLen = 32) -> 8 big-endian words
// (RFC 7914 §6 step 3).  4 compressions.
PD_HD void pbkdf2_final(const HmacMid &m, const uint32_t (&lo)[16], const uint32_t (&hi)[16], uint32_t (&out_be)[8]) {
    uint32_t st[8], w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = m.inner[i];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = bswap32(lo[i]);
    sha256_compress(st, w);
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = bswap32(hi[i]);
    sha256_compress(st, w);
    w[0] = 1; w[1] = 0x80000000u;
#pragma unroll
    for (int i = 2; i < 15; i++) w[i] = 0;
    w[15] = (64 + 128 + 4) * 8;
    sha256_compress(st, w);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = st[i];
    w[8] = 0x80000000u;
#pragma unroll
    for (int i = 9; i < 15; i++) w[i] = 0;
    w[15] = (64 + 32) * 8;
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = m.outer[i];
    sha256_compress(st, w);
#pragma unroll
    for (int i = 0; i < 8; i++) out_be[i] = st[i];
}

}  // namespace b200post

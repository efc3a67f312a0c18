// aes_device.cuh — AES-128 single-block encryption for sm_100a kernels (FIPS-197, T-table form).
//
// Used by the verify epilogue (and the proving scan): one 16-byte label is encrypted under a per-proof key
// and one ciphertext byte is compared with the proving difficulty (ASSUMED post-rs Prover8_56 scheme, see
// include/b200post_verify.h).  Round keys are expanded on the host (AES-NI) and read as 11 x uint4.
//
// State words are little-endian columns (byte 0 = row 0).  One 1-KiB table T0 lives in shared memory
// (T0[x] = {2·S[x], S[x], S[x], 3·S[x]}); T1..T3 are byte rotations of it (one PRMT each).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace b200post {

struct AesTables { uint32_t t0[256]; uint8_t sbox[256]; };   // host-built once, uploaded to global memory

// fills `t` (host): S-box from the GF(2^8) inverse + affine map, then the MixColumns-folded table
inline void aes_build_tables(AesTables &t) {
    auto xtime = [](uint8_t a) { return (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1b : 0)); };
    // log/antilog over generator 3
    uint8_t exp[256], log[256];
    uint8_t x = 1;
    for (int i = 0; i < 255; i++) { exp[i] = x; log[x] = (uint8_t)i; x = (uint8_t)(x ^ xtime(x)); }
    exp[255] = exp[0];
    for (int v = 0; v < 256; v++) {
        const uint8_t inv = v ? exp[(255 - log[v]) % 255] : 0;
        uint8_t s = inv, r = inv;
        for (int k = 0; k < 4; k++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
        s ^= 0x63;
        t.sbox[v] = s;
        const uint8_t s2 = xtime(s), s3 = (uint8_t)(s2 ^ s);
        t.t0[v] = (uint32_t)s2 | ((uint32_t)s << 8) | ((uint32_t)s << 16) | ((uint32_t)s3 << 24);
    }
}

#if defined(__CUDACC__)
// shared-memory image used by the device functions below
struct AesSmem { uint32_t t0[256]; uint32_t sbox[256]; };   // sbox widened to words: conflict-free byte picks

__device__ __forceinline__ void aes_load_smem(AesSmem &s, const AesTables *__restrict__ g) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { s.t0[i] = g->t0[i]; s.sbox[i] = g->sbox[i]; }
    __syncthreads();
}

__device__ __forceinline__ uint32_t rotl8(uint32_t v) { return __byte_perm(v, 0, 0x2103); }    // bytes (b3,b0,b1,b2) -> T1
__device__ __forceinline__ uint32_t rotl16(uint32_t v) { return __byte_perm(v, 0, 0x1032); }   // T2
__device__ __forceinline__ uint32_t rotl24(uint32_t v) { return __byte_perm(v, 0, 0x0321); }   // T3

// out = AES-128-Encrypt(rk, in); rk = 11 round keys as little-endian column words
__device__ __forceinline__ uint4 aes128_encrypt(const AesSmem &s, const uint4 *__restrict__ rk, uint4 in) {
    uint4 k = rk[0];
    uint32_t s0 = in.x ^ k.x, s1 = in.y ^ k.y, s2 = in.z ^ k.z, s3 = in.w ^ k.w;
#pragma unroll 1
    for (int r = 1; r < 10; r++) {
        k = rk[r];
        const uint32_t t0 = s.t0[s0 & 0xff] ^ rotl8(s.t0[(s1 >> 8) & 0xff]) ^ rotl16(s.t0[(s2 >> 16) & 0xff]) ^ rotl24(s.t0[s3 >> 24]) ^ k.x;
        const uint32_t t1 = s.t0[s1 & 0xff] ^ rotl8(s.t0[(s2 >> 8) & 0xff]) ^ rotl16(s.t0[(s3 >> 16) & 0xff]) ^ rotl24(s.t0[s0 >> 24]) ^ k.y;
        const uint32_t t2 = s.t0[s2 & 0xff] ^ rotl8(s.t0[(s3 >> 8) & 0xff]) ^ rotl16(s.t0[(s0 >> 16) & 0xff]) ^ rotl24(s.t0[s1 >> 24]) ^ k.z;
        const uint32_t t3 = s.t0[s3 & 0xff] ^ rotl8(s.t0[(s0 >> 8) & 0xff]) ^ rotl16(s.t0[(s1 >> 16) & 0xff]) ^ rotl24(s.t0[s2 >> 24]) ^ k.w;
        s0 = t0; s1 = t1; s2 = t2; s3 = t3;
    }
    k = rk[10];
    uint4 o;
    o.x = (s.sbox[s0 & 0xff] | (s.sbox[(s1 >> 8) & 0xff] << 8) | (s.sbox[(s2 >> 16) & 0xff] << 16) | (s.sbox[s3 >> 24] << 24)) ^ k.x;
    o.y = (s.sbox[s1 & 0xff] | (s.sbox[(s2 >> 8) & 0xff] << 8) | (s.sbox[(s3 >> 16) & 0xff] << 16) | (s.sbox[s0 >> 24] << 24)) ^ k.y;
    o.z = (s.sbox[s2 & 0xff] | (s.sbox[(s3 >> 8) & 0xff] << 8) | (s.sbox[(s0 >> 16) & 0xff] << 16) | (s.sbox[s1 >> 24] << 24)) ^ k.z;
    o.w = (s.sbox[s3 & 0xff] | (s.sbox[(s0 >> 8) & 0xff] << 8) | (s.sbox[(s1 >> 16) & 0xff] << 16) | (s.sbox[s2 >> 24] << 24)) ^ k.w;
    return o;
}
__device__ __forceinline__ uint32_t uint4_byte(const uint4 &v, uint32_t i) {
    const uint32_t w = i < 4 ? v.x : i < 8 ? v.y : i < 12 ? v.z : v.w;
    return (w >> (8 * (i & 3))) & 0xff;
}
#endif

}  // namespace b200post

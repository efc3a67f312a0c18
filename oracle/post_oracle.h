/*
 * post_oracle.h — CPU ORACLE for the POST label path.  TEST INFRASTRUCTURE ONLY.
 *
 * A from-spec restatement of the arithmetic go-spacemesh reaches through github.com/spacemeshos/post v0.12.9
 * (go.mod:48) -> libpost (post-rs, Makefile-libs.Inc:49-51).  Neither dependency is present under
 * /root/reference.  The label function is scrypt-jane's scrypt with ChaCha20/8 as the mix core and
 * HMAC-Keccak-512 (original 0x01 padding) inside PBKDF2:
 *     label32(i) = scrypt_jane(P = commitment || LE64(i) || 0^32, S = "", N, r = 1, p = 1, dkLen = 32)
 *     commitment = blake3(nodeID || commitmentATX);  label = label32[0:16];  VRF compare: 32 bytes big-endian.
 *
 *   PARITY STATUS
 *   - label function, commitment and the VRF comparison: PINNED against real data.  The reference's
 *     checkpoint/checkpointdata.json holds 42 identities of a LabelsPerUnit = 1024, N = 8192 network with their
 *     VRF nonces; under this function every nonce's label32 is within a factor of 5 of 2^256/numLabels (the
 *     arg-min of numLabels uniform draws; 26 of 42 below the threshold = 1 - 1/e), under RFC 7914 scrypt or any
 *     other convention tried none is (tools/pin_search.py, tests/golden/checkpoint_vrf.json).  Recomputing all
 *     1 899 520 labels of the 42 POSTs with this file, every recorded nonce is exactly the arg-min of its POST.
 *   - primitives: Keccak-f pinned against hashlib's SHA3-512 (same permutation, other pad byte), ChaCha20/8,
 *     PBKDF2 and ROMix against an independent numpy restatement (oracle/pyoracle.py), RFC 7914 / FIPS vectors
 *     for the SHA-256 / Salsa building blocks, BLAKE3 and FIPS-197 known answers.
 *   - the proof-side conventions (AES keys, index packing, K3 subset) stay "parity unpinned": the reference
 *     tree holds no proof vector.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (libb200post.so) never links or calls it.
 */
#ifndef POST_ORACLE_H
#define POST_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- primitives ------------------------------------------------------------------- */
void oracle_sha256(const uint8_t *msg, size_t len, uint8_t out[32]);
void oracle_hmac_sha256(const uint8_t *key, size_t klen, const uint8_t *msg, size_t mlen, uint8_t out[32]);
void oracle_pbkdf2_sha256(const uint8_t *pw, size_t pwlen, const uint8_t *salt, size_t saltlen,
                          uint32_t iters, uint8_t *out, size_t dklen);
/* Salsa20/8 core on 16 little-endian words, in place (RFC 7914 §3). */
void oracle_salsa20_8(uint32_t b[16]);
/* scryptBlockMix on 2r 64-byte blocks (RFC 7914 §4). `y` is scratch of the same size. */
void oracle_blockmix(uint32_t *b, uint32_t *y, uint32_t r);
/* Full scrypt (RFC 7914 §6). Returns 0, or -1 on bad parameters / allocation failure. */
int oracle_scrypt(const uint8_t *pw, size_t pwlen, const uint8_t *salt, size_t saltlen,
                  uint64_t N, uint32_t r, uint32_t p, uint8_t *out, size_t dklen);
/* ChaCha20/8 core on 16 little-endian words, in place (scrypt-jane's mix function). */
void oracle_chacha20_8(uint32_t b[16]);
/* Keccak[r = 576, c = 1024] with 64 bytes of output; pad = 0x01 (Keccak-512 as scrypt-jane uses it) or 0x06 (SHA3-512). */
void oracle_keccak512(const uint8_t *msg, size_t len, uint8_t pad, uint8_t out[64]);
void oracle_hmac_keccak512(const uint8_t *key, size_t klen, const uint8_t *msg, size_t mlen, uint8_t out[64]);
void oracle_pbkdf2_keccak512(const uint8_t *pw, size_t pwlen, const uint8_t *salt, size_t saltlen, uint8_t *out, size_t dklen);
/* scrypt-jane (ChaCha20/8 + Keccak-512): the POST label function's core.  0, or -1 on bad parameters. */
int oracle_scrypt_jane(const uint8_t *pw, size_t pwlen, const uint8_t *salt, size_t saltlen,
                       uint64_t N, uint32_t r, uint32_t p, uint8_t *out, size_t dklen);
/* BLAKE3-256 of an arbitrary-length message (hash/hash.go:16-25 uses zeebo/blake3). */
void oracle_blake3_256(const uint8_t *msg, size_t len, uint8_t out[32]);
/* BLAKE3 with extended output (XOF), used for the verify epilogue's key derivation. */
void oracle_blake3_xof(const uint8_t *msg, size_t len, uint8_t *out, size_t outlen);
/* AES-128 single-block encryption (FIPS-197), key schedule done per call. */
void oracle_aes128_encrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16]);

/* ---- the label path (Appendix A) ---------------------- */
/* commitment = blake3(nodeID || commitmentATX)   (activation/post.go:355-361 passes both) */
void oracle_commitment(const uint8_t node_id[32], const uint8_t commitment_atx[32], uint8_t out[32]);
/* label32(i) = scrypt_jane(P = commitment || LE64(i) || 0^32, S = "", N, r, p, dkLen = 32).  0 on success. */
int oracle_label32(const uint8_t commitment[32], uint64_t index, uint64_t N, uint32_t r, uint32_t p,
                   uint8_t out[32]);
/* labels for the contiguous range [start, start+count): 16 bytes each into out16
 * (activation/post.go:295 -> Initialize).  If vrf_difficulty != NULL also runs the VRF-nonce
 * scan: *found = 1 and (best_index, best_label32) set when some label32 is lexicographically
 * < the running minimum (initially vrf_difficulty); first index wins ties.  `threads` <= 1
 * is single-threaded. Returns 0 on success. */
int oracle_labels_range(const uint8_t commitment[32], uint64_t N, uint32_t r, uint32_t p,
                        uint64_t start, uint64_t count, uint8_t *out16,
                        const uint8_t *vrf_difficulty, int *found, uint64_t *best_index,
                        uint8_t best_label32[32], int threads);
/* labels for scattered (commitment, index) pairs (activation/post_verifier.go:159 ->
 * verify_proof recomputes K2/K3 labels).  commitments is n x 32 bytes. out is n x 16. */
int oracle_labels_gather(size_t n, const uint8_t *commitments, const uint64_t *indices,
                         uint64_t N, uint32_t r, uint32_t p, uint8_t *out16, int threads);
/* VRF target = floor(2^256 / num_labels) as 32 big-endian bytes (activation/validation.go:261-282). */
void oracle_vrf_difficulty(uint64_t num_labels, uint8_t out[32]);

/* timing helper for bench.py's cpu_baseline: computes `count`
 * labels starting at `start` on `threads` threads, returns elapsed seconds. */
/* ROMix implementation used by the label functions: 0 = scalar restatement, 1 = SSE2, 2 = AVX2 with two labels
 * per thread in lock-step (the default where available), 3 = AVX-512 with four labels per thread (run-time CPU check;
 * opt-in: bench.py times both and reports the faster).  All cross-checked against 0 in tests/.  Returns 0, or -1 if
 * unsupported. */
int oracle_set_impl(int impl);
int oracle_get_impl(void);

double oracle_time_labels(const uint8_t commitment[32], uint64_t N, uint64_t start, uint64_t count,
                          int threads, uint8_t *out16);

#ifdef __cplusplus
}
#endif
#endif

/*
 * post_compat.h — the subset of libpost's C header (`post.h`, post-rs v0.7.13) that the POST *label*
 * path of github.com/spacemeshos/post v0.12.9 binds through cgo, re-declared so that libb200post.so can
 * stand in for `-lpost` on that path (link mechanics: Makefile-libs.Inc:5-6,59-69 in the reference;
 * call sites: activation/post.go:295,355-361, activation/post_supervisor.go:106,121).
 *
 * `post.h` is NOT present under /root/reference (it is fetched at build time, Makefile-libs.Inc:78-94),
 * so these declarations are a restatement of the published header and are marked "unpinned" in
 * a maintainer must diff them against the real post.h before linking.
 *
 * Semantics kept from libpost:
 *   - `initialize(init, start, end, out, &nonce)`: `end` is INCLUSIVE; `out` receives 16 bytes per label;
 *     returns InitializeOk when a VRF nonce below the (running) difficulty was found in this range and
 *     stores its index in *nonce, else InitializeOkNonceNotFound.
 *   - the callee never frees or retains caller memory; an Initializer is used by one thread at a time.
 * Difference: provider id 0xffffffff (CPU) is refused — this library has no CPU path by design.
 */
#ifndef B200POST_POST_COMPAT_H
#define B200POST_POST_COMPAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum DeviceClass { DeviceClassCPU = 1, DeviceClassGPU = 2 } DeviceClass;

typedef struct Provider {
    char name[64];
    uint32_t id;
    DeviceClass class_;
} Provider;

typedef enum DeviceInfoResult {
    DeviceInfoOk = 0,
    DeviceInfoInvalidArgument = 1,
    DeviceInfoBufferTooSmall = 2,
    DeviceInfoFailed = 3
} DeviceInfoResult;

typedef enum InitializeResult {
    InitializeOk = 0,
    InitializeOkNonceNotFound = 1,
    InitializeInvalidLabelsRange = 2,
    InitializeError = 3,
    InitializeInvalidArgument = 4
} InitializeResult;

typedef struct Initializer Initializer;

size_t get_providers_count(void);
DeviceInfoResult get_providers(Provider *out, size_t out_len);

/* n = scrypt N; commitment = 32 bytes; vrf_difficulty = 32 bytes big-endian or NULL. NULL on error. */
Initializer *new_initializer(uint32_t provider_id, size_t n, const uint8_t *commitment, const uint8_t *vrf_difficulty);
InitializeResult initialize(Initializer *init, uint64_t start, uint64_t end, uint8_t *out, uint64_t *nonce);
void free_initializer(Initializer *init);

#ifdef __cplusplus
}
#endif
#endif

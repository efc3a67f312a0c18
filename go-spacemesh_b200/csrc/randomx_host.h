// randomx_host.h — host half of the k2pow (RandomX) engine: everything that is sequential and runs once per cache
// key before the GPU takes over.  RandomX (tevador/RandomX v1.1.x, doc/specs.md) is what go-spacemesh's k2pow is at
// this revision (cmd/root.go:254-259 "randomx-based proof of work", activation/post_types.go:116-121 RandomXMode;
// asked for over the RPC at activation/nipost.go:171 and checked inside Verify, activation/post_verifier.go:150-160).
//
//   * Blake2b (RFC 7693)                               — seeds, the SuperscalarHash generator's byte stream
//   * Argon2d v0x13, 1 lane, 262144 KiB, 3 passes       — the 256 MiB cache (spec §7.1); a strictly sequential fill
//   * SuperscalarHash program generator (spec §6)      — 8 programs per key
// The dataset (2080 MiB), scratchpad fill, program generation, the VM and the final hash run on the device
// (randomx_kernels.cu).  No part of oracle/ is used here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace b200post {
namespace rx {

constexpr uint32_t kCacheKiB = 262144, kArgonPasses = 3, kCacheAccesses = 8;
constexpr uint64_t kDatasetBase = 2147483648ull, kDatasetExtra = 33554368ull;
constexpr uint64_t kDatasetItems = (kDatasetBase + kDatasetExtra) / 64;   // 34 078 719 x 64 B
constexpr uint32_t kScratchpadL3 = 2097152, kScratchpadL2 = 262144, kScratchpadL1 = 16384;
constexpr int kProgramSize = 256, kProgramIterations = 2048, kProgramCount = 8;

void blake2b(void *out, size_t outlen, const void *in, size_t inlen);

// One SuperscalarHash instruction as the device interprets it (16 bytes): the reciprocal of IMUL_RCP is resolved here.
struct SsOp {
    uint8_t opcode, dst, src, shift;
    uint32_t imm32;
    uint64_t rcp;
};
enum SsOpcode : uint8_t { SS_ISUB_R, SS_IXOR_R, SS_IADD_RS, SS_IMUL_R, SS_IROR_C, SS_IADD_C, SS_IXOR_C, SS_IMULH_R, SS_ISMULH_R, SS_IMUL_RCP };

struct SsProgram {
    std::vector<SsOp> ops;
    uint32_t address_reg = 0;
};

struct CacheImage {
    std::vector<uint64_t> memory;          // 256 MiB Argon2d memory
    SsProgram programs[kCacheAccesses];
};

uint64_t reciprocal(uint32_t divisor);     // floor(2^x / divisor) with the top bit set (spec §5.2.8)
// Fills `out` for `key` (any length; the generator uses its first 60 bytes).  ~0.7 s on one core.
void build_cache(const void *key, size_t keylen, CacheImage &out);

}  // namespace rx
}  // namespace b200post

// randomx_host.cpp — see randomx_host.h.  Written from the RandomX specification (doc/specs.md §3.1, §6, §7.1),
// RFC 7693 (Blake2b) and RFC 9106 (Argon2).
#include "randomx_host.h"

#include <algorithm>
#include <array>
#include <cstring>

namespace b200post {
namespace rx {
namespace {

inline uint64_t rotr(uint64_t v, unsigned n) { return (v >> n) | (v << ((64 - n) & 63)); }
inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline void wr32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }

// ---------------------------------------------------------------- Blake2b
constexpr std::array<uint64_t, 8> kB2Iv = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                           0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
constexpr uint8_t kB2Sigma[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

class Blake2b {
public:
    explicit Blake2b(size_t outlen) : outlen_(outlen) {
        h_ = kB2Iv;
        h_[0] ^= 0x01010000ull ^ outlen;
    }
    void update(const void *data, size_t len) {
        const uint8_t *p = static_cast<const uint8_t *>(data);
        while (len) {
            if (fill_ == 128) { count_ += 128; compress(false); fill_ = 0; }
            const size_t take = std::min(len, (size_t)128 - fill_);
            memcpy(block_ + fill_, p, take);
            fill_ += take; p += take; len -= take;
        }
    }
    void finish(void *out) {
        count_ += fill_;
        memset(block_ + fill_, 0, 128 - fill_);
        compress(true);
        uint8_t full[64];
        for (int i = 0; i < 8; i++) memcpy(full + 8 * i, &h_[i], 8);
        memcpy(out, full, outlen_);
    }

private:
    void compress(bool last) {
        uint64_t m[16], v[16];
        for (int i = 0; i < 16; i++) m[i] = rd64(block_ + 8 * i);
        for (int i = 0; i < 8; i++) { v[i] = h_[i]; v[i + 8] = kB2Iv[i]; }
        v[12] ^= count_;            // message lengths on this path are far below 2^64: the high counter word stays 0
        if (last) v[14] = ~v[14];
        auto g = [&](int r, int i, int a, int b, int c, int d) {
            v[a] += v[b] + m[kB2Sigma[r % 10][2 * i]];     v[d] = rotr(v[d] ^ v[a], 32);
            v[c] += v[d];                                   v[b] = rotr(v[b] ^ v[c], 24);
            v[a] += v[b] + m[kB2Sigma[r % 10][2 * i + 1]]; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] += v[d];                                   v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; r++) {
            g(r, 0, 0, 4, 8, 12); g(r, 1, 1, 5, 9, 13); g(r, 2, 2, 6, 10, 14); g(r, 3, 3, 7, 11, 15);
            g(r, 4, 0, 5, 10, 15); g(r, 5, 1, 6, 11, 12); g(r, 6, 2, 7, 8, 13); g(r, 7, 3, 4, 9, 14);
        }
        for (int i = 0; i < 8; i++) h_[i] ^= v[i] ^ v[i + 8];
    }
    std::array<uint64_t, 8> h_{};
    uint8_t block_[128] = {0};
    size_t fill_ = 0, outlen_;
    uint64_t count_ = 0;
};

// ---------------------------------------------------------------- Argon2d (RFC 9106 §3), one lane
// H' of RFC 9106 §3.3: arbitrary-length output from Blake2b
void long_hash(uint8_t *out, uint32_t outlen, const uint8_t *in, size_t inlen) {
    uint8_t len_le[4]; wr32(len_le, outlen);
    if (outlen <= 64) { Blake2b h(outlen); h.update(len_le, 4); h.update(in, inlen); h.finish(out); return; }
    uint8_t cur[64], nxt[64];
    { Blake2b h(64); h.update(len_le, 4); h.update(in, inlen); h.finish(cur); }
    memcpy(out, cur, 32);
    uint32_t pos = 32, left = outlen - 32;
    while (left > 64) {
        blake2b(nxt, 64, cur, 64); memcpy(cur, nxt, 64);
        memcpy(out + pos, cur, 32); pos += 32; left -= 32;
    }
    blake2b(nxt, left, cur, 64);
    memcpy(out + pos, nxt, left);
}

struct Block { uint64_t w[128]; };   // 1 KiB

inline void bla_mix(uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
    auto f = [](uint64_t x, uint64_t y) { return x + y + 2 * (x & 0xffffffffull) * (y & 0xffffffffull); };
    a = f(a, b); d = rotr(d ^ a, 32); c = f(c, d); b = rotr(b ^ c, 24);
    a = f(a, b); d = rotr(d ^ a, 16); c = f(c, d); b = rotr(b ^ c, 63);
}
// the permutation P on 16 words picked out of `w` by `idx`
inline void permute(uint64_t *w, const int (&idx)[16]) {
    auto G = [&](int a, int b, int c, int d) { bla_mix(w[idx[a]], w[idx[b]], w[idx[c]], w[idx[d]]); };
    G(0, 4, 8, 12); G(1, 5, 9, 13); G(2, 6, 10, 14); G(3, 7, 11, 15);
    G(0, 5, 10, 15); G(1, 6, 11, 12); G(2, 7, 8, 13); G(3, 4, 9, 14);
}
// compression G of RFC 9106 §3.5; xor_into = the v1.3 "XOR over" of passes > 0
void compress_block(const Block &x, const Block &y, Block &dst, bool xor_into) {
    Block r, z;
    for (int i = 0; i < 128; i++) { r.w[i] = x.w[i] ^ y.w[i]; z.w[i] = r.w[i]; }
    for (int row = 0; row < 8; row++) {
        int idx[16];
        for (int k = 0; k < 16; k++) idx[k] = 16 * row + k;
        permute(z.w, idx);
    }
    for (int col = 0; col < 8; col++) {
        int idx[16];
        for (int k = 0; k < 8; k++) { idx[2 * k] = 16 * k + 2 * col; idx[2 * k + 1] = 16 * k + 2 * col + 1; }
        permute(z.w, idx);
    }
    for (int i = 0; i < 128; i++) dst.w[i] = (xor_into ? dst.w[i] : 0) ^ r.w[i] ^ z.w[i];
}

void argon2d_fill(Block *mem, uint32_t blocks, uint32_t passes, const void *pwd, uint32_t pwdlen, const void *salt, uint32_t saltlen) {
    uint8_t h0[72], le[4];
    Blake2b h(64);
    const uint32_t header[6] = {1u /*lanes*/, 0u /*tag length: RandomX reads the memory, not a tag*/, blocks /*KiB*/, passes, 0x13u, 0u /*Argon2d*/};
    for (uint32_t v : header) { wr32(le, v); h.update(le, 4); }
    wr32(le, pwdlen); h.update(le, 4); h.update(pwd, pwdlen);
    wr32(le, saltlen); h.update(le, 4); h.update(salt, saltlen);
    wr32(le, 0); h.update(le, 4); h.update(le, 4);
    h.finish(h0);
    wr32(h0 + 68, 0);
    wr32(h0 + 64, 0); long_hash(reinterpret_cast<uint8_t *>(&mem[0]), 1024, h0, 72);
    wr32(h0 + 64, 1); long_hash(reinterpret_cast<uint8_t *>(&mem[1]), 1024, h0, 72);
    const uint32_t seg = blocks / 4;
    for (uint32_t pass = 0; pass < passes; pass++)
        for (uint32_t slice = 0; slice < 4; slice++)
            for (uint32_t i = (pass == 0 && slice == 0) ? 2 : 0; i < seg; i++) {
                const uint32_t cur = slice * seg + i, prev = cur ? cur - 1 : blocks - 1;
                // data-dependent reference block: J1 = low word of the previous block, mapped quadratically onto the
                // window of blocks already written (RFC 9106 §3.4.2)
                const uint64_t j1 = mem[prev].w[0] & 0xffffffffull;
                const uint64_t window = pass == 0 ? (uint64_t)cur - 1 : (uint64_t)blocks - seg + i - 1;
                const uint64_t x = (j1 * j1) >> 32;
                const uint64_t rel = window - 1 - ((window * x) >> 32);
                const uint64_t base = (pass == 0 || slice == 3) ? 0 : (uint64_t)(slice + 1) * seg;
                compress_block(mem[prev], mem[(base + rel) % blocks], mem[cur], pass != 0);
            }
}

// ---------------------------------------------------------------- SuperscalarHash generator (spec §6)
// Simulates a 3-port (P0, P1, P5) out-of-order core with a 16-byte/cycle decoder: the program is whatever keeps
// that core busy for 170 cycles.
enum Port : int { kP0 = 1, kP1 = 2, kP5 = 4, kP01 = 3, kP05 = 5, kP015 = 7 };
struct MacroOp { int bytes, latency, uop_a, uop_b; bool chained; };
constexpr MacroOp kSubRR{3, 1, kP015, 0, false}, kXorRR{3, 1, kP015, 0, false}, kMulWide{3, 4, kP1, kP5, false},
    kMovRR{3, 0, 0, 0, false}, kMovRRChained{3, 0, 0, 0, true}, kLea{4, 1, kP01, 0, false}, kImulRR{4, 3, kP1, 0, false},
    kImulRRChained{4, 3, kP1, 0, true}, kRorRI{4, 1, kP05, 0, false}, kAluRI{7, 1, kP015, 0, false}, kMovImm64{10, 1, kP015, 0, false};

// the generator's 14 instruction kinds (the 7/8/9-byte encodings of add/xor imm differ only in padding)
enum Kind : int { kISUB_R, kIXOR_R, kIADD_RS, kIMUL_R, kIROR_C, kIADD_C7, kIXOR_C7, kIADD_C8, kIXOR_C8, kIADD_C9, kIXOR_C9,
                  kIMULH_R, kISMULH_R, kIMUL_RCP, kKinds, kNone = -1 };
struct KindInfo { Kind kind; int n; MacroOp ops[3]; int result_at, dst_at, src_at; };
constexpr KindInfo kInfo[kKinds] = {
    {kISUB_R, 1, {kSubRR}, 0, 0, 0},    {kIXOR_R, 1, {kXorRR}, 0, 0, 0},    {kIADD_RS, 1, {kLea}, 0, 0, 0},
    {kIMUL_R, 1, {kImulRR}, 0, 0, 0},   {kIROR_C, 1, {kRorRI}, 0, 0, -1},   {kIADD_C7, 1, {kAluRI}, 0, 0, -1},
    {kIXOR_C7, 1, {kAluRI}, 0, 0, -1},  {kIADD_C8, 1, {kAluRI}, 0, 0, -1},  {kIXOR_C8, 1, {kAluRI}, 0, 0, -1},
    {kIADD_C9, 1, {kAluRI}, 0, 0, -1},  {kIXOR_C9, 1, {kAluRI}, 0, 0, -1},
    {kIMULH_R, 3, {kMovRR, kMulWide, kMovRRChained}, 1, 0, 1},
    {kISMULH_R, 3, {kMovRR, kMulWide, kMovRRChained}, 1, 0, 1},
    {kIMUL_RCP, 2, {kMovImm64, kImulRRChained}, 1, 1, -1}};
constexpr KindInfo kNoInstr{kNone, 0, {}, 0, 0, 0};

struct Fetch { int slots, id; int bytes[4]; };
constexpr Fetch kF484{3, 0, {4, 8, 4}}, kF7333{4, 1, {7, 3, 3, 3}}, kF3733{4, 2, {3, 7, 3, 3}}, kF493{3, 3, {4, 9, 3}},
    kF4444{4, 4, {4, 4, 4, 4}}, kF3310{3, 5, {3, 3, 10}};
constexpr const Fetch *kRandomFetch[4] = {&kF484, &kF7333, &kF3733, &kF493};

constexpr int kTargetLatency = 170, kPortMapCycles = kTargetLatency + 4, kMaxProgram = 3 * kTargetLatency + 2;
constexpr int kLookAhead = 4, kMaxDiscards = 256, kLeaNoDst = 5;

// the generator's byte stream: Blake2b-512 re-hashed over its own 64-byte state (spec §3.1.1 "BlakeGenerator")
class ByteStream {
public:
    ByteStream(const void *seed, size_t len) {
        memset(buf_, 0, sizeof buf_);
        memcpy(buf_, seed, std::min<size_t>(len, 60));
    }
    uint8_t byte() { need(1); return buf_[pos_++]; }
    uint32_t word() { need(4); const uint32_t v = rd32(buf_ + pos_); pos_ += 4; return v; }

private:
    void need(size_t n) {
        if (pos_ + n > 64) { uint8_t t[64]; blake2b(t, 64, buf_, 64); memcpy(buf_, t, 64); pos_ = 0; }
    }
    uint8_t buf_[64];
    size_t pos_ = 64;
};

struct RegState { int ready = 0; int last_group = kNone; int last_par = -1; };

struct Candidate {
    const KindInfo *info = &kNoInstr;
    int src = -1, dst = -1, mod = 0, group = kNone, group_par = 0;
    uint32_t imm = 0;
    bool dst_may_equal_src = false, par_follows_src = false;

    void make(const KindInfo *ki, ByteStream &rng) {
        *this = Candidate{};
        info = ki;
        switch (ki->kind) {
            case kISUB_R: group = kIADD_RS; par_follows_src = true; break;
            case kIXOR_R: group = kIXOR_R; par_follows_src = true; break;
            case kIADD_RS: mod = rng.byte(); group = kIADD_RS; par_follows_src = true; break;
            case kIMUL_R: group = kIMUL_R; par_follows_src = true; break;
            case kIROR_C: do { imm = rng.byte() & 63; } while (imm == 0); group = kIROR_C; group_par = -1; break;
            case kIADD_C7: case kIADD_C8: case kIADD_C9: imm = rng.word(); group = kIADD_C7; group_par = -1; break;
            case kIXOR_C7: case kIXOR_C8: case kIXOR_C9: imm = rng.word(); group = kIXOR_C7; group_par = -1; break;
            case kIMULH_R: dst_may_equal_src = true; group = kIMULH_R; group_par = (int)rng.word(); break;
            case kISMULH_R: dst_may_equal_src = true; group = kISMULH_R; group_par = (int)rng.word(); break;
            case kIMUL_RCP: do { imm = rng.word(); } while ((imm & (imm - 1)) == 0); group = kIMUL_RCP; group_par = -1; break;
            default: break;
        }
    }
    void make_for_slot(int bytes, int fetch_id, bool last_slot, ByteStream &rng) {
        static constexpr Kind s3[2] = {kISUB_R, kIXOR_R}, s3l[4] = {kISUB_R, kIXOR_R, kIMULH_R, kISMULH_R}, s4[2] = {kIROR_C, kIADD_RS},
                              s7[2] = {kIXOR_C7, kIADD_C7}, s8[2] = {kIXOR_C8, kIADD_C8}, s9[2] = {kIXOR_C9, kIADD_C9};
        switch (bytes) {
            case 3: make(&kInfo[last_slot ? s3l[rng.byte() & 3] : s3[rng.byte() & 1]], rng); break;
            case 4: if (fetch_id == 4 && !last_slot) make(&kInfo[kIMUL_R], rng); else make(&kInfo[s4[rng.byte() & 1]], rng); break;
            case 7: make(&kInfo[s7[rng.byte() & 1]], rng); break;
            case 8: make(&kInfo[s8[rng.byte() & 1]], rng); break;
            case 9: make(&kInfo[s9[rng.byte() & 1]], rng); break;
            default: make(&kInfo[kIMUL_RCP], rng); break;
        }
    }
    static bool pick(const int *pool, int n, ByteStream &rng, int &out) {
        if (n == 0) return false;
        out = pool[n > 1 ? rng.word() % (uint32_t)n : 0];
        return true;
    }
    bool choose_src(int cycle, const RegState (&regs)[8], ByteStream &rng) {
        int pool[8], n = 0;
        for (int i = 0; i < 8; i++) if (regs[i].ready <= cycle) pool[n++] = i;
        if (n == 2 && info->kind == kIADD_RS && (pool[0] == kLeaNoDst || pool[1] == kLeaNoDst)) { group_par = src = kLeaNoDst; return true; }
        if (!pick(pool, n, rng, src)) return false;
        if (par_follows_src) group_par = src;
        return true;
    }
    bool choose_dst(int cycle, bool allow_mul_chain, const RegState (&regs)[8], ByteStream &rng) {
        int pool[8], n = 0;
        for (int i = 0; i < 8; i++) {
            const RegState &r = regs[i];
            if (r.ready > cycle) continue;
            if (!dst_may_equal_src && i == src) continue;
            if (!allow_mul_chain && group == kIMUL_R && r.last_group == kIMUL_R) continue;
            if (r.last_group == group && r.last_par == group_par) continue;
            if (info->kind == kIADD_RS && i == kLeaNoDst) continue;
            pool[n++] = i;
        }
        return pick(pool, n, rng, dst);
    }
};

class PortMap {
public:
    PortMap() { memset(busy_, 0, sizeof busy_); }
    // earliest cycle >= `from` at which the macro-op can issue (all of its uops in the same cycle); -1 if none
    int place(const MacroOp &m, int from, int after, bool commit) {
        if (m.uop_a == 0) return from;                 // register move: eliminated, needs neither a port nor its operand early
        if (m.chained) from = std::max(from, after);
        if (m.uop_b == 0) return place_uop(m.uop_a, from, commit);
        for (; from < kPortMapCycles; from++) {
            const int a = place_uop(m.uop_a, from, false), b = place_uop(m.uop_b, from, false);
            if (a >= 0 && a == b) {
                if (commit) { place_uop(m.uop_a, a, true); place_uop(m.uop_b, b, true); }
                return a;
            }
        }
        return -1;
    }

private:
    int place_uop(int uop, int from, bool commit) {
        static constexpr int order[3] = {2, 0, 1};   // P5 first, then P0, P1 last: keep the multiplier port free
        static constexpr int bit[3] = {kP0, kP1, kP5};
        for (; from < kPortMapCycles; from++)
            for (int k : order)
                if ((uop & bit[k]) && !busy_[from][k]) { if (commit) busy_[from][k] = true; return from; }
        return -1;
    }
    bool busy_[kPortMapCycles][3];
};

void generate(SsProgram &prog, ByteStream &rng) {
    PortMap ports;
    RegState regs[8];
    Candidate cur;
    const Fetch *fetch = nullptr;
    int op_at = 0, cycle = 0, chain_cycle = 0, muls = 0, discards = 0;
    bool done = false;
    struct Raw { int kind, dst, src, mod; uint32_t imm; };
    std::vector<Raw> raw;
    for (int decode = 0; decode < kTargetLatency && !done && (int)raw.size() < kMaxProgram; decode++) {
        const Kind k = cur.info->kind;
        if (k == kIMULH_R || k == kISMULH_R) fetch = &kF3310;
        else if (muls < decode + 1) fetch = &kF4444;
        else if (k == kIMUL_RCP) fetch = (rng.byte() & 1) ? &kF484 : &kF493;
        else fetch = kRandomFetch[rng.byte() & 3];
        for (int slot = 0; slot < fetch->slots;) {
            const int cycle_at_slot = cycle;
            if (op_at >= cur.info->n) {
                if (done || (int)raw.size() >= kMaxProgram) break;
                cur.make_for_slot(fetch->bytes[slot], fetch->id, slot + 1 == fetch->slots, rng);
                op_at = 0;
            }
            const MacroOp &m = cur.info->ops[op_at];
            int at = ports.place(m, cycle, chain_cycle, false);
            if (at < 0) { done = true; break; }
            bool discarded = false, abandon = false;
            auto wait_for = [&](auto &&chooser) {
                int tries = 0;
                while (tries < kLookAhead && !chooser(at)) { tries++; at++; cycle++; }
                if (tries < kLookAhead) return;
                if (discards < kMaxDiscards) { discards++; op_at = cur.info->n; discarded = true; }
                else { cur = Candidate{}; abandon = true; }
            };
            if (op_at == cur.info->src_at) wait_for([&](int c) { return cur.choose_src(c, regs, rng); });
            if (!discarded && !abandon && op_at == cur.info->dst_at)
                wait_for([&](int c) { return cur.choose_dst(c, discards > 0, regs, rng); });
            if (discarded) continue;
            if (abandon) break;
            discards = 0;
            at = ports.place(m, at, at, true);
            if (at < 0) { done = true; break; }
            chain_cycle = at + m.latency;
            if (op_at == cur.info->result_at) {
                RegState &r = regs[cur.dst];
                r.ready = chain_cycle; r.last_group = cur.group; r.last_par = cur.group_par;
            }
            slot++; op_at++;
            if (at >= kTargetLatency) done = true;
            cycle = cycle_at_slot;
            if (op_at >= cur.info->n) {
                raw.push_back({cur.info->kind, cur.dst, cur.src >= 0 ? cur.src : cur.dst, cur.mod, cur.imm});
                muls += (cur.info->kind == kIMUL_R || cur.info->kind == kIMULH_R || cur.info->kind == kISMULH_R || cur.info->kind == kIMUL_RCP);
            }
        }
        cycle++;
    }
    // the register with the longest dependency chain (unit latencies, unlimited width) addresses the next cache line
    int depth[8] = {0};
    for (const Raw &o : raw) {
        const int via_dst = depth[o.dst] + 1, via_src = o.dst != o.src ? depth[o.src] + 1 : 0;
        depth[o.dst] = std::max(via_dst, via_src);
    }
    int deepest = 0;
    prog.address_reg = 0;
    for (int i = 0; i < 8; i++) if (depth[i] > deepest) { deepest = depth[i]; prog.address_reg = (uint32_t)i; }
    prog.ops.clear();
    for (const Raw &o : raw) {
        SsOp d{};
        d.dst = (uint8_t)o.dst; d.src = (uint8_t)o.src; d.imm32 = o.imm;
        switch (o.kind) {
            case kISUB_R: d.opcode = SS_ISUB_R; break;
            case kIXOR_R: d.opcode = SS_IXOR_R; break;
            case kIADD_RS: d.opcode = SS_IADD_RS; d.shift = (uint8_t)((o.mod >> 2) & 3); break;
            case kIMUL_R: d.opcode = SS_IMUL_R; break;
            case kIROR_C: d.opcode = SS_IROR_C; break;
            case kIADD_C7: case kIADD_C8: case kIADD_C9: d.opcode = SS_IADD_C; break;
            case kIXOR_C7: case kIXOR_C8: case kIXOR_C9: d.opcode = SS_IXOR_C; break;
            case kIMULH_R: d.opcode = SS_IMULH_R; break;
            case kISMULH_R: d.opcode = SS_ISMULH_R; break;
            default: d.opcode = SS_IMUL_RCP; d.rcp = reciprocal(o.imm); break;
        }
        prog.ops.push_back(d);
    }
}

}  // namespace

void blake2b(void *out, size_t outlen, const void *in, size_t inlen) {
    Blake2b h(outlen);
    h.update(in, inlen);
    h.finish(out);
}

uint64_t reciprocal(uint32_t divisor) {
    // floor(2^(63 + bits(divisor)) / divisor), by long division of 2^63 continued for bits(divisor) more binary digits
    uint64_t q = (1ull << 63) / divisor, r = (1ull << 63) % divisor;
    int bits = 0;
    for (uint32_t t = divisor; t; t >>= 1) bits++;
    for (int i = 0; i < bits; i++) {
        if (r >= divisor - r) { q = 2 * q + 1; r = 2 * r - divisor; }
        else { q = 2 * q; r = 2 * r; }
    }
    return q;
}

void build_cache(const void *key, size_t keylen, CacheImage &out) {
    out.memory.assign((size_t)kCacheKiB * 128, 0);
    argon2d_fill(reinterpret_cast<Block *>(out.memory.data()), kCacheKiB, kArgonPasses, key, (uint32_t)keylen, "RandomX\x03", 8);
    ByteStream rng(key, keylen);
    for (auto &p : out.programs) generate(p, rng);
}

}  // namespace rx
}  // namespace b200post

// compat_capi.cu — the verifying / proving half of the libpost-compatible symbol set (include/post_compat.h), on top of
// the batched verifier (verifier.cu), the prover (prover.cu) and the k2pow engine.  One proof per call, as libpost.
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/b200post_prove.h"
#include "../../include/b200post_verify.h"
#include "../../include/post_compat.h"
#include "engine.h"

using namespace b200post;

struct Verifier { uint32_t provider; };

namespace {

VerifyResult result(VerifyResultTag tag, uint32_t idx = 0) { VerifyResult r; r.tag = tag; r.invalid_index = idx; return r; }

VerifyResult run(const Verifier *v, const Proof &proof, const ProofMetadata *m, const ProofConfig &cfg, const InitConfig &ic,
                 const b200post_verify_options &opt) {
    if (!v || !m || (proof.indices.len && !proof.indices.ptr)) return result(VerifyInvalidArgument);
    if (ic.scrypt.r != 1 || ic.scrypt.p != 1) return result(VerifyInvalidArgument);
    b200post_proof p{proof.nonce, proof.indices.ptr, proof.indices.len, proof.pow};
    b200post_proof_metadata md;
    memcpy(md.node_id, m->node_id, 32); memcpy(md.commitment_atx_id, m->commitment_atx_id, 32); memcpy(md.challenge, m->challenge, 32);
    md.num_units = m->num_units; md.labels_per_unit = m->labels_per_unit;
    b200post_verify_params q{};
    q.k1 = cfg.k1; q.k2 = cfg.k2; q.scrypt_n = ic.scrypt.n;
    memcpy(q.pow_difficulty, cfg.pow_difficulty, 32);
    int status = B200POST_OK;
    uint64_t bad = 0;
    const int rc = b200post_verify_batch(v->provider, 1, &p, &md, &q, &opt, nullptr /* builtin k2pow check */, &status, &bad);
    if (rc != B200POST_OK) return result(VerifyFailed);
    switch (status) {
        case B200POST_OK: return result(VerifyOk);
        case B200POST_ERR_INVALID_PROOF: return bad == ~0ull ? result(VerifyFailed) : result(VerifyInvalidIndex, (uint32_t)bad);
        case B200POST_ERR_EMPTY_PROOF: case B200POST_ERR_INVALID_ARGUMENT: return result(VerifyInvalidArgument);
        default: return result(VerifyFailed);
    }
}

}  // namespace

extern "C" {

VerifyResult new_verifier(uint32_t /*flags*/, Verifier **out) {
    if (!out) return result(VerifyInvalidArgument);
    *out = nullptr;
    if (!engine_for(0)) return result(VerifyFailedToCreateVerifier);     // libpost's verifier is CPU-wide; ours lives on device 0
    *out = new (std::nothrow) Verifier{0};
    return *out ? result(VerifyOk) : result(VerifyFailedToCreateVerifier);
}

void free_verifier(Verifier *verifier) { delete verifier; }

VerifyResult verify_proof(const Verifier *verifier, Proof proof, const ProofMetadata *metadata, ProofConfig cfg, InitConfig init_cfg) {
    b200post_verify_options o{};
    o.mode = B200POST_VERIFY_ALL;
    return run(verifier, proof, metadata, cfg, init_cfg, o);
}

VerifyResult verify_proof_index(const Verifier *verifier, Proof proof, const ProofMetadata *metadata, ProofConfig cfg, InitConfig init_cfg,
                                size_t index) {
    if (index > 0xffffffffu) return result(VerifyInvalidArgument);
    b200post_verify_options o{};
    o.mode = B200POST_VERIFY_SELECTED_INDEX; o.selected_index = (uint32_t)index;
    return run(verifier, proof, metadata, cfg, init_cfg, o);
}

VerifyResult verify_proof_subset(const Verifier *verifier, Proof proof, const ProofMetadata *metadata, ProofConfig cfg, InitConfig init_cfg,
                                 size_t k3, const uint8_t *seed, size_t seed_len) {
    if (k3 > 0xffffffffu || (seed_len && !seed)) return result(VerifyInvalidArgument);
    b200post_verify_options o{};
    o.mode = B200POST_VERIFY_SUBSET; o.k3 = (uint32_t)k3; o.seed = seed; o.seed_len = seed_len;
    return run(verifier, proof, metadata, cfg, init_cfg, o);
}

Proof *generate_proof(const char *datadir, const uint8_t *challenge, ProofConfig cfg, size_t nonces, size_t /*threads*/, uint32_t /*pow_flags*/) {
    if (!datadir || !challenge || nonces == 0 || nonces > 4096) { set_error("generate_proof: invalid argument"); return nullptr; }
    b200post_post_config pc{};
    pc.k1 = cfg.k1; pc.k2 = cfg.k2; pc.k3 = cfg.k2;
    memcpy(pc.pow_difficulty, cfg.pow_difficulty, 32);
    b200post_prove_opts po{};
    po.provider = 0; po.nonces = (uint32_t)nonces; po.pow_mode = B200POST_POW_BUILTIN;
    b200post_proof_out out;
    if (b200post_generate_proof(datadir, challenge, &pc, &po, &out, nullptr, nullptr) != B200POST_OK) return nullptr;
    Proof *p = static_cast<Proof *>(malloc(sizeof(Proof)));
    uint8_t *buf = static_cast<uint8_t *>(malloc(out.indices_len ? out.indices_len : 1));
    if (!p || !buf) { free(p); free(buf); return nullptr; }
    memcpy(buf, out.indices, out.indices_len);
    p->nonce = out.nonce; p->pow = out.pow;
    p->indices.ptr = buf; p->indices.len = out.indices_len; p->indices.cap = out.indices_len;
    return p;
}

void free_proof(Proof *proof) {
    if (!proof) return;
    free(proof->indices.ptr);
    free(proof);
}

}  // extern "C"

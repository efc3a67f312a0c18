"""GPU tier: the batched PostVerifier against the oracle's restatement of the post-rs verifier.

Shapes follow the reference's verify tests: e2e params K1=12..K2=8 (activation/e2e/nipost_test.go:76-84) scaled so
that a brute-force prover finds proofs in a 1024-label space; the tamper test is
systest/tests/distributed_post_verification_test.go:254-267 (Indices[i] += 1 -> *verifying.ErrInvalidIndex)."""
import importlib
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vf(b2, gpu_ready):
    return importlib.import_module("go-spacemesh_b200.verify")


@pytest.fixture(scope="module")
def small_space(orc, vf):
    """A 4 x 256-label space at N=2 with a brute-forced valid proof per nonce."""
    rng = np.random.default_rng(42)
    node_id, atx, challenge = (bytes(rng.integers(0, 256, 32, dtype=np.uint8)) for _ in range(3))
    meta = vf.ProofMetadata(node_id, atx, challenge, num_units=4, labels_per_unit=256)
    params = vf.VerifyParams(k1=200, k2=8, scrypt_n=2)
    proofs = {}
    for nonce, pow_ in ((0, 0), (5, 77), (17, 2**40 + 3), (255, 1)):
        packed, hits = orc.py_prove(node_id, atx, challenge, 4, 256, params.k1, params.k2, 2, nonce=nonce, pow_=pow_)
        assert packed is not None, "prover found fewer than K2 hits; loosen K1"
        proofs[nonce] = (vf.Proof(nonce, packed, pow_), hits)
    return meta, params, proofs


def _oracle_verdict(orc, proof, meta, params, **kw):
    return orc.py_verify(proof.nonce, proof.indices, proof.pow, meta.node_id, meta.commitment_atx_id, meta.challenge,
                         meta.num_units, meta.labels_per_unit, params.k1, params.k2, params.scrypt_n, **kw)


def test_valid_proofs_pass_in_every_mode(vf, orc, small_space):
    meta, params, proofs = small_space
    v = vf.PostVerifier(pow="skip")
    try:
        for nonce, (proof, hits) in proofs.items():
            assert _oracle_verdict(orc, proof, meta, params) == (True, None)
            v.verify(proof, meta, params)
            v.verify(proof, meta, params, mode=vf.MODE_SUBSET, k3=3, seed=b"local-peer-id")
            for pos in range(params.k2):
                v.verify(proof, meta, params, mode=vf.MODE_SELECTED_INDEX, selected_index=pos, prioritized=True)
    finally:
        v.close()


def test_tampered_indices_are_reported(vf, orc, small_space):
    """Indices[i] += 1 for every position: verdict and failing POSITION (ErrInvalidIndex.Index) must equal the oracle's."""
    meta, params, proofs = small_space
    proof, hits = proofs[5]
    bits = vf.bits_per_index(meta.num_units * meta.labels_per_unit)
    v = vf.PostVerifier(pow="skip")
    try:
        seen_invalid = 0
        for pos in range(params.k2):
            bad = list(hits)
            bad[pos] = (bad[pos] + 1) % (meta.num_units * meta.labels_per_unit)
            tampered = vf.Proof(proof.nonce, vf.pack_indices(bad, bits), proof.pow)
            ok, idx = _oracle_verdict(orc, tampered, meta, params)
            if ok:
                v.verify(tampered, meta, params)
            else:
                seen_invalid += 1
                with pytest.raises(vf.ErrInvalidIndex) as e:
                    v.verify(tampered, meta, params)
                assert e.value.index == idx
                # SelectedIndex on the tampered position (activation/malfeasance.go:161-166)
                with pytest.raises(vf.ErrInvalidIndex) as e:
                    v.verify(tampered, meta, params, mode=vf.MODE_SELECTED_INDEX, selected_index=pos)
                assert e.value.index == pos
        assert seen_invalid >= params.k2 // 2
    finally:
        v.close()


def test_subset_selection_matches_oracle(vf, orc, small_space):
    """Corrupt everything: Subset(k3, seed) must flag exactly the index the oracle's shuffle reaches first."""
    meta, params, proofs = small_space
    proof, hits = proofs[17]
    bits = vf.bits_per_index(meta.num_units * meta.labels_per_unit)
    garbage = vf.Proof(proof.nonce, vf.pack_indices([(h * 7 + 3) % 1024 for h in hits], bits), proof.pow)
    v = vf.PostVerifier(pow="skip")
    try:
        for seed in (b"", b"peer-A", b"peer-B" * 5):
            for k3 in (1, 2, 5, 8, 20):
                ok, idx = _oracle_verdict(orc, garbage, meta, params, mode="subset", k3=k3, seed=seed)
                if ok:
                    v.verify(garbage, meta, params, mode=vf.MODE_SUBSET, k3=k3, seed=seed)
                else:
                    with pytest.raises(vf.ErrInvalidIndex) as e:
                        v.verify(garbage, meta, params, mode=vf.MODE_SUBSET, k3=k3, seed=seed)
                    assert e.value.index == idx, (seed, k3)
    finally:
        v.close()


def test_malformed_proofs(vf, b2, small_space):
    meta, params, proofs = small_space
    proof, _ = proofs[0]
    v = vf.PostVerifier(pow="skip")
    try:
        with pytest.raises(vf.ErrEmptyProof):                       # "proof indices are empty"
            v.verify(vf.Proof(0, b"", 0), meta, params)
        with pytest.raises(b2.B200PostError) as e:                  # one byte short
            v.verify(vf.Proof(0, proof.indices[:-1], 0), meta, params)
        assert e.value.code == b2.ERR_INVALID_ARGUMENT
        with pytest.raises(b2.B200PostError):
            v.verify(proof, meta, params, mode=vf.MODE_SELECTED_INDEX, selected_index=params.k2)
        with pytest.raises(b2.B200PostError):
            v.verify(proof, vf.ProofMetadata(meta.node_id, meta.commitment_atx_id, meta.challenge, 0, 256), params)
    finally:
        v.close()


def test_closed_verifier(vf, small_space):
    """post_verifier_test.go:37-62,64-91: Verify after Close -> "verifier is closed"; Close is idempotent."""
    meta, params, proofs = small_space
    proof, _ = proofs[0]
    v = vf.PostVerifier(pow="skip")
    v.verify(proof, meta, params)
    v.close()
    v.close()
    with pytest.raises(vf.ErrVerifierClosed) as e:
        v.verify(proof, meta, params)
    assert str(e.value) == "verifier is closed"


@pytest.mark.parametrize("multi", [False, True])
def test_concurrent_callers_are_coalesced(vf, orc, b2, small_space, multi):
    """Safe for concurrent use (post_verifier.go:227); concurrent Verify calls share GPU batches.  `multi`: one
    dispatcher with a worker per device (a one-GPU box lists its device twice: two workers, one engine)."""
    meta, params, proofs = small_space
    bits = vf.bits_per_index(1024)
    gpus = [p["id"] for p in b2.providers() if p["id"] != b2.CPU_PROVIDER_ID]
    v = vf.PostVerifier(pow="skip", providers=gpus if len(gpus) > 1 else gpus * 2) if multi else vf.PostVerifier(pow="skip")
    work = []
    for nonce, (proof, hits) in proofs.items():
        work.append((proof, None))
        bad = list(hits); bad[3] = (bad[3] + 1) % 1024
        t = vf.Proof(proof.nonce, vf.pack_indices(bad, bits), proof.pow)
        ok, idx = _oracle_verdict(orc, t, meta, params)
        work.append((t, None if ok else idx))
    errors = []

    def caller(k):
        try:
            for rep in range(6):
                proof, expect = work[(k + rep) % len(work)]
                try:
                    v.verify(proof, meta, params, prioritized=(k % 5 == 0))
                    got = None
                except vf.ErrInvalidIndex as e:
                    got = e.index
                if got != expect:
                    errors.append((k, rep, got, expect))
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=caller, args=(k,)) for k in range(32)]
    for t in threads: t.start()
    for t in threads: t.join()
    batches, n = v.stats()
    v.close()
    assert not errors, errors[:3]
    assert n == 32 * 6 and batches < n, (batches, n)


def test_pow_callback_contract(vf, small_space):
    meta, params, proofs = small_space
    proof, _ = proofs[17]
    seen = []

    def pow_ok(ctx, pow_, nonce_group, challenge8, difficulty, node_id):
        seen.append((pow_, nonce_group, bytes(challenge8[:8]), bytes(difficulty[:32]), bytes(node_id[:32])))
        return 0

    q = vf.VerifyParams(params.k1, params.k2, params.scrypt_n, pow_difficulty=bytes([0, 0x0d, 0xfb, 0x23]) + b"\xff" * 28)
    v = vf.PostVerifier(pow=pow_ok)
    v.verify(proof, meta, q)
    v.close()
    scaled = (int.from_bytes(q.pow_difficulty, "big") // meta.num_units).to_bytes(32, "big")
    assert seen == [(proof.pow, proof.nonce // 16, meta.challenge[:8], scaled, meta.node_id)]
    v = vf.PostVerifier(pow=lambda *a: 1)
    with pytest.raises(vf.ErrInvalidIndex) as e:
        v.verify(proof, meta, q)
    assert e.value.index == 2**64 - 1          # the pow, not a label, was rejected
    v.close()


def test_batch_at_mainnet_shape(vf, orc, b2):
    """BASELINE.json configs[2] shape, scaled down: proofs x K2=37 random indices in a 4-SU space at N=8192.
    Random indices are almost never valid; K1 is inflated so that verdicts are mixed.  A sample of proofs is
    checked in full against the oracle, every proof for the (cheap) structural invariants."""
    rng = np.random.default_rng(3)
    n_proofs, k2, num_labels = 192, 37, 2**34
    bits = vf.bits_per_index(num_labels)
    params = vf.VerifyParams(k1=2**31, k2=k2, scrypt_n=8192)       # difficulty 2^61: msb threshold 0x20
    proofs, metas, idxs = [], [], []
    for i in range(n_proofs):
        node_id, atx, ch = (bytes(rng.integers(0, 256, 32, dtype=np.uint8)) for _ in range(3))
        ix = [int(x) for x in rng.integers(0, num_labels, k2)]
        idxs.append(ix)
        proofs.append(vf.Proof(int(rng.integers(0, 288)), vf.pack_indices(ix, bits), int(rng.integers(0, 2**56))))
        metas.append(vf.ProofMetadata(node_id, atx, ch, 4, 2**32))
    opts = [dict(mode=vf.MODE_SUBSET, k3=4, seed=b"p") if i % 3 == 0 else dict() for i in range(n_proofs)]
    st, bad = vf.verify_batch(proofs, metas, params, options=opts, pow="skip")
    assert set(st) <= {b2.OK, b2.ERR_INVALID_PROOF}
    for i in range(n_proofs):
        if st[i] == b2.ERR_INVALID_PROOF:
            assert 0 <= bad[i] < k2          # a position in the K2 list (ErrInvalidIndex.Index), not a label index
    for i in list(range(0, n_proofs, 16)):
        kw = dict(mode="subset", k3=4, seed=b"p") if i % 3 == 0 else {}
        ok, idx = _oracle_verdict(orc, proofs[i], metas[i], params, **kw)
        assert (st[i] == b2.OK) == ok and (ok or bad[i] == idx), i


def test_device_epilogue_against_oracle_densely(vf, orc, b2):
    """The on-device AES epilogue (incl. the rare 'MSB equal -> lazy cipher' branch, ~1/256 labels) must give the
    oracle's verdict for every proof of a large random batch (N = 2 keeps the oracle side cheap)."""
    rng = np.random.default_rng(77)
    n_proofs, k2, num_units, lpu = 1500, 8, 4, 256
    bits = vf.bits_per_index(num_units * lpu)
    params = vf.VerifyParams(k1=900, k2=k2, scrypt_n=2)          # difficulty msb ~0xe0: most labels pass, verdicts mixed
    node_id, atx = (bytes(rng.integers(0, 256, 32, dtype=np.uint8)) for _ in range(2))
    proofs, metas = [], []
    for i in range(n_proofs):
        ch = bytes(rng.integers(0, 256, 32, dtype=np.uint8))
        ix = [int(x) for x in rng.integers(0, num_units * lpu, k2)]
        proofs.append(vf.Proof(int(rng.integers(0, 320)), vf.pack_indices(ix, bits), int(rng.integers(0, 2**60))))
        metas.append(vf.ProofMetadata(node_id, atx, ch, num_units, lpu))
    st, bad = vf.verify_batch(proofs, metas, params, pow="skip")
    n_ok = 0
    for i in range(n_proofs):
        ok, idx = _oracle_verdict(orc, proofs[i], metas[i], params)
        assert (st[i] == b2.OK) == ok and (ok or bad[i] == idx), i
        n_ok += ok
    assert 0 < n_ok < n_proofs


def test_batch_split_over_devices(vf, b2):
    """b200post_verify_batch_multi: same verdicts, in the caller's order, as the one-device batch.  On a one-GPU
    box the provider is listed twice, which still exercises the split, the per-device threads and the merge."""
    rng = np.random.default_rng(5)
    n_proofs, k2, num_units, lpu = 701, 8, 4, 256
    bits = vf.bits_per_index(num_units * lpu)
    params = vf.VerifyParams(k1=900, k2=k2, scrypt_n=2)
    proofs, metas = [], []
    for i in range(n_proofs):
        ident = bytes(rng.integers(0, 256, 96, dtype=np.uint8))
        ix = [int(x) for x in rng.integers(0, num_units * lpu, k2)]
        proofs.append(vf.Proof(int(rng.integers(0, 64)), vf.pack_indices(ix, bits) if i != 300 else b"", i))
        metas.append(vf.ProofMetadata(ident[:32], ident[32:64], ident[64:], num_units, lpu))
    gpus = [p["id"] for p in b2.providers() if p["id"] != b2.CPU_PROVIDER_ID]
    devs = gpus if len(gpus) > 1 else [gpus[0], gpus[0]]
    one = vf.verify_batch(proofs, metas, params, provider=gpus[0], pow="skip")
    many = vf.verify_batch(proofs, metas, params, providers=devs, pow="skip")
    assert one == many and one[0][300] == b2.ERR_EMPTY_PROOF and 0 < sum(1 for s in one[0] if s == b2.OK) < n_proofs
    assert vf.verify_batch(proofs[:1], metas[:1], params, providers=devs, pow="skip") == ([one[0][0]], [one[1][0]])
    with pytest.raises(b2.B200PostError):
        vf.verify_batch(proofs, metas, params, providers=[gpus[0], 12345], pow="skip")


def test_metrics_follow_the_work(vf, b2, small_space):
    import re
    meta, params, proofs = small_space
    proof, _ = proofs[0]

    def val(name):
        m = re.search(rf"^{re.escape(name)} (\S+)$", b2.metrics_text(), re.M)
        return float(m.group(1))

    before = {k: val(k) for k in ("b200post_verify_proofs_total", "b200post_labels_gather_total",
                                  "b200post_post_verification_seconds_count")}
    v = vf.PostVerifier(pow="skip")
    for _ in range(3):
        v.verify(proof, meta, params)
    v.close()
    assert val("b200post_verify_proofs_total") == before["b200post_verify_proofs_total"] + 3
    assert val("b200post_labels_gather_total") == before["b200post_labels_gather_total"] + 3 * params.k2
    assert val("b200post_post_verification_seconds_count") == before["b200post_post_verification_seconds_count"] + 3
    assert val("b200post_post_verification_waiting_total") == 0


def test_prioritized_calls_jump_the_queue(vf, orc):
    """post_verifier_test.go:93-142 (prioritised jobs are taken first): with one proof per GPU batch and a backlog
    of ordinary calls, a PrioritizedCall() submitted last must not finish last."""
    import time
    rng = np.random.default_rng(5)
    num_labels, k2 = 2**20, 12
    bits = vf.bits_per_index(num_labels)
    params = vf.VerifyParams(k1=2**19, k2=k2, scrypt_n=8192)       # every call costs a full N=8192 launch pair
    meta = vf.ProofMetadata(bytes(32), bytes(32), bytes(32), 1, num_labels)

    def mk():
        return vf.Proof(0, vf.pack_indices([int(x) for x in rng.integers(0, num_labels, k2)], bits), 0)

    v = vf.PostVerifier(pow="skip", max_batch_proofs=1)
    order, lock = [], threading.Lock()

    def call(tag, prioritized):
        try:
            v.verify(mk(), meta, params, prioritized=prioritized)
        except vf.ErrInvalidIndex:
            pass
        with lock:
            order.append(tag)

    normal = [threading.Thread(target=call, args=(f"n{i}", False)) for i in range(10)]
    for t in normal:
        t.start()
    time.sleep(0.05)                                                # let the backlog form
    pr = threading.Thread(target=call, args=("PRIO", True))
    pr.start()
    for t in normal + [pr]:
        t.join()
    batches, n = v.stats()
    v.close()
    assert n == 11 and batches == 11
    assert order.index("PRIO") <= 5, order


def test_builtin_k2pow_check(vf, b2, small_space):
    """Default verifier = the RandomX pow check on the device (activation/post_verifier.go:150-160): a mined pow passes,
    any other fails with the 'k2pow, not a label' marker; the difficulty is pow_difficulty / num_units."""
    k2 = importlib.import_module("go-spacemesh_b200.k2pow")
    meta, params, proofs = small_space
    proof, hits = proofs[17]
    easy = bytes([0x20]) + b"\x00" * 31                        # 1/8 of all hashes pass before scaling, 1/32 after / 4 units
    q = vf.VerifyParams(params.k1, params.k2, params.scrypt_n, pow_difficulty=easy)
    scaled = k2.scale_difficulty(easy, meta.num_units)
    found, _ = k2.search(proof.nonce // 16, meta.challenge[:8], meta.node_id, scaled, 0, 4096)
    assert found is not None
    # the labels of `proof` were proven under pow = proof.pow: keep the label check out of it (SELECTED_INDEX would still
    # bind to the pow through the AES key), so this test drives verify_batch on statuses only
    good = vf.Proof(proof.nonce, proof.indices, found)
    wrong = vf.Proof(proof.nonce, proof.indices, found + 1 if not k2.verify(found + 1, proof.nonce // 16, meta.challenge[:8], meta.node_id, scaled) else found + 2)
    st, bad = vf.verify_batch([good, wrong, wrong], [meta] * 3, q, pow="builtin")
    assert st[1] == st[2] == b2.ERR_INVALID_PROOF and bad[1] == bad[2] == vf.POW_INVALID
    assert st[0] in (b2.OK, b2.ERR_INVALID_PROOF) and bad[0] != vf.POW_INVALID   # pow accepted; labels judged under the new key
    st2, _ = vf.verify_batch([good], [meta], q, pow="skip")
    assert st2[0] == st[0]
    # through the queueing verifier (opts == builtin by default)
    v = vf.PostVerifier()
    try:
        with pytest.raises(vf.ErrInvalidIndex) as e:
            v.verify(wrong, meta, q)
        assert e.value.index == vf.POW_INVALID
        big = vf.Proof(proof.nonce, proof.indices, 2**56 + 5)     # does not fit the 7 hashed bytes
        with pytest.raises(vf.ErrInvalidIndex) as e:
            v.verify(big, meta, q)
        assert e.value.index == vf.POW_INVALID
    finally:
        v.close()


def test_pow_policy_is_explicit(vf, b2):
    """A NULL callback no longer means 'skip': CALLBACK without a function is refused, SKIP must be asked for."""
    with pytest.raises(b2.B200PostError) as e:
        vf.PostVerifier(pow="callback-missing")
    assert e.value.code == b2.ERR_UNSUPPORTED
    vf.PostVerifier(pow="skip").close()


def test_k2_above_16_bits_is_rejected(vf, b2):
    meta = vf.ProofMetadata(bytes(32), bytes(32), bytes(32), 1, 2**20)
    bits = vf.bits_per_index(2**20)
    for k2 in (65536, 70000):
        params = vf.VerifyParams(k1=10, k2=k2, scrypt_n=2)
        proof = vf.Proof(0, bytes((k2 * bits + 7) // 8), 0)
        st, _ = vf.verify_batch([proof], [meta], params, options=[dict(mode=vf.MODE_SUBSET, k3=2, seed=b"s")], pow="skip")
        assert st == [b2.ERR_INVALID_ARGUMENT]

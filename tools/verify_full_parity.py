"""Dev tool (GPU box): BASELINE.json configs[2] at full size — 10 000 proofs x K2 = 37 at N = 8192 — with EVERY
recomputed label and EVERY verdict compared with the oracle ("on the full 370 000-label set
once per kernel revision")."""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b2 = importlib.import_module("go-spacemesh_b200")
vf = importlib.import_module("go-spacemesh_b200.verify")
from oracle import pyoracle as orc

n_proofs, k2, num_labels, N = 10000, 37, 2**34, 8192
rng = np.random.default_rng(3)
bits = vf.bits_per_index(num_labels)
ids = rng.integers(0, 256, (n_proofs, 96), dtype=np.uint8)
idx = rng.integers(0, num_labels, (n_proofs, k2), dtype=np.uint64)
params = vf.VerifyParams(k1=2**32 - 1, k2=k2, scrypt_n=N)
proofs = [vf.Proof(int(i % 288), vf.pack_indices(idx[i].tolist(), bits), int(i)) for i in range(n_proofs)]
metas = [vf.ProofMetadata(ids[i, :32].tobytes(), ids[i, 32:64].tobytes(), ids[i, 64:].tobytes(), 4, 2**32) for i in range(n_proofs)]
batch = vf.PreparedBatch(proofs, metas, params)
t0 = time.perf_counter(); st, bad = batch.run(); t_gpu = time.perf_counter() - t0

comms = np.repeat(np.stack([np.frombuffer(b2.commitment(m.node_id, m.commitment_atx_id), dtype=np.uint8) for m in metas]), k2, axis=0)
flat = idx.reshape(-1)
t0 = time.perf_counter(); gpu_labels = b2.labels_gather(comms, flat, N); t_gather = time.perf_counter() - t0
t0 = time.perf_counter(); cpu_labels = orc.c_labels_gather(comms, flat, N); t_cpu = time.perf_counter() - t0
labels_ok = bool((gpu_labels == cpu_labels).all())
diff = orc.py_proving_difficulty(params.k1, num_labels)
mism = 0
for i in range(n_proofs):
    exp_bad = None
    for k in range(k2):
        if not orc.py_label_passes(cpu_labels[i * k2 + k].tobytes(), metas[i].challenge, proofs[i].nonce, proofs[i].pow, diff):
            exp_bad = int(idx[i, k]); break
    got_bad = None if st[i] == 0 else bad[i]
    mism += (exp_bad != got_bad)
res = dict(proofs=n_proofs, labels=int(flat.size), labels_equal=labels_ok, verdict_mismatches=mism, invalid=int(sum(1 for s in st if s)),
           gpu_batch_s=t_gpu, gpu_gather_s=t_gather, oracle_labels_s=t_cpu, oracle_threads=orc.default_threads())
print(json.dumps(res))
sys.exit(0 if labels_ok and mism == 0 else 1)

"""Dev tool (GPU box): one verify batch (2000 proofs x 37, N=8192) and one proving scan (2^22 labels, 288 nonces),
for an ncu launch list of the non-init kernels (K1, K2p, K3, K5, K6a, K6b)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b2 = importlib.import_module("go-spacemesh_b200")
vf = importlib.import_module("go-spacemesh_b200.verify")
pr = importlib.import_module("go-spacemesh_b200.prove")
rng = np.random.default_rng(3)
n, k2, nl = 2000, 37, 2**34
bits = vf.bits_per_index(nl)
ids = rng.integers(0, 256, (n, 96), dtype=np.uint8)
idx = rng.integers(0, nl, (n, k2), dtype=np.uint64)
proofs = [vf.Proof(int(i % 288), vf.pack_indices(idx[i].tolist(), bits), int(i)) for i in range(n)]
metas = [vf.ProofMetadata(ids[i, :32].tobytes(), ids[i, 32:64].tobytes(), ids[i, 64:].tobytes(), 4, 2**32) for i in range(n)]
st, _ = vf.verify_batch(proofs, metas, vf.VerifyParams(k1=2**32 - 1, k2=k2, scrypt_n=8192))
labels = rng.integers(0, 256, (1 << 22, 16), dtype=np.uint8)
try:
    pr.prove_scan(labels, bytes(32), 288, list(range(18)), 1, 37, 2**40)
except b2.B200PostError:
    pass
print("done", sum(1 for s in st if s))

"""Dev tool (GPU box): throughput of the proving scan over labels in host memory (full scan, no early exit)."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b2 = importlib.import_module("go-spacemesh_b200")
pr = importlib.import_module("go-spacemesh_b200.prove")
n = 1 << 26                                     # 1 GiB of labels
labels = np.random.default_rng(0).integers(0, 256, (n, 16), dtype=np.uint8)
for nonces in (16, 64, 288):
    pows = list(range(nonces // 16))
    for rep in range(2):
        t0 = time.perf_counter()
        try:
            pr.prove_scan(labels, bytes(32), nonces, pows, 1, 37, 2**40)   # difficulty ~2^24: practically no hits
        except b2.B200PostError as e:
            assert e.code == b2.ERR_INVALID_PROOF
        dt = time.perf_counter() - t0
    print(f"nonces={nonces}: {n / dt / 1e6:.1f} M labels/s = {n * 16 / dt / 1e9:.2f} GB/s of POST data ({dt:.3f} s per GiB)", flush=True)

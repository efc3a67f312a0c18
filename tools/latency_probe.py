"""Latency of small label jobs and small verify batches on cuda:0 (device time of the engine's stream, and wall time
of the C-ABI call with host buffers).  Usage: python tools/latency_probe.py"""
import importlib, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
pkg = importlib.import_module("go-spacemesh_b200")
vf = importlib.import_module("go-spacemesh_b200.verify")
rng = np.random.default_rng(9)
out = {"gather": [], "verify": []}
pkg.labels_gather(rng.integers(0, 256, (64, 32), dtype=np.uint8), rng.integers(0, 2**34, 64, dtype=np.uint64), 8192)   # warm-up
for lowlat in (4096, 0):
    pkg.set_option("lowlat_max_labels", lowlat)
    for n in (1, 37, 592, 2048, 4096):
        comms = rng.integers(0, 256, (n, 32), dtype=np.uint8); idx = rng.integers(0, 2**34, n, dtype=np.uint64)
        best_dev, best_wall = 1e9, 1e9
        for _ in range(3):
            t = time.perf_counter(); pkg.labels_gather(comms, idx, 8192); w = (time.perf_counter() - t) * 1e3
            best_wall = min(best_wall, w); best_dev = min(best_dev, pkg.last_call_ms())
        out["gather"].append({"lowlat_max_labels": lowlat, "labels": n, "device_ms": round(best_dev, 2), "wall_ms": round(best_wall, 2)})
        print(json.dumps(out["gather"][-1]), flush=True)
pkg.set_option("lowlat_max_labels", 4096)
out["rotate_form"] = []
for mask in (0, 1):          # 16/8-bit rotates as SHF (0) or PRMT (1): does the form change the serial chain's latency?
    pkg.set_option("rotate_mask", mask)
    comms = rng.integers(0, 256, (37, 32), dtype=np.uint8); idx = rng.integers(0, 2**34, 37, dtype=np.uint64)
    pkg.labels_gather(comms, idx, 8192)
    best = min((pkg.labels_gather(comms, idx, 8192), pkg.last_call_ms())[1] for _ in range(3))
    out["rotate_form"].append({"rotate_mask": mask, "labels": 37, "device_ms": round(best, 3)})
    print(json.dumps(out["rotate_form"][-1]), flush=True)
pkg.set_option("rotate_mask", 0)
k2, num_labels = 37, 2**34
bits = vf.bits_per_index(num_labels)
params = vf.VerifyParams(k1=2**31, k2=k2, scrypt_n=8192)
for n in (1, 16, 256, 10000):
    proofs, metas = [], []
    for i in range(n):
        node_id, atx, ch = (bytes(rng.integers(0, 256, 32, dtype=np.uint8)) for _ in range(3))
        proofs.append(vf.Proof(int(rng.integers(0, 288)), vf.pack_indices([int(x) for x in rng.integers(0, num_labels, k2)], bits), int(rng.integers(0, 2**56))))
        metas.append(vf.ProofMetadata(node_id, atx, ch, 4, 2**32))
    batch = vf.PreparedBatch(proofs, metas, params)
    for pow_mode in ("skip", "builtin"):
        best = 1e9
        for _ in range(2 if n > 1000 else 3):
            t = time.perf_counter(); batch.run(0, pow_mode); best = min(best, (time.perf_counter() - t) * 1e3)
        out["verify"].append({"proofs": n, "pow": pow_mode, "wall_ms": round(best, 2), "proofs_per_s": round(n / (best / 1e3), 1)})
        print(json.dumps(out["verify"][-1]), flush=True)
print(json.dumps(out))

"""k2pow nonce search over every GPU of the box (BASELINE.json configs[4]: the nonce range split over the devices,
no data-path collective).  One process, one host thread per device (b200post_k2pow_search_multi).
Prints one JSON line with the aggregate hashes/s over whole batches (wall clock around the call; datasets resident)."""
import importlib, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
b2 = importlib.import_module("go-spacemesh_b200")
k2 = importlib.import_module("go-spacemesh_b200.k2pow")
gpus = [p["id"] for p in b2.providers()]
rng = np.random.default_rng(5)
ch, node = bytes(rng.integers(0, 256, 8, dtype=np.uint8)), bytes(rng.integers(0, 256, 32, dtype=np.uint8))
t = time.perf_counter()
for g in gpus:
    k2.prepare(provider=g)
prep = time.perf_counter() - t
batch = k2.batch_size(gpus[0])
n = 3 * batch * len(gpus)
k2.search(0, ch, node, b"\x00" * 32, 0, batch * len(gpus), providers=gpus)          # warm-up: allocate the batches
t = time.perf_counter()
found, done = k2.search(0, ch, node, b"\x00" * 32, 0, n, providers=gpus)
wall = time.perf_counter() - t
one = None
if len(gpus) > 1:
    t = time.perf_counter(); _, d1 = k2.search(0, ch, node, b"\x00" * 32, 0, 3 * batch, provider=gpus[0]); one = d1 / (time.perf_counter() - t)
# correctness of the split: a threshold that exactly one nonce of a range meets is found by the multi-device search
hs = k2.hashes(1, ch, node, 0, 2 * batch + 5, provider=gpus[-1])
order = sorted(range(len(hs)), key=lambda i: bytes(hs[i]))
hit, _ = k2.search(1, ch, node, bytes(hs[order[1]]), 0, len(hs), providers=gpus)
print(json.dumps({"n_gpus": len(gpus), "hashes": done, "seconds": wall, "hashes_per_s": done / wall, "single_gpu_hashes_per_s": one,
                  "batch_per_gpu": batch, "dataset_prepare_all_s": prep, "found": found, "split_search_finds_the_unique_hit": hit == order[0]}))

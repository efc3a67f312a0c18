"""Dev tool (GPU box): SURVEY.md §8d cfg2 measurement — steady-state 4-SU-style init to a /dev/null sink.

Streams >= 2^28 labels (N = 8192) through consecutive initialize()-sized host-buffer calls (labels copied to host,
then dropped), starting below index 2^32 so that the run crosses the first space-unit boundary, and parity-checks
2^16 labels sampled from the produced range (seed 2) plus the range's edges and the last index of a 4-SU space
against the oracle."""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b2 = importlib.import_module("go-spacemesh_b200")
from oracle import pyoracle as orc

N, total = 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
n_check = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
commitment = b2.commitment(bytes(range(32)), bytes(range(32, 64)))
wave = b2.wave_slots(N)
batch = 16 * wave
start0 = 2**32 - total // 2
b2.labels_range(commitment, N, start0 - batch, batch, discard=True)       # warm-up; primes the pipeline for start0
keep = {}
rng = np.random.default_rng(2)
picks = np.unique(np.concatenate([rng.integers(0, total, n_check), [0, total - 1, total // 2 - 1, total // 2]])).astype(np.int64)
b2.romix_time(reset=True)
t0 = time.perf_counter()
done = 0
while done < total:
    n = min(batch, total - done)
    out, _ = b2.labels_range(commitment, N, start0 + done, n)              # host buffer, then dropped (/dev/null sink)
    sel = picks[(picks >= done) & (picks < done + n)]
    for i in sel:
        keep[int(i)] = out[int(i) - done].copy()
    done += n
wall = time.perf_counter() - t0
ms, k, lab = b2.romix_time(reset=True)
idx = np.array(sorted(keep), dtype=np.uint64)
got = np.stack([keep[int(i)] for i in idx])
comms = np.tile(np.frombuffer(commitment, dtype=np.uint8), (len(idx), 1))
t1 = time.perf_counter()
exp = orc.c_labels_gather(comms, idx + np.uint64(start0), N)
ok = bool((got == exp).all())
# the last index of a 4-SU space, through the scattered path
last = np.array([2**34 - 1], dtype=np.uint64)
ok_last = bool((b2.labels_gather(comms[:1], last, N) == orc.c_labels_gather(comms[:1], last, N)).all())
res = dict(labels=total, seconds=wall, labels_per_s=total / wall, MiB_per_s_of_post_data=total * 16 / wall / 2**20,
           romix_share=ms / 1e3 / wall, start_index=start0, crosses_2_32=True, checked=len(idx), parity_ok=ok, last_index_ok=ok_last,
           oracle_seconds=time.perf_counter() - t1, extrapolated_4SU_seconds=2**34 / (total / wall))
print(json.dumps(res))
sys.exit(0 if ok and ok_last else 1)

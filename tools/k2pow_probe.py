"""Times the k2pow engine on cuda:0: dataset build, then one full batch of RandomX hashes (device time from the
library's CUDA events).  Usage: python tools/k2pow_probe.py [vms_per_sm ...]"""
import importlib
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

pkg = importlib.import_module("go-spacemesh_b200")
k2 = importlib.import_module("go-spacemesh_b200.k2pow")
t = time.time(); k2.prepare(); t_ds = time.time() - t
rng = np.random.default_rng(1)
ch, node = bytes(rng.integers(0, 256, 8, dtype=np.uint8)), bytes(rng.integers(0, 256, 32, dtype=np.uint8))
out = {"dataset_s": round(t_ds, 3), "runs": []}
for per_sm in [int(a) for a in sys.argv[1:]] or [256]:
    pkg.set_option("rx_vms_per_sm", per_sm)
    n = k2.batch_size()
    t = time.time()
    found, done = k2.search(0, ch, node, b"\x00" * 32, 0, n)
    wall = time.time() - t
    tm = k2.last_timing()
    out["runs"].append({"vms_per_sm": per_sm, "batch": n, "wall_s": round(wall, 3), "device_ms": round(tm["total_ms"], 1),
                        "vm_kernel_ms": round(tm["vm_kernel_ms"], 1), "hashes_per_s": round(done / (tm["total_ms"] / 1e3), 1)})
    print(json.dumps(out["runs"][-1]), flush=True)
print(json.dumps(out))

"""Times the k2pow engine on cuda:0: dataset build, then one full batch of RandomX hashes per setting (device time from
the library's CUDA events).  Usage: python tools/k2pow_probe.py [mode:vms_per_sm ...]   e.g. 2:32 1:64 0:256"""
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
ref = None
for arg in sys.argv[1:] or ["1:48"]:
    mode, per_sm = (int(x) for x in arg.split(":"))
    pkg.set_option("rx_vm_mode", mode)
    pkg.set_option("rx_vms_per_sm", per_sm)
    n = k2.batch_size()
    hs = k2.hashes(0, ch, node, 0, n)
    tm = k2.last_timing()
    if ref is None:
        ref = hs[:64].copy()
    same = bool((hs[:64] == ref).all())
    out["runs"].append({"mode": mode, "vms_per_sm": per_sm, "batch": n, "device_ms": round(tm["total_ms"], 1),
                        "vm_kernel_ms": round(tm["vm_kernel_ms"], 1), "hashes_per_s": round(n / (tm["total_ms"] / 1e3), 1),
                        "same_hashes_as_first_run": same})
    print(json.dumps(out["runs"][-1]), flush=True)
print(json.dumps(out))

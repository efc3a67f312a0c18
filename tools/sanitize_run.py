"""Dev tool (GPU box): a small workload touching every kernel and every ROMix variant, for
`compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_run.py` (the reference runs its unit
tests under the Go race detector, Makefile:106; this is the native-code analogue)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b2 = importlib.import_module("go-spacemesh_b200")
from oracle import pyoracle as orc
b2.set_option("max_scratch_mib", 256)
c = bytes(range(32))
ok = True
for variant, tpbs in ((4, (64, 128, 256, 512)), (1, (128,)), (0, (128,)), (2, (128,))):
    for tpb in tpbs:
        b2.set_option("romix_variant", variant); b2.set_option("tpb", tpb)
        n, start, count = 32, 2**32 - 40, 1200 if variant == 4 else 300
        diff = orc.py_vrf_difficulty(64)
        got, vrf = b2.labels_range(c, n, start, count, vrf_difficulty_=diff)
        exp, f, bi, bl = orc.c_labels_range(c, n, start, count, diff)
        good = bool((got == exp).all()) and vrf == ((bi, bl) if f else None)
        ok &= good
        print(f"variant {variant} tpb {tpb}: {'ok' if good else 'MISMATCH'}", flush=True)
b2.set_option("romix_variant", 4); b2.set_option("tpb", 512)
rng = np.random.default_rng(1)
comms = rng.integers(0, 256, (700, 32), dtype=np.uint8); idx = rng.integers(0, 2**34, 700, dtype=np.uint64)
good = bool((b2.labels_gather(comms, idx, 16) == orc.c_labels_gather(comms, idx, 16)).all())
ok &= good
print("gather:", "ok" if good else "MISMATCH", flush=True)
# round 2: the low-latency ROMix kernel (small jobs) ...
b2.set_option("lowlat_max_labels", 4096)
for n_items in (1, 37, 600):
    comms = rng.integers(0, 256, (n_items, 32), dtype=np.uint8); idx = rng.integers(0, 2**34, n_items, dtype=np.uint64)
    good = bool((b2.labels_gather(comms, idx, 16) == orc.c_labels_gather(comms, idx, 16)).all())
    ok &= good
    print(f"lowlat gather {n_items}:", "ok" if good else "MISMATCH", flush=True)
# ... and the k2pow (RandomX) kernels: a handful of VMs through dataset build, fill, program decode, the VM, the final hash
if os.environ.get("SANITIZE_K2POW", "1") != "0":
    k2 = importlib.import_module("go-spacemesh_b200.k2pow")
    b2.set_option("rx_vms_per_sm", 1)
    got = k2.randomx_hash(b"test key 000", [b"This is a test"])[0].hex()
    good = got == "639183aae1bf4c9a35884cb46b09cad9175f04efd7684e7262a0ac1c2f0b4e3f"
    ok &= good
    print("k2pow RandomX KAT:", "ok" if good else "MISMATCH " + got, flush=True)
sys.exit(0 if ok else 1)

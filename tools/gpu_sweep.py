"""Dev tool (GPU box): parity + timing sweep of the ROMix kernel variants through the C ABI.

usage: python tools/gpu_sweep.py parity <variant>      -> gpurun_out/parity_v<variant>.json
       python tools/gpu_sweep.py sweep  <spec.json|->  -> gpurun_out/sweep_<tag>.json
Uses the oracle only as the checker (same rule as tests/).
"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b2 = importlib.import_module("go-spacemesh_b200")
from oracle import pyoracle as orc  # noqa: E402

OUT = "gpurun_out"
os.makedirs(OUT, exist_ok=True)


def parity(variant: int, mws=(0, 1)):
    res = {"variant": variant, "cases": []}
    rng = np.random.default_rng(7)
    commitment = bytes(rng.integers(0, 256, 32, dtype=np.uint8))
    ok_all = True
    for mw in mws:
        b2.set_option("romix_variant", variant)
        b2.set_option("rotate_mask", mw)
        for n, start, count in [(2, 0, 1024), (2, 2**32 - 100, 333), (16, 5, 4100), (1024, 2**40, 515), (8192, 2**32 - 64, 160)]:
            t = time.time()
            diff = orc.py_vrf_difficulty(max(count // 4, 2))
            got, vrf = b2.labels_range(commitment, n, start, count, vrf_difficulty_=diff)
            exp, f, bi, bl = orc.c_labels_range(commitment, n, start, count, diff)
            ok = bool((got == exp).all()) and ((vrf is None and not f) or (vrf is not None and f and vrf == (bi, bl)))
            ok_all &= ok
            res["cases"].append(dict(mw=mw, n=n, start=start, count=count, ok=ok, vrf=str(vrf)[:40], exp_vrf=str((bi, bl))[:40], s=round(time.time() - t, 3)))
            print(res["cases"][-1], flush=True)
        # gather
        m = 700
        comms = rng.integers(0, 256, (m, 32), dtype=np.uint8)
        idx = rng.integers(0, 2**34, m, dtype=np.uint64)
        for n in (4, 8192):
            k = m if n == 4 else 96
            got = b2.labels_gather(comms[:k], idx[:k], n)
            exp = orc.c_labels_gather(comms[:k], idx[:k], n)
            ok = bool((got == exp).all())
            ok_all &= ok
            res["cases"].append(dict(mw=mw, gather_n=n, items=k, ok=ok))
            print(res["cases"][-1], flush=True)
    res["ok"] = ok_all
    json.dump(res, open(f"{OUT}/parity_v{variant}.json", "w"), indent=1)
    print("PARITY", variant, "OK" if ok_all else "FAIL")
    return ok_all


def sweep(specs, tag):
    """specs: list of dicts(variant, mw, tpb, ctas, n=8192, waves=2)"""
    commitment = bytes(range(32))
    prov = b2.providers()[0]
    results = []
    for sp in specs:
        n = sp.get("n", 8192)
        try:
            b2.set_option("romix_variant", sp["variant"]); b2.set_option("rotate_mask", sp["mw"])
            b2.set_option("tpb", sp["tpb"]); b2.set_option("ctas_per_sm", sp["ctas"])
            b2.set_option("debug_skip_phase", sp.get("skip", 0)); b2.set_option("dr_unroll", sp.get("dr", 4))
            slots = b2.wave_slots(n)
            waves = sp.get("waves", 2)
            b2.labels_range(commitment, n, 0, slots, discard=True)  # warm-up layer: allocates scratch
            b2.romix_time(reset=True)
            t0 = time.time()
            b2.labels_range(commitment, n, slots, slots * waves, discard=True)
            wall = time.time() - t0
            dev_ms = b2.last_call_ms()
            ms, k, lab = b2.romix_time(reset=True)
            lps_k2 = lab / (ms / 1e3)
            r = dict(sp, slots=slots, romix_ms=round(ms / max(k, 1), 3), launches=k, labels_per_s_k2=round(lps_k2),
                     labels_per_s_call=round(slots * waves / (dev_ms / 1e3)), labels_per_s_wall=round(slots * waves / wall),
                     GBps=round(lps_k2 * (256 * n + 16) / 1e9, 1))
        except Exception as e:  # noqa: BLE001
            r = dict(sp, error=str(e))
        results.append(r)
        print(r, flush=True)
        json.dump(results, open(f"{OUT}/sweep_{tag}.json", "w"), indent=1)
    return results


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "parity":
        sys.exit(0 if parity(int(sys.argv[2])) else 1)
    elif mode == "sweep":
        tag = sys.argv[2]
        specs = json.load(open(sys.argv[3])) if len(sys.argv) > 3 else json.loads(sys.stdin.read())
        sweep(specs, tag)

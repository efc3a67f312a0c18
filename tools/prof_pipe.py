"""Dev tool (GPU box): run a few layers of the N=8192 init with explicit options so ncu can capture K2p.
usage: python tools/prof_pipe.py key=value ... [layers=4]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b2 = importlib.import_module("go-spacemesh_b200")
layers = 4
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "layers": layers = int(v)
    else: b2.set_option(k, int(v, 0))
slots = b2.wave_slots(8192)
b2.labels_range(bytes(range(32)), 8192, 0, slots * layers, discard=True)
ms, k, lab = b2.romix_time()
print(f"slots {slots}: {k} ROMix launches, {ms / k:.2f} ms avg, {lab / ms * 1e3:.0f} labels/s (under profiler if any)")

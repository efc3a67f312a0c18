"""Dev tool (GPU box): run a few full waves of the N=8192 init so ncu can capture the ROMix kernel.
usage: python tools/prof_one.py [variant] [waves] [ctas] [skip]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b2 = importlib.import_module("go-spacemesh_b200")
variant = int(sys.argv[1]) if len(sys.argv) > 1 else b2.get_option("romix_variant")
waves = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if len(sys.argv) > 3: b2.set_option("ctas_per_sm", int(sys.argv[3]))
if len(sys.argv) > 4: b2.set_option("debug_skip_phase", int(sys.argv[4]))
b2.set_option("romix_variant", variant)
slots = b2.wave_slots(8192)
b2.labels_range(bytes(range(32)), 8192, 0, slots * waves, discard=True)
ms, k, lab = b2.romix_time()
print(f"variant {variant}: {k} ROMix launches, {ms / k:.2f} ms each, {lab / ms * 1e3:.0f} labels/s (under profiler if any)")

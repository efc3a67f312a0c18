"""One small k2pow batch for an ncu capture of the VM kernel (tools/, not product)."""
import importlib, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
pkg = importlib.import_module("go-spacemesh_b200")
k2 = importlib.import_module("go-spacemesh_b200.k2pow")
pkg.set_option("rx_vms_per_sm", int(sys.argv[1]) if len(sys.argv) > 1 else 256)
n = k2.batch_size()
print(k2.search(0, bytes(8), bytes(32), b"\x00" * 32, 0, n), k2.last_timing())

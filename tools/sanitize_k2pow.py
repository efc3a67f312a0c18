"""Dev tool (GPU box): RandomX's own known answers through the k2pow engine, for
`compute-sanitizer --tool racecheck|memcheck --kernel-regex kns=execute_kernel python tools/sanitize_k2pow.py [mode]`
(only the VM kernel instrumented: the 2 GiB dataset build under racecheck would take hours).  A second launch runs
3 VMs (a full 2-warp CTA and a half-empty one) and checks that VM 0 is unchanged by its neighbours."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b2 = importlib.import_module("go-spacemesh_b200")
k2 = importlib.import_module("go-spacemesh_b200.k2pow")
b2.set_option("rx_vms_per_sm", 1)
b2.set_option("rx_vm_mode", int(sys.argv[1]) if len(sys.argv) > 1 else 1)
got = k2.randomx_hash(b"test key 000", [b"This is a test"])[0].hex()
ok = got == "639183aae1bf4c9a35884cb46b09cad9175f04efd7684e7262a0ac1c2f0b4e3f"
print("RandomX KAT (1 VM):", "ok" if ok else "MISMATCH " + got, flush=True)
three = k2.randomx_hash(b"test key 000", [b"This is a test", b"This is a tesu", b"This is a tesv"])
same = three[0].hex() == got and len({bytes(h) for h in three}) == 3
print("3 VMs in one launch:", "ok" if same else "MISMATCH", flush=True)
sys.exit(0 if ok and same else 1)

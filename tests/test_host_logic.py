"""CPU tier: the product's host-side logic and the kernels' arithmetic (compiled for the host)."""
import ctypes
import importlib
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_device_arithmetic_on_host_matches_oracle(host_emul, orc):
    """post_device.cuh (Keccak-f, the two PBKDF2 passes, ChaCha20/8, BlockMix exactly as the kernels inline them) == oracle."""
    rng = np.random.default_rng(21)
    for n in (2, 4, 64, 1024, 8192):
        for idx in (0, 1, 2**32 - 1, 2**32, 2**40 + 12345, 2**64 - 1, int(rng.integers(0, 2**63))):
            c = bytes(rng.integers(0, 256, 32, dtype=np.uint8))
            out = ctypes.create_string_buffer(32)
            host_emul.emul_label32(c, ctypes.c_uint64(idx), n, out)
            assert out.raw == orc.c_label32(c, idx, n), (n, idx)


def test_device_arithmetic_golden(host_emul, golden):
    for it in golden["gather"]["items"]:
        out = ctypes.create_string_buffer(32)
        host_emul.emul_label32(bytes.fromhex(it["commitment"]), ctypes.c_uint64(it["index"]), it["N"], out)
        assert out.raw.hex() == it["label32"]
    # the kernels' arithmetic reproduces the real VRF-nonce labels of the reference's checkpoint fixture
    for it in golden["checkpoint_vrf"]["items"][::3]:
        out = ctypes.create_string_buffer(32)
        host_emul.emul_label32(bytes.fromhex(it["commitment"]), ctypes.c_uint64(it["vrf_nonce"]), it["N"], out)
        assert out.raw.hex() == it["label32"]
    rng = np.random.default_rng(5)
    for _ in range(4):      # the interleaved fill+mix step equals the two separate steps
        seed = (ctypes.c_uint32 * 64)(*[int(x) for x in rng.integers(0, 2**32, 64, dtype=np.uint64)])
        assert host_emul.emul_dual_step_matches(seed) == 1


def test_library_exports_every_declared_symbol(b2):
    """The C-ABI library loads on a CPU-only box and exports everything include/*.h declares."""
    lib = b2.lib()
    declared = set()
    for hdr in ("b200post.h", "post_compat.h", "b200post_verify.h", "b200post_setup.h", "b200post_prove.h", "b200post_poet.h", "b200post_k2pow.h"):
        p = ROOT / "include" / hdr
        if not p.exists():
            continue
        text = re.sub(r"/\*.*?\*/", "", p.read_text(), flags=re.S)
        text = re.sub(r"typedef[^;{]*\(\s*\*[^;]*;", "", text)        # function-pointer typedefs are not symbols
        declared |= set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text))
    declared -= {"defined"}
    assert {"b200post_labels_range", "b200post_labels_gather", "initialize", "new_initializer"} <= declared
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_commitment_and_difficulty_host_helpers(b2, orc, golden):
    rng = np.random.default_rng(22)
    for _ in range(8):
        a = bytes(rng.integers(0, 256, 32, dtype=np.uint8)); b = bytes(rng.integers(0, 256, 32, dtype=np.uint8))
        assert b2.commitment(a, b) == orc.py_commitment(a, b)
    for n, hx in golden["vrf_difficulty"].items():
        assert b2.vrf_difficulty(int(n)).hex() == hx


def test_no_cpu_fallback(b2):
    """Without a GPU every compute entry point must fail loudly; the CPU provider id is refused outright."""
    if b2.providers():
        pytest.skip("a CUDA device is present")
    with pytest.raises(b2.B200PostError) as e:
        b2.labels_range(bytes(32), 2, 0, 8)
    assert e.value.code == b2.ERR_NO_DEVICE
    with pytest.raises(b2.B200PostError) as e:
        b2.labels_gather(np.zeros((1, 32), np.uint8), np.zeros(1, np.uint64), 2)
    assert e.value.code == b2.ERR_NO_DEVICE
    with pytest.raises(b2.B200PostError) as e:
        b2.labels_gather_indexed(np.zeros((1, 32), np.uint8), np.zeros(1, np.uint32), np.zeros(1, np.uint64), 2)
    assert e.value.code == b2.ERR_NO_DEVICE
    with pytest.raises(b2.B200PostError) as e:
        b2.labels_range(bytes(32), 2, 0, 8, provider=b2.CPU_PROVIDER_ID)
    assert e.value.code == b2.ERR_UNSUPPORTED
    vf = importlib.import_module("go-spacemesh_b200.verify")
    proof, meta = vf.Proof(0, vf.pack_indices([1, 2], 11), 0), vf.ProofMetadata(bytes(32), bytes(32), bytes(32), 4, 256)
    params = vf.VerifyParams(k1=200, k2=2, scrypt_n=2)
    for kw in ({"provider": 0}, {"providers": [0, 1]}):
        with pytest.raises(b2.B200PostError) as e:
            vf.verify_batch([proof, proof], [meta, meta], params, **kw)
        assert e.value.code == b2.ERR_NO_DEVICE
    with pytest.raises(b2.B200PostError) as e:
        vf.PostVerifier(providers=[0, 1])
    assert e.value.code == b2.ERR_NO_DEVICE


def test_argument_validation(b2):
    for n in (0, 1, 3, 12, 2**21, 2**32):
        with pytest.raises(b2.B200PostError) as e:
            b2.labels_range(bytes(32), n, 0, 8)
        assert e.value.code == b2.ERR_INVALID_ARGUMENT
    with pytest.raises(b2.B200PostError) as e:
        b2.labels_range(bytes(32), 2, 2**64 - 4, 8)   # index overflow: last index would be 2^64 + 3
    assert e.value.code == b2.ERR_INVALID_ARGUMENT
    with pytest.raises(b2.B200PostError) as e:     # a row past the end of the commitment table
        b2.labels_gather_indexed(np.zeros((2, 32), np.uint8), np.array([2], np.uint32), np.zeros(1, np.uint64), 2)
    assert e.value.code == b2.ERR_INVALID_ARGUMENT
    with pytest.raises(b2.B200PostError):
        b2.set_option("romix_variant", 9)
    with pytest.raises(b2.B200PostError):
        b2.set_option("rotate_mask", 2)
    with pytest.raises(b2.B200PostError):
        b2.set_option("no_such_option", 1)


def test_product_does_not_reference_the_oracle():
    """The shipped sources must not import, link or call anything under oracle/."""
    pkg = ROOT / "go-spacemesh_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")) + list(pkg.rglob("Makefile")):
        if "build" in p.parts:
            continue
        text = p.read_text()
        assert "post_oracle" not in text and "pyoracle" not in text and "from oracle" not in text, p


def test_verify_helpers_match_oracle(b2, orc):
    """Index packing / difficulty helpers of the verify path (pure host code) against the Python restatement."""
    import importlib
    vf = importlib.import_module("go-spacemesh_b200.verify")
    rng = np.random.default_rng(23)
    for n in (1, 2, 3, 1000, 1024, 2**34, 2**34 + 1, 2**64 - 1):
        assert vf.bits_per_index(n) == orc.py_bits_per_index(n)
        for k1 in (1, 26, 2**31):
            assert vf.proving_difficulty(k1, n) == orc.py_proving_difficulty(k1, n)
    assert vf.bits_per_index(0) == 0
    for bits in (1, 7, 8, 11, 34, 35, 63, 64):
        idx = [int(x) for x in rng.integers(0, 2**min(bits, 63), 37)]
        packed = vf.pack_indices(idx, bits)
        assert packed == orc.py_pack_indices(idx, bits)
        assert vf.unpack_indices(packed, bits, 37) == idx
    # wire cap: 37 x 35-bit indices of a 4-SU space fit the 800-byte Indices cap (activation/wire/wire_v1.go:43)
    assert len(vf.pack_indices([0] * 37, vf.bits_per_index(2**34))) == 162


def test_verifier_needs_a_device(b2):
    import importlib
    vf = importlib.import_module("go-spacemesh_b200.verify")
    if b2.providers():
        pytest.skip("a CUDA device is present")
    with pytest.raises(b2.B200PostError) as e:
        vf.PostVerifier()
    assert e.value.code == b2.ERR_NO_DEVICE


def test_metrics_text_exposition(b2):
    """Prometheus text with the reference's POST metric names (activation/metrics/metrics.go:40-52)."""
    t = b2.metrics_text()
    for name in ("b200post_labels_range_total", "b200post_post_verification_waiting_total",
                 'b200post_post_verification_seconds_bucket{le="1"}', 'b200post_post_verification_seconds_bucket{le="512"}',
                 'b200post_post_verification_seconds_bucket{le="+Inf"}', "b200post_post_verification_seconds_count"):
        assert name in t
    for line in t.splitlines():
        assert line.startswith("#") or len(line.split(" ")) == 2


def test_reference_label_checker_matches_oracle(b2, orc, golden):
    """b200post_reference_label (the fault detector's CPU checker inside the product) == oracle; pinned on real data too."""
    rng = np.random.default_rng(4)
    for n in (2, 64, 8192):
        for idx in (0, 2**32, 2**64 - 1, int(rng.integers(0, 2**62))):
            c = bytes(rng.integers(0, 256, 32, dtype=np.uint8))
            assert b2.reference_label(c, idx, n) == orc.c_label32(c, idx, n)
    for it in golden["checkpoint_vrf"]["items"][:4]:
        assert b2.reference_label(bytes.fromhex(it["commitment"]), it["vrf_nonce"], it["N"]).hex() == it["label32"]


def test_vrf_comm_needs_a_device_and_valid_arguments(b2):
    """b200post_vrf_comm_* (the C library's NCCL min-reduce for one-process-per-GPU hosts): argument checks on a CPU box."""
    L = b2.lib()
    L.b200post_vrf_comm_init.argtypes = [ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
    comm = ctypes.c_void_p()
    assert L.b200post_vrf_comm_init(0, 2, 2, bytes(128), ctypes.byref(comm)) == b2.ERR_INVALID_ARGUMENT      # rank >= world
    assert L.b200post_vrf_comm_init(0, 0, 1, None, ctypes.byref(comm)) == b2.ERR_INVALID_ARGUMENT
    if not b2.providers():             # (with a device this would load libnccl into the test process: left to tests/test_gpu_nccl_vrf.py)
        rc = L.b200post_vrf_comm_init(0, 0, 1, bytes(128), ctypes.byref(comm))
        assert rc == b2.ERR_NO_DEVICE and not comm
    L.b200post_vrf_comm_free(None)

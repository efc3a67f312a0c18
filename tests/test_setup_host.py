"""CPU tier: PostSetupManager call-sequence / validation behaviour that needs no GPU
(mirrors activation/post_test.go:77-103, 105-116, 141-166, 168-178, 180-205)."""
import importlib

import pytest


@pytest.fixture()
def su(b2):
    return importlib.import_module("go-spacemesh_b200.setup")


def _opts(su, tmp_path, **kw):
    d = dict(data_dir=str(tmp_path / "post"), num_units=2, max_file_size=4096, provider_id=0, scrypt_n=2,
             compute_batch_size=1 << 10)
    d.update(kw)
    return su.PostSetupOpts(**d)


NODE, ATX = bytes(range(32)), bytes(range(32, 64))


def test_initial_status(su):
    mgr = su.PostSetupManager()
    st = mgr.status()
    assert st.state == su.STATE_NOT_STARTED and st.num_labels_written == 0


def test_prepare_initializer_validates_options(su, b2, tmp_path):
    """post_test.go:77-103: ComputeBatchSize = 3, NumUnits out of [Min, Max], Scrypt.N = 0 are rejected."""
    mgr = su.PostSetupManager()
    mgr.prepare_initializer(_opts(su, tmp_path), NODE, ATX)          # good options: no error
    mgr2 = su.PostSetupManager()
    for bad in (dict(compute_batch_size=3), dict(num_units=mgr2.cfg.max_num_units + 1), dict(num_units=mgr2.cfg.min_num_units - 1),
                dict(scrypt_n=0), dict(scrypt_n=12), dict(scrypt_r=8), dict(max_file_size=100), dict(data_dir="")):
        with pytest.raises(b2.B200PostError) as e:
            mgr2.prepare_initializer(_opts(su, tmp_path, **bad), NODE, ATX)
        assert e.value.code == b2.ERR_INVALID_ARGUMENT, bad
        assert mgr2.status().state == su.STATE_ERROR                 # post_test.go:168-178 StateError


def test_call_sequence_is_enforced(su, b2, tmp_path):
    """post_test.go:141-166."""
    mgr = su.PostSetupManager()
    with pytest.raises(b2.B200PostError) as e:                       # StartSession before PrepareInitializer
        mgr.start_session()
    assert e.value.code == su.ERR_STATE and "post session not prepared" in str(e.value)
    mgr.prepare_initializer(_opts(su, tmp_path), NODE, ATX)
    assert mgr.status().state == su.STATE_PREPARED
    with pytest.raises(b2.B200PostError) as e:                       # PrepareInitializer twice
        mgr.prepare_initializer(_opts(su, tmp_path), NODE, ATX)
    assert e.value.code == su.ERR_STATE and "post setup session in progress" in str(e.value)


def test_start_session_without_provider(su, b2, tmp_path):
    """post_test.go:105-116: prepare works without a provider, StartSession fails with "no provider specified"."""
    mgr = su.PostSetupManager()
    mgr.prepare_initializer(_opts(su, tmp_path, provider_id=None), NODE, ATX)
    with pytest.raises(b2.B200PostError) as e:
        mgr.start_session()
    assert e.value.code == su.ERR_NO_PROVIDER and "no provider specified" in str(e.value)
    assert mgr.status().state == su.STATE_ERROR


def test_metadata_written_by_prepare_and_pins_commitment(su, b2, tmp_path):
    """post.go:374-377: an existing postdata_metadata.json decides the commitment ATX; another identity is refused."""
    mgr = su.PostSetupManager()
    o = _opts(su, tmp_path)
    mgr.prepare_initializer(o, NODE, ATX)
    md = su.load_metadata(o.data_dir)
    assert md["node_id"] == NODE and md["commitment_atx_id"] == ATX and md["labels_per_unit"] == 512
    assert md["num_units"] == 2 and md["scrypt_n"] == 2 and md["nonce"] is None
    mgr2 = su.PostSetupManager()
    mgr2.prepare_initializer(o, NODE, b"\x07" * 32)                  # different ATX offered: metadata wins
    assert mgr2.commitment_atx() == ATX
    mgr3 = su.PostSetupManager()
    with pytest.raises(b2.B200PostError) as e:
        mgr3.prepare_initializer(o, b"\x01" * 32, ATX)               # someone else's data
    assert e.value.code == su.ERR_CONFIG_MISMATCH
    with pytest.raises(b2.B200PostError):
        su.load_metadata(str(tmp_path / "nowhere"))


def test_reset_before_any_session(su, b2):
    mgr = su.PostSetupManager()
    with pytest.raises(b2.B200PostError) as e:
        mgr.reset()
    assert e.value.code == su.ERR_STATE

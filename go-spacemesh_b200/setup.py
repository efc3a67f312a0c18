"""PostSetupManager over libb200post.so — host-side mirror of activation.PostSetupManager
(activation/post.go:185-449; interface postSetupProvider, activation/interface.go:114-119) for tests/tools.

Method names and the state machine are the reference's: prepare_initializer / start_session / status / reset,
states NotStarted(1) .. Error(6).  The state machine itself lives in C++ (csrc/setup.cu); this is ctypes glue.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

from . import B200PostError, ERR_CANCELLED, OK, lib

(STATE_NOT_STARTED, STATE_PREPARED, STATE_IN_PROGRESS, STATE_STOPPED, STATE_COMPLETE, STATE_ERROR) = range(1, 7)
ERR_STATE, ERR_NO_PROVIDER, ERR_IO, ERR_LABEL_MISMATCH, ERR_CONFIG_MISMATCH = 10, 11, 12, 13, 14
PROVIDER_UNSET, PROVIDER_ALL = -1, -2


class _PostConfig(ctypes.Structure):
    _fields_ = [("min_num_units", ctypes.c_uint32), ("max_num_units", ctypes.c_uint32), ("labels_per_unit", ctypes.c_uint64),
                ("k1", ctypes.c_uint32), ("k2", ctypes.c_uint32), ("k3", ctypes.c_uint32), ("pow_difficulty", ctypes.c_uint8 * 32)]


class _SetupOpts(ctypes.Structure):
    _fields_ = [("data_dir", ctypes.c_char_p), ("num_units", ctypes.c_uint32), ("max_file_size", ctypes.c_uint64),
                ("provider_id", ctypes.c_int64), ("scrypt_n", ctypes.c_uint64), ("scrypt_r", ctypes.c_uint64),
                ("scrypt_p", ctypes.c_uint64), ("compute_batch_size", ctypes.c_uint64), ("self_check_every", ctypes.c_uint32)]


class _Status(ctypes.Structure):
    _fields_ = [("state", ctypes.c_int32), ("num_labels_written", ctypes.c_uint64)]


class _Metadata(ctypes.Structure):
    _fields_ = [("node_id", ctypes.c_uint8 * 32), ("commitment_atx_id", ctypes.c_uint8 * 32), ("labels_per_unit", ctypes.c_uint64),
                ("num_units", ctypes.c_uint32), ("max_file_size", ctypes.c_uint64), ("scrypt_n", ctypes.c_uint64),
                ("scrypt_r", ctypes.c_uint64), ("scrypt_p", ctypes.c_uint64), ("has_nonce", ctypes.c_uint32),
                ("nonce", ctypes.c_uint64), ("nonce_value", ctypes.c_uint8 * 32), ("last_position", ctypes.c_uint64)]


@dataclass
class PostConfig:            # activation/post.go:27-38
    min_num_units: int = 1
    max_num_units: int = 10
    labels_per_unit: int = 512
    k1: int = 26
    k2: int = 37
    k3: int = 37
    pow_difficulty: bytes | None = None   # None = the library default (config/mainnet.go:41's prefix)


@dataclass
class PostSetupOpts:         # activation/post.go:53-61
    data_dir: str = ""
    num_units: int = 2
    max_file_size: int = 4 << 30
    provider_id: int | None = None
    scrypt_n: int = 8192
    scrypt_r: int = 1
    scrypt_p: int = 1
    compute_batch_size: int = 1 << 20
    self_check_every: int = 16


@dataclass
class PostSetupStatus:       # activation/post.go:121-125
    state: int
    num_labels_written: int


def _bind():
    L = lib()
    if getattr(L, "_setup_bound", False):
        return L
    vp = ctypes.c_void_p
    L.b200post_setup_manager_new.argtypes = [ctypes.POINTER(_PostConfig), ctypes.POINTER(vp)]
    L.b200post_setup_manager_free.argtypes = [vp]
    L.b200post_setup_manager_free.restype = None
    L.b200post_setup_prepare_initializer.argtypes = [vp, ctypes.POINTER(_SetupOpts), ctypes.c_char_p, ctypes.c_char_p]
    L.b200post_setup_start_session.argtypes = [vp, vp]
    L.b200post_setup_get_status.argtypes = [vp, ctypes.POINTER(_Status)]
    L.b200post_setup_reset.argtypes = [vp]
    L.b200post_setup_commitment_atx.argtypes = [vp, vp]
    L.b200post_load_metadata.argtypes = [ctypes.c_char_p, ctypes.POINTER(_Metadata)]
    L.b200post_default_post_config.argtypes = [ctypes.POINTER(_PostConfig)]
    L.b200post_default_post_config.restype = None
    L._setup_bound = True
    return L


def _err(rc: int):
    if rc != OK:
        raise B200PostError(rc, lib().b200post_last_error().decode(errors="replace"))


def load_metadata(data_dir: str) -> dict:
    """initialization.LoadMetadata."""
    m = _Metadata()
    _err(_bind().b200post_load_metadata(data_dir.encode(), ctypes.byref(m)))
    return dict(node_id=bytes(m.node_id), commitment_atx_id=bytes(m.commitment_atx_id), labels_per_unit=m.labels_per_unit,
                num_units=m.num_units, max_file_size=m.max_file_size, scrypt_n=m.scrypt_n,
                nonce=int(m.nonce) if m.has_nonce else None, nonce_value=bytes(m.nonce_value) if m.has_nonce else None,
                last_position=m.last_position)


class PostSetupManager:
    def __init__(self, cfg: PostConfig | None = None):
        L = _bind()
        cfg = cfg or PostConfig()
        c = _PostConfig()
        L.b200post_default_post_config(ctypes.byref(c))
        c.min_num_units, c.max_num_units, c.labels_per_unit = cfg.min_num_units, cfg.max_num_units, cfg.labels_per_unit
        c.k1, c.k2, c.k3 = cfg.k1, cfg.k2, cfg.k3
        self.cfg = cfg
        self._h = ctypes.c_void_p()
        _err(L.b200post_setup_manager_new(ctypes.byref(c), ctypes.byref(self._h)))

    def prepare_initializer(self, opts: PostSetupOpts, node_id: bytes, commitment_atx_id: bytes) -> None:
        o = _SetupOpts(opts.data_dir.encode(), opts.num_units, opts.max_file_size,
                       PROVIDER_UNSET if opts.provider_id is None else opts.provider_id,
                       opts.scrypt_n, opts.scrypt_r, opts.scrypt_p, opts.compute_batch_size, opts.self_check_every)
        _err(_bind().b200post_setup_prepare_initializer(self._h, ctypes.byref(o), node_id, commitment_atx_id))

    def start_session(self, cancel: ctypes.c_int | None = None) -> None:
        """Blocking.  `cancel` = a ctypes.c_int another thread sets to 1 (ctx cancel); raises code ERR_CANCELLED."""
        _err(_bind().b200post_setup_start_session(self._h, ctypes.addressof(cancel) if cancel is not None else None))

    def status(self) -> PostSetupStatus:
        s = _Status()
        _err(_bind().b200post_setup_get_status(self._h, ctypes.byref(s)))
        return PostSetupStatus(int(s.state), int(s.num_labels_written))

    def reset(self) -> None:
        _err(_bind().b200post_setup_reset(self._h))

    def commitment_atx(self) -> bytes:
        out = ctypes.create_string_buffer(32)
        _err(_bind().b200post_setup_commitment_atx(self._h, out))
        return out.raw

    def __del__(self):
        try:
            if self._h:
                _bind().b200post_setup_manager_free(self._h)
                self._h = None
        except Exception:  # noqa: BLE001
            pass

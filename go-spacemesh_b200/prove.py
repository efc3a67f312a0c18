"""Proof generation over libb200post.so — the AES-scan half of activation.PostClient.Proof
(activation/interface.go:204-207; api/grpcserver/post_client.go:69-143 polls the external post-service for it).

`generate_proof(data_dir, challenge, cfg)` returns (verify.Proof, verify.ProofMetadata) ready for
PostVerifier.verify.  The k2pow (RandomX upstream) is a caller-supplied hook; conventions are ASSUMED
(include/b200post_prove.h)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import B200PostError, OK, lib
from .setup import PostConfig, _PostConfig, _bind as _bind_setup
from .verify import Proof, ProofMetadata, _Meta

POW_PROVE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_uint8, ctypes.POINTER(ctypes.c_uint8),
                                ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint64))


class _ProveOpts(ctypes.Structure):
    _fields_ = [("provider", ctypes.c_uint32), ("nonces", ctypes.c_uint32), ("chunk_labels", ctypes.c_uint64),
                ("pow_prove", POW_PROVE_FN), ("pow_ctx", ctypes.c_void_p), ("pow_mode", ctypes.c_uint32),
                ("pow_cache_key", ctypes.c_char_p), ("pow_cache_key_len", ctypes.c_size_t)]


class _ProofOut(ctypes.Structure):
    _fields_ = [("nonce", ctypes.c_uint32), ("pow", ctypes.c_uint64), ("indices_len", ctypes.c_size_t),
                ("indices", ctypes.c_uint8 * 800), ("labels_scanned", ctypes.c_uint64)]


def _bind():
    L = lib()
    if getattr(L, "_prove_bound", False):
        return L
    L.b200post_generate_proof.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(_PostConfig), ctypes.POINTER(_ProveOpts),
                                          ctypes.POINTER(_ProofOut), ctypes.POINTER(_Meta), ctypes.c_void_p]
    L.b200post_prove_scan.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_uint32,
                                      ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.POINTER(_ProofOut)]
    L._prove_bound = True
    return L


def _c_cfg(cfg: PostConfig) -> _PostConfig:
    c = _PostConfig()
    _bind_setup().b200post_default_post_config(ctypes.byref(c))
    c.min_num_units, c.max_num_units, c.labels_per_unit = cfg.min_num_units, cfg.max_num_units, cfg.labels_per_unit
    c.k1, c.k2, c.k3 = cfg.k1, cfg.k2, cfg.k3
    if getattr(cfg, "pow_difficulty", None) is not None:
        ctypes.memmove(c.pow_difficulty, cfg.pow_difficulty, 32)
    return c


def _err(rc):
    if rc != OK:
        raise B200PostError(rc, lib().b200post_last_error().decode(errors="replace"))


def generate_proof(data_dir: str, challenge: bytes, cfg: PostConfig, *, provider: int = 0, nonces: int = 16,
                   chunk_labels: int = 0, pow="builtin"):
    """PostClient.Proof(ctx, challenge) -> (Post, PostInfo-like metadata); also returns labels scanned.
    pow: "builtin" (k2pow search on the device, the library default), "skip" (pow = 0, explicit) or a callable
    (ctx, nonce_group, challenge8, difficulty32, node_id32, pow_out) -> 0."""
    L = _bind()
    if callable(pow):
        cb, mode = POW_PROVE_FN(pow), 1
    else:
        cb, mode = ctypes.cast(None, POW_PROVE_FN), {"builtin": 0, "skip": 2, "callback-missing": 1}[pow]
    opts = _ProveOpts(provider, nonces, chunk_labels, cb, None, mode, None, 0)
    out, meta, c = _ProofOut(), _Meta(), _c_cfg(cfg)
    _err(L.b200post_generate_proof(data_dir.encode(), challenge, ctypes.byref(c), ctypes.byref(opts), ctypes.byref(out),
                                   ctypes.byref(meta), None))
    proof = Proof(int(out.nonce), bytes(out.indices[: out.indices_len]), int(out.pow))
    pm = ProofMetadata(bytes(meta.node_id), bytes(meta.commitment_atx_id), bytes(meta.challenge), int(meta.num_units),
                       int(meta.labels_per_unit))
    return proof, pm, int(out.labels_scanned)


def prove_scan(labels: np.ndarray, challenge: bytes, nonces: int, pows, k1: int, k2: int, num_labels: int, *,
               first_index: int = 0, provider: int = 0):
    """The scan alone over labels in host memory (uint8[n,16])."""
    labels = np.ascontiguousarray(labels, dtype=np.uint8).reshape(-1, 16)
    arr = (ctypes.c_uint64 * len(pows))(*[int(p) for p in pows])
    out = _ProofOut()
    _err(_bind().b200post_prove_scan(provider, labels.ctypes.data, first_index, labels.shape[0], challenge, nonces, arr, k1, k2,
                                     num_labels, ctypes.byref(out)))
    return int(out.nonce), bytes(out.indices[: out.indices_len]), int(out.pow), int(out.labels_scanned)

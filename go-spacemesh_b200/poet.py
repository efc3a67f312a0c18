"""PoET registration PoW search over libb200post.so (shared.FindSubmitPowNonce, activation/poet.go:529-535)."""
from __future__ import annotations

import ctypes

from . import B200PostError, OK, lib


def _bind():
    L = lib()
    if not getattr(L, "_poet_bound", False):
        L.b200post_poet_pow_find.argtypes = [ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                                             ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                                             ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.c_void_p]
        L.b200post_poet_pow_hash.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p,
                                             ctypes.c_uint64, ctypes.c_void_p]
        L.b200post_poet_pow_hash.restype = None
        L._poet_bound = True
    return L


def pow_hash(pow_challenge: bytes, poet_challenge: bytes, node_id: bytes, nonce: int) -> bytes:
    out = ctypes.create_string_buffer(32)
    _bind().b200post_poet_pow_hash(pow_challenge, len(pow_challenge), poet_challenge, len(poet_challenge), node_id, nonce, out)
    return out.raw


def find_submit_pow_nonce(pow_challenge: bytes, poet_challenge: bytes, node_id: bytes, difficulty: int, *, provider: int = 0,
                          start_nonce: int = 0, max_nonces: int = 1 << 40):
    """Returns (lowest valid nonce, candidates evaluated)."""
    nonce, hashes = ctypes.c_uint64(0), ctypes.c_uint64(0)
    rc = _bind().b200post_poet_pow_find(provider, pow_challenge, len(pow_challenge), poet_challenge, len(poet_challenge), node_id,
                                        difficulty, start_nonce, max_nonces, ctypes.byref(nonce), ctypes.byref(hashes), None)
    if rc != OK:
        raise B200PostError(rc, lib().b200post_last_error().decode(errors="replace"))
    return int(nonce.value), int(hashes.value)

"""TEST INFRASTRUCTURE (oracle): numpy twin of the counter-based dropout mask of the CUDA kernels
(csrc/common.cuh: drop_bits / drop_keep).  Gives the tests the exact keep-mask a kernel used, so dropout can be
checked against the reference's arithmetic (torch.nn.functional.dropout semantics: zero with probability p, scale
the rest by 1/(1-p); self_attention.py:40, encoder_block.py:67-69, decoder_block.py:132,234) element by element."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def thr16_of(p: float) -> int:
    return int(round(p * 65536.0))


def scale_of(thr16: int) -> float:
    return float(np.float32(65536.0) / (np.float32(65536.0) - np.float32(thr16)))


def _mul(a, c):
    return (a.astype(np.uint64) * np.uint64(c)) & M32


def _base(seed: int, site: int) -> np.uint64:
    """drop_base(seed, site) of common.cuh: the avalanche-mixed per-(step, call) base of the hash."""
    h = (int(seed & 0xFFFFFFFF) * 0x9E3779B1 + int(site) * 0x7F4A7C15 + 0x165667B1) & 0xFFFFFFFF
    h ^= h >> 16
    h = (h * 0x7FEB352D) & 0xFFFFFFFF
    h ^= h >> 15
    h = (h * 0x846CA68B) & 0xFFFFFFFF
    h ^= h >> 16
    return np.uint64(h)


def keep_mask(seed: int, site: int, rows, cols, thr16: int) -> np.ndarray:
    """rows: int array [R] (global row ids), cols: int array [C] -> bool [R, C] (True = kept)."""
    rows = np.asarray(rows, dtype=np.uint64)[:, None]
    cols = np.asarray(cols, dtype=np.uint64)[None, :]
    h = _base(seed, site)
    h = h ^ _mul(rows, 0x85EBCA77)
    h = h ^ _mul(cols >> np.uint64(1), 0xC2B2AE3D)
    h = h ^ (h >> np.uint64(16))
    h = _mul(h, 0x7FEB352D)
    h = h ^ (h >> np.uint64(15))
    h = _mul(h, 0x846CA68B)
    h = h ^ (h >> np.uint64(16))
    bits = np.where((cols & np.uint64(1)) == 1, h >> np.uint64(16), h & np.uint64(0xFFFF))
    return bits >= np.uint64(thr16)

"""TEST / BENCH SUPPORT ONLY -- the position-addressable synthetic text of SURVEY.md 8(d) (see tests/support/acgen.cu).

generate_text(..., device_ptr=...) fills device memory through tests/support/libacgen.so (a CUDA kernel); without device_ptr the
text is produced on the host by the numpy restatement below -- no library at all is loaded, which is what bench.py's reference arm
and cpu_baseline leg use.  host_text_c() runs the C loop of libacgen.so (tests hold the three against each other)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def build(quiet=True):
    subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL if quiet else None)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libacgen.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is missing: run `make -C tests/support` (done by __graft_entry__.build())")
        _LIB = ctypes.CDLL(path)
        vp, u64, i = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
        _LIB.acgen_generate_text.restype = i
        _LIB.acgen_generate_text.argtypes = [vp, i, u64, u64, i, u64, u64, u64, vp, vp, u64, vp]
    return _LIB


def _mix64(x):
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(30))
        x = x * np.uint64(0xBF58476D1CE4E5B9)
        x = x ^ (x >> np.uint64(27))
        x = x * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def _raw64(seed, k):
    with np.errstate(over="ignore"):
        return _mix64(np.uint64(seed) + k.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15))


def host_text(nb, first=0, kind=0, seed=0xC0FFEE, plant_seed=0x5EED, plant_period=0, dict_flat=None, dict_offsets=None):
    """numpy restatement of gen_byte (acgen.cu) for positions [first, first + nb)."""
    nb, first = int(nb), int(first)
    w0, w1 = first >> 3, (first + nb + 7) >> 3
    words = _raw64(seed, np.arange(w0, w1, dtype=np.uint64))
    out = words.view(np.uint8)[first - 8 * w0: first - 8 * w0 + nb].copy()  # little-endian: byte i mod 8 of word i / 8
    if kind == 1:
        out = (0x20 + ((out.astype(np.uint32) * 95) >> 8)).astype(np.uint8)
    nbk = 0 if dict_offsets is None else len(dict_offsets) - 1
    if plant_period and nbk:
        flat = np.asarray(dict_flat, dtype=np.uint8)
        offs = np.asarray(dict_offsets, dtype=np.uint64)
        b0 = max(first // plant_period - 1, 0)
        b1 = (first + nb - 1) // plant_period if nb else b0 - 1
        blocks = np.arange(b0, b1 + 1, dtype=np.uint64)
        r = _raw64(plant_seed, blocks)
        kws = ((r >> np.uint64(20)) % np.uint64(nbk)).astype(np.int64)
        ats = ((r & np.uint64(0xFFFFF)) % np.uint64(plant_period)).astype(np.int64)
        # ascending block order: a plant that spills into the next period is overwritten there by that period's own plant,
        # but only inside [at, at + len) of the later plant -- exactly gen_byte's "own block first, then the previous one"
        for b, kw, at in zip(blocks.astype(np.int64), kws, ats):
            lo, hi = int(offs[kw]), int(offs[kw + 1])
            start = int(b) * plant_period + int(at)
            stop = min(start + (hi - lo), (int(b) + 2) * plant_period)  # gen_byte looks back one period only
            a, z = max(start, first), min(stop, first + nb)
            if a < z:
                out[a - first: z - first] = flat[lo + (a - start): lo + (z - start)]
    return out


def host_text_c(nb, first=0, kind=0, seed=0xC0FFEE, plant_seed=0x5EED, plant_period=0, dict_flat=None, dict_offsets=None):
    """The same text from the C loop of libacgen.so (cross-check of the numpy restatement)."""
    nbk = 0 if dict_offsets is None else len(dict_offsets) - 1
    df = np.ascontiguousarray(dict_flat, dtype=np.uint8) if nbk else None
    do = np.ascontiguousarray(dict_offsets, dtype=np.uint64) if nbk else None
    out = np.empty(nb, dtype=np.uint8)
    rc = _lib().acgen_generate_text(out.ctypes.data, 0, first, nb, kind, seed, plant_seed, plant_period if nbk else 0, df.ctypes.data if nbk else None,
                                    do.ctypes.data if nbk else None, nbk, None)
    if rc:
        raise RuntimeError(f"acgen_generate_text failed ({rc})")
    return out


def generate_text(nb, first=0, kind=0, seed=0xC0FFEE, plant_seed=0x5EED, plant_period=0, dict_flat=None, dict_offsets=None, device_ptr=None, stream=None):
    """Host array (numpy path) unless device_ptr is given (CUDA kernel of libacgen.so writing device memory)."""
    if device_ptr is None:
        return host_text(nb, first, kind, seed, plant_seed, plant_period, dict_flat, dict_offsets)
    nbk = 0 if dict_offsets is None else len(dict_offsets) - 1
    df = np.ascontiguousarray(dict_flat, dtype=np.uint8) if nbk else None
    do = np.ascontiguousarray(dict_offsets, dtype=np.uint64) if nbk else None
    rc = _lib().acgen_generate_text(device_ptr, 1, first, nb, kind, seed, plant_seed, plant_period if nbk else 0, df.ctypes.data if nbk else None,
                                    do.ctypes.data if nbk else None, nbk, stream)
    if rc:
        raise RuntimeError(f"acgen_generate_text failed (cudaError {rc})")
    return None

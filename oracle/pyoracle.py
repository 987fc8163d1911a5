"""TEST INFRASTRUCTURE ONLY -- ctypes access to the oracles.

kinds:
  "ref_meyer", "ref_classic" : the UNMODIFIED reference source compiled in oracle/_ref/ (oracle/Makefile) and driven
                               through its public API by oracle/ref_harness.c;
  "port"                     : oracle/ac_port.c, the plain-C restatement.

Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import this module; the
product (aho-corasick-1975_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MATCH_DTYPE = np.dtype([("end", "<u8"), ("id", "<u4"), ("len", "<u4")])

_PATHS = {
    "ref_meyer": (os.path.join(HERE, "_ref", "libacref_meyer.so"), "refh"),
    "ref_classic": (os.path.join(HERE, "_ref", "libacref_classic.so"), "refh"),
    "port": (os.path.join(HERE, "libacport.so"), "acport"),
}
_LIBS = {}


def build(quiet=True):
    """Builds oracle/libacport.so and, when /root/reference is present, oracle/_ref/ (building the checker is not using it)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.DEVNULL if quiet else None)


def available(kind):
    return os.path.exists(_PATHS[kind][0])


def _lib(kind):
    if kind in _LIBS:
        return _LIBS[kind]
    path, pfx = _PATHS[kind]
    if not os.path.exists(path):
        raise FileNotFoundError(f"oracle library {path} is not built (run `make -C oracle`)")
    L = ctypes.CDLL(path)
    f = lambda name: getattr(L, f"{pfx}_{name}")
    vp, u64, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32
    f("create").restype, f("create").argtypes = vp, [ctypes.c_size_t]
    f("release").restype, f("release").argtypes = None, [vp]
    f("insert").restype, f("insert").argtypes = u32, [vp, vp, ctypes.c_size_t]
    f("insert_many").restype, f("insert_many").argtypes = None, [vp, vp, vp, ctypes.c_size_t, vp]
    f("nb_keywords").restype, f("nb_keywords").argtypes = ctypes.c_size_t, [vp]
    f("reset_cursor").restype, f("reset_cursor").argtypes = None, [vp]
    f("scan").restype, f("scan").argtypes = u64, [vp, vp, u64, u64, vp, u64, ctypes.c_int]
    f("scan_mt").restype, f("scan_mt").argtypes = u64, [vp, vp, u64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    f("incremental").restype, f("incremental").argtypes = ctypes.c_int, []
    if pfx == "refh":
        L.refh_foreach_ranks.restype, L.refh_foreach_ranks.argtypes = ctypes.c_size_t, [vp, vp, ctypes.c_size_t]
    if pfx == "acport":
        L.acport_scan_lead.restype = u64
        L.acport_scan_lead.argtypes = [vp, vp, u64, u64, u64, vp, u64, ctypes.c_int, ctypes.POINTER(u32)]
        L.acport_prepare.restype, L.acport_prepare.argtypes = None, [vp]
        L.acport_lmax.restype, L.acport_lmax.argtypes = u32, [vp]
        L.acport_nb_states.restype, L.acport_nb_states.argtypes = u32, [vp]
    _LIBS[kind] = (L, f)
    return _LIBS[kind]


def _sym_dtype(width):
    return {1: np.uint8, 2: np.uint16, 4: np.uint32}[width]


def pack_keywords(keywords, width):
    """list of bytes / 1-D arrays -> (flat symbol array, uint64 offsets[nb+1])."""
    dt = _sym_dtype(width)
    arrs = [np.frombuffer(k, dtype=dt) if isinstance(k, (bytes, bytearray)) else np.asarray(k, dtype=dt) for k in keywords]
    offsets = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        offsets[1:] = np.cumsum([len(a) for a in arrs])
        flat = np.ascontiguousarray(np.concatenate(arrs)) if offsets[-1] else np.zeros(0, dtype=dt)
    else:
        flat = np.zeros(0, dtype=dt)
    return flat, offsets


class Oracle:
    """One dictionary + one carried scan cursor, mirroring the reference's machine/cursor pair."""

    def __init__(self, kind="ref_meyer", width=1):
        self.kind, self.width = kind, width
        self._L, self._f = _lib(kind)
        self._h = self._f("create")(width)
        if not self._h:
            raise RuntimeError("oracle create failed")

    def close(self):
        if self._h:
            self._f("release")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def insert(self, keyword):
        a = np.ascontiguousarray(np.frombuffer(keyword, dtype=_sym_dtype(self.width)) if isinstance(keyword, (bytes, bytearray)) else np.asarray(keyword, dtype=_sym_dtype(self.width)))
        return int(self._f("insert")(self._h, a.ctypes.data, len(a)))

    def insert_many(self, keywords=None, flat=None, offsets=None):
        if flat is None:
            flat, offsets = pack_keywords(keywords, self.width)
        flat = np.ascontiguousarray(flat, dtype=_sym_dtype(self.width))
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ranks = np.zeros(len(offsets) - 1, dtype=np.uint32)
        self._f("insert_many")(self._h, flat.ctypes.data, offsets.ctypes.data, len(ranks), ranks.ctypes.data)
        return ranks

    @property
    def nb_keywords(self):
        return int(self._f("nb_keywords")(self._h))

    def reset_cursor(self):
        self._f("reset_cursor")(self._h)

    def foreach_ranks(self):
        """Keyword ranks in the order the reference's acm_foreach_keyword enumerates them (reference builds only)."""
        out = np.zeros(self.nb_keywords, dtype=np.uint32)
        n = int(self._L.refh_foreach_ranks(self._h, out.ctypes.data, len(out)))
        assert n == len(out)
        return out

    def _text(self, text):
        if isinstance(text, (bytes, bytearray)):
            text = np.frombuffer(text, dtype=_sym_dtype(self.width))
        return np.ascontiguousarray(text, dtype=_sym_dtype(self.width))

    def scan(self, text, base=0, cap=None):
        """Continues from the carried cursor; returns the records in the reference's emission order."""
        t = self._text(text)
        cap = int(cap if cap is not None else min(max(1024, 64 * len(t)), 1 << 26))
        if True:
            out = np.zeros(cap, dtype=MATCH_DTYPE)
            n = int(self._f("scan")(self._h, t.ctypes.data, len(t), base, out.ctypes.data, cap, 1))
            if n <= cap:
                return out[:n]
            raise RuntimeError(f"oracle scan produced {n} > cap {cap} records; pass a larger cap")

    def count(self, text, get_match=True):
        t = self._text(text)
        return int(self._f("scan")(self._h, t.ctypes.data, len(t), 0, None, 0, 1 if get_match else 0))

    def scan_mt(self, text, nthreads, get_match=True):
        """Timing helper: thread k scans slice k from state 0. Returns (matches counted, seconds)."""
        t = self._text(text)
        sec = ctypes.c_double(0)
        n = int(self._f("scan_mt")(self._h, t.ctypes.data, len(t), nthreads, 1 if get_match else 0, ctypes.byref(sec)))
        return n, sec.value

    # port-only conveniences for big checkers
    def scan_lead(self, text, lead, base=0, cap=None):
        assert self.kind == "port"
        t = self._text(text)
        cap = int(cap if cap is not None else min(max(1024, 64 * len(t)), 1 << 26))
        out = np.zeros(cap, dtype=MATCH_DTYPE)
        cur = ctypes.c_uint32(0)
        n = int(self._L.acport_scan_lead(self._h, t.ctypes.data, len(t), lead, base, out.ctypes.data, cap, 1, ctypes.byref(cur)))
        if n > cap:
            raise RuntimeError("cap too small")
        return out[:n]

    @property
    def lmax(self):
        assert self.kind == "port"
        return int(self._L.acport_lmax(self._h))

    @property
    def nb_states(self):
        assert self.kind == "port"
        return int(self._L.acport_nb_states(self._h))


def fnv1a64_records(records):
    """FNV-1a-64 over (end, id, len) as three little-endian u64 each, in the given order."""
    a = np.empty((len(records), 3), dtype="<u8")
    a[:, 0], a[:, 1], a[:, 2] = records["end"], records["id"], records["len"]
    h = 0xCBF29CE484222325
    for b in a.tobytes():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def sort_records(records):
    """Canonical order: (end asc, len desc) == the reference's emission order (SURVEY.md 8(b) 'Result order')."""
    order = np.lexsort((-records["len"].astype(np.int64), records["end"]))
    return records[order]

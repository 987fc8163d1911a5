"""ctypes binding of libac75.so (include/aho_corasick.h + include/acm_b200.h).  Plumbing only."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MATCH_DTYPE = np.dtype([("end", "<u8"), ("id", "<u4"), ("len", "<u4")])  # ACMB200Match (field names follow the oracle's)
ENGINES = {0: "auto", 1: "dfa_smem", 2: "dfa_global", 3: "filter"}
ERRORS = {1: "invalid argument", 2: "no CUDA device (there is no CPU fallback)", 3: "CUDA error", 4: "out of memory", 5: "alphabet needs remap", 6: "capacity"}


class AcmError(RuntimeError):
    def __init__(self, code, detail=""):
        super().__init__(f"libac75 error {code} ({ERRORS.get(code, '?')}): {detail}")
        self.code = code


def library_path():
    return os.path.join(HERE, "libac75.so")


def build_library(quiet=True):
    """nvcc/gcc build of csrc/ into libac75.so (sm_100a only); cross-compiles without a GPU."""
    subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], check=True, stdout=subprocess.DEVNULL if quiet else None)
    return library_path()


class _Scan(ctypes.Structure):
    _fields_ = [("text", ctypes.c_void_p), ("nb_symbols", ctypes.c_uint64), ("lead", ctypes.c_uint64), ("base", ctypes.c_uint64),
                ("text_on_device", ctypes.c_int), ("matches_on_device", ctypes.c_int), ("matches", ctypes.c_void_p), ("capacity", ctypes.c_uint64),
                ("cursor", ctypes.POINTER(ctypes.c_void_p)), ("stream", ctypes.c_void_p)]


class Stats(ctypes.Structure):
    _fields_ = [("engine", ctypes.c_int), ("symbol_width", ctypes.c_int), ("nb_states", ctypes.c_uint32), ("nb_keywords", ctypes.c_uint32),
                ("max_keyword_length", ctypes.c_uint32), ("min_keyword_length", ctypes.c_uint32), ("nb_classes", ctypes.c_uint32),
                ("table_bytes", ctypes.c_uint64), ("smem_bytes", ctypes.c_uint64), ("finalise_count", ctypes.c_uint64), ("finalise_ms", ctypes.c_double),
                ("scan_kernel_ms", ctypes.c_double), ("main_kernel_ms", ctypes.c_double), ("h2d_ms", ctypes.c_double), ("d2h_ms", ctypes.c_double),
                ("last_nb_symbols", ctypes.c_uint64), ("last_nb_matches", ctypes.c_uint64), ("last_nb_candidates", ctypes.c_uint64),
                ("main_kernel_launches", ctypes.c_uint64), ("total_kernel_launches", ctypes.c_uint64), ("fallback_count", ctypes.c_uint64), ("filter_fp", ctypes.c_double),
                ("hot_spans", ctypes.c_uint64), ("dfa_event_scans", ctypes.c_uint64), ("filter_stride", ctypes.c_uint64),
                ("dense_scans", ctypes.c_uint64), ("patch_count", ctypes.c_uint64), ("blob_loads", ctypes.c_uint64), ("dfa_tma_scans", ctypes.c_uint64), ("dfa_lean_scans", ctypes.c_uint64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["engine"] = ENGINES.get(d["engine"], d["engine"])
        return d


_LIB = None


def lib():
    """Loads libac75.so; raises if it was not built (no silent fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: run __graft_entry__.build() / make -C csrc (the scan path has no fallback)")
    L = ctypes.CDLL(path)
    vp, u64, u32, i = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    sigs = {
        "acm_create": (vp, [vp, vp, vp]), "acm_release": (None, [vp]), "acm_initiate": (vp, [vp]), "acm_nb_keywords": (ctypes.c_size_t, [vp]),
        "acm_insert_letter_of_keyword": (None, [ctypes.POINTER(vp), vp]), "acm_insert_end_of_keyword": (vp, [ctypes.POINTER(vp), vp, vp]),
        "acm_match": (ctypes.c_size_t, [ctypes.POINTER(vp), vp]),
        "acm_b200_device_count": (i, []), "acm_b200_finalise": (i, [vp, i]), "acm_b200_scan_ex": (i, [vp, ctypes.POINTER(_Scan), ctypes.POINTER(u64)]),
        "acm_b200_insert_keywords": (i, [vp, vp, vp, u64, vp]), "acm_b200_symbol_width": (ctypes.c_size_t, [vp]), "acm_b200_max_keyword_length": (u32, [vp]),
        "acm_b200_set_option": (i, [vp, ctypes.c_char_p, ctypes.c_char_p]), "acm_b200_get_stats": (i, [vp, ctypes.POINTER(Stats)]),
        "acm_b200_last_error": (ctypes.c_char_p, []), "acm_b200_version": (ctypes.c_char_p, []),
        "acm_b200_remap_text": (i, [vp, vp, ctypes.c_size_t, u64, vp]),
        "acm_b200_keyword_order": (i, [vp, vp, u64, ctypes.POINTER(u64)]),
        "acm_b200_save": (i, [vp, ctypes.c_char_p]), "acm_b200_load": (vp, [ctypes.c_char_p, ctypes.POINTER(i)]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _LIB = L
    return L


def device_count():
    return int(lib().acm_b200_device_count())


def _check(rc, allow_capacity=False):
    if rc and not (allow_capacity and rc == 6):
        raise AcmError(rc, (lib().acm_b200_last_error() or b"").decode())


_SYM = {1: np.uint8, 2: np.uint16, 4: np.uint32}


class Machine:
    """An ACMachine created with ACM_CMP_DEFAULT over `width`-byte letters, plus one carried scan cursor."""

    def __init__(self, width=1, _handle=None):
        L = lib()
        self.width = width
        if _handle is not None:  # a machine loaded from a blob: its trie is rebuilt only when somebody needs it
            self._m = _handle
            self._cursor = ctypes.c_void_p(None)
            return
        self._size = ctypes.c_size_t(width)  # cmp_arg is borrowed for the machine's life (reference aho_corasick.c:147)
        cmp_default = ctypes.c_void_p.in_dll(L, "ACM_CMP_DEFAULT")
        self._m = L.acm_create(cmp_default, ctypes.addressof(self._size), None)
        self._cursor = ctypes.c_void_p(L.acm_initiate(self._m))

    @classmethod
    def load(cls, path):
        """acm_b200_load: a machine from a blob written by save()."""
        err = ctypes.c_int(0)
        h = lib().acm_b200_load(os.fsencode(path), ctypes.byref(err))
        if not h:
            raise AcmError(err.value, f"cannot load {path}")
        return cls(int(lib().acm_b200_symbol_width(h)), _handle=h)

    def save(self, path):
        _check(lib().acm_b200_save(self._m, os.fsencode(path)))

    def keyword_order(self):
        """Keyword ids in acm_foreach_keyword order."""
        n = ctypes.c_uint64(0)
        _check(lib().acm_b200_keyword_order(self._m, None, 0, ctypes.byref(n)))
        ids = np.zeros(n.value, dtype=np.uint32)
        _check(lib().acm_b200_keyword_order(self._m, ids.ctypes.data, len(ids), ctypes.byref(n)))
        return ids

    def close(self):
        if getattr(self, "_m", None):
            lib().acm_release(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- dictionary --
    def insert_many(self, keywords=None, flat=None, offsets=None):
        if flat is None:
            dt = _SYM[self.width]
            arrs = [np.frombuffer(k, dtype=dt) if isinstance(k, (bytes, bytearray)) else np.asarray(k, dtype=dt) for k in keywords]
            offsets = np.zeros(len(arrs) + 1, dtype=np.uint64)
            offsets[1:] = np.cumsum([len(a) for a in arrs]) if arrs else []
            flat = np.concatenate(arrs) if arrs and offsets[-1] else np.zeros(0, dtype=dt)
        flat = np.ascontiguousarray(flat, dtype=_SYM[self.width])
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = np.zeros(len(offsets) - 1, dtype=np.uint32)
        _check(lib().acm_b200_insert_keywords(self._m, flat.ctypes.data, offsets.ctypes.data, len(ids), ids.ctypes.data))
        return ids

    @property
    def nb_keywords(self):
        return int(lib().acm_nb_keywords(self._m))

    @property
    def max_keyword_length(self):
        return int(lib().acm_b200_max_keyword_length(self._m))

    def set_option(self, key, value):
        _check(lib().acm_b200_set_option(self._m, key.encode(), str(value).encode()))

    def finalise(self, device=-1):
        _check(lib().acm_b200_finalise(self._m, device))

    def stats(self):
        s = Stats()
        _check(lib().acm_b200_get_stats(self._m, ctypes.byref(s)))
        return s.as_dict()

    def stats_into(self, s=None):
        """Fills a Stats structure in place (a new one if none is given) and returns it -- no dictionary: for timed loops, where host
        time between two scans is GPU idle time."""
        if s is None:
            s = Stats()
        _check(lib().acm_b200_get_stats(self._m, ctypes.byref(s)))
        return s

    def reset_cursor(self):
        self._cursor = ctypes.c_void_p(lib().acm_initiate(self._m))

    # -- host-side per-symbol API (the reference's loop), for tests --
    def host_match_count(self, text):
        L = lib()
        t = np.ascontiguousarray(text if not isinstance(text, (bytes, bytearray)) else np.frombuffer(text, dtype=_SYM[self.width]), dtype=_SYM[self.width])
        cur = ctypes.c_void_p(L.acm_initiate(self._m))
        total, base = 0, t.ctypes.data
        for k in range(len(t)):
            total += L.acm_match(ctypes.byref(cur), base + k * self.width)
        return total

    # -- batch scans --
    def scan(self, text, lead=0, base=0, capacity=None, carry=False, count_only=False):
        """Host text (bytes / numpy) -> records in the reference's emission order (numpy structured array)."""
        t = np.ascontiguousarray(np.frombuffer(text, dtype=_SYM[self.width]) if isinstance(text, (bytes, bytearray)) else np.asarray(text, dtype=_SYM[self.width]))
        cap = 0 if count_only else int(capacity if capacity is not None else min(max(1024, 64 * len(t)), 1 << 26))
        out = np.zeros(cap, dtype=MATCH_DTYPE)
        s = _Scan(text=t.ctypes.data if len(t) else None, nb_symbols=len(t), lead=lead, base=base, text_on_device=0, matches_on_device=0,
                  matches=out.ctypes.data if cap else None, capacity=cap, cursor=ctypes.pointer(self._cursor) if carry else None, stream=None)
        n = ctypes.c_uint64(0)
        _check(lib().acm_b200_scan_ex(self._m, ctypes.byref(s), ctypes.byref(n)), allow_capacity=True)
        if count_only:
            return int(n.value)
        if n.value > cap:
            raise AcmError(6, f"{n.value} records > capacity {cap}")
        return out[: n.value]

    def scan_device(self, d_text_ptr, nb_symbols, lead=0, base=0, d_matches_ptr=None, capacity=0, stream=None):
        """Device-resident text (raw pointer, 16-byte aligned); records stay on the device. Returns the total number found."""
        s = _Scan(text=d_text_ptr, nb_symbols=nb_symbols, lead=lead, base=base, text_on_device=1, matches_on_device=1, matches=d_matches_ptr, capacity=capacity,
                  cursor=None, stream=stream)
        n = ctypes.c_uint64(0)
        _check(lib().acm_b200_scan_ex(self._m, ctypes.byref(s), ctypes.byref(n)), allow_capacity=True)
        return int(n.value)

    def scan_host_to_host(self, host_ptr, nb_symbols, out_ptr, capacity, lead=0, base=0):
        """Raw host pointers (e.g. pinned torch tensors) in and out: the end-to-end path. Returns the total number found."""
        s = _Scan(text=host_ptr, nb_symbols=nb_symbols, lead=lead, base=base, text_on_device=0, matches_on_device=0, matches=out_ptr, capacity=capacity,
                  cursor=None, stream=None)
        n = ctypes.c_uint64(0)
        _check(lib().acm_b200_scan_ex(self._m, ctypes.byref(s), ctypes.byref(n)), allow_capacity=True)
        return int(n.value)

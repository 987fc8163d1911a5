"""a finalised dictionary on disk.  save -> load gives a machine with the same keyword ids whose batch scan
uploads the stored tables as they are (no insertion, no build) and returns identical records; the per-symbol API, carried
cursors and further insertions work on it too (the trie is rebuilt from the packed dictionary on demand)."""
import numpy as np
import pytest

from helpers import ac75, generate_text, oracle_records, pack, random_patterns
from oracle import pyoracle

CASES = [  # (width, dictionary)
    (1, lambda: random_patterns(3000, seed=5)),                            # filter engine, stride-2 tables
    (1, lambda: pack([b"he", b"she", b"his", b"hers", b"ushers"])),       # shared-memory DFA
    (1, lambda: random_patterns(4000, lmin=2, lmax=12, seed=6)),           # global-memory DFA (short keywords)
    (2, lambda: random_patterns(2000, lmin=2, lmax=9, seed=7, alphabet=700, width=2)),
    (4, lambda: random_patterns(2000, lmin=2, lmax=8, seed=8, alphabet=50_000, width=4)),
]


@pytest.mark.parametrize("width,make", CASES)
def test_blob_round_trip_on_the_host(tmp_path, width, make):
    flat, offsets = make()
    m = ac75().Machine(width)
    ids = m.insert_many(flat=flat, offsets=offsets)
    path = tmp_path / "dict.ac75"
    m.save(path)
    l = ac75().Machine.load(path)
    assert l.width == width and l.nb_keywords == m.nb_keywords and l.max_keyword_length == m.max_keyword_length  # no trie needed for these
    assert np.array_equal(l.keyword_order(), m.keyword_order())  # (rebuilds the trie of the loaded machine)
    rng = np.random.default_rng(width)
    text = rng.integers(0, 256 if width == 1 else 700, size=20000).astype({1: np.uint8, 2: np.uint16, 4: np.uint32}[width])
    k = flat[int(offsets[1]):int(offsets[2])]
    text[100:100 + len(k)] = k
    assert l.host_match_count(text) == m.host_match_count(text) > 0
    assert np.array_equal(l.insert_many(flat=flat, offsets=offsets), ids)  # every keyword kept its id
    m.close(), l.close()


def test_load_rejects_what_is_not_a_blob(tmp_path):
    bad = tmp_path / "bad.ac75"
    bad.write_bytes(b"not a blob at all" * 10)
    with pytest.raises(ac75().AcmError):
        ac75().Machine.load(bad)
    with pytest.raises(ac75().AcmError):
        ac75().Machine.load(tmp_path / "missing.ac75")
    # a blob of another format version (its tables may be hashed differently: version 3 changed the slot hash) is refused, not misread
    m = ac75().Machine(1)
    m.insert_many([b"abcde", b"bcdef", b"xyz12"])
    good = tmp_path / "good.ac75"
    m.save(good)
    m.close()
    ac75().Machine.load(good).close()
    # equal dictionaries give equal files (no pointer of the saving process is written)
    m2 = ac75().Machine(1)
    m2.insert_many([b"abcde", b"bcdef", b"xyz12"])
    again = tmp_path / "again.ac75"
    m2.save(again)
    m2.close()
    assert again.read_bytes() == good.read_bytes()
    raw = bytearray(good.read_bytes())
    version = int.from_bytes(raw[8:12], "little")
    raw[8:12] = (version - 1).to_bytes(4, "little")
    old = tmp_path / "old.ac75"
    old.write_bytes(bytes(raw))
    with pytest.raises(ac75().AcmError):
        ac75().Machine.load(old)


@pytest.mark.gpu
@pytest.mark.parametrize("width,make", CASES)
def test_blob_scan_equals_original_scan(tmp_path, width, make):
    flat, offsets = make()
    m = ac75().Machine(width)
    m.insert_many(flat=flat, offsets=offsets)
    if width == 1:
        text = generate_text(2 << 20, kind=0, plant_period=512, dict_flat=flat, dict_offsets=offsets)
        if m.nb_keywords < 10:
            text = np.frombuffer(b"To ushers: he found his pencil, but she could not find hers. " * 20000, dtype=np.uint8)
    else:
        rng = np.random.default_rng(9)
        text = rng.integers(0, 700 if width == 2 else 50_000, size=1 << 20).astype({2: np.uint16, 4: np.uint32}[width])
        for p in range(0, len(text) - 64, 997):
            k = int(rng.integers(0, len(offsets) - 1))
            kw = flat[int(offsets[k]):int(offsets[k + 1])]
            text[p:p + len(kw)] = kw
    want = m.scan(text)
    engine = m.stats()["engine"]
    path = tmp_path / "dict.ac75"
    m.save(path)
    l = ac75().Machine.load(path)
    got = l.scan(text)
    st = l.stats()
    assert st["engine"] == engine and st["blob_loads"] == 1 and st["finalise_count"] == 1
    assert len(want) > 100 and np.array_equal(got, want)
    assert np.array_equal(got, oracle_records(flat=flat, offsets=offsets, text=text, width=width, kind="port"))
    # carried cursor across two scans of the loaded machine == one scan
    l.reset_cursor()
    half = len(text) // 2
    a = l.scan(text[:half], carry=True)
    b = l.scan(text[half:], base=half, carry=True)
    assert np.array_equal(np.concatenate([a, b]), want)
    # an insertion after loading: tables follow, ids continue
    extra = np.asarray(text[777:777 + 5])
    new_id = l.insert_many([extra.tobytes() if width == 1 else extra])
    o = pyoracle.Oracle("port", width)
    o.insert_many(flat=flat, offsets=offsets)
    assert int(o.insert_many([extra.tobytes() if width == 1 else extra])[0]) == int(new_id[0])
    assert np.array_equal(l.scan(text), o.scan(text, cap=max(1 << 16, 8 * len(text))))
    m.close(), l.close(), o.close()

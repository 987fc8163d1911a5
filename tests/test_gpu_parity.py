"""GPU parity tests (B200): every record the CUDA path returns through the C-ABI == the oracle's, bit for bit and in the
reference's emission order, on the BASELINE.json configs at test scale, for every engine, with carried cursors, leads and shards."""
import numpy as np
import pytest

from helpers import ac75, best_oracle_kind, config1_keywords, config2_keywords, generate_text, oracle_records, pack, random_patterns
from oracle import pyoracle

pytestmark = pytest.mark.gpu


def gpu_scan(keywords=None, flat=None, offsets=None, text=None, width=1, engine="auto", **opts):
    m = ac75().Machine(width)
    m.insert_many(keywords, flat=flat, offsets=offsets)
    if engine != "auto":
        m.set_option("engine", engine)
    for k, v in opts.items():
        m.set_option(k, v)
    r = m.scan(text, capacity=max(1 << 16, 8 * len(text)))
    st = m.stats()
    m.close()
    return r, st


def test_readme_example_all_engines():
    words = [b"he", b"she", b"his", b"hers"]
    text = b"To ushers: he found his pencil, but she could not find hers."
    want = oracle_records(words, text=text)
    assert len(want) == 9
    for engine in ("dfa_smem", "dfa_global", "filter"):
        got, st = gpu_scan(words, text=text, engine=engine)
        assert st["engine"] == engine
        assert np.array_equal(got, want), engine


@pytest.mark.parametrize("full", [False, True])
@pytest.mark.parametrize("engine", ["auto", "dfa_global", "filter"])
def test_config1_novel(novel, golden_config1, full, engine):
    g = golden_config1["readme_plus_wordlist" if full else "readme_only"]
    got, st = gpu_scan(config1_keywords(novel, full), text=novel, engine=engine)
    assert len(got) == g["matches"]
    assert "%016x" % pyoracle.fnv1a64_records(got) == g["fnv1a64"], st  # order-sensitive hash == the reference's emission order
    assert [list(map(int, x)) for x in got[:48].tolist()] == g["head"]


@pytest.mark.parametrize("engine", ["auto", "dfa_global", "filter"])
def test_config2_words_over_ascii(novel, engine):
    words = config2_keywords(novel)
    flat, offsets = pack(words)
    text = generate_text(4 << 20, kind=1, plant_period=4096, dict_flat=flat, dict_offsets=offsets)
    want = oracle_records(words, text=text, kind="port")
    got, st = gpu_scan(words, text=text, engine=engine)
    if engine == "auto":
        assert st["engine"] == "dfa_smem"
    assert len(want) > 1000 and np.array_equal(got, want), (engine, len(got), len(want))


@pytest.mark.parametrize("engine", ["auto", "dfa_global"])
def test_config3_random_patterns(engine):
    flat, offsets = random_patterns(5000)
    text = generate_text(8 << 20, kind=0, plant_period=4096, dict_flat=flat, dict_offsets=offsets)
    want = oracle_records(flat=flat, offsets=offsets, text=text, kind="port")
    got, st = gpu_scan(flat=flat, offsets=offsets, text=text, engine=engine)
    if engine == "auto":
        assert st["engine"] == "filter" and st["fallback_count"] == 0 and st["filter_stride"] == 2
    assert len(want) >= 2000 and np.array_equal(got, want), (engine, len(got), len(want))


@pytest.mark.parametrize("width,alphabet,lmin", [(2, 300, 2), (4, 50000, 2), (4, 3, 1), (2, 2, 1), (1, 2, 1), (1, 4, 3)])
def test_generic_alphabets_and_dense_matches(width, alphabet, lmin):
    rng = np.random.default_rng(width * 1000 + alphabet)
    flat, offsets = random_patterns(300, lmin=lmin, lmax=8, seed=alphabet, alphabet=alphabet, width=width)
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[width]
    text = rng.integers(0, alphabet, size=200_000).astype(dt)
    # plant some keywords so that sparse alphabets have matches too
    for k in rng.integers(0, 300, size=500):
        kw = flat[int(offsets[k]):int(offsets[k + 1])]
        at = int(rng.integers(0, len(text) - len(kw)))
        text[at:at + len(kw)] = kw
    want = oracle_records(flat=flat, offsets=offsets, text=text, width=width, kind="port")
    engines = ["auto", "filter"] + (["dfa_global"] if width == 1 else [])
    for engine in engines:
        got, st = gpu_scan(flat=flat, offsets=offsets, text=text, width=width, engine=engine)
        assert len(want) > 0 and np.array_equal(got, want), (engine, st, len(got), len(want))


@pytest.mark.parametrize("engine", ["dfa_smem", "dfa_global", "filter"])
def test_carried_cursor_kats(golden_kats, engine):
    """Insertions interleaved with scanning on one carried cursor (Meyer): each insertion triggers a rebuild + re-upload."""
    for kat in golden_kats:
        m = ac75().Machine(kat["width"])
        m.set_option("engine", engine)
        for step, want in zip(kat["steps"], kat["results"]):
            if step[0] == "insert":
                assert m.insert_many([k.encode("latin1") for k in step[1]]).tolist() == want["ranks"], kat["name"]
            elif step[0] == "scan":
                got = [list(map(int, x)) for x in m.scan(step[1].encode("latin1"), carry=True).tolist()]
                assert got == want["records"], (kat["name"], engine)
            else:
                m.reset_cursor()
        m.close()


@pytest.mark.parametrize("engine", ["dfa_global", "filter"])
def test_incremental_rounds_random(engine):
    """config-5 shape at test scale: rounds of insertions, each followed by a scan of the next slice with the cursor carried."""
    rng = np.random.default_rng(5)
    o = pyoracle.Oracle(best_oracle_kind(), 1)
    m = ac75().Machine(1)
    m.set_option("engine", engine)
    stream = rng.integers(97, 101, size=60_000).astype(np.uint8)
    for rnd in range(6):
        kws = [stream[a:a + l].tobytes() for a, l in zip(rng.integers(0, 59_000, size=200), rng.integers(2, 9, size=200))]
        assert np.array_equal(o.insert_many(kws), m.insert_many(kws))
        sl = stream[rnd * 10_000:(rnd + 1) * 10_000]
        want = o.scan(sl, base=rnd * 10_000, cap=1 << 22)
        got = m.scan(sl, base=rnd * 10_000, carry=True, capacity=1 << 22)
        assert np.array_equal(got, want), (engine, rnd, len(got), len(want))
    assert m.stats()["finalise_count"] == 6
    m.close()
    o.close()


@pytest.mark.parametrize("engine", ["auto", "dfa_global"])
def test_lead_base_and_shards(engine):
    flat, offsets = random_patterns(2000, lmin=4, lmax=32, seed=99)
    text = generate_text(1 << 20, kind=0, plant_period=1024, dict_flat=flat, dict_offsets=offsets)
    want = oracle_records(flat=flat, offsets=offsets, text=text, kind="port")
    m = ac75().Machine(1)
    m.insert_many(flat=flat, offsets=offsets)
    m.set_option("engine", engine)
    lmax = m.max_keyword_length
    for world in (1, 2, 3, 8):
        parts = []
        for (a, b, lead) in ac75().plan_shards(len(text), world, lmax):
            parts.append(m.scan(text[a - lead:b], lead=lead, base=a - lead))
        got = np.concatenate(parts)
        assert np.array_equal(got, want), (engine, world)
    m.close()


def test_empty_and_tiny_inputs():
    for engine in ("dfa_smem", "dfa_global", "filter"):
        m = ac75().Machine(1)
        m.set_option("engine", engine)
        assert len(m.scan(b"anything at all")) == 0  # empty dictionary (reference generic test :70)
        m.insert_many([b"abcd"])
        assert len(m.scan(b"")) == 0
        assert len(m.scan(b"abc")) == 0
        assert m.scan(b"abcd").tolist() == [(3, 0, 4)]
        assert m.scan(b"xabcdabcd").tolist() == [(4, 0, 4), (8, 0, 4)]
        m.close()


def test_capacity_is_reported():
    m = ac75().Machine(1)
    m.insert_many([b"a"])
    assert m.scan(b"a" * 1000, count_only=True) == 1000
    with pytest.raises(ac75().AcmError):
        m.scan(b"a" * 1000, capacity=10)
    m.close()


def test_custom_comparator_machine_through_c_abi(tmp_path, novel):
    """wchar_t letters + case-insensitive user comparator (the reference's own test setup): class-id remap + batch scan must equal
    the per-symbol loop of the same machine, across insertions on a carried cursor.  Driven from C, as a user of the C-ABI would."""
    import os
    import subprocess

    from conftest import ROOT

    libdir = os.path.join(ROOT, "aho-corasick-1975_b200")
    exe = tmp_path / "custom_cmp_parity"
    subprocess.run(["gcc", "-std=c11", "-O2", "-D_XOPEN_SOURCE=700", f"-I{os.path.join(ROOT, 'include')}", "-o", str(exe), os.path.join(ROOT, "tests", "csrc", "custom_cmp_parity.c"),
                    f"-L{libdir}", "-lac75", f"-Wl,-rpath,{libdir}"], check=True)
    txt = tmp_path / "novel.txt"
    txt.write_bytes(bytes(b if b < 128 else 32 for b in novel))
    r = subprocess.run([str(exe), str(txt)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("records identical") == 4 and "width 4" in r.stdout
    # the same workload through the ORACLE (case folded by hand: under the comparator "The" and "the" are one keyword and one text):
    # keyword ids, record counts and the FNV-1a-64 of the records (SURVEY Appendix B's hash) must be what the GPU printed
    import re

    raw = txt.read_bytes()[: 1 << 20]
    low = np.frombuffer(raw.lower(), dtype=np.uint8)
    o = pyoracle.Oracle(best_oracle_kind(), 1)
    alpha = np.array([chr(c).isascii() and chr(c).isalpha() for c in range(256)])[np.frombuffer(raw, dtype=np.uint8)]
    rounds, per, nk, n = 4, len(raw) // 4, 0, len(raw)
    printed = re.findall(r"round (\d): (\d+) keywords, (\d+) records identical, gpu fnv ([0-9a-f]{16})", r.stdout)
    assert len(printed) == rounds
    for rnd in range(rounds):
        i = rnd * per
        while i < (rnd + 1) * per and nk < 4000:
            if alpha[i]:
                j = i
                while j < n and alpha[j]:
                    j += 1
                before = o.nb_keywords
                o.insert(bytes(low[i:j]))
                nk += o.nb_keywords > before
                i = j + 37
            else:
                i += 1
        want = o.scan(low[rnd * per:(rnd + 1) * per], base=rnd * per, cap=2_000_000)
        assert (int(printed[rnd][1]), int(printed[rnd][2])) == (o.nb_keywords, len(want)), (rnd, printed[rnd], o.nb_keywords, len(want))
        assert int(printed[rnd][3], 16) == pyoracle.fnv1a64_records(want), rnd
    o.close()


@pytest.mark.parametrize("engine", ["dfa_smem", "dfa_global", "filter"])
def test_streaming_ingest_equals_single_run(engine):
    """Host text larger than the streaming threshold goes through two bounded device buffers in chunks (each re-copied with
    max depth - 1 symbols of left context); records, their order, leads and the carried cursor must equal a single run's."""
    words = [b"he", b"she", b"his", b"hers", b"ushers", b"pencil", b"abcdefghijklmnopqrstuvwxyz"]
    rng = np.random.default_rng(8)
    pieces = [b"To ushers: he found his pencil, but she could not find hers. ", b"abcdefghijklmnopqrstuvwxyz", b"xx", b"shershe"]
    text = b"".join(pieces[i] for i in rng.integers(0, len(pieces), size=60_000))
    text = np.frombuffer(text, dtype=np.uint8)
    o = pyoracle.Oracle(best_oracle_kind(), 1)
    o.insert_many(words)
    want1 = o.scan(text[:1000], cap=1 << 22)
    want2 = o.scan(text[1000:], base=1000, cap=1 << 23)  # cursor carried
    m = ac75().Machine(1)
    m.insert_many(words)
    m.set_option("engine", engine)
    m.set_option("stream_bytes", 256 * 1024)  # ~2.3 MB of text -> 9 chunks
    got1 = m.scan(text[:1000], carry=True, capacity=1 << 22)
    got2 = m.scan(text[1000:], base=1000, carry=True, capacity=1 << 23)
    assert np.array_equal(got1, want1) and len(want2) > 100_000 and np.array_equal(got2, want2), (engine, len(got2), len(want2))
    # a lead that spans several chunks, as a shard with a long left context would have
    lead = 700_000
    got3 = m.scan(text, lead=lead, capacity=1 << 23)
    o.reset_cursor()
    want3 = o.scan(text, cap=1 << 23)
    assert np.array_equal(got3, want3[want3["end"] >= lead])
    m.close(), o.close()


def test_batch_api_error_codes():
    """The batch ABI reports misuse through return codes (the legacy API keeps the reference's ACM_ASSERT convention)."""
    import ctypes

    import torch

    pkg = ac75()
    a, b = pkg.Machine(1), pkg.Machine(1)
    a.insert_many([b"abc"]), b.insert_many([b"abc"])
    # capacity too small: ERR_CAPACITY, the total is still reported and the first records are valid
    out = np.zeros(2, dtype=pkg.MATCH_DTYPE)
    text = np.frombuffer(b"abc" * 10, dtype=np.uint8)
    from importlib import import_module

    binding = import_module("ac75.binding")
    s = binding._Scan(text=text.ctypes.data, nb_symbols=len(text), lead=0, base=0, text_on_device=0, matches_on_device=0, matches=out.ctypes.data, capacity=2,
                      cursor=None, stream=None, sorted=1)
    n = ctypes.c_uint64(0)
    assert pkg.lib().acm_b200_scan_ex(a._m, ctypes.byref(s), ctypes.byref(n)) == 6 and n.value == 10
    assert out.tolist() == [(2, 0, 3), (5, 0, 3)]
    # lead larger than the text, null text with symbols: ERR_INVALID
    s.lead = len(text) + 1
    assert pkg.lib().acm_b200_scan_ex(a._m, ctypes.byref(s), ctypes.byref(n)) == 1
    s.lead, s.text = 0, None
    assert pkg.lib().acm_b200_scan_ex(a._m, ctypes.byref(s), ctypes.byref(n)) == 1
    # a cursor of another machine: ERR_INVALID
    s.text = text.ctypes.data
    foreign = ctypes.c_void_p(pkg.lib().acm_initiate(b._m))
    s.cursor = ctypes.pointer(foreign)
    assert pkg.lib().acm_b200_scan_ex(a._m, ctypes.byref(s), ctypes.byref(n)) == 1
    # misaligned device text: ERR_INVALID
    d = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    with pytest.raises(pkg.AcmError) as e:
        a.scan_device(d.data_ptr() + 1, 100)
    assert e.value.code == 1
    assert a.scan_device(d.data_ptr(), 1024) == 0
    a.close(), b.close()


@pytest.mark.parametrize("engine", ["filter", "dfa_global", "auto"])
def test_all_ones_and_zero_bytes(engine):
    """Keys made of 0xFF / 0x00 bytes: 0xFFFFFFFF is the empty marker of the filter's compact key set and must still be found."""
    words = [b"\xff\xff\xff\xff", b"\xff\xff\xff\xff\xff\xff", b"\x00\x00\x00\x00", b"\x00\xff\x00\xff", b"\xff\x00\x00\x00\x00\xff"]
    rng = np.random.default_rng(3)
    text = rng.choice(np.array([0, 255], dtype=np.uint8), size=300_000, p=[0.5, 0.5])
    text[1000:1100] = 255
    text[5000:5100] = 0
    want = oracle_records(words, text=text)
    got, st = gpu_scan(words, text=text, engine=engine)
    assert len(want) > 10_000 and np.array_equal(got, want), (engine, st["engine"], len(got), len(want))


def test_stride2_kernel_tile_and_span_boundaries():
    """Stride-2 filter kernel (byte alphabet, shortest keyword >= 4): keywords of both end parities and every length class planted
    across the 2 KiB tile, 32 KiB span and text boundaries; with the kernel switched off the records must be the same."""
    rng = np.random.default_rng(77)
    kws = [rng.integers(0, 256, size=int(L), dtype=np.uint8).tobytes() for L in list(range(4, 12)) * 40 + [32, 31, 17] * 10]
    kws += [bytes(4), b"\x07" + bytes(5)]  # all-zero windows are keys: the zero padding of the last, partial tile must not count as hits
    n = 3 * 32768 + 2048 + 777
    o = pyoracle.Oracle("port", 1)
    o.insert_many(kws)
    m = ac75().Machine(1)
    m.insert_many(kws)
    m.set_option("engine", "filter")
    m1 = ac75().Machine(1)
    m1.insert_many(kws)
    m1.set_option("engine", "filter")
    m1.set_option("stride2", 0)
    total = 0
    for delta in range(-8, 9):
        text = rng.integers(0, 256, size=n, dtype=np.uint8)
        for i, b in enumerate([4, 16, 512, 2048, 4096, 32768, 32768 + 2048, 65536, 3 * 32768, n]):
            k = np.frombuffer(kws[(i * 17 + delta + 8) % len(kws)], dtype=np.uint8)
            end = min(n, max(len(k), b + delta))  # one past the last byte of the planted keyword
            text[end - len(k):end] = k
        for lead in (0, 5, 2049):
            want = o.scan(text, cap=1 << 20)
            o.reset_cursor()
            want = want[want["end"] >= lead]
            got = m.scan(text, lead=lead, capacity=1 << 20)
            got1 = m1.scan(text, lead=lead, capacity=1 << 20)
            st, st1 = m.stats(), m1.stats()
            assert st["filter_stride"] == 2 and st1["filter_stride"] == 1 and st["fallback_count"] == 0
            assert np.array_equal(got, want), (delta, lead, len(got), len(want))
            assert np.array_equal(got1, want), (delta, lead, len(got1), len(want))
            total += len(want)
    assert total > 300
    m.close(), m1.close(), o.close()


def test_stride2_kernel_hot_spans_are_redone_exactly():
    """A run of equal bytes whose 3-byte window is a filter key makes every test of several tiles hit; the stage holds ~100 hits.
    Such spans are handed to the exact follow-up kernel (no whole-scan fallback) and the records still equal the oracle's."""
    rng = np.random.default_rng(5)
    kws = [rng.integers(0, 256, size=int(L), dtype=np.uint8).tobytes() for L in list(range(4, 20)) * 30]
    kws += [b"\x01\x09\x09\x09", b"\x09\x09\x09\x02\x03"]  # (9,9,9) is a key of both roles; a run of 9s matches neither keyword
    text = rng.integers(0, 256, size=200_000, dtype=np.uint8)
    text[50_000:57_000] = 9
    text[57_000:57_002] = (2, 3)  # ... until the run ends in the second keyword
    text[120_001:123_500] = 9
    want = oracle_records(kws, text=text, kind="port")
    m = ac75().Machine(1)
    m.insert_many(kws)
    m.set_option("engine", "filter")
    got = m.scan(text, capacity=1 << 16)
    st = m.stats()
    assert st["filter_stride"] == 2 and st["fallback_count"] == 0 and 2 <= st["hot_spans"] <= 4, st
    assert len(want) >= 1 and np.array_equal(got, want), (len(got), len(want))
    # a span with more true candidates than the warp's candidate buffer holds (a keyword planted every 64 bytes) takes the same route
    text2 = rng.integers(0, 256, size=200_000, dtype=np.uint8)
    k = np.frombuffer(kws[3], dtype=np.uint8)
    for at in range(70_000, 80_000, 64):
        text2[at:at + len(k)] = k
    want2 = oracle_records(kws, text=text2, kind="port")
    got2 = m.scan(text2, capacity=1 << 16)
    assert m.stats()["fallback_count"] == 0 and m.stats()["hot_spans"] > st["hot_spans"]
    assert len(want2) >= 150 and np.array_equal(got2, want2), (len(got2), len(want2))
    m.close()


def test_dfa_second_pass_from_recorded_events(novel):
    """Shared-memory DFA engine: pass 1 records where it met an output state and pass 2 expands those events; with the recording
    switched off pass 2 walks the text again.  Both must give the oracle's records; a text with more events than slots (one event
    slot per 4 symbols) must fall back to the walk by itself."""
    words = config2_keywords(novel, 300)
    flat, offsets = pack(words)
    text = generate_text(3 << 20, kind=1, plant_period=512, dict_flat=flat, dict_offsets=offsets)
    want = oracle_records(words, text=text, kind="port")
    for events, lean in ((1, 0), (1, 1), (0, 0)):
        m = ac75().Machine(1)
        m.insert_many(words)
        m.set_option("engine", "dfa_smem")
        m.set_option("dfa_events", events)
        m.set_option("dfa_lean", lean)  # 1: pass 1 records events only, the records are counted from the event lists afterwards
        got = m.scan(text, capacity=1 << 22)
        lead_got = m.scan(text, lead=100_001, base=7, capacity=1 << 22)
        st = m.stats()
        assert st["engine"] == "dfa_smem" and st["dfa_event_scans"] == (2 if events else 0) and (st["dfa_lean_scans"] > 0) == bool(lean), st
        assert len(want) > 100_000 and np.array_equal(got, want)
        w2 = want[want["end"] >= 100_001].copy()
        w2["end"] += 7
        assert np.array_equal(lead_got, w2)
        m.close()
    # every symbol ends a keyword: more events than slots, the second pass walks
    m = ac75().Machine(1)
    m.insert_many([b"a", b"aa", b"ab"])
    m.set_option("engine", "dfa_smem")
    dense = np.frombuffer(b"aab" * 50_000, dtype=np.uint8)
    want = oracle_records([b"a", b"aa", b"ab"], text=dense, kind="port")
    got = m.scan(dense, capacity=1 << 20)
    assert m.stats()["dfa_event_scans"] == 0 and np.array_equal(got, want) and len(want) > 150_000
    m.close()

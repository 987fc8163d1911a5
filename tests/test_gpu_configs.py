"""GPU parity at the shapes of BASELINE.json configs[3], configs[4] and at configs[2]'s full size.

Sizes the oracle finishes in seconds are compared record for record; the full-size run is checked through size-independent
properties (exact prefix, recall of every sampled planted keyword, shards summing to the whole, engines agreeing)."""
import ctypes

import numpy as np
import pytest

from helpers import ac75, generate_text, random_patterns
from oracle import pyoracle

pytestmark = pytest.mark.gpu


def zipf_tokens(n, vocab, seed):
    """Zipf(1.0) over `vocab` ids via the inverse CDF."""
    rng = np.random.default_rng(seed)
    cdf = np.cumsum(1.0 / np.arange(1, vocab + 1))
    cdf /= cdf[-1]
    return np.searchsorted(cdf, rng.random(n)).astype(np.uint32)


def test_config5_token_ngrams_incremental_meyer():
    """uint32 alphabet, 50k-symbol vocabulary, n-gram keywords cut from the stream (so they occur, densely for frequent tokens),
    then rounds of insertions each followed by a scan of the next slice with the cursor carried: every round is one DFA/filter
    rebuild + re-upload, and the result must equal the reference-semantics oracle's carried-cursor scan."""
    vocab, n, rounds = 50_000, 1_200_000, 6
    stream = zipf_tokens(n, vocab, 42)
    rng = np.random.default_rng(43)

    def cut(k):
        starts, lens = rng.integers(0, n - 8, size=k), rng.integers(2, 9, size=k)
        flat = np.concatenate([stream[a:a + l] for a, l in zip(starts, lens)])
        return flat, np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)

    o = pyoracle.Oracle("port", 4)
    m = ac75().Machine(4)
    flat, offsets = cut(20_000)
    assert np.array_equal(o.insert_many(flat=flat, offsets=offsets), m.insert_many(flat=flat, offsets=offsets))
    per = n // (rounds + 1)
    total = 0
    for r in range(rounds + 1):
        if r:
            flat, offsets = cut(200)  # Meyer phase: a few insertions between two scans
            assert np.array_equal(o.insert_many(flat=flat, offsets=offsets), m.insert_many(flat=flat, offsets=offsets))
        sl = stream[r * per:(r + 1) * per]
        want = o.scan(sl, base=r * per, cap=1 << 24)
        got = m.scan(sl, base=r * per, carry=True, capacity=1 << 24)
        assert len(want) > 1000 and np.array_equal(got, want), (r, len(got), len(want))
        total += len(got)
    st = m.stats()
    assert st["engine"] == "filter" and st["symbol_width"] == 4 and st["finalise_count"] == rounds + 1
    m.close(), o.close()


def test_config4_shape_large_dictionary():
    """configs[3] shape at test scale: a dictionary whose tables exceed shared memory by far (300k patterns, ~5M states)."""
    flat, offsets = random_patterns(300_000, seed=4)
    text = generate_text(16 << 20, kind=0, plant_period=2048, dict_flat=flat, dict_offsets=offsets)
    o = pyoracle.Oracle("port", 1)
    o.insert_many(flat=flat, offsets=offsets)
    want = o.scan(text, cap=1 << 22)
    m = ac75().Machine(1)
    m.insert_many(flat=flat, offsets=offsets)
    got = m.scan(text, capacity=1 << 22)
    st = m.stats()
    assert st["engine"] == "filter" and st["nb_states"] > 4_000_000 and st["fallback_count"] == 0
    assert len(want) >= 8000 and np.array_equal(got, want)
    m.close(), o.close()


@pytest.mark.slow
def test_config3_full_size_properties():
    """configs[2] at its full size (100k patterns, 8 GiB on one GPU), device-resident."""
    import torch

    flat, offsets = random_patterns(100_000, seed=0xD1C7)
    n = 8 << 30
    m = ac75().Machine(1)
    ids = m.insert_many(flat=flat, offsets=offsets)
    d_text = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    generate_text(n, kind=0, plant_period=4096, dict_flat=flat, dict_offsets=offsets, device_ptr=d_text.data_ptr())
    cap = 1 << 22
    d_out = torch.empty(cap * 16, dtype=torch.uint8, device="cuda")
    total = m.scan_device(d_text.data_ptr(), n, d_matches_ptr=d_out.data_ptr(), capacity=cap)
    assert m.stats()["engine"] == "filter" and m.stats()["fallback_count"] == 0 and total <= cap
    recs = d_out[: total * 16].cpu().numpy().view(ac75().MATCH_DTYPE)
    # (1) the reference's emission order: end ascending, longest first
    assert np.array_equal(pyoracle.sort_records(recs), recs)
    # (2) exact equality with the oracle on a 32 MiB prefix
    pre = 32 << 20
    host_prefix = d_text[:pre].cpu().numpy()
    o = pyoracle.Oracle("port", 1)
    o.insert_many(flat=flat, offsets=offsets)
    want = o.scan(host_prefix, cap=1 << 20)
    got_prefix = recs[recs["end"] < pre]
    assert len(want) > 8000 and np.array_equal(got_prefix, want)
    # (3) recall: the keyword planted in sampled 4 KiB periods is reported at its position (generator of acm_kernels.cuh::gen_byte)
    def mix64(x):
        x = np.uint64(x)
        x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))
    rng = np.random.default_rng(1)
    ends = recs["end"]
    checked = 0
    with np.errstate(over="ignore"):
        for b in rng.integers(1, (n // 4096) - 1, size=4000):
            r = mix64(np.uint64(0x5EED) + np.uint64(b) * np.uint64(0x9E3779B97F4A7C15))
            kw, at = int((r >> np.uint64(20)) % np.uint64(len(ids))), int((r & np.uint64(0xFFFFF)) % np.uint64(4096))
            length = int(offsets[kw + 1] - offsets[kw])
            r2 = mix64(np.uint64(0x5EED) + np.uint64(b + 1) * np.uint64(0x9E3779B97F4A7C15))
            at2 = int((r2 & np.uint64(0xFFFFF)) % np.uint64(4096))
            if at + length > 4096 + at2:  # overwritten by the next period's plant
                continue
            end = int(b) * 4096 + at + length - 1
            lo, hi = np.searchsorted(ends, end, "left"), np.searchsorted(ends, end, "right")
            assert any(int(x["id"]) == int(ids[kw]) and int(x["len"]) == length for x in recs[lo:hi]), (b, kw, end)
            checked += 1
    assert checked > 3000
    # (4) shards with leads sum to the whole and reproduce it
    parts, lmax = [], m.max_keyword_length
    for (a, bnd, lead) in ac75().plan_shards(n, 8, lmax):
        k = m.scan_device(d_text.data_ptr() + a - lead, bnd - a + lead, lead=lead, base=a - lead, d_matches_ptr=d_out.data_ptr(), capacity=cap)
        parts.append(d_out[: k * 16].cpu().numpy().view(ac75().MATCH_DTYPE).copy())
    assert sum(len(p) for p in parts) == total and np.array_equal(np.concatenate(parts), recs)
    # (5) the two filter kernels are independent implementations of the same first stage: at the full size the stride-2 kernel
    # (used above) and the one-test-per-position kernel must produce the same records, byte for byte
    assert m.stats()["filter_stride"] == 2
    m.set_option("stride2", 0)
    total1 = m.scan_device(d_text.data_ptr(), n, d_matches_ptr=d_out.data_ptr(), capacity=cap)
    assert m.stats()["filter_stride"] == 1 and total1 == total
    assert np.array_equal(d_out[: total1 * 16].cpu().numpy().view(ac75().MATCH_DTYPE), recs)
    m.close(), o.close()


def test_dense_fallback_runs_in_segments():
    """A text whose candidate density is far above the filter's false-positive rate (4-letter alphabet: one position in five ends
    with some keyword's last 4 symbols) overflows the sparse staging; the scan must fall back to the dense mode, which walks the
    text in 64 Mi-symbol segments with left context, and still return exactly the oracle's records."""
    rng = np.random.default_rng(21)
    flat, offsets = random_patterns(60, lmin=7, lmax=11, seed=21, alphabet=4)
    n = (64 << 20) + (3 << 20) + 12345  # crosses one segment boundary
    text = rng.integers(0, 4, size=n).astype(np.uint8)
    o = pyoracle.Oracle("port", 1)
    o.insert_many(flat=flat, offsets=offsets)
    want = o.scan(text, cap=1 << 24)
    m = ac75().Machine(1)
    m.insert_many(flat=flat, offsets=offsets)
    m.set_option("engine", "filter")
    got = m.scan(text, capacity=1 << 24)
    st = m.stats()
    # a prefix probe finds the text dense in candidates: it goes straight to the dense mode, no doomed sparse attempt
    assert st["engine"] == "filter" and st["fallback_count"] == 0 and st["dense_scans"] == 1
    assert len(want) > 10_000 and np.array_equal(got, want), (len(got), len(want))
    # with a lead and a base, as a shard would be scanned
    lead = 5_000_000
    got2 = m.scan(text, lead=lead, base=10**12, capacity=1 << 24)
    w2 = want[want["end"] >= lead].copy()
    w2["end"] += 10**12
    assert np.array_equal(got2, w2)
    # the second dense text went straight to the dense mode too ...
    assert m.stats()["fallback_count"] == 0 and m.stats()["dense_scans"] == 2
    # ... and a sparse text afterwards is still scanned exactly, after which the fast mode is back
    sparse = rng.integers(4, 256, size=1 << 20).astype(np.uint8)
    k = flat[int(offsets[3]):int(offsets[4])]
    sparse[5000:5000 + len(k)] = k
    o.reset_cursor()
    want3 = o.scan(sparse, cap=1 << 16)
    for _ in range(2):
        assert np.array_equal(m.scan(sparse, capacity=1 << 16), want3) and len(want3) >= 1
    assert m.stats()["fallback_count"] == 0
    m.close(), o.close()


@pytest.mark.slow
def test_config2_full_size_against_the_oracle():
    """configs[1] at its FULL size: the 1k most frequent words of the novel over 1 GiB of synthetic printable ASCII, every one of the
    ~72 M records compared with the oracle's, in order."""
    import gzip
    import os
    import re
    from collections import Counter

    from helpers import ROOT

    novel = gzip.open(os.path.join(ROOT, "tests", "golden", "mrs_dalloway.txt.gz"), "rb").read()
    cnt = Counter(re.findall(rb"[a-z]+", novel.lower()))
    words = [w for w, _ in sorted(cnt.items(), key=lambda kv: (-kv[1], kv[0]))[:1000]]
    flat = np.frombuffer(b"".join(words), dtype=np.uint8)
    offsets = np.concatenate([[0], np.cumsum([len(w) for w in words])]).astype(np.uint64)
    n = 1 << 30
    text = generate_text(n, kind=1, plant_period=4096, dict_flat=flat, dict_offsets=offsets)
    m = ac75().Machine(1)
    m.insert_many(flat=flat, offsets=offsets)
    got = m.scan(text, capacity=n // 8)
    st = m.stats()
    assert st["engine"] == "dfa_smem" and len(got) > 60_000_000
    o = pyoracle.Oracle("port", 1)
    o.insert_many(flat=flat, offsets=offsets)
    piece = 128 << 20  # the oracle's carried cursor makes the pieces one scan; compared piece by piece to bound host memory
    at = 0
    for p0 in range(0, n, piece):
        want = o.scan(text[p0:p0 + piece], base=p0, cap=piece // 8)
        assert np.array_equal(got[at:at + len(want)], want), p0
        at += len(want)
    assert at == len(got)
    m.close(), o.close()


@pytest.mark.slow
def test_config4_full_dictionary_prefix_and_shard_boundaries():
    """configs[3] at its stated dictionary size: 1,000,000 patterns (DFA far beyond shared memory).  Records compared with the oracle
    on a 64 MiB prefix of the 16 GiB text and on windows around every boundary of its 8-way split (scanned as shards: lead + base)."""
    flat, offsets = random_patterns(1_000_000, seed=0xD1C7)
    o = pyoracle.Oracle("port", 1)
    ranks = o.insert_many(flat=flat, offsets=offsets)
    m = ac75().Machine(1)
    ids = m.insert_many(flat=flat, offsets=offsets)
    assert np.array_equal(ranks, ids)
    pre = 64 << 20
    text = generate_text(pre, kind=0, plant_period=4096, dict_flat=flat, dict_offsets=offsets)
    want = o.scan(text, cap=1 << 22)
    got = m.scan(text, capacity=1 << 22)
    st = m.stats()
    assert st["engine"] == "filter" and st["nb_keywords"] == m.nb_keywords and st["nb_states"] > 15_000_000
    assert len(want) > 16_000 and np.array_equal(got, want)
    lmax = m.max_keyword_length
    total, world, half = 16 << 30, 8, 1 << 20
    for (a, b, lead) in ac75().plan_shards(total, world, lmax)[1:]:
        # the shard that starts at a, cut down to its first MiB, exactly as rank g would scan it ...
        win = generate_text(half + lead, first=a - lead, kind=0, plant_period=4096, dict_flat=flat, dict_offsets=offsets)
        right = m.scan(win, lead=lead, base=a - lead, capacity=1 << 16)
        # ... and the last MiB of the shard before it
        prev = generate_text(half, first=a - half, kind=0, plant_period=4096, dict_flat=flat, dict_offsets=offsets)
        left = m.scan(prev, base=a - half, capacity=1 << 16)
        # the oracle over the two MiB as one text: what a single scan reports there, minus what needs context before the window
        o.reset_cursor()
        both = o.scan(np.concatenate([prev, win[lead:]]), base=a - half, cap=1 << 16)
        mine = np.concatenate([left, right])
        keep = both["end"] - (both["len"].astype(np.uint64) - 1) >= np.uint64(a - half)  # occurrences wholly inside the two MiB
        assert np.array_equal(mine[mine["end"] - (mine["len"].astype(np.uint64) - 1) >= np.uint64(a - half)], both[keep]) and len(both) > 300
    m.close(), o.close()


@pytest.mark.slow
def test_config5_full_dictionary_ten_meyer_rounds():
    """configs[4] at its stated dictionary size: 200,000 n-grams over a 50k vocabulary, then 10 Meyer rounds of +2,000 n-grams, each
    followed by a scan of the next slice with the cursor carried (reference aho_corasick_generic_test.c:184-228 interleaves the
    same way).  Every round's records equal the oracle's carried-cursor scan; the tables are updated, not rebuilt, where possible."""
    vocab, rounds, add = 50_000, 10, 2000
    per = 1 << 20
    stream = zipf_tokens(per * (rounds + 1), vocab, 42)
    rng = np.random.default_rng(43)

    def cut(k):
        starts, lens = rng.integers(0, len(stream) - 8, size=k), rng.integers(2, 9, size=k)
        flat = np.concatenate([stream[a:a + l] for a, l in zip(starts, lens)])
        return flat, np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)

    o = pyoracle.Oracle("port", 4)
    m = ac75().Machine(4)
    flat, offsets = cut(200_000)
    assert np.array_equal(o.insert_many(flat=flat, offsets=offsets), m.insert_many(flat=flat, offsets=offsets))
    for r in range(rounds + 1):
        if r:
            flat, offsets = cut(add)
            assert np.array_equal(o.insert_many(flat=flat, offsets=offsets), m.insert_many(flat=flat, offsets=offsets))
        sl = stream[r * per:(r + 1) * per]
        want = o.scan(sl, base=r * per, cap=1 << 24)
        got = m.scan(sl, base=r * per, carry=True, capacity=1 << 24)
        assert len(want) > 100_000 and np.array_equal(got, want), (r, len(got), len(want))
    st = m.stats()
    assert st["engine"] == "filter" and st["symbol_width"] == 4 and st["finalise_count"] == rounds + 1 and st["nb_keywords"] == m.nb_keywords > 190_000
    assert st["patch_count"] == rounds  # every Meyer round updated the resident tables in place
    m.close(), o.close()


@pytest.mark.parametrize("width,alphabet", [(4, 50_000), (4, 40), (2, 300), (2, 6)])
def test_in_place_table_update_equals_rebuild(width, alphabet):
    """Append-only insertions between scans (Meyer-style): the tables patched in place give the records of the oracle's
    carried-cursor scan and of a machine that rebuilds everything (option patch=0), for nested / suffix / extending keywords too."""
    rng = np.random.default_rng(width * 1000 + alphabet)
    dt = {2: np.uint16, 4: np.uint32}[width]
    n = 400_000
    text = rng.integers(0, alphabet, size=n).astype(dt)

    def cut(k, lmin=2):
        out = []
        for _ in range(k):
            a, l = int(rng.integers(0, n - 12)), int(rng.integers(lmin, 10))
            kw = text[a:a + l].copy()
            r = rng.random()
            if r < 0.2 and out:  # an earlier keyword with symbols prepended: the old one becomes a proper suffix
                kw = np.concatenate([text[a:a + 2], out[int(rng.integers(0, len(out)))]])
            elif r < 0.3 and out and len(out[-1]) > 3:  # a proper suffix of an earlier keyword
                kw = out[-1][1:].copy()
            out.append(kw)
        return out

    o = pyoracle.Oracle("port", width)
    m, full = ac75().Machine(width), ac75().Machine(width)
    full.set_option("patch", 0)
    first = cut(3000)
    for mach in (o, m, full):
        mach.insert_many(first)
    rounds, per = 8, n // 9
    for r in range(rounds + 1):
        if r:
            more = cut(int(rng.integers(1, 400)))
            ids = [mach.insert_many(more) for mach in (o, m, full)]
            assert np.array_equal(ids[0], ids[1]) and np.array_equal(ids[0], ids[2])
        sl = text[r * per:(r + 1) * per]
        want = o.scan(sl, base=r * per, cap=1 << 22)
        got = m.scan(sl, base=r * per, carry=True, capacity=1 << 22)
        ref = full.scan(sl, base=r * per, carry=True, capacity=1 << 22)
        assert np.array_equal(got, want) and np.array_equal(ref, want), (r, len(got), len(ref), len(want))
    # (a round may still rebuild: when the room the build left for growth is used up -- always, for the tiny tables of a 6-letter alphabet)
    assert m.stats()["patch_count"] >= (rounds - 2 if alphabet >= 40 else 0) and full.stats()["patch_count"] == 0 and full.stats()["finalise_count"] == rounds + 1
    for mach in (o, m, full):
        mach.close()

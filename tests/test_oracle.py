"""Pins the oracle (CPU only): every golden vector / known answer the reference holds for the scan path.

 - README.md:92-93 (exact match list and order of examples/test.c)
 - the reference's own generic test numbers (examples/aho_corasick_generic_test.c sub-test 3)
 - SURVEY.md Appendix B: config-1 counts and FNV-1a-64 hashes, both reference builds
 - oracle/ac_port.c (the restatement) == unmodified reference (Meyer'85 build) == unmodified reference (classic build)
   on random dictionaries with heavy overlap, with insertions interleaved with scanning on a carried cursor.
"""
import os
import re

import numpy as np
import pytest

from oracle import pyoracle
from conftest import oracle_kinds, GOLDEN

KINDS = oracle_kinds()


def config1_keywords(text, full):
    words = [b"he", b"she", b"his", b"hers"]
    if full:
        words += re.findall(rb"[A-Za-z]+", text)
    return words


@pytest.mark.parametrize("kind", KINDS)
def test_readme_example(kind):
    # README.md:92-93: " 6:he 5:she 6:hers 12:he 21:his 38:he 37:she 56:he 56:hers" with j = nb..1 (shortest first)
    o = pyoracle.Oracle(kind, 1)
    words = [b"he", b"she", b"his", b"hers"]
    assert o.insert_many(words).tolist() == [0, 1, 2, 3]
    text = b"To ushers: he found his pencil, but she could not find hers."
    r = o.scan(text)
    got = []
    for end in np.unique(r["end"]):
        grp = r[r["end"] == end]
        for rec in grp[::-1]:  # the example prints index nb-1 .. 0
            got.append(f"{int(rec['end']) + 2 - int(rec['len'])}:{words[rec['id']].decode()}")
    assert " " + " ".join(got) == " 6:he 5:she 6:hers 12:he 21:his 38:he 37:she 56:he 56:hers"
    line = open(os.path.join(GOLDEN, "readme_example.stdout")).read().splitlines()[1]
    assert line == " " + " ".join(got)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("which", ["readme_only", "readme_plus_wordlist"])
def test_config1_golden(kind, which, novel, golden_config1):
    g = golden_config1[which]
    o = pyoracle.Oracle(kind, 1)
    o.insert_many(config1_keywords(novel, which != "readme_only"))
    r = o.scan(novel)
    assert o.nb_keywords == g["keywords"]
    assert len(r) == g["matches"]
    assert len(np.unique(r["end"])) == g["positions"]
    assert "%016x" % pyoracle.fnv1a64_records(r) == g["fnv1a64"]
    assert [list(map(int, x)) for x in r[:48].tolist()] == g["head"]
    # emission order is (end asc, len desc)
    assert np.array_equal(pyoracle.sort_records(r), r)


def test_survey_appendix_b_numbers(golden_config1):
    assert golden_config1["readme_only"]["matches"] == 11676 and golden_config1["readme_only"]["fnv1a64"] == "4cb7510699888d13"
    assert golden_config1["readme_only"]["positions"] == 10403
    g = golden_config1["readme_plus_wordlist"]
    assert (g["keywords"], g["lmax"], g["matches"], g["positions"], g["fnv1a64"]) == (7394, 17, 298855, 186817, "e6ce99c887bfa45c")


@pytest.mark.parametrize("kind", KINDS)
def test_small_kats(kind, golden_kats):
    for kat in golden_kats:
        o = pyoracle.Oracle(kind, kat["width"])
        for step, want in zip(kat["steps"], kat["results"]):
            if step[0] == "insert":
                assert o.insert_many([k.encode("latin1") for k in step[1]]).tolist() == want["ranks"], kat["name"]
            elif step[0] == "scan":
                got = [list(map(int, x)) for x in o.scan(step[1].encode("latin1")).tolist()]
                assert got == want["records"], kat["name"]
            else:
                o.reset_cursor()


def test_carried_cursor_known_answers(golden_kats):
    # {zz}, scan "he", insert "hers", scan "rs" -> 0 matches; a fresh scan of "hers" gives 1
    k = {x["name"]: x for x in golden_kats}
    res = k["carry_zz_hers"]["results"]
    assert res[3]["records"] == [] and len(res[5]["records"]) == 1
    # cursor inside "abcd", insert "bc", continue "cd" -> bc@2 (local 0) and abcd@3 (local 1)
    assert k["carry_abcd_bc"]["results"][3]["records"] == [[0, 1, 2], [1, 0, 4]]


def _rand_keywords(rng, n, alphabet, lmin, lmax, width):
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[width]
    return [rng.choice(alphabet, size=rng.integers(lmin, lmax + 1)).astype(dt) for _ in range(n)]


@pytest.mark.parametrize("width,alphabet", [(1, [97, 98]), (1, [0, 1, 255]), (2, [7, 300, 65535]), (4, [5, 70000, 2**32 - 1, 9])])
def test_three_oracles_agree_random_interleaved(width, alphabet):
    if len(KINDS) < 3:
        pytest.skip("reference oracle not built (no /root/reference and no prebuilt oracle/_ref)")
    rng = np.random.default_rng(1234 + width + len(alphabet))
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[width]
    for trial in range(12):
        os_ = [pyoracle.Oracle(k, width) for k in KINDS]
        for rnd in range(5):
            kws = _rand_keywords(rng, int(rng.integers(1, 40)), alphabet, 1, 9, width)
            ranks = [o.insert_many(kws).tolist() for o in os_]
            assert ranks[0] == ranks[1] == ranks[2]
            text = rng.choice(alphabet, size=int(rng.integers(0, 600))).astype(dt)
            recs = [o.scan(text, base=1000 * rnd) for o in os_]  # cursor carried across rounds
            assert np.array_equal(recs[0], recs[1]) and np.array_equal(recs[0], recs[2]), (trial, rnd)
        for o in os_:
            o.close()


def test_reference_generic_test3_numbers():
    import json

    g = json.load(open(os.path.join(GOLDEN, "generic_test3.json")))
    assert g["keywords"] == [25000, 50000, 74999, 99996, 124996, 149996, 174996, 199996, 224996, 249995]
    assert g["matches"] == [3, 3, 3, 13, 30, 21, 19, 26, 31, 30]

"""keyword id -> MatchHolder (acm_b200_keyword) and the id table in acm_foreach_keyword order
(acm_b200_keyword_order), against the library's own reference-shaped API (C program) and against the reference itself."""
import os
import subprocess

import numpy as np
import pytest

from helpers import ac75, random_patterns
from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "aho-corasick-1975_b200", "csrc")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    d = tmp_path_factory.mktemp("kwcheck")
    stub = d / "stub.c"
    stub.write_text("struct acm_device_image; void acm_device_release (struct acm_device_image *i) { (void)i; }\n")
    exe = d / "keyword_lookup_check"
    subprocess.run(["gcc", "-std=c11", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", str(exe), os.path.join(ROOT, "tests", "csrc", "keyword_lookup_check.c"),
                    os.path.join(CSRC, "acm_host.c"), os.path.join(CSRC, "acm_finalise.c"), str(stub), "-lpthread"], check=True)
    return str(exe)


@pytest.mark.parametrize("args", [("2000", "6"), ("20000", "3"), ("500", "26")])
def test_keyword_by_id_equals_get_match(checker, args):
    r = subprocess.run([checker, *args], capture_output=True, text=True)
    assert r.returncode == 0 and "errors 0" in r.stdout, (r.stdout, r.stderr[-1000:])


@pytest.mark.parametrize("kind", ["ref_meyer", "ref_classic"])
@pytest.mark.parametrize("width,alphabet,n", [(1, 256, 3000), (1, 4, 2000), (2, 1000, 1500), (4, 50000, 1500)])
def test_keyword_order_equals_the_reference_enumeration(kind, width, alphabet, n):
    if not pyoracle.available(kind):
        pytest.skip("reference build not present")
    flat, offsets = random_patterns(n, lmin=1, lmax=9, seed=width * 100 + n, alphabet=alphabet, width=width)
    o = pyoracle.Oracle(kind, width)
    ranks = o.insert_many(flat=flat, offsets=offsets)
    m = ac75().Machine(width)
    ids = m.insert_many(flat=flat, offsets=offsets)
    assert np.array_equal(ranks, ids)
    want = o.foreach_ranks()
    got = m.keyword_order()
    assert len(want) == o.nb_keywords == m.nb_keywords and np.array_equal(got, want)
    m.close(), o.close()

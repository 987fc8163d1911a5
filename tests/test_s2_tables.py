"""CPU check of the stride-2 filter tables (window choice, distance table, kw_dist): tests/csrc/s2_tables_check.c emulates what the
stride-2 kernels do with them and compares with the per-symbol host API -- every occurrence must be reported exactly once."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "aho-corasick-1975_b200", "csrc")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    d = tmp_path_factory.mktemp("s2check")
    stub = d / "stub.c"
    stub.write_text("struct acm_device_image; void acm_device_release (struct acm_device_image *i) { (void)i; }\n")
    exe = d / "s2_tables_check"
    subprocess.run(["gcc", "-std=c11", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", str(exe), os.path.join(ROOT, "tests", "csrc", "s2_tables_check.c"),
                    os.path.join(CSRC, "acm_host.c"), os.path.join(CSRC, "acm_finalise.c"), str(stub), "-lpthread"], check=True)
    return str(exe)


@pytest.mark.parametrize("args", [
    ("3000", "256", "300000"),                       # random bytes, nested keywords, dense plants
    ("20000", "256", "500000"),
    ("100000", "256", "1000000"),                    # config-3 sized dictionary
    ("2000", "256", "100000", "64"),                 # keywords longer than the largest distance a window may choose
    ("100", "256", "50000", "8"),                    # short keywords: few windows to choose from
    ("3000", "4", "200000", "12", "195", "filter"),  # tiny alphabets: shared windows, chained distance-table words
    ("5000", "2", "100000", "20", "195", "filter"),
    ("300", "1", "5000", "30", "195", "filter"),
    ("3000", "256", "300000", "32", "100"),          # a smaller shared-memory carve-out
])
def test_every_occurrence_reported_exactly_once(checker, args):
    r = subprocess.run([checker, *args], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-2000:])
    assert "errors 0" in r.stdout

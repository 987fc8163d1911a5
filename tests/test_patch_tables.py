"""CPU check of the in-place table update: tests/csrc/patch_tables_check.c appends keywords in rounds, patches the
filter tables in place and compares them, after every round, with a from-scratch build of the same dictionary."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "aho-corasick-1975_b200", "csrc")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    d = tmp_path_factory.mktemp("patchcheck")
    stub = d / "stub.c"
    stub.write_text("struct acm_device_image; void acm_device_release (struct acm_device_image *i) { (void)i; }\n")
    exe = d / "patch_tables_check"
    subprocess.run(["gcc", "-std=c11", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", str(exe), os.path.join(ROOT, "tests", "csrc", "patch_tables_check.c"),
                    os.path.join(CSRC, "acm_host.c"), os.path.join(CSRC, "acm_finalise.c"), str(stub), "-lpthread"], check=True)
    return str(exe)


@pytest.mark.parametrize("args", [("3000", "100", "40"), ("20000", "500", "50000"), ("2000", "60", "5"), ("50000", "2000", "300")])
def test_patched_tables_equal_rebuilt_tables(checker, args):
    r = subprocess.run([checker, *args], capture_output=True, text=True)
    assert r.returncode == 0 and "errors 0" in r.stdout, (r.stdout[-1500:], r.stderr[-500:])

"""CPU check of the open-addressing tables the verification kernel walks (q-gram -> node, reverse-trie edges):
tests/csrc/slot_tables_check.c builds the filter tables of random dictionaries and checks their fill (at most a third), that every key
is reachable by the GPU's probe sequence, and that the probe sequences are as short as linear probing over a uniform hash gives
(the 4-instruction multiplicative hash of acm_tables.h, which replaced a 21-instruction finaliser)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "aho-corasick-1975_b200", "csrc")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    d = tmp_path_factory.mktemp("slotcheck")
    stub = d / "stub.c"
    stub.write_text("struct acm_device_image; void acm_device_release (struct acm_device_image *i) { (void)i; }\n")
    exe = d / "slot_tables_check"
    subprocess.run(["gcc", "-std=c11", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", str(exe), os.path.join(ROOT, "tests", "csrc", "slot_tables_check.c"),
                    os.path.join(CSRC, "acm_host.c"), os.path.join(CSRC, "acm_finalise.c"), str(stub), "-lpthread"], check=True)
    return str(exe)


@pytest.mark.parametrize("args", [
    ("20000", "256", "1"),     # random bytes
    ("100000", "256", "1"),    # config-3 sized dictionary
    ("30000", "5", "1"),       # tiny alphabet: few distinct q-grams, a bushy reverse trie
    ("50000", "3000", "4"),    # 32-bit symbols: 64-bit keys (node << 32 | symbol, two packed symbols)
])
def test_slot_tables_are_sparse_and_probe_sequences_short(checker, args):
    r = subprocess.run([checker, *args], capture_output=True, text=True)
    assert r.returncode == 0 and "errors 0" in r.stdout, (r.stdout[-1500:], r.stderr[-500:])

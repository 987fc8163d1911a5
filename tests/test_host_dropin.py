"""CPU tests of the drop-in boundary: the reference's own example programs, UNMODIFIED, compiled against include/aho_corasick.h
and linked to libac75.so must print what the reference prints (golden stdout generated from the unmodified reference, tests/golden)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE, ROOT
from helpers import ac75
from oracle import pyoracle

LIBDIR = os.path.join(ROOT, "aho-corasick-1975_b200")
INCLUDE = os.path.join(ROOT, "include")
needs_reference = pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "examples", "test.c")), reason="reference examples not mounted")


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    ac75().build_library()
    return tmp_path_factory.mktemp("dropin")


def compile_against_product(src, out, extra=()):
    subprocess.run(["gcc", "-std=c11", "-O2", *extra, f"-I{INCLUDE}", "-o", str(out), src, f"-L{LIBDIR}", "-lac75", f"-Wl,-rpath,{LIBDIR}", "-lpthread"], check=True)


@needs_reference
def test_readme_example_program(built):
    exe = built / "test"
    compile_against_product(os.path.join(REFERENCE, "examples", "test.c"), exe)
    out = subprocess.run([str(exe)], capture_output=True, check=True).stdout
    assert out == open(os.path.join(GOLDEN, "readme_example.stdout"), "rb").read()
    assert out.splitlines()[1] == b" 6:he 5:she 6:hers 12:he 21:his 38:he 37:she 56:he 56:hers"  # README.md:92-93


@needs_reference
def test_generic_test_program_subtests_1_2(built):
    """wchar_t letters, case-insensitive comparator, malloc'd letters with free as destructor, duplicate keywords, values,
    acm_foreach_keyword order, acm_print layout, insertions interleaved with scanning (6,966 keywords)."""
    exe = built / "generic"
    compile_against_product(os.path.join(REFERENCE, "examples", "aho_corasick_generic_test.c"), exe, extra=("-D_XOPEN_SOURCE=700",))
    env = dict(os.environ, LC_ALL="C.utf8")
    out = subprocess.run([str(exe), "3"], capture_output=True, check=True, cwd=os.path.join(REFERENCE, "examples"), env=env).stdout
    got = b"\n".join(l for l in out.split(b"\n") if not l.startswith(b"Elapsed CPU time") and b"in use." not in l)
    assert got == open(os.path.join(GOLDEN, "generic_test12.stdout"), "rb").read()
    assert b"Incremental string matching (Meyer, 1985) in use." in out


@needs_reference
@pytest.mark.slow
def test_generic_test_program_subtest_3(built):
    import json

    exe = built / "generic3"
    compile_against_product(os.path.join(REFERENCE, "examples", "aho_corasick_generic_test.c"), exe, extra=("-D_XOPEN_SOURCE=700",))
    out = subprocess.run([str(exe), "4"], capture_output=True, check=True, cwd=os.path.join(REFERENCE, "examples")).stdout.decode()
    g = json.load(open(os.path.join(GOLDEN, "generic_test3.json")))
    assert [int(x) for x in re.findall(r"amongst (\d+) keywords", out)] == g["keywords"]
    assert [int(x) for x in re.findall(r"\] (\d+) matches found", out)] == g["matches"]


@pytest.fixture(scope="module")
def product_harness(built):
    """oracle/ref_harness.c drives a library through the reference's public API only -- compile it against OUR library."""
    so = built / "libproduct_harness.so"
    subprocess.run(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-D_POSIX_C_SOURCE=200809L", f"-I{INCLUDE}", "-o", str(so), os.path.join(ROOT, "oracle", "ref_harness.c"),
                    f"-L{LIBDIR}", "-lac75", f"-Wl,-rpath,{LIBDIR}", "-lpthread"], check=True)
    pyoracle._PATHS["product_host"] = (str(so), "refh")
    return "product_host"


def test_per_symbol_api_equals_reference_on_novel(product_harness, novel, golden_config1):
    for which, full in (("readme_only", False), ("readme_plus_wordlist", True)):
        words = [b"he", b"she", b"his", b"hers"] + (re.findall(rb"[A-Za-z]+", novel) if full else [])
        o = pyoracle.Oracle(product_harness, 1)
        o.insert_many(words)
        r = o.scan(novel)
        g = golden_config1[which]
        assert (o.nb_keywords, len(r)) == (g["keywords"], g["matches"])
        assert "%016x" % pyoracle.fnv1a64_records(r) == g["fnv1a64"]
        o.close()


def test_per_symbol_api_kats_and_random_interleaved(product_harness, golden_kats):
    for kat in golden_kats:
        o = pyoracle.Oracle(product_harness, kat["width"])
        for step, want in zip(kat["steps"], kat["results"]):
            if step[0] == "insert":
                assert o.insert_many([k.encode("latin1") for k in step[1]]).tolist() == want["ranks"], kat["name"]
            elif step[0] == "scan":
                assert [list(map(int, x)) for x in o.scan(step[1].encode("latin1")).tolist()] == want["records"], kat["name"]
            else:
                o.reset_cursor()
        o.close()
    rng = np.random.default_rng(11)
    for width, alphabet in ((1, [97, 98]), (2, [7, 300, 65535]), (4, [5, 70000, 2**32 - 1, 9])):
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[width]
        for trial in range(8):
            a, b = pyoracle.Oracle(product_harness, width), pyoracle.Oracle("port", width)
            for rnd in range(5):
                kws = [rng.choice(alphabet, size=rng.integers(1, 9)).astype(dt) for _ in range(int(rng.integers(1, 40)))]
                assert a.insert_many(kws).tolist() == b.insert_many(kws).tolist()
                text = rng.choice(alphabet, size=int(rng.integers(0, 500))).astype(dt)
                assert np.array_equal(a.scan(text), b.scan(text)), (width, trial, rnd)
            a.close(), b.close()


def test_bulk_insert_equals_incremental(product_harness):
    """acm_b200_insert_keywords (one BFS pass) must leave the machine exactly as per-letter insertion (Meyer incremental) does."""
    rng = np.random.default_rng(5)
    kws = [rng.integers(97, 100, size=rng.integers(1, 8)).astype(np.uint8).tobytes() for _ in range(1500)]
    text = rng.integers(97, 100, size=20000).astype(np.uint8)
    m = ac75().Machine(1)
    ids = m.insert_many(kws)  # >= 256 keywords: bulk path
    o = pyoracle.Oracle(product_harness, 1)
    assert np.array_equal(o.insert_many(kws), ids)
    assert m.host_match_count(text) == o.count(text)
    m.close(), o.close()


def test_abi_exports_every_declared_symbol():
    """libac75.so loads without a GPU and exports every symbol include/*.h declares, and nothing internal."""
    lib = ctypes.CDLL(ac75().library_path())
    declared = set()
    for h in ("aho_corasick.h", "acm_b200.h"):
        src = open(os.path.join(INCLUDE, h)).read()
        declared |= set(re.findall(r"\b(acm_\w+)\s*\(", src))
        declared |= set(re.findall(r"extern const \w+ (ACM_\w+);", src))
    assert len(declared) >= 28
    for name in sorted(declared):
        assert hasattr(lib, name), name
    exported = subprocess.run(["nm", "-D", "--defined-only", ac75().library_path()], capture_output=True, text=True, check=True).stdout
    names = {l.split()[-1] for l in exported.splitlines() if l.split() and l.split()[-2] in "TDBR"}
    assert {n for n in names if n.lower().startswith("acm_")} == declared


def test_batch_scan_fails_loudly_without_gpu():
    if ac75().device_count() > 0:
        pytest.skip("a GPU is present")
    m = ac75().Machine(1)
    m.insert_many([b"abc"])
    with pytest.raises(ac75().AcmError) as e:
        m.scan(b"xxabcxx")
    assert e.value.code == 2  # ACM_B200_ERR_NO_DEVICE: there is no CPU fallback
    m.close()


def test_error_convention_of_legacy_api(built, tmp_path):
    """Contract violations: message on stderr, thrd_exit(EXIT_FAILURE) (reference aho_corasick.c:24-36); from main the process ends with 0."""
    src = tmp_path / "bad.c"
    src.write_text('#include "aho_corasick.h"\nint main(void){ ACMachine *m = acm_create (ACM_CMP_DEFAULT, &(size_t){1}, 0); ACState *s = acm_initiate (m);'
                   ' printf("before\\n"); acm_insert_end_of_keyword (&s, 0, 0); printf("not reached\\n"); return 7; }\n')
    exe = tmp_path / "bad"
    compile_against_product(str(src), exe)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert "before" in r.stdout and "not reached" not in r.stdout
    assert r.stderr.startswith("FATAL ERROR: A prerequisite is not fulfilled in function acm_insert_end_of_keyword.")
    assert r.returncode == 0

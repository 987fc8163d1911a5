"""Regenerates tests/golden/ from the reference mounted at /root/reference (run in the build container only).

Everything written here comes from the UNMODIFIED reference source compiled by oracle/Makefile (oracle/_ref/), or is the
reference's own data fixture:
  mrs_dalloway.txt.gz      : examples/mrs_dalloway.txt (the config-1 text; data fixture, not source), gzip'd
  readme_example.stdout    : stdout of examples/test.c (must equal README.md:92-93)
  generic_test12.stdout    : stdout of examples/aho_corasick_generic_test.c sub-tests 1+2 under LC_ALL=C.utf8, timing line removed
  generic_test3.json       : keyword totals / match counts printed by sub-test 3 (glibc rand(), unseeded)
  config1.json             : counts, FNV-1a-64 and sample records of the two config-1 dictionaries, both reference builds
  kat_small.json           : small known-answer cases (carried cursor, duplicates, nested keywords) with full record lists
"""
import gzip, json, os, re, subprocess, sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def config1_keywords(text, full):
    words = [b"he", b"she", b"his", b"hers"]
    if full:
        words += re.findall(rb"[A-Za-z]+", text)  # maximal runs, case preserved, first-occurrence order
    return words


def main():
    pyoracle.build()
    text = open(f"{REF}/examples/mrs_dalloway.txt", "rb").read()
    with gzip.GzipFile(os.path.join(OUT, "mrs_dalloway.txt.gz"), "wb", mtime=0) as f:
        f.write(text)
    env = dict(os.environ, LC_ALL="C.utf8")
    outs = {}
    for mode in ("meyer", "classic"):
        outs[mode, "t"] = subprocess.run([f"{ROOT}/oracle/_ref/ref_test_{mode}"], capture_output=True, check=True).stdout
        g = subprocess.run([f"{ROOT}/oracle/_ref/ref_generic_test_{mode}", "3"], capture_output=True, check=True, cwd=f"{REF}/examples", env=env).stdout
        outs[mode, "g"] = b"\n".join(l for l in g.split(b"\n") if not l.startswith(b"Elapsed CPU time") and b"in use." not in l)
        g3 = subprocess.run([f"{ROOT}/oracle/_ref/ref_generic_test_{mode}", "4"], capture_output=True, check=True, cwd=f"{REF}/examples", env=env).stdout.decode()
        outs[mode, "3"] = {"keywords": [int(x) for x in re.findall(r"amongst (\d+) keywords", g3)], "matches": [int(x) for x in re.findall(r"\] (\d+) matches found", g3)]}
    assert outs["meyer", "t"] == outs["classic", "t"] and outs["meyer", "g"] == outs["classic", "g"] and outs["meyer", "3"] == outs["classic", "3"]
    open(os.path.join(OUT, "readme_example.stdout"), "wb").write(outs["meyer", "t"])
    open(os.path.join(OUT, "generic_test12.stdout"), "wb").write(outs["meyer", "g"])
    json.dump(outs["meyer", "3"], open(os.path.join(OUT, "generic_test3.json"), "w"), indent=1)

    cfg = {}
    for full in (0, 1):
        per_mode = []
        for kind in ("ref_meyer", "ref_classic"):
            o = pyoracle.Oracle(kind, 1)
            o.insert_many(config1_keywords(text, full))
            r = o.scan(text)
            per_mode.append({
                "keywords": o.nb_keywords, "matches": len(r), "positions": int(len(np.unique(r["end"]))),
                "lmax": int(r["len"].max()), "fnv1a64": "%016x" % pyoracle.fnv1a64_records(r),
                "head": [[int(a), int(b), int(c)] for a, b, c in r[:48].tolist()],
                "tail": [[int(a), int(b), int(c)] for a, b, c in r[-48:].tolist()],
                "per_id_counts_first8": np.bincount(r["id"], minlength=8)[:8].tolist(),
            })
            o.close()
        assert per_mode[0] == per_mode[1], "Meyer and classic builds of the reference disagree"
        cfg["readme_plus_wordlist" if full else "readme_only"] = per_mode[0]
    json.dump(cfg, open(os.path.join(OUT, "config1.json"), "w"), indent=1)

    # small KATs, full record lists, from the reference (Meyer build), cross-checked with the classic build
    kats = []

    def kat(name, steps, width=1):
        """steps: list of ("insert", [keywords]) / ("scan", text) / ("reset",); records of every scan are stored."""
        res = []
        for kind in ("ref_meyer", "ref_classic"):
            o = pyoracle.Oracle(kind, width)
            scans = []
            for st in steps:
                if st[0] == "insert":
                    scans.append({"ranks": o.insert_many(st[1]).tolist()})
                elif st[0] == "scan":
                    scans.append({"records": [[int(a), int(b), int(c)] for a, b, c in o.scan(st[1]).tolist()]})
                else:
                    o.reset_cursor()
                    scans.append({})
            res.append(scans)
            o.close()
        assert res[0] == res[1], name
        enc = [[s[0]] + ([[k.decode("latin1") for k in s[1]]] if s[0] == "insert" else [s[1].decode("latin1")] if s[0] == "scan" else []) for s in steps]
        kats.append({"name": name, "width": width, "steps": enc, "results": res[0]})

    kat("readme", [("insert", [b"he", b"she", b"his", b"hers"]), ("scan", b"To ushers: he found his pencil, but she could not find hers.")])
    kat("paper_graph_dups", [("insert", [b"he", b"she", b"sheers", b"his", b"hi", b"hers", b"ushers", b"abcde", b"bcd", b"hers", b"hen", b"hen", b"bcdef", b"pen", b"cdefg", b"pen", b"bcd", b"abc", b"abcd", b"abcde", b"bcde", b"cde", b"cd", b"bc", b"u", b"uu"]),
                              ("scan", b"he found his pencil, but she could not find hers (hi! ushers !! --abcdefgh--)")])
    kat("carry_zz_hers", [("insert", [b"zz"]), ("scan", b"he"), ("insert", [b"hers"]), ("scan", b"rs"), ("reset",), ("scan", b"hers")])
    kat("carry_abcd_bc", [("insert", [b"abcd"]), ("scan", b"ab"), ("insert", [b"bc"]), ("scan", b"cd")])
    kat("carry_she_he", [("insert", [b"she"]), ("scan", b"she"), ("insert", [b"he"]), ("scan", b"she")])
    kat("aaa_runs", [("insert", [b"a", b"ba", b"baa", b"baaa"]), ("scan", b"abaaabaa"), ("insert", [b"aa"]), ("scan", b"baaaabaa"), ("insert", [b"aaa"]), ("scan", b"aaaabaaab")])
    kat("nested_binary", [("insert", [b"\x00", b"\x00\x00", b"\x00\xff\x00", b"\xff\x00\xff\x00"]), ("scan", b"\x00\x00\xff\x00\xff\x00\x00")])
    json.dump(kats, open(os.path.join(OUT, "kat_small.json"), "w"), indent=1)
    print("golden written:", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()

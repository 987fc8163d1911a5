"""Shared test helpers: package loader, dictionaries and texts of the BASELINE.json configs at test scale."""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
SUPPORT = os.path.join(ROOT, "tests", "support")
if SUPPORT not in sys.path:
    sys.path.insert(0, SUPPORT)
import __graft_entry__ as entry  # noqa: E402
import textgen  # noqa: E402  (tests/support: the synthetic-text generator, not part of the product library)
from oracle import pyoracle  # noqa: E402

generate_text = textgen.generate_text


def ac75():
    return entry.load_package()


def best_oracle_kind():
    return "ref_meyer" if pyoracle.available("ref_meyer") else "port"


def pack(keywords, width=1):
    return pyoracle.pack_keywords(keywords, width)


def config1_keywords(novel, full):
    words = [b"he", b"she", b"his", b"hers"]
    if full:
        words += re.findall(rb"[A-Za-z]+", novel)
    return words


def config2_keywords(novel, n=1000):
    """n most frequent lowercase alphabetic words of the novel (ties lexicographic), SURVEY.md 8(d)."""
    from collections import Counter

    cnt = Counter(re.findall(rb"[a-z]+", novel.lower()))
    return [w for w, _ in sorted(cnt.items(), key=lambda kv: (-kv[1], kv[0]))[:n]]


def random_patterns(n, lmin=4, lmax=32, seed=0xD1C7, alphabet=256, width=1):
    rng = np.random.default_rng(seed)
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[width]
    lens = rng.integers(lmin, lmax + 1, size=n)
    flat = rng.integers(0, alphabet, size=int(lens.sum())).astype(dt)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    return flat, offsets


def oracle_records(keywords=None, flat=None, offsets=None, text=None, width=1, kind=None):
    o = pyoracle.Oracle(kind or best_oracle_kind(), width)
    o.insert_many(keywords, flat=flat, offsets=offsets)
    r = o.scan(text, cap=max(1 << 16, 8 * len(text)))
    o.close()
    return r

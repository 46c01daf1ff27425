"""strings.ToLower tables: the product's (suggest_b200/csrc/unicode_lower.inc, tools/gen_unicode_lower.py: Python
unicodedata) against the oracle's (oracle/unicode_lower.inc, oracle/gen_unicode_lower.pl: Perl Unicode::UCD + Age).

Both implement Unicode 11.0.0, the tables of Go 1.13 (the release in the reference's go.mod); the two files are derived
from different copies of the Unicode Character Database by different code, so non-ASCII lower-casing of the product is
not checked against itself.  No GPU.
"""
import ctypes as C
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT
from oracle import oracle
from suggest_b200 import _capi

PRODUCT = os.path.join(ROOT, "suggest_b200", "csrc", "unicode_lower.inc")
ORACLE = os.path.join(ROOT, "oracle", "unicode_lower.inc")


def expand(path):
    table = {}
    for lo, hi, delta, step in re.findall(r"\{0x([0-9A-F]+), 0x([0-9A-F]+), (-?\d+), (\d+)\}", open(path).read()):
        for cp in range(int(lo, 16), int(hi, 16) + 1, int(step)):
            assert cp not in table
            table[cp] = cp + int(delta)
    return table


def product_lower(text):
    t = text.encode("utf-8")
    out = C.create_string_buffer(3 * len(t) + 8)
    n = _capi.lib().sg_host_to_lower(t, len(t), out, 3 * len(t) + 8)
    return out.raw[:n].decode("utf-8")


def test_tables_come_from_different_generators_and_agree():
    assert "gen_unicode_lower.pl" in open(ORACLE).readline() and "gen_unicode_lower.py" in open(PRODUCT).readline()
    p, o = expand(PRODUCT), expand(ORACLE)
    assert len(p) == len(o) == 1383  # simple lower-case mappings of Unicode 11.0.0
    assert p == o


@pytest.mark.skipif(shutil.which("perl") is None, reason="perl not installed")
def test_oracle_table_is_what_its_generator_writes():
    out = subprocess.run(["perl", os.path.join(ROOT, "oracle", "gen_unicode_lower.pl")], capture_output=True, text=True)
    if out.returncode != 0 and "Unicode/UCD" in out.stderr:
        pytest.skip("perl without Unicode::UCD")
    assert out.returncode == 0, out.stderr
    if "Unicode::UCD 15.0.0" in out.stdout.splitlines()[0]:  # the committed file was written by this version
        assert out.stdout == open(ORACLE).read()
    else:  # another UCD version must still give the same Unicode 11 table
        tmp = os.path.join(os.environ.get("TMPDIR", "/tmp"), "sg_unicode_lower_regen.inc")
        open(tmp, "w").write(out.stdout)
        assert expand(tmp) == expand(ORACLE)


# (input, lowered) - what Go 1.13's unicode.ToLower does
VECTORS = [
    ("İ", "i"),                  # simple mapping of I WITH DOT ABOVE (Python's full mapping gives two runes)
    ("ᲐᲿ", "აჿ"),  # Georgian Mtavruli, added BY Unicode 11.0: lowered
    ("Ꞹ", "ꞹ"),             # U WITH STROKE, 11.0: lowered
    ("\U00016E40", "\U00016E60"),     # Medefaidrin, 11.0: lowered
    ("ẞ", "ß"),             # capital sharp s
    ("ΩKÅ", "ωkå"),  # Ohm, Kelvin, Angstrom signs
    ("ǅ", "ǆ"),             # title-case digraph
    ("Ꭰ", "ꭰ"),             # Cherokee
    # added AFTER Unicode 11.0: Go 1.13 has no mapping for them, they stay as they are
    ("ꞺꞼꞾ", "ꞺꞼꞾ"),  # 12.0 glottal A / I / U
    ("ꟂꟄꟅꟆ", "ꟂꟄꟅꟆ"),  # 12.0
    ("ꟇꟉꟵ", "ꟇꟉꟵ"),  # 13.0
    ("ⰯꟀꟐꟖꟘ", "ⰯꟀꟐꟖꟘ"),  # 14.0
    ("\U00010570\U00010595", "\U00010570\U00010595"),  # 14.0 Vithkuqi
]


@pytest.mark.parametrize("text,want", VECTORS)
def test_lower_vectors(text, want):
    assert oracle.to_lower(text.encode("utf-8")).decode("utf-8") == want
    assert product_lower(text) == want


def test_runes_changed_after_unicode_11_are_exactly_these():
    """the difference between this table and the newest one Python knows (15.0) is the set of later additions"""
    import unicodedata
    if unicodedata.unidata_version != "15.0.0":
        pytest.skip("written against unicodedata 15.0.0")
    table = expand(PRODUCT)
    newer = set()
    for cp in range(0x110000):
        if 0xD800 <= cp <= 0xDFFF or cp == 0x130:
            continue
        low = chr(cp).lower()
        if len(low) == 1 and ord(low) != cp and cp not in table:
            newer.add(cp)
    want = ({0xA7BA, 0xA7BC, 0xA7BE, 0xA7C2, 0xA7C4, 0xA7C5, 0xA7C6, 0xA7C7, 0xA7C9, 0xA7F5, 0x2C2F, 0xA7C0, 0xA7D0, 0xA7D6, 0xA7D8}
            | (set(range(0x10570, 0x10596)) - {0x1057B, 0x1058B, 0x10593}))
    assert newer == want

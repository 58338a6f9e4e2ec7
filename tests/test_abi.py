"""The C ABI: header <-> ctypes table <-> exported symbols agree; the library loads on a CPU-only
box (no compute call is made here)."""
import os
import re
import subprocess

import pytest

from deepctr_torch_b200 import _build, _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "ctr_b200.h")


def parse_header():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|int64_t|const char\*)\s+(ctr_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(2)] = n
    return protos


def test_header_matches_ctypes_table():
    protos = parse_header()
    table = dict((k, len(v)) for k, v in _lib.SIGNATURES.items())
    table.update((k, len(v[0])) for k, v in _lib.SPECIAL.items())
    assert set(protos) == set(table), set(protos) ^ set(table)
    for name, n in protos.items():
        assert table[name] == n, "%s: header has %d parameters, ctypes table %d" % (name, n, table[name])


def test_library_builds_and_exports_every_symbol():
    path = _build.build(verbose=False)
    assert os.path.exists(path)
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    for name in parse_header():
        assert name in exported, name
    lib = _lib.load()
    assert lib.ctr_version() == 2
    for name in _lib.exported_symbols():
        assert hasattr(lib, name)


def test_sass_is_sm100a():
    path = _build.build(verbose=False)
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout

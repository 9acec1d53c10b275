"""CPU: the shared library loads and exports exactly what include/cabinet_b200.h declares; the ctypes
prototypes in cabinet_b200/_lib.py agree with the header (argument count and kind)."""

import ctypes
import re
from pathlib import Path

import pytest

from cabinet_b200 import _lib

HEADER = Path(__file__).resolve().parent.parent / "include" / "cabinet_b200.h"


def header_prototypes():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    protos = {}
    for m in re.finditer(r"(?:const\s+char\s*\*|long\s+long|int)\s+(cabinet_\w+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        kinds = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a or "cabinet_stream_t" in a:
                    kinds.append("p")
                elif a.startswith("long long"):
                    kinds.append("ll")
                elif a.startswith("float"):
                    kinds.append("f")
                elif a.startswith("int"):
                    kinds.append("i")
                else:
                    raise AssertionError(f"unparsed argument {a!r} of {name}")
        protos[name] = kinds
    return protos


def kind(t):
    if t is ctypes.c_int:
        return "i"
    if t is ctypes.c_longlong:
        return "ll"
    if t is ctypes.c_float:
        return "f"
    return "p"


def test_library_builds_and_loads():
    if not _lib.LIB_PATH.is_file():
        _lib.build()
    lib = _lib.load()
    assert lib.cabinet_abi_version() == 2


def test_header_symbols_exported_and_prototypes_agree():
    protos = header_prototypes()
    assert len(protos) >= 15
    assert set(protos) == set(_lib.SIGNATURES), set(protos) ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name, kinds in protos.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert [kind(t) for t in _lib.SIGNATURES[name][0]] == kinds, name


def test_invalid_arguments_are_rejected_without_a_gpu():
    lib = _lib.load()
    rc = lib.cabinet_dwconv(None, 8, None, None, None, 8, _lib.BF16, 1, 8, 8, 8, 3, 1, 8, 8, 0, None, None)
    assert rc == 1 and b"null" in lib.cabinet_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "dwconv")


def test_header_constants_match_the_python_side():
    """Enum / flag values the host side hard-codes must equal the header's."""
    from cabinet_b200.engine import Engine

    text = HEADER.read_text()
    flag = re.search(r"#define\s+CABINET_CONV_REVERSE_TILES\s+(0x[0-9a-fA-F]+)", text)
    assert flag and int(flag.group(1), 16) == Engine.REVERSE_TILES
    enum = dict(re.findall(r"(CABINET_(?:ACT_\w+|F32|BF16))\s*=\s*(\d+)", text))
    assert int(enum["CABINET_ACT_NONE"]) == _lib.ACT_NONE and int(enum["CABINET_ACT_RELU"]) == _lib.ACT_RELU
    assert int(enum["CABINET_ACT_HSWISH"]) == _lib.ACT_HSWISH and int(enum["CABINET_ACT_HSIGMOID"]) == _lib.ACT_HSIGMOID
    assert int(enum["CABINET_ACT_SIGMOID"]) == _lib.ACT_SIGMOID
    assert int(enum["CABINET_F32"]) == _lib.F32 and int(enum["CABINET_BF16"]) == _lib.BF16
    assert all(int(v) < 0x100 for k, v in enum.items() if k.startswith("CABINET_ACT_"))  # the flag bit stays free

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


def test_round2_entry_points_validate_their_arguments_without_a_gpu():
    """cabinet_mbconv_t / cabinet_expand_sums / cabinet_stem_tc2: rejected shapes return CABINET_ERR_INVALID with a
    message before any CUDA call; an empty batch is a no-op (the reference's empty-loader case)."""
    import ctypes

    lib = _lib.load()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.addressof(buf) + (-ctypes.addressof(buf)) % 16   # any 16-byte aligned non-null address (never dereferenced)

    def mbt(**kw):
        a = dict(N=1, H=16, W=16, Cin=24, Cexp=72, act_e=_lib.ACT_RELU, k=3, s=1, act_dw=_lib.ACT_RELU, proj=p, Cout=24,
                 res=1, ldy=24, OH=16, OW=16, gap=None, scale=None)
        a.update(kw)
        return lib.cabinet_mbconv_t(p, a["Cin"], a["N"], a["H"], a["W"], a["Cin"], p, p, a["Cexp"], a["act_e"], a["k"], a["s"],
                                    a["act_dw"], a["proj"], p if a["proj"] else None, a["Cout"], a["res"], p, a["ldy"],
                                    a["OH"], a["OW"], a["gap"], a["scale"], None)

    assert mbt(N=0) == 0                                                   # empty batch
    assert mbt(k=4) == 1 and b"k must be" in lib.cabinet_last_error()
    assert mbt(Cin=60) == 1 and b"Cin" in lib.cabinet_last_error()         # not a multiple of 8
    assert mbt(Cin=120, Cout=120) == 1                                     # Cin % 64 > 56: no room for the bias slots
    assert mbt(act_e=_lib.ACT_SIGMOID) == 1 and b"activations" in lib.cabinet_last_error()
    assert mbt(OH=15) == 1 and b"output size" in lib.cabinet_last_error()
    assert mbt(Cout=136, ldy=136, res=0) == 1 and b"Cout" in lib.cabinet_last_error()
    assert mbt(s=2, OH=8, OW=8) == 1 and b"identity" in lib.cabinet_last_error()   # residual needs stride 1
    assert mbt(gap=p) == 1 and b"pooling sums" in lib.cabinet_last_error()         # only in the depthwise-output mode
    assert mbt(proj=None, Cout=0, res=0, ldy=72, scale=p) == 1                     # the gate input needs project mode
    assert mbt(proj=None, Cout=0, res=0, ldy=64) == 1                              # output narrower than Cexp
    assert mbt(k=5, s=2, OH=8, OW=8, res=0) == 1 and b"depthwise-output mode only" in lib.cabinet_last_error()

    def sums(**kw):
        a = dict(N=1, H=16, W=16, Cin=24, Cexp=72, k=3, nsplit=1)
        a.update(kw)
        return lib.cabinet_expand_sums(p, a["Cin"], a["N"], a["H"], a["W"], a["Cin"], p, p, a["Cexp"], _lib.ACT_RELU, a["k"],
                                       a["nsplit"], p, None)

    assert sums(N=0) == 0
    assert sums(W=300) == 1 and b"256" in lib.cabinet_last_error()         # a border column / row must fit one MMA
    assert sums(k=5, H=3) == 1                                             # image narrower than the border
    assert sums(Cexp=64) == 1                                              # the replicated-lane packing is not summed
    assert sums(k=7) == 1
    assert lib.cabinet_stem_tc2(p, 0, 64, 64, p, p, 64, p, 16, 32, 32, None) == 0
    assert lib.cabinet_stem_tc2(p, 1, 64, 62, p, p, 64, p, 16, 32, 31, None) == 1 and b"multiple of 4" in lib.cabinet_last_error()
    assert lib.cabinet_stem_tc2(p, 1, 64, 64, p, p, 64, p, 16, 31, 32, None) == 1
    assert lib.cabinet_stem_tc2(p, 1, 64, 64, None, p, 64, p, 16, 32, 32, None) == 1 and b"null" in lib.cabinet_last_error()

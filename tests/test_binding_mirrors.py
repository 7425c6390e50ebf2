"""CPU: the three statements of the C ABI's structs -- include/regnde.h (authoritative), the ctypes mirror the Python host
uses, and the Julia structs a maintainer of the reference would add (julia/RegNeuralDEB200.jl) -- must name
the same fields with the same types in the same order, and the Julia file may only ccall symbols the header declares."""
import ctypes as C
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "regnde.h").read_text()
JULIA = (ROOT / "julia" / "RegNeuralDEB200.jl").read_text()

CTYPE = {"int32_t": (C.c_int32, "Int32"), "int64_t": (C.c_int64, "Int64"), "float": (C.c_float, "Float32")}


def header_struct(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HEADER, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ty, rest = decl.split(None, 1)
        for item in rest.split(","):
            m = re.fullmatch(r"\s*(\w+)(?:\[(\d+)\])?\s*", item)
            fields.append((m.group(1), ty, int(m.group(2)) if m.group(2) else 0))
    return fields


def julia_struct(name):
    body = re.search(r"struct %s\n(.*?)\nend" % name, JULIA, re.S).group(1)
    fields = []
    for line in body.splitlines():
        line = line.split("#")[0]
        if "= new(" in line:
            continue
        for item in line.split(";"):
            item = item.strip()
            if not item:
                continue
            fname, ty = item.split("::")
            m = re.fullmatch(r"NTuple\{(\d+),(\w+)\}", ty)
            fields.append((fname, m.group(2), int(m.group(1))) if m else (fname, ty, 0))
    return fields


def check(cname, pycls, jname):
    hf = header_struct(cname)
    assert [(n, (CTYPE[t][0] * k) if k else CTYPE[t][0]) for n, t, k in hf] == \
        [(n, t) for n, t in pycls._fields_], f"ctypes mirror of {cname} differs from the header"
    assert [(n, CTYPE[t][1], k) for n, t, k in hf] == julia_struct(jname), f"Julia mirror of {cname} differs from the header"


def test_struct_mirrors_agree_with_the_header():
    from regneuralde.jl_b200 import _lib as L
    check("rnde_config", L.Config, "RndeConfig")
    check("rnde_stats", L.Stats, "RndeStats")
    check("rnde_gru_config", L.GruConfig, "RndeGruConfig")
    assert C.sizeof(L.Config) == 184 and C.sizeof(L.Stats) == 32 and C.sizeof(L.GruConfig) == 32


def test_julia_binding_only_calls_declared_symbols_and_integration_lists_all():
    declared = set(re.findall(r"\b(rnde_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)))
    called = set(re.findall(r"ccall\(\(:(\w+), LIB\)", JULIA))
    assert called and called <= declared, called - declared
    integ = (ROOT / "INTEGRATION.md").read_text()
    missing = [s for s in sorted(declared) if s not in integ]
    assert not missing, missing


def test_enum_values_agree():
    from regneuralde.jl_b200 import _lib as L
    enums = dict(re.findall(r"\b(RNDE_\w+)\s*=\s*(-?\d+)", HEADER))
    for py, c in [("ACT_IDENTITY", "RNDE_ACT_IDENTITY"), ("ACT_TANH", "RNDE_ACT_TANH"), ("ALG_TSIT5", "RNDE_ALG_TSIT5"),
                  ("ALG_AUTO_TSIT5", "RNDE_ALG_AUTO_TSIT5"), ("REG_NONE", "RNDE_REG_NONE"), ("REG_ERR_DT", "RNDE_REG_ERR_DT"),
                  ("REG_STIFF_DT_ABS", "RNDE_REG_STIFF_DT_ABS"), ("REG_STIFF_SCALED", "RNDE_REG_STIFF_SCALED"),
                  ("REG_ERR_PLUS_STIFF", "RNDE_REG_ERR_PLUS_STIFF"), ("KERNEL_AUTO", "RNDE_KERNEL_AUTO"), ("KERNEL_CHAIN", "RNDE_KERNEL_CHAIN"),
                  ("DIST_SINGLE", "RNDE_DIST_SINGLE"), ("DIST_EXACT", "RNDE_DIST_EXACT"), ("DIST_INDEPENDENT", "RNDE_DIST_INDEPENDENT"),
                  ("ARITH_FMA_CHAIN", "RNDE_ARITH_FMA_CHAIN"), ("ARITH_FIXED24", "RNDE_ARITH_FIXED24"), ("ARITH_SPLITK", "RNDE_ARITH_SPLITK"), ("OK", "RNDE_OK"),
                  ("DETACH_ALL", "RNDE_DETACH_ALL"), ("DETACH_ALL_BUT_FIRST", "RNDE_DETACH_ALL_BUT_FIRST")]:
        assert c in enums, c
        assert getattr(L, py) == int(enums[c]), (py, c)
    for name, val in re.findall(r"const (\w+)\s*=\s*Int32\((\d+)\)", JULIA):
        assert int(enums["RNDE_" + name]) == int(val), name


def test_julia_ccall_arities_match_the_header_prototypes():
    """Every ccall of the Julia stub passes as many arguments as the C prototype takes (a changed signature in
    include/regnde.h must be followed in julia/RegNeuralDEB200.jl)."""
    H = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(rnde_\w+)\s*\(([^;{]*?)\)\s*;", H, re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("void", "") else len(args.split(","))
    seen = 0
    for m in re.finditer(r"ccall\(\(:(\w+), LIB\),\s*([\w{}]+),\s*\((.*?)\),\s", JULIA, re.S):
        name, _, types = m.groups()
        depth, n = 0, 1 if types.strip() else 0
        for ch in types:
            depth += ch == "{"
            depth -= ch == "}"
            n += ch == "," and depth == 0
        if types.strip().endswith(","):
            n -= 1
        assert name in protos and n == protos[name], (name, n, protos.get(name))
        seen += 1
    assert seen >= 10


def _top_level_args(text, start):
    """arguments of the call whose opening parenthesis is at text[start]"""
    depth, args, cur = 0, [], ""
    for ch in text[start:]:
        if ch in "([{":
            depth += 1
            if depth == 1:
                continue
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                args.append(cur.strip())
                return args
        if ch == "," and depth == 1:
            args.append(cur.strip()); cur = ""
        else:
            cur += ch
    raise AssertionError("unbalanced call")


def test_julia_positional_struct_constructors_fill_every_field():
    nfields = len(header_struct("rnde_config"))
    calls = [m.end() - 1 for m in re.finditer(r"(?<!struct )RndeConfig\(", JULIA)]
    assert len(calls) >= 2
    for pos in calls:
        args = _top_level_args(JULIA, pos)
        assert len(args) == nfields, (len(args), nfields, args[:4])
        assert args[0] == "sizeof(RndeConfig)"

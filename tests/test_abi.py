"""The C-ABI boundary: libvgpu.so loads on a CPU-only box and exports every symbol include/vgpu.h
declares; without a GPU the entry points fail loudly (there is no CPU path). No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vgpu_[a-z_0-9]+)\s*\(", hdr)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("vgpu_init", "vgpu_table_create", "vgpu_segment_put", "vgpu_query_agg", "vgpu_result_get",
                 "vgpu_result_free", "vgpu_comm_init", "vgpu_last_error", "vgpu_shutdown"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    from viyadb_b200 import _native as N
    lib = C.CDLL(N.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/vgpu.h but not exported"
    bound = {s[0] for s in N.SYMBOLS}
    assert bound == set(declared_symbols()), "viyadb_b200/_native.py must bind exactly the declared symbols"
    assert lib.vgpu_abi_version() == N.VGPU_ABI_VERSION


def test_struct_layouts_match_the_header(built_lib):
    """Every public struct of include/vgpu.h against its ctypes mirror (viyadb_b200/_native.py): sizeof AND the offset of
    every field, taken from a C program compiled from the header with gcc. A binding that drifts from the header would
    hand the library misplaced pointers — silently."""
    import subprocess
    import tempfile
    from viyadb_b200 import _native as N
    pairs = [("vgpu_column", N.Column), ("vgpu_schema", N.Schema), ("vgpu_bitset_csr", N.BitsetCsr), ("vgpu_pred_node", N.PredNode),
             ("vgpu_key", N.Key), ("vgpu_plan", N.Plan), ("vgpu_result_view", N.ResultView), ("vgpu_gen_col", N.GenCol),
             ("vgpu_rows_plan", N.RowsPlan), ("vgpu_rows_view", N.RowsView), ("vgpu_search_plan", N.SearchPlan),
             ("vgpu_search_view", N.SearchView)]
    lines = []
    for cname, cls in pairs:
        lines.append(f'printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf(" {fname}:%zu", offsetof({cname}, {fname}));')
        lines.append('printf("\\n");')
    src = '#include "vgpu.h"\n#include <stddef.h>\n#include <stdio.h>\nint main(){' + "".join(lines) + "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        # a field the binding names but the header lacks fails to compile here
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == len(pairs)
    for line, (cname, cls) in zip(out, pairs):
        parts = line.split()
        assert parts[0] == cname and int(parts[1]) == C.sizeof(cls), (cname, parts[1], C.sizeof(cls))
        got = {p.split(":")[0]: int(p.split(":")[1]) for p in parts[2:]}
        want = {fname: getattr(cls, fname).offset for fname, _ in cls._fields_}
        assert got == want, (cname, got, want)
    # and the header has no struct the binding does not mirror
    import re
    declared = set(re.findall(r"typedef struct (vgpu_\w+) \{", open(os.path.join(ROOT, "include", "vgpu.h")).read()))
    assert declared == {c for c, _ in pairs}, declared ^ {c for c, _ in pairs}


def test_constants_match_the_header():
    """every enumerator and flag of include/vgpu.h against the constant of the same name (without the VGPU_ prefix) in the
    ctypes binding: an enum that drifts would send the library a different column kind or operator than the host meant"""
    import re
    import subprocess
    import tempfile
    from viyadb_b200 import _native as N
    text = open(os.path.join(ROOT, "include", "vgpu.h")).read()
    names = set()
    for body in re.findall(r"typedef enum \w+ \{(.*?)\}", text, re.S):
        names.update(re.findall(r"\b(VGPU_[A-Z0-9_]+)\s*=", body))
    names.update(re.findall(r"#define (VGPU_[A-Z0-9_]+) +[-0-9xa-fA-Fu(]", text))
    names.discard("VGPU_H_")
    assert len(names) > 50
    src = '#include "vgpu.h"\n#include <stdio.h>\nint main(){' + "".join(
        f'printf("{n} %lld\\n", (long long)({n}));' for n in sorted(names)) + "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split("\n")
    header = {l.split()[0]: int(l.split()[1]) for l in out if l.strip()}
    checked = 0
    for name, value in header.items():
        py = name[len("VGPU_"):]
        if name == "VGPU_ABI_VERSION":
            py = "VGPU_ABI_VERSION"
        if hasattr(N, py):
            assert getattr(N, py) == value or getattr(N, py) == value & 0xFFFFFFFF, (name, getattr(N, py), value)
            checked += 1
    assert checked >= 45, checked      # the binding names (nearly) everything the header defines


def test_no_cpu_fallback_without_a_device(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import viyadb_b200 as v
    with pytest.raises(v.VgpuError) as e:
        v.Database({"tables": []}, device=0)
    assert e.value.code == -3 and "no CPU path" in str(e.value)
    # plan-only databases have no device store: scanning must fail, not fall back
    db = v.Database({"tables": [{"name": "t", "dimensions": [{"name": "a"}], "metrics": [{"name": "count", "type": "count"}]}]}, device=None)
    with pytest.raises(RuntimeError):
        db.query({"type": "aggregate", "table": "t", "dimensions": ["a"], "metrics": ["count"]})


def test_product_never_imports_the_oracle():
    """Nothing under viyadb_b200/ may import, call, link or execute anything under oracle/ (comments that
    cite it are fine)."""
    bad = ("import viya_oracle", "from viya_oracle", "viya_oracle.", "oracle/_ref", "\"oracle\"", "'oracle'",
           "#include \"../../oracle", "#include \"oracle")
    for base, _, files in os.walk(os.path.join(ROOT, "viyadb_b200")):
        if "_build" in base or "__pycache__" in base:
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                continue
            for line in open(os.path.join(base, f), errors="replace"):
                code = line.split("#")[0] if f.endswith(".py") else line.split("//")[0]
                if f.endswith(".py") is False and line.lstrip().startswith("#include"):
                    code = line
                assert not any(b in code for b in bad), (os.path.join(base, f), line.strip())

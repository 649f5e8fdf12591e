"""INTEGRATION.md §3 shows the Rust binding a maintainer of the reference would add.  No Rust toolchain exists in this
image, so the transcription cannot be compiled — this test keeps it honest instead: every constant, every
#[repr(C)] struct (field names, order, types) and every extern "C" function (name, arity, parameter and return
types) in that block is compared with include/gpr.h.  A stale binding is undefined behaviour on the caller's side."""
import os
import re

import abi_parse as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRIM = {"uint32_t": "u32", "int32_t": "i32", "uint64_t": "u64", "int64_t": "i64", "double": "f64", "float": "f32",
        "int": "c_int", "size_t": "usize", "uint8_t": "u8", "char": "c_char", "void": "c_void"}


def _camel(name):
    return "".join(p.capitalize() for p in name.split("_"))


def _rust_type(base, stars, arr=None):
    const = base.startswith("const ")
    b = base[6:] if const else base
    t = PRIM.get(b) or _camel(b)
    if stars == 0:
        return f"[{t}; {arr}]" if arr else t
    for i in range(stars):
        t = ("*const " if (const and i == 0) else "*mut ") + t
    return t


def _rust_block():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```rust\n(.*?)```", doc, flags=re.S)
    return next(b for b in blocks if 'extern "C"' in b)


def test_rust_structs_match_the_header():
    rust = _rust_block()
    found = {}
    for m in re.finditer(r"#\[repr\(C\)\]\s*pub struct (\w+) \{(.*?)\}", rust, flags=re.S):
        found[m.group(1)] = [(f, " ".join(t.split())) for f, t in re.findall(r"pub (\w+):\s*([^,]+),", m.group(2))]
    structs = A.structs()
    assert len(structs) == 6
    for name, fields in structs.items():
        want = [(f, _rust_type(base, stars, arr)) for base, stars, f, arr in fields]
        assert found.get(_camel(name)) == want, name
    assert set(found) - {"GprCtx"} == {_camel(n) for n in structs}


def test_rust_functions_match_the_header():
    rust = _rust_block()
    ext = rust[rust.index('extern "C" {'):]
    ext = ext[:ext.index("\n}")]
    found = {}
    for m in re.finditer(r"pub fn (gpr_\w+)\((.*?)\)\s*(?:->\s*([^;]+))?;", ext, flags=re.S):
        params = [tuple(" ".join(x.split()) for x in p.split(":", 1)) for p in m.group(2).split(",") if p.strip()]
        found[m.group(1)] = (params, " ".join((m.group(3) or "").split()))
    funcs = A.functions()
    assert len(funcs) >= 40 and set(found) == set(funcs)
    for name, (ret, params) in funcs.items():
        want_params = [(p, _rust_type(base, stars)) for base, stars, p in params]
        want_ret = "" if ret == ("void", 0) else _rust_type(*ret)
        assert found[name] == (want_params, want_ret), name


def test_rust_constants_match_the_header():
    rust = _rust_block()
    consts = {n: int(v, 0) for n, v in re.findall(r"pub const (GPR_\w+): \w+ = (-?(?:0x)?[0-9a-fA-F]+);", rust)}
    hdr = re.sub(r"/\*.*?\*/", " ", open(A.HEADER).read(), flags=re.S)
    want = {n: int(v) for n, v in re.findall(r"^\s*(GPR_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr, flags=re.M)}
    for n, v in re.findall(r"^#define (GPR_[A-Z0-9_]+) (0x[0-9a-fA-F]+|\d+)u?\s*$", hdr, flags=re.M):
        if not n.startswith("GPR_VERSION"):
            want[n] = int(v, 0)
    assert len(want) >= 23 and consts == want

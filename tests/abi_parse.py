"""Parse include/gpr.h (plain C, regular layout) into structs and prototypes — shared by the ABI tests
(TEST INFRASTRUCTURE)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gpr.h")


def _strip_comments(src):
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def _decl(text):
    """'const float *util' -> (('const float', 1), 'util', None);  'char name[64]' -> (('char', 0), 'name', 64)"""
    text = " ".join(text.split())
    m = re.match(r"^(.*?)([A-Za-z_]\w*)(\[(\d+)\])?$", text)
    base, name, _, arr = m.groups()
    stars = base.count("*")
    base = " ".join(base.replace("*", " ").split())
    return (base, stars), name, int(arr) if arr else None


def structs():
    """{name: [(ctype, stars, field, array_len)]} in declaration order"""
    src = _strip_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} (\w+);", src, flags=re.S):
        assert m.group(1) == m.group(3)
        fields = []
        for stmt in m.group(2).split(";"):
            stmt = " ".join(stmt.split())
            if not stmt:
                continue
            first, *more = [s.strip() for s in stmt.split(",")]
            (base, stars), name, arr = _decl(first)
            fields.append((base, stars, name, arr))
            for extra in more:   # `int32_t cc_major, cc_minor;`
                fields.append((base, extra.count("*"), extra.replace("*", "").strip(), None))
        out[m.group(1)] = fields
    return out


def functions():
    """{name: (return (ctype, stars), [(ctype, stars, param)])} for every GPR_API prototype"""
    src = _strip_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"^GPR_API\s+(.*?)\b(gpr_\w+)\s*\((.*?)\)\s*;", src, flags=re.M | re.S):
        ret = " ".join(m.group(1).split())
        rstars = ret.count("*")
        ret = " ".join(ret.replace("*", " ").split())
        params = []
        body = " ".join(m.group(3).split())
        if body != "void":
            for p in body.split(","):
                (base, stars), name, _ = _decl(p.strip())
                params.append((base, stars, name))
        out[m.group(2)] = ((ret, rstars), params)
    return out

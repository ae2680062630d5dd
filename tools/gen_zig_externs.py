#!/usr/bin/env python
"""Extern declarations for bindings/zigz_b200.zig, derived from include/*.h (the same parser the ctypes loader uses).

Hand-written declarations in the .zig file win (they carry nicer pointer types); everything the headers declare and the
file does not yet mention is appended below the marker line, so the binding covers the whole C ABI.
    python tools/gen_zig_externs.py            # rewrite the generated section in place
    python tools/gen_zig_externs.py --check    # exit 1 if the file is missing a declaration
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zigz_b200 import _cabi  # noqa: E402

ZIG = os.path.join(ROOT, "bindings", "zigz_b200.zig")
MARK = "// ---- generated from include/*.h by tools/gen_zig_externs.py (do not edit below) ----"
BASE = {"int32_t": "i32", "int64_t": "i64", "uint32_t": "u32", "uint64_t": "u64", "size_t": "usize", "int": "c_int",
        "uint8_t": "u8", "zb_mle": "Mle", "zb_tree": "Tree", "float": "f32", "double": "f64", "char": "u8",
        "zb_rank_fn": "*const fn (*Ctx, i32, i32, ?*anyopaque) callconv(.C) i32"}
OPAQUE = {"zb_ctx": "Ctx", "zh_transcript": "Transcript"}
KEYWORDS = {"error", "type", "test", "var", "const", "fn", "align", "union", "struct", "enum", "opaque", "export", "extern"}


def zig_type(decl: str, is_return: bool = False):
    """(zig type, parameter name) of one C declaration."""
    d = decl.strip()
    m = re.match(r"^(.*?)([A-Za-z_]\w*)?\s*((?:\[[^\]]*\])*)\s*$", d)
    head, name, arrays = m.group(1).strip(), m.group(2), m.group(3)
    if not head:  # a bare type such as `void` or `zb_mle`
        head, name = name, None
    toks = head.replace("*", " * ").split()
    base = [t for t in toks if t not in ("const", "*", "struct", "volatile")][0]
    # pointer levels, innermost first, each with the constness of what it points to
    levels = []
    const_pending = False
    for t in toks:
        if t == "const":
            const_pending = True
        elif t == "*":
            levels.append(const_pending)
            const_pending = False
    dims = re.findall(r"\[([^\]]*)\]", arrays)
    if base == "void":
        inner = "anyopaque"
    elif base in OPAQUE:
        inner = OPAQUE[base]
    else:
        inner = BASE[base]
    if dims:  # T x[32] decays to *[32]T, T x[][32] to [*c][32]T
        const_elem = "const " if (const_pending or (toks and toks[0] == "const" and not levels)) else ""
        if len(dims) == 1:
            t = f"*{const_elem}[{dims[0]}]{inner}" if dims[0] else f"[*c]{const_elem}{inner}"
        else:
            t = f"[*c]{const_elem}[{dims[1]}]{inner}"
        return t, name
    t = inner
    for k, is_const in enumerate(levels):
        c = "const " if is_const else ""
        if k == 0 and base in OPAQUE:
            t = f"*{c}{t}"
        elif k == 0 and base == "void":
            t = f"?*{c}anyopaque"
        elif k == 0 and base == "char":
            t = "[*:0]const u8" if is_const else "[*c]u8"
        else:
            t = f"[*c]{c}{t}"
    if not levels and base == "void":
        t = "void"
    return t, name


def extern_line(name, ret, params):
    ps = []
    for k, p in enumerate(params):
        t, pname = zig_type(p)
        pname = pname or f"a{k}"
        if pname in KEYWORDS:
            pname += "_"
        ps.append(f"{pname}: {t}")
    rt, _ = zig_type(ret, True)
    return f"pub extern fn {name}({', '.join(ps)}) {rt};"


def main():
    src = open(ZIG).read()
    head = src.split(MARK)[0].rstrip() + "\n"
    have = set(re.findall(r"pub extern fn (\w+)\(", head))
    protos = _cabi.declared_prototypes()
    missing = [(n, r, p) for n, r, p in protos if n not in have]
    gen = "\n".join(extern_line(*m) for m in missing)
    new = head + "\n" + MARK + "\n" + gen + "\n"
    if "--check" in sys.argv:
        declared = set(re.findall(r"pub extern fn (\w+)\(", src))
        lacking = [n for n, _, _ in protos if n not in declared]
        unknown = [n for n in declared if n not in {q[0] for q in protos}]
        if lacking or unknown:
            print("missing:", lacking, "unknown:", unknown)
            sys.exit(1)
        return
    open(ZIG, "w").write(new)
    print(f"{len(have)} hand-written, {len(missing)} generated")


if __name__ == "__main__":
    main()

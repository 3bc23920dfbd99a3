#!/usr/bin/env python3
"""wgsl2cpp.py — mechanical WGSL -> C++ source translator (TEST INFRASTRUCTURE, part of the oracle).

Purpose: the reference (pudnax/vokselis) is Rust + WGSL and cannot be built here (no Rust, no
Vulkan, no naga). Its *algorithm* for the raycast path, however, lives entirely in four WGSL files.
This tool reads those files WHERE THEY LIE under /root/reference/shaders/ and rewrites their syntax
token by token into C++ that compiles against oracle/wgsl_rt.hpp (vector types + WGSL builtins).
Nothing is restated by hand: expressions, evaluation order, constants and control flow are the
reference's own text. The output goes to oracle/_ref/ only (git-ignored); reference sources are
never copied into the repository.

What it handles (the WGSL subset those four files use, 2022-era syntax):
  struct decls, `type` aliases, module `let` constants, `var<private>`, resource `var`s with
  @group/@binding, `fn` with attributes, let/var statements (incl. WGSL's same-scope shadowing,
  e.g. `let tmp = tmp + ...;` -> fresh C++ names), typed and untyped vec/mat constructors,
  swizzles (multi-component -> method calls), float literals (-> `f` suffix so all arithmetic
  stays fp32), i32()/u32()/f32() conversions.

Usage: wgsl2cpp.py <in.wgsl> <out.hpp> --ns <namespace>
"""
from __future__ import annotations

import argparse
import re
import sys

TOKEN_RE = re.compile(
    r"""
    (?P<ws>\s+)
  | (?P<num>(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?f?|\d+[eE][+-]?\d+f?|0[xX][0-9a-fA-F]+[iu]?|\d+[iuf]?)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op><<=|>>=|->|<<|>>|<=|>=|==|!=|&&|\|\||\+=|-=|\*=|/=|%=|&=|\|=|\^=|\+\+|--|[-+*/%&|^~!<>=.,;:(){}\[\]@])
    """,
    re.X,
)

SCALARS = {"f32": "f", "i32": "i", "u32": "u", "bool": "b"}
TEX_TYPES = {
    "texture_storage_3d": "StorageTex3D",
    "texture_storage_2d": "StorageTex2D",
    "texture_3d": "Tex3D",
    "texture_2d": "Tex2D",
}
SWZ = re.compile(r"^(?:[xyzw]{2,4}|[rgba]{2,4})$")
# WGSL builtin functions are renamed w_<name> so they can never collide with <cmath>'s globals.
BUILTINS = {
    "min", "max", "clamp", "smoothstep", "mix", "pow", "abs", "floor", "fract", "ceil", "sin", "cos", "sqrt",
    "length", "normalize", "dot", "cross", "any", "all", "exp", "exp2", "log", "log2", "sign", "step", "round",
    "trunc", "tan", "inverseSqrt", "distance", "reflect",
}


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def tokenize(src: str):
    out, pos = [], 0
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {src[pos:pos + 40]!r}")
        kind = m.lastgroup
        if kind != "ws":
            out.append((kind, m.group()))
        pos = m.end()
    return out


class Tokens:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        j = self.i + k
        return self.t[j] if j < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, text):
        if self.peek()[1] == text:
            self.i += 1
            return True
        return False

    def expect(self, text):
        tok = self.next()
        if tok[1] != text:
            ctx = " ".join(t[1] for t in self.t[max(0, self.i - 8): self.i + 4])
            raise SyntaxError(f"expected {text!r}, got {tok[1]!r} near: {ctx}")
        return tok

    def eof(self):
        return self.i >= len(self.t)


def skip_attributes(ts: Tokens):
    """@name or @name(args...)"""
    attrs = []
    while ts.peek()[1] == "@":
        ts.next()
        name = ts.next()[1]
        args = []
        if ts.peek()[1] == "(":
            depth = 0
            while True:
                tok = ts.next()[1]
                if tok == "(":
                    depth += 1
                elif tok == ")":
                    depth -= 1
                    if depth == 0:
                        break
                else:
                    args.append(tok)
        attrs.append((name, args))
    return attrs


def parse_type(ts: Tokens) -> str:
    """Consume a WGSL type, return the C++ spelling."""
    name = ts.next()[1]
    if ts.peek()[1] == "<":
        ts.next()
        args, depth, cur = [], 1, []
        while depth:
            tok = ts.next()[1]
            if tok == "<":
                depth += 1
            elif tok == ">":
                depth -= 1
                if depth == 0:
                    break
            if tok == "," and depth == 1:
                args.append("".join(cur)); cur = []
            else:
                cur.append(tok)
        args.append("".join(cur))
        m = re.fullmatch(r"vec([234])", name)
        if m:
            return f"vec{m.group(1)}{SCALARS[args[0]]}"
        m = re.fullmatch(r"mat([234])x([234])", name)
        if m and m.group(1) == m.group(2):
            return f"mat{m.group(1)}{SCALARS[args[0]]}"
        if name in TEX_TYPES:
            return TEX_TYPES[name]
        raise SyntaxError(f"unsupported generic type {name}<{args}>")
    if name == "sampler":
        return "Sampler"
    return name  # f32/i32/u32 (aliases in wgsl_rt.hpp), struct names, `type` aliases


class Scope:
    """WGSL allows re-declaring a name in the scope that already holds it; C++ does not."""

    def __init__(self):
        self.stack = [{}]
        self.counter = {}

    def push(self):
        self.stack.append({})

    def pop(self):
        self.stack.pop()

    def lookup(self, name):
        for s in reversed(self.stack):
            if name in s:
                return s[name]
        return name

    def declare(self, name):
        cur = self.stack[-1]
        if name in cur:
            k = self.counter.get(name, 0) + 1
            self.counter[name] = k
            cur[name] = f"{name}_{k}"
        else:
            cur[name] = name
        return cur[name]


def float_lit(tok: str) -> str:
    if re.fullmatch(r"(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+", tok):
        return tok + "f"
    if tok.endswith("i") and tok[:-1].isdigit():
        return tok[:-1]
    return tok


def emit_expr_tokens(ts: Tokens, scope: Scope, stop: set[str], out: list[str]):
    """Copy an expression up to (not including) a depth-0 token in `stop`, applying the rewrites."""
    depth = 0
    prev = ""
    while True:
        kind, tok = ts.peek()
        if kind == "eof":
            raise SyntaxError("unexpected eof in expression")
        if depth == 0 and tok in stop:
            return
        ts.next()
        if tok in "([{":
            depth += 1
        elif tok in ")]}":
            depth -= 1
        if kind == "num":
            out.append(float_lit(tok))
        elif kind == "id":
            if prev == ".":
                # member access or swizzle
                if SWZ.match(tok):
                    out.append(tok + "()")
                else:
                    out.append(tok)
            elif re.fullmatch(r"vec[234]|mat[234]x[234]", tok):
                if ts.peek()[1] == "<":
                    ts.i -= 1
                    out.append(parse_type(ts))
                else:  # untyped constructor: every use in these shaders is f32
                    m = re.fullmatch(r"mat([234])x\1", tok)
                    out.append((f"mat{m.group(1)}f") if m else tok + "f")
            elif tok in ("i32", "u32", "f32") and ts.peek()[1] == "(":
                out.append(f"to_{tok}")
            elif tok in BUILTINS and ts.peek()[1] == "(" and scope.lookup(tok) == tok:
                out.append("w_" + tok)
            else:
                out.append(scope.lookup(tok))
        else:
            out.append(tok)
        prev = tok


def join(parts: list[str]) -> str:
    s = ""
    for p in parts:
        if s and (s[-1].isalnum() or s[-1] == "_") and (p[0].isalnum() or p[0] == "_"):
            s += " "
        elif s and s[-1] in ",;" :
            s += " "
        s += p
    return s


def emit_block(ts: Tokens, scope: Scope, out: list[str], indent: int):
    """Translate statements until the matching '}' (consumed)."""
    pad = "    " * indent
    while True:
        kind, tok = ts.peek()
        if tok == "}":
            ts.next()
            return
        if tok == "{":
            ts.next()
            out.append(pad + "{")
            scope.push()
            emit_block(ts, scope, out, indent + 1)
            scope.pop()
            out.append(pad + "}")
            continue
        if tok in ("let", "var"):
            out.append(pad + emit_decl(ts, scope) + ";")
            ts.expect(";")
            continue
        if tok == "if":
            ts.next()
            parts = ["if ("]
            emit_expr_tokens(ts, scope, {"{"}, parts)
            parts.append(") {")
            ts.expect("{")
            out.append(pad + join(parts))
            scope.push()
            emit_block(ts, scope, out, indent + 1)
            scope.pop()
            while ts.peek()[1] == "else":
                ts.next()
                if ts.peek()[1] == "if":
                    ts.next()
                    parts = ["} else if ("]
                    emit_expr_tokens(ts, scope, {"{"}, parts)
                    parts.append(") {")
                else:
                    parts = ["} else {"]
                ts.expect("{")
                out.append(pad + join(parts))
                scope.push()
                emit_block(ts, scope, out, indent + 1)
                scope.pop()
            out.append(pad + "}")
            continue
        if tok == "for":
            ts.next()
            ts.expect("(")
            scope.push()
            init = emit_decl(ts, scope) if ts.peek()[1] in ("let", "var") else join(expr_until(ts, scope, {";"}))
            ts.expect(";")
            cond = join(expr_until(ts, scope, {";"}))
            ts.expect(";")
            step = join(expr_until(ts, scope, {")"}))
            ts.expect(")")
            ts.expect("{")
            out.append(f"{pad}for ({init}; {cond}; {step}) {{")
            scope.push()
            emit_block(ts, scope, out, indent + 1)
            scope.pop()
            scope.pop()
            out.append(pad + "}")
            continue
        # plain statement: return / break / assignment / call
        parts = expr_until(ts, scope, {";"})
        ts.expect(";")
        out.append(pad + join(parts) + ";")


def expr_until(ts: Tokens, scope: Scope, stop: set[str]) -> list[str]:
    parts: list[str] = []
    emit_expr_tokens(ts, scope, stop, parts)
    return parts


def emit_decl(ts: Tokens, scope: Scope) -> str:
    """`let|var name [: T] [= expr]` (terminator not consumed)."""
    kw = ts.next()[1]
    if ts.peek()[1] == "<":  # var<function> etc.
        while ts.next()[1] != ">":
            pass
    name = ts.next()[1]
    ctype = None
    if ts.accept(":"):
        ctype = parse_type(ts)
    init = None
    if ts.accept("="):
        init = join(expr_until(ts, scope, {";"}))  # evaluated BEFORE the new name is visible
    cname = scope.declare(name)
    const = "const " if kw == "let" else ""
    if init is None:
        return f"{ctype} {cname}{{}}"
    if ctype is None:
        return f"{const}auto {cname} = {init}"
    return f"{const}{ctype} {cname} = {init}"


def translate(src: str, ns: str, origin: str) -> str:
    ts = Tokens(tokenize(strip_comments(src)))
    out = [
        f"// GENERATED by oracle/wgsl2cpp.py from {origin} — do not edit, do not commit.",
        "#pragma once",
        '#include "wgsl_rt.hpp"',
        f"namespace {ns} {{",
        "using namespace wgsl;",
    ]
    entry_points = []
    while not ts.eof():
        attrs = skip_attributes(ts)
        kind, tok = ts.peek()
        if tok == ";":
            ts.next()
            continue
        if tok == "struct":
            ts.next()
            name = ts.next()[1]
            ts.expect("{")
            out.append(f"struct {name} {{")
            while not ts.accept("}"):
                skip_attributes(ts)
                fname = ts.next()[1]
                ts.expect(":")
                ftype = parse_type(ts)
                ts.accept(",")
                ts.accept(";")
                out.append(f"    {ftype} {fname}{{}};")
            ts.accept(";")
            out.append("};")
        elif tok == "type":
            ts.next()
            name = ts.next()[1]
            ts.expect("=")
            out.append(f"using {name} = {parse_type(ts)};")
            ts.expect(";")
        elif tok == "let":
            scope = Scope()
            out.append("static " + emit_decl(ts, scope) + ";")
            ts.expect(";")
        elif tok == "var":
            ts.next()
            space = None
            if ts.accept("<"):
                space = ts.next()[1]
                while ts.next()[1] != ">":
                    pass
            name = ts.next()[1]
            ts.expect(":")
            ctype = parse_type(ts)
            init = None
            if ts.accept("="):
                init = join(expr_until(ts, Scope(), {";"}))
            ts.expect(";")
            if space == "private":
                out.append(f"static thread_local {ctype} {name}" + (f" = {init};" if init else "{};"))
            else:  # uniform / storage / handle: bound by the driver before each dispatch
                out.append(f"static {ctype} {name}{{}};  // resource binding {attrs}")
        elif tok == "fn":
            ts.next()
            name = ts.next()[1]
            ts.expect("(")
            scope = Scope()
            params = []
            while not ts.accept(")"):
                skip_attributes(ts)
                pname = ts.next()[1]
                ts.expect(":")
                ptype = parse_type(ts)
                ts.accept(",")
                params.append(f"{ptype} {scope.declare(pname)}")
            ret = "void"
            if ts.accept("->"):
                skip_attributes(ts)
                ret = parse_type(ts)
            ts.expect("{")
            out.append(f"static inline {ret} {name}({', '.join(params)}) {{")
            emit_block(ts, scope, out, 1)
            out.append("}")
            stage = [a for a in attrs if a[0] in ("compute", "vertex", "fragment")]
            if stage:
                wg = [a[1] for a in attrs if a[0] == "workgroup_size"]
                entry_points.append((name, stage[0][0], wg[0] if wg else []))
        else:
            raise SyntaxError(f"unexpected top-level token {tok!r}")
    for name, stage, wg in entry_points:
        dims = [w for w in wg if w != ","]
        if dims:
            out.append(f"static const unsigned {name}_workgroup_size[3] = {{{', '.join(dims)}}};")
    out.append(f"}}  // namespace {ns}")
    return "\n".join(out) + "\n"


def emit_swizzles(dirname: str):
    """swz{2,3,4}.inc: every 2/3/4-component swizzle method for wgsl_rt.hpp's vecN<T>."""
    import itertools
    import os

    for n in (2, 3, 4):
        lines = []
        for names in ("xyzw"[:n], "rgba"[:n]):
            for length in (2, 3, 4):
                for combo in itertools.product(names, repeat=length):
                    lines.append(f"    vec{length}<T> {''.join(combo)}() const {{ return {{{', '.join(combo)}}}; }}")
        with open(os.path.join(dirname, f"swz{n}.inc"), "w") as f:
            f.write("\n".join(lines) + "\n")


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--emit-swizzles":
        emit_swizzles(sys.argv[2])
        return 0
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--ns", required=True)
    a = ap.parse_args()
    with open(a.src) as f:
        text = f.read()
    cpp = translate(text, a.ns, a.src)
    with open(a.dst, "w") as f:
        f.write(cpp)
    return 0


if __name__ == "__main__":
    sys.exit(main())

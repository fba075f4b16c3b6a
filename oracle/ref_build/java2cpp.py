#!/usr/bin/env python3
"""java2cpp.py — stdin -> stdout filter used by build_ref.sh.  TEST INFRASTRUCTURE.

Purely SYNTACTIC rewrites that let g++ compile excerpts of the reference's Java sources behind
java_shim.hpp.  No arithmetic, control flow, constant or call is touched:
  package line, `public`/`private` modifiers        dropped
  `class X {`                                       -> `struct X {`
  `static final T[][] n` / `T[] n` / `T n`          -> `static inline const JArr<..> n` / `static constexpr T n`
  array types  `T[] x`, `T[][] x`, `T[][][] x`      -> JArr<T> x, JArr<JArr<T> > x, ...
  `new T[]{`                                        -> `JArr<T>{`
  `new T[expr]` followed by k empty `[]`            -> JArr<..k+1 deep..>(expr)   (bracket matched)
  `this.`                                           -> `this->`
  `System.out.println(...);`                        -> `;`   (string concatenation has no C++ spelling)
"""
import re
import sys

PRIMS = ("float", "double", "int")


def jarr(t: str, depth: int) -> str:
    s = t
    for _ in range(depth):
        s = f"JArr<{s} >"
    return s


def rewrite_new(line: str) -> str:
    out = []
    i = 0
    while True:
        m = re.compile(r"new (float|double|int)\[").search(line, i)
        if not m:
            out.append(line[i:])
            break
        out.append(line[i:m.start()])
        t = m.group(1)
        j = m.end()
        if line[j] == "]":  # new T[]{...}
            k = j + 1
            depth = 1
            while line.startswith("[]", k):
                depth += 1
                k += 2
            out.append(jarr(t, depth))
            i = k
            continue
        level = 1
        k = j
        while level:
            c = line[k]
            level += (c == "[") - (c == "]")
            k += 1
        expr = line[j:k - 1]
        depth = 1
        while line.startswith("[]", k):
            depth += 1
            k += 2
        out.append(f"{jarr(t, depth)}({expr})")
        i = k
    return "".join(out)


def convert(line: str) -> str:
    if line.startswith("package "):
        return "\n"
    line = re.sub(r"\b(public|private) ", "", line)
    line = re.sub(r"\bclass ([A-Za-z]+) \{", r"struct \1 {", line)
    line = rewrite_new(line)
    for t in PRIMS:
        for depth in (3, 2, 1):
            line = re.sub(r"static final %s%s ?" % (t, r"\[\]" * depth), "static inline const %s " % jarr(t, depth), line)
        line = re.sub(r"static final %s " % t, "static constexpr %s " % t, line)
        for depth in (3, 2, 1):
            line = re.sub(r"\b%s%s " % (t, r"\[\]" * depth), "%s " % jarr(t, depth), line)
    line = line.replace("this.", "this->")
    line = re.sub(r"System\.out\.println\(.*\);", ";", line)
    return line


if __name__ == "__main__":
    for ln in sys.stdin:
        sys.stdout.write(convert(ln))

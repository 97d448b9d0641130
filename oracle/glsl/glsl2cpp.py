#!/usr/bin/env python3
"""glsl2cpp.py -- TEST INFRASTRUCTURE. Makes the reference's own GLSL compile as C++ (against the vendored glm and
oracle/glsl/glsl_shim.hpp), so that its shaders - not a restatement of them - produce the golden vectors for the
encoder, the scene gather, the dst codec, the loss gradients, the backward pass, the dW reduction and the optimizer.

    glsl2cpp.py <reference root> <out dir> <relative file> [...]

The sources are read where they lie under the reference tree; the output goes to <out dir> (oracle/_ref_glsl/gen, which
is git-ignored): nothing of the reference is copied into this repository. The translation is purely lexical and leaves
every expression, constant and statement of the shader as written. The complete rule list:

  R1  `#version` / `#extension` lines are dropped (no C++ meaning).
  R2  floating-point literals get an `f` suffix: an unsuffixed GLSL literal is a 32-bit float, an unsuffixed C++ one is a
      double. (`1e-8` -> `1e-8f`, `2.0` -> `2.0f`; integer literals are untouched.)
  R3  parameter qualifiers: `in` is dropped; `out T x` / `inout T x` become `T &x`; for array parameters the qualifier is
      simply dropped (a C++ array parameter already aliases the caller's array, which is GLSL's copy-in/copy-out result
      for the non-aliased calls these shaders make).
  R4  interface blocks and opaque uniforms become plain globals the runtime binds:
        layout(...) [readonly|writeonly] buffer B { T a[]; };          ->  static T *a;
        layout(...) uniform|buffer B { T x; U y, z; };                  ->  static T x; static U y, z;
        layout(...) buffer B { T a[]; } inst[N];                        ->  static struct { T *a; } inst[N];
        layout(...) [readonly|writeonly] uniform image2D|sampler2D n;   ->  static image2D|sampler2D n;
        layout(constant_id = k) const T n = v;                          ->  static const T n = v;
        layout(local_size_x...) in;                                     ->  dropped (the runtime dispatches 128 / 64 / 1 lanes)
  R5  `shared` -> `static thread_local` (one workgroup runs on one OS thread at a time).
  R6  array constructors: `T x[N] = T[N](a, b, ...)` -> `T x[N] = {a, b, ...}`; as an expression `T[N](a, ...)` ->
      `glsl_array<T, N>{{a, ...}}` (converts to `const T *`).
  R7  per-file compatibility patches, each a plain string replacement listed in PATCHES below with its justification
      (GLSL constructor conversions C++ / glm do not have).
"""
import os
import re
import sys

PATCHES = {
    # ivec2(uint, uvec3): GLSL constructors consume components left to right and may leave the rest of the last argument
    # unused (GLSL 4.60 spec 5.4.2), i.e. only .x of the uvec3 quotient is used. glm has no such constructor.
    "test/mlp_learning_an_image/inference.comp": [("gl_GlobalInvocationID / WINDOW_SIZE", "gl_GlobalInvocationID.x / WINDOW_SIZE")],
    # test/train_NV.comp is stale against its own NN_nv.glsl: it calls NNLoadDA3_L2Loss with 3 arguments, the
    # function takes 4 (loss_scale). train_32.spv predates the parameter; LOSS_SCALE is 1.0 everywhere (Constant.glsl:8).
    "test/train_NV.comp": [("NNLoadDA3_L2Loss(predict, vec3(f16vec3(target.x, target.y, target.z)), out_coopmats)",
                            "NNLoadDA3_L2Loss(predict, vec3(f16vec3(target.x, target.y, target.z)), 1.0f, out_coopmats)")],
}

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")


def strip_comments_keep(text):
    """Returns a list of (is_code, chunk): comments are passed through untouched."""
    out, i, n = [], 0, len(text)
    start = 0
    while i < n:
        if text.startswith("//", i):
            j = text.find("\n", i)
            j = n if j < 0 else j
            out.append((True, text[start:i])), out.append((False, text[i:j]))
            i = start = j
        elif text.startswith("/*", i):
            j = text.find("*/", i)
            j = n if j < 0 else j + 2
            out.append((True, text[start:i])), out.append((False, text[i:j]))
            i = start = j
        else:
            i += 1
    out.append((True, text[start:]))
    return out


def match_paren(s, i, open_ch="(", close_ch=")"):
    depth = 0
    for j in range(i, len(s)):
        if s[j] == open_ch:
            depth += 1
        elif s[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced " + open_ch)


def rewrite_block_members(body, as_struct):
    """members of an interface block -> declarations (unsized arrays become pointers)."""
    decls = []
    for m in [x.strip() for x in body.split(";") if x.strip()]:
        mm = re.match(r"^(.*?)(\w+)\s*\[\s*\]$", m, re.S)
        if mm:
            decls.append(f"{mm.group(1).strip()} *{mm.group(2)};")
        else:
            decls.append(m + ";")
    return " ".join(decls) if as_struct else " ".join("static " + d for d in decls)


def rewrite_layouts(code):
    out, i = [], 0
    pat = re.compile(r"\blayout\s*\(")
    while True:
        m = pat.search(code, i)
        if not m:
            out.append(code[i:])
            break
        out.append(code[i:m.start()])
        close = match_paren(code, m.end() - 1)
        rest = code[close + 1:]
        semi = None
        mq = re.match(r"\s*((?:(?:readonly|writeonly|coherent|volatile|restrict)\s+)*)(uniform|buffer|in|const)\b", rest)
        if not mq:
            raise ValueError("unhandled layout declaration: " + code[m.start():close + 40])
        kind = mq.group(2)
        after = rest[mq.end():]
        if kind == "in":  # layout(local_size_x = ...) in;
            semi = after.index(";")
            out.append("/* local size: set by the runtime */")
        elif kind == "const":  # specialisation constant
            semi = after.index(";")
            out.append("static const" + after[:semi + 1])
        else:
            mb = re.match(r"\s*(\w+)\s*\{", after)
            if mb and mb.group(1) not in ("image2D", "sampler2D"):
                bclose = match_paren(after, mb.end() - 1, "{", "}")
                body = after[mb.end():bclose]
                tail = after[bclose + 1:]
                semi_rel = tail.index(";")
                inst = tail[:semi_rel].strip()
                if inst:
                    out.append("static struct { " + rewrite_block_members(body, True) + " } " + inst + ";")
                else:
                    out.append(rewrite_block_members(body, False))
                semi = bclose + 1 + semi_rel
            else:  # opaque uniform
                semi = after.index(";")
                out.append("static" + after[:semi + 1])
        i = close + 1 + mq.end() + semi + 1
    return "".join(out)


def rewrite_qualifiers(code):
    # out / inout parameters
    def fix(m):
        typ, name, nxt = m.group(2), m.group(3), m.group(4)
        return f"{typ} {name}{nxt}" if nxt.startswith("[") else f"{typ} &{name}{nxt}"
    code = re.sub(r"\b(out|inout)\s+(\w+(?:\s*<[^<>]*>)?)\s+(\w+)(\s*[\[,)])", lambda m: fix(m).replace("  ", " "), code)
    # `in` as a parameter qualifier (always followed by a type or `const`)
    code = re.sub(r"\bin\s+(?=const\b|\w)", "", code)
    return code


def rewrite_array_ctors(code):
    # declaration form: T x[N] = T[N](...)
    pat = re.compile(r"=\s*(\w+)\s*\[\s*(\w+)\s*\]\s*\(")
    while True:
        m = pat.search(code)
        if not m:
            break
        close = match_paren(code, m.end() - 1)
        code = code[:m.start()] + "= {" + code[m.end():close] + "}" + code[close + 1:]
    pat = re.compile(r"(?<![\w\]])(u?i?vec[234]|float|uint|int)\s*\[\s*(\w+)\s*\]\s*\(")
    while True:
        m = pat.search(code)
        if not m:
            break
        close = match_paren(code, m.end() - 1)
        code = code[:m.start()] + f"glsl_array<{m.group(1)}, {m.group(2)}>{{{{" + code[m.end():close] + "}}" + code[close + 1:]
    return code


def translate(text, rel):
    for a, b in PATCHES.get(rel, []):
        assert a in text, f"{rel}: compatibility patch no longer applies: {a}"
    lines = [l for l in text.split("\n") if not re.match(r"\s*#\s*(version|extension)\b", l)]
    text = "\n".join(lines)
    chunks = strip_comments_keep(text)
    code = "\x00".join(c for is_code, c in chunks if is_code)  # comments cut out, positions remembered by the separators
    code = FLOAT_LIT.sub(lambda m: m.group(1) + "f", code)
    for a, b in PATCHES.get(rel, []):
        a2 = FLOAT_LIT.sub(lambda m: m.group(1) + "f", a)
        assert a2 in code, f"{rel}: compatibility patch does not match after R2: {a2}"
        code = code.replace(a2, b)
    code = rewrite_layouts(code)
    code = rewrite_qualifiers(code)
    code = rewrite_array_ctors(code)
    code = re.sub(r"\bshared\b", "static thread_local", code)
    parts = code.split("\x00")
    res, k = [], 0
    for is_code, c in chunks:
        if is_code:
            res.append(parts[k])
            k += 1
        else:
            res.append(c)
    return "".join(res)


def main():
    root, out = sys.argv[1], sys.argv[2]
    for rel in sys.argv[3:]:
        src = open(os.path.join(root, rel)).read()
        dst = os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as f:
            f.write(f"// GENERATED by oracle/glsl/glsl2cpp.py from {rel} of the reference tree -- do not commit\n")
            f.write(translate(src, rel))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle/_ref recipe) -- not product code.

Turns the reference's GLSL shader sources, read where they lie under /root/reference/resources/shaders, into C++ that
compiles against the reference's vendored GLM, so that the reference's OWN shading / post-pass code can run on the CPU and
pin the oracle restatement.  Output goes to a scratch directory given on the command line (never into the repository);
only the compiled oracle/_ref/libref_shaders.so is kept (git-ignored).

The rewriting is purely lexical and changes no arithmetic:
  * `#version` / `#extension` lines and `layout(...)` qualifiers are dropped;
  * interface blocks (`uniform X {..} x;`, `buffer X { T v[]; } x;`) become structs (unsized arrays become pointers);
  * `restrict`, `uniform` in front of opaque types are dropped; `inout T x` becomes `T& x`;
  * float literals get an `f` suffix (GLSL literals are 32-bit floats; C++ ones would be doubles);
  * multi-component swizzles `.xyz` become GLM's function form `.xyz()`.
"""
import os
import re
import sys

SRC = "/root/reference/resources/shaders"
FILES = ["raygen.rgen", "raygen.h", "closesthit.rchit", "miss.rmiss", "shadowMiss.rmiss", "payload.h", "raytracer_bindings.h", "compute.h",
         "rough_prepare.comp", "rough_blur.h", "rough_blur_h.comp", "rough_blur_v.comp", "postprocess.comp", "fxaa.comp", "fxaa.h",
         "gpu_material.def", "vertex.def", "uniform_buffer_object.def", "instance_offset_table.def", "compute_shader_shared.def"]

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")
SWIZZLE = re.compile(r"\.([xyzw]{2,4}|[rgba]{2,4})\b(?!\s*\()")


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def convert(text):
    text = strip_comments(text)
    out = []
    for line in text.split("\n"):
        s = line.strip()
        if s.startswith("#version") or s.startswith("#extension"):
            continue
        if re.match(r"^\s*layout\s*\(\s*local_size", line):
            continue
        if s.startswith("#include") or s.startswith("#pragma"):
            out.append(line)
            continue
        line = re.sub(r"layout\s*\([^)]*\)", "", line)
        line = re.sub(r"\brestrict\b", "", line)
        line = re.sub(r"\buniform\s+(\w+)\s*\{", r"struct \1 {", line)        # uniform block
        line = re.sub(r"\bbuffer\s+(\w+)", r"struct \1", line)                  # storage block
        line = re.sub(r"\buniform\s+", "", line)                               # opaque uniforms (images, samplers, AS)
        line = re.sub(r"\b(\w+)\s+(\w+)\[\];", r"\1* \2;", line)               # unsized array member
        line = re.sub(r"\binout\s+(\w+)\s+(\w+)", r"\1& \2", line)
        if not s.startswith("#if") and not s.startswith("#elif"):
            line = FLOAT_LIT.sub(lambda m: m.group(1) + "f", line)
            line = SWIZZLE.sub(lambda m: "." + m.group(1) + "()", line)
        out.append(line)
    return "\n".join(out) + "\n"


def main():
    dst = sys.argv[1]
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        with open(os.path.join(SRC, f)) as fh:
            text = fh.read()
        with open(os.path.join(dst, f), "w") as fh:
            fh.write(convert(text))
    print("generated", len(FILES), "files into", dst)


if __name__ == "__main__":
    main()

"""Builds oracle/_ref/libref_glsl.so: the REFERENCE's own GLSL shaders, compiled for the CPU.  TEST INFRASTRUCTURE ONLY.

The shader sources are read where they lie under /root/reference/Core/src/Shaders at build time; nothing of them is copied into
this repository (the translation units are written to a temporary directory and deleted, only the .so lands in oracle/_ref/,
which is git-ignored).  Each shader becomes `namespace glsl { namespace shader_<name> { <shader text> <our driver> } }` on top
of oracle/glsl_cpu.h (GLSL types and built-ins as C++).  The ONLY textual changes made to the shader text, all mechanical:
  * `#include "x.glsl"` is expanded (what the reference's own loader, pangolin::GlSlProgram, does) and `#version` dropped;
  * storage qualifiers at global scope (`uniform`, `in`, `out`, `flat`, `layout(...)`) are dropped: the variables become
    namespace-scope C++ variables that the driver sets / reads;
  * parameter qualifiers: `in` dropped; `out` / `inout` on a non-array parameter becomes a C++ reference (`T& name`); on an array
    parameter it is dropped (a C++ array parameter already aliases the caller's array).
Compiled with -fsingle-precision-constant (GLSL literals are fp32) and -ffp-contract=off.

    python oracle/build_ref_glsl.py            # no-op when /root/reference is absent or the library is up to date
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REF_SHADERS", "/root/reference/Core/src/Shaders")
OUT = os.path.join(HERE, "_ref", "libref_glsl.so")
# name -> shader file; the driver is oracle/glsl_drivers/<name>.inc
SHADERS = {
    "predict_hrbf": "predict_hrbf.frag",
    "depth_bilateral": "depth_bilateral.frag",
    "depth_metric_raw": "depth_metric_raw.frag",
    "depth_metric_filtered": "depth_metric_filtered.frag",
    "depth_vertex_normal_radius": "depth_vertex_normal_radius.frag",
    "depth_curvature_gradient": "depth_curvature_gradient.frag",
    "depth_confidence_evaluation": "depth_confidence_evaluation.frag",
    "fill_vertex": "fill_vertex.frag",
    "fill_normal": "fill_normal.frag",
    "fill_curvature": "fill_curvature.frag",
    "fill_rgb": "fill_rgb.frag",
    "index_map": "index_map.vert",
    "data_vert": "data.vert",
    "update_vert": "update.vert",
    "copy_unstable": "copy_unstable.vert",
    "init_unstable": "init_unstableTex.vert",
    "resize": "resize.frag",
    "update_delta_trans": "update_delta_trans.vert",
    "depth_update_normalrad": "depth_update_normalrad.frag",
}
TYPES = r"(?:float|int|uint|bool|vec[234]|mat[34]|sampler2D|usampler2D)"


def expand_includes(path, seen):
    out = []
    for line in open(path, encoding="utf-8", errors="replace"):
        m = re.match(r'\s*#include\s*"([^"]+)"', line)
        if m:
            inc = os.path.join(os.path.dirname(path), m.group(1))
            if inc not in seen:
                seen.add(inc)
                out.append(expand_includes(inc, seen))
            continue
        if re.match(r"\s*#version", line):
            continue
        out.append(line.rstrip("\n"))
    return "\n".join(out)


def translate(src):
    # parameter qualifiers (inside parentheses: preceded by '(' or ',')
    def param(m):
        lead, qual, typ, name, arr = m.group(1), m.group(2), m.group(3), m.group(4), m.group(5)
        if arr or qual == "in":
            return f"{lead}{typ} {name}{arr}"
        return f"{lead}{typ}& {name}"
    src = re.sub(r"([(,]\s*)(inout|out|in)\s+(" + TYPES + r")\s+(\w+)(\s*\[)?", lambda m: param(m), src)
    # storage qualifiers at global scope (start of a line)
    src = re.sub(r"(?m)^\s*(?:layout\s*\([^)]*\)\s*)?(?:flat\s+)?(?:uniform|in|out)\s+(?=" + TYPES + r"\b)", "", src)
    return src


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, "glsl_cpu.h"), os.path.abspath(__file__)] + [os.path.join(HERE, "glsl_drivers", f) for f in os.listdir(os.path.join(HERE, "glsl_drivers"))]
    deps += [os.path.join(REF, f) for f in os.listdir(REF) if f.endswith((".glsl", ".frag", ".vert"))]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def main():
    if not os.path.isdir(REF):
        print(f"build_ref_glsl.py: {REF} not present -- keeping prebuilt {OUT}")
        return 0
    if not stale():
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        tus = []
        for name, fname in SHADERS.items():
            body = translate(expand_includes(os.path.join(REF, fname), set()))
            driver = os.path.join(HERE, "glsl_drivers", name + ".inc")
            tu = os.path.join(tmp, name + ".cpp")
            with open(tu, "w") as f:
                prelude = '#include "vertex_stage.inc"\n' if fname.endswith(".vert") else ""
                f.write('#include "glsl_cpu.h"\nnamespace glsl { namespace shader_%s {\n%s#line 1 "%s"\n%s\n#line 1 "%s"\n#include "%s"\n} }\n'
                        % (name, prelude, fname, body, os.path.basename(driver), driver))
            tus.append(tu)
        cmd = ["g++", "-O2", "-std=c++17", "-fsingle-precision-constant", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", HERE, "-I", os.path.join(HERE, "glsl_drivers"), "-o", OUT] + tus
        subprocess.check_call(cmd)
    print("built", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())

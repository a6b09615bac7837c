#!/usr/bin/env python
"""Builds oracle/_ref/libref_odometry.so: the REFERENCE's own tracking loop -- Core/src/Utils/RGBDOdometry.cpp: constructor, the
init* functions, populateRGBDData and getIncrementalTransformation, VERBATIM -- on top of the reference's own kernels
(Core/src/Cuda/reduce.cu, cudafuncs.cu, containers/device_memory.cpp, compiled unmodified as in oracle/build_ref.sh).
TEST INFRASTRUCTURE ONLY: it pins row 4 of the oracle and is the GPU "reference arm" of the tracking stage.

What is NOT the reference here, stated:
  * Eigen (not installed) is oracle/eigen_mini: same members, eager evaluation, cofactor inverses, pivoted LDL^T;
  * GPUTexture / Pangolin / the GL interop calls are served from plain cudaArrays (oracle/host_shims/odom, oracle/ref_glinterop_shim.h);
  * RGBDOdometry.cpp is TRIMMED mechanically: the member functions that are never called on the path and need PlaneExtraction /
    file output (listed in DROP) are cut out by name with brace matching; every other line of the file is compiled as it stands.
    The trimmed copy and a verbatim copy of RGBDOdometry.h are written to oracle/_ref/odom_gen/ (git-ignored), never into the repo."""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REF_CORE", "/root/reference/Core/src")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "odom_gen")
DROP = ("pseudocolor", "correspondPlaneSearch", "correspondPlaneSearchRANSAC", "getLastCorrespondence", "addToPoseGraph", "savePoseGraph",
        "getCovariance", "DownloadGPUMaps", "savefilePLY", "saveCorrepICPsave", "saveCudaAttrib")


def trim(src):
    """drop the top-level function definitions named in DROP (brace matching from the definition's first line)"""
    lines = src.split("\n")
    out, i = [], 0
    head = re.compile(r"^(?:inline\s+)?[A-Za-z_][\w:<>,\s\*&]*?\b(?:RGBDOdometry::)?(\w+)\s*\(")
    while i < len(lines):
        m = head.match(lines[i])
        if m and m.group(1) in DROP and not lines[i].startswith((" ", "\t", "//")):
            depth, seen = 0, False
            while i < len(lines):
                depth += lines[i].count("{") - lines[i].count("}")
                seen = seen or "{" in lines[i]
                i += 1
                if seen and depth == 0:
                    break
            continue
        out.append(lines[i])
        i += 1
    return "\n".join(out)


def main():
    if not os.path.isdir(REF):
        print("build_ref_odometry.py: %s not present (GPU box?) -- keeping prebuilt %s" % (REF, OUT))
        return 0
    # two builds: the reference's own nvcc flags (Core/src/CMakeLists.txt:74-75: --ftz=true --prec-div=false --prec-sqrt=false), and the
    # same sources with IEEE division / square root / denormals, which separates "the algorithm" from "the build flags" when the
    # oracle (plain C, IEEE) is compared with it
    so = os.path.join(OUT, "libref_odometry.so")
    so_ieee = os.path.join(OUT, "libref_odometry_ieee.so")
    deps = [os.path.join(REF, "Utils", "RGBDOdometry.cpp"), os.path.join(REF, "Utils", "RGBDOdometry.h"), os.path.join(REF, "Cuda", "reduce.cu"),
            os.path.join(REF, "Cuda", "cudafuncs.cu"), os.path.join(HERE, "ref_shim_odometry.cpp"), os.path.join(HERE, "eigen_mini", "Eigen", "Core"),
            os.path.join(HERE, "ref_glinterop_shim.h"), os.path.join(HERE, "ref_texshim.h"), os.path.abspath(__file__)]
    if all(os.path.exists(x) and all(os.path.getmtime(x) > os.path.getmtime(d) for d in deps) for x in (so, so_ieee)):
        return 0
    shutil.rmtree(GEN, ignore_errors=True)
    os.makedirs(os.path.join(GEN, "Utils"))
    os.symlink(os.path.join(REF, "Cuda"), os.path.join(GEN, "Cuda"))
    os.symlink(os.path.join(REF, "Defines.h"), os.path.join(GEN, "Defines.h"))
    for f in os.listdir(os.path.join(REF, "Utils")):
        if f.endswith(".h") and f != "RGBDOdometry.h":
            os.symlink(os.path.join(REF, "Utils", f), os.path.join(GEN, "Utils", f))
    shims = os.path.join(HERE, "host_shims", "odom")
    shutil.copy(os.path.join(shims, "GPUTexture.h"), os.path.join(GEN, "GPUTexture.h"))
    shutil.copy(os.path.join(shims, "PlaneExtraction.h"), os.path.join(GEN, "PlaneExtraction.h"))
    os.makedirs(os.path.join(GEN, "Shaders"))
    shutil.copy(os.path.join(shims, "Shaders", "Shaders.h"), os.path.join(GEN, "Shaders", "Shaders.h"))
    shutil.copy(os.path.join(REF, "Utils", "RGBDOdometry.h"), os.path.join(GEN, "Utils", "RGBDOdometry.h"))
    src = open(os.path.join(REF, "Utils", "RGBDOdometry.cpp"), errors="replace").read()
    trimmed = trim(src)
    for name in ("getIncrementalTransformation", "initICPModel", "populateRGBDData", "initCurvatureModel", "initICPweight", "initFirstRGB"):
        assert "RGBDOdometry::" + name in trimmed, name
    for name in DROP[1:]:
        assert "RGBDOdometry::" + name + "(" not in trimmed, name
    open(os.path.join(GEN, "Utils", "RGBDOdometry.cpp"), "w").write(trimmed)
    cuda_inc = "/usr/local/cuda/include"
    arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
    nvflags = ["--ftz=true", "--prec-div=false", "--prec-sqrt=false", "-O3", "-Xcompiler", "-fPIC", "-w", "-I" + os.path.join(REF, "Cuda")]
    gxx = ["g++", "-O2", "-std=c++17", "-fPIC", "-w", "-I" + os.path.join(HERE, "eigen_mini"), "-I" + cuda_inc, "-I" + GEN, "-I" + os.path.join(GEN, "Utils")]
    run = lambda c: subprocess.check_call(c)
    run(gxx + ["-include", os.path.join(HERE, "ref_glinterop_shim.h"), "-c", os.path.join(GEN, "Utils", "RGBDOdometry.cpp"), "-o", os.path.join(GEN, "RGBDOdometry.o")])
    run(gxx + ["-include", os.path.join(HERE, "ref_glinterop_shim.h"), "-c", os.path.join(HERE, "ref_shim_odometry.cpp"), "-o", os.path.join(GEN, "shim.o")])
    for target, flags in ((so, nvflags), (so_ieee, ["--ftz=false", "--prec-div=true", "--prec-sqrt=true"] + nvflags[3:])):
        run(["nvcc"] + arch + flags + ["-c", os.path.join(REF, "Cuda", "reduce.cu"), "-o", os.path.join(GEN, "reduce.o")])
        run(["nvcc"] + arch + flags + ["-include", os.path.join(HERE, "ref_texshim.h"), "-c", os.path.join(REF, "Cuda", "cudafuncs.cu"), "-o", os.path.join(GEN, "cudafuncs.o")])
        run(["nvcc"] + arch + flags + ["-x", "cu", "-c", os.path.join(REF, "Cuda", "containers", "device_memory.cpp"), "-o", os.path.join(GEN, "device_memory.o")])
        run(["nvcc"] + arch + ["-shared", "-o", target] + [os.path.join(GEN, o) for o in ("RGBDOdometry.o", "shim.o", "reduce.o", "cudafuncs.o", "device_memory.o")] + ["-lcudart"])
        print("built", target)
    shutil.rmtree(GEN, ignore_errors=True)      # the generated copies are build intermediates: nothing of the reference's text stays behind
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""TEST INFRASTRUCTURE ONLY.  Build oracle/_ref/scn_rules_ref.so = the reference's own CPU rule builders
(/root/reference/sparseconvnet/SCN/Metadata/*.h) compiled around oracle/rules_shim.cpp.

The reference headers cannot be #included whole (Metadata.h pulls in cudpp, sparsehash and the CUDA 9 runtime), so
the self-contained template functions are CUT OUT OF THE HEADERS AT BUILD TIME, by the markers below, into
oracle/_build/*.inc (git-ignored) -- nothing from the reference is copied into the repository.  Only runs where
/root/reference exists (the authoring container); the GPU box uses the prebuilt .so that travels with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/sparseconvnet/SCN/Metadata"
OUT = os.path.join(HERE, "_ref")
BUILD = os.path.join(HERE, "_build")
NAME = "scn_rules_ref"

# (output file, header, first-line marker [exclusive], last-line marker [exclusive])
SPANS = [
    ("rules_extracted_float3.inc", "Metadata.h", "using RuleBook = std::vector<std::vector<Int>>;",
     "template <Int dimension> using Points = std::vector<Point<dimension>>;"),
    ("rules_extracted_regions.inc", "RectangularRegions.h", "#define RECTANGULARREGIONS_H",
     "#endif /* RECTANGULARREGIONS_H */"),
    ("rules_extracted_input.inc", "IOLayersRules.h", '#include "Metadata.h"', "#ifdef GPU_GRID"),
    ("rules_extracted_submanifold.inc", "SubmanifoldConvolutionRules.h", '#include "../CUDA/SubmanifoldRules_cuda.cpp"',
     "#ifdef GPU_GRID"),
    ("rules_extracted_convolution.inc", "ConvolutionRules.h", "#include <algorithm>", "#ifdef GPU_GRID"),
]


def so_path():
    return os.path.join(OUT, NAME + ".so")


def _cut(header, first, last):
    lines = open(os.path.join(REF, header)).read().split("\n")
    a = next(i for i, l in enumerate(lines) if l.strip() == first)
    b = next(i for i, l in enumerate(lines) if i > a and l.strip() == last)
    return f"// lines {a + 2}-{b} of {header}, cut at build time\n" + "\n".join(lines[a + 1:b]) + "\n"


def build(force=False, verbose=False):
    """Returns the .so path, or None when the reference tree is absent and nothing is prebuilt."""
    target = so_path()
    src = os.path.join(HERE, "rules_shim.cpp")
    if os.path.exists(target) and not force and os.path.getmtime(target) >= os.path.getmtime(src):
        return target
    if not os.path.isdir(REF):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(BUILD, exist_ok=True)
    for out, header, first, last in SPANS:
        with open(os.path.join(BUILD, out), "w") as f:
            f.write(_cut(header, first, last))
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", f"-I{BUILD}", src, "-o", target]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("built" if p else "reference tree absent and no prebuilt .so", p)

"""TEST INFRASTRUCTURE ONLY.  Build oracle/_ref/scn_cpu_ref*.so = the reference's own CPU arithmetic
(/root/reference/sparseconvnet/SCN/CPU/*.cpp, #included unmodified by oracle/ref_shim.cpp).

Only runs where /root/reference exists (the authoring container).  The GPU box uses the prebuilt
.so that travels with the snapshot (oracle/_ref/ is git-ignored but not gpurun-ignored).
The reference's own build system (setup.py + cudpp + easy_profiler CMake) is NOT used.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/sparseconvnet/SCN"
OUT = os.path.join(HERE, "_ref")
NAME = "scn_cpu_ref"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(force=False, verbose=False):
    """Returns the .so path, or None when the reference tree is not present and nothing is prebuilt."""
    target = so_path()
    src = os.path.join(HERE, "ref_shim.cpp")
    if os.path.exists(target) and not force and os.path.getmtime(target) >= os.path.getmtime(src):
        return target
    if not os.path.isdir(REF):
        return target if os.path.exists(target) else None
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}", f"-I{REF}"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = (
        ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-w",
         f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
         f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
        + inc + [src, "-o", target, f"-L{libdir}", f"-Wl,-rpath,{libdir}",
                 "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python", "-lgomp"]
    )
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return target


def load():
    """Import the compiled reference module (None if unavailable)."""
    path = build()
    if path is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)

    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("built" if p else "reference tree absent and no prebuilt .so", p)

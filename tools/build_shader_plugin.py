"""Compiles a shader plug-in (a functor written against include/axr_shader_plugin.cuh) into a shared library that
axr_load_shader_plugin() opens at run time. Same compiler flags as the library itself (axiomr_b200/build.py): sm_100a, no FMA
contraction, static CUDA runtime.
usage: python tools/build_shader_plugin.py my_shader.cu [-o my_shader.so]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from axiomr_b200 import build as b  # noqa: E402


def build_plugin(src: str, out: str | None = None, force: bool = False) -> str:
    out = out or os.path.splitext(src)[0] + ".so"
    deps = [src] + b.DEPS + [os.path.join(ROOT, "include", "axr_shader_plugin.cuh")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps if os.path.exists(d)):
        return out
    cmd = [b.nvcc()] + b.NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "axiomr_b200", "csrc"), "-o", out, src]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True)
    except OSError as e:  # no compiler on this machine
        if os.path.exists(out) and not force:
            return out    # a copy built elsewhere travels with the tree; axr_load_shader_plugin refuses it if it is out of date
        raise RuntimeError(f"cannot run nvcc to build the shader plug-in {src}: {e}")
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed building the shader plug-in {src}")
    return out


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "-o"]
    if not args:
        raise SystemExit(__doc__)
    print(build_plugin(args[0], args[1] if len(args) > 1 else None, force=True))

"""Builds A/B variants of libkyd.so into ky_b200/lib/ab/ (compile-time switches of the kernels)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
variants = {
    "base": [],
    "traits1": ["-DKYD_TRAITS=1"],
    "traits1_inline": ["-DKYD_TRAITS=1", "-DKYD_MATH_INLINE=1"],
}
names = sys.argv[1:] or list(variants)
os.makedirs(os.path.join(g.LIB, "ab"), exist_ok=True)
procs = []
for n in names:
    out = os.path.join(g.LIB, "ab", f"libkyd_{n}.so")
    cmd = ["nvcc"] + g.NVCC_FLAGS + variants[n] + ["-shared", "-o", out, os.path.join(g.CSRC, "kyd_kernels.cu"), os.path.join(g.CSRC, "kyd_api.cu")]
    procs.append((n, subprocess.Popen(cmd, cwd=g.ROOT)))
for n, p in procs:
    print(n, "rc", p.wait())

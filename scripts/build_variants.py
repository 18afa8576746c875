"""Builds A/B variants of libkyd.so into ky_b200/lib/ab/ (compile-time switches of the kernels)."""
import itertools, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
variants = {
    "base": [],
    "noprefetch": ["-DKYD_PREFETCH=0"],
    "inline": ["-DKYD_MATH_INLINE=1"],
    "slowrsqrt": ["-DKYD_FAST_RSQRT=0"],
    "noprefetch_inline": ["-DKYD_PREFETCH=0", "-DKYD_MATH_INLINE=1"],
    "noprefetch_inline_mb3": ["-DKYD_PREFETCH=0", "-DKYD_MATH_INLINE=1", "-DKYD_SHADE_MIN_BLOCKS=3"],
    "noprefetch_inline_mb5": ["-DKYD_PREFETCH=0", "-DKYD_MATH_INLINE=1", "-DKYD_SHADE_MIN_BLOCKS=5"],
}
names = sys.argv[1:] or list(variants)
procs = []
for n in names:
    out = os.path.join(g.LIB, "ab", f"libkyd_{n}.so")
    cmd = ["nvcc"] + g.NVCC_FLAGS + variants[n] + ["-shared", "-o", out, os.path.join(g.CSRC, "kyd_kernels.cu"), os.path.join(g.CSRC, "kyd_api.cu")]
    procs.append((n, subprocess.Popen(cmd, cwd=g.ROOT)))
for n, p in procs:
    print(n, "rc", p.wait())

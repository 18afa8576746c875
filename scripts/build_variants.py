"""Builds A/B variants of libkyd.so into ky_b200/lib/ab/ (compile-time switches of the kernels)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
variants = {
    "base": [],
    "cur": [],
    "imb3": ["-DKYD_INTERSECT_MIN_BLOCKS=3"],
    "imb4": ["-DKYD_INTERSECT_MIN_BLOCKS=4"],
    "onephase": ["-DKYD_TWO_PHASE=0"],
    "unroll4": ["-DKYD_TRAVERSAL_UNROLL=4"],
    "nocull": ["-DKYD_NEE_CULL=0"],
    "powinline": ["-DKYD_POW_INLINE=1"],
    "rsqrtcall": ["-DKYD_RSQRT_NOINLINE=1"],
    "specnopf": ["-DKYD_SHADE_PREFETCH_SPECULAR=0"],
    "neenopf": ["-DKYD_NEE_PREFETCH=0"],
    "manynopf": ["-DKYD_SHADE_PREFETCH_MANY=0"],
    "nee5": ["-DKYD_NEE_MIN_BLOCKS=5"],
    "nee6": ["-DKYD_NEE_MIN_BLOCKS=6"],
    "nee8": ["-DKYD_NEE_MIN_BLOCKS=8"],
    "nodefer": ["-DKYD_NEE_DEFER=0"],
    "vmajor": ["-DKYD_NEE_LIGHT_MAJOR=0"],
    "nee7": ["-DKYD_NEE_MIN_BLOCKS=7"],
    "nosum": ["-DKYD_NEE_SUMMARY=0"],
    "lock1": ["-DKYD_SHADE_LOCKSTEP=1"],
    "smb5": ["-DKYD_SHADE_MIN_BLOCKS=5"],
    "t256": ["-DSHADE_THREADS=256", "-DKYD_SHADE_MIN_BLOCKS=2"],
    "t192": ["-DSHADE_THREADS=192", "-DKYD_SHADE_MIN_BLOCKS=2"],
    "lock7": ["-DKYD_SHADE_LOCKSTEP=7"],
    "neelock": ["-DKYD_SHADE_LOCKSTEP=3", "-DKYD_NEE_LOCKSTEP=1"],
    "lock2": ["-DKYD_SHADE_LOCKSTEP=2"],
    "lock3": ["-DKYD_SHADE_LOCKSTEP=3"],
    "sumlive": ["-DKYD_NEE_SUMMARY=2"],
    "sum5": ["-DKYD_NEE_MIN_BLOCKS=5"],
    "occ6": ["-DKYD_SHADE_MIN_BLOCKS_SPECULAR=6", "-DKYD_SHADE_MIN_BLOCKS_MANY=6"],
    "occ8": ["-DKYD_SHADE_MIN_BLOCKS_SPECULAR=8", "-DKYD_SHADE_MIN_BLOCKS_MANY=8"],
    "occ5": ["-DKYD_SHADE_MIN_BLOCKS_SPECULAR=5", "-DKYD_SHADE_MIN_BLOCKS_MANY=5"],
    "nee5d": ["-DKYD_NEE_MIN_BLOCKS=5"],
    "nopf": ["-DKYD_SHADE_PREFETCH=0"],
    "nopf_mb5": ["-DKYD_SHADE_PREFETCH=0", "-DKYD_SHADE_MIN_BLOCKS=5"],
    "nopf_mb6": ["-DKYD_SHADE_PREFETCH=0", "-DKYD_SHADE_MIN_BLOCKS=6"],
    "mb3": ["-DKYD_SHADE_MIN_BLOCKS=3"],
    "u1_mb6": ["-DKYD_SHADE_MIN_BLOCKS=6"],
    "u1_mb5": ["-DKYD_SHADE_MIN_BLOCKS=5"],
    "noinline": ["-DKYD_WF_NOINLINE=1"],
    "noinline_mb6": ["-DKYD_WF_NOINLINE=1", "-DKYD_SHADE_MIN_BLOCKS=6"],
    "imb2": ["-DKYD_INTERSECT_MIN_BLOCKS=2"],
    "mb5": ["-DKYD_SHADE_MIN_BLOCKS=5"],
    "mb6": ["-DKYD_SHADE_MIN_BLOCKS=6"],
    "slowsincos": ["-DKYD_FAST_SINCOS=0"],
    "inshadow": ["-DKYD_INLINE_SHADOW=1"],
    "inshadow_mb3": ["-DKYD_INLINE_SHADOW=1", "-DKYD_SHADE_MIN_BLOCKS=3"],
    "inshadow_pf": ["-DKYD_INLINE_SHADOW=1", "-DKYD_SHADE_PREFETCH=1"],
    "inshadow_pf_mb3": ["-DKYD_INLINE_SHADOW=1", "-DKYD_SHADE_PREFETCH=1", "-DKYD_SHADE_MIN_BLOCKS=3"],
    "pf": ["-DKYD_SHADE_PREFETCH=1"],
}
names = sys.argv[1:] or list(variants)
os.makedirs(os.path.join(g.LIB, "ab"), exist_ok=True)
procs = []
for n in names:
    out = os.path.join(g.LIB, "ab", f"libkyd_{n}.so")
    cmd = ["nvcc"] + g.NVCC_FLAGS + variants[n] + ["-shared", "-o", out, os.path.join(g.CSRC, "kyd_kernels.cu"), os.path.join(g.CSRC, "kyd_kernels_big.cu"), os.path.join(g.CSRC, "kyd_api.cu"), os.path.join(g.CSRC, "kyd_film.cu"), os.path.join(g.CSRC, "kyd_smallpt.cu"), "-Xptxas", "-v"]
    procs.append((n, subprocess.Popen(cmd, cwd=g.ROOT, stderr=open(os.path.join(g.LIB, "ab", f"{n}.ptxas.log"), "w"))))
for n, p in procs:
    print(n, "rc", p.wait())

"""Shared list of (scene, integrator, strategy) cases used by the golden generator and the parity tests."""
import ky_b200 as ky

W, H, SPP = 48, 32, 4

SCENES = {
    "cornell": (ky.SCENE_CORNELL, ky.CB_DEFAULT),
    "cornell_point": (ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_POINT),
    "cornell_direction": (ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_DIRECTION),
    "cornell_environment": (ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_ENVIRONMENT),
    "cornell_large_glass": (ky.SCENE_CORNELL, ky.CB_LARGE_GLASS | ky.CB_LIGHT_AREA),
    "cornell_large_mirror_all_lights": (ky.SCENE_CORNELL, ky.CB_LARGE_MIRROR | 15),
    "veach": (ky.SCENE_VEACH, 0),
    "smallpt": (ky.SCENE_SMALLPT, 0),
    "shapes": (ky.SCENE_SHAPES, 0),
}

STRATEGIES = {"idle": ky.DS_IDLE, "bsdf": ky.DS_BSDF, "light": ky.DS_LIGHT, "bsdf_mis": ky.DS_BSDF_MIS,
              "light_mis": ky.DS_LIGHT_MIS, "both_mis": ky.DS_BOTH_MIS}

INTEGRATORS = {"position": ky.INT_POSITION, "normal": ky.INT_NORMAL, "basecolor": ky.INT_BASECOLOR,
               "direct_lighting": ky.INT_DIRECT_LIGHTING, "simple_pt_recursion": ky.INT_SIMPLE_PT_RECURSION,
               "pt_recursion": ky.INT_PT_RECURSION, "pt_recursion_defered": ky.INT_PT_RECURSION_DEFERED,
               "pt_iteration": ky.INT_PT_ITERATION}


def film_cases():
    """(name, scene_key, integrator, direct_sample, max_depth, spp)"""
    out = []
    for sk in SCENES:
        for ik in ("position", "normal", "basecolor"):
            out.append((f"{sk}/{ik}", sk, INTEGRATORS[ik], ky.DS_IDLE, 0, 1))
        for dk, ds in STRATEGIES.items():
            out.append((f"{sk}/pt_iteration/{dk}", sk, ky.INT_PT_ITERATION, ds, 5, SPP))
        for ik in ("direct_lighting", "simple_pt_recursion", "pt_recursion", "pt_recursion_defered"):
            out.append((f"{sk}/{ik}/both_mis", sk, INTEGRATORS[ik], ky.DS_BOTH_MIS, 5, SPP))
    # BASELINE config 4 uses depth 8; config 2 compares strategies under direct lighting
    out.append(("cornell/pt_iteration/both_mis/depth8", "cornell", ky.INT_PT_ITERATION, ky.DS_BOTH_MIS, 8, SPP))
    out.append(("cornell/direct_lighting/bsdf", "cornell", ky.INT_DIRECT_LIGHTING, ky.DS_BSDF, 0, SPP))
    out.append(("cornell/direct_lighting/light", "cornell", ky.INT_DIRECT_LIGHTING, ky.DS_LIGHT, 0, SPP))
    out.append(("veach/pt_iteration/both_mis/depth1", "veach", ky.INT_PT_ITERATION, ky.DS_BOTH_MIS, 1, SPP))
    return out


def make_scene(scene_key, w=W, h=H):
    sid, flags = SCENES[scene_key]
    return ky.Scene(sid, w, h, flags)


# ---- film output stage -------------------------------------------------------------------------------------------
FILM_EDGE_VALUES = [0.0, -0.0, 1.0, 1.5, -3.7, float("nan"), float("inf"), float("-inf"), 1e-33, 1e-32, 9.9e-33, 1e-20,
                    300.0, 1e30, -1e30, 3e9, -3e9, 0.5, 2.0 ** -126, 1e-45, 0.0031308, 0.999999, 1.0000001, 255.0, 256.0]


def stage_film(width, height, seed, scale=False):
    """A film for the output stage: uniform values, the edge values above sprinkled in, optionally wild magnitudes."""
    import numpy as np
    rng = np.random.default_rng(seed)
    f = rng.random((height, width, 3)).astype(np.float32)
    if scale:
        f *= rng.choice(np.array([1, 10, 1e-3, 1e5, -1, 1e-30, 1e25], np.float32), size=f.shape)
    flat = f.reshape(-1)
    edge = np.array(FILM_EDGE_VALUES, np.float32)
    for k in range(0, flat.size - edge.size, max(edge.size + 7, flat.size // 5)):
        flat[k:k + edge.size] = edge if (k // 7) % 2 == 0 else edge[::-1]
    return f


# (name, width, height, seed, scale): odd sizes so that 3*w is not a multiple of 4 and words straddle bmp lines
STAGE_FILMS = [("37x21", 37, 21, 1, False), ("37x21s", 37, 21, 2, True), ("1x1", 1, 1, 3, False), ("5x3", 5, 3, 4, True),
               ("64x48", 64, 48, 5, False), ("3x7", 3, 7, 6, True)]

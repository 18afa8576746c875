// kyd_kernels_big.cu -- the large-scene build of the kernels: the same sources as kyd_kernels.cu compiled into namespace
// kyd_big with the per-surface data in global memory and the three scene queries walking a bounding-volume hierarchy
// (kyd_device.cuh, KYD_BIG_SCENE).  kyd_api.cu picks this build when a scene has more than KYD_MAX_SURFACES surfaces.
#define KYD_BIG_SCENE 1
#include "kyd_kernels.cu"

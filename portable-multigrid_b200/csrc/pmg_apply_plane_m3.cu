// one epilogue mode of the plane-per-step apply kernel per translation unit (parallel compilation)
#define PMG_PLANE_TU_MODE 3
#include "pmg_apply_plane_launch.h"

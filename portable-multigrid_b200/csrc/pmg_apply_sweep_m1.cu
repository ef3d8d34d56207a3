// pmg_apply_sweep_m1.cu -- the line-marching apply kernel compiled for epilogue mode 1 (RESIDUAL: out = b - A u).
#define PMG_SWEEP_TU_MODE 1
#include "pmg_apply_sweep_launch.h"

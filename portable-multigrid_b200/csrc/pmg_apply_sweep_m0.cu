// pmg_apply_sweep_m0.cu -- the line-marching apply kernel compiled for epilogue mode 0 (APPLY: out = A u).
#define PMG_SWEEP_TU_MODE 0
#include "pmg_apply_sweep_launch.h"

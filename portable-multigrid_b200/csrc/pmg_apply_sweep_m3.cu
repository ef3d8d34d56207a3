// pmg_apply_sweep_m3.cu -- the line-marching apply kernel compiled for epilogue mode 3 (CHEB_STEP: out = u + f1 (u - xold) + f2 Dinv (b - A u)).
#define PMG_SWEEP_TU_MODE 3
#include "pmg_apply_sweep_launch.h"

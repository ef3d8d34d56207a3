// pmg_apply_sweep_m2.cu -- the line-marching apply kernel compiled for epilogue mode 2 (CHEB_FIRST: out = u + f2 Dinv (b - A u)).
#define PMG_SWEEP_TU_MODE 2
#include "pmg_apply_sweep_launch.h"

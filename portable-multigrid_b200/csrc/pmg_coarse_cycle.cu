// pmg_coarse_cycle.cu -- CUDA kernel + launcher of the single-CTA coarse V-cycle (csrc/pmg_coarse_cycle.h).  sm_100a only.
#include "pmg_coarse_cycle.h"
#include "pmg_cuda_common.h"
#include "pmg_kernels.h"

namespace {

constexpr int kThreads = 1024;

struct DeviceExec {
  template <class F> __device__ __forceinline__ void for_each_thread(F f) { f((int)threadIdx.x); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
};

template <int P>
__global__ void __launch_bounds__(kThreads, 1) k_coarse_cycle(const __grid_constant__ PmgCoarseParams q)
{
  DeviceExec ex;
  PmgCoarseCycle<kThreads, P>::run(q, ex);
}

} // namespace

extern "C" int pmgk_coarse_cycle_supported(const pmgk_level *lv)
{
  if (!lv || lv->dim != 3 || lv->coef || lv->degree + 1 > PMG_CC_MAX_N1) return 0;
  /* the whole level on this GPU */
  return lv->z0 == 0 && lv->nzl == lv->Nz && lv->cz_lo == 0 && lv->cz_hi == lv->nz;
}

extern "C" int pmgk_coarse_cycle(const pmgk_coarse_level *levels, int n_levels, int pre, int post, const double *P1d_host, void *stream)
{
  if (!levels || n_levels < 1 || n_levels > PMG_CC_MAX_LEVELS || pre < 0 || post < 0 || (n_levels > 1 && !P1d_host)) return PMG_ERR_ARG;
  PmgCoarseParams q;
  const pmgk_level *l0 = levels[0].lv;
  const int p = l0->degree, n1 = p + 1;
  q.n_levels = n_levels; q.pre = pre; q.post = post; q.p = p; q.faces = l0->faces;
  for (int i = 0; i < n1 * n1; ++i) { q.M[i] = l0->Mref[i]; q.K[i] = l0->Kref[i]; }
  for (int i = 0; i < n1 * (2 * p + 1); ++i) q.P1d[i] = (n_levels > 1) ? P1d_host[i] : 0.0;
  for (int l = 0; l < n_levels; ++l) {
    const pmgk_level *lv = levels[l].lv;
    if (!pmgk_coarse_cycle_supported(lv) || lv->degree != p || lv->faces != q.faces) return PMG_ERR_ARG;
    if (l > 0 && (lv->nx != 2 * levels[l - 1].lv->nx || lv->ny != 2 * levels[l - 1].lv->ny || lv->nz != 2 * levels[l - 1].lv->nz)) return PMG_ERR_ARG;
    PmgCoarseLevel &c = q.lv[l];
    c.nx = lv->nx; c.ny = lv->ny; c.nz = lv->nz; c.Nx = lv->Nx; c.Ny = lv->Ny; c.Nz = lv->Nz;
    c.cx = lv->h[1] * lv->h[2] / lv->h[0]; c.cy = lv->h[0] * lv->h[2] / lv->h[1]; c.cz = lv->h[0] * lv->h[1] / lv->h[2];
    c.degree = levels[l].cheb_degree; c.theta = levels[l].theta; c.delta = levels[l].delta;
    c.dinv_tab = lv->dinv_tab;
    c.sol = levels[l].sol; c.rhs = levels[l].rhs; c.tmp = levels[l].tmp; c.res = levels[l].res;
    if (!c.sol || !c.rhs || !c.tmp || !c.res || !c.dinv_tab) return PMG_ERR_ARG;
  }
  switch (p) {
    case 1: k_coarse_cycle<1><<<1, kThreads, 0, (cudaStream_t)stream>>>(q); break;
    case 2: k_coarse_cycle<2><<<1, kThreads, 0, (cudaStream_t)stream>>>(q); break;
    case 3: k_coarse_cycle<3><<<1, kThreads, 0, (cudaStream_t)stream>>>(q); break;
    case 4: k_coarse_cycle<4><<<1, kThreads, 0, (cudaStream_t)stream>>>(q); break;
    default: return PMG_ERR_UNSUPPORTED;
  }
  PMG_CUDA_CHECK(cudaGetLastError());
  pmg_count_launch(1);
  return 0;
}

// batched_reg.cu -- instantiations + launcher of the register-resident batched solver (batched_reg.cuh).
#include <stdlib.h>
#include <string.h>
#include "ctx.h"
#include "batched_reg.cuh"

using namespace lfpsqp;

namespace {
template <class Fam, int NPL, int ME, bool INEQ>
int launch(lfpsqp_ctx *c, BatchedArgs &A) {
  auto kern = batched_reg_kernel<Fam, NPL, ME, INEQ>;
  const int NA = A.n + A.p;
  const size_t smem = (size_t)(((INEQ ? 5 * NA : 0) + 1) & ~1) * 8;
  int resident = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, 128, smem);
  if (e != cudaSuccess || resident < 1) return c->cuda_fail(e, "occupancy query (batched_reg_kernel)");
  int64_t grid = (int64_t)c->sm_count * resident, need = (A.B + 3) / 4;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  cudaMemsetAsync(c->work_counter, 0, 8, c->stream);
  A.work_counter = c->work_counter;
  cudaEventRecord(c->ev0, c->stream);
  kern<<<(unsigned)grid, 128, smem, c->stream>>>(A);
  cudaEventRecord(c->ev1, c->stream);
  c->last_launches = 1;
  c->last_cfg_warps = 4; c->last_cfg_grid = (int)grid; c->last_cfg_smem = (int)smem; c->last_cfg_resident = resident;
  e = cudaGetLastError();
  if (e != cudaSuccess) return c->cuda_fail(e, "batched_reg_kernel launch");
  return LFPSQP_OK;
}
}  // namespace

// returns LFPSQP_OK when launched, 1 when this problem shape has no register-resident instantiation (caller falls
// back to the shared-memory warp kernel), < 0 on errors.  LFPSQP_BATCHED_KERNEL=smem forces the fallback (A/B tests).
int launch_batched_reg(lfpsqp_ctx *c, BatchedArgs &A) {
  const char *force = getenv("LFPSQP_BATCHED_KERNEL");
  if (force && strcmp(force, "smem") == 0) return 1;
  if (A.prm.linesearch != 0 && !A.prm.disable_linesearch) return 1;   // exact_linesearch! lives in the shared-memory solver
  const int NA = A.n + A.p, ME = A.m + A.p;
  switch (A.family) {
    case LFPSQP_FAM_README_INEQ:
      if (ME != 1 || !A.ineq) return 1;
      if (NA <= 32) return launch<SepReadmeIneq, 1, 1, true>(c, A);
      if (NA <= 64) return launch<SepReadmeIneq, 2, 1, true>(c, A);
      if (NA <= 128) return launch<SepReadmeIneq, 4, 1, true>(c, A);
      return 1;
    case LFPSQP_FAM_README_EQ:
      if (ME != 1 || NA > 64) return 1;
      return A.ineq ? launch<SepReadmeEq, 2, 1, true>(c, A) : launch<SepReadmeEq, 2, 1, false>(c, A);
    case LFPSQP_FAM_BOXQUAD:
      if (!A.ineq || NA > 32) return 1;
      return ME == 0 ? launch<SepBoxQuad, 1, 0, true>(c, A) : launch<SepBoxQuad, 1, 1, true>(c, A);
    default: return 1;
  }
}

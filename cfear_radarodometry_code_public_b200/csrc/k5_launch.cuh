// Host-side launchers of K5.  The 36 instantiations of k5_register<COST, LOSS, AUX> are spread over one translation unit
// per cost metric (k5_cost0.cu / k5_cost1.cu / k5_cost2.cu, bodies in k5_cost_tu.cuh) so that they compile in parallel.
#pragma once
#include "k5_register.cuh"

namespace cfear {

// cudaFuncAttributeMaxDynamicSharedMemorySize for every instantiation of the unit
cudaError_t k5_set_smem_cost0(int bytes);
cudaError_t k5_set_smem_cost1(int bytes);
cudaError_t k5_set_smem_cost2(int bytes);
// launch k5_register<COST, p.loss, p.solver_mode != ceres_lm> on `stream`; false if p.loss is not instantiated
bool k5_launch_cost0(const RegParams& p, int nprob, int smem, cudaStream_t stream, int prio = 0);
bool k5_launch_cost1(const RegParams& p, int nprob, int smem, cudaStream_t stream, int prio = 0);
bool k5_launch_cost2(const RegParams& p, int nprob, int smem, cudaStream_t stream, int prio = 0);

}  // namespace cfear

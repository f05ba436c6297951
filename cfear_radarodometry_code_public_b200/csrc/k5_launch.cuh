// Host-side launchers of K5.  The 36 + 2 x 18 instantiations of k5_register<COST, LOSS, AUX, NT> are spread over one translation unit
// per cost metric (k5_cost0.cu / k5_cost1.cu / k5_cost2.cu, bodies in k5_cost_tu.cuh) so that they compile in parallel.
#pragma once
#include "k5_register.cuh"

namespace cfear {

// cudaFuncAttributeMaxDynamicSharedMemorySize for every instantiation of the unit (bytes_mid / bytes_wide: the 192- and
// 384-thread forms, 0 = leave)
cudaError_t k5_set_smem_cost0(int bytes, int bytes_mid, int bytes_wide);
cudaError_t k5_set_smem_cost1(int bytes, int bytes_mid, int bytes_wide);
cudaError_t k5_set_smem_cost2(int bytes, int bytes_mid, int bytes_wide);
// launch k5_register<COST, p.loss, p.solver_mode != ceres_lm> on `stream`; false if p.loss is not instantiated.
// form 0: 128 threads, three CTAs per SM (`smem` bytes each); 1: 192 threads, two per SM; 2: 384 threads, one per SM -- with
// smem_form bytes of dynamic shared memory; forms 1 and 2 exist for the product instantiations only.
bool k5_launch_cost0(const RegParams& p, int nprob, int smem, int form, int smem_form, cudaStream_t stream, int prio = 0);
bool k5_launch_cost1(const RegParams& p, int nprob, int smem, int form, int smem_form, cudaStream_t stream, int prio = 0);
bool k5_launch_cost2(const RegParams& p, int nprob, int smem, int form, int smem_form, cudaStream_t stream, int prio = 0);

}  // namespace cfear

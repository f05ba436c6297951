// Body of one K5 translation unit; CFEAR_K5_TU_COST selects the cost metric (0 P2P, 1 P2L, 2 P2D).
#include "k5_launch.cuh"

#define K5_CAT2(a, b) a##b
#define K5_CAT(a, b) K5_CAT2(a, b)

namespace cfear {

#ifdef CFEAR_K5_MINIMAL      /* experiment builds (profiles/ab): only the Huber instantiations */
#define K5_FOR_EACH_LOSS(X) X(1)
#else
#define K5_FOR_EACH_LOSS(X) X(0) X(1) X(2) X(3) X(4) X(5)
#endif

cudaError_t K5_CAT(k5_set_smem_cost, CFEAR_K5_TU_COST)(int bytes, int bytes_mid, int bytes_wide) {
  cudaError_t e = cudaSuccess;
#define K5_ATTR(LO)                                                                                                              \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k5_register<CFEAR_K5_TU_COST, LO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k5_register<CFEAR_K5_TU_COST, LO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);  \
  if (e == cudaSuccess && bytes_mid > 0)                                                                                         \
    e = cudaFuncSetAttribute(k5_register<CFEAR_K5_TU_COST, LO, false, K5_THREADS_MID>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes_mid); \
  if (e == cudaSuccess && bytes_wide > 0)                                                                                        \
    e = cudaFuncSetAttribute(k5_register<CFEAR_K5_TU_COST, LO, false, K5_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes_wide);
  K5_FOR_EACH_LOSS(K5_ATTR)
#undef K5_ATTR
  return e;
}

bool K5_CAT(k5_launch_cost, CFEAR_K5_TU_COST)(const RegParams& p, int nprob, int smem, int form, int smem_form, cudaStream_t stream, int prio) {
  // gn_fixed / cost only, and the ceres_lm loop with association outputs or the soft prior: the AUX instantiations
  const bool aux = p.solver_mode != 0 || p.assoc != nullptr || p.assoc_sim != nullptr || p.soft_L != nullptr;
  // form 2 / 1: the caller found the batch small enough for one 384-thread / two 192-thread CTAs per SM (k5_register.cuh)
  RegParams pw = p;
  pw.smem_bytes = smem_form;
  switch (p.loss) {
#define K5_CASE(LO)                                                                                       \
  case LO:                                                                                                \
    if (aux) launch_with_priority(k5_register<CFEAR_K5_TU_COST, LO, true>, nprob, K5_THREADS, smem, stream, prio, p);   \
    else if (form == 2)                                                                                   \
      launch_with_priority(k5_register<CFEAR_K5_TU_COST, LO, false, K5_THREADS_WIDE>, nprob, K5_THREADS_WIDE, smem_form, stream, prio, pw); \
    else if (form == 1)                                                                                   \
      launch_with_priority(k5_register<CFEAR_K5_TU_COST, LO, false, K5_THREADS_MID>, nprob, K5_THREADS_MID, smem_form, stream, prio, pw);   \
    else launch_with_priority(k5_register<CFEAR_K5_TU_COST, LO, false>, nprob, K5_THREADS, smem, stream, prio, p);      \
    return true;
    K5_FOR_EACH_LOSS(K5_CASE)
#undef K5_CASE
    default: return false;
  }
}

}  // namespace cfear

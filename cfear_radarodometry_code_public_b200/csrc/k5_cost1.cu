// K5 instantiations for cost metric 1 (P2L); see k5_launch.cuh
#define CFEAR_K5_TU_COST 1
#include "k5_cost_tu.cuh"

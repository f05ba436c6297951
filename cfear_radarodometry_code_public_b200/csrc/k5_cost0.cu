// K5 instantiations for cost metric 0 (P2P); see k5_launch.cuh
#define CFEAR_K5_TU_COST 0
#include "k5_cost_tu.cuh"

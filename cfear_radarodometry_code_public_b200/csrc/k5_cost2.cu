// K5 instantiations for cost metric 2 (P2D); see k5_launch.cuh
#define CFEAR_K5_TU_COST 2
#include "k5_cost_tu.cuh"

// GENERATED -- instantiates the solve and filter kernels for model 'hs_nk'.
#include "model_hs_nk.cuh"
#define DYB_MODEL dyb::Model_hs_nk
#include "../model_tu.cuh"

// GENERATED -- instantiates the solve and filter kernels for model 'ls2003'.
#include "model_ls2003.cuh"
#define DYB_MODEL dyb::Model_ls2003
#include "../model_tu.cuh"

// GENERATED -- instantiates the solve and filter kernels for model 'mc4'.
#include "model_mc4.cuh"
#define DYB_MODEL dyb::Model_mc4
#include "../model_tu.cuh"

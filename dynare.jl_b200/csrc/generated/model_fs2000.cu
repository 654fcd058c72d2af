// GENERATED -- instantiates the solve and filter kernels for model 'fs2000'.
#include "model_fs2000.cuh"
#define DYB_MODEL dyb::Model_fs2000
#include "../model_tu.cuh"

// Beam-search kernels (placeholder translation unit; filled in by the beam-search milestone).
#include "kernels.cuh"
namespace gic {}

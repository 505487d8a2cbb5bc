// Step kernels of one contact model (see ../launch.cuh): its own translation unit so that the library builds in parallel.
#include "../launch.cuh"
namespace od { OD_INSTANTIATE_STEP(planar_push, PlanarPushModel, true, true) }

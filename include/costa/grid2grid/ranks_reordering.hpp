// costa::optimal_reordering lives with the volume graph it works on (comm_volume.hpp); this header keeps the reference's
// include path (libs/COSTA/src/costa/grid2grid/ranks_reordering.hpp) working.
#pragma once
#include <costa/grid2grid/comm_volume.hpp>

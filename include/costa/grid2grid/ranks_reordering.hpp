#pragma once
#include <costa/grid2grid/comm_volume.hpp>

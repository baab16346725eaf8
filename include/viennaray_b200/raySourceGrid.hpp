// Drop-in name of the reference header include/viennaray/raySourceGrid.hpp: the host-side helpers live
// in vr_host_extras.hpp (on top of vr_host.hpp, the B200 host mirror of the ViennaRay interface).
#pragma once
#include "vr_host_extras.hpp"

// Drop-in name of the reference header include/viennaray/rayParticle.hpp: everything
// lives in vr_host.hpp (B200 host mirror of the ViennaRay interface).
#pragma once
#include "vr_host.hpp"

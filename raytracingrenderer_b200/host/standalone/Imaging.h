// Stand-alone build of the C++ host layer: code written against RTBase's headers (#include "Imaging.h")
// resolves here and gets the product's own scene API (host/rtb_scene.hpp) instead of the reference's.
#pragma once
#include "../rtb_standalone.hpp"

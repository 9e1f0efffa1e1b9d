// cvlite stand-in for <opencv2/calib3d.hpp>; see cvlite.hpp
#pragma once
#include "cvlite.hpp"

// cvlite stand-in for <opencv2/opencv_modules.hpp>; see cvlite.hpp
#pragma once
#include "cvlite.hpp"

// cvlite stand-in for <opencv2/imgproc.hpp>; see cvlite.hpp
#pragma once
#include "cvlite.hpp"

// cvlite stand-in for <opencv2/xfeatures2d.hpp>; see cvlite.hpp
#pragma once
#include "cvlite.hpp"

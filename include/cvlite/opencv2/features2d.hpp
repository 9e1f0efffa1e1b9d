// cvlite stand-in for <opencv2/features2d.hpp>; see cvlite.hpp
#pragma once
#include "cvlite.hpp"

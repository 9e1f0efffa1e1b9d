// cvlite stand-in for <opencv2/core/ocl.hpp>; see cvlite.hpp
#pragma once
#include "../cvlite.hpp"

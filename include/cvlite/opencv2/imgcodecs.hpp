// cvlite stand-in for <opencv2/imgcodecs.hpp>; see cvlite.hpp
#pragma once
#include "cvlite.hpp"

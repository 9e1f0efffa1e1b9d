// cvlite stand-in for <opencv2/highgui.hpp>; see cvlite.hpp
#pragma once
#include "cvlite.hpp"

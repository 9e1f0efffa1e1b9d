// cvlite stand-in for <opencv2/core.hpp>; see cvlite.hpp
#pragma once
#include "cvlite.hpp"

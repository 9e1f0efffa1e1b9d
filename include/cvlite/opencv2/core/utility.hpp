// cvlite stand-in for <opencv2/core/utility.hpp>; see cvlite.hpp
#pragma once
#include "../cvlite.hpp"

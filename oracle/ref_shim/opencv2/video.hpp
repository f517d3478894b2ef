// TEST INFRASTRUCTURE: see lvb_cv.hpp in this directory (stand-in for the OpenCV headers the reference includes).
#include "lvb_cv.hpp"

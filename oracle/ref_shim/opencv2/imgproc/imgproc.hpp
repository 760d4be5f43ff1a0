#include "cv_standin.hpp"  // stand-in, see cv_standin.hpp

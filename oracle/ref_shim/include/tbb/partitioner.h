// stand-in, see shim_core.hpp
#include "shim_core.hpp"

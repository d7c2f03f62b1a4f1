// Stand-in for the CMake-generated rakau/config.hpp (reference config.hpp.in): version macros only, no
// accelerator backend (the shimmed build exercises the reference's CPU path).
#ifndef RAKAU_CONFIG_HPP
#define RAKAU_CONFIG_HPP
#define RAKAU_VERSION_STRING "0.1-shim"
#define RAKAU_VERSION_MAJOR 0
#define RAKAU_VERSION_MINOR 1
#endif

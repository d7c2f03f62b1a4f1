// Counterpart of the reference's benchmark/benchmark_pot.cpp (potentials): see accpot_main.hpp.
#include "accpot_main.hpp"

int main(int argc, char **argv) { return rakau_benchmark::accpot_main<1>(argc, argv); }

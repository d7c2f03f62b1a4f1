// Counterpart of the reference's benchmark/benchmark_acc_pot.cpp (accelerations + potentials): see accpot_main.hpp.
#include "accpot_main.hpp"

int main(int argc, char **argv) { return rakau_benchmark::accpot_main<2>(argc, argv); }

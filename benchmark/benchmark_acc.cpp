// Counterpart of the reference's benchmark/benchmark_acc.cpp (accelerations): see accpot_main.hpp.
#include "accpot_main.hpp"

int main(int argc, char **argv) { return rakau_benchmark::accpot_main<0>(argc, argv); }

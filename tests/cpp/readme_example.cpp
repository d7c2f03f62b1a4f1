// The README snippet of the reference (test/readme_example.cpp) must compile and run unchanged.
#include <array>
#include <initializer_list>
#include <vector>

#include <rakau/tree.hpp>

using namespace rakau;
using namespace rakau::kwargs;

int main()
{
    // Create an octree from a set of particle coordinates and masses.
    octree<float> t{x_coords = {1, 2, 3}, y_coords = {4, 5, 6}, z_coords = {7, 8, 9}, masses = {1, 1, 1}};

    // Prepare output vectors for the accelerations.
    std::array<std::vector<float>, 3> accs;

    // Compute the accelerations with a theta parameter of 0.4.
    t.accs_u(accs, 0.4f);
    return accs[0].size() == 3u ? 0 : 1;
}

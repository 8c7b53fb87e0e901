// test helper: dump draws of simu_b200/host/rng.h so that tests can compare them with an
// independent Python restatement of the same published algorithms
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../simu_b200/host/rng.h"

int main(int argc, char **argv)
{
    const long long seed = argc > 1 ? atoll(argv[1]) : 1;
    const int n = argc > 2 ? atoi(argv[2]) : 10;
    simu::Mt19937 rng(seed);
    std::vector<int> perm(n);
    for (int i = 0; i < n; i++) perm[i] = i;
    simu::random_shuffle(perm, rng);
    printf("shuffle");
    for (int v : perm) printf(" %d", v);
    printf("\nnormal");
    {
        simu::NormalDraws d(rng);
        for (int i = 0; i < n; i++) printf(" %.17g", d());
    }
    printf("\nnormal2");
    {
        simu::NormalDraws d(rng);   // fresh object: cached half of the pair is dropped
        for (int i = 0; i < 3; i++) printf(" %.17g", d());
    }
    const unsigned a = rng(), b = rng();
    printf("\nraw %u %u\n", a, b);
    return 0;
}

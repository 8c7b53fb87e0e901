// TextWriter must print doubles exactly like the reference's `ostream << double`
// (/root/reference/src/source.cc:1066, :1101-1103): compares 6e6 values, incl. specials.
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <random>
#include <sstream>

#include "../../simu_b200/host/fastio.h"

int main(int argc, char **argv)
{
    const std::string path = argv[1];
    std::mt19937_64 r(1);
    std::normal_distribution<double> nd;
    std::vector<double> vals = {0.0, -0.0, 1, 2, -9, 0.5, 1e6, 999999.5, 1e-5, 123456.5, 1234567, 0.1, 100000, 1e-4, 0.00001,
                                std::numeric_limits<double>::infinity(), -std::numeric_limits<double>::infinity(),
                                std::nan(""), 5e-324, 1.7976931348623157e308, 0.9999995, 9.9999995e-5};
    for (int i = 0; i < 2000000; i++) {
        unsigned long long u = r();
        double v;
        memcpy(&v, &u, 8);
        if (!std::isnan(v)) vals.push_back(v);
        vals.push_back(nd(r));
        vals.push_back(std::round(nd(r) * 1000) / 8);
    }
    {
        simu::TextWriter w(path);
        for (double v : vals) w << v << '\n';
        w << (long long)-1234567890123LL << '\t' << 42 << '\t' << std::string("x") << "y\n";
    }
    std::ifstream in(path);
    std::string line;
    size_t i = 0, bad = 0;
    for (; i < vals.size() && std::getline(in, line); i++) {
        std::ostringstream os;
        os << vals[i];
        if (os.str() != line) bad++;
    }
    std::getline(in, line);
    if (line != "-1234567890123\t42\txy") bad++;
    printf("%zu values, %zu differences\n", i, bad);
    return (bad == 0 && i == vals.size()) ? 0 : 1;
}

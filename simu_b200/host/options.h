// Command-line surface of `simu` without Boost: the same option names, arities, defaults
// and validation messages as /root/reference/src/source.cc:1116-1162 (declaration),
// :157-185 (numbers -- including negative ones -- are always values, never options) and
// :423-700 (fix_and_validate).  Parsing errors carry boost::program_options' wording so that
// the "Exception  : ..." line a user sees is unchanged.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <functional>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace simu {

#define SIMU_VERSION "v0.9.4"
#define DEFAULT_SIGSQ 1.0f
#define DEFAULT_CAUSAL_PI 0.001f
#define DEFAULT_HSQ 0.7f
#define DEFAULT_RG 0.0f
#define DEFAULT_K 0.1f

struct SimuOptions {                       // source.cc:57-105
    std::string bfile, bfile_chr;
    std::vector<std::string> chr_labels;
    bool qt = false, cc = false;
    int num_traits = 1;
    std::vector<float> k;
    std::vector<int> ncas, ncon;
    std::vector<float> hsq;
    int num_components = 1;
    std::vector<std::string> causal_variants;
    std::vector<int> causal_n;
    std::vector<float> causal_pi;
    std::vector<std::string> causal_regions;
    std::vector<float> trait1_sigsq, trait2_sigsq, trait1_s_pow, trait2_s_pow, rg;
    bool gcta_sigma = false, norm_effect = false;
    std::string out = "simu";
    int64_t seed = 0;
    int trait2_snp_offset = 0;
    bool help = false;
    // derived
    std::vector<std::string> bfiles;
    int num_samples = -1, num_variants = 0;
    std::string out_pheno;
    std::vector<std::string> out_causals;
    std::set<std::string> seen;            // vm.count(name) > 0
    bool count(const char *name) const { return seen.count(name) != 0; }
};

class OptionError : public std::runtime_error {
public:
    explicit OptionError(const std::string &m) : std::runtime_error(m) {}
};

enum class Arity { Switch, One, Many };

struct OptionSpec {
    const char *name;
    Arity arity;
    const char *help;
    std::function<void(SimuOptions &, const std::string &, const std::string &)> set;   // (opts, value, optname)
};

const std::vector<OptionSpec> &option_table();
void parse_command_line(int argc, char *argv[], SimuOptions *o);
std::string help_text();

}  // namespace simu

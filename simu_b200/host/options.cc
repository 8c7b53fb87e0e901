#include "options.h"

#include <cerrno>
#include <cstring>
#include <sstream>

namespace simu {

namespace {

bool is_number(const std::string &s)
{
    // boost::lexical_cast<double>(arg) succeeds (source.cc:163-168): whole token, no blanks
    if (s.empty() || s[0] == ' ' || s[0] == '\t') return false;
    if (s.size() > 1 && (s[1] == 'x' || s[1] == 'X')) return false;
    if (s.size() > 2 && (s[0] == '-' || s[0] == '+') && (s[2] == 'x' || s[2] == 'X')) return false;
    char *end = nullptr;
    strtod(s.c_str(), &end);
    return end && *end == '\0' && end != s.c_str();
}

[[noreturn]] void invalid(const std::string &value, const std::string &opt)
{
    throw OptionError("the argument ('" + value + "') for option '--" + opt + "' is invalid");
}

int to_int(const std::string &v, const std::string &opt)
{
    char *end = nullptr;
    errno = 0;
    long x = strtol(v.c_str(), &end, 10);
    if (v.empty() || !end || *end != '\0' || errno || x < INT32_MIN || x > INT32_MAX) invalid(v, opt);
    return (int)x;
}

int64_t to_i64(const std::string &v, const std::string &opt)
{
    char *end = nullptr;
    errno = 0;
    long long x = strtoll(v.c_str(), &end, 10);
    if (v.empty() || !end || *end != '\0' || errno) invalid(v, opt);
    return (int64_t)x;
}

float to_float(const std::string &v, const std::string &opt)
{
    if (!is_number(v)) invalid(v, opt);
    return strtof(v.c_str(), nullptr);
}

template <typename T, typename F>
std::function<void(SimuOptions &, const std::string &, const std::string &)> push(std::vector<T> SimuOptions::*m, F conv)
{
    return [m, conv](SimuOptions &o, const std::string &v, const std::string &n) { (o.*m).push_back(conv(v, n)); };
}

std::string ident(const std::string &v, const std::string &) { return v; }

}  // namespace

const std::vector<OptionSpec> &option_table()
{
    using S = SimuOptions;
    static const std::vector<OptionSpec> t = {
        {"help", Arity::Switch, "produce this help message", [](S &o, const std::string &, const std::string &) { o.help = true; }},
        {"bfile", Arity::One, "prefix for plink .bed/.bim/.fam file",
         [](S &o, const std::string &v, const std::string &) { o.bfile = v; }},
        {"bfile-chr", Arity::One,
         "same as --bfile, but will automatically concatenate .bed/.bim/.fam files split across 22 chromosomes. "
         "If the filename prefix contains the symbol @, SIMU will replace the @ symbol with chromosome numbers. "
         "Otherwise, SIMU will append chromosome numbers to the end of the filename prefix.",
         [](S &o, const std::string &v, const std::string &) { o.bfile_chr = v; }},
        {"chr-labels", Arity::Many,
         "Set of chromosome labels. Defaults to '1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16 17 18 19 20 21 22'",
         push(&S::chr_labels, ident)},
        {"qt", Arity::Switch, "simulate quantitative trait", [](S &o, const std::string &, const std::string &) { o.qt = true; }},
        {"cc", Arity::Switch, "simulate case/control trait", [](S &o, const std::string &, const std::string &) { o.cc = true; }},
        {"num-traits", Arity::One, "number of traits (either 1 or 2 traits are supported)",
         [](S &o, const std::string &v, const std::string &n) { o.num_traits = to_int(v, n); }},
        {"k", Arity::Many, "prevalence for case/control traits, by default 0.1; one value per trait", push(&S::k, to_float)},
        {"ncas", Arity::Many, "number of cases, by default N*k, where N is sample size; one value per trait",
         push(&S::ncas, to_int)},
        {"ncon", Arity::Many, "number of controls, by default N*(1-k), where N is sample size; one value per trait",
         push(&S::ncon, to_int)},
        {"hsq", Arity::Many, "heritability, by default 0.7; one value per trait", push(&S::hsq, to_float)},
        {"num-components", Arity::One, "number of components in the mixture",
         [](S &o, const std::string &v, const std::string &n) { o.num_components = to_int(v, n); }},
        {"causal-pi", Arity::Many, "proportion of causal variants; by default 0.001; one value per mixture component",
         push(&S::causal_pi, to_float)},
        {"causal-n", Arity::Many,
         "number of causal variants (alternative to --causal-pi option); one value per mixture component",
         push(&S::causal_n, to_int)},
        {"causal-variants", Arity::Many,
         "file with a list of causal variants and, optionally, their effect sizes; one file per mixture component. "
         "This is an alternative option to --causal-pi and --causal-n. See README.md file for detailed description "
         "of file formats.",
         push(&S::causal_variants, ident)},
        {"causal-regions", Arity::Many,
         "file with a list of non-overlapping regions to distribute causal variants. Regions must be defined as a "
         "list of variant names (e.g. RS numbers, one value per line). Supplements --causal-pi or --causal-n "
         "options; can not be used together with --causal-variants option. One file per mixture component",
         push(&S::causal_regions, ident)},
        {"trait1-sigsq", Arity::Many,
         "variance of effect sizes for trait1 per causal marker; by default 1.0; one value per mixture component",
         push(&S::trait1_sigsq, to_float)},
        {"trait2-sigsq", Arity::Many,
         "variance of effect sizes for trait2 per causal marker; by default 1.0; one value per mixture component",
         push(&S::trait2_sigsq, to_float)},
        {"trait1-s-pow", Arity::Many,
         "draw effect sizes on the first trait with variance proportional to (2*p(1-p))^S, where parameter S is "
         "defined by trait1-s-pow, and p is allele frequency; by default 0.0; one value per mixture component",
         push(&S::trait1_s_pow, to_float)},
        {"trait2-s-pow", Arity::Many,
         "draw effect sizes on the second trait with variance proportional to (2*p(1-p))^S, where parameter S is "
         "defined by trait2-s-pow, where p is allele frequency; by default 0.0; one value per mixture component",
         push(&S::trait2_s_pow, to_float)},
        {"gcta-sigma", Arity::Switch,
         "draw effect sizes with variance inversely proportional to 2*p(1-p), where p is allele frequency; this "
         "corresponds to --trait1-s-pow and --trait2-s-pow set to -1.0",
         [](S &o, const std::string &, const std::string &) { o.gcta_sigma = true; }},
        {"norm-effect", Arity::Switch,
         "report effect sizes w.r.t. normalized genotypes (e.i. additively coded 0,1,2 genotypes devided by "
         "sqrt(2*p(1-p)), where p is allele frequency). Default behavior without --norm-effect is to report effect "
         "size w.r.t. additively coded 0,1,2 genotypes.",
         [](S &o, const std::string &, const std::string &) { o.norm_effect = true; }},
        {"rg", Arity::Many, "coefficient of genetic correlation; by default 0.0; one value per mixture component",
         push(&S::rg, to_float)},
        {"seed", Arity::One, "seed for random numbers generator (default is time-dependent seed)",
         [](S &o, const std::string &v, const std::string &n) { o.seed = to_i64(v, n); }},
        {"trait2-snp-offset", Arity::One,
         "shifts causal variant in trait2 by a given number of positions (for example, '--trait2-snp-offset 1' "
         "indicates that causal SNPs will correspond to adjacent rows in the input bim file)",
         [](S &o, const std::string &v, const std::string &n) { o.trait2_snp_offset = to_int(v, n); }},
        {"out", Arity::One,
         "prefix of the output files; will generate .pheno file containing synthesized phenotypes; and .*.causals "
         "files (one file per trait) containing lists of causal variants and their effect sizes for each component "
         "in the mixture. See README.md file for detailed description of file formats.",
         [](S &o, const std::string &v, const std::string &) { o.out = v; }},
    };
    return t;
}

static const OptionSpec *find_option(const std::string &token)
{
    const auto &t = option_table();
    for (const auto &s : t)
        if (token == s.name) return &s;
    // unambiguous prefix, as program_options' default style allows
    std::vector<const OptionSpec *> hits;
    for (const auto &s : t)
        if (strncmp(s.name, token.c_str(), token.size()) == 0) hits.push_back(&s);
    if (hits.size() == 1) return hits[0];
    if (hits.size() > 1) {
        std::string m = "option '--" + token + "' is ambiguous and matches ";
        for (size_t i = 0; i < hits.size(); i++) {
            if (i) m += (i + 1 == hits.size()) ? ", and " : ", ";
            m += std::string("'--") + hits[i]->name + "'";
        }
        throw OptionError(m);
    }
    throw OptionError("unrecognised option '--" + token + "'");
}

void parse_command_line(int argc, char *argv[], SimuOptions *o)
{
    int i = 1;
    while (i < argc) {
        std::string tok = argv[i];
        if (is_number(tok) || tok.size() < 2 || tok[0] != '-') {
            // a bare value that no multitoken option claimed: program_options turns it into a
            // nameless positional option and, with no positional description given, store()
            // skips it (source.cc:157-185 makes every number such a positional)
            i++;
            continue;
        }
        std::string name, adjacent;
        bool has_adjacent = false;
        if (tok == "-h") name = "help";
        else if (tok[1] == '-') {
            name = tok.substr(2);
            const size_t eq = name.find('=');
            if (eq != std::string::npos) { adjacent = name.substr(eq + 1); name = name.substr(0, eq); has_adjacent = true; }
        } else throw OptionError("unrecognised option '" + tok + "'");
        const OptionSpec *spec = find_option(name);
        const std::string canon = spec->name;
        i++;
        std::vector<std::string> values;
        if (has_adjacent) values.push_back(adjacent);
        if (spec->arity == Arity::Switch) {
            if (has_adjacent) throw OptionError("option '--" + canon + "' does not take any arguments");
        } else {
            const size_t max_tokens = spec->arity == Arity::One ? 1 : (size_t)-1;
            // following tokens are values until one of them is a known option; numbers
            // (negative ones included) are always values
            while (i < argc && values.size() < max_tokens) {
                std::string v = argv[i];
                if (!is_number(v) && v.size() >= 2 && v[0] == '-') break;
                values.push_back(v);
                i++;
            }
            if (values.empty()) throw OptionError("the required argument for option '--" + canon + "' is missing");
        }
        if (spec->arity == Arity::One && o->seen.count(canon))
            throw OptionError("option '--" + canon + "' cannot be specified more than once");
        o->seen.insert(canon);
        if (spec->arity == Arity::Switch) spec->set(*o, "", canon);
        else for (const auto &v : values) spec->set(*o, v, canon);
    }
}

std::string help_text()
{
    std::ostringstream ss;
    ss << "SIMU " SIMU_VERSION " - library for simulation of GWAS summary statistics:\n";
    for (const auto &s : option_table()) {
        std::string head = std::string("  ") + (strcmp(s.name, "help") == 0 ? "-h [ --help ]" : std::string("--") + s.name);
        if (s.arity != Arity::Switch) head += strcmp(s.name, "num-traits") == 0 || strcmp(s.name, "num-components") == 0
                                                  ? " arg (=1)" : (strcmp(s.name, "out") == 0 ? " arg (=simu)" : " arg");
        if (head.size() < 28) head.resize(28, ' '); else head += "\n" + std::string(28, ' ');
        // wrap the description at 52 columns like program_options does
        std::string d = s.help, line;
        std::istringstream words(d);
        std::string w;
        bool first = true;
        while (words >> w) {
            if (line.size() + w.size() + 1 > 52 && !line.empty()) {
                ss << (first ? head : std::string(28, ' ')) << line << "\n";
                first = false;
                line.clear();
            }
            line += (line.empty() ? "" : " ") + w;
        }
        ss << (first ? head : std::string(28, ' ')) << line << "\n";
    }
    return ss.str();
}

}  // namespace simu

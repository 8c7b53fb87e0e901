// simu_b200 -- drop-in command line for precimed/simu's phenotype synthesis (y = Gx + e)
// with the hot path on a B200.  Same options, same .pheno / .causals / .log outputs as the
// reference's main() (/root/reference/src/source.cc:1111-1344); what changed is below the
// four calls
//     find_freq                  source.cc:1235  -> simu_b200_panel_load_bed + simu_b200_find_freq
//     find_pheno                 source.cc:1295  -> simu_b200_find_pheno
//     apply_heritability         source.cc:1299  -> simu_b200_apply_heritability
//     apply_liability_threshold  source.cc:1304  -> simu_b200_liability_order (+ host shuffles)
// Causal selection, effect-size and noise draws stay on the host with the reference's RNG
// order (rng.h).  Environment: SIMU_B200_MODE=exact|fast (default exact: find_pheno bit-identical to the
// reference on one GPU; fast = the tensor-core kernel, ~1e-12 of max|y|; anything else is an error),
// SIMU_B200_DEVICE=n, SIMU_B200_DEVICES=all|0,1,...  The mode and the devices are written to <out>.log.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <ctime>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <thread>

#include <sys/stat.h>

#include "../../include/simu_b200.h"
#include "fastio.h"
#include "options.h"
#include "plink_text.h"
#include "rng.h"

using namespace simu;

namespace {

class Logger {                                   // source.cc:107-133: tee to stdout and <out>.log
public:
    explicit Logger(const std::string &path) : file_(path) {}
    template <typename T> Logger &operator<<(const T &v) { std::cout << v; file_ << v; file_.flush(); return *this; }
    template <typename T> Logger &operator<<(const std::vector<T> &v)
    {
        for (size_t i = 0; i < v.size(); i++) (*this) << (i > 0 ? " " : "") << v[i];
        return *this;
    }
private:
    std::ofstream file_;
};

bool file_exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }

// SIMU_B200_TIMING=1: wall time per stage as one JSON line on stderr (the .log and stdout stay as the reference's)
struct StageTimer {
    std::vector<std::pair<std::string, double>> stages;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *name)
    {
        const auto t1 = std::chrono::steady_clock::now();
        stages.emplace_back(name, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    }
    void report() const
    {
        if (!getenv("SIMU_B200_TIMING")) return;
        double total = 0;
        std::cerr << "{";
        for (const auto &s : stages) { std::cerr << "\"" << s.first << "_s\": " << s.second << ", "; total += s.second; }
        std::cerr << "\"total_s\": " << total << "}\n";
    }
};

std::string now_string(std::time_t t)            // boost::posix_time::ptime operator<<: 2018-Jan-01 12:00:00
{
    char buf[64];
    std::tm tm;
    localtime_r(&t, &tm);
    strftime(buf, sizeof(buf), "%Y-%b-%d %H:%M:%S", &tm);
    return buf;
}

std::string duration_string(long s)              // boost time_duration operator<<: HH:MM:SS
{
    char buf[64];
    snprintf(buf, sizeof(buf), "%02ld:%02ld:%02ld", s / 3600, (s / 60) % 60, s % 60);
    return buf;
}

void log_header(int argc, char *argv[], Logger &log)   // source.cc:135-154
{
    log << "*********************************************************************\n"
           "* SIMU - library for simulation of GWAS summary statistics\n"
           "* Version " SIMU_VERSION " (simu_b200: B200-native hot path)\n"
           "* (C) 2018 Oleksandr Frei\n"
           "* Norwegian Centre for Mental Disorders Research / University of Oslo\n"
           "* GNU General Public License v3\n"
           "*********************************************************************\n";
    if (argc == 0) return;
    log << "Call:\n" << argv[0] << " ";
    for (int i = 1; i < argc; i++) {
        if (strlen(argv[i]) == 0) continue;
        if (argv[i][0] == '-') log << "\\\n\t";
        log << argv[i] << " ";
    }
    log << "\n\n";
}

std::string replace_all(std::string s, const std::string &from, const std::string &to)
{
    for (size_t p = 0; (p = s.find(from, p)) != std::string::npos; p += to.size()) s.replace(p, from.size(), to);
    return s;
}

// ---------------------------------------------------------------- causal-variants files

void read_causal_variants_file(const std::string &file, Logger *log, const PlinkFiles &pio,
                               std::vector<int> *variant_indices, std::vector<double> *effect1,
                               std::vector<double> *effect2)          // source.cc:363-421
{
    std::ifstream in(file);
    std::string line;
    variant_indices->clear();
    if (effect1) effect1->clear();
    if (effect2) effect2->clear();
    int num_tokens = -1, line_num = 1;
    if (log) (*log) << "Reading '" << file << "'... ";
    while (std::getline(in, line)) {
        std::vector<std::string> tok;
        {   // trim + split on any of " \t\n\r", compressing runs
            size_t i = 0;
            while (i < line.size()) {
                while (i < line.size() && strchr(" \t\n\r", line[i])) i++;
                size_t j = i;
                while (j < line.size() && !strchr(" \t\n\r", line[j])) j++;
                if (j > i) tok.emplace_back(line, i, j - i);
                i = j;
            }
        }
        if (tok.empty()) continue;
        if (num_tokens == -1) num_tokens = (int)tok.size();
        if ((int)tok.size() != num_tokens) {
            std::stringstream ss;
            ss << "ERROR: failed to parse line " << line_num << " from '" << file << "', invalid number of tokens ("
               << tok.size() << " found, " << num_tokens << " expected)";
            throw std::runtime_error(ss.str());
        }
        const int variant_index = pio.get_locus_id(tok[0]);
        double e1 = 0, e2 = 0;
        for (int c = 1; c <= 2 && c < (int)tok.size(); c++) {
            double v;
            const bool ok = full_double(tok[c], &v) && !strchr(" \t", tok[c][0]);
            if (!ok) {
                std::stringstream ss;
                ss << "ERROR: failed to parse line " << line_num << " from '" << file << "', effect size '" << tok[c]
                   << "' is not a number";
                throw std::runtime_error(ss.str());
            }
            (c == 1 ? e1 : e2) = v;
        }
        variant_indices->push_back(variant_index);
        if (effect1 && tok.size() > 1) effect1->push_back(e1);
        if (effect2 && tok.size() > 2) effect2->push_back(e2);
        line_num++;
    }
    if (variant_indices->empty())
        throw std::runtime_error("ERROR: '" + file + "' is empty or contains no variants");
    if (log) (*log) << variant_indices->size() << " variants found.\n";
}

// ---------------------------------------------------------------- fix_and_validate

void fix_and_validate(SimuOptions &o, Logger &log, PlinkFiles *pio)    // source.cc:423-700
{
    if (o.chr_labels.empty())
        for (int i = 1; i <= 22; i++) o.chr_labels.push_back(std::to_string(i));

    if (o.bfile.empty() && o.bfile_chr.empty())
        throw std::invalid_argument("ERROR: Either --bfile or --bfile-chr must be specified");
    if (!o.bfile_chr.empty()) {
        if (o.bfile_chr.find('@') == std::string::npos) o.bfile_chr.append("@");
        for (const auto &lab : o.chr_labels) o.bfiles.push_back(replace_all(o.bfile_chr, "@", lab));
    } else {
        o.bfiles.push_back(o.bfile);
    }

    if (!o.cc && !o.qt) throw std::invalid_argument("ERROR: Either --qt or --cc must be specified");
    if (o.num_traits <= 0 || o.num_traits >= 3) throw std::invalid_argument("ERROR: --num-traits must be either 1 or 2");
    if (o.num_components <= 0) throw std::invalid_argument("ERROR: --num-components must be non-negative number");
    if (o.out.empty()) throw std::invalid_argument("ERROR: --out option must be specified");
    o.out_pheno = o.out + ".pheno";
    for (int i = 0; i < o.num_traits; i++) o.out_causals.push_back(o.out + "." + std::to_string(i + 1) + ".causals");
    if (file_exists(o.out_pheno))
        log << "WARNING: Target file " << o.out_pheno << " already exists and will be overwritten\n";

    if (o.qt && (o.count("k") || o.count("ncas") || o.count("ncon"))) {
        log << "WARNING: Options --k, --ncas, --ncon are not relevant to --qt, and will be ignored\n";
        o.k.clear(); o.ncas.clear(); o.ncon.clear();
    }
    if (o.num_traits == 1 && (o.count("trait2-sigsq") || o.count("rg"))) {
        log << "WARNING: Options --trait2-sigsq and --rg are not relevant for a single-trait simulations, and will be ignored\n";
        o.trait2_sigsq.clear(); o.rg.clear();
    }
    if (o.cc && (o.ncas.empty() != o.ncon.empty()))
        throw std::invalid_argument("ERROR: Options --ncas and --ncon must be either both present, or both absent");

    const size_t nt = (size_t)o.num_traits, nc = (size_t)o.num_components;
    if (o.cc) {
        if (!o.k.empty() && o.k.size() != nt)
            throw std::invalid_argument("ERROR: Number of --k values does not match the number of traits");
        if (!o.ncas.empty() && o.ncas.size() != nt)
            throw std::invalid_argument("ERROR: Number of --ncas values does not match the number of traits");
        if (!o.ncon.empty() && o.ncon.size() != nt)
            throw std::invalid_argument("ERROR: Number of --ncon values does not match the number of traits");
    }
    if (!o.hsq.empty() && o.hsq.size() != nt)
        throw std::invalid_argument("ERROR: Number of --hsq values does not match the number of traits");
    if (o.cc) {
        if (o.k.empty()) o.k.assign(nt, DEFAULT_K);
        for (float v : o.k)
            if (v <= 0 || v >= 1) throw std::invalid_argument("ERROR: Prevalence --k must be between 0.0 and 1.0");
    }
    if (o.hsq.empty()) o.hsq.assign(nt, DEFAULT_HSQ);
    for (float v : o.hsq)
        if (v < 0 || v > 1) throw std::invalid_argument("ERROR: Heritability --hsq must be between 0.0 and 1.0");

    if (!o.causal_pi.empty() && o.causal_pi.size() != nc)
        throw std::invalid_argument("ERROR: Number of --causal-pi values does not match the number of components");
    if (!o.causal_n.empty() && o.causal_n.size() != nc)
        throw std::invalid_argument("ERROR: Number of --causal-n values does not match the number of components");
    if (!o.causal_variants.empty() && o.causal_variants.size() != nc)
        throw std::invalid_argument("ERROR: Number of --causal-variants values does not match the number of components");
    if (((int)o.causal_pi.empty() + (int)o.causal_n.empty() + (int)o.causal_variants.empty()) < 2)
        throw std::invalid_argument("ERROR: At most one of --causal-pi, --causal-n, --causal-variants options must be used");
    if (o.causal_pi.empty() && o.causal_n.empty() && o.causal_variants.empty()) o.causal_pi.assign(nc, DEFAULT_CAUSAL_PI);
    for (float v : o.causal_pi)
        if (v <= 0 || v > 1) throw std::invalid_argument("ERROR: --causal-pi values must be between 0.0 and 1.0");
    for (int v : o.causal_n)
        if (v <= 0) throw std::invalid_argument("ERROR: --causal-n values must be a positive");

    if (!o.causal_regions.empty() && !o.causal_variants.empty()) {
        log << "WARNING: Option --causal-regions can not be used together with --causal-variants, and will be ignored\n";
        o.causal_regions.clear();
    }
    if (!o.causal_regions.empty() && o.causal_regions.size() != nc)
        throw std::invalid_argument("ERROR: Number of --causal-regions values does not match the number of components");
    for (const auto &v : o.causal_regions)
        if (!file_exists(v)) throw std::runtime_error("ERROR: input file " + v + " does not exist");
    for (const auto &v : o.causal_variants)
        if (!file_exists(v)) throw std::runtime_error("ERROR: input file " + v + " does not exist");

    auto sigsq_default = [&](std::vector<float> &t, const char *name) {
        if (!t.empty() && t.size() != nc)
            throw std::invalid_argument(std::string("ERROR: Number of --") + name + " values does not match the number of components");
        if (t.empty()) t.assign(nc, DEFAULT_SIGSQ);
        for (float v : t)
            if (v < 0) throw std::invalid_argument(std::string("ERROR: --") + name + " must be non-negative");
    };
    sigsq_default(o.trait1_sigsq, "trait1-sigsq");
    if (o.num_traits >= 2) {
        sigsq_default(o.trait2_sigsq, "trait2-sigsq");
        if (!o.rg.empty() && o.rg.size() != nc)
            throw std::invalid_argument("ERROR: Number of --rg values does not match the number of components");
        if (o.rg.empty()) o.rg.assign(nc, DEFAULT_RG);
        for (float v : o.rg)
            if (v < -1.0f || v > 1.0f) throw std::invalid_argument("ERROR: --rg values must be between -1.0 and 1.0");
    }
    for (int t = 0; t < o.num_traits; t++) {
        auto &sp = t == 0 ? o.trait1_s_pow : o.trait2_s_pow;
        if (!sp.empty() && sp.size() != nc)
            throw std::invalid_argument("ERROR: Number of --trait-s-pow values does not match the number of components");
    }
    if (o.num_traits == 1 && !o.trait2_s_pow.empty()) {
        log << "WARNING: Option --trait2-s-pow will be ignored (--num-traits 1)\n";
        o.trait2_s_pow.clear();
    }
    if (o.num_traits == 2 && !o.trait1_s_pow.empty() && o.trait2_s_pow.empty())
        throw std::invalid_argument("ERROR: --trait1-s-pow and --trait2-s-pow must be both specified or absent (when --num-traits=2)");
    if (o.gcta_sigma && (!o.trait1_s_pow.empty() || !o.trait2_s_pow.empty()))
        throw std::invalid_argument("ERROR: --gcta-sigma is incompatible with --trait1-s-pow / --trait2-s-pow");

    for (const auto &bfile : o.bfiles) {            // source.cc:589-600
        log << "Opening " << bfile << "... ";
        PlinkFile pf(bfile);
        log << "done, " << pf.num_samples() << " samples, " << pf.num_loci() << " variants.\n";
        if (o.num_samples < 0) o.num_samples = pf.num_samples();
        if (o.num_samples != pf.num_samples())
            throw std::runtime_error("ERROR: Inconsistent number of subjects in --bfile-chr input");
        pio->push_back(std::move(pf));
    }
    o.num_variants = pio->num_variants();
    if (o.bfiles.size() > 1) log << "Found " << o.num_variants << " variants in total.\n";

    if (!o.causal_variants.empty() || !o.causal_regions.empty()) {
        log << "Validate that all variants in --bfile / --bfile-chr have unique names (required for --causal-variants "
               "and --causal-regions options)...\n";
        pio->find_variant_index_map();
    }

    if (o.cc) {                                      // source.cc:613-628
        if (o.ncas.empty() || o.ncon.empty())
            for (int i = 0; i < o.num_traits; i++) {
                o.ncas.push_back(static_cast<int>(o.k[i] * o.num_samples));
                o.ncon.push_back(o.num_samples - o.ncas[i]);
            }
        for (int i = 0; i < o.num_traits; i++) {
            const int max_cas = static_cast<int>(o.k[i] * o.num_samples), max_con = o.num_samples - max_cas;
            if (o.ncas[i] <= 0 || o.ncas[i] > max_cas)
                throw std::invalid_argument("ERROR: --ncas out of range, must be 0 to k*N, where k is prevalence");
            if (o.ncon[i] <= 0 || o.ncon[i] > max_con)
                throw std::invalid_argument("ERROR: --ncon out of range, must be 0 to (1-k)*N, where k is prevalence");
        }
    }

    if (o.causal_regions.empty()) {                  // source.cc:630-642
        if (!o.causal_pi.empty()) {
            for (int i = 0; i < o.num_components; i++)
                o.causal_n.push_back(static_cast<int>(o.causal_pi[i] * o.num_variants));
            o.causal_pi.clear();
        }
        if (!o.causal_n.empty()) {
            int total = 0;
            for (int i = 0; i < o.num_components; i++) total += o.causal_n[i];
            if (total > o.num_variants)
                throw std::invalid_argument("ERROR: sum of --causal-pi must be less than 1.0, or sum of --causal-n must be "
                                            "less than number of variants present in the input file.");
        }
    } else {                                         // source.cc:643-677
        std::vector<int> region_per_variant(o.num_variants, -1);
        for (int c = 0; c < o.num_components; c++) {
            std::vector<int> idx;
            read_causal_variants_file(o.causal_regions[c], &log, *pio, &idx, nullptr, nullptr);
            for (int v : idx) {
                if (region_per_variant[v] != -1)
                    throw std::runtime_error("ERROR: Variant '" + pio->get_locus(v)->name +
                                             "' is duplicated in --causal-regions input; --causal-regions must be non-overlapping.");
                region_per_variant[v] = c;
            }
            if (!o.causal_pi.empty()) {
                o.causal_n.push_back(static_cast<int>(o.causal_pi[c] * idx.size()));
                if (o.causal_n.back() <= 0) {
                    std::stringstream ss;
                    ss << "--causal-pi " << o.causal_pi[c] << " is too low because --causal-region " << o.causal_regions[c]
                       << " contains only " << idx.size() << " variants.";
                    throw std::runtime_error(ss.str());
                }
            }
            if (o.causal_n[c] > (int)idx.size()) {
                std::stringstream ss;
                ss << "--causal-n " << o.causal_n[c] << " is too high because --causal-region " << o.causal_regions[c]
                   << " contains only " << idx.size() << " variants.";
                throw std::runtime_error(ss.str());
            }
        }
        o.causal_pi.clear();
    }

    if (o.trait2_snp_offset != 0) {                  // source.cc:679-696
        if (o.num_traits != 2) throw std::invalid_argument("ERROR: --trait2-snp-offset is applicable only to --num-traits 2");
        if (o.num_components != 1) throw std::invalid_argument("ERROR: --trait2-snp-offset is supported only for --num-components 1");
        if (!o.causal_regions.empty()) throw std::invalid_argument("ERROR: --trait2-snp-offset is incompatible with --causal-regions");
        if (!o.causal_variants.empty()) throw std::invalid_argument("ERROR: --trait2-snp-offset is incompatible with --causal-variants");
        if (o.rg[0] != DEFAULT_RG) throw std::invalid_argument("ERROR: --trait2-snp-offset is incompatible with --rg");
        o.num_components = 3;
    }
}

void describe_simu_options(const SimuOptions &s, Logger &log)        // source.cc:702-729
{
    log << "Options in effect (after applying default setting to non-specified parameters):\n";
    if (!s.bfile.empty()) log << "\t--bfile " << s.bfile << " \\\n";
    if (!s.bfile_chr.empty()) log << "\t--bfile-chr " << s.bfile_chr << " \\\n";
    if (s.qt) log << "\t--qt \\\n";
    if (s.cc) log << "\t--cc \\\n";
    log << "\t--num-traits " << s.num_traits << " \\\n";
    if (!s.k.empty()) log << "\t--k " << s.k << " \\\n";
    if (!s.ncas.empty()) log << "\t--ncas " << s.ncas << " \\\n";
    if (!s.ncon.empty()) log << "\t--ncon " << s.ncon << " \\\n";
    if (!s.hsq.empty()) log << "\t--hsq " << s.hsq << " \\\n";
    log << "\t--num-components " << s.num_components << " \\\n";
    if (!s.causal_variants.empty()) log << "\t--causal-variants " << s.causal_variants << " \\\n";
    if (!s.causal_n.empty()) log << "\t--causal-n " << s.causal_n << " \\\n";
    if (!s.causal_pi.empty()) log << "\t--causal-pi " << s.causal_pi << " \\\n";
    if (!s.causal_regions.empty()) log << "\t--causal-regions " << s.causal_regions << " \\\n";
    if (!s.trait1_sigsq.empty()) log << "\t--trait1-sigsq " << s.trait1_sigsq << " \\\n";
    if (!s.trait2_sigsq.empty()) log << "\t--trait2-sigsq " << s.trait2_sigsq << " \\\n";
    if (!s.rg.empty()) log << "\t--rg " << s.rg << " \\\n";
    if (!s.out.empty()) log << "\t--out " << s.out << " \\\n";
    log << "\t--seed " << s.seed << " \\\n";
    if (s.gcta_sigma) log << "\t--gcta-sigma \\\n";
    if (!s.trait1_s_pow.empty()) log << "\t--trait1-s-pow " << s.trait1_s_pow << " \\\n";
    if (!s.trait2_s_pow.empty()) log << "\t--trait2-s-pow " << s.trait2_s_pow << " \\\n";
    if (s.norm_effect) log << "\t--norm-effect \\\n";
    log << "\n";
}

// ---------------------------------------------------------------- causal selection

void find_causals_trait2_snp_offset(const SimuOptions &o, Mt19937 &rng, std::vector<int> *comp)   // source.cc:731-757
{
    comp->assign(o.num_variants, -1);
    std::vector<int> perm(o.num_variants);
    for (int i = 0; i < o.num_variants; i++) perm[i] = i;
    random_shuffle(perm, rng);
    for (int c = 0; c < o.causal_n[0]; c++) {
        const int s1 = perm[c], s2 = (s1 + o.trait2_snp_offset) % o.num_variants;
        comp->at(s1) += 1;
        comp->at(s2) += 2;
    }
    for (int i = 0; i < o.num_variants; i++)
        if (comp->at(i) >= 3) throw std::runtime_error("error in simu logic, (component_per_variant->at(i) >= 3)");
}

void find_causals(const SimuOptions &o, Mt19937 &rng, Logger &log, const PlinkFiles &pio, std::vector<int> *comp,
                  std::vector<double> *effect1, std::vector<double> *effect2)                     // source.cc:759-841
{
    comp->assign(o.num_variants, -1);
    if (!o.causal_variants.empty()) {
        bool from_file = false;
        for (int c = 0; c < o.num_components; ++c) {
            std::vector<int> idx;
            std::vector<double> e1, e2;
            read_causal_variants_file(o.causal_variants[c], &log, pio, &idx, &e1, &e2);
            if (c == 0) {
                if (!e1.empty()) effect1->assign(o.num_variants, 0.0);
                if (!e2.empty()) effect2->assign(o.num_variants, 0.0);
                from_file = !e1.empty();
                if (from_file && e2.empty() && o.num_traits == 2)
                    throw std::runtime_error("ERROR: issue with --causal-variants files; effect sizes are not provided for "
                                             "the second trait, file '" + o.causal_variants[0] + "'");
                if (from_file && !e2.empty() && o.num_traits == 1)
                    log << "WARNING: minor issue with --causal-variants files; effect sizes provided for the second trait "
                           "will be ignored, '" << o.causal_variants[0] << "'\n";
            }
            if (from_file && e1.empty())
                throw std::runtime_error("ERROR: issue with --causal-variants files; effect sizes are present in '" +
                                         o.causal_variants[0] + "' but not in '" + o.causal_variants[c] + "'");
            for (size_t i = 0; i < idx.size(); i++) {
                const int v = idx[i];
                if (comp->at(v) != -1)
                    throw std::runtime_error("ERROR: Variant '" + pio.get_locus(v)->name +
                                             "' is duplicated in --causal-variants input");
                comp->at(v) = c;
                if (from_file && !e1.empty()) effect1->at(v) = e1[i];
                if (from_file && !e2.empty()) effect2->at(v) = e2[i];
            }
        }
        return;
    }
    if (!o.causal_regions.empty()) {
        for (int c = 0; c < o.num_components; c++) {
            std::vector<int> idx;
            read_causal_variants_file(o.causal_regions[c], nullptr, pio, &idx, nullptr, nullptr);
            random_shuffle(idx, rng);
            for (int i = 0; i < o.causal_n[c]; i++) comp->at(idx[i]) = c;
        }
        return;
    }
    std::vector<int> perm(o.num_variants);
    for (int i = 0; i < o.num_variants; i++) perm[i] = i;
    random_shuffle(perm, rng);
    int p = 0;
    for (int c = 0; c < o.num_components; c++)
        for (int i = 0; i < o.causal_n[c]; i++, p++) comp->at(perm[p]) = c;
}

// SIMU_B200_MODE: "exact" (the default: per-sample FP64 sums in ascending SNP order, bit-identical to
// source.cc:986-996 on one GPU) or "fast" (tensor-core kernel, exact sum of 40-bit fixed-point weights, ~1e-12
// of max|y|).  Anything else is an error -- a typo must not silently change the arithmetic.
int find_pheno_mode()
{
    const char *m = getenv("SIMU_B200_MODE");
    if (!m || !*m || std::string(m) == "exact") return SIMU_B200_MODE_EXACT;
    if (std::string(m) == "fast") return SIMU_B200_MODE_FAST;
    throw std::runtime_error(std::string("ERROR: SIMU_B200_MODE must be 'exact' or 'fast', got '") + m + "'");
}

// ---------------------------------------------------------------- the GPU session
//
// One context per GPU.  SIMU_B200_DEVICE=n (default 0) selects a single GPU; SIMU_B200_DEVICES=all or a
// list "0,1,2,3" shards the path by SNP over several GPUs of the box: the ascending causal
// list is cut into contiguous slices balanced by causal-row count, every GPU stages and processes its
// slice from its own host thread, frequencies are concatenated and the partial phenotypes summed on the
// host in device order (deterministic; with more than one GPU the sum order differs from the reference's,
// so EXACT mode is then tolerance-matched, not bit-identical).

struct Gpu {
    std::vector<simu_b200_ctx *> ctxs;
    simu_b200_ctx *ctx = nullptr;     // ctxs[0]: the O(N) reductions and the threshold run there
    std::vector<int> causal;          // ascending variant indices staged in the panels, in order
    std::vector<size_t> lo, hi;       // slice [lo[d], hi[d]) of `causal` owned by GPU d
    ~Gpu() { for (simu_b200_ctx *c : ctxs) if (c) simu_b200_close(c); }
    void check(int rc) const { check(ctx, rc); }
    static void check(const simu_b200_ctx *c, int rc)
    {
        if (rc != SIMU_B200_OK) throw std::runtime_error(simu_b200_last_error(c));
    }
    void open(int num_samples)
    {
        std::vector<int> devices;
        if (const char *all = getenv("SIMU_B200_DEVICES")) {
            if (std::string(all) == "all") {
                int n = 0;
                if (simu_b200_device_count(&n) != SIMU_B200_OK) throw std::runtime_error(std::string("ERROR: ") + simu_b200_last_error(nullptr));
                for (int d = 0; d < n; d++) devices.push_back(d);
            } else {
                std::stringstream ss(all);
                std::string tok;
                while (std::getline(ss, tok, ',')) if (!tok.empty()) devices.push_back(atoi(tok.c_str()));
            }
        }
        if (devices.empty()) { const char *d = getenv("SIMU_B200_DEVICE"); devices.push_back(d ? atoi(d) : 0); }
        for (int d : devices) {
            simu_b200_ctx *c = nullptr;
            if (simu_b200_open(&c, d, num_samples) != SIMU_B200_OK)
                throw std::runtime_error(std::string("ERROR: ") + simu_b200_last_error(nullptr));
            ctxs.push_back(c);
        }
        ctx = ctxs[0];
    }
    // run fn(d) for every GPU, each on its own host thread; rethrow the first failure
    template <typename F> void for_each_gpu(F fn) const
    {
        if (ctxs.size() == 1) { fn((size_t)0); return; }
        std::vector<std::thread> pool;
        std::vector<std::string> errors(ctxs.size());
        for (size_t d = 0; d < ctxs.size(); d++)
            pool.emplace_back([&, d] { try { fn(d); } catch (std::exception &e) { errors[d] = e.what(); if (errors[d].empty()) errors[d] = "error"; } });
        for (auto &t : pool) t.join();
        for (const auto &e : errors) if (!e.empty()) throw std::runtime_error(e);
    }
    // stage the causal rows of every bfile, in variant order (the rows pio_next_row would read)
    void stage(const SimuOptions &o, const std::vector<int> &comp, const PlinkFiles &pio)
    {
        if ((int)comp.size() < o.num_variants) {
            std::stringstream ss;
            ss << "ERROR: expect to find " << o.num_variants << " variants across input files; found only " << comp.size()
               << ". Are input Plink files corrupted?";
            throw std::runtime_error(ss.str());
        }
        causal.clear();
        for (int v = 0; v < o.num_variants; v++)
            if (comp[v] != -1) causal.push_back(v);
        const size_t D = ctxs.size(), mc = causal.size();
        lo.assign(D, 0); hi.assign(D, 0);
        for (size_t d = 0, at = 0; d < D; d++) {          // contiguous slices, balanced by causal-row count
            lo[d] = at;
            at += mc / D + (d < mc % D ? 1 : 0);
            hi[d] = at;
        }
        // A GPU's slice normally fits its HBM as ONE panel, staged once and read by both passes.  When it does
        // not (e.g. --causal-pi 1 on 11M variants x 100K samples = 275 GB on one GPU) the slice is cut into
        // panel-sized batches that are staged again for the second pass -- the reference reads the .bed twice
        // as well (source.cc:1235, :1295) -- so memory, not the command line, decides how a run is laid out.
        files = &pio;
        batches.assign(D, {});
        for (size_t d = 0; d < D; d++) {
            int64_t free_b = 0, total_b = 0;
            check(ctxs[d], simu_b200_mem_info(ctxs[d], &free_b, &total_b));
            const int64_t per_group = simu_b200_panel_bytes(ctxs[d], 16, SIMU_B200_LAYOUT_SAMPLE16);
            // headroom: staging buffers, per-sample vectors and 64 B per causal row of weights / frequencies / effects
            int64_t budget = free_b - (int64_t(1) << 30) - 64 * (int64_t)(hi[d] - lo[d]) - 64 * (int64_t)o.num_samples;
            if (const char *cap = getenv("SIMU_B200_MAX_PANEL_MB")) budget = std::min<int64_t>(budget, (int64_t)atoll(cap) << 20);
            const size_t max_rows = (size_t)std::max<int64_t>(1, budget / per_group) * 16;
            for (size_t at = lo[d]; at < hi[d]; at += max_rows) batches[d].push_back({at, std::min(hi[d], at + max_rows)});
            if (batches[d].empty()) batches[d].push_back({lo[d], hi[d]});
        }
        loaded.assign(D, (size_t)-1);
        if (resident()) for_each_gpu([&](size_t d) { load(d, 0); });
    }
    bool resident() const
    {
        for (const auto &b : batches) if (b.size() != 1) return false;
        return true;
    }
    // stage batch b of GPU d's slice into its panel (compact causal panel: only the rows pio_next_row would read,
    // sample-major in groups of sixteen rows -- the layout of the vertical counts and of both find_pheno kernels)
    void load(size_t d, size_t b)
    {
        if (loaded[d] == b) return;
        simu_b200_ctx *c = ctxs[d];
        const size_t blo = batches[d][b].first, bhi = batches[d][b].second;
        size_t largest = 1;
        for (const auto &r : batches[d]) largest = std::max(largest, r.second - r.first);
        if ((size_t)simu_b200_panel_rows(c) < largest || simu_b200_panel_layout(c) != SIMU_B200_LAYOUT_SAMPLE16)
            check(c, simu_b200_panel_alloc_layout(c, (int64_t)largest, SIMU_B200_LAYOUT_SAMPLE16));
        size_t pos = blo;
        int64_t dst = 0;
        for (size_t f = 0; f < files->num_files() && pos < bhi; f++) {
            const int flo = files->offset(f), fhi = flo + files->file(f).num_loci();
            std::vector<int64_t> local;
            while (pos < bhi && causal[pos] < fhi) local.push_back(causal[pos++] - flo);
            if (local.empty()) continue;
            check(c, simu_b200_panel_load_bed(c, (files->file(f).prefix + ".bed").c_str(), files->file(f).num_loci(), local.data(),
                                              (int64_t)local.size(), dst));
            dst += (int64_t)local.size();
        }
        loaded[d] = b;
    }
    const PlinkFiles *files = nullptr;
    std::vector<std::vector<std::pair<size_t, size_t>>> batches;      // per GPU: ranges of `causal`, one panel each
    std::vector<size_t> loaded;                                       // per GPU: the batch its panel holds
};

void find_freq(const SimuOptions &o, const std::vector<int> &comp, const PlinkFiles &pio, Gpu &gpu,
               std::vector<double> *freq_vec)                                                     // source.cc:843-885
{
    freq_vec->assign(o.num_variants, std::numeric_limits<double>::quiet_NaN());
    gpu.stage(o, comp, pio);
    if (gpu.causal.empty()) return;
    std::vector<double> f(gpu.causal.size());
    gpu.for_each_gpu([&](size_t d) {
        for (size_t b = 0; b < gpu.batches[d].size(); b++) {
            const size_t l = gpu.batches[d][b].first, n = gpu.batches[d][b].second - l;
            if (!n) continue;
            gpu.load(d, b);
            Gpu::check(gpu.ctxs[d], simu_b200_find_freq(gpu.ctxs[d], nullptr, (int64_t)n, nullptr, f.data() + l));
        }
    });
    for (size_t j = 0; j < gpu.causal.size(); j++) (*freq_vec)[gpu.causal[j]] = f[j];
}

void find_pheno(const SimuOptions &o, const std::vector<double> &freq_vec, const std::vector<double> &effect1,
                const std::vector<double> &effect2, Gpu &gpu, std::vector<double> *pheno1, std::vector<double> *pheno2)
{                                                                                                // source.cc:951-1006
    const bool bivariate = o.num_traits == 2;
    const size_t mc = gpu.causal.size();
    std::vector<double> f(mc), x1(mc), x2(bivariate ? mc : 0);
    for (size_t j = 0; j < mc; j++) {
        f[j] = freq_vec[gpu.causal[j]];
        x1[j] = effect1[gpu.causal[j]];
        if (bivariate) x2[j] = effect2[gpu.causal[j]];
    }
    pheno1->assign(o.num_samples, 0.0);
    pheno2->clear();
    if (bivariate) pheno2->assign(o.num_samples, 0.0);
    const int mode = find_pheno_mode();
    const size_t D = gpu.ctxs.size();
    // the rows of one GPU, batch by batch in ascending order, summed into (q1, q2): one call when the slice is
    // resident, else one panel load + accumulate per batch (EXACT: the per-sample chain of additions continues)
    auto run = [&](size_t d, double *q1, double *q2) {
        if (gpu.batches[d].size() == 1) {
            const size_t l = gpu.lo[d], n = gpu.hi[d] - l;
            gpu.load(d, 0);
            Gpu::check(gpu.ctxs[d], simu_b200_find_pheno(gpu.ctxs[d], nullptr, (int64_t)n, f.data() + l, x1.data() + l,
                                                         bivariate ? x2.data() + l : nullptr, q1, bivariate ? q2 : nullptr, mode));
            return;
        }
        for (size_t b = 0; b < gpu.batches[d].size(); b++) {
            const size_t l = gpu.batches[d][b].first, n = gpu.batches[d][b].second - l;
            gpu.load(d, b);
            Gpu::check(gpu.ctxs[d], simu_b200_find_pheno_accumulate(gpu.ctxs[d], (int64_t)n, f.data() + l, x1.data() + l,
                                                                    bivariate ? x2.data() + l : nullptr, q1,
                                                                    bivariate ? q2 : nullptr, mode));
        }
    };
    if (D == 1) {
        run(0, pheno1->data(), bivariate ? pheno2->data() : nullptr);
        return;
    }
    // one partial N x K phenotype per GPU, summed in device order (the path's one exchange step)
    std::vector<std::vector<double>> p1(D, std::vector<double>(o.num_samples, 0.0)), p2(D);
    if (bivariate) for (auto &v : p2) v.assign(o.num_samples, 0.0);
    gpu.for_each_gpu([&](size_t d) { run(d, p1[d].data(), bivariate ? p2[d].data() : nullptr); });
    for (size_t d = 0; d < D; d++)
        for (int i = 0; i < o.num_samples; i++) {
            (*pheno1)[i] += p1[d][i];
            if (bivariate) (*pheno2)[i] += p2[d][i];
        }
}

void apply_heritability(float hsq, Mt19937 &rng, Gpu &gpu, std::vector<double> *pheno, std::vector<double> *effect)
{                                                                                                // source.cc:1008-1029
    std::vector<double> noise(pheno->size());
    NormalDraws random_normal(rng);
    for (size_t i = 0; i < noise.size(); i++) noise[i] = random_normal();
    double gen = 0, env = 0;
    gpu.check(simu_b200_apply_heritability(gpu.ctx, hsq, pheno->data(), (int64_t)pheno->size(), noise.data(),
                                           effect->data(), (int64_t)effect->size(), &gen, &env));
}

void apply_liability_threshold(float k, int ncas, int ncon, Mt19937 &rng, Gpu &gpu, std::vector<double> *pheno)
{                                                                                                // source.cc:1031-1055
    const int n = static_cast<int>(pheno->size());
    const int max_ncas = (int)floor(k * n), max_ncon = n - max_ncas;
    std::vector<int> cas_inds, con_inds;
    for (int i = 0; i < max_ncon; i++) con_inds.push_back(i);
    for (int i = max_ncon; i < n; i++) cas_inds.push_back(i);
    random_shuffle(con_inds, rng);
    random_shuffle(cas_inds, rng);
    if ((int)con_inds.size() < ncon) throw std::runtime_error("ERROR: too many controls requested");
    if ((int)cas_inds.size() < ncas) throw std::runtime_error("ERROR: too many cases requested");
    if (ncon == max_ncon && ncas == max_ncas) {
        // every sample of both pools is labelled (the default, source.cc:613-628): the shuffles above only
        // consumed RNG state and the pools follow from the order statistic -- radix select instead of a sort
        std::vector<uint8_t> is_case(n);
        gpu.check(simu_b200_liability_split(gpu.ctx, pheno->data(), n, max_ncon, is_case.data()));
        for (int i = 0; i < n; i++) pheno->at(i) = is_case[i] ? 2 : 1;
        return;
    }
    std::vector<int32_t> index(n);
    gpu.check(simu_b200_liability_order(gpu.ctx, pheno->data(), n, index.data()));
    for (int i = 0; i < n; i++) pheno->at(i) = -9;
    for (int i = 0; i < ncon; i++) pheno->at(index[con_inds[i]]) = 1;
    for (int i = 0; i < ncas; i++) pheno->at(index[cas_inds[i]]) = 2;
}

// ---------------------------------------------------------------- effect sizes (host, RNG order of source.cc:887-949)

void draw_effect_sizes(const SimuOptions &o, Mt19937 &rng, const std::vector<int> &comp, std::vector<double> *e1,
                       std::vector<double> *e2)
{
    e1->assign(o.num_variants, 0.0);
    if (o.num_traits == 2) e2->assign(o.num_variants, 0.0);
    NormalDraws random_normal(rng);
    for (int i = 0; i < o.num_variants; i++) {
        const int c = comp[i];
        if (c == -1) continue;
        if (o.trait2_snp_offset != 0) {                       // :903-924
            const double s1 = sqrt(o.trait1_sigsq[0]), s2 = sqrt(o.trait2_sigsq[0]);
            const double x1 = random_normal(), x2 = random_normal();
            if (c == 0 || c == 2) (*e1)[i] = s1 * x1;
            if (c == 1 || c == 2) (*e2)[i] = s2 * x2;
        } else if (o.num_traits == 1) {                       // :887-901
            (*e1)[i] = sqrt(o.trait1_sigsq[c]) * random_normal();
        } else {                                              // :926-949
            const double s1 = sqrt(o.trait1_sigsq[c]), s2 = sqrt(o.trait2_sigsq[c]), rg = o.rg[c];
            const double x1 = random_normal(), x2 = random_normal();
            (*e1)[i] = s1 * x1;
            (*e2)[i] = s2 * (rg * x1 + sqrt(1 - rg * rg) * x2);
        }
    }
}

// ---------------------------------------------------------------- writers (source.cc:1057-1107)

void save_pheno_file(const SimuOptions &o, const PlinkFiles &pio, const std::vector<double> &p1, const std::vector<double> &p2)
{
    TextWriter file(o.out_pheno);          // same characters as the reference's ofstream << (fastio.h)
    const bool bivariate = o.num_traits == 2;
    file << "FID\tIID\ttrait1" << (bivariate ? "\ttrait2" : "") << "\n";
    for (size_t i = 0; i < p1.size(); i++) {
        const Sample &s = pio.get_sample(i);
        file << s.fid << "\t" << s.iid << "\t" << p1[i];
        if (bivariate) file << "\t" << p2[i];
        file << "\n";
    }
}

void save_causals_file(const SimuOptions &o, const std::string &path, const PlinkFiles &pio, const std::vector<double> &freq,
                       const std::vector<int> &comp, const std::vector<double> &effect_per_variant)
{
    TextWriter file(path);
    file << "SNP\tCHR\tPOS\tA1\tA2\tFRQ";
    for (int i = 0; i < o.num_components; i++) file << "\tBETA_c" << (i + 1);
    file << "\n";
    std::vector<int> causal;
    for (int v = 0; v < o.num_variants; v++)
        if (comp[v] != -1) causal.push_back(v);
    write_rows(file, causal.size(), [&](auto &out, size_t j) {          // one line, exactly as source.cc:1089-1105 prints it
        const int v = causal[j];
        const Locus *l = pio.get_locus(v);
        double effect = effect_per_variant[v];
        if (o.norm_effect) effect *= sqrt(fabs(2 * freq[v] * (1 - freq[v])));
        out << l->name << "\t" << static_cast<int>(l->chromosome) << "\t" << l->bp_position << "\t" << l->allele1 << "\t"
            << l->allele2 << "\t" << freq[v];
        for (int i = 0; i < o.num_components; i++) out << "\t" << ((i == comp[v]) ? effect : 0.0);
        out << "\n";
    });
}

}  // namespace

int main(int argc, char *argv[])
{
    try {
        SimuOptions o;
        o.seed = std::chrono::duration_cast<std::chrono::microseconds>(
                     std::chrono::system_clock::now().time_since_epoch()).count();       // source.cc:101-102
        parse_command_line(argc, argv, &o);
        if (o.help) {                                                                     // source.cc:1176-1207
            std::cerr << help_text();
            std::cerr << "\nExamples:\n\n"
                         "* Simulate a quantitative trait with 1% causal markers, with the heritability of 0.5:\n"
                         "  simu --bfile test --qt --causal-pi 0.01 --hsq 0.5\n\n"
                         "* Simulate a quantitative trait as above, using simulation model from GCTA GWAS Simulation:\n"
                         "  simu --bfile test --qt --causal-pi 0.01 --hsq 0.5 --gcta-sigma --norm-effect \n\n"
                         "* Simulate two quantitative traits with genetic correlation of 0.8 and the heritabilities of 0.2 and 0.6:\n"
                         "  simu --bfile test --qt --causal-pi 0.01 --num-traits 2 --hsq 0.2 0.6 --rg 0.8 \n\n"
                         "* Simulate 500 cases and 500 controls with the heritability of liability of 0.5 and disease prevalence of 0.1:\n"
                         "  simu --bfile test --cc --k 0.1 --ncas 500 --ncon 500 --causal-pi 0.01 --hsq 0.5 \n\n"
                         "* Simulate a quantitative trait with specific location of causal markers:\n"
                         "  simu --bfile test --qt --causal-variants causal.snplist --hsq 0.5\n\n"
                         "* Simulate a quantitative trait where n=10 causal markers are randomly distributed within specific region:\n"
                         "  simu --bfile test --qt --causal-n 10 --causal-regions snplist --hsq 0.5\n\n";
            exit(EXIT_FAILURE);
        }

        Logger log(o.out + ".log");
        log_header(argc, argv, log);
        try {
            const std::time_t started = std::time(nullptr);
            log << "Analysis started: " << now_string(started) << "\n";

            StageTimer timer;
            PlinkFiles pio;
            fix_and_validate(o, log, &pio);
            timer.mark("parse_and_validate");
            describe_simu_options(o, log);

            Mt19937 rng(o.seed);                                                          // source.cc:1222
            std::vector<int> comp;
            std::vector<double> effect1, effect2;
            if (o.trait2_snp_offset == 0) find_causals(o, rng, log, pio, &comp, &effect1, &effect2);
            else find_causals_trait2_snp_offset(o, rng, &comp);

            timer.mark("causal_selection");
            const int mode = find_pheno_mode();          // rejects a bad SIMU_B200_MODE before any work is done
            Gpu gpu;
            gpu.open(o.num_samples);
            timer.mark("cuda_context");
            log << "simu_b200: find_pheno mode " << (mode == SIMU_B200_MODE_EXACT ? "exact" : "fast") << ", "
                << gpu.ctxs.size() << " GPU" << (gpu.ctxs.size() == 1 ? "" : "s (phenotypes summed over SNP shards: tolerance-matched)")
                << "\n";

            log << "Calculate allele frequencies (for causal markers only)... \n";
            std::vector<double> freq_vec;
            find_freq(o, comp, pio, gpu, &freq_vec);
            timer.mark("stage_bed_and_find_freq");

            auto het_or_throw = [&](int v, const char *what) {                            // source.cc:1249-1255
                const double freq = freq_vec[v], het = fabs(2 * freq * (1 - freq));
                if (het < std::numeric_limits<float>::epsilon()) {
                    const Locus *l = pio.get_locus(v);
                    throw std::runtime_error("Causal variant " + (l ? l->name : std::string("rs???")) +
                                             " has zero frequency. Can not apply " + what);
                }
                return het;
            };

            if (effect1.empty()) {                                                        // source.cc:1238-1265
                log << "Generate effect sizes from normal distribution...\n";
                draw_effect_sizes(o, rng, comp, &effect1, &effect2);
                if (o.gcta_sigma || !o.trait1_s_pow.empty()) {
                    log << "Apply --gcta-sigma or --trait1-s-pow, --trait2-s-pow options to effect sizes...\n";
                    for (int v = 0; v < o.num_variants; v++) {
                        if (comp[v] == -1) continue;
                        const double het = het_or_throw(v, "--gcta-sigma or --trait1-s-pow, --trait2-s-pow ");
                        const double s1 = o.gcta_sigma ? -1.0 : o.trait1_s_pow[comp[v]];
                        effect1[v] *= sqrt(pow(het, s1));
                        if (o.num_traits == 2) {
                            const double s2 = o.gcta_sigma ? -1.0 : o.trait2_s_pow[comp[v]];
                            effect2[v] *= sqrt(pow(het, s2));
                        }
                    }
                }
            } else {                                                                      // source.cc:1266-1290
                log << "Use effect sizes provided in --causal-variants file.\n";
                if (o.num_components > 1)
                    log << "WARNING: when --causal-variants contain effect sizes there is no real need to use more than one "
                           "component (--num-components), isn't it so?\n";
                if (o.gcta_sigma || o.count("trait1-sigsq") || o.count("trait2-sigsq") || o.count("rg") ||
                    o.count("trait1-s-pow") || o.count("trait2-s-pow"))
                    log << "WARNING: Options --gcta-sigma, --trait1-sigsq, --trait2-sigsq, --rg, --trait1-s-pow, --trait2-s-pow "
                           "are irrelevant and will be ignored.\n";
                if (o.norm_effect) {
                    log << "NB! due to --norm-effect option, all effect sizes from --causal-variants will be treated as "
                           "per-normalized-genotypes effect sizes\n";
                    for (int v = 0; v < o.num_variants; v++) {
                        if (comp[v] == -1) continue;
                        const double het = het_or_throw(v, "--norm-effect.");
                        effect1[v] /= sqrt(het);
                        if (o.num_traits == 2) effect2[v] /= sqrt(het);
                    }
                } else {
                    log << "NB! without --norm-effect option all effect sizes from --causal-variants will be treated w.r.t. "
                           "additive coded 0,1,2 genotypes\n";
                }
            }

            timer.mark("effect_sizes");
            log << "Calculate phenotypes... \n";
            std::vector<double> pheno1, pheno2;
            find_pheno(o, freq_vec, effect1, effect2, gpu, &pheno1, &pheno2);
            timer.mark("find_pheno");

            apply_heritability(o.hsq[0], rng, gpu, &pheno1, &effect1);                    // source.cc:1299-1300
            if (o.num_traits == 2) apply_heritability(o.hsq[1], rng, gpu, &pheno2, &effect2);

            if (o.cc) {                                                                   // source.cc:1303-1306
                apply_liability_threshold(o.k[0], o.ncas[0], o.ncon[0], rng, gpu, &pheno1);
                if (o.num_traits == 2) apply_liability_threshold(o.k[1], o.ncas[1], o.ncon[1], rng, gpu, &pheno2);
            }

            timer.mark("heritability_and_threshold");
            log << "Save phenotypes to " << o.out_pheno << "...\n";
            save_pheno_file(o, pio, pheno1, pheno2);
            log << "Save causal variants and their effect sizes to " << o.out << ((o.num_traits == 2) ? ".*" : ".1")
                << ".causals...\n";
            save_causals_file(o, o.out_causals[0], pio, freq_vec, comp, effect1);
            if (o.num_traits == 2) save_causals_file(o, o.out_causals[1], pio, freq_vec, comp, effect2);

            timer.mark("write_outputs");
            timer.report();
            const std::time_t finished = std::time(nullptr);
            log << "Analysis finished: " << now_string(finished) << "\n";
            log << "Elapsed time: " << duration_string((long)(finished - started)) << "\n";
        } catch (std::exception &e) {
            log << e.what() << "\n";
            return EXIT_FAILURE;
        }
    } catch (std::exception &e) {
        std::cerr << "Exception  : " << e.what() << "\n";
        return EXIT_FAILURE;
    } catch (...) {
        std::cerr << "Unknown error occurred.";
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}

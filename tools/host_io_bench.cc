// Host-side text I/O of the drop-in CLI measured on the host cores, and
// checked against a plain iostream / std::map restatement of what the reference does
// (source.cc:327-352 name map, :1057-1107 writers; libplinkio's .bim parse):
//     host_io_bench <n_variants> <n_samples> <workdir>
// writes a synthetic .bim / .fam, then times and compares
//   parse .bim            FileText tokeniser        vs  getline + per-field std::string
//   name -> index map     unordered_map<string_view> vs  std::map<std::string,int>
//   .causals / .pheno     TextWriter (to_chars)      vs  ofstream <<
// and prints one JSON line.  Exit code 1 if any output differs.
#include <chrono>
#include <cmath>
#include <fstream>
#include <iostream>
#include <map>
#include <random>

#include "../simu_b200/host/plink_text.h"

using namespace simu;
using clk = std::chrono::steady_clock;
static double secs(clk::time_point a) { return std::chrono::duration<double>(clk::now() - a).count(); }

static bool parse_bim_iostream(const std::string &path, std::vector<Locus> *loci)   // the pre-fastio parser
{
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    while (std::getline(f, line)) {
        std::vector<std::string> t = split_blank(line);
        if (t.empty()) continue;
        if (t.size() > 6) return false;
        Locus l;
        long long v;
        double d;
        if (t.size() > 0) { if (!full_long(t[0], &v)) return false; l.chromosome = (unsigned char)v; }
        if (t.size() > 1) l.name = t[1];
        if (t.size() > 2) { if (!full_double(t[2], &d)) return false; l.position = (float)d; }
        if (t.size() > 3) { if (!full_long(t[3], &v)) return false; l.bp_position = v; }
        if (t.size() > 4) l.allele1 = t[4];
        if (t.size() > 5) l.allele2 = t[5];
        loci->push_back(std::move(l));
    }
    return true;
}

static bool same_file(const std::string &a, const std::string &b)
{
    std::ifstream fa(a, std::ios::binary), fb(b, std::ios::binary);
    std::vector<char> ba(1 << 20), bb(1 << 20);
    while (true) {
        fa.read(ba.data(), ba.size()); fb.read(bb.data(), bb.size());
        if (fa.gcount() != fb.gcount() || memcmp(ba.data(), bb.data(), (size_t)fa.gcount())) return false;
        if (fa.gcount() == 0) return true;
    }
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: host_io_bench n_variants n_samples workdir\n"); return 2; }
    const long m = atol(argv[1]), n = atol(argv[2]);
    const std::string dir = argv[3], prefix = dir + "/hostio";
    {   // fixture: names rs<i>, 22 chromosomes, odd blanks on some lines like real files have
        TextWriter bim(prefix + ".bim"), fam(prefix + ".fam");
        for (long i = 0; i < m; i++) {
            bim << (int)(1 + i * 22 / m) << (i % 7 == 0 ? "\t" : " ") << "rs" << (long long)(i * 7919 % 100000007) << "_" << (long long)i
                << "\t" << (i % 5 ? "0" : "0.25") << " " << (long long)(i + 1) << "\tA\t" << (i % 3 ? "G" : "CCT") << (i % 11 == 0 ? "\r\n" : "\n");
        }
        for (long i = 0; i < n; i++) fam << "id1_" << (long long)i << " id2_" << (long long)i << " 0 0 0 -9\n";
    }
    bool ok = true;
    std::vector<Locus> a, b;
    auto t0 = clk::now();
    ok &= parse_bim_file(prefix + ".bim", &a);
    const double t_parse_new = secs(t0);
    t0 = clk::now();
    ok &= parse_bim_iostream(prefix + ".bim", &b);
    const double t_parse_old = secs(t0);
    ok &= a.size() == b.size() && (long)a.size() == m;
    for (size_t i = 0; ok && i < a.size(); i++)
        ok &= a[i].chromosome == b[i].chromosome && a[i].name == b[i].name && a[i].position == b[i].position &&
              a[i].bp_position == b[i].bp_position && a[i].allele1 == b[i].allele1 && a[i].allele2 == b[i].allele2;
    std::vector<Sample> fam;
    ok &= parse_fam_file(prefix + ".fam", &fam) && (long)fam.size() == n;

    // name map: build + look every 3rd name up (a --causal-variants file)
    t0 = clk::now();
    NameIndex index;                                   // the product's name index (plink_text.h)
    {
        std::vector<std::string_view> names;
        names.reserve(a.size());
        for (const Locus &l : a) names.emplace_back(l.name);
        index.build(names);
    }
    long sum_new = 0;
    for (size_t i = 0; i < a.size(); i += 3) sum_new += index.find(std::string_view(b[i].name));
    ok &= index.find("no_such_variant") == -1;
    const double t_map_new = secs(t0);
    t0 = clk::now();
    std::map<std::string, int> tm;
    for (size_t i = 0; i < b.size(); i++) ok &= tm.insert({b[i].name, (int)i}).second;
    long sum_old = 0;
    for (size_t i = 0; i < b.size(); i += 3) sum_old += tm.find(b[i].name)->second;
    const double t_map_old = secs(t0);
    ok &= sum_new == sum_old;

    // writers: every variant causal, 2 components; phenotypes with case/control style and real values
    std::mt19937_64 rng(1);
    std::normal_distribution<double> nd;
    std::vector<double> freq(m), eff(m), y1(n), y2(n);
    for (long i = 0; i < m; i++) { freq[i] = (double)(rng() % 200001) / 200000.0; eff[i] = nd(rng) * std::pow(10.0, (double)(rng() % 9) - 4.0); }
    for (long i = 0; i < n; i++) { y1[i] = nd(rng) * 3.0; y2[i] = (i % 3 == 0) ? -9 : (i % 3); }
    t0 = clk::now();
    {
        TextWriter file(prefix + ".new.causals");
        file << "SNP\tCHR\tPOS\tA1\tA2\tFRQ\tBETA_c1\tBETA_c2\n";
        write_rows(file, (size_t)m, [&](auto &out, size_t v) {          // the product's writer path (fastio.h)
            const Locus &l = a[v];
            out << l.name << "\t" << static_cast<int>(l.chromosome) << "\t" << l.bp_position << "\t" << l.allele1 << "\t"
                << l.allele2 << "\t" << freq[v];
            for (int c = 0; c < 2; c++) out << "\t" << ((c == (int)(v & 1)) ? eff[v] : 0.0);
            out << "\n";
        });
        TextWriter ph(prefix + ".new.pheno");
        ph << "FID\tIID\ttrait1\ttrait2\n";
        for (long i = 0; i < n; i++) ph << fam[i].fid << "\t" << fam[i].iid << "\t" << y1[i] << "\t" << y2[i] << "\n";
    }
    const double t_write_new = secs(t0);
    t0 = clk::now();
    {
        std::ofstream file(prefix + ".old.causals");
        file << "SNP\tCHR\tPOS\tA1\tA2\tFRQ\tBETA_c1\tBETA_c2\n";
        for (long v = 0; v < m; v++) {
            const Locus &l = b[v];
            file << l.name << "\t" << static_cast<int>(l.chromosome) << "\t" << l.bp_position << "\t" << l.allele1 << "\t"
                 << l.allele2 << "\t" << freq[v];
            for (int c = 0; c < 2; c++) file << "\t" << ((c == (int)(v & 1)) ? eff[v] : 0.0);
            file << "\n";
        }
        std::ofstream ph(prefix + ".old.pheno");
        ph << "FID\tIID\ttrait1\ttrait2\n";
        for (long i = 0; i < n; i++) ph << fam[i].fid << "\t" << fam[i].iid << "\t" << y1[i] << "\t" << y2[i] << "\n";
    }
    const double t_write_old = secs(t0);
    ok &= same_file(prefix + ".new.causals", prefix + ".old.causals") && same_file(prefix + ".new.pheno", prefix + ".old.pheno");
    for (const char *ext : {".bim", ".fam", ".new.causals", ".old.causals", ".new.pheno", ".old.pheno"}) remove((prefix + ext).c_str());
    printf("{\"n_variants\": %ld, \"n_samples\": %ld, \"identical\": %s, \"parse_bim_s\": {\"fastio\": %.3f, \"iostream\": %.3f}, "
           "\"name_map_s\": {\"hash\": %.3f, \"std_map\": %.3f}, \"writers_s\": {\"fastio\": %.3f, \"ofstream\": %.3f}}\n",
           m, n, ok ? "true" : "false", t_parse_new, t_parse_old, t_map_new, t_map_old, t_write_new, t_write_old);
    return ok ? 0 : 1;
}

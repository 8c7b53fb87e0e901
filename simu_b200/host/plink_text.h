// .bim / .fam text parsing and the concatenated fileset view, for the C++ host driver.
//
// Behaviour follows the reference's vendored libplinkio (paths under
// /root/reference/libplinkio/src): fields split on runs of blanks or tabs
// (bim_parse.c:67-71, libcsv skips leading blanks and empty lines), an integer chromosome
// parsed with strtol and narrowed to unsigned char (bim_parse.c:112-125), float genetic and
// long long bp positions that must parse completely (:139-175), non-empty strings
// (:85-100), a 7th field is an error (:224-226); .fam: sex must be one character
// (fam_parse.c:113-137), phenotype 1/2/0/-9 or a number (:149-199).  Any error makes the
// open fail the way pio_open does: "ERROR: Could not open <prefix>" (source.cc:190-194).
// The fileset view mirrors PioFile / PioFiles (source.cc:187-361).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "fastio.h"

namespace simu {

struct Locus {
    unsigned char chromosome = 0;
    std::string name;
    float position = 0.f;
    long long bp_position = 0;
    std::string allele1, allele2;
};

struct Sample {
    std::string fid, iid;
};

inline std::vector<std::string> split_blank(const std::string &line)
{
    std::vector<std::string> out;
    size_t i = 0;
    const size_t n = line.size();
    while (i < n) {
        while (i < n && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r')) i++;
        size_t j = i;
        while (j < n && line[j] != ' ' && line[j] != '\t' && line[j] != '\r') j++;
        if (j > i) out.emplace_back(line, i, j - i);
        i = j;
    }
    return out;
}

// One .bim line into *l.  Returns 0 for a blank line, 1 for a record, -1 for a malformed line.
inline int parse_bim_line(char *b, char *e, Locus *l)
{
    std::string_view t[6];
    const int n = FileText::split(b, e, t, 6);
    if (n == 0) return 0;
    if (n > 6) return -1;
    long long v;
    double d;
    // rows with fewer than six fields are kept with defaults, as libplinkio's new_row does
    if (n > 0) { if (!full_long(t[0], &v)) return -1; l->chromosome = (unsigned char)v; }
    if (n > 1) l->name.assign(t[1]);
    if (n > 2) { if (!full_double(t[2], &d)) return -1; l->position = (float)d; }
    if (n > 3) { if (!full_long(t[3], &v)) return -1; l->bp_position = v; }
    if (n > 4) l->allele1.assign(t[4]);
    if (n > 5) l->allele2.assign(t[5]);
    return 1;
}

inline bool parse_bim_file(const std::string &path, std::vector<Locus> *loci)
{
    FileText f(path);                      // whole file in memory, fields tokenised in place
    if (!f.ok()) return false;
    // big files: the text is cut at line boundaries and the pieces are parsed by several threads --
    // pass 1 counts the records of each piece, pass 2 parses them straight into their final slots
    const int parts = f.size() > (4u << 20) ? text_threads() : 1;
    const std::vector<size_t> cuts = f.cut_at_lines(parts);
    std::vector<size_t> count((size_t)parts, 0);
    parallel_for_threads(parts, [&](int p) {
        size_t pos = cuts[(size_t)p];
        char *b, *e;
        std::string_view one[1];
        while (f.next_line_in(&pos, cuts[(size_t)p + 1], &b, &e))
            if (FileText::split(b, e, one, 1) > 0) count[(size_t)p]++;
    });
    std::vector<size_t> first((size_t)parts + 1, loci->size());
    for (int p = 0; p < parts; p++) first[(size_t)p + 1] = first[(size_t)p] + count[(size_t)p];
    loci->resize(first[(size_t)parts]);
    std::vector<char> bad((size_t)parts, 0);
    parallel_for_threads(parts, [&](int p) {
        size_t pos = cuts[(size_t)p], at = first[(size_t)p];
        char *b, *e;
        while (f.next_line_in(&pos, cuts[(size_t)p + 1], &b, &e)) {
            const int rc = parse_bim_line(b, e, &(*loci)[at < loci->size() ? at : 0]);
            if (rc < 0) { bad[(size_t)p] = 1; return; }
            at += (size_t)rc;
        }
    });
    for (char x : bad) if (x) return false;
    return true;
}

inline bool parse_fam_file(const std::string &path, std::vector<Sample> *samples)
{
    FileText f(path);
    if (!f.ok()) return false;
    char *b, *e;
    std::string_view t[6];
    samples->reserve(samples->size() + f.count_lines());
    while (f.next_line(&b, &e)) {
        const int n = FileText::split(b, e, t, 6);
        if (n == 0) continue;
        if (n > 6) return false;
        if (n > 4 && t[4].size() != 1) return false;                        // parse_sex
        if (n > 5) {                                                         // parse_phenotype
            double d;
            const std::string_view p = t[5];
            const bool code = p.size() == 1 && (p[0] == '0' || p[0] == '1' || p[0] == '2');
            const bool miss = strncmp(std::string(p).c_str(), "-9", p.size()) == 0;
            if (!code && !miss && !full_double(p, &d)) return false;
        }
        Sample s;
        s.fid.assign(t[0]);
        if (n > 1) s.iid.assign(t[1]);
        samples->push_back(std::move(s));
    }
    return true;
}

// Variant name -> index (PioFiles::snp_to_id_, source.cc:327-352).  The reference's std::map costs
// O(log M) string compares per insert and lookup; this is a hash map keyed by views of the loci's own
// names, in `shards` independent tables (shard = hash mod shards) that are filled by as many threads --
// every thread hashes all names and inserts its own.  A duplicate is reported as the sequential build
// would: the name whose SECOND occurrence comes first.
class NameIndex {
public:
    // names[v] must stay alive and in place while the index is used
    void build(const std::vector<std::string_view> &names)
    {
        const int shards = names.size() > 4000000 ? text_threads() : 1;     // below that one table is faster (0.3 s per 1M names)
        maps_.assign((size_t)shards, {});
        std::vector<long long> dup((size_t)shards, -1);
        parallel_for_threads(shards, [&](int t) {
            auto &m = maps_[(size_t)t];
            m.reserve(names.size() / (size_t)shards * 2 + 16);
            const std::hash<std::string_view> h;
            for (size_t v = 0; v < names.size(); v++) {
                if (shards > 1 && h(names[v]) % (size_t)shards != (size_t)t) continue;
                if (!m.emplace(names[v], (int)v).second) { dup[(size_t)t] = (long long)v; return; }
            }
        });
        long long first = -1;
        for (long long d : dup) if (d >= 0 && (first < 0 || d < first)) first = d;
        if (first >= 0) throw std::runtime_error("ERROR: duplicated variant name: " + std::string(names[(size_t)first]));
    }
    bool empty() const { return maps_.empty(); }
    // -1 when absent
    int find(std::string_view name) const
    {
        const auto &m = maps_[maps_.size() > 1 ? std::hash<std::string_view>()(name) % maps_.size() : 0];
        auto it = m.find(name);
        return it == m.end() ? -1 : it->second;
    }

private:
    std::vector<std::unordered_map<std::string_view, int>> maps_;
};

// One bfile prefix (PioFile, source.cc:187-267).
struct PlinkFile {
    std::string prefix;
    std::vector<Locus> loci;
    std::vector<Sample> samples;

    explicit PlinkFile(const std::string &bfile) : prefix(bfile)
    {
        const std::string fail = "ERROR: Could not open " + bfile;
        if (!parse_fam_file(bfile + ".fam", &samples)) throw std::runtime_error(fail);
        if (!parse_bim_file(bfile + ".bim", &loci)) throw std::runtime_error(fail);
        FILE *fp = fopen((bfile + ".bed").c_str(), "rb");
        if (!fp) throw std::runtime_error(fail);
        unsigned char h[3];
        const size_t got = fread(h, 1, 3, fp);
        fclose(fp);
        if (got != 3) throw std::runtime_error(fail);                      // bed.c:52-58
        // bed_header.c:80-102: v1.00 = magic 6c 1b + order byte; v0.99 = order byte first; older = sample-major
        const bool v100 = h[0] == 0x6c && h[1] == 0x1b;
        const bool v099 = !v100 && (h[0] & ~0x01) == 0;
        const bool snp_major = v100 ? (h[2] == 0x01) : (v099 ? h[0] == 0x01 : false);
        if (!snp_major)                                                     // source.cc:196-200
            throw std::runtime_error("ERROR: Unsupported format in " + bfile +
                                     ", snps must be rows, samples must be columns");
        if (samples.empty())                                                // source.cc:202-206
            throw std::runtime_error("ERROR: Unsupported format in " + bfile + ", rows appears to have zero size");
    }
    int num_samples() const { return (int)samples.size(); }
    int num_loci() const { return (int)loci.size(); }
};

// Concatenation in label order (PioFiles, source.cc:269-361).
class PlinkFiles {
public:
    void push_back(PlinkFile &&f)
    {
        offsets_.push_back(num_variants_);
        num_variants_ += f.num_loci();
        files_.push_back(std::move(f));
    }
    int num_variants() const { return num_variants_; }
    size_t num_files() const { return files_.size(); }
    const PlinkFile &file(size_t i) const { return files_[i]; }
    int offset(size_t i) const { return offsets_[i]; }
    const Sample &get_sample(size_t i) const { return files_[0].samples[i]; }     // source.cc:309-311
    const Locus *get_locus(int variant_index) const                                // source.cc:313-320
    {
        for (size_t i = 0; i < files_.size(); i++) {
            if (variant_index < files_[i].num_loci()) return &files_[i].loci[variant_index];
            variant_index -= files_[i].num_loci();
        }
        return nullptr;
    }
    void find_variant_index_map()                                                  // source.cc:327-343
    {
        if (!snp_to_id_.empty())
            throw std::runtime_error("internal error - find_variant_index_map must be called once");
        // keyed by views of the loci's own names (the loci never move once all files are pushed)
        std::vector<std::string_view> names;
        names.reserve((size_t)num_variants_);
        for (const PlinkFile &f : files_)
            for (const Locus &l : f.loci) names.emplace_back(l.name);
        snp_to_id_.build(names);
    }
    int get_locus_id(const std::string &name) const                                // source.cc:345-352
    {
        const int id = snp_to_id_.empty() ? -1 : snp_to_id_.find(std::string_view(name));
        if (id < 0)
            throw std::runtime_error("ERROR: locus with name " + name +
                                     " does not exist in the input; check your --bfile and/or --causal-variants, "
                                     "--causal-regions arguments");
        return id;
    }

private:
    std::vector<PlinkFile> files_;
    std::vector<int> offsets_;
    NameIndex snp_to_id_;
    int num_variants_ = 0;
};

}  // namespace simu

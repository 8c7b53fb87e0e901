// parse_bim_file with several threads (text cut at line boundaries, two passes) must give exactly what one
// thread gives: blank lines, leading blanks, CRLF, short rows, a last line without newline; a malformed
// line anywhere makes the whole parse fail (pio_open would: bim_parse.c:112-175, :224-226).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "../../simu_b200/host/plink_text.h"

using namespace simu;

// the getline + per-field std::string parser that fastio replaced (kept here as the differential reference)
static bool parse_bim_iostream(const std::string &path, std::vector<Locus> *loci)
{
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    while (std::getline(f, line)) {
        std::vector<std::string> t = split_blank(line);
        if (t.empty()) continue;
        if (t.size() > 6) return false;
        Locus l;
        char *end = nullptr;
        if (t.size() > 0) { long long v = strtoll(t[0].c_str(), &end, 10); if (!end || *end) return false; l.chromosome = (unsigned char)v; }
        if (t.size() > 1) l.name = t[1];
        if (t.size() > 2) { double d = strtod(t[2].c_str(), &end); if (!end || *end) return false; l.position = (float)d; }
        if (t.size() > 3) { long long v = strtoll(t[3].c_str(), &end, 10); if (!end || *end) return false; l.bp_position = v; }
        if (t.size() > 4) l.allele1 = t[4];
        if (t.size() > 5) l.allele2 = t[5];
        loci->push_back(std::move(l));
    }
    return true;
}

static bool same(const std::vector<Locus> &a, const std::vector<Locus> &b)
{
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++)
        if (a[i].chromosome != b[i].chromosome || a[i].name != b[i].name || memcmp(&a[i].position, &b[i].position, sizeof(float)) != 0 ||
            a[i].bp_position != b[i].bp_position || a[i].allele1 != b[i].allele1 || a[i].allele2 != b[i].allele2)
            return false;
    return true;
}

int main(int argc, char **argv)
{
    const std::string path = argv[1];
    const long m = 250000;
    long records = 0;
    {
        TextWriter w(path);
        for (long i = 0; i < m; i++) {
            if (i % 1000 == 17) w << "\n";                               // blank line
            if (i % 1000 == 18) w << "   \t \r\n";                       // blank line with blanks
            if (i % 777 == 5) { w << "  7 short_" << (long long)i << "\n"; records++; continue; }   // fewer than six fields
            w << (i % 13 == 0 ? " \t" : "") << (int)(1 + i % 22) << "\t" << "rs" << (long long)(i * 31) << (i % 9 ? " " : "\t\t")
              << (i % 4 ? "0" : "1.5e-3") << " " << (long long)(i + 1) << " A " << (i % 3 ? "G" : "ACGT") << (i % 11 ? "\n" : "\r\n");
            records++;
        }
        w << "22 last_no_newline 0 99 C T";
        records++;
    }
    std::vector<Locus> one, many;
    setenv("SIMU_B200_IO_THREADS", "1", 1);
    const bool ok1 = parse_bim_file(path, &one);
    setenv("SIMU_B200_IO_THREADS", "5", 1);
    const bool ok5 = parse_bim_file(path, &many);
    bool ok = ok1 && ok5 && (long)one.size() == records && same(one, many) && many.back().name == "last_no_newline" &&
              many[5].name == "short_5" && many[5].chromosome == 7 && many[5].bp_position == 0;
    // a malformed line in the last third: both fail
    {
        FILE *fp = fopen(path.c_str(), "ab");
        fputs("\n3 bad_pos 0 12x A G\n4 fine 0 5 A G\n", fp);
        fclose(fp);
    }
    std::vector<Locus> x, y;
    setenv("SIMU_B200_IO_THREADS", "1", 1);
    ok = ok && !parse_bim_file(path, &x);
    setenv("SIMU_B200_IO_THREADS", "5", 1);
    ok = ok && !parse_bim_file(path, &y);
    // differential fuzz: random token soup, fastio parser vs the iostream parser (accept/reject and every field)
    {
        unsigned long long rs = 88172645463325252ull;
        auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; };
        const char *tokens[] = {"1", "22", "0", "-3", "+7", "255", "256", "x", "rs12", "1e3", "0.5", ".", "-", "--5", "0x10", "12x",
                                "99999999999999999999", "123456789012345678", "nan", "inf", "1.5e-3", "A", "ACGT", "", "7.", "-0"};
        const char *seps[] = {" ", "\t", "  ", " \t ", "\t\t"};
        int fuzz_bad = 0, accepted = 0;
        for (int it = 0; it < 3000; it++) {
            {
                TextWriter w(path);
                const int lines = 1 + (int)(rnd() % 6);
                for (int l = 0; l < lines; l++) {
                    const int nt = (int)(rnd() % 8);
                    if (rnd() % 5 == 0) w << seps[rnd() % 5];
                    for (int t = 0; t < nt; t++) {
                        // mostly well-formed columns so that a good share of the files is accepted
                        const char *tok = (rnd() % 4) ? (t == 0 ? "3" : t == 2 ? "0" : t == 3 ? "1234" : "rs1") : tokens[rnd() % 26];
                        w << tok << seps[rnd() % 5];
                    }
                    w << (rnd() % 4 == 0 ? "\r\n" : "\n");
                }
            }
            std::vector<Locus> fa, fb;
            const bool ra = parse_bim_file(path, &fa), rb = parse_bim_iostream(path, &fb);
            if (ra != rb || (ra && !same(fa, fb))) { if (fuzz_bad < 5) { FileText ft(path); fprintf(stderr, "DISAGREE fast=%d old=%d nfa=%zu nfb=%zu: [%.*s]\n", ra, rb, fa.size(), fb.size(), (int)ft.size(), ft.begin()); } fuzz_bad++; }
            accepted += ra;
        }
        printf("fuzz: 3000 files, %d accepted by both, %d disagreements\n", accepted, fuzz_bad);
        ok = ok && fuzz_bad == 0 && accepted > 100;
    }
    printf("%ld records, threaded == single: %s\n", records, ok ? "yes" : "NO");
    remove(path.c_str());
    return ok ? 0 : 1;
}

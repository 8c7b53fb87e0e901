// parse_bim_file with several threads (text cut at line boundaries, two passes) must give exactly what one
// thread gives: blank lines, leading blanks, CRLF, short rows, a last line without newline; a malformed
// line anywhere makes the whole parse fail (pio_open would: bim_parse.c:112-175, :224-226).
#include <cstdio>
#include <cstdlib>

#include "../../simu_b200/host/plink_text.h"

using namespace simu;

static bool same(const std::vector<Locus> &a, const std::vector<Locus> &b)
{
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++)
        if (a[i].chromosome != b[i].chromosome || a[i].name != b[i].name || a[i].position != b[i].position ||
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
    printf("%ld records, threaded == single: %s\n", records, ok ? "yes" : "NO");
    remove(path.c_str());
    return ok ? 0 : 1;
}

// ref_source_harness.cc -- the UNMODIFIED reference source, /root/reference/src/source.cc, included
// verbatim at build time (never copied into this repository) and compiled against
// oracle/boost_shim/, so that its literal find_freq / find_pheno / apply_heritability /
// apply_liability_threshold / main can be called in-process.  TEST INFRASTRUCTURE ONLY:
// parity anchor for the oracle restatement and for the CUDA path, and the CPU baseline
// (`kind: "reference"`) of bench.py.  Built by oracle/Makefile into oracle/_ref/.
#define main simu_reference_main
#include SIMU_REF_SOURCE
#undef main

#include <chrono>

static std::string g_ref_error;

extern "C" {

const char *ref_src_last_error() { return g_ref_error.c_str(); }

// the reference's main(), argv[0] included
int ref_src_main(int argc, char **argv) { return simu_reference_main(argc, argv); }

// find_freq (source.cc:843-885) then find_pheno (:951-1006) on one bfile, exactly as main() calls
// them (:1235, :1295).  comp = component_per_variant (-1 non-causal); freq/effect arrays have one
// entry per variant.  secs[0], secs[1] receive the wall time of the two calls.
int ref_src_freq_pheno(const char *prefix, int num_traits, const int *comp, long long m, double *freq_out,
                       const double *e1, const double *e2, double *p1, double *p2, double *secs)
{
    try {
        SimuOptions o;
        PioFiles pio;
        auto pf = std::make_shared<PioFile>(std::string(prefix));
        pio.push_back(pf);
        o.num_samples = pf->num_samples();
        o.num_variants = pio.num_variants();
        o.num_traits = num_traits;
        if (m != o.num_variants) { g_ref_error = "variant count mismatch"; return -5; }
        std::vector<int> c(comp, comp + m);
        std::vector<double> freq, x1(e1, e1 + m), x2, y1, y2;
        if (num_traits == 2) x2.assign(e2, e2 + m);
        const auto t0 = std::chrono::steady_clock::now();
        find_freq(o, c, &pio, &freq);
        const auto t1 = std::chrono::steady_clock::now();
        find_pheno(o, c, freq, x1, x2, &pio, &y1, &y2);
        const auto t2 = std::chrono::steady_clock::now();
        for (long long i = 0; i < m; i++) freq_out[i] = freq[i];
        for (size_t i = 0; i < y1.size(); i++) p1[i] = y1[i];
        if (num_traits == 2) for (size_t i = 0; i < y2.size(); i++) p2[i] = y2[i];
        if (secs) {
            secs[0] = std::chrono::duration<double>(t1 - t0).count();
            secs[1] = std::chrono::duration<double>(t2 - t1).count();
        }
        return 0;
    } catch (std::exception &e) {
        g_ref_error = e.what();
        return -1;
    }
}

// apply_heritability (source.cc:1008-1029) with a generator seeded by `seed`; the noise it draws is
// returned so that callers can feed the same draws to the path under test.
int ref_src_apply_heritability(float hsq, long long seed, double *pheno, long long n, double *effect, long long m,
                               double *noise_out)
{
    try {
        boost::mt19937 rng(seed);
        if (noise_out) {                       // the draws apply_heritability is about to make (:1014-1016)
            boost::mt19937 peek(seed);
            boost::variate_generator<boost::mt19937 &, boost::normal_distribution<> > random_normal(peek, boost::normal_distribution<>());
            for (long long i = 0; i < n; i++) noise_out[i] = random_normal();
        }
        std::vector<double> p(pheno, pheno + n), e(effect, effect + m);
        apply_heritability(hsq, rng, &p, &e);
        for (long long i = 0; i < n; i++) pheno[i] = p[i];
        for (long long i = 0; i < m; i++) effect[i] = e[i];
        return 0;
    } catch (std::exception &ex) {
        g_ref_error = ex.what();
        return -1;
    }
}

// apply_liability_threshold (source.cc:1031-1055) with a generator seeded by `seed`
int ref_src_apply_liability_threshold(float k, int ncas, int ncon, long long seed, double *pheno, long long n)
{
    try {
        boost::mt19937 rng(seed);
        std::vector<double> p(pheno, pheno + n);
        apply_liability_threshold(k, ncas, ncon, rng, &p);
        for (long long i = 0; i < n; i++) pheno[i] = p[i];
        return 0;
    } catch (std::exception &ex) {
        g_ref_error = ex.what();
        return -1;
    }
}

}  // extern "C"

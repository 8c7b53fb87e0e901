// placeholder until the LUT kernel lands (next commit): FAST mode reports an error, it never
// silently falls back to another path.
#include "ctx.cuh"
int launch_pheno_lut(simu_b200_ctx *ctx, const int32_t *, int64_t, const double *, const double *,
                     const double *, double *, double *)
{
    return ctx_fail(ctx, SIMU_B200_ERROR, "find_pheno: FAST mode kernel not built yet");
}

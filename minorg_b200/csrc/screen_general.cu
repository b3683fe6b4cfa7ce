// K4a/K5: general screen (patterns, gaps, PAM proximity).  Placeholder until the verifier lands.
#include "common.cuh"
namespace mg {
int launch_screen_general(const mg_genome*, const mg_queries*, const mg_criteria*, const mg_pam*, int64_t, int64_t,
                          uint32_t*, cudaStream_t) {
    set_error("general screen (pattern / gaps / PAM proximity) not built yet");
    return MG_ERR_UNSUPPORTED;
}
}  // namespace mg

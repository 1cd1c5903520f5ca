// sgdx_exact_words.cu -- one instantiation set of sgdx_match_exact_first per read length in words; compiled
// eight times with -DSGDX_LW=1..8 (Makefile) so the objects build in parallel and each kernel keeps the register
// allocation ptxas gives it on its own.
#include "sgdx_exact_first.cuh"

#ifndef SGDX_LW
#error "compile with -DSGDX_LW=<words per read, 1..8>"
#endif

namespace sgdx {
#define SGDX_CAT2(a, b) a##b
#define SGDX_CAT(a, b) SGDX_CAT2(a, b)
cudaError_t SGDX_CAT(launch_exact_words_, SGDX_LW)(const DevParams &p, uint32_t mode, int sms, cudaStream_t stream) {
    return launch_exact_words<SGDX_LW>(p, mode, sms, stream);
}
}  // namespace sgdx

// ssw_score_c.cu -- score kernel instances for strip heights 11..13 (ssw_score_impl.cuh).  The instances are
// spread over four translation units only so that they compile in parallel.
#include "ssw_score_impl.cuh"

namespace sswb {

cudaError_t launch_score_c(int K, bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st)
{
    switch (K) {
        case 11: return launch_k<11>(a, trunc, rev, blocks, st);
        case 12: return launch_k<12>(a, trunc, rev, blocks, st);
        case 13: return launch_k<13>(a, trunc, rev, blocks, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sswb

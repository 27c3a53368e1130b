// ssw_score_a.cu -- score kernel instances for strip heights 1..6 (ssw_score_impl.cuh).  The instances are
// spread over four translation units only so that they compile in parallel.
#include "ssw_score_impl.cuh"

namespace sswb {

cudaError_t launch_score_a(int K, bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st)
{
    switch (K) {
        case 1: return launch_k<1>(a, trunc, rev, blocks, st);
        case 2: return launch_k<2>(a, trunc, rev, blocks, st);
        case 3: return launch_k<3>(a, trunc, rev, blocks, st);
        case 4: return launch_k<4>(a, trunc, rev, blocks, st);
        case 5: return launch_k<5>(a, trunc, rev, blocks, st);
        case 6: return launch_k<6>(a, trunc, rev, blocks, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sswb

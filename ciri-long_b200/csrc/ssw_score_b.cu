// ssw_score_b.cu -- score kernel instances for strip heights 7..10 (ssw_score_impl.cuh).  The instances are
// spread over four translation units only so that they compile in parallel.
#include "ssw_score_impl.cuh"

namespace sswb {

cudaError_t launch_score_b(int K, bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st)
{
    switch (K) {
        case 7: return launch_k<7>(a, trunc, rev, blocks, st);
        case 8: return launch_k<8>(a, trunc, rev, blocks, st);
        case 9: return launch_k<9>(a, trunc, rev, blocks, st);
        case 10: return launch_k<10>(a, trunc, rev, blocks, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sswb

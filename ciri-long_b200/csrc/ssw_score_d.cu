// ssw_score_d.cu -- score kernel instances for strip heights 14..16 (ssw_score_impl.cuh).  The instances are
// spread over four translation units only so that they compile in parallel.
#include "ssw_score_impl.cuh"

namespace sswb {

cudaError_t launch_score_d(int K, bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st)
{
    switch (K) {
        case 14: return launch_k<14>(a, trunc, rev, blocks, st);
        case 15: return launch_k<15>(a, trunc, rev, blocks, st);
        case 16: return launch_k<16>(a, trunc, rev, blocks, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sswb

// ssw_score.cu -- launch table of the score kernels (forward and reverse score pass, ssw_score_impl.cuh).
#include "ssw_common.cuh"
#include "ssw_kernels.h"

namespace sswb {

cudaError_t launch_score_a(int K, bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st);   // K 1..6
cudaError_t launch_score_b(int K, bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st);   // K 7..10
cudaError_t launch_score_c(int K, bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st);   // K 11..13
cudaError_t launch_score_d(int K, bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st);   // K 14..16

cudaError_t launch_score(int K, bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st)
{
    if (K >= 1 && K <= 6) return launch_score_a(K, trunc, rev, a, blocks, st);
    if (K <= 10) return launch_score_b(K, trunc, rev, a, blocks, st);
    if (K <= 13) return launch_score_c(K, trunc, rev, a, blocks, st);
    if (K <= 16) return launch_score_d(K, trunc, rev, a, blocks, st);
    return cudaErrorInvalidValue;
}

}  // namespace sswb

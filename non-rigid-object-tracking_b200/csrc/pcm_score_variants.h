// pcm_score_variants.h -- table of the K1 (score_kernel) instantiations.
// One translation unit per tile height (pcm_score_inst.cu compiled with -DPCM_PPT=6 / 7 / 8) so that they
// build in parallel; pcm_api.cu picks a variant per launch.
#pragma once
#include "pcm_score.cuh"

namespace pcm {

typedef void (*ScoreFn)(const CUtensorMap, const ScoreArgs);
struct ScoreVariant {
    bool smem;          // forests staged in shared memory (else read through L1)
    int depth;          // walk depth fixed at compile time (config.yaml:26 -> 5; benchmark.py:44 -> 7, 10), 0 = run-time
    int ppt;            // rows per thread; tile height = ROW_GROUPS * ppt
    ScoreFn fn;
    const char* name;
};
constexpr int N_SCORE_VARIANTS_PER_PPT = 8;     // {smem, L1} x {5, 7, 10, run-time}
const ScoreVariant* score_variants_ppt6();
const ScoreVariant* score_variants_ppt7();
const ScoreVariant* score_variants_ppt8();

}  // namespace pcm

// pcm_score_inst.cu -- the K1 instantiations of ONE tile height (-DPCM_PPT=6 | 7 | 8), see pcm_score_variants.h
#include "pcm_score_variants.h"

#ifndef PCM_PPT
#error "compile with -DPCM_PPT=6, 7 or 8"
#endif
#define PCM_STR2(x) #x
#define PCM_STR(x) PCM_STR2(x)
#define PCM_CAT2(a, b) a##b
#define PCM_CAT(a, b) PCM_CAT2(a, b)
#define PCM_SV(S, D) {S, D, PCM_PPT, score_kernel<S, D, PCM_PPT>, "score_kernel<" #S "," #D "," PCM_STR(PCM_PPT) ">"}

namespace pcm {

static_assert(PCM_PPT >= MIN_PPT && PCM_PPT <= MAX_PPT, "tile height outside the range the shared-memory layout is sized for");

const ScoreVariant* PCM_CAT(score_variants_ppt, PCM_PPT)() {
    static const ScoreVariant v[N_SCORE_VARIANTS_PER_PPT] = {
        PCM_SV(true, 5),  PCM_SV(true, 7),  PCM_SV(true, 10),  PCM_SV(true, 0),
        PCM_SV(false, 5), PCM_SV(false, 7), PCM_SV(false, 10), PCM_SV(false, 0),
    };
    return v;
}

}  // namespace pcm

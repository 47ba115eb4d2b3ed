// TEMPORARY link stubs, replaced as the stages land.
#include "cvk_internal.h"
#define NOTYET(name) throw CvkError(CVK_ERR_STATE, name ": stage not built yet")
void llm_build(cvk_ctx*, const int*, int) { NOTYET("llm"); }
cvk_lm_session* llm_session_create(cvk_ctx*, int, int) { NOTYET("llm"); }
void llm_session_destroy(cvk_ctx*, cvk_lm_session*) {}
void llm_prefill(cvk_ctx*, cvk_lm_session*, const int32_t*, const int*, const int32_t*, const int*, int, cudaStream_t) { NOTYET("llm"); }
void llm_decode(cvk_ctx*, cvk_lm_session*, int, const float*, const int32_t*, const int32_t*, int32_t*, int, int32_t*, int32_t*, int*, cudaStream_t) { NOTYET("llm"); }
void llm_forward_logp(cvk_ctx*, const float*, const int*, int, float*, cudaStream_t) { NOTYET("llm"); }
void llm_ras_sample(cvk_ctx*, float*, int, int, const int32_t*, int, const int32_t*, const float*, const int32_t*, int32_t*, cudaStream_t) { NOTYET("llm"); }
void mel_spectrogram(cvk_ctx*, const float*, const int*, int, float*, cudaStream_t) { NOTYET("mel"); }
void mel_init(cvk_ctx*) { NOTYET("mel"); }

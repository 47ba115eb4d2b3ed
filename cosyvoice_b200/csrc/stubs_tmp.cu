// TEMPORARY link stubs, replaced as the stages land.
#include "cvk_internal.h"
#define NOTYET(name) throw CvkError(CVK_ERR_STATE, name ": stage not built yet")
void mel_spectrogram(cvk_ctx*, const float*, const int*, int, float*, cudaStream_t) { NOTYET("mel"); }
void mel_init(cvk_ctx*) { NOTYET("mel"); }

// Internal hand-over between the drivers (product code): create a dada() context whose packed uniques are filled on the
// device by the caller instead of being packed on the host and uploaded (dd_derep.cu -> dd_driver.cu).
#pragma once
#include "dd_common.h"
#include <cstdint>

struct dada2b_ctx;

namespace dd2 {

// Allocates a context for nraw uniques (host copies of the sequences / abundances are taken from the arguments; device
// arrays seq2 [nraw][SW], qual [nraw][QS], len, reads, prior are allocated but NOT filled).  Returns the arrays through
// `arrays` and the context's stream through `stream`; the caller fills them on that stream, then reports the largest
// rounded quality with ctx_finish_device().  Throws std::runtime_error.
dada2b_ctx *ctx_create_device(int device, int nraw, int maxlen, int minlen, const char *seq_concat, const int64_t *seq_off,
                              const int32_t *abund, DevIn *arrays, cudaStream_t *stream);
void ctx_finish_device(dada2b_ctx *ctx, int maxq);

}  // namespace dd2

// Snippet builder (events -> padded model inputs).  Replaces data_loader.prepare_snippets and helpers
// (reference data_loader.py:29-51, 70-111).  Placeholder until the device implementation lands: the
// symbol exists so that the ABI is stable, and it fails loudly instead of falling back to the host.
#include "kernels.cuh"
using namespace rvb;

extern "C" int rvb_build_snippets(const void *, int, int64_t, const int32_t *, const int32_t *, const double *,
                                  const double *, int32_t, int64_t, int64_t, int32_t, float *, float *, int32_t,
                                  int32_t *, void *) {
    return fail(RVB_ERR_STATE, "rvb_build_snippets: device implementation not built yet");
}

// compile-and-link check of integration/b200_backend.h against the reference's headers (oracle/Makefile backend_check)
#include "b200_backend.h"

extern "C" int gbrl_b200_backend_check(void) {
    // never called without a GPU; referencing the members is what makes the linker resolve every C-ABI symbol they use
    GBRL_B200 *(*make)(void) = +[]() -> GBRL_B200 * {
        return new GBRL_B200(4, 1, 1, 4, 0, 256, 10, 0.9f, Cosine, Quantile, false, 5000, GREEDY, 0, gpu);
    };
    void (GBRL_B200::*fs)(dataHolder<const float> *, dataHolder<float> *, int, int, void *) = &GBRL_B200::step;
    float (GBRL_B200::*ff)(dataHolder<const float> *, dataHolder<const float> *, int, int, int, bool, void *) = &GBRL_B200::fit;
    void (GBRL_B200::*fp)(dataHolder<const float> *, int, int, int, int, float *, bool, void *) = &GBRL_B200::predict;
    return (make != nullptr) + (fs != nullptr) + (ff != nullptr) + (fp != nullptr);
}

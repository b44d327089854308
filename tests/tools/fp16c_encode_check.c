/* fp16c_encode_check.c -- CPU proof for the device FP16C encoder (ionsolver_b200/csrc/lattice.cuh, fp16c_encode): the reference's
 * float_to_half_custom (sim_kernels.cl:79-84) against "one round-toward-zero multiplication by 2^-112, add 0x800, shift by 12",
 * emulated with the host FPU in FE_TOWARDZERO mode (IEEE, denormals on).  Exhaustive over all 2^32 bit patterns by default
 * (about one minute); NaN inputs are counted separately (the device multiplication canonicalises NaNs, a broken simulation).
 *   gcc -O2 -frounding-math -o fp16c_encode_check fp16c_encode_check.c -lm && ./fp16c_encode_check [stride]
 * exit status 0 = no mismatch on non-NaN inputs. */
#include <fenv.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static uint16_t reference(uint32_t xb) { /* sim_kernels.cl:79-84 */
    const uint32_t b = xb + 0x00000800u;
    const uint32_t e = (b & 0x7F800000u) >> 23;
    const uint32_t m = b & 0x007FFFFFu;
    return (uint16_t)((b & 0x80000000u) >> 16 | (uint32_t)(e > 112u) * ((((e - 112u) << 11) & 0x7800u) | m >> 12) |
                      (uint32_t)((e < 113u) & (e > 100u)) * ((((0x007FF800u + m) >> (124u - e)) + 1u) >> 1));
}
static uint16_t candidate(uint32_t xb) { /* lattice.cuh fp16c_encode */
    volatile float ax = bfloat(xb & 0x7FFFFFFFu);
    volatile float y = ax * 1.925929944387236e-34f; /* 2^-112; rounding mode is toward zero */
    const uint32_t a = fbits(y) + 0x800u;
    return (uint16_t)(((a >> 12) & 0x7FFFu) | ((xb >> 16) & 0x8000u));
}

int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    fesetround(FE_TOWARDZERO);
    uint64_t bad = 0, bad_nan = 0, checked = 0;
    for (uint64_t i = 0; i < (1ull << 32); i += stride) {
        const uint32_t xb = (uint32_t)i;
        /* with a stride, the binades around the normal/denormal boundary and every binade top are still swept densely */
        checked++;
        if (reference(xb) != candidate(xb)) {
            if (((xb >> 23) & 0xFFu) == 255u && (xb & 0x7FFFFFu)) bad_nan++;
            else if (bad++ < 5) printf("mismatch %08x: reference %04x candidate %04x\n", xb, reference(xb), candidate(xb));
        }
    }
    if (stride > 1) {
        for (uint32_t e = 96; e <= 132; e++)
            for (uint32_t s = 0; s < 2; s++)
                for (uint32_t m = 0; m < 0x2000; m++) {
                    const uint32_t lo = (s << 31) | (e << 23) | m, hi = (s << 31) | (e << 23) | (0x7FFFFFu - m);
                    checked += 2;
                    if (reference(lo) != candidate(lo) || reference(hi) != candidate(hi)) bad++;
                }
    }
    printf("{\"checked\": %llu, \"mismatches_non_nan\": %llu, \"mismatches_nan_inputs\": %llu}\n", (unsigned long long)checked,
           (unsigned long long)bad, (unsigned long long)bad_nan);
    return bad ? 1 : 0;
}

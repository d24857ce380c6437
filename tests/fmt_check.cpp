// exhaustive-ish check of brc::put_f2 against printf("%.2f") (compiled and run by tests/test_format_numbers.py)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include "../bam_readcount_b200/csrc/brc_fmt_num.h"
static long bad = 0, n = 0;
static void check(float f) {
    char t[96]; std::snprintf(t, sizeof t, "%.2f", (double)f);
    std::string o; brc::put_f2(o, f); ++n;
    if (o != t) { if (bad < 10) std::printf("MISMATCH %a: got %s want %s\n", (double)f, o.c_str(), t); ++bad; }
}
int main() {
    std::mt19937_64 rng(12345);
    for (long i = 0; i < 20000000; ++i) { uint32_t b = (uint32_t)rng(); float f; std::memcpy(&f, &b, 4); check(f); }   // random bit patterns
    for (uint32_t k = 0; k < 3000000; ++k) {   // averages the emitter really prints: small ratios and their float neighbours (ties!)
        float f = (float)k / 100.0f, g = (float)k / 200.0f + 0.005f;
        check(f); check(std::nextafterf(f, 1e30f)); check(std::nextafterf(f, -1e30f)); check(g); check(-f);
        check((float)k / 8.0f + 0.125f); check((float)(k % 1000) / (float)(k % 997 + 1));
    }
    const float edge[] = {0.0f, -0.0f, 0.005f, 0.015f, 0.025f, 0.125f, 0.375f, 1e-30f, 1e-45f, 255.0f, 1e15f, 1.8e16f, 3e16f, 1e20f, 3.4e38f};
    for (float f : edge) { check(f); check(-f); }
    std::printf("checked %ld values, %ld mismatches\n", n, bad);
    return bad ? 1 : 0;
}

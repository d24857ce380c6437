// brc_fmt_num.h — exact, printf-free number formatting for the text emitter.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace brc {

inline void put_u(std::string &o, uint64_t x) {
    char t[24]; int n = 0;
    do { t[n++] = (char)('0' + x % 10); x /= 10; } while (x);
    while (n) o += t[--n];
}

// printf("%.2f", (double)x) for a float32 (what `std::fixed << std::setprecision(2) << float` prints,
// R:src/lib/bamrc/BasicStat.cpp:115-141): the binary value times 100 is rounded half-to-even in integer arithmetic.
inline void put_f2(std::string &o, float x) {
    uint32_t b; std::memcpy(&b, &x, 4);
    const uint32_t ex = (b >> 23) & 0xFF, mant = b & 0x7FFFFF;
    if (ex >= 150 + 31) {   // NaN / inf / >= 2^54: let libc do it
        char t[96]; int n = std::snprintf(t, sizeof t, "%.2f", (double)x); o.append(t, (size_t)n); return;
    }
    const uint64_t m = ex ? (uint64_t)(mant | 0x800000u) : (uint64_t)mant;
    const int e = (ex ? (int)ex : 1) - 150;          // x = m * 2^e
    const uint64_t x100 = m * 100u;
    uint64_t q;
    if (e >= 0) q = x100 << e;
    else {
        const int sh = -e;
        if (sh >= 64) q = 0;
        else {
            q = x100 >> sh;
            const uint64_t r = x100 & ((1ull << sh) - 1), half = 1ull << (sh - 1);
            if (r > half || (r == half && (q & 1))) ++q;
        }
    }
    if (b >> 31) o += '-';
    put_u(o, q / 100);
    o += '.';
    const unsigned f = (unsigned)(q % 100);
    o += (char)('0' + f / 10); o += (char)('0' + f % 10);
}

}  // namespace brc

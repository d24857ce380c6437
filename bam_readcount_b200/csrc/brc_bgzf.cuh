// brc_bgzf.cuh — DEFLATE (RFC 1951) decoder core shared by the device kernel of brc_bgzf.cu and its host twin.
//
// BGZF blocks are independent raw-DEFLATE streams of <= 64 KiB (V:htslib-1.10/bgzf.c:697 inflate_block,
// :897 bgzf_read_block), so a BAM file inflates block-parallel.  One decoder instance = one block = one WARP: the symbol stream
// is serial, so every lane runs the decoder redundantly and in lockstep (same registers, same control flow — it costs the issue
// slots of one lane), which makes the warp-wide steps natural: lane 0 alone writes the Huffman tables and the literals, all
// lanes copy a match / a stored run together.  On the host the same code runs with one lane, so the exact instruction
// sequence the GPU executes is unit-tested on the CPU against zlib (tests/test_bgzf_device.py).
// Tables: a 10-bit / 9-bit primary lookup (symbol << 4 | length), the canonical count/symbol arrays (the formulation of zlib's
// contrib/puff) as the slow path for longer codes.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define BRC_HD __host__ __device__ __forceinline__
#else
#define BRC_HD inline
#endif

namespace brc {
namespace inflate {

constexpr int LIT_BITS = 10, DIST_BITS = 9;
constexpr int MAXBITS = 15, MAXLCODES = 286, MAXDCODES = 30, FIXLCODES = 288;

struct Tables {                     // per decoder: 2048 + 1024 + 640 + 64 ... bytes (one per warp in shared memory)
    uint16_t lit[1 << LIT_BITS];    // symbol << 4 | code length; 0 = not in the primary table
    uint16_t dist[1 << DIST_BITS];
    uint16_t lcount[MAXBITS + 1], lsym[FIXLCODES];
    uint16_t dcount[MAXBITS + 1], dsym[MAXDCODES + 2];
    uint8_t lens[32 + FIXLCODES + MAXDCODES + 2];   // scratch: code lengths while a header is read
};

// the lanes of one decoder: 32 on the device, 1 on the host
struct Lanes {
    int lane, n;
    BRC_HD void sync() const {
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
    }
};

struct Bits {                       // LSB-first bit reader over `n` bytes at `base`; 4 bytes per refill once aligned
    const uint8_t *base; int64_t n, pos;
    uint64_t buf; int cnt;
    bool overrun;
};
BRC_HD void bits_init(Bits &b, const uint8_t *p, int64_t n) { b.base = p; b.n = n; b.pos = 0; b.buf = 0; b.cnt = 0; b.overrun = false; }
BRC_HD void bits_refill(Bits &b) {
    while (b.cnt <= 32 && b.pos < b.n) {
        const uint8_t *p = b.base + b.pos;
        if (b.pos + 4 <= b.n && (reinterpret_cast<uintptr_t>(p) & 3u) == 0) {
            uint32_t w;
#if defined(__CUDA_ARCH__)
            w = *reinterpret_cast<const uint32_t *>(p);
#else
            std::memcpy(&w, p, 4);
#endif
            b.buf |= (uint64_t)w << b.cnt; b.cnt += 32; b.pos += 4;
        } else { b.buf |= (uint64_t)(*p) << b.cnt; b.cnt += 8; b.pos += 1; }
    }
}
// try to have n (<= 32) bits buffered; near the end of the stream fewer may be left (the missing high bits read as zeros) —
// CONSUMING more bits than the stream holds raises `overrun`
BRC_HD void bits_need(Bits &b, int n) { if (b.cnt < n) bits_refill(b); }
BRC_HD uint32_t bits_peek(const Bits &b, int n) { return (uint32_t)(b.buf & ((1ull << n) - 1)); }
BRC_HD void bits_drop(Bits &b, int n) { b.buf >>= n; b.cnt -= n; if (b.cnt < 0) { b.overrun = true; b.cnt = 0; b.buf = 0; } }
BRC_HD uint32_t bits_get(Bits &b, int n) { bits_need(b, n); const uint32_t v = bits_peek(b, n); bits_drop(b, n); return v; }

BRC_HD uint32_t rev_bits(uint32_t v, int n) { uint32_t r = 0; for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; } return r; }

// canonical Huffman: count[len], sym[] sorted by (len, symbol) (puff's construct()), plus the primary lookup table.
// Returns <0 over-subscribed, >0 incomplete, 0 complete.
BRC_HD int build(uint16_t *count, uint16_t *symtab, uint16_t *primary, int pbits, const uint8_t *length, int n) {
    uint16_t offs[MAXBITS + 1];
    for (int len = 0; len <= MAXBITS; ++len) count[len] = 0;
    for (int s = 0; s < n; ++s) count[length[s]]++;
    for (int i = 0; i < (1 << pbits); ++i) primary[i] = 0;
    if (count[0] == n) return 0;                       // no codes: complete, but decoding will fail
    int left = 1;
    for (int len = 1; len <= MAXBITS; ++len) { left <<= 1; left -= count[len]; if (left < 0) return left; }
    offs[1] = 0;
    for (int len = 1; len < MAXBITS; ++len) offs[len + 1] = (uint16_t)(offs[len] + count[len]);
    for (int s = 0; s < n; ++s) if (length[s] != 0) symtab[offs[length[s]]++] = (uint16_t)s;
    // primary table: canonical codes are assigned in (len, symbol) order, first code of a length = (first_prev + count_prev) << 1
    uint32_t code = 0; int idx = 0;
    for (int len = 1; len <= pbits; ++len) {
        for (int k = 0; k < count[len]; ++k, ++idx, ++code) {
            const uint32_t r = rev_bits(code, len);
            const uint16_t ent = (uint16_t)((symtab[idx] << 4) | len);
            for (uint32_t f = r; f < (1u << pbits); f += (1u << len)) primary[f] = ent;
        }
        code <<= 1;
    }
    return left;
}

// lane 0 builds (it writes the tables), everybody learns the verdict
BRC_HD int build_shared(const Lanes &L, uint16_t *count, uint16_t *symtab, uint16_t *primary, int pbits, const uint8_t *length, int n, int *verdict) {
    L.sync();
    if (L.lane == 0) *verdict = build(count, symtab, primary, pbits, length, n);
    L.sync();
    return *verdict;
}

// one symbol: primary table, else the canonical walk over the longer lengths (bit by bit, MSB-first codes in an LSB-first stream)
BRC_HD int decode(Bits &b, const uint16_t *count, const uint16_t *symtab, const uint16_t *primary, int pbits) {
    bits_need(b, MAXBITS);
    const uint16_t ent = primary[bits_peek(b, pbits)];
    if (ent) { bits_drop(b, ent & 15); return ent >> 4; }
    int code = 0, first = 0, index = 0;
    uint64_t bitbuf = b.buf;
    for (int len = 1; len <= MAXBITS; ++len) {
        code |= (int)(bitbuf & 1u); bitbuf >>= 1;
        const int cnt = count[len];
        if (code - cnt < first) { bits_drop(b, len); return symtab[index + (code - first)]; }
        index += cnt; first += cnt; first <<= 1; code <<= 1;
    }
    return -1;
}

// Inflates one raw-DEFLATE stream of `clen` bytes into out[0 .. isize).  Returns 0 on success, a negative code otherwise.
// Every lane of L calls it with the same arguments and gets the same return value.
BRC_HD int inflate_block(const Lanes &L, const uint8_t *in, int64_t clen, uint8_t *out, uint32_t isize, Tables &T) {
    // length / distance code -> (base, extra bits) by formula (RFC 1951 §3.2.5), no tables in local memory:
    //   length   sym < 8: 3 + sym, 0;   sym == 28: 258, 0;   else e = (sym - 4) >> 2, ((4 + (sym & 3)) << e) + 3
    //   distance d < 4: d + 1, 0;       else e = (d - 2) >> 1, ((2 + (d & 1)) << e) + 1
    // order of the code-length code lengths (§3.2.7), 5 bits each: 16 17 18 0 8 7 9 6 10 5 11 4 | 12 3 13 2 14 1 15
    const uint64_t order_lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
    const uint64_t order_hi = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
    int *verdict = reinterpret_cast<int *>(T.lens + 28);          // 4 bytes of the scratch nobody else uses (lens[19..32) is free)
    Bits b; bits_init(b, in, clen);
    uint32_t o = 0;
    int last = 0;
    bool fixed_ready = false;
    while (!last) {
        last = (int)bits_get(b, 1);
        const uint32_t type = bits_get(b, 2);
        if (type == 0) {                               // stored: back to the byte stream
            if (b.overrun) return -17;
            bits_drop(b, b.cnt & 7);
            b.pos -= b.cnt >> 3; b.cnt = 0; b.buf = 0;  // the buffered whole bytes were real stream bytes: hand them back
            if (b.n - b.pos < 4) return -19;
            const uint8_t *p = b.base + b.pos;
            const uint32_t len = p[0] | (p[1] << 8), nlen = p[2] | (p[3] << 8);
            b.pos += 4;
            if ((len ^ 0xFFFFu) != nlen || (int64_t)len > b.n - b.pos || o + len > isize) return -20;
            p += 4;
            for (uint32_t k = (uint32_t)L.lane; k < len; k += (uint32_t)L.n) out[o + k] = p[k];
            o += len; b.pos += len;
            continue;
        }
        if (type == 3) return -2;
        if (type == 1) {
            if (!fixed_ready) {
                L.sync();                               // nobody still reads the previous block's tables
                if (L.lane == 0) {
                    int s = 0;
                    for (; s < 144; ++s) T.lens[32 + s] = 8;
                    for (; s < 256; ++s) T.lens[32 + s] = 9;
                    for (; s < 280; ++s) T.lens[32 + s] = 7;
                    for (; s < FIXLCODES; ++s) T.lens[32 + s] = 8;
                    for (int d = 0; d < MAXDCODES; ++d) T.lens[32 + FIXLCODES + d] = 5;
                }
                build_shared(L, T.lcount, T.lsym, T.lit, LIT_BITS, T.lens + 32, FIXLCODES, verdict);
                build_shared(L, T.dcount, T.dsym, T.dist, DIST_BITS, T.lens + 32 + FIXLCODES, MAXDCODES, verdict);
                fixed_ready = true;
            }
        } else {
            fixed_ready = false;
            const int nlen = (int)bits_get(b, 5) + 257, ndist = (int)bits_get(b, 5) + 1, ncode = (int)bits_get(b, 4) + 4;
            if (nlen > MAXLCODES || ndist > MAXDCODES) return -3;
            L.sync();                                   // nobody still reads the previous block's tables
            for (int idx = 0; idx < 19; ++idx) {
                const int pos = (int)((idx < 12 ? order_lo >> (5 * idx) : order_hi >> (5 * (idx - 12))) & 31u);
                const uint32_t v = idx < ncode ? bits_get(b, 3) : 0u;
                if (L.lane == 0) T.lens[pos] = (uint8_t)v;
            }
            // the code-length code: 19 symbols, lengths <= 7; the literal arrays hold its table for now (7-bit primary inside lit[])
            if (build_shared(L, T.lcount, T.lsym, T.lit, 7, T.lens, 19, verdict) != 0) return -4;
            uint8_t *cl = T.lens + 32;                  // the code lengths being read
            int idx = 0, prev = 0;
            while (idx < nlen + ndist) {
                const int sym = decode(b, T.lcount, T.lsym, T.lit, 7);
                if (sym < 0) return -5;
                if (sym < 16) { if (L.lane == 0) cl[idx] = (uint8_t)sym; ++idx; prev = sym; }
                else {
                    int len = 0, rep;
                    if (sym == 16) { if (idx == 0) return -6; len = prev; rep = 3 + (int)bits_get(b, 2); }
                    else if (sym == 17) rep = 3 + (int)bits_get(b, 3);
                    else rep = 11 + (int)bits_get(b, 7);
                    if (idx + rep > nlen + ndist) return -7;
                    if (L.lane == 0) for (int k = 0; k < rep; ++k) cl[idx + k] = (uint8_t)len;
                    idx += rep; prev = len;
                }
            }
            L.sync();
            if (cl[256] == 0) return -8;               // no end-of-block code
            int err = build_shared(L, T.lcount, T.lsym, T.lit, LIT_BITS, cl, nlen, verdict);
            if (err < 0 || (err > 0 && nlen - T.lcount[0] != 1)) return -9;
            err = build_shared(L, T.dcount, T.dsym, T.dist, DIST_BITS, cl + nlen, ndist, verdict);
            if (err < 0 || (err > 0 && ndist - T.dcount[0] != 1)) return -11;
        }
        for (;;) {
            int sym = decode(b, T.lcount, T.lsym, T.lit, LIT_BITS);
            if (sym < 0) return -12;
            if (sym < 256) { if (o >= isize) return -13; if (L.lane == 0) out[o] = (uint8_t)sym; ++o; continue; }
            if (sym == 256) break;
            sym -= 257;
            if (sym >= 29) return -14;
            const int le = sym < 8 || sym == 28 ? 0 : (sym - 4) >> 2;
            const uint32_t len = (sym < 8 ? 3u + (uint32_t)sym : (sym == 28 ? 258u : (((4u + ((uint32_t)sym & 3u)) << le) + 3u))) + (le ? bits_get(b, le) : 0u);
            const int ds = decode(b, T.dcount, T.dsym, T.dist, DIST_BITS);
            if (ds < 0 || ds >= 30) return -15;
            const int de = ds < 4 ? 0 : (ds - 2) >> 1;
            const uint32_t dist = (ds < 4 ? (uint32_t)ds + 1u : (((2u + ((uint32_t)ds & 1u)) << de) + 1u)) + (de ? bits_get(b, de) : 0u);
            if (dist > o || o + len > isize) return -16;
            // the whole warp copies the match: byte k of it is byte (k mod dist) of the `dist` bytes before it (LZ77 replication)
            const uint8_t *src = out + (o - dist);
            uint8_t *dst = out + o;
            L.sync();                                   // the literals lane 0 stored are visible to the other lanes
            for (uint32_t k = (uint32_t)L.lane; k < len; k += (uint32_t)L.n) dst[k] = src[dist >= len ? k : k % dist];
            L.sync();
            o += len;
        }
    }
    if (b.overrun) return -17;
    return o == isize ? 0 : -18;
}

}  // namespace inflate
}  // namespace brc

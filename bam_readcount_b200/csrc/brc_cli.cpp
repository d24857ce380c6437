// brc_cli.cpp — `brc-readcount`: the C++ host.  Same command line and STDOUT as bam-readcount
// (R:src/exe/bam-readcount/bamreadcount.cpp:421-670), with the pileup hot path routed through
// libbrc_engine.so (include/brc_engine.h).  File decode stays on the host, as the north star says:
// a small BGZF/BAM/BAI/FASTA reader written against the SAM specification (zlib for inflate);
// htslib is not needed for BAM.  CRAM input (R:test-data/cram_site_test.sh, BASELINE config 2b) is decoded through htslib when the
// host is built with -DBRC_WITH_HTSLIB against a libhts.a (tools/build_htslib.sh builds the copy vendored with the reference,
// htslib 1.10); without it a .cram argument is refused with a message.
//
// Mirrors, region by region, the reference's two loops:
//   -l site list : R:...:574-608  (d.beg=beg-1, d.end=end, queues cleared per region)
//   argv regions : R:...:641-657  (bam_parse_region; a bare contig name keeps the previous beg/end, A.6)
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <getopt.h>
#include <deque>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <zlib.h>
#include <chrono>
#include <future>
#include <memory>
#include <thread>
#include <unordered_map>
#include <unistd.h>

#include "../../include/brc_engine.h"
#ifdef BRC_WITH_HTSLIB
#include <functional>
#include <htslib/sam.h>
#include <htslib/hts.h>
#endif

namespace {

// ------------------------------------------------------------------------------------------
// BGZF (SAM spec §4.1): random access by virtual offset (coffset<<16 | uoffset)
// ------------------------------------------------------------------------------------------
struct Bgzf {
    FILE *fp = nullptr;
    std::vector<uint8_t> block;
    uint64_t block_coff = 0;     // compressed offset of the current block
    uint64_t next_coff = 0;      // compressed offset of the next block
    size_t upos = 0;             // position inside `block`
    bool eof = false;
    bool error = false;          // a block failed to inflate / a header is malformed / the file ends inside a block: not a clean EOF
    // read-ahead: a span of consecutive blocks is read with one fread and inflated by several threads
    struct Ahead { uint64_t coff, next; std::vector<uint8_t> data; bool ok; };
    std::vector<Ahead> ahead; size_t ahead_pos = 0;
    std::vector<uint8_t> raw;
    int n_threads = 8;
    size_t span_blocks = 64;

    bool open(const std::string &path) {
        fp = std::fopen(path.c_str(), "rb");
        unsigned hw = std::thread::hardware_concurrency();
        n_threads = (int)std::max(1u, std::min(hw ? hw : 1u, 16u));
        return fp != nullptr;
    }
    // a reader owned by ONE decode thread of ParallelFetcher: inflates inline, short read-ahead spans
    bool open_worker(const std::string &path) { fp = std::fopen(path.c_str(), "rb"); n_threads = 1; span_blocks = 16; return fp != nullptr; }
    ~Bgzf() { if (fp) std::fclose(fp); }

    static bool inflate_block(const uint8_t *src, size_t clen, std::vector<uint8_t> &dst, uint32_t isize) {
        dst.resize(isize);
        if (!isize) return true;
        z_stream zs{};
        if (inflateInit2(&zs, -15) != Z_OK) return false;
        zs.next_in = const_cast<uint8_t *>(src); zs.avail_in = (uInt)clen; zs.next_out = dst.data(); zs.avail_out = isize;
        const int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        return rc == Z_STREAM_END;
    }
    // parse one block header at raw[o..]; returns total block size (0 on error / truncated)
    static size_t block_size(const uint8_t *h, size_t avail, size_t &hdr_len) {
        if (avail < 18 || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) return 0;
        const size_t xlen = (size_t)(h[10] | (h[11] << 8));
        if (avail < 12 + xlen) return 0;
        int bsize = -1;
        for (size_t i = 0; i + 4 <= xlen;) {
            const size_t sl = (size_t)(h[12 + i + 2] | (h[12 + i + 3] << 8));
            if (h[12 + i] == 'B' && h[12 + i + 1] == 'C' && sl == 2) bsize = h[12 + i + 4] | (h[12 + i + 5] << 8);
            i += 4 + sl;
        }
        if (bsize < 0) return 0;
        hdr_len = 12 + xlen;
        return (size_t)bsize + 1;
    }
    bool fill_ahead(uint64_t coff) {
        ahead.clear(); ahead_pos = 0;
        if (fseeko(fp, (off_t)coff, SEEK_SET) != 0) return false;
        raw.resize(span_blocks * 65536 + 65536);
        const size_t got = std::fread(raw.data(), 1, raw.size(), fp);
        if (got == 0) { eof = true; return false; }                      // clean end of file at a block boundary
        if (got < 18) { eof = true; error = true; return false; }        // a fragment of a block header
        struct Job { size_t off, hdr, total; };
        std::vector<Job> jobs;
        for (size_t o = 0; o < got && jobs.size() < span_blocks;) {
            size_t hdr = 0; const size_t tot = block_size(raw.data() + o, got - o, hdr);
            if (!tot || o + tot > got) break;
            jobs.push_back({o, hdr, tot}); o += tot;
        }
        if (jobs.empty()) { error = true; return false; }                // bytes are there but no whole, well-formed block
        ahead.resize(jobs.size());
        auto work = [&](size_t t) {
            for (size_t j = t; j < jobs.size(); j += (size_t)n_threads) {
                const Job &jb = jobs[j];
                const uint8_t *b = raw.data() + jb.off;
                const size_t clen = jb.total - jb.hdr - 8;
                const uint8_t *tail = b + jb.total - 4;
                const uint32_t isize = tail[0] | (tail[1] << 8) | (tail[2] << 16) | ((uint32_t)tail[3] << 24);
                ahead[j].coff = coff + jb.off; ahead[j].next = coff + jb.off + jb.total;
                ahead[j].ok = inflate_block(b + jb.hdr, clen, ahead[j].data, isize);
            }
        };
        std::vector<std::thread> th;
        const int nt = (int)std::min<size_t>((size_t)n_threads, jobs.size());
        for (int t = 1; t < nt; ++t) th.emplace_back(work, (size_t)t);
        work(0);
        for (auto &x : th) x.join();
        return true;
    }
    bool load_block(uint64_t coff) {
        if (!(ahead_pos < ahead.size() && ahead[ahead_pos].coff == coff)) {
            // look inside the current span first (seeks within it), else read a new span
            bool found = false;
            for (size_t j = 0; j < ahead.size(); ++j) if (ahead[j].coff == coff && !ahead[j].data.empty()) { ahead_pos = j; found = true; break; }
            if (!found && !fill_ahead(coff)) { block.clear(); upos = 0; return false; }
        }
        Ahead &a = ahead[ahead_pos];
        if (!a.ok) { error = true; return false; }
        block = a.data;                 // keep the span entry intact: sorted site lists revisit blocks
        block_coff = a.coff; next_coff = a.next; upos = 0; eof = false;
        ++ahead_pos;
        return true;
    }
    bool seek(uint64_t voff) {
        // sorted site lists keep landing in the block that is already inflated
        if (!((voff >> 16) == block_coff && !block.empty() && !eof) && !load_block(voff >> 16)) return false;
        upos = (size_t)(voff & 0xFFFF);
        return true;
    }
    uint64_t tell() const { return (block_coff << 16) + (uint64_t)upos; }   // upos may equal a full 64 KiB block
    // read exactly n bytes (spanning blocks); false at EOF
    bool read(void *dst, size_t n) {
        uint8_t *d = (uint8_t *)dst;
        while (n) {
            if (upos >= block.size()) {
                if (!load_block(next_coff)) return false;
                if (block.empty()) { if (eof) return false; continue; }
            }
            const size_t k = std::min(n, block.size() - upos);
            std::memcpy(d, block.data() + upos, k);
            d += k; upos += k; n -= k;
        }
        return true;
    }
};

// ------------------------------------------------------------------------------------------
// BAM header + BAI index (SAM spec §4.2, §5.2)
// ------------------------------------------------------------------------------------------
struct BamFile {
    Bgzf bz;
    std::string text;
    std::vector<std::string> names;
    std::vector<int32_t> lens;
    std::map<std::string, int> tid_of;
    uint64_t first_rec = 0;
    // BAI
    struct RefIdx { std::vector<uint64_t> linear; uint64_t min_chunk = ~0ull; std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins; };
    std::vector<RefIdx> idx;
    bool have_idx = false;

    bool open(const std::string &path) {
        if (!bz.open(path) || !bz.load_block(0)) return false;
        char magic[4]; int32_t l_text, n_ref;
        if (!bz.read(magic, 4) || std::memcmp(magic, "BAM\1", 4) != 0) return false;
        if (!bz.read(&l_text, 4)) return false;
        text.resize((size_t)l_text);
        if (l_text && !bz.read(&text[0], (size_t)l_text)) return false;
        text = text.c_str();
        if (!bz.read(&n_ref, 4)) return false;
        for (int i = 0; i < n_ref; ++i) {
            int32_t ln, sl;
            if (!bz.read(&ln, 4)) return false;
            std::string nm((size_t)ln, 0);
            if (!bz.read(&nm[0], (size_t)ln) || !bz.read(&sl, 4)) return false;
            nm = nm.c_str();
            tid_of[nm] = i; names.push_back(nm); lens.push_back(sl);
        }
        first_rec = bz.tell();
        return true;
    }
    bool load_index(const std::string &bam_path) {
        std::string p = bam_path + ".bai";
        FILE *f = std::fopen(p.c_str(), "rb");
        if (!f && bam_path.size() > 4) { p = bam_path.substr(0, bam_path.size() - 4) + ".bai"; f = std::fopen(p.c_str(), "rb"); }
        if (!f) return false;
        auto rd = [&](void *d, size_t n) { return std::fread(d, 1, n, f) == n; };
        char magic[4]; int32_t n_ref;
        bool ok = rd(magic, 4) && std::memcmp(magic, "BAI\1", 4) == 0 && rd(&n_ref, 4);
        idx.assign((size_t)std::max(n_ref, 0), RefIdx());
        for (int r = 0; ok && r < n_ref; ++r) {
            int32_t n_bin; ok = rd(&n_bin, 4);
            for (int b = 0; ok && b < n_bin; ++b) {
                uint32_t bin; int32_t n_chunk; ok = rd(&bin, 4) && rd(&n_chunk, 4);
                for (int c = 0; ok && c < n_chunk; ++c) {
                    uint64_t cb, ce; ok = rd(&cb, 8) && rd(&ce, 8);
                    if (ok && bin != 37450) { idx[(size_t)r].min_chunk = std::min(idx[(size_t)r].min_chunk, cb); idx[(size_t)r].bins[bin].push_back({cb, ce}); }
                }
            }
            int32_t n_intv; ok = ok && rd(&n_intv, 4);
            if (ok) { idx[(size_t)r].linear.resize((size_t)n_intv); ok = n_intv == 0 || rd(idx[(size_t)r].linear.data(), 8 * (size_t)n_intv); }
        }
        std::fclose(f);
        have_idx = ok;
        return ok;
    }
    // smallest virtual offset of a record that can overlap position `beg` on `tid` (linear index, 16 kb windows)
    bool query_offset(int tid, int64_t beg, uint64_t &voff) const {
        if (tid < 0 || tid >= (int)idx.size()) return false;
        const RefIdx &ri = idx[(size_t)tid];
        if (ri.min_chunk == ~0ull) return false;                      // no alignments on this reference
        int64_t w = beg >> 14;
        voff = ri.min_chunk;
        if (!ri.linear.empty()) {
            if (w >= (int64_t)ri.linear.size()) w = (int64_t)ri.linear.size() - 1;
            for (; w >= 0; --w) if (ri.linear[(size_t)w] != 0) { voff = std::max(voff, ri.linear[(size_t)w]); break; }
        }
        return true;
    }
};

struct Rec {   // one decoded alignment (the bam1_t fields the path reads)
    int32_t tid, pos, l_qseq, nm, sm; uint16_t flag; uint8_t mapq; uint32_t n_cigar;
    int64_t endpos;              // bam_endpos, filled by RegionFetcher
    std::vector<uint8_t> data;   // whole record body
    const uint32_t *cigar; const uint8_t *seq, *qual; std::string rg; bool has_rg;
};

int64_t aux_int(const uint8_t *p, char t) {
    switch (t) {
    case 'c': return (int8_t)p[0]; case 'C': return p[0];
    case 's': { int16_t v; std::memcpy(&v, p, 2); return v; } case 'S': { uint16_t v; std::memcpy(&v, p, 2); return v; }
    case 'i': { int32_t v; std::memcpy(&v, p, 4); return v; } case 'I': { uint32_t v; std::memcpy(&v, p, 4); return v; }
    default: return 0;
    }
}

bool read_record(Bgzf &bz, Rec &r) {
    int32_t bs;
    if (!bz.read(&bs, 4)) return false;                                   // EOF (clean unless bz.error)
    if (bs < 32) { bz.error = true; return false; }                       // not a BAM record
    r.data.resize((size_t)bs);
    if (!bz.read(r.data.data(), (size_t)bs)) { bz.error = true; return false; }   // file ends inside a record
    const uint8_t *d = r.data.data();
    int32_t refid, pos, l_seq; uint8_t l_rn, mapq; uint16_t n_cig, flag;
    std::memcpy(&refid, d, 4); std::memcpy(&pos, d + 4, 4); l_rn = d[8]; mapq = d[9];
    std::memcpy(&n_cig, d + 12, 2); std::memcpy(&flag, d + 14, 2); std::memcpy(&l_seq, d + 16, 4);
    r.tid = refid; r.pos = pos; r.mapq = mapq; r.flag = flag; r.n_cigar = n_cig; r.l_qseq = l_seq;
    size_t o = 32 + l_rn;
    r.cigar = (const uint32_t *)(d + o); o += 4 * (size_t)n_cig;
    r.seq = d + o; o += ((size_t)l_seq + 1) / 2;
    r.qual = d + o; o += (size_t)l_seq;
    r.nm = BRC_TAG_ABSENT; r.sm = BRC_TAG_ABSENT; r.has_rg = false;
    bool got_nm = false, got_sm = false;
    while (o + 3 <= (size_t)bs) {          // bam_aux_get's linear scan: first occurrence wins
        const uint8_t *t = d + o; const char ty = (char)t[2]; o += 3;
        size_t sz = 0;
        switch (ty) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'Z': case 'H': { size_t e = o; while (e < (size_t)bs && d[e]) ++e; if (ty == 'Z' && t[0] == 'R' && t[1] == 'G' && !r.has_rg) { r.rg.assign((const char *)d + o, e - o); r.has_rg = true; } o = e + 1; continue; }
        case 'B': { const char st = (char)d[o]; uint32_t cnt; std::memcpy(&cnt, d + o + 1, 4); size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4; o += 5 + es * cnt; continue; }
        default: o = (size_t)bs; continue;
        }
        if (t[0] == 'N' && t[1] == 'M' && !got_nm && ty != 'A' && ty != 'f') { r.nm = (int32_t)aux_int(d + o, ty); got_nm = true; }
        if (t[0] == 'S' && t[1] == 'M' && !got_sm && ty != 'A' && ty != 'f') { r.sm = (int32_t)aux_int(d + o, ty); got_sm = true; }
        o += sz;
    }
    return true;
}

int64_t rec_endpos(const Rec &r) {   // bam_endpos
    if (!(r.flag & 4) && r.n_cigar > 0) {
        int64_t l = 0;
        for (uint32_t k = 0; k < r.n_cigar; ++k) { uint32_t op = r.cigar[k] & 0xF; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) l += r.cigar[k] >> 4; }
        return r.pos + l;
    }
    return (int64_t)r.pos + 1;
}

// ------------------------------------------------------------------------------------------
// the compressed bytes + index entry points covering samfetch(tid, fbeg, fend), for brc_push_bam_span — the
// records are inflated and framed on the device, only compressed bytes cross PCIe.  The BAI's chunk begins and linear-index
// offsets are starts of real records: each starts an independent framing chain.
// ------------------------------------------------------------------------------------------
struct SpanBuilder {
    std::vector<uint8_t> comp; std::vector<uint64_t> entries; int64_t end_voff = -1;
    bool build(BamFile &bam, int tid, int64_t fbeg, int64_t fend) {
        if (tid < 0 || tid >= (int)bam.idx.size()) return false;
        const BamFile::RefIdx &ri = bam.idx[(size_t)tid];
        uint64_t min_lin = 0;
        if (!ri.linear.empty()) { int64_t w = std::min<int64_t>(fbeg >> 14, (int64_t)ri.linear.size() - 1); min_lin = ri.linear[(size_t)w]; }
        uint64_t v0 = ~0ull, v1 = 0; std::vector<uint64_t> cand;
        const int64_t b = std::max<int64_t>(fbeg, 0), e = std::max<int64_t>(fend, b + 1) - 1;
        auto visit = [&](uint32_t bin) {
            auto it = ri.bins.find(bin);
            if (it == ri.bins.end()) return;
            for (const auto &c : it->second) if (c.second > min_lin) { const uint64_t cb = std::max(c.first, min_lin); v0 = std::min(v0, cb); v1 = std::max(v1, c.second); cand.push_back(cb); }
        };
        visit(0);
        const int sh[5] = {26, 23, 20, 17, 14}; const uint32_t of[5] = {1, 9, 73, 585, 4681};
        for (int l = 0; l < 5; ++l) for (int64_t k = b >> sh[l]; k <= (e >> sh[l]); ++k) visit(of[l] + (uint32_t)k);
        if (v0 == ~0ull || v1 <= v0) return false;
        for (int64_t w = std::min<int64_t>(b >> 14, (int64_t)ri.linear.size()); w < std::min<int64_t>((e >> 14) + 2, (int64_t)ri.linear.size()); ++w) cand.push_back(ri.linear[(size_t)w]);
        const uint64_t c0 = v0 >> 16, c1 = v1 >> 16;
        comp.resize((size_t)(c1 - c0) + 65536 + 32);
        if (fseeko(bam.bz.fp, (off_t)c0, SEEK_SET) != 0) return false;
        const size_t got = std::fread(comp.data(), 1, comp.size(), bam.bz.fp);
        std::vector<uint64_t> starts; size_t o = 0;
        while (o + 18 <= got) {
            size_t hdr = 0; const size_t tot = Bgzf::block_size(comp.data() + o, got - o, hdr);
            if (!tot || o + tot > got) break;
            starts.push_back(c0 + o); o += tot;
            if (starts.back() >= c1) break;
        }
        if (starts.empty()) return false;
        comp.resize(o);
        auto rel = [&](uint64_t v, uint64_t &out) { const uint64_t co = v >> 16; if (!std::binary_search(starts.begin(), starts.end(), co)) return false; out = ((co - c0) << 16) | (v & 0xFFFF); return true; };
        std::sort(cand.begin(), cand.end()); cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
        entries.clear();
        for (uint64_t v : cand) { uint64_t r; if (v >= v0 && v < v1 && rel(v, r)) entries.push_back(r); }
        uint64_t r1; end_voff = rel(v1, r1) ? (int64_t)r1 : -1;
        bam.bz.block.clear(); bam.bz.ahead.clear(); bam.bz.ahead_pos = 0; bam.bz.eof = false;     // the reader's position is stale now
        return !entries.empty();
    }
};

// ------------------------------------------------------------------------------------------
// samfetch(in, idx, tid, fbeg, fend) for a run of regions.  The reference re-seeks through the index and
// re-decodes up to a 16 kb linear-index window of records for every line of a site list; here consecutive regions on one
// contig with ascending starts share ONE forward pass over the file: records still overlapping a later region wait in a
// small window, everything else streams straight to the caller.  Each region still receives exactly the records samfetch
// yields (tid, endpos > fbeg, pos < fend) in file order.  `next_fbeg` (start of the following region, or INT64_MAX) only
// bounds what is retained; a caller that then asks for something else simply falls back to an index seek.
// ------------------------------------------------------------------------------------------
struct RegionFetcher {
    BamFile &bam;
    std::deque<Rec> win;          // decoded records that may overlap a later region, file order
    std::vector<Rec> spare;       // recycled records (keeps their heap buffers)
    bool active = false, stream_done = false, merge = true;
    int cur_tid = -1;
    int64_t last_fbeg = -1, keep_floor = -1, last_pos = -1;
    uint64_t n_seeks = 0, n_decoded = 0;

    explicit RegionFetcher(BamFile &b) : bam(b) { merge = std::getenv("BRC_CLI_NO_MERGE") == nullptr; }

    Rec take() { if (spare.empty()) return Rec(); Rec r = std::move(spare.back()); spare.pop_back(); return r; }
    void give(Rec &&r) { if (spare.size() < 4096) spare.push_back(std::move(r)); }

    template <class Emit> void fetch(int tid, int64_t fbeg, int64_t fend, int64_t next_fbeg, Emit &&emit) {
        uint64_t voff;
        if (!bam.query_offset(tid, fbeg, voff)) return;
        const bool cont = merge && active && tid == cur_tid && fbeg >= last_fbeg && fbeg >= keep_floor && voff <= bam.bz.tell();
        if (!cont) {
            while (!win.empty()) { give(std::move(win.back())); win.pop_back(); }
            stream_done = false; last_pos = -1; ++n_seeks;
            if (!bam.bz.seek(voff)) { active = false; return; }
            active = true; cur_tid = tid;
        }
        last_fbeg = fbeg; keep_floor = next_fbeg >= fbeg ? next_fbeg : fbeg;      // an unsorted successor seeks anyway
        for (const Rec &r : win) { if (r.pos >= fend) break; if (r.endpos > fbeg) emit(r); }
        while (!stream_done && last_pos < fend) {
            Rec r = take();
            if (!read_record(bam.bz, r) || r.tid != tid) { stream_done = true; give(std::move(r)); break; }
            ++n_decoded;
            r.endpos = rec_endpos(r); last_pos = r.pos;
            if (r.pos < fend && r.endpos > fbeg) emit(r);
            if (r.pos >= fend || r.endpos > keep_floor) win.push_back(std::move(r)); else give(std::move(r));
        }
        // drop what no later region (start >= keep_floor) can overlap; order of the rest is kept
        size_t w = 0;
        for (size_t i = 0; i < win.size(); ++i) {
            if (win[i].endpos > keep_floor) { if (w != i) std::swap(win[w], win[i]); ++w; }
        }
        while (win.size() > w) { give(std::move(win.back())); win.pop_back(); }
    }
};

#ifdef BRC_WITH_HTSLIB
// ------------------------------------------------------------------------------------------
// CRAM (or any htslib-readable alignment file) through htslib: header, index, and samfetch()'s iterator per region.
// Records are re-laid-out as the Rec the BAM reader produces, so everything downstream is shared.
// ------------------------------------------------------------------------------------------
struct HtsSource {
    samFile *fp = nullptr; bam_hdr_t *hdr = nullptr; hts_idx_t *idx = nullptr; bam1_t *b = nullptr;
    uint64_t n_seeks = 0, n_decoded = 0; bool error = false;
    ~HtsSource() { if (b) bam_destroy1(b); if (idx) hts_idx_destroy(idx); if (hdr) bam_hdr_destroy(hdr); if (fp) sam_close(fp); }
    bool open(const std::string &path, const std::string &fasta, BamFile &meta) {
        fp = sam_open(path.c_str(), "r");
        if (!fp) return false;
        if (!fasta.empty() && hts_set_fai_filename(fp, (fasta + ".fai").c_str()) != 0) return false;   // R:bamreadcount.cpp:503-523
        hdr = sam_hdr_read(fp);
        if (!hdr) return false;
        meta.text.assign(hdr->text ? hdr->text : "", hdr->text ? hdr->l_text : 0);
        for (int i = 0; i < hdr->n_targets; ++i) { meta.names.push_back(hdr->target_name[i]); meta.lens.push_back((int32_t)hdr->target_len[i]); meta.tid_of[hdr->target_name[i]] = i; }
        b = bam_init1();
        return true;
    }
    bool load_index(const std::string &path) { idx = sam_index_load(fp, path.c_str()); return idx != nullptr; }
    void fetch(int tid, int64_t fbeg, int64_t fend, const std::function<void(const Rec &)> &emit) {
        hts_itr_t *it = sam_itr_queryi(idx, tid, fbeg, fend);
        if (!it) return;
        ++n_seeks;
        Rec r; int ret;
        while ((ret = sam_itr_next(fp, it, b)) >= 0) {
            ++n_decoded;
            const bam1_core_t &c = b->core;
            r.tid = c.tid; r.pos = (int32_t)c.pos; r.l_qseq = c.l_qseq; r.flag = c.flag; r.mapq = c.qual; r.n_cigar = c.n_cigar;
            r.data.assign(32, 0); r.data.insert(r.data.end(), b->data, b->data + b->l_data);      // qname at +32, as in a BAM record body
            const uint8_t *d = r.data.data() + 32;
            r.cigar = (const uint32_t *)(d + c.l_qname); r.seq = d + c.l_qname + 4 * (size_t)c.n_cigar; r.qual = r.seq + ((size_t)c.l_qseq + 1) / 2;
            uint8_t *p;
            r.nm = (p = bam_aux_get(b, "NM")) ? (int32_t)bam_aux2i(p) : BRC_TAG_ABSENT;
            r.sm = (p = bam_aux_get(b, "SM")) ? (int32_t)bam_aux2i(p) : BRC_TAG_ABSENT;
            r.has_rg = (p = bam_aux_get(b, "RG")) != nullptr && *p == 'Z';
            if (r.has_rg) r.rg = (const char *)(p + 1);
            r.endpos = rec_endpos(r);
            emit(r);
        }
        if (ret < -1) error = true;
        hts_itr_destroy(it);
    }
};
#endif

// ------------------------------------------------------------------------------------------
// Per-read warning lines (R:src/lib/bamrc/ReadWarnings.hpp:12-50; call sites R:BasicStat.cpp:85,100 and R:bamreadcount.cpp:282).
// The engine returns per-type event COUNTS; the text names reads, in the order the reference meets the events: region by
// region, site by site, spanning reads in file order, and inside one event process_read's own order (SM before NM; an
// indel event calls process_read for the indel key and again for the base key).  The host replays exactly that walk over the
// reads that can warn at all (no NM tag, proper pair without SM tag, no library under -p) until every type has used up its
// -w budget, then stops collecting.  With -w -1 the reference prints one line per offending EVENT without bound; this host
// prints the first 1000 per type and then the engine's exact total (the one documented STDERR difference).
// ------------------------------------------------------------------------------------------
struct Warner {
    enum { SM = 0, NM = 1, ZM = 2, LIB = 3, NT = 4 };
    struct Cand { int32_t pos; int64_t endpos; uint16_t flag; uint8_t mapq; bool no_nm, no_sm, no_lib; std::vector<uint32_t> cigar; std::vector<uint8_t> qual; std::string qname; };
    struct Reg { int tid, beg, end; bool skip_halo; size_t c0, c1; };
    long long max_per_type; bool unlimited; int min_mapq, min_bq; bool per_lib, ic;
    long long counts[NT] = {0, 0, 0, 0};
    std::vector<Cand> cands; std::vector<Reg> regs;
    Warner(long long mw, int q, int b, bool p, bool i) : max_per_type(mw < 0 ? 1000 : mw), unlimited(mw < 0), min_mapq(q), min_bq(b), per_lib(p), ic(i) {}
    bool budget_left() const { return counts[SM] < max_per_type || counts[NM] < max_per_type || (per_lib && counts[LIB] < max_per_type); }
    bool collecting() const { return max_per_type > 0 && budget_left(); }
    void begin_region(int tid, int beg, int end, bool skip_halo) { regs.push_back({tid, beg, end, skip_halo, cands.size(), cands.size()}); }
    void consider(const Rec &r, bool no_lib) {
        if (!collecting() || (r.flag & 4) || regs.empty()) return;
        const bool no_nm = r.nm == BRC_TAG_ABSENT, no_sm = (r.flag & 2) && r.sm == BRC_TAG_ABSENT;
        if (!no_nm && !no_sm && !(per_lib && no_lib)) return;
        Cand c; c.pos = r.pos; c.endpos = r.endpos; c.flag = r.flag; c.mapq = r.mapq; c.no_nm = no_nm; c.no_sm = no_sm; c.no_lib = per_lib && no_lib;
        c.cigar.assign(r.cigar, r.cigar + r.n_cigar); c.qual.assign(r.qual, r.qual + r.l_qseq);
        c.qname = (const char *)(r.data.data() + 32);
        cands.push_back(std::move(c)); regs.back().c1 = cands.size();
    }
    // candidates collected by a parallel decode (ParallelFetcher), already in file order
    void take(std::vector<Cand> &more) {
        if (!collecting() || regs.empty()) { more.clear(); return; }
        for (Cand &c : more) cands.push_back(std::move(c));
        more.clear(); regs.back().c1 = cands.size();
    }
    void emit(int type, const std::string &qname) {
        static const char *msg[NT] = {"Couldn't find single-end mapping quality. Check to see if the SM tag is in BAM.",
                                      "Couldn't find number of mismatches. Check to see if the NM tag is in BAM.",
                                      "Couldn't find the generated tag.",
                                      "Library unavailable. Check to make sure the LB tag is present in the @RG entries of the header."};
        ++counts[type];
        if (counts[type] > max_per_type) return;
        std::fprintf(stderr, "WARNING: In read %s: %s\n", qname.c_str(), msg[type]);
        if (!unlimited && counts[type] == max_per_type) std::fprintf(stderr, "The previous warning has been emitted %lld times and will be disabled.\n", counts[type]);
    }
    // stateless resolve_cigar2 (V:htslib-1.10/sam.c:3964-4041): qpos / is_del / indel of `site` in a read
    static bool resolve(const Cand &c, int64_t site, int &qpos, int &indel) {
        int64_t x = c.pos; int y = 0; size_t k = 0; uint32_t op = 0; int len = 0; const size_t n = c.cigar.size();
        auto refop = [](uint32_t o) { return o == 0 || o == 2 || o == 3 || o == 7 || o == 8; };
        auto matchop = [](uint32_t o) { return o == 0 || o == 7 || o == 8; };
        for (; k < n; ++k) {
            op = c.cigar[k] & 0xF; len = (int)(c.cigar[k] >> 4);
            if (refop(op)) { if (site < x + len) break; x += len; if (matchop(op)) y += len; }
            else if (op == 1 || op == 4) y += len;
        }
        indel = 0;
        if (k >= n) return false;
        const bool is_del = !matchop(op);
        qpos = is_del ? y : y + (int)(site - x);
        if (x + len - 1 == site && k + 1 < n) {
            const uint32_t op2 = c.cigar[k + 1] & 0xF; const int l2 = (int)(c.cigar[k + 1] >> 4);
            if (op2 == 2) indel = -l2;
            else if (op2 == 1) indel = l2;
            else if (op2 == 6) { int l3 = 0; for (size_t m = k + 2; m < n; ++m) { const uint32_t o3 = c.cigar[m] & 0xF; if (o3 == 1) l3 += (int)(c.cigar[m] >> 4); else if (refop(o3)) break; } if (l3 > 0) indel = l3; }
        }
        return !is_del;
    }
    void replay() {
        for (const Reg &g : regs) {
            if (!budget_left()) break;
            if (g.c0 == g.c1) continue;
            const int64_t first = g.skip_halo ? g.beg : std::max(g.beg - 1, 0);
            size_t lo = g.c0;
            for (int64_t p = std::max<int64_t>(first, cands[g.c0].pos); p < g.end && budget_left(); ++p) {
                while (lo < g.c1 && cands[lo].endpos <= p) ++lo;          // leading reads that ended (later ones are re-checked below)
                if (lo >= g.c1) break;
                if (cands[lo].pos > p) { p = cands[lo].pos - 1; continue; }
                for (size_t i = lo; i < g.c1 && cands[i].pos <= p; ++i) {
                    const Cand &c = cands[i];
                    if (c.endpos <= p) continue;
                    if (c.no_lib) { emit(LIB, c.qname); break; }            // pileup_func returns: nothing after it at this site (R:...:281-284)
                    int qpos = 0, indel = 0;
                    if (!resolve(c, p, qpos, indel)) continue;               // is_del
                    if ((int)c.mapq < min_mapq || qpos >= (int)c.qual.size() || (int)c.qual[(size_t)qpos] < min_bq || (c.flag & (4 | 256 | 512 | 1024))) continue;
                    const int calls = (indel != 0 ? 1 : 0) + ((indel < 1 || !ic) ? 1 : 0);
                    for (int k = 0; k < calls; ++k) { if (c.no_sm) emit(SM, c.qname); if (c.no_nm) emit(NM, c.qname); }
                }
            }
        }
        cands.clear(); regs.clear();
    }
    void finish(const int64_t engine_counts[4]) const {
        if (!unlimited) return;
        static const char *nm[NT] = {"SM tag missing", "NM tag missing", "generated tag missing", "library unavailable"};
        for (int t = 0; t < NT; ++t) if (engine_counts[t] > max_per_type) std::fprintf(stderr, "WARNING: %s: %lld events in total (only the first %lld are listed)\n", nm[t], (long long)engine_counts[t], max_per_type);
    }
};

// ------------------------------------------------------------------------------------------
// Parallel decode of one big fetch (a window of a cut region): the position range is cut at 16 kb linear-index boundaries
// into one sub-range per thread; every thread has its own BGZF reader, seeks through the index and keeps the records whose
// START lies in its sub-range (the first thread also keeps the earlier records that reach into the fetch), so the slices
// concatenated in thread order are exactly samfetch's records in file order.  Each slice is decoded straight into the
// struct-of-arrays layout of brc_read_batch; the slices are then copied (in parallel) into ONE page-locked batch that
// brc_push_reads borrows, so the engine's pipelined upload runs out of it with no further host copy.
// Unmapped records are dropped here (the pileup buffer refuses them, V:htslib-1.10/sam.c:4488-4490).
// ------------------------------------------------------------------------------------------
struct Slice {
    std::vector<int32_t> pos, l_qseq, nm, sm; std::vector<uint16_t> flag, lib; std::vector<uint8_t> mapq;
    std::vector<uint64_t> cigar_off, seq_off, qual_off; std::vector<uint32_t> cigar; std::vector<uint8_t> seq, qual;
    std::vector<Warner::Cand> cands; bool cand_overflow = false, error = false; uint64_t n_decoded = 0;
    void clear() {
        pos.clear(); l_qseq.clear(); nm.clear(); sm.clear(); flag.clear(); lib.clear(); mapq.clear();
        cigar_off.assign(1, 0); seq_off.assign(1, 0); qual_off.assign(1, 0); cigar.clear(); seq.clear(); qual.clear();
        cands.clear(); cand_overflow = false; error = false; n_decoded = 0;
    }
    size_t n() const { return pos.size(); }
};

struct PinnedBuf {   // grow-only host buffer: page-locked through the engine when there is one (brc_host_alloc), plain otherwise
    void *p = nullptr; size_t cap = 0; bool pinned = false;
    bool reserve(size_t bytes, bool want_pinned) {
        if (bytes <= cap) return true;
        release();
        const size_t want = bytes + bytes / 8 + 4096;
        if (want_pinned && brc_host_alloc(want, &p) == BRC_OK && p) { pinned = true; cap = want; return true; }
        p = std::malloc(want); pinned = false; cap = p ? want : 0;
        return p != nullptr;
    }
    void release() { if (p) { if (pinned) brc_host_free(p); else std::free(p); } p = nullptr; cap = 0; }
    ~PinnedBuf() { release(); }
    PinnedBuf() = default; PinnedBuf(const PinnedBuf &) = delete; PinnedBuf &operator=(const PinnedBuf &) = delete;
};

struct WindowJob {   // one decoded window: the slices and the concatenated batch
    std::vector<Slice> slices;
    PinnedBuf buf[13];
    brc_read_batch batch{};
    uint64_t n_decoded = 0; bool error = false, cand_overflow = false;
    int tid = -1; int64_t fbeg = 0, fend = 0;
};

struct ParallelFetcher {
    const BamFile &bam; std::string path; int n_threads; bool per_lib, want_pinned;
    std::unordered_map<std::string, uint16_t> rg_lib;      // @RG ID -> library rank
    std::vector<std::unique_ptr<Bgzf>> readers;
    static constexpr size_t MAX_CANDS = 65536;

    ParallelFetcher(const BamFile &b, const std::string &p, bool pl, bool pin) : bam(b), path(p), per_lib(pl), want_pinned(pin) {
        unsigned hw = std::thread::hardware_concurrency();
        n_threads = (int)std::max(1u, std::min(hw ? hw : 1u, 16u));
        if (const char *ov = std::getenv("BRC_CLI_DECODE_THREADS")) n_threads = std::max(1, std::atoi(ov));
    }
    static bool eligible(int64_t fbeg, int64_t fend) { return fend - fbeg >= (int64_t(1) << 18); }   // >= 16 linear-index windows

    // records of contig `tid` with lo <= pos < hi (first slice: also pos < lo with endpos > fbeg), in file order
    void decode_slice(Bgzf &bz, int tid, int64_t lo, int64_t hi, bool first, int64_t fbeg, bool collect, Slice &out) const {
        out.clear();
        uint64_t voff;
        if (!bam.query_offset(tid, lo, voff)) return;
        if (!bz.seek(voff)) { out.error = bz.error; return; }
        Rec r;
        for (;;) {
            if (!read_record(bz, r)) { out.error = bz.error; break; }
            if (r.tid != tid || r.pos >= hi) break;
            ++out.n_decoded;
            if (r.pos < lo) {
                if (!first) continue;
                r.endpos = rec_endpos(r);
                if (r.endpos <= fbeg) continue;
            }
            if (r.flag & 4) continue;
            uint16_t lib = 0; bool no_lib = false;
            if (per_lib) {
                lib = (uint16_t)BRC_LIB_NONE;
                if (r.has_rg) { auto it = rg_lib.find(r.rg); if (it != rg_lib.end()) lib = it->second; }
                no_lib = lib == (uint16_t)BRC_LIB_NONE;
            }
            if (collect && !out.cand_overflow) {
                const bool no_nm = r.nm == BRC_TAG_ABSENT, no_sm = (r.flag & 2) && r.sm == BRC_TAG_ABSENT;
                if (no_nm || no_sm || no_lib) {
                    if (out.cands.size() >= MAX_CANDS) out.cand_overflow = true;
                    else {
                        Warner::Cand c; c.pos = r.pos; c.endpos = rec_endpos(r); c.flag = r.flag; c.mapq = r.mapq; c.no_nm = no_nm; c.no_sm = no_sm; c.no_lib = no_lib;
                        c.cigar.assign(r.cigar, r.cigar + r.n_cigar); c.qual.assign(r.qual, r.qual + r.l_qseq); c.qname = (const char *)(r.data.data() + 32);
                        out.cands.push_back(std::move(c));
                    }
                }
            }
            out.pos.push_back(r.pos); out.flag.push_back(r.flag); out.mapq.push_back(r.mapq); out.lib.push_back(lib); out.l_qseq.push_back(r.l_qseq);
            out.nm.push_back(r.nm); out.sm.push_back(r.sm);
            out.cigar.insert(out.cigar.end(), r.cigar, r.cigar + r.n_cigar); out.cigar_off.push_back(out.cigar.size());
            out.seq.insert(out.seq.end(), r.seq, r.seq + ((size_t)r.l_qseq + 1) / 2); out.seq_off.push_back(out.seq.size());
            out.qual.insert(out.qual.end(), r.qual, r.qual + r.l_qseq); out.qual_off.push_back(out.qual.size());
        }
    }

    // decode [fbeg, fend) of `tid` into job.slices and job.batch; false when a reader could not be opened
    bool run(int tid, int64_t fbeg, int64_t fend, bool collect, WindowJob &job) {
        job.tid = tid; job.fbeg = fbeg; job.fend = fend; job.error = false; job.cand_overflow = false; job.n_decoded = 0;
        const int64_t w0 = fbeg >> 14, w1 = (fend + 16383) >> 14;
        const int T = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, (w1 - w0) / 4));
        while ((int)readers.size() < T) { readers.emplace_back(new Bgzf()); if (!readers.back()->open_worker(path)) return false; }
        job.slices.resize((size_t)T);
        std::vector<int64_t> cut((size_t)T + 1);
        for (int t = 0; t <= T; ++t) cut[(size_t)t] = t == 0 ? fbeg : t == T ? fend : ((w0 + (w1 - w0) * t / T) << 14);
        auto work = [&](int t) { decode_slice(*readers[(size_t)t], tid, cut[(size_t)t], cut[(size_t)t + 1], t == 0, fbeg, collect, job.slices[(size_t)t]); };
        const bool timing = std::getenv("BRC_CLI_TIMING") != nullptr;
        auto clk = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t0 = clk();
        {
            std::vector<std::thread> th;
            for (int t = 1; t < T; ++t) th.emplace_back(work, t);
            work(0);
            for (auto &x : th) x.join();
        }
        const double t1 = clk();
        // ---- concatenate: prefix sums, then every thread copies its slice ----
        std::vector<size_t> rn((size_t)T + 1, 0), cn((size_t)T + 1, 0), sn((size_t)T + 1, 0), qn((size_t)T + 1, 0);
        for (int t = 0; t < T; ++t) {
            const Slice &sl = job.slices[(size_t)t];
            rn[(size_t)t + 1] = rn[(size_t)t] + sl.n(); cn[(size_t)t + 1] = cn[(size_t)t] + sl.cigar.size();
            sn[(size_t)t + 1] = sn[(size_t)t] + sl.seq.size(); qn[(size_t)t + 1] = qn[(size_t)t] + sl.qual.size();
            job.error = job.error || sl.error; job.cand_overflow = job.cand_overflow || sl.cand_overflow; job.n_decoded += sl.n_decoded;
        }
        const size_t n = rn[(size_t)T];
        const size_t bytes[13] = {n * 4, n * 2, n, n * 2, n * 4, n * 4, n * 4, (n + 1) * 8, cn[(size_t)T] * 4, (n + 1) * 8, sn[(size_t)T], (n + 1) * 8, qn[(size_t)T]};
        for (int k = 0; k < 13; ++k) if (!job.buf[k].reserve(bytes[k] + 64, want_pinned)) return false;
        int32_t *pos = (int32_t *)job.buf[0].p; uint16_t *flag = (uint16_t *)job.buf[1].p; uint8_t *mapq = (uint8_t *)job.buf[2].p; uint16_t *lib = (uint16_t *)job.buf[3].p;
        int32_t *lq = (int32_t *)job.buf[4].p, *nm = (int32_t *)job.buf[5].p, *sm = (int32_t *)job.buf[6].p;
        uint64_t *coff = (uint64_t *)job.buf[7].p; uint32_t *cig = (uint32_t *)job.buf[8].p; uint64_t *soff = (uint64_t *)job.buf[9].p; uint8_t *seq = (uint8_t *)job.buf[10].p;
        uint64_t *qoff = (uint64_t *)job.buf[11].p; uint8_t *qual = (uint8_t *)job.buf[12].p;
        auto copy = [&](int t) {
            const Slice &sl = job.slices[(size_t)t];
            const size_t r0 = rn[(size_t)t], m = sl.n();
            if (m) {
                std::memcpy(pos + r0, sl.pos.data(), m * 4); std::memcpy(flag + r0, sl.flag.data(), m * 2); std::memcpy(mapq + r0, sl.mapq.data(), m);
                std::memcpy(lib + r0, sl.lib.data(), m * 2); std::memcpy(lq + r0, sl.l_qseq.data(), m * 4); std::memcpy(nm + r0, sl.nm.data(), m * 4);
                std::memcpy(sm + r0, sl.sm.data(), m * 4);
                if (!sl.cigar.empty()) std::memcpy(cig + cn[(size_t)t], sl.cigar.data(), sl.cigar.size() * 4);
                if (!sl.seq.empty()) std::memcpy(seq + sn[(size_t)t], sl.seq.data(), sl.seq.size());
                if (!sl.qual.empty()) std::memcpy(qual + qn[(size_t)t], sl.qual.data(), sl.qual.size());
                for (size_t i = 0; i < m; ++i) { coff[r0 + i] = cn[(size_t)t] + sl.cigar_off[i]; soff[r0 + i] = sn[(size_t)t] + sl.seq_off[i]; qoff[r0 + i] = qn[(size_t)t] + sl.qual_off[i]; }
            }
            if (t == T - 1) { coff[n] = cn[(size_t)T]; soff[n] = sn[(size_t)T]; qoff[n] = qn[(size_t)T]; }
        };
        {
            std::vector<std::thread> th;
            for (int t = 1; t < T; ++t) th.emplace_back(copy, t);
            copy(0);
            for (auto &x : th) x.join();
        }
        if (timing) std::fprintf(stderr, "[brc timing] window %d:%lld-%lld: %d threads decode %.3fs, concatenate %.3fs (%zu reads)\n", tid, (long long)fbeg, (long long)fend, T, t1 - t0, clk() - t1, n);
        brc_read_batch &b = job.batch;
        b = brc_read_batch{};
        b.n_reads = (int64_t)n; b.tid = nullptr; b.pos = pos; b.flag = flag; b.mapq = mapq; b.lib = lib; b.l_qseq = lq; b.nm = nm; b.sm = sm;
        b.cigar_off = coff; b.cigar = cig; b.seq_off = soff; b.seq = seq; b.qual_off = qoff; b.qual = qual;
        return true;
    }
};

// ------------------------------------------------------------------------------------------
// FASTA + .fai (fai_fetch of a whole chromosome, R:...:87)
// ------------------------------------------------------------------------------------------
struct Fasta {
    std::string path;
    struct Ent { int64_t len, off, lb, lw; };
    std::map<std::string, Ent> ents;
    bool open(const std::string &p) {
        path = p;
        std::ifstream f(p + ".fai");
        if (!f) return false;
        std::string line;
        while (std::getline(f, line)) { std::istringstream ss(line); std::string n; Ent e; if (ss >> n >> e.len >> e.off >> e.lb >> e.lw) ents[n] = e; }
        return true;
    }
    bool fetch(const std::string &name, std::string &out) const {
        auto it = ents.find(name);
        if (it == ents.end()) return false;
        const Ent &e = it->second;
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) return false;
        const int64_t n_lines = (e.len + e.lb - 1) / e.lb;
        std::vector<char> raw((size_t)(n_lines * e.lw + 8));
        fseeko(f, (off_t)e.off, SEEK_SET);
        const size_t got = std::fread(raw.data(), 1, raw.size(), f);
        std::fclose(f);
        out.clear(); out.reserve((size_t)e.len);
        for (size_t i = 0; i < got && (int64_t)out.size() < e.len; ++i) if (raw[i] != '\n' && raw[i] != '\r') out.push_back(raw[i]);
        return (int64_t)out.size() == e.len;
    }
};

void usage() {
    std::printf("Usage: bam-readcount [OPTIONS] bam_file|cram_file [region]\nGenerate metrics for bam_file at single nucleotide positions.\n"
                "Example: bam-readcount -f ref.fa some.bam|some.cram\n\nAvailable options:\n"
                "  -h [ --help ]                         produce this message\n"
                "  -v [ --version ]                      output the version number\n"
                "  -q [ --min-mapping-quality ] arg (=0) minimum mapping quality of reads used for counting.\n"
                "  -b [ --min-base-quality ] arg (=0)    minimum base quality at a position to use the read for counting.\n"
                "  -d [ --max-count ] arg (=10000000)    max depth to avoid excessive memory usage.\n"
                "  -l [ --site-list ] arg                file containing a list of regions to report readcounts within.\n"
                "  -f [ --reference-fasta ] arg          reference sequence in the fasta format.\n"
                "  -D [ --print-individual-mapq ] arg    report the mapping qualities as a comma separated list.\n"
                "  -p [ --per-library ]                  report results by library.\n"
                "  -w [ --max-warnings ] arg             maximum number of warnings of each type to emit. -1 gives an unlimited number.\n"
                "  -i [ --insertion-centric ]            generate indel centric readcounts. Reads containing insertions will not be\n"
                "                                        included in per-base counts\n"
                "  --shard RANK/COUNT                    (this host) compute only shard RANK of COUNT: the regions are cut into COUNT runs of\n"
                "                                        about equal BAI-estimated coverage, one process per GPU (BRC_DEVICE); outputs\n"
                "                                        concatenate in rank order\n\n");
}

// samtools region string "name[:beg[-end]]" as bam_parse_region (V:bam_aux.c:65-75) handles it.  Returns 0 when beg and end were
// set; -1 when only the contig is known — a bare name, an open end ("chr:100", "chr:100-": the 64-bit end exceeds INT_MAX so the
// legacy wrapper bails out after setting the contig) or unparsable coordinates: the caller's beg/end keep their previous values
// (initially 0 .. 0x7fffffff; probe with the reference binary: "21:10405200" prints the whole contig, and as a second region it
// repeats the previous one).  Thousands separators are accepted ("21:10,402,985-10,402,990").
int parse_region(const BamFile &bam, const std::string &s, int &tid, int &beg, int &end) {
    std::string name = s; tid = -1;
    const size_t colon = s.rfind(':');
    bool ranged = false;
    int64_t b = 0, e = 0;
    if (colon != std::string::npos && bam.tid_of.find(s) == bam.tid_of.end()) {
        std::string coords = s.substr(colon + 1); name = s.substr(0, colon);
        coords.erase(std::remove(coords.begin(), coords.end(), ','), coords.end());
        const char *c = coords.c_str();
        auto digits = [](const char *&q, int64_t &v) { const char *q0 = q; v = 0; while (*q >= '0' && *q <= '9') { v = v * 10 + (*q - '0'); ++q; } return q != q0; };
        if (*c == '-') { ++c; b = 1; ranged = digits(c, e); }                     // "chr:-end": from the first base
        else if (digits(c, b) && *c == '-') { ++c; ranged = digits(c, e); }       // "chr:beg-end"; "chr:beg" / "chr:beg-" stay open
        if (ranged) b = b > 0 ? b - 1 : 0;
    }
    auto it = bam.tid_of.find(name);
    if (it == bam.tid_of.end()) return -1;
    tid = it->second;
    if (!ranged) return -1;
    beg = (int)std::min<int64_t>(b, 0x7fffffff); end = (int)std::min<int64_t>(e, 0x7fffffff);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    const double t_main0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    int min_mapq = 0, min_bq = 0, max_cnt = 10000000; bool per_lib = false, ic = false; long long max_warn = -1;
    std::string fn_pos, fn_fa, dist_arg;
    static option lo[] = {{"help", 0, 0, 'h'}, {"version", 0, 0, 'v'}, {"min-mapping-quality", 1, 0, 'q'}, {"min-base-quality", 1, 0, 'b'},
                          {"max-count", 1, 0, 'd'}, {"site-list", 1, 0, 'l'}, {"reference-fasta", 1, 0, 'f'}, {"print-individual-mapq", 1, 0, 'D'},
                          {"per-library", 0, 0, 'p'}, {"max-warnings", 1, 0, 'w'}, {"insertion-centric", 0, 0, 'i'}, {"shard", 1, 0, 1000}, {0, 0, 0, 0}};
    int shard_rank = 0, shard_count = 1;
    bool help = false, version = false;
    for (int c; (c = getopt_long(argc, argv, "hvq:b:d:l:f:D:pw:i", lo, nullptr)) != -1;) {
        switch (c) {
        case 'h': help = true; break; case 'v': version = true; break;
        case 'q': min_mapq = std::atoi(optarg); break; case 'b': min_bq = std::atoi(optarg); break; case 'd': max_cnt = std::atoi(optarg); break;
        case 'l': fn_pos = optarg; break; case 'f': fn_fa = optarg; break; case 'D': dist_arg = optarg; break;
        case 'p': per_lib = true; break; case 'w': max_warn = std::atoll(optarg); break; case 'i': ic = true; break;
        case 1000: if (std::sscanf(optarg, "%d/%d", &shard_rank, &shard_count) != 2 || shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) { std::fprintf(stderr, "--shard wants RANK/COUNT with 0 <= RANK < COUNT\n"); return 1; } break;
        default: usage(); return 1;
        }
    }
    if (version) { std::printf("bam-readcount version: b200 (engine ABI %d)\n", brc_abi_version()); return 1; }   // R:...:467-470 (exit 1)
    if (help || optind >= argc) { usage(); return 1; }                                                              // R:...:472-475
    const std::string bam_path = argv[optind];
    std::vector<std::string> region_args(argv + optind + 1, argv + argc);
    std::fprintf(stderr, "Minimum mapping quality is set to %d\n", min_mapq);
    if (dist_arg == "1" || dist_arg == "true") { std::fprintf(stderr, "Not currently supporting distributions\n"); return 1; }

    // CUDA context creation takes a second or two: do it while the BAM header, index and FASTA index are read
    brc_config cfg{}; cfg.min_mapq = min_mapq; cfg.min_bq = min_bq; cfg.max_cnt = max_cnt; cfg.per_lib = per_lib; cfg.insertion_centric = ic;
    cfg.n_libs = 0; cfg.device = std::getenv("BRC_DEVICE") ? std::atoi(std::getenv("BRC_DEVICE")) : 0;
    const bool decode_only = std::getenv("BRC_CLI_DECODE_ONLY") != nullptr;   // test hook: exercise BGZF/BAI/region fetch without a GPU
    int warm_rc = BRC_OK;
    brc_config warm_cfg = cfg; warm_cfg.per_lib = 0;
    std::thread warm([&warm_rc, warm_cfg, decode_only] { if (decode_only) return; brc_engine *tmp = nullptr; warm_rc = brc_create(&warm_cfg, &tmp); if (tmp) brc_destroy(tmp); });

    BamFile bam;
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{warm};
    const bool is_cram = bam_path.size() > 5 && bam_path.substr(bam_path.size() - 5) == ".cram";
#ifdef BRC_WITH_HTSLIB
    HtsSource hts;
    if (is_cram) { if (!hts.open(bam_path, fn_fa, bam)) { std::fprintf(stderr, "Fail to open BAM file %s\n", bam_path.c_str()); return 1; } }
    else
#else
    if (is_cram) { std::fprintf(stderr, "CRAM input needs a host built with htslib (tools/build_htslib.sh, then python -m bam_readcount_b200.build); convert to BAM\n"); return 1; }
#endif
    if (!bam.open(bam_path)) { std::fprintf(stderr, "Fail to open BAM file %s\n", bam_path.c_str()); return 1; }
    Fasta fa;
    const bool have_fa = !fn_fa.empty() && fa.open(fn_fa);
    if (!fn_fa.empty() && !have_fa) { std::fprintf(stderr, "Fail to open reference file %s\n", fn_fa.c_str()); return 1; }

    // @RG ID -> LB; libraries in std::set order (R:...:92-111, 526-529)
    std::map<std::string, std::string> rg_lb; std::set<std::string> libs;
    {
        std::istringstream ss(bam.text); std::string line;
        while (std::getline(ss, line)) {
            if (line.compare(0, 3, "@RG") != 0) continue;
            std::istringstream ls(line); std::string tok, id, lb; bool has_lb = false;
            while (std::getline(ls, tok, '\t')) { if (tok.compare(0, 3, "ID:") == 0) id = tok.substr(3); else if (tok.compare(0, 3, "LB:") == 0) { lb = tok.substr(3); has_lb = true; } }
            if (has_lb) { libs.insert(lb); if (!id.empty() && !rg_lb.count(id)) rg_lb[id] = lb; }
        }
    }
    for (const auto &l : libs) std::fprintf(stderr, "Expect library: %s in BAM\n", l.c_str());
    std::vector<std::string> lib_names(libs.begin(), libs.end());
    std::map<std::string, uint16_t> lib_rank;
    for (size_t i = 0; i < lib_names.size(); ++i) lib_rank[lib_names[i]] = (uint16_t)i;
    std::vector<const char *> lib_ptrs; for (auto &s : lib_names) lib_ptrs.push_back(s.c_str());
    if (lib_ptrs.empty()) lib_ptrs.push_back("");

    if (fn_pos.empty() && region_args.empty()) {
        std::fprintf(stderr, "Whole-file mode is not supported (the reference skips its per-read pre-processing there, R:...:624); give regions or -l\n");
        return 1;
    }
#ifdef BRC_WITH_HTSLIB
    if (is_cram) { if (!hts.load_index(bam_path)) { std::fprintf(stderr, "BAM indexing file is not available.\n"); return 1; } }
    else
#endif
    if (!bam.load_index(bam_path)) { std::fprintf(stderr, "BAM indexing file is not available.\n"); return 1; }
    if (!have_fa && !decode_only) { std::fprintf(stderr, "A reference FASTA (-f) is required in region / site-list mode\n"); return 1; }

    cfg.n_libs = (int32_t)lib_names.size();
    brc_engine *eng = nullptr;      // created below, once the first window is already being decoded (the CUDA context is still coming up)
    int rc = BRC_OK;

    struct Region { int tid, beg, end; bool site_list; bool cont = false; };   // cont: a later window of a cut region (its halo site belongs to the window before)
    std::vector<Region> regions;
    if (!fn_pos.empty()) {
        std::ifstream fp(fn_pos);
        if (!fp) { std::fprintf(stderr, "Failed to open region list file: %s\n", fn_pos.c_str()); return 1; }
        std::string line;
        while (std::getline(fp, line)) {
            std::istringstream ss(line); std::string name; int b, e;
            if (!(ss >> name >> b >> e)) continue;
            auto it = bam.tid_of.find(name);
            if (it == bam.tid_of.end()) { std::fprintf(stderr, "%s not found in bam file. Region %s %i %i skipped.\n", name.c_str(), name.c_str(), b, e); continue; }
            regions.push_back({it->second, b - 1, e, true});
        }
    } else {
        int beg = 0, end = 0x7fffffff;
        for (const auto &rs : region_args) {
            int tid;
            parse_region(bam, rs, tid, beg, end);
            if (tid < 0) { std::fprintf(stderr, "Invalid region %s\n", rs.c_str()); brc_destroy(eng); return 1; }
            regions.push_back({tid, beg, end, false});
        }
    }

    // Long regions are cut into consecutive windows so host staging and result buffers stay bounded (a 250 Mb chromosome
    // at 30x would otherwise need tens of GB).  A window is just a region of the -l loop: it recomputes the site to its left
    // (the 1-site halo), so the concatenated output equals the unsplit region's.  Only the last window of an argv region keeps
    // argv semantics.  (With a tiny -d the max-count rule sees the window's own fetch order)
    {
        // 2 Mb windows: each is decoded by all threads in ~40 ms one window ahead of the engine and needs 110 MB of page-locked memory
        // (8 Mb windows were 0.6 s slower on a 10 Mb BAM, r02t); the record-by-record paths (CRAM, BRC_CLI_SEQUENTIAL) keep 8 Mb
        const bool par_windows_ok = !is_cram && std::getenv("BRC_CLI_DEVICE_DECODE") == nullptr && std::getenv("BRC_CLI_SEQUENTIAL") == nullptr;
        int64_t W = std::getenv("BRC_CLI_WINDOW") ? std::atoll(std::getenv("BRC_CLI_WINDOW")) : (par_windows_ok ? 2000000 : 8000000);
        if (shard_count > 1 && !std::getenv("BRC_CLI_WINDOW")) W = 1000000;      // finer units so the shards can balance
        std::vector<Region> cut;
        for (const Region &g : regions) {
            const int64_t clen = bam.lens[(size_t)g.tid];
            const int64_t e_eff = std::min<int64_t>(g.end, std::max<int64_t>(clen, (int64_t)g.beg + 1));
            if (e_eff - g.beg <= W) { cut.push_back(g); continue; }
            for (int64_t b = g.beg; b < e_eff; b += W) {
                const bool last = b + W >= e_eff;
                Region w{g.tid, (int)b, last ? g.end : (int)(b + W), last ? g.site_list : true};
                w.cont = b != g.beg;
                cut.push_back(w);
            }
        }
        regions.swap(cut);
    }

    // --shard RANK/COUNT: one process per GPU.  The (windowed) regions are the units; each unit's weight is the
    // compressed-byte span the BAI linear index gives for it — the coverage estimate the index offers without touching the data
    // — and the units are cut into COUNT contiguous runs of about equal weight; this process computes run RANK.  Units are
    // independent (every window recomputes its left halo site), so the concatenation of the ranks' outputs in rank order is the
    // unsharded output; only the reference's never-cleared argv deletion queue does not cross a shard boundary.
    if (shard_count > 1) {
        std::vector<double> wgt(regions.size(), 1.0);
        for (size_t i = 0; i < regions.size(); ++i) {
            const Region &g = regions[i];
            const int64_t clen = bam.lens[(size_t)g.tid];
            const int64_t e_eff = std::min<int64_t>(g.end, std::max<int64_t>(clen, (int64_t)g.beg + 1));
            double w = (double)std::max<int64_t>(e_eff - g.beg, 1) * 0.05;                 // no index information: 0.05 bytes per base
            if (g.tid < (int)bam.idx.size() && !bam.idx[(size_t)g.tid].linear.empty()) {
                const auto &lin = bam.idx[(size_t)g.tid].linear;
                auto off_at = [&](int64_t p) { int64_t k = std::min<int64_t>(std::max<int64_t>(p >> 14, 0), (int64_t)lin.size() - 1); while (k > 0 && lin[(size_t)k] == 0) --k; return (double)(lin[(size_t)k] >> 16); };
                const double d = off_at(e_eff + 16384) - off_at(g.beg);
                if (d > 0) w = d * (double)(e_eff - g.beg) / (double)((((e_eff + 16384) >> 14) - (g.beg >> 14)) * 16384);
            }
            wgt[i] = w;
        }
        double tot = 0; for (double w : wgt) tot += w;
        std::vector<size_t> cutpt((size_t)shard_count + 1, regions.size()); cutpt[0] = 0;
        { double acc = 0; int r = 1; for (size_t i = 0; i < regions.size() && r < shard_count; ++i) { acc += wgt[i]; while (r < shard_count && acc >= tot * r / shard_count) cutpt[(size_t)r++] = i + 1; } }
        std::vector<Region> mine(regions.begin() + (long)cutpt[(size_t)shard_rank], regions.begin() + (long)cutpt[(size_t)shard_rank + 1]);
        if (std::getenv("BRC_CLI_TIMING")) std::fprintf(stderr, "[brc shard] %d/%d: units %zu..%zu of %zu\n", shard_rank, shard_count, cutpt[(size_t)shard_rank], cutpt[(size_t)shard_rank + 1], regions.size());
        regions.swap(mine);
    }

    std::set<int> ref_loaded;
    std::string chrom;
    const bool timing = std::getenv("BRC_CLI_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_decode = 0, t_compute = 0, t_format = 0, t_write = 0, t_ref = 0, t_results = 0, t_reset = 0;
    // BRC_CLI_DEVICE_DECODE=1: BGZF inflate + BAM framing on the GPU (per-read warning lines need host-decoded reads: counts only)
    const bool device_decode = std::getenv("BRC_CLI_DEVICE_DECODE") != nullptr && !decode_only;
    std::vector<std::string> rg_id_store; std::vector<const char *> rg_ids; std::vector<uint16_t> rg_libs;
    for (const auto &kv : rg_lb) rg_id_store.push_back(kv.first);
    for (const auto &id : rg_id_store) { rg_ids.push_back(id.c_str()); rg_libs.push_back(lib_rank[rg_lb[id]]); }
    Warner warner(max_warn, min_mapq, min_bq, per_lib, ic);
    int64_t warn_total[4] = {0, 0, 0, 0};
    auto flush = [&]() -> int {
        const double c0 = now();
        int r = brc_compute(eng);
        t_compute += now() - c0;
        if (r != BRC_OK) { std::fprintf(stderr, "brc_compute: %s\n", brc_last_error(eng)); return r; }
        brc_results res{};
        const double g0 = now();
        if ((r = brc_get_results(eng, &res)) != BRC_OK) { std::fprintf(stderr, "brc_get_results: %s\n", brc_last_error(eng)); return r; }
        std::fflush(stdout);
        const double f0 = now();
        t_results += f0 - g0;
        const bool argv_chain = res.n_regions > 1 && !res.regions[0].site_list_mode;   // never-cleared deletion queue: one sequential pass
        int64_t total_slots = 0;
        for (int64_t g = 0; g < res.n_regions; ++g) total_slots += res.regions[g].n_slots;
        const bool many_small = res.n_regions > 1 && total_slots <= (int64_t(1) << 22);  // a site list: all regions in one formatting pass
        if (argv_chain || many_small) {
            if (brc_write_text(eng, -1, 0, -1, lib_ptrs.data(), STDOUT_FILENO) < 0) { std::fprintf(stderr, "format: %s\n", brc_last_error(eng)); return -1; }
        } else {
            const int64_t WIN = 1 << 21;   // stream big regions in 2M-site windows (each formatted by several threads)
            for (int64_t g = 0; g < res.n_regions; ++g)
                for (int64_t first = 0; first < res.regions[g].n_slots; first += WIN)
                    if (brc_write_text(eng, g, first, WIN, lib_ptrs.data(), STDOUT_FILENO) < 0) { std::fprintf(stderr, "format: %s\n", brc_last_error(eng)); return -1; }
        }
        t_format += now() - f0;
        { int64_t wc[4]; if (brc_get_warning_counts(eng, wc) == BRC_OK) for (int k = 0; k < 4; ++k) warn_total[k] += wc[k]; }
        warner.replay();
        const double r0 = now();
        const int rr = brc_reset(eng);
        t_reset += now() - r0;
        return rr;
    };
    int64_t pushed = 0;
    RegionFetcher fetcher(bam);
    // big fetches (windows of a cut region) are decoded by several threads, one window ahead of the engine (ParallelFetcher)
    ParallelFetcher pf(bam, bam_path, per_lib, !decode_only);
    for (const auto &kv : rg_lb) pf.rg_lib[kv.first] = lib_rank[kv.second];
    const bool allow_parallel = !is_cram && !device_decode && std::getenv("BRC_CLI_SEQUENTIAL") == nullptr;
    std::unique_ptr<WindowJob[]> jobs(new WindowJob[2]);
    std::future<bool> ahead; size_t ahead_gi = (size_t)-1; int ahead_slot = 0;
    bool par_decode_error = false; uint64_t par_decoded = 0, par_windows = 0;
    auto fetch_range = [&](size_t gi, int64_t &fb, int64_t &fe) {
        const Region &g = regions[gi];
        const int64_t clen = bam.lens[(size_t)g.tid];
        fb = std::max<int64_t>((int64_t)g.beg - 1, 0);
        fe = std::min<int64_t>(g.end, std::max<int64_t>(clen, (int64_t)g.beg + 1));
    };
    auto par_ok = [&](size_t gi) {
        if (!allow_parallel || gi >= regions.size()) return false;
        int64_t fb, fe; fetch_range(gi, fb, fe);
        return ParallelFetcher::eligible(fb, fe);
    };
    auto start_decode = [&](size_t gi, int slot) {
        int64_t fb, fe; fetch_range(gi, fb, fe);
        const int tid = regions[gi].tid; const bool collect = warner.collecting();
        WindowJob *job = &jobs[slot]; ParallelFetcher *pfp = &pf;
        return std::async(std::launch::async, [pfp, job, tid, fb, fe, collect] { return pfp->run(tid, fb, fe, collect, *job); });
    };
    // the decoded window for region gi (prefetched or decoded now); nullptr = take the sequential path
    auto decoded_window = [&](size_t gi) -> WindowJob * {
        int slot = 0; bool ok;
        if (ahead.valid() && ahead_gi == gi) { ok = ahead.get(); slot = ahead_slot; }
        else { if (ahead.valid()) ahead.get(); ok = start_decode(gi, 0).get(); }
        ahead_gi = (size_t)-1;
        WindowJob *job = &jobs[slot];
        if (ok && gi + 1 < regions.size() && par_ok(gi + 1)) { ahead_slot = slot ^ 1; ahead_gi = gi + 1; ahead = start_decode(gi + 1, ahead_slot); }
        if (!ok) return nullptr;
        if (job->cand_overflow && warner.collecting()) return nullptr;      // a BAM full of untagged reads: the per-read warning replay wants them all
        par_decode_error = par_decode_error || job->error; par_decoded += job->n_decoded; ++par_windows;
        return job;
    };
    if (!regions.empty() && par_ok(0)) { ahead_slot = 0; ahead_gi = 0; ahead = start_decode(0, 0); }   // decode under the CUDA start-up
    warm.join();
    rc = decode_only ? BRC_OK : (warm_rc != BRC_OK ? warm_rc : brc_create(&cfg, &eng));
    if (rc != BRC_OK) { std::fprintf(stderr, "brc_create: %s\n", brc_strerror(rc)); return 1; }
    if (eng) brc_set_queue_carry(eng, 1);
    const double t_loop0 = now();
    auto next_fbeg = [&](size_t gi) -> int64_t {   // start of the following fetch when it continues this one, else "keep nothing"
        if (gi + 1 >= regions.size() || regions[gi + 1].tid != regions[gi].tid) return INT64_MAX;
        return std::max<int64_t>((int64_t)regions[gi + 1].beg - 1, 0);
    };
    for (size_t gi = 0; gi < regions.size(); ++gi) {
        const Region &g = regions[gi];
        const double d0 = now();
        if (decode_only) {
            const int64_t fbeg = std::max<int64_t>((int64_t)g.beg - 1, 0), fend = g.end;
            int64_t n = 0, psum = 0, qsum = 0;
            auto count = [&](const Rec &r) { if (r.flag & 4) return; ++n; psum += r.pos; for (int k = 0; k < r.l_qseq; ++k) qsum += r.qual[k]; };
            WindowJob *job = par_ok(gi) ? decoded_window(gi) : nullptr;
            if (job) {
                const brc_read_batch &b = job->batch;
                n = b.n_reads;
                for (int64_t i = 0; i < b.n_reads; ++i) psum += b.pos[i];
                for (uint64_t k = b.qual_off[0]; k < b.qual_off[b.n_reads]; ++k) qsum += b.qual[k];
                fetcher.active = false;
            } else
#ifdef BRC_WITH_HTSLIB
            if (is_cram) hts.fetch(g.tid, fbeg, fend, count); else
#endif
            fetcher.fetch(g.tid, fbeg, fend, next_fbeg(gi), count);
            std::printf("%d\t%d\t%d\t%lld\t%lld\t%lld\n", g.tid, g.beg, g.end, (long long)n, (long long)psum, (long long)qsum);
            continue;
        }
        if (!ref_loaded.count(g.tid)) {   // load_reference: whole chromosome
            if (!fa.fetch(bam.names[(size_t)g.tid], chrom)) { std::fprintf(stderr, "Failed to fetch %s from %s\n", bam.names[(size_t)g.tid].c_str(), fn_fa.c_str()); brc_destroy(eng); return 1; }
            rc = brc_set_reference(eng, g.tid, bam.names[(size_t)g.tid].c_str(), (int64_t)chrom.size(), 0, chrom.data(), (int64_t)chrom.size());
            if (rc != BRC_OK) { std::fprintf(stderr, "brc_set_reference: %s\n", brc_last_error(eng)); brc_destroy(eng); return 1; }
            ref_loaded.insert(g.tid);
            t_ref += now() - d0;
        }
        const double d1 = now();
        if (par_ok(gi)) {
            // one window = one batch: whatever smaller regions are pending goes out first, then the window is pushed as ONE borrowed
            // batch (the engine streams it to the GPU in chunks) while the next window is already being decoded
            if (pushed > 0) { if (flush() != BRC_OK) { brc_destroy(eng); return 1; } pushed = 0; }
            if (WindowJob *job = decoded_window(gi)) {
                brc_begin_region(eng, g.tid, g.beg, g.end, g.site_list ? 1 : 0);
                warner.begin_region(g.tid, g.beg, g.end, g.cont);
                for (Slice &sl : job->slices) warner.take(sl.cands);
                if (job->batch.n_reads > 0) {
                    const int prc = brc_push_reads(eng, &job->batch);
                    if (prc != BRC_OK) { std::fprintf(stderr, "brc_push_reads: %s\n", brc_last_error(eng)); brc_destroy(eng); return 1; }
                }
                brc_end_region(eng);
                fetcher.active = false;
                t_decode += now() - d1;
                if (flush() != BRC_OK) { brc_destroy(eng); return 1; }     // the batch is borrowed until the text is out
                pushed = 0;
                continue;
            }
        }
        brc_begin_region(eng, g.tid, g.beg, g.end, g.site_list ? 1 : 0);
        warner.begin_region(g.tid, g.beg, g.end, g.cont);
        // samfetch(in, idx, ref, d.beg-1, d.end): records with tid, endpos > max(beg-1,0), pos < end, in file order
        const int64_t fbeg = std::max<int64_t>((int64_t)g.beg - 1, 0), fend = g.end;
        int push_rc = BRC_OK;
        if (device_decode && !is_cram) {
            // f-2: hand the engine the compressed span; it inflates, frames and computes on the device.  One span per batch.
            SpanBuilder sb;
            if (sb.build(bam, g.tid, fbeg, fend)) {
                brc_bam_span sp{}; sp.comp = sb.comp.data(); sp.comp_len = (int64_t)sb.comp.size(); sp.n_entry = (int64_t)sb.entries.size(); sp.entry = sb.entries.data();
                sp.end_voff = sb.end_voff; sp.tid = g.tid; sp.n_rg = (int32_t)rg_ids.size(); sp.rg_id = rg_ids.data(); sp.rg_lib = rg_libs.data();
                const int prc = brc_push_bam_span(eng, &sp);
                if (prc != BRC_OK) { std::fprintf(stderr, "brc_push_bam_span: %s\n", brc_last_error(eng)); brc_destroy(eng); return 1; }
            }
            fetcher.active = false;
            brc_end_region(eng);
            t_decode += now() - d1;
            if (flush() != BRC_OK) { brc_destroy(eng); return 1; }
            pushed = 0;
            continue;
        }
        auto push = [&](const Rec &r) {
            uint16_t lib = 0;
            if (per_lib) {
                lib = (uint16_t)BRC_LIB_NONE;
                if (r.has_rg) { auto it = rg_lb.find(r.rg); if (it != rg_lb.end()) lib = lib_rank[it->second]; }
            }
            warner.consider(r, lib == (uint16_t)BRC_LIB_NONE);
            const int prc = brc_push_read(eng, r.tid, r.pos, r.flag, r.mapq, lib, r.l_qseq, r.nm, r.sm, r.n_cigar, r.cigar, r.seq, r.qual);
            if (prc != BRC_OK && push_rc == BRC_OK) push_rc = prc;
            ++pushed;
        };
#ifdef BRC_WITH_HTSLIB
        if (is_cram) hts.fetch(g.tid, fbeg, fend, push); else
#endif
        fetcher.fetch(g.tid, fbeg, fend, next_fbeg(gi), push);
        if (push_rc != BRC_OK) { std::fprintf(stderr, "brc_push_read: %s\n", brc_last_error(eng)); brc_destroy(eng); return 1; }
        brc_end_region(eng);
        t_decode += now() - d1;
        // flush in batches at region boundaries; the deletion queue of the argv loop is carried by the engine (brc_set_queue_carry)
        if (gi + 1 == regions.size() || pushed > 1500000) { if (flush() != BRC_OK) { brc_destroy(eng); return 1; } pushed = 0; }
    }
    warner.finish(warn_total);
#ifdef BRC_WITH_HTSLIB
    const bool decode_error = bam.bz.error || hts.error || par_decode_error;
#else
    const bool decode_error = bam.bz.error || par_decode_error;
#endif
    if (ahead.valid()) ahead.get();
    if (decode_error) std::fprintf(stderr, "[E::bgzf_read] %s: truncated or corrupt BGZF block / BAM record — the output above is incomplete\n", bam_path.c_str());
    if (timing) std::fprintf(stderr, "[brc timing] index seeks %llu  records decoded %llu  (+ %llu records in %llu windows decoded by %d threads)\n", (unsigned long long)fetcher.n_seeks,
                             (unsigned long long)fetcher.n_decoded, (unsigned long long)par_decoded, (unsigned long long)par_windows, pf.n_threads);
    if (decode_only) return decode_error ? 1 : 0;
    if (timing) std::fprintf(stderr, "[brc timing] reference %.3fs  decode+push %.3fs  compute %.3fs  results %.3fs  format %.3fs  reset %.3fs  write %.3fs  | region loop %.3fs, since main() %.3fs\n",
                             t_ref, t_decode, t_compute, t_results, t_format, t_reset, t_write, now() - t_loop0, now() - t_main0);
    if (timing) std::fprintf(stderr, "[brc timing] startup (CUDA context, header, index, first window) %.3fs\n", t_loop0 - t_main0);
    if (!std::getenv("BRC_CLI_CLEAN_EXIT")) {
        // everything is printed: leave without tearing down the CUDA context, the page-locked buffers and the thread pools one by
        // one (0.3-0.9 s on a B200 box, r02t) — the kernel reclaims them
        std::fflush(nullptr);
        _exit(decode_error ? 1 : 0);
    }
    brc_destroy(eng);
    return decode_error ? 1 : 0;
}

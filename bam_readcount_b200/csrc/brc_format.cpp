// brc_format.cpp — host text emitter: turns the engine's binary per-site records into the
// reference's STDOUT lines.
//
//   line assembly + deletion queue   R:src/exe/bam-readcount/bamreadcount.cpp:351-416
//   BasicStat printer                R:src/lib/bamrc/BasicStat.cpp:110-159
//   IndelQueue::process              R:src/lib/bamrc/IndelQueue.cpp:3-15
//
// Only formatting, ordering of allele strings and the p -> p+1 deletion shift live here; every
// number printed was accumulated on the GPU.  Averages are float32 divisions printed with
// "%.2f" of the value promoted to double, exactly like `std::fixed << setprecision(2)`.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "brc_engine_internal.h"

namespace {

struct Stat { uint32_t v[BRC_N_STATS]; };
inline float f32(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

void put_f2(std::string &o, float x) { char t[64]; int n = std::snprintf(t, sizeof t, "%.2f", (double)x); o.append(t, (size_t)n); }
void put_u(std::string &o, uint32_t x) { char t[16]; int n = std::snprintf(t, sizeof t, "%u", x); o.append(t, (size_t)n); }

// operator<<(std::ostream&, const BasicStat&)
void put_stat(std::string &o, const Stat *s, bool is_indel) {
    if (!s || s->v[BRC_S_COUNT] == 0) { o += "0:0.00:0.00:0.00:0:0:0.00:0.00:0.00:0:0.00:0.00:0.00"; return; }
    const uint32_t *v = s->v;
    const float rc = (float)v[BRC_S_COUNT];
    put_u(o, v[BRC_S_COUNT]); o += ':';
    put_f2(o, (float)v[BRC_S_MAPQ] / rc); o += ':';
    if (is_indel) o += "0.00"; else put_f2(o, (float)v[BRC_S_BASEQ] / rc);
    o += ':';
    put_f2(o, (float)v[BRC_S_SE_MAPQ] / rc); o += ':';
    put_u(o, v[BRC_S_PLUS]); o += ':'; put_u(o, v[BRC_S_MINUS]); o += ':';
    put_f2(o, f32(v[BRC_S_POS_FRAC]) / rc); o += ':';
    put_f2(o, f32(v[BRC_S_NM_FRAC]) / rc); o += ':';
    put_f2(o, (float)v[BRC_S_MMQS] / rc); o += ':';
    put_u(o, v[BRC_S_NQ2]); o += ':';
    if (v[BRC_S_NQ2] > 0) put_f2(o, f32(v[BRC_S_Q2_DIST]) / (float)v[BRC_S_NQ2]); else o += "0.00";
    o += ':';
    put_f2(o, (float)v[BRC_S_CLIP_LEN] / rc); o += ':';
    put_f2(o, f32(v[BRC_S_3P_DIST]) / rc);
}

struct QEnt { int32_t tid; int64_t pos; Stat st; std::string allele; };
struct EmitState {
    std::vector<std::deque<QEnt>> q;
    std::vector<char> q_exists;
    explicit EmitState(int rows) : q((size_t)rows), q_exists((size_t)rows, 0) {}
    void clear() { for (auto &d : q) d.clear(); std::fill(q_exists.begin(), q_exists.end(), 0); }
};

const char kNt[] = "=ACGTN";
const uint8_t kCanon[16] = {0, 1, 2, 5, 3, 5, 5, 5, 4, 5, 5, 5, 5, 5, 5, 5};

void format_region(const brc_engine *e, int64_t g, const char *const *lib_names, EmitState &st, std::string &out) {
    const brc_region &rg = e->regions[(size_t)g];
    const brc::HostRef *ref = brc::find_ref(e, rg.tid);
    const int rows = e->n_rows;
    const int64_t NS = e->n_slots, RS = (int64_t)rows * NS;
    const uint32_t *ncover = e->h_ncover.as<uint32_t>(), *npass = e->h_npass.as<uint32_t>(), *pstats = e->h_pstats.as<uint32_t>();
    const uint8_t *flags = e->h_flags.as<uint8_t>(), *pbase = e->h_pbase.as<uint8_t>(), *skind = e->h_sec_kind.as<uint8_t>();
    const int32_t *shead = e->h_sec_head.as<int32_t>(), *snext = e->h_sec_next.as<int32_t>(), *slen = e->h_sec_len.as<int32_t>(),
                  *sqpos = e->h_sec_qpos.as<int32_t>();
    const int64_t *sread = e->h_sec_read.as<int64_t>();
    const uint32_t *sstats = e->h_sec_stats.as<uint32_t>();
    const int64_t SC = e->h_sec_cap;
    const uint8_t *h_seq = e->host_seq(); const uint64_t *h_seq_off = e->host_seq_off();
    std::string rec;
    struct Indel { std::string allele; Stat st; };
    std::vector<Indel> indels;
    for (int32_t s = 0; s < rg.n_slots; ++s) {
        const int64_t slot = rg.slot_base + s;
        const int64_t pos = (int64_t)rg.first_pos + s;
        uint64_t n_total = 0, mapq_n = 0; bool abandoned = false;
        for (int r = 0; r < rows; ++r) { n_total += ncover[r * NS + slot]; mapq_n += npass[r * NS + slot]; abandoned |= (flags[r * NS + slot] & 1) != 0; }
        if (n_total == 0 && !abandoned) continue;       // no read spans the site: the callback never fires
        if (abandoned) continue;                        // -p and a read without library: `return 0` before anything is kept
        rec.clear();
        int64_t extra_depth = 0;
        for (int r = 0; r < rows; ++r) {
            const int64_t idx = r * NS + slot;
            if (ncover[idx] == 0) continue;
            if (e->cfg.per_lib) { rec += '\t'; rec += lib_names ? lib_names[r] : "?"; rec += "\t{"; }
            Stat base[6]; bool have[6] = {false, false, false, false, false, false};
            indels.clear();
            if (pbase[idx] < 6) { for (int k = 0; k < BRC_N_STATS; ++k) base[pbase[idx]].v[k] = pstats[(int64_t)k * RS + idx]; have[pbase[idx]] = true; }
            for (int32_t j = shead[idx]; j >= 0; j = snext[j]) {
                Stat t; for (int k = 0; k < BRC_N_STATS; ++k) t.v[k] = sstats[(int64_t)k * SC + j];
                if (skind[j] < 6) { base[skind[j]] = t; have[skind[j]] = true; continue; }
                Indel in; in.st = t;
                if (skind[j] == BRC_KIND_INS) {          // "+" + canonicalised read bases qpos+1..qpos+len  (R:...:324-330)
                    in.allele = "+";
                    const uint8_t *sq = h_seq + h_seq_off[(size_t)sread[j]];
                    for (int k = 1; k <= slen[j]; ++k) { int i = sqpos[j] + k; uint8_t b = sq[i >> 1]; in.allele += kNt[kCanon[(i & 1) ? (b & 15) : (b >> 4)]]; }
                } else {                                 // "-" + raw reference characters pos+1..pos+len (R:...:331-339)
                    in.allele = "-";
                    for (int k = 1; k <= slen[j]; ++k) {
                        int64_t p = pos + k; char c = 'N';
                        if (ref && p >= ref->win_beg && p < ref->win_beg + ref->win_len && p < ref->chrom_len) c = ref->seq[(size_t)(p - ref->win_beg)];
                        in.allele += c;
                    }
                }
                indels.push_back(std::move(in));
            }
            for (int j = 0; j < 6; ++j) { rec += '\t'; rec += kNt[j]; rec += ':'; put_stat(rec, have[j] ? &base[j] : nullptr, false); }
            std::sort(indels.begin(), indels.end(), [](const Indel &a, const Indel &b) { return a.allele < b.allele; });
            for (auto &in : indels) {
                if (in.allele[0] == '-') { st.q[(size_t)r].push_back(QEnt{rg.tid, pos + 1, in.st, in.allele}); st.q_exists[(size_t)r] = 1; }
                else { rec += '\t'; rec += in.allele; rec += ':'; put_stat(rec, &in.st, true); }
            }
            if (st.q_exists[(size_t)r]) {               // IndelQueue::process(tid, pos, record)
                auto &q = st.q[(size_t)r];
                while (!q.empty() && ((q.front().tid == rg.tid && q.front().pos < pos) || q.front().tid != rg.tid)) q.pop_front();
                while (!q.empty() && q.front().tid == rg.tid && q.front().pos == pos) {
                    rec += '\t'; rec += q.front().allele; rec += ':'; put_stat(rec, &q.front().st, true);
                    extra_depth += q.front().st.v[BRC_S_COUNT];
                    q.pop_front();
                }
            }
            if (e->cfg.per_lib) rec += "\t}";
        }
        if (pos >= rg.beg && pos < rg.end) {
            char rb = 'N';
            if (ref && pos < ref->chrom_len && pos >= ref->win_beg && pos < ref->win_beg + ref->win_len) rb = ref->seq[(size_t)(pos - ref->win_beg)];
            out += ref ? ref->name : std::string("?"); out += '\t';
            char t[32]; int n = std::snprintf(t, sizeof t, "%lld", (long long)(pos + 1)); out.append(t, (size_t)n);
            out += '\t'; out += rb; out += '\t';
            n = std::snprintf(t, sizeof t, "%lld", (long long)((int64_t)mapq_n + extra_depth)); out.append(t, (size_t)n);
            out += rec; out += '\n';
        }
    }
    if (rg.site_list_mode) st.clear();                  // d.indel_queue_map.clear()  (R:...:605)
}

}  // namespace

extern "C" int64_t brc_format_text(brc_engine *e, int64_t region_index, const char *const *lib_names, char *buf, int64_t cap) {
    if (!e) return BRC_E_INVALID;
    if (!e->results_valid) return brc::set_error(e, BRC_E_INVALID, "format_text: no results");
    if (e->n_host_reads() == 0 && e->h_n_sec > 0) return brc::set_error(e, BRC_E_INVALID, "format_text: needs the pushed reads (push path only)");
    if (region_index >= (int64_t)e->regions.size()) return BRC_E_INVALID;
    std::string out;
    EmitState st(e->n_rows);
    if (region_index < 0) for (int64_t g = 0; g < (int64_t)e->regions.size(); ++g) format_region(e, g, lib_names, st, out);
    else format_region(e, region_index, lib_names, st, out);
    if (buf && cap > 0) {
        int64_t n = std::min<int64_t>((int64_t)out.size(), cap - 1);
        std::memcpy(buf, out.data(), (size_t)n); buf[n] = 0;
    }
    return (int64_t)out.size();
}

#!/usr/bin/env python
"""TEST INFRASTRUCTURE — applies the INTEGRATION.md §2 binding to a COPY of the reference's main file.

Reads <reference>/src/exe/bam-readcount/bamreadcount.cpp, writes the patched translation unit to <out>; the copy lives only
under the git-ignored oracle/_ref/work (build_patched_ref.sh deletes it after compiling).  Nothing of the reference is stored
here: the script holds a handful of one-line anchors to find the call sites and the glue code that replaces them — the code a
bam-readcount maintainer would add to route fetch_func / pileup_func / the region drivers through libbrc_engine.so:

  fetch_func(b)                          -> brc_push_read(...)                 (R:bamreadcount.cpp:114-261)
  bam_plbuf_init + bam_plp_set_maxcnt    -> brc_set_reference + brc_begin_region   (R:...:591-592, 650-651)
  bam_plbuf_push(0) + bam_plbuf_destroy  -> brc_end_region [+ compute + emit]      (R:...:603-604, 655-656)
  every pileup_func() + cout             -> brc_compute + brc_format_text
"""
import sys

GLUE = r'''
// ---- libbrc_engine.so binding (INTEGRATION.md section 2) ----
#include "brc_engine.h"
#include <vector>
struct brc_glue_t {
    brc_engine *eng; bam_header_t *header; std::map<std::string, uint16_t> lib_rank; std::vector<const char *> lib_names;
    bool per_lib; int cur_tid; bool pending;
} g_brc = {0, 0, std::map<std::string, uint16_t>(), std::vector<const char *>(), false, -1, false};

static inline int32_t brc_aux_int(const bam1_t *b, const char tag[2]) {
    uint8_t *p = bam_aux_get(b, tag);                       // the lookup process_read does
    return p ? (int32_t)bam_aux2i(p) : BRC_TAG_ABSENT;
}
static int brc_fetch(const bam1_t *b) {                      // fetch_func + bam_plbuf_push(b)
    uint16_t lib = 0;
    if (g_brc.per_lib) {
        const char *lb = bam_get_library(g_brc.header, b);
        lib = lb ? g_brc.lib_rank[lb] : (uint16_t)BRC_LIB_NONE;
    }
    return brc_push_read(g_brc.eng, b->core.tid, b->core.pos, b->core.flag, b->core.qual, lib, b->core.l_qseq,
                         brc_aux_int(b, "NM"), brc_aux_int(b, "SM"), b->core.n_cigar, bam1_cigar(b), bam1_seq(b), bam1_qual(b));
}
static void brc_die(const char *what) { fprintf(stderr, "brc engine: %s: %s\n", what, brc_last_error(g_brc.eng)); exit(1); }
template <class D> static void brc_glue_init(D *d, const std::set<std::string> &names) {
    brc_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.min_mapq = d->min_mapq; cfg.min_bq = d->min_bq; cfg.max_cnt = d->max_cnt; cfg.per_lib = d->per_lib; cfg.insertion_centric = d->insertion_centric;
    uint16_t r = 0;
    for (std::set<std::string>::const_iterator it = names.begin(); it != names.end(); ++it) { g_brc.lib_rank[*it] = r++; g_brc.lib_names.push_back(it->c_str()); }
    cfg.n_libs = (int32_t)names.size(); cfg.device = 0;
    g_brc.header = d->in->header; g_brc.per_lib = d->per_lib;
    if (brc_create(&cfg, &g_brc.eng) != BRC_OK) { fprintf(stderr, "brc engine: no usable CUDA device\n"); exit(1); }
}
template <class D> static void brc_region_begin(D *d, int ref, int site_list_mode) {
    if (d->ref && ref != g_brc.cur_tid) {                    // after load_reference(): hand the contig to the engine once
        if (brc_set_reference(g_brc.eng, ref, d->in->header->target_name[ref], d->len, 0, d->ref, d->len) != BRC_OK) brc_die("set_reference");
        g_brc.cur_tid = ref;
    }
    if (brc_begin_region(g_brc.eng, ref, d->beg, d->end, site_list_mode) != BRC_OK) brc_die("begin_region");
}
static void brc_emit() {                                     // every pileup_func() invocation of the pushed regions + their cout lines
    if (!g_brc.pending) return;
    if (brc_compute(g_brc.eng) != BRC_OK) brc_die("compute");
    const char *const *names = g_brc.lib_names.empty() ? 0 : &g_brc.lib_names[0];
    int64_t need = brc_format_text(g_brc.eng, -1, names, 0, 0);
    if (need < 0) brc_die("format_text");
    std::vector<char> out((size_t)need + 1);
    brc_format_text(g_brc.eng, -1, names, &out[0], need + 1);
    fwrite(&out[0], 1, (size_t)need, stdout);
    brc_reset(g_brc.eng);
    g_brc.pending = false;
}
static void brc_region_end(int site_list_mode) {
    if (brc_end_region(g_brc.eng) != BRC_OK) brc_die("end_region");
    g_brc.pending = true;
    if (site_list_mode) brc_emit();                          // the -l loop clears its queues per line: emit per line; argv regions share one queue: emit once
}
// ---- end of binding ----
'''


def patch(src: str) -> str:
    def once(s, old, new, nth=0, count=1):
        idx = -1
        for _ in range(nth + 1):
            idx = s.index(old, idx + 1)
        return s[:idx] + new + s[idx + len(old):]

    # 1. glue after the WARN global (all htslib / std headers are in scope there)
    anchor = "std::auto_ptr<ReadWarnings> WARN;"
    assert anchor in src
    src = src.replace(anchor, anchor + "\n" + GLUE, 1)
    # 2. fetch_func: route the record to the engine
    a = "static int fetch_func(const bam1_t *b, void *data) {"
    assert a in src
    src = src.replace(a, a + "\n    if (g_brc.eng) return brc_fetch(b);", 1)
    # 3. the two region drivers (site list first, argv regions second)
    init = "bam_plbuf_t *buf = bam_plbuf_init(pileup_func, &d); // initialize pileup"
    setm = "bam_plp_set_maxcnt(buf->iter, d.max_cnt);"
    fin = "bam_plbuf_push(0, buf); // finalize pileup"
    des = "bam_plbuf_destroy(buf);"
    assert src.count(init) == 2 and src.count(setm) == 2 and src.count(fin) == 2 and src.count(des) == 2
    for mode in (1, 0):
        src = once(src, init, f"brc_region_begin(&d, ref, {mode}); bam_plbuf_t *buf = 0;")
        src = once(src, setm, "")
        src = once(src, fin, f"brc_region_end({mode});")
        src = once(src, des, "")
    # 4. argv regions: one batch for the whole loop (the deletion queue is never cleared between them), emitted after it
    tail = "hts_idx_destroy(idx);"
    assert src.count(tail) == 2
    src = once(src, tail, "brc_emit(); hts_idx_destroy(idx);", nth=1)
    # 5. engine creation once the options and the header are known
    a = "d.indel_queue_map = indel_queue_map_t();"
    assert a in src
    src = src.replace(a, a + "\n    if (vm.count(\"region\") || !fn_pos.empty()) brc_glue_init(&d, lib_names);", 1)
    return src


if __name__ == "__main__":
    ref_root, out = sys.argv[1], sys.argv[2]
    text = open(f"{ref_root}/src/exe/bam-readcount/bamreadcount.cpp").read()
    open(out, "w").write(patch(text))

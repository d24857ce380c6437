// brc_kernels.cu — hand-written sm_100a kernels of the pileup-readcount hot path.
//
//   K0  read_precompute_kernel  ≙ fetch_func           R:src/exe/bam-readcount/bamreadcount.cpp:114-253
//                                 + bam_plp_push admit  V:htslib-1.10/sam.c:4484-4531 (FUNMAP / tid<0)
//                                 + bam_endpos          V:htslib-1.10/sam.c:507-513
//   K1  pileup_kernel           ≙ bam_plp64_next        V:htslib-1.10/sam.c:4416-4466   (which reads span a site)
//                                 + resolve_cigar2      V:htslib-1.10/sam.c:3964-4041   (qpos / is_del / indel)
//                                 + pileup_func         R:...bamreadcount.cpp:265-346   (filters, classification)
//                                 + BasicStat::process_read  R:src/lib/bamrc/BasicStat.cpp:28-107
//
// Formulation: site-centric gather.  One thread owns one (site, library-row);
// it walks the reads overlapping its warp's 32 sites IN FILE ORDER and accumulates the 13
// statistics of the site's primary allele in registers.  File order per key is exactly the
// reference's accumulation order, so the four float32 sums (and the one double-rounded add)
// are bit-identical to the CPU reference at any depth (SURVEY.md §7 hard part 1) — no
// event tuples are ever written to HBM.  Rare keys (a second base class at a site, indel
// alleles) go to an L2-resident record pool owned by the same thread.
//
// All float arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction, no
// fast-math): the results must match the reference's x86-64 SSE arithmetic bit for bit.
#include "brc_device.cuh"

namespace brc {

// seq_nt16_table (V:htslib-1.10/hts.c:73-91): ASCII -> 4-bit IUPAC code, 15 for anything else
__constant__ uint8_t c_nt16[256];
// bam_nt16_canonical_table (R:bamreadcount.cpp:36-39): nibble -> index into "=ACGTN"
__device__ __forceinline__ uint32_t canonical16(uint32_t nib) {
    // packed 16 x 4-bit LUT: {0,1,2,5,3,5,5,5,4,5,5,5,5,5,5,5}
    return (0x5555555455535210ull >> (nib * 4)) & 0xFu;
}

static uint8_t h_nt16[256];
static bool h_nt16_ready = false;
static void build_nt16() {
    for (int i = 0; i < 256; ++i) h_nt16[i] = 15;
    const char *s = "=ACMGRSVTWYHKDBN";
    for (int i = 0; i < 16; ++i) {
        h_nt16[(unsigned char)s[i]] = (uint8_t)i;
        if (s[i] >= 'A' && s[i] <= 'Z') h_nt16[(unsigned char)(s[i] + 32)] = (uint8_t)i;
    }
    h_nt16['0'] = 1; h_nt16['1'] = 2; h_nt16['2'] = 4; h_nt16['3'] = 8;
    h_nt16_ready = true;
}

__device__ __forceinline__ bool is_refop(uint32_t op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
__device__ __forceinline__ bool is_matchop(uint32_t op) { return op == 0 || op == 7 || op == 8; }
__device__ __forceinline__ uint32_t seq_nib(const uint8_t *seq, uint64_t off, int i) {
    uint32_t b = seq[off + (uint32_t)(i >> 1)];
    return (i & 1) ? (b & 0xFu) : (b >> 4);
}

// ---------------------------------------------------------------------------------------------
// init: tile read ranges + counters
// ---------------------------------------------------------------------------------------------
__global__ void init_tiles_kernel(int32_t *tile_lo, int32_t *tile_hi, int64_t n_tiles, int32_t *sec_count,
                                  unsigned long long *warn) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_tiles) { tile_lo[i] = 0x7fffffff; tile_hi[i] = 0; }
    if (i == 0) { *sec_count = 0; warn[0] = 0; warn[1] = 0; }
}

cudaError_t launch_init_tiles(int32_t *tile_lo, int32_t *tile_hi, int64_t n_tiles, int32_t *sec_count,
                              unsigned long long *warn, cudaStream_t s) {
    int64_t n = n_tiles > 1 ? n_tiles : 1;
    init_tiles_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tile_lo, tile_hi, n_tiles, sec_count, warn);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K0: per-read precompute.  One thread per read (v0).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ char ref_at(const RefWin &rw, int64_t p) {
    if (p < 0 || p >= rw.chrom_len) return 0;          // the reference's string ends with NUL at chrom_len
    if (p < rw.win_beg || p >= rw.win_beg + rw.win_len) return 'N';
    return rw.seq[p - rw.win_beg];
}

__global__ void __launch_bounds__(128) read_precompute_kernel(PrecomputeParams P) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P.reads.n_reads) return;
    const ReadsDev &R = P.reads;
    int32_t g = P.region_of_read ? P.region_of_read[i] : 0;
    const RegionDev rg = P.regions[g];
    const RefWin rw = P.refs[rg.tid_slot];

    const int32_t pos = R.pos[i];
    const uint32_t flag = R.flag[i];
    const uint32_t mapq = R.mapq[i];
    const int32_t l_qseq = R.l_qseq[i];
    const uint64_t coff = R.cigar_off[i];
    const uint32_t n_cigar = (uint32_t)(R.cigar_off[i + 1] - coff);
    const uint32_t *cig = R.cigar + coff;
    const uint64_t soff = R.seq_off[i], qoffb = R.qual_off[i];
    const uint8_t *qual = R.qual + qoffb;

    // --- fetch_func CIGAR/reference walk (R:...:133-199) + bam_cigar2rlen + SIMPLE detection ---
    uint32_t sum_mmq = 0;
    int left_clip = 0, clipped_length = l_qseq, right_clip = l_qseq;
    int last_mm_pos = -1, last_mm_qual = 0;
    int64_t reference_position = pos;
    int read_position = 0;
    bool walking = true;            // false after the reference's `break` out of the op loop
    int64_t rlen = 0;               // reference span (all ref-consuming ops, incl. = and X)
    int n_refops = 0, qoff = 0;
    bool simple = true, seen_ref = false;
    for (uint32_t k = 0; k < n_cigar; ++k) {
        const uint32_t c = cig[k];
        const int op_length = (int)(c >> 4);
        const uint32_t op = c & 0xFu;
        if (is_refop(op)) { rlen += op_length; n_refops++; seen_ref = true; if (!is_matchop(op)) simple = false; }
        else if (op == 1 || op == 6) simple = false;
        else if (op == 4 && !seen_ref) qoff += op_length;
        if (!walking) continue;
        if (op == 0) {
            int j;
            for (j = 0; j < op_length; ++j) {
                const int cur = read_position + j;
                const int64_t refpos = reference_position + j;
                if (rg.ref_len_check && refpos > rw.chrom_len) continue;
                const char rc = ref_at(rw, refpos);
                if (rc == 0) break;
                const uint32_t ref_base = c_nt16[(unsigned char)rc];
                const uint32_t read_base = seq_nib(R.seq, soff, cur);
                if (read_base != ref_base && ref_base != 15 && read_base != 0) {
                    const int q = qual[cur];
                    if (last_mm_pos != -1) {
                        if (last_mm_pos + 1 != cur) { sum_mmq += (uint32_t)last_mm_qual; last_mm_qual = q; }
                        else if (last_mm_qual < q) last_mm_qual = q;
                        last_mm_pos = cur;
                    } else { last_mm_pos = cur; last_mm_qual = q; }
                }
            }
            if (j < op_length) { walking = false; continue; }
            reference_position += op_length; read_position += op_length;
        } else if (op == 2 || op == 3) reference_position += op_length;
        else if (op == 1) read_position += op_length;
        else if (op == 4) {
            read_position += op_length; clipped_length -= op_length;
            if (k == 0) left_clip += op_length; else right_clip -= op_length;
        }
    }
    sum_mmq += (uint32_t)last_mm_qual;
    if (n_refops != 1) simple = false;

    // --- Q2 run / effective 3' end (R:...:202-238) ---
    int tpi, q2_pos = -1, kk, inc;
    const bool reverse = (flag & 16u) != 0;
    if (reverse) { kk = tpi = 0; inc = 1; if (tpi < left_clip) tpi = left_clip; }
    else { kk = tpi = l_qseq - 1; inc = -1; if (tpi > right_clip) tpi = right_clip; }
    while (kk >= 0 && kk < l_qseq) {
        if (qual[kk] != 2) { q2_pos = kk - 1; break; }
        kk += inc;
    }
    if (reverse) { if (tpi < q2_pos) tpi = q2_pos; }
    else { if (tpi > q2_pos && q2_pos != -1) tpi = q2_pos; }

    // --- admission (bam_plp_push) and span (bam_endpos) ---
    const int32_t tid_ok = 1;  // tid<0 reads are dropped by the host batcher / caller contract
    const bool unmapped = (flag & 4u) != 0;
    int64_t end = (!unmapped && n_cigar > 0) ? (int64_t)pos + rlen : (int64_t)pos + 1;
    if (unmapped || !tid_ok) end = pos;      // never admitted to the pileup: covers nothing

    ReadDesc d;
    d.pos = pos; d.end = (int32_t)end; d.l_qseq = l_qseq;
    uint32_t fm = (flag & 0xFFFFu) | (mapq << 16);
    if (simple) fm |= FM_SIMPLE;
    const int32_t nm = R.nm[i], sm = R.sm[i];
    if (nm == INT32_MIN) fm |= FM_NM_ABSENT;
    int32_t se;
    if (flag & 2u) { if (sm != INT32_MIN) se = sm; else { se = 0; fm |= FM_SM_MISSING; } }
    else se = (int32_t)mapq;
    d.fm = fm;
    d.mmq = (int32_t)sum_mmq; d.clen = clipped_length; d.lclip = left_clip; d.tpi = tpi;
    d.q2 = q2_pos;
    d.nmfrac = (nm == INT32_MIN) ? 0.0f : __fdiv_rn((float)nm, (float)clipped_length);
    d.se = se;
    const uint32_t lib = R.lib ? (uint32_t)R.lib[i] : 0u;
    d.lib_nc = lib | ((n_cigar > 0xFFFFu ? 0xFFFFu : n_cigar) << 16);
    d.seq_off = soff; d.qual_off = qoffb;
    d.qoff = qoff; d.cigar_off = (uint32_t)coff; d.n_cigar = n_cigar; d.pad = 0;
    // 5 x 16-byte stores
    int4 *dst = reinterpret_cast<int4 *>(P.desc + i);
    const int4 *src = reinterpret_cast<const int4 *>(&d);
#pragma unroll
    for (int q = 0; q < 5; ++q) dst[q] = src[q];

    // --- which tiles of this read's region does it overlap?  (first/last read per tile) ---
    int64_t a = pos > rg.first_pos ? pos : rg.first_pos;
    int64_t b = end < rg.end ? end : rg.end;
    if (b > a) {
        int64_t t0 = rg.tile_base + (a - rg.first_pos) / TILE;
        int64_t t1 = rg.tile_base + (b - 1 - rg.first_pos) / TILE;
        const int32_t idx = (int32_t)i;
        for (int64_t t = t0; t <= t1; ++t) { atomicMin(P.tile_lo + t, idx); atomicMax(P.tile_hi + t, idx + 1); }
    }
}

cudaError_t launch_precompute(const PrecomputeParams &p, cudaStream_t s) {
    if (!h_nt16_ready) build_nt16();
    static bool uploaded[64] = {false};
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 64 && !uploaded[dev]) {
        cudaError_t e = cudaMemcpyToSymbol(c_nt16, h_nt16, 256);
        if (e != cudaSuccess) return e;
        uploaded[dev] = true;
    }
    if (p.reads.n_reads == 0) return cudaSuccess;
    const int bs = 128;
    read_precompute_kernel<<<(unsigned)((p.reads.n_reads + bs - 1) / bs), bs, 0, s>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K1: site-centric pileup + ordered accumulation
// ---------------------------------------------------------------------------------------------
struct Acc {  // the 13 accumulators, print order (BRC_S_*)
    uint32_t count, mapq, baseq, se, plus, minus;
    float posf, nmf;
    uint32_t mmqs, nq2;
    float q2d;
    uint32_t clip;
    float d3p;
};
struct Event {  // one (site, read) event's contributions (computed once, used for base and indel keys)
    uint32_t mapq, baseq, mmq, clen;
    int32_t se;
    bool minus, has_q2;
    float q2term, d3pterm, nmterm;
    double posterm;   // 1.0 - |(qpos-left_clip) - rc| / rc   evaluated in double (R:BasicStat.cpp:70)
};

__device__ __forceinline__ void acc_zero(Acc &a) {
    a.count = a.mapq = a.baseq = a.se = a.plus = a.minus = a.mmqs = a.nq2 = a.clip = 0;
    a.posf = a.nmf = a.q2d = a.d3p = 0.0f;
}
__device__ __forceinline__ void acc_add(Acc &a, const Event &e, bool is_indel) {
    a.count++;
    a.mapq += e.mapq;
    if (e.minus) a.minus++; else a.plus++;
    a.mmqs += e.mmq;
    if (e.has_q2) { a.q2d = __fadd_rn(a.q2d, e.q2term); a.nq2++; }
    a.d3p = __fadd_rn(a.d3p, e.d3pterm);
    a.clip += e.clen;
    a.posf = __double2float_rn(__dadd_rn((double)a.posf, e.posterm));
    a.se += (uint32_t)e.se;
    a.nmf = __fadd_rn(a.nmf, e.nmterm);
    if (!is_indel) a.baseq += e.baseq;
}

// secondary key record: find-or-append in the thread's private chain, then accumulate in place
__device__ __noinline__ void sec_accumulate(const PileupParams &P, int32_t &head, int kind, int len, int64_t read,
                                            int qpos, const Event &e, bool is_indel) {
    const ResultsDev &S = P.res;
    int32_t j = head;
    while (j >= 0) {
        if (S.sec_kind[j] == (uint8_t)kind && S.sec_len[j] == len) {
            if (kind != KIND_INS) break;
            // same inserted bases?  compare canonicalised read bases (R:bamreadcount.cpp:324-330)
            const ReadDesc *da = P.desc + read, *db = P.desc + S.sec_read[j];
            const uint64_t oa = da->seq_off, ob = db->seq_off;
            const int qb = S.sec_qpos[j];
            bool same = true;
            for (int k = 1; k <= len && same; ++k)
                same = canonical16(seq_nib(P.seq, oa, qpos + k)) == canonical16(seq_nib(P.seq, ob, qb + k));
            if (same) break;
        }
        j = S.sec_next[j];
    }
    if (j < 0) {
        j = atomicAdd(S.sec_count, 1);
        if ((int64_t)j >= S.sec_cap) return;   // overflow: host sees sec_count > cap and retries with a larger pool
        S.sec_next[j] = head; S.sec_kind[j] = (uint8_t)kind; S.sec_len[j] = len; S.sec_read[j] = read; S.sec_qpos[j] = qpos;
#pragma unroll
        for (int k = 0; k < N_STATS; ++k) S.sec_stats[(int64_t)k * S.sec_cap + j] = 0u;
        head = j;
    }
    uint32_t *st = S.sec_stats + j;
    const int64_t c = S.sec_cap;
    st[0 * c] += 1u;
    st[1 * c] += e.mapq;
    if (!is_indel) st[2 * c] += e.baseq;
    st[3 * c] += (uint32_t)e.se;
    if (e.minus) st[5 * c] += 1u; else st[4 * c] += 1u;
    st[6 * c] = __float_as_uint(__double2float_rn(__dadd_rn((double)__uint_as_float(st[6 * c]), e.posterm)));
    st[7 * c] = __float_as_uint(__fadd_rn(__uint_as_float(st[7 * c]), e.nmterm));
    st[8 * c] += e.mmq;
    if (e.has_q2) { st[9 * c] += 1u; st[10 * c] = __float_as_uint(__fadd_rn(__uint_as_float(st[10 * c]), e.q2term)); }
    st[11 * c] += e.clen;
    st[12 * c] = __float_as_uint(__fadd_rn(__uint_as_float(st[12 * c]), e.d3pterm));
}

// stateless resolve_cigar2: which op holds `site`, qpos, is_del, indel
__device__ __noinline__ void resolve_general(const uint32_t *cig, uint32_t n_cigar, int32_t pos, int32_t site, int &qpos,
                                             int &indel, bool &is_del) {
    int64_t x = pos; int y = 0; uint32_t k = 0; uint32_t op = 0; int len = 0;
    for (; k < n_cigar; ++k) {
        const uint32_t c = cig[k]; op = c & 0xFu; len = (int)(c >> 4);
        if (is_refop(op)) {
            if ((int64_t)site < x + len) break;
            x += len; if (is_matchop(op)) y += len;
        } else if (op == 1 || op == 4) y += len;
    }
    indel = 0; is_del = false; qpos = 0;
    if (k >= n_cigar) { is_del = true; return; }  // cannot happen for pos <= site < end
    if (is_matchop(op)) qpos = y + (int)(site - x); else { is_del = true; qpos = y; }
    if (x + len - 1 == site && k + 1 < n_cigar) {
        const uint32_t c2 = cig[k + 1]; const uint32_t op2 = c2 & 0xFu; const int l2 = (int)(c2 >> 4);
        if (op2 == 2) indel = -l2;
        else if (op2 == 1) indel = l2;
        else if (op2 == 6 && k + 2 < n_cigar) {
            int l3 = 0;
            for (uint32_t m = k + 2; m < n_cigar; ++m) {
                const uint32_t c3 = cig[m]; const uint32_t op3 = c3 & 0xFu;
                if (op3 == 1) l3 += (int)(c3 >> 4);
                else if (op3 == 2 || op3 == 0 || op3 == 3 || op3 == 7 || op3 == 8) break;
            }
            if (l3 > 0) indel = l3;
        }
    }
}

template <bool PER_LIB>
__global__ void __launch_bounds__(TILE) pileup_kernel(PileupParams P) {
    const int64_t tile = blockIdx.x;
    const uint32_t row = blockIdx.y;
    const TileInfo ti = P.tiles[tile];
    const int sl = threadIdx.x;
    const bool active = sl < ti.n;
    const int32_t site = ti.pos0 + sl;
    const int32_t lo = P.tile_lo[tile], hi = P.tile_hi[tile];
    const int warp0 = sl & ~31;
    const int32_t wfirst = ti.pos0 + warp0;
    const int32_t wlast = ti.pos0 + min(warp0 + 31, ti.n - 1);
    if (warp0 >= ti.n) return;   // whole warp beyond the tile

    Acc acc; acc_zero(acc);
    uint32_t ncover = 0, npass = 0, flags = 0, pbase = NO_BASE;
    int32_t sec_head = -1;
    uint32_t warn_sm = 0, warn_nm = 0;
    const int4 *desc4 = reinterpret_cast<const int4 *>(P.desc);

    for (int32_t r = lo; r < hi; ++r) {
        const int4 q0 = __ldg(desc4 + (int64_t)r * 5 + 0);       // pos,end,l_qseq,fm
        if (q0.x > wlast) break;                                 // reads are position-sorted within a region
        if (q0.y <= wfirst) continue;
        const int4 q2 = __ldg(desc4 + (int64_t)r * 5 + 2);       // q2,nmfrac,se,lib_nc
        const bool cover = active && site >= q0.x && site < q0.y;
        if (PER_LIB) {
            const uint32_t lib = (uint32_t)q2.w & 0xFFFFu;
            if (lib == LIB_NONE) { if (cover) flags |= 1u; continue; }
            if (lib != row) continue;
            // -p: pileup_func returns at the first read without a library (R:...:281-284); nothing after
            // it in pileup (= file) order is processed or warned about at this site
            if (flags & 1u) continue;
        }
        if (!cover) continue;
        ncover++;
        const uint32_t fm = (uint32_t)q0.w;
        const int4 q3 = __ldg(desc4 + (int64_t)r * 5 + 3);       // seq_off, qual_off
        const int4 q4 = __ldg(desc4 + (int64_t)r * 5 + 4);       // qoff,cigar_off,n_cigar
        const uint64_t seq_off = ((uint64_t)(uint32_t)q3.y << 32) | (uint32_t)q3.x;
        const uint64_t qual_off = ((uint64_t)(uint32_t)q3.w << 32) | (uint32_t)q3.z;
        int qpos, indel = 0; bool is_del = false;
        if (fm & FM_SIMPLE) qpos = site - q0.x + q4.x;
        else resolve_general(P.cigar + (uint32_t)q4.y, (uint32_t)q4.z, q0.x, site, qpos, indel, is_del);
        if (is_del) continue;
        const uint32_t mapq = (fm >> 16) & 0xFFu;
        if ((int)mapq < P.min_mapq) continue;
        const uint32_t bq = P.qual[qual_off + (uint32_t)qpos];
        if ((int)bq < P.min_bq) continue;
        if (fm & FLAG_FILTER) continue;
        npass++;

        const int4 q1 = __ldg(desc4 + (int64_t)r * 5 + 1);       // mmq,clen,lclip,tpi
        Event e;
        e.mapq = mapq; e.baseq = bq; e.mmq = (uint32_t)q1.x; e.clen = (uint32_t)q1.y; e.se = q2.z;
        e.minus = (fm & 16u) != 0;
        const float fl = (float)q0.z;
        e.has_q2 = q2.x > -1;
        e.q2term = e.has_q2 ? __fdiv_rn((float)abs(qpos - q2.x), fl) : 0.0f;
        e.d3pterm = __fdiv_rn((float)abs(qpos - q1.w), fl);
        e.nmterm = __int_as_float(q2.y);
        const float rc = __fmul_rn((float)q1.y, 0.5f);
        const float f = __fdiv_rn(fabsf(__fsub_rn((float)(qpos - q1.z), rc)), rc);
        e.posterm = __dsub_rn(1.0, (double)f);
        const bool nm_absent = (fm & FM_NM_ABSENT) != 0, sm_missing = (fm & FM_SM_MISSING) != 0;
        if (nm_absent) e.nmterm = 0.0f;

        if (indel != 0) {
            sec_accumulate(P, sec_head, indel > 0 ? KIND_INS : KIND_DEL, indel > 0 ? indel : -indel, (int64_t)r, qpos, e, true);
            warn_nm += nm_absent; warn_sm += sm_missing;
        }
        if (indel < 1 || !P.insertion_centric) {
            const uint32_t base = canonical16(seq_nib(P.seq, seq_off, qpos));
            if (pbase == NO_BASE) pbase = base;
            if (base == pbase) {
                acc_add(acc, e, false);   // NM missing: nmterm is +0.0f, which leaves the float sum unchanged
            } else sec_accumulate(P, sec_head, (int)base, 0, (int64_t)r, qpos, e, false);
            warn_nm += nm_absent; warn_sm += sm_missing;
        }
    }

    if (active) {
        const ResultsDev &S = P.res;
        const int64_t idx = (int64_t)row * S.n_slots + ti.slot0 + sl;
        const int64_t stride = (int64_t)S.n_rows * S.n_slots;
        S.ncover[idx] = ncover; S.npass[idx] = npass; S.flags[idx] = (uint8_t)flags; S.pbase[idx] = (uint8_t)pbase;
        S.sec_head[idx] = sec_head;
        uint32_t *ps = S.pstats + idx;
        ps[0 * stride] = acc.count; ps[1 * stride] = acc.mapq; ps[2 * stride] = acc.baseq; ps[3 * stride] = acc.se;
        ps[4 * stride] = acc.plus; ps[5 * stride] = acc.minus; ps[6 * stride] = __float_as_uint(acc.posf);
        ps[7 * stride] = __float_as_uint(acc.nmf); ps[8 * stride] = acc.mmqs; ps[9 * stride] = acc.nq2;
        ps[10 * stride] = __float_as_uint(acc.q2d); ps[11 * stride] = acc.clip; ps[12 * stride] = __float_as_uint(acc.d3p);
    }
    // warning counters: warp-reduce then one atomic per warp
    for (int o = 16; o; o >>= 1) { warn_sm += __shfl_xor_sync(0xffffffffu, warn_sm, o); warn_nm += __shfl_xor_sync(0xffffffffu, warn_nm, o); }
    if ((sl & 31) == 0) {
        if (warn_sm) atomicAdd(P.res.warn + 0, (unsigned long long)warn_sm);
        if (warn_nm) atomicAdd(P.res.warn + 1, (unsigned long long)warn_nm);
    }
}

cudaError_t launch_pileup(const PileupParams &p, cudaStream_t s) {
    if (p.n_tiles == 0) return cudaSuccess;
    // grid.x is limited to 2^31-1 tiles; grid.y = library rows
    dim3 grid((unsigned)p.n_tiles, (unsigned)p.res.n_rows, 1);
    if (p.per_lib) pileup_kernel<true><<<grid, TILE, 0, s>>>(p);
    else pileup_kernel<false><<<grid, TILE, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace brc

// fcx_dazz.cu -- reading a Dazzler read database (.db / .idx / .bps) and a local-alignment file (.las)
// directly, in place of the text hop
//        LA4Falcon -H$CUTOFF -fo db las | python -m falcon_kit.mains.consensus ...
// (falcon_kit/mains/consensus_task.py:81-90, falcon_kit/bash.py:349-358).
//
// What replaces what
//   * the read store: every (trimmed) read of the DB is uploaded ONCE, 2-bit packed, in both
//     orientations (pool entry 2r = read r, 2r + 1 = its reverse complement) -- the .bps bytes go to the
//     device as they are and k_repack_bps turns them into the engine's layout (no ASCII at any point);
//     LA4Falcon re-ships the whole B read as text for every overlap;
//   * the block lists: the overlap records are walked in file order with LA4Falcon's -f -o -H rules
//     and the consensus parser's rules (falcon_kit/mains/consensus.py:161-209, :26-45) applied to
//     (read id, length) instead of text lines; a block is a list of pool ids.
//
// FORMAT PROVENANCE (parity unpinned).  Neither the Dazzler tools nor any .db/.las sample are in the
// reference tree (LA4Falcon is an un-vendored third-party dependency).  The layouts
// below restate the public headers of DAZZ_DB (DB.h: HITS_DB, HITS_READ; 4 bases per byte, first base
// in the top two bits, a=0 c=1 g=2 t=3) and DALIGNER (align.h: Overlap / Path; a record on disk is the
// Overlap struct minus its leading trace pointer = 40 bytes, followed by tlen trace bytes -- or 2-byte
// values when tspace > 125) of the FALCON-integrate era, and LA4Falcon's record loop as far as the
// reference's parser pins it.  tests/dazz_writer.py writes files of exactly this layout; the tests
// prove that text path and binary path give identical blocks and consensus, not that the layout
// matches files produced by the real tools.
#include "../../include/falcon_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

// DB.h (FALCON-era DAZZ_DB), 64-bit layout
struct HitsRead { int32_t origin, rlen, fpulse, pad0; int64_t boff, coff; int32_t flags, pad1; };   // 40 bytes
struct HitsDbHeader {                                                                              // 112 bytes
    int32_t ureads, treads, cutoff, all; float freq[4]; int32_t maxlen, pad0; int64_t totlen;
    int32_t nreads, trimmed, part, ufirst, tfirst, pad1; uint64_t path; int32_t loaded, pad2; uint64_t bases, reads, tracks;
};
static_assert(sizeof(HitsRead) == 40, "HITS_READ layout");
static_assert(sizeof(HitsDbHeader) == 112, "HITS_DB layout");
constexpr int DB_BEST = 0x800;
// align.h: what Write_Overlap puts on disk (Overlap minus the trace pointer)
struct OvlRec { int32_t tlen, diffs, abpos, bbpos, aepos, bepos; uint32_t flags; int32_t aread, bread, pad; };
static_assert(sizeof(OvlRec) == 40, "Overlap I/O layout");
constexpr uint32_t COMP_FLAG = 0x1;
constexpr int TRACE_XOVR = 125;

struct Mapped {
    const uint8_t* p = nullptr; size_t n = 0; int fd = -1;
    bool open(const std::string& path, std::string& err) {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) { err = "cannot open " + path; return false; }
        struct stat st;
        if (fstat(fd, &st) != 0) { err = "cannot stat " + path; return false; }
        n = (size_t)st.st_size;
        if (n == 0) { p = nullptr; return true; }
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { err = "cannot mmap " + path; return false; }
        p = (const uint8_t*)m;
        madvise(m, n, MADV_SEQUENTIAL);
        return true;
    }
    void close() { if (p) munmap((void*)p, n); if (fd >= 0) ::close(fd); p = nullptr; n = 0; fd = -1; }
};

}  // namespace

struct fcx_dazz {
    std::string err;
    Mapped idx, bps, las;
    // trimmed reads (the numbering .las files use): length and offset of the compressed bases
    std::vector<int32_t> rlen;
    std::vector<uint64_t> boff;
    // .las cursor
    int64_t novl = 0; int32_t tspace = 0; int tbytes = 1;
    size_t las_pos = 0; int64_t ovl_seen = 0;
    bool las_done = true;
    // block under construction (the parser's state, consensus.py:161-209)
    int32_t p_aread = -1;
    struct Rd { uint32_t pool_id; int32_t len; };
    std::vector<Rd> reads; std::unordered_set<int32_t> ids;
    size_t n_seqs = 0, seed_len = 0; unsigned long long read_cov = 0; int32_t seed_read = -1;
    // output of the last take
    std::vector<uint32_t> boff_out, rids_out; std::vector<char> ids_out;
    uint64_t pairs_out = 0;
};

static thread_local std::string g_dazz_err;
extern "C" const char* fcx_dazz_last_error(const fcx_dazz* d) { return d ? d->err.c_str() : g_dazz_err.c_str(); }

extern "C" void fcx_dazz_close(fcx_dazz* d) {
    if (!d) return;
    d->idx.close(); d->bps.close(); d->las.close();
    delete d;
}

// db_path: the ".db" stub (or the path without the extension); the hidden .<root>.idx / .<root>.bps
// files live next to it (DB.c:Open_DB).
extern "C" int fcx_dazz_open(const char* db_path, fcx_dazz** out) {
    *out = nullptr;
    std::string p(db_path);
    if (p.size() > 3 && p.compare(p.size() - 3, 3, ".db") == 0) p.resize(p.size() - 3);
    const size_t slash = p.find_last_of('/');
    const std::string dir = slash == std::string::npos ? std::string(".") : p.substr(0, slash);
    const std::string root = slash == std::string::npos ? p : p.substr(slash + 1);
    fcx_dazz* d = new fcx_dazz();
    if (!d->idx.open(dir + "/." + root + ".idx", g_dazz_err) || !d->bps.open(dir + "/." + root + ".bps", g_dazz_err)) {
        fcx_dazz_close(d); return 1;
    }
    if (d->idx.n < sizeof(HitsDbHeader)) { g_dazz_err = "truncated .idx (no HITS_DB header)"; fcx_dazz_close(d); return 1; }
    HitsDbHeader h;
    memcpy(&h, d->idx.p, sizeof h);
    if (h.ureads < 0 || h.treads < 0 || h.treads > h.ureads ||
        d->idx.n != sizeof(HitsDbHeader) + (size_t)h.ureads * sizeof(HitsRead)) {
        g_dazz_err = "unexpected .idx layout (this reader knows the 112-byte HITS_DB header + 40-byte HITS_READ records)";
        fcx_dazz_close(d); return 1;
    }
    // Trim_DB (DB.c): reads shorter than the cutoff and, unless `all`, reads that are not the best
    // of their well drop out of the numbering
    const HitsRead* rd = reinterpret_cast<const HitsRead*>(d->idx.p + sizeof(HitsDbHeader));
    const int cutoff = h.cutoff < 0 ? 0 : h.cutoff;
    const int need = h.all ? 0 : DB_BEST;
    for (int32_t i = 0; i < h.ureads; i++) {
        if ((rd[i].flags & need) != need || rd[i].rlen < cutoff) continue;
        if (rd[i].rlen < 0 || (uint64_t)rd[i].boff + (uint64_t)(rd[i].rlen + 3) / 4 > d->bps.n) {
            g_dazz_err = "read " + std::to_string(i) + " points outside the .bps file"; fcx_dazz_close(d); return 1;
        }
        d->rlen.push_back(rd[i].rlen); d->boff.push_back((uint64_t)rd[i].boff);
    }
    *out = d;
    return 0;
}

extern "C" uint32_t fcx_dazz_nreads(const fcx_dazz* d) { return (uint32_t)d->rlen.size(); }
extern "C" int32_t fcx_dazz_read_length(const fcx_dazz* d, uint32_t r) { return r < d->rlen.size() ? d->rlen[r] : -1; }

// The whole DB into the engine's pool, both orientations: entry 2r = read r, 2r + 1 = reverse complement.
extern "C" int fcx_dazz_upload(fcx_dazz* d, fcx_ctx* ctx) {
    const int rc = fcx_pool_upload_bps(ctx, d->bps.p, d->bps.n, d->boff.data(), d->rlen.data(), (uint32_t)d->rlen.size());
    if (rc) d->err = std::string("fcx_pool_upload_bps: ") + fcx_last_error(ctx);
    return rc;
}

extern "C" int fcx_las_open(fcx_dazz* d, const char* las_path) {
    d->las.close();
    d->las_done = true; d->p_aread = -1;
    if (!d->las.open(las_path, d->err)) return 1;
    if (d->las.n < 12) { d->err = "truncated .las (no header)"; return 1; }
    memcpy(&d->novl, d->las.p, 8); memcpy(&d->tspace, d->las.p + 8, 4);
    if (d->novl < 0 || d->tspace < 0) { d->err = "bad .las header"; return 1; }
    d->tbytes = d->tspace <= TRACE_XOVR && d->tspace != 0 ? 1 : 2;
    d->las_pos = 12; d->ovl_seen = 0; d->las_done = false;
    d->reads.clear(); d->ids.clear(); d->n_seqs = 0; d->read_cov = 0; d->seed_read = -1;
    return 0;
}

namespace {

// get_seq_data's "+" (consensus.py:191-195) followed by get_longest_reads(sort=True) (:26-45)
void emit_block(fcx_dazz* d, unsigned min_n_read, unsigned max_n_read, unsigned min_cov_aln, unsigned max_cov_aln) {
    if (d->n_seqs == 0) return;
    if (!(d->n_seqs >= min_n_read && d->read_cov / d->seed_len >= min_cov_aln)) return;
    std::vector<uint32_t> rest(d->reads.size());
    for (uint32_t i = 0; i < rest.size(); i++) rest[i] = i;                 // [seed copy, r1, r2, ...]
    std::stable_sort(rest.begin(), rest.end(), [&](uint32_t a, uint32_t b) { return d->reads[a].len > d->reads[b].len; });
    size_t keep = max_n_read;
    if (max_cov_aln > 0) {
        keep = 1; unsigned long long cov = 0;
        for (size_t i = 0; i < rest.size(); i++) {
            if (cov / d->seed_len > max_cov_aln) break;
            keep++; cov += (unsigned long long)d->reads[rest[i]].len;
        }
        keep = std::min<size_t>(keep, max_n_read);
    }
    size_t n = 0;
    if (keep > 0) { d->rids_out.push_back(d->reads[0].pool_id); n = 1; }        // seqs[0] = seed
    for (size_t i = 0; i < rest.size() && n < keep; i++, n++) d->rids_out.push_back(d->reads[rest[i]].pool_id);
    d->boff_out.push_back((uint32_t)d->rids_out.size());
    char id[32];
    snprintf(id, sizeof id, "%08d", d->seed_read);                              // LA4Falcon prints ids as %08d
    d->ids_out.insert(d->ids_out.end(), id, id + strlen(id) + 1);
    d->pairs_out += n > 0 ? n - 1 : 0;
}

void reset_block(fcx_dazz* d) { d->reads.clear(); d->ids.clear(); d->n_seqs = 0; d->read_cov = 0; d->seed_read = -1; }

// one "<id> <sequence>" line of the LA4Falcon stream, as the parser sees it
void feed_read(fcx_dazz* d, int32_t read, bool comp, unsigned min_len_aln) {
    const int32_t full = d->rlen[read];
    const int32_t slen = full > 100000 ? 99999 : full;                          // consensus.py:178-179
    if ((unsigned)slen < min_len_aln) return;
    if (d->n_seqs == 0) { d->seed_len = (size_t)slen; d->seed_read = read; d->n_seqs = 1; }
    if (d->ids.insert(read).second) {                                            // the seed again, by design
        d->reads.push_back({2u * (uint32_t)read + (comp ? 1u : 0u), slen});
        d->n_seqs++; d->read_cov += (unsigned long long)slen;
    }
}

}  // namespace

// Next batch of seed blocks from the .las (at most max_blocks blocks / max_pairs pairs, at least one
// block).  seed_cutoff is LA4Falcon's -H; the rest are the consensus CLI's parser options.
extern "C" int fcx_las_take(fcx_dazz* d, int seed_cutoff, unsigned min_n_read, unsigned min_len_aln, unsigned max_n_read,
                            unsigned min_cov_aln, unsigned max_cov_aln, uint32_t max_blocks, uint64_t max_pairs,
                            const uint32_t** block_off, const uint32_t** read_ids, uint32_t* n_blocks,
                            const char** seed_ids, int* done) {
    d->boff_out.assign(1, 0u); d->rids_out.clear(); d->ids_out.clear(); d->pairs_out = 0;
    const uint32_t nreads = (uint32_t)d->rlen.size();
    while (!d->las_done && d->boff_out.size() - 1 < max_blocks && d->pairs_out < max_pairs) {
        if (d->ovl_seen >= d->novl) {
            // LA4Falcon's epilogue: "+ +" for the open group, then "- -"
            if (d->p_aread != -1) { emit_block(d, min_n_read, max_n_read, min_cov_aln, max_cov_aln); reset_block(d); }
            d->las_done = true;
            break;
        }
        if (d->las_pos + sizeof(OvlRec) > d->las.n) { d->err = "truncated .las (overlap record)"; return 1; }
        OvlRec o;
        memcpy(&o, d->las.p + d->las_pos, sizeof o);
        d->las_pos += sizeof(OvlRec);
        if (o.tlen < 0 || d->las_pos + (size_t)o.tlen * d->tbytes > d->las.n) { d->err = "truncated .las (trace)"; return 1; }
        d->las_pos += (size_t)o.tlen * d->tbytes;
        d->ovl_seen++;
        if ((uint32_t)o.aread >= nreads || (uint32_t)o.bread >= nreads) { d->err = "overlap refers to a read outside the DB"; return 1; }
        const int32_t alen = d->rlen[o.aread], blen = d->rlen[o.bread];
        // -o: proper overlaps only
        if (o.abpos != 0 && o.bbpos != 0) continue;
        if (o.aepos != alen && o.bepos != blen) continue;
        // -H: seeds of at least this length
        if (alen < seed_cutoff) continue;
        if (o.aread != d->p_aread) {
            if (d->p_aread != -1) { emit_block(d, min_n_read, max_n_read, min_cov_aln, max_cov_aln); reset_block(d); }
            d->p_aread = o.aread;
            feed_read(d, o.aread, false, min_len_aln);                           // the seed line
        }
        feed_read(d, o.bread, (o.flags & COMP_FLAG) != 0, min_len_aln);          // -f: the whole B read, oriented
    }
    *block_off = d->boff_out.data(); *read_ids = d->rids_out.data();
    *n_blocks = (uint32_t)(d->boff_out.size() - 1);
    d->ids_out.push_back('\0');
    *seed_ids = d->ids_out.data();
    if (done) *done = d->las_done ? 1 : 0;
    return 0;
}

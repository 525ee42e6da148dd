// fcx_parser.cu -- host-side parser of the LA4Falcon block stream.
//
// Restates, in C++, what falcon_kit/mains/consensus.py does per line in Python:
//   get_seq_data      consensus.py:161-209  (2-token lines only; sequences > 100000 cut to 99999; the
//                     first read of a block is the seed and is appended twice; duplicate ids are
//                     dropped; "+" emits the block if len(seqs) >= min_n_read and
//                     read_cov // seed_len >= min_cov_aln; "*" discards; "-" stops)
//   get_longest_reads consensus.py:26-45    (seed + stable sort of the rest by -len, capped at
//                     max_n_read / by max_cov_aln)
// and hands the blocks out in the shape fcx_pool_upload / fcx_consensus_blocks take.  Pure host
// code (text handling); no arithmetic of the hot path lives here.
#include "../../include/falcon_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_set>
#include <vector>

namespace {

struct Rd { uint64_t off; uint32_t len; };

struct Block {
    std::string seed_id;
    std::vector<char> data;            // read bytes in arrival order, back to back (the seed once)
    std::vector<Rd> reads;             // distinct reads in arrival order; reads[0] is the seed
    std::vector<uint32_t> order;       // the block as generate_consensus sees it: indices into reads
};

inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// First ASCII-whitespace byte in p[i, n) (or n): eight bytes at a time.  Every whitespace byte is < 0x21, and
// (x - 0x21..21) & ~x & 0x80..80 flags the LOWEST byte of a word that is < 0x21 exactly (higher flags may be
// borrow artefacts), so the candidate is checked with is_space and the scan resumes behind it otherwise.
inline size_t find_space(const char* p, size_t i, size_t n) {
    while (i + 8 <= n) {
        uint64_t x;
        memcpy(&x, p + i, 8);
        const uint64_t m = (x - 0x2121212121212121ull) & ~x & 0x8080808080808080ull;
        if (m == 0) { i += 8; continue; }
        const size_t at = i + (size_t)(__builtin_ctzll(m) >> 3);
        if (is_space((unsigned char)p[at])) return at;
        i = at + 1;
    }
    while (i < n && !is_space((unsigned char)p[i])) i++;
    return i;
}

}  // namespace

// The parser's rules, per line, building seed blocks into `done` (see the header comment).  One
// builder holds the state of ONE open block; a block is closed by a control line, after which the state is
// fresh -- which is what lets complete blocks of a chunk be parsed by independent builders in parallel.
struct BlockBuilder {
    unsigned min_n_read = 0, min_len_aln = 0, max_n_read = 0, min_cov_aln = 0, max_cov_aln = 0;
    bool stopped = false;
    // block under construction.  seqs = [seed, seed, r1, r2, ...] in the reference; here the seed
    // is stored once and referenced twice.
    Block cur;
    size_t n_seqs = 0;                 // len(seqs) of the reference (seed counted twice)
    size_t seed_len = 0;
    unsigned long long read_cov = 0;
    std::unordered_set<std::string> ids;
    std::deque<Block>* done = nullptr;
    // byte buffers of blocks already handed out, kept for reuse: a fresh 1.5 MB vector per seed block is a fresh
    // mmap, i.e. a page fault per 4 kB, every time
    struct Spare { std::mutex m; std::vector<std::vector<char>> v; };
    Spare* spare = nullptr;
    void fresh_data() {
        cur.data = std::vector<char>();
        if (!spare) return;
        std::lock_guard<std::mutex> g(spare->m);
        if (!spare->v.empty()) { cur.data = std::move(spare->v.back()); spare->v.pop_back(); cur.data.clear(); }
    }

    void reset_block() {
        cur.seed_id.clear(); cur.data.clear(); cur.reads.clear(); cur.order.clear();
        ids.clear(); n_seqs = 0; read_cov = 0;
    }

    void emit_block() {
        if (n_seqs == 0) return;       // reference: ZeroDivisionError territory (read_cov // 0); skipped here
        if (!(n_seqs >= min_n_read && read_cov / seed_len >= min_cov_aln)) return;
        // seqs[1:] in arrival order: the seed copy is present iff the seed's id was not a duplicate,
        // which it never is (it is the first id of the block); then the other distinct reads
        std::vector<uint32_t> rest;
        rest.reserve(cur.reads.size());
        for (uint32_t i = 0; i < cur.reads.size(); i++) rest.push_back(i);      // [seed copy, r1, r2, ...]
        // get_longest_reads(sort=True): stable sort of seqs[1:] by -len
        std::stable_sort(rest.begin(), rest.end(),
                         [&](uint32_t a, uint32_t b) { return cur.reads[a].len > cur.reads[b].len; });
        size_t keep = max_n_read;
        if (max_cov_aln > 0) {
            keep = 1; unsigned long long cov = 0;
            for (size_t i = 0; i < rest.size(); i++) {
                if (cov / seed_len > max_cov_aln) break;
                keep++; cov += cur.reads[rest[i]].len;
            }
            keep = std::min<size_t>(keep, max_n_read);
        }
        cur.order.clear();
        cur.order.push_back(0);                                                  // seqs[0] = seed
        for (size_t i = 0; i < rest.size() && cur.order.size() < keep; i++) cur.order.push_back(rest[i]);
        if (keep == 0) cur.order.clear();                                       // seqs[:0]
        done->push_back(std::move(cur));
        cur = Block();
        fresh_data();
    }

    // l.strip().split(): tokens separated by ASCII whitespace; exactly two are required.
    // Returns the number of tokens found (3 = "more than two").
    static int split2(const char* p, size_t n, const char* tok[2], size_t len[2]) {
        size_t i = 0; int nt = 0;
        while (i < n) {
            while (i < n && is_space((unsigned char)p[i])) i++;
            if (i >= n) break;
            size_t s = i;
            i = find_space(p, i, n);
            if (nt < 2) { tok[nt] = p + s; len[nt] = i - s; }
            nt++;
            if (nt > 2) return 3;
        }
        return nt;
    }
    static bool is_control(const char* tok0, size_t len0) {
        return len0 == 1 && (tok0[0] == '+' || tok0[0] == '-' || tok0[0] == '*');
    }

    void line(const char* p, size_t n) {
        if (stopped) return;
        const char* tok[2] = {nullptr, nullptr}; size_t len[2] = {0, 0};
        if (split2(p, n, tok, len) != 2) return;
        size_t slen = len[1];
        if (slen > 100000) slen = 99999;                                  // consensus.py:178-179
        if (!is_control(tok[0], len[0])) {
            if (slen >= min_len_aln) {
                std::string id(tok[0], len[0]);
                const bool first = n_seqs == 0;
                if (first) { seed_len = slen; cur.seed_id = id; n_seqs = 1; }                 // the seed
                if (ids.insert(id).second) {                                                   // seed again, by design
                    Rd r; r.off = cur.data.size(); r.len = (uint32_t)slen;
                    cur.data.insert(cur.data.end(), tok[1], tok[1] + slen);
                    cur.reads.push_back(r);
                    n_seqs++; read_cov += slen;
                }
            }
        } else if (tok[0][0] == '+') { emit_block(); reset_block(); }
        else if (tok[0][0] == '*') { reset_block(); }
        else { stopped = true; }
    }

    // every complete line of p[0, n); returns the number of bytes consumed (the rest is a partial line)
    size_t lines(const char* p, size_t n) {
        size_t at = 0;
        while (at < n && !stopped) {
            const char* nl = (const char*)memchr(p + at, '\n', n - at);
            if (!nl) break;
            line(p + at, (size_t)(nl - (p + at)));
            at = (size_t)(nl - p) + 1;
        }
        return stopped ? n : at;
    }
};

struct fcx_parser {
    BlockBuilder::Spare spare;
    BlockBuilder B;                    // the open block of the stream
    std::string carry;                 // partial last line of the previous chunk
    std::deque<Block> ready;
    unsigned n_threads = 1;
    size_t par_min = (size_t)1 << 20;  // chunks smaller than this are parsed by the calling thread alone
    // storage handed out by fcx_parser_take: two sets used alternately, so that the batch of the
    // previous take() stays valid while the caller (another thread) is still consuming it
    struct Out { std::vector<char> bases, ids; std::vector<uint64_t> off; std::vector<uint32_t> boff, rids; };
    Out outs[2];
    unsigned take_no = 0;
};

extern "C" fcx_parser* fcx_parser_create(unsigned min_n_read, unsigned min_len_aln, unsigned max_n_read,
                                         unsigned min_cov_aln, unsigned max_cov_aln) {
    fcx_parser* p = new fcx_parser();
    p->B.min_n_read = min_n_read; p->B.min_len_aln = min_len_aln; p->B.max_n_read = max_n_read;
    p->B.min_cov_aln = min_cov_aln; p->B.max_cov_aln = max_cov_aln;
    p->B.done = &p->ready;
    p->B.spare = &p->spare;
    // FCX_PARSER_THREADS > 1: complete blocks of a chunk are parsed by worker threads.  Off by default: on the
    // machines measured the single-threaded parser (~2 GB/s of feed) is bound by memory traffic, not by tokenising.
    p->n_threads = 1;
    if (const char* e = getenv("FCX_PARSER_THREADS")) p->n_threads = (unsigned)std::max(1, atoi(e));
    if (const char* e = getenv("FCX_PARSER_PAR_MIN")) p->par_min = (size_t)std::max(0, atoi(e));
    return p;
}
extern "C" void fcx_parser_destroy(fcx_parser* p) { delete p; }

// Control lines ("+", "*", "-" as the first of exactly two tokens) of p[0, n), complete lines only:
// {end offset of the line (behind its newline), kind}.  Read lines are ~15 kB each, so this pass is one memchr
// per line plus a tokenisation of the few lines that start with a control character.
static void find_control_lines(const char* p, size_t n, std::vector<std::pair<size_t, char>>& out) {
    size_t at = 0;
    while (at < n) {
        const char* nl = (const char*)memchr(p + at, '\n', n - at);
        if (!nl) break;
        const size_t ln = (size_t)(nl - (p + at));
        size_t i = at;
        while (i < at + ln && is_space((unsigned char)p[i])) i++;
        if (i < at + ln && (p[i] == '+' || p[i] == '*' || p[i] == '-') && (i + 1 == at + ln || is_space((unsigned char)p[i + 1]))) {
            const char* tok[2]; size_t len[2];
            if (BlockBuilder::split2(p + at, ln, tok, len) == 2 && BlockBuilder::is_control(tok[0], len[0]))
                out.emplace_back((size_t)(nl - p) + 1, tok[0][0]);
        }
        at = (size_t)(nl - p) + 1;
    }
}

extern "C" int fcx_parser_feed(fcx_parser* ps, const char* data, size_t n, int eof) {
    BlockBuilder& B = ps->B;
    size_t start = 0;
    if (!ps->carry.empty()) {
        const char* nl = (const char*)memchr(data, '\n', n);
        if (!nl) { ps->carry.append(data, n); start = n; }
        else {
            ps->carry.append(data, (size_t)(nl - data));
            B.line(ps->carry.data(), ps->carry.size());
            ps->carry.clear();
            start = (size_t)(nl - data) + 1;
        }
    }
    if (start < n && !B.stopped) {
        const char* p = data + start; const size_t m = n - start;
        std::vector<std::pair<size_t, char>> ctl;
        if (ps->n_threads > 1 && m >= ps->par_min) find_control_lines(p, m, ctl);
        // a "-" ends the stream: nothing behind it counts
        size_t n_ctl = ctl.size();
        for (size_t i = 0; i < ctl.size(); i++) if (ctl[i].second == '-') { n_ctl = i + 1; break; }
        if (n_ctl >= 3) {
            // [0, ctl[0]) continues the open block (this thread); every [ctl[i-1], ctl[i]) is one complete block
            // with fresh state (worker threads); what follows the last control line is parsed afterwards
            const size_t n_seg = n_ctl - 1;
            std::vector<std::deque<Block>> seg_out(n_seg);
            std::vector<char> seg_stop(n_seg, 0);
            std::atomic<size_t> next(0);
            auto work = [&]() {
                for (;;) {
                    const size_t k = next.fetch_add(1);
                    if (k >= n_seg) break;
                    BlockBuilder W;
                    W.min_n_read = B.min_n_read; W.min_len_aln = B.min_len_aln; W.max_n_read = B.max_n_read;
                    W.min_cov_aln = B.min_cov_aln; W.max_cov_aln = B.max_cov_aln;
                    W.done = &seg_out[k];
                    W.spare = &ps->spare;
                    W.fresh_data();
                    W.lines(p + ctl[k].first, ctl[k + 1].first - ctl[k].first);
                    seg_stop[k] = W.stopped ? 1 : 0;
                }
            };
            std::vector<std::thread> th;
            const unsigned nt = (unsigned)std::min<size_t>(ps->n_threads - 1, n_seg);
            for (unsigned i = 0; i < nt; i++) th.emplace_back(work);
            B.lines(p, ctl[0].first);                      // the open block, up to and including its control line
            work();                                        // then this thread helps with the segments
            for (auto& t : th) t.join();
            for (size_t k = 0; k < n_seg && !B.stopped; k++) {
                for (auto& blk : seg_out[k]) ps->ready.push_back(std::move(blk));
                if (seg_stop[k]) B.stopped = true;
            }
            start += ctl[n_ctl - 1].first;
        }
    }
    if (start < n && !B.stopped) {
        const size_t used = B.lines(data + start, n - start);
        if (used < n - start) ps->carry.assign(data + start + used, n - start - used);
    }
    if (eof && !ps->carry.empty()) { B.line(ps->carry.data(), ps->carry.size()); ps->carry.clear(); }
    return B.stopped ? -(int)ps->ready.size() - 1 : (int)ps->ready.size();
}

extern "C" int fcx_parser_pending(const fcx_parser* ps) { return (int)ps->ready.size(); }
extern "C" int fcx_parser_stopped(const fcx_parser* ps) { return ps->B.stopped ? 1 : 0; }

extern "C" int fcx_parser_take(fcx_parser* ps, uint32_t max_blocks, uint64_t max_bases, const char** bases,
                               const uint64_t** offsets, uint32_t* n_reads, const uint32_t** block_off,
                               const uint32_t** read_ids, uint32_t* n_blocks, const char** seed_ids) {
    fcx_parser::Out& o = ps->outs[ps->take_no++ & 1u];
    o.bases.clear(); o.ids.clear(); o.off.assign(1, 0); o.boff.assign(1, 0); o.rids.clear();
    uint32_t nb = 0; uint64_t total = 0;
    while (!ps->ready.empty() && nb < max_blocks) {
        Block& b = ps->ready.front();
        const uint64_t sz = b.data.size();
        if (nb > 0 && total + sz > max_bases) break;
        const uint32_t base_id = (uint32_t)(o.off.size() - 1);
        const uint64_t base_off = o.bases.size();
        o.bases.insert(o.bases.end(), b.data.begin(), b.data.end());      // one copy per block
        for (auto& r : b.reads) o.off.push_back(base_off + r.off + r.len);
        for (uint32_t k : b.order) o.rids.push_back(base_id + k);
        o.boff.push_back((uint32_t)o.rids.size());
        o.ids.insert(o.ids.end(), b.seed_id.begin(), b.seed_id.end());
        o.ids.push_back('\0');
        total += sz; nb++;
        {   // the block's byte buffer goes back to the pool (bounded: what a few batches need)
            std::lock_guard<std::mutex> g(ps->spare.m);
            if (ps->spare.v.size() < 4096) ps->spare.v.push_back(std::move(b.data));
        }
        ps->ready.pop_front();
    }
    o.bases.push_back('\0');
    *bases = o.bases.data(); *offsets = o.off.data(); *n_reads = (uint32_t)(o.off.size() - 1);
    *block_off = o.boff.data(); *read_ids = o.rids.data(); *n_blocks = nb; *seed_ids = o.ids.data();
    return 0;
}

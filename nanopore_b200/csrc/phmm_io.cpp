// phmm_io.cpp -- libphmm_io.so: native ingest / chain / pack / emit of the realignment path (include/phmm_io.h).
//
// What the reference does record by record in Python 2 on pysam (nanopore/analyses/utils.py:233-245, 287-469,
// 557-574, 591-609) is done here on whole files with host threads: the SAM text is split into lines once, records
// are parsed in parallel, (read, reference) groups are chained in parallel, and the output text is formatted in
// parallel chunks.  The semantics (grouping order, chain tie-breaking, sort order of the chained file, which
// fields are re-serialised) follow nanopore_b200/realign.py and nanopore_b200/sam.py, which tests compare byte
// for byte.  Host C++ only: no CUDA, no Python, nothing from oracle/.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/phmm_io.h"

namespace {

using sv = std::string_view;

struct IoError {
    int code;
    std::string msg;
};

[[noreturn]] void fail(int code, std::string msg) { throw IoError{code, std::move(msg)}; }

std::string read_file(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) fail(PHMM_IO_E_FILE, std::string("cannot open ") + path + ": " + strerror(errno));
    std::string s;
    if (fseek(f, 0, SEEK_END) == 0) {
        long n = ftell(f);
        if (n > 0) s.reserve((size_t)n);
        fseek(f, 0, SEEK_SET);
    }
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, n);
    const bool bad = ferror(f) != 0;
    fclose(f);
    if (bad) fail(PHMM_IO_E_FILE, std::string("read error on ") + path);
    return s;
}

// runs fn(i) for i in [0, n) on up to `threads` host threads; the first exception is rethrown
void parallel_for(int64_t n, int threads, const std::function<void(int64_t)> &fn) {
    if (n <= 0) return;
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n));
    if (nt == 1) { for (int64_t i = 0; i < n; i++) fn(i); return; }
    std::atomic<int64_t> next{0};
    std::mutex mu;
    bool failed = false;
    IoError err{0, ""};
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&] {
            for (;;) {
                const int64_t i = next.fetch_add(1);
                if (i >= n) break;
                try { fn(i); }
                catch (const IoError &e) { std::lock_guard<std::mutex> g(mu); if (!failed) { failed = true; err = e; } next = n; }
                catch (const std::exception &e) { std::lock_guard<std::mutex> g(mu); if (!failed) { failed = true; err = IoError{PHMM_IO_E_ARG, e.what()}; } next = n; }
            }
        });
    for (auto &th : pool) th.join();
    if (failed) throw err;
}

sv rstrip_crlf(sv s) {
    while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.remove_suffix(1);
    return s;
}

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

sv first_word(sv s) {
    size_t a = 0;
    while (a < s.size() && is_space(s[a])) a++;
    size_t b = a;
    while (b < s.size() && !is_space(s[b])) b++;
    return s.substr(a, b - a);
}

// lines of a text buffer, '\n' kept out
std::vector<sv> split_lines(const std::string &buf) {
    std::vector<sv> out;
    const char *p = buf.data(), *e = p + buf.size();
    while (p < e) {
        const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
        if (!q) q = e;
        out.emplace_back(p, (size_t)(q - p));
        p = q + 1;
    }
    return out;
}

int64_t parse_int(sv s, const char *what) {
    if (s.empty()) fail(PHMM_IO_E_FORMAT, std::string("empty ") + what + " field");
    size_t i = 0;
    bool neg = false;
    if (s[0] == '-' || s[0] == '+') { neg = s[0] == '-'; i = 1; }
    if (i == s.size()) fail(PHMM_IO_E_FORMAT, std::string("bad ") + what + " field " + std::string(s));
    int64_t v = 0;
    for (; i < s.size(); i++) {
        if (s[i] < '0' || s[i] > '9') fail(PHMM_IO_E_FORMAT, std::string("bad ") + what + " field " + std::string(s));
        v = v * 10 + (s[i] - '0');
    }
    return neg ? -v : v;
}

uint8_t base_code(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

std::string reverse_complement(sv s) {
    std::string out(s.size(), 'N');
    for (size_t i = 0; i < s.size(); i++) {
        char c = s[s.size() - 1 - i];
        switch (c) {
            case 'A': c = 'T'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; case 'T': c = 'A'; break;
            case 'a': c = 't'; break; case 'c': c = 'g'; break; case 'g': c = 'c'; break; case 't': c = 'a'; break;
            default: break;                                       // N, n and anything else stay
        }
        out[i] = c;
    }
    return out;
}

struct CigOp { uint8_t code; int64_t len; };
const char OPS[] = "MIDNSHP=X";

// "<number><op>..." -> ops; '*' / empty -> none.  Same validation as sam.parse_cigar_arrays.
void parse_cigar(sv s, std::vector<CigOp> &out) {
    out.clear();
    if (s.empty() || s == "*") return;
    int64_t v = 0;
    int nd = 0;
    for (char c : s) {
        if (c >= '0' && c <= '9') {
            v = v * 10 + (c - '0');
            if (++nd > 18) fail(PHMM_IO_E_FORMAT, "malformed CIGAR " + std::string(s.substr(0, 60)));
        } else {
            const char *p = (const char *)memchr(OPS, c, 9);
            if (!p || nd == 0) fail(PHMM_IO_E_FORMAT, "malformed CIGAR " + std::string(s.substr(0, 60)));
            out.push_back({(uint8_t)(p - OPS), v});
            v = 0; nd = 0;
        }
    }
    if (nd != 0) fail(PHMM_IO_E_FORMAT, "malformed CIGAR " + std::string(s.substr(0, 60)));
}

void append_int(std::string &s, int64_t v) {
    char buf[24];
    int n = snprintf(buf, sizeof(buf), "%lld", (long long)v);
    s.append(buf, (size_t)n);
}

struct Rec {
    sv qname;
    int64_t flag = 0;
    int32_t rname = -1, rnext = -1;
    int64_t pos = -1, mapq = 0, pnext = -1, tlen = 0;
    sv cigar;            // empty: none ("*")
    sv seq;              // empty: none ("*")
    sv qual;             // empty: none ("*")
    sv tags;             // the optional fields as they stood, tabs included, without the leading tab
    sv rname_s, rnext_s; // as read (resolved to ids sequentially after the parallel parse)
    bool reverse() const { return (flag & 0x10) != 0; }
};

struct Seqs {
    std::string buf;                                  // file contents (FASTQ) or the concatenated sequences (FASTA)
    std::vector<std::string> names;
    std::vector<sv> seq;                              // views into buf / owned
    std::deque<std::string> owned;
    std::unordered_map<std::string, int64_t> index;
    bool loaded = false;
    void clear() { buf.clear(); names.clear(); seq.clear(); owned.clear(); index.clear(); loaded = false; }
};

}  // namespace

struct phmm_io {
    int threads = 1;
    std::string err;
    Seqs ref, reads;
    // packed reference (all contigs in file order)
    std::vector<uint8_t> ref_codes;
    std::vector<int64_t> ref_offset;
    // SAM
    std::string sam_buf;
    std::vector<std::string> header_lines;
    std::vector<std::string> references;
    std::unordered_map<std::string, int32_t> tid;
    std::vector<Rec> recs;
    std::deque<std::string> arena;                    // strings owned by chained records
    bool have_records = false;
    // batch
    std::vector<int64_t> mapped;                      // indices of the mapped records
    std::vector<uint8_t> b_reads;
    std::vector<int64_t> b_read_off, b_ref_start, b_ref_end, b_ops_off;
    std::vector<uint32_t> b_ops;
    bool have_batch = false;
};

namespace {

void load_fasta(Seqs &s, const char *path) {
    s.clear();
    const std::string file = read_file(path);
    const std::vector<sv> lines = split_lines(file);
    // fastaRead: header lines start with '>', sequence lines lose every white-space character (bioio.fastaRead)
    std::vector<std::pair<std::string, std::string>> recs;
    bool open = false;
    for (sv ln : lines) {
        ln = rstrip_crlf(ln);
        if (!ln.empty() && ln[0] == '>') {
            recs.emplace_back(std::string(ln.substr(1)), std::string());
            open = true;
        } else if (open) {
            std::string &dst = recs.back().second;
            for (char c : ln) if (!is_space(c)) dst.push_back(c);
        }
    }
    for (auto &r : recs) {
        const sv w = first_word(r.first);
        if (w.empty()) fail(PHMM_IO_E_FORMAT, std::string("fasta header without a name in ") + path);
        std::string name(w);
        if (s.index.count(name)) fail(PHMM_IO_E_FORMAT, "duplicate sequence name " + name + " in " + path);
        s.index[name] = (int64_t)s.names.size();
        s.names.push_back(name);
        s.owned.push_back(std::move(r.second));
        s.seq.emplace_back(s.owned.back());
    }
    s.loaded = true;
}

void load_fastq(Seqs &s, const char *path) {
    s.clear();
    s.buf = read_file(path);
    const std::vector<sv> lines = split_lines(s.buf);
    size_t i = 0;
    while (i < lines.size()) {
        const sv line = lines[i];
        if (!line.empty() && line[0] == '@') {
            const sv name_line = rstrip_crlf(line.substr(1));
            const sv seq = i + 1 < lines.size() ? rstrip_crlf(lines[i + 1]) : sv();
            const sv plus = i + 2 < lines.size() ? lines[i + 2] : sv();
            if (plus.empty() || plus[0] != '+') fail(PHMM_IO_E_FORMAT, "Got unexpected line: " + std::string(plus.substr(0, 60)));
            const sv qual = i + 3 < lines.size() ? rstrip_crlf(lines[i + 3]) : sv();
            if (seq.size() == qual.size())
                for (char c : qual)
                    if ((unsigned char)c < 33 || (unsigned char)c > 126)
                        fail(PHMM_IO_E_FORMAT, "Got a qual value out of range for sequence " + std::string(name_line));
            const sv w = first_word(name_line);
            if (w.empty()) fail(PHMM_IO_E_FORMAT, std::string("fastq record without a name in ") + path);
            std::string name(w);
            if (s.index.count(name)) fail(PHMM_IO_E_FORMAT, "duplicate sequence name " + name + " in " + path);
            s.index[name] = (int64_t)s.names.size();
            s.names.push_back(name);
            s.seq.push_back(seq);
            i += 4;
        } else i++;
    }
    s.loaded = true;
}

int32_t tid_of(phmm_io *io, sv name) {
    if (name == "*") return -1;
    std::string n(name);
    auto it = io->tid.find(n);
    if (it != io->tid.end()) return it->second;
    const int32_t id = (int32_t)io->references.size();       // header-less SAM: registered on the fly (sam.Samfile._tid_of)
    io->tid[n] = id;
    io->references.push_back(n);
    return id;
}

void add_header(phmm_io *io, sv line) {
    io->header_lines.emplace_back(line);
    if (line.substr(0, 3) == "@SQ") {
        size_t p = 0;
        bool first = true, have = false;
        std::string name;
        while (p <= line.size()) {
            size_t q = line.find('\t', p);
            if (q == sv::npos) q = line.size();
            const sv f = line.substr(p, q - p);
            if (!first && f.substr(0, 3) == "SN:") { name = std::string(f.substr(3)); have = true; }
            first = false;
            p = q + 1;
        }
        if (have && !io->tid.count(name)) {
            io->tid[name] = (int32_t)io->references.size();
            io->references.push_back(name);
        }
    }
}

// splits the SAM text into header and record lines, parses the records in parallel, resolves names sequentially
void read_sam(phmm_io *io, const char *path) {
    io->sam_buf = read_file(path);
    io->header_lines.clear(); io->references.clear(); io->tid.clear(); io->recs.clear(); io->arena.clear();
    io->have_records = false; io->have_batch = false;
    const std::vector<sv> lines = split_lines(io->sam_buf);
    size_t i = 0;
    for (; i < lines.size(); i++) {                          // the header block: leading '@' lines
        const sv ln = lines[i];
        if (!ln.empty() && ln[0] == '@') add_header(io, rstrip_crlf(ln));
        else break;
    }
    std::vector<sv> rl;
    rl.reserve(lines.size() - i);
    for (; i < lines.size(); i++) {
        const sv ln = rstrip_crlf(lines[i]);
        if (ln.empty() || ln[0] == '@') continue;
        rl.push_back(ln);
    }
    io->recs.resize(rl.size());
    const int64_t n = (int64_t)rl.size();
    const int64_t chunk = 256;
    parallel_for((n + chunk - 1) / chunk, io->threads, [&](int64_t c) {
        for (int64_t k = c * chunk; k < std::min(n, (c + 1) * chunk); k++) {
            const sv line = rl[k];
            sv f[11];
            size_t p = 0;
            int nf = 0;
            while (nf < 11) {
                size_t q = line.find('\t', p);
                if (q == sv::npos) { f[nf++] = line.substr(p); p = line.size() + 1; break; }
                f[nf++] = line.substr(p, q - p);
                p = q + 1;
            }
            if (nf < 11) fail(PHMM_IO_E_FORMAT, "SAM record with " + std::to_string(nf) + " fields: " + std::string(line.substr(0, 60)));
            Rec &r = io->recs[k];
            r.qname = f[0];
            r.flag = parse_int(f[1], "FLAG");
            r.rname_s = f[2];
            r.pos = parse_int(f[3], "POS") - 1;
            r.mapq = parse_int(f[4], "MAPQ");
            r.cigar = (f[5] == "*") ? sv() : f[5];
            r.rnext_s = f[6];
            r.pnext = parse_int(f[7], "PNEXT") - 1;
            r.tlen = parse_int(f[8], "TLEN");
            r.seq = (f[9] == "*") ? sv() : f[9];
            r.qual = (f[10] == "*") ? sv() : f[10];
            r.tags = p <= line.size() ? line.substr(p) : sv();
            if (p > line.size()) r.tags = sv();
        }
    });
    for (Rec &r : io->recs) {                                 // ids in file order, exactly as a sequential reader hands them out
        r.rname = tid_of(io, r.rname_s);
        r.rnext = (r.rnext_s == "=") ? r.rname : tid_of(io, r.rnext_s);
    }
    io->have_records = true;
}

void format_record(const phmm_io *io, const Rec &r, sv cigar_text, std::string &out) {
    out.append(r.qname); out.push_back('\t');
    append_int(out, r.flag); out.push_back('\t');
    if (r.rname >= 0) out.append(io->references[(size_t)r.rname]); else out.push_back('*');
    out.push_back('\t');
    append_int(out, r.pos + 1); out.push_back('\t');
    append_int(out, r.mapq); out.push_back('\t');
    if (cigar_text.empty()) out.push_back('*'); else out.append(cigar_text);
    out.push_back('\t');
    if (r.rnext < 0) out.push_back('*');
    else if (r.rnext == r.rname) out.push_back('=');
    else out.append(io->references[(size_t)r.rnext]);
    out.push_back('\t');
    append_int(out, r.pnext + 1); out.push_back('\t');
    append_int(out, r.tlen); out.push_back('\t');
    if (r.seq.empty()) out.push_back('*'); else out.append(r.seq);
    out.push_back('\t');
    if (r.qual.empty()) out.push_back('*'); else out.append(r.qual);
    if (!r.tags.empty()) { out.push_back('\t'); out.append(r.tags); }
    out.push_back('\n');
}

void write_chunks(const char *path, const std::vector<std::string> &header, const std::vector<std::string> &chunks) {
    FILE *f = fopen(path, "wb");
    if (!f) fail(PHMM_IO_E_FILE, std::string("cannot create ") + path + ": " + strerror(errno));
    bool ok = true;
    for (const std::string &h : header) ok = ok && fwrite(h.data(), 1, h.size(), f) == h.size() && fputc('\n', f) != EOF;
    for (const std::string &c : chunks) ok = ok && fwrite(c.data(), 1, c.size(), f) == c.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) fail(PHMM_IO_E_FILE, std::string("write error on ") + path);
}

// ---------------------------------------------------------------------------------------------------------
// chaining (utils.py:295-469 as restated in nanopore_b200/realign.py)
// ---------------------------------------------------------------------------------------------------------
struct Hit {
    int64_t rec;                       // record index
    std::vector<CigOp> ops;
    int64_t score = 0, rStart = 0, qStart = 0, rEnd = 0, qEnd = 0;
    int64_t readOffset = 0;            // signed position in the read of the first non-clipped base
};

// getAbsoluteReadOffset on the parsed cigar (utils.py:155-166)
int64_t read_offset(const std::vector<CigOp> &ops, bool reverse, int64_t read_len) {
    int64_t off = (!ops.empty() && ops[0].code == 5) ? ops[0].len : 0;
    if (reverse) off = -(read_len - 1 - off);
    for (const CigOp &o : ops) {
        if (o.code == 5) continue;
        if (o.code == 4) off += o.len;
        break;
    }
    return off;
}

void summarise(Hit &h, const Rec &r, int64_t read_len) {
    parse_cigar(r.cigar, h.ops);
    h.readOffset = read_offset(h.ops, r.reverse(), read_len);
    int64_t q = 0, x = r.pos;
    bool any = false;
    h.score = 0;
    for (const CigOp &o : h.ops) {
        if (o.code == 0) {
            if (!any) { h.rStart = x; h.qStart = h.readOffset + q; any = true; }
            h.rEnd = x + o.len - 1; h.qEnd = h.readOffset + q + o.len - 1;
            h.score += o.len;
        }
        if (o.code == 0 || o.code == 1) q += o.len;
        if (o.code == 0 || o.code == 2) x += o.len;
    }
    if (!any) fail(PHMM_IO_E_FORMAT, "alignment of " + std::string(r.qname) + " has no aligned positions");
}

// chainFn (utils.py:388-426): returns the chain as indices into hits, first to last
std::vector<int> chain_hits(std::vector<Hit> &hits, const std::vector<Rec> &recs, int64_t max_gap = 200) {
    const int n = (int)hits.size();
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hits[a].rStart < hits[b].rStart; });
    std::vector<int64_t> score(n);
    std::vector<int> ptr(n, -1);
    for (int i = 0; i < n; i++) score[i] = hits[i].score;
    for (int oi = 0; oi < n; oi++) {
        const int i = order[oi];
        const Hit &a = hits[i];
        const int64_t own = a.score;
        for (int oj = 0; oj < oi; oj++) {
            const int j = order[oj];
            const Hit &b = hits[j];
            if (a.rStart > b.rEnd && a.qStart > b.qEnd && recs[a.rec].reverse() == recs[b.rec].reverse() &&
                a.rStart - b.rEnd + a.qStart - b.qEnd <= max_gap && own + score[j] > score[i]) {
                score[i] = own + score[j];
                ptr[i] = j;
            }
        }
    }
    // sorted(..., key=score)[-1]: of the hits with the top score, the last in reference order
    int best = order[0];
    for (int oi = 1; oi < n; oi++) if (score[order[oi]] >= score[best]) best = order[oi];
    std::vector<int> chain;
    for (int k = best; k >= 0; k = ptr[k]) chain.push_back(k);
    std::reverse(chain.begin(), chain.end());
    return chain;
}

struct Chained {
    Rec rec;
    std::string seq, cigar;
};

// mergeChainedAlignedReads (utils.py:295-386)
void merge_chain(const std::vector<Hit> &hits, const std::vector<int> &chain, const std::vector<Rec> &recs, sv refSeq, sv readSeq,
                 Chained &out) {
    const Rec &first = recs[hits[chain[0]].rec];
    const bool rev = first.reverse();
    out.rec = Rec();
    out.rec.qname = first.qname;
    out.rec.flag = rev ? 0x10 : 0;
    out.rec.rname = first.rname;
    out.rec.rnext = -1;
    out.rec.pos = 0; out.rec.mapq = 0; out.rec.pnext = -1; out.rec.tlen = 0;
    out.seq = rev ? reverse_complement(readSeq) : std::string(readSeq);
    std::string &cg = out.cigar;
    cg.clear();
    auto emit = [&](int code, int64_t len) { append_int(cg, len); cg.push_back(OPS[code]); };
    const int64_t lref = (int64_t)refSeq.size(), lread = (int64_t)readSeq.size();
    int64_t pPos = 0, pQPos = rev ? -(lread - 1) : 0, spanX = 0, spanY = 0;
    for (int k : chain) {
        const Hit &h = hits[k];
        const Rec &r = recs[h.rec];
        if (r.reverse() != rev) fail(PHMM_IO_E_FORMAT, "chain of " + std::string(first.qname) + " mixes strands");
        if (r.pos < pPos) fail(PHMM_IO_E_FORMAT, "chain of " + std::string(first.qname) + " overlaps on the reference");
        if (r.pos > pPos) { emit(2, r.pos - pPos); spanX += r.pos - pPos; pPos = r.pos; }
        for (const CigOp &o : h.ops)
            if (!(o.code <= 2 || o.code == 4 || o.code == 5))
                fail(PHMM_IO_E_FORMAT, "cigar of " + std::string(first.qname) + " holds an operation other than M, I, D, S, H");
        const int64_t qPos = h.readOffset;
        if (qPos < pQPos) fail(PHMM_IO_E_FORMAT, "chain of " + std::string(first.qname) + " overlaps on the read");
        if (qPos > pQPos) { emit(1, qPos - pQPos); spanY += qPos - pQPos; pQPos = qPos; }
        for (const CigOp &o : h.ops) {
            if (o.code > 2) continue;
            emit(o.code, o.len);
            if (o.code == 0 || o.code == 2) { pPos += o.len; spanX += o.len; }
            if (o.code == 0 || o.code == 1) { pQPos += o.len; spanY += o.len; }
        }
    }
    if (pPos > lref) fail(PHMM_IO_E_FORMAT, "alignment of " + std::string(first.qname) + " runs past the end of its reference");
    if (pPos < lref) { emit(2, lref - pPos); spanX += lref - pPos; }
    if (rev) {
        if (pQPos > 1) fail(PHMM_IO_E_FORMAT, "alignment of " + std::string(first.qname) + " runs past the end of the read");
        if (pQPos < 1) { emit(1, -pQPos + 1); spanY += -pQPos + 1; }
    } else {
        if (pQPos > lread) fail(PHMM_IO_E_FORMAT, "alignment of " + std::string(first.qname) + " runs past the end of the read");
        if (pQPos < lread) { emit(1, lread - pQPos); spanY += lread - pQPos; }
    }
    if (spanX != lref || spanY != lread)
        fail(PHMM_IO_E_FORMAT, "chained alignment of " + std::string(first.qname) + " does not span both sequences (utils.py:381-382)");
}

void chain_sam(phmm_io *io, const char *path) {
    if (!io->ref.loaded || !io->reads.loaded) fail(PHMM_IO_E_STATE, "phmm_io_chain_sam needs the reference and the reads");
    read_sam(io, path);
    // groups in first-seen order of (read name, reference id) over the mapped records (samIterator, utils.py:287-293)
    struct Group { std::vector<int64_t> recs; int64_t read_idx; };
    std::vector<Group> groups;
    std::unordered_map<std::string, size_t> gidx;
    for (int64_t k = 0; k < (int64_t)io->recs.size(); k++) {
        const Rec &r = io->recs[k];
        if (r.rname == -1) continue;
        auto rit = io->reads.index.find(std::string(r.qname));
        if (rit == io->reads.index.end()) fail(PHMM_IO_E_FORMAT, "Aligned read name: " + std::string(r.qname) + " not in read sequences names");
        std::string key(r.qname);
        key.push_back('\t');
        key += std::to_string(r.rname);
        auto it = gidx.find(key);
        if (it == gidx.end()) { it = gidx.emplace(key, groups.size()).first; groups.push_back({{}, rit->second}); }
        groups[it->second].recs.push_back(k);
    }
    std::vector<Chained> out(groups.size());
    const std::vector<Rec> &recs = io->recs;
    parallel_for((int64_t)groups.size(), io->threads, [&](int64_t g) {
        const Group &G = groups[g];
        const Rec &r0 = recs[G.recs[0]];
        auto fit = io->ref.index.find(io->references[(size_t)r0.rname]);
        if (fit == io->ref.index.end())
            fail(PHMM_IO_E_FORMAT, "Reference sequence " + io->references[(size_t)r0.rname] + " of read " + std::string(r0.qname) + " not in the reference fasta");
        const sv refSeq = io->ref.seq[(size_t)fit->second], readSeq = io->reads.seq[(size_t)G.read_idx];
        std::vector<Hit> hits(G.recs.size());
        for (size_t h = 0; h < hits.size(); h++) { hits[h].rec = G.recs[h]; summarise(hits[h], recs[G.recs[h]], (int64_t)readSeq.size()); }
        const std::vector<int> chain = chain_hits(hits, recs);
        merge_chain(hits, chain, recs, refSeq, readSeq, out[g]);
    });
    // chained.sort(): (reference id, pos, name, flag)
    std::vector<size_t> ord(out.size());
    for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
        const Rec &x = out[a].rec, &y = out[b].rec;
        if (x.rname != y.rname) return x.rname < y.rname;
        if (x.qname != y.qname) return x.qname < y.qname;
        return x.flag < y.flag;
    });
    std::vector<Rec> fresh;
    fresh.reserve(out.size());
    for (size_t i : ord) {
        Chained &c = out[i];
        io->arena.push_back(std::string(c.rec.qname));            // names outlive the input buffer's reuse
        c.rec.qname = io->arena.back();
        io->arena.push_back(std::move(c.seq));
        c.rec.seq = io->arena.back();
        io->arena.push_back(std::move(c.cigar));
        c.rec.cigar = io->arena.back();
        fresh.push_back(c.rec);
    }
    io->recs.swap(fresh);
    io->have_batch = false;
}

// ---------------------------------------------------------------------------------------------------------
// packing (packAlignedReads in nanopore_b200/realign.py; utils.py:168-180,570)
// ---------------------------------------------------------------------------------------------------------
void pack_reference(phmm_io *io) {
    if (!io->ref_offset.empty() || io->ref.names.empty()) return;
    int64_t total = 0;
    io->ref_offset.resize(io->ref.names.size() + 1);
    for (size_t i = 0; i < io->ref.names.size(); i++) { io->ref_offset[i] = total; total += (int64_t)io->ref.seq[i].size(); }
    io->ref_offset[io->ref.names.size()] = total;
    io->ref_codes.resize((size_t)total);
    parallel_for((int64_t)io->ref.names.size(), io->threads, [&](int64_t i) {
        const sv s = io->ref.seq[(size_t)i];
        uint8_t *dst = io->ref_codes.data() + io->ref_offset[(size_t)i];
        for (size_t k = 0; k < s.size(); k++) dst[k] = base_code(s[k]);
    });
}

void build_batch(phmm_io *io) {
    if (io->have_batch) return;
    if (!io->have_records) fail(PHMM_IO_E_STATE, "no SAM records loaded");
    if (!io->ref.loaded) fail(PHMM_IO_E_STATE, "no reference loaded");
    pack_reference(io);
    io->mapped.clear();
    for (int64_t k = 0; k < (int64_t)io->recs.size(); k++) if (io->recs[k].rname != -1) io->mapped.push_back(k);
    const int64_t n = (int64_t)io->mapped.size();
    struct Item { std::vector<uint32_t> ops; int64_t qstart = 0, qend = 0, rs = 0, re = 0; };
    std::vector<Item> items((size_t)n);
    parallel_for(n, io->threads, [&](int64_t i) {
        const Rec &r = io->recs[(size_t)io->mapped[(size_t)i]];
        const std::string &rn = io->references[(size_t)r.rname];
        auto fit = io->ref.index.find(rn);
        if (fit == io->ref.index.end())
            fail(PHMM_IO_E_FORMAT, "Reference sequence " + rn + " of read " + std::string(r.qname) + " not in the reference fasta");
        std::vector<CigOp> ops;
        parse_cigar(r.cigar, ops);
        Item &it = items[(size_t)i];
        int64_t alen = 0;
        for (const CigOp &o : ops) {
            if (!(o.code <= 2 || o.code == 4 || o.code == 5))
                fail(PHMM_IO_E_FORMAT, "cigar of " + std::string(r.qname) + " holds an operation other than M, I, D, S, H");
            if (o.code == 0 || o.code == 2) alen += o.len;
            if (o.code <= 2 && o.len > 0) {
                if (o.len > 0x3fffffff) fail(PHMM_IO_E_FORMAT, "cigar operation of " + std::string(r.qname) + " too long");
                if (!it.ops.empty() && (it.ops.back() & 3u) == o.code) it.ops.back() += (uint32_t)o.len << 2;   // pack_ops merges neighbours
                else it.ops.push_back(((uint32_t)o.len << 2) | o.code);
            }
        }
        // soft clips bound the query; hard clips are not in SEQ
        std::vector<CigOp> nh;
        for (const CigOp &o : ops) if (o.code != 5) nh.push_back(o);
        const int64_t ls = (int64_t)r.seq.size();
        int64_t qs = (!nh.empty() && nh.front().code == 4) ? nh.front().len : 0;
        int64_t qe = ls - ((nh.size() > 1 && nh.back().code == 4) ? nh.back().len : 0);
        qs = std::max<int64_t>(0, std::min(qs, ls));                                // Python slice semantics
        qe = std::max<int64_t>(0, std::min(qe, ls));
        if (qe < qs) qe = qs;
        it.qstart = qs; it.qend = qe;
        const int64_t len = (int64_t)io->ref.seq[(size_t)fit->second].size();
        if (r.pos < 0 || r.pos + alen > len) fail(PHMM_IO_E_FORMAT, "Alignment of " + std::string(r.qname) + " runs past the end of " + rn);
        it.rs = io->ref_offset[(size_t)fit->second] + r.pos;
        it.re = it.rs + alen;
    });
    io->b_read_off.assign((size_t)n + 1, 0); io->b_ops_off.assign((size_t)n + 1, 0);
    io->b_ref_start.resize((size_t)n); io->b_ref_end.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        io->b_read_off[(size_t)i + 1] = io->b_read_off[(size_t)i] + (items[(size_t)i].qend - items[(size_t)i].qstart);
        io->b_ops_off[(size_t)i + 1] = io->b_ops_off[(size_t)i] + (int64_t)items[(size_t)i].ops.size();
        io->b_ref_start[(size_t)i] = items[(size_t)i].rs; io->b_ref_end[(size_t)i] = items[(size_t)i].re;
    }
    io->b_reads.resize((size_t)io->b_read_off[(size_t)n]);
    io->b_ops.resize((size_t)io->b_ops_off[(size_t)n]);
    parallel_for(n, io->threads, [&](int64_t i) {
        const Rec &r = io->recs[(size_t)io->mapped[(size_t)i]];
        const Item &it = items[(size_t)i];
        uint8_t *dst = io->b_reads.data() + io->b_read_off[(size_t)i];
        for (int64_t k = it.qstart; k < it.qend; k++) dst[k - it.qstart] = base_code(r.seq[(size_t)k]);
        if (!it.ops.empty()) memcpy(io->b_ops.data() + io->b_ops_off[(size_t)i], it.ops.data(), it.ops.size() * 4);
    });
    io->have_batch = true;
}

void write_records(phmm_io *io, const char *path, const std::vector<int64_t> &which, const uint32_t *ops, const int64_t *off) {
    const int64_t n = (int64_t)which.size();
    const int64_t chunk = 64;
    std::vector<std::string> chunks((size_t)((n + chunk - 1) / chunk));
    parallel_for((int64_t)chunks.size(), io->threads, [&](int64_t c) {
        std::string &s = chunks[(size_t)c];
        std::string cg;
        for (int64_t k = c * chunk; k < std::min(n, (c + 1) * chunk); k++) {
            const Rec &r = io->recs[(size_t)which[(size_t)k]];
            if (ops) {
                cg.clear();
                for (int64_t j = off[k]; j < off[k + 1]; j++) { append_int(cg, (int64_t)(ops[j] >> 2)); cg.push_back(OPS[ops[j] & 3u]); }
                format_record(io, r, cg, s);
            } else format_record(io, r, r.cigar, s);
        }
    });
    write_chunks(path, io->header_lines, chunks);
}

template <typename F>
int guarded(phmm_io *io, F f) {
    if (!io) return PHMM_IO_E_ARG;
    try { f(); io->err.clear(); return PHMM_IO_OK; }
    catch (const IoError &e) { io->err = e.msg; return e.code; }
    catch (const std::exception &e) { io->err = e.what(); return PHMM_IO_E_ARG; }
}

}  // namespace

extern "C" {

int phmm_io_version(void) { return 1; }

// Estimated DP cells per read from the guide cigars alone (see phmm_io.h).  Same formula as batch.estimate_cells.
int phmm_io_estimate_cells(int64_t n, const uint32_t *ops, const int64_t *ops_off, const int64_t *read_off, const int64_t *ref_start,
                           const int64_t *ref_end, int band, int anchor_trim, int64_t split_side, int threads, int64_t *out_cells) {
    if (n < 0 || (n > 0 && (!ops_off || !read_off || !ref_start || !ref_end || !out_cells))) return PHMM_IO_E_ARG;
    int hw = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = hw > 0 ? hw : 1;
    const double e = (double)band, side2 = (double)split_side * (double)split_side;
    auto block = [&](double dx, double dy) {
        if (dx < 0) dx = 0;
        if (dy < 0) dy = 0;
        if (dx * dy > side2) {
            const double hx = std::min(std::floor(dx / 2), (double)split_side), hy = std::min(std::floor(dy / 2), (double)split_side);
            return 2.0 * (hx + e + 1) * (hy + e + 1);
        }
        return (dx + e + 1) * (dy + e + 1);
    };
    const int64_t chunk = 256;
    try {
        parallel_for((n + chunk - 1) / chunk, threads, [&](int64_t c) {
            for (int64_t i = c * chunk; i < std::min(n, (c + 1) * chunk); i++) {
                const int64_t lX = ref_end[i] - ref_start[i], lY = read_off[i + 1] - read_off[i];
                double cost = 0.0;
                int64_t x = 0, y = 0, px = 0, py = 0;
                for (int64_t k = ops_off[i]; k < ops_off[i + 1]; k++) {
                    const int64_t len = ops[k] >> 2;
                    const int code = (int)(ops[k] & 3u);
                    if (code == 0) {
                        if (len > 2 * (int64_t)anchor_trim) {
                            const int64_t an = len - 2 * anchor_trim;
                            cost += 2.0 * (double)an * (e + 1);
                            cost += block((double)(x + anchor_trim - px), (double)(y + anchor_trim - py));
                            px = x + anchor_trim + an; py = y + anchor_trim + an;
                        }
                        x += len; y += len;
                    } else if (code == 1) y += len;
                    else if (code == 2) x += len;
                }
                cost += block((double)(lX - px), (double)(lY - py));
                out_cells[i] = cost < 1.0 ? 1 : (int64_t)cost;
            }
        });
    } catch (...) { return PHMM_IO_E_ARG; }
    return PHMM_IO_OK;
}

// out[out_off[k] .. out_off[k+1]) = data[off[idx[k]] .. off[idx[k]+1]) (elements of elem_size bytes), threaded.
int phmm_io_gather_ranges(const void *data, const int64_t *off, const int64_t *idx, int64_t n_idx, int elem_size, int threads,
                          void *out, const int64_t *out_off) {
    if (n_idx < 0 || elem_size < 1 || (n_idx > 0 && (!data || !off || !idx || !out || !out_off))) return PHMM_IO_E_ARG;
    int hw = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = hw > 0 ? hw : 1;
    const int64_t chunk = 128;
    try {
        parallel_for((n_idx + chunk - 1) / chunk, threads, [&](int64_t c) {
            for (int64_t k = c * chunk; k < std::min(n_idx, (c + 1) * chunk); k++) {
                const int64_t a = off[idx[k]], b = off[idx[k] + 1];
                if (b > a) memcpy((char *)out + out_off[k] * elem_size, (const char *)data + a * elem_size, (size_t)(b - a) * elem_size);
            }
        });
    } catch (...) { return PHMM_IO_E_ARG; }
    return PHMM_IO_OK;
}

phmm_io *phmm_io_create(int threads) {
    phmm_io *io = new (std::nothrow) phmm_io();
    if (!io) return nullptr;
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    io->threads = threads > 0 ? threads : hw;
    return io;
}

void phmm_io_destroy(phmm_io *io) { delete io; }

const char *phmm_io_last_error(phmm_io *io) { return io ? io->err.c_str() : "null handle"; }

int phmm_io_load_reference(phmm_io *io, const char *fasta_path) {
    return guarded(io, [&] {
        if (!fasta_path) fail(PHMM_IO_E_ARG, "fasta_path is NULL");
        load_fasta(io->ref, fasta_path);
        io->ref_codes.clear(); io->ref_offset.clear(); io->have_batch = false;
    });
}

int phmm_io_load_reads(phmm_io *io, const char *fastq_path) {
    return guarded(io, [&] {
        if (!fastq_path) fail(PHMM_IO_E_ARG, "fastq_path is NULL");
        load_fastq(io->reads, fastq_path);
    });
}

int phmm_io_chain_sam(phmm_io *io, const char *sam_path) {
    return guarded(io, [&] {
        if (!sam_path) fail(PHMM_IO_E_ARG, "sam_path is NULL");
        chain_sam(io, sam_path);
    });
}

int phmm_io_load_sam(phmm_io *io, const char *sam_path) {
    return guarded(io, [&] {
        if (!sam_path) fail(PHMM_IO_E_ARG, "sam_path is NULL");
        read_sam(io, sam_path);
    });
}

int phmm_io_counts(phmm_io *io, int64_t *n_records, int64_t *n_mapped) {
    return guarded(io, [&] {
        if (!io->have_records) fail(PHMM_IO_E_STATE, "no SAM records loaded");
        int64_t m = 0;
        for (const Rec &r : io->recs) m += r.rname != -1;
        if (n_records) *n_records = (int64_t)io->recs.size();
        if (n_mapped) *n_mapped = m;
    });
}

int phmm_io_batch_view(phmm_io *io, phmm_io_batch *out) {
    return guarded(io, [&] {
        if (!out) fail(PHMM_IO_E_ARG, "out is NULL");
        build_batch(io);
        out->n_reads = (int64_t)io->mapped.size();
        out->ref = io->ref_codes.data(); out->ref_len = (int64_t)io->ref_codes.size();
        out->reads = io->b_reads.data(); out->read_off = io->b_read_off.data();
        out->ref_start = io->b_ref_start.data(); out->ref_end = io->b_ref_end.data();
        out->ops = io->b_ops.data(); out->ops_off = io->b_ops_off.data();
    });
}

int phmm_io_write_sam(phmm_io *io, const char *out_path) {
    return guarded(io, [&] {
        if (!out_path) fail(PHMM_IO_E_ARG, "out_path is NULL");
        if (!io->have_records) fail(PHMM_IO_E_STATE, "no SAM records loaded");
        std::vector<int64_t> all(io->recs.size());
        for (size_t i = 0; i < all.size(); i++) all[i] = (int64_t)i;
        write_records(io, out_path, all, nullptr, nullptr);
    });
}

int phmm_io_write_realigned_sam(phmm_io *io, const char *out_path, const uint32_t *ops, const int64_t *off, int64_t n) {
    return guarded(io, [&] {
        if (!out_path || !off || (!ops && n > 0 && off[n] > 0)) fail(PHMM_IO_E_ARG, "NULL argument");
        if (!io->have_records) fail(PHMM_IO_E_STATE, "no SAM records loaded");
        std::vector<int64_t> mapped;
        for (int64_t k = 0; k < (int64_t)io->recs.size(); k++) if (io->recs[(size_t)k].rname != -1) mapped.push_back(k);
        if ((int64_t)mapped.size() != n)
            fail(PHMM_IO_E_ARG, "got " + std::to_string(n) + " cigars for " + std::to_string(mapped.size()) + " mapped records (utils.py:588-589)");
        static const uint32_t none = 0;
        write_records(io, out_path, mapped, ops ? ops : &none, off);
    });
}

}  // extern "C"

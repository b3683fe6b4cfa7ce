// Native FASTA loader: the data format on the input side of the screen.  The reference reads subject genomes with
// Biopython / pyfaidx (fasta.py, MINORg.py:2119 `IndexedFasta(ifasta)`); the Python mirror's numpy reader takes 0.7 s per
// 135 Mb genome, 7x the time the GPU needs to screen 10 000 gRNA against it.  This reader maps the file and strips it
// with a pool of host threads (three passes over the bytes: header scan, count, compacting copy) into one buffer that
// mg_genome_from_ascii / mg_genome_from_fasta hands to the device.
//
// Record semantics (identical to minorg_b200.fasta.read_fasta_arrays, checked byte for byte in tests/test_host_mirror.py):
// a header line is a line whose first byte is '>'; the record id is the first whitespace-delimited token of the header;
// the sequence is every byte up to the next header line except '\n', '\r', ' ' and '\t', case preserved; bytes before
// the first header are ignored.  gzip files (magic 1f 8b; the subjects MINORg users keep are often .fasta.gz) are
// inflated with zlib first.  mg_genome_from_fasta refuses a file without a single record: a subject with no contigs
// would let every gRNA pass the background check.
#include <fcntl.h>
#include <zlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

struct mg_fasta {
    std::vector<std::string> names;
    std::vector<int64_t> offsets;     // n + 1
    char* seq = nullptr;              // offsets[n] bytes (malloc)
};

namespace mg {

static inline bool dropped(unsigned char c) { return c == '\n' || c == '\r' || c == ' ' || c == '\t'; }

template <class F>
static void parallel_for(int n_items, int threads, F f) {
    if (n_items <= 0) return;
    threads = std::max(1, std::min(threads, n_items));
    std::atomic<int> next(0);
    auto work = [&] { for (int i = next.fetch_add(1); i < n_items; i = next.fetch_add(1)) f(i); };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
}

static int fasta_parse(const char* buf, int64_t size, int threads, mg_fasta* out) {
    constexpr int64_t kChunk = 4 << 20;
    // pass 1: header line starts ('>' at the start of a line)
    const int n_chunks = (int)((size + kChunk - 1) / kChunk);
    std::vector<std::vector<int64_t>> found(n_chunks);
    parallel_for(n_chunks, threads, [&](int c) {
        const int64_t a = (int64_t)c * kChunk, b = std::min(size, a + kChunk);
        for (const char* p = (const char*)memchr(buf + a, '>', (size_t)(b - a)); p; ) {
            const int64_t pos = p - buf;
            if (pos == 0 || buf[pos - 1] == '\n') found[c].push_back(pos);
            const int64_t rest = b - (pos + 1);
            p = rest > 0 ? (const char*)memchr(p + 1, '>', (size_t)rest) : nullptr;
        }
    });
    std::vector<int64_t> headers;
    for (auto& v : found) headers.insert(headers.end(), v.begin(), v.end());
    const int n = (int)headers.size();
    out->names.resize(n);
    out->offsets.assign(n + 1, 0);
    // bodies, cut into work items
    struct Item { int rec; int64_t a, b, kept, dst; };
    std::vector<Item> items;
    std::vector<int64_t> body_begin(n), body_end(n);
    for (int r = 0; r < n; ++r) {
        const int64_t h = headers[r], nxt = r + 1 < n ? headers[r + 1] : size;
        const char* nl = (const char*)memchr(buf + h, '\n', (size_t)(nxt - h));
        const int64_t eol = nl ? nl - buf : nxt;
        // id = first whitespace-delimited token of the header text
        int64_t i = h + 1;
        while (i < eol && isspace((unsigned char)buf[i])) ++i;
        int64_t j = i;
        while (j < eol && !isspace((unsigned char)buf[j])) ++j;
        out->names[r].assign(buf + i, (size_t)(j - i));
        body_begin[r] = eol;
        body_end[r] = nxt;
        for (int64_t a = eol; a < nxt || a == eol; a += kChunk) {
            items.push_back(Item{r, a, std::min(nxt, a + kChunk), 0, 0});
            if (a + kChunk >= nxt) break;
        }
    }
    // pass 2: bytes kept per item
    parallel_for((int)items.size(), threads, [&](int k) {
        Item& it = items[k];
        int64_t cnt = 0;
        const unsigned char* p = (const unsigned char*)buf + it.a;
        const int64_t len = it.b - it.a;
        for (int64_t i = 0; i < len; ++i) cnt += !dropped(p[i]);
        it.kept = cnt;
    });
    int64_t total = 0;
    for (auto& it : items) {
        it.dst = total;
        total += it.kept;
        out->offsets[it.rec + 1] = total;
    }
    for (int r = 0; r < n; ++r) out->offsets[r + 1] = std::max(out->offsets[r + 1], out->offsets[r]);   // empty records
    out->seq = (char*)malloc((size_t)std::max<int64_t>(total, 1));
    if (!out->seq) { set_error("out of memory reading FASTA (%lld bytes)", (long long)total); return MG_ERR_NOMEM; }
    // pass 3: compacting copy
    char* dst = out->seq;
    parallel_for((int)items.size(), threads, [&](int k) {
        const Item& it = items[k];
        const unsigned char* p = (const unsigned char*)buf + it.a;
        char* o = dst + it.dst;
        const int64_t len = it.b - it.a;
        int64_t i = 0;
        while (i < len) {
            // copy whole lines: most FASTA bodies hold nothing to drop but the line feeds
            const unsigned char* nl = (const unsigned char*)memchr(p + i, '\n', (size_t)(len - i));
            const int64_t e = nl ? nl - p : len;
            bool clean = true;
            for (int64_t j = i; j < e; ++j) clean &= !dropped(p[j]);
            if (clean) { memcpy(o, p + i, (size_t)(e - i)); o += e - i; }
            else for (int64_t j = i; j < e; ++j) { *o = (char)p[j]; o += !dropped(p[j]); }
            i = e + 1;
        }
    });
    return MG_OK;
}

}  // namespace mg

using namespace mg;

extern "C" {

void mg_fasta_free(mg_fasta* f) {
    if (!f) return;
    free(f->seq);
    delete f;
}

int mg_fasta_read(const char* path, int threads, mg_fasta** out) {
    MG_CHECK_ARG(path != nullptr && out != nullptr);
    if (threads <= 0) threads = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { set_error("cannot open %s: %s", path, strerror(errno)); return MG_ERR_ARG; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); set_error("cannot stat %s: %s", path, strerror(errno)); return MG_ERR_ARG; }
    const int64_t size = (int64_t)st.st_size;
    mg_fasta* f = new mg_fasta();
    int rc = MG_OK;
    if (size == 0) {
        f->offsets.assign(1, 0);
        f->seq = (char*)malloc(1);
    } else {
        void* map = mmap(nullptr, (size_t)size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (map == MAP_FAILED) { close(fd); delete f; set_error("cannot map %s: %s", path, strerror(errno)); return MG_ERR_ARG; }
        madvise(map, (size_t)size, MADV_SEQUENTIAL);
        const unsigned char* m = (const unsigned char*)map;
        if (size >= 2 && m[0] == 0x1f && m[1] == 0x8b) {
            // gzip: inflate the whole file (multi-member files included) into a host buffer, then parse that
            std::vector<char> plain;
            gzFile gz = gzopen(path, "rb");
            if (!gz) { rc = MG_ERR_ARG; set_error("cannot open %s as gzip", path); }
            else {
                gzbuffer(gz, 1 << 20);
                size_t used = 0;
                plain.resize((size_t)std::max<int64_t>(size * 4, 1 << 20));
                for (;;) {
                    if (plain.size() - used < (1u << 20)) plain.resize(plain.size() * 2);
                    const int got = gzread(gz, plain.data() + used, (unsigned)std::min<size_t>(plain.size() - used, 1u << 30));
                    if (got < 0) { int en = 0; set_error("gzip error in %s: %s", path, gzerror(gz, &en)); rc = MG_ERR_ARG; break; }
                    if (got == 0) break;
                    used += (size_t)got;
                }
                gzclose(gz);
                if (rc == MG_OK) rc = fasta_parse(plain.data(), (int64_t)used, threads, f);
            }
        } else {
            rc = fasta_parse((const char*)map, size, threads, f);
        }
        munmap(map, (size_t)size);
    }
    close(fd);
    if (rc != MG_OK) { mg_fasta_free(f); return rc; }
    *out = f;
    return MG_OK;
}

int mg_fasta_n_records(const mg_fasta* f) { return f ? (int)f->names.size() : 0; }
const char* mg_fasta_name(const mg_fasta* f, int i) { return (f && i >= 0 && i < (int)f->names.size()) ? f->names[i].c_str() : ""; }
const int64_t* mg_fasta_offsets(const mg_fasta* f) { return f ? f->offsets.data() : nullptr; }
const char* mg_fasta_sequence(const mg_fasta* f) { return f ? f->seq : nullptr; }

int mg_genome_from_fasta(const char* path, int threads, void* stream, mg_genome** out, mg_fasta** fasta_out) {
    MG_CHECK_ARG(out != nullptr);
    mg_fasta* f = nullptr;
    int rc = mg_fasta_read(path, threads, &f);
    if (rc != MG_OK) return rc;
    if (f->names.empty()) {      // a subject without contigs would let every gRNA pass the background check
        mg_fasta_free(f);
        set_error("%s holds no FASTA record (no line starts with '>'): not a FASTA file?", path);
        return MG_ERR_ARG;
    }
    rc = mg_genome_from_ascii(f->seq, f->offsets.data(), (int)f->names.size(), stream, out);
    if (rc != MG_OK || fasta_out == nullptr) mg_fasta_free(f);
    else *fasta_out = f;
    return rc;
}

}  // extern "C"

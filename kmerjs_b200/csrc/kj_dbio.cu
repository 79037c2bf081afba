// kj_dbio.cu -- template databases from disk into GPU memory (kj_db_load / kj_db_save_packed): the step before the
// path.  The reference keeps the k-mer -> template-list store in MongoDB / Redis and fills it from JSON; the layouts
// accepted here are the reference's own (paths relative to the kmerjs repository):
//
//   KJ_DB_KMER_DOCS       [{"kmer": K, "templates": [{"sequence","lengths","ulengths","species"}, ..]}, ..]
//                         lib/kmerFinderServer.js:68-92; the Redis lists of :184-199 hold the same records as strings
//   KJ_DB_TEMPLATE_DOCS   [{"sequence","lengths","ulenght","species","reads":[K, ..]}, ..]
//                         src/kmerPyToMongo.py:35-42 (the field really is spelled `ulenght`)
//   KJ_DB_KMERFINDER_MAP  {K: "T1,T2,.."}  lib/index.js:184-192, src/kmerPyToMongo.py:15-24, with optional side tables
//                         <path>.lengths.json / .ulengths.json / .descriptions.json ({template: value})
//   KJ_DB_PACKED          the versioned binary cache written by kj_db_save_packed (and TemplateDB.save): plain arrays
//
// The Summary record {"templates","uniqueLens","totalLen"} (lib/kmerFinderServer.js:716-724, test_data/summary.json)
// comes from `summary_path`, or is derived from the templates when that is NULL.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include "kj_internal.hpp"

namespace {

// ---------------------------------------------------------------------------------------- a small JSON reader
struct JVal {
    enum Type { NUL, BOOL, NUM, STR, ARR, OBJ } type = NUL;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;       // insertion order kept: it is the DB order
    const JVal *get(const char *key) const {
        for (const auto &kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char *p, *end;
    std::string err;
    explicit JParser(const std::string &s) : p(s.data()), end(s.data() + s.size()) {}
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool fail(const char *what) { if (err.empty()) err = what; return false; }
    bool string(std::string &out) {
        if (p >= end || *p != '"') return fail("expected a string");
        ++p;
        out.clear();
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail("bad escape");
                switch (*p) {
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    case '/': out += '/'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'n': out += '\n'; break;
                    case 'r': out += '\r'; break;
                    case 't': out += '\t'; break;
                    case 'u': {
                        if (end - p < 5) return fail("bad \\u escape");
                        unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                        p += 4;
                        if (cp < 0x80) out += (char)cp;                       // UTF-8
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: return fail("bad escape");
                }
                ++p;
            } else {
                out += *p++;
            }
        }
        if (p >= end) return fail("unterminated string");
        ++p;
        return true;
    }
    bool value(JVal &v, int depth = 0) {
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end");
        if (*p == '{') {
            v.type = JVal::OBJ;
            ++p; ws();
            if (p < end && *p == '}') { ++p; return true; }
            for (;;) {
                ws();
                std::string key;
                if (!string(key)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':'");
                ++p;
                v.obj.emplace_back(std::move(key), JVal());
                if (!value(v.obj.back().second, depth + 1)) return false;
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v.type = JVal::ARR;
            ++p; ws();
            if (p < end && *p == ']') { ++p; return true; }
            for (;;) {
                v.arr.emplace_back();
                if (!value(v.arr.back(), depth + 1)) return false;
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (*p == '"') { v.type = JVal::STR; return string(v.str); }
        if (end - p >= 4 && !strncmp(p, "true", 4)) { v.type = JVal::BOOL; v.b = true; p += 4; return true; }
        if (end - p >= 5 && !strncmp(p, "false", 5)) { v.type = JVal::BOOL; v.b = false; p += 5; return true; }
        if (end - p >= 4 && !strncmp(p, "null", 4)) { v.type = JVal::NUL; p += 4; return true; }
        char *q = nullptr;
        errno = 0;
        v.num = strtod(p, &q);
        if (q == p || q > end) return fail("bad value");
        v.type = JVal::NUM;
        p = q;
        return true;
    }
};

bool read_file(const std::string &path, std::string &out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

bool parse_json_file(const std::string &path, JVal &v, std::string &err) {
    std::string text;
    if (!read_file(path, text)) { err = "cannot read " + path; return false; }
    JParser ps(text);
    if (!ps.value(v)) { err = path + ": " + ps.err; return false; }
    ps.ws();
    if (ps.p != ps.end) { err = path + ": trailing characters"; return false; }
    return true;
}

uint64_t as_u64(const JVal *v) {
    if (!v) return 0;
    if (v->type == JVal::NUM) return v->num > 0 ? (uint64_t)(v->num + 0.5) : 0;
    if (v->type == JVal::STR) return strtoull(v->str.c_str(), nullptr, 10);
    return 0;
}

// ---------------------------------------------------------------------------------------- host description of a DB
struct HostDb {
    std::vector<uint8_t> kmer_bytes;
    std::vector<uint32_t> kmer_len;
    std::vector<uint64_t> list_off{0};
    std::vector<uint32_t> tmpl_ids;
    std::vector<std::string> names, species;
    std::vector<uint64_t> lengths, ulengths;
    uint64_t s_templates = 0, s_unique_lens = 0, s_total_len = 0;
    bool have_summary = false;
    std::map<std::string, uint32_t> tid;

    uint32_t template_id(const std::string &name, uint64_t len, uint64_t ulen, const std::string &sp, bool update) {
        auto it = tid.find(name);
        if (it != tid.end()) {
            if (update) { lengths[it->second] = len; ulengths[it->second] = ulen; species[it->second] = sp; }
            return it->second;
        }
        const uint32_t id = (uint32_t)names.size();
        tid.emplace(name, id);
        names.push_back(name); lengths.push_back(len); ulengths.push_back(ulen); species.push_back(sp);
        return id;
    }
    void add_kmer(const std::string &k) {
        kmer_bytes.insert(kmer_bytes.end(), k.begin(), k.end());
        kmer_len.push_back((uint32_t)k.size());
    }
    void close_list() { list_off.push_back(tmpl_ids.size()); }
    void derive_summary() {
        if (have_summary) return;
        s_templates = names.size();
        s_unique_lens = s_total_len = 0;
        for (size_t i = 0; i < names.size(); ++i) { s_unique_lens += ulengths[i]; s_total_len += lengths[i]; }
    }
};

bool load_kmer_docs(const JVal &doc, HostDb &db, std::string &err) {
    for (const JVal &d : doc.arr) {
        const JVal *k = d.get("kmer"), *ts = d.get("templates");
        if (!k || k->type != JVal::STR || !ts || ts->type != JVal::ARR) { err = "per-k-mer document without kmer / templates"; return false; }
        db.add_kmer(k->str);
        for (const JVal &t0 : ts->arr) {
            JVal parsed;
            const JVal *t = &t0;
            if (t0.type == JVal::STR) {                  // Redis list entries are JSON strings (lib/kmerFinderServer.js:184-186)
                JParser ps(t0.str);
                if (!ps.value(parsed)) { err = "bad template record string"; return false; }
                t = &parsed;
            }
            const JVal *seq = t->get("sequence");
            if (!seq || seq->type != JVal::STR) { err = "template record without sequence"; return false; }
            const JVal *sp = t->get("species");
            db.tmpl_ids.push_back(db.template_id(seq->str, as_u64(t->get("lengths")), as_u64(t->get("ulengths")),
                                                 sp && sp->type == JVal::STR ? sp->str : std::string(), false));
        }
        db.close_list();
    }
    return true;
}

bool load_template_docs(const JVal &doc, HostDb &db, std::string &err) {
    // k-mer lists in document order (what the Mongo unwind / group of lib/kmerFinderServer.js:70-92 yields for an ordered collection)
    std::map<std::string, uint32_t> kid;
    std::vector<std::vector<uint32_t>> lists;
    std::vector<std::string> kmers;
    for (const JVal &d : doc.arr) {
        const JVal *seq = d.get("sequence");
        if (!seq || seq->type != JVal::STR) { err = "per-template document without sequence"; return false; }
        const JVal *ul = d.get("ulenght") ? d.get("ulenght") : d.get("ulengths");
        const JVal *sp = d.get("species");
        const uint32_t t = db.template_id(seq->str, as_u64(d.get("lengths")), as_u64(ul), sp && sp->type == JVal::STR ? sp->str : std::string(), true);
        const JVal *reads = d.get("reads");
        if (reads && reads->type == JVal::ARR)
            for (const JVal &r : reads->arr) {
                if (r.type != JVal::STR) continue;
                auto it = kid.find(r.str);
                if (it == kid.end()) { it = kid.emplace(r.str, (uint32_t)kmers.size()).first; kmers.push_back(r.str); lists.emplace_back(); }
                lists[it->second].push_back(t);
            }
    }
    for (size_t i = 0; i < kmers.size(); ++i) {
        db.add_kmer(kmers[i]);
        db.tmpl_ids.insert(db.tmpl_ids.end(), lists[i].begin(), lists[i].end());
        db.close_list();
    }
    return true;
}

bool load_kmerfinder_map(const JVal &doc, const std::string &path, HostDb &db, std::string &err) {
    JVal lens, ulens, descr;
    std::string e2;
    const bool hl = parse_json_file(path + ".lengths.json", lens, e2), hu = parse_json_file(path + ".ulengths.json", ulens, e2),
               hd = parse_json_file(path + ".descriptions.json", descr, e2);
    for (const auto &kv : doc.obj) {
        if (kv.second.type != JVal::STR) { err = "KmerFinder map value is not a string"; return false; }
        db.add_kmer(kv.first);
        const std::string &csv = kv.second.str;
        size_t a = 0;
        while (a <= csv.size()) {
            size_t b = csv.find(',', a);
            if (b == std::string::npos) b = csv.size();
            if (b > a) {
                const std::string name = csv.substr(a, b - a);
                const JVal *sp = hd ? descr.get(name.c_str()) : nullptr;
                db.tmpl_ids.push_back(db.template_id(name, hl ? as_u64(lens.get(name.c_str())) : 0,
                                                     hu ? as_u64(ulens.get(name.c_str())) : 0,
                                                     sp && sp->type == JVal::STR ? sp->str : std::string(), false));
            }
            a = b + 1;
        }
        db.close_list();
    }
    return true;
}

// ---------------------------------------------------------------------------------------- packed binary
const char KJ_PACKED_MAGIC[8] = {'K', 'J', 'D', 'B', 'v', '0', '0', '1'};
struct PackedHeader {
    char magic[8];
    uint64_t n_kmers, kmer_bytes, n_pairs, n_templates, names_bytes, species_bytes;
    uint64_t s_templates, s_unique_lens, s_total_len;
};

template <class T>
bool wr(FILE *f, const T *p, size_t n) { return n == 0 || fwrite(p, sizeof(T), n, f) == n; }
template <class T>
bool rd(FILE *f, std::vector<T> &v, size_t n) { v.resize(n); return n == 0 || fread(v.data(), sizeof(T), n, f) == n; }

std::string join0(const std::vector<std::string> &v) {
    std::string s;
    for (const auto &x : v) { s += x; s += '\0'; }
    return s;
}
void split0(const std::vector<char> &blob, std::vector<std::string> &out, size_t n) {
    out.clear();
    size_t a = 0;
    for (size_t i = 0; i < blob.size() && out.size() < n; ++i)
        if (blob[i] == '\0') { out.emplace_back(blob.data() + a, i - a); a = i + 1; }
    out.resize(n);
}

bool load_packed(const std::string &path, HostDb &db, std::string &err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    PackedHeader h;
    bool ok = fread(&h, sizeof(h), 1, f) == 1 && !memcmp(h.magic, KJ_PACKED_MAGIC, 8);
    if (!ok) { fclose(f); err = path + ": not a kmerjs_b200 packed database (magic / version)"; return false; }
    std::vector<char> nb, sb;
    ok = rd(f, db.kmer_len, h.n_kmers) && rd(f, db.kmer_bytes, h.kmer_bytes) && rd(f, db.list_off, h.n_kmers + 1) &&
         rd(f, db.tmpl_ids, h.n_pairs) && rd(f, db.lengths, h.n_templates) && rd(f, db.ulengths, h.n_templates) &&
         rd(f, nb, h.names_bytes) && rd(f, sb, h.species_bytes);
    fclose(f);
    if (!ok) { err = path + ": truncated packed database"; return false; }
    uint64_t kb = 0;
    for (uint32_t l : db.kmer_len) kb += l;
    if (kb != h.kmer_bytes || db.list_off.front() != 0 || db.list_off.back() != h.n_pairs) { err = path + ": inconsistent packed database"; return false; }
    for (size_t i = 0; i + 1 < db.list_off.size(); ++i)
        if (db.list_off[i] > db.list_off[i + 1]) { err = path + ": inconsistent packed database"; return false; }
    split0(nb, db.names, h.n_templates);
    split0(sb, db.species, h.n_templates);
    db.s_templates = h.s_templates; db.s_unique_lens = h.s_unique_lens; db.s_total_len = h.s_total_len;
    db.have_summary = true;
    return true;
}

}  // namespace

extern "C" int kj_db_save_packed(const char *path, const kj_db_desc *d, const char *const *names, const char *const *species) {
    if (!path || !d) return kj_fail(nullptr, KJ_E_INVALID, "kj_db_save_packed: null argument");
    FILE *f = fopen(path, "wb");
    if (!f) return kj_fail(nullptr, KJ_E_IO, std::string("cannot create ") + path);
    std::vector<std::string> nv(d->n_templates), sv(d->n_templates);
    for (uint32_t i = 0; i < d->n_templates; ++i) { if (names && names[i]) nv[i] = names[i]; if (species && species[i]) sv[i] = species[i]; }
    const std::string nb = join0(nv), sb = join0(sv);
    PackedHeader h;
    memcpy(h.magic, KJ_PACKED_MAGIC, 8);
    h.n_kmers = d->n_kmers;
    h.kmer_bytes = 0;
    for (uint64_t i = 0; i < d->n_kmers; ++i) h.kmer_bytes += d->kmer_len[i];
    h.n_pairs = d->n_kmers ? d->list_off[d->n_kmers] : 0;
    h.n_templates = d->n_templates;
    h.names_bytes = nb.size(); h.species_bytes = sb.size();
    h.s_templates = d->summary_templates; h.s_unique_lens = d->summary_unique_lens; h.s_total_len = d->summary_total_len;
    const uint64_t zero = 0;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && wr(f, d->kmer_len, d->n_kmers) && wr(f, d->kmer_bytes, h.kmer_bytes) &&
              (d->n_kmers ? wr(f, d->list_off, d->n_kmers + 1) : wr(f, &zero, 1)) && wr(f, d->tmpl_ids, h.n_pairs) &&
              wr(f, d->lengths, d->n_templates) && wr(f, d->ulengths, d->n_templates) && wr(f, nb.data(), nb.size()) &&
              wr(f, sb.data(), sb.size());
    ok = (fclose(f) == 0) && ok;
    if (!ok) return kj_fail(nullptr, KJ_E_IO, std::string("write error on ") + path);
    return KJ_OK;
}

extern "C" int kj_db_load(kj_ctx *ctx, const char *path, int format, const char *summary_path, uint32_t part,
                          uint32_t n_parts, kj_db **out) {
    if (!ctx || !path || !out) return kj_fail(ctx, KJ_E_INVALID, "kj_db_load: null argument");
    HostDb db;
    std::string err;
    const std::string p(path);
    if (format == KJ_DB_AUTO) {
        FILE *f = fopen(path, "rb");
        if (!f) return kj_fail(ctx, KJ_E_IO, "cannot open " + p);
        char m[8] = {0};
        const size_t got = fread(m, 1, 8, f);
        fclose(f);
        if (got == 8 && !memcmp(m, KJ_PACKED_MAGIC, 4)) format = KJ_DB_PACKED;
    }
    if (format == KJ_DB_PACKED) {
        if (!load_packed(p, db, err)) return kj_fail(ctx, KJ_E_IO, err);
    } else {
        JVal doc;
        if (!parse_json_file(p, doc, err)) return kj_fail(ctx, KJ_E_IO, err);
        if (format == KJ_DB_AUTO) {
            if (doc.type == JVal::OBJ) format = KJ_DB_KMERFINDER_MAP;
            else if (doc.type == JVal::ARR && !doc.arr.empty() && doc.arr[0].get("reads")) format = KJ_DB_TEMPLATE_DOCS;
            else format = KJ_DB_KMER_DOCS;
        }
        bool ok = false;
        if (format == KJ_DB_KMER_DOCS && doc.type == JVal::ARR) ok = load_kmer_docs(doc, db, err);
        else if (format == KJ_DB_TEMPLATE_DOCS && doc.type == JVal::ARR) ok = load_template_docs(doc, db, err);
        else if (format == KJ_DB_KMERFINDER_MAP && doc.type == JVal::OBJ) ok = load_kmerfinder_map(doc, p, db, err);
        else err = "the document does not have the layout of the requested format";
        if (!ok) return kj_fail(ctx, KJ_E_INVALID, p + ": " + err);
    }
    if (summary_path) {
        JVal sdoc;
        if (!parse_json_file(summary_path, sdoc, err)) return kj_fail(ctx, KJ_E_IO, err);
        const JVal *s = sdoc.type == JVal::ARR && !sdoc.arr.empty() ? &sdoc.arr[0] : &sdoc;
        if (!s->get("templates") || !s->get("uniqueLens")) return kj_fail(ctx, KJ_E_INVALID, std::string(summary_path) + ": not a Summary record");
        db.s_templates = as_u64(s->get("templates")); db.s_unique_lens = as_u64(s->get("uniqueLens")); db.s_total_len = as_u64(s->get("totalLen"));
        db.have_summary = true;
    }
    db.derive_summary();
    kj_db_desc d{};
    d.n_kmers = db.kmer_len.size();
    d.kmer_bytes = db.kmer_bytes.data(); d.kmer_len = db.kmer_len.data();
    d.list_off = db.list_off.data(); d.tmpl_ids = db.tmpl_ids.data();
    d.n_templates = (uint32_t)db.names.size();
    d.lengths = db.lengths.data(); d.ulengths = db.ulengths.data();
    d.summary_templates = db.s_templates; d.summary_unique_lens = db.s_unique_lens; d.summary_total_len = db.s_total_len;
    d.part = part; d.n_parts = n_parts;
    kj_db *h = nullptr;
    int rc = kj_db_create(ctx, &d, &h);
    if (rc) return rc;
    h->names = db.names;
    h->species = db.species;
    *out = h;
    return KJ_OK;
}

// template attributes of a database (names and species only when it came through kj_db_load)
extern "C" int kj_db_template(const kj_db *db, uint32_t id, const char **name, const char **species, uint64_t *lengths,
                              uint64_t *ulength) {
    if (!db || id >= db->n_templates) return KJ_E_INVALID;
    if (name) *name = id < db->names.size() ? db->names[id].c_str() : "";
    if (species) *species = id < db->species.size() ? db->species[id].c_str() : "";
    if (lengths) *lengths = db->lengths[id];
    if (ulength) *ulength = db->ulengths[id];
    return KJ_OK;
}

extern "C" int kj_db_summary(const kj_db *db, uint64_t *templates, uint64_t *unique_lens, uint64_t *total_len) {
    if (!db) return KJ_E_INVALID;
    if (templates) *templates = db->s_templates;
    if (unique_lens) *unique_lens = db->s_unique_lens;
    if (total_len) *total_len = db->s_total_len;
    return KJ_OK;
}

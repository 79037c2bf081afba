// node/addon.cc -- N-API addon over libkmerjs_b200.so (include/kmerjs_b200.h).
//
// NOT compiled in this repository's build: neither the build image nor the GPU box ships Node.js or
// node_api.h.  It is the binding a kmerjs maintainer adds (INTEGRATION.md has the binding.gyp); the
// Python package under kmerjs_b200/ binds the same ABI through ctypes and is what the tests drive.
//
// Exposed to JS (node/index.js wraps them in the reference's classes):
//   init(device) -> ctx            countFile(ctx, path, prefix, k, step) -> Promise<counts>
//   countsExport(counts) -> {keys: string[], counts: number[], lines, bytesRead}
//   countsFromMap(ctx, keys, counts, prefix, k, step) -> counts
//   dbCreate(ctx, {...arrays, summary}) -> db
//   firstMatch(ctx, counts, db) -> match | throws 'No hits were found!'
//   matchScores(match) -> {uScore, tScore, order, hits}
//   wtaNext(match) -> row | null        (synchronous, like the generator's .next())
//   countsAlive(counts) -> Uint8Array   (to mirror kmerMap.delete on the JS Map)
#include <node_api.h>
#include <string.h>
#include <string>
#include <vector>
#include "../include/kmerjs_b200.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, nullptr, #call); return nullptr; } } while (0)

static napi_value throw_kj(napi_env env, kj_ctx *ctx, int rc) {
    const char *msg = kj_last_error(ctx);
    napi_throw_error(env, nullptr, msg && *msg ? msg : "kmerjs_b200 error");
    (void)rc;
    return nullptr;
}

template <class T> static T *unwrap(napi_env env, napi_value v) {
    void *p = nullptr;
    napi_get_value_external(env, v, &p);
    return static_cast<T *>(p);
}

static napi_value Init(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    int32_t device = 0;
    if (argc) napi_get_value_int32(env, argv[0], &device);
    kj_ctx *ctx = nullptr;
    int rc = kj_init(device, nullptr, &ctx);
    if (rc) return throw_kj(env, nullptr, rc);          // KJ_E_NO_SM100: there is no CPU fallback
    napi_value out;
    NAPI_OK(napi_create_external(env, ctx, [](napi_env, void *p, void *) { kj_destroy((kj_ctx *)p); }, nullptr, &out));
    return out;
}

// ---- countFile: async work on the libuv pool, resolves an external holding kj_counts* ----------
struct CountJob {
    napi_async_work work; napi_deferred deferred;
    kj_ctx *ctx; std::string path, prefix; uint32_t k, step;
    kj_counts *counts = nullptr; int rc = 0; std::string err;
};

static void CountExecute(napi_env, void *data) {
    CountJob *j = static_cast<CountJob *>(data);
    kj_count_params p{};
    p.prefix = (const uint8_t *)j->prefix.data(); p.prefix_len = (uint32_t)j->prefix.size();
    p.k = j->k; p.step = j->step;
    j->rc = kj_counts_create(j->ctx, &p, &j->counts);
    if (!j->rc) j->rc = kj_counts_add_file(j->counts, j->path.c_str());
    if (!j->rc) j->rc = kj_counts_finish(j->counts);
    if (j->rc) { j->err = kj_last_error(j->ctx); kj_counts_free(j->counts); j->counts = nullptr; }
}

static void CountComplete(napi_env env, napi_status, void *data) {
    CountJob *j = static_cast<CountJob *>(data);
    if (j->rc) {
        napi_value msg, err;
        napi_create_string_utf8(env, j->err.c_str(), NAPI_AUTO_LENGTH, &msg);
        napi_create_error(env, nullptr, msg, &err);
        napi_reject_deferred(env, j->deferred, err);
    } else {
        napi_value ext;
        napi_create_external(env, j->counts, [](napi_env, void *p, void *) { kj_counts_free((kj_counts *)p); }, nullptr, &ext);
        napi_resolve_deferred(env, j->deferred, ext);
    }
    napi_delete_async_work(env, j->work);
    delete j;
}

static std::string get_string(napi_env env, napi_value v) {
    size_t n = 0; napi_get_value_string_latin1(env, v, nullptr, 0, &n);
    std::string s(n, '\0'); napi_get_value_string_latin1(env, v, &s[0], n + 1, &n);
    return s;
}

static napi_value CountFile(napi_env env, napi_callback_info info) {
    size_t argc = 5; napi_value argv[5];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    CountJob *j = new CountJob();
    j->ctx = unwrap<kj_ctx>(env, argv[0]);
    j->path = get_string(env, argv[1]); j->prefix = get_string(env, argv[2]);
    napi_get_value_uint32(env, argv[3], &j->k); napi_get_value_uint32(env, argv[4], &j->step);
    napi_value promise, name;
    NAPI_OK(napi_create_promise(env, &j->deferred, &promise));
    napi_create_string_utf8(env, "kmerjs_b200.countFile", NAPI_AUTO_LENGTH, &name);
    NAPI_OK(napi_create_async_work(env, nullptr, name, CountExecute, CountComplete, j, &j->work));
    NAPI_OK(napi_queue_async_work(env, j->work));
    return promise;
}

static napi_value CountsExport(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_counts *c = unwrap<kj_counts>(env, argv[0]);
    const uint64_t n = kj_counts_size(c);
    std::vector<uint8_t> keys(n * 32 + 1); std::vector<uint32_t> len(n + 1); std::vector<uint64_t> cnt(n + 1);
    if (int rc = kj_counts_export(c, keys.data(), len.data(), cnt.data())) return throw_kj(env, nullptr, rc);
    napi_value out, jk, jc, v;
    napi_create_object(env, &out); napi_create_array_with_length(env, n, &jk); napi_create_array_with_length(env, n, &jc);
    for (uint64_t i = 0; i < n; ++i) {
        napi_create_string_latin1(env, (const char *)&keys[32 * i], len[i], &v); napi_set_element(env, jk, (uint32_t)i, v);
        napi_create_double(env, (double)cnt[i], &v); napi_set_element(env, jc, (uint32_t)i, v);
    }
    napi_set_named_property(env, out, "keys", jk); napi_set_named_property(env, out, "counts", jc);
    napi_create_double(env, (double)kj_counts_lines(c), &v); napi_set_named_property(env, out, "lines", v);
    napi_create_double(env, (double)kj_counts_bytes_read(c), &v); napi_set_named_property(env, out, "bytesRead", v);
    return out;
}

// kj_row -> the numeric fields of a result row (lib/kmerFinderClient.js:75-89); template / species are resolved in index.js
static napi_value row_to_js(napi_env env, const kj_row &r) {
    napi_value out, v;
    napi_create_object(env, &out);
#define SETD(name, val) napi_create_double(env, (double)(val), &v); napi_set_named_property(env, out, name, v)
    SETD("templateId", r.template_id); SETD("score", r.score); SETD("expected", r.expected); SETD("z", r.z);
    SETD("probability", r.probability); SETD("frac-q", r.frac_q); SETD("frac-d", r.frac_d); SETD("depth", r.depth);
    SETD("kmers-template", r.kmers_template); SETD("total-frac-q", r.total_frac_q);
    SETD("total-frac-d", r.total_frac_d); SETD("total-temp-cover", r.total_temp_cover);
#undef SETD
    return out;
}

static napi_value WtaNext(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_match *m = unwrap<kj_match>(env, argv[0]);
    kj_row r;
    int rc = kj_wta_next(m, &r);
    if (rc < 0) return throw_kj(env, nullptr, rc);       // the two 'No hits were found! (...)' texts
    if (rc == 0) { napi_value out; napi_get_null(env, &out); return out; }
    return row_to_js(env, r);
}

// ---- helpers for array arguments ---------------------------------------------------------------
static bool get_u64_vec(napi_env env, napi_value arr, std::vector<uint64_t> &out) {
    uint32_t n = 0;
    if (napi_get_array_length(env, arr, &n) != napi_ok) return false;
    out.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        napi_value v; double d = 0;
        napi_get_element(env, arr, i, &v);
        napi_get_value_double(env, v, &d);
        out[i] = (uint64_t)d;
    }
    return true;
}

// countsFromMap(ctx, keys: string[], counts: number[], prefix, k, step) -> counts
// A k-mer Map that did not come from countFile (e.g. parsed from JSON): keys in Map order get the
// ordinals 0..n-1.  Regular keys travel as {2-bit key, count, ordinal} records, the others as
// 56-byte byte-string records (kj_counts_merge_host_records / kj_counts_irregular_merge).
static napi_value CountsFromMap(napi_env env, napi_callback_info info) {
    size_t argc = 6; napi_value argv[6];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_ctx *ctx = unwrap<kj_ctx>(env, argv[0]);
    std::vector<uint64_t> counts;
    if (!get_u64_vec(env, argv[2], counts)) return throw_kj(env, ctx, KJ_E_INVALID);
    std::string prefix = get_string(env, argv[3]);
    uint32_t k = 16, step = 1;
    napi_get_value_uint32(env, argv[4], &k); napi_get_value_uint32(env, argv[5], &step);
    kj_count_params p{};
    p.prefix = (const uint8_t *)prefix.data(); p.prefix_len = (uint32_t)prefix.size(); p.k = k; p.step = step;
    kj_counts *c = nullptr;
    int rc = kj_counts_create(ctx, &p, &c);
    if (rc) return throw_kj(env, ctx, rc);
    std::vector<uint64_t> reg;            // {key, count, ordinal} triples
    std::vector<uint8_t> irr;             // 56-byte records {u8 key[32], u64 len, u64 count, u64 ordinal}
    for (uint32_t i = 0; i < counts.size(); ++i) {
        napi_value v; napi_get_element(env, argv[1], i, &v);
        std::string key = get_string(env, v);
        bool regular = key.size() == k && k <= 32;
        uint64_t packed = 0;
        for (char ch : key) {
            if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') { regular = false; break; }
            packed = (packed << 2) | (uint64_t)(((unsigned char)ch >> 1) & 3);     // A=0 C=1 T=2 G=3
        }
        if (regular && packed != ~0ull) { reg.push_back(packed); reg.push_back(counts[i]); reg.push_back(i); }
        else if (key.size() <= 32) {
            size_t at = irr.size(); irr.resize(at + 56, 0);
            memcpy(&irr[at], key.data(), key.size());
            uint64_t tail[3] = {key.size(), counts[i], i};
            memcpy(&irr[at + 32], tail, 24);
        } else { kj_counts_free(c); napi_throw_error(env, nullptr, "k-mers longer than 32 bytes are not supported"); return nullptr; }
    }
    if (!rc && !reg.empty()) rc = kj_counts_merge_host_records(c, reg.data(), reg.size() / 3);
    if (!rc && !irr.empty()) rc = kj_counts_irregular_merge(c, irr.data(), irr.size() / 56);
    if (!rc) rc = kj_counts_finish(c);
    if (rc) { kj_counts_free(c); return throw_kj(env, ctx, rc); }
    napi_value ext;
    NAPI_OK(napi_create_external(env, c, [](napi_env, void *q, void *) { kj_counts_free((kj_counts *)q); }, nullptr, &ext));
    return ext;
}

// dbCreate(ctx, {kmers: string[], lists: number[][], lengths: number[], ulengths: number[],
//                summary: {templates, uniqueLens, totalLen}}) -> db
static napi_value DbCreate(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_ctx *ctx = unwrap<kj_ctx>(env, argv[0]);
    napi_value kmers, lists, lengths, ulengths, summary, v;
    napi_get_named_property(env, argv[1], "kmers", &kmers);
    napi_get_named_property(env, argv[1], "lists", &lists);
    napi_get_named_property(env, argv[1], "lengths", &lengths);
    napi_get_named_property(env, argv[1], "ulengths", &ulengths);
    napi_get_named_property(env, argv[1], "summary", &summary);
    uint32_t n = 0; napi_get_array_length(env, kmers, &n);
    std::vector<uint8_t> bytes; std::vector<uint32_t> klen(n), tm; std::vector<uint64_t> off(1, 0), len, ulen;
    for (uint32_t i = 0; i < n; ++i) {
        napi_get_element(env, kmers, i, &v);
        std::string key = get_string(env, v);
        bytes.insert(bytes.end(), key.begin(), key.end()); klen[i] = (uint32_t)key.size();
        napi_get_element(env, lists, i, &v);
        std::vector<uint64_t> lst; get_u64_vec(env, v, lst);
        for (uint64_t t : lst) tm.push_back((uint32_t)t);
        off.push_back(tm.size());
    }
    get_u64_vec(env, lengths, len); get_u64_vec(env, ulengths, ulen);
    auto num = [&](const char *name) { napi_value x; double d = 0; napi_get_named_property(env, summary, name, &x); napi_get_value_double(env, x, &d); return (uint64_t)d; };
    kj_db_desc d{};
    d.n_kmers = n; d.kmer_bytes = bytes.data(); d.kmer_len = klen.data(); d.list_off = off.data(); d.tmpl_ids = tm.data();
    d.n_templates = (uint32_t)len.size(); d.lengths = len.data(); d.ulengths = ulen.data();
    d.summary_templates = num("templates"); d.summary_unique_lens = num("uniqueLens"); d.summary_total_len = num("totalLen");
    kj_db *db = nullptr;
    int rc = kj_db_create(ctx, &d, &db);
    if (rc) return throw_kj(env, ctx, rc);
    napi_value ext;
    NAPI_OK(napi_create_external(env, db, [](napi_env, void *q, void *) { kj_db_free((kj_db *)q); }, nullptr, &ext));
    return ext;
}

// firstMatch(ctx, counts, db) -> match   (throws 'No hits were found!', lib/kmerFinderClient.js:159-161)
static napi_value FirstMatch(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value argv[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_ctx *ctx = unwrap<kj_ctx>(env, argv[0]);
    kj_match *m = nullptr;
    int rc = kj_first_match(ctx, unwrap<kj_counts>(env, argv[1]), unwrap<kj_db>(env, argv[2]), &m);
    if (rc) return throw_kj(env, ctx, rc);
    napi_value ext;
    NAPI_OK(napi_create_external(env, m, [](napi_env, void *q, void *) { kj_match_free((kj_match *)q); }, nullptr, &ext));
    return ext;
}

// matchScores(match, nTemplates) -> {uScore: number[], tScore: number[], order: number[], hits}
static napi_value MatchScores(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_match *m = unwrap<kj_match>(env, argv[0]);
    uint32_t T = 0; napi_get_value_uint32(env, argv[1], &T);
    std::vector<uint64_t> u(T + 1), t(T + 1); std::vector<uint32_t> order(kj_match_n_matched(m) + 1);
    if (int rc = kj_match_scores(m, u.data(), t.data(), order.data())) return throw_kj(env, nullptr, rc);
    napi_value out, ju, jt, jo, v;
    napi_create_object(env, &out);
    napi_create_array_with_length(env, T, &ju); napi_create_array_with_length(env, T, &jt);
    napi_create_array_with_length(env, kj_match_n_matched(m), &jo);
    for (uint32_t i = 0; i < T; ++i) {
        napi_create_double(env, (double)u[i], &v); napi_set_element(env, ju, i, v);
        napi_create_double(env, (double)t[i], &v); napi_set_element(env, jt, i, v);
    }
    for (uint32_t i = 0; i < kj_match_n_matched(m); ++i) { napi_create_uint32(env, order[i], &v); napi_set_element(env, jo, i, v); }
    napi_set_named_property(env, out, "uScore", ju); napi_set_named_property(env, out, "tScore", jt);
    napi_set_named_property(env, out, "order", jo);
    napi_create_double(env, (double)kj_match_hits(m), &v); napi_set_named_property(env, out, "hits", v);
    return out;
}

// templateKmers(match, templateId) -> number[]  positions (Map order) of the k-mers in the template's `kmers` Set
static napi_value TemplateKmers(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_match *m = unwrap<kj_match>(env, argv[0]);
    uint32_t t = 0; napi_get_value_uint32(env, argv[1], &t);
    uint64_t n = 0;
    if (int rc = kj_match_template_kmers(m, t, nullptr, 0, &n)) return throw_kj(env, nullptr, rc);
    std::vector<uint64_t> idx(n + 1);
    if (int rc = kj_match_template_kmers(m, t, idx.data(), n, &n)) return throw_kj(env, nullptr, rc);
    napi_value out, v;
    napi_create_array_with_length(env, n, &out);
    for (uint64_t i = 0; i < n; ++i) { napi_create_double(env, (double)idx[i], &v); napi_set_element(env, out, (uint32_t)i, v); }
    return out;
}

// countsAlive(counts) -> Uint8Array in Map order (0 = deleted by a winner, kmerMap.delete of lib/kmerFinderClient.js:223-225)
static napi_value CountsAlive(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_counts *c = unwrap<kj_counts>(env, argv[0]);
    const uint64_t n = kj_counts_size(c);
    void *data = nullptr; napi_value buf, out;
    NAPI_OK(napi_create_arraybuffer(env, n, &data, &buf));
    if (int rc = kj_counts_alive(c, (uint8_t *)data)) return throw_kj(env, nullptr, rc);
    NAPI_OK(napi_create_typedarray(env, napi_uint8_array, n, buf, 0, &out));
    return out;
}


// countLine(ctx, line, prefix, k, step) -> {keys: string[], counts: number[]}
// KmerJS#kmersInLine (lib/kmers.js:88-100): the windows of one line, this strand only, no length gate: the
// byte stream API with KJ_F_FORWARD_ONLY | KJ_F_NO_LINE_GATE and base_line = 1 (the buffer IS a sequence line).
static napi_value CountLine(napi_env env, napi_callback_info info) {
    size_t argc = 5; napi_value argv[5];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_ctx *ctx = unwrap<kj_ctx>(env, argv[0]);
    std::string line = get_string(env, argv[1]), prefix = get_string(env, argv[2]);
    uint32_t k = 16, step = 1;
    napi_get_value_uint32(env, argv[3], &k); napi_get_value_uint32(env, argv[4], &step);
    kj_count_params p{};
    p.prefix = (const uint8_t *)prefix.data(); p.prefix_len = (uint32_t)prefix.size();
    p.k = k; p.step = step; p.flags = KJ_F_FORWARD_ONLY | KJ_F_NO_LINE_GATE; p.base_line = 1;
    kj_counts *c = nullptr;
    int rc = kj_counts_create(ctx, &p, &c);
    if (rc == KJ_OK) rc = kj_counts_add_buffer(c, (const uint8_t *)line.data(), line.size(), line.size(), KJ_MEM_HOST, 1);
    if (rc == KJ_OK) rc = kj_counts_finish(c);
    if (rc) { kj_counts_free(c); return throw_kj(env, ctx, rc); }
    const uint64_t n = kj_counts_size(c);
    std::vector<uint8_t> keys(32 * (n + 1)); std::vector<uint32_t> len(n + 1); std::vector<uint64_t> cnt(n + 1);
    rc = kj_counts_export(c, keys.data(), len.data(), cnt.data());
    kj_counts_free(c);
    if (rc) return throw_kj(env, ctx, rc);
    napi_value out, ks, cs, v;
    NAPI_OK(napi_create_object(env, &out));
    NAPI_OK(napi_create_array_with_length(env, n, &ks));
    NAPI_OK(napi_create_array_with_length(env, n, &cs));
    for (uint64_t i = 0; i < n; ++i) {
        napi_create_string_latin1(env, (const char *)keys.data() + 32 * i, len[i], &v); napi_set_element(env, ks, (uint32_t)i, v);
        napi_create_double(env, (double)cnt[i], &v); napi_set_element(env, cs, (uint32_t)i, v);
    }
    napi_set_named_property(env, out, "keys", ks);
    napi_set_named_property(env, out, "counts", cs);
    return out;
}

// dbLoad(ctx, path, summaryPath | null) -> {handle, names, species, lengths, ulengths, summary}   (kj_db_load, any layout)
static napi_value DbLoad(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value argv[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_ctx *ctx = unwrap<kj_ctx>(env, argv[0]);
    std::string path = get_string(env, argv[1]);
    napi_valuetype ty; napi_typeof(env, argv[2], &ty);
    std::string spath = ty == napi_string ? get_string(env, argv[2]) : std::string();
    kj_db *db = nullptr;
    int rc = kj_db_load(ctx, path.c_str(), KJ_DB_AUTO, spath.empty() ? nullptr : spath.c_str(), 0, 1, &db);
    if (rc) return throw_kj(env, ctx, rc);
    napi_value out, ext, names, species, lengths, ulengths, summary, v;
    NAPI_OK(napi_create_object(env, &out));
    NAPI_OK(napi_create_external(env, db, [](napi_env, void *q, void *) { kj_db_free((kj_db *)q); }, nullptr, &ext));
    const uint32_t T = kj_db_n_templates(db);
    napi_create_array_with_length(env, T, &names); napi_create_array_with_length(env, T, &species);
    napi_create_array_with_length(env, T, &lengths); napi_create_array_with_length(env, T, &ulengths);
    for (uint32_t t = 0; t < T; ++t) {
        const char *nm = "", *sp = ""; uint64_t ln = 0, ul = 0;
        kj_db_template(db, t, &nm, &sp, &ln, &ul);
        napi_create_string_utf8(env, nm, NAPI_AUTO_LENGTH, &v); napi_set_element(env, names, t, v);
        napi_create_string_utf8(env, sp, NAPI_AUTO_LENGTH, &v); napi_set_element(env, species, t, v);
        napi_create_double(env, (double)ln, &v); napi_set_element(env, lengths, t, v);
        napi_create_double(env, (double)ul, &v); napi_set_element(env, ulengths, t, v);
    }
    uint64_t st = 0, su = 0, sl = 0;
    kj_db_summary(db, &st, &su, &sl);
    napi_create_object(env, &summary);
    napi_create_double(env, (double)st, &v); napi_set_named_property(env, summary, "templates", v);
    napi_create_double(env, (double)su, &v); napi_set_named_property(env, summary, "uniqueLens", v);
    napi_create_double(env, (double)sl, &v); napi_set_named_property(env, summary, "totalLen", v);
    napi_set_named_property(env, out, "_handle", ext);
    napi_set_named_property(env, out, "names", names); napi_set_named_property(env, out, "species", species);
    napi_set_named_property(env, out, "lengths", lengths); napi_set_named_property(env, out, "ulengths", ulengths);
    napi_set_named_property(env, out, "summary", summary);
    return out;
}

// zScore(roundingMode, r1, n1, r2, n2) -> decimal string (lib/stats.js:19-45, exact, 20 places)
static napi_value ZScore(napi_env env, napi_callback_info info) {
    size_t argc = 5; napi_value argv[5];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    double a[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 5; ++i) napi_get_value_double(env, argv[i], &a[i]);
    double z = 0; char text[256];
    int rc = kj_stats_zscore((int)a[0], (uint64_t)a[1], (uint64_t)a[2], (uint64_t)a[3], (uint64_t)a[4], &z, text, sizeof(text));
    if (rc) return throw_kj(env, nullptr, rc);
    napi_value out; NAPI_OK(napi_create_string_utf8(env, text, NAPI_AUTO_LENGTH, &out));
    return out;
}

// fastp(zText) -> number (lib/stats.js:52-115, exact compare against the thresholds)
static napi_value Fastp(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    double p = 1.0;
    int rc = kj_stats_fastp_text(get_string(env, argv[0]).c_str(), &p);
    if (rc) return throw_kj(env, nullptr, rc);
    napi_value out; NAPI_OK(napi_create_double(env, p, &out));
    return out;
}

// standardScoring(match, nTemplates) -> row[]   (lib/kmerFinderServer.js:857-874)
static napi_value StandardScoring(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_match *m = unwrap<kj_match>(env, argv[0]);
    uint32_t T = 0; napi_get_value_uint32(env, argv[1], &T);
    std::vector<kj_row> rows(T + 1);
    uint32_t n = 0;
    if (int rc = kj_standard_scoring(m, rows.data(), T + 1, &n)) return throw_kj(env, nullptr, rc);
    napi_value out; NAPI_OK(napi_create_array_with_length(env, n, &out));
    for (uint32_t i = 0; i < n; ++i) napi_set_element(env, out, i, row_to_js(env, rows[i]));
    return out;
}

static napi_value ModuleInit(napi_env env, napi_value exports) {
    napi_property_descriptor props[] = {
        {"init", nullptr, Init, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"countFile", nullptr, CountFile, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"countsExport", nullptr, CountsExport, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"wtaNext", nullptr, WtaNext, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"countsFromMap", nullptr, CountsFromMap, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"dbCreate", nullptr, DbCreate, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"firstMatch", nullptr, FirstMatch, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"matchScores", nullptr, MatchScores, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"templateKmers", nullptr, TemplateKmers, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"countsAlive", nullptr, CountsAlive, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"countLine", nullptr, CountLine, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"dbLoad", nullptr, DbLoad, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"zScore", nullptr, ZScore, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"fastp", nullptr, Fastp, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"standardScoring", nullptr, StandardScoring, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
    return exports;
}
NAPI_MODULE(kmerjs_b200, ModuleInit)

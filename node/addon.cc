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

static napi_value WtaNext(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    kj_match *m = unwrap<kj_match>(env, argv[0]);
    kj_row r;
    int rc = kj_wta_next(m, &r);
    if (rc < 0) return throw_kj(env, nullptr, rc);       // the two 'No hits were found! (...)' texts
    napi_value out, v;
    if (rc == 0) { napi_get_null(env, &out); return out; }
    napi_create_object(env, &out);
#define SETD(name, val) napi_create_double(env, (double)(val), &v); napi_set_named_property(env, out, name, v)
    SETD("templateId", r.template_id); SETD("score", r.score); SETD("expected", r.expected); SETD("z", r.z);
    SETD("probability", r.probability); SETD("frac-q", r.frac_q); SETD("frac-d", r.frac_d); SETD("depth", r.depth);
    SETD("kmers-template", r.kmers_template); SETD("total-frac-q", r.total_frac_q);
    SETD("total-frac-d", r.total_frac_d); SETD("total-temp-cover", r.total_temp_cover);
#undef SETD
    return out;
}

// countsFromMap, dbCreate, firstMatch, matchScores, countsAlive follow the same pattern (external
// handles + kj_* calls) and are listed in INTEGRATION.md; omitted here for brevity of the sketch.

static napi_value ModuleInit(napi_env env, napi_value exports) {
    napi_property_descriptor props[] = {
        {"init", nullptr, Init, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"countFile", nullptr, CountFile, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"countsExport", nullptr, CountsExport, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"wtaNext", nullptr, WtaNext, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
    return exports;
}
NAPI_MODULE(kmerjs_b200, ModuleInit)

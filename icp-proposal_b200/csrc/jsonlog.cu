// jsonlog.cu - the chain log in the reference's JSON wire format, at the C boundary (host code only).
//
// Replaces api/sampling/loggers/JSONAcceptRejectLogger.scala:
//   jsonLogFormat (:35)      {index, name, logvalue{...}, status, rigid[9], coeff[K], datetime}
//   accept / reject (:93-106) rejected records carry the CURRENT state's log-values and empty rigid / coeff arrays
//   writeLog (:112-122)       the reference re-serialises the whole list on every call (O(n^2) over a run); here records
//                             are APPENDED: the file is a valid JSON array after every icp_jsonlog_append
//   loadLog (:124-127)        icp_jsonlog_load, which also reads files written by the reference (spray-json prettyPrint)
// The arrays are the device log of icp_chain_run ([step][chain] records), so a chain's log goes from the runner to
// the reference's file format without a host-side mirror in between. spray-json writes null for NaN / infinities; so
// does this writer, and the loader turns null back into NaN.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <map>
#include <string>
#include <vector>

#include "icp_internal.h"

using namespace icp;

struct icp_jsonlog_s {
    FILE *f = nullptr;
    int K = 0;
    std::vector<std::string> names, keys;
    long long n_records = 0;
    bool compact = false;
};

namespace {

std::string json_escape(const std::string &s) {
    std::string o;
    for (char ch : s) {
        switch (ch) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\n': o += "\\n"; break;
            case '\t': o += "\\t"; break;
            default:
                if ((unsigned char)ch < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", ch); o += b; }
                else o += ch;
        }
    }
    return o;
}

void put_number(std::string &o, double v) {
    if (!std::isfinite(v)) { o += "null"; return; }    // spray-json: JsNumber(NaN / Infinity) is JsNull
    char b[40];
    snprintf(b, sizeof b, "%.17g", v);                 // round-trips every double
    o += b;
    if (!strpbrk(b, ".eEn")) o += ".0";                // keep it a JSON float like the reference's Double fields
}

// ---- a small JSON reader (objects, arrays, strings, numbers, true / false / null) ---------------------------------
struct Reader {
    const char *p, *end;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++; }
    bool eat(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
    void need(char c) { if (!eat(c)) throw ArgError{std::string("chain log: expected '") + c + "'"}; }
    std::string str() {
        need('"');
        std::string o;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                p++;
                switch (*p) {
                    case 'n': o += '\n'; break;
                    case 't': o += '\t'; break;
                    case 'r': o += '\r'; break;
                    case 'b': o += '\b'; break;
                    case 'f': o += '\f'; break;
                    case 'u': {
                        if (p + 4 >= end) throw ArgError{"chain log: bad \\u escape"};
                        unsigned v = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                        o += (char)(v < 0x80 ? v : '?');
                        p += 4;
                        break;
                    }
                    default: o += *p;
                }
                p++;
            } else o += *p++;
        }
        need('"');
        return o;
    }
    double num_or_null() {
        ws();
        if (end - p >= 4 && !strncmp(p, "null", 4)) { p += 4; return NAN; }
        char *e = nullptr;
        double v = strtod(p, &e);
        if (e == p) throw ArgError{"chain log: expected a number"};
        p = e;
        return v;
    }
    bool boolean() {
        ws();
        if (end - p >= 4 && !strncmp(p, "true", 4)) { p += 4; return true; }
        if (end - p >= 5 && !strncmp(p, "false", 5)) { p += 5; return false; }
        throw ArgError{"chain log: expected true / false"};
    }
    std::vector<double> numbers() {
        std::vector<double> v;
        need('[');
        if (eat(']')) return v;
        do v.push_back(num_or_null()); while (eat(','));
        need(']');
        return v;
    }
    void skip_value() {
        ws();
        if (p >= end) throw ArgError{"chain log: truncated"};
        if (*p == '"') { str(); return; }
        if (*p == '{') { p++; if (eat('}')) return; do { str(); need(':'); skip_value(); } while (eat(',')); need('}'); return; }
        if (*p == '[') { p++; if (eat(']')) return; do skip_value(); while (eat(',')); need(']'); return; }
        if (*p == 't' || *p == 'f') { boolean(); return; }
        num_or_null();
    }
};

struct Record {
    long long index = 0;
    std::string name;
    std::vector<std::pair<std::string, double>> logvalue;
    bool status = false;
    std::vector<double> rigid, coeff;
};

std::vector<Record> parse_log(const std::string &text) {
    Reader r{text.data(), text.data() + text.size()};
    std::vector<Record> out;
    r.need('[');
    if (r.eat(']')) return out;
    do {
        Record rec;
        r.need('{');
        if (!r.eat('}')) {
            do {
                const std::string key = r.str();
                r.need(':');
                if (key == "index") rec.index = (long long)r.num_or_null();
                else if (key == "name") rec.name = r.str();
                else if (key == "status") rec.status = r.boolean();
                else if (key == "rigid") rec.rigid = r.numbers();
                else if (key == "coeff") rec.coeff = r.numbers();
                else if (key == "logvalue") {
                    r.need('{');
                    if (!r.eat('}')) {
                        do { std::string k = r.str(); r.need(':'); rec.logvalue.push_back({k, r.num_or_null()}); } while (r.eat(','));
                        r.need('}');
                    }
                } else r.skip_value();
            } while (r.eat(','));
            r.need('}');
        }
        out.push_back(std::move(rec));
    } while (r.eat(','));
    r.need(']');
    return out;
}

int32_t host_error(const char *msg) { set_error(nullptr, msg); return ICP_ERR_INVALID_ARGUMENT; }

}  // namespace

extern "C" int32_t icp_jsonlog_open(const char *path, int32_t K, const char *const *component_names, int32_t n_components,
                                    const char *const *value_keys, icp_jsonlog *out) {
    if (!path || !out || K < 1 || n_components < 1 || !component_names || !value_keys) return host_error("icp_jsonlog_open: bad arguments");
    FILE *f = fopen(path, "wb");
    if (!f) return host_error("Writing JSON log file failed!");      // JSONAcceptRejectLogger.scala:119
    icp_jsonlog lg = new icp_jsonlog_s();
    lg->f = f; lg->K = K;
    for (int i = 0; i < n_components; i++) lg->names.push_back(component_names[i] ? component_names[i] : "");
    for (int i = 0; i < 3; i++) lg->keys.push_back(value_keys[i] ? value_keys[i] : "");
    fputs("[]", f);
    fflush(f);
    *out = lg;
    return ICP_OK;
}

extern "C" int32_t icp_jsonlog_append(icp_jsonlog lg, int32_t n_steps, int32_t C, int32_t chain, const int32_t *log_component,
                                      const uint8_t *log_accepted, const double *log_values, const double *log_theta) {
    if (!lg || !lg->f || n_steps < 0 || C < 1 || chain < 0 || chain >= C || !log_component || !log_accepted || !log_values ||
        !log_theta)
        return host_error("icp_jsonlog_append: bad arguments");
    const int K = lg->K, Lt = K + kTheta0;
    char stamp[32];
    {
        time_t now = time(nullptr);
        struct tm tmv;
        localtime_r(&now, &tmv);
        strftime(stamp, sizeof stamp, "%Y-%m-%d %H:%M:%S", &tmv);      // datetimeFormat, JSONAcceptRejectLogger.scala:56
    }
    std::string buf;
    buf.reserve((size_t)n_steps * (64 + 24 * (size_t)(K + 12)));
    for (int s = 0; s < n_steps; s++) {
        const size_t rec = (size_t)s * C + chain;
        const int comp = log_component[rec];
        if (comp < 0 || comp >= (int)lg->names.size()) return host_error("icp_jsonlog_append: component index outside the name table");
        const bool ok = log_accepted[rec] != 0;
        buf += lg->n_records ? ", {\n" : "{\n";
        buf += "  \"index\": " + std::to_string(lg->n_records) + ",\n";
        buf += "  \"name\": \"" + json_escape(lg->names[comp]) + "\",\n";
        buf += "  \"logvalue\": {\n";
        bool first = true;
        for (int k = 0; k < 3; k++) {
            if (lg->keys[k].empty()) continue;
            if (!first) buf += ",\n";
            first = false;
            buf += "    \"" + json_escape(lg->keys[k]) + "\": ";
            put_number(buf, log_values[3 * rec + k]);
        }
        buf += "\n  },\n";
        buf += std::string("  \"status\": ") + (ok ? "true" : "false") + ",\n";
        // a rejected record carries empty parameter arrays (:102-104)
        buf += "  \"rigid\": [";
        if (ok) for (int j = 1; j < kTheta0; j++) { if (j > 1) buf += ", "; put_number(buf, log_theta[rec * Lt + j]); }
        buf += "],\n  \"coeff\": [";
        if (ok) for (int j = 0; j < K; j++) { if (j) buf += ", "; put_number(buf, log_theta[rec * Lt + kTheta0 + j]); }
        buf += "],\n";
        buf += std::string("  \"datetime\": \"") + stamp + "\"\n}";
        lg->n_records++;
    }
    // overwrite the closing bracket, append, close again: a valid JSON array after every call, O(new records) work
    if (fseek(lg->f, -1, SEEK_END) != 0) return host_error("Writing JSON log file failed!");
    buf += "]";
    if (fwrite(buf.data(), 1, buf.size(), lg->f) != buf.size() || fflush(lg->f) != 0) return host_error("Writing JSON log file failed!");
    return ICP_OK;
}

extern "C" int32_t icp_jsonlog_close(icp_jsonlog lg) {
    if (!lg) return ICP_OK;
    if (lg->f) fclose(lg->f);
    delete lg;
    return ICP_OK;
}

extern "C" int32_t icp_jsonlog_load(const char *path, int32_t K, int64_t capacity, int64_t *n_records, int64_t *index,
                                    uint8_t *status, double *values, double *theta, char *names, char *value_keys) {
    try {
        ICP_REQUIRE(path && n_records && K >= 1 && capacity >= 0, "icp_jsonlog_load: bad arguments");
        FILE *f = fopen(path, "rb");
        ICP_REQUIRE(f != nullptr, std::string("cannot open chain log ") + path);
        std::string text;
        char chunk[1 << 16];
        size_t got;
        while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) text.append(chunk, got);
        fclose(f);
        std::vector<Record> recs = parse_log(text);
        *n_records = (int64_t)recs.size();
        if (capacity == 0) return ICP_OK;                    // size query
        ICP_REQUIRE((int64_t)recs.size() <= capacity, "icp_jsonlog_load: capacity smaller than the log");
        // key order of the value triple: product, prior, then the distance evaluator's key (ProductEvaluators.scala:49-53)
        std::string k3[3] = {"product", "prior", ""};
        for (const Record &r : recs)
            for (const auto &kv : r.logvalue)
                if (kv.first != "product" && kv.first != "prior" && k3[2].empty()) k3[2] = kv.first;
        if (value_keys)
            for (int k = 0; k < 3; k++) {
                memset(value_keys + 64 * k, 0, 64);
                strncpy(value_keys + 64 * k, k3[k].c_str(), 63);
            }
        const int Lt = K + kTheta0;
        for (size_t i = 0; i < recs.size(); i++) {
            const Record &r = recs[i];
            if (index) index[i] = r.index;
            if (status) status[i] = r.status ? 1 : 0;
            if (values)
                for (int k = 0; k < 3; k++) {
                    double v = NAN;
                    for (const auto &kv : r.logvalue) if (kv.first == k3[k]) v = kv.second;
                    values[3 * i + k] = v;
                }
            if (theta) {
                double *th = theta + i * Lt;
                for (int j = 0; j < Lt; j++) th[j] = NAN;
                if (!r.rigid.empty() || !r.coeff.empty()) {
                    ICP_REQUIRE(r.rigid.size() == 9 && (int)r.coeff.size() == K,
                                "chain log record " + std::to_string(i) + ": rigid must have 9 entries and coeff K entries");
                    th[0] = 1.0;                                    // sampleToModelParameters (:139-146): scale is not logged
                    for (int j = 0; j < 9; j++) th[1 + j] = r.rigid[j];
                    for (int j = 0; j < K; j++) th[kTheta0 + j] = r.coeff[j];
                }
            }
            if (names) {
                memset(names + 64 * i, 0, 64);
                strncpy(names + 64 * i, r.name.c_str(), 63);
            }
        }
        return ICP_OK;
    } catch (...) {
        return translate_exception(nullptr);
    }
}

// LogHelper.samplesFromLog (apps/util/LogHelper.scala:27-37) / ReplayFittingFromLog.scala:53-66: the log indices burnIn,
// burnIn + takeEveryN, ... below min(n, total), each mapped to the closest accepted record at or before it.
extern "C" int32_t icp_chainlog_sample_indices(int64_t n_records, const uint8_t *status, int32_t take_every_n, int64_t total,
                                               int64_t burn_in, int64_t capacity, int64_t *n_out, int64_t *indices) {
    try {
        ICP_REQUIRE(status && n_out && n_records >= 0 && take_every_n >= 1 && burn_in >= 0 && total >= 0, "icp_chainlog_sample_indices: bad arguments");
        const int64_t stop = std::min<int64_t>(n_records, total);
        int64_t cnt = 0;
        for (int64_t i = burn_in; i < stop; i += take_every_n) {
            int64_t j = i;
            while (j >= 0 && !status[j]) j--;
            ICP_REQUIRE(j >= 0, "no accepted sample at or before the requested log index");    // the reference recurses below 0 and throws
            if (cnt < total) {
                if (indices) { ICP_REQUIRE(cnt < capacity, "icp_chainlog_sample_indices: capacity too small"); indices[cnt] = j; }
                cnt++;
            }
        }
        *n_out = cnt;
        return ICP_OK;
    } catch (...) {
        return translate_exception(nullptr);
    }
}

// Host-side mirror of the digest path of kubegems/modelx's pkg/client, pkg/types and the local FS
// store of pkg/registry, in C++ on top of the modelx_digest.h C ABI (the reference is Go, which this
// build environment cannot compile).  Every function cites the reference code it follows
// (file:line under the modelx tree).  No hashing happens here: digests come from the GPU engine.
#include "../../../include/modelx_client.h"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <dirent.h>
#include <fcntl.h>
#include <map>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(int rc, const std::string& m) { g_err = m; return rc; }
int fail_errno(const std::string& what) { return fail(MXD_ERR_IO, what + ": " + strerror(errno)); }

// ---- pkg/client/push.go:17-23 ---------------------------------------------------------------------
const char* kMediaTypeModelManifestJson = "application/vnd.modelx.model.manifest.v1.json";
const char* kMediaTypeModelConfigYaml = "application/vnd.modelx.model.config.v1.yaml";
const char* kMediaTypeModelFile = "application/vnd.modelx.model.file.v1";
const char* kMediaTypeModelDirectoryTarGz = "application/vnd.modelx.model.directory.v1.tar+gz";
// EmptyFileDigiest, push.go:25
const char* kEmptyFileDigest = "sha256:e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855";
const char* kZeroTime = "0001-01-01T00:00:00Z";  // Go's zero time.Time as JSON

// ---- pkg/types/types.go:28-37, 60-66 ----------------------------------------------------------------
struct Descriptor {
    std::string name, mediaType, digest;
    int64_t size = 0;
    uint32_t mode = 0;                  // Go os.FileMode bits
    std::vector<std::string> urls;
    std::string modified = kZeroTime;   // RFC 3339 (Nano) text, as Go marshals time.Time
    std::map<std::string, std::string> annotations;
};
struct Manifest {
    int64_t schemaVersion = 0;
    std::string mediaType;
    Descriptor config;
    std::vector<Descriptor> blobs;
    bool blobs_null = true;             // a nil slice marshals as null
    std::map<std::string, std::string> annotations;
};

// ---- encoding/json compatible output ------------------------------------------------------------------
// encoding/json string encoding with HTML escaping on (the default of json.Marshal): ", \\, control
// characters, <, >, &, U+2028/U+2029 are escaped; invalid UTF-8 becomes \ufffd.
void json_string(std::string& o, const std::string& s) {
    static const char* hex = "0123456789abcdef";
    o += '"';
    size_t i = 0;
    while (i < s.size()) {
        const unsigned char c = (unsigned char)s[i];
        if (c < 0x80) {
            if (c == '"') o += "\\\"";
            else if (c == '\\') o += "\\\\";
            else if (c == '\n') o += "\\n";
            else if (c == '\r') o += "\\r";
            else if (c == '\t') o += "\\t";
            else if (c < 0x20 || c == '<' || c == '>' || c == '&') { o += "\\u00"; o += hex[c >> 4]; o += hex[c & 15]; }
            else o += (char)c;
            ++i; continue;
        }
        // decode one UTF-8 sequence (shortest form, no surrogates, <= U+10FFFF), as Go's utf8.DecodeRuneInString
        int len = 0; uint32_t cp = 0;
        if (c >= 0xC2 && c <= 0xDF) { len = 2; cp = c & 0x1F; }
        else if (c >= 0xE0 && c <= 0xEF) { len = 3; cp = c & 0x0F; }
        else if (c >= 0xF0 && c <= 0xF4) { len = 4; cp = c & 0x07; }
        bool ok = len != 0 && i + (size_t)len <= s.size();
        for (int k = 1; ok && k < len; ++k) {
            const unsigned char cc = (unsigned char)s[i + k];
            if ((cc & 0xC0) != 0x80) ok = false; else cp = (cp << 6) | (cc & 0x3F);
        }
        if (ok && ((len == 3 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) || (len == 4 && (cp < 0x10000 || cp > 0x10FFFF)))) ok = false;
        if (!ok) { o += "\\ufffd"; ++i; continue; }
        if (cp == 0x2028) o += "\\u2028";
        else if (cp == 0x2029) o += "\\u2029";
        else o.append(s, i, (size_t)len);
        i += (size_t)len;
    }
    o += '"';
}
void json_map(std::string& o, const std::map<std::string, std::string>& m) {
    o += '{';
    bool first = true;
    for (auto& kv : m) { if (!first) o += ','; first = false; json_string(o, kv.first); o += ':'; json_string(o, kv.second); }
    o += '}';
}
void json_descriptor(std::string& o, const Descriptor& d) {
    o += "{\"name\":"; json_string(o, d.name);
    if (!d.mediaType.empty()) { o += ",\"mediaType\":"; json_string(o, d.mediaType); }
    if (!d.digest.empty()) { o += ",\"digest\":"; json_string(o, d.digest); }
    if (d.size != 0) { o += ",\"size\":" + std::to_string(d.size); }
    if (d.mode != 0) { o += ",\"mode\":" + std::to_string(d.mode); }
    if (!d.urls.empty()) {
        o += ",\"urls\":[";
        for (size_t i = 0; i < d.urls.size(); ++i) { if (i) o += ','; json_string(o, d.urls[i]); }
        o += ']';
    }
    o += ",\"modified\":"; json_string(o, d.modified);   // struct-typed field: omitempty never drops it
    if (!d.annotations.empty()) { o += ",\"annotations\":"; json_map(o, d.annotations); }
    o += '}';
}
std::string json_manifest(const Manifest& m) {
    std::string o = "{\"schemaVersion\":" + std::to_string(m.schemaVersion);
    if (!m.mediaType.empty()) { o += ",\"mediaType\":"; json_string(o, m.mediaType); }
    o += ",\"config\":"; json_descriptor(o, m.config);
    o += ",\"blobs\":";
    if (m.blobs.empty() && m.blobs_null) o += "null";
    else { o += '['; for (size_t i = 0; i < m.blobs.size(); ++i) { if (i) o += ','; json_descriptor(o, m.blobs[i]); } o += ']'; }
    if (!m.annotations.empty()) { o += ",\"annotations\":"; json_map(o, m.annotations); }
    o += '}';
    return o;
}

// ---- a small JSON reader (enough for manifests) ------------------------------------------------------
struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false; double num = 0; std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal* get(const char* k) const { for (auto& kv : obj) if (kv.first == k) return &kv.second; return nullptr; }
};
struct JParser {
    const char* p; const char* e; std::string err;
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    bool lit(const char* s) { size_t n = strlen(s); if ((size_t)(e - p) >= n && !memcmp(p, s, n)) { p += n; return true; } return false; }
    static void utf8(std::string& o, unsigned cp) {
        if (cp < 0x80) o += (char)cp;
        else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 63)); }
        else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 63)); o += (char)(0x80 | (cp & 63)); }
        else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 63)); o += (char)(0x80 | ((cp >> 6) & 63)); o += (char)(0x80 | (cp & 63)); }
    }
    bool str(std::string& o) {
        if (p >= e || *p != '"') { err = "expected string"; return false; }
        ++p;
        while (p < e && *p != '"') {
            if (*p == '\\') {
                if (++p >= e) break;
                char c = *p++;
                switch (c) {
                    case 'n': o += '\n'; break; case 't': o += '\t'; break; case 'r': o += '\r'; break;
                    case 'b': o += '\b'; break; case 'f': o += '\f'; break; case '/': o += '/'; break;
                    case '\\': o += '\\'; break; case '"': o += '"'; break;
                    case 'u': {
                        if (e - p < 4) { err = "bad \\u"; return false; }
                        unsigned cp = (unsigned)strtoul(std::string(p, 4).c_str(), nullptr, 16); p += 4;
                        if (cp >= 0xD800 && cp < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                            unsigned lo = (unsigned)strtoul(std::string(p + 2, 4).c_str(), nullptr, 16);
                            if (lo >= 0xDC00 && lo < 0xE000) { cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); p += 6; }
                        }
                        utf8(o, cp); break;
                    }
                    default: err = "bad escape"; return false;
                }
            } else o += *p++;
        }
        if (p >= e) { err = "unterminated string"; return false; }
        ++p; return true;
    }
    bool val(JVal& v, int depth = 0) {
        if (depth > 64) { err = "too deep"; return false; }
        ws();
        if (p >= e) { err = "unexpected end"; return false; }
        if (*p == '{') {
            v.kind = JVal::Obj; ++p; ws();
            if (p < e && *p == '}') { ++p; return true; }
            for (;;) {
                ws(); std::string k; if (!str(k)) return false;
                ws(); if (p >= e || *p != ':') { err = "expected ':'"; return false; } ++p;
                JVal c; if (!val(c, depth + 1)) return false;
                v.obj.emplace_back(std::move(k), std::move(c));
                ws(); if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == '}') { ++p; return true; }
                err = "expected ',' or '}'"; return false;
            }
        }
        if (*p == '[') {
            v.kind = JVal::Arr; ++p; ws();
            if (p < e && *p == ']') { ++p; return true; }
            for (;;) {
                JVal c; if (!val(c, depth + 1)) return false;
                v.arr.push_back(std::move(c));
                ws(); if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == ']') { ++p; return true; }
                err = "expected ',' or ']'"; return false;
            }
        }
        if (*p == '"') { v.kind = JVal::Str; return str(v.str); }
        if (lit("null")) { v.kind = JVal::Null; return true; }
        if (lit("true")) { v.kind = JVal::Bool; v.b = true; return true; }
        if (lit("false")) { v.kind = JVal::Bool; v.b = false; return true; }
        char* endp = nullptr;
        v.num = strtod(p, &endp);
        if (endp == p) { err = "unexpected character"; return false; }
        v.kind = JVal::Num; v.str.assign(p, (size_t)(endp - p)); p = endp; return true;
    }
};

bool descriptor_from(const JVal& v, Descriptor* d, std::string* err) {
    if (v.kind != JVal::Obj) { *err = "descriptor is not an object"; return false; }
    if (auto* x = v.get("name")) d->name = x->str;
    if (auto* x = v.get("mediaType")) d->mediaType = x->str;
    if (auto* x = v.get("digest")) d->digest = x->str;
    if (auto* x = v.get("size")) d->size = strtoll(x->str.c_str(), nullptr, 10);
    if (auto* x = v.get("mode")) d->mode = (uint32_t)strtoull(x->str.c_str(), nullptr, 10);
    if (auto* x = v.get("urls")) for (auto& u : x->arr) d->urls.push_back(u.str);
    if (auto* x = v.get("modified")) if (x->kind == JVal::Str) d->modified = x->str;
    if (auto* x = v.get("annotations")) for (auto& kv : x->obj) d->annotations[kv.first] = kv.second.str;
    return true;
}
bool manifest_from_json(const char* text, Manifest* m, std::string* err) {
    JParser jp{text, text + strlen(text), ""};
    JVal root;
    if (!jp.val(root)) { *err = jp.err; return false; }
    if (root.kind != JVal::Obj) { *err = "manifest is not an object"; return false; }
    if (auto* x = root.get("schemaVersion")) m->schemaVersion = strtoll(x->str.c_str(), nullptr, 10);
    if (auto* x = root.get("mediaType")) m->mediaType = x->str;
    if (auto* x = root.get("config")) if (!descriptor_from(*x, &m->config, err)) return false;
    if (auto* x = root.get("blobs")) {
        if (x->kind == JVal::Arr) {
            m->blobs_null = false;
            for (auto& b : x->arr) { Descriptor d; if (!descriptor_from(b, &d, err)) return false; m->blobs.push_back(std::move(d)); }
        }
    }
    if (auto* x = root.get("annotations")) for (auto& kv : x->obj) m->annotations[kv.first] = kv.second.str;
    return true;
}

// ---- os.FileMode / time.Time as the Go client would see them -------------------------------------------
uint32_t go_file_mode(mode_t m) {   // os.FileMode bits (Go's os/types.go), from a Unix st_mode
    uint32_t g = (uint32_t)(m & 0777);
    switch (m & S_IFMT) {
        case S_IFDIR: g |= 1u << 31; break;           // ModeDir
        case S_IFLNK: g |= 1u << 27; break;           // ModeSymlink
        case S_IFBLK: g |= 1u << 26; break;           // ModeDevice
        case S_IFCHR: g |= (1u << 26) | (1u << 21); break;  // ModeDevice | ModeCharDevice
        case S_IFIFO: g |= 1u << 25; break;           // ModeNamedPipe
        case S_IFSOCK: g |= 1u << 24; break;          // ModeSocket
        default: break;
    }
    if (m & S_ISUID) g |= 1u << 23;
    if (m & S_ISGID) g |= 1u << 22;
    if (m & S_ISVTX) g |= 1u << 20;
    return g;
}
std::string go_time_json(const struct timespec& ts) {   // time.Time.MarshalJSON: RFC3339Nano in the local zone
    struct tm tmv;
    time_t sec = ts.tv_sec;
    localtime_r(&sec, &tmv);
    char buf[64];
    strftime(buf, sizeof buf, "%Y-%m-%dT%H:%M:%S", &tmv);
    std::string o = buf;
    if (ts.tv_nsec) {
        char frac[16]; snprintf(frac, sizeof frac, "%09ld", ts.tv_nsec);
        std::string f = frac; while (!f.empty() && f.back() == '0') f.pop_back();
        o += "." + f;
    }
    long off = tmv.tm_gmtoff;
    if (off == 0) o += "Z";
    else { char z[48]; long a = off < 0 ? -off : off; snprintf(z, sizeof z, "%c%02ld:%02ld", off < 0 ? '-' : '+', a / 3600, (a % 3600) / 60); o += z; }
    return o;
}

std::string join(const std::string& a, const std::string& b) {
    if (a.empty()) return b;
    return a.back() == '/' ? a + b : a + "/" + b;
}
char* dup_out(const std::string& s) { char* p = (char*)malloc(s.size() + 1); if (p) memcpy(p, s.c_str(), s.size() + 1); return p; }

// ---- ParseManifest, pkg/client/push.go:67-100 ------------------------------------------------------------
int parse_manifest(const std::string& basedir, const std::string& configfile, Manifest* m) {
    m->mediaType = kMediaTypeModelManifestJson;
    DIR* d = opendir(basedir.c_str());
    if (!d) return fail_errno("open " + basedir);
    std::vector<std::pair<std::string, bool>> entries;   // name, is_dir
    while (struct dirent* de = readdir(d)) {
        std::string name = de->d_name;
        if (name == "." || name == "..") continue;
        bool is_dir = de->d_type == DT_DIR;
        if (de->d_type == DT_UNKNOWN) { struct stat st; if (lstat(join(basedir, name).c_str(), &st) == 0) is_dir = S_ISDIR(st.st_mode); }
        entries.emplace_back(name, is_dir);
    }
    closedir(d);
    std::sort(entries.begin(), entries.end());            // os.ReadDir returns entries sorted by filename
    for (auto& e : entries) {
        if (!e.first.empty() && e.first[0] == '.') continue;             // push.go:76-78
        if (e.first == configfile) { m->config.name = e.first; m->config.mediaType = kMediaTypeModelConfigYaml; continue; }
        Descriptor desc; desc.name = e.first;
        desc.mediaType = e.second ? kMediaTypeModelDirectoryTarGz : kMediaTypeModelFile;
        m->blobs.push_back(std::move(desc));
        m->blobs_null = false;
    }
    std::sort(m->blobs.begin(), m->blobs.end(), [](const Descriptor& a, const Descriptor& b) { return a.name < b.name; });  // :98
    return MXD_OK;
}

// the stat half of pushFile, push.go:120-142 (fields are only filled when still zero)
int push_file_fill(const std::string& path, Descriptor* d) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0) return fail_errno("stat " + path);
    if (d->size == 0) d->size = (int64_t)st.st_size;
    if (d->mode == 0) d->mode = go_file_mode(st.st_mode);
    if (d->modified == kZeroTime) d->modified = go_time_json(st.st_mtim);
    return MXD_OK;
}

std::string digest_str(const uint8_t* d) { char s[72]; mxd_digest_string(d, s); return s; }

// ---- opt-in digest cache: <basedir>/.modelx/digests.json = {"<name>":{"size":..,"mtime_ns":..,"digest":".."}} ----
struct CacheEntry { int64_t size = 0; int64_t mtime_ns = 0; std::string digest; };
std::string cache_path(const std::string& basedir) { return join(join(basedir, ".modelx"), "digests.json"); }
int write_file(const std::string& path, const std::string& data, mode_t mode);
int mkdir_all(const std::string& path, mode_t mode);
void cache_load(const std::string& basedir, std::map<std::string, CacheEntry>* out) {
    FILE* f = fopen(cache_path(basedir).c_str(), "rb");
    if (!f) return;
    std::string text; char buf[65536]; size_t r;
    while ((r = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, r);
    fclose(f);
    JParser jp{text.c_str(), text.c_str() + text.size(), ""};
    JVal root;
    if (!jp.val(root) || root.kind != JVal::Obj) return;        // unreadable cache = no cache
    for (auto& kv : root.obj) {
        CacheEntry e;
        if (auto* x = kv.second.get("size")) e.size = strtoll(x->str.c_str(), nullptr, 10);
        if (auto* x = kv.second.get("mtime_ns")) e.mtime_ns = strtoll(x->str.c_str(), nullptr, 10);
        if (auto* x = kv.second.get("digest")) e.digest = x->str;
        uint8_t tmp[32];
        if (mxd_digest_parse(e.digest.c_str(), tmp) == MXD_OK) (*out)[kv.first] = e;
    }
}
void cache_store(const std::string& basedir, const std::map<std::string, CacheEntry>& m) {
    std::string o = "{";
    bool first = true;
    for (auto& kv : m) {
        if (!first) o += ',';
        first = false;
        json_string(o, kv.first);
        o += ":{\"size\":" + std::to_string(kv.second.size) + ",\"mtime_ns\":" + std::to_string(kv.second.mtime_ns) + ",\"digest\":";
        json_string(o, kv.second.digest); o += '}';
    }
    o += '}';
    if (mkdir_all(join(basedir, ".modelx"), 0755) == MXD_OK) write_file(cache_path(basedir), o, 0644);   // best effort
}

// Digest phase of Client.Push, push.go:29-52: every file blob + the config through pushFile's
// "stat, digest" (push.go:120-142); the digests come from one lock-step GPU batch.
int push_digest(mxd_ctx* ctx, const std::string& basedir, const std::string& configfile, int flags, Manifest* m) {
    const bool with_tree = (flags & MXC_PUSH_TREE) != 0, use_cache = (flags & MXC_PUSH_CACHE) != 0;
    int rc = parse_manifest(basedir, configfile, m);
    if (rc != MXD_OK) return rc;
    std::vector<Descriptor*> files;
    for (auto& b : m->blobs) {
        if (b.mediaType == kMediaTypeModelDirectoryTarGz)
            return fail(MXC_ERR_UNSUPPORTED, "directory blob '" + b.name + "': tar+gzip is outside the digest path (DESIGN.md section 8)");
        files.push_back(&b);
    }
    if (m->config.name.empty()) return fail(MXD_ERR_IO, "stat " + join(basedir, configfile) + ": no such file or directory");
    files.push_back(&m->config);
    std::vector<std::string> paths;
    for (auto* d : files) paths.push_back(join(basedir, d->name));
    std::map<std::string, CacheEntry> cache;
    if (use_cache) cache_load(basedir, &cache);
    std::vector<struct stat> sts(files.size());
    std::vector<size_t> todo;                         // indices that really need hashing
    for (size_t i = 0; i < files.size(); ++i) {
        if (stat(paths[i].c_str(), &sts[i]) != 0) return fail_errno("stat " + paths[i]);
        const int64_t mt = (int64_t)sts[i].st_mtim.tv_sec * 1000000000ll + sts[i].st_mtim.tv_nsec;
        auto it = cache.find(files[i]->name);
        if (use_cache && it != cache.end() && it->second.size == (int64_t)sts[i].st_size && it->second.mtime_ns == mt)
            files[i]->digest = it->second.digest;
        else todo.push_back(i);
    }
    if (!todo.empty()) {
        std::vector<const char*> cpaths;
        for (size_t i : todo) cpaths.push_back(paths[i].c_str());
        std::vector<uint8_t> out(32 * todo.size());
        rc = mxd_sha256_files(ctx, cpaths.data(), todo.size(), out.data(), nullptr);
        if (rc != MXD_OK) return fail(rc, std::string("digest: ") + mxd_last_error());
        for (size_t k = 0; k < todo.size(); ++k) files[todo[k]]->digest = digest_str(&out[32 * k]);
    }
    for (size_t i = 0; i < files.size(); ++i) {
        rc = push_file_fill(paths[i], files[i]);
        if (rc != MXD_OK) return rc;
        if (use_cache) {
            CacheEntry e; e.size = (int64_t)sts[i].st_size;
            e.mtime_ns = (int64_t)sts[i].st_mtim.tv_sec * 1000000000ll + sts[i].st_mtim.tv_nsec; e.digest = files[i]->digest;
            cache[files[i]->name] = e;
        }
        if (with_tree) {
            uint8_t root[32]; uint64_t nch = 0, sz = 0;
            rc = mxd_tree_digest_file(ctx, paths[i].c_str(), nullptr, nullptr, 0, &nch, &sz, root);
            if (rc != MXD_OK) return fail(rc, std::string("tree digest: ") + mxd_last_error());
            files[i]->annotations["modelx.tree.v1"] = digest_str(root) + ";leaf=16384;fanout=8;chunk=8388608;chunks=" + std::to_string(nch);
        }
    }
    if (use_cache) cache_store(basedir, cache);
    return MXD_OK;
}

// ---- pkg/registry local FS store ------------------------------------------------------------------------
// BlobDigestPath, store.go:56-61: path.Join(repository, "blobs", algorithm, hex)
int blob_digest_path(const std::string& repo, const std::string& digest, std::string* out) {
    size_t colon = digest.find(':');
    if (colon == std::string::npos) return fail(MXC_ERR_DIGEST_INVALID, "digest invalid: " + digest);
    *out = join(join(join(repo, "blobs"), digest.substr(0, colon)), digest.substr(colon + 1));
    return MXD_OK;
}
int mkdir_all(const std::string& path, mode_t mode) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); ++i) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty() && mkdir(cur.c_str(), mode) != 0 && errno != EEXIST) return fail_errno("mkdir " + cur);
        }
        if (i < path.size()) cur += path[i];
    }
    return MXD_OK;
}
std::string dir_of(const std::string& p) { size_t s = p.rfind('/'); return s == std::string::npos ? "." : p.substr(0, s); }
int write_file(const std::string& path, const std::string& data, mode_t mode) {
    int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, mode);
    if (fd < 0) return fail_errno("open " + path);
    size_t off = 0;
    while (off < data.size()) { ssize_t w = write(fd, data.data() + off, data.size() - off); if (w < 0) { if (errno == EINTR) continue; int e = errno; close(fd); errno = e; return fail_errno("write " + path); } off += (size_t)w; }
    close(fd);
    return MXD_OK;
}
int copy_file(const std::string& src, const std::string& dst, mode_t mode, int64_t* copied) {
    int in = open(src.c_str(), O_RDONLY | O_CLOEXEC);
    if (in < 0) return fail_errno("open " + src);
    int out = open(dst.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, mode);
    if (out < 0) { int e = errno; close(in); errno = e; return fail_errno("open " + dst); }
    std::vector<char> buf(4 << 20);
    int64_t total = 0; int rc = MXD_OK;
    for (;;) {
        ssize_t r = read(in, buf.data(), buf.size());
        if (r < 0) { if (errno == EINTR) continue; rc = fail_errno("read " + src); break; }
        if (r == 0) break;
        ssize_t off = 0;
        while (off < r) { ssize_t w = write(out, buf.data() + off, (size_t)(r - off)); if (w < 0) { if (errno == EINTR) continue; rc = fail_errno("write " + dst); break; } off += w; }
        if (rc != MXD_OK) break;
        total += r;
    }
    close(in); close(out);
    if (copied) *copied = total;
    return rc;
}
// localFileMeta + json.MarshalIndent(meta, "", "  "), fs_local.go:41-44,155-169
std::string meta_json(const std::string& content_type, int64_t content_length) {
    std::string o = "{";
    bool any = false;
    if (!content_type.empty()) { o += "\n  \"contentType\": "; json_string(o, content_type); any = true; }
    if (content_length != 0) { o += any ? ",\n  \"contentLength\": " : "\n  \"contentLength\": "; o += std::to_string(content_length); any = true; }
    o += any ? "\n}" : "}";
    return o;
}
// FSRegistryStore.PutBlob (store_fs.go:358-364) -> LocalFSProvider.Put (fs_local.go:46-51): meta first, then data
int fs_put_file(const std::string& basepath, const std::string& rel, const std::string& content_type, int64_t content_length,
                const std::string* srcfile, const std::string* inline_data) {
    const std::string datafile = join(basepath, rel), metafile = datafile + ".meta";
    int rc = mkdir_all(dir_of(metafile), 0755);
    if (rc != MXD_OK) return rc;
    rc = write_file(metafile, meta_json(content_type, content_length), 0644);
    if (rc != MXD_OK) return rc;
    if (srcfile) return copy_file(*srcfile, datafile, 0644, nullptr);
    return write_file(datafile, *inline_data, 0644);
}
int fs_exists(const std::string& basepath, const std::string& rel) {   // LocalFSProvider.Exists, fs_local.go:76-85
    struct stat st;
    if (stat(join(basepath, rel).c_str(), &st) == 0) return 1;
    if (errno == ENOENT || errno == ENOTDIR) return 0;
    return fail_errno("stat " + join(basepath, rel));
}
int fs_put_blob(mxd_ctx* ctx, const std::string& basepath, const std::string& repo, const std::string& digest,
                const std::string& content_type, const std::string& srcfile, int verify) {
    uint8_t want[32];
    if (mxd_digest_parse(digest.c_str(), want) != MXD_OK) return fail(MXC_ERR_DIGEST_INVALID, "digest invalid: " + digest);  // BlobDigestFun, registry.go:218-227
    if (content_type.empty()) return fail(MXD_ERR_INVALID, "content type invalid: empty");                                // registry.go:147-151
    std::string rel;
    int rc = blob_digest_path(repo, digest, &rel);
    if (rc != MXD_OK) return rc;
    struct stat st;
    if (stat(srcfile.c_str(), &st) != 0) return fail_errno("stat " + srcfile);
    rc = fs_put_file(basepath, rel, content_type, (int64_t)st.st_size, &srcfile, nullptr);
    if (rc != MXD_OK) return rc;
    if (verify) {   // new: digest verification of what was stored
        if (!ctx) return fail(MXD_ERR_INVALID, "verify needs an engine context");
        const std::string stored = join(basepath, rel);
        const char* p[1] = {stored.c_str()};
        uint8_t ok = 0;
        if (verify == 2) {          // the key is a modelx.tree.v1 root (mxc_push_local_tree)
            uint8_t root[32]; uint64_t nch = 0, sz = 0;
            rc = mxd_tree_digest_file(ctx, stored.c_str(), nullptr, nullptr, 0, &nch, &sz, root);
            ok = rc == MXD_OK && memcmp(root, want, 32) == 0;
        } else {
            rc = mxd_verify_files(ctx, p, want, 1, &ok);
        }
        if (rc != MXD_OK) return fail(rc, std::string("verify: ") + mxd_last_error());
        if (!ok) { unlink(stored.c_str()); unlink((stored + ".meta").c_str()); return fail(MXC_ERR_DIGEST_INVALID, "digest invalid: " + digest); }
    }
    return MXD_OK;
}
int read_file(const std::string& path, std::string* out) {
    int fd = open(path.c_str(), O_RDONLY | O_CLOEXEC);
    if (fd < 0) return errno == ENOENT ? fail(MXC_ERR_NOT_FOUND, "not found: " + path) : fail_errno("open " + path);
    char buf[65536];
    for (;;) { ssize_t r = read(fd, buf, sizeof buf); if (r < 0) { if (errno == EINTR) continue; int e = errno; close(fd); errno = e; return fail_errno("read " + path); } if (r == 0) break; out->append(buf, (size_t)r); }
    close(fd);
    return MXD_OK;
}

const char* kTreeMode = "tree.v1";
const char* kTreeAnnotation = "modelx.digest";
bool is_tree_keyed(const Descriptor& d) {
    auto it = d.annotations.find(kTreeAnnotation);
    return it != d.annotations.end() && it->second.compare(0, strlen(kTreeMode), kTreeMode) == 0;
}

// pullFile's check, pull.go:111-136, for a list of descriptors.  Whole-file digests of all present files are one
// GPU batch; descriptors annotated as tree-keyed (mxc_push_local_tree) are checked with the tree digest instead.
struct PullState { std::string name, digest, state; };
int pull_check(mxd_ctx* ctx, const std::string& basedir, const std::vector<Descriptor>& descs, std::vector<PullState>* out) {
    std::vector<size_t> present;
    std::vector<std::string> paths;
    for (size_t i = 0; i < descs.size(); ++i) {
        const Descriptor& d = descs[i];
        if (d.mediaType == kMediaTypeModelDirectoryTarGz)
            return fail(MXC_ERR_UNSUPPORTED, "directory blob '" + d.name + "': tar+gzip is outside the digest path");
        out->push_back({d.name, d.digest, "missing"});
        struct stat st;
        const std::string p = join(basedir, d.name);
        if (stat(p.c_str(), &st) == 0) {
            // os.Open succeeds on a directory and digest.FromReader then fails with EISDIR (pull.go:116-119)
            if (S_ISDIR(st.st_mode)) return fail(MXD_ERR_IO, "read " + p + ": is a directory");
            if (is_tree_keyed(d)) {
                uint8_t root[32]; uint64_t nch = 0, sz = 0;
                int rc = mxd_tree_digest_file(ctx, p.c_str(), nullptr, nullptr, 0, &nch, &sz, root);
                if (rc != MXD_OK) return fail(rc, std::string("tree digest: ") + mxd_last_error());
                (*out)[i].state = (digest_str(root) == d.digest) ? "already exists" : "differs";
            } else {
                present.push_back(i); paths.push_back(p);
            }
        } else if (errno != ENOENT && errno != ENOTDIR) {
            return fail_errno("open " + p);                                      // pull.go:125-127
        }
    }
    if (!present.empty()) {
        std::vector<const char*> cp; for (auto& p : paths) cp.push_back(p.c_str());
        std::vector<uint8_t> got(32 * present.size());
        int rc = mxd_sha256_files(ctx, cp.data(), present.size(), got.data(), nullptr);
        if (rc != MXD_OK) return fail(rc, std::string("digest: ") + mxd_last_error());
        for (size_t k = 0; k < present.size(); ++k)                           // pull.go:120 string equality
            (*out)[present[k]].state = (digest_str(&got[32 * k]) == descs[present[k]].digest) ? "already exists" : "differs";
    }
    for (size_t i = 0; i < descs.size(); ++i) {
        PullState& s = (*out)[i];
        if (s.state == "already exists") continue;
        const bool empty = is_tree_keyed(descs[i]) ? descs[i].size == 0 : s.digest == kEmptyFileDigest;   // pull.go:134-136
        if (empty) s.state = "empty";
    }
    return MXD_OK;
}

std::string manifest_path(const std::string& repo, const std::string& ref) { return join(join(repo, "manifests"), ref); }   // store.go:67-69

}  // namespace

extern "C" {

const char* mxc_last_error(void) { return g_err.c_str(); }
void mxc_free(char* p) { free(p); }

int mxc_parse_manifest(const char* basedir, const char* configfile, char** manifest_json) {
    if (!basedir || !configfile || !manifest_json) return fail(MXD_ERR_INVALID, "parse_manifest: null argument");
    Manifest m;
    int rc = parse_manifest(basedir, configfile, &m);
    if (rc != MXD_OK) return rc;
    *manifest_json = dup_out(json_manifest(m));
    return MXD_OK;
}

int mxc_push_digest(mxd_ctx* ctx, const char* basedir, const char* configfile, int flags, char** manifest_json) {
    if (!ctx || !basedir || !configfile || !manifest_json) return fail(MXD_ERR_INVALID, "push_digest: null argument");
    Manifest m;
    int rc = push_digest(ctx, basedir, configfile, flags, &m);
    if (rc != MXD_OK) return rc;
    *manifest_json = dup_out(json_manifest(m));
    return MXD_OK;
}

int mxc_pull_check(mxd_ctx* ctx, const char* basedir, const char* manifest_json, char** report_json) {
    if (!ctx || !basedir || !manifest_json || !report_json) return fail(MXD_ERR_INVALID, "pull_check: null argument");
    Manifest m; std::string err;
    if (!manifest_from_json(manifest_json, &m, &err)) return fail(MXC_ERR_MANIFEST, "manifest invalid: " + err);
    std::vector<Descriptor> descs = m.blobs;      // append(manifest.Blobs, manifest.Config), pull.go:38
    descs.push_back(m.config);
    std::vector<PullState> st;
    int rc = pull_check(ctx, basedir, descs, &st);
    if (rc != MXD_OK) return rc;
    std::string o = "[";
    for (size_t i = 0; i < st.size(); ++i) {
        if (i) o += ',';
        o += "{\"name\":"; json_string(o, st[i].name); o += ",\"state\":"; json_string(o, st[i].state);
        o += ",\"digest\":"; json_string(o, st[i].digest); o += '}';
    }
    o += ']';
    *report_json = dup_out(o);
    return MXD_OK;
}

int mxc_blob_digest_path(const char* repository, const char* digest, char** path) {
    if (!repository || !digest || !path) return fail(MXD_ERR_INVALID, "blob_digest_path: null argument");
    std::string rel;
    int rc = blob_digest_path(repository, digest, &rel);
    if (rc != MXD_OK) return rc;
    *path = dup_out(rel);
    return MXD_OK;
}

int mxc_fs_put_blob(mxd_ctx* ctx, const char* basepath, const char* repository, const char* digest,
                    const char* content_type, const char* srcfile, int verify) {
    if (!basepath || !repository || !digest || !content_type || !srcfile) return fail(MXD_ERR_INVALID, "fs_put_blob: null argument");
    return fs_put_blob(ctx, basepath, repository, digest, content_type, srcfile, verify);
}

int mxc_fs_exists_blob(const char* basepath, const char* repository, const char* digest) {
    if (!basepath || !repository || !digest) return fail(MXD_ERR_INVALID, "fs_exists_blob: null argument");
    std::string rel;
    int rc = blob_digest_path(repository, digest, &rel);
    if (rc != MXD_OK) return rc;
    return fs_exists(basepath, rel);
}

int mxc_fs_put_manifest(const char* basepath, const char* repository, const char* reference, const char* content_type,
                        const char* manifest_json) {
    if (!basepath || !repository || !reference || !manifest_json) return fail(MXD_ERR_INVALID, "fs_put_manifest: null argument");
    Manifest m; std::string err;
    if (!manifest_from_json(manifest_json, &m, &err)) return fail(MXC_ERR_MANIFEST, "manifest invalid: " + err);
    const std::string content = json_manifest(m);          // json.Marshal(manifest), store_fs.go:88
    return fs_put_file(basepath, manifest_path(repository, reference), content_type ? content_type : "", (int64_t)content.size(),
                       nullptr, &content);
}

int mxc_fs_get_manifest(const char* basepath, const char* repository, const char* reference, char** manifest_json) {
    if (!basepath || !repository || !reference || !manifest_json) return fail(MXD_ERR_INVALID, "fs_get_manifest: null argument");
    std::string text;
    int rc = read_file(join(basepath, manifest_path(repository, reference)), &text);
    if (rc != MXD_OK) return rc;
    *manifest_json = dup_out(text);
    return MXD_OK;
}

int mxc_push_local(mxd_ctx* ctx, const char* basedir, const char* configfile, const char* basepath, const char* repository,
                   const char* version, int verify, char** report_json) {
    if (!ctx || !basedir || !configfile || !basepath || !repository || !version || !report_json)
        return fail(MXD_ERR_INVALID, "push_local: null argument");
    Manifest m;
    int rc = push_digest(ctx, basedir, configfile, 0, &m);
    if (rc != MXD_OK) return rc;
    std::vector<Descriptor*> all;
    for (auto& b : m.blobs) all.push_back(&b);
    all.push_back(&m.config);
    std::string blobs = "[";
    for (size_t i = 0; i < all.size(); ++i) {
        const Descriptor& d = *all[i];
        std::string status;
        if (d.digest == kEmptyFileDigest) status = "empty";                       // push.go:165-168
        else {
            std::string rel; rc = blob_digest_path(repository, d.digest, &rel); if (rc != MXD_OK) return rc;
            int ex = fs_exists(basepath, rel);                                      // HeadBlob, push.go:169-177
            if (ex < 0) return ex;
            if (ex) status = "exists";
            else {
                // fallback upload through the server: Content-Type application/octet-stream (client/registry.go:109-120)
                rc = fs_put_blob(ctx, basepath, repository, d.digest, "application/octet-stream", join(basedir, d.name), verify ? 1 : 0);
                if (rc != MXD_OK) return rc;
                status = "done";
            }
        }
        if (i) blobs += ',';
        blobs += "{\"name\":"; json_string(blobs, d.name); blobs += ",\"status\":"; json_string(blobs, status);
        blobs += ",\"digest\":"; json_string(blobs, d.digest); blobs += '}';
    }
    blobs += ']';
    const std::string mj = json_manifest(m);
    rc = fs_put_file(basepath, manifest_path(repository, version), kMediaTypeModelManifestJson, (int64_t)mj.size(), nullptr, &mj);  // push.go:57-64
    if (rc != MXD_OK) return rc;
    *report_json = dup_out("{\"manifest\":" + mj + ",\"blobs\":" + blobs + "}");
    return MXD_OK;
}

static int pwrite_sink(void* user, uint64_t off, const void* data, uint64_t n) {
    const int fd = *static_cast<int*>(user);
    const char* p = static_cast<const char*>(data);
    uint64_t done = 0;
    while (done < n) {
        ssize_t w = pwrite(fd, p + done, n - done, (off_t)(off + done));
        if (w < 0) { if (errno == EINTR) continue; return 1; }
        done += (uint64_t)w;
    }
    return 0;
}

int mxc_push_local_tree(mxd_ctx* ctx, const char* basedir, const char* configfile, const char* basepath,
                        const char* repository, const char* version, char** report_json) {
    if (!ctx || !basedir || !configfile || !basepath || !repository || !version || !report_json)
        return fail(MXD_ERR_INVALID, "push_local_tree: null argument");
    Manifest m;
    int rc = parse_manifest(basedir, configfile, &m);
    if (rc != MXD_OK) return rc;
    if (m.config.name.empty()) return fail(MXD_ERR_IO, "stat " + join(basedir, configfile) + ": no such file or directory");
    std::vector<Descriptor*> all;
    for (auto& b : m.blobs) {
        if (b.mediaType == kMediaTypeModelDirectoryTarGz)
            return fail(MXC_ERR_UNSUPPORTED, "directory blob '" + b.name + "': tar+gzip is outside the digest path (DESIGN.md section 8)");
        all.push_back(&b);
    }
    all.push_back(&m.config);
    const std::string incoming_dir = join(join(join(basepath, repository), "blobs"), "sha256");
    rc = mkdir_all(incoming_dir, 0755);
    if (rc != MXD_OK) return rc;
    std::string blobs = "[";
    for (size_t i = 0; i < all.size(); ++i) {
        Descriptor& d = *all[i];
        const std::string src = join(basedir, d.name);
        rc = push_file_fill(src, &d);
        if (rc != MXD_OK) return rc;
        const std::string tmp = join(incoming_dir, ".incoming-" + std::to_string((long)getpid()) + "-" + std::to_string(i));
        int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
        if (fd < 0) return fail_errno("open " + tmp);
        uint8_t root[32]; uint64_t nch = 0, sz = 0;
        rc = mxd_tree_digest_file_tee(ctx, src.c_str(), nullptr, nullptr, 0, &nch, &sz, root, pwrite_sink, &fd);   // one read: GPU + store
        close(fd);
        if (rc != MXD_OK) { unlink(tmp.c_str()); return fail(rc, std::string("tree digest: ") + mxd_last_error()); }
        d.digest = digest_str(root);
        d.annotations[kTreeAnnotation] = std::string(kTreeMode) + ";leaf=16384;fanout=8;chunk=8388608;chunks=" + std::to_string(nch);
        std::string status, rel;
        rc = blob_digest_path(repository, d.digest, &rel);
        if (rc != MXD_OK) { unlink(tmp.c_str()); return rc; }
        if (sz == 0) { unlink(tmp.c_str()); status = "empty"; }                   // like EmptyFileDigiest: nothing to upload
        else {
            const int ex = fs_exists(basepath, rel);
            if (ex < 0) { unlink(tmp.c_str()); return ex; }
            if (ex) { unlink(tmp.c_str()); status = "exists"; }
            else {
                const std::string datafile = join(basepath, rel);
                rc = write_file(datafile + ".meta", meta_json("application/octet-stream", (int64_t)sz), 0644);
                if (rc == MXD_OK && rename(tmp.c_str(), datafile.c_str()) != 0) rc = fail_errno("rename " + tmp);
                if (rc != MXD_OK) { unlink(tmp.c_str()); return rc; }
                status = "done";
            }
        }
        if (i) blobs += ',';
        blobs += "{\"name\":"; json_string(blobs, d.name); blobs += ",\"status\":"; json_string(blobs, status);
        blobs += ",\"digest\":"; json_string(blobs, d.digest); blobs += '}';
    }
    blobs += ']';
    const std::string mj = json_manifest(m);
    rc = fs_put_file(basepath, manifest_path(repository, version), kMediaTypeModelManifestJson, (int64_t)mj.size(), nullptr, &mj);
    if (rc != MXD_OK) return rc;
    *report_json = dup_out("{\"manifest\":" + mj + ",\"blobs\":" + blobs + "}");
    return MXD_OK;
}

int mxc_pull_local(mxd_ctx* ctx, const char* basepath, const char* repository, const char* version, const char* into,
                   char** report_json) {
    if (!ctx || !basepath || !repository || !version || !into || !report_json) return fail(MXD_ERR_INVALID, "pull_local: null argument");
    struct stat st;                                                                  // pull.go:20-32
    if (stat(into, &st) != 0) {
        if (errno != ENOENT) return fail_errno(std::string("stat ") + into);
        int rc = mkdir_all(into, 0755); if (rc != MXD_OK) return rc;
    } else if (!S_ISDIR(st.st_mode)) return fail(MXD_ERR_IO, std::string(into) + " is not a directory");
    std::string text;
    int rc = read_file(join(basepath, manifest_path(repository, version)), &text);
    if (rc != MXD_OK) return rc;
    Manifest m; std::string err;
    if (!manifest_from_json(text.c_str(), &m, &err)) return fail(MXC_ERR_MANIFEST, "manifest invalid: " + err);
    std::vector<Descriptor> descs = m.blobs; descs.push_back(m.config);
    std::vector<PullState> states;
    rc = pull_check(ctx, into, descs, &states);
    if (rc != MXD_OK) return rc;
    std::string o = "[";
    for (size_t i = 0; i < descs.size(); ++i) {
        std::string status = states[i].state;
        if (status != "already exists") {
            const std::string dst = join(into, descs[i].name);
            mode_t perm = (mode_t)(descs[i].mode & 0777); if (perm == 0) perm = 0644;   // OpenWriteFile, pull.go:65-73
            rc = mkdir_all(dir_of(dst), 0777); if (rc != MXD_OK) return rc;
            if (status == "empty") { rc = write_file(dst, "", perm); if (rc != MXD_OK) return rc; }
            else {
                std::string rel; rc = blob_digest_path(repository, descs[i].digest, &rel); if (rc != MXD_OK) return rc;
                const std::string src = join(basepath, rel);
                if (access(src.c_str(), R_OK) != 0) return fail(MXC_ERR_NOT_FOUND, "blob not found: " + descs[i].digest);
                rc = copy_file(src, dst, perm, nullptr); if (rc != MXD_OK) return rc;
                chmod(dst.c_str(), perm);
                status = "done";
            }
        }
        if (i) o += ',';
        o += "{\"name\":"; json_string(o, descs[i].name); o += ",\"status\":"; json_string(o, status); o += '}';
    }
    o += ']';
    *report_json = dup_out(o);
    return MXD_OK;
}

}  // extern "C"

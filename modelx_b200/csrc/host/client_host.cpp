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
#include <atomic>
#include <functional>
#include <mutex>
#include <condition_variable>
#include <zlib.h>
#include <memory>

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

// A blob name is one path element of the model directory (ParseManifest only lists top-level entries,
// push.go:71-96); a manifest that says otherwise ("../x", "a/b", "") must not be joined onto a local path.
bool valid_entry_name(const std::string& n) {
    return !n.empty() && n != "." && n != ".." && n.find('/') == std::string::npos && n.find('\0') == std::string::npos;
}
bool valid_digest(const std::string& d) { uint8_t tmp[32]; return mxd_digest_parse(d.c_str(), tmp) == MXD_OK; }   // BlobDigestFun, registry.go:218-227

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

// Manifest-supplied strings are not trusted (the server validates digests, registry.go:218-227; names are never
// validated by the reference, which joins them onto the target directory, pull.go:113).
bool validate_manifest(const Manifest& m, std::string* err) {
    auto check = [&](const Descriptor& d, bool is_config) {
        if (is_config && d.name.empty() && d.digest.empty()) return true;      // manifest without a config descriptor
        if (!valid_entry_name(d.name)) { *err = "descriptor name '" + d.name + "' is not a single path element"; return false; }
        if (!d.digest.empty() && !valid_digest(d.digest)) { *err = "descriptor '" + d.name + "': digest invalid: " + d.digest; return false; }
        return true;
    };
    for (auto& b : m.blobs) if (!check(b, false)) return false;
    return check(m.config, true);
}

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
int read_file(const std::string& path, std::string* out) {
    int fd = open(path.c_str(), O_RDONLY | O_CLOEXEC);
    if (fd < 0) return errno == ENOENT ? fail(MXC_ERR_NOT_FOUND, "not found: " + path) : fail_errno("open " + path);
    char buf[65536];
    for (;;) { ssize_t r = read(fd, buf, sizeof buf); if (r < 0) { if (errno == EINTR) continue; int e = errno; close(fd); errno = e; return fail_errno("read " + path); } if (r == 0) break; out->append(buf, (size_t)r); }
    close(fd);
    return MXD_OK;
}
int pwrite_all(int fd, const void* data, uint64_t n, uint64_t off) {
    const char* p = static_cast<const char*>(data);
    uint64_t done = 0;
    while (done < n) {
        ssize_t w = pwrite(fd, p + done, n - done, (off_t)(off + done));
        if (w < 0) { if (errno == EINTR) continue; return -1; }
        done += (uint64_t)w;
    }
    return 0;
}

// =====================================================================================================================
// Directory blobs: tar + gzip (pkg/client/helper.go:24-53 TGZ, :55-83 UnTGZ).
// The reference archives with mholt/archiver (Tar + klauspost gzip) and clears owner and time attributes
// (ClearAttributes: true), so a directory always packs to the same bytes; its digest is computed while the archive
// is written (io.MultiWriter into the digester, helper.go:46-50).  Here: a ustar writer with zeroed owner / times,
// zlib's gzip (level 6, header mtime 0), and the digest from the GPU hasher (mxd_hasher_*, the hash.Hash seam) fed
// by the same writes.  Compressed bytes differ from the Go compressor's -- any gzip stream extracts the same -- so a
// directory digest is comparable between runs of this client, not with the stock client's.
// =====================================================================================================================
struct TgzOut {
    int fd = -1;                  // optional archive file
    mxd_hasher* hasher = nullptr;
    z_stream z{};
    bool z_open = false;
    uint64_t written = 0;
    std::vector<uint8_t> obuf = std::vector<uint8_t>(1 << 20);
    int rc = MXD_OK;
    int flush_out(size_t n) {
        if (n == 0) return MXD_OK;
        if (fd >= 0) {
            size_t off = 0;
            while (off < n) { ssize_t w = write(fd, obuf.data() + off, n - off); if (w < 0) { if (errno == EINTR) continue; return fail_errno("write archive"); } off += (size_t)w; }
        }
        int r = mxd_hasher_write(hasher, obuf.data(), n);
        if (r != MXD_OK) return fail(r, std::string("hasher: ") + mxd_last_error());
        written += n;
        return MXD_OK;
    }
    int begin() {
        if (deflateInit2(&z, 6, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return fail(MXD_ERR_NOMEM, "deflateInit2");
        z_open = true;
        return MXD_OK;
    }
    int put(const void* data, size_t n, int flush = Z_NO_FLUSH) {
        z.next_in = (Bytef*)data; z.avail_in = (uInt)n;
        do {
            z.next_out = obuf.data(); z.avail_out = (uInt)obuf.size();
            int zr = deflate(&z, flush);
            if (zr == Z_STREAM_ERROR) return fail(MXD_ERR_INVALID, "deflate");
            int r = flush_out(obuf.size() - z.avail_out);
            if (r != MXD_OK) return r;
            if (flush == Z_FINISH && zr == Z_STREAM_END) break;
        } while (z.avail_in > 0 || z.avail_out == 0);
        return MXD_OK;
    }
    ~TgzOut() { if (z_open) deflateEnd(&z); }
};

void tar_octal(char* dst, size_t width, uint64_t v) { snprintf(dst, width, "%0*llo", (int)width - 1, (unsigned long long)v); }
int tar_header(TgzOut& out, const std::string& name, char type, uint64_t size, uint32_t mode) {
    char h[512]; memset(h, 0, sizeof h);
    std::string nm = name;
    if (nm.size() > 100) {      // GNU long name record, understood by Go's archive/tar and every tar
        int r = tar_header(out, "././@LongLink", 'L', nm.size() + 1, 0644);
        if (r != MXD_OK) return r;
        std::string body = nm; body.push_back('\0');
        body.resize((body.size() + 511) / 512 * 512, '\0');
        if ((r = out.put(body.data(), body.size())) != MXD_OK) return r;
        nm.resize(100);
    }
    memcpy(h, nm.data(), nm.size());
    tar_octal(h + 100, 8, mode & 07777);
    tar_octal(h + 108, 8, 0); tar_octal(h + 116, 8, 0);          // uid, gid cleared
    tar_octal(h + 124, 12, size);
    tar_octal(h + 136, 12, 0);                                   // mtime cleared
    memset(h + 148, ' ', 8);
    h[156] = type;
    memcpy(h + 257, "ustar", 6); memcpy(h + 263, "00", 2);
    unsigned sum = 0; for (int i = 0; i < 512; ++i) sum += (unsigned char)h[i];
    snprintf(h + 148, 8, "%06o", sum); h[154] = '\0'; h[155] = ' ';
    return out.put(h, 512);
}
int tar_walk(TgzOut& out, const std::string& root, const std::string& rel) {
    const std::string dir = rel.empty() ? root : join(root, rel);
    DIR* d = opendir(dir.c_str());
    if (!d) return fail_errno("open " + dir);
    std::vector<std::string> names;
    while (struct dirent* de = readdir(d)) { std::string n = de->d_name; if (n != "." && n != "..") names.push_back(n); }
    closedir(d);
    std::sort(names.begin(), names.end());                       // filepath.WalkDir order
    for (auto& n : names) {
        const std::string r = rel.empty() ? n : rel + "/" + n, full = join(root, r);
        struct stat st;
        if (lstat(full.c_str(), &st) != 0) return fail_errno("lstat " + full);
        int rc;
        if (S_ISDIR(st.st_mode)) {
            if ((rc = tar_header(out, r + "/", '5', 0, st.st_mode)) != MXD_OK) return rc;
            if ((rc = tar_walk(out, root, r)) != MXD_OK) return rc;
        } else if (S_ISREG(st.st_mode)) {
            if ((rc = tar_header(out, r, '0', (uint64_t)st.st_size, st.st_mode)) != MXD_OK) return rc;
            int fd = open(full.c_str(), O_RDONLY | O_CLOEXEC);
            if (fd < 0) return fail_errno("open " + full);
            std::vector<char> buf(1 << 20);
            uint64_t left = (uint64_t)st.st_size;
            while (left) {
                ssize_t g = read(fd, buf.data(), std::min<uint64_t>(buf.size(), left));
                if (g < 0) { if (errno == EINTR) continue; int e = errno; close(fd); errno = e; return fail_errno("read " + full); }
                if (g == 0) { close(fd); return fail(MXD_ERR_IO, "read " + full + ": file shrank while archiving"); }
                if ((rc = out.put(buf.data(), (size_t)g)) != MXD_OK) { close(fd); return rc; }
                left -= (uint64_t)g;
            }
            close(fd);
            const size_t pad = (512 - (size_t)(st.st_size % 512)) % 512;
            if (pad) { char z[512] = {0}; if ((rc = out.put(z, pad)) != MXD_OK) return rc; }
        } else if (S_ISLNK(st.st_mode)) {
            char target[4096]; ssize_t tl = readlink(full.c_str(), target, sizeof target - 1);
            if (tl < 0) return fail_errno("readlink " + full);
            target[tl] = 0;
            char h[512]; (void)h;
            // symlinks: header with linkname (<= 100 bytes; longer targets are rejected rather than truncated)
            if (tl > 100) return fail(MXC_ERR_UNSUPPORTED, "symlink target longer than 100 bytes: " + full);
            TgzOut& o = out;
            char hdr[512]; memset(hdr, 0, sizeof hdr);
            std::string nm = r; if (nm.size() > 100) return fail(MXC_ERR_UNSUPPORTED, "symlink with a name longer than 100 bytes: " + full);
            memcpy(hdr, nm.data(), nm.size());
            tar_octal(hdr + 100, 8, 0777); tar_octal(hdr + 108, 8, 0); tar_octal(hdr + 116, 8, 0);
            tar_octal(hdr + 124, 12, 0); tar_octal(hdr + 136, 12, 0);
            memset(hdr + 148, ' ', 8); hdr[156] = '2'; memcpy(hdr + 157, target, (size_t)tl);
            memcpy(hdr + 257, "ustar", 6); memcpy(hdr + 263, "00", 2);
            unsigned sum = 0; for (int i = 0; i < 512; ++i) sum += (unsigned char)hdr[i];
            snprintf(hdr + 148, 8, "%06o", sum); hdr[154] = '\0'; hdr[155] = ' ';
            if ((rc = o.put(hdr, 512)) != MXD_OK) return rc;
        }   // sockets, devices, fifos are skipped
    }
    return MXD_OK;
}

// TGZ(ctx, dir, intofile), helper.go:24-53: archive `dir`'s contents, optionally into a file, digest on the fly.
int tgz_directory(mxd_ctx* ctx, const std::string& dir, const std::string& intofile, std::string* digest, uint64_t* archive_size) {
    struct stat st;
    if (stat(dir.c_str(), &st) != 0) return fail_errno("stat " + dir);
    if (!S_ISDIR(st.st_mode)) return fail(MXD_ERR_IO, dir + " is not a directory");
    TgzOut out;
    if (!intofile.empty()) {
        int rc = mkdir_all(dir_of(intofile), 0755);
        if (rc != MXD_OK) return rc;
        out.fd = open(intofile.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
        if (out.fd < 0) return fail_errno("create " + intofile);
    }
    int rc = mxd_hasher_new(ctx, &out.hasher);
    if (rc != MXD_OK) { if (out.fd >= 0) close(out.fd); return fail(rc, std::string("hasher: ") + mxd_last_error()); }
    rc = out.begin();
    if (rc == MXD_OK) rc = tar_walk(out, dir, "");
    if (rc == MXD_OK) { char end[1024] = {0}; rc = out.put(end, sizeof end, Z_FINISH); }
    uint8_t d[32];
    if (rc == MXD_OK) { int r = mxd_hasher_sum(out.hasher, d); if (r != MXD_OK) rc = fail(r, std::string("hasher: ") + mxd_last_error()); }
    mxd_hasher_free(out.hasher);
    if (out.fd >= 0) close(out.fd);
    if (rc != MXD_OK) { if (!intofile.empty()) unlink(intofile.c_str()); return rc; }
    *digest = digest_str(d);
    if (archive_size) *archive_size = out.written;
    return MXD_OK;
}

// UnTGZ(ctx, intodir, reader), helper.go:55-83: directories with their mode, regular files created/truncated with
// their mode.  Entry names that would leave `intodir` are refused.
int untgz_file(const std::string& archive, const std::string& intodir) {
    gzFile gz = gzopen(archive.c_str(), "rb");
    if (!gz) return fail_errno("open " + archive);
    gzbuffer(gz, 1 << 20);
    int rc = mkdir_all(intodir, 0755);
    std::string longname;
    char h[512];
    auto rd = [&](void* dst, size_t n) -> bool { size_t got = 0; while (got < n) { int r = gzread(gz, (char*)dst + got, (unsigned)(n - got)); if (r <= 0) return false; got += (size_t)r; } return true; };
    while (rc == MXD_OK) {
        if (!rd(h, 512)) { rc = fail(MXD_ERR_IO, "untgz: truncated archive " + archive); break; }
        bool allzero = true; for (int i = 0; i < 512; ++i) if (h[i]) { allzero = false; break; }
        if (allzero) break;
        const uint64_t size = strtoull(std::string(h + 124, 12).c_str(), nullptr, 8);
        const uint32_t mode = (uint32_t)strtoul(std::string(h + 100, 8).c_str(), nullptr, 8);
        const char type = h[156];
        std::string name = longname.empty() ? std::string(h, strnlen(h, 100)) : longname;
        if (longname.empty() && h[345]) name = std::string(h + 345, strnlen(h + 345, 155)) + "/" + name;   // ustar prefix
        longname.clear();
        const uint64_t padded = (size + 511) / 512 * 512;
        if (type == 'L' || type == 'x' || type == 'g') {          // GNU long name / pax headers
            if (size > (1u << 20)) { rc = fail(MXC_ERR_MANIFEST, "untgz: extended header of " + std::to_string(size) + " bytes"); break; }
            std::string body(padded, '\0');
            if (padded && !rd(&body[0], padded)) { rc = fail(MXD_ERR_IO, "untgz: truncated archive"); break; }
            body.resize(size);
            if (type == 'L') longname = body.c_str();
            else if (type == 'x') {                                // "len path=value\n" records
                size_t p = 0;
                while (p < body.size()) {
                    size_t sp = body.find(' ', p); if (sp == std::string::npos) break;
                    size_t len = strtoul(body.substr(p, sp - p).c_str(), nullptr, 10); if (len == 0 || p + len > body.size()) break;
                    std::string rec = body.substr(sp + 1, p + len - sp - 2);
                    if (rec.compare(0, 5, "path=") == 0) longname = rec.substr(5);
                    p += len;
                }
            }
            continue;
        }
        while (!name.empty() && name.back() == '/') name.pop_back();
        bool bad = name.empty() || name[0] == '/';
        { size_t p = 0; while (!bad && p <= name.size()) { size_t q = name.find('/', p); if (q == std::string::npos) q = name.size(); if (name.substr(p, q - p) == "..") bad = true; p = q + 1; } }
        if (bad && !(name.empty() && type == '5')) { rc = fail(MXC_ERR_MANIFEST, "untgz: entry '" + name + "' would leave the target directory"); break; }
        const std::string dst = name.empty() ? intodir : join(intodir, name);
        if (type == '5') {
            rc = mkdir_all(dst, 0755);
            if (rc == MXD_OK && !name.empty()) chmod(dst.c_str(), mode & 07777);
        } else if (type == '0' || type == '\0') {
            rc = mkdir_all(dir_of(dst), 0755);
            int fd = rc == MXD_OK ? open(dst.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, mode & 07777) : -1;
            if (rc == MXD_OK && fd < 0) rc = fail_errno("create " + dst);
            std::vector<char> buf(1 << 20);
            uint64_t left = size;
            while (rc == MXD_OK && left) {
                const size_t n = (size_t)std::min<uint64_t>(buf.size(), left);
                if (!rd(buf.data(), n)) { rc = fail(MXD_ERR_IO, "untgz: truncated archive"); break; }
                if (pwrite_all(fd, buf.data(), n, size - left) != 0) { rc = fail_errno("write " + dst); break; }
                left -= n;
            }
            if (fd >= 0) { fchmod(fd, mode & 07777); close(fd); }
            if (rc == MXD_OK && padded > size) { char pad[512]; if (!rd(pad, padded - size)) rc = fail(MXD_ERR_IO, "untgz: truncated archive"); }
        } else if (type == '2') {
            const std::string target(h + 157, strnlen(h + 157, 100));
            // a link may only point inside its own subtree: an absolute or ../ target would let a later entry write
            // through it to a place outside intodir
            bool tbad = target.empty() || target[0] == '/';
            { size_t p2 = 0; while (!tbad && p2 <= target.size()) { size_t q = target.find('/', p2); if (q == std::string::npos) q = target.size(); if (target.substr(p2, q - p2) == "..") tbad = true; p2 = q + 1; } }
            if (tbad) { rc = fail(MXC_ERR_MANIFEST, "untgz: symlink '" + name + "' -> '" + target + "' would leave the target directory"); break; }
            rc = mkdir_all(dir_of(dst), 0755);
            unlink(dst.c_str());
            if (rc == MXD_OK && symlink(target.c_str(), dst.c_str()) != 0) rc = fail_errno("symlink " + dst);
        } else {                                                  // other entry types: skip the body
            std::vector<char> skip(64 << 10);
            for (uint64_t left = padded; rc == MXD_OK && left;) {
                const size_t n = (size_t)std::min<uint64_t>(skip.size(), left);
                if (!rd(skip.data(), n)) rc = fail(MXD_ERR_IO, "untgz: truncated archive");
                left -= n;
            }
        }
    }
    gzclose(gz);
    return rc;
}

// ---- opt-in digest cache: <basedir>/.modelx/digests.json = {"<name>":{"size":..,"mtime_ns":..,"digest":".."}} ----
struct CacheEntry { int64_t size = 0; int64_t mtime_ns = 0; std::string digest; };
std::string cache_path(const std::string& basedir) { return join(join(basedir, ".modelx"), "digests.json"); }
void cache_load(const std::string& basedir, std::map<std::string, CacheEntry>* out) {
    std::string text;
    if (read_file(cache_path(basedir), &text) != MXD_OK) return;
    JParser jp{text.c_str(), text.c_str() + text.size(), ""};
    JVal root;
    if (!jp.val(root) || root.kind != JVal::Obj) return;        // unreadable cache = no cache
    for (auto& kv : root.obj) {
        CacheEntry e;
        if (auto* x = kv.second.get("size")) e.size = strtoll(x->str.c_str(), nullptr, 10);
        if (auto* x = kv.second.get("mtime_ns")) e.mtime_ns = strtoll(x->str.c_str(), nullptr, 10);
        if (auto* x = kv.second.get("digest")) e.digest = x->str;
        if (valid_digest(e.digest)) (*out)[kv.first] = e;
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
int64_t mtime_ns(const struct stat& st) { return (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec; }

const char* kTreeMode = "tree.v1";
const char* kTreeAnnotation = "modelx.digest";
bool is_tree_keyed(const Descriptor& d) {
    auto it = d.annotations.find(kTreeAnnotation);
    return it != d.annotations.end() && it->second.compare(0, strlen(kTreeMode), kTreeMode) == 0;
}
std::string tree_annotation(uint64_t nchunks) { return std::string(kTreeMode) + ";leaf=16384;fanout=8;chunk=8388608;chunks=" + std::to_string(nchunks); }

// One blob of a push as the pipeline sees it: where its bytes are and which descriptor it fills.
struct PushItem {
    Descriptor* desc = nullptr;
    std::string path;               // file to hash / upload (for directory blobs: the .modelx/<name>.tar.gz just written)
    bool is_dir = false;
    struct stat st{};
};

// ParseManifest, then per blob the stat + (for directories) tar.gz half of pushDirectory / pushFile (push.go:102-147).
// Directory blobs are packed to <basedir>/.modelx/<name>.tar.gz with their digest taken while the archive is written.
int push_prepare(mxd_ctx* ctx, const std::string& basedir, const std::string& configfile, Manifest* m, std::vector<PushItem>* items) {
    int rc = parse_manifest(basedir, configfile, m);
    if (rc != MXD_OK) return rc;
    if (m->config.name.empty()) return fail(MXD_ERR_IO, "stat " + join(basedir, configfile) + ": no such file or directory");
    std::vector<Descriptor*> all;
    for (auto& b : m->blobs) all.push_back(&b);
    all.push_back(&m->config);
    for (Descriptor* d : all) {
        PushItem it; it.desc = d;
        const std::string src = join(basedir, d->name);
        if (d->mediaType == kMediaTypeModelDirectoryTarGz) {
            struct stat ds;
            if (stat(src.c_str(), &ds) != 0) return fail_errno("stat " + src);
            d->mode = go_file_mode(ds.st_mode);                                   // push.go:107-108
            d->modified = go_time_json(ds.st_mtim);
            it.is_dir = true;
            it.path = join(join(basedir, ".modelx"), d->name + ".tar.gz");       // push.go:111
            uint64_t asz = 0;
            rc = tgz_directory(ctx, src, it.path, &d->digest, &asz);
            if (rc != MXD_OK) return rc;
        } else {
            it.path = src;
        }
        if (stat(it.path.c_str(), &it.st) != 0) return fail_errno("stat " + it.path);
        if (S_ISDIR(it.st.st_mode)) return fail(MXD_ERR_IO, "read " + it.path + ": is a directory");
        items->push_back(std::move(it));
    }
    return MXD_OK;
}

// Digest phase of Client.Push, push.go:29-52: every blob + the config through pushFile's "stat, digest"
// (push.go:120-142); the whole-file digests come from ONE coalesced GPU batch (the reference: 3 goroutines).
int push_digest(mxd_ctx* ctx, const std::string& basedir, const std::string& configfile, int flags, Manifest* m) {
    const bool with_tree = (flags & MXC_PUSH_TREE) != 0, use_cache = (flags & MXC_PUSH_CACHE) != 0;
    std::vector<PushItem> items;
    int rc = push_prepare(ctx, basedir, configfile, m, &items);
    if (rc != MXD_OK) return rc;
    std::map<std::string, CacheEntry> cache;
    if (use_cache) cache_load(basedir, &cache);
    std::vector<size_t> todo;                         // items that really need hashing
    for (size_t i = 0; i < items.size(); ++i) {
        PushItem& it = items[i];
        if (!it.desc->digest.empty()) continue;       // directory blob: digest taken while archiving
        auto c = cache.find(it.desc->name);
        if (use_cache && c != cache.end() && c->second.size == (int64_t)it.st.st_size && c->second.mtime_ns == mtime_ns(it.st))
            it.desc->digest = c->second.digest;
        else todo.push_back(i);
    }
    if (!todo.empty()) {
        std::vector<const char*> cpaths;
        for (size_t i : todo) cpaths.push_back(items[i].path.c_str());
        std::vector<uint8_t> out(32 * todo.size());
        rc = mxd_sha256_files(ctx, cpaths.data(), todo.size(), out.data(), nullptr);
        if (rc != MXD_OK) return fail(rc, std::string("digest: ") + mxd_last_error());
        for (size_t k = 0; k < todo.size(); ++k) items[todo[k]].desc->digest = digest_str(&out[32 * k]);
    }
    for (auto& it : items) {
        rc = push_file_fill(it.path, it.desc);
        if (rc != MXD_OK) return rc;
        if (use_cache && !it.is_dir) {
            CacheEntry e; e.size = (int64_t)it.st.st_size; e.mtime_ns = mtime_ns(it.st); e.digest = it.desc->digest;
            cache[it.desc->name] = e;
        }
        if (with_tree) {
            uint8_t root[32]; uint64_t nch = 0, sz = 0;
            rc = mxd_tree_digest_file(ctx, it.path.c_str(), nullptr, nullptr, 0, &nch, &sz, root);
            if (rc != MXD_OK) return fail(rc, std::string("tree digest: ") + mxd_last_error());
            it.desc->annotations["modelx.tree.v1"] = digest_str(root) + ";leaf=16384;fanout=8;chunk=8388608;chunks=" + std::to_string(nch);
        }
    }
    if (use_cache) cache_store(basedir, cache);
    return MXD_OK;
}

// =====================================================================================================================
// pkg/registry local FS store
// =====================================================================================================================
// BlobDigestPath, store.go:56-61: path.Join(repository, "blobs", algorithm, hex).  The digest is validated first
// (BlobDigestFun, registry.go:218-227): anything but sha256:<64 lower hex> could name a file outside the store.
int blob_digest_path(const std::string& repo, const std::string& digest, std::string* out) {
    if (!valid_digest(digest)) return fail(MXC_ERR_DIGEST_INVALID, "digest invalid: " + digest);
    const size_t colon = digest.find(':');
    *out = join(join(join(repo, "blobs"), digest.substr(0, colon)), digest.substr(colon + 1));
    return MXD_OK;
}
bool valid_repository(const std::string& repo) {       // <project>/<name>: relative, no empty / dot components
    if (repo.empty() || repo[0] == '/') return false;
    size_t p = 0;
    while (p <= repo.size()) {
        size_t q = repo.find('/', p); if (q == std::string::npos) q = repo.size();
        const std::string part = repo.substr(p, q - p);
        if (part.empty() || part == "." || part == "..") return false;
        p = q + 1;
    }
    return true;
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
// LocalFSProvider.Put (fs_local.go:46-51) for small inline content: meta first, then data
int fs_put_inline(const std::string& basepath, const std::string& rel, const std::string& content_type, const std::string& data) {
    const std::string datafile = join(basepath, rel), metafile = datafile + ".meta";
    int rc = mkdir_all(dir_of(metafile), 0755);
    if (rc != MXD_OK) return rc;
    rc = write_file(metafile, meta_json(content_type, (int64_t)data.size()), 0644);
    if (rc != MXD_OK) return rc;
    return write_file(datafile, data, 0644);
}
int fs_exists(const std::string& basepath, const std::string& rel) {   // LocalFSProvider.Exists, fs_local.go:76-85
    struct stat st;
    if (stat(join(basepath, rel).c_str(), &st) == 0) return 1;
    if (errno == ENOENT || errno == ENOTDIR) return 0;
    return fail_errno("stat " + join(basepath, rel));
}
std::string manifest_path(const std::string& repo, const std::string& ref) { return join(join(repo, "manifests"), ref); }   // store.go:67-69
const char* kIndexFile = "index.json";                                  // RegistryIndexFileName, store_fs.go:21
const char* kMediaTypeModelIndexJson = "application/vnd.modelx.model.index.v1.json";

// types.Index, types.go:53-58
std::string json_index(const std::vector<Descriptor>& manifests, const std::map<std::string, std::string>* annotations) {
    std::string o = "{\"schemaVersion\":0,\"manifests\":";
    if (manifests.empty()) o += "null";
    else { o += '['; for (size_t i = 0; i < manifests.size(); ++i) { if (i) o += ','; json_descriptor(o, manifests[i]); } o += ']'; }
    if (annotations && !annotations->empty()) { o += ",\"annotations\":"; json_map(o, *annotations); }
    o += '}';
    return o;
}

// FSRegistryStore.RefreshIndex (store_fs.go:185-238) + PutIndex (:145-172) + RefreshGlobalIndex (:287-330):
// after every PutManifest the repository's index.json lists one descriptor per version {name, modified = the
// manifest file's mtime, annotations = the manifest's, size = config + blobs}, sorted by name, with the first
// non-nil annotations promoted to the index; the registry's index.json lists every repository that has one.
int fs_refresh_index(const std::string& basepath, const std::string& repo) {
    const std::string mdir = join(basepath, join(repo, "manifests"));
    DIR* d = opendir(mdir.c_str());
    if (!d) return fail_errno("open " + mdir);
    std::vector<std::string> names;
    while (struct dirent* de = readdir(d)) {
        std::string n = de->d_name;
        if (n == "." || n == ".." || (n.size() > 5 && n.compare(n.size() - 5, 5, ".meta") == 0)) continue;
        struct stat st;
        if (stat(join(mdir, n).c_str(), &st) != 0 || S_ISDIR(st.st_mode)) continue;
        names.push_back(n);
    }
    closedir(d);
    std::sort(names.begin(), names.end());
    std::vector<Descriptor> list;
    const std::map<std::string, std::string>* idx_ann = nullptr;
    std::vector<Manifest> keep(names.size());
    for (size_t i = 0; i < names.size(); ++i) {
        std::string text, err;
        int rc = read_file(join(mdir, names[i]), &text);
        if (rc != MXD_OK) return rc;
        if (!manifest_from_json(text.c_str(), &keep[i], &err)) return fail(MXC_ERR_MANIFEST, "manifest invalid: " + names[i] + ": " + err);
        struct stat st; stat(join(mdir, names[i]).c_str(), &st);
        Descriptor desc; desc.name = names[i]; desc.modified = go_time_json(st.st_mtim); desc.annotations = keep[i].annotations;
        // Go sums int64 sizes with wrap-around (store_fs.go:205-212); same bits here, without the signed-overflow UB
        uint64_t total = (uint64_t)keep[i].config.size; for (auto& b : keep[i].blobs) total += (uint64_t)b.size;
        desc.size = (int64_t)total;
        list.push_back(std::move(desc));
        if (!idx_ann && !keep[i].annotations.empty()) idx_ann = &keep[i].annotations;
    }
    if (!list.empty()) {
        int rc = fs_put_inline(basepath, join(repo, kIndexFile), kMediaTypeModelIndexJson, json_index(list, idx_ann));
        if (rc != MXD_OK) return rc;
    }
    // global index: every <repository>/index.json below the base path
    std::vector<Descriptor> repos;
    std::function<int(const std::string&)> walk = [&](const std::string& rel) -> int {
        const std::string dir = rel.empty() ? basepath : join(basepath, rel);
        DIR* dd = opendir(dir.c_str());
        if (!dd) return MXD_OK;
        std::vector<std::pair<std::string, bool>> ents;
        while (struct dirent* de = readdir(dd)) {
            std::string n = de->d_name; if (n == "." || n == "..") continue;
            struct stat st; if (stat(join(dir, n).c_str(), &st) != 0) continue;
            ents.emplace_back(n, S_ISDIR(st.st_mode));
        }
        closedir(dd);
        for (auto& e : ents) {
            if (e.second) { if (e.first != "blobs" && e.first != "manifests") { int rc = walk(rel.empty() ? e.first : rel + "/" + e.first); if (rc != MXD_OK) return rc; } continue; }
            if (e.first != kIndexFile || rel.empty()) continue;
            std::string text; if (read_file(join(dir, e.first), &text) != MXD_OK) continue;
            JParser jp{text.c_str(), text.c_str() + text.size(), ""}; JVal root;
            Descriptor desc; desc.name = rel; desc.mediaType = kMediaTypeModelIndexJson;
            if (jp.val(root) && root.kind == JVal::Obj) if (auto* x = root.get("annotations")) for (auto& kv : x->obj) desc.annotations[kv.first] = kv.second.str;
            repos.push_back(std::move(desc));
        }
        return MXD_OK;
    };
    int rc = walk("");
    if (rc != MXD_OK) return rc;
    std::sort(repos.begin(), repos.end(), [](const Descriptor& a, const Descriptor& b) { return a.name < b.name; });
    return fs_put_inline(basepath, kIndexFile, kMediaTypeModelIndexJson, json_index(repos, nullptr));
}

// FSRegistryStore.PutManifest, store_fs.go:87-104
int fs_put_manifest(const std::string& basepath, const std::string& repo, const std::string& reference, const std::string& content_type,
                    const std::string& manifest_json_text) {
    if (!valid_repository(repo)) return fail(MXD_ERR_INVALID, "repository invalid: " + repo);
    if (!valid_entry_name(reference)) return fail(MXD_ERR_INVALID, "reference invalid: " + reference);
    int rc = fs_put_inline(basepath, manifest_path(repo, reference), content_type, manifest_json_text);
    if (rc != MXD_OK) return rc;
    return fs_refresh_index(basepath, repo);
}

struct TempWriter {     // sink target: pwrite into an open temp file
    int fd = -1;
    static int sink(void* user, uint64_t off, const void* data, uint64_t n) { return pwrite_all(static_cast<TempWriter*>(user)->fd, data, n, off); }
};

// Registry.PutBlob (registry.go:144-164) -> FSRegistryStore.PutBlob (store_fs.go:358-364) -> LocalFSProvider.Put.
// The body lands in a temp file beside its final place and is renamed in only when complete, so a reader never sees
// a partial or unverified blob, and an existing (good) blob is never overwritten or deleted.
// verify != 0 (new): the body is hashed WHILE it is written -- one read of the source, the bytes are
// teed to the GPU and to the temp file -- and a mismatch with `digest` is rejected with DIGEST_INVALID before
// anything becomes visible.  verify == 1: whole-file SHA-256; 2: modelx.tree.v1 root.
int fs_put_blob(mxd_ctx* ctx, const std::string& basepath, const std::string& repo, const std::string& digest,
                const std::string& content_type, const std::string& srcfile, int verify) {
    if (!valid_repository(repo)) return fail(MXD_ERR_INVALID, "repository invalid: " + repo);
    if (content_type.empty()) return fail(MXD_ERR_INVALID, "content type invalid: empty");                                // registry.go:147-151
    std::string rel;
    int rc = blob_digest_path(repo, digest, &rel);
    if (rc != MXD_OK) return rc;
    uint8_t want[32]; mxd_digest_parse(digest.c_str(), want);
    struct stat st;
    if (stat(srcfile.c_str(), &st) != 0) return fail_errno("stat " + srcfile);
    const std::string datafile = join(basepath, rel);
    if (fs_exists(basepath, rel) == 1 && !verify) return MXD_OK;       // content-addressed: same key, same bytes
    rc = mkdir_all(dir_of(datafile), 0755);
    if (rc != MXD_OK) return rc;
    static std::atomic<uint64_t> seq{0};
    const std::string tmp = join(dir_of(datafile), ".incoming-" + std::to_string((long)getpid()) + "-" + std::to_string(seq++));
    TempWriter tw; tw.fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
    if (tw.fd < 0) return fail_errno("open " + tmp);
    uint8_t got[32]; bool match = true;
    if (verify) {
        if (!ctx) { close(tw.fd); unlink(tmp.c_str()); return fail(MXD_ERR_INVALID, "verify needs an engine context"); }
        uint64_t sz = 0, nch = 0;
        if (verify == 2) rc = mxd_tree_digest_file_tee(ctx, srcfile.c_str(), nullptr, nullptr, 0, &nch, &sz, got, TempWriter::sink, &tw);
        else { mxd_part whole{0, (int64_t)st.st_size}; rc = mxd_sha256_file_ranges(ctx, srcfile.c_str(), &whole, 1, got, &sz, TempWriter::sink, &tw); }
        if (rc != MXD_OK) { close(tw.fd); unlink(tmp.c_str()); return fail(rc, std::string("verify: ") + mxd_last_error()); }
        match = memcmp(got, want, 32) == 0;
    } else {                                                         // the reference stores the body unverified
        int in = open(srcfile.c_str(), O_RDONLY | O_CLOEXEC);
        if (in < 0) { int e = errno; close(tw.fd); unlink(tmp.c_str()); errno = e; return fail_errno("open " + srcfile); }
        std::vector<char> buf(4 << 20); uint64_t off = 0;
        for (;;) {
            ssize_t r = read(in, buf.data(), buf.size());
            if (r < 0) { if (errno == EINTR) continue; rc = fail_errno("read " + srcfile); break; }
            if (r == 0) break;
            if (pwrite_all(tw.fd, buf.data(), (uint64_t)r, off) != 0) { rc = fail_errno("write " + tmp); break; }
            off += (uint64_t)r;
        }
        close(in);
    }
    close(tw.fd);
    if (rc != MXD_OK) { unlink(tmp.c_str()); return rc; }
    if (!match) { unlink(tmp.c_str()); return fail(MXC_ERR_DIGEST_INVALID, "digest invalid: " + digest + " (body hashes to " + digest_str(got) + ")"); }
    rc = write_file(datafile + ".meta", meta_json(content_type, (int64_t)st.st_size), 0644);
    if (rc == MXD_OK && rename(tmp.c_str(), datafile.c_str()) != 0) rc = fail_errno("rename " + tmp);
    if (rc != MXD_OK) unlink(tmp.c_str());
    return rc;
}

// =====================================================================================================================
// The read-once push: every blob is read from disk ONCE.  Its bytes stream through the pinned
// ring to the GPU -- whole-file SHA-256 (the reference's identity) and, in the same rounds, the SHA-256 of every
// multipart part (calcParts over the server's part count, extension_s3.go:99-112, store_s3.go:198-203,273-279) --
// and are teed, part by part, to an uploader (S3Extension.Upload's role, extension_s3.go:52-89: at most
// `max_concurrent` writes in flight, a failed part is re-sent up to 3 times, :133-148).  The reference reads every
// file twice (push.go:160, then extension_s3.go:71-82) and sends parts unhashed.
// =====================================================================================================================
struct Semaphore {
    std::mutex mu; std::condition_variable cv; int avail;
    explicit Semaphore(int n) : avail(n) {}
    void acquire() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return avail > 0; }); --avail; }
    void release() { { std::lock_guard<std::mutex> lk(mu); ++avail; } cv.notify_one(); }
};
struct UploadBlob {
    const mxc_uploader* up = nullptr; Semaphore* sem = nullptr;
    uint64_t index = 0, size = 0;
    std::vector<mxd_part> parts;
    std::vector<std::atomic<int>> failed;          // per part: a write was refused during the pass
    std::atomic<uint64_t> delivered{0};
    explicit UploadBlob(size_t nparts) : failed(nparts) { for (auto& f : failed) f.store(0); }
    int write_part(uint64_t part, uint64_t off, const void* data, uint64_t n) {
        if (sem) sem->acquire();
        const int r = up->part_write(up->user, index, part, off, data, n);
        if (sem) sem->release();
        return r;
    }
    // the tee: split a piece at part boundaries
    static int sink(void* user, uint64_t off, const void* data, uint64_t n) {
        auto* b = static_cast<UploadBlob*>(user);
        const uint8_t* p = static_cast<const uint8_t*>(data);
        uint64_t done = 0;
        while (done < n) {
            const uint64_t at = off + done;
            size_t pi = 0;       // parts are consecutive and cover [0, size): find the one holding `at`
            { size_t lo = 0, hi = b->parts.size(); while (lo + 1 < hi) { size_t mid = (lo + hi) / 2; if ((uint64_t)b->parts[mid].offset <= at) lo = mid; else hi = mid; } pi = lo; }
            const uint64_t pend = (uint64_t)b->parts[pi].offset + (uint64_t)b->parts[pi].length;
            const uint64_t take = std::min(n - done, pend - at);
            if (!b->failed[pi].load() && b->write_part(pi, at, p + done, take) != 0) b->failed[pi].store(1);
            b->delivered += take;
            done += take;
        }
        return 0;                 // a refused part is retried after the pass; the digest pass itself goes on
    }
};

int push_stream(mxd_ctx* ctx, const std::string& basedir, const std::string& configfile, const mxc_uploader* up, int flags,
                Manifest* m, std::string* blobs_json) {
    if (!up || !up->begin || !up->part_write || !up->complete) return fail(MXD_ERR_INVALID, "push_stream: uploader needs begin, part_write and complete");
    std::vector<PushItem> items;
    int rc = push_prepare(ctx, basedir, configfile, m, &items);
    if (rc != MXD_OK) return rc;
    const bool force_multi = (flags & MXC_PUSH_FORCE_MULTIPART) != 0;
    Semaphore sem(up->max_concurrent > 0 ? up->max_concurrent : 1);
    std::vector<std::unique_ptr<UploadBlob>> ubs;
    std::vector<std::vector<mxd_part>> ranges(items.size());
    std::vector<std::vector<uint8_t>> outs(items.size());
    std::vector<mxd_file_job> jobs(items.size());
    size_t begun = 0;
    auto abort_all = [&] { if (up->abort) for (size_t i = 0; i < begun; ++i) up->abort(up->user, i); };
    for (size_t i = 0; i < items.size(); ++i) {
        const uint64_t size = (uint64_t)items[i].st.st_size;
        const int64_t np = size ? mxd_server_part_count((int64_t)size, force_multi ? 1 : 0) : 1;
        auto ub = std::unique_ptr<UploadBlob>(new UploadBlob((size_t)np));
        ub->up = up; ub->sem = up->max_concurrent > 0 ? &sem : nullptr; ub->index = i; ub->size = size;
        ub->parts.resize((size_t)np);
        mxd_calc_parts((int64_t)size, np, ub->parts.data());
        ranges[i].push_back(mxd_part{0, (int64_t)size});                  // chain 0: the blob's content address
        if (np > 1) for (auto& p : ub->parts) ranges[i].push_back(p);     // chains 1..np: per-part SHA-256, same pass
        outs[i].resize(32 * ranges[i].size());
        rc = up->begin(up->user, i, items[i].desc->name.c_str(), size, ub->parts.data(), (uint64_t)np);
        if (rc != 0) { abort_all(); return fail(MXD_ERR_IO, "uploader refused blob '" + items[i].desc->name + "'"); }
        ++begun;
        jobs[i] = mxd_file_job{};
        jobs[i].path = items[i].path.c_str(); jobs[i].ranges = ranges[i].data(); jobs[i].nranges = ranges[i].size();
        jobs[i].out = outs[i].data(); jobs[i].sink = UploadBlob::sink; jobs[i].sink_user = ub.get();
        ubs.push_back(std::move(ub));
    }
    rc = mxd_sha256_file_jobs(ctx, jobs.data(), jobs.size());            // ONE pass over every file: digests + tee
    if (rc != MXD_OK) { abort_all(); return fail(rc, std::string("digest: ") + mxd_last_error()); }
    // retry refused parts from the file (extension_s3.go:133-148: up to 3 attempts per part)
    uint64_t reread = 0;
    for (size_t i = 0; i < items.size(); ++i) {
        UploadBlob& b = *ubs[i];
        for (size_t pi = 0; pi < b.parts.size(); ++pi) {
            if (!b.failed[pi].load()) continue;
            bool ok = false;
            for (int attempt = 2; attempt <= 3 && !ok; ++attempt) {
                if (up->part_restart && up->part_restart(up->user, i, pi) != 0) continue;
                int fd = open(items[i].path.c_str(), O_RDONLY | O_CLOEXEC);
                if (fd < 0) { abort_all(); return fail_errno("open " + items[i].path); }
                std::vector<char> buf(4 << 20);
                uint64_t off = (uint64_t)b.parts[pi].offset; const uint64_t end = off + (uint64_t)b.parts[pi].length;
                ok = true;
                while (off < end && ok) {
                    ssize_t r = pread(fd, buf.data(), std::min<uint64_t>(buf.size(), end - off), (off_t)off);
                    if (r <= 0) { if (r < 0 && errno == EINTR) continue; ok = false; break; }
                    reread += (uint64_t)r;
                    if (b.write_part(pi, off, buf.data(), (uint64_t)r) != 0) ok = false;
                    off += (uint64_t)r;
                }
                close(fd);
            }
            if (!ok) { abort_all(); return fail(MXD_ERR_IO, "upload of part " + std::to_string(pi) + " of '" + items[i].desc->name + "' failed 3 times"); }
        }
    }
    std::string bj = "[";
    for (size_t i = 0; i < items.size(); ++i) {
        Descriptor& d = *items[i].desc;
        const std::string dg = digest_str(outs[i].data());
        if (items[i].is_dir && d.digest != dg) { abort_all(); return fail(MXC_ERR_DIGEST_INVALID, "archive of '" + d.name + "' changed while it was pushed"); }
        d.digest = dg;
        rc = push_file_fill(items[i].path, &d);
        if (rc != MXD_OK) { abort_all(); return rc; }
        const size_t np = ubs[i]->parts.size();
        const uint8_t* pd = np > 1 ? outs[i].data() + 32 : outs[i].data();   // a single part is the whole blob
        char status[16] = "done";
        rc = up->complete(up->user, i, d.digest.c_str(), pd, np, status);
        if (rc != 0) { abort_all(); return fail(MXD_ERR_IO, "uploader could not complete blob '" + d.name + "'" + (g_err.empty() ? "" : ": " + g_err)); }
        status[15] = 0;
        if (i) bj += ',';
        bj += "{\"name\":"; json_string(bj, d.name); bj += ",\"status\":"; json_string(bj, status);
        bj += ",\"digest\":"; json_string(bj, d.digest);
        bj += ",\"size\":" + std::to_string(ubs[i]->size) + ",\"parts\":[";
        for (size_t pi = 0; pi < np; ++pi) {
            if (pi) bj += ',';
            bj += "{\"offset\":" + std::to_string(ubs[i]->parts[pi].offset) + ",\"length\":" + std::to_string(ubs[i]->parts[pi].length) + ",\"sha256\":";
            json_string(bj, digest_str(pd + 32 * pi).substr(7)); bj += '}';
        }
        bj += "]}";
    }
    bj += ']';
    *blobs_json = bj + ",\"reread_bytes\":" + std::to_string(reread);
    return MXD_OK;
}

// The uploader that is the local FS store: parts are pwritten into one temp file beside the blob's final place;
// complete() applies PushBlob's decisions (push.go:163-194) now that the digest is known: EmptyFileDigiest ->
// "empty", already stored -> "exists", else meta + rename -> "done".
struct FsUploader {
    std::string basepath, repo, dir;
    struct Blob { int fd = -1; std::string tmp; uint64_t size = 0; };
    std::vector<Blob> blobs;
    std::mutex mu;
    static int begin(void* u, uint64_t i, const char*, uint64_t size, const mxd_part*, uint64_t) {
        auto* f = static_cast<FsUploader*>(u);
        std::lock_guard<std::mutex> lk(f->mu);
        if (f->blobs.size() <= i) f->blobs.resize(i + 1);
        static std::atomic<uint64_t> seq{0};
        f->blobs[i].tmp = join(f->dir, ".incoming-" + std::to_string((long)getpid()) + "-" + std::to_string(seq++));
        f->blobs[i].size = size;
        f->blobs[i].fd = open(f->blobs[i].tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
        return f->blobs[i].fd < 0 ? -1 : 0;
    }
    static int part_write(void* u, uint64_t i, uint64_t, uint64_t off, const void* data, uint64_t n) {
        return pwrite_all(static_cast<FsUploader*>(u)->blobs[i].fd, data, n, off);
    }
    static int complete(void* u, uint64_t i, const char* digest, const uint8_t*, uint64_t, char status[16]) {
        auto* f = static_cast<FsUploader*>(u);
        Blob& b = f->blobs[i];
        if (b.fd >= 0) { close(b.fd); b.fd = -1; }
        std::string rel;
        if (blob_digest_path(f->repo, digest, &rel) != MXD_OK) { unlink(b.tmp.c_str()); return -1; }
        if (kEmptyFileDigest == std::string(digest)) { unlink(b.tmp.c_str()); snprintf(status, 16, "empty"); return 0; }     // push.go:165-168
        const int ex = fs_exists(f->basepath, rel);                                                                       // HeadBlob, push.go:169-177
        if (ex < 0) { unlink(b.tmp.c_str()); return -1; }
        if (ex) { unlink(b.tmp.c_str()); snprintf(status, 16, "exists"); return 0; }
        const std::string datafile = join(f->basepath, rel);
        // fallback upload through the server: Content-Type application/octet-stream (client/registry.go:109-120)
        int rc = write_file(datafile + ".meta", meta_json("application/octet-stream", (int64_t)b.size), 0644);
        if (rc == MXD_OK && rename(b.tmp.c_str(), datafile.c_str()) != 0) rc = fail_errno("rename " + b.tmp);
        if (rc != MXD_OK) { unlink(b.tmp.c_str()); return -1; }
        snprintf(status, 16, "done");
        return 0;
    }
    static void abort(void* u, uint64_t i) {
        auto* f = static_cast<FsUploader*>(u);
        if (i >= f->blobs.size()) return;
        if (f->blobs[i].fd >= 0) { close(f->blobs[i].fd); f->blobs[i].fd = -1; }
        if (!f->blobs[i].tmp.empty()) unlink(f->blobs[i].tmp.c_str());
    }
};

// =====================================================================================================================
// Pull
// =====================================================================================================================
// pullFile's check, pull.go:111-136 (and pullDirectory's, :146-156), for a list of descriptors.  Whole-file digests of
// all present files are one coalesced GPU batch with a status PER FILE; descriptors annotated as tree-keyed
// (mxc_push_local_tree) are checked with the tree digest instead; directories by re-archiving them (TGZ(dir, "")).
struct PullState { std::string name, digest, state; };
int pull_check(mxd_ctx* ctx, const std::string& basedir, const std::vector<Descriptor>& descs, std::vector<PullState>* out) {
    std::vector<size_t> present, tree_present;
    std::vector<std::string> paths, tree_paths;
    for (size_t i = 0; i < descs.size(); ++i) {
        const Descriptor& d = descs[i];
        out->push_back({d.name, d.digest, "missing"});
        struct stat st;
        const std::string p = join(basedir, d.name);
        const bool have = stat(p.c_str(), &st) == 0;
        if (!have && errno != ENOENT && errno != ENOTDIR) return fail_errno("open " + p);     // pull.go:125-127
        if (d.mediaType == kMediaTypeModelDirectoryTarGz) {
            if (have && S_ISDIR(st.st_mode)) {
                std::string dg; uint64_t asz = 0;
                int rc = tgz_directory(ctx, p, "", &dg, &asz);                                // pull.go:149
                if (rc != MXD_OK) return rc;
                (*out)[i].state = dg == d.digest ? "already exists" : "differs";
            } else if (have) return fail(MXD_ERR_IO, p + " is not a directory");
            continue;
        }
        if (!have) continue;
        // os.Open succeeds on a directory and digest.FromReader then fails with EISDIR (pull.go:116-119)
        if (S_ISDIR(st.st_mode)) return fail(MXD_ERR_IO, "read " + p + ": is a directory");
        if (is_tree_keyed(d)) { tree_present.push_back(i); tree_paths.push_back(p); }
        else { present.push_back(i); paths.push_back(p); }
    }
    if (!tree_present.empty()) {       // tree-keyed blobs: one pipelined pass over all of them
        std::vector<const char*> cp; for (auto& p : tree_paths) cp.push_back(p.c_str());
        std::vector<uint8_t> roots(32 * tree_present.size());
        int rc = mxd_tree_digest_files(ctx, cp.data(), cp.size(), nullptr, roots.data(), nullptr, nullptr);
        if (rc != MXD_OK) return fail(rc, std::string("tree digest: ") + mxd_last_error());
        for (size_t k = 0; k < tree_present.size(); ++k)
            (*out)[tree_present[k]].state = (digest_str(&roots[32 * k]) == descs[tree_present[k]].digest) ? "already exists" : "differs";
    }
    if (!present.empty()) {
        std::vector<mxd_file_job> jobs(present.size());
        std::vector<uint8_t> got(32 * present.size());
        for (size_t k = 0; k < present.size(); ++k) { jobs[k] = mxd_file_job{}; jobs[k].path = paths[k].c_str(); jobs[k].out = &got[32 * k]; }
        int rc = mxd_sha256_file_jobs(ctx, jobs.data(), jobs.size());
        if (rc != MXD_OK) return fail(rc, std::string("digest: ") + mxd_last_error());
        for (size_t k = 0; k < present.size(); ++k)                           // pull.go:120 string equality
            (*out)[present[k]].state = (digest_str(&got[32 * k]) == descs[present[k]].digest) ? "already exists" : "differs";
    }
    for (size_t i = 0; i < descs.size(); ++i) {
        PullState& s = (*out)[i];
        if (s.state == "already exists" || descs[i].mediaType == kMediaTypeModelDirectoryTarGz) continue;
        const bool empty = is_tree_keyed(descs[i]) ? descs[i].size == 0 : s.digest == kEmptyFileDigest;   // pull.go:134-136
        if (empty) s.state = "empty";
    }
    return MXD_OK;
}

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

int mxc_tgz(mxd_ctx* ctx, const char* dir, const char* intofile, char** digest, uint64_t* archive_size) {
    if (!ctx || !dir || !digest) return fail(MXD_ERR_INVALID, "tgz: null argument");
    std::string dg;
    int rc = tgz_directory(ctx, dir, intofile ? intofile : "", &dg, archive_size);
    if (rc != MXD_OK) return rc;
    *digest = dup_out(dg);
    return MXD_OK;
}

int mxc_untgz(const char* archive, const char* intodir) {
    if (!archive || !intodir) return fail(MXD_ERR_INVALID, "untgz: null argument");
    return untgz_file(archive, intodir);
}

int mxc_pull_check(mxd_ctx* ctx, const char* basedir, const char* manifest_json, char** report_json) {
    if (!ctx || !basedir || !manifest_json || !report_json) return fail(MXD_ERR_INVALID, "pull_check: null argument");
    Manifest m; std::string err;
    if (!manifest_from_json(manifest_json, &m, &err) || !validate_manifest(m, &err)) return fail(MXC_ERR_MANIFEST, "manifest invalid: " + err);
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
    if (!manifest_from_json(manifest_json, &m, &err) || !validate_manifest(m, &err)) return fail(MXC_ERR_MANIFEST, "manifest invalid: " + err);
    return fs_put_manifest(basepath, repository, reference, content_type ? content_type : "", json_manifest(m));   // json.Marshal(manifest), store_fs.go:88
}

int mxc_fs_get_manifest(const char* basepath, const char* repository, const char* reference, char** manifest_json) {
    if (!basepath || !repository || !reference || !manifest_json) return fail(MXD_ERR_INVALID, "fs_get_manifest: null argument");
    if (!valid_repository(repository) || !valid_entry_name(reference)) return fail(MXD_ERR_INVALID, "repository or reference invalid");
    std::string text;
    int rc = read_file(join(basepath, manifest_path(repository, reference)), &text);
    if (rc != MXD_OK) return rc;
    *manifest_json = dup_out(text);
    return MXD_OK;
}

int mxc_fs_get_index(const char* basepath, const char* repository, char** index_json) {
    if (!basepath || !index_json) return fail(MXD_ERR_INVALID, "fs_get_index: null argument");
    const std::string repo = repository ? repository : "";
    if (!repo.empty() && !valid_repository(repo)) return fail(MXD_ERR_INVALID, "repository invalid: " + repo);
    std::string text;
    int rc = read_file(join(basepath, repo.empty() ? std::string(kIndexFile) : join(repo, kIndexFile)), &text);
    if (rc != MXD_OK) return rc;
    *index_json = dup_out(text);
    return MXD_OK;
}

int mxc_push_stream(mxd_ctx* ctx, const char* basedir, const char* configfile, const mxc_uploader* up, int flags, char** report_json) {
    if (!ctx || !basedir || !configfile || !up || !report_json) return fail(MXD_ERR_INVALID, "push_stream: null argument");
    Manifest m; std::string blobs;
    int rc = push_stream(ctx, basedir, configfile, up, flags, &m, &blobs);
    if (rc != MXD_OK) return rc;
    *report_json = dup_out("{\"manifest\":" + json_manifest(m) + ",\"blobs\":" + blobs + "}");
    return MXD_OK;
}

int mxc_push_local(mxd_ctx* ctx, const char* basedir, const char* configfile, const char* basepath, const char* repository,
                   const char* version, int flags, char** report_json) {
    if (!ctx || !basedir || !configfile || !basepath || !repository || !version || !report_json)
        return fail(MXD_ERR_INVALID, "push_local: null argument");
    if (!valid_repository(repository)) return fail(MXD_ERR_INVALID, std::string("repository invalid: ") + repository);
    if (!valid_entry_name(version)) return fail(MXD_ERR_INVALID, std::string("version invalid: ") + version);
    FsUploader fs; fs.basepath = basepath; fs.repo = repository;
    fs.dir = join(join(join(basepath, repository), "blobs"), "sha256");
    int rc = mkdir_all(fs.dir, 0755);
    if (rc != MXD_OK) return rc;
    mxc_uploader up{};
    up.user = &fs; up.max_concurrent = 0;
    up.begin = FsUploader::begin; up.part_write = FsUploader::part_write; up.complete = FsUploader::complete; up.abort = FsUploader::abort;
    Manifest m; std::string blobs;
    rc = push_stream(ctx, basedir, configfile, &up, flags, &m, &blobs);
    if (rc != MXD_OK) return rc;
    const std::string mj = json_manifest(m);
    rc = fs_put_manifest(basepath, repository, version, kMediaTypeModelManifestJson, mj);          // push.go:57-64
    if (rc != MXD_OK) return rc;
    *report_json = dup_out("{\"manifest\":" + mj + ",\"blobs\":" + blobs + "}");
    return MXD_OK;
}

int mxc_push_local_tree(mxd_ctx* ctx, const char* basedir, const char* configfile, const char* basepath,
                        const char* repository, const char* version, char** report_json) {
    if (!ctx || !basedir || !configfile || !basepath || !repository || !version || !report_json)
        return fail(MXD_ERR_INVALID, "push_local_tree: null argument");
    if (!valid_repository(repository)) return fail(MXD_ERR_INVALID, std::string("repository invalid: ") + repository);
    if (!valid_entry_name(version)) return fail(MXD_ERR_INVALID, std::string("version invalid: ") + version);
    Manifest m;
    std::vector<PushItem> items;
    int rc = push_prepare(ctx, basedir, configfile, &m, &items);
    if (rc != MXD_OK) return rc;
    const std::string incoming_dir = join(join(join(basepath, repository), "blobs"), "sha256");
    rc = mkdir_all(incoming_dir, 0755);
    if (rc != MXD_OK) return rc;
    std::string blobs = "[";
    for (size_t i = 0; i < items.size(); ++i) {
        Descriptor& d = *items[i].desc;
        const std::string& src = items[i].path;
        d.digest.clear();                      // directory blobs: the archive is keyed by its tree root like everything else
        rc = push_file_fill(src, &d);
        if (rc != MXD_OK) return rc;
        const std::string tmp = join(incoming_dir, ".incoming-" + std::to_string((long)getpid()) + "-t" + std::to_string(i));
        TempWriter tw; tw.fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
        if (tw.fd < 0) return fail_errno("open " + tmp);
        uint8_t root[32]; uint64_t nch = 0, sz = 0;
        rc = mxd_tree_digest_file_tee(ctx, src.c_str(), nullptr, nullptr, 0, &nch, &sz, root, TempWriter::sink, &tw);   // one read: GPU + store
        close(tw.fd);
        if (rc != MXD_OK) { unlink(tmp.c_str()); return fail(rc, std::string("tree digest: ") + mxd_last_error()); }
        d.digest = digest_str(root);
        d.annotations[kTreeAnnotation] = tree_annotation(nch);
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
    rc = fs_put_manifest(basepath, repository, version, kMediaTypeModelManifestJson, mj);
    if (rc != MXD_OK) return rc;
    *report_json = dup_out("{\"manifest\":" + mj + ",\"blobs\":" + blobs + "}");
    return MXD_OK;
}

// Client.Pull against the same store (pull.go:19-39, pullFile :111-143, pullDirectory :146-208).  New:
// what is copied out of the store is hashed WHILE it is copied (the bytes are teed to the GPU and to a temp file next
// to the destination) and only a file whose digest equals the descriptor's is renamed into place; a corrupted store
// blob is reported as DIGEST_INVALID and leaves nothing behind.  The reference writes whatever it downloads
// (pull.go:137-142).
int mxc_pull_local(mxd_ctx* ctx, const char* basepath, const char* repository, const char* version, const char* into,
                   char** report_json) {
    if (!ctx || !basepath || !repository || !version || !into || !report_json) return fail(MXD_ERR_INVALID, "pull_local: null argument");
    if (!valid_repository(repository) || !valid_entry_name(version)) return fail(MXD_ERR_INVALID, "repository or version invalid");
    struct stat st;                                                                  // pull.go:20-32
    if (stat(into, &st) != 0) {
        if (errno != ENOENT) return fail_errno(std::string("stat ") + into);
        int rc = mkdir_all(into, 0755); if (rc != MXD_OK) return rc;
    } else if (!S_ISDIR(st.st_mode)) return fail(MXD_ERR_IO, std::string(into) + " is not a directory");
    std::string text;
    int rc = read_file(join(basepath, manifest_path(repository, version)), &text);
    if (rc != MXD_OK) return rc;
    Manifest m; std::string err;
    if (!manifest_from_json(text.c_str(), &m, &err) || !validate_manifest(m, &err)) return fail(MXC_ERR_MANIFEST, "manifest invalid: " + err);
    std::vector<Descriptor> descs = m.blobs; descs.push_back(m.config);
    std::vector<PullState> states;
    rc = pull_check(ctx, into, descs, &states);
    if (rc != MXD_OK) return rc;

    struct Fetch { size_t i; std::string src, dst, tmp; TempWriter tw; uint8_t got[32]; bool tree; bool is_dir; };
    std::vector<std::unique_ptr<Fetch>> fetches;
    std::vector<std::string> status(descs.size());
    for (size_t i = 0; i < descs.size(); ++i) {
        status[i] = states[i].state;
        if (status[i] == "already exists") continue;
        const bool is_dir = descs[i].mediaType == kMediaTypeModelDirectoryTarGz;
        const std::string dst = is_dir ? join(join(into, ".modelx"), descs[i].name + ".tar.gz") : join(into, descs[i].name);   // pull.go:160
        rc = mkdir_all(dir_of(dst), 0777); if (rc != MXD_OK) return rc;
        mode_t perm = (mode_t)(descs[i].mode & 0777); if (perm == 0 || is_dir) perm = 0644;   // OpenWriteFile, pull.go:65-73
        if (status[i] == "empty") { rc = write_file(dst, "", perm); if (rc != MXD_OK) return rc; continue; }
        std::string rel; rc = blob_digest_path(repository, descs[i].digest, &rel); if (rc != MXD_OK) return rc;
        auto f = std::unique_ptr<Fetch>(new Fetch());
        f->i = i; f->src = join(basepath, rel); f->dst = dst; f->tmp = dst + ".modelx-partial"; f->tree = is_tree_keyed(descs[i]); f->is_dir = is_dir;
        if (access(f->src.c_str(), R_OK) != 0) return fail(MXC_ERR_NOT_FOUND, "blob not found: " + descs[i].digest);
        f->tw.fd = open(f->tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, perm);
        if (f->tw.fd < 0) return fail_errno("open " + f->tmp);
        fetches.push_back(std::move(f));
    }
    auto cleanup = [&] { for (auto& f : fetches) { if (f->tw.fd >= 0) { close(f->tw.fd); f->tw.fd = -1; } unlink(f->tmp.c_str()); } };
    // whole-file keyed blobs: ONE coalesced pass, each copy teed through its own chain; tree keyed: one by one
    std::vector<mxd_file_job> jobs;
    std::vector<Fetch*> jf;
    for (auto& f : fetches) {
        if (f->tree) {
            uint64_t nch = 0, sz = 0;
            rc = mxd_tree_digest_file_tee(ctx, f->src.c_str(), nullptr, nullptr, 0, &nch, &sz, f->got, TempWriter::sink, &f->tw);
            if (rc != MXD_OK) { cleanup(); return fail(rc, std::string("tree digest: ") + mxd_last_error()); }
        } else {
            mxd_file_job jb{}; jb.path = f->src.c_str(); jb.out = f->got; jb.sink = TempWriter::sink; jb.sink_user = &f->tw;
            jobs.push_back(jb); jf.push_back(f.get());
        }
    }
    if (!jobs.empty()) {
        rc = mxd_sha256_file_jobs(ctx, jobs.data(), jobs.size());
        if (rc != MXD_OK) { cleanup(); return fail(rc, std::string("digest: ") + mxd_last_error()); }
    }
    std::string bad;
    for (auto& f : fetches) {
        close(f->tw.fd); f->tw.fd = -1;
        if (digest_str(f->got) != descs[f->i].digest) {
            unlink(f->tmp.c_str());
            if (bad.empty()) bad = "blob '" + descs[f->i].name + "' in the store hashes to " + digest_str(f->got) + ", manifest says " + descs[f->i].digest;
            status[f->i] = "digest invalid";
            continue;
        }
        if (rename(f->tmp.c_str(), f->dst.c_str()) != 0) { int e = errno; cleanup(); errno = e; return fail_errno("rename " + f->tmp); }
        if (f->is_dir) {
            rc = untgz_file(f->dst, join(into, descs[f->i].name));                   // pull.go:178-186
            if (rc != MXD_OK) { cleanup(); return rc; }
            chmod(join(into, descs[f->i].name).c_str(), (mode_t)(descs[f->i].mode & 0777));
        } else {
            mode_t perm = (mode_t)(descs[f->i].mode & 0777); if (perm == 0) perm = 0644;
            chmod(f->dst.c_str(), perm);
        }
        status[f->i] = "done";
    }
    std::string o = "[";
    for (size_t i = 0; i < descs.size(); ++i) {
        if (i) o += ',';
        o += "{\"name\":"; json_string(o, descs[i].name); o += ",\"status\":"; json_string(o, status[i]); o += '}';
    }
    o += ']';
    *report_json = dup_out(o);
    if (!bad.empty()) return fail(MXC_ERR_DIGEST_INVALID, "digest invalid: " + bad);
    return MXD_OK;
}

}  // extern "C"

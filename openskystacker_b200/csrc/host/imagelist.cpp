#include "imagelist.hpp"

#include <algorithm>
#include <cctype>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>

#include "image_io.hpp"

namespace openskystacker {
namespace {

// A JSON value: just enough of RFC 8259 for the image list.
struct Json {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;
    const Json *get(const std::string &k) const
    {
        for (auto &kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct Parser {
    const std::string &s;
    size_t p = 0;
    bool ok = true;
    explicit Parser(const std::string &t) : s(t) {}
    void ws() { while (p < s.size() && isspace((unsigned char)s[p])) p++; }
    bool eat(char c) { ws(); if (p < s.size() && s[p] == c) { p++; return true; } return false; }
    std::string string()
    {
        std::string out;
        if (!eat('"')) { ok = false; return out; }
        while (p < s.size() && s[p] != '"') {
            char c = s[p++];
            if (c == '\\' && p < s.size()) {
                char e = s[p++];
                switch (e) {
                case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                case 'u': {
                    unsigned cp = 0;
                    for (int i = 0; i < 4 && p < s.size(); i++) cp = cp * 16 + (unsigned)(isdigit((unsigned char)s[p]) ? s[p] - '0' : (tolower(s[p]) - 'a' + 10)), p++;
                    if (cp < 0x80) out += (char)cp;
                    else if (cp < 0x800) { out += (char)(0xc0 | cp >> 6); out += (char)(0x80 | (cp & 63)); }
                    else { out += (char)(0xe0 | cp >> 12); out += (char)(0x80 | ((cp >> 6) & 63)); out += (char)(0x80 | (cp & 63)); }
                    break;
                }
                default: out += e;
                }
            } else out += c;
        }
        if (p >= s.size()) ok = false; else p++;
        return out;
    }
    Json value()
    {
        Json v;
        ws();
        if (p >= s.size()) { ok = false; return v; }
        char c = s[p];
        if (c == '[') {
            p++; v.kind = Json::Array;
            if (eat(']')) return v;
            do { v.arr.push_back(value()); } while (ok && eat(','));
            if (!eat(']')) ok = false;
        } else if (c == '{') {
            p++; v.kind = Json::Object;
            if (eat('}')) return v;
            do { ws(); std::string k = string(); if (!eat(':')) ok = false; v.obj.emplace_back(k, value()); } while (ok && eat(','));
            if (!eat('}')) ok = false;
        } else if (c == '"') { v.kind = Json::String; v.str = string(); }
        else if (!s.compare(p, 4, "true")) { v.kind = Json::Bool; v.b = true; p += 4; }
        else if (!s.compare(p, 5, "false")) { v.kind = Json::Bool; p += 5; }
        else if (!s.compare(p, 4, "null")) { p += 4; }
        else {
            size_t q = p;
            while (q < s.size() && (isdigit((unsigned char)s[q]) || strchr("+-.eE", s[q]))) q++;
            if (q == p) { ok = false; return v; }
            v.kind = Json::Number; v.num = atof(s.substr(p, q - p).c_str()); p = q;
        }
        return v;
    }
};

std::string lowerExt(const std::string &f)
{
    size_t slash = f.find_last_of("/\\"), dot = f.find_last_of('.');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return "";
    std::string e = f.substr(dot + 1);
    std::transform(e.begin(), e.end(), e.begin(), [](unsigned char c) { return (char)tolower(c); });
    return e;
}

}  // namespace

ImageType getImageType(const std::string &filename)
{
    static const char *fits[] = {"fit", "fits", "fts"};
    static const char *raw[] = {"3fr", "ari", "arw", "bay", "crw", "cr2", "cap", "data", "dcs", "dcr", "dng", "drf", "eip", "erf",
                                "fff", "gpr", "iiq", "k25", "kdc", "mdc", "mef", "mos", "mrw", "nef", "nrw", "obm", "orf", "pef",
                                "ptx", "pxn", "r3d", "raf", "raw", "rwl", "rw2", "rwz", "sr2", "srf", "srw", "x3f"};
    std::string e = lowerExt(filename);
    for (auto x : fits) if (e == x) return FITS_IMAGE;
    for (auto x : raw) if (e == x) return RAW_IMAGE;
    return RGB_IMAGE;
}

ImageRecord *getImageRecord(const std::string &filename)
{
    ImageRecord *r = new ImageRecord();
    r->filename = filename;
    r->checked = true;
    ImageInfo info;
    std::string err;
    // util.cpp:112-113 (FITS: axis(0) x axis(1)) and :134-143 (RGB, without EXIF)
    if (getImageType(filename) != RAW_IMAGE && probeImage(filename, &info, &err)) {
        r->width = info.width;
        r->height = info.height;
    }
    return r;
}

std::vector<ImageRecord *> loadImageList(const std::string &filename, int *err)
{
    std::vector<ImageRecord *> result;
    if (err) *err = 0;
    std::ifstream f(filename);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string text = ss.str();
    Parser p(text);
    Json doc = p.value();
    if (!f || !p.ok || doc.kind != Json::Array) { if (err) *err = -1; return result; }
    std::string dir = ".";
    size_t slash = filename.find_last_of('/');
    if (slash != std::string::npos) dir = filename.substr(0, slash);
    static const char *types[] = {"light", "dark", "darkflat", "flat", "bias"};
    for (const Json &img : doc.arr) {
        if (img.kind != Json::Object || img.obj.empty()) { if (err) *err = -2; return result; }
        const Json *fn = img.get("filename");
        if (!fn || fn->kind != Json::String) { if (err) *err = -3; return result; }
        std::string name = fn->str;
        if (name.empty() || name[0] != '/') name = dir + "/" + name;
        const Json *ty = img.get("type");
        std::string type = ty && ty->kind == Json::String ? ty->str : "";
        const Json *ck = img.get("checked");
        bool checked = ck && ck->kind == Json::Bool && ck->b; // QJsonValue::toBool(): false unless a true bool
        std::unique_ptr<ImageRecord> rec(getImageRecord(name));
        if (type == "ref") {
            rec->type = ImageRecord::LIGHT;
            rec->reference = true;
        } else {
            int idx = -1;
            for (int i = 0; i < 5; i++) if (type == types[i]) idx = i;
            if (idx < 0) { if (err) *err = -4; return result; }
            rec->type = (ImageRecord::FrameType)idx;
        }
        rec->checked = checked;
        result.push_back(rec.release());
    }
    return result;
}

std::string getTimeString(int seconds)
{
    auto unit = [](int n, const char *one, const char *many) { return std::to_string(n) + " " + (n == 1 ? one : many); };
    if (seconds < 60) return unit(seconds, "second", "seconds");
    if (seconds < 3600) return unit(seconds / 60, "minute", "minutes") + " and " + unit(seconds % 60, "second", "seconds");
    if (seconds < 86400) {
        int rem = seconds % 3600;
        return unit(seconds / 3600, "hour", "hours") + ", " + unit(rem / 60, "minute", "minutes") + ", and " + unit(rem % 60, "second", "seconds");
    }
    int drem = seconds % 86400, hrem = drem % 3600;
    return unit(seconds / 86400, "day", "days") + ", " + unit(drem / 3600, "hour", "hours") + ", " + unit(hrem / 60, "minute", "minutes") +
           ", and " + unit(hrem % 60, "second", "seconds");
}

}  // namespace openskystacker

// Scenario files of the reference application: config.json
// {"plyFilename", "fixedVertices":[..], "forces":[[node,dof,magnitude]..]}
// [REF src/utilities.hpp:50-95] and the ASCII PLY edge mesh with per-edge list
// properties edge_normal / beam_dimensions / beam_material
// [REF src/edgemesh.hpp:26-28; src/edgemesh.cpp:115-135; Models/TwoLineSegments.ply].
// The reference reads them through nlohmann::json and vcglib's nanoply; this is
// a dependency-free reader for exactly that schema.  Pure C++17, no CUDA.
#include <cctype>
#include <cerrno>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "mbfea_internal.hpp"

namespace mbfea {

namespace {

// ---- a small JSON value + recursive-descent parser ---------------------------
struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<JVal> arr;
    std::map<std::string, JVal> obj;
};

struct JParser {
    const std::string &s;
    size_t i = 0;
    std::string err;
    explicit JParser(const std::string &src) : s(src) {}
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    bool fail(const char *m) { if (err.empty()) err = std::string(m) + " at byte " + std::to_string(i); return false; }
    bool parse_string(std::string &out)
    {
        if (s[i] != '"') return fail("expected string");
        ++i;
        out.clear();
        while (i < s.size() && s[i] != '"') {
            char ch = s[i++];
            if (ch == '\\') {
                if (i >= s.size()) return fail("bad escape");
                char e = s[i++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (i + 4 > s.size()) return fail("bad \\u escape");
                        unsigned cp = (unsigned)std::strtoul(s.substr(i, 4).c_str(), nullptr, 16);
                        i += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += e;  // \" \\ \/
                }
            } else out += ch;
        }
        if (i >= s.size()) return fail("unterminated string");
        ++i;
        return true;
    }
    int depth = 0;
    struct Nest {   // bounded nesting: a hostile file must not overflow the stack
        int &d;
        explicit Nest(int &dd) : d(dd) { ++d; }
        ~Nest() { --d; }
    };
    bool parse(JVal &v)
    {
        Nest nest(depth);
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (i >= s.size()) return fail("unexpected end");
        const char ch = s[i];
        if (ch == '{') {
            v.kind = JVal::Obj; ++i; ws();
            if (i < s.size() && s[i] == '}') { ++i; return true; }
            for (;;) {
                ws();
                std::string key;
                if (i >= s.size() || !parse_string(key)) return fail("expected key");
                ws();
                if (i >= s.size() || s[i] != ':') return fail("expected ':'");
                ++i;
                JVal child;
                if (!parse(child)) return false;
                v.obj[key] = std::move(child);
                ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == '}') { ++i; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (ch == '[') {
            v.kind = JVal::Arr; ++i; ws();
            if (i < s.size() && s[i] == ']') { ++i; return true; }
            for (;;) {
                JVal child;
                if (!parse(child)) return false;
                v.arr.push_back(std::move(child));
                ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == ']') { ++i; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (ch == '"') { v.kind = JVal::Str; return parse_string(v.str); }
        if (s.compare(i, 4, "true") == 0) { v.kind = JVal::Bool; v.b = true; i += 4; return true; }
        if (s.compare(i, 5, "false") == 0) { v.kind = JVal::Bool; v.b = false; i += 5; return true; }
        if (s.compare(i, 4, "null") == 0) { v.kind = JVal::Null; i += 4; return true; }
        const char *b = s.c_str() + i;
        char *e = nullptr;
        errno = 0;
        const double d = std::strtod(b, &e);
        if (e == b) return fail("unexpected character");
        v.kind = JVal::Num; v.num = d; i += (size_t)(e - b);
        return true;
    }
};

bool read_file(const char *path, std::string &out)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    std::ostringstream ss;
    ss << in.rdbuf();
    out = ss.str();
    return true;
}

struct PlyProp {
    std::string name, type, count_type;
    bool is_list = false;
};
struct PlyElem {
    std::string name;
    int64_t count = 0;
    std::vector<PlyProp> props;
};

constexpr long kMaxListLen = 1 << 16;   // the reference's lists hold 2 or 3 values (count type uchar)

bool is_float_type(const std::string &t) { return t == "float" || t == "float32"; }

// ASCII token -> value of the declared PLY type, widened to double
double typed_value(const std::string &tok, const std::string &type, bool &ok)
{
    char *e = nullptr;
    const double d = std::strtod(tok.c_str(), &e);
    ok = e != tok.c_str() && *e == '\0';
    if (is_float_type(type)) return (double)(float)d;
    return d;
}

}  // namespace

int load_ply_edge_mesh(const char *ply_path, Scenario &sc, std::string &msg)
{
    std::ifstream in(ply_path);
    if (!in) { msg = std::string("Error: Unable to open ") + ply_path; return MBFEA_EIO; }
    // every element row takes at least two bytes of text: a declared count beyond the file size is a
    // corrupt header, not a reason to allocate
    in.seekg(0, std::ios::end);
    const int64_t file_bytes = (int64_t)in.tellg();
    in.seekg(0, std::ios::beg);
    std::string line;
    if (!std::getline(in, line) || line.substr(0, 3) != "ply") { msg = "not a PLY file"; return MBFEA_EIO; }
    std::vector<PlyElem> elems;
    bool ascii = false, header_done = false;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string kw;
        ls >> kw;
        if (kw == "format") {
            std::string f;
            ls >> f;
            ascii = (f == "ascii");
        } else if (kw == "element") {
            PlyElem el;
            if (!(ls >> el.name >> el.count) || el.count < 0 || el.count > file_bytes) {
                msg = "PLY: bad element count in '" + line + "'";
                return MBFEA_EIO;
            }
            elems.push_back(el);
        } else if (kw == "property") {
            if (elems.empty()) { msg = "PLY property before any element"; return MBFEA_EIO; }
            PlyProp p;
            std::string t;
            ls >> t;
            if (t == "list") { p.is_list = true; ls >> p.count_type >> p.type >> p.name; }
            else { p.type = t; ls >> p.name; }
            // anything after the name (the fixture carries a free-text note) is ignored
            elems.back().props.push_back(p);
        } else if (kw == "end_header") { header_done = true; break; }
        // comment / obj_info lines are skipped
    }
    if (!header_done) { msg = "PLY header has no end_header"; return MBFEA_EIO; }
    if (!ascii) { msg = "only ASCII PLY edge meshes are supported"; return MBFEA_EIO; }

    auto next_token = [&](std::string &tok) -> bool { return (bool)(in >> tok); };  // skips blank lines
    bool have_v = false, have_e = false;
    for (const PlyElem &el : elems) {
        if (el.name == "vertex") {
            have_v = true;
            sc.V = el.count;
            sc.nodes_cm.assign((size_t)3 * el.count, 0.0);
            for (int64_t v = 0; v < el.count; ++v)
                for (const PlyProp &p : el.props) {
                    std::string tok;
                    bool ok = false;
                    if (p.is_list) {
                        if (!next_token(tok)) { msg = "PLY truncated (vertex list)"; return MBFEA_EIO; }
                        const long n = std::strtol(tok.c_str(), nullptr, 10);
                        if (n < 0 || n > kMaxListLen) { msg = "PLY: bad list length '" + tok + "'"; return MBFEA_EIO; }
                        for (long k = 0; k < n; ++k) if (!next_token(tok)) { msg = "PLY truncated"; return MBFEA_EIO; }
                        continue;
                    }
                    if (!next_token(tok)) { msg = "PLY truncated (vertices)"; return MBFEA_EIO; }
                    const double d = typed_value(tok, p.type, ok);
                    if (!ok) { msg = "PLY: bad number '" + tok + "'"; return MBFEA_EIO; }
                    if (p.name == "x") sc.nodes_cm[(size_t)v] = d;
                    else if (p.name == "y") sc.nodes_cm[(size_t)(el.count + v)] = d;
                    else if (p.name == "z") sc.nodes_cm[(size_t)(2 * el.count + v)] = d;
                }
        } else if (el.name == "edge") {
            have_e = true;
            sc.E = el.count;
            const int64_t E = el.count;
            sc.elems_cm.assign((size_t)2 * E, 0);
            // defaults of VCGEdgeMesh::setDefaultAttributes [REF src/edgemesh.cpp:226-232]
            sc.normals_cm.assign((size_t)3 * E, 0.0);
            for (int64_t e = 0; e < E; ++e) sc.normals_cm[(size_t)(2 * E + e)] = 1.0;
            sc.dims_bh.assign((size_t)2 * E, 0.01f);
            sc.mat_nuE.assign((size_t)2 * E, 0.0f);
            for (int64_t e = 0; e < E; ++e) { sc.mat_nuE[(size_t)2 * e] = 0.3f; sc.mat_nuE[(size_t)2 * e + 1] = 210.0f; }
            bool has_n = false;
            for (int64_t e = 0; e < E; ++e)
                for (const PlyProp &p : el.props) {
                    std::string tok;
                    bool ok = false;
                    if (!p.is_list) {
                        if (!next_token(tok)) { msg = "PLY truncated (edges)"; return MBFEA_EIO; }
                        char *end = nullptr;
                        const long long idx = std::strtoll(tok.c_str(), &end, 10);
                        if (end == tok.c_str() || idx < INT32_MIN || idx > INT32_MAX) {   // no silent wrap into range
                            msg = "PLY: bad vertex index '" + tok + "'";
                            return MBFEA_EIO;
                        }
                        if (p.name == "vertex1") sc.elems_cm[(size_t)e] = (int32_t)idx;
                        else if (p.name == "vertex2") sc.elems_cm[(size_t)(E + e)] = (int32_t)idx;
                        continue;
                    }
                    if (!next_token(tok)) { msg = "PLY truncated (edge list)"; return MBFEA_EIO; }
                    const long n = std::strtol(tok.c_str(), nullptr, 10);
                    if (n < 0 || n > kMaxListLen) { msg = "PLY: bad list length '" + tok + "'"; return MBFEA_EIO; }
                    std::vector<double> vals((size_t)n);
                    for (long k = 0; k < n; ++k) {
                        if (!next_token(tok)) { msg = "PLY truncated (edge list values)"; return MBFEA_EIO; }
                        vals[(size_t)k] = typed_value(tok, p.type, ok);
                        if (!ok) { msg = "PLY: bad number '" + tok + "'"; return MBFEA_EIO; }
                    }
                    if (p.name == "edge_normal") {
                        if (n != 3) { msg = "PLY: edge_normal must have 3 entries"; return MBFEA_EIO; }
                        has_n = true;
                        for (int k = 0; k < 3; ++k) sc.normals_cm[(size_t)(k * E + e)] = vals[(size_t)k];
                    } else if (p.name == "beam_dimensions") {
                        if (n != 2) { msg = "PLY: beam_dimensions must have 2 entries {b,h}"; return MBFEA_EIO; }
                        sc.dims_bh[(size_t)2 * e] = (float)vals[0]; sc.dims_bh[(size_t)2 * e + 1] = (float)vals[1];
                    } else if (p.name == "beam_material") {
                        if (n != 2) { msg = "PLY: beam_material must have 2 entries {nu,E_GPa}"; return MBFEA_EIO; }
                        sc.mat_nuE[(size_t)2 * e] = (float)vals[0]; sc.mat_nuE[(size_t)2 * e + 1] = (float)vals[1];
                    }
                }
            (void)has_n;
        } else {
            // skip unknown elements
            for (int64_t r = 0; r < el.count; ++r)
                for (const PlyProp &p : el.props) {
                    std::string tok;
                    if (!next_token(tok)) { msg = "PLY truncated"; return MBFEA_EIO; }
                    if (p.is_list) {
                        const long n = std::strtol(tok.c_str(), nullptr, 10);
                        if (n < 0 || n > kMaxListLen) { msg = "PLY: bad list length '" + tok + "'"; return MBFEA_EIO; }
                        for (long k = 0; k < n; ++k) if (!next_token(tok)) { msg = "PLY truncated"; return MBFEA_EIO; }
                    }
                }
        }
    }
    if (!have_v || !have_e) { msg = "PLY needs 'vertex' and 'edge' elements"; return MBFEA_EIO; }
    return MBFEA_OK;
}

int load_scenario_files(const char *json_path, const char *ply_override, Scenario &sc, std::string &msg)
{
    std::string text;
    if (!read_file(json_path, text)) { msg = std::string("cannot open ") + json_path; return MBFEA_EIO; }
    JParser jp(text);
    JVal root;
    if (!jp.parse(root) || root.kind != JVal::Obj) { msg = "config.json: " + (jp.err.empty() ? std::string("not an object") : jp.err); return MBFEA_EIO; }
    std::string ply;
    if (ply_override && *ply_override) ply = ply_override;
    else {
        auto it = root.obj.find("plyFilename");  // [REF src/utilities.hpp:57-59]
        if (it == root.obj.end() || it->second.kind != JVal::Str) { msg = "config.json has no plyFilename"; return MBFEA_EIO; }
        ply = it->second.str;
        if (!ply.empty() && ply[0] != '/') {  // relative to the scenario file
            std::string dir(json_path);
            const size_t slash = dir.find_last_of('/');
            if (slash != std::string::npos) ply = dir.substr(0, slash + 1) + ply;
        }
    }
    if (int rc = load_ply_edge_mesh(ply.c_str(), sc, msg)) return rc;
    sc.fixed.clear(); sc.f_node.clear(); sc.f_dof.clear(); sc.f_mag.clear();
    auto fv = root.obj.find("fixedVertices");  // [REF src/utilities.hpp:91-94]
    if (fv != root.obj.end()) {
        if (fv->second.kind != JVal::Arr) { msg = "config.json: fixedVertices must be an array"; return MBFEA_EIO; }
        for (const JVal &v : fv->second.arr) {
            if (v.kind != JVal::Num || !(std::fabs(v.num) <= (double)INT32_MAX)) {   // also rejects NaN
                msg = "config.json: fixedVertices entries must be integers";
                return MBFEA_EIO;
            }
            sc.fixed.push_back((int32_t)v.num);
        }
    }
    auto fo = root.obj.find("forces");  // [REF src/utilities.hpp:70-79]
    if (fo != root.obj.end()) {
        if (fo->second.kind != JVal::Arr) { msg = "config.json: forces must be an array"; return MBFEA_EIO; }
        for (const JVal &v : fo->second.arr) {
            if (v.kind != JVal::Arr || v.arr.size() < 3 || v.arr[0].kind != JVal::Num || v.arr[1].kind != JVal::Num ||
                v.arr[2].kind != JVal::Num || !(std::fabs(v.arr[0].num) < 9.0e18) || !(std::fabs(v.arr[1].num) < 9.0e18)) {
                msg = "config.json: a force is [node, dof, magnitude]";
                return MBFEA_EIO;
            }
            const int64_t node = (int64_t)v.arr[0].num, dof = (int64_t)v.arr[1].num;
            const double mag = v.arr[2].num;
            // Expects(nf.dof >= 0 && nf.dof < 6 && nf.index >= 0 && nf.magnitude >= 0) [REF :77]
            if (!(dof >= 0 && dof < 6 && node >= 0 && mag >= 0)) {
                msg = "precondition violated: scenario force needs dof in [0,6), index >= 0, magnitude >= 0";
                return MBFEA_EPRECOND;
            }
            sc.f_node.push_back(node); sc.f_dof.push_back(dof); sc.f_mag.push_back(mag);
        }
    }
    return MBFEA_OK;
}

}  // namespace mbfea

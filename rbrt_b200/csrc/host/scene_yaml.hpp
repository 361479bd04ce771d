// scene_yaml.hpp — the scene description of rbrt_lib/src/blueprints.rs:15-48 read from YAML the way the reference reads it
// (`serde_yaml::from_reader` into `SceneBlueprint`, blueprints.rs:76-92), for the C++ host (rbrt_cli.cpp).
//
// serde_yaml `0.9` and serde are not under /root/reference and only the two scene files pin them, so this
// restates the published behaviour a scene author can observe:
//   YAML syntax  block mappings and sequences (also a sequence at its key's indentation), flow `[a, b]` / `{k: v}` collections
//                (nested, over several lines), plain / 'single' / "double" quoted scalars, `# comments`, `---` / `...` markers,
//                `&anchor` / `*alias`.  Refused with a message: a second document (serde_yaml refuses it too), tags, block
//                scalars (`|`, `>`), multi-line plain or quoted scalars, `? ` complex keys, tabs as indentation.
//   structs      every field of SceneBlueprint / CameraBluePrint / TriangleMeshBlueprint / SphereBlueprint / Vec3 is required
//                except the two `Option`s (`albedo`, `material_param`: absent or null = None); `mesh_blueprints` and
//                `sphere_blueprints` must be there (write `[]` for none); unknown keys are ignored, a key given twice is an
//                error; a struct may also be written as the sequence of its fields (`center: [0, 1, 2]`), as serde allows.
//   f32          a PLAIN scalar that is an integer (decimal without leading zeros, `0x` / `0o` / `0b`) or a float (`.inf`,
//                `.nan`, or what Rust's f64::from_str takes and is finite); floats are rounded to f64 first and then to f32,
//                integers straight to f32, as serde's `v as f32` does.  null / booleans / quoted scalars are type errors.
//   String       any scalar, its text taken as written.
//   tabs         as libyaml (serde_yaml's parser): never as indentation, not where a block value would start (`key:<tab>v`), fine after a value.
//                (PyYAML, the Python host's parser, refuses a tab after a value too: the one known difference between the hosts.)
// Every failure is the reference's panic: message on stderr, exit code 101.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/rbrt_gpu.h"

namespace scene_yaml {

struct ParseError { std::string msg; };

struct Node {
    enum Kind { Null, Scalar, Map, Seq } kind = Null;
    std::string scalar;                                       // Scalar: the text (escapes resolved); Null: "" / "~" / "null"
    bool quoted = false;
    int line = 0;
    std::vector<std::pair<std::shared_ptr<Node>, std::shared_ptr<Node>>> map;
    std::vector<std::shared_ptr<Node>> seq;
};
using NodeP = std::shared_ptr<Node>;

class Parser {
  public:
    explicit Parser(const std::string& text) {
        size_t a = 0;
        while (a <= text.size()) {
            size_t b = text.find('\n', a);
            if (b == std::string::npos) b = text.size();
            std::string l = text.substr(a, b - a);
            if (!l.empty() && l.back() == '\r') l.pop_back();
            lines_.push_back(l);
            a = b + 1;
        }
        if (!lines_.empty() && lines_[0].compare(0, 3, "\xEF\xBB\xBF") == 0) lines_[0].erase(0, 3);
    }

    NodeP document() {
        skip_blank();
        if (li_ < lines_.size() && is_marker(lines_[li_], "---")) {
            std::string rest = strip(strip_comment(lines_[li_].substr(3)));
            if (rest.empty()) ++li_;
            else lines_[li_] = "   " + lines_[li_].substr(3);                       // `--- value`
            skip_blank();
        }
        NodeP root = std::make_shared<Node>();
        if (li_ < lines_.size() && !is_marker(lines_[li_], "---") && !is_marker(lines_[li_], "...")) root = block_node(indent_of(lines_[li_]));
        skip_blank();
        while (li_ < lines_.size() && is_marker(lines_[li_], "...")) { ++li_; skip_blank(); }
        if (li_ < lines_.size()) {
            if (is_marker(lines_[li_], "---")) fail("deserializing from YAML containing more than one document is not supported");
            fail("unexpected content after the document");
        }
        return root;
    }

  private:
    std::vector<std::string> lines_;
    size_t li_ = 0;
    std::map<std::string, NodeP> anchors_;

    [[noreturn]] void fail(const std::string& what, int line = -1) const {
        std::ostringstream o;
        o << what << " at line " << (line >= 0 ? line : (int)li_ + 1);
        throw ParseError{o.str()};
    }
    static bool is_blank_ch(char c) { return c == ' ' || c == '\t'; }
    static std::string strip(const std::string& s) {
        size_t a = 0, b = s.size();
        while (a < b && is_blank_ch(s[a])) ++a;
        while (b > a && is_blank_ch(s[b - 1])) --b;
        return s.substr(a, b - a);
    }
    static bool is_marker(const std::string& l, const char* m) { return l.compare(0, 3, m) == 0 && (l.size() == 3 || is_blank_ch(l[3])); }
    // a `#` starts a comment at the start of the text or after white space, outside quotes
    static std::string strip_comment(const std::string& s) {
        char q = 0;
        for (size_t i = 0; i < s.size(); ++i) {
            char c = s[i];
            if (q) {
                if (q == '"' && c == '\\') ++i;
                else if (c == q) { if (q == '\'' && i + 1 < s.size() && s[i + 1] == '\'') ++i; else q = 0; }
            } else if ((c == '"' || c == '\'') && (i == 0 || is_blank_ch(s[i - 1]) || strchr("[]{},:", s[i - 1]))) q = c;
            else if (c == '#' && (i == 0 || is_blank_ch(s[i - 1]))) return s.substr(0, i);
        }
        return s;
    }
    int indent_of(const std::string& l) const {
        int n = 0;
        while (n < (int)l.size() && l[n] == ' ') ++n;
        if (n < (int)l.size() && l[n] == '\t') fail("a tab cannot indent a YAML block");
        return n;
    }
    void skip_blank() {
        while (li_ < lines_.size() && strip(strip_comment(lines_[li_])).empty()) ++li_;
    }
    std::string content(size_t li) const { return strip(strip_comment(lines_[li])); }

    static bool starts_seq_item(const std::string& t) { return t == "-" || (t.size() > 1 && t[0] == '-' && is_blank_ch(t[1])); }

    // position of the `:` that ends an implicit key of a block mapping entry (followed by blank or end), or npos
    static size_t key_colon(const std::string& t) {
        size_t i = 0;
        if (!t.empty() && (t[0] == '"' || t[0] == '\'')) {              // quoted key
            char q = t[0];
            for (i = 1; i < t.size(); ++i) {
                if (q == '"' && t[i] == '\\') { ++i; continue; }
                if (t[i] == q) { if (q == '\'' && i + 1 < t.size() && t[i + 1] == '\'') { ++i; continue; } break; }
            }
            if (i >= t.size()) return std::string::npos;
            ++i;
            while (i < t.size() && is_blank_ch(t[i])) ++i;
            return (i < t.size() && t[i] == ':' && (i + 1 == t.size() || is_blank_ch(t[i + 1]))) ? i : std::string::npos;
        }
        if (!t.empty() && strchr("[{&*!|>%@`", t[0])) return std::string::npos;     // a flow collection / anchored value / ..., not a key
        for (; i < t.size(); ++i)
            if (t[i] == ':' && (i + 1 == t.size() || is_blank_ch(t[i + 1]))) return i;
        return std::string::npos;
    }

    // libyaml (PyYAML, serde_yaml's unsafe-libyaml) does not skip a tab where a block-context token may start: `key:\tvalue` and `-\tvalue` are
    // "found character that cannot start any token"
    void no_tab_before_value(const std::string& t, size_t from) const {
        size_t i = from;
        while (i < t.size() && is_blank_ch(t[i])) { if (t[i] == '\t') fail("found character '\\t' that cannot start any token"); ++i; }
    }

    NodeP block_node(int indent) {
        const std::string t = content(li_);
        if (starts_seq_item(t)) return block_seq(indent);
        if (key_colon(t) != std::string::npos) return block_map(indent);
        // a scalar or a flow collection standing alone
        std::string rest = t;
        ++li_;
        NodeP n = inline_value(rest, indent);
        return n;
    }

    NodeP block_seq(int indent) {
        NodeP n = std::make_shared<Node>();
        n->kind = Node::Seq; n->line = (int)li_ + 1;
        while (true) {
            skip_blank();
            if (li_ >= lines_.size() || is_marker(lines_[li_], "---") || is_marker(lines_[li_], "...")) break;
            int ind = indent_of(lines_[li_]);
            if (ind < indent) break;
            std::string t = content(li_);
            if (ind > indent) fail("bad indentation of a sequence entry");
            if (!starts_seq_item(t)) break;                                  // (a sequence written at its key's indentation ends at the next key)
            no_tab_before_value(strip_comment(lines_[li_]), (size_t)ind + 1);
            // the entry's content starts after "- ": treat it as a block that begins on this line, indented to that column
            std::string& raw = lines_[li_];
            raw[ind] = ' ';
            if (strip(strip_comment(raw)).empty()) {
                ++li_;
                skip_blank();
                if (li_ < lines_.size() && !is_marker(lines_[li_], "---") && !is_marker(lines_[li_], "...") && indent_of(lines_[li_]) > indent) n->seq.push_back(block_node(indent_of(lines_[li_])));
                else n->seq.push_back(std::make_shared<Node>());
            } else n->seq.push_back(block_node(indent_of(raw)));
        }
        return n;
    }

    NodeP block_map(int indent) {
        NodeP n = std::make_shared<Node>();
        n->kind = Node::Map; n->line = (int)li_ + 1;
        while (true) {
            skip_blank();
            if (li_ >= lines_.size() || is_marker(lines_[li_], "---") || is_marker(lines_[li_], "...")) break;
            int ind = indent_of(lines_[li_]);
            if (ind < indent) break;
            std::string t = content(li_);
            if (ind > indent) fail("bad indentation of a mapping entry");
            if (t.size() > 1 && t[0] == '?' && is_blank_ch(t[1])) fail("complex `? ` keys are not supported");
            size_t c = key_colon(t);
            if (c == std::string::npos) fail("expected `key: value`");
            size_t used = 0;
            NodeP key = flow_scalar(strip(t.substr(0, c)), used, false);
            key->line = (int)li_ + 1;
            no_tab_before_value(strip_comment(lines_[li_]), (size_t)ind + c + 1);      // (the raw line: `key:<tab>` at the end of a line is refused too)
            std::string rest = strip(t.substr(c + 1));
            ++li_;
            NodeP val;
            std::string anchor = take_anchor(rest);
            if (rest.empty()) {
                skip_blank();
                const bool more = li_ < lines_.size() && !is_marker(lines_[li_], "---") && !is_marker(lines_[li_], "...");
                if (more && indent_of(lines_[li_]) > indent) val = block_node(indent_of(lines_[li_]));
                else if (more && indent_of(lines_[li_]) == indent && starts_seq_item(content(li_))) val = block_seq(indent);
                else { val = std::make_shared<Node>(); val->line = (int)li_; }
            } else val = inline_value(rest, indent);
            if (!anchor.empty()) anchors_[anchor] = val;
            n->map.emplace_back(key, val);
        }
        return n;
    }

    // `&name` in front of a value: returns the name and removes it from `rest`
    std::string take_anchor(std::string& rest) {
        if (rest.empty() || rest[0] != '&') return "";
        size_t e = 1;
        while (e < rest.size() && !is_blank_ch(rest[e]) && !strchr("[]{},", rest[e])) ++e;
        std::string name = rest.substr(1, e - 1);
        if (name.empty()) fail("empty anchor name", (int)li_);
        rest = strip(rest.substr(e));
        return name;
    }

    // a value written on the key's line: alias, flow collection (may continue on the following lines), quoted or plain scalar
    NodeP inline_value(std::string text, int parent_indent) {
        const int line = (int)li_;                                       // (li_ already points past the line the text came from)
        std::string anchor = take_anchor(text);
        NodeP n;
        if (text.empty()) { n = std::make_shared<Node>(); n->line = line; }
        else if (text[0] == '!') fail("tags are not supported", line);
        else if (text[0] == '|' || text[0] == '>') fail("block scalars are not supported", line);
        else if (text[0] == '[' || text[0] == '{') {
            // pull in the following lines until the brackets balance
            while (!balanced(text)) {
                if (li_ >= lines_.size()) fail("unterminated flow collection", line);
                text += " " + content(li_);
                ++li_;
            }
            size_t pos = 0;
            n = flow_value(text, pos, line);
            skip_ws(text, pos);
            if (pos != text.size()) fail("unexpected text after a flow collection", line);
        } else {
            size_t used = 0;
            n = flow_scalar(text, used, false);
            n->line = line;
            if (used != text.size()) fail("unexpected text after a scalar", line);
            // a plain scalar continued on more-indented lines is one multi-line scalar in YAML: not supported here
            size_t save = li_;
            skip_blank();
            if (li_ < lines_.size() && !is_marker(lines_[li_], "---") && !is_marker(lines_[li_], "...") && indent_of(lines_[li_]) > parent_indent)
                fail("multi-line scalars are not supported");
            li_ = save;
        }
        if (!anchor.empty()) anchors_[anchor] = n;
        return n;
    }

    static bool balanced(const std::string& s) {
        int depth = 0; char q = 0;
        for (size_t i = 0; i < s.size(); ++i) {
            char c = s[i];
            if (q) {
                if (q == '"' && c == '\\') ++i;
                else if (c == q) { if (q == '\'' && i + 1 < s.size() && s[i + 1] == '\'') ++i; else q = 0; }
            } else if (c == '"' || c == '\'') q = c;
            else if (c == '[' || c == '{') ++depth;
            else if (c == ']' || c == '}') --depth;
        }
        return depth <= 0 && !q;
    }
    static void skip_ws(const std::string& s, size_t& p) { while (p < s.size() && is_blank_ch(s[p])) ++p; }

    NodeP flow_value(const std::string& s, size_t& p, int line) {
        skip_ws(s, p);
        if (p >= s.size()) fail("unexpected end of a flow collection", line);
        std::string anchor;
        if (s[p] == '&') {
            size_t e = p + 1;
            while (e < s.size() && !is_blank_ch(s[e]) && !strchr("[]{},", s[e])) ++e;
            anchor = s.substr(p + 1, e - p - 1);
            p = e; skip_ws(s, p);
        }
        NodeP n;
        if (p < s.size() && s[p] == '[') {
            n = std::make_shared<Node>(); n->kind = Node::Seq; n->line = line;
            ++p;
            while (true) {
                skip_ws(s, p);
                if (p >= s.size()) fail("unterminated flow sequence", line);
                if (s[p] == ']') { ++p; break; }
                n->seq.push_back(flow_value(s, p, line));
                skip_ws(s, p);
                if (p < s.size() && s[p] == ',') { ++p; continue; }
                if (p < s.size() && s[p] == ']') { ++p; break; }
                if (p < s.size() && s[p] == ':') fail("single-pair mappings inside a flow sequence are not supported", line);
                fail("expected `,` or `]` in a flow sequence", line);
            }
        } else if (p < s.size() && s[p] == '{') {
            n = std::make_shared<Node>(); n->kind = Node::Map; n->line = line;
            ++p;
            while (true) {
                skip_ws(s, p);
                if (p >= s.size()) fail("unterminated flow mapping", line);
                if (s[p] == '}') { ++p; break; }
                NodeP key = flow_value(s, p, line);
                skip_ws(s, p);
                NodeP val = std::make_shared<Node>(); val->line = line;
                if (p < s.size() && s[p] == ':') {
                    ++p; skip_ws(s, p);
                    if (p < s.size() && s[p] != ',' && s[p] != '}') val = flow_value(s, p, line);
                }
                n->map.emplace_back(key, val);
                skip_ws(s, p);
                if (p < s.size() && s[p] == ',') { ++p; continue; }
                if (p < s.size() && s[p] == '}') { ++p; break; }
                fail("expected `,` or `}` in a flow mapping", line);
            }
        } else if (p < s.size() && s[p] == '!') fail("tags are not supported", line);
        else {
            size_t used = 0;
            n = flow_scalar(s.substr(p), used, true);
            n->line = line;
            p += used;
        }
        if (!anchor.empty()) anchors_[anchor] = n;
        return n;
    }

    // one scalar at the start of `s`: alias, quoted, or plain (in a flow collection a plain scalar ends at `,[]{}` and at `: `)
    NodeP flow_scalar(const std::string& s, size_t& used, bool in_flow) {
        NodeP n = std::make_shared<Node>();
        if (s.empty()) { used = 0; return n; }
        if (s[0] == '*') {
            size_t e = 1;
            while (e < s.size() && !is_blank_ch(s[e]) && !(in_flow && strchr("[]{},", s[e]))) ++e;
            auto it = anchors_.find(s.substr(1, e - 1));
            if (it == anchors_.end()) fail("unknown anchor `" + s.substr(1, e - 1) + "`", (int)li_);
            used = e;
            return it->second;
        }
        if (s[0] == '"' || s[0] == '\'') {
            const char q = s[0];
            std::string out;
            size_t i = 1;
            for (;; ++i) {
                if (i >= s.size()) fail("unterminated quoted scalar (multi-line quoted scalars are not supported)", (int)li_);
                char c = s[i];
                if (q == '\'') {
                    if (c == '\'') { if (i + 1 < s.size() && s[i + 1] == '\'') { out += '\''; ++i; continue; } break; }
                    out += c;
                } else {
                    if (c == '"') break;
                    if (c != '\\') { out += c; continue; }
                    if (++i >= s.size()) fail("unterminated escape", (int)li_);
                    switch (s[i]) {
                        case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break; case '0': out += '\0'; break;
                        case '\\': out += '\\'; break; case '"': out += '"'; break; case '/': out += '/'; break; case ' ': out += ' '; break;
                        case 'a': out += '\a'; break; case 'b': out += '\b'; break; case 'e': out += '\x1b'; break; case 'f': out += '\f'; break; case 'v': out += '\v'; break;
                        case 'x': case 'u': case 'U': {
                            const int nd = s[i] == 'x' ? 2 : (s[i] == 'u' ? 4 : 8);
                            if (i + nd >= s.size()) fail("bad escape", (int)li_);
                            unsigned long cp = strtoul(s.substr(i + 1, nd).c_str(), nullptr, 16);
                            i += nd;
                            if (cp < 0x80) out += (char)cp;
                            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                            else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                            else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                            break;
                        }
                        default: fail("unknown escape in a double-quoted scalar", (int)li_);
                    }
                }
            }
            used = i + 1;
            n->kind = Node::Scalar; n->scalar = out; n->quoted = true;
            return n;
        }
        size_t e = 0;
        for (; e < s.size(); ++e) {
            char c = s[e];
            if (in_flow && strchr(",[]{}", c)) break;
            if (c == ':' && (e + 1 == s.size() || is_blank_ch(s[e + 1]) || (in_flow && strchr(",[]{}", s[e + 1])))) break;
        }
        std::string text = strip(s.substr(0, e));
        used = e;
        if (text.empty() || text == "~" || text == "null" || text == "Null" || text == "NULL") { n->kind = Node::Null; n->scalar = text; return n; }
        n->kind = Node::Scalar; n->scalar = text;
        return n;
    }
};

// ------------------------------------------------------------------ serde's view of the tree

[[noreturn]] inline void type_error(const std::string& what, const Node* n) {
    std::ostringstream o;
    o << what;
    if (n) o << " at line " << n->line;
    throw ParseError{o.str()};
}

inline bool digits_but_not_number(const std::string& s) {            // "007": a string in YAML 1.2 (serde_yaml)
    size_t a = (!s.empty() && (s[0] == '-' || s[0] == '+')) ? 1 : 0;
    if (s.size() - a <= 1 || s[a] != '0') return false;
    for (size_t i = a + 1; i < s.size(); ++i) if (s[i] < '0' || s[i] > '9') return false;
    return true;
}

// Rust's f64::from_str grammar (finite results only are numbers for serde_yaml)
inline bool rust_float_syntax(const std::string& s) {
    size_t i = 0, n = s.size();
    if (i < n && (s[i] == '+' || s[i] == '-')) ++i;
    size_t d0 = i; while (i < n && isdigit((unsigned char)s[i])) ++i;
    size_t int_digits = i - d0, frac_digits = 0;
    if (i < n && s[i] == '.') { ++i; size_t f0 = i; while (i < n && isdigit((unsigned char)s[i])) ++i; frac_digits = i - f0; }
    if (int_digits + frac_digits == 0) return false;
    if (i < n && (s[i] == 'e' || s[i] == 'E')) {
        ++i;
        if (i < n && (s[i] == '+' || s[i] == '-')) ++i;
        size_t e0 = i; while (i < n && isdigit((unsigned char)s[i])) ++i;
        if (i == e0) return false;
    }
    return i == n;
}

inline float as_f32(const Node* n, const char* field) {
    if (!n) type_error(std::string("missing field `") + field + "`", nullptr);
    if (n->kind == Node::Null) type_error(std::string("`") + field + "`: invalid type: unit value, expected f32", n);
    if (n->kind != Node::Scalar) type_error(std::string("`") + field + "`: invalid type: " + (n->kind == Node::Map ? "map" : "sequence") + ", expected f32", n);
    const std::string& s = n->scalar;
    auto bad = [&]() { type_error(std::string("`") + field + "`: invalid type: string \"" + s + "\", expected f32", n); };
    if (n->quoted) bad();
    if (s == "true" || s == "True" || s == "TRUE" || s == "false" || s == "False" || s == "FALSE") type_error(std::string("`") + field + "`: invalid type: boolean, expected f32", n);
    // integers: [+-]? then 0x / 0o / 0b digits, or decimal digits without a leading zero
    {
        std::string u = s;
        bool neg = false;
        if (!u.empty() && (u[0] == '+' || u[0] == '-')) { neg = u[0] == '-'; u = u.substr(1); }
        int radix = 10;
        if (u.compare(0, 2, "0x") == 0) radix = 16; else if (u.compare(0, 2, "0o") == 0) radix = 8; else if (u.compare(0, 2, "0b") == 0) radix = 2;
        std::string digits = radix == 10 ? u : u.substr(2);
        bool ok = !digits.empty() && !(radix == 10 && digits_but_not_number(s));
        for (char c : digits) {
            int v = isdigit((unsigned char)c) ? c - '0' : (isalpha((unsigned char)c) ? tolower((unsigned char)c) - 'a' + 10 : 99);
            if (v >= radix) ok = false;
        }
        if (ok) {
            // exact value as a long double sum would round twice; accumulate in unsigned __int128 (serde_yaml tries up to i128 / u128)
            unsigned __int128 acc = 0; bool overflow = false;
            for (char c : digits) {
                int v = isdigit((unsigned char)c) ? c - '0' : tolower((unsigned char)c) - 'a' + 10;
                if (acc > (~(unsigned __int128)0 - v) / radix) { overflow = true; break; }
                acc = acc * radix + v;
            }
            if (!overflow) { float f = (float)acc; return (neg && acc != 0) ? -f : f; }   // integer -> f32, rounded once (`v as f32`); "-0" is the integer 0
            if (radix != 10) bad();
        }
    }
    if (digits_but_not_number(s)) bad();
    std::string u = s;
    if (!u.empty() && u[0] == '+') { u = u.substr(1); if (!u.empty() && (u[0] == '+' || u[0] == '-')) bad(); }
    if (u == ".inf" || u == ".Inf" || u == ".INF") return INFINITY;
    if (s == "-.inf" || s == "-.Inf" || s == "-.INF") return -INFINITY;
    if (s == ".nan" || s == ".NaN" || s == ".NAN") return NAN;
    if (!rust_float_syntax(u)) bad();
    double d = strtod(u.c_str(), nullptr);                                        // correctly rounded to f64 ...
    if (!std::isfinite(d)) bad();                                                 // (serde_yaml: an overflowing literal stays a string)
    return (float)d;                                                              // ... then `as f32`
}

inline const Node* map_get(const Node* m, const char* key) {
    const Node* found = nullptr;
    for (auto& kv : m->map)
        if (kv.first->kind == Node::Scalar && kv.first->scalar == key) {
            if (found) type_error(std::string("duplicate field `") + key + "`", kv.first.get());
            found = kv.second.get();
        }
    return found;
}

// a struct: a mapping with the named fields, or the sequence of its fields in declaration order
struct Fields {
    const Node* n; std::vector<const char*> names; const char* what;
    Fields(const Node* node, std::vector<const char*> nm, const char* w) : n(node), names(std::move(nm)), what(w) {
        if (!n) type_error(std::string("missing field `") + what + "`", nullptr);
        if (n->kind != Node::Map && n->kind != Node::Seq) type_error(std::string("`") + what + "`: invalid type: " + (n->kind == Node::Null ? "unit value" : "scalar") + ", expected struct", n);
        // (serde's derived visit_seq wants every field, the Options too — they may be null)
        if (n->kind == Node::Seq && n->seq.size() != names.size()) type_error(std::string("`") + what + "`: invalid length " + std::to_string(n->seq.size()) + ", expected struct with " + std::to_string(names.size()) + " elements", n);
    }
    const Node* get(const char* key, bool required = true) const {
        const Node* f = nullptr;
        if (n->kind == Node::Map) f = map_get(n, key);
        else for (size_t i = 0; i < names.size(); ++i) if (!strcmp(names[i], key) && i < n->seq.size()) f = n->seq[i].get();
        if (!f && required) {
            type_error(std::string("`") + what + "`: missing field `" + key + "`", n);
        }
        return f;
    }
};

inline rbrt_vec3 as_vec3(const Node* n, const char* what) {
    Fields f(n, {"x", "y", "z"}, what);
    rbrt_vec3 v;
    v.x = as_f32(f.get("x"), "x"); v.y = as_f32(f.get("y"), "y"); v.z = as_f32(f.get("z"), "z");
    return v;
}
inline std::string as_string(const Node* n, const char* what) {
    if (!n) type_error(std::string("missing field `") + what + "`", nullptr);
    if (n->kind == Node::Null) return n->scalar;                                   // `~` read as a String is the text "~"
    if (n->kind != Node::Scalar) type_error(std::string("`") + what + "`: invalid type: " + (n->kind == Node::Map ? "map" : "sequence") + ", expected a string", n);
    return n->scalar;
}
inline const std::vector<NodeP>& as_seq(const Node* n, const char* what) {
    if (!n) type_error(std::string("missing field `") + what + "`", nullptr);
    if (n->kind != Node::Seq) type_error(std::string("`") + what + "`: invalid type: " + (n->kind == Node::Null ? "unit value" : (n->kind == Node::Map ? "map" : "scalar")) + ", expected a sequence", n);
    return n->seq;
}

// ------------------------------------------------------------------ blueprints.rs:15-48
struct MaterialBp { std::string type; bool has_albedo = false; rbrt_vec3 albedo{0, 0, 0}; bool has_param = false; float param = 0; };
struct MeshBp { std::string obj; float scale; rbrt_vec3 translation, rotation; MaterialBp mat; };
struct SphereBp { float radius; rbrt_vec3 center; MaterialBp mat; };
struct SceneBp { rbrt_vec3 up, look_at, position; float focal; std::vector<MeshBp> meshes; std::vector<SphereBp> spheres; };

inline MaterialBp material_bp(const Fields& f) {
    MaterialBp m;
    m.type = as_string(f.get("material_type"), "material_type");
    if (const Node* a = f.get("albedo", false)) if (a->kind != Node::Null) { m.has_albedo = true; m.albedo = as_vec3(a, "albedo"); }
    if (const Node* p = f.get("material_param", false)) if (p->kind != Node::Null) { m.has_param = true; m.param = as_f32(p, "material_param"); }
    return m;
}

inline SceneBp scene_blueprint(const Node* root) {
    SceneBp bp;
    Fields top(root, {"camera_blueprint", "mesh_blueprints", "sphere_blueprints"}, "SceneBlueprint");
    Fields cam(top.get("camera_blueprint"), {"camera_up", "camera_look_at", "camera_position", "camera_focal_length_mm"}, "camera_blueprint");
    bp.up = as_vec3(cam.get("camera_up"), "camera_up");
    bp.look_at = as_vec3(cam.get("camera_look_at"), "camera_look_at");
    bp.position = as_vec3(cam.get("camera_position"), "camera_position");
    bp.focal = as_f32(cam.get("camera_focal_length_mm"), "camera_focal_length_mm");
    for (auto& it : as_seq(top.get("mesh_blueprints"), "mesh_blueprints")) {
        Fields f(it.get(), {"obj_filepath", "scale", "translation", "rotation_rad", "material_type", "albedo", "material_param"}, "TriangleMeshBlueprint");
        MeshBp m;
        m.obj = as_string(f.get("obj_filepath"), "obj_filepath");
        m.scale = as_f32(f.get("scale"), "scale");
        m.translation = as_vec3(f.get("translation"), "translation");
        m.rotation = as_vec3(f.get("rotation_rad"), "rotation_rad");
        m.mat = material_bp(f);
        bp.meshes.push_back(m);
    }
    for (auto& it : as_seq(top.get("sphere_blueprints"), "sphere_blueprints")) {
        Fields f(it.get(), {"radius", "center", "material_type", "albedo", "material_param"}, "SphereBlueprint");
        SphereBp s;
        s.radius = as_f32(f.get("radius"), "radius");
        s.center = as_vec3(f.get("center"), "center");
        s.mat = material_bp(f);
        bp.spheres.push_back(s);
    }
    return bp;
}

inline SceneBp parse_scene(const std::string& text) {
    Parser p(text);
    NodeP root = p.document();
    return scene_blueprint(root.get());
}

}  // namespace scene_yaml

// obj_loader.cpp — load_mesh_vertices_from_file (mesh.rs:78-121) behind the C-ABI: the step before scene upload.
//
// The reference reads the .obj with tobj `^4.0.2` (rbrt_lib/Cargo.toml:12, no lockfile; `tobj::load_obj` with default
// LoadOptions, mesh.rs:84-88), walks every model's `mesh.indices` in triples (mesh.rs:92-107: `indices.len() / 3`, whatever
// the faces were), scales, rotates and translates every corner in f32 (mesh.rs:102-112) and panics when the load failed
// (`assert!(loaded_mesh.is_ok())`, mesh.rs:89).  tobj is not under /root/reference and no reference test pins it (SURVEY.md
// §8c: parity unpinned), so what follows restates tobj 4's published behaviour for default options:
//   * lines are split at '\n' (a trailing '\r' dropped), words at white space, the first word selects the record;
//   * `v x y z [...]`: three f32 (Rust's `f32::from_str`: correctly rounded, optional sign, inf / nan), fewer or a bad one fail
//     the load; `vt u v`, `vn x y z` likewise (only counted here: face records may refer to them);
//   * `f` AND `l` records: every word is `v[/vt[/vn]]` (1-based, negative = relative to the count so far, empty = absent, a
//     fourth field or a non-integer fails the load); without `triangulate` the position indices of a record are appended as
//     they stand, 1 (point), 2 (line), 3, 4 or n of them;
//   * `o` / `g`: when face records are pending they become a model; `usemtl name`: likewise when the material ID changes
//     (IDs come from the `newmtl` names of the `mtllib` files loaded so far, relative to the .obj; a library that cannot be
//     read or parsed contributes none; unknown names are "no material"); the remaining records become the last model;
//   * exporting a model checks every index against the positions / texcoords / normals read SO FAR and fails the load on
//     one out of bounds (index 0 wraps to usize::MAX and fails too);
//   * anything else (`#`, `s`, `vp`, ...) is ignored.
// rbrt then cuts each model's index list into triples and drops the remainder.  A file of triangles only gives the same
// soup however it is split into models; the rules above matter for files with lines, quads or polygons, where the reference
// produces the triangles a drop-in must produce too.
//
// B200-side reason to have this in the library: a C3-sized .obj is ~50 MB of text; parsed line by line on one core it costs
// 100x the 33 ms the frame takes to render.  The file is mmap'ed and cut at line ends into one piece per host thread; the
// pieces are parsed concurrently (std::from_chars for the floats), then stitched serially: relative indices get the piece's
// base, model boundaries are resolved, and the gather + transform of the corners runs on all threads again.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <new>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "../../include/rbrt_gpu.h"

namespace rbrt { void set_error(const char* fmt, ...); }

namespace {

inline bool is_ws(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// Rust's f32::from_str: [+-]? ( "inf" | "infinity" | "nan" | decimal with optional exponent ), whole token, correctly rounded
bool parse_f32(const char* b, const char* e, float* out) {
    if (b == e) return false;
    if (*b == '+') { ++b; if (b == e || *b == '-' || *b == '+') return false; }
    for (const char* p = b; p < e; ++p) if (*p == '(' || *p == 'x' || *p == 'X') return false;   // from_chars would take nan(...); never hex
    auto r = std::from_chars(b, e, *out, std::chars_format::general);
    if (r.ec == std::errc() && r.ptr == e) return true;
    if (r.ec == std::errc::result_out_of_range && r.ptr == e) {                                // Rust: +-inf on overflow, +-0 / subnormal on underflow
        std::string s(b, e);
        char* end = nullptr;
        *out = strtof(s.c_str(), &end);
        return end == s.c_str() + s.size();
    }
    return false;
}

// Rust's isize::from_str: [+-]? digits, whole token, no overflow
bool parse_isize(const char* b, const char* e, int64_t* out) {
    if (b == e) return false;
    bool neg = false;
    if (*b == '+' || *b == '-') { neg = *b == '-'; ++b; if (b == e) return false; }
    uint64_t v = 0;
    for (; b < e; ++b) {
        if (*b < '0' || *b > '9') return false;
        if (v > (uint64_t)INT64_MAX / 10) return false;
        v = v * 10 + (uint64_t)(*b - '0');
        if (v > (uint64_t)INT64_MAX + (neg ? 1 : 0)) return false;
    }
    *out = neg ? (int64_t)(0 - v) : (int64_t)v;
    return true;
}

constexpr int64_t OUT_OF_BOUNDS = INT64_MAX;                 // an index tobj wraps to a huge usize

enum EvKind { EV_GROUP, EV_USEMTL, EV_MTLLIB };
struct Event { EvKind kind; uint64_t idx_at, faces_at, pos_at, vt_at, vn_at; std::string name; };

struct Piece {
    const char *begin, *end;
    std::vector<float> pos;
    std::vector<int64_t> idx, vt_idx, vn_idx;                // per face corner; relative ones still lack the piece's base; -1 = absent
    std::vector<uint64_t> rel, rel_vt, rel_vn;               // corners whose index is relative
    std::vector<Event> events;
    uint64_t faces = 0, n_vt = 0, n_vn = 0;
    std::string error;                                       // first failure in this piece
    uint64_t error_line = 0;                                 // line within the piece
};

struct Tok { const char *b, *e; };
inline bool next_tok(const char*& p, const char* le, Tok* t) {
    while (p < le && is_ws(*p)) ++p;
    if (p == le) return false;
    t->b = p;
    while (p < le && !is_ws(*p)) ++p;
    t->e = p;
    return true;
}
inline bool tok_is(const Tok& t, const char* s) { size_t n = strlen(s); return (size_t)(t.e - t.b) == n && !memcmp(t.b, s, n); }

// floats of a `v` / `vt` / `vn` record: the first n words must parse, the rest of the line is not looked at
// (tobj tries to read vertex colours from it and ignores a failure)
bool take_floats(const char*& p, const char* le, int n, float* out) {
    Tok t;
    for (int i = 0; i < n; ++i) { if (!next_tok(p, le, &t) || !parse_f32(t.b, t.e, out + i)) return false; }
    return true;
}

std::string rest_of_line(const char* p, const char* le) {
    while (p < le && is_ws(*p)) ++p;
    while (le > p && is_ws(le[-1])) --le;
    return std::string(p, le);
}

void parse_piece(Piece& pc) {
    const char* p = pc.begin;
    uint64_t line_no = 0;
    auto fail = [&](const char* what) { pc.error = what; pc.error_line = line_no; };
    while (p < pc.end && pc.error.empty()) {
        const char* le = (const char*)memchr(p, '\n', (size_t)(pc.end - p));
        const char* next = le ? le + 1 : pc.end;
        if (!le) le = pc.end;
        ++line_no;
        const char* q = p;
        Tok t;
        if (next_tok(q, le, &t)) {
            if (tok_is(t, "v")) {
                float v[3];
                if (!take_floats(q, le, 3, v)) fail("a `v` record needs three numbers (tobj: PositionParseError)");
                else pc.pos.insert(pc.pos.end(), v, v + 3);
            } else if (tok_is(t, "vt")) {
                float v[2];
                if (!take_floats(q, le, 2, v)) fail("a `vt` record needs two numbers (tobj: TexcoordParseError)"); else ++pc.n_vt;
            } else if (tok_is(t, "vn")) {
                float v[3];
                if (!take_floats(q, le, 3, v)) fail("a `vn` record needs three numbers (tobj: NormalParseError)"); else ++pc.n_vn;
            } else if (tok_is(t, "f") || tok_is(t, "l")) {
                Tok w;
                while (next_tok(q, le, &w)) {
                    int64_t field[3] = {-1, -1, -1}; bool relative[3] = {false, false, false};
                    int k = 0; const char* fb = w.b;
                    for (;; ++k) {
                        const char* fe = (const char*)memchr(fb, '/', (size_t)(w.e - fb));
                        if (!fe) fe = w.e;
                        if (fe > fb) {
                            int64_t x;
                            if (k > 2 || !parse_isize(fb, fe, &x)) { fail("bad face corner (tobj: FaceParseError)"); break; }
                            const uint64_t count = k == 0 ? pc.pos.size() / 3 : (k == 1 ? pc.n_vt : pc.n_vn);
                            if (x < 0) { field[k] = (int64_t)count + x; relative[k] = true; }      // + the piece's base, later
                            else field[k] = x == 0 ? (k == 0 ? OUT_OF_BOUNDS : -1) : x - 1;   // 0 wraps to usize::MAX: no position / tobj's "absent" marker
                        }
                        if (fe == w.e) break;
                        fb = fe + 1;
                        if (fb == w.e) break;                                                    // "1/" : an empty last field
                    }
                    if (!pc.error.empty()) break;
                    if (field[0] == -1 && !relative[0]) field[0] = OUT_OF_BOUNDS;                // "/1": no position index (tobj: usize::MAX)
                    if (relative[0]) pc.rel.push_back(pc.idx.size());
                    if (relative[1]) pc.rel_vt.push_back(pc.idx.size());
                    if (relative[2]) pc.rel_vn.push_back(pc.idx.size());
                    const bool has_vt = field[1] != -1 || relative[1], has_vn = field[2] != -1 || relative[2];
                    pc.idx.push_back(field[0]);
                    // texcoord / normal indices are only bounds-checked; kept sparse: most files have none or few
                    if (has_vt || !pc.vt_idx.empty()) { pc.vt_idx.resize(pc.idx.size() - 1, -1); pc.vt_idx.push_back(has_vt ? field[1] : -1); }
                    if (has_vn || !pc.vn_idx.empty()) { pc.vn_idx.resize(pc.idx.size() - 1, -1); pc.vn_idx.push_back(has_vn ? field[2] : -1); }
                }
                ++pc.faces;
            } else if (tok_is(t, "o") || tok_is(t, "g")) {
                pc.events.push_back(Event{EV_GROUP, pc.idx.size(), pc.faces, pc.pos.size() / 3, pc.n_vt, pc.n_vn, std::string()});
            } else if (tok_is(t, "usemtl") || tok_is(t, "mtllib")) {
                std::string name = rest_of_line(q, le);
                if (tok_is(t, "usemtl") && name.empty()) fail("`usemtl` without a name (tobj: MaterialParseError)");
                else pc.events.push_back(Event{tok_is(t, "usemtl") ? EV_USEMTL : EV_MTLLIB, pc.idx.size(), pc.faces, pc.pos.size() / 3, pc.n_vt, pc.n_vn, name});
            }
        }
        p = next;
    }
}

// `newmtl` names of a material library, in order; false when tobj's load_mtl would fail (the library then contributes nothing)
bool mtl_names(const std::string& path, std::vector<std::string>* names) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::string data; char buf[65536]; size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.append(buf, n);
    fclose(f);
    bool ok = true;
    const char *p = data.data(), *end = p + data.size();
    while (p < end && ok) {
        const char* le = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* next = le ? le + 1 : end;
        if (!le) le = end;
        const char* q = p; Tok t;
        if (next_tok(q, le, &t)) {
            if (tok_is(t, "newmtl")) {
                std::string name = rest_of_line(q, le);
                if (name.empty()) ok = false; else names->push_back(name);
            } else if (tok_is(t, "Ka") || tok_is(t, "Kd") || tok_is(t, "Ks")) { float v[3]; ok = take_floats(q, le, 3, v); }
            else if (tok_is(t, "Ns") || tok_is(t, "Ni") || tok_is(t, "d")) { float v[1]; ok = take_floats(q, le, 1, v); }
            else if (tok_is(t, "illum")) { Tok w; int64_t x; ok = next_tok(q, le, &w) && parse_isize(w.b, w.e, &x) && x >= 0 && x <= 255; }
            else if (tok_is(t, "map_Ka") || tok_is(t, "map_Kd") || tok_is(t, "map_Ks") || tok_is(t, "map_Ns") || tok_is(t, "map_Bump") || tok_is(t, "map_bump") ||
                     tok_is(t, "bump") || tok_is(t, "map_d")) ok = !rest_of_line(q, le).empty();
        }
        p = next;
    }
    return ok;
}

// the scalar coefficients of Vec3::rotate_point (vec3.rs:139-155), evaluated per vertex exactly as rbrt_transform_vertices does
inline void transform_range(float* xyz, uint64_t n_vertices, float scale, const float sc[6], rbrt_vec3 tr) {
    const float s_x = sc[0], s_y = sc[1], s_z = sc[2], c_x = sc[3], c_y = sc[4], c_z = sc[5];
    for (uint64_t i = 0; i < n_vertices; ++i) {
        float x = xyz[3 * i] * scale, y = xyz[3 * i + 1] * scale, z = xyz[3 * i + 2] * scale;   // mesh.rs:102-106
        float rx = (c_x * c_z - c_y * s_x * s_z) * x - (c_x * s_z + c_y * c_z * s_x) * y + s_x * s_y * z;   // vec3.rs:150-152
        float ry = (c_z * s_x + c_x * c_y * s_z) * x + (c_x * c_y * c_z - s_x * s_z) * y - c_x * s_y * z;
        float rz = s_y * s_z * x + c_z * s_y * y + c_y * z;
        xyz[3 * i] = rx + tr.x; xyz[3 * i + 1] = ry + tr.y; xyz[3 * i + 2] = rz + tr.z;          // mesh.rs:108-112
    }
}

unsigned host_threads() {
    unsigned n = std::thread::hardware_concurrency();
    if (const char* e = getenv("RBRT_HOST_THREADS")) { int v = atoi(e); if (v > 0) n = (unsigned)v; }
    return n ? std::min(n, 64u) : 1u;
}

// body(begin, end) on `threads` host threads.  An exception in a worker (std::bad_alloc) is carried to the caller; when a thread
// cannot be started the rest of the range runs on the calling thread.
template <class F> void parallel_for(uint64_t n, unsigned threads, F&& body) {
    if (threads <= 1 || n < 2) { body((uint64_t)0, n); return; }
    threads = (unsigned)std::min<uint64_t>(threads, n);
    std::vector<std::thread> pool;
    pool.reserve(threads);                                                 // (so that only these two allocations can fail before any work has started)
    std::vector<std::exception_ptr> errors(threads);
    unsigned started = 0;
    for (; started + 1 < threads; ++started) {
        const unsigned t = started;
        try {
            pool.emplace_back([&, t] { try { body(n * t / threads, n * (t + 1) / threads); } catch (...) { errors[t] = std::current_exception(); } });
        } catch (const std::system_error&) { break; }
    }
    try { body(n * started / threads, n); } catch (...) { errors[threads - 1] = std::current_exception(); }    // the caller takes the last share (and what no thread took)
    for (auto& th : pool) th.join();
    for (auto& e : errors) if (e) std::rethrow_exception(e);
}

}  // namespace

extern "C" int rbrt_transform_vertices(float* xyz, uint64_t n_vertices, float scale, rbrt_vec3 rot, rbrt_vec3 tr) {
    if (n_vertices && !xyz) return RBRT_E_INVALID;
    const float sc[6] = {sinf(rot.x), sinf(rot.y), sinf(rot.z), cosf(rot.x), cosf(rot.y), cosf(rot.z)};
    // every vertex is independent: large soups are split over the host threads (same arithmetic per vertex)
    try {
        parallel_for(n_vertices, n_vertices >= (1u << 16) ? host_threads() : 1, [&](uint64_t b, uint64_t e) { transform_range(xyz + 3 * b, e - b, scale, sc, tr); });
    } catch (...) { transform_range(xyz, n_vertices, scale, sc, tr); }               // transform_range cannot throw: this is parallel_for failing to allocate, before any vertex was touched
    return RBRT_OK;
}

extern "C" void rbrt_mesh_free(float* tri_vertices) { free(tri_vertices); }

static int load_obj(const char* filepath, rbrt_vec3 translation, rbrt_vec3 rotation_rad, float scale, float** tri_vertices_out, uint64_t* num_triangles_out);

extern "C" int rbrt_mesh_load_obj(const char* filepath, rbrt_vec3 translation, rbrt_vec3 rotation_rad, float scale,
                                  float** tri_vertices_out, uint64_t* num_triangles_out) {
    try {                                                                             // the C-ABI never unwinds
        return load_obj(filepath, translation, rotation_rad, scale, tri_vertices_out, num_triangles_out);
    } catch (const std::bad_alloc&) {
        rbrt::set_error("out of host memory while reading %s", filepath ? filepath : "(null)");
    } catch (const std::exception& e) {
        rbrt::set_error("%s: %s", filepath ? filepath : "(null)", e.what());
    } catch (...) {
        rbrt::set_error("%s: unknown failure", filepath ? filepath : "(null)");
    }
    if (tri_vertices_out && *tri_vertices_out) { free(*tri_vertices_out); *tri_vertices_out = nullptr; }
    if (num_triangles_out) *num_triangles_out = 0;
    return RBRT_E_ALLOC;
}

static int load_obj(const char* filepath, rbrt_vec3 translation, rbrt_vec3 rotation_rad, float scale, float** tri_vertices_out, uint64_t* num_triangles_out) {
    if (!filepath || !tri_vertices_out || !num_triangles_out) { rbrt::set_error("null argument"); return RBRT_E_INVALID; }
    *tri_vertices_out = nullptr; *num_triangles_out = 0;
    int fd = open(filepath, O_RDONLY);
    if (fd < 0) { rbrt::set_error("cannot open %s (the reference panics: assertion failed: loaded_mesh.is_ok(), mesh.rs:89)", filepath); return RBRT_E_INVALID; }
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); rbrt::set_error("%s is not a regular file", filepath); return RBRT_E_INVALID; }
    const size_t size = (size_t)st.st_size;
    const char* data = nullptr;
    if (size) {
        void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); rbrt::set_error("cannot map %s", filepath); return RBRT_E_ALLOC; }
        madvise(m, size, MADV_SEQUENTIAL);
        data = (const char*)m;
    }
    close(fd);
    struct Unmap { const char* p; size_t n; ~Unmap() { if (p) munmap((void*)p, n); } } unmap{data, size};

    // ---- pieces: cut at line ends, parsed concurrently
    const unsigned threads = host_threads();
    size_t piece_bytes = 1u << 20;                                        // below ~1 MB per thread the threads cost more than they save
    if (const char* e = getenv("RBRT_OBJ_PIECE_BYTES")) { long v = atol(e); if (v > 0) piece_bytes = (size_t)v; }   // (tests: many pieces of a small file)
    const size_t n_pieces = std::max<size_t>(1, std::min<size_t>(threads, size / piece_bytes));
    std::vector<Piece> pieces(n_pieces);
    {
        const char* cur = data;
        for (size_t i = 0; i < n_pieces; ++i) {
            const char* stop = data + size;
            if (i + 1 < n_pieces) {
                const char* want = data + size * (i + 1) / n_pieces;
                if (want < cur) want = cur;
                const char* nl = (const char*)memchr(want, '\n', (size_t)(data + size - want));
                stop = nl ? nl + 1 : data + size;
            }
            pieces[i].begin = cur; pieces[i].end = stop; cur = stop;
        }
    }
    parallel_for(n_pieces, (unsigned)n_pieces, [&](uint64_t b, uint64_t e) { for (uint64_t i = b; i < e; ++i) parse_piece(pieces[i]); });

    // ---- stitch: bases, the first failure in file order, model boundaries
    std::vector<uint64_t> pos_base(n_pieces + 1, 0), idx_base(n_pieces + 1, 0), vt_base(n_pieces + 1, 0), vn_base(n_pieces + 1, 0), face_base(n_pieces + 1, 0);
    uint64_t lines_before = 0;
    for (size_t i = 0; i < n_pieces; ++i) {
        const Piece& pc = pieces[i];
        if (!pc.error.empty()) {
            rbrt::set_error("%s line %llu: %s; the reference panics here (assertion failed: loaded_mesh.is_ok(), mesh.rs:89)", filepath,
                            (unsigned long long)(lines_before + pc.error_line), pc.error.c_str());
            return RBRT_E_INVALID;
        }
        lines_before += (uint64_t)std::count(pc.begin, pc.end, '\n');
        pos_base[i + 1] = pos_base[i] + pc.pos.size() / 3; idx_base[i + 1] = idx_base[i] + pc.idx.size();
        vt_base[i + 1] = vt_base[i] + pc.n_vt; vn_base[i + 1] = vn_base[i] + pc.n_vn; face_base[i + 1] = face_base[i] + pc.faces;
    }
    const uint64_t n_pos = pos_base[n_pieces], n_idx = idx_base[n_pieces];
    parallel_for(n_pieces, (unsigned)n_pieces, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            Piece& pc = pieces[i];
            for (uint64_t k : pc.rel) { pc.idx[k] += (int64_t)pos_base[i]; if (pc.idx[k] < 0) pc.idx[k] = OUT_OF_BOUNDS; }
            // (a relative texcoord / normal index that lands on -1 wraps to usize::MAX, which tobj reads as "absent")
            for (uint64_t k : pc.rel_vt) { pc.vt_idx[k] += (int64_t)vt_base[i]; if (pc.vt_idx[k] < -1) pc.vt_idx[k] = OUT_OF_BOUNDS; }
            for (uint64_t k : pc.rel_vn) { pc.vn_idx[k] += (int64_t)vn_base[i]; if (pc.vn_idx[k] < -1) pc.vn_idx[k] = OUT_OF_BOUNDS; }
        }
    });

    struct Model { uint64_t begin, end, n_pos, n_vt, n_vn; };              // corner range + what had been read when tobj exported it
    std::vector<Model> models;
    {
        std::map<std::string, uint64_t> mat_map; uint64_t n_materials = 0;
        int64_t mat_id = -1;                                               // -1 = None
        uint64_t model_begin = 0, faces_done = 0;
        std::string dir(filepath);
        { size_t s = dir.find_last_of('/'); dir = s == std::string::npos ? std::string() : dir.substr(0, s + 1); }
        for (size_t i = 0; i < n_pieces; ++i)
            for (const Event& ev : pieces[i].events) {
                const uint64_t at = idx_base[i] + ev.idx_at, faces = face_base[i] + ev.faces_at;
                bool split = false;
                if (ev.kind == EV_MTLLIB) {
                    std::vector<std::string> names;
                    if (!ev.name.empty() && mtl_names(dir + ev.name, &names)) {
                        for (size_t k = 0; k < names.size(); ++k) mat_map[names[k]] = n_materials + k;
                        n_materials += names.size();
                    }
                    continue;
                }
                if (ev.kind == EV_GROUP) split = faces > faces_done;
                else {
                    auto it = mat_map.find(ev.name);
                    const int64_t new_mat = it == mat_map.end() ? -1 : (int64_t)it->second;
                    split = new_mat != mat_id && faces > faces_done;
                    mat_id = new_mat;
                }
                if (split) {
                    models.push_back(Model{model_begin, at, pos_base[i] + ev.pos_at, vt_base[i] + ev.vt_at, vn_base[i] + ev.vn_at});
                    model_begin = at; faces_done = faces;
                }
            }
        models.push_back(Model{model_begin, n_idx, n_pos, vt_base[n_pieces], vn_base[n_pieces]});   // tobj always pushes the last one
    }

    // ---- one flat view of the corners and positions (pieces stay where they are; lookups go through the bases)
    auto piece_of = [&](const std::vector<uint64_t>& base, uint64_t g) { return (size_t)(std::upper_bound(base.begin(), base.end(), g) - base.begin()) - 1; };
    uint64_t n_tris = 0;
    std::vector<uint64_t> tri_base(models.size() + 1, 0);
    for (size_t m = 0; m < models.size(); ++m) { tri_base[m + 1] = tri_base[m] + (models[m].end - models[m].begin) / 3; }
    n_tris = tri_base[models.size()];

    // bounds of every exported corner (all of them, also the ones rbrt's `/ 3` drops: tobj fails before rbrt looks)
    for (size_t m = 0; m < models.size(); ++m) {
        const Model& md = models[m];
        for (uint64_t g = md.begin; g < md.end;) {
            const size_t pi = piece_of(idx_base, g);
            const Piece& pc = pieces[pi];
            const uint64_t lo = g - idx_base[pi], hi = std::min<uint64_t>(pc.idx.size(), md.end - idx_base[pi]);
            for (uint64_t k = lo; k < hi; ++k) {
                const bool bad = pc.idx[k] < 0 || (uint64_t)pc.idx[k] >= md.n_pos ||
                                 (md.n_vt && k < pc.vt_idx.size() && pc.vt_idx[k] != -1 && (uint64_t)pc.vt_idx[k] >= md.n_vt) ||   // only looked at when
                                 (md.n_vn && k < pc.vn_idx.size() && pc.vn_idx[k] != -1 && (uint64_t)pc.vn_idx[k] >= md.n_vn);     // the file has any
                if (bad) {
                    rbrt::set_error("%s: a face refers to a vertex, texcoord or normal that has not been read (tobj: Face*OutOfBounds); the reference panics "
                                    "here (assertion failed: loaded_mesh.is_ok(), mesh.rs:89)", filepath);
                    return RBRT_E_INVALID;
                }
            }
            g = idx_base[pi] + hi;
            if (hi == lo) break;
        }
    }

    if (!n_tris) return RBRT_OK;
    float* out = (float*)malloc(sizeof(float) * 9 * n_tris);
    if (!out) { rbrt::set_error("out of host memory for %llu triangles", (unsigned long long)n_tris); return RBRT_E_ALLOC; }
    *tri_vertices_out = out;                                              // (the wrapper frees it should the gather throw)
    const float sc[6] = {sinf(rotation_rad.x), sinf(rotation_rad.y), sinf(rotation_rad.z), cosf(rotation_rad.x), cosf(rotation_rad.y), cosf(rotation_rad.z)};
    parallel_for(n_tris, n_tris >= (1u << 14) ? threads : 1, [&](uint64_t tb, uint64_t te) {
        size_t m = (size_t)(std::upper_bound(tri_base.begin(), tri_base.end(), tb) - tri_base.begin()) - 1;
        for (uint64_t t = tb; t < te; ++t) {
            while (t >= tri_base[m + 1]) ++m;
            const uint64_t g0 = models[m].begin + 3 * (t - tri_base[m]);
            for (int k = 0; k < 3; ++k) {
                const uint64_t g = g0 + k;
                const size_t pi = piece_of(idx_base, g);
                const uint64_t v = (uint64_t)pieces[pi].idx[g - idx_base[pi]];
                const size_t pp = piece_of(pos_base, v);
                const float* src = pieces[pp].pos.data() + 3 * (v - pos_base[pp]);
                float* dst = out + 9 * t + 3 * k;
                dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
            }
        }
        transform_range(out + 9 * tb, 3 * (te - tb), scale, sc, translation);
    });
    *num_triangles_out = n_tris;
    return RBRT_OK;
}

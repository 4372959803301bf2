// loaders.cpp — host authoring that rides on the Mesh path: the OBJ and STL model loaders and the Mesh
// utilities the example scenes call on loaded models.  Behaviour follows PTSharpCore/OBJ.cs, STL.cs and Mesh.cs:141-289,
// quirks included (each is named where it is reproduced); nothing here traces a ray.
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

#include "ptsharp.hpp"

namespace ptsharp {

Vector Box::Anchor(const Vector& anchor) const { return Add(Min, Mul(Size(), anchor)); }  // Box.cs:48
Vector Matrix::MulDirection(const Vector& b) const {                                       // Matrix.cs:144-150
    return Normalize(Vector(m[0] * b.X() + m[1] * b.Y() + m[2] * b.Z(), m[4] * b.X() + m[5] * b.Y() + m[6] * b.Z(),
                            m[8] * b.X() + m[9] * b.Y() + m[10] * b.Z()));
}

// ---- Mesh utilities ------------------------------------------------------------------------------------------------
namespace {
// Dictionary<Vector, ...> keys: Vector equality is component-wise ==, so -0 and +0 are one key
struct VKey {
    uint32_t x, y, z;
    bool operator==(const VKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VKeyHash { size_t operator()(const VKey& k) const { return ((size_t)k.x * 0x9E3779B1u) ^ ((size_t)k.y * 0x85EBCA77u) ^ ((size_t)k.z * 0xC2B2AE3Du); } };
uint32_t key_bits(float f) { if (f == 0.0f) f = 0.0f; uint32_t u; std::memcpy(&u, &f, 4); return u; }
VKey key_of(const Vector& v) { return VKey{key_bits(v.x), key_bits(v.y), key_bits(v.z)}; }
void dirty(Mesh& m) { m.haveBox = false; m.tree.reset(); }  // Mesh.cs:28-32
}  // namespace

void Mesh::SmoothNormals() {
    std::unordered_map<VKey, Vector, VKeyHash> sum;
    for (const Triangle& t : Triangles) {  // accumulated in triangle order, like the reference's second loop
        sum[key_of(t.V1)] = Add(sum[key_of(t.V1)], t.N1);
        sum[key_of(t.V2)] = Add(sum[key_of(t.V2)], t.N2);
        sum[key_of(t.V3)] = Add(sum[key_of(t.V3)], t.N3);
    }
    for (auto& kv : sum) kv.second = Normalize(kv.second);
    for (Triangle& t : Triangles) { t.N1 = sum[key_of(t.V1)]; t.N2 = sum[key_of(t.V2)]; t.N3 = sum[key_of(t.V3)]; }
}

void Mesh::SmoothNormalsThreshold(double radians) {
    // Quirk kept (Mesh.cs:166-175): the candidate list of a vertex is NOT the normals meeting at it but a snapshot of the
    // running list of ALL N1 (resp. N2, N3) seen up to the last triangle that names the vertex as its V1 (resp. V2, V3).
    const double threshold = std::cos(radians);
    struct Snap { int list; size_t count; };
    std::unordered_map<VKey, Snap, VKeyHash> lookup;
    for (size_t i = 0; i < Triangles.size(); i++) {
        lookup[key_of(Triangles[i].V1)] = Snap{0, i + 1};
        lookup[key_of(Triangles[i].V2)] = Snap{1, i + 1};
        lookup[key_of(Triangles[i].V3)] = Snap{2, i + 1};
    }
    auto smooth = [&](const Vector& normal, const Snap& sn, const std::vector<Triangle>& src) {
        Vector result;
        for (size_t k = 0; k < sn.count; k++) {
            const Vector& x = sn.list == 0 ? src[k].N1 : sn.list == 1 ? src[k].N2 : src[k].N3;
            if ((double)Dot(x, normal) >= threshold) result = Add(result, x);
        }
        return Normalize(result);
    };
    const std::vector<Triangle> src = Triangles;  // the lists hold the normals as they were before the update
    for (Triangle& t : Triangles) {
        t.N1 = smooth(t.N1, lookup[key_of(t.V1)], src);
        t.N2 = smooth(t.N2, lookup[key_of(t.V2)], src);
        t.N3 = smooth(t.N3, lookup[key_of(t.V3)], src);
    }
}

void Mesh::Transform(const Matrix& matrix) {
    for (Triangle& t : Triangles) {
        t.V1 = matrix.MulPosition(t.V1); t.V2 = matrix.MulPosition(t.V2); t.V3 = matrix.MulPosition(t.V3);
        t.N1 = matrix.MulDirection(t.N1); t.N2 = matrix.MulDirection(t.N2); t.N3 = matrix.MulDirection(t.N3);
    }
    dirty(*this);
}

void Mesh::MoveTo(const Vector& position, const Vector& anchor) { Transform(Matrix::Translate(Sub(position, BoundingBox().Anchor(anchor)))); }

void Mesh::FitInside(const Box& box, const Vector& anchor) {
    const Box bb = BoundingBox();
    const Vector r = Div(box.Size(), bb.Size());
    const double scale = NetMin(NetMin(r.X(), r.Y()), r.Z());  // Vector.MinComponent (Vector.cs:491)
    const Vector extra = Sub(box.Size(), MulScalar(bb.Size(), scale));
    Matrix matrix = Matrix::Identity();
    matrix = Matrix::Translate(Vector(-bb.Min.X(), -bb.Min.Y(), -bb.Min.Z())).Mul(matrix);
    matrix = Matrix::Scale(Vector(scale, scale, scale)).Mul(matrix);
    matrix = Matrix::Translate(Add(box.Min, Mul(extra, anchor))).Mul(matrix);
    Transform(matrix);
}

void Mesh::SetMaterial(const Material& material) { for (Triangle& t : Triangles) t.Mat = material; }

// ---- OBJ -------------------------------------------------------------------------------------------------------------
namespace {
std::map<std::string, Material>& mat_list() { static std::map<std::string, Material> m; return m; }  // OBJ.cs:9 (static: survives loads)

std::vector<std::string> split_spaces(const std::string& line) {  // Split(' ') + RemoveAll(empty): only U+0020 separates
    std::vector<std::string> words;
    size_t i = 0;
    while (i <= line.size()) {
        size_t j = line.find(' ', i);
        if (j == std::string::npos) j = line.size();
        if (j > i) words.push_back(line.substr(i, j - i));
        i = j + 1;
    }
    return words;
}
bool read_line(std::istream& in, std::string& line) {  // StreamReader.ReadLine: \n, \r or \r\n end a line
    if (!std::getline(in, line)) return false;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    return true;
}
float parse_float(const std::string& s) {
    char* end = nullptr;
    const float v = std::strtof(s.c_str(), &end);
    if (end == s.c_str()) throw std::runtime_error("OBJ: bad number \"" + s + "\"");
    return v;
}
int parse_int(const std::string& s) {
    char* end = nullptr;
    const long v = std::strtol(s.c_str(), &end, 10);
    if (end == s.c_str()) throw std::runtime_error("OBJ: bad index \"" + s + "\"");
    return (int)v;
}
// arg.Split({"//", "/"}, RemoveEmptyEntries): "7//3" gives {"7", "3"} - the normal index lands in the texture slot (kept)
std::vector<std::string> split_face_vertex(const std::string& arg) {
    std::vector<std::string> parts;
    std::string cur;
    for (char c : arg) {
        if (c == '/') { if (!cur.empty()) parts.push_back(cur); cur.clear(); }
        else cur.push_back(c);
    }
    if (!cur.empty()) parts.push_back(cur);
    return parts;
}
void load_mtl(const std::string& path, const Material& parent) {
    // OBJ.cs:165-218.  Material is a struct: `matList[name] = material` stores a COPY of the parent taken at `newmtl`, and
    // the Kd / Ke / map_* lines that follow edit a local the dictionary never sees - a material library only registers names.
    std::ifstream in(path);
    if (!in) return;  // "MTL file not found": recolours a by-value copy of the parent, i.e. nothing
    std::string line;
    while (read_line(in, line)) {
        const std::vector<std::string> w = split_spaces(line);
        if (w.size() >= 2 && w[0] == "newmtl") mat_list()[w[1]] = parent;
    }
}
}  // namespace

std::shared_ptr<Mesh> OBJ::Load(const std::string& path, const Material& parent) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Unable to open \"" + path + "\", does not exist.");
    std::vector<Vector> vs, vts, vns;
    vns.push_back(Vector(0, 0, 0));  // OBJ.cs:16: a dummy first normal, so `vn` index k names the (k-1)-th normal of the file
    std::vector<Triangle> triangles;
    Material material = parent;
    std::string line;
    while (read_line(in, line)) {
        std::transform(line.begin(), line.end(), line.begin(), [](unsigned char c) { return (char)std::tolower(c); });
        std::vector<std::string> words = split_spaces(line);
        if (words.empty()) continue;
        const std::string type = words[0];
        words.erase(words.begin());
        if (type == "mtllib") {
            if (!words.empty()) load_mtl(words[0], parent);  // relative to the working directory, like the reference
        } else if (type == "usemtl") {
            auto it = words.empty() ? mat_list().end() : mat_list().find(words[0]);
            if (it != mat_list().end()) material = it->second;
        } else if (type == "v") {
            if (words.size() < 3) throw std::runtime_error("OBJ: short v line");
            vs.push_back(Vector(parse_float(words[0]), parse_float(words[1]), parse_float(words[2])));
        } else if (type == "vt") {
            if (words.size() < 2) throw std::runtime_error("OBJ: short vt line");
            vts.push_back(Vector(parse_float(words[0]), parse_float(words[1]), 0));
        } else if (type == "vn") {
            if (words.size() < 3) throw std::runtime_error("OBJ: short vn line");
            vns.push_back(Vector(parse_float(words[0]), parse_float(words[1]), parse_float(words[2])));
        } else if (type == "f") {
            const size_t n = words.size();
            std::vector<int> fvs(n, 0), fvts(n, 0), fvns(n, 0);
            for (size_t c = 0; c < n; c++) {
                const std::vector<std::string> vertex = split_face_vertex(words[c]);
                if (vertex.size() > 0) fvs[c] = parse_int(vertex[0]) - 1;
                if (vertex.size() > 1) fvts[c] = parse_int(vertex[1]) - 1;
                if (vertex.size() > 2) fvns[c] = parse_int(vertex[2]) - 1;
            }
            auto at = [&](const std::vector<Vector>& list, int i) -> const Vector& {
                if (i < 0 || (size_t)i >= list.size()) throw std::runtime_error("OBJ: index out of range in \"" + line + "\"");
                return list[(size_t)i];
            };
            for (size_t i = 1; i + 1 < n; i++) {  // fan triangulation (0, i, i + 1)
                Triangle t;
                t.Mat = material;
                if (!vs.empty()) { t.V1 = at(vs, fvs[0]); t.V2 = at(vs, fvs[i]); t.V3 = at(vs, fvs[i + 1]); }
                if (!vts.empty()) { t.T1 = at(vts, fvts[0]); t.T2 = at(vts, fvts[i]); t.T3 = at(vts, fvts[i + 1]); }
                t.N1 = at(vns, fvns[0]); t.N2 = at(vns, fvns[i]); t.N3 = at(vns, fvns[i + 1]);
                t.FixNormals();
                triangles.push_back(t);
            }
        }
    }
    return Mesh::NewMesh(std::move(triangles));
}

// ---- STL -------------------------------------------------------------------------------------------------------------
namespace {
bool starts_with_ci(const std::string& s, const char* prefix) {
    size_t i = 0;
    for (; prefix[i]; i++) if (i >= s.size() || std::tolower((unsigned char)s[i]) != std::tolower((unsigned char)prefix[i])) return false;
    return true;
}
std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
bool is_binary_stl(const std::string& bytes) {  // STL.cs:52-71
    if (bytes.size() < 84) return false;
    std::string header = bytes.substr(0, 80);
    size_t a = 0;
    while (a < header.size() && std::isspace((unsigned char)header[a])) a++;
    if (starts_with_ci(header.substr(a), "solid")) return bytes.substr(0, 256).find("facet") == std::string::npos;
    return true;
}
// the three numbers after "facet normal" / "vertex" (STL.cs:73, 129-143): double.Parse, then Vector's float storage
bool parse_vector(const std::string& line, Vector& out) {
    std::string rest;
    if (starts_with_ci(line, "facet normal")) rest = line.substr(12);
    else if (starts_with_ci(line, "vertex")) rest = line.substr(6);
    else return false;
    std::istringstream ss(rest);
    std::string tok[3];
    if (!(ss >> tok[0] >> tok[1] >> tok[2])) return false;
    double v[3];
    for (int k = 0; k < 3; k++) {
        char* end = nullptr;
        v[k] = std::strtod(tok[k].c_str(), &end);
        if (end == tok[k].c_str() || *end) return false;
    }
    out = Vector(v[0], v[1], v[2]);
    return true;
}
float le_float(const unsigned char* p) { uint32_t u = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); float f; std::memcpy(&f, &u, 4); return f; }
}  // namespace

std::shared_ptr<Mesh> STL::Load(const std::string& path, const Material& material) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("Unable to open \"" + path + "\"");
    const std::string bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::vector<Triangle> triangles;
    auto add = [&](const Vector& a, const Vector& b, const Vector& c) {
        Triangle t;
        t.V1 = a; t.V2 = b; t.V3 = c; t.Mat = material;  // the file's facet normal is read and dropped (STL.cs:94, 190-194)
        t.FixNormals();
        triangles.push_back(t);
    };
    if (is_binary_stl(bytes)) {  // STL.cs:146-222: 80-byte header, int32 count, 50-byte facets; a short file keeps what was read
        const unsigned char* p = reinterpret_cast<const unsigned char*>(bytes.data());
        const uint32_t count = (uint32_t)p[80] | ((uint32_t)p[81] << 8) | ((uint32_t)p[82] << 16) | ((uint32_t)p[83] << 24);
        for (uint32_t i = 0; i < count; i++) {
            const size_t off = 84 + (size_t)i * 50;
            if (off + 50 > bytes.size()) break;
            const unsigned char* f = p + off;
            add(Vector(le_float(f + 12), le_float(f + 16), le_float(f + 20)), Vector(le_float(f + 24), le_float(f + 28), le_float(f + 32)),
                Vector(le_float(f + 36), le_float(f + 40), le_float(f + 44)));
        }
        return Mesh::NewMesh(std::move(triangles));
    }
    // STL.cs:73-127: any failure while reading text gives an empty mesh
    std::istringstream text(bytes);
    std::string line;
    if (!read_line(text, line) || line.find("solid") == std::string::npos) return Mesh::NewMesh({});
    std::vector<Vector> vertices;
    while (read_line(text, line)) {
        line = trim(line);
        Vector v;
        if (starts_with_ci(line, "facet normal")) { if (!parse_vector(line, v)) return Mesh::NewMesh({}); }
        else if (starts_with_ci(line, "vertex")) { if (!parse_vector(line, v)) return Mesh::NewMesh({}); vertices.push_back(v); }
        else if (starts_with_ci(line, "endfacet")) {
            if (vertices.size() >= 3) { add(vertices[0], vertices[1], vertices[2]); vertices.clear(); }
        } else if (starts_with_ci(line, "endsolid")) break;
    }
    return Mesh::NewMesh(std::move(triangles));
}

}  // namespace ptsharp

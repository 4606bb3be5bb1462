// scene_loader.cpp -- see scene_loader.h.  Everything is double precision with the reference's operation order
// (compiled with -ffp-contract=off) so that the flattened scene is bit-identical to what the reference hands to
// its path tracer; tests/test_host.py checks that against dumps of the compiled reference.
#include "scene_loader.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <set>
#include <sstream>

#include "xml_mini.h"

namespace dsrt_host {
namespace {

const double kPI = 3.14159265358979323;     // CMU462/misc.h:10
const float kEPS_F = 0.00001f;              // misc.h:13
inline double radians(double deg) { return deg * (kPI / 180); }     // misc.h:24-26
inline double degrees(double rad) { return rad * (180 / kPI); }     // misc.h:31-33

struct V3 { double x = 0, y = 0, z = 0; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double c, V3 v) { return {c * v.x, c * v.y, c * v.z}; }
inline V3 cross(V3 u, V3 v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
inline double norm(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V3 unit(V3 a) { double rn = 1. / std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); return {rn * a.x, rn * a.y, rn * a.z}; }   // vector3D.h:121
inline V3 normalized(V3 a) { double c = 1. / norm(a); return {a.x * c, a.y * c, a.z * c}; }                                      // vector3D.h:129
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }

struct V4 { double x = 0, y = 0, z = 0, w = 0; };
// column-major 4x4 like CMU462::Matrix4x4 (entries[j] = column j; the default constructor leaves zeros)
struct M4 {
  double c[4][4];   // c[col][row]
  M4() { for (auto& col : c) for (double& v : col) v = 0; }
  static M4 identity() { M4 m; for (int i = 0; i < 4; i++) m.c[i][i] = 1.; return m; }
  double& at(int i, int j) { return c[j][i]; }
  double at(int i, int j) const { return c[j][i]; }
  M4 operator*(const M4& B) const {                          // matrix4x4.cpp:155-171
    M4 C;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
      double s = 0.;
      for (int k = 0; k < 4; k++) s += at(i, k) * B.at(k, j);
      C.at(i, j) = s;
    }
    return C;
  }
  V4 operator*(const V4& x) const {                          // matrix4x4.cpp:174-179: x0*col0 + x1*col1 + x2*col2 + x3*col3
    V4 r;
    double* o = &r.x; const double xs[4] = {x.x, x.y, x.z, x.w};
    for (int i = 0; i < 4; i++) o[i] = ((xs[0] * c[0][i] + xs[1] * c[1][i]) + xs[2] * c[2][i]) + xs[3] * c[3][i];
    return r;
  }
};
inline V4 v4(V3 v, double w) { return {v.x, v.y, v.z, w}; }
inline V3 to3D(V4 v) { return {v.x, v.y, v.z}; }
inline V3 projectTo3D(V4 v) { double iw = 1.0 / v.w; return {v.x * iw, v.y * iw, v.z * iw}; }   // vector4D.cpp:14-17

struct Bsdf { int type = 0; float a[3] = {0, 0, 0}, b[3] = {0, 0, 0}; float ior = 0; };
enum InstanceType { CAMERA, LIGHT, SPHERE, POLYMESH, NONE };
enum LightType { L_NONE, L_AMBIENT, L_DIRECTIONAL, L_AREA, L_POINT, L_SPOT };
struct Instance {
  InstanceType type = NONE;
  // camera
  float hFov = 0, vFov = 0, nClip = 0, fClip = 0; V3 view_dir, up_dir;
  // light
  LightType light_type = L_NONE; float spectrum[3] = {1, 1, 1}; V3 position{0, 0, 0}, direction{0, 0, -1}, up{0, 1, 0};
  // sphere
  float radius = 0;
  // polymesh
  std::vector<V3> vertices; std::vector<std::vector<size_t>> polygons;
  bool has_material = false; Bsdf material;
};
struct Node { M4 transform = M4::identity(); Instance inst; bool has_instance = false; };

struct Parser {
  std::map<std::string, XmlElement*> sources;
  V3 up;
  M4 transform;              // static Matrix4x4 ColladaParser::transform -> zero until <asset> is read
  std::vector<Node> nodes;
  std::string err;

  bool fail(const std::string& m) { if (err.empty()) err = m; return false; }

  void uri_load(XmlElement* xml) {                           // collada.cpp:55-68
    if (xml->Attribute("id")) sources[xml->Attribute("id")] = xml;
    for (XmlElement* c = xml->FirstChildElement(); c; c = c->NextSiblingElement()) uri_load(c);
  }
  XmlElement* uri_find(const std::string& id) { auto it = sources.find(id); return it == sources.end() ? nullptr : it->second; }
  XmlElement* get_element(XmlElement* xml, const std::string& query) {   // collada.cpp:78-99
    std::stringstream ss(query);
    XmlElement* e = xml; std::string token;
    while (e && std::getline(ss, token, '/')) e = e->FirstChildElement(token.c_str());
    if (e) { const char* url = e->Attribute("url"); if (url) e = uri_find(std::string(url + 1)); }
    return e;
  }
  XmlElement* get_technique_common(XmlElement* xml) {        // collada.cpp:102-116
    if (XmlElement* cp = xml->FirstChildElement("profile_COMMON")) {
      for (XmlElement* t = cp->FirstChildElement("technique"); t; t = t->NextSiblingElement("technique")) {
        const char* sid = t->Attribute("sid");
        if (sid && std::string(sid) == "common") return t;
      }
    }
    return xml->FirstChildElement("technique_common");
  }
  XmlElement* get_technique_cmu462(XmlElement* xml) {        // collada.cpp:119-130
    for (XmlElement* t = get_element(xml, "extra/technique"); t; t = t->NextSiblingElement("technique")) {
      const char* p = t->Attribute("profile");
      if (p && std::string(p) == "CMU462") return t;
    }
    return nullptr;
  }
  static void spectrum_from_string(const char* s, float out[3]) {   // collada.cpp:27-39 (operator>> into float)
    out[0] = out[1] = out[2] = 0;
    if (!s) return;
    char* e = nullptr;
    for (int i = 0; i < 3; i++) { out[i] = strtof(s, &e); if (e == s) { out[i] = 0; break; } s = e; }
  }
  static const char* text(XmlElement* e) { return e ? e->GetText() : nullptr; }

  bool parse_material(XmlElement* xml, Bsdf& m) {            // collada.cpp:852-936
    XmlElement* e_effect = get_element(xml, "instance_effect");
    if (!e_effect) return fail("no target effects found for material");
    XmlElement* tc = get_technique_common(e_effect);
    XmlElement* t462 = get_technique_cmu462(e_effect);
    m = Bsdf(); m.type = 0; m.a[0] = m.a[1] = m.a[2] = .5f;
    if (t462) {
      for (XmlElement* b = t462->FirstChildElement(); b; b = b->NextSiblingElement()) {
        std::string type = b->Name();
        if (type == "emission") { m = Bsdf(); m.type = 4; spectrum_from_string(text(get_element(b, "radiance")), m.a); }
        else if (type == "mirror") { m = Bsdf(); m.type = 1; spectrum_from_string(text(get_element(b, "reflectance")), m.a); }
        else if (type == "refraction") {
          m = Bsdf(); m.type = 2; spectrum_from_string(text(get_element(b, "transmittance")), m.b);
          const char* io = text(get_element(b, "ior")); m.ior = io ? (float)atof(io) : 0.f;
        } else if (type == "glass") {
          m = Bsdf(); m.type = 3; spectrum_from_string(text(get_element(b, "transmittance")), m.b);
          spectrum_from_string(text(get_element(b, "reflectance")), m.a);
          const char* io = text(get_element(b, "ior")); m.ior = io ? (float)atof(io) : 0.f;
        }
      }
    } else if (tc) {
      XmlElement* d = get_element(tc, "phong/diffuse/color");
      if (d) { m.type = 0; spectrum_from_string(text(d), m.a); }
    }
    return true;
  }

  bool bind_material(XmlElement* node_xml, Instance& inst) { // collada.cpp:369-389 / 399-418
    XmlElement* im = get_element(node_xml, "instance_geometry/bind_material/technique_common/instance_material");
    if (!im) return true;
    const char* target = im->Attribute("target");
    if (!target) return fail("no target material in instance");
    XmlElement* e_mat = uri_find(std::string(target + 1));
    if (!e_mat) return fail(std::string("invalid target material id: ") + (target + 1));
    inst.has_material = true;
    return parse_material(e_mat, inst.material);
  }

  bool parse_camera(XmlElement* xml, Instance& c) {          // collada.cpp:432-473
    c.type = CAMERA; c.up_dir = up; c.view_dir = {0, 0, -1};
    XmlElement* p = get_element(xml, "optics/technique_common/perspective");
    if (!p) return fail("no perspective defined in camera");
    XmlElement *xf = p->FirstChildElement("xfov"), *yf = p->FirstChildElement("yfov"), *zn = p->FirstChildElement("znear"), *zf = p->FirstChildElement("zfar");
    c.hFov = xf && xf->GetText() ? (float)atof(xf->GetText()) : 50.0f;
    c.vFov = yf && yf->GetText() ? (float)atof(yf->GetText()) : 35.0f;
    c.nClip = zn && zn->GetText() ? (float)atof(zn->GetText()) : 0.001f;
    c.fClip = zf && zf->GetText() ? (float)atof(zf->GetText()) : 1000.0f;
    if (!yf) {
      XmlElement* ar = get_element(p, "aspect_ratio");
      if (!ar || !ar->GetText()) return fail("incomplete perspective definition in camera");
      float aspect_ratio = (float)atof(ar->GetText());
      c.vFov = (float)(2 * degrees(std::atan(std::tan(radians(0.5 * c.hFov)) / aspect_ratio)));
    }
    return true;
  }

  bool parse_light(XmlElement* xml, Instance& l) {           // collada.cpp:475-576
    l.type = LIGHT;
    XmlElement* tc = get_technique_common(xml);
    XmlElement* t462 = get_technique_cmu462(xml);
    XmlElement* technique = t462 ? t462 : tc;
    if (!technique) return fail("no supported profile defined in light");
    XmlElement* e = technique->FirstChildElement();
    if (!e) return true;
    std::string type = e->Name();
    XmlElement* col = get_element(e, "color");
    if (type == "ambient") l.light_type = L_AMBIENT;
    else if (type == "directional") l.light_type = L_DIRECTIONAL;
    else if (type == "area") l.light_type = L_AREA;
    else if (type == "point") {
      l.light_type = L_POINT;
      if (!(col && get_element(e, "constant_attenuation") && get_element(e, "linear_attenuation") && get_element(e, "quadratic_attenuation")))
        return fail("incomplete definition of point light");
    } else if (type == "spot") {
      l.light_type = L_SPOT;
    } else return fail("light type " + type + " is not supported");
    if (!col) return fail("no color definition in light");
    spectrum_from_string(col->GetText(), l.spectrum);
    return true;
  }

  bool parse_sphere(XmlElement* xml, Instance& s) {          // collada.cpp:578-601
    s.type = SPHERE;
    XmlElement* t = get_technique_cmu462(xml);
    if (!t) return fail("no 462 profile technique in sphere geometry");
    XmlElement* r = get_element(t, "sphere/radius");
    if (!r || !r->GetText()) return fail("invalid sphere definition in geometry");
    s.radius = (float)atof(r->GetText());
    return true;
  }

  // a `count` attribute is untrusted: it must be non-negative and cannot exceed what the element's text could hold
  // (one character per value plus a separator); anything else is a corrupt file, not an allocation request
  static bool plausible_count(long long count, const char* text) {
    return count >= 0 && (unsigned long long)count <= (text ? (unsigned long long)strlen(text) / 2 + 1 : 0ull);
  }
  static bool read_floats(const char* s, size_t n, std::vector<float>& out) {
    out.clear(); out.reserve(n);
    if (!s) return n == 0;
    char* e = nullptr;
    bool failed = false;
    for (size_t i = 0; i < n; i++) {
      float f = 0.f;              // a failed stream extraction stores 0 (C++11), and every later one fails too
      if (!failed) { f = strtof(s, &e); if (e == s) { failed = true; f = 0.f; } else s = e; }
      out.push_back(f);
    }
    return true;
  }
  static void read_sizes(const char* s, size_t n, std::vector<size_t>& out) {
    out.clear(); out.reserve(n);
    char* e = nullptr; bool failed = (s == nullptr);
    for (size_t i = 0; i < n; i++) {
      size_t v = 0;
      if (!failed) { unsigned long long u = strtoull(s, &e, 10); if (e != s) { v = (size_t)u; s = e; } else failed = true; }
      out.push_back(v);
    }
  }

  bool parse_polymesh(XmlElement* xml, Instance& pm) {       // collada.cpp:604-850
    pm.type = POLYMESH;
    XmlElement* e_mesh = xml->FirstChildElement("mesh");
    if (!e_mesh) return fail("no mesh data defined in geometry");
    std::map<std::string, std::vector<float>> arr;
    for (XmlElement* s = e_mesh->FirstChildElement("source"); s; s = s->NextSiblingElement("source")) {
      const char* sid = s->Attribute("id");
      XmlElement* fa = s->FirstChildElement("float_array");
      if (sid && fa) {
        if (!plausible_count(fa->IntAttribute("count"), fa->GetText())) return fail("implausible float_array count in geometry");
        std::vector<float> f; read_floats(fa->GetText(), (size_t)fa->IntAttribute("count"), f); arr[sid] = f;
      }
    }
    XmlElement* e_vertices = e_mesh->FirstChildElement("vertices");
    if (!e_vertices || !e_vertices->Attribute("id")) return fail("no vertices defined in geometry");
    std::string vertices_id = e_vertices->Attribute("id");
    std::vector<V3> vertices;
    for (XmlElement* in = e_vertices->FirstChildElement("input"); in; in = in->NextSiblingElement("input")) {
      const char* sem = in->Attribute("semantic"); const char* src = in->Attribute("source");
      if (sem && src && std::string(sem) == "POSITION") {
        auto it = arr.find(std::string(src + 1));
        if (it == arr.end()) return fail(std::string("undefined input source: ") + (src + 1));
        const std::vector<float>& f = it->second;
        for (size_t i = 0; i + 2 < f.size(); i += 3) vertices.push_back({f[i], f[i + 1], f[i + 2]});
      }
    }
    XmlElement* pl = e_mesh->FirstChildElement("polylist");
    if (!pl) return true;
    bool hv = false, hn = false, ht = false; size_t vo = 0;
    for (XmlElement* in = pl->FirstChildElement("input"); in; in = in->NextSiblingElement("input")) {
      const char* sem = in->Attribute("semantic"); const char* src = in->Attribute("source");
      if (!sem || !src) return fail("polylist input without semantic/source");
      std::string semantic = sem, source = src + 1;
      if (in->IntAttribute("offset") < 0 || in->IntAttribute("offset") > 16) return fail("implausible polylist input offset");
      size_t offset = (size_t)in->IntAttribute("offset");
      if (semantic == "VERTEX") {
        hv = true; vo = offset;
        if (source != vertices_id) return fail("undefined source for VERTEX semantic: " + source);
        pm.vertices = vertices;
      }
      if (semantic == "NORMAL") { hn = true; if (arr.find(source) == arr.end()) return fail("undefined source for NORMAL semantic: " + source); }
      if (semantic == "TEXCOORD") { ht = true; if (arr.find(source) == arr.end()) return fail("undefined source for TEXCOORD semantic: " + source); }
    }
    const size_t stride = (hv ? 1 : 0) + (hn ? 1 : 0) + (ht ? 1 : 0);
    XmlElement* e_vcount = pl->FirstChildElement("vcount");
    if (!e_vcount) return fail("polygon sizes undefined in geometry");
    if (!plausible_count(pl->IntAttribute("count"), e_vcount->GetText())) return fail("implausible polylist count in geometry");
    const size_t num_polygons = (size_t)pl->IntAttribute("count");
    std::vector<size_t> sizes; read_sizes(e_vcount->GetText(), num_polygons, sizes);
    XmlElement* e_p = pl->FirstChildElement("p");
    if (!e_p) return fail("no index array defined in geometry");
    const size_t index_text = e_p->GetText() ? strlen(e_p->GetText()) : 0;
    size_t num_indices = 0;
    for (size_t s : sizes) {
      if (s > index_text) return fail("implausible polygon size in geometry");
      num_indices += s * stride;
      if (num_indices > index_text + 1) return fail("index array too short in geometry");
    }
    std::vector<size_t> indices; read_sizes(e_p->GetText(), num_indices, indices);
    pm.polygons.assign(num_polygons, {});
    if (hv) {
      size_t k = 0;
      for (size_t i = 0; i < num_polygons; i++) for (size_t j = 0; j < sizes[i]; j++) {
        size_t at = k * stride + vo;
        if (at >= indices.size()) return fail("index array too short in geometry");
        pm.polygons[i].push_back(indices[at]); k++;
      }
    }
    return true;
  }

  bool parse_node(XmlElement* xml) {                         // collada.cpp:234-430
    Node node;
    for (XmlElement* e = xml->FirstChildElement(); e; e = e->NextSiblingElement()) {
      std::string name = e->Name();
      if (name == "matrix") {
        // 16 stream extractions into a zero-initialised Matrix4x4; a short list leaves zeros (CBgems.dae's camera
        // node has 15 numbers), collada.cpp:253-266
        std::vector<double> v(16, 0.0); const char* s = e->GetText(); char* end = nullptr;
        for (int i = 0; i < 16 && s; i++) { double d = strtod(s, &end); if (end == s) break; v[i] = d; s = end; }
        M4 mat; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) mat.at(i, j) = v[i * 4 + j];
        node.transform = mat; break;
      }
      if (name == "rotate" || name == "translate" || name == "scale") {
        // the reference fills a ZERO-initialised Matrix4x4 (collada.cpp:272-322); restated as is
        std::vector<double> v; const char* s = e->GetText(); char* end = nullptr;
        while (s) { double d = strtod(s, &end); if (end == s) break; v.push_back(d); s = end; }
        M4 m;
        auto get = [&](size_t i) { return i < v.size() ? v[i] : 0.0; };
        if (name == "rotate") {
          const char* sid = e->Attribute("sid");
          if (!sid || !*sid) return fail("<rotate> without sid");
          switch (std::string(sid).back()) {
            case 'X': m.at(1, 1) = get(0); m.at(1, 2) = get(1); m.at(2, 1) = get(2); m.at(2, 2) = get(3); break;
            case 'Y': m.at(0, 0) = get(0); m.at(2, 0) = get(1); m.at(0, 2) = get(2); m.at(2, 2) = get(3); break;
            case 'Z': m.at(0, 0) = get(0); m.at(0, 1) = get(1); m.at(1, 0) = get(2); m.at(1, 1) = get(3); break;
            default: break;
          }
        } else if (name == "translate") { m.at(0, 3) = get(0); m.at(1, 3) = get(1); m.at(2, 3) = get(2); }
        else { m.at(0, 0) = get(0); m.at(1, 1) = get(2); }       // m(1,1) is written twice, collada.cpp:319
        node.transform = m * node.transform;
      }
    }
    const M4 save = transform;
    node.transform = transform * node.transform;
    transform = node.transform;
    for (XmlElement* c = get_element(xml, "node"); c; c = c->NextSiblingElement("node")) if (!parse_node(c)) return false;
    transform = save;
    XmlElement* e_camera = get_element(xml, "instance_camera");
    XmlElement* e_light = get_element(xml, "instance_light");
    XmlElement* e_geometry = get_element(xml, "instance_geometry");
    if (e_camera) { node.has_instance = true; if (!parse_camera(e_camera, node.inst)) return false; }
    else if (e_light) { node.has_instance = true; if (!parse_light(e_light, node.inst)) return false; }
    else if (e_geometry) {
      if (get_element(e_geometry, "mesh")) {
        node.has_instance = true;
        if (!parse_polymesh(e_geometry, node.inst) || !bind_material(xml, node.inst)) return false;
      } else if (get_element(e_geometry, "extra")) {
        node.has_instance = true;
        if (!parse_sphere(e_geometry, node.inst) || !bind_material(xml, node.inst)) return false;
      }
    }
    nodes.push_back(node);
    return true;
  }

  bool load(const std::string& path) {                       // collada.cpp:131-225
    std::ifstream in(path, std::ios::binary);
    if (!in.is_open()) return fail("cannot open " + path);
    std::stringstream buf; buf << in.rdbuf();
    const std::string content = buf.str();
    XmlDocument doc;
    if (!doc.Parse(content)) return fail("XML error: " + doc.error());
    XmlElement* root = doc.FirstChildElement("COLLADA");
    if (!root) return fail("not a COLLADA file");
    uri_load(root);
    if (XmlElement* asset = get_element(root, "asset")) {
      XmlElement* up_axis = get_element(asset, "up_axis");
      if (!up_axis || !up_axis->GetText()) return fail("no up direction defined in COLLADA file");
      std::string up_dir = up_axis->GetText();
      transform = M4::identity();
      if (up_dir == "X_UP") {
        transform.at(0, 0) = 0; transform.at(0, 1) = 1; transform.at(1, 0) = 1; transform.at(1, 1) = 0; transform.at(2, 2) = -1;
        up = {1, 0, 0};
      } else if (up_dir == "Z_UP") {
        transform.at(1, 1) = 0; transform.at(1, 2) = 1; transform.at(2, 1) = 1; transform.at(2, 2) = 0; transform.at(0, 0) = -1;
        up = {0, 0, 1};
      } else if (up_dir == "Y_UP") up = {0, 1, 0};
      else return fail("invalid up direction in COLLADA file");
    }
    XmlElement* e_scene = get_element(root, "scene/instance_visual_scene");
    if (!e_scene) return fail("no scene description found in file");
    for (XmlElement* n = get_element(e_scene, "node"); n; n = n->NextSiblingElement("node")) if (!parse_node(n)) return false;
    return true;
  }
};

// ---- HalfedgeMesh::build + Vertex::computeNormal + StaticScene::Mesh, restated over index arrays -------------------
struct HalfedgeMesh {
  std::vector<int> next, twin, vert, face;       // per halfedge (creation order); face >= n_faces means boundary loop
  std::vector<int> v_he;                         // per vertex (first-encounter order)
  std::vector<V3> v_pos, v_nrm;
  std::vector<int> f_he;                         // per face: its representative halfedge (the LAST one created)
  int n_faces = 0;

  int new_he() { next.push_back(-1); twin.push_back(-1); vert.push_back(-1); face.push_back(-1); return (int)next.size() - 1; }
  bool he_boundary(int h) const { return face[h] >= n_faces; }
  bool v_boundary(int v) const {                 // Vertex::isBoundary, halfEdgeMesh.h:530-545
    int h = v_he[v];
    do { if (he_boundary(h)) return true; h = next[twin[h]]; } while (h != v_he[v]);
    return false;
  }

  bool build(const std::vector<std::vector<size_t>>& polygons, const std::vector<V3>& positions, std::string& err) {
    std::map<size_t, int> index_to_vertex;
    std::vector<size_t> degree;
    for (const auto& p : polygons) {
      if (p.size() < 3) { err = "each polygon must have at least three vertices"; return false; }
      std::set<size_t> distinct;
      for (size_t i : p) {
        distinct.insert(i);
        auto it = index_to_vertex.find(i);
        if (it == index_to_vertex.end()) { index_to_vertex[i] = (int)v_he.size(); v_he.push_back(-1); degree.push_back(1); }
        else degree[it->second]++;
      }
      if (distinct.size() < p.size()) { err = "one of the input polygons does not have distinct vertices"; return false; }
    }
    n_faces = (int)polygons.size();
    f_he.assign(n_faces, -1);
    std::map<std::pair<size_t, size_t>, int> pair_to_he;
    for (int f = 0; f < n_faces; f++) {
      const auto& p = polygons[f];
      const size_t deg = p.size();
      std::vector<int> fh;
      for (size_t i = 0; i < deg; i++) {
        const size_t a = p[i], b = p[(i + 1) % deg];
        if (pair_to_he.count({a, b})) { err = "found multiple oriented edges with the same indices (non-manifold or inconsistently oriented mesh)"; return false; }
        const int hab = new_he();
        pair_to_he[{a, b}] = hab;
        face[hab] = f; f_he[f] = hab;
        vert[hab] = index_to_vertex[a]; v_he[vert[hab]] = hab;
        fh.push_back(hab);
        auto it = pair_to_he.find({b, a});
        if (it != pair_to_he.end()) { twin[hab] = it->second; twin[it->second] = hab; }
      }
      for (size_t i = 0; i < deg; i++) next[fh[i]] = fh[(i + 1) % deg];
    }
    // boundary vertices point at a twinless halfedge (halfEdgeMesh.cpp:225-239)
    for (size_t v = 0; v < v_he.size(); v++) {
      int h = v_he[v];
      do {
        if (twin[h] < 0) { v_he[v] = h; break; }
        h = next[twin[h]];
      } while (h != v_he[v]);
    }
    // one virtual face per boundary loop (halfEdgeMesh.cpp:242-313); the loop also runs over the halfedges it appends
    int n_boundaries = 0;
    for (int h = 0; h < (int)next.size(); h++) {
      if (twin[h] >= 0) continue;
      const int b = n_faces + n_boundaries++;
      std::vector<int> bh;
      int i = h;
      do {
        const int t = new_he();
        bh.push_back(t);
        twin[i] = t; twin[t] = i; face[t] = b; vert[t] = vert[next[i]];
        i = next[i];
        while (i != h && twin[i] >= 0) {
          i = twin[i];
          if (next[i] < 0) { err = "non-manifold boundary"; return false; }
          i = next[i];
        }
      } while (i != h);
      const size_t deg = bh.size();
      for (size_t p = 0; p < deg; p++) next[bh[p]] = bh[(p + deg - 1) % deg];
    }
    for (size_t v = 0; v < v_he.size(); v++) v_he[v] = next[twin[v_he[v]]];          // halfEdgeMesh.cpp:321-323
    for (size_t v = 0; v < v_he.size(); v++) {                                       // manifold check, :326-352
      size_t count = 0; int h = v_he[v];
      do { if (!he_boundary(h)) count++; h = next[twin[h]]; } while (h != v_he[v]);
      if (count != degree[v]) { err = "at least one of the vertices is nonmanifold"; return false; }
    }
    if (positions.size() != v_he.size()) { err = "number of vertex positions is different from the number of distinct vertices"; return false; }
    v_pos.assign(v_he.size(), V3());
    { size_t i = 0; for (const auto& kv : index_to_vertex) v_pos[kv.second] = positions[i++]; }   // sorted-index order, :379-390
    v_nrm.assign(v_he.size(), V3());
    for (size_t v = 0; v < v_he.size(); v++) {                                       // Vertex::computeNormal, halfEdgeMesh.h:492-515
      V3 n{0, 0, 0}; const V3 pi = v_pos[v];
      int h = v_he[v];
      const bool boundary = v_boundary((int)v);
      do {
        const V3 pj = v_pos[vert[next[h]]], pk = v_pos[vert[next[next[h]]]];
        n = n + cross(pj - pi, pk - pi);
        h = boundary ? twin[next[h]] : next[twin[h]];
      } while (h != v_he[v]);
      v_nrm[v] = normalized(n);
    }
    return true;
  }
};

void box_grow(double lo[3], double hi[3], V3 p) {
  lo[0] = std::min(lo[0], p.x); lo[1] = std::min(lo[1], p.y); lo[2] = std::min(lo[2], p.z);
  hi[0] = std::max(hi[0], p.x); hi[1] = std::max(hi[1], p.y); hi[2] = std::max(hi[2], p.z);
}

}  // namespace

bool g_direct_triangles_storage = false;

// ---- Camera --------------------------------------------------------------------------------------------------------
void HostCamera::configure(double hFov_, double vFov_, double nClip_, double fClip_, size_t w, size_t h) {
  screenW = w; screenH = h; nClip = nClip_; fClip = fClip_; hFov = hFov_; vFov = vFov_;
  double ar1 = std::tan(radians(hFov) / 2) / std::tan(radians(vFov) / 2);
  ar = static_cast<double>(screenW) / screenH;
  if (ar1 < ar) hFov = 2 * degrees(std::atan(std::tan(radians(vFov) / 2) * ar));
  else if (ar1 > ar) vFov = 2 * degrees(std::atan(std::tan(radians(hFov) / 2) / ar));
  screenDist = ((double)screenH) / (2.0 * std::tan(radians(vFov) / 2));
}
void HostCamera::place(const double target[3], double phi_, double theta_, double r_, double minR_, double maxR_) {
  double rr = std::min(std::max(r_, minR_), maxR_);
  double ph = (std::sin(phi_) == 0) ? (phi_ + kEPS_F) : phi_;
  for (int k = 0; k < 3; k++) targetPos[k] = target[k];
  phi = ph; theta = theta_; r = rr; minR = minR_; maxR = maxR_;
  compute_position();
}
void HostCamera::compute_position() {
  double sinPhi = std::sin(phi);
  if (sinPhi == 0) { phi += kEPS_F; sinPhi = std::sin(phi); }
  const V3 dirToCamera{r * sinPhi * std::sin(theta), r * std::cos(phi), r * sinPhi * std::cos(theta)};
  pos[0] = targetPos[0] + dirToCamera.x; pos[1] = targetPos[1] + dirToCamera.y; pos[2] = targetPos[2] + dirToCamera.z;
  const V3 upVec{0, sinPhi > 0 ? 1.0 : -1.0, 0};
  V3 sx = normalized(cross(upVec, dirToCamera));
  V3 sy = normalized(cross(dirToCamera, sx));
  V3 sz = unit(dirToCamera);
  c2w[0] = sx.x; c2w[1] = sx.y; c2w[2] = sx.z; c2w[3] = sy.x; c2w[4] = sy.y; c2w[5] = sy.z; c2w[6] = sz.x; c2w[7] = sz.y; c2w[8] = sz.z;
}
bool HostCamera::load_info(const std::string& path, std::string& err) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) { err = "cannot open camera file " + path; return false; }
  double m[9]; int n = 0;
  n += fscanf(f, "%lf %lf %lf", &pos[0], &pos[1], &pos[2]);
  n += fscanf(f, "%lf %lf %lf", &targetPos[0], &targetPos[1], &targetPos[2]);
  n += fscanf(f, "%lf", &phi); n += fscanf(f, "%lf", &theta); n += fscanf(f, "%lf", &minR); n += fscanf(f, "%lf", &maxR);
  n += fscanf(f, "%lf %lf %lf %lf %lf %lf %lf %lf %lf", &m[0], &m[1], &m[2], &m[3], &m[4], &m[5], &m[6], &m[7], &m[8]);
  fclose(f);
  if (n != 19) { err = "malformed camera file " + path; return false; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c2w[j * 3 + i] = m[i * 3 + j];   // file is row-major c2w(i,j)
  return true;
}

// ---- Application::load ---------------------------------------------------------------------------------------------
void set_direct_triangle_fallback(bool on) { g_direct_triangles_storage = on; }

bool load_collada(const std::string& path, size_t width, size_t height, FlatScene& out, HostCamera& camera, std::string& err) {
  const bool g_direct_triangles = g_direct_triangles_storage;
  Parser P;
  if (!P.load(path)) { err = P.err; return false; }
  out = FlatScene();
  camera = HostCamera();
  camera.configure(50, 35, 0.01, 100, 600, 600);                  // Application::init, application.cpp:93-99
  V3 c_pos{0, 0, 0}, c_dir{0, 0, 0};
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  auto add_bsdf = [&](const Instance& in) {
    Bsdf b; if (in.has_material) b = in.material; else { b.type = 0; b.a[0] = b.a[1] = b.a[2] = 0.5f; }   // mesh.cpp:31-35
    out.bsdf_type.push_back(b.type);
    const float row[8] = {b.a[0], b.a[1], b.a[2], b.b[0], b.b[1], b.b[2], b.ior, 0.f};
    out.bsdf_param.insert(out.bsdf_param.end(), row, row + 8);
    return (int32_t)out.bsdf_type.size() - 1;
  };
  for (const Node& node : P.nodes) {
    if (!node.has_instance) continue;
    const Instance& in = node.inst; const M4& T = node.transform;
    switch (in.type) {
      case CAMERA:
        c_pos = to3D(T * v4(c_pos, 1));
        c_dir = unit(to3D(T * v4(in.view_dir, 1)));
        camera.configure(in.hFov, in.vFov, in.nClip, in.fClip, width, height);
        break;
      case LIGHT: {
        double q[28]; for (double& v : q) v = 0;
        for (int k = 0; k < 3; k++) q[k] = in.spectrum[k];
        int type = -1;
        if (in.light_type == L_AMBIENT) {                        // ambient_light.h -> InfiniteHemisphereLight, light.cpp:27-32
          type = 1; const double s2w[9] = {1, 0, 0, 0, 0, -1, 0, 1, 0};
          for (int k = 0; k < 9; k++) q[16 + k] = s2w[k];
        } else if (in.light_type == L_DIRECTIONAL) {             // directional_light.h:16-20, light.cpp:11-15
          type = 0;
          V3 d = normalized(neg(to3D(T * v4(in.direction, 1))));
          V3 dtl = neg(unit(d));
          q[3] = dtl.x; q[4] = dtl.y; q[5] = dtl.z;
        } else if (in.light_type == L_POINT) {                   // point_light.h:18-21
          type = 2; V3 p = to3D(T * v4(in.position, 1)); q[3] = p.x; q[4] = p.y; q[5] = p.z;
        } else if (in.light_type == L_AREA) {                    // area_light.h:16-27, light.cpp:73-77
          type = 3;
          V3 position = to3D(T * v4(in.position, 1));
          V3 direction = normalized(to3D(T * v4(in.direction, 1)) - position);
          V3 dim_y0 = in.up, dim_x0 = cross(in.up, in.direction);
          V3 dim_x = to3D(T * v4(dim_x0, 1)) - position, dim_y = to3D(T * v4(dim_y0, 1)) - position;
          q[3] = position.x; q[4] = position.y; q[5] = position.z; q[6] = direction.x; q[7] = direction.y; q[8] = direction.z;
          q[9] = dim_x.x; q[10] = dim_x.y; q[11] = dim_x.z; q[12] = dim_y.x; q[13] = dim_y.y; q[14] = dim_y.z;
          q[15] = (float)(norm(dim_x) * norm(dim_y));
        }
        // spot lights are empty stubs in the reference (light.cpp:61-69): skipped
        if (type >= 0) { out.light_type.push_back(type); out.light_param.insert(out.light_param.end(), q, q + 28); }
        break;
      }
      case SPHERE: {                                             // application.cpp:342-347, dynamic_scene/sphere.cpp:8-16
        const V3 position = projectTo3D(T * v4({0, 0, 0}, 1));
        const double scale = norm(to3D(T * V4{1, 0, 0, 0}));
        const double r = in.radius * scale;
        const int32_t b = add_bsdf(in);
        out.prim_type.push_back(0); out.prim_bsdf.push_back(b);
        for (int k = 0; k < 9; k++) { out.tri_pos.push_back(0); out.tri_nrm.push_back(0); }
        out.sphere.push_back(position.x); out.sphere.push_back(position.y); out.sphere.push_back(position.z); out.sphere.push_back(r);
        box_grow(lo, hi, {position.x - r, position.y - r, position.z - r}); box_grow(lo, hi, {position.x + r, position.y + r, position.z + r});
        break;
      }
      case POLYMESH: {                                           // dynamic_scene/mesh.cpp:16-35, static_scene/object.cpp:16-57
        std::vector<V3> verts = in.vertices;
        for (V3& v : verts) v = projectTo3D(T * v4(v, 1));
        HalfedgeMesh hm; std::string e;
        if (!hm.build(in.polygons, verts, e)) {
          if (!g_direct_triangles) { err = "error converting polygons to halfedge mesh: " + e; return false; }
          // Direct indexed-triangle import (opt-in, the reference exit(1)s here, halfEdgeMesh.cpp:165-175):
          // polygons are fan-triangulated instead of truncated to their first three vertices (object.cpp:36-41), vertex
          // normals are the normalised sums of cross(pj - pi, pk - pi) over the incident triangles -- the terms of
          // Vertex::computeNormal without the manifold walk.
          std::vector<V3> nrm(verts.size(), V3());
          std::vector<std::array<size_t, 3>> tris;
          for (const auto& poly : in.polygons) {
            if (poly.size() < 3) { err = "each polygon must have at least three vertices"; return false; }
            for (size_t q : poly) if (q >= verts.size()) { err = "polygon index out of range"; return false; }
            for (size_t i = 1; i + 1 < poly.size(); i++) tris.push_back({poly[0], poly[i], poly[i + 1]});
          }
          for (const auto& t : tris)
            for (int k = 0; k < 3; k++) {
              const V3 pi = verts[t[k]], pj = verts[t[(k + 1) % 3]], pk = verts[t[(k + 2) % 3]];
              nrm[t[k]] = nrm[t[k]] + cross(pj - pi, pk - pi);
            }
          int32_t b2 = -1;
          for (const auto& t : tris) {
            if (b2 < 0) b2 = add_bsdf(in);
            out.prim_type.push_back(1); out.prim_bsdf.push_back(b2);
            for (int k = 0; k < 3; k++) { const V3& p = verts[t[k]]; box_grow(lo, hi, p); out.tri_pos.push_back(p.x); out.tri_pos.push_back(p.y); out.tri_pos.push_back(p.z); }
            for (int k = 0; k < 3; k++) { const V3 n = normalized(nrm[t[k]]); out.tri_nrm.push_back(n.x); out.tri_nrm.push_back(n.y); out.tri_nrm.push_back(n.z); }
            for (int k = 0; k < 4; k++) out.sphere.push_back(0);
          }
          break;
        }
        for (const V3& p : hm.v_pos) box_grow(lo, hi, p);
        int32_t b = -1;
        for (int f = 0; f < hm.n_faces; f++) {
          if (b < 0) b = add_bsdf(in);
          const int h = hm.f_he[f];
          const int vi[3] = {hm.vert[h], hm.vert[hm.next[h]], hm.vert[hm.next[hm.next[h]]]};
          out.prim_type.push_back(1); out.prim_bsdf.push_back(b);
          for (int k = 0; k < 3; k++) { const V3& p = hm.v_pos[vi[k]]; out.tri_pos.push_back(p.x); out.tri_pos.push_back(p.y); out.tri_pos.push_back(p.z); }
          for (int k = 0; k < 3; k++) { const V3& n = hm.v_nrm[vi[k]]; out.tri_nrm.push_back(n.x); out.tri_nrm.push_back(n.y); out.tri_nrm.push_back(n.z); }
          for (int k = 0; k < 4; k++) out.sphere.push_back(0);
        }
        break;
      }
      default: break;
    }
  }
  const bool empty = lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2];
  if (!empty) {                                                  // application.cpp:267-291
    const V3 ext{hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    const double target[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
    const double cvd = norm(ext) / 2 * 1.5;
    camera.place(target, std::acos(c_dir.y), std::atan2(c_dir.x, c_dir.z), cvd * 2, cvd / 10.0, cvd * 20.0);
  }
  return true;
}

}  // namespace dsrt_host

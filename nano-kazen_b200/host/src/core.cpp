/* core.cpp -- Object / ObjectFactory / PropertyList / Transform / file resolver / XML loader.
 * Behaviour follows src/kazen/object.cpp:5-20, proplist.cpp:5-41, parser.cpp:10-305,
 * common.cpp:236-296 (string conversions); the XML reader is a small hand-written one
 * (the reference links pugixml, which this image does not have). */
#include <kazen/object.h>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <set>
#include <sstream>
#include <sys/stat.h>

namespace kazen {

std::string fmt(const char *format, ...) {
    char buf[2048];
    va_list ap; va_start(ap, format);
    vsnprintf(buf, sizeof(buf), format, ap);
    va_end(ap);
    return std::string(buf);
}

/* ------------------------------------------------------------------ math */
Mat4 Mat4::identity() { Mat4 r; memset(r.m, 0, sizeof(r.m)); for (int i = 0; i < 4; ++i) r.m[i][i] = 1.f; return r; }
Mat4 Mat4::operator*(const Mat4 &o) const {
    Mat4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float acc = 0.f;
            for (int k = 0; k < 4; ++k) acc += m[i][k] * o.m[k][j];
            r.m[i][j] = acc;
        }
    return r;
}
Mat4 Mat4::transpose() const { Mat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = m[j][i]; return r; }
Mat4 Mat4::inverse() const {
    double a[4][8];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a[i][j] = m[i][j]; a[i][j + 4] = (i == j); }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
        if (a[piv][c] == 0.0) throw Exception("singular transform");
        for (int j = 0; j < 8; ++j) std::swap(a[c][j], a[piv][j]);
        const double inv = 1.0 / a[c][c];
        for (int j = 0; j < 8; ++j) a[c][j] *= inv;
        for (int r = 0; r < 4; ++r) if (r != c) { const double f = a[r][c]; for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j]; }
    }
    Mat4 r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = (float)a[i][j + 4];
    return r;
}
Vec3 Transform::point(const Vec3 &p) const {
    const float (*M)[4] = matrix.m;
    const float x = M[0][0] * p.x + M[0][1] * p.y + M[0][2] * p.z + M[0][3];
    const float y = M[1][0] * p.x + M[1][1] * p.y + M[1][2] * p.z + M[1][3];
    const float z = M[2][0] * p.x + M[2][1] * p.y + M[2][2] * p.z + M[2][3];
    const float w = M[3][0] * p.x + M[3][1] * p.y + M[3][2] * p.z + M[3][3];
    return Vec3{x / w, y / w, z / w};
}
Vec3 Transform::vector(const Vec3 &v) const {
    const float (*M)[4] = matrix.m;
    return Vec3{M[0][0] * v.x + M[0][1] * v.y + M[0][2] * v.z, M[1][0] * v.x + M[1][1] * v.y + M[1][2] * v.z, M[2][0] * v.x + M[2][1] * v.y + M[2][2] * v.z};
}
Vec3 Transform::normal(const Vec3 &n) const {
    const float (*I)[4] = inv.m;      /* transpose of the inverse's upper 3x3 */
    return Vec3{I[0][0] * n.x + I[1][0] * n.y + I[2][0] * n.z, I[0][1] * n.x + I[1][1] * n.y + I[2][1] * n.z, I[0][2] * n.x + I[1][2] * n.y + I[2][2] * n.z};
}

/* ------------------------------------------------------------------ properties / objects */
void PropertyList::set(const std::string &n, Value v, const char *type) {
    if (m_values.count(n)) std::cerr << "Property \"" << n << "\" was specified multiple times!" << std::endl;
    m_values[n] = Entry{std::move(v), type};
}

void Object::addChild(Object *) { throw Exception("Object::addChild() is not implemented for objects of type '" + classTypeName(getClassType()) + "'!"); }
void Object::setParent(Object *) {}
void Object::activate() {}
std::string Object::classTypeName(EClassType type) {
    switch (type) {
        case EScene: return "scene"; case EMesh: return "mesh"; case EBSDF: return "bsdf"; case ELight: return "light";
        case ECamera: return "camera"; case EIntegrator: return "integrator"; case ESampler: return "sampler";
        case ETexture: return "texture"; case EMedium: return "medium"; case EPhaseFunction: return "phase";
        case EReconstructionFilter: return "rfilter"; case EAccel: return "accel"; default: return "<unknown>";
    }
}

std::map<std::string, ObjectFactory::Constructor> *ObjectFactory::m_constructors = nullptr;
void ObjectFactory::registerClass(const std::string &name, const Constructor &constr) {
    if (!m_constructors) m_constructors = new std::map<std::string, Constructor>();
    (*m_constructors)[name] = constr;
}
Object *ObjectFactory::createInstance(const std::string &name, const PropertyList &propList) {
    if (!m_constructors || m_constructors->find(name) == m_constructors->end())
        throw Exception("A constructor for class \"" + name + "\" could not be found!");
    return (*m_constructors)[name](propList);
}
bool ObjectFactory::isRegistered(const std::string &name) { return m_constructors && m_constructors->count(name); }
std::vector<std::string> ObjectFactory::registeredNames() {
    std::vector<std::string> r;
    if (m_constructors) for (auto &kv : *m_constructors) r.push_back(kv.first);
    return r;
}

/* ------------------------------------------------------------------ file resolver */
static std::vector<std::string> &searchPaths() { static std::vector<std::string> p{"."}; return p; }
void resolverPrepend(const std::string &dir) { searchPaths().insert(searchPaths().begin(), dir.empty() ? "." : dir); }
std::string resolvePath(const std::string &path) {
    if (!path.empty() && path[0] == '/') return path;
    struct stat st;
    for (const std::string &d : searchPaths()) {
        const std::string c = d + "/" + path;
        if (stat(c.c_str(), &st) == 0) return c;
    }
    return path;
}

/* ------------------------------------------------------------------ string conversions (common.cpp:236-296) */
static std::vector<std::string> tokenize(const std::string &s, const std::string &delim = ", ", bool includeEmpty = false) {
    std::vector<std::string> tokens;
    std::string::size_type last = 0, pos = s.find_first_of(delim, last);
    while (last != std::string::npos) {
        if (pos != last || includeEmpty) tokens.push_back(s.substr(last, pos - last));
        last = pos;
        if (last != std::string::npos) { last += 1; pos = s.find_first_of(delim, last); }
    }
    return tokens;
}
static bool toBool(const std::string &str) {
    std::string v = str;
    for (char &c : v) c = (char)tolower(c);
    if (v == "false") return false;
    if (v == "true") return true;
    throw Exception("Could not parse boolean value \"" + str + "\"");
}
static int toInt(const std::string &str) {
    char *end = nullptr;
    const int r = (int)strtol(str.c_str(), &end, 10);
    if (*end != '\0') throw Exception("Could not parse integer value \"" + str + "\"");
    return r;
}
static float toFloat(const std::string &str) {
    char *end = nullptr;
    const float r = strtof(str.c_str(), &end);
    if (*end != '\0') throw Exception("Could not parse floating point value \"" + str + "\"");
    return r;
}
static Vec3 toVector3f(const std::string &str) {
    const std::vector<std::string> t = tokenize(str);
    if (t.size() != 3) throw Exception("Expected 3 values");
    return Vec3{toFloat(t[0]), toFloat(t[1]), toFloat(t[2])};
}
static Vec3 normalized(Vec3 v) { const float l = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); return Vec3{v.x / l, v.y / l, v.z / l}; }
static Vec3 cross(Vec3 a, Vec3 b) { return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

/* ------------------------------------------------------------------ a minimal XML reader */
namespace {
struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<XmlNode> children;
    size_t offset = 0;
    const std::string *attr(const std::string &n) const { for (auto &a : attrs) if (a.first == n) return &a.second; return nullptr; }
    std::string value(const std::string &n) const { const std::string *a = attr(n); return a ? *a : std::string(); }
};

class XmlReader {
public:
    XmlReader(const std::string &text, const std::string &file) : s(text), fname(file) {}
    XmlNode parseDocument() {
        skipMisc();
        if (eof()) fail("no root element");
        XmlNode root = parseElement();
        skipMisc();
        if (!eof()) fail("content after the root element");
        return root;
    }
    std::string where(size_t pos) const {
        size_t line = 1, col = 1;
        for (size_t i = 0; i < pos && i < s.size(); ++i) { if (s[i] == '\n') { ++line; col = 1; } else ++col; }
        return fmt("line %zu, col %zu", line, col);
    }
private:
    const std::string &s; std::string fname; size_t p = 0;
    bool eof() const { return p >= s.size(); }
    [[noreturn]] void fail(const std::string &msg) const { throw Exception("Error while parsing \"" + fname + "\": " + msg + " (at " + where(p) + ")"); }
    bool starts(const char *t) const { return s.compare(p, strlen(t), t) == 0; }
    void skipWs() { while (!eof() && isspace((unsigned char)s[p])) ++p; }
    void skipMisc() {      /* whitespace, comments, declaration, doctype */
        for (;;) {
            skipWs();
            if (starts("<!--")) { size_t e = s.find("-->", p); if (e == std::string::npos) fail("unterminated comment"); p = e + 3; }
            else if (starts("<?")) { size_t e = s.find("?>", p); if (e == std::string::npos) fail("unterminated declaration"); p = e + 2; }
            else if (starts("<!")) { size_t e = s.find('>', p); if (e == std::string::npos) fail("unterminated markup"); p = e + 1; }
            else return;
        }
    }
    static std::string unescape(const std::string &v) {
        std::string r; r.reserve(v.size());
        for (size_t i = 0; i < v.size(); ++i) {
            if (v[i] != '&') { r += v[i]; continue; }
            if (v.compare(i, 4, "&lt;") == 0) { r += '<'; i += 3; } else if (v.compare(i, 4, "&gt;") == 0) { r += '>'; i += 3; }
            else if (v.compare(i, 5, "&amp;") == 0) { r += '&'; i += 4; } else if (v.compare(i, 6, "&quot;") == 0) { r += '"'; i += 5; }
            else if (v.compare(i, 6, "&apos;") == 0) { r += '\''; i += 5; } else r += v[i];
        }
        return r;
    }
    std::string parseName() {
        size_t b = p;
        while (!eof() && (isalnum((unsigned char)s[p]) || s[p] == '_' || s[p] == '-' || s[p] == ':' || s[p] == '.')) ++p;
        if (p == b) fail("expected a name");
        return s.substr(b, p - b);
    }
    XmlNode parseElement() {
        if (eof() || s[p] != '<') fail("unexpected content");
        XmlNode n; n.offset = p; ++p;
        n.name = parseName();
        for (;;) {
            skipWs();
            if (eof()) fail("unterminated tag");
            if (starts("/>")) { p += 2; return n; }
            if (s[p] == '>') { ++p; break; }
            std::string an = parseName();
            skipWs();
            if (eof() || s[p] != '=') fail("expected '=' after attribute name");
            ++p; skipWs();
            if (eof() || (s[p] != '"' && s[p] != '\'')) fail("expected a quoted attribute value");
            const char q = s[p++];
            size_t e = s.find(q, p);
            if (e == std::string::npos) fail("unterminated attribute value");
            n.attrs.emplace_back(an, unescape(s.substr(p, e - p)));
            p = e + 1;
        }
        for (;;) {
            skipMisc();
            if (eof()) fail("missing closing tag for <" + n.name + ">");
            if (starts("</")) {
                p += 2;
                const std::string cn = parseName();
                if (cn != n.name) fail("mismatched closing tag </" + cn + "> for <" + n.name + ">");
                skipWs();
                if (eof() || s[p] != '>') fail("malformed closing tag");
                ++p;
                return n;
            }
            if (s[p] != '<') fail("unexpected content");      /* parser.cpp:122-125: text nodes are an error */
            n.children.push_back(parseElement());
        }
    }
};
}  // namespace

/* ------------------------------------------------------------------ loadFromXML (parser.cpp:10-305) */
Object *loadFromXML(const std::string &filename, const ParseOverrides *overrides) {
    std::ifstream is(filename);
    if (is.fail()) throw Exception("Error while parsing \"" + filename + "\": cannot open file");
    std::stringstream ss; ss << is.rdbuf();
    const std::string text = ss.str();
    XmlReader reader(text, filename);
    XmlNode root = reader.parseDocument();

    enum ETag {
        EBoolean = Object::EClassTypeCount, EInteger, EFloat, EString, EPoint, EVector, EColor, ETransform,
        ETranslate, EMatrix, ERotate, EScale, ELookAt, EInvalid
    };
    std::map<std::string, int> tags = {
        {"scene", Object::EScene}, {"mesh", Object::EMesh}, {"bsdf", Object::EBSDF}, {"light", Object::ELight}, {"camera", Object::ECamera},
        {"medium", Object::EMedium}, {"phase", Object::EPhaseFunction}, {"integrator", Object::EIntegrator}, {"sampler", Object::ESampler},
        {"texture", Object::ETexture}, {"rfilter", Object::EReconstructionFilter}, {"accel", Object::EAccel},
        {"boolean", EBoolean}, {"integer", EInteger}, {"float", EFloat}, {"string", EString}, {"point", EPoint}, {"vector", EVector},
        {"color", EColor}, {"transform", ETransform}, {"translate", ETranslate}, {"matrix", EMatrix}, {"rotate", ERotate},
        {"scale", EScale}, {"lookat", ELookAt}};

    auto at = [&](const XmlNode &n) { return reader.where(n.offset); };
    auto checkAttributes = [&](const XmlNode &node, std::set<std::string> attrs) {
        for (auto &a : node.attrs) {
            auto it = attrs.find(a.first);
            if (it == attrs.end())
                throw Exception("Error while parsing \"" + filename + "\": unexpected attribute \"" + a.first + "\" in \"" + node.name + "\" at " + at(node));
            attrs.erase(it);
        }
        if (!attrs.empty())
            throw Exception("Error while parsing \"" + filename + "\": missing attribute \"" + *attrs.begin() + "\" in \"" + node.name + "\" at " + at(node));
    };

    Mat4 transform = Mat4::identity();

    std::function<Object *(XmlNode &, PropertyList &, int)> parseTag = [&](XmlNode &node, PropertyList &list, int parentTag) -> Object * {
        auto it = tags.find(node.name);
        if (it == tags.end()) throw Exception("Error while parsing \"" + filename + "\": unexpected tag \"" + node.name + "\" at " + at(node));
        const int tag = it->second;
        const bool hasParent = parentTag != EInvalid;
        const bool parentIsObject = hasParent && parentTag < Object::EClassTypeCount;
        const bool currentIsObject = tag < Object::EClassTypeCount;
        const bool parentIsTransform = parentTag == ETransform;
        const bool currentIsTransformOp = tag == ETranslate || tag == ERotate || tag == EScale || tag == ELookAt || tag == EMatrix;
        if (!hasParent && !currentIsObject)
            throw Exception("Error while parsing \"" + filename + "\": root element \"" + node.name + "\" must be a kazen object (at " + at(node) + ")");
        if (parentIsTransform != currentIsTransformOp)
            throw Exception("Error while parsing \"" + filename + "\": transform nodes can only contain transform operations (at " + at(node) + ")");
        if (hasParent && !parentIsObject && !(parentIsTransform && currentIsTransformOp))
            throw Exception("Error while parsing \"" + filename + "\": node \"" + node.name + "\" requires a kazen object as parent (at " + at(node) + ")");

        std::string type = node.value("type");
        if (tag == Object::EScene) type = "scene";
        else if (tag == ETransform) transform = Mat4::identity();

        PropertyList propList;
        std::vector<Object *> children;
        for (XmlNode &ch : node.children) {
            Object *child = parseTag(ch, propList, tag);
            if (child) children.push_back(child);
        }

        Object *result = nullptr;
        try {
            if (currentIsObject) {
                if (overrides) {
                    /* values are "<t>:<text>" with t in {i,f,b,s}; a property the XML already set was
                     * replaced while it was parsed, anything else is injected here with the given type */
                    auto ov = overrides->find(node.name);
                    if (ov != overrides->end())
                        for (auto &kv : ov->second) {
                            const std::string &v = kv.second;
                            const std::string body = v.size() > 2 && v[1] == ':' ? v.substr(2) : v;
                            if (kv.first == "type") { type = body; continue; }
                            if (propList.has(kv.first)) continue;
                            switch (v.size() > 2 && v[1] == ':' ? v[0] : 's') {
                                case 'i': propList.setInteger(kv.first, toInt(body)); break;
                                case 'f': propList.setFloat(kv.first, toFloat(body)); break;
                                case 'b': propList.setBoolean(kv.first, toBool(body)); break;
                                default: propList.setString(kv.first, body); break;
                            }
                        }
                }
                result = ObjectFactory::createInstance(type, propList);
                if ((int)result->getClassType() != tag)
                    throw Exception("Unexpectedly constructed an object of type <" + Object::classTypeName(result->getClassType()) + "> (expected type <" +
                                    Object::classTypeName((Object::EClassType)tag) + ">): " + result->toString());
                result->setId(node.value("id"));
                for (Object *ch : children) { result->addChild(ch); ch->setParent(result); }
                result->activate();
            } else {
                const std::string name = node.value("name"), value = node.value("value");
                auto overridden = [&](const std::string &fallback) -> std::string {
                    /* CLI overrides address properties as <parent tag>.<property name> */
                    if (!overrides) return fallback;
                    for (auto &kv : tags) if (kv.second == parentTag) {
                        auto ov = overrides->find(kv.first);
                        if (ov != overrides->end()) {
                            auto pv = ov->second.find(name);
                            if (pv != ov->second.end()) return pv->second.size() > 2 && pv->second[1] == ':' ? pv->second.substr(2) : pv->second;
                        }
                    }
                    return fallback;
                };
                switch (tag) {
                    case EString: checkAttributes(node, {"name", "value"}); list.setString(name, overridden(value)); break;
                    case EFloat: checkAttributes(node, {"name", "value"}); list.setFloat(name, toFloat(overridden(value))); break;
                    case EInteger: checkAttributes(node, {"name", "value"}); list.setInteger(name, toInt(overridden(value))); break;
                    case EBoolean: checkAttributes(node, {"name", "value"}); list.setBoolean(name, toBool(overridden(value))); break;
                    case EPoint: checkAttributes(node, {"name", "value"}); list.setPoint(name, toVector3f(value)); break;
                    case EVector: checkAttributes(node, {"name", "value"}); list.setVector(name, toVector3f(value)); break;
                    case EColor: { checkAttributes(node, {"name", "value"}); const Vec3 v = toVector3f(value); list.setColor(name, Color3{v.x, v.y, v.z}); } break;
                    case ETransform: checkAttributes(node, {"name"}); list.setTransform(name, Transform(transform)); break;
                    case ETranslate: {
                        checkAttributes(node, {"value"});
                        const Vec3 v = toVector3f(value);
                        Mat4 t = Mat4::identity(); t.m[0][3] = v.x; t.m[1][3] = v.y; t.m[2][3] = v.z;
                        transform = t * transform;
                    } break;
                    case EMatrix: {
                        checkAttributes(node, {"value"});
                        const std::vector<std::string> tk = tokenize(value);
                        if (tk.size() != 16) throw Exception("Expected 16 values");
                        Mat4 mm;
                        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) mm.m[i][j] = toFloat(tk[(size_t)i * 4 + j]);
                        transform = mm * transform;
                    } break;
                    case EScale: {
                        checkAttributes(node, {"value"});
                        const Vec3 v = toVector3f(value);
                        Mat4 t = Mat4::identity(); t.m[0][0] = v.x; t.m[1][1] = v.y; t.m[2][2] = v.z;
                        transform = t * transform;
                    } break;
                    case ERotate: {
                        checkAttributes(node, {"angle", "axis"});
                        const float angle = toFloat(node.value("angle")) * (3.14159265358979323846f / 180.0f);
                        const Vec3 a = toVector3f(node.value("axis"));
                        const float c = std::cos(angle), s = std::sin(angle), k = 1.f - c;     /* angle-axis -> matrix (axis taken as given) */
                        Mat4 t = Mat4::identity();
                        t.m[0][0] = c + k * a.x * a.x;       t.m[0][1] = k * a.x * a.y - s * a.z; t.m[0][2] = k * a.x * a.z + s * a.y;
                        t.m[1][0] = k * a.y * a.x + s * a.z; t.m[1][1] = c + k * a.y * a.y;       t.m[1][2] = k * a.y * a.z - s * a.x;
                        t.m[2][0] = k * a.z * a.x - s * a.y; t.m[2][1] = k * a.z * a.y + s * a.x; t.m[2][2] = c + k * a.z * a.z;
                        transform = t * transform;
                    } break;
                    case ELookAt: {
                        checkAttributes(node, {"origin", "target", "up"});
                        const Vec3 origin = toVector3f(node.value("origin")), target = toVector3f(node.value("target")), up = toVector3f(node.value("up"));
                        const Vec3 dir = normalized(Vec3{target.x - origin.x, target.y - origin.y, target.z - origin.z});
                        const Vec3 left = normalized(cross(normalized(up), dir));
                        const Vec3 newUp = normalized(cross(dir, left));
                        Mat4 t = Mat4::identity();
                        t.m[0][0] = left.x; t.m[1][0] = left.y; t.m[2][0] = left.z;
                        t.m[0][1] = newUp.x; t.m[1][1] = newUp.y; t.m[2][1] = newUp.z;
                        t.m[0][2] = dir.x; t.m[1][2] = dir.y; t.m[2][2] = dir.z;
                        t.m[0][3] = origin.x; t.m[1][3] = origin.y; t.m[2][3] = origin.z;
                        transform = t * transform;
                    } break;
                    default: throw Exception("Unhandled element \"" + node.name + "\"");
                }
            }
        } catch (const Exception &e) {
            const std::string w = e.what();
            if (w.rfind("Error while parsing", 0) == 0) throw;      /* already located */
            throw Exception("Error while parsing \"" + filename + "\": " + w + " (at " + at(node) + ")");
        }
        return result;
    };

    PropertyList list;
    return parseTag(root, list, EInvalid);
}

}  // namespace kazen

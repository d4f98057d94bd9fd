/* kazen/object.h -- kazen's plugin/object API for the B200 host (C++17, no third-party deps).
 *
 * Keeps the reference interface a plugin author sees (include/kazen/object.h:13-152,
 * include/kazen/proplist.h, src/kazen/object.cpp:14-20): Object with getClassType / addChild /
 * setParent / activate / toString / setId, ObjectFactory::registerClass / createInstance, the
 * KAZEN_REGISTER_CLASS(cls, "name") macro and the typed PropertyList getters with defaults.
 * Added: EAccel (the accelerator becomes a plugin) -- appended after the reference's
 * class types so their numeric values are unchanged.
 */
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

namespace kazen {

/* printf-style exception (the reference formats with fmt's {}; callers here pass finished strings
 * or use fmt()). */
class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string &msg) : std::runtime_error(msg) {}
};
std::string fmt(const char *format, ...);

struct Vec3 { float x = 0, y = 0, z = 0; };
struct Color3 { float r = 0, g = 0, b = 0; };

/* Row-major 4x4 float matrix + its inverse (the reference's Transform keeps both, transform.h:20-30). */
struct Mat4 {
    float m[4][4];
    static Mat4 identity();
    Mat4 operator*(const Mat4 &o) const;
    Mat4 inverse() const;           /* double Gauss-Jordan, rounded to float */
    Mat4 transpose() const;
};
struct Transform {
    Mat4 matrix = Mat4::identity(), inv = Mat4::identity();
    Transform() {}
    explicit Transform(const Mat4 &mm) : matrix(mm), inv(mm.inverse()) {}
    Vec3 point(const Vec3 &p) const;     /* homogeneous divide, transform.h:59-62 */
    Vec3 vector(const Vec3 &v) const;    /* upper 3x3, transform.h:49-51 */
    Vec3 normal(const Vec3 &n) const;    /* inverse transpose, transform.h:54-56 */
};

class PropertyList {
public:
    using Value = std::variant<bool, int, float, std::string, Color3, Vec3, Transform>;
    void setBoolean(const std::string &n, bool v) { set(n, Value(v), "boolean"); }
    void setInteger(const std::string &n, int v) { set(n, Value(v), "integer"); }
    void setFloat(const std::string &n, float v) { set(n, Value(v), "float"); }
    void setString(const std::string &n, const std::string &v) { set(n, Value(v), "string"); }
    void setColor(const std::string &n, const Color3 &v) { set(n, Value(v), "color"); }
    void setPoint(const std::string &n, const Vec3 &v) { set(n, Value(v), "point"); }
    void setVector(const std::string &n, const Vec3 &v) { set(n, Value(v), "vector"); }
    void setTransform(const std::string &n, const Transform &v) { set(n, Value(v), "transform"); }

    bool getBoolean(const std::string &n) const { return get<bool>(n, "boolean"); }
    bool getBoolean(const std::string &n, bool d) const { return has(n) ? get<bool>(n, "boolean") : d; }
    int getInteger(const std::string &n) const { return get<int>(n, "integer"); }
    int getInteger(const std::string &n, int d) const { return has(n) ? get<int>(n, "integer") : d; }
    float getFloat(const std::string &n) const { return get<float>(n, "float"); }
    float getFloat(const std::string &n, float d) const { return has(n) ? get<float>(n, "float") : d; }
    std::string getString(const std::string &n) const { return get<std::string>(n, "string"); }
    std::string getString(const std::string &n, const std::string &d) const { return has(n) ? get<std::string>(n, "string") : d; }
    Color3 getColor(const std::string &n) const { return get<Color3>(n, "color"); }
    Color3 getColor(const std::string &n, const Color3 &d) const { return has(n) ? get<Color3>(n, "color") : d; }
    Vec3 getPoint(const std::string &n) const { return get<Vec3>(n, "point"); }
    Vec3 getPoint(const std::string &n, const Vec3 &d) const { return has(n) ? get<Vec3>(n, "point") : d; }
    Vec3 getVector(const std::string &n) const { return get<Vec3>(n, "vector"); }
    Vec3 getVector(const std::string &n, const Vec3 &d) const { return has(n) ? get<Vec3>(n, "vector") : d; }
    Transform getTransform(const std::string &n) const { return get<Transform>(n, "transform"); }
    Transform getTransform(const std::string &n, const Transform &d) const { return has(n) ? get<Transform>(n, "transform") : d; }
    bool has(const std::string &n) const { return m_values.count(n) != 0; }

private:
    struct Entry { Value v; const char *type; };
    std::map<std::string, Entry> m_values;
    void set(const std::string &n, Value v, const char *type);
    template <typename T> T get(const std::string &n, const char *type) const {
        auto it = m_values.find(n);
        if (it == m_values.end()) throw Exception("Property '" + n + "' is missing!");
        /* point and vector share a C++ type: compare the XML type name like the reference does */
        if (std::string(it->second.type) != type) throw Exception("Property '" + n + "' has the wrong type! (expected <" + type + ">)!");
        return std::get<T>(it->second.v);
    }
};

class Object {
public:
    enum EClassType {
        EScene = 0, EMesh, EBSDF, EPhaseFunction, ELight, EMedium, ECamera, EIntegrator, ESampler,
        EReconstructionFilter, ETexture,
        EAccel,                 /* new: accelerator plugins ("gpu_bvh") */
        EClassTypeCount
    };
    virtual ~Object() {}
    virtual EClassType getClassType() const = 0;
    virtual void addChild(Object *child);
    virtual void setParent(Object *parent);
    virtual void activate();
    virtual std::string toString() const = 0;
    void setId(const std::string &id) { m_id = id; }
    const std::string &getId() const { return m_id; }
    static std::string classTypeName(EClassType type);

protected:
    std::string m_id;
};

class ObjectFactory {
public:
    typedef std::function<Object *(const PropertyList &)> Constructor;
    /* last registration of a name wins (object.cpp:19) */
    static void registerClass(const std::string &name, const Constructor &constr);
    static Object *createInstance(const std::string &name, const PropertyList &propList);
    static bool isRegistered(const std::string &name);
    static std::vector<std::string> registeredNames();

private:
    static std::map<std::string, Constructor> *m_constructors;
};

#define KAZEN_REGISTER_CLASS(cls, name)                                                   \
    cls *cls##_create(const ::kazen::PropertyList &list) { return new cls(list); }        \
    static struct cls##_ {                                                                \
        cls##_() { ::kazen::ObjectFactory::registerClass(name, cls##_create); }           \
    } cls##__KAZEN_;

/* parser.h: load a scene description; overrides["camera"]["width"] = "512" replaces/sets a
 * property of every object of that tag before its constructor runs (CLI --spp/--size). */
typedef std::map<std::string, std::map<std::string, std::string>> ParseOverrides;
Object *loadFromXML(const std::string &filename, const ParseOverrides *overrides = nullptr);

/* file resolver (main.cpp:52): directories searched for relative asset paths */
void resolverPrepend(const std::string &dir);
std::string resolvePath(const std::string &path);

}  // namespace kazen

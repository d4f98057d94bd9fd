/* kazen/scene.h -- host-side scene objects of the B200 build.
 *
 * The plugin classes keep kazen's registered names, XML properties and defaults, but they are
 * DESCRIPTORS: instead of implementing eval()/sample()/rayIntersect() on the CPU they flatten
 * into the POD tables of include/kzgpu.h, and the render runs on the GPU behind the C ABI.
 * Reference counterparts are cited at each class.
 */
#pragma once
#include <kazen/object.h>
#include <memory>
#include "kzgpu.h"

namespace kazen {

class Scene;

/* texture.h:9-21 / texture.cpp:10-270 */
class Texture : public Object {
public:
    EClassType getClassType() const override { return ETexture; }
    /* appends this node (children first) to `out`, returns its index */
    virtual int flatten(struct FlattenCtx &ctx) const = 0;
};

/* bsdf.h:58-127 */
class BSDF : public Object {
public:
    EClassType getClassType() const override { return EBSDF; }
    virtual int flatten(struct FlattenCtx &ctx) const = 0;
};

/* light.h / light.cpp:7-66 */
class Light : public Object {
public:
    EClassType getClassType() const override { return ELight; }
    virtual kz_light_desc describe() const = 0;
};

/* rfilter.h:16-37 */
class ReconstructionFilter : public Object {
public:
    EClassType getClassType() const override { return EReconstructionFilter; }
    float getRadius() const { return m_radius; }
    virtual float eval(float x) const = 0;
    kz_filter_desc tabulate() const;           /* block.cpp:13-21 */
protected:
    float m_radius = 0.f;
};

/* camera.h:14-56 */
class Camera : public Object {
public:
    ~Camera() override { delete m_rfilter; }
    EClassType getClassType() const override { return ECamera; }
    void addChild(Object *obj) override;
    void activate() override;                  /* camera.cpp:35-68 / 156-189 */
    const ReconstructionFilter *getReconstructionFilter() const { return m_rfilter; }
    int width() const { return m_width; }
    int height() const { return m_height; }
    virtual kz_camera_desc describe() const = 0;
protected:
    void readCommon(const PropertyList &p);
    int m_width = 1280, m_height = 720;
    float m_fov = 30.f, m_near = 1e-4f, m_far = 1e4f;
    Transform m_cameraToWorld;
    Mat4 m_sampleToCamera = Mat4::identity();
    ReconstructionFilter *m_rfilter = nullptr;
};

/* sampler.h:44-107; here only the constructor's rounding rules live on the host */
class Sampler : public Object {
public:
    EClassType getClassType() const override { return ESampler; }
    virtual kz_sampler_desc describe() const = 0;
    uint32_t getSampleCount() const { return m_sampleCount; }
protected:
    uint32_t m_sampleCount = 1;
    uint64_t m_seed = 1;
};

/* mesh.h:63-185; buffers as the reference holds them (packed xyz / uv / index triples) */
class Mesh : public Object {
public:
    ~Mesh() override;
    EClassType getClassType() const override { return EMesh; }
    void addChild(Object *child) override;     /* mesh.cpp:136-161 */
    void activate() override;                  /* mesh.cpp:24-45: default diffuse BSDF */
    std::string toString() const override;
    bool isLight() const { return m_light != nullptr; }
    const BSDF *getBSDF() const { return m_bsdf; }
    const Light *getLight() const { return m_light; }
    std::vector<float> m_V, m_N, m_UV;
    std::vector<uint32_t> m_F;
    std::string m_name;
protected:
    BSDF *m_bsdf = nullptr;
    Light *m_light = nullptr;
};

/* The frame the GPU integrator fills: block.h ImageBlock for the whole film, (r*w, g*w, b*w, w). */
struct ImageBlock {
    int width = 0, height = 0, border = 0;
    std::vector<float> data;       /* (height+2b) x (width+2b) x 4 */
};

/* integrator.h:16-43 + the whole-frame hook of SURVEY 8b */
class Integrator : public Object {
public:
    EClassType getClassType() const override { return EIntegrator; }
    virtual void preprocess(const Scene *) {}
    /* Whole-frame render; returns false if the integrator only implements the per-ray interface. */
    virtual bool renderFrame(Scene *scene, ImageBlock &result) = 0;
    virtual kz_integrator_desc describe() const = 0;
};

/* accel.h:14-64 made a plugin: addMesh / build keep their meaning, rayIntersect becomes a batch call */
class Accel : public Object {
public:
    EClassType getClassType() const override { return EAccel; }
    virtual void addMesh(Mesh *mesh) = 0;
    virtual void build() = 0;                        /* host side only records the choice; the BVH is built on upload */
    virtual int builder() const = 0;                 /* KZ_BUILD_* */
};

/* Owner of everything a kz_scene_desc points to. */
struct FlattenCtx {
    std::vector<kz_texture_desc> textures;
    std::vector<kz_image_desc> images;
    std::vector<std::vector<float>> image_data;
    std::vector<kz_bsdf_desc> bsdfs;
    std::vector<kz_light_desc> lights;
    std::vector<kz_mesh_desc> meshes;
    std::vector<uint16_t> blue_noise;
    std::vector<uint32_t> pmj;
    std::map<const Object *, int> memo;       /* shared objects are flattened once */
    kz_scene_desc desc{};
};

/* scene.h:15-138 / scene.cpp:17-126 */
class Scene : public Object {
public:
    explicit Scene(const PropertyList &props);
    ~Scene() override;
    EClassType getClassType() const override { return EScene; }
    void addChild(Object *obj) override;
    void activate() override;
    std::string toString() const override;
    const std::vector<Mesh *> &getMeshes() const { return m_meshes; }
    const Camera *getCamera() const { return m_camera; }
    Sampler *getSampler() const { return m_sampler; }
    Integrator *getIntegrator() const { return m_integrator; }
    const Accel *getAccel() const { return m_accel; }
    /* builds (once) the POD tables the C ABI consumes */
    const kz_scene_desc &flatten();
    int gpus = 1;                                   /* devices the GPU integrator may use (CLI --gpus) */
    /* progressive / resumable accumulation: render only the sample indices [sppBegin, sppEnd) of the sampler's sampleCount
     * (sppEnd < 0: all) and start from a previously saved raw frame instead of an empty one (CLI --spp-range, --resume) */
    int sppBegin = 0, sppEnd = -1;
    std::string resumeFrame;
private:
    std::vector<Mesh *> m_meshes;
    Camera *m_camera = nullptr;
    Sampler *m_sampler = nullptr;
    Integrator *m_integrator = nullptr;
    Texture *m_background = nullptr;
    Accel *m_accel = nullptr;
    std::unique_ptr<FlattenCtx> m_flat;
};

/* renderer.h: render the scene and write <stem>.png (+ optional raw float frame) */
namespace renderer {
void render(Scene *scene, const std::string &outputName, bool writeRaw = false);
}

/* image I/O used by imagetexture and the PNG writer (bitmap.cpp:38-64) */
void writePNG(const std::string &path, int w, int h, const uint8_t *rgb8);
void writeEXR(const std::string &path, int w, int h, const float *rgb);        /* linear float RGB, bitmap.cpp:23-36 */
bool readImage(const std::string &path, int &w, int &h, std::vector<float> &rgb, std::string &err);   /* PNG (8/16 bit) / JPEG / PFM / HDR / EXR */
bool readJPEG(const std::string &path, int &w, int &h, std::vector<float> &rgb, std::string &err);    /* baseline / sequential Huffman, jpeg.cpp */

/* Stand-in pmj02bn / blue-noise tables (the reference's table sources are missing from its
 * public tree): Owen-scrambled Sobol' (0,2) point sets + hashed dither.  NOT pbrt's tables. */
void fallbackPmjTables(std::vector<uint16_t> &blueNoise, std::vector<uint32_t> &pmj);

}  // namespace kazen

// idkpt.hpp -- header-only C++17 host mirror of IDKEngine.Render.PathTracer over the C ABI of idkpt.h.
//
// The reference host is compiled C#; .NET is not available in this image, so the class a maintainer would write as
// `PathTracerNative : IDisposable` is provided here in C++ with the reference's member names, argument
// meaning and reset-on-set behaviour (PathTracer.cs:12-125,170-346). Errors become idk::Error carrying the IdkPtStatus and
// idkpt_last_error(). Link with -lidkpt (no torch, no CUDA headers needed on the host side).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "idkpt.h"
#include "idkvx.h"

namespace idk {

class Error : public std::runtime_error {
public:
    Error(int status, const std::string& what) : std::runtime_error(what), status_(status) {}
    int status() const { return status_; }

private:
    int status_;
};

// new PathTracer.GpuSettings() defaults (PathTracer.cs:127-138) + the class defaults (RayDepth 7, 1 spp; :12,211)
inline IdkPtSettings DefaultSettings() {
    IdkPtSettings s = {};
    s.Gpu.FocalLength = 8.0f;
    s.Gpu.LenseRadius = 0.0f;
    s.Gpu.DoDebugBVHTraversal = 0;
    s.Gpu.DoTraceLights = 0;
    s.Gpu.DoRussianRoulette = 1;
    s.RayDepth = 7;
    s.SamplesPerPixel = 1;
    return s;
}

// TonemapAndGammaCorrect.GpuSettings + Bloom.GpuSettings + Application.IsBloom defaults
inline IdkPtPostSettings DefaultPostSettings() {
    IdkPtPostSettings p = {};
    p.Exposure = 0.45f; p.Saturation = 1.06f; p.Linear = 0.18f; p.Peak = 1.0f; p.Compression = 0.1f;
    p.DoTonemapAndSrgbTransform = 1;
    p.IsBloom = 1; p.BloomThreshold = 1.5f; p.BloomMaxColor = 3.8f; p.BloomMinusLods = 3;
    return p;
}

struct Tile {
    int StripeHeight = 8, Index = 0, Count = 1;   // multi-GPU screen split: one PathTracer per GPU
    bool GlobalSlots = false;                     // IDKPT_CREATE_GLOBAL_SLOTS: the N tiles reproduce the untiled image bit for bit
};

class PathTracer {
public:
    // PathTracer(int width, int height, in GpuSettings settings)   PathTracer.cs:170
    PathTracer(int width, int height, const IdkPtGpuSettings& gpuSettings = DefaultSettings().Gpu, int device = 0, Tile tile = Tile(), int lanes = 0)
        : settings_(DefaultSettings()), width_(width), height_(height) {
        settings_.Gpu = gpuSettings;
        IdkPtCreateInfo ci = {};
        ci.Device = device; ci.Width = width; ci.Height = height;
        ci.TileStripeHeight = tile.StripeHeight; ci.TileIndex = tile.Index; ci.TileCount = tile.Count;
        ci.Flags = IDKPT_CREATE_LANES(lanes) | (tile.GlobalSlots ? IDKPT_CREATE_GLOBAL_SLOTS : 0u);
        const int rc = idkpt_create(&ci, &ctx_);
        if (rc != IDKPT_OK) {
            const char* msg = idkpt_last_error(nullptr);
            throw Error(rc, std::string("idkpt_create failed: ") + (msg ? msg : ""));
        }
    }
    ~PathTracer() { Dispose(); }
    PathTracer(const PathTracer&) = delete;
    PathTracer& operator=(const PathTracer&) = delete;
    PathTracer(PathTracer&& o) noexcept : ctx_(o.ctx_), settings_(o.settings_), width_(o.width_), height_(o.height_) { o.ctx_ = nullptr; }
    PathTracer& operator=(PathTracer&& o) noexcept {
        if (this != &o) { Dispose(); ctx_ = o.ctx_; settings_ = o.settings_; width_ = o.width_; height_ = o.height_; o.ctx_ = nullptr; }
        return *this;
    }
    void Dispose() {   // PathTracer.cs:344
        if (ctx_) { idkpt_destroy(ctx_); ctx_ = nullptr; }
    }

    // ---- the reference's public surface -------------------------------------------------------------------------------
    // Compute(): PathTracer.cs:214-271. Asynchronous (several samples in flight); pass `stats` for the synchronous, timed form.
    void Compute(const GpuPerFrameData& frame, IdkPtStats* stats = nullptr) { check(idkpt_compute(ctx_, &frame, &settings_, stats), "idkpt_compute"); }
    void Sync() { check(idkpt_sync(ctx_), "idkpt_sync"); }
    void SetSize(int width, int height) {   // :299-332
        check(idkpt_resize(ctx_, width, height), "idkpt_resize");
        width_ = width; height_ = height;
    }
    void ResetAccumulation() { check(idkpt_reset_accumulation(ctx_), "idkpt_reset_accumulation"); }   // :334
    const IdkPtGpuSettings& GetGpuSettings() const { return settings_.Gpu; }                           // :339
    uint32_t AccumulatedSamples() const { return idkpt_accumulated_samples(ctx_); }                    // :27-37
    int Width() const { return width_; }
    int Height() const { return height_; }

    // properties; the setters that reset the accumulation in the reference do so here (:16-25, :39-97)
    int RayDepth() const { return settings_.RayDepth; }
    void RayDepth(int v) { settings_.RayDepth = v; ResetAccumulation(); }
    float FocalLength() const { return settings_.Gpu.FocalLength; }
    void FocalLength(float v) { settings_.Gpu.FocalLength = v; ResetAccumulation(); }
    float LenseRadius() const { return settings_.Gpu.LenseRadius; }
    void LenseRadius(float v) { settings_.Gpu.LenseRadius = v; ResetAccumulation(); }
    bool DoDebugBVHTraversal() const { return settings_.Gpu.DoDebugBVHTraversal != 0; }
    void DoDebugBVHTraversal(bool v) { settings_.Gpu.DoDebugBVHTraversal = v; ResetAccumulation(); }
    bool DoTraceLights() const { return settings_.Gpu.DoTraceLights != 0; }
    void DoTraceLights(bool v) { settings_.Gpu.DoTraceLights = v; ResetAccumulation(); }
    bool DoRussianRoulette() const { return settings_.Gpu.DoRussianRoulette != 0; }
    void DoRussianRoulette(bool v) { settings_.Gpu.DoRussianRoulette = v; ResetAccumulation(); }
    int SamplesPerPixel() const { return settings_.SamplesPerPixel; }        // :12
    void SamplesPerPixel(int v) { settings_.SamplesPerPixel = v; }
    bool DoRaySorting() const { return settings_.DoRaySorting != 0; }        // :101-111
    void DoRaySorting(bool v) { settings_.DoRaySorting = v; }
    bool OutputAOVs() const { return settings_.OutputAOVs != 0; }            // :113-125
    void OutputAOVs(bool v) { settings_.OutputAOVs = v; }
    bool CollectStats() const { return settings_.CollectStats != 0; }
    void CollectStats(bool v) { settings_.CollectStats = v; }

    // Result / AlbedoTexture / NormalTexture (:143,167-168): rgba32f, width*height*4 floats, full-image layout
    std::vector<float> Result() const { return read(IDKPT_IMAGE_RESULT); }
    std::vector<float> AlbedoTexture() const { return read(IDKPT_IMAGE_ALBEDO); }
    std::vector<float> NormalTexture() const { return read(IDKPT_IMAGE_NORMAL); }

    // ---- what the reference passes implicitly through bound buffers ------------------------------------------------------
    void SetScene(const IdkPtSceneDesc& scene) { check(idkpt_set_scene(ctx_, &scene), "idkpt_set_scene"); }
    void UpdateRange(IdkPtArrayId which, uint64_t first, uint64_t count, const void* data) { check(idkpt_update_range(ctx_, which, first, count, data), "idkpt_update_range"); }
    void SetTextures(const IdkPtTextureDesc* textures, uint64_t count) { check(idkpt_set_textures(ctx_, textures, count), "idkpt_set_textures"); }
    void SetSky(const IdkPtSkyDesc& sky) { check(idkpt_set_sky(ctx_, &sky), "idkpt_set_sky"); }

    // ---- presentation / interop -------------------------------------------------------------------------------------------
    void PresentAsync(void* pinnedHostRgba32f, uint64_t bytes, IdkPtImage which = IDKPT_IMAGE_RESULT) { check(idkpt_present_async(ctx_, which, pinnedHostRgba32f, bytes), "idkpt_present_async"); }
    void PresentWait() { check(idkpt_present_wait(ctx_), "idkpt_present_wait"); }
    float TlasBuild(int searchRadius = 15) { float ms = 0.0f; check(idkpt_tlas_build(ctx_, searchRadius, &ms), "idkpt_tlas_build"); return ms; }   // BVH.TlasBuild on the device
    float Denoise(const IdkPtDenoiseSettings& s) { float ms = 0.0f; check(idkpt_denoise(ctx_, &s, &ms), "idkpt_denoise"); return ms; }   // PathTracerPipeline.Denoise
    std::vector<float> Denoised() const { return read(IDKPT_IMAGE_DENOISED); }
    void RegisterHostBuffer(void* hostPtr, uint64_t bytes) { check(idkpt_register_host_buffer(ctx_, hostPtr, bytes), "idkpt_register_host_buffer"); }
    void UnregisterHostBuffer(void* hostPtr) { check(idkpt_unregister_host_buffer(ctx_, hostPtr), "idkpt_unregister_host_buffer"); }
    // Bloom + TonemapAndGammaCorrect -> RGBA8 (Application.cs:217-223); out may be null to keep the frame on the device
    float PostProcess(const IdkPtPostSettings& post, uint8_t* rgba8Out, IdkPtImage source = IDKPT_IMAGE_RESULT) {
        float ms = 0.0f;
        check(idkpt_post_process(ctx_, &post, source, rgba8Out, &ms), "idkpt_post_process");
        return ms;
    }
    void* StreamHandle() const { void* s = nullptr; check(idkpt_stream_handle(ctx_, &s), "idkpt_stream_handle"); return s; }
    std::pair<void*, uint64_t> ResultDevicePtr(IdkPtImage which = IDKPT_IMAGE_RESULT) const {
        void* p = nullptr; uint64_t n = 0;
        check(idkpt_result_device_ptr(ctx_, which, &p, &n), "idkpt_result_device_ptr");
        return {p, n};
    }
    std::vector<int32_t> TileRows() const {
        int32_t n = 0;
        check(idkpt_tile_rows(ctx_, &n, nullptr, 0), "idkpt_tile_rows");
        std::vector<int32_t> rows((size_t)n);
        check(idkpt_tile_rows(ctx_, &n, rows.data(), n), "idkpt_tile_rows");
        return rows;
    }

    // ---- neighbours of the path ------------------------------------------------------------------------------------
    std::vector<IdkPtHit> TraceRays(const std::vector<IdkPtRay>& rays, bool traceLights = false, bool anyHit = false) {
        std::vector<IdkPtHit> hits(rays.size());
        float ms = 0.0f;
        check((anyHit ? idkpt_trace_rays_any : idkpt_trace_rays)(ctx_, rays.data(), rays.size(), traceLights, hits.data(), &ms), "idkpt_trace_rays");
        return hits;
    }
    void ShadowsRayTraced(const GpuPerFrameData& frame, const float* depth, const float* normalRG, int width, int height, int lightIndex, int samples,
                          uint32_t noiseIndex, const float* taaJitter, float* visibilityInOut) {
        check(idkpt_shadows_ray_traced(ctx_, &frame, depth, normalRG, width, height, lightIndex, samples, noiseIndex, taaJitter, visibilityInOut, nullptr), "idkpt_shadows_ray_traced");
    }
    void SetSkinningData(const GpuUnskinnedVertex* vertices, uint64_t count) { check(idkpt_set_skinning_data(ctx_, vertices, count), "idkpt_set_skinning_data"); }
    void SkinVertices(const float* jointMatrices3x4, uint64_t jointCount, const IdkPtSkinningCmd* cmds, uint32_t cmdCount) {
        check(idkpt_skin_vertices(ctx_, jointMatrices3x4, jointCount, cmds, cmdCount, nullptr), "idkpt_skin_vertices");
    }
    void BlasRefit(uint32_t firstBlas, uint32_t count = 1) { check(idkpt_blas_refit(ctx_, firstBlas, count, nullptr), "idkpt_blas_refit"); }
    void ReadRange(IdkPtArrayId which, uint64_t first, uint64_t count, void* out) const { check(idkpt_read_range(ctx_, which, first, count, out), "idkpt_read_range"); }

    IdkPtCtx* Handle() const { return ctx_; }

    // One process driving N GPUs (like the engine): tracers in tile order (tracer r created with Tile{.., r, N}). Afterwards every
    // Compute() also delivers this tile into every tracer's full frame over peer memory (PresentAsync(.., IDKPT_IMAGE_GATHERED)).
    // With one host thread only queue work (Compute without stats): a synchronous call would wait for peers not yet submitted.
    static void ConnectPeers(const std::vector<PathTracer*>& tracers) {
        std::vector<IdkPtCtx*> ctxs;
        for (PathTracer* t : tracers) ctxs.push_back(t->ctx_);
        const int rc = idkpt_gather_connect(ctxs.data(), (int32_t)ctxs.size());
        if (rc != IDKPT_OK) {
            std::string msg;
            for (IdkPtCtx* c : ctxs) { const char* m = idkpt_last_error(c); if (m && *m) { msg = m; break; } }
            throw Error(rc, "idkpt_gather_connect failed: " + msg);
        }
    }

private:
    void check(int rc, const char* what) const {
        if (rc != IDKPT_OK) {
            const char* msg = idkpt_last_error(ctx_);
            throw Error(rc, std::string(what) + " failed: " + (msg ? msg : ""));
        }
    }
    std::vector<float> read(IdkPtImage which) const {
        std::vector<float> img((size_t)width_ * height_ * 4);
        check(idkpt_read_result(ctx_, which, img.data(), img.size() * sizeof(float)), "idkpt_read_result");
        return img;
    }

    IdkPtCtx* ctx_ = nullptr;
    IdkPtSettings settings_;
    int width_, height_;
};

// Voxelizer.Render() + ConeTracer.Compute() (Source/Render/VXGI/Voxelizer/Voxelizer.cs:57-114, ConeTracing/ConeTracer.cs:10-50)
class Voxelizer {
public:
    Voxelizer(int width, int height, int depth, const float gridMin[3], const float gridMax[3], int device = 0) {
        IdkVxCreateInfo ci = {};
        ci.Device = device; ci.Width = width; ci.Height = height; ci.Depth = depth;
        for (int i = 0; i < 3; i++) { ci.GridMin[i] = gridMin[i]; ci.GridMax[i] = gridMax[i]; }
        const int rc = idkvx_create(&ci, &ctx_);
        if (rc != IDKPT_OK) {
            const char* msg = idkvx_last_error(nullptr);
            throw Error(rc, std::string("idkvx_create failed: ") + (msg ? msg : ""));
        }
    }
    ~Voxelizer() { Dispose(); }
    Voxelizer(const Voxelizer&) = delete;
    Voxelizer& operator=(const Voxelizer&) = delete;
    void Dispose() { if (ctx_) { idkvx_destroy(ctx_); ctx_ = nullptr; } }

    void SetScene(const IdkPtSceneDesc& scene) { check(idkvx_set_scene(ctx_, &scene), "idkvx_set_scene"); }
    void SetGrid(const float gridMin[3], const float gridMax[3]) { check(idkvx_set_grid(ctx_, gridMin, gridMax), "idkvx_set_grid"); }   // GridMin / GridMax setters
    int LevelCount() const { return idkvx_level_count(ctx_); }
    IdkVxStats Render() { IdkVxStats st = {}; check(idkvx_voxelize(ctx_, &st), "idkvx_voxelize"); return st; }
    // ConeTracer.Compute: G-buffer attachments in, rgba32f indirect light out (width * height * 4 floats)
    std::vector<float> ConeTrace(const GpuPerFrameData& frame, const IdkVxConeSettings& settings, const float* depth, const float* normalRG,
                                 const float* metallicRoughness, int width, int height, const float skyColor[3], IdkVxStats* stats = nullptr) {
        std::vector<float> out((size_t)width * height * 4);
        check(idkvx_cone_trace(ctx_, &frame, &settings, depth, normalRG, metallicRoughness, width, height, skyColor, out.data(), stats), "idkvx_cone_trace");
        return out;
    }
    // ConeTracer.GpuSettings defaults (ConeTracer.cs:10-22)
    static IdkVxConeSettings DefaultConeSettings() {
        IdkVxConeSettings c = {};
        c.MaxSamples = 4; c.StepMultiplier = 0.16f; c.GIBoost = 1.3f; c.GISkyBoxBoost = 1.0f / 1.3f; c.NormalRayOffset = 1.0f; c.NoiseIndex = 0;
        return c;
    }

private:
    void check(int rc, const char* what) const {
        if (rc != IDKPT_OK) {
            const char* msg = idkvx_last_error(ctx_);
            throw Error(rc, std::string(what) + " failed: " + (msg ? msg : ""));
        }
    }
    IdkVxCtx* ctx_ = nullptr;
};

}  // namespace idk

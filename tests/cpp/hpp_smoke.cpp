// Compile + link check of the C++ host mirror (include/idkpt.hpp). Without a CUDA device the constructor must throw
// idk::Error(IDKPT_ERR_NO_DEVICE): the product has no CPU fallback. With a device it renders nothing (no scene) and checks
// that Compute() without a scene reports IDKPT_ERR_NO_SCENE.
#include <cstdio>
#include <cstring>

#include "idkpt.hpp"

int main() {
    static_assert(sizeof(IdkPtSettings) == 40, "IdkPtSettings layout");
    try {
        idk::PathTracer pt(64, 48);
        GpuPerFrameData frame;
        std::memset(&frame, 0, sizeof(frame));
        try {
            pt.Compute(frame);
            std::puts("FAIL: Compute without a scene succeeded");
            return 1;
        } catch (const idk::Error& e) {
            if (e.status() != IDKPT_ERR_NO_SCENE) { std::printf("FAIL: status %d\n", e.status()); return 1; }
        }
        pt.RayDepth(5);
        if (pt.RayDepth() != 5 || pt.AccumulatedSamples() != 0) return 1;
        const float mn[3] = {-1, -1, -1}, mx[3] = {1, 1, 1};
        idk::Voxelizer vx(16, 16, 16, mn, mx);
        if (vx.LevelCount() != 5) return 1;
        try { vx.Render(); return 1; } catch (const idk::Error& e) { if (e.status() != IDKPT_ERR_NO_SCENE) return 1; }
        {   // single-process multi-GPU wiring: two tile contexts with global slots connect; a mismatched pair is refused
            idk::Tile t0, t1;
            t0.Index = 0; t0.Count = 2; t0.GlobalSlots = true;
            t1.Index = 1; t1.Count = 2; t1.GlobalSlots = true;
            idk::PathTracer a(64, 48, idk::DefaultSettings().Gpu, 0, t0, 2), b(64, 48, idk::DefaultSettings().Gpu, 0, t1, 2);
            idk::PathTracer::ConnectPeers({&a, &b});
            try { idk::PathTracer::ConnectPeers({&b, &a}); return 1; } catch (const idk::Error& e) { if (e.status() != IDKPT_ERR_INVALID_ARGUMENT) return 1; }
        }
        std::puts("OK device");
        return 0;
    } catch (const idk::Error& e) {
        if (e.status() == IDKPT_ERR_NO_DEVICE) { std::printf("OK no-device: %s\n", e.what()); return 0; }
        std::printf("FAIL: %d %s\n", e.status(), e.what());
        return 1;
    }
}

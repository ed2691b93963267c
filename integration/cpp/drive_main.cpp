// drive_main.cpp -- tiny C++ driver that uses the host mirror the way the reference's loadbsp step uses
// raytracer.Environment (cmd/tasks/loadbsp/main.go:276,148): add geometry, build, trace packets.
// Prints one line per packet lane; tests/test_gpu_cpp_driver.py compares the output with the oracle.
#include <cstdio>
#include "vrad_environment.hpp"

int main() {
    raytracer::Environment env;
    // a 256^3 room (6 inward quads) with two box occluders, ids as loadbsp assigns them
    env.AddAxisAlignedRectangularSolid(raytracer::TRACE_ID_OPAQUE, {0, 0, 0}, {256, 256, 256});
    env.AddAxisAlignedRectangularSolid(raytracer::TRACE_ID_OPAQUE + 16, {40, 40, 0}, {100, 90, 64});
    env.AddAxisAlignedRectangularSolid(raytracer::TRACE_ID_OPAQUE + 32, {150, 120, 0}, {220, 200, 128});
    env.SetupAccelerationStructure();
    const float dirs[4][3] = {{0.6f, 0.0f, -0.8f}, {-0.6f, 0.48f, -0.64f}, {0.0f, 1.0f, 0.0f}, {0.36f, 0.48f, 0.8f}};
    for (int p = 0; p < 4; p++) {
        raytracer::FourRays rays;
        raytracer::Flt4x tmin{0, 0, 0, 0}, tmax{1000, 1000, 1000, 1000};
        for (int l = 0; l < 4; l++) {
            rays.Origin.X[l] = 20.0f + 50.0f * l + 3.0f * p; rays.Origin.Y[l] = 30.0f + 40.0f * p; rays.Origin.Z[l] = 200.0f;
            rays.Direction.X[l] = dirs[(l + p) % 4][0]; rays.Direction.Y[l] = dirs[(l + p) % 4][1]; rays.Direction.Z[l] = dirs[(l + p) % 4][2];
        }
        raytracer::RayTracingResult res;
        env.Trace4Rays(rays, tmin, tmax, &res);
        for (int l = 0; l < 4; l++)
            std::printf("%d %d %d %.9g %.9g %.9g %.9g\n", p, l, res.HitIds[l], res.HitDistance[l],
                        res.SurfaceNormal.X[l], res.SurfaceNormal.Y[l], res.SurfaceNormal.Z[l]);
    }
    raytracer::FourVectors a, b;
    for (int l = 0; l < 4; l++) { a.X[l] = 10; a.Y[l] = 10 + 60.0f * l; a.Z[l] = 10; b.X[l] = 250; b.Y[l] = 240 - 50.0f * l; b.Z[l] = 20 + 60.0f * l; }
    raytracer::Flt4x vis;
    trace::TestLineDoesHitSky(env, a, b, &vis);
    std::printf("vis %g %g %g %g\n", vis[0], vis[1], vis[2], vis[3]);
    return 0;
}

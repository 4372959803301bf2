// Example.simplesphere (PTSharpCore/Example.cs:1670-1697) written against the C++ host mirror of the reference API.
// The scene-building lines are the reference's, statement for statement; only the language changed.
//   g++ -O2 -std=c++17 -I. examples/simplesphere.cpp -Lptsharp_b200/_lib -lpthost -lptgpu -Wl,-rpath,$PWD/ptsharp_b200/_lib -o simplesphere
#include <cstdio>

#include "ptsharp_b200/host/ptsharp.hpp"

using namespace ptsharp;

int main(int argc, char** argv) {
    int width = 512, height = 512;
    Scene scene;
    // create a material
    auto material = Material::DiffuseMaterial(Colour::White);
    // add the floor (a plane)
    auto plane = Plane::NewPlane(Vector(0, 0, 0), Vector(0, 0, 1), material);
    scene.Add(plane);
    // add the ball (a sphere)
    auto sphere = Sphere::NewSphere(Vector(0, 0, 1), 1.0F, material);
    scene.Add(sphere);
    // add a spherical light source
    auto light = Sphere::NewSphere(Vector(0, 0, 5.0F), 1.0F, Material::LightMaterial(Colour::White, 8));
    scene.Add(light);
    // position the camera
    auto camera = Camera::LookAt(Vector(3, 3, 3), Vector(0, 0, 0.5F), Vector(0, 0, 1), 50);
    // render the scene with progressive refinement
    auto sampler = DefaultSampler::NewSampler(16, 4);
    auto renderer = Renderer::NewRenderer(scene, camera, sampler, width, height, true);
    renderer.SamplesPerPixel = 16;
    renderer.IterativeRender("simplesphere.ppm", argc > 1 ? std::atoi(argv[1]) : 4);
    ptgpu_counters c = renderer.Counters();
    std::printf("%llu camera samples, %llu path segments, %llu shadow rays, last pass %.2f ms\n", (unsigned long long)c.cameraSamples,
                (unsigned long long)c.segments, (unsigned long long)c.shadowRays, c.lastPassMs);
    return 0;
}
